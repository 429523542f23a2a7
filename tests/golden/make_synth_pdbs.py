#!/usr/bin/env python3
"""CA-only PDB files of a synthetic mid-size pair (230 and 250 residues, seeded random walk
+ a rigidly moved noisy copy with indels; tests/helpers.py) so that the UNMODIFIED reference
can be run on a problem that needs the multi-pass DP kernel.  Outputs are committed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import random_chain, related_chain  # noqa: E402


def write_pdb(path, pts):
    with open(path, "w") as f:
        for k, (x, y, z) in enumerate(pts, 1):
            f.write("ATOM  %5d  %-3s %3s %1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n"
                    % (k, "CA", "GLY", "A", k, x, y, z, 1.00, 0.00))
        f.write("END\n")


def write_backbone_pdb(path, pts, names=("N", "CA", "C", "O")):
    """four backbone atoms per residue (pts: [residues * 4, 3])"""
    with open(path, "w") as f:
        for k, (x, y, z) in enumerate(pts):
            res, a = k // len(names) + 1, names[k % len(names)]
            f.write("ATOM  %5d  %-3s %3s %1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n"
                    % (k + 1, a, "GLY", "A", res, x, y, z, 1.00, 0.00))
        f.write("END\n")


def main():
    # a short pair with the four backbone atoms of share/README.md's atoms_used.txt example
    rng = np.random.default_rng(6066)
    A = random_chain(rng, 60, 4)
    B = related_chain(rng, A, 66, 4, noise=0.6)
    write_backbone_pdb(os.path.join(HERE, "synth_bb1.pdb"), A)
    write_backbone_pdb(os.path.join(HERE, "synth_bb2.pdb"), B)
    with open(os.path.join(HERE, "atoms_used_backbone.txt"), "w") as f:
        f.write(" N\n C\n CA\n O\n")
    print("wrote synth_bb1.pdb (60 x 4 atoms), synth_bb2.pdb (66 x 4 atoms), atoms_used_backbone.txt")
    rng = np.random.default_rng(230250)
    A = random_chain(rng, 230)
    B = related_chain(rng, A, 250, noise=0.9)
    write_pdb(os.path.join(HERE, "synth_a.pdb"), A)
    write_pdb(os.path.join(HERE, "synth_b.pdb"), B)
    print("wrote synth_a.pdb (230 CA), synth_b.pdb (250 CA)")
    # a longer pair (molecule 2 beyond 256 residues: CTA-per-orientation NW filter)
    rng = np.random.default_rng(280300)
    A = random_chain(rng, 280)
    B = related_chain(rng, A, 300, noise=0.8)
    write_pdb(os.path.join(HERE, "synth_c.pdb"), A)
    write_pdb(os.path.join(HERE, "synth_d.pdb"), B)
    print("wrote synth_c.pdb (280 CA), synth_d.pdb (300 CA)")
    # two short chains for the all-vs-all command line (with synth_bb1 / synth_bb2: six pairs)
    rng = np.random.default_rng(50505)
    A = random_chain(rng, 52, 1)
    write_pdb(os.path.join(HERE, "synth_e.pdb"), np.round(A, 3))
    B = related_chain(rng, A, 47, 1, noise=1.0)
    write_pdb(os.path.join(HERE, "synth_f.pdb"), np.round(B, 3))
    print("wrote synth_e.pdb (52 CA), synth_f.pdb (47 CA)")


if __name__ == "__main__":
    main()
