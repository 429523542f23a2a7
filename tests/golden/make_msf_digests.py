#!/usr/bin/env python3
"""Digest the files the UNMODIFIED reference CLI (oracle/_ref/strict/minrms, strict build)
writes for configs 1 and 2 on the fixture pair: align<n>.msf (timestamp line masked,
lib/pair_alignment/pair_alignment.cc:1553-1556) and align_chimera.plot.
Needs /root/reference (for the build); outputs tests/golden/msf_digests.json (committed)."""
import glob
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CONFIGS = {
    "c1": ["-fm", "4", "-of", "8.0,0.33", "-HS", "-minN", "0.33", "-max-rmsd", "3.5", "-ir", "-r"],
    "c2": ["-fm", "4", "-HS", "-ir", "-r"],
    # synthetic pairs of make_synth_pdbs.py: 280 x 300 under config-1 flags with a fragment RMSD
    # cut-off (multi-pass DP, CTA NW filter), and 60 x 66 with four backbone atoms per residue
    "long": ["-fm", "4,0.6", "-of", "8.0,0.33", "-HS", "-minN", "0.33", "-max-rmsd", "4.0", "-ir", "-r"],
    "bb": ["-fm", "4", "-HS", "-ir", "-r"],
    # alternate solutions (bin/minrms/minrms_parser.cc:755-840): three ranks on config 1 with an
    # RMSD tolerance; four ranks, no tolerance, four backbone atoms; -alt-o-orig (implies -IR);
    # no refinement with a distance cut-off
    "alt": ["-fm", "4", "-of", "8.0,0.33", "-HS", "-minN", "0.33", "-max-rmsd", "3.5", "-ir", "-r",
            "-alt-o", "3,2.0,1.5"],
    "alt_bb": ["-fm", "4", "-HS", "-ir", "-r", "-alt-o", "4,1.0"],
    "alt_orig": ["-fm", "4", "-of", "8.0,0.33", "-HS", "-minN", "0.33", "-max-rmsd", "3.5", "-r",
                 "-alt-o-orig", "3,1.5,1.6"],
    "alt_bbnr": ["-fm", "4", "-HS", "-ir", "-alt-o", "3,0.5,1.3", "-max-dist", "6.0"],
    # config 2 (no cut-offs: every size of all 20 424 orientations is a candidate record)
    "alt_c2": ["-fm", "4", "-HS", "-ir", "-r", "-alt-o", "3,3.0,1.3"],
    # -r-all: the alternates are refined too (bin/minrms/search_or_dyn3d.cc:2344-2640 with r > 0)
    "alt_rall": ["-fm", "4", "-HS", "-ir", "-r-all", "-alt-o", "3,1.0,1.6"],
    "alt_rall_orig": ["-fm", "4", "-HS", "-r-all", "-alt-o-orig", "3,1.0,1.6", "-max-dist", "7.0"],
    # -alt-r (ALT_METHOD_PAIRS): one more search per rank, at most K common pairs with the lower ranks
    "altr": ["-fm", "4", "-of", "8.0,0.33", "-HS", "-minN", "0.33", "-max-rmsd", "3.5", "-ir", "-r",
             "-alt-r", "3,40,1.5"],
    "altr_bb": ["-fm", "4", "-HS", "-ir", "-r", "-alt-r", "3,5"],
    "altr_ir": ["-fm", "4", "-HS", "-IR", "-r-all", "-alt-r", "4,20,1.4", "-max-dist", "7.0"],
}
MOLS = {"long": ("synth_c.pdb", "synth_d.pdb"), "bb": ("synth_bb1.pdb", "synth_bb2.pdb"),
        "alt_bb": ("synth_bb1.pdb", "synth_bb2.pdb"), "alt_bbnr": ("synth_bb1.pdb", "synth_bb2.pdb"),
        "alt_rall": ("synth_bb1.pdb", "synth_bb2.pdb"), "alt_rall_orig": ("synth_bb1.pdb", "synth_bb2.pdb"),
        "altr_bb": ("synth_bb1.pdb", "synth_bb2.pdb"), "altr_ir": ("synth_bb1.pdb", "synth_bb2.pdb")}
PAIR = ("hemoglobin.pdb,*.A", "myoglobin.pdb")


def digest_dir(d):
    out = {}
    # with alternates the files of rank r are align<r>_<n>.msf, align<r>_chimera.plot / .info
    # (bin/minrms/search_or_dyn3d.cc:2682-2688, 2968-2990)
    for f in sorted(glob.glob(os.path.join(d, "align*.msf")) + glob.glob(os.path.join(d, "align_chimera.plot")) +
                    glob.glob(os.path.join(d, "align[0-9]*_chimera.*"))):
        lines = open(f, "rb").read().split(b"\n")
        lines = [b"<MSF header line masked>" if b" MSF: " in ln and b"Check:" in ln else ln for ln in lines]
        out[os.path.basename(f)] = hashlib.sha256(b"\n".join(lines)).hexdigest()
    return out


def run(exe, args, workdir, cfg=None):
    mols = MOLS.get(cfg, PAIR)
    for f in mols:
        subprocess.check_call(["cp", os.path.join(HERE, f.split(",")[0]), workdir])
    if cfg in ("bb", "alt_bb", "alt_rall", "altr_bb"):   # the reference reads the atom names from atoms_used.txt in its cwd
        subprocess.check_call(["cp", os.path.join(HERE, "atoms_used_backbone.txt"), os.path.join(workdir, "atoms_used.txt")])
    subprocess.check_call([exe] + args + list(mols), cwd=workdir,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return digest_dir(workdir)


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "strict", "minrms")
    res = {}
    for cfg, args in CONFIGS.items():
        pre = os.path.join("/tmp/ref_runs", cfg + "_strict_r")
        if os.path.isdir(pre) and glob.glob(os.path.join(pre, "align*.msf")):
            res[cfg] = digest_dir(pre)
        else:
            with tempfile.TemporaryDirectory() as d:
                res[cfg] = run(exe, args, d, cfg)
        print(cfg, len(res[cfg]), "files")
    json.dump(res, open(os.path.join(HERE, "msf_digests.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    sys.exit(main())
