#!/usr/bin/env python3
"""Golden values of the align_chimera.plot statistics from the UNMODIFIED reference.

Needs /root/reference and oracle/_ref (oracle/build_ref.sh); run in the build container,
output committed as tests/golden/ref_plot.npz.

For each config below, the reference's best tables stored in ref_<cfg>.npz (strict build,
after FirstPass and after Refine) are handed to oracle/_ref/strict/ref_plot, which evaluates
FindFurthestResPair and LevittGerstein98::Sstr_Probability (lib/pair_alignment/
pair_alignment.cc:2351-2379, 2690-2925) and the printed expressions of
SearchOrDyn3d::WritePlotFile (bin/minrms/search_or_dyn3d.cc:2703-2760) with the reference's
own code.  Stored per config and stage: stats[R][9] (float64) =
{has, rmsd, furthest, S_str, mu_str, delta_str, Z, probability, -log10(probability)}.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden import CONFIGS  # noqa: E402

TMP = os.environ.get("GOLDEN_TMP", "/tmp/plot_dumps")
PICK = ["c1", "c2", "mid", "bb", "maxn"]


def write_table(path, ssd, T, pairs):
    R, N = pairs.shape[0], pairs.shape[1]
    with open(path, "wb") as f:
        f.write(np.array([R, N], np.int32).tobytes())
        for r in range(R):
            f.write(np.float32(ssd[r]).tobytes())
            f.write(np.ascontiguousarray(T[r], np.float32).tobytes())
            f.write(np.ascontiguousarray(pairs[r], np.int16).tobytes())


def main():
    os.makedirs(TMP, exist_ok=True)
    out = {}
    exe = os.path.join(ROOT, "oracle", "_ref", "strict", "ref_plot")
    for cfg in PICK:
        stride, nsample, args, mols, fm_cutoff = CONFIGS[cfg]
        g = np.load(os.path.join(HERE, "ref_%s.npz" % cfg))
        for f in mols:
            subprocess.check_call(["cp", os.path.join(HERE, f.split(",")[0]), TMP])
        au = os.path.join(TMP, "atoms_used.txt")       # read by the reference from its cwd
        if cfg.startswith("bb"):
            subprocess.check_call(["cp", os.path.join(HERE, "atoms_used_backbone.txt"), au])
        elif os.path.exists(au):
            os.remove(au)
        for stage in ("search", "refine"):
            tin = os.path.join(TMP, "%s_%s.in" % (cfg, stage))
            tout = os.path.join(TMP, "%s_%s.out" % (cfg, stage))
            write_table(tin, g["strict_%s_ssd" % stage], g["strict_%s_T" % stage], g["strict_%s_pairs" % stage])
            subprocess.check_call([exe, tin, tout] + args + list(mols), cwd=TMP, stdout=subprocess.DEVNULL)
            raw = open(tout, "rb").read()
            R = g["strict_%s_ssd" % stage].shape[0]
            stats = np.frombuffer(raw[: R * 72], np.float64).reshape(R, 9)
            sizes = np.frombuffer(raw[R * 72:], np.int32)
            out["%s_%s" % (cfg, stage)] = stats
            out["%s_sizes" % cfg] = sizes
            print(cfg, stage, "sizes with a solution:", int(stats[:, 0].sum()), "sizeof(rmsd, -log10):", sizes)
    path = os.path.join(HERE, "ref_plot.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
