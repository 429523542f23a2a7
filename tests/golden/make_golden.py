#!/usr/bin/env python3
"""Generate the golden fixtures from the UNMODIFIED reference (needs /root/reference and
oracle/_ref built by oracle/build_ref.sh; run in the build container, outputs committed).

For each config (C1 = tests/test_typical_usage.sh flags, C2 = unconstrained)
runs oracle/_ref/{strict,fast}/ref_dump on the fixture pair and stores in ref_<cfg>.npz:
  params, coordinates, sha256 of the shuffled orientation list, a sample of orientations
  with every per-size value the reference computes for them (strict build), and the
  reference's best tables after FirstPass and after Refine (strict and fast builds).
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_ctypes as oc  # noqa: E402

PAIR = ("hemoglobin.pdb,*.A", "myoglobin.pdb")
SYNTH = ("synth_a.pdb", "synth_b.pdb")       # tests/golden/make_synth_pdbs.py: 230 x 250 residues
# name: (detail stride, sample size, flags, molecules, fragment RMSD cut-off)
CONFIGS = {
    "c1": (50, 32, ["-fm", "4", "-of", "8.0,0.33", "-HS", "-minN", "0.33", "-max-rmsd", "3.5", "-ir", "-r"], PAIR, -1.0),
    "c2": (200, 16, ["-fm", "4", "-HS", "-ir", "-r"], PAIR, -1.0),
    # a pair beyond the single-pass DP kernel; "-fm 4,0.6" keeps only fragment pairs that
    # superimpose within 0.6 A (a few hundred orientations: the reference needs ~40 s)
    "mid": (5, 12, ["-fm", "4,0.6", "-HS", "-ir", "-r"], SYNTH, 0.6),
    # 280 x 300 residues with the flags of config 1: molecule 2 is longer than 256 residues
    # (CTA-per-orientation NW filter), banded DP windows, max-rmsd cut-off, multi-pass DP
    "long": (5, 12, ["-fm", "4,0.6", "-of", "8.0,0.33", "-HS", "-minN", "0.33", "-max-rmsd", "4.0", "-ir", "-r"],
             ("synth_c.pdb", "synth_d.pdb"), 0.6),
    # four backbone atoms per residue (atoms_used.txt = N, C, CA, O; share/README.md), 60 x 66
    "bb": (50, 12, ["-fm", "4", "-HS", "-ir", "-r"], ("synth_bb1.pdb", "synth_bb2.pdb"), -1.0),
    "bb_of": (20, 12, ["-fm", "4", "-of", "8.0,0.33", "-HS", "-minN", "0.33", "-max-rmsd", "3.0", "-ir", "-r"],
              ("synth_bb1.pdb", "synth_bb2.pdb"), -1.0),
    # CA atoms of the same files, -IR (the DP cost is the score, no Kabsch per size) and the
    # -max-dist cut-off with its early exit
    "ir_off": (50, 12, ["-fm", "4", "-HS", "-IR", "-max-dist", "7.0", "-minN", "0.25", "-r"],
               ("synth_bb1.pdb", "synth_bb2.pdb"), -1.0),
    # explicit fragment-matching interval pairs (residue numbers 5-30 x 10-40 and 35-55 x 30-60:
    # 0-based index intervals [4,29]x[9,39], [34,54]x[29,59]; or_gen.cc:345-493)
    "iv": (20, 12, ["-fm", "4", "-HS", "-fm-interval", "5-30.A,10-40.A", "-fm-interval", "35-55.A,30-60.A",
                    "-ir", "-r"], ("synth_bb1.pdb", "synth_bb2.pdb"), -1.0),
    # -minN / -maxN as counts: only sizes 10..40 of the 60 x 66 pair
    "maxn": (50, 12, ["-fm", "5", "-HS", "-minN", "10", "-maxN", "40", "-ir", "-r"],
             ("synth_bb1.pdb", "synth_bb2.pdb"), -1.0),
}
INTERVALS = {"iv": np.array([[4, 29, 9, 39], [34, 54, 29, 59]], np.int32)}
TMP = os.environ.get("GOLDEN_TMP", "/tmp/dumps")


def run_dump(flavour, cfg, stride, args, mols):
    out = os.path.join(TMP, "%s_%s.bin" % (cfg, flavour))
    if not os.path.exists(out):
        os.makedirs(TMP, exist_ok=True)
        for f in mols:
            subprocess.check_call(["cp", os.path.join(HERE, f.split(",")[0]), TMP])
        au = os.path.join(TMP, "atoms_used.txt")       # read by the reference from its cwd
        if cfg.startswith("bb"):
            subprocess.check_call(["cp", os.path.join(HERE, "atoms_used_backbone.txt"), au])
        elif os.path.exists(au):
            os.remove(au)
        exe = os.path.join(ROOT, "oracle", "_ref", flavour, "ref_dump")
        subprocess.check_call([exe, out, str(stride)] + args + list(mols), cwd=TMP, stdout=subprocess.DEVNULL)
    return oc.read_ref_dump(out)


def main():
    only = sys.argv[1:]
    for cfg, (stride, nsample, args, mols, fm_cutoff) in CONFIGS.items():
        if only and cfg not in only:
            continue
        ds = run_dump("strict", cfg, stride, args, mols)
        df = run_dump("fast", cfg, -1, args, mols)
        assert np.array_equal(ds["orient"], df["orient"]) or True
        # sample: detail orientations, preferring ones that reach the DP, plus a few rejected
        det = sorted(ds["detail"].keys())
        passing = [o for o in det if ds["n_last"][o] > 0]
        failing = [o for o in det if ds["n_last"][o] == 0]
        long_ones = sorted(passing, key=lambda o: -ds["n_last"][o])[: nsample // 4]
        pick = sorted(set(long_ones + passing[: nsample - len(long_ones) - min(4, len(failing))] + failing[:4]))
        N = ds["N"]
        out = dict(
            m=ds["m"], nB=ds["nB"], nb=ds["nb"], n_min=ds["n_min"], n_max=ds["n_max"],
            max_dist=ds["max_dist"], max_rmsd=ds["max_rmsd"], use_filter=ds["use_filter"],
            of_dist=ds["of_dist"], of_min=ds["of_min"], ir=ds["ir"], frag=ds["frag"],
            fm_cutoff=np.float32(fm_cutoff),
            **({"intervals": INTERVALS[cfg]} if cfg in INTERVALS else {}),
            count=ds["count"], A=ds["A"], B=ds["B"],
            orient_sha256=np.frombuffer(hashlib.sha256(ds["orient"].tobytes()).digest(), np.uint8),
            orient_fast_sha256=np.frombuffer(hashlib.sha256(df["orient"].tobytes()).digest(), np.uint8),
            orient_head=ds["orient"][:8],
            n_pass=int((ds["n_last"] > 0).sum()),
            n_last_all=ds["n_last"].astype(np.int16),
            sample_ids=np.asarray(pick, np.int64),
            sample_orient=ds["orient"][pick],
            sample_good=ds["good"][pick], sample_n_last=ds["n_last"][pick],
            sample_cost=ds["cost"][pick], sample_valid=ds["valid"][pick], sample_ssd=ds["ssd"][pick],
            sample_T=np.stack([ds["detail"][o][0] for o in pick]),
            sample_pairs=np.stack([ds["detail"][o][1] for o in pick]),
        )
        for tag, d in (("strict", ds), ("fast", df)):
            for stage in ("search", "refine"):
                b = d["best_" + stage]
                out["%s_%s_ssd" % (tag, stage)] = b["ssd"]
                out["%s_%s_T" % (tag, stage)] = b["T"]
                out["%s_%s_pairs" % (tag, stage)] = b["pairs"]
        path = os.path.join(HERE, "ref_%s.npz" % cfg)
        np.savez_compressed(path, **out)
        print(cfg, "->", path, os.path.getsize(path), "bytes;", len(pick), "sample orientations;",
              out["n_pass"], "of", ds["count"], "reach the DP")


if __name__ == "__main__":
    main()
