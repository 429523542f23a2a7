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
}


def digest_dir(d):
    out = {}
    for f in sorted(glob.glob(os.path.join(d, "align*.msf")) + glob.glob(os.path.join(d, "align_chimera.plot"))):
        lines = open(f, "rb").read().split(b"\n")
        lines = [b"<MSF header line masked>" if b" MSF: " in ln and b"Check:" in ln else ln for ln in lines]
        out[os.path.basename(f)] = hashlib.sha256(b"\n".join(lines)).hexdigest()
    return out


def run(exe, args, workdir):
    for f in ("hemoglobin.pdb", "myoglobin.pdb"):
        subprocess.check_call(["cp", os.path.join(HERE, f), workdir])
    subprocess.check_call([exe] + args + ["hemoglobin.pdb,*.A", "myoglobin.pdb"], cwd=workdir,
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
                res[cfg] = run(exe, args, d)
        print(cfg, len(res[cfg]), "files")
    json.dump(res, open(os.path.join(HERE, "msf_digests.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    sys.exit(main())
