"""CPU check of the HOST side of the drop-in: minrms_b200/host/{main_cuda,search_or_cuda}.cc
compiled against the reference's own objects (oracle/_ref/strict/*.a, prebuilt by
oracle/build_ref.sh) with the C ABI answered by the CPU oracle (tests/host_shim/mr_oracle_shim.c)
in the place of libminrms_b200.so.  What is checked is the logic that lives above the C ABI --
table merging, the alternate-solutions bookkeeping of -alt-o / -alt-o-orig
(bin/minrms/search_or_dyn3d.cc:477-542, 626-952, 1253-1447 restated without temp files), the
orientation subset and the per-rank searches of -alt-r (:367-416, 960-1035), -r-all, the
per-rank writers -- against the digests of the files the UNMODIFIED reference binary writes
(tests/golden/msf_digests.json).  The GPU suite runs the same configurations through the CUDA
library (tests/test_gpu_cli.py).  Needs the reference's headers, i.e. /root/reference."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_msf_digests import CONFIGS, run  # noqa: E402

REF = os.environ.get("REF", "/root/reference")
STRICT = os.path.join(ROOT, "oracle", "_ref", "strict")
LIBDIRS = "global_utils parse_utils vect_3d eigen pdb++ fast_rot_metric biopolymer dyn superimpose or_gen pair_alignment"

pytestmark = pytest.mark.skipif(
    not (os.path.isdir(os.path.join(REF, "lib", "dyn")) and os.path.exists(os.path.join(STRICT, "libminrms_ref.a"))),
    reason="needs the reference's headers and oracle/_ref (oracle/build_ref.sh)")


@pytest.fixture(scope="module")
def hostcheck(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("hostcheck"))
    inc = ["-I%s/lib/%s" % (REF, x) for x in LIBDIRS.split()] + ["-I%s/bin/minrms" % REF, "-I" + os.path.join(ROOT, "include")]
    strict = ["-O2", "-ffp-contract=off"]      # the flags of oracle/_ref/strict and of the oracle
    for src, obj in ((os.path.join(ROOT, "tests", "host_shim", "mr_oracle_shim.c"), "shim.o"),
                     (os.path.join(ROOT, "oracle", "minrms_oracle.c"), "oracle.o")):
        subprocess.check_call(["gcc"] + strict + ["-c", "-I" + os.path.join(ROOT, "include"),
                                                  "-I" + os.path.join(ROOT, "oracle"), src, "-o", os.path.join(d, obj)])
    exe = os.path.join(d, "minrms_hostcheck")
    host = os.path.join(ROOT, "minrms_b200", "host")
    subprocess.check_call(["g++", "-std=c++11", "-DNDEBUG", "-w"] + strict + inc +
                          [os.path.join(host, "main_cuda.cc"), os.path.join(host, "search_or_cuda.cc"),
                           os.path.join(d, "shim.o"), os.path.join(d, "oracle.o"),
                           os.path.join(STRICT, "libminrms_bin.a"), os.path.join(STRICT, "libminrms_ref.a"),
                           "-lpthread", "-lm", "-o", exe])
    return exe


@pytest.mark.parametrize("cfg", ["bb", "alt_bb", "alt_bbnr", "alt", "alt_rall", "alt_rall_orig", "altr_bb", "altr_ir"])
def test_host_logic_writes_the_reference_files(cfg, hostcheck, tmp_path):
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "msf_digests.json")))[cfg]
    got = run(hostcheck, CONFIGS[cfg], str(tmp_path), cfg)
    assert sorted(got) == sorted(want)
    diff = [f for f in want if got[f] != want[f]]
    assert not diff, "files differ from the reference: %s" % diff[:5]


def test_two_slices_give_the_same_alternates(hostcheck, tmp_path):
    """MINRMS_CUDA_GPUS=2: the scores of the two slices are concatenated in search order, so the
    candidate records and the alternates are those of one search"""
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "msf_digests.json")))["alt_bbnr"]
    os.environ["MINRMS_CUDA_GPUS"] = "2"
    try:
        got = run(hostcheck, CONFIGS["alt_bbnr"], str(tmp_path), "alt_bbnr")
    finally:
        del os.environ["MINRMS_CUDA_GPUS"]
    assert got == want


def test_two_slices_give_the_same_pairs_alternates(hostcheck, tmp_path):
    """-alt-r with MINRMS_CUDA_GPUS=2: every rank's search is cut in two and merged"""
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "msf_digests.json")))["altr_ir"]
    os.environ["MINRMS_CUDA_GPUS"] = "2"
    try:
        got = run(hostcheck, CONFIGS["altr_ir"], str(tmp_path), "altr_ir")
    finally:
        del os.environ["MINRMS_CUDA_GPUS"]
    assert got == want
