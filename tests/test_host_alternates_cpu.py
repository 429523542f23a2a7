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
    # the all-vs-all driver over the same host sources
    subprocess.check_call(["g++", "-std=c++11", "-DNDEBUG", "-w"] + strict + inc +
                          [os.path.join(host, "batch_cuda.cc"), os.path.join(host, "search_or_cuda.cc"),
                           os.path.join(d, "shim.o"), os.path.join(d, "oracle.o"),
                           os.path.join(STRICT, "libminrms_bin.a"), os.path.join(STRICT, "libminrms_ref.a"),
                           "-lpthread", "-lm", "-o", os.path.join(d, "minrms_batch_hostcheck")])
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


BATCH_STRUCTURES = ["synth_e.pdb", "synth_f.pdb", "synth_bb1.pdb", "synth_bb2.pdb"]
BATCH_FLAGS = ["-fm", "4", "-HS", "-ir", "-r", "-minN", "0.4"]


def check_batch_against_pairwise_reference(batch_exe, ref_exe, workdir, extra=()):
    """every pair directory of the batch driver holds the files `minrms [flags] A B` writes"""
    from make_msf_digests import digest_dir
    for f in BATCH_STRUCTURES:
        subprocess.check_call(["cp", os.path.join(ROOT, "tests", "golden", f), workdir])
    with open(os.path.join(workdir, "list.txt"), "w") as f:
        f.write("# structures\n" + "\n".join(BATCH_STRUCTURES) + "\n")
    r = subprocess.run([batch_exe] + BATCH_FLAGS + ["-batch", "list.txt", "-batch-out", "out"] + list(extra),
                       cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-300:])
    assert "6 pairs in" in r.stdout
    n_files = 0
    for i, a in enumerate(BATCH_STRUCTURES):
        for b in BATCH_STRUCTURES[i + 1:]:
            got = digest_dir(os.path.join(workdir, "out", a[:-4] + "__" + b[:-4]))
            refdir = os.path.join(workdir, "ref_" + a[:-4] + "_" + b[:-4])
            os.mkdir(refdir)
            subprocess.check_call(["cp", os.path.join(workdir, a), os.path.join(workdir, b), refdir])
            subprocess.check_call([ref_exe] + BATCH_FLAGS + [a, b], cwd=refdir,
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            want = digest_dir(refdir)
            assert got == want and len(want) > 10, (a, b)
            n_files += len(want)
    return n_files


@pytest.mark.parametrize("workers", [1, 3])
def test_batch_driver_writes_what_pairwise_runs_write(workers, hostcheck, tmp_path):
    """minrms_b200/host/batch_cuda.cc (all-vs-all: worker processes that run pair after pair, the
    parser's static atom table emptied in between) against one run of the unmodified reference
    binary per pair; with one worker all six pairs go through one process"""
    batch = os.path.join(os.path.dirname(hostcheck), "minrms_batch_hostcheck")
    n = check_batch_against_pairwise_reference(batch, os.path.join(STRICT, "minrms"), str(tmp_path),
                                               ["-batch-workers-per-gpu", str(workers)])
    assert n > 100


def test_batch_driver_survives_a_pair_the_reference_refuses(hostcheck, tmp_path):
    """a structure too short for the fragment size makes the reference end the process (ERR):
    the pairs with it are reported as failed, every other pair of the worker's share still runs"""
    from make_msf_digests import digest_dir
    work = str(tmp_path)
    for f in BATCH_STRUCTURES[:3]:
        subprocess.check_call(["cp", os.path.join(ROOT, "tests", "golden", f), work])
    with open(os.path.join(work, "tiny.pdb"), "w") as f:      # two residues: no fragment of four
        for k in range(2):
            f.write("ATOM  %5d  %-3s %3s %1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n" % (k + 1, "CA", "GLY", "A", k + 1, 3.8 * k, 0.0, 0.0, 1.0, 0.0))
        f.write("END\n")
    with open(os.path.join(work, "list.txt"), "w") as f:
        f.write("\n".join(BATCH_STRUCTURES[:2] + ["tiny.pdb", BATCH_STRUCTURES[2]]) + "\n")
    batch = os.path.join(os.path.dirname(hostcheck), "minrms_batch_hostcheck")
    r = subprocess.run([batch] + BATCH_FLAGS + ["-batch", "list.txt", "-batch-out", "out"], cwd=work,
                       capture_output=True, text=True)
    assert r.returncode == 1 and "3 pair(s) FAILED" in r.stdout, (r.stdout[-300:], r.stderr[-300:])
    good = [d for d in os.listdir(os.path.join(work, "out")) if "tiny" not in d]
    assert len(good) == 3
    for d in good:
        assert os.path.exists(os.path.join(work, "out", d, "minrms.done"))
        assert len(digest_dir(os.path.join(work, "out", d))) > 10
