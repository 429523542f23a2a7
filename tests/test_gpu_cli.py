"""GPU test of the drop-in boundary: the reference's own command line, parser, orientation
generator and MSF / Chimera writers (unmodified objects from /root/reference, prebuilt into
oracle/_ref by oracle/build_ref.sh) with SearchOrCuda (minrms_b200/host) in the place of
SearchOrDyn3d.  Every file it writes must be byte-identical (timestamp line masked) to
what the unmodified reference binary writes -- compared against digests committed in
tests/golden/msf_digests.json and, for config 1, against the reference binary run here."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_msf_digests import CONFIGS, digest_dir, run  # noqa: E402

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "oracle", "_ref", "strict", "minrms_cuda")
REF = os.path.join(ROOT, "oracle", "_ref", "strict", "minrms")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("cfg", ["c1", "c2", "long", "bb", "alt", "alt_c2", "alt_bb", "alt_orig", "alt_bbnr", "alt_rall", "alt_rall_orig",
                                 "altr", "altr_bb", "altr_ir"])
def test_cli_output_matches_reference_digests(cfg, tmp_path):
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "msf_digests.json")))[cfg]
    got = run(EXE, CONFIGS[cfg], str(tmp_path), cfg)
    assert sorted(got) == sorted(want)
    diff = [f for f in want if got[f] != want[f]]
    assert not diff, "files differ from the reference: %s" % diff[:5]


@pytest.mark.skipif(not (os.path.exists(EXE) and os.path.exists(REF)), reason="oracle/_ref not built")
def test_cli_config1_side_by_side(tmp_path):
    a, b = tmp_path / "cuda", tmp_path / "ref"
    a.mkdir(); b.mkdir()
    got = run(EXE, CONFIGS["c1"], str(a))
    ref = run(REF, CONFIGS["c1"], str(b))
    assert got == ref and len(ref) == 96      # align47..141.msf + align_chimera.plot
    # SearchOrCuda::WriteChimeraInfo writes the plot from mr_plot_stats and the info file itself
    assert (a / "align_chimera.info").read_bytes() == (b / "align_chimera.info").read_bytes()
    assert (a / "align_chimera.plot").read_bytes() == (b / "align_chimera.plot").read_bytes()


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref not built")
def test_cli_two_gpu_slices_give_identical_files(tmp_path):
    """MINRMS_CUDA_GPUS=2 partitions Search(og,start,stop); with one physical GPU both
    contexts land on device 0 only if two devices exist, so this runs when >= 2 GPUs"""
    out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
    if out.count("GPU ") < 2:
        pytest.skip("needs 2 GPUs")
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "msf_digests.json")))["c1"]
    env = dict(os.environ, MINRMS_CUDA_GPUS="2")
    for f in ("hemoglobin.pdb", "myoglobin.pdb"):
        subprocess.check_call(["cp", os.path.join(ROOT, "tests", "golden", f), str(tmp_path)])
    r = subprocess.run([EXE] + CONFIGS["c1"] + ["hemoglobin.pdb,*.A", "myoglobin.pdb"], cwd=str(tmp_path),
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-500:]
    assert digest_dir(str(tmp_path)) == want


BATCH = os.path.join(ROOT, "oracle", "_ref", "strict", "minrms_batch_cuda")


@pytest.mark.skipif(not (os.path.exists(BATCH) and os.path.exists(REF)), reason="oracle/_ref not built")
def test_batch_cli_matches_pairwise_reference_runs(tmp_path):
    """the C++ all-vs-all driver (minrms_b200/host/batch_cuda.cc) on four small structures, two
    slots on the GPU: every pair directory holds the files of `minrms [flags] A B`"""
    from test_host_alternates_cpu import check_batch_against_pairwise_reference
    n = check_batch_against_pairwise_reference(BATCH, REF, str(tmp_path), ["-batch-workers-per-gpu", "2"])
    assert n > 100
