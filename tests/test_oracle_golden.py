"""CPU tests: the oracle (oracle/minrms_oracle.c) against the golden fixtures generated
from the unmodified reference (tests/golden/make_golden.py, strict build)."""
import os

import numpy as np
import pytest

import oracle_ctypes as oc
from helpers import (assert_eval_equal, assert_table_equal, bits, params_from_golden, sha)


def _orientations(g):
    return oc.gen_orientations(g["A"], g["B"], int(g["nb"]), int(g["frag"]), int(g["n_min"]),
                               cutoff_rmsd=float(g["fm_cutoff"]) if "fm_cutoff" in g else -1.0,
                               intervals=g["intervals"] if "intervals" in g else None)


@pytest.mark.parametrize("cfg", ["golden_c1", "golden_c2", "golden_mid", "golden_long", "golden_bb", "golden_bb_of", "golden_ir_off", "golden_iv", "golden_maxn"])
def test_orientation_list_matches_reference(cfg, request):
    g = request.getfixturevalue(cfg)
    og = _orientations(g)
    assert og.shape[0] == int(g["count"])
    assert np.array_equal(bits(og[:8]), bits(g["orient_head"]))
    assert np.array_equal(sha(og), g["orient_sha256"])          # includes the shuffle order
    assert np.array_equal(og[g["sample_ids"]], g["sample_orient"])


@pytest.mark.parametrize("cfg", ["golden_c1", "golden_c2", "golden_mid", "golden_long", "golden_bb", "golden_bb_of", "golden_ir_off", "golden_iv", "golden_maxn"])
def test_per_orientation_values_match_reference(cfg, request):
    g = request.getfixturevalue(cfg)
    ws = oc.Workspace(g["A"], g["B"], params_from_golden(g))
    for k, o in enumerate(g["sample_ids"]):
        res = ws.eval(g["sample_orient"][k])
        assert (res["n_last"] > 0) == bool(g["sample_good"][k])
        assert_eval_equal(res, g["sample_n_last"][k], g["sample_cost"][k], g["sample_valid"][k],
                          g["sample_ssd"][k], g["sample_T"][k], g["sample_pairs"][k], "o=%d" % o)


def test_config1_search_and_refine_bit_exact(golden_c1):
    """full config 1 (tests/test_typical_usage.sh flags): FirstPass then Refine"""
    g = golden_c1
    ws = oc.Workspace(g["A"], g["B"], params_from_golden(g))
    og = oc.gen_orientations(g["A"], g["B"], int(g["nb"]), int(g["frag"]), int(g["n_min"]))
    best = ws.search(og)
    assert_table_equal(best["ssd"], best["T"], best["pairs"], g["strict_search_ssd"],
                       g["strict_search_T"], g["strict_search_pairs"], int(g["n_min"]), "search")
    ws.refine(best)
    assert_table_equal(best["ssd"], best["T"], best["pairs"], g["strict_refine_ssd"],
                       g["strict_refine_T"], g["strict_refine_pairs"], int(g["n_min"]), "refine")


@pytest.mark.parametrize("cfg", ["golden_mid", "golden_long", "golden_bb", "golden_bb_of", "golden_ir_off", "golden_iv", "golden_maxn"])
def test_synthetic_search_and_refine_bit_exact(cfg, request):
    """synthetic pairs run through the unmodified reference: 230 x 250 unconstrained and
    280 x 300 under the flags of config 1 (fragment RMSD cut-off -fm 4,0.6), and a 60 x 66 pair
    with four backbone atoms per residue, unconstrained and filtered"""
    g = request.getfixturevalue(cfg)
    ws = oc.Workspace(g["A"], g["B"], params_from_golden(g))
    og = _orientations(g)
    assert og.shape[0] == int(g["count"])
    best = ws.search(og)
    assert_table_equal(best["ssd"], best["T"], best["pairs"], g["strict_search_ssd"],
                       g["strict_search_T"], g["strict_search_pairs"], int(g["n_min"]), "search")
    ws.refine(best)
    assert_table_equal(best["ssd"], best["T"], best["pairs"], g["strict_refine_ssd"],
                       g["strict_refine_T"], g["strict_refine_pairs"], int(g["n_min"]), "refine")


@pytest.mark.skipif(not os.environ.get("ORACLE_FULL"), reason="200 s of CPU; set ORACLE_FULL=1")
def test_config2_search_and_refine_bit_exact(golden_c2):
    g = golden_c2
    ws = oc.Workspace(g["A"], g["B"], params_from_golden(g))
    og = oc.gen_orientations(g["A"], g["B"], int(g["nb"]), int(g["frag"]), int(g["n_min"]))
    best = ws.search(og)
    assert_table_equal(best["ssd"], best["T"], best["pairs"], g["strict_search_ssd"],
                       g["strict_search_T"], g["strict_search_pairs"], int(g["n_min"]), "search")
    ws.refine(best)
    assert_table_equal(best["ssd"], best["T"], best["pairs"], g["strict_refine_ssd"],
                       g["strict_refine_T"], g["strict_refine_pairs"], int(g["n_min"]), "refine")


@pytest.mark.parametrize("cfg", ["golden_c1", "golden_c2"])
def test_shipped_flags_reference_within_tie_tolerance(cfg, request):
    """Oracle-F (the reference with its shipped -O3 -ffast-math) vs Oracle-S (strict):
    after FirstPass the alignments must be identical wherever the scores differ by more
    than the tie tolerance tau = 4e-4 relative on SSD, and RMSD within 1.5e-4 A
    (the reference's own float32 noise floor is 1.0e-4..1.24e-4 A)."""
    g = request.getfixturevalue(cfg)
    n_min = int(g["n_min"])
    s, f = g["strict_search_ssd"], g["fast_search_ssd"]
    have = s < 1e26
    assert np.array_equal(have, f < 1e26)
    n = n_min + np.nonzero(have)[0]
    rs, rf = np.sqrt(s[have] / n), np.sqrt(f[have] / n)
    # E0 - 2*lambda cancels catastrophically, so for near-zero RMSD (n = 2, 3 of config 2:
    # 0.0016 A) the two reference builds disagree by 6.6e-4 A; state the bound for RMSD >= 0.01 A
    big = rs >= 0.01
    assert np.max(np.abs(rs - rf)[big]) < 1.5e-4
    assert np.max(np.abs(s[have] - f[have])) < 0.1        # absolute SSD noise (float32 ulp of E0 ~ 1e5 is 8e-3)
    same = [np.array_equal(g["strict_search_pairs"][r][:n_min + r], g["fast_search_pairs"][r][:n_min + r])
            for r in np.nonzero(have)[0]]
    rel = np.abs(s[have] - f[have]) / np.maximum(s[have], 1e-30)
    # any alignment difference must be explained by a score gap below the tie tolerance
    for ok, gap in zip(same, rel):
        assert ok or gap < 4e-4
    assert all(same)   # measured: FirstPass alignments are identical on both configs


def test_known_rmsd_values(golden_c1, golden_c2):
    """the RMSD values the reference prints (SURVEY.md 8c golden values, shipped flags)"""
    def rmsd(g, n):
        return float(np.sqrt(g["fast_refine_ssd"][n - int(g["n_min"])] / np.float32(n)))
    assert abs(rmsd(golden_c1, 47) - 0.475068) < 2e-6
    assert abs(rmsd(golden_c1, 100) - 0.923256) < 2e-6
    assert abs(rmsd(golden_c1, 141) - 1.51285) < 1e-5
    assert abs(rmsd(golden_c2, 10) - 0.105252) < 2e-6
    assert abs(rmsd(golden_c2, 47) - 0.474806) < 2e-6
    assert abs(rmsd(golden_c2, 141) - 1.51285) < 1e-5


def test_superimpose_properties():
    """Kabsch/Diamond: recovers a known rigid motion; rmsd 0 for identical sets"""
    rng = np.random.default_rng(7)
    P = rng.normal(scale=10, size=(40, 3)).astype(np.float32)
    r, T = oc.superimpose(P, P)
    assert r < 1e-3
    th = 0.7
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    Q = (P @ Rz.T + np.array([3.0, -2.0, 5.0])).astype(np.float32)
    r, T = oc.superimpose(Q, P)            # move P onto Q
    assert r < 2e-3
    assert np.allclose(T.reshape(3, 4)[:, :3], Rz, atol=1e-4)
    assert np.allclose(oc.apply_transform(T, P), Q, atol=2e-3)


def test_dp_cell_counts_closed_form():
    assert oc.dp_cells(141, 151, 1, 141) == 1044481
    assert oc.dp_cells(141, 151, 47, 141) == 794770      # C1 (banded)
    assert oc.dp_cells(600, 600, 1, 600) == 72180100     # C3


def test_edge_cases_small():
    """tiny and ragged inputs: 1 x k, k x 1, n_min = n_max, -maxN below min(m,n)"""
    rng = np.random.default_rng(3)
    for (m, nB, n_min, n_max) in [(3, 3, 1, 3), (2, 7, 1, 2), (7, 2, 2, 2), (5, 9, 3, 4), (1, 4, 1, 1)]:
        A = rng.normal(scale=5, size=(m, 3)).astype(np.float32)
        B = rng.normal(scale=5, size=(nB, 3)).astype(np.float32)
        p = oc.make_params(n_min=n_min, n_max=n_max, m=m, nB=nB)
        ws = oc.Workspace(A, B, p)
        T = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
        res = ws.eval(T)
        assert res["n_last"] == max(1, n_max) if n_max >= 2 else res["n_last"] == 1
        for n in range(max(2, n_min), n_max + 1):
            assert res["valid"][n - 1]
            pr = res["pairs"][n - 1][:n]
            assert np.all(np.diff(pr[:, 0]) > 0) and np.all(np.diff(pr[:, 1]) > 0)
        # costs are non-decreasing in n at a fixed orientation
        assert np.all(np.diff(res["cost"][:res["n_last"]]) >= 0)
