"""Per-size statistics of align_chimera.plot (SearchOrDyn3d::WritePlotFile,
bin/minrms/search_or_dyn3d.cc:2703-2760; FindFurthestResPair and LevittGerstein98 in
lib/pair_alignment/pair_alignment.cc:2351-2379, 2690-2925).

tests/golden/ref_plot.npz holds what the UNMODIFIED reference (strict build) computes for its
own best tables (tests/golden/make_plot_golden.py).  CPU tests pin the oracle to it; the GPU
tests pin mr_plot_stats (C ABI) to it -- all nine columns bit for bit -- and the plot file the
Python mirror writes to the digest of the file the reference binary writes."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_ctypes as oc
from helpers import random_chain, related_chain, settings_from_golden, settings_from_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
# config -> index of " CA" among the atoms used per residue (atoms_used_backbone.txt: N, C, CA, O)
CONFIGS = {"c1": 0, "c2": 0, "mid": 0, "bb": 2, "maxn": 0}


def _load(cfg):
    return np.load(os.path.join(GOLD, "ref_%s.npz" % cfg)), np.load(os.path.join(GOLD, "ref_plot.npz"))


def _same(got, want, what):
    assert got.shape == want.shape, what
    bad = np.argwhere(got.view(np.uint64) != want.view(np.uint64))
    assert bad.size == 0, "%s: first difference at (slot, column) %s: %r != %r" % (
        what, bad[0], got[tuple(bad[0])], want[tuple(bad[0])])


@pytest.mark.parametrize("stage", ["search", "refine"])
@pytest.mark.parametrize("cfg", sorted(CONFIGS))
def test_oracle_plot_rows_match_reference(cfg, stage):
    g, p = _load(cfg)
    rows = oc.plot_stats(g["A"], g["B"], int(g["nb"]), CONFIGS[cfg], int(g["n_min"]),
                         g["strict_%s_ssd" % stage], g["strict_%s_T" % stage], g["strict_%s_pairs" % stage])
    _same(rows, p["%s_%s" % (cfg, stage)], "%s/%s" % (cfg, stage))
    assert rows[:, 0].sum() >= 31


def test_oracle_probability_closed_form():
    """both branches of Sstr_Probability (n < 120 / n >= 120, Z below / above ln 5000)"""
    import ctypes as C
    out = np.zeros(4, np.float32)
    fp = out.ctypes.data_as(C.POINTER(C.c_float))
    oc.lib().orc_sstr_probability(C.c_float(100.0), 50, fp)
    lnN = np.log(np.float32(50.0), dtype=np.float32).astype(np.float64)
    mu = 18.411424 * lnN * lnN - 4.501719 * lnN + 2.637163
    delta = 21.351857 * lnN - 37.521269
    assert abs(out[0] - mu) < 1e-3 and abs(out[1] - delta) < 1e-3
    z = (100.0 - mu) / delta
    assert abs(out[2] - z) < 1e-4 and abs(out[3] - (1.0 - np.exp(-np.exp(-z)))) < 1e-6
    oc.lib().orc_sstr_probability(C.c_float(2500.0), 141, fp)      # large Z: exp(-Z) branch
    assert out[2] > np.log(5000.0) and 0.0 < out[3] < 2.1e-4
    assert abs(out[1] - (21.351857 * np.log(120.0) - 37.521269)) < 1e-4


# ----------------------------------------------------------------------------------- GPU
def _search(g, cfg):
    from minrms_b200 import SearchOrCuda
    st = settings_from_golden(g)
    st.ca_index = CONFIGS[cfg]
    return SearchOrCuda(g["A"], g["B"], st)


@pytest.mark.gpu
@pytest.mark.parametrize("stage", ["search", "refine"])
@pytest.mark.parametrize("cfg", sorted(CONFIGS))
def test_cuda_plot_rows_match_reference(cfg, stage):
    g, p = _load(cfg)
    s = _search(g, cfg)
    k0 = s.stats()["kernel_launches"]
    rows = s.plot_stats(dict(ssd=g["strict_%s_ssd" % stage], T=g["strict_%s_T" % stage],
                             pairs=g["strict_%s_pairs" % stage]))
    assert s.stats()["kernel_launches"] == k0 + 1
    _same(rows, p["%s_%s" % (cfg, stage)], "%s/%s" % (cfg, stage))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["c2", "bb"])
def test_python_plot_file_equals_the_reference_binary_s(cfg, tmp_path):
    """Search + Refine on the GPU, then WritePlotFile: the file must hash to the digest of the
    align_chimera.plot the unmodified reference binary wrote (tests/golden/msf_digests.json)"""
    from minrms_b200 import OrientationGenerator
    g, _ = _load(cfg)
    s = _search(g, cfg)
    og = OrientationGenerator(s)
    s.Search(og, 0, og.size())
    s.Refine()
    path = str(tmp_path / "align_chimera.plot")
    s.WritePlotFile(path)
    want = json.load(open(os.path.join(GOLD, "msf_digests.json")))[cfg]["align_chimera.plot"]
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == want


@pytest.mark.gpu
def test_cuda_plot_rows_match_oracle_on_synthetic_pairs():
    """seeded pairs with gaps, sizes without a solution and several atoms per residue"""
    from minrms_b200 import OrientationGenerator, SearchOrCuda
    rng = np.random.default_rng(7)
    for (m, nB, nb, ca, n_min, max_rmsd) in [(40, 46, 1, 0, 10, 2.0), (33, 30, 2, 1, 1, -1.0), (64, 70, 3, 0, 5, 4.0)]:
        A = random_chain(rng, m, nb)
        B = related_chain(rng, A, nB, nb)
        p = oc.make_params(nb=nb, n_min=n_min, m=m, nB=nB, max_rmsd=max_rmsd)
        st = settings_from_params(p, 4)
        st.ca_index = ca
        s = SearchOrCuda(A, B, st)
        og = OrientationGenerator(s)
        s.Search(og, 0, min(og.size(), 400))
        t = s.best_table()
        rows = s.plot_stats()
        want = oc.plot_stats(A, B, nb, ca, n_min, t["ssd"], t["T"], t["pairs"])
        _same(rows, want, "%dx%d nb=%d" % (m, nB, nb))
        assert rows[:, 0].sum() > 0


@pytest.mark.gpu
def test_plot_stats_error_paths():
    from minrms_b200 import MinrmsError, SearchOrCuda
    rng = np.random.default_rng(3)
    A = random_chain(rng, 12)
    B = related_chain(rng, A, 14)
    st = settings_from_params(oc.make_params(n_min=2, m=12, nB=14), 4)
    s = SearchOrCuda(A, B, st)
    st.ca_index = 1                       # not among the nb = 1 atoms used
    with pytest.raises(MinrmsError, match="C-alpha"):
        s.plot_stats()
    st.ca_index = 0
    t = s.best_table()                    # empty table: no rows, no failure
    assert s.plot_stats()[:, 0].sum() == 0
    t["ssd"][0] = 1.0
    t["pairs"][0, :2] = [[3, 3], [2, 5]]  # not ascending
    with pytest.raises(MinrmsError, match="ascending"):
        s.plot_stats(t)
