"""All-vs-all batch driver (config 5 shape) on the GPU against the oracle, pair by pair."""
import numpy as np
import pytest

import oracle_ctypes as oc
from helpers import assert_table_equal, random_chain, related_chain, settings_from_params

pytestmark = pytest.mark.gpu


def _chains():
    rng = np.random.default_rng(5000)
    base = random_chain(rng, 44)
    lens = [44, 31, 38, 27, 3]                  # the 3-residue chain is shorter than the fragment
    return [base] + [related_chain(rng, base, n, noise=0.7) for n in lens[1:]]


@pytest.mark.parametrize("workers", [1, 3])
def test_all_vs_all_matches_oracle_pair_by_pair(workers):
    """workers = 3: three host threads, each with its own context on the GPU, share the list"""
    from minrms_b200 import Settings
    from minrms_b200.batch import all_vs_all
    chains = _chains()
    st = Settings(fm_fragment_size=4)
    seen = []
    res = all_vs_all(chains, st, on_result=lambda i, j, r: seen.append((i, j)), workers=workers)
    # every pair of the four long chains; pairs with the 3-residue chain are refused
    # (no 4-residue fragment, lib/or_gen/or_gen.cc:548-580)
    want = {(i, j) for i in range(4) for j in range(i + 1, 4)}
    assert set(res) == want and sorted(seen) == sorted(want)
    for (i, j), r in res.items():
        A, B = chains[i], chains[j]
        p = oc.make_params(n_min=1, m=A.shape[0], nB=B.shape[0])
        og = oc.gen_orientations(A, B, 1, 4, 1)
        ws = oc.Workspace(A, B, p)
        best = ws.search(og)
        ev = ws.refine(best)
        assert r.orientations == og.shape[0] and r.refine_evals == ev
        assert_table_equal(r.ssd, r.T, r.pairs, best["ssd"], best["T"], best["pairs"], 1, "pair %d-%d" % (i, j))


def test_all_vs_all_ranks_partition_the_pairs():
    """two ranks (run one after the other on the same GPU): disjoint shares, identical results"""
    from minrms_b200 import Settings
    from minrms_b200.batch import all_vs_all
    chains = _chains()[:4]
    st = Settings(fm_fragment_size=4)
    whole = all_vs_all(chains, st, refine=False, min_n=0.5)
    parts = [all_vs_all(chains, st, rank=k, world=2, refine=False, min_n=0.5) for k in range(2)]
    assert not set(parts[0]) & set(parts[1])
    assert set(parts[0]) | set(parts[1]) == set(whole)
    for part in parts:
        for key, r in part.items():
            w = whole[key]
            assert r.n_min == w.n_min == max(1, int(np.floor(0.5 * min(chains[key[0]].shape[0], chains[key[1]].shape[0]) + 0.5)))
            assert np.array_equal(r.ssd.view(np.uint32), w.ssd.view(np.uint32)) and np.array_equal(r.pairs, w.pairs)


def test_batch_cli_writes_the_reference_s_msf_files(tmp_path):
    """`python -m minrms_b200.batch DIR --msf-dir OUT` on the reference's test pair = config 2
    (-fm 4 -HS -ir -r): the 140 align<n>.msf files must hash to the digests of the files the
    unmodified reference binary wrote (tests/golden/msf_digests.json)"""
    import json
    import os
    import shutil
    import sys
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold)
    from make_msf_digests import digest_dir
    from minrms_b200.batch import main
    src = tmp_path / "pdb"
    src.mkdir()
    for f in ("hemoglobin.pdb", "myoglobin.pdb"):
        shutil.copy(os.path.join(gold, f), str(src))
    assert main([str(src), "--out", str(tmp_path / "r.npz"), "--msf-dir", str(tmp_path / "msf"), "--workers", "2"]) == 0
    want = {k: v for k, v in json.load(open(os.path.join(gold, "msf_digests.json")))["c2"].items() if k.endswith(".msf")}
    got = digest_dir(str(tmp_path / "msf" / "hemoglobin__myoglobin"))
    assert sorted(got) == sorted(want) and len(got) == 140
    assert not [f for f in want if got[f] != want[f]]
