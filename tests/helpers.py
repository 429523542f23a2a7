"""shared helpers for the parity tests"""
import hashlib

import numpy as np

import oracle_ctypes as oc


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def params_from_golden(g):
    return oc.OrcParams(int(g["nb"]), int(g["n_min"]), int(g["n_max"]), float(g["max_dist"]),
                        float(g["max_rmsd"]), int(g["use_filter"]), float(g["of_dist"]),
                        int(g["of_min"]), int(g["ir"]))


def settings_from_golden(g):
    from minrms_b200 import Settings
    return Settings(nb=int(g["nb"]), n_min=int(g["n_min"]), n_max=int(g["n_max"]),
                    dyn3d_max_dist=float(g["max_dist"]), dyn3d_max_rmsd=float(g["max_rmsd"]),
                    orientation_filter_nw=bool(g["use_filter"]),
                    orientation_filter_nw_max_dist=float(g["of_dist"]),
                    orientation_filter_nw_min_num_matches=int(g["of_min"]),
                    refine_while_searching=bool(g["ir"]), fm_fragment_size=int(g["frag"]),
                    fm_cutoff_rmsd=float(g["fm_cutoff"]) if "fm_cutoff" in g else -1.0)


def settings_from_params(p, frag=4):
    from minrms_b200 import Settings
    return Settings(nb=p.nb, n_min=p.n_min, n_max=p.n_max, dyn3d_max_dist=p.max_dist,
                    dyn3d_max_rmsd=p.max_rmsd, orientation_filter_nw=bool(p.use_filter),
                    orientation_filter_nw_max_dist=p.of_dist,
                    orientation_filter_nw_min_num_matches=p.of_min,
                    refine_while_searching=bool(p.refine_while_searching), fm_fragment_size=frag)


def assert_eval_equal(res, n_last, cost, valid, ssd, T=None, pairs=None, what=""):
    """res: dict from Workspace.eval / SearchOrCuda.eval_one; the rest: expected."""
    assert res["n_last"] == n_last, (what, res["n_last"], n_last)
    nl = int(n_last)
    assert np.array_equal(bits(res["cost"][:nl]), bits(cost[:nl])), what + " cost"
    assert np.array_equal(res["valid"].astype(bool), np.asarray(valid).astype(bool)), what + " valid"
    v = np.asarray(valid).astype(bool)
    assert np.array_equal(bits(res["ssd"][v]), bits(np.asarray(ssd)[v])), what + " ssd"
    if T is not None:
        assert np.array_equal(bits(res["T"][v]), bits(np.asarray(T)[v])), what + " T"
    if pairs is not None:
        for n in np.nonzero(v)[0] + 1:
            assert np.array_equal(res["pairs"][n - 1][:n], pairs[n - 1][:n]), (what, "pairs", n)


def assert_table_equal(ssd, T, pairs, exp_ssd, exp_T, exp_pairs, n_min, what=""):
    assert np.array_equal(bits(ssd), bits(exp_ssd)), what + " ssd"
    have = np.asarray(exp_ssd) < np.float32(1e26)
    assert np.array_equal(bits(T[have]), bits(exp_T[have])), what + " T"
    for r in np.nonzero(have)[0]:
        n = n_min + r
        assert np.array_equal(pairs[r][:n], exp_pairs[r][:n]), (what, "pairs", n)


from minrms_b200.synth import random_chain, related_chain  # noqa: E402,F401  (seeded synthetic chains)
