"""Small end-to-end cases for compute-sanitizer (racecheck / memcheck), one tool per gpurun call:
  compute-sanitizer --tool racecheck python scratch/sanitize_case.py
Covers: single-pass DP (ping-pong and in-place buckets), NW filter, traceback/Kabsch, select,
winners' payload, Refine batches; the multi-pass DP (230 x 250 pair, two passes) and its
cluster-split mode inside Refine batches (300 x 280 pair); the -IR costs-only kernel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minrms_b200 import OrientationGenerator, SearchOrCuda, Settings
from minrms_b200.synth import random_chain, related_chain


def run(name, m, nB, limit, settings, refine_iters):
    rng = np.random.default_rng(m * 1000 + nB)
    A = random_chain(rng, m)
    B = related_chain(rng, A, nB)
    s = SearchOrCuda(A, B, settings)
    og = OrientationGenerator(s)
    n = min(limit, og.size())
    s.Search(og, 0, n)
    ev = s.Refine(refine_iters) if refine_iters else 0
    t = s.best_table()
    print("%s: %dx%d, %d orientations, %d sizes solved, %d refine evaluations, kernels launched %d" % (
        name, m, nB, n, int((t["ssd"] < 1e26).sum()), ev, s.stats()["kernel_launches"]), flush=True)
    s.close()


run("small, filter + max-rmsd (in-place tile update)", 40, 46, 400,
    Settings(n_min=10, dyn3d_max_rmsd=5.0, orientation_filter_nw=True, orientation_filter_nw_max_dist=8.0,
             orientation_filter_nw_min_num_matches=10), 6)
run("test-pair size (ping-pong tile update)", 141, 151, 160, Settings(), 3)
run("-IR (costs-only kernel)", 100, 110, 200, Settings(refine_while_searching=False), 0)
run("multi-pass DP", 230, 250, 150, Settings(), 0)
run("multi-pass DP, cluster-split Refine batches", 300, 280, 150, Settings(), 2)
