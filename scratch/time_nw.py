"""NW filter alone: decision-only (cost) kernels against the counting kernels.
python scratch/time_nw.py m nB n_orientations of_frac"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minrms_b200 import OrientationGenerator, SearchOrCuda, Settings
from minrms_b200.synth import random_chain, related_chain
m, nB, n_or, frac = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
rng = np.random.default_rng(4000)
A = random_chain(rng, m); B = related_chain(np.random.default_rng(4001), A, nB)
k = int(frac * min(m, nB))
st = Settings(orientation_filter_nw=True, orientation_filter_nw_max_dist=8.0, orientation_filter_nw_min_num_matches=k)
for by_cost in (1, 0):
    s = SearchOrCuda(A, B, st)
    s.set_option("nw_cost_filter", by_cost)
    og = OrientationGenerator(s)
    n = min(n_or, og.size())
    s.upload_orientations(og.orientations[:n])
    for rep in range(2):
        s.reset_best(); s.stats(reset=True)
        s.search_resident(0, n, 0)
        stt = s.stats()
    print("%dx%d of_min %d, %d orientations, by_cost=%d: nw %.2f ms = %.0f GCUPS, survivors %d" % (
        m, nB, k, n, by_cost, stt["nw_kernel_ms"], stt["nw_cells"] / stt["nw_kernel_ms"] / 1e6, stt["orientations_dp"]))
    s.close()
