"""Refine cost of one config-5-shaped pair: python scratch/refine_pair.py m n  (MR_DEBUG_REFINE=1 for per-batch lines)"""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import random_chain
from minrms_b200 import Settings, SearchOrCuda, OrientationGenerator
m, n = int(sys.argv[1]), int(sys.argv[2])
A = random_chain(np.random.default_rng(5000), m)
B = random_chain(np.random.default_rng(5001), n)
s = SearchOrCuda(A, B, Settings(fm_fragment_size=4))
og = OrientationGenerator(s)
t0 = time.perf_counter(); s.Search(og, 0, og.size()); t1 = time.perf_counter()
st0 = s.stats()
ev = s.Refine(); t2 = time.perf_counter()
st = s.stats()
print('%dx%d: %d orientations, search %.1f ms, refine %.1f ms: %d evaluations, %d batches, %d launched, kernels %d, d2h %.1f MB' % (
    m, n, og.size(), (t1 - t0) * 1e3, (t2 - t1) * 1e3, ev, st['refine_batches'], st['refine_evals_launched'],
    st['kernel_launches'] - st0['kernel_launches'], (st['d2h_bytes'] - st0['d2h_bytes']) / 1e6))
