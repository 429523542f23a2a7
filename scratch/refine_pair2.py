"""Refine of one related pair with a bounded number of iterations: python scratch/refine_pair2.py m n iters split"""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle_ctypes as oc
from helpers import random_chain, related_chain
from minrms_b200 import Settings, SearchOrCuda
m, n, iters, split = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rng = np.random.default_rng(17 * m + n)
A = random_chain(rng, m); B, idx = related_chain(rng, A, n, noise=0.9, return_idx=True)
r, T = oc.superimpose(A[idx], B[:len(idx)])
s = SearchOrCuda(A, B, Settings(fm_fragment_size=4))
s.set_option("refine_split", split)
s.Search(T[None])
t0 = time.perf_counter(); ev = s.Refine(max_iters=iters); t1 = time.perf_counter()
st = s.stats()
print('%dx%d split=%d: refine %.1f ms: %d evaluations, %d batches (%.2f ms per batch)' % (m, n, split, (t1 - t0) * 1e3, ev, st['refine_batches'], (t1 - t0) * 1e3 / max(1, st['refine_batches'])))
