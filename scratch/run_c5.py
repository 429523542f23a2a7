"""Config-5-shaped all-vs-all slice: python scratch/run_c5.py n_chains  (lengths ~U{150..400}, seeds 5000+i)"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import random_chain
from minrms_b200 import Settings
from minrms_b200.batch import all_vs_all, pair_cost
n = int(sys.argv[1])
chains = []
for i in range(n):
    rng = np.random.default_rng(5000 + i)
    chains.append(random_chain(rng, int(rng.integers(150, 401))))
st = Settings(fm_fragment_size=4)
lens = [c.shape[0] for c in chains]
cells = sum(pair_cost(lens[i], lens[j], st) for i in range(n) for j in range(i + 1, n))
t0 = time.perf_counter()
last=[time.perf_counter()]
def show(i,j,r):
    now=time.perf_counter(); print('pair %d-%d (%dx%d) %d orientations: %.1f ms' % (i,j,lens[i],lens[j],r.orientations,(now-last[0])*1e3), flush=True); last[0]=now
import os
res = all_vs_all(chains, st, refine=bool(int(os.environ.get('C5_REFINE','1'))), on_result=show if len(sys.argv)>2 else None, workers=int(os.environ.get('C5_WORKERS','1')))
t1 = time.perf_counter()
ori = sum(r.orientations for r in res.values())
print('%d chains (%d..%d residues), %d pairs, %d orientations, %.3e DP cells: %.2f s -> %.2f pairs/s, %.0f orientations/s, %.0f GCUPS end to end' % (
    n, min(lens), max(lens), len(res), ori, cells, t1 - t0, len(res) / (t1 - t0), ori / (t1 - t0), cells / (t1 - t0) / 1e9))
