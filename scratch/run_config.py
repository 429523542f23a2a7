"""Full FirstPass at BASELINE config 3 / 4 size on synthetic chains: python scratch/run_config.py 3|4 [max_orientations]"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle_ctypes as oc
from helpers import *
from minrms_b200 import SearchOrCuda, OrientationGenerator
from minrms_b200.search import dp_cells_per_orientation
cfg = int(sys.argv[1]); limit = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if cfg == 3:
    m, nB, frag, filt = 600, 600, 3, None
else:
    m, nB, frag, filt = 2000, 1500, 4, (8.0, 495)
rng = np.random.default_rng(1000 * cfg)
A = random_chain(rng, m); B = related_chain(rng, A, nB, noise=1.5 if cfg == 3 else 1.0)
p = oc.make_params(n_min=1, m=m, nB=nB, use_filter=filt is not None, of_dist=filt[0] if filt else 0.0, of_min=filt[1] if filt else 0)
s = SearchOrCuda(A, B, settings_from_params(p, frag))
t0 = time.perf_counter(); og = OrientationGenerator(s); t1 = time.perf_counter()
n = og.size() if not limit else min(limit, og.size())
print('config %d: %d x %d, %d orientations (%.2f s to generate), searching %d' % (cfg, m, nB, og.size(), t1 - t0, n), flush=True)
s.stats(reset=True)
t2 = time.perf_counter(); s.Search(og, 0, n); t3 = time.perf_counter()
st = s.stats()
cells = dp_cells_per_orientation(m, nB, 1, min(m, nB))
t = s.best_table()
found = int(np.sum(t['ssd'] < 1e26))
print('search %.2f s: %.0f orientations/s; nw %.1f ms (%.0f GCUPS), survivors %d, dp %.1f ms (%.0f GCUPS), tb %.1f ms; sizes with a solution %d; rmsd(n=100) %.4f' % (
    t3 - t2, n / (t3 - t2), st['nw_kernel_ms'], st['nw_cells'] / max(st['nw_kernel_ms'], 1e-9) / 1e6, st['orientations_dp'], st['dp_kernel_ms'],
    cells * st['orientations_dp'] / max(st['dp_kernel_ms'], 1e-9) / 1e6, st['tb_kernel_ms'], found, s.rmsd(100)))
if cfg != 3: sys.exit(0)   # configs 3-5 are FirstPass only (-R); Refine at config 3 size is informative
t4 = time.perf_counter(); ev = s.Refine(); t5 = time.perf_counter()
st = s.stats()
print('refine %.2f s: %d evaluations in %d batches (%d launched); rmsd(n=100) %.4f' % (t5 - t4, ev, st['refine_batches'], st['refine_evals_launched'], s.rmsd(100)))
