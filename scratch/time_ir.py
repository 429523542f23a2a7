"""-IR FirstPass (score = DP cost): costs-only DP kernel against the full kernel, test pair.
python scratch/time_ir.py"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from minrms_b200 import Settings, SearchOrCuda, OrientationGenerator
from minrms_b200.search import load_ca_pdb
A = load_ca_pdb(os.path.join(ROOT, 'tests/golden/hemoglobin.pdb'))
B = load_ca_pdb(os.path.join(ROOT, 'tests/golden/myoglobin.pdb'))
tabs = []
for costs_only in (1, 0):
    s = SearchOrCuda(A, B, Settings(fm_fragment_size=4, refine_while_searching=False))
    s.set_option("ir_costs_only", costs_only)
    og = OrientationGenerator(s)
    s.upload_orientations(og.orientations)
    for rep in range(3):
        s.reset_best(); s.stats(reset=True)
        t0 = time.perf_counter()
        s.search_resident(0, og.size(), 0)
        t1 = time.perf_counter()
        st = s.stats()
    s.fetch_best()
    tabs.append(s.best_table())
    print('costs_only=%d: %d orientations, search %.2f ms (dp kernel %.2f ms, traceback %.2f ms)' % (
        costs_only, og.size(), (t1 - t0) * 1e3, st['dp_kernel_ms'], st['tb_kernel_ms']))
    s.close()
a, b = tabs
same = all(np.array_equal(np.asarray(a[k]).view(np.uint8), np.asarray(b[k]).view(np.uint8)) for k in ('ssd', 'T', 'pairs', 'o_id'))
print('tables identical:', same)
