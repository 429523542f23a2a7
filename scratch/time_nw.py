"""NW filter timing: python scratch/time_nw.py m nB count"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle_ctypes as oc
from helpers import *
from minrms_b200 import SearchOrCuda, OrientationGenerator
m, nB, count = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(7 * m + nB)
A = random_chain(rng, m); B = related_chain(rng, A, nB, noise=1.2)
of_min = int(0.33 * min(m, nB))
p = oc.make_params(n_min=1, m=m, nB=nB, use_filter=True, of_dist=8.0, of_min=of_min)
s = SearchOrCuda(A, B, settings_from_params(p, 4))
og = OrientationGenerator(s)
pick = og.orientations[np.linspace(0, og.size() - 1, count).astype(int)]
s.upload_orientations(pick)
for rep in range(2):
    s.reset_best(); s.stats(reset=True)
    s.search_resident(0, count, 0); s.fetch_keys()
    st = s.stats()
print('m %d nB %d count %d: nw %.2f ms (%d cells, %.1f GCUPS) survivors %d dp %.2f ms' % (
    m, nB, count, st['nw_kernel_ms'], st['nw_cells'], st['nw_cells'] / st['nw_kernel_ms'] / 1e6, st['orientations_dp'], st['dp_kernel_ms']))
