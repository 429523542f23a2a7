"""DP kernel timing on mid/large synthetic pairs: python scratch/time_mid.py m nB count [kernel] [mp_threads]"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle_ctypes as oc
from helpers import *
from minrms_b200 import SearchOrCuda, OrientationGenerator
from minrms_b200.search import dp_cells_per_orientation
m, nB, count = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kernels = [int(k) for k in sys.argv[4].split(',')] if len(sys.argv) > 4 else [0]
mpt = int(sys.argv[5]) if len(sys.argv) > 5 else 0
rng = np.random.default_rng(7 * m + nB)
A = random_chain(rng, m); B = related_chain(rng, A, nB, noise=1.2)
p = oc.make_params(n_min=1, m=m, nB=nB)
ref = None
for kern in kernels:
    s = SearchOrCuda(A, B, settings_from_params(p, 4))
    if mpt: s.set_option("mp_threads", mpt)
    s.set_option("dp_kernel", kern)
    og = OrientationGenerator(s)
    pick = og.orientations[np.linspace(0, og.size() - 1, count).astype(int)]
    cells = dp_cells_per_orientation(m, nB, 1, min(m, nB))
    s.upload_orientations(pick)
    for rep in range(2):
        s.reset_best(); s.stats(reset=True)
        t2 = time.perf_counter(); s.search_resident(0, count, 0); s.fetch_keys(); t3 = time.perf_counter()
        st = s.stats()
    s.fetch_best()
    t = s.best_table()
    same = '' if ref is None else ' same_as_first=%s' % (np.array_equal(bits(t['ssd']), bits(ref['ssd'])) and np.array_equal(t['pairs'], ref['pairs']))
    if ref is None: ref = t
    print('m %d nB %d count %d kernel %d mp_threads %d: search %.1f ms dp %.2f ms tb %.2f ms; %.1f GCUPS%s' % (
        m, nB, count, kern, mpt, (t3 - t2) * 1e3, st['dp_kernel_ms'], st['tb_kernel_ms'],
        cells * count / st['dp_kernel_ms'] / 1e6 if cells else 0, same))
    del s
