"""phase timing of one full pair search: python scratch/time_pair.py m nB"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import random_chain
from minrms_b200 import Settings, SearchOrCuda, OrientationGenerator
from minrms_b200.search import dp_cells_per_orientation
m, nB = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(1); A = random_chain(rng, m); B = random_chain(rng, nB)
for rep in range(2):
    t0 = time.perf_counter()
    s = SearchOrCuda(A, B, Settings(fm_fragment_size=4))
    t1 = time.perf_counter()
    og = OrientationGenerator(s)
    t2 = time.perf_counter()
    s.stats(reset=True)
    s.Search(og, 0, og.size())
    t3 = time.perf_counter()
    st = s.stats()
    s.close()
    t4 = time.perf_counter()
    cells = dp_cells_per_orientation(m, nB, 1, min(m, nB)) * og.size()
    print('%dx%d %d orientations: create %.1f ms, gen %.1f ms, search %.1f ms (dp %.1f in %d launches, tb %.1f, other %.1f), close %.1f ms; dp kernel %.0f GCUPS, end to end %.0f' % (
        m, nB, og.size(), (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, st['dp_kernel_ms'], st['dp_kernel_launches'], st['tb_kernel_ms'],
        (t3-t2)*1e3 - st['dp_kernel_ms'] - st['tb_kernel_ms'], (t4-t3)*1e3, cells/st['dp_kernel_ms']/1e6, cells/(t4-t0)/1e9))
