"""Cost of the alternate-solutions modes on the test pair (config 2: 141 x 151, 20 424 orientations):
mr_search against mr_search_scores (-alt-o: one more store per compared size + the D2H of the
score array) and mr_search_dissimilar with one and two lower ranks (-alt-r: the common-pairs count
in the traceback kernel), host buffers, wall clock around the C ABI call; then the wall time of the
command line with and without -alt-o / -alt-r next to the reference binary's.
python scratch/time_alternates.py [--cli]"""
import os, subprocess, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from minrms_b200 import Settings, SearchOrCuda, OrientationGenerator
from minrms_b200.search import load_ca_pdb
A = load_ca_pdb(os.path.join(ROOT, 'tests/golden/hemoglobin.pdb'))
B = load_ca_pdb(os.path.join(ROOT, 'tests/golden/myoglobin.pdb'))
s = SearchOrCuda(A, B, Settings(fm_fragment_size=4))
og = OrientationGenerator(s)
orient = og.orientations
R, N = s.R, s.N
def fresh():
    s.load_table(dict(ssd=np.full(R, np.float32(1e26), np.float32), T=np.zeros((R, 12), np.float32),
                      pairs=np.full((R, N, 2), -1, np.int16), o_id=np.full(R, -1, np.int64)))
def timed(f, reps=5):
    best = 1e9
    for _ in range(reps):
        fresh(); s.stats(reset=True)
        t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, s.stats()
ms0, st0 = timed(lambda: s.Search(orient))
rank0 = s.best_table()
ms1, st1 = timed(lambda: s.SearchScores(orient))
ms2, st2 = timed(lambda: s.SearchDissimilar(orient, [rank0], 60))
rank1 = s.best_table()
ms3, st3 = timed(lambda: s.SearchDissimilar(orient, [rank0, rank1], 60))
print('orientations %d, sizes %d' % (orient.shape[0], R))
for name, ms, st in (('mr_search', ms0, st0), ('mr_search_scores', ms1, st1), ('mr_search_dissimilar 1 lower', ms2, st2),
                     ('mr_search_dissimilar 2 lower', ms3, st3)):
    print('%-30s %8.2f ms  (dp %.2f ms, traceback %.2f ms, d2h %.1f MB)' % (name, ms, st['dp_kernel_ms'], st['tb_kernel_ms'], st['d2h_bytes'] / 1e6))
s.close()
if '--cli' in sys.argv:
    from make_msf_digests import CONFIGS, run
    import tempfile
    for cfg in ('c1', 'alt', 'altr'):
        for who in ('minrms_cuda', 'minrms'):
            exe = os.path.join(ROOT, 'oracle', '_ref', 'strict', who)
            with tempfile.TemporaryDirectory() as d:
                t0 = time.perf_counter(); run(exe, CONFIGS[cfg], d, cfg); dt = time.perf_counter() - t0
            print('cli %-5s %-12s %7.2f s' % (cfg, who, dt))
