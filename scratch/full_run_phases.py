"""Where a whole run goes (warm process): python scratch/full_run_phases.py [config]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from minrms_b200 import OrientationGenerator, SearchOrCuda
w = bench.workload(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
for rep in range(4):
    t = [time.perf_counter()]
    s = SearchOrCuda(w["A"], w["B"], w["settings"]); t.append(time.perf_counter())
    og = OrientationGenerator(s); t.append(time.perf_counter())
    s.Search(og, 0, og.size()); t.append(time.perf_counter())
    ev = s.Refine(); t.append(time.perf_counter())
    s.close(); t.append(time.perf_counter())
    d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
    print("rep %d: create %.1f ms, orientations %.1f, Search %.1f, Refine %.1f (%d evaluations), close %.1f, total %.1f" % (
        rep, d[0], d[1], d[2], d[3], ev, d[4], sum(d)), flush=True)
