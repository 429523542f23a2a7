"""DP kernel by chain length: min-with-predicate form (dp_form 0; multi-pass: 384- and 320-thread
buckets) against the two-min / ping-pong forms (dp_form 1).
python scratch/time_forms.py 40 60 100 141x151 170 300 600"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minrms_b200 import OrientationGenerator, SearchOrCuda, Settings
from minrms_b200.search import dp_cells_per_orientation
from minrms_b200.synth import random_chain, related_chain
NMIN = int(os.environ.get("NMIN", "1"))   # > 1: the windowed kernel variants
for arg in sys.argv[1:]:
    m, nB = (int(x) for x in arg.split("x")) if "x" in arg else (int(arg), int(arg))
    rng = np.random.default_rng(7 * m + nB)
    A = random_chain(rng, m); B = related_chain(rng, A, nB)
    cells = dp_cells_per_orientation(m, nB, NMIN, min(m, nB))
    count = int(min(60000, max(592, 2.0e10 / cells)))
    variants = [("vp", {"dp_form": 0}), ("legacy", {"dp_form": 1})]
    keys = []
    for label, opts in variants:
        s = SearchOrCuda(A, B, Settings(n_min=NMIN))
        for k, v in opts.items():
            s.set_option(k, v)
        og = OrientationGenerator(s)
        pick = og.orientations[np.arange(count) % og.size()]
        s.upload_orientations(pick)
        for rep in range(2):
            s.reset_best(); s.stats(reset=True)
            s.search_resident(0, count, 0)
            keys_now = s.fetch_keys().copy()
            st = s.stats()
        keys.append(keys_now)
        print("%dx%d x%d %-10s: dp %.2f ms = %.0f GCUPS, traceback %.2f ms" % (
            m, nB, count, label, st["dp_kernel_ms"], cells * count / st["dp_kernel_ms"] / 1e6, st["tb_kernel_ms"]), flush=True)
        s.close()
    print("   keys identical:", all(np.array_equal(keys[0], k) for k in keys[1:]))
