"""Single-pass DP kernel by chain length: min-with-predicate form (dp_form 0) against the two-min /
ping-pong forms (dp_form 1).  python scratch/time_forms.py 40 60 100 141x151 170"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minrms_b200 import OrientationGenerator, SearchOrCuda, Settings
from minrms_b200.search import dp_cells_per_orientation
from minrms_b200.synth import random_chain, related_chain
for arg in sys.argv[1:]:
    m, nB = (int(x) for x in arg.split("x")) if "x" in arg else (int(arg), int(arg))
    rng = np.random.default_rng(7 * m + nB)
    A = random_chain(rng, m); B = related_chain(rng, A, nB)
    cells = dp_cells_per_orientation(m, nB, 1, min(m, nB))
    count = int(min(60000, max(2000, 2.0e10 / cells)))
    keys = []
    for form in (0, 1):
        s = SearchOrCuda(A, B, Settings())
        s.set_option("dp_form", form)
        og = OrientationGenerator(s)
        pick = og.orientations[np.arange(count) % og.size()]
        s.upload_orientations(pick)
        for rep in range(2):
            s.reset_best(); s.stats(reset=True)
            s.search_resident(0, count, 0)
            keys_now = s.fetch_keys().copy()
            st = s.stats()
        keys.append(keys_now)
        print("%dx%d x%d dp_form %d: dp %.2f ms = %.0f GCUPS, traceback %.2f ms" % (
            m, nB, count, form, st["dp_kernel_ms"], cells * count / st["dp_kernel_ms"] / 1e6, st["tb_kernel_ms"]), flush=True)
        s.close()
    print("   keys identical:", np.array_equal(keys[0], keys[1]))
