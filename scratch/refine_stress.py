"""Refine with the cluster-split DP against one CTA per orientation, repeated: tables must be identical."""
import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import random_chain, related_chain
from minrms_b200 import Settings, SearchOrCuda, OrientationGenerator
for (m, n, seed) in [(300, 280, 1), (233, 310, 2), (380, 400, 3), (520, 470, 4)]:
    rng = np.random.default_rng(seed)
    A = random_chain(rng, m); B = related_chain(rng, A, n, noise=1.5)
    ref = None
    for rep, split in enumerate([0, 1, 1, 1]):
        s = SearchOrCuda(A, B, Settings(fm_fragment_size=4, fm_cutoff_rmsd=0.8))
        s.set_option("refine_split", split)
        og = OrientationGenerator(s)
        s.Search(og, 0, min(og.size(), 3000))
        ev = s.Refine()
        t = s.best_table()
        key = (ev, t["ssd"].tobytes(), t["T"].tobytes(), t["pairs"].tobytes())
        if ref is None: ref = key
        assert key == ref, "%dx%d rep %d split %d differs" % (m, n, rep, split)
        s.close()
    print("%dx%d: %d evaluations, split == single-CTA, 3 repeats identical" % (m, n, ref[0]))
