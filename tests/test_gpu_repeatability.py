"""Race evidence without a sanitizer (compute-sanitizer is closed on this GPU pool): the DP
kernels exchange tile borders through double-buffered shared memory and a carry buffer that
is rewritten in place, so a missing barrier would show up as run-to-run differences.  The
same searches are repeated with different chunk sizes (different CTA-to-orientation
assignments and wave boundaries) and the 64-bit per-size keys compared bit for bit."""
import numpy as np
import pytest

from helpers import settings_from_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,kernel", [("golden_c2", 0), ("golden_c1", 0), ("golden_mid", 3), ("golden_mid", 0)])
def test_repeated_searches_are_bit_identical(cfg, kernel, request):
    from minrms_b200 import OrientationGenerator, SearchOrCuda
    g = request.getfixturevalue(cfg)
    s = SearchOrCuda(g["A"], g["B"], settings_from_golden(g))
    if kernel == 3:
        s.set_option("mp_threads", 64)
        s.set_option("dp_kernel", 3)
    og = OrientationGenerator(s, intervals=g["intervals"] if "intervals" in g else None)
    n = min(og.size(), 6000)
    s.upload_orientations(og.orientations[:n])
    ref = None
    for chunk in (0, 0, 700, 149, 0):
        s.set_option("chunk", chunk)
        s.reset_best()
        s.search_resident(0, n, 0)
        keys = s.fetch_keys()
        if ref is None:
            ref = keys
        assert np.array_equal(keys, ref), "chunk=%d" % chunk
    s.close()
