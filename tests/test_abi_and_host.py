"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares,
fails loudly without a GPU, and the host-side logic (shuffle, parameter mapping)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_ctypes as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        src = open(os.path.join(inc, f)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(mr_[a-z_0-9]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from minrms_b200 import _lib, build
    build.build()
    L = C.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), "missing export " + s
    assert set(_lib.EXPORTS) == set(syms)


def test_sass_is_sm100a():
    from minrms_b200 import _lib, build
    build.build()
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_traceback_kernel_instances_and_registers():
    """The FirstPass traceback / superposition kernel exists per layout (tiles, columns) in three
    instances -- plain, emitting (payload, Refine) and with the common-pairs test of
    mr_search_dissimilar -- and the two throughput instances stay at the 40 registers their
    launch bounds (12 blocks of 128 threads per SM) allow: the -alt-r variant is a separate
    instantiation, the FirstPass kernel is not touched by it."""
    from minrms_b200 import _lib, build
    build.build()
    out = subprocess.run(["cuobjdump", "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    regs = {}
    for fn, r in re.findall(r"Function (\S*k_traceback_kabsch\S*):\s*\n\s*REG:(\d+)", out):
        layout = "tiles" if "TbTiled" in fn else "columns"
        emit, excl = re.search(r"Lb([01])ELb([01])E", fn).groups()
        regs[(layout, int(emit), int(excl))] = int(r)
    for layout in ("tiles", "columns"):
        assert regs[(layout, 0, 0)] <= 40 and regs[(layout, 0, 1)] <= 40, regs
        assert (layout, 1, 0) in regs and (layout, 1, 1) not in regs, regs


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="CPU-only check")
def test_create_fails_loudly_without_gpu():
    from minrms_b200 import SearchOrCuda, Settings, MinrmsError
    A = np.zeros((5, 3), np.float32)
    with pytest.raises(MinrmsError, match="no CPU fallback"):
        SearchOrCuda(A, A, Settings())


def test_interpret_num_matches():
    from minrms_b200 import interpret_num_matches
    assert interpret_num_matches(0.33, 141, 151) == 47      # config 1: -minN 0.33
    assert interpret_num_matches(47, 141, 151) == 47
    assert interpret_num_matches(0.33, 1500, 2000) == 495   # config 4 filter minimum


def test_shuffle_matches_libc_rand():
    """the restated glibc rand()/std::random_shuffle against the real libc in a fresh
    process (never-seeded rand(), as in lib/or_gen/or_gen.cc:179)"""
    code = ("import ctypes,sys; L=ctypes.CDLL('libc.so.6'); n=int(sys.argv[1]); p=list(range(n))\n"
            "for i in range(1,n):\n j=L.rand()%(i+1)\n p[i],p[j]=p[j],p[i]\n"
            "print(' '.join(map(str,p)))")
    out = subprocess.run([sys.executable, "-c", code, "5000"], capture_output=True, text=True, check=True)
    expect = np.array(out.stdout.split(), np.int64)
    assert np.array_equal(oc.shuffle_perm(5000), expect)


def test_partition_covers_everything():
    from minrms_b200.dist import partition
    for count in (0, 1, 7, 20424):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = partition(count, r, world)
                seen.extend(range(a, b))
            assert seen == list(range(count))


def test_all_vs_all_schedule_covers_every_pair_once_and_balances():
    """minrms_b200/batch.py: LPT partition of the unordered pairs (config 5 driver)"""
    from minrms_b200 import Settings
    from minrms_b200.batch import pair_cost, pair_settings, schedule
    rng = np.random.default_rng(5000)
    lengths = [int(x) for x in rng.integers(150, 401, 24)]
    st = Settings(fm_fragment_size=4)
    all_pairs = {(i, j) for i in range(24) for j in range(i + 1, 24)}
    for world in (1, 2, 8):
        parts = schedule(lengths, world, st)
        assert len(parts) == world
        flat = [p for part in parts for p in part]
        assert len(flat) == len(all_pairs) and set(flat) == all_pairs
        assert schedule(lengths, world, st) == parts                      # deterministic
        loads = [sum(pair_cost(lengths[i], lengths[j], st) for i, j in part) for part in parts]
        biggest = max(pair_cost(lengths[i], lengths[j], st) for i, j in all_pairs)
        assert max(loads) - min(loads) <= biggest                         # LPT bound
    # (m-L+1)(n-L+1) orientations x closed-form cell count
    assert pair_cost(141, 151, st) == 138 * 148 * 1044481
    assert pair_cost(3, 200, st) == 0                                     # shorter than the fragment
    ps = pair_settings(st, 141, 151, min_n=0.33, of_min=0.33)
    assert ps.n_min == 47 and ps.orientation_filter_nw_min_num_matches == 47 and st.n_min == 1


def test_bench_config_is_the_same_object_in_both_arms():
    """bench.py: the `config` both arms print comes from one function of the workload and the launch
    alone (the driver compares the two), and its closed-form cell count is the library's."""
    import argparse
    import importlib.util
    import os
    from minrms_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    L = _lib.load()
    for m, nB, n_min in ((141, 151, 1), (141, 151, 47), (600, 600, 1), (2000, 1500, 1), (40, 46, 10), (7, 2, 2)):
        n_max = min(m, nB)
        assert bench.dp_cells_closed_form(m, nB, n_min, n_max) == int(L.mr_dp_cells_per_orientation(m, nB, n_min, n_max))
    args = argparse.Namespace(scaling="weak")
    for cfg in (1, 2):
        w = bench.workload(cfg)
        a = bench.common_config(w, args, 20424)
        b = bench.common_config(w, args, 20424)
        assert a == b and a["config"] == cfg and a["dp_cells_per_orientation"] in (794770, 1044481)
        assert set(a) >= {"workload", "orientations_in_list", "orientations_per_gpu_per_step", "alignment_sizes", "l2", "parallelism"}
