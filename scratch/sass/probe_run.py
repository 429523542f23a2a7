"""Run probe_p.cubin (see probe_patch.py) through the driver API: python scratch/sass/probe_run.py"""
import ctypes, os, sys
import numpy as np
here = os.path.dirname(os.path.abspath(__file__))
cu = ctypes.CDLL("libcuda.so.1")
def ck(r, what=""):
    if r != 0:
        s = ctypes.c_char_p(); cu.cuGetErrorString(r, ctypes.byref(s)); raise RuntimeError("%s: %s" % (what, s.value))
ck(cu.cuInit(0)); dev = ctypes.c_int(); ck(cu.cuDeviceGet(ctypes.byref(dev), 0))
ctx = ctypes.c_void_p(); ck(cu.cuDevicePrimaryCtxRetain(ctypes.byref(ctx), dev)); ck(cu.cuCtxSetCurrent(ctx))
def load(path):
    m = ctypes.c_void_p(); ck(cu.cuModuleLoadData(ctypes.byref(m), open(path, "rb").read()), "load " + path); return m
def fn(m, name):
    f = ctypes.c_void_p(); ck(cu.cuModuleGetFunction(ctypes.byref(f), m, name.encode()), name); return f
def alloc(n):
    p = ctypes.c_uint64(); ck(cu.cuMemAlloc_v2(ctypes.byref(p), n)); return p
def launch(f, grid, block, *args):
    arr = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.byref(a), ctypes.c_void_p) for a in args])
    ck(cu.cuLaunchKernel(f, grid, 1, 1, block, 1, 1, 0, None, arr, None), "launch"); ck(cu.cuCtxSynchronize(), "sync")
mod = load(os.path.join(here, "probe_p.cubin")); mod0 = load(os.path.join(here, "probe.cubin"))
# --- semantics
a = np.array([1, 5, 5, -3, 7, 0x7f800000, 100, -2**31] + [0] * 24, dtype=np.int32)
b = np.array([2, 5, 4, -3, -7, 0x7f800000, 101, 2**31 - 1] + [0] * 24, dtype=np.int32)
da, db, do = alloc(128), alloc(128), alloc(4 * 192)
for m, tag in ((mod0, "unpatched"), (mod, "patched (VIMNMX R, P0, P1, a, b)")):
    ck(cu.cuMemcpyHtoD_v2(da, a.ctypes.data, 128)); ck(cu.cuMemcpyHtoD_v2(db, b.ctypes.data, 128)); ck(cu.cuMemsetD8_v2(do, 0xff, 768))
    launch(fn(m, "sem"), 1, 32, da, db, do)
    o = np.zeros(192, dtype=np.int32); ck(cu.cuMemcpyDtoH_v2(o.ctypes.data, do, 768))
    print(tag)
    print("  a      ", a[:8]); print("  b      ", b[:8]); print("  min    ", o[:8])
    print("  P0     ", 1 - o[64:72], " (unpatched: a > b)")
    print("  P1     ", 1 - o[128:136], " (unpatched: (a^0x55) >= b+7)")
# --- throughput
seed = np.random.default_rng(1).integers(1, 1 << 30, 2048).astype(np.int32)
ds, dout, dclk = alloc(8192), alloc(4096), alloc(8)
ck(cu.cuMemcpyHtoD_v2(ds, seed.ctypes.data, 8192))
iters = 2000
for name, m, per_it in (("thr_vimnmx", mod, 32), ("thr_vimnmx_p", mod, 32), ("thr_fmnmx", mod, 32), ("thr_shf", mod, 32), ("thr_p2r", mod, 32),
                        ("thr_fsetp_pfadd", mod, 32), ("thr_isetp_pfadd", mod, 32)):
    row = "%-18s" % name
    for w in (4, 8, 12, 16, 32):
        launch(fn(m, name), 1, 32 * w, ctypes.c_int(iters), ds, dout, dclk)
        c = np.zeros(1, dtype=np.int64); ck(cu.cuMemcpyDtoH_v2(c.ctypes.data, dclk, 8))
        row += "  w%-2d %5.2f" % (w, c[0] / (iters * per_it * (w / 4.0)))
    print(row, "  clk per loop item per SMSP")
