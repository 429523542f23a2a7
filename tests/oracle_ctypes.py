"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs -- never from minrms_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
UPPER_BOUND = np.float32(1.0e26)


class OrcParams(C.Structure):
    _fields_ = [("nb", C.c_int32), ("n_min", C.c_int32), ("n_max", C.c_int32),
                ("max_dist", C.c_float), ("max_rmsd", C.c_float),
                ("use_filter", C.c_int32), ("of_dist", C.c_float), ("of_min", C.c_int32),
                ("refine_while_searching", C.c_int32)]


class OrcEval(C.Structure):
    _fields_ = [("n_last", C.c_int32), ("nw_matches", C.c_int32),
                ("cost", C.POINTER(C.c_float)), ("valid", C.POINTER(C.c_uint8)),
                ("ssd", C.POINTER(C.c_float)), ("T", C.POINTER(C.c_float)),
                ("pairs", C.POINTER(C.c_int16))]


class OrcBest(C.Structure):
    _fields_ = [("ssd", C.POINTER(C.c_float)), ("T", C.POINTER(C.c_float)),
                ("pairs", C.POINTER(C.c_int16)), ("o_id", C.POINTER(C.c_int64))]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "minrms_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        L.orc_superimpose.restype = C.c_float
        L.orc_superimpose.argtypes = [C.c_int, fp, fp, fp]
        L.orc_apply_transform.argtypes = [fp, fp, C.c_int, fp]
        L.orc_gen_orientations.restype = C.c_int64
        L.orc_gen_orientations.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_float,
                                           C.c_int, C.POINTER(C.c_int32), C.c_int, fp, C.c_int64]
        L.orc_shuffle_orientations.argtypes = [fp, C.c_int64]
        L.orc_shuffle_perm.argtypes = [C.POINTER(C.c_int64), C.c_int64]
        L.orc_nw_num_matches.restype = C.c_int
        L.orc_nw_num_matches.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.c_float, C.c_int]
        L.orc_ws_create.restype = C.c_void_p
        L.orc_ws_create.argtypes = [C.c_int, C.c_int, C.POINTER(OrcParams)]
        L.orc_ws_destroy.argtypes = [C.c_void_p]
        L.orc_eval_orientation.argtypes = [C.c_void_p, fp, fp, fp, C.c_int, C.POINTER(OrcEval)]
        L.orc_best_init.argtypes = [C.POINTER(OrcParams), C.c_int, C.POINTER(OrcBest)]
        L.orc_search.restype = C.c_int64
        L.orc_search.argtypes = [C.c_void_p, fp, fp, fp, C.c_int64, C.c_int64, C.POINTER(OrcBest)]
        L.orc_refine.restype = C.c_int
        L.orc_refine.argtypes = [C.c_void_p, fp, fp, C.POINTER(OrcBest), C.c_int]
        L.orc_plot_stats.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(OrcBest), C.POINTER(C.c_double)]
        L.orc_sstr_probability.argtypes = [C.c_float, C.c_int, fp]
        L.orc_dp_cells.restype = C.c_int64
        L.orc_dp_cells.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def make_params(nb=1, n_min=1, n_max=None, max_dist=-1.0, max_rmsd=-1.0, use_filter=False,
                of_dist=0.0, of_min=0, refine_while_searching=True, m=None, nB=None):
    if n_max is None:
        n_max = min(m, nB)
    return OrcParams(nb, n_min, n_max, max_dist, max_rmsd, int(use_filter), of_dist, of_min,
                     int(refine_while_searching))


def superimpose(fixed, moveable):
    fixed = np.ascontiguousarray(fixed, np.float32)
    moveable = np.ascontiguousarray(moveable, np.float32)
    T = np.zeros(12, np.float32)
    r = lib().orc_superimpose(fixed.shape[0], _fp(fixed), _fp(moveable), _fp(T))
    return np.float32(r), T


def apply_transform(T, pts):
    pts = np.ascontiguousarray(pts, np.float32)
    T = np.ascontiguousarray(T, np.float32)
    out = np.empty_like(pts)
    lib().orc_apply_transform(_fp(T), _fp(pts), pts.shape[0], _fp(out))
    return out


def gen_orientations(A, B, nb, frag_len, n_min, cutoff_rmsd=-1.0, intervals=None, shuffle=True):
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    m, nB = A.shape[0] // nb, B.shape[0] // nb
    iv = None
    niv = 0
    if intervals is not None and len(intervals):
        ivarr = np.ascontiguousarray(intervals, np.int32).reshape(-1, 4)
        iv = ivarr.ctypes.data_as(C.POINTER(C.c_int32))
        niv = ivarr.shape[0]
    cap = m * nB
    out = np.zeros((cap, 12), np.float32)
    cnt = lib().orc_gen_orientations(_fp(A), m, _fp(B), nB, nb, frag_len, cutoff_rmsd, n_min,
                                     iv, niv, _fp(out), cap)
    out = np.ascontiguousarray(out[:cnt])
    if shuffle:
        lib().orc_shuffle_orientations(_fp(out), cnt)
    return out


def shuffle_perm(count):
    perm = np.zeros(count, np.int64)
    lib().orc_shuffle_perm(perm.ctypes.data_as(C.POINTER(C.c_int64)), count)
    return perm


def nw_num_matches(A, Bp, nb, of_dist, of_min):
    A = np.ascontiguousarray(A, np.float32)
    Bp = np.ascontiguousarray(Bp, np.float32)
    return lib().orc_nw_num_matches(_fp(A), A.shape[0] // nb, _fp(Bp), Bp.shape[0] // nb, nb,
                                    of_dist, of_min)


class Workspace:
    def __init__(self, A, B, params):
        self.A = np.ascontiguousarray(A, np.float32)
        self.B = np.ascontiguousarray(B, np.float32)
        self.p = params
        self.m = self.A.shape[0] // params.nb
        self.nB = self.B.shape[0] // params.nb
        self.N = min(self.m, self.nB)
        self.R = params.n_max - params.n_min + 1
        self.ws = lib().orc_ws_create(self.m, self.nB, C.byref(params))

    def __del__(self):
        if getattr(self, "ws", None):
            lib().orc_ws_destroy(self.ws)
            self.ws = None

    def eval(self, T12, apply_filter=True):
        N = self.N
        T12 = np.ascontiguousarray(T12, np.float32)
        res = dict(cost=np.zeros(N, np.float32), valid=np.zeros(N, np.uint8),
                   ssd=np.zeros(N, np.float32), T=np.zeros((N, 12), np.float32),
                   pairs=np.zeros((N, N, 2), np.int16))
        ev = OrcEval(0, 0, _fp(res["cost"]), res["valid"].ctypes.data_as(C.POINTER(C.c_uint8)),
                     _fp(res["ssd"]), _fp(res["T"]),
                     res["pairs"].ctypes.data_as(C.POINTER(C.c_int16)))
        lib().orc_eval_orientation(self.ws, _fp(self.A), _fp(self.B), _fp(T12), int(apply_filter),
                                   C.byref(ev))
        res["n_last"] = ev.n_last
        res["nw_matches"] = ev.nw_matches
        return res

    def new_best(self):
        R, N = self.R, self.N
        b = dict(ssd=np.zeros(R, np.float32), T=np.zeros((R, 12), np.float32),
                 pairs=np.full((R, N, 2), -1, np.int16), o_id=np.zeros(R, np.int64))
        lib().orc_best_init(C.byref(self.p), N, C.byref(self._best_struct(b)))
        return b

    @staticmethod
    def _best_struct(b):
        return OrcBest(_fp(b["ssd"]), _fp(b["T"]), b["pairs"].ctypes.data_as(C.POINTER(C.c_int16)),
                       b["o_id"].ctypes.data_as(C.POINTER(C.c_int64)))

    def search(self, orient, best=None, first_id=0):
        orient = np.ascontiguousarray(orient, np.float32).reshape(-1, 12)
        if best is None:
            best = self.new_best()
        cells = lib().orc_search(self.ws, _fp(self.A), _fp(self.B), _fp(orient), orient.shape[0],
                                 first_id, C.byref(self._best_struct(best)))
        best["cells"] = best.get("cells", 0) + cells
        return best

    def refine(self, best, max_iters=-1):
        return lib().orc_refine(self.ws, _fp(self.A), _fp(self.B),
                                C.byref(self._best_struct(best)), max_iters)


def plot_stats(A, B, nb, ca, n_min, ssd, T, pairs):
    """Rows of align_chimera.plot for a best table: [R][9] float64 (see orc_plot_stats)."""
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    ssd = np.ascontiguousarray(ssd, np.float32)
    T = np.ascontiguousarray(T, np.float32)
    pairs = np.ascontiguousarray(pairs, np.int16)
    R, N = pairs.shape[0], pairs.shape[1]
    b = OrcBest(_fp(ssd), _fp(T), pairs.ctypes.data_as(C.POINTER(C.c_int16)), None)
    out = np.zeros((R, 9), np.float64)
    lib().orc_plot_stats(_fp(A), A.size // (3 * nb), _fp(B), B.size // (3 * nb), nb, ca, n_min, R, N,
                         C.byref(b), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def dp_cells(m, nB, n_min, n_max):
    return lib().orc_dp_cells(m, nB, n_min, n_max)


def read_ref_dump(path):
    """Parse a file written by oracle/ref_dump.cc."""
    buf = open(path, "rb").read()
    off = [0]

    def take(dt, count=1):
        a = np.frombuffer(buf, dt, count, off[0])
        off[0] += a.nbytes
        return a

    magic = int(take(np.int32)[0])
    assert magic == 0x4d524d53
    m, nB, nb, n_min, n_max = (int(x) for x in take(np.int32, 5))
    max_dist, max_rmsd = (float(x) for x in take(np.float32, 2))
    use_filter = int(take(np.int32)[0])
    of_dist = float(take(np.float32)[0])
    of_min = int(take(np.int32)[0])
    ir = int(take(np.int32)[0])
    refine = int(take(np.int32)[0])
    frag = int(take(np.int32)[0])
    count = int(take(np.int64)[0])
    stride = int(take(np.int32)[0])
    N = min(m, nB)
    R = n_max - n_min + 1
    d = dict(m=m, nB=nB, nb=nb, n_min=n_min, n_max=n_max, max_dist=max_dist, max_rmsd=max_rmsd,
             use_filter=use_filter, of_dist=of_dist, of_min=of_min, ir=ir, refine=refine,
             frag=frag, count=count, stride=stride, N=N, R=R)
    d["A"] = take(np.float32, m * nb * 3).reshape(-1, 3).copy()
    d["B"] = take(np.float32, nB * nb * 3).reshape(-1, 3).copy()
    d["orient"] = take(np.float32, count * 12).reshape(-1, 12).copy()
    if stride >= 0:
        good = np.zeros(count, np.int32)
        n_last = np.zeros(count, np.int32)
        cost = np.zeros((count, N), np.float32)
        valid = np.zeros((count, N), np.uint8)
        ssd = np.zeros((count, N), np.float32)
        detail = {}
        for o in range(count):
            good[o], n_last[o] = take(np.int32, 2)
            cost[o] = take(np.float32, N)
            valid[o] = take(np.uint8, N)
            ssd[o] = take(np.float32, N)
            if stride > 0 and o % stride == 0:
                T = take(np.float32, 12 * N).reshape(N, 12).copy()
                pairs = take(np.int16, 2 * N * N).reshape(N, N, 2).copy()
                detail[o] = (T, pairs)
        d.update(good=good, n_last=n_last, cost=cost, valid=valid, ssd=ssd, detail=detail)

    def best():
        ssd = np.zeros(R, np.float32)
        T = np.zeros((R, 12), np.float32)
        pairs = np.zeros((R, N, 2), np.int16)
        for r in range(R):
            ssd[r] = take(np.float32)[0]
            T[r] = take(np.float32, 12)
            pairs[r] = take(np.int16, 2 * N).reshape(N, 2)
        return dict(ssd=ssd, T=T, pairs=pairs)

    d["best_search"] = best()
    if refine:
        d["best_refine"] = best()
    assert off[0] == len(buf), (off[0], len(buf))
    return d


def params_from_dump(d):
    return OrcParams(d["nb"], d["n_min"], d["n_max"], d["max_dist"], d["max_rmsd"], d["use_filter"],
                     d["of_dist"], d["of_min"], d["ir"])
