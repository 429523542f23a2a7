"""Host-side mirror of the reference's search interface over the C ABI.

Names follow the reference: Settings fields are those of SearchOrDyn3d::Settings /
OrientationGenerator::Settings (bin/minrms/search_or_dyn3d.h:159-215,
bin/minrms/search_or.h:149-308, lib/or_gen/or_gen.h), SearchOrCuda has the methods of
minrms::SearchOr (bin/minrms/search_or.h:57-130).  All compute happens in the CUDA
library; this file only marshals numpy arrays."""
import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _lib

UPPER_BOUND = np.float32(1.0e26)


class MinrmsError(RuntimeError):
    """Raised where the reference would ERR(...) -> exit(-1) (lib/global_utils/ERR.h:20-21)."""


class NoOrientationsError(MinrmsError):
    """the orientation generator has no fragment pair to fit (lib/or_gen/or_gen.cc:548-580)"""


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@dataclass
class Settings:
    nb: int = 1                                   # atoms per residue (" CA" by default)
    ca_index: int = 0                             # position of " CA" among those atoms (atoms_used.txt order)
    n_min: int = 1                                # -minN
    n_max: int = 0                                # -maxN (0 = min(m, n))
    dyn3d_max_dist: float = -1.0                  # -max-dist (NO_MAX_PHYS_DISTANCE = -1)
    dyn3d_max_rmsd: float = -1.0                  # -max-rmsd (dyn3d_NO_MAX_RMSD = -1)
    orientation_filter_nw: bool = False           # -of / -OF
    orientation_filter_nw_max_dist: float = 0.0
    orientation_filter_nw_min_num_matches: int = 0
    refine_while_searching: bool = True           # -ir / -IR
    refine_max_iters: int = -1                    # -r [iters]; ITERATE_UNTIL_CONVERGENCE = -1
    fm_fragment_size: int = 4                     # -fm L
    fm_cutoff_rmsd: float = -1.0                  # -fm L,rmsd (fm_NO_CUTOFF_RMSD)


def interpret_num_matches(x, m, n):
    """-minN / -maxN / -of k argument: values < 1 are fractions of min(m,n)
    (InterpretNumMatches, bin/minrms/minrms_parser.cc:1288-1327)."""
    if x < 1.0:
        return int(math.floor(x * min(m, n) + 0.5))
    return int(x)


def load_ca_pdb(path, atom_names=("CA",), chain=None):
    """Minimal fixed-column ATOM reader (the reference's PDB layer, lib/pdb++ and
    lib/biopolymer/load_pdb.cc, is out of scope; this feeds tests and the benchmark).
    Returns float32 [residues*nb, 3] with one point per (residue, atom name)."""
    residues, order = {}, []
    with open(path) as f:
        for line in f:
            if not line.startswith("ATOM"):
                continue
            name, ch = line[12:16].strip(), line[21]
            if chain is not None and ch != chain:
                continue
            key = (ch, line[22:27])
            if key not in residues:
                residues[key] = {}
                order.append(key)
            if name in atom_names and name not in residues[key]:
                residues[key][name] = (float(line[30:38]), float(line[38:46]), float(line[46:54]))
    pts = []
    for key in order:
        r = residues[key]
        if all(a in r for a in atom_names):
            pts.extend(r[a] for a in atom_names)
    return np.asarray(pts, np.float32).reshape(-1, 3)


def dp_cells_per_orientation(m, nB, n_min, n_max):
    return int(_lib.load().mr_dp_cells_per_orientation(m, nB, n_min, n_max))


class SearchOrCuda:
    """GPU replacement of SearchOrDyn3d (bin/minrms/search_or_dyn3d.{h,cc}).

    A, B: float32 [m*nb,3] / [n*nb,3] -- settings.aMol_f[0/1] flattened exactly like
    search_or_dyn3d.cc:186-224.  Results live in aSumSqdDists / aTransforms / aAlignments
    (the reference's aaSumSqdDists[0] etc.), slot r = n - n_min."""

    def __init__(self, A, B, settings, device=0):
        self.lib = _lib.load()
        self.A = np.ascontiguousarray(A, np.float32).reshape(-1, 3)
        self.B = np.ascontiguousarray(B, np.float32).reshape(-1, 3)
        s = self.settings = settings
        self.m = self.A.shape[0] // s.nb
        self.nB = self.B.shape[0] // s.nb
        self.N = min(self.m, self.nB)
        self.n_min = s.n_min
        self.n_max = s.n_max if s.n_max > 0 else self.N
        self.R = self.n_max - self.n_min + 1
        self.params = _lib.MrParams(s.nb, self.n_min, self.n_max, s.dyn3d_max_dist, s.dyn3d_max_rmsd,
                                    int(s.orientation_filter_nw), s.orientation_filter_nw_max_dist,
                                    s.orientation_filter_nw_min_num_matches,
                                    int(s.refine_while_searching))
        self.ctx = C.c_void_p()
        rc = self.lib.mr_create(C.byref(self.ctx), _fp(self.A), self.m, _fp(self.B), self.nB,
                                C.byref(self.params), device)
        if rc != 0:
            self.ctx = None
            raise MinrmsError(self.lib.mr_last_error(None).decode())
        R, N = max(self.R, 1), self.N
        self.aSumSqdDists = np.zeros(R, np.float32)
        self.aTransforms = np.zeros((R, 12), np.float32)
        self.aAlignments = np.full((R, N, 2), -1, np.int16)
        self.aOrientationIds = np.full(R, -1, np.int64)
        self._best = self._best_struct()
        self._check(self.lib.mr_best_init(self.ctx, C.byref(self._best)))

    # -- plumbing ---------------------------------------------------------------
    def _best_struct(self):
        return _lib.MrBest(_fp(self.aSumSqdDists), _fp(self.aTransforms),
                           self.aAlignments.ctypes.data_as(C.POINTER(C.c_int16)),
                           self.aOrientationIds.ctypes.data_as(C.POINTER(C.c_int64)))

    def _check(self, rc):
        if rc != 0:
            raise MinrmsError(self.lib.mr_last_error(self.ctx).decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.mr_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        self.close()

    def set_option(self, name, value):
        self._check(self.lib.mr_set_option(self.ctx, name.encode(), int(value)))

    def info(self, name):
        """what the context will run (mr_get_info): dp_path, dp_threads, dp_vp_bucket, dp_ctas_per_sm, ..."""
        import ctypes as C
        v = C.c_int64(0)
        self._check(self.lib.mr_get_info(self.ctx, name.encode(), C.byref(v)))
        return int(v.value)

    # -- SearchOr interface -----------------------------------------------------
    def Search(self, og, start=0, stop=None, first_id=None):
        """SearchOr::Search(og, start, stop) (search_or.h:87-104): FirstPass over
        og[start:stop]; ids are the positions in og, so disjoint ranges merge exactly
        (first_id: id of og[start] when several lists share one id space)."""
        orient = og.orientations if hasattr(og, "orientations") else np.asarray(og, np.float32)
        orient = np.ascontiguousarray(orient, np.float32).reshape(-1, 12)
        stop = orient.shape[0] if stop is None else stop
        part = np.ascontiguousarray(orient[start:stop])
        self._check(self.lib.mr_search(self.ctx, _fp(part), part.shape[0], start if first_id is None else first_id,
                                       C.byref(self._best)))

    def SearchScores(self, og, start=0, stop=None, first_id=None):
        """Search() that also returns the [stop - start, R] float32 array of every score FirstPass
        compares (mr_search_scores): what SearchOrDyn3d::KeepTrackOfAltSolutions is handed
        (search_or_dyn3d.cc:477-542, 1897-1904), UPPER_BOUND where the reference never gets to
        that size.  The alternate-solutions bookkeeping (-alt-o) starts from it."""
        orient = og.orientations if hasattr(og, "orientations") else np.asarray(og, np.float32)
        orient = np.ascontiguousarray(orient, np.float32).reshape(-1, 12)
        stop = orient.shape[0] if stop is None else stop
        part = np.ascontiguousarray(orient[start:stop])
        scores = np.full((part.shape[0], max(self.R, 1)), UPPER_BOUND, np.float32)
        self._check(self.lib.mr_search_scores(self.ctx, _fp(part), part.shape[0],
                                              start if first_id is None else first_id,
                                              C.byref(self._best), _fp(scores)))
        return scores[:, :self.R]

    def SearchDissimilar(self, og, lower, max_pairs_common, start=0, stop=None, first_id=None):
        """One pass of SearchOrDyn3d::MultiPass (-alt-r, search_or_dyn3d.cc:1975-2250) into this
        object's table: Search() in which an alignment of size n is recorded only if it shares at
        most max_pairs_common pairs with the size-n alignment of every table in `lower`
        (dicts with ssd [R] and pairs [R, N, 2], e.g. best_table() of the lower ranks)."""
        orient = og.orientations if hasattr(og, "orientations") else np.asarray(og, np.float32)
        orient = np.ascontiguousarray(orient, np.float32).reshape(-1, 12)
        stop = orient.shape[0] if stop is None else stop
        part = np.ascontiguousarray(orient[start:stop])
        keep = [(np.ascontiguousarray(t["ssd"], np.float32), np.ascontiguousarray(t["pairs"], np.int16)) for t in lower]
        arr = (_lib.MrBest * max(len(keep), 1))()
        for q, (ssd, pairs) in enumerate(keep):
            assert ssd.shape[0] >= self.R and pairs.shape == (max(self.R, 1), self.N, 2)
            arr[q] = _lib.MrBest(_fp(ssd), None, pairs.ctypes.data_as(C.POINTER(C.c_int16)), None)
        self._check(self.lib.mr_search_dissimilar(self.ctx, _fp(part), part.shape[0],
                                                  start if first_id is None else first_id,
                                                  C.byref(self._best), len(keep), arr, int(max_pairs_common)))

    def check_solutions(self):
        """the reference's final check in FirstPass (search_or_dyn3d.cc:1929-1941)"""
        if not np.any(self.aSumSqdDists[:self.R] < UPPER_BOUND):
            raise MinrmsError("Error: No solutions were found satisfying your criteria.")

    def Refine(self, max_iters=None):
        """SearchOr::Refine (search_or_dyn3d.cc:2261-2658); returns #evaluations.
        max_iters defaults to settings.refine_max_iters (-1: until nothing changes)."""
        n = C.c_int32(0)
        if max_iters is None:
            max_iters = self.settings.refine_max_iters
        self._check(self.lib.mr_refine(self.ctx, C.byref(self._best), max_iters, C.byref(n)))
        return n.value

    # -- device-resident variant (kernel timing, multi-GPU reduction) -----------
    def upload_orientations(self, orient):
        orient = np.ascontiguousarray(orient, np.float32).reshape(-1, 12)
        self._check(self.lib.mr_upload_orientations(self.ctx, _fp(orient), orient.shape[0]))

    def search_resident(self, first, count, first_id):
        self._check(self.lib.mr_search_resident(self.ctx, first, count, first_id))

    def reset_best(self):
        self._check(self.lib.mr_reset_best(self.ctx))
        self._check(self.lib.mr_best_init(self.ctx, C.byref(self._best)))

    def fetch_best(self):
        self._check(self.lib.mr_fetch_best(self.ctx, C.byref(self._best)))

    def fetch_keys(self):
        keys = np.zeros(max(self.R, 1), np.uint64)
        self._check(self.lib.mr_fetch_keys(self.ctx, keys.ctypes.data_as(C.POINTER(C.c_uint64))))
        return keys[:self.R]

    def device_keys(self):
        """(device address, R) of the per-size keys; also snapshots them (mr_device_keys)"""
        ptr, n = C.c_void_p(), C.c_int32(0)
        self._check(self.lib.mr_device_keys(self.ctx, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def fetch_rows_device(self):
        """(device address, int32 per row) of the payload rows of the sizes this context owns
        after the cross-rank MIN over device_keys() (mr_fetch_rows_device)"""
        ptr, n = C.c_void_p(), C.c_int32(0)
        self._check(self.lib.mr_fetch_rows_device(self.ctx, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def load_table(self, table):
        """make `table` (dict ssd, T, pairs, o_id) the stored best table"""
        self.aSumSqdDists[:self.R] = table["ssd"]
        self.aTransforms[:self.R] = table["T"]
        self.aAlignments[:self.R] = table["pairs"]
        self.aOrientationIds[:self.R] = table["o_id"]

    # -- single orientation (Refine building block, parity tests) ---------------
    def eval_one(self, T12, apply_filter=True):
        N = self.N
        T12 = np.ascontiguousarray(T12, np.float32).reshape(12)
        res = dict(cost=np.zeros(N, np.float32), valid=np.zeros(N, np.uint8),
                   ssd=np.zeros(N, np.float32), T=np.zeros((N, 12), np.float32),
                   pairs=np.zeros((N, N, 2), np.int16))
        nl, nw = C.c_int32(0), C.c_int32(0)
        self._check(self.lib.mr_eval_one(
            self.ctx, _fp(T12), int(apply_filter), C.byref(nl), C.byref(nw), _fp(res["cost"]),
            res["valid"].ctypes.data_as(C.POINTER(C.c_uint8)), _fp(res["ssd"]), _fp(res["T"]),
            res["pairs"].ctypes.data_as(C.POINTER(C.c_int16))))
        res["n_last"], res["nw_matches"] = nl.value, nw.value
        return res

    def event_record(self, slot):
        self._check(self.lib.mr_event_record(self.ctx, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float(0)
        self._check(self.lib.mr_event_elapsed_ms(self.ctx, a, b, C.byref(ms)))
        return ms.value

    def stats(self, reset=False):
        st = _lib.MrStats()
        self._check(self.lib.mr_get_stats(self.ctx, C.byref(st)))
        if reset:
            self._check(self.lib.mr_reset_stats(self.ctx))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def best_table(self):
        return dict(ssd=self.aSumSqdDists[:self.R].copy(), T=self.aTransforms[:self.R].copy(),
                    pairs=self.aAlignments[:self.R].copy(), o_id=self.aOrientationIds[:self.R].copy())

    # -- output statistics -----------------------------------
    def plot_stats(self, table=None):
        """Rows of align_chimera.plot for the stored table (or `table` = dict(ssd, T, pairs)):
        float64 [R, 9] = {has solution, rmsd, longest distance, S_str, mu_str, delta_str, Z,
        probability, -log10(probability)} (SearchOrDyn3d::WritePlotFile,
        search_or_dyn3d.cc:2703-2760)."""
        best = self._best
        if table is not None:
            keep = [np.ascontiguousarray(table["ssd"], np.float32), np.ascontiguousarray(table["T"], np.float32),
                    np.ascontiguousarray(table["pairs"], np.int16)]
            best = _lib.MrBest(_fp(keep[0]), _fp(keep[1]), keep[2].ctypes.data_as(C.POINTER(C.c_int16)), None)
        rows = np.zeros((max(self.R, 1), 9), np.float64)
        self._check(self.lib.mr_plot_stats(self.ctx, C.byref(best), self.settings.ca_index,
                                           rows.ctypes.data_as(C.POINTER(C.c_double))))
        return rows[:self.R]

    def WritePlotFile(self, path="align_chimera.plot"):
        """The file SearchOr::WriteChimeraInfo -> WritePlotFile writes (search_or_dyn3d.cc:2662-2760);
        numbers are formatted like the reference's ostream (6 significant digits)."""
        rows = self.plot_stats()
        with open(path, "w") as f:
            f.write("# Matches,RMSD,Longest Distance,-log(probability)\n")
            for r in range(self.R):
                if rows[r, 0] == 0.0:
                    continue
                f.write("%d %g %g %g\n" % (self.n_min + r, rows[r, 1], rows[r, 2], rows[r, 8]))

    def rmsd(self, n):
        """RMSD of the stored size-n alignment as WriteMSF prints it (search_or_dyn3d.cc:2878)."""
        return float(np.sqrt(self.aSumSqdDists[n - self.n_min] / np.float32(n * self.settings.nb)))


class OrientationGenerator:
    """lib/or_gen/or_gen.{h,cc}: all candidate orientations, precomputed and shuffled.
    og[i] is a 12-float row-major Matrix3x4; len(og) = og.size()."""

    def __init__(self, search, intervals=None, shuffle=True):
        s = search.settings
        lib = search.lib
        iv, niv = None, 0
        if intervals is not None and len(intervals):
            self._iv = np.ascontiguousarray(intervals, np.int32).reshape(-1, 4)
            iv, niv = self._iv.ctypes.data_as(C.POINTER(C.c_int32)), self._iv.shape[0]
        cnt = C.c_int64(0)
        if s.fm_fragment_size <= search.N:   # a longer fragment fits nowhere: count stays 0
            search._check(lib.mr_gen_orientations(search.ctx, s.fm_fragment_size, s.fm_cutoff_rmsd, iv, niv,
                                                  int(shuffle), None, 0, C.byref(cnt)))
        if cnt.value == 0:
            raise NoOrientationsError("Error: there are no valid regions from which to draw the "
                              "%d residue blocks (lib/or_gen/or_gen.cc:548-580)" % s.fm_fragment_size)
        self.orientations = np.zeros((cnt.value, 12), np.float32)
        search._check(lib.mr_gen_orientations(search.ctx, s.fm_fragment_size, s.fm_cutoff_rmsd, iv, niv,
                                              int(shuffle), _fp(self.orientations), cnt.value,
                                              C.byref(cnt)))

    def size(self):
        return self.orientations.shape[0]

    __len__ = size

    def __getitem__(self, i):
        return self.orientations[i]


def superimpose_batch(search, fixed, moveable):
    """Superimpose::FindTransformMinRMSD for a batch [batch, np, 3] on the GPU."""
    fixed = np.ascontiguousarray(fixed, np.float32)
    moveable = np.ascontiguousarray(moveable, np.float32)
    batch, npts = fixed.shape[0], fixed.shape[1]
    T = np.zeros((batch, 12), np.float32)
    r = np.zeros(batch, np.float32)
    search._check(search.lib.mr_superimpose_batch(search.ctx, batch, npts, _fp(fixed), _fp(moveable),
                                                  _fp(T), _fp(r)))
    return r, T
