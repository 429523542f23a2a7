"""ctypes binding of the C ABI declared in include/minrms_b200.h.  There is no CPU
fallback: if the CUDA library is missing the import of a compute entry point fails loudly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MINRMS_B200_LIB: another build of the same library (kernel A/B measurements)
LIB_PATH = os.environ.get("MINRMS_B200_LIB") or os.path.join(HERE, "libminrms_b200.so")


class MrParams(C.Structure):
    _fields_ = [("nb", C.c_int32), ("n_min", C.c_int32), ("n_max", C.c_int32),
                ("max_dist", C.c_float), ("max_rmsd", C.c_float),
                ("use_filter", C.c_int32), ("of_dist", C.c_float), ("of_min", C.c_int32),
                ("refine_while_searching", C.c_int32)]


class MrBest(C.Structure):
    _fields_ = [("ssd", C.POINTER(C.c_float)), ("T", C.POINTER(C.c_float)),
                ("pairs", C.POINTER(C.c_int16)), ("o_id", C.POINTER(C.c_int64))]


class MrStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("orientations", C.c_int64),
                ("orientations_dp", C.c_int64), ("dp_cells", C.c_int64), ("nw_cells", C.c_int64),
                ("dp_kernel_ms", C.c_double), ("dp_kernel_launches", C.c_int64),
                ("nw_kernel_ms", C.c_double), ("tb_kernel_ms", C.c_double),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("refine_batches", C.c_int64), ("refine_evals_launched", C.c_int64)]


EXPORTS = ["mr_create", "mr_destroy", "mr_last_error", "mr_best_init", "mr_search", "mr_search_scores", "mr_search_dissimilar",
           "mr_upload_orientations", "mr_search_resident", "mr_reset_best", "mr_fetch_best",
           "mr_fetch_keys", "mr_device_keys", "mr_fetch_rows_device", "mr_eval_one", "mr_refine", "mr_gen_orientations",
           "mr_superimpose_batch", "mr_plot_stats", "mr_get_stats", "mr_reset_stats", "mr_set_option", "mr_get_info",
           "mr_dp_cells_per_orientation", "mr_event_record", "mr_event_elapsed_ms"]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "minrms_b200: %s is missing -- build it with `python -m minrms_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    fp, vp = C.POINTER(C.c_float), C.c_void_p
    L.mr_create.restype = C.c_int
    L.mr_create.argtypes = [C.POINTER(vp), fp, C.c_int32, fp, C.c_int32, C.POINTER(MrParams), C.c_int32]
    L.mr_destroy.restype = None
    L.mr_destroy.argtypes = [vp]
    L.mr_last_error.restype = C.c_char_p
    L.mr_last_error.argtypes = [vp]
    L.mr_best_init.argtypes = [vp, C.POINTER(MrBest)]
    L.mr_search.argtypes = [vp, fp, C.c_int64, C.c_int64, C.POINTER(MrBest)]
    L.mr_search_scores.argtypes = [vp, fp, C.c_int64, C.c_int64, C.POINTER(MrBest), fp]
    L.mr_search_dissimilar.argtypes = [vp, fp, C.c_int64, C.c_int64, C.POINTER(MrBest), C.c_int32,
                                       C.POINTER(MrBest), C.c_int32]
    L.mr_upload_orientations.argtypes = [vp, fp, C.c_int64]
    L.mr_search_resident.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64]
    L.mr_reset_best.argtypes = [vp]
    L.mr_fetch_best.argtypes = [vp, C.POINTER(MrBest)]
    L.mr_fetch_keys.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mr_device_keys.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
    L.mr_fetch_rows_device.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
    L.mr_eval_one.argtypes = [vp, fp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), fp,
                              C.POINTER(C.c_uint8), fp, fp, C.POINTER(C.c_int16)]
    L.mr_refine.argtypes = [vp, C.POINTER(MrBest), C.c_int32, C.POINTER(C.c_int32)]
    L.mr_gen_orientations.argtypes = [vp, C.c_int32, C.c_float, C.POINTER(C.c_int32), C.c_int32,
                                      C.c_int32, fp, C.c_int64, C.POINTER(C.c_int64)]
    L.mr_superimpose_batch.argtypes = [vp, C.c_int32, C.c_int32, fp, fp, fp, fp]
    L.mr_plot_stats.argtypes = [vp, C.POINTER(MrBest), C.c_int32, C.POINTER(C.c_double)]
    L.mr_get_stats.argtypes = [vp, C.POINTER(MrStats)]
    L.mr_reset_stats.argtypes = [vp]
    L.mr_event_record.argtypes = [vp, C.c_int32]
    L.mr_event_elapsed_ms.argtypes = [vp, C.c_int32, C.c_int32, fp]
    L.mr_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.mr_get_info.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64)]
    L.mr_dp_cells_per_orientation.restype = C.c_int64
    L.mr_dp_cells_per_orientation.argtypes = [C.c_int32] * 4
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("mr_create",):
            fn.restype = C.c_int
    _lib = L
    return L
