"""ctypes binding of libwildcat_b200.so — the extern "C" surface declared in include/wildcat_b200.h.

There is no CPU fallback: if the CUDA library is missing or no device is present, every call fails loudly.
"""
import ctypes as C
import os
import re

from . import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("WC_LIB_OVERRIDE") or os.path.join(_HERE, "libwildcat_b200.so")  # override: A/B experiments of tools/
HEADER = os.path.join(_HERE, "..", "include", "wildcat_b200.h")
_lib = None


class WildcatError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: status {status} {detail}")


def declared_symbols():
    """every wc_* function the header declares (used by the CPU-side export test)."""
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"\b(wc_[a-z0-9_]+)\s*\(", txt)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise WildcatError(T.WC_ECUDA, "load", f"{SO_PATH} is missing: run `python -m wildcat_slam_b200.build` "
                           "(the product path has no CPU fallback)")
    lib = C.CDLL(SO_PATH)
    vp, sz, i32, dbl = C.c_void_p, C.c_size_t, C.c_int, C.c_double
    P = C.POINTER
    lib.wc_abi_version.restype = i32
    lib.wc_default_params.argtypes = [P(T.Params)]
    lib.wc_default_solve_opts.argtypes = [P(T.SolveOpts)]
    lib.wc_create.argtypes = [P(T.Params), i32, P(vp)]
    lib.wc_destroy.argtypes = [vp]
    lib.wc_last_error.argtypes = [vp]
    lib.wc_last_error.restype = C.c_char_p
    lib.wc_status_str.argtypes = [i32]
    lib.wc_status_str.restype = C.c_char_p
    lib.wc_host_alloc.argtypes = [sz]
    lib.wc_host_alloc.restype = vp
    lib.wc_host_free.argtypes = [vp]
    lib.wc_host_free.restype = None
    lib.wc_stream.argtypes = [vp]
    lib.wc_stream.restype = vp
    lib.wc_build_surfels.argtypes = [vp, vp, sz, vp, sz, P(sz), vp, P(dbl)]
    lib.wc_points_upload.argtypes = [vp, vp, sz]
    lib.wc_points_prefetch.argtypes = [vp, vp, sz, i32]
    lib.wc_comm_shard_upload.argtypes = [vp, i32]
    lib.wc_build_surfels_resident.argtypes = [vp, P(sz), P(dbl), P(dbl), P(dbl)]
    lib.wc_surfels_fetch.argtypes = [vp, vp, sz, P(sz)]
    lib.wc_default_sweep_filter.argtypes = [P(T.SweepFilter)]
    lib.wc_filter_points.argtypes = [vp, P(T.SweepFilter), vp, sz, vp, sz, P(sz)]
    lib.wc_undistort_sweep.argtypes = [vp, vp, sz, vp, sz, vp]
    lib.wc_undistort_upload.argtypes = [vp, vp, sz, vp, sz]
    lib.wc_surfel_markers.argtypes = [vp, vp, sz, vp]
    lib.wc_window_residuals.argtypes = [vp, P(T.SolveOpts), vp, vp, vp, sz, P(sz), vp, sz, P(sz)]
    lib.wc_unpack_pointcloud2.argtypes = [vp, vp, sz, P(T.Pc2Layout), vp]
    lib.wc_update_surfel_poses.argtypes = [vp, vp, sz, vp, sz]
    lib.wc_match.argtypes = [vp, vp, sz, vp, sz, i32, vp, sz, P(sz), vp, P(dbl)]
    lib.wc_knn6.argtypes = [vp, vp, sz, vp, sz, i32, vp, vp]
    win = [vp, sz, vp, sz, vp, sz, vp, sz, vp, sz, vp, sz]
    lib.wc_window_solve.argtypes = [vp, *win, P(T.SolveOpts), P(T.SolveSummary)]
    lib.wc_window_upload.argtypes = [vp, *win]
    lib.wc_window_solve_resident.argtypes = [vp, P(T.SolveOpts), P(T.SolveSummary), vp]
    lib.wc_window_evaluate.argtypes = [vp, *win, P(T.SolveOpts), P(dbl), vp, vp]
    lib.wc_spline_fit_eval.argtypes = [vp, vp, vp, sz, vp, sz, vp, vp]
    lib.wc_apply_corrections.argtypes = [vp, vp, sz, vp, sz]
    lib.wc_predict_states.argtypes = [vp, vp, sz, vp, vp, vp, dbl, dbl, sz, vp]
    lib.wc_pass_upload.argtypes = [vp, vp, sz, vp, sz, vp, sz]
    lib.wc_pass_upload_windows.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, i32]
    lib.wc_window_shrink.argtypes = [vp, dbl, dbl, i32, P(sz), P(sz)]
    lib.wc_windows_fetch.argtypes = [vp, vp, sz, P(sz), vp, sz, P(sz)]
    lib.wc_window_pass_resident.argtypes = [vp, P(T.SolveOpts), P(T.SolveSummary), vp, P(T.PassStats)]
    lib.wc_launch_count.argtypes = [vp]
    lib.wc_launch_count.restype = C.c_int64
    lib.wc_comm_export.argtypes = [vp, vp]
    lib.wc_comm_connect.argtypes = [vp, i32, i32, vp]
    lib.wc_comm_disconnect.argtypes = [vp]
    lib.wc_comm_bench.argtypes = [vp, sz, i32, P(dbl)]
    for name in declared_symbols():
        f = getattr(lib, name)  # raises AttributeError if the library lacks a declared entry point
        if name not in ("wc_abi_version", "wc_default_params", "wc_default_solve_opts", "wc_destroy", "wc_last_error",
                        "wc_status_str", "wc_stream", "wc_launch_count", "wc_host_alloc", "wc_host_free", "wc_default_sweep_filter"):
            f.restype = i32
    _lib = lib
    return lib
