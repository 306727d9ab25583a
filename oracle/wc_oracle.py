"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/libwc_oracle.so (the CPU restatement of the reference's hot path).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (wildcat_slam_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from wildcat_slam_b200 import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libwc_oracle.so")
_lib = None


def build(force=False):
    """Compile the restatement (g++ only; no GPU, no reference sources needed)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cc", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "wildcat_b200.h"))
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.wco_build_surfels.restype = C.c_int64
        _lib.wco_match.restype = C.c_int64
        _lib.wco_cubic_bspline_approx.restype = C.c_double
        _lib.wco_cubic_bspline_approx.argtypes = [C.c_double] * 5
        _lib.wco_cubic_spline_interpolate.restype = C.c_double
        _lib.wco_cubic_spline_interpolate.argtypes = [C.c_double] * 9
    return _lib


def _p(a):
    return T.ptr(a)


def build_surfels(points, params=None, want_assign=False, want_info=False, rel_margin=1e-6):
    """BuildSurfels (surfel_extraction.cc:316-337).  Returns dict(surfels, assign, info, near_threshold)."""
    prm = params or T.default_params()
    points = np.ascontiguousarray(points, dtype=T.POINT48)
    n = len(points)
    cap = max(1024, n)
    out = np.zeros(cap, dtype=T.SURFEL)
    assign = np.zeros(n, dtype=T.ASSIGN) if want_assign else None
    info_dt = np.dtype([("n_points", "<i4"), ("layer", "<i4"), ("evals", "<f8", (3,)), ("likeness", "<f8")])
    info = np.zeros(cap, dtype=info_dt) if want_info else None
    near = C.c_int64(0)
    ns = lib().wco_build_surfels(C.byref(prm), _p(points), C.c_int64(n), _p(out), C.c_int64(cap), _p(assign), _p(info),
                                 C.c_double(rel_margin), C.byref(near))
    assert ns >= 0
    return dict(surfels=out[:ns].copy(), assign=assign, info=None if info is None else info[:ns].copy(),
                near_threshold=near.value)


def update_surfel_poses(imu, surfels):
    imu = np.ascontiguousarray(imu, dtype=T.IMU)
    surfels = np.ascontiguousarray(surfels, dtype=T.SURFEL).copy()
    st = lib().wco_update_surfel_poses(_p(imu), C.c_int64(len(imu)), _p(surfels), C.c_int64(len(surfels)))
    return st, surfels


def undistort_sweep(imu, points):
    imu = np.ascontiguousarray(imu, dtype=T.IMU)
    points = np.ascontiguousarray(points, dtype=T.POINT48)
    out = np.zeros_like(points)
    st = lib().wco_undistort_sweep(_p(imu), C.c_int64(len(imu)), _p(points), C.c_int64(len(points)), _p(out))
    return st, out


def predict_states(imu, ba, bg, grav, t_last_sample, sample_dt, n_new):
    imu = np.ascontiguousarray(imu, dtype=T.IMU).copy()
    out = np.zeros(max(1, n_new), dtype=T.SAMPLE)
    v = [np.ascontiguousarray(a, dtype=np.float64) for a in (ba, bg, grav)]
    st = lib().wco_predict_states(_p(imu), C.c_int64(len(imu)), _p(v[0]), _p(v[1]), _p(v[2]), C.c_double(t_last_sample),
                                  C.c_double(sample_dt), C.c_int64(n_new), _p(out))
    return st, imu, out[:n_new].copy()


def filter_points(points, flt=None):
    flt = flt or T.default_sweep_filter()
    points = np.ascontiguousarray(points, dtype=T.POINT48)
    out = np.zeros_like(points)
    f = lib().wco_filter_points
    f.restype = C.c_int64
    n = f(C.byref(flt), _p(points), C.c_int64(len(points)), _p(out))
    return (int(-n), None) if n < 0 else (0, out[:n].copy())


def match(query, target, self_match, params=None, use_kdtree=True):
    prm = params or T.default_params()
    query = np.ascontiguousarray(query, dtype=T.SURFEL)
    target = np.ascontiguousarray(target, dtype=T.SURFEL)
    out = np.zeros(max(1, len(query)), dtype=T.CORR)
    fit = np.zeros(max(1, len(query)), dtype=np.uint8)
    n = lib().wco_match(C.byref(prm), _p(query), C.c_int64(len(query)), _p(target), C.c_int64(len(target)),
                        C.c_int(int(self_match)), C.c_int(int(use_kdtree)), _p(out), _p(fit))
    return out[:n].copy(), fit[:n].copy()


def knn6(query6, target6, k, use_kdtree=True):
    query6 = np.ascontiguousarray(query6, dtype=np.float64)
    target6 = np.ascontiguousarray(target6, dtype=np.float64)
    nq = len(query6)
    idx = np.zeros((nq, k), dtype=np.int32)
    d2 = np.zeros((nq, k), dtype=np.float64)
    lib().wco_knn6(_p(query6), C.c_int64(nq), _p(target6), C.c_int64(len(target6)), C.c_int(k), C.c_int(int(use_kdtree)),
                   _p(idx), _p(d2))
    return idx, d2


def _window_args(sld, fix, sld_corr, fix_corr, imu, samples):
    sld = np.ascontiguousarray(sld, dtype=T.SURFEL)
    fix = np.ascontiguousarray(fix if fix is not None else np.zeros(0, T.SURFEL), dtype=T.SURFEL)
    sld_corr = np.ascontiguousarray(sld_corr if sld_corr is not None else np.zeros(0, T.CORR), dtype=T.CORR)
    fix_corr = np.ascontiguousarray(fix_corr if fix_corr is not None else np.zeros(0, T.CORR), dtype=T.CORR)
    imu = np.ascontiguousarray(imu if imu is not None else np.zeros(0, T.IMU), dtype=T.IMU)
    samples = np.ascontiguousarray(samples, dtype=T.SAMPLE).copy()
    keep = (sld, fix, sld_corr, fix_corr, imu, samples)
    args = [_p(sld), C.c_int64(len(sld)), _p(fix), C.c_int64(len(fix)), _p(sld_corr), C.c_int64(len(sld_corr)),
            _p(fix_corr), C.c_int64(len(fix_corr)), _p(imu), C.c_int64(len(imu)), _p(samples), C.c_int64(len(samples))]
    return keep, args, samples


def window_evaluate(sld, fix, sld_corr, fix_corr, imu, samples, params=None, opts=None, want_jtj=True):
    prm, o = params or T.default_params(), opts or T.default_solve_opts()
    keep, args, samples = _window_args(sld, fix, sld_corr, fix_corr, imu, samples)
    n = 12 * len(samples)
    cost = C.c_double(0)
    grad = np.zeros(n)
    jtj = np.zeros((n, n)) if want_jtj else None
    st = lib().wco_window_evaluate(C.byref(prm), C.byref(o), *args, C.byref(cost), _p(grad), _p(jtj))
    return st, cost.value, grad, jtj


def window_solve(sld, fix, sld_corr, fix_corr, imu, samples, params=None, opts=None):
    prm, o = params or T.default_params(), opts or T.default_solve_opts()
    keep, args, samples = _window_args(sld, fix, sld_corr, fix_corr, imu, samples)
    summ = T.SolveSummary()
    st = lib().wco_window_solve(C.byref(prm), C.byref(o), *args, C.byref(summ))
    return st, samples, summ


def window_residuals(sld, fix, sld_corr, fix_corr, imu, samples, params=None, opts=None):
    """PrintSurfelResiduals / PrintImuResiduals numbers: (status, sliding residuals, fixed residuals, imu residuals (n, 12))."""
    prm, o = params or T.default_params(), opts or T.default_solve_opts()
    keep, args, samples = _window_args(sld, fix, sld_corr, fix_corr, imu, samples)
    ns, nf = len(keep[2]), len(keep[3])
    lid = np.zeros(max(1, ns + nf))
    ires = np.zeros((max(1, len(keep[4])), 12))
    nb = C.c_int64(0)
    st = lib().wco_window_residuals(C.byref(prm), C.byref(o), *args, _p(lid), _p(ires), C.byref(nb))
    return st, lid[:ns].copy(), lid[ns:ns + nf].copy(), ires[:nb.value].copy()


def surfel_markers(surfels):
    surfels = np.ascontiguousarray(surfels, dtype=T.SURFEL)
    out = np.zeros(max(1, len(surfels)), dtype=T.MARKER)
    lib().wco_surfel_markers(_p(surfels), C.c_int64(len(surfels)), _p(out))
    return out[:len(surfels)].copy()


def lidar_factor(s1, s2, unary, sample_ts, x, params=None, jacobian_mode=T.WC_JAC_REFERENCE_OVERWRITE):
    prm = params or T.default_params()
    s1 = np.ascontiguousarray(s1, dtype=T.SURFEL).reshape(1)
    s2 = np.ascontiguousarray(s2, dtype=T.SURFEL).reshape(1)
    ts = np.ascontiguousarray(sample_ts, dtype=np.float64)
    K = len(ts)
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(K * 12)
    r = C.c_double(0)
    w = C.c_double(0)
    jac = np.zeros(12 * K)
    nrm = np.zeros(3)
    st = lib().wco_lidar_factor(C.byref(prm), C.c_int(jacobian_mode), _p(s1), _p(s2), C.c_int(int(unary)), _p(ts), None,
                                _p(x), C.c_int64(K), C.byref(r), _p(jac), C.byref(w), _p(nrm))
    return st, r.value, jac, w.value, nrm


def imu_factor(i3, sample_ts, mode, grav, x, params=None):
    prm = params or T.default_params()
    i3 = np.ascontiguousarray(i3, dtype=T.IMU)
    ts = np.ascontiguousarray(sample_ts, dtype=np.float64)
    nblk = 3 if mode == 0 else 2
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(nblk * 12)
    grav = np.ascontiguousarray(grav, dtype=np.float64)
    res = np.zeros(12)
    jac = np.zeros((12, 12 * nblk))
    st = lib().wco_imu_factor(C.byref(prm), _p(i3), _p(ts), C.c_int(mode), _p(grav), _p(x), _p(res), _p(jac))
    return st, res, jac


def spline_fit_eval(ts, pts3, query_t):
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    pts3 = np.ascontiguousarray(pts3, dtype=np.float64)
    q = np.ascontiguousarray(query_t, dtype=np.float64)
    out = np.zeros((len(q), 3))
    valid = np.zeros(len(q), dtype=np.uint8)
    ctrl = np.zeros((len(ts), 3))
    lib().wco_spline_fit_eval(_p(ts), _p(pts3), C.c_int64(len(ts)), _p(q), C.c_int64(len(q)), _p(out), _p(valid), _p(ctrl))
    return out, valid.astype(bool), ctrl


def apply_corrections(samples, imu):
    samples = np.ascontiguousarray(samples, dtype=T.SAMPLE).copy()
    imu = np.ascontiguousarray(imu, dtype=T.IMU).copy()
    st = lib().wco_apply_corrections(_p(samples), C.c_int64(len(samples)), _p(imu), C.c_int64(len(imu)))
    return st, samples, imu


def so3(op, v):
    names = {"exp": (0, 4), "log": (1, 3), "jl": (2, 9), "jl_inv": (3, 9), "jr": (4, 9), "jr_inv": (5, 9), "eig": (6, 12)}
    code, nout = names[op]
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(nout)
    lib().wco_so3(C.c_int(code), _p(v), _p(out))
    return out
