"""Host-side mirror of the reference's src/odometry interface for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as the reference's C++ symbols (with CHECK aborts turned into
WildcatError carrying a wc_status):

  BuildSurfels(cloud)                        <- surfel_extraction.h:145-147
  UpdateSurfelPoses(imu_states, surfels)     <- lidar_odometry.cc:160-170
  KnnSurfelMatcher().BuildIndex / .Match     <- knn_surfel_matcher.h:17-19
  SolveWindow(...)                           <- Build*Residuals + ceres::Solve, lidar_odometry.cc:254-363,541-561
  CubicBSplineInterpolator(ts, pts).Interp   <- spline_interpolation.h:42-113
  ApplyCorrections(samples, imu)             <- UpdateImuPoses + UpdateSamplePoses, lidar_odometry.cc:172-215

Arrays are numpy structured arrays of the dtypes in wildcat_slam_b200.types (layout-identical to the reference's
structs).  Every call runs on the CUDA device owned by the Context; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import abi
from . import types as T


class Context:
    """wc_ctx: device memory, stream and (multi-GPU) peer mappings.  One per thread / per GPU."""

    def __init__(self, device=0, params=None):
        self.lib = abi.load()
        self.params = params or T.default_params()
        self._h = C.c_void_p()
        st = self.lib.wc_create(C.byref(self.params), int(device), C.byref(self._h))
        if st != T.WC_OK:
            raise abi.WildcatError(st, "wc_create", "(no CUDA device? the library has no CPU fallback)")
        self.device = device
        self._bufs = {}
        self._pinned = []

    def pinned(self, n, dtype):
        """numpy array of n records backed by page-locked memory from wc_host_alloc (freed with the context)"""
        dtype = np.dtype(dtype)
        nbytes = max(int(n), 1) * dtype.itemsize
        p = self.lib.wc_host_alloc(nbytes)
        if not p:
            raise abi.WildcatError(T.WC_ECUDA, "wc_host_alloc", f"{nbytes} bytes")
        self._pinned.append(p)
        raw = (C.c_uint8 * nbytes).from_address(p)
        return np.frombuffer(raw, dtype=dtype, count=max(int(n), 1))

    def buffer(self, name, n, dtype):
        """persistent pinned host buffer (grown on demand): no capacity-sized allocation + memset per call, and the
        D2H of results / H2D of arrays handed on to the next stage run at full PCIe rate"""
        b = self._bufs.get(name)
        if b is None or len(b) < n or b.dtype != np.dtype(dtype):
            b = self.pinned(max(n, 1), dtype)
            self._bufs[name] = b
        return b

    def close(self):
        if self._h:
            self._bufs = {}
            self.lib.wc_destroy(self._h)
            for p in self._pinned:
                self.lib.wc_host_free(p)
            self._pinned = []
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st, where):
        if st != T.WC_OK:
            raise abi.WildcatError(st, where, self.lib.wc_last_error(self._h).decode(errors="replace"))

    @property
    def handle(self):
        return self._h

    def stream(self):
        return self.lib.wc_stream(self._h)

    def prefetch(self, cloud, at_solve=False):
        """wc_points_prefetch: host-to-device copy of the NEXT sweep beside whatever runs meanwhile; the upload of the same
        array (ResidentSweep / ResidentPass / ResidentWindows) then finds it on the device.  at_solve: the copy starts when
        the next window pass reaches its solve stage (it would slow the memory-bound extraction and matching stages).
        `cloud` must be a contiguous POINT48 array — pinned (Context.pinned) for a truly asynchronous copy — left untouched
        until that upload."""
        if cloud.dtype != T.POINT48 or not cloud.flags["C_CONTIGUOUS"]:
            raise ValueError("prefetch needs a contiguous POINT48 array (the upload must see the same buffer)")
        self._prefetched = cloud  # keeps the buffer alive
        self.check(self.lib.wc_points_prefetch(self._h, T.ptr(cloud), len(cloud), 1 if at_solve else 0), "wc_points_prefetch")

    # ---- multi-GPU residual sharding -------------------------------------------------------------------------
    def comm_export(self):
        h = np.zeros(64, dtype=np.uint8)
        self.check(self.lib.wc_comm_export(self._h, T.ptr(h)), "wc_comm_export")
        return h

    def comm_connect(self, rank, world, all_handles):
        all_handles = np.ascontiguousarray(all_handles, dtype=np.uint8).reshape(world * 64)
        self.check(self.lib.wc_comm_connect(self._h, rank, world, T.ptr(all_handles)), "wc_comm_connect")

    def comm_shard_upload(self, on=True):
        """sweep uploads become collective: each rank copies its 1/world slab over PCIe, the rest comes from the peers"""
        self.check(self.lib.wc_comm_shard_upload(self._h, 1 if on else 0), "wc_comm_shard_upload")

    def comm_disconnect(self):
        self.check(self.lib.wc_comm_disconnect(self._h), "wc_comm_disconnect")


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def BuildSurfels(cloud, ctx=None, want_assign=False, timing=None, copy=True):
    """void BuildSurfels(const std::vector<hilti_ros::Point>&, std::deque<Surfel::Ptr>&, GlobalMap&).

    cloud: POINT48 array with non-decreasing time.  Returns the surfels sorted by timestamp (world frame), and the
    per-point (voxel, leaf) assignment when want_assign."""
    ctx = ctx or default_context()
    cloud = np.ascontiguousarray(cloud, dtype=T.POINT48)
    cap = int(ctx.params.max_surfels)
    out = ctx.buffer("surfels", cap, T.SURFEL)
    assign = np.zeros(len(cloud), dtype=T.ASSIGN) if want_assign else None
    n_out = C.c_size_t(0)
    ms = C.c_double(0)
    st = ctx.lib.wc_build_surfels(ctx.handle, T.ptr(cloud), len(cloud), T.ptr(out), cap, C.byref(n_out), T.ptr(assign),
                                  C.byref(ms))
    ctx.check(st, "wc_build_surfels")
    if timing is not None:
        timing["gpu_ms"] = ms.value
    # copy=False: a view of the context's pinned result buffer, valid until the next BuildSurfels on this context
    surfels = out[: n_out.value].copy() if copy else out[: n_out.value]
    return (surfels, assign) if want_assign else surfels


def UpdateSurfelPoses(imu_states, surfels, ctx=None, inplace=False):
    """UpdateSurfelPoses(const std::deque<ImuState>&, std::deque<Surfel::Ptr>&): returns the updated copy (or updates the
    given array in place, like the reference, with inplace=True)."""
    ctx = ctx or default_context()
    imu = np.ascontiguousarray(imu_states, dtype=T.IMU)
    s = np.ascontiguousarray(surfels, dtype=T.SURFEL)
    if not inplace:
        s = s.copy()
    ctx.check(ctx.lib.wc_update_surfel_poses(ctx.handle, T.ptr(imu), len(imu), T.ptr(s), len(s)), "wc_update_surfel_poses")
    return s


class KnnSurfelMatcher:
    """class KnnSurfelMatcher (knn_surfel_matcher.h:10-42)."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self.target_surfels_ = None

    def BuildIndex(self, surfels):
        if len(surfels) == 0:  # knn_surfel_matcher.cc:4-6
            return
        self.target_surfels_ = np.ascontiguousarray(surfels, dtype=T.SURFEL)

    def Match(self, surfels, timing=None, copy=True):
        """Returns (correspondences CORR[], first_is_target uint8[]).  When the query array is the array the index
        was built from, both indices address it (sliding-window matcher); otherwise s1/s2 are time ordered and
        first_is_target tells which array s1 indexes (fixed-window matcher: always the target)."""
        q = np.ascontiguousarray(surfels, dtype=T.SURFEL)
        if self.target_surfels_ is None or len(q) == 0:
            return np.zeros(0, T.CORR), np.zeros(0, np.uint8)
        t = self.target_surfels_
        self_match = int(t.ctypes.data == q.ctypes.data or (len(t) == len(q) and t.tobytes() == q.tobytes()))
        out = self.ctx.buffer("corr", len(q), T.CORR)
        fit = self.ctx.buffer("fit", len(q), np.dtype(np.uint8))
        n = C.c_size_t(0)
        ms = C.c_double(0)
        st = self.ctx.lib.wc_match(self.ctx.handle, T.ptr(q), len(q), T.ptr(t), len(t), self_match, T.ptr(out), len(out),
                                   C.byref(n), T.ptr(fit), C.byref(ms))
        self.ctx.check(st, "wc_match")
        if timing is not None:
            timing["gpu_ms"] = ms.value
        if not copy:  # views of the context's pinned buffers, valid until the next Match on this context
            return out[: n.value], fit[: n.value]
        return out[: n.value].copy(), fit[: n.value].copy()

    def KNearestSearchVectors(self, query6, target6, k=10):
        """FLANNBuildIndex + FLANNKNearestSearch on raw 6-vectors (knn_surfel_matcher_test.cc:19-43)."""
        q = np.ascontiguousarray(query6, dtype=np.float64)
        t = np.ascontiguousarray(target6, dtype=np.float64)
        idx = np.zeros((len(q), k), dtype=np.int32)
        d2 = np.zeros((len(q), k), dtype=np.float64)
        st = self.ctx.lib.wc_knn6(self.ctx.handle, T.ptr(q), len(q), T.ptr(t), len(t), k, T.ptr(idx), T.ptr(d2))
        self.ctx.check(st, "wc_knn6")
        return idx, d2


def _window_arrays(sld, fix, sld_corr, fix_corr, imu, samples):
    sld = np.ascontiguousarray(sld, dtype=T.SURFEL)
    fix = np.ascontiguousarray(fix if fix is not None else np.zeros(0, T.SURFEL), dtype=T.SURFEL)
    sc = np.ascontiguousarray(sld_corr if sld_corr is not None else np.zeros(0, T.CORR), dtype=T.CORR)
    fc = np.ascontiguousarray(fix_corr if fix_corr is not None else np.zeros(0, T.CORR), dtype=T.CORR)
    imu = np.ascontiguousarray(imu if imu is not None else np.zeros(0, T.IMU), dtype=T.IMU)
    smp = np.ascontiguousarray(samples, dtype=T.SAMPLE).copy()
    args = [T.ptr(sld), len(sld), T.ptr(fix), len(fix), T.ptr(sc), len(sc), T.ptr(fc), len(fc), T.ptr(imu), len(imu),
            T.ptr(smp), len(smp)]
    return (sld, fix, sc, fc, imu, smp), args


def SolveWindow(sld, fix, sld_corr, fix_corr, imu, samples, opts=None, ctx=None):
    """BuildSldWinLidarResiduals + BuildFixWinLidarResiduals + BuildImuResiduals + ceres::Solve.

    Returns (samples with data_cor overwritten by the solution, SolveSummary)."""
    ctx = ctx or default_context()
    o = opts or T.default_solve_opts()
    keep, args = _window_arrays(sld, fix, sld_corr, fix_corr, imu, samples)
    summ = T.SolveSummary()
    st = ctx.lib.wc_window_solve(ctx.handle, *args, C.byref(o), C.byref(summ))
    ctx.check(st, "wc_window_solve")
    return keep[5], summ


def EvaluateWindow(sld, fix, sld_corr, fix_corr, imu, samples, opts=None, ctx=None, want_jtj=True):
    """ceres::Problem::Evaluate-like hook: cost, gradient and J^T J of the robustified problem at samples.data_cor."""
    ctx = ctx or default_context()
    o = opts or T.default_solve_opts()
    keep, args = _window_arrays(sld, fix, sld_corr, fix_corr, imu, samples)
    n = 12 * len(keep[5])
    cost = C.c_double(0)
    grad = np.zeros(n)
    jtj = np.zeros((n, n)) if want_jtj else None
    st = ctx.lib.wc_window_evaluate(ctx.handle, *args, C.byref(o), C.byref(cost), T.ptr(grad), T.ptr(jtj))
    ctx.check(st, "wc_window_evaluate")
    return cost.value, grad, jtj


class ResidentWindow:
    """Upload a window once, then solve repeatedly from the uploaded starting point (benchmark / multi-GPU path)."""

    def __init__(self, sld, fix, sld_corr, fix_corr, imu, samples, ctx=None):
        self.ctx = ctx or default_context()
        self._keep, args = _window_arrays(sld, fix, sld_corr, fix_corr, imu, samples)
        self.K = len(self._keep[5])
        self.ctx.check(self.ctx.lib.wc_window_upload(self.ctx.handle, *args), "wc_window_upload")

    def solve(self, opts=None):
        o = opts or T.default_solve_opts()
        summ = T.SolveSummary()
        x = np.zeros((self.K, 12))
        st = self.ctx.lib.wc_window_solve_resident(self.ctx.handle, C.byref(o), C.byref(summ), T.ptr(x))
        self.ctx.check(st, "wc_window_solve_resident")
        return x, summ


class CubicBSplineInterpolator:
    """class CubicBSplineInterpolator (spline_interpolation.h:42-113); Interp returns None where the reference
    returns nullptr."""

    def __init__(self, timestamps, points, ctx=None):
        self.ctx = ctx or default_context()
        self.timestamps_ = np.ascontiguousarray(timestamps, dtype=np.float64)
        self.points_ = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        if len(self.timestamps_) != len(self.points_):  # CHECK_EQ :47
            raise abi.WildcatError(T.WC_EINVAL, "CubicBSplineInterpolator", "timestamps.size() != points.size()")

    def InterpMany(self, ts):
        q = np.ascontiguousarray(np.atleast_1d(ts), dtype=np.float64)
        out = np.zeros((len(q), 3))
        valid = np.zeros(len(q), dtype=np.uint8)
        st = self.ctx.lib.wc_spline_fit_eval(self.ctx.handle, T.ptr(self.timestamps_), T.ptr(self.points_), len(self.timestamps_),
                                             T.ptr(q), len(q), T.ptr(out), T.ptr(valid))
        self.ctx.check(st, "wc_spline_fit_eval")
        return out, valid.astype(bool)

    def Interp(self, timestamp):
        out, valid = self.InterpMany([timestamp])
        return out[0] if valid[0] else None


def ApplyCorrections(samples, imu_states, ctx=None):
    """UpdateImuPoses(sample_states, imu_states) then UpdateSamplePoses(sample_states): returns updated copies."""
    ctx = ctx or default_context()
    s = np.ascontiguousarray(samples, dtype=T.SAMPLE).copy()
    imu = np.ascontiguousarray(imu_states, dtype=T.IMU).copy()
    ctx.check(ctx.lib.wc_apply_corrections(ctx.handle, T.ptr(s), len(s), T.ptr(imu), len(imu)), "wc_apply_corrections")
    return s, imu


def PredictStates(imu_states, ba, bg, grav, t_last_sample, sample_dt, n_new, ctx=None):
    """Steps 2-3 of LidarOdometry::PredictImuStatesAndSampleStates (lidar_odometry.cc:403-453): forward-predict the poses
    of imu_states[2:] (PredictPoseOfNewImuState) and create n_new sample states at t_last_sample + i * sample_dt."""
    ctx = ctx or default_context()
    imu = np.ascontiguousarray(imu_states, dtype=T.IMU).copy()
    out = np.zeros(max(1, n_new), dtype=T.SAMPLE)
    v = [np.ascontiguousarray(a, dtype=np.float64) for a in (ba, bg, grav)]
    st = ctx.lib.wc_predict_states(ctx.handle, T.ptr(imu), len(imu), T.ptr(v[0]), T.ptr(v[1]), T.ptr(v[2]),
                                   C.c_double(t_last_sample), C.c_double(sample_dt), n_new, T.ptr(out))
    ctx.check(st, "wc_predict_states")
    return imu, out[:n_new].copy()


def FilterPoints(cloud, flt=None, ctx=None):
    """The per-point loop at the top of LidarOdometry::AddLidarScan (lidar_odometry.cc:489-496): lidar -> IMU extrinsic,
    range and blind-box filter; returns the kept points in order."""
    ctx = ctx or default_context()
    flt = flt or T.default_sweep_filter()
    cloud = np.ascontiguousarray(cloud, dtype=T.POINT48)
    out = np.zeros(max(1, len(cloud)), dtype=T.POINT48)
    n = C.c_size_t(0)
    st = ctx.lib.wc_filter_points(ctx.handle, C.byref(flt), T.ptr(cloud), len(cloud), T.ptr(out), len(out), C.byref(n))
    ctx.check(st, "wc_filter_points")
    return out[: n.value].copy()


def UndistortSweep(sweep_in, imu_states, ctx=None):
    """UndistortSweep (lidar_odometry.cc:143-158): every point to the world frame with the IMU pose interpolated at its
    own timestamp."""
    ctx = ctx or default_context()
    sweep_in = np.ascontiguousarray(sweep_in, dtype=T.POINT48)
    imu = np.ascontiguousarray(imu_states, dtype=T.IMU)
    out = np.zeros_like(sweep_in)
    st = ctx.lib.wc_undistort_sweep(ctx.handle, T.ptr(imu), len(imu), T.ptr(sweep_in), len(sweep_in), T.ptr(out))
    ctx.check(st, "wc_undistort_sweep")
    return out


class ResidentSweep:
    """Upload a sweep once, extract repeatedly (the "inputs resident in HBM" leg of the benchmark).  With `imu_states`
    the cloud is the RAW (distorted, IMU-frame) sweep and the undistortion runs on the device, fused with the repack
    into the resident layout (wc_undistort_upload)."""

    def __init__(self, cloud, ctx=None, imu_states=None):
        self.ctx = ctx or default_context()
        self.cloud = np.ascontiguousarray(cloud, dtype=T.POINT48)
        if imu_states is None:
            self.ctx.check(self.ctx.lib.wc_points_upload(self.ctx.handle, T.ptr(self.cloud), len(self.cloud)), "wc_points_upload")
        else:
            imu = np.ascontiguousarray(imu_states, dtype=T.IMU)
            st = self.ctx.lib.wc_undistort_upload(self.ctx.handle, T.ptr(imu), len(imu), T.ptr(self.cloud), len(self.cloud))
            self.ctx.check(st, "wc_undistort_upload")

    def extract(self):
        n = C.c_size_t(0)
        k, e, t = C.c_double(0), C.c_double(0), C.c_double(0)
        st = self.ctx.lib.wc_build_surfels_resident(self.ctx.handle, C.byref(n), C.byref(k), C.byref(e), C.byref(t))
        self.ctx.check(st, "wc_build_surfels_resident")
        return n.value, dict(keys_ms=k.value, emit_ms=e.value, total_ms=t.value)

    def fetch(self):
        cap = int(self.ctx.params.max_surfels)
        out = self.ctx.buffer("surfels", cap, T.SURFEL)
        n = C.c_size_t(0)
        self.ctx.check(self.ctx.lib.wc_surfels_fetch(self.ctx.handle, T.ptr(out), cap, C.byref(n)), "wc_surfels_fetch")
        return out[: n.value].copy()


class ResidentPass:
    """Device-resident window pass (steps 7-13 of AddLidarScan in one call): upload sweep + IMU/sample states + fixed
    window once, then run extract -> poses -> match x2 -> assemble -> solve repeatedly without host round trips."""

    def __init__(self, cloud, imu, samples, fix_body, ctx=None, keep_fix=False):
        """keep_fix: the fixed window uploaded by an earlier pass on this context stays where it is (it only changes at
        ShrinkToFit, lidar_odometry.cc:228-250); fix_body is ignored."""
        self.ctx = ctx or default_context()
        self.cloud = np.ascontiguousarray(cloud, dtype=T.POINT48)
        self.imu = np.ascontiguousarray(imu, dtype=T.IMU)
        self.samples = np.ascontiguousarray(samples, dtype=T.SAMPLE)
        self.fix = np.ascontiguousarray(fix_body if fix_body is not None else np.zeros(0, T.SURFEL), dtype=T.SURFEL)
        lib, h = self.ctx.lib, self.ctx.handle
        self.ctx.check(lib.wc_points_upload(h, T.ptr(self.cloud), len(self.cloud)), "wc_points_upload")
        if keep_fix:
            st = lib.wc_pass_upload_windows(h, T.ptr(self.imu), len(self.imu), T.ptr(self.samples), len(self.samples), None, 0, None, 0, 2)
            self.ctx.check(st, "wc_pass_upload_windows")
        else:
            self.ctx.check(lib.wc_pass_upload(h, T.ptr(self.imu), len(self.imu), T.ptr(self.samples), len(self.samples),
                                              T.ptr(self.fix), len(self.fix)), "wc_pass_upload")

    def run(self, opts=None):
        o = opts or T.default_solve_opts()
        summ, stats = T.SolveSummary(), T.PassStats()
        x = np.zeros((len(self.samples), 12))
        st = self.ctx.lib.wc_window_pass_resident(self.ctx.handle, C.byref(o), C.byref(summ), T.ptr(x), C.byref(stats))
        self.ctx.check(st, "wc_window_pass_resident")
        return x, summ, stats


class ResidentWindows:
    """surfels_sld_win_ / surfels_fix_win_ of LidarOdometry kept in HBM across sweeps (lidar_odometry.cc:527-528,228-250):
    AddSweep appends the new sweep's surfels to the sliding window and runs steps 8-13 over the whole window, ShrinkToFit
    moves the surfels that fell out of the sliding window to the front of the fixed window.  Only the new sweep, the IMU
    states and the sample states cross PCIe; the sample / IMU deques themselves (a few KB) stay with the caller."""

    WC_KEEP_SLD, WC_KEEP_FIX = 1, 2

    def __init__(self, ctx=None, fix_body=None, sld_body=None):
        self.ctx = ctx or default_context()
        self._fix0 = np.ascontiguousarray(fix_body if fix_body is not None else np.zeros(0, T.SURFEL), dtype=T.SURFEL)
        self._sld0 = np.ascontiguousarray(sld_body if sld_body is not None else np.zeros(0, T.SURFEL), dtype=T.SURFEL)
        self._resident = False
        self.n_sld, self.n_fix = len(self._sld0), len(self._fix0)

    def AddSweep(self, cloud, imu, samples, opts=None):
        lib, h = self.ctx.lib, self.ctx.handle
        cloud = np.ascontiguousarray(cloud, dtype=T.POINT48)
        imu = np.ascontiguousarray(imu, dtype=T.IMU)
        samples = np.ascontiguousarray(samples, dtype=T.SAMPLE)
        self.ctx.check(lib.wc_points_upload(h, T.ptr(cloud), len(cloud)), "wc_points_upload")
        keep = (self.WC_KEEP_SLD | self.WC_KEEP_FIX) if self._resident else 0
        st = lib.wc_pass_upload_windows(h, T.ptr(imu), len(imu), T.ptr(samples), len(samples), T.ptr(self._fix0), len(self._fix0),
                                        T.ptr(self._sld0), len(self._sld0), keep)
        self.ctx.check(st, "wc_pass_upload_windows")
        self._resident = True
        o = opts or T.default_solve_opts()
        summ, stats = T.SolveSummary(), T.PassStats()
        x = np.zeros((len(samples), 12))
        self.ctx.check(lib.wc_window_pass_resident(h, C.byref(o), C.byref(summ), T.ptr(x), C.byref(stats)), "wc_window_pass_resident")
        self.n_sld += int(stats.n_surfels)
        return x, summ, stats

    def ShrinkToFit(self, t_front_imu, fix_window_duration=20.0, trim_fixed=False):
        ns, nf = C.c_size_t(0), C.c_size_t(0)
        st = self.ctx.lib.wc_window_shrink(self.ctx.handle, C.c_double(t_front_imu), C.c_double(fix_window_duration), int(trim_fixed),
                                           C.byref(ns), C.byref(nf))
        self.ctx.check(st, "wc_window_shrink")
        self.n_sld, self.n_fix = ns.value, nf.value
        return self.n_sld, self.n_fix

    def Fetch(self):
        sld = np.zeros(max(1, self.n_sld), dtype=T.SURFEL)
        fix = np.zeros(max(1, self.n_fix), dtype=T.SURFEL)
        ns, nf = C.c_size_t(0), C.c_size_t(0)
        st = self.ctx.lib.wc_windows_fetch(self.ctx.handle, T.ptr(sld), len(sld), C.byref(ns), T.ptr(fix), len(fix), C.byref(nf))
        self.ctx.check(st, "wc_windows_fetch")
        return sld[: ns.value], fix[: nf.value]
