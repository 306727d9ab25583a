"""Observability outputs of the reference on top of the device kernels:

  * Histogram            common/histogram.{h,cc}: Add / ToString(buckets), float arithmetic like the original
  * WindowResiduals      PrintSurfelResiduals + PrintImuResiduals (lidar_odometry.cc:56-93): residuals after the loss
                         corrector of the uploaded window, evaluated on the device (wc_window_residuals)
  * SurfelMarkers        PubSurfels (surfel_extraction.cc:360-417): marker pose / scale / colour per surfel
  * pose_stamped         the transform broadcast at lidar_odometry.cc:596-601 (last sample state)
"""
import ctypes as C

import numpy as np

from . import types as T


class Histogram:
    def __init__(self, values=None):
        self.values_ = [] if values is None else [float(v) for v in values]

    def Add(self, value):
        self.values_.append(float(value))

    def ToString(self, buckets):
        assert buckets >= 1  # CHECK_GE(buckets, 1)
        if not self.values_:
            return "Count: 0"
        f32 = np.float32
        v = np.asarray(self.values_, dtype=np.float64)
        vmin, vmax = f32(v.min()), f32(v.max())
        acc = f32(0.0)
        for x in v:  # std::accumulate(..., 0.f): float accumulator
            acc = f32(acc + f32(x))
        mean = f32(acc / f32(len(v)))
        out = f"Count: {len(v)}  Min: {_g(vmin)}  Max: {_g(vmax)}  Mean: {_g(mean)}"
        if vmin == vmax:
            return out
        vf = v.astype(np.float32)
        lower, total = vmin, 0
        for i in range(buckets):
            last = i + 1 == buckets
            upper = vmax if last else f32(f32(vmax * f32(i + 1)) / f32(buckets) + f32(vmin * f32(buckets - i - 1)) / f32(buckets))
            count = int(np.count_nonzero((lower <= vf) & ((vf <= upper) if last else (vf < upper))))
            total += count
            out += "\n[%f, %f%c" % (lower, upper, "]" if last else ")")
            bar = (count * 20 + len(v) // 2) // len(v)
            out += "\t" + " " * (20 - bar) + "#" * bar
            out += f"\tCount: {count} ({_g(f32(count * f32(1e2)) / f32(len(v)))}%)\tTotal: {total} ({_g(f32(total * f32(1e2)) / f32(len(v)))}%)"
            lower = upper
        return out


def _g(x):
    """absl::StrCat(float): shortest representation with 6 significant digits (%g)"""
    return "%g" % float(x)


def WindowResiduals(ctx, K, data_cor=None, opts=None, n_lidar_cap=None, n_imu_cap=None):
    """Residuals of the window last uploaded to `ctx` (wc_window_upload / ResidentWindow / SolveWindow), evaluated at
    data_cor (K x 12; None: the uploaded starting point).  Returns dict(sld=..., fix=..., imu=(n, 12))."""
    o = opts or T.default_solve_opts()
    cap = int(n_lidar_cap if n_lidar_cap is not None else ctx.params.max_corrs)
    icap = int(n_imu_cap if n_imu_cap is not None else ctx.params.max_imu_states)
    res = np.zeros(max(1, cap))
    fixf = np.zeros(max(1, cap), dtype=np.uint8)
    ires = np.zeros((max(1, icap), 12))
    n, nb = C.c_size_t(0), C.c_size_t(0)
    x = None if data_cor is None else np.ascontiguousarray(data_cor, dtype=np.float64).reshape(K * 12)
    st = ctx.lib.wc_window_residuals(ctx.handle, C.byref(o), T.ptr(x), T.ptr(res), T.ptr(fixf), cap, C.byref(n), T.ptr(ires), icap, C.byref(nb))
    ctx.check(st, "wc_window_residuals")
    res, fixf = res[: n.value], fixf[: n.value].astype(bool)
    return dict(sld=res[~fixf].copy(), fix=res[fixf].copy(), imu=ires[: nb.value].copy())


def residual_report(r, buckets=10):
    """the LOG(INFO) lines of PrintSurfelResiduals / PrintImuResiduals (costs: 1/2 sum of squared corrected residuals)"""
    lines = []
    for name, v in (("Sliding window", r["sld"]), ("Fixed window", r["fix"])):
        if len(v):
            lines.append(f"{name} Surfel residuals, cost: {_g(0.5 * float(np.sum(v * v)))}, dist: {Histogram(v).ToString(buckets)}")
    if len(r["imu"]):
        cost = 0.5 * float(np.sum(r["imu"] ** 2))
        for j, typ in enumerate(("gyro", "acc", "gyro_bias", "acc_bias")):
            h = Histogram(np.linalg.norm(r["imu"][:, 3 * j:3 * j + 3], axis=1))
            lines.append(f"Imu residuals with type {typ}, cost: {_g(cost)}, dist: {h.ToString(buckets)}")
    return "\n".join(lines)


def SurfelMarkers(surfels, ctx=None):
    from . import odometry as od

    ctx = ctx or od.default_context()
    s = np.ascontiguousarray(surfels, dtype=T.SURFEL)
    out = np.zeros(max(1, len(s)), dtype=T.MARKER)
    ctx.check(ctx.lib.wc_surfel_markers(ctx.handle, T.ptr(s), len(s), T.ptr(out)), "wc_surfel_markers")
    return out[: len(s)]


def pose_stamped(samples):
    """world -> imu_link transform of the newest sample state (lidar_odometry.cc:596-601)"""
    s = samples[-1]
    return dict(stamp=float(s["timestamp"]), frame_id="world", child_frame_id="imu_link", translation=s["pos"].copy(), rotation=s["rot"].copy())
