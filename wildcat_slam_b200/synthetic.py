"""Deterministic synthetic lidar/IMU windows of the BASELINE.json shapes.

Counter-based RNG (splitmix64 of seed/stream/index -> uniform -> Box-Muller), so every array is a pure
function of (config, seed) and the same on every rank and every box.  Nothing here reads /root/reference.

Scene: axis-aligned room + 8 seeded vertical panels.  Sensor: spinning lidar, R rings, A azimuth steps per
revolution, 10 rev/s, ring-fastest firing order, strictly increasing timestamps (lidar_odometry.cc:491).
Truth trajectory analytic; the *prior* the window starts from is IMU dead reckoning with biased, noisy
measurements through the reference's own prediction scheme (lidar_odometry.cc:112-123), and the sweep is
"undistorted" into the world with that prior exactly like UndistortSweep (lidar_odometry.cc:143-158).
"""
from dataclasses import dataclass, field

import numpy as np

from . import types as T

GRAV = np.array([0.0, 0.0, -9.81])


# ----------------------------------------------------------------------------------------------- RNG
def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform(seed, stream, idx):
    """U(0,1) for counters idx (uint64 array)."""
    with np.errstate(over="ignore"):
        key = np.uint64(seed) ^ (np.uint64(stream) << np.uint64(40))
        z = _splitmix64(_splitmix64(np.asarray(idx, dtype=np.uint64) ^ key))
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def normal(seed, stream, idx):
    u1 = uniform(seed, 2 * stream, idx)
    u2 = uniform(seed, 2 * stream + 1, idx)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


# ----------------------------------------------------------------------------------------- SO(3), numpy
def so3_exp(w):
    """rotation vectors (n,3) -> quaternions (n,4) xyzw."""
    w = np.atleast_2d(w)
    th = np.linalg.norm(w, axis=1)
    half = 0.5 * th
    small = th < 1e-10
    k = np.where(small, 0.5 - th * th / 48.0, np.sin(half) / np.where(small, 1.0, th))
    return np.concatenate([w * k[:, None], np.cos(half)[:, None]], axis=1)


def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack(
        [aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
         aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def quat_conj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def quat_rotate(q, v):
    u = q[..., :3]
    uv = 2.0 * np.cross(u, v)
    return v + q[..., 3:4] * uv + np.cross(u, uv)


def quat_log(q):
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    n = np.linalg.norm(q[..., :3], axis=-1)
    w = q[..., 3]
    ang = np.where(w < 0, np.arctan2(-n, -w), np.arctan2(n, w))
    k = np.where(n < 1e-10, 2.0 / np.where(w == 0, 1.0, w), 2.0 * ang / np.where(n < 1e-10, 1.0, n))
    return q[..., :3] * k[..., None]


def quat_slerp(a, t, b):
    """Eigen slerp, vectorised (lidar_odometry.cc:153)."""
    d = np.sum(a * b, axis=-1)
    ad = np.abs(d)
    lin = ad >= 1.0 - np.finfo(np.float64).eps
    th = np.arccos(np.clip(ad, -1.0, 1.0))
    st = np.where(lin, 1.0, np.sin(th))
    s0 = np.where(lin, 1.0 - t, np.sin((1.0 - t) * th) / st)
    s1 = np.where(lin, t, np.sin(t * th) / st)
    s1 = np.where(d < 0, -s1, s1)
    return s0[..., None] * a + s1[..., None] * b


# --------------------------------------------------------------------------------------------- truth
def truth_pose(t, start=np.zeros(3)):
    """p(t), q(t) of SURVEY §8d, t measured from the window origin."""
    t = np.asarray(t, dtype=np.float64)
    p = np.stack([0.5 * t, 0.3 * np.sin(0.8 * t), 0.05 * np.sin(1.3 * t)], axis=-1) + start
    w = np.stack([0.05 * np.sin(0.7 * t), 0.04 * np.cos(0.9 * t), 0.2 * t], axis=-1)
    return p, so3_exp(w)


@dataclass
class Config:
    name: str
    rings: int
    az_steps: int          # azimuth steps per revolution
    n_points: int
    room: tuple            # (Lx, Ly, H) full extents
    K: int                 # control poses (SampleStates)
    el_fov_deg: float      # +- elevation
    az_span_deg: float = 360.0
    az0_deg: float = 0.0
    rev_hz: float = 10.0
    start: tuple = (0.0, 0.0, 1.5)
    range_sigma: float = 0.01
    n_panels: int = 8
    fix_points: int = 0    # points of the preceding (already optimised) sweep that feeds the fixed window
    t0: float = 1000.0
    bg_true: tuple = (0.004, -0.003, 0.005)
    ba_true: tuple = (0.06, -0.05, 0.04)


CONFIGS = {
    # C1: 10 k points, K=4.  16 rings over a 60 deg sector facing the +x wall of a 6x5x3 m room, so that a
    # 10 k-point sweep is dense enough for >=20-point clusters (a full 360 deg scan of 10 k points is not).
    "C1": Config("C1", rings=16, az_steps=125, n_points=10_000, room=(6.0, 5.0, 3.0), K=4, el_fov_deg=15.0,
                 az_span_deg=60.0, az0_deg=-30.0, start=(0.8, 0.0, 1.5), fix_points=4_000),
    # C2: VLP-16 shaped, 99 840 points over 2 s, K=8
    "C2": Config("C2", rings=16, az_steps=312, n_points=99_840, room=(12.0, 10.0, 4.0), K=8, el_fov_deg=15.0,
                 fix_points=24_960),
    # C3: OS1-128 shaped, 2 M points over 1.526 s, K=12, 200 Hz IMU
    "C3": Config("C3", rings=128, az_steps=1024, n_points=2_000_000, room=(40.0, 30.0, 8.0), K=12, el_fov_deg=22.5,
                 fix_points=393_216),
}


def _planes(cfg: Config, seed):
    """Room walls + panels as (point c, normal n, axis u, axis v, half extents hu, hv)."""
    Lx, Ly, H = cfg.room
    hx, hy = Lx / 2, Ly / 2
    P = []
    ex, ey, ez = np.eye(3)
    big = 1e9
    P.append((np.array([hx, 0, H / 2]), ex, ey, ez, big, big))
    P.append((np.array([-hx, 0, H / 2]), -ex, ey, ez, big, big))
    P.append((np.array([0, hy, H / 2]), ey, ex, ez, big, big))
    P.append((np.array([0, -hy, H / 2]), -ey, ex, ez, big, big))
    P.append((np.array([0, 0, 0.0]), -ez, ex, ey, big, big))
    P.append((np.array([0, 0, H]), ez, ex, ey, big, big))
    idx = np.arange(cfg.n_panels * 8, dtype=np.uint64)
    u = uniform(seed, 900, idx).reshape(cfg.n_panels, 8)
    for i in range(cfg.n_panels):
        cx = (u[i, 0] - 0.5) * 0.8 * Lx
        cy = (u[i, 1] - 0.5) * 0.8 * Ly
        if abs(cx - cfg.start[0]) < 1.0 and abs(cy - cfg.start[1]) < 1.0:
            cx += 2.0
        yaw = u[i, 2] * np.pi
        w = 1.0 + u[i, 3] * min(Lx, Ly) * 0.15
        h = min(H, 1.0 + u[i, 4] * 0.4 * H)
        n = np.array([np.cos(yaw), np.sin(yaw), 0.0])
        a = np.array([-np.sin(yaw), np.cos(yaw), 0.0])
        P.append((np.array([cx, cy, h / 2]), n, a, ez, w / 2, h / 2))
    return P


def _cast(origins, dirs, planes):
    """nearest positive ray/plane hit distance (n,), inf if none."""
    best = np.full(len(origins), np.inf)
    for (c, n, a, b, ha, hb) in planes:
        denom = dirs @ n
        with np.errstate(divide="ignore", invalid="ignore"):
            t = ((c - origins) @ n) / denom
        ok = np.isfinite(t) & (t > 1e-6)
        if ha < 1e8:
            hit = origins + dirs * np.where(ok, t, 0.0)[:, None] - c
            ok &= (np.abs(hit @ a) <= ha) & (np.abs(hit @ b) <= hb)
        best = np.where(ok & (t < best), t, best)
    return best


def _imu_truth(cfg: Config, ts, start):
    """gyr/acc at IMU times consistent with the reference's discrete prediction scheme."""
    dt = ts[1] - ts[0]
    tt = ts - cfg.t0
    p, q = truth_pose(tt, start)
    h = 1e-4
    _, qm = truth_pose(tt - h, start)
    _, qp = truth_pose(tt + h, start)
    gyr = quat_log(quat_mul(quat_conj(qm), qp)) / (2 * h)
    # p3 = (R1 (a1 - ba) + g) dt^2 + 2 p2 - p1  =>  a1 = R1^T ((p3 - 2 p2 + p1)/dt^2 - g)
    p_ext, _ = truth_pose(np.concatenate([tt, tt[-1:] + dt, tt[-1:] + 2 * dt]), start)
    dd = (p_ext[2:] - 2 * p_ext[1:-1] + p_ext[:-2]) / (dt * dt)
    acc = quat_rotate(quat_conj(q), dd - GRAV)
    return p, q, gyr, acc


def _predict(imu):
    """forward-predict imu[2:] poses in place, lidar_odometry.cc:112-123 with ba = bg = 0."""
    for i in range(2, len(imu)):
        i1, i2, i3 = imu[i - 2], imu[i - 1], imu[i]
        dt = i3["timestamp"] - i2["timestamp"]
        dq = so3_exp(((i2["gyr"] + i3["gyr"]) / 2 * dt)[None])[0]
        imu["rot"][i] = quat_mul(i2["rot"], dq)
        imu["pos"][i] = (quat_rotate(i1["rot"], i1["acc"]) + GRAV) * dt * dt + 2 * i2["pos"] - i1["pos"]


def _pose_at(imu, t):
    """lerp/slerp of IMU states at times t (vectorised UndistortSweep pose lookup, lower_bound semantics)."""
    idx = np.searchsorted(imu["timestamp"], t, side="left")
    assert idx.min() >= 1 and idx.max() < len(imu)
    t0, t1 = imu["timestamp"][idx - 1], imu["timestamp"][idx]
    f = (t - t0) / (t1 - t0)
    pos = imu["pos"][idx - 1] * (1 - f)[:, None] + imu["pos"][idx] * f[:, None]
    rot = quat_slerp(imu["rot"][idx - 1], f, imu["rot"][idx])
    return pos, rot


def _scan(cfg: Config, seed, stream, n_points, t_first, imu_world, planes, start, chunk=1 << 18):
    """n_points firings starting at t_first; world coordinates via the poses in imu_world."""
    out = np.zeros(n_points, dtype=T.POINT48)
    R, A = cfg.rings, cfg.az_steps
    rate = cfg.rev_hz * A * R
    el = np.deg2rad(np.linspace(-cfg.el_fov_deg, cfg.el_fov_deg, R))
    for s in range(0, n_points, chunk):
        k = np.arange(s, min(n_points, s + chunk), dtype=np.int64)
        t = t_first + k / rate
        ring = k % R
        col = k // R
        az = np.deg2rad(cfg.az0_deg + cfg.az_span_deg * (col % A) / A)
        d_b = np.stack([np.cos(el[ring]) * np.cos(az), np.cos(el[ring]) * np.sin(az), np.sin(el[ring])], axis=-1)
        p_t, q_t = truth_pose(t - cfg.t0, start)
        rho = _cast(p_t, quat_rotate(q_t, d_b), planes)
        rho = rho + cfg.range_sigma * normal(seed, stream, k.astype(np.uint64))
        body = d_b * rho[:, None]
        pos, rot = _pose_at(imu_world, t)
        w = quat_rotate(rot, body) + pos
        out["x"][k], out["y"][k], out["z"][k] = w[:, 0], w[:, 1], w[:, 2]
        out["intensity"][k] = 1.0
        out["time"][k] = t
        out["ring"][k] = ring
        out["pad"][k] = np.where(np.isfinite(rho) & (rho >= 0.3) & (rho <= 120.0), 1.0, 0.0)  # validity, stripped below
    keep = out["pad"] > 0
    out = out[keep].copy()
    out["pad"] = 1.0  # PCL_ADD_POINT4D sets data[3] = 1
    return out


@dataclass
class Window:
    cfg: Config
    seed: int
    points: np.ndarray            # POINT48, world frame via the prior, time ordered
    imu: np.ndarray               # IMU states of the window (prior poses, measured acc/gyr)
    samples: np.ndarray           # K SampleStates (prior poses, zero corrections)
    fix_points: np.ndarray        # POINT48 of the preceding sweep (truth poses)
    fix_imu: np.ndarray           # IMU states covering fix_points (truth poses)
    truth_sample_pos: np.ndarray = field(default=None)
    truth_sample_rot: np.ndarray = field(default=None)


def make_window(cfg, seed=20240116) -> Window:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    start = np.array(cfg.start)
    planes = _planes(cfg, seed)
    dt = 1.0 / 200.0
    rate = cfg.rev_hz * cfg.az_steps * cfg.rings
    delta = 1e-3
    T_pts = cfg.n_points / rate
    m = int(np.ceil((delta + T_pts + 0.002) / dt))
    span = (m + 0.5) * dt
    ts_imu = cfg.t0 + dt * np.arange(m + 2)
    p, q, gyr, acc = _imu_truth(cfg, ts_imu, start)
    n_imu = len(ts_imu)
    idx = np.arange(n_imu * 3, dtype=np.uint64)
    sg = 0.00015198973532354657 * np.sqrt(200.0)
    sa = 0.006308226052016165 * np.sqrt(200.0)
    imu = np.zeros(n_imu, dtype=T.IMU)
    imu["timestamp"] = ts_imu
    imu["gyr"] = gyr + np.array(cfg.bg_true) + sg * normal(seed, 11, idx).reshape(n_imu, 3)
    imu["acc"] = acc + np.array(cfg.ba_true) + sa * normal(seed, 12, idx).reshape(n_imu, 3)
    imu["pos"][:2], imu["rot"][:2] = p[:2], q[:2]
    _predict(imu)

    pts = _scan(cfg, seed, 21, cfg.n_points, cfg.t0 + delta, imu, planes, start)

    K = cfg.K
    samples = np.zeros(K, dtype=T.SAMPLE)
    samples["timestamp"] = cfg.t0 + span * np.arange(K) / (K - 1)
    samples["timestamp"][-1] = cfg.t0 + span
    samples["grav"] = GRAV
    # sample poses: lerp/slerp of the IMU states (lidar_odometry.cc:439-449); sample 0 sits on imu[0]
    ts_s = samples["timestamp"].copy()
    pos_s, rot_s = _pose_at(imu, np.maximum(ts_s, ts_imu[0] + 1e-12))
    samples["pos"], samples["rot"] = pos_s, rot_s
    samples["pos"][0], samples["rot"][0] = imu["pos"][0], imu["rot"][0]
    tp, tq = truth_pose(ts_s - cfg.t0, start)

    # preceding sweep for the fixed window: truth poses (already optimised), ends before the window starts
    fix_pts = np.zeros(0, dtype=T.POINT48)
    fix_imu = np.zeros(0, dtype=T.IMU)
    if cfg.fix_points > 0:
        T_fix = cfg.fix_points / rate
        mf = int(np.ceil((T_fix + 0.004) / dt)) + 1
        ts_f = cfg.t0 - dt * np.arange(mf, -1, -1)  # ... up to t0
        pf, qf, gf, af = _imu_truth(cfg, ts_f, start)
        fix_imu = np.zeros(len(ts_f), dtype=T.IMU)
        fix_imu["timestamp"], fix_imu["pos"], fix_imu["rot"], fix_imu["gyr"], fix_imu["acc"] = ts_f, pf, qf, gf, af
        fix_pts = _scan(cfg, seed, 22, cfg.fix_points, cfg.t0 - 0.002 - T_fix, fix_imu, planes, start)
    return Window(cfg, seed, pts, imu, samples, fix_pts, fix_imu, tp, tq)


# --------------------------------------------------------------------------------- C5: correspondence stress
@dataclass
class StressWindow:
    """BASELINE config 5: surfels sampled directly on the scene planes and correspondences generated
    directly (extraction and matching skipped), K control poses over a long window."""
    name: str
    seed: int
    surfels: np.ndarray      # SURFEL, body frame, prior poses, time ordered
    corr: np.ndarray         # CORR sliding-window pairs, timestamp(s1) < timestamp(s2)
    samples: np.ndarray      # K SampleStates (prior poses, zero corrections)
    truth_sample_pos: np.ndarray


def prior_pose(t, start, dp_scale=0.02, dth_scale=0.002):
    """truth composed with a smooth prior error of the SURVEY §8d form: dp = dp_scale [sin .5t, cos .4t, sin .3t] m,
    dtheta = dth_scale [cos .6t, sin .5t, cos .35t] rad.  (With the 5 cm / 10 mrad of the sweep configs the 20 m lever arms
    of this room saturate the Cauchy loss and the solver crawls for ~100 iterations; 2 cm / 2 mrad converges in ~20.)"""
    p, q = truth_pose(t, start)
    dp = dp_scale * np.stack([np.sin(0.5 * t), np.cos(0.4 * t), np.sin(0.3 * t)], axis=-1)
    dth = dth_scale * np.stack([np.cos(0.6 * t), np.sin(0.5 * t), np.cos(0.35 * t)], axis=-1)
    return p + dp, quat_mul(so3_exp(dth), q)


def quat_to_matrix(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                     np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                     np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)


def make_stress_window(n_corr=10_000_000, K=64, seed=20240116, span=8.0, room=(40.0, 30.0, 8.0), t0=1000.0, name="C5",
                       chunk=1 << 20) -> StressWindow:
    """n_corr surfels (one pair each): surfel i lies on room plane i mod 6 (centre uniform on the plane plus N(0, 1 cm)
    along the normal, covariance diag(1e-4, 0.04, 0.05) in the plane's frame), observed at t_i (stratified uniform over
    `span` seconds, so the array is time ordered) from the TRUE pose and stored in the body frame with the PRIOR pose;
    it is paired with a later surfel of the same plane, j = i + 6 m with m uniform, so the interval pairs cover the
    upper triangle uniformly."""
    S = int(n_corr)
    start = np.array([0.0, 0.0, 1.5])
    Lx, Ly, H = room
    ex, ey, ez = np.eye(3)
    # (point on plane, normal, in-plane axes, half extents)
    planes = [(np.array([Lx / 2, 0, H / 2]), ex, ey, ez, Ly / 2, H / 2), (np.array([-Lx / 2, 0, H / 2]), -ex, ey, ez, Ly / 2, H / 2),
              (np.array([0, Ly / 2, H / 2]), ey, ex, ez, Lx / 2, H / 2), (np.array([0, -Ly / 2, H / 2]), -ey, ex, ez, Lx / 2, H / 2),
              (np.array([0, 0, 0.0]), -ez, ex, ey, Lx / 2, Ly / 2), (np.array([0, 0, H]), ez, ex, ey, Lx / 2, Ly / 2)]
    pc = np.stack([p[0] for p in planes]); pn = np.stack([p[1] for p in planes])
    pa = np.stack([p[2] for p in planes]); pb = np.stack([p[3] for p in planes])
    pha = np.array([p[4] for p in planes]); phb = np.array([p[5] for p in planes])
    lam = np.array([1e-4, 0.04, 0.05])
    out = np.zeros(S, dtype=T.SURFEL)
    corr = np.zeros(S, dtype=T.CORR)
    for s in range(0, S, chunk):
        i = np.arange(s, min(S, s + chunk), dtype=np.int64)
        iu = i.astype(np.uint64)
        t_rel = span * (i + uniform(seed, 31, iu)) / S
        pl = (i % 6).astype(np.int64)
        ua, ub = uniform(seed, 32, iu), uniform(seed, 33, iu)
        c_true = pc[pl] + pa[pl] * ((2 * ua - 1) * pha[pl])[:, None] + pb[pl] * ((2 * ub - 1) * phb[pl])[:, None]
        c_true = c_true + pn[pl] * (0.01 * normal(seed, 34, iu))[:, None]
        Rw = np.stack([pn[pl], pa[pl], pb[pl]], axis=-1)                  # columns: normal, in-plane axes
        cov_w = np.einsum("nij,j,nkj->nik", Rw, lam, Rw)
        p_t, q_t = truth_pose(t_rel, start)
        p_p, q_p = prior_pose(t_rel, start)
        Rt = quat_to_matrix(q_t)
        # body-frame observation through the TRUE pose; the prior pose is what the window starts from
        c_b = np.einsum("nji,nj->ni", Rt, c_true - p_t)
        cov_b = np.einsum("nji,njk,nkl->nil", Rt, cov_w, Rt)
        n_b = np.einsum("nji,nj->ni", Rt, pn[pl])
        flip = np.sum(n_b * c_b, axis=1) < 0   # normal away from the view point (the sensor origin in the body frame)
        n_b = np.where(flip[:, None], -n_b, n_b)
        out["timestamp"][i] = t0 + t_rel
        out["resolution"][i] = 0.8
        out["plane_std_deviation"][i] = 0.01
        out["rot"][i], out["pos"][i] = q_p, p_p
        out["center"][i], out["covariance"][i], out["norm"][i] = c_b, cov_b.reshape(-1, 9), n_b
        out["is_in_body_frame"][i] = 1
        # partner: a later surfel of the same plane
        room_m = (S - 1 - i) // 6
        m = 1 + np.floor(uniform(seed, 35, iu) * np.maximum(room_m, 1)).astype(np.int64)
        j = np.minimum(i + 6 * m, S - 1 - ((S - 1 - i) % 6))
        corr["s1"][i], corr["s2"][i] = i, j
    corr = corr[corr["s2"] > corr["s1"]]
    corr = corr[out["timestamp"][corr["s1"]] < out["timestamp"][corr["s2"]]].copy()
    smp = np.zeros(K, dtype=T.SAMPLE)
    eps = 1e-3
    ts = -eps + (span + 2 * eps) * np.arange(K) / (K - 1)
    smp["timestamp"] = t0 + ts
    smp["grav"] = GRAV
    smp["pos"], smp["rot"] = prior_pose(ts, start)
    return StressWindow(name, seed, out, corr, smp, truth_pose(ts, start)[0])
