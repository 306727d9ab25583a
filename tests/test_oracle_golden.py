"""The oracle against every known answer the reference's own tests hold for this path,
the reference notebook's output, and finite-difference / self-consistency checks of the restated factors.
CPU only."""
import os

import numpy as np
import pytest

from wildcat_slam_b200 import synthetic as S
from wildcat_slam_b200 import types as T

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# --- spline_interpolation_test.cc:10-41 -----------------------------------------------------------------
def test_bspline_approx_known_answers(oracle):
    f = oracle.lib().wco_cubic_bspline_approx
    for s, want in [(0, 2), (1, 3), (0.4, 2.4), (0.5, 2.5)]:
        assert f(1, 2, 3, 4, s) == pytest.approx(want, rel=4 * np.finfo(float).eps)  # EXPECT_DOUBLE_EQ = 4 ulp
    for s in (0, 1, 0.5, 0.4):
        assert f(2, 2, 2, 2, s) == pytest.approx(2, rel=4 * np.finfo(float).eps)


def test_spline_interpolate_known_answers(oracle):
    f = oracle.lib().wco_cubic_spline_interpolate
    for s, want in [(0, 2), (1, 3), (0.4, 2.4), (0.5, 2.5)]:
        assert f(-1, 1, 0, 2, 1, 3, 2, 4, s) == pytest.approx(want, rel=4 * np.finfo(float).eps)
    for s in (0, 1, 0.5, 0.4):
        assert f(-1, 2, 0, 2, 1, 2, 2, 2, s) == pytest.approx(2, rel=4 * np.finfo(float).eps)
    assert f(-1, 2, 0, 3, 1, 1, 2, 2, 0) == pytest.approx(3, rel=4 * np.finfo(float).eps)
    assert f(-1, 2, 0, 3, 1, 1, 2, 2, 1) == pytest.approx(1, rel=4 * np.finfo(float).eps)


# --- spline_interpolation_test.cc:79-96 -----------------------------------------------------------------
P8 = np.array([[1, 1, 1], [2, 3, 2], [4, 5, 5], [6, 6, 3], [5, 4, 1], [6, 7, 1], [9, 9, 8], [12, 15, 11]], dtype=float)
TS8 = np.array([0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0])


def test_interpolator_passes_through_samples(oracle):
    out, valid, _ = oracle.spline_fit_eval(TS8, P8, TS8)
    assert valid.all()
    for i in range(8):  # Eigen isApprox(1e-6): |a-b| <= 1e-6 * min(|a|,|b|)
        assert np.linalg.norm(out[i] - P8[i]) <= 1e-6 * min(np.linalg.norm(out[i]), np.linalg.norm(P8[i]))
    out, valid, _ = oracle.spline_fit_eval(TS8, P8, np.array([0.29999, 1.00001]))
    assert not valid.any()  # Interp returns nullptr outside [t0, tK-1]


# --- scripts/CubicBSpline3D.ipynb -----------------------------------------------------------------------
def test_interpolator_matches_reference_notebook(oracle):
    g = np.load(os.path.join(GOLD, "bspline_notebook.npz"))
    Nbs, Nq = int(g["Nbs"]), 8
    i = np.arange(1, Nbs + 1)
    u = Nq * (i / Nbs)
    u = u[u >= 1]
    assert len(u) == len(g["BSpline"]) == 438
    # the notebook evaluates at index_f = u; the C++ class maps t -> (t-t0)/(tK-t0)*(K-1)+1.  Only u <= K is
    # inside the class's domain (u in [1, 8]).
    t = TS8[0] + (u - 1.0) / (Nq - 1) * (TS8[-1] - TS8[0])
    out, valid, ctrl = oracle.spline_fit_eval(TS8, g["p"], np.minimum(t, TS8[-1]))
    assert valid.all()
    np.testing.assert_allclose(ctrl, g["Q"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out, g["BSpline"], rtol=0, atol=2e-6)  # t(u) round trip costs ~1e-7 in u


# --- utils_test.cc:5-21 ---------------------------------------------------------------------------------
def test_jl_jl_inv(oracle):
    v = np.array([1.0, 2.0, 3.0])
    jl = oracle.so3("jl", v).reshape(3, 3)
    jli = oracle.so3("jl_inv", v).reshape(3, 3)
    ref = np.linalg.inv(jl)
    assert np.linalg.norm(jli - ref) <= 1e-12 * min(np.linalg.norm(jli), np.linalg.norm(ref))  # isApprox default


def test_jl_jr(oracle):
    v = np.array([1.0, 2.0, 3.0])
    np.testing.assert_allclose(oracle.so3("jl", v), oracle.so3("jr", -v), rtol=1e-12)


def test_exp_log_roundtrip_and_scipy(oracle):
    from scipy.spatial.transform import Rotation as R

    rng = np.random.default_rng(1)
    for _ in range(50):
        v = rng.normal(size=3) * rng.choice([1e-12, 1e-3, 0.5, 2.0])
        if np.linalg.norm(v) > 3.0:  # stay inside the principal branch of Log
            v *= 3.0 / np.linalg.norm(v)
        q = oracle.so3("exp", v)
        np.testing.assert_allclose(q, R.from_rotvec(v).as_quat(), atol=1e-15)
        np.testing.assert_allclose(oracle.so3("log", q), v, atol=1e-14)


def test_sym_eig_matches_lapack(oracle):
    rng = np.random.default_rng(2)
    for _ in range(100):
        a = rng.normal(size=(3, 3))
        a = a @ a.T * rng.choice([1e-6, 1.0, 1e4])
        out = oracle.so3("eig", a.reshape(-1))
        w, v = out[:3], out[3:].reshape(3, 3)
        wl = np.linalg.eigvalsh(a)
        np.testing.assert_allclose(w, wl, rtol=1e-12, atol=1e-15 * np.abs(wl).max())
        np.testing.assert_allclose(a @ v, v * w, atol=1e-12 * np.abs(wl).max())
        np.testing.assert_allclose(v.T @ v, np.eye(3), atol=1e-13)


# --- knn_surfel_matcher_test.cc:19-43 -------------------------------------------------------------------
def test_knn_self_is_nearest(oracle):
    rng = np.random.default_rng(3)
    vecs = rng.uniform(-1, 1, size=(10_000, 6))  # Vector6::Random()
    idx, d2 = oracle.knn6(vecs, vecs, 10, use_kdtree=True)
    assert idx.shape == (10_000, 10)
    assert (idx[:, 0] == np.arange(10_000)).all()
    assert (d2[:, 0] == 0).all() and (np.diff(d2, axis=1) >= 0).all()
    # kd-tree == brute force (exact search)
    sub = slice(0, 500)
    idx_b, d2_b = oracle.knn6(vecs[sub], vecs, 10, use_kdtree=False)
    assert (idx_b == idx[sub]).all() and (d2_b == d2[sub]).all()


# --- restated factors: finite differences (exact mode) and the Q1 overwrite ---------------------------
def _fd(fun, x, eps=1e-6):
    x = x.reshape(-1)
    cols = []
    for i in range(len(x)):
        xp, xm = x.copy(), x.copy()
        xp[i] += eps
        xm[i] -= eps
        cols.append((np.atleast_1d(fun(xp)) - np.atleast_1d(fun(xm))) / (2 * eps))
    return np.stack(cols, axis=-1)


def _c1_surfels(oracle, w):
    r = oracle.build_surfels(w.points)
    st, sld = oracle.update_surfel_poses(w.imu, r["surfels"])
    assert st == 0
    cs, _ = oracle.match(sld, sld, True)
    return sld, cs


def test_lidar_factor_jacobian_fd_and_q1(oracle, c1_window):
    w = c1_window
    sld, cs = _c1_surfels(oracle, w)
    ts = w.samples["timestamp"]
    K = len(ts)
    rng = np.random.default_rng(4)
    x = rng.normal(size=K * 12) * 0.02
    seen_modes = set()
    for c in cs[::7]:
        s1, s2 = sld[c["s1"]], sld[c["s2"]]
        st, r0, jac, wt, n = oracle.lidar_factor(s1, s2, False, ts, x, jacobian_mode=T.WC_JAC_EXACT)
        assert st == 1
        fd = _fd(lambda xx: oracle.lidar_factor(s1, s2, False, ts, xx)[1], x)[0]
        np.testing.assert_allclose(jac, fd, atol=1e-6 * max(1.0, np.abs(fd).max()))
        _, _, jq, _, _ = oracle.lidar_factor(s1, s2, False, ts, x, jacobian_mode=T.WC_JAC_REFERENCE_OVERWRITE)
        i1 = np.searchsorted(ts, s1["timestamp"], side="right")
        i2 = np.searchsorted(ts, s2["timestamp"], side="right")
        mode = 0 if i1 < i2 - 1 else (1 if i1 == i2 - 1 else 2)
        seen_modes.add(mode)
        if mode == 0:
            np.testing.assert_array_equal(jq, jac)  # no aliasing -> identical
        else:
            assert np.abs(jq - jac).max() > 0  # Q1: the s1 term is lost in the aliased block(s)
            if mode == 2:
                # both blocks hold only the s2 terms: J_s2 * (1-f2), J_s2 * f2 (cost_functor.h:152-175,225-228)
                f2 = (s2["timestamp"] - ts[i2 - 1]) / (ts[i2] - ts[i2 - 1])
                a = jq[12 * (i2 - 1):12 * (i2 - 1) + 6] / (1 - f2)
                b = jq[12 * i2:12 * i2 + 6] / f2
                np.testing.assert_allclose(a, b, rtol=1e-9)
    assert {1, 2} <= seen_modes or {0, 2} <= seen_modes
    # unary factor (fixed-window surfel as s1)
    s1, s2 = sld[cs[0]["s1"]], sld[cs[0]["s2"]]
    st, r0, jac, _, _ = oracle.lidar_factor(s1, s2, True, ts, x)
    fd = _fd(lambda xx: oracle.lidar_factor(s1, s2, True, ts, xx)[1], x)[0]
    np.testing.assert_allclose(jac, fd, atol=1e-6 * max(1.0, np.abs(fd).max()))


def test_imu_factor_fd_except_reference_quirks(oracle, c1_window):
    """All Jacobian blocks match finite differences except the two the reference itself gets wrong:
    d(gyro)/d(rot) uses F with +r (cost_functor.h:303,313,446-448) and d(gyro)/d(bg) is written for both
    i1 and i2 (:304,314).  The restatement reproduces those as written."""
    w = c1_window
    ts = w.samples["timestamp"]
    K = len(ts)
    rng = np.random.default_rng(5)
    for i0 in (0, 30, 60, len(w.imu) - 4):
        i3 = w.imu[i0:i0 + 3]
        k = np.searchsorted(ts, i3["timestamp"][0], side="right")
        mode = 1 if k == K - 1 else 0
        tss = ts[k - 1:k + 2] if mode == 0 else ts[k - 1:k + 1]
        nb = 3 if mode == 0 else 2
        x = rng.normal(size=nb * 12) * 0.02
        st, res, jac = oracle.imu_factor(i3, tss, mode, S.GRAV, x)
        assert st == 12
        fd = _fd(lambda xx: oracle.imu_factor(i3, tss, mode, S.GRAV, xx)[1], x)
        mask = np.ones_like(jac, dtype=bool)
        for b in range(nb):
            mask[0:3, 12 * b + 0:12 * b + 3] = False  # gyro / rot
            mask[0:3, 12 * b + 6:12 * b + 9] = False  # gyro / bg
        np.testing.assert_allclose(jac[mask], fd[mask], atol=2e-5 * np.abs(fd).max())
        assert np.abs(jac[~mask] - fd[~mask]).max() > 1e-3  # the quirk is really there


def test_window_evaluate_consistent_with_factors(oracle, c1_window):
    """gradient of the robustified cost == J^T r; J^T J symmetric PSD; exact-mode gradient == FD of the cost."""
    w = c1_window
    sld, cs = _c1_surfels(oracle, w)
    rng = np.random.default_rng(6)
    smp = w.samples.copy()
    smp["data_cor"] = rng.normal(size=(len(smp), 12)) * 1e-3
    o = T.default_solve_opts()
    o.jacobian_mode = T.WC_JAC_EXACT
    o.use_imu_factors = 0
    st, cost, g, H = oracle.window_evaluate(sld, None, cs, None, None, smp, opts=o)
    assert st == 0
    np.testing.assert_allclose(H, H.T, atol=1e-9 * np.abs(H).max())
    assert np.linalg.eigvalsh(H).min() > -1e-8 * np.abs(H).max()

    def f(xx):
        s2 = smp.copy()
        s2["data_cor"] = xx.reshape(-1, 12)
        return oracle.window_evaluate(sld, None, cs, None, None, s2, opts=o, want_jtj=False)[1]

    fd = _fd(f, smp["data_cor"].copy(), eps=1e-7)[0]
    np.testing.assert_allclose(g, fd, atol=2e-5 * np.abs(fd).max())


def test_c1_golden_regression(oracle, c1_window):
    """The committed C1 fixture (tests/golden/make_golden.py) still reproduces: pins the restatement itself."""
    g = np.load(os.path.join(GOLD, "c1_oracle.npz"))
    w = c1_window
    np.testing.assert_array_equal(np.stack([w.points["x"], w.points["y"], w.points["z"]], 1), g["points_xyz"])
    np.testing.assert_array_equal(w.points["time"], g["points_t"])
    r = oracle.build_surfels(w.points, want_assign=True)
    assert r["surfels"].tobytes() == g["surfels"].tobytes()
    assert r["assign"].tobytes() == g["assign"].tobytes()
    st, sld = oracle.update_surfel_poses(w.imu, r["surfels"])
    cs, _ = oracle.match(sld, sld, True)
    assert cs.tobytes() == g["sld_corr"].tobytes()
    st, smp, summ = oracle.window_solve(sld, g["fix_body"], cs, g["fix_corr"], w.imu, w.samples)
    n = summ.num_iterations
    np.testing.assert_allclose(np.array(summ.iter_cost[1:n + 1]), g["iter_cost"][1:], rtol=1e-12)
    np.testing.assert_allclose(smp["data_cor"], g["data_cor"], rtol=1e-9, atol=1e-12)


def test_matcher_bruteforce_equals_kdtree(oracle, c2_window):
    w = c2_window
    r = oracle.build_surfels(w.points)
    st, sld = oracle.update_surfel_poses(w.imu, r["surfels"])
    a, _ = oracle.match(sld, sld, True, use_kdtree=True)
    b, _ = oracle.match(sld, sld, True, use_kdtree=False)
    assert a.tobytes() == b.tobytes() and len(a) > 1000
    # every pair is time ordered and respects the gates' time threshold
    t = sld["timestamp"]
    assert (t[a["s1"]] < t[a["s2"]]).all() and (t[a["s2"]] - t[a["s1"]] >= 0.06).all()
    # de-dup: no unordered pair twice
    pairs = set(map(tuple, np.stack([a["s1"], a["s2"]], 1).tolist()))
    assert len(pairs) == len(a)


def test_solve_reduces_cost_and_recovers_bias(oracle, c2_window):
    w = c2_window
    r = oracle.build_surfels(w.points)
    st, sld = oracle.update_surfel_poses(w.imu, r["surfels"])
    rf = oracle.build_surfels(w.fix_points)
    st, fix = oracle.update_surfel_poses(w.fix_imu, rf["surfels"])
    cs, _ = oracle.match(sld, sld, True)
    cf, fit = oracle.match(sld, fix, False)
    assert fit.all()  # fixed surfels are older: s1 always indexes the target array
    st, smp, summ = oracle.window_solve(sld, fix, cs, cf, w.imu, w.samples)
    assert st == 0 and summ.final_cost < 0.2 * summ.initial_cost
    err0 = np.linalg.norm(w.samples["pos"] - w.truth_sample_pos, axis=1)
    err1 = np.linalg.norm(w.samples["pos"] + smp["data_cor"][:, 3:6] - w.truth_sample_pos, axis=1)
    assert err1[-1] < 0.5 * err0[-1]
    np.testing.assert_allclose(smp["data_cor"][-1, 6:9], w.cfg.bg_true, atol=1.5e-3)


def test_sweep_row_golden_regression(oracle):
    """the oracle's sweep-preparation row (lidar_odometry.cc:489-496, 143-158) against the committed fixture
    (tests/golden/make_golden.py: sweep) — a regression pin of the restatement; upstream has no test for it."""
    import os

    from wildcat_slam_b200 import synthetic as S
    from wildcat_slam_b200 import types as T

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_sweep_oracle.npz"))
    w = S.make_window("C1")
    n = len(g["raw_t"])
    pts = np.zeros(n, dtype=T.POINT48)
    pts["x"], pts["y"], pts["z"], pts["time"] = g["raw_xyz"][:, 0], g["raw_xyz"][:, 1], g["raw_xyz"][:, 2], g["raw_t"]
    st, kept = oracle.filter_points(pts)
    assert st == 0 and len(kept) == len(g["kept_t"]) and 0 < len(kept) < n
    assert np.array_equal(np.stack([kept["x"], kept["y"], kept["z"]], 1), g["kept_xyz"]) and np.array_equal(kept["time"], g["kept_t"])
    # the default extrinsic maps (x, y, z) -> (-y - 0.001, -x - 0.00855, -z + 0.055) up to the 5e-8 off-diagonal terms
    raw = g["raw_xyz"].astype(np.float64)
    approx = np.stack([-raw[:, 1] - 0.001, -raw[:, 0] - 0.00855, -raw[:, 2] + 0.055], 1)
    keep = (np.linalg.norm(approx, axis=1) >= 0.3) & (np.linalg.norm(approx, axis=1) <= 120.0) & ~(
        (approx[:, 0] >= -0.8) & (approx[:, 0] <= 0.3) & (np.abs(approx[:, 1]) <= 0.5) & (np.abs(approx[:, 2]) <= 0.4))
    assert abs(int(keep.sum()) - len(kept)) <= 2  # independent numpy restatement (boundary ties aside)
    st, und = oracle.undistort_sweep(w.imu, w.points[:n])
    assert st == 0 and np.array_equal(np.stack([und["x"], und["y"], und["z"]], 1), g["undistorted_xyz"])
    # time-order violation against the last KEPT point aborts (CHECK lidar_odometry.cc:491)
    bad = pts.copy()
    bad["time"][n // 2] = bad["time"][0] - 1.0
    assert oracle.filter_points(bad)[0] == T.WC_EINVAL_TIME_ORDER


def test_predict_states_against_numpy_restatement(oracle):
    """the oracle's IMU forward prediction + sample-state creation (lidar_odometry.cc:112-123, 403-453) against an
    independent numpy restatement (wildcat_slam_b200.synthetic: _predict / _pose_at, written for the generator)."""
    from wildcat_slam_b200 import synthetic as S
    from wildcat_slam_b200 import types as T

    w = S.make_window("C1")
    imu = w.imu.copy()
    imu["pos"][2:] = 0
    imu["rot"][2:] = 0
    t_last, sdt = float(w.samples["timestamp"][0]), 0.08
    n_new = int((imu["timestamp"][-1] - t_last) / sdt)
    st, out, smp = oracle.predict_states(imu, np.zeros(3), np.zeros(3), S.GRAV, t_last, sdt, n_new)
    assert st == 0 and n_new >= 3
    np.testing.assert_allclose(out["rot"], w.imu["rot"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["pos"], w.imu["pos"], rtol=0, atol=1e-12)
    pos, rot = S._pose_at(w.imu, smp["timestamp"])
    np.testing.assert_allclose(smp["pos"], pos, rtol=0, atol=1e-12)
    assert np.minimum(np.abs(smp["rot"] - rot).max(1), np.abs(smp["rot"] + rot).max(1)).max() < 1e-12
    assert np.array_equal(smp["timestamp"], t_last + sdt * np.arange(1, n_new + 1)) and not smp["data_cor"].any()
    ba, bg = np.array([0.01, -0.02, 0.03]), np.array([0.001, 0.002, -0.003])
    st, out2, smp2 = oracle.predict_states(imu, ba, bg, S.GRAV, t_last, sdt, 1)
    chk = imu.copy()
    for i in range(2, len(chk)):
        i1, i2, i3 = chk[i - 2], chk[i - 1], chk[i]
        dt = i3["timestamp"] - i2["timestamp"]
        chk["rot"][i] = S.quat_mul(i2["rot"], S.so3_exp((((i2["gyr"] + i3["gyr"]) / 2 - bg) * dt)[None])[0])
        chk["pos"][i] = (S.quat_rotate(i1["rot"], i1["acc"] - ba) + S.GRAV) * dt * dt + 2 * i2["pos"] - i1["pos"]
    np.testing.assert_allclose(out2["rot"], chk["rot"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out2["pos"], chk["pos"], rtol=0, atol=1e-12)
    assert np.array_equal(smp2["data_cor"][0, 6:9], bg) and np.array_equal(smp2["data_cor"][0, 9:12], ba)
    bad = imu.copy()
    bad["timestamp"][5] += 1e-3
    assert oracle.predict_states(bad, ba, bg, S.GRAV, t_last, sdt, 0)[0] == T.WC_EINVAL_TIME_ORDER
    assert oracle.predict_states(imu, ba, bg, S.GRAV, float(imu["timestamp"][-1]), sdt, 1)[0] == T.WC_EOUT_OF_SPAN
