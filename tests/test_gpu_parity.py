"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs, against the committed
golden fixtures, and through size-independent properties at the full BASELINE sizes.  Needs a B200: -m gpu."""
import os

import numpy as np
import pytest

from wildcat_slam_b200 import synthetic as S
from wildcat_slam_b200 import types as T

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

# ---- tolerances (fp64 path; stated per quantity) ---------------------------------------------------------------
TOL_CENTER = 6e-9        # m: surfel centre (world).  The GPU sums exact integers of a 2^-27 m (7.45e-9 m) fixed-point grid: a float32
#                          coordinate at least 2^-4 m away from its coordinate plane lies ON the grid (no rounding at all); a closer one is
#                          rounded by at most 0.75 units = 5.6e-9 m, which bounds the error of a mean of such points (typical: 1e-9)
TOL_COV = 2e-9           # m^2: covariance entries (oracle restates the reference's single-pass E[xx^T]-mu mu^T, Q3)
TOL_TIME = 1e-9          # s: mean timestamp (GPU sums exact 2^-36 s fixed point; oracle sums fp64 sequentially)
TOL_NORMAL = 2e-6        # eigenvector of a covariance known to TOL_COV with eigen-gaps >= 1e-3 m^2
TOL_COST_REL = 1e-9      # relative, per-iteration LM cost
TOL_X = 1e-8             # data_cor entries (pose corrections, rad / m)


@pytest.fixture(scope="module")
def od():
    from wildcat_slam_b200 import odometry

    return odometry


@pytest.fixture(scope="module")
def ctx(od):
    c = od.Context(0)
    yield c
    c.close()


def _canon(s):
    """canonical order for comparison: the reference's std::sort orders by timestamp only and leaves (near-)ties
    unspecified (Q5) — surfels of a parent and a child node built from the same points share their mean timestamp up
    to summation round-off — so both sides are re-ordered by (timestamp rounded to 1e-8 s, resolution desc, centre
    rounded to 1 um)."""
    key = np.lexsort((np.round(s["center"][:, 2], 6), np.round(s["center"][:, 1], 6), np.round(s["center"][:, 0], 6),
                      -s["resolution"], np.round(s["timestamp"], 8)))
    return s[key]


def _assert_surfels_close(g, o):
    assert len(g) == len(o)
    assert (np.diff(g["timestamp"]) >= 0).all()  # sorted by timestamp, surfel_extraction.cc:334
    # positions may differ from the oracle's only inside groups of (near-)equal timestamps
    np.testing.assert_allclose(g["timestamp"], o["timestamp"], rtol=0, atol=TOL_TIME)
    g, o = _canon(g), _canon(o)
    np.testing.assert_allclose(g["timestamp"], o["timestamp"], rtol=0, atol=TOL_TIME)
    np.testing.assert_array_equal(g["resolution"], o["resolution"])
    np.testing.assert_allclose(g["center"], o["center"], rtol=0, atol=TOL_CENTER)
    np.testing.assert_allclose(g["covariance"], o["covariance"], rtol=0, atol=TOL_COV)
    np.testing.assert_allclose(g["plane_std_deviation"] ** 2, o["plane_std_deviation"] ** 2, rtol=0, atol=TOL_COV)
    dots = np.sum(g["norm"] * o["norm"], axis=1)
    assert (dots > 1 - TOL_NORMAL).all(), dots.min()
    np.testing.assert_array_equal(g["rot"], o["rot"])
    np.testing.assert_array_equal(g["is_in_body_frame"], o["is_in_body_frame"])


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_extract_matches_oracle(od, ctx, oracle, name):
    w = S.make_window(name)
    ref = oracle.build_surfels(w.points, want_assign=True, rel_margin=1e-6)
    assert ref["near_threshold"] == 0, "seeded input sits on a planarity threshold; pick another seed"
    g, assign = od.BuildSurfels(w.points, ctx=ctx, want_assign=True)
    # voxel / octree-cell index assignment: bit exact
    assert assign.tobytes() == ref["assign"].tobytes()
    _assert_surfels_close(g, ref["surfels"])


def test_extract_matches_committed_golden(od, ctx):
    gold = np.load(os.path.join(GOLD, "c1_oracle.npz"))
    w = S.make_window("C1")
    g, assign = od.BuildSurfels(w.points, ctx=ctx, want_assign=True)
    assert assign.tobytes() == gold["assign"].tobytes()
    _assert_surfels_close(g, gold["surfels"])


def test_extract_edge_cases(od, ctx, oracle):
    # empty sweep
    assert len(od.BuildSurfels(np.zeros(0, T.POINT48), ctx=ctx)) == 0
    # fewer points than any threshold
    w = S.make_window("C1")
    assert len(od.BuildSurfels(w.points[:15], ctx=ctx)) == 0
    # ragged sizes around the warp / block granularity
    for n in (31, 33, 255, 257, 4097):
        ref = oracle.build_surfels(w.points[:n], want_assign=True)
        g, a = od.BuildSurfels(w.points[:n], ctx=ctx, want_assign=True)
        assert a.tobytes() == ref["assign"].tobytes()
        _assert_surfels_close(g, ref["surfels"])
    # time order violated -> WC_EINVAL_TIME_ORDER (CHECK lidar_odometry.cc:491)
    bad = w.points[:1000].copy()
    bad["time"][500] = bad["time"][0] - 1.0
    from wildcat_slam_b200.abi import WildcatError

    with pytest.raises(WildcatError) as e:
        od.BuildSurfels(bad, ctx=ctx)
    assert e.value.status == T.WC_EINVAL_TIME_ORDER
    # negative coordinates / voxel boundaries: points exactly on voxel and octree-cell faces
    pts = np.zeros(4 * 64, dtype=T.POINT48)
    vs = float(np.float32(0.8))
    k = np.arange(len(pts))
    pts["x"] = np.float32((k % 8 - 4) * vs / 4)
    pts["y"] = np.float32(((k // 8) % 8 - 4) * vs / 8)
    pts["z"] = np.float32(-0.05 + 1e-3 * (k % 5))
    pts["time"] = 10.0 + 1e-4 * k
    ref = oracle.build_surfels(pts, want_assign=True)
    g, a = od.BuildSurfels(pts, ctx=ctx, want_assign=True)
    assert a.tobytes() == ref["assign"].tobytes()


def test_extract_is_bitwise_reproducible(od, ctx):
    """exact int64 accumulation => identical bytes run to run, whatever the atomic order."""
    w = S.make_window("C2")
    a = od.BuildSurfels(w.points, ctx=ctx)
    b = od.BuildSurfels(w.points, ctx=ctx)
    assert a.tobytes() == b.tobytes()


def test_update_surfel_poses_matches_oracle(od, ctx, oracle):
    w = S.make_window("C1")
    ref = oracle.build_surfels(w.points)["surfels"]
    st, o = oracle.update_surfel_poses(w.imu, ref)
    g = od.UpdateSurfelPoses(w.imu, ref, ctx=ctx)
    for f in ("pos", "rot", "center", "norm", "covariance"):
        np.testing.assert_allclose(g[f], o[f], rtol=0, atol=1e-12)
    assert (g["is_in_body_frame"] == 1).all()
    from wildcat_slam_b200.abi import WildcatError

    with pytest.raises(WildcatError) as e:  # surfel outside the IMU span -> CHECK :164
        od.UpdateSurfelPoses(w.imu[50:], ref, ctx=ctx)
    assert e.value.status == T.WC_EOUT_OF_SPAN


def _body_surfels(oracle, w):
    sld = oracle.update_surfel_poses(w.imu, oracle.build_surfels(w.points)["surfels"])[1]
    fix = oracle.update_surfel_poses(w.fix_imu, oracle.build_surfels(w.fix_points)["surfels"])[1]
    return sld, fix


def test_knn_known_answer(od, ctx, oracle):
    """knn_surfel_matcher_test.cc:19-43 on the CUDA kNN: every vector's nearest neighbour is itself, k = 10."""
    rng = np.random.default_rng(3)
    vecs = rng.uniform(-1, 1, size=(10_000, 6))
    m = od.KnnSurfelMatcher(ctx)
    idx, d2 = m.KNearestSearchVectors(vecs, vecs, 10)
    assert idx.shape == (10_000, 10) and (idx[:, 0] == np.arange(10_000)).all()
    oi, od2 = oracle.knn6(vecs, vecs, 10, use_kdtree=False)
    np.testing.assert_array_equal(idx, oi)
    np.testing.assert_array_equal(d2, od2)  # same accumulation order, no FMA: bit exact


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_match_identical_to_oracle(od, ctx, oracle, name):
    w = S.make_window(name)
    sld, fix = _body_surfels(oracle, w)
    m = od.KnnSurfelMatcher(ctx)
    m.BuildIndex(sld)
    g, _ = m.Match(sld)
    o, _ = oracle.match(sld, sld, True, use_kdtree=False)
    assert g.tobytes() == o.tobytes() and len(g) > 50
    m2 = od.KnnSurfelMatcher(ctx)
    m2.BuildIndex(fix)
    g2, fit = m2.Match(sld)
    o2, ofit = oracle.match(sld, fix, False, use_kdtree=False)
    assert g2.tobytes() == o2.tobytes() and (fit == ofit).all()
    # empty target set -> no correspondences (knn_surfel_matcher.cc:18-20)
    m3 = od.KnnSurfelMatcher(ctx)
    m3.BuildIndex(np.zeros(0, T.SURFEL))
    assert len(m3.Match(sld)[0]) == 0


def _window(oracle, name):
    w = S.make_window(name)
    sld, fix = _body_surfels(oracle, w)
    cs, _ = oracle.match(sld, sld, True)
    cf, _ = oracle.match(sld, fix, False)
    return w, sld, fix, cs, cf


@pytest.mark.parametrize("jac_mode", [T.WC_JAC_REFERENCE_OVERWRITE, T.WC_JAC_EXACT])
def test_evaluate_matches_oracle(od, ctx, oracle, jac_mode):
    w, sld, fix, cs, cf = _window(oracle, "C1")
    rng = np.random.default_rng(7)
    smp = w.samples.copy()
    smp["data_cor"] = rng.normal(size=(len(smp), 12)) * 1e-3
    o = T.default_solve_opts()
    o.jacobian_mode = jac_mode
    st, c_o, g_o, H_o = oracle.window_evaluate(sld, fix, cs, cf, w.imu, smp, opts=o)
    assert st == 0
    c_g, g_g, H_g = od.EvaluateWindow(sld, fix, cs, cf, w.imu, smp, opts=o, ctx=ctx)
    assert c_g == pytest.approx(c_o, rel=1e-12)
    np.testing.assert_allclose(g_g, g_o, rtol=0, atol=1e-10 * np.abs(g_o).max())
    np.testing.assert_allclose(H_g, H_o, rtol=0, atol=1e-10 * np.abs(H_o).max())


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_solve_matches_oracle_iteration_by_iteration(od, ctx, oracle, name):
    w, sld, fix, cs, cf = _window(oracle, name)
    st, smp_o, so = oracle.window_solve(sld, fix, cs, cf, w.imu, w.samples)
    smp_g, sg = od.SolveWindow(sld, fix, cs, cf, w.imu, w.samples, ctx=ctx)
    assert sg.num_iterations == so.num_iterations and sg.termination == so.termination
    n = so.num_iterations
    assert list(sg.iter_accepted[1:n + 1]) == list(so.iter_accepted[1:n + 1])
    np.testing.assert_allclose(np.array(sg.iter_cost[1:n + 1]), np.array(so.iter_cost[1:n + 1]), rtol=TOL_COST_REL)
    assert sg.initial_cost == pytest.approx(so.initial_cost, rel=1e-12)
    assert sg.final_cost == pytest.approx(so.final_cost, rel=TOL_COST_REL)
    np.testing.assert_allclose(smp_g["data_cor"], smp_o["data_cor"], rtol=0, atol=TOL_X)
    assert sg.num_residual_blocks_imu == so.num_residual_blocks_imu


def test_solve_matches_committed_golden(od, ctx):
    gold = np.load(os.path.join(GOLD, "c1_oracle.npz"))
    w = S.make_window("C1")
    smp, sg = od.SolveWindow(gold["sld_body"], gold["fix_body"], gold["sld_corr"], gold["fix_corr"], w.imu, w.samples, ctx=ctx)
    n = sg.num_iterations
    np.testing.assert_allclose(np.array(sg.iter_cost[1:n + 1]), gold["iter_cost"][1:], rtol=TOL_COST_REL)
    np.testing.assert_allclose(smp["data_cor"], gold["data_cor"], rtol=0, atol=TOL_X)


def test_solve_error_codes(od, ctx, oracle):
    from wildcat_slam_b200.abi import WildcatError

    w, sld, fix, cs, cf = _window(oracle, "C1")
    bad = cs.copy()
    bad["s1"][0], bad["s2"][0] = cs["s2"][0], cs["s1"][0]  # CHECK_LT(s1.t, s2.t), lidar_odometry.cc:256
    with pytest.raises(WildcatError) as e:
        od.SolveWindow(sld, fix, bad, cf, w.imu, w.samples, ctx=ctx)
    assert e.value.status == T.WC_EINVAL_TIME_ORDER
    short = w.samples[:-1].copy()  # surfels beyond the last sample: CHECK(sp2r_it != end), :266
    with pytest.raises(WildcatError) as e:
        od.SolveWindow(sld, fix, cs, cf, None, short, ctx=ctx)
    assert e.value.status == T.WC_EOUT_OF_SPAN
    # no correspondences, no IMU: zero cost, gradient tolerance at iteration 0
    smp, sg = od.SolveWindow(sld, None, None, None, None, w.samples, ctx=ctx)
    assert sg.num_iterations == 0 and sg.initial_cost == 0 and (smp["data_cor"] == 0).all()


def test_spline_known_answers_and_notebook(od, ctx, oracle):
    P8 = np.array([[1, 1, 1], [2, 3, 2], [4, 5, 5], [6, 6, 3], [5, 4, 1], [6, 7, 1], [9, 9, 8], [12, 15, 11]], dtype=float)
    TS8 = np.array([0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0])
    it = od.CubicBSplineInterpolator(TS8, P8, ctx=ctx)
    for i in range(8):  # spline_interpolation_test.cc:79-96
        p = it.Interp(TS8[i])
        assert p is not None and np.linalg.norm(p - P8[i]) <= 1e-6 * min(np.linalg.norm(p), np.linalg.norm(P8[i]))
    assert it.Interp(0.2999) is None and it.Interp(1.0001) is None
    g = np.load(os.path.join(GOLD, "bspline_notebook.npz"))
    u = 8 * (np.arange(1, 501) / 500)
    u = u[u >= 1]
    t = np.minimum(TS8[0] + (u - 1.0) / 7 * (TS8[-1] - TS8[0]), TS8[-1])
    out, valid = it.InterpMany(t)
    assert valid.all()
    np.testing.assert_allclose(out, g["BSpline"], rtol=0, atol=2e-6)
    o_out, o_valid, _ = oracle.spline_fit_eval(TS8, P8, t)
    np.testing.assert_allclose(out, o_out, rtol=0, atol=1e-10)


def test_apply_corrections_matches_oracle(od, ctx, oracle):
    w, sld, fix, cs, cf = _window(oracle, "C1")
    st, smp_o, _ = oracle.window_solve(sld, fix, cs, cf, w.imu, w.samples)
    st, s_o, i_o = oracle.apply_corrections(smp_o, w.imu)
    assert st == 0
    s_g, i_g = od.ApplyCorrections(smp_o, w.imu, ctx=ctx)
    for f in ("rot", "pos", "data_cor"):
        np.testing.assert_allclose(s_g[f], s_o[f], rtol=0, atol=1e-10)
    for f in ("rot", "pos"):
        np.testing.assert_allclose(i_g[f], i_o[f], rtol=0, atol=1e-10)
    assert (s_g["data_cor"][:, :6] == 0).all()


def test_full_window_pipeline_c2(od, ctx, oracle):
    """GPU end to end (extract -> poses -> match x2 -> solve) against the oracle end to end.  Surfels with (near-)equal
    mean timestamps may be ordered differently (std::sort ties, Q5), which permutes indices and can re-route a handful
    of order-dependent de-duplications, so the comparison is on the solution, not on index lists."""
    w = S.make_window("C2")
    sld = od.UpdateSurfelPoses(w.imu, od.BuildSurfels(w.points, ctx=ctx), ctx=ctx)
    fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
    m = od.KnnSurfelMatcher(ctx)
    m.BuildIndex(sld)
    cs, _ = m.Match(sld)
    m2 = od.KnnSurfelMatcher(ctx)
    m2.BuildIndex(fix)
    cf, _ = m2.Match(sld)
    smp, sg = od.SolveWindow(sld, fix, cs, cf, w.imu, w.samples, ctx=ctx)
    o_sld, o_fix = _body_surfels(oracle, w)
    o_cs, _ = oracle.match(o_sld, o_sld, True)
    o_cf, _ = oracle.match(o_sld, o_fix, False)
    assert abs(len(cs) - len(o_cs)) <= 0.002 * len(o_cs) and abs(len(cf) - len(o_cf)) <= 0.002 * len(o_cf)
    # the same correspondences as (timestamp, timestamp) pairs, up to the few re-routed ones
    key = lambda s, c: set(zip(np.round(s["timestamp"][c["s1"]], 7).tolist(), np.round(s["timestamp"][c["s2"]], 7).tolist()))  # noqa: E731
    a, b = key(sld, cs), key(o_sld, o_cs)
    assert len(a ^ b) <= 0.01 * len(b)
    st, smp_o, so = oracle.window_solve(o_sld, o_fix, o_cs, o_cf, w.imu, w.samples)
    np.testing.assert_allclose(smp["data_cor"], smp_o["data_cor"], rtol=0, atol=2e-4)
    assert sg.final_cost == pytest.approx(so.final_cost, rel=2e-3)
    # and the device-resident fused pass gives the host-API result exactly
    rp = od.ResidentPass(w.points, w.imu, w.samples, fix, ctx=ctx)
    x, s2, stats = rp.run()
    assert stats.n_surfels == len(sld) and stats.n_sld_corr == len(cs) and stats.n_fix_corr == len(cf)
    assert s2.num_iterations == sg.num_iterations
    np.testing.assert_allclose(x, smp["data_cor"], rtol=0, atol=1e-10)


def test_extract_recovers_after_errors(od, ctx, oracle):
    """A call that fails with a documented, recoverable error (time order, capacity) must leave the context clean: the
    next BuildSurfels on the same context equals the oracle (the emit kernels have already counted surfels into the
    time-bucket histogram by the time the error is known)."""
    from wildcat_slam_b200.abi import WildcatError

    w = S.make_window("C2")
    ref = oracle.build_surfels(w.points, want_assign=True)
    bad = w.points.copy()
    bad["time"][len(bad) // 2] = bad["time"][0] - 1.0  # full sweep, one timestamp out of order
    with pytest.raises(WildcatError) as e:
        od.BuildSurfels(bad, ctx=ctx)
    assert e.value.status == T.WC_EINVAL_TIME_ORDER
    g, a = od.BuildSurfels(w.points, ctx=ctx, want_assign=True)
    assert a.tobytes() == ref["assign"].tobytes()
    _assert_surfels_close(g, ref["surfels"])
    # capacity: a context that cannot hold the sweep's surfels
    prm = T.default_params()
    prm.max_surfels = 64
    small = od.Context(0, params=prm)
    try:
        with pytest.raises(WildcatError) as e:
            od.BuildSurfels(w.points, ctx=small)
        assert e.value.status == T.WC_ECAPACITY
        few = w.points[:6000]
        r2 = oracle.build_surfels(few)
        if len(r2["surfels"]) <= 64:
            _assert_surfels_close(od.BuildSurfels(few, ctx=small), r2["surfels"])
    finally:
        small.close()
    g2 = od.BuildSurfels(w.points, ctx=ctx)
    assert g2.tobytes() == g.tobytes()


def test_c3_matches_oracle(od, ctx, oracle):
    """BASELINE headline config (2 M points, K = 12) against the oracle itself: per-point assignment bit exact, surfel set
    within the stated tolerances, both correspondence lists byte-identical on the same body-frame surfels, the LM solve
    iteration by iteration (cost, accept sequence, termination) and the final data_cor."""
    w = S.make_window("C3")
    ref = oracle.build_surfels(w.points, want_assign=True, rel_margin=1e-6)
    assert ref["near_threshold"] == 0, "seeded input sits on a planarity threshold; pick another seed"
    g, assign = od.BuildSurfels(w.points, ctx=ctx, want_assign=True)
    assert assign.tobytes() == ref["assign"].tobytes()
    _assert_surfels_close(g, ref["surfels"])
    # matcher + solve on the ORACLE's surfels (so index lists are comparable entry by entry)
    sld, fix = _body_surfels(oracle, w)
    gs = od.UpdateSurfelPoses(w.imu, ref["surfels"], ctx=ctx)
    for f in ("pos", "rot", "center", "norm", "covariance"):
        np.testing.assert_allclose(gs[f], sld[f], rtol=0, atol=1e-12)
    m = od.KnnSurfelMatcher(ctx)
    m.BuildIndex(sld)
    cs, _ = m.Match(sld)
    o_cs, _ = oracle.match(sld, sld, True)
    assert cs.tobytes() == o_cs.tobytes() and len(cs) > 10_000
    m2 = od.KnnSurfelMatcher(ctx)
    m2.BuildIndex(fix)
    cf, fit = m2.Match(sld)
    o_cf, o_fit = oracle.match(sld, fix, False)
    assert cf.tobytes() == o_cf.tobytes() and (fit == o_fit).all() and len(cf) > 10_000
    st, smp_o, so = oracle.window_solve(sld, fix, o_cs, o_cf, w.imu, w.samples)
    assert st == 0
    smp_g, sg = od.SolveWindow(sld, fix, cs, cf, w.imu, w.samples, ctx=ctx)
    assert sg.num_iterations == so.num_iterations and sg.termination == so.termination
    n = so.num_iterations
    assert list(sg.iter_accepted[1:n + 1]) == list(so.iter_accepted[1:n + 1])
    np.testing.assert_allclose(np.array(sg.iter_cost[1:n + 1]), np.array(so.iter_cost[1:n + 1]), rtol=TOL_COST_REL)
    assert sg.final_cost == pytest.approx(so.final_cost, rel=TOL_COST_REL)
    np.testing.assert_allclose(smp_g["data_cor"], smp_o["data_cor"], rtol=0, atol=TOL_X)
    # the device-resident fused pass on the GPU's own surfels reaches the same solution (tie order may permute indices)
    fix_g = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
    x, s2, stats = od.ResidentPass(w.points, w.imu, w.samples, fix_g, ctx=ctx).run()
    assert stats.n_surfels == len(g)
    np.testing.assert_allclose(x, smp_o["data_cor"], rtol=0, atol=2e-4)
    assert s2.final_cost == pytest.approx(so.final_cost, rel=2e-3)


def test_c3_full_size_properties(od, ctx):
    """BASELINE full size (2 M points, K = 12): size-independent properties instead of the (slow) oracle."""
    w = S.make_window("C3")
    s1, a1 = od.BuildSurfels(w.points, ctx=ctx, want_assign=True)
    # (1) keys: numpy restatement of VoxelLoc on all 2 M points — bit exact
    vs = np.float64(np.float32(0.8))
    for ax, f in zip("xyz", ("vx", "vy", "vz")):
        np.testing.assert_array_equal(a1[f], np.floor(w.points[ax].astype(np.float64) / vs).astype(np.int32))
    assert a1["leaf"].min() >= 0 and a1["leaf"].max() <= 63
    # (2) sortedness by timestamp, (3) idempotence / bitwise reproducibility
    assert (np.diff(s1["timestamp"]) >= 0).all() and len(s1) > 10_000
    s2 = od.BuildSurfels(w.points, ctx=ctx)
    assert s1.tobytes() == s2.tobytes()
    # (4) every surfel obeys the acceptance tests it was emitted under
    lam = np.linalg.eigvalsh(s1["covariance"].reshape(-1, 3, 3))
    assert (lam[:, 0] <= 0.01 + 1e-9).all() and (2 * (lam[:, 1] - lam[:, 0]) / lam.sum(1) >= 0.1 - 1e-9).all()
    np.testing.assert_allclose(np.sqrt(np.maximum(lam[:, 0], 0)), s1["plane_std_deviation"], atol=1e-7)
    # (5) linearity: extracting the time-shifted sweep shifts the timestamps and nothing else
    sh = w.points.copy()
    sh["time"] += 64.0
    s3 = od.BuildSurfels(sh, ctx=ctx)
    assert len(s3) == len(s1)
    np.testing.assert_allclose(s3["timestamp"] - 64.0, s1["timestamp"], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(s3["center"], s1["center"])
    # (6) solve: cost decreases monotonically over accepted steps and the pose error shrinks 10x
    sld = od.UpdateSurfelPoses(w.imu, s1, ctx=ctx)
    fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
    m = od.KnnSurfelMatcher(ctx)
    m.BuildIndex(sld)
    cs, _ = m.Match(sld)
    m2 = od.KnnSurfelMatcher(ctx)
    m2.BuildIndex(fix)
    cf, _ = m2.Match(sld)
    t = sld["timestamp"]
    assert (t[cs["s1"]] < t[cs["s2"]]).all() and len(cs) > 10_000 and len(cf) > 10_000
    smp, sg = od.SolveWindow(sld, fix, cs, cf, w.imu, w.samples, ctx=ctx)
    acc = [sg.iter_cost[i] for i in range(1, sg.num_iterations + 1) if sg.iter_accepted[i]]
    assert all(b < a for a, b in zip([sg.initial_cost] + acc, acc))
    err0 = np.linalg.norm(w.samples["pos"][-1] - w.truth_sample_pos[-1])
    err1 = np.linalg.norm(w.samples["pos"][-1] + smp["data_cor"][-1, 3:6] - w.truth_sample_pos[-1])
    assert err1 < 0.1 * err0


def test_c3_matcher_grid_is_exact(od, ctx, monkeypatch):
    """Full-size matcher (46 k surfels): the uniform-grid search with every cell size (and its ring / box-scan phases) must
    return exactly the correspondences of the exhaustive scan — the grid is an index, never an approximation."""
    w = S.make_window("C3")
    sld = od.UpdateSurfelPoses(w.imu, od.BuildSurfels(w.points, ctx=ctx), ctx=ctx)
    fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)

    def both(c):
        m = od.KnnSurfelMatcher(c); m.BuildIndex(sld); a, _ = m.Match(sld)
        m2 = od.KnnSurfelMatcher(c); m2.BuildIndex(fix); b, _ = m2.Match(sld)
        return a.tobytes(), b.tobytes()

    monkeypatch.setenv("WC_KNN_GRID_MIN", str(1 << 40))  # read at context creation: always the exhaustive scan
    brute_ctx = od.Context(0)
    monkeypatch.delenv("WC_KNN_GRID_MIN")
    try:
        ref = both(brute_ctx)
    finally:
        brute_ctx.close()
    assert both(ctx) == ref                                # default: unit cells
    for cells in ("1", "2", "4"):
        monkeypatch.setenv("WC_KNN_CELLS_PER_UNIT", cells)  # read at every match call
        assert both(ctx) == ref, cells
    monkeypatch.delenv("WC_KNN_CELLS_PER_UNIT")


def test_solve_wide_system_64_control_poses(od, ctx, oracle):
    """BASELINE config 5 shape in small: 64 control poses (12 * 64 - 3 = 765 unknowns).  The LM system no longer fits
    shared memory, so the Cholesky runs out of global memory (lm_step<false>) with the CTA-wide backward sweep."""
    import dataclasses

    w = S.make_window(dataclasses.replace(S.CONFIGS["C2"], name="C2K64", K=64))
    g = od.BuildSurfels(w.points, ctx=ctx)
    sld = od.UpdateSurfelPoses(w.imu, g, ctx=ctx)
    fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
    m = od.KnnSurfelMatcher(ctx); m.BuildIndex(sld); cs, _ = m.Match(sld)
    m2 = od.KnnSurfelMatcher(ctx); m2.BuildIndex(fix); cf, _ = m2.Match(sld)
    smp, sg = od.SolveWindow(sld, fix, cs, cf, w.imu, w.samples, ctx=ctx)
    st, smp_o, so = oracle.window_solve(sld, fix, cs, cf, w.imu, w.samples)
    assert st == 0 and len(w.samples) == 64
    assert sg.num_iterations == so.num_iterations and sg.termination == so.termination
    assert abs(sg.final_cost / so.final_cost - 1) < TOL_COST_REL
    np.testing.assert_allclose(smp["data_cor"], smp_o["data_cor"], rtol=0, atol=TOL_X)
