"""Parity of the sweep-preparation row (AddLidarScan's extrinsic + range / blind-box filter,
lidar_odometry.cc:489-496, and UndistortSweep, :143-158) against the CPU oracle.  Needs a B200: -m gpu."""
import numpy as np
import pytest

from wildcat_slam_b200 import synthetic as S
from wildcat_slam_b200 import types as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def od():
    from wildcat_slam_b200 import odometry

    return odometry


@pytest.fixture(scope="module")
def ctx(od):
    c = od.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    from oracle import wc_oracle

    wc_oracle.build()
    return wc_oracle


def _same(a, b):
    """field-wise bit equality (the 48-byte record has padding bytes that carry no meaning)."""
    return len(a) == len(b) and all(np.array_equal(np.ascontiguousarray(a[f]).view(np.uint8), np.ascontiguousarray(b[f]).view(np.uint8))
                                  for f in a.dtype.names)


def _raw_cloud(name, extra=True):
    """a lidar-frame cloud with points on both sides of every filter decision."""
    w = S.make_window(name)
    pts = w.points.copy()
    if extra:
        rng = np.random.default_rng(11)
        n = len(pts)
        k = n // 7
        pts["x"][:k] = rng.uniform(-1.0, 1.0, k).astype(np.float32)        # around the blind box and the 0.3 m sphere
        pts["y"][:k] = rng.uniform(-1.0, 1.0, k).astype(np.float32)
        pts["z"][:k] = rng.uniform(-0.6, 0.6, k).astype(np.float32)
        pts["x"][k:2 * k] = rng.uniform(-130.0, 130.0, k).astype(np.float32)  # around the 120 m sphere
        pts["y"][k:2 * k] = rng.uniform(-130.0, 130.0, k).astype(np.float32)
        # exact boundary cases of the box (in imu_link: undo the default extrinsic is not needed, any frame will do)
    return w, pts


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_filter_points_bit_exact(od, ctx, oracle, name):
    w, pts = _raw_cloud(name)
    st, ref = oracle.filter_points(pts)
    assert st == 0
    got = od.FilterPoints(pts, ctx=ctx)
    assert 0 < len(ref) < len(pts)                     # the cloud exercises both outcomes
    assert len(got) == len(ref)
    assert _same(got, ref)                             # order, kept set and the float32 coordinates: bit exact
    # another extrinsic / limits
    f = T.default_sweep_filter()
    q = np.array([0.1, -0.2, 0.3, 0.9]); q /= np.linalg.norm(q)
    f.ext_q[:] = q.tolist()
    f.ext_t[:] = [0.5, -0.25, 0.125]
    f.min_range, f.max_range = 1.0, 60.0
    st, ref = oracle.filter_points(pts, f)
    got = od.FilterPoints(pts, f, ctx=ctx)
    assert st == 0 and _same(got, ref)


def test_filter_points_edge_cases(od, ctx, oracle):
    empty = np.zeros(0, dtype=T.POINT48)
    assert len(od.FilterPoints(empty, ctx=ctx)) == 0
    one = np.zeros(1, dtype=T.POINT48)
    one["x"], one["time"] = 5.0, 1.0
    assert _same(od.FilterPoints(one, ctx=ctx), oracle.filter_points(one)[1])
    # ragged size (not a multiple of the 1024-point CTA), everything dropped, everything kept
    w, pts = _raw_cloud("C1", extra=False)
    pts = pts[:1500]
    far = pts.copy(); far["x"] = 500.0
    assert len(od.FilterPoints(far, ctx=ctx)) == 0 == len(oracle.filter_points(far)[1])
    assert _same(od.FilterPoints(pts, ctx=ctx), oracle.filter_points(pts)[1])
    # time order violation -> WC_EINVAL_TIME_ORDER (CHECK lidar_odometry.cc:491)
    bad = pts.copy(); bad["time"][700] = bad["time"][0] - 1.0
    assert oracle.filter_points(bad)[0] == T.WC_EINVAL_TIME_ORDER
    from wildcat_slam_b200.abi import WildcatError
    with pytest.raises(WildcatError) as e:
        od.FilterPoints(bad, ctx=ctx)
    assert e.value.status == T.WC_EINVAL_TIME_ORDER


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_undistort_matches_oracle(od, ctx, oracle, name):
    w = S.make_window(name)
    pts = w.points  # any IMU-frame coordinates will do: every timestamp lies inside the IMU span
    st, ref = oracle.undistort_sweep(w.imu, pts)
    assert st == 0
    got = od.UndistortSweep(pts, w.imu, ctx=ctx)
    for fld in ("intensity", "time", "ring"):
        assert np.array_equal(got[fld], ref[fld])
    # fp64 slerp (acos / sin from different math libraries) then a float32 store: identical up to one float32 ulp
    for fld in ("x", "y", "z"):
        ulp = np.spacing(np.abs(ref[fld]).astype(np.float32))
        assert (np.abs(got[fld].astype(np.float64) - ref[fld].astype(np.float64)) <= ulp).all()
        assert np.mean(got[fld] == ref[fld]) > 0.999


def test_undistort_error_and_fused_resident_path(od, ctx, oracle):
    w = S.make_window("C1")
    from wildcat_slam_b200.abi import WildcatError
    bad = w.points.copy()
    bad["time"][-1] = w.imu["timestamp"][-1] + 1.0  # beyond the IMU span: CHECK lidar_odometry.cc:150
    assert oracle.undistort_sweep(w.imu, bad)[0] == T.WC_EOUT_OF_SPAN
    with pytest.raises(WildcatError) as e:
        od.UndistortSweep(bad, w.imu, ctx=ctx)
    assert e.value.status == T.WC_EOUT_OF_SPAN
    # fused: raw sweep -> (undistort + repack on the device) -> BuildSurfels == BuildSurfels(UndistortSweep(raw))
    und = od.UndistortSweep(w.points, w.imu, ctx=ctx)
    ref = od.BuildSurfels(und, ctx=ctx)
    rs = od.ResidentSweep(w.points, ctx=ctx, imu_states=w.imu)
    n, _ = rs.extract()
    got = rs.fetch()
    assert n == len(ref) > 0 and got.tobytes() == ref.tobytes()


def test_predict_states_matches_oracle(od, ctx, oracle):
    """SURVEY section 8(f) rank 2, first part: IMU forward prediction (PredictPoseOfNewImuState, lidar_odometry.cc:112-123)
    and the new sample states (:430-453).  fp64 with a different sincos library and FMA contraction on the device:
    stated tolerances, 400-step recurrences included."""
    from wildcat_slam_b200.abi import WildcatError

    w = S.make_window("C2")
    imu = w.imu.copy()
    imu["pos"][2:] = 0
    imu["rot"][2:] = 0
    ba, bg = np.array([0.01, -0.02, 0.03]), np.array([0.001, 0.002, -0.003])
    t_last, sdt = float(w.samples["timestamp"][0]), 0.08
    n_new = int((imu["timestamp"][-1] - t_last) / sdt)
    st, imu_o, smp_o = oracle.predict_states(imu, ba, bg, S.GRAV, t_last, sdt, n_new)
    assert st == 0 and n_new >= 20
    imu_g, smp_g = od.PredictStates(imu, ba, bg, S.GRAV, t_last, sdt, n_new, ctx=ctx)
    for f in ("timestamp", "acc", "gyr"):
        assert np.array_equal(imu_g[f], imu_o[f])
    np.testing.assert_allclose(imu_g["rot"], imu_o["rot"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(imu_g["pos"], imu_o["pos"], rtol=0, atol=1e-10)
    assert np.array_equal(smp_g["timestamp"], smp_o["timestamp"]) and np.array_equal(smp_g["data_cor"], smp_o["data_cor"])
    assert np.array_equal(smp_g["grav"], smp_o["grav"])
    np.testing.assert_allclose(smp_g["rot"], smp_o["rot"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(smp_g["pos"], smp_o["pos"], rtol=0, atol=1e-10)
    # with zero biases the prediction reproduces the generator's own (independent, numpy) forward prediction
    imu0, _ = od.PredictStates(imu, np.zeros(3), np.zeros(3), S.GRAV, t_last, sdt, 0, ctx=ctx)
    np.testing.assert_allclose(imu0["rot"], w.imu["rot"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(imu0["pos"], w.imu["pos"], rtol=0, atol=1e-10)
    # CHECK_NEAR on the IMU spacing (:119) and CHECK_NE on the sample bracket (:441-442)
    bad = imu.copy()
    bad["timestamp"][10] += 1e-3
    assert oracle.predict_states(bad, ba, bg, S.GRAV, t_last, sdt, 0)[0] == T.WC_EINVAL_TIME_ORDER
    with pytest.raises(WildcatError) as e:
        od.PredictStates(bad, ba, bg, S.GRAV, t_last, sdt, 0, ctx=ctx)
    assert e.value.status == T.WC_EINVAL_TIME_ORDER
    assert oracle.predict_states(imu, ba, bg, S.GRAV, float(imu["timestamp"][-1]), sdt, 1)[0] == T.WC_EOUT_OF_SPAN
    with pytest.raises(WildcatError) as e:
        od.PredictStates(imu, ba, bg, S.GRAV, float(imu["timestamp"][-1]), sdt, 1, ctx=ctx)
    assert e.value.status == T.WC_EOUT_OF_SPAN


def test_prefetched_sweep_equals_plain_upload(od, ctx):
    """wc_points_prefetch: a sweep copied in ahead of time on the copy stream gives bitwise the same surfels as the plain
    upload; an unclaimed prefetch (another buffer is uploaded instead) is ignored; a second prefetch replaces the first."""
    w = S.make_window("C2")
    plain = od.ResidentSweep(w.points, ctx=ctx)
    plain.extract()
    ref = plain.fetch()
    a = ctx.pinned(len(w.points), T.POINT48)
    a[:] = w.points
    b = ctx.pinned(len(w.points) // 2, T.POINT48)
    b[:] = w.points[: len(b)]
    ctx.prefetch(b)           # replaced below
    ctx.prefetch(a)
    rs = od.ResidentSweep(a, ctx=ctx)   # claims the prefetched copy
    rs.extract()
    got = rs.fetch()
    assert got.tobytes() == ref.tobytes()
    ctx.prefetch(b)           # never claimed: the next upload is of another buffer, which drops the prefetch
    rs2 = od.ResidentSweep(w.points, ctx=ctx)
    rs2.extract()
    assert rs2.fetch().tobytes() == ref.tobytes()
    b[:] = w.points[len(b): 2 * len(b)]  # the buffer changes after its prefetch was dropped: the upload must see the new content
    half = od.ResidentSweep(b, ctx=ctx)
    n_half, _ = half.extract()
    plain_half = od.ResidentSweep(np.array(b), ctx=ctx)
    assert plain_half.extract()[0] == n_half and half.fetch().tobytes() == plain_half.fetch().tobytes()


def test_prefetch_at_solve_is_issued_by_the_pass(od, ctx):
    """WC_PREFETCH_AT_SOLVE: the copy of the next sweep starts when the window pass reaches its solve stage; the next
    pass claims it and gives the same result; without a pass in between the upload copies the sweep itself."""
    w = S.make_window("C2")
    fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
    a = ctx.pinned(len(w.points), T.POINT48)
    a[:] = w.points
    b = ctx.pinned(len(w.points), T.POINT48)
    b[:] = w.points
    rp = od.ResidentPass(a, w.imu, w.samples, fix, ctx=ctx)
    ctx.prefetch(b, at_solve=True)
    x1, s1, st1 = rp.run()                      # issues the copy of b beside its solve
    rp2 = od.ResidentPass(b, w.imu, w.samples, None, ctx=ctx, keep_fix=True)   # finds b on the device
    x2, s2, st2 = rp2.run()
    assert st2.n_surfels == st1.n_surfels and st2.n_sld_corr == st1.n_sld_corr and st2.n_fix_corr == st1.n_fix_corr
    assert s2.num_iterations == s1.num_iterations
    np.testing.assert_allclose(x2, x1, rtol=0, atol=1e-9)
    ctx.prefetch(a, at_solve=True)              # no pass follows: the upload below copies the sweep itself
    rs = od.ResidentSweep(a, ctx=ctx)
    n, _ = rs.extract()
    assert n == st1.n_surfels
