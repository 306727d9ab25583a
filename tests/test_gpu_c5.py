"""BASELINE config 5 (correspondence stress: surfels and correspondences generated directly, 64 control poses) at a
reduced size the oracle finishes in seconds: the CUDA solve against the oracle iteration by iteration, and the
reduced-precision modes of the fused lidar kernel against the fp64 one (the fp64-vs-fp32 tolerance sweep; the full-size
table is profiles/round2_c5_precision_sweep.txt)."""
import numpy as np
import pytest

from wildcat_slam_b200 import synthetic as S
from wildcat_slam_b200 import types as T

pytestmark = pytest.mark.gpu

TOL_COST_REL = 1e-9   # per-iteration LM cost, fp64 path against the oracle
TOL_X = 1e-8          # data_cor (rad / m), fp64 path against the oracle
# reduced precision against the fp64 GPU solve (float32 records: positions of ~20 m carry 2e-6 m, weights of ~60 / m)
TOL_COST_REL_F32 = 2e-5
TOL_X_F32 = 2e-4


@pytest.fixture(scope="module")
def c5():
    from wildcat_slam_b200 import odometry as od

    w = S.make_stress_window(50_000, K=64)
    prm = T.default_params()
    ctx = od.Context(0, params=prm)
    yield od, ctx, w
    ctx.close()


def _opts(mode=T.WC_PREC_F64):
    o = T.default_solve_opts()
    o.use_imu_factors, o.precision = 0, mode
    return o


def test_c5_reduced_matches_oracle(c5, oracle):
    od, ctx, w = c5
    assert len(w.samples) == 64 and len(w.corr) > 49_000
    t = w.surfels["timestamp"]
    assert (np.diff(t) > 0).all() and (t[w.corr["s1"]] < t[w.corr["s2"]]).all()
    st, smp_o, so = oracle.window_solve(w.surfels, None, w.corr, None, None, w.samples, opts=_opts())
    assert st == 0
    smp_g, sg = od.SolveWindow(w.surfels, None, w.corr, None, None, w.samples, opts=_opts(), ctx=ctx)
    assert sg.num_iterations == so.num_iterations and sg.termination == so.termination
    n = so.num_iterations
    assert list(sg.iter_accepted[1:n + 1]) == list(so.iter_accepted[1:n + 1])
    np.testing.assert_allclose(np.array(sg.iter_cost[1:n + 1]), np.array(so.iter_cost[1:n + 1]), rtol=TOL_COST_REL)
    np.testing.assert_allclose(smp_g["data_cor"], smp_o["data_cor"], rtol=0, atol=TOL_X)
    # the bias unknowns are touched by no factor: they stay where they started
    assert (smp_g["data_cor"][:, 6:] == w.samples["data_cor"][:, 6:]).all()


@pytest.mark.parametrize("mode", [T.WC_PREC_MIXED, T.WC_PREC_F32])
def test_c5_reduced_precision_within_tolerance(c5, mode):
    od, ctx, w = c5
    rw = od.ResidentWindow(w.surfels, None, w.corr, None, None, w.samples, ctx)
    x64, s64 = rw.solve(_opts())
    x, s = rw.solve(_opts(mode))
    assert s.termination == s64.termination and abs(s.num_iterations - s64.num_iterations) <= 2
    assert abs(s.final_cost / s64.final_cost - 1) < TOL_COST_REL_F32
    np.testing.assert_allclose(x, x64, rtol=0, atol=TOL_X_F32)
    # fp64 again after a reduced-precision solve: the fp32 records do not leak into the default path
    x64b, _ = rw.solve(_opts())
    np.testing.assert_allclose(x64b, x64, rtol=0, atol=1e-10)


def test_unknown_precision_is_rejected(c5):
    from wildcat_slam_b200.abi import WildcatError

    od, ctx, w = c5
    o = _opts()
    o.precision = 7
    with pytest.raises(WildcatError) as e:
        od.SolveWindow(w.surfels[:1000], None, w.corr[:10], None, None, w.samples, opts=o, ctx=ctx)
    assert e.value.status == T.WC_EINVAL


def test_tma_staged_and_direct_linearize_agree(c5, monkeypatch):
    """The fused lidar kernel reads its record tiles either straight from global memory (small windows) or from shared
    memory stages filled by TMA bulk copies (large windows); both must give the same normal equations."""
    od, ctx, w = c5
    rng = np.random.default_rng(5)
    smp = w.samples.copy()
    smp["data_cor"][:, :6] = rng.normal(size=(len(smp), 6)) * 1e-3
    out = {}
    for staged in ("0", "1"):
        monkeypatch.setenv("WC_LIN_STAGED", staged)
        c2 = od.Context(0)
        try:
            out[staged] = od.EvaluateWindow(w.surfels, None, w.corr, None, None, smp, opts=_opts(), ctx=c2)
        finally:
            c2.close()
    monkeypatch.delenv("WC_LIN_STAGED")
    (c0, g0, H0), (c1, g1, H1) = out["0"], out["1"]
    assert c1 == pytest.approx(c0, rel=1e-13)
    np.testing.assert_allclose(g1, g0, rtol=0, atol=1e-11 * np.abs(g0).max())
    np.testing.assert_allclose(H1, H0, rtol=0, atol=1e-11 * np.abs(H0).max())
