"""Observability outputs: Histogram::ToString restated (CPU), residual dump and surfel markers on the
device against the oracle (-m gpu)."""
import numpy as np
import pytest

from wildcat_slam_b200 import report as R
from wildcat_slam_b200 import synthetic as S
from wildcat_slam_b200 import types as T


def test_histogram_to_string_known_answers():
    """common/histogram.cc:29-77 by hand: empty, constant, and a 4-value / 2-bucket case"""
    assert R.Histogram().ToString(10) == "Count: 0"
    assert R.Histogram([2.5, 2.5]).ToString(3) == "Count: 2  Min: 2.5  Max: 2.5  Mean: 2.5"
    s = R.Histogram([0.0, 1.0, 2.0, 4.0]).ToString(2).split("\n")
    assert s[0] == "Count: 4  Min: 0  Max: 4  Mean: 1.75"
    # buckets [0, 2) with 2 values (bar = (2 * 20 + 2) / 4 = 10 chars) and [2, 4] with 2 values
    assert s[1] == "[0.000000, 2.000000)\t" + " " * 10 + "#" * 10 + "\tCount: 2 (50%)\tTotal: 2 (50%)"
    assert s[2] == "[2.000000, 4.000000]\t" + " " * 10 + "#" * 10 + "\tCount: 2 (50%)\tTotal: 4 (100%)"
    h = R.Histogram()
    for v in np.linspace(-1, 1, 101):
        h.Add(v)
    lines = h.ToString(10).split("\n")
    assert len(lines) == 11 and lines[-1].endswith("Total: 101 (100%)")


@pytest.mark.gpu
def test_window_residuals_and_markers_match_oracle(oracle):
    from wildcat_slam_b200 import odometry as od

    w = S.make_window("C1")
    sld = oracle.update_surfel_poses(w.imu, oracle.build_surfels(w.points)["surfels"])[1]
    fix = oracle.update_surfel_poses(w.fix_imu, oracle.build_surfels(w.fix_points)["surfels"])[1]
    cs, _ = oracle.match(sld, sld, True)
    cf, _ = oracle.match(sld, fix, False)
    ctx = od.Context(0)
    try:
        smp, sg = od.SolveWindow(sld, fix, cs, cf, w.imu, w.samples, ctx=ctx)
        for x in (None, smp["data_cor"]):
            at = w.samples.copy()
            if x is not None:
                at["data_cor"] = x
            st, o_sld, o_fix, o_imu = oracle.window_residuals(sld, fix, cs, cf, w.imu, at)
            assert st == 0
            g = R.WindowResiduals(ctx, len(w.samples), data_cor=x)
            assert len(g["sld"]) == len(cs) and len(g["fix"]) == len(cf) and g["imu"].shape == o_imu.shape
            np.testing.assert_allclose(np.sort(g["sld"]), np.sort(o_sld), rtol=0, atol=1e-10)
            np.testing.assert_allclose(np.sort(g["fix"]), np.sort(o_fix), rtol=0, atol=1e-10)
            np.testing.assert_allclose(g["imu"], o_imu, rtol=0, atol=1e-9 * max(1.0, np.abs(o_imu).max()))
        # the cost the solver reports is the cost of these residuals
        tot = 0.5 * (np.sum(g["sld"] ** 2) + np.sum(g["fix"] ** 2))
        assert tot < sg.final_cost  # lidar part of a cost that also holds the IMU blocks
        txt = R.residual_report(g)
        assert "Sliding window Surfel residuals" in txt and "Imu residuals with type acc_bias" in txt
        # markers: world and body-frame surfels
        for s in (oracle.build_surfels(w.points)["surfels"], sld):
            gm, om = R.SurfelMarkers(s, ctx=ctx), oracle.surfel_markers(s)
            np.testing.assert_allclose(gm["position"], om["position"], rtol=0, atol=1e-12)
            np.testing.assert_allclose(gm["scale"], om["scale"], rtol=0, atol=1e-9)
            np.testing.assert_allclose(gm["color"], om["color"], rtol=0, atol=1e-6)
            # q and -q are the same rotation
            d = np.abs(np.sum(gm["orientation"] * om["orientation"], axis=1))
            assert (d > 1 - 1e-9).all()
            assert (np.abs(np.linalg.norm(gm["orientation"], axis=1) - 1) < 1e-12).all()
    finally:
        ctx.close()
