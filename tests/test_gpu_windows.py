"""Windows resident across sweeps (second half): three consecutive sweeps through the device-resident
sliding / fixed windows against the oracle driving the same sequence on the host — append, pose update of the whole
window, both matchers, the solve, and ShrinkToFit's push_front order (lidar_odometry.cc:527-528,228-250)."""
import numpy as np
import pytest

from wildcat_slam_b200 import synthetic as S
from wildcat_slam_b200 import types as T

pytestmark = pytest.mark.gpu


def test_three_sweeps_resident_windows_match_oracle(oracle):
    from wildcat_slam_b200 import odometry as od

    w = S.make_window("C2")
    n = len(w.points)
    sweeps = [w.points[: n // 3], w.points[n // 3: 2 * n // 3], w.points[2 * n // 3:]]
    ctx = od.Context(0)
    try:
        fix0 = oracle.update_surfel_poses(w.fix_imu, oracle.build_surfels(w.fix_points)["surfels"])[1]
        rw = od.ResidentWindows(ctx, fix_body=fix0)
        o_sld, o_fix = np.zeros(0, T.SURFEL), fix0
        for k, pts in enumerate(sweeps):
            # ---- oracle: AddLidarScan steps 7-13 on host deques
            new = oracle.build_surfels(pts)["surfels"]
            o_sld = np.concatenate([o_sld, new])
            st, o_sld = oracle.update_surfel_poses(w.imu, o_sld)
            assert st == 0
            cs, _ = oracle.match(o_sld, o_sld, True)
            cf, _ = oracle.match(o_sld, o_fix, False)
            st, smp_o, so = oracle.window_solve(o_sld, o_fix, cs, cf, w.imu, w.samples)
            assert st == 0
            # ---- device: only the sweep crosses the boundary after the first call
            x, sg, stats = rw.AddSweep(pts, w.imu, w.samples)
            assert stats.n_surfels == len(new) and rw.n_sld == len(o_sld)
            # timestamp ties inside a sweep may be ordered differently (Q5): pair counts within 0.3 %, same solution
            assert abs(stats.n_sld_corr - len(cs)) <= 0.003 * len(cs) + 2 and abs(stats.n_fix_corr - len(cf)) <= 0.003 * len(cf) + 2
            np.testing.assert_allclose(x, smp_o["data_cor"], rtol=0, atol=2e-4)
            assert sg.final_cost == pytest.approx(so.final_cost, rel=2e-3)
            g_sld, g_fix = rw.Fetch()
            np.testing.assert_allclose(np.sort(g_sld["timestamp"]), np.sort(o_sld["timestamp"]), rtol=0, atol=1e-9)
            for f in ("pos", "rot"):  # whole window re-interpolated from the current IMU states
                np.testing.assert_allclose(g_sld[f][np.argsort(g_sld["timestamp"], kind="stable")],
                                           o_sld[f][np.argsort(o_sld["timestamp"], kind="stable")], rtol=0, atol=1e-9)
            assert (g_sld["is_in_body_frame"] == 1).all()
            # ---- ShrinkToFit after the second sweep: the first ~0.4 s leave the sliding window
            if k == 1:
                t_cut = w.points["time"][0] + 0.4
                m = int(np.searchsorted(o_sld["timestamp"], t_cut, side="left"))
                assert 0 < m < len(o_sld)
                o_fix = np.concatenate([o_sld[:m][::-1], o_fix])  # push_front one by one
                o_sld = o_sld[m:]
                ns, nf = rw.ShrinkToFit(t_cut)
                assert (ns, nf) == (len(o_sld), len(o_fix))
                g_sld, g_fix = rw.Fetch()
                np.testing.assert_allclose(g_fix["timestamp"][:m], o_fix["timestamp"][:m], rtol=0, atol=1e-9)  # newest first
                assert g_fix[m:].tobytes() == fix0.tobytes() and (np.diff(g_fix["timestamp"][:m]) <= 0).all()
                assert (np.diff(g_sld["timestamp"]) >= 0).all() and g_sld["timestamp"][0] >= t_cut
        # reference behaviour (Q6): the fixed window never shrinks; with trim_fixed it keeps `duration` behind its newest surfel
        ns, nf = rw.ShrinkToFit(-1.0, fix_window_duration=0.05, trim_fixed=False)
        assert nf == len(o_fix)
        ns, nf = rw.ShrinkToFit(-1.0, fix_window_duration=0.05, trim_fixed=True)
        t = o_fix["timestamp"]
        keep = len(t)
        while keep > 0 and t[0] - t[keep - 1] > 0.05:
            keep -= 1
        assert nf == keep and 0 < keep < len(o_fix)
    finally:
        ctx.close()
