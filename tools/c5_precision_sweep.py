"""BASELINE config 5: fp64 vs fp32 tolerance sweep of the fused lidar kernel.  For each size, solves the same stress
window with WC_PREC_F64 / WC_PREC_MIXED / WC_PREC_F32 and prints iterations, final cost, the largest difference of the
solution from the fp64 one, and the time of one linearisation pass."""
import os
import sys

os.environ.setdefault("WC_TIME_PASSES", "1")
sys.path.insert(0, ".")
import numpy as np

from wildcat_slam_b200 import odometry as od, synthetic as S, types as T

sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [20_000, 200_000, 2_000_000]
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
print(f"{'corr':>9} {'mode':>6} {'iters':>5} {'final cost':>16} {'rel dcost':>10} {'|dx|max':>10} {'lin us':>9} {'solve ms':>9} {'GB/s':>7}")
for n in sizes:
    w = S.make_stress_window(n, K=K)
    prm = T.default_params()
    prm.max_surfels, prm.max_corrs, prm.max_samples = max(int(prm.max_surfels), len(w.surfels)), max(int(prm.max_corrs), len(w.corr)), max(128, K)
    ctx = od.Context(0, params=prm)
    rw = od.ResidentWindow(w.surfels, None, w.corr, None, None, w.samples, ctx)
    ref = None
    for name, mode in (("f64", T.WC_PREC_F64), ("mixed", T.WC_PREC_MIXED), ("f32", T.WC_PREC_F32)):
        o = T.default_solve_opts()
        o.use_imu_factors, o.precision = 0, mode
        for _ in range(2):
            x, sm = rw.solve(o)
        if ref is None:
            ref = (x.copy(), sm.final_cost)
        lin = sm.gpu_ms_linearize / max(1, sm.num_linearizations)
        bytes_ = (128 if mode == T.WC_PREC_F64 else 64) * len(w.corr)
        print(f"{len(w.corr):9d} {name:>6} {sm.num_iterations:5d} {sm.final_cost:16.6f} {abs(sm.final_cost / ref[1] - 1):10.2e} "
              f"{np.abs(x - ref[0]).max():10.2e} {lin * 1e3:9.1f} {sm.gpu_ms_total:9.3f} {bytes_ / (lin * 1e-3) / 1e9:7.1f}", flush=True)
    ctx.close()
