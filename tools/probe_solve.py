"""solve-only timing probe on C3: with / without IMU factors, exact vs reference Jacobians."""
import sys
sys.path.insert(0, ".")
import numpy as np
from wildcat_slam_b200 import odometry as od, synthetic as S, types as T
w = S.make_window("C3")
ctx = od.Context(0)
sld = od.UpdateSurfelPoses(w.imu, od.BuildSurfels(w.points, ctx=ctx), ctx=ctx)
fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
m = od.KnnSurfelMatcher(ctx); m.BuildIndex(sld); cs, _ = m.Match(sld)
m2 = od.KnnSurfelMatcher(ctx); m2.BuildIndex(fix); cf, _ = m2.Match(sld)
rw = od.ResidentWindow(sld, fix, cs, cf, w.imu, w.samples, ctx)
for imu in (1, 0):
    o = T.default_solve_opts(); o.use_imu_factors = imu
    for rep in range(3):
        x, s = rw.solve(o)
    print(f"use_imu={imu}: iters {s.num_iterations} solve_ms {s.gpu_ms_total:.3f} per-iter {1e3 * s.gpu_ms_total / max(1, s.num_iterations):.1f} us")
rw0 = od.ResidentWindow(sld[:0], fix[:0], cs[:0], cf[:0], w.imu, w.samples, ctx)
x, s = rw0.solve()
print(f"imu only: iters {s.num_iterations} solve_ms {s.gpu_ms_total:.3f} per-iter {1e3 * s.gpu_ms_total / max(1, s.num_iterations):.1f} us")
