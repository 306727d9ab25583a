"""C3 (2 M points, K = 12) timing probe: per-stage device times, run 3x."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from wildcat_slam_b200 import odometry as od, synthetic as S, types as T
t = time.time(); w = S.make_window(sys.argv[1] if len(sys.argv) > 1 else "C3"); print("gen", round(time.time() - t, 2), len(w.points), flush=True)
ctx = od.Context(0)
rs = od.ResidentSweep(w.points, ctx)
for i in range(4):
    n, tm = rs.extract(); print("extract", n, tm, flush=True)
s = rs.fetch()
t = time.time(); sld = od.UpdateSurfelPoses(w.imu, s, ctx=ctx); print("poses wall", round(time.time() - t, 4))
fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
print("S", len(sld), "S_fix", len(fix), flush=True)
m = od.KnnSurfelMatcher(ctx); m.BuildIndex(sld)
for i in range(2):
    tm = {}; cs, _ = m.Match(sld, timing=tm); print("match sld", len(cs), tm, flush=True)
m2 = od.KnnSurfelMatcher(ctx); m2.BuildIndex(fix); tm = {}; cf, _ = m2.Match(sld, timing=tm); print("match fix", len(cf), tm, flush=True)
rw = od.ResidentWindow(sld, fix, cs, cf, w.imu, w.samples, ctx)
for i in range(3):
    x, sg = rw.solve(); print("solve iters", sg.num_iterations, "ms", round(sg.gpu_ms_total, 3), "cost", sg.initial_cost, sg.final_cost, T.TERMINATION[sg.termination], flush=True)
err0 = np.linalg.norm(w.samples["pos"][-1] - w.truth_sample_pos[-1]); err1 = np.linalg.norm(w.samples["pos"][-1] + x[-1, 3:6] - w.truth_sample_pos[-1])
print("pos err", err0, "->", err1)
