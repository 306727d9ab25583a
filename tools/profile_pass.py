"""Runs the device-resident C3 window pass a few times (target of the ncu captures under profiles/)."""
import sys
sys.path.insert(0, ".")
from wildcat_slam_b200 import odometry as od, synthetic as S
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = S.make_window(cfg)
ctx = od.Context(0)
fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
rp = od.ResidentPass(w.points, w.imu, w.samples, fix, ctx=ctx)
for i in range(reps):
    x, summ, st = rp.run()
    print("pass", i, "ms", round(st.ms_total, 3), "extract", round(st.ms_extract, 3), "match", round(st.ms_match, 3), "solve", round(st.ms_solve, 3),
          "iters", summ.num_iterations, "launches", st.n_launches, flush=True)
