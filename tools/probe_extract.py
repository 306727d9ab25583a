"""Runs the resident surfel extraction a few times on a config (target of the per-kernel ncu captures)."""
import sys
sys.path.insert(0, ".")
from wildcat_slam_b200 import odometry as od, synthetic as S
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = S.make_window(cfg)
ctx = od.Context(0)
rs = od.ResidentSweep(w.points, ctx=ctx)
for rep in range(reps):
    n, st = rs.extract()
    print(rep, n, {k: round(v, 4) for k, v in st.items()}, flush=True)
ctx.close()
