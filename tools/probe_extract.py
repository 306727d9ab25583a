import sys
import numpy as np
sys.path.insert(0, ".")
from wildcat_slam_b200 import odometry as od, synthetic as S
w = S.make_window(sys.argv[1] if len(sys.argv) > 1 else "C1")
ctx = od.Context(0)
g, a = od.BuildSurfels(w.points, ctx=ctx, want_assign=True)
print("surfels", len(g))
