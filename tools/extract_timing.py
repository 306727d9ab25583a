"""Times the resident surfel extraction (K1 cell_key_link / the rest) on a config and prints the stage medians."""
import sys

sys.path.insert(0, ".")
import numpy as np

from wildcat_slam_b200 import odometry as od, synthetic as S

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
w = S.make_window(cfg)
ctx = od.Context(0)
rs = od.ResidentSweep(w.points, ctx=ctx)
ks, es, ts = [], [], []
for rep in range(reps):
    n, st = rs.extract()
    ks.append(st["keys_ms"]), es.append(st["emit_ms"]), ts.append(st["total_ms"])
s = rs.fetch()
print(f"{cfg}: K1 {np.median(ks)*1e3:.1f} us (min {min(ks)*1e3:.1f})  rest {np.median(es)*1e3:.1f} us  total {np.median(ts)*1e3:.1f} us  surfels {len(s)}", flush=True)
ctx.close()
