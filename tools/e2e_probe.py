"""Streaming end-to-end step time on C3 (one GPU): upload (prefetched) + pass, as bench.py's e2e leg, without the rest."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from wildcat_slam_b200 import odometry as od, synthetic as S, types as T
w = S.make_window("C3")
ctx = od.Context(0)
fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
bufs = []
for _ in range(2):
    p = ctx.pinned(len(w.points), T.POINT48); p[:] = w.points; bufs.append(p)
fix_p = ctx.pinned(len(fix), T.SURFEL); fix_p[:] = fix
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def step(k, stream=True):
    flush.zero_(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    rp = od.ResidentPass(bufs[k & 1], w.imu, w.samples, fix_p, ctx=ctx, keep_fix=k > 0)
    if stream: ctx.prefetch(bufs[(k + 1) & 1], at_solve=stream == 2)
    x, sg, st = rp.run()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, st.ms_total, st.ms_extract, st.ms_match, st.ms_pack, st.ms_solve
for mode in (2, 1, 0):
    for k in range(3): step(k, mode)
    r = [step(k, mode) for k in range(3, 13)]
    m = np.mean(np.array(r), axis=0)
    print(f"stream={mode}: step {1e3*m[0]:.3f} ms (pass device time {m[1]:.3f} ms: extract {m[2]:.3f} match {m[3]:.3f} pack {m[4]:.3f} solve {m[5]:.3f})", flush=True)
