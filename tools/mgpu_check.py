"""Multi-GPU residual sharding check (run under torchrun, one rank per GPU): the sharded solve must give bitwise
identical corrections on every rank and agree with the single-GPU solve."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from wildcat_slam_b200 import odometry as od, sharding, synthetic as S  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
w = S.make_window(cfg)
ctx = od.Context(local)
sld = od.UpdateSurfelPoses(w.imu, od.BuildSurfels(w.points, ctx=ctx), ctx=ctx)
fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
m = od.KnnSurfelMatcher(ctx); m.BuildIndex(sld); cs, _ = m.Match(sld)
m2 = od.KnnSurfelMatcher(ctx); m2.BuildIndex(fix); cf, _ = m2.Match(sld)
x1, s1 = od.ResidentWindow(sld, fix, cs, cf, w.imu, w.samples, ctx).solve()      # single GPU (world == 1 in the ctx)
handles = sharding.exchange_handles(ctx.comm_export(), dist, device="cuda")
ctx.comm_connect(rank, world, handles)
# sharded matcher: k-NN queries split over the ranks, index lists all-gathered over NVLink -> identical pair lists
m = od.KnnSurfelMatcher(ctx); m.BuildIndex(sld); cs2, _ = m.Match(sld)
m2 = od.KnnSurfelMatcher(ctx); m2.BuildIndex(fix); cf2, _ = m2.Match(sld)
assert cs2.tobytes() == cs.tobytes() and cf2.tobytes() == cf.tobytes(), "sharded matcher differs from the single-GPU matcher"
# sharded sweep upload: each rank copies its slab over PCIe, the repack reads every point from its owner over NVLink ->
# bitwise the same surfels as the plain upload, with and without the prefetch
from wildcat_slam_b200 import types as T  # noqa: E402
rs0 = od.ResidentSweep(w.points, ctx=ctx); rs0.extract(); g0 = rs0.fetch()
ctx.comm_shard_upload(True)
rs1 = od.ResidentSweep(w.points, ctx=ctx); rs1.extract(); g1 = rs1.fetch()
pin = ctx.pinned(len(w.points), T.POINT48); pin[:] = w.points
ctx.prefetch(pin)
rs2 = od.ResidentSweep(pin, ctx=ctx); rs2.extract(); g2 = rs2.fetch()
rs3 = od.ResidentSweep(w.points[: len(w.points) // 3], ctx=ctx); n3, _ = rs3.extract()   # another size, other parity
ctx.comm_shard_upload(False)
rs4 = od.ResidentSweep(w.points[: len(w.points) // 3], ctx=ctx); n4, _ = rs4.extract()
assert g1.tobytes() == g0.tobytes() and g2.tobytes() == g0.tobytes() and n3 == n4 and rs3.fetch().tobytes() == rs4.fetch().tobytes(), \
    "sharded sweep upload differs from the plain upload"
print(f"[rank {rank}] sharded upload identical: {len(g1)} surfels", flush=True)
rw = od.ResidentWindow(sld, fix, cs, cf, w.imu, w.samples, ctx)                  # sharded: this rank packs its block only
for rep in range(3):
    xs, ss = rw.solve()
t = torch.from_numpy(xs.copy()).cuda()
allx = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(allx, t)
same = all(torch.equal(allx[0], a) for a in allx)
d = float(np.abs(xs - x1).max())
print(f"[rank {rank}] sharded matcher identical: {len(cs2)}+{len(cf2)} pairs", flush=True)
print(f"[rank {rank}] iters single={s1.num_iterations} sharded={ss.num_iterations} cost single={s1.final_cost:.12g} sharded={ss.final_cost:.12g} "
      f"|x_sharded - x_single|max={d:.3e} bitwise_identical_across_ranks={same} solve_ms single={s1.gpu_ms_total:.3f} sharded={ss.gpu_ms_total:.3f}", flush=True)
assert same and d < 1e-9 and ss.num_iterations == s1.num_iterations
dist.barrier()
ctx.comm_disconnect()
dist.destroy_process_group()
