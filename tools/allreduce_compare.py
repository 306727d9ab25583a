"""Exchange step of the sharded solve, side by side (run under torchrun, one rank per GPU): the library's own peer-memory
reduction kernel (comm_allreduce: flag round + rank-ordered sum over cudaIpc-mapped buffers, packed lower triangle) and
ncclAllReduce (torch.distributed, sum, fp64) on a buffer of the same payload, for the C3 (12 poses) and C5 (64 poses)
normal equations.  Prints microseconds per call, max over ranks."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from wildcat_slam_b200 import odometry as od, sharding, types as T  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prm = T.default_params()
ctx = od.Context(local, params=prm)
ctx.comm_connect(rank, world, sharding.exchange_handles(ctx.comm_export(), dist, device="cuda"))
reps = 200
for K in (12, 64):
    N = 12 * K
    n_packed = N * (N + 1) // 2 + N + 1
    ms = C.c_double(0)
    ctx.check(ctx.lib.wc_comm_bench(ctx.handle, K, reps, C.byref(ms)), "wc_comm_bench")
    buf = torch.zeros(n_packed, dtype=torch.float64, device="cuda")
    for _ in range(5):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ms.value * 1e3, e0.elapsed_time(e1) / reps * 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world {world}  K {K:3d}  payload {n_packed * 8 / 1024:8.1f} KiB   peer-memory kernel {t[0].item():7.1f} us   ncclAllReduce {t[1].item():7.1f} us",
              flush=True)
dist.barrier()
ctx.comm_disconnect()
dist.destroy_process_group()
