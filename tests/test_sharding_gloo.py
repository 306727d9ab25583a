"""The N > 1 host logic on CPU: world_size-2 gloo processes shard the correspondences exactly like the CUDA path does,
linearise their block with the oracle, all-reduce the normal equations and must reproduce the unsharded evaluation;
the IPC-handle exchange returns the same table on every rank."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import wc_oracle as O
    from wildcat_slam_b200 import sharding
    from wildcat_slam_b200 import synthetic as S
    from wildcat_slam_b200 import types as T

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = S.make_window("C1")
    sld = O.update_surfel_poses(w.imu, O.build_surfels(w.points)["surfels"])[1]
    fix = O.update_surfel_poses(w.fix_imu, O.build_surfels(w.fix_points)["surfels"])[1]
    cs, _ = O.match(sld, sld, True)
    cf, _ = O.match(sld, fix, False)
    rng = np.random.default_rng(11)
    smp = w.samples.copy()
    smp["data_cor"] = rng.normal(size=(len(smp), 12)) * 1e-3
    my_cs, my_cf = sharding.split_corrs(cs, cf, rank, world)
    o = T.default_solve_opts()
    o.use_imu_factors = 1 if rank == 0 else 0  # IMU factors live on rank 0
    st, c, g, H = O.window_evaluate(sld, fix, my_cs, my_cf, w.imu if rank == 0 else None, smp, opts=o)
    assert st == 0
    packed = torch.from_numpy(np.concatenate([H.reshape(-1), g, [c]]))
    dist.all_reduce(packed)  # the exchange step (NVLink peer-memory reduction on the GPU path)
    st, c_f, g_f, H_f = O.window_evaluate(sld, fix, cs, cf, w.imu, smp)
    n = len(g_f)
    np.testing.assert_allclose(packed[:n * n].numpy().reshape(n, n), H_f, rtol=0, atol=1e-9 * np.abs(H_f).max())
    np.testing.assert_allclose(packed[n * n:n * n + n].numpy(), g_f, rtol=0, atol=1e-9 * np.abs(g_f).max())
    assert abs(packed[-1].item() - c_f) <= 1e-12 * c_f
    counts = torch.tensor([len(my_cs), len(my_cf)])
    dist.all_reduce(counts)
    assert counts.tolist() == [len(cs), len(cf)]
    handles = sharding.exchange_handles(np.full(64, rank + 1, dtype=np.uint8), dist)
    assert handles.shape == (world, 64) and all((handles[r] == r + 1).all() for r in range(world))
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


def test_shard_ranges_partition_exactly():
    sys.path.insert(0, ROOT)
    from wildcat_slam_b200 import sharding

    for C in (0, 1, 7, 90424, 10_000_000):
        for world in (1, 2, 4, 8):
            edges = [sharding.shard_range(C, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == C
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            assert max(b - a for a, b in edges) - min(b - a for a, b in edges) <= 1


def test_two_rank_sharded_normal_equations(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_sweep_slab_owner_matches_the_block_partition():
    """the owner lookup of the sharded sweep upload (one guess + one correction step, as in repack_points_sharded) agrees
    with the block partition shard_range for every point, also when slabs are empty (n < world) or uneven"""
    from wildcat_slam_b200 import sharding

    rng = np.random.default_rng(5)
    cases = [(1, 2), (3, 8), (7, 8), (8, 8), (9, 8), (100_003, 3), (2_000_000, 8), (393_216, 7)]
    cases += [(int(rng.integers(1, 50_000)), int(rng.integers(2, 9))) for _ in range(40)]
    for n, world in cases:
        i = np.arange(n) if n <= 200_000 else rng.integers(0, n, 200_000)
        owner = sharding.slab_owner(i, n, world)
        lo = np.array([sharding.shard_range(n, r, world)[0] for r in range(world)] + [n])
        want = np.searchsorted(lo, i, side="right") - 1
        # empty slabs share a boundary: the partition's owner is the LAST rank starting at or before i
        assert np.array_equal(lo[owner] <= i, np.ones_like(i, bool)) and np.array_equal(i < lo[owner + 1], np.ones_like(i, bool)), (n, world)
        assert np.array_equal(owner, want), (n, world)
