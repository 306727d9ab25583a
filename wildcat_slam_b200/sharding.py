"""Host-side logic of the multi-GPU residual sharding: which correspondences a rank linearises and how the
ranks exchange the IPC handles of their exchange buffers.  The slicing mirrors wc_window_prepare_device (wc_solve.cu)."""
import numpy as np


def shard_range(n_corr, rank, world):
    """contiguous block partition of the concatenated (sliding ++ fixed) correspondence list."""
    return (n_corr * rank) // world, (n_corr * (rank + 1)) // world


def split_corrs(sld_corr, fix_corr, rank, world):
    """the (sliding, fixed) correspondences of this rank's block."""
    ns, nf = len(sld_corr), len(fix_corr)
    c0, c1 = shard_range(ns + nf, rank, world)
    return sld_corr[min(c0, ns):min(c1, ns)], fix_corr[max(c0 - ns, 0):max(c1 - ns, 0)]


def slab_owner(i, n, world):
    """rank whose slab of a sharded sweep upload holds point i (wc_comm_shard_upload: rank r copies points
    [n r // world, n (r + 1) // world); repack_points_sharded, wc_extract.cu, finds the owner the same way: first guess
    i world // n, then steps down / up until the slab brackets i — more than one step only when slabs are empty)."""
    i = np.asarray(i, dtype=np.int64)
    r = (i * world) // n
    lo = lambda q: (n * q) // world  # noqa: E731
    for _ in range(world):
        r = np.where((r > 0) & (i < lo(r)), r - 1, r)
    for _ in range(world):
        r = np.where((r + 1 < world) & (i >= lo(np.minimum(r + 1, world))), r + 1, r)
    return r


def exchange_handles(handle, dist, device=None):
    """all-gather of the 64-byte exchange-buffer handles; returns a (world, 64) uint8 array identical on every rank."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(handle, dtype=np.uint8))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()
