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


def exchange_handles(handle, dist, device=None):
    """all-gather of the 64-byte exchange-buffer handles; returns a (world, 64) uint8 array identical on every rank."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(handle, dtype=np.uint8))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()
