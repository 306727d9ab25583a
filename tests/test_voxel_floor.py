"""The device path computes VoxelLoc = floor(p / (double)0.8f) (surfel_extraction.h:59-64) without a division
(wc_extract.cu: voxel_floor).  This checks the arithmetic claim behind it in numpy fp64 (same IEEE operations, no
contraction): reciprocal-multiply proposal + two exact product comparisons == floor of the rounded quotient, on random
float32 coordinates and on the float32 neighbours of every voxel boundary."""
import numpy as np


def voxel_floor(x, v):
    inv_v = 1.0 / v
    q = np.floor(x * inv_v)
    q = np.where(x < q * v, q - 1.0, np.where(x >= (q + 1.0) * v, q + 1.0, q))
    return q


def test_voxel_floor_is_bit_exact():
    v = np.float64(np.float32(0.8))
    rng = np.random.default_rng(7)
    xs = [rng.uniform(-130.0, 130.0, 2_000_000).astype(np.float32),
          (rng.standard_normal(200_000) * 1e-3).astype(np.float32),
          np.array([0.0, -0.0, 1e-38, -1e-38, 1e-45, -1e-45, 0.8, -0.8, 120.0, -120.0, 13107.0, -13107.0], np.float32)]
    # float32 neighbours of every boundary k * v, |k| <= 20000 (beyond the +-16384-voxel key range of the device path)
    k = np.arange(-20000, 20001, dtype=np.float64)
    b = (k * v).astype(np.float32)
    for _ in range(3):
        xs.append(b.copy())
        b = np.nextafter(b, np.float32(np.inf))
    b = (k * v).astype(np.float32)
    for _ in range(3):
        b = np.nextafter(b, np.float32(-np.inf))
        xs.append(b.copy())
    for other in (np.float32(0.4), np.float32(0.2), np.float32(1.0), np.float32(0.1)):  # other voxel sizes
        vo = np.float64(other)
        x = rng.uniform(-100.0, 100.0, 500_000).astype(np.float32).astype(np.float64)
        assert np.array_equal(voxel_floor(x, vo), np.floor(x / vo))
    x = np.concatenate(xs).astype(np.float64)
    assert np.array_equal(voxel_floor(x, v), np.floor(x / v))
