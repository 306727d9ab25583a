"""The device path computes VoxelLoc = floor(p / (double)0.8f) (surfel_extraction.h:59-64), the two octree child codes
(surfel_extraction.cc:148-166) and the fixed-point offset from the leaf-cell centre in float32 + integer arithmetic
(wc_extract.cu: axis_cell).  This checks the arithmetic claim behind it against the reference's own double-precision
formulas, in numpy: same IEEE operations (the float32 fused multiply-add is emulated through an 80-bit intermediate),
on random float32 coordinates, on the float32 neighbours of every cell face, and for other voxel sizes."""
import numpy as np

F = np.float32
SCALE = 134217728.0  # 2^27


def fma32(a, b, c):
    """RN_f32(a * b + c): the product of two float32 is exact in 80-bit, the sum of values this close as well."""
    return (a.astype(np.longdouble) * b.astype(np.longdouble) + c.astype(np.longdouble)).astype(F)


def axis_cell(x, vsf):
    """numpy restatement of the device function, operation by operation."""
    v4, v8 = F(vsf / F(4)), F(vsf / F(8))
    inv_v4 = F(1.0) / v4
    magic = F(12582912.0)
    y = (x * inv_v4).astype(F)
    gm = (y + magic).astype(F)
    g0 = gm.view(np.int32) - np.int32(0x4B400000)
    gf = (gm - magic).astype(F)
    r0 = fma32(-gf, np.full_like(x, v4), x)
    dec = (r0 < 0) | ((r0 == 0) & ((g0 & 3) != 0))  # below the rounded cell, or on a face inside the voxel
    g = g0 - dec.astype(np.int32)                    # leaf-cell index along the axis = 4 voxel + child
    gf = np.where(dec, gf - F(1), gf).astype(F)
    q = g >> 2
    f = g & 3
    gc = (gf * F(2) + F(1)).astype(F)
    rel = np.rint(fma32(-gc, np.full_like(x, v8), x).astype(np.float64) * SCALE).astype(np.int64)
    return q, f, rel


def reference_cell(x, vsf):
    """surfel_extraction.h:59-64 + .cc:148-166,209-211 in double, as the reference evaluates them."""
    xd = x.astype(np.float64)
    v = np.float64(vsf)
    q0, q1 = np.float64(F(vsf / F(4))), np.float64(F(F(vsf / F(4)) / F(2)))
    k = np.floor(xd / v)
    c = (0.5 + k) * v
    b1 = xd > c
    c = c + np.where(b1, q0, -q0)
    b2 = xd > c
    c = c + np.where(b2, q1, -q1)
    return k.astype(np.int64), 2 * b1.astype(np.int64) + b2.astype(np.int64), (xd - c) * SCALE


def _inputs(vsf, rng):
    v4 = np.float64(F(vsf / F(4)))
    xs = [rng.uniform(-130.0, 130.0, 2_000_000).astype(F), (rng.standard_normal(200_000) * 1e-3).astype(F),
          rng.uniform(-0.3, 0.3, 200_000).astype(F),
          np.array([0.0, -0.0, 1e-38, -1e-38, 1e-45, -1e-45, 0.8, -0.8, 120.0, -120.0, 13107.0, -13107.0], F)]
    # float32 neighbours of every cell face k * v4 (|k| up to beyond the +-16384-voxel key range of the device path)
    k = np.arange(-80000, 80001, dtype=np.float64)
    for start in ((k * v4).astype(F),):
        b = start.copy()
        for _ in range(3):
            xs.append(b.copy())
            b = np.nextafter(b, F(np.inf))
        b = start.copy()
        for _ in range(3):
            b = np.nextafter(b, F(-np.inf))
            xs.append(b.copy())
    return np.concatenate(xs)


def test_axis_cell_matches_the_reference_formulas():
    rng = np.random.default_rng(7)
    for vsf in (F(0.8), F(0.4), F(0.2), F(1.0), F(0.1)):
        x = _inputs(vsf, rng)
        q, f, rel = axis_cell(x, vsf)
        k, child, rel_ref = reference_cell(x, vsf)
        assert np.array_equal(q, k), vsf                      # VoxelLoc: bit exact
        assert np.array_equal(f, child), vsf                  # both child codes: bit exact (strict '>' on the faces)
        # offset from the leaf centre in units of 2^-27 m: for |x| >= 2^-4 m the float32 value is exact, so the only
        # rounding is the final one to an integer (none at all for the reference's 0.8 m voxels, whose cell centres are
        # multiples of 2^-27 m); closer to the axis the float32 rounding adds at most a quarter unit
        err = np.abs(rel - rel_ref)
        far = np.abs(x) >= 2.0 ** -4
        assert err[far].max() <= 0.5, vsf
        if vsf == F(0.8):
            assert (err[far] == 0).all()
        assert err.max() <= 0.75, (vsf, err.max())
        assert np.abs(rel).max() <= np.float64(F(vsf / F(8))) * SCALE + 1
