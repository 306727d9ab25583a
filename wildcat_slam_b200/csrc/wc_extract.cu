// Surfel extraction on the device — replaces BuildSurfels / BuildVoxelMap / OctoTree::{InitOctoTree, CutOctoTree,
// InitPlane, ExtractSurfelInfo} / ClusterSurfels (src/odometry/surfel_extraction.cc:12-65,82-220,304-337).
//
//   K0 repack_points        48-byte hilti_ros::Point records -> float4 xyz + double t (resident layout)
//   K1 voxel_key_moments    per point: VoxelLoc (fp64 floor(p/(double)0.8f)), two octree child codes, time bin;
//                           open-addressed hash of (voxel, leaf cell, time bin) -> slot; exact int64 moments by
//                           native RED atomics after warp-level aggregation of equal keys
//   K2a..c voxel_index / scan / scatter     group the occupied slots by voxel
//   K2  cluster_eig_emit    per voxel: octree levels from additive moments, planarity flags, exact time clusters
//                           (gap test on the original fp64 timestamps), 3x3 Jacobi eigen-solve, Surfel records
//   K2s bitonic sort        final order by (timestamp, resolution desc, center.x)  (surfel_extraction.cc:334)
//
// Exactness: the time clusters of a node are maximal runs of its time-ordered points with consecutive gaps
// <= 0.05 s (surfel_extraction.cc:22-29).  Points sharing a 2^-5 s bin can never be split (spread < gap), and a
// boundary between consecutive non-empty bins exists iff t_min(next) - t_max(prev) > gap — the same fp64
// subtraction on the same two operands the reference performs, because those two points are consecutive in the
// node's time-ordered list.
#include <math.h>
#include <stdlib.h>

#include "wc_ctx.h"
#include "wc_device_math.cuh"

using namespace wcd;

namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

struct ExtractParams {
  double t_first;
  float  v4, inv_v4, v8;  // leaf-cell width voxel/4 (exact in float32), its rounded reciprocal, half width voxel/8
  int    vox0[3];
  int    goffm[3];        // 4 (WC_VOX_BIAS - vox0) - 0x4B400000: leaf-cell index bias of the tile keys, minus the magic exponent
  int    n;
};

// axis_cell: x float32, v4 = voxel/4 (exact in float32).  y = RN(x * RN(1/v4)) is within 2^-6 of x / v4 (|x / v4| < 2^22),
// so g0 = round-to-nearest(y) (magic-number add) is floor(x / v4) or floor + 1.  r0 = fma(-g0, v4, x) is the CORRECTLY
// ROUNDED value of x - g0 v4: its sign is exact (and it is zero iff x lies exactly on the cell face g0 v4), hence
// g = g0 - [r0 < 0] is exactly floor(x / v4).  The reference (surfel_extraction.h:59-64, .cc:148-166,209-211) evaluates
// floor(x / v), (0.5 + k) v, c +- v/4, c +- v/8 in double; all of these are exact there (products of a <= 22-bit integer
// and the 24-bit v), so its cell decisions are the exact ones too: voxel = g >> 2, child index along the axis g & 3 —
// except ON a face (r0 == 0), where the strict '>' sends the point to the lower cell unless that would leave the voxel
// (g0 & 3 == 0).  Both corrections are the same decrement, so the leaf-cell index along the axis is
// g' = g0 - [r0 < 0 or (r0 == 0 and g0 & 3 != 0)] = 4 voxel + child.  The fixed-point offset from the leaf centre
// (2 g' + 1) v8 is one more fused multiply-add: exact whenever |x| >= 2^-4 m (the offset is then a multiple of 2^-27
// below 2^-3), rounded by at most 0.75 * 2^-27 m closer to the axis otherwise.  tests/test_voxel_floor.py checks all of
// this against the reference's double-precision formulas.  Returns g' + bias in `a`.
__device__ __forceinline__ bool axis_cell(float x, const ExtractParams& P, int goffm, int& a, int& rel) {
  const float MAGIC = 12582912.f;  // 1.5 * 2^23
  const float y     = __fmul_rn(x, P.inv_v4);
  const float gm    = __fadd_rn(y, MAGIC);
  float       gf    = __fadd_rn(gm, -MAGIC);
  const float r0    = __fmaf_rn(-gf, P.v4, x);
  const int   a0    = __float_as_int(gm) + goffm;  // g0 + bias (bias and exponent are multiples of 4: a0 & 3 == g0 & 3)
  const bool  dec   = (r0 < 0.f) || (r0 == 0.f && (a0 & 3) != 0);
  a                 = a0;
  if (dec) a = a0 - 1, gf = __fadd_rn(gf, -1.f);
  const float gc = __fmaf_rn(gf, 2.f, 1.f);  // 2 g' + 1: exact
  rel            = __float2int_rn(__fmul_rn(__fmaf_rn(-gc, P.v8, x), (float)WC_COORD_SCALE));
  return fabsf(y) < 4.0e6f;  // also false for NaN / Inf
}

// ---------------------------------------------------------------------------------------------- K0
__global__ void repack_points(const wc_point48* __restrict__ raw, int n, float4* __restrict__ xyz,
                              double* __restrict__ t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4* r = reinterpret_cast<const float4*>(raw + i);
  float4        a = r[0];
  float4        b = r[1];  // intensity, pad, time (as two floats)
  xyz[i]          = a;
  t[i]            = __hiloint2double(__float_as_int(b.w), __float_as_int(b.z));
}

// sharded upload: point i sits in the raw area of the rank that copied its slab; read it there (NVLink peer load,
// cache-volatile: a peer's memory may change between sweeps) and repack into this rank's resident layout
struct RawAreas {
  const wc_point48* area[8];
  int               row0[9];  // slab of rank r: points [row0[r], row0[r + 1])
  int               world;
};
__global__ void repack_points_sharded(RawAreas A, int n, float4* __restrict__ xyz, double* __restrict__ t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = (int)(((long long)i * A.world) / n);
  while (r > 0 && i < A.row0[r]) --r;
  while (r + 1 < A.world && i >= A.row0[r + 1]) ++r;
  const uint4* q = reinterpret_cast<const uint4*>(A.area[r] + i);
  const uint4  a = __ldcv(q), b = __ldcv(q + 1);  // b: intensity, pad, time (as two words)
  xyz[i]         = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
  t[i]           = __hiloint2double((int)b.w, (int)b.z);
}

// ---------------------------------------------------------------------------------------------- K1
// Cell table entry: key and published slot id share one 16-byte word, so a probe touches one sector.
struct __align__(16) HEnt {
  unsigned long long key;
  int                slot;  // -1 until the creator has written its slot record
  int                pad;
};

// Tile kernel: a CTA takes KT consecutive points.  Each point computes its cell key (leaf-cell index per axis + time
// bin) and its exact fixed-point payload, finds the key's position in a shared-memory hash table (one 64-bit CAS for the
// first point of a key) and pushes itself onto that position's list with ONE native 32-bit exchange; the payload, with
// the link to the previous list head packed into its spare bits, goes to shared memory.  The first point of a key also
// appends the position to the tile's compact run list.  After one barrier a thread per run walks its list, sums the
// points in exact 64-bit integers (order-independent, so the records are bitwise reproducible whatever order the
// exchanges happened in) and writes ONE 128-byte run record with plain stores (the tile's records are contiguous: one
// counter bump per tile reserves them).  No counting sort, no scan, no global atomics, fences or dependent global probes
// here: a cell that straddles a tile boundary simply yields two records, which voxel_index merges through the global
// cell table afterwards.  A spinning lidar revisits a 0.2 m cell with a handful of adjacent rings x azimuth columns, all
// inside one tile, so the record traffic per point drops by the run length (4-6x at C3).
// 72 KB of shared memory and <= 40 registers: three CTAs (48 warps) per SM, so the key phase of one tile overlaps the
// list walks of another.
constexpr int KT = 2048, KNT = 512, KPT = KT / KNT, KTAB = 2816;
constexpr unsigned K_END = 0xFFFFu;  // list terminator (links are point ids x 16 = byte offsets into pay[], < 2^15)
constexpr int K1_OFF_HEAD = KTAB * 8, K1_OFF_PAY = KTAB * 12, K1_OFF_NXT = K1_OFF_PAY + KT * 16, K1_OFF_RUN = K1_OFF_NXT + KT * 2;
constexpr int K1_SMEM = K1_OFF_RUN + KT * 2;
static_assert(K1_OFF_PAY % 16 == 0, "table region is cleared with 16-byte stores");

__global__ void __launch_bounds__(KNT, 3)
voxel_key_moments(const float4* __restrict__ xyz, const double* __restrict__ time, ExtractParams P, wc_slot_planes slots,
                  int slot_cap, wc_extract_status* __restrict__ st, wc_point_assign* __restrict__ assign) {
  extern __shared__ __align__(16) unsigned char k1_smem[];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(k1_smem);           // KTAB keys (all-ones = empty)
  unsigned*           head = reinterpret_cast<unsigned*>(k1_smem + K1_OFF_HEAD);        // KTAB list heads (all-ones = none)
  unsigned char*      payb = k1_smem + K1_OFF_PAY;                                      // KT payloads int4 (x, y, z, t)
  unsigned char*      nxtb = k1_smem + K1_OFF_NXT;                                      // KT links, u16
  unsigned short*     runh = reinterpret_cast<unsigned short*>(k1_smem + K1_OFF_RUN);   // table position of each run
  __shared__ int      s_nruns, s_base;
  const int t = threadIdx.x;
  const int base = blockIdx.x * KT;

  for (int k = t; k < K1_OFF_PAY / 16; k += KNT) reinterpret_cast<int4*>(k1_smem)[k] = make_int4(-1, -1, -1, -1);
  if (t == 0) s_nruns = 0, s_base = -1;
  // the tile's points: all loads first, so that they are in flight together
  float4 p[KPT];
  double tt[KPT], tp[KPT];
#pragma unroll
  for (int u = 0; u < KPT; ++u) {
    const int i = base + t + KNT * u;
    if (i < P.n) {
      p[u]  = xyz[i];
      tt[u] = time[i];
      tp[u] = i > 0 ? time[i - 1] : tt[u];
    }
  }
  __syncthreads();

#pragma unroll
  for (int u = 0; u < KPT; ++u) {
    const int l = t + KNT * u, i = base + l;
    if (i >= P.n) continue;
    if (tt[u] < tp[u]) st->err_time_order = 1;  // CHECK lidar_odometry.cc:491
    // VoxelLoc (surfel_extraction.h:59-64), the two child descents (.cc:148-166) and the exact fixed-point offset
    // from the leaf-cell centre, in float32 / integer arithmetic (axis_cell)
    int  ax, ay, az;
    int4 q;
    bool ok = axis_cell(p[u].x, P, P.goffm[0], ax, q.x);
    ok      = axis_cell(p[u].y, P, P.goffm[1], ay, q.y) && ok;
    ok      = axis_cell(p[u].z, P, P.goffm[2], az, q.z) && ok;
    if (assign) {
      const int fx = ax & 3, fy = ay & 3, fz = az & 3;
      const int leaf = ((fx >> 1) << 5) | ((fy >> 1) << 4) | ((fz >> 1) << 3) | ((fx & 1) << 2) | ((fy & 1) << 1) | (fz & 1);
      assign[i] = wc_point_assign{(ax >> 2) - WC_VOX_BIAS + P.vox0[0], (ay >> 2) - WC_VOX_BIAS + P.vox0[1],
                                  (az >> 2) - WC_VOX_BIAS + P.vox0[2], leaf};
    }
    // time: Q = rint((t - t_first) 2^36) read off the mantissa of (t - t_first) + 1.5 * 2^16 (ulp 2^-36, ties to even like
    // rint); bin = Q >> 31, valid iff 0 <= Q < 2^43 (128 s)
    const double dm  = __dadd_rn(__dadd_rn(tt[u], -P.t_first), 98304.0);
    const int    qlo = __double2loint(dm), qhi = __double2hiint(dm) - 0x40F80000;
    const int    bin = (qhi << 1) | (int)((unsigned)qlo >> 31);
    q.w              = qlo & 0x7fffffff;
    if (!ok || (((unsigned)ax | (unsigned)ay | (unsigned)az) >> 17) != 0u || (unsigned)qhi >= (unsigned)(WC_MAX_BINS >> 1)) {
      st->err_range = 1;
      continue;
    }
    const unsigned klo = ((unsigned)ay << 29) | ((unsigned)az << 12) | (unsigned)bin;
    const unsigned khi = ((unsigned)ax << 14) | ((unsigned)ay >> 3);
    const unsigned long long key = ((unsigned long long)khi << 32) | klo;
    unsigned m = klo * 0x9E3779B1u + khi * 0x85EBCA77u;  // tile-local table: a cheap mix
    m          = (m ^ (m >> 15)) * 0x2C1B3C6Du;
    unsigned h = __umulhi(m, (unsigned)KTAB);
    bool created = false;
    for (;;) {  // at most KT distinct keys in KTAB positions: always terminates
      unsigned long long k = skey[h];
      if (k == key) break;
      if (k == WC_KEY_EMPTY) {
        k = atomicCAS(&skey[h], WC_KEY_EMPTY, key);
        if (k == WC_KEY_EMPTY) {
          created = true;
          break;
        }
        if (k == key) break;
      }
      h = h + 1 == KTAB ? 0u : h + 1;
    }
    const unsigned l16  = (unsigned)l << 4;
    const unsigned prev = atomicExch(&head[h], l16);
    *reinterpret_cast<int4*>(payb + l16)                = q;
    *reinterpret_cast<unsigned short*>(nxtb + (l << 1)) = (unsigned short)prev;
    if (created) {
      // plain shared atomic (inline PTX: the compiler's warp-aggregated form costs ~25 instructions in divergent code)
      unsigned rid;
      asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(rid) : "r"((unsigned)__cvta_generic_to_shared(&s_nruns)) : "memory");
      runh[rid] = (unsigned short)h;
    }
  }
  __syncthreads();
  const int nruns = s_nruns;
  // reserve one slot id per run of this tile: thread 0 publishes the base through shared memory; it is needed only after
  // the walks, by which time the counter bump has long returned (no second barrier: a warp writes its records as soon
  // as its own lists are summed)
  if (t == 0) atomicExch(&s_base, atomicAdd(&st->n_slots, nruns));  // (shared-memory atomics on both sides of the hand-over)
  // ---- one thread per run: walk its list and sum it
  for (int r0 = 0; r0 < nruns; r0 += KNT) {
    const int r = r0 + t;
    long long a_t = 0, a_x = 0, a_y = 0, a_z = 0, a_xx = 0, a_xy = 0, a_xz = 0, a_yy = 0, a_yz = 0, a_zz = 0;
    unsigned  lmin = 0xffffffffu, lmax = 0u, np = 0;
    unsigned long long key = 0;
    if (r < nruns) {
      const int h = runh[r];
      key         = skey[h];
      unsigned lo = head[h];
      do {
        const int4     e  = *reinterpret_cast<const int4*>(payb + lo);
        const unsigned nl = *reinterpret_cast<const unsigned short*>(nxtb + (lo >> 3));
        lmin = min(lmin, lo), lmax = max(lmax, lo);
        ++np;
        a_t += e.w, a_x += e.x, a_y += e.y, a_z += e.z;
        a_xx += (long long)e.x * e.x, a_xy += (long long)e.x * e.y, a_xz += (long long)e.x * e.z;
        a_yy += (long long)e.y * e.y, a_yz += (long long)e.y * e.z, a_zz += (long long)e.z * e.z;
        lo = nl;
      } while (lo != K_END);
    }
    if (r >= nruns) continue;
    int sb;
    while ((sb = atomicAdd(&s_base, 0)) < 0) {
    }
    const int fresh = sb + r;
    if (fresh >= slot_cap) {
      st->err_capacity = 1;
      continue;
    }
    // record key in the layout the later stages use: [vx:15][vy:15][vz:15][leaf:6][bin:12]
    const unsigned ax = (unsigned)(key >> 46), ay = (unsigned)(key >> 29) & 0x1ffffu, az = (unsigned)(key >> 12) & 0x1ffffu;
    const unsigned fx = ax & 3u, fy = ay & 3u, fz = az & 3u;
    const unsigned leaf = ((fx >> 1) << 5) | ((fy >> 1) << 4) | ((fz >> 1) << 3) | ((fx & 1u) << 2) | ((fy & 1u) << 1) | (fz & 1u);
    const unsigned long long rkey = ((unsigned long long)(ax >> 2) << 48) | ((unsigned long long)(ay >> 2) << 33) |
                                    ((unsigned long long)(az >> 2) << 18) | ((unsigned long long)leaf << 12) | (key & 4095ull);
    // timestamps are non-decreasing in the point index (checked above): the run's earliest / latest points are its
    // lowest / highest indices
    // consecutive lanes hold consecutive slots: every store below fills half of a 32-byte sector per lane, side by side
    slots.key[fresh]        = rkey;
    slots.p0[2 * fresh]     = make_int4((int)np, base + (int)(lmin >> 4), base + (int)(lmax >> 4), 0);
    slots.p0[2 * fresh + 1] = make_int4((int)a_t, (int)(a_t >> 32), (int)a_x, (int)(a_x >> 32));
    slots.p1[2 * fresh]     = make_int4((int)a_y, (int)(a_y >> 32), (int)a_z, (int)(a_z >> 32));
    slots.p1[2 * fresh + 1] = make_int4((int)a_xx, (int)(a_xx >> 32), (int)a_xy, (int)(a_xy >> 32));
    slots.p2[2 * fresh]     = make_int4((int)a_xz, (int)(a_xz >> 32), (int)a_yy, (int)(a_yy >> 32));
    slots.p2[2 * fresh + 1] = make_int4((int)a_yz, (int)(a_yz >> 32), (int)a_zz, (int)(a_zz >> 32));
  }
}

// ---------------------------------------------------------------------------------------------- K2a-c
// One thread per run record.  (1) Merge records of the same (voxel, leaf cell, time bin) through the global cell table:
// the first record to claim the key becomes the cell's slot, later ones add their exact integer moments to it with
// native 64-bit RED atomics and retire (n = 0).  The records were completed by the previous kernel, so no fences are
// needed — the only dependent chain is one CAS.  (2) Surviving slots register with their voxel (voxel table -> dense
// voxel id, per-voxel slot count).
__global__ void voxel_index(wc_slot_planes slots, wc_extract_status* __restrict__ st, HEnt* __restrict__ tab,
                            unsigned long long capmask, unsigned long long* __restrict__ vkeys, int* __restrict__ vslot,
                            unsigned long long vmask, int* __restrict__ vox_count, unsigned long long* __restrict__ vox_key,
                            int* __restrict__ vox_hpos, int vox_cap, int2* __restrict__ rec_info, int* __restrict__ rec_tpos) {
  const int ns = min(st->n_slots, INT_MAX);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += gridDim.x * blockDim.x) {
    const unsigned long long ckey = slots.key[s];
    int                      owner = -2, tpos = -1;
    {
      unsigned long long h = mix64(ckey) & capmask;
      for (unsigned long long probe = 0; probe <= capmask; ++probe, h = (h + 1) & capmask) {
        const unsigned long long k = atomicCAS(&tab[h].key, WC_KEY_EMPTY, ckey);
        if (k == WC_KEY_EMPTY) {
          *((volatile int*)&tab[h].slot) = s;
          owner                          = s;
          tpos                           = (int)h;
          break;
        }
        if (k == ckey) {
          while ((owner = *((volatile int*)&tab[h].slot)) == -1) {
          }
          break;
        }
      }
    }
    if (owner != s) {
      if (owner < 0) {
        st->err_capacity = 1;
      } else {
        const int4 a0 = slots.p0[2 * s];
        atomicAdd(&slots.p0[2 * owner].x, a0.x);
        atomicMin(&slots.p0[2 * owner].y, a0.y);
        atomicMax(&slots.p0[2 * owner].z, a0.z);
        const unsigned long long* m0 = reinterpret_cast<const unsigned long long*>(slots.p0 + 2 * s + 1);
        const unsigned long long* m1 = reinterpret_cast<const unsigned long long*>(slots.p1 + 2 * s);
        const unsigned long long* m2 = reinterpret_cast<const unsigned long long*>(slots.p2 + 2 * s);
        unsigned long long*       o0 = reinterpret_cast<unsigned long long*>(slots.p0 + 2 * owner + 1);
        unsigned long long*       o1 = reinterpret_cast<unsigned long long*>(slots.p1 + 2 * owner);
        unsigned long long*       o2 = reinterpret_cast<unsigned long long*>(slots.p2 + 2 * owner);
        atomicAdd(o0, m0[0]), atomicAdd(o0 + 1, m0[1]);
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(o1 + k, m1[k]), atomicAdd(o2 + k, m2[k]);
      }
      rec_info[s] = make_int2(-1, 0);  // retired: voxel_scatter skips it
      rec_tpos[s] = -1;
      continue;
    }
    const unsigned long long key = ckey >> 18;
    unsigned long long       h   = mix64(key) & vmask;
    int                      vid = -2;
    for (unsigned long long probe = 0; probe <= vmask; ++probe, h = (h + 1) & vmask) {
      unsigned long long k = *((volatile unsigned long long*)&vkeys[h]);
      if (k == WC_KEY_EMPTY) {
        k = atomicCAS(&vkeys[h], WC_KEY_EMPTY, key);
        if (k == WC_KEY_EMPTY) {
          vid = atomicAdd(&st->n_voxels, 1);
          if (vid >= vox_cap) {
            st->err_capacity = 1;
            vid              = -2;
          } else {
            vox_key[vid]  = key;
            vox_hpos[vid] = (int)h;  // remember the table position for cleanup
          }
          atomicExch(&vslot[h], vid);
          break;
        }
      }
      if (k == key) {
        while ((vid = *((volatile int*)&vslot[h])) == -1) {
        }
        break;
      }
    }
    // the record's rank among its voxel's slots: voxel_scatter places it without another atomic
    rec_info[s] = make_int2(vid, vid >= 0 ? atomicAdd(&vox_count[vid], 1) : 0);
    rec_tpos[s] = tpos;  // the cell-table entry this record claimed (extract_cleanup empties exactly those)
  }
}

// exclusive scan of vox_count[0..n_voxels) by one CTA
__global__ void __launch_bounds__(1024) voxel_scan(const int* __restrict__ cnt, int* __restrict__ off,
                                                   const wc_extract_status* __restrict__ st) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int      nv = st->n_voxels;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nv; base += 1024) {
    const int i    = base + threadIdx.x;
    const int v    = i < nv ? cnt[i] : 0;
    int       incl = v;
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((threadIdx.x & 31) >= d) incl += o;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_sums[threadIdx.x], wi = w;
      for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, wi, d);
        if (threadIdx.x >= d) wi += o;
      }
      warp_sums[threadIdx.x] = wi - w;  // exclusive
    }
    __syncthreads();
    const int excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
    if (i < nv) off[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[nv] = carry;
}

__global__ void voxel_scatter(const int2* __restrict__ rec_info, const wc_extract_status* __restrict__ st, int slot_cap,
                              const int* __restrict__ off, int* __restrict__ seg) {
  const int ns = min(st->n_slots, slot_cap);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += gridDim.x * blockDim.x) {
    const int2 vr = rec_info[s];
    if (vr.x < 0) continue;
    seg[off[vr.x] + vr.y] = s;
  }
}

// ---------------------------------------------------------------------------------------------- K2
struct EmitParams {
  double voxel, q0, q1;
  double t_first;
  double thr;        // (double)planer_threshold (float)
  double min_like;
  double gap;
  double view[3];
  int    vox0[3];
  int    lps[3];     // layer_point_size
  int    cmin;       // cluster_min_points
  int    max_layer;
  int    surf_cap;
  double bscale;     // time buckets of the final sort: bucket = (t - t_first) * bscale, clamped to [0, SORT_NB)
};
constexpr int SORT_NB = 4096;
__device__ __forceinline__ int sort_bucket(double t, double t_first, double bscale) {
  const double b = (t - t_first) * bscale;  // monotone in t: buckets never contradict the exact key order
  return b >= (double)(SORT_NB - 1) ? SORT_NB - 1 : (b > 0.0 ? (int)b : 0);
}

// moments about the voxel centre, fp64: n, s[3], ss[6]
struct Mom {
  double n, s[3], ss[6];
};
__device__ __forceinline__ void mom_zero(Mom& m) {
  m.n = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) m.s[k] = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) m.ss[k] = 0;
}
// one slot's values, gathered from the planes (six 16-byte loads)
struct SlotVals {
  long long n, st, s[3], ss[6];
  int       imin, imax;
};
__device__ __forceinline__ long long ll_of(int lo, int hi) { return (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo); }
__device__ __forceinline__ SlotVals load_slot(const wc_slot_planes& S, int s) {
  const int4 a = S.p0[2 * s], b = S.p0[2 * s + 1], c = S.p1[2 * s], d = S.p1[2 * s + 1], e = S.p2[2 * s], f = S.p2[2 * s + 1];
  SlotVals   v;
  v.n = a.x, v.imin = a.y, v.imax = a.z;
  v.st = ll_of(b.x, b.y), v.s[0] = ll_of(b.z, b.w), v.s[1] = ll_of(c.x, c.y), v.s[2] = ll_of(c.z, c.w);
  v.ss[0] = ll_of(d.x, d.y), v.ss[1] = ll_of(d.z, d.w), v.ss[2] = ll_of(e.x, e.y), v.ss[3] = ll_of(e.z, e.w);
  v.ss[4] = ll_of(f.x, f.y), v.ss[5] = ll_of(f.z, f.w);
  return v;
}
// add one slot's exact integer moments (about its leaf centre) shifted to the voxel centre
__device__ __forceinline__ void mom_add_slot(Mom& m, const SlotVals* __restrict__ sl, int leaf, double q0, double q1) {
  const int    c1 = leaf >> 3, c2 = leaf & 7;
  const double dx = ((c1 & 4) ? q0 : -q0) + ((c2 & 4) ? q1 : -q1);
  const double dy = ((c1 & 2) ? q0 : -q0) + ((c2 & 2) ? q1 : -q1);
  const double dz = ((c1 & 1) ? q0 : -q0) + ((c2 & 1) ? q1 : -q1);
  const double inv = 1.0 / WC_COORD_SCALE, inv2 = inv * inv;
  const double n  = (double)sl->n;
  const double sx = (double)sl->s[0] * inv, sy = (double)sl->s[1] * inv, sz = (double)sl->s[2] * inv;
  m.n += n;
  m.s[0] += sx + n * dx, m.s[1] += sy + n * dy, m.s[2] += sz + n * dz;
  m.ss[0] += (double)sl->ss[0] * inv2 + 2.0 * dx * sx + n * dx * dx;
  m.ss[1] += (double)sl->ss[1] * inv2 + dx * sy + dy * sx + n * dx * dy;
  m.ss[2] += (double)sl->ss[2] * inv2 + dx * sz + dz * sx + n * dx * dz;
  m.ss[3] += (double)sl->ss[3] * inv2 + 2.0 * dy * sy + n * dy * dy;
  m.ss[4] += (double)sl->ss[4] * inv2 + dy * sz + dz * sy + n * dy * dz;
  m.ss[5] += (double)sl->ss[5] * inv2 + 2.0 * dz * sz + n * dz * dz;
}
__device__ __forceinline__ void mom_add(Mom& a, const Mom& b) {
  a.n += b.n;
#pragma unroll
  for (int k = 0; k < 3; ++k) a.s[k] += b.s[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) a.ss[k] += b.ss[k];
}
// per-entry record: moments about the voxel centre + exact time sums + time extrema, 14 doubles
//   [0] n, [1..3] s, [4..9] ss, [10] n*bin, [11] sum(Q - bin<<31) (both exact integers < 2^53), [12] index of the earliest point, [13] of the latest
constexpr int REC = 14;
__device__ __forceinline__ void rec_from_slot(const wc_slot_planes& S, int slot, unsigned lb, double q0, double q1, double* r) {
  const SlotVals v = load_slot(S, slot);
  const SlotVals* sl = &v;
  Mom m;
  mom_zero(m);
  mom_add_slot(m, sl, (int)(lb >> 12), q0, q1);
  r[0] = m.n;
#pragma unroll
  for (int k = 0; k < 3; ++k) r[1 + k] = m.s[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) r[4 + k] = m.ss[k];
  r[10] = (double)(sl->n * (long long)(lb & 4095u));
  r[11] = (double)sl->st;
  r[12] = (double)sl->imin;
  r[13] = (double)sl->imax;
}
__device__ __forceinline__ void mom_add_rec(Mom& m, const double* r) {
  m.n += r[0];
#pragma unroll
  for (int k = 0; k < 3; ++k) m.s[k] += r[1 + k];
#pragma unroll
  for (int k = 0; k < 6; ++k) m.ss[k] += r[4 + k];
}

// InitPlane / ClusterSurfels statistics: mean, population covariance, ascending eigen-pairs, likeness
struct PlaneFit {
  double mu[3], cov[6], ev[3], like;
  M3     evec;
};
__device__ __forceinline__ void plane_fit(const Mom& m, PlaneFit& f) {
  const double n = m.n;
  f.mu[0] = m.s[0] / n, f.mu[1] = m.s[1] / n, f.mu[2] = m.s[2] / n;
  f.cov[0] = m.ss[0] / n - f.mu[0] * f.mu[0];
  f.cov[1] = m.ss[1] / n - f.mu[0] * f.mu[1];
  f.cov[2] = m.ss[2] / n - f.mu[0] * f.mu[2];
  f.cov[3] = m.ss[3] / n - f.mu[1] * f.mu[1];
  f.cov[4] = m.ss[4] / n - f.mu[1] * f.mu[2];
  f.cov[5] = m.ss[5] / n - f.mu[2] * f.mu[2];
  SymEig3(f.cov[0], f.cov[1], f.cov[2], f.cov[3], f.cov[4], f.cov[5], f.ev, f.evec);
  f.like = 2.0 * (f.ev[1] - f.ev[0]) / (f.ev[0] + f.ev[1] + f.ev[2]);
}

__device__ __forceinline__ unsigned level_key(unsigned leafbin, int level) {
  // leafbin = leaf<<12 | bin.  Sort keys: L2 (leaf, bin); L1 (leaf>>3, bin, leaf&7); L0 (bin, leaf)
  const unsigned leaf = leafbin >> 12, bin = leafbin & 4095u;
  if (level == 2) return leafbin;
  if (level == 1) return ((leaf >> 3) << 15) | (bin << 3) | (leaf & 7u);
  return (bin << 6) | leaf;
}
__device__ __forceinline__ unsigned group_of(unsigned lk, int level) { return level == 2 ? lk : (level == 1 ? lk >> 3 : lk >> 6); }
__device__ __forceinline__ unsigned node_of(unsigned lk, int level) { return level == 2 ? lk >> 12 : (level == 1 ? lk >> 15 : 0u); }

template <int NT>
__device__ void bitonic_sort_smem(unsigned* keys, int npad) {
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned a = keys[i], b = keys[ixj];
          const bool     up = (i & k) == 0;
          if ((a > b) == up) keys[i] = b, keys[ixj] = a;
        }
      }
      __syncthreads();
    }
}

// One CTA per voxel whose entry count E lies in (E_LO, ECAP].
// STAGE: the voxel's entry records are converted once and kept in shared memory (the per-node / per-cluster loops then
// read ~30-cycle shared memory instead of chasing 128-byte slots through L2).
#ifndef WC_EMIT_MINB
#define WC_EMIT_MINB(NT) 1
#endif
template <int ECAP, int NT, bool STAGE>
__global__ void __launch_bounds__(NT, WC_EMIT_MINB(NT))
cluster_eig_emit(const wc_slot_planes slots, const double* __restrict__ time, const int* __restrict__ seg, const int* __restrict__ vox_off,
                 const unsigned long long* __restrict__ vox_key, wc_extract_status* __restrict__ st, EmitParams P, int e_lo,
                 wc_surfel* __restrict__ out, unsigned long long* __restrict__ sort_hi,
                 unsigned long long* __restrict__ sort_lo, int* __restrict__ bcnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned*      skey    = reinterpret_cast<unsigned*>(smem_raw);      // ECAP sort words: levelkey<<14 | entry
  unsigned*      leafbin = skey + ECAP;                                // ECAP
  int*           sid     = reinterpret_cast<int*>(leafbin + ECAP);     // ECAP slot ids
  unsigned char* flag    = reinterpret_cast<unsigned char*>(sid + ECAP);  // ECAP: 1 group head, 2 node head, 4 cluster head
  unsigned short* jobs   = reinterpret_cast<unsigned short*>(flag + ECAP);  // ECAP: cluster heads that need a plane fit
  double*        rec     = reinterpret_cast<double*>(smem_raw + ECAP * 16);  // STAGE: ECAP x REC doubles
  // entry e -> its record (shared memory when staged, else built from the slot in L2)
  auto get_rec = [&](int e, double* tmp) -> const double* {
    if (STAGE) return rec + e * REC;
    rec_from_slot(slots, sid[e], leafbin[e], P.q0, P.q1, tmp);
    return tmp;
  };
  __shared__ Mom           tot[73];   // 0 root, 1..8 layer 1, 9..72 layer 2
  __shared__ unsigned char emit[73];
  __shared__ unsigned char l1_cut[8];  // layer-1 child analysed and not planar => its children exist
  __shared__ unsigned      bitmap[128];   // presence bits of the 64 x 64 (node, local bin) key space
  __shared__ unsigned      bprefix[128];
  __shared__ int           s_bmin, s_bmax, s_njobs;

  const int nv = st->n_voxels;
  for (int v = blockIdx.x; v < nv; v += gridDim.x) {
    const int off = vox_off[v];
    const int E   = vox_off[v + 1] - off;
    if (E <= e_lo || E > ECAP) {
      if (E > 8192 && threadIdx.x == 0) st->err_capacity = 1;
      continue;
    }
    __syncthreads();
    int npad = 1;
    while (npad < E) npad <<= 1;
    for (int e = threadIdx.x; e < npad; e += NT) {
      if (e < E) {
        const int                s   = seg[off + e];
        const unsigned long long key = slots.key[s];
        sid[e]                       = s;
        leafbin[e]                   = (unsigned)(key & 0x3ffffu);
        skey[e]                      = ((unsigned)(key & 0x3ffffu) << 14) | (unsigned)e;
        if (STAGE) rec_from_slot(slots, s, (unsigned)(key & 0x3ffffu), P.q0, P.q1, rec + e * REC);
      } else {
        skey[e] = 0xffffffffu;
      }
    }
    for (int k = threadIdx.x; k < 73; k += NT) mom_zero(tot[k]), emit[k] = 0;
    if (threadIdx.x == 0) s_bmin = 4096, s_bmax = -1;
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += NT) {
      const int b = (int)(leafbin[e] & 4095u);
      atomicMin(&s_bmin, b);
      atomicMax(&s_bmax, b);
    }
    __syncthreads();
    const int  bmin = s_bmin, nbl = s_bmax - s_bmin + 1;
    const bool rank_sort = nbl <= 64;  // (leaf, local bin) fits a 4096-bit map: sort by ranking unique keys, no comparisons
    // voxel centre (surfel_extraction.cc:209-211)
    const unsigned long long vk = vox_key[v];
    const int vx = (int)((vk >> 30) & 32767u) - WC_VOX_BIAS + P.vox0[0], vy = (int)((vk >> 15) & 32767u) - WC_VOX_BIAS + P.vox0[1],
              vz = (int)(vk & 32767u) - WC_VOX_BIAS + P.vox0[2];
    const double ccx = (0.5 + (double)vx) * P.voxel, ccy = (0.5 + (double)vy) * P.voxel, ccz = (0.5 + (double)vz) * P.voxel;

    for (int level = 2; level >= 0; --level) {
      if (level != 2) {
        for (int e = threadIdx.x; e < npad; e += NT)
          skey[e] = e < E ? ((level_key(leafbin[e], level) << 14) | (unsigned)e) : 0xffffffffu;
        __syncthreads();
      }
      if (rank_sort) {
        // keys are unique per entry, so the sorted position of an entry is the number of present keys below its own:
        // set a presence bit per entry in the compact key space, prefix-popcount the 128 words, rank = prefix + popc.
        auto compact = [&](unsigned lb) -> unsigned {
          const unsigned leaf = lb >> 12, bl = (lb & 4095u) - (unsigned)bmin;
          if (level == 2) return leaf * (unsigned)nbl + bl;
          if (level == 1) return ((leaf >> 3) * (unsigned)nbl + bl) * 8u + (leaf & 7u);
          return bl * 64u + leaf;
        };
        for (int w = threadIdx.x; w < 128; w += NT) bitmap[w] = 0u;
        __syncthreads();
        for (int e = threadIdx.x; e < E; e += NT) {
          const unsigned ck = compact(leafbin[e]);
          atomicOr(&bitmap[ck >> 5], 1u << (ck & 31u));
        }
        __syncthreads();
        if (threadIdx.x < 32) {  // exclusive prefix of the word popcounts, 4 words per lane
          unsigned c4[4], sum = 0;
#pragma unroll
          for (int u = 0; u < 4; ++u) c4[u] = __popc(bitmap[threadIdx.x * 4 + u]), sum += c4[u];
          unsigned incl = sum;
          for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)threadIdx.x >= d) incl += o;
          }
          unsigned run = incl - sum;
#pragma unroll
          for (int u = 0; u < 4; ++u) bprefix[threadIdx.x * 4 + u] = run, run += c4[u];
        }
        __syncthreads();
        unsigned mykey[(ECAP + NT - 1) / NT];
        int      cnt = 0;
        for (int e = threadIdx.x; e < E; e += NT) mykey[cnt++] = skey[e];
        __syncthreads();
        cnt = 0;
        for (int e = threadIdx.x; e < E; e += NT) {
          const unsigned ck   = compact(leafbin[e]);
          const unsigned rank = bprefix[ck >> 5] + __popc(bitmap[ck >> 5] & ((1u << (ck & 31u)) - 1u));
          skey[rank]          = mykey[cnt++];
        }
        __syncthreads();
      } else {
        bitonic_sort_smem<NT>(skey, npad);
      }
      // head flags
      for (int i = threadIdx.x; i < E; i += NT) {
        const unsigned lk = skey[i] >> 14;
        unsigned char  f  = 0;
        if (i == 0) f = 3;
        else {
          const unsigned pk = skey[i - 1] >> 14;
          if (group_of(lk, level) != group_of(pk, level)) f |= 1;
          if (node_of(lk, level) != node_of(pk, level)) f |= 3;
        }
        flag[i] = f;
      }
      __syncthreads();
      if (level == 2) {
        // leaf totals (thread per leaf run), then the tree 64 -> 8 -> 1 and the planarity flags
        for (int i = threadIdx.x; i < E; i += NT) {
          if (!(flag[i] & 2)) continue;
          const int leaf = (int)(skey[i] >> 26);
          Mom       m;
          mom_zero(m);
          double tmp[REC];
          for (int j = i; j < E && (j == i || !(flag[j] & 2)); ++j) mom_add_rec(m, get_rec(skey[j] & 16383u, tmp));
          tot[9 + leaf] = m;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
          Mom m;
          mom_zero(m);
          for (int c = 0; c < 8; ++c) mom_add(m, tot[9 + 8 * threadIdx.x + c]);
          tot[1 + threadIdx.x] = m;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          Mom m;
          mom_zero(m);
          for (int c = 0; c < 8; ++c) mom_add(m, tot[1 + c]);
          tot[0] = m;
        }
        __syncthreads();
        // InitOctoTree / CutOctoTree decisions (surfel_extraction.cc:128-184)
        const bool root_analysed = tot[0].n > (double)P.lps[0];
        for (int k = threadIdx.x; k < 9; k += NT) {
          const int layer    = k == 0 ? 0 : 1;
          bool      analysed = k == 0 ? root_analysed : (root_analysed && P.max_layer >= 1 && tot[k].n > (double)P.lps[1]);
          bool      plane    = false;
          if (analysed) {
            PlaneFit f;
            plane_fit(tot[k], f);
            plane = f.ev[0] < P.thr && f.like > P.min_like;
          }
          emit[k] = analysed && plane;
          if (layer == 1) l1_cut[k - 1] = analysed && !plane && P.max_layer >= 2;
        }
        __syncthreads();
        for (int k = 9 + threadIdx.x; k < 73; k += NT) {
          const bool analysed = l1_cut[(k - 9) >> 3] && tot[k].n > (double)P.lps[2];
          bool       plane    = false;
          if (analysed) {
            PlaneFit f;
            plane_fit(tot[k], f);
            plane = f.ev[0] < P.thr && f.like > P.min_like;
          }
          emit[k] = analysed && plane;
        }
        __syncthreads();
      }
      const int node_base = level == 2 ? 9 : (level == 1 ? 1 : 0);
      // cluster boundaries: at each group head compare with the previous group of the same node
      unsigned long long my_starts = 0;  // bit per loop trip: written to flag[] only after every thread has read it
      int                trip      = 0;
      for (int i = threadIdx.x; i < E; i += NT, ++trip) {
        const unsigned char f = flag[i];
        if (!(f & 1)) continue;
        const unsigned lk = skey[i] >> 14;
        if (!emit[node_base + node_of(lk, level)]) continue;
        bool start = (f & 2) != 0;
        if (!start) {
          int gmin = INT_MAX, pmax = -1;  // earliest point of this group, latest point of the previous one
          for (int j = i; j < E && (j == i || !(flag[j] & 1)); ++j) {
            const int e = skey[j] & 16383u;
            gmin        = min(gmin, STAGE ? (int)rec[e * REC + 12] : slots.p0[2 * sid[e]].y);
          }
          for (int j = i - 1; j >= 0; --j) {
            const int e = skey[j] & 16383u;
            pmax        = max(pmax, STAGE ? (int)rec[e * REC + 13] : slots.p0[2 * sid[e]].z);
            if (flag[j] & 1) break;
          }
          // points[i].timestamp - cluster.back().timestamp > 0.05  (surfel_extraction.cc:24)
          start = __dsub_rn(time[gmin], time[pmax]) > P.gap;
        }
        if (start) my_starts |= 1ull << trip;
      }
      __syncthreads();
      trip = 0;
      for (int i = threadIdx.x; i < E; i += NT, ++trip)
        if ((my_starts >> trip) & 1ull) flag[i] |= 4;
      __syncthreads();
      // One thread per cluster counts its points; the clusters that reach cluster_min_points (surfel_extraction.cc:33) become
      // fit jobs.  The jobs are then taken by consecutive threads, so the eigen-solves of a voxel run on one or two full
      // warps instead of a few lanes of every warp.
      if (threadIdx.x == 0) s_njobs = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < E; i += NT) {
        if (!(flag[i] & 4)) continue;
        double n = 0.0;
        for (int j = i; j < E && (j == i || !(flag[j] & 6)); ++j) {
          const int e = skey[j] & 16383u;
          n += STAGE ? rec[e * REC] : (double)slots.p0[2 * sid[e]].x;
        }
        if (n < (double)P.cmin) continue;
        jobs[atomicAdd(&s_njobs, 1)] = (unsigned short)i;
      }
      __syncthreads();
      const int njobs = s_njobs;
      for (int q = threadIdx.x; q < njobs; q += NT) {
        const int      i   = jobs[q];
        const unsigned lk0 = skey[i] >> 14;
        Mom            m;
        mom_zero(m);
        double tA = 0.0, tB = 0.0, tmp[REC];  // exact integer sums (< 2^53)
        for (int j = i; j < E && (j == i || !(flag[j] & 6)); ++j) {
          const double* r = get_rec(skey[j] & 16383u, tmp);
          mom_add_rec(m, r);
          tA += r[10];
          tB += r[11];
        }
        PlaneFit f;
        plane_fit(m, f);
        const bool accepted = !(f.ev[0] > P.thr || f.like < P.min_like);  // :54
        // output slots: one counter bump per warp (the lanes of this trip that accepted their cluster)
        const unsigned act = __activemask(), acc = __ballot_sync(act, accepted);
        if (!accepted) continue;
        const int leader = __ffs(acc) - 1, lane_id = threadIdx.x & 31;
        int       idx    = 0;
        if (lane_id == leader) idx = atomicAdd(&st->n_surfels, __popc(acc));
        idx = __shfl_sync(acc, idx, leader) + __popc(acc & ((1u << lane_id) - 1u));
        const double cxw = ccx + f.mu[0], cyw = ccy + f.mu[1], czw = ccz + f.mu[2];
        V3           nrm = col(f.evec, 0);
        if (nrm.x * (cxw - P.view[0]) + nrm.y * (cyw - P.view[1]) + nrm.z * (czw - P.view[2]) < 0) nrm = -nrm;
        if (idx >= P.surf_cap) {
          st->err_capacity = 1;
          continue;
        }
        const double tmean = P.t_first + (tA * 2147483648.0 + tB) / (m.n * WC_TIME_SCALE);
        wc_surfel    s;
        s.timestamp           = tmean;
        s.resolution          = level == 0 ? P.voxel : (level == 1 ? 2.0 * P.q0 : 2.0 * P.q1);  // (float)(quarter*4)
        s.plane_std_deviation = sqrt(f.ev[0]);
        s.rot[0] = s.rot[1] = s.rot[2] = 0.0, s.rot[3] = 1.0;
        s.pos[0] = s.pos[1] = s.pos[2] = 0.0;
        s.center[0] = cxw, s.center[1] = cyw, s.center[2] = czw;
        s.covariance[0] = f.cov[0], s.covariance[1] = f.cov[1], s.covariance[2] = f.cov[2];
        s.covariance[3] = f.cov[1], s.covariance[4] = f.cov[3], s.covariance[5] = f.cov[4];
        s.covariance[6] = f.cov[2], s.covariance[7] = f.cov[4], s.covariance[8] = f.cov[5];
        s.norm[0] = nrm.x, s.norm[1] = nrm.y, s.norm[2] = nrm.z;
        s.is_in_body_frame = 0, s._pad = 0;
        out[idx]     = s;
        sort_hi[idx] = OrderedBits(tmean);
        sort_lo[idx] = ((unsigned long long)level << 62) | (OrderedBits(cxw) >> 2);
        atomicAdd(&bcnt[sort_bucket(tmean, P.t_first, P.bscale)], 1);
        (void)lk0;
      }
      __syncthreads();
    }
  }
}

// leave the cell table, the voxel table and the counters clean for the next call: only the touched entries are reset
// (the tables are initialised once, at allocation; the slot records are fully rewritten by their creators)
__global__ void extract_cleanup(const wc_extract_status* __restrict__ st, int slot_cap, int vox_cap, HEnt* __restrict__ tab,
                                const int* __restrict__ rec_tpos, unsigned long long* __restrict__ vkeys, int* __restrict__ vslot,
                                const int* __restrict__ vox_hpos, int* __restrict__ vox_count) {
  const int nv  = min(st->n_voxels, vox_cap), ns = min(st->n_slots, slot_cap);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int s = tid; s < ns; s += nth) {
    const int h = rec_tpos[s];
    if (h >= 0) *reinterpret_cast<int4*>(tab + h) = make_int4(-1, -1, -1, 0);  // {key = WC_KEY_EMPTY, slot = -1}
  }
  for (int v = tid; v < nv; v += nth) {
    const int h  = vox_hpos[v];
    vkeys[h]     = WC_KEY_EMPTY;
    vslot[h]     = -1;
    vox_count[v] = 0;
  }
}

// ---------------------------------------------------------------------------------------------- final sort
// std::sort(surfels by timestamp) (surfel_extraction.cc:334) as a bucket sort: the emit kernels count the surfels per
// time bucket (SORT_NB uniform buckets over the sweep; the bucket map is monotone in the timestamp), one CTA scans the
// counts, the surfel ids are scattered into their buckets, and one warp per bucket orders its handful of entries by the
// exact key (timestamp bits, then level / centre.x bits, then emission id) by rank counting.  Four launches instead of
// a 20-launch global bitonic network.
struct SortRec {
  unsigned long long hi, lo;
  unsigned           idx;
};
__device__ __forceinline__ bool rec_less(const SortRec& a, const SortRec& b) {
  if (a.hi != b.hi) return a.hi < b.hi;
  if (a.lo != b.lo) return a.lo < b.lo;
  return a.idx < b.idx;
}

__global__ void __launch_bounds__(1024) bsort_scan(const int* __restrict__ bcnt, int* __restrict__ boff, int* __restrict__ bcur) {
  __shared__ int warp_sums[32];
  const int t = threadIdx.x, lane = t & 31;
  int       v[SORT_NB / 1024], tot = 0;
#pragma unroll
  for (int k = 0; k < SORT_NB / 1024; ++k) v[k] = bcnt[t * (SORT_NB / 1024) + k], tot += v[k];
  int incl = tot;
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) warp_sums[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    const int w  = warp_sums[t];
    int       wi = w;
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, wi, d);
      if (t >= d) wi += o;
    }
    warp_sums[t] = wi - w;
  }
  __syncthreads();
  int run = warp_sums[t >> 5] + incl - tot;
#pragma unroll
  for (int k = 0; k < SORT_NB / 1024; ++k) {
    boff[t * (SORT_NB / 1024) + k] = run, bcur[t * (SORT_NB / 1024) + k] = 0;
    run += v[k];
  }
  if (t == 1023) boff[SORT_NB] = run;
}

// (the surfel count is read on the device: the host does not wait for the emit kernels before enqueueing the sort)
__global__ void bsort_scatter(const unsigned long long* __restrict__ hi, const wc_extract_status* __restrict__ st, int cap, double t_first,
                              double bscale, const int* __restrict__ boff, int* __restrict__ bcur, unsigned* __restrict__ perm) {
  const int n = min(st->n_surfels, cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int b = sort_bucket(FromOrderedBits(hi[i]), t_first, bscale);
    perm[boff[b] + atomicAdd(&bcur[b], 1)] = (unsigned)i;
  }
}

// one warp per bucket: out_idx[boff[b] + rank] = id, rank = number of bucket entries with a smaller exact key
__global__ void __launch_bounds__(256)
bsort_rank(const unsigned long long* __restrict__ hi, const unsigned long long* __restrict__ lo, const int* __restrict__ boff,
           const unsigned* __restrict__ perm, int* __restrict__ bcnt, unsigned* __restrict__ out_idx) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= SORT_NB) return;
  const int p0 = boff[b], m = boff[b + 1] - p0;
  if (lane == 0) bcnt[b] = 0;  // leave the counters clean for the next call
  for (int base = 0; base < m; base += 32) {
    const int e = base + lane;
    SortRec   me{~0ull, ~0ull, ~0u};
    if (e < m) me.idx = perm[p0 + e], me.hi = hi[me.idx], me.lo = lo[me.idx];
    int rank = 0;
    for (int cb = 0; cb < m; cb += 32) {  // all entries of the bucket, 32 at a time through the lanes
      const int c = cb + lane;
      SortRec   ot{~0ull, ~0ull, ~0u};
      if (c < m) ot.idx = perm[p0 + c], ot.hi = hi[ot.idx], ot.lo = lo[ot.idx];
      const int lim = min(32, m - cb);
      for (int j = 0; j < lim; ++j) {
        SortRec o;
        o.hi  = __shfl_sync(0xffffffffu, ot.hi, j);
        o.lo  = __shfl_sync(0xffffffffu, ot.lo, j);
        o.idx = __shfl_sync(0xffffffffu, ot.idx, j);
        rank += rec_less(o, me);
      }
    }
    if (e < m) out_idx[p0 + rank] = me.idx;
  }
}

__global__ void gather_surfels(const wc_surfel* __restrict__ in, const unsigned* __restrict__ idx, const wc_extract_status* __restrict__ st,
                               int cap, wc_surfel* __restrict__ out) {
  // 13 x 16-byte chunks per surfel; consecutive threads move consecutive chunks
  const long long n13 = 13ll * min(st->n_surfels, cap);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n13; t += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(t / 13), c = (int)(t % 13);
    reinterpret_cast<uint4*>(out + s)[c] = reinterpret_cast<const uint4*>(in + idx[s])[c];
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host
static wc_status extract_alloc(wc_ctx* c) {
  if (c->d_xyz) return WC_OK;
  const size_t np = (size_t)c->prm.max_points;
  c->hcap         = wc_next_pow2(np < 1024 ? 1024 : np);
  c->slot_cap     = np;
  c->vcap         = wc_next_pow2(2 * np);
  WC_CUDA(c, cudaMalloc(&c->d_raw, np * sizeof(wc_point48)));
  WC_CUDA(c, cudaMalloc(&c->d_xyz, np * sizeof(float4)));
  WC_CUDA(c, cudaMalloc(&c->d_time, np * sizeof(double)));
  WC_CUDA(c, cudaMalloc(&c->d_htab, c->hcap * sizeof(HEnt)));
  c->slot_cap     = (np + 3) & ~(size_t)3;  // keeps every plane 32-byte aligned
  WC_CUDA(c, cudaMalloc(&c->d_slots, c->slot_cap * WC_SLOT_BYTES));
  WC_CUDA(c, cudaMalloc(&c->d_vkeys, c->vcap * 8));
  WC_CUDA(c, cudaMalloc(&c->d_vslot, c->vcap * 4));
  WC_CUDA(c, cudaMalloc(&c->d_vox_count, (np + 1) * 4));
  WC_CUDA(c, cudaMalloc(&c->d_vox_off, (np + 1) * 4));
  WC_CUDA(c, cudaMalloc(&c->d_rec_info, np * sizeof(int2)));
  WC_CUDA(c, cudaMalloc(&c->d_rec_tpos, np * 4));
  WC_CUDA(c, cudaMalloc(&c->d_vox_key, (np + 1) * 8));
  WC_CUDA(c, cudaMalloc(&c->d_vox_hpos, (np + 1) * 4));
  WC_CUDA(c, cudaMalloc(&c->d_seg, np * 4));
  WC_CUDA(c, cudaMalloc(&c->d_xstat, sizeof(wc_extract_status)));
  WC_CUDA(c, cudaMallocHost(&c->h_xstat, sizeof(wc_extract_status)));
  const size_t sc = wc_next_pow2((size_t)c->prm.max_surfels < 2048 ? 2048 : (size_t)c->prm.max_surfels);
  WC_CUDA(c, cudaMalloc(&c->d_surf_raw, sc * sizeof(wc_surfel)));
  WC_CUDA(c, cudaMalloc(&c->d_surf, sc * sizeof(wc_surfel)));
  WC_CUDA(c, cudaMalloc(&c->d_sort_hi, sc * 8));
  WC_CUDA(c, cudaMalloc(&c->d_sort_lo, sc * 8));
  WC_CUDA(c, cudaMalloc(&c->d_sort_idx, sc * 4));
  WC_CUDA(c, cudaMalloc(&c->d_sort_perm, sc * 4));
  WC_CUDA(c, cudaMalloc(&c->d_bcnt, SORT_NB * 4));
  WC_CUDA(c, cudaMalloc(&c->d_boff, (SORT_NB + 1) * 4));
  WC_CUDA(c, cudaMalloc(&c->d_bcur, SORT_NB * 4));
  WC_CUDA(c, cudaMemsetAsync(c->d_bcnt, 0, SORT_NB * 4, c->stream));
  WC_CUDA(c, cudaMalloc(&c->d_assign, np * sizeof(wc_point_assign)));
  WC_CUDA(c, cudaMemsetAsync(c->d_vkeys, 0xff, c->vcap * 8, c->stream));
  WC_CUDA(c, cudaMemsetAsync(c->d_vslot, 0xff, c->vcap * 4, c->stream));
  WC_CUDA(c, cudaMemsetAsync(c->d_vox_count, 0, (np + 1) * 4, c->stream));
  WC_CUDA(c, cudaMemsetAsync(c->d_htab, 0xff, c->hcap * sizeof(HEnt), c->stream));
  WC_CUDA(c, cudaFuncSetAttribute(voxel_key_moments, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM));
  WC_CUDA(c, cudaFuncSetAttribute((cluster_eig_emit<512, 128, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * (16 + 8 * REC)));
  WC_CUDA(c, cudaFuncSetAttribute((cluster_eig_emit<8192, 256, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
  return WC_OK;
}

void wc_extract_free(wc_ctx* c) {
  void* ptrs[] = {c->d_raw,  c->d_raw_next,   c->d_xyz,      c->d_time,    c->d_htab,   c->d_slots,    c->d_vkeys,
                  c->d_vslot,   c->d_vox_count, c->d_vox_off, c->d_rec_info, c->d_rec_tpos, c->d_vox_key, c->d_vox_hpos, c->d_seg,   c->d_xstat,
                  c->d_surf_raw, c->d_surf,    c->d_sort_hi, c->d_sort_lo, c->d_sort_idx, c->d_assign, c->d_sort_perm, c->d_bcnt, c->d_boff, c->d_bcur};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (c->h_xstat) cudaFreeHost(c->h_xstat);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream), cudaStreamDestroy(c->copy_stream);
  if (c->ev_prefetch) cudaEventDestroy(c->ev_prefetch);
}

void*     wc_comm_raw_area(wc_ctx* c, unsigned long long epoch);  // wc_comm.cu
void      wc_comm_raw_areas(wc_ctx* c, unsigned long long epoch, const void* areas[8]);
wc_status wc_comm_raw_ready(wc_ctx* c, unsigned long long epoch);
wc_status wc_comm_check(wc_ctx* c);

// sharded sweep upload: world > 1 and switched on by wc_comm_shard_upload (the uploads are then collective calls)
static bool shard_upload(const wc_ctx* c) { return c->shard_upload && c->world > 1 && c->comm_ready; }
static void slab_of(const wc_ctx* c, size_t n, int r, size_t* lo, size_t* hi) {
  *lo = n * (size_t)r / (size_t)c->world, *hi = n * (size_t)(r + 1) / (size_t)c->world;
}

static wc_slot_planes slot_planes(const wc_ctx* c) {
  unsigned char* b = (unsigned char*)c->d_slots;
  const size_t   n = c->slot_cap;
  return wc_slot_planes{(unsigned long long*)b, (int4*)(b + 8 * n), (int4*)(b + 40 * n), (int4*)(b + 72 * n)};
}

// The background copy goes out in pieces: the pass that runs meanwhile reads a few small results back between its stages,
// and a copy engine works through its queue in order — behind one 96 MB transfer such a read-back would wait up to 2 ms.
static wc_status copy_chunked(wc_ctx* c, wc_point48* dst, const wc_point48* src, size_t n) {
  static const long long chunk_mb = getenv("WC_PREFETCH_CHUNK_MB") ? atoll(getenv("WC_PREFETCH_CHUNK_MB")) : 1;
  const size_t per = chunk_mb > 0 ? (size_t)chunk_mb * (1u << 20) / sizeof(wc_point48) : n;
  for (size_t i = 0; i < n; i += per) {
    const size_t m = n - i < per ? n - i : per;
    WC_CUDA(c, cudaMemcpyAsync(dst + i, src + i, m * sizeof(wc_point48), cudaMemcpyHostToDevice, c->copy_stream));
  }
  return WC_OK;
}

// Streaming ingestion: copy the NEXT sweep into a second staging buffer on a dedicated copy stream while the current window
// pass runs; the wc_points_upload of that same buffer then finds the data already on the device (it only waits for the
// copy's event).  pts must stay valid and unchanged until that upload, and should be pinned memory (wc_host_alloc) for the
// copy to be truly asynchronous.  Only the NEXT wc_points_upload can claim a prefetch; one that is not claimed is dropped.
// when = WC_PREFETCH_NOW: the copy starts at once.  when = WC_PREFETCH_AT_SOLVE: it starts when the next
// wc_window_pass_resident reaches its solve stage — extraction and matching are memory bound and lose 0.45 ms at C3 to a
// transfer running beside them, the latency-bound solve (2.2 ms >= the 1.9 ms copy) loses nothing; without a pass before
// the upload, the upload simply copies the sweep itself.
static wc_status prefetch_issue(wc_ctx* c, const wc_point48* pts, size_t n) {
  wc_status s = WC_OK;
  if (!c->copy_stream) {
    WC_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    WC_CUDA(c, cudaEventCreateWithFlags(&c->ev_prefetch, cudaEventDisableTiming));
  }
  if (!shard_upload(c) && !c->d_raw_next) WC_CUDA(c, cudaMalloc(&c->d_raw_next, (size_t)c->prm.max_points * sizeof(wc_point48)));
  // (no dependency on the main stream: the staging buffer written here was last read by the repack of an upload that
  //  synchronised the stream before it returned; the exported raw areas are protected by the flag protocol)
  if (shard_upload(c)) {
    // only this rank's slab crosses PCIe, into this rank's exported raw area of the next epoch (every peer has finished
    // reading that area: it signalled the epoch in between after its repack)
    size_t lo, hi;
    slab_of(c, n, c->rank, &lo, &hi);
    c->prefetch_epoch = ++c->raw_epoch;
    wc_point48* area  = (wc_point48*)wc_comm_raw_area(c, c->prefetch_epoch);
    if ((s = copy_chunked(c, area + lo, pts + lo, hi - lo))) return s;
  } else {
    if ((s = copy_chunked(c, (wc_point48*)c->d_raw_next, pts, n))) return s;
  }
  WC_CUDA(c, cudaEventRecord(c->ev_prefetch, c->copy_stream));
  c->prefetch_src = pts, c->prefetch_n = n;
  return WC_OK;
}

extern "C" wc_status wc_points_prefetch(wc_ctx* c, const wc_point48* pts, size_t n, int when) {
  if (!c || (!pts && n) || (when != WC_PREFETCH_NOW && when != WC_PREFETCH_AT_SOLVE)) return WC_EINVAL;
  if (n > (size_t)c->prm.max_points) WC_FAIL(c, WC_ECAPACITY, "n=%zu exceeds max_points=%lld", n, (long long)c->prm.max_points);
  wc_status s = extract_alloc(c);
  if (s) return s;
  c->prefetch_src = nullptr, c->defer_src = nullptr;
  if (n == 0) return WC_OK;
  if (when == WC_PREFETCH_AT_SOLVE) {
    c->defer_src = pts, c->defer_n = n;
    return WC_OK;
  }
  return prefetch_issue(c, pts, n);
}

// called by the window pass when its solve stage is about to be enqueued
wc_status wc_points_prefetch_deferred(wc_ctx* c) {
  if (!c->defer_src) return WC_OK;
  const wc_point48* pts = (const wc_point48*)c->defer_src;
  c->defer_src          = nullptr;
  return prefetch_issue(c, pts, c->defer_n);
}

extern "C" wc_status wc_points_upload(wc_ctx* c, const wc_point48* pts, size_t n) {
  if (!c || (!pts && n)) return WC_EINVAL;
  if (n > (size_t)c->prm.max_points) WC_FAIL(c, WC_ECAPACITY, "n=%zu exceeds max_points=%lld", n, (long long)c->prm.max_points);
  wc_status s = extract_alloc(c);
  if (s) return s;
  c->n_pts = n;
  if (n == 0) return WC_OK;
  // a prefetch can only be claimed by the upload that directly follows it (same buffer, same size): any other upload
  // drops it, so a stale copy can never be mistaken for a later buffer that happens to live at the same address
  const bool claimed = c->prefetch_src == pts && c->prefetch_n == n;
  c->prefetch_src    = nullptr;
  c->defer_src       = nullptr;  // a prefetch still waiting for a pass's solve stage: this upload copies the sweep itself
  if (shard_upload(c)) {
    // Collective over the ranks (every rank uploads the same sweep): this rank copies its slab into its exported raw area
    // (or finds it there, prefetched), one flag round, then every point is repacked straight from its owner's area.
    unsigned long long epoch;
    if (claimed) {
      epoch = c->prefetch_epoch;
      WC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_prefetch, 0));
    } else {
      epoch = ++c->raw_epoch;
      size_t lo, hi;
      slab_of(c, n, c->rank, &lo, &hi);
      wc_point48* area = (wc_point48*)wc_comm_raw_area(c, epoch);
      if (hi > lo) WC_CUDA(c, cudaMemcpyAsync(area + lo, pts + lo, (hi - lo) * sizeof(wc_point48), cudaMemcpyHostToDevice, c->stream));
    }
    if ((s = wc_comm_raw_ready(c, epoch))) return s;
    RawAreas A;
    memset(&A, 0, sizeof(A));
    const void* areas[8];
    wc_comm_raw_areas(c, epoch, areas);
    for (int r = 0; r < 8; ++r) A.area[r] = (const wc_point48*)areas[r];
    for (int r = 0; r <= c->world; ++r) A.row0[r] = (int)(n * (size_t)r / (size_t)c->world);
    A.world = c->world;
    { ++c->n_launches; repack_points_sharded<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(A, (int)n, c->d_xyz, c->d_time); }
    WC_CUDA(c, cudaGetLastError());
  } else {
    if (claimed) {
      // this sweep was prefetched: swap the staging buffers and wait (on the device) for the copy
      void* tmp = c->d_raw; c->d_raw = c->d_raw_next; c->d_raw_next = tmp;
      WC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_prefetch, 0));
    } else {
      WC_CUDA(c, cudaMemcpyAsync(c->d_raw, pts, n * sizeof(wc_point48), cudaMemcpyHostToDevice, c->stream));
    }
    { ++c->n_launches; repack_points<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const wc_point48*)c->d_raw, (int)n, c->d_xyz, c->d_time); }
    WC_CUDA(c, cudaGetLastError());
  }
  // voxel of the first point / first timestamp anchor the relative keys (host copy of element 0 is at hand)
  const double vs = (double)c->prm.voxel_size;
  c->vox0[0] = (int)floor((double)pts[0].x / vs), c->vox0[1] = (int)floor((double)pts[0].y / vs), c->vox0[2] = (int)floor((double)pts[0].z / vs);
  c->t_first = pts[0].time, c->t_last = pts[n - 1].time;
  WC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (shard_upload(c)) return wc_comm_check(c);
  return WC_OK;
}

extern "C" wc_status wc_build_surfels_resident(wc_ctx* c, size_t* n_out, double* gpu_ms_keys, double* gpu_ms_emit,
                                               double* gpu_ms_total) {
  if (!c) return WC_EINVAL;
  wc_status s = extract_alloc(c);
  if (s) return s;
  if (c->prm.max_layer < 0 || c->prm.max_layer > 2) WC_FAIL(c, WC_EINVAL, "max_layer must be 0..2");
  if (!(c->prm.cluster_time_gap > 0.03125)) WC_FAIL(c, WC_EINVAL, "cluster_time_gap must exceed the 2^-5 s time bin");
  const int n  = (int)c->n_pts;
  c->n_surfels = 0;
  if (n_out) *n_out = 0;
  if (n == 0) return WC_OK;
  cudaStream_t st = c->stream;
  const float  vsf = c->prm.voxel_size;
  ExtractParams P;
  P.v4 = vsf / 4, P.v8 = vsf / 8, P.inv_v4 = 1.0f / P.v4;  // divisions by powers of two: exact
  if (!(vsf > 0.f) || !isfinite(vsf) || (double)P.v8 * WC_COORD_SCALE >= 1073741824.0 || (double)P.v4 * 4 != (double)vsf)
    WC_FAIL(c, WC_EINVAL, "voxel_size out of range");
  P.t_first = c->t_first, P.n = n;
  for (int k = 0; k < 3; ++k) P.vox0[k] = c->vox0[k], P.goffm[k] = 4 * (WC_VOX_BIAS - c->vox0[k]) - 0x4B400000;

  WC_CUDA(c, cudaEventRecord(c->ev[0], st));
  WC_CUDA(c, cudaMemsetAsync(c->d_xstat, 0, sizeof(wc_extract_status), st));
  // the emit kernels add to the time-bucket counters before any error is known: clear them on every call, so that a
  // call that returned early (time order / range / capacity) cannot leave stale counts behind
  WC_CUDA(c, cudaMemsetAsync(c->d_bcnt, 0, SORT_NB * 4, st));
  // cell table sized to this sweep (worst case one cell per point still fits; typical load is ~0.2)
  const size_t hcap = wc_next_pow2(n < 1024 ? 1024 : (size_t)n);
  WC_CUDA(c, cudaEventRecord(c->ev[4], st));
  const int ntiles = (n + KT - 1) / KT;
  const wc_slot_planes SP = slot_planes(c);
  { ++c->n_launches; voxel_key_moments<<<ntiles, KNT, K1_SMEM, st>>>(c->d_xyz, c->d_time, P, SP, (int)c->slot_cap, c->d_xstat,
                                           c->want_assign ? c->d_assign : nullptr); }
  WC_CUDA(c, cudaEventRecord(c->ev[1], st));
  { ++c->n_launches; voxel_index<<<c->num_sms * 8, 256, 0, st>>>(SP, c->d_xstat, (HEnt*)c->d_htab, hcap - 1, c->d_vkeys, c->d_vslot, c->vcap - 1, c->d_vox_count,
                                              c->d_vox_key, c->d_vox_hpos, (int)c->prm.max_points, c->d_rec_info, c->d_rec_tpos); }
  { ++c->n_launches; voxel_scan<<<1, 1024, 0, st>>>(c->d_vox_count, c->d_vox_off, c->d_xstat); }
  { ++c->n_launches; voxel_scatter<<<c->num_sms * 4, 256, 0, st>>>(c->d_rec_info, c->d_xstat, (int)c->slot_cap, c->d_vox_off, c->d_seg); }
  EmitParams E;
  E.voxel = (double)vsf, E.q0 = (double)P.v4, E.q1 = (double)P.v8, E.t_first = P.t_first;
  E.thr = (double)c->prm.planer_threshold, E.min_like = c->prm.min_plane_likeness, E.gap = c->prm.cluster_time_gap;
  for (int k = 0; k < 3; ++k) E.view[k] = c->prm.view_point[k], E.vox0[k] = c->vox0[k], E.lps[k] = c->prm.layer_point_size[k];
  E.cmin = c->prm.cluster_min_points, E.max_layer = c->prm.max_layer;
  E.surf_cap = (int)c->prm.max_surfels;
  E.bscale   = c->t_last > c->t_first ? (double)SORT_NB / (c->t_last - c->t_first) : 0.0;
  // the three entry-count tiers are independent (surfel slots are claimed by an atomic counter): they run concurrently
  // on the main stream and two side streams, so the small tiers fill the SMs the large tier's tail leaves idle
  c->n_launches += 3;
  WC_CUDA(c, cudaEventRecord(c->ev_fork, st));
  for (int i = 0; i < 3; ++i) WC_CUDA(c, cudaStreamWaitEvent(c->side[i], c->ev_fork, 0));
  c->n_launches += 1;
  cluster_eig_emit<512, 128, true><<<c->num_sms * 3, 128, 512 * (16 + 8 * REC), st>>>(SP, c->d_time, c->d_seg, c->d_vox_off, c->d_vox_key, c->d_xstat,
                                                                                        E, 256, c->d_surf_raw, c->d_sort_hi, c->d_sort_lo, c->d_bcnt);
  cluster_eig_emit<256, 128, true><<<c->num_sms * 6, 128, 256 * (16 + 8 * REC), c->side[2]>>>(SP, c->d_time, c->d_seg, c->d_vox_off, c->d_vox_key,
                                                                                                c->d_xstat, E, 128, c->d_surf_raw, c->d_sort_hi, c->d_sort_lo, c->d_bcnt);
  cluster_eig_emit<128, 64, true><<<c->num_sms * 12, 64, 128 * (16 + 8 * REC), c->side[0]>>>(SP, c->d_time, c->d_seg, c->d_vox_off, c->d_vox_key,
                                                                                               c->d_xstat, E, 0, c->d_surf_raw, c->d_sort_hi, c->d_sort_lo, c->d_bcnt);
  cluster_eig_emit<8192, 256, false><<<c->num_sms, 256, 8192 * 16, c->side[1]>>>(SP, c->d_time, c->d_seg, c->d_vox_off, c->d_vox_key, c->d_xstat, E,
                                                                                 512, c->d_surf_raw, c->d_sort_hi, c->d_sort_lo, c->d_bcnt);
  for (int i = 0; i < 3; ++i) {
    WC_CUDA(c, cudaEventRecord(c->ev_join[i], c->side[i]));
    WC_CUDA(c, cudaStreamWaitEvent(st, c->ev_join[i], 0));
  }
  WC_CUDA(c, cudaMemcpyAsync(c->h_xstat, c->d_xstat, sizeof(wc_extract_status), cudaMemcpyDeviceToHost, st));
  { ++c->n_launches; extract_cleanup<<<c->num_sms * 4, 256, 0, st>>>(c->d_xstat, (int)c->slot_cap, (int)c->prm.max_points, (HEnt*)c->d_htab, c->d_rec_tpos,
                                                  c->d_vkeys, c->d_vslot, c->d_vox_hpos, c->d_vox_count); }
  WC_CUDA(c, cudaEventRecord(c->ev[2], st));
  // final order by timestamp: enqueued without waiting for the surfel count (the kernels read it on the device), so the
  // whole extraction has ONE host synchronisation
  c->n_launches += 4;
  bsort_scan<<<1, 1024, 0, st>>>(c->d_bcnt, c->d_boff, c->d_bcur);
  bsort_scatter<<<c->num_sms * 2, 256, 0, st>>>(c->d_sort_hi, c->d_xstat, E.surf_cap, E.t_first, E.bscale, c->d_boff, c->d_bcur, c->d_sort_perm);
  bsort_rank<<<SORT_NB / 8, 256, 0, st>>>(c->d_sort_hi, c->d_sort_lo, c->d_boff, c->d_sort_perm, c->d_bcnt, c->d_sort_idx);
  gather_surfels<<<c->num_sms * 8, 256, 0, st>>>(c->d_surf_raw, c->d_sort_idx, c->d_xstat, E.surf_cap, c->d_surf);
  WC_CUDA(c, cudaEventRecord(c->ev[3], st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  const wc_extract_status hs = *c->h_xstat;
  if (hs.err_time_order) WC_FAIL(c, WC_EINVAL_TIME_ORDER, "point timestamps are not non-decreasing");
  if (hs.err_range) WC_FAIL(c, WC_EINVAL, "sweep exceeds the key range (+-16384 voxels around the first point, 128 s)");
  if (hs.err_capacity) WC_FAIL(c, WC_ECAPACITY, "capacity exceeded (slots=%d voxels=%d surfels=%d)", hs.n_slots, hs.n_voxels, hs.n_surfels);
  const int S = hs.n_surfels;
  c->n_surfels = (size_t)S;
  c->last_slots = hs.n_slots, c->last_voxels = hs.n_voxels;
  float ms;
  if (gpu_ms_keys) { cudaEventElapsedTime(&ms, c->ev[4], c->ev[1]); *gpu_ms_keys = ms; }
  if (gpu_ms_emit) { cudaEventElapsedTime(&ms, c->ev[1], c->ev[3]); *gpu_ms_emit = ms; }
  if (gpu_ms_total) { cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]); *gpu_ms_total = ms; }
  if (n_out) *n_out = (size_t)S;
  return WC_OK;
}

extern "C" wc_status wc_surfels_fetch(wc_ctx* c, wc_surfel* out, size_t cap, size_t* n_out) {
  if (!c || (!out && c->n_surfels)) return WC_EINVAL;
  if (n_out) *n_out = c->n_surfels;
  if (c->n_surfels > cap) WC_FAIL(c, WC_ECAPACITY, "%zu surfels exceed the output capacity %zu", c->n_surfels, cap);
  if (c->n_surfels)
    WC_CUDA(c, cudaMemcpyAsync(out, c->d_surf, c->n_surfels * sizeof(wc_surfel), cudaMemcpyDeviceToHost, c->stream));
  WC_CUDA(c, cudaStreamSynchronize(c->stream));
  return WC_OK;
}

extern "C" wc_status wc_build_surfels(wc_ctx* c, const wc_point48* pts, size_t n, wc_surfel* out, size_t cap,
                                      size_t* n_out, wc_point_assign* assign, double* gpu_ms) {
  if (!c) return WC_EINVAL;
  wc_status s = wc_points_upload(c, pts, n);
  if (s) return s;
  c->want_assign = assign != nullptr;
  size_t S = 0;
  s        = wc_build_surfels_resident(c, &S, nullptr, nullptr, gpu_ms);
  c->want_assign = 0;
  if (s) return s;
  if (assign && n) WC_CUDA(c, cudaMemcpyAsync(assign, c->d_assign, n * sizeof(wc_point_assign), cudaMemcpyDeviceToHost, c->stream));
  return wc_surfels_fetch(c, out, cap, n_out);
}
