// Surfel correspondence search on the device — replaces KnnSurfelMatcher::{BuildIndex, Match, KNearestSearch,
// ToVector, FLANNBuildIndex, FLANNKNearestSearch} (src/odometry/knn_surfel_matcher.cc:3-98).
//
//   surfel_features   world centre / normal of every surfel (surfel.h:67-80) and the 6-D feature
//                     [c_w / 1.0, n_w / (5 deg)] (knn_surfel_matcher.cc:91-98)
//   knn6_bruteforce   exact k nearest neighbours in squared L2 (flann::L2_Simple accumulation order, no FMA
//                     contraction), targets staged through shared memory; ties ordered by target index
//   gate_candidates   the three gates of Match (:26-34) applied to the k candidates in ascending distance
//   resolve_pairs     the order-dependent pair de-duplication (:35-39) as a parallel fixed-point iteration:
//                     query i takes its first gated candidate c unless c < i already took i.  acc[i] depends
//                     only on acc[c] for c < i, so the fixed point is unique and equals the sequential result.
//   compact_pairs     output in query order, each pair ordered by time (:41-45)
#include <stdlib.h>

#include "wc_ctx.h"
#include "wc_device_math.cuh"

using namespace wcd;

namespace {

constexpr int KMAX  = 16;
constexpr int FSTR  = 16;  // doubles per surfel feature record: f[6], c_w[3], n_w[3], t, pad[3]
constexpr int TILE  = 256;

__global__ void surfel_features(const wc_surfel* __restrict__ s, int n, double inv_center, double ang, double* __restrict__ f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Q4 q = ldq(s[i].rot);
  const V3 c = q * ld3(s[i].center) + ld3(s[i].pos);
  const V3 m = q * ld3(s[i].norm);
  double*  o = f + (size_t)i * FSTR;
  // Vector3d / double  (Eigen: per-coefficient division)
  o[0] = __ddiv_rn(c.x, inv_center), o[1] = __ddiv_rn(c.y, inv_center), o[2] = __ddiv_rn(c.z, inv_center);
  o[3] = __ddiv_rn(m.x, ang), o[4] = __ddiv_rn(m.y, ang), o[5] = __ddiv_rn(m.z, ang);
  o[6] = c.x, o[7] = c.y, o[8] = c.z, o[9] = m.x, o[10] = m.y, o[11] = m.z;
  o[12] = s[i].timestamp, o[13] = o[14] = o[15] = 0.0;
}

// q: nq records of stride qstr (first 6 doubles = feature); t likewise.
template <int KM>
__global__ void __launch_bounds__(TILE)
knn6_bruteforce(const double* __restrict__ q, int qstr, int nq, const double* __restrict__ t, int tstr, int nt, int k,
                int* __restrict__ out_idx, double* __restrict__ out_d2) {
  __shared__ double tile[TILE * 6];
  const int  i      = blockIdx.x * TILE + threadIdx.x;
  const bool active = i < nq;
  double     f[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) f[d] = active ? q[(size_t)i * qstr + d] : 0.0;
  double dk[KM];
  int    ik[KM];
#pragma unroll
  for (int j = 0; j < KM; ++j) dk[j] = INFINITY, ik[j] = -1;
  double worst = INFINITY;  // dk[k-1]
  for (int base = 0; base < nt; base += TILE) {
    const int m = min(TILE, nt - base);
    __syncthreads();
    for (int e = threadIdx.x; e < m * 6; e += TILE) tile[e] = t[(size_t)(base + e / 6) * tstr + (e % 6)];
    __syncthreads();
    if (!active) continue;
    for (int j = 0; j < m; ++j) {
      // flann::L2_Simple: result += diff*diff, in dimension order, no contraction
      double r = 0.0;
#pragma unroll
      for (int d = 0; d < 6; ++d) {
        const double diff = __dsub_rn(f[d], tile[j * 6 + d]);
        r                 = __dadd_rn(r, __dmul_rn(diff, diff));
      }
      if (r < worst) {
        const int id = base + j;
#pragma unroll
        for (int p = KM - 1; p > 0; --p) {
          if (p < k) {
            const bool shift = r < dk[p - 1];
            const bool place = !shift && r < dk[p];
            if (shift) dk[p] = dk[p - 1], ik[p] = ik[p - 1];
            else if (place) dk[p] = r, ik[p] = id;
          }
        }
        if (r < dk[0]) dk[0] = r, ik[0] = id;
#pragma unroll
        for (int p = 0; p < KM; ++p)
          if (p == k - 1) worst = dk[p];
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int j = 0; j < KM; ++j)
    if (j < k) out_idx[(size_t)i * k + j] = ik[j], out_d2[(size_t)i * k + j] = dk[j];
}


// ---- exact kNN through a uniform grid over the centre part of the feature (cell = 1 feature unit) ------------------
// Targets are bucketed by cell (open-addressed cell table -> dense cell id -> counting-sort scatter).  A query scans
// Chebyshev rings of cells around its own cell; after ring r every unvisited target differs from the query by at least
// b = (distance from the query to the faces of the visited cube) in one centre coordinate, so its squared 6-D distance
// is >= b^2: the search stops as soon as the k-th best distance is strictly below b^2 — the result is the exact k-NN,
// identical to brute force, ordered by (distance, target index).

__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}
__device__ __forceinline__ unsigned long long cell_key(long long ix, long long iy, long long iz) {
  return ((unsigned long long)(ix + (1 << 20)) << 42) | ((unsigned long long)(iy + (1 << 20)) << 21) | (unsigned long long)(iz + (1 << 20));
}
#define WC_CELL_EMPTY 0xFFFFFFFFFFFFFFFFull

struct GridBufs {
  unsigned long long* keys;  // cell table
  int*                cid;   // published dense cell id
  unsigned long long  mask;
  int*                cnt;   // per cell
  int*                off;
  int*                cur;
  int*                hpos;  // table position per cell (cleanup)
  int*                tcell; // per target
  int*                ncells;
  double*             sfeat; // sorted by cell, structure of arrays: 6 feature columns + the target index column, stride scap
  size_t              scap;
  double              cs;    // grid cells per feature unit along the three centre axes (a power of two: f * cs is exact)
  double*             cbox;  // per cell: min[6], max[6] of the member features
  unsigned long long* ckey;  // per cell: its key
  int*                err;
};

__global__ void grid_count(const double* __restrict__ tf, int tstr, int nt, GridBufs G, int cell_cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  const double* f  = tf + (size_t)i * tstr;
  const double  fx = floor(f[0] * G.cs), fy = floor(f[1] * G.cs), fz = floor(f[2] * G.cs);
  if (!(fabs(fx) < 1e6 && fabs(fy) < 1e6 && fabs(fz) < 1e6)) {
    *G.err = 1;  // outside the 2^20-cell key range (or non-finite)
    G.tcell[i] = -1;
    return;
  }
  const unsigned long long key = cell_key((long long)fx, (long long)fy, (long long)fz);
  unsigned long long       h   = mix64(key) & G.mask;
  int                      id  = -2;
  for (;; h = (h + 1) & G.mask) {
    unsigned long long k = *((volatile unsigned long long*)&G.keys[h]);
    if (k == WC_CELL_EMPTY) {
      k = atomicCAS(&G.keys[h], WC_CELL_EMPTY, key);
      if (k == WC_CELL_EMPTY) {
        id = atomicAdd(G.ncells, 1);
        if (id >= cell_cap) *G.err = 1, id = -2;
        else G.hpos[id] = (int)h, G.ckey[id] = key;
        atomicExch(&G.cid[h], id);
        break;
      }
    }
    if (k == key) {
      while ((id = *((volatile int*)&G.cid[h])) == -1) {
      }
      break;
    }
  }
  G.tcell[i] = id;
  if (id >= 0) atomicAdd(&G.cnt[id], 1);
}

__global__ void __launch_bounds__(1024) grid_scan(GridBufs G) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int      n = *G.ncells;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? G.cnt[i] : 0;
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
      warp_sums[threadIdx.x] = wi - w;
    }
    __syncthreads();
    const int excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
    if (i < n) G.off[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) G.off[n] = carry;
}

__global__ void grid_scatter(const double* __restrict__ tf, int tstr, int nt, GridBufs G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  const int id = G.tcell[i];
  if (id < 0) return;
  const int     pos = G.off[id] + atomicAdd(&G.cur[id], 1);
  const double* f   = tf + (size_t)i * tstr;
#pragma unroll
  for (int d = 0; d < 6; ++d) G.sfeat[d * G.scap + pos] = f[d];
  G.sfeat[6 * G.scap + pos] = __longlong_as_double((long long)i);
}

// 6-D bounding box of every cell's members (for the pruned scan of the cells the ring search does not reach): eight
// lanes per cell stride over its members, then a 3-step shuffle reduction
__global__ void __launch_bounds__(256) grid_boxes(GridBufs G) {
  const int n = *G.ncells, sub = threadIdx.x & 7;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  double    lo[6], hi[6];
#pragma unroll
  for (int d = 0; d < 6; ++d) lo[d] = INFINITY, hi[d] = -INFINITY;
  if (c < n)
    for (int p = G.off[c] + sub; p < G.off[c + 1]; p += 8) {
#pragma unroll
      for (int d = 0; d < 6; ++d) {
        const double td = G.sfeat[d * G.scap + p];
        lo[d] = fmin(lo[d], td), hi[d] = fmax(hi[d], td);
      }
    }
#pragma unroll
  for (int s = 4; s > 0; s >>= 1)
#pragma unroll
    for (int d = 0; d < 6; ++d) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
    }
  if (c < n && sub == 0)
#pragma unroll
    for (int d = 0; d < 6; ++d) G.cbox[(size_t)c * 12 + d] = lo[d], G.cbox[(size_t)c * 12 + 6 + d] = hi[d];
}

__global__ void grid_cleanup(GridBufs G) {
  const int n = *G.ncells;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int h = G.hpos[i];
    G.keys[h]   = WC_CELL_EMPTY;
    G.cid[h]    = -1;
    G.cnt[i]    = 0;
    G.cur[i]    = 0;
  }
}
__global__ void grid_reset_count(GridBufs G) { *G.ncells = 0; }

__device__ __forceinline__ bool cand_less(double d, int id, double dk, int ik) { return d < dk || (d == dk && id < ik); }
#ifdef WC_KNN_STATS
__device__ unsigned long long g_knn_stats[8];  // queries, phase-1 queries, members scanned (phase 0 / 1), cells scanned in phase 1, inserts
__global__ void knn_stats_print() {
  printf("knn stats: queries %llu ring2 %llu phase1 %llu members p0 %llu p0.5+p1 %llu p1-cells-taken %llu inserts %llu\n", g_knn_stats[0],
         g_knn_stats[6], g_knn_stats[1], g_knn_stats[2], g_knn_stats[3], g_knn_stats[4], g_knn_stats[5]);
  for (int i = 0; i < 8; ++i) g_knn_stats[i] = 0;
}
#define KSTAT(i, v) do { if (lane == 0) atomicAdd(&g_knn_stats[i], (unsigned long long)(v)); } while (0)
#else
#define KSTAT(i, v)
#endif

// One warp per query.  The sorted candidate list lives in registers, one entry per lane (lane j = j-th best, k <= 32),
// so an insertion is a ballot + popc + shfl_up instead of a per-thread array shuffle, and all 32 lanes always work on
// the same query: the ring cells are probed in parallel (one cell per lane), the members of a cell are scanned 32 at a
// time with coalesced 64-byte loads, and the far cells are box-tested 32 at a time.
//   phase 0: the 27 cells of Chebyshev rings 0..1 around the query's cell; stop if the k-th best distance is strictly
//            inside the visited cube (nothing outside can be closer);
//   phase 1: every other cell whose 6-D bounding box can still hold a better (or tying) candidate.
// Exact k-NN in (distance, target index) order — identical to the brute-force scan.
__global__ void __launch_bounds__(256)
knn6_warp(const double* __restrict__ q, int qstr, int nq, int k, GridBufs G, int* __restrict__ out_idx, double* __restrict__ out_d2) {
  const int lane = threadIdx.x & 31;
  const int nw   = gridDim.x * (blockDim.x >> 5);
  const int nc   = *G.ncells;
#pragma unroll 1
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < nq; i += nw) {
    double f[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) f[d] = q[(size_t)i * qstr + d];
    double my_d = INFINITY, worst = INFINITY;  // lane j holds the j-th best candidate
    int    my_i = 0x7fffffff, worsti = 0x7fffffff;

    // scan the members [p0, p1) of one cell, 32 per step, inserting the ones that beat the current k-th best
    bool fresh = true;  // warp-uniform: the list is still empty
    // one step: lane holds sorted-feature row p (or -1); the candidates that beat the current k-th best are inserted
    auto scan_batch = [&](int p) {
      double cd = INFINITY;
      int    ci = 0x7fffffff;
      if (p >= 0) {
        double td[7];  // coalesced column loads, all issued before the first use
#pragma unroll
        for (int d = 0; d < 7; ++d) td[d] = G.sfeat[d * G.scap + p];
        double rr = 0.0;  // flann::L2_Simple accumulation order, no contraction
#pragma unroll
        for (int d = 0; d < 6; ++d) {
          const double diff = __dsub_rn(f[d], td[d]);
          rr                = __dadd_rn(rr, __dmul_rn(diff, diff));
        }
        cd = rr, ci = (int)__double_as_longlong(td[6]);
      }
      if (fresh) {
        // empty list (first batch of the query): a 32-lane bitonic sort of the batch IS the list — 15 compare-exchange
        // steps instead of ~20 serial insertions; lanes >= k keep larger candidates, which the insert path ignores
        if (__any_sync(0xffffffffu, ci != 0x7fffffff)) {
#pragma unroll
          for (int kk = 2; kk <= 32; kk <<= 1)
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
              const double od = __shfl_xor_sync(0xffffffffu, cd, j);
              const int    oi = __shfl_xor_sync(0xffffffffu, ci, j);
              const bool   up = (lane & kk) == 0, lower = (lane & j) == 0;
              const bool   o_less = cand_less(od, oi, cd, ci);
              if ((lower == up) ? o_less : !o_less && !(od == cd && oi == ci)) cd = od, ci = oi;
            }
          my_d = cd, my_i = ci;
          worst  = __shfl_sync(0xffffffffu, my_d, k - 1);
          worsti = __shfl_sync(0xffffffffu, my_i, k - 1);
          fresh  = false;
        }
        return;
      }
      unsigned m = __ballot_sync(0xffffffffu, cand_less(cd, ci, worst, worsti));
      while (m) {
        const int    src = __ffs(m) - 1;
        const double bd  = __shfl_sync(0xffffffffu, cd, src);
        const int    bi  = __shfl_sync(0xffffffffu, ci, src);
        m &= m - 1;
        if (!cand_less(bd, bi, worst, worsti)) continue;  // the list moved on since the ballot
        KSTAT(5, 1);
        const int    pos = __popc(__ballot_sync(0xffffffffu, lane < k && cand_less(my_d, my_i, bd, bi)));
        const double ud  = __shfl_up_sync(0xffffffffu, my_d, 1);
        const int    ui  = __shfl_up_sync(0xffffffffu, my_i, 1);
        if (lane == pos) my_d = bd, my_i = bi;
        else if (lane > pos) my_d = ud, my_i = ui;
        worst  = __shfl_sync(0xffffffffu, my_d, k - 1);
        worsti = __shfl_sync(0xffffffffu, my_i, k - 1);
      }
    };
    int  kphase    = 0;
    // One cell per lane (c0 = first member row, cn = member count or 0, lb = lower bound of the 6-D distance from the
    // query to anything in the cell): the cells are taken nearest first and their members are scanned as ONE flat index
    // space, 32 per step, until the bound of the next cell exceeds the k-th best distance.
    auto flat_scan = [&](int c0, int cn, double lb, const bool sorted) {
        if (!__any_sync(0xffffffffu, cn > 0)) return;  // a round of misses
        // Nearest cells first (32-lane bitonic sort on the box bound): the members of cells whose bound already exceeds the
        // k-th best distance are never loaded — a wall query skips the floor surfels that share its cells, because their
        // normals put the whole cell far away in feature space — and the scan stops at the first such cell.  Exact: a
        // skipped cell cannot hold a better or tying candidate (same test as the far-cell phase below).
        if (sorted) {
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1)
#pragma unroll
          for (int j = kk >> 1; j > 0; j >>= 1) {
            const double ol = __shfl_xor_sync(0xffffffffu, lb, j);
            const int    o0 = __shfl_xor_sync(0xffffffffu, c0, j), on = __shfl_xor_sync(0xffffffffu, cn, j);
            const bool   up = (lane & kk) == 0, lower = (lane & j) == 0;
            // ascending by (lb, c0): take the partner's entry when it belongs on this side
            const bool o_less = ol < lb || (ol == lb && o0 < c0);
            const bool differ = ol != lb || o0 != c0;
            if (differ && ((lower == up) ? o_less : !o_less)) lb = ol, c0 = o0, cn = on;
          }
        }
        int incl = cn;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl  = incl - cn;
#pragma unroll 1
        for (int base = 0; base < total; base += 32) {
          const int g  = base + lane;
          int       lo = 0;  // last cell whose exclusive prefix is <= g
#pragma unroll
          for (int stp = 16; stp > 0; stp >>= 1) {
            const int pv = __shfl_sync(0xffffffffu, excl, lo + stp);  // (lo + stp <= 31)
            if (pv <= g) lo += stp;
          }
          const int    ce = __shfl_sync(0xffffffffu, excl, lo), cs = __shfl_sync(0xffffffffu, c0, lo);
          const double cl = __shfl_sync(0xffffffffu, lb, lo);
          // sorted: lane 0 holds the batch's nearest cell — if even that one is out of reach, so is everything after it
          if (sorted && __shfl_sync(0xffffffffu, cl, 0) * (1.0 - 1e-12) > worst) break;
          const bool take = g < total && !(cl * (1.0 - 1e-12) > worst);
          if (!__any_sync(0xffffffffu, take)) continue;
#ifdef WC_KNN_STATS
          const int n_take = __popc(__ballot_sync(0xffffffffu, take));
          KSTAT(2 + kphase, n_take);
#endif
          scan_batch(take ? cs + (g - ce) : -1);
        }
    };
    const double    gx = f[0] * G.cs, gy = f[1] * G.cs, gz = f[2] * G.cs;  // grid coordinates of the query centre
    const double    cfx = floor(gx), cfy = floor(gy), cfz = floor(gz);
    const bool      rings_ok = fabs(cfx) < 1e6 && fabs(cfy) < 1e6 && fabs(cfz) < 1e6;
    const long long ix = rings_ok ? (long long)cfx : 0, iy = rings_ok ? (long long)cfy : 0, iz = rings_ok ? (long long)cfz : 0;
    bool            done = false;
    int             ring = 1;  // half-width of the cube of cells already scanned
    if (rings_ok) {
      // probe one cell per lane (cn = 0 for absent cells), then scan the members of all 32 lanes' cells as ONE flat
      // index space (inclusive prefix of the cell sizes over the lanes): every step scans 32 members regardless of how
      // they are spread over the cells
      auto probe_and_scan = [&](bool active, int dx, int dy, int dz) {
        int    c0 = 0, cn = 0;
        double lb = INFINITY;  // lower bound of the 6-D distance from the query to anything in this lane's cell
        if (active) {
          const unsigned long long key = cell_key(ix + dx, iy + dy, iz + dz);
          unsigned long long       h   = mix64(key) & G.mask;
          for (;; h = (h + 1) & G.mask) {
            const unsigned long long kk = G.keys[h];
            if (kk == key) {
              const int id = G.cid[h];
              c0 = G.off[id], cn = G.off[id + 1] - c0;
              const double* bx = G.cbox + (size_t)id * 12;  // 6-D bounding box of the cell's members (centre AND normal part)
              lb               = 0.0;
#pragma unroll
              for (int d = 0; d < 6; ++d) {
                const double ee = fmax(fmax(bx[d] - f[d], f[d] - bx[6 + d]), 0.0);
                lb += ee * ee;
              }
              break;
            }
            if (kk == WC_CELL_EMPTY) break;
          }
        }
        flat_scan(c0, cn, lb, true);
      };
      // phase 0: the 3 x 3 x 3 block; lane 0 takes the query's own cell so that its members come first and tighten the
      // k-th distance early
      {
        const int cc = lane == 0 ? 13 : (lane <= 13 ? lane - 1 : lane);
        probe_and_scan(lane < 27, cc % 3 - 1, (cc / 3) % 3 - 1, cc / 9 - 1);
      }
      // distance from the query to the nearest face of the scanned cube, in feature units
      double b = fmin(fmin(fmin(gx - (cfx - 1), (cfx + 2) - gx), fmin(gy - (cfy - 1), (cfy + 2) - gy)),
                      fmin(gz - (cfz - 1), (cfz + 2) - gz)) / G.cs;
      done = worst < b * b * (1.0 - 1e-12);
      // phase 0.5: the shell of the 5 x 5 x 5 block, if the k-th best distance still reaches outside the scanned cube — for
      // a sparse target set this settles most queries before the global box scan.  (Measured at C3: going on to the 7^3 and
      // 9^3 shells before the global scan is slower, 907 vs 782 us for both matchers.)
      KSTAT(6, done ? 0 : 1);
#pragma unroll 1
      for (int r = 2; r <= 2 && !done; ++r) {
        kphase = 1;
        const int w = 2 * r + 1, w3 = w * w * w;
#pragma unroll 1
        for (int o0 = 0; o0 < w3; o0 += 32) {
          const int  o  = o0 + lane;
          const int  dx = o % w - r, dy = (o / w) % w - r, dz = o / (w * w) - r;
          const bool shell = o < w3 && (abs(dx) == r || abs(dy) == r || abs(dz) == r);
          probe_and_scan(shell, dx, dy, dz);
        }
        ring = r;
        const double b = fmin(fmin(fmin(gx - (cfx - r), (cfx + r + 1) - gx), fmin(gy - (cfy - r), (cfy + r + 1) - gy)),
                              fmin(gz - (cfz - r), (cfz + r + 1) - gz)) / G.cs;
        done = worst < b * b * (1.0 - 1e-12);
      }
    }
    KSTAT(0, 1);
    if (!done) {
      KSTAT(1, 1);
      kphase = 1;
      // phase 1: box-test the cells outside the scanned cube 32 at a time; the survivors of a round are scanned together
      // (nearest first, flat over their members — most cells of a sparse index hold a handful of surfels)
#pragma unroll 1
      for (int cb = 0; cb < nc; cb += 32) {
        const int cell = cb + lane;
        int       p0 = 0, pn = 0;
        double    lb = INFINITY;
        if (cell < nc) {
          bool visited = false;
          if (rings_ok) {
            const unsigned long long key = G.ckey[cell];
            const long long cx = (long long)(key >> 42) - (1 << 20), cy = (long long)((key >> 21) & 0x1fffff) - (1 << 20),
                            cz = (long long)(key & 0x1fffff) - (1 << 20);
            visited = llabs(cx - ix) <= ring && llabs(cy - iy) <= ring && llabs(cz - iz) <= ring;
          }
          if (!visited) {
            const double* bx = G.cbox + (size_t)cell * 12;
            double        l2 = 0.0;
#pragma unroll
            for (int d = 0; d < 6; ++d) {
              const double ee = fmax(fmax(bx[d] - f[d], f[d] - bx[6 + d]), 0.0);
              l2 += ee * ee;
            }
            if (!(l2 * (1.0 - 1e-12) > worst)) {  // skipped only if it cannot hold a better or tying candidate
              lb = l2;
              p0 = G.off[cell], pn = G.off[cell + 1] - p0;
            }
          }
        }
#ifdef WC_KNN_STATS
        const int n_cells_taken = __popc(__ballot_sync(0xffffffffu, pn > 0));
        KSTAT(4, n_cells_taken);
#endif
        flat_scan(p0, pn, lb, false);  // (no ordering here: a sort per 32-cell round costs more than it prunes)
      }
    }
    if (lane < k) out_idx[(size_t)i * k + lane] = (my_i == 0x7fffffff) ? -1 : my_i, out_d2[(size_t)i * k + lane] = my_d;
  }
}

struct GateParams {
  double time_diff, ang, dist;
  int    k;
};

// gated[i*k + j]: the candidates of query i that pass the gates, ascending distance, -1 terminated
__global__ void gate_candidates(const double* __restrict__ qf, int nq, const double* __restrict__ tf, const int* __restrict__ knn,
                                GateParams G, int* __restrict__ gated) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const double* a  = qf + (size_t)i * FSTR;
  const V3      ca = ld3(a + 6), na = ld3(a + 9);
  const double  ta = a[12];
  int           w  = 0;
  for (int j = 0; j < G.k; ++j) {
    const int c = knn[(size_t)i * G.k + j];
    if (c < 0) break;
    const double* b = tf + (size_t)c * FSTR;
    if (fabs(__dsub_rn(b[12], ta)) < G.time_diff) continue;                       // knn_surfel_matcher.cc:26
    const V3 nb = ld3(b + 9), cb = ld3(b + 6);
    const double dn = __dadd_rn(__dadd_rn(__dmul_rn(na.x, nb.x), __dmul_rn(na.y, nb.y)), __dmul_rn(na.z, nb.z));
    if (acos(dn) > G.ang) continue;                                               // :29, un-clamped acos (Q9)
    const V3     d  = ca - cb;
    const double pd = __dadd_rn(__dadd_rn(__dmul_rn(na.x, d.x), __dmul_rn(na.y, d.y)), __dmul_rn(na.z, d.z));
    if (fabs(pd) > G.dist) continue;                                              // :32
    gated[(size_t)i * G.k + w++] = c;
  }
  for (; w < G.k; ++w) gated[(size_t)i * G.k + w] = -1;
}

// one relaxation sweep of the de-duplication recurrence; *changed is set if any entry moved
__global__ void resolve_pairs(const int* __restrict__ gated, int nq, int k, int self_match, const int* __restrict__ acc_in,
                              int* __restrict__ acc_out, int* __restrict__ changed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int take = -1;
  for (int j = 0; j < k; ++j) {
    const int c = gated[(size_t)i * k + j];
    if (c < 0) break;
    // surfel_pairs already holds {c, i} iff the earlier query c accepted i  (:35-39)
    if (self_match && c < i && acc_in[c] == i) continue;
    take = c;
    break;
  }
  if (take != acc_in[i]) *changed = 1;
  acc_out[i] = take;
}

// Ordered compaction of the accepted pairs in two launches over all SMs: per-CTA counts, then every CTA sums the counts
// of the CTAs before it (a few dozen values), scans its own 1024 queries and scatters — output order = query order.
__global__ void __launch_bounds__(1024) pair_count(const int* __restrict__ acc, int nq, int* __restrict__ blk_cnt) {
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int n = __syncthreads_count(i < nq && acc[i] >= 0);
  if (threadIdx.x == 0) blk_cnt[blockIdx.x] = n;
}
__global__ void __launch_bounds__(1024)
pair_scatter(const int* __restrict__ acc, int nq, const int* __restrict__ blk_cnt, const double* __restrict__ qf,
             const double* __restrict__ tf, wc_corr_idx* __restrict__ out, unsigned char* __restrict__ first_is_target,
             int* __restrict__ n_out /* [0] pair count, [2] set when a pair has the QUERY first */) {
  __shared__ int warp_sums[32];
  __shared__ int s_base;
  const int t = threadIdx.x, lane = t & 31, i = blockIdx.x * 1024 + t;
  if (t < 32) {
    int s = 0;
    for (int b = lane; b < (int)blockIdx.x; b += 32) s += blk_cnt[b];
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if (lane == 0) s_base = s;
  }
  const int c = i < nq ? acc[i] : -1;
  const int v = c >= 0;
  int       incl = v;
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
  const int pos = s_base + warp_sums[t >> 5] + incl - v;
  if (v) {
    const bool query_first = qf[(size_t)i * FSTR + 12] < tf[(size_t)c * FSTR + 12];  // knn_surfel_matcher.cc:41
    out[pos].s1            = query_first ? i : c;
    out[pos].s2            = query_first ? c : i;
    first_is_target[pos]   = query_first ? 0 : 1;
    if (query_first) n_out[2] = 1;
  }
  if (blockIdx.x == gridDim.x - 1 && t == 1023) *n_out = pos + v;
}

}  // namespace

static wc_status match_alloc(wc_ctx* c) {
  if (c->d_qfeat) return WC_OK;
  const size_t ns = (size_t)c->prm.max_surfels;
  WC_CUDA(c, cudaMalloc(&c->d_msurf_q, ns * sizeof(wc_surfel)));
  WC_CUDA(c, cudaMalloc(&c->d_msurf_t, ns * sizeof(wc_surfel)));
  WC_CUDA(c, cudaMalloc(&c->d_qfeat, ns * FSTR * 8));
  WC_CUDA(c, cudaMalloc(&c->d_tfeat, ns * FSTR * 8));
  WC_CUDA(c, cudaMalloc(&c->d_knn_idx, ns * KMAX * 4));
  WC_CUDA(c, cudaMalloc(&c->d_knn_d2, ns * KMAX * 8));
  WC_CUDA(c, cudaMalloc(&c->d_gated, ns * KMAX * 4));
  WC_CUDA(c, cudaMalloc(&c->d_acc, ns * 4));
  WC_CUDA(c, cudaMalloc(&c->d_acc2, ns * 4));
  WC_CUDA(c, cudaMalloc(&c->d_flag, 16));
  WC_CUDA(c, cudaMalloc(&c->d_scan_tmp, (ns / 1024 + 2) * 4));
  WC_CUDA(c, cudaMalloc(&c->d_corr_out, ns * sizeof(wc_corr_idx)));
  WC_CUDA(c, cudaMalloc(&c->d_fit_out, ns));
  WC_CUDA(c, cudaMallocHost(&c->h_flag, 16));
  // uniform-grid index over the targets
  GridBufs* G = (GridBufs*)calloc(1, sizeof(GridBufs));
  c->d_grid   = G;
  const size_t cap = wc_next_pow2(2 * ns);
  G->mask          = cap - 1;
  WC_CUDA(c, cudaMalloc(&G->keys, cap * 8));
  WC_CUDA(c, cudaMalloc(&G->cid, cap * 4));
  WC_CUDA(c, cudaMalloc(&G->cnt, (ns + 1) * 4));
  WC_CUDA(c, cudaMalloc(&G->off, (ns + 1) * 4));
  WC_CUDA(c, cudaMalloc(&G->cur, (ns + 1) * 4));
  WC_CUDA(c, cudaMalloc(&G->hpos, (ns + 1) * 4));
  WC_CUDA(c, cudaMalloc(&G->tcell, ns * 4));
  WC_CUDA(c, cudaMalloc(&G->ncells, 8));
  G->err = c->d_flag + 2;
  G->scap = ns;
  WC_CUDA(c, cudaMalloc(&G->sfeat, ns * 7 * 8));
  WC_CUDA(c, cudaMalloc(&G->cbox, (ns + 1) * 12 * 8));
  WC_CUDA(c, cudaMalloc(&G->ckey, (ns + 1) * 8));
  WC_CUDA(c, cudaMemsetAsync(G->keys, 0xff, cap * 8, c->stream));
  WC_CUDA(c, cudaMemsetAsync(G->cid, 0xff, cap * 4, c->stream));
  WC_CUDA(c, cudaMemsetAsync(G->cnt, 0, (ns + 1) * 4, c->stream));
  WC_CUDA(c, cudaMemsetAsync(G->cur, 0, (ns + 1) * 4, c->stream));
  WC_CUDA(c, cudaMemsetAsync(G->ncells, 0, 8, c->stream));
  return WC_OK;
}

void wc_match_free(wc_ctx* c) {
  void* ptrs[] = {c->d_msurf_q, c->d_msurf_t, c->d_qfeat, c->d_tfeat, c->d_knn_idx, c->d_knn_d2, c->d_gated,
                  c->d_acc,     c->d_acc2,    c->d_flag,  c->d_scan_tmp, c->d_corr_out, c->d_fit_out, c->d_part_d, c->d_part_i};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (c->h_flag) cudaFreeHost(c->h_flag);
  GridBufs* G = (GridBufs*)c->d_grid;
  if (G) {
    void* gp[] = {G->keys, G->cid, G->cnt, G->off, G->cur, G->hpos, G->tcell, G->ncells, G->sfeat, G->cbox, G->ckey};
    for (void* p : gp)
      if (p) cudaFree(p);
    free(G);
    c->d_grid = nullptr;
  }
}

int*      wc_comm_gather_region(wc_ctx* c);  // wc_comm.cu
wc_status wc_comm_check(wc_ctx* c);
wc_status wc_comm_allgather_rows(wc_ctx* c, int* dst, const int* row0, int width);

// Device-resident matcher core: query/target surfels already at d_q / d_t.  Leaves the pairs in d_corr_out.
// With world > 1 the exact k-NN search (the bulk of the matcher) is sharded over the ranks by query blocks and the index
// lists are all-gathered through NVLink peer memory; index build, gating and pair resolution are replicated (they are
// order-dependent across queries and cheap), so every rank ends with the identical correspondence list.
wc_status wc_match_device(wc_ctx* c, const wc_surfel* d_q, size_t nq, const wc_surfel* d_t, size_t nt, int self_match,
                          size_t* n_out) {
  *n_out = 0;
  if (nq == 0 || nt == 0) return WC_OK;  // knn_surfel_matcher.cc:18-20
  wc_status as = match_alloc(c);
  if (as) return as;
  cudaStream_t st = c->stream;
  WC_CUDA(c, cudaMemsetAsync(c->d_flag, 0, 16, st));
  const int    k  = c->prm.knn_candidates;
  if (k < 1 || k > KMAX) WC_FAIL(c, WC_EINVAL, "knn_candidates must be 1..%d", KMAX);
  const unsigned gq = (unsigned)((nq + 255) / 256), gt = (unsigned)((nt + 255) / 256);
  { ++c->n_launches; surfel_features<<<gq, 256, 0, st>>>(d_q, (int)nq, c->prm.center_dist_threshold, c->prm.angular_dist_threshold, c->d_qfeat); }
  const double* tfeat = c->d_qfeat;
  if (!self_match) {
    { ++c->n_launches; surfel_features<<<gt, 256, 0, st>>>(d_t, (int)nt, c->prm.center_dist_threshold, c->prm.angular_dist_threshold, c->d_tfeat); }
    tfeat = c->d_tfeat;
  }
  // query shard of this rank
  const bool sharded = c->world > 1 && c->comm_ready && nq >= 1024;
  int        row0[9] = {0};
  for (int r = 0; r <= c->world && r <= 8; ++r) row0[r] = sharded ? (int)((nq * (size_t)r) / (size_t)c->world) : (r ? (int)nq : 0);
  const int     q0 = sharded ? row0[c->rank] : 0, nq_my = (sharded ? row0[c->rank + 1] : (int)nq) - q0;
  const double* qf_my   = c->d_qfeat + (size_t)q0 * FSTR;
  int*          knn_out = (sharded ? wc_comm_gather_region(c) : c->d_knn_idx) + (size_t)q0 * k;
  double*       d2_out  = c->d_knn_d2 + (size_t)q0 * k;
  if (nt < (size_t)c->knn_grid_min) {
    if (nq_my > 0) {
      if (k <= 10) { ++c->n_launches; knn6_bruteforce<10><<<(unsigned)((nq_my + TILE - 1) / TILE), TILE, 0, st>>>(qf_my, FSTR, nq_my, tfeat, FSTR, (int)nt, k,
                                                                           knn_out, d2_out); }
      else { ++c->n_launches; knn6_bruteforce<KMAX><<<(unsigned)((nq_my + TILE - 1) / TILE), TILE, 0, st>>>(qf_my, FSTR, nq_my, tfeat, FSTR, (int)nt, k,
                                                                           knn_out, d2_out); }
    }
  } else {
    GridBufs GB = *(GridBufs*)c->d_grid;
    // unit cells (1 m / 5 deg feature units): measured at C3, half-size cells scan fewer candidates in the 3 x 3 x 3 block
    // but escalate to the ring-2 / box-scan phases more often and end up slower (930 vs 890 us); exact either way
    GB.cs = 1.0;
    if (const char* e = getenv("WC_KNN_CELLS_PER_UNIT")) GB.cs = atof(e);  // test hook (1, 2, 4)
    { ++c->n_launches; grid_count<<<gt, 256, 0, st>>>(tfeat, FSTR, (int)nt, GB, (int)c->prm.max_surfels); }
    { ++c->n_launches; grid_scan<<<1, 1024, 0, st>>>(GB); }
    { ++c->n_launches; grid_scatter<<<gt, 256, 0, st>>>(tfeat, FSTR, (int)nt, GB); }
    { ++c->n_launches; grid_boxes<<<(unsigned)((nt * 8 + 255) / 256), 256, 0, st>>>(GB); }  // cells <= targets
    if (nq_my > 0) { ++c->n_launches; knn6_warp<<<c->num_sms * 8, 256, 0, st>>>(qf_my, FSTR, nq_my, k, GB, knn_out, d2_out); }
#ifdef WC_KNN_STATS
    knn_stats_print<<<1, 1, 0, st>>>();
    { int ncell = 0; cudaMemcpyAsync(&ncell, GB.ncells, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); printf("grid cells %d targets %zu\n", ncell, nt); }
#endif
    { ++c->n_launches; grid_cleanup<<<c->num_sms, 256, 0, st>>>(GB); }
    { ++c->n_launches; grid_reset_count<<<1, 1, 0, st>>>(GB); }
  }
  if (sharded) {
    wc_status gs = wc_comm_allgather_rows(c, c->d_knn_idx, row0, k);
    if (gs) return gs;
  }
  GateParams GP{c->prm.time_diff_threshold, c->prm.angular_dist_threshold, c->prm.surfel_dist_threshold, k};
  { ++c->n_launches; gate_candidates<<<gq, 256, 0, st>>>(c->d_qfeat, (int)nq, tfeat, c->d_knn_idx, GP, c->d_gated); }
  WC_CUDA(c, cudaMemsetAsync(c->d_acc, 0xff, nq * 4, st));
  int* a = c->d_acc;
  int* b = c->d_acc2;
  const unsigned gc = (unsigned)((nq + 1023) / 1024);
  for (int it = 0;; ++it) {
    // acc <- F(acc) is a deterministic sweep; it has converged when a sweep reproduces its input.  Eight sweeps (the
    // recurrence settles in a handful) AND the ordered compaction of their result are enqueued before the one host
    // check; the flag is cleared right before the last sweep so that it reports that sweep alone.  In the rare case that
    // the recurrence had not settled, more sweeps follow and the compaction is redone.
    const int sweeps = self_match ? 8 : 1;
    for (int s = 0; s < sweeps; ++s) {
      if (s == sweeps - 1) WC_CUDA(c, cudaMemsetAsync(c->d_flag, 0, 4, st));
      { ++c->n_launches; resolve_pairs<<<gq, 256, 0, st>>>(c->d_gated, (int)nq, k, self_match, a, b, c->d_flag); }
      int* tmp = a; a = b; b = tmp;
    }
    WC_CUDA(c, cudaMemsetAsync(c->d_flag + 3, 0, 4, st));
    { ++c->n_launches; pair_count<<<gc, 1024, 0, st>>>(a, (int)nq, c->d_scan_tmp); }
    { ++c->n_launches; pair_scatter<<<gc, 1024, 0, st>>>(a, (int)nq, c->d_scan_tmp, c->d_qfeat, tfeat, c->d_corr_out, c->d_fit_out, c->d_flag + 1); }
    WC_CUDA(c, cudaMemcpyAsync(c->h_flag, c->d_flag, 16, cudaMemcpyDeviceToHost, st));
    WC_CUDA(c, cudaStreamSynchronize(st));
    if (!self_match || c->h_flag[0] == 0) break;
    if (it > (int)nq) WC_FAIL(c, WC_ENUMERIC, "pair de-duplication did not converge");
  }
  WC_CUDA(c, cudaGetLastError());
  if (c->h_flag[2]) WC_FAIL(c, WC_EINVAL, "surfel centres outside the matcher grid range (+-1e6 cells) or non-finite");
  if (sharded) {
    wc_status cs = wc_comm_check(c);
    if (cs) return cs;
  }
  *n_out = (size_t)c->h_flag[1];
  c->match_query_first = c->h_flag[3];
  if (getenv("WC_DEBUG")) fprintf(stderr, "[wc_match] nq=%zu nt=%zu self=%d pairs=%d\n", nq, nt, self_match, c->h_flag[1]);
  return WC_OK;
}

extern "C" wc_status wc_match(wc_ctx* c, const wc_surfel* query, size_t nq, const wc_surfel* target, size_t nt,
                              int self_match, wc_corr_idx* out, size_t cap, size_t* n_out, uint8_t* first_is_target,
                              double* gpu_ms) {
  if (!c || !n_out || (nq && !query) || (nt && !target)) return WC_EINVAL;
  *n_out = 0;
  if (nq > (size_t)c->prm.max_surfels || nt > (size_t)c->prm.max_surfels)
    WC_FAIL(c, WC_ECAPACITY, "surfel count exceeds max_surfels=%lld", (long long)c->prm.max_surfels);
  if (self_match && (nq != nt)) WC_FAIL(c, WC_EINVAL, "self_match requires target == query");
  wc_status s = match_alloc(c);
  if (s) return s;
  cudaStream_t st = c->stream;
  if (nq) WC_CUDA(c, cudaMemcpyAsync(c->d_msurf_q, query, nq * sizeof(wc_surfel), cudaMemcpyHostToDevice, st));
  if (!self_match && nt) WC_CUDA(c, cudaMemcpyAsync(c->d_msurf_t, target, nt * sizeof(wc_surfel), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaEventRecord(c->ev[0], st));
  size_t np = 0;
  s         = wc_match_device(c, c->d_msurf_q, nq, self_match ? c->d_msurf_q : c->d_msurf_t, nt, self_match, &np);
  if (s) return s;
  WC_CUDA(c, cudaEventRecord(c->ev[1], st));
  if (np > cap) WC_FAIL(c, WC_ECAPACITY, "%zu pairs exceed the output capacity %zu", np, cap);
  if (np) {
    WC_CUDA(c, cudaMemcpyAsync(out, c->d_corr_out, np * sizeof(wc_corr_idx), cudaMemcpyDeviceToHost, st));
    if (first_is_target) WC_CUDA(c, cudaMemcpyAsync(first_is_target, c->d_fit_out, np, cudaMemcpyDeviceToHost, st));
  }
  WC_CUDA(c, cudaStreamSynchronize(st));
  if (gpu_ms) {
    float ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    *gpu_ms = ms;
  }
  *n_out = np;
  return WC_OK;
}

extern "C" wc_status wc_knn6(wc_ctx* c, const double* query6, size_t nq, const double* target6, size_t nt, int k,
                             int32_t* out_idx, double* out_dist2) {
  if (!c || !query6 || !target6 || !out_idx || !out_dist2) return WC_EINVAL;
  if (k < 1 || k > KMAX) WC_FAIL(c, WC_EINVAL, "k must be 1..%d", KMAX);
  if (nq > (size_t)c->prm.max_surfels || nt > (size_t)c->prm.max_surfels) WC_FAIL(c, WC_ECAPACITY, "too many vectors");
  // the reference reads k_indices[0..k) whatever the target count (undefined behaviour, knn_surfel_matcher.cc:60-62)
  if (nq && nt < (size_t)k) WC_FAIL(c, WC_ETOO_FEW_TARGETS, "%zu targets for a %d-nearest-neighbour search", nt, k);
  wc_status s = match_alloc(c);
  if (s) return s;
  cudaStream_t st = c->stream;
  // the feature buffers have room for FSTR doubles per record: pack the 6-vectors at the front
  WC_CUDA(c, cudaMemcpyAsync(c->d_qfeat, query6, nq * 48, cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemcpyAsync(c->d_tfeat, target6, nt * 48, cudaMemcpyHostToDevice, st));
  if (nq && nt)
    { ++c->n_launches; knn6_bruteforce<KMAX><<<(unsigned)((nq + TILE - 1) / TILE), TILE, 0, st>>>(c->d_qfeat, 6, (int)nq, c->d_tfeat, 6, (int)nt, k,
                                                                         c->d_knn_idx, c->d_knn_d2); }
  WC_CUDA(c, cudaGetLastError());
  WC_CUDA(c, cudaMemcpyAsync(out_idx, c->d_knn_idx, nq * k * 4, cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaMemcpyAsync(out_dist2, c->d_knn_d2, nq * k * 8, cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  return WC_OK;
}
