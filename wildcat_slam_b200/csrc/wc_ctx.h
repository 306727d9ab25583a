// Device context of the B200 window-odometry library (internal; the public surface is include/wildcat_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/wildcat_b200.h"

#define WC_NUM_SMS_FALLBACK 148

// ---- surfel extraction: fixed-point accumulation layout ------------------------------------------------
// One slot per (voxel, leaf cell, time bin).  All accumulators are 64-bit integers fed by native RED
// atomics, so the sums are exact and independent of the order the points arrive in (bitwise reproducible).
//   coordinates: relative to the leaf-cell centre, unit 2^-27 m  (|rel| <= 0.1 m -> 24 bits; products < 2^48)
//   time       : relative to the first point, unit 2^-36 s; bin = Q >> 31 (2^-5 s = 31.25 ms < cluster gap)
#define WC_COORD_SCALE 134217728.0   /* 2^27 */
#define WC_TIME_SCALE 68719476736.0  /* 2^36 */
#define WC_BIN_SHIFT 31
#define WC_MAX_BINS 4096
#define WC_VOX_BIAS 16384 /* voxel coordinates relative to the first point's voxel, 15 bits per axis */
#define WC_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull

// Run / cell records ("slots") live in four planes, so that a warp whose lanes hold consecutive slots writes whole
// 32-byte sectors side by side (a 128-byte array-of-structures record per lane costs one memory transaction per lane
// and per 16 bytes: measured 38 of 80 us in K1) and so that the cell merge reads the keys as one dense array:
//   key[s]                    [vx:15][vy:15][vz:15][leaf:6][bin:12]
//   p0[s] = {n, imin, imax, 0 | st, s[0]}        n points; first / last point of the cell (indices into the sweep:
//                                                timestamps are non-decreasing in the index, so these are its earliest /
//                                                latest point; the emit stage reads their exact fp64 times);
//                                                st = sum of (Q - bin<<31); s = sum rel
//   p1[s] = {s[1], s[2] | ss[0], ss[1]}          ss = xx xy xz yy yz zz
//   p2[s] = {ss[2], ss[3] | ss[4], ss[5]}
// 104 bytes per slot.
struct wc_slot_planes {
  unsigned long long* key;
  int4*               p0;  // two int4 per slot in each body plane
  int4*               p1;
  int4*               p2;
};
#define WC_SLOT_BYTES 104

struct wc_extract_status {
  int      err_time_order;
  int      err_range;     // voxel offset or bin out of key range
  int      err_capacity;  // hash table / slot / voxel-entry capacity
  int      n_slots;
  int      n_voxels;
  int      n_surfels;
  int      pad[2];
};

// packed correspondence record (one lidar residual block), 128 bytes
struct __align__(16) wc_corr_rec {
  double v1[3];  // rot1 * center1 (body -> world rotation only)
  double v2[3];
  double d0[3];  // binary: pos1 - pos2 ; unary: GetCenterInWorld(s1) - pos2
  double n[3];   // weight * normal
  double f1, f2;
  int    b1l, b2l;  // left sample index of each interval (b1l = -1 for unary factors)
  int    mode, pad;
};
static_assert(sizeof(wc_corr_rec) == 128, "corr record layout");

struct wc_imu_rec {
  wc_imu_state i1, i2, i3;
  double       ts[3];
  int          blk[3];
  int          mode;
};

struct wc_ctx {
  wc_params    prm;
  int          device;
  int          num_sms;
  cudaStream_t stream;
  cudaStream_t side[3];   // side streams for independent launches inside one call (fork / join by events)
  cudaEvent_t  ev_fork, ev_join[3];
  cudaEvent_t  ev[8];
  char         err[512];
  long long    n_launches;  // kernels launched so far (bench.py's gpu_launches)

  // ---- extraction
  void*               d_raw;      // wc_point48 staging (raw upload)
  void*               d_raw_next; // second staging buffer: the NEXT sweep, copied in by wc_points_prefetch while a pass runs
  cudaStream_t        copy_stream;
  cudaEvent_t         ev_prefetch;
  const void*         prefetch_src;  // host buffer of the prefetched sweep (nullptr: none pending)
  size_t              prefetch_n;
  const void*         defer_src;     // prefetch requested for the solve stage of the next window pass, not issued yet
  size_t              defer_n;
  float4*             d_xyz;      // resident points
  double*             d_time;
  size_t              n_pts;
  void*               d_htab;     // cell hash: 16-byte {key, published slot id} entries
  size_t              hcap;
  void*               d_slots;      // slot planes (wc_slot_planes over one allocation)
  size_t              slot_cap;
  unsigned long long* d_vkeys;    // voxel hash
  int*                d_vslot;
  size_t              vcap;
  int*                d_vox_count;  // per voxel: number of slots
  int*                d_vox_off;
  int2*               d_rec_info;   // per run record: {voxel id or -1 (merged away), rank among the voxel's slots}
  int*                d_rec_tpos;   // per run record: the cell-table entry it claimed, or -1
  unsigned long long* d_vox_key;
  int*                d_vox_hpos;   // voxel-table position of each voxel (for cleanup)
  int*                d_seg;        // slot ids grouped by voxel
  wc_extract_status*  d_xstat;
  wc_extract_status*  h_xstat;      // pinned
  wc_surfel*          d_surf_raw;   // emitted, unsorted
  wc_surfel*          d_surf;       // sorted
  unsigned long long* d_sort_hi;
  unsigned long long* d_sort_lo;
  unsigned int*       d_sort_idx;
  unsigned int*       d_sort_perm;  // surfel ids grouped by time bucket
  int*                d_bcnt;       // per time bucket: count / offset / cursor
  int*                d_boff;
  int*                d_bcur;
  wc_point_assign*    d_assign;
  size_t              n_surfels;
  int                 vox0[3];
  double              t_first, t_last;
  int                 want_assign;
  int                 last_slots, last_voxels;

  void*               d_sweep;    // wc_sweep_mem (filter / undistort staging)

  // ---- matcher
  double* d_qfeat;  // nq x 8: 6 features + timestamp + pad
  double* d_tfeat;
  int*    d_knn_idx;
  double* d_knn_d2;
  int*    d_gated;     // nq x k gated candidate lists (-1 terminated)
  int*    d_acc;       // accepted candidate per query
  int*    d_acc2;
  int*    d_flag;
  int*    h_flag;      // pinned
  wc_surfel* d_msurf_q;
  wc_surfel* d_msurf_t;
  wc_corr_idx* d_corr_out;
  unsigned char* d_fit_out;
  int*    d_scan_tmp;
  void*   d_grid;      // GridBufs (host copy of the device pointers)
  double* d_part_d;    // partial top-k lists of the tiled exhaustive kNN stage
  int*    d_part_i;
  int     match_query_first;  // last match produced at least one pair whose FIRST surfel is the query
  long long knn_grid_min;  // target count from which the uniform-grid kNN replaces the brute-force scan

  // ---- window solve
  wc_surfel*       d_sld;
  wc_surfel*       d_fix;
  wc_corr_idx*     d_sld_corr;
  wc_corr_idx*     d_fix_corr;
  wc_imu_state*    d_imu;
  wc_sample_state* d_samples;
  wc_corr_rec*     d_rec;
  wc_imu_rec*      d_imu_rec;
  size_t           n_sld, n_fix, n_sld_corr, n_fix_corr, n_imu, K, n_rec, n_imu_rec;
  size_t           n_sld_prev;  // sliding-window surfels of earlier sweeps in d_sld[0 .. n_sld_prev): the pass appends behind them
  wc_surfel*       d_fix_tmp;   // staging of wc_window_shrink
  double*          d_x;       // 12K current point
  double*          d_xc;      // candidate
  double*          d_x0;      // uploaded start
  double*          d_H;       // (12K)^2 full symmetric, accumulated
  double*          d_g;       // 12K
  double*          d_cost;    // [2]: lidar+imu cost at linearisation point / candidate
  double*          d_Hs;      // scaled reduced system + workspace
  double*          d_work;
  void*            d_lm;      // LM state block (device)
  void*            h_lm;      // pinned mirror
  int*             d_status;  // assemble errors
  void*            d_spline;  // wc_spline_mem
  int              n_imu_blocks, first_imu_block;
  int              lm_batch;  // LM iterations enqueued between host checks of the termination flag
  int              last_lm_iters;  // iterations of the previous solve (how far to enqueue ahead)

  // ---- multi-GPU exchange
  int     rank, world;
  double* d_xchg;         // my exchange buffer (IPC-exported)
  size_t  xchg_bytes;
  double* peer_xchg[8];   // peer exchange buffers (own entry = d_xchg)
  int     comm_ready;
  unsigned long long comm_epoch;
  unsigned long long gather_epoch;
  int                shard_upload;     // wc_comm_shard_upload: sweep uploads are collective, each rank copies its slab only
  unsigned long long raw_epoch;        // sharded sweep upload: one epoch per upload / prefetch (parity selects the raw area)
  unsigned long long prefetch_epoch;   // epoch taken by the pending prefetch
  int*    d_comm_err;
};

#define WC_CUDA(ctx, call)                                                                        \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s -> %s", __FILE__, __LINE__, #call,       \
               cudaGetErrorString(e_));                                                           \
      return WC_ECUDA;                                                                            \
    }                                                                                             \
  } while (0)

#define WC_FAIL(ctx, code, ...)                                  \
  do {                                                           \
    snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);       \
    return (code);                                               \
  } while (0)

static inline size_t wc_next_pow2(size_t v) {
  size_t p = 1;
  while (p < v) p <<= 1;
  return p;
}
