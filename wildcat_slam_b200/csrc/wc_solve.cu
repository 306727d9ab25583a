// Window solve on the device — replaces LidarOdometry::Build{SldWin,FixWin}LidarResiduals + BuildImuResiduals +
// ceres::Solve (src/odometry/lidar_odometry.cc:254-363,541-561) with the cost functors of src/odometry/cost_functor.h.
//
//   K4 corr_pack        factor construction (cost_functor.h:17-26,102-114): eigen of Sigma1w+Sigma2w -> weight*normal,
//                       v_i = R_i c_i, bracketing sample intervals (upper_bound, lidar_odometry.cc:258-268,303-307),
//                       interpolation factors; records bucketed by interval pair, stored as 16 SoA columns
//   K5 lidar_linearize  residual + Cauchy corrector + analytic 1x24 Jacobian row (cost_functor.h:28-59,116-179, Q1
//                       switch) per correspondence; J^T J / J^T r of a 256-record tile by a register-tiled SYRK out of
//                       shared memory; per-bucket 24x24 blocks flushed to the dense normal equations with fp64 RED
//      imu_linearize    ImuFactor residuals/Jacobians (cost_functor.h:264-472), one warp per IMU triplet
//   K6 (fused into K5)  the candidate-cost pass of Ceres' step evaluation is the same launch that linearises at the
//                       candidate: an accepted step re-uses that J^T J, a rejected one discards it
//   K7 lm_*             Ceres TrustRegionMinimizer + LevenbergMarquardtStrategy: Jacobi scaling,
//                       damping, dense Cholesky in shared memory, step, accept/reject, radius update, termination —
//                       one CTA, state resident on the device
#include <float.h>
#include <type_traits>
#include <stdlib.h>

#include "wc_ctx.h"
#include <cooperative_groups.h>

#include "wc_device_math.cuh"
#include "wc_tma.cuh"

using namespace wcd;
namespace cg = cooperative_groups;

wc_status wc_comm_allreduce(wc_ctx* c, int at_candidate);  // wc_comm.cu; no-op when world == 1
wc_status wc_comm_check(wc_ctx* c);
wc_status wc_comm_begin_solve(wc_ctx* c);
void      wc_comm_partial_views(wc_ctx* c, double** H, double** g, double** cost);

namespace {

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may start while
// its predecessor on the stream is still running; pdl_wait() blocks until the predecessor grid has completed and its
// writes are visible, pdl_trigger() lets the successor be scheduled early.  Both are no-ops in a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int REC_COLS = 16;
constexpr int LT       = 128;       // linearize tile = threads per CTA (43 KB of shared memory: 4-5 CTAs per SM)
constexpr int JR       = 28;        // augmented row count: 24 Jacobian columns, residual, 3 pad
constexpr int JS       = LT + 1;    // padded row stride (doubles)
constexpr int NGRP     = LT / 32;
constexpr int LIN_SMEM_BASE = ((JR * JS + NGRP * 28 * 16) * 8 + 127) / 128 * 128;  // Jacobian tile + flush staging
constexpr int LIN_SMEM      = LIN_SMEM_BASE + 2 * REC_COLS * LT * 8;                   // + two record stages (TMA), fp64 records
constexpr int LIN_SMEM_R32  = LIN_SMEM_BASE + 2 * REC_COLS * LT * 4;                   // fp32 records

struct LMState {
  // control
  int    cur;            // which of the two normal-equation buffers belongs to the current point x
  int    done;
  int    termination;
  int    iteration;
  int    step_valid;
  int    last_successful;
  int    reuse_diagonal;
  int    num_consecutive_invalid;
  int    num_successful, num_unsuccessful, num_linearizations;
  int    D;              // reduced dimension
  int    err;            // assembly / evaluation errors (wc_status)
  int    pending;        // a candidate has been evaluated and awaits lm_decide
  double radius, decrease_factor;
  double x_cost, x_norm, grad_max, model_cost_change, step_norm, initial_cost;
  double iter_cost[WC_MAX_ITER_LOG];
  double iter_radius[WC_MAX_ITER_LOG];
  signed char iter_accepted[WC_MAX_ITER_LOG];
};

struct SolveBufs {
  double* H[2];
  double* g[2];
  double* cost[2];  // each one double
  double* x;        // current point (N)
  double* xc;       // candidate
  double* scale;    // D
  double* diag;     // D
  double* step;     // D
  double* A;        // D*D workspace when it does not fit shared memory
  int*    act;      // wide systems: reduced columns some factor touches (H_cc != 0 at the first linearisation); act[D] = count
  LMState* st;
  int     N;        // 12 K
  int     fix_first;
};

__device__ __forceinline__ int upper_bound_ts(const double* __restrict__ ts, int K, double t) {
  int lo = 0, hi = K;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t < ts[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------ K4
struct PackArgs {
  const wc_surfel*   sld;
  const wc_surfel*   fix;
  const wc_corr_idx* sld_corr;
  const wc_corr_idx* fix_corr;
  int                n_sld, n_fix, n_sld_corr, n_fix_corr;
  int                c0, c1;  // this rank's slice of the concatenated correspondence list
  const double*      ts;
  int                K;
  double             weight_floor;
  double*            tmp;     // REC_COLS x stride, unsorted
  int                stride;
  int*               bucket;  // per record
  int*               hist;
  int*               err;   // pack errors (wc_status), separate from the LM state the solve re-initialises
};

__global__ void corr_pack(PackArgs a) {
  const int i = a.c0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.c1) return;
  const bool       unary = i >= a.n_sld_corr;
  const wc_corr_idx ci   = unary ? a.fix_corr[i - a.n_sld_corr] : a.sld_corr[i];
  const int        n1    = unary ? a.n_fix : a.n_sld;
  if (ci.s1 < 0 || ci.s1 >= n1 || ci.s2 < 0 || ci.s2 >= a.n_sld) {
    *a.err = WC_EINVAL;
    return;
  }
  const wc_surfel& s1 = unary ? a.fix[ci.s1] : a.sld[ci.s1];
  const wc_surfel& s2 = a.sld[ci.s2];
  if (!(s1.timestamp < s2.timestamp)) *a.err = WC_EINVAL_TIME_ORDER;  // CHECK_LT lidar_odometry.cc:256,301
  const Q4 q1 = ldq(s1.rot), q2 = ldq(s2.rot);
  const M3 R1 = ToMatrix(q1), R2 = ToMatrix(q2);
  // GetCovarianceInWorld (surfel.h:89-91) of both, summed; weight and direction (cost_functor.h:22-25,110-113)
  const M3 cw = (R1 * ld33(s1.covariance)) * transpose(R1) + (R2 * ld33(s2.covariance)) * transpose(R2);
  double   ev[3];
  M3       V;
  SymEig3(cw.m[0][0], cw.m[1][0], cw.m[2][0], cw.m[1][1], cw.m[2][1], cw.m[2][2], ev, V);  // lower triangle like Eigen
  const double w  = 1.0 / sqrt(a.weight_floor + ev[0]);
  const V3     wn = w * col(V, 0);
  const V3     v1 = q1 * ld3(s1.center), v2 = q2 * ld3(s2.center);
  const V3     p1 = ld3(s1.pos), p2 = ld3(s2.pos);
  const int    sp2r = upper_bound_ts(a.ts, a.K, s2.timestamp);
  int          b1l = -1, b2l = sp2r - 1, mode = 0;
  double       f1 = 0.0, f2 = 0.0;
  bool         ok = sp2r > 0 && sp2r < a.K;  // CHECKs lidar_odometry.cc:265-266,304-305
  if (ok) f2 = (s2.timestamp - a.ts[b2l]) / (a.ts[b2l + 1] - a.ts[b2l]);
  if (!unary) {
    const int sp1r = upper_bound_ts(a.ts, a.K, s1.timestamp);
    ok             = ok && sp1r > 0 && sp1r < a.K;  // :259-260
    b1l            = sp1r - 1;
    if (ok) {
      f1   = (s1.timestamp - a.ts[b1l]) / (a.ts[b1l + 1] - a.ts[b1l]);
      mode = (a.ts[b1l + 1] < a.ts[b2l]) ? 0 : ((b1l + 1 == b2l) ? 1 : 2);  // :271,280
    }
  }
  if (!ok) {
    *a.err = WC_EOUT_OF_SPAN;
    return;
  }
  const int    r  = i - a.c0;
  double*      o  = a.tmp + r;
  const size_t S  = (size_t)a.stride;
  const V3     d0 = unary ? (v1 + p1) - p2 : p1 - p2;
  const V3     w1 = unary ? mk(0, 0, 0) : v1;
  o[0 * S] = w1.x, o[1 * S] = w1.y, o[2 * S] = w1.z;
  o[3 * S] = v2.x, o[4 * S] = v2.y, o[5 * S] = v2.z;
  o[6 * S] = d0.x, o[7 * S] = d0.y, o[8 * S] = d0.z;
  o[9 * S] = wn.x, o[10 * S] = wn.y, o[11 * S] = wn.z;
  o[12 * S] = f1, o[13 * S] = f2;
  const int bk = (b1l + 1) * a.K + b2l;
  o[14 * S]    = __longlong_as_double(((long long)(unsigned)b1l << 32) | (unsigned)b2l);
  o[15 * S]    = __longlong_as_double(((long long)(unsigned)mode << 32) | (unsigned)bk);
  a.bucket[r]  = bk;
  atomicAdd(&a.hist[bk], 1);
}

__global__ void __launch_bounds__(1024) bucket_scan(const int* __restrict__ hist, int nb, int* __restrict__ off) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? hist[i] : 0;
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
    if (i < nb) off[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

__global__ void bucket_scatter(const double* __restrict__ tmp, const int* __restrict__ bucket, int n, int stride,
                               const int* __restrict__ off, int* __restrict__ cursor, double* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int bk  = bucket[i];
  const int pos = off[bk] + atomicAdd(&cursor[bk], 1);
#pragma unroll
  for (int c = 0; c < REC_COLS; ++c) rec[(size_t)c * stride + pos] = tmp[(size_t)c * stride + i];
}

// ------------------------------------------------------------------------------------------------ K5
template <int PREC> struct TJSel { using type = double; };
template <> struct TJSel<WC_PREC_F32> { using type = float; };

struct LinArgs {
  const double* rec;
  const float*  rec32;  // fp32 copy of the sorted records (WC_PREC_MIXED / WC_PREC_F32)
  int           stride, n_rec;
  SolveBufs     B;
  int           at_candidate;  // 0: linearise at x into buffer cur; 1: at xc into buffer 1-cur
  int           jac_mode;
  double        cauchy_b, cauchy_c;
};

__device__ __forceinline__ void lidar_eval(const double* __restrict__ rec, size_t S, int i, const double* __restrict__ x,
                                           int jac_mode, double cb, double cc, double J[24], double& r_out, double& cost_out,
                                           int& b1l, int& b2l, int& bk) {
  const double* o  = rec + i;
  const V3      v1 = mk(o[0 * S], o[1 * S], o[2 * S]), v2 = mk(o[3 * S], o[4 * S], o[5 * S]);
  const V3      d0 = mk(o[6 * S], o[7 * S], o[8 * S]), wn = mk(o[9 * S], o[10 * S], o[11 * S]);
  const double  f1 = o[12 * S], f2 = o[13 * S];
  const long long w14 = __double_as_longlong(o[14 * S]), w15 = __double_as_longlong(o[15 * S]);
  b1l = (int)(w14 >> 32), b2l = (int)(w14 & 0xffffffffll);
  const int mode = (int)(w15 >> 32);
  bk             = (int)(w15 & 0xffffffffll);
  const double* x2l = x + 12 * b2l;
  const V3 r2 = (1 - f2) * ld3(x2l) + f2 * ld3(x2l + 12);
  const V3 t2 = (1 - f2) * ld3(x2l + 3) + f2 * ld3(x2l + 15);
  const Q4 E2 = Exp(r2);
  V3       r1 = mk(0, 0, 0), t1 = mk(0, 0, 0);
  Q4       E1 = Q4{1, 0, 0, 0};
  const bool unary = b1l < 0;
  if (!unary) {
    const double* x1l = x + 12 * b1l;
    r1 = (1 - f1) * ld3(x1l) + f1 * ld3(x1l + 12);
    t1 = (1 - f1) * ld3(x1l + 3) + f1 * ld3(x1l + 15);
    E1 = Exp(r1);
  }
  // residual (cost_functor.h:39,140); wn carries the weight
  const V3     e   = (E1 * v1 + t1 + d0) - (E2 * v2 + t2);
  const double r   = dot(wn, e);
  const double s   = r * r;
  const double sum = 1.0 + s * cc, inv = 1.0 / sum;  // ceres::CauchyLoss
  cost_out         = 0.5 * cb * log(sum);
  const double sr  = sqrt(fmax(DBL_MIN, inv));       // Corrector: rho'' < 0 => scale by sqrt(rho')
  r_out            = r * sr;
  // jacobian_s2 (:42-45,162-165), jacobian_s1 (:147-150)
  const V3 a2 = vTm(vTm(vTm(wn, ToMatrix(E2)), Hat(v2)), Jr(r2));
  double   g1l = 1 - f1, g1r = f1;
  const double g2l = 1 - f2, g2r = f2;
  // Q1: in the aliased modes the later '=' overwrites the s1 part (cost_functor.h:152-175 + 215-229)
  if (jac_mode == WC_JAC_REFERENCE_OVERWRITE) {
    if (mode == 1) g1r = 0.0;
    if (mode == 2) g1l = 0.0, g1r = 0.0;
  }
  if (unary) {
#pragma unroll
    for (int k = 0; k < 12; ++k) J[k] = 0.0;
  } else {
    const V3 a1 = vTm(vTm(vTm(-wn, ToMatrix(E1)), Hat(v1)), Jr(r1));
    J[0] = sr * g1l * a1.x, J[1] = sr * g1l * a1.y, J[2] = sr * g1l * a1.z;
    J[3] = sr * g1l * wn.x, J[4] = sr * g1l * wn.y, J[5] = sr * g1l * wn.z;
    J[6] = sr * g1r * a1.x, J[7] = sr * g1r * a1.y, J[8] = sr * g1r * a1.z;
    J[9] = sr * g1r * wn.x, J[10] = sr * g1r * wn.y, J[11] = sr * g1r * wn.z;
  }
  J[12] = sr * g2l * a2.x, J[13] = sr * g2l * a2.y, J[14] = sr * g2l * a2.z;
  J[15] = -sr * g2l * wn.x, J[16] = -sr * g2l * wn.y, J[17] = -sr * g2l * wn.z;
  J[18] = sr * g2r * a2.x, J[19] = sr * g2r * a2.y, J[20] = sr * g2r * a2.z;
  J[21] = -sr * g2r * wn.x, J[22] = -sr * g2r * wn.y, J[23] = -sr * g2r * wn.z;
}


// ---- fp32 evaluation of one lidar factor (WC_PREC_MIXED / WC_PREC_F32): same formulas as lidar_eval, single precision.
// The rotation matrix comes from Rodrigues' formula instead of the quaternion detour; Jr(r) = A I + B a a^T - C Hat(a).
struct F3 {
  float x, y, z;
};
__device__ __forceinline__ F3 f3(float x, float y, float z) { return F3{x, y, z}; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return F3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 operator*(float s, F3 a) { return F3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float fdot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 fcross(F3 a, F3 b) { return F3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
struct Rod {
  float A, B, C, D;  // sin(th)/th, (1 - sin(th)/th)/th^2, (1 - cos(th))/th^2, th^2
};
__device__ __forceinline__ Rod rodrigues(F3 r) {
  const float t2 = fdot(r, r);
  Rod         k;
  k.D = t2;
  if (t2 < 1e-8f) {
    k.A = 1.f - t2 * (1.f / 6.f), k.B = 1.f / 6.f - t2 * (1.f / 120.f), k.C = 0.5f - t2 * (1.f / 24.f);
  } else {
    const float th = sqrtf(t2);
    float       sn, cs;
    sincosf(th, &sn, &cs);
    k.A = sn / th, k.B = (1.f - k.A) / t2, k.C = (1.f - cs) / t2;
  }
  return k;
}
// Exp(r) v = v + A (r x v) + C r x (r x v);  Exp(r)^T u = Exp(-r) u
__device__ __forceinline__ F3 rot_apply(const Rod& k, F3 r, F3 v, float sign) {
  const F3 rv = fcross(r, v);
  return v + (sign * k.A) * rv + k.C * fcross(r, rv);
}
// w^T Jr(r) = A w + B (w . r) r - C (w x r)   (Jr(r) = Jl(-r), utils.h:35-49; A, B, C per unit of |r|^2 absorbed)
__device__ __forceinline__ F3 row_times_Jr(const Rod& k, F3 r, F3 w) { return k.A * w + (k.B * fdot(w, r)) * r - k.C * fcross(w, r); }

template <typename TJ>
__device__ __forceinline__ void lidar_eval32(const float* __restrict__ rec, size_t S, int i, const double* __restrict__ x, int jac_mode,
                                             float cb, float cc, TJ J[24], TJ& r_out, double& cost_out, int& b1l, int& b2l, int& bk) {
  const float* o  = rec + i;
  const F3     v1 = f3(o[0 * S], o[1 * S], o[2 * S]), v2 = f3(o[3 * S], o[4 * S], o[5 * S]);
  const F3     d0 = f3(o[6 * S], o[7 * S], o[8 * S]), wn = f3(o[9 * S], o[10 * S], o[11 * S]);
  const float  f1 = o[12 * S], f2 = o[13 * S];
  const int    w14 = __float_as_int(o[14 * S]), w15 = __float_as_int(o[15 * S]);
  b1l = w14 >> 16, b2l = w14 & 0xffff;
  const int mode = w15 >> 24;
  bk             = w15 & 0xffffff;
  const double* x2l = x + 12 * b2l;
  const F3 r2 = (1.f - f2) * f3((float)x2l[0], (float)x2l[1], (float)x2l[2]) + f2 * f3((float)x2l[12], (float)x2l[13], (float)x2l[14]);
  const F3 t2 = (1.f - f2) * f3((float)x2l[3], (float)x2l[4], (float)x2l[5]) + f2 * f3((float)x2l[15], (float)x2l[16], (float)x2l[17]);
  const Rod k2 = rodrigues(r2);
  F3        r1 = f3(0, 0, 0), t1 = f3(0, 0, 0);
  Rod       k1 = Rod{1.f, 1.f / 6.f, 0.5f, 0.f};
  const bool unary = b1l < 0;
  if (!unary) {
    const double* x1l = x + 12 * b1l;
    r1 = (1.f - f1) * f3((float)x1l[0], (float)x1l[1], (float)x1l[2]) + f1 * f3((float)x1l[12], (float)x1l[13], (float)x1l[14]);
    t1 = (1.f - f1) * f3((float)x1l[3], (float)x1l[4], (float)x1l[5]) + f1 * f3((float)x1l[15], (float)x1l[16], (float)x1l[17]);
    k1 = rodrigues(r1);
  }
  const F3    e   = (rot_apply(k1, r1, v1, 1.f) + t1 + d0) - (rot_apply(k2, r2, v2, 1.f) + t2);
  const float r   = fdot(wn, e);
  const float s   = r * r;
  const float sum = 1.f + s * cc, inv = 1.f / sum;
  cost_out        = 0.5 * (double)(cb * logf(sum));
  const float sr  = sqrtf(fmaxf(FLT_MIN, inv));
  r_out           = (TJ)(r * sr);
  // a2 = wn^T R2 Hat(v2) Jr(r2) = ((R2^T wn) x v2)^T Jr(r2);  a1 likewise with -wn
  const F3 a2 = row_times_Jr(k2, r2, fcross(rot_apply(k2, r2, wn, -1.f), v2));
  float    g1l = 1.f - f1, g1r = f1;
  const float g2l = 1.f - f2, g2r = f2;
  if (jac_mode == WC_JAC_REFERENCE_OVERWRITE) {
    if (mode == 1) g1r = 0.f;
    if (mode == 2) g1l = 0.f, g1r = 0.f;
  }
  if (unary) {
#pragma unroll
    for (int k = 0; k < 12; ++k) J[k] = (TJ)0;
  } else {
    const F3 a1 = row_times_Jr(k1, r1, fcross(rot_apply(k1, r1, -1.f * wn, -1.f), v1));
    J[0] = (TJ)(sr * g1l * a1.x), J[1] = (TJ)(sr * g1l * a1.y), J[2] = (TJ)(sr * g1l * a1.z);
    J[3] = (TJ)(sr * g1l * wn.x), J[4] = (TJ)(sr * g1l * wn.y), J[5] = (TJ)(sr * g1l * wn.z);
    J[6] = (TJ)(sr * g1r * a1.x), J[7] = (TJ)(sr * g1r * a1.y), J[8] = (TJ)(sr * g1r * a1.z);
    J[9] = (TJ)(sr * g1r * wn.x), J[10] = (TJ)(sr * g1r * wn.y), J[11] = (TJ)(sr * g1r * wn.z);
  }
  J[12] = (TJ)(sr * g2l * a2.x), J[13] = (TJ)(sr * g2l * a2.y), J[14] = (TJ)(sr * g2l * a2.z);
  J[15] = (TJ)(-sr * g2l * wn.x), J[16] = (TJ)(-sr * g2l * wn.y), J[17] = (TJ)(-sr * g2l * wn.z);
  J[18] = (TJ)(sr * g2r * a2.x), J[19] = (TJ)(sr * g2r * a2.y), J[20] = (TJ)(sr * g2r * a2.z);
  J[21] = (TJ)(-sr * g2r * wn.x), J[22] = (TJ)(-sr * g2r * wn.y), J[23] = (TJ)(-sr * g2r * wn.z);
}

// block (bi, bj), bi <= bj, of the 7x7 grid of 4x4 blocks, enumerated by lane 0..27
__device__ __forceinline__ void lane_block(int lane, int& bi, int& bj) {
  int l = lane;
  bi    = 0;
  while (l >= 7 - bi) l -= 7 - bi, ++bi;
  bj = bi + l;
}

template <int PREC, bool STAGED>
__device__ __forceinline__ void lidar_linearize_body(const LinArgs& a, double* sm, int cta, int ncta) {
  using TJ = typename TJSel<PREC>::type;   // element type of the tile's Jacobian rows and of the J^T J partial blocks
  TJ* Jt    = reinterpret_cast<TJ*>(sm);  // JR x JS
  TJ* stage = Jt + JR * JS;               // NGRP x 28 x 16 (flush staging; separate from Jt: a flush can happen mid-tile)
  __shared__ int sbk[LT];
  __shared__ int sb1[LT], sb2[LT];
  __shared__ int heads[LT];
  __shared__ int nheads;
  __shared__ double wcost[NGRP];

  LMState* st = a.B.st;
  if (st->done || (a.at_candidate && !st->step_valid)) return;
  const int     buf = a.at_candidate ? 1 - st->cur : st->cur;
  const double* x   = a.at_candidate ? a.B.xc : a.B.x;
  double*       H   = a.B.H[buf];
  double*       g   = a.B.g[buf];
  const int     N   = a.B.N;
  const size_t  S   = (size_t)a.stride;

  const int t = threadIdx.x, lane = t & 31, grp = t >> 5;
  int       bi, bj;
  lane_block(lane < 28 ? lane : 0, bi, bj);
  TJ acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = (TJ)0;
  double cost_local = 0.0;
  int    cur_b1 = -2, cur_b2 = -2, cur_bk = -1;

  const int ntiles = (a.n_rec + LT - 1) / LT;
  const int per    = (ntiles + ncta - 1) / ncta;
  const int tile0 = cta * per, tile1 = min(ntiles, tile0 + per);

  // Record tiles are staged in shared memory by the TMA unit (cp.async.bulk: one 1-D bulk copy per SoA column, 16 per
  // tile, all completing on one mbarrier), double buffered: the copy of tile k + 1 is in flight while tile k is evaluated
  // and its J^T J accumulated.  Full tiles are always copied (the column stride is padded, rows past n_rec are ignored).
  using TR = typename std::conditional<PREC == WC_PREC_F64, double, float>::type;
  TR* srec = reinterpret_cast<TR*>(reinterpret_cast<unsigned char*>(sm) + LIN_SMEM_BASE);  // [2][REC_COLS][LT]
  __shared__ __align__(8) uint64_t full_bar[2];
  const TR* grec = PREC == WC_PREC_F64 ? reinterpret_cast<const TR*>(a.rec) : reinterpret_cast<const TR*>(a.rec32);
  auto issue = [&](int tile, int stg) {  // threads 0 .. REC_COLS - 1: one column each; thread 0 also arms the barrier
    if (t == 0) wctma::mbar_expect_tx(&full_bar[stg], (unsigned)(REC_COLS * LT * sizeof(TR)));
    wctma::bulk_g2s(srec + (stg * REC_COLS + t) * LT, grec + (size_t)t * S + (size_t)tile * LT, (unsigned)(LT * sizeof(TR)), &full_bar[stg]);
  };
  if constexpr (STAGED) {
    if (t == 0) {
      wctma::mbar_init(&full_bar[0], 1);
      wctma::mbar_init(&full_bar[1], 1);
      wctma::mbar_fence_init();
    }
    __syncthreads();
    if (t < REC_COLS && tile0 < tile1) issue(tile0, 0);
  }
  // flush the per-thread 4x4 partial blocks of the current bucket into the dense normal equations
  auto flush = [&]() {
    __syncthreads();
    if (lane < 28)
#pragma unroll
      for (int k = 0; k < 16; ++k) stage[(grp * 28 + lane) * 16 + k] = acc[k], acc[k] = (TJ)0;
    __syncthreads();
    if (cur_bk >= 0) {
      for (int o = t; o < 28 * 16; o += LT) {
        const int blk = o >> 4, k = o & 15;
        int       pbi, pbj;
        lane_block(blk, pbi, pbj);
        const int p = 4 * pbi + (k >> 2), q = 4 * pbj + (k & 3);
        if (p > q || p >= 24 || q > 24) continue;
        double v = 0.0;
#pragma unroll
        for (int gi = 0; gi < NGRP; ++gi) v += (double)stage[(gi * 28 + blk) * 16 + k];
        if (v == 0.0) continue;
        const int sa = p / 6, gpb = (sa < 2 ? cur_b1 + sa : cur_b2 + sa - 2);
        if (sa < 2 && cur_b1 < 0) continue;  // unary factor: no s1 blocks
        const int gp = 12 * gpb + p % 6;
        if (q == 24) {
          atomicAdd(&g[gp], v);
        } else {
          const int sb = q / 6, gqb = (sb < 2 ? cur_b1 + sb : cur_b2 + sb - 2);
          if (sb < 2 && cur_b1 < 0) continue;
          const int gq = 12 * gqb + q % 6;
          // lower triangle only (the LM step reads nothing else; wc_window_evaluate mirrors it for the caller).  Blocks of
          // two factors' shared sample states overlap, so an entry and its mirror image can both occur: they are summed.
          // (two different local columns on the same unknown — aliased intervals in WC_JAC_EXACT mode — meet on the
          // diagonal, where the old mirrored pair of additions counted them twice, as (J_p + J_q)^2 requires)
          atomicAdd(&H[(size_t)max(gp, gq) * N + min(gp, gq)], (gp == gq && p != q) ? 2.0 * v : v);
        }
      }
    }
    __syncthreads();
  };

  for (int tile = tile0; tile < tile1; ++tile) {
    const int  i     = tile * LT + t;
    const bool valid = i < a.n_rec;
    const int  stg   = (tile - tile0) & 1;
    // the other stage was last read two barriers ago (every thread is past the previous tile's barriers): refill it
    const TR* trec = grec + (size_t)tile * LT;  // !STAGED: straight from global memory (one tile per CTA: nothing to overlap)
    size_t    tS   = S;
    if constexpr (STAGED) {
      if (t < REC_COLS && tile + 1 < tile1) issue(tile + 1, stg ^ 1);
      wctma::mbar_wait(&full_bar[stg], (unsigned)(((tile - tile0) >> 1) & 1));
      trec = srec + stg * REC_COLS * LT, tS = (size_t)LT;
    }
    TJ         J[24], r = (TJ)0;
    double     cst = 0.0;
    int        b1l = -2, b2l = -2, bk = -1;
    if (valid) {
      if constexpr (PREC == WC_PREC_F64) lidar_eval(trec, tS, t, x, a.jac_mode, a.cauchy_b, a.cauchy_c, J, r, cst, b1l, b2l, bk);
      else lidar_eval32<TJ>(trec, tS, t, x, a.jac_mode, (float)a.cauchy_b, (float)a.cauchy_c, J, r, cst, b1l, b2l, bk);
    } else {
#pragma unroll
      for (int k = 0; k < 24; ++k) J[k] = (TJ)0;
    }
    cost_local += cst;
    __syncthreads();  // previous tile's SYRK reads are done
#pragma unroll
    for (int k = 0; k < 24; ++k) Jt[k * JS + t] = J[k];
    Jt[24 * JS + t] = r;
    Jt[25 * JS + t] = (TJ)0, Jt[26 * JS + t] = (TJ)0, Jt[27 * JS + t] = (TJ)0;
    sbk[t] = bk, sb1[t] = b1l, sb2[t] = b2l;
    if (t == 0) nheads = 0;
    __syncthreads();
    if (valid && (t == 0 || sbk[t - 1] != bk)) heads[atomicAdd(&nheads, 1)] = t;
    __syncthreads();
    const int nh = nheads;
    for (int h = 0; h < nh; ++h) {
      const int hd = heads[h];
      const int b  = sbk[hd];
      if (b != cur_bk) {
        if (cur_bk >= 0) flush();
        cur_bk = b, cur_b1 = sb1[hd], cur_b2 = sb2[hd];
      }
      if (lane < 28) {
        const int c0 = grp * 32;
        for (int c = c0; c < c0 + 32; ++c) {
          if (sbk[c] != b) continue;
          TJ ra[4], rb[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) ra[k] = Jt[(4 * bi + k) * JS + c], rb[k] = Jt[(4 * bj + k) * JS + c];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[4 * u + v] += ra[u] * rb[v];
        }
      }
    }
  }
  if (cur_bk >= 0) flush();
  // cost: warp reduce, then one RED per CTA
  for (int d = 16; d > 0; d >>= 1) cost_local += __shfl_down_sync(0xffffffffu, cost_local, d);
  if (lane == 0) wcost[grp] = cost_local;
  __syncthreads();
  if (t == 0) {
    double c = 0.0;
    for (int k = 0; k < NGRP; ++k) c += wcost[k];
    atomicAdd(a.B.cost[buf], c);
  }
}

// ---- IMU factors -----------------------------------------------------------------------------------------------
struct ImuArgs {
  const wc_imu_state* imu;
  int                 n_imu;
  const double*       ts;
  int                 K;
  SolveBufs           B;
  int                 at_candidate;
  double              wg, wa, wbg, wba, dt;
  double              grav[3];
};

struct StateCorr {
  V3 r, t, bg, ba;
};
__device__ __forceinline__ void state_corr(const double* xl, const double* xr, double f, StateCorr& c) {
  c.r  = (1 - f) * ld3(xl) + f * ld3(xr);
  c.t  = (1 - f) * ld3(xl + 3) + f * ld3(xr + 3);
  c.bg = (1 - f) * ld3(xl + 6) + f * ld3(xr + 6);
  c.ba = (1 - f) * ld3(xl + 9) + f * ld3(xr + 9);
}
// F (cost_functor.h:446-448)
__device__ __forceinline__ M3 imu_F(const Q4& L, const Q4& R, const V3& r) {
  return (Jr_inv(Log((L * Exp(r)) * R)) * ToMatrix(conj(R))) * Jr(r);
}
__device__ __forceinline__ void add_block(double* jac, int ld, int r0, int c0, const M3& m, double s) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) jac[(r0 + i) * ld + c0 + j] += s * m.m[i][j];
}

constexpr int IMU_WARPS = LT / 32;
constexpr int IMU_LD    = 37;  // 36 Jacobian columns + residual

// One warp per IMU triplet (BuildImuResiduals, lidar_odometry.cc:319-363).
__device__ __forceinline__ void imu_linearize_body(const ImuArgs& a, double* sm, int cta) {
  double (*sj)[12 * IMU_LD] = reinterpret_cast<double (*)[12 * IMU_LD]>(sm);
  __shared__ int sblk[IMU_WARPS][4];
  LMState* st = a.B.st;
  if (st->done || (a.at_candidate && !st->step_valid)) return;
  const int     buf  = a.at_candidate ? 1 - st->cur : st->cur;
  const double* x    = a.at_candidate ? a.B.xc : a.B.x;
  double*       H    = a.B.H[buf];
  double*       g    = a.B.g[buf];
  const int     N    = a.B.N;
  const int     lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int     i    = cta * IMU_WARPS + wid;
  if (i + 2 >= a.n_imu) return;
  const wc_imu_state &i1 = a.imu[i], &i2 = a.imu[i + 1], &i3 = a.imu[i + 2];
  if (i1.timestamp < a.ts[0] || i3.timestamp > a.ts[a.K - 1]) return;  // :324-329 (timestamps increase)
  double* jac = sj[wid];
  for (int k = lane; k < 12 * IMU_LD; k += 32) jac[k] = 0.0;
  __shared__ double smat[IMU_WARPS][4][9];  // F0, F1, A30, A39 of this warp's triplet
  __syncwarp();
  // The factor's independent pieces run on four lanes side by side (each repeats the cheap common part): lane 0 the
  // residuals, lane 1 F0, lane 2 F1, lane 3 A30 / A39 — the serial chain of one lane doing everything was the tail of the
  // whole linearisation launch.
  int    l[3] = {0, 0, 0};
  double f[3] = {0.0, 0.0, 0.0};
  if (lane < 4) {
    const int sp2  = upper_bound_ts(a.ts, a.K, i1.timestamp);
    const int mode = (sp2 == a.K - 1) ? 1 : 0;
    const int nblk = mode == 0 ? 3 : 2;
    const double tsv[3] = {a.ts[sp2 - 1], a.ts[sp2], mode == 0 ? a.ts[sp2 + 1] : DBL_MAX};
    const double* xs[3] = {x + 12 * (sp2 - 1), x + 12 * sp2, x + 12 * (mode == 0 ? sp2 + 1 : sp2)};
    // ComputeStateCorr for the three IMU states (cost_functor.h:358-400)
    StateCorr c[3];
    const double tt[3] = {i1.timestamp, i2.timestamp, i3.timestamp};
    bool         ok    = true;
    for (int s = 0; s < 3; ++s) {
      const double t = tt[s];
      if (mode == 0) {
        const bool in12 = t >= tsv[0] && t < tsv[1], in23 = t >= tsv[1] && t <= tsv[2];
        ok   = ok && (in12 || in23);
        l[s] = in12 ? 0 : 1;
      } else {
        ok   = ok && t >= tsv[0] && t <= tsv[1];
        l[s] = 0;
      }
      f[s] = (t - tsv[l[s]]) / (tsv[l[s] + 1] - tsv[l[s]]);
      state_corr(xs[l[s]], xs[l[s] + 1], f[s], c[s]);
    }
    const Q4 R1 = ldq(i1.rot), R2 = ldq(i2.rot);
    const Q4 E1R1 = Exp(c[0].r) * R1;
    if (lane == 0) {
      if (!ok) st->err = WC_EOUT_OF_SPAN;
      const V3 gyr_est = Log((conj(E1R1) * Exp(c[1].r)) * R2) / a.dt;
      const V3 acc_est = ((c[2].t + ld3(i3.pos)) + (c[0].t + ld3(i1.pos)) - 2.0 * (c[1].t + ld3(i2.pos))) / (a.dt * a.dt);
      const V3 rg  = a.wg * ((ld3(i1.gyr) + ld3(i2.gyr)) / 2.0 - gyr_est - c[0].bg);
      const V3 ra  = a.wa * (E1R1 * (ld3(i1.acc) - c[0].ba) - acc_est + ld3(a.grav));
      const V3 rbg = a.wbg * (c[0].bg - c[1].bg);
      const V3 rba = a.wba * (c[0].ba - c[1].ba);
      const double res[12] = {rg.x, rg.y, rg.z, ra.x, ra.y, ra.z, rbg.x, rbg.y, rbg.z, rba.x, rba.y, rba.z};
      for (int k = 0; k < 12; ++k) jac[k * IMU_LD + 36] = res[k];
      sblk[wid][0] = sp2 - 1, sblk[wid][1] = sp2, sblk[wid][2] = mode == 0 ? sp2 + 1 : -1, sblk[wid][3] = nblk;
      double cst = 0.0;
      for (int k = 0; k < 12; ++k) cst += 0.5 * res[k] * res[k];  // TrivialLoss
      atomicAdd(a.B.cost[buf], cst);
    } else if (lane == 1) {
      st33(smat[wid][0], imu_F(conj(R1), Exp(c[1].r) * R2, c[0].r));
    } else if (lane == 2) {
      st33(smat[wid][1], imu_F(conj(E1R1), R2, c[1].r));
    } else {
      st33(smat[wid][2], (ToMatrix(Exp(c[0].r)) * Hat(R1 * (ld3(i1.acc) - c[0].ba))) * Jr(c[0].r));
      st33(smat[wid][3], ToMatrix(E1R1));
    }
  }
  __syncwarp();
  if (lane == 0) {
    // jacobian_tau, tau1, tau2 (:301-321) dispatched to the bracketing blocks (:402-444)
    const M3     I   = eye3();
    const double idt = 1 / a.dt, idt2 = 1 / a.dt / a.dt;
    {
      const int    cl = 12 * l[0], cr = cl + 12;
      const double wl = 1 - f[0], wr = f[0];
      const M3     F0 = ld33(smat[wid][0]), A30 = ld33(smat[wid][2]), A39 = ld33(smat[wid][3]);
      for (int side = 0; side < 2; ++side) {
        const int    c0 = side ? cr : cl;
        const double w  = side ? wr : wl;
        add_block(jac, IMU_LD, 0, c0 + 0, F0, w * a.wg * idt);
        add_block(jac, IMU_LD, 0, c0 + 6, I, -w * a.wg);
        add_block(jac, IMU_LD, 3, c0 + 0, A30, -w * a.wa);
        add_block(jac, IMU_LD, 3, c0 + 3, I, -w * a.wa * idt2);
        add_block(jac, IMU_LD, 3, c0 + 9, A39, -w * a.wa);
        add_block(jac, IMU_LD, 6, c0 + 6, I, w * a.wbg);
        add_block(jac, IMU_LD, 9, c0 + 9, I, w * a.wba);
      }
    }
    {
      const int    cl = 12 * l[1], cr = cl + 12;
      const double wl = 1 - f[1], wr = f[1];
      const M3     F1 = ld33(smat[wid][1]);
      for (int side = 0; side < 2; ++side) {
        const int    c0 = side ? cr : cl;
        const double w  = side ? wr : wl;
        add_block(jac, IMU_LD, 0, c0 + 0, F1, -w * a.wg * idt);
        add_block(jac, IMU_LD, 0, c0 + 6, I, -w * a.wg);
        add_block(jac, IMU_LD, 3, c0 + 3, I, w * a.wa * (2 / a.dt / a.dt));
        add_block(jac, IMU_LD, 6, c0 + 6, I, -w * a.wbg);
        add_block(jac, IMU_LD, 9, c0 + 9, I, -w * a.wba);
      }
    }
    {
      const int    cl = 12 * l[2], cr = cl + 12;
      const double wl = 1 - f[2], wr = f[2];
      add_block(jac, IMU_LD, 3, cl + 3, I, -wl * a.wa * idt2);
      add_block(jac, IMU_LD, 3, cr + 3, I, -wr * a.wa * idt2);
    }
  }
  __syncwarp();
  const int nc = 12 * sblk[wid][3];
  // J^T J (both triangles) and J^T r of this factor, lanes over outputs
  for (int o = lane; o < nc * (nc + 1); o += 32) {
    const int p = o / (nc + 1), q = o % (nc + 1);
    const int qc = q == nc ? 36 : q;
    double    v  = 0.0;
#pragma unroll
    for (int r = 0; r < 12; ++r) v = fma(jac[r * IMU_LD + p], jac[r * IMU_LD + qc], v);
    if (v == 0.0) continue;
    const int gp = 12 * sblk[wid][p / 12] + p % 12;
    if (q == nc) {
      atomicAdd(&g[gp], v);
    } else {
      const int gq = 12 * sblk[wid][q / 12] + q % 12;
      if (gq <= gp) atomicAdd(&H[(size_t)gp * N + gq], v);  // lower triangle only
    }
  }
}

// IMU triplets on the first n_imu CTAs, lidar tiles on the rest: one launch per linearisation.  The IMU warps run long
// serial chains (Log / Exp / 12 x 36 Jacobians per triplet); scheduled first, they overlap with the lidar tiles
// instead of forming the kernel's tail.
template <int PREC, bool STAGED>
__global__ void __launch_bounds__(LT) window_linearize(LinArgs a, ImuArgs b, int n_lidar, int n_imu) {
  extern __shared__ __align__(16) double sm[];
  pdl_trigger();  // the LM step that follows may be scheduled now; it waits for this grid before reading anything
  pdl_wait();
  if ((int)blockIdx.x < n_imu) imu_linearize_body(b, sm, blockIdx.x);
  else lidar_linearize_body<PREC, STAGED>(a, sm, blockIdx.x - n_imu, n_lidar);
}

// one-time conversion of the sorted fp64 records to the 64-byte fp32 records of WC_PREC_MIXED / WC_PREC_F32
__global__ void rec_to_f32(const double* __restrict__ rec, int n, int stride, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t S = (size_t)stride;
#pragma unroll
  for (int c = 0; c < 14; ++c) out[c * S + i] = (float)rec[c * S + i];
  const long long w14 = __double_as_longlong(rec[14 * S + i]), w15 = __double_as_longlong(rec[15 * S + i]);
  const int b1l = (int)(w14 >> 32), b2l = (int)(w14 & 0xffffffffll), mode = (int)(w15 >> 32), bk = (int)(w15 & 0xffffffffll);
  out[14 * S + i] = __int_as_float((b1l << 16) | (b2l & 0xffff));
  out[15 * S + i] = __int_as_float((mode << 24) | bk);
}

// ------------------------------------------------------------------------------------------------ K7
constexpr int LMT = 512;
constexpr int CB  = 8;  // Cholesky block width

__device__ __forceinline__ int col_of(int i, int fix_first) { return fix_first ? (i < 3 ? i : (i < 6 ? -1 : i - 3)) : i; }
__device__ __forceinline__ int amb_of(int c, int fix_first) { return fix_first ? (c < 3 ? c : c + 3) : c; }

__device__ double block_sum(double v, double* red) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int k = 0; k < LMT / 32; ++k) s += red[k];  // fixed order: identical on every thread and every rank
  return s;
}
__device__ double block_max(double v, double* red) {
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, d));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int k = 0; k < LMT / 32; ++k) s = fmax(s, red[k]);
  return s;
}

__global__ void zero_buffers(SolveBufs B, int which /*0 cur, 1 other*/) {
  const LMState* st  = B.st;
  const int      buf = which ? 1 - st->cur : st->cur;
  const size_t   n   = (size_t)B.N * B.N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) B.H[buf][i] = 0.0;
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < B.N; i += blockDim.x) B.g[buf][i] = 0.0;
    if (threadIdx.x == 0) *B.cost[buf] = 0.0;
  }
}

// after the first linearisation: Jacobi scaling, gradient norm, initial state (TrustRegionMinimizer::Init)
__global__ void __launch_bounds__(LMT) lm_init(SolveBufs B, wc_solve_opts o) {
  __shared__ double red[LMT / 32];
  LMState*      st = B.st;
  const int     N = B.N, ff = B.fix_first;
  const double* H = B.H[st->cur];
  const double* g = B.g[st->cur];
  const int     D = ff ? N - 3 : N;
  double        gm = 0.0, xn = 0.0;
  for (int i = threadIdx.x; i < N; i += LMT) {
    const int c = col_of(i, ff);
    if (c >= 0) {
      B.scale[c] = 1.0 / (1.0 + sqrt(H[(size_t)i * N + i]));  // jacobian_scaling_, computed once
      gm         = fmax(gm, fabs(g[i]));
    }
    xn += B.x[i] * B.x[i];
  }
  gm = block_max(gm, red);
  xn = block_sum(xn, red);
  if (threadIdx.x == 0) {
    // Unknowns no factor touches (zero row / column of J^T J: e.g. the bias blocks of a lidar-only window) decouple: their
    // damped system is diag / radius * delta = 0, so delta = 0 exactly.  The wide-system step leaves them out.
    int na = 0;
    for (int c = 0; c < D; ++c)
      if (H[(size_t)amb_of(c, ff) * N + amb_of(c, ff)] != 0.0) B.act[na++] = c;
    B.act[D] = na;
    st->D = D;
    st->x_cost = st->initial_cost = *B.cost[st->cur];
    st->grad_max = gm, st->x_norm = sqrt(xn);
    st->radius = o.initial_trust_region_radius, st->decrease_factor = 2.0;
    st->reuse_diagonal = 0, st->num_consecutive_invalid = 0, st->iteration = 0, st->last_successful = 1;
    st->num_successful = st->num_unsuccessful = 0, st->num_linearizations = 1;
    st->termination = WC_TERM_NO_CONVERGENCE, st->done = 0, st->step_valid = 0, st->pending = 0;
    if (!isfinite(st->x_cost)) st->done = 1, st->termination = WC_TERM_FAILURE;
  }
}

// ---- decide: after the candidate evaluation — tolerances, accept / reject, trust-region update
__device__ void lm_decide_dev(const SolveBufs& B, LMState* st, const wc_solve_opts& o, double* red, int* s_accept) {
  const int t = threadIdx.x, N = B.N, ff = B.fix_first;
  const int it = st->iteration < WC_MAX_ITER_LOG ? st->iteration : WC_MAX_ITER_LOG - 1;
  if (t == 0) {
    *s_accept   = 0;
    st->pending = 0;
    if (!st->step_valid) {  // HandleInvalidStep
      st->num_unsuccessful += 1;
      st->iter_cost[it] = nan(""), st->iter_accepted[it] = 0;
      if (++st->num_consecutive_invalid >= 5) st->done = 1, st->termination = WC_TERM_FAILURE;
      else st->radius = st->radius / st->decrease_factor, st->decrease_factor *= 2.0, st->reuse_diagonal = 1;
    } else {
      st->num_consecutive_invalid = 0;
      st->num_linearizations += 1;
      double cand = *B.cost[1 - st->cur];
      if (!isfinite(cand)) cand = DBL_MAX;
      st->iter_cost[it]     = cand;
      st->iter_accepted[it] = 0;
      if (st->step_norm <= o.parameter_tolerance * (st->x_norm + o.parameter_tolerance)) {
        st->done = 1, st->termination = WC_TERM_PARAMETER_TOL;
      } else if (fabs(st->x_cost - cand) <= o.function_tolerance * st->x_cost) {
        st->done = 1, st->termination = WC_TERM_FUNCTION_TOL;
      } else {
        const double rd = (st->x_cost - cand) / st->model_cost_change;
        if (rd > o.min_relative_decrease) {
          *s_accept  = 1;
          st->x_cost = cand;
          st->cur    = 1 - st->cur;
          st->iter_accepted[it] = 1;
          st->num_successful += 1;
          st->last_successful = 1;
          const double d      = 1.0 - pow(2.0 * rd - 1.0, 3);
          st->radius          = fmin(o.max_trust_region_radius, st->radius / fmax(1.0 / 3.0, d));
          st->decrease_factor = 2.0, st->reuse_diagonal = 0;
        } else {
          st->num_unsuccessful += 1;
          st->radius = st->radius / st->decrease_factor, st->decrease_factor *= 2.0, st->reuse_diagonal = 1;
        }
      }
    }
  }
  __syncthreads();
  if (!*s_accept) return;  // uniform
  const double* g = B.g[st->cur];
  double gm = 0.0, xn = 0.0;
  for (int i = t; i < N; i += LMT) {
    const double v = B.xc[i];
    B.x[i]         = v;
    xn += v * v;
    if (col_of(i, ff) >= 0) gm = fmax(gm, fabs(g[i]));
  }
  gm = block_max(gm, red);
  xn = block_sum(xn, red);
  if (t == 0) st->grad_max = gm, st->x_norm = sqrt(xn);
  __syncthreads();
}

// ---- blocked right-looking Cholesky with the right-hand side carried as an extra row, with look-ahead.
// Layout: Dp = D rounded up to the block width CB (padding rows/columns are identity), leading dimension LD even with
// LD/2 odd (16-byte row segments of consecutive rows fall into distinct bank groups), Dp + 4 rows: rows 0..Dp-1 the SPD
// matrix (lower triangle), row Dp the right-hand side g^T, rows Dp+1..Dp+3 zero (so every row tile of 4 is full).
// After the call rows 0..Dp-1 hold L and row Dp holds z = L^-1 g (the forward substitution rides along).
// Per block column k the critical path is  panel(k) -> update of block column k+1 -> factor(k+1):
//   S1 panel rows = A_panel * Lkk^-T by forward substitution, one thread per row (no explicit inverse);
//   S2 the next block column only (all threads, 1 x 4 strips);
//   S3 warp 0 factorises diagonal block k+1 in registers (shuffles, reciprocal square roots, no divisions on the pivot
//      chain) WHILE the warps of the other three schedulers finish the trailing update in 4 x 4 register tiles fed by
//      16-byte shared loads.
__device__ __forceinline__ int chol_ld(int Dp) { return ((Dp >> 1) & 1) ? Dp : Dp + 2; }

// Diagonal block factor by one warp: the CB x CB block is spread over the lanes (lane = 8 * (c & 3) + r holds a[r][c]
// and a[r][c + 4]) and factorised by a ROLLED pivot loop — a few dozen instructions that stay in the L0 instruction
// cache, instead of a fully unrolled register-array version whose straight-line code is instruction-fetch bound on a
// single warp.  Per pivot j: broadcast d = a[j][j], rs = rsqrt(d), gather column j for the lane's row and its two
// columns, then a[r][c] -= (a[r][j] rs)(a[c][j] rs) for r, c > j and a[r][j] *= rs.  Runs while the other warps are
// idle or in light phases (shuffles share the shared-memory pipe).
__device__ __forceinline__ void chol_factor_diag(double* A, int LD, int k0, double* rinv_out, int* s_fail) {
  const int lane = threadIdx.x & 31, r = lane & 7, c0 = lane >> 3;
  double    e0 = A[(k0 + r) * LD + k0 + c0], e1 = A[(k0 + r) * LD + k0 + c0 + 4];  // upper-triangle values are ignored
  bool      bad = false;
#pragma unroll 1
  for (int j = 0; j < CB; ++j) {
    const int    src = (j & 3) * 8;              // lanes holding column j
    const double colv = j < 4 ? e0 : e1;          // this lane's value IF it holds column j (only read from such lanes)
    const double d    = __shfl_sync(0xffffffffu, colv, src + j);
    const double ar   = __shfl_sync(0xffffffffu, colv, src + r);
    const double ac0  = __shfl_sync(0xffffffffu, colv, src + c0);
    const double ac1  = __shfl_sync(0xffffffffu, colv, src + c0 + 4);
    if (!(d > 0.0) || !isfinite(d)) bad = true;
    const double rs = rsqrt(d);
    const double lr = ar * rs;
    if (r > j) {
      if (c0 > j) e0 = fma(-lr, ac0 * rs, e0);
      if (c0 + 4 > j) e1 = fma(-lr, ac1 * rs, e1);
    }
    if (r >= j) {  // column j becomes final: L[r][j] = a[r][j] rs (the diagonal: d rs = sqrt(d))
      if (c0 == j) e0 = lr;
      if (c0 + 4 == j) e1 = lr;
    }
    if (lane == j) rinv_out[j] = rs;
  }
  if (bad && lane == 0) *s_fail = 1;
  if (c0 <= r) A[(k0 + r) * LD + k0 + c0] = e0;
  if (c0 + 4 <= r) A[(k0 + r) * LD + k0 + c0 + 4] = e1;
}

#ifndef WC_FACTOR_ROLLED
// default (measured in the kernel, 17 block factors: 44.7 k cycles vs 51.0 k for the rolled lane-grid variant above and
// 43.6 k for an all-in-registers variant that needs 196 registers): lanes 0..CB-1 hold one row each (CB registers),
// fully unrolled pivot loop, shuffles for the pivot row
__device__ __forceinline__ void chol_factor_diag_rows(double* A, int LD, int k0, double* rinv_out, int* s_fail) {
  const int lane = threadIdx.x & 31;
  double    a[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) a[c] = (lane < CB && c <= lane) ? A[(k0 + lane) * LD + k0 + c] : (c == lane ? 1.0 : 0.0);
  bool   bad  = false;
  double rinv = 1.0;
#pragma unroll
  for (int j = 0; j < CB; ++j) {
    const double djj = __shfl_sync(0xffffffffu, a[j], j);
    if (!(djj > 0.0) || !isfinite(djj)) bad = true;
    const double rs = rsqrt(djj);
    if (lane == j) rinv = rs;
    if (lane >= j) a[j] = (lane == j) ? djj * rs : a[j] * rs;
#pragma unroll
    for (int k = j + 1; k < CB; ++k) {
      const double lkj = __shfl_sync(0xffffffffu, a[j], k);
      if (lane >= k) a[k] -= a[j] * lkj;
    }
  }
  if (bad && lane == 0) *s_fail = 1;
  if (lane < CB) {
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c <= lane) A[(k0 + lane) * LD + k0 + c] = a[c];
    rinv_out[lane] = rinv;
  }
}
#define chol_factor_diag chol_factor_diag_rows
#endif

// load 8 consecutive doubles (16-byte aligned) as four 128-bit accesses
__device__ __forceinline__ void ld8(const double* p, double* v) {
  const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double2 d = q[k];
    v[2 * k] = d.x, v[2 * k + 1] = d.y;
  }
}

// panel row i: L[i][k0..k0+CB) = A[i][k0..k0+CB) * Lkk^-T by forward substitution
__device__ __forceinline__ void chol_panel_row(double* A, int LD, int i, int k0, const double* rinv) {
  double ai[CB], li[CB];
  ld8(A + i * LD + k0, ai);
#pragma unroll
  for (int b = 0; b < CB; ++b) {
    double v = ai[b];
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < b) v = fma(-li[c], A[(k0 + b) * LD + k0 + c], v);
    li[b] = v * rinv[b];
  }
  double2* q = reinterpret_cast<double2*>(A + i * LD + k0);
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = make_double2(li[2 * k], li[2 * k + 1]);
}

// warp 0: next diagonal block D' = A[r0..r0+CB)[r0..r0+CB) - P P^T (lower triangle, 36 entries over the lanes)
__device__ __forceinline__ void chol_update_next_diag(double* A, int LD, int k0, int r0) {
  const int lane = threadIdx.x & 31;
  for (int e = lane; e < CB * (CB + 1) / 2; e += 32) {
    int r = 0;
    while ((r + 1) * (r + 2) / 2 <= e) ++r;
    const int c = e - r * (r + 1) / 2;
    double    pr[CB], pc[CB];
    ld8(A + (r0 + r) * LD + k0, pr);
    ld8(A + (r0 + c) * LD + k0, pc);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int b = 0; b < CB; b += 2) s0 = fma(pr[b], pc[b], s0), s1 = fma(pr[b + 1], pc[b + 1], s1);
    A[(r0 + r) * LD + r0 + c] -= s0 + s1;
  }
}

// Trailing update: rows [r1, Dp] with r1 = r0 + CB (row Dp = right-hand side), columns [r0, min(row, Dp - 1)].
// One warp per 16 x 16 tile; lane (rho = lane & 7, gam = lane >> 3) owns the INTERLEAVED elements rows {rho, rho + 8} x
// columns {gam, gam + 4, gam + 8, gam + 12} of the tile, so that every 16-byte shared load of a warp touches eight
// consecutive rows (li) or four consecutive rows (lj) — conflict-free with the 16-byte row stride pattern of LD (a
// contiguous 4 x 4 tile per thread would put the lanes 4 rows apart: 4-way bank conflicts on every load).
// Row band bi (rows r1 + 16 bi ..) has column bands 0 .. bi + 1.
__device__ __forceinline__ void chol_trailing_tiles(double* A, int LD, int Dp, int k0, int r0, int wid, int nwarps) {
  const int lane = threadIdx.x & 31, rho = lane & 7, gam = lane >> 3;
  const int r1    = r0 + CB;
  const int nb    = (Dp + 1 - r1 + 15) >> 4;
  const int ntile = nb * (nb + 3) / 2;
  for (int q = wid; q < ntile; q += nwarps) {
    int bi = (int)((sqrtf(9.f + 8.f * (float)q) - 3.f) * 0.5f);
    while (bi * (bi + 3) / 2 > q) --bi;
    while ((bi + 1) * (bi + 4) / 2 <= q) ++bi;
    const int bj = q - bi * (bi + 3) / 2;
    const int i0 = r1 + 16 * bi + rho, j0 = r0 + 16 * bj + gam;
    double    lj[4][CB], li[CB];
#pragma unroll
    for (int c = 0; c < 4; ++c) ld8(A + min(j0 + 4 * c, Dp) * LD + k0, lj[c]);  // clamped rows are masked below
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int i = i0 + 8 * a;
      if (i > Dp) continue;
      ld8(A + i * LD + k0, li);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + 4 * c;
        if (j >= Dp || (j > i && i < Dp)) continue;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int b = 0; b < CB; b += 2) s0 = fma(li[b], lj[c][b], s0), s1 = fma(li[b + 1], lj[c][b + 1], s1);
        A[i * LD + j] -= s0 + s1;
      }
    }
  }
}

#ifdef WC_LM_TIMING
__device__ unsigned long long g_lm_stamps[512];
__device__ int                g_lm_nstamp;
__device__ __forceinline__ unsigned long long wc_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__global__ void lm_stamps_print() {
  printf("lm_step start/end stamps (ns since the first):");
  for (int i = 0; i < g_lm_nstamp && i < 512; ++i) printf(" %llu", g_lm_stamps[i] - g_lm_stamps[0]);
  printf("\n");
  g_lm_nstamp = 0;
}
#define WC_TICK() c0 = clock64()
#define WC_TOCK(acc) acc += clock64() - c0
#else
#define WC_TICK()
#define WC_TOCK(acc)
#endif

// Per block column k, two barriers:
//   phase 1  panel(k): warp 0 solves the CB rows of the NEXT diagonal block and updates that block; the other warps
//            solve the panel rows below;
//   phase 2  warp 0 factorises the next diagonal block (the serial pivot chain)  ||  warps 1..15: the remaining
//            trailing update, one 16 x 16 tile per warp.
__device__ void cholesky_blocked_rhs(double* A, int Dp, double* rinv, int* s_fail) {
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, LD = chol_ld(Dp);
#ifdef WC_LM_TIMING
  long long c_p1 = 0, c_p2 = 0, c_fact = 0, c_trail = 0, c0;
#endif
  if (warp == 0) chol_factor_diag(A, LD, 0, rinv, s_fail);
  __syncthreads();
  for (int k0 = 0; k0 < Dp; k0 += CB) {
    if (*s_fail) return;  // uniform (written before the last barrier)
    const int r0 = k0 + CB;
    WC_TICK();
    if (warp == 0) {
      const int i = r0 + lane;
      if (lane < CB && i <= Dp) chol_panel_row(A, LD, i, k0, rinv + k0);
      if (r0 < Dp) {
        __syncwarp();
        chol_update_next_diag(A, LD, k0, r0);
      }
    } else {
      for (int i = r0 + CB + (t - 32); i <= Dp; i += LMT - 32) chol_panel_row(A, LD, i, k0, rinv + k0);
    }
    __syncthreads();
    WC_TOCK(c_p1);
    if (r0 >= Dp) break;  // last block column: only the right-hand-side row remained
    WC_TICK();
    if (warp == 0) {
      chol_factor_diag(A, LD, r0, rinv + r0, s_fail);
#ifdef WC_LM_TIMING
      c_fact += clock64() - c0;
#endif
    } else {
      chol_trailing_tiles(A, LD, Dp, k0, r0, warp - 1, LMT / 32 - 1);
#ifdef WC_LM_TIMING
      if (t == LMT - 1) c_trail += clock64() - c0;
#endif
    }
    __syncthreads();
    WC_TOCK(c_p2);
  }
#ifdef WC_LM_TIMING
  if (t == 0) printf("chol cycles: phase1 (panel, diag update) %lld phase2 %lld of which warp 0's factor %lld\n", c_p1, c_p2, c_fact);
  if (t == LMT - 1) printf("chol cycles: last warp's trailing %lld\n", c_trail);
#endif
}

// Backward substitution L^T x = z by ONE warp, no block barriers: z (row Dp of A) lives in registers, element i in
// lane i % 32, slot i / 32.  For j = Dp-1 .. 0: x_j = z_j * rinv_j (broadcast by one shuffle), then z_i -= L[j][i] x_j
// for i < j — row j of L is contiguous, so the loads are conflict-free and independent of the pivot chain.  The slot
// of the pivot is a compile-time constant of each 32-column sweep, so the loop body is a dozen instructions (a lone
// warp pays ~4 cycles per dependent instruction: instruction count, not flops, is what matters here).
// The result is written negated into y (length D).  Systems wider than 32 * BS_SLOTS fall back to a serial sweep.
constexpr int BS_SLOTS = 8;  // up to 256 unknowns in registers
template <int SJ>
__device__ __forceinline__ void backward_sweep(const double* A, int LD, int D, int Dp, const double* rinv, double (&z)[BS_SLOTS],
                                               double* y) {
  const int lane = threadIdx.x & 31;
  if (32 * SJ >= Dp) return;
  const int jhi = min(31, Dp - 1 - 32 * SJ);
#pragma unroll 1
  for (int jj = jhi; jj >= 0; --jj) {
    const int     j  = 32 * SJ + jj;
    const double* Lj = A + (size_t)j * LD + lane;
    double        l[SJ + 1];
#pragma unroll
    for (int s = 0; s < SJ; ++s) l[s] = Lj[32 * s];
    l[SJ] = lane < jj ? Lj[32 * SJ] : 0.0;
    const double xj = __shfl_sync(0xffffffffu, z[SJ], jj) * rinv[j];
#pragma unroll
    for (int s = 0; s <= SJ; ++s) z[s] = fma(-l[s], xj, z[s]);
    if (lane == jj && j < D) y[j] = -xj;
  }
}
__device__ void chol_backward_warp(double* A, int D, int Dp, const double* rinv, double* y) {
  const int LD = chol_ld(Dp), lane = threadIdx.x & 31;
  double*   zrow = A + (size_t)Dp * LD;
  if (Dp <= 32 * BS_SLOTS) {
    if (threadIdx.x >= 32) return;
    double z[BS_SLOTS];
#pragma unroll
    for (int s = 0; s < BS_SLOTS; ++s) z[s] = (32 * s + lane < Dp) ? zrow[32 * s + lane] : 0.0;
    backward_sweep<7>(A, LD, D, Dp, rinv, z, y);
    backward_sweep<6>(A, LD, D, Dp, rinv, z, y);
    backward_sweep<5>(A, LD, D, Dp, rinv, z, y);
    backward_sweep<4>(A, LD, D, Dp, rinv, z, y);
    backward_sweep<3>(A, LD, D, Dp, rinv, z, y);
    backward_sweep<2>(A, LD, D, Dp, rinv, z, y);
    backward_sweep<1>(A, LD, D, Dp, rinv, z, y);
    backward_sweep<0>(A, LD, D, Dp, rinv, z, y);
  } else {
    // wide systems (K > 21 control poses): column sweep by the whole CTA, one barrier per column
    for (int j = Dp - 1; j >= 0; --j) {
      const double  xj = zrow[j] * rinv[j];
      const double* Lj = A + (size_t)j * LD;
      for (int i = threadIdx.x; i < j; i += LMT) zrow[i] = fma(-Lj[i], xj, zrow[i]);
      if (threadIdx.x == 0 && j < D) y[j] = -xj;
      __syncthreads();
    }
  }
}

// One LM iteration boundary in one launch:
//   (1) lm_decide on the outstanding candidate (if any),
//   (2) FinalizeIterationAndCheckIfMinimizerCanContinue,
//   (3) LevenbergMarquardtStrategy::ComputeStep + model cost change + candidate point,
//   (4) clears the normal-equation buffer the next linearisation accumulates into.
template <bool A_IN_SMEM>
__global__ void __launch_bounds__(LMT) lm_step(SolveBufs B, wc_solve_opts o, int zero_next) {
  extern __shared__ __align__(16) double sA[];
  __shared__ double red[LMT / 32];
  __shared__ int    s_fail, s_accept, s_done, s_pending;
  // the LM state block is staged in shared memory (one coalesced read, one coalesced write-back): thread 0's
  // bookkeeping then costs shared-memory latencies instead of a chain of dependent L2 round trips
  __shared__ __align__(16) LMState sst;
  static_assert(sizeof(LMState) % 8 == 0, "LMState is copied as 8-byte words");
  const int t  = threadIdx.x;
  pdl_trigger();
  pdl_wait();
#ifdef WC_LM_TIMING
  if (t == 0) { const int k = atomicAdd(&g_lm_nstamp, 1); if (k < 512) g_lm_stamps[k] = wc_globaltimer(); }
#endif
  for (int k = t; k < (int)(sizeof(LMState) / 8); k += LMT)
    reinterpret_cast<unsigned long long*>(&sst)[k] = reinterpret_cast<const unsigned long long*>(B.st)[k];
  __syncthreads();
  LMState* st = &sst;
  auto write_back = [&]() {
    __syncthreads();
    for (int k = t; k < (int)(sizeof(LMState) / 8); k += LMT)
      reinterpret_cast<unsigned long long*>(B.st)[k] = reinterpret_cast<const unsigned long long*>(&sst)[k];
  };
  // control flags are broadcast through shared memory: thread 0 rewrites them below while other warps may lag
  if (t == 0) s_done = st->done, s_pending = st->pending;
  __syncthreads();
  if (s_done) return;
  const int N = B.N, ff = B.fix_first, D = st->D;
#ifdef WC_LM_TIMING
  long long tk[8];
  tk[0] = clock64();
#endif
  if (s_pending) lm_decide_dev(B, st, o, red, &s_accept);
  __syncthreads();
#ifdef WC_LM_TIMING
  tk[1] = clock64();
#endif
  if (t == 0) {
    int term = -1;
    if (!st->done) {
      if (st->iteration >= o.max_num_iterations) term = WC_TERM_NO_CONVERGENCE;
      else if (st->last_successful && st->grad_max <= o.gradient_tolerance) term = WC_TERM_GRADIENT_TOL;
      else if (st->radius < o.min_trust_region_radius) term = WC_TERM_MIN_RADIUS;
      if (term >= 0) st->done = 1, st->termination = term;
    }
    s_fail = 0;
    s_done = st->done;
  }
  __syncthreads();
  if (s_done) {
    write_back();
    return;
  }
  const double* H = B.H[st->cur];
  const double* g = B.g[st->cur];
  const int     Dp = (D + CB - 1) / CB * CB, LD = chol_ld(Dp);
  double*       A    = A_IN_SMEM ? sA : B.A;            // (Dp + 4) x LD; compile-time choice: shared accesses are LDS/STS
  double*       rinv = A + (size_t)(Dp + 4) * LD;       // reciprocal pivots
  const double  radius = st->radius;
  // A = S H S + diag / radius, lower triangle only; padding rows are identity, row Dp is the right-hand side g_s = S g,
  // rows Dp+1..Dp+3 are zero.
  double* sscale = rinv + Dp;    // A_IN_SMEM: jacobian scaling and LM diagonal of this step, staged next to the pivots
  double* sdiag  = sscale + Dp;
  if constexpr (A_IN_SMEM) {
    // The raw lower-triangle rows of H (and g) come in by asynchronous 16-byte copies (cp.async: global -> shared without
    // a register stage), ~10 per thread and all in flight at once, where register-staged loads paid ~5 dependent L2 round
    // trips.  (One bulk copy per row through the TMA unit was measured too: 142 small copies serialise in the copy engine
    // and take as long as the loads did.)  Row r lands at the start of A's row r in H's own column numbering; the scaling
    // pass below compacts it in place (the three fixed position columns drop out).
    const double spre = t < D ? B.scale[t] : 1.0;
    const double dpre = (t < D && st->reuse_diagonal) ? B.diag[t] : 0.0;
    {
      const int lane = t & 31, nw = LMT / 32;
      for (int r = t >> 5; r <= D; r += nw) {
        const double* src = r < D ? H + (size_t)amb_of(r, ff) * N : g;            // 16-byte aligned rows (N is even)
        double*       dst = A + (size_t)(r < D ? r : Dp) * LD;
        const int     nch = r < D ? (amb_of(r, ff) >> 1) + 1 : N / 2;              // 16-byte chunks: columns 0 .. amb(r) (+1)
        for (int j = lane; j < nch; j += 32)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wctma::smem_addr(dst + 2 * j)), "l"(src + 2 * j) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (t < D) {
      double d = dpre;
      if (!st->reuse_diagonal) {
        d         = fmin(fmax(A[(size_t)t * LD + amb_of(t, ff)] * spre * spre, o.min_lm_diagonal), o.max_lm_diagonal);
        B.diag[t] = d;
      }
      sscale[t] = spre, sdiag[t] = d;
    }
    __syncthreads();
    // (rolled loops on purpose: this code runs once per launch, from a cold instruction cache — an unrolled version is
    //  fetch bound)
    const int lane = t & 31, nw = LMT / 32;
#pragma unroll 1
    for (int r = t >> 5; r < Dp + 4; r += nw) {
      double*      Ar = A + (size_t)r * LD;
      const double sr = r < D ? sscale[r] : 1.0;
      const int    c1 = r < Dp ? r : Dp - 1;  // last column of this row
#pragma unroll 1
      for (int c = lane; c - lane <= c1; c += 32) {
        // the row is compacted in place, 32 columns at a time in ascending order: a chunk reads columns up to c + 3 and
        // writes columns <= c, so later chunks still find their raw values
        double x = 0.0;
        if (c < D && (r == Dp || (r < D && c <= r))) x = Ar[amb_of(c, ff)];
        __syncwarp();
        if (c > c1) continue;
        if (r < D) {
          x *= sr * sscale[c];
          if (r == c) {
            const double sq = sqrt(sdiag[r] / radius);
            x += sq * sq;
          }
        } else if (r < Dp) {
          x = c == r ? 1.0 : 0.0;
        } else if (r == Dp) {
          x = c < D ? x * sscale[c] : 0.0;
        } else {
          x = 0.0;
        }
        Ar[c] = x;
      }
    }
  } else {
  if (!st->reuse_diagonal)
    for (int c = t; c < D; c += LMT) {
      const int    i = amb_of(c, ff);
      const double d = H[(size_t)i * N + i] * B.scale[c] * B.scale[c];
      B.diag[c]      = fmin(fmax(d, o.min_lm_diagonal), o.max_lm_diagonal);
    }
  __syncthreads();
  // A = S H S + diag / radius, lower triangle only; padding rows are identity, row Dp is the right-hand side g_s = S g,
  // rows Dp+1..Dp+3 are zero.  Warps over rows, lanes over columns (coalesced rows of H); two rows x five column chunks
  // of loads are issued before the first use so that ~10 L2 round trips are in flight per thread.
  {
    const int lane = t & 31, nw = LMT / 32;
    for (int rb = t >> 5; rb < Dp + 4; rb += 2 * nw) {
      double v[2][5];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = rb + u * nw;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int c = lane + 32 * q;
          v[u][q]     = 0.0;
          if (r < D && c <= r) v[u][q] = H[(size_t)amb_of(r, ff) * N + amb_of(c, ff)];
          else if (r == Dp && c < D) v[u][q] = g[amb_of(c, ff)];
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = rb + u * nw;
        if (r >= Dp + 4) continue;
        double*      Ar = A + (size_t)r * LD;
        const double sr = r < D ? B.scale[r] : 1.0;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int c = lane + 32 * q;
          if (c >= Dp || (r < Dp && c > r)) continue;
          double x = v[u][q];
          if (r < D) {
            x *= sr * B.scale[c];
            if (r == c) {
              const double sq = sqrt(B.diag[r] / radius);
              x += sq * sq;
            }
          } else if (r < Dp) {
            x = c == r ? 1.0 : 0.0;
          } else if (r == Dp && c < D) {
            x *= B.scale[c];
          }
          Ar[c] = x;
        }
        for (int c = lane + 160; c < Dp && (r >= Dp || c <= r); c += 32) {  // D > 160: remaining columns, plain loop
          double x = 0.0;
          if (r < D) {
            x = H[(size_t)amb_of(r, ff) * N + amb_of(c, ff)] * sr * B.scale[c];
            if (r == c) {
              const double sq = sqrt(B.diag[r] / radius);
              x += sq * sq;
            }
          } else if (r < Dp) {
            x = c == r ? 1.0 : 0.0;
          } else if (r == Dp && c < D) {
            x = g[amb_of(c, ff)] * B.scale[c];
          }
          Ar[c] = x;
        }
      }
    }
  }
  }  // !A_IN_SMEM
  __syncthreads();
#ifdef WC_LM_TIMING
  tk[2] = clock64();
#endif
  cholesky_blocked_rhs(A, Dp, rinv, &s_fail);
  __syncthreads();
#ifdef WC_LM_TIMING
  tk[3] = clock64();
#endif
  bool    valid = !s_fail;
  double* y     = B.step;
  // the backward substitution of a system this size occupies one warp: the other warps clear the normal-equation buffer
  // the next linearisation accumulates into, instead of doing that after the solve with everybody waiting
  const bool zero_early = zero_next && Dp <= 32 * BS_SLOTS && t >= 32;
  if (zero_early) {
    const int nb = 1 - st->cur;
    double2*  Hz = reinterpret_cast<double2*>(B.H[nb]);  // N = 12 K: N * N is even, cudaMalloc alignment
    for (int i = t - 32; i < N * N / 2; i += LMT - 32) Hz[i] = make_double2(0.0, 0.0);
    for (int i = t - 32; i < N; i += LMT - 32) B.g[nb][i] = 0.0;
    if (t == 32) *B.cost[nb] = 0.0;
  }
  if (valid) chol_backward_warp(A, D, Dp, rinv, y);  // y = -(S H S + diag/radius)^-1 g_s
  __syncthreads();
#ifdef WC_LM_TIMING
  tk[4] = clock64();
#endif
  // model_cost_change = -y^T (g_s + H_s y / 2).  y solves (H_s + Dg) y = -g_s with Dg = diag / radius, hence
  // H_s y = -g_s - Dg y and the change is (y^T Dg y - y^T g_s) / 2: no matrix-vector product needed.
  double part = 0.0, bad = 0.0;
  if (valid)
    for (int c = t; c < D; c += LMT) {
      const double sq = sqrt(B.diag[c] / radius);
      part += 0.5 * y[c] * (sq * sq * y[c] - g[amb_of(c, ff)] * B.scale[c]);
      if (!isfinite(y[c])) bad = 1.0;
    }
  const double mcc  = block_sum(part, red);
  const double nbad = block_sum(bad, red);
  valid             = valid && nbad == 0.0 && mcc > 0.0;
  double sn = 0.0;
  for (int i = t; i < N; i += LMT) {
    const int    c = col_of(i, ff);
    const double d = (valid && c >= 0) ? y[c] * B.scale[c] : 0.0;
    B.xc[i]        = B.x[i] + d;
    sn += d * d;
  }
  sn = block_sum(sn, red);
#ifdef WC_LM_TIMING
  tk[5] = clock64();
#endif
  if (zero_next && Dp > 32 * BS_SLOTS) {  // wide systems: every thread took part in the backward sweep
    const int nb = 1 - st->cur;
    double2* Hz = reinterpret_cast<double2*>(B.H[nb]);  // N = 12 K: N * N is even, cudaMalloc alignment
    for (int i = t; i < N * N / 2; i += LMT) Hz[i] = make_double2(0.0, 0.0);
    for (int i = t; i < N; i += LMT) B.g[nb][i] = 0.0;
    if (t == 0) *B.cost[nb] = 0.0;
  }
  if (t == 0) {
    st->iteration += 1;
    st->last_successful   = 0;
    st->reuse_diagonal    = 1;
    st->step_valid        = valid ? 1 : 0;
    st->pending           = 1;
    st->model_cost_change = mcc;
    st->step_norm         = sqrt(sn);
    const int it          = st->iteration < WC_MAX_ITER_LOG ? st->iteration : WC_MAX_ITER_LOG - 1;
    st->iter_radius[it]   = radius;
#ifdef WC_LM_TIMING
    tk[6] = clock64();
    if (st->iteration == 3)
      printf("lm_step cycles: decide %lld build %lld chol %lld back %lld mcc %lld zero %lld total %lld\n", tk[1] - tk[0],
             tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4], tk[6] - tk[5], tk[6] - tk[0]);
#endif
  }
  write_back();
#ifdef WC_LM_TIMING
  __syncthreads();
  if (t == 0) { const int k = atomicAdd(&g_lm_nstamp, 1); if (k < 512) g_lm_stamps[k] = wc_globaltimer(); }
#endif
}


// ---------------------------------------------------------------------------------------------- K7, wide systems
// Systems that do not fit one SM's shared memory (more than ~21 control poses): one COOPERATIVE launch over all SMs.
// Blocked right-looking Cholesky (block WNB = 32) of the active unknowns in global memory (L2 resident), the right-hand
// side carried as an extra row; per block column: every CTA factorises the diagonal block redundantly in shared memory
// (no broadcast, no extra barrier), the panel rows are spread over the CTAs, grid barrier, the trailing 32 x 32 tiles are
// spread over the CTAs, grid barrier.  CTA 0 does the LM bookkeeping before and the blocked backward substitution after.
constexpr int WNB = 32;

// Cholesky of a WNB x WNB block held in shared memory (leading dimension WNB + 1) by the whole CTA: thread e owns the
// elements e and e + LMT of the block; per pivot ONE barrier: everybody reads the (still unscaled) pivot column, applies
// the rank-1 update to its elements of the trailing block, and the owners of the pivot column write its scaled values at
// the start of the next round (nobody reads that column again).  Leaves the reciprocal pivots; *fail is set on a
// non-positive pivot.  (A one-warp register version of this block is 18x slower here: its 32-entry row does not fit the
// 128 registers a 512-thread CTA leaves per thread.)
__device__ __forceinline__ void wide_factor_diag(double (*L)[WNB + 1], double* rinv, int* fail) {
  const int t = threadIdx.x;
  static_assert(WNB * WNB == 2 * LMT, "two elements per thread");
  const int r0 = t >> 5, c0 = t & 31, r1 = r0 + WNB / 2;  // elements (r0, c0) and (r1, c0)
  if (t == 0) *fail = 0;
  __syncthreads();
  double rs_prev = 0.0;
#pragma unroll 1
  for (int j = 0; j < WNB; ++j) {
    if (j > 0 && c0 == j - 1) {  // scale the previous pivot column (its readers are past the barrier)
      if (r0 >= j - 1) L[r0][c0] *= rs_prev;
      if (r1 >= j - 1) L[r1][c0] *= rs_prev;
    }
    const double d = L[j][j];
    if (t == 0 && (!(d > 0.0) || !isfinite(d))) *fail = 1;
    const double rs = rsqrt(d), rs2 = rs * rs;
    if (t == 0) rinv[j] = rs;
    if (c0 > j) {
      const double lc = L[c0][j] * rs2;
      if (c0 <= r0) L[r0][c0] -= L[r0][j] * lc;  // (j < c0 <= r: the row is below the pivot too)
      if (c0 <= r1) L[r1][c0] -= L[r1][j] * lc;
    }
    rs_prev = rs;
    __syncthreads();
  }
  if (c0 == WNB - 1) {
    if (r0 >= WNB - 1) L[r0][c0] *= rs_prev;
    if (r1 >= WNB - 1) L[r1][c0] *= rs_prev;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(LMT) lm_step_wide(SolveBufs B, wc_solve_opts o, int zero_next) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[LMT / 32];
  __shared__ int    s_accept, s_fail;
  __shared__ __align__(16) LMState sst;  // CTA 0's working copy of the LM state
  __shared__ double sL[WNB][WNB + 1], sPi[WNB][WNB + 1], sPj[WNB][WNB + 1];
  __shared__ double sRinv[WNB], sx[WNB];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, cta = blockIdx.x, ncta = gridDim.x;
  LMState*  gst = B.st;
#ifdef WC_LM_TIMING
  long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c_fac = 0, c_pan = 0, c_upd = 0, c0;
  tk[0] = clock64();
#endif
  if (cta == 0) {
    for (int k = t; k < (int)(sizeof(LMState) / 8); k += LMT)
      reinterpret_cast<unsigned long long*>(&sst)[k] = reinterpret_cast<const unsigned long long*>(gst)[k];
    __syncthreads();
    if (!sst.done) {
      if (sst.pending) lm_decide_dev(B, &sst, o, red, &s_accept);
      __syncthreads();
      if (t == 0 && !sst.done) {
        int term = -1;
        if (sst.iteration >= o.max_num_iterations) term = WC_TERM_NO_CONVERGENCE;
        else if (sst.last_successful && sst.grad_max <= o.gradient_tolerance) term = WC_TERM_GRADIENT_TOL;
        else if (sst.radius < o.min_trust_region_radius) term = WC_TERM_MIN_RADIUS;
        if (term >= 0) sst.done = 1, sst.termination = term;
      }
      __syncthreads();
      for (int k = t; k < (int)(sizeof(LMState) / 8); k += LMT)
        reinterpret_cast<unsigned long long*>(gst)[k] = reinterpret_cast<const unsigned long long*>(&sst)[k];
      __threadfence();
    }
  }
  grid.sync();
#ifdef WC_LM_TIMING
  tk[1] = clock64();
#endif
  if (*(volatile int*)&gst->done) return;  // uniform over the grid
  const int     N = B.N, ff = B.fix_first, D = *(volatile int*)&gst->D, cur = *(volatile int*)&gst->cur;
  const double  radius = *(volatile double*)&gst->radius;
  const int     reuse  = *(volatile int*)&gst->reuse_diagonal;
  const double* H = B.H[cur];
  const double* g = B.g[cur];
  const int     Da = B.act[D], nblk = (Da + WNB - 1) / WNB, Dp = nblk * WNB, LD = Dp;
  double*       A  = B.A;  // (Dp + 1) x LD: lower triangle of the damped scaled system, row Dp = scaled gradient
  // ---- build
  for (int r = cta * (LMT / 32) + warp; r <= Dp; r += ncta * (LMT / 32)) {
    double* Ar = A + (size_t)r * LD;
    if (r < Da) {
      const int    cr = B.act[r], ir = amb_of(cr, ff);
      const double sr = B.scale[cr];
      for (int c = lane; c <= r; c += 32) {
        const int cc = B.act[c];
        double    x  = H[(size_t)ir * N + amb_of(cc, ff)] * sr * B.scale[cc];
        if (c == r) {
          double d = B.diag[cr];
          if (!reuse) {
            d          = fmin(fmax(H[(size_t)ir * N + ir] * sr * sr, o.min_lm_diagonal), o.max_lm_diagonal);
            B.diag[cr] = d;
          }
          const double sq = sqrt(d / radius);
          x += sq * sq;
        }
        Ar[c] = x;
      }
    } else if (r < Dp) {
      for (int c = lane; c <= r; c += 32) Ar[c] = c == r ? 1.0 : 0.0;
    } else {
      for (int c = lane; c < Dp; c += 32) Ar[c] = c < Da ? g[amb_of(B.act[c], ff)] * B.scale[B.act[c]] : 0.0;
    }
  }
  grid.sync();
#ifdef WC_LM_TIMING
  tk[2] = clock64();
#endif
  // ---- factorisation
  bool ok = true;
  for (int k = 0; k < nblk && ok; ++k) {
    const int k0 = k * WNB;
    WC_TICK();
    // (a) diagonal block, redundantly in every CTA
    for (int e = t; e < WNB * WNB; e += LMT) {
      const int r = e / WNB, c = e % WNB;
      sL[r][c] = c <= r ? __ldcg(A + (size_t)(k0 + r) * LD + k0 + c) : 0.0;
    }
    __syncthreads();
    wide_factor_diag(sL, sRinv, &s_fail);
    ok = !s_fail;
    if (!ok) break;  // every CTA computed the same block: uniform over the grid
    if (cta == 0)
      for (int e = t; e < WNB * WNB; e += LMT) {
        const int r = e / WNB, c = e % WNB;
        if (c <= r) A[(size_t)(k0 + r) * LD + k0 + c] = sL[r][c];
      }
    WC_TOCK(c_fac);
    WC_TICK();
    // (b) panel rows below (and the right-hand-side row): L[i][k0..k0+WNB) = A[i][k0..) * Lkk^-T, one thread per row,
    //     rows dealt to the CTAs first so that the latency-bound solves run on as many SMs as possible
    const int nrow = Dp + 1 - (k0 + WNB);
    for (int q = cta + ncta * t; q < nrow; q += ncta * LMT) {
      double*   Ai = A + (size_t)(k0 + WNB + q) * LD + k0;
      double    li[WNB];
#pragma unroll
      for (int b = 0; b < WNB; ++b) li[b] = __ldcg(Ai + b);
#pragma unroll
      for (int b = 0; b < WNB; ++b) {
        double v = li[b];
#pragma unroll
        for (int c = 0; c < WNB; ++c)
          if (c < b) v = fma(-li[c], sL[b][c], v);
        li[b] = v * sRinv[b];
      }
#pragma unroll
      for (int b = 0; b < WNB; ++b) Ai[b] = li[b];
    }
    grid.sync();
    WC_TOCK(c_pan);
    WC_TICK();
    // (c) trailing update: tiles (bi >= bj > k) of WNB x WNB plus the right-hand-side row, dealt round robin
    const int nb = nblk - k - 1;                   // block rows / columns left
    const int ntile = nb * (nb + 1) / 2 + nb;      // lower triangle + the right-hand-side row's nb tiles
    for (int q = cta; q < ntile; q += ncta) {
      int bi, bj;
      if (q < nb * (nb + 1) / 2) {
        bi = (int)((sqrtf(1.f + 8.f * (float)q) - 1.f) * 0.5f);
        while (bi * (bi + 1) / 2 > q) --bi;
        while ((bi + 1) * (bi + 2) / 2 <= q) ++bi;
        bj = q - bi * (bi + 1) / 2;
      } else {
        bi = nb, bj = q - nb * (nb + 1) / 2;       // bi == nb: the single right-hand-side row
      }
      const int i0 = k0 + WNB + bi * WNB, j0 = k0 + WNB + bj * WNB;
      const int ni = bi == nb ? 1 : WNB;
      __syncthreads();
      for (int e = t; e < WNB * WNB; e += LMT) {
        const int r = e / WNB, c = e % WNB;
        sPi[r][c] = r < ni ? __ldcg(A + (size_t)(i0 + r) * LD + k0 + c) : 0.0;
        sPj[r][c] = __ldcg(A + (size_t)(j0 + r) * LD + k0 + c);
      }
      __syncthreads();
      for (int e = t; e < WNB * WNB; e += LMT) {
        const int r = e / WNB, c = e % WNB;
        if (r >= ni || (bi == bj && c > r)) continue;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int b = 0; b < WNB; b += 2) s0 = fma(sPi[r][b], sPj[c][b], s0), s1 = fma(sPi[r][b + 1], sPj[c][b + 1], s1);
        A[(size_t)(i0 + r) * LD + j0 + c] -= s0 + s1;
      }
    }
    grid.sync();
    WC_TOCK(c_upd);
  }
#ifdef WC_LM_TIMING
  tk[3] = clock64();
#endif
  // ---- the other CTAs clear the normal-equation buffer the next linearisation accumulates into, CTA 0 finishes the step
  if (cta != 0) {
    if (zero_next) {
      const int nbuf = 1 - cur;
      double2*  Hz   = reinterpret_cast<double2*>(B.H[nbuf]);
      for (size_t i = (size_t)(cta - 1) * LMT + t; i < (size_t)N * N / 2; i += (size_t)(ncta - 1) * LMT) Hz[i] = make_double2(0.0, 0.0);
      if (cta == 1) {
        for (int i = t; i < N; i += LMT) B.g[nbuf][i] = 0.0;
        if (t == 0) *B.cost[nbuf] = 0.0;
      }
    }
    return;
  }
  // blocked backward substitution L^T x = z (z = row Dp), last block first; y = -x scattered to the reduced columns
  double* y    = B.step;
  double* zrow = A + (size_t)Dp * LD;
  for (int c = t; c < D; c += LMT) y[c] = 0.0;
  __syncthreads();
  if (ok) {
    for (int kb = nblk - 1; kb >= 0; --kb) {
      const int k0 = kb * WNB;
      // z[k0 + c] -= sum_{i >= k0 + WNB} L[i][k0 + c] x_i : warp w sums rows i = k0 + WNB + w, + 16, ... for all 32 columns
      double part = 0.0, part2 = 0.0;
      int    i    = k0 + WNB + warp;
      for (; i + LMT / 32 < Dp; i += 2 * (LMT / 32)) {  // two independent loads in flight
        const double l0 = __ldcg(A + (size_t)i * LD + k0 + lane), l1 = __ldcg(A + (size_t)(i + LMT / 32) * LD + k0 + lane);
        part = fma(l0, zrow[i], part), part2 = fma(l1, zrow[i + LMT / 32], part2);
      }
      if (i < Dp) part = fma(__ldcg(A + (size_t)i * LD + k0 + lane), zrow[i], part);
      sPi[warp][lane] = part + part2;  // LMT / 32 = 16 partial rows
      for (int e = t; e < WNB * WNB; e += LMT) {
        const int r = e / WNB, c = e % WNB;
        sL[r][c] = c <= r ? A[(size_t)(k0 + r) * LD + k0 + c] : 0.0;
      }
      __syncthreads();
      if (warp == 0) {
        double z = zrow[k0 + lane];
#pragma unroll
        for (int w = 0; w < LMT / 32; ++w) z -= sPi[w][lane];
        // upper-triangular solve L_kk^T x = z: column lane, last unknown first (reciprocal pivots: no division in the chain)
        const double rp = 1.0 / sL[lane][lane];
#pragma unroll
        for (int j = WNB - 1; j >= 0; --j) {
          const double xj = __shfl_sync(0xffffffffu, z * rp, j);
          if (lane == j) z = xj;
          else if (lane < j) z = fma(-sL[j][lane], xj, z);
        }
        zrow[k0 + lane] = z;  // x overwrites z
        if (k0 + lane < Da) y[B.act[k0 + lane]] = -z;
      }
      __syncthreads();
    }
  }
  __syncthreads();
#ifdef WC_LM_TIMING
  tk[4] = clock64();
#endif
  bool   valid = ok;
  double part = 0.0, bad = 0.0;
  if (valid)
    for (int c = t; c < D; c += LMT) {
      const double sq = sqrt(B.diag[c] / radius);
      part += 0.5 * y[c] * (sq * sq * y[c] - g[amb_of(c, ff)] * B.scale[c]);  // inactive columns: y = 0
      if (!isfinite(y[c])) bad = 1.0;
    }
  const double mcc  = block_sum(part, red);
  const double nbad = block_sum(bad, red);
  valid             = valid && nbad == 0.0 && mcc > 0.0;
  double sn = 0.0;
  for (int i = t; i < N; i += LMT) {
    const int    c = col_of(i, ff);
    const double d = (valid && c >= 0) ? y[c] * B.scale[c] : 0.0;
    B.xc[i]        = B.x[i] + d;
    sn += d * d;
  }
  sn = block_sum(sn, red);
  if (t == 0) {
    sst.iteration += 1;
    sst.last_successful   = 0;
    sst.reuse_diagonal    = 1;
    sst.step_valid        = valid ? 1 : 0;
    sst.pending           = 1;
    sst.model_cost_change = mcc;
    sst.step_norm         = sqrt(sn);
    const int it          = sst.iteration < WC_MAX_ITER_LOG ? sst.iteration : WC_MAX_ITER_LOG - 1;
    sst.iter_radius[it]   = radius;
#ifdef WC_LM_TIMING
    tk[5] = clock64();
    if (sst.iteration == 3)
      printf("lm_step_wide cycles: decide+sync %lld build %lld chol %lld (factor %lld, panel+sync %lld, update+sync %lld) back %lld finish %lld total %lld, Da %d\n",
             tk[1] - tk[0], tk[2] - tk[1], tk[3] - tk[2], c_fac, c_pan, c_upd, tk[4] - tk[3], tk[5] - tk[4], tk[5] - tk[0], Da);
#endif
  }
  __syncthreads();
  for (int k = t; k < (int)(sizeof(LMState) / 8); k += LMT)
    reinterpret_cast<unsigned long long*>(gst)[k] = reinterpret_cast<const unsigned long long*>(&sst)[k];
}

// mirror the lower triangle (what the linearisation accumulates) into the upper one, for callers that want J^T J
__global__ void symmetrize_lower(double* __restrict__ H, int N) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)N * N) return;
  const int r = (int)(i / N), c = (int)(i % N);
  if (c > r) H[i] = H[(size_t)c * N + r];
}


// ---------------------------------------------------------------------------------------------- outputs (8f rank 4)
// residual after the Cauchy corrector of every packed lidar record, in record order; unary (fixed-window) flag
__global__ void lidar_residuals(const double* __restrict__ rec, int n, int stride, const double* __restrict__ x, int jac_mode, double cb,
                                double cc, double* __restrict__ out, unsigned char* __restrict__ is_fix) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double J[24], r, cst;
  int    b1l, b2l, bk;
  lidar_eval(rec, (size_t)stride, i, x, jac_mode, cb, cc, J, r, cst, b1l, b2l, bk);
  out[i]    = r;
  is_fix[i] = b1l < 0;
}

// the 12 residuals of every IMU triplet inside the sample span (BuildImuResiduals, lidar_odometry.cc:319-363), one thread
// per triplet; out[12 * slot ..], slot = rank of the triplet among the valid ones (they form one contiguous range)
__global__ void imu_residuals(ImuArgs a, const double* __restrict__ x, int first_valid, double* __restrict__ out, int* __restrict__ err) {
  const int i = first_valid + blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 2 >= a.n_imu) return;
  const wc_imu_state &i1 = a.imu[i], &i2 = a.imu[i + 1], &i3 = a.imu[i + 2];
  if (i1.timestamp < a.ts[0] || i3.timestamp > a.ts[a.K - 1]) return;
  const int sp2  = upper_bound_ts(a.ts, a.K, i1.timestamp);
  const int mode = (sp2 == a.K - 1) ? 1 : 0;
  const double tsv[3] = {a.ts[sp2 - 1], a.ts[sp2], mode == 0 ? a.ts[sp2 + 1] : DBL_MAX};
  const double* xs[3] = {x + 12 * (sp2 - 1), x + 12 * sp2, x + 12 * (mode == 0 ? sp2 + 1 : sp2)};
  StateCorr    c[3];
  const double tt[3] = {i1.timestamp, i2.timestamp, i3.timestamp};
  for (int s = 0; s < 3; ++s) {
    const double t = tt[s];
    int          l = 0;
    if (mode == 0) {
      const bool in12 = t >= tsv[0] && t < tsv[1], in23 = t >= tsv[1] && t <= tsv[2];
      if (!(in12 || in23)) *err = WC_EOUT_OF_SPAN;
      l = in12 ? 0 : 1;
    } else if (!(t >= tsv[0] && t <= tsv[1])) {
      *err = WC_EOUT_OF_SPAN;
    }
    state_corr(xs[l], xs[l + 1], (t - tsv[l]) / (tsv[l + 1] - tsv[l]), c[s]);
  }
  const Q4 R1 = ldq(i1.rot), R2 = ldq(i2.rot);
  const Q4 E1R1 = Exp(c[0].r) * R1;
  const V3 gyr_est = Log((conj(E1R1) * Exp(c[1].r)) * R2) / a.dt;
  const V3 acc_est = ((c[2].t + ld3(i3.pos)) + (c[0].t + ld3(i1.pos)) - 2.0 * (c[1].t + ld3(i2.pos))) / (a.dt * a.dt);
  const V3 rg  = a.wg * ((ld3(i1.gyr) + ld3(i2.gyr)) / 2.0 - gyr_est - c[0].bg);
  const V3 ra  = a.wa * (E1R1 * (ld3(i1.acc) - c[0].ba) - acc_est + ld3(a.grav));
  const V3 rbg = a.wbg * (c[0].bg - c[1].bg);
  const V3 rba = a.wba * (c[0].ba - c[1].ba);
  double*  o   = out + 12 * (size_t)(i - first_valid);
  o[0] = rg.x, o[1] = rg.y, o[2] = rg.z, o[3] = ra.x, o[4] = ra.y, o[5] = ra.z;
  o[6] = rbg.x, o[7] = rbg.y, o[8] = rbg.z, o[9] = rba.x, o[10] = rba.y, o[11] = rba.z;
}

// PubSurfels (surfel_extraction.cc:360-417) without the ROS message: marker pose / scale / colour per surfel
__global__ void surfel_markers(const wc_surfel* __restrict__ s, int n, wc_marker* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Q4 q = ldq(s[i].rot);
  const M3 R = ToMatrix(q);
  const M3 cw = (R * ld33(s[i].covariance)) * transpose(R);  // GetCovarianceInWorld (surfel.h:89-91)
  double   ev[3];
  M3       V;
  SymEig3(cw.m[0][0], cw.m[1][0], cw.m[2][0], cw.m[1][1], cw.m[2][1], cw.m[2][2], ev, V);  // lower triangle like Eigen
  // makeRightHanded (:340-358)
  auto unit = [](V3 v) { return v / sqrt(dot(v, v)); };
  V3 c0 = unit(col(V, 0)), c1 = unit(col(V, 1)), c2 = unit(col(V, 2));
  if (dot(cross(c0, c1), c2) < 0) {
    const V3 t = c0;
    c0 = c1, c1 = t;
    const double e = ev[0];
    ev[0] = ev[1], ev[1] = e;
  }
  const double m[3][3] = {{c0.x, c1.x, c2.x}, {c0.y, c1.y, c2.y}, {c0.z, c1.z, c2.z}};
  // Eigen::Quaterniond(rot): trace branch, else the largest diagonal element
  double qq[4];  // x y z w
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0) {
    t     = sqrt(t + 1.0);
    qq[3] = 0.5 * t;
    t     = 0.5 / t;
    qq[0] = (m[2][1] - m[1][2]) * t, qq[1] = (m[0][2] - m[2][0]) * t, qq[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int a = 0;
    if (m[1][1] > m[0][0]) a = 1;
    if (m[2][2] > m[a][a]) a = 2;
    const int b = (a + 1) % 3, c = (b + 1) % 3;
    t     = sqrt(m[a][a] - m[b][b] - m[c][c] + 1.0);
    qq[a] = 0.5 * t;
    t     = 0.5 / t;
    qq[3] = (m[c][b] - m[b][c]) * t;
    qq[b] = (m[b][a] + m[a][b]) * t;
    qq[c] = (m[c][a] + m[a][c]) * t;
  }
  const V3 cen = s[i].is_in_body_frame ? q * ld3(s[i].center) + ld3(s[i].pos) : ld3(s[i].center);  // GetCenterInWorld
  const V3 nw  = s[i].is_in_body_frame ? q * ld3(s[i].norm) : ld3(s[i].norm);
  wc_marker k;
  k.position[0] = cen.x, k.position[1] = cen.y, k.position[2] = cen.z;
  for (int j = 0; j < 4; ++j) k.orientation[j] = qq[j];
  for (int j = 0; j < 3; ++j) k.scale[j] = 3.0 * sqrt(ev[j]);
  k.color[0] = (float)((nw.x + 1) / 2), k.color[1] = (float)((nw.y + 1) / 2), k.color[2] = (float)((nw.z + 1) / 2), k.color[3] = 1.f;
  out[i] = k;
}

__global__ void extract_ts(const wc_sample_state* __restrict__ s, int K, double* __restrict__ ts, double* __restrict__ x) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  ts[k] = s[k].timestamp;
  for (int j = 0; j < 12; ++j) x[12 * k + j] = s[k].data_cor[j];
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host
struct wc_solve_mem {
  double*  ts;
  double*  tmp;
  double*  rec;
  float*   rec32;        // fp32 copy of rec (allocated and filled on the first reduced-precision solve of a window)
  int      rec32_valid;
  int*     bucket;
  int*     hist;
  int*     off;
  int*     cursor;
  int      stride;
  int      nb_cap;
  double*  Hbuf[2];
  double*  gbuf[2];
  double*  cost;  // 2 doubles
  double*  scale;
  double*  diag;
  double*  step;
  double*  A;
  int      Ncap;
  int*     act;
  LMState* st;
  LMState* h_st;
  double*  h_x;
  int*     d_perr;  // pack error word
  int*     h_perr;  // pinned
  double   grav[3];
  wc_surfel*  mk_in;      // surfel staging of wc_surfel_markers (grown on demand)
  size_t      mk_cap;
  void*       rec_markers(wc_ctx* c, size_t n) {
    if (n > mk_cap) {
      if (mk_in) cudaFree(mk_in);
      mk_in = nullptr, mk_cap = 0;
      if (cudaMalloc(&mk_in, n * sizeof(wc_surfel)) != cudaSuccess) return nullptr;
      mk_cap = n;
    }
    (void)c;
    return mk_in;
  }
  cudaEvent_t lin_ev[2 * WC_MAX_ITER_LOG + 4];  // begin / end of the linearisation passes of one solve
  int         n_lin_ev;
  int         lin_timing_off;  // no event records while a batch is being captured into a graph
  cudaGraphExec_t gexec;       // one batch of LM iterations, kept across solves and updated in place
};

static wc_status solve_alloc(wc_ctx* c) {
  if (c->d_lm) return WC_OK;
  wc_solve_mem* m = (wc_solve_mem*)calloc(1, sizeof(wc_solve_mem));
  c->d_lm         = m;
  const size_t ns = (size_t)c->prm.max_surfels, nc = (size_t)c->prm.max_corrs, K = (size_t)c->prm.max_samples;
  const size_t N  = 12 * K;
  m->stride       = (int)((nc + 255) & ~(size_t)255);
  m->nb_cap       = (int)(K * K);
  m->Ncap         = (int)N;
  WC_CUDA(c, cudaMalloc(&c->d_sld, ns * sizeof(wc_surfel)));
  WC_CUDA(c, cudaMalloc(&c->d_fix, ns * sizeof(wc_surfel)));
  WC_CUDA(c, cudaMalloc(&c->d_sld_corr, nc * sizeof(wc_corr_idx)));
  WC_CUDA(c, cudaMalloc(&c->d_fix_corr, nc * sizeof(wc_corr_idx)));
  WC_CUDA(c, cudaMalloc(&c->d_imu, (size_t)c->prm.max_imu_states * sizeof(wc_imu_state)));
  WC_CUDA(c, cudaMalloc(&c->d_samples, K * sizeof(wc_sample_state)));
  WC_CUDA(c, cudaMalloc(&m->ts, K * 8));
  WC_CUDA(c, cudaMalloc(&m->tmp, (size_t)REC_COLS * m->stride * 8));
  WC_CUDA(c, cudaMalloc(&m->rec, (size_t)REC_COLS * m->stride * 8));
  WC_CUDA(c, cudaMalloc(&m->bucket, (size_t)m->stride * 4));
  WC_CUDA(c, cudaMalloc(&m->hist, (size_t)m->nb_cap * 4));
  WC_CUDA(c, cudaMalloc(&m->off, (size_t)m->nb_cap * 4));
  WC_CUDA(c, cudaMalloc(&m->cursor, (size_t)m->nb_cap * 4));
  for (int b = 0; b < 2; ++b) {
    WC_CUDA(c, cudaMalloc(&m->Hbuf[b], N * N * 8));
    WC_CUDA(c, cudaMalloc(&m->gbuf[b], N * 8));
  }
  WC_CUDA(c, cudaMalloc(&m->cost, 16));
  WC_CUDA(c, cudaMalloc(&m->scale, N * 8));
  WC_CUDA(c, cudaMalloc(&m->diag, N * 8));
  WC_CUDA(c, cudaMalloc(&m->step, N * 8));
  WC_CUDA(c, cudaMalloc(&m->A, ((N + 40) * (N + 40) + N + 8) * 8));
  WC_CUDA(c, cudaMalloc(&m->act, (N + 1) * 4));
  WC_CUDA(c, cudaMalloc(&c->d_x, N * 8));
  WC_CUDA(c, cudaMalloc(&c->d_xc, N * 8));
  WC_CUDA(c, cudaMalloc(&c->d_x0, N * 8));
  WC_CUDA(c, cudaMalloc(&m->st, sizeof(LMState)));
  WC_CUDA(c, cudaMallocHost(&m->h_st, sizeof(LMState)));
  WC_CUDA(c, cudaMallocHost(&m->h_x, N * 8));
  WC_CUDA(c, cudaMalloc(&m->d_perr, 16));
  WC_CUDA(c, cudaMallocHost(&m->h_perr, 16));
  m->h_perr[0] = 0;
  WC_CUDA(c, cudaFuncSetAttribute((window_linearize<WC_PREC_F64, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, LIN_SMEM));
  WC_CUDA(c, cudaFuncSetAttribute((window_linearize<WC_PREC_MIXED, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, LIN_SMEM_R32));
  WC_CUDA(c, cudaFuncSetAttribute((window_linearize<WC_PREC_F32, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, LIN_SMEM_R32));
  WC_CUDA(c, cudaFuncSetAttribute((window_linearize<WC_PREC_F64, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, LIN_SMEM_BASE));
  WC_CUDA(c, cudaFuncSetAttribute((window_linearize<WC_PREC_MIXED, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, LIN_SMEM_BASE));
  WC_CUDA(c, cudaFuncSetAttribute((window_linearize<WC_PREC_F32, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, LIN_SMEM_BASE));
  for (int i = 0; i < 2 * WC_MAX_ITER_LOG + 4; ++i) WC_CUDA(c, cudaEventCreate(&m->lin_ev[i]));
  WC_CUDA(c, cudaFuncSetAttribute(lm_step<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return WC_OK;
}

void wc_solve_free(wc_ctx* c) {
  wc_solve_mem* m = (wc_solve_mem*)c->d_lm;
  void* ptrs[] = {c->d_sld, c->d_fix, c->d_sld_corr, c->d_fix_corr, c->d_imu, c->d_samples, c->d_x, c->d_xc, c->d_x0, c->d_fix_tmp, c->d_status};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (!m) return;
  if (m->gexec) cudaGraphExecDestroy(m->gexec);
  for (auto& e : m->lin_ev)
    if (e) cudaEventDestroy(e);
  void* mp[] = {m->mk_in, m->rec32, m->ts, m->tmp, m->rec, m->bucket, m->hist, m->off, m->cursor, m->Hbuf[0], m->Hbuf[1], m->gbuf[0], m->gbuf[1],
                m->cost, m->scale, m->diag, m->step, m->A, m->st, m->act};
  for (void* p : mp)
    if (p) cudaFree(p);
  if (m->h_st) cudaFreeHost(m->h_st);
  if (m->h_x) cudaFreeHost(m->h_x);
  if (m->h_perr) cudaFreeHost(m->h_perr);
  if (m->d_perr) cudaFree(m->d_perr);
  free(m);
  c->d_lm = nullptr;
}

static SolveBufs make_bufs(wc_ctx* c, int fix_first) {
  wc_solve_mem* m = (wc_solve_mem*)c->d_lm;
  SolveBufs     B;
  for (int b = 0; b < 2; ++b) B.H[b] = m->Hbuf[b], B.g[b] = m->gbuf[b], B.cost[b] = m->cost + b;
  B.x = c->d_x, B.xc = c->d_xc, B.scale = m->scale, B.diag = m->diag, B.step = m->step, B.A = m->A, B.st = m->st, B.act = m->act;
  B.N = (int)(12 * c->K), B.fix_first = fix_first;
  return B;
}

// Device-side window preparation: sample timestamps / start point, factor construction for this rank's slice of the
// correspondence list, bucketing by interval pair.  Expects d_sld/d_fix/d_sld_corr/d_fix_corr/d_imu/d_samples and the
// counts in the ctx.
static wc_status pack_error(wc_ctx* c, int e) {
  if (e == WC_EINVAL) WC_FAIL(c, WC_EINVAL, "correspondence index out of range");
  if (e == WC_EINVAL_TIME_ORDER) WC_FAIL(c, WC_EINVAL_TIME_ORDER, "correspondence with timestamp(s1) >= timestamp(s2)");
  if (e == WC_EOUT_OF_SPAN) WC_FAIL(c, WC_EOUT_OF_SPAN, "surfel timestamp outside the sample-state span");
  return WC_OK;
}

// defer != 0: no host synchronisation; the error word is read back by the solve's first host check
wc_status wc_window_prepare_device(wc_ctx* c, int defer) {
  wc_solve_mem* m  = (wc_solve_mem*)c->d_lm;
  cudaStream_t  st = c->stream;
  const size_t  K = c->K, n_sld = c->n_sld, n_fix = c->n_fix, n_sld_corr = c->n_sld_corr, n_fix_corr = c->n_fix_corr;
  if (n_sld_corr + n_fix_corr > (size_t)c->prm.max_corrs) WC_FAIL(c, WC_ECAPACITY, "too many correspondences");
  { ++c->n_launches; extract_ts<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(c->d_samples, (int)K, m->ts, c->d_x0); }
  // factor construction for this rank's slice of the correspondence list, bucketed by interval pair
  const int C  = (int)(n_sld_corr + n_fix_corr);
  const int c0 = (int)((long long)C * c->rank / c->world), c1 = (int)((long long)C * (c->rank + 1) / c->world);
  c->n_rec     = (size_t)(c1 - c0);
  m->rec32_valid = 0;
  const int nb = (int)(K * K);
  WC_CUDA(c, cudaMemsetAsync(m->d_perr, 0, 16, st));
  WC_CUDA(c, cudaMemsetAsync(m->hist, 0, (size_t)nb * 4, st));
  WC_CUDA(c, cudaMemsetAsync(m->cursor, 0, (size_t)nb * 4, st));
  if (c->n_rec) {
    PackArgs a;
    a.sld = c->d_sld, a.fix = c->d_fix, a.sld_corr = c->d_sld_corr, a.fix_corr = c->d_fix_corr;
    a.n_sld = (int)n_sld, a.n_fix = (int)n_fix, a.n_sld_corr = (int)n_sld_corr, a.n_fix_corr = (int)n_fix_corr;
    a.c0 = c0, a.c1 = c1, a.ts = m->ts, a.K = (int)K, a.weight_floor = c->prm.weight_floor;
    a.tmp = m->tmp, a.stride = m->stride, a.bucket = m->bucket, a.hist = m->hist, a.err = m->d_perr;
    const unsigned grid = (unsigned)((c->n_rec + 255) / 256);
    { ++c->n_launches; corr_pack<<<grid, 256, 0, st>>>(a); }
    { ++c->n_launches; bucket_scan<<<1, 1024, 0, st>>>(m->hist, nb, m->off); }
    { ++c->n_launches; bucket_scatter<<<grid, 256, 0, st>>>(m->tmp, m->bucket, (int)c->n_rec, m->stride, m->off, m->cursor, m->rec); }
  }
  WC_CUDA(c, cudaMemcpyAsync(m->h_perr, m->d_perr, 4, cudaMemcpyDeviceToHost, st));
  if (defer) return WC_OK;
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  return pack_error(c, m->h_perr[0]);
}

extern "C" wc_status wc_window_upload(wc_ctx* c, const wc_surfel* sld, size_t n_sld, const wc_surfel* fix, size_t n_fix,
                                      const wc_corr_idx* sld_corr, size_t n_sld_corr, const wc_corr_idx* fix_corr,
                                      size_t n_fix_corr, const wc_imu_state* imu, size_t n_imu,
                                      const wc_sample_state* samples, size_t K) {
  if (!c || !samples || K < 2 || (n_sld && !sld) || (n_fix && !fix) || (n_sld_corr && !sld_corr) || (n_fix_corr && !fix_corr) ||
      (n_imu && !imu))
    return WC_EINVAL;
  wc_status s = solve_alloc(c);
  if (s) return s;
  wc_solve_mem* m = (wc_solve_mem*)c->d_lm;
  if (n_sld > (size_t)c->prm.max_surfels || n_fix > (size_t)c->prm.max_surfels) WC_FAIL(c, WC_ECAPACITY, "too many surfels");
  if (n_sld_corr + n_fix_corr > (size_t)c->prm.max_corrs) WC_FAIL(c, WC_ECAPACITY, "too many correspondences");
  if (K > (size_t)c->prm.max_samples) WC_FAIL(c, WC_ECAPACITY, "too many sample states");
  if (n_imu > (size_t)c->prm.max_imu_states) WC_FAIL(c, WC_ECAPACITY, "too many IMU states");
  cudaStream_t st = c->stream;
  c->n_sld = n_sld, c->n_fix = n_fix, c->n_sld_corr = n_sld_corr, c->n_fix_corr = n_fix_corr, c->n_imu = n_imu, c->K = K;
  c->n_sld_prev = 0;
  if (n_sld) WC_CUDA(c, cudaMemcpyAsync(c->d_sld, sld, n_sld * sizeof(wc_surfel), cudaMemcpyHostToDevice, st));
  if (n_fix) WC_CUDA(c, cudaMemcpyAsync(c->d_fix, fix, n_fix * sizeof(wc_surfel), cudaMemcpyHostToDevice, st));
  if (n_sld_corr) WC_CUDA(c, cudaMemcpyAsync(c->d_sld_corr, sld_corr, n_sld_corr * sizeof(wc_corr_idx), cudaMemcpyHostToDevice, st));
  if (n_fix_corr) WC_CUDA(c, cudaMemcpyAsync(c->d_fix_corr, fix_corr, n_fix_corr * sizeof(wc_corr_idx), cudaMemcpyHostToDevice, st));
  if (n_imu) WC_CUDA(c, cudaMemcpyAsync(c->d_imu, imu, n_imu * sizeof(wc_imu_state), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemcpyAsync(c->d_samples, samples, K * sizeof(wc_sample_state), cudaMemcpyHostToDevice, st));
  for (int k = 0; k < 3; ++k) m->grav[k] = samples[K - 1].grav[k];  // lidar_odometry.cc:341,355
  c->n_imu_blocks = 0, c->first_imu_block = 0;  // BuildImuResiduals' blocks (:320-329): one contiguous range of triplets
  for (size_t i = 0; i + 2 < n_imu; ++i)
    if (imu[i].timestamp >= samples[0].timestamp && imu[i + 2].timestamp <= samples[K - 1].timestamp) {
      if (c->n_imu_blocks == 0) c->first_imu_block = (int)i;
      ++c->n_imu_blocks;
    }
  return wc_window_prepare_device(c, 0);
}

// Upload of everything a device-resident window pass needs besides the sweep itself: IMU states, sample states and
// the (body-frame) fixed-window surfels.  The sliding-window surfels and both correspondence lists are produced on
// the device by the pass.
wc_status wc_window_upload_aux(wc_ctx* c, const wc_imu_state* imu, size_t n_imu, const wc_sample_state* samples, size_t K,
                               const wc_surfel* fix, size_t n_fix) {
  if (!c || !samples || K < 2 || (n_imu && !imu) || (n_fix && !fix)) return WC_EINVAL;
  wc_status s = solve_alloc(c);
  if (s) return s;
  wc_solve_mem* m = (wc_solve_mem*)c->d_lm;
  if (n_fix > (size_t)c->prm.max_surfels) WC_FAIL(c, WC_ECAPACITY, "too many surfels");
  if (K > (size_t)c->prm.max_samples) WC_FAIL(c, WC_ECAPACITY, "too many sample states");
  if (n_imu > (size_t)c->prm.max_imu_states) WC_FAIL(c, WC_ECAPACITY, "too many IMU states");
  cudaStream_t st = c->stream;
  c->n_fix = n_fix, c->n_imu = n_imu, c->K = K;
  if (n_fix) WC_CUDA(c, cudaMemcpyAsync(c->d_fix, fix, n_fix * sizeof(wc_surfel), cudaMemcpyHostToDevice, st));
  if (n_imu) WC_CUDA(c, cudaMemcpyAsync(c->d_imu, imu, n_imu * sizeof(wc_imu_state), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemcpyAsync(c->d_samples, samples, K * sizeof(wc_sample_state), cudaMemcpyHostToDevice, st));
  for (int k = 0; k < 3; ++k) m->grav[k] = samples[K - 1].grav[k];  // lidar_odometry.cc:341,355
  c->n_imu_blocks = 0, c->first_imu_block = 0;
  for (size_t i = 0; i + 2 < n_imu; ++i)
    if (imu[i].timestamp >= samples[0].timestamp && imu[i + 2].timestamp <= samples[K - 1].timestamp) {
      if (c->n_imu_blocks == 0) c->first_imu_block = (int)i;
      ++c->n_imu_blocks;
    }
  WC_CUDA(c, cudaStreamSynchronize(st));
  return WC_OK;
}


// kernel launch with (optionally) the programmatic-dependent-launch attribute
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// enqueue one linearisation (lidar + IMU [+ exchange]) on the ctx stream
static wc_status enqueue_linearize(wc_ctx* c, const SolveBufs& B_in, const wc_solve_opts* o, int at_candidate, int zeroed) {
  wc_solve_mem* m  = (wc_solve_mem*)c->d_lm;
  cudaStream_t  st = c->stream;
  SolveBufs     B  = B_in;
  if (c->world > 1) {
    // sharded: accumulate into this rank's exchange partial; comm_allreduce sums the ranks into the real buffers
    double *pH, *pg, *pc;
    wc_comm_partial_views(c, &pH, &pg, &pc);
    B.H[0] = B.H[1] = pH, B.g[0] = B.g[1] = pg, B.cost[0] = B.cost[1] = pc;
    zeroed = 1;  // cleared by wc_comm_begin_solve / by the reduction two epochs back
  }
  if (!zeroed) { ++c->n_launches; zero_buffers<<<64, 256, 0, st>>>(B, at_candidate); }
  const int prec = o->precision;
  if (prec != WC_PREC_F64 && prec != WC_PREC_MIXED && prec != WC_PREC_F32) WC_FAIL(c, WC_EINVAL, "unknown precision mode %d", prec);
  if (prec != WC_PREC_F64 && !m->rec32_valid) {
    if (!m->rec32) WC_CUDA(c, cudaMalloc(&m->rec32, (size_t)REC_COLS * m->stride * 4));
    if (c->K * c->K >= (1u << 24) || c->K >= 32768) WC_FAIL(c, WC_ECAPACITY, "too many sample states for the fp32 record format");
    if (c->n_rec) { ++c->n_launches; rec_to_f32<<<(unsigned)((c->n_rec + 255) / 256), 256, 0, st>>>(m->rec, (int)c->n_rec, m->stride, m->rec32); }
    m->rec32_valid = 1;
  }
  LinArgs a;
  a.rec = m->rec, a.rec32 = m->rec32, a.stride = m->stride, a.n_rec = (int)c->n_rec, a.B = B, a.at_candidate = at_candidate;
  a.jac_mode = o->jacobian_mode, a.cauchy_b = c->prm.cauchy_a * c->prm.cauchy_a, a.cauchy_c = 1.0 / a.cauchy_b;
  const int ntiles  = (int)((c->n_rec + LT - 1) / LT);
  // Small windows (a few waves of tiles): one tile per CTA, records straight from global memory.  Large windows:
  // resident CTAs with contiguous chunks of tiles, the record tiles staged by TMA bulk copies so that the copy of a CTA's
  // next tile overlaps the evaluation of its current one.  (Measured at C3, 707 tiles: staging a CTA's only tile costs
  // 6 us per launch — there is nothing to overlap with; at 2 M records it gains 12 %.)
  static const int force_staged = getenv("WC_LIN_STAGED") ? atoi(getenv("WC_LIN_STAGED")) : -1;  // test hook
  const bool staged  = force_staged >= 0 ? force_staged != 0 : ntiles > 8 * c->num_sms;
  const int  n_lidar = !staged ? ntiles : (ntiles <= 3 * c->num_sms ? ntiles : 3 * c->num_sms);
  ImuArgs b;
  memset(&b, 0, sizeof(b));
  int n_imu_cta = 0;
  static const int dbg_no_imu = getenv("WC_DBG_NO_IMU") != nullptr;  // timing experiment only
  if (o->use_imu_factors && c->n_imu >= 3 && c->rank == 0 && !dbg_no_imu) {
    b.imu = c->d_imu, b.n_imu = (int)c->n_imu, b.ts = m->ts, b.K = (int)c->K, b.B = B, b.at_candidate = at_candidate;
    b.wg = c->prm.weight_gyr, b.wa = c->prm.weight_acc, b.wbg = c->prm.weight_bg, b.wba = c->prm.weight_ba;
    b.dt = 1.0 / c->prm.imu_rate;
    for (int k = 0; k < 3; ++k) b.grav[k] = m->grav[k];
    n_imu_cta = (int)((c->n_imu - 2 + IMU_WARPS - 1) / IMU_WARPS);
  }
  // per-pass event records (summary.gpu_ms_linearize) cost ~5 us of stream time each: only on request (WC_TIME_PASSES=1)
  static const int time_passes = getenv("WC_TIME_PASSES") ? atoi(getenv("WC_TIME_PASSES")) : 0;
  const bool timed = time_passes && !m->lin_timing_off && m->n_lin_ev + 2 <= 2 * WC_MAX_ITER_LOG + 4;
  if (timed) WC_CUDA(c, cudaEventRecord(m->lin_ev[m->n_lin_ev++], st));
  if (n_lidar + n_imu_cta > 0) {
    ++c->n_launches;
    const unsigned grid = (unsigned)(n_lidar + n_imu_cta);
    // inside the LM loop (at_candidate) on one GPU the launch may overlap the tail of the LM step before it (PDL)
    static const int use_pdl = getenv("WC_LM_PDL") ? atoi(getenv("WC_LM_PDL")) : 0;  // measured at C3: solve 2.23 ms without, 2.33 ms with
    const bool       pdl     = use_pdl && at_candidate && c->world == 1;
    const dim3       g3(grid), b3(LT);
    cudaError_t      le;
    if (staged) {
      if (prec == WC_PREC_F64) le = launch_pdl(window_linearize<WC_PREC_F64, true>, g3, b3, LIN_SMEM, st, pdl, a, b, n_lidar, n_imu_cta);
      else if (prec == WC_PREC_MIXED) le = launch_pdl(window_linearize<WC_PREC_MIXED, true>, g3, b3, LIN_SMEM_R32, st, pdl, a, b, n_lidar, n_imu_cta);
      else le = launch_pdl(window_linearize<WC_PREC_F32, true>, g3, b3, LIN_SMEM_R32, st, pdl, a, b, n_lidar, n_imu_cta);
    } else {
      if (prec == WC_PREC_F64) le = launch_pdl(window_linearize<WC_PREC_F64, false>, g3, b3, LIN_SMEM_BASE, st, pdl, a, b, n_lidar, n_imu_cta);
      else if (prec == WC_PREC_MIXED) le = launch_pdl(window_linearize<WC_PREC_MIXED, false>, g3, b3, LIN_SMEM_BASE, st, pdl, a, b, n_lidar, n_imu_cta);
      else le = launch_pdl(window_linearize<WC_PREC_F32, false>, g3, b3, LIN_SMEM_BASE, st, pdl, a, b, n_lidar, n_imu_cta);
    }
    WC_CUDA(c, le);
  }
  if (timed) WC_CUDA(c, cudaEventRecord(m->lin_ev[m->n_lin_ev++], st));
  WC_CUDA(c, cudaGetLastError());
  return WC_OK;
}

static void fill_summary(const wc_ctx* c, const LMState* s, wc_solve_summary* sum) {
  memset(sum, 0, sizeof(*sum));
  sum->initial_cost = s->initial_cost, sum->final_cost = s->x_cost;
  sum->num_iterations = s->iteration, sum->num_successful_steps = s->num_successful;
  sum->num_unsuccessful_steps = s->num_unsuccessful, sum->termination = s->termination;
  sum->num_residual_blocks_sld = (int)c->n_sld_corr, sum->num_residual_blocks_fix = (int)c->n_fix_corr;
  sum->num_linearizations = s->num_linearizations;
  sum->num_residual_blocks_imu = c->n_imu_blocks;
  for (int i = 0; i < WC_MAX_ITER_LOG; ++i)
    sum->iter_cost[i] = s->iter_cost[i], sum->iter_radius[i] = s->iter_radius[i], sum->iter_accepted[i] = s->iter_accepted[i];
}

wc_status wc_points_prefetch_deferred(wc_ctx* c);  // wc_extract.cu

extern "C" wc_status wc_window_solve_resident(wc_ctx* c, const wc_solve_opts* opts, wc_solve_summary* summary,
                                              double* data_cor_out) {
  if (!c || !c->d_lm || c->K < 2) return WC_EINVAL;
  wc_solve_opts o;
  if (opts) o = *opts; else wc_default_solve_opts(&o);
  wc_solve_mem* m  = (wc_solve_mem*)c->d_lm;
  cudaStream_t  st = c->stream;
  const int     N  = (int)(12 * c->K);
  SolveBufs     B  = make_bufs(c, o.fix_first_position ? 1 : 0);
  const int     D  = o.fix_first_position ? N - 3 : N;
  const int     Dp          = (D + CB - 1) / CB * CB, LD = ((Dp >> 1) & 1) ? Dp : Dp + 2;
  const size_t  a_bytes     = ((size_t)(Dp + 4) * LD + 3 * (size_t)Dp) * 8;  // A, reciprocal pivots, scaling, LM diagonal
  const int     a_in_smem   = a_bytes <= 200 * 1024;
  const size_t  smem        = a_in_smem ? a_bytes : 0;
  static const int no_wide  = getenv("WC_LM_NO_WIDE") != nullptr;  // test hook: the single-CTA global-memory step
  const bool    use_wide    = !a_in_smem && !no_wide;
  WC_CUDA(c, cudaEventRecord(c->ev[4], st));
  WC_CUDA(c, cudaMemsetAsync(m->st, 0, sizeof(LMState), st));
  WC_CUDA(c, cudaMemcpyAsync(c->d_x, c->d_x0, (size_t)N * 8, cudaMemcpyDeviceToDevice, st));
  m->n_lin_ev = 0;
  wc_status s = wc_comm_begin_solve(c);
  if (s) return s;
  if ((s = enqueue_linearize(c, B, &o, 0, 0))) return s;
  if ((s = wc_comm_allreduce(c, 0))) return s;
  { ++c->n_launches; lm_init<<<1, LMT, 0, st>>>(B, o); }
  const int batch = c->lm_batch > 0 ? c->lm_batch : 8;
  static const int dbg_ev = getenv("WC_LM_EVENTS") != nullptr;  // debug: per-launch stream timeline of the LM loop
  static cudaEvent_t evs[256];
  static int evs_init = 0;
  int n_ev = 0;
  if (dbg_ev && !evs_init) { for (auto& e : evs) cudaEventCreate(&e); evs_init = 1; }
  auto enqueue_batch = [&]() -> wc_status {
    for (int b = 0; b < batch; ++b) {
      // decide(previous candidate) + next trust-region step + clear the candidate buffer, then linearise there
      if (dbg_ev && n_ev + 3 < 256) cudaEventRecord(evs[n_ev++], st);
      ++c->n_launches;
      if (a_in_smem) {
        static const int use_pdl = getenv("WC_LM_PDL") ? atoi(getenv("WC_LM_PDL")) : 0;  // measured at C3: solve 2.23 ms without, 2.33 ms with
        WC_CUDA(c, launch_pdl(lm_step<true>, dim3(1), dim3(LMT), smem, st, use_pdl && c->world == 1, B, o, (int)(c->world == 1)));
      } else if (use_wide) {
        int   zn      = c->world == 1;
        void* args[3] = {(void*)&B, (void*)&o, (void*)&zn};
        WC_CUDA(c, cudaLaunchCooperativeKernel((const void*)lm_step_wide, dim3((unsigned)c->num_sms), dim3(LMT), args, 0, st));
      } else {
        lm_step<false><<<1, LMT, 0, st>>>(B, o, c->world == 1);
      }
      if (dbg_ev && n_ev + 3 < 256) cudaEventRecord(evs[n_ev++], st);
      wc_status es;
      if ((es = enqueue_linearize(c, B, &o, 1, 1))) return es;
      if ((es = wc_comm_allreduce(c, 1))) return es;
    }
    return WC_OK;
  };
  // One batch of LM iterations is a fixed sequence of launches (every kernel reads its branch from the device-resident
  // LM state): it is captured once per solve and replayed as a CUDA graph, which shortens the kernel-to-kernel gaps
  // (WC_LM_GRAPH=0 turns that off).
  static const int use_graph = getenv("WC_LM_GRAPH") ? atoi(getenv("WC_LM_GRAPH")) : 1;  // measured at C3: solve 2.54 -> 2.35 ms
  static const int time_passes_g = getenv("WC_TIME_PASSES") ? atoi(getenv("WC_TIME_PASSES")) : 0;
  cudaGraphExec_t  gexec = nullptr;
  if (use_graph && !dbg_ev && !time_passes_g && c->world == 1 && !use_wide) {
    // capture (a few microseconds per launch), then update the executable graph kept from the previous solve in place:
    // same topology, new kernel parameters — instantiating from scratch costs more than the gaps it saves
    cudaGraph_t g = nullptr;
    m->lin_timing_off = 1;
    WC_CUDA(c, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    s = enqueue_batch();
    cudaError_t ce = cudaStreamEndCapture(st, &g);
    m->lin_timing_off = 0;
    if (s) return s;
    WC_CUDA(c, ce);
    if (m->gexec) {
      cudaGraphExecUpdateResultInfo info;
      if (cudaGraphExecUpdate(m->gexec, g, &info) != cudaSuccess) {
        cudaGetLastError();
        cudaGraphExecDestroy(m->gexec);
        m->gexec = nullptr;
      }
    }
    if (!m->gexec) WC_CUDA(c, cudaGraphInstantiate(&m->gexec, g, 0));
    cudaGraphDestroy(g);
    gexec = m->gexec;
  }
  // Consecutive windows take about the same number of iterations: that many batches are enqueued before the first host
  // check of the termination flag (iterations enqueued after the solver terminated return at once).
  int ahead = (c->last_lm_iters + 1 + batch - 1) / batch;
  if (ahead < 1) ahead = 1;
  if (ahead > 8) ahead = 8;
  for (int done = 0, it = 0; !done && it <= o.max_num_iterations + 2 * batch;) {
    for (int r = 0; r < ahead; ++r, it += batch) {
      if (gexec) WC_CUDA(c, cudaGraphLaunch(gexec, st));
      else if ((s = enqueue_batch())) return s;
    }
    ahead = 1;
    // the LM iterations are enqueued and the host is about to wait for them: the sweep announced by
    // wc_points_prefetch(..., WC_PREFETCH_AT_SOLVE) starts crossing PCIe now, beside the solve (no-op when none is pending)
    if ((s = wc_points_prefetch_deferred(c))) return s;
    WC_CUDA(c, cudaMemcpyAsync(m->h_st, m->st, sizeof(LMState), cudaMemcpyDeviceToHost, st));
    WC_CUDA(c, cudaStreamSynchronize(st));
    if ((s = pack_error(c, m->h_perr[0]))) return s;
    done = m->h_st->done;
  }
  c->last_lm_iters = m->h_st->iteration;
  WC_CUDA(c, cudaMemcpyAsync(m->h_x, c->d_x, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaEventRecord(c->ev[5], st));
  WC_CUDA(c, cudaStreamSynchronize(st));
#ifdef WC_LM_TIMING
  lm_stamps_print<<<1, 1, 0, st>>>();
  cudaStreamSynchronize(st);
#endif
  if (dbg_ev) {
    printf("LM timeline (us between consecutive events: step, linearize, step, ...):");
    for (int i = 0; i + 1 < n_ev; ++i) { float ms; cudaEventElapsedTime(&ms, evs[i], evs[i + 1]); printf(" %.1f", ms * 1e3f); }
    printf("\n");
  }
  WC_CUDA(c, cudaGetLastError());
  if (m->h_st->err == WC_EOUT_OF_SPAN) WC_FAIL(c, WC_EOUT_OF_SPAN, "IMU state outside its sample interval (cost_functor.h:367,390)");
  if ((s = wc_comm_check(c))) return s;
  if (summary) {
    fill_summary(c, m->h_st, summary);
    float ms;
    cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]);
    summary->gpu_ms_total = ms;
    // executed passes only: launches enqueued after the solver terminated return at once
    double lin = 0.0;
    for (int i = 0; i + 1 < m->n_lin_ev && i / 2 < m->h_st->num_linearizations; i += 2) {
      cudaEventElapsedTime(&ms, m->lin_ev[i], m->lin_ev[i + 1]);
      lin += ms;
    }
    summary->gpu_ms_linearize = lin;
  }
  if (data_cor_out) memcpy(data_cor_out, m->h_x, (size_t)N * 8);
  if (m->h_st->termination == WC_TERM_FAILURE) WC_FAIL(c, WC_ENUMERIC, "solve failed: non-finite cost or 5 consecutive invalid steps");
  return WC_OK;
}

extern "C" wc_status wc_window_solve(wc_ctx* c, const wc_surfel* sld, size_t n_sld, const wc_surfel* fix, size_t n_fix,
                                     const wc_corr_idx* sld_corr, size_t n_sld_corr, const wc_corr_idx* fix_corr,
                                     size_t n_fix_corr, const wc_imu_state* imu, size_t n_imu, wc_sample_state* samples,
                                     size_t K, const wc_solve_opts* opts, wc_solve_summary* summary) {
  wc_status s = wc_window_upload(c, sld, n_sld, fix, n_fix, sld_corr, n_sld_corr, fix_corr, n_fix_corr, imu, n_imu, samples, K);
  if (s) return s;
  wc_solve_mem* m = (wc_solve_mem*)c->d_lm;
  s               = wc_window_solve_resident(c, opts, summary, nullptr);
  if (s && s != WC_ENUMERIC) return s;
  for (size_t k = 0; k < K; ++k)
    for (int j = 0; j < 12; ++j) samples[k].data_cor[j] = m->h_x[12 * k + j];  // in place, like SampleState::data_cor
  return s;
}

extern "C" wc_status wc_window_evaluate(wc_ctx* c, const wc_surfel* sld, size_t n_sld, const wc_surfel* fix, size_t n_fix,
                                        const wc_corr_idx* sld_corr, size_t n_sld_corr, const wc_corr_idx* fix_corr,
                                        size_t n_fix_corr, const wc_imu_state* imu, size_t n_imu,
                                        const wc_sample_state* samples, size_t K, const wc_solve_opts* opts, double* cost,
                                        double* grad, double* jtj) {
  wc_status s = wc_window_upload(c, sld, n_sld, fix, n_fix, sld_corr, n_sld_corr, fix_corr, n_fix_corr, imu, n_imu, samples, K);
  if (s) return s;
  wc_solve_opts o;
  if (opts) o = *opts; else wc_default_solve_opts(&o);
  wc_solve_mem* m  = (wc_solve_mem*)c->d_lm;
  cudaStream_t  st = c->stream;
  const size_t  N  = 12 * K;
  SolveBufs     B  = make_bufs(c, 0);
  WC_CUDA(c, cudaMemsetAsync(m->st, 0, sizeof(LMState), st));
  WC_CUDA(c, cudaMemcpyAsync(c->d_x, c->d_x0, N * 8, cudaMemcpyDeviceToDevice, st));
  m->n_lin_ev = 0;
  if ((s = wc_comm_begin_solve(c))) return s;
  if ((s = enqueue_linearize(c, B, &o, 0, 0))) return s;
  if ((s = wc_comm_allreduce(c, 0))) return s;
  if (cost) WC_CUDA(c, cudaMemcpyAsync(cost, m->cost, 8, cudaMemcpyDeviceToHost, st));
  if (grad) WC_CUDA(c, cudaMemcpyAsync(grad, m->gbuf[0], N * 8, cudaMemcpyDeviceToHost, st));
  if (jtj) {
    { ++c->n_launches; symmetrize_lower<<<(unsigned)((N * N + 255) / 256), 256, 0, st>>>(m->Hbuf[0], (int)N); }
    WC_CUDA(c, cudaMemcpyAsync(jtj, m->Hbuf[0], N * N * 8, cudaMemcpyDeviceToHost, st));
  }
  WC_CUDA(c, cudaMemcpyAsync(m->h_st, m->st, sizeof(LMState), cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  if (m->h_st->err == WC_EOUT_OF_SPAN) WC_FAIL(c, WC_EOUT_OF_SPAN, "IMU state outside its sample interval");
  return WC_OK;
}

// accessors used by the exchange layer (wc_comm.cu): the packed [H | g | cost] of a buffer
void wc_solve_exchange_views(wc_ctx* c, int which, double** H, double** g, double** cost, int* N, void** state) {
  wc_solve_mem* m = (wc_solve_mem*)c->d_lm;
  (void)which;
  H[0] = m->Hbuf[0], H[1] = m->Hbuf[1], g[0] = m->gbuf[0], g[1] = m->gbuf[1], cost[0] = m->cost, cost[1] = m->cost + 1;
  *N     = (int)(12 * c->K);
  *state = m->st;
}

extern "C" wc_status wc_window_residuals(wc_ctx* c, const wc_solve_opts* opts, const double* data_cor, double* lidar_res,
                                         uint8_t* lidar_is_fix, size_t lidar_cap, size_t* n_lidar, double* imu_res, size_t imu_cap,
                                         size_t* n_imu_blocks) {
  if (!c || !c->d_lm || c->K < 2 || !n_lidar || !n_imu_blocks) return WC_EINVAL;
  wc_solve_opts o;
  if (opts) o = *opts; else wc_default_solve_opts(&o);
  wc_solve_mem* m  = (wc_solve_mem*)c->d_lm;
  cudaStream_t  st = c->stream;
  const size_t  N  = 12 * c->K;
  *n_lidar = c->n_rec, *n_imu_blocks = 0;
  if (c->n_rec > lidar_cap) WC_FAIL(c, WC_ECAPACITY, "%zu lidar residuals exceed the output capacity %zu", c->n_rec, lidar_cap);
  if (c->n_rec && (!lidar_res || !lidar_is_fix)) return WC_EINVAL;
  // evaluation point
  if (data_cor) WC_CUDA(c, cudaMemcpyAsync(c->d_xc, data_cor, N * 8, cudaMemcpyHostToDevice, st));
  else WC_CUDA(c, cudaMemcpyAsync(c->d_xc, c->d_x0, N * 8, cudaMemcpyDeviceToDevice, st));
  // scratch: the unsorted staging columns of the pack stage are free after wc_window_upload
  double*        d_res = m->tmp;
  unsigned char* d_fix = (unsigned char*)(m->tmp + m->stride);
  if (c->n_rec) {
    ++c->n_launches;
    lidar_residuals<<<(unsigned)((c->n_rec + 255) / 256), 256, 0, st>>>(m->rec, (int)c->n_rec, m->stride, c->d_xc, o.jacobian_mode,
                                                                     c->prm.cauchy_a * c->prm.cauchy_a, 1.0 / (c->prm.cauchy_a * c->prm.cauchy_a),
                                                                     d_res, d_fix);
    WC_CUDA(c, cudaMemcpyAsync(lidar_res, d_res, c->n_rec * 8, cudaMemcpyDeviceToHost, st));
    WC_CUDA(c, cudaMemcpyAsync(lidar_is_fix, d_fix, c->n_rec, cudaMemcpyDeviceToHost, st));
  }
  // IMU blocks: the triplets inside the sample span are contiguous; their range comes from the host copy of the counts
  if (o.use_imu_factors && c->n_imu >= 3 && c->n_imu_blocks > 0 && c->rank == 0) {
    if ((size_t)c->n_imu_blocks > imu_cap) WC_FAIL(c, WC_ECAPACITY, "%d IMU blocks exceed the output capacity %zu", c->n_imu_blocks, imu_cap);
    if (!imu_res) return WC_EINVAL;
    if ((size_t)c->n_imu_blocks * 12 > (size_t)(REC_COLS - 2) * m->stride) WC_FAIL(c, WC_ECAPACITY, "IMU residual scratch too small");
    ImuArgs b;
    memset(&b, 0, sizeof(b));
    b.imu = c->d_imu, b.n_imu = (int)c->n_imu, b.ts = m->ts, b.K = (int)c->K;
    b.wg = c->prm.weight_gyr, b.wa = c->prm.weight_acc, b.wbg = c->prm.weight_bg, b.wba = c->prm.weight_ba;
    b.dt = 1.0 / c->prm.imu_rate;
    for (int k = 0; k < 3; ++k) b.grav[k] = m->grav[k];
    double* d_imu_res = m->tmp + 2 * (size_t)m->stride;
    WC_CUDA(c, cudaMemsetAsync(m->st, 0, sizeof(LMState), st));
    ++c->n_launches;
    imu_residuals<<<(unsigned)((c->n_imu_blocks + 127) / 128), 128, 0, st>>>(b, c->d_xc, c->first_imu_block, d_imu_res, &m->st->err);
    WC_CUDA(c, cudaMemcpyAsync(imu_res, d_imu_res, (size_t)c->n_imu_blocks * 96, cudaMemcpyDeviceToHost, st));
    WC_CUDA(c, cudaMemcpyAsync(m->h_st, m->st, sizeof(LMState), cudaMemcpyDeviceToHost, st));
    *n_imu_blocks = (size_t)c->n_imu_blocks;
  }
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  if (*n_imu_blocks && m->h_st->err == WC_EOUT_OF_SPAN) WC_FAIL(c, WC_EOUT_OF_SPAN, "IMU state outside its sample interval");
  return WC_OK;
}

extern "C" wc_status wc_surfel_markers(wc_ctx* c, const wc_surfel* surfels, size_t n, wc_marker* out) {
  if (!c || (n && (!surfels || !out))) return WC_EINVAL;
  if (n == 0) return WC_OK;
  if (n > (size_t)c->prm.max_surfels) WC_FAIL(c, WC_ECAPACITY, "n=%zu exceeds max_surfels", n);
  wc_status s = solve_alloc(c);
  if (s) return s;
  wc_solve_mem* m  = (wc_solve_mem*)c->d_lm;
  cudaStream_t  st = c->stream;
  if (n * sizeof(wc_marker) > (size_t)REC_COLS * m->stride * 8) WC_FAIL(c, WC_ECAPACITY, "marker scratch too small (raise max_corrs)");
  // d_fix is free between windows only if the caller is not mid-solve: markers use their own staging, the pack scratch
  wc_surfel* d_in = (wc_surfel*)m->rec_markers(c, n);
  if (!d_in) WC_FAIL(c, WC_ECUDA, "marker staging allocation failed");
  WC_CUDA(c, cudaMemcpyAsync(d_in, surfels, n * sizeof(wc_surfel), cudaMemcpyHostToDevice, st));
  wc_marker* d_out = (wc_marker*)m->tmp;
  { ++c->n_launches; surfel_markers<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, (int)n, d_out); }
  WC_CUDA(c, cudaMemcpyAsync(out, d_out, n * sizeof(wc_marker), cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  return WC_OK;
}

// exchange-step timing hook (wc_comm_bench): a K-pose window with a cleared LM state (done = 0, cur = 0)
wc_status wc_solve_exchange_reset(wc_ctx* c, size_t K) {
  wc_status s = solve_alloc(c);
  if (s) return s;
  if (K < 2 || K > (size_t)c->prm.max_samples) WC_FAIL(c, WC_ECAPACITY, "K out of range");
  wc_solve_mem* m = (wc_solve_mem*)c->d_lm;
  c->K = K;
  WC_CUDA(c, cudaMemsetAsync(m->st, 0, sizeof(LMState), c->stream));
  return WC_OK;
}
