// Cubic B-spline correction spreading and pose updates on the device — replaces CubicBSplineInterpolator
// (src/odometry/spline_interpolation.h:42-113), CubicBSplineSampleCorrector and UpdateSurfelPoses / UpdateSamplePoses /
// UpdateImuPoses / PredictPoseOfNewImuState (src/odometry/lidar_odometry.cc:22-54,112-123,160-215), Surfel::UpdatePose
// (src/odometry/surfel.h:48-58).
//
//   bspline_fit     control points Q = (N^T N)^-1 N^T P for up to 6 channels at once; N has rows
//                   [1,4,1,0]/6 at clamp(i-1..i+2) (spline_interpolation.h:86-99).  One CTA, Gaussian elimination
//                   with partial pivoting on the augmented normal system in shared memory.
//   bspline_eval    u = (t-t0)/(tK-1 - t0)*(K-1)+1, i = floor(u), tau = u-i, [tau^3 tau^2 tau 1] M Q4 / 6 (:51-72)
//   update_surfel_poses / apply_imu_corrections / repredict_last / update_samples
#include "wc_ctx.h"
#include "wc_device_math.cuh"

using namespace wcd;

namespace {

__global__ void __launch_bounds__(256) bspline_fit(const double* __restrict__ pts, int K, int nch, double* __restrict__ Q) {
  extern __shared__ __align__(16) double sm[];
  const int W = K + nch;  // augmented width
  double*   A = sm;       // K x W : [N^T N | N^T P]
  __shared__ int    piv;
  __shared__ double red_v[256];
  __shared__ int    red_i[256];
  const int t = threadIdx.x;
  for (int e = t; e < K * W; e += 256) A[e] = 0.0;
  __syncthreads();
  // N(i, clamp(i-1+j)) += [1,4,1,0][j]/6 ; accumulate N^T N and N^T P by one thread per output row block
  // (K <= 128: the triple loop is tiny).  Thread r owns output row r.
  for (int r = t; r < K; r += 256) {
    for (int i = 0; i < K; ++i) {
      double nrow[3];
      int    idx[3];
      const double w[3] = {1.0 / 6.0, 4.0 / 6.0, 1.0 / 6.0};
      for (int j = 0; j < 3; ++j) idx[j] = min(max(i - 1 + j, 0), K - 1), nrow[j] = w[j];
      double nir = 0.0;  // N(i, r)
      for (int j = 0; j < 3; ++j)
        if (idx[j] == r) nir += nrow[j];
      if (nir == 0.0) continue;
      for (int j = 0; j < 3; ++j) A[r * W + idx[j]] += nir * nrow[j];
      for (int c = 0; c < nch; ++c) A[r * W + K + c] += nir * pts[i * nch + c];
    }
  }
  __syncthreads();
  for (int col = 0; col < K; ++col) {
    // partial pivot
    double best = -1.0;
    int    bi   = col;
    for (int r = col + t; r < K; r += 256) {
      const double v = fabs(A[r * W + col]);
      if (v > best) best = v, bi = r;
    }
    red_v[t] = best, red_i[t] = bi;
    __syncthreads();
    if (t == 0) {
      double b = -1.0;
      int    p = col;
      for (int k = 0; k < 256; ++k)
        if (red_v[k] > b) b = red_v[k], p = red_i[k];
      piv = p;
    }
    __syncthreads();
    const int p = piv;
    if (p != col)
      for (int c = t; c < W; c += 256) {
        const double tmp = A[col * W + c];
        A[col * W + c]   = A[p * W + c];
        A[p * W + c]     = tmp;
      }
    __syncthreads();
    const double d = A[col * W + col];
    __syncthreads();
    for (int c = t; c < W; c += 256) A[col * W + c] /= d;
    __syncthreads();
    for (int e = t; e < K * W; e += 256) {
      const int r = e / W, c = e % W;
      if (r == col || c == col) continue;
      A[e] -= A[r * W + col] * A[col * W + c];
    }
    __syncthreads();
    for (int r = t; r < K; r += 256)
      if (r != col) A[r * W + col] = 0.0;
    __syncthreads();
  }
  for (int e = t; e < K * nch; e += 256) Q[e] = A[(e / nch) * W + K + e % nch];
}

__device__ __forceinline__ bool bspline_interp(const double* __restrict__ Q, int K, int nch, double t0, double t1, double t,
                                               double* out) {
  if (t < t0 || t > t1) return false;  // Interp returns nullptr (spline_interpolation.h:52-54)
  const double index_f   = (t - t0) / (t1 - t0) * (double)(K - 1) + 1.0;
  const int    index_int = (int)floor(index_f);
  const double tau       = index_f - index_int;
  const double tv[4]     = {tau * tau * tau, tau * tau, tau, 1.0};
  const double M[4][4]   = {{-1, 3, -3, 1}, {3, -6, 3, 0}, {-3, 0, 3, 0}, {1, 4, 1, 0}};
  double       tm[4];
  for (int j = 0; j < 4; ++j) tm[j] = tv[0] * M[0][j] + tv[1] * M[1][j] + tv[2] * M[2][j] + tv[3] * M[3][j];
  for (int c = 0; c < nch; ++c) {
    double s = 0.0;
    for (int j = 0; j < 4; ++j) s += tm[j] * Q[min(max(index_int - 2 + j, 0), K - 1) * nch + c];
    out[c] = s / 6.0;
  }
  return true;
}

__global__ void bspline_eval(const double* __restrict__ Q, int K, int nch, double t0, double t1, const double* __restrict__ tq,
                             int nq, double* __restrict__ out, unsigned char* __restrict__ valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  double o[6] = {0, 0, 0, 0, 0, 0};
  const bool ok = bspline_interp(Q, K, nch, t0, t1, tq[i], o);
  for (int c = 0; c < nch; ++c) out[(size_t)i * nch + c] = o[c];
  if (valid) valid[i] = ok ? 1 : 0;
}

__device__ __forceinline__ int imu_lower_bound(const wc_imu_state* __restrict__ imu, int n, double t) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (imu[mid].timestamp < t) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// UpdateSurfelPoses (lidar_odometry.cc:160-170) + Surfel::UpdatePose (surfel.h:48-58)
__global__ void update_surfel_poses(const wc_imu_state* __restrict__ imu, int n_imu, wc_surfel* __restrict__ s, int n,
                                    int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t   = s[i].timestamp;
  const int    idx = imu_lower_bound(imu, n_imu, t);
  if (idx == 0 || idx == n_imu) {  // CHECK :164
    *err = WC_EOUT_OF_SPAN;
    return;
  }
  const wc_imu_state &a = imu[idx - 1], &b = imu[idx];
  const double        f   = (t - a.timestamp) / (b.timestamp - a.timestamp);
  const V3            pos = ld3(a.pos) * (1 - f) + ld3(b.pos) * f;
  const Q4            rot = Slerp(ldq(a.rot), f, ldq(b.rot));
  st3(s[i].pos, pos);
  stq(s[i].rot, rot);
  if (!s[i].is_in_body_frame) {
    s[i].is_in_body_frame = 1;
    const Q4 rc = conj(rot);
    st3(s[i].center, rc * (ld3(s[i].center) - pos));
    st3(s[i].norm, rc * ld3(s[i].norm));
    st33(s[i].covariance, (ToMatrix(rc) * ld33(s[i].covariance)) * ToMatrix(rot));
  }
}

// UpdateImuPoses, first loop (lidar_odometry.cc:192-203): Q holds 6 channels [rot_cor, pos_cor]
__global__ void apply_imu_corrections(const double* __restrict__ Q, int K, double t0, double t1, wc_imu_state* __restrict__ imu,
                                      int n_imu, int* __restrict__ first_last) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_imu) return;
  double c[6];
  if (!bspline_interp(Q, K, 6, t0, t1, imu[i].timestamp, c)) return;
  stq(imu[i].rot, Exp(mk(c[0], c[1], c[2])) * ldq(imu[i].rot));
  st3(imu[i].pos, mk(c[3], c[4], c[5]) + ld3(imu[i].pos));
  atomicMin(&first_last[0], i);
  atomicMax(&first_last[1], i);
}

// PredictPoseOfNewImuState for the trailing IMU state (:205-214, 112-123) and UpdateSamplePoses (:172-179)
__global__ void repredict_and_update_samples(wc_imu_state* __restrict__ imu, int n_imu, wc_sample_state* __restrict__ s, int K,
                                             const int* __restrict__ first_last, int* __restrict__ err) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0 && first_last[1] >= 0) {
    if (first_last[0] != 0 || first_last[1] != n_imu - 2) {
      *err = WC_EOUT_OF_SPAN;  // CHECK_EQ :209-210
    } else {
      const wc_imu_state &i1 = imu[n_imu - 3], &i2 = imu[n_imu - 2];
      wc_imu_state&       i3 = imu[n_imu - 1];
      const wc_sample_state& b = s[K - 1];
      const V3 ba = ld3(b.data_cor + 9), bg = ld3(b.data_cor + 6), grav = ld3(b.grav);
      const double dt = i3.timestamp - i2.timestamp;
      if (!(fabs(dt - (i2.timestamp - i1.timestamp)) <= 1e-6)) *err = WC_EINVAL_TIME_ORDER;  // CHECK_NEAR :119
      stq(i3.rot, ldq(i2.rot) * Exp(((ld3(i2.gyr) + ld3(i3.gyr)) / 2.0 - bg) * dt));
      st3(i3.pos, (ldq(i1.rot) * (ld3(i1.acc) - ba) + grav) * dt * dt + 2.0 * ld3(i2.pos) - ld3(i1.pos));
    }
  }
  if (k < K) {
    // note: sample K-1's bias entries are read by thread 0 above; only rot/pos corrections are folded and zeroed
    const Q4 q = Exp(ld3(s[k].data_cor)) * ldq(s[k].rot);
    const V3 p = ld3(s[k].data_cor + 3) + ld3(s[k].pos);
    stq(s[k].rot, q);
    st3(s[k].pos, p);
    for (int c = 0; c < 6; ++c) s[k].data_cor[c] = 0.0;
  }
}

__global__ void gather_corrections(const wc_sample_state* __restrict__ s, int K, double* __restrict__ pts6, double* __restrict__ ts) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  for (int c = 0; c < 6; ++c) pts6[6 * k + c] = s[k].data_cor[c];
  ts[k] = s[k].timestamp;
}

// PredictImuStatesAndSampleStates steps 2-3 (lidar_odometry.cc:403-453) in one CTA:
//   A (parallel)   dq_k = Exp(((g_{k-1} + g_k) / 2 - bg) dt_k), parked in imu[k].rot; spacing check (CHECK_NEAR :119)
//   B (one thread) rot_k = rot_{k-1} * dq_k                      — the chain itself is sequential in the reference too
//   C (parallel)   a_k = (R_{k-2} (acc_{k-2} - ba) + grav) dt dt, parked in imu[k].pos
//   D (one thread) pos_k = a_k + 2 pos_{k-1} - pos_{k-2}
//   E (parallel)   new sample states: lerp / slerp of the bracketing IMU states (:439-449)
// The expensive parts (Exp: sincos; the rotations; slerp) run in parallel, the two recurrences cost a few FMAs per step.
__global__ void __launch_bounds__(256)
predict_states(wc_imu_state* __restrict__ imu, int n, V3 ba, V3 bg, V3 grav, double t_last, double sample_dt, int n_new,
               wc_sample_state* __restrict__ out, int* __restrict__ err) {
  const int t = threadIdx.x;
  for (int k = 2 + t; k < n; k += 256) {
    const double d3 = imu[k].timestamp - imu[k - 1].timestamp, d2 = imu[k - 1].timestamp - imu[k - 2].timestamp;
    if (!(fabs(d3 - d2) <= 1e-6)) err[0] = WC_EINVAL_TIME_ORDER;
    stq(imu[k].rot, Exp(((ld3(imu[k - 1].gyr) + ld3(imu[k].gyr)) / 2.0 - bg) * d3));
  }
  __syncthreads();
  if (t == 0)
    for (int k = 2; k < n; ++k) stq(imu[k].rot, ldq(imu[k - 1].rot) * ldq(imu[k].rot));
  __syncthreads();
  for (int k = 2 + t; k < n; k += 256) {
    const double dt = imu[k].timestamp - imu[k - 1].timestamp;
    st3(imu[k].pos, (ldq(imu[k - 2].rot) * (ld3(imu[k - 2].acc) - ba) + grav) * dt * dt);
  }
  __syncthreads();
  if (t == 0)
    for (int k = 2; k < n; ++k) st3(imu[k].pos, ld3(imu[k].pos) + 2.0 * ld3(imu[k - 1].pos) - ld3(imu[k - 2].pos));
  __syncthreads();
  for (int i = 1 + t; i <= n_new; i += 256) {
    const double     ts = t_last + (double)i * sample_dt;
    wc_sample_state& ss = out[i - 1];
    ss.timestamp = ts;
#pragma unroll
    for (int c = 0; c < 6; ++c) ss.data_cor[c] = 0.0;
    ss.data_cor[6] = bg.x, ss.data_cor[7] = bg.y, ss.data_cor[8] = bg.z;
    ss.data_cor[9] = ba.x, ss.data_cor[10] = ba.y, ss.data_cor[11] = ba.z;
    st3(ss.grav, grav);
    const int idx = imu_lower_bound(imu, n, ts);
    if (idx == 0 || idx == n) {  // CHECK_NE :441-442
      err[0] = WC_EOUT_OF_SPAN;
      continue;
    }
    const wc_imu_state &a = imu[idx - 1], &b = imu[idx];
    const double        f = (ts - a.timestamp) / (b.timestamp - a.timestamp);
    stq(ss.rot, Slerp(ldq(a.rot), f, ldq(b.rot)));
    st3(ss.pos, (1 - f) * ld3(a.pos) + f * ld3(b.pos));
  }
}

}  // namespace

struct wc_spline_mem {
  double* pts;  // K x 6
  double* Q;    // K x 6
  double* tq;
  double* out;
  unsigned char* valid;
  double* ts;
  int*    flags;   // [0] first, [1] last, [2] err
  int*    h_flags;
  size_t  qcap;
  wc_surfel*       surf;
  wc_imu_state*    imu;
  wc_sample_state* samples;
};

static wc_status spline_alloc(wc_ctx* c, size_t nq) {
  wc_spline_mem* m = (wc_spline_mem*)c->d_spline;
  if (!m) {
    m           = (wc_spline_mem*)calloc(1, sizeof(wc_spline_mem));
    c->d_spline = m;
    const size_t K = (size_t)c->prm.max_samples;
    WC_CUDA(c, cudaMalloc(&m->pts, K * 6 * 8));
    WC_CUDA(c, cudaMalloc(&m->Q, K * 6 * 8));
    WC_CUDA(c, cudaMalloc(&m->ts, K * 8));
    WC_CUDA(c, cudaMalloc(&m->flags, 16));
    WC_CUDA(c, cudaMallocHost(&m->h_flags, 16));
    WC_CUDA(c, cudaMalloc(&m->surf, (size_t)c->prm.max_surfels * sizeof(wc_surfel)));
    WC_CUDA(c, cudaMalloc(&m->imu, (size_t)c->prm.max_imu_states * sizeof(wc_imu_state)));
    WC_CUDA(c, cudaMalloc(&m->samples, K * sizeof(wc_sample_state)));
    WC_CUDA(c, cudaFuncSetAttribute(bspline_fit, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  if (nq > m->qcap) {
    if (m->tq) cudaFree(m->tq), cudaFree(m->out), cudaFree(m->valid);
    m->qcap = nq < 4096 ? 4096 : nq;
    WC_CUDA(c, cudaMalloc(&m->tq, m->qcap * 8));
    WC_CUDA(c, cudaMalloc(&m->out, m->qcap * 6 * 8));
    WC_CUDA(c, cudaMalloc(&m->valid, m->qcap));
  }
  return WC_OK;
}

void wc_spline_free(wc_ctx* c) {
  wc_spline_mem* m = (wc_spline_mem*)c->d_spline;
  if (!m) return;
  void* p[] = {m->pts, m->Q, m->tq, m->out, m->valid, m->ts, m->flags, m->surf, m->imu, m->samples};
  for (void* q : p)
    if (q) cudaFree(q);
  if (m->h_flags) cudaFreeHost(m->h_flags);
  free(m);
  c->d_spline = nullptr;
}

extern "C" wc_status wc_spline_fit_eval(wc_ctx* c, const double* ts, const double* pts3, size_t K, const double* query_t,
                                        size_t nq, double* out3, uint8_t* valid) {
  if (!c || !ts || !pts3 || K < 2 || (nq && (!query_t || !out3))) return WC_EINVAL;
  if (K > (size_t)c->prm.max_samples) WC_FAIL(c, WC_ECAPACITY, "K exceeds max_samples");
  wc_status s = spline_alloc(c, nq);
  if (s) return s;
  wc_spline_mem* m  = (wc_spline_mem*)c->d_spline;
  cudaStream_t   st = c->stream;
  WC_CUDA(c, cudaMemcpyAsync(m->pts, pts3, K * 3 * 8, cudaMemcpyHostToDevice, st));
  if (nq) WC_CUDA(c, cudaMemcpyAsync(m->tq, query_t, nq * 8, cudaMemcpyHostToDevice, st));
  { ++c->n_launches; bspline_fit<<<1, 256, K * (K + 3) * 8, st>>>(m->pts, (int)K, 3, m->Q); }
  if (nq) {
    { ++c->n_launches; bspline_eval<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(m->Q, (int)K, 3, ts[0], ts[K - 1], m->tq, (int)nq, m->out, m->valid); }
    WC_CUDA(c, cudaMemcpyAsync(out3, m->out, nq * 3 * 8, cudaMemcpyDeviceToHost, st));
    if (valid) WC_CUDA(c, cudaMemcpyAsync(valid, m->valid, nq, cudaMemcpyDeviceToHost, st));
  }
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  return WC_OK;
}

extern "C" wc_status wc_update_surfel_poses(wc_ctx* c, const wc_imu_state* imu, size_t n_imu, wc_surfel* surfels, size_t n) {
  if (!c || (n && (!imu || !surfels))) return WC_EINVAL;
  if (n == 0) return WC_OK;
  if (n > (size_t)c->prm.max_surfels || n_imu > (size_t)c->prm.max_imu_states) WC_FAIL(c, WC_ECAPACITY, "capacity exceeded");
  wc_status s = spline_alloc(c, 0);
  if (s) return s;
  wc_spline_mem* m  = (wc_spline_mem*)c->d_spline;
  cudaStream_t   st = c->stream;
  WC_CUDA(c, cudaMemcpyAsync(m->imu, imu, n_imu * sizeof(wc_imu_state), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemcpyAsync(m->surf, surfels, n * sizeof(wc_surfel), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemsetAsync(m->flags, 0, 16, st));
  { ++c->n_launches; update_surfel_poses<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->imu, (int)n_imu, m->surf, (int)n, m->flags + 2); }
  WC_CUDA(c, cudaMemcpyAsync(m->h_flags, m->flags, 16, cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaMemcpyAsync(surfels, m->surf, n * sizeof(wc_surfel), cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  if (m->h_flags[2]) WC_FAIL(c, WC_EOUT_OF_SPAN, "surfel timestamp outside the IMU state span (lidar_odometry.cc:164)");
  return WC_OK;
}

// device-resident variant for the fused window pass (surfels already on the device)
// defer != 0: no host synchronisation here — the caller checks wc_pose_update_check() after its next one
wc_status wc_update_surfel_poses_device(wc_ctx* c, const wc_imu_state* d_imu, size_t n_imu, wc_surfel* d_surf, size_t n, int defer) {
  wc_status s = spline_alloc(c, 0);
  if (s) return s;
  wc_spline_mem* m = (wc_spline_mem*)c->d_spline;
  m->h_flags[2]    = 0;
  if (n == 0) return WC_OK;
  WC_CUDA(c, cudaMemsetAsync(m->flags, 0, 16, c->stream));
  { ++c->n_launches; update_surfel_poses<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_imu, (int)n_imu, d_surf, (int)n, m->flags + 2); }
  WC_CUDA(c, cudaMemcpyAsync(m->h_flags, m->flags, 16, cudaMemcpyDeviceToHost, c->stream));
  if (defer) return WC_OK;
  WC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (m->h_flags[2]) WC_FAIL(c, WC_EOUT_OF_SPAN, "surfel timestamp outside the IMU state span (lidar_odometry.cc:164)");
  return WC_OK;
}
wc_status wc_pose_update_check(wc_ctx* c) {  // after a host synchronisation that followed a deferred pose update
  wc_spline_mem* m = (wc_spline_mem*)c->d_spline;
  if (m && m->h_flags[2]) WC_FAIL(c, WC_EOUT_OF_SPAN, "surfel timestamp outside the IMU state span (lidar_odometry.cc:164)");
  return WC_OK;
}

extern "C" wc_status wc_predict_states(wc_ctx* c, wc_imu_state* imu, size_t n_imu, const double* ba3, const double* bg3,
                                       const double* grav3, double t_last_sample, double sample_dt, size_t n_new,
                                       wc_sample_state* samples_out) {
  if (!c || !imu || n_imu < 2 || !ba3 || !bg3 || !grav3 || (n_new && !samples_out)) return WC_EINVAL;
  if (n_new > (size_t)c->prm.max_samples || n_imu > (size_t)c->prm.max_imu_states) WC_FAIL(c, WC_ECAPACITY, "capacity exceeded");
  wc_status s = spline_alloc(c, 0);
  if (s) return s;
  wc_spline_mem* m  = (wc_spline_mem*)c->d_spline;
  cudaStream_t   st = c->stream;
  WC_CUDA(c, cudaMemcpyAsync(m->imu, imu, n_imu * sizeof(wc_imu_state), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemsetAsync(m->flags, 0, 16, st));
  { ++c->n_launches; predict_states<<<1, 256, 0, st>>>(m->imu, (int)n_imu, mk(ba3[0], ba3[1], ba3[2]), mk(bg3[0], bg3[1], bg3[2]),
                                                        mk(grav3[0], grav3[1], grav3[2]), t_last_sample, sample_dt, (int)n_new, m->samples,
                                                        m->flags); }
  WC_CUDA(c, cudaMemcpyAsync(m->h_flags, m->flags, 16, cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaMemcpyAsync(imu, m->imu, n_imu * sizeof(wc_imu_state), cudaMemcpyDeviceToHost, st));
  if (n_new) WC_CUDA(c, cudaMemcpyAsync(samples_out, m->samples, n_new * sizeof(wc_sample_state), cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  if (m->h_flags[0] == WC_EINVAL_TIME_ORDER) WC_FAIL(c, WC_EINVAL_TIME_ORDER, "IMU samples are not uniformly spaced within 1e-6 s (lidar_odometry.cc:119)");
  if (m->h_flags[0] == WC_EOUT_OF_SPAN) WC_FAIL(c, WC_EOUT_OF_SPAN, "new sample time outside the predicted IMU states (lidar_odometry.cc:441-442)");
  return WC_OK;
}

extern "C" wc_status wc_apply_corrections(wc_ctx* c, wc_sample_state* samples, size_t K, wc_imu_state* imu, size_t n_imu) {
  if (!c || !samples || K < 2 || (n_imu && !imu)) return WC_EINVAL;
  if (K > (size_t)c->prm.max_samples || n_imu > (size_t)c->prm.max_imu_states) WC_FAIL(c, WC_ECAPACITY, "capacity exceeded");
  if (n_imu > 0 && n_imu < 3) WC_FAIL(c, WC_EINVAL, "need at least 3 IMU states");
  wc_status s = spline_alloc(c, 0);
  if (s) return s;
  wc_spline_mem* m  = (wc_spline_mem*)c->d_spline;
  cudaStream_t   st = c->stream;
  WC_CUDA(c, cudaMemcpyAsync(m->samples, samples, K * sizeof(wc_sample_state), cudaMemcpyHostToDevice, st));
  if (n_imu) WC_CUDA(c, cudaMemcpyAsync(m->imu, imu, n_imu * sizeof(wc_imu_state), cudaMemcpyHostToDevice, st));
  const int init[4] = {0x7fffffff, -1, 0, 0};
  WC_CUDA(c, cudaMemcpyAsync(m->flags, init, 16, cudaMemcpyHostToDevice, st));
  { ++c->n_launches; gather_corrections<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(m->samples, (int)K, m->pts, m->ts); }
  { ++c->n_launches; bspline_fit<<<1, 256, K * (K + 6) * 8, st>>>(m->pts, (int)K, 6, m->Q); }
  if (n_imu)
    { ++c->n_launches; apply_imu_corrections<<<(unsigned)((n_imu + 127) / 128), 128, 0, st>>>(m->Q, (int)K, samples[0].timestamp,
                                                                           samples[K - 1].timestamp, m->imu, (int)n_imu, m->flags); }
  { ++c->n_launches; repredict_and_update_samples<<<(unsigned)((K + 127) / 128), 128, 0, st>>>(m->imu, (int)n_imu, m->samples, (int)K, m->flags,
                                                                            m->flags + 2); }
  // the flags come back first: on an error the caller's arrays stay untouched (the reference would have aborted)
  WC_CUDA(c, cudaMemcpyAsync(m->h_flags, m->flags, 16, cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  if (m->h_flags[2] == WC_EINVAL_TIME_ORDER)
    WC_FAIL(c, WC_EINVAL_TIME_ORDER, "trailing IMU samples are not uniformly spaced within 1e-6 s (lidar_odometry.cc:119)");
  if (m->h_flags[2]) WC_FAIL(c, WC_EOUT_OF_SPAN, "IMU states are not exactly covered by the sample span (lidar_odometry.cc:209-210)");
  WC_CUDA(c, cudaMemcpyAsync(samples, m->samples, K * sizeof(wc_sample_state), cudaMemcpyDeviceToHost, st));
  if (n_imu) WC_CUDA(c, cudaMemcpyAsync(imu, m->imu, n_imu * sizeof(wc_imu_state), cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  return WC_OK;
}
