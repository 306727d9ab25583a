// Device-resident window pass: one call runs the whole hot path of LidarOdometry::AddLidarScan steps 7-13
// (src/odometry/lidar_odometry.cc:523-561) on data that already sits in HBM — BuildSurfels, UpdateSurfelPoses, the
// sliding-window and fixed-window matchers, problem assembly and the solve — without any surfel / correspondence
// round trip through the host.  This is the "inputs resident" leg of the benchmark; the host-buffer entry points
// (wc_build_surfels, wc_match, wc_window_solve) are the drop-in boundary.
#include "wc_ctx.h"

wc_status wc_window_upload_aux(wc_ctx* c, const wc_imu_state* imu, size_t n_imu, const wc_sample_state* samples, size_t K,
                               const wc_surfel* fix, size_t n_fix);
wc_status wc_window_prepare_device(wc_ctx* c, int defer);
wc_status wc_match_device(wc_ctx* c, const wc_surfel* d_q, size_t nq, const wc_surfel* d_t, size_t nt, int self_match,
                          size_t* n_out);
wc_status wc_update_surfel_poses_device(wc_ctx* c, const wc_imu_state* d_imu, size_t n_imu, wc_surfel* d_surf, size_t n, int defer);
wc_status wc_pose_update_check(wc_ctx* c);

extern "C" wc_status wc_pass_upload(wc_ctx* c, const wc_imu_state* imu, size_t n_imu, const wc_sample_state* samples,
                                    size_t K, const wc_surfel* fix, size_t n_fix) {
  wc_status s = wc_window_upload_aux(c, imu, n_imu, samples, K, fix, n_fix);
  if (!s) c->n_sld_prev = 0;  // single-sweep window
  return s;
}

extern "C" wc_status wc_pass_upload_windows(wc_ctx* c, const wc_imu_state* imu, size_t n_imu, const wc_sample_state* samples, size_t K,
                                            const wc_surfel* fix, size_t n_fix, const wc_surfel* sld_prev, size_t n_sld_prev,
                                            int keep_flags) {
  if (!c || (!(keep_flags & WC_KEEP_SLD) && n_sld_prev && !sld_prev)) return WC_EINVAL;
  const size_t keep_fix = c->n_fix, keep_sld = c->n_sld;
  const bool   kf = (keep_flags & WC_KEEP_FIX) != 0, ks = (keep_flags & WC_KEEP_SLD) != 0;
  if ((kf || ks) && !c->d_lm) WC_FAIL(c, WC_EINVAL, "no resident window to keep");
  wc_status s = wc_window_upload_aux(c, imu, n_imu, samples, K, kf ? nullptr : fix, kf ? 0 : n_fix);
  if (s) return s;
  if (kf) c->n_fix = keep_fix;
  if (ks) {
    c->n_sld_prev = keep_sld;
  } else {
    if (n_sld_prev > (size_t)c->prm.max_surfels) WC_FAIL(c, WC_ECAPACITY, "too many surfels");
    if (n_sld_prev) WC_CUDA(c, cudaMemcpyAsync(c->d_sld, sld_prev, n_sld_prev * sizeof(wc_surfel), cudaMemcpyHostToDevice, c->stream));
    WC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->n_sld_prev = n_sld_prev;
  }
  c->n_sld = c->n_sld_prev;
  return WC_OK;
}

extern "C" wc_status wc_window_pass_resident(wc_ctx* c, const wc_solve_opts* opts, wc_solve_summary* summary,
                                             double* data_cor_out, wc_pass_stats* stats) {
  if (!c || !c->d_lm || c->K < 2) return WC_EINVAL;
  cudaStream_t st = c->stream;
  wc_pass_stats ps;
  memset(&ps, 0, sizeof(ps));
  cudaEvent_t e0 = c->ev[6], e1 = c->ev[7];
  float       ms;
  WC_CUDA(c, cudaEventRecord(e0, st));
  // 7. BuildSurfels
  size_t    S = 0;
  wc_status s = wc_build_surfels_resident(c, &S, &ps.ms_extract_keys, &ps.ms_extract_emit, &ps.ms_extract);
  if (s) return s;
  ps.n_surfels = (int64_t)S;
  if (S > (size_t)c->prm.max_surfels) WC_FAIL(c, WC_ECAPACITY, "too many surfels");
  // surfels_sld_win_.insert(end, sweep) (lidar_odometry.cc:527), then 8. UpdateSurfelPoses over the WHOLE window (:528):
  // new surfels go world -> body, the surfels of earlier sweeps get their poses re-interpolated from the current IMU states
  WC_CUDA(c, cudaEventRecord(e1, st));
  const size_t base = c->n_sld_prev;
  if (base + S > (size_t)c->prm.max_surfels) WC_FAIL(c, WC_ECAPACITY, "sliding window exceeds max_surfels");
  WC_CUDA(c, cudaMemcpyAsync(c->d_sld + base, c->d_surf, S * sizeof(wc_surfel), cudaMemcpyDeviceToDevice, st));
  const size_t W = base + S;
  if ((s = wc_update_surfel_poses_device(c, c->d_imu, c->n_imu, c->d_sld, W, /*defer=*/1))) return s;  // checked after the matcher's sync
  c->n_sld = W;
  // 9./10. sliding-window and fixed-window matchers
  size_t n_sc = 0, n_fc = 0;
  if ((s = wc_match_device(c, c->d_sld, W, c->d_sld, W, 1, &n_sc))) return s;
  if ((s = wc_pose_update_check(c))) return s;
  if (n_sc > (size_t)c->prm.max_corrs) WC_FAIL(c, WC_ECAPACITY, "too many correspondences");
  if (n_sc) WC_CUDA(c, cudaMemcpyAsync(c->d_sld_corr, c->d_corr_out, n_sc * sizeof(wc_corr_idx), cudaMemcpyDeviceToDevice, st));
  if ((s = wc_match_device(c, c->d_sld, W, c->d_fix, c->n_fix, 0, &n_fc))) return s;
  if (n_sc + n_fc > (size_t)c->prm.max_corrs) WC_FAIL(c, WC_ECAPACITY, "too many correspondences");
  // the fixed-window residuals read a pair as (fixed, sliding): CHECK_LT(s1.t, s2.t), lidar_odometry.cc:301
  if (n_fc && c->match_query_first)
    WC_FAIL(c, WC_EINVAL_TIME_ORDER, "a fixed-window surfel is not older than the sliding-window surfel it was matched to");
  if (n_fc) WC_CUDA(c, cudaMemcpyAsync(c->d_fix_corr, c->d_corr_out, n_fc * sizeof(wc_corr_idx), cudaMemcpyDeviceToDevice, st));
  c->n_sld_corr = n_sc, c->n_fix_corr = n_fc;
  ps.n_sld_corr = (int64_t)n_sc, ps.n_fix_corr = (int64_t)n_fc;
  WC_CUDA(c, cudaEventRecord(c->ev[0], st));
  // 11. problem assembly, 13. solve
  if ((s = wc_window_prepare_device(c, /*defer=*/1))) return s;  // pack errors surface at the solve's first host check
  WC_CUDA(c, cudaEventRecord(c->ev[1], st));
  wc_solve_summary local;
  s = wc_window_solve_resident(c, opts, summary ? summary : &local, data_cor_out);
  if (s) return s;
  WC_CUDA(c, cudaEventRecord(c->ev[2], st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ms, e1, c->ev[0]);
  ps.ms_match = ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  ps.ms_pack = ms;
  cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]);
  ps.ms_solve = ms;
  cudaEventElapsedTime(&ms, e0, c->ev[2]);
  ps.ms_total = ms;
  ps.n_launches = c->n_launches;
  if (stats) *stats = ps;
  return WC_OK;
}

// ---- ShrinkToFit, surfel half (lidar_odometry.cc:242-249) ------------------------------------------------------------
namespace {
// first index whose timestamp is not below t_cut (the window is time ordered: sweeps arrive in order, each one sorted)
__global__ void window_cut(const wc_surfel* __restrict__ s, int n, double t_cut, int* __restrict__ out) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s[mid].timestamp < t_cut) lo = mid + 1; else hi = mid;
  }
  out[0] = lo;
}
// new fixed window = reversed(moved prefix) ++ old fixed window [0 .. keep): 13 x 16-byte chunks per surfel
__global__ void window_move(const wc_surfel* __restrict__ sld, int m, const wc_surfel* __restrict__ fix_old, int keep,
                            wc_surfel* __restrict__ fix_new) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t / 13, c = t % 13;
  if (i >= m + keep) return;
  const wc_surfel* src = i < m ? sld + (m - 1 - i) : fix_old + (i - m);
  reinterpret_cast<uint4*>(fix_new + i)[c] = reinterpret_cast<const uint4*>(src)[c];
}
__global__ void window_shift(const wc_surfel* __restrict__ in, int m, int n, wc_surfel* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t / 13, c = t % 13;
  if (i >= n - m) return;
  reinterpret_cast<uint4*>(out + i)[c] = reinterpret_cast<const uint4*>(in + m + i)[c];
}
// trim_fixed: number of leading entries (newest first) within `dur` of the newest one
__global__ void window_trim(const wc_surfel* __restrict__ fix, int n, double dur, int* __restrict__ out) {
  int keep = n;
  if (n > 0) {
    const double t0 = fix[0].timestamp;
    while (keep > 0 && t0 - fix[keep - 1].timestamp > dur) --keep;  // newest first: the back holds the oldest
  }
  out[1] = keep;
}
}  // namespace

extern "C" wc_status wc_window_shrink(wc_ctx* c, double t_front_imu, double fix_window_duration, int trim_fixed, size_t* n_sld,
                                      size_t* n_fix) {
  if (!c || !c->d_lm) return WC_EINVAL;
  cudaStream_t st = c->stream;
  if (!c->d_fix_tmp) WC_CUDA(c, cudaMalloc(&c->d_fix_tmp, (size_t)c->prm.max_surfels * sizeof(wc_surfel)));
  int* d_cnt = c->d_status ? c->d_status : nullptr;
  if (!d_cnt) {
    WC_CUDA(c, cudaMalloc(&c->d_status, 16));
    d_cnt = c->d_status;
  }
  int h[2] = {0, 0};
  { ++c->n_launches; window_cut<<<1, 1, 0, st>>>(c->d_sld, (int)c->n_sld, t_front_imu, d_cnt); }
  WC_CUDA(c, cudaMemcpyAsync(h, d_cnt, 4, cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  const size_t m = (size_t)h[0];
  if (m) {
    if (c->n_fix + m > (size_t)c->prm.max_surfels) WC_FAIL(c, WC_ECAPACITY, "fixed window exceeds max_surfels");
    const size_t tot = m + c->n_fix;
    { ++c->n_launches; window_move<<<(unsigned)((tot * 13 + 255) / 256), 256, 0, st>>>(c->d_sld, (int)m, c->d_fix, (int)c->n_fix, c->d_fix_tmp); }
    wc_surfel* t = c->d_fix;
    c->d_fix = c->d_fix_tmp, c->d_fix_tmp = t;
    c->n_fix = tot;
    // the remaining sliding-window surfels move to the front (through the staging buffer: the ranges overlap)
    const size_t rest = c->n_sld - m;
    if (rest) {
      { ++c->n_launches; window_shift<<<(unsigned)((rest * 13 + 255) / 256), 256, 0, st>>>(c->d_sld, (int)m, (int)c->n_sld, c->d_fix_tmp); }
      WC_CUDA(c, cudaMemcpyAsync(c->d_sld, c->d_fix_tmp, rest * sizeof(wc_surfel), cudaMemcpyDeviceToDevice, st));
    }
    c->n_sld = rest;
  }
  if (trim_fixed && c->n_fix) {
    { ++c->n_launches; window_trim<<<1, 1, 0, st>>>(c->d_fix, (int)c->n_fix, fix_window_duration, d_cnt); }
    WC_CUDA(c, cudaMemcpyAsync(h, d_cnt, 8, cudaMemcpyDeviceToHost, st));
    WC_CUDA(c, cudaStreamSynchronize(st));
    c->n_fix = (size_t)h[1];
  }
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  c->n_sld_prev = c->n_sld;
  if (n_sld) *n_sld = c->n_sld;
  if (n_fix) *n_fix = c->n_fix;
  return WC_OK;
}

extern "C" wc_status wc_windows_fetch(wc_ctx* c, wc_surfel* sld, size_t sld_cap, size_t* n_sld, wc_surfel* fix, size_t fix_cap,
                                      size_t* n_fix) {
  if (!c || !c->d_lm) return WC_EINVAL;
  if (n_sld) *n_sld = c->n_sld;
  if (n_fix) *n_fix = c->n_fix;
  if ((sld && c->n_sld > sld_cap) || (fix && c->n_fix > fix_cap)) WC_FAIL(c, WC_ECAPACITY, "output capacity too small");
  if (sld && c->n_sld) WC_CUDA(c, cudaMemcpyAsync(sld, c->d_sld, c->n_sld * sizeof(wc_surfel), cudaMemcpyDeviceToHost, c->stream));
  if (fix && c->n_fix) WC_CUDA(c, cudaMemcpyAsync(fix, c->d_fix, c->n_fix * sizeof(wc_surfel), cudaMemcpyDeviceToHost, c->stream));
  WC_CUDA(c, cudaStreamSynchronize(c->stream));
  return WC_OK;
}

extern "C" int64_t wc_launch_count(const wc_ctx* c) { return c ? c->n_launches : 0; }
