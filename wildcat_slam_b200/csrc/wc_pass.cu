// Device-resident window pass: one call runs the whole hot path of LidarOdometry::AddLidarScan steps 7-13
// (src/odometry/lidar_odometry.cc:523-561) on data that already sits in HBM — BuildSurfels, UpdateSurfelPoses, the
// sliding-window and fixed-window matchers, problem assembly and the solve — without any surfel / correspondence
// round trip through the host.  This is the "inputs resident" leg of the benchmark; the host-buffer entry points
// (wc_build_surfels, wc_match, wc_window_solve) are the drop-in boundary.
#include "wc_ctx.h"

wc_status wc_window_upload_aux(wc_ctx* c, const wc_imu_state* imu, size_t n_imu, const wc_sample_state* samples, size_t K,
                               const wc_surfel* fix, size_t n_fix);
wc_status wc_window_prepare_device(wc_ctx* c);
wc_status wc_match_device(wc_ctx* c, const wc_surfel* d_q, size_t nq, const wc_surfel* d_t, size_t nt, int self_match,
                          size_t* n_out);
wc_status wc_update_surfel_poses_device(wc_ctx* c, const wc_imu_state* d_imu, size_t n_imu, wc_surfel* d_surf, size_t n);

extern "C" wc_status wc_pass_upload(wc_ctx* c, const wc_imu_state* imu, size_t n_imu, const wc_sample_state* samples,
                                    size_t K, const wc_surfel* fix, size_t n_fix) {
  return wc_window_upload_aux(c, imu, n_imu, samples, K, fix, n_fix);
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
  // 8. UpdateSurfelPoses (world -> body on first call), on the extraction's own output buffer
  WC_CUDA(c, cudaEventRecord(e1, st));
  if ((s = wc_update_surfel_poses_device(c, c->d_imu, c->n_imu, c->d_surf, S))) return s;
  WC_CUDA(c, cudaMemcpyAsync(c->d_sld, c->d_surf, S * sizeof(wc_surfel), cudaMemcpyDeviceToDevice, st));
  c->n_sld = S;
  // 9./10. sliding-window and fixed-window matchers
  size_t n_sc = 0, n_fc = 0;
  if ((s = wc_match_device(c, c->d_sld, S, c->d_sld, S, 1, &n_sc))) return s;
  if (n_sc > (size_t)c->prm.max_corrs) WC_FAIL(c, WC_ECAPACITY, "too many correspondences");
  if (n_sc) WC_CUDA(c, cudaMemcpyAsync(c->d_sld_corr, c->d_corr_out, n_sc * sizeof(wc_corr_idx), cudaMemcpyDeviceToDevice, st));
  if ((s = wc_match_device(c, c->d_sld, S, c->d_fix, c->n_fix, 0, &n_fc))) return s;
  if (n_sc + n_fc > (size_t)c->prm.max_corrs) WC_FAIL(c, WC_ECAPACITY, "too many correspondences");
  // the fixed-window residuals read a pair as (fixed, sliding): CHECK_LT(s1.t, s2.t), lidar_odometry.cc:301
  if (n_fc && c->match_query_first)
    WC_FAIL(c, WC_EINVAL_TIME_ORDER, "a fixed-window surfel is not older than the sliding-window surfel it was matched to");
  if (n_fc) WC_CUDA(c, cudaMemcpyAsync(c->d_fix_corr, c->d_corr_out, n_fc * sizeof(wc_corr_idx), cudaMemcpyDeviceToDevice, st));
  c->n_sld_corr = n_sc, c->n_fix_corr = n_fc;
  ps.n_sld_corr = (int64_t)n_sc, ps.n_fix_corr = (int64_t)n_fc;
  WC_CUDA(c, cudaEventRecord(c->ev[0], st));
  // 11. problem assembly, 13. solve
  if ((s = wc_window_prepare_device(c))) return s;
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

extern "C" int64_t wc_launch_count(const wc_ctx* c) { return c ? c->n_launches : 0; }
