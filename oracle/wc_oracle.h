/* ORACLE — TEST INFRASTRUCTURE ONLY (see wc_math.h).  C API of the CPU restatement of the reference's
 * sliding-window odometry hot path, consumed through ctypes by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  POD types come from include/wildcat_b200.h.
 *
 * Parity pinning: the reference cannot be compiled here (Eigen, Ceres, FLANN, PCL, glog,
 * gtest absent), so this restatement is pinned only by
 *   - the reference's own known-answer unit tests re-expressed on it (tests/test_oracle_golden.py:
 *     spline_interpolation_test.cc:10-41,79-96, utils_test.cc:5-21, knn_surfel_matcher_test.cc:19-43),
 *   - the output of scripts/CubicBSpline3D.ipynb (runs here; vectors in tests/golden/).
 * BuildSurfels, Match gating, the cost functors and the Ceres LM loop are touched by NO reference test
 * or fixture: for those rows PARITY IS UNPINNED (restated from the cited lines and, for Ceres 1.14
 * trust_region_minimizer / levenberg_marquardt_strategy / loss_function / corrector, from their
 * published behaviour).
 */
#ifndef WC_ORACLE_H_
#define WC_ORACLE_H_
#include "../include/wildcat_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wco_surfel_info {
  int32_t n_points;
  int32_t layer;        /* 0,1,2 */
  double  evals[3];     /* ascending */
  double  likeness;
} wco_surfel_info;

/* BuildSurfels, surfel_extraction.cc:316-337.  Returns number of surfels (or -1 if cap too small).
 * assign/info may be NULL.  near_threshold: number of planarity decisions (node or cluster) whose
 * eigen test sits within rel_margin of a threshold. */
int64_t wco_build_surfels(const wc_params* prm, const wc_point48* pts, int64_t n, wc_surfel* out, int64_t cap,
                          wc_point_assign* assign, wco_surfel_info* info, double rel_margin,
                          int64_t* near_threshold);

/* UpdateSurfelPoses, lidar_odometry.cc:160-170 + Surfel::UpdatePose surfel.h:48-58.  0 ok, else wc_status. */
int wco_update_surfel_poses(const wc_imu_state* imu, int64_t n_imu, wc_surfel* surfels, int64_t n);

/* KnnSurfelMatcher::BuildIndex + Match, knn_surfel_matcher.cc:3-49.  use_kdtree: 1 = kd-tree exact search
 * (what FLANN does; the timed baseline), 0 = brute force (the definition).  Returns number of pairs. */
int64_t wco_match(const wc_params* prm, const wc_surfel* query, int64_t nq, const wc_surfel* target, int64_t nt,
                  int self_match, int use_kdtree, wc_corr_idx* out, uint8_t* first_is_target);
/* knn_surfel_matcher.cc:64-89 */
void wco_knn6(const double* query6, int64_t nq, const double* target6, int64_t nt, int k, int use_kdtree,
              int32_t* out_idx, double* out_dist2);

/* problem.Evaluate-like test hook: cost, gradient (12K) and J^T J (12K x 12K row-major) of the robustified
 * problem at samples[].data_cor.  Any output may be NULL.  Returns wc_status. */
int wco_window_evaluate(const wc_params* prm, const wc_solve_opts* opts, const wc_surfel* sld, int64_t n_sld,
                        const wc_surfel* fix, int64_t n_fix, const wc_corr_idx* sld_corr, int64_t n_sld_corr,
                        const wc_corr_idx* fix_corr, int64_t n_fix_corr, const wc_imu_state* imu, int64_t n_imu,
                        const wc_sample_state* samples, int64_t K, double* cost, double* grad, double* jtj);

/* Build*Residuals + ceres::Solve, lidar_odometry.cc:254-363,541-561 (Ceres LM restated). */
int wco_window_solve(const wc_params* prm, const wc_solve_opts* opts, const wc_surfel* sld, int64_t n_sld,
                     const wc_surfel* fix, int64_t n_fix, const wc_corr_idx* sld_corr, int64_t n_sld_corr,
                     const wc_corr_idx* fix_corr, int64_t n_fix_corr, const wc_imu_state* imu, int64_t n_imu,
                     wc_sample_state* samples, int64_t K, wc_solve_summary* summary);

/* Single-factor hooks for finite-difference checks (cost_functor.h).  params: nblk blocks of 12.  jac: rows x
 * (12*nblk) row-major.  kind 0 = SurfelMatchUnaryFactor, 1 = SurfelMatchBinaryFactor (mode chosen from
 * sample indices), 2 = ImuFactor.  Returns number of residual rows, <0 on error. */
int wco_lidar_factor(const wc_params* prm, int jacobian_mode, const wc_surfel* s1, const wc_surfel* s2, int unary,
                     const double* sample_ts /*4: sp1l,sp1r,sp2l,sp2r*/, const int32_t* blk /*4 block ids*/,
                     const double* x /*K*12*/, int64_t K, double* residual, double* jac /*12K*/, double* weight,
                     double* normal3);
int wco_imu_factor(const wc_params* prm, const wc_imu_state* i3 /*3 states*/, const double* sample_ts /*3*/,
                   int mode, const double* grav3, const double* x /*3*12 (or 2*12)*/, double* residual12,
                   double* jac /*12 x 36 (or 24) row-major*/);

/* CubicBSplineInterpolator, spline_interpolation.h:42-113 */
void wco_spline_fit_eval(const double* ts, const double* pts3, int64_t K, const double* query_t, int64_t nq,
                         double* out3, uint8_t* valid, double* ctrl3 /* K*3 control points, may be NULL */);
/* CubicBSplineApprox / CubicSplineInterpolate (test-only helpers, spline_interpolation.h:9-40), N = 1 */
double wco_cubic_bspline_approx(double p_1, double p0, double p1, double p2, double s);
double wco_cubic_spline_interpolate(double s_1, double p_1, double s0, double p0, double s1, double p1, double s2,
                                    double p2, double s);

/* UpdateImuPoses + UpdateSamplePoses, lidar_odometry.cc:172-215 */
int wco_apply_corrections(wc_sample_state* samples, int64_t K, wc_imu_state* imu, int64_t n_imu);

/* PredictImuStatesAndSampleStates steps 2-3 (lidar_odometry.cc:403-453) with PredictPoseOfNewImuState (:112-123) */
int wco_predict_states(wc_imu_state* imu, int64_t n_imu, const double* ba3, const double* bg3, const double* grav3,
                       double t_last_sample, double sample_dt, int64_t n_new, wc_sample_state* samples_out);

/* UndistortSweep, lidar_odometry.cc:143-158 (a "next" row, used by the synthetic generator's checks) */
int wco_undistort_sweep(const wc_imu_state* imu, int64_t n_imu, const wc_point48* in, int64_t n, wc_point48* out);

/* AddLidarScan's per-point extrinsic + range / blind-box filter, lidar_odometry.cc:489-496; returns kept count or -status */
int64_t wco_filter_points(const wc_sweep_filter* f, const wc_point48* in, int64_t n, wc_point48* out);

/* utils.h / Sophus helpers for the known-answer tests: op 0 Exp (out: quat xyzw), 1 Log (in: quat xyzw,
 * out 3), 2 Jl, 3 Jl_inv, 4 Jr, 5 Jr_inv (out 9 row-major), 6 sym-eig (in 9, out 3 evals + 9 evecs) */
void wco_so3(int op, const double* in, double* out);

#ifdef __cplusplus
}
#endif
#endif
