/*
 * wildcat_b200.h — C ABI of the B200-native Wildcat sliding-window odometry hot path.
 *
 * The reference (kekeliu-whu/Wildcat-SLAM) has no plugin/FFI layer: the hot path is the set of C++
 * symbols called from LidarOdometry::AddLidarScan (src/odometry/lidar_odometry.cc:523-566).
 * Every entry point below names the reference interface it replaces.  All structs are plain C PODs
 * with fixed layout (static-asserted in wc_types_check.h and in the Python ctypes mirror); no torch /
 * Eigen / STL types cross this boundary.  The caller owns every host buffer; a wc_ctx owns device
 * memory, streams and (for multi-GPU) peer mappings.  Calls are synchronous on return.  Violated
 * reference CHECK() preconditions become wc_status error codes instead of aborting the process.
 *
 * There is NO CPU fallback behind this ABI: every compute entry point requires a CUDA device and
 * fails with WC_ECUDA otherwise.
 */
#ifndef WILDCAT_B200_H_
#define WILDCAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WC_ABI_VERSION 2

typedef enum wc_status {
  WC_OK                = 0,
  WC_EINVAL            = 1, /* null pointer / bad size / bad option                                  */
  WC_EINVAL_TIME_ORDER = 2, /* CHECK(pt.time >= back().time) lidar_odometry.cc:491; s1.t < s2.t :256  */
  WC_EOUT_OF_SPAN      = 3, /* surfel/IMU time outside the sample-state span  lidar_odometry.cc:259-266*/
  WC_ETOO_FEW_TARGETS  = 4, /* <k targets for kNN (UB in the reference, knn_surfel_matcher.cc:60-62)   */
  WC_ECAPACITY         = 5, /* caller output buffer or ctx capacity too small                         */
  WC_ECUDA             = 6, /* CUDA runtime error / no device                                         */
  WC_ECOMM             = 7, /* multi-GPU peer mapping / exchange error                                */
  WC_ENUMERIC          = 8  /* non-finite cost, failed factorisation beyond the LM retry budget       */
} wc_status;

/* ------------------------------------------------------------------------------------------------
 * Value types (mirror src/common/common.h and src/odometry/surfel.h field for field)
 * ---------------------------------------------------------------------------------------------- */

/* hilti_ros::Point, src/common/common.h:12-28: PCL_ADD_POINT4D (x,y,z,pad), float intensity,
 * double time, uint16 ring; EIGEN_ALIGN16 => sizeof == 48. */
typedef struct wc_point48 {
  float    x, y, z, pad;
  float    intensity;
  float    _pad0;
  double   time;
  uint16_t ring;
  uint16_t _pad1[7];
} wc_point48;

/* Surfel, src/odometry/surfel.h:35-127.  rot is Eigen coefficient order (x,y,z,w).  covariance is the
 * full symmetric 3x3 (row-major == column-major).  center/covariance/norm are in the WORLD frame when
 * is_in_body_frame == 0 (as constructed by BuildSurfels) and in the BODY frame after UpdatePose. */
typedef struct wc_surfel {
  double  timestamp;
  double  resolution;
  double  plane_std_deviation;
  double  rot[4];
  double  pos[3];
  double  center[3];
  double  covariance[9];
  double  norm[3];
  int32_t is_in_body_frame;
  int32_t _pad;
} wc_surfel; /* 208 bytes */

/* SurfelCorrespondence, surfel.h:124-127, as indices.  For sliding-window correspondences both
 * indices address the sliding-window surfel array.  For fixed-window correspondences s1 addresses the
 * fixed-window array and s2 the sliding-window array.  timestamp(s1) < timestamp(s2). */
typedef struct wc_corr_idx {
  int32_t s1;
  int32_t s2;
} wc_corr_idx;

/* SampleState ("control pose"), surfel.h:9-23.  data_cor = [rot_cor(3), pos_cor(3), bg(3), ba(3)]. */
typedef struct wc_sample_state {
  double timestamp;
  double data_cor[12];
  double grav[3];
  double rot[4]; /* x,y,z,w */
  double pos[3];
} wc_sample_state; /* 184 bytes */

/* ImuState, surfel.h:25-33. */
typedef struct wc_imu_state {
  double timestamp;
  double pos[3];
  double rot[4]; /* x,y,z,w */
  double acc[3];
  double gyr[3];
} wc_imu_state; /* 112 bytes */

/* Per-point assignment written by wc_build_surfels when requested (test/diagnostic output; the
 * reference has no such output, it is what "voxel/surfel index assignment bit-exact" is checked on).
 * vx,vy,vz = VoxelLoc (surfel_extraction.h:59-64); leaf = 8*code(layer1)+code(layer2) with
 * code = 4[x>cx]+2[y>cy]+[z>cz] (surfel_extraction.cc:148-158). */
typedef struct wc_point_assign {
  int32_t vx, vy, vz;
  int32_t leaf;
} wc_point_assign;

/* ------------------------------------------------------------------------------------------------
 * Parameters.  Defaults (wc_default_params) are the reference's compile-time constants.
 * ---------------------------------------------------------------------------------------------- */
typedef struct wc_params {
  /* surfel extraction, surfel_extraction.cc:327 and :24,33 */
  float   voxel_size;            /* 0.8f (float on purpose, Q2)                     */
  int32_t max_layer;             /* 2  (only 2 is supported by the CUDA path)       */
  int32_t layer_point_size[3];   /* {20,20,20}: node analysed iff n >  this         */
  int32_t cluster_min_points;    /* 20: cluster kept iff n >= this                  */
  float   planer_threshold;      /* 0.01f                                           */
  double  min_plane_likeness;    /* 0.1                                             */
  double  cluster_time_gap;      /* 0.05 s                                          */
  double  view_point[3];         /* 0,0,0                                           */
  /* matcher, knn_surfel_matcher.h:37-41 */
  double  center_dist_threshold;  /* 1.0            */
  double  angular_dist_threshold; /* 5 pi / 180     */
  double  surfel_dist_threshold;  /* 0.1            */
  int32_t knn_candidates;         /* 10             */
  int32_t _pad0;
  double  time_diff_threshold;    /* 0.06           */
  /* factors, cost_functor.h:24,112; lidar_odometry.cc:270,309; lio_config.h:10-14,32,40-45 */
  double  cauchy_a;               /* 0.4            */
  double  weight_floor;           /* (0.05/6)^2     */
  double  imu_rate;               /* 200            */
  double  weight_gyr, weight_acc, weight_bg, weight_ba;
  /* capacities of the device context */
  int64_t max_points;             /* points per wc_build_surfels call               */
  int64_t max_surfels;            /* surfels per window (sliding or fixed)          */
  int64_t max_corrs;              /* correspondences per window                     */
  int32_t max_samples;            /* control poses per window                       */
  int32_t max_imu_states;
} wc_params;

/* arithmetic of the fused lidar residual + Jacobian + J^T J kernel (BASELINE config 5: fp64 vs fp32 tolerance sweep).
 * The IMU factors, the normal equations the tiles are summed into and the LM step are always fp64. */
enum { WC_PREC_F64 = 0,   /* fp64 records, evaluation and accumulation (the reference's arithmetic; default)      */
       WC_PREC_MIXED = 1, /* fp32 records (64 B instead of 128 B) and evaluation, fp64 J^T J accumulation           */
       WC_PREC_F32 = 2 }; /* fp32 records, evaluation and per-tile J^T J accumulation, fp64 sums across tiles       */

enum { WC_JAC_REFERENCE_OVERWRITE = 0, /* reproduce cost_functor.h:152-175 aliasing (Q1) */
       WC_JAC_EXACT = 1 };

/* ceres::Solver::Options as set / defaulted at lidar_odometry.cc:551-554. */
typedef struct wc_solve_opts {
  int32_t max_num_iterations;          /* 100  */
  int32_t jacobian_mode;               /* WC_JAC_REFERENCE_OVERWRITE */
  int32_t fix_first_position;          /* SubsetParameterization(12,{3,4,5}) on sample 0, :556-560 */
  int32_t use_imu_factors;             /* 1 */
  double  initial_trust_region_radius; /* 1e4  */
  double  max_trust_region_radius;     /* 1e16 */
  double  min_trust_region_radius;     /* 1e-32 */
  double  min_relative_decrease;       /* 1e-3 */
  double  min_lm_diagonal;             /* 1e-6 */
  double  max_lm_diagonal;             /* 1e32 */
  double  function_tolerance;          /* 1e-6 */
  double  gradient_tolerance;          /* 1e-10 */
  double  parameter_tolerance;         /* 1e-8 */
  int32_t precision;                   /* WC_PREC_F64 */
  int32_t _pad;
} wc_solve_opts;

enum { WC_TERM_NO_CONVERGENCE = 0, WC_TERM_FUNCTION_TOL = 1, WC_TERM_GRADIENT_TOL = 2,
       WC_TERM_PARAMETER_TOL = 3, WC_TERM_MIN_RADIUS = 4, WC_TERM_FAILURE = 5 };

#define WC_MAX_ITER_LOG 128
typedef struct wc_solve_summary {
  double  initial_cost;
  double  final_cost;
  int32_t num_iterations;       /* LM iterations executed (successful + unsuccessful)          */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t termination;
  int32_t num_residual_blocks_sld, num_residual_blocks_fix, num_residual_blocks_imu;
  int32_t num_linearizations;   /* passes of the fused residual+Jacobian+JtJ kernel            */
  double  iter_cost[WC_MAX_ITER_LOG];      /* candidate cost evaluated at iteration i (1-based) */
  double  iter_radius[WC_MAX_ITER_LOG];
  int8_t  iter_accepted[WC_MAX_ITER_LOG];
  double  gpu_ms_total;         /* CUDA-event time of the whole solve on the ctx stream        */
  double  gpu_ms_linearize;     /* sum over passes of the fused kernel (+reduction/collective) */
} wc_solve_summary;

typedef struct wc_ctx wc_ctx;

/* ------------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------------- */
int         wc_abi_version(void);
void        wc_default_params(wc_params* p);
void        wc_default_solve_opts(wc_solve_opts* o);
wc_status   wc_create(const wc_params* p, int device, wc_ctx** out);
void        wc_destroy(wc_ctx* ctx);
const char* wc_last_error(const wc_ctx* ctx);
const char* wc_status_str(wc_status s);
/* page-locked host memory (cudaMallocHost) for the caller's POD arrays: H2D / D2H of pinned buffers run at full PCIe rate */
void*       wc_host_alloc(size_t bytes);
void        wc_host_free(void* p);
/* the CUDA stream (cudaStream_t) all work of this ctx is issued on; for CUDA-event timing by callers */
void*       wc_stream(wc_ctx* ctx);

/* ------------------------------------------------------------------------------------------------
 * Surfel extraction — replaces
 *   void BuildSurfels(const std::vector<hilti_ros::Point>&, std::deque<Surfel::Ptr>&, GlobalMap&)
 *   src/odometry/surfel_extraction.h:145-147, .cc:316-337
 * pts: n points with non-decreasing time (WC_EINVAL_TIME_ORDER otherwise).  out: capacity cap.
 * Surfels are returned sorted by (timestamp, resolution descending, center) — the reference's
 * std::sort orders by timestamp only and leaves ties unspecified (Q5).  assign (optional, may be NULL):
 * n records.  gpu_ms (optional): CUDA-event time of the device work, excluding H2D/D2H.
 * ---------------------------------------------------------------------------------------------- */
wc_status wc_build_surfels(wc_ctx* ctx, const wc_point48* pts, size_t n, wc_surfel* out, size_t cap,
                           size_t* n_out, wc_point_assign* assign, double* gpu_ms);

/* Device-resident variant used by the benchmark's "inputs already in HBM" leg: upload once, then run
 * extraction repeatedly without host traffic.  Results stay on the device until fetched. */
wc_status wc_points_upload(wc_ctx* ctx, const wc_point48* pts, size_t n);
/* Streaming ingestion (no reference counterpart: the reference processes one rosbag message at a time,
 * wildcat_slam_node.cc:83-99): copies the NEXT sweep to the device on a copy stream while the window pass of the current
 * sweep runs; the NEXT wc_points_upload, if it names the same (pts, n), finds the points on the device (any other upload
 * drops the prefetch).  pts must stay valid and unchanged until that call; pinned memory (wc_host_alloc) makes the copy
 * asynchronous.  when: WC_PREFETCH_NOW starts the copy at once; WC_PREFETCH_AT_SOLVE starts it when the next
 * wc_window_pass_resident reaches its solve stage (the memory-bound extraction and matching stages are slowed by a
 * transfer running beside them, the solve is not); if no pass runs before the upload, the upload copies normally. */
#define WC_PREFETCH_NOW 0
#define WC_PREFETCH_AT_SOLVE 1
wc_status wc_points_prefetch(wc_ctx* ctx, const wc_point48* pts, size_t n, int when);
wc_status wc_build_surfels_resident(wc_ctx* ctx, size_t* n_out, double* gpu_ms_keys,
                                    double* gpu_ms_emit, double* gpu_ms_total);
wc_status wc_surfels_fetch(wc_ctx* ctx, wc_surfel* out, size_t cap, size_t* n_out);

/* ------------------------------------------------------------------------------------------------
 * Sweep preparation — the per-point steps immediately before BuildSurfels.
 *
 * wc_filter_points replaces the loop at the top of LidarOdometry::AddLidarScan (lidar_odometry.cc:489-496):
 *   p <- (float)(ext_lidar2imu * p); drop if |p| < min_range, |p| > max_range or p inside the blind box;
 *   the kept points keep their order.  WC_EINVAL_TIME_ORDER replaces the CHECK at :491 and is slightly stricter:
 *   the reference compares each point with the last KEPT point, this call requires the raw timestamps to be
 *   non-decreasing (identical on every valid sweep).
 * wc_undistort_sweep replaces UndistortSweep (lidar_odometry.cc:143-158): every point is moved to the world
 *   frame with the IMU pose interpolated at its own timestamp; WC_EOUT_OF_SPAN replaces the CHECK at :150.
 * wc_undistort_upload is the resident variant: raw sweep in, the undistorted sweep is left on the device in
 *   the layout wc_build_surfels_resident consumes (the undistorted 48-byte sweep is never materialised).
 * ---------------------------------------------------------------------------------------------- */
typedef struct wc_sweep_filter {
  double ext_q[4];          /* lidar -> IMU rotation, Eigen coeff order (x, y, z, w); lio_config.h:24-30 */
  double ext_t[3];          /* lidar -> IMU translation                                                 */
  double min_range;         /* 0.3    lio_config.h:19                                                  */
  double max_range;         /* 120    lio_config.h:18                                                  */
  double blind_box_min[3];  /* (-0.8, -0.5, -0.4)  lio_config.h:20-22, in imu_link                      */
  double blind_box_max[3];  /* ( 0.3,  0.5,  0.4)                                                       */
} wc_sweep_filter;
void      wc_default_sweep_filter(wc_sweep_filter* f);
wc_status wc_filter_points(wc_ctx* ctx, const wc_sweep_filter* f, const wc_point48* in, size_t n, wc_point48* out,
                           size_t cap, size_t* n_out);
wc_status wc_undistort_sweep(wc_ctx* ctx, const wc_imu_state* imu, size_t n_imu, const wc_point48* in, size_t n,
                             wc_point48* out);
wc_status wc_undistort_upload(wc_ctx* ctx, const wc_imu_state* imu, size_t n_imu, const wc_point48* in, size_t n);

/* ------------------------------------------------------------------------------------------------
 * Wire ingestion — replaces pcl::fromROSMsg(*msg, *cloud) in HandleLidarMessage
 * (wildcat_slam_node.cc:46-52) for the point type registered at common.h:21-28: fields x, y, z, intensity
 * (FLOAT32), timestamp (FLOAT64), ring (UINT16) of a sensor_msgs/PointCloud2 payload, found by NAME and
 * datatype at arbitrary byte offsets inside a point_step-byte record, become the 48-byte hilti_ros::Point
 * records every other entry point consumes.  A field the message does not carry (offset -1) stays zero, as
 * pcl::fromROSMsg leaves it.  The raw payload is uploaded as it came off the wire and unpacked by a kernel.
 * ---------------------------------------------------------------------------------------------- */
typedef struct wc_pc2_layout {
  uint32_t point_step;                                       /* bytes per point in the message            */
  int32_t  off_x, off_y, off_z, off_intensity;               /* FLOAT32 fields, byte offsets (-1: absent) */
  int32_t  off_time;                                         /* FLOAT64 "timestamp"                       */
  int32_t  off_ring;                                         /* UINT16 "ring"                             */
} wc_pc2_layout;
wc_status wc_unpack_pointcloud2(wc_ctx* ctx, const uint8_t* data, size_t n_points, const wc_pc2_layout* layout,
                                wc_point48* out);

/* ------------------------------------------------------------------------------------------------
 * Observability outputs — the numeric half of PrintSurfelResiduals /
 * PrintImuResiduals (lidar_odometry.cc:56-93: ceres::Problem::Evaluate with apply_loss_function = true, i.e. the
 * residuals after the Cauchy corrector) and of PubSurfels (surfel_extraction.cc:360-417: per-surfel marker pose,
 * scale and colour from the eigen-decomposition of the world covariance).  Formatting (Histogram::ToString,
 * histogram.cc:29-77) and message publishing stay on the host: wildcat_slam_b200/report.py.
 *
 * wc_window_residuals works on the window last handed to wc_window_upload / wc_window_solve* and evaluates at
 * `data_cor` (K x 12; NULL: the window's uploaded starting point).  Lidar residuals come back in the solver's
 * bucket order (a histogram does not care) with a per-residual flag telling the fixed-window (unary) blocks from
 * the sliding-window (binary) ones; IMU residuals are 12 per block (gyro, acc, gyro bias, acc bias x 3).
 * On a multi-GPU context every rank returns the residuals of its own block of correspondences.
 * ---------------------------------------------------------------------------------------------- */
typedef struct wc_marker {
  double position[3];     /* GetCenterInWorld                                                   */
  double orientation[4];  /* Eigen::Quaterniond(right-handed eigenvectors), coeff order x y z w */
  double scale[3];        /* 3 sqrt(eigenvalue), after makeRightHanded's swap                    */
  float  color[4];        /* ((n + 1) / 2, a = 1), n = GetNormInWorld                            */
} wc_marker;
wc_status wc_surfel_markers(wc_ctx* ctx, const wc_surfel* surfels, size_t n, wc_marker* out);
wc_status wc_window_residuals(wc_ctx* ctx, const wc_solve_opts* opts, const double* data_cor, double* lidar_res,
                              uint8_t* lidar_is_fix, size_t lidar_cap, size_t* n_lidar, double* imu_res, size_t imu_cap,
                              size_t* n_imu_blocks);

/* ------------------------------------------------------------------------------------------------
 * Surfel poses — replaces UpdateSurfelPoses (lidar_odometry.cc:160-170) + Surfel::UpdatePose
 * (surfel.h:48-58): interpolate the IMU pose at each surfel time (lerp / Eigen slerp) and move the
 * surfel to the body frame on first call.  In-place on host surfels.
 * ---------------------------------------------------------------------------------------------- */
wc_status wc_update_surfel_poses(wc_ctx* ctx, const wc_imu_state* imu, size_t n_imu, wc_surfel* surfels,
                                 size_t n);

/* ------------------------------------------------------------------------------------------------
 * Correspondence search — replaces KnnSurfelMatcher::BuildIndex + ::Match
 *   src/odometry/knn_surfel_matcher.h:17-19, .cc:3-49
 * self_match != 0: target set == query set (sliding-window matcher, lidar_odometry.cc:532-534), pair
 * de-duplication applies and out indices both address `query`.  Otherwise (fixed-window matcher,
 * :536-538) out.s1/s2 are ordered by time with s1 = the earlier one; `first_is_target[i]` (optional)
 * tells whether s1 is the target (always true in the reference's use because fixed surfels are older).
 * Ties in kNN distance are broken by the smaller target index (FLANN leaves them unspecified).
 * ---------------------------------------------------------------------------------------------- */
wc_status wc_match(wc_ctx* ctx, const wc_surfel* query, size_t nq, const wc_surfel* target, size_t nt,
                   int self_match, wc_corr_idx* out, size_t cap, size_t* n_out, uint8_t* first_is_target,
                   double* gpu_ms);

/* exact k nearest neighbours in the 6-D matcher feature space — replaces
 * KnnSurfelMatcher::FLANNBuildIndex/FLANNKNearestSearch (knn_surfel_matcher.cc:64-89), exposed because
 * the reference unit-tests exactly this (knn_surfel_matcher_test.cc:19-43).  feat: n x 6 doubles. */
wc_status wc_knn6(wc_ctx* ctx, const double* query6, size_t nq, const double* target6, size_t nt, int k,
                  int32_t* out_idx, double* out_dist2);

/* ------------------------------------------------------------------------------------------------
 * Window solve — replaces Build{SldWin,FixWin}LidarResiduals + BuildImuResiduals + ceres::Solve
 *   src/odometry/lidar_odometry.cc:254-363,541-561 with cost_functor.h factors.
 * samples[K].data_cor is read as the starting point and overwritten with the solution (in place,
 * like SampleState::data_cor).  Surfels must be in the body frame.  imu may be NULL / n_imu == 0.
 * gravity is taken from samples[K-1].grav (lidar_odometry.cc:341,355).
 * ---------------------------------------------------------------------------------------------- */
wc_status wc_window_solve(wc_ctx* ctx, const wc_surfel* sld, size_t n_sld, const wc_surfel* fix, size_t n_fix,
                          const wc_corr_idx* sld_corr, size_t n_sld_corr, const wc_corr_idx* fix_corr,
                          size_t n_fix_corr, const wc_imu_state* imu, size_t n_imu, wc_sample_state* samples,
                          size_t K, const wc_solve_opts* opts, wc_solve_summary* summary);

/* Device-resident variant: upload the window once (surfels, correspondences, IMU states, samples), then
 * solve repeatedly from the uploaded starting point. */
wc_status wc_window_upload(wc_ctx* ctx, const wc_surfel* sld, size_t n_sld, const wc_surfel* fix, size_t n_fix,
                           const wc_corr_idx* sld_corr, size_t n_sld_corr, const wc_corr_idx* fix_corr,
                           size_t n_fix_corr, const wc_imu_state* imu, size_t n_imu,
                           const wc_sample_state* samples, size_t K);
wc_status wc_window_solve_resident(wc_ctx* ctx, const wc_solve_opts* opts, wc_solve_summary* summary,
                                   double* data_cor_out /* K*12, may be NULL */);

/* Evaluate cost / gradient / JtJ at the given data_cor without stepping (test hook replacing
 * ceres::Problem::Evaluate, lidar_odometry.cc:56-94).  Outputs: cost = 1/2 sum rho(r^2); grad[12K];
 * jtj[12K*12K] row-major full symmetric; all for the robustified problem (Cauchy corrector applied). */
wc_status wc_window_evaluate(wc_ctx* ctx, const wc_surfel* sld, size_t n_sld, const wc_surfel* fix, size_t n_fix,
                             const wc_corr_idx* sld_corr, size_t n_sld_corr, const wc_corr_idx* fix_corr,
                             size_t n_fix_corr, const wc_imu_state* imu, size_t n_imu,
                             const wc_sample_state* samples, size_t K, const wc_solve_opts* opts, double* cost,
                             double* grad, double* jtj);


/* ------------------------------------------------------------------------------------------------
 * Device-resident window pass: steps 7-13 of LidarOdometry::AddLidarScan (lidar_odometry.cc:523-561) in one
 * call on data already in HBM.  wc_points_upload provides the sweep, wc_pass_upload the IMU states, the
 * sample states and the body-frame fixed-window surfels; the pass extracts the surfels, moves them to the body
 * frame, runs both matchers, assembles and solves.  Nothing but the summary crosses the host boundary.
 * ---------------------------------------------------------------------------------------------- */
typedef struct wc_pass_stats {
  double  ms_total, ms_extract, ms_extract_keys, ms_extract_emit, ms_match, ms_pack, ms_solve;
  int64_t n_surfels, n_sld_corr, n_fix_corr;
  int64_t n_launches; /* kernels launched by this ctx so far (cumulative) */
} wc_pass_stats;
wc_status wc_pass_upload(wc_ctx* ctx, const wc_imu_state* imu, size_t n_imu, const wc_sample_state* samples,
                         size_t K, const wc_surfel* fix, size_t n_fix);
wc_status wc_window_pass_resident(wc_ctx* ctx, const wc_solve_opts* opts, wc_solve_summary* summary,
                                  double* data_cor_out /* K*12, may be NULL */, wc_pass_stats* stats);

/* Windows resident across sweeps (second half; lidar_odometry.cc:527-528,228-250,574-580).
 * The reference keeps every surfel of the last ~6 s in surfels_sld_win_, appends each new sweep's surfels to it, updates
 * the poses of ALL of them, and matches / solves over the whole window; ShrinkToFit then moves the surfels older than the
 * first IMU state of the trimmed window to the front of the fixed window.  Here both windows live in HBM:
 *   wc_pass_upload_windows  like wc_pass_upload, plus the body-frame sliding-window surfels of the earlier sweeps.  With
 *                           WC_KEEP_SLD / WC_KEEP_FIX the window left on the device by the previous pass (and
 *                           wc_window_shrink) is kept and the corresponding host array is ignored: only the new sweep,
 *                           the IMU states and the sample states cross PCIe.
 *   wc_window_pass_resident appends the new sweep's surfels behind the earlier ones, then steps 8-13 on the whole window.
 *   wc_window_shrink        the surfel half of ShrinkToFit: sliding-window surfels with timestamp < t_front_imu move, newest
 *                           first, to the front of the fixed window (std::deque::push_front order).  trim_fixed = 0
 *                           reproduces the reference, whose trim loop compares back() with back() and never removes
 *                           anything; trim_fixed = 1 drops fixed-window surfels older than
 *                           fix_window_duration behind the newest one.
 *   wc_windows_fetch        copies both resident windows to the host (tests, hand-over to another context). */
enum { WC_KEEP_SLD = 1, WC_KEEP_FIX = 2 };
wc_status wc_pass_upload_windows(wc_ctx* ctx, const wc_imu_state* imu, size_t n_imu, const wc_sample_state* samples, size_t K,
                                 const wc_surfel* fix, size_t n_fix, const wc_surfel* sld_prev, size_t n_sld_prev, int keep_flags);
wc_status wc_window_shrink(wc_ctx* ctx, double t_front_imu, double fix_window_duration, int trim_fixed, size_t* n_sld,
                           size_t* n_fix);
wc_status wc_windows_fetch(wc_ctx* ctx, wc_surfel* sld, size_t sld_cap, size_t* n_sld, wc_surfel* fix, size_t fix_cap,
                           size_t* n_fix);
/* timing hook: `reps` peer-memory reductions of the packed normal equations of a K-pose window (multi-GPU ctx; collective) */
wc_status wc_comm_bench(wc_ctx* ctx, size_t K, int reps, double* ms_per_call);
/* number of CUDA kernels this ctx has launched since creation */
int64_t   wc_launch_count(const wc_ctx* ctx);

/* ------------------------------------------------------------------------------------------------
 * Spline — replaces CubicBSplineInterpolator(timestamps, points) + Interp(t)
 *   src/odometry/spline_interpolation.h:42-113.  valid[i] == 0 <=> Interp returned nullptr.
 * ---------------------------------------------------------------------------------------------- */
wc_status wc_spline_fit_eval(wc_ctx* ctx, const double* ts, const double* pts3, size_t K, const double* query_t,
                             size_t nq, double* out3, uint8_t* valid);

/* IMU forward prediction + new sample states — replaces steps 2-3 of
 * LidarOdometry::PredictImuStatesAndSampleStates (lidar_odometry.cc:403-453) with PredictPoseOfNewImuState (:112-123).
 * imu: n_imu states in/out; [0] and [1] carry poses (the tail of the window), [2..n) carry timestamp / acc / gyr and
 * receive rot = rot_prev * Exp(((gyr_prev + gyr) / 2 - bg) dt) and pos = (R_pp (acc_pp - ba) + grav) dt^2 + 2 pos_prev - pos_pp.
 * ba, bg, grav: those of the last sample state (:404-406).  samples_out: n_new sample states at
 * t_last_sample + i * sample_dt, i = 1..n_new: lerp / slerp pose of the bracketing IMU states (:439-449), biases and
 * gravity copied, pose corrections zero.  WC_EINVAL_TIME_ORDER replaces CHECK_NEAR :119 (uniform IMU spacing within
 * 1e-6 s), WC_EOUT_OF_SPAN the CHECK_NE at :441-442. */
wc_status wc_predict_states(wc_ctx* ctx, wc_imu_state* imu, size_t n_imu, const double* ba3, const double* bg3,
                            const double* grav3, double t_last_sample, double sample_dt, size_t n_new,
                            wc_sample_state* samples_out);

/* Post-solve updates — replaces UpdateImuPoses + UpdateSamplePoses (lidar_odometry.cc:172-215):
 * spreads the sample corrections over the IMU states with the cubic B-spline, re-predicts the last IMU
 * state, folds the corrections into the sample poses and zeroes them.  In place. */
wc_status wc_apply_corrections(wc_ctx* ctx, wc_sample_state* samples, size_t K, wc_imu_state* imu, size_t n_imu);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU residual sharding.  One process per GPU.  Each rank creates its ctx, exports
 * a 64-byte IPC handle for its exchange buffer, the host layer all-gathers the handles (torch.distributed
 * / any out-of-band channel) and every rank opens its peers.  Afterwards wc_window_* on every rank must
 * be called collectively with identical inputs; rank r linearises its contiguous slice of the
 * correspondence buckets and the packed [JtJ | Jtr | cost] buffer is all-reduced over NVLink peer
 * memory inside the reduction kernel (fixed summation order => bitwise identical steps on all ranks).
 * ---------------------------------------------------------------------------------------------- */
#define WC_IPC_HANDLE_BYTES 64
wc_status wc_comm_export(wc_ctx* ctx, uint8_t handle[WC_IPC_HANDLE_BYTES]);
wc_status wc_comm_connect(wc_ctx* ctx, int rank, int world, const uint8_t* all_handles /* world*64 */);
/* Sharded sweep upload (upload half).  After wc_comm_shard_upload(ctx, 1) the calls wc_points_upload and
 * wc_points_prefetch are COLLECTIVE: every rank passes the same (pts, n); each copies only its 1/world slab of the raw
 * sweep over its own PCIe link into its exported raw area, and every rank repacks each point from its owner's area over
 * NVLink.  Off by default (a rank may then upload on its own). */
wc_status wc_comm_shard_upload(wc_ctx* ctx, int on);
wc_status wc_comm_disconnect(wc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* WILDCAT_B200_H_ */
