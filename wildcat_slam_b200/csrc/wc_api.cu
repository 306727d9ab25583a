// Context management and defaults of the C ABI (include/wildcat_b200.h).
#include <math.h>
#include <stdlib.h>

#include "wc_ctx.h"

void wc_extract_free(wc_ctx* c);
void wc_match_free(wc_ctx* c);
void wc_solve_free(wc_ctx* c);
void wc_comm_free(wc_ctx* c);
void wc_spline_free(wc_ctx* c);
void wc_sweep_free(wc_ctx* c);

extern "C" int wc_abi_version(void) { return WC_ABI_VERSION; }

extern "C" void wc_default_params(wc_params* p) {
  memset(p, 0, sizeof(*p));
  // surfel_extraction.cc:327 (0.8, 2, {20,20,20,20}, 0.01, 0.1), :24 (0.05 s), :33 (20)
  p->voxel_size = 0.8f;
  p->max_layer  = 2;
  p->layer_point_size[0] = p->layer_point_size[1] = p->layer_point_size[2] = 20;
  p->cluster_min_points = 20;
  p->planer_threshold   = 0.01f;
  p->min_plane_likeness = 0.1;
  p->cluster_time_gap   = 0.05;
  // knn_surfel_matcher.h:37-41
  p->center_dist_threshold  = 1.0;
  p->angular_dist_threshold = 5.0 * M_PI / 180.0;
  p->surfel_dist_threshold  = 0.1;
  p->knn_candidates         = 10;
  p->time_diff_threshold    = 0.06;
  // lidar_odometry.cc:270,309; cost_functor.h:24,112
  p->cauchy_a     = 0.4;
  p->weight_floor = pow(0.05 / 6, 2);
  // lio_config.h:10-14,32,42-45
  const double gnd = 0.00015198973532354657, and_ = 0.006308226052016165, grw = 0.00011673723527962174,
               arw = 2.664506559330434e-06, w = 0.01;
  p->imu_rate   = 200;
  p->weight_gyr = 1 / (gnd * sqrt(p->imu_rate)) * w;
  p->weight_acc = 1 / (and_ * sqrt(p->imu_rate)) * w;
  p->weight_bg  = 1 / (grw / sqrt(p->imu_rate)) * w;
  p->weight_ba  = 1 / (arw / sqrt(p->imu_rate)) * w;
  p->max_points     = 1 << 21;
  p->max_surfels    = 1 << 18;
  p->max_corrs      = 1 << 19;
  p->max_samples    = 128;
  p->max_imu_states = 8192;
}

extern "C" void wc_default_solve_opts(wc_solve_opts* o) {
  // lidar_odometry.cc:551-554 sets three options; everything else is the Ceres default
  memset(o, 0, sizeof(*o));
  o->max_num_iterations          = 100;
  o->jacobian_mode               = WC_JAC_REFERENCE_OVERWRITE;
  o->fix_first_position          = 1;
  o->use_imu_factors             = 1;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius     = 1e16;
  o->min_trust_region_radius     = 1e-32;
  o->min_relative_decrease       = 1e-3;
  o->min_lm_diagonal             = 1e-6;
  o->max_lm_diagonal             = 1e32;
  o->function_tolerance          = 1e-6;
  o->gradient_tolerance          = 1e-10;
  o->parameter_tolerance         = 1e-8;
  o->precision                   = WC_PREC_F64;
}

extern "C" const char* wc_status_str(wc_status s) {
  switch (s) {
    case WC_OK: return "WC_OK";
    case WC_EINVAL: return "WC_EINVAL";
    case WC_EINVAL_TIME_ORDER: return "WC_EINVAL_TIME_ORDER";
    case WC_EOUT_OF_SPAN: return "WC_EOUT_OF_SPAN";
    case WC_ETOO_FEW_TARGETS: return "WC_ETOO_FEW_TARGETS";
    case WC_ECAPACITY: return "WC_ECAPACITY";
    case WC_ECUDA: return "WC_ECUDA";
    case WC_ECOMM: return "WC_ECOMM";
    case WC_ENUMERIC: return "WC_ENUMERIC";
  }
  return "WC_?";
}

extern "C" wc_status wc_create(const wc_params* p, int device, wc_ctx** out) {
  if (!out) return WC_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return WC_ECUDA;  // no CPU fallback
  wc_ctx* c = (wc_ctx*)calloc(1, sizeof(wc_ctx));
  if (!c) return WC_EINVAL;
  if (p) c->prm = *p; else wc_default_params(&c->prm);
  c->device = device;
  // every failure below goes through wc_destroy, which tolerates the still-zeroed fields
  bool ok = cudaSetDevice(device) == cudaSuccess;
  cudaDeviceProp prop;
  ok = ok && cudaGetDeviceProperties(&prop, device) == cudaSuccess;
  if (ok) c->num_sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : WC_NUM_SMS_FALLBACK;
  ok = ok && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; ok && i < 3; ++i)
    ok = cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < 8; ++i) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
  if (!ok) {
    wc_destroy(c);
    return WC_ECUDA;
  }
  c->rank = 0, c->world = 1;
  c->knn_grid_min = 512;
  c->lm_batch = 8;
  if (const char* e = getenv("WC_LM_BATCH")) c->lm_batch = atoi(e);
  if (const char* e = getenv("WC_KNN_GRID_MIN")) c->knn_grid_min = atoll(e);  // test hook: force either exact search
  *out = c;
  return WC_OK;
}

extern "C" void wc_destroy(wc_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 3; ++i)
    if (c->side[i]) cudaStreamSynchronize(c->side[i]);
  wc_comm_free(c);
  wc_extract_free(c);
  wc_match_free(c);
  wc_solve_free(c);
  wc_spline_free(c);
  wc_sweep_free(c);
  for (int i = 0; i < 8; ++i)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 3; ++i) {
    if (c->side[i]) cudaStreamDestroy(c->side[i]);
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->stream) cudaStreamDestroy(c->stream);
  free(c);
}

extern "C" const char* wc_last_error(const wc_ctx* c) { return c ? c->err : "null context"; }
extern "C" void*       wc_stream(wc_ctx* c) { return c ? (void*)c->stream : nullptr; }

// Pinned host memory for callers that want fast, truly asynchronous transfers of the POD arrays they pass to the
// host-buffer entry points (bench.py's e2e leg, the Python mirror's persistent output buffers).
extern "C" void* wc_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void wc_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
