// Sweep preparation on the device — SURVEY §8(f) rank 1, the step immediately before the hot path:
//   filter_points     the per-point loop of LidarOdometry::AddLidarScan (src/odometry/lidar_odometry.cc:489-496):
//                     p <- (float)(ext_lidar2imu * (double)p), drop if |p| < min_range, |p| > max_range (float norm) or
//                     inside the blind bounding box; order-preserving compaction in two launches
//   undistort_points  UndistortSweep (lidar_odometry.cc:143-158): lower_bound over the IMU states, lerp / Eigen slerp of
//                     the bracketing poses, p_w = (float)(R (double)p + t)
//   undistort_repack  the same, written straight into the resident float4 + double layout K1 consumes (the
//                     undistorted 48-byte sweep is never materialised)
#include "wc_ctx.h"
#include "wc_device_math.cuh"

using namespace wcd;

namespace {

__device__ __forceinline__ int imu_lower_bound_s(const wc_imu_state* __restrict__ imu, int n, double t) {
  int lo = 0, hi = n;  // first idx with imu[idx].timestamp >= t  (std::lower_bound, lidar_odometry.cc:148)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (imu[mid].timestamp < t) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct FilterArgs {
  double q[4];    // ext_lidar2imu rotation (x, y, z, w)
  double t[3];    // ext_lidar2imu translation
  double min_range, max_range;
  double box_min[3], box_max[3];
};

// transformed point + keep flag of raw point i
__device__ __forceinline__ bool filter_one(const wc_point48& in, const FilterArgs& a, float& ox, float& oy, float& oz) {
  const Q4 q = ldq(a.q);  // Eigen coefficient order (x, y, z, w)
  const V3 p = q * mk((double)in.x, (double)in.y, (double)in.z) + mk(a.t[0], a.t[1], a.t[2]);  // Rigid3d * point, :490
  ox = (float)p.x, oy = (float)p.y, oz = (float)p.z;
  // getVector3fMap().norm(): float squared norm summed in x, y, z order, correctly rounded square root  (:492)
  const float  n2 = __fadd_rn(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)), __fmul_rn(oz, oz));
  const double nr = (double)__fsqrt_rn(n2);
  if (nr < a.min_range || nr > a.max_range) return false;
  const double x = (double)ox, y = (double)oy, z = (double)oz;  // AlignedBox::contains: min <= p <= max
  const bool   in_box = x >= a.box_min[0] && x <= a.box_max[0] && y >= a.box_min[1] && y <= a.box_max[1] && z >= a.box_min[2] &&
                      z <= a.box_max[2];
  return !in_box;
}

__global__ void __launch_bounds__(1024) filter_count(const wc_point48* __restrict__ in, int n, FilterArgs a, int* __restrict__ blk_cnt,
                                                     int* __restrict__ err) {
  const int i = blockIdx.x * 1024 + threadIdx.x;
  bool      keep = false;
  if (i < n) {
    float x, y, z;
    keep = filter_one(in[i], a, x, y, z);
    if (i > 0 && in[i].time < in[i - 1].time) *err = WC_EINVAL_TIME_ORDER;  // CHECK :491
  }
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) blk_cnt[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) filter_scatter(const wc_point48* __restrict__ in, int n, FilterArgs a,
                                                       const int* __restrict__ blk_cnt, wc_point48* __restrict__ out,
                                                       int* __restrict__ n_out) {
  __shared__ int warp_sums[32];
  __shared__ int s_base;
  const int t = threadIdx.x, lane = t & 31, i = blockIdx.x * 1024 + t;
  if (t < 32) {
    int s = 0;
    for (int b = lane; b < (int)blockIdx.x; b += 32) s += blk_cnt[b];
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if (lane == 0) s_base = s;
  }
  float x = 0.f, y = 0.f, z = 0.f;
  const int v = (i < n) ? (int)filter_one(in[i], a, x, y, z) : 0;
  int incl = v;
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
    const uint4* src = reinterpret_cast<const uint4*>(in + i);
    uint4*       dst = reinterpret_cast<uint4*>(out + pos);
    uint4        r0 = src[0];
    r0.x = __float_as_uint(x), r0.y = __float_as_uint(y), r0.z = __float_as_uint(z);
    dst[0] = r0, dst[1] = src[1], dst[2] = src[2];
  }
  if (blockIdx.x == gridDim.x - 1 && t == 1023) *n_out = pos + v;
}

// the undistorted position of one point (float, as the reference stores it); false if outside the IMU span (CHECK :150)
__device__ __forceinline__ bool undistort_one(const wc_imu_state* __restrict__ imu, int n_imu, float x, float y, float z, double t,
                                              float& ox, float& oy, float& oz) {
  const int idx = imu_lower_bound_s(imu, n_imu, t);
  if (!(idx >= 1 && idx < n_imu)) return false;
  const wc_imu_state &a = imu[idx - 1], &b = imu[idx];
  const double        f   = (t - a.timestamp) / (b.timestamp - a.timestamp);
  const V3            pos = ld3(a.pos) * (1 - f) + ld3(b.pos) * f;
  const Q4            rot = Slerp(ldq(a.rot), f, ldq(b.rot));
  const V3            p   = rot * mk((double)x, (double)y, (double)z) + pos;
  ox = (float)p.x, oy = (float)p.y, oz = (float)p.z;
  return true;
}

__global__ void undistort_points(const wc_imu_state* __restrict__ imu, int n_imu, const wc_point48* __restrict__ in, int n,
                                 wc_point48* __restrict__ out, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* src = reinterpret_cast<const uint4*>(in + i);
  uint4        r0 = src[0], r1 = src[1], r2 = src[2];
  const double t  = __hiloint2double((int)r1.w, (int)r1.z);
  float        x, y, z;
  if (!undistort_one(imu, n_imu, __uint_as_float(r0.x), __uint_as_float(r0.y), __uint_as_float(r0.z), t, x, y, z)) {
    *err = WC_EOUT_OF_SPAN;
    return;
  }
  r0.x = __float_as_uint(x), r0.y = __float_as_uint(y), r0.z = __float_as_uint(z);
  uint4* dst = reinterpret_cast<uint4*>(out + i);
  dst[0] = r0, dst[1] = r1, dst[2] = r2;
}

__global__ void undistort_repack(const wc_imu_state* __restrict__ imu, int n_imu, const wc_point48* __restrict__ in, int n,
                                 float4* __restrict__ xyz, double* __restrict__ time, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4* src = reinterpret_cast<const float4*>(in + i);
  const float4  r0 = src[0], r1 = src[1];
  const double  t  = __hiloint2double(__float_as_int(r1.w), __float_as_int(r1.z));
  float         x, y, z;
  if (!undistort_one(imu, n_imu, r0.x, r0.y, r0.z, t, x, y, z)) {
    *err = WC_EOUT_OF_SPAN;
    return;
  }
  xyz[i]  = make_float4(x, y, z, r0.w);
  time[i] = t;
}


// ---- PointCloud2 payload -> 48-byte records (pcl::fromROSMsg for hilti_ros::Point, wildcat_slam_node.cc:49)
// One thread per point; the fields sit at arbitrary (possibly unaligned) byte offsets of the point_step-byte record.
__device__ __forceinline__ unsigned ld_u32_unaligned(const unsigned char* p) {
  return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
}
__global__ void unpack_pointcloud2(const unsigned char* __restrict__ data, int n, wc_pc2_layout L, wc_point48* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* r = data + (size_t)i * L.point_step;
  wc_point48           p;
  memset(&p, 0, sizeof(p));
  if (L.off_x >= 0) p.x = __uint_as_float(ld_u32_unaligned(r + L.off_x));
  if (L.off_y >= 0) p.y = __uint_as_float(ld_u32_unaligned(r + L.off_y));
  if (L.off_z >= 0) p.z = __uint_as_float(ld_u32_unaligned(r + L.off_z));
  if (L.off_intensity >= 0) p.intensity = __uint_as_float(ld_u32_unaligned(r + L.off_intensity));
  if (L.off_time >= 0)
    p.time = __longlong_as_double((long long)(((unsigned long long)ld_u32_unaligned(r + L.off_time + 4) << 32) | ld_u32_unaligned(r + L.off_time)));
  if (L.off_ring >= 0) p.ring = (uint16_t)((unsigned)r[L.off_ring] | ((unsigned)r[L.off_ring + 1] << 8));
  out[i] = p;
}

}  // namespace

struct wc_sweep_mem {
  wc_point48*   in;
  wc_point48*   out;
  wc_imu_state* imu;
  int*          blk_cnt;
  int*          flags;    // [0] n_out, [1] err
  int*          h_flags;  // pinned
  float4*       h_first;  // pinned: first undistorted point (anchors the relative voxel keys)
  double*       h_t0;
  unsigned char* raw;     // PointCloud2 payload staging (grown on demand)
  size_t        raw_cap;
};

static wc_status sweep_alloc(wc_ctx* c) {
  if (c->d_sweep) return WC_OK;
  wc_sweep_mem* m  = (wc_sweep_mem*)calloc(1, sizeof(wc_sweep_mem));
  c->d_sweep       = m;
  const size_t np  = (size_t)c->prm.max_points;
  WC_CUDA(c, cudaMalloc(&m->in, np * sizeof(wc_point48)));
  WC_CUDA(c, cudaMalloc(&m->out, np * sizeof(wc_point48)));
  WC_CUDA(c, cudaMalloc(&m->imu, (size_t)c->prm.max_imu_states * sizeof(wc_imu_state)));
  WC_CUDA(c, cudaMalloc(&m->blk_cnt, (np / 1024 + 2) * 4));
  WC_CUDA(c, cudaMalloc(&m->flags, 16));
  WC_CUDA(c, cudaMallocHost(&m->h_flags, 16));
  WC_CUDA(c, cudaMallocHost(&m->h_first, sizeof(float4)));
  WC_CUDA(c, cudaMallocHost(&m->h_t0, 8));
  return WC_OK;
}

void wc_sweep_free(wc_ctx* c) {
  wc_sweep_mem* m = (wc_sweep_mem*)c->d_sweep;
  if (!m) return;
  void* ptrs[] = {m->in, m->out, m->imu, m->blk_cnt, m->flags, m->raw};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (m->h_flags) cudaFreeHost(m->h_flags);
  if (m->h_first) cudaFreeHost(m->h_first);
  if (m->h_t0) cudaFreeHost(m->h_t0);
  free(m);
  c->d_sweep = nullptr;
}

extern "C" void wc_default_sweep_filter(wc_sweep_filter* f) {
  // lio_config.h:18-30
  memset(f, 0, sizeof(*f));
  f->min_range = 0.3, f->max_range = 120.0;
  f->blind_box_min[0] = -0.8, f->blind_box_min[1] = -0.5, f->blind_box_min[2] = -0.4;
  f->blind_box_max[0] = 0.3, f->blind_box_max[1] = 0.5, f->blind_box_max[2] = 0.4;
  f->ext_t[0] = -0.001, f->ext_t[1] = -0.00855, f->ext_t[2] = 0.055;
  // Eigen::Quaterniond of the rotation matrix [[-5.32125e-08, -1, 0], [-1, -5.32125e-08, 0], [0, 0, -1]]
  // (Eigen's matrix -> quaternion conversion, trace <= 0 branch with i = 0): stated by the caller for other sensors
  const double m00 = -5.32125e-08, m11 = -5.32125e-08, m22 = -1.0, m01 = -1.0, m10 = -1.0;
  const double t = sqrt(m00 - m11 - m22 + 1.0);  // i = 0, j = 1, k = 2
  f->ext_q[0] = 0.5 * t;
  const double s = 0.5 / t;
  f->ext_q[3] = (0.0 - 0.0) * s;         // (m(k,j) - m(j,k)) * t
  f->ext_q[1] = (m10 + m01) * s;
  f->ext_q[2] = (0.0 + 0.0) * s;
}

extern "C" wc_status wc_filter_points(wc_ctx* c, const wc_sweep_filter* f, const wc_point48* in, size_t n, wc_point48* out,
                                      size_t cap, size_t* n_out) {
  if (!c || !f || (n && (!in || !out)) || !n_out) return WC_EINVAL;
  *n_out = 0;
  if (n == 0) return WC_OK;
  if (n > (size_t)c->prm.max_points) WC_FAIL(c, WC_ECAPACITY, "n=%zu exceeds max_points", n);
  wc_status s = sweep_alloc(c);
  if (s) return s;
  wc_sweep_mem* m  = (wc_sweep_mem*)c->d_sweep;
  cudaStream_t  st = c->stream;
  FilterArgs    a;
  for (int k = 0; k < 4; ++k) a.q[k] = f->ext_q[k];
  for (int k = 0; k < 3; ++k) a.t[k] = f->ext_t[k], a.box_min[k] = f->blind_box_min[k], a.box_max[k] = f->blind_box_max[k];
  a.min_range = f->min_range, a.max_range = f->max_range;
  WC_CUDA(c, cudaMemcpyAsync(m->in, in, n * sizeof(wc_point48), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemsetAsync(m->flags, 0, 16, st));
  const unsigned g = (unsigned)((n + 1023) / 1024);
  { ++c->n_launches; filter_count<<<g, 1024, 0, st>>>(m->in, (int)n, a, m->blk_cnt, m->flags + 1); }
  { ++c->n_launches; filter_scatter<<<g, 1024, 0, st>>>(m->in, (int)n, a, m->blk_cnt, m->out, m->flags); }
  WC_CUDA(c, cudaMemcpyAsync(m->h_flags, m->flags, 16, cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  if (m->h_flags[1]) WC_FAIL(c, WC_EINVAL_TIME_ORDER, "point timestamps are not non-decreasing (lidar_odometry.cc:491)");
  const size_t k = (size_t)m->h_flags[0];
  *n_out         = k;
  if (k > cap) WC_FAIL(c, WC_ECAPACITY, "%zu kept points exceed the output capacity %zu", k, cap);
  if (k) WC_CUDA(c, cudaMemcpyAsync(out, m->out, k * sizeof(wc_point48), cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  return WC_OK;
}

extern "C" wc_status wc_undistort_sweep(wc_ctx* c, const wc_imu_state* imu, size_t n_imu, const wc_point48* in, size_t n,
                                        wc_point48* out) {
  if (!c || (n && (!imu || !in || !out))) return WC_EINVAL;
  if (n == 0) return WC_OK;
  if (n > (size_t)c->prm.max_points || n_imu > (size_t)c->prm.max_imu_states) WC_FAIL(c, WC_ECAPACITY, "capacity exceeded");
  wc_status s = sweep_alloc(c);
  if (s) return s;
  wc_sweep_mem* m  = (wc_sweep_mem*)c->d_sweep;
  cudaStream_t  st = c->stream;
  WC_CUDA(c, cudaMemcpyAsync(m->imu, imu, n_imu * sizeof(wc_imu_state), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemcpyAsync(m->in, in, n * sizeof(wc_point48), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemsetAsync(m->flags, 0, 16, st));
  { ++c->n_launches; undistort_points<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->imu, (int)n_imu, m->in, (int)n, m->out, m->flags + 1); }
  WC_CUDA(c, cudaMemcpyAsync(m->h_flags, m->flags, 16, cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaMemcpyAsync(out, m->out, n * sizeof(wc_point48), cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  if (m->h_flags[1]) WC_FAIL(c, WC_EOUT_OF_SPAN, "point timestamp outside the IMU state span (lidar_odometry.cc:150)");
  return WC_OK;
}

// raw (distorted, IMU-frame) sweep in, undistorted sweep resident for wc_build_surfels_resident — one upload, one kernel
extern "C" wc_status wc_undistort_upload(wc_ctx* c, const wc_imu_state* imu, size_t n_imu, const wc_point48* in, size_t n) {
  if (!c || (n && (!imu || !in))) return WC_EINVAL;
  if (n > (size_t)c->prm.max_points || n_imu > (size_t)c->prm.max_imu_states) WC_FAIL(c, WC_ECAPACITY, "capacity exceeded");
  wc_status s = wc_points_upload(c, nullptr, 0);  // allocates the extraction buffers
  if (s) return s;
  if ((s = sweep_alloc(c))) return s;
  c->n_pts = n;
  if (n == 0) return WC_OK;
  wc_sweep_mem* m  = (wc_sweep_mem*)c->d_sweep;
  cudaStream_t  st = c->stream;
  WC_CUDA(c, cudaMemcpyAsync(m->imu, imu, n_imu * sizeof(wc_imu_state), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemcpyAsync(m->in, in, n * sizeof(wc_point48), cudaMemcpyHostToDevice, st));
  WC_CUDA(c, cudaMemsetAsync(m->flags, 0, 16, st));
  { ++c->n_launches; undistort_repack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->imu, (int)n_imu, m->in, (int)n, c->d_xyz, c->d_time, m->flags + 1); }
  WC_CUDA(c, cudaMemcpyAsync(m->h_flags, m->flags, 16, cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaMemcpyAsync(m->h_first, c->d_xyz, sizeof(float4), cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  if (m->h_flags[1]) WC_FAIL(c, WC_EOUT_OF_SPAN, "point timestamp outside the IMU state span (lidar_odometry.cc:150)");
  const double vs = (double)c->prm.voxel_size;
  c->vox0[0] = (int)floor((double)m->h_first->x / vs), c->vox0[1] = (int)floor((double)m->h_first->y / vs),
  c->vox0[2] = (int)floor((double)m->h_first->z / vs);
  c->t_first = in[0].time, c->t_last = in[n - 1].time;
  return WC_OK;
}

extern "C" wc_status wc_unpack_pointcloud2(wc_ctx* c, const uint8_t* data, size_t n, const wc_pc2_layout* L, wc_point48* out) {
  if (!c || !L || (n && (!data || !out))) return WC_EINVAL;
  if (n == 0) return WC_OK;
  if (n > (size_t)c->prm.max_points) WC_FAIL(c, WC_ECAPACITY, "n=%zu exceeds max_points", n);
  const int32_t offs[6] = {L->off_x, L->off_y, L->off_z, L->off_intensity, L->off_time, L->off_ring};
  const int     size[6] = {4, 4, 4, 4, 8, 2};
  for (int k = 0; k < 6; ++k)
    if (offs[k] >= 0 && (size_t)offs[k] + size[k] > L->point_step) WC_FAIL(c, WC_EINVAL, "PointCloud2 field %d lies outside the %u-byte point", k, L->point_step);
  if (L->point_step == 0) WC_FAIL(c, WC_EINVAL, "point_step is zero");
  wc_status s = sweep_alloc(c);
  if (s) return s;
  wc_sweep_mem* m  = (wc_sweep_mem*)c->d_sweep;
  cudaStream_t  st = c->stream;
  const size_t  bytes = n * (size_t)L->point_step;
  if (bytes > m->raw_cap) {
    if (m->raw) cudaFree(m->raw);
    m->raw = nullptr, m->raw_cap = 0;
    WC_CUDA(c, cudaMalloc(&m->raw, bytes));
    m->raw_cap = bytes;
  }
  WC_CUDA(c, cudaMemcpyAsync(m->raw, data, bytes, cudaMemcpyHostToDevice, st));
  { ++c->n_launches; unpack_pointcloud2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m->raw, (int)n, *L, m->out); }
  WC_CUDA(c, cudaMemcpyAsync(out, m->out, n * sizeof(wc_point48), cudaMemcpyDeviceToHost, st));
  WC_CUDA(c, cudaStreamSynchronize(st));
  WC_CUDA(c, cudaGetLastError());
  return WC_OK;
}
