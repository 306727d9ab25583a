// fp64 3-vector / 3x3 / quaternion / SO(3) device arithmetic for the window-odometry kernels.
// Follows the arithmetic the reference takes from Eigen and Sophus:
//   quaternion product / rotate / toRotationMatrix / slerp   (Eigen Geometry, un-vendored)
//   Exp / Log                                                  3rd-party/Sophus-1.22.10/sophus/so3.hpp:264-309,694-729
//   Hat / Jl / Jl_inv / Jr / Jr_inv                             src/common/utils.h:15-67
//   SelfAdjointEigenSolver<Matrix3d> -> cyclic Jacobi in registers (ascending eigenvalues)
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace wcd {

struct V3 {
  double x, y, z;
};
__host__ __device__ inline V3     mk(double x, double y, double z) { return V3{x, y, z}; }
__host__ __device__ inline V3     ld3(const double* p) { return V3{p[0], p[1], p[2]}; }
__host__ __device__ inline void   st3(double* p, const V3& v) { p[0] = v.x, p[1] = v.y, p[2] = v.z; }
__host__ __device__ inline V3     operator+(const V3& a, const V3& b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ inline V3     operator-(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ inline V3     operator-(const V3& a) { return V3{-a.x, -a.y, -a.z}; }
__host__ __device__ inline V3     operator*(double s, const V3& a) { return V3{s * a.x, s * a.y, s * a.z}; }
__host__ __device__ inline V3     operator*(const V3& a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
__host__ __device__ inline V3     operator/(const V3& a, double s) { return V3{a.x / s, a.y / s, a.z / s}; }
__host__ __device__ inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline V3     cross(const V3& a, const V3& b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__host__ __device__ inline double sqnorm(const V3& a) { return dot(a, a); }
__host__ __device__ inline double norm(const V3& a) { return sqrt(dot(a, a)); }
__host__ __device__ inline double get(const V3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

struct M3 {
  double m[3][3];
};
__host__ __device__ inline M3 zero3() {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = 0.0;
  return r;
}
__host__ __device__ inline M3 eye3() {
  M3 r = zero3();
  r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0;
  return r;
}
__host__ __device__ inline M3 ld33(const double* p) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = p[3 * i + j];
  return r;
}
__host__ __device__ inline void st33(double* p, const M3& a) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) p[3 * i + j] = a.m[i][j];
}
__host__ __device__ inline M3 operator*(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
__host__ __device__ inline V3 operator*(const M3& a, const V3& v) {
  return V3{a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
__host__ __device__ inline M3 operator+(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r;
}
__host__ __device__ inline M3 operator-(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j];
  return r;
}
__host__ __device__ inline M3 operator*(double s, const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = s * a.m[i][j];
  return r;
}
__host__ __device__ inline M3 transpose(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
__host__ __device__ inline M3 outer(const V3& a, const V3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = get(a, i) * get(b, j);
  return r;
}
// v^T A
__host__ __device__ inline V3 vTm(const V3& v, const M3& a) {
  return V3{v.x * a.m[0][0] + v.y * a.m[1][0] + v.z * a.m[2][0], v.x * a.m[0][1] + v.y * a.m[1][1] + v.z * a.m[2][1],
            v.x * a.m[0][2] + v.y * a.m[1][2] + v.z * a.m[2][2]};
}
__host__ __device__ inline V3 col(const M3& a, int j) { return V3{a.m[0][j], a.m[1][j], a.m[2][j]}; }

// utils.h:15-22
__host__ __device__ inline M3 Hat(const V3& v) {
  M3 r = zero3();
  r.m[0][1] = -v.z, r.m[0][2] = v.y;
  r.m[1][0] = v.z, r.m[1][2] = -v.x;
  r.m[2][0] = -v.y, r.m[2][1] = v.x;
  return r;
}

struct Q4 {
  double w, x, y, z;
};
__host__ __device__ inline Q4   ldq(const double* c) { return Q4{c[3], c[0], c[1], c[2]}; }  // Eigen coeffs(): x,y,z,w
__host__ __device__ inline void stq(double* c, const Q4& q) { c[0] = q.x, c[1] = q.y, c[2] = q.z, c[3] = q.w; }
__host__ __device__ inline V3   vec(const Q4& q) { return V3{q.x, q.y, q.z}; }
__host__ __device__ inline Q4   conj(const Q4& q) { return Q4{q.w, -q.x, -q.y, -q.z}; }
__host__ __device__ inline Q4   operator*(const Q4& a, const Q4& b) {
  return Q4{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
__host__ __device__ inline V3 operator*(const Q4& q, const V3& v) {
  V3 uv = cross(vec(q), v);
  uv    = uv + uv;
  return v + q.w * uv + cross(vec(q), uv);
}
__host__ __device__ inline M3 ToMatrix(const Q4& q) {
  M3           r;
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r.m[0][0] = 1 - (tyy + tzz), r.m[0][1] = txy - twz, r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz, r.m[1][1] = 1 - (txx + tzz), r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy, r.m[2][1] = tyz + twx, r.m[2][2] = 1 - (txx + tyy);
  return r;
}
__host__ __device__ inline Q4 Normalized(const Q4& q) {
  double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return Q4{q.w / n, q.x / n, q.y / n, q.z / n};
}
// Eigen slerp (lidar_odometry.cc:153,167)
__host__ __device__ inline Q4 Slerp(const Q4& a, double t, const Q4& b) {
  const double one  = 1.0 - 2.220446049250313e-16;
  double       d    = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
  double       absD = fabs(d);
  double       s0, s1;
  if (absD >= one) {
    s0 = 1 - t, s1 = t;
  } else {
    double theta = acos(absD), sinTheta = sin(theta);
    s0 = sin((1 - t) * theta) / sinTheta;
    s1 = sin(t * theta) / sinTheta;
  }
  if (d < 0) s1 = -s1;
  return Q4{s0 * a.w + s1 * b.w, s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z};
}
// Sophus SO3::exp
__host__ __device__ inline Q4 Exp(const V3& omega) {
  const double eps      = 1e-10;
  double       theta_sq = sqnorm(omega);
  double       imag, real;
  if (theta_sq < eps * eps) {
    double theta_po4 = theta_sq * theta_sq;
    imag             = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real             = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    double theta = sqrt(theta_sq), half = 0.5 * theta, sh, ch;
    sincos(half, &sh, &ch);
    imag = sh / theta;
    real = ch;
  }
  return Q4{real, imag * omega.x, imag * omega.y, imag * omega.z};
}
// Sophus SO3(q).log() (constructor normalises)
__host__ __device__ inline V3 Log(const Q4& q_in) {
  const double eps       = 1e-10;
  Q4           q         = Normalized(q_in);
  double       squared_n = sqnorm(vec(q));
  double       w         = q.w;
  double       k;
  if (squared_n < eps * eps) {
    double squared_w = w * w;
    k                = 2.0 / w - (2.0 / 3.0) * (squared_n) / (w * squared_w);
  } else {
    double n         = sqrt(squared_n);
    double atan_nbyw = (w < 0) ? atan2(-n, -w) : atan2(n, w);
    k                = 2.0 * atan_nbyw / n;
  }
  return k * vec(q);
}
// utils.h:32-67
__host__ __device__ inline M3 Jl(const V3& v) {
  const double tol = 1e-10;
  double       n   = norm(v);
  if (n > tol) {
    V3     a = v / n;
    double s, c;
    sincos(n, &s, &c);
    return (s / n) * eye3() + (1 - s / n) * outer(a, a) + ((1 - c) / n) * Hat(a);
  }
  return eye3();
}
__host__ __device__ inline M3 Jl_inv(const V3& v) {
  const double tol = 1e-10;
  double       n   = norm(v);
  if (n > tol) {
    double s, c;
    sincos(n / 2, &s, &c);
    M3 h = Hat(v);
    return eye3() - 0.5 * h + ((1 - n * c / 2 / s) / sqnorm(v)) * (h * h);
  }
  return eye3();
}
__host__ __device__ inline M3 Jr(const V3& v) { return Jl(-v); }
__host__ __device__ inline M3 Jr_inv(const V3& v) { return Jl_inv(-v); }

// Symmetric 3x3 eigen-decomposition in registers: cyclic Jacobi, eigenvalues ascending, eigenvectors in
// columns.  Input: a00,a01,a02,a11,a12,a22.
__host__ __device__ inline void SymEig3(double a00, double a01, double a02, double a11, double a12, double a22,
                                        double ev[3], M3& V) {
  double a[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2, r = 3 - p - q;
      if (a[p][q] == 0.0) continue;
      {  // an off-diagonal entry that cannot change either diagonal entry any more is rounded off (classic cyclic Jacobi)
        const double g = 100.0 * fabs(a[p][q]);
        if (fabs(a[p][p]) + g == fabs(a[p][p]) && fabs(a[q][q]) + g == fabs(a[q][q])) {
          a[p][q] = a[q][p] = 0.0;
          continue;
        }
      }
      // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)) with theta = tau / (2 a_pq), written with one division, one square
      // root and one reciprocal square root: t = sgn * 2|a_pq| / (|tau| + sqrt(tau^2 + 4 a_pq^2)), c = rsqrt(t^2 + 1)
      const double tau = a[q][q] - a[p][p], b = 2.0 * a[p][q];
      const double sg  = (tau == 0.0 || ((tau > 0.0) == (b > 0.0))) ? 1.0 : -1.0;
      double       t   = sg * fabs(b) / (fabs(tau) + sqrt(tau * tau + b * b));
      if (!isfinite(t)) t = 0.0;
#ifdef __CUDA_ARCH__
      const double c = rsqrt(t * t + 1.0), s = t * c;
#else
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#endif
      double apq = a[p][q];
      a[p][p] -= t * apq;
      a[q][q] += t * apq;
      a[p][q] = a[q][p] = 0.0;
      double arp = a[r][p], arq = a[r][q];
      a[r][p] = a[p][r] = c * arp - s * arq;
      a[r][q] = a[q][r] = s * arp + c * arq;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double vkp = v[k][p], vkq = v[k][q];
        v[k][p] = c * vkp - s * vkq;
        v[k][q] = s * vkp + c * vkq;
      }
    }
  }
  // sort ascending (3-element network), permuting eigenvector columns
  double d0 = a[0][0], d1 = a[1][1], d2 = a[2][2];
  int    i0 = 0, i1 = 1, i2 = 2;
  if (d1 < d0) { double t = d0; d0 = d1; d1 = t; int ti = i0; i0 = i1; i1 = ti; }
  if (d2 < d1) { double t = d1; d1 = d2; d2 = t; int ti = i1; i1 = i2; i2 = ti; }
  if (d1 < d0) { double t = d0; d0 = d1; d1 = t; int ti = i0; i0 = i1; i1 = ti; }
  ev[0] = d0, ev[1] = d1, ev[2] = d2;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    V.m[k][0] = i0 == 0 ? v[k][0] : (i0 == 1 ? v[k][1] : v[k][2]);
    V.m[k][1] = i1 == 0 ? v[k][0] : (i1 == 1 ? v[k][1] : v[k][2]);
    V.m[k][2] = i2 == 0 ? v[k][0] : (i2 == 1 ? v[k][1] : v[k][2]);
  }
}

// order-preserving map double -> uint64 (for 64-bit atomicMin/Max on timestamps)
__host__ __device__ inline unsigned long long OrderedBits(double d) {
#ifdef __CUDA_ARCH__
  unsigned long long b = (unsigned long long)__double_as_longlong(d);
#else
  unsigned long long b;
  memcpy(&b, &d, 8);
#endif
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double FromOrderedBits(unsigned long long u) {
  unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}

}  // namespace wcd
