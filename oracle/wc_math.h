// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under wildcat_slam_b200/ may
// include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker / CPU baseline.
//
// wc_math.h — fp64 small-matrix / quaternion / SO(3) arithmetic the reference takes from Eigen and
// Sophus, restated without either library (neither is installed here; Sophus is vendored under
// /root/reference/3rd-party/Sophus-1.22.10 but needs Eigen).
//
//   Quaternion product / rotate / toRotationMatrix / slerp : Eigen 3.3 Geometry/Quaternion.h (un-vendored,
//       version unpinned by the reference: CMakeLists.txt:28-34 has find_package without versions).
//   Exp / Log                                             : sophus/so3.hpp:264-309, 694-729 (vendored)
//   Hat / Jl / Jl_inv / Jr / Jr_inv                        : src/common/utils.h:15-67
//   SelfAdjointEigenSolver<Matrix3d>                       : replaced by cyclic Jacobi (same eigen-pairs to
//       fp64 round-off, ascending order; Eigen's QL iteration itself cannot be restated bit-for-bit).
//
// Compile with -ffp-contract=off: the reference's default x86-64 build has no FMA contraction.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

namespace wco {

struct V3 {
  double x = 0, y = 0, z = 0;
  V3() {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  explicit V3(const double* p) : x(p[0]), y(p[1]), z(p[2]) {}
  double  operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  void    store(double* p) const { p[0] = x, p[1] = y, p[2] = z; }
};
inline V3     operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3     operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3     operator-(const V3& a) { return {-a.x, -a.y, -a.z}; }
inline V3     operator*(double s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3     operator*(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3     operator/(const V3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3     cross(const V3& a, const V3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double sqnorm(const V3& a) { return dot(a, a); }
inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }

struct M3 {
  double m[3][3];
  M3() { std::memset(m, 0, sizeof(m)); }
  static M3 Identity() {
    M3 r;
    r.m[0][0] = r.m[1][1] = r.m[2][2] = 1;
    return r;
  }
  static M3 FromRowMajor(const double* p) {
    M3 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r.m[i][j] = p[3 * i + j];
    return r;
  }
  void store(double* p) const {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) p[3 * i + j] = m[i][j];
  }
  V3 col(int j) const { return {m[0][j], m[1][j], m[2][j]}; }
  V3 row(int i) const { return {m[i][0], m[i][1], m[i][2]}; }
};
inline M3 operator*(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
inline V3 operator*(const M3& a, const V3& v) { return {dot(a.row(0), v), dot(a.row(1), v), dot(a.row(2), v)}; }
inline M3 operator+(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r;
}
inline M3 operator-(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j];
  return r;
}
inline M3 operator*(double s, const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = s * a.m[i][j];
  return r;
}
inline M3 operator/(const M3& a, double s) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] / s;
  return r;
}
inline M3 transpose(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
inline M3 outer(const V3& a, const V3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a[i] * b[j];
  return r;
}
// row vector times matrix: (v^T A)
inline V3 vTm(const V3& v, const M3& a) { return {dot(v, a.col(0)), dot(v, a.col(1)), dot(v, a.col(2))}; }

// utils.h:15-22
inline M3 Hat(const V3& v) {
  M3 r;
  r.m[0][1] = -v.z, r.m[0][2] = v.y;
  r.m[1][0] = v.z, r.m[1][2] = -v.x;
  r.m[2][0] = -v.y, r.m[2][1] = v.x;
  return r;
}

// Eigen::Quaterniond; stored w,x,y,z here, (x,y,z,w) in the ABI structs (Eigen coeffs() order).
struct Q4 {
  double w = 1, x = 0, y = 0, z = 0;
  Q4() {}
  Q4(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}
  static Q4 FromCoeffs(const double* c) { return Q4(c[3], c[0], c[1], c[2]); }
  void      storeCoeffs(double* c) const { c[0] = x, c[1] = y, c[2] = z, c[3] = w; }
  V3        vec() const { return {x, y, z}; }
  Q4        conjugate() const { return Q4(w, -x, -y, -z); }
};
// Eigen quat product (Quaternion.h, internal::quat_product generic)
inline Q4 operator*(const Q4& a, const Q4& b) {
  return Q4(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x);
}
// Eigen QuaternionBase::_transformVector
inline V3 operator*(const Q4& q, const V3& v) {
  V3 uv = cross(q.vec(), v);
  uv    = uv + uv;
  return v + q.w * uv + cross(q.vec(), uv);
}
// Eigen QuaternionBase::toRotationMatrix
inline M3 ToMatrix(const Q4& q) {
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
inline Q4 Normalized(const Q4& q) {
  double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return Q4(q.w / n, q.x / n, q.y / n, q.z / n);
}
// Eigen QuaternionBase::slerp
inline Q4 Slerp(const Q4& a, double t, const Q4& b) {
  const double one  = 1.0 - 2.220446049250313e-16;
  double       d    = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
  double       absD = std::fabs(d);
  double       s0, s1;
  if (absD >= one) {
    s0 = 1 - t, s1 = t;
  } else {
    double theta = std::acos(absD), sinTheta = std::sin(theta);
    s0 = std::sin((1 - t) * theta) / sinTheta;
    s1 = std::sin(t * theta) / sinTheta;
  }
  if (d < 0) s1 = -s1;
  return Q4(s0 * a.w + s1 * b.w, s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z);
}

// Sophus SO3::expAndTheta, so3.hpp:694-729; utils.h:24-26
inline Q4 Exp(const V3& omega) {
  const double eps      = 1e-10;  // Sophus::Constants<double>::epsilon()
  double       theta_sq = sqnorm(omega);
  double       imag, real;
  if (theta_sq < eps * eps) {
    double theta_po4 = theta_sq * theta_sq;
    imag             = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real             = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    double theta = std::sqrt(theta_sq), half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  return Q4(real, imag * omega.x, imag * omega.y, imag * omega.z);
}
// Sophus SO3(quat) normalises, then logAndTheta, so3.hpp:264-309; utils.h:28-30
inline V3 Log(const Q4& q_in) {
  const double eps       = 1e-10;
  Q4           q         = Normalized(q_in);
  double       squared_n = sqnorm(q.vec());
  double       w         = q.w;
  double       two_atan_nbyw_by_n;
  if (squared_n < eps * eps) {
    double squared_w   = w * w;
    two_atan_nbyw_by_n = 2.0 / w - (2.0 / 3.0) * (squared_n) / (w * squared_w);
  } else {
    double n           = std::sqrt(squared_n);
    double atan_nbyw   = (w < 0) ? std::atan2(-n, -w) : std::atan2(n, w);
    two_atan_nbyw_by_n = 2.0 * atan_nbyw / n;
  }
  return two_atan_nbyw_by_n * q.vec();
}

// utils.h:32-67
inline M3 Jl(const V3& v) {
  const double tol = 1e-10;
  if (norm(v) > tol) {
    double theta = norm(v);
    V3     a     = v / theta;
    return (std::sin(theta) / theta) * M3::Identity() + (1 - std::sin(theta) / theta) * outer(a, a) +
           ((1 - std::cos(theta)) / theta) * Hat(a);
  }
  return M3::Identity();
}
inline M3 Jl_inv(const V3& v) {
  const double tol = 1e-10;
  if (norm(v) > tol) {
    double n = norm(v);
    return M3::Identity() - 0.5 * Hat(v) + ((1 - n * std::cos(n / 2) / 2 / std::sin(n / 2)) * (Hat(v) * Hat(v))) / sqnorm(v);
  }
  return M3::Identity();
}
inline M3 Jr(const V3& v) { return Jl(-v); }
inline M3 Jr_inv(const V3& v) { return Jl_inv(-v); }

// Symmetric 3x3 eigen-decomposition, eigenvalues ascending, eigenvectors in columns (the contract of
// Eigen::SelfAdjointEigenSolver<Matrix3d>, used at surfel_extraction.cc:49,98 and cost_functor.h:23,111).
// Cyclic Jacobi; reads the lower triangle like Eigen does.
inline void SymEig3(const M3& A_in, double evals[3], M3& evecs) {
  double a[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j <= i; ++j) a[i][j] = a[j][i] = A_in.m[i][j];
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t     = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        if (!std::isfinite(theta)) t = 0.0;  // |theta| overflow: rotation angle ~ 0
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        double app = a[p][p], aqq = a[q][q], apq = a[p][q];
        a[p][p] = app - t * apq;
        a[q][q] = aqq + t * apq;
        a[p][q] = a[q][p] = 0.0;
        int r             = 3 - p - q;
        double arp = a[r][p], arq = a[r][q];
        a[r][p] = a[p][r] = c * arp - s * arq;
        a[r][q] = a[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; ++k) {
          double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int    idx[3] = {0, 1, 2};
  double d[3]   = {a[0][0], a[1][1], a[2][2]};
  std::sort(idx, idx + 3, [&](int i, int j) { return d[i] < d[j]; });
  for (int j = 0; j < 3; ++j) {
    evals[j] = d[idx[j]];
    for (int k = 0; k < 3; ++k) evecs.m[k][j] = v[k][idx[j]];
  }
}

}  // namespace wco
