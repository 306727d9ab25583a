// ORACLE — TEST INFRASTRUCTURE ONLY (see wc_math.h / wc_oracle.h).
// Cubic B-spline interpolator and the post-solve pose updates restated from
//   src/odometry/spline_interpolation.h:9-113, src/odometry/lidar_odometry.cc:22-54,112-123,172-215.
// Pinned by spline_interpolation_test.cc:10-41,79-96 and scripts/CubicBSpline3D.ipynb (tests/golden/).
#include <cstdint>
#include <vector>

#include "wc_math.h"
#include <cmath>
#include <cstring>

#include "wc_oracle.h"

using namespace wco;

namespace {

// dense inverse by Gauss-Jordan with partial pivoting (Eigen's MatrixXd::inverse() is PartialPivLU based)
bool Invert(std::vector<double>& a, int n) {
  std::vector<double> inv((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[(size_t)r * n + c]) > std::fabs(a[(size_t)piv * n + c])) piv = r;
    if (a[(size_t)piv * n + c] == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) std::swap(a[(size_t)c * n + j], a[(size_t)piv * n + j]), std::swap(inv[(size_t)c * n + j], inv[(size_t)piv * n + j]);
    double d = a[(size_t)c * n + c];
    for (int j = 0; j < n; ++j) a[(size_t)c * n + j] /= d, inv[(size_t)c * n + j] /= d;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      double f = a[(size_t)r * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) a[(size_t)r * n + j] -= f * a[(size_t)c * n + j], inv[(size_t)r * n + j] -= f * inv[(size_t)c * n + j];
    }
  }
  a = inv;
  return true;
}

const double kM[4][4] = {{-1, 3, -3, 1}, {3, -6, 3, 0}, {-3, 0, 3, 0}, {1, 4, 1, 0}};  // spline_interpolation.h:83

struct Interpolator {
  std::vector<double> ts;
  int                 Np = 0;
  std::vector<double> Q;  // Np x 3

  // Init(), spline_interpolation.h:74-104
  void Init(const double* timestamps, const double* pts3, int n) {
    ts.assign(timestamps, timestamps + n);
    Np = n;
    std::vector<double> N((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) {
      double temp[4];
      for (int j = 0; j < 4; ++j) temp[j] = (0 * kM[0][j] + 0 * kM[1][j] + 0 * kM[2][j] + 1.0 * kM[3][j]) / 6.0;
      for (int j = 0; j < 4; ++j) {
        int idx = std::min(std::max(i - 1 + j, 0), n - 1);
        N[(size_t)i * n + idx] += temp[j];
      }
    }
    // (N^T N)^-1 N^T p
    std::vector<double> NtN((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int k = 0; k < n; ++k) s += N[(size_t)k * n + i] * N[(size_t)k * n + j];
        NtN[(size_t)i * n + j] = s;
      }
    Invert(NtN, n);
    std::vector<double> W((size_t)n * n, 0.0);  // inv(NtN) * N^T
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int k = 0; k < n; ++k) s += NtN[(size_t)i * n + k] * N[(size_t)j * n + k];
        W[(size_t)i * n + j] = s;
      }
    Q.assign((size_t)n * 3, 0.0);
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) {
        double s = 0;
        for (int k = 0; k < n; ++k) s += W[(size_t)i * n + k] * pts3[3 * k + c];
        Q[3 * i + c] = s;
      }
  }

  // Interp(), spline_interpolation.h:51-72
  bool Interp(double timestamp, double* out3) const {
    if (timestamp < ts.front() || timestamp > ts.back()) return false;
    double index_f   = (timestamp - ts.front()) / (ts.back() - ts.front()) * (double)(Np - 1) + 1.0;
    int    index_int = (int)std::floor(index_f);
    double t         = index_f - index_int;
    double tv[4]     = {t * t * t, t * t, t, 1.0};
    double tm[4];
    for (int j = 0; j < 4; ++j) tm[j] = tv[0] * kM[0][j] + tv[1] * kM[1][j] + tv[2] * kM[2][j] + tv[3] * kM[3][j];
    for (int c = 0; c < 3; ++c) {
      double s = 0;
      for (int j = 0; j < 4; ++j) {
        int idx = std::min(std::max(index_int - 2 + j, 0), Np - 1);
        s += tm[j] * Q[3 * idx + c];
      }
      out3[c] = s / 6.0;
    }
    return true;
  }
};

// PredictPoseOfNewImuState, lidar_odometry.cc:112-123
void PredictPoseOfNewImuState(const wc_imu_state& i1, const wc_imu_state& i2, const V3& ba, const V3& bg,
                              const V3& grav, wc_imu_state& i3) {
  double dt  = i3.timestamp - i2.timestamp;
  Q4     rot = Q4::FromCoeffs(i2.rot) * Exp(((V3(i2.gyr) + V3(i3.gyr)) / 2 - bg) * dt);
  V3     pos = (Q4::FromCoeffs(i1.rot) * (V3(i1.acc) - ba) + grav) * dt * dt + 2 * V3(i2.pos) - V3(i1.pos);
  rot.storeCoeffs(i3.rot);
  pos.store(i3.pos);
}

}  // namespace

extern "C" void wco_spline_fit_eval(const double* ts, const double* pts3, int64_t K, const double* query_t,
                                    int64_t nq, double* out3, uint8_t* valid, double* ctrl3) {
  Interpolator it;
  it.Init(ts, pts3, (int)K);
  if (ctrl3)
    for (int64_t i = 0; i < 3 * K; ++i) ctrl3[i] = it.Q[i];
  for (int64_t i = 0; i < nq; ++i) {
    double o[3] = {0, 0, 0};
    bool   ok   = it.Interp(query_t[i], o);
    out3[3 * i] = o[0], out3[3 * i + 1] = o[1], out3[3 * i + 2] = o[2];
    if (valid) valid[i] = ok ? 1 : 0;
  }
}

// spline_interpolation.h:9-20
extern "C" double wco_cubic_bspline_approx(double p_1, double p0, double p1, double p2, double s) {
  double s2 = s * s, s3 = s * s * s;
  return (p_1 * std::pow(1 - s, 3) + p0 * (3 * s3 - 6 * s2 + 4) + p1 * (-3 * s3 + 3 * s2 + 3 * s + 1) + p2 * s3) / 6;
}
// spline_interpolation.h:22-40
extern "C" double wco_cubic_spline_interpolate(double s_1, double p_1, double s0, double p0, double s1, double p1,
                                               double s2, double p2, double s) {
  double m0 = 0.5 * ((p0 - p_1) / (s0 - s_1) + (p1 - p0) / (s1 - s0));
  double m1 = 0.5 * ((p1 - p0) / (s1 - s0) + (p2 - p1) / (s2 - s1));
  double t = (s - s0) / (s1 - s0), t2 = t * t, t3 = t * t * t;
  return (2 * t3 - 3 * t2 + 1) * p0 + (t3 - 2 * t2 + t) * (s1 - s0) * m0 + (-2 * t3 + 3 * t2) * p1 + (t3 - t2) * (s1 - s0) * m1;
}

// UpdateImuPoses (lidar_odometry.cc:187-215) then UpdateSamplePoses (:172-179)
extern "C" int wco_apply_corrections(wc_sample_state* samples, int64_t K, wc_imu_state* imu, int64_t n_imu) {
  std::vector<double> ts(K), rc(3 * K), pc(3 * K);
  for (int64_t k = 0; k < K; ++k) {
    ts[k] = samples[k].timestamp;
    for (int c = 0; c < 3; ++c) rc[3 * k + c] = samples[k].data_cor[c], pc[3 * k + c] = samples[k].data_cor[3 + c];
  }
  Interpolator rot_i, pos_i;
  rot_i.Init(ts.data(), rc.data(), (int)K);
  pos_i.Init(ts.data(), pc.data(), (int)K);
  int64_t first = -1, last = -1;
  for (int64_t i = 0; i < n_imu; ++i) {
    double r[3] = {0, 0, 0}, p[3] = {0, 0, 0};
    bool   ok = rot_i.Interp(imu[i].timestamp, r);
    pos_i.Interp(imu[i].timestamp, p);
    if (ok) {
      Q4 q = Exp(V3(r)) * Q4::FromCoeffs(imu[i].rot);
      V3 t = V3(p) + V3(imu[i].pos);
      q.storeCoeffs(imu[i].rot), t.store(imu[i].pos);
      if (first == -1) first = i;
      last = i;
    }
  }
  if (first != -1) {
    if (first != 0 || last != n_imu - 2) return WC_EOUT_OF_SPAN;  // CHECK_EQ at :209-210
    const wc_sample_state& b = samples[K - 1];
    PredictPoseOfNewImuState(imu[n_imu - 3], imu[n_imu - 2], V3(b.data_cor + 9), V3(b.data_cor + 6), V3(b.grav), imu[n_imu - 1]);
  }
  for (int64_t k = 0; k < K; ++k) {
    Q4 q = Exp(V3(samples[k].data_cor)) * Q4::FromCoeffs(samples[k].rot);
    V3 t = V3(samples[k].data_cor + 3) + V3(samples[k].pos);
    q.storeCoeffs(samples[k].rot), t.store(samples[k].pos);
    for (int c = 0; c < 6; ++c) samples[k].data_cor[c] = 0;
  }
  return WC_OK;
}

// PredictImuStatesAndSampleStates steps 2-3, lidar_odometry.cc:403-453: forward-predict imu[2..n) from the two states in
// front of them (PredictPoseOfNewImuState, :112-123, CHECK_NEAR on the spacing :119), then append n_new sample states at
// t_last_sample + i * sample_dt with the lerp / slerp pose of the bracketing IMU states and the given biases / gravity.
extern "C" int wco_predict_states(wc_imu_state* imu, int64_t n_imu, const double* ba3, const double* bg3, const double* grav3,
                                  double t_last_sample, double sample_dt, int64_t n_new, wc_sample_state* samples_out) {
  if (n_imu < 2) return WC_EINVAL;
  const V3 ba(ba3), bg(bg3), grav(grav3);
  for (int64_t k = 2; k < n_imu; ++k) {
    const double d3 = imu[k].timestamp - imu[k - 1].timestamp, d2 = imu[k - 1].timestamp - imu[k - 2].timestamp;
    if (!(std::fabs(d3 - d2) <= 1e-6)) return WC_EINVAL_TIME_ORDER;  // CHECK_NEAR :119
    PredictPoseOfNewImuState(imu[k - 2], imu[k - 1], ba, bg, grav, imu[k]);
  }
  for (int64_t i = 1; i <= n_new; ++i) {
    const double     t = t_last_sample + (double)i * sample_dt;  // :431
    wc_sample_state& ss = samples_out[i - 1];
    std::memset(&ss, 0, sizeof(ss));
    ss.timestamp = t;
    for (int c = 0; c < 3; ++c) ss.data_cor[6 + c] = bg3[c], ss.data_cor[9 + c] = ba3[c], ss.grav[c] = grav3[c];
    int64_t lo = 0, hi = n_imu;  // std::lower_bound :439
    while (lo < hi) {
      const int64_t mid = (lo + hi) / 2;
      if (imu[mid].timestamp < t) lo = mid + 1; else hi = mid;
    }
    if (lo == 0 || lo == n_imu) return WC_EOUT_OF_SPAN;  // CHECK_NE :441-442
    const double f   = (t - imu[lo - 1].timestamp) / (imu[lo].timestamp - imu[lo - 1].timestamp);
    const Q4     rot = Slerp(Q4::FromCoeffs(imu[lo - 1].rot), f, Q4::FromCoeffs(imu[lo].rot));
    const V3     pos = (1 - f) * V3(imu[lo - 1].pos) + f * V3(imu[lo].pos);
    rot.storeCoeffs(ss.rot), pos.store(ss.pos);
  }
  return WC_OK;
}

extern "C" void wco_so3(int op, const double* in, double* out) {
  switch (op) {
    case 0: Exp(V3(in)).storeCoeffs(out); break;
    case 1: Log(Q4::FromCoeffs(in)).store(out); break;
    case 2: Jl(V3(in)).store(out); break;
    case 3: Jl_inv(V3(in)).store(out); break;
    case 4: Jr(V3(in)).store(out); break;
    case 5: Jr_inv(V3(in)).store(out); break;
    case 6: {
      M3 ev;
      SymEig3(M3::FromRowMajor(in), out, ev);
      ev.store(out + 3);
      break;
    }
  }
}
