// ORACLE — TEST INFRASTRUCTURE ONLY (see wc_math.h / wc_oracle.h).
// Cost functors, problem assembly and the trust-region solve restated from
//   src/odometry/cost_functor.h, src/odometry/lidar_odometry.cc:254-363,551-561
// plus the un-vendored third-party pieces the reference calls (version unpinned; Ubuntu 20.04 => Ceres 1.14):
//   ceres::CauchyLoss / TrivialLoss / Corrector, SubsetParameterization, TrustRegionMinimizer with
//   LevenbergMarquardtStrategy and a normal-equation Cholesky (SPARSE_NORMAL_CHOLESKY forms J^T J + D^2
//   and factorises it; a dense factorisation is the same linear solve up to round-off).
// PARITY UNPINNED: no reference test or fixture exercises any of this.
#include <cfloat>
#include <cstdint>
#include <vector>

#include "wc_math.h"
#include "wc_oracle.h"

using namespace wco;

namespace {

// index of the first sample with timestamp > t (std::upper_bound, lidar_odometry.cc:258,264,303,330)
int64_t SampleUpperBound(const wc_sample_state* s, int64_t K, double t) {
  int64_t lo = 0, hi = K;
  while (lo < hi) {
    int64_t mid = (lo + hi) / 2;
    if (t < s[mid].timestamp) hi = mid; else lo = mid + 1;
  }
  return lo;
}

struct WorldSurfel {
  V3     v;    // rot * CenterInBody()
  V3     pos;  // surfel pos
  V3     cw;   // GetCenterInWorld()
  M3     covw;
  double t;
};
WorldSurfel World(const wc_surfel& s) {
  Q4          q = Q4::FromCoeffs(s.rot);
  M3          R = ToMatrix(q);
  WorldSurfel w;
  w.v    = q * V3(s.center);
  w.pos  = V3(s.pos);
  w.cw   = w.v + w.pos;
  w.covw = (R * M3::FromRowMajor(s.covariance)) * ToMatrix(q.conjugate());  // surfel.h:89-91
  w.t    = s.timestamp;
  return w;
}

// One lidar residual block after construction (cost_functor.h:17-26,102-114).
struct LidarFactor {
  bool   unary;
  V3     n;       // norm_
  double w;       // weight_
  V3     c1w;     // unary: s1->GetCenterInWorld()
  V3     v1, p1;  // binary: s1 rot*center, pos
  V3     v2, p2;
  double f1, f2;  // interpolation factors (fixed by timestamps)
  int    b1l, b1r, b2l, b2r;  // sample (block) indices
  int    mode;                // binary: 0,1,2
};

int MakeLidarFactor(const wc_params* prm, const wc_surfel& s1, const wc_surfel& s2, bool unary,
                    const wc_sample_state* samples, int64_t K, LidarFactor* f) {
  if (!(s1.timestamp < s2.timestamp)) return WC_EINVAL_TIME_ORDER;  // lidar_odometry.cc:256,301
  WorldSurfel w1 = World(s1), w2 = World(s2);
  double      evals[3];
  M3          evecs;
  SymEig3(w1.covw + w2.covw, evals, evecs);
  f->unary = unary;
  f->w     = 1 / std::sqrt(prm->weight_floor + evals[0]);  // pow(0.05/6,2) + lambda_min
  f->n     = evecs.col(0);
  f->c1w = w1.cw, f->v1 = w1.v, f->p1 = w1.pos, f->v2 = w2.v, f->p2 = w2.pos;
  int64_t sp2r = SampleUpperBound(samples, K, s2.timestamp);
  if (sp2r == 0 || sp2r == K) return WC_EOUT_OF_SPAN;  // CHECKs at :265-266,304-305
  f->b2l = (int)sp2r - 1, f->b2r = (int)sp2r;
  f->f2 = (s2.timestamp - samples[f->b2l].timestamp) / (samples[f->b2r].timestamp - samples[f->b2l].timestamp);
  f->b1l = f->b1r = -1, f->f1 = 0, f->mode = 0;
  if (!unary) {
    int64_t sp1r = SampleUpperBound(samples, K, s1.timestamp);
    if (sp1r == 0 || sp1r == K) return WC_EOUT_OF_SPAN;  // :259-260
    f->b1l = (int)sp1r - 1, f->b1r = (int)sp1r;
    f->f1 = (s1.timestamp - samples[f->b1l].timestamp) / (samples[f->b1r].timestamp - samples[f->b1l].timestamp);
    if (samples[f->b1r].timestamp < samples[f->b2l].timestamp) f->mode = 0;  // :271
    else if (f->b1r == f->b2l) f->mode = 1;                                   // :280
    else f->mode = 2;
  }
  return WC_OK;
}

// Evaluate (cost_functor.h:28-59, 116-179).  x: K*12.  Outputs the raw residual and, per distinct parameter
// block, the 1x6 non-zero part of the Jacobian (columns 6..11 are zero).  nblk/blk/J describe the blocks as
// Ceres sees them after DispatchPtr aliasing.
struct LidarEval {
  double r;
  int    nblk;
  int    blk[4];
  double J[4][6];
};
void EvalLidar(const LidarFactor& f, const double* x, int jacobian_mode, bool want_jac, LidarEval* e) {
  const double* x2l = x + 12 * f.b2l;
  const double* x2r = x + 12 * f.b2r;
  V3            r_s2 = (1 - f.f2) * V3(x2l) + f.f2 * V3(x2r);
  V3            t_s2 = (1 - f.f2) * V3(x2l + 3) + f.f2 * V3(x2r + 3);
  Q4            E2   = Exp(r_s2);
  V3            r_s1, t_s1;
  Q4            E1;
  if (f.unary) {
    e->r = f.w * dot(f.n, f.c1w - E2 * f.v2 - t_s2 - f.p2);
  } else {
    const double* x1l = x + 12 * f.b1l;
    const double* x1r = x + 12 * f.b1r;
    r_s1 = (1 - f.f1) * V3(x1l) + f.f1 * V3(x1r);
    t_s1 = (1 - f.f1) * V3(x1l + 3) + f.f1 * V3(x1r + 3);
    E1   = Exp(r_s1);
    e->r = f.w * dot(f.n, E1 * f.v1 + t_s1 + f.p1 - E2 * f.v2 - t_s2 - f.p2);
  }
  if (!want_jac) return;
  // jacobian_s2 (:42-45,162-165)
  double J2[6], J1[6];
  {
    V3 a  = vTm(vTm(vTm(f.w * f.n, ToMatrix(E2)), Hat(f.v2)), Jr(r_s2));
    V3 b  = -(f.w * f.n);
    J2[0] = a.x, J2[1] = a.y, J2[2] = a.z, J2[3] = b.x, J2[4] = b.y, J2[5] = b.z;
  }
  if (f.unary) {
    e->nblk = 2, e->blk[0] = f.b2l, e->blk[1] = f.b2r;
    for (int k = 0; k < 6; ++k) e->J[0][k] = J2[k] * (1 - f.f2), e->J[1][k] = J2[k] * f.f2;
    return;
  }
  {
    V3 a  = vTm(vTm(vTm(-(f.w * f.n), ToMatrix(E1)), Hat(f.v1)), Jr(r_s1));
    V3 b  = f.w * f.n;
    J1[0] = a.x, J1[1] = a.y, J1[2] = a.z, J1[3] = b.x, J1[4] = b.y, J1[5] = b.z;
  }
  // the four writes of :152-175 in order, through the DispatchPtr aliasing of :215-229
  int    slot_of[4];  // which Ceres block each of (sp1l, sp1r, sp2l, sp2r) maps to
  if (f.mode == 0) {
    e->nblk = 4, e->blk[0] = f.b1l, e->blk[1] = f.b1r, e->blk[2] = f.b2l, e->blk[3] = f.b2r;
    slot_of[0] = 0, slot_of[1] = 1, slot_of[2] = 2, slot_of[3] = 3;
  } else if (f.mode == 1) {
    e->nblk = 3, e->blk[0] = f.b1l, e->blk[1] = f.b1r, e->blk[2] = f.b2r;
    slot_of[0] = 0, slot_of[1] = 1, slot_of[2] = 1, slot_of[3] = 2;
  } else {
    e->nblk = 2, e->blk[0] = f.b1l, e->blk[1] = f.b1r;
    slot_of[0] = 0, slot_of[1] = 1, slot_of[2] = 0, slot_of[3] = 1;
  }
  const double  g[4]  = {1 - f.f1, f.f1, 1 - f.f2, f.f2};
  const double* Js[4] = {J1, J1, J2, J2};
  bool          written[4] = {false, false, false, false};
  for (int w = 0; w < 4; ++w) {
    int s = slot_of[w];
    for (int k = 0; k < 6; ++k) {
      double val = Js[w][k] * g[w];
      if (jacobian_mode == WC_JAC_EXACT && written[s]) e->J[s][k] += val;
      else e->J[s][k] = val;  // WC_JAC_REFERENCE_OVERWRITE: the later '=' wins (Q1)
    }
    written[s] = true;
  }
}

// ceres::CauchyLoss(a) + Corrector (rho'' < 0 always => plain sqrt(rho') scaling)
struct Cauchy {
  double b, c;
  explicit Cauchy(double a) : b(a * a), c(1 / (a * a)) {}
  void Evaluate(double s, double rho[3]) const {
    double sum = 1 + s * c, inv = 1 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(DBL_MIN, inv);
    rho[2] = -c * (inv * inv);
  }
};

// ---- IMU factor (cost_functor.h:264-472) -----------------------------------------------------------
struct ImuFactorDef {
  wc_imu_state i1, i2, i3;
  double       ts[3];  // sp1, sp2, sp3 timestamps (sp3 = DBL_MAX in mode 1)
  int          blk[3];
  int          mode;  // 0: three blocks, 1: two blocks
};

struct StateCorr {
  V3 r, t, bg, ba;
};
// ComputeStateCorr (:358-400).  Returns false if the CHECK would fire.
bool ComputeStateCorr(const ImuFactorDef& f, const double* const xs[3], double timestamp, StateCorr* c, int* left,
                      double* factor_out) {
  int l;
  if (f.mode == 0) {
    bool in12 = timestamp >= f.ts[0] && timestamp < f.ts[1];
    bool in23 = timestamp >= f.ts[1] && timestamp <= f.ts[2];
    if (!(in12 || in23)) return false;
    l = in12 ? 0 : 1;
  } else {
    if (!(timestamp >= f.ts[0] && timestamp <= f.ts[1])) return false;
    l = 0;
  }
  double        factor = (timestamp - f.ts[l]) / (f.ts[l + 1] - f.ts[l]);
  const double *a = xs[l], *b = xs[l + 1];
  c->r  = (1 - factor) * V3(a) + factor * V3(b);
  c->t  = (1 - factor) * V3(a + 3) + factor * V3(b + 3);
  c->bg = (1 - factor) * V3(a + 6) + factor * V3(b + 6);
  c->ba = (1 - factor) * V3(a + 9) + factor * V3(b + 9);
  *left = l, *factor_out = factor;
  return true;
}

// F (:446-448)
M3 ImuF(const Q4& L, const Q4& R, const V3& r) { return (Jr_inv(Log((L * Exp(r)) * R)) * ToMatrix(R.conjugate())) * Jr(r); }

void SetBlock(double* J, int ld, int r0, int c0, const M3& m) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) J[(r0 + i) * ld + c0 + j] = m.m[i][j];
}

// Evaluate.  res[12]; jac: 12 x (12*nblk) row-major (nblk = 3 or 2), zero-initialised here.
bool EvalImu(const wc_params* prm, const ImuFactorDef& f, const double* const xs[3], const V3& gravity, double* res,
             double* jac) {
  const double wg = prm->weight_gyr, wa = prm->weight_acc, wbg = prm->weight_bg, wba = prm->weight_ba;
  const double dt = 1 / prm->imu_rate;
  StateCorr    c1, c2, c3;
  int          l1, l2, l3;
  double       f1, f2, f3;
  if (!ComputeStateCorr(f, xs, f.i1.timestamp, &c1, &l1, &f1)) return false;
  if (!ComputeStateCorr(f, xs, f.i2.timestamp, &c2, &l2, &f2)) return false;
  if (!ComputeStateCorr(f, xs, f.i3.timestamp, &c3, &l3, &f3)) return false;
  Q4 R1 = Q4::FromCoeffs(f.i1.rot), R2 = Q4::FromCoeffs(f.i2.rot);
  Q4 E1R1    = Exp(c1.r) * R1;
  V3 gyr_est = Log((E1R1.conjugate() * Exp(c2.r)) * R2) / dt;
  V3 acc_est = ((c3.t + V3(f.i3.pos)) + (c1.t + V3(f.i1.pos)) - 2 * (c2.t + V3(f.i2.pos))) / (dt * dt);
  V3 rg  = wg * ((V3(f.i1.gyr) + V3(f.i2.gyr)) / 2 - gyr_est - c1.bg);
  V3 ra  = wa * (E1R1 * (V3(f.i1.acc) - c1.ba) - acc_est + gravity);
  V3 rbg = wbg * (c1.bg - c2.bg);
  V3 rba = wba * (c1.ba - c2.ba);
  rg.store(res), ra.store(res + 3), rbg.store(res + 6), rba.store(res + 9);
  if (!jac) return true;

  const int nblk = f.mode == 0 ? 3 : 2;
  const int ld   = 12 * nblk;
  for (int i = 0; i < 12 * ld; ++i) jac[i] = 0;
  double tau[144] = {0}, tau1[144] = {0}, tau2[144] = {0};
  M3     I        = M3::Identity();
  SetBlock(tau, 12, 0, 0, (wg * (1 / dt)) * ImuF(R1.conjugate(), Exp(c2.r) * R2, c1.r));
  SetBlock(tau, 12, 0, 6, (-wg) * I);
  SetBlock(tau, 12, 3, 0, (-wa) * ((ToMatrix(Exp(c1.r)) * Hat(R1 * (V3(f.i1.acc) - c1.ba))) * Jr(c1.r)));
  SetBlock(tau, 12, 3, 3, (-wa * (1 / dt / dt)) * I);
  SetBlock(tau, 12, 3, 9, (-wa) * ToMatrix(E1R1));
  SetBlock(tau, 12, 6, 6, wbg * I);
  SetBlock(tau, 12, 9, 9, wba * I);
  SetBlock(tau1, 12, 0, 0, (-wg * (1 / dt)) * ImuF(E1R1.conjugate(), R2, c2.r));
  SetBlock(tau1, 12, 0, 6, (-wg) * I);
  SetBlock(tau1, 12, 3, 3, (wa * (2 / dt / dt)) * I);
  SetBlock(tau1, 12, 6, 6, (-wbg) * I);
  SetBlock(tau1, 12, 9, 9, (-wba) * I);
  SetBlock(tau2, 12, 3, 3, (-wa * (1 / dt / dt)) * I);
  // DispatchJacobians (:402-444)
  const double* taus[3] = {tau, tau1, tau2};
  const int     ls[3]   = {l1, l2, l3};
  const double  fs[3]   = {f1, f2, f3};
  for (int s = 0; s < 3; ++s)
    for (int i = 0; i < 12; ++i)
      for (int j = 0; j < 12; ++j) {
        double v = taus[s][12 * i + j];
        if (v == 0) continue;
        jac[i * ld + 12 * ls[s] + j] += v * (1 - fs[s]);
        jac[i * ld + 12 * (ls[s] + 1) + j] += v * fs[s];
      }
  return true;
}

// BuildImuResiduals (lidar_odometry.cc:319-363)
void BuildImuFactors(const wc_imu_state* imu, int64_t n_imu, const wc_sample_state* samples, int64_t K,
                     std::vector<ImuFactorDef>* out) {
  for (int64_t i = 0; i + 2 < n_imu; ++i) {
    const wc_imu_state &i1 = imu[i], &i2 = imu[i + 1], &i3 = imu[i + 2];
    if (i1.timestamp < samples[0].timestamp) continue;
    if (i3.timestamp > samples[K - 1].timestamp) break;
    int64_t      sp2 = SampleUpperBound(samples, K, i1.timestamp);
    ImuFactorDef f;
    f.i1 = i1, f.i2 = i2, f.i3 = i3;
    f.blk[0] = (int)sp2 - 1, f.blk[1] = (int)sp2;
    f.ts[0] = samples[sp2 - 1].timestamp, f.ts[1] = samples[sp2].timestamp;
    if (sp2 == K - 1) {
      f.mode = 1, f.blk[2] = -1, f.ts[2] = DBL_MAX;
    } else {
      f.mode = 0, f.blk[2] = (int)sp2 + 1, f.ts[2] = samples[sp2 + 1].timestamp;
    }
    out->push_back(f);
  }
}

// ---- the assembled problem -------------------------------------------------------------------------
struct Problem {
  const wc_params*          prm;
  wc_solve_opts             opts;
  int64_t                   K;
  std::vector<LidarFactor>  lidar;
  int64_t                   n_sld = 0, n_fix = 0;
  std::vector<ImuFactorDef> imu;
  V3                        gravity;
  std::vector<int>          col_of;  // ambient index -> reduced column, -1 if held constant
  int                       D = 0;   // reduced size

  // cost = 1/2 sum rho; optionally g (ambient 12K) and H (ambient, full symmetric, row-major)
  int Evaluate(const double* x, double* cost, double* g, double* H) const {
    const int64_t N = 12 * K;
    if (g) std::fill(g, g + N, 0.0);
    if (H) std::fill(H, H + N * N, 0.0);
    Cauchy loss(prm->cauchy_a);
    double c = 0;
    for (const LidarFactor& f : lidar) {
      LidarEval e;
      EvalLidar(f, x, opts.jacobian_mode, g || H, &e);
      double rho[3];
      loss.Evaluate(e.r * e.r, rho);
      c += 0.5 * rho[0];
      if (!(g || H)) continue;
      double sr = std::sqrt(rho[1]);
      double r  = e.r * sr;
      for (int a = 0; a < e.nblk; ++a)
        for (int k = 0; k < 6; ++k) e.J[a][k] *= sr;
      for (int a = 0; a < e.nblk; ++a) {
        if (g)
          for (int k = 0; k < 6; ++k) g[12 * e.blk[a] + k] += e.J[a][k] * r;
        if (H)
          for (int b = 0; b < e.nblk; ++b)
            for (int i = 0; i < 6; ++i)
              for (int j = 0; j < 6; ++j) H[(12 * e.blk[a] + i) * N + 12 * e.blk[b] + j] += e.J[a][i] * e.J[b][j];
      }
    }
    double res[12], jac[12 * 36];
    for (const ImuFactorDef& f : imu) {
      const double* xs[3] = {x + 12 * f.blk[0], x + 12 * f.blk[1], f.mode == 0 ? x + 12 * f.blk[2] : nullptr};
      if (!EvalImu(prm, f, xs, gravity, res, (g || H) ? jac : nullptr)) return WC_EOUT_OF_SPAN;
      for (int i = 0; i < 12; ++i) c += 0.5 * res[i] * res[i];  // TrivialLoss
      if (!(g || H)) continue;
      const int nblk = f.mode == 0 ? 3 : 2, ld = 12 * nblk;
      for (int a = 0; a < nblk; ++a)
        for (int i = 0; i < 12; ++i) {
          const int gi = 12 * f.blk[a] + i;
          if (g) {
            double s = 0;
            for (int r = 0; r < 12; ++r) s += jac[r * ld + 12 * a + i] * res[r];
            g[gi] += s;
          }
          if (H)
            for (int b = 0; b < nblk; ++b)
              for (int j = 0; j < 12; ++j) {
                double s = 0;
                for (int r = 0; r < 12; ++r) s += jac[r * ld + 12 * a + i] * jac[r * ld + 12 * b + j];
                H[gi * N + 12 * f.blk[b] + j] += s;
              }
        }
    }
    *cost = c;
    return WC_OK;
  }
};

int Assemble(const wc_params* prm, const wc_solve_opts* opts, const wc_surfel* sld, int64_t n_sld,
             const wc_surfel* fix, int64_t n_fix, const wc_corr_idx* sld_corr, int64_t n_sld_corr,
             const wc_corr_idx* fix_corr, int64_t n_fix_corr, const wc_imu_state* imu, int64_t n_imu,
             const wc_sample_state* samples, int64_t K, Problem* p) {
  if (K < 2) return WC_EINVAL;
  p->prm = prm, p->opts = *opts, p->K = K;
  p->lidar.reserve(n_sld_corr + n_fix_corr);
  for (int64_t i = 0; i < n_sld_corr; ++i) {
    if (sld_corr[i].s1 < 0 || sld_corr[i].s1 >= n_sld || sld_corr[i].s2 < 0 || sld_corr[i].s2 >= n_sld) return WC_EINVAL;
    LidarFactor f;
    int         st = MakeLidarFactor(prm, sld[sld_corr[i].s1], sld[sld_corr[i].s2], false, samples, K, &f);
    if (st) return st;
    p->lidar.push_back(f);
  }
  for (int64_t i = 0; i < n_fix_corr; ++i) {
    if (fix_corr[i].s1 < 0 || fix_corr[i].s1 >= n_fix || fix_corr[i].s2 < 0 || fix_corr[i].s2 >= n_sld) return WC_EINVAL;
    LidarFactor f;
    int         st = MakeLidarFactor(prm, fix[fix_corr[i].s1], sld[fix_corr[i].s2], true, samples, K, &f);
    if (st) return st;
    p->lidar.push_back(f);
  }
  p->n_sld = n_sld_corr, p->n_fix = n_fix_corr;
  if (opts->use_imu_factors && imu && n_imu >= 3) BuildImuFactors(imu, n_imu, samples, K, &p->imu);
  p->gravity = V3(samples[K - 1].grav);
  p->col_of.assign(12 * K, 0);
  int d = 0;
  for (int64_t i = 0; i < 12 * K; ++i) {
    bool fixed   = opts->fix_first_position && i >= 3 && i <= 5;  // SubsetParameterization(12,{3,4,5}) on sample 0
    p->col_of[i] = fixed ? -1 : d++;
  }
  p->D = d;
  return WC_OK;
}

// in-place dense Cholesky (lower) of an n x n row-major SPD matrix; false if not positive definite
bool Cholesky(std::vector<double>& A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0) || !std::isfinite(d)) return false;
    d                     = std::sqrt(d);
    A[(size_t)j * n + j]  = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  return true;
}
void CholSolve(const std::vector<double>& L, int n, std::vector<double>& b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}

}  // namespace

extern "C" int wco_window_evaluate(const wc_params* prm, const wc_solve_opts* opts, const wc_surfel* sld,
                                   int64_t n_sld, const wc_surfel* fix, int64_t n_fix, const wc_corr_idx* sld_corr,
                                   int64_t n_sld_corr, const wc_corr_idx* fix_corr, int64_t n_fix_corr,
                                   const wc_imu_state* imu, int64_t n_imu, const wc_sample_state* samples, int64_t K,
                                   double* cost, double* grad, double* jtj) {
  Problem p;
  int     st = Assemble(prm, opts, sld, n_sld, fix, n_fix, sld_corr, n_sld_corr, fix_corr, n_fix_corr, imu, n_imu,
                        samples, K, &p);
  if (st) return st;
  std::vector<double> x(12 * K);
  for (int64_t k = 0; k < K; ++k)
    for (int j = 0; j < 12; ++j) x[12 * k + j] = samples[k].data_cor[j];
  double c;
  st = p.Evaluate(x.data(), &c, grad, jtj);
  if (cost) *cost = c;
  return st;
}

// ceres::Solve with the options of lidar_odometry.cc:551-554 — TrustRegionMinimizer::Minimize restated.
extern "C" int wco_window_solve(const wc_params* prm, const wc_solve_opts* opts, const wc_surfel* sld, int64_t n_sld,
                                const wc_surfel* fix, int64_t n_fix, const wc_corr_idx* sld_corr, int64_t n_sld_corr,
                                const wc_corr_idx* fix_corr, int64_t n_fix_corr, const wc_imu_state* imu,
                                int64_t n_imu, wc_sample_state* samples, int64_t K, wc_solve_summary* sum) {
  Problem p;
  int     st = Assemble(prm, opts, sld, n_sld, fix, n_fix, sld_corr, n_sld_corr, fix_corr, n_fix_corr, imu, n_imu,
                        samples, K, &p);
  if (st) return st;
  const int64_t       N = 12 * K;
  const int           D = p.D;
  std::vector<double> x(N), cand(N), g(N), H((size_t)N * N);
  for (int64_t k = 0; k < K; ++k)
    for (int j = 0; j < 12; ++j) x[12 * k + j] = samples[k].data_cor[j];

  std::memset(sum, 0, sizeof(*sum));
  sum->num_residual_blocks_sld = (int32_t)p.n_sld, sum->num_residual_blocks_fix = (int32_t)p.n_fix;
  sum->num_residual_blocks_imu = (int32_t)p.imu.size();

  std::vector<double> scale(D, 1.0), gs(D), Hs((size_t)D * D), A((size_t)D * D), step(D), diag(D), delta(N);
  bool                have_scale = false;
  double              x_cost = 0, x_norm = 0, grad_max = 0;

  // EvaluateGradientAndJacobian: robustified g, H at x; Jacobi scaling fixed at the first call
  auto Linearize = [&]() -> int {
    int s = p.Evaluate(x.data(), &x_cost, g.data(), H.data());
    if (s) return s;
    ++sum->num_linearizations;
    grad_max = 0;
    for (int64_t i = 0; i < N; ++i)
      if (p.col_of[i] >= 0) grad_max = std::max(grad_max, std::fabs(g[i]));  // unscaled gradient, local space
    if (!have_scale) {
      for (int64_t i = 0; i < N; ++i)
        if (p.col_of[i] >= 0) scale[p.col_of[i]] = 1.0 / (1.0 + std::sqrt(H[(size_t)i * N + i]));
      have_scale = true;
    }
    for (int64_t i = 0; i < N; ++i) {
      int ci = p.col_of[i];
      if (ci < 0) continue;
      gs[ci] = g[i] * scale[ci];
      for (int64_t j = 0; j < N; ++j) {
        int cj = p.col_of[j];
        if (cj >= 0) Hs[(size_t)ci * D + cj] = H[(size_t)i * N + j] * scale[ci] * scale[cj];
      }
    }
    return WC_OK;
  };
  auto Norm = [&](const std::vector<double>& v) {
    double s = 0;
    for (double e : v) s += e * e;
    return std::sqrt(s);
  };

  st = Linearize();
  if (st) return st;
  sum->initial_cost = x_cost;
  x_norm            = Norm(x);
  double radius = opts->initial_trust_region_radius, decrease_factor = 2.0;
  bool   reuse_diagonal = false;
  int    num_consecutive_invalid = 0;
  int    iteration = 0;
  bool   last_successful = true;  // iteration 0 counts as successful (Init())
  sum->termination       = WC_TERM_NO_CONVERGENCE;

  auto StepRejected = [&]() { radius = radius / decrease_factor, decrease_factor *= 2.0, reuse_diagonal = true; };

  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iteration >= opts->max_num_iterations) break;
    if (last_successful && grad_max <= opts->gradient_tolerance) { sum->termination = WC_TERM_GRADIENT_TOL; break; }
    if (radius < opts->min_trust_region_radius) { sum->termination = WC_TERM_MIN_RADIUS; break; }
    ++iteration;
    last_successful = false;
    const int it    = iteration < WC_MAX_ITER_LOG ? iteration : WC_MAX_ITER_LOG - 1;

    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal)
      for (int i = 0; i < D; ++i) diag[i] = std::min(std::max(Hs[(size_t)i * D + i], opts->min_lm_diagonal), opts->max_lm_diagonal);
    A = Hs;
    for (int i = 0; i < D; ++i) {
      double lm = std::sqrt(diag[i] / radius);
      A[(size_t)i * D + i] += lm * lm;
    }
    bool valid = Cholesky(A, D);
    if (valid) {
      step = gs;
      CholSolve(A, D, step);
      for (int i = 0; i < D; ++i) {
        step[i] = -step[i];
        if (!std::isfinite(step[i])) valid = false;
      }
    }
    reuse_diagonal = true;
    double model_cost_change = 0;
    if (valid) {
      // -(J d)^T (r + J d / 2) = -d^T (g + H d / 2)
      for (int i = 0; i < D; ++i) {
        double hd = 0;
        for (int j = 0; j < D; ++j) hd += Hs[(size_t)i * D + j] * step[j];
        model_cost_change -= step[i] * (gs[i] + 0.5 * hd);
      }
      valid = model_cost_change > 0.0;
    }
    sum->iter_radius[it] = radius;
    if (!valid) {
      // HandleInvalidStep
      ++sum->num_unsuccessful_steps;
      sum->iter_cost[it] = NAN, sum->iter_accepted[it] = 0;
      if (++num_consecutive_invalid >= 5) { sum->termination = WC_TERM_FAILURE; break; }
      StepRejected();
      continue;
    }
    num_consecutive_invalid = 0;
    for (int64_t i = 0; i < N; ++i) delta[i] = p.col_of[i] >= 0 ? step[p.col_of[i]] * scale[p.col_of[i]] : 0.0;
    // ComputeCandidatePointAndEvaluateCost
    for (int64_t i = 0; i < N; ++i) cand[i] = x[i] + delta[i];
    double cand_cost;
    st = p.Evaluate(cand.data(), &cand_cost, nullptr, nullptr);
    if (st) return st;
    if (!std::isfinite(cand_cost)) cand_cost = DBL_MAX;
    sum->iter_cost[it] = cand_cost;
    // ParameterToleranceReached
    if (Norm(delta) <= opts->parameter_tolerance * (x_norm + opts->parameter_tolerance)) {
      sum->termination = WC_TERM_PARAMETER_TOL;
      break;
    }
    // FunctionToleranceReached
    if (std::fabs(x_cost - cand_cost) <= opts->function_tolerance * x_cost) {
      sum->termination = WC_TERM_FUNCTION_TOL;
      break;
    }
    // IsStepSuccessful
    double relative_decrease = (x_cost - cand_cost) / model_cost_change;
    if (relative_decrease > opts->min_relative_decrease) {
      // HandleSuccessfulStep
      x      = cand;
      x_norm = Norm(x);
      st     = Linearize();
      if (st) return st;
      last_successful         = true;
      sum->iter_accepted[it]  = 1;
      ++sum->num_successful_steps;
      double d = 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3);
      radius   = radius / std::max(1.0 / 3.0, d);
      radius   = std::min(opts->max_trust_region_radius, radius);
      decrease_factor = 2.0, reuse_diagonal = false;
    } else {
      sum->iter_accepted[it] = 0;
      ++sum->num_unsuccessful_steps;
      StepRejected();
    }
  }
  sum->num_iterations = iteration;
  sum->final_cost     = x_cost;
  for (int64_t k = 0; k < K; ++k)
    for (int j = 0; j < 12; ++j) samples[k].data_cor[j] = x[12 * k + j];
  return WC_OK;
}

extern "C" int wco_lidar_factor(const wc_params* prm, int jacobian_mode, const wc_surfel* s1, const wc_surfel* s2,
                                int unary, const double* sample_ts, const int32_t* blk, const double* x, int64_t K,
                                double* residual, double* jac, double* weight, double* normal3) {
  // build a throw-away sample array carrying only timestamps
  std::vector<wc_sample_state> samples(K);
  std::memset(samples.data(), 0, sizeof(wc_sample_state) * K);
  (void)blk;
  for (int64_t k = 0; k < K; ++k) samples[k].timestamp = sample_ts[k];
  LidarFactor f;
  int         st = MakeLidarFactor(prm, *s1, *s2, unary != 0, samples.data(), K, &f);
  if (st) return -st;
  LidarEval e;
  EvalLidar(f, x, jacobian_mode, jac != nullptr, &e);
  *residual = e.r;
  if (weight) *weight = f.w;
  if (normal3) f.n.store(normal3);
  if (jac) {
    for (int64_t i = 0; i < 12 * K; ++i) jac[i] = 0;
    for (int a = 0; a < e.nblk; ++a)
      for (int k = 0; k < 6; ++k) jac[12 * e.blk[a] + k] = e.J[a][k];
  }
  return 1;
}

extern "C" int wco_imu_factor(const wc_params* prm, const wc_imu_state* i3, const double* sample_ts, int mode,
                              const double* grav3, const double* x, double* residual12, double* jac) {
  ImuFactorDef f;
  f.i1 = i3[0], f.i2 = i3[1], f.i3 = i3[2];
  f.ts[0] = sample_ts[0], f.ts[1] = sample_ts[1], f.ts[2] = mode == 0 ? sample_ts[2] : DBL_MAX;
  f.blk[0] = 0, f.blk[1] = 1, f.blk[2] = mode == 0 ? 2 : -1;
  f.mode               = mode;
  const double* xs[3]  = {x, x + 12, mode == 0 ? x + 24 : nullptr};
  if (!EvalImu(prm, f, xs, V3(grav3), residual12, jac)) return -WC_EOUT_OF_SPAN;
  return 12;
}

// ---- observability outputs --------------------------------------------------------------------------
// PrintSurfelResiduals / PrintImuResiduals (lidar_odometry.cc:56-93): ceres::Problem::Evaluate with apply_loss_function =
// true returns the residuals after the loss corrector, r * sqrt(rho') for the Cauchy blocks, unchanged for the IMU
// blocks (TrivialLoss).  lidar_res: sliding-window blocks first, then the fixed-window blocks, in construction order.
extern "C" int wco_window_residuals(const wc_params* prm, const wc_solve_opts* opts, const wc_surfel* sld, int64_t n_sld,
                                    const wc_surfel* fix, int64_t n_fix, const wc_corr_idx* sld_corr, int64_t n_sld_corr,
                                    const wc_corr_idx* fix_corr, int64_t n_fix_corr, const wc_imu_state* imu, int64_t n_imu,
                                    const wc_sample_state* samples, int64_t K, double* lidar_res, double* imu_res,
                                    int64_t* n_imu_blocks) {
  Problem p;
  int     st = Assemble(prm, opts, sld, n_sld, fix, n_fix, sld_corr, n_sld_corr, fix_corr, n_fix_corr, imu, n_imu, samples, K, &p);
  if (st) return st;
  std::vector<double> x(12 * K);
  for (int64_t k = 0; k < K; ++k)
    for (int j = 0; j < 12; ++j) x[12 * k + j] = samples[k].data_cor[j];
  Cauchy loss(prm->cauchy_a);
  size_t i = 0;
  for (const LidarFactor& f : p.lidar) {
    LidarEval e;
    EvalLidar(f, x.data(), p.opts.jacobian_mode, false, &e);
    double rho[3];
    loss.Evaluate(e.r * e.r, rho);
    lidar_res[i++] = e.r * std::sqrt(rho[1]);
  }
  *n_imu_blocks = (int64_t)p.imu.size();
  size_t b = 0;
  for (const ImuFactorDef& f : p.imu) {
    const double* xs[3] = {x.data() + 12 * f.blk[0], x.data() + 12 * f.blk[1], f.mode == 0 ? x.data() + 12 * f.blk[2] : nullptr};
    if (!EvalImu(prm, f, xs, p.gravity, imu_res + 12 * b, nullptr)) return WC_EOUT_OF_SPAN;
    ++b;
  }
  return 0;
}

// PubSurfels (surfel_extraction.cc:360-417) without the ROS message.  The eigenvector signs follow this file's SymEig3
// restatement of Eigen's solver (parity unpinned: Eigen is absent from this image).
extern "C" void wco_surfel_markers(const wc_surfel* s, int64_t n, wc_marker* out) {
  for (int64_t i = 0; i < n; ++i) {
    WorldSurfel w = World(s[i]);
    Q4          q = Q4::FromCoeffs(s[i].rot);
    double      ev[3];
    M3          V;
    SymEig3(w.covw, ev, V);
    auto unit = [](V3 v) { return v / std::sqrt(dot(v, v)); };
    V3   c0 = unit(V.col(0)), c1 = unit(V.col(1)), c2 = unit(V.col(2));
    if (dot(cross(c0, c1), c2) < 0) {  // makeRightHanded :340-358
      std::swap(c0, c1);
      std::swap(ev[0], ev[1]);
    }
    const double m[3][3] = {{c0.x, c1.x, c2.x}, {c0.y, c1.y, c2.y}, {c0.z, c1.z, c2.z}};
    double       qq[4];
    double       t = m[0][0] + m[1][1] + m[2][2];
    if (t > 0.0) {  // Eigen::Quaterniond(Matrix3d)
      t     = std::sqrt(t + 1.0);
      qq[3] = 0.5 * t;
      t     = 0.5 / t;
      qq[0] = (m[2][1] - m[1][2]) * t, qq[1] = (m[0][2] - m[2][0]) * t, qq[2] = (m[1][0] - m[0][1]) * t;
    } else {
      int a = 0;
      if (m[1][1] > m[0][0]) a = 1;
      if (m[2][2] > m[a][a]) a = 2;
      const int b = (a + 1) % 3, c = (b + 1) % 3;
      t     = std::sqrt(m[a][a] - m[b][b] - m[c][c] + 1.0);
      qq[a] = 0.5 * t;
      t     = 0.5 / t;
      qq[3] = (m[c][b] - m[b][c]) * t;
      qq[b] = (m[b][a] + m[a][b]) * t;
      qq[c] = (m[c][a] + m[a][c]) * t;
    }
    const bool body = s[i].is_in_body_frame != 0;
    V3         cen  = body ? w.cw : V3(s[i].center);
    V3         nw   = body ? q * V3(s[i].norm) : V3(s[i].norm);
    wc_marker& k    = out[i];
    k.position[0] = cen.x, k.position[1] = cen.y, k.position[2] = cen.z;
    for (int j = 0; j < 4; ++j) k.orientation[j] = qq[j];
    for (int j = 0; j < 3; ++j) k.scale[j] = 3.0 * std::sqrt(ev[j]);
    k.color[0] = (float)((nw.x + 1) / 2), k.color[1] = (float)((nw.y + 1) / 2), k.color[2] = (float)((nw.z + 1) / 2), k.color[3] = 1.f;
  }
}
