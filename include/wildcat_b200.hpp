// C++ host mirror of the reference's src/odometry entry points over the C ABI of wildcat_b200.h.
//
// The reference is compiled C++ whose dependencies (Eigen, Ceres, FLANN, PCL, glog) are absent from this image, so this
// header restates its interface for the hot path with plain standard-library types: the same names, argument order,
// ownership (std::shared_ptr in std::deque / std::vector) and error behaviour (the reference's CHECK aborts become a
// thrown wildcat_b200::Error carrying the wc_status).  Every function is a thin shim: it flattens the shared_ptr
// containers into the layout-identical POD arrays of the C ABI, calls libwildcat_b200.so, and scatters the results
// back.  There is no CPU implementation behind it.
//
//   reference (file:line)                                                   here
//   hilti_ros::Point                       src/common/common.h:12-28        wildcat_b200::Point (= wc_point48)
//   ImuState / SampleState / Surfel        src/odometry/surfel.h:9-127      same names, std::array instead of Eigen
//   BuildSurfels                           surfel_extraction.h:145-147      BuildSurfels(cloud, surfels)
//   UndistortSweep                         lidar_odometry.cc:143-158        UndistortSweep(in, imu_states, out)
//   UpdateSurfelPoses                      lidar_odometry.cc:160-170        UpdateSurfelPoses(imu_states, surfels)
//   KnnSurfelMatcher::{BuildIndex,Match}   knn_surfel_matcher.h:17-19       KnnSurfelMatcher
//   Build*Residuals + ceres::Solve         lidar_odometry.cc:254-363,541-561  SolveWindow(...)
//   CubicBSplineInterpolator               spline_interpolation.h:44-51     CubicBSplineInterpolator
//   UpdateImuPoses + UpdateSamplePoses     lidar_odometry.cc:172-215        ApplyCorrections(samples, imu_states)
//   AddLidarScan's extrinsic + filter      lidar_odometry.cc:489-496        FilterPoints(msg, points_buff)
//   PredictImuStatesAndSampleStates 2-3    lidar_odometry.cc:403-453        PredictStates(imu_states, n_predicted, samples, dt, n_new)
#pragma once
#include <array>
#include <cstring>
#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "wildcat_b200.h"

namespace wildcat_b200 {

using Vector3d    = std::array<double, 3>;
using Quaterniond = std::array<double, 4>;  // Eigen coefficient order: x, y, z, w
using Matrix3d    = std::array<double, 9>;  // symmetric where the reference's is; row-major otherwise

struct Error : std::runtime_error {
  wc_status status;
  Error(wc_status s, const std::string& what) : std::runtime_error(what), status(s) {}
};

// One device context per process, like the reference's per-process LidarOdometry singleton.
class Context {
 public:
  explicit Context(int device = 0, const wc_params* params = nullptr) {
    wc_status s = wc_create(params, device, &ctx_);
    if (s != WC_OK) throw Error(s, std::string("wc_create: ") + wc_status_str(s) + " (no CPU fallback exists)");
  }
  ~Context() { wc_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  wc_ctx* get() const { return ctx_; }
  void    Check(wc_status s, const char* where) const {
    if (s != WC_OK) throw Error(s, std::string(where) + ": " + wc_status_str(s) + " — " + wc_last_error(ctx_));
  }
  static Context& Default() {
    static Context c(0);
    return c;
  }

 private:
  wc_ctx* ctx_ = nullptr;
};

// ---- value types (layout-identical to the PODs of the C ABI) -------------------------------------------------------
using Point = wc_point48;  // hilti_ros::Point

struct ImuState {  // surfel.h:25-33
  double      timestamp = 0;
  Vector3d    pos{};
  Quaterniond rot{{0, 0, 0, 1}};
  Vector3d    acc{}, gyr{};
};
static_assert(sizeof(ImuState) == sizeof(wc_imu_state) && std::is_standard_layout<ImuState>::value, "ImuState layout");

struct SampleState {  // surfel.h:9-23: data_cor = [rot_cor, pos_cor, bg, ba]
  using Ptr = std::shared_ptr<SampleState>;
  double      timestamp = 0;
  double      data_cor[12] = {0};
  Vector3d    grav{};
  Quaterniond rot{{0, 0, 0, 1}};
  Vector3d    pos{};
  double*     rot_cor() { return data_cor; }
  double*     pos_cor() { return data_cor + 3; }
  double*     bg() { return data_cor + 6; }
  double*     ba() { return data_cor + 9; }
};
static_assert(sizeof(SampleState) == sizeof(wc_sample_state) && std::is_standard_layout<SampleState>::value, "SampleState layout");

struct Surfel {  // surfel.h:35-121
  using Ptr = std::shared_ptr<Surfel>;
  double      timestamp = 0, resolution = 0, plane_std_deviation = 0;
  Quaterniond rot{{0, 0, 0, 1}};
  Vector3d    pos{}, center{};
  Matrix3d    covariance{};
  Vector3d    norm{};
  int32_t     is_in_body_frame = 0, pad_ = 0;
};
static_assert(sizeof(Surfel) == sizeof(wc_surfel) && std::is_standard_layout<Surfel>::value, "Surfel layout");
static_assert(std::is_trivially_copyable<Surfel>::value && std::is_trivially_copyable<SampleState>::value &&
                  std::is_trivially_copyable<ImuState>::value,
              "the value types are copied to / from the PODs of the C ABI bytewise");

struct SurfelCorrespondence {  // surfel.h:124-127; s1 is the earlier surfel
  Surfel::Ptr s1, s2;
};

namespace detail {
template <class T, class Pod, class Container>
std::vector<Pod> Flatten(const Container& in) {
  std::vector<Pod> out(in.size());
  size_t           i = 0;
  for (const auto& p : in) std::memcpy(&out[i++], &*p, sizeof(Pod));
  return out;
}
inline std::vector<wc_imu_state> Flatten(const std::deque<ImuState>& in) {
  std::vector<wc_imu_state> out(in.size());
  size_t                    i = 0;
  for (const auto& s : in) std::memcpy(&out[i++], &s, sizeof(wc_imu_state));
  return out;
}
}  // namespace detail

// ---- BuildSurfels, surfel_extraction.h:145-147 (the GlobalMap scratch argument has no counterpart) ------------------
inline void BuildSurfels(const std::vector<Point>& cloud, std::deque<Surfel::Ptr>& surfels, Context& ctx = Context::Default()) {
  wc_params prm;
  wc_default_params(&prm);
  std::vector<wc_surfel> out((size_t)prm.max_surfels);
  size_t                 n = 0;
  ctx.Check(wc_build_surfels(ctx.get(), cloud.data(), cloud.size(), out.data(), out.size(), &n, nullptr, nullptr), "BuildSurfels");
  for (size_t i = 0; i < n; ++i) {  // already sorted by timestamp (surfel_extraction.cc:334)
    auto s = std::make_shared<Surfel>();
    std::memcpy(static_cast<void*>(s.get()), &out[i], sizeof(wc_surfel));
    surfels.push_back(std::move(s));
  }
}

// ---- the per-point loop at the top of AddLidarScan, lidar_odometry.cc:489-496: extrinsic + range / blind-box filter -----
inline void FilterPoints(const std::vector<Point>& msg, std::deque<Point>& points_buff, const wc_sweep_filter* filter = nullptr,
                         Context& ctx = Context::Default()) {
  wc_sweep_filter f;
  if (filter) f = *filter; else wc_default_sweep_filter(&f);
  std::vector<Point> kept(msg.size());
  size_t             n = 0;
  ctx.Check(wc_filter_points(ctx.get(), &f, msg.data(), msg.size(), kept.data(), kept.size(), &n), "FilterPoints");
  points_buff.insert(points_buff.end(), kept.begin(), kept.begin() + (std::ptrdiff_t)n);
}

// ---- steps 2-3 of PredictImuStatesAndSampleStates, lidar_odometry.cc:403-453 -------------------------------------------
// imu_states: the last two states carry poses, the ones appended after them (timestamp / acc / gyr set) receive theirs;
// n_new sample states are appended at sample_states.back()->timestamp + i * sample_dt with its biases and gravity.
inline void PredictStates(std::deque<ImuState>& imu_states, size_t n_already_predicted, std::deque<SampleState::Ptr>& sample_states,
                          double sample_dt, size_t n_new, Context& ctx = Context::Default()) {
  if (n_already_predicted < 2 || n_already_predicted > imu_states.size() || sample_states.empty())
    throw Error(WC_EINVAL, "PredictStates: need two predicted IMU states and one sample state");
  std::vector<wc_imu_state> imu(imu_states.size() - n_already_predicted + 2);
  for (size_t i = 0; i < imu.size(); ++i) std::memcpy(&imu[i], &imu_states[n_already_predicted - 2 + i], sizeof(wc_imu_state));
  SampleState&                 last = *sample_states.back();
  std::vector<wc_sample_state> added(n_new ? n_new : 1);
  ctx.Check(wc_predict_states(ctx.get(), imu.data(), imu.size(), last.ba(), last.bg(), last.grav.data(), last.timestamp, sample_dt, n_new,
                              added.data()),
            "PredictStates");
  for (size_t i = 2; i < imu.size(); ++i) std::memcpy(static_cast<void*>(&imu_states[n_already_predicted - 2 + i]), &imu[i], sizeof(wc_imu_state));
  for (size_t i = 0; i < n_new; ++i) {
    auto s = std::make_shared<SampleState>();
    std::memcpy(static_cast<void*>(s.get()), &added[i], sizeof(wc_sample_state));
    sample_states.push_back(std::move(s));
  }
}

// ---- UndistortSweep, lidar_odometry.cc:143-158 ---------------------------------------------------------------------------
inline void UndistortSweep(const std::vector<Point>& sweep_in, const std::deque<ImuState>& imu_states, std::vector<Point>& sweep_out,
                           Context& ctx = Context::Default()) {
  const auto imu = detail::Flatten(imu_states);
  sweep_out.resize(sweep_in.size());
  ctx.Check(wc_undistort_sweep(ctx.get(), imu.data(), imu.size(), sweep_in.data(), sweep_in.size(), sweep_out.data()), "UndistortSweep");
}

// ---- UpdateSurfelPoses, lidar_odometry.cc:160-170 + Surfel::UpdatePose, surfel.h:48-58 ----------------------------------
inline void UpdateSurfelPoses(const std::deque<ImuState>& imu_states, std::deque<Surfel::Ptr>& surfels, Context& ctx = Context::Default()) {
  const auto imu = detail::Flatten(imu_states);
  auto       s   = detail::Flatten<Surfel, wc_surfel>(surfels);
  ctx.Check(wc_update_surfel_poses(ctx.get(), imu.data(), imu.size(), s.data(), s.size()), "UpdateSurfelPoses");
  for (size_t i = 0; i < s.size(); ++i) std::memcpy(static_cast<void*>(surfels[i].get()), &s[i], sizeof(wc_surfel));
}

// ---- KnnSurfelMatcher, knn_surfel_matcher.h:17-19 ----------------------------------------------------------------------
class KnnSurfelMatcher {
 public:
  explicit KnnSurfelMatcher(Context& ctx = Context::Default()) : ctx_(ctx) {}
  void BuildIndex(const std::deque<Surfel::Ptr>& surfels) { target_surfels_ = surfels; }  // knn_surfel_matcher.cc:3-16
  void Match(std::deque<Surfel::Ptr>& surfels, std::vector<SurfelCorrespondence>& surfel_corrs) {  // :18-49
    surfel_corrs.clear();
    if (target_surfels_.empty() || surfels.empty()) return;
    const bool self = target_surfels_.size() == surfels.size() && target_surfels_.front() == surfels.front() &&
                      target_surfels_.back() == surfels.back();
    const auto q = detail::Flatten<Surfel, wc_surfel>(surfels);
    const auto t = self ? std::vector<wc_surfel>() : detail::Flatten<Surfel, wc_surfel>(target_surfels_);
    std::vector<wc_corr_idx> out(q.size());
    std::vector<uint8_t>     first_is_target(q.size());
    size_t                   n = 0;
    ctx_.Check(wc_match(ctx_.get(), q.data(), q.size(), self ? q.data() : t.data(), self ? q.size() : t.size(), self ? 1 : 0, out.data(),
                        out.size(), &n, first_is_target.data(), nullptr),
               "KnnSurfelMatcher::Match");
    for (size_t i = 0; i < n; ++i) {
      const bool tgt_first = !self && first_is_target[i];
      const auto& a = self ? surfels : (tgt_first ? target_surfels_ : surfels);
      const auto& b = self ? surfels : (tgt_first ? surfels : target_surfels_);
      surfel_corrs.push_back({a[out[i].s1], b[out[i].s2]});
    }
  }

 private:
  Context&                ctx_;
  std::deque<Surfel::Ptr> target_surfels_;
};

// ---- problem assembly + ceres::Solve, lidar_odometry.cc:254-363,541-561 ------------------------------------------------
// sld_corrs pair sliding-window surfels; fix_corrs pair a fixed-window surfel (s1) with a sliding-window surfel (s2).
// The corrections are written in place into SampleState::data_cor, like Ceres writes the parameter blocks.
inline wc_solve_summary SolveWindow(const std::deque<Surfel::Ptr>& surfels_sld_win, const std::deque<Surfel::Ptr>& surfels_fix_win,
                                    const std::vector<SurfelCorrespondence>& sld_corrs, const std::vector<SurfelCorrespondence>& fix_corrs,
                                    const std::deque<ImuState>& imu_states, std::deque<SampleState::Ptr>& sample_states,
                                    const wc_solve_opts* options = nullptr, Context& ctx = Context::Default()) {
  const auto sld = detail::Flatten<Surfel, wc_surfel>(surfels_sld_win), fix = detail::Flatten<Surfel, wc_surfel>(surfels_fix_win);
  const auto imu = detail::Flatten(imu_states);
  auto       smp = detail::Flatten<SampleState, wc_sample_state>(sample_states);
  std::unordered_map<const Surfel*, int32_t> sld_index, fix_index;
  for (size_t i = 0; i < surfels_sld_win.size(); ++i) sld_index[surfels_sld_win[i].get()] = (int32_t)i;
  for (size_t i = 0; i < surfels_fix_win.size(); ++i) fix_index[surfels_fix_win[i].get()] = (int32_t)i;
  std::vector<wc_corr_idx> cs(sld_corrs.size()), cf(fix_corrs.size());
  for (size_t i = 0; i < sld_corrs.size(); ++i) cs[i] = {sld_index.at(sld_corrs[i].s1.get()), sld_index.at(sld_corrs[i].s2.get())};
  for (size_t i = 0; i < fix_corrs.size(); ++i) cf[i] = {fix_index.at(fix_corrs[i].s1.get()), sld_index.at(fix_corrs[i].s2.get())};
  wc_solve_summary summary;
  ctx.Check(wc_window_solve(ctx.get(), sld.data(), sld.size(), fix.data(), fix.size(), cs.data(), cs.size(), cf.data(), cf.size(), imu.data(),
                            imu.size(), smp.data(), smp.size(), options, &summary),
            "SolveWindow");
  for (size_t k = 0; k < smp.size(); ++k) std::memcpy(sample_states[k]->data_cor, smp[k].data_cor, sizeof(smp[k].data_cor));
  return summary;
}

// ---- CubicBSplineInterpolator, spline_interpolation.h:44-51 --------------------------------------------------------------
class CubicBSplineInterpolator {
 public:
  CubicBSplineInterpolator(const std::vector<double>& timestamps, const std::vector<Vector3d>& points, Context& ctx = Context::Default())
      : ctx_(ctx), ts_(timestamps), pts_(points) {
    if (ts_.size() != pts_.size() || ts_.size() < 2) throw Error(WC_EINVAL, "CubicBSplineInterpolator: need >= 2 samples");
  }
  std::shared_ptr<Vector3d> Interp(double timestamp) const {  // nullptr outside [t_0, t_{K-1}]  (:52-54)
    Vector3d out{};
    uint8_t  valid = 0;
    ctx_.Check(wc_spline_fit_eval(ctx_.get(), ts_.data(), pts_[0].data(), ts_.size(), &timestamp, 1, out.data(), &valid), "Interp");
    return valid ? std::make_shared<Vector3d>(out) : nullptr;
  }

 private:
  Context&              ctx_;
  std::vector<double>   ts_;
  std::vector<Vector3d> pts_;
};

// ---- UpdateImuPoses + UpdateSamplePoses, lidar_odometry.cc:172-215 ------------------------------------------------------
inline void ApplyCorrections(std::deque<SampleState::Ptr>& sample_states, std::deque<ImuState>& imu_states, Context& ctx = Context::Default()) {
  auto smp = detail::Flatten<SampleState, wc_sample_state>(sample_states);
  auto imu = detail::Flatten(imu_states);
  ctx.Check(wc_apply_corrections(ctx.get(), smp.data(), smp.size(), imu.data(), imu.size()), "ApplyCorrections");
  for (size_t k = 0; k < smp.size(); ++k) std::memcpy(static_cast<void*>(sample_states[k].get()), &smp[k], sizeof(wc_sample_state));
  for (size_t i = 0; i < imu.size(); ++i) std::memcpy(static_cast<void*>(&imu_states[i]), &imu[i], sizeof(wc_imu_state));
}

// ---- streaming ingestion (no reference counterpart: wildcat_slam_node.cc:83-99 hands the bag over message by message) ------
// Announce the NEXT sweep before the window pass of the current one: its host-to-device copy then runs beside that pass's
// solve stage (at_solve) or starts at once, and the next BuildSurfels / wc_points_upload of the same vector finds the points
// on the device.  The vector must stay untouched until then.
inline void PrefetchSweep(const std::vector<Point>& next_sweep, bool at_solve = true, Context& ctx = Context::Default()) {
  ctx.Check(wc_points_prefetch(ctx.get(), next_sweep.data(), next_sweep.size(), at_solve ? WC_PREFETCH_AT_SOLVE : WC_PREFETCH_NOW),
            "PrefetchSweep");
}

// ---- ImuResampler, src/sensor/imu_resampler.h:12-53 (host logic at the IMU rate; same interface) ---------------------------
struct ImuData {  // common.h:31-35
  double   timestamp;
  Vector3d linear_acceleration;
  Vector3d angular_velocity;
};

class ImuResampler {
 public:
  explicit ImuResampler(int freq) : period_(1.0 / freq) {}

  void AddImuData(const ImuData& d) {  // the bracket: the two most recent raw samples
    if (have_ == 2) raw_[0] = raw_[1], have_ = 1;
    raw_[have_++] = d;
  }

  // one sample of the fixed-rate sequence, or nullptr: the first raw sample itself, afterwards t_prev + 1 / freq whenever
  // that instant lies inside the bracket (ends included), linearly interpolated
  std::shared_ptr<ImuData> AdvanceGetResampledImuData() {
    if (have_ < 2) return nullptr;
    if (!started_) {
      started_ = true, t_prev_ = raw_[0].timestamp;
      return std::make_shared<ImuData>(raw_[0]);
    }
    const double t = t_prev_ + period_;
    if (t < raw_[0].timestamp || t > raw_[1].timestamp) return nullptr;
    const double f = (t - raw_[0].timestamp) / (raw_[1].timestamp - raw_[0].timestamp);
    auto         out = std::make_shared<ImuData>();
    out->timestamp   = t;
    for (int k = 0; k < 3; ++k) {
      out->linear_acceleration[k] = (1 - f) * raw_[0].linear_acceleration[k] + f * raw_[1].linear_acceleration[k];
      out->angular_velocity[k]    = (1 - f) * raw_[0].angular_velocity[k] + f * raw_[1].angular_velocity[k];
    }
    t_prev_ = t;
    return out;
  }

 private:
  ImuData raw_[2];
  int     have_    = 0;
  bool    started_ = false;
  double  period_, t_prev_ = 0.0;
};

}  // namespace wildcat_b200
