// Exercises the C++ host mirror (include/wildcat_b200.hpp) the way LidarOdometry::AddLidarScan uses the reference's own
// entry points (src/odometry/lidar_odometry.cc:523-566): BuildSurfels -> UpdateSurfelPoses -> KnnSurfelMatcher x2 ->
// SolveWindow -> ApplyCorrections, plus CubicBSplineInterpolator and an error path.  Inputs are the raw arrays of a
// synthetic window written by the pytest driver; outputs are raw arrays the driver compares with the Python mirror.
//   mirror_test <dir>      (reads <dir>/{points,imu,samples,fix_points,fix_imu}.bin)
#include <cstdio>
#include <fstream>
#include <iostream>
#include <unordered_map>

#include "wildcat_b200.hpp"

namespace wb = wildcat_b200;

template <class T>
static std::vector<T> ReadAll(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("cannot open " + path);
  const size_t bytes = (size_t)f.tellg();
  std::vector<T> v(bytes / sizeof(T));
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
  return v;
}
template <class T>
static void WriteAll(const std::string& path, const std::vector<T>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}
static std::deque<wb::ImuState> ToDeque(const std::vector<wb::ImuState>& v) { return {v.begin(), v.end()}; }

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  try {
    const auto points     = ReadAll<wb::Point>(dir + "/points.bin");
    const auto fix_points = ReadAll<wb::Point>(dir + "/fix_points.bin");
    auto       imu_states = ToDeque(ReadAll<wb::ImuState>(dir + "/imu.bin"));
    const auto fix_imu    = ToDeque(ReadAll<wb::ImuState>(dir + "/fix_imu.bin"));
    std::deque<wb::SampleState::Ptr> sample_states;
    for (const auto& s : ReadAll<wb::SampleState>(dir + "/samples.bin")) sample_states.push_back(std::make_shared<wb::SampleState>(s));

    // lidar_odometry.cc:523-527
    std::deque<wb::Surfel::Ptr> surfels_sld_win, surfels_fix_win;
    wb::BuildSurfels(points, surfels_sld_win);
    wb::UpdateSurfelPoses(imu_states, surfels_sld_win);
    wb::BuildSurfels(fix_points, surfels_fix_win);
    wb::UpdateSurfelPoses(fix_imu, surfels_fix_win);
    // :532-538
    std::vector<wb::SurfelCorrespondence> surfel_corrs_sld, surfel_corrs_fix;
    wb::KnnSurfelMatcher matcher_sld, matcher_fix;
    matcher_sld.BuildIndex(surfels_sld_win);
    matcher_sld.Match(surfels_sld_win, surfel_corrs_sld);
    matcher_fix.BuildIndex(surfels_fix_win);
    matcher_fix.Match(surfels_sld_win, surfel_corrs_fix);
    // :541-561
    const wc_solve_summary summary = wb::SolveWindow(surfels_sld_win, surfels_fix_win, surfel_corrs_sld, surfel_corrs_fix, imu_states, sample_states);

    std::vector<wb::Surfel> sld_out;
    for (const auto& s : surfels_sld_win) sld_out.push_back(*s);
    WriteAll(dir + "/cpp_sld.bin", sld_out);
    std::unordered_map<const wb::Surfel*, int32_t> si, fi;
    for (size_t i = 0; i < surfels_sld_win.size(); ++i) si[surfels_sld_win[i].get()] = (int32_t)i;
    for (size_t i = 0; i < surfels_fix_win.size(); ++i) fi[surfels_fix_win[i].get()] = (int32_t)i;
    std::vector<int32_t> cs, cf;
    for (const auto& c : surfel_corrs_sld) cs.push_back(si.at(c.s1.get())), cs.push_back(si.at(c.s2.get()));
    for (const auto& c : surfel_corrs_fix) cf.push_back(fi.at(c.s1.get())), cf.push_back(si.at(c.s2.get()));
    WriteAll(dir + "/cpp_corr_sld.bin", cs);
    WriteAll(dir + "/cpp_corr_fix.bin", cf);
    std::vector<double> cor;
    for (const auto& s : sample_states) cor.insert(cor.end(), s->data_cor, s->data_cor + 12);
    WriteAll(dir + "/cpp_data_cor.bin", cor);

    // :564-566 (UpdateImuPoses + UpdateSamplePoses): corrections folded into the poses and zeroed
    wb::ApplyCorrections(sample_states, imu_states);
    double residual_cor = 0;
    for (const auto& s : sample_states)
      for (int k = 0; k < 6; ++k) residual_cor += std::abs(s->data_cor[k]);

    // CubicBSplineInterpolator: Interp(t_i) reproduces sample i (spline_interpolation_test.cc:79-96), nullptr outside
    std::vector<double>       ts;
    std::vector<wb::Vector3d> pts;
    for (int i = 0; i < 8; ++i) ts.push_back(0.3 + 0.1 * i), pts.push_back({1.0 + i, 2.0 * i, 0.5 * i * i});
    wb::CubicBSplineInterpolator interp(ts, pts);
    double spline_err = 0;
    for (int i = 0; i < 8; ++i) {
      auto p = interp.Interp(ts[i]);
      if (!p) return 3;
      for (int k = 0; k < 3; ++k) spline_err = std::max(spline_err, std::abs((*p)[k] - pts[i][k]));
    }
    const bool outside_null = interp.Interp(0.2) == nullptr && interp.Interp(1.1) == nullptr;

    // the rows around the path: FilterPoints (lidar_odometry.cc:489-496), UndistortSweep (:143-158), PredictStates (:403-453)
    std::deque<wb::Point> points_buff;
    wb::FilterPoints(points, points_buff);
    std::vector<wb::Point> undistorted;
    wb::UndistortSweep(points, ToDeque(ReadAll<wb::ImuState>(dir + "/imu.bin")), undistorted);
    double und_sum = 0;
    for (const auto& p : undistorted) und_sum += (double)p.x + (double)p.y + (double)p.z;
    auto imu_pred = ToDeque(ReadAll<wb::ImuState>(dir + "/imu.bin"));  // poses of [2..) wiped, then re-predicted with zero biases
    const auto imu_ref = imu_pred;
    for (size_t i = 2; i < imu_pred.size(); ++i) imu_pred[i].pos = {0, 0, 0}, imu_pred[i].rot = {0, 0, 0, 0};
    std::deque<wb::SampleState::Ptr> smp_pred;
    smp_pred.push_back(std::make_shared<wb::SampleState>());
    smp_pred[0]->timestamp = imu_ref[0].timestamp;
    smp_pred[0]->grav      = {0.0, 0.0, -9.81};
    wb::PredictStates(imu_pred, 2, smp_pred, 0.08, 2);
    double pred_err = 0;
    for (size_t i = 0; i < imu_pred.size(); ++i)
      for (int k = 0; k < 3; ++k) pred_err = std::max(pred_err, std::abs(imu_pred[i].pos[k] - imu_ref[i].pos[k]));

    // error path: CHECK(pt.time >= back().time) (lidar_odometry.cc:491) -> Error{WC_EINVAL_TIME_ORDER}
    int  thrown = 0;
    auto bad    = points;
    bad[bad.size() / 2].time = bad[0].time - 1.0;
    try {
      std::deque<wb::Surfel::Ptr> tmp;
      wb::BuildSurfels(bad, tmp);
    } catch (const wb::Error& e) {
      thrown = (int)e.status;
    }

    std::printf("surfels %zu fix %zu corr_sld %zu corr_fix %zu iterations %d termination %d initial_cost %.17g final_cost %.17g "
                "residual_cor %.3g spline_err %.3g outside_null %d thrown %d kept %zu und_sum %.17g pred_err %.3g new_samples %zu\n",
                surfels_sld_win.size(), surfels_fix_win.size(), surfel_corrs_sld.size(), surfel_corrs_fix.size(), summary.num_iterations,
                summary.termination, summary.initial_cost, summary.final_cost, residual_cor, spline_err, (int)outside_null, thrown,
                points_buff.size(), und_sum, pred_err, smp_pred.size() - 1);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "mirror_test failed: %s\n", e.what());
    return 1;
  }
}
