// The C++ mirror's ImuResampler against the known answers of the reference's own test (src/sensor/imu_resampler_test.cc:7-31):
// 10 Hz re-sampling of two samples one second apart gives timestamps 0, 0.1, 0.2 — compared exactly, as EXPECT_EQ does
// upstream — and the third sample carries the 0.8 / 0.2 interpolation weights.  Host logic only: no device call.
#include <cmath>
#include <cstdio>

#include "wildcat_b200.hpp"

namespace wb = wildcat_b200;

static int fails = 0;
#define EXPECT(cond)                                               \
  do {                                                             \
    if (!(cond)) std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #cond), ++fails; \
  } while (0)

static bool Approx(const wb::Vector3d& a, const wb::Vector3d& b) {  // Eigen's isApprox: |a - b|^2 <= 1e-24 min(|a|^2, |b|^2)
  double d = 0, na = 0, nb = 0;
  for (int k = 0; k < 3; ++k) d += (a[k] - b[k]) * (a[k] - b[k]), na += a[k] * a[k], nb += b[k] * b[k];
  return d <= 1e-24 * std::fmin(na, nb);
}

int main() {
  wb::ImuResampler  ir(10);
  const wb::ImuData imu1{0, {1, 2, 3}, {435, 342, 434}}, imu2{1, {11, 234, 453}, {234, 46, 32}};
  ir.AddImuData(imu1);
  EXPECT(!ir.AdvanceGetResampledImuData());  // one raw sample is not a bracket yet (imu_resampler.h:24)
  ir.AddImuData(imu2);
  auto r = ir.AdvanceGetResampledImuData();
  EXPECT(r && r->timestamp == 0);
  r = ir.AdvanceGetResampledImuData();
  EXPECT(r && r->timestamp == 0.1);
  r = ir.AdvanceGetResampledImuData();
  EXPECT(r && r->timestamp == 0.2);
  wb::Vector3d gyr, acc;
  for (int k = 0; k < 3; ++k) {
    gyr[k] = 0.8 * imu1.angular_velocity[k] + 0.2 * imu2.angular_velocity[k];
    acc[k] = 0.8 * imu1.linear_acceleration[k] + 0.2 * imu2.linear_acceleration[k];
  }
  EXPECT(r && Approx(gyr, r->angular_velocity) && Approx(acc, r->linear_acceleration));
  // a third raw sample replaces the older end of the bracket; an instant outside the new bracket yields nothing
  ir.AddImuData(wb::ImuData{1.05, {0, 0, 0}, {0, 0, 0}});
  EXPECT(!ir.AdvanceGetResampledImuData());  // 0.3 is not inside [1, 1.05]
  std::printf(fails ? "resampler_test: %d failure(s)\n" : "resampler_test ok\n", fails);
  return fails ? 1 : 0;
}
