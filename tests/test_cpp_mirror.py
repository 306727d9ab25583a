"""The C++ host mirror (include/wildcat_b200.hpp: the reference's src/odometry entry points restated over the C ABI).

CPU: the header and its test program compile warning-free and link against the library (no device call is made).
GPU: tests/cpp/mirror_test.cc drives BuildSurfels -> UpdateSurfelPoses -> KnnSurfelMatcher x2 -> SolveWindow ->
ApplyCorrections -> CubicBSplineInterpolator exactly as LidarOdometry::AddLidarScan drives the reference
(lidar_odometry.cc:523-566); its outputs must equal those of the Python mirror on the same window."""
import os
import subprocess

import numpy as np
import pytest

from wildcat_slam_b200 import abi
from wildcat_slam_b200 import synthetic as S
from wildcat_slam_b200 import types as T

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _build(tmp_path):
    if not os.path.exists(abi.SO_PATH):
        from wildcat_slam_b200 import build

        build.build()
    exe = str(tmp_path / "mirror_test")
    libdir = os.path.dirname(abi.SO_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_test.cc"), "-o", exe, "-L", libdir, "-l:libwildcat_b200.so",
                           f"-Wl,-rpath,{libdir}"])
    return exe


def test_mirror_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "missing")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr  # fails on its inputs, before any device call


def test_cpp_imu_resampler_known_answer_of_the_reference(tmp_path):
    """tests/cpp/resampler_test.cc replays src/sensor/imu_resampler_test.cc:7-31 on the C++ mirror's ImuResampler (host only)"""
    exe = str(tmp_path / "resampler_test")
    libdir = os.path.dirname(abi.SO_PATH)
    if not os.path.exists(abi.SO_PATH):
        from wildcat_slam_b200 import build

        build.build()
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "resampler_test.cc"), "-o", exe, "-L", libdir, "-l:libwildcat_b200.so",
                           f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "resampler_test ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_mirror_pipeline_equals_python_mirror(tmp_path):
    from wildcat_slam_b200 import odometry as od

    exe = _build(tmp_path)
    w = S.make_window("C1")
    for name, arr in (("points", w.points), ("imu", w.imu), ("samples", w.samples), ("fix_points", w.fix_points), ("fix_imu", w.fix_imu)):
        np.ascontiguousarray(arr).tofile(tmp_path / f"{name}.bin")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    tok = r.stdout.split()
    out = {tok[i]: tok[i + 1] for i in range(0, len(tok), 2)}

    ctx = od.Context(0)
    try:
        sld = od.UpdateSurfelPoses(w.imu, od.BuildSurfels(w.points, ctx=ctx), ctx=ctx)
        fix = od.UpdateSurfelPoses(w.fix_imu, od.BuildSurfels(w.fix_points, ctx=ctx), ctx=ctx)
        m = od.KnnSurfelMatcher(ctx); m.BuildIndex(sld); cs, _ = m.Match(sld)
        m2 = od.KnnSurfelMatcher(ctx); m2.BuildIndex(fix); cf, _ = m2.Match(sld)
        smp, sg = od.SolveWindow(sld, fix, cs, cf, w.imu, w.samples, ctx=ctx)
        kept = od.FilterPoints(w.points, ctx=ctx)
        und = od.UndistortSweep(w.points, w.imu, ctx=ctx)
    finally:
        ctx.close()
    assert int(out["surfels"]) == len(sld) > 0 and int(out["fix"]) == len(fix)
    cpp_sld = np.fromfile(tmp_path / "cpp_sld.bin", dtype=T.SURFEL)
    for f in T.SURFEL.names:
        if not f.startswith("_"):
            assert np.array_equal(cpp_sld[f], sld[f]), f                     # same kernels, same inputs: bitwise
    assert np.array_equal(np.fromfile(tmp_path / "cpp_corr_sld.bin", dtype=np.int32).reshape(-1, 2), np.stack([cs["s1"], cs["s2"]], 1))
    assert np.array_equal(np.fromfile(tmp_path / "cpp_corr_fix.bin", dtype=np.int32).reshape(-1, 2), np.stack([cf["s1"], cf["s2"]], 1))
    assert int(out["iterations"]) == sg.num_iterations and int(out["termination"]) == sg.termination
    assert float(out["initial_cost"]) == pytest.approx(sg.initial_cost, rel=1e-12)
    assert float(out["final_cost"]) == pytest.approx(sg.final_cost, rel=1e-9)   # fp64 atomics: summation order varies
    np.testing.assert_allclose(np.fromfile(tmp_path / "cpp_data_cor.bin").reshape(-1, 12), smp["data_cor"], rtol=0, atol=1e-9)
    assert float(out["residual_cor"]) == 0.0                                     # UpdateSamplePoses zeroes the corrections
    assert float(out["spline_err"]) < 1e-6 and out["outside_null"] == "1"        # spline_interpolation_test.cc:79-96, :52-54
    assert int(out["thrown"]) == T.WC_EINVAL_TIME_ORDER                           # CHECK lidar_odometry.cc:491
    assert int(out["kept"]) == len(kept)                                          # FilterPoints
    und_sum = float(und["x"].astype(np.float64).sum() + und["y"].astype(np.float64).sum() + und["z"].astype(np.float64).sum())
    assert float(out["und_sum"]) == pytest.approx(und_sum, rel=1e-9)              # UndistortSweep (same kernel; summation order only)
    assert float(out["pred_err"]) < 1e-10 and int(out["new_samples"]) == 2        # PredictStates reproduces the generator's prediction
