"""Wire / disk ingestion: ROS1 PointCloud2 / Imu deserialisation, pcl::fromROSMsg's field matching
for hilti_ros::Point, and sequential rosbag 2.0 reading — host logic on the CPU; the device unpack kernel is compared
with the numpy restatement under -m gpu."""
import numpy as np
import pytest

from wildcat_slam_b200 import ingest as I
from wildcat_slam_b200 import types as T


def _points(n=1000, seed=3):
    rng = np.random.default_rng(seed)
    p = np.zeros(n, dtype=T.POINT48)
    p["x"], p["y"], p["z"] = rng.normal(size=(3, n)).astype(np.float32) * 20
    p["intensity"] = rng.uniform(0, 255, n).astype(np.float32)
    p["time"] = 1.6e9 + np.cumsum(rng.uniform(1e-6, 1e-5, n))
    p["ring"] = rng.integers(0, 128, n)
    return p


# an on-wire layout unlike the struct: reordered fields, an unrelated field, unaligned offsets, a trailing pad
ODD_FIELDS = [("ring", 1, I.UINT16), ("timestamp", 3, I.FLOAT64), ("reflectivity", 11, I.UINT8), ("z", 12, I.FLOAT32),
              ("y", 17, I.FLOAT32), ("x", 21, I.FLOAT32), ("intensity", 26, I.FLOAT32)]


def test_pointcloud2_round_trip_default_and_odd_layout():
    p = _points()
    for fields, step in ((None, None), (ODD_FIELDS, 33)):
        msg = I.parse_pointcloud2(I.serialize_pointcloud2(p, p["time"][0], fields=fields, point_step=step))
        assert msg.n_points == len(p) and msg.stamp == pytest.approx(p["time"][0], abs=1e-6)
        q = I.unpack_pointcloud2_host(msg)
        for f in ("x", "y", "z", "intensity", "time", "ring"):
            np.testing.assert_array_equal(q[f], p[f])


def test_field_matching_follows_pcl():
    p = _points(50)
    # a 'timestamp' of the wrong datatype and a missing 'ring': both members stay zero, everything else is mapped
    fields = [("x", 0, I.FLOAT32), ("y", 4, I.FLOAT32), ("z", 8, I.FLOAT32), ("intensity", 12, I.FLOAT32), ("timestamp", 16, I.FLOAT32)]
    msg = I.parse_pointcloud2(I.serialize_pointcloud2(p, 5.0, fields=fields, point_step=20))
    L = I.pointcloud2_layout(msg)
    assert (L.off_time, L.off_ring, L.off_x, L.point_step) == (-1, -1, 0, 20)
    q = I.unpack_pointcloud2_host(msg)
    assert (q["time"] == 0).all() and (q["ring"] == 0).all()
    np.testing.assert_array_equal(q["x"], p["x"])


def test_imu_round_trip():
    m = I.parse_imu(I.serialize_imu(1234.5678, [0.1, -0.2, 0.3], [9.7, 0.1, -0.4]))
    assert m.stamp == pytest.approx(1234.5678, abs=1e-9)
    np.testing.assert_array_equal(m.angular_velocity, [0.1, -0.2, 0.3])
    np.testing.assert_array_equal(m.linear_acceleration, [9.7, 0.1, -0.4])


@pytest.mark.parametrize("compression", ["none", "bz2"])
def test_rosbag_sequential_replay(tmp_path, compression):
    """the offline mode of wildcat_slam_node.cc:83-99: IMU and lidar messages come back in file order with their types"""
    path = str(tmp_path / "t.bag")
    w = I.BagWriter(path, compression=compression, chunk_msgs=5)
    sent = []
    p = _points(300)
    for k in range(12):
        t = 100.0 + 0.01 * k
        if k % 4 == 3:
            raw = I.serialize_pointcloud2(p[k * 20:(k + 1) * 20], t)
            w.write("/hesai/pandar", "sensor_msgs/PointCloud2", t, raw)
            sent.append(("/hesai/pandar", "sensor_msgs/PointCloud2", raw))
        else:
            raw = I.serialize_imu(t, [k, 0, 0], [0, 0, 9.81])
            w.write("/alphasense/imu", "sensor_msgs/Imu", t, raw)
            sent.append(("/alphasense/imu", "sensor_msgs/Imu", raw))
    w.close()
    got = [(topic, typ, bytes(raw)) for topic, typ, t, raw in I.BagReader(path)]
    assert got == sent
    clouds = [I.unpack_pointcloud2_host(I.parse_pointcloud2(raw)) for topic, typ, raw in got if typ.endswith("PointCloud2")]
    np.testing.assert_array_equal(np.concatenate(clouds)["time"], np.concatenate([p[k * 20:(k + 1) * 20] for k in (3, 7, 11)])["time"])


@pytest.mark.gpu
def test_device_unpack_matches_host_mapping():
    from wildcat_slam_b200 import odometry as od

    ctx = od.Context(0)
    try:
        p = _points(100_003)
        for fields, step in ((None, None), (ODD_FIELDS, 33)):
            msg = I.parse_pointcloud2(I.serialize_pointcloud2(p, p["time"][0], fields=fields, point_step=step))
            g = I.UnpackPointCloud2(msg, ctx=ctx)
            assert g.tobytes() == I.unpack_pointcloud2_host(msg).tobytes()
        # unpacked records feed the sweep preparation unchanged
        flt = T.default_sweep_filter()
        assert len(od.FilterPoints(g, flt, ctx=ctx)) <= len(g)
    finally:
        ctx.close()


def test_imu_resampler_known_answer_of_the_reference():
    """src/sensor/imu_resampler_test.cc:7-31 — 10 Hz re-sampling of two samples one second apart: timestamps 0, 0.1, 0.2
    (compared with EXPECT_EQ upstream, i.e. exactly) and the 0.8 / 0.2 interpolation weights of the third sample"""
    ir = I.ImuResampler(10)
    acc1, gyr1 = np.array([1.0, 2.0, 3.0]), np.array([435.0, 342.0, 434.0])      # ImuData{t, linear_acceleration, angular_velocity}
    acc2, gyr2 = np.array([11.0, 234.0, 453.0]), np.array([234.0, 46.0, 32.0])
    ir.add(0, acc1, gyr1)
    assert ir.advance() is None                                               # one sample only: nothing yet (:24)
    ir.add(1, acc2, gyr2)
    t, a, g = ir.advance()
    assert t == 0 and np.array_equal(a, acc1) and np.array_equal(g, gyr1)
    t, a, g = ir.advance()
    assert t == 0.1
    t, a, g = ir.advance()
    assert t == 0.2
    np.testing.assert_allclose(g, 0.8 * gyr1 + 0.2 * gyr2, rtol=1e-12)         # isApprox upstream
    np.testing.assert_allclose(a, 0.8 * acc1 + 0.2 * acc2, rtol=1e-12)


def test_imu_resampler_stream_follows_the_node():
    """one add + one advance per message (wildcat_slam_node.cc:30-44): a 400 Hz recording re-sampled to 200 Hz gives a
    uniform 5 ms grid that the state prediction's spacing check accepts; a bracket that does not contain the next instant
    yields nothing, and the queue never holds more than two samples"""
    rng = np.random.default_rng(2)
    n = 2000
    t = 100.0 + np.cumsum(rng.uniform(0.0023, 0.0027, n))                      # ~400 Hz with jitter
    acc, gyr = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    ts, a, g = I.resample_imu(t, acc, gyr, 200)
    assert ts[0] == t[0] and len(ts) > 0.45 * n
    np.testing.assert_allclose(np.diff(ts), 0.005, rtol=0, atol=1e-9)
    k = 7                                                                        # every sample is a lerp of its bracket
    j = np.searchsorted(t, ts[k], side="right") - 1
    f = (ts[k] - t[j]) / (t[j + 1] - t[j])
    np.testing.assert_allclose(a[k], (1 - f) * acc[j] + f * acc[j + 1], rtol=1e-9, atol=1e-12)
    slow = I.ImuResampler(200)                                                   # raw data slower than the grid: stalls like upstream
    slow.add(0.0, acc[0], gyr[0]); slow.add(1.0, acc[1], gyr[1])
    assert slow.advance()[0] == 0.0 and slow.advance()[0] == 0.005
    slow.add(1.5, acc[2], gyr[2])                                                # bracket [1, 1.5] does not contain 0.01
    assert slow.advance() is None and len(slow.pair) == 2
