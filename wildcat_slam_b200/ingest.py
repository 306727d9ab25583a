"""Wire / disk ingestion without ROS: what wildcat_slam_node.cc does with rosbag::View,
sensor_msgs/PointCloud2 and sensor_msgs/Imu (wildcat_slam_node.cc:30-52,83-99), restated for the raw bytes.

  * parse_pointcloud2 / parse_imu   ROS1 message deserialisers (little-endian wire format)
  * pointcloud2_layout              pcl::fromROSMsg's field matching for the point type registered at common.h:21-28:
                                    a struct field is filled from the message field with the same NAME, datatype and
                                    count 1; anything else is left zero
  * UnpackPointCloud2               payload -> 48-byte hilti_ros::Point records, on the device (wc_unpack_pointcloud2)
  * BagReader / BagWriter           rosbag format 2.0, sequential: bag header, chunks (none / bz2), connection and
                                    message-data records; index records are skipped (a replay reads front to back like
                                    rosbag::View on an unfiltered bag).  The writer exists for tests and for exporting
                                    synthetic sweeps.

  * ImuResampler                    the fixed-rate IMU re-sampling between the bag and LidarOdometry::AddImuData
                                    (src/sensor/imu_resampler.h:12-53, HandleImuMessage wildcat_slam_node.cc:30-44)

The numeric path of the payload (field extraction) runs on the GPU; everything here is header parsing and O(200 Hz) host
logic.
"""
import bz2
import struct
from dataclasses import dataclass, field

import numpy as np

from . import types as T

# sensor_msgs/PointField datatypes
INT8, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 = 1, 2, 3, 4, 5, 6, 7, 8
# the registered point type (common.h:21-28): struct member <- (message field name, datatype)
HILTI_FIELDS = {"x": ("x", FLOAT32), "y": ("y", FLOAT32), "z": ("z", FLOAT32), "intensity": ("intensity", FLOAT32),
                "time": ("timestamp", FLOAT64), "ring": ("ring", UINT16)}


@dataclass
class PointField:
    name: str
    offset: int
    datatype: int
    count: int = 1


@dataclass
class PointCloud2:
    stamp: float
    frame_id: str
    height: int
    width: int
    fields: list
    is_bigendian: bool
    point_step: int
    row_step: int
    data: bytes
    is_dense: bool
    seq: int = 0

    @property
    def n_points(self):
        return self.height * self.width


@dataclass
class Imu:
    stamp: float
    frame_id: str
    angular_velocity: np.ndarray
    linear_acceleration: np.ndarray
    orientation: np.ndarray = field(default=None)


class _Cursor:
    def __init__(self, buf, pos=0):
        self.b, self.p = memoryview(buf), pos

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def string(self):
        n = self.take("I")
        s = bytes(self.b[self.p:self.p + n]).decode("utf-8", "replace")
        self.p += n
        return s

    def blob(self):
        n = self.take("I")
        v = self.b[self.p:self.p + n]
        self.p += n
        return v


def _header(c):
    seq = c.take("I")
    sec, nsec = c.take("II")
    return seq, sec + nsec * 1e-9, c.string()


def parse_pointcloud2(buf) -> PointCloud2:
    c = _Cursor(buf)
    seq, stamp, frame = _header(c)
    height, width = c.take("II")
    fields = []
    for _ in range(c.take("I")):
        name = c.string()
        off, dt, cnt = c.take("IBI")
        fields.append(PointField(name, off, dt, cnt))
    be = bool(c.take("B"))
    point_step, row_step = c.take("II")
    data = c.blob()
    dense = bool(c.take("B"))
    return PointCloud2(stamp, frame, height, width, fields, be, point_step, row_step, data, dense, seq)


def parse_imu(buf) -> Imu:
    c = _Cursor(buf)
    _, stamp, frame = _header(c)
    q = np.array(c.take("4d"))
    c.take("9d")
    w = np.array(c.take("3d"))
    c.take("9d")
    a = np.array(c.take("3d"))
    return Imu(stamp, frame, w, a, q)


def pointcloud2_layout(msg: PointCloud2) -> T.Pc2Layout:
    """pcl::fromROSMsg's mapping for hilti_ros::Point: by name, datatype and count; unmatched members stay zero."""
    if msg.is_bigendian:
        raise ValueError("big-endian PointCloud2 payloads are not supported")
    off = {}
    for member, (name, dt) in HILTI_FIELDS.items():
        off[member] = -1
        for f in msg.fields:
            if f.name == name and f.datatype == dt and f.count == 1:
                off[member] = f.offset
    return T.Pc2Layout(msg.point_step, off["x"], off["y"], off["z"], off["intensity"], off["time"], off["ring"])


def unpack_pointcloud2_host(msg: PointCloud2) -> np.ndarray:
    """numpy restatement of the same mapping (test oracle for the device kernel; not used by the product path)."""
    L = pointcloud2_layout(msg)
    n = msg.n_points
    raw = np.frombuffer(msg.data, dtype=np.uint8, count=n * msg.point_step).reshape(n, msg.point_step)
    out = np.zeros(n, dtype=T.POINT48)
    for member, off, dt in (("x", L.off_x, "<f4"), ("y", L.off_y, "<f4"), ("z", L.off_z, "<f4"), ("intensity", L.off_intensity, "<f4"),
                            ("time", L.off_time, "<f8"), ("ring", L.off_ring, "<u2")):
        if off >= 0:
            w = np.dtype(dt).itemsize
            out[member] = np.ascontiguousarray(raw[:, off:off + w]).view(dt)[:, 0]
    return out


def UnpackPointCloud2(msg: PointCloud2, ctx=None, out=None) -> np.ndarray:
    """pcl::fromROSMsg(*msg, *cloud) on the device: PointCloud2 payload -> hilti_ros::Point records."""
    from . import odometry as od

    ctx = ctx or od.default_context()
    L = pointcloud2_layout(msg)
    n = msg.n_points
    if len(msg.data) < n * msg.point_step:
        raise ValueError("PointCloud2 data shorter than height * width * point_step")
    out = np.zeros(n, dtype=T.POINT48) if out is None else out[:n]
    data = np.frombuffer(msg.data, dtype=np.uint8)
    ctx.check(ctx.lib.wc_unpack_pointcloud2(ctx.handle, T.ptr(data), n, L, T.ptr(out)), "wc_unpack_pointcloud2")
    return out


# ----------------------------------------------------------------------------------------------- serialisers (tests)
def _ser_header(seq, stamp, frame):
    sec = int(np.floor(stamp))
    nsec = int(round((stamp - sec) * 1e9))
    if nsec >= 1_000_000_000:
        sec, nsec = sec + 1, nsec - 1_000_000_000
    f = frame.encode()
    return struct.pack("<III", seq, sec, nsec) + struct.pack("<I", len(f)) + f


def serialize_pointcloud2(points: np.ndarray, stamp, frame_id="lidar", fields=None, point_step=None, extra_pad=0) -> bytes:
    """hilti_ros::Point records -> sensor_msgs/PointCloud2 bytes.  `fields` lets a test choose another on-wire layout:
    list of (name, offset, datatype); default is the packed layout x y z intensity timestamp ring."""
    points = np.ascontiguousarray(points, dtype=T.POINT48)
    if fields is None:
        fields = [("x", 0, FLOAT32), ("y", 4, FLOAT32), ("z", 8, FLOAT32), ("intensity", 12, FLOAT32), ("timestamp", 16, FLOAT64),
                  ("ring", 24, UINT16)]
        point_step = 26 + extra_pad
    n = len(points)
    raw = np.zeros((n, point_step), dtype=np.uint8)
    src = {"x": ("x", "<f4"), "y": ("y", "<f4"), "z": ("z", "<f4"), "intensity": ("intensity", "<f4"), "timestamp": ("time", "<f8"),
           "ring": ("ring", "<u2")}
    for name, off, dt in fields:
        if name in src:
            member, np_dt = src[name]
            want = {"<f4": FLOAT32, "<f8": FLOAT64, "<u2": UINT16}[np_dt]
            if dt != want:   # a test layout with a mismatching datatype: write something, the reader must ignore it
                continue
            b = np.ascontiguousarray(points[member].astype(np_dt)).view(np.uint8).reshape(n, -1)
            raw[:, off:off + b.shape[1]] = b
    out = _ser_header(0, stamp, frame_id) + struct.pack("<II", 1, n) + struct.pack("<I", len(fields))
    for name, off, dt in fields:
        nm = name.encode()
        out += struct.pack("<I", len(nm)) + nm + struct.pack("<IBI", off, dt, 1)
    data = raw.tobytes()
    out += struct.pack("<BII", 0, point_step, point_step * n) + struct.pack("<I", len(data)) + data + struct.pack("<B", 1)
    return out


def serialize_imu(stamp, gyr, acc, frame_id="imu") -> bytes:
    z9 = struct.pack("<9d", *([0.0] * 9))
    return (_ser_header(0, stamp, frame_id) + struct.pack("<4d", 0, 0, 0, 1) + z9 + struct.pack("<3d", *gyr) + z9 +
            struct.pack("<3d", *acc) + z9)


# ------------------------------------------------------------------------------------------------------- rosbag 2.0
_MAGIC = b"#ROSBAG V2.0\n"
OP_MSG, OP_BAG_HEADER, OP_INDEX, OP_CHUNK, OP_CHUNK_INFO, OP_CONNECTION = 0x02, 0x03, 0x04, 0x05, 0x06, 0x07


def _parse_fields(hdr):
    out, p = {}, 0
    while p < len(hdr):
        n = struct.unpack_from("<I", hdr, p)[0]
        p += 4
        k, _, v = bytes(hdr[p:p + n]).partition(b"=")
        out[k.decode()] = v
        p += n
    return out


def _records(buf, pos, end):
    while pos < end:
        hl = struct.unpack_from("<I", buf, pos)[0]
        hdr = _parse_fields(buf[pos + 4:pos + 4 + hl])
        pos += 4 + hl
        dl = struct.unpack_from("<I", buf, pos)[0]
        data = buf[pos + 4:pos + 4 + dl]
        pos += 4 + dl
        yield hdr, data


class BagReader:
    """for topic, msg_type, t, raw_message in BagReader(path): ...   (front to back, like rosbag::View on the whole bag)"""

    def __init__(self, path):
        self.buf = memoryview(open(path, "rb").read())
        if bytes(self.buf[:len(_MAGIC)]) != _MAGIC:
            raise ValueError("not a rosbag 2.0 file")
        self.connections = {}

    def _connection(self, hdr, data):
        f = _parse_fields(data)
        self.connections[struct.unpack("<I", hdr["conn"])[0]] = (hdr["topic"].decode(), f.get("type", b"").decode())

    def __iter__(self):
        for hdr, data in _records(self.buf, len(_MAGIC), len(self.buf)):
            op = hdr["op"][0]
            if op == OP_CHUNK:
                comp = hdr["compression"].decode()
                if comp == "none":
                    chunk = data
                elif comp == "bz2":
                    chunk = memoryview(bz2.decompress(bytes(data)))
                else:
                    raise ValueError(f"chunk compression '{comp}' is not supported (no lz4 in this image)")
                for h2, d2 in _records(chunk, 0, len(chunk)):
                    op2 = h2["op"][0]
                    if op2 == OP_CONNECTION:
                        self._connection(h2, d2)
                    elif op2 == OP_MSG:
                        conn = struct.unpack("<I", h2["conn"])[0]
                        sec, nsec = struct.unpack("<II", h2["time"])
                        topic, typ = self.connections[conn]
                        yield topic, typ, sec + nsec * 1e-9, d2
            elif op == OP_CONNECTION:
                self._connection(hdr, data)
            # bag header, index data and chunk info records carry nothing a sequential replay needs


def _rec(fields, data):
    h = b""
    for k, v in fields:
        kv = k.encode() + b"=" + v
        h += struct.pack("<I", len(kv)) + kv
    return struct.pack("<I", len(h)) + h + struct.pack("<I", len(data)) + data


class BagWriter:
    """minimal rosbag 2.0 writer (one chunk per flush, optional bz2): enough structure for rosbag tools to read the
    file sequentially; no index records."""

    def __init__(self, path, compression="none", chunk_msgs=64):
        self.f = open(path, "wb")
        self.compression, self.chunk_msgs = compression, chunk_msgs
        self.conns, self.pending, self.n_chunks = {}, [], 0
        self.f.write(_MAGIC)
        self._write_bag_header()

    def _write_bag_header(self):
        rec = _rec([("op", bytes([OP_BAG_HEADER])), ("index_pos", struct.pack("<Q", 0)), ("conn_count", struct.pack("<I", len(self.conns))),
                    ("chunk_count", struct.pack("<I", self.n_chunks))], b"")
        pad = 4096 - len(rec)   # the bag header record is padded to 4096 bytes
        rec = rec[:-4] + struct.pack("<I", pad) + b" " * pad
        self.f.seek(len(_MAGIC))
        self.f.write(rec)
        self.f.seek(0, 2)

    def write(self, topic, msg_type, t, raw):
        if topic not in self.conns:
            cid = len(self.conns)
            self.conns[topic] = cid
            ch = b""
            for k, v in (("topic", topic), ("type", msg_type), ("md5sum", "*"), ("message_definition", "")):
                kv = k.encode() + b"=" + v.encode()
                ch += struct.pack("<I", len(kv)) + kv
            self.pending.append(_rec([("op", bytes([OP_CONNECTION])), ("conn", struct.pack("<I", cid)), ("topic", topic.encode())], ch))
        sec = int(np.floor(t))
        self.pending.append(_rec([("op", bytes([OP_MSG])), ("conn", struct.pack("<I", self.conns[topic])),
                                  ("time", struct.pack("<II", sec, int(round((t - sec) * 1e9))))], raw))
        if len(self.pending) >= self.chunk_msgs:
            self.flush()

    def flush(self):
        if not self.pending:
            return
        body = b"".join(self.pending)
        data = bz2.compress(body) if self.compression == "bz2" else body
        self.f.write(_rec([("op", bytes([OP_CHUNK])), ("compression", self.compression.encode()), ("size", struct.pack("<I", len(body)))], data))
        self.pending, self.n_chunks = [], self.n_chunks + 1

    def close(self):
        self.flush()
        self._write_bag_header()
        self.f.close()


# ---- fixed-rate IMU re-sampling ---------------------------------------------------------------------------------------
class ImuResampler:
    """src/sensor/imu_resampler.h:12-53.  Keeps the two most recent raw samples; every call of advance() yields at most one
    sample of the fixed-rate sequence: the first raw sample itself, then t_prev + 1 / freq whenever that instant lies
    inside the bracket of the two raw samples (both ends inclusive), linearly interpolated.  A bracket that does not
    contain the next instant yields nothing (the caller feeds the next raw sample and asks again, as HandleImuMessage does,
    wildcat_slam_node.cc:30-44) — so with raw data slower than 2 x freq the sequence stalls, exactly like upstream."""

    def __init__(self, freq):
        self.period = 1.0 / freq  # the reference divides 1.0 by the int rate on every call: the same double
        self.pair = []            # [(t, acc[3], gyr[3])], at most two
        self.t_prev = None

    def add(self, t, acc, gyr):
        """AddImuData (:16-21)"""
        self.pair.append((float(t), np.asarray(acc, dtype=np.float64), np.asarray(gyr, dtype=np.float64)))
        del self.pair[:-2]

    def advance(self):
        """AdvanceGetResampledImuData (:23-46): (t, acc, gyr) or None"""
        if len(self.pair) < 2:
            return None
        (t0, a0, g0), (t1, a1, g1) = self.pair
        if self.t_prev is None:
            self.t_prev = t0
            return t0, a0.copy(), g0.copy()
        t = self.t_prev + self.period
        if not (t0 <= t <= t1):
            return None
        f = (t - t0) / (t1 - t0)
        self.t_prev = t
        return t, (1 - f) * a0 + f * a1, (1 - f) * g0 + f * g1


def resample_imu(stamps, acc, gyr, freq):
    """HandleImuMessage over a whole recording (wildcat_slam_node.cc:30-44): one add + ONE advance per raw message, as the
    node does.  Returns (t[M], acc[M, 3], gyr[M, 3]) of what LidarOdometry::AddImuData would have received."""
    r, out = ImuResampler(freq), []
    for t, a, g in zip(stamps, acc, gyr):
        r.add(t, a, g)
        s = r.advance()
        if s is not None:
            out.append(s)
    if not out:
        return np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3))
    return np.array([o[0] for o in out]), np.stack([o[1] for o in out]), np.stack([o[2] for o in out])
