"""Minimal ROS bag v2.0 / sensor_msgs::PointCloud2 reader (host side, numpy only).

Replaces, for the hot path's input side, what the reference node does with rosbag::View and
pcl::fromROSMsg (ilcc2/test/get_lidar_corners.cpp:136-164): open the bag, take the FIRST message on the
LiDAR topic, hand its point records on.  Only what the bundled bags need is implemented: uncompressed
chunks (compression=none), record ops 2 (message), 5 (chunk), 7 (connection).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

_MAGIC = b"#ROSBAG V2.0\n"


def _parse_header(buf: bytes) -> dict:
    out, pos = {}, 0
    while pos < len(buf):
        (flen,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        field = buf[pos:pos + flen]
        pos += flen
        k, _, v = field.partition(b"=")
        out[k.decode()] = v
    return out


def _records(buf: bytes, pos: int = 0):
    n = len(buf)
    while pos + 8 <= n:
        (hlen,) = struct.unpack_from("<I", buf, pos)
        hdr = _parse_header(buf[pos + 4:pos + 4 + hlen])
        pos += 4 + hlen
        (dlen,) = struct.unpack_from("<I", buf, pos)
        data = buf[pos + 4:pos + 4 + dlen]
        pos += 4 + dlen
        yield hdr, data


def iter_messages(path: str, topic: str):
    """Yield (time_sec, serialized message bytes) for `topic`, in file order."""
    with open(path, "rb") as f:
        raw = f.read()
    if not raw.startswith(_MAGIC):
        raise ValueError(f"{path}: not a ROS bag v2.0")
    conns = {}
    for hdr, data in _records(raw, len(_MAGIC)):
        op = hdr.get("op", b"\xff")[0]
        if op == 5:  # chunk
            if hdr.get("compression", b"none") != b"none":
                raise NotImplementedError("compressed bag chunks are not supported")
            for h2, d2 in _records(data):
                op2 = h2["op"][0]
                if op2 == 7:
                    conns[struct.unpack("<I", h2["conn"])[0]] = h2["topic"].decode()
                elif op2 == 2:
                    c = struct.unpack("<I", h2["conn"])[0]
                    if conns.get(c) == topic:
                        sec, nsec = struct.unpack("<II", h2["time"])
                        yield sec + nsec * 1e-9, d2
        elif op == 7:
            conns[struct.unpack("<I", hdr["conn"])[0]] = hdr["topic"].decode()


@dataclass
class PointCloud2:
    """The fields of sensor_msgs/PointCloud2 that the hot path needs; `data` is the raw record buffer."""
    frame_id: str
    height: int
    width: int
    fields: dict  # name -> (offset, datatype, count)
    is_bigendian: bool
    point_step: int
    row_step: int
    data: np.ndarray  # uint8, width*height*point_step
    is_dense: bool

    @property
    def n_points(self) -> int:
        return self.width * self.height

    def offsets_xyzi(self):
        """Byte offsets of x, y, z, intensity (all FLOAT32=7), i.e. what pcl::fromROSMsg maps by field name."""
        offs = []
        for name in ("x", "y", "z", "intensity"):
            off, dt, _ = self.fields[name]
            if dt != 7:
                raise ValueError(f"field {name} is not FLOAT32")
            offs.append(off)
        return tuple(offs)

    def records_f32(self) -> np.ndarray:
        """(N, point_step/4) float32 view of the records (zero copy)."""
        assert self.point_step % 4 == 0 and not self.is_bigendian
        return self.data.view(np.float32).reshape(self.n_points, self.point_step // 4)

    def packed_xyzi(self) -> np.ndarray:
        """(N,4) float32 copy in the packed 16-byte XYZI layout."""
        rec = self.records_f32()
        ox, oy, oz, oi = (o // 4 for o in self.offsets_xyzi())
        return np.ascontiguousarray(rec[:, [ox, oy, oz, oi]])

    def field_u16(self, name: str) -> np.ndarray:
        off, dt, _ = self.fields[name]
        assert dt == 4
        return self.data.reshape(self.n_points, self.point_step)[:, off:off + 2].copy().view(np.uint16).reshape(-1)


def parse_pointcloud2(msg: bytes) -> PointCloud2:
    pos = 0

    def u32():
        nonlocal pos
        (v,) = struct.unpack_from("<I", msg, pos)
        pos += 4
        return v

    def u8():
        nonlocal pos
        v = msg[pos]
        pos += 1
        return v

    def string():
        nonlocal pos
        n = u32()
        s = msg[pos:pos + n].decode()
        pos += n
        return s

    u32(); u32(); u32()  # seq, stamp.sec, stamp.nsec
    frame_id = string()
    height, width = u32(), u32()
    fields = {}
    for _ in range(u32()):
        name = string()
        off = u32()
        dt = u8()
        cnt = u32()
        fields[name] = (off, dt, cnt)
    big = bool(u8())
    point_step, row_step = u32(), u32()
    dlen = u32()
    data = np.frombuffer(msg, np.uint8, dlen, pos).copy()
    pos += dlen
    dense = bool(u8())
    return PointCloud2(frame_id, height, width, fields, big, point_step, row_step, data, dense)


def first_cloud(path: str, topic: str = "/velodyne_points") -> PointCloud2:
    """First PointCloud2 on `topic` (get_lidar_corners.cpp:144-154 keeps the first and breaks)."""
    for _, msg in iter_messages(path, topic):
        return parse_pointcloud2(msg)
    raise ValueError(f"{path}: no message on {topic}")


def write_pointcloud2_bag(path: str, records: np.ndarray, topic: str = "/velodyne_points", frame_id: str = "velodyne"):
    """Write a one-message bag v2.0 (uncompressed chunk) holding `records` ((N,8) float32 = 32-byte velodyne records:
    x@0 y@4 z@8 intensity@16 ring(u16)@20), in the layout of the bundled bags. Used to feed the C++ host tool in tests."""
    rec = np.ascontiguousarray(records, np.float32)
    assert rec.ndim == 2 and rec.shape[1] == 8

    def hdr(fields):
        b = b"".join(struct.pack("<I", len(k) + 1 + len(v)) + k.encode() + b"=" + v for k, v in fields)
        return struct.pack("<I", len(b)) + b

    def record(h, data):
        return h + struct.pack("<I", len(data)) + data

    n = rec.shape[0]
    flds = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 16, 7), ("ring", 20, 4)]
    msg = struct.pack("<III", 0, 0, 0) + struct.pack("<I", len(frame_id)) + frame_id.encode() + struct.pack("<II", 1, n)
    msg += struct.pack("<I", len(flds))
    for name, off, dt in flds:
        msg += struct.pack("<I", len(name)) + name.encode() + struct.pack("<IBI", off, dt, 1)
    msg += struct.pack("<BII", 0, 32, 32 * n) + struct.pack("<I", 32 * n) + rec.tobytes() + struct.pack("<B", 1)
    conn = record(hdr([("op", b"\x07"), ("conn", struct.pack("<I", 0)), ("topic", topic.encode())]),
                  hdr([("type", b"sensor_msgs/PointCloud2"), ("topic", topic.encode())]))
    m = record(hdr([("op", b"\x02"), ("conn", struct.pack("<I", 0)), ("time", struct.pack("<II", 1, 0))]), msg)
    chunk = record(hdr([("op", b"\x05"), ("compression", b"none"), ("size", struct.pack("<I", len(conn + m)))]), conn + m)
    with open(path, "wb") as f:
        f.write(_MAGIC + record(hdr([("op", b"\x03"), ("chunk_count", struct.pack("<I", 1)),
                                     ("conn_count", struct.pack("<I", 1)), ("index_pos", struct.pack("<Q", 0))]), b" " * 16))
        f.write(chunk)
