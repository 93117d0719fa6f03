"""CPU: the C-ABI library loads, exports every symbol include/ilcc2_b200.h declares, and its host-only entry points
behave like the reference's host code (no compute calls: there is no GPU here, and the library has no CPU path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ilcc2_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ilcc2_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(product_lib):
    from lidar_camera_calibration_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(product_lib, n), f"{n} declared in include/ilcc2_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS)
    assert product_lib.ilcc2_version() == 100


def test_struct_layouts_match_the_oracle_and_the_header(product_lib, oracle):
    from lidar_camera_calibration_b200 import _lib
    assert C.sizeof(_lib.Params) == C.sizeof(oracle.Params) == 96          # sizeof(ilcc2_params) in the header
    assert C.sizeof(_lib.Diag) == C.sizeof(oracle.Diag) == 320             # sizeof(ilcc2_diag)
    assert [f[0] for f in _lib.Diag._fields_] == [f[0] for f in oracle.Diag._fields_]


def test_default_params_are_the_reference_literals(product_lib):
    from lidar_camera_calibration_b200 import api
    p = api.default_params()
    assert (p.roi_half_x, p.roi_half_y, p.roi_half_z) == (1.0, 1.5, 2.0)            # LidarCornersEst.cpp:54-64
    assert abs(p.cluster_tolerance - 0.12) < 1e-7 and (p.min_cluster_size, p.max_cluster_size) == (100, 25000)
    assert p.ransac_threshold == 0.03 and p.hist_bins == 100 and p.gray_rate == 2.5 and p.huber_delta == 0.1
    assert p.grid_length == 0.15 and (p.cells_x, p.cells_y) == (6, 8)              # pointgrey.yaml: 7x5 corners


def test_set_chessboard_param_reads_the_reference_yaml(product_lib, fixtures, tmp_path):
    from lidar_camera_calibration_b200 import api
    y = tmp_path / "pointgrey.yaml"
    y.write_text(fixtures["yaml_text"])
    p = api.default_params(grid_length=1.0, cells_x=2, cells_y=2)
    assert api.set_chessboard_param(p, str(y))
    assert p.grid_length == fixtures["yaml"]["grid_length"] == 0.15
    # m_grid_size = (corner_in_x+1, corner_in_y+1) sorted ascending (LidarCornersEst.cpp:31-39)
    assert (p.cells_x, p.cells_y) == (6, 8)
    swapped = tmp_path / "swapped.yaml"
    swapped.write_text("%YAML:1.0\ngrid_length: 0.2\ncorner_in_x: 5\ncorner_in_y: 9\n")
    assert api.set_chessboard_param(p, str(swapped)) and (p.cells_x, p.cells_y, p.grid_length) == (6, 10, 0.2)
    assert not api.set_chessboard_param(p, str(tmp_path / "missing.yaml"))           # "can not open": returns false


def test_save_corners2txt_reproduces_the_shipped_file_byte_for_byte(product_lib, fixtures, tmp_path):
    """get_lidar_corners.cpp:27-36 streams floats with the default precision; writing the fixture's own numbers back
    must give the shipped process_data/pointgrey_lidar_1.txt exactly."""
    from lidar_camera_calibration_b200 import api
    corners = np.asarray(fixtures["lidar_corners"]["1"], np.float32)
    out = tmp_path / "pointgrey_lidar_1.txt"
    api.save_corners2txt(corners, str(out))
    assert out.read_text() == fixtures["lidar_corner_text_1"]


def test_shard_range_is_a_contiguous_partition(product_lib):
    from lidar_camera_calibration_b200 import api, dist
    for n, w in [(8192, 8), (1000, 3), (5, 8), (0, 4)]:
        cover = []
        for r in range(w):
            b, c = api.shard_range(n, w, r)
            assert (b, c) == dist.shard_range(n, w, r)
            cover += list(range(b, b + c))
        assert cover == list(range(n))


def test_create_fails_loudly_without_a_gpu(product_lib):
    """No CPU fallback: without a usable sm_100 device ilcc2_create returns ILCC2_ERR_CUDA and says why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lidar_camera_calibration_b200 import api
    with pytest.raises(api.Ilcc2Error) as e:
        api.CornerExtractor(1, 1000)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_rosbag_reader_round_trip(tmp_path):
    """The bag reader used for the golden fixtures parses a minimal hand-written bag v2.0 (uncompressed chunk)."""
    import struct
    from lidar_camera_calibration_b200 import rosbag_lite as rb

    def hdr(**kw):
        b = b"".join(struct.pack("<I", len(k) + 1 + len(v)) + k.encode() + b"=" + v for k, v in kw.items())
        return struct.pack("<I", len(b)) + b

    def rec(h, data):
        return h + struct.pack("<I", len(data)) + data

    pts = np.arange(5 * 8, dtype=np.float32).reshape(5, 8)
    fields = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 16, 7), ("ring", 20, 4)]
    msg = struct.pack("<III", 1, 2, 3) + struct.pack("<I", 8) + b"velodyne" + struct.pack("<II", 1, 5)
    msg += struct.pack("<I", len(fields))
    for name, off, dt in fields:
        msg += struct.pack("<I", len(name)) + name.encode() + struct.pack("<IBI", off, dt, 1)
    msg += struct.pack("<BII", 0, 32, 160) + struct.pack("<I", 160) + pts.tobytes() + struct.pack("<B", 1)
    conn = rec(hdr(op=b"\x07", conn=struct.pack("<I", 1), topic=b"/velodyne_points"), hdr(type=b"sensor_msgs/PointCloud2"))
    m = rec(hdr(op=b"\x02", conn=struct.pack("<I", 1), time=struct.pack("<II", 10, 0)), msg)
    chunk = rec(hdr(op=b"\x05", compression=b"none", size=struct.pack("<I", len(conn + m))), conn + m)
    bag = tmp_path / "t.bag"
    bag.write_bytes(b"#ROSBAG V2.0\n" + rec(hdr(op=b"\x03"), b" " * 8) + chunk)
    pc = rb.first_cloud(str(bag))
    assert pc.n_points == 5 and pc.point_step == 32 and pc.offsets_xyzi() == (0, 4, 8, 16)
    assert np.array_equal(pc.records_f32(), pts)
    assert np.array_equal(pc.packed_xyzi(), pts[:, [0, 1, 2, 4]])


def test_calibration_host_entry_points_match_the_oracle_and_the_shipped_files(product_lib, oracle, fixtures, tmp_path):
    """The host-only entry points around the PnP (rough axes, text readers, camera yaml, composing and writing the
    extrinsic) need no GPU: they must agree with the oracle's restatement and with the reference's shipped artefacts."""
    from lidar_camera_calibration_b200 import api
    for cam in ("front", "car_left", "pointgrey", "left", "right", "back"):
        assert np.array_equal(api.axis_roughly(cam), oracle.axis_roughly(cam))
    with pytest.raises(api.Ilcc2Error):
        api.axis_roughly("no_such_camera")
    y = tmp_path / "pointgrey.yaml"
    y.write_text(fixtures["yaml_text"])
    K, cx, cy = api.read_camera_yaml(str(y))
    assert np.array_equal(K, np.asarray(fixtures["yaml"]["K"])) and (cx, cy) == (7, 5)
    for b in range(1, 7):
        lp, cp = tmp_path / f"l{b}.txt", tmp_path / f"c{b}.txt"
        lp.write_text(fixtures["lidar_corner_text"][str(b)])
        cp.write_text(fixtures["cam_corner_text"][str(b)])
        assert np.array_equal(api.read_lidar_corners(str(lp), 35), oracle.read_lidar_corners(str(lp), 35))
        assert np.array_equal(api.read_cam_corners(str(cp), 35, 5), oracle.read_cam_corners(str(cp), 35, 5))
    with pytest.raises(api.Ilcc2Error):
        api.read_lidar_corners(str(tmp_path / "missing.txt"), 35)
    # composing the shipped pose from the oracle's solution and writing it reproduces config/pointgrey.bin
    Tax = oracle.axis_roughly("pointgrey")
    P3, P2 = [], []
    for b in range(1, 7):
        p3 = oracle.read_lidar_corners(str(tmp_path / f"l{b}.txt"), 35) @ Tax[:3, :3].T + Tax[:3, 3]
        P3.append(oracle.order_lidar(p3, 7, 5))
        P2.append(oracle.order_cam(oracle.read_cam_corners(str(tmp_path / f"c{b}.txt"), 35, 5), 7, 5))
    r, t, _ = oracle.solve_pnp(np.concatenate(P3), np.concatenate(P2), [K[0, 0], K[0, 2], K[1, 1], K[1, 2]])
    T = api.compose_extrinsic(r, t, Tax)
    assert np.allclose(T, oracle.compose_extrinsic(r, t, Tax), rtol=0, atol=1e-15)
    out = tmp_path / "pose.bin"
    api.extrinsic2bin(str(out), T)
    got = np.fromfile(out, dtype="<f8")
    ref = np.frombuffer(bytes.fromhex(fixtures["extrinsic_bin_hex"]), dtype="<f8")
    assert got.shape == (16,) and np.abs(got - ref).max() < 1e-12
    assert np.array_equal(got.reshape(4, 4).T, T)  # column-major on disk (Eigen::Matrix4d)
