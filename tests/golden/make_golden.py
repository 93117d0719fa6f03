"""Regenerates tests/golden/* from the read-only reference checkout (run in the build container only).

    python tests/golden/make_golden.py [/root/reference]

Writes
  bag_clouds.npz        first /velodyne_points cloud of bag/2018-12-03-{1..6}.bag: xyzi float32 (N,4) + ring uint16
                        (the 32-byte PointCloud2 records are rebuilt from these in tests/conftest.py)
  fixtures.json         ilcc2/process_data/pointgrey_lidar_{1..6}.txt (35 x 3 each), pointgrey{1..6}.txt camera
                        corners, config/pointgrey.yaml board/intrinsics values, config/pointgrey.bin (4x4 extrinsic)
Nothing in tests/, smoke() or bench.py reads /root/reference at run time; they read these files.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from lidar_camera_calibration_b200 import rosbag_lite as rb  # noqa: E402


def main(ref="/root/reference"):
    clouds = {}
    for b in range(1, 7):
        pc = rb.first_cloud(f"{ref}/bag/2018-12-03-{b}.bag")
        assert pc.point_step == 32 and pc.offsets_xyzi() == (0, 4, 8, 16) and pc.fields["ring"][0] == 20
        clouds[f"xyzi_{b}"] = pc.packed_xyzi()
        clouds[f"ring_{b}"] = pc.field_u16("ring")
    np.savez_compressed(os.path.join(HERE, "bag_clouds.npz"), **clouds)

    fx = {"lidar_corners": {}, "cam_corners": {}}
    for b in range(1, 7):
        fx["lidar_corners"][str(b)] = np.loadtxt(f"{ref}/ilcc2/process_data/pointgrey_lidar_{b}.txt").tolist()
        fx["cam_corners"][str(b)] = np.loadtxt(f"{ref}/ilcc2/process_data/pointgrey{b}.txt").tolist()
    fx["lidar_corner_text_1"] = open(f"{ref}/ilcc2/process_data/pointgrey_lidar_1.txt").read()
    # raw text of every corner file, so that the tests can lay out a process_data/ directory for the host tools
    fx["lidar_corner_text"] = {str(b): open(f"{ref}/ilcc2/process_data/pointgrey_lidar_{b}.txt").read() for b in range(1, 7)}
    fx["cam_corner_text"] = {str(b): open(f"{ref}/ilcc2/process_data/pointgrey{b}.txt").read() for b in range(1, 7)}
    fx["extrinsic_bin_hex"] = open(f"{ref}/ilcc2/config/pointgrey.bin", "rb").read().hex()
    import cv2
    fs = cv2.FileStorage(f"{ref}/ilcc2/config/pointgrey.yaml", cv2.FILE_STORAGE_READ)
    fx["yaml"] = {
        "K": fs.getNode("K").mat().tolist(), "d": fs.getNode("d").mat().reshape(-1).tolist(),
        "width": int(fs.getNode("Camera.width").real()), "height": int(fs.getNode("Camera.height").real()),
        "grid_length": fs.getNode("grid_length").real(),
        "corner_in_x": int(fs.getNode("corner_in_x").real()), "corner_in_y": int(fs.getNode("corner_in_y").real()),
    }
    fx["yaml_text"] = open(f"{ref}/ilcc2/config/pointgrey.yaml").read()
    T = np.fromfile(f"{ref}/ilcc2/config/pointgrey.bin", dtype="<f8").reshape(4, 4).T  # column-major on disk
    fx["extrinsic_lidar2cam"] = T.tolist()
    # which keys the author pressed, inferred from the fixtures' corner order and colour phase (DESIGN.md "Oracle"):
    # 'd' once for bags 1,2,3,6; none for bag 5; fixture 4 belongs to a recording that is not bundled.
    fx["key_flags"] = {"1": 1, "2": 1, "3": 1, "5": 0, "6": 1}
    json.dump(fx, open(os.path.join(HERE, "fixtures.json"), "w"))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main(*sys.argv[1:])
