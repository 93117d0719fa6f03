"""CPU tests of the corner-ordering / PnP oracle (oracle/ilcc2_pnp_oracle.c) against the reference's shipped artefacts.

The known answer is ilcc2/config/pointgrey.bin: the extrinsic the author's calib_lidar_cam run wrote from the six
shipped fixture pairs (calib_lidar_cam.cpp:101-165).  The oracle's restatement of Ceres' trust-region minimiser
reproduces it to the last bit or two, which also pins the trust-region machinery shared with the board-pose solver.
"""
import os

import numpy as np
import pytest

BW, BH = 7, 5  # corner_in_x, corner_in_y of config/pointgrey.yaml


def load_pairs(fixtures, oracle, tmp_path, bags=range(1, 7)):
    """The reference's data-ready loop (calib_lidar_cam.cpp:101-135) over the shipped corner files."""
    Tax = oracle.axis_roughly("pointgrey")
    P3, P2 = [], []
    for b in bags:
        lp, cp = tmp_path / f"pointgrey_lidar_{b}.txt", tmp_path / f"pointgrey{b}.txt"
        lp.write_text(fixtures["lidar_corner_text"][str(b)])
        cp.write_text(fixtures["cam_corner_text"][str(b)])
        p2 = oracle.read_cam_corners(str(cp), BW * BH, BH)
        p3 = oracle.read_lidar_corners(str(lp), BW * BH)
        p3 = p3 @ Tax[:3, :3].T + Tax[:3, 3]
        P3.append(oracle.order_lidar(p3, BW, BH))
        P2.append(oracle.order_cam(p2, BW, BH))
    K = np.asarray(fixtures["yaml"]["K"])
    return np.concatenate(P3), np.concatenate(P2), np.array([K[0, 0], K[0, 2], K[1, 1], K[1, 2]]), Tax


def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra @ Rb.T) - 1.0) / 2.0, -1.0, 1.0)))


def test_axis_roughly_matches_eigen_angle_axis_products(oracle):
    T = oracle.axis_roughly("pointgrey")
    c, s = np.cos(1.57), np.sin(1.57)
    Ry = np.array([[np.cos(-1.57), 0, np.sin(-1.57)], [0, 1, 0], [-np.sin(-1.57), 0, np.cos(-1.57)]])
    Rx = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    assert np.allclose(T[:3, :3], Ry @ Rx, atol=1e-15) and np.allclose(T[:3, 3], 0) and T[3, 3] == 1
    assert np.allclose(oracle.axis_roughly("left")[:3, :3], Rx, atol=1e-15)


def test_read_cam_corners_is_column_major_for_7_rows(fixtures, oracle, tmp_path):
    p = tmp_path / "c.txt"
    p.write_text(fixtures["cam_corner_text"]["1"])
    pts = oracle.read_cam_corners(str(p), 35, BH)
    raw = np.asarray(fixtures["cam_corners"]["1"])  # np.loadtxt: 7 rows of x then 7 rows of y, 5 columns
    assert pts.shape == (35, 2)
    for w in range(5):
        for h in range(7):
            assert pts[w * 7 + h, 0] == raw[h, w] and pts[w * 7 + h, 1] == raw[7 + h, w]


def test_order_checks_are_literal(oracle):
    rng = np.random.default_rng(0)
    g = np.stack(np.meshgrid(np.arange(BW), np.arange(BH)), -1).reshape(-1, 2).astype(np.float64)  # row-major, x fastest
    for flip_rows in (False, True):
        for flip_cols in (False, True):
            q = g.reshape(BH, BW, 2).copy()
            if flip_rows:
                q = q[::-1]
            if flip_cols:
                q = q[:, ::-1]
            q = q.reshape(-1, 2) + rng.normal(0, 0.01, (BW * BH, 2))
            out = oracle.order_cam(q, BW, BH)
            assert np.abs(out - g).max() < 0.1
            q3 = np.concatenate([q, rng.normal(5, 0.1, (BW * BH, 1))], 1)
            out3 = oracle.order_lidar(q3, BW, BH)
            assert np.abs(out3[:, :2] - g).max() < 0.1


def test_pnp_residual_jacobian_vs_finite_differences(oracle):
    rng = np.random.default_rng(1)
    cam = np.array([1061.4, 980.7, 1061.0, 601.7])
    for x in (np.zeros(6), np.array([1e-9, 0, 0, 0, 0, 0]), rng.normal(0, 0.2, 6)):  # small-angle branch and general
        p3 = np.array([0.3, -0.2, 2.5]) + rng.normal(0, 0.1, 3)
        p2 = np.array([1000.0, 600.0])
        res, J = oracle.pnp_residual(p3, p2, cam, x)
        Jn = np.zeros((2, 6))
        for k in range(6):
            d = np.zeros(6)
            d[k] = 1e-6
            Jn[:, k] = (oracle.pnp_residual(p3, p2, cam, x + d, False)[0] -
                        oracle.pnp_residual(p3, p2, cam, x - d, False)[0]) / 2e-6
        assert np.abs(J - Jn).max() < 1e-5 * max(1.0, np.abs(J).max())


def test_pnp_over_the_six_shipped_pairs_reproduces_pointgrey_bin(fixtures, oracle, tmp_path):
    P3, P2, cam, Tax = load_pairs(fixtures, oracle, tmp_path)
    r, t, sm = oracle.solve_pnp(P3, P2, cam)
    assert sm["n_pairs"] == 210 and sm["termination"] == 3 and sm["final_cost"] < sm["initial_cost"]
    T = oracle.compose_extrinsic(r, t, Tax)
    G = np.frombuffer(bytes.fromhex(fixtures["extrinsic_bin_hex"]), dtype="<f8").reshape(4, 4).T  # column-major file
    assert np.allclose(G, np.asarray(fixtures["extrinsic_lidar2cam"]))
    # north-star gate is 1e-3 rad / 1e-3 m; the restatement is ~13 orders of magnitude inside it
    assert rot_angle(T[:3, :3], G[:3, :3]) < 1e-9 and np.abs(T[:3, 3] - G[:3, 3]).max() < 1e-12
    assert np.abs(T - G).max() < 1e-12


def test_pnp_ignores_masked_pairs_and_recovers_a_planted_pose(oracle):
    rng = np.random.default_rng(2)
    cam = np.array([1061.4, 980.7, 1061.0, 601.7])
    r_true, t_true = np.array([0.02, -0.03, 0.01]), np.array([0.05, -0.07, 0.08])
    p3 = np.stack([rng.uniform(-1, 1, 400), rng.uniform(-0.6, 0.6, 400), rng.uniform(2, 4, 400)], 1)
    th = np.linalg.norm(r_true)
    k = r_true / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    q = p3 @ R.T + t_true
    p2 = np.stack([cam[0] * q[:, 0] / q[:, 2] + cam[1], cam[2] * q[:, 1] / q[:, 2] + cam[3]], 1)
    valid = np.ones(400, np.uint8)
    valid[::7] = 0
    p2[::7] += 300.0  # garbage in the masked pairs must not matter
    r, t, sm = oracle.solve_pnp(p3, p2, cam, valid=valid)
    assert sm["n_pairs"] == int(valid.sum())
    assert np.abs(r - r_true).max() < 1e-6 and np.abs(t - t_true).max() < 1e-6


def test_project_points_oracle_is_literal(fixtures, oracle, bag_clouds):
    """spaceToPlane + HSVtoRGB against a direct numpy transcription (ImageCornersEst.cpp:135-155, :372-428)."""
    rec = bag_clouds[1]
    T = np.asarray(fixtures["extrinsic_lidar2cam"])
    K = np.asarray(fixtures["yaml"]["K"])
    cam = np.array([K[0, 0], K[0, 2], K[1, 1], K[1, 2]])
    W, H = fixtures["yaml"]["width"], fixtures["yaml"]["height"]
    pix, rgbv = oracle.project_points(rec, T, cam, W, H)
    P = rec[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        u, v = cam[0] * (P[:, 0] / P[:, 2]) + cam[1], cam[2] * (P[:, 1] / P[:, 2]) + cam[3]
    valid = ~((P[:, 2] < 0) | (P[:, 2] > 5.0)) & (u > 0) & (u < W) & (v > 0) & (v < H)
    assert np.array_equal(rgbv[:, 3] == 1, valid) and 3000 < valid.sum() < 10000
    assert np.all(pix[~valid] == -1)
    assert np.abs(pix[valid, 0] - np.trunc(u[valid])).max() <= 1 and np.abs(pix[valid, 1] - np.trunc(v[valid])).max() <= 1
    for i in np.flatnonzero(valid)[::97]:
        h = int((float(rec[i, 4]) - 0.0) / 60.0 * 255)
        mx, adj = np.float32(255.0), np.float32(255.0) * np.float32(h % 60) / np.float32(60.0)
        r, g, b = {0: (mx, adj, 0), 1: (mx - adj, mx, 0), 2: (0, mx, adj), 3: (0, mx - adj, mx), 4: (adj, 0, mx)}.get(
            h // 60, (mx, 0, mx - adj))
        assert tuple(rgbv[i, :3]) == (int(r), int(g), int(b))
    img = np.random.default_rng(0).integers(0, 255, (H, W, 3), dtype=np.uint8)
    pix2, rgb2 = oracle.project_points(np.ascontiguousarray(rec[:, [0, 1, 2, 4]]), T, cam, W, H, image_bgr=img)
    assert np.array_equal(pix2, pix)
    assert np.array_equal(rgb2[valid][:, :3], img[pix[valid, 1], pix[valid, 0]][:, ::-1])
