"""CPU: the oracle against everything the reference ships that pins this path.

The reference has no tests; its only pins are data fixtures: bag/2018-12-03-{1,2,3,5,6}.bag ->
ilcc2/process_data/pointgrey_lidar_{1,2,3,5,6}.txt (produced interactively by the author, click point not
recorded) and the README excerpt.  tests/golden/ holds the first cloud of each bag and the fixtures.
"""
import numpy as np
import pytest

# known-answer stage sizes of the oracle on the bundled bags (click = fixture centroid): ROI, clusters, chosen
# cluster, chessboard plane.  SURVEY.md Appendix B obtained the same numbers (+-2 plane points) independently.
STAGES = {1: (5430, 3, 2600, 2590), 2: (4036, 4, 1829, 1808), 3: (2950, 3, 1986, 1719),
          5: (2891, 3, 1861, 1821), 6: (3384, 2, 3066, 1377)}


def test_bags_reproduce_shipped_corner_files_in_order(oracle, bag_clouds, fixtures):
    """With the author's key presses ('d' for bags 1,2,3,6; none for 5) the oracle reproduces the five shipped
    process_data/pointgrey_lidar_<i>.txt files corner for corner, in file order, at the mm level."""
    worst = 0.0
    for b, flags in fixtures["key_flags"].items():
        fx = np.asarray(fixtures["lidar_corners"][b])
        st, corners, d = oracle.extract_frame(bag_clouds[int(b)], fx.mean(0), flags=flags)
        assert st == 0
        assert (d["n_roi"], d["n_clusters"], d["n_cluster_pts"], d["n_board"]) == STAGES[int(b)]
        err = np.linalg.norm(corners - fx, axis=1).max()
        worst = max(worst, err)
        assert err < 7e-3, (b, err)
        # fixture corners lie in the oracle's board plane x' = 0: the front half (ROI, cluster, plane, PCA) is
        # reproduced essentially exactly for bags 1,2,3,5; bag 6's cluster is attached to other structure, so its
        # plane depends on the unrecorded click
        T = d["pca"].reshape(4, 4).astype(np.float64)
        off_plane = np.abs(fx @ T[0, :3] + T[0, 3]).max()
        assert off_plane < (4e-3 if b == "6" else 2e-4), (b, off_plane)
    assert worst > 0  # not trivially identical: the click point and the Ceres build are unknown


def test_bag3_matches_to_a_fraction_of_a_millimetre(oracle, bag_clouds, fixtures):
    fx = np.asarray(fixtures["lidar_corners"]["3"])
    st, corners, d = oracle.extract_frame(bag_clouds[3], fx.mean(0), flags=1)
    assert st == 0 and np.linalg.norm(corners - fx, axis=1).max() < 5e-4


def test_auto_mode_picks_the_authors_colour_phase(oracle, bag_clouds, fixtures):
    from lidar_camera_calibration_b200.synth import corner_set_distance
    for b, flags in fixtures["key_flags"].items():
        fx = np.asarray(fixtures["lidar_corners"][b])
        st, corners, d = oracle.extract_frame(bag_clouds[int(b)], fx.mean(0), flags=oracle.FLAG_AUTO)
        assert st == 0 and d["hyp_chosen"] == flags
        assert corner_set_distance(corners, fx) < 7e-3
        # the wrong colour hypothesis is one square (0.15 m) off and costs several times more
        assert d["cost"][1 - flags] > 3 * d["cost"][flags]


def test_readme_excerpt(fixtures):
    """README.md:150-158 quotes the first/last lines of pointgrey_lidar_1.txt."""
    lines = fixtures["lidar_corner_text_1"].strip().splitlines()
    assert len(lines) == 35 and lines[0].split() == ["2.08881", "0.609289", "0.302605"]


def test_packed16_and_stride32_layouts_agree(oracle, bag_clouds, fixtures):
    fx = np.asarray(fixtures["lidar_corners"]["1"])
    rec = bag_clouds[1]
    a = oracle.extract_frame(rec, fx.mean(0), flags=1)
    b = oracle.extract_frame(np.ascontiguousarray(rec[:, [0, 1, 2, 4]]), fx.mean(0), flags=1)
    assert a[0] == b[0] == 0 and np.array_equal(a[1], b[1])


def test_eigenvector_signs_follow_the_fixture_axes(oracle, bag_clouds, fixtures):
    """The corner ORDER of a fixture reveals the in-plane PCA axes of the author's Eigen build:
    y' runs along fixture[7]-fixture[0], z' along fixture[1]-fixture[0] (LidarCornersEst.cpp:524-534)."""
    for b in fixtures["key_flags"]:
        fx = np.asarray(fixtures["lidar_corners"][b])
        st, _, d = oracle.extract_frame(bag_clouds[int(b)], fx.mean(0), flags=0)
        T = d["pca"].reshape(4, 4)
        e1 = (fx[7] - fx[0]) / np.linalg.norm(fx[7] - fx[0])
        e2 = (fx[1] - fx[0]) / np.linalg.norm(fx[1] - fx[0])
        assert e1 @ T[1, :3] > 0.95 and e2 @ T[2, :3] > 0.95


def test_synthetic_frames_recover_planted_corners(oracle):
    from lidar_camera_calibration_b200 import synth
    pts, clicks, gt = synth.make_batch(12, want_truth=True)
    st, corners = oracle.extract_batch(pts, clicks, flags=oracle.FLAG_AUTO, threads=4)
    assert (st == 0).all()
    err = np.array([synth.corner_set_distance(corners[f], gt[f]) for f in range(12)])
    assert np.median(err) < 4e-3 and (err < 1e-2).sum() >= 10   # method accuracy: mm on most frames


@pytest.mark.parametrize("threads", [1, 3])
def test_batch_equals_frame_by_frame(oracle, threads):
    from lidar_camera_calibration_b200 import synth
    pts, clicks = synth.make_batch(5, start=40)
    st, corners = oracle.extract_batch(pts, clicks, flags=oracle.FLAG_AUTO, threads=threads)
    for f in range(5):
        s1, c1, _ = oracle.extract_frame(pts[f], clicks[f], flags=oracle.FLAG_AUTO)
        assert s1 == st[f] and np.array_equal(c1, corners[f], equal_nan=True)
