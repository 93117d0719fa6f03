"""ctypes binding of the CPU oracle (oracle/libilcc2_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libilcc2_oracle.so")

FLAG_TOPLEFT_WHITE = 1
FLAG_INVERSE = 2
FLAG_AUTO = 4

STATUS_NAMES = {0: "OK", 1: "NO_POINTS_IN_ROI", 2: "NO_CLUSTER", 3: "PLANE_TOO_SMALL",
                4: "DEGENERATE_INTENSITY", 5: "NO_OPT_POINTS", 6: "RANSAC_FAILED"}


class Params(C.Structure):
    _fields_ = [
        ("roi_half_x", C.c_float), ("roi_half_y", C.c_float), ("roi_half_z", C.c_float),
        ("cluster_tolerance", C.c_float),
        ("min_cluster_size", C.c_int32), ("max_cluster_size", C.c_int32),
        ("ransac_threshold", C.c_double),
        ("ransac_max_iterations", C.c_int32),
        ("ransac_probability", C.c_double),
        ("hist_bins", C.c_int32),
        ("gray_rate", C.c_double), ("huber_delta", C.c_double), ("grid_length", C.c_double),
        ("cells_x", C.c_int32), ("cells_y", C.c_int32),
        ("min_board_points", C.c_int32), ("max_solver_iterations", C.c_int32),
    ]


class Diag(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("n_roi", C.c_int32), ("nn_index", C.c_int32), ("n_clusters", C.c_int32),
        ("click_in_cluster", C.c_int32), ("cluster_rank", C.c_int32), ("n_cluster_pts", C.c_int32),
        ("ransac_iterations", C.c_int32), ("n_ransac_inliers", C.c_int32), ("n_board", C.c_int32),
        ("n_opt", C.c_int32), ("hyp_chosen", C.c_int32),
        ("iters", (C.c_int32 * 2) * 2), ("term", (C.c_int32 * 2) * 2), ("n_evals", C.c_int32 * 2),
        ("plane", C.c_float * 4), ("pca", C.c_float * 16), ("eigvals", C.c_float * 3),
        ("hist_min", C.c_double), ("hist_max", C.c_double), ("hist_mean", C.c_double),
        ("r_low", C.c_double), ("r_high", C.c_double), ("gray", C.c_double * 2),
        ("theta_t", (C.c_double * 3) * 2), ("cost", C.c_double * 2), ("cost_a", C.c_double * 2),
    ]


def diag_to_dict(d: Diag) -> dict:
    out = {}
    for name, _ in Diag._fields_:
        v = getattr(d, name)
        out[name] = np.ctypeslib.as_array(v).copy() if hasattr(v, "_length_") else v
    return out


class PnpSummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("iterations", C.c_int32),
                ("termination", C.c_int32), ("n_evals", C.c_int32), ("n_pairs", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc). Returns the library path."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("ilcc2_oracle.c", "ilcc2_pnp_oracle.c", "ilcc2_oracle.h", "Makefile")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None
_lib_fast = None


def _cpu_tag() -> str:
    """Short hash of this machine's CPU model + ISA flags: the -march=native build is only valid where it was made."""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        keep = [ln for ln in txt.splitlines() if ln.startswith(("model name", "flags"))][:2]
    except OSError:
        keep = []
    return hashlib.sha1("\n".join(keep).encode()).hexdigest()[:10]


def lib_fast() -> C.CDLL:
    """The -O3 -march=native CPU-baseline build (bench.py's cpu_baseline / --impl reference legs only), compiled on
    first use on the machine that runs it.  Same entry points as lib(); never used as the parity checker."""
    global _lib_fast
    if _lib_fast is None:
        name = f"libilcc2_oracle_fast_{_cpu_tag()}.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path) or any(os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(path)
                                           for f in ("ilcc2_oracle.c", "ilcc2_pnp_oracle.c", "ilcc2_oracle.h")):
            subprocess.check_call(["make", "-C", _HERE, "-s", "fast", f"FAST={name}"])
        L = C.CDLL(path)
        _declare(L)
        _lib_fast = L
    return _lib_fast


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


def _declare(L) -> None:
    if True:
        L.ilcc2o_default_params.argtypes = [C.POINTER(Params)]
        L.ilcc2o_extract_frame.restype = C.c_int
        L.ilcc2o_extract_frame.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.POINTER(C.c_float), C.c_uint32, C.c_int32,
                                           C.POINTER(Params), C.POINTER(C.c_float), C.POINTER(Diag),
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        L.ilcc2o_extract_batch.restype = C.c_int
        L.ilcc2o_extract_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32,
                                           C.POINTER(Params), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ilcc2o_max_threads.restype = C.c_int
        L.ilcc2o_set_variant.argtypes = [C.c_uint32, C.c_int, C.c_int]
        L.ilcc2o_get_variant.restype = C.c_uint32
        L.ilcc2o_pcl_rnd.restype = C.c_uint32
        L.ilcc2o_pcl_rnd.argtypes = [C.c_int]
        L.ilcc2o_eigen3f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ilcc2o_gray_zone.restype = C.c_int
        L.ilcc2o_gray_zone.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Params), C.c_void_p]
        L.ilcc2o_residual.restype = C.c_double
        L.ilcc2o_residual.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                      C.c_int, C.c_int, C.c_void_p]
        L.ilcc2o_tr_boundary_2d.restype = C.c_int
        L.ilcc2o_tr_boundary_2d.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.ilcc2o_solve.restype = C.c_int
        L.ilcc2o_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                   C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ilcc2o_axis_roughly.restype = C.c_int
        L.ilcc2o_axis_roughly.argtypes = [C.c_char_p, C.c_void_p]
        L.ilcc2o_order_lidar.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ilcc2o_order_cam.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ilcc2o_pnp_residual.argtypes = [C.c_void_p] * 6
        L.ilcc2o_solve_pnp.restype = C.c_int
        L.ilcc2o_solve_pnp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_double,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(PnpSummary)]
        L.ilcc2o_compose_extrinsic.argtypes = [C.c_void_p] * 4
        L.ilcc2o_project_points.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                            C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]


def default_params(**overrides) -> Params:
    p = Params()
    lib().ilcc2o_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _layout(points: np.ndarray):
    """(n, point_step, off_x, off_y, off_z, off_i) for an (N,4) packed XYZI or (N,8) PointCloud2-like array."""
    assert points.dtype == np.float32 and points.ndim == 2 and points.flags.c_contiguous
    if points.shape[1] == 4:
        return points.shape[0], 16, 0, 4, 8, 12
    if points.shape[1] == 8:
        return points.shape[0], 32, 0, 4, 8, 16
    raise ValueError("points must be (N,4) packed XYZI or (N,8) 32-byte-stride records")


def extract_frame(points: np.ndarray, click, flags: int = 0, plane_index_delta: int = 0, params: Params | None = None,
                  want_indices: bool = False):
    """Run the whole path on one frame. Returns (status, corners[ncorn,3], diag dict[, roi, cluster, board])."""
    L = lib()
    p = params or default_params()
    n, step, ox, oy, oz, oi = _layout(points)
    ncorn = (p.cells_x - 1) * (p.cells_y - 1)
    corners = np.empty((ncorn, 3), np.float32)
    d = Diag()
    ck = (C.c_float * 3)(*[float(v) for v in click])
    roi = clu = brd = None
    if want_indices:
        roi = np.empty(max(n, 1), np.int32)
        clu = np.empty(max(n, 1), np.int32)
        brd = np.empty(max(n, 1), np.int32)
    st = L.ilcc2o_extract_frame(points.ctypes.data, n, step, ox, oy, oz, oi, ck, flags, plane_index_delta,
                                C.byref(p), corners.ctypes.data_as(C.POINTER(C.c_float)), C.byref(d),
                                roi.ctypes.data if want_indices else None,
                                clu.ctypes.data if want_indices else None,
                                brd.ctypes.data if want_indices else None)
    dd = diag_to_dict(d)
    if want_indices:
        return st, corners, dd, roi[:d.n_roi].copy(), clu[:d.n_cluster_pts].copy(), brd[:d.n_board].copy()
    return st, corners, dd


def extract_batch(points: np.ndarray, clicks: np.ndarray, flags=FLAG_AUTO, params: Params | None = None,
                  threads: int = 1, want_diag: bool = False, fast: bool = False):
    """points: (F, N, 4|8) float32; clicks (F,3). Returns (status[F], corners[F,ncorn,3][, diags]).
    fast: the -O3 -march=native baseline build (timing only) instead of the parity build."""
    L = lib_fast() if fast else lib()
    p = params or default_params()
    assert points.ndim == 3 and points.dtype == np.float32 and points.flags.c_contiguous
    F, N, w = points.shape
    _, step, ox, oy, oz, oi = _layout(points.reshape(F * N, w))
    clicks = np.ascontiguousarray(clicks, np.float32)
    ncorn = (p.cells_x - 1) * (p.cells_y - 1)
    corners = np.empty((F, ncorn, 3), np.float32)
    status = np.empty(F, np.int32)
    diags = (Diag * F)() if want_diag else None
    if np.isscalar(flags):
        fl_ptr, dflt = None, int(flags)
    else:
        fl = np.ascontiguousarray(flags, np.uint32)
        fl_ptr, dflt = fl.ctypes.data, 0
    L.ilcc2o_extract_batch(points.ctypes.data, F, N, step, ox, oy, oz, oi, clicks.ctypes.data, fl_ptr, dflt,
                           C.byref(p), threads, corners.ctypes.data, status.ctypes.data,
                           C.cast(diags, C.c_void_p) if want_diag else None)
    if want_diag:
        return status, corners, [diag_to_dict(d) for d in diags]
    return status, corners


def max_threads() -> int:
    return lib().ilcc2o_max_threads()


def pcl_rnd(k: int) -> int:
    return int(lib().ilcc2o_pcl_rnd(k))


def eigen3f(cov: np.ndarray):
    cov = np.ascontiguousarray(cov, np.float32).reshape(9)
    ev = np.empty(3, np.float32)
    V = np.empty(9, np.float32)
    lib().ilcc2o_eigen3f(cov.ctypes.data, ev.ctypes.data, V.ctypes.data)
    return ev, V.reshape(3, 3).T.copy()  # columns = eigenvectors


def gray_zone(intensity: np.ndarray, params: Params | None = None):
    p = params or default_params()
    inten = np.ascontiguousarray(intensity, np.float32)
    out = np.zeros(7)
    st = lib().ilcc2o_gray_zone(inten.ctypes.data, inten.size, C.byref(p), out.ctypes.data)
    return st, dict(min=out[0], max=out[1], mean=out[2], r_low=out[3], r_high=out[4], gray=(out[5], out[6]))


def residual(y, z, white, x, W=6, H=8, g=0.15, topleft_white=False, use_oob=True):
    xx = np.ascontiguousarray(x, np.float64)
    jac = np.zeros(3)
    r = lib().ilcc2o_residual(float(y), float(z), int(white), xx.ctypes.data, W, H, g, int(topleft_white),
                              int(use_oob), jac.ctypes.data)
    return r, jac


def tr_boundary_2d(B, g, radius):
    B = np.ascontiguousarray(B, np.float64)
    g = np.ascontiguousarray(g, np.float64)
    x = np.zeros(2)
    ok = lib().ilcc2o_tr_boundary_2d(B.ctypes.data, g.ctypes.data, float(radius), x.ctypes.data)
    return bool(ok), x


def solve(yz: np.ndarray, white: np.ndarray, x0, W=6, H=8, g=0.15, topleft_white=False, use_oob=True,
          huber=0.1, max_iter=50):
    yz = np.ascontiguousarray(yz, np.float64)
    white = np.ascontiguousarray(white, np.uint8)
    x = np.array(x0, np.float64)
    cost = C.c_double()
    it = C.c_int32()
    ne = C.c_int32()
    term = lib().ilcc2o_solve(yz.ctypes.data, white.ctypes.data, white.size, W, H, g, int(topleft_white),
                              int(use_oob), huber, max_iter, x.ctypes.data, C.byref(cost), C.byref(it), C.byref(ne))
    return term, x, cost.value, it.value, ne.value


# ---- corner ordering + LiDAR->camera PnP (ilcc2_pnp_oracle.c) ----
def axis_roughly(cam_name: str) -> np.ndarray:
    """get_lidar2cam_axis_roughly (calib_lidar_cam.cpp:50-70) as a 4x4."""
    T = np.zeros((4, 4))
    lib().ilcc2o_axis_roughly(cam_name.encode(), T.ctypes.data)
    return T


def order_lidar(pts3: np.ndarray, bw: int, bh: int) -> np.ndarray:
    p = np.ascontiguousarray(pts3, np.float64).copy()
    lib().ilcc2o_order_lidar(p.ctypes.data, bw, bh)
    return p


def order_cam(pts2: np.ndarray, bw: int, bh: int) -> np.ndarray:
    p = np.ascontiguousarray(pts2, np.float64).copy()
    lib().ilcc2o_order_cam(p.ctypes.data, bw, bh)
    return p


def pnp_residual(p3, p2, cam, x, want_jac=True):
    p3, p2, cam, x = (np.ascontiguousarray(a, np.float64) for a in (p3, p2, cam, x))
    res, jac = np.zeros(2), np.zeros((2, 6))
    lib().ilcc2o_pnp_residual(p3.ctypes.data, p2.ctypes.data, cam.ctypes.data, x.ctypes.data, res.ctypes.data,
                              jac.ctypes.data if want_jac else None)
    return res, jac


def solve_pnp(p3, p2, cam, r0=(0, 0, 0), t0=(0, 0, 0), valid=None, huber=0.1, max_iter=50):
    """Optimization::solvePose3d2dError (Optimization.cpp:13-91). Returns (r, t, summary dict)."""
    p3, p2, cam = (np.ascontiguousarray(a, np.float64) for a in (p3, p2, cam))
    r, t = np.array(r0, np.float64), np.array(t0, np.float64)
    v = None if valid is None else np.ascontiguousarray(valid, np.uint8)
    sm = PnpSummary()
    lib().ilcc2o_solve_pnp(p3.ctypes.data, p2.ctypes.data, None if v is None else v.ctypes.data, p3.shape[0],
                           cam.ctypes.data, huber, max_iter, r.ctypes.data, t.ctypes.data, C.byref(sm))
    return r, t, {k: getattr(sm, k) for k, _ in PnpSummary._fields_}


def compose_extrinsic(r, t, T_axis) -> np.ndarray:
    r, t, T_axis = (np.ascontiguousarray(a, np.float64) for a in (r, t, T_axis))
    T = np.zeros((4, 4))
    lib().ilcc2o_compose_extrinsic(r.ctypes.data, t.ctypes.data, T_axis.ctypes.data, T.ctypes.data)
    return T


def read_cam_corners(path: str, num: int, board_h: int) -> np.ndarray:
    """ImageCornersEst::read_cam_corners (ImageCornersEst.cpp:213-279): lines of x until >= num values, then lines
    of y into the same [row][col]; emitted column-major when the row count differs from the board height."""
    lines = [ln.split() for ln in open(path).read().splitlines()]
    rows, counter, i = [], 0, 0
    while i < len(lines):
        vals = [float(v) for v in lines[i]]
        i += 1
        rows.append([[v, 0.0] for v in vals])
        counter += len(vals)
        if counter >= num:
            break
    k = 0
    while i < len(lines) and k < len(rows):
        for j, v in enumerate(lines[i][:len(rows[k])]):
            rows[k][j][1] = float(v)
        i += 1
        k += 1
    out = []
    if len(rows) != board_h:
        for w in range(len(rows[0])):
            for h in range(len(rows)):
                out.append(rows[h][w])
    else:
        for h in range(len(rows)):
            out.extend(rows[h])
    return np.asarray(out, np.float64)


def read_lidar_corners(path: str, num: int) -> np.ndarray:
    """ImageCornersEst::read_lidar_corners (:281-299): whitespace-separated floats parsed as float32."""
    v = np.array(open(path).read().split()[: 3 * num], dtype=np.float32).reshape(-1, 3)
    return v.astype(np.float64)


def project_points(points: np.ndarray, T_lidar2cam, camera, width, height, distance_valid=5.0, inten_low=0.0,
                   inten_high=60.0, image_bgr=None):
    """spaceToPlane + colouring (pcd2image.cpp:59-84 / rgblidar.cpp:53-73). Returns (pix (n,2) int32, rgbv (n,4) uint8)."""
    n, step, ox, oy, oz, oi = _layout(points)
    T, cam = np.ascontiguousarray(T_lidar2cam, np.float64), np.ascontiguousarray(camera, np.float64)
    pix, rgbv = np.empty((n, 2), np.int32), np.empty((n, 4), np.uint8)
    img = None if image_bgr is None else np.ascontiguousarray(image_bgr, np.uint8)
    lib().ilcc2o_project_points(points.ctypes.data, n, step, ox, oy, oz, oi, T.ctypes.data, cam.ctypes.data, width, height,
                                distance_valid, inten_low, inten_high, None if img is None else img.ctypes.data,
                                pix.ctypes.data, rgbv.ctypes.data)
    return pix, rgbv


# ---- parity of a whole batch against this oracle (bench.py's cpu_baseline leg and tests/) ----
def parity_report(points: np.ndarray, clicks: np.ndarray, status_gpu: np.ndarray, corners_gpu: np.ndarray,
                  diags_gpu=None, flags: int = FLAG_AUTO, params: Params | None = None, threads: int = 1,
                  tol_m: float = 1e-3) -> dict:
    """Runs the oracle on every frame of the batch and compares: per-frame status, corner distance (same lattice
    order, so index for index), and -- when the product's ilcc2_diag records are given -- solver iteration counts /
    termination codes / chosen hypothesis.  Returns the `parity` block of bench.py's JSON line."""
    t0 = __import__("time").time()
    st, corners, diags = extract_batch(points, clicks, flags=flags, params=params, threads=threads, want_diag=True)
    dt = __import__("time").time() - t0
    status_gpu = np.asarray(status_gpu)
    corners_gpu = np.asarray(corners_gpu, np.float64)
    F = int(st.size)
    same_status = st == status_gpu
    ok = (st == 0) & (status_gpu == 0)
    err = np.zeros(F)
    if ok.any():
        err[ok] = np.linalg.norm(corners_gpu[ok] - corners[ok].astype(np.float64), axis=2).max(axis=1)
    out = {"frames": F, "frames_ok": int(ok.sum()), "status_equal": bool(same_status.all()),
           "status_mismatch_frames": np.flatnonzero(~same_status)[:16].tolist(),
           "max_corner_err_m": float(err.max()) if F else 0.0,
           "frac_within_1e-3": float((err[ok] <= tol_m).mean()) if ok.any() else None,
           "frames_beyond_tol": np.flatnonzero(err > tol_m)[:16].tolist(),
           "p99_corner_err_m": float(np.percentile(err[ok], 99)) if ok.any() else None,
           "oracle_s": dt, "oracle_threads": threads}
    if diags_gpu is not None:
        it_eq = hyp_eq = term_eq = 0
        worst_theta = 0.0
        for f in np.flatnonzero(ok):
            a, b = diags[f], diags_gpu[f]
            it_eq += int(np.array_equal(a["iters"], b["iters"]))
            term_eq += int(np.array_equal(a["term"], b["term"]))
            hyp_eq += int(a["hyp_chosen"] == b["hyp_chosen"])
            worst_theta = max(worst_theta, float(np.abs(np.asarray(a["theta_t"]) - np.asarray(b["theta_t"])).max()))
        n = max(int(ok.sum()), 1)
        out.update({"iters_equal_frac": it_eq / n, "term_equal_frac": term_eq / n, "hyp_equal_frac": hyp_eq / n,
                    "max_theta_t_diff": worst_theta})
    return out
