"""Host-side mirror of the reference's interface for the LiDAR corner path, over the C ABI.

Two layers:
  * CornerExtractor      -- thin wrapper of ilcc2_create / ilcc2_extract_corners for batches (numpy host arrays or
                            torch CUDA tensors).  This is what bench.py and the parity tests call.
    PipelinedExtractor   -- several of them on streams of their own taking consecutive batches in turn (throughput mode).
  * LidarCornersEst      -- same method names, argument meaning and return values as the reference class
                            (ilcc2/include/ilcc2/LidarCornersEst.h:40-67) as driven by ilcc2/test/get_lidar_corners.cpp:183-198:
                            set_chessboard_param(yaml) / setROI(cloud, point) / EuclideanCluster() / PCA() /
                            get_corners(corners); the GUI key presses become constructor/call arguments.
All arithmetic happens in libilcc2_b200.so on the GPU; nothing here computes corners.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Diag, Params

FLAG_TOPLEFT_WHITE = 1
FLAG_INVERSE = 2
FLAG_AUTO = 4
MEM_HOST, MEM_DEVICE = 0, 1
STATUS_NAMES = {0: "OK", 1: "NO_POINTS_IN_ROI", 2: "NO_CLUSTER", 3: "PLANE_TOO_SMALL",
                4: "DEGENERATE_INTENSITY", 5: "NO_OPT_POINTS", 6: "RANSAC_FAILED", 7: "ROI_CAPACITY"}
GATHER_PEER, GATHER_MULTICAST = 0, 1
STAGES = ("roi_filter", "cluster", "plane_pca", "board_fit")


class Ilcc2Error(RuntimeError):
    pass


def default_params(**overrides) -> Params:
    p = Params()
    _lib.load().ilcc2_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def set_chessboard_param(p: Params, cam_yaml: str) -> bool:
    """LidarCornersEst::set_chessboard_param (LidarCornersEst.cpp:20-46): grid_length and sorted (cells_x, cells_y)."""
    return _lib.load().ilcc2_set_chessboard_param(C.byref(p), cam_yaml.encode()) == 0


def diag_to_dict(d: Diag) -> dict:
    out = {}
    for name, _ in Diag._fields_:
        v = getattr(d, name)
        out[name] = np.ctypeslib.as_array(v).copy() if hasattr(v, "_length_") else v
    return out


def _layout(width: int):
    if width == 4:
        return 16, 0, 4, 8, 12
    if width == 8:
        return 32, 0, 4, 8, 16
    raise ValueError("points must have 4 (packed XYZI) or 8 (32-byte PointCloud2 record) float32 columns")


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class CornerExtractor:
    """Owns one ilcc2_handle (device scratch for max_frames x max_points_per_frame)."""

    def __init__(self, max_frames: int, max_points_per_frame: int, params: Params | None = None, device: int = 0,
                 roi_capacity: int = 0, overflow_frames: int = 0):
        """roi_capacity / overflow_frames: ilcc2_create_sized (0 = worst-case sizing, as ilcc2_create)."""
        self._L = _lib.load()
        self.params = params or default_params()
        self.device = device
        self.max_frames = max_frames
        self.max_points = max_points_per_frame
        self.n_corners = (self.params.cells_x - 1) * (self.params.cells_y - 1)
        h = C.c_void_p()
        rc = self._L.ilcc2_create_sized(C.byref(self.params), device, max_frames, max_points_per_frame,
                                        int(roi_capacity), int(overflow_frames), C.byref(h))
        if rc != 0:
            raise Ilcc2Error(f"ilcc2_create failed ({rc}): {self._L.ilcc2_last_error(None).decode()}")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.ilcc2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise Ilcc2Error(f"ilcc2 error {rc}: {self._L.ilcc2_last_error(self._h).decode()}")

    # ---- batch call, host buffers (numpy) ----
    def extract(self, points: np.ndarray, clicks: np.ndarray, flags=FLAG_AUTO, plane_index_delta=None,
                frame_offsets=None, want_diag: bool = False):
        """points: (F, N, 4|8) float32, or (total, 4|8) with frame_offsets (F+1 int64). clicks (F,3).
        flags: scalar applied to all frames, or (F,) uint8. Returns (status[F], corners[F,ncorn,3][, diags])."""
        assert points.dtype == np.float32 and points.flags.c_contiguous
        if frame_offsets is None:
            assert points.ndim == 3
            F, N, wdt = points.shape
            fo_ptr = None
        else:
            assert points.ndim == 2
            frame_offsets = np.ascontiguousarray(frame_offsets, np.int64)
            F, N, wdt = frame_offsets.size - 1, 0, points.shape[1]
            fo_ptr = frame_offsets.ctypes.data
        step, ox, oy, oz, oi = _layout(wdt)
        clicks = np.ascontiguousarray(clicks, np.float32).reshape(F, 3)
        corners = np.empty((F, self.n_corners, 3), np.float32)
        status = np.empty(F, np.int32)
        diags = (Diag * F)() if want_diag else None
        fl_ptr, dflt = None, 0
        if np.isscalar(flags):
            dflt = int(flags)
        else:
            fl = np.ascontiguousarray(flags, np.uint8)
            assert fl.size == F
            fl_ptr = fl.ctypes.data
        pd_ptr = None
        if plane_index_delta is not None:
            pd = np.ascontiguousarray(plane_index_delta, np.int8)
            assert pd.size == F
            pd_ptr = pd.ctypes.data
        rc = self._L.ilcc2_extract_corners(self._h, points.ctypes.data, F, N, fo_ptr, step, ox, oy, oz, oi, MEM_HOST,
                                           clicks.ctypes.data, fl_ptr, dflt, pd_ptr, corners.ctypes.data,
                                           status.ctypes.data, C.cast(diags, C.c_void_p) if want_diag else None, None)
        self._check(rc)
        if want_diag:
            return status, corners, [diag_to_dict(d) for d in diags]
        return status, corners

    # ---- batch call, raw pointers (pinned host or device), asynchronous for device memory ----
    def extract_raw(self, points_ptr: int, n_frames: int, points_per_frame: int, width: int, clicks_ptr: int,
                    corners_ptr: int, status_ptr: int, mem_kind: int, flags: int = FLAG_AUTO, stream: int = 0,
                    flags_ptr: int = 0, diag_ptr: int = 0):
        step, ox, oy, oz, oi = _layout(width)
        rc = self._L.ilcc2_extract_corners(self._h, points_ptr, n_frames, points_per_frame, None, step, ox, oy, oz, oi,
                                           mem_kind, clicks_ptr, flags_ptr or None, flags, None, corners_ptr,
                                           status_ptr, diag_ptr or None, stream or None)
        self._check(rc)

    def extract_torch(self, points, clicks, corners=None, status=None, flags: int = FLAG_AUTO, stream=None):
        """points: CUDA float32 tensor (F, N, 4|8); clicks CUDA (F,3). Asynchronous on `stream` (default: current)."""
        import torch
        assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous() and points.dim() == 3
        F, N, wdt = points.shape
        if corners is None:
            corners = torch.empty((F, self.n_corners, 3), dtype=torch.float32, device=points.device)
        if status is None:
            status = torch.empty((F,), dtype=torch.int32, device=points.device)
        st = stream if stream is not None else torch.cuda.current_stream(points.device)
        self.extract_raw(points.data_ptr(), F, N, wdt, clicks.data_ptr(), corners.data_ptr(), status.data_ptr(),
                         MEM_DEVICE, flags, st.cuda_stream)
        return status, corners

    # ---- introspection of the last call ----
    def set_gather_targets(self, corner_ptrs, status_ptrs, slot_base: int, kind: int = GATHER_PEER):
        """ilcc2_set_gather_targets: K5 also stores every frame's corners / status into these device-accessible
        buffers (kind GATHER_PEER: peer mappings of the ranks; GATHER_MULTICAST: one multicast mapping, stored to with
        multimem.st) at frame slot slot_base + f."""
        n = len(corner_ptrs)
        assert n == len(status_ptrs)
        ca = (C.c_void_p * max(n, 1))(*[int(p) for p in corner_ptrs])
        sa = (C.c_void_p * max(n, 1))(*[int(p) for p in status_ptrs])
        self._check(self._L.ilcc2_set_gather_targets(self._h, n, ca, sa, int(slot_base), int(kind)))

    def fetch_indices(self, frame: int, which: int) -> np.ndarray:
        """which: 0 m_cloud_ROI, 1 chosen cluster, 2 m_cloud_chessboard -- original point indices of `frame`."""
        buf = np.empty(self.max_points, np.int32)
        n = self._L.ilcc2_fetch_indices(self._h, frame, which, buf.ctypes.data, buf.size)
        if n < 0:
            raise Ilcc2Error(f"ilcc2_fetch_indices failed ({n})")
        return buf[:n].copy()

    def fetch_board_points(self, frame: int):
        yz = np.empty((self.max_points, 2), np.float32)
        wh = np.empty(self.max_points, np.uint8)
        n = self._L.ilcc2_fetch_board_points(self._h, frame, yz.ctypes.data, wh.ctypes.data, self.max_points)
        if n < 0:
            raise Ilcc2Error(f"ilcc2_fetch_board_points failed ({n})")
        return yz[:n].copy(), wh[:n].copy()

    def set_lanes(self, lanes: int):
        """ilcc2_set_lanes: this handle is one of `lanes` handles that take consecutive device-resident batches in turn."""
        self._check(self._L.ilcc2_set_lanes(self._h, int(lanes)))

    def set_profiling(self, enabled: bool):
        self._check(self._L.ilcc2_set_profiling(self._h, int(enabled)))

    def stage_times(self) -> dict:
        ms = (C.c_float * 8)()
        names = (C.c_char_p * 8)()
        n = self._L.ilcc2_stage_times(self._h, ms, names, 8)
        if n < 0:
            raise Ilcc2Error("ilcc2_stage_times failed")
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    @property
    def launch_count(self) -> int:
        return int(self._L.ilcc2_launch_count(self._h))


class PipelinedExtractor:
    """Throughput mode for a STREAM of device-resident batches: `lanes` handles (ilcc2_create_sized each), every one
    on a CUDA stream of its own, take the batches in turn.  ilcc2_extract_corners is asynchronous for device memory,
    so the kernels of batch k+1 (K1-K4, then the first wave of K5) run on the SMs that the tail of batch k's K5 leaves
    idle -- K5 is a chain of ~100 dependent evaluations per solve, its last CTAs keep a few SMs busy for the length of
    a whole solve.  Every batch goes through exactly the same launches as with one handle (results are identical); only
    their placement in time changes.  Measured per 1024 frames: one handle 2.02 ms, three lanes 1.63 ms (K5 then runs one
    warp per solve: ilcc2_set_lanes), four lanes 1.62.

    submit() makes the lane's stream wait for the caller's current stream (inputs are ready), enqueues the batch, and
    returns the lane; join() makes a stream wait for everything submitted so far.  Output buffers of a lane must not be
    reused before that lane's next submit() has been joined (bench.py keeps one pair per lane)."""

    def __init__(self, lanes: int, max_frames: int, max_points_per_frame: int, device: int = 0, **kw):
        import torch
        self.device = torch.device("cuda", device)
        self.extractors = [CornerExtractor(max_frames, max_points_per_frame, device=device, **kw) for _ in range(lanes)]
        for e in self.extractors:
            e.set_lanes(lanes)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(lanes)]
        self._ready = [torch.cuda.Event() for _ in range(lanes)]
        self._done = [torch.cuda.Event() for _ in range(lanes)]
        self._next = 0
        self.active = lanes   # lanes that take batches (set to 1 for "one batch in flight" on the same object)
        self.n_corners = self.extractors[0].n_corners

    @property
    def lanes(self) -> int:
        return len(self.extractors)

    def submit(self, points, clicks, corners, status, flags: int = FLAG_AUTO, after=None) -> int:
        """Enqueue one batch on the next lane. `after(lane)` (optional) is called with the lane's stream current, right
        behind the batch (e.g. the exchange step of a multi-GPU run). `corners` / `status` may be lists with one
        tensor per lane. Returns the lane index."""
        import torch
        j = self._next % self.active
        self._next = (j + 1) % self.active
        if isinstance(corners, (list, tuple)):
            corners = corners[j]
        if isinstance(status, (list, tuple)):
            status = status[j]
        st = self.streams[j]
        self._ready[j].record(torch.cuda.current_stream(self.device))
        st.wait_event(self._ready[j])
        self.extractors[j].extract_torch(points, clicks, corners, status, flags=flags, stream=st)
        if after is not None:
            with torch.cuda.stream(st):
                after(j)
        self._done[j].record(st)
        return j

    def join(self, stream=None):
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        for e in self._done:
            s.wait_event(e)

    # ---- host buffers: ilcc2_extract_corners(ILCC2_MEM_HOST) blocks until the results are in the caller's buffers, so
    # every lane gets a host thread; batch k+1's H2D copies then run while batch k's last chunk is still in K5 and its
    # corners travel back (measured, 1024 frames from page-locked memory: 9.32 -> 8.66 ms = the PCIe rate of 472 MB)
    def _host_worker(self, j):
        import torch
        torch.cuda.set_device(self.device)
        while True:
            item = self._hq[j].get()
            try:
                if item is None:
                    return
                try:
                    self.extractors[j].extract_raw(*item, self.streams[j].cuda_stream)
                except Exception as e:  # noqa: BLE001 -- reported by wait_lane() in the submitting thread
                    self._herr[j] = e
            finally:
                self._hq[j].task_done()

    def submit_host(self, points_ptr: int, n_frames: int, points_per_frame: int, width: int, clicks_ptr: int,
                    corners_ptr: int, status_ptr: int, flags: int = FLAG_AUTO) -> int:
        """Hand one host batch to the next lane's thread (arguments as CornerExtractor.extract_raw with MEM_HOST); returns
        the lane. The buffers belong to the lane until wait_lane(lane) / wait_all()."""
        import queue
        import threading
        if not hasattr(self, "_hq"):
            n = len(self.extractors)
            self._hq = [queue.Queue() for _ in range(n)]
            self._herr = [None] * n
            self._ht = [threading.Thread(target=self._host_worker, args=(j,), daemon=True) for j in range(n)]
            for t in self._ht:
                t.start()
        j = self._next % self.active
        self._next = (j + 1) % self.active
        self.wait_lane(j)
        self._hq[j].put((points_ptr, n_frames, points_per_frame, width, clicks_ptr, corners_ptr, status_ptr, MEM_HOST, flags))
        return j

    def wait_lane(self, j: int):
        if hasattr(self, "_hq"):
            self._hq[j].join()
            if self._herr[j] is not None:
                e, self._herr[j] = self._herr[j], None
                raise e

    def wait_all(self):
        for j in range(len(self.extractors)):
            self.wait_lane(j)

    @property
    def next_lane(self) -> int:
        return self._next % self.active

    @property
    def launch_count(self) -> int:
        return sum(e.launch_count for e in self.extractors)

    def close(self):
        if hasattr(self, "_hq"):
            for q in self._hq:
                q.put(None)
            for t in self._ht:
                t.join(timeout=10)
            del self._hq
        for e in self.extractors:
            e.close()


def save_corners2txt(corners: np.ndarray, path: str):
    """get_lidar_corners.cpp:27-36: 'x y z' per line, 6 significant digits."""
    c = np.ascontiguousarray(corners, np.float32).reshape(-1, 3)
    rc = _lib.load().ilcc2_save_corners2txt(c.ctypes.data, c.shape[0], path.encode())
    if rc != 0:
        raise Ilcc2Error(f"ilcc2_save_corners2txt({path}) failed ({rc})")


def shard_range(n_frames: int, world_size: int, rank: int):
    b, c = C.c_int(), C.c_int()
    _lib.load().ilcc2_shard_range(n_frames, world_size, rank, C.byref(b), C.byref(c))
    return b.value, c.value


class LidarCornersEst:
    """Drop-in shaped like the reference class; one frame at a time (what the ROS node does).

    Key presses of the two PCL viewer windows become attributes set before the call:
      plane_index_delta  net 'w'/'s' presses in EuclideanCluster()   (Visualization.cpp:22-29)
      topleft_white      True = one press of 'd' in get_corners()     (Visualization.cpp:45-48)
      inverse            True = one press of 'a'                      (Visualization.cpp:49-52)
      auto               True = unattended colour-phase selection (not in the reference)
    """

    def __init__(self, device: int = 0, max_points: int = 150000):
        self.params = default_params()
        self._device = device
        self._max_points = max_points
        self._ex: CornerExtractor | None = None
        self.plane_index_delta = 0
        self.topleft_white = False
        self.inverse = False
        self.auto = False
        self.m_click_point = None
        self.m_cloud_ROI = np.empty((0, 4), np.float32)
        self.m_cloud_chessboard = np.empty((0, 4), np.float32)
        self.m_cloud_PCA = np.empty((0, 4), np.float32)
        self.m_cloud_corners = np.empty((0, 3), np.float32)
        self.pca_matrix = np.eye(4, dtype=np.float32)
        self.m_gray_zone = np.zeros(2)
        self.diag: dict | None = None
        self._cloud = None
        self._status = None
        self._corners = None

    # LidarCornersEst.cpp:20-46
    def set_chessboard_param(self, cam_yaml: str) -> bool:
        ok = set_chessboard_param(self.params, cam_yaml)
        self._ex = None
        return ok

    @property
    def m_grid_length(self):
        return self.params.grid_length

    @property
    def m_grid_size(self):
        return (self.params.cells_x, self.params.cells_y)

    # LidarCornersEst.cpp:48-70.  cloud: (N,4) packed XYZI or (N,8) 32-byte records, float32
    def setROI(self, cloud: np.ndarray, point) -> None:
        self._cloud = np.ascontiguousarray(cloud, np.float32)
        self.m_click_point = np.asarray(point, np.float32).reshape(3)
        self._status = None

    def _xyzi(self, idx):
        c = self._cloud
        return c[idx][:, [0, 1, 2, 3]] if c.shape[1] == 4 else c[idx][:, [0, 1, 2, 4]]

    def _run(self):
        if self._ex is None:
            self._ex = CornerExtractor(1, self._max_points, self.params, self._device)
        flags = (FLAG_AUTO if self.auto else 0) | (FLAG_TOPLEFT_WHITE if self.topleft_white else 0) | \
                (FLAG_INVERSE if self.inverse else 0)
        st, corners, diags = self._ex.extract(self._cloud[None], self.m_click_point[None], flags=flags,
                                              plane_index_delta=[self.plane_index_delta], want_diag=True)
        self._status, self._corners, self.diag = int(st[0]), corners[0], diags[0]
        self.m_cloud_ROI = self._xyzi(self._ex.fetch_indices(0, 0))
        self.m_cloud_chessboard = self._xyzi(self._ex.fetch_indices(0, 2))
        self.pca_matrix = self.diag["pca"].reshape(4, 4)
        self.m_gray_zone = self.diag["gray"]

    # LidarCornersEst.cpp:124-186: True when a chessboard plane was found ('o' pressed)
    def EuclideanCluster(self) -> bool:
        if self._cloud is None:
            raise Ilcc2Error("setROI() must be called first")
        self._run()
        return self._status not in (1, 2, 3, 6)

    # LidarCornersEst.cpp:366-372
    def PCA(self) -> None:
        if self._status is None:
            self._run()
        if len(self.m_cloud_chessboard):
            T = self.pca_matrix
            xyz = self.m_cloud_chessboard[:, :3] @ T[:3, :3].T + T[:3, 3]
            self.m_cloud_PCA = np.concatenate([xyz, self.m_cloud_chessboard[:, 3:4]], 1).astype(np.float32)

    # LidarCornersEst.cpp:374-450: fills `corners` (list) with (x,y,z) tuples; True when accepted ('k')
    def get_corners(self, corners: list) -> bool:
        if self._status is None:
            self._run()
        if self._status != 0:
            return False
        self.m_cloud_corners = self._corners.copy()
        corners.extend([tuple(float(v) for v in c) for c in self._corners])
        return True


# ---------------------------------------------------------------------------------------------------------------
# What follows the corner path (calib_lidar_cam.cpp:101-165): corner ordering + LiDAR->camera PnP, on the GPU.
# ---------------------------------------------------------------------------------------------------------------
def axis_roughly(camera_name: str) -> np.ndarray:
    """get_lidar2cam_axis_roughly (calib_lidar_cam.cpp:50-70) as a row-major 4x4."""
    T = np.zeros((4, 4))
    if _lib.load().ilcc2_axis_roughly(camera_name.encode(), T.ctypes.data) != 0:
        raise Ilcc2Error(f"unknown camera name {camera_name!r}")
    return T


def compose_extrinsic(r, t, T_axis) -> np.ndarray:
    r, t, T_axis = (np.ascontiguousarray(a, np.float64) for a in (r, t, T_axis))
    T = np.zeros((4, 4))
    _lib.load().ilcc2_compose_extrinsic(r.ctypes.data, t.ctypes.data, T_axis.ctypes.data, T.ctypes.data)
    return T


def extrinsic2bin(path: str, T: np.ndarray):
    T = np.ascontiguousarray(T, np.float64)
    if _lib.load().ilcc2_extrinsic2bin(path.encode(), T.ctypes.data) != 0:
        raise Ilcc2Error(f"can not write {path}")


def read_cam_corners(path: str, num: int, board_h: int) -> np.ndarray:
    out = np.zeros((4 * num, 2))
    n = _lib.load().ilcc2_read_cam_corners(path.encode(), num, board_h, out.ctypes.data, out.shape[0])
    if n < 0:
        raise Ilcc2Error(f"can not read {path}")
    return out[:n].copy()


def read_lidar_corners(path: str, num: int) -> np.ndarray:
    out = np.zeros((num, 3))
    n = _lib.load().ilcc2_read_lidar_corners(path.encode(), num, out.ctypes.data)
    if n < 0:
        raise Ilcc2Error(f"can not read {path}")
    return out[:n].copy()


def read_camera_yaml(path: str):
    K = np.zeros((3, 3))
    cx, cy = C.c_int(0), C.c_int(0)
    if _lib.load().ilcc2_read_camera_yaml(path.encode(), K.ctypes.data, C.byref(cx), C.byref(cy)) != 0:
        raise Ilcc2Error(f"can not read K from {path}")
    return K, cx.value, cy.value


class PnpSolver:
    """ilcc2_pnp_* : check_order_lidar / check_order_cam for a batch of frames and Optimization::solvePose3d2dError,
    both on the GPU (csrc/ilcc2_pnp.cu)."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        rc = self._L.ilcc2_pnp_create(device, C.byref(self._h))
        if rc != 0:
            raise Ilcc2Error(f"ilcc2_pnp_create failed ({rc}): {self._L.ilcc2_pnp_last_error(None).decode()}")
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.ilcc2_pnp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise Ilcc2Error(f"ilcc2 pnp error {rc}: {self._L.ilcc2_pnp_last_error(self._h).decode()}")

    @property
    def launch_count(self) -> int:
        return int(self._L.ilcc2_pnp_launch_count(self._h))

    def order_corners(self, lidar_corners, T_axis, board_w: int, board_h: int, status=None, cam_corners=None):
        """numpy host arrays or torch CUDA tensors. Returns (pts3 (F*nc,3) f64, pts2 (F*nc,2) f64 | None, valid (F*nc,) u8)."""
        T_axis = np.ascontiguousarray(T_axis, np.float64)
        nc = board_w * board_h
        if _is_torch(lidar_corners):
            import torch
            lc = lidar_corners.contiguous()
            F = lc.shape[0]
            assert lc.dtype == torch.float32 and lc.shape[1] == nc
            p3 = torch.empty((F * nc, 3), dtype=torch.float64, device=lc.device)
            p2 = torch.empty((F * nc, 2), dtype=torch.float64, device=lc.device) if cam_corners is not None else None
            va = torch.empty((F * nc,), dtype=torch.uint8, device=lc.device)
            cc = cam_corners.contiguous() if cam_corners is not None else None
            stt = status.contiguous() if status is not None else None
            stream = torch.cuda.current_stream(lc.device).cuda_stream
            self._check(self._L.ilcc2_pnp_order_corners(
                self._h, lc.data_ptr(), stt.data_ptr() if stt is not None else None,
                cc.data_ptr() if cc is not None else None, F, board_w, board_h, T_axis.ctypes.data, p3.data_ptr(),
                p2.data_ptr() if p2 is not None else None, va.data_ptr(), MEM_DEVICE, stream))
            return p3, p2, va
        lc = np.ascontiguousarray(lidar_corners, np.float32)
        F = lc.shape[0]
        assert lc.shape[1:] == (nc, 3)
        p3 = np.empty((F * nc, 3))
        cc = np.ascontiguousarray(cam_corners, np.float64) if cam_corners is not None else None
        p2 = np.empty((F * nc, 2)) if cc is not None else None
        va = np.empty((F * nc,), np.uint8)
        stt = np.ascontiguousarray(status, np.int32) if status is not None else None
        self._check(self._L.ilcc2_pnp_order_corners(
            self._h, lc.ctypes.data, stt.ctypes.data if stt is not None else None,
            cc.ctypes.data if cc is not None else None, F, board_w, board_h, T_axis.ctypes.data, p3.ctypes.data,
            p2.ctypes.data if p2 is not None else None, va.ctypes.data, MEM_HOST, None))
        return p3, p2, va

    def solve(self, pts3, pts2, camera, r0=(0.0, 0.0, 0.0), t0=(0.0, 0.0, 0.0), valid=None, huber=0.1,
              max_iterations=50):
        """Optimization::solvePose3d2dError. Returns (r, t, summary dict)."""
        camera = np.ascontiguousarray(camera, np.float64)
        r, t = np.array(r0, np.float64), np.array(t0, np.float64)
        sm = _lib.PnpSummary()
        if _is_torch(pts3):
            import torch
            p3, p2 = pts3.contiguous(), pts2.contiguous()
            assert p3.dtype == torch.float64 and p2.dtype == torch.float64
            v = valid.contiguous() if valid is not None else None
            stream = torch.cuda.current_stream(p3.device).cuda_stream
            self._check(self._L.ilcc2_pnp_solve(self._h, p3.data_ptr(), p2.data_ptr(),
                                                v.data_ptr() if v is not None else None, p3.shape[0],
                                                camera.ctypes.data, huber, max_iterations, r.ctypes.data,
                                                t.ctypes.data, C.byref(sm), MEM_DEVICE, stream))
        else:
            p3, p2 = np.ascontiguousarray(pts3, np.float64), np.ascontiguousarray(pts2, np.float64)
            v = np.ascontiguousarray(valid, np.uint8) if valid is not None else None
            self._check(self._L.ilcc2_pnp_solve(self._h, p3.ctypes.data, p2.ctypes.data,
                                                v.ctypes.data if v is not None else None, p3.shape[0],
                                                camera.ctypes.data, huber, max_iterations, r.ctypes.data,
                                                t.ctypes.data, C.byref(sm), MEM_HOST, None))
        return r, t, {k: getattr(sm, k) for k, _ in _lib.PnpSummary._fields_}

    def project_points(self, points, T_lidar2cam, camera, width: int, height: int, distance_valid: float = 5.0,
                       inten_low: float = 0.0, inten_high: float = 60.0, image_bgr=None, pix=None, rgbv=None):
        """ImageCornersEst::spaceToPlane for every point + the colouring of pcd2image.cpp:59-84 (image_bgr None: HSV of
        the intensity) or rgblidar.cpp:53-73 (colour of the pixel).  points (n, 4|8) float32, numpy or torch CUDA.
        Returns (pix (n,2) int32, rgbv (n,4) uint8 = r, g, b, valid)."""
        T = np.ascontiguousarray(T_lidar2cam, np.float64)
        cam = np.ascontiguousarray(camera, np.float64)
        if _is_torch(points):
            import torch
            assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous() and points.dim() == 2
            n, wdt = points.shape
            step, ox, oy, oz, oi = _layout(wdt)
            if pix is None:
                pix = torch.empty((n, 2), dtype=torch.int32, device=points.device)
            if rgbv is None:
                rgbv = torch.empty((n, 4), dtype=torch.uint8, device=points.device)
            img = image_bgr.contiguous() if image_bgr is not None else None
            stream = torch.cuda.current_stream(points.device).cuda_stream
            self._check(self._L.ilcc2_project_points(
                self._h, points.data_ptr(), n, step, ox, oy, oz, oi, T.ctypes.data, cam.ctypes.data, width, height,
                distance_valid, inten_low, inten_high, img.data_ptr() if img is not None else None, pix.data_ptr(),
                rgbv.data_ptr(), MEM_DEVICE, stream))
            return pix, rgbv
        pts = np.ascontiguousarray(points, np.float32)
        n, wdt = pts.shape
        step, ox, oy, oz, oi = _layout(wdt)
        pix, rgbv = np.empty((n, 2), np.int32), np.empty((n, 4), np.uint8)
        img = np.ascontiguousarray(image_bgr, np.uint8) if image_bgr is not None else None
        self._check(self._L.ilcc2_project_points(
            self._h, pts.ctypes.data, n, step, ox, oy, oz, oi, T.ctypes.data, cam.ctypes.data, width, height,
            distance_valid, inten_low, inten_high, img.ctypes.data if img is not None else None, pix.ctypes.data,
            rgbv.ctypes.data, MEM_HOST, None))
        return pix, rgbv
