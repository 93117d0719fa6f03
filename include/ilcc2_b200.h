/*
 * ilcc2_b200.h -- C ABI of the B200-native LiDAR chessboard-corner extractor (libilcc2_b200.so).
 *
 * Drop-in boundary for ONE path of icameling/lidar_camera_calibration: what the ilcc2 node
 * ilcc2/test/get_lidar_corners.cpp:183-198 does with a LidarCornersEst object between receiving a
 * PointCloud2 + clicked point and writing process_data/<cam>_lidar_<i>.txt:
 *
 *     setROI(cloud, click)      ilcc2/src/LidarCornersEst.cpp:48-70
 *     EuclideanCluster()        ilcc2/src/LidarCornersEst.cpp:124-186   (+ getPlane :190-221)
 *     PCA()                     ilcc2/src/LidarCornersEst.cpp:366-372   (+ transformbyPCA :330-364, get_gray_zone/calHist :224-328)
 *     get_corners(corners)      ilcc2/src/LidarCornersEst.cpp:374-450   (+ Optimization::get_theta_t, Optimization.cpp:94-160,
 *                                                                         VirtualboardError, Optimization.h:12-121, getPCDcorners :501-556)
 *     save_corners2txt(...)     ilcc2/test/get_lidar_corners.cpp:27-36
 *
 * The reference has no FFI; its seam is that C++ class (ilcc2/include/ilcc2/LidarCornersEst.h:40-67).
 * This header is what a binding for that seam calls; include/ilcc2_b200/LidarCornersEst.hpp is the
 * source-compatible C++ adaptor and INTEGRATION.md shows the reference-side change.
 *
 * All entry points are plain C: pointers, sizes, no C++/torch types.  The GUI key presses of the
 * reference (Visualization.cpp:20-57) become per-frame flags.  There is no CPU fallback: every compute
 * entry point fails with ILCC2_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef ILCC2_B200_H_
#define ILCC2_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ILCC2_VERSION 100

/* ---- function return codes (0 = success) ---- */
enum {
  ILCC2_SUCCESS = 0,
  ILCC2_ERR_INVALID_ARGUMENT = 1,
  ILCC2_ERR_CUDA = 2,          /* CUDA runtime error or no usable device; see ilcc2_last_error */
  ILCC2_ERR_CAPACITY = 3,      /* more frames / points than the handle was created for */
  ILCC2_ERR_IO = 4
};

/* ---- per-frame status (status_out); the reference's bool returns / exceptions / UB cases ---- */
enum {
  ILCC2_OK = 0,
  ILCC2_NO_POINTS_IN_ROI = 1,     /* setROI left nothing */
  ILCC2_NO_CLUSTER = 2,           /* cluster_indices empty: reference throws std::out_of_range (:167) */
  ILCC2_PLANE_TOO_SMALL = 3,      /* fewer plane inliers than min_board_points (online gate, :111) or < 3 */
  ILCC2_DEGENERATE_INTENSITY = 4, /* max==min (division by zero, :235) or the mode walk runs off the map (:266-282) */
  ILCC2_NO_OPT_POINTS = 5,        /* every board point fell in the gray zone */
  ILCC2_RANSAC_FAILED = 6,        /* no non-degenerate 3-point sample */
  ILCC2_ROI_CAPACITY = 7          /* the frame's ROI holds more points than the handle's roi_capacity, or more frames of
                                     the call needed the global-memory clustering tables than the handle holds
                                     (ilcc2_create_sized); never with ilcc2_create's defaults */
};

/* ---- per-frame flags = the reference's key presses ---- */
#define ILCC2_FLAG_TOPLEFT_WHITE 1u /* 'd' once: re-optimise with topleftWhite=true, warm-started (:398-409) */
#define ILCC2_FLAG_INVERSE       2u /* 'a' once: swap (W,H) in the corner lattice only (:394-397, :504-507) */
#define ILCC2_FLAG_AUTO          4u /* unattended: both colour hypotheses from zero, keep the lower final cost */

/* ---- where the caller's buffers live ---- */
#define ILCC2_MEM_HOST   0 /* points/clicks/flags/outputs are host pointers; the call returns when outputs are valid */
#define ILCC2_MEM_DEVICE 1 /* all of them are device pointers on the handle's device; asynchronous on `stream` */

/* Algorithm constants. ilcc2_default_params() fills the literals hard-coded in the reference. */
typedef struct ilcc2_params {
  float roi_half_x;              /* 1.0   LidarCornersEst.cpp:59 */
  float roi_half_y;              /* 1.5   :64 */
  float roi_half_z;              /* 2.0   :54 */
  float cluster_tolerance;       /* 0.12  :131   (0.10 in get_chessboard_by_point, :79) */
  int32_t min_cluster_size;      /* 100   :132 */
  int32_t max_cluster_size;      /* 25000 :133 */
  double ransac_threshold;       /* 0.03  :201 */
  int32_t ransac_max_iterations; /* 50    pcl::SACSegmentation default */
  double ransac_probability;     /* 0.99  pcl::SACSegmentation default */
  int32_t hist_bins;             /* 100   :226 */
  double gray_rate;              /* 2.5   :371 */
  double huber_delta;            /* 0.1   Optimization.cpp:137 */
  double grid_length;            /* 0.15  config/pointgrey.yaml:17 */
  int32_t cells_x;               /* (corner_in_x+1, corner_in_y+1) sorted ascending -> (6, 8)  :31-39 */
  int32_t cells_y;
  int32_t min_board_points;      /* 0 offline; 500 reproduces the online gate (:111) */
  int32_t max_solver_iterations; /* 50    ceres::Solver::Options default */
} ilcc2_params;

/* Optional per-frame diagnostics (what the reference prints to stdout, plus solver state). */
typedef struct ilcc2_diag {
  int32_t status;
  int32_t n_roi;             /* points kept by setROI */
  int32_t nn_index;          /* ROI index of the click's nearest neighbour (:141) */
  int32_t n_clusters;        /* clusters with min<=size<=max */
  int32_t click_in_cluster;  /* 1 if that neighbour lies in a kept cluster (:144-153) */
  int32_t cluster_rank;      /* size-descending rank of the chosen cluster (= m_plane_index) */
  int32_t n_cluster_pts;
  int32_t ransac_iterations;
  int32_t n_ransac_inliers;  /* inliers of the best 3-point model */
  int32_t n_board;           /* "chessboard plane size" (:184) */
  int32_t n_opt;             /* black + white points (gray ones are skipped, Optimization.cpp:122-125) */
  int32_t hyp_chosen;        /* 0: topleftWhite=false, 1: true */
  int32_t iters[2][2];       /* [hypothesis][phase: 0 useOutofBoard, 1 not] solver iterations */
  int32_t term[2][2];        /* termination: 0 max iterations, 1 gradient, 2 parameter, 3 function tol., 4 radius, 5 invalid */
  int32_t n_evals[2];        /* residual evaluations actually executed per hypothesis */
  float plane[4];            /* refined plane a,b,c,d */
  float pca[16];             /* pca_matrix, row-major 4x4 (:347-349) */
  float eigvals[3];          /* eigenValuesPCA (:339) */
  double hist_min, hist_max, hist_mean, r_low, r_high; /* calHist (:256, :300) */
  double gray[2];            /* gray_zone (:322-323) */
  double theta_t[2][3];      /* per hypothesis: theta, t_y, t_z */
  double cost[2];            /* per hypothesis: final Huber cost of the second phase */
  double cost_a[2];          /* per hypothesis: final Huber cost of the first (useOutofBoard) phase */
} ilcc2_diag;

typedef struct ilcc2_handle ilcc2_handle;

int ilcc2_version(void);
void ilcc2_default_params(ilcc2_params* p);

/* Reads grid_length / corner_in_x / corner_in_y from an OpenCV-YAML file the way
 * LidarCornersEst::set_chessboard_param does (LidarCornersEst.cpp:20-46) and fills
 * p->grid_length, p->cells_x <= p->cells_y. Other fields of *p are left untouched. */
int ilcc2_set_chessboard_param(ilcc2_params* p, const char* cam_yaml_path);

/* Creates a handle that owns all device scratch for up to max_frames frames of up to
 * max_points_per_frame points each, including the staging ring that host batches stream through (no allocation per
 * call afterwards; the first call with PAGEABLE host memory adds a page-locked ring and copy threads, once).
 * One handle per (GPU, stream); a handle is not thread-safe. */
int ilcc2_create(const ilcc2_params* params, int device, int max_frames, int max_points_per_frame,
                 ilcc2_handle** out);
/* Same, with the scratch sized for the caller's data instead of the worst case: the ROI-sized arrays (41 bytes per
 * entry and frame) hold roi_capacity points per frame (0 = max_points_per_frame: every point may survive setROI; the
 * bundled VLP-16 bags keep 2.9-5.4 k of 24 k points), and the global-memory clustering tables that ROIs above 4096
 * points (or very sparse ones) use exist for overflow_frames frames of a call (0 = every frame when roi_capacity >
 * 4096, else max_frames / 8).  A frame beyond either limit reports ILCC2_ROI_CAPACITY; nothing is truncated. */
int ilcc2_create_sized(const ilcc2_params* params, int device, int max_frames, int max_points_per_frame,
                       int roi_capacity, int overflow_frames, ilcc2_handle** out);
int ilcc2_destroy(ilcc2_handle* h);

/* Last error text of the handle (or of the last failed ilcc2_create when h == NULL). */
const char* ilcc2_last_error(const ilcc2_handle* h);

/* The hot path for a batch of frames: replaces setROI + EuclideanCluster + PCA + get_corners.
 *   points        n_frames frames of point records, back to back.  A record is `point_step` bytes with
 *                 float32 x, y, z, intensity at byte offsets off_x/off_y/off_z/off_intensity:
 *                 (16; 0,4,8,12) = packed XYZI float4;  (32; 0,4,8,16) = pcl::PointXYZI and the velodyne
 *                 PointCloud2 records of the bundled bags (zero-copy from msg.data).
 *   frame_offsets NULL -> every frame has points_per_frame records.  Else n_frames+1 record offsets (host
 *                 memory, int64) so frames may have different sizes; points_per_frame is then ignored.
 *   click_xyz     [n_frames][3] float32: /clicked_point per frame (get_lidar_corners.cpp:38-50).
 *   flags         NULL (-> default_flags for all frames) or [n_frames] uint8 of ILCC2_FLAG_*.
 *   plane_index_delta NULL or [n_frames] int8: net 'w' (+) / 's' (-) presses (Visualization.cpp:22-29).
 *   corners_out   [n_frames][(cells_x-1)*(cells_y-1)][3] float32, LiDAR frame, reference lattice order
 *                 (LidarCornersEst.cpp:524-534); NaN when status != ILCC2_OK.
 *   status_out    [n_frames] int32 ILCC2_OK...
 *   diag_out      NULL or [n_frames] ilcc2_diag.
 *   stream        cudaStream_t (NULL = default stream).
 * mem_kind says where points / click_xyz / flags / plane_index_delta / *_out live. */
int ilcc2_extract_corners(ilcc2_handle* h, const void* points, int n_frames, int points_per_frame,
                          const int64_t* frame_offsets, int point_step, int off_x, int off_y, int off_z,
                          int off_intensity, int mem_kind, const float* click_xyz, const uint8_t* flags,
                          uint32_t default_flags, const int8_t* plane_index_delta, float* corners_out,
                          int32_t* status_out, ilcc2_diag* diag_out, void* stream);

/* Intermediate clouds of the LAST call for one frame, as the reference publishes them
 * (m_cloud_ROI / cluster / m_cloud_chessboard; get_lidar_corners.cpp:192-201): original point indices
 * into that frame, in input order.  which: 0 ROI, 1 chosen cluster, 2 chessboard plane.
 * Synchronises the stream. Returns the count (<= cap entries written) or a negative error. */
int ilcc2_fetch_indices(ilcc2_handle* h, int frame, int which, int32_t* out_host, int cap);

/* Non-gray chessboard points in the PCA frame (m_cloud_PCA y,z + colour class) of the LAST call:
 * yz_out [cap][2] float32, white_out [cap] uint8. Returns the count or a negative error. */
int ilcc2_fetch_board_points(ilcc2_handle* h, int frame, float* yz_out_host, uint8_t* white_out_host, int cap);

/* Per-kernel device times (ms, CUDA events on the call's stream) of the LAST call when profiling is on.
 * names_out receives pointers to static strings. Returns the number of stages. */
int ilcc2_set_profiling(ilcc2_handle* h, int enabled);
int ilcc2_stage_times(ilcc2_handle* h, float* ms_out, const char** names_out, int cap);

/* Throughput mode for a STREAM of device-resident batches: `lanes` handles, each with a CUDA stream of its own, take
 * consecutive batches in turn (ilcc2_extract_corners is asynchronous for ILCC2_MEM_DEVICE), so that the kernels of batch
 * k+1 run on the SMs that the tail of batch k's K5 leaves idle.  This call tells a handle that it is one of `lanes` such
 * handles: launch shapes that depend on how many frames are in flight on the device (K5's warps per solve) are then
 * chosen for lanes x the call's frames.  Results do not depend on it (same launches, same arithmetic; the partial sums
 * of a solve are grouped by its warps, which moves theta/t in the 15th digit and leaves the fp32 corners as they are).
 * lanes = 1 (default): one batch in flight.  The reference processes one cloud at a time (get_lidar_corners.cpp:136-198);
 * this is the batch-stream counterpart of its bag loop. */
int ilcc2_set_lanes(ilcc2_handle* h, int lanes);

/* Kernels launched by this handle since creation (bench.py's gpu_launches). */
int64_t ilcc2_launch_count(const ilcc2_handle* h);

/* Output contract: process_data/<camera_name>_lidar_<bag_idx>.txt, "x y z\n" per corner with the default
 * ostream precision of 6 significant digits (get_lidar_corners.cpp:27-36). */
int ilcc2_save_corners2txt(const float* corners_xyz, int n_corners, const char* path);

/* Multi-GPU: the all-gather of the per-frame corners fused into the kernel that produces them.
 * After this call every ilcc2_extract_corners of the handle ALSO stores frame f's corners and status into each target at
 * frame slot slot_base + f: corners_targets[t] is a device-accessible float[total_frames][n_corners][3],
 * status_targets[t] an int32[total_frames].
 *   target_kind ILCC2_GATHER_PEER       n_targets = world size: peer-mapped buffers of the ranks of one NVLink / NVSwitch
 *                                       domain (torch symmetric memory, cudaIpc mappings, peer-enabled pointers): plain
 *                                       stores, one per target.
 *               ILCC2_GATHER_MULTICAST  n_targets = 1: ONE multicast (NVLS) mapping of such a buffer; stored to with
 *                                       multimem.st, NVSwitch fans the store out to every rank.
 * The caller separates producers and consumers with one barrier per step (no data-path collective) AND must make sure
 * that no rank starts the next step's extract into a buffer that another rank is still reading: alternate between two
 * buffers (call this function before every step; it only updates the handle) or add a second barrier after the read.
 * n_targets = 0 switches it off.  At most 8 targets. */
#define ILCC2_GATHER_PEER      0
#define ILCC2_GATHER_MULTICAST 1
int ilcc2_set_gather_targets(ilcc2_handle* h, int n_targets, float* const* corners_targets,
                             int32_t* const* status_targets, int slot_base, int target_kind);

/* The same for a C / C++ host that runs one rank per GPU over an NCCL communicator (what a multi-GPU build of
 * ilcc2/test/calib_lidar_cam.cpp:116-165, the consumer of every frame's corners, would hold): allocates a double-
 * buffered gather buffer for total_frames frames on this rank, exchanges cudaIpc handles (or, between ranks of one
 * process, peer-enabled pointers) through ncclAllGather on `nccl_comm` (an ncclComm_t; NCCL is resolved from the host
 * process at run time, this library does not link it) and points the fused stores of K5 at every rank's buffer.  If any
 * rank cannot map a peer, all ranks fall back to a plain ncclAllGather inside ilcc2_gather_sync.  Collective: every
 * rank of the communicator calls it; slot_base = this rank's first frame in the gathered order.
 *   ilcc2_gather_sync     after the step's ilcc2_extract_corners (ILCC2_MEM_DEVICE, same stream): the barrier between
 *                         producers and consumers (one ncclAllReduce of an int) -- or the ncclAllGather of the
 *                         fallback -- enqueued on `stream`; returns the device pointers that hold ALL ranks' corners /
 *                         status once `stream` reaches this point, and switches the handle to the other half of the
 *                         double buffer for the next step.
 *   ilcc2_gather_is_fused 1: K5 stores over NVLink; 0: NCCL fallback. */
int ilcc2_gather_setup_nccl(ilcc2_handle* h, void* nccl_comm, int total_frames, int slot_base);
int ilcc2_gather_sync(ilcc2_handle* h, const float** corners_all, const int32_t** status_all, void* stream);
int ilcc2_gather_is_fused(const ilcc2_handle* h);
int ilcc2_gather_teardown(ilcc2_handle* h);

/* Frame sharding across ranks (one process per GPU): contiguous blocks. */
void ilcc2_shard_range(int n_frames, int world_size, int rank, int* begin, int* count);

/* =====================================================================================================
 * What follows the corner path in the reference: ordering the 3-D / 2-D corners
 * and the LiDAR->camera pose solve of ilcc2/test/calib_lidar_cam.cpp:101-165.  Kernels: csrc/ilcc2_pnp.cu.
 * ===================================================================================================== */
typedef struct ilcc2_pnp ilcc2_pnp;

typedef struct ilcc2_pnp_summary { /* ceres::Solver::Summary of Optimization.cpp:68-69, the fields the node prints */
  double initial_cost, final_cost;
  int32_t iterations;  /* trust-region iterations (successful + unsuccessful) */
  int32_t termination; /* 0 max iterations, 1 gradient, 2 parameter, 3 function tolerance, 4 radius, 5 invalid steps */
  int32_t n_evals;     /* residual + Jacobian evaluations as Ceres counts them */
  int32_t n_pairs;     /* -1: not counted on the device */
} ilcc2_pnp_summary;

int ilcc2_pnp_create(int device, ilcc2_pnp** out);
int ilcc2_pnp_destroy(ilcc2_pnp* h);
const char* ilcc2_pnp_last_error(const ilcc2_pnp* h);
int64_t ilcc2_pnp_launch_count(const ilcc2_pnp* h);

/* get_lidar2cam_axis_roughly (calib_lidar_cam.cpp:50-70): row-major 4x4 for camera_name in
 * front / car_left / pointgrey / left / right / back; ILCC2_ERR_INVALID_ARGUMENT otherwise. Host only. */
int ilcc2_axis_roughly(const char* camera_name, double T_rowmajor[16]);

/* The "data ready" loop of calib_lidar_cam.cpp:119-123 for a batch of frames, on the device:
 *   pts3_out[f][k]  = check_order_lidar(T_axis * (double)lidar_corners[f][k])   (ImageCornersEst.cpp:461-487)
 *   pts2_out[f][k]  = check_order_cam(cam_corners[f][k])                        (ImageCornersEst.cpp:430-459)
 *   valid_out[f*nc+k] = status == NULL || status[f] == ILCC2_OK                 (nullable)
 * lidar_corners [n_frames][board_w*board_h][3] float32 as ilcc2_extract_corners writes them (board_w = the inner
 * lattice index, e.g. 7; board_h = 5); cam_corners / pts2_out nullable.  T_axis is a HOST pointer.
 * mem_kind: where lidar_corners / status / cam_corners / *_out live. */
int ilcc2_pnp_order_corners(ilcc2_pnp* h, const float* lidar_corners, const int32_t* status, const double* cam_corners,
                            int n_frames, int board_w, int board_h, const double T_axis[16], double* pts3_out,
                            double* pts2_out, uint8_t* valid_out, int mem_kind, void* stream);

/* Optimization::solvePose3d2dError (Optimization.cpp:13-91): Huber(huber_delta) reprojection error of pts3 under
 * (angle-axis r, translation t) against pts2 with camera = fx, cx, fy, cy; Ceres trust region / dense normal
 * Cholesky / subspace dogleg with its defaults (max_iterations 50 in the reference).
 * pts3 [n_pairs][3], pts2 [n_pairs][2] float64, valid [n_pairs] (nullable) live where mem_kind says; camera, r, t
 * (in: start, out: result) and summary are HOST pointers.  The call returns when the result is valid. */
int ilcc2_pnp_solve(ilcc2_pnp* h, const double* pts3, const double* pts2, const uint8_t* valid, int n_pairs,
                    const double camera[4], double huber_delta, int max_iterations, double r[3], double t[3],
                    ilcc2_pnp_summary* summary, int mem_kind, void* stream);

/* Using the calibration: ImageCornersEst::spaceToPlane (ImageCornersEst.cpp:135-155) for every
 * point of a cloud plus the colouring loops of ilcc2/test/pcd2image.cpp:59-84 and ilcc2/test/rgblidar.cpp:53-73.
 *   P_c = R * P_w + t with [R|t] = T_lidar2cam (row-major 4x4, what txt2extrinsic reads back from <pose_file>);
 *   a point is valid iff 0 <= P_c.z <= distance_valid and 0 < u < width, 0 < v < height (u = fx x/z + cx, ...).
 *   pix_xy_out  [n][2] int32: (int)u, (int)v -- or (-1, -1) for an invalid point
 *   rgbv_out    [n][4] uint8: r, g, b, valid.  image_bgr == NULL: HSVtoRGB((I - inten_low) / (inten_high - inten_low)
 *               * 255, 100, 100) as pcd2image does (it uses inten_low 0, inten_high 60, distance_valid 5);
 *               else the bgr8 pixel [v][u] of image_bgr ([height][width][3]) as rgblidar does.
 * Point records as in ilcc2_extract_corners.  camera = fx, cx, fy, cy and T_lidar2cam are HOST pointers. */
int ilcc2_project_points(ilcc2_pnp* h, const void* points, int64_t n_points, int point_step, int off_x, int off_y,
                         int off_z, int off_intensity, const double T_lidar2cam[16], const double camera[4], int width,
                         int height, double distance_valid, double inten_low, double inten_high,
                         const uint8_t* image_bgr, int32_t* pix_xy_out, uint8_t* rgbv_out, int mem_kind, void* stream);

/* T_lidar2cam = [AngleAxis(|r|, r/|r|) | t] * T_axis_roughly (Optimization.cpp:75-88, calib_lidar_cam.cpp:157-160). */
int ilcc2_compose_extrinsic(const double r[3], const double t[3], const double T_axis[16], double T_out_rowmajor[16]);

/* ImageCornersEst::extrinsic2txt (ImageCornersEst.cpp:301-306): 128 bytes, Eigen::Matrix4d column-major. */
int ilcc2_extrinsic2bin(const char* path, const double T_rowmajor[16]);

/* The reference's text readers (ImageCornersEst.cpp:213-299) and camK of getRectifyParam (:15-31). Host only.
 * The two corner readers return the number of corners read, or a negative ILCC2_ERR_*. */
int ilcc2_read_cam_corners(const char* path, int num, int board_h, double* out_xy, int cap);
int ilcc2_read_lidar_corners(const char* path, int num, double* out_xyz);
int ilcc2_read_camera_yaml(const char* path, double K_rowmajor[9], int* corner_in_x, int* corner_in_y);

#ifdef __cplusplus
}
#endif
#endif /* ILCC2_B200_H_ */
