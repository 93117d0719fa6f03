// Internal declarations shared by the kernels (ilcc2_kernels.cu) and the C-ABI host code (ilcc2_api.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ilcc2_b200.h"

namespace ilcc2 {

// Kernel-side copy of ilcc2_params plus derived constants (computed once on the host, in double).
struct DevParams {
  float roi_hx, roi_hy, roi_hz;
  float tol2;            // (float)(tol*tol): FLANN radius search compares squared distances
  float cell;            // cluster hash cell edge = 0.95 * tol / sqrt(3): same cell => linked
  float inv_cell;        // 1 / cell (the 5 % slack of the cell edge covers the rounding of x * inv_cell vs x / cell)
  int ncx, ncy, ncz;     // cells per axis over the ROI box
  int min_cl, max_cl;
  double ransac_thr;
  int ransac_max_iter;
  double ransac_log_prob;
  int hist_bins;
  double gray_rate;
  double huber_a, huber_b;
  double g, inv_g, wg2, hg2;
  int W, H;
  int min_board;
  int max_solver_iter;
  int ncorners;
};

constexpr int kRoiThreads = 512;      // K1 CTA size for small batches (128 for batches of >= 4 frames per SM)
constexpr int kRoiUnroll = 4;         // K1 records per thread per tile
constexpr int kClusterThreads = 256;  // K2 CTA size
constexpr int kClusterSmemPts = 4096; // K2: frames with at most this many ROI points cluster in shared memory
constexpr int kClusterSmemSlots = 4096;  // K2: hash slots (cells) in shared memory; 16-bit tables -> 66 KB per CTA, 3 CTAs per SM
constexpr int kClusterMaxOcc = 3072;     // K2: more occupied cells than this -> the frame goes to the global-memory tables
constexpr int kClusterBitmapRows = 2944; // K2: occupancy bitmap rows (ncy*ncz) that fit the shared-memory layout
constexpr int kClusterQueue = 192;       // K2: per-warp queue of (cell, neighbour) pairs: 32 lanes x 5 dx of a round + up to 31 waiting
constexpr int kPlaneThreads = 128;    // K3/K4 CTA size == canonical reduction width of their double sums
// K5: one CTA per (frame, colour hypothesis); the CTA is a team of T warps that share the solve.  T is chosen per
// launch so that the batch fills the machine (launch_board_fit): 2 for large batches, 4 / 8 for small ones, where the
// latency of one solve is what the caller waits for.  The CTA that finishes second picks the lower cost and emits
// the corners.
constexpr int kFitMaxTeam = 8;
constexpr int kFitGlobalTeam = 2;     // team size of the global-memory instantiation (frames above kFitBitsCap points)
#ifndef FIT_TEAM_CAP
#define FIT_TEAM_CAP 4096
#endif
// K5: optimiser points of a frame that the shared-memory instantiation takes (multiple of 32 * kFitMaxTeam).  Each of
// the T warps of a team owns a candidate list of kFitTeamCap / T 16-bit entries in shared memory, so the footprint
// (2 x 8 KB of lists + 0.5 KB of colour bits per CTA) does not depend on T and most of the SM's 256 KB stay L1 for the
// points.  Frames with more points take the global-memory instantiation (tables from the handle's overflow pool).
constexpr int kFitTeamCap = FIT_TEAM_CAP;
constexpr int kFitBitsCap = kFitTeamCap;
// Handles created for big frames (max_points_per_frame > kFitBigFrame, e.g. the 64 x 2048 HDL-64-like clouds whose
// boards hold 10-14 k optimiser points) use a larger shared-memory layout instead: 2 KB of colour bits + 2 x 32 KB of lists.
constexpr int kFitBigFrame = 65536;
constexpr int kFitTeamCapBig = 16384;
constexpr int kFitBitsCapBig = kFitTeamCapBig;
#ifndef FIT_BATCH
#define FIT_BATCH 8
#endif
constexpr int kFitBatch = FIT_BATCH;   // K5: candidates (radius halvings) prepared per trust-region model, one per lane
constexpr int kRndTable = 8192;       // PCL RANSAC draw table length

constexpr int kMaxGatherTargets = 8;   // ranks of one NVSwitch domain (or 1: a multicast mapping that reaches all of them)

// Everything a launch needs; passed by value.
struct Workspace {
  // caller input (device pointers)
  const unsigned char* points;
  const long long* frame_off;  // nullable
  int pts_per_frame;
  int step, ox, oy, oz, oi;
  const float* clicks;
  const uint8_t* flags;        // nullable
  uint32_t default_flags;
  const int8_t* pdelta;        // nullable
  int n_frames;                // frames in THIS launch (a chunk of the call's batch)
  int frame_base;              // first frame of the chunk; kernels process frames frame_base + blockIdx.x
  int load_frames;             // frames in flight on the device at the same time (the whole call when its chunks run
                               // concurrently, else n_frames): what CTA sizes / K5 team sizes are chosen for
  int cap;                     // per-frame capacity R of the ROI-sized arrays
  // scratch
  float4* roi_pts;             // [F][R] ROI points (x,y,z,intensity), input order
  int* roi_src;                // [F][R] original index of each ROI point
  int* roi_count;              // [F]
  unsigned long long* nn_key;  // [F] (dist2 bits << 32 | ROI index) of the click's nearest neighbour
  int* sel_idx;                // [F][R] chosen cluster, ROI indices ascending
  int* sel_count;              // [F]
  int* brd_idx;                // [F][R] plane inliers, ROI indices ascending
  int* brd_count;              // [F]
  int* tmp_idx;                // [F][R] RANSAC permutation, then inliers of the sample model
  int* cl_scratch;             // [cl_pool_slots][cl_stride] global-memory clustering tables for large ROIs (a pool:
                               // frames that need them take the next free slot through cl_pool_next)
  int cl_pool_slots;
  int* cl_pool_next;
  int cl_ts;                   // slots in the global hash table (power of two)
  long long cl_stride;
  float2* opt_yz;              // [F][R] non-gray board points in the PCA frame (y', z')
  uint8_t* opt_white;          // [F][R]
  int* opt_count;              // [F]
  int* fit_order;              // [F] block index -> frame for K5 (work-balanced order)
  int* fit_done;               // [F] K5 with one CTA per hypothesis: CTAs of the frame that have finished
  int fit_bits_cap;            // K5 shared-memory layout of this handle: colour-bit capacity (points) ...
  int fit_team_cap;            // ... and points per chunk of a solve (list entries per team)
  int num_sms;
  uint32_t* fit_scratch;       // K5 global-memory path: colour bits + index lists, in slots of the SAME pool as
  long long fit_stride;        // cl_scratch (K2 is done with it by then): slot stride = cl_stride, taken through
  int* fit_pool_next;          // fit_pool_next
  const uint32_t* rnd_table;   // [kRndTable] PCL RANSAC draws
  // outputs (device)
  float* corners;              // [F][ncorners][3]
  int32_t* status;             // [F]
  ilcc2_diag* diag;            // [F] (always present internally)
  // multi-GPU: K5 also stores corners / status into peer-mapped gather buffers (ilcc2_set_gather_targets)
  int n_targets;               // 0 = off; else entries of gather_corners / gather_status in use
  int gather_multicast;        // 1: target 0 is a multicast (NVLS) mapping -> multimem.st; 0: ordinary peer pointers
  int gather_slot_base;        // frame f of this handle goes to slot gather_slot_base + f of every target
  float* gather_corners[kMaxGatherTargets];
  int32_t* gather_status[kMaxGatherTargets];
};

void launch_roi_filter(const Workspace& w, const DevParams& p, cudaStream_t s);
void launch_cluster(const Workspace& w, const DevParams& p, cudaStream_t s);
void launch_plane_pca(const Workspace& w, const DevParams& p, cudaStream_t s);
void launch_board_fit(const Workspace& w, const DevParams& p, cudaStream_t s);
cudaError_t configure_kernels();
size_t cluster_smem_bytes();
size_t fit_smem_bytes(const Workspace& w, int team);
int fit_team_size(int frames, int num_sms);

}  // namespace ilcc2
