// ilcc2_api.cu -- host side of the C ABI declared in include/ilcc2_b200.h.
// Owns the handle (device scratch sized once at creation), validates arguments, stages host buffers,
// launches the five stage kernels on the caller's stream and returns results.  No CPU compute path exists:
// every compute entry point fails loudly when CUDA is unavailable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "ilcc2_internal.h"

using namespace ilcc2;

namespace {

thread_local std::string g_create_error;

constexpr int kStages = 5;
const char* kStageNames[kStages] = {"roi_filter", "cluster", "plane", "pca_gray", "board_fit"};

}  // namespace

struct ilcc2_handle {
  ilcc2_params params;
  DevParams dp;
  int device = 0;
  int max_frames = 0;
  int max_points = 0;
  int cap = 0;
  Workspace ws{};
  // owned device memory
  std::vector<void*> allocs;
  unsigned char* d_points_stage = nullptr;
  size_t points_stage_bytes = 0;
  float* d_clicks = nullptr;
  uint8_t* d_flags = nullptr;
  int8_t* d_pdelta = nullptr;
  long long* d_frame_off = nullptr;
  float* d_corners = nullptr;
  int32_t* d_status = nullptr;
  ilcc2_diag* d_diag = nullptr;
  uint32_t* d_rnd = nullptr;
  // last call
  int last_n_frames = 0;
  std::vector<long long> last_frame_off;  // host copy (empty when fixed size)
  int last_ppf = 0;
  cudaStream_t last_stream = nullptr;
  // profiling
  bool profiling = false;
  cudaEvent_t ev[kStages + 1] = {};
  // host-buffer calls: chunked H2D on a private copy stream overlapped with the kernels of the previous chunk
  cudaStream_t copy_stream = nullptr;
  static constexpr int kMaxChunks = 64;
  cudaEvent_t ev_copied[kMaxChunks] = {};
  cudaEvent_t ev_start = nullptr;
  static constexpr int kWorkStreams = 4;   // chunks are computed round-robin on these, so their kernels overlap
  cudaStream_t work_stream[kWorkStreams] = {};
  cudaEvent_t ev_work_done[kWorkStreams] = {};
  int e2e_chunk_frames = 256;
  int dev_chunks = 1;        // > 1 (ILCC2_DEV_CHUNKS): device-resident batches are cut into this many concurrent chunks
  bool ev_valid = false;
  int64_t launches = 0;
  std::string error;
};

namespace {

int fail(ilcc2_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h)
    h->error = buf;
  else
    g_create_error = buf;
  return code;
}

#define CU_TRY(h, expr)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess) return fail(h, ILCC2_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

template <typename T>
cudaError_t dev_alloc(ilcc2_handle* h, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 1);
  if (e == cudaSuccess) {
    h->allocs.push_back(q);
    *p = static_cast<T*>(q);
  }
  return e;
}

bool params_valid(const ilcc2_params& p) {
  return p.roi_half_x > 0 && p.roi_half_y > 0 && p.roi_half_z > 0 && p.cluster_tolerance > 0 &&
         p.min_cluster_size >= 1 && p.max_cluster_size >= p.min_cluster_size && p.ransac_threshold > 0 &&
         p.ransac_max_iterations >= 1 && p.ransac_probability > 0 && p.ransac_probability < 1 && p.hist_bins >= 1 &&
         p.hist_bins <= 4096 && p.gray_rate > 1 && p.huber_delta > 0 && p.grid_length > 0 && p.cells_x >= 2 &&
         p.cells_y >= 2 && p.max_solver_iterations >= 1;
}

void derive(const ilcc2_params& p, DevParams& d) {
  d.roi_hx = p.roi_half_x;
  d.roi_hy = p.roi_half_y;
  d.roi_hz = p.roi_half_z;
  d.tol2 = (float)((double)p.cluster_tolerance * (double)p.cluster_tolerance);
  d.cell = (float)(0.95 * (double)p.cluster_tolerance / std::sqrt(3.0));
  d.ncx = (int)std::floor(2.0 * p.roi_half_x / d.cell) + 2;
  d.ncy = (int)std::floor(2.0 * p.roi_half_y / d.cell) + 2;
  d.ncz = (int)std::floor(2.0 * p.roi_half_z / d.cell) + 2;
  d.min_cl = p.min_cluster_size;
  d.max_cl = p.max_cluster_size;
  d.ransac_thr = p.ransac_threshold;
  d.ransac_max_iter = p.ransac_max_iterations;
  d.ransac_log_prob = std::log(1.0 - p.ransac_probability);
  d.hist_bins = p.hist_bins;
  d.gray_rate = p.gray_rate;
  d.huber_a = p.huber_delta;
  d.huber_b = p.huber_delta * p.huber_delta;
  d.g = p.grid_length;
  d.inv_g = 1.0 / p.grid_length;
  d.W = p.cells_x;
  d.H = p.cells_y;
  d.wg2 = (double)p.cells_x * p.grid_length / 2.0;
  d.hg2 = (double)p.cells_y * p.grid_length / 2.0;
  d.min_board = p.min_board_points;
  d.max_solver_iter = p.max_solver_iterations;
  d.ncorners = (p.cells_x - 1) * (p.cells_y - 1);
}

}  // namespace

extern "C" {

int ilcc2_version(void) { return ILCC2_VERSION; }

void ilcc2_default_params(ilcc2_params* p) {
  if (!p) return;
  p->roi_half_x = 1.0f;
  p->roi_half_y = 1.5f;
  p->roi_half_z = 2.0f;
  p->cluster_tolerance = 0.12f;
  p->min_cluster_size = 100;
  p->max_cluster_size = 25000;
  p->ransac_threshold = 0.03;
  p->ransac_max_iterations = 50;
  p->ransac_probability = 0.99;
  p->hist_bins = 100;
  p->gray_rate = 2.5;
  p->huber_delta = 0.1;
  p->grid_length = 0.15;
  p->cells_x = 6;
  p->cells_y = 8;
  p->min_board_points = 0;
  p->max_solver_iterations = 50;
}

int ilcc2_set_chessboard_param(ilcc2_params* p, const char* cam_yaml_path) {
  if (!p || !cam_yaml_path) return fail(nullptr, ILCC2_ERR_INVALID_ARGUMENT, "null argument");
  std::ifstream in(cam_yaml_path);
  if (!in.is_open()) return fail(nullptr, ILCC2_ERR_IO, "can not open %s", cam_yaml_path);
  double grid = -1.0;
  int cx = -1, cy = -1;
  std::string line;
  while (std::getline(in, line)) {
    const size_t c = line.find(':');
    if (c == std::string::npos) continue;
    std::string key = line.substr(0, c);
    key.erase(0, key.find_first_not_of(" \t"));
    key.erase(key.find_last_not_of(" \t") + 1);
    const std::string val = line.substr(c + 1);
    if (key == "grid_length") grid = atof(val.c_str());
    if (key == "corner_in_x") cx = atoi(val.c_str());
    if (key == "corner_in_y") cy = atoi(val.c_str());
  }
  if (!(grid > 0) || cx < 1 || cy < 1)
    return fail(nullptr, ILCC2_ERR_IO, "%s: grid_length / corner_in_x / corner_in_y missing", cam_yaml_path);
  p->grid_length = grid;
  int a = cx + 1, b = cy + 1;  // LidarCornersEst.cpp:31-39
  if (a > b) std::swap(a, b);
  p->cells_x = a;
  p->cells_y = b;
  return ILCC2_SUCCESS;
}

int ilcc2_create(const ilcc2_params* params, int device, int max_frames, int max_points_per_frame,
                 ilcc2_handle** out) {
  if (!out) return fail(nullptr, ILCC2_ERR_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  if (max_frames < 1 || max_points_per_frame < 1)
    return fail(nullptr, ILCC2_ERR_INVALID_ARGUMENT, "max_frames and max_points_per_frame must be >= 1");
  ilcc2_params p;
  if (params)
    p = *params;
  else
    ilcc2_default_params(&p);
  if (!params_valid(p)) return fail(nullptr, ILCC2_ERR_INVALID_ARGUMENT, "invalid ilcc2_params");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, ILCC2_ERR_CUDA, "no CUDA device (%s); this library has no CPU path",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return fail(nullptr, ILCC2_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  CU_TRY(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, ILCC2_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);

  ilcc2_handle* h = new ilcc2_handle();
  h->params = p;
  derive(p, h->dp);
  if (h->dp.ncx > 1023 || h->dp.ncy > 1023 || h->dp.ncz > 1023) {
    delete h;
    return fail(nullptr, ILCC2_ERR_INVALID_ARGUMENT, "ROI / cluster_tolerance give too many hash cells");
  }
  h->device = device;
  h->max_frames = max_frames;
  h->max_points = max_points_per_frame;
  h->cap = max_points_per_frame;
  const size_t F = (size_t)max_frames, R = (size_t)h->cap;
  Workspace& w = h->ws;
  int ts = 1;
  while (ts < 2 * h->cap) ts <<= 1;
  w.cl_ts = ts;
  w.cl_stride = 2ll * ts + 3ll * h->cap;
  const bool need_global_cluster = true;  // also the overflow path of the shared-memory tables
  cudaError_t ce = cudaSuccess;
  auto A = [&](cudaError_t r) { if (ce == cudaSuccess) ce = r; };
  A(dev_alloc(h, &w.roi_pts, F * R));
  A(dev_alloc(h, &w.roi_src, F * R));
  A(dev_alloc(h, &w.roi_count, F));
  A(dev_alloc(h, &w.nn_key, F));
  A(dev_alloc(h, &w.sel_idx, F * R));
  A(dev_alloc(h, &w.sel_count, F));
  A(dev_alloc(h, &w.brd_idx, F * R));
  A(dev_alloc(h, &w.brd_count, F));
  A(dev_alloc(h, &w.tmp_idx, F * R));
  A(dev_alloc(h, &w.cl_scratch, need_global_cluster ? F * (size_t)w.cl_stride : 1));
  A(dev_alloc(h, &w.opt_yz, F * R + 1024));  // K5 prefetches one team stride (<= 8 x 64 points) ahead without clamping
  A(dev_alloc(h, &w.opt_white, F * R));
  A(dev_alloc(h, &w.opt_count, F));
  A(dev_alloc(h, &w.fit_order, F));
  A(dev_alloc(h, &w.fit_done, F));
  w.num_sms = prop.multiProcessorCount;
  w.fit_stride = (h->cap + 31) / 32 + 2ll * kFitGlobalTeam * h->cap;
  bool big_layout = h->cap > kFitBigFrame;
  if (const char* e = getenv("ILCC2_FIT_LAYOUT")) big_layout = std::string(e) == "big";  // tests: force either K5 layout
  w.fit_bits_cap = big_layout ? kFitBitsCapBig : kFitBitsCap;
  w.fit_team_cap = big_layout ? kFitTeamCapBig : kFitTeamCap;
  A(dev_alloc(h, &w.fit_scratch, h->cap > w.fit_bits_cap ? F * (size_t)w.fit_stride : 1));
  A(dev_alloc(h, &h->d_clicks, F * 3));
  A(dev_alloc(h, &h->d_flags, F));
  A(dev_alloc(h, &h->d_pdelta, F));
  A(dev_alloc(h, &h->d_frame_off, F + 1));
  A(dev_alloc(h, &h->d_corners, F * (size_t)h->dp.ncorners * 3));
  A(dev_alloc(h, &h->d_status, F));
  A(dev_alloc(h, &h->d_diag, F));
  A(dev_alloc(h, &h->d_rnd, (size_t)kRndTable));
  if (ce == cudaSuccess) {
    // PCL SampleConsensusModel draws: boost::mt19937 seeded 12345 through uniform_int<>(0, INT_MAX) = mt() >> 1
    std::vector<uint32_t> tab(kRndTable);
    std::mt19937 gen(12345u);
    for (int k = 0; k < kRndTable; ++k) tab[k] = (uint32_t)gen() >> 1;
    A(cudaMemcpy(h->d_rnd, tab.data(), sizeof(uint32_t) * kRndTable, cudaMemcpyHostToDevice));
  }
  if (ce == cudaSuccess) A(configure_kernels());
  for (int i = 0; i <= kStages && ce == cudaSuccess; ++i) A(cudaEventCreate(&h->ev[i]));
  if (ce == cudaSuccess) A(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < ilcc2_handle::kMaxChunks && ce == cudaSuccess; ++i)
    A(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
  if (ce == cudaSuccess) A(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
  for (int i = 0; i < ilcc2_handle::kWorkStreams && ce == cudaSuccess; ++i) {
    A(cudaStreamCreateWithFlags(&h->work_stream[i], cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&h->ev_work_done[i], cudaEventDisableTiming));
  }
  if (const char* e = getenv("ILCC2_E2E_CHUNK")) h->e2e_chunk_frames = atoi(e) > 0 ? atoi(e) : h->e2e_chunk_frames;
  if (const char* e = getenv("ILCC2_DEV_CHUNKS")) h->dev_chunks = atoi(e) > 0 ? atoi(e) : 1;
  if (ce != cudaSuccess) {
    const int rc = fail(nullptr, ILCC2_ERR_CUDA, "ilcc2_create: %s", cudaGetErrorString(ce));
    for (void* q : h->allocs) cudaFree(q);
    delete h;
    return rc;
  }
  w.rnd_table = h->d_rnd;
  w.cap = h->cap;
  *out = h;
  return ILCC2_SUCCESS;
}

int ilcc2_destroy(ilcc2_handle* h) {
  if (!h) return ILCC2_SUCCESS;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void* q : h->allocs) cudaFree(q);
  if (h->d_points_stage) cudaFree(h->d_points_stage);
  for (int i = 0; i <= kStages; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 0; i < ilcc2_handle::kMaxChunks; ++i)
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int i = 0; i < ilcc2_handle::kWorkStreams; ++i) {
    if (h->work_stream[i]) cudaStreamDestroy(h->work_stream[i]);
    if (h->ev_work_done[i]) cudaEventDestroy(h->ev_work_done[i]);
  }
  delete h;
  return ILCC2_SUCCESS;
}

const char* ilcc2_last_error(const ilcc2_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int ilcc2_extract_corners(ilcc2_handle* h, const void* points, int n_frames, int points_per_frame,
                          const int64_t* frame_offsets, int point_step, int off_x, int off_y, int off_z,
                          int off_intensity, int mem_kind, const float* click_xyz, const uint8_t* flags,
                          uint32_t default_flags, const int8_t* plane_index_delta, float* corners_out,
                          int32_t* status_out, ilcc2_diag* diag_out, void* stream) {
  if (!h) return ILCC2_ERR_INVALID_ARGUMENT;
  if (n_frames == 0) return ILCC2_SUCCESS;
  if (!points || !click_xyz || !corners_out || !status_out)
    return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "points, click_xyz, corners_out and status_out are required");
  if (n_frames < 0 || n_frames > h->max_frames)
    return fail(h, ILCC2_ERR_CAPACITY, "n_frames %d exceeds handle capacity %d", n_frames, h->max_frames);
  if (mem_kind != ILCC2_MEM_HOST && mem_kind != ILCC2_MEM_DEVICE)
    return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "mem_kind must be ILCC2_MEM_HOST or ILCC2_MEM_DEVICE");
  if (point_step < 16 || point_step % 4 != 0) return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "point_step must be a multiple of 4, >= 16");
  const int offs[4] = {off_x, off_y, off_z, off_intensity};
  for (int o : offs)
    if (o < 0 || o % 4 != 0 || o + 4 > point_step) return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "field offset %d invalid", o);
  const bool fast = (point_step == 16 && off_x == 0 && off_y == 4 && off_z == 8 && off_intensity == 12) ||
                    (point_step == 32 && off_x == 0 && off_y == 4 && off_z == 8 && off_intensity == 16);
  if (fast && mem_kind == ILCC2_MEM_DEVICE && (reinterpret_cast<uintptr_t>(points) & 15u) != 0)
    return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "device point buffers of 16- / 32-byte records must be 16-byte aligned (128-bit loads)");
  long long total_points = 0;
  h->last_frame_off.clear();
  if (frame_offsets) {
    for (int f = 0; f < n_frames; ++f) {
      const long long n = frame_offsets[f + 1] - frame_offsets[f];
      if (n < 0 || n > h->max_points)
        return fail(h, ILCC2_ERR_CAPACITY, "frame %d has %lld points; handle capacity %d", f, n, h->max_points);
    }
    if (frame_offsets[0] != 0) return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "frame_offsets[0] must be 0");
    total_points = frame_offsets[n_frames];
    h->last_frame_off.assign(frame_offsets, frame_offsets + n_frames + 1);
  } else {
    if (points_per_frame < 0 || points_per_frame > h->max_points)
      return fail(h, ILCC2_ERR_CAPACITY, "points_per_frame %d exceeds handle capacity %d", points_per_frame, h->max_points);
    total_points = (long long)n_frames * points_per_frame;
  }
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Workspace w = h->ws;
  const size_t bytes = (size_t)total_points * point_step;
  if (mem_kind == ILCC2_MEM_HOST) {
    if (bytes > h->points_stage_bytes) {
      if (h->d_points_stage) cudaFree(h->d_points_stage);
      h->d_points_stage = nullptr;
      h->points_stage_bytes = 0;
      const size_t want = (size_t)h->max_frames * h->max_points * point_step;
      CU_TRY(h, cudaMalloc(&h->d_points_stage, want > bytes ? want : bytes));
      h->points_stage_bytes = want > bytes ? want : bytes;
    }
    CU_TRY(h, cudaMemcpyAsync(h->d_clicks, click_xyz, sizeof(float) * 3 * n_frames, cudaMemcpyHostToDevice, s));
    if (flags) CU_TRY(h, cudaMemcpyAsync(h->d_flags, flags, n_frames, cudaMemcpyHostToDevice, s));
    if (plane_index_delta) CU_TRY(h, cudaMemcpyAsync(h->d_pdelta, plane_index_delta, n_frames, cudaMemcpyHostToDevice, s));
    w.points = h->d_points_stage;
    w.clicks = h->d_clicks;
    w.flags = flags ? h->d_flags : nullptr;
    w.pdelta = plane_index_delta ? h->d_pdelta : nullptr;
    w.corners = h->d_corners;
    w.status = h->d_status;
  } else {
    if ((reinterpret_cast<uintptr_t>(points) & 15u) != 0 && fast)
      return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "device points must be 16-byte aligned");
    w.points = static_cast<const unsigned char*>(points);
    w.clicks = click_xyz;
    w.flags = flags;
    w.pdelta = plane_index_delta;
    w.corners = corners_out;
    w.status = status_out;
  }
  if (frame_offsets) {
    CU_TRY(h, cudaMemcpyAsync(h->d_frame_off, frame_offsets, sizeof(long long) * (n_frames + 1), cudaMemcpyHostToDevice, s));
    w.frame_off = h->d_frame_off;
  } else {
    w.frame_off = nullptr;
  }
  w.pts_per_frame = points_per_frame;
  w.step = point_step;
  w.ox = off_x;
  w.oy = off_y;
  w.oz = off_z;
  w.oi = off_intensity;
  w.default_flags = default_flags;
  w.n_frames = n_frames;
  w.frame_base = 0;
  w.diag = h->d_diag;
  CU_TRY(h, cudaMemsetAsync(h->d_diag, 0, sizeof(ilcc2_diag) * n_frames, s));

  const bool prof = h->profiling;
  static const bool debug_sync = getenv("ILCC2_DEBUG_SYNC") != nullptr;  // debugging aid: sync + report per stage
  auto stage_done = [&](int i) -> cudaError_t {
    if (prof) {
      cudaError_t e = cudaEventRecord(h->ev[i], s);
      if (e != cudaSuccess) return e;
    }
    if (debug_sync && i > 0) {
      cudaError_t e = cudaStreamSynchronize(s);
      fprintf(stderr, "[ilcc2] stage %s done: %s\n", kStageNames[i - 1], cudaGetErrorString(e));
      fflush(stderr);
      return e;
    }
    return cudaSuccess;
  };
  // Host buffers: the batch is cut into chunks of frames.  Chunks are copied back to back on a private copy stream
  // and computed round-robin on private work streams, so the five kernels of chunk c overlap both the copy of the
  // following chunks and the kernels of its neighbours (a small chunk leaves most SMs idle on its own: its run
  // time is the latency of its slowest frame).  Needs page-locked host memory to actually overlap.
  // Device buffers: the batch is resident, so all chunks (dev_chunks equal ones) start at once on the work streams.
  // K5 keeps the SMs' register files full for most of its run while K2-K4 are latency-bound kernels that need few
  // resources per frame: with the chunks staggered, K1-K4 of one chunk run in the issue slots and the CTA slots that K5
  // of the chunks before it leaves free (its tail, above all), instead of in front of it with the machine mostly idle.
  // CTA / team sizes are chosen for the whole call's frames (load_frames), which are all in flight together.
  // Profiling runs and ragged host batches: one chunk on the caller's stream.
  // Chunk schedule: equal chunks of e2e_chunk_frames, except that the batch ENDS with two short chunks (3/4 and 1/4
  // of a chunk).  After the last copy only the last chunk's kernels remain, and K5's run time is the latency of its
  // slowest solve, which launch_board_fit shortens for small launches by putting more warps on each solve.
  std::vector<int> cb, cn;  // first frame / frame count of every chunk
  bool dev_chunked = false;
  if (mem_kind == ILCC2_MEM_HOST && !prof && !debug_sync && !frame_offsets && n_frames > h->e2e_chunk_frames) {
    int C = h->e2e_chunk_frames;
    while ((n_frames + C - 1) / C + 2 > ilcc2_handle::kMaxChunks) C *= 2;
    const int tail1 = std::max(1, C / 4), tail2 = std::max(1, 3 * C / 4);
    int body = n_frames - tail1 - tail2;
    int f0 = 0;
    if (body > 0) {
      const int nb = (body + C - 1) / C;
      for (int i = 0; i < nb; ++i) {  // spread the remainder over the body chunks
        const int nf = body / nb + (i < body % nb ? 1 : 0);
        cb.push_back(f0);
        cn.push_back(nf);
        f0 += nf;
      }
    }
    for (int t : {tail2, tail1}) {
      const int nf = std::min(t, n_frames - f0);
      if (nf > 0) {
        cb.push_back(f0);
        cn.push_back(nf);
        f0 += nf;
      }
    }
  } else if (mem_kind == ILCC2_MEM_DEVICE && !prof && !debug_sync && h->dev_chunks > 1 && n_frames >= 128 * h->dev_chunks) {
    const int nc = std::min(h->dev_chunks, (int)ilcc2_handle::kMaxChunks);
    for (int c = 0; c < nc; ++c) {
      const int b = (int)((long long)n_frames * c / nc), e = (int)((long long)n_frames * (c + 1) / nc);
      cb.push_back(b);
      cn.push_back(e - b);
    }
    dev_chunked = true;
  } else {
    cb.push_back(0);
    cn.push_back(n_frames);
  }
  const int n_chunks = (int)cb.size();
  if (mem_kind == ILCC2_MEM_HOST) {
    if (n_chunks == 1) {
      CU_TRY(h, cudaMemcpyAsync(h->d_points_stage, points, bytes, cudaMemcpyHostToDevice, s));
    } else {
      CU_TRY(h, cudaEventRecord(h->ev_start, s));  // everything queued on s so far (small copies, memset) comes first
      CU_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ev_start, 0));
      for (int i = 0; i < ilcc2_handle::kWorkStreams; ++i) CU_TRY(h, cudaStreamWaitEvent(h->work_stream[i], h->ev_start, 0));
      const size_t frame_bytes = (size_t)points_per_frame * point_step;
      for (int c = 0; c < n_chunks; ++c) {
        CU_TRY(h, cudaMemcpyAsync(h->d_points_stage + (size_t)cb[c] * frame_bytes,
                                  static_cast<const unsigned char*>(points) + (size_t)cb[c] * frame_bytes,
                                  (size_t)cn[c] * frame_bytes, cudaMemcpyHostToDevice, h->copy_stream));
        CU_TRY(h, cudaEventRecord(h->ev_copied[c], h->copy_stream));
      }
    }
  }
  if (dev_chunked) {
    CU_TRY(h, cudaEventRecord(h->ev_start, s));  // everything queued on s so far (memset, earlier calls) comes first
    for (int i = 0; i < ilcc2_handle::kWorkStreams; ++i) CU_TRY(h, cudaStreamWaitEvent(h->work_stream[i], h->ev_start, 0));
  }
  CU_TRY(h, stage_done(0));
  for (int c = 0; c < n_chunks; ++c) {
    w.frame_base = cb[c];
    w.n_frames = cn[c];
    w.load_frames = dev_chunked ? n_frames : cn[c];
    cudaStream_t cs = n_chunks > 1 ? h->work_stream[c % ilcc2_handle::kWorkStreams] : s;
    if (n_chunks > 1 && !dev_chunked) CU_TRY(h, cudaStreamWaitEvent(cs, h->ev_copied[c], 0));
    launch_roi_filter(w, h->dp, cs);
    if (n_chunks == 1) CU_TRY(h, stage_done(1));
    launch_cluster(w, h->dp, cs);
    if (n_chunks == 1) CU_TRY(h, stage_done(2));
    launch_plane(w, h->dp, cs);
    if (n_chunks == 1) CU_TRY(h, stage_done(3));
    launch_pca_gray(w, h->dp, cs);
    if (n_chunks == 1) CU_TRY(h, stage_done(4));
    launch_board_fit(w, h->dp, cs);
    if (n_chunks == 1) CU_TRY(h, stage_done(5));
    CU_TRY(h, cudaGetLastError());  // a launch failure is reported before the next chunk is queued
    h->launches += kStages + 1 + (h->cap > h->ws.fit_bits_cap ? 1 : 0) + 1;
  }
  if (n_chunks > 1) {  // join the work streams back into the caller's stream
    for (int i = 0; i < ilcc2_handle::kWorkStreams; ++i) {
      CU_TRY(h, cudaEventRecord(h->ev_work_done[i], h->work_stream[i]));
      CU_TRY(h, cudaStreamWaitEvent(s, h->ev_work_done[i], 0));
    }
  }
  h->ev_valid = prof;
  CU_TRY(h, cudaGetLastError());

  h->last_n_frames = n_frames;
  h->last_ppf = points_per_frame;
  h->last_stream = s;
  const size_t cbytes = sizeof(float) * 3 * (size_t)h->dp.ncorners * n_frames;
  if (mem_kind == ILCC2_MEM_HOST) {
    CU_TRY(h, cudaMemcpyAsync(corners_out, h->d_corners, cbytes, cudaMemcpyDeviceToHost, s));
    CU_TRY(h, cudaMemcpyAsync(status_out, h->d_status, sizeof(int32_t) * n_frames, cudaMemcpyDeviceToHost, s));
    if (diag_out) CU_TRY(h, cudaMemcpyAsync(diag_out, h->d_diag, sizeof(ilcc2_diag) * n_frames, cudaMemcpyDeviceToHost, s));
    CU_TRY(h, cudaStreamSynchronize(s));
  } else if (diag_out) {
    CU_TRY(h, cudaMemcpyAsync(diag_out, h->d_diag, sizeof(ilcc2_diag) * n_frames, cudaMemcpyDeviceToDevice, s));
  }
  return ILCC2_SUCCESS;
}

int ilcc2_fetch_indices(ilcc2_handle* h, int frame, int which, int32_t* out_host, int cap) {
  if (!h || !out_host || frame < 0 || frame >= h->last_n_frames || which < 0 || which > 2) return -ILCC2_ERR_INVALID_ARGUMENT;
  if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->last_stream) != cudaSuccess) return -ILCC2_ERR_CUDA;
  const Workspace& w = h->ws;
  const size_t R = (size_t)h->cap;
  int n = 0;
  const int* cnt = which == 0 ? w.roi_count : which == 1 ? w.sel_count : w.brd_count;
  if (cudaMemcpy(&n, cnt + frame, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -ILCC2_ERR_CUDA;
  if (n <= 0) return 0;
  const int m = n < cap ? n : cap;
  std::vector<int> src(R), idx;
  int nroi = 0;
  cudaMemcpy(&nroi, w.roi_count + frame, sizeof(int), cudaMemcpyDeviceToHost);
  if (cudaMemcpy(src.data(), w.roi_src + frame * R, sizeof(int) * (size_t)nroi, cudaMemcpyDeviceToHost) != cudaSuccess)
    return -ILCC2_ERR_CUDA;
  if (which == 0) {
    for (int i = 0; i < m; ++i) out_host[i] = src[i];
    return n;
  }
  idx.resize(n);
  const int* list = which == 1 ? w.sel_idx : w.brd_idx;
  if (cudaMemcpy(idx.data(), list + frame * R, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess)
    return -ILCC2_ERR_CUDA;
  for (int i = 0; i < m; ++i) out_host[i] = src[idx[i]];
  return n;
}

int ilcc2_fetch_board_points(ilcc2_handle* h, int frame, float* yz_out_host, uint8_t* white_out_host, int cap) {
  if (!h || !yz_out_host || !white_out_host || frame < 0 || frame >= h->last_n_frames) return -ILCC2_ERR_INVALID_ARGUMENT;
  if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->last_stream) != cudaSuccess) return -ILCC2_ERR_CUDA;
  const Workspace& w = h->ws;
  const size_t R = (size_t)h->cap;
  int n = 0;
  if (cudaMemcpy(&n, w.opt_count + frame, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -ILCC2_ERR_CUDA;
  const int m = n < cap ? n : cap;
  if (m > 0) {
    if (cudaMemcpy(yz_out_host, w.opt_yz + frame * R, sizeof(float2) * (size_t)m, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(white_out_host, w.opt_white + frame * R, (size_t)m, cudaMemcpyDeviceToHost) != cudaSuccess)
      return -ILCC2_ERR_CUDA;
  }
  return n;
}

int ilcc2_set_profiling(ilcc2_handle* h, int enabled) {
  if (!h) return ILCC2_ERR_INVALID_ARGUMENT;
  h->profiling = enabled != 0;
  return ILCC2_SUCCESS;
}

int ilcc2_stage_times(ilcc2_handle* h, float* ms_out, const char** names_out, int cap) {
  if (!h) return -ILCC2_ERR_INVALID_ARGUMENT;
  if (!h->ev_valid) return 0;
  if (cudaSetDevice(h->device) != cudaSuccess || cudaEventSynchronize(h->ev[kStages]) != cudaSuccess) return -ILCC2_ERR_CUDA;
  for (int i = 0; i < kStages && i < cap; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
    if (ms_out) ms_out[i] = ms;
    if (names_out) names_out[i] = kStageNames[i];
  }
  return kStages;
}

int ilcc2_set_gather_targets(ilcc2_handle* h, int n_targets, float* const* corners_targets,
                             int32_t* const* status_targets, int slot_base) {
  if (!h) return ILCC2_ERR_INVALID_ARGUMENT;
  if (n_targets < 0 || n_targets > kMaxGatherTargets || slot_base < 0 ||
      (n_targets > 0 && (!corners_targets || !status_targets)))
    return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "set_gather_targets: 0..%d targets, non-null pointer arrays", kMaxGatherTargets);
  for (int i = 0; i < n_targets; ++i)
    if (!corners_targets[i] || !status_targets[i])
      return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "set_gather_targets: target %d is null", i);
  Workspace& w = h->ws;
  w.n_targets = n_targets;
  w.gather_slot_base = slot_base;
  for (int i = 0; i < kMaxGatherTargets; ++i) {
    w.gather_corners[i] = i < n_targets ? corners_targets[i] : nullptr;
    w.gather_status[i] = i < n_targets ? status_targets[i] : nullptr;
  }
  return ILCC2_SUCCESS;
}

int64_t ilcc2_launch_count(const ilcc2_handle* h) { return h ? h->launches : 0; }

int ilcc2_save_corners2txt(const float* corners_xyz, int n_corners, const char* path) {
  if (!corners_xyz || !path || n_corners < 0) return ILCC2_ERR_INVALID_ARGUMENT;
  std::ofstream out(path, std::ios_base::trunc);
  if (!out.is_open()) return ILCC2_ERR_IO;
  for (int i = 0; i < n_corners; ++i)
    out << corners_xyz[3 * i] << " " << corners_xyz[3 * i + 1] << " " << corners_xyz[3 * i + 2] << std::endl;
  out.close();
  return out.fail() ? ILCC2_ERR_IO : ILCC2_SUCCESS;
}

void ilcc2_shard_range(int n_frames, int world_size, int rank, int* begin, int* count) {
  if (world_size < 1) world_size = 1;
  const long long b = (long long)n_frames * rank / world_size;
  const long long e = (long long)n_frames * (rank + 1) / world_size;
  if (begin) *begin = (int)b;
  if (count) *count = (int)(e - b);
}

}  // extern "C"
