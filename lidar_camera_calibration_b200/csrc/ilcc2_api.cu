// ilcc2_api.cu -- host side of the C ABI declared in include/ilcc2_b200.h.
// Owns the handle (device scratch sized once at creation), validates arguments, stages host buffers,
// launches the five stage kernels on the caller's stream and returns results.  No CPU compute path exists:
// every compute entry point fails loudly when CUDA is unavailable.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>
#include <xmmintrin.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
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

constexpr int kStages = 4;
const char* kStageNames[kStages] = {"roi_filter", "cluster", "plane_pca", "board_fit"};

}  // namespace

// A few threads that copy slices of a pageable chunk into page-locked staging memory (one memcpy stream per core
// moves ~10 GB/s; PCIe 5 x16 takes ~55 GB/s).  Created on the first pageable host call.
class CopyPool {
 public:
  explicit CopyPool(int n) : n_(n < 1 ? 1 : n) {
    for (int i = 0; i < n_; ++i) th_.emplace_back([this, i] { run(i); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> l(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int threads() const { return n_; }
  // repack = false: plain copy of `bytes` bytes.  repack = true: `bytes` counts 32-byte PointCloud2 / pcl::PointXYZI
  // records in src (x, y, z @0/4/8, intensity @16); dst receives packed 16-byte (x, y, z, intensity) records, written
  // with streaming stores -- the padding and the ring field never cross PCIe.
  void copy(void* dst, const void* src, size_t bytes, bool repack = false) {
    if (!repack && (bytes < (size_t)(1 << 20) || n_ == 1)) {
      memcpy(dst, src, bytes);
      return;
    }
    {
      std::lock_guard<std::mutex> l(m_);
      dst_ = static_cast<unsigned char*>(dst);
      src_ = static_cast<const unsigned char*>(src);
      bytes_ = bytes;
      repack_ = repack;
      pending_ = n_;
      ++gen_;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> l(m_);
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void run(int i) {
    unsigned long long seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> l(m_);
      cv_.wait(l, [&] { return gen_ != seen; });
      seen = gen_;
      if (stop_) return;
      unsigned char* d = dst_;
      const unsigned char* sp = src_;
      const bool rp = repack_;
      const size_t per = ((bytes_ + n_ - 1) / n_ + 4095) & ~(size_t)4095, b = std::min(bytes_, per * i),
                   e = std::min(bytes_, per * (i + 1));
      l.unlock();
      if (e > b) {
        if (!rp) {
          memcpy(d + b, sp + b, e - b);
        } else {
          const float* in = reinterpret_cast<const float*>(sp + (b >> 5 << 5));
          float* out = reinterpret_cast<float*>(d + (b >> 5 << 4));
          for (size_t k = 0, n = (e >> 5) - (b >> 5); k < n; ++k, in += 8, out += 4) {
            const __m128 a = _mm_loadu_ps(in);                               // x y z pad
            const __m128 zi = _mm_shuffle_ps(a, _mm_load_ss(in + 4), _MM_SHUFFLE(0, 0, 2, 2));  // z z i i
            _mm_stream_ps(out, _mm_shuffle_ps(a, zi, _MM_SHUFFLE(2, 0, 1, 0)));  // x y z i
          }
          _mm_sfence();
        }
      }
      l.lock();
      if (--pending_ == 0) done_.notify_one();
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  unsigned long long gen_ = 0;
  bool stop_ = false;
  unsigned char* dst_ = nullptr;
  const unsigned char* src_ = nullptr;
  size_t bytes_ = 0;
  bool repack_ = false;
  int pending_ = 0;
};

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
  // host-buffer calls: device staging ring (K1 is the only reader of the staged records) and, for pageable caller
  // memory, a page-locked host ring the caller's chunks are copied through by the pool's threads
  static constexpr int kDevSlots = 4, kPinSlots = 3;
  unsigned char* d_stage[kDevSlots] = {};
  unsigned char* h_stage[kPinSlots] = {};
  size_t slot_bytes = 0;
  int slot_frames = 0;
  cudaEvent_t ev_k1_done[kDevSlots] = {};
  cudaEvent_t ev_h2d_done[kPinSlots] = {};
  bool dev_slot_used[kDevSlots] = {};
  bool pin_slot_used[kPinSlots] = {};
  CopyPool* pool = nullptr;
  unsigned char* d_points_big = nullptr;  // profiling / debug runs of host batches beyond the ring: one buffer, grown on demand
  size_t points_big_bytes = 0;
  int roi_cap = 0, overflow_frames = 0;
  struct ilcc2_gather* gather = nullptr;        // ilcc2_gather_setup_nccl
  const float* last_corners_dev = nullptr;      // outputs of the last ILCC2_MEM_DEVICE call (NCCL fallback of the gather)
  const int32_t* last_status_dev = nullptr;
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
  int e2e_chunk_frames = 64;
  int lanes = 1;             // ilcc2_set_lanes: handles that take consecutive device-resident batches in turn
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
  d.inv_cell = (float)(1.0 / (double)d.cell);
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
  return ilcc2_create_sized(params, device, max_frames, max_points_per_frame, 0, 0, out);
}

int ilcc2_create_sized(const ilcc2_params* params, int device, int max_frames, int max_points_per_frame,
                       int roi_capacity, int overflow_frames, ilcc2_handle** out) {
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
  if (prop.major != 10 || prop.minor != 0)  // sm_100a code is architecture-specific: no other 10.x part can run it
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
  // ROI-sized arrays hold roi_capacity entries per frame (default: every point of the frame may survive setROI);
  // a frame whose ROI is larger reports ILCC2_ROI_CAPACITY.  The global-memory clustering tables exist for a pool of
  // overflow_frames frames (default: every frame when the ROI can exceed the shared-memory tables, else an eighth).
  h->cap = roi_capacity > 0 ? std::min(roi_capacity, max_points_per_frame) : max_points_per_frame;
  h->roi_cap = h->cap;
  h->overflow_frames = overflow_frames > 0 ? std::min(overflow_frames, max_frames)
                                           : (h->cap > kClusterSmemPts ? max_frames : std::max(std::min(max_frames, 8), max_frames / 8));
  const size_t F = (size_t)max_frames, R = (size_t)h->cap;
  Workspace& w = h->ws;
  int ts = 1;
  while (ts < 2 * h->cap) ts <<= 1;
  w.cl_ts = ts;
  w.cl_stride = 2ll * ts + 3ll * h->cap;
  w.cl_pool_slots = h->overflow_frames;
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
  A(dev_alloc(h, &w.cl_scratch, (size_t)h->overflow_frames * (size_t)w.cl_stride));
  A(dev_alloc(h, &w.cl_pool_next, 1));
  A(dev_alloc(h, &w.opt_yz, F * R + 1024));  // K5 prefetches one team stride (<= 8 x 64 points) ahead without clamping
  A(dev_alloc(h, &w.opt_white, F * R));
  A(dev_alloc(h, &w.opt_count, F));
  A(dev_alloc(h, &w.fit_order, F));
  A(dev_alloc(h, &w.fit_done, F));
  w.num_sms = prop.multiProcessorCount;
  bool big_layout = h->max_points > kFitBigFrame;
  if (const char* e = getenv("ILCC2_FIT_LAYOUT")) big_layout = std::string(e) == "big";  // tests: force either K5 layout
  w.fit_bits_cap = big_layout ? kFitBitsCapBig : kFitBitsCap;
  w.fit_team_cap = big_layout ? kFitTeamCapBig : kFitTeamCap;
  A(dev_alloc(h, &w.fit_pool_next, 1));
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
  // device staging ring for host-buffer calls: kDevSlots chunks of 32-byte records (whole batches of up to
  // kDevSlots chunks fit at once, larger ones stream through it)
  h->slot_frames = std::min(max_frames, h->e2e_chunk_frames);
  h->slot_bytes = (size_t)h->slot_frames * (size_t)max_points_per_frame * 32;
  for (int i = 0; i < ilcc2_handle::kDevSlots && ce == cudaSuccess; ++i) {
    if (i == 0 || max_frames > h->slot_frames) {  // batches beyond one slot stream through all of them
      void* q = nullptr;
      A(cudaMalloc(&q, h->slot_bytes));
      h->d_stage[i] = static_cast<unsigned char*>(q);
    }
    A(cudaEventCreateWithFlags(&h->ev_k1_done[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < ilcc2_handle::kPinSlots && ce == cudaSuccess; ++i)
    A(cudaEventCreateWithFlags(&h->ev_h2d_done[i], cudaEventDisableTiming));
  if (const char* e = getenv("ILCC2_DEV_CHUNKS")) h->dev_chunks = atoi(e) > 0 ? atoi(e) : 1;
  if (ce != cudaSuccess) {
    const int rc = fail(nullptr, ILCC2_ERR_CUDA, "ilcc2_create: %s", cudaGetErrorString(ce));
    for (void* q : h->allocs) cudaFree(q);
    for (unsigned char* q : h->d_stage)
      if (q) cudaFree(q);
    delete h;
    return rc;
  }
  w.rnd_table = h->d_rnd;
  w.cap = h->cap;
  w.fit_scratch = reinterpret_cast<uint32_t*>(w.cl_scratch);  // K5's global tables reuse the pool K2 has finished with
  w.fit_stride = w.cl_stride;
  *out = h;
  return ILCC2_SUCCESS;
}

int ilcc2_destroy(ilcc2_handle* h) {
  if (!h) return ILCC2_SUCCESS;
  ilcc2_gather_teardown(h);
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void* q : h->allocs) cudaFree(q);
  delete h->pool;
  for (unsigned char* q : h->d_stage)
    if (q) cudaFree(q);
  for (unsigned char* q : h->h_stage)
    if (q) cudaFreeHost(q);
  if (h->d_points_big) cudaFree(h->d_points_big);
  for (cudaEvent_t e : h->ev_k1_done)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_h2d_done)
    if (e) cudaEventDestroy(e);
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
  const bool host = mem_kind == ILCC2_MEM_HOST;
  if (host) {
    CU_TRY(h, cudaMemcpyAsync(h->d_clicks, click_xyz, sizeof(float) * 3 * n_frames, cudaMemcpyHostToDevice, s));
    if (flags) CU_TRY(h, cudaMemcpyAsync(h->d_flags, flags, n_frames, cudaMemcpyHostToDevice, s));
    if (plane_index_delta) CU_TRY(h, cudaMemcpyAsync(h->d_pdelta, plane_index_delta, n_frames, cudaMemcpyHostToDevice, s));
    w.clicks = h->d_clicks;
    w.flags = flags ? h->d_flags : nullptr;
    w.pdelta = plane_index_delta ? h->d_pdelta : nullptr;
    w.corners = h->d_corners;
    w.status = h->d_status;
  } else {
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
  CU_TRY(h, cudaMemsetAsync(w.cl_pool_next, 0, sizeof(int), s));
  CU_TRY(h, cudaMemsetAsync(w.fit_pool_next, 0, sizeof(int), s));

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
  // Host buffers: the batch is cut into chunks of frames that stream through the handle's staging ring.  A chunk is
  // copied host -> device on a private copy stream into one of kDevSlots device slots (reused once K1 of the chunk
  // that held it has run: K1 is the only reader of the staged records) and computed on one of the private work
  // streams, so the kernels of chunk c overlap the copies of the chunks behind it and the kernels of its neighbours.
  // Page-locked caller memory is copied from directly; pageable memory (what a ROS node holds: the `data` vector of a
  // sensor_msgs::PointCloud2, get_lidar_corners.cpp:136-164) first goes through kPinSlots page-locked slots, filled
  // by the copy pool's threads while the DMA of the previous chunk is in flight.
  // Chunk schedule: equal chunks of e2e_chunk_frames, except that the batch ENDS with two short chunks (3/4 and 1/4
  // of a chunk).  After the last copy only the last chunk's kernels remain, and K5's run time is the latency of its
  // slowest solve, which launch_board_fit shortens for small launches by putting more warps on each solve.
  // Device buffers: one chunk on the caller's stream (ILCC2_DEV_CHUNKS > 1: concurrent chunks on the work streams).
  // Profiling / debug runs: one chunk on the caller's stream, per-stage events.
  std::vector<int> cb, cn;  // first frame / frame count of every chunk
  bool dev_chunked = false;
  const bool serial = prof || debug_sync;
  const size_t ring_bytes = h->slot_bytes * (size_t)ilcc2_handle::kDevSlots;
  const int chunk_frames = point_step <= 32 ? h->slot_frames : std::max(1, (int)((long long)h->slot_frames * 32 / point_step));
  if (host && !serial && n_frames > chunk_frames) {
    int C = chunk_frames;
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
  } else if (!host && !serial && h->dev_chunks > 1 && n_frames >= 128 * h->dev_chunks) {
    const int nc = h->dev_chunks;
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
  const bool multi = n_chunks > 1;
  // record offset (in point records) of a frame within the caller's buffer
  auto rec_off = [&](int f) -> long long { return frame_offsets ? (long long)frame_offsets[f] : (long long)f * points_per_frame; };
  bool pageable = false, repack = false;
  unsigned char* big = nullptr;
  if (host) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, points) != cudaSuccess) {
      cudaGetLastError();
      pageable = true;
    } else {
      pageable = pa.type == cudaMemoryTypeUnregistered;
    }
    static const bool driver_staging = getenv("ILCC2_DRIVER_STAGING") != nullptr;  // A/B: let cudaMemcpyAsync stage pageable memory
    if (driver_staging) pageable = false;
    if (!multi && bytes > h->slot_bytes) {
      // one chunk that does not fit a slot (profiling / debug runs of large host batches): a buffer grown on demand
      if (bytes > h->points_big_bytes) {
        if (h->d_points_big) cudaFree(h->d_points_big);
        h->d_points_big = nullptr;
        h->points_big_bytes = 0;
        CU_TRY(h, cudaMalloc(&h->d_points_big, bytes));
        h->points_big_bytes = bytes;
      }
      big = h->d_points_big;
    }
    if (pageable && multi && !h->h_stage[0]) {  // first pageable call: page-locked ring + copy threads (one-time)
      for (int i = 0; i < ilcc2_handle::kPinSlots; ++i) {
        void* q = nullptr;
        CU_TRY(h, cudaHostAlloc(&q, h->slot_bytes, cudaHostAllocDefault));
        h->h_stage[i] = static_cast<unsigned char*>(q);
      }
      int nt = (int)std::thread::hardware_concurrency();
      cpu_set_t set;
      if (sched_getaffinity(0, sizeof(set), &set) == 0) nt = CPU_COUNT(&set);
      nt = std::min(nt, 12);
      if (const char* e = getenv("ILCC2_STAGE_THREADS")) nt = atoi(e);
      h->pool = new CopyPool(std::max(1, std::min(nt, 64)));
    }
    // 32-byte records in pageable memory are compacted to packed XYZI on their way through the page-locked ring (the
    // host pass exists anyway): half the PCIe bytes, and K1 reads packed records.  Same floats, same results.
    repack = pageable && multi && point_step == 32 && off_x == 0 && off_y == 4 && off_z == 8 && off_intensity == 16 &&
             getenv("ILCC2_NO_REPACK") == nullptr;
    if (repack) {
      w.step = 16;
      w.oi = 12;
    }
  }
  const int dev_step = repack ? 16 : point_step;
  if (multi) {
    CU_TRY(h, cudaEventRecord(h->ev_start, s));  // everything queued on s so far (small copies, memsets) comes first
    if (host) CU_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->ev_start, 0));
    for (int i = 0; i < ilcc2_handle::kWorkStreams; ++i) CU_TRY(h, cudaStreamWaitEvent(h->work_stream[i], h->ev_start, 0));
  }
  CU_TRY(h, stage_done(0));
  for (int c = 0; c < n_chunks; ++c) {
    w.frame_base = cb[c];
    w.n_frames = cn[c];
    w.load_frames = dev_chunked ? n_frames : cn[c];
    if (!host) w.load_frames *= h->lanes;  // other handles' batches are in flight next to this one
    cudaStream_t cs = multi ? h->work_stream[c % ilcc2_handle::kWorkStreams] : s;
    const int ds = c % ilcc2_handle::kDevSlots;
    if (host) {
      const long long r0 = rec_off(cb[c]), r1 = frame_offsets ? (long long)frame_offsets[cb[c] + cn[c]] : rec_off(cb[c] + cn[c]);
      const size_t cbytes = (size_t)(r1 - r0) * point_step;
      const unsigned char* src = static_cast<const unsigned char*>(points) + (size_t)r0 * point_step;
      unsigned char* dst = big ? big : h->d_stage[ds];
      cudaStream_t xs = multi ? h->copy_stream : s;
      if (multi && h->dev_slot_used[ds]) CU_TRY(h, cudaStreamWaitEvent(xs, h->ev_k1_done[ds], 0));  // slot free again?
      if (pageable && multi) {
        const int ps = c % ilcc2_handle::kPinSlots;
        if (h->pin_slot_used[ps]) CU_TRY(h, cudaEventSynchronize(h->ev_h2d_done[ps]));
        h->pool->copy(h->h_stage[ps], src, cbytes, repack);
        CU_TRY(h, cudaMemcpyAsync(dst, h->h_stage[ps], repack ? cbytes / 2 : cbytes, cudaMemcpyHostToDevice, xs));
        CU_TRY(h, cudaEventRecord(h->ev_h2d_done[ps], xs));
        h->pin_slot_used[ps] = true;
      } else {
        CU_TRY(h, cudaMemcpyAsync(dst, src, cbytes, cudaMemcpyHostToDevice, xs));
      }
      if (multi) {
        CU_TRY(h, cudaEventRecord(h->ev_copied[c % ilcc2_handle::kMaxChunks], xs));
        CU_TRY(h, cudaStreamWaitEvent(cs, h->ev_copied[c % ilcc2_handle::kMaxChunks], 0));
      }
      w.points = dst - (size_t)r0 * dev_step;  // kernels index records by their offset within the whole batch
    }
    launch_roi_filter(w, h->dp, cs);
    if (host && multi) {
      CU_TRY(h, cudaEventRecord(h->ev_k1_done[ds], cs));
      h->dev_slot_used[ds] = true;
    }
    if (!multi) CU_TRY(h, stage_done(1));
    launch_cluster(w, h->dp, cs);
    if (!multi) CU_TRY(h, stage_done(2));
    launch_plane_pca(w, h->dp, cs);
    if (!multi) CU_TRY(h, stage_done(3));
    launch_board_fit(w, h->dp, cs);
    if (!multi) CU_TRY(h, stage_done(4));
    CU_TRY(h, cudaGetLastError());  // a launch failure is reported before the next chunk is queued
    h->launches += 6 + (h->cap > h->ws.fit_bits_cap ? 1 : 0);  // K1, K2 (2 instantiations), K3+K4, fit_order, K5 (1 or 2)
  }
  if (multi) {  // join the work streams back into the caller's stream
    for (int i = 0; i < ilcc2_handle::kWorkStreams; ++i) {
      CU_TRY(h, cudaEventRecord(h->ev_work_done[i], h->work_stream[i]));
      CU_TRY(h, cudaStreamWaitEvent(s, h->ev_work_done[i], 0));
    }
  }
  h->ev_valid = prof && !multi;
  CU_TRY(h, cudaGetLastError());
  (void)ring_bytes;

  h->last_n_frames = n_frames;
  h->last_ppf = points_per_frame;
  h->last_stream = s;
  h->last_corners_dev = host ? nullptr : corners_out;
  h->last_status_dev = host ? nullptr : status_out;
  const size_t cbytes = sizeof(float) * 3 * (size_t)h->dp.ncorners * n_frames;
  if (host) {
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

int ilcc2_set_lanes(ilcc2_handle* h, int lanes) {
  if (!h) return ILCC2_ERR_INVALID_ARGUMENT;
  if (lanes < 1 || lanes > 64) return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "lanes must be in [1, 64]");
  h->lanes = lanes;
  return ILCC2_SUCCESS;
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
                             int32_t* const* status_targets, int slot_base, int target_kind) {
  if (!h) return ILCC2_ERR_INVALID_ARGUMENT;
  if (n_targets < 0 || n_targets > kMaxGatherTargets || slot_base < 0 ||
      (n_targets > 0 && (!corners_targets || !status_targets)))
    return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "set_gather_targets: 0..%d targets, non-null pointer arrays", kMaxGatherTargets);
  if (target_kind != ILCC2_GATHER_PEER && target_kind != ILCC2_GATHER_MULTICAST)
    return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "set_gather_targets: target_kind must be ILCC2_GATHER_PEER or ILCC2_GATHER_MULTICAST");
  if (target_kind == ILCC2_GATHER_MULTICAST && n_targets != 1)
    return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "set_gather_targets: a multicast mapping is ONE target that reaches every rank");
  for (int i = 0; i < n_targets; ++i)
    if (!corners_targets[i] || !status_targets[i])
      return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "set_gather_targets: target %d is null", i);
  Workspace& w = h->ws;
  w.n_targets = n_targets;
  w.gather_multicast = (n_targets > 0 && target_kind == ILCC2_GATHER_MULTICAST) ? 1 : 0;
  w.gather_slot_base = slot_base;
  for (int i = 0; i < kMaxGatherTargets; ++i) {
    w.gather_corners[i] = i < n_targets ? corners_targets[i] : nullptr;
    w.gather_status[i] = i < n_targets ? status_targets[i] : nullptr;
  }
  return ILCC2_SUCCESS;
}

// ---- multi-GPU from a C / C++ host: gather buffers over an existing NCCL communicator ------------------------------
// NCCL is resolved from the host process at run time (the host links the NCCL it created the communicator with); this
// library has no link-time dependency on it.
namespace {
struct NcclApi {
  int (*CommCount)(void*, int*) = nullptr;
  int (*CommUserRank)(void*, int*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi load_nccl() {
  NcclApi a;
  void* self = dlopen(nullptr, RTLD_NOW | RTLD_GLOBAL);
  auto sym = [&](const char* n) -> void* {
    void* p = self ? dlsym(self, n) : nullptr;
    if (!p) p = dlsym(RTLD_DEFAULT, n);
    return p;
  };
  a.CommCount = reinterpret_cast<decltype(a.CommCount)>(sym("ncclCommCount"));
  a.CommUserRank = reinterpret_cast<decltype(a.CommUserRank)>(sym("ncclCommUserRank"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  a.ok = a.CommCount && a.CommUserRank && a.AllGather && a.AllReduce;
  return a;
}
constexpr int kNcclChar = 0, kNcclInt32 = 2, kNcclMin = 3;  // ncclDataType_t / ncclRedOp_t values (nccl.h, stable ABI)
struct PeerInfo {  // what every rank publishes about its gather allocation
  cudaIpcMemHandle_t ipc;
  unsigned long long ptr;
  long long pid;
  int device;
  int pad;
};
}  // namespace

struct ilcc2_gather {
  NcclApi nccl;
  void* comm = nullptr;
  int rank = 0, world = 1, total_frames = 0, slot_base = 0;
  bool fused = false;
  int parity = 0;
  unsigned char* local = nullptr;               // this rank's allocation: [2] x (corners [total][nc][3] f32, status [total] i32)
  size_t half_bytes = 0, corners_bytes = 0;
  unsigned char* peer_base[kMaxGatherTargets] = {};
  bool peer_opened[kMaxGatherTargets] = {};
  int* d_flag = nullptr;
};

int ilcc2_gather_setup_nccl(ilcc2_handle* h, void* nccl_comm, int total_frames, int slot_base) {
  if (!h || !nccl_comm || total_frames < 1 || slot_base < 0) return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "gather_setup_nccl: bad argument");
  if (h->gather) return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "gather_setup_nccl: already set up (ilcc2_gather_teardown first)");
  CU_TRY(h, cudaSetDevice(h->device));
  ilcc2_gather* G = new ilcc2_gather();
  G->nccl = load_nccl();
  if (!G->nccl.ok) {
    delete G;
    return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "gather_setup_nccl: the host process exports no NCCL (ncclAllGather / ncclCommCount not found)");
  }
  G->comm = nccl_comm;
  if (G->nccl.CommCount(nccl_comm, &G->world) != 0 || G->nccl.CommUserRank(nccl_comm, &G->rank) != 0 || G->world < 1 ||
      G->world > kMaxGatherTargets) {
    delete G;
    return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "gather_setup_nccl: communicator of 1..%d ranks expected", kMaxGatherTargets);
  }
  G->total_frames = total_frames;
  G->slot_base = slot_base;
  G->corners_bytes = sizeof(float) * 3 * (size_t)h->dp.ncorners * total_frames;
  G->half_bytes = (G->corners_bytes + sizeof(int32_t) * (size_t)total_frames + 255) & ~(size_t)255;
  cudaError_t ce = cudaMalloc(reinterpret_cast<void**>(&G->local), 2 * G->half_bytes);
  if (ce == cudaSuccess) ce = cudaMemset(G->local, 0, 2 * G->half_bytes);
  PeerInfo* d_info = nullptr;
  if (ce == cudaSuccess) ce = cudaMalloc(reinterpret_cast<void**>(&d_info), sizeof(PeerInfo) * (G->world + 1));
  if (ce == cudaSuccess) ce = cudaMalloc(reinterpret_cast<void**>(&G->d_flag), 2 * sizeof(int));
  if (ce != cudaSuccess) {
    if (G->local) cudaFree(G->local);
    if (d_info) cudaFree(d_info);
    delete G;
    return fail(h, ILCC2_ERR_CUDA, "gather_setup_nccl: %s", cudaGetErrorString(ce));
  }
  // publish (IPC handle, raw pointer, pid, device) of the local allocation to every rank
  PeerInfo mine{};
  int good = cudaIpcGetMemHandle(&mine.ipc, G->local) == cudaSuccess ? 1 : 0;
  cudaGetLastError();
  mine.ptr = reinterpret_cast<unsigned long long>(G->local);
  mine.pid = (long long)getpid();
  mine.device = h->device;
  std::vector<PeerInfo> all(G->world);
  cudaMemcpy(d_info + G->world, &mine, sizeof(PeerInfo), cudaMemcpyHostToDevice);
  int rc = G->nccl.AllGather(d_info + G->world, d_info, sizeof(PeerInfo), kNcclChar, nccl_comm, nullptr);
  if (rc == 0 && cudaStreamSynchronize(nullptr) == cudaSuccess)
    cudaMemcpy(all.data(), d_info, sizeof(PeerInfo) * G->world, cudaMemcpyDeviceToHost);
  else
    good = 0;
  for (int r = 0; r < G->world && good; ++r) {
    if (r == G->rank) {
      G->peer_base[r] = G->local;
    } else if (all[r].pid == mine.pid) {  // same process, another device: the pointer itself, with peer access on
      int can = 0;
      cudaDeviceCanAccessPeer(&can, h->device, all[r].device);
      if (can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
        cudaGetLastError();
      }
      if (can) G->peer_base[r] = reinterpret_cast<unsigned char*>(all[r].ptr); else good = 0;
    } else {
      void* q = nullptr;
      if (cudaIpcOpenMemHandle(&q, all[r].ipc, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
        G->peer_base[r] = static_cast<unsigned char*>(q);
        G->peer_opened[r] = true;
      } else {
        cudaGetLastError();
        good = 0;
      }
    }
  }
  // every rank must have mapped every peer, else all fall back to ncclAllGather
  cudaMemcpy(G->d_flag, &good, sizeof(int), cudaMemcpyHostToDevice);
  int agreed = 0;
  if (G->nccl.AllReduce(G->d_flag, G->d_flag + 1, 1, kNcclInt32, kNcclMin, nccl_comm, nullptr) == 0 &&
      cudaStreamSynchronize(nullptr) == cudaSuccess)
    cudaMemcpy(&agreed, G->d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_info);
  G->fused = agreed == 1;
  h->gather = G;
  if (!G->fused) {
    for (int r = 0; r < G->world; ++r)
      if (G->peer_opened[r]) {
        cudaIpcCloseMemHandle(G->peer_base[r]);
        G->peer_opened[r] = false;
      }
    return ilcc2_set_gather_targets(h, 0, nullptr, nullptr, 0, ILCC2_GATHER_PEER);
  }
  float* ct[kMaxGatherTargets];
  int32_t* st[kMaxGatherTargets];
  for (int r = 0; r < G->world; ++r) {
    ct[r] = reinterpret_cast<float*>(G->peer_base[r]);
    st[r] = reinterpret_cast<int32_t*>(G->peer_base[r] + G->corners_bytes);
  }
  return ilcc2_set_gather_targets(h, G->world, ct, st, slot_base, ILCC2_GATHER_PEER);
}

int ilcc2_gather_sync(ilcc2_handle* h, const float** corners_all, const int32_t** status_all, void* stream) {
  if (!h || !h->gather) return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "gather_sync: ilcc2_gather_setup_nccl first");
  ilcc2_gather* G = h->gather;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned char* half = G->local + (size_t)G->parity * G->half_bytes;
  int rc;
  if (G->fused) {
    // producers (K5 of every rank, stores into all ranks' buffers) | one all-reduce as the barrier | consumers
    rc = G->nccl.AllReduce(G->d_flag, G->d_flag + 1, 1, kNcclInt32, kNcclMin, G->comm, s);
  } else {
    if (!h->last_corners_dev || !h->last_status_dev)
      return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "gather_sync: the NCCL fallback needs a preceding ILCC2_MEM_DEVICE call");
    const int per = G->total_frames / G->world;  // the fallback exchanges equal shards
    if (per * G->world != G->total_frames || h->last_n_frames != per)
      return fail(h, ILCC2_ERR_INVALID_ARGUMENT, "gather_sync: the NCCL fallback needs total_frames / world frames per rank");
    rc = G->nccl.AllGather(h->last_corners_dev, half, sizeof(float) * 3 * (size_t)h->dp.ncorners * per, kNcclChar, G->comm, s);
    if (rc == 0)
      rc = G->nccl.AllGather(h->last_status_dev, half + G->corners_bytes, sizeof(int32_t) * (size_t)per, kNcclChar, G->comm, s);
  }
  if (rc != 0) return fail(h, ILCC2_ERR_CUDA, "gather_sync: NCCL error %d (%s)", rc, G->nccl.GetErrorString ? G->nccl.GetErrorString(rc) : "?");
  if (corners_all) *corners_all = reinterpret_cast<const float*>(half);
  if (status_all) *status_all = reinterpret_cast<const int32_t*>(half + G->corners_bytes);
  // The next extract writes the OTHER half: a rank that runs ahead cannot overwrite what a slower rank is still reading
  // (it can reach this half again only after the next sync, which every rank enters after it has consumed this one).
  G->parity ^= 1;
  if (G->fused) {
    float* ct[kMaxGatherTargets];
    int32_t* st[kMaxGatherTargets];
    for (int r = 0; r < G->world; ++r) {
      unsigned char* b = G->peer_base[r] + (size_t)G->parity * G->half_bytes;
      ct[r] = reinterpret_cast<float*>(b);
      st[r] = reinterpret_cast<int32_t*>(b + G->corners_bytes);
    }
    return ilcc2_set_gather_targets(h, G->world, ct, st, G->slot_base, ILCC2_GATHER_PEER);
  }
  return ILCC2_SUCCESS;
}

int ilcc2_gather_is_fused(const ilcc2_handle* h) { return h && h->gather && h->gather->fused ? 1 : 0; }

int ilcc2_gather_teardown(ilcc2_handle* h) {
  if (!h || !h->gather) return ILCC2_SUCCESS;
  ilcc2_gather* G = h->gather;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < G->world; ++r)
    if (G->peer_opened[r]) cudaIpcCloseMemHandle(G->peer_base[r]);
  if (G->local) cudaFree(G->local);
  if (G->d_flag) cudaFree(G->d_flag);
  delete G;
  h->gather = nullptr;
  return ilcc2_set_gather_targets(h, 0, nullptr, nullptr, 0, ILCC2_GATHER_PEER);
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
