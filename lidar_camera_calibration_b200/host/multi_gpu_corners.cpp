// multi_gpu_corners -- a C++ host that runs the corner path on G GPUs of one node WITHOUT Python: one process per GPU,
// contiguous frame shards (ilcc2_shard_range), the all-gather of the per-frame corners fused into K5 through
// ilcc2_gather_setup_nccl / ilcc2_gather_sync over an NCCL communicator.  It is what a multi-GPU build of the step that
// consumes every frame's corners (ilcc2/test/calib_lidar_cam.cpp:116-165) would do before its PnP, and the test
// vehicle of those two entry points (tests/test_gpu_multi.py).
//
//   multi_gpu_corners --frames frames.bin --clicks clicks.bin --n-frames N --points P --gpus G --steps K --out prefix
//
// frames.bin: N x P packed XYZI float32 records; clicks.bin: N x 3 float32.  Step s processes every shard rotated by s
// frames (so consecutive steps produce different corners in every slot: the double buffer of the gather is exercised).
// Every rank writes what it holds after ilcc2_gather_sync to <prefix>.rank<r>.step<s>.bin (N x ncorners x 3 float32,
// then N int32 status).
#include <cuda_runtime.h>
#include <nccl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "ilcc2_b200.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "[rank %d] %s: %s\n", g_rank, #x, cudaGetErrorString(e_));       \
      _exit(2);                                                                        \
    }                                                                                  \
  } while (0)
#define NK(x)                                                                          \
  do {                                                                                 \
    ncclResult_t e_ = (x);                                                             \
    if (e_ != ncclSuccess) {                                                           \
      fprintf(stderr, "[rank %d] %s: %s\n", g_rank, #x, ncclGetErrorString(e_));       \
      _exit(3);                                                                        \
    }                                                                                  \
  } while (0)
#define IK(h, x)                                                                       \
  do {                                                                                 \
    int e_ = (x);                                                                      \
    if (e_ != ILCC2_SUCCESS) {                                                         \
      fprintf(stderr, "[rank %d] %s: %d %s\n", g_rank, #x, e_, ilcc2_last_error(h));   \
      _exit(4);                                                                        \
    }                                                                                  \
  } while (0)

static int g_rank = -1;

static int run_rank(int rank, int world, const std::map<std::string, std::string>& a) {
  g_rank = rank;
  const int N = atoi(a.at("--n-frames").c_str()), P = atoi(a.at("--points").c_str()), K = atoi(a.at("--steps").c_str());
  const std::string prefix = a.at("--out"), idfile = prefix + ".ncclid";
  ncclUniqueId id;
  if (rank == 0) {
    NK(ncclGetUniqueId(&id));
    const std::string tmp = idfile + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f || fwrite(&id, sizeof(id), 1, f) != 1) return 5;
    fclose(f);
    rename(tmp.c_str(), idfile.c_str());
  } else {
    for (int tries = 0;; ++tries) {
      FILE* f = fopen(idfile.c_str(), "rb");
      if (f) {
        const size_t n = fread(&id, sizeof(id), 1, f);
        fclose(f);
        if (n == 1) break;
      }
      if (tries > 3000) return 6;
      usleep(10000);
    }
  }
  CK(cudaSetDevice(rank));
  ncclComm_t comm;
  NK(ncclCommInitRank(&comm, world, id, rank));
  int begin = 0, count = 0;
  ilcc2_shard_range(N, world, rank, &begin, &count);
  // this rank's shard from the files
  std::vector<float> pts((size_t)count * P * 4), clk((size_t)count * 3);
  {
    std::ifstream fp(a.at("--frames"), std::ios::binary), fc(a.at("--clicks"), std::ios::binary);
    fp.seekg((std::streamoff)begin * P * 16);
    fp.read(reinterpret_cast<char*>(pts.data()), (std::streamsize)pts.size() * 4);
    fc.seekg((std::streamoff)begin * 12);
    fc.read(reinterpret_cast<char*>(clk.data()), (std::streamsize)clk.size() * 4);
    if (!fp || !fc) return 7;
  }
  ilcc2_handle* h = nullptr;
  IK(nullptr, ilcc2_create_sized(nullptr, rank, count, P, 8192, 0, &h));
  IK(h, ilcc2_gather_setup_nccl(h, comm, N, begin));
  const int ncorn = 35;
  float *d_pts, *d_clk, *d_pts_rot, *d_clk_rot, *d_corners;
  int32_t* d_status;
  CK(cudaMalloc(&d_pts, pts.size() * 4));
  CK(cudaMalloc(&d_pts_rot, pts.size() * 4));
  CK(cudaMalloc(&d_clk, clk.size() * 4));
  CK(cudaMalloc(&d_clk_rot, clk.size() * 4));
  CK(cudaMalloc(&d_corners, (size_t)count * ncorn * 3 * 4));
  CK(cudaMalloc(&d_status, (size_t)count * 4));
  CK(cudaMemcpy(d_pts, pts.data(), pts.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_clk, clk.data(), clk.size() * 4, cudaMemcpyHostToDevice));
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  std::vector<float> hc((size_t)N * ncorn * 3);
  std::vector<int32_t> hs(N);
  if (rank == 0) {
    printf("world %d frames %d fused %d\n", world, N, ilcc2_gather_is_fused(h));
    fflush(stdout);
  }
  for (int step = 0; step < K; ++step) {
    const int r = count ? step % count : 0;  // rotate the shard by r frames: frame i of this step = frame (i + r) % count
    const size_t fb = (size_t)P * 16;
    CK(cudaMemcpyAsync(d_pts_rot, reinterpret_cast<char*>(d_pts) + r * fb, (count - r) * fb, cudaMemcpyDeviceToDevice, s));
    if (r) CK(cudaMemcpyAsync(reinterpret_cast<char*>(d_pts_rot) + (count - r) * fb, d_pts, r * fb, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(d_clk_rot, d_clk + 3 * r, (size_t)(count - r) * 12, cudaMemcpyDeviceToDevice, s));
    if (r) CK(cudaMemcpyAsync(d_clk_rot + 3 * (count - r), d_clk, (size_t)r * 12, cudaMemcpyDeviceToDevice, s));
    IK(h, ilcc2_extract_corners(h, d_pts_rot, count, P, nullptr, 16, 0, 4, 8, 12, ILCC2_MEM_DEVICE, d_clk_rot, nullptr,
                                ILCC2_FLAG_AUTO, nullptr, d_corners, d_status, nullptr, s));
    const float* all_c = nullptr;
    const int32_t* all_s = nullptr;
    IK(h, ilcc2_gather_sync(h, &all_c, &all_s, s));
    // no host synchronisation between steps except for the read-back of this step's gathered buffer
    CK(cudaMemcpyAsync(hc.data(), all_c, hc.size() * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hs.data(), all_s, hs.size() * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const std::string out = prefix + ".rank" + std::to_string(rank) + ".step" + std::to_string(step) + ".bin";
    FILE* f = fopen(out.c_str(), "wb");
    if (!f) return 8;
    fwrite(hc.data(), 4, hc.size(), f);
    fwrite(hs.data(), 4, hs.size(), f);
    fclose(f);
  }
  IK(h, ilcc2_gather_teardown(h));
  ilcc2_destroy(h);
  ncclCommDestroy(comm);
  return 0;
}

int main(int argc, char** argv) {
  std::map<std::string, std::string> a = {{"--gpus", "2"}, {"--steps", "3"}, {"--out", "gather"}};
  for (int i = 1; i + 1 < argc; i += 2) a[argv[i]] = argv[i + 1];
  for (const char* k : {"--frames", "--clicks", "--n-frames", "--points"})
    if (!a.count(k)) {
      fprintf(stderr, "usage: multi_gpu_corners --frames F --clicks C --n-frames N --points P [--gpus G --steps K --out prefix]\n");
      return 1;
    }
  const int world = atoi(a["--gpus"].c_str());
  unlink((a["--out"] + ".ncclid").c_str());
  std::vector<pid_t> kids;
  for (int r = 0; r < world; ++r) {  // fork BEFORE any CUDA / NCCL call: every rank is its own process
    const pid_t p = fork();
    if (p == 0) _exit(run_rank(r, world, a));
    kids.push_back(p);
  }
  int bad = 0;
  for (pid_t p : kids) {
    int st = 0;
    waitpid(p, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1;
  }
  unlink((a["--out"] + ".ncclid").c_str());
  return bad;
}
