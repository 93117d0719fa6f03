// ilcc2_kernels.cu -- sm_100a kernels of the LiDAR chessboard-corner path.
//
// One launch per stage over the whole batch, one CTA per frame in every stage:
//   K1 roi_filter   setROI                     (ilcc2/src/LidarCornersEst.cpp:48-70)      HBM-bound stream
//   K2 cluster      EuclideanCluster           (LidarCornersEst.cpp:124-153)              shared-memory hash + union-find
//   K3 plane        getPlane                   (LidarCornersEst.cpp:190-221)              RANSAC + covariance refit   } one launch:
//   K4 pca_gray     transformbyPCA + calHist   (LidarCornersEst.cpp:224-372)                                          } plane_pca_kernel
//   K5 board_fit    get_theta_t + getPCDcorners(Optimization.cpp:94-160, Optimization.h:31-107, LidarCornersEst.cpp:374-450, 501-556)
//
// Compiled with -fmad=false: every fp32 comparison that decides set membership (ROI box, 0.12 m linkage,
// 0.03 m plane distance) is an explicit sequence of IEEE operations, so results do not depend on
// contraction choices; the fp64 sums that feed decisions are reduced in a fixed order (strided
// per-thread partial sums + xor butterfly).  There is no dense contraction anywhere on this path
// (largest matrix: 3x3), hence no tensor-core / TMEM use; K1 is the only stage that touches all input bytes.
#include "ilcc2_internal.h"

#include <float.h>
#include <math.h>

#include <type_traits>

namespace ilcc2 {

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  // FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  r = __fadd_rn(r, __fmul_rn(dz, dz));
  return r;
}

__device__ __forceinline__ float plane_dist(const float m0, const float m1, const float m2, const float m3,
                                            float x, float y, float z) {
  // Eigen SSE Vector4f dot: (a*x + c*z) + (b*y + d)
  const float t0 = __fadd_rn(__fmul_rn(m0, x), __fmul_rn(m2, z));
  const float t1 = __fadd_rn(__fmul_rn(m1, y), m3);
  return fabsf(__fadd_rn(t0, t1));
}

__device__ __forceinline__ double warp_butterfly(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// Canonical 128-wide reduction for a 128-thread CTA: thread t holds the partial sum of elements k = t (mod 128).
// total = butterfly over strides 64, 32 (through shared memory), 16..1 (shuffles). Result valid in all threads
// of warp 0 and broadcast through shared memory to everyone.
template <int NQ>
__device__ __forceinline__ void block_reduce128(double (&v)[NQ], double* s_part /* [NQ][128] */, double* s_out /* [NQ] */) {
  const int t = threadIdx.x;
#pragma unroll
  for (int q = 0; q < NQ; ++q) s_part[q * 128 + t] = v[q];
  __syncthreads();
  if (t < 32) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double* p = s_part + q * 128;
      double s = (p[t] + p[t + 64]) + (p[t + 32] + p[t + 96]);
      s = warp_butterfly(s);
      if (t == 0) s_out[q] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NQ; ++q) v[q] = s_out[q];
  __syncthreads();
}

// Order-preserving compaction step for one tile of NT elements: returns this thread's output offset within
// the tile; tile_total = number of selected elements in the tile. All threads of the CTA must call.
template <int NT>
__device__ __forceinline__ int block_ordered_offset(bool pred, int* s_warp, int& tile_total) {
  const unsigned b = __ballot_sync(FULL, pred);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(b);
  __syncthreads();
  int off = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) {
    const int c = s_warp[i];
    if (i < warp) off += c;
    tot += c;
  }
  __syncthreads();
  tile_total = tot;
  return off + __popc(b & ((1u << lane) - 1u));
}

// L2 residency of the intermediates.  K1 streams 472 MB of input through the L2 (evict-first loads) and writes ~50 MB
// of ROI points that K2, K3 and K4 gather from; K4 writes the optimiser's points that K5 reads ~150 times.  Stored with
// an evict-last policy they stay in the 126 MB L2 across the streaming pass instead of being re-read from DRAM (ncu,
// 1024 frames: 581 -> see profiles/ MB of DRAM traffic per step for 472 MB of algorithmic bytes).
#ifndef ILCC2_NO_L2_HINTS
__device__ __forceinline__ unsigned long long l2_evict_last_policy() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_keep(float4* p, float4 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_keep(float2* p, float2 v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_keep(int* p, int v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
#else
__device__ __forceinline__ unsigned long long l2_evict_last_policy() { return 0ull; }
template <typename P, typename V>
__device__ __forceinline__ void st_keep(P* p, V v, unsigned long long) { *p = v; }
#endif

__device__ __forceinline__ void roi_limits(const float* click, const DevParams& p, float lim[6]) {
  // PassThrough::setFilterLimits takes floats; the reference computes click -/+ literal in double
  lim[0] = (float)((double)click[0] - (double)p.roi_hx);
  lim[1] = (float)((double)click[0] + (double)p.roi_hx);
  lim[2] = (float)((double)click[1] - (double)p.roi_hy);
  lim[3] = (float)((double)click[1] + (double)p.roi_hy);
  lim[4] = (float)((double)click[2] - (double)p.roi_hz);
  lim[5] = (float)((double)click[2] + (double)p.roi_hz);
}

// ------------------------------------------------------------------------------------------------
// K1  roi_filter: the only pass over all input bytes.  One CTA per frame streams the frame's records with
// 128-bit loads (two tiles in flight per thread), applies the three PassThrough box tests and writes the
// survivors in input order (ballot + block scan), tracking the click's nearest neighbour on the way.
// MODE 0: packed 16-byte XYZI; MODE 1: 32-byte records x,y,z@0 intensity@16; MODE 2: generic offsets.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ float4 load_point(const unsigned char* rec, int ox, int oy, int oz, int oi) {
  if (MODE == 0) {
    return __ldcs(reinterpret_cast<const float4*>(rec));
  } else if (MODE == 1) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(rec));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(rec + 16));
    return make_float4(a.x, a.y, a.z, b.x);
  } else {
    return make_float4(__ldcs(reinterpret_cast<const float*>(rec + ox)), __ldcs(reinterpret_cast<const float*>(rec + oy)),
                       __ldcs(reinterpret_cast<const float*>(rec + oz)), __ldcs(reinterpret_cast<const float*>(rec + oi)));
  }
}

// NT threads per CTA, 1024 threads (64 registers each) resident per SM.  Large batches use small CTAs (NT = 128: all
// 1024 frames of the bench batch are resident in ONE wave, 0.091 ms = 5.2 TB/s; NT = 512 has 3.5 waves and a 13 %
// tail, 0.098 ms), small batches use large ones (a single frame is streamed by whatever one CTA has in flight).
template <int MODE, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) roi_filter_kernel(const Workspace w, const DevParams p) {
  constexpr int U = kRoiUnroll, TILE = NT * U, NW = NT / 32;
  __shared__ int s_cnt[2][U * NW];
  __shared__ unsigned long long s_key[NW];

  const int f = blockIdx.x + w.frame_base, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long start;
  int n;
  if (w.frame_off) {
    start = w.frame_off[f];
    n = (int)(w.frame_off[f + 1] - start);
  } else {
    start = (long long)f * w.pts_per_frame;
    n = w.pts_per_frame;
  }
  const unsigned char* base = w.points + (size_t)start * w.step;
  const float cx = w.clicks[3 * f + 0], cy = w.clicks[3 * f + 1], cz = w.clicks[3 * f + 2];
  float lim[6];
  {
    const float ck[3] = {cx, cy, cz};
    roi_limits(ck, p, lim);
  }
  float4* out_pts = w.roi_pts + (size_t)f * w.cap;
  int* out_src = w.roi_src + (size_t)f * w.cap;
  const unsigned long long keep = l2_evict_last_policy();

  float4 cur[U], nxt[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int idx = u * NT + tid;
    cur[u] = (idx < n) ? load_point<MODE>(base + (size_t)idx * w.step, w.ox, w.oy, w.oz, w.oi)
                       : make_float4(NAN, NAN, NAN, 0.f);
  }
  int total = 0;
  unsigned long long best = ~0ull;
  for (int tile0 = 0; tile0 < n; tile0 += TILE) {
    // prefetch the next tile before the scan of this one
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = tile0 + TILE + u * NT + tid;
      nxt[u] = (idx < n) ? load_point<MODE>(base + (size_t)idx * w.step, w.ox, w.oy, w.oz, w.oi)
                         : make_float4(NAN, NAN, NAN, 0.f);
    }
    bool pr[U];
    int rank[U];
    int* cnt = s_cnt[(tile0 / TILE) & 1];  // double-buffered: one barrier per tile is enough
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 v = cur[u];
      bool ok = isfinite(v.x) && isfinite(v.y) && isfinite(v.z);
      ok = ok && !(v.z < lim[4] || v.z > lim[5]);
      ok = ok && !(v.x < lim[0] || v.x > lim[1]);
      ok = ok && !(v.y < lim[2] || v.y > lim[3]);
      pr[u] = ok;
      const unsigned b = __ballot_sync(FULL, ok);
      rank[u] = __popc(b & ((1u << lane) - 1u));
      if (lane == 0) cnt[u * NW + warp] = __popc(b);
    }
    __syncthreads();
    // every warp scans the U*NW (= 64) counts itself (u-major = input order): no serial section, no second barrier
    static_assert(U * NW <= 64, "scan below handles at most 64 counts");
    int my_off[U], tile_total;
    {
      const int a = lane < U * NW ? cnt[lane] : 0, b = 32 + lane < U * NW ? cnt[32 + lane] : 0;
      int sa = a, sb = b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(FULL, sa, o), tb = __shfl_up_sync(FULL, sb, o);
        if (lane >= o) {
          sa += ta;
          sb += tb;
        }
      }
      const int tot_a = __shfl_sync(FULL, sa, 31);
      tile_total = tot_a + __shfl_sync(FULL, sb, 31);
      const int ex_a = sa - a, ex_b = tot_a + sb - b;  // exclusive offsets of counts lane and 32 + lane
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = u * NW + warp;  // this warp's count for sub-tile u
        const int va = __shfl_sync(FULL, ex_a, idx & 31), vb = __shfl_sync(FULL, ex_b, idx & 31);
        my_off[u] = idx < 32 ? va : vb;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (pr[u]) {
        const int k = total + my_off[u] + rank[u];
        const float4 v = cur[u];
        if (k < w.cap) {
          st_keep(out_pts + k, v, keep);
          st_keep(out_src + k, tile0 + u * NT + tid, keep);
        }
        const float d2 = sqdist3(v.x, v.y, v.z, cx, cy, cz);
        const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)k;
        best = key < best ? key : best;
      }
    }
    total += tile_total;
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = nxt[u];
  }
  // nearest neighbour of the click: min over (dist2, index)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(FULL, best, o);
    best = other < best ? other : best;
  }
  if (lane == 0) s_key[warp] = best;
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < NW; ++i) best = s_key[i] < best ? s_key[i] : best;
    const bool over = total > w.cap;  // more ROI points than the handle's roi_capacity: reported, nothing is truncated
    const int m = over ? 0 : total;
    w.roi_count[f] = m;
    w.nn_key[f] = best;
    ilcc2_diag* d = w.diag + f;
    d->n_roi = total;
    d->nn_index = m == 0 ? 0 : (int)(best & 0xffffffffull);
    d->status = over ? ILCC2_ROI_CAPACITY : (m == 0) ? ILCC2_NO_POINTS_IN_ROI : ILCC2_OK;
  }
}

template <int NT>
static void launch_roi_filter_nt(const Workspace& w, const DevParams& p, cudaStream_t s) {
  if (w.step == 16 && w.ox == 0 && w.oy == 4 && w.oz == 8 && w.oi == 12)
    roi_filter_kernel<0, NT><<<w.n_frames, NT, 0, s>>>(w, p);
  else if (w.step == 32 && w.ox == 0 && w.oy == 4 && w.oz == 8 && w.oi == 16)
    roi_filter_kernel<1, NT><<<w.n_frames, NT, 0, s>>>(w, p);
  else
    roi_filter_kernel<2, NT><<<w.n_frames, NT, 0, s>>>(w, p);
}

void launch_roi_filter(const Workspace& w, const DevParams& p, cudaStream_t s) {
  if (w.load_frames >= 4 * w.num_sms)
    launch_roi_filter_nt<128>(w, p, s);
  else
    launch_roi_filter_nt<kRoiThreads>(w, p, s);
}

// ------------------------------------------------------------------------------------------------
// K2  cluster: exact single-linkage components at distance < tol (pcl::EuclideanClusterExtraction) of the
// ROI points of one frame.  Points are hashed into cells of edge 0.95*tol/sqrt(3): two points of one cell
// are always linked, and linked points are at most 2 cells apart per axis.  Union-find runs over points
// (parent of a point = head of its cell to begin with); every occupied cell is joined with each occupied
// cell of the "forward" half of its 5x5x5 neighbourhood as soon as one point pair is closer than tol
// (skipped when the two already share a root).  The 3x3x3 core goes first so that most outer pairs are
// already connected when they are visited.
// Shared-memory layout (frames up to kClusterSmemPts ROI points): open-addressing hash cell -> slot
// (key, head of the cell's point list), 16-bit next pointers, union-find parents, list of occupied slots,
// and an occupancy BITMAP with one 32-bit word per (y,z) row of cells, so the five x-neighbours of a row are
// tested with one load and the hash is only probed for cells that exist.  Larger frames use the same code on
// a per-frame global scratch without the bitmap.
// ------------------------------------------------------------------------------------------------
// 16-bit tables in the shared-memory instantiation (point indices and slots are < 4096), 32-bit in the global one
__device__ __forceinline__ int cas_idx(int* a, int cmp, int val) { return atomicCAS(a, cmp, val); }
__device__ __forceinline__ int cas_idx(uint16_t* a, int cmp, int val) {
  return (int)atomicCAS(reinterpret_cast<unsigned short*>(a), (unsigned short)cmp, (unsigned short)val);
}
__device__ __forceinline__ int exch_idx(int* a, int val) { return atomicExch(a, val); }
__device__ __forceinline__ int exch_idx(uint16_t* a, int val) {  // no 16-bit atomicExch: CAS loop
  int old = *reinterpret_cast<volatile uint16_t*>(a), assumed;
  do {
    assumed = old;
    old = cas_idx(a, assumed, val);
  } while (old != assumed);
  return old;
}

template <typename ParT>
__device__ __forceinline__ int uf_find(ParT* par, int x) {
  volatile ParT* vp = par;
  int cur = (int)vp[x];
  if (cur != x) {
    int prev = x, nx;
    while (cur != (nx = (int)vp[cur])) {
      vp[prev] = (ParT)nx;
      prev = cur;
      cur = nx;
    }
  }
  return cur;
}

template <typename ParT>
__device__ __forceinline__ void uf_union(ParT* par, int a, int b) {
  int ra = uf_find(par, a), rb = uf_find(par, b);
  while (ra != rb) {
    if (ra < rb) {
      const int t = ra;
      ra = rb;
      rb = t;
    }
    const int old = cas_idx(&par[ra], ra, rb);  // link the larger root under the smaller
    if (old == ra) break;
    ra = uf_find(par, old);
    rb = uf_find(par, rb);
  }
}

__device__ __forceinline__ unsigned hash_cell(int k) {
  unsigned h = (unsigned)k * 2654435761u;
  return h ^ (h >> 15);
}

// cell key: x | y << 10 | z << 20 (each axis < 1024 cells; checked at handle creation)
__device__ __forceinline__ int pack_cell(int x, int y, int z) { return x | (y << 10) | (z << 20); }

// Slot of cell `kb` in the open-addressing table, or -1 when the cell is not occupied.
__device__ __forceinline__ int find_cell(const int* key, int mask, int kb) {
  unsigned h = hash_cell(kb) & mask;
  for (;;) {
    const int kk = key[h];
    if (kk == kb) return (int)h;
    if (kk == -1) return -1;
    h = (h + 1) & mask;
  }
}

// Link the cells whose point lists start at ha / hb as soon as one point pair is closer than tol (nothing to do when
// they already share a root).
template <typename NxtT>
__device__ __forceinline__ void pair_join(const float4* __restrict__ P, const NxtT* nxt, NxtT* parent, int ha, int hb,
                                          float tol2) {
  if (uf_find(parent, ha) == uf_find(parent, hb)) return;
  constexpr int kEnd = (NxtT)(-1);
  for (int pa = ha; pa != kEnd; pa = (int)nxt[pa]) {
    const float4 qa = __ldg(P + pa);
    for (int pb = hb; pb != kEnd; pb = (int)nxt[pb]) {
      const float4 qb = __ldg(P + pb);
      if (sqdist3(qa.x, qa.y, qa.z, qb.x, qb.y, qb.z) < tol2) {
        uf_union(parent, pa, pb);
        return;
      }
    }
  }
}

// forward half of the 5x5 (dy, dz) rows: row 0 is the cell's own row (dx > 0 only)
__constant__ signed char c_row_dy[13] = {0, 1, 2, -2, -1, 0, 1, 2, -2, -1, 0, 1, 2};
__constant__ signed char c_row_dz[13] = {0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2};

template <bool SMEM>
__global__ void __launch_bounds__(kClusterThreads, SMEM ? 3 : 1) cluster_kernel(const Workspace w, const DevParams p) {
  constexpr int NT = kClusterThreads;
  using NxtT = typename std::conditional<SMEM, uint16_t, int>::type;
  constexpr int kEnd = (NxtT)(-1);
  extern __shared__ int sm[];
  __shared__ int s_warp[NT / 32];
  __shared__ int s_nocc, s_nkept, s_chosen_root, s_slot;
  __shared__ unsigned s_queue_fallback[SMEM ? 1 : (NT / 32) * kClusterQueue];  // code words
  __shared__ int s_queue_slot[SMEM ? 1 : (NT / 32) * kClusterQueue];         // slot ids (do not fit 13 bits)
  __shared__ int s_q2[2][(NT / 32) * 64];  // per-warp second queue: list heads of the cell pairs whose roots differ

  const int f = blockIdx.x + w.frame_base, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ilcc2_diag* d = w.diag + f;
  int* sel = w.sel_idx + (size_t)f * w.cap;
  const int M = w.roi_count[f];
  // the shared-memory instantiation runs first and owns the frames that fit; it hands a frame over (sel_count = -1)
  // when its hash table gets too full
#ifdef CLUSTER_FORCE_GLOBAL  // A/B: every frame on the global-memory tables
  if (SMEM) return;
#else
  if (SMEM ? (M > kClusterSmemPts) : (M <= kClusterSmemPts && w.sel_count[f] != -1)) return;
#endif
  if (M == 0) {
    if (tid == 0) w.sel_count[f] = 0;
    return;
  }
  int TS;
  int* key;
  NxtT *head, *parent, *nxt, *occ;
  int* min_idx;  // step 3: lowest point index per component
  unsigned* bitmap = nullptr;
  unsigned* queue = nullptr;  // per-warp queue of (cell, neighbour) pairs
  int* qslot = nullptr;
  const int nrows = p.ncy * p.ncz;
  const bool use_bitmap = SMEM && p.ncx <= 32 && nrows <= kClusterBitmapRows;
  if (SMEM) {
    TS = kClusterSmemSlots;
    key = sm;
    head = reinterpret_cast<NxtT*>(key + TS);
    parent = head + TS;
    nxt = parent + kClusterSmemPts;
    occ = nxt + kClusterSmemPts;
    bitmap = reinterpret_cast<unsigned*>(occ + kClusterSmemPts);
    queue = bitmap + kClusterBitmapRows + warp * kClusterQueue;
    min_idx = reinterpret_cast<int*>(bitmap);  // bitmap + queues (>= kClusterSmemPts ints) are free by then
  } else {
    TS = w.cl_ts;
    if (tid == 0) s_slot = atomicAdd(w.cl_pool_next, 1);  // next free table of the pool
    __syncthreads();
    if (s_slot >= w.cl_pool_slots) {  // more overflow frames in this call than the handle holds tables for
      if (tid == 0) {
        d->status = ILCC2_ROI_CAPACITY;
        w.sel_count[f] = 0;
      }
      return;
    }
    int* g = w.cl_scratch + (size_t)s_slot * w.cl_stride;
    key = g;
    head = reinterpret_cast<NxtT*>(key + TS);
    parent = head + TS;
    nxt = parent + w.cap;
    occ = nxt + w.cap;
    min_idx = reinterpret_cast<int*>(head);
    queue = s_queue_fallback + warp * kClusterQueue;
    qslot = s_queue_slot + warp * kClusterQueue;
  }
  for (int i = tid; i < TS; i += NT) {
    key[i] = -1;
    head[i] = (NxtT)kEnd;
  }
  if (use_bitmap)
    for (int i = tid; i < nrows; i += NT) bitmap[i] = 0u;
  if (tid == 0) s_nocc = 0;
  __syncthreads();

  const float4* P = w.roi_pts + (size_t)f * w.cap;
  float lim[6];
  roi_limits(w.clicks + 3 * f, p, lim);
  const int mask = TS - 1;
  // 1. insert points into the cell hash; per-cell linked lists; occupancy bitmap
  for (int i = tid; i < M; i += NT) {
    const float4 q = P[i];
    int cx = (int)floorf((q.x - lim[0]) * p.inv_cell);
    int cy = (int)floorf((q.y - lim[2]) * p.inv_cell);
    int cz = (int)floorf((q.z - lim[4]) * p.inv_cell);
    cx = min(max(cx, 0), p.ncx - 1);
    cy = min(max(cy, 0), p.ncy - 1);
    cz = min(max(cz, 0), p.ncz - 1);
    const int k = pack_cell(cx, cy, cz);
    unsigned h = hash_cell(k) & mask;
    for (;;) {
      const int prev = atomicCAS(&key[h], -1, k);
      if (prev == -1) {
        occ[atomicAdd(&s_nocc, 1)] = (NxtT)h;
        if (use_bitmap) atomicOr(&bitmap[cz * p.ncy + cy], 1u << cx);
        break;
      }
      if (prev == k) break;
      h = (h + 1) & mask;
    }
    nxt[i] = (NxtT)exch_idx(&head[h], i);
    parent[i] = (NxtT)h;  // slot for now
  }
  __syncthreads();
  if (SMEM && s_nocc > kClusterMaxOcc) {  // lookups of absent cells need empty slots: leave the frame to the global tables
    if (tid == 0) w.sel_count[f] = -1;
    return;
  }
  for (int i = tid; i < M; i += NT) parent[i] = head[(int)parent[i]];  // same cell => same set
  __syncthreads();
  // 2. join neighbouring cells: G lanes per cell, lane r of a group owns row r of the forward (dy,dz) rows
  const int nocc = s_nocc;
  const float tol2 = p.tol2;
  int* q2a = s_q2[0] + warp * 64;
  int* q2b = s_q2[1] + warp * 64;
  int n2 = 0, qn = 0;  // pairs waiting in the second / first queue of this warp
  // stage 1 for queue entry e of every lane (`act`: the lane has one), then stage 2 whenever 32 pairs have come together
  auto drain = [&](int e, bool act) {
    int ha = 0, hb = 0;
    bool need = false;
    if (act) {
      const unsigned q = queue[e];
      const int qa = SMEM ? (int)(q & 8191u) : qslot[e];
      const int bit = (int)((q >> 13) & 7u), r = (int)(q >> 16);
      const int ka = key[qa];
      const int bx = (ka & 1023) + bit - 2, by = ((ka >> 10) & 1023) + c_row_dy[r], bz = (ka >> 20) + c_row_dz[r];
      const int hs = find_cell(key, mask, pack_cell(bx, by, bz));
      if (hs >= 0) {
        ha = (int)head[qa];
        hb = (int)head[hs];
        need = uf_find(parent, ha) != uf_find(parent, hb);
      }
    }
    const unsigned m = __ballot_sync(FULL, need);
    if (need) {
      const int at = n2 + __popc(m & ((1u << lane) - 1u));
      q2a[at] = ha;
      q2b[at] = hb;
    }
    n2 += __popc(m);
    __syncwarp();
    if (n2 >= 32) {
      n2 -= 32;
      pair_join<NxtT>(P, nxt, parent, q2a[n2 + lane], q2b[n2 + lane], tol2);
      __syncwarp();
    }
  };
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    // pass 0: the 13 offsets of the 3x3x3 core (5 rows x dx in [-1,1]); pass 1: the 49 outer ones (13 rows)
    const int G = pass == 0 ? 8 : 16;
    const int nrow = pass == 0 ? 5 : 13;
    const int cells_per_iter = NT / G;
    const int gl = tid % G;
    int dy = 0, dz = 0;
    if (gl < nrow) {
      // pass 0 rows: (0,0) (1,0) (-1,1) (0,1) (1,1);  pass 1: all 13 rows
      const int r = pass == 0 ? (gl == 0 ? 0 : gl == 1 ? 1 : gl + 2) : gl;
      dy = c_row_dy[r];
      dz = c_row_dz[r];
    }
    for (int c0 = 0; c0 < nocc; c0 += cells_per_iter) {  // warp-uniform trip count
      const int ci = c0 + tid / G;
      // (a) each lane finds the occupied neighbour cells of its (cell, row): a 5-bit mask over dx
      unsigned want = 0u;
      int a = 0;
      if (ci < nocc && gl < nrow) {
        a = (int)occ[ci];
        const int ka = key[a];
        const int ax = ka & 1023, ay = (ka >> 10) & 1023, az = ka >> 20;
        const int by = ay + dy, bz = az + dz;
        if ((unsigned)by < (unsigned)p.ncy && (unsigned)bz < (unsigned)p.ncz) {
          const bool own = dy == 0 && dz == 0;
          const bool core_row = dy >= -1 && dy <= 1 && dz <= 1;
          if (pass == 0)
            want = own ? 0x08u : 0x0eu;                       // dx = 1 | dx in {-1,0,1}
          else
            want = own ? 0x10u : (core_row ? 0x11u : 0x1fu);  // dx = 2 | dx = +-2 | all five
          // drop dx that fall outside the grid (bit j <-> bx = ax + j - 2)
          if (ax < 2) want &= (0x1fu << (2 - ax)) & 0x1fu;
          if (ax + 2 >= p.ncx) want &= 0x1fu >> (ax + 3 - p.ncx);
          if (use_bitmap) {
            const unsigned row = bitmap[bz * p.ncy + by];
            want &= (unsigned)((((unsigned long long)row) << 2) >> ax) & 0x1fu;  // bits ax-2 .. ax+2
          }
        }
      }
      // (b) compact the (cell, neighbour) pairs of the warp into its queue: all lanes then work on real pairs
      // one ballot per dx: the lanes that want this neighbour append behind each other (uniform, no scan, no per-lane loop)
      const int rcode = pass == 0 ? (gl == 0 ? 0 : gl == 1 ? 1 : gl + 2) : gl;
      const unsigned lt = (1u << lane) - 1u;
#pragma unroll
      for (int bit = 0; bit < 5; ++bit) {
        const bool on = (want >> bit) & 1u;
        const unsigned mb = __ballot_sync(FULL, on);
        if (on) {
          const int at = qn + __popc(mb & lt);
          if (SMEM) {
            queue[at] = (unsigned)a | ((unsigned)bit << 13) | ((unsigned)rcode << 16);
          } else {
            qslot[at] = a;
            queue[at] = ((unsigned)bit << 13) | ((unsigned)rcode << 16);
          }
        }
        qn += __popc(mb);
      }
      __syncwarp();
      // (c) two stages, so that the lanes of a warp do the same kind of work at the same time (a single stage ran
      // with 8 of 32 lanes active: most pairs end at the root comparison, a few walk two point lists).  Stage 1, one
      // pair per lane: find the neighbour's slot, compare the roots of the two cells.  Pairs whose roots differ are
      // compacted into a second per-warp queue; stage 2 takes 32 of them at a time (distance tests + union).
      // The queue is drained 32 pairs at a time (a round adds ~12 pairs per warp); what is left stays for the next round.
      while (qn >= 32) {
        qn -= 32;
        drain(qn + lane, true);
      }
    }
    if (qn > 0) drain(lane, lane < qn);  // the pairs left over from this pass
    qn = 0;
    if (lane < n2) pair_join<NxtT>(P, nxt, parent, q2a[lane], q2b[lane], tol2);  // the pairs left over from this pass
    n2 = 0;
    __syncthreads();
  }
  // 3. roots, sizes, lowest index per component (key/head tables are free now)
  NxtT* root = nxt;
  int* size_of = key;  // TS >= M entries
  NxtT* kept = occ;
  for (int i = tid; i < M; i += NT) {
    size_of[i] = 0;
    min_idx[i] = 0x7fffffff;
  }
  if (tid == 0) s_nkept = 0;
  __syncthreads();
  for (int i = tid; i < M; i += NT) {
    const int r = uf_find(parent, i);
    root[i] = (NxtT)r;
    atomicAdd(&size_of[r], 1);
    atomicMin(&min_idx[r], i);
  }
  __syncthreads();
  for (int i = tid; i < M; i += NT)
    if ((int)root[i] == i && size_of[i] >= p.min_cl && size_of[i] <= p.max_cl) kept[atomicAdd(&s_nkept, 1)] = (NxtT)i;
  __syncthreads();
  const int nk = s_nkept;
  if (nk == 0) {
    if (tid == 0) {
      d->status = ILCC2_NO_CLUSTER;
      d->n_clusters = 0;
      w.sel_count[f] = 0;
    }
    return;
  }
  // 4. rank kept clusters: size descending, lowest first index first (PCL leaves ties unspecified)
  NxtT* rank_of = parent;  // indexed by root; parent[] is not needed once root[] is final
  for (int e = tid; e < nk; e += NT) {
    const int r = (int)kept[e];
    const int sz = size_of[r], mi = min_idx[r];
    int rank = 0;
    for (int j = 0; j < nk; ++j) {
      const int r2 = (int)kept[j];
      const int s2 = size_of[r2];
      if (s2 > sz || (s2 == sz && min_idx[r2] < mi)) ++rank;
    }
    rank_of[r] = (NxtT)rank;
  }
  __syncthreads();
  if (tid == 0) {
    const int nn = (int)(w.nn_key[f] & 0xffffffffull);
    const int rn = (int)root[nn];
    const bool in = size_of[rn] >= p.min_cl && size_of[rn] <= p.max_cl;
    int rank = in ? (int)rank_of[rn] : 0;
    const int delta = w.pdelta ? (int)w.pdelta[f] : 0;
    if (delta > 0) rank = min(rank + delta, nk - 1);
    if (delta < 0) rank = max(rank + delta, 0);
    int chosen = (int)kept[0];
    for (int j = 0; j < nk; ++j)
      if ((int)rank_of[(int)kept[j]] == rank) chosen = (int)kept[j];
    s_chosen_root = chosen;
    d->n_clusters = nk;
    d->click_in_cluster = in ? 1 : 0;
    d->cluster_rank = rank;
    d->n_cluster_pts = size_of[chosen];
  }
  __syncthreads();
  // 5. chosen cluster in input order
  const int chosen = s_chosen_root;
  int total = 0;
  for (int base = 0; base < M; base += NT) {
    const int i = base + tid;
    const bool pr = i < M && (int)root[i] == chosen;
    int tile_total;
    const int off = block_ordered_offset<NT>(pr, s_warp, tile_total);
    if (pr) sel[total + off] = i;
    total += tile_total;
  }
  if (tid == 0) w.sel_count[f] = total;
  (void)kEnd;
  (void)lane;
  (void)warp;
}

size_t cluster_smem_bytes() {  // 66 KB: three CTAs per SM
  static_assert(sizeof(unsigned) * (kClusterBitmapRows + (kClusterThreads / 32) * kClusterQueue) >= sizeof(int) * kClusterSmemPts,
                "step 3 reuses the bitmap + queue region for one int per point");
  return sizeof(int) * kClusterSmemSlots + sizeof(uint16_t) * (kClusterSmemSlots + 3 * kClusterSmemPts) +
         sizeof(unsigned) * (kClusterBitmapRows + (kClusterThreads / 32) * kClusterQueue);
}

void launch_cluster(const Workspace& w, const DevParams& p, cudaStream_t s) {
  cluster_kernel<true><<<w.n_frames, kClusterThreads, cluster_smem_bytes(), s>>>(w, p);
  cluster_kernel<false><<<w.n_frames, kClusterThreads, 0, s>>>(w, p);  // frames the first one could not take
}

// ------------------------------------------------------------------------------------------------
// symmetric 3x3 eigen solvers (device restatements; see DESIGN.md for what each one mirrors)
// ------------------------------------------------------------------------------------------------
// double cyclic Jacobi, fixed 12 sweeps: used for the least-squares plane refit.
__device__ void jacobi3d(const double A[6], double ev[3], double V[3][3]) {
  double a[3][3] = {{A[0], A[1], A[2]}, {A[1], A[3], A[4]}, {A[2], A[4], A[5]}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  // One thread runs this while its CTA waits: the sweep loop stays a loop (three rotations = ~550 instructions that
  // stay in the instruction cache) instead of 12 x 3 unrolled rotations (6561 instructions fetched once each, cold).
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
#pragma unroll
    for (int pi = 0; pi < 2; ++pi)
#pragma unroll
      for (int q = pi + 1; q < 3; ++q) {
        const double apq = a[pi][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[pi][pi]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0);
        const double s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][pi], akq = a[k][q];
          a[k][pi] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[pi][k], aqk = a[q][k];
          a[pi][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][pi], vkq = V[k][q];
          V[k][pi] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  ev[0] = a[0][0];
  ev[1] = a[1][1];
  ev[2] = a[2][2];
}

__device__ __forceinline__ float eig_hypotf(float x, float y) {
  const float ax = fabsf(x), ay = fabsf(y);
  float pp, qp;
  if (ax > ay) {
    pp = ax;
    qp = ay / pp;
  } else {
    pp = ay;
    qp = ax / pp;
  }
  if (pp == 0.0f) return 0.0f;
  return pp * sqrtf(1.0f + qp * qp);
}

// float, Eigen-3.2-style SelfAdjointEigenSolver<Matrix3f>::compute (tridiagonalise, implicit QR with
// Wilkinson shift, deflation at 1e-5, ascending sort): defines the SIGN of the PCA axes
// (LidarCornersEst.cpp:337-343).  cov row-major; Q[r][c] columns = eigenvectors.
__device__ void eigen3f(const float cov[9], float evals[3], float Q[3][3]) {
  float m00 = cov[0], m10 = cov[3], m11 = cov[4], m20 = cov[6], m21 = cov[7], m22 = cov[8];
  float scale = fabsf(m00);
  scale = fmaxf(scale, fabsf(m10));
  scale = fmaxf(scale, fabsf(m11));
  scale = fmaxf(scale, fabsf(m20));
  scale = fmaxf(scale, fabsf(m21));
  scale = fmaxf(scale, fabsf(m22));
  if (scale == 0.0f) scale = 1.0f;
  m00 /= scale; m10 /= scale; m11 /= scale; m20 /= scale; m21 /= scale; m22 /= scale;
  float diag[3], sub[2];
  diag[0] = m00;
  const float v1norm2 = m20 * m20;
  if (v1norm2 <= FLT_MIN) {
    diag[1] = m11;
    diag[2] = m22;
    sub[0] = m10;
    sub[1] = m21;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Q[r][c] = (r == c) ? 1.0f : 0.0f;
  } else {
    const float beta = sqrtf(m10 * m10 + v1norm2);
    const float inv_beta = 1.0f / beta;
    const float m01 = m10 * inv_beta;
    const float m02 = m20 * inv_beta;
    const float q = 2.0f * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q;
    diag[2] = m22 - m02 * q;
    sub[0] = beta;
    sub[1] = m21 - m01 * q;
    Q[0][0] = 1; Q[0][1] = 0;   Q[0][2] = 0;
    Q[1][0] = 0; Q[1][1] = m01; Q[1][2] = m02;
    Q[2][0] = 0; Q[2][1] = m02; Q[2][2] = -m01;
  }
  int end = 2, start = 0, iter = 0;
  const float precision = 1e-5f;
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (fabsf(sub[i]) <= (fabsf(diag[i]) + fabsf(diag[i + 1])) * precision) sub[i] = 0.0f;
    while (end > 0 && sub[end - 1] == 0.0f) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 90) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0f) start--;
    const float td = (diag[end - 1] - diag[end]) * 0.5f;
    const float e = sub[end - 1];
    float mu = diag[end];
    if (td == 0.0f) {
      mu -= fabsf(e);
    } else if (e != 0.0f) {
      const float e2 = e * e;
      const float h = eig_hypotf(td, e);
      if (e2 == 0.0f)
        mu -= e / ((td + (td > 0.0f ? h : -h)) / e);
      else
        mu -= e2 / (td + (td > 0.0f ? h : -h));
    }
    float x = diag[start] - mu;
    float z = sub[start];
    for (int k = start; k < end; ++k) {
      float c, s;
      if (z == 0.0f) {
        c = x < 0.0f ? -1.0f : 1.0f;
        s = 0.0f;
      } else if (x == 0.0f) {
        c = 0.0f;
        s = z < 0.0f ? 1.0f : -1.0f;
      } else if (fabsf(x) > fabsf(z)) {
        const float t = z / x;
        float u = sqrtf(1.0f + t * t);
        if (x < 0.0f) u = -u;
        c = 1.0f / u;
        s = -t * c;
      } else {
        const float t = x / z;
        float u = sqrtf(1.0f + t * t);
        if (z < 0.0f) u = -u;
        s = -1.0f / u;
        c = -t * s;
      }
      const float sdk = s * diag[k] + c * sub[k];
      const float dkp1 = s * sub[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      sub[k] = c * sdk - s * dkp1;
      if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
      x = sub[k];
      if (k < end - 1) {
        z = -s * sub[k + 1];
        sub[k + 1] = c * sub[k + 1];
      }
      for (int r = 0; r < 3; ++r) {
        const float xi = Q[r][k], yi = Q[r][k + 1];
        Q[r][k] = c * xi - s * yi;
        Q[r][k + 1] = s * xi + c * yi;
      }
    }
  }
  for (int i = 0; i < 2; ++i) {
    int k = 0;
    for (int j = 1; j < 3 - i; ++j)
      if (diag[i + j] < diag[i + k]) k = j;
    if (k > 0) {
      const float t = diag[i];
      diag[i] = diag[i + k];
      diag[i + k] = t;
      for (int r = 0; r < 3; ++r) {
        const float q = Q[r][i];
        Q[r][i] = Q[r][i + k];
        Q[r][i + k] = q;
      }
    }
  }
  for (int i = 0; i < 3; ++i) evals[i] = diag[i] * scale;
}

// ------------------------------------------------------------------------------------------------
// K3  plane: pcl::SACSegmentation(PLANE, RANSAC, 0.03, optimizeCoefficients) on the chosen cluster.
// Thread 0 walks PCL's sample sequence (shuffled-index swaps driven by the fixed-seed draw table), every
// hypothesis is scored by the whole CTA; then the LS refit (fp64 covariance, 128-wide canonical sums,
// smallest eigenvector) and the final inlier re-selection, written in input order.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sample_good(const float4 p0, const float4 p1, const float4 p2) {
  const float r0 = (p1.x - p0.x) / (p2.x - p0.x);
  const float r1 = (p1.y - p0.y) / (p2.y - p0.y);
  const float r2 = (p1.z - p0.z) / (p2.z - p0.z);
  return (r0 != r1) || (r2 != r1);
}

__device__ __forceinline__ bool plane_from_3(const float4 p0, const float4 p1, const float4 p2, float m[4]) {
  const float a0 = p1.x - p0.x, a1 = p1.y - p0.y, a2 = p1.z - p0.z;
  const float b0 = p2.x - p0.x, b1 = p2.y - p0.y, b2 = p2.z - p0.z;
  const float r0 = a0 / b0, r1 = a1 / b1, r2 = a2 / b2;
  if (r0 == r1 && r2 == r1) return false;
  float n0 = a1 * b2 - a2 * b1;
  float n1 = a2 * b0 - a0 * b2;
  float n2 = a0 * b1 - a1 * b0;
  float nn = n0 * n0;
  nn = nn + n1 * n1;
  nn = nn + n2 * n2;
  const float norm = sqrtf(nn);
  n0 = n0 / norm;
  n1 = n1 / norm;
  n2 = n2 / norm;
  float dd = n0 * p0.x;
  dd = dd + n1 * p0.y;
  dd = dd + n2 * p0.z;
  m[0] = n0;
  m[1] = n1;
  m[2] = n2;
  m[3] = -1.0f * dd;
  return true;
}

__device__ __forceinline__ void plane_stage(const Workspace& w, const DevParams& p, double* s_part /* [9 * 128] */) {
  constexpr int NT = kPlaneThreads;
  __shared__ float s_model[4];
  __shared__ int s_ctrl;  // 1: score s_model, 0: stop
  __shared__ int s_have;  // after the loop: a model was found (separate word: s_ctrl may still be being read)
  __shared__ int s_warp[NT / 32];
  __shared__ int s_cnt[NT / 32];
  __shared__ double s_out[9];

  const int f = blockIdx.x + w.frame_base, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ilcc2_diag* d = w.diag + f;
  if (d->status != ILCC2_OK) {
    if (tid == 0) w.brd_count[f] = 0;
    return;
  }
  const int C = w.sel_count[f];
  const int* S = w.sel_idx + (size_t)f * w.cap;
  const float4* P = w.roi_pts + (size_t)f * w.cap;
  int* perm = w.tmp_idx + (size_t)f * w.cap;
  int* brd = w.brd_idx + (size_t)f * w.cap;
  const double thr = p.ransac_thr;

  for (int i = tid; i < C; i += NT) perm[i] = i;
  __syncthreads();

  // RANSAC control state lives in thread 0's registers
  int iterations = 0, skipped = 0, n_best = -0x7fffffff, draw = 0;
  bool have_model = false;
  double k = 1.0;
  float best[4] = {0.f, 0.f, 0.f, 0.f};
  const int max_skip = p.ransac_max_iter * 10;
  for (;;) {
    if (tid == 0) {
      int ctrl = 0;
      if (C >= 3) {
        while ((double)iterations < k && skipped < max_skip) {
          bool good = false;
          float4 q0, q1, q2;
          for (int chk = 0; chk < 1000 && !good; ++chk) {
            if (draw + 3 > kRndTable) break;
            for (int i = 0; i < 3; ++i) {
              const int j = i + (int)(w.rnd_table[draw++] % (unsigned)(C - i));
              const int t = perm[i];
              perm[i] = perm[j];
              perm[j] = t;
            }
            q0 = P[S[perm[0]]];
            q1 = P[S[perm[1]]];
            q2 = P[S[perm[2]]];
            good = sample_good(q0, q1, q2);
          }
          if (!good) break;
          float m[4];
          if (!plane_from_3(q0, q1, q2, m)) {
            ++skipped;
            continue;
          }
          s_model[0] = m[0];
          s_model[1] = m[1];
          s_model[2] = m[2];
          s_model[3] = m[3];
          ctrl = 1;
          break;
        }
      }
      s_ctrl = ctrl;
    }
    __syncthreads();
    if (s_ctrl == 0) break;
    const float m0 = s_model[0], m1 = s_model[1], m2 = s_model[2], m3 = s_model[3];
    int cnt = 0;
    for (int i = tid; i < C; i += NT) {
      const float4 q = __ldg(P + S[i]);
      cnt += ((double)plane_dist(m0, m1, m2, m3, q.x, q.y, q.z) < thr) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
      int total = 0;
      for (int i = 0; i < NT / 32; ++i) total += s_cnt[i];
      if (total > n_best) {
        n_best = total;
        best[0] = m0;
        best[1] = m1;
        best[2] = m2;
        best[3] = m3;
        have_model = true;
        const double wr = (double)n_best * (1.0 / (double)C);
        double p_no = 1.0 - pow(wr, 3.0);
        p_no = fmax(p_no, DBL_EPSILON);
        p_no = fmin(p_no, 1.0 - DBL_EPSILON);
        k = p.ransac_log_prob / log(p_no);
      }
      ++iterations;
      if (iterations > p.ransac_max_iter) k = -1.0;  // PCL breaks out of the loop here
    }
    __syncthreads();
  }
  if (tid == 0) {
    s_have = have_model ? 1 : 0;
    s_model[0] = best[0];
    s_model[1] = best[1];
    s_model[2] = best[2];
    s_model[3] = best[3];
    d->ransac_iterations = iterations;
  }
  __syncthreads();
  if (s_have == 0) {
    if (tid == 0) {
      d->status = ILCC2_RANSAC_FAILED;
      w.brd_count[f] = 0;
    }
    return;
  }
  // inliers of the best sample model, input order (ROI indices), into perm[] (no longer needed)
  float m0 = s_model[0], m1 = s_model[1], m2 = s_model[2], m3 = s_model[3];
  int* inl = perm;
  __syncthreads();
  int ni = 0;
  {
    int rn = tid < C ? S[tid] : 0;  // one tile ahead of the compaction barriers, as below
    float4 qn = __ldg(P + rn);
    for (int base = 0; base < C; base += NT) {
      const int i = base + tid;
      const int ri = rn;
      const float4 q = qn;
      if (i + NT < C) {
        rn = S[i + NT];
        qn = __ldg(P + rn);
      }
      const bool pr = i < C && (double)plane_dist(m0, m1, m2, m3, q.x, q.y, q.z) < thr;
      int tile_total;
      const int off = block_ordered_offset<NT>(pr, s_warp, tile_total);
      if (pr) inl[ni + off] = ri;
      ni += tile_total;
    }
  }
  __syncthreads();
  int nb;
  if (ni >= 4) {
    // LS refit: moments about the first inlier, fp64, canonical 128-wide order
    const float4 ref = P[inl[0]];
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int kk = tid; kk < ni; kk += NT) {
      const float4 q = __ldg(P + inl[kk]);
      const double dx = (double)q.x - (double)ref.x;
      const double dy = (double)q.y - (double)ref.y;
      const double dz = (double)q.z - (double)ref.z;
      v[0] += dx;
      v[1] += dy;
      v[2] += dz;
      v[3] += dx * dx;
      v[4] += dx * dy;
      v[5] += dx * dz;
      v[6] += dy * dy;
      v[7] += dy * dz;
      v[8] += dz * dz;
    }
    block_reduce128<9>(v, s_part, s_out);
    if (tid == 0) {
      const double inv_n = 1.0 / (double)ni;
      const double mx = v[0] * inv_n, my = v[1] * inv_n, mz = v[2] * inv_n;
      double A[6];
      A[0] = v[3] * inv_n - mx * mx;
      A[1] = v[4] * inv_n - mx * my;
      A[2] = v[5] * inv_n - mx * mz;
      A[3] = v[6] * inv_n - my * my;
      A[4] = v[7] * inv_n - my * mz;
      A[5] = v[8] * inv_n - mz * mz;
      double ev[3], V[3][3];
      jacobi3d(A, ev, V);
      int c0 = 0;
      if (ev[1] < ev[c0]) c0 = 1;
      if (ev[2] < ev[c0]) c0 = 2;
      double nx = V[0][c0], ny = V[1][c0], nz = V[2][c0];
      const double nn = sqrt(nx * nx + ny * ny + nz * nz);
      nx /= nn;
      ny /= nn;
      nz /= nn;
      const double cx = (double)ref.x + mx, cy = (double)ref.y + my, cz = (double)ref.z + mz;
      s_model[0] = (float)nx;
      s_model[1] = (float)ny;
      s_model[2] = (float)nz;
      s_model[3] = (float)(-(nx * cx + ny * cy + nz * cz));
    }
    __syncthreads();
    m0 = s_model[0];
    m1 = s_model[1];
    m2 = s_model[2];
    m3 = s_model[3];
    nb = 0;
    int rn = tid < C ? S[tid] : 0;  // next tile's index and point, fetched one tile ahead of the compaction barriers
    float4 qn = __ldg(P + rn);
    for (int base = 0; base < C; base += NT) {
      const int i = base + tid;
      const int ri = rn;
      const float4 q = qn;
      bool pr = false;
      if (i + NT < C) {
        rn = S[i + NT];
        qn = __ldg(P + rn);
      }
      if (i < C) pr = (double)plane_dist(m0, m1, m2, m3, q.x, q.y, q.z) < thr;
      int tile_total;
      const int off = block_ordered_offset<NT>(pr, s_warp, tile_total);
      if (pr) brd[nb + off] = ri;
      nb += tile_total;
    }
  } else {
    for (int i = tid; i < ni; i += NT) brd[i] = inl[i];
    nb = ni;
  }
  if (tid == 0) {
    d->n_ransac_inliers = ni;
    d->n_board = nb;
    d->plane[0] = m0;
    d->plane[1] = m1;
    d->plane[2] = m2;
    d->plane[3] = m3;
    w.brd_count[f] = nb;
    if (nb < 3 || nb < p.min_board) d->status = (ni == 0 && nb == 0) ? ILCC2_RANSAC_FAILED : ILCC2_PLANE_TOO_SMALL;
  }
}


// ------------------------------------------------------------------------------------------------
// K4  pca_gray: transformbyPCA (centroid, covariance/N, eigenvectors, T = [V^T | -V^T c]) + calHist /
// get_gray_zone + the black/white classification of get_theta_t; emits the optimiser's point list.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void xform_pt(const float* T, float x, float y, float z, float out[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float v = __fmul_rn(T[4 * r + 0], x);
    v = __fadd_rn(v, __fmul_rn(T[4 * r + 1], y));
    v = __fadd_rn(v, __fmul_rn(T[4 * r + 2], z));
    out[r] = __fadd_rn(v, T[4 * r + 3]);
  }
}

__device__ __forceinline__ void pca_gray_stage(const Workspace& w, const DevParams& p, double* s_part /* [6 * 128] */) {
  constexpr int NT = kPlaneThreads;
  extern __shared__ int s_hist[];  // hist_bins + 1 counters (dynamic: 404 B for the reference's 100 bins)
  __shared__ double s_out[6];
  __shared__ float s_T[16];
  __shared__ float s_c[3];
  __shared__ float s_mm[2 * (NT / 32)];
  __shared__ double s_gray[2];
  __shared__ int s_status;
  __shared__ int s_warp[NT / 32];

  const int f = blockIdx.x + w.frame_base, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ilcc2_diag* d = w.diag + f;
  if (d->status != ILCC2_OK) {
    if (tid == 0) w.opt_count[f] = 0;
    return;
  }
  const int B = w.brd_count[f];
  const int* BI = w.brd_idx + (size_t)f * w.cap;
  const float4* P = w.roi_pts + (size_t)f * w.cap;
  const int NB = p.hist_bins;

  // centroid + intensity statistics
  double v[4] = {0, 0, 0, 0};
  float imin = FLT_MAX, imax = -FLT_MAX;
  for (int k = tid; k < B; k += NT) {
    const float4 q = __ldg(P + BI[k]);
    v[0] += (double)q.x;
    v[1] += (double)q.y;
    v[2] += (double)q.z;
    v[3] += (double)q.w;
    imin = fminf(imin, q.w);
    imax = fmaxf(imax, q.w);
  }
  block_reduce128<4>(v, s_part, s_out);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    imin = fminf(imin, __shfl_xor_sync(FULL, imin, o));
    imax = fmaxf(imax, __shfl_xor_sync(FULL, imax, o));
  }
  if (lane == 0) {
    s_mm[warp] = imin;
    s_mm[NT / 32 + warp] = imax;
  }
  if (tid == 0) {
    s_c[0] = (float)(v[0] / (double)B);
    s_c[1] = (float)(v[1] / (double)B);
    s_c[2] = (float)(v[2] / (double)B);
  }
  for (int i = tid; i <= NB; i += NT) s_hist[i] = 0;
  __syncthreads();
  const double isum = v[3];
  for (int i = 0; i < NT / 32; ++i) {
    imin = fminf(imin, s_mm[i]);
    imax = fmaxf(imax, s_mm[NT / 32 + i]);
  }
  const float c0 = s_c[0], c1 = s_c[1], c2 = s_c[2];
  const double mn = (double)imin, mx = (double)imax;
  const double mean = isum / (double)B;
  const double factor = (double)NB / (mx - mn);
  // covariance (centred with the float centroid, products in double) + histogram
  double m[6] = {0, 0, 0, 0, 0, 0};
  for (int k = tid; k < B; k += NT) {
    const float4 q = __ldg(P + BI[k]);
    const double dx = (double)__fsub_rn(q.x, c0);
    const double dy = (double)__fsub_rn(q.y, c1);
    const double dz = (double)__fsub_rn(q.z, c2);
    m[0] += dx * dx;
    m[1] += dx * dy;
    m[2] += dx * dz;
    m[3] += dy * dy;
    m[4] += dy * dz;
    m[5] += dz * dz;
    if (mx > mn) {
      int bin = (int)round(((double)q.w - mn) * factor);
      bin = min(max(bin, 0), NB);
      atomicAdd(&s_hist[bin], 1);
    }
  }
  block_reduce128<6>(m, s_part, s_out);
  if (tid == 0) {
    float cov[9];
    const float cxx = (float)(m[0] / (double)B), cxy = (float)(m[1] / (double)B), cxz = (float)(m[2] / (double)B);
    const float cyy = (float)(m[3] / (double)B), cyz = (float)(m[4] / (double)B), czz = (float)(m[5] / (double)B);
    cov[0] = cxx; cov[1] = cxy; cov[2] = cxz;
    cov[3] = cxy; cov[4] = cyy; cov[5] = cyz;
    cov[6] = cxz; cov[7] = cyz; cov[8] = czz;
    float ev[3], Q[3][3];
    eigen3f(cov, ev, Q);
    float v0[3] = {Q[0][0], Q[1][0], Q[2][0]}, v1[3] = {Q[0][1], Q[1][1], Q[2][1]}, v2[3];
    v2[0] = v0[1] * v1[2] - v0[2] * v1[1];
    v2[1] = v0[2] * v1[0] - v0[0] * v1[2];
    v2[2] = v0[0] * v1[1] - v0[1] * v1[0];
    const float* rows[3] = {v0, v1, v2};
    for (int r = 0; r < 3; ++r) {
      s_T[4 * r + 0] = rows[r][0];
      s_T[4 * r + 1] = rows[r][1];
      s_T[4 * r + 2] = rows[r][2];
      float t = rows[r][0] * c0;
      t = t + rows[r][1] * c1;
      t = t + rows[r][2] * c2;
      s_T[4 * r + 3] = -1.0f * t;
    }
    s_T[12] = 0.f; s_T[13] = 0.f; s_T[14] = 0.f; s_T[15] = 1.f;
    for (int i = 0; i < 16; ++i) d->pca[i] = s_T[i];
    for (int i = 0; i < 3; ++i) d->eigvals[i] = ev[i];
    // calHist mode rule
    int st = ILCC2_OK;
    double low = -1.0, high = -1.0;
    d->hist_min = mn;
    d->hist_max = mx;
    d->hist_mean = mean;
    if (!(mx > mn) || NB <= 0) {
      st = ILCC2_DEGENERATE_INTENSITY;
    } else {
      const double bin_width = (mx - mn) / (double)NB;
      bool low_found = false, high_found = false;
      int prev_count = 0x7fffffff;
      for (;;) {
        int bestc = -1, best_bin = -1;
        for (int b = 0; b < NB; ++b) {
          const int h = s_hist[b];
          if (h < prev_count && h > bestc) {
            bestc = h;
            best_bin = b;
          }
        }
        if (bestc < 0) break;
        prev_count = bestc;
        const double edge = bin_width * (double)best_bin + mn;
        if (edge > mean && !high_found) {
          high_found = true;
          high = edge;
        }
        if (edge < mean && !low_found) {
          low_found = true;
          low = edge;
        }
        if (low_found && high_found) break;
      }
      if (!low_found || !high_found) st = ILCC2_DEGENERATE_INTENSITY;
    }
    d->r_low = low;
    d->r_high = high;
    const double rate = p.gray_rate;
    double g0 = 0.0, g1 = 0.0;
    if (st == ILCC2_OK) {
      g0 = ((rate - 1.0) * low + high) / rate;
      g1 = (low + (rate - 1.0) * high) / rate;
    }
    d->gray[0] = g0;
    d->gray[1] = g1;
    s_gray[0] = g0;
    s_gray[1] = g1;
    s_status = st;
    if (st != ILCC2_OK) d->status = st;
  }
  __syncthreads();
  if (s_status != ILCC2_OK) {
    if (tid == 0) w.opt_count[f] = 0;
    return;
  }
  // classify (Optimization.cpp:114-125) and emit (y', z', white) in board order
  const double g0 = s_gray[0], g1 = s_gray[1];
  float2* oyz = w.opt_yz + (size_t)f * w.cap;
  uint8_t* owh = w.opt_white + (size_t)f * w.cap;
  const unsigned long long keep = l2_evict_last_policy();
  int ns = 0;
  // the index -> point gather of the NEXT tile is in flight while this tile goes through the two barriers of the ordered
  // compaction (21 % of the kernel's samples sat on this load)
  float4 qn = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < B) qn = __ldg(P + BI[tid]);
  for (int base = 0; base < B; base += NT) {
    const int k = base + tid;
    bool pr = false;
    int white = 0;
    const float4 q = qn;
    if (k + NT < B) qn = __ldg(P + BI[k + NT]);
    if (k < B) {
      const double I = (double)q.w;
      if (I < g0) {
        pr = true;
        white = 0;
      } else if (I > g1) {
        pr = true;
        white = 1;
      }
    }
    int tile_total;
    const int off = block_ordered_offset<NT>(pr, s_warp, tile_total);
    if (pr) {
      float t[3];
      xform_pt(s_T, q.x, q.y, q.z, t);
      st_keep(oyz + ns + off, make_float2(t[1], t[2]), keep);
      owh[ns + off] = (uint8_t)white;
    }
    ns += tile_total;
  }
  if (tid == 0) {
    w.opt_count[f] = ns;
    d->n_opt = ns;
    if (ns == 0) d->status = ILCC2_NO_OPT_POINTS;
  }
}

// K3 + K4 in one launch: the same 128-thread CTA per frame runs the plane stage and then, on the board indices it has
// just written (still in L1 / L2), the PCA + gray-zone stage: one launch and one pass over the frame's ROI points less.
// Footprint: the two stages share one reduction buffer and the histogram is sized by hist_bins, 11 KB of shared memory
// and 64 registers per thread instead of 33 KB / 72 -- a CTA of this kernel spends most of its life waiting for its
// thread 0 (RANSAC control, the two eigen-solvers), so what it costs the machine is the resources it holds meanwhile.
__global__ void __launch_bounds__(kPlaneThreads, 8) plane_pca_kernel(const Workspace w, const DevParams p) {
  __shared__ double s_part[9 * 128];
  plane_stage(w, p, s_part);
  __syncthreads();  // brd_idx / brd_count / diag.status written by this CTA above are visible to all of its threads
  pca_gray_stage(w, p, s_part);
}

void launch_plane_pca(const Workspace& w, const DevParams& p, cudaStream_t s) {
  plane_pca_kernel<<<w.n_frames, kPlaneThreads, sizeof(int) * (size_t)(p.hist_bins + 1), s>>>(w, p);
}

// ------------------------------------------------------------------------------------------------
// K5  board_fit: the virtual-chessboard pose optimisation (one ceres::Solve per phase in the reference).
// One CTA per (frame, colour hypothesis); the CTA is a team of T warps that share the solve (T = 8 / 4 / 2 / 1 chosen
// per launch, see launch_board_fit).  The frame's points (float2 y',z') are read through L1, one colour bit per point
// and the candidate lists live in shared memory.  A residual evaluation has two passes, each warp on its share of
// the points:
//   classify (fp32, every point): which virtual cell does the point fall in, with an error margin derived from
//     a forward error bound of the fp32 arithmetic.  Points that are CERTAINLY inside a cell of their own
//     colour (about 90 %) have a residual of exactly zero and are dropped; the others are compacted (ballot)
//     into two dense lists: L1 = colour mismatch or not certain (near a cell border / the board edge),
//     L2 = certainly outside the board (only contributes while useOutofBoard is on).
//   evaluate (fp64, listed points only, all lanes busy): the exact residual of Optimization.h:31-107, Huber
//     cost, and -- when the step is accepted -- the 3-column Jacobian, J^T r and J^T J, from the SAME lists
//     (Ceres evaluates the accepted point a second time; here nothing is recomputed for the dropped points).
// Sums are lane-strided over the lists + xor butterfly, then combined over the team's warps in fixed order through
// one named barrier (team_sum): deterministic, but grouped differently from the CPU oracle's sums (agreement ~1e-15
// relative, see tests).  The 3x3 dogleg step (subspace dogleg, Optimization.cpp:145-149) is computed redundantly by
// all lanes of all warps from state kept in shared memory (ScalarWs): no exchange besides the sums.  The CTA of a
// frame that finishes second picks the lower final cost and emits the corners (to the caller's buffers and to the
// gather targets of the other ranks, if any).
// ------------------------------------------------------------------------------------------------
#ifndef FIT_PAIR
#define FIT_PAIR 1
#endif
#if !FIT_PAIR
#error "FIT_PAIR = 0 (one CTA per hypothesis, round 1) is no longer maintained: the overflow-pool slots are per frame"
#endif
struct VbModel {
  int tlw;
  double Wd, Hd, g, inv_g, wg2, hg2, ha, hb, two_ha;
  float ymax_zmax;  // max|y'| + max|z'| over the frame's points (for the fp32 error bound)
};

// list entries per warp: the warps of a team take the 64-point groups of a chunk in turn
__host__ __device__ inline int fit_list_cap(int team_cap, int team) { return ((team_cap / 64 + team - 1) / team) * 64; }

template <typename IdxT>
struct FitPts {
  const float2* yz;
  const uint32_t* wbits;
  int n;
  float* timing; // debug builds (FIT_TIMING): kilo-cycles spent in evaluation / scalar solver
  int bits_bytes; // shared-memory instantiation: size of the colour-bit region in front of the lists
  IdxT* list;    // this warp's list buffer: L1 grows up from 0, L2 grows down from list_cap-1
  int list_cap;  // entries per warp: enough for the warp's share of a chunk (every T-th group of 64 points)
  int team_cap;  // points per chunk
};

// Lattice coordinates of a point, (y' + W g / 2) / g (Optimization.h:45-46), the way Ceres evaluates them:
//   rotation  AngleAxisRotatePoint((theta, 0, 0)): y' = y cos - (w0 z) sin, z' = z cos + (w0 y) sin with
//             w0 = |theta| * (1 / |theta|) (1 or 1 - 2^-53; 1 in the small-angle branch) -- `w0` below;
//   JET true  (residuals + Jacobian at an accepted point, T = Jet): Jet / Jet multiplies by the reciprocal of g;
//   JET false (cost of a candidate, T = double): a true division, taken here as q = x (1/g), q + fma(-q, g, x) (1/g),
//             which is the correctly rounded quotient for the correctly rounded reciprocal (Markstein), 2 FMAs instead of
//             the ~50-instruction IEEE division sequence.
template <bool JET>
__device__ __forceinline__ void vb_lattice(const VbModel& M, double y, double z, double c, double s, double w0, double ty,
                                           double tz, double& fi, double& fj) {
  const double yy = (y * c - (w0 * z) * s) + ty;
  const double zz = (z * c + (w0 * y) * s) + tz;
  const double xi = yy + M.wg2, xj = zz + M.hg2;
  const double qi = xi * M.inv_g, qj = xj * M.inv_g;
  if (JET) {
    fi = qi;
    fj = qj;
  } else {
    fi = fma(fma(-qi, M.g, xi), M.inv_g, qi);
    fj = fma(fma(-qj, M.g, xj), M.inv_g, qj);
  }
}

// exact residual of one point (Optimization.h:31-107) with the signs of its autodiff Jacobian, branch-free
// except for the (warp-uniform) useOutofBoard switch. wx = laser_white ^ topleftWhite.
template <bool JET>
__device__ __forceinline__ double vb_eval(const VbModel& M, bool oob, double y, double z, unsigned wx, double c,
                                          double s, double w0, double ty, double tz, double& si, double& sj) {
  double fi, fj;
  vb_lattice<JET>(M, y, z, c, s, w0, ty, tz, fi, fj);
  const bool in = fi > 0.0 && fi < M.Wd && fj > 0.0 && fj < M.Hd;
  // floor for 0 <= f < 2^51: (f + 2^52) - 2^52 rounds to the nearest integer, step back if it went up
  const double kMagic = 4503599627370496.0;
  const double ti = fi + kMagic, tj = fj + kMagic;
  const double ri = ti - kMagic, rj = tj - kMagic;
  const bool ci = ri > fi, cj = rj > fj;
  const double ifl = ci ? ri - 1.0 : ri;
  const double jfl = cj ? rj - 1.0 : rj;
  const unsigned par = ((unsigned)__double2loint(ti) ^ (unsigned)__double2loint(tj) ^ (unsigned)ci ^ (unsigned)cj) & 1u;
  const bool mismatch = ((wx ^ par) & 1u) != 0u;  // laser colour != colour of the virtual cell
  const double di = fi - ifl, dj = fj - jfl;
  const bool bi = di > 0.5, bj = dj > 0.5;
  const double ei = bi ? (ifl + 1.0) - fi : di;
  const double ej = bj ? (jfl + 1.0) - fj : dj;
  double r = mismatch ? ei + ej : 0.0;
  si = bi ? -1.0 : 1.0;
  sj = bj ? -1.0 : 1.0;
  if (oob) {
    const double fiw = fi - M.Wd, fjh = fj - M.Hd;
    const double a0 = fabs(fi), a1 = fabs(fiw);
    const double b0 = fabs(fj), b1 = fabs(fjh);
    const bool li = a0 < a1, lj = b0 < b1;
    const double ro = (li ? a0 : a1) + (lj ? b0 : b1);
    const double sio = (li ? fi : fiw) < 0.0 ? -1.0 : 1.0;
    const double sjo = (lj ? fj : fjh) < 0.0 ? -1.0 : 1.0;
    r = in ? r : ro;
    si = in ? si : sio;
    sj = in ? sj : sjo;
  } else {
    r = in ? r : 0.0;
  }
  return r;
}

// residual (and Jacobian signs) of a point that is certainly outside the board: the useOutofBoard branch,
// Optimization.h:86-101 -- same expressions as the corresponding part of vb_eval.
template <bool JET>
__device__ __forceinline__ double vb_eval_out(const VbModel& M, double y, double z, double c, double s, double w0,
                                              double ty, double tz, double& si, double& sj) {
  double fi, fj;
  vb_lattice<JET>(M, y, z, c, s, w0, ty, tz, fi, fj);
  const double fiw = fi - M.Wd, fjh = fj - M.Hd;
  const double a0 = fabs(fi), a1 = fabs(fiw);
  const double b0 = fabs(fj), b1 = fabs(fjh);
  const bool li = a0 < a1, lj = b0 < b1;
  si = (li ? fi : fiw) < 0.0 ? -1.0 : 1.0;
  sj = (lj ? fj : fjh) < 0.0 ? -1.0 : 1.0;
  return (li ? a0 : a1) + (lj ? b0 : b1);
}

// IEEE double division / square root expand to ~50 instructions each at every use; the scalar solver below has
// ~100 of them.  One out-of-line copy keeps the kernel image (and its instruction-cache footprint) small.
__device__ __noinline__ double ddiv_ieee(double a, double b) { return a / b; }
__device__ __noinline__ double dsqrt_ieee(double a) { return sqrt(a); }

// Cheap reciprocal / square root for the INSIDE of self-correcting iterations (the secular-equation Newton of
// tr_boundary_2d): MUFU seed (about 20 bits) + two Newton steps, about 1-2 ulp, 6-9 instructions instead of the
// ~50 of the IEEE sequences.  Arguments there are finite, positive and far from the exponent limits.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
__device__ __forceinline__ double fast_sqrt_pos(double a) {  // a > 0
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  y = y * fma(-0.5 * a * y, y, 1.5);
  y = y * fma(-0.5 * a * y, y, 1.5);
  double g = a * y;
  g = fma(0.5 * y, fma(-g, g, a), g);
  return g;
}

__device__ __forceinline__ double ddiv(double a, double b) { return ddiv_ieee(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return dsqrt_ieee(a); }

// sincos() inlines a large slow path; keep one copy of it in the kernel
__device__ __noinline__ void sincos_once(double x, double* s, double* c) { sincos(x, s, c); }

__device__ __forceinline__ double huber_cost(const VbModel& M, double r) {
  const double sq = r * r;
  return (sq > M.hb) ? 0.5 * (M.two_ha * r - M.hb) : 0.5 * sq;
}

// sm_100 packed fp32: two IEEE fp32 operations per instruction on a 64-bit register pair
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// fp32 classification of points [k0, k1) (k1 - k0 <= list_cap) at pose (c, s, ty, tz) -> lists L1 (up), L2 (down).
//   fi = ((y c - z s) + ty + W g/2) / g  evaluated with two FMAs; forward error bound 5u(|y|+|z|+|t|)/g with
//   u = 2^-24 (+1 rounding for the pre-scaled constants): 8u is used, times a safety factor 4 -> margin m.
//   A point is CERTAINLY inside cell (floor fi, floor fj) of the board when both coordinates are at least m away
//   from every integer and from the board edges; it is CERTAINLY outside when one coordinate is m beyond an edge.
template <typename IdxT>
__device__ __forceinline__ void classify_pass(const float2* __restrict__ yz, const uint32_t* __restrict__ wbits,
                                              IdxT* list, int list_cap, const VbModel& M, bool oob, double c, double s,
                                              double ty, double tz, int lane, int k0, int k1, int stride, int& n1,
                                              int& n2) {
  // points k0, k0 + stride, ... in groups of 32 * U: the warps of a team take the groups of a chunk in turn, so that
  // each gets the same mix of the board's regions (balanced list lengths -> short waits at the team barrier)
  const float ig = (float)M.inv_g;
  const float cg = (float)c * ig, sg = (float)s * ig;
  const float tyw = (float)(ty + M.wg2) * ig, tzh = (float)(tz + M.hg2) * ig;
  const float Wh = 0.5f * (float)M.Wd, Hh = 0.5f * (float)M.Hd;
  const float bound = (8.0f * 5.9604645e-8f) * (M.ymax_zmax * ig + fmaxf(fabsf(tyw), fabsf(tzh)));
  const float m = 4.0f * bound + 1e-6f;
  const float kMagicF = 12582912.0f;  // 1.5 * 2^23: (f + magic) - magic = rint(f) for |f| < 2^22
  const unsigned tlw = (unsigned)M.tlw;
  int c1 = 0, c2 = 0;
  const unsigned lt = (1u << lane) - 1u;
#ifndef CLS_UNROLL
#define CLS_UNROLL 2
#endif
  constexpr int U = CLS_UNROLL;  // independent points per lane per iteration: the warp is latency-bound, not issue-bound
  const f32x2 cg2 = pack2(cg, cg), sg2 = pack2(sg, sg), t2 = pack2(tyw, tzh), magic2 = pack2(kMagicF, kMagicF),
              nmagic2 = pack2(-kMagicF, -kMagicF), nhalf2 = pack2(-Wh, -Hh);
  // software pipeline: the points of the next iteration are loaded before this iteration's arithmetic (the first
  // FFMA2 otherwise waits ~7 % of the kernel's samples for its load).  No clamping anywhere: lanes at or beyond k1 read
  // whatever follows (opt_yz is allocated 128 entries longer than the last frame needs, the colour bits are followed by
  // the lists) and are masked out by `valid` below.
  float2 qn[U];
#pragma unroll
  for (int u = 0; u < U; ++u) qn[u] = __ldg(yz + k0 + 32 * u + lane);
  for (int base = k0; base < k1; base += stride) {
    bool is1[U], is2[U];
    float2 qc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      qc[u] = qn[u];
      qn[u] = __ldg(yz + base + stride + 32 * u + lane);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = base + 32 * u + lane;
      const bool valid = k < k1;
      const float2 q = qc[u];
      const unsigned wword = wbits[(base + 32 * u) >> 5];
#ifndef CLS_NO_F32X2
      // Blackwell packed fp32 (FFMA2 / FADD2): the i and j coordinates go through identical operations, so every
      // step is one instruction for both (bit-identical to the scalar sequence: each half is an IEEE fp32 op)
      const f32x2 F = fma2(pack2(q.x, q.y), cg2, fma2(pack2(-q.y, q.x), sg2, t2));
      const f32x2 Tm = add2(F, magic2);
      const f32x2 D = sub2(F, add2(Tm, nmagic2));  // signed distance to the nearest integer
      const f32x2 E0 = add2(F, nhalf2);            // (fi - Wh, fj - Hh)
      float fi, fj, ti, tj, di, dj, ei, ej;
      unpack2(F, fi, fj);
      unpack2(Tm, ti, tj);
      unpack2(D, di, dj);
      unpack2(E0, ei, ej);
      (void)fi;
      (void)fj;
      const float edge = fmaxf(fabsf(ei) - Wh, fabsf(ej) - Hh);  // < 0 inside the board, > 0 outside
#else
      const float fi = fmaf(q.x, cg, fmaf(-q.y, sg, tyw));
      const float fj = fmaf(q.y, cg, fmaf(q.x, sg, tzh));
      const float ti = fi + kMagicF, tj = fj + kMagicF;
      const float di = fi - (ti - kMagicF), dj = fj - (tj - kMagicF);  // signed distance to the nearest integer
      const float edge = fmaxf(fabsf(fi - Wh) - Wh, fabsf(fj - Hh) - Hh);  // < 0 inside the board, > 0 outside
#endif
      // cell parity: parity(rint) ^ (d < 0)
      const unsigned par =
          (__float_as_uint(ti) ^ __float_as_uint(tj) ^ ((__float_as_uint(di) ^ __float_as_uint(dj)) >> 31)) & 1u;
      const bool interior = fminf(fabsf(di), fabsf(dj)) >= m;
      const bool certain_in = interior && edge <= -m;
      const bool outb = edge >= m;
      const bool mism = ((((wword >> lane) ^ tlw ^ par) & 1u) != 0u);
      is1[u] = valid && (certain_in ? mism : !outb);
      is2[u] = valid && outb;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = base + 32 * u + lane;
      const unsigned b1 = __ballot_sync(FULL, is1[u]);
      if (is1[u]) list[c1 + __popc(b1 & lt)] = (IdxT)k;
      c1 += __popc(b1);
      if (oob) {
        const unsigned b2 = __ballot_sync(FULL, is2[u]);
        if (is2[u]) list[list_cap - 1 - (c2 + __popc(b2 & lt))] = (IdxT)k;
        c2 += __popc(b2);
      }
    }
  }
  __syncwarp();
  n1 = c1;
  n2 = c2;
}

// accumulate one evaluated point: Huber cost (want_jac false) or J^T r / J^T J with the Huber corrector
__device__ __forceinline__ void accumulate_point(const VbModel& M, bool want_jac, double y, double z, double c, double s,
                                                 double r, double si, double sj, double (&a)[9]) {
  if (!want_jac) {
    a[0] += huber_cost(M, r);
    return;
  }
  const double dyy = -(y * s) - (z * c);
  const double dzz = (y * c) - (z * s);
  double j0v = (si * dyy) * M.inv_g + (sj * dzz) * M.inv_g;
  double j1 = si * M.inv_g;
  double j2 = sj * M.inv_g;
  double rr = r;
  const bool big = r * r > M.hb;
  if (__any_sync(FULL, big)) {  // Huber corrector: scale row and residual by sqrt(rho') = sqrt(a / r)
    // residuals here lie in (huber_delta, a few cells]: MUFU-seeded reciprocal and square root inline (1-2 ulp from the
    // IEEE results, which shows as <= 1e-15 in theta/t and not at all in the corners; ~35 instructions per list
    // iteration instead of two out-of-line IEEE routines: K5 2.17 -> 2.07 ms)
    const double rsafe = big ? r : 1.0;
    const double rho1 = M.ha * fast_rcp(rsafe);
    const double sc = big ? fast_sqrt_pos(rho1) : 1.0;
    j0v *= sc;
    j1 *= sc;
    j2 *= sc;
    rr = sc * r;
  }
  if (r == 0.0) {
    j0v = 0.0;
    j1 = 0.0;
    j2 = 0.0;
  }
  a[0] += j0v * rr;
  a[1] += j1 * rr;
  a[2] += j2 * rr;
  a[3] += j0v * j0v;
  a[4] += j0v * j1;
  a[5] += j0v * j2;
  a[6] += j1 * j1;
  a[7] += j1 * j2;
  a[8] += j2 * j2;
}

// exact fp64 evaluation of the listed points (all lanes busy; uniform trip counts because of __any_sync).
// JAC = false: a[0] += Huber cost (Ceres' T = double pass).  JAC = true: a[0..2] += J^T r, a[3..8] += J^T J (Jet pass).
template <typename IdxT, bool JAC>
__device__ __forceinline__ void list_pass(const float2* __restrict__ yz, const uint32_t* __restrict__ wbits,
                                          const IdxT* list, int list_cap, const VbModel& M, bool oob, double c, double s,
                                          double w0, double ty, double tz, int lane, int n1, int n2, double (&a)[9]) {
  // L1: colour mismatch or not certain -> full residual.  LIST_U points per lane and iteration: their residual chains
  // (~50 dependent fp64 operations each) are independent, the sums are still taken in list order.  Software pipeline:
  // index, point and colour bit of the next iteration are fetched (shared-memory list -> L1 gather, a dependent chain)
  // while this one is evaluated.
#ifndef LIST_U
#define LIST_U 1
#endif
  constexpr int LU = LIST_U;
  int kn[LU];
  float2 qn[LU];
  unsigned wn[LU];
#pragma unroll
  for (int u = 0; u < LU; ++u) {
    const int en = 32 * u + lane;
    kn[u] = en < n1 ? (int)list[en] : 0;
    qn[u] = __ldg(yz + kn[u]);
    wn[u] = wbits[kn[u] >> 5];
  }
  for (int e0 = 0; e0 < n1; e0 += 32 * LU) {
    bool act[LU];
    double y[LU], z[LU], r[LU], si[LU], sj[LU];
    unsigned wx[LU];
#pragma unroll
    for (int u = 0; u < LU; ++u) {
      act[u] = e0 + 32 * u + lane < n1;
      y[u] = (double)qn[u].x;
      z[u] = (double)qn[u].y;
      wx[u] = ((wn[u] >> (kn[u] & 31)) & 1u) ^ (unsigned)M.tlw;
      const int en = e0 + 32 * (LU + u) + lane;
      kn[u] = en < n1 ? (int)list[en] : 0;
      qn[u] = __ldg(yz + kn[u]);
      wn[u] = wbits[kn[u] >> 5];
    }
#pragma unroll
    for (int u = 0; u < LU; ++u) r[u] = vb_eval<JAC>(M, oob, y[u], z[u], wx[u], c, s, w0, ty, tz, si[u], sj[u]);
#pragma unroll
    for (int u = 0; u < LU; ++u) {
      if (!act[u]) r[u] = 0.0;
      accumulate_point(M, JAC, y[u], z[u], c, s, r[u], si[u], sj[u], a);
    }
  }
  if (oob) {
    for (int e0 = 0; e0 < n2; e0 += 32) {  // L2: certainly outside the board -> distance to the nearest edges
      const int e = e0 + lane;
      const bool act = e < n2;
      const int k = act ? (int)list[list_cap - 1 - e] : 0;
      const float2 q = yz[k];
      const double y = (double)q.x, z = (double)q.y;
      double si, sj;
      double r = vb_eval_out<JAC>(M, y, z, c, s, w0, ty, tz, si, sj);
      if (!act) r = 0.0;
      accumulate_point(M, JAC, y, z, c, s, r, si, sj, a);
    }
  }
}

// --- 3x3 trust-region machinery (all lanes execute it redundantly on identical inputs) ---
struct Dogleg {
  double radius, mu;
  int reuse;
  double diag[3], grad[3], gn[3], alpha;
  int sub1d;
  double U[3][2], sub_g[2], sub_B[3];
  double step_norm;
  // values that only change when the Gauss-Newton step / subspace model is rebuilt, cached for the (many) rejected
  // iterations that reuse it: |gn|, and eigenvalues of sub_B and |sub_g| for the boundary problem
  double gnn, tb_l1, tb_l2, tb_gn;
  double inv_diag[3];  // 1 / diag: x / diag is taken as x * inv_diag (<= 1.5 ulp instead of 0.5; one division per axis)
};

// Per-warp scratch of the scalar solver, kept in SHARED memory: the out-of-line dogleg routines take these by
// reference, which would otherwise put them on the local-memory stack -- 128 B of L1 per 4-byte slot and warp, i.e.
// ~8 KB per warp that competes with the frame's points for L1 (ncu: long-scoreboard stalls in the scalar code and
// 85 MB of local-memory write-back per launch).  All lanes of a warp hold identical values, so they all store the
// same data to the same address.
struct ScalarWs {
  Dogleg S;
  double Hs[6], gs[3], step[3];
  // iterate / bookkeeping of solve_tr: only touched between evaluations, so keeping them here frees ~50 registers
  // inside the classification and list loops
  double xc[3], cand[3], g[3], H[6], scale[3], xbest[3];
  double x_cost, xnorm, best_cost, model_change, gmax;
  // candidate batch (solve_tr): entry k = candidate for radius 2^-k: cand[3], model decrease (<= 0: invalid step),
  // |step|, sin / cos of the candidate angle, unit-axis factor
  double batch[kFitBatch][8];
};

__device__ __forceinline__ double norm3(const double v[3]) { return dsqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// Read-modify-write of a warp-replicated scalar in shared memory (all 32 lanes hold and store the same value): every
// lane must have READ the old value before any lane stores the new one.  Converged lanes do that anyway; the barriers
// make it independent of convergence (compute-sanitizer's racecheck serialises lanes and does expose it otherwise).
__device__ __forceinline__ void ws_update(double& dst, double v) {
  __syncwarp();
  dst = v;
  __syncwarp();
}

__device__ __forceinline__ double quad3(const double H[6], const double u[3], const double v[3]) {
  const double h0 = H[0] * v[0] + H[1] * v[1] + H[2] * v[2];
  const double h1 = H[1] * v[0] + H[3] * v[1] + H[4] * v[2];
  const double h2 = H[2] * v[0] + H[4] * v[1] + H[5] * v[2];
  return u[0] * h0 + u[1] * h1 + u[2] * h2;
}

// 3x3 Cholesky solve; the three pivots are inverted once (3 rsqrt, 1 ulp each, instead of 3 sqrt + 13 divisions)
__device__ bool chol3_solve(const double A[6], const double b[3], double x[3]) {
  if (!(A[0] > 0.0)) return false;
  const double i00 = rsqrt(A[0]);
  const double l10 = A[1] * i00;
  const double l20 = A[2] * i00;
  double t = A[3] - l10 * l10;
  if (!(t > 0.0)) return false;
  const double i11 = rsqrt(t);
  const double l21 = (A[4] - l20 * l10) * i11;
  t = A[5] - l20 * l20 - l21 * l21;
  if (!(t > 0.0)) return false;
  const double i22 = rsqrt(t);
  const double y0 = b[0] * i00;
  const double y1 = (b[1] - l10 * y0) * i11;
  const double y2 = (b[2] - l20 * y0 - l21 * y1) * i22;
  x[2] = y2 * i22;
  x[1] = (y1 - l21 * x[2]) * i11;
  x[0] = (y0 - l10 * x[1] - l20 * x[2]) * i00;
  return isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]);
}

__device__ bool tr2_eval(double a, double b, double dd, double g0, double g1, double y, double x[2], double& n2,
                         double& q) {
  const double ay = a + y, dy = dd + y;
  const double det = ay * dy - b * b;
  if (!(ay > 0.0) || !(det > 0.0)) return false;
  if (det > 1e-280 && det < 1e280) {
    const double inv = fast_rcp(det);
    x[0] = -(dy * g0 - b * g1) * inv;
    x[1] = -(ay * g1 - b * g0) * inv;
    n2 = x[0] * x[0] + x[1] * x[1];
    q = x[0] * ((dy * x[0] - b * x[1]) * inv) + x[1] * ((ay * x[1] - b * x[0]) * inv);
    return true;
  }
  x[0] = ddiv(-(dy * g0 - b * g1), det);
  x[1] = ddiv(-(ay * g1 - b * g0), det);
  n2 = x[0] * x[0] + x[1] * x[1];
  const double w0 = ddiv(dy * x[0] - b * x[1], det);
  const double w1 = ddiv(ay * x[1] - b * x[0], det);
  q = x[0] * w0 + x[1] * w1;
  return true;
}

// min 1/2 x'Bx + g'x on |x| = radius: safeguarded Newton on 1/|x(y)| - 1/radius, x(y) = -(B + yI)^-1 g
__device__ __noinline__ bool tr_boundary_2d(const double B[3], const double g[2], double l1, double l2, double gn,
                                            double radius, double x_out[2]) {
  const double a = B[0], b = B[1], dd = B[2];
  x_out[0] = 0.0;
  x_out[1] = 0.0;
  if (!(gn > 0.0) || !(radius > 0.0)) return false;
  const double gr = ddiv(gn, radius);
  double lo = gr - l2;
  if (lo < -l1) lo = -l1;
  double hi = gr - l1;
  if (!(hi > lo)) hi = lo + fabs(lo) * 1e-12 + 1e-300;
  double y = 0.0;
  if (y < lo || y > hi) y = lo;
  double x[2] = {0.0, 0.0}, n2 = 0.0, q = 0.0;
  bool have = false;
  const double inv_radius = (radius > 1e-280 && radius < 1e280) ? fast_rcp(radius) : ddiv(1.0, radius);  // Newton update only
  double n_last = 0.0;
  for (int it = 0; it < 100; ++it) {
    if (!tr2_eval(a, b, dd, g[0], g[1], y, x, n2, q)) {
      lo = y;
      y = 0.5 * (lo + hi);
      if (!(y > lo)) break;
      continue;
    }
    have = true;
    const bool fast_ok = n2 > 1e-280 && n2 < 1e280 && fabs(q) > 1e-280 && fabs(q) < 1e280;
    const double n = fast_ok ? fast_sqrt_pos(n2) : dsqrt(n2);
    n_last = n;
    if (fabs(n - radius) <= 1e-14 * radius) break;
    if (n > radius) lo = y; else hi = y;
    double yn = fast_ok ? y + ((n - radius) * inv_radius) * (n2 * fast_rcp(q)) : y + ddiv(n - radius, radius) * ddiv(n2, q);
    if (!(yn > lo && yn < hi)) yn = 0.5 * (lo + hi);
    if (yn == y || !(hi > lo)) break;
    y = yn;
  }
  if (!have) return false;
  if (!(fabs(n_last - radius) <= 1e-6 * radius)) return false;  // n_last = |x| of the last evaluated multiplier
  x_out[0] = x[0];
  x_out[1] = x[1];
  return true;
}

// Trust-region step for radius `radius` from the model held in S (read-only here, so that different lanes can ask for
// different radii at once): traditional dogleg.  Returns |step| in the scaled space.
__device__ double dogleg_traditional(const Dogleg& S, double radius, double step[3]) {
  const double gnorm = norm3(S.grad), gnn = S.gnn;
  if (gnn <= radius) {
    for (int i = 0; i < 3; ++i) step[i] = S.gn[i] * S.inv_diag[i];
    return gnn;
  }
  if (gnorm * S.alpha >= radius) {
    const double sc = -ddiv(radius, gnorm);
    for (int i = 0; i < 3; ++i) step[i] = (sc * S.grad[i]) * S.inv_diag[i];
    return radius;
  }
  const double b_dot_a = -S.alpha * (S.grad[0] * S.gn[0] + S.grad[1] * S.gn[1] + S.grad[2] * S.gn[2]);
  const double a2 = (S.alpha * gnorm) * (S.alpha * gnorm);
  const double bma2 = a2 - 2.0 * b_dot_a + gnn * gnn;
  const double c = b_dot_a - a2;
  const double dd = dsqrt(c * c + bma2 * (radius * radius - a2));
  const double beta = (c <= 0.0) ? ddiv(dd - c, bma2) : ddiv(radius * radius - a2, dd + c);
  double dl[3];
  for (int i = 0; i < 3; ++i) dl[i] = (-S.alpha * (1.0 - beta)) * S.grad[i] + beta * S.gn[i];
  const double n = norm3(dl);
  for (int i = 0; i < 3; ++i) step[i] = dl[i] * S.inv_diag[i];
  return n;
}

// Subspace dogleg step (Ceres DoglegStrategy::ComputeSubspaceDoglegStep) for radius `radius`; returns |step| (scaled).
__device__ __noinline__ double dogleg_step_for_radius(const Dogleg& S, double radius, double step[3]) {
  if (S.gnn <= radius) {
    for (int i = 0; i < 3; ++i) step[i] = S.gn[i] * S.inv_diag[i];
    return S.gnn;
  }
  if (S.sub1d) {
    const double sc = -ddiv(radius, norm3(S.grad));
    for (int i = 0; i < 3; ++i) step[i] = (sc * S.grad[i]) * S.inv_diag[i];
    return radius;
  }
  double m2[2];
  if (!tr_boundary_2d(S.sub_B, S.sub_g, S.tb_l1, S.tb_l2, S.tb_gn, radius, m2)) return dogleg_traditional(S, radius, step);
  for (int i = 0; i < 3; ++i) step[i] = (S.U[i][0] * m2[0] + S.U[i][1] * m2[1]) * S.inv_diag[i];
  return radius;
}

// The part of DoglegStrategy::ComputeStep that does not depend on the radius: scaling diagonal, gradient, Cauchy step
// length, mu-regularised Gauss-Newton step, 2-D subspace model.  False = linear solver failure (an invalid step).
__device__ __noinline__ bool dogleg_build_model(Dogleg& S, const double Hs[6], const double gs[3]) {
  const double hd[3] = {Hs[0], Hs[3], Hs[5]};
  for (int i = 0; i < 3; ++i) {
    double v = hd[i];
    if (v < 1e-6) v = 1e-6;
    if (v > 1e32) v = 1e32;
    S.inv_diag[i] = rsqrt(v);  // v in [1e-6, 1e32]
    S.diag[i] = v * S.inv_diag[i];
  }
  for (int i = 0; i < 3; ++i) S.grad[i] = gs[i] * S.inv_diag[i];
  double sg[3];
  for (int i = 0; i < 3; ++i) sg[i] = S.grad[i] * S.inv_diag[i];
  const double jg2 = quad3(Hs, sg, sg);
  S.alpha = ddiv(S.grad[0] * S.grad[0] + S.grad[1] * S.grad[1] + S.grad[2] * S.grad[2], jg2);
  bool ok = false;
  double y[3] = {0.0, 0.0, 0.0};
  double mu = S.mu;
  while (mu < 1.0) {
    double A[6] = {Hs[0], Hs[1], Hs[2], Hs[3], Hs[4], Hs[5]};
    const double sm = dsqrt(mu);
    const double l0 = S.diag[0] * sm, l1 = S.diag[1] * sm, l2 = S.diag[2] * sm;
    A[0] += l0 * l0;
    A[3] += l1 * l1;
    A[5] += l2 * l2;
    if (chol3_solve(A, gs, y)) {
      ok = true;
      break;
    }
    mu *= 10.0;
  }
  ws_update(S.mu, mu);
  if (!ok) return false;
  for (int i = 0; i < 3; ++i) S.gn[i] = -S.diag[i] * y[i];
  const double n_g = norm3(S.grad), n_n = norm3(S.gn);
  S.gnn = n_n;
  const bool gfirst = n_g >= n_n;
  const double na = gfirst ? n_g : n_n;
  if (!(na > 0.0)) return false;
  double u0[3], bp[3], A1[3];
  const double inv_na = ddiv(1.0, na);
  for (int i = 0; i < 3; ++i) {
    u0[i] = (gfirst ? S.grad[i] : S.gn[i]) * inv_na;
    A1[i] = gfirst ? S.gn[i] : S.grad[i];
  }
  const double pr = A1[0] * u0[0] + A1[1] * u0[1] + A1[2] * u0[2];
  for (int i = 0; i < 3; ++i) bp[i] = A1[i] - pr * u0[i];
  const double nb = norm3(bp);
  if (!(nb > na * (2.0 * DBL_EPSILON))) {
    S.sub1d = 1;
  } else {
    S.sub1d = 0;
    double w0[3], w1[3];
    const double inv_nb = ddiv(1.0, nb);
    for (int i = 0; i < 3; ++i) {
      S.U[i][0] = u0[i];
      S.U[i][1] = bp[i] * inv_nb;
    }
    S.sub_g[0] = S.U[0][0] * S.grad[0] + S.U[1][0] * S.grad[1] + S.U[2][0] * S.grad[2];
    S.sub_g[1] = S.U[0][1] * S.grad[0] + S.U[1][1] * S.grad[1] + S.U[2][1] * S.grad[2];
    for (int i = 0; i < 3; ++i) {
      w0[i] = S.U[i][0] * S.inv_diag[i];
      w1[i] = S.U[i][1] * S.inv_diag[i];
    }
    S.sub_B[0] = quad3(Hs, w0, w0);
    S.sub_B[1] = quad3(Hs, w0, w1);
    S.sub_B[2] = quad3(Hs, w1, w1);
    const double tr = S.sub_B[0] + S.sub_B[2], hd2 = 0.5 * (S.sub_B[0] - S.sub_B[2]);
    const double disc = dsqrt(hd2 * hd2 + S.sub_B[1] * S.sub_B[1]);
    S.tb_l1 = 0.5 * tr - disc;
    S.tb_l2 = 0.5 * tr + disc;
    S.tb_gn = dsqrt(S.sub_g[0] * S.sub_g[0] + S.sub_g[1] * S.sub_g[1]);
  }
  return true;
}


enum { TERM_MAX_ITER = 0, TERM_GRADIENT = 1, TERM_PARAMETER = 2, TERM_FUNCTION = 3, TERM_MIN_RADIUS = 4, TERM_INVALID = 5 };

// Sum of per-warp totals over the T warps that share a solve (= the CTA).  Every warp ends up with the bit-identical
// total (fixed order: warp 0 first), so the scalar solver can run redundantly on all of them without any further
// exchange.  One barrier per call; the slots are double-buffered (`parity`), which is enough because a warp cannot
// run two reductions ahead of its team mates.  (An immediate barrier id: with a register id ptxas reserves all 16
// named barriers for the CTA, and the barrier pool of the SM then caps the number of resident CTAs.)
template <int NQ, int T>
__device__ __forceinline__ void team_sum(double (&v)[NQ], double* slots /* [2][T][9] */, int wt, int lane, int& parity, int team) {
  if (T == 1) return;
  double* buf = slots + parity * (T * 9);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) buf[wt * 9 + q] = v[q];
  }
  if (team == 0)  // the two teams of a CTA (one per colour hypothesis) synchronise independently
    asm volatile("bar.sync 1, %0;" ::"n"(32 * T) : "memory");
  else
    asm volatile("bar.sync 2, %0;" ::"n"(32 * T) : "memory");
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double t = buf[q];
#pragma unroll
    for (int w = 1; w < T; ++w) t += buf[w * 9 + q];
    v[q] = t;
  }
  parity ^= 1;
}

// One ceres::Solve (TrustRegionMinimizer, Ceres 1.13/1.14 defaults) on one warp.  Written as a single loop
// whose head evaluates the cost at the current candidate (the first candidate is the start point and is always
// "accepted"), so that the classification / evaluation code exists exactly once in the kernel image.
//
// Candidate batches.  Between two accepted steps the model (Gauss-Newton step, 2-D subspace model) is fixed and every
// rejection just halves the radius, so the candidates of a whole run of rejections are known in advance: the step for
// radius r 2^-k, k = 0, 1, ...  The solver's per-candidate scalar work (the secular-equation Newton of the boundary
// problem, the model decrease, sincos of the candidate angle) is therefore done ONCE per model, with lane k of the warp
// working on radius r 2^-k -- all lanes were executing this code redundantly anyway -- and parked in shared memory; a
// rejected iteration then costs one evaluation pass plus a table look-up instead of one pass plus ~2000 cycles of
// dependent fp64 arithmetic (3.3 candidates are evaluated per accepted step on the bench batch).  Every candidate is
// bit-identical to what the sequential formulation computes (same functions of the same radius).
// Rejected Gauss-Newton steps that would be recomputed identically (radius still above |GN|) are fast-forwarded:
// iteration counter and radius advance exactly as if the same candidate had been evaluated and rejected again.
// Inlined into its single call site (the phase loop of board_fit_kernel): the model constants then come straight
// from the kernel parameters and nothing is spilled around the call.
template <typename IdxT, int T>
__device__ __forceinline__ int solve_tr(const FitPts<IdxT>* Fp, const VbModel* Mp, int oob_i, int max_iter, int lane,
                                     double* x, double* final_cost, int* iters_out, int* n_evals_io, double* slots,
                                     int* parity_io, ScalarWs* ws_shared) {
  const float2* __restrict__ yz = Fp->yz;
  const uint32_t* wbits = Fp->wbits;
  IdxT* list = Fp->list;
  int list_cap = Fp->list_cap;
  if (std::is_same<IdxT, uint16_t>::value) {
    // shared-memory instantiation: derive the pointers from the shared array itself so that the compiler emits
    // LDS/STS with 32-bit addresses instead of generic loads/stores
    extern __shared__ unsigned char fit_sm[];
    wbits = reinterpret_cast<const uint32_t*>(fit_sm);
    list_cap = Fp->list_cap;
    list = reinterpret_cast<IdxT*>(fit_sm + Fp->bits_bytes) + (size_t)(threadIdx.x >> 5) * list_cap;
  }
  const int n = Fp->n;
  const VbModel M = *Mp;
  const bool oob = oob_i != 0;
  const int wt = (threadIdx.x >> 5) % T, team = (threadIdx.x >> 5) / T;  // warp of the team; team of the CTA
  int parity = *parity_io;

  ScalarWs& ws = *ws_shared;
  double *xc = ws.xc, *cand = ws.cand, *g = ws.g, *H = ws.H, *scale = ws.scale, *xbest = ws.xbest;
  double &x_cost = ws.x_cost, &xnorm = ws.xnorm, &best_cost = ws.best_cost, &model_change = ws.model_change, &gmax = ws.gmax;
  for (int i = 0; i < 3; ++i) {
    xc[i] = cand[i] = xbest[i] = x[i];
    g[i] = 0.0;
    scale[i] = 1.0;
  }
  for (int i = 0; i < 6; ++i) H[i] = 0.0;
  x_cost = xnorm = best_cost = model_change = gmax = 0.0;
  Dogleg& S = ws.S;
  S.radius = 1e4;
  S.mu = 1e-8;
  S.reuse = 0;
  S.sub1d = 0;
  S.alpha = 0.0;
  S.step_norm = 0.0;
  S.gnn = 0.0;
  S.tb_l1 = S.tb_l2 = S.tb_gn = 0.0;
  S.inv_diag[0] = S.inv_diag[1] = S.inv_diag[2] = 1.0;
  int iter = 0, invalid = 0, term = TERM_MAX_ITER, n_evals = 0;
  int bi = 0, nb = 0;  // cursor into / entries of the candidate batch (entry k: radius at build time x 2^-k)
  bool succ = true, first = true;
  // rotation of the candidate: sin, cos and AngleAxisRotatePoint's unit axis theta * (1 / |theta|), which is 1 or
  // 1 - 2^-53 (small-angle branch: 1)
  double sn, cs, w0 = 1.0;
  sincos_once(cand[0], &sn, &cs);
  {
    const double at = fabs(cand[0]);
    if (at * at > DBL_EPSILON) w0 = at * ddiv(1.0, at);
  }
#ifdef FIT_TIMING
  long long t_eval = 0, t_scalar = 0, t_cls = 0, t_jac = 0, t0 = clock64();
#endif
  for (;;) {
#ifdef FIT_TIMING
    { const long long t1 = clock64(); t_scalar += t1 - t0; t0 = t1; }
#endif
    // ---- cost at the candidate ----
    double a[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) a[q] = 0.0;
    int n1 = 0, n2 = 0;
    {
#ifdef FIT_TIMING
      const long long tc0 = clock64();
#endif
      // this warp's share of the frame's points: every T-th group of 32 * CLS_UNROLL points
      classify_pass<IdxT>(yz, wbits, list, list_cap, M, oob, cs, sn, cand[1], cand[2], lane, wt * (32 * CLS_UNROLL), n,
                          T * 32 * CLS_UNROLL, n1, n2);
#ifdef FIT_TIMING
      t_cls += clock64() - tc0;
#endif
      list_pass<IdxT, false>(yz, wbits, list, list_cap, M, oob, cs, sn, w0, cand[1], cand[2], lane, n1, n2, a);
    }
    double cc1[1] = {warp_butterfly(a[0])};
    team_sum<1, T>(cc1, slots, wt, lane, parity, team);
    const double cc = cc1[0];
    n_evals++;
#ifdef FIT_TIMING
    { const long long t1 = clock64(); t_eval += t1 - t0; t0 = t1; }
#endif
    bool accept = true;
    double quality = 1.0;
    if (!first) {
      const double d0 = xc[0] - cand[0], d1 = xc[1] - cand[1], d2 = xc[2] - cand[2];
      const double step_norm = dsqrt(d0 * d0 + d1 * d1 + d2 * d2);
      if (step_norm <= 1e-8 * (xnorm + 1e-8)) { term = TERM_PARAMETER; break; }
      const double cost_change = x_cost - cc;
      if (fabs(cost_change) <= 1e-6 * x_cost) { term = TERM_FUNCTION; break; }
      quality = ddiv(cost_change, model_change);
      accept = quality > 1e-3;
    }
    if (accept) {
      // ---- Jacobian at the accepted point: from the lists of the evaluation that has just been accepted ----
      for (int i = 0; i < 3; ++i) xc[i] = cand[i];
      xnorm = norm3(xc);
      x_cost = cc;
#pragma unroll
      for (int q = 0; q < 9; ++q) a[q] = 0.0;
      list_pass<IdxT, true>(yz, wbits, list, list_cap, M, oob, cs, sn, w0, xc[1], xc[2], lane, n1, n2, a);
#pragma unroll
      for (int q = 0; q < 9; ++q) a[q] = warp_butterfly(a[q]);
      team_sum<9, T>(a, slots, wt, lane, parity, team);
      g[0] = a[0];
      g[1] = a[1];
      g[2] = a[2];
#pragma unroll
      for (int q = 0; q < 6; ++q) H[q] = a[3 + q];
      gmax = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
      succ = true;
      nb = 0;  // new model: the parked candidates are stale
#ifdef FIT_TIMING
      { const long long t1 = clock64(); t_jac += t1 - t0; t0 = t1; }
#endif
      if (first) {
        scale[0] = ddiv(1.0, 1.0 + dsqrt(H[0]));
        scale[1] = ddiv(1.0, 1.0 + dsqrt(H[3]));
        scale[2] = ddiv(1.0, 1.0 + dsqrt(H[5]));
        best_cost = x_cost;
      } else {
        {
          double r = S.radius;
          if (quality < 0.25) r *= 0.5;
          if (quality > 0.75) r = fmax(r, 3.0 * S.step_norm);
          ws_update(S.radius, r);
          ws_update(S.mu, fmax(1e-8, ddiv(2.0 * S.mu, 10.0)));
        }
        S.reuse = 0;
        if (x_cost < best_cost) {
          best_cost = x_cost;
          for (int i = 0; i < 3; ++i) xbest[i] = xc[i];
        }
      }
    } else {
      succ = false;
      const double gnn = S.gnn;
      double r = S.radius;
      const bool was_gn = gnn <= r;
      r *= 0.5;
      S.reuse = 1;
      ++bi;
      if (was_gn) {
        // the identical Gauss-Newton candidate would be evaluated and rejected again
        while (iter < max_iter && gnn <= r && r > 1e-32) {
          iter++;
          r *= 0.5;
          ++bi;
        }
      }
      ws_update(S.radius, r);
    }
    first = false;
    // ---- next trust-region step (loops over invalid steps) ----
    bool stop = false;
    for (;;) {
      if (iter >= max_iter) { term = TERM_MAX_ITER; stop = true; break; }
      if (succ && gmax <= 1e-10) { term = TERM_GRADIENT; stop = true; break; }
      if (S.radius <= 1e-32) { term = TERM_MIN_RADIUS; stop = true; break; }
      iter++;
      if (bi >= nb) {
        // (re)build the batch from the current radius: lane k takes radius 2^-k
        double* Hs = ws.Hs;
        double* gs = ws.gs;
        bool ok = true;
        if (!S.reuse) {
          Hs[0] = H[0] * scale[0] * scale[0];
          Hs[1] = H[1] * scale[0] * scale[1];
          Hs[2] = H[2] * scale[0] * scale[2];
          Hs[3] = H[3] * scale[1] * scale[1];
          Hs[4] = H[4] * scale[1] * scale[2];
          Hs[5] = H[5] * scale[2] * scale[2];
          for (int i = 0; i < 3; ++i) gs[i] = g[i] * scale[i];
          S.reuse = 1;
          ok = dogleg_build_model(S, Hs, gs);
        }
        __syncwarp();
        const int kk = lane < kFitBatch ? lane : kFitBatch - 1;
        const double rk = S.radius * __hiloint2double((1023 - kk) << 20, 0);  // radius 2^-k, exact
        double st[3] = {0.0, 0.0, 0.0};
        double stepn = 0.0, mc = -1.0;
        if (ok) {
          stepn = dogleg_step_for_radius(S, rk, st);
          mc = -((st[0] * gs[0] + st[1] * gs[1] + st[2] * gs[2]) + 0.5 * quad3(Hs, st, st));
        }
        const double c0 = xc[0] + st[0] * scale[0], c1 = xc[1] + st[1] * scale[1], c2 = xc[2] + st[2] * scale[2];
        double bsn, bcs, bw0 = 1.0;
        sincos_once(c0, &bsn, &bcs);
        {
          const double at = fabs(c0);
          if (at * at > DBL_EPSILON) bw0 = at * ddiv(1.0, at);
        }
        if (lane < kFitBatch) {
          double* e = ws.batch[lane];
          e[0] = c0;
          e[1] = c1;
          e[2] = c2;
          e[3] = mc;
          e[4] = stepn;
          e[5] = bsn;
          e[6] = bcs;
          e[7] = bw0;
        }
        __syncwarp();
        nb = kFitBatch;
        bi = 0;
      }
      const double* e = ws.batch[bi];
      const double mc = e[3];
      if (mc > 0.0) {
        cand[0] = e[0];
        cand[1] = e[1];
        cand[2] = e[2];
        model_change = mc;
        S.step_norm = e[4];
        sn = e[5];
        cs = e[6];
        w0 = e[7];
        break;
      }
      if (++invalid >= 5) { term = TERM_INVALID; stop = true; break; }
      ws_update(S.mu, S.mu * 10.0);
      S.reuse = 0;
      nb = 0;
      succ = false;
    }
    if (stop) break;
    invalid = 0;
  }
  for (int i = 0; i < 3; ++i) x[i] = xbest[i];
  *final_cost = best_cost;
  *iters_out = iter;
  *n_evals_io += n_evals;
  *parity_io = parity;
#ifdef FIT_TIMING
  t_scalar += clock64() - t0;
  if (lane == 0) {
    if ((threadIdx.x >> 5) == 0 && (FIT_PAIR || (blockIdx.x & 1) == 0)) {
      atomicAdd(&Fp->timing[0], (float)t_cls * 1e-3f);
      atomicAdd(&Fp->timing[1], (float)(t_eval - t_cls) * 1e-3f);
      atomicAdd(&Fp->timing[2], (float)t_jac * 1e-3f);
      atomicAdd(&Fp->timing[3], (float)t_scalar * 1e-3f);
    }
  }
#endif
  return term;
}


__device__ __forceinline__ void rigid_inverse(const float T[16], float O[16]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) O[4 * r + c] = T[4 * c + r];
  for (int r = 0; r < 3; ++r) {
    float v = O[4 * r + 0] * T[3];
    v = v + O[4 * r + 1] * T[7];
    v = v + O[4 * r + 2] * T[11];
    O[4 * r + 3] = -1.0f * v;
  }
  O[12] = 0.f; O[13] = 0.f; O[14] = 0.f; O[15] = 1.f;
}

// Work-balancing permutation for K5: frames bucketed by their number of optimiser points (the work of a solve is
// ~ points x evaluations), largest first, and dealt to thread blocks in snake order over rounds of `num_sms`
// blocks, so that the frames resident on one SM add up to about the same work.  One CTA; any number of frames.
__global__ void __launch_bounds__(1024) fit_order_kernel(const Workspace w, int num_sms) {
  constexpr int NB = 128;
  __shared__ int s_hist[NB], s_start[NB];
  const int tid = threadIdx.x, F = w.n_frames;
  for (int i = tid; i < NB; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int base = w.frame_base;
  auto bucket = [&](int fl) {
    const int f = base + fl;
    const int n = (w.diag[f].status == ILCC2_OK) ? w.opt_count[f] : 0;
    return NB - 1 - min(n >> 5, NB - 1);  // descending size
  };
  for (int f = tid; f < F; f += blockDim.x) {
    atomicAdd(&s_hist[bucket(f)], 1);
    w.fit_done[base + f] = 0;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int b = 0; b < NB; ++b) {
      s_start[b] = acc;
      acc += s_hist[b];
    }
  }
  __syncthreads();
  const int full = (F / num_sms) * num_sms;
  for (int f = tid; f < F; f += blockDim.x) {
    const int pos = atomicAdd(&s_start[bucket(f)], 1);
    int blk = pos;
    if (pos < full) {
      const int round = pos / num_sms, j = pos % num_sms;
      blk = round * num_sms + ((round & 1) ? num_sms - 1 - j : j);
    }
    w.fit_order[base + blk] = base + f;
  }
}

// SMEM = true : frames with at most kFitBitsCap optimiser points (16-bit candidate lists + colour bits in shared memory;
//               frames above kFitTeamCap points are evaluated in chunks of kFitTeamCap)
// SMEM = false: larger frames; same code on global-memory scratch.  Both are launched; each skips the other's frames.
#ifndef FIT_WARPS_PER_SM
#define FIT_WARPS_PER_SM 16
#endif
// The all-gather of SURVEY.md 8(e) fused into the producer: besides the caller's own buffers, K5 stores a frame's
// corners and status into every gather target -- peer-mapped buffers of the ranks (plain stores over NVLink), or ONE
// multicast (NVLS) mapping that NVSwitch fans out to all ranks.  A multicast address may only be touched by multimem.*
// instructions (PTX ISA: any other access to it is undefined), so that mode stores with multimem.st.
__device__ __forceinline__ void gather_store_corner(const Workspace& w, int f, int ncorners, int i, float v) {
  const size_t at = (size_t)(w.gather_slot_base + f) * ncorners * 3 + i;
  if (w.gather_multicast) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(w.gather_corners[0] + at), "f"(v) : "memory");
  } else {
    for (int t = 0; t < w.n_targets; ++t) w.gather_corners[t][at] = v;
  }
}
__device__ __forceinline__ void gather_store_status(const Workspace& w, int f, int32_t st) {
  if (w.gather_multicast) {
    asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(w.gather_status[0] + w.gather_slot_base + f), "r"(st) : "memory");
  } else {
    for (int t = 0; t < w.n_targets; ++t) w.gather_status[t][w.gather_slot_base + f] = st;
  }
}

// FIT_PAIR = 1: one CTA per FRAME holding two teams of T warps, one per colour hypothesis (auto mode), that run their
// solves independently (a named barrier each) on the same points: the frame's points are then resident in the SM's L1
// once instead of once per hypothesis on two different SMs (L1 hit rate of the point loads 69 % -> see profiles/), each
// CTA carries one cheap (right colour phase) and one expensive (wrong phase: half of the points mismatch) solve, and the
// CTA-wide barrier at the end replaces the global-memory handshake between the two CTAs of a frame.
// FIT_PAIR = 0: one CTA per (frame, hypothesis) as in round 1.
template <bool SMEM, int T>
__global__ void __launch_bounds__(32 * T * (FIT_PAIR ? 2 : 1), (SMEM ? FIT_WARPS_PER_SM : 8) / (T * (FIT_PAIR ? 2 : 1)))
    board_fit_kernel(const Workspace w, const DevParams p) {
  using IdxT = typename std::conditional<SMEM, uint16_t, uint32_t>::type;
  constexpr int kTeams = FIT_PAIR ? 2 : 1;
  constexpr int kWarps = T * kTeams, kThreads = 32 * kWarps;
  extern __shared__ unsigned char fit_sm[];
  __shared__ double s_x[2][3];
  __shared__ double s_cost[2];
  __shared__ float s_ext[kWarps][2];
  __shared__ double s_red[kTeams * 2 * T * 9];  // team_sum slots: [team][parity][warp][9]
  __shared__ int s_last;
  __shared__ ScalarWs s_ws[kWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int team_of_cta = kTeams == 2 ? warp / T : 0;
  const int fblock = kTeams == 2 ? blockIdx.x : blockIdx.x >> 1;
  const int hyp_cta = kTeams == 2 ? team_of_cta : (blockIdx.x & 1);
  const int f = w.fit_order[w.frame_base + fblock];  // big frames first (fit_order_kernel)
  ilcc2_diag* d = w.diag + f;
  float* corners = w.corners + (size_t)f * p.ncorners * 3;
  const int st = d->status;
  const int n = w.opt_count[f];
  if (st == ILCC2_OK && (n <= w.fit_bits_cap) != SMEM) return;  // the other instantiation owns this frame
  if (st != ILCC2_OK) {
    if (SMEM && (kTeams == 2 || hyp_cta == 0)) {
      for (int i = tid; i < p.ncorners * 3; i += kThreads) {
        corners[i] = NAN;
        gather_store_corner(w, f, p.ncorners, i, NAN);
      }
      if (tid == 0) {
        w.status[f] = st;
        gather_store_status(w, f, st);
      }
    }
    return;
  }
  const unsigned flags = w.flags ? (unsigned)w.flags[f] : w.default_flags;
  const bool autom = (flags & ILCC2_FLAG_AUTO) != 0;
  if (kTeams == 1 && !autom && hyp_cta == 1) return;  // manual mode: one CTA runs the requested hypotheses in sequence
  const float2* gyz = w.opt_yz + (size_t)f * w.cap;
  const uint8_t* gwh = w.opt_white + (size_t)f * w.cap;
  FitPts<IdxT> F;
  F.n = n;
  uint32_t* wbits;
  if (SMEM) {
    // shared memory holds the two candidate lists (16-bit indices) of every warp and the colour bits; the points
    // themselves are read through L1 (a frame's optimiser points are ~15 KB), which keeps the shared-memory footprint
    // small so that registers, not shared memory, bound the residency.
    F.bits_bytes = (int)sizeof(uint32_t) * (w.fit_bits_cap / 32);
    IdxT* lists = reinterpret_cast<IdxT*>(fit_sm + F.bits_bytes);
    F.team_cap = w.fit_team_cap;
    F.list_cap = fit_list_cap(w.fit_team_cap, T);
    F.list = lists + (size_t)warp * F.list_cap;
    F.yz = gyz;
    wbits = reinterpret_cast<uint32_t*>(fit_sm);
  } else {
    if (tid == 0) s_last = atomicAdd(w.fit_pool_next, 1);  // next free slot of the overflow pool
    __syncthreads();
    if (s_last >= w.cl_pool_slots) {  // more frames beyond the shared-memory layout than the handle holds tables for
      for (int i = tid; i < p.ncorners * 3; i += kThreads) {
        corners[i] = NAN;
        gather_store_corner(w, f, p.ncorners, i, NAN);
      }
      if (tid == 0) {
        d->status = ILCC2_ROI_CAPACITY;
        w.status[f] = ILCC2_ROI_CAPACITY;
        gather_store_status(w, f, ILCC2_ROI_CAPACITY);
      }
      return;
    }
    uint32_t* g = w.fit_scratch + (size_t)s_last * w.fit_stride;
    wbits = g;
    F.yz = gyz;
    F.bits_bytes = 0;
    F.list = reinterpret_cast<IdxT*>(g + (w.cap + 31) / 32 + (size_t)(kTeams == 2 ? warp : hyp_cta * T + warp) * w.cap);
    F.list_cap = w.cap;
    F.team_cap = w.cap;
  }
  float ymax = 0.f, zmax = 0.f;
  for (int base = warp * 32; base < n; base += kThreads) {  // stage: colour flags -> one bit per point; extent
    const int k = base + lane;
    const bool v = k < n;
    const float2 q = v ? gyz[k] : make_float2(0.f, 0.f);
    ymax = fmaxf(ymax, fabsf(q.x));
    zmax = fmaxf(zmax, fabsf(q.y));
    const unsigned m = __ballot_sync(FULL, v && gwh[k] != 0);
    if (lane == 0) wbits[base >> 5] = m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ymax = fmaxf(ymax, __shfl_xor_sync(FULL, ymax, o));
    zmax = fmaxf(zmax, __shfl_xor_sync(FULL, zmax, o));
  }
  if (lane == 0) {
    s_ext[warp][0] = ymax;
    s_ext[warp][1] = zmax;
  }
  F.wbits = wbits;
  F.timing = d->plane;
#ifdef FIT_TIMING
  if (tid < 4 && hyp_cta == 0) F.timing[tid] = 0.f;
#endif
  if (tid < 2) s_cost[tid] = 0.0;
  __syncthreads();

  VbModel M;
  M.Wd = (double)p.W;
  M.Hd = (double)p.H;
  M.g = p.g;
  M.inv_g = p.inv_g;
  M.wg2 = p.wg2;
  M.hg2 = p.hg2;
  M.ha = p.huber_a;
  M.hb = p.huber_b;
  M.two_ha = 2.0 * p.huber_a;
  {
    float ym = 0.f, zm = 0.f;
    for (int i = 0; i < kWarps; ++i) {
      ym = fmaxf(ym, s_ext[i][0]);
      zm = fmaxf(zm, s_ext[i][1]);
    }
    M.ymax_zmax = ym + zm;
  }
  const int wt = warp % T;
  if (autom || team_of_cta == 0) {  // manual mode: team 0 runs the requested hypotheses in sequence, team 1 waits
    const int h_first = autom ? hyp_cta : 0;
    const int h_last = autom ? hyp_cta : ((flags & ILCC2_FLAG_TOPLEFT_WHITE) ? 1 : 0);
    double x[3] = {0.0, 0.0, 0.0};
    int parity = 0;
    for (int h = h_first; h <= h_last; ++h) {  // manual mode: 'd' warm-starts hypothesis 1 from hypothesis 0
      int evals = 0;
      double fc = 0.0;
      M.tlw = h;
#pragma unroll 1
      for (int phase = 0; phase < 2; ++phase) {  // useOutofBoard first, then without (LidarCornersEst.cpp:402-409); one call site
        int it = 0;
        const int term = solve_tr<IdxT, T>(&F, &M, 1 - phase, p.max_solver_iter, lane, x, &fc, &it, &evals,
                                           s_red + team_of_cta * (2 * T * 9), &parity, &s_ws[warp]);
        if (lane == 0 && wt == 0) {
          d->iters[h][phase] = it;
          d->term[h][phase] = term;
          if (phase == 0) d->cost_a[h] = fc;
        }
      }
      if (lane == 0 && wt == 0) {
        d->n_evals[h] = evals;
        d->cost[h] = fc;
        for (int i = 0; i < 3; ++i) {
          d->theta_t[h][i] = x[i];
          s_x[h][i] = x[i];
        }
        s_cost[h] = fc;
      }
    }
  }
  __syncthreads();
  if (kTeams == 1 && autom) {
    // the CTA of the other hypothesis publishes through diag; whichever CTA arrives second finishes the frame
    if (tid == 0) {
      __threadfence();
      s_last = atomicAdd(&w.fit_done[f], 1);
    }
    __syncthreads();
    if (s_last == 0) return;
    __threadfence();
    if (tid < 2) {
      const int oh = 1 - hyp_cta;
      volatile const double* vd = reinterpret_cast<volatile const double*>(d->theta_t[oh]);
      if (tid == 0) s_cost[oh] = *reinterpret_cast<volatile const double*>(&d->cost[oh]);
      for (int i = tid; i < 3; i += 2) s_x[oh][i] = vd[i];
    }
    __syncthreads();
  }
  const int chosen = autom ? ((s_cost[1] < s_cost[0]) ? 1 : 0) : ((flags & ILCC2_FLAG_TOPLEFT_WHITE) ? 1 : 0);
  // getPCDcorners: lattice -> transOptim^-1 -> transPCA^-1, fp32
  int W = p.W, H = p.H;
  if (flags & ILCC2_FLAG_INVERSE) {
    W = p.H;
    H = p.W;
  }
  const float thf = (float)s_x[chosen][0];
  const float E = (float)cos((double)thf), Fs = (float)sin((double)thf);
  const float Topt[16] = {1.f, 0.f, 0.f, 0.f, 0.f, E, -Fs, (float)s_x[chosen][1], 0.f, Fs, E, (float)s_x[chosen][2],
                          0.f, 0.f, 0.f, 1.f};
  float pca[16], Ti[16], Pi[16];
  for (int i = 0; i < 16; ++i) pca[i] = d->pca[i];
  rigid_inverse(Topt, Ti);
  rigid_inverse(pca, Pi);
  for (int k = tid; k < p.ncorners; k += kThreads) {
    const int i = 1 + k / (H - 1), j = 1 + k % (H - 1);
    const float py = (float)(((double)i - (double)W / 2.0) * p.g);
    const float pz = (float)(((double)j - (double)H / 2.0) * p.g);
    float q[3], r[3];
    xform_pt(Ti, 0.0f, py, pz, q);
    xform_pt(Pi, q[0], q[1], q[2], r);
    corners[3 * k + 0] = r[0];
    corners[3 * k + 1] = r[1];
    corners[3 * k + 2] = r[2];
    for (int c = 0; c < 3; ++c) gather_store_corner(w, f, p.ncorners, 3 * k + c, r[c]);
  }
  if (tid == 0) {
    d->hyp_chosen = chosen;
    w.status[f] = ILCC2_OK;
    gather_store_status(w, f, ILCC2_OK);
  }
}

size_t fit_smem_bytes(const Workspace& w, int team) {
  return sizeof(uint32_t) * (w.fit_bits_cap / 32) +
         sizeof(uint16_t) * (size_t)team * (FIT_PAIR ? 2 : 1) * fit_list_cap(w.fit_team_cap, team);
}

// Team size for a launch of `frames` frames: the largest T in {2, 4, 8} for which all 2 * frames CTAs are resident at
// once (128 registers per thread -> 16 warps per SM).  Large batches are throughput-bound (T = 2 keeps the redundant
// scalar work low); small ones are bound by the latency of a single solve, which more warps per solve shorten.
int fit_team_size(int frames, int num_sms) {
#ifdef FIT_FORCE_TEAM
  return FIT_FORCE_TEAM;
#endif
  const long long warps = (long long)FIT_WARPS_PER_SM * num_sms;
  if (2ll * frames * 8 <= warps) return 8;
  if (2ll * frames * 4 <= warps) return 4;
  if (2ll * frames * 4 <= 5 * warps) return 2;  // up to 2.5 waves of two-warp teams (measured: 1024 frames 3.0 vs 3.7 ms)
  return 1;                                     // throughput regime: no redundant scalar work (2048 frames 5.2 vs 6.0 ms)
}

template <int T>
static void launch_fit_smem(const Workspace& w, const DevParams& p, cudaStream_t s) {
  board_fit_kernel<true, T><<<(FIT_PAIR ? 1 : 2) * w.n_frames, 32 * T * (FIT_PAIR ? 2 : 1), fit_smem_bytes(w, T), s>>>(w, p);
}

void launch_board_fit(const Workspace& w, const DevParams& p, cudaStream_t s) {
  fit_order_kernel<<<1, 1024, 0, s>>>(w, w.num_sms);
  switch (fit_team_size(w.load_frames, w.num_sms)) {
    case 1: launch_fit_smem<1>(w, p, s); break;
    case 2: launch_fit_smem<2>(w, p, s); break;
    case 4: launch_fit_smem<4>(w, p, s); break;
    default: launch_fit_smem<8>(w, p, s); break;
  }
  if (w.cap > w.fit_bits_cap)
    board_fit_kernel<false, kFitGlobalTeam><<<(FIT_PAIR ? 1 : 2) * w.n_frames, 32 * kFitGlobalTeam * (FIT_PAIR ? 2 : 1), 0, s>>>(w, p);
}

cudaError_t configure_kernels() {
  cudaError_t e;
  e = cudaFuncSetAttribute(cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cluster_smem_bytes());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(board_fit_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(uint32_t) * (kFitBitsCapBig / 32) + sizeof(uint16_t) * kFitTeamCapBig * (FIT_PAIR ? 2 : 1)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(board_fit_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(uint32_t) * (kFitBitsCapBig / 32) + sizeof(uint16_t) * kFitTeamCapBig * (FIT_PAIR ? 2 : 1)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(board_fit_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(uint32_t) * (kFitBitsCapBig / 32) + sizeof(uint16_t) * kFitTeamCapBig * (FIT_PAIR ? 2 : 1)));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(board_fit_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(uint32_t) * (kFitBitsCapBig / 32) + sizeof(uint16_t) * kFitTeamCapBig * (FIT_PAIR ? 2 : 1)));
  return e;
}

}  // namespace ilcc2
