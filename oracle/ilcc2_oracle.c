/*
 * ilcc2_oracle.c -- CPU ORACLE (test infrastructure, never shipped; see ilcc2_oracle.h).
 *
 * Plain-C restatement of the reference's LiDAR chessboard-corner path.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC   (oracle/Makefile)
 * Floating-point contraction is off on purpose: the fp32 stages are written so that every
 * comparison that decides set membership (ROI box, 0.12 m linkage, 0.03 m plane distance) is an
 * explicit sequence of IEEE single-precision operations that the CUDA product repeats in the same
 * order, so point sets can be compared index-for-index.
 *
 * Double-precision sums that feed decisions (centroid / covariance / normal equations) use one
 * canonical order (strided partial sums + xor butterfly, see racc below) so that a parallel
 * implementation can reproduce them; where the reference accumulates in float (PCL) this is a
 * documented deviation (more accurate, <= 3e-6 m on the centroid).
 */
#include "ilcc2_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* test-only restatement switches (ilcc2o_set_variant): each bit replaces ONE restatement choice */
/* by its alternative so that tests/tools can measure how far the end result moves (DESIGN.md    */
/* "deviation table").  0 = the choices documented in the header.  Not thread-local: set it      */
/* before a batch, never during one.                                                             */
/* ------------------------------------------------------------------------------------------ */
static uint32_t g_variant = 0;
static int g_extra_d_pairs = 0;   /* extra pairs of 'd' presses (each press toggles topleftWhite and re-optimises warm) */
static int g_tr_perturb_ulps = 0; /* boundary-problem solution moved by this many ulps (sensitivity probe) */
void ilcc2o_set_variant(uint32_t mask, int extra_d_pairs, int tr_perturb_ulps) {
  g_variant = mask;
  g_extra_d_pairs = extra_d_pairs < 0 ? 0 : extra_d_pairs;
  g_tr_perturb_ulps = tr_perturb_ulps;
}
uint32_t ilcc2o_get_variant(void) { return g_variant; }

/* ------------------------------------------------------------------------------------------ */
/* canonical reduction                                                                        */
/* ------------------------------------------------------------------------------------------ */
/* Element k is added to partial sum k mod w (in increasing k); the w partial sums are combined by an xor
 * butterfly from stride w/2 down to 1.  w = 128 for the centroid / covariance / intensity sums (a 128-thread
 * CTA reproduces it with one partial per thread), w = 32 for the optimiser's normal equations (one warp). */
#define RW_MAX 128
#define RW_BLOCK 128
#define RW_WARP 32

typedef struct { double p[RW_MAX]; int w; } racc;

static void racc_zero_w(racc* a, int w) { memset(a->p, 0, sizeof(a->p)); a->w = w; }
static void racc_zero(racc* a) { racc_zero_w(a, RW_BLOCK); }
static inline void racc_add(racc* a, int k, double v) { a->p[k & (a->w - 1)] += v; }
static double racc_total(const racc* a) {
  double cur[RW_MAX], nxt[RW_MAX];
  const int w = a->w;
  memcpy(cur, a->p, sizeof(cur));
  for (int s = w / 2; s >= 1; s >>= 1) {
    for (int v = 0; v < w; ++v) nxt[v] = cur[v] + cur[v ^ s];
    memcpy(cur, nxt, sizeof(cur));
  }
  return cur[0];
}

/* ------------------------------------------------------------------------------------------ */
/* parameters                                                                                 */
/* ------------------------------------------------------------------------------------------ */
void ilcc2o_default_params(ilcc2o_params* p) {
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

int ilcc2o_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static inline float ldf(const unsigned char* base, int off) {
  float v;
  memcpy(&v, base + off, sizeof(float));
  return v;
}

/* ------------------------------------------------------------------------------------------ */
/* a3: setROI  (LidarCornersEst.cpp:48-70): three in-place pcl::PassThrough, z then x then y.  */
/* [UPSTREAM PCL] limits are stored as float; a point is dropped if any of x,y,z is not       */
/* finite, or field < min or field > max (inclusive limits); input order is preserved.        */
/* The reference computes point.z-2.0 in double (float - double literal) then narrows.        */
/* ------------------------------------------------------------------------------------------ */
static int roi_filter(const unsigned char* pts, int n, int step, int ox, int oy, int oz,
                      const float click[3], const ilcc2o_params* P, int* roi_idx) {
  const float zmin = (float)((double)click[2] - (double)P->roi_half_z);
  const float zmax = (float)((double)click[2] + (double)P->roi_half_z);
  const float xmin = (float)((double)click[0] - (double)P->roi_half_x);
  const float xmax = (float)((double)click[0] + (double)P->roi_half_x);
  const float ymin = (float)((double)click[1] - (double)P->roi_half_y);
  const float ymax = (float)((double)click[1] + (double)P->roi_half_y);
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const unsigned char* r = pts + (size_t)i * step;
    const float x = ldf(r, ox), y = ldf(r, oy), z = ldf(r, oz);
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;
    if (z < zmin || z > zmax) continue;
    if (x < xmin || x > xmax) continue;
    if (y < ymin || y > ymax) continue;
    roi_idx[m++] = i;
  }
  return m;
}

/* squared distance the way FLANN's L2_Simple<float> accumulates it: ((dx*dx)+dy*dy)+dz*dz */
static inline float sqdist3(const float* a, const float* b) {
  const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  float r = dx * dx;
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}

/* ------------------------------------------------------------------------------------------ */
/* a4: EuclideanCluster (LidarCornersEst.cpp:124-153)                                         */
/* [UPSTREAM PCL] extractEuclideanClusters: BFS from every unprocessed point in index order,   */
/* neighbours = radiusSearch(tolerance) i.e. squared distance < (float)(tol*tol); clusters with*/
/* min<=size<=max kept, indices sorted ascending, clusters sorted by size descending (tie order*/
/* is unspecified in PCL's std::sort; here: discovery order, i.e. lowest first index first).   */
/* Neighbour search uses a uniform grid (cell = tolerance) in place of the kd-tree: same set.  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int* label;      /* per ROI point: cluster id in discovery order (all clusters, also small ones) */
  int* sizes;      /* per discovered cluster */
  int* first;      /* lowest point index per discovered cluster */
  int n_found;
} cc_result;

static void cluster_cc(const float* xyz /* M x 3 */, int M, float tol, cc_result* out) {
  const float r2 = (float)((double)tol * (double)tol);
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < M; ++i)
    for (int k = 0; k < 3; ++k) {
      if (xyz[3 * i + k] < mn[k]) mn[k] = xyz[3 * i + k];
      if (xyz[3 * i + k] > mx[k]) mx[k] = xyz[3 * i + k];
    }
  const double cs = (double)tol * 1.0001; /* a hair larger than tol: 27-cell stencil is a superset */
  int dim[3];
  for (int k = 0; k < 3; ++k) {
    dim[k] = (int)floor(((double)mx[k] - (double)mn[k]) / cs) + 1;
    if (dim[k] < 1) dim[k] = 1;
  }
  const size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
  int* cell_of = (int*)malloc(sizeof(int) * (size_t)M);
  int* start = (int*)calloc(ncell + 1, sizeof(int));
  int* order = (int*)malloc(sizeof(int) * (size_t)M);
  for (int i = 0; i < M; ++i) {
    int c[3];
    for (int k = 0; k < 3; ++k) {
      c[k] = (int)floor(((double)xyz[3 * i + k] - (double)mn[k]) / cs);
      if (c[k] < 0) c[k] = 0;
      if (c[k] >= dim[k]) c[k] = dim[k] - 1;
    }
    cell_of[i] = (c[2] * dim[1] + c[1]) * dim[0] + c[0];
    start[cell_of[i] + 1]++;
  }
  for (size_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
  int* fill = (int*)malloc(sizeof(int) * ncell);
  memcpy(fill, start, sizeof(int) * ncell);
  for (int i = 0; i < M; ++i) order[fill[cell_of[i]]++] = i;
  free(fill);

  int* queue = (int*)malloc(sizeof(int) * (size_t)M);
  out->label = (int*)malloc(sizeof(int) * (size_t)M);
  out->sizes = (int*)malloc(sizeof(int) * (size_t)(M > 0 ? M : 1));
  out->first = (int*)malloc(sizeof(int) * (size_t)(M > 0 ? M : 1));
  out->n_found = 0;
  for (int i = 0; i < M; ++i) out->label[i] = -1;

  for (int i = 0; i < M; ++i) {
    if (out->label[i] >= 0) continue;
    const int id = out->n_found++;
    int qh = 0, qt = 0;
    queue[qt++] = i;
    out->label[i] = id;
    while (qh < qt) {
      const int p = queue[qh++];
      const int c = cell_of[p];
      const int cx = c % dim[0], cy = (c / dim[0]) % dim[1], cz = c / (dim[0] * dim[1]);
      for (int dz = -1; dz <= 1; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= dim[2]) continue;
        for (int dy = -1; dy <= 1; ++dy) {
          const int y = cy + dy;
          if (y < 0 || y >= dim[1]) continue;
          for (int dx = -1; dx <= 1; ++dx) {
            const int x = cx + dx;
            if (x < 0 || x >= dim[0]) continue;
            const int cc = (z * dim[1] + y) * dim[0] + x;
            for (int s = start[cc]; s < start[cc + 1]; ++s) {
              const int q = order[s];
              if (out->label[q] >= 0) continue;
              if (sqdist3(xyz + 3 * p, xyz + 3 * q) < r2) {
                out->label[q] = id;
                queue[qt++] = q;
              }
            }
          }
        }
      }
    }
    out->sizes[id] = qt;
    out->first[id] = i;
  }
  free(queue);
  free(order);
  free(start);
  free(cell_of);
}

static void cc_free(cc_result* r) {
  free(r->label);
  free(r->sizes);
  free(r->first);
}

/* ------------------------------------------------------------------------------------------ */
/* [UPSTREAM PCL/boost] SampleConsensusModel RNG: boost::mt19937 seeded 12345u (random=false),  */
/* drawn through boost::uniform_int<>(0, INT_MAX): range 2^31 on a 2^32 engine -> bucket size 2 */
/* -> value = mt() / 2.  A fresh model (fresh seed) is created by every getPlane() call          */
/* (LidarCornersEst.cpp:195), so the draw sequence is the same for every frame.                  */
/* ------------------------------------------------------------------------------------------ */
#define PCL_RND_TABLE 8192
static uint32_t g_rnd_table[PCL_RND_TABLE];
static int g_rnd_ready = 0;

static void rnd_table_init(void) {
  uint32_t mt[624];
  mt[0] = 12345u;
  for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
  int idx = 624;
  for (int k = 0; k < PCL_RND_TABLE; ++k) {
    if (idx >= 624) {
      for (int i = 0; i < 624; ++i) {
        const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    g_rnd_table[k] = y >> 1;
  }
}

static void rnd_ensure(void) {
  if (!g_rnd_ready) {
#ifdef _OPENMP
#pragma omp critical(ilcc2o_rnd)
#endif
    {
      if (!g_rnd_ready) {
        rnd_table_init();
        g_rnd_ready = 1;
      }
    }
  }
}

uint32_t ilcc2o_pcl_rnd(int k) {
  rnd_ensure();
  return g_rnd_table[k & (PCL_RND_TABLE - 1)];
}

/* plane distance as Eigen's SSE Vector4f dot reduces it: (a*x + c*z) + (b*y + d) */
static inline float plane_dist(const float m[4], const float* p) {
  const float t0 = m[0] * p[0] + m[2] * p[2];
  const float t1 = m[1] * p[1] + m[3];
  return fabsf(t0 + t1);
}

/* [UPSTREAM PCL] SampleConsensusModelPlane::computeModelCoefficients (float) */
static int plane_from_3(const float* p0, const float* p1, const float* p2, float m[4]) {
  const float a0 = p1[0] - p0[0], a1 = p1[1] - p0[1], a2 = p1[2] - p0[2];
  const float b0 = p2[0] - p0[0], b1 = p2[1] - p0[1], b2 = p2[2] - p0[2];
  const float r0 = a0 / b0, r1 = a1 / b1, r2 = a2 / b2;
  if (r0 == r1 && r2 == r1) return 0; /* collinear */
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
  float d = n0 * p0[0];
  d = d + n1 * p0[1];
  d = d + n2 * p0[2];
  m[0] = n0;
  m[1] = n1;
  m[2] = n2;
  m[3] = -1.0f * d;
  return 1;
}

/* [UPSTREAM PCL] isSampleGood: the three points are not collinear by the ratio test */
static int sample_good(const float* p0, const float* p1, const float* p2) {
  const float r0 = (p1[0] - p0[0]) / (p2[0] - p0[0]);
  const float r1 = (p1[1] - p0[1]) / (p2[1] - p0[1]);
  const float r2 = (p1[2] - p0[2]) / (p2[2] - p0[2]);
  return (r0 != r1) || (r2 != r1);
}

/* symmetric 3x3 eigen decomposition in double: cyclic Jacobi, fixed sweep count.
 * A: 00 01 02 11 12 22.  evals unsorted, V columns = eigenvectors (row-major V[r][c]). */
static void jacobi3d(const double A[6], double ev[3], double V[3][3]) {
  double a[3][3] = {{A[0], A[1], A[2]}, {A[1], A[3], A[4]}, {A[2], A[4], A[5]}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0);
        const double s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  ev[0] = a[0][0];
  ev[1] = a[1][1];
  ev[2] = a[2][2];
}

/* ------------------------------------------------------------------------------------------ */
/* a5: getPlane (LidarCornersEst.cpp:190-221) = pcl::SACSegmentation(PLANE, RANSAC, 0.03,      */
/* optimizeCoefficients).  [UPSTREAM PCL 1.7/1.8] RandomSampleConsensus::computeModel +         */
/* SACSegmentation::segment: adaptive k, best model by inlier count, LS refit on its inliers    */
/* (optimizeModelCoefficients), inliers re-selected against the refined plane.                  */
/* Deviation: the LS refit accumulates in double (canonical order, coordinates shifted by the    */
/* first inlier) and takes the smallest eigenvector from a double Jacobi solver; PCL accumulates */
/* raw moments in float and uses a closed-form float eigen33.                                    */
/* ------------------------------------------------------------------------------------------ */
/* [UPSTREAM PCL 1.7] SampleConsensusModelPlane::optimizeModelCoefficients as PCL writes it: raw moments
 * accumulated sequentially in fp32 (computeMeanAndCovarianceMatrix), covariance = E[xx] - E[x]E[x], smallest
 * eigenvector by the closed-form pcl::eigen33 (computeRoots + row cross products), d = -n.c. */
static void pcl_roots2_f32(float b, float c, float roots[3]) {
  roots[0] = 0.0f;
  float d = (float)(b * b - 4.0 * c);
  if (d < 0.0f) d = 0.0f;
  const float sd = sqrtf(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}
static void pcl_plane_refit_f32(const float* xyz, const int* inl, int ni, float out[4]) {
  float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < ni; ++k) {
    const float* p = xyz + 3 * inl[k];
    a[0] += p[0] * p[0];
    a[1] += p[0] * p[1];
    a[2] += p[0] * p[2];
    a[3] += p[1] * p[1];
    a[4] += p[1] * p[2];
    a[5] += p[2] * p[2];
    a[6] += p[0];
    a[7] += p[1];
    a[8] += p[2];
  }
  const float inv = 1.0f / (float)ni;
  for (int q = 0; q < 9; ++q) a[q] *= inv;
  float m[3][3];
  m[0][0] = a[0] - a[6] * a[6];
  m[0][1] = a[1] - a[6] * a[7];
  m[0][2] = a[2] - a[6] * a[8];
  m[1][1] = a[3] - a[7] * a[7];
  m[1][2] = a[4] - a[7] * a[8];
  m[2][2] = a[5] - a[8] * a[8];
  m[1][0] = m[0][1]; m[2][0] = m[0][2]; m[2][1] = m[1][2];
  float scale = 0.0f;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      if (fabsf(m[r][c]) > scale) scale = fabsf(m[r][c]);
  if (scale <= FLT_MIN) scale = 1.0f;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m[r][c] = m[r][c] / scale;
  float roots[3];
  {
    const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
                     m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
    const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] -
                     m[1][2] * m[1][2];
    const float c2 = m[0][0] + m[1][1] + m[2][2];
    if (fabsf(c0) < FLT_EPSILON) {
      pcl_roots2_f32(c2, c1, roots);
    } else {
      const float s_inv3 = (float)(1.0 / 3.0), s_sqrt3 = sqrtf(3.0f);
      const float c2_over_3 = c2 * s_inv3;
      float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
      if (a_over_3 > 0.0f) a_over_3 = 0.0f;
      const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
      float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
      if (q > 0.0f) q = 0.0f;
      const float rho = sqrtf(-a_over_3);
      const float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
      const float ct = cosf(theta), st = sinf(theta);
      roots[0] = c2_over_3 + 2.0f * rho * ct;
      roots[1] = c2_over_3 - rho * (ct + s_sqrt3 * st);
      roots[2] = c2_over_3 - rho * (ct - s_sqrt3 * st);
      float t;
      if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
      if (roots[1] >= roots[2]) {
        t = roots[1]; roots[1] = roots[2]; roots[2] = t;
        if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
      }
      if (roots[0] <= 0.0f) pcl_roots2_f32(c2, c1, roots);
    }
  }
  for (int r = 0; r < 3; ++r) m[r][r] -= roots[0];
  float v[3][3];
  const int pr[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  float len[3];
  for (int k = 0; k < 3; ++k) {
    const float* A = m[pr[k][0]];
    const float* B = m[pr[k][1]];
    v[k][0] = A[1] * B[2] - A[2] * B[1];
    v[k][1] = A[2] * B[0] - A[0] * B[2];
    v[k][2] = A[0] * B[1] - A[1] * B[0];
    len[k] = v[k][0] * v[k][0] + v[k][1] * v[k][1] + v[k][2] * v[k][2];
  }
  int best = 2;
  if (len[0] >= len[1] && len[0] >= len[2]) best = 0;
  else if (len[1] >= len[0] && len[1] >= len[2]) best = 1;
  const float nrm = sqrtf(len[best]);
  out[0] = v[best][0] / nrm;
  out[1] = v[best][1] / nrm;
  out[2] = v[best][2] / nrm;
  out[3] = -1.0f * (out[0] * a[6] + out[1] * a[7] + out[2] * a[8]);
}

static int plane_ransac(const float* xyz /* C x 3 */, int C, const ilcc2o_params* P, float plane[4],
                        int* inl /* out, C */, int* n_best_out, int* iters_out) {
  rnd_ensure();
  *n_best_out = 0;
  *iters_out = 0;
  if (C < 3) return 0;
  int* sh = (int*)malloc(sizeof(int) * (size_t)C);
  for (int i = 0; i < C; ++i) sh[i] = i;
  const double thr = P->ransac_threshold;
  const double log_prob = log(1.0 - P->ransac_probability);
  const double one_over = 1.0 / (double)C;
  int iterations = 0, skipped = 0, n_best = -INT32_MAX, draw = 0, have_model = 0;
  const int max_skip = P->ransac_max_iterations * 10;
  double k = 1.0;
  float best[4] = {0, 0, 0, 0};
  while ((double)iterations < k && skipped < max_skip) {
    int s[3] = {0, 0, 0}, good = 0;
    for (int chk = 0; chk < 1000 && !good; ++chk) {
      if (draw + 3 > PCL_RND_TABLE) break;
      for (int i = 0; i < 3; ++i) {
        const int j = i + (int)(g_rnd_table[draw++] % (uint32_t)(C - i));
        const int t = sh[i];
        sh[i] = sh[j];
        sh[j] = t;
      }
      s[0] = sh[0];
      s[1] = sh[1];
      s[2] = sh[2];
      good = sample_good(xyz + 3 * s[0], xyz + 3 * s[1], xyz + 3 * s[2]);
    }
    if (!good) break;
    float m[4];
    if (!plane_from_3(xyz + 3 * s[0], xyz + 3 * s[1], xyz + 3 * s[2], m)) {
      ++skipped;
      continue;
    }
    int cnt = 0;
    for (int i = 0; i < C; ++i)
      if ((double)plane_dist(m, xyz + 3 * i) < thr) ++cnt;
    if (cnt > n_best) {
      n_best = cnt;
      memcpy(best, m, sizeof(best));
      have_model = 1;
      const double w = (double)n_best * one_over;
      double p_no = 1.0 - pow(w, 3.0);
      if (p_no < DBL_EPSILON) p_no = DBL_EPSILON;
      if (p_no > 1.0 - DBL_EPSILON) p_no = 1.0 - DBL_EPSILON;
      k = log_prob / log(p_no);
    }
    ++iterations;
    if (iterations > P->ransac_max_iterations) break;
  }
  free(sh);
  *iters_out = iterations;
  if (!have_model) return 0;
  /* inliers of the best model */
  int ni = 0;
  for (int i = 0; i < C; ++i)
    if ((double)plane_dist(best, xyz + 3 * i) < thr) inl[ni++] = i;
  *n_best_out = ni;
  memcpy(plane, best, sizeof(best));
  if (ni < 4) return ni; /* PCL: needs at least 4 points to refit; keeps the sample model */
  /* LS refit */
  racc acc[9];
  for (int q = 0; q < 9; ++q) racc_zero(&acc[q]);
  const float* ref = xyz + 3 * inl[0];
  for (int kk = 0; kk < ni; ++kk) {
    const float* p = xyz + 3 * inl[kk];
    const double dx = (double)p[0] - (double)ref[0];
    const double dy = (double)p[1] - (double)ref[1];
    const double dz = (double)p[2] - (double)ref[2];
    racc_add(&acc[0], kk, dx);
    racc_add(&acc[1], kk, dy);
    racc_add(&acc[2], kk, dz);
    racc_add(&acc[3], kk, dx * dx);
    racc_add(&acc[4], kk, dx * dy);
    racc_add(&acc[5], kk, dx * dz);
    racc_add(&acc[6], kk, dy * dy);
    racc_add(&acc[7], kk, dy * dz);
    racc_add(&acc[8], kk, dz * dz);
  }
  double S[9];
  for (int q = 0; q < 9; ++q) S[q] = racc_total(&acc[q]);
  const double inv_n = 1.0 / (double)ni;
  const double mx = S[0] * inv_n, my = S[1] * inv_n, mz = S[2] * inv_n;
  double A[6];
  A[0] = S[3] * inv_n - mx * mx;
  A[1] = S[4] * inv_n - mx * my;
  A[2] = S[5] * inv_n - mx * mz;
  A[3] = S[6] * inv_n - my * my;
  A[4] = S[7] * inv_n - my * mz;
  A[5] = S[8] * inv_n - mz * mz;
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
  const double cx = (double)ref[0] + mx, cy = (double)ref[1] + my, cz = (double)ref[2] + mz;
  float refined[4];
  refined[0] = (float)nx;
  refined[1] = (float)ny;
  refined[2] = (float)nz;
  refined[3] = (float)(-(nx * cx + ny * cy + nz * cz));
  if (g_variant & ILCC2O_VAR_REFIT_PCL_F32) pcl_plane_refit_f32(xyz, inl, ni, refined);
  /* re-select inliers against the refined plane */
  int nr = 0;
  for (int i = 0; i < C; ++i)
    if ((double)plane_dist(refined, xyz + 3 * i) < thr) inl[nr++] = i;
  memcpy(plane, refined, sizeof(refined));
  return nr;
}

/* ------------------------------------------------------------------------------------------ */
/* [UPSTREAM Eigen 3.2.x] SelfAdjointEigenSolver<Matrix3f>::compute (iterative path, the one the */
/* reference's constructor call takes, LidarCornersEst.cpp:337): scale by max|coeff|, 3x3       */
/* tridiagonalisation special case, implicit symmetric QR steps with Wilkinson shift, sort      */
/* ascending.  Restated so that the eigenvector SIGNS follow the same construction: the corner   */
/* ORDER of the five shipped fixtures fixes the sign of the in-plane axes the author's build     */
/* produced, and this restatement reproduces all five (tests/test_oracle_golden.py) when the     */
/* deflation test uses Eigen 3.2's dummy_precision<float>() = 1e-5 (Eigen 3.3's 2*epsilon gives  */
/* one more QR sweep on bag 6 and flips two columns), so 3.2 semantics are what is restated.     */
/* ------------------------------------------------------------------------------------------ */
static inline float eig_hypotf(float x, float y) {
  float ax = fabsf(x), ay = fabsf(y), p, qp;
  if (ax > ay) {
    p = ax;
    qp = ay / p;
  } else {
    p = ay;
    qp = ax / p;
  }
  if (p == 0.0f) return 0.0f;
  return p * sqrtf(1.0f + qp * qp);
}

void ilcc2o_eigen3f(const float cov[9], float evals[3], float evecs[9]) {
  /* lower triangle of the input, row-major cov[r*3+c] */
  float m00 = cov[0], m10 = cov[3], m11 = cov[4], m20 = cov[6], m21 = cov[7], m22 = cov[8];
  float scale = fabsf(m00);
  if (fabsf(m10) > scale) scale = fabsf(m10);
  if (fabsf(m11) > scale) scale = fabsf(m11);
  if (fabsf(m20) > scale) scale = fabsf(m20);
  if (fabsf(m21) > scale) scale = fabsf(m21);
  if (fabsf(m22) > scale) scale = fabsf(m22);
  if (scale == 0.0f) scale = 1.0f;
  m00 /= scale; m10 /= scale; m11 /= scale; m20 /= scale; m21 /= scale; m22 /= scale;
  float diag[3], sub[2];
  float Q[3][3]; /* Q[r][c] */
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
  const int n = 3;
  int end = n - 1, start = 0, iter = 0;
  const float precision = 1e-5f; /* NumTraits<float>::dummy_precision() */
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if ((g_variant & ILCC2O_VAR_EIGEN33_DEFL)
              ? (fabsf(sub[i]) <= (fabsf(diag[i]) + fabsf(diag[i + 1])) * (2.0f * FLT_EPSILON) || fabsf(sub[i]) <= FLT_MIN)
              : (fabsf(sub[i]) <= (fabsf(diag[i]) + fabsf(diag[i + 1])) * precision))
        sub[i] = 0.0f;
    while (end > 0 && sub[end - 1] == 0.0f) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 30 * n) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0f) start--;
    /* tridiagonal_qr_step */
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
      /* JacobiRotation::makeGivens(x, z) */
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
      /* Q = Q * G : columns k, k+1 */
      for (int r = 0; r < 3; ++r) {
        const float xi = Q[r][k], yi = Q[r][k + 1];
        Q[r][k] = c * xi - s * yi;
        Q[r][k + 1] = s * xi + c * yi;
      }
    }
  }
  /* sort ascending (selection, swapping columns) */
  for (int i = 0; i < n - 1; ++i) {
    int k = 0;
    for (int j = 1; j < n - i; ++j)
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
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) evecs[c * 3 + r] = Q[r][c]; /* column-major */
}

/* ------------------------------------------------------------------------------------------ */
/* a6: transformbyPCA (LidarCornersEst.cpp:330-364).  Centroid / covariance accumulate in       */
/* double in canonical order and are narrowed to float (PCL accumulates in float, sequentially). */
/* ------------------------------------------------------------------------------------------ */
/* [UPSTREAM PCL 1.7 + Eigen 3.2] compute3DCentroid / computeCovarianceMatrixNormalized as PCL writes them:
 * sequential fp32 accumulation in input order; Eigen 3.2's `vec /= n` multiplies by (1/n). */
static void pcl_centroid_cov_f32(const float* xyz, int B, float c[3], float cov6[6]) {
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
  for (int i = 0; i < B; ++i) {
    s0 += xyz[3 * i + 0];
    s1 += xyz[3 * i + 1];
    s2 += xyz[3 * i + 2];
  }
  const float inv = 1.0f / (float)B;
  c[0] = s0 * inv;
  c[1] = s1 * inv;
  c[2] = s2 * inv;
  float xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
  for (int i = 0; i < B; ++i) {
    const float px = xyz[3 * i + 0] - c[0], py = xyz[3 * i + 1] - c[1], pz = xyz[3 * i + 2] - c[2];
    yy += py * py;
    yz += py * pz;
    zz += pz * pz;
    xx += px * px;
    xy += py * px;
    xz += pz * px;
  }
  cov6[0] = xx * inv; cov6[1] = xy * inv; cov6[2] = xz * inv; cov6[3] = yy * inv; cov6[4] = yz * inv; cov6[5] = zz * inv;
}

static void pca_frame(const float* xyz /* B x 3 */, int B, float T[16], float evals[3]) {
  racc a[3];
  for (int q = 0; q < 3; ++q) racc_zero(&a[q]);
  for (int i = 0; i < B; ++i) {
    racc_add(&a[0], i, (double)xyz[3 * i + 0]);
    racc_add(&a[1], i, (double)xyz[3 * i + 1]);
    racc_add(&a[2], i, (double)xyz[3 * i + 2]);
  }
  float c[3];
  for (int q = 0; q < 3; ++q) c[q] = (float)(racc_total(&a[q]) / (double)B);
  racc m[6];
  for (int q = 0; q < 6; ++q) racc_zero(&m[q]);
  for (int i = 0; i < B; ++i) {
    const double dx = (double)(xyz[3 * i + 0] - c[0]);
    const double dy = (double)(xyz[3 * i + 1] - c[1]);
    const double dz = (double)(xyz[3 * i + 2] - c[2]);
    racc_add(&m[0], i, dx * dx);
    racc_add(&m[1], i, dx * dy);
    racc_add(&m[2], i, dx * dz);
    racc_add(&m[3], i, dy * dy);
    racc_add(&m[4], i, dy * dz);
    racc_add(&m[5], i, dz * dz);
  }
  float cov[9];
  float cxx = (float)(racc_total(&m[0]) / (double)B), cxy = (float)(racc_total(&m[1]) / (double)B);
  float cxz = (float)(racc_total(&m[2]) / (double)B), cyy = (float)(racc_total(&m[3]) / (double)B);
  float cyz = (float)(racc_total(&m[4]) / (double)B), czz = (float)(racc_total(&m[5]) / (double)B);
  if (g_variant & ILCC2O_VAR_PCA_PCL_F32) {
    float c6[6];
    pcl_centroid_cov_f32(xyz, B, c, c6);
    cxx = c6[0]; cxy = c6[1]; cxz = c6[2]; cyy = c6[3]; cyz = c6[4]; czz = c6[5];
  }
  cov[0] = cxx; cov[1] = cxy; cov[2] = cxz;
  cov[3] = cxy; cov[4] = cyy; cov[5] = cyz;
  cov[6] = cxz; cov[7] = cyz; cov[8] = czz;
  float V[9];
  ilcc2o_eigen3f(cov, evals, V);
  /* col2 = col0 x col1 (:343) */
  const float* v0 = V;
  const float* v1 = V + 3;
  float v2[3];
  v2[0] = v0[1] * v1[2] - v0[2] * v1[1];
  v2[1] = v0[2] * v1[0] - v0[0] * v1[2];
  v2[2] = v0[0] * v1[1] - v0[1] * v1[0];
  const float* rows[3] = {v0, v1, v2}; /* rows of V^T */
  for (int r = 0; r < 3; ++r) {
    T[4 * r + 0] = rows[r][0];
    T[4 * r + 1] = rows[r][1];
    T[4 * r + 2] = rows[r][2];
    float t = rows[r][0] * c[0];
    t = t + rows[r][1] * c[1];
    t = t + rows[r][2] * c[2];
    T[4 * r + 3] = -1.0f * t;
  }
  T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

/* pcl::transformPointCloud with an affine float matrix: ((m0*x + m1*y) + m2*z) + m3 */
static inline void xform_pt(const float T[16], const float* p, float out[3]) {
  for (int r = 0; r < 3; ++r) {
    float v = T[4 * r + 0] * p[0];
    v = v + T[4 * r + 1] * p[1];
    v = v + T[4 * r + 2] * p[2];
    out[r] = v + T[4 * r + 3];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* a7: calHist (LidarCornersEst.cpp:224-301) + get_gray_zone (:303-328)                          */
/* out: min, max, mean, R_L, R_H, gray0, gray1.                                                  */
/* Faithful quirks: bin = round(sample*factor) may equal hist_bins (the max) -> written out of   */
/* bounds in the reference and never read: ignored here.  std::map<double,int>::insert keeps the  */
/* FIRST (lowest) bin for each distinct count.  The walk runs off rend() (UB) if no bin edge lies */
/* on one side of the mean: reported as DEGENERATE_INTENSITY.  max==min divides by zero: same.    */
/* The mean is summed in canonical order (the reference sums the sorted values sequentially;      */
/* identical for integer-valued intensities, which is what Velodyne drivers emit).               */
/* ------------------------------------------------------------------------------------------ */
int ilcc2o_gray_zone(const float* inten, int32_t n, const ilcc2o_params* P, double out[7]) {
  const int NB = P->hist_bins;
  for (int i = 0; i < 7; ++i) out[i] = 0.0;
  if (n <= 0 || NB <= 0 || NB > 4096) return ILCC2O_DEGENERATE_INTENSITY;
  double mn = (double)inten[0], mx = (double)inten[0];
  racc a;
  racc_zero(&a);
  for (int i = 0; i < n; ++i) {
    const double v = (double)inten[i];
    if (v < mn) mn = v;
    if (v > mx) mx = v;
    racc_add(&a, i, v);
  }
  const double mean = racc_total(&a) / (double)n;
  out[0] = mn;
  out[1] = mx;
  out[2] = mean;
  if (!(mx > mn)) return ILCC2O_DEGENERATE_INTENSITY;
  const double factor = (double)NB / (mx - mn);
  int* hist = (int*)calloc((size_t)NB + 1, sizeof(int));
  for (int i = 0; i < n; ++i) {
    const double sample = (double)inten[i] - mn;
    int bin = (int)round(sample * factor);
    if (bin < 0) bin = 0;
    if (bin > NB) bin = NB;
    hist[bin]++;
  }
  const double bin_width = (mx - mn) / (double)NB;
  /* distinct counts, descending; lowest bin per count */
  int low_found = 0, high_found = 0;
  double low = -1.0, high = -1.0;
  int prev_count = INT32_MAX;
  for (;;) {
    /* next distinct count below prev_count */
    int best = -1, best_bin = -1;
    for (int b = 0; b < NB; ++b)
      if (hist[b] < prev_count && hist[b] > best) {
        best = hist[b];
        best_bin = b;
      }
    if (best < 0) break;
    prev_count = best;
    const double edge = bin_width * (double)best_bin + mn;
    if (edge > mean && !high_found) {
      high_found = 1;
      high = edge;
    }
    if (edge < mean && !low_found) {
      low_found = 1;
      low = edge;
    }
    if (low_found && high_found) break;
  }
  free(hist);
  out[3] = low;
  out[4] = high;
  if (!low_found || !high_found) return ILCC2O_DEGENERATE_INTENSITY;
  const double rate = P->gray_rate;
  out[5] = ((rate - 1.0) * low + high) / rate;
  out[6] = (low + (rate - 1.0) * high) / rate;
  return ILCC2O_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* a9: VirtualboardError::operator() (Optimization.h:31-107) with its autodiff Jacobian.        */
/* x = (theta, t_y, t_z).  [UPSTREAM Ceres rotation.h] AngleAxisRotatePoint with angle_axis =     */
/* (theta, 0, 0): for theta^2 > DBL_EPSILON  w0 = theta * (1 / |theta|)  (+-1 or +-(1 - 2^-53)),   */
/* y' = y cos|theta| + (-(w0 z)) sin|theta|,  z' = z cos|theta| + (w0 y) sin|theta|  (the w (w.pt)  */
/* (1 - cos) term is an exact zero); else  y' = y - theta z,  z' = z + theta y.                   */
/* Ceres calls the functor with T = double when only the cost is wanted (candidate evaluation):  */
/* i = (y' + W g / 2) / g is then a true division; with T = Jet (residuals + Jacobian at an      */
/* accepted point) Jet / Jet multiplies by the reciprocal of the divisor: i = (...) * (1 / g).   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int W, H, tlw, oob;
  double g, inv_g, wg2, hg2, huber_a, huber_b;
} vb_model;

typedef struct {
  double c, s; /* cos|theta|, sin|theta| (1, theta in the small-angle branch) */
  double w0;   /* theta / |theta| as Ceres computes it (1 in the small-angle branch, where s carries the sign) */
  double cs, ss; /* cos(theta), sin(theta): the analytic Jacobian below uses the signed pair */
} vb_rot;

static void vb_rot_init(vb_rot* R, double theta) {
  const double theta2 = theta * theta;
  R->cs = cos(theta);
  R->ss = sin(theta);
  if (theta2 > DBL_EPSILON) {
    const double t = sqrt(theta2);
    R->c = cos(t);
    R->s = sin(t);
    const double tinv = 1.0 / t;
    R->w0 = (g_variant & ILCC2O_VAR_AXIS_UNIT) ? (theta < 0.0 ? -1.0 : 1.0) : theta * tinv;
  } else {
    R->c = 1.0;
    R->s = theta;
    R->w0 = 1.0;
  }
}

static inline double vb_point(const vb_model* M, const vb_rot* R, int jet, double y, double z, int white,
                              double ty, double tz, double* si, double* sj) {
  const double yy = (y * R->c + (-(R->w0 * z)) * R->s) + ty;
  const double zz = (z * R->c + (R->w0 * y) * R->s) + tz;
  const int recip = jet || (g_variant & ILCC2O_VAR_COST_RECIP);
  const double fi = recip ? (yy + M->wg2) * M->inv_g : (yy + M->wg2) / M->g;
  const double fj = recip ? (zz + M->hg2) * M->inv_g : (zz + M->hg2) / M->g;
  *si = 0.0;
  *sj = 0.0;
  if (fi > 0.0 && fi < (double)M->W && fj > 0.0 && fj < (double)M->H) {
    const double ifl = floor(fi), jfl = floor(fj);
    const int same = ((((int)ifl) ^ ((int)jfl)) & 1) == 0;
    const int cell_white = same ? M->tlw : !M->tlw;
    if (white == cell_white) return 0.0;
    double ei, ej;
    const double di = fi - ifl, dj = fj - jfl;
    if (di > 0.5) {
      ei = ceil(fi) - fi;
      *si = -1.0;
    } else {
      ei = di;
      *si = 1.0;
    }
    if (dj > 0.5) {
      ej = ceil(fj) - fj;
      *sj = -1.0;
    } else {
      ej = dj;
      *sj = 1.0;
    }
    return ei + ej;
  }
  if (M->oob && !jet && (g_variant & ILCC2O_VAR_INT_ABS)) {
    /* unqualified abs(double) resolving to ::abs(int) (glibc <stdlib.h> under libstdc++ < 6): operands truncated */
    const int a0 = abs((int)fi), a1 = abs((int)(fi - (double)M->W));
    const int b0 = abs((int)fj), b1 = abs((int)(fj - (double)M->H));
    *si = 0.0;
    *sj = 0.0;
    return (double)(a0 < a1 ? a0 : a1) + (double)(b0 < b1 ? b0 : b1);
  }
  if (M->oob) {
    double ei, ej;
    const double a0 = fabs(fi), a1 = fabs(fi - (double)M->W);
    if (a0 < a1) {
      ei = a0;
      *si = fi < 0.0 ? -1.0 : 1.0;
    } else {
      ei = a1;
      *si = (fi - (double)M->W) < 0.0 ? -1.0 : 1.0;
    }
    const double b0 = fabs(fj), b1 = fabs(fj - (double)M->H);
    if (b0 < b1) {
      ej = b0;
      *sj = fj < 0.0 ? -1.0 : 1.0;
    } else {
      ej = b1;
      *sj = (fj - (double)M->H) < 0.0 ? -1.0 : 1.0;
    }
    return ei + ej;
  }
  return 0.0;
}

static void vb_model_init(vb_model* M, int W, int H, double g, int tlw, int oob, double huber) {
  M->W = W;
  M->H = H;
  M->tlw = tlw ? 1 : 0;
  M->oob = oob ? 1 : 0;
  M->g = g;
  M->inv_g = 1.0 / g;
  M->wg2 = (double)W * g / 2.0;
  M->hg2 = (double)H * g / 2.0;
  M->huber_a = huber;
  M->huber_b = huber * huber;
}

double ilcc2o_residual(double y, double z, int white, const double x[3], int W, int H, double g,
                       int topleft_white, int use_oob, double jac[3]) {
  vb_model M;
  vb_model_init(&M, W, H, g, topleft_white, use_oob, 0.1);
  vb_rot R;
  vb_rot_init(&R, x[0]);
  const double c = R.cs, s = R.ss;
  double si, sj;
  const double r = vb_point(&M, &R, jac != NULL, y, z, white, x[1], x[2], &si, &sj);
  if (jac) {
    const double dyy = -(y * s) - (z * c);
    const double dzz = (y * c) - (z * s);
    jac[0] = (si * dyy) * M.inv_g + (sj * dzz) * M.inv_g;
    jac[1] = si * M.inv_g;
    jac[2] = sj * M.inv_g;
  }
  return r;
}

/* [UPSTREAM Ceres] HuberLoss(a): rho(s) = s for s <= a^2 else 2 a sqrt(s) - a^2; cost = rho/2 */
static inline double huber_cost(const vb_model* M, double r) {
  const double s = r * r;
  if (s > M->huber_b) return 0.5 * (2.0 * M->huber_a * r - M->huber_b);
  return 0.5 * s;
}

typedef struct {
  const vb_model* M;
  const double* yz;
  const uint8_t* white;
  int n;
  int n_evals;
} vb_problem;

/* test-only trace of the evaluation points of a solve (per thread): 4 doubles per evaluation = x[0..2], kind
 * (0 cost-only evaluation of a candidate, 1 cost + Jacobian at an accepted point) */
static __thread double* g_trace = NULL;
static __thread int g_trace_cap = 0, g_trace_n = 0;
void ilcc2o_trace(double* buf, int cap) { g_trace = buf; g_trace_cap = cap; g_trace_n = 0; }
int ilcc2o_trace_count(void) { return g_trace_n; }
static void trace_eval(const double x[3], int kind) {
  if (g_trace && g_trace_n < g_trace_cap) {
    double* t = g_trace + 4 * (size_t)g_trace_n;
    t[0] = x[0]; t[1] = x[1]; t[2] = x[2]; t[3] = (double)kind;
  }
  if (g_trace) g_trace_n++;
}

static double eval_cost(vb_problem* Pb, const double x[3]) {
  const vb_model* M = Pb->M;
  trace_eval(x, 0);
  vb_rot R;
  vb_rot_init(&R, x[0]);
  racc a;
  racc_zero_w(&a, RW_WARP);
  for (int k = 0; k < Pb->n; ++k) {
    double si, sj;
    const double r = vb_point(M, &R, 0, Pb->yz[2 * k], Pb->yz[2 * k + 1], Pb->white[k], x[1], x[2], &si, &sj);
    if (r != 0.0) racc_add(&a, k, huber_cost(M, r));
  }
  Pb->n_evals++;
  return racc_total(&a);
}

/* cost, g = J^T r, H = J^T J with the Huber corrector applied ([UPSTREAM Ceres] Corrector:
 * rho'' <= 0 for Huber, so residual and Jacobian row are both scaled by sqrt(rho')).
 * Canonical summation order: the cost is summed like eval_cost (point k -> partial k mod 32); the gradient and
 * J^T J terms only exist for points with a non-zero residual, and the j-th such point (in k order) goes to
 * partial j mod 32 -- the order in which a warp walks the compacted list of non-zero residuals. */
static void eval_jac(vb_problem* Pb, const double x[3], double* cost, double g[3], double H[6]) {
  const vb_model* M = Pb->M;
  trace_eval(x, 1);
  vb_rot R;
  vb_rot_init(&R, x[0]);
  const double c = R.cs, s = R.ss;
  racc a[10];
  for (int q = 0; q < 10; ++q) racc_zero_w(&a[q], RW_WARP);
  int j = 0;
  for (int k = 0; k < Pb->n; ++k) {
    double si, sj;
    const double y = Pb->yz[2 * k], z = Pb->yz[2 * k + 1];
    const double r = vb_point(M, &R, 1, y, z, Pb->white[k], x[1], x[2], &si, &sj);
    if (r == 0.0) continue;
    const double dyy = -(y * s) - (z * c);
    const double dzz = (y * c) - (z * s);
    double j0 = (si * dyy) * M->inv_g + (sj * dzz) * M->inv_g;
    double j1 = si * M->inv_g;
    double j2 = sj * M->inv_g;
    double rr = r;
    const double sq = r * r;
    if (sq > M->huber_b) {
      const double rho1 = M->huber_a / r;
      const double sc = sqrt(rho1);
      j0 *= sc;
      j1 *= sc;
      j2 *= sc;
      rr = sc * r;
    }
    racc_add(&a[0], k, huber_cost(M, r));
    racc_add(&a[1], j, j0 * rr);
    racc_add(&a[2], j, j1 * rr);
    racc_add(&a[3], j, j2 * rr);
    racc_add(&a[4], j, j0 * j0);
    racc_add(&a[5], j, j0 * j1);
    racc_add(&a[6], j, j0 * j2);
    racc_add(&a[7], j, j1 * j1);
    racc_add(&a[8], j, j1 * j2);
    racc_add(&a[9], j, j2 * j2);
    ++j;
  }
  *cost = racc_total(&a[0]);
  for (int q = 0; q < 3; ++q) g[q] = racc_total(&a[1 + q]);
  for (int q = 0; q < 6; ++q) H[q] = racc_total(&a[4 + q]);
  Pb->n_evals++;
}

/* ------------------------------------------------------------------------------------------ */
/* 2-D trust-region boundary problem  min 1/2 x'Bx + g'x  s.t. |x| = radius  (B sym. PSD).       */
/* [UPSTREAM Ceres] DoglegStrategy::FindMinimumOnTrustRegionBoundary builds the quartic in the   */
/* Lagrange multiplier y, takes the real parts of its roots from a companion-matrix eigen solve, */
/* and keeps the candidate x(y) = -(B + yI)^-1 g with the lowest model value.  The global        */
/* minimiser is the root y* > -lambda_min(B); deviation: it is found here by a safeguarded       */
/* Newton iteration on 1/|x(y)| - 1/radius (More-Sorensen), using +,-,*,/,sqrt only.              */
/* ------------------------------------------------------------------------------------------ */
static int tr2_eval(double a, double b, double d, double g0, double g1, double y, double x[2],
                    double* n2, double* q) {
  const double ay = a + y, dy = d + y;
  const double det = ay * dy - b * b;
  if (!(ay > 0.0) || !(det > 0.0)) return 0;
  x[0] = -(dy * g0 - b * g1) / det;
  x[1] = -(ay * g1 - b * g0) / det;
  *n2 = x[0] * x[0] + x[1] * x[1];
  /* w = (B+yI)^-1 x */
  const double w0 = (dy * x[0] - b * x[1]) / det;
  const double w1 = (ay * x[1] - b * x[0]) / det;
  *q = x[0] * w0 + x[1] * w1;
  return 1;
}

/* test-only statistics: calls and Newton iterations of the boundary solver on the calling thread */
static __thread long long g_tr2_calls = 0, g_tr2_iters = 0;
void ilcc2o_tr2_stats(long long* calls, long long* iters, int reset) {
  if (calls) *calls = g_tr2_calls;
  if (iters) *iters = g_tr2_iters;
  if (reset) g_tr2_calls = g_tr2_iters = 0;
}

int ilcc2o_tr_boundary_2d(const double B[3], const double g[2], double radius, double x_out[2]) {
  g_tr2_calls++;
  const double a = B[0], b = B[1], d = B[2];
  const double tr = a + d, hd = 0.5 * (a - d);
  const double disc = sqrt(hd * hd + b * b);
  const double l1 = 0.5 * tr - disc, l2 = 0.5 * tr + disc;
  const double gn = sqrt(g[0] * g[0] + g[1] * g[1]);
  x_out[0] = 0.0;
  x_out[1] = 0.0;
  if (!(gn > 0.0) || !(radius > 0.0)) return 0;
  double lo = gn / radius - l2;
  if (lo < -l1) lo = -l1;
  double hi = gn / radius - l1;
  if (!(hi > lo)) hi = lo + fabs(lo) * 1e-12 + 1e-300;
  double y = 0.0;
  if (y < lo || y > hi) y = lo;
  double x[2] = {0, 0}, n2 = 0, q = 0;
  int have = 0;
  for (int it = 0; it < 100; ++it) {
    g_tr2_iters++;
    if (!tr2_eval(a, b, d, g[0], g[1], y, x, &n2, &q)) {
      lo = y;
      y = 0.5 * (lo + hi);
      if (!(y > lo)) break;
      continue;
    }
    have = 1;
    const double n = sqrt(n2);
    if (fabs(n - radius) <= 1e-14 * radius) break;
    if (n > radius) lo = y; else hi = y;
    double yn = y + ((n - radius) / radius) * (n2 / q);
    if (!(yn > lo && yn < hi)) yn = 0.5 * (lo + hi);
    if (yn == y || !(hi > lo)) break;
    y = yn;
  }
  if (!have) return 0;
  /* accept only if on the boundary to 1e-6: otherwise (hard case) report failure */
  const double n = sqrt(n2);
  if (!(fabs(n - radius) <= 1e-6 * radius)) return 0;
  x_out[0] = x[0];
  x_out[1] = x[1];
  if (g_tr_perturb_ulps) { /* sensitivity probe: Ceres gets this point from quartic roots, not from this iteration */
    x_out[0] = x[0] * (1.0 + (double)g_tr_perturb_ulps * DBL_EPSILON);
    x_out[1] = x[1] * (1.0 - (double)g_tr_perturb_ulps * DBL_EPSILON);
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* [UPSTREAM Ceres 1.13/1.14] DoglegStrategy (SUBSPACE_DOGLEG) on the 3x3 normal equations.      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  double radius, mu;
  int reuse;
  double diag[3], grad[3], gn[3], alpha;
  int sub1d;
  double U[3][2], sub_g[2], sub_B[3];
  double step_norm;
} dogleg_t;

static int chol3_solve(const double A[6] /*00 01 02 11 12 22*/, const double b[3], double x[3]) {
  double l00, l10, l20, l11, l21, l22;
  if (!(A[0] > 0.0)) return 0;
  l00 = sqrt(A[0]);
  l10 = A[1] / l00;
  l20 = A[2] / l00;
  double t = A[3] - l10 * l10;
  if (!(t > 0.0)) return 0;
  l11 = sqrt(t);
  l21 = (A[4] - l20 * l10) / l11;
  t = A[5] - l20 * l20 - l21 * l21;
  if (!(t > 0.0)) return 0;
  l22 = sqrt(t);
  const double y0 = b[0] / l00;
  const double y1 = (b[1] - l10 * y0) / l11;
  const double y2 = (b[2] - l20 * y0 - l21 * y1) / l22;
  x[2] = y2 / l22;
  x[1] = (y1 - l21 * x[2]) / l11;
  x[0] = (y0 - l10 * x[1] - l20 * x[2]) / l00;
  return isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]);
}

static inline double quad3(const double H[6], const double u[3], const double v[3]) {
  /* u' H v */
  const double h0 = H[0] * v[0] + H[1] * v[1] + H[2] * v[2];
  const double h1 = H[1] * v[0] + H[3] * v[1] + H[4] * v[2];
  const double h2 = H[2] * v[0] + H[4] * v[1] + H[5] * v[2];
  return u[0] * h0 + u[1] * h1 + u[2] * h2;
}

static inline double norm3(const double v[3]) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

static void dogleg_traditional(dogleg_t* S, double step[3]) {
  const double gnorm = norm3(S->grad), gnn = norm3(S->gn);
  if (gnn <= S->radius) {
    for (int i = 0; i < 3; ++i) step[i] = S->gn[i] / S->diag[i];
    S->step_norm = gnn;
    return;
  }
  if (gnorm * S->alpha >= S->radius) {
    for (int i = 0; i < 3; ++i) step[i] = (-(S->radius / gnorm) * S->grad[i]) / S->diag[i];
    S->step_norm = S->radius;
    return;
  }
  const double b_dot_a = -S->alpha * (S->grad[0] * S->gn[0] + S->grad[1] * S->gn[1] + S->grad[2] * S->gn[2]);
  const double a2 = (S->alpha * gnorm) * (S->alpha * gnorm);
  const double bma2 = a2 - 2.0 * b_dot_a + gnn * gnn;
  const double c = b_dot_a - a2;
  const double d = sqrt(c * c + bma2 * (S->radius * S->radius - a2));
  const double beta = (c <= 0.0) ? (d - c) / bma2 : (S->radius * S->radius - a2) / (d + c);
  double dl[3];
  for (int i = 0; i < 3; ++i) dl[i] = (-S->alpha * (1.0 - beta)) * S->grad[i] + beta * S->gn[i];
  S->step_norm = norm3(dl);
  for (int i = 0; i < 3; ++i) step[i] = dl[i] / S->diag[i];
}

static void dogleg_subspace_step(dogleg_t* S, double step[3]) {
  const double gnn = norm3(S->gn);
  if (gnn <= S->radius) {
    for (int i = 0; i < 3; ++i) step[i] = S->gn[i] / S->diag[i];
    S->step_norm = gnn;
    return;
  }
  if (S->sub1d) {
    const double gnorm = norm3(S->grad);
    for (int i = 0; i < 3; ++i) step[i] = (-(S->radius / gnorm) * S->grad[i]) / S->diag[i];
    S->step_norm = S->radius;
    return;
  }
  double m2[2];
  if (!ilcc2o_tr_boundary_2d(S->sub_B, S->sub_g, S->radius, m2)) {
    dogleg_traditional(S, step);
    return;
  }
  for (int i = 0; i < 3; ++i) step[i] = (S->U[i][0] * m2[0] + S->U[i][1] * m2[1]) / S->diag[i];
  S->step_norm = S->radius;
}

/* returns 1 on LINEAR_SOLVER_SUCCESS, 0 on failure (step invalid) */
static int dogleg_compute_step(dogleg_t* S, const double Hs[6], const double gs[3], double step[3]) {
  if (S->reuse) {
    dogleg_subspace_step(S, step);
    return 1;
  }
  S->reuse = 1;
  const double hd[3] = {Hs[0], Hs[3], Hs[5]};
  for (int i = 0; i < 3; ++i) {
    double v = hd[i];
    if (v < 1e-6) v = 1e-6;
    if (v > 1e32) v = 1e32;
    S->diag[i] = sqrt(v);
  }
  for (int i = 0; i < 3; ++i) S->grad[i] = gs[i] / S->diag[i];
  /* Cauchy point */
  double sg[3];
  for (int i = 0; i < 3; ++i) sg[i] = S->grad[i] / S->diag[i];
  const double jg2 = quad3(Hs, sg, sg);
  S->alpha = (S->grad[0] * S->grad[0] + S->grad[1] * S->grad[1] + S->grad[2] * S->grad[2]) / jg2;
  /* Gauss-Newton step with mu-regularisation */
  int ok = 0;
  double y[3] = {0, 0, 0};
  while (S->mu < 1.0) {
    double A[6];
    memcpy(A, Hs, sizeof(A));
    const double sm = sqrt(S->mu);
    const double l0 = S->diag[0] * sm, l1 = S->diag[1] * sm, l2 = S->diag[2] * sm;
    A[0] += l0 * l0;
    A[3] += l1 * l1;
    A[5] += l2 * l2;
    if (chol3_solve(A, gs, y)) {
      ok = 1;
      break;
    }
    S->mu *= 10.0;
  }
  if (!ok) return 0;
  for (int i = 0; i < 3; ++i) S->gn[i] = -S->diag[i] * y[i];
  /* subspace model: orthonormal basis of span{grad, gn}, larger column first */
  const double n_g = norm3(S->grad), n_n = norm3(S->gn);
  const double* A0 = (n_g >= n_n) ? S->grad : S->gn;
  const double* A1 = (n_g >= n_n) ? S->gn : S->grad;
  const double na = (n_g >= n_n) ? n_g : n_n;
  if (!(na > 0.0)) return 0; /* rank 0 */
  double u0[3], bp[3];
  for (int i = 0; i < 3; ++i) u0[i] = A0[i] / na;
  const double pr = A1[0] * u0[0] + A1[1] * u0[1] + A1[2] * u0[2];
  for (int i = 0; i < 3; ++i) bp[i] = A1[i] - pr * u0[i];
  const double nb = norm3(bp);
  if (!(nb > na * (2.0 * DBL_EPSILON))) {
    S->sub1d = 1;
  } else {
    S->sub1d = 0;
    double w0[3], w1[3];
    for (int i = 0; i < 3; ++i) {
      S->U[i][0] = u0[i];
      S->U[i][1] = bp[i] / nb;
    }
    S->sub_g[0] = S->U[0][0] * S->grad[0] + S->U[1][0] * S->grad[1] + S->U[2][0] * S->grad[2];
    S->sub_g[1] = S->U[0][1] * S->grad[0] + S->U[1][1] * S->grad[1] + S->U[2][1] * S->grad[2];
    for (int i = 0; i < 3; ++i) {
      w0[i] = S->U[i][0] / S->diag[i];
      w1[i] = S->U[i][1] / S->diag[i];
    }
    S->sub_B[0] = quad3(Hs, w0, w0);
    S->sub_B[1] = quad3(Hs, w0, w1);
    S->sub_B[2] = quad3(Hs, w1, w1);
  }
  dogleg_subspace_step(S, step);
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* a8: Optimization::get_theta_t (Optimization.cpp:94-160) = one ceres::Solve.                  */
/* [UPSTREAM Ceres 1.13/1.14] TrustRegionMinimizer, defaults: 50 iterations, function tolerance  */
/* 1e-6, gradient tolerance 1e-10, parameter tolerance 1e-8, min_relative_decrease 1e-3, initial  */
/* radius 1e4, Jacobi scaling 1/(1+|J_col|) fixed at iteration 0, monotonic steps, at most 5      */
/* consecutive invalid steps; the parameter/function tests run on the candidate before acceptance. */
/* ------------------------------------------------------------------------------------------ */
enum { TERM_MAX_ITER = 0, TERM_GRADIENT = 1, TERM_PARAMETER = 2, TERM_FUNCTION = 3, TERM_MIN_RADIUS = 4,
       TERM_INVALID = 5 };

static int solve_tr(vb_problem* Pb, int max_iter, double x[3], double* final_cost, int* iters_out) {
  double xc[3] = {x[0], x[1], x[2]};
  double x_cost, g[3], H[6];
  eval_jac(Pb, xc, &x_cost, g, H);
  double scale[3];
  scale[0] = 1.0 / (1.0 + sqrt(H[0]));
  scale[1] = 1.0 / (1.0 + sqrt(H[3]));
  scale[2] = 1.0 / (1.0 + sqrt(H[5]));
  double xnorm = norm3(xc);
  double best_cost = x_cost, xbest[3] = {xc[0], xc[1], xc[2]};
  dogleg_t S;
  memset(&S, 0, sizeof(S));
  S.radius = 1e4;
  S.mu = 1e-8;
  S.reuse = 0;
  int iter = 0, invalid = 0, succ = 1, term = TERM_MAX_ITER;
  double gmax = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
  for (;;) {
    if (iter >= max_iter) { term = TERM_MAX_ITER; break; }
    if (succ && gmax <= 1e-10) { term = TERM_GRADIENT; break; }
    if (S.radius <= 1e-32) { term = TERM_MIN_RADIUS; break; }
    iter++;
    double Hs[6], gs[3], step[3];
    Hs[0] = H[0] * scale[0] * scale[0];
    Hs[1] = H[1] * scale[0] * scale[1];
    Hs[2] = H[2] * scale[0] * scale[2];
    Hs[3] = H[3] * scale[1] * scale[1];
    Hs[4] = H[4] * scale[1] * scale[2];
    Hs[5] = H[5] * scale[2] * scale[2];
    for (int i = 0; i < 3; ++i) gs[i] = g[i] * scale[i];
    const int ok = dogleg_compute_step(&S, Hs, gs, step);
    int valid = 0;
    double model_change = 0.0;
    if (ok) {
      model_change = -((step[0] * gs[0] + step[1] * gs[1] + step[2] * gs[2]) + 0.5 * quad3(Hs, step, step));
      valid = model_change > 0.0;
    }
    if (!valid) {
      if (++invalid >= 5) { term = TERM_INVALID; break; }
      S.mu *= 10.0;
      S.reuse = 0;
      succ = 0;
      continue;
    }
    invalid = 0;
    double cand[3];
    for (int i = 0; i < 3; ++i) cand[i] = xc[i] + step[i] * scale[i];
    const double cc = eval_cost(Pb, cand);
    const double d0 = xc[0] - cand[0], d1 = xc[1] - cand[1], d2 = xc[2] - cand[2];
    const double step_norm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    if (step_norm <= 1e-8 * (xnorm + 1e-8)) { term = TERM_PARAMETER; break; }
    const double cost_change = x_cost - cc;
    if (fabs(cost_change) <= 1e-6 * x_cost) { term = TERM_FUNCTION; break; }
    const double quality = cost_change / model_change;
    if (quality > 1e-3) {
      for (int i = 0; i < 3; ++i) xc[i] = cand[i];
      xnorm = norm3(xc);
      eval_jac(Pb, xc, &x_cost, g, H);
      gmax = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
      succ = 1;
      if (quality < 0.25) S.radius *= 0.5;
      if (quality > 0.75) S.radius = fmax(S.radius, 3.0 * S.step_norm);
      S.mu = fmax(1e-8, 2.0 * S.mu / 10.0);
      S.reuse = 0;
      if (x_cost < best_cost) {
        best_cost = x_cost;
        for (int i = 0; i < 3; ++i) xbest[i] = xc[i];
      }
    } else {
      succ = 0;
      S.radius *= 0.5;
      S.reuse = 1;
    }
  }
  for (int i = 0; i < 3; ++i) x[i] = xbest[i];
  *final_cost = best_cost;
  *iters_out = iter;
  return term;
}

int ilcc2o_solve(const double* yz, const uint8_t* white, int32_t n, int W, int H, double g,
                 int topleft_white, int use_oob, double huber, int max_iter,
                 double x[3], double* final_cost, int32_t* iterations, int32_t* n_evals) {
  vb_model M;
  vb_model_init(&M, W, H, g, topleft_white, use_oob, huber);
  vb_problem Pb = {&M, yz, white, n, 0};
  int it = 0;
  const int term = solve_tr(&Pb, max_iter, x, final_cost, &it);
  if (iterations) *iterations = it;
  if (n_evals) *n_evals = Pb.n_evals;
  return term;
}

/* rigid inverse of a row-major float 4x4 [R|t] -> [R^T | -(R^T t)] */
static void rigid_inverse(const float T[16], float O[16]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) O[4 * r + c] = T[4 * c + r];
  for (int r = 0; r < 3; ++r) {
    float v = O[4 * r + 0] * T[3];
    v = v + O[4 * r + 1] * T[7];
    v = v + O[4 * r + 2] * T[11];
    O[4 * r + 3] = -1.0f * v;
  }
  O[12] = 0; O[13] = 0; O[14] = 0; O[15] = 1;
}

/* a11: getPCDcorners (LidarCornersEst.cpp:501-556): lattice in the board frame, then
 * transOptim^-1, then transPCA^-1, all fp32; a10: transf = getTransformation(0,ty,tz,theta,0,0) (:412) */
static void emit_corners(const float pca[16], const double x[3], const ilcc2o_params* P, int inverse,
                         float* corners) {
  int W = P->cells_x, H = P->cells_y;
  if (inverse) {
    W = P->cells_y;
    H = P->cells_x;
  }
  const float thf = (float)x[0];
  const float E = (float)cos((double)thf), F = (float)sin((double)thf);
  float Topt[16] = {1, 0, 0, 0, 0, E, -F, (float)x[1], 0, F, E, (float)x[2], 0, 0, 0, 1};
  float Ti[16], Pi[16];
  rigid_inverse(Topt, Ti);
  rigid_inverse(pca, Pi);
  int k = 0;
  for (int i = 1; i < W; ++i)
    for (int j = 1; j < H; ++j) {
      float p[3], q[3];
      p[0] = 0.0f;
      p[1] = (float)(((double)i - (double)W / 2.0) * P->grid_length);
      p[2] = (float)(((double)j - (double)H / 2.0) * P->grid_length);
      xform_pt(Ti, p, q);
      xform_pt(Pi, q, corners + 3 * k);
      ++k;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* whole path, one frame                                                                      */
/* ------------------------------------------------------------------------------------------ */
int ilcc2o_extract_frame(const void* points, int32_t n, int32_t point_step,
                         int32_t off_x, int32_t off_y, int32_t off_z, int32_t off_i,
                         const float click[3], uint32_t flags, int32_t plane_index_delta,
                         const ilcc2o_params* P, float* corners_out, ilcc2o_diag* diag,
                         int32_t* roi_idx_out, int32_t* cluster_idx_out, int32_t* board_idx_out) {
  const unsigned char* pts = (const unsigned char*)points;
  ilcc2o_diag D;
  memset(&D, 0, sizeof(D));
  const int ncorn = (P->cells_x - 1) * (P->cells_y - 1);
  for (int i = 0; i < 3 * ncorn; ++i) corners_out[i] = NAN;
  int status = ILCC2O_OK;
  int* roi = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  float* rxyz = NULL;
  float* cxyz = NULL;
  int* csel = NULL;
  int* inl = NULL;
  float* bxyz = NULL;
  float* binten = NULL;
  double* yz = NULL;
  uint8_t* white = NULL;
  cc_result cc;
  int have_cc = 0;

  /* a3 */
  const int M = roi_filter(pts, n, point_step, off_x, off_y, off_z, click, P, roi);
  D.n_roi = M;
  if (roi_idx_out) memcpy(roi_idx_out, roi, sizeof(int) * (size_t)M);
  if (M == 0) { status = ILCC2O_NO_POINTS_IN_ROI; goto done; }
  rxyz = (float*)malloc(sizeof(float) * 3 * (size_t)M);
  for (int i = 0; i < M; ++i) {
    const unsigned char* r = pts + (size_t)roi[i] * point_step;
    rxyz[3 * i + 0] = ldf(r, off_x);
    rxyz[3 * i + 1] = ldf(r, off_y);
    rxyz[3 * i + 2] = ldf(r, off_z);
  }
  /* a4 */
  cluster_cc(rxyz, M, P->cluster_tolerance, &cc);
  have_cc = 1;
  {
    /* nearest neighbour of the click (:139-141), ties -> lowest index */
    int nn = 0;
    float best = sqdist3(rxyz, click);
    for (int i = 1; i < M; ++i) {
      const float d = sqdist3(rxyz + 3 * i, click);
      if (d < best) {
        best = d;
        nn = i;
      }
    }
    D.nn_index = nn;
    /* kept clusters ranked by (size desc, first index asc) */
    int nk = 0;
    int* kept = (int*)malloc(sizeof(int) * (size_t)cc.n_found);
    for (int c = 0; c < cc.n_found; ++c)
      if (cc.sizes[c] >= P->min_cluster_size && cc.sizes[c] <= P->max_cluster_size) kept[nk++] = c;
    D.n_clusters = nk;
    if (nk == 0) {
      free(kept);
      status = ILCC2O_NO_CLUSTER;
      goto done;
    }
    /* insertion sort, stable: discovery order already has ascending first index */
    for (int i = 1; i < nk; ++i) {
      const int c = kept[i];
      int j = i - 1;
      while (j >= 0 && cc.sizes[kept[j]] < cc.sizes[c]) {
        kept[j + 1] = kept[j];
        --j;
      }
      kept[j + 1] = c;
    }
    int rank = 0;
    D.click_in_cluster = 0;
    for (int i = 0; i < nk; ++i)
      if (kept[i] == cc.label[nn]) {
        rank = i;
        D.click_in_cluster = 1;
        break;
      }
    /* 'w' / 's' presses (Visualization.cpp:22-29, LidarCornersEst.cpp:164-165) */
    if (plane_index_delta > 0) {
      int r = rank + plane_index_delta;
      if (r > nk - 1) r = nk - 1;
      rank = r;
    } else if (plane_index_delta < 0) {
      int r = rank + plane_index_delta;
      if (r < 0) r = 0;
      rank = r;
    }
    D.cluster_rank = rank;
    const int chosen = kept[rank];
    free(kept);
    const int C = cc.sizes[chosen];
    D.n_cluster_pts = C;
    cxyz = (float*)malloc(sizeof(float) * 3 * (size_t)C);
    csel = (int*)malloc(sizeof(int) * (size_t)C);
    int k = 0;
    for (int i = 0; i < M; ++i)
      if (cc.label[i] == chosen) {
        csel[k] = i;
        memcpy(cxyz + 3 * k, rxyz + 3 * i, sizeof(float) * 3);
        ++k;
      }
    if (cluster_idx_out)
      for (int i = 0; i < C; ++i) cluster_idx_out[i] = roi[csel[i]];
    /* a5 */
    inl = (int*)malloc(sizeof(int) * (size_t)C);
    int n_best = 0, r_iters = 0;
    const int B = plane_ransac(cxyz, C, P, D.plane, inl, &n_best, &r_iters);
    D.ransac_iterations = r_iters;
    D.n_ransac_inliers = n_best;
    D.n_board = B;
    if (board_idx_out)
      for (int i = 0; i < B; ++i) board_idx_out[i] = roi[csel[inl[i]]];
    if (B < 3 || B < P->min_board_points) {
      status = (n_best == 0 && B == 0) ? ILCC2O_RANSAC_FAILED : ILCC2O_PLANE_TOO_SMALL;
      goto done;
    }
    /* a6 */
    bxyz = (float*)malloc(sizeof(float) * 3 * (size_t)B);
    binten = (float*)malloc(sizeof(float) * (size_t)B);
    for (int i = 0; i < B; ++i) {
      memcpy(bxyz + 3 * i, cxyz + 3 * inl[i], sizeof(float) * 3);
      binten[i] = ldf(pts + (size_t)roi[csel[inl[i]]] * point_step, off_i);
    }
    pca_frame(bxyz, B, D.pca, D.eigvals);
    /* a7 */
    double gz[7];
    const int gst = ilcc2o_gray_zone(binten, B, P, gz);
    D.hist_min = gz[0];
    D.hist_max = gz[1];
    D.hist_mean = gz[2];
    D.r_low = gz[3];
    D.r_high = gz[4];
    D.gray[0] = gz[5];
    D.gray[1] = gz[6];
    if (gst != ILCC2O_OK) { status = gst; goto done; }
    /* a8: classify (Optimization.cpp:108-127), points in the PCA frame */
    yz = (double*)malloc(sizeof(double) * 2 * (size_t)B);
    white = (uint8_t*)malloc((size_t)B);
    int S = 0;
    for (int i = 0; i < B; ++i) {
      const double I = (double)binten[i];
      int w;
      if (I < D.gray[0]) w = 0;
      else if (I > D.gray[1]) w = 1;
      else continue;
      float q[3];
      xform_pt(D.pca, bxyz + 3 * i, q);
      yz[2 * S] = (double)q[1];
      yz[2 * S + 1] = (double)q[2];
      white[S] = (uint8_t)w;
      ++S;
    }
    D.n_opt = S;
    if (S == 0) { status = ILCC2O_NO_OPT_POINTS; goto done; }
    /* a10: get_corners default key path (:374-450) */
    const int W = P->cells_x, H = P->cells_y;
    vb_model Mo;
    /* hypothesis sequence: auto = (0, 1) each from zero; manual = 0, then 1 if 'd' was pressed, then (test-only
     * variant) extra pairs of 'd' presses -- every press toggles topleftWhite and re-optimises from the current
     * theta_t (LidarCornersEst.cpp:398-409) */
    int seq[2 + 64], nseq = 0;
    seq[nseq++] = 0;
    if (flags & (ILCC2O_FLAG_AUTO | ILCC2O_FLAG_TOPLEFT_WHITE)) seq[nseq++] = 1;
    if (!(flags & ILCC2O_FLAG_AUTO))
      for (int e = 0; e < 2 * g_extra_d_pairs && nseq < 66; ++e) seq[nseq] = 1 - seq[nseq - 1], ++nseq;
    double xw[3] = {0.0, 0.0, 0.0};
    for (int qi = 0; qi < nseq; ++qi) {
      const int h = seq[qi];
      double x[3] = {0.0, 0.0, 0.0};
      if (qi > 0 && !(flags & ILCC2O_FLAG_AUTO)) memcpy(x, xw, sizeof(x)); /* 'd': warm start */
      int evals = 0;
      for (int phase = 0; phase < 2; ++phase) {
        vb_model_init(&Mo, W, H, P->grid_length, h, phase == 0, P->huber_delta);
        vb_problem Pb = {&Mo, yz, white, S, 0};
        int it = 0;
        double fc = 0.0;
        D.term[h][phase] = solve_tr(&Pb, P->max_solver_iterations, x, &fc, &it);
        D.iters[h][phase] = it;
        D.cost[h] = fc;
        if (phase == 0) D.cost_a[h] = fc;
        evals += Pb.n_evals;
      }
      D.n_evals[h] = evals;
      memcpy(D.theta_t[h], x, sizeof(x));
      memcpy(xw, x, sizeof(x));
    }
    if (flags & ILCC2O_FLAG_AUTO)
      D.hyp_chosen = (D.cost[1] < D.cost[0]) ? 1 : 0;
    else
      D.hyp_chosen = (flags & ILCC2O_FLAG_TOPLEFT_WHITE) ? 1 : 0;
    emit_corners(D.pca, D.theta_t[D.hyp_chosen], P, (flags & ILCC2O_FLAG_INVERSE) != 0, corners_out);
  }
done:
  D.status = status;
  if (diag) *diag = D;
  if (have_cc) cc_free(&cc);
  free(roi);
  free(rxyz);
  free(cxyz);
  free(csel);
  free(inl);
  free(bxyz);
  free(binten);
  free(yz);
  free(white);
  return status;
}

int ilcc2o_extract_batch(const void* points, int32_t n_frames, int32_t n_per_frame, int32_t point_step,
                         int32_t off_x, int32_t off_y, int32_t off_z, int32_t off_i,
                         const float* clicks, const uint32_t* flags, uint32_t default_flags,
                         const ilcc2o_params* P, int32_t threads,
                         float* corners_out, int32_t* status_out, ilcc2o_diag* diag_out) {
  const int ncorn = (P->cells_x - 1) * (P->cells_y - 1);
  int ok = 0;
  rnd_ensure();
#ifdef _OPENMP
  if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : ok)
#endif
  for (int f = 0; f < n_frames; ++f) {
    const unsigned char* base = (const unsigned char*)points + (size_t)f * n_per_frame * point_step;
    const int st = ilcc2o_extract_frame(base, n_per_frame, point_step, off_x, off_y, off_z, off_i,
                                        clicks + 3 * f, flags ? flags[f] : default_flags, 0, P,
                                        corners_out + (size_t)f * ncorn * 3,
                                        diag_out ? diag_out + f : NULL, NULL, NULL, NULL);
    if (status_out) status_out[f] = st;
    if (st == ILCC2O_OK) ok++;
  }
  return ok;
}
