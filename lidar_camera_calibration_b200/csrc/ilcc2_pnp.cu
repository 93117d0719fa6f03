// ilcc2_pnp.cu -- the step that follows the corner path in the reference, on the device:
//   * order_corners_kernel: p <- T_axis_roughly * p, check_order_lidar / check_order_cam
//       (ilcc2/test/calib_lidar_cam.cpp:119-123, ilcc2/src/ImageCornersEst.cpp:430-505) for a batch of frames;
//   * pnp_eval_kernel: one evaluation of Optimization::solvePose3d2dError's problem (ilcc2/src/Optimization.cpp:13-91,
//       Pose3d2dError ilcc2/include/ilcc2/Optimization.h:126-189): Huber cost, J^T r and J^T J over all 3D-2D pairs in
//       fp64, reduced deterministically (fixed per-block partials, summed in block order by the last block to finish),
//       followed IN THE SAME LAUNCH by the trust-region bookkeeping (Ceres TrustRegionMinimizer + subspace dogleg on
//       the 6x6 normal equations) that decides accept/reject and writes the next candidate into device memory.
//     The host enqueues max_iterations + 1 of these launches back to back; a launch that finds the `done` flag set
//     returns at once, so there is no host round trip per iteration (a 50-iteration solve is one stream of launches
//     and one D2H of the result).
// There is no CPU path: ilcc2_pnp_create fails without an sm_100 device.
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "ilcc2_b200.h"

namespace {

constexpr int NP = 6;
constexpr int NSUM = 1 + NP + NP * (NP + 1) / 2;  // cost, J^T r, upper triangle of J^T J = 28
constexpr int kPnpThreads = 256;
constexpr unsigned FULL = 0xffffffffu;

enum { TERM_MAX_ITER = 0, TERM_GRADIENT = 1, TERM_PARAMETER = 2, TERM_FUNCTION = 3, TERM_MIN_RADIUS = 4, TERM_INVALID = 5 };

struct PnpState {
  double x[NP], cand[NP], g[NP], H[NP * NP], scale[NP], xbest[NP];
  double x_cost, xnorm, best_cost, initial_cost, model_change, gmax;
  double radius, mu, diag[NP], grad[NP], gn[NP], alpha, U[NP][2], sub_g[2], sub_B[3], step_norm;
  int reuse, sub1d;
  int iter, invalid, succ, term, done, first, n_evals, n_pairs;
  unsigned int block_counter;
};

struct PnpProblem {
  const double* p3;
  const double* p2;
  const uint8_t* valid;  // nullable
  int m;
  double fx, cx, fy, cy, ha, hb;
  int max_iter;
};

// ---- forward-mode duals with respect to the rotation vector (the 3-D point carries no derivative) ----
struct J3 {
  double a, v0, v1, v2;
};
__device__ __forceinline__ J3 jmul(const J3& f, const J3& g) {
  return {f.a * g.a, f.a * g.v0 + f.v0 * g.a, f.a * g.v1 + f.v1 * g.a, f.a * g.v2 + f.v2 * g.a};
}
__device__ __forceinline__ J3 jmuls(const J3& f, double s) { return {f.a * s, f.v0 * s, f.v1 * s, f.v2 * s}; }
__device__ __forceinline__ J3 jadd(const J3& f, const J3& g) { return {f.a + g.a, f.v0 + g.v0, f.v1 + g.v1, f.v2 + g.v2}; }
__device__ __forceinline__ J3 jsub(const J3& f, const J3& g) { return {f.a - g.a, f.v0 - g.v0, f.v1 - g.v1, f.v2 - g.v2}; }

// Everything of AngleAxisRotatePoint (Ceres rotation.h) that depends on the pose only: computed once per thread.
struct RotTerms {
  bool big;
  J3 ct, st, w[3], omc;  // cos(theta), sin(theta), unit axis, 1 - cos(theta)
  J3 aa[3];             // the rotation vector itself (small-angle branch)
};

__device__ RotTerms rot_terms(const double* x) {
  RotTerms R;
  R.aa[0] = {x[0], 1.0, 0.0, 0.0};
  R.aa[1] = {x[1], 0.0, 1.0, 0.0};
  R.aa[2] = {x[2], 0.0, 0.0, 1.0};
  const J3 th2 = jadd(jadd(jmul(R.aa[0], R.aa[0]), jmul(R.aa[1], R.aa[1])), jmul(R.aa[2], R.aa[2]));
  R.big = th2.a > DBL_EPSILON;
  if (R.big) {
    const double th = sqrt(th2.a), h = 1.0 / (2.0 * th);
    const J3 theta = {th, th2.v0 * h, th2.v1 * h, th2.v2 * h};
    double s, c;
    sincos(th, &s, &c);
    R.ct = {c, -s * theta.v0, -s * theta.v1, -s * theta.v2};
    R.st = {s, c * theta.v0, c * theta.v1, c * theta.v2};
    const double ti = 1.0 / th;  // Jet 1/theta: (0 - ti * theta.v) * ti
    const J3 tinv = {ti, (0.0 - ti * theta.v0) * ti, (0.0 - ti * theta.v1) * ti, (0.0 - ti * theta.v2) * ti};
    for (int i = 0; i < 3; ++i) R.w[i] = jmul(R.aa[i], tinv);
    R.omc = {1.0 - c, 0.0 - R.ct.v0, 0.0 - R.ct.v1, 0.0 - R.ct.v2};
  }
  return R;
}

// residual (2) and 2 x 6 Jacobian of one pair (Optimization.h:155-183 under Ceres autodiff)
__device__ __forceinline__ void pair_residual(const RotTerms& R, const double* x, const PnpProblem& P, double px,
                                              double py, double pz, double ox, double oy, double e[2],
                                              double J[2][NP]) {
  J3 X[3];
  if (R.big) {
    const J3 wxp[3] = {jsub(jmuls(R.w[1], pz), jmuls(R.w[2], py)), jsub(jmuls(R.w[2], px), jmuls(R.w[0], pz)),
                       jsub(jmuls(R.w[0], py), jmuls(R.w[1], px))};
    const J3 tmp = jmul(jadd(jadd(jmuls(R.w[0], px), jmuls(R.w[1], py)), jmuls(R.w[2], pz)), R.omc);
    const double p[3] = {px, py, pz};
    for (int i = 0; i < 3; ++i) X[i] = jadd(jadd(jmuls(R.ct, p[i]), jmul(wxp[i], R.st)), jmul(R.w[i], tmp));
  } else {
    const J3 wxp[3] = {jsub(jmuls(R.aa[1], pz), jmuls(R.aa[2], py)), jsub(jmuls(R.aa[2], px), jmuls(R.aa[0], pz)),
                       jsub(jmuls(R.aa[0], py), jmuls(R.aa[1], px))};
    const double p[3] = {px, py, pz};
    for (int i = 0; i < 3; ++i) X[i] = {p[i] + wxp[i].a, wxp[i].v0, wxp[i].v1, wxp[i].v2};
  }
  const double Xa = X[0].a + x[3], Ya = X[1].a + x[4], Za = X[2].a + x[5];
  const double zi = 1.0 / Za, u = Xa * zi, v = Ya * zi;
  // quotient rule exactly as Jet division does it: (f' - q g') / g
  const double du[NP] = {(X[0].v0 - u * X[2].v0) * zi, (X[0].v1 - u * X[2].v1) * zi, (X[0].v2 - u * X[2].v2) * zi,
                         (1.0 - u * 0.0) * zi, (0.0 - u * 0.0) * zi, (0.0 - u * 1.0) * zi};
  const double dv[NP] = {(X[1].v0 - v * X[2].v0) * zi, (X[1].v1 - v * X[2].v1) * zi, (X[1].v2 - v * X[2].v2) * zi,
                         (0.0 - v * 0.0) * zi, (1.0 - v * 0.0) * zi, (0.0 - v * 1.0) * zi};
  e[0] = ox - (P.fx * u + P.cx);
  e[1] = oy - (P.fy * v + P.cy);
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    J[0][k] = 0.0 - P.fx * du[k];
    J[1][k] = 0.0 - P.fy * dv[k];
  }
}

// ---- 6-parameter subspace dogleg (Ceres DoglegStrategy), executed by one thread ----
__device__ double norm6(const double* v) {
  double s = 0.0;
  for (int i = 0; i < NP; ++i) s += v[i] * v[i];
  return sqrt(s);
}
__device__ double dot6(const double* a, const double* b) {
  double s = 0.0;
  for (int i = 0; i < NP; ++i) s += a[i] * b[i];
  return s;
}
__device__ double quad6(const double* H, const double* u, const double* v) {
  double s = 0.0;
  for (int i = 0; i < NP; ++i) {
    double h = 0.0;
    for (int j = 0; j < NP; ++j) h += H[NP * i + j] * v[j];
    s += u[i] * h;
  }
  return s;
}

__device__ bool chol6_solve(const double* A, const double* b, double* x) {
  double L[NP * NP];
  for (int i = 0; i < NP * NP; ++i) L[i] = 0.0;
  for (int j = 0; j < NP; ++j) {
    double t = A[NP * j + j];
    for (int k = 0; k < j; ++k) t -= L[NP * j + k] * L[NP * j + k];
    if (!(t > 0.0)) return false;
    L[NP * j + j] = sqrt(t);
    for (int i = j + 1; i < NP; ++i) {
      double u = A[NP * i + j];
      for (int k = 0; k < j; ++k) u -= L[NP * i + k] * L[NP * j + k];
      L[NP * i + j] = u / L[NP * j + j];
    }
  }
  double y[NP];
  for (int i = 0; i < NP; ++i) {
    double u = b[i];
    for (int k = 0; k < i; ++k) u -= L[NP * i + k] * y[k];
    y[i] = u / L[NP * i + i];
  }
  for (int i = NP - 1; i >= 0; --i) {
    double u = y[i];
    for (int k = i + 1; k < NP; ++k) u -= L[NP * k + i] * x[k];
    x[i] = u / L[NP * i + i];
  }
  for (int i = 0; i < NP; ++i)
    if (!isfinite(x[i])) return false;
  return true;
}

// min 1/2 x'Bx + g'x on |x| = radius (2-D): safeguarded Newton on the secular equation, as in K5 (ilcc2_kernels.cu)
__device__ bool tr2_point(double a, double b, double d, double g0, double g1, double y, double x[2], double& n2, double& q) {
  const double ay = a + y, dy = d + y, det = ay * dy - b * b;
  if (!(ay > 0.0) || !(det > 0.0)) return false;
  x[0] = -(dy * g0 - b * g1) / det;
  x[1] = -(ay * g1 - b * g0) / det;
  n2 = x[0] * x[0] + x[1] * x[1];
  const double w0 = (dy * x[0] - b * x[1]) / det, w1 = (ay * x[1] - b * x[0]) / det;
  q = x[0] * w0 + x[1] * w1;
  return true;
}

__device__ bool tr_boundary_2d(const double B[3], const double g[2], double radius, double out[2]) {
  const double a = B[0], b = B[1], d = B[2];
  const double tr = a + d, hd = 0.5 * (a - d), disc = sqrt(hd * hd + b * b);
  const double l1 = 0.5 * tr - disc, l2 = 0.5 * tr + disc;
  const double gn = sqrt(g[0] * g[0] + g[1] * g[1]);
  out[0] = out[1] = 0.0;
  if (!(gn > 0.0) || !(radius > 0.0)) return false;
  double lo = gn / radius - l2;
  if (lo < -l1) lo = -l1;
  double hi = gn / radius - l1;
  if (!(hi > lo)) hi = lo + fabs(lo) * 1e-12 + 1e-300;
  double y = 0.0;
  if (y < lo || y > hi) y = lo;
  double x[2] = {0.0, 0.0}, n2 = 0.0, q = 0.0;
  bool have = false;
  for (int it = 0; it < 100; ++it) {
    if (!tr2_point(a, b, d, g[0], g[1], y, x, n2, q)) {
      lo = y;
      y = 0.5 * (lo + hi);
      if (!(y > lo)) break;
      continue;
    }
    have = true;
    const double n = sqrt(n2);
    if (fabs(n - radius) <= 1e-14 * radius) break;
    if (n > radius) lo = y; else hi = y;
    double yn = y + ((n - radius) / radius) * (n2 / q);
    if (!(yn > lo && yn < hi)) yn = 0.5 * (lo + hi);
    if (yn == y || !(hi > lo)) break;
    y = yn;
  }
  if (!have) return false;
  if (!(fabs(sqrt(n2) - radius) <= 1e-6 * radius)) return false;
  out[0] = x[0];
  out[1] = x[1];
  return true;
}

__device__ void dl_traditional(PnpState& S, double* step) {
  const double gnorm = norm6(S.grad), gnn = norm6(S.gn);
  if (gnn <= S.radius) {
    for (int i = 0; i < NP; ++i) step[i] = S.gn[i] / S.diag[i];
    S.step_norm = gnn;
    return;
  }
  if (gnorm * S.alpha >= S.radius) {
    for (int i = 0; i < NP; ++i) step[i] = (-(S.radius / gnorm) * S.grad[i]) / S.diag[i];
    S.step_norm = S.radius;
    return;
  }
  const double b_dot_a = -S.alpha * dot6(S.grad, S.gn);
  const double a2 = (S.alpha * gnorm) * (S.alpha * gnorm);
  const double bma2 = a2 - 2.0 * b_dot_a + gnn * gnn;
  const double c = b_dot_a - a2;
  const double d = sqrt(c * c + bma2 * (S.radius * S.radius - a2));
  const double beta = (c <= 0.0) ? (d - c) / bma2 : (S.radius * S.radius - a2) / (d + c);
  double dl[NP];
  for (int i = 0; i < NP; ++i) dl[i] = (-S.alpha * (1.0 - beta)) * S.grad[i] + beta * S.gn[i];
  S.step_norm = norm6(dl);
  for (int i = 0; i < NP; ++i) step[i] = dl[i] / S.diag[i];
}

__device__ void dl_subspace_step(PnpState& S, double* step) {
  const double gnn = norm6(S.gn);
  if (gnn <= S.radius) {
    for (int i = 0; i < NP; ++i) step[i] = S.gn[i] / S.diag[i];
    S.step_norm = gnn;
    return;
  }
  if (S.sub1d) {
    const double gnorm = norm6(S.grad);
    for (int i = 0; i < NP; ++i) step[i] = (-(S.radius / gnorm) * S.grad[i]) / S.diag[i];
    S.step_norm = S.radius;
    return;
  }
  double m2[2];
  if (!tr_boundary_2d(S.sub_B, S.sub_g, S.radius, m2)) {
    dl_traditional(S, step);
    return;
  }
  for (int i = 0; i < NP; ++i) step[i] = (S.U[i][0] * m2[0] + S.U[i][1] * m2[1]) / S.diag[i];
  S.step_norm = S.radius;
}

__device__ bool dl_compute_step(PnpState& S, const double* Hs, const double* gs, double* step) {
  if (S.reuse) {
    dl_subspace_step(S, step);
    return true;
  }
  S.reuse = 1;
  for (int i = 0; i < NP; ++i) {
    double v = Hs[NP * i + i];
    if (v < 1e-6) v = 1e-6;
    if (v > 1e32) v = 1e32;
    S.diag[i] = sqrt(v);
  }
  double sg[NP];
  for (int i = 0; i < NP; ++i) {
    S.grad[i] = gs[i] / S.diag[i];
    sg[i] = S.grad[i] / S.diag[i];
  }
  S.alpha = dot6(S.grad, S.grad) / quad6(Hs, sg, sg);
  bool ok = false;
  double y[NP];
  while (S.mu < 1.0) {
    double A[NP * NP];
    for (int i = 0; i < NP * NP; ++i) A[i] = Hs[i];
    const double sm = sqrt(S.mu);
    for (int i = 0; i < NP; ++i) {
      const double l = S.diag[i] * sm;
      A[NP * i + i] += l * l;
    }
    if (chol6_solve(A, gs, y)) {
      ok = true;
      break;
    }
    S.mu *= 10.0;
  }
  if (!ok) return false;
  for (int i = 0; i < NP; ++i) S.gn[i] = -S.diag[i] * y[i];
  const double n_g = norm6(S.grad), n_n = norm6(S.gn);
  const bool gfirst = n_g >= n_n;
  const double* A0 = gfirst ? S.grad : S.gn;
  const double* A1 = gfirst ? S.gn : S.grad;
  const double na = gfirst ? n_g : n_n;
  if (!(na > 0.0)) return false;
  double u0[NP], bp[NP];
  for (int i = 0; i < NP; ++i) u0[i] = A0[i] / na;
  const double pr = dot6(A1, u0);
  for (int i = 0; i < NP; ++i) bp[i] = A1[i] - pr * u0[i];
  const double nb = norm6(bp);
  if (!(nb > na * (2.0 * DBL_EPSILON))) {
    S.sub1d = 1;
  } else {
    S.sub1d = 0;
    double w0[NP], w1[NP];
    S.sub_g[0] = 0.0;
    S.sub_g[1] = 0.0;
    for (int i = 0; i < NP; ++i) {
      S.U[i][0] = u0[i];
      S.U[i][1] = bp[i] / nb;
      S.sub_g[0] += S.U[i][0] * S.grad[i];
      S.sub_g[1] += S.U[i][1] * S.grad[i];
      w0[i] = S.U[i][0] / S.diag[i];
      w1[i] = S.U[i][1] / S.diag[i];
    }
    S.sub_B[0] = quad6(Hs, w0, w0);
    S.sub_B[1] = quad6(Hs, w0, w1);
    S.sub_B[2] = quad6(Hs, w1, w1);
  }
  dl_subspace_step(S, step);
  return true;
}

// Trust-region bookkeeping after one evaluation at S.cand (sums = cost, J^T r, upper J^T J): Ceres
// TrustRegionMinimizer with the defaults solvePose3d2dError leaves in place (50 iterations, function tolerance 1e-6,
// gradient tolerance 1e-10, parameter tolerance 1e-8, min_relative_decrease 1e-3, radius 1e4, Jacobi scaling).
__device__ void trust_region_advance(PnpState& S, const double* sums, int max_iter) {
  const double cc = sums[0];
  bool accept = true;
  double quality = 1.0;
  S.n_evals++;
  if (!S.first) {
    double dx[NP];
    for (int i = 0; i < NP; ++i) dx[i] = S.x[i] - S.cand[i];
    if (norm6(dx) <= 1e-8 * (S.xnorm + 1e-8)) { S.term = TERM_PARAMETER; S.done = 1; return; }
    const double cost_change = S.x_cost - cc;
    if (fabs(cost_change) <= 1e-6 * S.x_cost) { S.term = TERM_FUNCTION; S.done = 1; return; }
    quality = cost_change / S.model_change;
    accept = quality > 1e-3;
  }
  if (accept) {
    for (int i = 0; i < NP; ++i) S.x[i] = S.cand[i];
    S.xnorm = norm6(S.x);
    S.x_cost = cc;
    int q = 1 + NP;
    for (int a = 0; a < NP; ++a) {
      S.g[a] = sums[1 + a];
      for (int b = a; b < NP; ++b) {
        S.H[NP * a + b] = sums[q];
        S.H[NP * b + a] = sums[q];
        ++q;
      }
    }
    S.gmax = 0.0;
    for (int i = 0; i < NP; ++i) S.gmax = fmax(S.gmax, fabs(S.g[i]));
    S.succ = 1;
    if (S.first) {
      for (int i = 0; i < NP; ++i) S.scale[i] = 1.0 / (1.0 + sqrt(S.H[NP * i + i]));
      S.best_cost = cc;
      S.initial_cost = cc;
      for (int i = 0; i < NP; ++i) S.xbest[i] = S.x[i];
    } else {
      S.n_evals++;  // Ceres evaluates the accepted point a second time for the Jacobian
      if (quality < 0.25) S.radius *= 0.5;
      if (quality > 0.75) S.radius = fmax(S.radius, 3.0 * S.step_norm);
      S.mu = fmax(1e-8, 2.0 * S.mu / 10.0);
      S.reuse = 0;
      if (S.x_cost < S.best_cost) {
        S.best_cost = S.x_cost;
        for (int i = 0; i < NP; ++i) S.xbest[i] = S.x[i];
      }
    }
  } else {
    S.succ = 0;
    S.radius *= 0.5;
    S.reuse = 1;
  }
  S.first = 0;
  for (;;) {
    if (S.iter >= max_iter) { S.term = TERM_MAX_ITER; S.done = 1; return; }
    if (S.succ && S.gmax <= 1e-10) { S.term = TERM_GRADIENT; S.done = 1; return; }
    if (S.radius <= 1e-32) { S.term = TERM_MIN_RADIUS; S.done = 1; return; }
    S.iter++;
    double Hs[NP * NP], gs[NP], step[NP];
    for (int i = 0; i < NP; ++i) {
      gs[i] = S.g[i] * S.scale[i];
      for (int j = 0; j < NP; ++j) Hs[NP * i + j] = S.H[NP * i + j] * S.scale[i] * S.scale[j];
    }
    const bool ok = dl_compute_step(S, Hs, gs, step);
    bool valid = false;
    if (ok) {
      S.model_change = -(dot6(step, gs) + 0.5 * quad6(Hs, step, step));
      valid = S.model_change > 0.0;
    }
    if (valid) {
      S.invalid = 0;
      for (int i = 0; i < NP; ++i) S.cand[i] = S.x[i] + step[i] * S.scale[i];
      return;
    }
    if (++S.invalid >= 5) { S.term = TERM_INVALID; S.done = 1; return; }
    S.mu *= 10.0;
    S.reuse = 0;
    S.succ = 0;
  }
}

__global__ void __launch_bounds__(kPnpThreads) pnp_eval_kernel(const PnpProblem P, PnpState* st, double* partials) {
  __shared__ double s_warp[kPnpThreads / 32][NSUM];
  __shared__ double s_sum[NSUM];
  __shared__ bool s_last;
  if (st->done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double x[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) x[i] = st->cand[i];
  const RotTerms R = rot_terms(x);
  double acc[NSUM];
#pragma unroll
  for (int q = 0; q < NSUM; ++q) acc[q] = 0.0;
  for (int i = blockIdx.x * kPnpThreads + tid; i < P.m; i += gridDim.x * kPnpThreads) {
    if (P.valid && !P.valid[i]) continue;
    double e[2], J[2][NP];
    pair_residual(R, x, P, P.p3[3 * i], P.p3[3 * i + 1], P.p3[3 * i + 2], P.p2[2 * i], P.p2[2 * i + 1], e, J);
    // HuberLoss(a) on s = |e|^2 and Ceres' Corrector (rho'' <= 0: scale residual and Jacobian rows by sqrt(rho'))
    const double s = e[0] * e[0] + e[1] * e[1];
    double rho0 = s, rho1 = 1.0;
    if (s > P.hb) {
      const double r = sqrt(s);
      rho0 = 2.0 * P.ha * r - P.hb;
      rho1 = fmax(P.ha / r, DBL_MIN);
    }
    const double sc = sqrt(rho1);
    const double r0 = e[0] * sc, r1 = e[1] * sc;
    acc[0] += 0.5 * rho0;
    int q = 1 + NP;
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      const double ja0 = J[0][a] * sc, ja1 = J[1][a] * sc;
      acc[1 + a] += ja0 * r0 + ja1 * r1;
#pragma unroll
      for (int b = a; b < NP; ++b) {
        acc[q] += ja0 * (J[0][b] * sc) + ja1 * (J[1][b] * sc);
        ++q;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NSUM; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0) s_warp[warp][q] = v;
  }
  __syncthreads();
  if (tid < NSUM) {
    double v = 0.0;
    for (int w = 0; w < kPnpThreads / 32; ++w) v += s_warp[w][tid];
    partials[(size_t)blockIdx.x * NSUM + tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&st->block_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < NSUM) {  // fixed order over blocks: the result does not depend on which block finished last
    double v = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(partials + (size_t)b * NSUM + tid);
    s_sum[tid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    st->block_counter = 0;
    trust_region_advance(*st, s_sum, P.max_iter);
  }
}

// calib_lidar_cam.cpp:119-123 for a batch: one thread per (frame, corner).  The two flip decisions of check_order_*
// (rows first, then columns ON THE ROW-FLIPPED array) are recomputed by every thread from the four corners involved.
__global__ void order_corners_kernel(const float* __restrict__ lidar, const int32_t* __restrict__ status,
                                     const double* __restrict__ cam_in, int n_frames, int bw, int bh, const double* T,
                                     double* __restrict__ p3_out, double* __restrict__ p2_out,
                                     uint8_t* __restrict__ valid_out) {
  const int nc = bw * bh;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n_frames * nc) return;
  const int f = (int)(gid / nc), k = (int)(gid % nc), h = k / bw, w = k % bw;
  const float* L = lidar + (size_t)f * nc * 3;
  auto axis = [&](int idx, double out[3]) {  // Eigen::Isometry3d * Vector3d on the float corner widened to double
    const double x = (double)L[3 * idx], y = (double)L[3 * idx + 1], z = (double)L[3 * idx + 2];
    for (int r = 0; r < 3; ++r) out[r] = ((T[4 * r] * x + T[4 * r + 1] * y) + T[4 * r + 2] * z) + T[4 * r + 3];
  };
  {
    double a[3], b[3];
    axis(0, a);
    axis(bw + 1, b);
    const bool rowflip = a[1] > b[1];
    const int r0 = rowflip ? bh - 1 : 0;
    axis(r0 * bw, a);
    axis(r0 * bw + 1, b);
    const bool colflip = a[0] > b[0];
    const int src = (rowflip ? bh - 1 - h : h) * bw + (colflip ? bw - 1 - w : w);
    double o[3];
    axis(src, o);
    double* dst = p3_out + (size_t)gid * 3;
    dst[0] = o[0];
    dst[1] = o[1];
    dst[2] = o[2];
  }
  if (cam_in) {
    const double* Cm = cam_in + (size_t)f * nc * 2;
    const bool rowflip = Cm[1] > Cm[2 * (bw + 1) + 1];
    const int r0 = rowflip ? bh - 1 : 0;
    const bool colflip = Cm[2 * (r0 * bw)] > Cm[2 * (r0 * bw + 1)];
    const int src = (rowflip ? bh - 1 - h : h) * bw + (colflip ? bw - 1 - w : w);
    p2_out[(size_t)gid * 2] = Cm[2 * src];
    p2_out[(size_t)gid * 2 + 1] = Cm[2 * src + 1];
  }
  if (valid_out) valid_out[gid] = (!status || status[f] == ILCC2_OK) ? 1 : 0;
}

// ImageCornersEst::spaceToPlane (ImageCornersEst.cpp:135-155) + the per-point colouring of pcd2image.cpp:59-84
// (HSVtoRGB of the intensity, ImageCornersEst.cpp:372-428) or rgblidar.cpp:53-73 (colour of the pixel the point falls on).
// One thread per point, grid-stride; 16 or 32 B in, 12 B out per point: HBM-bound.
struct ProjParams {
  double R[9], t[3], fx, cx, fy, cy, dist, inten_low, inten_high;
  int width, height, step, ox, oy, oz, oi;
};

__device__ __forceinline__ uchar4 hsv_to_rgb_100(int h) {  // HSVtoRGB(h, 100, 100, ...): S = V = 100
  const float RGB_max = 100 * 2.55f;
  const float RGB_min = RGB_max * (100 - 100) / 100.0f;
  const int i = h / 60, difs = h % 60;
  const float adj = (RGB_max - RGB_min) * difs / 60.0f;
  float r, g, b;
  switch (i) {
    case 0: r = RGB_max; g = RGB_min + adj; b = RGB_min; break;
    case 1: r = RGB_max - adj; g = RGB_max; b = RGB_min; break;
    case 2: r = RGB_min; g = RGB_max; b = RGB_min + adj; break;
    case 3: r = RGB_min; g = RGB_max - adj; b = RGB_max; break;
    case 4: r = RGB_min + adj; g = RGB_min; b = RGB_max; break;
    default: r = RGB_max; g = RGB_min; b = RGB_max - adj; break;
  }
  // float -> unsigned char as the reference's implicit conversion (truncation); values outside 0..255 are undefined
  // behaviour there (negative hue) and are clamped here
  auto cv8 = [](float v) { return (unsigned char)(int)fminf(fmaxf(v, 0.f), 255.f); };
  return make_uchar4(cv8(r), cv8(g), cv8(b), 1);
}

struct ProjPoint {
  float x, y, z, in;
};

template <int MODE>  // 0: packed 16-byte XYZI, 1: 32-byte records (x,y,z@0, intensity@16), 2: generic offsets
__device__ __forceinline__ ProjPoint proj_load(const unsigned char* rec, const ProjParams& P) {
  ProjPoint q;
  if (MODE == 0) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(rec));
    q.x = v.x; q.y = v.y; q.z = v.z; q.in = v.w;
  } else if (MODE == 1) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(rec));
    q.x = a.x; q.y = a.y; q.z = a.z;
    q.in = __ldcs(reinterpret_cast<const float*>(rec + 16));
  } else {
    q.x = *reinterpret_cast<const float*>(rec + P.ox);
    q.y = *reinterpret_cast<const float*>(rec + P.oy);
    q.z = *reinterpret_cast<const float*>(rec + P.oz);
    q.in = *reinterpret_cast<const float*>(rec + P.oi);
  }
  return q;
}

__device__ __forceinline__ void proj_one(const ProjPoint& q, const ProjParams& P, const unsigned char* __restrict__ image,
                                         int2& px, uchar4& c) {
  const double X = (double)q.x, Y = (double)q.y, Z = (double)q.z;
  px = make_int2(-1, -1);
  c = make_uchar4(0, 0, 0, 0);
  const double pcz = ((P.R[6] * X + P.R[7] * Y) + P.R[8] * Z) + P.t[2];  // Eigen: m_R * P_w + m_t
  if (pcz < 0.0 || pcz > P.dist || pcz != pcz) return;
  const double pcx = ((P.R[0] * X + P.R[1] * Y) + P.R[2] * Z) + P.t[0];
  const double pcy = ((P.R[3] * X + P.R[4] * Y) + P.R[5] * Z) + P.t[1];
  const double u = P.fx * (pcx / pcz) + P.cx, v = P.fy * (pcy / pcz) + P.cy;
  if (u > 0.0 && u < (double)P.width && v > 0.0 && v < (double)P.height) {
    px = make_int2((int)u, (int)v);
    if (image) {
      const unsigned char* g = image + ((size_t)px.y * P.width + px.x) * 3;  // bgr8
      c = make_uchar4(g[2], g[1], g[0], 1);
    } else {
      const double h = ((double)q.in - P.inten_low) / (P.inten_high - P.inten_low) * 255.0;
      c = hsv_to_rgb_100(h == h ? (int)h : 0);
    }
  }
}

// four independent points per thread and iteration: the loads are issued before any of the arithmetic
template <int MODE>
__global__ void __launch_bounds__(256) project_points_kernel(const unsigned char* __restrict__ pts, long long n,
                                                             const ProjParams P, const unsigned char* __restrict__ image,
                                                             int2* __restrict__ pix, uchar4* __restrict__ rgb) {
  constexpr int U = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
    ProjPoint q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n) q[u] = proj_load<MODE>(pts + (size_t)i * P.step, P);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n) {
        int2 px;
        uchar4 c;
        proj_one(q[u], P, image, px, c);
        __stcs(pix + i, px);
        __stcs(rgb + i, c);
      }
    }
  }
}

thread_local std::string g_pnp_create_error;

}  // namespace

struct ilcc2_pnp {
  int device = 0;
  int num_sms = 148;
  PnpState* d_state = nullptr;
  double* d_partials = nullptr;
  int partial_blocks = 0;
  double* d_T = nullptr;
  int64_t launches = 0;
  std::string error;
};

namespace {

int pfail(ilcc2_pnp* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h)
    h->error = buf;
  else
    g_pnp_create_error = buf;
  return code;
}

#define PNP_TRY(h, expr)                                                                               \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) return pfail(h, ILCC2_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// device copy of a host array for ILCC2_MEM_HOST calls (the PnP is not the hot path: plain malloc/free per call)
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t upload(const void* host, size_t bytes, cudaStream_t s) {
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) return e;
    return host && bytes ? cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
  }
};

void rot_axis(int axis, double ang, double R[9]) {
  const double c = std::cos(ang), s = std::sin(ang);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::memcpy(R, I, sizeof(I));
  if (axis == 0) { R[4] = c; R[5] = -s; R[7] = s; R[8] = c; }
  if (axis == 1) { R[0] = c; R[2] = s; R[6] = -s; R[8] = c; }
  if (axis == 2) { R[0] = c; R[1] = -s; R[3] = s; R[4] = c; }
}

void mul33(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

}  // namespace

extern "C" {

int ilcc2_pnp_create(int device, ilcc2_pnp** out) {
  if (!out) return pfail(nullptr, ILCC2_ERR_INVALID_ARGUMENT, "out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return pfail(nullptr, ILCC2_ERR_CUDA, "no CUDA device (%s); this library has no CPU path",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return pfail(nullptr, ILCC2_ERR_INVALID_ARGUMENT, "device %d out of range", device);
  PNP_TRY(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  PNP_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return pfail(nullptr, ILCC2_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                 prop.major, prop.minor);
  ilcc2_pnp* h = new ilcc2_pnp();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->partial_blocks = 2 * h->num_sms;
  cudaError_t ce = cudaMalloc((void**)&h->d_state, sizeof(PnpState));
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&h->d_partials, sizeof(double) * NSUM * h->partial_blocks);
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&h->d_T, sizeof(double) * 16);
  if (ce != cudaSuccess) {
    ilcc2_pnp_destroy(h);
    return pfail(nullptr, ILCC2_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(ce));
  }
  *out = h;
  return ILCC2_SUCCESS;
}

int ilcc2_pnp_destroy(ilcc2_pnp* h) {
  if (!h) return ILCC2_SUCCESS;
  cudaSetDevice(h->device);
  if (h->d_state) cudaFree(h->d_state);
  if (h->d_partials) cudaFree(h->d_partials);
  if (h->d_T) cudaFree(h->d_T);
  delete h;
  return ILCC2_SUCCESS;
}

const char* ilcc2_pnp_last_error(const ilcc2_pnp* h) { return h ? h->error.c_str() : g_pnp_create_error.c_str(); }

int64_t ilcc2_pnp_launch_count(const ilcc2_pnp* h) { return h ? h->launches : 0; }

int ilcc2_axis_roughly(const char* cam, double T[16]) {
  if (!cam || !T) return ILCC2_ERR_INVALID_ARGUMENT;
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, A[9], B[9];
  const std::string c(cam);
  if (c == "front" || c == "car_left" || c == "pointgrey") {
    rot_axis(1, -1.57, A);
    rot_axis(0, 1.57, B);
    mul33(A, B, R);
  } else if (c == "left") {
    rot_axis(0, 1.57, R);
  } else if (c == "right") {
    rot_axis(0, 1.57, A);
    rot_axis(2, 3.14, B);
    mul33(A, B, R);
  } else if (c == "back") {
    rot_axis(1, 1.57, A);
    rot_axis(0, 1.57, B);
    mul33(A, B, R);
  } else {
    return ILCC2_ERR_INVALID_ARGUMENT;  // the reference leaves the rotation uninitialised for other names
  }
  for (int i = 0; i < 16; ++i) T[i] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
  T[15] = 1.0;
  return ILCC2_SUCCESS;
}

int ilcc2_pnp_order_corners(ilcc2_pnp* h, const float* lidar_corners, const int32_t* status, const double* cam_corners,
                            int n_frames, int board_w, int board_h, const double T_axis[16], double* pts3_out,
                            double* pts2_out, uint8_t* valid_out, int mem_kind, void* stream) {
  if (!h) return ILCC2_ERR_INVALID_ARGUMENT;
  if (!lidar_corners || !T_axis || !pts3_out || n_frames < 1 || board_w < 2 || board_h < 2 || (cam_corners && !pts2_out))
    return pfail(h, ILCC2_ERR_INVALID_ARGUMENT, "order_corners: null or out-of-range argument");
  if (mem_kind != ILCC2_MEM_HOST && mem_kind != ILCC2_MEM_DEVICE)
    return pfail(h, ILCC2_ERR_INVALID_ARGUMENT, "mem_kind must be ILCC2_MEM_HOST or ILCC2_MEM_DEVICE");
  PNP_TRY(h, cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)n_frames * board_w * board_h;
  PNP_TRY(h, cudaMemcpyAsync(h->d_T, T_axis, sizeof(double) * 16, cudaMemcpyHostToDevice, s));
  const int threads = 256, blocks = (int)((n + threads - 1) / threads);
  if (mem_kind == ILCC2_MEM_DEVICE) {
    order_corners_kernel<<<blocks, threads, 0, s>>>(lidar_corners, status, cam_corners, n_frames, board_w, board_h,
                                                    h->d_T, pts3_out, pts2_out, valid_out);
    h->launches++;
    PNP_TRY(h, cudaGetLastError());
    return ILCC2_SUCCESS;
  }
  DevBuf dl, ds, dc, o3, o2, ov;
  PNP_TRY(h, dl.upload(lidar_corners, n * 3 * sizeof(float), s));
  if (status) PNP_TRY(h, ds.upload(status, (size_t)n_frames * sizeof(int32_t), s));
  if (cam_corners) PNP_TRY(h, dc.upload(cam_corners, n * 2 * sizeof(double), s));
  PNP_TRY(h, o3.upload(nullptr, n * 3 * sizeof(double), s));
  PNP_TRY(h, o2.upload(nullptr, n * 2 * sizeof(double), s));
  PNP_TRY(h, ov.upload(nullptr, n, s));
  order_corners_kernel<<<blocks, threads, 0, s>>>((const float*)dl.p, status ? (const int32_t*)ds.p : nullptr,
                                                  cam_corners ? (const double*)dc.p : nullptr, n_frames, board_w,
                                                  board_h, h->d_T, (double*)o3.p, (double*)o2.p, (uint8_t*)ov.p);
  h->launches++;
  PNP_TRY(h, cudaGetLastError());
  PNP_TRY(h, cudaMemcpyAsync(pts3_out, o3.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (cam_corners) PNP_TRY(h, cudaMemcpyAsync(pts2_out, o2.p, n * 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (valid_out) PNP_TRY(h, cudaMemcpyAsync(valid_out, ov.p, n, cudaMemcpyDeviceToHost, s));
  PNP_TRY(h, cudaStreamSynchronize(s));
  return ILCC2_SUCCESS;
}

int ilcc2_pnp_solve(ilcc2_pnp* h, const double* pts3, const double* pts2, const uint8_t* valid, int n_pairs,
                    const double cam[4], double huber_delta, int max_iterations, double r[3], double t[3],
                    ilcc2_pnp_summary* summary, int mem_kind, void* stream) {
  if (!h) return ILCC2_ERR_INVALID_ARGUMENT;
  if (!pts3 || !pts2 || !cam || !r || !t || n_pairs < 1 || !(huber_delta > 0) || max_iterations < 1 || max_iterations > 10000)
    return pfail(h, ILCC2_ERR_INVALID_ARGUMENT, "solve: null or out-of-range argument");
  if (mem_kind != ILCC2_MEM_HOST && mem_kind != ILCC2_MEM_DEVICE)
    return pfail(h, ILCC2_ERR_INVALID_ARGUMENT, "mem_kind must be ILCC2_MEM_HOST or ILCC2_MEM_DEVICE");
  PNP_TRY(h, cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DevBuf d3, d2, dv;
  PnpProblem P;
  if (mem_kind == ILCC2_MEM_HOST) {
    PNP_TRY(h, d3.upload(pts3, (size_t)n_pairs * 3 * sizeof(double), s));
    PNP_TRY(h, d2.upload(pts2, (size_t)n_pairs * 2 * sizeof(double), s));
    if (valid) PNP_TRY(h, dv.upload(valid, (size_t)n_pairs, s));
    P.p3 = (const double*)d3.p;
    P.p2 = (const double*)d2.p;
    P.valid = valid ? (const uint8_t*)dv.p : nullptr;
  } else {
    P.p3 = pts3;
    P.p2 = pts2;
    P.valid = valid;
  }
  P.m = n_pairs;
  P.fx = cam[0];
  P.cx = cam[1];
  P.fy = cam[2];
  P.cy = cam[3];
  P.ha = huber_delta;
  P.hb = huber_delta * huber_delta;
  P.max_iter = max_iterations;
  PnpState init;
  std::memset(&init, 0, sizeof(init));
  for (int i = 0; i < 3; ++i) {
    init.x[i] = init.cand[i] = init.xbest[i] = r[i];
    init.x[3 + i] = init.cand[3 + i] = init.xbest[3 + i] = t[i];
  }
  init.radius = 1e4;
  init.mu = 1e-8;
  init.first = 1;
  init.succ = 1;
  PNP_TRY(h, cudaMemcpyAsync(h->d_state, &init, sizeof(init), cudaMemcpyHostToDevice, s));
  int blocks = (n_pairs + kPnpThreads - 1) / kPnpThreads;
  if (blocks > h->partial_blocks) blocks = h->partial_blocks;
  // one evaluation per trust-region iteration (+ the start point); launches after convergence return immediately
  for (int it = 0; it <= max_iterations; ++it) pnp_eval_kernel<<<blocks, kPnpThreads, 0, s>>>(P, h->d_state, h->d_partials);
  h->launches += max_iterations + 1;
  PNP_TRY(h, cudaGetLastError());
  PnpState fin;
  PNP_TRY(h, cudaMemcpyAsync(&fin, h->d_state, sizeof(fin), cudaMemcpyDeviceToHost, s));
  PNP_TRY(h, cudaStreamSynchronize(s));
  for (int i = 0; i < 3; ++i) {
    r[i] = fin.xbest[i];
    t[i] = fin.xbest[3 + i];
  }
  if (summary) {
    summary->initial_cost = fin.initial_cost;
    summary->final_cost = fin.best_cost;
    summary->iterations = fin.iter;
    summary->termination = fin.term;
    summary->n_evals = fin.n_evals;
    summary->n_pairs = -1;  // not counted on the device
  }
  return ILCC2_SUCCESS;
}

int ilcc2_project_points(ilcc2_pnp* h, const void* points, int64_t n_points, int point_step, int off_x, int off_y,
                         int off_z, int off_intensity, const double T_lidar2cam[16], const double camera[4], int width,
                         int height, double distance_valid, double inten_low, double inten_high,
                         const uint8_t* image_bgr, int32_t* pix_xy_out, uint8_t* rgbv_out, int mem_kind, void* stream) {
  if (!h) return ILCC2_ERR_INVALID_ARGUMENT;
  if (!points || !T_lidar2cam || !camera || !pix_xy_out || !rgbv_out || n_points < 1 || point_step < 16 || width < 1 ||
      height < 1 || off_x < 0 || off_y < 0 || off_z < 0 || off_intensity < 0 || off_x + 4 > point_step ||
      off_y + 4 > point_step || off_z + 4 > point_step || off_intensity + 4 > point_step || (point_step & 3) ||
      ((off_x | off_y | off_z | off_intensity) & 3))
    return pfail(h, ILCC2_ERR_INVALID_ARGUMENT, "project_points: null or out-of-range argument");
  if (mem_kind != ILCC2_MEM_HOST && mem_kind != ILCC2_MEM_DEVICE)
    return pfail(h, ILCC2_ERR_INVALID_ARGUMENT, "mem_kind must be ILCC2_MEM_HOST or ILCC2_MEM_DEVICE");
  PNP_TRY(h, cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ProjParams P;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) P.R[3 * i + j] = T_lidar2cam[4 * i + j];
    P.t[i] = T_lidar2cam[4 * i + 3];
  }
  P.fx = camera[0]; P.cx = camera[1]; P.fy = camera[2]; P.cy = camera[3];
  P.dist = distance_valid; P.inten_low = inten_low; P.inten_high = inten_high;
  P.width = width; P.height = height; P.step = point_step;
  P.ox = off_x; P.oy = off_y; P.oz = off_z; P.oi = off_intensity;
  const unsigned char* d_pts = static_cast<const unsigned char*>(points);
  const unsigned char* d_img = image_bgr;
  int2* d_pix = reinterpret_cast<int2*>(pix_xy_out);
  uchar4* d_rgb = reinterpret_cast<uchar4*>(rgbv_out);
  DevBuf bp, bi, bo, bc;
  if (mem_kind == ILCC2_MEM_HOST) {
    PNP_TRY(h, bp.upload(points, (size_t)n_points * point_step, s));
    if (image_bgr) PNP_TRY(h, bi.upload(image_bgr, (size_t)width * height * 3, s));
    PNP_TRY(h, bo.upload(nullptr, (size_t)n_points * sizeof(int2), s));
    PNP_TRY(h, bc.upload(nullptr, (size_t)n_points * sizeof(uchar4), s));
    d_pts = (const unsigned char*)bp.p;
    d_img = image_bgr ? (const unsigned char*)bi.p : nullptr;
    d_pix = (int2*)bo.p;
    d_rgb = (uchar4*)bc.p;
  }
  const int threads = 256;
  long long blocks = (n_points + threads - 1) / threads;
  const long long cap = 8ll * h->num_sms * 4;  // 8 resident CTAs per SM, 4 points per thread at full size
  if (blocks > cap) blocks = cap;
  const bool aligned16 = (reinterpret_cast<uintptr_t>(d_pts) & 15u) == 0;  // 128-bit loads; else the generic path
  const bool packed = aligned16 && point_step == 16 && off_x == 0 && off_y == 4 && off_z == 8 && off_intensity == 12;
  const bool pcl32 = aligned16 && point_step == 32 && off_x == 0 && off_y == 4 && off_z == 8 && off_intensity == 16;
  if (packed)
    project_points_kernel<0><<<(int)blocks, threads, 0, s>>>(d_pts, n_points, P, d_img, d_pix, d_rgb);
  else if (pcl32)
    project_points_kernel<1><<<(int)blocks, threads, 0, s>>>(d_pts, n_points, P, d_img, d_pix, d_rgb);
  else
    project_points_kernel<2><<<(int)blocks, threads, 0, s>>>(d_pts, n_points, P, d_img, d_pix, d_rgb);
  h->launches++;
  PNP_TRY(h, cudaGetLastError());
  if (mem_kind == ILCC2_MEM_HOST) {
    PNP_TRY(h, cudaMemcpyAsync(pix_xy_out, d_pix, (size_t)n_points * sizeof(int2), cudaMemcpyDeviceToHost, s));
    PNP_TRY(h, cudaMemcpyAsync(rgbv_out, d_rgb, (size_t)n_points * sizeof(uchar4), cudaMemcpyDeviceToHost, s));
    PNP_TRY(h, cudaStreamSynchronize(s));
  }
  return ILCC2_SUCCESS;
}

// Optimization.cpp:75-88 + calib_lidar_cam.cpp:157-160: [AngleAxis(|r|, r/|r|) | t] * T_axis_roughly (row-major 4x4)
int ilcc2_compose_extrinsic(const double r[3], const double t[3], const double T_axis[16], double T_out[16]) {
  if (!r || !t || !T_axis || !T_out) return ILCC2_ERR_INVALID_ARGUMENT;
  const double th = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (th > 0.0) {
    const double k[3] = {r[0] / th, r[1] / th, r[2] / th};
    const double c = std::cos(th), s = std::sin(th), v = 1.0 - c;
    R[0] = c + k[0] * k[0] * v;
    R[1] = k[0] * k[1] * v - k[2] * s;
    R[2] = k[0] * k[2] * v + k[1] * s;
    R[3] = k[1] * k[0] * v + k[2] * s;
    R[4] = c + k[1] * k[1] * v;
    R[5] = k[1] * k[2] * v - k[0] * s;
    R[6] = k[2] * k[0] * v - k[1] * s;
    R[7] = k[2] * k[1] * v + k[0] * s;
    R[8] = c + k[2] * k[2] * v;
  }
  const double T[16] = {R[0], R[1], R[2], t[0], R[3], R[4], R[5], t[1], R[6], R[7], R[8], t[2], 0, 0, 0, 1};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += T[4 * i + k] * T_axis[4 * k + j];
      T_out[4 * i + j] = s;
    }
  return ILCC2_SUCCESS;
}

// ImageCornersEst::extrinsic2txt (ImageCornersEst.cpp:301-306): the raw Eigen::Matrix4d, i.e. 16 doubles column-major
int ilcc2_extrinsic2bin(const char* path, const double T_rowmajor[16]) {
  if (!path || !T_rowmajor) return ILCC2_ERR_INVALID_ARGUMENT;
  double cm[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) cm[4 * j + i] = T_rowmajor[4 * i + j];
  std::ofstream out(path, std::ios_base::binary | std::ios_base::trunc);
  if (!out.is_open()) return ILCC2_ERR_IO;
  out.write(reinterpret_cast<const char*>(cm), sizeof(cm));
  return out.good() ? ILCC2_SUCCESS : ILCC2_ERR_IO;
}

// ImageCornersEst::read_cam_corners (ImageCornersEst.cpp:213-279): lines of x until >= num values were read, then lines
// of y into the same [row][col]; emitted column-major when the number of rows differs from the board height.
// Returns the number of corners written to out_xy ([cap][2]) or a negative error.
int ilcc2_read_cam_corners(const char* path, int num, int board_h, double* out_xy, int cap) {
  if (!path || !out_xy || num < 1) return -ILCC2_ERR_INVALID_ARGUMENT;
  std::ifstream in(path);
  if (!in.is_open()) return -ILCC2_ERR_IO;
  std::vector<std::vector<std::pair<double, double>>> rows;
  std::string line;
  int counter = 0;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::vector<std::pair<double, double>> row;
    double v;
    while (ss >> v) {
      row.emplace_back(v, 0.0);
      counter++;
    }
    rows.push_back(row);
    if (counter >= num) break;
  }
  size_t k = 0;
  while (std::getline(in, line) && k < rows.size()) {
    std::istringstream ss(line);
    double v;
    for (size_t i = 0; i < rows[k].size() && (ss >> v); ++i) rows[k][i].second = v;
    k++;
  }
  if (rows.empty() || rows[0].empty()) return -ILCC2_ERR_IO;
  int n = 0;
  auto emit = [&](const std::pair<double, double>& p) {
    if (n < cap) {
      out_xy[2 * n] = p.first;
      out_xy[2 * n + 1] = p.second;
    }
    n++;
  };
  if ((int)rows.size() != board_h) {
    for (size_t w = 0; w < rows[0].size(); ++w)
      for (size_t hh = 0; hh < rows.size(); ++hh)
        if (w < rows[hh].size()) emit(rows[hh][w]);
  } else {
    for (auto& rw : rows)
      for (auto& p : rw) emit(p);
  }
  return n;
}

// ImageCornersEst::read_lidar_corners (:281-299): `float x, y, z; infile >> x >> y >> z` widened to double.
int ilcc2_read_lidar_corners(const char* path, int num, double* out_xyz) {
  if (!path || !out_xyz || num < 1) return -ILCC2_ERR_INVALID_ARGUMENT;
  std::ifstream in(path);
  if (!in.is_open()) return -ILCC2_ERR_IO;
  int n = 0;
  while (n < num) {
    float x, y, z;
    if (!(in >> x >> y >> z)) break;
    out_xyz[3 * n] = x;
    out_xyz[3 * n + 1] = y;
    out_xyz[3 * n + 2] = z;
    n++;
  }
  return n;
}

// camK of ImageCornersEst::getRectifyParam (ImageCornersEst.cpp:15-31): K (row-major 3x3) and the board size from the
// OpenCV-YAML camera file.
int ilcc2_read_camera_yaml(const char* path, double K[9], int* corner_in_x, int* corner_in_y) {
  if (!path || !K) return ILCC2_ERR_INVALID_ARGUMENT;
  std::ifstream in(path);
  if (!in.is_open()) return ILCC2_ERR_IO;
  std::stringstream buf;
  buf << in.rdbuf();
  const std::string txt = buf.str();
  size_t pos = 0;
  bool have_k = false;
  while (pos < txt.size()) {  // top-level key "K:" at the start of a line
    const size_t eol = txt.find('\n', pos);
    const std::string line = txt.substr(pos, eol == std::string::npos ? std::string::npos : eol - pos);
    if (line.rfind("K:", 0) == 0) {
      const size_t d = txt.find("data:", pos), lb = txt.find('[', d), rb = txt.find(']', lb);
      if (d == std::string::npos || lb == std::string::npos || rb == std::string::npos) return ILCC2_ERR_IO;
      std::string nums = txt.substr(lb + 1, rb - lb - 1);
      for (auto& c : nums)
        if (c == ',') c = ' ';
      std::istringstream ss(nums);
      int n = 0;
      while (n < 9 && (ss >> K[n])) n++;
      have_k = n == 9;
    }
    const size_t c = line.find(':');
    if (c != std::string::npos) {
      const std::string key = line.substr(0, c), val = line.substr(c + 1);
      if (key == "corner_in_x" && corner_in_x) *corner_in_x = atoi(val.c_str());
      if (key == "corner_in_y" && corner_in_y) *corner_in_y = atoi(val.c_str());
    }
    if (eol == std::string::npos) break;
    pos = eol + 1;
  }
  return have_k ? ILCC2_SUCCESS : ILCC2_ERR_IO;
}

}  // extern "C"
