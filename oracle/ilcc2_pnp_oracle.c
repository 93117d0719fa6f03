/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the step that follows the LiDAR corner path in the
 * reference: ordering the 3-D / 2-D corners and the Huber reprojection solve for the LiDAR->camera pose.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this file; the product
 * (lidar_camera_calibration_b200/csrc/ilcc2_pnp.cu) never does.
 *
 * Follows (under /root/reference/ilcc2):
 *   test/calib_lidar_cam.cpp:50-70      get_lidar2cam_axis_roughly
 *   test/calib_lidar_cam.cpp:119-123    T_axis_roughly * p, check_order_lidar, check_order_cam
 *   src/ImageCornersEst.cpp:430-505     check_order_cam / check_order_lidar
 *   include/ilcc2/Optimization.h:126-189  Pose3d2dError (autodiff <2,3,3>)
 *   src/Optimization.cpp:13-91          solvePose3d2dError: HuberLoss(0.1) per 2-D residual block, TRUST_REGION,
 *                                       DENSE_NORMAL_CHOLESKY, DOGLEG + SUBSPACE_DOGLEG, defaults otherwise
 * Third-party arithmetic restated from the published algorithms (not under /root/reference; versions unpinned):
 *   Ceres 1.13/1.14 jet.h (forward-mode duals), rotation.h AngleAxisRotatePoint, HuberLoss + Corrector,
 *   TrustRegionMinimizer, DoglegStrategy (same restatement as the 3-parameter solver in ilcc2_oracle.c, here
 *   for N = 6 parameters in two blocks).  Known deviations: as documented there (2-D boundary problem solved by a
 *   safeguarded Newton iteration; orthonormal basis of span{g, gn} by pivoted Gram-Schmidt instead of
 *   ColPivHouseholderQR -- same subspace).
 * Pinned by: config/pointgrey.bin, which the solve over the six shipped fixture pairs reproduces
 * (tests/test_pnp_oracle.py).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "ilcc2_oracle.h"

#define NP 6

/* ---------------------------------------------------------------------------------------------- */
/* calib_lidar_cam.cpp:50-70                                                                        */
/* ---------------------------------------------------------------------------------------------- */
static void rot_axis(int axis, double ang, double R[9]) {
  const double c = cos(ang), s = sin(ang);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(R, I, sizeof(I));
  if (axis == 0) { R[4] = c; R[5] = -s; R[7] = s; R[8] = c; }
  if (axis == 1) { R[0] = c; R[2] = s; R[6] = -s; R[8] = c; }
  if (axis == 2) { R[0] = c; R[1] = -s; R[3] = s; R[4] = c; }
}

static void mul33(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

/* row-major 4x4; returns 0 for an unknown camera name (the reference leaves the rotation uninitialised then) */
int ilcc2o_axis_roughly(const char* cam, double T[16]) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, A[9], B[9];
  int known = 1;
  if (!strcmp(cam, "front") || !strcmp(cam, "car_left") || !strcmp(cam, "pointgrey")) {
    rot_axis(1, -1.57, A);
    rot_axis(0, 1.57, B);
    mul33(A, B, R);
  } else if (!strcmp(cam, "left")) {
    rot_axis(0, 1.57, R);
  } else if (!strcmp(cam, "right")) {
    rot_axis(0, 1.57, A);
    rot_axis(2, 3.14, B);
    mul33(A, B, R);
  } else if (!strcmp(cam, "back")) {
    rot_axis(1, 1.57, A);
    rot_axis(0, 1.57, B);
    mul33(A, B, R);
  } else {
    known = 0;
  }
  memset(T, 0, 16 * sizeof(double));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
  T[15] = 1.0;
  return known;
}

/* ImageCornersEst.cpp:461-487 on points already multiplied by T_axis_roughly; pts [bw*bh][3] in place */
void ilcc2o_order_lidar(double* p, int bw, int bh) {
  if (p[3 * 0 + 1] > p[3 * (bw + 1) + 1]) {
    for (int h = 0; h < bh / 2; ++h) {
      const int front = bw * h, end = bw * (bh - 1 - h);
      for (int w = 0; w < bw; ++w)
        for (int k = 0; k < 3; ++k) {
          const double t = p[3 * (front + w) + k];
          p[3 * (front + w) + k] = p[3 * (end + w) + k];
          p[3 * (end + w) + k] = t;
        }
    }
  }
  if (p[0] > p[3]) {
    for (int h = 0; h < bh; ++h) {
      const int front = bw * h;
      for (int w = 0; w < bw / 2; ++w)
        for (int k = 0; k < 3; ++k) {
          const double t = p[3 * (front + w) + k];
          p[3 * (front + w) + k] = p[3 * (front + bw - 1 - w) + k];
          p[3 * (front + bw - 1 - w) + k] = t;
        }
    }
  }
}

/* ImageCornersEst.cpp:430-459; pts [bw*bh][2] in place */
void ilcc2o_order_cam(double* p, int bw, int bh) {
  if (p[2 * 0 + 1] > p[2 * (bw + 1) + 1]) {
    for (int h = 0; h < bh / 2; ++h) {
      const int front = bw * h, end = bw * (bh - 1 - h);
      for (int w = 0; w < bw; ++w)
        for (int k = 0; k < 2; ++k) {
          const double t = p[2 * (front + w) + k];
          p[2 * (front + w) + k] = p[2 * (end + w) + k];
          p[2 * (end + w) + k] = t;
        }
    }
  }
  if (p[0] > p[2]) {
    for (int h = 0; h < bh; ++h) {
      const int front = bw * h;
      for (int w = 0; w < bw / 2; ++w)
        for (int k = 0; k < 2; ++k) {
          const double t = p[2 * (front + w) + k];
          p[2 * (front + w) + k] = p[2 * (front + bw - 1 - w) + k];
          p[2 * (front + bw - 1 - w) + k] = t;
        }
    }
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* [UPSTREAM Ceres jet.h] forward-mode dual numbers with NP derivatives                            */
/* ---------------------------------------------------------------------------------------------- */
typedef struct { double a; double v[NP]; } jet;

static jet j_const(double a) { jet r; r.a = a; memset(r.v, 0, sizeof(r.v)); return r; }
static jet j_var(double a, int k) { jet r = j_const(a); r.v[k] = 1.0; return r; }
static jet j_add(jet f, jet g) { jet r; r.a = f.a + g.a; for (int k = 0; k < NP; ++k) r.v[k] = f.v[k] + g.v[k]; return r; }
static jet j_sub(jet f, jet g) { jet r; r.a = f.a - g.a; for (int k = 0; k < NP; ++k) r.v[k] = f.v[k] - g.v[k]; return r; }
static jet j_mul(jet f, jet g) {
  jet r; r.a = f.a * g.a;
  for (int k = 0; k < NP; ++k) r.v[k] = f.a * g.v[k] + f.v[k] * g.a;
  return r;
}
static jet j_div(jet f, jet g) {
  const double gi = 1.0 / g.a, q = f.a * gi;
  jet r; r.a = q;
  for (int k = 0; k < NP; ++k) r.v[k] = (f.v[k] - q * g.v[k]) * gi;
  return r;
}
static jet j_sqrt(jet f) {
  const double t = sqrt(f.a), h = 1.0 / (2.0 * t);
  jet r; r.a = t;
  for (int k = 0; k < NP; ++k) r.v[k] = f.v[k] * h;
  return r;
}
static jet j_cos(jet f) {
  const double c = cos(f.a), s = -sin(f.a);
  jet r; r.a = c;
  for (int k = 0; k < NP; ++k) r.v[k] = s * f.v[k];
  return r;
}
static jet j_sin(jet f) {
  const double s = sin(f.a), c = cos(f.a);
  jet r; r.a = s;
  for (int k = 0; k < NP; ++k) r.v[k] = c * f.v[k];
  return r;
}

/* [UPSTREAM Ceres rotation.h] AngleAxisRotatePoint */
static void angle_axis_rotate(const jet aa[3], const jet pt[3], jet out[3]) {
  const jet theta2 = j_add(j_add(j_mul(aa[0], aa[0]), j_mul(aa[1], aa[1])), j_mul(aa[2], aa[2]));
  if (theta2.a > DBL_EPSILON) {
    const jet theta = j_sqrt(theta2);
    const jet ct = j_cos(theta), st = j_sin(theta);
    const jet ti = j_div(j_const(1.0), theta);
    const jet w[3] = {j_mul(aa[0], ti), j_mul(aa[1], ti), j_mul(aa[2], ti)};
    const jet wxp[3] = {j_sub(j_mul(w[1], pt[2]), j_mul(w[2], pt[1])), j_sub(j_mul(w[2], pt[0]), j_mul(w[0], pt[2])),
                        j_sub(j_mul(w[0], pt[1]), j_mul(w[1], pt[0]))};
    const jet tmp = j_mul(j_add(j_add(j_mul(w[0], pt[0]), j_mul(w[1], pt[1])), j_mul(w[2], pt[2])),
                          j_sub(j_const(1.0), ct));
    for (int i = 0; i < 3; ++i) out[i] = j_add(j_add(j_mul(pt[i], ct), j_mul(wxp[i], st)), j_mul(w[i], tmp));
  } else {
    const jet wxp[3] = {j_sub(j_mul(aa[1], pt[2]), j_mul(aa[2], pt[1])), j_sub(j_mul(aa[2], pt[0]), j_mul(aa[0], pt[2])),
                        j_sub(j_mul(aa[0], pt[1]), j_mul(aa[1], pt[0]))};
    for (int i = 0; i < 3; ++i) out[i] = j_add(pt[i], wxp[i]);
  }
}

/* Optimization.h:155-183: residual (2) and its 2 x 6 Jacobian for one pair; x = (r, t); cam = fx, cx, fy, cy */
void ilcc2o_pnp_residual(const double p3[3], const double p2[2], const double cam[4], const double x[6],
                         double res[2], double jac[12]) {
  jet r[3], t[3], pt[3], pr[3];
  for (int i = 0; i < 3; ++i) {
    r[i] = j_var(x[i], i);
    t[i] = j_var(x[3 + i], 3 + i);
    pt[i] = j_const(p3[i]);
  }
  angle_axis_rotate(r, pt, pr);
  for (int i = 0; i < 3; ++i) pr[i] = j_add(pr[i], t[i]);
  pr[0] = j_div(pr[0], pr[2]);
  pr[1] = j_div(pr[1], pr[2]);
  pr[0] = j_add(j_mul(j_const(cam[0]), pr[0]), j_const(cam[1]));
  pr[1] = j_add(j_mul(j_const(cam[2]), pr[1]), j_const(cam[3]));
  const jet e0 = j_sub(j_const(p2[0]), pr[0]), e1 = j_sub(j_const(p2[1]), pr[1]);
  res[0] = e0.a;
  res[1] = e1.a;
  if (jac)
    for (int k = 0; k < NP; ++k) {
      jac[k] = e0.v[k];
      jac[NP + k] = e1.v[k];
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* cost, g = J^T r, H = J^T J with HuberLoss(a) on the squared norm of each 2-D block.             */
/* [UPSTREAM Ceres] HuberLoss::Evaluate: s <= a^2: rho = (s, 1, 0); else r = sqrt(s):             */
/* rho = (2 a r - a^2, max(DBL_MIN, a / r), -rho' / (2 s)).  Corrector: rho'' <= 0 -> residual and   */
/* Jacobian rows are scaled by sqrt(rho') only.  cost = 1/2 sum rho.                                */
/* ---------------------------------------------------------------------------------------------- */
typedef struct {
  const double* p3;
  const double* p2;
  const uint8_t* valid; /* nullable */
  int m;
  double cam[4];
  double ha, hb;
  int n_evals;
} pnp_problem;

static double pnp_eval(pnp_problem* P, const double x[NP], double* g, double* H) {
  double cost = 0.0;
  if (g) {
    memset(g, 0, NP * sizeof(double));
    memset(H, 0, NP * NP * sizeof(double));
  }
  for (int i = 0; i < P->m; ++i) {
    if (P->valid && !P->valid[i]) continue;
    double res[2], J[12];
    ilcc2o_pnp_residual(P->p3 + 3 * i, P->p2 + 2 * i, P->cam, x, res, g ? J : NULL);
    const double s = res[0] * res[0] + res[1] * res[1];
    double rho0 = s, rho1 = 1.0;
    if (s > P->hb) {
      const double r = sqrt(s);
      rho0 = 2.0 * P->ha * r - P->hb;
      rho1 = P->ha / r;
      if (rho1 < DBL_MIN) rho1 = DBL_MIN;
    }
    cost += 0.5 * rho0;
    if (g) {
      const double sc = sqrt(rho1);
      double rr[2] = {res[0] * sc, res[1] * sc};
      for (int k = 0; k < 12; ++k) J[k] *= sc;
      for (int a = 0; a < NP; ++a) {
        g[a] += J[a] * rr[0] + J[NP + a] * rr[1];
        for (int b = a; b < NP; ++b) H[NP * a + b] += J[a] * J[b] + J[NP + a] * J[NP + b];
      }
    }
  }
  if (g)
    for (int a = 0; a < NP; ++a)
      for (int b = 0; b < a; ++b) H[NP * a + b] = H[NP * b + a];
  P->n_evals++;
  return cost;
}

/* ---------------------------------------------------------------------------------------------- */
/* [UPSTREAM Ceres 1.13/1.14] DoglegStrategy (SUBSPACE_DOGLEG), N = 6, dense normal Cholesky.      */
/* ---------------------------------------------------------------------------------------------- */
typedef struct {
  double radius, mu;
  int reuse;
  double diag[NP], grad[NP], gn[NP], alpha;
  int sub1d;
  double U[NP][2], sub_g[2], sub_B[3];
  double step_norm;
} dogleg6;

static double normN(const double* v) {
  double s = 0.0;
  for (int i = 0; i < NP; ++i) s += v[i] * v[i];
  return sqrt(s);
}
static double dotN(const double* a, const double* b) {
  double s = 0.0;
  for (int i = 0; i < NP; ++i) s += a[i] * b[i];
  return s;
}
static double quadN(const double* H, const double* u, const double* v) { /* u' H v */
  double s = 0.0;
  for (int i = 0; i < NP; ++i) {
    double h = 0.0;
    for (int j = 0; j < NP; ++j) h += H[NP * i + j] * v[j];
    s += u[i] * h;
  }
  return s;
}

static int cholN_solve(const double* A, const double* b, double* x) {
  double L[NP * NP];
  memset(L, 0, sizeof(L));
  for (int j = 0; j < NP; ++j) {
    double t = A[NP * j + j];
    for (int k = 0; k < j; ++k) t -= L[NP * j + k] * L[NP * j + k];
    if (!(t > 0.0)) return 0;
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
    if (!isfinite(x[i])) return 0;
  return 1;
}

static void dl6_traditional(dogleg6* S, double* step) {
  const double gnorm = normN(S->grad), gnn = normN(S->gn);
  if (gnn <= S->radius) {
    for (int i = 0; i < NP; ++i) step[i] = S->gn[i] / S->diag[i];
    S->step_norm = gnn;
    return;
  }
  if (gnorm * S->alpha >= S->radius) {
    for (int i = 0; i < NP; ++i) step[i] = (-(S->radius / gnorm) * S->grad[i]) / S->diag[i];
    S->step_norm = S->radius;
    return;
  }
  const double b_dot_a = -S->alpha * dotN(S->grad, S->gn);
  const double a2 = (S->alpha * gnorm) * (S->alpha * gnorm);
  const double bma2 = a2 - 2.0 * b_dot_a + gnn * gnn;
  const double c = b_dot_a - a2;
  const double d = sqrt(c * c + bma2 * (S->radius * S->radius - a2));
  const double beta = (c <= 0.0) ? (d - c) / bma2 : (S->radius * S->radius - a2) / (d + c);
  double dl[NP];
  for (int i = 0; i < NP; ++i) dl[i] = (-S->alpha * (1.0 - beta)) * S->grad[i] + beta * S->gn[i];
  S->step_norm = normN(dl);
  for (int i = 0; i < NP; ++i) step[i] = dl[i] / S->diag[i];
}

static void dl6_subspace_step(dogleg6* S, double* step) {
  const double gnn = normN(S->gn);
  if (gnn <= S->radius) {
    for (int i = 0; i < NP; ++i) step[i] = S->gn[i] / S->diag[i];
    S->step_norm = gnn;
    return;
  }
  if (S->sub1d) {
    const double gnorm = normN(S->grad);
    for (int i = 0; i < NP; ++i) step[i] = (-(S->radius / gnorm) * S->grad[i]) / S->diag[i];
    S->step_norm = S->radius;
    return;
  }
  double m2[2];
  if (!ilcc2o_tr_boundary_2d(S->sub_B, S->sub_g, S->radius, m2)) {
    dl6_traditional(S, step);
    return;
  }
  for (int i = 0; i < NP; ++i) step[i] = (S->U[i][0] * m2[0] + S->U[i][1] * m2[1]) / S->diag[i];
  S->step_norm = S->radius;
}

static int dl6_compute_step(dogleg6* S, const double* Hs, const double* gs, double* step) {
  if (S->reuse) {
    dl6_subspace_step(S, step);
    return 1;
  }
  S->reuse = 1;
  for (int i = 0; i < NP; ++i) {
    double v = Hs[NP * i + i];
    if (v < 1e-6) v = 1e-6;
    if (v > 1e32) v = 1e32;
    S->diag[i] = sqrt(v);
  }
  for (int i = 0; i < NP; ++i) S->grad[i] = gs[i] / S->diag[i];
  double sg[NP];
  for (int i = 0; i < NP; ++i) sg[i] = S->grad[i] / S->diag[i];
  S->alpha = dotN(S->grad, S->grad) / quadN(Hs, sg, sg);
  int ok = 0;
  double y[NP];
  while (S->mu < 1.0) {
    double A[NP * NP];
    memcpy(A, Hs, sizeof(A));
    const double sm = sqrt(S->mu);
    for (int i = 0; i < NP; ++i) {
      const double l = S->diag[i] * sm;
      A[NP * i + i] += l * l;
    }
    if (cholN_solve(A, gs, y)) {
      ok = 1;
      break;
    }
    S->mu *= 10.0;
  }
  if (!ok) return 0;
  for (int i = 0; i < NP; ++i) S->gn[i] = -S->diag[i] * y[i];
  const double n_g = normN(S->grad), n_n = normN(S->gn);
  const double* A0 = (n_g >= n_n) ? S->grad : S->gn;
  const double* A1 = (n_g >= n_n) ? S->gn : S->grad;
  const double na = (n_g >= n_n) ? n_g : n_n;
  if (!(na > 0.0)) return 0;
  double u0[NP], bp[NP];
  for (int i = 0; i < NP; ++i) u0[i] = A0[i] / na;
  const double pr = dotN(A1, u0);
  for (int i = 0; i < NP; ++i) bp[i] = A1[i] - pr * u0[i];
  const double nb = normN(bp);
  if (!(nb > na * (2.0 * DBL_EPSILON))) {
    S->sub1d = 1;
  } else {
    S->sub1d = 0;
    double w0[NP], w1[NP];
    for (int i = 0; i < NP; ++i) {
      S->U[i][0] = u0[i];
      S->U[i][1] = bp[i] / nb;
    }
    S->sub_g[0] = 0.0;
    S->sub_g[1] = 0.0;
    for (int i = 0; i < NP; ++i) {
      S->sub_g[0] += S->U[i][0] * S->grad[i];
      S->sub_g[1] += S->U[i][1] * S->grad[i];
      w0[i] = S->U[i][0] / S->diag[i];
      w1[i] = S->U[i][1] / S->diag[i];
    }
    S->sub_B[0] = quadN(Hs, w0, w0);
    S->sub_B[1] = quadN(Hs, w0, w1);
    S->sub_B[2] = quadN(Hs, w1, w1);
  }
  dl6_subspace_step(S, step);
  return 1;
}

/* [UPSTREAM Ceres 1.13/1.14] TrustRegionMinimizer with the defaults solvePose3d2dError leaves in place */
int ilcc2o_solve_pnp(const double* p3, const double* p2, const uint8_t* valid, int32_t m, const double cam[4],
                     double huber, int32_t max_iter, double r[3], double t[3], ilcc2o_pnp_summary* out) {
  pnp_problem P = {p3, p2, valid, m, {cam[0], cam[1], cam[2], cam[3]}, huber, huber * huber, 0};
  enum { T_MAX_ITER = 0, T_GRADIENT = 1, T_PARAMETER = 2, T_FUNCTION = 3, T_MIN_RADIUS = 4, T_INVALID = 5 };
  double xc[NP] = {r[0], r[1], r[2], t[0], t[1], t[2]};
  double g[NP], H[NP * NP], scale[NP];
  double x_cost = pnp_eval(&P, xc, g, H);
  const double initial_cost = x_cost;
  for (int i = 0; i < NP; ++i) scale[i] = 1.0 / (1.0 + sqrt(H[NP * i + i]));
  double xnorm = normN(xc), best_cost = x_cost, xbest[NP];
  memcpy(xbest, xc, sizeof(xc));
  dogleg6 S;
  memset(&S, 0, sizeof(S));
  S.radius = 1e4;
  S.mu = 1e-8;
  int iter = 0, invalid = 0, succ = 1, term = T_MAX_ITER;
  double gmax = 0.0;
  for (int i = 0; i < NP; ++i) gmax = fmax(gmax, fabs(g[i]));
  for (;;) {
    if (iter >= max_iter) { term = T_MAX_ITER; break; }
    if (succ && gmax <= 1e-10) { term = T_GRADIENT; break; }
    if (S.radius <= 1e-32) { term = T_MIN_RADIUS; break; }
    iter++;
    double Hs[NP * NP], gs[NP], step[NP];
    for (int i = 0; i < NP; ++i) {
      gs[i] = g[i] * scale[i];
      for (int j = 0; j < NP; ++j) Hs[NP * i + j] = H[NP * i + j] * scale[i] * scale[j];
    }
    const int ok = dl6_compute_step(&S, Hs, gs, step);
    int valid_step = 0;
    double model_change = 0.0;
    if (ok) {
      model_change = -(dotN(step, gs) + 0.5 * quadN(Hs, step, step));
      valid_step = model_change > 0.0;
    }
    if (!valid_step) {
      if (++invalid >= 5) { term = T_INVALID; break; }
      S.mu *= 10.0;
      S.reuse = 0;
      succ = 0;
      continue;
    }
    invalid = 0;
    double cand[NP], dx[NP];
    for (int i = 0; i < NP; ++i) {
      cand[i] = xc[i] + step[i] * scale[i];
      dx[i] = xc[i] - cand[i];
    }
    const double cc = pnp_eval(&P, cand, NULL, NULL);
    if (normN(dx) <= 1e-8 * (xnorm + 1e-8)) { term = T_PARAMETER; break; }
    const double cost_change = x_cost - cc;
    if (fabs(cost_change) <= 1e-6 * x_cost) { term = T_FUNCTION; break; }
    const double quality = cost_change / model_change;
    if (quality > 1e-3) {
      memcpy(xc, cand, sizeof(xc));
      xnorm = normN(xc);
      x_cost = pnp_eval(&P, xc, g, H);
      gmax = 0.0;
      for (int i = 0; i < NP; ++i) gmax = fmax(gmax, fabs(g[i]));
      succ = 1;
      if (quality < 0.25) S.radius *= 0.5;
      if (quality > 0.75) S.radius = fmax(S.radius, 3.0 * S.step_norm);
      S.mu = fmax(1e-8, 2.0 * S.mu / 10.0);
      S.reuse = 0;
      if (x_cost < best_cost) {
        best_cost = x_cost;
        memcpy(xbest, xc, sizeof(xc));
      }
    } else {
      succ = 0;
      S.radius *= 0.5;
      S.reuse = 1;
    }
  }
  for (int i = 0; i < 3; ++i) {
    r[i] = xbest[i];
    t[i] = xbest[3 + i];
  }
  if (out) {
    out->initial_cost = initial_cost;
    out->final_cost = best_cost;
    out->iterations = iter;
    out->termination = term;
    out->n_evals = P.n_evals;
    out->n_pairs = 0;
    for (int i = 0; i < m; ++i) out->n_pairs += (!valid || valid[i]) ? 1 : 0;
  }
  return term;
}

/* Optimization.cpp:75-88 + calib_lidar_cam.cpp:157-160: T = [AngleAxis(|r|, r/|r|) | t] * T_axis_roughly, row-major */
void ilcc2o_compose_extrinsic(const double r[3], const double t[3], const double T_axis[16], double T_out[16]) {
  const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (th > 0.0) {
    const double k[3] = {r[0] / th, r[1] / th, r[2] / th};
    const double c = cos(th), s = sin(th), v = 1.0 - c;
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
  double T[16] = {R[0], R[1], R[2], t[0], R[3], R[4], R[5], t[1], R[6], R[7], R[8], t[2], 0, 0, 0, 1};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += T[4 * i + k] * T_axis[4 * k + j];
      T_out[4 * i + j] = s;
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* ImageCornersEst::spaceToPlane (ImageCornersEst.cpp:135-155), HSVtoRGB (:372-428) and the per-point  */
/* loops of test/pcd2image.cpp:59-84 (image == NULL) / test/rgblidar.cpp:53-73 (image = bgr8).         */
/* pix [n][2] int32 ((-1,-1) = not valid), rgbv [n][4] = r, g, b, valid.                              */
/* ---------------------------------------------------------------------------------------------- */
static void hsv_to_rgb(int h, int s, int v, unsigned char* r, unsigned char* g, unsigned char* b) {
  float RGB_min, RGB_max;
  RGB_max = v * 2.55f;
  RGB_min = RGB_max * (100 - s) / 100.0f;
  const int i = h / 60;
  const int difs = h % 60;
  const float RGB_Adj = (RGB_max - RGB_min) * difs / 60.0f;
  float fr, fg, fb;
  switch (i) {
    case 0: fr = RGB_max; fg = RGB_min + RGB_Adj; fb = RGB_min; break;
    case 1: fr = RGB_max - RGB_Adj; fg = RGB_max; fb = RGB_min; break;
    case 2: fr = RGB_min; fg = RGB_max; fb = RGB_min + RGB_Adj; break;
    case 3: fr = RGB_min; fg = RGB_max - RGB_Adj; fb = RGB_max; break;
    case 4: fr = RGB_min + RGB_Adj; fg = RGB_min; fb = RGB_max; break;
    default: fr = RGB_max; fg = RGB_min; fb = RGB_max - RGB_Adj; break;
  }
  /* float -> unsigned char truncates; out-of-range values are undefined behaviour in the reference (negative hue):
   * defined here as clamping */
  const float v3[3] = {fr, fg, fb};
  unsigned char o[3];
  for (int k = 0; k < 3; ++k) {
    float c = v3[k];
    if (c < 0.f) c = 0.f;
    if (c > 255.f) c = 255.f;
    o[k] = (unsigned char)(int)c;
  }
  *r = o[0]; *g = o[1]; *b = o[2];
}

void ilcc2o_project_points(const void* points, int64_t n, int32_t step, int32_t ox, int32_t oy, int32_t oz, int32_t oi,
                           const double T[16], const double cam[4], int32_t width, int32_t height, double dist,
                           double inten_low, double inten_high, const uint8_t* image_bgr, int32_t* pix, uint8_t* rgbv) {
  const unsigned char* base = (const unsigned char*)points;
  for (int64_t i = 0; i < n; ++i) {
    float x, y, z, in;
    memcpy(&x, base + i * step + ox, 4);
    memcpy(&y, base + i * step + oy, 4);
    memcpy(&z, base + i * step + oz, 4);
    memcpy(&in, base + i * step + oi, 4);
    const double X = x, Y = y, Z = z;
    const double pcx = ((T[0] * X + T[1] * Y) + T[2] * Z) + T[3];
    const double pcy = ((T[4] * X + T[5] * Y) + T[6] * Z) + T[7];
    const double pcz = ((T[8] * X + T[9] * Y) + T[10] * Z) + T[11];
    pix[2 * i] = pix[2 * i + 1] = -1;
    rgbv[4 * i] = rgbv[4 * i + 1] = rgbv[4 * i + 2] = rgbv[4 * i + 3] = 0;
    if (pcz < 0 || pcz > dist) continue;
    const double u = cam[0] * (pcx / pcz) + cam[1], v = cam[2] * (pcy / pcz) + cam[3];
    if (!(u > 0 && u < width && v > 0 && v < height)) continue;
    const int px = (int)u, py = (int)v;
    pix[2 * i] = px;
    pix[2 * i + 1] = py;
    rgbv[4 * i + 3] = 1;
    if (image_bgr) {
      const uint8_t* q = image_bgr + ((size_t)py * width + px) * 3;
      rgbv[4 * i] = q[2];
      rgbv[4 * i + 1] = q[1];
      rgbv[4 * i + 2] = q[0];
    } else {
      const double h = ((double)in - inten_low) / (inten_high - inten_low) * 255;
      hsv_to_rgb(h == h ? (int)h : 0, 100, 100, &rgbv[4 * i], &rgbv[4 * i + 1], &rgbv[4 * i + 2]);
    }
  }
}
