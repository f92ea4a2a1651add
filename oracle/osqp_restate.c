/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 * See osqp_restate.h.  Function names in comments are those of OSQP 0.6.x
 * (third-party, un-vendored); the call sites being replaced are reference
 * sqp/dsqp_solver.cc:480 (osqp_set_default_settings), :487 (max_iter),
 * :498 (osqp_setup), :500 (osqp_warm_start_x), :502 (osqp_solve),
 * :511 (info->status_val), :523 (solution->x), :549 (osqp_cleanup).
 */
#include "osqp_restate.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "sparse_ldl.h"

#define RHO_MIN 1e-06
#define RHO_MAX 1e06
#define RHO_EQ_OVER_RHO_INEQ 1e03
#define RHO_TOL 1e-04
#define MIN_SCALING 1e-04
#define MAX_SCALING 1e+04
#define OSQP_INFTY 1e30
#define c_max(a, b) (((a) > (b)) ? (a) : (b))
#define c_min(a, b) (((a) < (b)) ? (a) : (b))
#define c_absval(x) (((x) < 0) ? -(x) : (x))

enum {
  OSQP_SOLVED = 1,
  OSQP_SOLVED_INACCURATE = 2,
  OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3,
  OSQP_DUAL_INFEASIBLE_INACCURATE = 4,
  OSQP_MAX_ITER_REACHED = -2,
  OSQP_PRIMAL_INFEASIBLE = -3,
  OSQP_DUAL_INFEASIBLE = -4,
  OSQP_NON_CVX = -7,
  OSQP_UNSOLVED = -10
};

void oq_default_settings(oq_settings *s) {
  /* osqp_set_default_settings (constants.h of 0.6.x) */
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6;
  s->eps_abs = 1e-3; s->eps_rel = 1e-3;
  s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->adaptive_rho_tolerance = 5.0;
  s->scaling = 10; s->check_termination = 25; s->adaptive_rho = 1;
  s->adaptive_rho_interval = 25; /* pinned, see header */
  s->max_iter = 4000;
  s->linsys = 0;
}

typedef struct work {
  int n, m;
  int *Pp, *Pi; double *Px;
  int *Ap, *Ai; double *Ax_;
  double *q, *l, *u;
  double *D, *Dinv, *E, *Einv; double c, cinv;
  double *rho_vec, *rho_inv_vec; int *constr_type;
  double *x, *y, *z, *xz_tilde, *x_prev, *z_prev;
  double *Ax, *Px_, *Aty, *delta_y, *Atdelta_y, *delta_x, *Pdelta_x, *Adelta_x;
  double *D_temp, *D_temp_A, *E_temp, *sol;
  oq_settings st;
  sldl *lin;
  /* triplet buffers for the linear system */
  int tcap, *ti, *tj; double *tv;
  /* CSR view of A for linsys 1 */
  int *Rp, *Rj; double *Rv;
  double pri_res, dua_res, obj_val; int status;
  long flops;
} work;

/* ---- lin_alg.c of OSQP: same loop orders ---- */
static double vec_norm_inf(const double *v, int l) {
  double mx = 0.0;
  for (int i = 0; i < l; i++) { double a = c_absval(v[i]); if (a > mx) mx = a; }
  return mx;
}
static double vec_scaled_norm_inf(const double *S, const double *v, int l) {
  double mx = 0.0;
  for (int i = 0; i < l; i++) { double a = c_absval(S[i] * v[i]); if (a > mx) mx = a; }
  return mx;
}
static void mat_vec(int ncol, int nrow, const int *p, const int *ix, const double *x,
                    const double *v, double *y, int plus_eq) {
  if (!plus_eq) for (int i = 0; i < nrow; i++) y[i] = 0;
  for (int j = 0; j < ncol; j++)
    for (int k = p[j]; k < p[j + 1]; k++) y[ix[k]] += x[k] * v[j];
}
static void mat_tpose_vec(int ncol, const int *p, const int *ix, const double *x,
                          const double *v, double *y, int plus_eq, int skip_diag) {
  if (!plus_eq) for (int j = 0; j < ncol; j++) y[j] = 0;
  for (int j = 0; j < ncol; j++)
    for (int k = p[j]; k < p[j + 1]; k++) {
      if (skip_diag && ix[k] == j) continue;
      y[j] += x[k] * v[ix[k]];
    }
}
static void mat_inf_norm_cols(int ncol, const int *p, const double *x, double *E) {
  for (int j = 0; j < ncol; j++) {
    E[j] = 0.;
    for (int k = p[j]; k < p[j + 1]; k++) E[j] = c_max(c_absval(x[k]), E[j]);
  }
}
static void mat_inf_norm_rows(int ncol, int nrow, const int *p, const int *ix,
                              const double *x, double *E) {
  for (int i = 0; i < nrow; i++) E[i] = 0.;
  for (int j = 0; j < ncol; j++)
    for (int k = p[j]; k < p[j + 1]; k++) E[ix[k]] = c_max(c_absval(x[k]), E[ix[k]]);
}
static void mat_inf_norm_cols_sym_triu(int n, const int *p, const int *ix,
                                       const double *x, double *E) {
  for (int j = 0; j < n; j++) E[j] = 0.;
  for (int j = 0; j < n; j++)
    for (int k = p[j]; k < p[j + 1]; k++) {
      int i = ix[k];
      double a = c_absval(x[k]);
      E[j] = c_max(a, E[j]);
      if (i != j) E[i] = c_max(a, E[i]);
    }
}
static void limit_scaling(double *D, int n) {
  for (int i = 0; i < n; i++) {
    D[i] = D[i] < MIN_SCALING ? 1.0 : D[i];
    D[i] = D[i] > MAX_SCALING ? MAX_SCALING : D[i];
  }
}

/* scaling.c: scale_data */
static void scale_data(work *w) {
  int n = w->n, m = w->m;
  w->c = 1.0;
  for (int i = 0; i < n; i++) w->D[i] = w->Dinv[i] = 1.0;
  for (int i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1.0;
  for (int it = 0; it < w->st.scaling; it++) {
    /* compute_inf_norm_cols_KKT */
    mat_inf_norm_cols_sym_triu(n, w->Pp, w->Pi, w->Px, w->D_temp);
    mat_inf_norm_cols(n, w->Ap, w->Ax_, w->D_temp_A);
    for (int i = 0; i < n; i++) w->D_temp[i] = c_max(w->D_temp[i], w->D_temp_A[i]);
    mat_inf_norm_rows(n, m, w->Ap, w->Ai, w->Ax_, w->E_temp);
    limit_scaling(w->D_temp, n);
    limit_scaling(w->E_temp, m);
    for (int i = 0; i < n; i++) w->D_temp[i] = 1.0 / sqrt(w->D_temp[i]);
    for (int i = 0; i < m; i++) w->E_temp[i] = 1.0 / sqrt(w->E_temp[i]);
    /* P <- D P D (premult then postmult), A <- E A D */
    for (int j = 0; j < n; j++)
      for (int k = w->Pp[j]; k < w->Pp[j + 1]; k++) w->Px[k] *= w->D_temp[w->Pi[k]];
    for (int j = 0; j < n; j++)
      for (int k = w->Pp[j]; k < w->Pp[j + 1]; k++) w->Px[k] *= w->D_temp[j];
    for (int j = 0; j < n; j++)
      for (int k = w->Ap[j]; k < w->Ap[j + 1]; k++) w->Ax_[k] *= w->E_temp[w->Ai[k]];
    for (int j = 0; j < n; j++)
      for (int k = w->Ap[j]; k < w->Ap[j + 1]; k++) w->Ax_[k] *= w->D_temp[j];
    for (int i = 0; i < n; i++) w->q[i] = w->D_temp[i] * w->q[i];
    for (int i = 0; i < n; i++) w->D[i] = w->D[i] * w->D_temp[i];
    for (int i = 0; i < m; i++) w->E[i] = w->E[i] * w->E_temp[i];
    /* cost normalization */
    mat_inf_norm_cols_sym_triu(n, w->Pp, w->Pi, w->Px, w->D_temp);
    double c_temp = 0.0;
    for (int i = 0; i < n; i++) c_temp += w->D_temp[i];
    c_temp /= (double)n; /* vec_mean */
    double inf_norm_q = vec_norm_inf(w->q, n);
    limit_scaling(&inf_norm_q, 1);
    c_temp = c_max(c_temp, inf_norm_q);
    limit_scaling(&c_temp, 1);
    c_temp = 1. / c_temp;
    for (int k = 0; k < w->Pp[n]; k++) w->Px[k] *= c_temp;
    for (int i = 0; i < n; i++) w->q[i] *= c_temp;
    w->c *= c_temp;
  }
  w->cinv = 1. / w->c;
  for (int i = 0; i < n; i++) w->Dinv[i] = 1.0 / w->D[i];
  for (int i = 0; i < m; i++) w->Einv[i] = 1.0 / w->E[i];
  for (int i = 0; i < m; i++) w->l[i] = w->E[i] * w->l[i];
  for (int i = 0; i < m; i++) w->u[i] = w->E[i] * w->u[i];
}

/* auxil.c: set_rho_vec */
static void set_rho_vec(work *w) {
  w->st.rho = c_min(c_max(w->st.rho, RHO_MIN), RHO_MAX);
  for (int i = 0; i < w->m; i++) {
    if ((w->l[i] < -OSQP_INFTY * MIN_SCALING) && (w->u[i] > OSQP_INFTY * MIN_SCALING)) {
      w->constr_type[i] = -1; w->rho_vec[i] = RHO_MIN;
    } else if (w->u[i] - w->l[i] < RHO_TOL) {
      w->constr_type[i] = 1; w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->st.rho;
    } else {
      w->constr_type[i] = 0; w->rho_vec[i] = w->st.rho;
    }
    w->rho_inv_vec[i] = 1. / w->rho_vec[i];
  }
}

/* linear system: form + factor (init_linsys_solver / update_rho_vec) */
static void tpush(work *w, int *nz, int i, int j, double v) {
  if (*nz >= w->tcap) {
    w->tcap = w->tcap ? 2 * w->tcap : 1024;
    w->ti = (int *)realloc(w->ti, sizeof(int) * (size_t)w->tcap);
    w->tj = (int *)realloc(w->tj, sizeof(int) * (size_t)w->tcap);
    w->tv = (double *)realloc(w->tv, sizeof(double) * (size_t)w->tcap);
  }
  w->ti[*nz] = i; w->tj[*nz] = j; w->tv[*nz] = v; (*nz)++;
}

static int factor_linsys(work *w, int first) {
  int n = w->n, m = w->m, nz = 0;
  /* P + sigma I (upper) */
  for (int j = 0; j < n; j++) {
    int has_diag = 0;
    for (int k = w->Pp[j]; k < w->Pp[j + 1]; k++) {
      int i = w->Pi[k];
      if (i == j) { tpush(w, &nz, i, j, w->Px[k] + w->st.sigma); has_diag = 1; }
      else tpush(w, &nz, i, j, w->Px[k]);
    }
    if (!has_diag) tpush(w, &nz, j, j, w->st.sigma);
  }
  if (w->st.linsys == 0) {
    /* [[P+sigma I, A'],[A, -diag(1/rho)]] */
    for (int j = 0; j < n; j++)
      for (int k = w->Ap[j]; k < w->Ap[j + 1]; k++) tpush(w, &nz, j, n + w->Ai[k], w->Ax_[k]);
    for (int i = 0; i < m; i++) tpush(w, &nz, n + i, n + i, -w->rho_inv_vec[i]);
  } else {
    /* reduced: P + sigma I + A' diag(rho) A  (Schur complement of the block above) */
    for (int i = 0; i < m; i++)
      for (int a = w->Rp[i]; a < w->Rp[i + 1]; a++)
        for (int b = a; b < w->Rp[i + 1]; b++)
          tpush(w, &nz, w->Rj[a], w->Rj[b], w->rho_vec[i] * w->Rv[a] * w->Rv[b]);
    w->flops += 3L * nz;
  }
  int rc = sldl_factor_triplets(w->lin, nz, w->ti, w->tj, w->tv, first);
  w->flops += w->lin->flops_factor;
  return rc;
}

/* update_xz_tilde: compute_rhs + solve_linsys_qdldl */
static void update_xz_tilde(work *w) {
  int n = w->n, m = w->m;
  for (int i = 0; i < n; i++) w->xz_tilde[i] = w->st.sigma * w->x_prev[i] - w->q[i];
  for (int i = 0; i < m; i++) w->xz_tilde[i + n] = w->z_prev[i] - w->rho_inv_vec[i] * w->y[i];
  if (w->st.linsys == 0) {
    double *b = w->xz_tilde;
    /* solve into s->sol, then x_tilde = sol_x, z_tilde = b_z + rho_inv * sol_nu */
    double *tmp = w->sol;
    memcpy(tmp, b, sizeof(double) * (size_t)(n + m));
    sldl_solve(w->lin, tmp);
    for (int j = 0; j < n; j++) b[j] = tmp[j];
    for (int j = 0; j < m; j++) b[j + n] += w->rho_inv_vec[j] * tmp[j + n];
  } else {
    /* eliminate nu: (P+sigma I+A'rho A) x~ = rhs_x + A'(rho .* rhs_z); z~ = A x~ */
    double *b = w->xz_tilde;
    for (int i = 0; i < m; i++) w->Adelta_x[i] = w->rho_vec[i] * b[n + i];
    mat_tpose_vec(n, w->Ap, w->Ai, w->Ax_, w->Adelta_x, b, 1, 0);
    sldl_solve(w->lin, b);
    mat_vec(n, m, w->Ap, w->Ai, w->Ax_, b, b + n, 0);
    w->flops += 4L * w->Ap[n] + m;
  }
  w->flops += w->lin->flops_solve;
}

static void update_x(work *w) {
  for (int i = 0; i < w->n; i++)
    w->x[i] = w->st.alpha * w->xz_tilde[i] + (1.0 - w->st.alpha) * w->x_prev[i];
  for (int i = 0; i < w->n; i++) w->delta_x[i] = w->x[i] - w->x_prev[i];
}
static void update_z(work *w) {
  int n = w->n;
  for (int i = 0; i < w->m; i++)
    w->z[i] = w->st.alpha * w->xz_tilde[i + n] + (1.0 - w->st.alpha) * w->z_prev[i] +
              w->rho_inv_vec[i] * w->y[i];
  for (int i = 0; i < w->m; i++) w->z[i] = c_min(c_max(w->z[i], w->l[i]), w->u[i]); /* project */
}
static void update_y(work *w) {
  int n = w->n;
  for (int i = 0; i < w->m; i++) {
    w->delta_y[i] = w->rho_vec[i] * (w->st.alpha * w->xz_tilde[i + n] +
                                     (1.0 - w->st.alpha) * w->z_prev[i] - w->z[i]);
    w->y[i] += w->delta_y[i];
  }
}

static double compute_obj_val(work *w, const double *x) {
  double quad = 0.0;
  for (int j = 0; j < w->n; j++)
    for (int k = w->Pp[j]; k < w->Pp[j + 1]; k++) {
      int i = w->Pi[k];
      if (i == j) quad += .5 * w->Px[k] * x[i] * x[i];
      else if (i < j) quad += w->Px[k] * x[i] * x[j];
    }
  double lin = 0.0;
  for (int i = 0; i < w->n; i++) lin += w->q[i] * x[i];
  double obj = quad + lin;
  if (w->st.scaling) obj *= w->cinv;
  return obj;
}

/* update_info: compute_pri_res / compute_dua_res (they leave the scaled
 * residual vectors in z_prev / x_prev, which compute_rho_estimate reads) */
static void update_info(work *w) {
  int n = w->n, m = w->m;
  w->obj_val = compute_obj_val(w, w->x);
  if (m == 0) w->pri_res = 0.;
  else {
    mat_vec(n, m, w->Ap, w->Ai, w->Ax_, w->x, w->Ax, 0);
    for (int i = 0; i < m; i++) w->z_prev[i] = w->Ax[i] - w->z[i];
    w->pri_res = w->st.scaling ? vec_scaled_norm_inf(w->Einv, w->z_prev, m)
                               : vec_norm_inf(w->z_prev, m);
  }
  for (int i = 0; i < n; i++) w->x_prev[i] = w->q[i];
  mat_vec(n, n, w->Pp, w->Pi, w->Px, w->x, w->Px_, 0);
  mat_tpose_vec(n, w->Pp, w->Pi, w->Px, w->x, w->Px_, 1, 1);
  for (int i = 0; i < n; i++) w->x_prev[i] = w->x_prev[i] + w->Px_[i];
  if (m > 0) {
    mat_tpose_vec(n, w->Ap, w->Ai, w->Ax_, w->y, w->Aty, 0, 0);
    for (int i = 0; i < n; i++) w->x_prev[i] = w->x_prev[i] + w->Aty[i];
  }
  w->dua_res = w->st.scaling ? w->cinv * vec_scaled_norm_inf(w->Dinv, w->x_prev, n)
                             : vec_norm_inf(w->x_prev, n);
  w->flops += 4L * w->Ap[n] + 4L * w->Pp[n] + 6L * (n + m);
}

static double compute_pri_tol(work *w, double eps_abs, double eps_rel) {
  double a, b;
  if (w->st.scaling) {
    a = vec_scaled_norm_inf(w->Einv, w->z, w->m);
    b = vec_scaled_norm_inf(w->Einv, w->Ax, w->m);
  } else {
    a = vec_norm_inf(w->z, w->m);
    b = vec_norm_inf(w->Ax, w->m);
  }
  return eps_abs + eps_rel * c_max(a, b);
}
static double compute_dua_tol(work *w, double eps_abs, double eps_rel) {
  double mx, t;
  if (w->st.scaling) {
    mx = vec_scaled_norm_inf(w->Dinv, w->q, w->n);
    t = vec_scaled_norm_inf(w->Dinv, w->Aty, w->n); mx = c_max(mx, t);
    t = vec_scaled_norm_inf(w->Dinv, w->Px_, w->n); mx = c_max(mx, t);
    mx *= w->cinv;
  } else {
    mx = vec_norm_inf(w->q, w->n);
    t = vec_norm_inf(w->Aty, w->n); mx = c_max(mx, t);
    t = vec_norm_inf(w->Px_, w->n); mx = c_max(mx, t);
  }
  return eps_abs + eps_rel * mx;
}

/* auxil.c: is_primal_infeasible -- evaluated literally, including the
 * (-inf) * 0 = NaN that a true IEEE -inf lower bound (reference
 * dsqp_solver.cc:1121-1123) produces in the support-function sum. */
static int is_primal_infeasible(work *w, double eps_prim_inf) {
  int m = w->m, n = w->n;
  double norm_delta_y, ineq_lhs = 0.0;
  for (int i = 0; i < m; i++) {
    if (w->u[i] > OSQP_INFTY * MIN_SCALING) {
      if (w->l[i] < -OSQP_INFTY * MIN_SCALING) w->delta_y[i] = 0.0;
      else w->delta_y[i] = c_min(w->delta_y[i], 0.0);
    } else if (w->l[i] < -OSQP_INFTY * MIN_SCALING) {
      w->delta_y[i] = c_max(w->delta_y[i], 0.0);
    }
  }
  if (w->st.scaling) {
    for (int i = 0; i < m; i++) w->Adelta_x[i] = w->E[i] * w->delta_y[i];
    norm_delta_y = vec_norm_inf(w->Adelta_x, m);
  } else norm_delta_y = vec_norm_inf(w->delta_y, m);
  if (norm_delta_y > eps_prim_inf) {
    for (int i = 0; i < m; i++)
      ineq_lhs += w->u[i] * c_max(w->delta_y[i], 0) + w->l[i] * c_min(w->delta_y[i], 0);
    if (ineq_lhs < -eps_prim_inf * norm_delta_y) {
      mat_tpose_vec(n, w->Ap, w->Ai, w->Ax_, w->delta_y, w->Atdelta_y, 0, 0);
      if (w->st.scaling)
        for (int i = 0; i < n; i++) w->Atdelta_y[i] = w->Dinv[i] * w->Atdelta_y[i];
      return vec_norm_inf(w->Atdelta_y, n) < eps_prim_inf * norm_delta_y;
    }
  }
  return 0;
}

/* auxil.c: is_dual_infeasible */
static int is_dual_infeasible(work *w, double eps_dual_inf) {
  int n = w->n, m = w->m;
  double norm_delta_x, cost_scaling;
  if (w->st.scaling) {
    norm_delta_x = vec_scaled_norm_inf(w->D, w->delta_x, n);
    cost_scaling = w->c;
  } else {
    norm_delta_x = vec_norm_inf(w->delta_x, n);
    cost_scaling = 1.0;
  }
  if (norm_delta_x > eps_dual_inf) {
    double qdx = 0.0;
    for (int i = 0; i < n; i++) qdx += w->q[i] * w->delta_x[i];
    if (qdx < -cost_scaling * eps_dual_inf * norm_delta_x) {
      mat_vec(n, n, w->Pp, w->Pi, w->Px, w->delta_x, w->Pdelta_x, 0);
      mat_tpose_vec(n, w->Pp, w->Pi, w->Px, w->delta_x, w->Pdelta_x, 1, 1);
      if (w->st.scaling)
        for (int i = 0; i < n; i++) w->Pdelta_x[i] = w->Dinv[i] * w->Pdelta_x[i];
      if (vec_norm_inf(w->Pdelta_x, n) < cost_scaling * eps_dual_inf * norm_delta_x) {
        mat_vec(n, m, w->Ap, w->Ai, w->Ax_, w->delta_x, w->Adelta_x, 0);
        if (w->st.scaling)
          for (int i = 0; i < m; i++) w->Adelta_x[i] = w->Einv[i] * w->Adelta_x[i];
        for (int i = 0; i < m; i++) {
          if (((w->u[i] < OSQP_INFTY * MIN_SCALING) &&
               (w->Adelta_x[i] > eps_dual_inf * norm_delta_x)) ||
              ((w->l[i] > -OSQP_INFTY * MIN_SCALING) &&
               (w->Adelta_x[i] < -eps_dual_inf * norm_delta_x)))
            return 0;
        }
        return 1;
      }
    }
  }
  return 0;
}

/* auxil.c: check_termination */
static int check_termination(work *w, int approximate) {
  double eps_abs = w->st.eps_abs, eps_rel = w->st.eps_rel;
  double eps_prim_inf = w->st.eps_prim_inf, eps_dual_inf = w->st.eps_dual_inf;
  int prim_res_check = 0, dual_res_check = 0, prim_inf_check = 0, dual_inf_check = 0;
  if ((w->pri_res > OSQP_INFTY) || (w->dua_res > OSQP_INFTY)) {
    w->status = OSQP_NON_CVX;
    return 1;
  }
  if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_prim_inf *= 10; eps_dual_inf *= 10; }
  if (w->m == 0) prim_res_check = 1;
  else {
    double eps_prim = compute_pri_tol(w, eps_abs, eps_rel);
    if (w->pri_res < eps_prim) prim_res_check = 1;
    else prim_inf_check = is_primal_infeasible(w, eps_prim_inf);
  }
  double eps_dual = compute_dua_tol(w, eps_abs, eps_rel);
  if (w->dua_res < eps_dual) dual_res_check = 1;
  else dual_inf_check = is_dual_infeasible(w, eps_dual_inf);
  if (prim_res_check && dual_res_check) {
    w->status = approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED;
    return 1;
  } else if (prim_inf_check) {
    w->status = approximate ? OSQP_PRIMAL_INFEASIBLE_INACCURATE : OSQP_PRIMAL_INFEASIBLE;
    return 1;
  } else if (dual_inf_check) {
    w->status = approximate ? OSQP_DUAL_INFEASIBLE_INACCURATE : OSQP_DUAL_INFEASIBLE;
    return 1;
  }
  return 0;
}

/* auxil.c: compute_rho_estimate / adapt_rho / osqp_update_rho */
static double compute_rho_estimate(work *w) {
  int n = w->n, m = w->m;
  double pri_res = vec_norm_inf(w->z_prev, m);
  double dua_res = vec_norm_inf(w->x_prev, n);
  double pri_res_norm = vec_norm_inf(w->z, m);
  double t = vec_norm_inf(w->Ax, m);
  pri_res_norm = c_max(pri_res_norm, t);
  pri_res /= (pri_res_norm + 1e-10);
  double dua_res_norm = vec_norm_inf(w->q, n);
  t = vec_norm_inf(w->Aty, n); dua_res_norm = c_max(dua_res_norm, t);
  t = vec_norm_inf(w->Px_, n); dua_res_norm = c_max(dua_res_norm, t);
  dua_res /= (dua_res_norm + 1e-10);
  double rho_estimate = w->st.rho * sqrt(pri_res / (dua_res + 1e-10));
  rho_estimate = c_min(c_max(rho_estimate, RHO_MIN), RHO_MAX);
  return rho_estimate;
}
static int adapt_rho(work *w, int *n_factor) {
  double rho_new = compute_rho_estimate(w);
  if ((rho_new > w->st.rho * w->st.adaptive_rho_tolerance) ||
      (rho_new < w->st.rho / w->st.adaptive_rho_tolerance)) {
    w->st.rho = c_min(c_max(rho_new, RHO_MIN), RHO_MAX);
    for (int i = 0; i < w->m; i++) {
      if (w->constr_type[i] == 0) {
        w->rho_vec[i] = w->st.rho; w->rho_inv_vec[i] = 1. / w->st.rho;
      } else if (w->constr_type[i] == 1) {
        w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->st.rho;
        w->rho_inv_vec[i] = 1. / w->rho_vec[i];
      }
    }
    (*n_factor)++;
    return factor_linsys(w, 0);
  }
  return 0;
}

static double *dcopy(const double *s, int n) {
  double *d = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (n > 0) memcpy(d, s, sizeof(double) * (size_t)n);
  return d;
}
static int *icopy(const int *s, int n) {
  int *d = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  if (n > 0) memcpy(d, s, sizeof(int) * (size_t)n);
  return d;
}
static double *dzero(int n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

int oq_solve(int n, int m, const int *Pp, const int *Pi, const double *Px,
             const double *q, const int *Ap, const int *Ai, const double *Ax,
             const double *l, const double *u, const double *x_warm,
             const oq_settings *st, const int *perm_kkt, const int *perm_x,
             double *x_out, double *y_out, oq_info *info) {
  work W; memset(&W, 0, sizeof(W));
  work *w = &W;
  w->n = n; w->m = m; w->st = *st;
  /* validate_data: l <= u */
  for (int i = 0; i < m; i++)
    if (l[i] > u[i]) { info->status = -1000; info->iter = 0; info->n_factor = 0; return 1; }
  /* osqp_setup: copy data */
  w->Pp = icopy(Pp, n + 1); w->Pi = icopy(Pi, Pp[n]); w->Px = dcopy(Px, Pp[n]);
  w->Ap = icopy(Ap, n + 1); w->Ai = icopy(Ai, Ap[n]); w->Ax_ = dcopy(Ax, Ap[n]);
  w->q = dcopy(q, n); w->l = dcopy(l, m); w->u = dcopy(u, m);
  w->D = dzero(n); w->Dinv = dzero(n); w->E = dzero(m); w->Einv = dzero(m);
  w->rho_vec = dzero(m); w->rho_inv_vec = dzero(m);
  w->constr_type = (int *)calloc((size_t)(m > 0 ? m : 1), sizeof(int));
  w->x = dzero(n); w->y = dzero(m); w->z = dzero(m); w->xz_tilde = dzero(n + m);
  w->x_prev = dzero(n); w->z_prev = dzero(m);
  w->Ax = dzero(m); w->Px_ = dzero(n); w->Aty = dzero(n); w->delta_y = dzero(m);
  w->Atdelta_y = dzero(n); w->delta_x = dzero(n); w->Pdelta_x = dzero(n);
  w->Adelta_x = dzero(m);
  w->D_temp = dzero(n); w->D_temp_A = dzero(n); w->E_temp = dzero(m); w->sol = dzero(n + m);
  if (w->st.scaling) scale_data(w);
  else { w->c = w->cinv = 1.0; for (int i = 0; i < n; i++) w->D[i] = w->Dinv[i] = 1.0;
         for (int i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1.0; }
  set_rho_vec(w);
  if (w->st.linsys == 1) {
    /* CSR of the scaled A */
    w->Rp = (int *)calloc((size_t)m + 1, sizeof(int));
    w->Rj = (int *)malloc(sizeof(int) * (size_t)(Ap[n] > 0 ? Ap[n] : 1));
    w->Rv = (double *)malloc(sizeof(double) * (size_t)(Ap[n] > 0 ? Ap[n] : 1));
    for (int k = 0; k < Ap[n]; k++) w->Rp[w->Ai[k] + 1]++;
    for (int i = 0; i < m; i++) w->Rp[i + 1] += w->Rp[i];
    int *fill = icopy(w->Rp, m + 1);
    for (int j = 0; j < n; j++)
      for (int k = w->Ap[j]; k < w->Ap[j + 1]; k++) {
        int p = fill[w->Ai[k]]++;
        w->Rj[p] = j; w->Rv[p] = w->Ax_[k];
      }
    free(fill);
    w->lin = sldl_new(n, perm_x);
  } else {
    w->lin = sldl_new(n + m, perm_kkt);
  }
  int n_factor = 1, rc = factor_linsys(w, 1);
  w->status = OSQP_UNSOLVED;
  int iter = 0, can_check_termination = 0;
  if (rc == 0) {
    /* osqp_warm_start_x: x <- Dinv x0, z <- A x, y stays 0 */
    if (x_warm) {
      for (int i = 0; i < n; i++) w->x[i] = x_warm[i];
      if (w->st.scaling) for (int i = 0; i < n; i++) w->x[i] = w->x[i] * w->Dinv[i];
      mat_vec(n, m, w->Ap, w->Ai, w->Ax_, w->x, w->z, 0);
    }
    /* osqp_solve main loop */
    for (iter = 1; iter <= w->st.max_iter; iter++) {
      double *t;
      t = w->x; w->x = w->x_prev; w->x_prev = t;
      t = w->z; w->z = w->z_prev; w->z_prev = t;
      update_xz_tilde(w);
      update_x(w);
      update_z(w);
      update_y(w);
      w->flops += 4L * n + 12L * m;
      can_check_termination = w->st.check_termination && (iter % w->st.check_termination == 0);
      if (can_check_termination) {
        update_info(w);
        if (check_termination(w, 0)) break;
      }
      if (w->st.adaptive_rho && w->st.adaptive_rho_interval &&
          (iter % w->st.adaptive_rho_interval == 0)) {
        if (!can_check_termination) update_info(w);
        if (adapt_rho(w, &n_factor)) { rc = -1; break; }
      }
    }
    if (iter > w->st.max_iter) iter = w->st.max_iter; /* loop ran to completion */
    if (!can_check_termination) {
      update_info(w);
      check_termination(w, 0);
    }
    if (w->status == OSQP_UNSOLVED) {
      if (!check_termination(w, 1)) w->status = OSQP_MAX_ITER_REACHED;
    }
  }
  /* store_solution */
  int has_solution = !(w->status == OSQP_PRIMAL_INFEASIBLE ||
                       w->status == OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
                       w->status == OSQP_DUAL_INFEASIBLE ||
                       w->status == OSQP_DUAL_INFEASIBLE_INACCURATE ||
                       w->status == OSQP_NON_CVX);
  if (has_solution && rc == 0) {
    w->obj_val = compute_obj_val(w, w->x);
    for (int i = 0; i < n; i++) x_out[i] = w->st.scaling ? w->D[i] * w->x[i] : w->x[i];
    if (y_out)
      for (int i = 0; i < m; i++) y_out[i] = w->st.scaling ? w->cinv * w->E[i] * w->y[i] : w->y[i];
  } else {
    for (int i = 0; i < n; i++) x_out[i] = NAN;
    if (y_out) for (int i = 0; i < m; i++) y_out[i] = NAN;
  }
  info->status = rc == 0 ? w->status : -1001;
  info->iter = iter; info->n_factor = n_factor;
  info->obj_val = w->obj_val; info->pri_res = w->pri_res; info->dua_res = w->dua_res;
  info->rho = w->st.rho; info->flops = w->flops;
  free(w->Pp); free(w->Pi); free(w->Px); free(w->Ap); free(w->Ai); free(w->Ax_);
  free(w->q); free(w->l); free(w->u); free(w->D); free(w->Dinv); free(w->E); free(w->Einv);
  free(w->rho_vec); free(w->rho_inv_vec); free(w->constr_type);
  free(w->x); free(w->y); free(w->z); free(w->xz_tilde); free(w->x_prev); free(w->z_prev);
  free(w->Ax); free(w->Px_); free(w->Aty); free(w->delta_y); free(w->Atdelta_y);
  free(w->delta_x); free(w->Pdelta_x); free(w->Adelta_x);
  free(w->D_temp); free(w->D_temp_A); free(w->E_temp); free(w->sol);
  free(w->ti); free(w->tj); free(w->tv); free(w->Rp); free(w->Rj); free(w->Rv);
  sldl_free(w->lin);
  return 0;
}
