/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 *
 * CPU restatement of the OSQP 0.6.x ADMM that the reference calls at
 * sqp/dsqp_solver.cc:457-549 (osqp_setup / osqp_warm_start_x / osqp_solve /
 * osqp_cleanup).  OSQP is a third-party dependency that is NOT under
 * /root/reference (README.md:16 "osqp version 0.63", CMakeLists.txt:9
 * find_package(osqp REQUIRED)); its source is not available offline, so this
 * file restates its published algorithm (Stellato et al., "OSQP: an operator
 * splitting solver for quadratic programs", and the 0.6.x C sources as
 * recalled).  PARITY UNPINNED: no OSQP binary, golden vector or fixture
 * exists in the reference tree to check this against.
 *
 * One deliberate pin: OSQP's default build chooses adaptive_rho_interval
 * from wall-clock time (settings->adaptive_rho_interval == 0 with PROFILING);
 * here the interval is a parameter (25 = what the time rule yields unless
 * setup costs more than ~95 ADMM iterations).
 */
#ifndef ORACLE_OSQP_RESTATE_H_
#define ORACLE_OSQP_RESTATE_H_

typedef struct oq_settings {
  double rho, sigma, alpha;
  double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  double adaptive_rho_tolerance;
  int scaling, check_termination, adaptive_rho, adaptive_rho_interval;
  int max_iter;
  int linsys; /* 0: LDL^T of the (n+m) KKT; 1: LDL^T of P+sigma I+A'rho A */
} oq_settings;

typedef struct oq_info {
  int status;     /* OSQP status_val */
  int iter;       /* ADMM iterations run */
  int n_factor;   /* numeric factorizations (1 + rho updates) */
  double obj_val; /* unscaled objective at the returned x */
  double pri_res, dua_res, rho;
  long flops;     /* counted floating-point work (factor + solves + matvecs) */
} oq_info;

void oq_default_settings(oq_settings *s);

/*
 * min 1/2 x'Px + q'x  s.t. l <= Ax <= u.
 * P: upper-triangular CSC (n x n), A: CSC (m x n).  x_warm may be NULL (cold).
 * perm_kkt: permutation (new->old) of the n+m KKT unknowns, perm_x: of the n
 * primal unknowns (used by linsys 1); NULL = identity.
 * x_out[n], y_out[m] (may be NULL).  Iterates at a given iteration k are
 * obtained by running with max_iter = k (status -2 keeps the iterate).
 */
int oq_solve(int n, int m, const int *Pp, const int *Pi, const double *Px,
             const double *q, const int *Ap, const int *Ai, const double *Ax,
             const double *l, const double *u, const double *x_warm,
             const oq_settings *st, const int *perm_kkt, const int *perm_x,
             double *x_out, double *y_out, oq_info *info);

#endif
