/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 *
 * Sparse LDL^T of a symmetric quasi-definite matrix, up-looking, elimination
 * tree based: the textbook algorithm that OSQP 0.6.x's default linear solver
 * QDLDL uses (third-party, not under /root/reference; named in reference
 * README.md:16, CMakeLists.txt:9).  OSQP orders the KKT with AMD; here the
 * caller supplies the permutation (any permutation gives the same solution
 * up to rounding).
 */
#ifndef ORACLE_SPARSE_LDL_H_
#define ORACLE_SPARSE_LDL_H_

typedef struct sldl {
  int n;
  int *perm;  /* new -> old, may be NULL */
  int *pinv;  /* old -> new */
  int *Lp, *Li;
  double *Lx, *D, *Dinv;
  int *parent, *Lnz;
  int *flag, *pattern;
  double *Y, *work;
  /* permuted upper CSC pattern of the matrix being factored */
  int *Up, *Ui;
  double *Ux;
  int unz_cap;
  long flops_factor, flops_solve;
} sldl;

/* triplets (i,j,v) in ORIGINAL numbering, either triangle, duplicates summed */
sldl *sldl_new(int n, const int *perm);
void sldl_free(sldl *s);
/* symbolic + numeric; returns 0 ok, <0 zero pivot */
int sldl_factor_triplets(sldl *s, int nz, const int *ti, const int *tj,
                         const double *tv, int redo_symbolic);
/* x <- K^{-1} x (original numbering) */
void sldl_solve(const sldl *s, double *x);

#endif
