/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 * See sparse_ldl.h.  Up-looking sparse LDL^T (elimination tree + row
 * patterns), written from the published algorithm (Davis, "Algorithm 849: a
 * concise sparse Cholesky factorization package"); QDLDL (OSQP 0.6.x linear
 * solver, called through osqp_setup / osqp_solve at reference
 * sqp/dsqp_solver.cc:498-502) implements the same recurrences.
 */
#include "sparse_ldl.h"

#include <stdlib.h>
#include <string.h>

typedef struct {
  int i, j;
  double v;
} trip;

static int trip_cmp(const void *a, const void *b) {
  const trip *x = (const trip *)a, *y = (const trip *)b;
  if (x->j != y->j) return x->j < y->j ? -1 : 1;
  if (x->i != y->i) return x->i < y->i ? -1 : 1;
  return 0;
}

sldl *sldl_new(int n, const int *perm) {
  sldl *s = (sldl *)calloc(1, sizeof(sldl));
  s->n = n;
  s->pinv = (int *)malloc(sizeof(int) * (size_t)n);
  s->perm = (int *)malloc(sizeof(int) * (size_t)n);
  for (int k = 0; k < n; ++k) {
    s->perm[k] = perm ? perm[k] : k;
    s->pinv[s->perm[k]] = k;
  }
  s->Lp = (int *)calloc((size_t)n + 1, sizeof(int));
  s->D = (double *)calloc((size_t)n, sizeof(double));
  s->Dinv = (double *)calloc((size_t)n, sizeof(double));
  s->parent = (int *)malloc(sizeof(int) * (size_t)n);
  s->Lnz = (int *)malloc(sizeof(int) * (size_t)n);
  s->flag = (int *)malloc(sizeof(int) * (size_t)n);
  s->pattern = (int *)malloc(sizeof(int) * (size_t)n);
  s->Y = (double *)calloc((size_t)n, sizeof(double));
  s->work = (double *)calloc((size_t)n, sizeof(double));
  s->Up = (int *)calloc((size_t)n + 1, sizeof(int));
  return s;
}

void sldl_free(sldl *s) {
  if (!s) return;
  free(s->perm); free(s->pinv); free(s->Lp); free(s->Li); free(s->Lx);
  free(s->D); free(s->Dinv); free(s->parent); free(s->Lnz); free(s->flag);
  free(s->pattern); free(s->Y); free(s->work); free(s->Up); free(s->Ui);
  free(s->Ux);
  free(s);
}

/* permuted, upper-triangular, column-compressed, duplicates summed */
static void build_upper(sldl *s, int nz, const int *ti, const int *tj,
                        const double *tv) {
  trip *t = (trip *)malloc(sizeof(trip) * (size_t)(nz > 0 ? nz : 1));
  for (int k = 0; k < nz; ++k) {
    int a = s->pinv[ti[k]], b = s->pinv[tj[k]];
    if (a > b) { int c = a; a = b; b = c; }
    t[k].i = a; t[k].j = b; t[k].v = tv[k];
  }
  qsort(t, (size_t)nz, sizeof(trip), trip_cmp);
  if (nz > s->unz_cap) {
    free(s->Ui); free(s->Ux);
    s->Ui = (int *)malloc(sizeof(int) * (size_t)nz);
    s->Ux = (double *)malloc(sizeof(double) * (size_t)nz);
    s->unz_cap = nz;
  }
  int n = s->n, cnt = 0;
  memset(s->Up, 0, sizeof(int) * ((size_t)n + 1));
  for (int k = 0; k < nz; ++k) {
    if (cnt > 0 && s->Ui[cnt - 1] == t[k].i && k > 0 && t[k - 1].j == t[k].j) {
      s->Ux[cnt - 1] += t[k].v;
    } else {
      s->Ui[cnt] = t[k].i; s->Ux[cnt] = t[k].v;
      s->Up[t[k].j + 1]++;
      cnt++;
    }
  }
  for (int j = 0; j < n; ++j) s->Up[j + 1] += s->Up[j];
  free(t);
}

static void symbolic(sldl *s) {
  int n = s->n;
  for (int k = 0; k < n; ++k) {
    s->parent[k] = -1; s->flag[k] = k; s->Lnz[k] = 0;
    for (int p = s->Up[k]; p < s->Up[k + 1]; ++p) {
      int i = s->Ui[p];
      if (i >= k) continue;
      for (; s->flag[i] != k; i = s->parent[i]) {
        if (s->parent[i] == -1) s->parent[i] = k;
        s->Lnz[i]++;
        s->flag[i] = k;
      }
    }
  }
  s->Lp[0] = 0;
  for (int k = 0; k < n; ++k) s->Lp[k + 1] = s->Lp[k] + s->Lnz[k];
  free(s->Li); free(s->Lx);
  int lnz = s->Lp[n] > 0 ? s->Lp[n] : 1;
  s->Li = (int *)malloc(sizeof(int) * (size_t)lnz);
  s->Lx = (double *)malloc(sizeof(double) * (size_t)lnz);
}

static int numeric(sldl *s) {
  int n = s->n;
  long fl = 0;
  for (int k = 0; k < n; ++k) {
    int top = n;
    s->Y[k] = 0.0; s->flag[k] = k; s->Lnz[k] = 0;
    for (int p = s->Up[k]; p < s->Up[k + 1]; ++p) {
      int i = s->Ui[p];
      if (i > k) continue;
      s->Y[i] += s->Ux[p];
      int len = 0;
      for (; s->flag[i] != k; i = s->parent[i]) {
        s->pattern[len++] = i;
        s->flag[i] = k;
      }
      while (len > 0) s->pattern[--top] = s->pattern[--len];
    }
    s->D[k] = s->Y[k];
    s->Y[k] = 0.0;
    for (; top < n; ++top) {
      int i = s->pattern[top];
      double yi = s->Y[i];
      s->Y[i] = 0.0;
      int p2 = s->Lp[i] + s->Lnz[i];
      for (int p = s->Lp[i]; p < p2; ++p) s->Y[s->Li[p]] -= s->Lx[p] * yi;
      fl += 2L * (p2 - s->Lp[i]) + 3;
      double lki = yi / s->D[i];
      s->D[k] -= lki * yi;
      s->Li[p2] = k;
      s->Lx[p2] = lki;
      s->Lnz[i]++;
    }
    if (s->D[k] == 0.0) return -(k + 1);
    s->Dinv[k] = 1.0 / s->D[k];
  }
  s->flops_factor = fl;
  s->flops_solve = 4L * s->Lp[n] + n;
  return 0;
}

int sldl_factor_triplets(sldl *s, int nz, const int *ti, const int *tj,
                         const double *tv, int redo_symbolic) {
  build_upper(s, nz, ti, tj, tv);
  if (redo_symbolic || !s->Li) symbolic(s);
  return numeric(s);
}

void sldl_solve(const sldl *s, double *x) {
  int n = s->n;
  double *w = s->work;
  for (int k = 0; k < n; ++k) w[k] = x[s->perm[k]];
  for (int j = 0; j < n; ++j) {
    double wj = w[j];
    for (int p = s->Lp[j]; p < s->Lp[j + 1]; ++p) w[s->Li[p]] -= s->Lx[p] * wj;
  }
  for (int j = 0; j < n; ++j) w[j] *= s->Dinv[j];
  for (int j = n - 1; j >= 0; --j) {
    double wj = w[j];
    for (int p = s->Lp[j]; p < s->Lp[j + 1]; ++p) wj -= s->Lx[p] * w[s->Li[p]];
    w[j] = wj;
  }
  for (int k = 0; k < n; ++k) x[s->perm[k]] = w[k];
}
