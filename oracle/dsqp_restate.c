/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 * See dsqp_oracle.h.  Every function cites the reference lines it follows.
 */
#include "dsqp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "osqp_restate.h"

/* ------------------------------------------------------------------ */
/* parameters: config.yaml + motion_planning.cc:54-109 + utils.cc:34-59 */
/* (lives in the oracle too so it does not link the product library)    */
static void oracle_settings(const csdo_params *p, oq_settings *s, int linsys) {
  oq_default_settings(s);
  s->rho = p->rho; s->sigma = p->sigma; s->alpha = p->alpha;
  s->eps_abs = p->eps_abs; s->eps_rel = p->eps_rel;
  s->eps_prim_inf = p->eps_prim_inf; s->eps_dual_inf = p->eps_dual_inf;
  s->adaptive_rho_tolerance = p->adaptive_rho_tolerance;
  s->scaling = p->scaling; s->check_termination = p->check_termination;
  s->adaptive_rho = p->adaptive_rho;
  s->adaptive_rho_interval = p->adaptive_rho_interval;
  s->max_iter = p->osqp_max_iter; /* dsqp_solver.cc:487 */
  s->linsys = linsys;
}

/* ------------------------------------------------------------------ */
/* corridor.cc                                                          */
typedef struct { double x_min, y_min, x_max, y_max; } box_t;

/* Box::ExpandBox, corridor.h:26-49 */
static box_t expand_box(box_t b, int direction, double ds) {
  if (direction == 0) b.y_max += ds;
  else if (direction == 1) b.x_min -= ds;
  else if (direction == 2) b.y_min -= ds;
  else if (direction == 3) b.x_max += ds;
  return b;
}

/* isBoxValid, corridor.cc:252-272 */
static int is_box_valid(const csdo_params *p, box_t box, const double *obs, int n_obs,
                        double dimx, double dimy) {
  double rv = p->rv;
  if (box.x_min < rv || box.x_max > dimx - rv || box.y_min < rv || box.y_max > dimy - rv)
    return 0;
  for (int o = 0; o < n_obs; ++o) {
    box_t d = box;
    for (int i = 0; i < 4; i++) d = expand_box(d, i, obs[3 * o + 2] + rv);
    if (d.x_min < obs[3 * o] && obs[3 * o] < d.x_max && d.y_min < obs[3 * o + 1] &&
        obs[3 * o + 1] < d.y_max)
      return 0;
  }
  return 1;
}

/* generateLocalBox, corridor.cc:278-324 */
static int generate_local_box(const csdo_params *p, double xc, double yc, const double *obs,
                              int n_obs, double dimx, double dimy, box_t *res) {
  int id[4] = {0, 1, 2, 3};
  double lens[4] = {0, 0, 0, 0};
  box_t box = {xc, yc, xc, yc};
  int num_expand = 0, n_id_valid = 4;
  while (n_id_valid > 0) {
    for (int k = 0; k < 4; ++k) {
      int i = id[k];
      if (i == -1) continue;
      box_t trial = expand_box(box, i, p->box_ds);
      if (is_box_valid(p, trial, obs, n_obs, dimx, dimy)) {
        num_expand++;
        lens[i] += p->box_ds;
        box = trial;
        if (lens[i] >= p->box_limit) { n_id_valid -= 1; id[i] = -1; }
      } else {
        n_id_valid -= 1; id[i] = -1;
      }
    }
  }
  *res = box;
  return num_expand > 0;
}

/* isPointCollision, corridor.cc:32-52: first obstacle in container order */
static int is_point_collision(const csdo_params *p, double x, double y, const double *obs,
                              int n_obs) {
  box_t box = {x, y, x, y};
  for (int o = 0; o < n_obs; ++o) {
    box_t d = box;
    for (int i = 0; i < 4; i++) d = expand_box(d, i, obs[3 * o + 2] + p->rv);
    if (d.x_min < obs[3 * o] && obs[3 * o] < d.x_max && d.y_min < obs[3 * o + 1] &&
        obs[3 * o + 1] < d.y_max)
      return o;
  }
  return -1;
}

/* generateLegalPoint, corridor.cc:84-122 */
static int generate_legal_point(const csdo_params *p, const double *oc, const double *obs,
                                int n_obs, double dimx, double dimy, double *x, double *y,
                                box_t *res) {
  int n_cand = 20;
  double x0 = *x, y0 = *y;
  double theta0 = atan2(y0 - oc[1], x0 - oc[0]);
  double d_safe = 0.2;
  double d = p->rv + oc[2] + d_safe;
  for (int i = 0; i < n_cand; i++) {
    int j = i / 2;
    if (i % 2 == 1) j = -j;
    double theta = theta0 + j * 2 * M_PI / n_cand;
    *x = oc[0] + d * cos(theta);
    *y = oc[1] + d * sin(theta);
    if (*x > p->rv && *x < dimx - p->rv && *y > p->rv && *y < dimy - p->rv) {
      box_t b = {0, 0, 0, 0};
      generate_local_box(p, *x, *y, obs, n_obs, dimx, dimy, &b);
      if (is_box_valid(p, b, obs, n_obs, dimx, dimy)) { *res = b; return 1; }
    }
  }
  box_t b = {*x, *y, *x, *y};
  *res = b;
  return 0;
}

/* generateBox, corridor.cc:124-159 (+ isPointOutOfMap :25-30, projectNearBorder :54-81) */
void orc_generate_box(const csdo_params *p, double dimx, double dimy, double x, double y,
                      const double *obs, int n_obs, double *box_out, int *status) {
  double rv = p->rv;
  int success = 0, initial = 0;
  box_t box = {x, y, x, y};
  if (x < rv || x > dimx - rv || y < rv || y > dimy - rv) {
    initial = 1;
    double eps = 1e-3;
    if (x < rv) x = rv + eps;
    else if (x > dimx - rv) x = dimx - rv - eps;
    if (y < rv) y = rv + eps;
    else if (y > dimy - rv) y = dimy - rv - eps;
  }
  int oc = is_point_collision(p, x, y, obs, n_obs);
  if (oc >= 0) {
    initial = 2;
    success = generate_legal_point(p, obs + 3 * oc, obs, n_obs, dimx, dimy, &x, &y, &box);
  } else {
    success = generate_local_box(p, x, y, obs, n_obs, dimx, dimy, &box);
  }
  box_out[0] = box.x_min; box_out[1] = box.y_min; box_out[2] = box.x_max; box_out[3] = box.y_max;
  status[0] = success; status[1] = initial;
}

/* calcCorridors :164-248 (float centres through State, motion_planning.h:113-118,201-206)
 * / updateCorridor dsqp_solver.cc:818-872 (double centres) */
int orc_agent_corridors(const csdo_params *p, int Nt, const double *x, const double *y,
                        const double *yaw, double dimx, double dimy, const double *obs,
                        int n_obs, int double_centres, double *corr, int *box_status) {
  int legal = 1;
  for (int t = 0; t < Nt; ++t) {
    double xf, yf, xr, yr;
    if (double_centres) {
      xf = x[t] + p->f2x * cos(yaw[t]); xr = x[t] + p->r2x * cos(yaw[t]);
      yf = y[t] + p->f2x * sin(yaw[t]); yr = y[t] + p->r2x * sin(yaw[t]);
    } else {
      xf = (double)(float)(x[t] + p->f2x * cos(yaw[t]));
      xr = (double)(float)(x[t] + p->r2x * cos(yaw[t]));
      yf = (double)(float)(y[t] + p->f2x * sin(yaw[t]));
      yr = (double)(float)(y[t] + p->r2x * sin(yaw[t]));
    }
    double bf[4], br[4]; int sf[2], sr[2];
    orc_generate_box(p, dimx, dimy, xf, yf, obs, n_obs, bf, sf);
    orc_generate_box(p, dimx, dimy, xr, yr, obs, n_obs, br, sr);
    if (sf[1] > 0 || sr[1] > 0) legal = 0;
    corr[0 * Nt + t] = bf[0]; corr[1 * Nt + t] = bf[2]; /* xf_min xf_max */
    corr[2 * Nt + t] = bf[1]; corr[3 * Nt + t] = bf[3]; /* yf_min yf_max */
    corr[4 * Nt + t] = br[0]; corr[5 * Nt + t] = br[2];
    corr[6 * Nt + t] = br[1]; corr[7 * Nt + t] = br[3];
    if (box_status) {
      box_status[4 * t + 0] = sf[0]; box_status[4 * t + 1] = sf[1];
      box_status[4 * t + 2] = sr[0]; box_status[4 * t + 3] = sr[1];
    }
  }
  return legal;
}

/* ------------------------------------------------------------------ */
/* inter_agent_cons.cc:12-140 + State (motion_planning.h:113-217)      */
typedef struct { double x, y, yaw; float xc, yc, xf, xr, yf, yr; } state_t;

static state_t make_state(const csdo_params *p, double x, double y, double yaw) {
  state_t s; s.x = x; s.y = y; s.yaw = yaw;
  s.xf = (float)(x + p->f2x * cos(yaw)); s.xr = (float)(x + p->r2x * cos(yaw));
  s.yf = (float)(y + p->f2x * sin(yaw)); s.yr = (float)(y + p->r2x * sin(yaw));
  float LF = (float)p->LF, LB = (float)p->LB;
  float d_center2real = (LF + LB) / 2 - LB;
  s.xc = (float)(x + d_center2real * cos(yaw));
  s.yc = (float)(y + d_center2real * sin(yaw));
  return s;
}
/* State::agentDistance :208-217: float differences, double squares */
static double agent_distance(const state_t *a, const state_t *b) {
  double d, e;
  d = (double)(a->xf - b->xf) * (double)(a->xf - b->xf) + (double)(a->yf - b->yf) * (double)(a->yf - b->yf);
  e = (double)(a->xf - b->xr) * (double)(a->xf - b->xr) + (double)(a->yf - b->yr) * (double)(a->yf - b->yr);
  d = e < d ? e : d;
  e = (double)(a->xr - b->xf) * (double)(a->xr - b->xf) + (double)(a->yr - b->yf) * (double)(a->yr - b->yf);
  d = e < d ? e : d;
  e = (double)(a->xr - b->xr) * (double)(a->xr - b->xr) + (double)(a->yr - b->yr) * (double)(a->yr - b->yr);
  d = e < d ? e : d;
  return sqrt(d);
}
/* State::agentCollision :140-183 (PRCISE_COLLISION branch), all float */
static int agent_collision(const csdo_params *p, const state_t *a, const state_t *o) {
  float length = (float)p->LF + (float)p->LB;
  float width = (float)p->car_width;
  float shift_x = o->xc - a->xc, shift_y = o->yc - a->yc;
  float cos_v = (float)cos(a->yaw), sin_v = (float)sin(a->yaw);
  float cos_o = (float)cos(o->yaw), sin_o = (float)sin(o->yaw);
  float half_l_v = length / 2, half_w_v = width / 2, half_l_o = length / 2, half_w_o = width / 2;
  float dx1 = cos_v * length / 2, dy1 = sin_v * length / 2;
  float dx2 = sin_v * width / 2, dy2 = -cos_v * width / 2;
  float dx3 = cos_o * length / 2, dy3 = sin_o * length / 2;
  float dx4 = sin_o * width / 2, dy4 = -cos_o * width / 2;
  return ((fabsf(shift_x * cos_v + shift_y * sin_v) <=
           fabsf(dx3 * cos_v + dy3 * sin_v) + fabsf(dx4 * cos_v + dy4 * sin_v) + half_l_v) &&
          (fabsf(shift_x * sin_v - shift_y * cos_v) <=
           fabsf(dx3 * sin_v - dy3 * cos_v) + fabsf(dx4 * sin_v - dy4 * cos_v) + half_w_v) &&
          (fabsf(shift_x * cos_o + shift_y * sin_o) <=
           fabsf(dx1 * cos_o + dy1 * sin_o) + fabsf(dx2 * cos_o + dy2 * sin_o) + half_l_o) &&
          (fabsf(shift_x * sin_o - shift_y * cos_o) <=
           fabsf(dx1 * sin_o - dy1 * cos_o) + fabsf(dx2 * sin_o - dy2 * cos_o) + half_w_o));
}
/* calcPerpendicular :54-69 */
static void calc_perpendicular(const csdo_params *p, double x1, double y1, double x2, double y2,
                               double *a, double *b, double *c1, double *c2) {
  double rv = p->rv;
  *a = x2 - x1; *b = y2 - y1;
  double c = (x1 * x1 + y1 * y1 - x2 * x2 - y2 * y2) / 2;
  double d = sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
  *c1 = c + rv * d; *c2 = c - rv * d;
}

int orc_instance_planes(const csdo_params *p, int Na, int Nt, const double *guess,
                        int *plane_cnt, int *plane_t, double *plane_abc,
                        const int *plane_ptr) {
  int legal = 1;
  int *fill = (int *)calloc((size_t)Na, sizeof(int));
  for (int a = 0; a < Na; ++a) plane_cnt[a] = 0;
  double thr = 2 * sqrt(2) * p->r_trust; /* :35 */
  for (int t = 0; t < Nt; ++t)
    for (int i = 0; i < Na - 1; ++i) {
      const double *gi = guess + (size_t)6 * Nt * i;
      state_t si = make_state(p, gi[t], gi[Nt + t], gi[2 * Nt + t]);
      for (int j = i + 1; j < Na; ++j) {
        const double *gj = guess + (size_t)6 * Nt * j;
        state_t sj = make_state(p, gj[t], gj[Nt + t], gj[2 * Nt + t]);
        double d = agent_distance(&si, &sj);
        if (!(d < thr)) continue;
        if (agent_collision(p, &si, &sj)) legal = 0;
        plane_cnt[i]++; plane_cnt[j]++;
        if (!plane_t) continue;
        /* calcEqualInterPlanes :71-140 */
        double xfi = si.xf, yfi = si.yf, xri = si.xr, yri = si.yr;
        double xfj = sj.xf, yfj = sj.yf, xrj = sj.xr, yrj = sj.yr;
        double a_f2f, b_f2f, c_f2f, c_f2f_, a_f2r, b_f2r, c_f2r, c_f2r_;
        double a_r2f, b_r2f, c_r2f, c_r2f_, a_r2r, b_r2r, c_r2r, c_r2r_;
        calc_perpendicular(p, xfi, yfi, xfj, yfj, &a_f2f, &b_f2f, &c_f2f, &c_f2f_);
        calc_perpendicular(p, xfi, yfi, xrj, yrj, &a_f2r, &b_f2r, &c_f2r, &c_f2r_);
        calc_perpendicular(p, xri, yri, xfj, yfj, &a_r2f, &b_r2f, &c_r2f, &c_r2f_);
        calc_perpendicular(p, xri, yri, xrj, yrj, &a_r2r, &b_r2r, &c_r2r, &c_r2r_);
        int ki = plane_ptr[i] + fill[i]++, kj = plane_ptr[j] + fill[j]++;
        double *pi = plane_abc + (size_t)12 * ki, *pj = plane_abc + (size_t)12 * kj;
        plane_t[ki] = t; plane_t[kj] = t;
        pi[0] = a_f2f; pi[1] = b_f2f; pi[2] = c_f2f; pi[3] = a_f2r; pi[4] = b_f2r; pi[5] = c_f2r;
        pi[6] = a_r2f; pi[7] = b_r2f; pi[8] = c_r2f; pi[9] = a_r2r; pi[10] = b_r2r; pi[11] = c_r2r;
        /* plane_j: note f2r <- r2f (:131-135) */
        pj[0] = -a_f2f; pj[1] = -b_f2f; pj[2] = -c_f2f_; pj[3] = -a_r2f; pj[4] = -b_r2f; pj[5] = -c_r2f_;
        pj[6] = -a_f2r; pj[7] = -b_f2r; pj[8] = -c_f2r_; pj[9] = -a_r2r; pj[10] = -b_r2r; pj[11] = -c_r2r_;
      }
    }
  free(fill);
  return legal;
}

/* ------------------------------------------------------------------ */
/* QP assembly: dsqp_solver.cc:103-205 and the six builders :646-1129  */
typedef struct { int r, c; double v; } ent_t;
static int ent_cmp(const void *a, const void *b) {
  const ent_t *x = (const ent_t *)a, *y = (const ent_t *)b;
  if (x->c != y->c) return x->c < y->c ? -1 : 1;
  if (x->r != y->r) return x->r < y->r ? -1 : 1;
  return 0;
}

int orc_assemble_qp(const csdo_params *p, int Nt, const double *lin, const double *trust,
                    const double *cfg, const double *corr, int K, const int *plane_t,
                    const double *plane_abc, int *Ap, int *Ai, double *Ax, double *l,
                    double *u, int *Pp, int *Pi, double *Px) {
  const double *x0 = lin, *y0 = lin + Nt, *yaw0 = lin + 2 * Nt, *steer0 = lin + 3 * Nt,
               *v0 = lin + 4 * Nt, *w0 = lin + 5 * Nt;
  (void)x0; (void)y0; (void)w0;
  int n = 6 * Nt - 2, m = 13 * Nt + 4 * K, Nm = Nt - 1;
  int nnz = 28 * Nt - 11 + 12 * K, cnt = 0;
  ent_t *e = (ent_t *)malloc(sizeof(ent_t) * (size_t)nnz);
  double dt = p->dt, WB = p->WB, f2x = p->f2x, r2x = p->r2x;
  int X = 0, Y = Nt, P = 2 * Nt, S = 3 * Nt, V = 4 * Nt, W = 4 * Nt + Nm;
#define PUSH(R, C, VAL) do { e[cnt].r = (R); e[cnt].c = (C); e[cnt].v = (VAL); cnt++; } while (0)
  for (int i = 0; i < m; ++i) { l[i] = 0; u[i] = 0; }
  /* calcKineConstraint :646-744 */
  for (int t = 0; t < Nm; ++t) {
    double s = sin(yaw0[t]), c = cos(yaw0[t]), cs = cos(steer0[t]);
    PUSH(t, X + t, 1); PUSH(t, X + t + 1, -1);
    PUSH(t, P + t, -dt * (v0[t] * s));                 /* a_yaw1 :670 */
    PUSH(t, V + t, dt * c);                            /* coeff_cos :704 */
    PUSH(Nm + t, Y + t, 1); PUSH(Nm + t, Y + t + 1, -1);
    PUSH(Nm + t, P + t, dt * (v0[t] * c));             /* a_yaw2 :676 */
    PUSH(Nm + t, V + t, dt * s);                       /* coeff_sin :705 */
    PUSH(2 * Nm + t, P + t, 1); PUSH(2 * Nm + t, P + t + 1, -1);
    PUSH(2 * Nm + t, S + t, (dt / WB * v0[t]) / (cs * cs)); /* a_steer :681 */
    PUSH(2 * Nm + t, V + t, dt / WB * tan(steer0[t]));      /* coeff_tan :706 */
    PUSH(3 * Nm + t, S + t, 1); PUSH(3 * Nm + t, S + t + 1, -1);
    PUSH(3 * Nm + t, W + t, dt * 1.0);                 /* eye_dt :707 */
    /* C :717-718; lb = ub = -C :741-742 */
    double C1 = dt * yaw0[t] * v0[t] * s;
    double C2 = -dt * yaw0[t] * v0[t] * c;
    double C3 = -dt * (steer0[t] * v0[t] / WB / (cs * cs));
    l[t] = u[t] = -C1; l[Nm + t] = u[Nm + t] = -C2; l[2 * Nm + t] = u[2 * Nm + t] = -C3;
    l[3 * Nm + t] = u[3 * Nm + t] = -0.0;
  }
  int si = 4 * Nm;
  /* calcCfgConstraint :746-788 */
  PUSH(si + 0, X, 1); PUSH(si + 1, X + Nt - 1, 1); PUSH(si + 2, Y, 1); PUSH(si + 3, Y + Nt - 1, 1);
  PUSH(si + 4, P, 1); PUSH(si + 5, P + Nt - 1, 1);
  for (int k = 0; k < 6; ++k) l[si + k] = u[si + k] = cfg[k];
  si += 6;
  /* calcCorridorConstraint :874-968; D, E reused by the inter rows */
  double *Dc = (double *)malloc(sizeof(double) * (size_t)(4 * Nt));
  double *Eo = (double *)malloc(sizeof(double) * (size_t)(4 * Nt));
  for (int t = 0; t < Nt; ++t) {
    double s = sin(yaw0[t]), c = cos(yaw0[t]);
    Dc[t] = -f2x * s; Dc[Nt + t] = f2x * c; Dc[2 * Nt + t] = -r2x * s; Dc[3 * Nt + t] = r2x * c;
    Eo[t] = f2x * (c + yaw0[t] * s); Eo[Nt + t] = f2x * (s - yaw0[t] * c);
    Eo[2 * Nt + t] = r2x * (c + yaw0[t] * s); Eo[3 * Nt + t] = r2x * (s - yaw0[t] * c);
    PUSH(si + t, X + t, 1); PUSH(si + t, P + t, Dc[t]);
    PUSH(si + Nt + t, Y + t, 1); PUSH(si + Nt + t, P + t, Dc[Nt + t]);
    PUSH(si + 2 * Nt + t, X + t, 1); PUSH(si + 2 * Nt + t, P + t, Dc[2 * Nt + t]);
    PUSH(si + 3 * Nt + t, Y + t, 1); PUSH(si + 3 * Nt + t, P + t, Dc[3 * Nt + t]);
    /* corridor_lbs/ubs order xf,yf,xr,yr (:800-812); Corridor fields via corr planes */
    l[si + t] = corr[0 * Nt + t] - Eo[t];               u[si + t] = corr[1 * Nt + t] - Eo[t];
    l[si + Nt + t] = corr[2 * Nt + t] - Eo[Nt + t];     u[si + Nt + t] = corr[3 * Nt + t] - Eo[Nt + t];
    l[si + 2 * Nt + t] = corr[4 * Nt + t] - Eo[2 * Nt + t]; u[si + 2 * Nt + t] = corr[5 * Nt + t] - Eo[2 * Nt + t];
    l[si + 3 * Nt + t] = corr[6 * Nt + t] - Eo[3 * Nt + t]; u[si + 3 * Nt + t] = corr[7 * Nt + t] - Eo[3 * Nt + t];
  }
  si += 4 * Nt;
  /* calcTrustRegionConstraint :970-994 */
  for (int t = 0; t < Nt; ++t) {
    PUSH(si + t, X + t, 1); PUSH(si + Nt + t, Y + t, 1);
    l[si + t] = -p->r_trust + trust[t]; u[si + t] = p->r_trust + trust[t];
    l[si + Nt + t] = -p->r_trust + trust[Nt + t]; u[si + Nt + t] = p->r_trust + trust[Nt + t];
  }
  si += 2 * Nt;
  /* calcMaxCtrlAndSteerConstraint :996-1039 */
  for (int t = 0; t < Nm; ++t) {
    PUSH(si + t, V + t, 1); PUSH(si + Nm + t, W + t, 1);
    l[si + t] = -p->max_v; u[si + t] = p->max_v;
    l[si + Nm + t] = -p->max_omega; u[si + Nm + t] = p->max_omega;
  }
  for (int t = 0; t < Nt; ++t) {
    PUSH(si + 2 * Nm + t, S + t, 1);
    l[si + 2 * Nm + t] = -p->steer_max; u[si + 2 * Nm + t] = p->steer_max;
  }
  si += 2 * Nm + Nt;
  /* calcInterVehicleConstraint :1097-1129: M_GD = G*D, ub = -(H + G*E), lb = -inf */
  for (int k = 0; k < K; ++k) {
    int t = plane_t[k];
    const double *pl = plane_abc + (size_t)12 * k;
    for (int r = 0; r < 4; ++r) {
      double a = pl[3 * r], b = pl[3 * r + 1], c = pl[3 * r + 2];
      int ox = (r < 2) ? 0 : 2 * Nt, oy = (r < 2) ? Nt : 3 * Nt;
      PUSH(si + 4 * k + r, X + t, a * 1.0);
      PUSH(si + 4 * k + r, Y + t, b * 1.0);
      PUSH(si + 4 * k + r, P + t, a * Dc[ox + t] + b * Dc[oy + t]);
      u[si + 4 * k + r] = -(c + (a * Eo[ox + t] + b * Eo[oy + t]));
      l[si + 4 * k + r] = -INFINITY;
    }
  }
#undef PUSH
  free(Dc); free(Eo);
  if (cnt != nnz) { free(e); return -1; }
  qsort(e, (size_t)cnt, sizeof(ent_t), ent_cmp);
  memset(Ap, 0, sizeof(int) * ((size_t)n + 1));
  for (int k = 0; k < cnt; ++k) { Ai[k] = e[k].r; Ax[k] = e[k].v; Ap[e[k].c + 1]++; }
  for (int j = 0; j < n; ++j) Ap[j + 1] += Ap[j];
  free(e);
  /* objective :163-197, triangularView<Upper> :449 */
  int pc = 0;
  for (int j = 0; j < n; ++j) {
    Pp[j] = pc;
    if (j >= V && j < V + Nm) {
      int t = j - V;
      if (t > 0) { Pi[pc] = j - 1; Px[pc] = -1; pc++; }
      Pi[pc] = j; Px[pc] = (t != 0 && t != Nt - 2) ? 2 : 1; pc++;
    } else if (j >= W) {
      Pi[pc] = j; Px[pc] = 1; pc++;
    }
  }
  Pp[n] = pc;
  return pc;
}

/* ------------------------------------------------------------------ */
int orc_osqp_solve(int n, int m, const int *Pp, const int *Pi, const double *Px,
                   const double *q, const int *Ap, const int *Ai, const double *Ax,
                   const double *l, const double *u, const double *x_warm,
                   const csdo_params *p, int max_iter, int linsys, double *x_out,
                   double *y_out, int *status, int *iters, int *n_factor, double *obj) {
  /* obj: [0] objective, [1] pri_res, [2] dua_res, [3] final rho */
  oq_settings st; oracle_settings(p, &st, linsys);
  st.max_iter = max_iter;
  oq_info info; memset(&info, 0, sizeof(info));
  oq_solve(n, m, Pp, Pi, Px, q, Ap, Ai, Ax, l, u, x_warm, &st, NULL, NULL, x_out, y_out, &info);
  *status = info.status; *iters = info.iter; *n_factor = info.n_factor;
  obj[0] = info.obj_val; obj[1] = info.pri_res; obj[2] = info.dua_res; obj[3] = info.rho;
  return 0;
}

/* ------------------------------------------------------------------ */
/* isFeasible, dsqp_solver.cc:292-420 (fully_check = false)            */
static int is_feasible(const csdo_params *p, int Nt, const double *s, const double *corr,
                       int K, const int *plane_t, const double *plane_abc) {
  const double *x0 = s, *y0 = s + Nt, *yaw0 = s + 2 * Nt, *st0 = s + 3 * Nt, *v0 = s + 4 * Nt,
               *w0 = s + 5 * Nt;
  double dt = p->dt, WB = p->WB;
  double e1 = 0, e2 = 0, e3 = 0, e4 = 0;
  for (int t = 0; t < Nt - 1; ++t) {
    double a = x0[t] + v0[t] * cos(yaw0[t]) * dt - x0[t + 1]; e1 += a * a;
    a = y0[t] + v0[t] * sin(yaw0[t]) * dt - y0[t + 1]; e2 += a * a;
    a = yaw0[t] + v0[t] * tan(st0[t]) / WB * dt - yaw0[t + 1]; e3 += a * a;
    a = st0[t] + w0[t] * dt - st0[t + 1]; e4 += a * a;
  }
  double err_kin = (e1 + e2 + e3 + e4) / Nt;
  if (err_kin > 1e-2) return 0;
  double err_cor_max = 0;
  /* Y = [xf | yf | xr | yr] true disc centres :336-353; lower then upper check */
  for (int pass = 0; pass < 2; ++pass) {
    for (int q = 0; q < 4; ++q)
      for (int t = 0; t < Nt; ++t) {
        double off = (q < 2) ? p->f2x : p->r2x;
        double Yv = (q % 2 == 0) ? x0[t] + off * cos(yaw0[t]) : y0[t] + off * sin(yaw0[t]);
        double lo = corr[(2 * q) * Nt + t], hi = corr[(2 * q + 1) * Nt + t];
        double err = pass == 0 ? lo - Yv : Yv - hi; /* check_less_than(small,big): small>big */
        if (pass == 0 ? !(lo <= Yv) : !(Yv <= hi))
          if (err > err_cor_max) err_cor_max = err;
      }
    if (err_cor_max > 1e-1) return 0;
  }
  double err_inter_max = 0;
  for (int k = 0; k < K; ++k) {
    int t = plane_t[k];
    const double *pl = plane_abc + (size_t)12 * k;
    double xf = x0[t] + p->f2x * cos(yaw0[t]), yf = y0[t] + p->f2x * sin(yaw0[t]);
    double xr = x0[t] + p->r2x * cos(yaw0[t]), yr = y0[t] + p->r2x * sin(yaw0[t]);
    for (int r = 0; r < 4; ++r) {
      double res = (r < 2) ? pl[3 * r] * xf + pl[3 * r + 1] * yf + pl[3 * r + 2]
                           : pl[3 * r] * xr + pl[3 * r + 1] * yr + pl[3 * r + 2];
      if (res > 0 && res > err_inter_max) err_inter_max = res;
    }
  }
  return (err_kin < 1e-2 && err_inter_max < 1e-1 && err_cor_max < 1e-1);
}

/* calcIndividualSQP, dsqp_solver.cc:36-269, one agent */
typedef struct { int status, sqp_iters, admm_iters, n_factor; double obj; long flops; } agent_stat;

static void individual_sqp(const csdo_params *p, int Nt, const double *guess, int K,
                           const int *plane_t, const double *plane_abc, double dimx,
                           double dimy, const double *obs, int n_obs, int linsys,
                           double *corr, double *traj, agent_stat *stat) {
  int n = 6 * Nt - 2, m = 13 * Nt + 4 * K, Nm = Nt - 1;
  int nnzA = 28 * Nt - 11 + 12 * K;
  size_t sz = (size_t)6 * Nt;
  double *lin = (double *)malloc(sizeof(double) * sz);
  double *sol0 = (double *)malloc(sizeof(double) * (size_t)n);
  double *sol = (double *)malloc(sizeof(double) * (size_t)n);
  int *Ap = (int *)malloc(sizeof(int) * ((size_t)n + 1));
  int *Ai = (int *)malloc(sizeof(int) * (size_t)nnzA);
  double *Ax = (double *)malloc(sizeof(double) * (size_t)nnzA);
  double *l = (double *)malloc(sizeof(double) * (size_t)m), *u = (double *)malloc(sizeof(double) * (size_t)m);
  int *Pp = (int *)malloc(sizeof(int) * ((size_t)n + 1)), *Pi = (int *)malloc(sizeof(int) * (size_t)(3 * Nt));
  double *Px = (double *)malloc(sizeof(double) * (size_t)(3 * Nt));
  double *q = (double *)calloc((size_t)n, sizeof(double));
  int *perm_x = (int *)malloc(sizeof(int) * (size_t)n);
  int *perm_k = (int *)malloc(sizeof(int) * (size_t)(n + m));
  /* time-major ordering of the unknowns (any ordering gives the same LDL^T solution) */
  int c = 0;
  for (int t = 0; t < Nt; ++t) {
    perm_x[c++] = t; perm_x[c++] = Nt + t; perm_x[c++] = 2 * Nt + t; perm_x[c++] = 3 * Nt + t;
    if (t < Nm) { perm_x[c++] = 4 * Nt + t; perm_x[c++] = 4 * Nt + Nm + t; }
  }
  for (int i = 0; i < m; ++i) perm_k[i] = n + i;
  for (int j = 0; j < n; ++j) perm_k[m + j] = perm_x[j];
  memcpy(lin, guess, sizeof(double) * sz);
  /* solution0 << x0,y0,yaw0,steer0,v0_,w0_ :63 */
  memcpy(sol0, lin, sizeof(double) * (size_t)(4 * Nt));
  memcpy(sol0 + 4 * Nt, lin + 4 * Nt, sizeof(double) * (size_t)Nm);
  memcpy(sol0 + 4 * Nt + Nm, lin + 5 * Nt, sizeof(double) * (size_t)Nm);
  double cfg[6] = {guess[0], guess[Nt - 1], guess[Nt], guess[2 * Nt - 1], guess[2 * Nt], guess[3 * Nt - 1]};
  oq_settings st; oracle_settings(p, &st, linsys);
  double th = p->delta_solution_threshold, delta = th + 1;
  int iter_count = 0, status = 1;
  stat->admm_iters = 0; stat->n_factor = 0; stat->obj = 0; stat->flops = 0;
  while (delta > th && iter_count < p->max_iter) {
    orc_assemble_qp(p, Nt, lin, guess /* x_trust,y_trust :59-60 */, cfg, corr, K, plane_t,
                    plane_abc, Ap, Ai, Ax, l, u, Pp, Pi, Px);
    oq_info info; memset(&info, 0, sizeof(info));
    oq_solve(n, m, Pp, Pi, Px, q, Ap, Ai, Ax, l, u, sol0, &st, perm_k, perm_x, sol, NULL, &info);
    status = info.status;
    stat->admm_iters += info.iter; stat->n_factor += info.n_factor; stat->flops += info.flops;
    if (abs(status) > 2) memcpy(sol, sol0, sizeof(double) * (size_t)n); /* :518-521 */
    /* objective at the returned point: 1/2 s'Ps */
    {
      double o = 0; const double *v = sol + 4 * Nt, *w = sol + 4 * Nt + Nm;
      for (int t = 0; t + 1 < Nm; ++t) o += 0.5 * (v[t + 1] - v[t]) * (v[t + 1] - v[t]);
      for (int t = 0; t < Nm; ++t) o += 0.5 * w[t] * w[t];
      stat->obj = o;
    }
    delta = 0;
    for (int i = 0; i < n; ++i) delta += (sol[i] - sol0[i]) * (sol[i] - sol0[i]); /* :228 */
    iter_count++;
    /* ExtractAndSimplify :243 */
    memcpy(lin, sol, sizeof(double) * (size_t)(4 * Nt));
    memcpy(lin + 4 * Nt, sol + 4 * Nt, sizeof(double) * (size_t)Nm); lin[5 * Nt - 1] = 0;
    memcpy(lin + 5 * Nt, sol + 4 * Nt + Nm, sizeof(double) * (size_t)Nm); lin[6 * Nt - 1] = 0;
    if (iter_count > p->max_iter / 2 && is_feasible(p, Nt, lin, corr, K, plane_t, plane_abc)) break;
    memcpy(sol0, sol, sizeof(double) * (size_t)n); /* :250 */
    if (!p->fixed_corridor) /* updateCorridor :251-253 */
      orc_agent_corridors(p, Nt, lin, lin + Nt, lin + 2 * Nt, dimx, dimy, obs, n_obs, 1, corr, NULL);
  }
  /* extractSingleSolutionVec2OptRes :577-617 (from solution_vec) */
  memcpy(traj, sol, sizeof(double) * (size_t)(4 * Nt));
  memcpy(traj + 4 * Nt, sol + 4 * Nt, sizeof(double) * (size_t)Nm); traj[5 * Nt - 1] = 0;
  memcpy(traj + 5 * Nt, sol + 4 * Nt + Nm, sizeof(double) * (size_t)Nm); traj[6 * Nt - 1] = 0;
  stat->status = status; stat->sqp_iters = iter_count;
  free(lin); free(sol0); free(sol); free(Ap); free(Ai); free(Ax); free(l); free(u);
  free(Pp); free(Pi); free(Px); free(q); free(perm_x); free(perm_k);
}

/* SolverDSQP::SolverDSQP, dsqp_solver.cc:1133-1248, for every instance */
int orc_refine(const csdo_params *p, const csdo_batch *in, csdo_result *out, int linsys,
               int nthreads, double *flops_out) {
  int A = in->n_agents;
  int *agent_inst = (int *)malloc(sizeof(int) * (size_t)(A > 0 ? A : 1));
  for (int i = 0; i < in->n_inst; ++i) {
    if (in->inst_nt[i] < 3) { free(agent_inst); return CSDO_ERR_INVALID; }
    for (int a = in->inst_agent_ptr[i]; a < in->inst_agent_ptr[i + 1]; ++a) agent_inst[a] = i;
  }
  int *legal = (int *)malloc(sizeof(int) * (size_t)(A > 0 ? A : 1));
  double flops = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : flops)
  for (int a = 0; a < A; ++a) {
    int i = agent_inst[a], Nt = in->inst_nt[i];
    const double *g = in->guess + 6 * in->agent_off[a];
    double *corr = out->corridors + 8 * in->agent_off[a];
    double *traj = out->traj + 6 * in->agent_off[a];
    const double *obs = in->obs + 3 * (size_t)in->obs_ptr[i];
    int n_obs = in->obs_ptr[i + 1] - in->obs_ptr[i];
    double dimx = in->inst_dims[2 * i], dimy = in->inst_dims[2 * i + 1];
    /* calcCorridors :1154 */
    legal[a] = orc_agent_corridors(p, Nt, g, g + Nt, g + 2 * Nt, dimx, dimy, obs, n_obs, 0, corr, NULL);
    int k0 = in->plane_ptr[a], K = in->plane_ptr[a + 1] - k0;
    agent_stat st;
    individual_sqp(p, Nt, g, K, in->plane_t + k0, in->plane_abc + (size_t)12 * k0, dimx, dimy,
                   obs, n_obs, linsys, corr, traj, &st);
    out->status[a] = st.status; out->sqp_iters[a] = st.sqp_iters;
    if (out->n_qp) out->n_qp[a] = st.sqp_iters;
    out->admm_iters[a] = st.admm_iters; out->n_factor[a] = st.n_factor;
    if (out->objective) out->objective[a] = st.obj;
    flops += (double)st.flops;
  }
  /* status aggregation :1224-1243; initial_static_legal :1154 */
  for (int i = 0; i < in->n_inst; ++i) {
    int any = 0, worst = 2, sl = 1;
    for (int a = in->inst_agent_ptr[i]; a < in->inst_agent_ptr[i + 1]; ++a) {
      if (!legal[a]) sl = 0;
      int s = out->status[a];
      if (abs(s) > 1) { any = 1; if (abs(s) > worst) worst = s; }
    }
    out->inst_status[i] = any ? worst : 1;
    out->inst_static_legal[i] = sl;
  }
  if (flops_out) *flops_out = flops;
  free(agent_inst); free(legal);
  return CSDO_OK;
}

/* ------------------------------------------------------------------ */
/* InterpolateInitalGuess chain, inter_agent_cons.cc:143-411           */
static double normalize_angle_abs_in_pi(double x) { /* motion_planning.h:70-75 (returns float!) */
  x = fmod(x + M_PI, 2 * M_PI);
  if (x < 0) x += 2 * M_PI;
  return (double)(float)(x - M_PI);
}

int orc_interpolate_guess(int n_states, const double *states, const int *actions,
                          const double *goal, int n_interp, double dt, double r_const,
                          double LF, double LB, int Nt_out, double *out) {
  int n = n_interp, n_act = n_states - 1;
  int x_size = n_act * (n + 1) + 1;
  if (Nt_out < x_size) return -1;
  double *sx = (double *)malloc(sizeof(double) * (size_t)x_size * 3);
  int *act = (int *)malloc(sizeof(int) * (size_t)(x_size > 1 ? x_size - 1 : 1));
  int ns = 0, na = 0;
  double s[3] = {states[0], states[1], states[2]};
  sx[0] = s[0]; sx[1] = s[1]; sx[2] = s[2]; ns = 1;
  for (int i = 0; i < n_act; ++i) {
    /* action_sample :194-277 */
    int action = actions[i];
    double s0[3] = {s[0], s[1], s[2]};
    double s1[3] = {states[3 * (i + 1)], states[3 * (i + 1) + 1], states[3 * (i + 1) + 2]};
    if (i + 1 == n_states - 1 && goal) { s1[0] = goal[0]; s1[1] = goal[1]; s1[2] = goal[2]; }
    for (int k = 0; k < n + 1; ++k) act[na++] = action;
    if (action == 6) {
      for (int k = 1; k < n + 2; ++k) { sx[3 * ns] = s0[0]; sx[3 * ns + 1] = s0[1]; sx[3 * ns + 2] = s0[2]; ns++; }
    } else {
      double deltat, r = r_const;
      if (action == 0 || action == 3) {
        deltat = sqrt((s1[0] - s0[0]) * (s1[0] - s0[0]) + (s1[1] - s0[1]) * (s1[1] - s0[1])) / r_const;
      } else {
        deltat = normalize_angle_abs_in_pi(s1[2] - s0[2]);
        double d = sqrt((s1[0] - s0[0]) * (s1[0] - s0[0]) + (s1[1] - s0[1]) * (s1[1] - s0[1]));
        r = d / (2.0 * sin(fabs(deltat) / 2.0));
      }
      double da = fabs(deltat) / (double)(n + 1);
      /* calcActionD :172-190 */
      double dxs[6] = {r * da, r * sin(da), r * sin(da), -r * da, -r * sin(da), -r * sin(da)};
      double dys[6] = {0, -r * (1 - cos(da)), r * (1 - cos(da)), 0, -r * (1 - cos(da)), r * (1 - cos(da))};
      double dyw[6] = {0, -da, da, 0, da, -da};
      double dx = dxs[action], dy = dys[action], dyaw = dyw[action];
      double c[3] = {s0[0], s0[1], s0[2]};
      for (int k = 1; k < n + 1; ++k) {
        double xs = c[0] + dx * cos(c[2]) - dy * sin(c[2]);
        double ys = c[1] + dx * sin(c[2]) + dy * cos(c[2]);
        double yw = c[2] + dyaw;
        sx[3 * ns] = xs; sx[3 * ns + 1] = ys; sx[3 * ns + 2] = yw; ns++;
        c[0] = xs; c[1] = ys; c[2] = yw;
      }
      double dy2 = (action == 0 || action == 3) ? 0 : deltat;
      sx[3 * ns] = s1[0]; sx[3 * ns + 1] = s1[1]; sx[3 * ns + 2] = dy2 + s0[2]; ns++;
    }
    s[0] = sx[3 * (ns - 1)]; s[1] = sx[3 * (ns - 1) + 1]; s[2] = sx[3 * (ns - 1) + 2];
  }
  /* calcVSteerW :313-411 */
  int Nt = Nt_out;
  double *X = out, *Y = out + Nt, *YAW = out + 2 * Nt, *ST = out + 3 * Nt, *V = out + 4 * Nt, *W = out + 5 * Nt;
  for (int i = 0; i < ns; ++i) { X[i] = sx[3 * i]; Y[i] = sx[3 * i + 1]; YAW[i] = sx[3 * i + 2]; }
  for (int i = ns; i < Nt; ++i) { X[i] = sx[3 * (ns - 1)]; Y[i] = sx[3 * (ns - 1) + 1]; YAW[i] = sx[3 * (ns - 1) + 2]; }
  ST[0] = 0;
  /* std::atan(float) -> float overload: (LF-LB)/r are all float Constants (:352-353) */
  double phi_action = (double)atanf(((float)LF - (float)LB) / (float)r_const);
  for (int i = 1; i < ns; ++i) {
    int a = act[i - 1];
    ST[i] = (a == 0 || a == 3 || a == 6) ? 0 : ((a == 1 || a == 4) ? -phi_action : phi_action);
  }
  for (int i = ns; i < Nt; ++i) ST[i] = 0;
  for (int i = 0; i < Nt; ++i) { V[i] = 0; W[i] = 0; }
  for (int i = 0; i + 1 < ns; ++i) {
    V[i] = ((X[i + 1] - X[i]) / dt) * cos(YAW[i]) + ((Y[i + 1] - Y[i]) / dt) * sin(YAW[i]);
    W[i] = (ST[i + 1] - ST[i]) / dt;
  }
  free(sx); free(act);
  return ns;
}
