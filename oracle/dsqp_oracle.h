/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 *
 * CPU restatement of the reference's DSQP refine stage and its feeders:
 *   sqp/dsqp_solver.cc (all), sqp/corridor.cc (all),
 *   sqp/inter_agent_cons.cc:12-140 (pairs + planes), :143-411 (x0_bar),
 *   sqp/utils.cc:93-123, common/motion_planning.h:113-217.
 * The QP step goes through osqp_restate.c (OSQP 0.6.x restated; PARITY
 * UNPINNED -- the reference tree holds no golden vector, test or fixture for
 * this path and neither the reference nor OSQP can be built offline).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef DSQP_ORACLE_H_
#define DSQP_ORACLE_H_

#include "../include/csdo_dsqp.h" /* boundary PODs only: csdo_params, csdo_batch, csdo_result */

#ifdef __cplusplus
extern "C" {
#endif

/* generateBox, sqp/corridor.cc:124-159.  box = {x_min,y_min,x_max,y_max};
 * status = {success, initial_status}. */
void orc_generate_box(const csdo_params *p, double dimx, double dimy, double x,
                      double y, const double *obs, int n_obs, double *box,
                      int *status);

/* calcCorridors (double_centres=0, float disc centres via State) or the box
 * regeneration of updateCorridor (double_centres=1) for one agent.
 * xyyaw: 3 planes of Nt; corr: 8 planes of Nt; box_status [2*Nt][2] or NULL.
 * returns 1 if every initial_status == 0. */
int orc_agent_corridors(const csdo_params *p, int Nt, const double *x,
                        const double *y, const double *yaw, double dimx,
                        double dimy, const double *obs, int n_obs,
                        int double_centres, double *corr, int *box_status);

/* findNeighborPairsByTrustRegion + calcEqualInterPlanes for one instance.
 * guess: Na agents, each 6 planes of Nt (x,y,yaw used).  Pass plane_t==NULL
 * to count only.  plane_cnt[Na] out.  Returns initial_inter_legal. */
int orc_instance_planes(const csdo_params *p, int Na, int Nt,
                        const double *guess, int *plane_cnt, int *plane_t,
                        double *plane_abc, const int *plane_ptr);

/* Assemble one agent QP exactly like calcIndividualSQP's body
 * (dsqp_solver.cc:103-205) in the reference's variable/row order.
 * lin: 6 planes of Nt (linearization point), trust: x[Nt],y[Nt] planes,
 * cfg[6], corr 8 planes.  Sizes: n=6Nt-2, m=13Nt+4K,
 * nnzA=28Nt-11+12K, nnzP=5(Nt-1)-... (returned).  Caller allocates
 * Ap[n+1], Ai/Ax[nnzA], l/u[m], Pp[n+1], Pi/Px[3Nt]. */
int orc_assemble_qp(const csdo_params *p, int Nt, const double *lin,
                    const double *trust, const double *cfg, const double *corr,
                    int K, const int *plane_t, const double *plane_abc,
                    int *Ap, int *Ai, double *Ax, double *l, double *u,
                    int *Pp, int *Pi, double *Px);

/* One generic QP through the OSQP restatement (for known-answer tests).
 * obj[4]: objective, pri_res, dua_res, final rho. */
int orc_osqp_solve(int n, int m, const int *Pp, const int *Pi, const double *Px,
                   const double *q, const int *Ap, const int *Ai,
                   const double *Ax, const double *l, const double *u,
                   const double *x_warm, const csdo_params *p, int max_iter,
                   int linsys, double *x_out, double *y_out, int *status,
                   int *iters, int *n_factor, double *obj);

/* SolverDSQP::SolverDSQP for a whole batch.  linsys 0: (n+m) KKT LDL^T,
 * 1: reduced banded system.  nthreads: OpenMP threads over agents.
 * flops_out (may be NULL): counted floating-point work. */
int orc_refine(const csdo_params *p, const csdo_batch *in, csdo_result *out,
               int linsys, int nthreads, double *flops_out);

/* InterpolateInitalGuess for one agent path (inter_agent_cons.cc:143-411).
 * states [n_states][3], actions [n_states-1]; goal[3] replaces the last
 * state (:149-151).  out: 6 planes of Nt_out >= (n_states-1)*(n_interp+1)+1,
 * padded by holding the last pose (:341-345).  r, LF, LB: Constants. */
int orc_interpolate_guess(int n_states, const double *states,
                          const int *actions, const double *goal, int n_interp,
                          double dt, double r, double LF, double LB, int Nt_out,
                          double *out);

#ifdef __cplusplus
}
#endif
#endif
