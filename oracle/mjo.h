/* mjo.h — CPU ORACLE (test infrastructure, not product code).
 *
 * Plain-C fp64 restatement of the part of MuJoCo 2.x that the reference's hot path calls
 * (mj_forward / mj_forwardSkip / mj_step / mju_quatIntegrate; call sites
 * /root/reference/src/mjderivative.cpp:64,68,92,103,124,134,168,178,191,198 and
 * /root/reference/inc/ilqr.h:86,128), of the reference's FD schedule
 * (/root/reference/src/mjderivative.cpp:43-255) and of its A/B assembly and Riccati recursion
 * (/root/reference/inc/differentiator.h:52-93, /root/reference/inc/ilqr.h:69-186).
 *
 * MuJoCo itself (un-vendored, un-pinned; API use brackets 2.1.2 <= v < 3.0, first written
 * against 2.0.0) is absent from /root/reference and from this image, and the reference has no
 * golden vectors: PARITY WITH UPSTREAM MUJOCO IS UNPINNED.  What pins this oracle instead:
 * closed-form cart-pole dynamics, energy conservation, compiler-vs-CRBA mass matrix cross-checks
 * and KKT residuals of the contact solve; everything re-derived from the TEXT of the MJCF files and
 * textbook mechanics — body mass / centre of mass / inertia by quadrature, forward kinematics, the
 * mass matrix as the Hessian of the kinetic energy, bias forces from Lagrange's equations, qacc_smooth
 * by the articulated-body algorithm, every contact of the narrow phase, the rows of every contact and
 * the solver parameters of every row, the constraint solve as the unique minimiser of its convex
 * problem, the RK4 and semi-implicit Euler steps (tests/test_oracle_*.py); and the reference's own FD
 * driver and iLQR classes compiled verbatim against the shim in oracle/shim/ (oracle/_ref, built by
 * oracle/Makefile).  tools/mujoco_fixtures.py writes the fixtures that would pin it to MuJoCo itself.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use anything under oracle/.
 */
#ifndef MJO_H
#define MJO_H

#include "../include/ilqg_model.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MJO_MAXCON 96
#define MJO_MAXEFC 320
#define MJO_MINVAL 1e-15

typedef struct mjo_contact {
    double dist;
    double pos[3];
    double frame[9]; /* rows: normal, tangent1, tangent2 */
    int pair;        /* index into the model's pair tables */
} mjo_contact;

typedef struct mjo_data {
    /* state and inputs — one heap block, qpos immediately followed by qvel as in mjData
       (the reference relies on that: /root/reference/inc/differentiator.h:62) */
    double time;
    double* qpos;
    double* qvel;
    double* ctrl;
    double* qacc;
    double* qacc_warmstart;
    double* qfrc_applied;
    double* xfrc_applied;
    void* block;

    /* position stage */
    double xpos[ILQG_MAXBODY][3], xquat[ILQG_MAXBODY][4], xmat[ILQG_MAXBODY][9], xipos[ILQG_MAXBODY][3];
    double xanchor[ILQG_MAXJNT][3], xaxis[ILQG_MAXJNT][3];
    double geom_xpos[ILQG_MAXGEOM][3], geom_xmat[ILQG_MAXGEOM][9];
    double subtree_com[ILQG_MAXBODY][3];
    double cinert[ILQG_MAXBODY][10];
    double cdof[ILQG_MAXV][6];
    double crb[ILQG_MAXBODY][10];
    double qM[ILQG_MAXV * ILQG_MAXV];  /* dense, row-major nv x nv */
    double qL[ILQG_MAXV * ILQG_MAXV];  /* Cholesky factor of qM (lower) */
    int ncon, nefc;
    mjo_contact contact[MJO_MAXCON];
    double efc_J[MJO_MAXEFC * ILQG_MAXV];
    double efc_pos[MJO_MAXEFC], efc_margin[MJO_MAXEFC], efc_diagApprox[MJO_MAXEFC];
    double efc_R[MJO_MAXEFC], efc_D[MJO_MAXEFC], efc_KBIP[MJO_MAXEFC][4];
    /* velocity stage */
    double cvel[ILQG_MAXBODY][6], cdof_dot[ILQG_MAXV][6];
    double qfrc_passive[ILQG_MAXV], qfrc_bias[ILQG_MAXV];
    double efc_vel[MJO_MAXEFC], efc_aref[MJO_MAXEFC];
    /* acceleration stage */
    double qfrc_actuator[ILQG_MAXV], qfrc_smooth[ILQG_MAXV], qacc_smooth[ILQG_MAXV];
    double qfrc_constraint[ILQG_MAXV], efc_force[MJO_MAXEFC];
    int solver_iter; /* Newton iterations of the last solve */
    /* instrumentation: fp64 operations executed since last reset (FMA counted as 2) */
    double flops;
    /* scratch stack for the shim's mj_stackAlloc (/root/reference/src/mjderivative.cpp:51-53) */
    double stack[4 * ILQG_MAXV];
    int pstack;
} mjo_data;

mjo_data* mjo_make_data(const ilqg_model* m);
void mjo_delete_data(mjo_data* d);
void mjo_reset_data(const ilqg_model* m, mjo_data* d); /* qpos = qpos0, rest zero */
/* the fields cpMjData copies, /root/reference/src/util.cpp:4-13 */
void mjo_copy_state(const ilqg_model* m, mjo_data* dst, const mjo_data* src);

/* pipeline stages (names after MuJoCo's) */
void mjo_fwd_position(const ilqg_model* m, mjo_data* d);
void mjo_fwd_velocity(const ilqg_model* m, mjo_data* d);
void mjo_fwd_actuation(const ilqg_model* m, mjo_data* d);
void mjo_fwd_acceleration(const ilqg_model* m, mjo_data* d);
void mjo_fwd_constraint(const ilqg_model* m, mjo_data* d, int iterations, double tolerance);

/* mj_forward / mj_forwardSkip / mj_step with explicit solver limits
   (the reference mutates m->opt.iterations/tolerance around FD, mjderivative.cpp:241-254) */
void mjo_forward_skip(const ilqg_model* m, mjo_data* d, int skipstage, int iterations, double tolerance);
void mjo_forward(const ilqg_model* m, mjo_data* d);
void mjo_step(const ilqg_model* m, mjo_data* d);
void mjo_quat_integrate(double quat[4], const double vel[3], double scale);
void mjo_integrate_pos(const ilqg_model* m, double* qpos, const double* qvel, double dt);
double mjo_energy(const ilqg_model* m, mjo_data* d); /* kinetic + gravitational potential (test anchor) */

/* step cost callback for the oracle's own drivers */
typedef double (*mjo_cost_fn)(const double* qpos, const double* qvel, const double* ctrl, void* user);

/* Restated calcMJDerivatives for one knot (serial; all columns; same schedule, eps, stencils,
   warm-start handling and stage skipping as /root/reference/src/mjderivative.cpp:43-209).
   `scratch` is a caller-owned mjo_data.  Returns flops executed. */
void mjo_fd_knot(const ilqg_model* m, const double* qpos, const double* qvel, const double* ctrl, const double* warmstart,
                 mjo_cost_fn cost, void* user, double eps, int niter, int nwarmup, double* deriv, double* qacc_center,
                 mjo_data* scratch);
/* all knots, OpenMP over knots with persistent per-thread scratch ("fair CPU baseline") */
void mjo_fd_batch(const ilqg_model* m, int nknots, const double* qpos, const double* qvel, const double* ctrl,
                  const double* warmstart, mjo_cost_fn cost, void* user, double eps, int niter, int nwarmup, double* deriv,
                  double* qacc_center, int nthreads, double* flops_out);
/* n independent states, nsteps x mj_step each (OpenMP over states) */
void mjo_step_batch(const ilqg_model* m, int n, int nsteps, double* qpos, double* qvel, const double* ctrl, double* warmstart,
                    double* qacc, int nthreads);
void mjo_forward_batch(const ilqg_model* m, int n, const double* qpos, const double* qvel, const double* ctrl,
                       double* warmstart, double* qacc, int nthreads);

/* quadratic+linear step cost of include/ilqg_b200.h (struct ilqg_cost), for mjo_cost_fn's `user` */
double mjo_cost_quadratic(const double* qpos, const double* qvel, const double* ctrl, void* ilqg_cost_ptr);

#ifdef __cplusplus
}
#endif
#endif
