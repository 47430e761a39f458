/* ilqg_b200.h — C ABI of the B200-native iLQG hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, returns an int status
 * (0 = ok) and never exits the process (the reference returns void and lets MuJoCo abort,
 * /root/reference/inc/mjderivative.h:7).  No torch / C++ types cross this boundary.
 *
 * Two flavours of each compute call:
 *   *_dev  : all buffers are DEVICE pointers on the handle's GPU; stream-ordered on `stream`
 *            (a cudaStream_t passed as void*; NULL = legacy default stream); asynchronous.
 *   *_host : all buffers are HOST pointers; the call copies in, launches, copies out and
 *            synchronises.  This is what the drop-in `calcMJDerivatives` / `mj_step` wrappers
 *            in ilqg-mujoco_b200/host/ use.
 * There is no CPU implementation behind these calls: without a CUDA device they fail with
 * ILQG_ERR_CUDA.
 *
 * Concurrency: a handle (and the iLQR workspaces created on it) owns device scratch — the cost block, the centre
 * accelerations, the work-class permutation, the profiling events — that every call reuses.  Calls on ONE handle must
 * therefore be issued on one stream at a time (stream-ordered, like the reference's single-threaded caller,
 * /root/reference/cmd/basic.cpp:158-164); create one handle per stream (or per thread) to overlap calls.
 *
 * Layouts are the reference's:
 *   deriv (per knot, ND = nv*(2nv+nu) + 2nv + nu doubles, /root/reference/inc/differentiator.h:56-61):
 *     [0, nv^2)                 d qacc_j / d qpos_i   at i + j*nv   (/root/reference/src/mjderivative.cpp:202)
 *     [nv^2, 2nv^2)             d qacc_j / d qvel_i   at i + j*nv   (:138)
 *     [2nv^2, 2nv^2+nv*nu)      d qacc_j / d ctrl_i   at i + j*nu   (:107)
 *     then dg/dqpos[nv], dg/dqvel[nv], dg/dctrl[nu]   (:174,:120,:88; forward differences)
 *   knot inputs: qpos[nq], qvel[nv], ctrl[nu], qacc_warmstart[nv] — the fields cpMjData copies
 *     (/root/reference/src/util.cpp:4-13) that the dynamics read (qfrc_applied/xfrc_applied are
 *     always zero in the reference and are not supported).
 */
#ifndef ILQG_B200_H
#define ILQG_B200_H

#include <stddef.h>

#include "ilqg_model.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ILQG_OK = 0,
    ILQG_ERR_ARG = 1,         /* null pointer / bad size */
    ILQG_ERR_MODEL = 2,       /* malformed model or MJCF outside the supported subset */
    ILQG_ERR_IO = 3,
    ILQG_ERR_CUDA = 4,        /* CUDA runtime error (text via ilqg_last_error) */
    ILQG_ERR_UNSUPPORTED = 5, /* no kernel instantiation for this model's shape */
    ILQG_ERR_NONFINITE = 6,   /* a rollout produced NaN/Inf (per-knot flags in `status`) */
    ILQG_ERR_CAPACITY = 7     /* more contacts / constraint rows than the kernel's shared-memory budget (per-knot flag) */
};

typedef struct ilqg_handle_s* ilqg_handle;

/* ---- model compilation (host only; replaces mj_loadXML, /root/reference/cmd/basic.cpp:123) */
int ilqg_compile_mjcf(const char* xml_path, ilqg_model* out, char* err, int errlen);
int ilqg_compile_mjcf_string(const char* xml, ilqg_model* out, char* err, int errlen);
int ilqg_model_save(const char* path, const ilqg_model* m);
int ilqg_model_load(const char* path, ilqg_model* m);
int ilqg_model_sizeof(void);
/* structural check of a table (counts, index ranges, tree order, supported joint / geom kinds); ilqg_create and
 * ilqg_model_load run it and refuse the table with ILQG_ERR_MODEL / ILQG_ERR_UNSUPPORTED.  err may be NULL. */
int ilqg_model_validate(const ilqg_model* m, char* err, int errlen);
/* byte offset / element count / type of a named ilqg_model field (bindings that hold the table as bytes) */
int ilqg_model_field(const char* name, int* offset, int* count, int* is_double);

/* ---- GPU-resident model (the role of mjModel* in every reference call) */
int ilqg_create(const ilqg_model* m, int device, ilqg_handle* out);
int ilqg_destroy(ilqg_handle h);
const char* ilqg_last_error(ilqg_handle h); /* h may be NULL: last creation error */
int ilqg_deriv_size(const ilqg_model* m);   /* ND */

/* ---- FD options: the file-static tunables of /root/reference/src/mjderivative.cpp:36-39 */
typedef struct ilqg_fd_opts {
    double eps;   /* 1e-6 */
    int niter;    /* 30: solver iterations pinned during FD (:241) */
    int nwarmup;  /* 3: centre-point repetitions (:67) */
} ilqg_fd_opts;
void ilqg_fd_opts_default(ilqg_fd_opts* o);

/* ---- device-evaluable step cost.  The reference's stepCostFn_t is a host function pointer
 *      (/root/reference/inc/mjderivative.h:5) that reads qpos/qvel/ctrl only; on the GPU the
 *      same role is played by  g = sum_i a2_i z_i^2 + a1_i z_i  over z = (qpos, qvel, ctrl).
 *      It covers /root/reference/inc/inverted_pendulum/cost.h:7-17 (q2={1,10}, v2={1,10}, u2={1})
 *      and the test's cost /root/reference/tst/test_derivatives.cpp:16-20 (q1={1}).  Terms are
 *      accumulated qpos, then qvel, then ctrl, index ascending, quadratic before linear. */
typedef struct ilqg_cost {
    double q2[ILQG_MAXQ], q1[ILQG_MAXQ];
    double v2[ILQG_MAXV], v1[ILQG_MAXV];
    double u2[ILQG_MAXU], u1[ILQG_MAXU];
} ilqg_cost;

/* ---- FD linearisation of all knots in one launch.
 * Replaces calcMJDerivatives (/root/reference/inc/mjderivative.h:7, body
 * /root/reference/src/mjderivative.cpp:43-255) called once per knot by
 * Differentiator::updateDerivatives (/root/reference/inc/differentiator.h:87).
 *   qpos[nknots*nq], qvel[nknots*nv], ctrl[nknots*nu], warmstart[nknots*nv]  (knot-major)
 *   cost      : NULL -> the 2nv+nu cost-gradient entries of deriv are left untouched
 *               (host wrappers fill them with the caller's stepCostFn)
 *   deriv     : [nknots*ND] out
 *   qacc_out  : optional [nknots*nv] centre accelerations after warm-up (the value the reference
 *               leaves in the workers' qacc_warmstart); may be NULL
 *   status    : optional [nknots] out; 0 ok, ILQG_ERR_NONFINITE
 */
int ilqg_fd_batch_dev(ilqg_handle h, int nknots, const double* qpos, const double* qvel, const double* ctrl,
                      const double* warmstart, const ilqg_cost* cost, const ilqg_fd_opts* opts, double* deriv,
                      double* qacc_out, int* status, void* stream);
int ilqg_fd_batch_host(ilqg_handle h, int nknots, const double* qpos, const double* qvel, const double* ctrl,
                       const double* warmstart, const ilqg_cost* cost, const ilqg_fd_opts* opts, double* deriv,
                       double* qacc_out, int* status);

/* Pageable caller buffers of ilqg_fd_batch_host (what a calcMJDerivatives caller's malloc'ed arrays are): batches of 4 MB and more are
 * staged by the call itself through a pinned mirror owned by the handle, by ILQG_HOST_THREADS copy threads (default by core count,
 * 0 = the driver's single-threaded staging) working chunk by chunk beside the copy engines; nothing changes for the caller.
 * opt-in: page-lock the caller's large host buffers of ilqg_fd_batch_host (cudaHostRegister) the first time they are seen; they
 * stay registered until ilqg_set_host_pinning(h, 0) or ilqg_destroy.  The reference's caller owns one malloc'ed deriv array for the
 * life of a Differentiator (/root/reference/inc/differentiator.h:56): pageable memory is copied device-to-host at about a fifth of
 * the pinned rate.  The caller must not free a registered buffer while the handle lives (or must switch pinning off first). */
int ilqg_set_host_pinning(ilqg_handle h, int on);

/* ---- optional per-knot diagnostics of the centre evaluation (mj_forward + warm-up solves, /root/reference/src/mjderivative.cpp:64-68).
 * The reference pins the solver to `niter` iterations / tolerance 0 (:241-242); the kernels' solver leaves earlier when it has
 * reached the exact minimiser of the convex piecewise-quadratic cost ("Solver") — these counters say how many Newton
 * iterations actually ran.  diag_dev: DEVICE array [nknots][ILQG_DIAG_INTS] written by the following ilqg_fd_batch_dev /
 * ilqg_fd_batch_dev_scatter calls on the handle (NULL switches it off; the *_host call does not write it). */
#define ILQG_DIAG_INTS 8
enum {
    ILQG_DIAG_NEFC = 0,        /* constraint rows at the centre point */
    ILQG_DIAG_ITERS_FIRST = 1, /* Newton iterations of the first centre solve (from the caller's warm start) */
    ILQG_DIAG_ITERS_ALL = 2,   /* Newton iterations of all nwarmup centre solves */
    ILQG_DIAG_NACTIVE = 3,     /* rows active (force > 0) at the centre solution */
    ILQG_DIAG_CYC_BUILD = 4,   /* SM cycles the rollout spent in the position / velocity / actuation stages */
    ILQG_DIAG_CYC_SOLVE = 5,   /* SM cycles it spent in the nwarmup solves */
    ILQG_DIAG_NCON = 6,        /* contact points at the centre point (nefc also counts joints at their limits) */
    ILQG_DIAG_CYC_COLUMNS = 7  /* one-launch kernel only: SM cycles until the knot's slowest perturbed solve was done (else 0) */
};
int ilqg_fd_set_diag(ilqg_handle h, int* diag_dev);

/* ---- multi-GPU: the knots of ONE long horizon sharded over the GPUs of a node (BASELINE configs[4]).
 * The reference computes the knots' derivatives one after the other inside ILQR::backwardPass
 * (/root/reference/inc/ilqr.h:144-154); FD at knot n reads only knot n's inputs (/root/reference/src/mjderivative.cpp:61,72),
 * so ranks take contiguous knot ranges.  Instead of an all-gather after the kernels, ilqg_fd_batch_dev_scatter lets the
 * FD kernels' write-out store every deriv block to `ndst` destinations at once: this rank's copy of the horizon's deriv
 * array and the peers' copies, mapped into this process with CUDA IPC (stores travel over NVLink / NVSwitch).
 *   ilqg_peer_alloc   cudaMalloc a peer-visible buffer on the handle's GPU and export its IPC handle (64 bytes)
 *   ilqg_peer_open    map another rank's buffer (same node) into this process; ilqg_peer_close unmaps it
 *   dsts[i]           device pointer to where THIS call's knot 0 goes in destination i (base + first_knot * ND doubles)
 * Completion is stream-ordered per rank; ranks need one barrier (ilqg_peer_barrier, or any other) before reading blocks
 * written by their peers. */
#define ILQG_MAX_PEERS 8
#define ILQG_IPC_HANDLE_BYTES 64
int ilqg_peer_alloc(ilqg_handle h, size_t bytes, void** dev_ptr, unsigned char* ipc_handle);
int ilqg_peer_open(ilqg_handle h, const unsigned char* ipc_handle, void** dev_ptr);
int ilqg_peer_close(ilqg_handle h, void* dev_ptr);
int ilqg_peer_free(ilqg_handle h, void* dev_ptr);
/* node-wide barrier through peer memory (one 1-warp kernel per rank, no NCCL call): flags[r] = rank r's flag array
 * (ILQG_MAX_PEERS zero-initialised ints inside a peer buffer); epoch increases by one per call on every rank. */
int ilqg_peer_barrier(ilqg_handle h, int* const* flags, int nranks, int rank, int epoch, void* stream);
int ilqg_peer_barrier_timed_out(ilqg_handle h);
int ilqg_fd_batch_dev_scatter(ilqg_handle h, int nknots, const double* qpos, const double* qvel, const double* ctrl,
                              const double* warmstart, const ilqg_cost* cost, const ilqg_fd_opts* opts, double* const* dsts,
                              int ndst, double* qacc_out, int* status, void* stream);

/* ---- forward dynamics / stepping of n independent states.
 * ilqg_forward_*: mj_forward (/root/reference/src/mjderivative.cpp:64): qacc[n*nv] out; warmstart in/out.
 * ilqg_step_*   : nsteps x mj_step (/root/reference/inc/ilqr.h:86,128): qpos, qvel, warmstart in/out;
 *                 ctrl is held constant; qacc (optional) receives the last step's acceleration.
 * iterations/tolerance are the model's (the XML's), as in the reference's rollouts. */
int ilqg_forward_batch_dev(ilqg_handle h, int n, const double* qpos, const double* qvel, const double* ctrl,
                           double* warmstart, double* qacc, void* stream);
int ilqg_forward_batch_host(ilqg_handle h, int n, const double* qpos, const double* qvel, const double* ctrl,
                            double* warmstart, double* qacc);
int ilqg_step_batch_dev(ilqg_handle h, int n, int nsteps, double* qpos, double* qvel, const double* ctrl,
                        double* warmstart, double* qacc, void* stream);
int ilqg_step_batch_host(ilqg_handle h, int n, int nsteps, double* qpos, double* qvel, const double* ctrl,
                         double* warmstart, double* qacc);

/* ---- batched iLQR: ninst independent problems of horizon N (T = N+1 knots) resident on the GPU.
 * Replaces ILQR<nv,nu,N> (/root/reference/inc/ilqr.h:14-188) for a whole batch:
 *   ilqg_ilqr_init_*      ILQR::ILQR (:69-97)  open-loop rollout under the initial control, K = k = 0
 *   ilqg_ilqr_set_state_* ILQR::setDInit (:110-113)
 *   ilqg_ilqr_iterate     niter x ILQR::iterate (:179-186): forwardPass (:116-130) for every line-search step size in one
 *                         launch, acceptance of the first alpha in ladder order whose trajectory cost improves (the A10
 *                         specification, oracle/mjo_ilqr.c), FD of all knots (differentiator.h:85-93), backwardPass (:133-176).
 *                         accept_always != 0 with alphas = {1} reproduces the reference (full step, no cost test).
 * Index n of a trajectory runs as in the reference: n = N is the initial knot, n = 0 the final one (ilqr.h:52).
 * K[n] is nu x 2nv column-major (Eigen's layout); V is 2nv x 2nv column-major; the model must have nq == nv. */
typedef struct ilqg_ilqr_s* ilqg_ilqr;
int ilqg_ilqr_create(ilqg_handle h, int ninst, int N, int nalpha, const double* alphas /* host, NULL: 1,1/2,1/4,.. */, ilqg_ilqr* out);
int ilqg_ilqr_destroy(ilqg_ilqr w);
int ilqg_ilqr_set_cost(ilqg_ilqr w, const ilqg_cost* cost); /* host struct; NULL selects host-cost mode */
int ilqg_ilqr_set_mu(ilqg_ilqr w, double mu);               /* Levenberg-Marquardt term, default 1000 (ilqr.h:65) */
/* corrected != 0: assemble A/B from deriv as d qacc_j / d x_i (undoes the column-major-view quirk of
 * /root/reference/inc/differentiator.h:68-71); default 0 = the reference's matrices (parity mode) */
int ilqg_ilqr_set_layout(ilqg_ilqr w, int corrected);
/* opt-in mu schedule (the README of the reference advertises regularisation it does not implement, README.md:9-17):
 * per instance, mu /= factor after an accepted line-search step, mu *= factor after a rejected ladder, clamped to
 * [mu_min, mu_max]; factor <= 1 (default) keeps the reference's constant mu */
int ilqg_ilqr_set_mu_schedule(ilqg_ilqr w, double factor, double mu_min, double mu_max);
int ilqg_ilqr_init_dev(ilqg_ilqr w, const double* qpos, const double* qvel, const double* ctrl, const double* warm, void* stream);
int ilqg_ilqr_init_host(ilqg_ilqr w, const double* qpos, const double* qvel, const double* ctrl, const double* warm);
int ilqg_ilqr_set_state_dev(ilqg_ilqr w, const double* qpos, const double* qvel, const double* warm, void* stream);
int ilqg_ilqr_set_state_host(ilqg_ilqr w, const double* qpos, const double* qvel, const double* warm);
int ilqg_ilqr_iterate(ilqg_ilqr w, int niter, int accept_always, void* stream); /* asynchronous on `stream` */
/* the phases of one iteration, separately (ilqg_ilqr_iterate = forward; linearise; backward) */
int ilqg_ilqr_forward(ilqg_ilqr w, int accept_always, void* stream);
int ilqg_ilqr_linearise(ilqg_ilqr w, void* stream);
int ilqg_ilqr_backward(ilqg_ilqr w, void* stream);
/* host-cost mode (ilqg_ilqr_set_cost(w, NULL)): the caller's stepCostFn_t cannot run on the device, so the caller evaluates
   the forward-difference cost rows itself and uploads them between linearise and backward: rows[ninst][T][2nv+nu] */
int ilqg_ilqr_put_cost_rows_host(ilqg_ilqr w, const double* rows);
int ilqg_ilqr_get_knots_host(ilqg_ilqr w, double* qpos, double* qvel, double* ctrl, double* warm);
int ilqg_ilqr_iterations_done(ilqg_ilqr w);
/* the first control of every problem, u0[ninst][nu] (dArray[N]->ctrl: what /root/reference/src/inverted_pendulum/inverted_pendulum.cpp:26
 * applies to the plant), and optionally the cost trace [ninst][min(iterations,256)]; host pointers, synchronises */
int ilqg_ilqr_get_first_control_host(ilqg_ilqr w, double* u0, double* Jtrace);
/* the same with the cost trace cut to the last min(nlast, iterations, 256) iterations, Jtrace[ninst][that many] — what one MPC step of
 * `nlast` iterations on a long-lived workspace reads back (the reference's InvertedPendulum::forward keeps no trace at all) */
int ilqg_ilqr_get_first_control_last_host(ilqg_ilqr w, int nlast, double* u0, double* Jtrace);
/* results, instance-major on the host; any pointer may be NULL.  Jtrace/accepted: [ninst][min(iterations,256)] */
int ilqg_ilqr_get_host(ilqg_ilqr w, double* qpos, double* qvel, double* ctrl, double* K, double* k, double* V, double* v,
                       double* Jtrace, int* accepted);

/* ---- diagnostics */
long ilqg_launch_count(ilqg_handle h);        /* kernels launched through this handle so far */
const char* ilqg_engine_name(ilqg_handle h);  /* name of the kernel instantiation serving this model */
int ilqg_fp64_peak(int device, double* tflops); /* measured fp64 FMA throughput (roofline denominator) */
int ilqg_set_profiling(ilqg_handle h, int on);  /* record CUDA events around each FD kernel on the launch stream */
int ilqg_fd_last_kernel_ms(ilqg_handle h, float* center_ms, float* perturb_ms); /* durations of the last FD call's kernels */
/* the same, with the perturbed evaluations split by kernel: qvel/ctrl columns (stage-skipping) and qpos columns */
int ilqg_fd_last_stage_ms(ilqg_handle h, float* center_ms, float* velctrl_ms, float* qpos_ms);

#ifdef __cplusplus
}
#endif
#endif
