/* mjo_fd.c — CPU ORACLE (test infrastructure): restatement of the reference's FD schedule
 * (/root/reference/src/mjderivative.cpp:43-209 `worker`, :212-255 `calcMJDerivatives`) on top of
 * mjo_engine.c, plus OpenMP batch drivers used as the "fair CPU baseline" of bench.py. */
#include "mjo.h"
#include "../include/ilqg_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void load_knot(const ilqg_model* m, mjo_data* d, const double* qpos, const double* qvel, const double* ctrl,
                      const double* warm) {
    d->time = 0;
    memcpy(d->qpos, qpos, sizeof(double) * m->nq);
    memcpy(d->qvel, qvel, sizeof(double) * m->nv);
    memcpy(d->ctrl, ctrl, sizeof(double) * m->nu);
    if (warm) memcpy(d->qacc_warmstart, warm, sizeof(double) * m->nv);
    else memset(d->qacc_warmstart, 0, sizeof(double) * m->nv);
    memset(d->qacc, 0, sizeof(double) * m->nv);
    memset(d->qfrc_applied, 0, sizeof(double) * m->nv);
    memset(d->xfrc_applied, 0, sizeof(double) * 6 * m->nbody);
}

void mjo_fd_knot(const ilqg_model* m, const double* qpos, const double* qvel, const double* ctrl, const double* warmstart,
                 mjo_cost_fn cost, void* user, double eps, int niter, int nwarmup, double* deriv, double* qacc_center,
                 mjo_data* d) {
    int nv = m->nv, nu = m->nu, nq = m->nq;
    double temp[ILQG_MAXV], warm[ILQG_MAXV];
    /* centre point: full pipeline, then nwarmup-1 extra solves (mjderivative.cpp:61-68);
       solver pinned to niter iterations, tolerance 0 (:241-242) */
    load_knot(m, d, qpos, qvel, ctrl, warmstart);
    mjo_forward_skip(m, d, ILQG_STAGE_NONE, niter, 0.0);
    for (int rep = 1; rep < nwarmup; rep++) mjo_forward_skip(m, d, ILQG_STAGE_VEL, niter, 0.0);
    const double* output = d->qacc;
    double costCenter = cost ? cost(qpos, qvel, ctrl, user) : 0.0;
    memcpy(warm, d->qacc_warmstart, sizeof(double) * nv);
    if (qacc_center) memcpy(qacc_center, d->qacc, sizeof(double) * nv);

    /* controls: skip = VEL, central difference (:78-111) */
    for (int i = 0; i < nv; i++) {
        if (i >= nu) break;
        d->ctrl[i] = ctrl[i] + eps;
        if (cost) deriv[2 * nv * nv + nv * nu + 2 * nv + i] = (cost(d->qpos, d->qvel, d->ctrl, user) - costCenter) / eps;
        memcpy(d->qacc_warmstart, warm, sizeof(double) * nv);
        mjo_forward_skip(m, d, ILQG_STAGE_VEL, niter, 0.0);
        memcpy(temp, output, sizeof(double) * nv);
        d->ctrl[i] = ctrl[i] - eps;
        memcpy(d->qacc_warmstart, warm, sizeof(double) * nv);
        mjo_forward_skip(m, d, ILQG_STAGE_VEL, niter, 0.0);
        for (int j = 0; j < nv; j++) deriv[2 * nv * nv + i + j * nu] = (temp[j] - output[j]) / (2 * eps);
        d->ctrl[i] = ctrl[i];
    }
    /* velocities: skip = POS (:114-142) */
    for (int i = 0; i < nv; i++) {
        d->qvel[i] = qvel[i] + eps;
        if (cost) deriv[2 * nv * nv + nv * nu + nv + i] = (cost(d->qpos, d->qvel, d->ctrl, user) - costCenter) / eps;
        memcpy(d->qacc_warmstart, warm, sizeof(double) * nv);
        mjo_forward_skip(m, d, ILQG_STAGE_POS, niter, 0.0);
        memcpy(temp, output, sizeof(double) * nv);
        d->qvel[i] = qvel[i] - eps;
        memcpy(d->qacc_warmstart, warm, sizeof(double) * nv);
        mjo_forward_skip(m, d, ILQG_STAGE_POS, niter, 0.0);
        for (int j = 0; j < nv; j++) deriv[nv * nv + i + j * nv] = (temp[j] - output[j]) / (2 * eps);
        d->qvel[i] = qvel[i];
    }
    /* positions: skip = NONE; quaternion dofs are perturbed in the tangent space (:145-206) */
    for (int i = 0; i < nv; i++) {
        int jid = m->dof_jntid[i];
        int quatadr = -1, dofpos = 0;
        if (m->jnt_type[jid] == ILQG_JNT_BALL) {
            quatadr = m->jnt_qposadr[jid];
            dofpos = i - m->jnt_dofadr[jid];
        } else if (m->jnt_type[jid] == ILQG_JNT_FREE && i >= m->jnt_dofadr[jid] + 3) {
            quatadr = m->jnt_qposadr[jid] + 3;
            dofpos = i - m->jnt_dofadr[jid] - 3;
        }
        int qi = m->jnt_qposadr[jid] + i - m->jnt_dofadr[jid];
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
            memcpy(d->qpos, qpos, sizeof(double) * nq);
            if (quatadr >= 0) {
                double angvel[3] = {0, 0, 0};
                angvel[dofpos] = sgn * eps;
                mjo_quat_integrate(d->qpos + quatadr, angvel, 1);
            } else if (sgn > 0)
                d->qpos[qi] += eps;
            else
                d->qpos[qi] -= eps;
            if (sgn > 0 && cost) deriv[2 * nv * nv + nv * nu + i] = (cost(d->qpos, d->qvel, d->ctrl, user) - costCenter) / eps;
            memcpy(d->qacc_warmstart, warm, sizeof(double) * nv);
            mjo_forward_skip(m, d, ILQG_STAGE_NONE, niter, 0.0);
            if (sgn > 0) memcpy(temp, output, sizeof(double) * nv);
        }
        for (int j = 0; j < nv; j++) deriv[i + j * nv] = (temp[j] - output[j]) / (2 * eps);
        memcpy(d->qpos, qpos, sizeof(double) * nq);
    }
}

void mjo_fd_batch(const ilqg_model* m, int nknots, const double* qpos, const double* qvel, const double* ctrl,
                  const double* warmstart, mjo_cost_fn cost, void* user, double eps, int niter, int nwarmup, double* deriv,
                  double* qacc_center, int nthreads, double* flops_out) {
    int nv = m->nv, nu = m->nu, nq = m->nq;
    int nd = nv * (2 * nv + nu) + 2 * nv + nu;
    double flops = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads) reduction(+ : flops)
    {
        mjo_data* d = mjo_make_data(m);
#pragma omp for schedule(dynamic, 4)
        for (int k = 0; k < nknots; k++)
            mjo_fd_knot(m, qpos + (size_t)k * nq, qvel + (size_t)k * nv, ctrl + (size_t)k * nu,
                        warmstart ? warmstart + (size_t)k * nv : NULL, cost, user, eps, niter, nwarmup, deriv + (size_t)k * nd,
                        qacc_center ? qacc_center + (size_t)k * nv : NULL, d);
        flops += d->flops;
        mjo_delete_data(d);
    }
    if (flops_out) *flops_out = flops;
}

void mjo_step_batch(const ilqg_model* m, int n, int nsteps, double* qpos, double* qvel, const double* ctrl, double* warmstart,
                    double* qacc, int nthreads) {
    int nv = m->nv, nu = m->nu, nq = m->nq;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        mjo_data* d = mjo_make_data(m);
#pragma omp for schedule(dynamic, 4)
        for (int k = 0; k < n; k++) {
            load_knot(m, d, qpos + (size_t)k * nq, qvel + (size_t)k * nv, ctrl + (size_t)k * nu,
                      warmstart ? warmstart + (size_t)k * nv : NULL);
            for (int s = 0; s < nsteps; s++) mjo_step(m, d);
            memcpy(qpos + (size_t)k * nq, d->qpos, sizeof(double) * nq);
            memcpy(qvel + (size_t)k * nv, d->qvel, sizeof(double) * nv);
            if (warmstart) memcpy(warmstart + (size_t)k * nv, d->qacc_warmstart, sizeof(double) * nv);
            if (qacc) memcpy(qacc + (size_t)k * nv, d->qacc, sizeof(double) * nv);
        }
        mjo_delete_data(d);
    }
}

void mjo_forward_batch(const ilqg_model* m, int n, const double* qpos, const double* qvel, const double* ctrl,
                       double* warmstart, double* qacc, int nthreads) {
    int nv = m->nv, nu = m->nu, nq = m->nq;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        mjo_data* d = mjo_make_data(m);
#pragma omp for schedule(dynamic, 4)
        for (int k = 0; k < n; k++) {
            load_knot(m, d, qpos + (size_t)k * nq, qvel + (size_t)k * nv, ctrl + (size_t)k * nu,
                      warmstart ? warmstart + (size_t)k * nv : NULL);
            mjo_forward(m, d);
            if (warmstart) memcpy(warmstart + (size_t)k * nv, d->qacc_warmstart, sizeof(double) * nv);
            memcpy(qacc + (size_t)k * nv, d->qacc, sizeof(double) * nv);
        }
        mjo_delete_data(d);
    }
}

typedef struct mjo_quad_ctx { const ilqg_model* m; const ilqg_cost* c; } mjo_quad_ctx;

double mjo_cost_quadratic(const double* qpos, const double* qvel, const double* ctrl, void* ctx) {
    const mjo_quad_ctx* q = (const mjo_quad_ctx*)ctx;
    const ilqg_model* m = q->m;
    const ilqg_cost* c = q->c;
    double g = 0;
    for (int i = 0; i < m->nq; i++) { g += c->q2[i] * qpos[i] * qpos[i]; g += c->q1[i] * qpos[i]; }
    for (int i = 0; i < m->nv; i++) { g += c->v2[i] * qvel[i] * qvel[i]; g += c->v1[i] * qvel[i]; }
    for (int i = 0; i < m->nu; i++) { g += c->u2[i] * ctrl[i] * ctrl[i]; g += c->u1[i] * ctrl[i]; }
    return g;
}

/* convenience for ctypes: FD batch with the quadratic cost (cost may be NULL) */
void mjo_fd_batch_quad(const ilqg_model* m, int nknots, const double* qpos, const double* qvel, const double* ctrl,
                       const double* warmstart, const ilqg_cost* cost, double eps, int niter, int nwarmup, double* deriv,
                       double* qacc_center, int nthreads, double* flops_out) {
    mjo_quad_ctx ctx = {m, cost};
    mjo_fd_batch(m, nknots, qpos, qvel, ctrl, warmstart, cost ? mjo_cost_quadratic : NULL, &ctx, eps, niter, nwarmup, deriv,
                 qacc_center, nthreads, flops_out);
}
