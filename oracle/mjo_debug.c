/* mjo_debug.c — CPU ORACLE (test infrastructure): introspection helpers for the anchor tests. */
#include "mjo.h"
#include <math.h>
#include <string.h>

/* mass matrix, bias force and energy at (qpos, qvel) */
void mjo_debug_mass_bias(const ilqg_model* m, const double* qpos, const double* qvel, double* M, double* bias, double* energy) {
    mjo_data* d = mjo_make_data(m);
    memcpy(d->qpos, qpos, sizeof(double) * m->nq);
    memcpy(d->qvel, qvel, sizeof(double) * m->nv);
    mjo_fwd_position(m, d);
    mjo_fwd_velocity(m, d);
    if (M) memcpy(M, d->qM, sizeof(double) * m->nv * m->nv);
    if (bias) memcpy(bias, d->qfrc_bias, sizeof(double) * m->nv);
    if (energy) *energy = mjo_energy(m, d);
    mjo_delete_data(d);
}

/* remove every non-conservative and constraint term: damping, limits, contacts, actuators stay (ctrl=0) */
void mjo_debug_strip(ilqg_model* m, int keep_springs) {
    for (int i = 0; i < m->nv; i++) m->dof_damping[i] = 0;
    for (int j = 0; j < m->njnt; j++) { m->jnt_limited[j] = 0; if (!keep_springs) m->jnt_stiffness[j] = 0; }
    m->npair = 0;
}

/* one forward evaluation; reports constraint state and the KKT residual of the solve:
   r = M a - qfrc_smooth - J' f,  f_i = -D_i min(0, J_i a - aref_i) */
void mjo_debug_forward(const ilqg_model* m, const double* qpos, const double* qvel, const double* ctrl, const double* warm,
                       int iterations, double tolerance, double* qacc, int* info /*ncon,nefc,iter,nactive*/, double* kkt,
                       double* efc_force /*MJO_MAXEFC*/, double* contact_dist /*MJO_MAXCON*/) {
    mjo_data* d = mjo_make_data(m);
    int nv = m->nv;
    memcpy(d->qpos, qpos, sizeof(double) * m->nq);
    memcpy(d->qvel, qvel, sizeof(double) * nv);
    memcpy(d->ctrl, ctrl, sizeof(double) * m->nu);
    if (warm) memcpy(d->qacc_warmstart, warm, sizeof(double) * nv);
    mjo_forward_skip(m, d, ILQG_STAGE_NONE, iterations, tolerance);
    memcpy(qacc, d->qacc, sizeof(double) * nv);
    double res[ILQG_MAXV];
    for (int i = 0; i < nv; i++) {
        double s = -d->qfrc_smooth[i];
        for (int k = 0; k < nv; k++) s += d->qM[i * nv + k] * d->qacc[k];
        res[i] = s;
    }
    int nact = 0;
    for (int r = 0; r < d->nefc; r++) {
        double jar = -d->efc_aref[r];
        for (int i = 0; i < nv; i++) jar += d->efc_J[(size_t)r * nv + i] * d->qacc[i];
        double f = jar < 0 ? -d->efc_D[r] * jar : 0;
        if (jar < 0) nact++;
        if (efc_force) efc_force[r] = f;
        for (int i = 0; i < nv; i++) res[i] -= d->efc_J[(size_t)r * nv + i] * f;
    }
    double n2 = 0;
    for (int i = 0; i < nv; i++) n2 += res[i] * res[i];
    if (kkt) *kkt = sqrt(n2);
    if (info) { info[0] = d->ncon; info[1] = d->nefc; info[2] = d->solver_iter; info[3] = nact; }
    if (contact_dist) for (int c = 0; c < d->ncon; c++) contact_dist[c] = d->contact[c].dist;
    mjo_delete_data(d);
}

/* fp64 operations of ONE mj_step at (qpos, qvel, ctrl) — the rollout's unit of work for bench.py's iLQR roofline */
void mjo_debug_step_flops(const ilqg_model* m, const double* qpos, const double* qvel, const double* ctrl, double* flops) {
    mjo_data* d = mjo_make_data(m);
    memcpy(d->qpos, qpos, sizeof(double) * m->nq);
    memcpy(d->qvel, qvel, sizeof(double) * m->nv);
    memcpy(d->ctrl, ctrl, sizeof(double) * m->nu);
    d->flops = 0;
    mjo_step(m, d);
    if (flops) *flops = d->flops;
    mjo_delete_data(d);
}
