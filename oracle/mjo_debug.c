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

/* Everything the MuJoCo cross-check (tools/mujoco_fixtures.py -> tests/test_mujoco_fixtures.py) and the hand-derived contact
   anchors compare, for one state: the names follow mjData's.  Any pointer may be NULL.
   con: [MJO_MAXCON][13] = dist, pos[3], frame[9];  con_geom: [MJO_MAXCON][2] = geom1, geom2;
   efc: [MJO_MAXEFC][7] = pos, margin, diagApprox, R, D, aref, force;  efc_J: [MJO_MAXEFC][nv] */
void mjo_debug_dump(const ilqg_model* m, const double* qpos, const double* qvel, const double* ctrl, const double* warm, int iterations,
                    double tolerance, double* xpos, double* xquat, double* xipos, double* subtree_com, double* cinert, double* cdof,
                    double* qM, double* qfrc_bias, double* qfrc_passive, double* qfrc_actuator, double* qacc_smooth, double* qacc,
                    int* counts /* ncon, nefc, solver iterations */, double* con, int* con_geom, double* efc_J, double* efc) {
    mjo_data* d = mjo_make_data(m);
    const int nv = m->nv, nb = m->nbody;
    memcpy(d->qpos, qpos, sizeof(double) * m->nq);
    memcpy(d->qvel, qvel, sizeof(double) * nv);
    memcpy(d->ctrl, ctrl, sizeof(double) * m->nu);
    if (warm) memcpy(d->qacc_warmstart, warm, sizeof(double) * nv);
    mjo_forward_skip(m, d, ILQG_STAGE_NONE, iterations, tolerance);
    for (int b = 0; b < nb; b++) {
        if (xpos) memcpy(xpos + 3 * b, d->xpos[b], sizeof(double) * 3);
        if (xquat) memcpy(xquat + 4 * b, d->xquat[b], sizeof(double) * 4);
        if (xipos) memcpy(xipos + 3 * b, d->xipos[b], sizeof(double) * 3);
        if (subtree_com) memcpy(subtree_com + 3 * b, d->subtree_com[b], sizeof(double) * 3);
        if (cinert) memcpy(cinert + 10 * b, d->cinert[b], sizeof(double) * 10);
    }
    for (int i = 0; i < nv; i++)
        if (cdof) memcpy(cdof + 6 * i, d->cdof[i], sizeof(double) * 6);
    if (qM) memcpy(qM, d->qM, sizeof(double) * nv * nv);
    if (qfrc_bias) memcpy(qfrc_bias, d->qfrc_bias, sizeof(double) * nv);
    if (qfrc_passive) memcpy(qfrc_passive, d->qfrc_passive, sizeof(double) * nv);
    if (qfrc_actuator) memcpy(qfrc_actuator, d->qfrc_actuator, sizeof(double) * nv);
    if (qacc_smooth) memcpy(qacc_smooth, d->qacc_smooth, sizeof(double) * nv);
    if (qacc) memcpy(qacc, d->qacc, sizeof(double) * nv);
    if (counts) { counts[0] = d->ncon; counts[1] = d->nefc; counts[2] = d->solver_iter; }
    for (int c = 0; c < d->ncon; c++) {
        if (con) {
            con[13 * c] = d->contact[c].dist;
            memcpy(con + 13 * c + 1, d->contact[c].pos, sizeof(double) * 3);
            memcpy(con + 13 * c + 4, d->contact[c].frame, sizeof(double) * 9);
        }
        if (con_geom) { con_geom[2 * c] = m->pair_geom1[d->contact[c].pair]; con_geom[2 * c + 1] = m->pair_geom2[d->contact[c].pair]; }
    }
    for (int r = 0; r < d->nefc; r++) {
        if (efc_J) memcpy(efc_J + (size_t)r * nv, d->efc_J + (size_t)r * nv, sizeof(double) * nv);
        if (efc) {
            double* e = efc + 7 * r;
            e[0] = d->efc_pos[r]; e[1] = d->efc_margin[r]; e[2] = d->efc_diagApprox[r]; e[3] = d->efc_R[r]; e[4] = d->efc_D[r];
            e[5] = d->efc_aref[r]; e[6] = d->efc_force[r];
        }
    }
    mjo_delete_data(d);
}
