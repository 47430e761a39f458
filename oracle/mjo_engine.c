/* mjo_engine.c — CPU ORACLE (test infrastructure): fp64 restatement of MuJoCo 2.x forward
 * dynamics for the MJCF subset of /root/reference/res/ (the three .xml models).  See mjo.h for what pins it.
 *
 * Follows the stage structure of mj_forwardSkip as the reference uses it
 * (/root/reference/src/mjderivative.cpp:64,68,92,124,178): position stage (kinematics, com,
 * CRBA, factor, collision, constraint rows), velocity stage (com velocities, passive forces,
 * reference acceleration, RNE bias), acceleration stage (actuation, smooth acceleration,
 * constraint solve with warm start).  SURVEY.md Appendix A/A.2 lists the formulas frozen here.
 */
#include "mjo.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FL(d, n) ((d)->flops += (n))

/* ------------------------------------------------------------------ small helpers */
static void v3_cross(double* r, const double* a, const double* b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static double v3_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double v3_normalize(double* a) {
    double n = sqrt(v3_dot(a, a));
    if (n < MJO_MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
    a[0] /= n; a[1] /= n; a[2] /= n;
    return n;
}
static void quat_mul(double* r, const double* a, const double* b) {
    double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_normalize(double* q) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < MJO_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void quat2mat(double* m, const double* q) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static void mat_vec(double* r, const double* m, const double* v) {
    double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
           z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void axisangle2quat(double* q, const double* axis, double angle) {
    double s = sin(angle * 0.5);
    q[0] = cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

void mjo_quat_integrate(double quat[4], const double vel[3], double scale) {
    /* mju_quatIntegrate: quat <- normalize(quat) * quat(axis = vel/|vel|, angle = scale*|vel|) */
    double ax[3] = {vel[0], vel[1], vel[2]};
    double n = v3_normalize(ax);
    double qr[4], out[4];
    axisangle2quat(qr, ax, scale * n);
    quat_normalize(quat);
    quat_mul(out, quat, qr);
    memcpy(quat, out, sizeof out);
}

void mjo_integrate_pos(const ilqg_model* m, double* qpos, const double* qvel, double dt) {
    for (int j = 0; j < m->njnt; j++) {
        int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
        if (m->jnt_type[j] == ILQG_JNT_FREE) {
            for (int k = 0; k < 3; k++) qpos[qa + k] += dt * qvel[da + k];
            mjo_quat_integrate(qpos + qa + 3, qvel + da + 3, dt);
        } else if (m->jnt_type[j] == ILQG_JNT_BALL)
            mjo_quat_integrate(qpos + qa, qvel + da, dt);
        else
            qpos[qa] += dt * qvel[da];
    }
}

/* ------------------------------------------------------------------ data */
mjo_data* mjo_make_data(const ilqg_model* m) {
    mjo_data* d = (mjo_data*)calloc(1, sizeof(mjo_data));
    if (!d) return NULL;
    size_t n = (size_t)m->nq + 5 * (size_t)m->nv + m->nu + 6 * (size_t)m->nbody;
    double* b = (double*)calloc(n, sizeof(double));
    d->block = b;
    d->qpos = b; b += m->nq;
    d->qvel = b; b += m->nv;
    d->qacc_warmstart = b; b += m->nv;
    d->ctrl = b; b += m->nu;
    d->qfrc_applied = b; b += m->nv;
    d->xfrc_applied = b; b += 6 * m->nbody;
    d->qacc = b; b += m->nv;
    mjo_reset_data(m, d);
    return d;
}
void mjo_delete_data(mjo_data* d) {
    if (!d) return;
    free(d->block);
    free(d);
}
void mjo_reset_data(const ilqg_model* m, mjo_data* d) {
    d->time = 0;
    memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
    memset(d->qvel, 0, sizeof(double) * m->nv);
    memset(d->qacc, 0, sizeof(double) * m->nv);
    memset(d->qacc_warmstart, 0, sizeof(double) * m->nv);
    memset(d->qfrc_applied, 0, sizeof(double) * m->nv);
    memset(d->xfrc_applied, 0, sizeof(double) * 6 * m->nbody);
    memset(d->ctrl, 0, sizeof(double) * m->nu);
    d->ncon = d->nefc = 0;
    d->flops = 0;
}
void mjo_copy_state(const ilqg_model* m, mjo_data* dst, const mjo_data* src) {
    dst->time = src->time;
    memcpy(dst->qpos, src->qpos, sizeof(double) * m->nq);
    memcpy(dst->qvel, src->qvel, sizeof(double) * m->nv);
    memcpy(dst->qacc, src->qacc, sizeof(double) * m->nv);
    memcpy(dst->qacc_warmstart, src->qacc_warmstart, sizeof(double) * m->nv);
    memcpy(dst->qfrc_applied, src->qfrc_applied, sizeof(double) * m->nv);
    memcpy(dst->xfrc_applied, src->xfrc_applied, sizeof(double) * 6 * m->nbody);
    memcpy(dst->ctrl, src->ctrl, sizeof(double) * m->nu);
}

/* ------------------------------------------------------------------ position stage */
static void kinematics(const ilqg_model* m, mjo_data* d) {
    /* mj_kinematics: normalise quaternions in qpos, then walk the tree */
    for (int j = 0; j < m->njnt; j++) {
        if (m->jnt_type[j] == ILQG_JNT_FREE) quat_normalize(d->qpos + m->jnt_qposadr[j] + 3);
        if (m->jnt_type[j] == ILQG_JNT_BALL) quat_normalize(d->qpos + m->jnt_qposadr[j]);
    }
    d->xpos[0][0] = d->xpos[0][1] = d->xpos[0][2] = 0;
    d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
    quat2mat(d->xmat[0], d->xquat[0]);
    d->xipos[0][0] = d->xipos[0][1] = d->xipos[0][2] = 0;
    for (int b = 1; b < m->nbody; b++) {
        int p = m->body_parentid[b];
        double xpos[3], xquat[4], t[3];
        mat_vec(t, d->xmat[p], m->body_pos[b]);
        for (int k = 0; k < 3; k++) xpos[k] = d->xpos[p][k] + t[k];
        quat_mul(xquat, d->xquat[p], m->body_quat[b]);
        for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
            int j = m->body_jntadr[b] + jj, qa = m->jnt_qposadr[j];
            double mat[9];
            if (m->jnt_type[j] == ILQG_JNT_FREE) {
                for (int k = 0; k < 3; k++) xpos[k] = d->qpos[qa + k];
                for (int k = 0; k < 4; k++) xquat[k] = d->qpos[qa + 3 + k];
                for (int k = 0; k < 3; k++) d->xanchor[j][k] = xpos[k];
                d->xaxis[j][0] = 0; d->xaxis[j][1] = 0; d->xaxis[j][2] = 1;
                continue;
            }
            quat2mat(mat, xquat);
            mat_vec(t, mat, m->jnt_pos[j]);
            for (int k = 0; k < 3; k++) d->xanchor[j][k] = xpos[k] + t[k];
            mat_vec(d->xaxis[j], mat, m->jnt_axis[j]);
            if (m->jnt_type[j] == ILQG_JNT_BALL) { /* rotate about the anchor by the joint's quaternion (mj_kinematics, mjJNT_BALL) */
                double nq[4];
                quat_mul(nq, xquat, d->qpos + qa);
                memcpy(xquat, nq, sizeof nq);
                quat2mat(mat, xquat);
                mat_vec(t, mat, m->jnt_pos[j]);
                for (int k = 0; k < 3; k++) xpos[k] = d->xanchor[j][k] - t[k];
                continue;
            }
            double q = d->qpos[qa] - m->qpos0[qa];
            if (m->jnt_type[j] == ILQG_JNT_SLIDE) {
                for (int k = 0; k < 3; k++) xpos[k] += d->xaxis[j][k] * q;
            } else { /* hinge: rotate about the anchor */
                double ql[4], nq[4];
                axisangle2quat(ql, m->jnt_axis[j], q);
                quat_mul(nq, xquat, ql);
                memcpy(xquat, nq, sizeof nq);
                quat2mat(mat, xquat);
                mat_vec(t, mat, m->jnt_pos[j]);
                for (int k = 0; k < 3; k++) xpos[k] = d->xanchor[j][k] - t[k];
            }
        }
        quat_normalize(xquat);
        memcpy(d->xpos[b], xpos, sizeof xpos);
        memcpy(d->xquat[b], xquat, sizeof xquat);
        quat2mat(d->xmat[b], xquat);
        mat_vec(t, d->xmat[b], m->body_ipos[b]);
        for (int k = 0; k < 3; k++) d->xipos[b][k] = xpos[k] + t[k];
        FL(d, 120.0 + 90.0 * m->body_jntnum[b]);
    }
    for (int g = 0; g < m->ngeom; g++) {
        int b = m->geom_bodyid[g];
        double t[3], gm[9];
        mat_vec(t, d->xmat[b], m->geom_pos[g]);
        for (int k = 0; k < 3; k++) d->geom_xpos[g][k] = d->xpos[b][k] + t[k];
        quat2mat(gm, m->geom_quat[g]);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                d->geom_xmat[g][3 * r + c] = d->xmat[b][3 * r] * gm[c] + d->xmat[b][3 * r + 1] * gm[3 + c] + d->xmat[b][3 * r + 2] * gm[6 + c];
        FL(d, 15 + 24 + 45);
    }
}

static void com_pos(const ilqg_model* m, mjo_data* d) {
    /* mj_comPos: subtree centres of mass, body inertias and joint motion axes in a frame
       centred at the com of each kinematic tree */
    double mass[ILQG_MAXBODY];
    for (int b = 0; b < m->nbody; b++) {
        mass[b] = m->body_mass[b];
        for (int k = 0; k < 3; k++) d->subtree_com[b][k] = m->body_mass[b] * d->xipos[b][k];
    }
    for (int b = m->nbody - 1; b > 0; b--) {
        int p = m->body_parentid[b];
        mass[p] += mass[b];
        for (int k = 0; k < 3; k++) d->subtree_com[p][k] += d->subtree_com[b][k];
    }
    for (int b = 0; b < m->nbody; b++) {
        if (mass[b] < MJO_MINVAL) for (int k = 0; k < 3; k++) d->subtree_com[b][k] = d->xipos[b][k];
        else for (int k = 0; k < 3; k++) d->subtree_com[b][k] /= mass[b];
    }
    FL(d, 12.0 * m->nbody);
    for (int b = 1; b < m->nbody; b++) {
        const double* R = d->xmat[b];
        const double* in = m->body_inertia[b];
        double Ib[9] = {in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2]};
        double T[9], Iw[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) T[3 * r + c] = R[3 * r] * Ib[c] + R[3 * r + 1] * Ib[3 + c] + R[3 * r + 2] * Ib[6 + c];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) Iw[3 * r + c] = T[3 * r] * R[3 * c] + T[3 * r + 1] * R[3 * c + 1] + T[3 * r + 2] * R[3 * c + 2];
        double dif[3], ms = m->body_mass[b];
        const double* com = d->subtree_com[m->body_rootid[b]];
        for (int k = 0; k < 3; k++) dif[k] = d->xipos[b][k] - com[k];
        double dd = v3_dot(dif, dif);
        double* ci = d->cinert[b];
        ci[0] = Iw[0] + ms * (dd - dif[0] * dif[0]);
        ci[1] = Iw[4] + ms * (dd - dif[1] * dif[1]);
        ci[2] = Iw[8] + ms * (dd - dif[2] * dif[2]);
        ci[3] = Iw[1] - ms * dif[0] * dif[1];
        ci[4] = Iw[2] - ms * dif[0] * dif[2];
        ci[5] = Iw[5] - ms * dif[1] * dif[2];
        ci[6] = ms * dif[0]; ci[7] = ms * dif[1]; ci[8] = ms * dif[2];
        ci[9] = ms;
        FL(d, 90 + 30);
    }
    for (int j = 0; j < m->njnt; j++) {
        int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
        const double* com = d->subtree_com[m->body_rootid[b]];
        double off[3];
        for (int k = 0; k < 3; k++) off[k] = com[k] - d->xanchor[j][k];
        if (m->jnt_type[j] == ILQG_JNT_FREE) {
            for (int i = 0; i < 3; i++) {
                for (int k = 0; k < 6; k++) d->cdof[da + i][k] = 0;
                d->cdof[da + i][3 + i] = 1;
            }
            for (int i = 0; i < 3; i++) {
                double ax[3] = {d->xmat[b][i], d->xmat[b][3 + i], d->xmat[b][6 + i]};
                for (int k = 0; k < 3; k++) d->cdof[da + 3 + i][k] = ax[k];
                v3_cross(d->cdof[da + 3 + i] + 3, ax, off);
            }
            FL(d, 27);
        } else if (m->jnt_type[j] == ILQG_JNT_BALL) { /* the body's three axes about the anchor (mj_comPos, mjJNT_BALL) */
            for (int i = 0; i < 3; i++) {
                double ax[3] = {d->xmat[b][i], d->xmat[b][3 + i], d->xmat[b][6 + i]};
                for (int k = 0; k < 3; k++) d->cdof[da + i][k] = ax[k];
                v3_cross(d->cdof[da + i] + 3, ax, off);
            }
            FL(d, 27);
        } else if (m->jnt_type[j] == ILQG_JNT_SLIDE) {
            for (int k = 0; k < 3; k++) { d->cdof[da][k] = 0; d->cdof[da][3 + k] = d->xaxis[j][k]; }
        } else {
            for (int k = 0; k < 3; k++) d->cdof[da][k] = d->xaxis[j][k];
            v3_cross(d->cdof[da] + 3, d->xaxis[j], off);
            FL(d, 12);
        }
    }
}

/* spatial inertia (10 numbers about the tree com) times spatial motion vector [w; v] */
static void inert_vec(double* r, const double* i, const double* v) {
    r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] + (i[7] * v[5] - i[8] * v[4]);
    r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + (i[8] * v[3] - i[6] * v[5]);
    r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] + (i[6] * v[4] - i[7] * v[3]);
    r[3] = i[9] * v[3] + (v[1] * i[8] - v[2] * i[7]);
    r[4] = i[9] * v[4] + (v[2] * i[6] - v[0] * i[8]);
    r[5] = i[9] * v[5] + (v[0] * i[7] - v[1] * i[6]);
}

static void crb(const ilqg_model* m, mjo_data* d) {
    /* mj_crb: composite rigid body inertias, then M[i][j] = cdof_j . (crb_body(i) cdof_i) up the chain */
    int nv = m->nv;
    for (int b = 0; b < m->nbody; b++) memcpy(d->crb[b], d->cinert[b], sizeof d->crb[b]);
    for (int k = 0; k < 10; k++) d->crb[0][k] = 0;
    for (int b = m->nbody - 1; b > 0; b--) {
        int p = m->body_parentid[b];
        if (p > 0) for (int k = 0; k < 10; k++) d->crb[p][k] += d->crb[b][k];
    }
    FL(d, 10.0 * m->nbody);
    memset(d->qM, 0, sizeof(double) * nv * nv);
    for (int i = 0; i < nv; i++) {
        double buf[6];
        inert_vec(buf, d->crb[m->dof_bodyid[i]], d->cdof[i]);
        FL(d, 48);
        for (int j = i; j >= 0; j = m->dof_parentid[j]) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += d->cdof[j][k] * buf[k];
            d->qM[i * nv + j] = d->qM[j * nv + i] = s;
            FL(d, 12);
        }
        d->qM[i * nv + i] += m->dof_armature[i];
    }
}

/* dense Cholesky; returns 0 on success */
static int chol(double* L, const double* A, int n, mjo_data* d) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
            FL(d, 2.0 * j + 1);
            if (i == j) {
                if (s < MJO_MINVAL) s = MJO_MINVAL;
                L[i * n + i] = sqrt(s);
            } else
                L[i * n + j] = s / L[j * n + j];
        }
    return 0;
}
static void chol_solve(const double* L, double* x, int n, mjo_data* d) {
    for (int i = 0; i < n; i++) {
        double s = x[i];
        for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
        x[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
        x[i] = s / L[i * n + i];
    }
    FL(d, 2.0 * n * n + 2 * n);
}

/* translational Jacobian of a world point attached to `body` (mj_jac, translational half) */
static void jac_point(const ilqg_model* m, const mjo_data* d, double* jacp /*3 x nv*/, const double* point, int body) {
    int nv = m->nv;
    memset(jacp, 0, sizeof(double) * 3 * nv);
    if (body <= 0) return;
    const double* com = d->subtree_com[m->body_rootid[body]];
    double off[3] = {point[0] - com[0], point[1] - com[1], point[2] - com[2]};
    /* last dof of the nearest ancestor-or-self body that has dofs, then up the dof chain */
    int b = body;
    while (b > 0 && m->body_dofnum[b] == 0) b = m->body_parentid[b];
    if (b <= 0) return;
    for (int i = m->body_dofadr[b] + m->body_dofnum[b] - 1; i >= 0; i = m->dof_parentid[i]) {
        double c[3];
        v3_cross(c, d->cdof[i], off);
        for (int k = 0; k < 3; k++) jacp[k * nv + i] = d->cdof[i][3 + k] + c[k];
    }
}

static void make_frame(double* f) {
    /* mju_makeFrame: f[0..2] = x axis (normal), f[3..5] = optional y hint */
    v3_normalize(f);
    double* y = f + 3;
    if (sqrt(v3_dot(y, y)) < 0.5) {
        y[0] = y[1] = y[2] = 0;
        if (f[1] < 0.5 && f[1] > -0.5) y[1] = 1; else y[2] = 1;
    }
    double dp = v3_dot(f, y);
    for (int k = 0; k < 3; k++) y[k] -= dp * f[k];
    if (sqrt(v3_dot(y, y)) < 1e-12) { /* hint parallel to the normal: fall back to the default rule */
        y[0] = y[1] = y[2] = 0;
        if (f[1] < 0.5 && f[1] > -0.5) y[1] = 1; else y[2] = 1;
        dp = v3_dot(f, y);
        for (int k = 0; k < 3; k++) y[k] -= dp * f[k];
    }
    v3_normalize(y);
    v3_cross(f + 6, f, y);
}

static void add_contact(mjo_data* d, int pair, double dist, const double* pos, const double* normal, const double* yhint) {
    if (d->ncon >= MJO_MAXCON) return;
    mjo_contact* c = &d->contact[d->ncon++];
    c->pair = pair;
    c->dist = dist;
    memcpy(c->pos, pos, sizeof(double) * 3);
    memcpy(c->frame, normal, sizeof(double) * 3);
    if (yhint) memcpy(c->frame + 3, yhint, sizeof(double) * 3);
    else c->frame[3] = c->frame[4] = c->frame[5] = 0;
    make_frame(c->frame);
}

static void sphere_sphere(mjo_data* d, int pair, const double* p1, double r1, const double* p2, double r2, double margin) {
    double n[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    double len = sqrt(v3_dot(n, n));
    double dist = len - r1 - r2;
    FL(d, 12);
    if (dist > margin) return;
    if (len < MJO_MINVAL) { n[0] = 1; n[1] = 0; n[2] = 0; }
    else { n[0] /= len; n[1] /= len; n[2] /= len; }
    double pos[3];
    for (int k = 0; k < 3; k++) pos[k] = p1[k] + n[k] * (r1 + 0.5 * dist);
    add_contact(d, pair, dist, pos, n, NULL);
}

static void plane_sphere(mjo_data* d, int pair, const double* ppos, const double* pn, const double* c, double r, double margin,
                         const double* yhint) {
    double dif[3] = {c[0] - ppos[0], c[1] - ppos[1], c[2] - ppos[2]};
    double dist = v3_dot(dif, pn) - r;
    FL(d, 9);
    if (dist > margin) return;
    double pos[3];
    for (int k = 0; k < 3; k++) pos[k] = c[k] - pn[k] * (r + 0.5 * dist);
    add_contact(d, pair, dist, pos, pn, yhint);
}

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

static void capsule_capsule(mjo_data* d, int pair, const double* p1, const double* a1, double r1, double h1, const double* p2,
                            const double* a2, double r2, double h2, double margin) {
    /* closest points of the two axis segments, then a sphere-sphere test (mjc_CapsuleCapsule) */
    double dif[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    double mb = -v3_dot(a1, a2), u = -v3_dot(a1, dif), v = v3_dot(a2, dif);
    double det = 1.0 - mb * mb;
    FL(d, 25);
    if (fabs(det) >= 1e-12) {
        double x1 = (u - mb * v) / det, x2 = (v - mb * u) / det;
        if (x1 > h1) { x1 = h1; x2 = v - mb * h1; }
        else if (x1 < -h1) { x1 = -h1; x2 = v + mb * h1; }
        if (x2 > h2) { x2 = h2; x1 = clampd(u - mb * h2, -h1, h1); }
        else if (x2 < -h2) { x2 = -h2; x1 = clampd(u + mb * h2, -h1, h1); }
        double c1[3], c2[3];
        for (int k = 0; k < 3; k++) { c1[k] = p1[k] + a1[k] * x1; c2[k] = p2[k] + a2[k] * x2; }
        sphere_sphere(d, pair, c1, r1, c2, r2, margin);
    } else {
        /* parallel axes: test both ends of segment 1 against segment 2 and vice versa, keep <= 2 */
        int before = d->ncon;
        for (int s = -1; s <= 1 && d->ncon - before < 2; s += 2) {
            double c1[3], c2[3];
            for (int k = 0; k < 3; k++) c1[k] = p1[k] + a1[k] * s * h1;
            double t = (c1[0] - p2[0]) * a2[0] + (c1[1] - p2[1]) * a2[1] + (c1[2] - p2[2]) * a2[2];
            if (t < -h2 || t > h2) continue;
            for (int k = 0; k < 3; k++) c2[k] = p2[k] + a2[k] * t;
            sphere_sphere(d, pair, c1, r1, c2, r2, margin);
        }
        for (int s = -1; s <= 1 && d->ncon - before < 2; s += 2) {
            double c1[3], c2[3];
            for (int k = 0; k < 3; k++) c2[k] = p2[k] + a2[k] * s * h2;
            double t = (c2[0] - p1[0]) * a1[0] + (c2[1] - p1[1]) * a1[1] + (c2[2] - p1[2]) * a1[2];
            if (t <= -h1 || t >= h1) continue;
            for (int k = 0; k < 3; k++) c1[k] = p1[k] + a1[k] * t;
            sphere_sphere(d, pair, c1, r1, c2, r2, margin);
        }
        if (d->ncon == before) { /* segments do not overlap along the axis: nearest end points */
            double best = 1e300, b1[3] = {0, 0, 0}, b2[3] = {0, 0, 0};
            for (int s = -1; s <= 1; s += 2)
                for (int t = -1; t <= 1; t += 2) {
                    double c1[3], c2[3], dd = 0;
                    for (int k = 0; k < 3; k++) { c1[k] = p1[k] + a1[k] * s * h1; c2[k] = p2[k] + a2[k] * t * h2; dd += (c1[k] - c2[k]) * (c1[k] - c2[k]); }
                    if (dd < best) { best = dd; memcpy(b1, c1, sizeof c1); memcpy(b2, c2, sizeof c2); }
                }
            sphere_sphere(d, pair, b1, r1, b2, r2, margin);
        }
    }
}

static void collision(const ilqg_model* m, mjo_data* d) {
    d->ncon = 0;
    for (int p = 0; p < m->npair; p++) {
        int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
        int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
        /* the solver sees includemargin (= margin - gap; gap is 0 in this subset) */
        double margin = m->pair_margin[p];
        const double *x1 = d->geom_xpos[g1], *x2 = d->geom_xpos[g2];
        const double *M1 = d->geom_xmat[g1], *M2 = d->geom_xmat[g2];
        double ax1[3] = {M1[2], M1[5], M1[8]}, ax2[3] = {M2[2], M2[5], M2[8]};
        if (t1 == ILQG_GEOM_PLANE && t2 == ILQG_GEOM_SPHERE) {
            plane_sphere(d, p, x1, ax1, x2, m->geom_size[g2][0], margin, NULL);
        } else if (t1 == ILQG_GEOM_PLANE && t2 == ILQG_GEOM_CAPSULE) {
            /* mjc_PlaneCapsule: the two end spheres, contact frames aligned with the capsule axis */
            double r = m->geom_size[g2][0], h = m->geom_size[g2][1], e[3];
            for (int k = 0; k < 3; k++) e[k] = x2[k] + ax2[k] * h;
            plane_sphere(d, p, x1, ax1, e, r, margin, ax2);
            for (int k = 0; k < 3; k++) e[k] = x2[k] - ax2[k] * h;
            plane_sphere(d, p, x1, ax1, e, r, margin, ax2);
        } else if (t1 == ILQG_GEOM_SPHERE && t2 == ILQG_GEOM_SPHERE) {
            sphere_sphere(d, p, x1, m->geom_size[g1][0], x2, m->geom_size[g2][0], margin);
        } else if (t1 == ILQG_GEOM_SPHERE && t2 == ILQG_GEOM_CAPSULE) {
            double h = m->geom_size[g2][1];
            double t = clampd((x1[0] - x2[0]) * ax2[0] + (x1[1] - x2[1]) * ax2[1] + (x1[2] - x2[2]) * ax2[2], -h, h);
            double c[3];
            for (int k = 0; k < 3; k++) c[k] = x2[k] + ax2[k] * t;
            FL(d, 12);
            sphere_sphere(d, p, x1, m->geom_size[g1][0], c, m->geom_size[g2][0], margin);
        } else if (t1 == ILQG_GEOM_CAPSULE && t2 == ILQG_GEOM_CAPSULE) {
            capsule_capsule(d, p, x1, ax1, m->geom_size[g1][0], m->geom_size[g1][1], x2, ax2, m->geom_size[g2][0], m->geom_size[g2][1], margin);
        }
    }
}

static double impedance(const double* solimp, double pos, double margin) {
    /* getimpedance: d(r) of MuJoCo's solimp; dmin/dmax clamped to [1e-4, 0.9999] */
    double dmin = clampd(solimp[0], 1e-4, 0.9999), dmax = clampd(solimp[1], 1e-4, 0.9999);
    double width = solimp[2], mid = clampd(solimp[3], 1e-4, 0.9999), power = solimp[4] < 1 ? 1 : solimp[4];
    if (dmin == dmax || width <= MJO_MINVAL) return 0.5 * (dmin + dmax);
    double x = fabs(pos - margin) / width;
    if (x >= 1) return dmax;
    if (x <= 0) return dmin;
    double y;
    if (power == 1) y = x;
    else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
    else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
    return dmin + y * (dmax - dmin);
}

static int add_row(mjo_data* d, int nv) {
    if (d->nefc >= MJO_MAXEFC) return -1;
    int r = d->nefc++;
    memset(d->efc_J + (size_t)r * nv, 0, sizeof(double) * nv);
    return r;
}

static void set_row_params(const ilqg_model* m, mjo_data* d, int r, const double* solref, const double* solimp, double pos,
                           double margin, double diagApprox) {
    /* mj_makeImpedance for one row: impedance, regulariser R, stiffness K and damping B */
    double tc = solref[0], dr = solref[1];
    if (tc < 2 * m->timestep) tc = 2 * m->timestep; /* refsafe */
    double dmax = clampd(solimp[1], 1e-4, 0.9999);
    double imp = impedance(solimp, pos, margin);
    double R = (1 - imp) / imp * diagApprox;
    if (R < MJO_MINVAL) R = MJO_MINVAL;
    double kk = dmax * dmax * tc * tc * dr * dr, bb = dmax * tc;
    d->efc_pos[r] = pos;
    d->efc_margin[r] = margin;
    d->efc_diagApprox[r] = diagApprox;
    d->efc_R[r] = R;
    d->efc_KBIP[r][0] = 1.0 / (kk < MJO_MINVAL ? MJO_MINVAL : kk);
    d->efc_KBIP[r][1] = 2.0 / (bb < MJO_MINVAL ? MJO_MINVAL : bb);
    d->efc_KBIP[r][2] = imp;
    d->efc_KBIP[r][3] = 0;
    FL(d, 20);
}

static void make_constraint(const ilqg_model* m, mjo_data* d) {
    int nv = m->nv;
    d->nefc = 0;
    /* joint limits (mj_instantiateLimit): slide and hinge */
    for (int j = 0; j < m->njnt; j++) {
        if (!m->jnt_limited[j] || m->jnt_type[j] == ILQG_JNT_FREE || m->jnt_type[j] == ILQG_JNT_BALL) continue;
        double value = d->qpos[m->jnt_qposadr[j]];
        for (int side = -1; side <= 1; side += 2) {
            double dist = side * (m->jnt_range[j][(side + 1) / 2] - value);
            if (dist < m->jnt_margin[j]) {
                int r = add_row(d, nv);
                if (r < 0) return;
                d->efc_J[(size_t)r * nv + m->jnt_dofadr[j]] = -side;
                set_row_params(m, d, r, m->jnt_solref[j], m->jnt_solimp[j], dist, m->jnt_margin[j],
                               m->dof_invweight0[m->jnt_dofadr[j]]);
            }
        }
    }
    /* contacts (mj_instantiateContact): frictionless = 1 row, pyramidal condim 3 = 4 rows */
    double j1[3 * ILQG_MAXV], j2[3 * ILQG_MAXV], jc[3 * ILQG_MAXV];
    for (int c = 0; c < d->ncon; c++) {
        const mjo_contact* con = &d->contact[c];
        int p = con->pair;
        int b1 = m->geom_bodyid[m->pair_geom1[p]], b2 = m->geom_bodyid[m->pair_geom2[p]];
        jac_point(m, d, j1, con->pos, b1);
        jac_point(m, d, j2, con->pos, b2);
        for (int r = 0; r < 3; r++)
            for (int i = 0; i < nv; i++) {
                double s = 0;
                for (int k = 0; k < 3; k++) s += con->frame[3 * r + k] * (j2[k * nv + i] - j1[k * nv + i]);
                jc[r * nv + i] = s;
            }
        FL(d, 2 * 9.0 * nv * 2 + 9.0 * nv);
        double tran = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
        if (tran < MJO_MINVAL) tran = MJO_MINVAL;
        double margin = m->pair_margin[p];
        if (m->pair_condim[p] == 1) {
            int r = add_row(d, nv);
            if (r < 0) return;
            memcpy(d->efc_J + (size_t)r * nv, jc, sizeof(double) * nv);
            set_row_params(m, d, r, m->pair_solref[p], m->pair_solimp[p], con->dist, margin, tran);
        } else {
            double mu = m->pair_friction[p];
            int first = d->nefc;
            for (int k = 0; k < 4; k++) {
                int r = add_row(d, nv);
                if (r < 0) return;
                const double* jt = jc + (1 + k / 2) * nv;
                double sg = (k % 2) ? -mu : mu;
                for (int i = 0; i < nv; i++) d->efc_J[(size_t)r * nv + i] = jc[i] + sg * jt[i];
                set_row_params(m, d, r, m->pair_solref[p], m->pair_solimp[p], con->dist, margin, tran * (1 + mu * mu));
            }
            FL(d, 8.0 * nv);
            /* pyramidal regulariser (impratio = 1): every facet gets 2 mu^2 R_first */
            double Rpy = 2 * mu * mu * d->efc_R[first];
            if (Rpy < MJO_MINVAL) Rpy = MJO_MINVAL;
            for (int k = 0; k < 4; k++) d->efc_R[first + k] = Rpy;
        }
    }
    for (int r = 0; r < d->nefc; r++) d->efc_D[r] = 1.0 / d->efc_R[r];
}

void mjo_fwd_position(const ilqg_model* m, mjo_data* d) {
    kinematics(m, d);
    com_pos(m, d);
    crb(m, d);
    chol(d->qL, d->qM, m->nv, d);
    collision(m, d);
    make_constraint(m, d);
}

/* ------------------------------------------------------------------ velocity stage */
static void cross_motion(double* r, const double* vel, const double* v) {
    double a[3], b[3], c[3];
    v3_cross(a, vel, v);         /* w x w2 */
    v3_cross(b, vel, v + 3);     /* w x v2 */
    v3_cross(c, vel + 3, v);     /* v x w2 */
    for (int k = 0; k < 3; k++) { r[k] = a[k]; r[3 + k] = b[k] + c[k]; }
}
static void cross_force(double* r, const double* vel, const double* f) {
    double a[3], b[3], c[3];
    v3_cross(a, vel, f);         /* w x tau */
    v3_cross(b, vel + 3, f + 3); /* v x f */
    v3_cross(c, vel, f + 3);     /* w x f */
    for (int k = 0; k < 3; k++) { r[k] = a[k] + b[k]; r[3 + k] = c[k]; }
}

static void com_vel(const ilqg_model* m, mjo_data* d) {
    /* mj_comVel */
    for (int k = 0; k < 6; k++) d->cvel[0][k] = 0;
    for (int b = 1; b < m->nbody; b++) {
        double cvel[6];
        memcpy(cvel, d->cvel[m->body_parentid[b]], sizeof cvel);
        for (int jj = 0; jj < m->body_jntnum[b]; jj++) {
            int j = m->body_jntadr[b] + jj, da = m->jnt_dofadr[j];
            if (m->jnt_type[j] == ILQG_JNT_FREE) {
                for (int i = 0; i < 3; i++) {
                    for (int k = 0; k < 6; k++) d->cdof_dot[da + i][k] = 0;
                    for (int k = 0; k < 6; k++) cvel[k] += d->cdof[da + i][k] * d->qvel[da + i];
                }
                for (int i = 3; i < 6; i++) cross_motion(d->cdof_dot[da + i], cvel, d->cdof[da + i]);
                for (int i = 3; i < 6; i++)
                    for (int k = 0; k < 6; k++) cvel[k] += d->cdof[da + i][k] * d->qvel[da + i];
                FL(d, 72 + 3 * 36);
            } else if (m->jnt_type[j] == ILQG_JNT_BALL) { /* all three axes turn with the velocity before the joint (mj_comVel) */
                for (int i = 0; i < 3; i++) cross_motion(d->cdof_dot[da + i], cvel, d->cdof[da + i]);
                for (int i = 0; i < 3; i++)
                    for (int k = 0; k < 6; k++) cvel[k] += d->cdof[da + i][k] * d->qvel[da + i];
                FL(d, 36 + 3 * 36);
            } else {
                cross_motion(d->cdof_dot[da], cvel, d->cdof[da]);
                for (int k = 0; k < 6; k++) cvel[k] += d->cdof[da][k] * d->qvel[da];
                FL(d, 12 + 36);
            }
        }
        memcpy(d->cvel[b], cvel, sizeof cvel);
    }
}

static void passive(const ilqg_model* m, mjo_data* d) {
    for (int i = 0; i < m->nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
    for (int j = 0; j < m->njnt; j++) {
        if (m->jnt_type[j] == ILQG_JNT_FREE || m->jnt_type[j] == ILQG_JNT_BALL || m->jnt_stiffness[j] == 0) continue;
        int qa = m->jnt_qposadr[j];
        d->qfrc_passive[m->jnt_dofadr[j]] -= m->jnt_stiffness[j] * (d->qpos[qa] - m->qpos_spring[qa]);
    }
    FL(d, 4.0 * m->nv);
}

static void rne_bias(const ilqg_model* m, mjo_data* d) {
    /* mj_rne with flg_acc = 0: Coriolis, centrifugal and gravity forces */
    double cacc[ILQG_MAXBODY][6], cfrc[ILQG_MAXBODY][6];
    for (int k = 0; k < 3; k++) { cacc[0][k] = 0; cacc[0][3 + k] = -m->gravity[k]; }
    for (int k = 0; k < 6; k++) cfrc[0][k] = 0;
    for (int b = 1; b < m->nbody; b++) {
        memcpy(cacc[b], cacc[m->body_parentid[b]], sizeof cacc[b]);
        for (int i = m->body_dofadr[b]; i < m->body_dofadr[b] + m->body_dofnum[b]; i++)
            for (int k = 0; k < 6; k++) cacc[b][k] += d->cdof_dot[i][k] * d->qvel[i];
        double ia[6], iv[6], cf[6];
        inert_vec(ia, d->cinert[b], cacc[b]);
        inert_vec(iv, d->cinert[b], d->cvel[b]);
        cross_force(cf, d->cvel[b], iv);
        for (int k = 0; k < 6; k++) cfrc[b][k] = ia[k] + cf[k];
        FL(d, 12.0 * m->body_dofnum[b] + 48 * 2 + 36 + 6);
    }
    for (int b = m->nbody - 1; b > 0; b--) {
        int p = m->body_parentid[b];
        for (int k = 0; k < 6; k++) cfrc[p][k] += cfrc[b][k];
    }
    for (int i = 0; i < m->nv; i++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += d->cdof[i][k] * cfrc[m->dof_bodyid[i]][k];
        d->qfrc_bias[i] = s;
    }
    FL(d, 6.0 * m->nbody + 12.0 * m->nv);
}

static void reference_constraint(const ilqg_model* m, mjo_data* d) {
    /* mj_referenceConstraint: aref = -B vel - K imp (pos - margin) */
    int nv = m->nv;
    for (int r = 0; r < d->nefc; r++) {
        double s = 0;
        for (int i = 0; i < nv; i++) s += d->efc_J[(size_t)r * nv + i] * d->qvel[i];
        d->efc_vel[r] = s;
        d->efc_aref[r] = -d->efc_KBIP[r][1] * s - d->efc_KBIP[r][0] * d->efc_KBIP[r][2] * (d->efc_pos[r] - d->efc_margin[r]);
    }
    FL(d, d->nefc * (2.0 * nv + 6));
}

void mjo_fwd_velocity(const ilqg_model* m, mjo_data* d) {
    com_vel(m, d);
    passive(m, d);
    reference_constraint(m, d);
    rne_bias(m, d);
}

/* ------------------------------------------------------------------ acceleration stage */
void mjo_fwd_actuation(const ilqg_model* m, mjo_data* d) {
    for (int i = 0; i < m->nv; i++) d->qfrc_actuator[i] = 0;
    for (int u = 0; u < m->nu; u++) {
        double c = d->ctrl[u];
        if (m->act_ctrllimited[u]) c = clampd(c, m->act_ctrlrange[u][0], m->act_ctrlrange[u][1]);
        d->qfrc_actuator[m->act_dofid[u]] += m->act_gear[u] * c;
    }
    FL(d, 2.0 * m->nu);
}

void mjo_fwd_acceleration(const ilqg_model* m, mjo_data* d) {
    int nv = m->nv;
    for (int i = 0; i < nv; i++) {
        d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_applied[i] + d->qfrc_actuator[i];
        d->qacc_smooth[i] = d->qfrc_smooth[i];
    }
    FL(d, 3.0 * nv);
    chol_solve(d->qL, d->qacc_smooth, nv, d);
}

/* cost of the constraint problem at acceleration a (used for warm-start selection) */
static double total_cost(const ilqg_model* m, mjo_data* d, const double* a) {
    int nv = m->nv;
    double cost = 0;
    for (int r = 0; r < d->nefc; r++) {
        double jar = -d->efc_aref[r];
        for (int i = 0; i < nv; i++) jar += d->efc_J[(size_t)r * nv + i] * a[i];
        if (jar < 0) cost += 0.5 * d->efc_D[r] * jar * jar;
    }
    for (int i = 0; i < nv; i++) {
        double Ma = 0;
        for (int k = 0; k < nv; k++) Ma += d->qM[i * nv + k] * a[k];
        cost += 0.5 * (Ma - d->qfrc_smooth[i]) * (a[i] - d->qacc_smooth[i]);
    }
    FL(d, d->nefc * (2.0 * nv + 4) + nv * (2.0 * nv + 4));
    return cost;
}

typedef struct newton_ctx {
    int nv, nefc;
    double Ma[ILQG_MAXV], grad[ILQG_MAXV], Mgrad[ILQG_MAXV], search[ILQG_MAXV], Mv[ILQG_MAXV];
    double jar[MJO_MAXEFC], jv[MJO_MAXEFC];
    double H[ILQG_MAXV * ILQG_MAXV], L[ILQG_MAXV * ILQG_MAXV];
    double cost;
} newton_ctx;

static void newton_update(const ilqg_model* m, mjo_data* d, newton_ctx* c) {
    int nv = c->nv, ne = c->nefc;
    double cost = 0;
    for (int i = 0; i < nv; i++) d->qfrc_constraint[i] = 0;
    memcpy(c->H, d->qM, sizeof(double) * nv * nv);
    for (int r = 0; r < ne; r++) {
        const double* J = d->efc_J + (size_t)r * nv;
        if (c->jar[r] < 0) {
            double D = d->efc_D[r];
            double f = -D * c->jar[r];
            d->efc_force[r] = f;
            cost += 0.5 * D * c->jar[r] * c->jar[r];
            for (int i = 0; i < nv; i++) d->qfrc_constraint[i] += J[i] * f;
            for (int i = 0; i < nv; i++) {
                double t = D * J[i];
                for (int k = 0; k <= i; k++) c->H[i * nv + k] += t * J[k];
            }
            FL(d, 5 + 2.0 * nv + nv + nv * (nv + 1.0));
        } else
            d->efc_force[r] = 0;
    }
    for (int i = 0; i < nv; i++)
        for (int k = i + 1; k < nv; k++) c->H[i * nv + k] = c->H[k * nv + i];
    for (int i = 0; i < nv; i++) {
        cost += 0.5 * (c->Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
        c->grad[i] = c->Ma[i] - d->qfrc_smooth[i] - d->qfrc_constraint[i];
        c->Mgrad[i] = c->grad[i];
    }
    FL(d, 7.0 * nv);
    c->cost = cost;
    chol(c->L, c->H, nv, d);
    chol_solve(c->L, c->Mgrad, nv, d);
}

/* exact minimiser of the convex piecewise-quadratic cost along `search` (safeguarded Newton on
   the derivative; the solver path is builder-defined — the minimiser it converges to is not) */
static double newton_linesearch(const ilqg_model* m, mjo_data* d, newton_ctx* c) {
    int nv = c->nv, ne = c->nefc;
    double g1 = 0, g2 = 0;
    for (int i = 0; i < nv; i++) {
        double s = 0;
        for (int k = 0; k < nv; k++) s += d->qM[i * nv + k] * c->search[k];
        c->Mv[i] = s;
    }
    for (int i = 0; i < nv; i++) {
        g1 += c->search[i] * (c->Ma[i] - d->qfrc_smooth[i]);
        g2 += c->search[i] * c->Mv[i];
    }
    for (int r = 0; r < ne; r++) {
        double s = 0;
        for (int i = 0; i < nv; i++) s += d->efc_J[(size_t)r * nv + i] * c->search[i];
        c->jv[r] = s;
    }
    FL(d, 2.0 * nv * nv + 5.0 * nv + 2.0 * ne * nv);
    double alpha = 0, lo = 0, hi = INFINITY;
    for (int it = 0; it < m->ls_iterations; it++) {
        double d1 = g1 + g2 * alpha, d2 = g2;
        for (int r = 0; r < ne; r++) {
            double x = c->jar[r] + alpha * c->jv[r];
            if (x < 0) {
                double t = d->efc_D[r] * c->jv[r];
                d1 += t * x;
                d2 += t * c->jv[r];
            }
        }
        FL(d, 4 + 6.0 * ne);
        if (it == 0 && d1 >= 0) return 0; /* not a descent direction */
        if (d1 == 0) break;
        if (d1 < 0) lo = alpha; else hi = alpha;
        if (d2 < MJO_MINVAL) break;
        double step = -d1 / d2;
        double an = alpha + step;
        if (!(an > lo && an < hi)) an = isinf(hi) ? 2 * alpha + 1 : 0.5 * (lo + hi);
        if (fabs(an - alpha) <= 1e-14 * fabs(an)) { alpha = an; break; }
        alpha = an;
    }
    return alpha;
}

static void newton_solve(const ilqg_model* m, mjo_data* d, int maxiter, double tol) {
    static _Thread_local newton_ctx ctx;
    newton_ctx* c = &ctx;
    int nv = m->nv, ne = d->nefc;
    c->nv = nv; c->nefc = ne;
    double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
    for (int i = 0; i < nv; i++) {
        double s = 0;
        for (int k = 0; k < nv; k++) s += d->qM[i * nv + k] * d->qacc[k];
        c->Ma[i] = s;
    }
    for (int r = 0; r < ne; r++) {
        double s = -d->efc_aref[r];
        for (int i = 0; i < nv; i++) s += d->efc_J[(size_t)r * nv + i] * d->qacc[i];
        c->jar[r] = s;
    }
    FL(d, 2.0 * nv * nv + 2.0 * ne * nv);
    newton_update(m, d, c);
    for (int i = 0; i < nv; i++) c->search[i] = -c->Mgrad[i];
    int iter = 0;
    while (iter < maxiter) {
        double alpha = newton_linesearch(m, d, c);
        if (alpha == 0) break;
        for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * c->search[i]; c->Ma[i] += alpha * c->Mv[i]; }
        for (int r = 0; r < ne; r++) c->jar[r] += alpha * c->jv[r];
        FL(d, 4.0 * nv + 2.0 * ne);
        double old = c->cost;
        newton_update(m, d, c);
        for (int i = 0; i < nv; i++) c->search[i] = -c->Mgrad[i];
        iter++;
        double gn = 0;
        for (int i = 0; i < nv; i++) gn += c->grad[i] * c->grad[i];
        double improvement = scale * (old - c->cost), gradient = scale * sqrt(gn);
        if (improvement < tol || gradient < tol) break;
    }
    d->solver_iter = iter;
}

void mjo_fwd_constraint(const ilqg_model* m, mjo_data* d, int iterations, double tolerance) {
    int nv = m->nv;
    if (d->nefc == 0) {
        memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
        memcpy(d->qacc_warmstart, d->qacc_smooth, sizeof(double) * nv);
        memset(d->qfrc_constraint, 0, sizeof(double) * nv);
        d->solver_iter = 0;
        return;
    }
    /* warm start: the cheaper of qacc_warmstart and qacc_smooth */
    double cw = total_cost(m, d, d->qacc_warmstart);
    double cs = total_cost(m, d, d->qacc_smooth);
    if (cw < cs) memcpy(d->qacc, d->qacc_warmstart, sizeof(double) * nv);
    else memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    newton_solve(m, d, iterations, tolerance);
    memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
}

/* ------------------------------------------------------------------ drivers */
void mjo_forward_skip(const ilqg_model* m, mjo_data* d, int skipstage, int iterations, double tolerance) {
    if (skipstage < ILQG_STAGE_POS) mjo_fwd_position(m, d);
    if (skipstage < ILQG_STAGE_VEL) mjo_fwd_velocity(m, d);
    mjo_fwd_actuation(m, d);
    mjo_fwd_acceleration(m, d);
    mjo_fwd_constraint(m, d, iterations, tolerance);
}
void mjo_forward(const ilqg_model* m, mjo_data* d) { mjo_forward_skip(m, d, ILQG_STAGE_NONE, m->iterations, m->tolerance); }

static void euler(const ilqg_model* m, mjo_data* d) {
    /* mj_Euler: joint damping integrated implicitly, (M + h diag(b)) a' = f_smooth + f_constraint */
    int nv = m->nv, damped = 0;
    double h = m->timestep;
    double a[ILQG_MAXV];
    for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) damped = 1;
    if (damped) {
        static _Thread_local double A[ILQG_MAXV * ILQG_MAXV], L[ILQG_MAXV * ILQG_MAXV];
        memcpy(A, d->qM, sizeof(double) * nv * nv);
        for (int i = 0; i < nv; i++) {
            A[i * nv + i] += h * m->dof_damping[i];
            a[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
        }
        chol(L, A, nv, d);
        chol_solve(L, a, nv, d);
    } else
        memcpy(a, d->qacc, sizeof(double) * nv);
    for (int i = 0; i < nv; i++) d->qvel[i] += h * a[i];
    mjo_integrate_pos(m, d->qpos, d->qvel, h);
    d->time += h;
    FL(d, 6.0 * nv);
}

static void rk4(const ilqg_model* m, mjo_data* d) {
    /* mj_RungeKutta(4); stage 0 is the mj_forward already done by mjo_step */
    static const double A[3][3] = {{0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 1}};
    static const double B[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
    static const double C[3] = {0.5, 0.5, 1.0};
    int nq = m->nq, nv = m->nv;
    double h = m->timestep, t0 = d->time;
    double q0[ILQG_MAXQ], v0[ILQG_MAXV], X[4][ILQG_MAXV], F[4][ILQG_MAXV], dX[ILQG_MAXV], dF[ILQG_MAXV];
    memcpy(q0, d->qpos, sizeof(double) * nq);
    memcpy(v0, d->qvel, sizeof(double) * nv);
    memcpy(X[0], d->qvel, sizeof(double) * nv);
    memcpy(F[0], d->qacc, sizeof(double) * nv);
    for (int i = 1; i < 4; i++) {
        for (int k = 0; k < nv; k++) {
            dX[k] = 0; dF[k] = 0;
            for (int j = 0; j < i; j++) { dX[k] += A[i - 1][j] * X[j][k]; dF[k] += A[i - 1][j] * F[j][k]; }
        }
        memcpy(d->qpos, q0, sizeof(double) * nq);
        mjo_integrate_pos(m, d->qpos, dX, h);
        for (int k = 0; k < nv; k++) d->qvel[k] = v0[k] + h * dF[k];
        d->time = t0 + h * C[i - 1];
        mjo_forward_skip(m, d, ILQG_STAGE_NONE, m->iterations, m->tolerance);
        memcpy(X[i], d->qvel, sizeof(double) * nv);
        memcpy(F[i], d->qacc, sizeof(double) * nv);
    }
    for (int k = 0; k < nv; k++) {
        dX[k] = 0; dF[k] = 0;
        for (int j = 0; j < 4; j++) { dX[k] += B[j] * X[j][k]; dF[k] += B[j] * F[j][k]; }
    }
    memcpy(d->qpos, q0, sizeof(double) * nq);
    for (int k = 0; k < nv; k++) d->qvel[k] = v0[k] + h * dF[k];
    mjo_integrate_pos(m, d->qpos, dX, h);
    d->time = t0 + h;
    FL(d, 40.0 * nv);
}

void mjo_step(const ilqg_model* m, mjo_data* d) {
    mjo_forward(m, d);
    if (m->integrator == ILQG_INT_RK4) rk4(m, d);
    else euler(m, d);
}

double mjo_energy(const ilqg_model* m, mjo_data* d) {
    /* 0.5 v'Mv - sum m g.x (needs the position stage to be current) */
    int nv = m->nv;
    double e = 0;
    for (int i = 0; i < nv; i++)
        for (int k = 0; k < nv; k++) e += 0.5 * d->qvel[i] * d->qM[i * nv + k] * d->qvel[k];
    for (int b = 1; b < m->nbody; b++) e -= m->body_mass[b] * v3_dot(m->gravity, d->xipos[b]);
    return e;
}
