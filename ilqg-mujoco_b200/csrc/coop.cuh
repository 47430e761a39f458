// coop.cuh — warp-cooperative fp64 forward dynamics and FD linearisation for models that do not fit the
// thread-per-rollout register budget (humanoid: nv = 27, 150 perturbed evaluations per knot, 161 collision pairs).
//
// One WARP per perturbed rollout (north star (3): "one warp or CTA per perturbed rollout"): the rollout's whole
// mjData-equivalent — frames, spatial inertias, motion axes, the dense mass matrix and its Cholesky factor, the contact
// list, the constraint Jacobian, the Newton Hessian — lives in that warp's slice of SHARED memory; lanes stride over
// independent items (bodies of one tree level, dofs, geoms, collision pairs, constraint rows, matrix entries) and meet at
// __syncwarp().  The model is read from a GModel block in global memory (uniform or lane-strided reads, L1/L2 resident).
// Unlike dyn.cuh nothing is specialised on the kinematic tree: this path serves any model of the supported subset.
//
// Replaces, for such models, the same reference calls as dyn.cuh: mj_forward / mj_forwardSkip inside
// /root/reference/src/mjderivative.cpp:64-198 (incl. the quaternion tangent perturbation :152-169,187-192).
#pragma once
#include <cuda_runtime.h>

#include "../../include/ilqg_b200.h"
#include "dyn.cuh"

namespace ilqg {

#define COOP_MAXCON 24
#define COOP_MAXEFC 72
#define COOP_MAXLEVEL 16

// flat model + host-precomputed helpers, uploaded once per handle
struct GModel {
    ilqg_model m;
    double geom_axis[ILQG_MAXGEOM][3];
    double jnt_K[ILQG_MAXJNT], jnt_B[ILQG_MAXJNT], jnt_imp[ILQG_MAXJNT];
    double pair_K[ILQG_MAXPAIR], pair_B[ILQG_MAXPAIR], pair_imp[ILQG_MAXPAIR];
    unsigned body_dofmask[ILQG_MAXBODY];   // bit i set: dof i moves this body (nv <= 32)
    int nlevel, level_start[COOP_MAXLEVEL + 1], level_body[ILQG_MAXBODY];  // bodies grouped by tree depth
    int any_damping;
};

inline bool gmodel_from_tables(const ilqg_model& s, GModel& g) {
    if (s.nv > 32 || s.nbody > ILQG_MAXBODY) return false;
    g.m = s;
    for (int k = 0; k < s.ngeom; k++) {
        const double* q = s.geom_quat[k];
        g.geom_axis[k][0] = 2 * (q[1] * q[3] + q[0] * q[2]);
        g.geom_axis[k][1] = 2 * (q[2] * q[3] - q[0] * q[1]);
        g.geom_axis[k][2] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
    }
    for (int j = 0; j < s.njnt; j++) host_row_consts(s.timestep, s.jnt_solref[j], s.jnt_solimp[j], g.jnt_K[j], g.jnt_B[j], g.jnt_imp[j]);
    for (int p = 0; p < s.npair; p++) host_row_consts(s.timestep, s.pair_solref[p], s.pair_solimp[p], g.pair_K[p], g.pair_B[p], g.pair_imp[p]);
    int depth[ILQG_MAXBODY] = {0}, maxd = 0;
    for (int b = 1; b < s.nbody; b++) { depth[b] = depth[s.body_parentid[b]] + 1; if (depth[b] > maxd) maxd = depth[b]; }
    if (maxd > COOP_MAXLEVEL) return false;
    g.nlevel = maxd;
    int pos = 0;
    for (int L = 1; L <= maxd; L++) {
        g.level_start[L - 1] = pos;
        for (int b = 1; b < s.nbody; b++) if (depth[b] == L) g.level_body[pos++] = b;
    }
    g.level_start[maxd] = pos;
    for (int b = 0; b < s.nbody; b++) {
        unsigned mk = 0;
        for (int a = b; a > 0; a = s.body_parentid[a])
            for (int i = s.body_dofadr[a]; i < s.body_dofadr[a] + s.body_dofnum[a]; i++) mk |= 1u << i;
        g.body_dofmask[b] = mk;
    }
    g.any_damping = 0;
    for (int i = 0; i < s.nv; i++) g.any_damping |= s.dof_damping[i] > 0;
    return true;
}

// ------------------------------------------------------------------ per-warp shared-memory slice
struct CoopMem {
    double *q, *v, *u, *warm;
    double *xpos, *xquat, *xmat, *xipos, *anchor, *axis, *gpos, *gax, *com;
    double *cinert, *crb, *cdof, *cdofdot, *cvel, *cacc, *cfrc;
    double *M, *H;                       // nv x nv row-major (H doubles as Cholesky workspace)
    double *fs, *as, *fc, *qacc, *Ma, *grad, *search, *Mv;
    double *cdist, *cpos, *cframe;       // contacts
    int* cpair;
    double *J, *D, *aref, *jar, *jv;     // constraint rows
    int ncon, nefc;
};

__host__ __device__ inline size_t coop_doubles(int nq, int nv, int nu, int nb, int nj, int ng) {
    size_t n = nq + 2 * (size_t)nv + nu;                                   // q v u warm
    n += (size_t)nb * (3 + 4 + 9 + 3) + (size_t)nj * 6 + (size_t)ng * 6 + (size_t)nb * 3;  // frames, anchors, geoms, com
    n += (size_t)nb * 20 + (size_t)nv * 12 + (size_t)nb * 18;               // cinert crb, cdof cdofdot, cvel cacc cfrc
    n += 2 * (size_t)nv * nv + 8 * (size_t)nv;                              // M H + 8 vectors
    n += (size_t)COOP_MAXCON * 13;                                           // contacts
    n += (size_t)COOP_MAXEFC * nv + 4 * (size_t)COOP_MAXEFC;                 // J D aref jar jv
    return n;
}
__host__ __device__ inline size_t coop_bytes_per_warp(const ilqg_model& m) {
    size_t d = coop_doubles(m.nq, m.nv, m.nu, m.nbody, m.njnt, m.ngeom);
    return d * sizeof(double) + ((COOP_MAXCON * sizeof(int) + 15) / 16) * 16;
}

DEV void coop_carve(CoopMem& w, double* base, const ilqg_model& m) {
    const int nq = m.nq, nv = m.nv, nu = m.nu, nb = m.nbody, nj = m.njnt, ng = m.ngeom;
    double* p = base;
    auto take = [&](size_t n) { double* r = p; p += n; return r; };
    w.q = take(nq); w.v = take(nv); w.u = take(nu); w.warm = take(nv);
    w.xpos = take(nb * 3); w.xquat = take(nb * 4); w.xmat = take(nb * 9); w.xipos = take(nb * 3);
    w.anchor = take(nj * 3); w.axis = take(nj * 3); w.gpos = take(ng * 3); w.gax = take(ng * 3); w.com = take(nb * 3);
    w.cinert = take(nb * 10); w.crb = take(nb * 10); w.cdof = take(nv * 6); w.cdofdot = take(nv * 6);
    w.cvel = take(nb * 6); w.cacc = take(nb * 6); w.cfrc = take(nb * 6);
    w.M = take((size_t)nv * nv); w.H = take((size_t)nv * nv);
    w.fs = take(nv); w.as = take(nv); w.fc = take(nv); w.qacc = take(nv); w.Ma = take(nv); w.grad = take(nv); w.search = take(nv); w.Mv = take(nv);
    w.cdist = take(COOP_MAXCON); w.cpos = take(COOP_MAXCON * 3); w.cframe = take(COOP_MAXCON * 9);
    w.J = take((size_t)COOP_MAXEFC * nv); w.D = take(COOP_MAXEFC); w.aref = take(COOP_MAXEFC); w.jar = take(COOP_MAXEFC); w.jv = take(COOP_MAXEFC);
    w.cpair = reinterpret_cast<int*>(p);
}

DEV double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// in-place dense Cholesky of the n x n row-major matrix A (lower triangle), cooperative over the warp
DEV void coop_chol(double* A, int n, int lane) {
    for (int j = 0; j < n; j++) {
        __syncwarp();
        double d = A[j * n + j];
        if (d < ILQG_MINVAL) d = ILQG_MINVAL;
        double rinv = rsqrt(d);
        __syncwarp();
        for (int i = j + lane; i < n; i += 32) A[i * n + j] = (i == j) ? d * rinv : A[i * n + j] * rinv;
        __syncwarp();
        // trailing update: rows i > j, columns j < k <= i
        for (int i = j + 1 + lane; i < n; i += 32) {
            double lij = A[i * n + j];
            for (int k = j + 1; k <= i; k++) A[i * n + k] -= lij * A[k * n + j];
        }
    }
    __syncwarp();
}
// x <- (L L^T)^-1 x, column-oriented so that no reductions are needed
DEV void coop_chol_solve(const double* L, double* x, int n, int lane) {
    for (int j = 0; j < n; j++) {
        __syncwarp();
        double xj = x[j] / L[j * n + j];
        __syncwarp();
        if (lane == 0) x[j] = xj;
        for (int i = j + 1 + lane; i < n; i += 32) x[i] -= L[i * n + j] * xj;
    }
    for (int j = n - 1; j >= 0; j--) {
        __syncwarp();
        double xj = x[j] / L[j * n + j];
        __syncwarp();
        if (lane == 0) x[j] = xj;
        for (int i = lane; i < j; i += 32) x[i] -= L[j * n + i] * xj;
    }
    __syncwarp();
}

DEV void g_inert_vec(double* r, const double* i, const double* s) {  // same 10-number spatial inertia as dyn.cuh's Inert
    r[0] = i[0] * s[0] + i[3] * s[1] + i[4] * s[2] + (i[7] * s[5] - i[8] * s[4]);
    r[1] = i[3] * s[0] + i[1] * s[1] + i[5] * s[2] + (i[8] * s[3] - i[6] * s[5]);
    r[2] = i[4] * s[0] + i[5] * s[1] + i[2] * s[2] + (i[6] * s[4] - i[7] * s[3]);
    r[3] = i[9] * s[3] + (s[1] * i[8] - s[2] * i[7]);
    r[4] = i[9] * s[4] + (s[2] * i[6] - s[0] * i[8]);
    r[5] = i[9] * s[5] + (s[0] * i[7] - s[1] * i[6]);
}
DEV V3 gl3(const double* p) { return {p[0], p[1], p[2]}; }
DEV void gs3(double* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

// ------------------------------------------------------------------ the pipeline up to the constraint problem
// Inputs w.q, w.v, w.u (already perturbed).  Returns false when the contact / row capacity is exceeded.
DEV bool coop_build(const GModel* __restrict__ g, CoopMem& w, int lane) {
    const ilqg_model& m = g->m;
    const int nv = m.nv, nb = m.nbody, nj = m.njnt, ng = m.ngeom;
    // ---- kinematics, level by level (a level's bodies are independent)
    if (lane == 0) {
        w.xpos[0] = w.xpos[1] = w.xpos[2] = 0;
        w.xquat[0] = 1; w.xquat[1] = w.xquat[2] = w.xquat[3] = 0;
        for (int k = 0; k < 9; k++) w.xmat[k] = (k % 4 == 0) ? 1.0 : 0.0;
        w.xipos[0] = w.xipos[1] = w.xipos[2] = 0;
    }
    __syncwarp();
    for (int L = 0; L < g->nlevel; L++) {
        for (int idx = g->level_start[L] + lane; idx < g->level_start[L + 1]; idx += 32) {
            const int b = g->level_body[idx], p = m.body_parentid[b];
            M3 Rp = {gl3(w.xmat + 9 * p), gl3(w.xmat + 9 * p + 3), gl3(w.xmat + 9 * p + 6)};
            V3 pos = gl3(w.xpos + 3 * p) + mulv(Rp, gl3(m.body_pos[b]));
            Q4 quat = qmul({w.xquat[4 * p], w.xquat[4 * p + 1], w.xquat[4 * p + 2], w.xquat[4 * p + 3]},
                           {m.body_quat[b][0], m.body_quat[b][1], m.body_quat[b][2], m.body_quat[b][3]});
            for (int jj = 0; jj < m.body_jntnum[b]; jj++) {
                const int j = m.body_jntadr[b] + jj, qa = m.jnt_qposadr[j], ty = m.jnt_type[j];
                if (ty == ILQG_JNT_FREE) {
                    pos = {w.q[qa], w.q[qa + 1], w.q[qa + 2]};
                    quat = qnormalized({w.q[qa + 3], w.q[qa + 4], w.q[qa + 5], w.q[qa + 6]});
                    gs3(w.anchor + 3 * j, pos);
                    gs3(w.axis + 3 * j, {0, 0, 1});
                    continue;
                }
                M3 R = q2m(quat);
                V3 an = pos + mulv(R, gl3(m.jnt_pos[j]));
                V3 ax = mulv(R, gl3(m.jnt_axis[j]));
                gs3(w.anchor + 3 * j, an);
                gs3(w.axis + 3 * j, ax);
                double qq = w.q[qa] - m.qpos0[qa];
                if (ty == ILQG_JNT_SLIDE) pos = pos + qq * ax;
                else {
                    double s, c;
                    sincos(0.5 * qq, &s, &c);
                    quat = qmul(quat, {c, m.jnt_axis[j][0] * s, m.jnt_axis[j][1] * s, m.jnt_axis[j][2] * s});
                    pos = an - mulv(q2m(quat), gl3(m.jnt_pos[j]));
                }
            }
            quat = qnormalized(quat);
            M3 R = q2m(quat);
            gs3(w.xpos + 3 * b, pos);
            w.xquat[4 * b] = quat.w; w.xquat[4 * b + 1] = quat.x; w.xquat[4 * b + 2] = quat.y; w.xquat[4 * b + 3] = quat.z;
            gs3(w.xmat + 9 * b, R.r0); gs3(w.xmat + 9 * b + 3, R.r1); gs3(w.xmat + 9 * b + 6, R.r2);
            gs3(w.xipos + 3 * b, pos + mulv(R, gl3(m.body_ipos[b])));
        }
        __syncwarp();
    }
    // ---- geoms; tree centres of mass
    for (int k = lane; k < ng; k += 32) {
        const int b = m.geom_bodyid[k];
        M3 R = {gl3(w.xmat + 9 * b), gl3(w.xmat + 9 * b + 3), gl3(w.xmat + 9 * b + 6)};
        gs3(w.gpos + 3 * k, gl3(w.xpos + 3 * b) + mulv(R, gl3(m.geom_pos[k])));
        gs3(w.gax + 3 * k, mulv(R, gl3(g->geom_axis[k])));
    }
    for (int r = 1; r < nb; r++) {
        if (m.body_rootid[r] != r) continue;   // uniform branch
        if (lane < 3) {
            double s = 0, ms = 0;
            for (int b = r; b < nb; b++)
                if (m.body_rootid[b] == r) { s += m.body_mass[b] * w.xipos[3 * b + lane]; ms += m.body_mass[b]; }
            w.com[3 * r + lane] = s / ms;
        }
    }
    __syncwarp();
    // ---- spatial inertias about the tree com, motion axes
    for (int b = 1 + lane; b < nb; b += 32) {
        const double* R = w.xmat + 9 * b;
        const double* in = m.body_inertia[b];
        double Ib[9] = {in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2]}, T[9], Iw[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) T[3 * r + c] = R[3 * r] * Ib[c] + R[3 * r + 1] * Ib[3 + c] + R[3 * r + 2] * Ib[6 + c];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) Iw[3 * r + c] = T[3 * r] * R[3 * c] + T[3 * r + 1] * R[3 * c + 1] + T[3 * r + 2] * R[3 * c + 2];
        const double* cm = w.com + 3 * m.body_rootid[b];
        double d0 = w.xipos[3 * b] - cm[0], d1 = w.xipos[3 * b + 1] - cm[1], d2 = w.xipos[3 * b + 2] - cm[2], ms = m.body_mass[b];
        double dd = d0 * d0 + d1 * d1 + d2 * d2;
        double* ci = w.cinert + 10 * b;
        ci[0] = Iw[0] + ms * (dd - d0 * d0); ci[1] = Iw[4] + ms * (dd - d1 * d1); ci[2] = Iw[8] + ms * (dd - d2 * d2);
        ci[3] = Iw[1] - ms * d0 * d1; ci[4] = Iw[2] - ms * d0 * d2; ci[5] = Iw[5] - ms * d1 * d2;
        ci[6] = ms * d0; ci[7] = ms * d1; ci[8] = ms * d2; ci[9] = ms;
    }
    for (int j = lane; j < nj; j += 32) {
        const int b = m.jnt_bodyid[j], da = m.jnt_dofadr[j], ty = m.jnt_type[j];
        V3 off = gl3(w.com + 3 * m.body_rootid[b]) - gl3(w.anchor + 3 * j);
        if (ty == ILQG_JNT_FREE) {
            for (int i = 0; i < 3; i++) { for (int k = 0; k < 6; k++) w.cdof[6 * (da + i) + k] = 0; w.cdof[6 * (da + i) + 3 + i] = 1; }
            for (int i = 0; i < 3; i++) {
                V3 ax = {w.xmat[9 * b + i], w.xmat[9 * b + 3 + i], w.xmat[9 * b + 6 + i]};
                gs3(w.cdof + 6 * (da + 3 + i), ax);
                gs3(w.cdof + 6 * (da + 3 + i) + 3, cross(ax, off));
            }
        } else if (ty == ILQG_JNT_SLIDE) {
            gs3(w.cdof + 6 * da, {0, 0, 0});
            gs3(w.cdof + 6 * da + 3, gl3(w.axis + 3 * j));
        } else {
            V3 ax = gl3(w.axis + 3 * j);
            gs3(w.cdof + 6 * da, ax);
            gs3(w.cdof + 6 * da + 3, cross(ax, off));
        }
    }
    if (lane < 6) { w.cvel[lane] = 0; w.cacc[lane] = lane < 3 ? 0.0 : -m.gravity[lane - 3]; w.cfrc[lane] = 0; }
    __syncwarp();
    // ---- com velocities, cdof_dot, RNE forward sweep (level by level)
    for (int L = 0; L < g->nlevel; L++) {
        for (int idx = g->level_start[L] + lane; idx < g->level_start[L + 1]; idx += 32) {
            const int b = g->level_body[idx], p = m.body_parentid[b];
            S6 cv = {gl3(w.cvel + 6 * p), gl3(w.cvel + 6 * p + 3)}, ca = {gl3(w.cacc + 6 * p), gl3(w.cacc + 6 * p + 3)};
            for (int jj = 0; jj < m.body_jntnum[b]; jj++) {
                const int j = m.body_jntadr[b] + jj, da = m.jnt_dofadr[j];
                auto cd = [&](int i) { return S6{gl3(w.cdof + 6 * i), gl3(w.cdof + 6 * i + 3)}; };
                auto putdot = [&](int i, S6 s) { gs3(w.cdofdot + 6 * i, s.w); gs3(w.cdofdot + 6 * i + 3, s.v); };
                if (m.jnt_type[j] == ILQG_JNT_FREE) {
                    for (int i = 0; i < 3; i++) { putdot(da + i, {{0, 0, 0}, {0, 0, 0}}); cv = cv + w.v[da + i] * cd(da + i); }
                    for (int i = 3; i < 6; i++) putdot(da + i, cross_motion(cv, cd(da + i)));
                    for (int i = 3; i < 6; i++) cv = cv + w.v[da + i] * cd(da + i);
                } else {
                    putdot(da, cross_motion(cv, cd(da)));
                    cv = cv + w.v[da] * cd(da);
                }
            }
            for (int i = m.body_dofadr[b]; i < m.body_dofadr[b] + m.body_dofnum[b]; i++)
                ca = ca + w.v[i] * S6{gl3(w.cdofdot + 6 * i), gl3(w.cdofdot + 6 * i + 3)};
            gs3(w.cvel + 6 * b, cv.w); gs3(w.cvel + 6 * b + 3, cv.v);
            gs3(w.cacc + 6 * b, ca.w); gs3(w.cacc + 6 * b + 3, ca.v);
            double cvv[6] = {cv.w.x, cv.w.y, cv.w.z, cv.v.x, cv.v.y, cv.v.z}, caa[6] = {ca.w.x, ca.w.y, ca.w.z, ca.v.x, ca.v.y, ca.v.z};
            double ia[6], iv[6];
            g_inert_vec(ia, w.cinert + 10 * b, caa);
            g_inert_vec(iv, w.cinert + 10 * b, cvv);
            S6 cf = cross_force(cv, {{iv[0], iv[1], iv[2]}, {iv[3], iv[4], iv[5]}});
            w.cfrc[6 * b] = ia[0] + cf.w.x; w.cfrc[6 * b + 1] = ia[1] + cf.w.y; w.cfrc[6 * b + 2] = ia[2] + cf.w.z;
            w.cfrc[6 * b + 3] = ia[3] + cf.v.x; w.cfrc[6 * b + 4] = ia[4] + cf.v.y; w.cfrc[6 * b + 5] = ia[5] + cf.v.z;
        }
        __syncwarp();
    }
    // ---- backward accumulations: one lane per component, serial over bodies (children have larger ids than parents)
    for (int k = lane; k < 10; k += 32) w.crb[k] = 0;
    for (int e = 10 + lane; e < 10 * nb; e += 32) w.crb[e] = w.cinert[e];
    __syncwarp();
    if (lane < 6) {
        for (int b = nb - 1; b > 0; b--) { int p = m.body_parentid[b]; if (p > 0) w.cfrc[6 * p + lane] += w.cfrc[6 * b + lane]; }
    } else if (lane < 16) {
        const int k = lane - 6;
        for (int b = nb - 1; b > 0; b--) { int p = m.body_parentid[b]; if (p > 0) w.crb[10 * p + k] += w.crb[10 * b + k]; }
    }
    for (int e = lane; e < nv * nv; e += 32) w.M[e] = 0;
    __syncwarp();
    // ---- qfrc_smooth = passive - bias + actuator ; mass matrix
    for (int i = lane; i < nv; i += 32) {
        const int b = m.dof_bodyid[i], j = m.dof_jntid[i];
        double f = 0;
        for (int k = 0; k < 6; k++) f -= w.cdof[6 * i + k] * w.cfrc[6 * b + k];
        f -= m.dof_damping[i] * w.v[i];
        if (m.jnt_type[j] != ILQG_JNT_FREE && m.jnt_stiffness[j] != 0) f -= m.jnt_stiffness[j] * (w.q[m.jnt_qposadr[j]] - m.qpos_spring[m.jnt_qposadr[j]]);
        for (int a = 0; a < m.nu; a++)
            if (m.act_dofid[a] == i) {
                double c = w.u[a];
                if (m.act_ctrllimited[a]) c = clampd(c, m.act_ctrlrange[a][0], m.act_ctrlrange[a][1]);
                f += m.act_gear[a] * c;
            }
        w.fs[i] = f;
        w.as[i] = f;
        double buf[6];
        g_inert_vec(buf, w.crb + 10 * b, w.cdof + 6 * i);
        for (int a = i; a >= 0; a = m.dof_parentid[a]) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += w.cdof[6 * a + k] * buf[k];
            if (a == i) s += m.dof_armature[i];
            w.M[i * nv + a] = s;
            w.M[a * nv + i] = s;
        }
    }
    __syncwarp();
    for (int e = lane; e < nv * nv; e += 32) w.H[e] = w.M[e];
    coop_chol(w.H, nv, lane);
    coop_chol_solve(w.H, w.as, nv, lane);
    // ---- constraint rows: joint limits
    int ne = 0;
    for (int j0 = 0; j0 < nj; j0 += 32) {
        const int j = j0 + lane;
        int side = 0;
        double dist = 0;
        if (j < nj && m.jnt_limited[j] && m.jnt_type[j] != ILQG_JNT_FREE) {
            double value = w.q[m.jnt_qposadr[j]];
            double dlo = value - m.jnt_range[j][0], dhi = m.jnt_range[j][1] - value;
            if (dlo < m.jnt_margin[j]) { side = -1; dist = dlo; }
            else if (dhi < m.jnt_margin[j]) { side = 1; dist = dhi; }
        }
        unsigned bal = __ballot_sync(0xffffffffu, side != 0);
        int r = ne + __popc(bal & ((1u << lane) - 1u));
        if (side != 0 && r < COOP_MAXEFC) {
            const int da = m.jnt_dofadr[j];
            double R, kt;
            row_params(g->jnt_K[j], g->jnt_imp[j], m.jnt_solimp[j], dist, m.jnt_margin[j], m.dof_invweight0[da], R, kt);
            for (int i = 0; i < nv; i++) w.J[r * nv + i] = 0;
            w.J[r * nv + da] = -side;
            w.D[r] = 1.0 / R;
            w.aref[r] = -g->jnt_B[j] * (-side * w.v[da]) - kt;
        }
        ne += __popc(bal);
    }
    // ---- collision: lanes over candidate pairs, contacts appended through warp prefix sums
    int ncon = 0;
    for (int p0 = 0; p0 < m.npair; p0 += 32) {
        const int p = p0 + lane;
        int cnt = 0;
        double cd[2], cp[2][3], cn[2][3], ch[3] = {0, 0, 0};
        bool hint = false;
        if (p < m.npair) {
            const int g1 = m.pair_geom1[p], g2 = m.pair_geom2[p], t1 = m.geom_type[g1], t2 = m.geom_type[g2];
            const double margin = m.pair_margin[p];
            V3 x1 = gl3(w.gpos + 3 * g1), x2 = gl3(w.gpos + 3 * g2), a1 = gl3(w.gax + 3 * g1), a2 = gl3(w.gax + 3 * g2);
            auto sphere_sphere = [&](V3 c1, double r1, V3 c2, double r2) {
                V3 n = c2 - c1;
                double len = sqrt(dot(n, n)), dist = len - r1 - r2;
                if (dist > margin || cnt >= 2) return;
                n = len < ILQG_MINVAL ? V3{1, 0, 0} : (1.0 / len) * n;
                V3 pos = c1 + (r1 + 0.5 * dist) * n;
                cd[cnt] = dist; gs3(cp[cnt], pos); gs3(cn[cnt], n); cnt++;
            };
            auto plane_sphere = [&](V3 c, double r) {
                double dist = dot(c - x1, a1) - r;
                if (dist > margin) return;
                V3 pos = c - (r + 0.5 * dist) * a1;
                cd[cnt] = dist; gs3(cp[cnt], pos); gs3(cn[cnt], a1); cnt++;
            };
            if (t1 == ILQG_GEOM_PLANE && t2 == ILQG_GEOM_SPHERE) plane_sphere(x2, m.geom_size[g2][0]);
            else if (t1 == ILQG_GEOM_PLANE && t2 == ILQG_GEOM_CAPSULE) {
                double r = m.geom_size[g2][0], h = m.geom_size[g2][1];
                hint = true; gs3(ch, a2);
                plane_sphere(x2 + h * a2, r);
                plane_sphere(x2 - h * a2, r);
            } else if (t1 == ILQG_GEOM_SPHERE && t2 == ILQG_GEOM_SPHERE) sphere_sphere(x1, m.geom_size[g1][0], x2, m.geom_size[g2][0]);
            else if (t1 == ILQG_GEOM_SPHERE && t2 == ILQG_GEOM_CAPSULE) {
                double h = m.geom_size[g2][1];
                double t = clampd(dot(x1 - x2, a2), -h, h);
                sphere_sphere(x1, m.geom_size[g1][0], x2 + t * a2, m.geom_size[g2][0]);
            } else if (t1 == ILQG_GEOM_CAPSULE && t2 == ILQG_GEOM_CAPSULE) {
                double r1 = m.geom_size[g1][0], h1 = m.geom_size[g1][1], r2 = m.geom_size[g2][0], h2 = m.geom_size[g2][1];
                V3 dif = x1 - x2;
                double mb = -dot(a1, a2), uu = -dot(a1, dif), vv = dot(a2, dif), det = 1.0 - mb * mb;
                if (fabs(det) >= 1e-12) {
                    double s1 = (uu - mb * vv) / det, s2 = (vv - mb * uu) / det;
                    if (s1 > h1) { s1 = h1; s2 = vv - mb * h1; }
                    else if (s1 < -h1) { s1 = -h1; s2 = vv + mb * h1; }
                    if (s2 > h2) { s2 = h2; s1 = clampd(uu - mb * h2, -h1, h1); }
                    else if (s2 < -h2) { s2 = -h2; s1 = clampd(uu + mb * h2, -h1, h1); }
                    sphere_sphere(x1 + s1 * a1, r1, x2 + s2 * a2, r2);
                } else {
                    for (int s = -1; s <= 1 && cnt < 2; s += 2) {
                        V3 c1 = x1 + (s * h1) * a1;
                        double t = dot(c1 - x2, a2);
                        if (t < -h2 || t > h2) continue;
                        sphere_sphere(c1, r1, x2 + t * a2, r2);
                    }
                    for (int s = -1; s <= 1 && cnt < 2; s += 2) {
                        V3 c2 = x2 + (s * h2) * a2;
                        double t = dot(c2 - x1, a1);
                        if (t <= -h1 || t >= h1) continue;
                        sphere_sphere(x1 + t * a1, r1, c2, r2);
                    }
                    if (cnt == 0) {
                        double best = 1e300;
                        V3 b1 = x1, b2 = x2;
                        for (int s = -1; s <= 1; s += 2)
                            for (int t = -1; t <= 1; t += 2) {
                                V3 c1 = x1 + (s * h1) * a1, c2 = x2 + (t * h2) * a2;
                                double dd = dot(c1 - c2, c1 - c2);
                                if (dd < best) { best = dd; b1 = c1; b2 = c2; }
                            }
                        sphere_sphere(b1, r1, b2, r2);
                    }
                }
            }
        }
        // exclusive prefix of contact counts over the warp (pair order = oracle's contact order)
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        int base = ncon + incl - cnt;
        for (int c = 0; c < cnt; c++) {
            int slot = base + c;
            if (slot < COOP_MAXCON) {
                V3 ta, tb;
                make_frame(gl3(cn[c]), gl3(ch), hint, ta, tb);
                w.cdist[slot] = cd[c];
                gs3(w.cpos + 3 * slot, gl3(cp[c]));
                gs3(w.cframe + 9 * slot, gl3(cn[c])); gs3(w.cframe + 9 * slot + 3, ta); gs3(w.cframe + 9 * slot + 6, tb);
                w.cpair[slot] = p;
            }
        }
        ncon += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    bool ok = ncon <= COOP_MAXCON;
    if (!ok) ncon = COOP_MAXCON;
    // ---- contact rows: sequential over contacts, lanes over dofs
    for (int c = 0; c < ncon; c++) {
        const int p = w.cpair[c], condim = m.pair_condim[p];
        const int b1 = m.geom_bodyid[m.pair_geom1[p]], b2 = m.geom_bodyid[m.pair_geom2[p]];
        const int nrows = condim == 3 ? 4 : 1;
        if (ne + nrows > COOP_MAXEFC) { ok = false; break; }
        V3 pos = gl3(w.cpos + 3 * c), n = gl3(w.cframe + 9 * c), ta = gl3(w.cframe + 9 * c + 3), tb = gl3(w.cframe + 9 * c + 6);
        double jn = 0, ja = 0, jb = 0;
        if (lane < nv) {
            const bool m1 = (g->body_dofmask[b1] >> lane) & 1u, m2 = (g->body_dofmask[b2] >> lane) & 1u;
            if (m1 != m2) {
                const int bd = m2 ? b2 : b1;
                V3 jp = gl3(w.cdof + 6 * lane + 3) + cross(gl3(w.cdof + 6 * lane), pos - gl3(w.com + 3 * m.body_rootid[bd]));
                if (!m2) jp = -1.0 * jp;
                jn = dot(n, jp); ja = dot(ta, jp); jb = dot(tb, jp);
            }
        }
        const double qv = lane < nv ? w.v[lane] : 0.0;
        const double vn = warp_sum(jn * qv), va = warp_sum(ja * qv), vb = warp_sum(jb * qv);
        double tran = m.body_invweight0[b1][0] + m.body_invweight0[b2][0];
        if (tran < ILQG_MINVAL) tran = ILQG_MINVAL;
        const double margin = m.pair_margin[p], dist = w.cdist[c], B = g->pair_B[p];
        if (condim == 1) {
            double R, kt;
            row_params(g->pair_K[p], g->pair_imp[p], m.pair_solimp[p], dist, margin, tran, R, kt);
            if (lane < nv) w.J[ne * nv + lane] = jn;
            if (lane == 0) { w.D[ne] = 1.0 / R; w.aref[ne] = -B * vn - kt; }
            ne += 1;
        } else {
            const double mu = m.pair_friction[p];
            double R0, kt;
            row_params(g->pair_K[p], g->pair_imp[p], m.pair_solimp[p], dist, margin, tran * (1 + mu * mu), R0, kt);
            double Rpy = 2 * mu * mu * R0;
            if (Rpy < ILQG_MINVAL) Rpy = ILQG_MINVAL;
            for (int k = 0; k < 4; k++) {
                const double sg = (k & 1) ? -mu : mu;
                if (lane < nv) w.J[(ne + k) * nv + lane] = jn + sg * (k < 2 ? ja : jb);
                if (lane == 0) { w.D[ne + k] = 1.0 / Rpy; w.aref[ne + k] = -B * (vn + sg * (k < 2 ? va : vb)) - kt; }
            }
            ne += 4;
        }
    }
    __syncwarp();
    w.ncon = ncon;
    w.nefc = ne;
    return ok;
}

// ------------------------------------------------------------------ constraint solve (same algorithm as dyn.cuh::solve)
DEV double coop_cost(const CoopMem& w, const double* a, int nv, int lane) {
    double c = 0;
    for (int r = lane; r < w.nefc; r += 32) {
        double jar = -w.aref[r];
        for (int i = 0; i < nv; i++) jar += w.J[r * nv + i] * a[i];
        if (jar < 0) c += 0.5 * w.D[r] * jar * jar;
    }
    for (int i = lane; i < nv; i += 32) {
        double Ma = 0;
        for (int k = 0; k < nv; k++) Ma += w.M[i * nv + k] * a[k];
        c += 0.5 * (Ma - w.fs[i]) * (a[i] - w.as[i]);
    }
    return warp_sum(c);
}

struct Mask128 { unsigned w[4]; };
DEV bool operator==(const Mask128& a, const Mask128& b) { return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2] && a.w[3] == b.w[3]; }

// result in w.qacc (and w.warm, the next warm start)
DEV void coop_solve(const GModel* __restrict__ g, CoopMem& w, int maxiter, double tol, int lane, bool need_forces = false) {
    const ilqg_model& m = g->m;
    const int nv = m.nv, ne = w.nefc;
    static_assert(COOP_MAXEFC <= 128, "mask width");
    if (ne == 0) {
        for (int i = lane; i < nv; i += 32) { w.qacc[i] = w.as[i]; w.warm[i] = w.as[i]; w.fc[i] = 0; }
        __syncwarp();
        return;
    }
    {
        double cw = coop_cost(w, w.warm, nv, lane), cs = coop_cost(w, w.as, nv, lane);
        for (int i = lane; i < nv; i += 32) w.qacc[i] = cw < cs ? w.warm[i] : w.as[i];
    }
    __syncwarp();
    const double scale = 1.0 / (m.meaninertia * (nv > 1 ? nv : 1));
    for (int i = lane; i < nv; i += 32) {
        double s = 0;
        for (int k = 0; k < nv; k++) s += w.M[i * nv + k] * w.qacc[k];
        w.Ma[i] = s;
    }
    for (int r = lane; r < ne; r += 32) {
        double s = -w.aref[r];
        for (int i = 0; i < nv; i++) s += w.J[r * nv + i] * w.qacc[i];
        w.jar[r] = s;
    }
    __syncwarp();
    double cost = 0, old = 0;
    int iter = 0;
    for (;;) {
        // ---- active set, cost, forces, gradient, Hessian, Newton direction
        Mask128 act = {{0, 0, 0, 0}};
        double c = 0;
        for (int r0 = 0; r0 < ne; r0 += 32) {
            const int r = r0 + lane;
            bool a = r < ne && w.jar[r] < 0;
            act.w[r0 >> 5] = __ballot_sync(0xffffffffu, a);
            if (a) c += 0.5 * w.D[r] * w.jar[r] * w.jar[r];
        }
        for (int i = lane; i < nv; i += 32) {
            double f = 0;
            for (int r = 0; r < ne; r++)
                if ((act.w[r >> 5] >> (r & 31)) & 1u) f += w.J[r * nv + i] * (-w.D[r] * w.jar[r]);
            w.fc[i] = f;
            c += 0.5 * (w.Ma[i] - w.fs[i]) * (w.qacc[i] - w.as[i]);
            w.grad[i] = w.Ma[i] - w.fs[i] - f;
            w.search[i] = w.grad[i];
        }
        cost = warp_sum(c);
        for (int e = lane; e < nv * (nv + 1) / 2; e += 32) {
            // e -> (i, j), j <= i
            int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while ((i + 1) * (i + 2) / 2 <= e) i++;
            while (i * (i + 1) / 2 > e) i--;
            const int j = e - i * (i + 1) / 2;
            double h = w.M[i * nv + j];
            for (int r = 0; r < ne; r++)
                if ((act.w[r >> 5] >> (r & 31)) & 1u) h += w.D[r] * w.J[r * nv + i] * w.J[r * nv + j];
            w.H[i * nv + j] = h;
        }
        __syncwarp();
        coop_chol(w.H, nv, lane);
        coop_chol_solve(w.H, w.search, nv, lane);
        for (int i = lane; i < nv; i += 32) w.search[i] = -w.search[i];
        __syncwarp();
        if (iter > 0) {
            double gn = 0;
            for (int i = lane; i < nv; i += 32) gn += w.grad[i] * w.grad[i];
            gn = warp_sum(gn);
            if (scale * (old - cost) < tol || scale * sqrt(gn) < tol) break;
        }
        if (iter >= maxiter) break;
        // ---- exact line search
        double g1 = 0, g2 = 0;
        for (int i = lane; i < nv; i += 32) {
            double s = 0;
            for (int k = 0; k < nv; k++) s += w.M[i * nv + k] * w.search[k];
            w.Mv[i] = s;
            g1 += w.search[i] * (w.Ma[i] - w.fs[i]);
            g2 += w.search[i] * s;
        }
        double d1 = 0, d2 = 0;
        for (int r = lane; r < ne; r += 32) {
            double s = 0;
            for (int i = 0; i < nv; i++) s += w.J[r * nv + i] * w.search[i];
            w.jv[r] = s;
            if (w.jar[r] < 0) { double t = w.D[r] * s; d1 += t * w.jar[r]; d2 += t * s; }
        }
        g1 = warp_sum(g1); g2 = warp_sum(g2);
        d1 = warp_sum(d1) + g1; d2 = warp_sum(d2) + g2;
        __syncwarp();
        if (d1 >= 0 || d2 < ILQG_MINVAL) break;
        double alpha = 0, lo = 0, hi = CUDART_INF;
        Mask128 cur = act;
        for (int it = 0; it < m.ls_iterations; it++) {
            if (d1 < 0) lo = alpha; else hi = alpha;
            double an = alpha - d1 / d2;
            if (!(an > lo && an < hi)) an = isinf(hi) ? 2 * alpha + 1 : 0.5 * (lo + hi);
            double e1 = 0, e2 = 0;
            Mask128 mk = {{0, 0, 0, 0}};
            for (int r0 = 0; r0 < ne; r0 += 32) {
                const int r = r0 + lane;
                bool a = false;
                if (r < ne) {
                    double jv = w.jv[r], x = w.jar[r] + an * jv;
                    if (x < 0) { double t = w.D[r] * jv; e1 += t * x; e2 += t * jv; a = true; }
                }
                mk.w[r0 >> 5] = __ballot_sync(0xffffffffu, a);
            }
            e1 = warp_sum(e1) + g1 + g2 * an;
            e2 = warp_sum(e2) + g2;
            bool same = mk == cur;
            alpha = an; d1 = e1; d2 = e2; cur = mk;
            if (same || d1 == 0 || d2 < ILQG_MINVAL) break;
        }
        if (alpha == 0) break;
        for (int i = lane; i < nv; i += 32) { w.qacc[i] += alpha * w.search[i]; w.Ma[i] += alpha * w.Mv[i]; }
        for (int r = lane; r < ne; r += 32) w.jar[r] += alpha * w.jv[r];
        __syncwarp();
        old = cost;
        iter++;
        if (cur == act) break;  // exact optimum: full Newton step inside the piece the Hessian was built for
    }
    if (need_forces) {  // qfrc_constraint at the final point (mj_Euler needs it)
        for (int i = lane; i < nv; i += 32) {
            double f = 0;
            for (int r = 0; r < ne; r++)
                if (w.jar[r] < 0) f += w.J[r * nv + i] * (-w.D[r] * w.jar[r]);
            w.fc[i] = f;
        }
    }
    for (int i = lane; i < nv; i += 32) w.warm[i] = w.qacc[i];
    __syncwarp();
}

// ------------------------------------------------------------------ kernels
// centre: one warp per knot -> qacc_center (the warm start of every perturbed solve)
__global__ void __launch_bounds__(128) coop_center_kernel(const GModel* __restrict__ g, int nknots, const double* __restrict__ qpos,
                                                          const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                          const double* __restrict__ warmstart, int niter, int nwarmup, size_t warp_bytes,
                                                          double* __restrict__ qacc_center, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char coop_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * (blockDim.x >> 5) + wib;
    if (k >= nknots) return;
    const ilqg_model& m = g->m;
    CoopMem w;
    coop_carve(w, reinterpret_cast<double*>(coop_smem + wib * warp_bytes), m);
    for (int i = lane; i < m.nq; i += 32) w.q[i] = qpos[(size_t)k * m.nq + i];
    for (int i = lane; i < m.nv; i += 32) { w.v[i] = qvel[(size_t)k * m.nv + i]; w.warm[i] = warmstart ? warmstart[(size_t)k * m.nv + i] : 0.0; }
    for (int i = lane; i < m.nu; i += 32) w.u[i] = ctrl[(size_t)k * m.nu + i];
    __syncwarp();
    bool ok = coop_build(g, w, lane);
    for (int rep = 0; rep < nwarmup; rep++) coop_solve(g, w, niter, 0.0, lane);
    bool fin = true;
    for (int i = lane; i < m.nv; i += 32) { qacc_center[(size_t)k * m.nv + i] = w.qacc[i]; fin = fin && isfinite(w.qacc[i]); }
    fin = __all_sync(0xffffffffu, fin);
    if (status && lane == 0) status[k] = !ok ? ILQG_ERR_CAPACITY : (fin ? 0 : ILQG_ERR_NONFINITE);
}

// perturbed: one warp per (knot, column); the warp evaluates +eps then -eps and writes the column of the deriv block
__global__ void __launch_bounds__(128) coop_perturb_kernel(const GModel* __restrict__ g, int nknots, const double* __restrict__ qpos,
                                                           const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                           const double* __restrict__ qacc_center, const ilqg_cost* __restrict__ cost, double eps,
                                                           int niter, size_t warp_bytes, double* __restrict__ deriv, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char coop_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const ilqg_model& m = g->m;
    const int nq = m.nq, nv = m.nv, nu = m.nu, ncol = 2 * nv + nu, nd = nv * ncol + ncol;
    const long item = (long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (item >= (long)nknots * ncol) return;
    const int k = (int)(item / ncol), col = (int)(item - (long)k * ncol);
    CoopMem w;
    coop_carve(w, reinterpret_cast<double*>(coop_smem + wib * warp_bytes), m);
    double plus[1];  // this lane's component of qacc(+eps) (nv <= 32)
    plus[0] = 0;
    bool ok = true;
    double dcost = 0;
    for (int sgn = 1; sgn >= -1; sgn -= 2) {
        for (int i = lane; i < nq; i += 32) w.q[i] = qpos[(size_t)k * nq + i];
        for (int i = lane; i < nv; i += 32) { w.v[i] = qvel[(size_t)k * nv + i]; w.warm[i] = qacc_center[(size_t)k * nv + i]; }
        for (int i = lane; i < nu; i += 32) w.u[i] = ctrl[(size_t)k * nu + i];
        __syncwarp();
        double c0 = 0;
        if (cost && sgn > 0 && lane == 0) {
            for (int i = 0; i < nq; i++) { c0 = __dadd_rn(c0, __dmul_rn(__dmul_rn(cost->q2[i], w.q[i]), w.q[i])); c0 = __dadd_rn(c0, __dmul_rn(cost->q1[i], w.q[i])); }
            for (int i = 0; i < nv; i++) { c0 = __dadd_rn(c0, __dmul_rn(__dmul_rn(cost->v2[i], w.v[i]), w.v[i])); c0 = __dadd_rn(c0, __dmul_rn(cost->v1[i], w.v[i])); }
            for (int i = 0; i < nu; i++) { c0 = __dadd_rn(c0, __dmul_rn(__dmul_rn(cost->u2[i], w.u[i]), w.u[i])); c0 = __dadd_rn(c0, __dmul_rn(cost->u1[i], w.u[i])); }
        }
        if (lane == 0) {
            const double se = sgn * eps;
            if (col < nu) w.u[col] += se;
            else if (col < nu + nv) w.v[col - nu] += se;
            else {
                const int i = col - nu - nv, j = m.dof_jntid[i];
                if (m.jnt_type[j] == ILQG_JNT_FREE && i >= m.jnt_dofadr[j] + 3) {
                    const int a = i - m.jnt_dofadr[j] - 3;
                    quat_integrate(&w.q[m.jnt_qposadr[j] + 3], V3{a == 0 ? se : 0.0, a == 1 ? se : 0.0, a == 2 ? se : 0.0}, 1.0);
                } else
                    w.q[m.jnt_qposadr[j] + i - m.jnt_dofadr[j]] += se;
            }
            if (cost && sgn > 0) {
                double c1 = 0;
                for (int i = 0; i < nq; i++) { c1 = __dadd_rn(c1, __dmul_rn(__dmul_rn(cost->q2[i], w.q[i]), w.q[i])); c1 = __dadd_rn(c1, __dmul_rn(cost->q1[i], w.q[i])); }
                for (int i = 0; i < nv; i++) { c1 = __dadd_rn(c1, __dmul_rn(__dmul_rn(cost->v2[i], w.v[i]), w.v[i])); c1 = __dadd_rn(c1, __dmul_rn(cost->v1[i], w.v[i])); }
                for (int i = 0; i < nu; i++) { c1 = __dadd_rn(c1, __dmul_rn(__dmul_rn(cost->u2[i], w.u[i]), w.u[i])); c1 = __dadd_rn(c1, __dmul_rn(cost->u1[i], w.u[i])); }
                dcost = __ddiv_rn(__dsub_rn(c1, c0), eps);
            }
        }
        __syncwarp();
        ok = coop_build(g, w, lane) && ok;
        coop_solve(g, w, niter, 0.0, lane);
        if (sgn > 0) plus[0] = lane < nv ? w.qacc[lane] : 0.0;
        __syncwarp();
    }
    // column `col`: d qacc_j / d input, j = lane
    bool fin = true;
    if (lane < nv) {
        double d = (plus[0] - w.qacc[lane]) / (2 * eps);
        fin = isfinite(d);
        size_t off;
        if (col < nu) off = 2 * (size_t)nv * nv + col + (size_t)lane * nu;
        else if (col < nu + nv) off = (size_t)nv * nv + (col - nu) + (size_t)lane * nv;
        else off = (col - nu - nv) + (size_t)lane * nv;
        deriv[(size_t)k * nd + off] = d;
    }
    if (cost && lane == 0) {
        size_t off = (size_t)nv * ncol;
        if (col < nu) off += 2 * nv + col;
        else if (col < nu + nv) off += nv + (col - nu);
        else off += col - nu - nv;
        deriv[(size_t)k * nd + off] = dcost;
    }
    fin = __all_sync(0xffffffffu, fin);
    if (status && lane == 0) {
        if (!ok) atomicExch(&status[k], ILQG_ERR_CAPACITY);
        else if (!fin) atomicCAS(&status[k], 0, ILQG_ERR_NONFINITE);
    }
}

// mj_forward for n states: one warp per state
__global__ void __launch_bounds__(128) coop_forward_kernel(const GModel* __restrict__ g, int n, const double* __restrict__ qpos,
                                                           const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                           double* __restrict__ warmstart, double* __restrict__ qacc_out, size_t warp_bytes) {
    extern __shared__ __align__(16) unsigned char coop_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * (blockDim.x >> 5) + wib;
    if (k >= n) return;
    const ilqg_model& m = g->m;
    CoopMem w;
    coop_carve(w, reinterpret_cast<double*>(coop_smem + wib * warp_bytes), m);
    for (int i = lane; i < m.nq; i += 32) w.q[i] = qpos[(size_t)k * m.nq + i];
    for (int i = lane; i < m.nv; i += 32) { w.v[i] = qvel[(size_t)k * m.nv + i]; w.warm[i] = warmstart ? warmstart[(size_t)k * m.nv + i] : 0.0; }
    for (int i = lane; i < m.nu; i += 32) w.u[i] = ctrl[(size_t)k * m.nu + i];
    __syncwarp();
    coop_build(g, w, lane);
    coop_solve(g, w, m.iterations, m.tolerance, lane);
    for (int i = lane; i < m.nv; i += 32) {
        qacc_out[(size_t)k * m.nv + i] = w.qacc[i];
        if (warmstart) warmstart[(size_t)k * m.nv + i] = w.warm[i];
    }
}

// nsteps x mj_step for n states (Euler with implicit joint damping; RK4 models use the thread-per-rollout path)
__global__ void __launch_bounds__(128) coop_step_kernel(const GModel* __restrict__ g, int n, int nsteps, double* __restrict__ qpos,
                                                        double* __restrict__ qvel, const double* __restrict__ ctrl, double* __restrict__ warmstart,
                                                        double* __restrict__ qacc_out, size_t warp_bytes) {
    extern __shared__ __align__(16) unsigned char coop_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * (blockDim.x >> 5) + wib;
    if (k >= n) return;
    const ilqg_model& m = g->m;
    const int nv = m.nv;
    const double h = m.timestep;
    CoopMem w;
    coop_carve(w, reinterpret_cast<double*>(coop_smem + wib * warp_bytes), m);
    for (int i = lane; i < m.nq; i += 32) w.q[i] = qpos[(size_t)k * m.nq + i];
    for (int i = lane; i < nv; i += 32) { w.v[i] = qvel[(size_t)k * nv + i]; w.warm[i] = warmstart ? warmstart[(size_t)k * nv + i] : 0.0; }
    for (int i = lane; i < m.nu; i += 32) w.u[i] = ctrl[(size_t)k * m.nu + i];
    __syncwarp();
    for (int s = 0; s < nsteps; s++) {
        coop_build(g, w, lane);
        coop_solve(g, w, m.iterations, m.tolerance, lane, true);
        // mj_Euler: (M + h diag(b)) a = qfrc_smooth + qfrc_constraint when any dof is damped
        if (g->any_damping) {
            for (int e = lane; e < nv * nv; e += 32) w.H[e] = w.M[e] + ((e / nv) == (e % nv) ? h * m.dof_damping[e / nv] : 0.0);
            for (int i = lane; i < nv; i += 32) w.search[i] = w.fs[i] + w.fc[i];
            __syncwarp();
            coop_chol(w.H, nv, lane);
            coop_chol_solve(w.H, w.search, nv, lane);
        } else {
            for (int i = lane; i < nv; i += 32) w.search[i] = w.qacc[i];
            __syncwarp();
        }
        for (int i = lane; i < nv; i += 32) w.v[i] += h * w.search[i];
        __syncwarp();
        for (int j = lane; j < m.njnt; j += 32) {
            const int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
            if (m.jnt_type[j] == ILQG_JNT_FREE) {
                for (int c = 0; c < 3; c++) w.q[qa + c] += h * w.v[da + c];
                quat_integrate(&w.q[qa + 3], V3{w.v[da + 3], w.v[da + 4], w.v[da + 5]}, h);
            } else
                w.q[qa] += h * w.v[da];
        }
        __syncwarp();
    }
    for (int i = lane; i < m.nq; i += 32) qpos[(size_t)k * m.nq + i] = w.q[i];
    for (int i = lane; i < nv; i += 32) {
        qvel[(size_t)k * nv + i] = w.v[i];
        if (warmstart) warmstart[(size_t)k * nv + i] = w.warm[i];
        if (qacc_out) qacc_out[(size_t)k * nv + i] = w.qacc[i];
    }
}

}  // namespace ilqg
