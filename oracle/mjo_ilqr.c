/* mjo_ilqr.c — CPU ORACLE (test infrastructure): Eigen-free restatement of the reference's
 * Differentiator (/root/reference/inc/differentiator.h:52-93), ILQR (/root/reference/inc/ilqr.h:69-186)
 * and InvertedPendulum MPC driver (/root/reference/src/inverted_pendulum/inverted_pendulum.cpp:6-30),
 * with every observable quirk of SURVEY.md Appendix B kept (Q1 column-major views of the row-major deriv
 * blocks, Q2 explicit-Euler A/B, Q3 mu added to V and never removed, Q4 v uses the updated V, Q5 K/k start
 * at zero, Q10 rank-1 Hessians, Q11 affine term c, Q12 nominal overwritten in place).
 *
 * A10 (absent from the reference, named by the north star; this file is its specification):
 *   forward pass with step size alpha:  u = K (x - x*) + alpha k + u*        (alpha = 1 is ilqr.h:126)
 *   trajectory cost                  :  J = sum_{n=0..N} stepCost(knot n) evaluated on the knots the pass stores
 *   backtracking ladder              :  alphas[0..nalpha) tried in order; the first with J(alpha) < J_prev is
 *                                       accepted and overwrites the nominal; if none is, the nominal is kept
 *   mu                               :  constant (ilqr.h:65)
 * With nalpha = 1, alphas = {1} and accept_always != 0 this is exactly ILQR::iterate().
 */
#include "mjo.h"
#include "../include/ilqg_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

extern int mjo_ilqr_corrected_layout;

typedef struct mjo_ilqr {
    const ilqg_model* m;
    int N, nq, nv, nu, nx, nd;
    double mu;
    mjo_cost_fn cost;
    void* user;
    mjo_data* d;                 /* ILQR::d — the rolling state */
    double *qpos, *qvel, *ctrl, *warm, *qacc; /* dArray[n], n = 0..N (n = N initial, n = 0 final) */
    double *K, *k;               /* K[n]: nu x nx column-major, k[n]: nu */
    double *V, *v;               /* nx x nx column-major, 1 x nx */
    double *deriv;               /* last FD buffer of every knot */
    double *A, *B;               /* Differentiator::A (nx x nx), B (nx x nu), column-major */
    mjo_data* scratch;
    double J;                    /* cost of the current nominal (A10) */
} mjo_ilqr;

#define CM(M, r, c, rows) ((M)[(r) + (size_t)(c) * (rows)])

/* x_a (-) x_b in the TANGENT space of the configuration manifold: out[0..nv) = position part, out[nv..2nv) = velocity part.
   For slide / hinge dofs this is the plain difference the reference takes on raw qpos memory (ilqr.h:126,161-163); for a free
   joint's orientation it is MuJoCo's mju_subQuat: the body-frame rotation vector w with q_b * quat(w) = q_a.  The reference's
   own state vector is "2 nv doubles starting at qpos" and is undefined when nq != nv (quirk Q9): this is the opt-in extension
   that makes Differentiator / ILQR meaningful for the humanoid, in the same coordinates the FD blocks of
   calcMJDerivatives already use (mjderivative.cpp:152-169). */
void mjo_state_diff(const ilqg_model* m, const double* qa, const double* va, const double* qb, const double* vb, double* out) {
    int nv = m->nv;
    for (int j = 0; j < m->njnt; j++) {
        int qadr = m->jnt_qposadr[j], dadr = m->jnt_dofadr[j];
        if (m->jnt_type[j] == ILQG_JNT_FREE || m->jnt_type[j] == ILQG_JNT_BALL) {
            const int fr = m->jnt_type[j] == ILQG_JNT_FREE ? 3 : 0;   /* a free joint carries a position in front of its quaternion */
            for (int k = 0; k < fr; k++) out[dadr + k] = qa[qadr + k] - qb[qadr + k];
            const double *a = qa + qadr + fr, *b = qb + qadr + fr;
            double na = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]), nb = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2] + b[3] * b[3]);
            double A[4] = {a[0] / na, a[1] / na, a[2] / na, a[3] / na}, B[4] = {b[0] / nb, -b[1] / nb, -b[2] / nb, -b[3] / nb}; /* conj(q_b) */
            double d[4] = {B[0] * A[0] - B[1] * A[1] - B[2] * A[2] - B[3] * A[3], B[0] * A[1] + B[1] * A[0] + B[2] * A[3] - B[3] * A[2],
                           B[0] * A[2] - B[1] * A[3] + B[2] * A[0] + B[3] * A[1], B[0] * A[3] + B[1] * A[2] - B[2] * A[1] + B[3] * A[0]};
            double sn = sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
            double ang = 2 * atan2(sn, d[0]);
            if (ang > 3.14159265358979323846) ang -= 2 * 3.14159265358979323846;
            double sc = sn < 1e-15 ? 0.0 : ang / sn;
            for (int k = 0; k < 3; k++) out[dadr + fr + k] = d[1 + k] * sc;
        } else
            out[dadr] = qa[qadr] - qb[qadr];
    }
    for (int i = 0; i < nv; i++) out[nv + i] = va[i] - vb[i];
}

static void knot_store(mjo_ilqr* il, int n, const mjo_data* d) {
    memcpy(il->qpos + (size_t)n * il->nq, d->qpos, sizeof(double) * il->nq);
    memcpy(il->qvel + (size_t)n * il->nv, d->qvel, sizeof(double) * il->nv);
    memcpy(il->ctrl + (size_t)n * il->nu, d->ctrl, sizeof(double) * il->nu);
    memcpy(il->warm + (size_t)n * il->nv, d->qacc_warmstart, sizeof(double) * il->nv);
    memcpy(il->qacc + (size_t)n * il->nv, d->qacc, sizeof(double) * il->nv);
}
__attribute__((unused)) static void knot_load(const mjo_ilqr* il, int n, mjo_data* d) {
    memcpy(d->qpos, il->qpos + (size_t)n * il->nq, sizeof(double) * il->nq);
    memcpy(d->qvel, il->qvel + (size_t)n * il->nv, sizeof(double) * il->nv);
    memcpy(d->ctrl, il->ctrl + (size_t)n * il->nu, sizeof(double) * il->nu);
    memcpy(d->qacc_warmstart, il->warm + (size_t)n * il->nv, sizeof(double) * il->nv);
    memcpy(d->qacc, il->qacc + (size_t)n * il->nv, sizeof(double) * il->nv);
}

mjo_ilqr* mjo_ilqr_create(const ilqg_model* m, int N, mjo_cost_fn cost, void* user) {
    /* the reference's state vector assumes nq == nv (quirk Q9); with quaternions only the tangent-space extension (corrected
       A/B layout, mjo_state_diff) is defined */
    if (m->nq != m->nv && !mjo_ilqr_corrected_layout) return NULL;
    mjo_ilqr* il = (mjo_ilqr*)calloc(1, sizeof(mjo_ilqr));
    il->m = m; il->N = N; il->nq = m->nq; il->nv = m->nv; il->nu = m->nu; il->nx = 2 * m->nv;
    il->nd = m->nv * (2 * m->nv + m->nu) + 2 * m->nv + m->nu;
    il->mu = 1000.0; /* ilqr.h:65 */
    il->cost = cost; il->user = user;
    il->d = mjo_make_data(m);
    il->scratch = mjo_make_data(m);
    int T = N + 1;
    il->qpos = (double*)calloc((size_t)T * il->nq, sizeof(double));
    il->qvel = (double*)calloc((size_t)T * il->nv, sizeof(double));
    il->ctrl = (double*)calloc((size_t)T * il->nu, sizeof(double));
    il->warm = (double*)calloc((size_t)T * il->nv, sizeof(double));
    il->qacc = (double*)calloc((size_t)T * il->nv, sizeof(double));
    il->K = (double*)calloc((size_t)T * il->nu * il->nx, sizeof(double)); /* Q5: zero, not garbage */
    il->k = (double*)calloc((size_t)T * il->nu, sizeof(double));
    il->V = (double*)calloc((size_t)il->nx * il->nx, sizeof(double));
    il->v = (double*)calloc(il->nx, sizeof(double));
    il->deriv = (double*)calloc((size_t)T * il->nd, sizeof(double));
    il->A = (double*)calloc((size_t)il->nx * il->nx, sizeof(double));
    il->B = (double*)calloc((size_t)il->nx * il->nu, sizeof(double));
    return il;
}
void mjo_ilqr_destroy(mjo_ilqr* il) {
    if (!il) return;
    mjo_delete_data(il->d); mjo_delete_data(il->scratch);
    free(il->qpos); free(il->qvel); free(il->ctrl); free(il->warm); free(il->qacc); free(il->K); free(il->k);
    free(il->V); free(il->v); free(il->deriv); free(il->A); free(il->B);
    free(il);
}

static double knot_cost(const mjo_ilqr* il, int n) {
    return il->cost(il->qpos + (size_t)n * il->nq, il->qvel + (size_t)n * il->nv, il->ctrl + (size_t)n * il->nu, il->user);
}
double mjo_ilqr_traj_cost(const mjo_ilqr* il) {
    double J = 0;
    for (int n = il->N; n >= 0; n--) J += knot_cost(il, n);
    return J;
}

/* ILQR::setDInit (ilqr.h:110-113) from raw state */
void mjo_ilqr_set_dinit(mjo_ilqr* il, const double* qpos, const double* qvel, const double* ctrl, const double* warm,
                        const double* qacc) {
    mjo_data* d = il->d;
    memcpy(d->qpos, qpos, sizeof(double) * il->nq);
    memcpy(d->qvel, qvel, sizeof(double) * il->nv);
    memcpy(d->ctrl, ctrl, sizeof(double) * il->nu);
    if (warm) memcpy(d->qacc_warmstart, warm, sizeof(double) * il->nv);
    if (qacc) memcpy(d->qacc, qacc, sizeof(double) * il->nv);
}

/* ILQR constructor body (ilqr.h:72-87): open-loop rollout under dmain's control */
void mjo_ilqr_init(mjo_ilqr* il, const double* qpos, const double* qvel, const double* ctrl, const double* warm) {
    mjo_ilqr_set_dinit(il, qpos, qvel, ctrl, warm, NULL);
    memset(il->d->qacc, 0, sizeof(double) * il->nv);
    for (int n = il->N; n >= 0; n--) {
        knot_store(il, n, il->d);
        mjo_step(il->m, il->d);
    }
    memset(il->K, 0, sizeof(double) * (size_t)(il->N + 1) * il->nu * il->nx);
    memset(il->k, 0, sizeof(double) * (size_t)(il->N + 1) * il->nu);
    il->J = mjo_ilqr_traj_cost(il);
}

/* ILQR::forwardPass (ilqr.h:116-130) with the A10 step size; rolls from il->d, overwrites the nominal */
double mjo_ilqr_forward_pass(mjo_ilqr* il, double alpha) {
    int nx = il->nx, nu = il->nu, nv = il->nv;
    mjo_data* d = il->d;
    double x[2 * ILQG_MAXV];
    for (int n = il->N; n >= 0; n--) {
        const double* xs_q = il->qpos + (size_t)n * il->nq;
        const double* xs_v = il->qvel + (size_t)n * nv;
        const double* us = il->ctrl + (size_t)n * nu;
        const double* K = il->K + (size_t)n * nu * nx;
        const double* k = il->k + (size_t)n * nu;
        mjo_state_diff(il->m, d->qpos, d->qvel, xs_q, xs_v, x);
        double unew[ILQG_MAXU];
        for (int r = 0; r < nu; r++) {
            double s = 0;
            for (int c = 0; c < nx; c++) s += CM(K, r, c, nu) * x[c];
            unew[r] = s + alpha * k[r] + us[r];
        }
        memcpy(d->ctrl, unew, sizeof(double) * nu);
        knot_store(il, n, d);
        mjo_step(il->m, d);
    }
    return mjo_ilqr_traj_cost(il);
}

/* process-wide switches of the opt-in extensions (tests only; the defaults are the reference's behaviour):
   mu schedule: after an iteration whose ladder accepted a step mu <- max(mu_min, mu / factor), after a
   rejected one mu <- min(mu_max, mu * factor); factor <= 1 keeps the reference's constant mu (ilqr.h:65,166) */
double mjo_ilqr_mu_factor = 1.0, mjo_ilqr_mu_min = 1e-6, mjo_ilqr_mu_max = 1e10;
void mjo_ilqr_set_mu_schedule(double factor, double mu_min, double mu_max) {
    mjo_ilqr_mu_factor = factor; mjo_ilqr_mu_min = mu_min; mjo_ilqr_mu_max = mu_max;
}
/* process-wide switch for the opt-in corrected A/B layout (tests only; default 0 = the reference's views) */
int mjo_ilqr_corrected_layout = 0;
void mjo_ilqr_set_corrected_layout(int on) { mjo_ilqr_corrected_layout = on ? 1 : 0; }

/* Differentiator::updateDerivatives at knot n (differentiator.h:85-93): FD, then A/B through the
   column-major views of the row-major deriv blocks (quirk Q1) */
static void linearise_knot(mjo_ilqr* il, int n) {
    int nv = il->nv, nu = il->nu, nx = il->nx;
    double dt = il->m->timestep;
    double* deriv = il->deriv + (size_t)n * il->nd;
    mjo_fd_knot(il->m, il->qpos + (size_t)n * il->nq, il->qvel + (size_t)n * nv, il->ctrl + (size_t)n * nu, il->warm + (size_t)n * nv,
                il->cost, il->user, 1e-6, 30, 3, deriv, NULL, il->scratch);
    double *A = il->A, *B = il->B;
    memset(A, 0, sizeof(double) * nx * nx);
    memset(B, 0, sizeof(double) * nx * nu);
    for (int i = 0; i < nv; i++) { CM(A, i, i, nx) = 1; CM(A, i, nv + i, nx) = dt; }
    const double *dq = deriv, *dv = deriv + nv * nv, *du = deriv + 2 * nv * nv;
    if (mjo_ilqr_corrected_layout) { /* opt-in, not the reference: d qacc_r / d x_c sits at c + r*stride (mjderivative.cpp:107,138,202) */
        for (int r = 0; r < nv; r++)
            for (int c = 0; c < nv; c++) {
                CM(A, nv + r, c, nx) = dq[c + r * nv] * dt;
                CM(A, nv + r, nv + c, nx) = (r == c ? 1.0 : 0.0) + dv[c + r * nv] * dt;
            }
        for (int r = 0; r < nv; r++)
            for (int c = 0; c < nu; c++) CM(B, nv + r, c, nx) = du[c + r * nu] * dt;
        return;
    }
    for (int r = 0; r < nv; r++)
        for (int c = 0; c < nv; c++) {
            CM(A, nv + r, c, nx) = CM(dq, r, c, nv) * dt;
            CM(A, nv + r, nv + c, nx) = (r == c ? 1.0 : 0.0) + CM(dv, r, c, nv) * dt;
        }
    for (int r = 0; r < nv; r++)
        for (int c = 0; c < nu; c++) CM(B, nv + r, c, nx) = CM(du, r, c, nv) * dt;
}

/* dense symmetric-indefinite solve S X = RHS (S is nu x nu): LDL^T with diagonal pivoting, as Eigen's LDLT */
static void ldlt_solve(const double* S, int n, double* X, int ncols) {
    double L[ILQG_MAXU * ILQG_MAXU], D[ILQG_MAXU], W[ILQG_MAXU * ILQG_MAXU];
    int perm[ILQG_MAXU];
    memcpy(W, S, sizeof(double) * n * n);
    for (int i = 0; i < n; i++) perm[i] = i;
    memset(L, 0, sizeof(double) * n * n);
    for (int j = 0; j < n; j++) {
        int p = j;
        for (int i = j + 1; i < n; i++) if (fabs(CM(W, i, i, n)) > fabs(CM(W, p, p, n))) p = i;
        if (p != j) {
            for (int c = 0; c < n; c++) { double t = CM(W, j, c, n); CM(W, j, c, n) = CM(W, p, c, n); CM(W, p, c, n) = t; }
            for (int r = 0; r < n; r++) { double t = CM(W, r, j, n); CM(W, r, j, n) = CM(W, r, p, n); CM(W, r, p, n) = t; }
            for (int c = 0; c < j; c++) { double t = CM(L, j, c, n); CM(L, j, c, n) = CM(L, p, c, n); CM(L, p, c, n) = t; }
            int t = perm[j]; perm[j] = perm[p]; perm[p] = t;
        }
        D[j] = CM(W, j, j, n);
        CM(L, j, j, n) = 1;
        for (int i = j + 1; i < n; i++) CM(L, i, j, n) = CM(W, i, j, n) / D[j];
        for (int r = j + 1; r < n; r++)
            for (int c = j + 1; c < n; c++) CM(W, r, c, n) -= CM(L, r, j, n) * D[j] * CM(L, c, j, n);
    }
    for (int col = 0; col < ncols; col++) {
        double y[ILQG_MAXU];
        for (int i = 0; i < n; i++) y[i] = CM(X, perm[i], col, n);
        for (int i = 0; i < n; i++) for (int c = 0; c < i; c++) y[i] -= CM(L, i, c, n) * y[c];
        for (int i = 0; i < n; i++) y[i] /= D[i];
        for (int i = n - 1; i >= 0; i--) for (int c = i + 1; c < n; c++) y[i] -= CM(L, c, i, n) * y[c];
        for (int i = 0; i < n; i++) CM(X, perm[i], col, n) = y[i];
    }
}

/* ILQR::backwardPass (ilqr.h:133-176), initV (:100-107) */
void mjo_ilqr_backward_pass(mjo_ilqr* il) {
    int nv = il->nv, nu = il->nu, nx = il->nx;
    double *V = il->V, *v = il->v, *A = il->A, *B = il->B;
    /* initV: v = dgdx at knot 0, V = v' v */
    linearise_knot(il, 0);
    {
        const double* q = il->deriv + 2 * nv * nv + nv * nu;
        for (int i = 0; i < nx; i++) v[i] = q[i];
        for (int r = 0; r < nx; r++) for (int c = 0; c < nx; c++) CM(V, r, c, nx) = v[r] * v[c];
    }
    size_t nn = (size_t)nx * nx;
    double* Vs = (double*)malloc(sizeof(double) * nn * 4);
    double *T1 = Vs + nn, *Acl = Vs + 2 * nn, *Vn = Vs + 3 * nn;
    for (int n = 1; n <= il->N; n++) {
        for (int r = 0; r < nx; r++) for (int c = 0; c < nx; c++) CM(Vs, r, c, nx) = (CM(V, r, c, nx) + CM(V, c, r, nx)) / 2;
        memcpy(V, Vs, sizeof(double) * nn);
        linearise_knot(il, n);
        const double* q = il->deriv + (size_t)n * il->nd + 2 * nv * nv + nv * nu;
        const double* r_ = q + 2 * nv;
        double c[2 * ILQG_MAXV];
        mjo_state_diff(il->m, il->qpos + (size_t)(n - 1) * il->nq, il->qvel + (size_t)(n - 1) * nv, il->qpos + (size_t)n * il->nq,
                       il->qvel + (size_t)n * nv, c);
        for (int i = 0; i < nx; i++) CM(V, i, i, nx) += il->mu; /* Q3 */
        /* VB = V B (nx x nu), VA = V A */
        double VB[2 * ILQG_MAXV * ILQG_MAXU], S[ILQG_MAXU * ILQG_MAXU], rhsK[ILQG_MAXU * 2 * ILQG_MAXV], rhsk[ILQG_MAXU];
        for (int r = 0; r < nx; r++) for (int cc = 0; cc < nu; cc++) {
            double s = 0; for (int t = 0; t < nx; t++) s += CM(V, r, t, nx) * CM(B, t, cc, nx);
            CM(VB, r, cc, nx) = s;
        }
        for (int r = 0; r < nx; r++) for (int cc = 0; cc < nx; cc++) {
            double s = 0; for (int t = 0; t < nx; t++) s += CM(V, r, t, nx) * CM(A, t, cc, nx);
            CM(T1, r, cc, nx) = s; /* V A */
        }
        for (int a = 0; a < nu; a++) for (int b = 0; b < nu; b++) {
            double s = 0; for (int t = 0; t < nx; t++) s += CM(B, t, a, nx) * CM(VB, t, b, nx);
            CM(S, a, b, nu) = -2 * s - 2 * r_[a] * r_[b];
        }
        for (int a = 0; a < nu; a++) for (int cc = 0; cc < nx; cc++) {
            double s = 0; for (int t = 0; t < nx; t++) s += CM(B, t, a, nx) * CM(T1, t, cc, nx);
            CM(rhsK, a, cc, nu) = 2 * s;
        }
        for (int a = 0; a < nu; a++) {
            double s = 0;
            for (int t = 0; t < nx; t++) {
                double vc = 0; for (int e = 0; e < nx; e++) vc += CM(V, t, e, nx) * c[e];
                s += CM(B, t, a, nx) * (v[t] + 2 * vc);
            }
            rhsk[a] = s + r_[a];
        }
        double* K = il->K + (size_t)n * nu * nx;
        double* k = il->k + (size_t)n * nu;
        ldlt_solve(S, nu, rhsK, nx);
        ldlt_solve(S, nu, rhsk, 1);
        memcpy(K, rhsK, sizeof(double) * nu * nx);
        memcpy(k, rhsk, sizeof(double) * nu);
        /* Acl = A + B K ; V <- Acl' V Acl + Q + K' R K */
        for (int r = 0; r < nx; r++) for (int cc = 0; cc < nx; cc++) {
            double s = CM(A, r, cc, nx); for (int a = 0; a < nu; a++) s += CM(B, r, a, nx) * CM(K, a, cc, nu);
            CM(Acl, r, cc, nx) = s;
        }
        for (int r = 0; r < nx; r++) for (int cc = 0; cc < nx; cc++) {
            double s = 0; for (int t = 0; t < nx; t++) s += CM(V, r, t, nx) * CM(Acl, t, cc, nx);
            CM(T1, r, cc, nx) = s; /* V Acl */
        }
        double rK[2 * ILQG_MAXV]; /* r K (1 x nx) */
        for (int cc = 0; cc < nx; cc++) { double s = 0; for (int a = 0; a < nu; a++) s += r_[a] * CM(K, a, cc, nu); rK[cc] = s; }
        for (int r = 0; r < nx; r++) for (int cc = 0; cc < nx; cc++) {
            double s = 0; for (int t = 0; t < nx; t++) s += CM(Acl, t, r, nx) * CM(T1, t, cc, nx);
            CM(Vn, r, cc, nx) = s + q[r] * q[cc] + rK[r] * rK[cc];
        }
        memcpy(V, Vn, sizeof(double) * nn);
        /* v <- 2 (k'B' + c') V_new Acl + v Acl + q + 2 k' R K   (Q4: V already updated) */
        double w_[2 * ILQG_MAXV], wV[2 * ILQG_MAXV], vn[2 * ILQG_MAXV];
        for (int t = 0; t < nx; t++) { double s = c[t]; for (int a = 0; a < nu; a++) s += CM(B, t, a, nx) * k[a]; w_[t] = s; }
        for (int cc = 0; cc < nx; cc++) { double s = 0; for (int t = 0; t < nx; t++) s += w_[t] * CM(V, t, cc, nx); wV[cc] = s; }
        double kr = 0; for (int a = 0; a < nu; a++) kr += k[a] * r_[a];
        for (int cc = 0; cc < nx; cc++) {
            double s1 = 0, s2 = 0;
            for (int t = 0; t < nx; t++) { s1 += wV[t] * CM(Acl, t, cc, nx); s2 += v[t] * CM(Acl, t, cc, nx); }
            vn[cc] = 2 * s1 + s2 + q[cc] + 2 * kr * rK[cc];
        }
        memcpy(v, vn, sizeof(double) * nx);
    }
    free(Vs);
}

/* ILQR::iterate (ilqr.h:179-186) */
double mjo_ilqr_iterate(mjo_ilqr* il) {
    il->J = mjo_ilqr_forward_pass(il, 1.0);
    int N = il->N;
    mjo_ilqr_set_dinit(il, il->qpos + (size_t)N * il->nq, il->qvel + (size_t)N * il->nv, il->ctrl + (size_t)N * il->nu,
                       il->warm + (size_t)N * il->nv, il->qacc + (size_t)N * il->nv);
    mjo_ilqr_backward_pass(il);
    return il->J;
}

/* A10: one iteration with a backtracking ladder.  Returns the index of the accepted alpha (-1: none). */
int mjo_ilqr_iterate_linesearch(mjo_ilqr* il, const double* alphas, int nalpha, int accept_always, double* J_out) {
    int N = il->N, T = N + 1;
    size_t sq = (size_t)T * il->nq, sv = (size_t)T * il->nv, su = (size_t)T * il->nu;
    double* save = (double*)malloc(sizeof(double) * (sq + 3 * sv + su));
    double *s_q = save, *s_v = s_q + sq, *s_u = s_v + sv, *s_w = s_u + su, *s_a = s_w + sv;
    memcpy(s_q, il->qpos, sizeof(double) * sq); memcpy(s_v, il->qvel, sizeof(double) * sv); memcpy(s_u, il->ctrl, sizeof(double) * su);
    memcpy(s_w, il->warm, sizeof(double) * sv); memcpy(s_a, il->qacc, sizeof(double) * sv);
    /* il->d at entry is the state the pass must start from (set by setDInit: the nominal's first knot, or a new
       measured state at the start of an MPC step) */
    double i_q[ILQG_MAXQ], i_v[ILQG_MAXV], i_u[ILQG_MAXU], i_w[ILQG_MAXV], i_a[ILQG_MAXV];
    memcpy(i_q, il->d->qpos, sizeof(double) * il->nq); memcpy(i_v, il->d->qvel, sizeof(double) * il->nv);
    memcpy(i_u, il->d->ctrl, sizeof(double) * il->nu); memcpy(i_w, il->d->qacc_warmstart, sizeof(double) * il->nv);
    memcpy(i_a, il->d->qacc, sizeof(double) * il->nv);
    int accepted = -1;
    double Jprev = il->J;
    for (int a = 0; a < nalpha; a++) {
        memcpy(il->qpos, s_q, sizeof(double) * sq); memcpy(il->qvel, s_v, sizeof(double) * sv); memcpy(il->ctrl, s_u, sizeof(double) * su);
        memcpy(il->warm, s_w, sizeof(double) * sv); memcpy(il->qacc, s_a, sizeof(double) * sv);
        mjo_ilqr_set_dinit(il, i_q, i_v, i_u, i_w, i_a);
        double J = mjo_ilqr_forward_pass(il, alphas[a]);
        if (accept_always || J < Jprev) { accepted = a; il->J = J; break; }
    }
    if (accepted < 0) { /* keep the nominal */
        memcpy(il->qpos, s_q, sizeof(double) * sq); memcpy(il->qvel, s_v, sizeof(double) * sv); memcpy(il->ctrl, s_u, sizeof(double) * su);
        memcpy(il->warm, s_w, sizeof(double) * sv); memcpy(il->qacc, s_a, sizeof(double) * sv);
    }
    free(save);
    if (mjo_ilqr_mu_factor > 1.0) {
        if (accepted >= 0) { il->mu = il->mu / mjo_ilqr_mu_factor; if (il->mu < mjo_ilqr_mu_min) il->mu = mjo_ilqr_mu_min; }
        else { il->mu = il->mu * mjo_ilqr_mu_factor; if (il->mu > mjo_ilqr_mu_max) il->mu = mjo_ilqr_mu_max; }
    }
    mjo_ilqr_set_dinit(il, il->qpos + (size_t)N * il->nq, il->qvel + (size_t)N * il->nv, il->ctrl + (size_t)N * il->nu,
                       il->warm + (size_t)N * il->nv, il->qacc + (size_t)N * il->nv);
    mjo_ilqr_backward_pass(il);
    if (J_out) *J_out = il->J;
    return accepted;
}

/* accessors for ctypes */
void mjo_ilqr_get(const mjo_ilqr* il, double* qpos, double* qvel, double* ctrl, double* K, double* k, double* V, double* v,
                  double* deriv) {
    int T = il->N + 1;
    if (qpos) memcpy(qpos, il->qpos, sizeof(double) * (size_t)T * il->nq);
    if (qvel) memcpy(qvel, il->qvel, sizeof(double) * (size_t)T * il->nv);
    if (ctrl) memcpy(ctrl, il->ctrl, sizeof(double) * (size_t)T * il->nu);
    if (K) memcpy(K, il->K, sizeof(double) * (size_t)T * il->nu * il->nx);
    if (k) memcpy(k, il->k, sizeof(double) * (size_t)T * il->nu);
    if (V) memcpy(V, il->V, sizeof(double) * (size_t)il->nx * il->nx);
    if (v) memcpy(v, il->v, sizeof(double) * il->nx);
    if (deriv) memcpy(deriv, il->deriv, sizeof(double) * (size_t)T * il->nd);
}
void mjo_ilqr_set_mu(mjo_ilqr* il, double mu) { il->mu = mu; }

typedef struct mjo_quad_ctx2 { const ilqg_model* m; const ilqg_cost* c; } mjo_quad_ctx2;

/* batch driver for ctypes / bench: ninst independent problems, reference cadence
   (InvertedPendulum::forward, inverted_pendulum.cpp:19-30, without the final mj_step):
   init from (qpos,qvel,ctrl,warm), then niter x iterate.  Outputs the cost trace [ninst*niter],
   accepted alpha indices, and the final nominal / gains of every instance. */
void mjo_ilqr_run_batch(const ilqg_model* m, int ninst, int N, int niter, const double* qpos0, const double* qvel0, const double* ctrl0,
                        const double* warm0, const ilqg_cost* cost, const double* alphas, int nalpha, int accept_always, double mu,
                        double* Jtrace, int* accepted, double* qpos, double* qvel, double* ctrl, double* K, double* k, double* V,
                        double* v, int nthreads) {
    (void)nthreads;
    int T = N + 1, nq = m->nq, nv = m->nv, nu = m->nu, nx = 2 * nv;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < ninst; i++) {
        mjo_quad_ctx2 ctx = {m, cost};
        mjo_ilqr* il = mjo_ilqr_create(m, N, mjo_cost_quadratic, &ctx);
        il->mu = mu;
        mjo_ilqr_init(il, qpos0 + (size_t)i * nq, qvel0 + (size_t)i * nv, ctrl0 + (size_t)i * nu, warm0 ? warm0 + (size_t)i * nv : NULL);
        /* InvertedPendulum::forward: setDInit(d) — the same initial state — before the iterations */
        mjo_ilqr_set_dinit(il, qpos0 + (size_t)i * nq, qvel0 + (size_t)i * nv, ctrl0 + (size_t)i * nu, warm0 ? warm0 + (size_t)i * nv : NULL, NULL);
        for (int it = 0; it < niter; it++) {
            double J;
            int acc = 0;
            if (nalpha <= 0) J = mjo_ilqr_iterate(il);
            else acc = mjo_ilqr_iterate_linesearch(il, alphas, nalpha, accept_always, &J);
            if (Jtrace) Jtrace[(size_t)i * niter + it] = J;
            if (accepted) accepted[(size_t)i * niter + it] = acc;
        }
        mjo_ilqr_get(il, qpos ? qpos + (size_t)i * T * nq : NULL, qvel ? qvel + (size_t)i * T * nv : NULL, ctrl ? ctrl + (size_t)i * T * nu : NULL,
                     K ? K + (size_t)i * T * nu * nx : NULL, k ? k + (size_t)i * T * nu : NULL, V ? V + (size_t)i * nx * nx : NULL,
                     v ? v + (size_t)i * nx : NULL, NULL);
        mjo_ilqr_destroy(il);
    }
}

/* Restated MPC demo (InvertedPendulum ctor + nmpc x forward(), inverted_pendulum.cpp:6-30; the headless core of
   cmd/basic.cpp:155-164).  trace[s] = (qpos, qvel, ctrl) of the simulated system after MPC step s;
   Jtrace[s*niter + it] = trajectory cost after the forward pass of iteration `it` (A10 bookkeeping). */
void mjo_mpc_run(const ilqg_model* m, const double* qpos0, const double* qvel0, int nwarm, int N, int niter, int nmpc,
                 const ilqg_cost* cost, const double* alphas, int nalpha, int accept_always, double* trace, double* Jtrace,
                 int* accepted, double* nom_qpos, double* nom_qvel, double* nom_ctrl, double* K, double* k, double* V, double* v) {
    mjo_quad_ctx2 ctx = {m, cost};
    int nq = m->nq, nv = m->nv, nu = m->nu;
    mjo_data* d = mjo_make_data(m);
    if (qpos0) memcpy(d->qpos, qpos0, sizeof(double) * nq);
    if (qvel0) memcpy(d->qvel, qvel0, sizeof(double) * nv);
    for (int i = 0; i < nwarm; i++) mjo_step(m, d);
    mjo_ilqr* il = mjo_ilqr_create(m, N, mjo_cost_quadratic, &ctx);
    /* ILQR ctor: cpMjData(d_ilqr, dmain) copies qacc too */
    mjo_ilqr_set_dinit(il, d->qpos, d->qvel, d->ctrl, d->qacc_warmstart, d->qacc);
    for (int n = N; n >= 0; n--) { knot_store(il, n, il->d); mjo_step(m, il->d); }
    il->J = mjo_ilqr_traj_cost(il);
    int stride = nq + nv + nu;
    for (int s = 0; s < nmpc; s++) {
        mjo_ilqr_set_dinit(il, d->qpos, d->qvel, d->ctrl, d->qacc_warmstart, d->qacc);
        for (int it = 0; it < niter; it++) {
            double J;
            int acc = 0;
            if (nalpha <= 0) J = mjo_ilqr_iterate(il);
            else acc = mjo_ilqr_iterate_linesearch(il, alphas, nalpha, accept_always, &J);
            if (Jtrace) Jtrace[(size_t)s * niter + it] = J;
            if (accepted) accepted[(size_t)s * niter + it] = acc;
        }
        memcpy(d->ctrl, il->ctrl + (size_t)N * nu, sizeof(double) * nu);
        mjo_step(m, d);
        if (trace) {
            memcpy(trace + (size_t)s * stride, d->qpos, sizeof(double) * nq);
            memcpy(trace + (size_t)s * stride + nq, d->qvel, sizeof(double) * nv);
            memcpy(trace + (size_t)s * stride + nq + nv, d->ctrl, sizeof(double) * nu);
        }
    }
    mjo_ilqr_get(il, nom_qpos, nom_qvel, nom_ctrl, K, k, V, v, NULL);
    mjo_ilqr_destroy(il);
    mjo_delete_data(d);
}
