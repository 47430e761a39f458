// coop.cuh — warp-cooperative fp64 forward dynamics and FD linearisation for models that do not fit the
// thread-per-rollout register budget (humanoid: nv = 27, 150 perturbed evaluations per knot, 161 collision pairs).
//
// One WARP per rollout (north star (3): "one warp or CTA per perturbed rollout"): the rollout's mjData-equivalent lives in
// that warp's slice of SHARED memory; lanes stride over independent items (bodies of one tree level, dofs, geoms, collision
// pairs, constraint rows, matrix entries) and meet at __syncwarp().  Triangular solves keep the vector in registers
// (lane i = dof i) and broadcast with shuffles.  The model is read from a GModel block in global memory (uniform or
// lane-strided reads, L1/L2 resident).  Nothing is specialised on the kinematic tree: any model of the subset runs here.
//
// The state is split the way MuJoCo splits its pipeline (and the reference its skip levels, mjderivative.cpp:92,124,178):
//   C-state  products of the POSITION stage (motion axes, spatial inertias, M, its factor, constraint rows J/D/B/k-term) plus
//            the knot's inputs and the centre's velocity-stage products — 27 KB for the humanoid
//   private  velocity stage + constraint solve working set (bias recursion, qfrc/qacc vectors, aref/jar/jv, Newton Hessian)
//            — 11 KB; the position stage's temporaries (frames, geoms, contact records) alias it
// FD of a knot = three launches:
//   coop_center_kernel   warp per knot: full pipeline + warm-up solves; leaves the C-state in HBM and the list of collision
//                        pairs within (margin + slack) of contact — the only pairs a +-eps perturbation can activate
//   coop_velctrl_kernel  CTA per knot: loads the centre's C-state once; its warps share it read-only and take the qvel / ctrl
//                        columns in turn, each re-running only the velocity stage (or only the actuation) and the solve.
//                        7 warps x 11 KB private + 28 KB shared: 14 resident warps per SM instead of 4.
//   coop_qpos_kernel     warp per (knot, qpos column): full pipeline at +eps and -eps, narrow phase on the candidate list only.
//
// Replaces, for such models, the same reference calls as dyn.cuh: mj_forward / mj_forwardSkip inside
// /root/reference/src/mjderivative.cpp:64-198 (incl. the quaternion tangent perturbation :152-169,187-192).
#pragma once
#include <cuda_runtime.h>

#include "../../include/ilqg_b200.h"
#include "dyn.cuh"
#include "ilqr.cuh"

namespace ilqg {

#define COOP_MAXCON 24
#define COOP_MAXEFC 72
#define COOP_MAXLEVEL 16
#define COOP_MAXCAND 63   // candidate pairs recorded per knot (more: the perturbed evaluations test every pair)

// flat model + host-precomputed helpers, uploaded once per handle
struct GModel {
    ilqg_model m;
    double geom_axis[ILQG_MAXGEOM][3];
    double jnt_K[ILQG_MAXJNT], jnt_B[ILQG_MAXJNT], jnt_imp[ILQG_MAXJNT];
    double pair_K[ILQG_MAXPAIR], pair_B[ILQG_MAXPAIR], pair_imp[ILQG_MAXPAIR];
    unsigned body_dofmask[ILQG_MAXBODY];   // bit i set: dof i moves this body (nv <= 32)
    int nlevel, level_start[COOP_MAXLEVEL + 1], level_body[ILQG_MAXBODY];  // bodies grouped by tree depth
    int any_damping;
    int dof_act[ILQG_MAXV];   // actuator driving dof i, -1: none, -2: several (loop over the actuators)
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
    for (int i = 0; i < s.nv; i++) g.dof_act[i] = -1;
    for (int a = 0; a < s.nu; a++) { int d = s.act_dofid[a]; g.dof_act[d] = g.dof_act[d] == -1 ? a : -2; }
    g.any_damping = 0;
    for (int i = 0; i < s.nv; i++) g.any_damping |= s.dof_damping[i] > 0;
    return true;
}

// ------------------------------------------------------------------ shared-memory views
struct CoopMem {
    // ---- C-state (contiguous; copied to / from HBM as one block)
    double *q, *v, *u, *center;            // knot inputs, centre qacc (warm start of the perturbed solves)
    double *cdof, *cinert, *com, *dspr;    // motion axes, spatial inertias about the tree com, tree com, q - qpos_spring
    double *M, *L;                         // packed lower triangles; L's diagonal holds 1 / L_ii
    double *J, *D, *rB, *rkt;              // constraint rows: Jacobian, 1/R, damping B, K*imp*(pos - margin)
    double *fb0, *aref0;                   // the centre's velocity-stage products (ctrl columns reuse them)
    double* Hc;                            // factor of the Newton Hessian M + J_A' D_A J_A for the centre solution's active set A
    int* hdr;                              // hdr[0] = nefc, hdr[1] = capacity ok, hdr[2] = Hc valid, hdr[3] = row bound of the knot's +-eps
                                           // neighbourhood (centre only), hdr[4..7] = active-set mask of Hc
    int maxefc;                            // row capacity of this block
    int ncon;                              // contact points of the last position stage (diagnostics)
    // ---- private working set
    double *pv, *pu;                       // perturbed qvel / ctrl
    double *cvel, *cacc, *cfrc, *cdofdot;
    double *fb, *fs, *as, *fc, *qacc, *warm, *Ma, *grad, *search, *Mv;
    double *aref, *jar, *jv, *H;
    int* alist;                            // active rows of the current Newton iteration
    // ---- position-stage temporaries (alias the private block)
    double *xpos, *xquat, *xmat, *xipos, *anchor, *axis, *gpos, *gax, *crb, *cdist, *cpos, *cframe;
    int* cpair;
};

__host__ __device__ inline int coop_nt(int nv) { return nv * (nv + 1) / 2; }
// `maxefc`: row capacity of the block (COOP_MAXEFC everywhere except the qpos-column kernel, which is launched once per capacity
// class with the knots whose row bound fits — the rows are the largest part of a rollout's shared-memory state)
// with_L = false: a block without the Cholesky factor of M (the qpos-column kernel: its evaluations start from the centre's
// solution and never need qacc_smooth, see coop_solve's warm_only)
__host__ __device__ inline size_t coop_cstate_doubles(const ilqg_model& m, bool full, int maxefc = COOP_MAXEFC, bool with_L = true) {
    const size_t nv = m.nv, nb = m.nbody;
    size_t n = m.nq + nv + m.nu + nv;                       // q v u center
    n += 6 * nv + 10 * nb + 3 * nb + nv;                    // cdof cinert com dspr
    n += (with_L ? 2 : 1) * (size_t)coop_nt(m.nv);          // M L
    n += (size_t)maxefc * nv + 3 * maxefc;                  // J D rB rkt
    n += 4;                                                 // hdr (8 ints)
    n = (n + 1) & ~(size_t)1;
    if (!full) return n;                                    // what a stand-alone rollout needs
    n += nv + maxefc;                                       // fb0 aref0   } products of the centre evaluation that
    n += coop_nt(m.nv);                                     // Hc          } the qvel / ctrl columns reuse
    return (n + 1) & ~(size_t)1;
}
__host__ __device__ inline size_t coop_priv_doubles(const ilqg_model& m, int maxefc = COOP_MAXEFC) {
    const size_t nv = m.nv, nb = m.nbody, nj = m.njnt, ng = m.ngeom;
    size_t rne = 18 * nb + 6 * nv, hes = coop_nt(m.nv);   // the bias recursion's scratch is dead before the Newton Hessian is built
    size_t priv = nv + m.nu + (rne > hes ? rne : hes) + 10 * nv + 3 * (size_t)maxefc + ((size_t)maxefc + 1) / 2;
    size_t tmp = 19 * nb + 6 * nj + 6 * ng + 10 * nb + (size_t)COOP_MAXCON * 13 + (COOP_MAXCON + 1) / 2;
    size_t n = priv > tmp ? priv : tmp;
    return (n + 1) & ~(size_t)1;
}

DEV void coop_carve_cstate(CoopMem& w, double* base, const ilqg_model& m, int maxefc = COOP_MAXEFC, bool with_L = true) {
    w.maxefc = maxefc;
    const int nq = m.nq, nv = m.nv, nu = m.nu, nb = m.nbody, nt = coop_nt(m.nv);
    double* p = base;
    auto take = [&](size_t n) { double* r = p; p += n; return r; };
    w.q = take(nq); w.v = take(nv); w.u = take(nu); w.center = take(nv);
    w.cdof = take(6 * nv); w.cinert = take(10 * nb); w.com = take(3 * nb); w.dspr = take(nv);
    w.M = take(nt); w.L = with_L ? take(nt) : nullptr;
    w.J = take((size_t)maxefc * nv); w.D = take(maxefc); w.rB = take(maxefc); w.rkt = take(maxefc);
    w.hdr = reinterpret_cast<int*>(p);
    p = base + coop_cstate_doubles(m, false, maxefc, with_L);
    w.fb0 = take(nv); w.aref0 = take(maxefc); w.Hc = take(nt);   // only valid where the full C-state is allocated
}
DEV void coop_carve_priv(CoopMem& w, double* base, const ilqg_model& m, int maxefc = COOP_MAXEFC) {
    const int nv = m.nv, nu = m.nu, nb = m.nbody, nj = m.njnt, ng = m.ngeom, nt = coop_nt(m.nv);
    double* p = base;
    auto take = [&](size_t n) { double* r = p; p += n; return r; };
    w.pv = take(nv); w.pu = take(nu);
    {
        const size_t rne = 18 * (size_t)nb + 6 * (size_t)nv;
        double* blk = take(rne > (size_t)nt ? rne : (size_t)nt);
        w.cvel = blk; w.cacc = blk + 6 * nb; w.cfrc = blk + 12 * nb; w.cdofdot = blk + 18 * nb;
        w.H = blk;   // Newton Hessian: built after the velocity stage has consumed its scratch
    }
    w.fb = take(nv); w.fs = take(nv); w.as = take(nv); w.fc = take(nv); w.qacc = take(nv); w.warm = take(nv);
    w.Ma = take(nv); w.grad = take(nv); w.search = take(nv); w.Mv = take(nv);
    w.aref = take(maxefc); w.jar = take(maxefc); w.jv = take(maxefc);
    w.alist = reinterpret_cast<int*>(p);
    // temporaries of the position stage share the same bytes
    p = base;
    w.xpos = take(3 * nb); w.xquat = take(4 * nb); w.xmat = take(9 * nb); w.xipos = take(3 * nb);
    w.anchor = take(3 * nj); w.axis = take(3 * nj); w.gpos = take(3 * ng); w.gax = take(3 * ng); w.crb = take(10 * nb);
    w.cdist = take(COOP_MAXCON); w.cpos = take(COOP_MAXCON * 3); w.cframe = take(COOP_MAXCON * 9);
    w.cpair = reinterpret_cast<int*>(p);
}

DEV double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__host__ __device__ constexpr int ptri(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i

#define COOP_NMAX 32   // lane i owns dof i / row i of the nv x nv matrices

// fp64 tensor-core tile product (mma.sync m8n8k4): {d0, d1} += A(8x4) B(4x8).  Fragment layout (PTX ISA, .f64):
// a: row = lane / 4, k = lane % 4;  b: k = lane % 4, col = lane / 4;  c/d: row = lane / 4, cols = 2 (lane % 4) + {0, 1}.
DEV void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// In-place Cholesky of the packed lower triangle A (n <= 32); the diagonal is left holding 1 / L_jj.
// Blocked left-looking over 8-wide panels: before a panel is factored, the contributions of all earlier panels are taken off
// its block column by DMMA tile updates (20 tensor-core instructions for n = 27 instead of ~260 of the 351 scalar
// multiply-adds per row-lane); the panel itself needs no cross-lane step (see below): 4 warp barriers per factorisation
// instead of 27 shuffle / rsqrt / barrier rounds.
// (Measured alternatives that LOST on B200: rows in registers with every loop unrolled — instruction-cache misses, the SM's
// warps sit at unrelated points of the pipeline; manual 4-way unrolling with split accumulators.)
DEV void coop_chol(double* A, int n, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const int nblk = (n + 7) >> 3;
    for (int p = 0; p < nblk; p++) {
        const int c0 = 8 * p;
        if (p > 0) {
            const int jb = c0 + g;                                  // row of L that feeds B's column g
            for (int ib = p; ib < nblk; ib++) {
                const int i = 8 * ib + g, j = c0 + 2 * t;
                const bool v0 = i < n && j <= i, v1 = i < n && j + 1 <= i;
                double d0 = v0 ? A[ptri(i, j)] : 0.0, d1 = v1 ? A[ptri(i, j + 1)] : 0.0;
                for (int k0 = 0; k0 < c0; k0 += 4) {
                    const double a = i < n ? -A[ptri(i, k0 + t)] : 0.0;     // k < c0 <= i, jb: strictly lower entries (true L values)
                    const double b = jb < n ? A[ptri(jb, k0 + t)] : 0.0;
                    dmma_8x8x4(d0, d1, a, b);
                }
                if (v0) A[ptri(i, j)] = d0;
                if (v1) A[ptri(i, j + 1)] = d1;
            }
            __syncwarp();
        }
        // The 8 x 8 diagonal tile is factored by EVERY lane in registers (same arithmetic, same result): no shuffles, no
        // barriers inside the panel.  Each row below the panel is then one lane's private forward substitution against it.
        double T[8][8], rinv[8];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int bb = 0; bb <= a; bb++) T[a][bb] = c0 + a < n ? A[ptri(c0 + a, c0 + bb)] : (a == bb ? 1.0 : 0.0);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            double d = T[j][j];
#pragma unroll
            for (int k = 0; k < j; k++) d -= T[j][k] * T[j][k];
            if (d < ILQG_MINVAL) d = ILQG_MINVAL;
            rinv[j] = rsqrt(d);
#pragma unroll
            for (int a = j + 1; a < 8; a++) {
                double sa = T[a][j];
#pragma unroll
                for (int k = 0; k < j; k++) sa -= T[a][k] * T[j][k];
                T[a][j] = sa * rinv[j];
            }
        }
        const int i = c0 + 8 + lane;
        if (i < n) {
            double* ri = A + ptri(i, c0);
            double x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double sx = ri[j];
#pragma unroll
                for (int k = 0; k < j; k++) sx -= x[k] * T[j][k];
                x[j] = sx * rinv[j];
            }
#pragma unroll
            for (int j = 0; j < 8; j++) ri[j] = x[j];
        }
#pragma unroll
        for (int a = 0; a < 8; a++)
            if (lane == a && c0 + a < n) {
                double* ra = A + ptri(c0 + a, c0);
#pragma unroll
                for (int bb = 0; bb < a; bb++) ra[bb] = T[a][bb];
                ra[a] = rinv[a];
            }
        __syncwarp();
    }
}
// x <- (L L^T)^-1 x with x in registers: lane i holds x_i (n <= 32); returns this lane's component
DEV double coop_chol_solve(const double* L, double x, int n, int lane) {
    for (int j = 0; j < n; j++) {
        const double xj = __shfl_sync(0xffffffffu, x, j) * L[ptri(j, j)];
        if (lane == j) x = xj;
        else if (lane > j && lane < n) x -= L[ptri(lane, j)] * xj;
    }
    for (int j = n - 1; j >= 0; j--) {
        const double xj = __shfl_sync(0xffffffffu, x, j) * L[ptri(j, j)];
        if (lane == j) x = xj;
        else if (lane < j) x -= L[ptri(j, lane)] * xj;
    }
    return x;
}
// y_i = sum_k M_ik x_k for the packed symmetric M; lane i < n returns y_i
DEV double coop_symv(const double* M, const double* x, int n, int lane) {
    double s = 0;
    if (lane < n) {
        const double* row = M + ptri(lane, 0);
        for (int k = 0; k <= lane; k++) s += row[k] * x[k];
        int idx = ptri(lane + 1, lane);
        for (int k = lane + 1; k < n; k++) { s += M[idx] * x[k]; idx += k + 1; }
    }
    return s;
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

// ------------------------------------------------------------------ position stage
// In: w.q.  Out: the C-state's position products and hdr.  The narrow phase runs over `cand` (ncand pair indices) when
// ncand >= 0, else over every pair; with cand_out != NULL (centre evaluation) it also records, in pair order, the pairs
// within margin + slack of contact: cand_out[0] = count (or -1 when more than COOP_MAXCAND), cand_out[1..] = pair indices.
DEV void coop_pos(const GModel* __restrict__ g, CoopMem& w, int lane, const int* __restrict__ cand, int ncand, int* __restrict__ cand_out,
                  double slack, bool factor_M = true) {
    const ilqg_model& m = g->m;
    const int nv = m.nv, nb = m.nbody, nj = m.njnt, ng = m.ngeom, nt = coop_nt(m.nv);
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
                if (ty == ILQG_JNT_BALL) {   // rotate about the anchor by the joint's own (normalised) quaternion
                    quat = qmul(quat, qnormalized({w.q[qa], w.q[qa + 1], w.q[qa + 2], w.q[qa + 3]}));
                    pos = an - mulv(q2m(quat), gl3(m.jnt_pos[j]));
                    continue;
                }
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
            for (int i = 0; i < 6; i++) w.dspr[da + i] = 0;
        } else if (ty == ILQG_JNT_BALL) {   // the body's three axes about the anchor (mj_comPos)
            for (int i = 0; i < 3; i++) {
                V3 ax = {w.xmat[9 * b + i], w.xmat[9 * b + 3 + i], w.xmat[9 * b + 6 + i]};
                gs3(w.cdof + 6 * (da + i), ax);
                gs3(w.cdof + 6 * (da + i) + 3, cross(ax, off));
                w.dspr[da + i] = 0;
            }
        } else {
            if (ty == ILQG_JNT_SLIDE) {
                gs3(w.cdof + 6 * da, {0, 0, 0});
                gs3(w.cdof + 6 * da + 3, gl3(w.axis + 3 * j));
            } else {
                V3 ax = gl3(w.axis + 3 * j);
                gs3(w.cdof + 6 * da, ax);
                gs3(w.cdof + 6 * da + 3, cross(ax, off));
            }
            w.dspr[da] = m.jnt_stiffness[j] != 0 ? w.q[m.jnt_qposadr[j]] - m.qpos_spring[m.jnt_qposadr[j]] : 0.0;
        }
    }
    // ---- composite inertias: one lane per component, serial over bodies (children have larger ids than parents)
    for (int k = lane; k < 10; k += 32) w.crb[k] = 0;
    for (int e = 10 + lane; e < 10 * nb; e += 32) w.crb[e] = w.cinert[e];
    for (int e = lane; e < nt; e += 32) w.M[e] = 0;
    __syncwarp();
    if (lane < 10)
        for (int b = nb - 1; b > 0; b--) { int p = m.body_parentid[b]; if (p > 0) w.crb[10 * p + lane] += w.crb[10 * b + lane]; }
    __syncwarp();
    // ---- mass matrix (packed lower) and its factor
    for (int i = lane; i < nv; i += 32) {
        const int b = m.dof_bodyid[i];
        double buf[6];
        g_inert_vec(buf, w.crb + 10 * b, w.cdof + 6 * i);
        for (int a = i; a >= 0; a = m.dof_parentid[a]) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += w.cdof[6 * a + k] * buf[k];
            if (a == i) s += m.dof_armature[i];
            w.M[ptri(i, a)] = s;
        }
    }
    __syncwarp();
    if (factor_M) {
        for (int e = lane; e < nt; e += 32) w.L[e] = w.M[e];
        __syncwarp();
        coop_chol(w.L, nv, lane);
    }
    // ---- constraint rows: joint limits
    // rowbound: the rows an evaluation within +-eps of this one can have at most — limits within margin + slack, and for every
    // pair within margin + slack of contact the most contacts its narrow phase can return times its rows per contact
    int ne = 0, rowbound = 0;
    for (int j0 = 0; j0 < nj; j0 += 32) {
        const int j = j0 + lane;
        int side = 0;
        bool nearlim = false;
        double dist = 0;
        if (j < nj && m.jnt_limited[j] && m.jnt_type[j] != ILQG_JNT_FREE && m.jnt_type[j] != ILQG_JNT_BALL) {
            double value = w.q[m.jnt_qposadr[j]];
            double dlo = value - m.jnt_range[j][0], dhi = m.jnt_range[j][1] - value;
            if (dlo < m.jnt_margin[j]) { side = -1; dist = dlo; }
            else if (dhi < m.jnt_margin[j]) { side = 1; dist = dhi; }
            nearlim = dlo < m.jnt_margin[j] + slack || dhi < m.jnt_margin[j] + slack;
        }
        rowbound += __popc(__ballot_sync(0xffffffffu, nearlim));
        unsigned bal = __ballot_sync(0xffffffffu, side != 0);
        int r = ne + __popc(bal & ((1u << lane) - 1u));
        if (side != 0 && r < w.maxefc) {
            const int da = m.jnt_dofadr[j];
            double R, kt;
            row_params(g->jnt_K[j], g->jnt_imp[j], m.jnt_solimp[j], dist, m.jnt_margin[j], m.dof_invweight0[da], R, kt);
            for (int i = 0; i < nv; i++) w.J[r * nv + i] = 0;
            w.J[r * nv + da] = -side;
            w.D[r] = 1.0 / R;
            w.rB[r] = g->jnt_B[j];
            w.rkt[r] = kt;
        }
        ne += __popc(bal);
    }
    // ---- collision: lanes over candidate pairs, contacts appended through warp prefix sums
    int ncon = 0, nrec = 0;
    const int npairs = ncand >= 0 ? ncand : m.npair;
    for (int p0 = 0; p0 < npairs; p0 += 32) {
        const int pi = p0 + lane;
        int cnt = 0, p = -1;
        bool near = false;
        double cd[2], cp[2][3], cn[2][3], ch[3] = {0, 0, 0};
        bool hint = false;
        if (pi < npairs) {
            p = ncand >= 0 ? cand[pi] : pi;
            const int g1 = m.pair_geom1[p], g2 = m.pair_geom2[p], t1 = m.geom_type[g1], t2 = m.geom_type[g2];
            const double margin = m.pair_margin[p], wide = margin + slack;
            V3 x1 = gl3(w.gpos + 3 * g1), x2 = gl3(w.gpos + 3 * g2), a1 = gl3(w.gax + 3 * g1), a2 = gl3(w.gax + 3 * g2);
            auto sphere_sphere = [&](V3 c1, double r1, V3 c2, double r2) {
                V3 n = c2 - c1;
                double len = sqrt(dot(n, n)), dist = len - r1 - r2;
                near = near || dist <= wide;
                if (dist > margin || cnt >= 2) return;
                n = len < ILQG_MINVAL ? V3{1, 0, 0} : (1.0 / len) * n;
                V3 pos = c1 + (r1 + 0.5 * dist) * n;
                cd[cnt] = dist; gs3(cp[cnt], pos); gs3(cn[cnt], n); cnt++;
            };
            auto plane_sphere = [&](V3 c, double r) {
                double dist = dot(c - x1, a1) - r;
                near = near || dist <= wide;
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
                    if (fabs(det) < 1e-9) near = true;   // close to the parallel-axes branch: keep the pair for the perturbed evaluations
                } else {
                    near = true;
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
        if (cand_out) {   // record the near pairs in pair order
            unsigned nb_ = __ballot_sync(0xffffffffu, near);
            int prow = 0;
            if (near) {
                const int t1 = m.geom_type[m.pair_geom1[p]], t2 = m.geom_type[m.pair_geom2[p]];
                prow = ((t2 == ILQG_GEOM_CAPSULE && (t1 == ILQG_GEOM_PLANE || t1 == ILQG_GEOM_CAPSULE)) ? 2 : 1) * (m.pair_condim[p] == 3 ? 4 : 1);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) prow += __shfl_xor_sync(0xffffffffu, prow, o);
            rowbound += prow;
            int slot = nrec + __popc(nb_ & ((1u << lane) - 1u));
            if (near && slot < COOP_MAXCAND) cand_out[1 + slot] = p;
            nrec += __popc(nb_);
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
    if (cand_out && lane == 0) {
        cand_out[0] = nrec <= COOP_MAXCAND ? nrec : -1;
        w.hdr[3] = rowbound;
    }
    __syncwarp();
    bool ok = ncon <= COOP_MAXCON && ne <= w.maxefc;
    if (ncon > COOP_MAXCON) ncon = COOP_MAXCON;
    w.ncon = ncon;
    if (ne > w.maxefc) ne = w.maxefc;
    // ---- contact rows: sequential over contacts, lanes over dofs
    for (int c = 0; c < ncon; c++) {
        const int p = w.cpair[c], condim = m.pair_condim[p];
        const int b1 = m.geom_bodyid[m.pair_geom1[p]], b2 = m.geom_bodyid[m.pair_geom2[p]];
        const int nrows = condim == 3 ? 4 : 1;
        if (ne + nrows > w.maxefc) { ok = false; break; }
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
        double tran = m.body_invweight0[b1][0] + m.body_invweight0[b2][0];
        if (tran < ILQG_MINVAL) tran = ILQG_MINVAL;
        const double margin = m.pair_margin[p], dist = w.cdist[c], B = g->pair_B[p];
        if (condim == 1) {
            double R, kt;
            row_params(g->pair_K[p], g->pair_imp[p], m.pair_solimp[p], dist, margin, tran, R, kt);
            if (lane < nv) w.J[ne * nv + lane] = jn;
            if (lane == 0) { w.D[ne] = 1.0 / R; w.rB[ne] = B; w.rkt[ne] = kt; }
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
                if (lane == 0) { w.D[ne + k] = 1.0 / Rpy; w.rB[ne + k] = B; w.rkt[ne + k] = kt; }
            }
            ne += 4;
        }
    }
    if (lane == 0) { w.hdr[0] = ne; w.hdr[1] = ok ? 1 : 0; w.hdr[2] = 0; }
    __syncwarp();
}

// ------------------------------------------------------------------ velocity stage
// In: the C-state's position products and the velocity vector vv (shared memory, nv doubles).
// Out (private): fb = qfrc_passive - qfrc_bias, aref of every row.
DEV void coop_vel(const GModel* __restrict__ g, CoopMem& w, const double* vv, int lane) {
    const ilqg_model& m = g->m;
    const int nv = m.nv, nb = m.nbody, ne = w.hdr[0];
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
                    for (int i = 0; i < 3; i++) { putdot(da + i, {{0, 0, 0}, {0, 0, 0}}); cv = cv + vv[da + i] * cd(da + i); }
                    for (int i = 3; i < 6; i++) putdot(da + i, cross_motion(cv, cd(da + i)));
                    for (int i = 3; i < 6; i++) cv = cv + vv[da + i] * cd(da + i);
                } else if (m.jnt_type[j] == ILQG_JNT_BALL) {   // all three axes turn with the velocity before the joint (mj_comVel)
                    for (int i = 0; i < 3; i++) putdot(da + i, cross_motion(cv, cd(da + i)));
                    for (int i = 0; i < 3; i++) cv = cv + vv[da + i] * cd(da + i);
                } else {
                    putdot(da, cross_motion(cv, cd(da)));
                    cv = cv + vv[da] * cd(da);
                }
            }
            for (int i = m.body_dofadr[b]; i < m.body_dofadr[b] + m.body_dofnum[b]; i++)
                ca = ca + vv[i] * S6{gl3(w.cdofdot + 6 * i), gl3(w.cdofdot + 6 * i + 3)};
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
    if (lane < 6)
        for (int b = nb - 1; b > 0; b--) { int p = m.body_parentid[b]; if (p > 0) w.cfrc[6 * p + lane] += w.cfrc[6 * b + lane]; }
    __syncwarp();
    for (int i = lane; i < nv; i += 32) {
        const int b = m.dof_bodyid[i], j = m.dof_jntid[i];
        double f = 0;
        for (int k = 0; k < 6; k++) f -= w.cdof[6 * i + k] * w.cfrc[6 * b + k];
        f -= m.dof_damping[i] * vv[i];
        if (m.jnt_type[j] != ILQG_JNT_FREE && m.jnt_type[j] != ILQG_JNT_BALL && m.jnt_stiffness[j] != 0) f -= m.jnt_stiffness[j] * w.dspr[i];
        w.fb[i] = f;
    }
    for (int r = lane; r < ne; r += 32) {
        double s = 0;
        const double* Jr = w.J + r * nv;
        for (int i = 0; i < nv; i++) s += Jr[i] * vv[i];
        w.aref[r] = -w.rB[r] * s - w.rkt[r];
    }
    __syncwarp();
}

// actuation, qfrc_smooth, qacc_smooth (in: private fb; uu: ctrl vector in shared memory)
DEV void coop_smooth(const GModel* __restrict__ g, CoopMem& w, const double* uu, int lane, bool solve_as = true) {
    const ilqg_model& m = g->m;
    const int nv = m.nv;
    double f = 0;
    if (lane < nv) {
        f = w.fb[lane];
        const int one = g->dof_act[lane];
        if (one != -1)
            for (int a = one >= 0 ? one : 0; a < (one >= 0 ? one + 1 : m.nu); a++)
                if (m.act_dofid[a] == lane) {
                    double c = uu[a];
                    if (m.act_ctrllimited[a]) c = clampd(c, m.act_ctrlrange[a][0], m.act_ctrlrange[a][1]);
                    f += m.act_gear[a] * c;
                }
        w.fs[lane] = f;
    }
    if (solve_as) {
        const double a = coop_chol_solve(w.L, f, nv, lane);
        if (lane < nv) w.as[lane] = a;
    }
    __syncwarp();
}

// ------------------------------------------------------------------ constraint solve (same algorithm as dyn.cuh::solve)
struct Mask128 { unsigned w[4]; };
DEV bool operator==(const Mask128& a, const Mask128& b) { return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2] && a.w[3] == b.w[3]; }

// Newton Hessian M + sum_active D_r J_r J_r' (packed) into `H`, then its Cholesky factor in place.
// The rank-na update is the one dense contraction of this path that ncu shows as dense enough for the tensor cores (13 % of
// the instructions of a full humanoid evaluation when done entry by entry, profiles/r01_humanoid_coop_kernels.md): it runs as
// fp64 DMMA (mma.sync m8n8k4) over 8 x 8 tiles of the lower triangle.  Per k-step of four active rows the warp loads the rows'
// D and four 8-wide slices of J once and issues one DMMA per tile: A = D_r J_r[i-slice] (8 x 4), B = J_r[j-slice] (4 x 8).
// Fragment layout (PTX ISA, m8n8k4 .f64): a: row = lane / 4, k = lane % 4;  b: k = lane % 4, col = lane / 4;
// c/d: row = lane / 4, cols = 2 (lane % 4) + {0, 1}.
DEV void coop_hessian_factor(const CoopMem& w, double* H, int nv, int na, int lane) {
    constexpr int NB8 = COOP_NMAX / 8;           // 8-wide blocks per side
    const int nblk = (nv + 7) >> 3;
    const int g = lane >> 2, t = lane & 3;       // fragment coordinates
    double acc[NB8 * (NB8 + 1) / 2][2];
#pragma unroll
    for (int ib = 0; ib < NB8; ib++)
#pragma unroll
        for (int jb = 0; jb <= ib; jb++) {
            const int tile = ib * (ib + 1) / 2 + jb, i = 8 * ib + g, j = 8 * jb + 2 * t;
            acc[tile][0] = (ib < nblk && i < nv && j <= i) ? w.M[ptri(i, j)] : 0.0;
            acc[tile][1] = (ib < nblk && i < nv && j + 1 <= i) ? w.M[ptri(i, j + 1)] : 0.0;
        }
    for (int a0 = 0; a0 < na; a0 += 4) {
        const bool in = a0 + t < na;
        const int r = in ? w.alist[a0 + t] : 0;
        const double D = in ? w.D[r] : 0.0;
        const double* Jr = w.J + r * nv;
        double Jv[NB8];
#pragma unroll
        for (int bb = 0; bb < NB8; bb++) Jv[bb] = (in && 8 * bb + g < nv) ? Jr[8 * bb + g] : 0.0;
#pragma unroll
        for (int ib = 0; ib < NB8; ib++) {
            if (ib < nblk) {   // uniform
                const double av = D * Jv[ib];
#pragma unroll
                for (int jb = 0; jb <= ib; jb++) dmma_8x8x4(acc[ib * (ib + 1) / 2 + jb][0], acc[ib * (ib + 1) / 2 + jb][1], av, Jv[jb]);
            }
        }
    }
#pragma unroll
    for (int ib = 0; ib < NB8; ib++)
#pragma unroll
        for (int jb = 0; jb <= ib; jb++) {
            const int tile = ib * (ib + 1) / 2 + jb, i = 8 * ib + g, j = 8 * jb + 2 * t;
            if (ib < nblk && i < nv) {
                if (j <= i) H[ptri(i, j)] = acc[tile][0];
                if (j + 1 <= i) H[ptri(i, j + 1)] = acc[tile][1];
            }
        }
    __syncwarp();
    coop_chol(H, nv, lane);
}

// in: private fs, as, aref, warm; out: private qacc (and warm, the next warm start); nv <= 32: lane i owns dof i
// reuse: take the C-state's cached factor Hc whenever the active set equals the one it was built for (the qvel / ctrl columns
// of a knot share M, J and D with the centre, so all of their Newton systems with that active set are the same matrix)
// returns the Newton iterations run; *exact (optional) = the solve left through the exact-optimum test or had no rows
// warm_only: the solve of a PERTURBED evaluation that starts from the centre's solution (mjderivative.cpp:75,91).  MuJoCo starts from
// the better of the warm start and qacc_smooth; within +-eps of the centre that is the warm start whenever a row is active, and when none
// is the first Newton step (Hessian = M) lands on qacc_smooth exactly — so neither qacc_smooth nor the Cholesky factor of M is needed:
// one 27 x 27 factorisation per evaluation instead of two.  The Gauss term of the cost is taken as a'Ma/2 - a'fs (same differences).
DEV int coop_solve(const GModel* __restrict__ g, CoopMem& w, int maxiter, double tol, int lane, bool need_forces = false, bool reuse = false,
                   bool* exact = nullptr, bool warm_only = false) {
    const ilqg_model& m = g->m;
    const int nv = m.nv, ne = w.hdr[0];
    static_assert(COOP_MAXEFC <= 128, "mask width");
    if (ne == 0) {
        if (warm_only) {   // qacc = M^-1 qfrc_smooth through the (one) factorisation of this evaluation
            coop_hessian_factor(w, w.H, nv, 0, lane);
            const double a = coop_chol_solve(w.H, lane < nv ? w.fs[lane] : 0.0, nv, lane);
            if (lane < nv) { w.qacc[lane] = a; w.warm[lane] = a; w.fc[lane] = 0; }
        } else if (lane < nv) { w.qacc[lane] = w.as[lane]; w.warm[lane] = w.as[lane]; w.fc[lane] = 0; }
        __syncwarp();
        if (exact) *exact = true;
        return 0;
    }
    if (exact) *exact = false;
    const bool dof = lane < nv;
    const double fs_i = dof ? w.fs[lane] : 0.0, as_i = (dof && !warm_only) ? w.as[lane] : 0.0, warm_i = dof ? w.warm[lane] : 0.0;
    double qacc_i, Ma_i;
    if (warm_only) {
        for (int r = lane; r < ne; r += 32) {
            double jw = -w.aref[r];
            const double* Jr = w.J + r * nv;
            for (int i = 0; i < nv; i++) jw += Jr[i] * w.warm[i];
            w.jar[r] = jw;
        }
        qacc_i = warm_i;
        Ma_i = coop_symv(w.M, w.warm, nv, lane);
    } else {
        // the better of the warm start and qacc_smooth; one pass over the rows evaluates both
        double cw = 0, cs = 0;
        for (int r = lane; r < ne; r += 32) {
            double jw = -w.aref[r], js = jw;
            const double* Jr = w.J + r * nv;
            for (int i = 0; i < nv; i++) { jw += Jr[i] * w.warm[i]; js += Jr[i] * w.as[i]; }
            const double D = w.D[r];
            if (jw < 0) cw += 0.5 * D * jw * jw;
            if (js < 0) cs += 0.5 * D * js * js;
            w.jar[r] = jw;
            w.jv[r] = js;
        }
        const double Mw = coop_symv(w.M, w.warm, nv, lane);
        if (dof) cw += 0.5 * (Mw - fs_i) * (warm_i - as_i);
        cw = warp_sum(cw);
        cs = warp_sum(cs);
        if (cw < cs) { qacc_i = warm_i; Ma_i = Mw; }
        else {
            qacc_i = as_i;
            Ma_i = coop_symv(w.M, w.as, nv, lane);
            __syncwarp();
            for (int r = lane; r < ne; r += 32) w.jar[r] = w.jv[r];
        }
    }
    __syncwarp();
    const double scale = 1.0 / (m.meaninertia * (nv > 1 ? nv : 1));
    double cost = 0, old = 0, fc_i = 0;
    int iter = 0;
    for (;;) {
        // ---- active set (compacted list), cost, forces, gradient, Hessian, Newton direction
        Mask128 act = {{0, 0, 0, 0}};
        double c = 0;
        int na = 0;
        for (int r0 = 0; r0 < ne; r0 += 32) {
            const int r = r0 + lane;
            const bool a = r < ne && w.jar[r] < 0;
            const unsigned bal = __ballot_sync(0xffffffffu, a);
            act.w[r0 >> 5] = bal;
            if (a) {
                c += 0.5 * w.D[r] * w.jar[r] * w.jar[r];
                w.alist[na + __popc(bal & ((1u << lane) - 1u))] = r;
            }
            na += __popc(bal);
        }
        __syncwarp();
        double f = 0;
        if (dof)
            for (int a = 0; a < na; a++) { const int r = w.alist[a]; f += w.J[r * nv + lane] * (-w.D[r] * w.jar[r]); }
        fc_i = f;
        const double grad_i = Ma_i - fs_i - f;
        if (dof) c += warm_only ? (0.5 * Ma_i - fs_i) * qacc_i : 0.5 * (Ma_i - fs_i) * (qacc_i - as_i);
        cost = warp_sum(c);
        const double* Hf = w.H;
        if (reuse && w.hdr[2] && act.w[0] == (unsigned)w.hdr[4] && act.w[1] == (unsigned)w.hdr[5] && act.w[2] == (unsigned)w.hdr[6] &&
            act.w[3] == (unsigned)w.hdr[7])
            Hf = w.Hc;
        else
            coop_hessian_factor(w, w.H, nv, na, lane);
        const double search_i = -coop_chol_solve(Hf, grad_i, nv, lane);
        if (iter > 0) {
            const double gn = warp_sum(dof ? grad_i * grad_i : 0.0);
            if (scale * (old - cost) < tol || scale * sqrt(gn) < tol) break;
        }
        if (iter >= maxiter) break;
        // ---- exact line search
        if (dof) w.search[lane] = search_i;
        __syncwarp();
        const double Mv_i = coop_symv(w.M, w.search, nv, lane);
        double g1 = dof ? search_i * (Ma_i - fs_i) : 0.0, g2 = dof ? search_i * Mv_i : 0.0;
        double d1 = 0, d2 = 0;
        for (int r = lane; r < ne; r += 32) {
            double s = 0;
            const double* Jr = w.J + r * nv;
            for (int i = 0; i < nv; i++) s += Jr[i] * w.search[i];
            w.jv[r] = s;
            if (w.jar[r] < 0) { double t = w.D[r] * s; d1 += t * w.jar[r]; d2 += t * s; }
        }
        g1 = warp_sum(g1); g2 = warp_sum(g2);
        d1 = warp_sum(d1) + g1; d2 = warp_sum(d2) + g2;
        if (d1 >= 0 || d2 < ILQG_MINVAL) break;
        // rows this lane owns stay in registers for the whole search (ne <= 128: at most 4 per lane)
        double rj[4], rv[4], rD[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int r = lane + 32 * t;
            const bool in = r < ne;
            rj[t] = in ? w.jar[r] : 0.0; rv[t] = in ? w.jv[r] : 0.0; rD[t] = in ? w.D[r] : 0.0;
        }
        double alpha = 0, lo = 0, hi = CUDART_INF;
        Mask128 cur = act;
        for (int it = 0; it < m.ls_iterations; it++) {
            if (d1 < 0) lo = alpha; else hi = alpha;
            double an = alpha - d1 / d2;
            if (!(an > lo && an < hi)) an = isinf(hi) ? 2 * alpha + 1 : 0.5 * (lo + hi);
            double e1 = 0, e2 = 0;
            Mask128 mk = {{0, 0, 0, 0}};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (32 * t < ne) {   // uniform
                    bool a = false;
                    if (lane + 32 * t < ne) {
                        const double x = rj[t] + an * rv[t];
                        if (x < 0) { const double tt = rD[t] * rv[t]; e1 += tt * x; e2 += tt * rv[t]; a = true; }
                    }
                    mk.w[t] = __ballot_sync(0xffffffffu, a);
                }
            }
            e1 = warp_sum(e1) + g1 + g2 * an;
            e2 = warp_sum(e2) + g2;
            const bool same = mk == cur;
            alpha = an; d1 = e1; d2 = e2; cur = mk;
            if (same || d1 == 0 || d2 < ILQG_MINVAL) break;
        }
        if (alpha == 0) break;
        qacc_i += alpha * search_i;
        Ma_i += alpha * Mv_i;
        for (int r = lane; r < ne; r += 32) w.jar[r] += alpha * w.jv[r];
        __syncwarp();
        old = cost;
        iter++;
        if (cur == act) {  // exact optimum: full Newton step inside the piece the Hessian was built for
            if (exact) *exact = true;
            break;
        }
    }
    if (need_forces) {  // qfrc_constraint at the final point (mj_Euler needs it)
        double f = 0;
        if (dof)
            for (int r = 0; r < ne; r++)
                if (w.jar[r] < 0) f += w.J[r * nv + lane] * (-w.D[r] * w.jar[r]);
        fc_i = f;
    }
    if (dof) { w.qacc[lane] = qacc_i; w.warm[lane] = qacc_i; w.fc[lane] = fc_i; }
    __syncwarp();
    return iter;
}

// step cost on the device, lane 0 semantics (same term order as ilqg.cu's cost_eval)
DEV double coop_cost_eval(const ilqg_cost* cost, const double* q, const double* v, const double* u, int nq, int nv, int nu) {
    double c = 0;
    for (int i = 0; i < nq; i++) { c = __dadd_rn(c, __dmul_rn(__dmul_rn(cost->q2[i], q[i]), q[i])); c = __dadd_rn(c, __dmul_rn(cost->q1[i], q[i])); }
    for (int i = 0; i < nv; i++) { c = __dadd_rn(c, __dmul_rn(__dmul_rn(cost->v2[i], v[i]), v[i])); c = __dadd_rn(c, __dmul_rn(cost->v1[i], v[i])); }
    for (int i = 0; i < nu; i++) { c = __dadd_rn(c, __dmul_rn(__dmul_rn(cost->u2[i], u[i]), u[i])); c = __dadd_rn(c, __dmul_rn(cost->u1[i], u[i])); }
    return c;
}

// ------------------------------------------------------------------ kernels
// centre: one warp per knot -> qacc_center, the knot's C-state and candidate pair list in HBM
__global__ void __launch_bounds__(32) coop_center_kernel(const GModel* __restrict__ g, int nknots, const double* __restrict__ qpos,
                                                         const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                         const double* __restrict__ warmstart, int niter, int nwarmup, double slack,
                                                         int cdbl, int pdbl, double* __restrict__ qacc_center, int* __restrict__ status,
                                                         double* __restrict__ cstate_out, int* __restrict__ cand_out,
                                                         int* __restrict__ rowbound_out, int* __restrict__ diag) {
    extern __shared__ __align__(16) double coop_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * (blockDim.x >> 5) + wib;
    if (k >= nknots) return;
    const ilqg_model& m = g->m;
    CoopMem w;
    double* base = coop_smem + (size_t)wib * (cdbl + pdbl);
    coop_carve_cstate(w, base, m);
    coop_carve_priv(w, base + cdbl, m);
    for (int i = lane; i < m.nq; i += 32) w.q[i] = qpos[(size_t)k * m.nq + i];
    for (int i = lane; i < m.nv; i += 32) w.v[i] = qvel[(size_t)k * m.nv + i];
    for (int i = lane; i < m.nu; i += 32) w.u[i] = ctrl[(size_t)k * m.nu + i];
    __syncwarp();
    const long long t0 = clock64();
    coop_pos(g, w, lane, nullptr, -1, cand_out ? cand_out + (size_t)k * (COOP_MAXCAND + 1) : nullptr, slack);
    const bool ok = w.hdr[1] != 0;
    coop_vel(g, w, w.v, lane);
    coop_smooth(g, w, w.u, lane);
    if (lane < m.nv) w.warm[lane] = warmstart ? warmstart[(size_t)k * m.nv + lane] : 0.0;
    __syncwarp();
    const long long t1 = clock64();
    int it_first = 0, it_all = 0;
    for (int rep = 0; rep < nwarmup; rep++) {   // repetitions stop at the first solve that is exact (see fd_center_kernel)
        bool exact = false;
        const int it = coop_solve(g, w, niter, 0.0, lane, false, false, &exact);
        if (rep == 0) it_first = it;
        it_all += it;
        if (exact) break;
    }
    if (diag && lane == 0) {   // ILQG_DIAG_* (include/ilqg_b200.h)
        const long long t2 = clock64();
        int na = 0;
        for (int r = 0; r < w.hdr[0]; r++) na += w.jar[r] < 0;
        int* d = diag + (size_t)k * ILQG_DIAG_INTS;
        d[0] = w.hdr[0]; d[1] = it_first; d[2] = it_all; d[3] = na; d[4] = (int)(t1 - t0); d[5] = (int)(t2 - t1); d[6] = w.ncon; d[7] = 0;
    }
    bool fin = true;
    for (int i = lane; i < m.nv; i += 32) { qacc_center[(size_t)k * m.nv + i] = w.qacc[i]; w.center[i] = w.qacc[i]; w.fb0[i] = w.fb[i]; fin = fin && isfinite(w.qacc[i]); }
    for (int r = lane; r < w.hdr[0]; r += 32) w.aref0[r] = w.aref[r];
    {   // the Newton factor at the centre solution: every qvel / ctrl column whose first iterate has this active set reuses it
        const int ne = w.hdr[0];
        int na = 0;
        unsigned mk[4] = {0, 0, 0, 0};
        for (int r0 = 0; r0 < ne; r0 += 32) {
            const int r = r0 + lane;
            const bool a = r < ne && w.jar[r] < 0;
            const unsigned bal = __ballot_sync(0xffffffffu, a);
            mk[r0 >> 5] = bal;
            if (a) w.alist[na + __popc(bal & ((1u << lane) - 1u))] = r;
            na += __popc(bal);
        }
        __syncwarp();
        if (ne > 0) coop_hessian_factor(w, w.Hc, m.nv, na, lane);
        if (lane == 0) { w.hdr[2] = ne > 0 ? 1 : 0; w.hdr[4] = (int)mk[0]; w.hdr[5] = (int)mk[1]; w.hdr[6] = (int)mk[2]; w.hdr[7] = (int)mk[3]; }
    }
    fin = __all_sync(0xffffffffu, fin);
    if (status && lane == 0) status[k] = !ok ? ILQG_ERR_CAPACITY : (fin ? 0 : ILQG_ERR_NONFINITE);
    if (rowbound_out && lane == 0)   // a knot whose candidate list overflowed tests every pair when perturbed: full capacity
        rowbound_out[k] = (cand_out && cand_out[(size_t)k * (COOP_MAXCAND + 1)] < 0) ? COOP_MAXEFC : w.hdr[3];
    __syncwarp();
    if (cstate_out) {
        double* dst = cstate_out + (size_t)k * cdbl;
        for (int e = lane; e < cdbl; e += 32) dst[e] = base[e];
    }
}

DEV size_t coop_deriv_off(int col, int j, int nv, int nu) {   // reference layout (differentiator.h:56-61)
    if (col < nu) return 2 * (size_t)nv * nv + col + (size_t)j * nu;
    if (col < nu + nv) return (size_t)nv * nv + (col - nu) + (size_t)j * nv;
    return (col - nu - nv) + (size_t)j * nv;
}
DEV size_t coop_grad_off(int col, int nv, int nu) {
    size_t off = (size_t)nv * (2 * nv + nu);
    if (col < nu) return off + 2 * nv + col;
    if (col < nu + nv) return off + nv + (col - nu);
    return off + col - nu - nv;
}

// qvel / ctrl columns: one CTA per knot; the warps share the centre's C-state and take columns in turn
__global__ void __launch_bounds__(256, 2) coop_velctrl_kernel(const GModel* __restrict__ g, int nknots, const double* __restrict__ cstate,
                                                              const ilqg_cost* __restrict__ cost, double eps, int niter, int cdbl, int pdbl,
                                                              const FdDst dst, int* __restrict__ status) {
    extern __shared__ __align__(16) double coop_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int k = blockIdx.x;
    const ilqg_model& m = g->m;
    const int nq = m.nq, nv = m.nv, nu = m.nu, ncol = 2 * nv + nu, nd = nv * ncol + ncol;
    {
        const double* src = cstate + (size_t)k * cdbl;
        for (int e = threadIdx.x; e < cdbl; e += blockDim.x) coop_smem[e] = src[e];
    }
    __syncthreads();
    CoopMem w;
    coop_carve_cstate(w, coop_smem, m);
    coop_carve_priv(w, coop_smem + cdbl + (size_t)wib * pdbl, m);
    const int ne = w.hdr[0];
    double c0 = 0;
    if (cost && lane == 0) c0 = coop_cost_eval(cost, w.q, w.v, w.u, nq, nv, nu);
    bool fin = true;
    for (int col = wib; col < nu + nv; col += nwarp) {
        const bool is_vel = col >= nu;
        double plus = 0, dcost = 0;
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
            const double se = sgn * eps;
            if (lane < nv) { w.pv[lane] = w.v[lane] + ((is_vel && col - nu == lane) ? se : 0.0); w.warm[lane] = w.center[lane]; }
            if (lane < nu) w.pu[lane] = w.u[lane] + ((!is_vel && col == lane) ? se : 0.0);
            __syncwarp();
            if (cost && sgn > 0 && lane == 0) dcost = __ddiv_rn(__dsub_rn(coop_cost_eval(cost, w.q, w.pv, w.pu, nq, nv, nu), c0), eps);
            if (is_vel) coop_vel(g, w, w.pv, lane);
            else {   // mjSTAGE_VEL skip: the centre's velocity-stage products
                if (lane < nv) w.fb[lane] = w.fb0[lane];
                for (int r = lane; r < ne; r += 32) w.aref[r] = w.aref0[r];
                __syncwarp();
            }
            coop_smooth(g, w, w.pu, lane);
            coop_solve(g, w, niter, 0.0, lane, false, true);
            const double a = lane < nv ? w.qacc[lane] : 0.0;
            if (sgn > 0) plus = a;
            else if (lane < nv) {
                const double d = (plus - a) / (2 * eps);
                fin = fin && isfinite(d);
                for (int t = 0; t < dst.n; t++) dst.p[t][(size_t)k * nd + coop_deriv_off(col, lane, nv, nu)] = d;
            }
            __syncwarp();
        }
        if (cost && lane == 0)
            for (int t = 0; t < dst.n; t++) dst.p[t][(size_t)k * nd + coop_grad_off(col, nv, nu)] = dcost;
    }
    fin = __all_sync(0xffffffffu, fin);
    if (status && lane == 0 && !fin) atomicCAS(&status[k], 0, ILQG_ERR_NONFINITE);
}

// qpos columns: one warp per (knot, column); +eps then -eps through the full pipeline (narrow phase on the candidates)
__global__ void __launch_bounds__(32) coop_qpos_kernel(const GModel* __restrict__ g, int nknots, const double* __restrict__ qpos,
                                                       const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                       const double* __restrict__ qacc_center, const int* __restrict__ cand,
                                                       const ilqg_cost* __restrict__ cost, double eps, int niter, int cdbl, int pdbl,
                                                       const int* __restrict__ rowbound, int cap_lo, int cap,
                                                       const FdDst dst, int* __restrict__ status) {
    // Launched once per row-capacity class (cap_lo, cap]: a warp whose knot's row bound falls outside leaves at once.  cdbl / pdbl
    // are the block sizes for capacity `cap` — the smaller class fits 8 rollouts per SM instead of 6.
    extern __shared__ __align__(16) double coop_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const ilqg_model& m = g->m;
    const int nq = m.nq, nv = m.nv, nu = m.nu, ncol = 2 * nv + nu, nd = nv * ncol + ncol;
    const long item = (long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (item >= (long)nknots * nv) return;
    const int k = (int)(item / nv), i = (int)(item - (long)k * nv), col = nu + nv + i;
    if (rowbound) {
        const int rb = rowbound[k];
        if (rb <= cap_lo || (rb > cap && cap < COOP_MAXEFC)) return;   // (bounds above the largest class run there and report capacity)
    }
    CoopMem w;
    double* base = coop_smem + (size_t)wib * (cdbl + pdbl);
    coop_carve_cstate(w, base, m, cap, false);   // no factor of M in this kernel (coop_solve's warm_only)
    coop_carve_priv(w, base + cdbl, m, cap);
    const int* kc = cand ? cand + (size_t)k * (COOP_MAXCAND + 1) : nullptr;
    const int ncand = kc ? kc[0] : -1;
    double plus = 0, dcost = 0;
    bool ok = true, fin = true;
    for (int sgn = 1; sgn >= -1; sgn -= 2) {
        for (int e = lane; e < nq; e += 32) w.q[e] = qpos[(size_t)k * nq + e];
        for (int e = lane; e < nv; e += 32) w.v[e] = qvel[(size_t)k * nv + e];
        for (int e = lane; e < nu; e += 32) w.u[e] = ctrl[(size_t)k * nu + e];
        __syncwarp();
        if (lane == 0) {
            const double se = sgn * eps;
            double c0 = 0;
            if (cost && sgn > 0) c0 = coop_cost_eval(cost, w.q, w.v, w.u, nq, nv, nu);
            const int j = m.dof_jntid[i];
            if (m.jnt_type[j] == ILQG_JNT_BALL) {   // mjderivative.cpp:152-156
                const int a = i - m.jnt_dofadr[j];
                quat_integrate(&w.q[m.jnt_qposadr[j]], V3{a == 0 ? se : 0.0, a == 1 ? se : 0.0, a == 2 ? se : 0.0}, 1.0);
            } else if (m.jnt_type[j] == ILQG_JNT_FREE && i >= m.jnt_dofadr[j] + 3) {
                const int a = i - m.jnt_dofadr[j] - 3;
                quat_integrate(&w.q[m.jnt_qposadr[j] + 3], V3{a == 0 ? se : 0.0, a == 1 ? se : 0.0, a == 2 ? se : 0.0}, 1.0);
            } else
                w.q[m.jnt_qposadr[j] + i - m.jnt_dofadr[j]] += se;
            if (cost && sgn > 0) dcost = __ddiv_rn(__dsub_rn(coop_cost_eval(cost, w.q, w.v, w.u, nq, nv, nu), c0), eps);
        }
        __syncwarp();
        coop_pos(g, w, lane, kc ? kc + 1 : nullptr, ncand, nullptr, 0.0, false);
        ok = ok && w.hdr[1] != 0;
        coop_vel(g, w, w.v, lane);
        coop_smooth(g, w, w.u, lane, false);
        if (lane < nv) w.warm[lane] = qacc_center[(size_t)k * nv + lane];
        __syncwarp();
        coop_solve(g, w, niter, 0.0, lane, false, false, nullptr, true);
        const double a = lane < nv ? w.qacc[lane] : 0.0;
        if (sgn > 0) plus = a;
        else if (lane < nv) {
            const double d = (plus - a) / (2 * eps);
            fin = isfinite(d);
            for (int t = 0; t < dst.n; t++) dst.p[t][(size_t)k * nd + coop_deriv_off(col, lane, nv, nu)] = d;
        }
        __syncwarp();
    }
    if (cost && lane == 0)
        for (int t = 0; t < dst.n; t++) dst.p[t][(size_t)k * nd + coop_grad_off(col, nv, nu)] = dcost;
    fin = __all_sync(0xffffffffu, fin);
    if (status && lane == 0) {
        if (!ok) atomicExch(&status[k], ILQG_ERR_CAPACITY);
        else if (!fin) atomicCAS(&status[k], 0, ILQG_ERR_NONFINITE);
    }
}

// mj_forward for n states: one warp per state
__global__ void __launch_bounds__(32) coop_forward_kernel(const GModel* __restrict__ g, int n, const double* __restrict__ qpos,
                                                          const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                          double* __restrict__ warmstart, double* __restrict__ qacc_out, int cdbl, int pdbl) {
    extern __shared__ __align__(16) double coop_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * (blockDim.x >> 5) + wib;
    if (k >= n) return;
    const ilqg_model& m = g->m;
    CoopMem w;
    double* base = coop_smem + (size_t)wib * (cdbl + pdbl);
    coop_carve_cstate(w, base, m);
    coop_carve_priv(w, base + cdbl, m);
    for (int i = lane; i < m.nq; i += 32) w.q[i] = qpos[(size_t)k * m.nq + i];
    for (int i = lane; i < m.nv; i += 32) w.v[i] = qvel[(size_t)k * m.nv + i];
    for (int i = lane; i < m.nu; i += 32) w.u[i] = ctrl[(size_t)k * m.nu + i];
    __syncwarp();
    coop_pos(g, w, lane, nullptr, -1, nullptr, 0.0);
    coop_vel(g, w, w.v, lane);
    coop_smooth(g, w, w.u, lane);
    if (lane < m.nv) w.warm[lane] = warmstart ? warmstart[(size_t)k * m.nv + lane] : 0.0;
    __syncwarp();
    coop_solve(g, w, m.iterations, m.tolerance, lane);
    for (int i = lane; i < m.nv; i += 32) {
        qacc_out[(size_t)k * m.nv + i] = w.qacc[i];
        if (warmstart) warmstart[(size_t)k * m.nv + i] = w.warm[i];
    }
}

// one mj_step (Euler with implicit joint damping; RK4 models use the thread-per-rollout path) of the state in w.q / w.v under w.u;
// warm_i: this lane's component of qacc_warmstart (kept in a register: the position stage's temporaries alias w.warm)
DEV void coop_step_once(const GModel* __restrict__ g, CoopMem& w, int lane, double& warm_i) {
    const ilqg_model& m = g->m;
    const int nv = m.nv, nt = coop_nt(m.nv);
    const double h = m.timestep;
    coop_pos(g, w, lane, nullptr, -1, nullptr, 0.0);
    coop_vel(g, w, w.v, lane);
    coop_smooth(g, w, w.u, lane);
    if (lane < nv) w.warm[lane] = warm_i;
    __syncwarp();
    coop_solve(g, w, m.iterations, m.tolerance, lane, true);
    if (lane < nv) warm_i = w.warm[lane];
    // mj_Euler: (M + h diag(b)) a = qfrc_smooth + qfrc_constraint when any dof is damped
    double a_i = lane < nv ? w.qacc[lane] : 0.0;
    if (g->any_damping) {
        for (int e = lane; e < nt; e += 32) w.H[e] = w.M[e];
        __syncwarp();
        if (lane < nv) w.H[ptri(lane, lane)] += h * m.dof_damping[lane];
        __syncwarp();
        coop_chol(w.H, nv, lane);
        a_i = coop_chol_solve(w.H, lane < nv ? w.fs[lane] + w.fc[lane] : 0.0, nv, lane);
    }
    if (lane < nv) w.v[lane] += h * a_i;
    __syncwarp();
    for (int j = lane; j < m.njnt; j += 32) {
        const int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
        if (m.jnt_type[j] == ILQG_JNT_FREE) {
            for (int c = 0; c < 3; c++) w.q[qa + c] += h * w.v[da + c];
            quat_integrate(&w.q[qa + 3], V3{w.v[da + 3], w.v[da + 4], w.v[da + 5]}, h);
        } else if (m.jnt_type[j] == ILQG_JNT_BALL)
            quat_integrate(&w.q[qa], V3{w.v[da], w.v[da + 1], w.v[da + 2]}, h);
        else
            w.q[qa] += h * w.v[da];
    }
    __syncwarp();
}

// nsteps x mj_step for n states
__global__ void __launch_bounds__(32) coop_step_kernel(const GModel* __restrict__ g, int n, int nsteps, double* __restrict__ qpos,
                                                       double* __restrict__ qvel, const double* __restrict__ ctrl, double* __restrict__ warmstart,
                                                       double* __restrict__ qacc_out, int cdbl, int pdbl) {
    extern __shared__ __align__(16) double coop_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * (blockDim.x >> 5) + wib;
    if (k >= n) return;
    const ilqg_model& m = g->m;
    const int nv = m.nv;
    CoopMem w;
    double* base = coop_smem + (size_t)wib * (cdbl + pdbl);
    coop_carve_cstate(w, base, m);
    coop_carve_priv(w, base + cdbl, m);
    for (int i = lane; i < m.nq; i += 32) w.q[i] = qpos[(size_t)k * m.nq + i];
    for (int i = lane; i < nv; i += 32) w.v[i] = qvel[(size_t)k * nv + i];
    for (int i = lane; i < m.nu; i += 32) w.u[i] = ctrl[(size_t)k * m.nu + i];
    double warm_i = (warmstart && lane < nv) ? warmstart[(size_t)k * nv + lane] : 0.0;
    __syncwarp();
    for (int s = 0; s < nsteps; s++) coop_step_once(g, w, lane, warm_i);
    for (int i = lane; i < m.nq; i += 32) qpos[(size_t)k * m.nq + i] = w.q[i];
    for (int i = lane; i < nv; i += 32) {
        qvel[(size_t)k * nv + i] = w.v[i];
        if (warmstart) warmstart[(size_t)k * nv + i] = warm_i;
        if (qacc_out) qacc_out[(size_t)k * nv + i] = w.qacc[i];
    }
}

// ------------------------------------------------------------------ iLQR on the warp-cooperative engine (nq != nv allowed)
// x_a (-) x_b in the tangent space, the state vector of the opt-in extension beyond quirk Q9 (the reference's "2 nv doubles at
// qpos" is undefined with a quaternion in qpos): slide / hinge dofs by plain difference, a free joint's orientation as the
// body-frame rotation vector w with q_b * quat(w) = q_a (mju_subQuat) — the coordinates the FD blocks already use
// (mjderivative.cpp:152-169).  Lanes stride over joints; out[0..nv) position part, out[nv..2nv) velocity part.
DEV void coop_state_diff(const ilqg_model& m, const double* qa, const double* va, const double* qb, const double* vb, double* out, int lane) {
    const int nv = m.nv;
    for (int j = lane; j < m.njnt; j += 32) {
        int qadr = m.jnt_qposadr[j], dadr = m.jnt_dofadr[j];
        if (m.jnt_type[j] == ILQG_JNT_FREE || m.jnt_type[j] == ILQG_JNT_BALL) {
            const int fr = m.jnt_type[j] == ILQG_JNT_FREE ? 3 : 0;   // a free joint carries a position in front of its quaternion
            for (int k = 0; k < fr; k++) out[dadr + k] = qa[qadr + k] - qb[qadr + k];
            qadr += fr; dadr += fr;
            const Q4 A = qnormalized({qa[qadr], qa[qadr + 1], qa[qadr + 2], qa[qadr + 3]});
            const Q4 Bq = qnormalized({qb[qadr], qb[qadr + 1], qb[qadr + 2], qb[qadr + 3]});
            const Q4 d = qmul({Bq.w, -Bq.x, -Bq.y, -Bq.z}, A);
            const double sn = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
            double ang = 2 * atan2(sn, d.w);
            if (ang > 3.14159265358979323846) ang -= 2 * 3.14159265358979323846;
            const double sc = sn < 1e-15 ? 0.0 : ang / sn;
            out[dadr] = d.x * sc; out[dadr + 1] = d.y * sc; out[dadr + 2] = d.z * sc;
        } else
            out[dadr] = qa[qadr] - qb[qadr];
    }
    for (int i = lane; i < nv; i += 32) out[nv + i] = va[i] - vb[i];
}

// ILQR::forwardPass (ilqr.h:116-130) for every (instance, line-search step size): one warp each, same bookkeeping as
// ilqr_rollout_kernel (knot snapshots into the alpha's candidate, cost of the stored knots, u = K (x (-) x*) + alpha k + u*)
__global__ void __launch_bounds__(32) coop_rollout_kernel(const GModel* __restrict__ g, IlqrBuffers b, const ilqg_cost* __restrict__ cost, int cdbl,
                                                          int pdbl) {
    extern __shared__ __align__(16) double coop_smem[];
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x;
    const int ninst = b.ninst;
    if (item >= ninst * b.nalpha) return;
    const int a = item / ninst, i = item - a * ninst;
    const ilqg_model& m = g->m;
    const int nq = m.nq, nv = m.nv, nu = m.nu, nx = 2 * nv;
    CoopMem w;
    coop_carve_cstate(w, coop_smem, m);
    coop_carve_priv(w, coop_smem + cdbl, m);
    const double alpha = b.alphas[a];
    for (int e = lane; e < nq; e += 32) w.q[e] = b.init_q[(size_t)i * nq + e];
    for (int e = lane; e < nv; e += 32) w.v[e] = b.init_v[(size_t)i * nv + e];
    double warm_i = lane < nv ? b.init_w[(size_t)i * nv + lane] : 0.0;
    double J = 0;
    const size_t T1 = (size_t)(b.N + 1) * ninst;
    double* dx = w.grad;   // 2 nv doubles (grad | search are adjacent); consumed before the position stage reuses the private block
    __syncwarp();
    for (int n = b.N; n >= 0; n--) {
        const size_t kn = (size_t)n * ninst + i;
        coop_state_diff(m, w.q, w.v, b.nom_q + kn * nq, b.nom_v + kn * nv, dx, lane);
        __syncwarp();
        if (lane < nu) {
            const double* K = b.K + kn * nu * nx;
            double s = 0;
            for (int c = 0; c < nx; c++) s += K[lane + c * nu] * dx[c];
            w.u[lane] = s + alpha * b.k[kn * nu + lane] + b.nom_u[kn * nu + lane];
        }
        __syncwarp();
        const size_t cn = (size_t)a * T1 + kn;
        for (int e = lane; e < nq; e += 32) b.cand_q[cn * nq + e] = w.q[e];
        for (int e = lane; e < nv; e += 32) { b.cand_v[cn * nv + e] = w.v[e]; }
        if (lane < nv) b.cand_w[cn * nv + lane] = warm_i;
        for (int e = lane; e < nu; e += 32) b.cand_u[cn * nu + e] = w.u[e];
        if (cost && lane == 0) J = __dadd_rn(J, coop_cost_eval(cost, w.q, w.v, w.u, nq, nv, nu));
        coop_step_once(g, w, lane, warm_i);
    }
    if (lane == 0) b.cand_J[(size_t)a * ninst + i] = J;
}

// c_n = x*_{n-1} (-) x*_n for every knot n >= 1 of every instance (the affine term of ilqr.h:161-163 in tangent coordinates)
__global__ void coop_cdiff_kernel(const GModel* __restrict__ g, IlqrBuffers b) {
    const int lane = threadIdx.x & 31;
    const size_t item = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int ninst = b.ninst;
    if (item >= (size_t)b.N * ninst) return;
    const int n = 1 + (int)(item / ninst), i = (int)(item % ninst);
    const ilqg_model& m = g->m;
    const size_t kn = (size_t)n * ninst + i, kp = (size_t)(n - 1) * ninst + i;
    coop_state_diff(m, b.nom_q + kp * m.nq, b.nom_v + kp * m.nv, b.nom_q + kn * m.nq, b.nom_v + kn * m.nv, b.cdiff + kn * 2 * m.nv, lane);
}

}  // namespace ilqg
