// dyn.cuh — per-rollout fp64 forward dynamics for small kinematic trees, one CUDA thread per rollout.
//
// This is the device-side replacement of what the reference reaches through
// mj_forward / mj_forwardSkip / mj_step (/root/reference/src/mjderivative.cpp:64,68,92,124,178;
// /root/reference/inc/ilqr.h:86,128).  It is written for sm_100a only, from the pipeline
// description in SURVEY.md Appendix A — not from the CPU oracle's code: every loop over bodies,
// joints, dofs, geoms and collision pairs is unrolled at compile time over a Topo_* description
// (topo_gen.h) so the mjData-equivalent of a rollout (frames, spatial inertias, motion axes, mass
// matrix, factors) lives in registers; only the variable-length constraint rows go to local memory.
// Model constants arrive in DevModel<T> as a __grid_constant__ kernel parameter (constant bank).
//
// Because every perturbed evaluation recomputes all stages from its own inputs, the reference's
// stage skipping (mjSTAGE_POS / mjSTAGE_VEL) needs no special handling: re-evaluating a skipped
// stage from unperturbed inputs reproduces the centre's values exactly.  Only the solver warm start
// is carried from the centre (mjderivative.cpp:75,91).
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

#include "../../include/ilqg_b200.h"
#include "topo_gen.h"
#include "planar.h"

namespace ilqg {

#define ILQG_MINVAL 1e-15
// unrolling of the solver's loops over the constraint rows (local memory): independent loads of the next rows issue early
#ifndef ILQG_ROW_UNROLL
#define ILQG_ROW_UNROLL 2
#endif
#ifndef ILQG_HESS_UNROLL
#define ILQG_HESS_UNROLL 0
#endif
#define ILQG_PRAGMA_(x) _Pragma(#x)
#define ILQG_PRAGMA(x) ILQG_PRAGMA_(x)
#if ILQG_ROW_UNROLL > 0
#define ILQG_ROW_PRAGMA ILQG_PRAGMA(unroll ILQG_ROW_UNROLL)
#else
#define ILQG_ROW_PRAGMA
#endif
#if ILQG_HESS_UNROLL > 0
#define ILQG_HESS_PRAGMA ILQG_PRAGMA(unroll ILQG_HESS_UNROLL)
#else
#define ILQG_HESS_PRAGMA
#endif
#define DEV __device__ __forceinline__

// Destinations of the deriv blocks.  n = 1: the caller's buffer.  n > 1: the same block is stored to every destination —
// the copies of a knot-sharded horizon's deriv array on the peer GPUs (mapped over NVLink with CUDA IPC), so that the
// all-gather of SURVEY 8(e) happens in the FD kernels' write-out instead of a separate collective.
struct FdDst {
    double* p[ILQG_MAX_PEERS];
    int n;
};

template <int I, int N, class F>
DEV void sfor(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        sfor<I + 1, N>(f);
    }
}
// descending: I = N-1 ... LO
template <int I, int LO, class F>
DEV void sfor_down(F&& f) {
    if constexpr (I >= LO) {
        f(std::integral_constant<int, I>{});
        sfor_down<I - 1, LO>(f);
    }
}
#define IDX(x) (decltype(x)::value)

__host__ __device__ constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
__host__ __device__ constexpr int nz(int n) { return n > 0 ? n : 1; }

// ------------------------------------------------------------------ model constants (kernel parameter)
template <class T>
struct DevModel {
    double timestep, gravity[3], tolerance, meaninertia;
    int iterations, ls_iterations, integrator, pad;
    double body_pos[T::NBODY][3], body_quat[T::NBODY][4], body_mass[T::NBODY], body_ipos[T::NBODY][3];
    double body_inertia[T::NBODY][6], body_invw[T::NBODY];
    double jnt_pos[T::NJNT][3], jnt_axis[T::NJNT][3], jnt_range[T::NJNT][2], jnt_stiffness[T::NJNT], jnt_margin[T::NJNT];
    double jnt_solref[T::NJNT][2], jnt_solimp[T::NJNT][5];
    double qpos0[T::NQ], qpos_spring[T::NQ];
    double dof_armature[T::NV], dof_damping[T::NV], dof_invw[T::NV];
    double geom_size[T::NGEOM][2], geom_pos[T::NGEOM][3], geom_axis[T::NGEOM][3];
    double pair_margin[nz(T::NPAIR)], pair_mu[nz(T::NPAIR)], pair_solref[nz(T::NPAIR)][2], pair_solimp[nz(T::NPAIR)][5];
    double act_gear[nz(T::NU)], act_range[nz(T::NU)][2];
    // per-row constants of mj_makeImpedance precomputed on the host: stiffness K, damping B (refsafe applied)
    // and the impedance when solimp is flat (dmin == dmax), else -1
    double jnt_K[T::NJNT], jnt_B[T::NJNT], jnt_imp[T::NJNT];
    double pair_K[nz(T::NPAIR)], pair_B[nz(T::NPAIR)], pair_imp[nz(T::NPAIR)];
};

inline void host_row_consts(double timestep, const double* solref, const double* solimp, double& K, double& B, double& imp) {
    auto cl = [](double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); };
    double tc = solref[0], dr = solref[1];
    if (tc < 2 * timestep) tc = 2 * timestep;
    double dmin = cl(solimp[0], 1e-4, 0.9999), dmax = cl(solimp[1], 1e-4, 0.9999);
    double kk = dmax * dmax * tc * tc * dr * dr, bb = dmax * tc;
    K = 1.0 / (kk < ILQG_MINVAL ? ILQG_MINVAL : kk);
    B = 2.0 / (bb < ILQG_MINVAL ? ILQG_MINVAL : bb);
    imp = (dmin == dmax || solimp[2] <= ILQG_MINVAL) ? 0.5 * (dmin + dmax) : -1.0;
}

// host: fill from the flat tables; returns false if the tables do not have this topology
template <class T>
bool dev_model_from_tables(const ilqg_model& s, DevModel<T>& d) {
    if (s.nq != T::NQ || s.nv != T::NV || s.nu != T::NU || s.nbody != T::NBODY || s.njnt != T::NJNT || s.ngeom != T::NGEOM ||
        s.npair != T::NPAIR)
        return false;
    for (int b = 0; b < T::NBODY; b++)
        if (s.body_parentid[b] != T::body_parent(b) || s.body_rootid[b] != T::body_root(b) || s.body_jntadr[b] != T::body_jntadr(b) ||
            s.body_jntnum[b] != T::body_jntnum(b) || s.body_dofadr[b] != T::body_dofadr(b) || s.body_dofnum[b] != T::body_dofnum(b))
            return false;
    for (int b = 0; b < T::NBODY; b++)   // constants the kernels fold at compile time
        if (T::body_quat_identity(b) && !(s.body_quat[b][0] == 1 && s.body_quat[b][1] == 0 && s.body_quat[b][2] == 0 && s.body_quat[b][3] == 0))
            return false;
    for (int j = 0; j < T::NJNT; j++)
        if (T::jnt_pos_zero(j) && !(s.jnt_pos[j][0] == 0 && s.jnt_pos[j][1] == 0 && s.jnt_pos[j][2] == 0)) return false;
    for (int j = 0; j < T::NJNT; j++)
        if (s.jnt_type[j] != T::jnt_type(j) || s.jnt_bodyid[j] != T::jnt_body(j) || s.jnt_qposadr[j] != T::jnt_qposadr(j) ||
            s.jnt_dofadr[j] != T::jnt_dofadr(j) || (s.jnt_limited[j] != 0) != (T::jnt_limited(j) != 0) ||
            (s.jnt_stiffness[j] != 0 && !T::jnt_hasspring(j)))
            return false;
    for (int i = 0; i < T::NV; i++)
        if (s.dof_bodyid[i] != T::dof_body(i) || s.dof_jntid[i] != T::dof_jnt(i) || s.dof_parentid[i] != T::dof_parent(i)) return false;
    for (int g = 0; g < T::NGEOM; g++)
        if (s.geom_type[g] != T::geom_type(g) || s.geom_bodyid[g] != T::geom_body(g)) return false;
    for (int p = 0; p < T::NPAIR; p++)
        if (s.pair_geom1[p] != T::pair_g1(p) || s.pair_geom2[p] != T::pair_g2(p) || s.pair_condim[p] != T::pair_condim(p)) return false;
    for (int u = 0; u < T::NU; u++)
        if (s.act_dofid[u] != T::act_dof(u) || (s.act_ctrllimited[u] != 0) != (T::act_limited(u) != 0)) return false;
    bool damped = false;
    for (int i = 0; i < T::NV; i++) damped |= s.dof_damping[i] > 0;
    if (damped && !T::ANY_DAMPING) return false;
    if (T::PLANAR_Y && !model_is_planar_y(s)) return false;   // the kernels' static zero pattern (Alg<true>) must hold exactly

    d.timestep = s.timestep;
    for (int k = 0; k < 3; k++) d.gravity[k] = s.gravity[k];
    d.tolerance = s.tolerance;
    d.meaninertia = s.meaninertia;
    d.iterations = s.iterations;
    d.ls_iterations = s.ls_iterations;
    d.integrator = s.integrator;
    d.pad = 0;
    for (int b = 0; b < T::NBODY; b++) {
        for (int k = 0; k < 3; k++) { d.body_pos[b][k] = s.body_pos[b][k]; d.body_ipos[b][k] = s.body_ipos[b][k]; }
        for (int k = 0; k < 4; k++) d.body_quat[b][k] = s.body_quat[b][k];
        for (int k = 0; k < 6; k++) d.body_inertia[b][k] = s.body_inertia[b][k];
        d.body_mass[b] = s.body_mass[b];
        d.body_invw[b] = s.body_invweight0[b][0];
    }
    for (int j = 0; j < T::NJNT; j++) {
        for (int k = 0; k < 3; k++) { d.jnt_pos[j][k] = s.jnt_pos[j][k]; d.jnt_axis[j][k] = s.jnt_axis[j][k]; }
        for (int k = 0; k < 2; k++) { d.jnt_range[j][k] = s.jnt_range[j][k]; d.jnt_solref[j][k] = s.jnt_solref[j][k]; }
        for (int k = 0; k < 5; k++) d.jnt_solimp[j][k] = s.jnt_solimp[j][k];
        d.jnt_stiffness[j] = s.jnt_stiffness[j];
        d.jnt_margin[j] = s.jnt_margin[j];
        host_row_consts(s.timestep, s.jnt_solref[j], s.jnt_solimp[j], d.jnt_K[j], d.jnt_B[j], d.jnt_imp[j]);
    }
    for (int i = 0; i < T::NQ; i++) { d.qpos0[i] = s.qpos0[i]; d.qpos_spring[i] = s.qpos_spring[i]; }
    for (int i = 0; i < T::NV; i++) { d.dof_armature[i] = s.dof_armature[i]; d.dof_damping[i] = s.dof_damping[i]; d.dof_invw[i] = s.dof_invweight0[i]; }
    for (int g = 0; g < T::NGEOM; g++) {
        d.geom_size[g][0] = s.geom_size[g][0];
        d.geom_size[g][1] = s.geom_size[g][1];
        for (int k = 0; k < 3; k++) d.geom_pos[g][k] = s.geom_pos[g][k];
        const double* q = s.geom_quat[g];  // local +z axis of the geom in the body frame (third column of its rotation)
        d.geom_axis[g][0] = 2 * (q[1] * q[3] + q[0] * q[2]);
        d.geom_axis[g][1] = 2 * (q[2] * q[3] - q[0] * q[1]);
        d.geom_axis[g][2] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
    }
    for (int p = 0; p < T::NPAIR; p++) {
        d.pair_margin[p] = s.pair_margin[p];
        d.pair_mu[p] = s.pair_friction[p];
        for (int k = 0; k < 2; k++) d.pair_solref[p][k] = s.pair_solref[p][k];
        for (int k = 0; k < 5; k++) d.pair_solimp[p][k] = s.pair_solimp[p][k];
        host_row_consts(s.timestep, s.pair_solref[p], s.pair_solimp[p], d.pair_K[p], d.pair_B[p], d.pair_imp[p]);
    }
    for (int u = 0; u < T::NU; u++) {
        d.act_gear[u] = s.act_gear[u];
        d.act_range[u][0] = s.act_ctrlrange[u][0];
        d.act_range[u][1] = s.act_ctrlrange[u][1];
    }
    return true;
}

// ------------------------------------------------------------------ small math
struct V3 { double x, y, z; };
DEV V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
DEV V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
DEV V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
DEV double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DEV V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
DEV V3 ld3(const double* p) { return {p[0], p[1], p[2]}; }
DEV V3 normalized(V3 a) {
    double n = sqrt(dot(a, a));
    if (n < ILQG_MINVAL) return {1, 0, 0};
    double r = 1.0 / n;
    return {a.x * r, a.y * r, a.z * r};
}
struct Q4 { double w, x, y, z; };
DEV Q4 qmul(Q4 a, Q4 b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
DEV Q4 qnormalized(Q4 q) {
    double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    if (n < ILQG_MINVAL) return {1, 0, 0, 0};
    double r = 1.0 / n;
    return {q.w * r, q.x * r, q.y * r, q.z * r};
}
struct M3 { V3 r0, r1, r2; };  // rows
DEV M3 q2m(Q4 q) {
    double w = q.w, x = q.x, y = q.y, z = q.z;
    return {{w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)},
            {2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)},
            {2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z}};
}
DEV V3 mulv(const M3& R, V3 v) { return {dot(R.r0, v), dot(R.r1, v), dot(R.r2, v)}; }
DEV V3 col(const M3& R, int k) { return k == 0 ? V3{R.r0.x, R.r1.x, R.r2.x} : k == 1 ? V3{R.r0.y, R.r1.y, R.r2.y} : V3{R.r0.z, R.r1.z, R.r2.z}; }
DEV double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// spatial vectors: [angular; linear] about the tree's centre of mass
struct S6 { V3 w, v; };
DEV S6 operator+(S6 a, S6 b) { return {a.w + b.w, a.v + b.v}; }
DEV S6 operator*(double s, S6 a) { return {s * a.w, s * a.v}; }
DEV double dot(S6 a, S6 b) { return dot(a.w, b.w) + dot(a.v, b.v); }
DEV S6 cross_motion(S6 vel, S6 x) { return {cross(vel.w, x.w), cross(vel.w, x.v) + cross(vel.v, x.w)}; }
DEV S6 cross_force(S6 vel, S6 f) { return {cross(vel.w, f.w) + cross(vel.v, f.v), cross(vel.w, f.v)}; }
// spatial inertia about the tree com: symmetric rotational part, first moment h = m (c - com), mass
struct Inert { double xx, yy, zz, xy, xz, yz; V3 h; double m; };
DEV Inert operator+(const Inert& a, const Inert& b) {
    return {a.xx + b.xx, a.yy + b.yy, a.zz + b.zz, a.xy + b.xy, a.xz + b.xz, a.yz + b.yz, a.h + b.h, a.m + b.m};
}
DEV S6 mul(const Inert& I, S6 s) {
    V3 Iw = {I.xx * s.w.x + I.xy * s.w.y + I.xz * s.w.z, I.xy * s.w.x + I.yy * s.w.y + I.yz * s.w.z,
             I.xz * s.w.x + I.yz * s.w.y + I.zz * s.w.z};
    return {Iw + cross(I.h, s.v), I.m * s.v + cross(s.w, I.h)};
}

// ------------------------------------------------------------------ statically sparse vectors (planar trees)
// A tree whose joints are slides in the xz-plane and hinges about y, with every frame, offset and geom axis in that plane
// (inverted pendulum, hopper — and most locomotion benchmarks), moves in a 3-dimensional subspace of the spatial algebra:
// angular parts are (0, wy, 0), linear parts and positions (x, 0, z), quaternions (w, 0, y, 0), inertias couple x and z only.
// MuJoCo computes the full 6-D algebra and multiplies by those zeros at run time.  Here the zero pattern is part of the
// TYPE: a component of type Z0 is exactly zero, products with it are Z0, sums drop it — the same formulas (one copy of the
// pipeline below, written once over the type family Alg<PLANAR>) instantiate to roughly a third of the arithmetic and half
// of the live registers.  Dropping a term that is exactly 0.0 leaves every other term's value unchanged; only the
// association of the remaining FMAs may differ from the dense instantiation (last-bit differences, far inside the FD
// tolerance).  The topology flag T::PLANAR_Y is verified against the runtime tables when a model is bound.
struct Z0 {
    constexpr Z0() = default;
    DEV constexpr Z0(int) {}   // from the literal 0 of a brace initialiser (a run-time double cannot narrow into it)
};
DEV constexpr Z0 operator+(Z0, Z0) { return {}; }
DEV constexpr Z0 operator-(Z0, Z0) { return {}; }
DEV constexpr Z0 operator*(Z0, Z0) { return {}; }
DEV constexpr Z0 operator-(Z0) { return {}; }
DEV constexpr double operator+(Z0, double b) { return b; }
DEV constexpr double operator+(double a, Z0) { return a; }
DEV constexpr double operator-(double a, Z0) { return a; }
DEV constexpr double operator-(Z0, double b) { return -b; }
DEV constexpr Z0 operator*(Z0, double) { return {}; }
DEV constexpr Z0 operator*(double, Z0) { return {}; }
DEV constexpr Z0 operator*(int, Z0) { return {}; }
DEV constexpr double val(double x) { return x; }
DEV constexpr double val(Z0) { return 0.0; }

template <class X, class Y, class W> struct V3T { X x; Y y; W z; };
template <class X, class Y, class W> DEV V3T<X, Y, W> mk3(X x, Y y, W z) { return {x, y, z}; }
#define V3T_A template <class A0, class A1, class A2>
#define V3T_AB template <class A0, class A1, class A2, class B0, class B1, class B2>
V3T_AB DEV auto operator+(V3T<A0, A1, A2> a, V3T<B0, B1, B2> b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
V3T_AB DEV auto operator-(V3T<A0, A1, A2> a, V3T<B0, B1, B2> b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
V3T_A DEV auto operator*(double s, V3T<A0, A1, A2> a) { return mk3(s * a.x, s * a.y, s * a.z); }
V3T_AB DEV auto dot(V3T<A0, A1, A2> a, V3T<B0, B1, B2> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
V3T_A DEV auto dot(V3 a, V3T<A0, A1, A2> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
V3T_AB DEV auto cross(V3T<A0, A1, A2> a, V3T<B0, B1, B2> b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
V3T_A DEV V3 full(V3T<A0, A1, A2> a) { return {val(a.x), val(a.y), val(a.z)}; }
DEV V3 full(V3 a) { return a; }

template <class W, class X, class Y, class Zc> struct Q4T { W w; X x; Y y; Zc z; };
template <class W, class X, class Y, class Zc> DEV Q4T<W, X, Y, Zc> mk4(W w, X x, Y y, Zc z) { return {w, x, y, z}; }
template <class A0, class A1, class A2, class A3, class B0, class B1, class B2, class B3>
DEV auto qmul(Q4T<A0, A1, A2, A3> a, Q4T<B0, B1, B2, B3> b) {
    return mk4(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
               a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w);
}
template <class A0, class A1, class A2, class A3>
DEV Q4T<A0, A1, A2, A3> qnormalized(Q4T<A0, A1, A2, A3> q) {
    double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    if (n < ILQG_MINVAL) return {1, 0, 0, 0};
    double r = 1.0 / n;
    return {q.w * r, q.x * r, q.y * r, q.z * r};
}
template <class R0, class R1, class R2> struct M3T { R0 r0; R1 r1; R2 r2; };  // rows
template <class R0, class R1, class R2> DEV M3T<R0, R1, R2> mkm(R0 a, R1 b, R2 c) { return {a, b, c}; }
template <class A0, class A1, class A2, class A3>
DEV auto q2m(Q4T<A0, A1, A2, A3> q) {
    auto w = q.w; auto x = q.x; auto y = q.y; auto z = q.z;
    return mkm(mk3(w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)),
               mk3(2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)),
               mk3(2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z));
}
template <class R0, class R1, class R2, class A0, class A1, class A2>
DEV auto mulv(const M3T<R0, R1, R2>& R, V3T<A0, A1, A2> v) { return mk3(dot(R.r0, v), dot(R.r1, v), dot(R.r2, v)); }

template <class W, class V> struct S6T { W w; V v; };
template <class W, class V> DEV S6T<W, V> mk6(W w, V v) { return {w, v}; }
#define S6T_AB template <class AW, class AV, class BW, class BV>
S6T_AB DEV auto operator+(S6T<AW, AV> a, S6T<BW, BV> b) { return mk6(a.w + b.w, a.v + b.v); }
template <class AW, class AV> DEV auto operator*(double s, S6T<AW, AV> a) { return mk6(s * a.w, s * a.v); }
S6T_AB DEV auto dot(S6T<AW, AV> a, S6T<BW, BV> b) { return dot(a.w, b.w) + dot(a.v, b.v); }
S6T_AB DEV auto cross_motion(S6T<AW, AV> vel, S6T<BW, BV> x) { return mk6(cross(vel.w, x.w), cross(vel.w, x.v) + cross(vel.v, x.w)); }
S6T_AB DEV auto cross_force(S6T<AW, AV> vel, S6T<BW, BV> f) { return mk6(cross(vel.w, f.w) + cross(vel.v, f.v), cross(vel.w, f.v)); }
template <class XX, class YY, class ZZ, class XY, class XZ, class YZ, class H>
struct InertT { XX xx; YY yy; ZZ zz; XY xy; XZ xz; YZ yz; H h; double m; };
#define INERTT_A template <class XX, class YY, class ZZ, class XY, class XZ, class YZ, class H>
INERTT_A DEV InertT<XX, YY, ZZ, XY, XZ, YZ, H> operator+(const InertT<XX, YY, ZZ, XY, XZ, YZ, H>& a, const InertT<XX, YY, ZZ, XY, XZ, YZ, H>& b) {
    return {a.xx + b.xx, a.yy + b.yy, a.zz + b.zz, a.xy + b.xy, a.xz + b.xz, a.yz + b.yz, a.h + b.h, a.m + b.m};
}
template <class XX, class YY, class ZZ, class XY, class XZ, class YZ, class H, class AW, class AV>
DEV auto mul(const InertT<XX, YY, ZZ, XY, XZ, YZ, H>& I, S6T<AW, AV> s) {
    auto Iw = mk3(I.xx * s.w.x + I.xy * s.w.y + I.xz * s.w.z, I.xy * s.w.x + I.yy * s.w.y + I.yz * s.w.z,
                  I.xz * s.w.x + I.yz * s.w.y + I.zz * s.w.z);
    return mk6(Iw + cross(I.h, s.v), I.m * s.v + cross(s.w, I.h));
}
// body-frame inertia tensor as loaded from the model tables
template <class XY, class YZ> struct IbT { double xx, yy, zz; XY xy; double xz; YZ yz; };

// The type family the pipeline is written over.  Alg<false>: the dense 3-D types above.  Alg<true>: the planar patterns.
template <bool PLANAR> struct Alg;
template <> struct Alg<false> {
    using P3 = V3; using A3 = V3; using QT = Q4; using MT = M3; using ST = S6; using CD = S6; using IT = Inert;
    using IB = IbT<double, double>;
    static DEV P3 ldP(const double* p) { return {p[0], p[1], p[2]}; }
    static DEV A3 ldA(const double* p) { return {p[0], p[1], p[2]}; }
    static DEV QT ldQ(const double* p) { return {p[0], p[1], p[2], p[3]}; }
    static DEV QT hinge_quat(double c, double s, const double* ax) { return {c, ax[0] * s, ax[1] * s, ax[2] * s}; }
    static DEV IB ldI(const double* in) { return {in[0], in[1], in[2], in[3], in[4], in[5]}; }
};
template <> struct Alg<true> {
    using P3 = V3T<double, Z0, double>;   // positions, linear parts, in-plane axes
    using A3 = V3T<Z0, double, Z0>;       // angular parts, hinge axes
    using QT = Q4T<double, Z0, double, Z0>;
    using MT = M3T<P3, A3, P3>;
    using ST = S6T<A3, P3>;               // motion and force vectors have the same pattern
    using CD = S6T<V3T<Z0, Z0, Z0>, P3>;  // cdof_dot: the angular part of (planar) x (planar) vanishes
    using IT = InertT<double, double, double, Z0, double, Z0, P3>;
    using IB = IbT<Z0, Z0>;
    static DEV P3 ldP(const double* p) { return {p[0], 0, p[2]}; }
    static DEV A3 ldA(const double* p) { return {0, p[1], 0}; }
    static DEV QT ldQ(const double* p) { return {p[0], 0, p[2], 0}; }
    static DEV QT hinge_quat(double c, double s, const double* ax) { return {c, 0, ax[1] * s, 0}; }
    static DEV IB ldI(const double* in) { return {in[0], in[1], in[2], 0, in[4], 0}; }
};
template <class T> using AlgOf = Alg<(T::PLANAR_Y != 0)>;

// ------------------------------------------------------------------ per-rollout solver inputs
// Where a rollout's constraint rows live.  The row count is data dependent (0 in flight, 4 per contact point, 1 per joint at its
// limit; up to T::MAXEFC), so the rows cannot be registers.  Two placements behind one interface (element accessors):
//   RowsLocal   per-thread local memory (L1-resident as long as the CTA's rows fit there) — every kernel but one;
//   RowsShared  shared memory, ONE copy of J / D / B / k-term per KNOT for the threads that linearise the knot's qvel / ctrl
//               columns on the same position stage (fd_velctrl_kernel), plus per-thread aref / jar / jv — see there.
template <class T>
struct RowsLocal {
    static constexpr int NV = T::NV, ME = nz(T::MAXEFC);
    double J_[ME][NV];
    double D_[ME], aref_[ME], jar_[ME], jv_[ME];
    double rB_[ME], rkt_[ME];  // per-row damping B and stiffness term K*imp*(pos-margin): aref = -B (J qvel) - rkt
    DEV double& J(int r, int i) { return J_[r][i]; }
    DEV double& D(int r) { return D_[r]; }
    DEV double& aref(int r) { return aref_[r]; }
    DEV double& jar(int r) { return jar_[r]; }
    DEV double& jv(int r) { return jv_[r]; }
    DEV double& rB(int r) { return rB_[r]; }
    DEV double& rkt(int r) { return rkt_[r]; }
    static constexpr bool private_rows = true;   // every thread owns (and writes) its rows
    DEV bool row_writer() const { return true; }
    DEV int capacity() const { return ME; }
};
// Shared-memory placement.  `knot`: base of the knot's block, element (field f, row r) at (r * KF + f) * kstride — consecutive
// knots of the CTA at consecutive addresses, the threads of one knot read the same word (a broadcast): conflict-free.
// `mine`: base of the thread's private block, element (field f, row r) at (r * 3 + f) * tstride.
template <class T>
struct RowsShared {
    static constexpr int NV = T::NV, KF = NV + 3;   // J[nv], D, B, k-term per row of the knot's block
    double* knot;
    double* mine;
    int kstride, tstride;
    int cap;       // rows the blocks hold (a knot with more is flagged Work::overflow and redone with RowsLocal)
    bool writer;   // one thread of the knot's group stores the shared fields (all compute the same values)
    static constexpr bool private_rows = false;
    DEV double& J(int r, int i) { return knot[(r * KF + i) * kstride]; }
    DEV double& D(int r) { return knot[(r * KF + NV) * kstride]; }
    DEV double& rB(int r) { return knot[(r * KF + NV + 1) * kstride]; }
    DEV double& rkt(int r) { return knot[(r * KF + NV + 2) * kstride]; }
    DEV double& aref(int r) { return mine[(r * 3 + 0) * tstride]; }
    DEV double& jar(int r) { return mine[(r * 3 + 1) * tstride]; }
    DEV double& jv(int r) { return mine[(r * 3 + 2) * tstride]; }
    DEV bool row_writer() const { return writer; }
    DEV int capacity() const { return cap; }
};

template <class T, class R = RowsLocal<T>>
struct Work {
    static constexpr int NV = T::NV, NT = T::NV * (T::NV + 1) / 2, ME = nz(T::MAXEFC);
    double M[NT];    // mass matrix (packed lower)
    double L[NT];    // Cholesky factor of M, diagonal entries hold 1/L_ii
    double fs[NV];   // qfrc_smooth
    double fb[NV];   // qfrc_passive - qfrc_bias (the velocity stage's product; qfrc_smooth = fb + actuation)
    double as[NV];   // qacc_smooth
    double fc[NV];   // qfrc_constraint of the last solve
    int nefc;
    int ncon;        // contact points the narrow phase found (diagnostics)
    int iters;       // Newton iterations of the last solve
    int exact;       // the last solve left through the exact-optimum test (or had no rows): solving again from its result is a no-op
    int overflow;    // more rows than the placement's capacity (RowsShared): the knot must be redone with RowsLocal
    using RowsT = R;
    R rows;
};

// what the position stage leaves for the velocity stage (mjSTAGE_POS products that mj_fwdVelocity reads)
template <class T>
struct PosStage {
    typename AlgOf<T>::ST cdof[T::NV];
    typename AlgOf<T>::IT cin[T::NBODY];
    double dspr[T::NV];  // q - qpos_spring of sprung dofs
};

// ---- flat (de)serialisation of the statically sparse types: only the components that exist are stored
template <bool LOAD> DEV void xfer(double*& p, double& x) { if constexpr (LOAD) x = *p; else *p = x; p++; }
template <bool LOAD> DEV void xfer(double*&, Z0&) {}
template <bool LOAD> DEV void xfer(double*& p, V3& a) { xfer<LOAD>(p, a.x); xfer<LOAD>(p, a.y); xfer<LOAD>(p, a.z); }
template <bool LOAD, class X, class Y, class W> DEV void xfer(double*& p, V3T<X, Y, W>& a) { xfer<LOAD>(p, a.x); xfer<LOAD>(p, a.y); xfer<LOAD>(p, a.z); }
template <bool LOAD> DEV void xfer(double*& p, S6& a) { xfer<LOAD>(p, a.w); xfer<LOAD>(p, a.v); }
template <bool LOAD, class W, class V> DEV void xfer(double*& p, S6T<W, V>& a) { xfer<LOAD>(p, a.w); xfer<LOAD>(p, a.v); }
template <bool LOAD> DEV void xfer(double*& p, Inert& a) {
    xfer<LOAD>(p, a.xx); xfer<LOAD>(p, a.yy); xfer<LOAD>(p, a.zz); xfer<LOAD>(p, a.xy); xfer<LOAD>(p, a.xz); xfer<LOAD>(p, a.yz); xfer<LOAD>(p, a.h); xfer<LOAD>(p, a.m);
}
template <bool LOAD, class XX, class YY, class ZZ, class XY, class XZ, class YZ, class H> DEV void xfer(double*& p, InertT<XX, YY, ZZ, XY, XZ, YZ, H>& a) {
    xfer<LOAD>(p, a.xx); xfer<LOAD>(p, a.yy); xfer<LOAD>(p, a.zz); xfer<LOAD>(p, a.xy); xfer<LOAD>(p, a.xz); xfer<LOAD>(p, a.yz); xfer<LOAD>(p, a.h); xfer<LOAD>(p, a.m);
}
// The position stage's products of one rollout — what mj_forwardSkip(mjSTAGE_POS) keeps (/root/reference/src/mjderivative.cpp:124):
// motion axes, spatial inertias, spring offsets, M and its factor, the constraint rows (J, D, damping B, k-term) — to or from
// a flat record: [nefc | PosStage | M | L | rows].  POS_RECORD_DOUBLES bounds it.
template <class T>
__host__ __device__ constexpr int pos_record_doubles() { return 2 + 6 * T::NV + 10 * T::NBODY + T::NV + T::NV * (T::NV + 1) + nz(T::MAXEFC) * (T::NV + 3); }
template <bool LOAD, class T, class W>
DEV void xfer_pos_stage(double* rec, PosStage<T>& ps, W& w) {
    double* p = rec;
    double ne = (double)w.nefc;
    xfer<LOAD>(p, ne);
    if constexpr (LOAD) w.nefc = (int)ne;
    p++;   // (keeps what follows 16-byte aligned)
    sfor<0, T::NV>([&](auto ii) { xfer<LOAD>(p, ps.cdof[IDX(ii)]); });
    sfor<1, T::NBODY>([&](auto bb) { xfer<LOAD>(p, ps.cin[IDX(bb)]); });
    sfor<0, T::NV>([&](auto ii) { xfer<LOAD>(p, ps.dspr[IDX(ii)]); });
    sfor<0, T::NV*(T::NV + 1) / 2>([&](auto tt) { xfer<LOAD>(p, w.M[IDX(tt)]); });
    sfor<0, T::NV*(T::NV + 1) / 2>([&](auto tt) { xfer<LOAD>(p, w.L[IDX(tt)]); });
    for (int r = 0; r < w.nefc; r++) {
        sfor<0, T::NV>([&](auto ii) { xfer<LOAD>(p, w.rows.J(r, IDX(ii))); });
        xfer<LOAD>(p, w.rows.D(r));
        xfer<LOAD>(p, w.rows.rB(r));
        xfer<LOAD>(p, w.rows.rkt(r));
    }
}

// packed Cholesky A = L L^T with reciprocal diagonal
template <int N>
DEV void chol_packed(const double* A, double* L) {
    sfor<0, N>([&](auto ii) {
        constexpr int i = IDX(ii);
        sfor<0, i + 1>([&](auto jj) {
            constexpr int j = IDX(jj);
            double s = A[tri(i, j)];
            sfor<0, j>([&](auto kk) { constexpr int k = IDX(kk); s -= L[tri(i, k)] * L[tri(j, k)]; });
            if constexpr (i == j) {
                if (s < ILQG_MINVAL) s = ILQG_MINVAL;
                L[tri(i, i)] = rsqrt(s);
            } else
                L[tri(i, j)] = s * L[tri(j, j)];
        });
    });
}
template <int N>
DEV void chol_solve_packed(const double* L, double* x) {
    sfor<0, N>([&](auto ii) {
        constexpr int i = IDX(ii);
        double s = x[i];
        sfor<0, i>([&](auto kk) { constexpr int k = IDX(kk); s -= L[tri(i, k)] * x[k]; });
        x[i] = s * L[tri(i, i)];
    });
    sfor_down<N - 1, 0>([&](auto ii) {
        constexpr int i = IDX(ii);
        double s = x[i];
        sfor<i + 1, N>([&](auto kk) { constexpr int k = IDX(kk); s -= L[tri(k, i)] * x[k]; });
        x[i] = s * L[tri(i, i)];
    });
}

template <class T>
__host__ __device__ constexpr bool dof_is_ancestor(int a, int i) {  // a is i or an ancestor of i in the dof chain
    while (i >= 0) {
        if (i == a) return true;
        i = T::dof_parent(i);
    }
    return false;
}
template <class T>
__host__ __device__ constexpr bool dof_moves_body(int dof, int body) {
    int b = body;
    while (b > 0) {
        if (T::dof_body(dof) == b) return true;
        b = T::body_parent(b);
    }
    return false;
}

// position-dependent impedance d(r) (getimpedance); one out-of-line copy — it holds two pow() expansions and
// is only reached for non-flat solimp
__device__ __noinline__ double impedance(const double* solimp, double pos, double margin) {
    double dmin = clampd(solimp[0], 1e-4, 0.9999), dmax = clampd(solimp[1], 1e-4, 0.9999);
    double width = solimp[2], mid = clampd(solimp[3], 1e-4, 0.9999), power = solimp[4] < 1 ? 1 : solimp[4];
    if (dmin == dmax || width <= ILQG_MINVAL) return 0.5 * (dmin + dmax);
    double x = fabs(pos - margin) / width;
    if (x >= 1) return dmax;
    if (x <= 0) return dmin;
    double y;
    if (power == 1) y = x;
    else if (power == 2) y = x <= mid ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);
    else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
    else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
    return dmin + y * (dmax - dmin);
}

// regulariser R and stiffness term K*imp*(pos-margin) of one row
// (mj_makeImpedance + mj_referenceConstraint: aref = -B*vel - kterm); K, B, flat impedance come precomputed
DEV void row_params(double K, double flat_imp, const double* solimp, double pos, double margin, double diagApprox, double& R,
                    double& kterm) {
    double imp = flat_imp >= 0 ? flat_imp : impedance(solimp, pos, margin);
    R = (1 - imp) / imp * diagApprox;
    if (R < ILQG_MINVAL) R = ILQG_MINVAL;
    kterm = K * imp * (pos - margin);
}

// contact frame from the normal and an optional tangent hint (mju_makeFrame)
DEV void make_frame(V3 n, V3 hint, bool has_hint, V3& t1, V3& t2) {
    V3 y = hint;
    if (!has_hint || sqrt(dot(y, y)) < 0.5) y = (n.y < 0.5 && n.y > -0.5) ? V3{0, 1, 0} : V3{0, 0, 1};
    y = y - dot(n, y) * n;
    if (sqrt(dot(y, y)) < 1e-12) {
        y = (n.y < 0.5 && n.y > -0.5) ? V3{0, 1, 0} : V3{0, 0, 1};
        y = y - dot(n, y) * n;
    }
    t1 = normalized(y);
    t2 = cross(n, t1);
}

// Free joints exist in dense (non-planar) trees only.  The explicit template argument makes the call sites dependent on the
// enclosing generic lambda's parameter, so the planar instantiation of the pipeline never looks inside.
template <int J, class P3, class Q4, class M3>
DEV void kin_free_joint(const double* q7, P3& pos, Q4& quat, M3& R) {
    pos = {q7[0], q7[1], q7[2]};
    quat = qnormalized(Q4{q7[3], q7[4], q7[5], q7[6]});
    R = q2m(quat);
}
template <int J, class M3, class P3, class S6>
DEV void cdof_free_rot(const M3& xmat, P3 off, S6* cdof3) {
    sfor<0, 3>([&](auto kk) { P3 a = col(xmat, IDX(kk)); cdof3[IDX(kk)] = {a, cross(a, off)}; });
}

// ------------------------------------------------------------------ the pipeline up to the constraint problem
// Split along MuJoCo's stage boundaries (mjSTAGE_POS / mjSTAGE_VEL, the skip levels the reference passes to
// mj_forwardSkip, /root/reference/src/mjderivative.cpp:92,124,178):
//   build_pos     everything that depends on qpos only: frames, com, spatial inertias, motion axes, M and its factor,
//                 the constraint rows' J, D and the per-row constants (B, K*imp*(pos-margin))
//   build_vel     everything that also depends on qvel: bias forces (RNE), passive forces, the rows' aref
//   finish_smooth actuation, qfrc_smooth, qacc_smooth
// A qvel / ctrl column of the FD re-runs only the later stages on the centre's position-stage products.
// SYNC: block-wide barriers between the stages.  The stage code is long and straight-line (fully unrolled); without them the
// warps of a CTA drift apart and each streams its own copy of the instructions through the instruction caches (ncu: the
// second-largest stall reason was "no instruction").  With them the CTA's warps walk the code together and share fetches.
// Every thread of the block must call the stage functions when SYNC is set.
// FUSED (qv != nullptr): the velocity is known while the rows are built — aref is finished here from the Jacobian still in
// registers and the per-row constants (B, k-term) never go to local memory; build_vel then skips its pass over the rows.
template <class T, bool SYNC = false, bool FUSED = false, class W>
DEV void build_vel(const DevModel<T>& m, const PosStage<T>& ps, const double (&qv)[T::NV], W& w);
template <class T, class W>
DEV void finish_smooth(const DevModel<T>& m, const double (&u)[nz(T::NU)], W& w);

template <class T, bool SYNC = false, bool FUSED = false, class W>
DEV void build_pos(const DevModel<T>& m, const double (&q)[T::NQ], PosStage<T>& ps, W& w, const double* qv = nullptr,
                   const double* uu = nullptr) {
    auto stage_sync = [&]() { if constexpr (SYNC) __syncthreads(); };
    constexpr int NB = T::NBODY, NV = T::NV, NJ = T::NJNT;
    using A = AlgOf<T>;   // dense 3-D types, or the planar patterns (see Alg)
    using P3 = typename A::P3; using A3 = typename A::A3; using Q4 = typename A::QT; using M3 = typename A::MT;
    using S6 = typename A::ST; using Inert = typename A::IT;
    P3 xpos[NB];
    M3 xmat[NB];
    Q4 xquat[NB];
    P3 anchor[NJ], axisP[NJ];   // slide axes (in-plane for planar trees)
    A3 axisA[NJ];               // hinge axes
    xpos[0] = {0, 0, 0};
    xquat[0] = {1, 0, 0, 0};
    xmat[0] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    // ---- mj_kinematics
    // Exact folds (topology flags, verified against the runtime tables when the model is bound): an identity body quaternion
    // leaves the parent's frame untouched (q * 1 = q and q2m of the same quaternion is the parent's xmat, bit for bit), a joint
    // at the body origin has anchor = pos and no off-centre correction (R * 0 = 0).
    sfor<1, NB>([&](auto bb) {
        constexpr int b = IDX(bb), p = T::body_parent(b);
        constexpr bool qid = T::body_quat_identity(b) != 0;
        P3 pos;
        Q4 quat;
        M3 R;   // rotation of `quat`, valid while only slide joints have been applied (see slide_only below)
        if constexpr (p == 0) {
            pos = A::ldP(m.body_pos[b]);
            if constexpr (qid) { quat = {1, 0, 0, 0}; R = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}; }
            else { quat = A::ldQ(m.body_quat[b]); R = q2m(quat); }
        } else {
            pos = xpos[p] + mulv(xmat[p], A::ldP(m.body_pos[b]));
            if constexpr (qid) { quat = xquat[p]; R = xmat[p]; }
            else { quat = qmul(xquat[p], A::ldQ(m.body_quat[b])); R = q2m(quat); }
        }
        sfor<0, T::body_jntnum(b)>([&](auto jj) {
            constexpr int j = T::body_jntadr(b) + IDX(jj), qa = T::jnt_qposadr(j), ty = T::jnt_type(j);
            if constexpr (ty == ILQG_JNT_FREE) {
                kin_free_joint<j>(&q[qa], pos, quat, R);
                anchor[j] = pos;
            } else {
                constexpr bool jz = T::jnt_pos_zero(j) != 0;
                if constexpr (jz) anchor[j] = pos;
                else anchor[j] = pos + mulv(R, A::ldP(m.jnt_pos[j]));
                double qq = q[qa] - m.qpos0[qa];
                if constexpr (ty == ILQG_JNT_SLIDE) {
                    axisP[j] = mulv(R, A::ldP(m.jnt_axis[j]));
                    pos = pos + qq * axisP[j];
                } else {
                    axisA[j] = mulv(R, A::ldA(m.jnt_axis[j]));
                    double s, c;
                    sincos(0.5 * qq, &s, &c);
                    quat = qmul(quat, A::hinge_quat(c, s, m.jnt_axis[j]));
                    constexpr bool last = IDX(jj) + 1 == T::body_jntnum(b);
                    if constexpr (!jz || !last) R = q2m(quat);   // the rotated frame: off-centre correction and / or the next joint
                    if constexpr (jz) pos = anchor[j];
                    else pos = anchor[j] - mulv(R, A::ldP(m.jnt_pos[j]));
                }
            }
        });
        quat = qnormalized(quat);
        xpos[b] = pos;
        xquat[b] = quat;
        xmat[b] = q2m(quat);
    });
    stage_sync();
    // contact records (narrow phase -> row construction); the narrow phase runs right after the kinematics so that the frames
    // (xpos, xmat, geom frames) are dead while M is assembled and factorised — fewer live registers where the pressure peaks
    constexpr int MC = nz(T::MAXCON);
    double cdist[MC];
    P3 cpos[MC], cnrm[MC], chint[MC];
    int cpair[MC];   // pair index, bit 8 set when the tangent hint is valid
    int nc = 0;
    if constexpr (T::NPAIR > 0) {
        // geom frames
        P3 gpos[T::NGEOM], gax[T::NGEOM];
        sfor<0, T::NGEOM>([&](auto gg) {
            constexpr int g = IDX(gg), b = T::geom_body(g);
            if constexpr (b == 0) { gpos[g] = A::ldP(m.geom_pos[g]); gax[g] = A::ldP(m.geom_axis[g]); }
            else { gpos[g] = xpos[b] + mulv(xmat[b], A::ldP(m.geom_pos[g])); gax[g] = mulv(xmat[b], A::ldP(m.geom_axis[g])); }
        });
        // Phase A (unrolled over the model's pair list): narrow phase only — every contact found is pushed as a small
        // record.  Phase B (one runtime loop over the records) builds the rows.  The row construction is by far the
        // largest piece of code of the pipeline; keeping ONE copy of it instead of one per pair shrinks the kernel's
        // instruction footprint by a third and lets lanes whose contacts come from different pairs share the same code.
        auto push = [&](int p, double dist, P3 pos, P3 n, P3 hint, bool has_hint) {
            if (nc < MC) { cdist[nc] = dist; cpos[nc] = pos; cnrm[nc] = n; chint[nc] = hint; cpair[nc] = p | (has_hint ? 256 : 0); nc++; }
        };
        sfor<0, T::NPAIR>([&](auto pp) {
            constexpr int p = IDX(pp), g1 = T::pair_g1(p), g2 = T::pair_g2(p), t1 = T::geom_type(g1), t2 = T::geom_type(g2);
            const double margin = m.pair_margin[p];
            auto sphere_sphere = [&](P3 p1, double r1, P3 p2, double r2) {
                P3 n = p2 - p1;
                double len = sqrt(dot(n, n));
                double dist = len - r1 - r2;
                if (dist > margin) return false;
                if (len < ILQG_MINVAL) n = {1, 0, 0};
                else n = (1.0 / len) * n;
                push(p, dist, p1 + (r1 + 0.5 * dist) * n, n, P3{0, 0, 0}, false);
                return true;
            };
            if constexpr (t1 == ILQG_GEOM_PLANE && (t2 == ILQG_GEOM_CAPSULE || t2 == ILQG_GEOM_SPHERE)) {
                P3 pn = gax[g1];
                double r = m.geom_size[g2][0];
                auto plane_sphere = [&](P3 c, bool hint) {
                    double dist = dot(c - gpos[g1], pn) - r;
                    if (dist > margin) return;
                    push(p, dist, c - (r + 0.5 * dist) * pn, pn, gax[g2], hint);
                };
                if constexpr (t2 == ILQG_GEOM_SPHERE) plane_sphere(gpos[g2], false);
                else {
                    double h = m.geom_size[g2][1];
#pragma unroll 1
                    for (int e = 0; e < 2; e++) plane_sphere(gpos[g2] + (e ? -h : h) * gax[g2], true);
                }
            } else if constexpr (t1 == ILQG_GEOM_SPHERE && t2 == ILQG_GEOM_SPHERE) {
                sphere_sphere(gpos[g1], m.geom_size[g1][0], gpos[g2], m.geom_size[g2][0]);
            } else if constexpr (t1 == ILQG_GEOM_SPHERE && t2 == ILQG_GEOM_CAPSULE) {
                double h = m.geom_size[g2][1];
                double t = clampd(dot(gpos[g1] - gpos[g2], gax[g2]), -h, h);
                sphere_sphere(gpos[g1], m.geom_size[g1][0], gpos[g2] + t * gax[g2], m.geom_size[g2][0]);
            } else if constexpr (t1 == ILQG_GEOM_CAPSULE && t2 == ILQG_GEOM_CAPSULE) {
                P3 p1 = gpos[g1], a1 = gax[g1], p2 = gpos[g2], a2 = gax[g2];
                double r1 = m.geom_size[g1][0], h1 = m.geom_size[g1][1], r2 = m.geom_size[g2][0], h2 = m.geom_size[g2][1];
                // exact cull: the capsules lie inside spheres of radius h + r about their centres; if even those are farther apart
                // than the margin the narrow phase below cannot produce a contact (its distance is at least this one).  Self-collision
                // pairs are almost always culled here, and the closest-point search is the most expensive piece of the position stage.
                P3 ca[2], cb[2];
                int ncand = 0;
                const double reach = h1 + r1 + h2 + r2 + margin;
                const P3 cc = p1 - p2;
                if (!(reach > 0 && dot(cc, cc) > reach * reach)) {
                // candidate closest-point pairs on the two axis segments (at most two survive the distance test)
                auto consider = [&](P3 c1, P3 c2) {
                    P3 d = c2 - c1;
                    if (sqrt(dot(d, d)) - r1 - r2 > margin) return false;
                    if (ncand == 0) { ca[0] = c1; cb[0] = c2; } else { ca[1] = c1; cb[1] = c2; }
                    ncand++;
                    return true;
                };
                P3 dif = p1 - p2;
                double mb = -dot(a1, a2), uu = -dot(a1, dif), vv = dot(a2, dif);
                double det = 1.0 - mb * mb;
                if (fabs(det) >= 1e-12) {
                    double x1 = (uu - mb * vv) / det, x2 = (vv - mb * uu) / det;
                    if (x1 > h1) { x1 = h1; x2 = vv - mb * h1; }
                    else if (x1 < -h1) { x1 = -h1; x2 = vv + mb * h1; }
                    if (x2 > h2) { x2 = h2; x1 = clampd(uu - mb * h2, -h1, h1); }
                    else if (x2 < -h2) { x2 = -h2; x1 = clampd(uu + mb * h2, -h1, h1); }
                    consider(p1 + x1 * a1, p2 + x2 * a2);
                } else {  // parallel axes: end points against the other segment, at most two contacts
                    for (int s = -1; s <= 1 && ncand < 2; s += 2) {
                        P3 c1 = p1 + (s * h1) * a1;
                        double t = dot(c1 - p2, a2);
                        if (t < -h2 || t > h2) continue;
                        consider(c1, p2 + t * a2);
                    }
                    for (int s = -1; s <= 1 && ncand < 2; s += 2) {
                        P3 c2 = p2 + (s * h2) * a2;
                        double t = dot(c2 - p1, a1);
                        if (t <= -h1 || t >= h1) continue;
                        consider(p1 + t * a1, c2);
                    }
                    if (ncand == 0) {
                        double best = 1e300;
                        P3 bq1 = p1, bq2 = p2;
                        for (int s = -1; s <= 1; s += 2)
                            for (int t = -1; t <= 1; t += 2) {
                                P3 c1 = p1 + (s * h1) * a1, c2 = p2 + (t * h2) * a2;
                                double dd = dot(c1 - c2, c1 - c2);
                                if (dd < best) { best = dd; bq1 = c1; bq2 = c2; }
                            }
                        consider(bq1, bq2);
                    }
                }
                }
#pragma unroll 1
                for (int c = 0; c < ncand; c++) sphere_sphere(c ? ca[1] : ca[0], r1, c ? cb[1] : cb[0], r2);
            }
        });
    }
    // ---- mj_comPos: tree centres of mass, spatial inertias, motion axes
    P3 xipos[NB], com[NB];
    double tmass[NB];
    sfor<1, NB>([&](auto bb) {
        constexpr int b = IDX(bb);
        xipos[b] = xpos[b] + mulv(xmat[b], A::ldP(m.body_ipos[b]));
        if constexpr (T::body_root(b) == b) { com[b] = m.body_mass[b] * xipos[b]; tmass[b] = m.body_mass[b]; }
        else { com[T::body_root(b)] = com[T::body_root(b)] + m.body_mass[b] * xipos[b]; tmass[T::body_root(b)] += m.body_mass[b]; }
    });
    sfor<1, NB>([&](auto bb) {
        constexpr int b = IDX(bb);
        if constexpr (T::body_root(b) == b) com[b] = (1.0 / tmass[b]) * com[b];
    });
    Inert (&cin)[NB] = ps.cin;
    sfor<1, NB>([&](auto bb) {
        constexpr int b = IDX(bb);
        const M3& R = xmat[b];
        const auto in = A::ldI(m.body_inertia[b]);
        // Iw = R Ib R^T
        P3 t0 = {R.r0.x * in.xx + R.r0.y * in.xy + R.r0.z * in.xz, R.r0.x * in.xy + R.r0.y * in.yy + R.r0.z * in.yz,
                 R.r0.x * in.xz + R.r0.y * in.yz + R.r0.z * in.zz};
        A3 t1 = {R.r1.x * in.xx + R.r1.y * in.xy + R.r1.z * in.xz, R.r1.x * in.xy + R.r1.y * in.yy + R.r1.z * in.yz,
                 R.r1.x * in.xz + R.r1.y * in.yz + R.r1.z * in.zz};
        P3 t2 = {R.r2.x * in.xx + R.r2.y * in.xy + R.r2.z * in.xz, R.r2.x * in.xy + R.r2.y * in.yy + R.r2.z * in.yz,
                 R.r2.x * in.xz + R.r2.y * in.yz + R.r2.z * in.zz};
        double ms = m.body_mass[b];
        P3 d = xipos[b] - com[T::body_root(b)];
        double dd = dot(d, d);
        cin[b] = {dot(t0, R.r0) + ms * (dd - d.x * d.x), dot(t1, R.r1) + ms * (dd - d.y * d.y), dot(t2, R.r2) + ms * (dd - d.z * d.z),
                  dot(t0, R.r1) - ms * d.x * d.y, dot(t0, R.r2) - ms * d.x * d.z, dot(t1, R.r2) - ms * d.y * d.z, ms * d, ms};
    });
    S6 (&cdof)[NV] = ps.cdof;
    sfor<0, NJ>([&](auto jj) {
        constexpr int j = IDX(jj), b = T::jnt_body(j), da = T::jnt_dofadr(j), ty = T::jnt_type(j);
        P3 off = com[T::body_root(b)] - anchor[j];
        if constexpr (ty == ILQG_JNT_FREE) {
            cdof[da] = {{0, 0, 0}, {1, 0, 0}};
            cdof[da + 1] = {{0, 0, 0}, {0, 1, 0}};
            cdof[da + 2] = {{0, 0, 0}, {0, 0, 1}};
            cdof_free_rot<j>(xmat[b], off, &cdof[da + 3]);
        } else if constexpr (ty == ILQG_JNT_SLIDE) {
            cdof[da] = {{0, 0, 0}, axisP[j]};
        } else {
            cdof[da] = {axisA[j], cross(axisA[j], off)};
        }
    });
    stage_sync();
    sfor<0, NV>([&](auto ii) {
        constexpr int i = IDX(ii), j = T::dof_jnt(i);
        if constexpr (T::jnt_type(j) != ILQG_JNT_FREE && T::jnt_hasspring(j)) ps.dspr[i] = q[T::jnt_qposadr(j)] - m.qpos_spring[T::jnt_qposadr(j)];
        else ps.dspr[i] = 0;
    });
    // FUSED (velocity and controls known): the bias recursion runs before M is assembled and qacc_smooth right after the
    // factorisation — neither the recursion's per-body quantities nor the factor are live while the rows are built
    if constexpr (FUSED) build_vel<T, SYNC, true>(m, ps, *reinterpret_cast<const double (*)[NV]>(qv), w);
    // ---- mj_crb + factor
    Inert crb[NB];
    sfor<1, NB>([&](auto bb) { crb[IDX(bb)] = cin[IDX(bb)]; });
    sfor_down<NB - 1, 1>([&](auto bb) {
        constexpr int b = IDX(bb), p = T::body_parent(b);
        if constexpr (p > 0) crb[p] = crb[p] + crb[b];
    });
    sfor<0, NV>([&](auto ii) {
        constexpr int i = IDX(ii);
        const S6 buf = mul(crb[T::dof_body(i)], cdof[i]);
        sfor<0, i + 1>([&](auto jj) {
            constexpr int j = IDX(jj);
            if constexpr (dof_is_ancestor<T>(j, i)) w.M[tri(i, j)] = dot(cdof[j], buf);
            else w.M[tri(i, j)] = 0;
        });
        w.M[tri(i, i)] += m.dof_armature[i];
    });
    chol_packed<NV>(w.M, w.L);
    if constexpr (FUSED) finish_smooth<T>(m, *reinterpret_cast<const double (*)[nz(T::NU)]>(uu), w);

    stage_sync();
    // ---- constraint rows: joint limits, then contacts
    static_assert(!FUSED || W::RowsT::private_rows, "the fused stages finish aref per rollout: rows must be private");
    const int cap = w.rows.capacity();
    const bool wr = w.rows.row_writer();
    w.overflow = 0;
    int ne = 0;
    sfor<0, NJ>([&](auto jj) {
        constexpr int j = IDX(jj);
        if constexpr (T::jnt_limited(j) && T::jnt_type(j) != ILQG_JNT_FREE) {
            constexpr int da = T::jnt_dofadr(j);
            double value = q[T::jnt_qposadr(j)];
            sfor<0, 2>([&](auto ss) {
                constexpr int side = 2 * IDX(ss) - 1;
                double dist = side * (m.jnt_range[j][IDX(ss)] - value);
                if (dist < m.jnt_margin[j]) {
                    double R, kt;
                    const double B = m.jnt_B[j];
                    row_params(m.jnt_K[j], m.jnt_imp[j], m.jnt_solimp[j], dist, m.jnt_margin[j], m.dof_invw[da], R, kt);
                    if (ne < cap) {
                        if (wr) {
                            sfor<0, NV>([&](auto ii) { w.rows.J(ne, IDX(ii)) = IDX(ii) == da ? -side : 0.0; });
                            w.rows.D(ne) = 1.0 / R;
                            if constexpr (FUSED) w.rows.aref(ne) = -B * (-side * qv[da]) - kt;
                            else { w.rows.rB(ne) = B; w.rows.rkt(ne) = kt; }
                        }
                        ne++;
                    } else
                        w.overflow = 1;
                }
            });
        }
    });
    stage_sync();
    if constexpr (T::NPAIR > 0) {
        // Phase B: rows of every contact found
#pragma unroll 1
        for (int c = 0; c < nc; c++) {
            const int p = cpair[c] & 255;
            const bool has_hint = (cpair[c] & 256) != 0;
            const double dist = cdist[c];
            const P3 pos = cpos[c], n = cnrm[c];
            // the pair's constants, selected from compile-time-indexed tables
            unsigned mk1 = 0, mk2 = 0;   // dofs that move body 1 / body 2
            int condim = 1;
            double margin = 0, mu = 0, tran = 0, K = 0, B = 0, imp = 0;
            P3 r1 = {0, 0, 0}, r2 = {0, 0, 0};  // contact point relative to the com of body 1's / body 2's tree
            sfor<0, T::NPAIR>([&](auto pp) {
                constexpr int P = IDX(pp), b1 = T::geom_body(T::pair_g1(P)), b2 = T::geom_body(T::pair_g2(P));
                if (p == P) {
                    unsigned a = 0, b = 0;
                    sfor<0, NV>([&](auto ii) {
                        if constexpr (b1 > 0 && dof_moves_body<T>(IDX(ii), b1)) a |= 1u << IDX(ii);
                        if constexpr (b2 > 0 && dof_moves_body<T>(IDX(ii), b2)) b |= 1u << IDX(ii);
                    });
                    mk1 = a; mk2 = b;
                    condim = T::pair_condim(P);
                    margin = m.pair_margin[P]; mu = m.pair_mu[P]; K = m.pair_K[P]; B = m.pair_B[P]; imp = m.pair_imp[P];
                    tran = m.body_invw[b1] + m.body_invw[b2];
                    if constexpr (b1 > 0) r1 = pos - com[T::body_root(b1)];
                    if constexpr (b2 > 0) r2 = pos - com[T::body_root(b2)];
                }
            });
            V3 ta, tb;   // the tangents are dense even for a planar tree (one of them is the out-of-plane direction)
            make_frame(full(n), full(chint[c]), has_hint, ta, tb);
            double jn[NV], ja[NV], jb[NV];
            sfor<0, NV>([&](auto ii) {
                constexpr int i = IDX(ii);
                const bool m1 = (mk1 >> i) & 1u, m2 = (mk2 >> i) & 1u;
                jn[i] = 0; ja[i] = 0; jb[i] = 0;
                if (m1 != m2) {  // a dof that moves both bodies or neither gives no relative motion
                    P3 jp = cdof[i].v + cross(cdof[i].w, m2 ? r2 : r1);
                    if (!m2) jp = -1.0 * jp;
                    jn[i] = dot(n, jp);
                    if (condim == 3) { ja[i] = dot(ta, jp); jb[i] = dot(tb, jp); }
                }
            });
            if (tran < ILQG_MINVAL) tran = ILQG_MINVAL;
            if (condim == 1) {
                double R, kt;
                row_params(K, imp, m.pair_solimp[p], dist, margin, tran, R, kt);
                if (ne < cap) {
                    if (wr) {
                        sfor<0, NV>([&](auto ii) { w.rows.J(ne, IDX(ii)) = jn[IDX(ii)]; });
                        w.rows.D(ne) = 1.0 / R;
                        if constexpr (FUSED) {
                            double s = 0;
                            sfor<0, NV>([&](auto ii) { s += jn[IDX(ii)] * qv[IDX(ii)]; });
                            w.rows.aref(ne) = -B * s - kt;
                        } else { w.rows.rB(ne) = B; w.rows.rkt(ne) = kt; }
                    }
                    ne++;
                } else
                    w.overflow = 1;
            } else {
                double R0, kt;
                // all four facets share pos/margin; R of the first facet sets the pyramid's regulariser
                row_params(K, imp, m.pair_solimp[p], dist, margin, tran * (1 + mu * mu), R0, kt);
                double Rpy = 2 * mu * mu * R0;
                if (Rpy < ILQG_MINVAL) Rpy = ILQG_MINVAL;
                double Dpy = 1.0 / Rpy;
                if (ne + 4 <= cap) {
                    if (wr) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const double sg = (k % 2) ? -mu : mu;
                            double s = 0;
                            sfor<0, NV>([&](auto ii) {
                                const double Jri = jn[IDX(ii)] + sg * (k < 2 ? ja[IDX(ii)] : jb[IDX(ii)]);
                                w.rows.J(ne + k, IDX(ii)) = Jri;
                                if constexpr (FUSED) s += Jri * qv[IDX(ii)];
                            });
                            w.rows.D(ne + k) = Dpy;
                            if constexpr (FUSED) w.rows.aref(ne + k) = -B * s - kt;
                            else { w.rows.rB(ne + k) = B; w.rows.rkt(ne + k) = kt; }
                        }
                    }
                    ne += 4;
                } else
                    w.overflow = 1;
            }
        }
    }
    w.nefc = ne;
    w.ncon = nc;
    stage_sync();
}

template <class T, bool SYNC, bool FUSED, class W>
DEV void build_vel(const DevModel<T>& m, const PosStage<T>& ps, const double (&qv)[T::NV], W& w) {
    auto stage_sync = [&]() { if constexpr (SYNC) __syncthreads(); };
    constexpr int NB = T::NBODY, NV = T::NV;
    using A = AlgOf<T>;
    using S6 = typename A::ST; using CD = typename A::CD; using Inert = typename A::IT;
    const S6 (&cdof)[NV] = ps.cdof;
    const Inert (&cin)[NB] = ps.cin;
    // ---- mj_comVel + mj_rne(flg_acc=0): bias forces
    S6 cvel[NB], cacc[NB], cfrc[NB];
    CD cdofdot[NV];
    cvel[0] = {{0, 0, 0}, {0, 0, 0}};
    cacc[0] = {{0, 0, 0}, -1.0 * A::ldP(m.gravity)};
    sfor<1, NB>([&](auto bb) {
        constexpr int b = IDX(bb), p = T::body_parent(b);
        S6 cv = cvel[p];
        S6 ca = cacc[p];
        sfor<0, T::body_jntnum(b)>([&](auto jj) {
            constexpr int j = T::body_jntadr(b) + IDX(jj), da = T::jnt_dofadr(j), ty = T::jnt_type(j);
            if constexpr (ty == ILQG_JNT_FREE) {
                sfor<0, 3>([&](auto kk) { cdofdot[da + IDX(kk)] = {{0, 0, 0}, {0, 0, 0}}; cv = cv + qv[da + IDX(kk)] * cdof[da + IDX(kk)]; });
                sfor<3, 6>([&](auto kk) { cdofdot[da + IDX(kk)] = cross_motion(cv, cdof[da + IDX(kk)]); });
                sfor<3, 6>([&](auto kk) { cv = cv + qv[da + IDX(kk)] * cdof[da + IDX(kk)]; });
            } else {
                cdofdot[da] = cross_motion(cv, cdof[da]);
                cv = cv + qv[da] * cdof[da];
            }
        });
        sfor<T::body_dofadr(b), T::body_dofadr(b) + T::body_dofnum(b)>([&](auto ii) { ca = ca + qv[IDX(ii)] * cdofdot[IDX(ii)]; });
        cvel[b] = cv;
        cacc[b] = ca;
        cfrc[b] = mul(cin[b], ca) + cross_force(cv, mul(cin[b], cv));
    });
    sfor_down<NB - 1, 1>([&](auto bb) {
        constexpr int b = IDX(bb), p = T::body_parent(b);
        if constexpr (p > 0) cfrc[p] = cfrc[p] + cfrc[b];
    });
    // qfrc_passive - qfrc_bias
    sfor<0, NV>([&](auto ii) {
        constexpr int i = IDX(ii), j = T::dof_jnt(i);
        double f = -dot(cdof[i], cfrc[T::dof_body(i)]);
        if constexpr (T::ANY_DAMPING) f -= m.dof_damping[i] * qv[i];
        if constexpr (T::jnt_type(j) != ILQG_JNT_FREE && T::jnt_hasspring(j)) f -= m.jnt_stiffness[j] * ps.dspr[i];
        w.fb[i] = f;
    });
    // reference accelerations of the rows (mj_referenceConstraint): aref = -B (J qvel) - K imp (pos - margin)
    if constexpr (!FUSED)
    ILQG_ROW_PRAGMA
    for (int r = 0; r < w.nefc; r++) {
        double s = 0;
        sfor<0, NV>([&](auto ii) { s += w.rows.J(r, IDX(ii)) * qv[IDX(ii)]; });
        w.rows.aref(r) = -w.rows.rB(r) * s - w.rows.rkt(r);
    }
    stage_sync();
}

template <class T, class W>
DEV void finish_smooth(const DevModel<T>& m, const double (&u)[nz(T::NU)], W& w) {
    constexpr int NV = T::NV;
    sfor<0, NV>([&](auto ii) { w.fs[IDX(ii)] = w.fb[IDX(ii)]; });
    sfor<0, T::NU>([&](auto uu) {
        constexpr int a = IDX(uu);
        double c = u[a];
        if constexpr (T::act_limited(a)) c = clampd(c, m.act_range[a][0], m.act_range[a][1]);
        w.fs[T::act_dof(a)] += m.act_gear[a] * c;
    });
    sfor<0, NV>([&](auto ii) { w.as[IDX(ii)] = w.fs[IDX(ii)]; });
    chol_solve_packed<NV>(w.L, w.as);
}

// the whole of mj_fwdPosition + mj_fwdVelocity + mj_fwdActuation + qacc_smooth for one rollout
template <class T, bool SYNC = false, class W>
DEV void build_problem(const DevModel<T>& m, const double (&q)[T::NQ], const double (&qv)[T::NV], const double (&u)[nz(T::NU)],
                       W& w) {
    PosStage<T> ps;
    build_pos<T, SYNC, true>(m, q, ps, w, qv, u);   // FUSED: includes the velocity stage and qacc_smooth
}

// ------------------------------------------------------------------ constraint solve (mj_fwdConstraint)
// Newton with exact linesearch on the convex piecewise-quadratic cost.  `warm` in: qacc_warmstart;
// out: the solution (which is also the next warm start, as in MuJoCo 2.x).  qacc out.
//
// Termination: MuJoCo's rule (scaled cost improvement or gradient below `tol`, or `maxiter`), plus an exact
// optimality test that makes tol = 0 cheap: when the line minimiser lands in the same quadratic piece the
// Hessian was built for (same active set before and after the step), the step was a full Newton step onto
// that piece's minimiser and the piece is the right one — the point is the global minimum up to the
// round-off of the linear solve, and further iterations (which the reference's tol = 0 setting would
// spend until the cost stops decreasing in fp64) only add rounding noise.
// fac (optional): a cached Newton factor (chol_packed layout) valid for the active set fac_mask — the factor of M + J_A' D_A J_A
// only depends on M, J, D and the active set A, so rollouts that share the position stage (the qvel / ctrl columns of a knot and
// its centre evaluation) share it: an iteration whose active set equals fac_mask skips the Hessian assembly and the factorisation.
// fac_out / fac_mask_out (optional): where this solve leaves the factor of its last iteration and, if it ended exact, that
// iteration's active set (else ~0: no set matches).
template <class T, class W>
DEV void solve(const DevModel<T>& m, W& w, double (&warm)[T::NV], double (&qacc)[T::NV], int maxiter, double tol,
               bool need_forces = false, const double* fac = nullptr, unsigned long long fac_mask = ~0ull, int fac_stride = 1,
               double* fac_out = nullptr, unsigned long long* fac_mask_out = nullptr) {
    constexpr int NV = T::NV, NT = NV * (NV + 1) / 2;
    static_assert(T::MAXEFC <= 64, "active-set masks are 64-bit: larger models use the cooperative kernel");
    typedef unsigned long long mask_t;
    const int ne = w.nefc;
    w.iters = 0;
    w.exact = 1;
    if (ne == 0) {
        sfor<0, NV>([&](auto ii) { qacc[IDX(ii)] = w.as[IDX(ii)]; warm[IDX(ii)] = w.as[IDX(ii)]; w.fc[IDX(ii)] = 0; });
        if (fac_mask_out) *fac_mask_out = ~0ull;
        return;
    }
    w.exact = 0;
    const double scale = 1.0 / (m.meaninertia * (NV > 1 ? NV : 1));
    double Ma[NV], grad[NV], search[NV], Mv[NV];
    {
        // start from the better of the warm start and qacc_smooth; one pass over the rows evaluates both candidates
        // (J lives in local memory: every pass over it is a round of L1/L2 traffic) and leaves jar of the chosen one
        double cw = 0, cs = 0;
        ILQG_ROW_PRAGMA
        for (int r = 0; r < ne; r++) {
            double jw = -w.rows.aref(r), js = jw;
            sfor<0, NV>([&](auto ii) { const double Jri = w.rows.J(r, IDX(ii)); jw += Jri * warm[IDX(ii)]; js += Jri * w.as[IDX(ii)]; });
            const double D = w.rows.D(r);
            if (jw < 0) cw += 0.5 * D * jw * jw;
            if (js < 0) cs += 0.5 * D * js * js;   // the Gauss term of cost(qacc_smooth) is zero
            w.rows.jar(r) = jw;
            w.rows.jv(r) = js;
        }
        sfor<0, NV>([&](auto ii) {
            constexpr int i = IDX(ii);
            double s = 0;
            sfor<0, NV>([&](auto kk) { s += w.M[tri(i, IDX(kk))] * warm[IDX(kk)]; });
            Ma[i] = s;
            cw += 0.5 * (s - w.fs[i]) * (warm[i] - w.as[i]);
        });
        if (cw < cs) sfor<0, NV>([&](auto ii) { qacc[IDX(ii)] = warm[IDX(ii)]; });
        else {
            sfor<0, NV>([&](auto ii) { qacc[IDX(ii)] = w.as[IDX(ii)]; });
            sfor<0, NV>([&](auto ii) {
                constexpr int i = IDX(ii);
                double s = 0;
                sfor<0, NV>([&](auto kk) { s += w.M[tri(i, IDX(kk))] * qacc[IDX(kk)]; });
                Ma[i] = s;
            });
            for (int r = 0; r < ne; r++) w.rows.jar(r) = w.rows.jv(r);
        }
    }
    double cost = 0, old = 0;
    int iter = 0;
    bool forces_current = false;
    mask_t act = 0;
    for (;;) {
        // ---- cost, gradient, Hessian factor and Newton direction at the current point
        act = 0;
        {
            double H[NT], Lh[NT];
            bool cached = false;
            if (fac) {   // the active set first: with the cached factor's set the Hessian is not assembled at all
                mask_t pre = 0;
                for (int r = 0; r < ne; r++) pre |= (mask_t)(w.rows.jar(r) < 0) << r;
                cached = pre == fac_mask;
            }
            if (!cached) sfor<0, NT>([&](auto tt) { H[IDX(tt)] = w.M[IDX(tt)]; });
            sfor<0, NV>([&](auto ii) { w.fc[IDX(ii)] = 0; });
            double c = 0;
            ILQG_HESS_PRAGMA
            for (int r = 0; r < ne; r++) {
                double jar = w.rows.jar(r);
                if (jar < 0) {
                    act |= (mask_t)1 << r;
                    double D = w.rows.D(r);
                    double Jr[NV];
                    sfor<0, NV>([&](auto ii) { Jr[IDX(ii)] = w.rows.J(r, IDX(ii)); });
                    double f = -D * jar;
                    c += 0.5 * D * jar * jar;
                    sfor<0, NV>([&](auto ii) { w.fc[IDX(ii)] += Jr[IDX(ii)] * f; });
                    if (!cached)
                        sfor<0, NV>([&](auto ii) {
                            constexpr int i = IDX(ii);
                            double t = D * Jr[i];
                            sfor<0, i + 1>([&](auto kk) { H[tri(i, IDX(kk))] += t * Jr[IDX(kk)]; });
                        });
                }
            }
            sfor<0, NV>([&](auto ii) {
                constexpr int i = IDX(ii);
                c += 0.5 * (Ma[i] - w.fs[i]) * (qacc[i] - w.as[i]);
                grad[i] = Ma[i] - w.fs[i] - w.fc[i];
                search[i] = grad[i];
            });
            cost = c;
            forces_current = true;
            if (cached) sfor<0, NT>([&](auto tt) { Lh[IDX(tt)] = fac[IDX(tt) * fac_stride]; });
            else {
                chol_packed<NV>(H, Lh);
                if (fac_out) sfor<0, NT>([&](auto tt) { fac_out[IDX(tt)] = Lh[IDX(tt)]; });
            }
            chol_solve_packed<NV>(Lh, search);
            sfor<0, NV>([&](auto ii) { search[IDX(ii)] = -search[IDX(ii)]; });
        }
        if (iter > 0) {
            double gn = 0;
            sfor<0, NV>([&](auto ii) { gn += grad[IDX(ii)] * grad[IDX(ii)]; });
            if (scale * (old - cost) < tol || scale * sqrt(gn) < tol) break;
        }
        if (iter >= maxiter) break;
        // ---- exact linesearch: root of the piecewise-linear derivative along `search`
        double g1 = 0, g2 = 0;
        sfor<0, NV>([&](auto ii) {
            constexpr int i = IDX(ii);
            double s = 0;
            sfor<0, NV>([&](auto kk) { s += w.M[tri(i, IDX(kk))] * search[IDX(kk)]; });
            Mv[i] = s;
        });
        sfor<0, NV>([&](auto ii) { constexpr int i = IDX(ii); g1 += search[i] * (Ma[i] - w.fs[i]); g2 += search[i] * Mv[i]; });
        double d1 = g1, d2 = g2;
        ILQG_ROW_PRAGMA
        for (int r = 0; r < ne; r++) {
            double s = 0;
            sfor<0, NV>([&](auto ii) { s += w.rows.J(r, IDX(ii)) * search[IDX(ii)]; });
            w.rows.jv(r) = s;
            if ((act >> r) & 1) {
                double t = w.rows.D(r) * s;
                d1 += t * w.rows.jar(r);
                d2 += t * s;
            }
        }
        if (d1 >= 0 || d2 < ILQG_MINVAL) break;  // not a descent direction: converged to round-off
        double alpha = 0, lo = 0, hi = CUDART_INF;
        mask_t cur = act;   // active set at `alpha`
        mask_t reached = act;
        for (int it = 0; it < m.ls_iterations; it++) {
            if (d1 < 0) lo = alpha; else hi = alpha;
            double an = alpha - d1 / d2;
            if (!(an > lo && an < hi)) an = isinf(hi) ? 2 * alpha + 1 : 0.5 * (lo + hi);
            double e1 = g1 + g2 * an, e2 = g2;
            mask_t mk = 0;
            ILQG_ROW_PRAGMA
            for (int r = 0; r < ne; r++) {
                double jv = w.rows.jv(r);
                double x = w.rows.jar(r) + an * jv;
                if (x < 0) {
                    double t = w.rows.D(r) * jv;
                    e1 += t * x;
                    e2 += t * jv;
                    mk |= (mask_t)1 << r;
                }
            }
            bool same = mk == cur;  // the step stayed inside one linear piece of the derivative: `an` is its root
            alpha = an;
            d1 = e1;
            d2 = e2;
            cur = mk;
            reached = mk;
            if (same || d1 == 0 || d2 < ILQG_MINVAL) break;
        }
        if (alpha == 0) break;
        sfor<0, NV>([&](auto ii) { constexpr int i = IDX(ii); qacc[i] += alpha * search[i]; Ma[i] += alpha * Mv[i]; });
        const bool exact = reached == act;  // exact optimum (see above)
        // (the rows' residuals at the new point are only read by a further iteration or by the force evaluation below: a solve
        //  that leaves here without either skips this pass over the rows)
        if (!exact || need_forces) {
            ILQG_ROW_PRAGMA
            for (int r = 0; r < ne; r++) w.rows.jar(r) += alpha * w.rows.jv(r);
        }
        old = cost;
        iter++;
        forces_current = false;
        if (exact) { w.exact = 1; break; }
    }
    if (need_forces && !forces_current) {  // qfrc_constraint at the final point (the integrators need it)
        sfor<0, NV>([&](auto ii) { w.fc[IDX(ii)] = 0; });
        ILQG_ROW_PRAGMA
        for (int r = 0; r < ne; r++) {
            double jar = w.rows.jar(r);
            if (jar < 0) {
                double f = -w.rows.D(r) * jar;
                sfor<0, NV>([&](auto ii) { w.fc[IDX(ii)] += w.rows.J(r, IDX(ii)) * f; });
            }
        }
    }
    w.iters = iter;
    if (fac_mask_out) *fac_mask_out = w.exact ? act : ~0ull;   // (an exact exit leaves `act` = the set the last factor was built for)
    sfor<0, NV>([&](auto ii) { warm[IDX(ii)] = qacc[IDX(ii)]; });
}

// ------------------------------------------------------------------ the same solve by a whole warp (small batches)
// A pass over a small batch (one horizon, a single trajectory) is the latency of its longest dependent chain, and the longest link is
// the centre evaluation's COLD solve of a stance knot (mjderivative.cpp:64-68): several Newton iterations, each a string of passes of
// one thread over its rows in local memory.  In the one-launch kernel the other lanes of the knot's warp have nothing to do meanwhile.
// Here they take the rows: the centre lane (`src`) stages its problem in shared memory (rows J | D | aref, then M, qfrc_smooth,
// qacc_smooth, warm start), lane r keeps rows r and r + 32 in registers, and every pass over the rows of solve() becomes one or two
// rows per lane and a butterfly sum: gradient, cost and the 21 entries of the Hessian's row term in one round of reductions, the line
// search two sums and a ballot per trial point.  The dense nv x nv algebra (products with M, Cholesky, triangular solves) is done by
// every lane on identical operands — SIMT executes it once either way, and the xor butterfly leaves bit-identical sums on all lanes, so
// the warp stays converged and every lane ends with the solution: the warm start of the knot's perturbed evaluations
// (mjderivative.cpp:75,91) needs no hand-over.  Same algorithm, same decisions (active sets as ballot masks, exact-optimum exit,
// MuJoCo's termination rule) as solve(); the sums associate differently, so results agree to round-off, not bit for bit.
template <class T>
struct CoopSolveShape {
    static constexpr int NV = T::NV, NT = NV * (NV + 1) / 2, ME = nz(T::MAXEFC), RS = NV + 2;
    static constexpr int NS = (ME + 31) / 32;              // row slots per lane
    static constexpr int DOUBLES = ME * RS + NT + 3 * NV;  // shared memory per warp
};
DEV double warp_allsum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
// Called by all 32 lanes, converged.  `w`, `warm`: only the src lane's are read.  Out, on every lane: qacc = warm = the solution;
// iters / exact as Work::iters / Work::exact of solve(); nact = active rows at the solution.
template <class T, class W>
DEV void solve_coop(const DevModel<T>& m, W& w, bool is_src, int src, double* __restrict__ sh, double (&warm)[T::NV], double (&qacc)[T::NV],
                    int maxiter, double tol, int& iters, int& exact, int& nact) {
    using S = CoopSolveShape<T>;
    constexpr int NV = S::NV, NT = S::NT, NS = S::NS, RS = S::RS;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int ne = __shfl_sync(FULL, w.nefc, src);
    double* shv = sh + S::ME * RS;   // M | fs | as | warm
    __syncwarp();                    // (readers of the previous call's block are done)
    if (is_src) {
        for (int r = 0; r < ne; r++) {
            sfor<0, NV>([&](auto ii) { sh[r * RS + IDX(ii)] = w.rows.J(r, IDX(ii)); });
            sh[r * RS + NV] = w.rows.D(r);
            sh[r * RS + NV + 1] = w.rows.aref(r);
        }
        sfor<0, NT>([&](auto tt) { shv[IDX(tt)] = w.M[IDX(tt)]; });
        sfor<0, NV>([&](auto ii) { shv[NT + IDX(ii)] = w.fs[IDX(ii)]; shv[NT + NV + IDX(ii)] = w.as[IDX(ii)]; shv[NT + 2 * NV + IDX(ii)] = warm[IDX(ii)]; });
    }
    __syncwarp();
    double fs[NV], as[NV];
    sfor<0, NV>([&](auto ii) { fs[IDX(ii)] = shv[NT + IDX(ii)]; as[IDX(ii)] = shv[NT + NV + IDX(ii)]; warm[IDX(ii)] = shv[NT + 2 * NV + IDX(ii)]; });
    iters = 0;
    exact = 1;
    nact = 0;
    if (ne == 0) {
        sfor<0, NV>([&](auto ii) { qacc[IDX(ii)] = as[IDX(ii)]; warm[IDX(ii)] = as[IDX(ii)]; });
        return;
    }
    exact = 0;
    const double scale = 1.0 / (m.meaninertia * (NV > 1 ? NV : 1));
    // this lane's rows; a lane without one holds a row that is never active (J = 0, D = 0, residual +1)
    double Jr[NS][NV], Dr[NS], jar[NS], jv[NS];
    double Ma[NV], grad[NV], search[NV], Mv[NV];
    {
        double cw = 0, cs = 0, jw[NS], js[NS];
        sfor<0, NS>([&](auto ss) {
            constexpr int s = IDX(ss);
            const int r = lane + 32 * s;
            const bool has = r < ne;
            const double* row = sh + (has ? r : 0) * RS;
            sfor<0, NV>([&](auto ii) { Jr[s][IDX(ii)] = has ? row[IDX(ii)] : 0.0; });
            Dr[s] = has ? row[NV] : 0.0;
            double a = has ? -row[NV + 1] : 1.0, b = a;
            sfor<0, NV>([&](auto ii) { a += Jr[s][IDX(ii)] * warm[IDX(ii)]; b += Jr[s][IDX(ii)] * as[IDX(ii)]; });
            jw[s] = a;
            js[s] = b;
            if (a < 0) cw += 0.5 * Dr[s] * a * a;
            if (b < 0) cs += 0.5 * Dr[s] * b * b;
        });
        cw = warp_allsum(cw);
        cs = warp_allsum(cs);
        sfor<0, NV>([&](auto ii) {
            constexpr int i = IDX(ii);
            double s = 0;
            sfor<0, NV>([&](auto kk) { s += shv[tri(i, IDX(kk))] * warm[IDX(kk)]; });
            Ma[i] = s;
            cw += 0.5 * (s - fs[i]) * (warm[i] - as[i]);
        });
        if (cw < cs) {
            sfor<0, NV>([&](auto ii) { qacc[IDX(ii)] = warm[IDX(ii)]; });
            sfor<0, NS>([&](auto ss) { jar[IDX(ss)] = jw[IDX(ss)]; });
        } else {
            sfor<0, NV>([&](auto ii) { qacc[IDX(ii)] = as[IDX(ii)]; });
            sfor<0, NV>([&](auto ii) {
                constexpr int i = IDX(ii);
                double s = 0;
                sfor<0, NV>([&](auto kk) { s += shv[tri(i, IDX(kk))] * qacc[IDX(kk)]; });
                Ma[i] = s;
            });
            sfor<0, NS>([&](auto ss) { jar[IDX(ss)] = js[IDX(ss)]; });
        }
    }
    double cost = 0, old = 0;
    int iter = 0;
    unsigned act[NS];
    for (;;) {
        // ---- cost, gradient, Hessian factor and Newton direction: one round of reductions over [fc | cost | row term of H]
        double red[NV + 1 + NT];
        sfor<0, NV + 1 + NT>([&](auto xx) { red[IDX(xx)] = 0; });
        sfor<0, NS>([&](auto ss) {
            constexpr int s = IDX(ss);
            const bool on = jar[s] < 0;
            act[s] = __ballot_sync(FULL, on);
            if (on) {
                const double D = Dr[s], f = -D * jar[s];
                red[NV] += 0.5 * D * jar[s] * jar[s];
                sfor<0, NV>([&](auto ii) {
                    constexpr int i = IDX(ii);
                    red[i] += Jr[s][i] * f;
                    const double t = D * Jr[s][i];
                    sfor<0, i + 1>([&](auto kk) { red[NV + 1 + tri(i, IDX(kk))] += t * Jr[s][IDX(kk)]; });
                });
            }
        });
        bool any = false;
        sfor<0, NS>([&](auto ss) { any = any || act[IDX(ss)] != 0; });
        if (any) sfor<0, NV + 1 + NT>([&](auto xx) { red[IDX(xx)] = warp_allsum(red[IDX(xx)]); });
        {
            double H[NT], Lh[NT];
            sfor<0, NT>([&](auto tt) { H[IDX(tt)] = shv[IDX(tt)] + red[NV + 1 + IDX(tt)]; });
            double c = red[NV];
            sfor<0, NV>([&](auto ii) {
                constexpr int i = IDX(ii);
                c += 0.5 * (Ma[i] - fs[i]) * (qacc[i] - as[i]);
                grad[i] = Ma[i] - fs[i] - red[i];
                search[i] = grad[i];
            });
            cost = c;
            chol_packed<NV>(H, Lh);
            chol_solve_packed<NV>(Lh, search);
            sfor<0, NV>([&](auto ii) { search[IDX(ii)] = -search[IDX(ii)]; });
        }
        if (iter > 0) {
            double gn = 0;
            sfor<0, NV>([&](auto ii) { gn += grad[IDX(ii)] * grad[IDX(ii)]; });
            if (scale * (old - cost) < tol || scale * sqrt(gn) < tol) break;
        }
        if (iter >= maxiter) break;
        // ---- exact linesearch: root of the piecewise-linear derivative along `search`
        double g1 = 0, g2 = 0;
        sfor<0, NV>([&](auto ii) {
            constexpr int i = IDX(ii);
            double s = 0;
            sfor<0, NV>([&](auto kk) { s += shv[tri(i, IDX(kk))] * search[IDX(kk)]; });
            Mv[i] = s;
        });
        sfor<0, NV>([&](auto ii) { constexpr int i = IDX(ii); g1 += search[i] * (Ma[i] - fs[i]); g2 += search[i] * Mv[i]; });
        double p1 = 0, p2 = 0;
        sfor<0, NS>([&](auto ss) {
            constexpr int s = IDX(ss);
            double x = 0;
            sfor<0, NV>([&](auto ii) { x += Jr[s][IDX(ii)] * search[IDX(ii)]; });
            jv[s] = x;
            if (jar[s] < 0) {
                const double t = Dr[s] * x;
                p1 += t * jar[s];
                p2 += t * x;
            }
        });
        double d1 = g1 + warp_allsum(p1), d2 = g2 + warp_allsum(p2);
        if (d1 >= 0 || d2 < ILQG_MINVAL) break;  // not a descent direction: converged to round-off
        double alpha = 0, lo = 0, hi = CUDART_INF;
        unsigned cur[NS], reached[NS];
        sfor<0, NS>([&](auto ss) { cur[IDX(ss)] = act[IDX(ss)]; reached[IDX(ss)] = act[IDX(ss)]; });
        for (int it = 0; it < m.ls_iterations; it++) {
            if (d1 < 0) lo = alpha; else hi = alpha;
            double an = alpha - d1 / d2;
            if (!(an > lo && an < hi)) an = isinf(hi) ? 2 * alpha + 1 : 0.5 * (lo + hi);
            double q1 = 0, q2 = 0;
            bool same = true;
            sfor<0, NS>([&](auto ss) {
                constexpr int s = IDX(ss);
                const double x = jar[s] + an * jv[s];
                const bool on = x < 0;
                if (on) {
                    const double t = Dr[s] * jv[s];
                    q1 += t * x;
                    q2 += t * jv[s];
                }
                const unsigned mk = __ballot_sync(FULL, on);
                same = same && mk == cur[s];
                cur[s] = mk;
                reached[s] = mk;
            });
            alpha = an;
            d1 = g1 + g2 * an + warp_allsum(q1);
            d2 = g2 + warp_allsum(q2);
            if (same || d1 == 0 || d2 < ILQG_MINVAL) break;  // `same`: the step stayed inside one linear piece, `an` is its root
        }
        if (alpha == 0) break;
        sfor<0, NV>([&](auto ii) { constexpr int i = IDX(ii); qacc[i] += alpha * search[i]; Ma[i] += alpha * Mv[i]; });
        bool ex = true;   // exact optimum: the line minimiser lies in the piece the Hessian was built for (see solve())
        sfor<0, NS>([&](auto ss) { ex = ex && reached[IDX(ss)] == act[IDX(ss)]; jar[IDX(ss)] += alpha * jv[IDX(ss)]; });
        old = cost;
        iter++;
        if (ex) { exact = 1; break; }
    }
    iters = iter;
    sfor<0, NS>([&](auto ss) { nact += __popc(__ballot_sync(FULL, jar[IDX(ss)] < 0)); });
    sfor<0, NV>([&](auto ii) { warm[IDX(ii)] = qacc[IDX(ii)]; });
}

// ------------------------------------------------------------------ integration (mj_Euler / mj_RungeKutta)
DEV void quat_integrate(double* quat, V3 vel, double scale) {
    double n = sqrt(dot(vel, vel));
    V3 ax = n < ILQG_MINVAL ? V3{1, 0, 0} : (1.0 / n) * vel;
    if (n < ILQG_MINVAL) n = 0;
    double s, c;
    sincos(0.5 * scale * n, &s, &c);
    Q4 q = qnormalized({quat[0], quat[1], quat[2], quat[3]});
    Q4 r = qmul(q, {c, ax.x * s, ax.y * s, ax.z * s});
    quat[0] = r.w; quat[1] = r.x; quat[2] = r.y; quat[3] = r.z;
}

template <class T>
DEV void integrate_pos(double (&q)[T::NQ], const double (&v)[T::NV], double dt) {
    sfor<0, T::NJNT>([&](auto jj) {
        constexpr int j = IDX(jj), qa = T::jnt_qposadr(j), da = T::jnt_dofadr(j);
        if constexpr (T::jnt_type(j) == ILQG_JNT_FREE) {
            q[qa] += dt * v[da]; q[qa + 1] += dt * v[da + 1]; q[qa + 2] += dt * v[da + 2];
            quat_integrate(&q[qa + 3], {v[da + 3], v[da + 4], v[da + 5]}, dt);
        } else
            q[qa] += dt * v[da];
    });
}

// one mj_step: forward dynamics with the model's solver settings, then the model's integrator
template <class T, class W>
DEV void step(const DevModel<T>& m, W& w, double (&q)[T::NQ], double (&v)[T::NV], const double (&u)[nz(T::NU)],
              double (&warm)[T::NV], double (&qacc)[T::NV]) {
    constexpr int NV = T::NV, NQ = T::NQ, NT = NV * (NV + 1) / 2;
    const double h = m.timestep;
    // ONE copy of the pipeline for both integrators (a rollout thread runs this chain alone on its SM sub-partition: the
    // instruction footprint is what limits it).  RK4: a 4-trip loop, the stage values folded into the combination as they
    // appear, in the tableau's order (1/6) k0 + (1/3) k1 + (1/3) k2 + (1/6) k3.  Euler: one trip, then the implicit-damping update.
    const bool rk4 = m.integrator == ILQG_INT_RK4;
    const int nstage = rk4 ? 4 : 1;
    double q0[NQ], v0[NV], Xp[NV], Fp[NV], dX[NV], dF[NV];
    sfor<0, NQ>([&](auto ii) { q0[IDX(ii)] = q[IDX(ii)]; });
    sfor<0, NV>([&](auto ii) { v0[IDX(ii)] = v[IDX(ii)]; dX[IDX(ii)] = 0; dF[IDX(ii)] = 0; Xp[IDX(ii)] = 0; Fp[IDX(ii)] = 0; });
#pragma unroll 1
    for (int s = 0; s < nstage; s++) {
        if (s > 0) {
            const double a = s == 3 ? 1.0 : 0.5;  // the only non-zero tableau entry of row s-1 sits at column s-1
            double sx[NV];
            sfor<0, NV>([&](auto ii) { sx[IDX(ii)] = a * Xp[IDX(ii)]; });
            sfor<0, NQ>([&](auto ii) { q[IDX(ii)] = q0[IDX(ii)]; });
            integrate_pos<T>(q, sx, h);
            sfor<0, NV>([&](auto ii) { v[IDX(ii)] = v0[IDX(ii)] + h * (a * Fp[IDX(ii)]); });
        }
        build_problem<T>(m, q, v, u, w);
        solve<T>(m, w, warm, qacc, m.iterations, m.tolerance, !rk4);
        const double wgt = (s == 0 || s == 3) ? 1.0 / 6 : 1.0 / 3;
        sfor<0, NV>([&](auto ii) {
            constexpr int i = IDX(ii);
            Xp[i] = v[i]; Fp[i] = qacc[i];
            dX[i] = fma(wgt, Xp[i], dX[i]);
            dF[i] = fma(wgt, Fp[i], dF[i]);
        });
    }
    if (rk4) {
        sfor<0, NQ>([&](auto ii) { q[IDX(ii)] = q0[IDX(ii)]; });
        sfor<0, NV>([&](auto ii) { v[IDX(ii)] = v0[IDX(ii)] + h * dF[IDX(ii)]; });
        integrate_pos<T>(q, dX, h);
    } else {
        double a[NV];
        if constexpr (T::ANY_DAMPING) {
            double A[NT], La[NT];
            sfor<0, NT>([&](auto tt) { A[IDX(tt)] = w.M[IDX(tt)]; });
            sfor<0, NV>([&](auto ii) { constexpr int i = IDX(ii); A[tri(i, i)] += h * m.dof_damping[i]; a[i] = w.fs[i] + w.fc[i]; });
            chol_packed<NV>(A, La);
            chol_solve_packed<NV>(La, a);
        } else
            sfor<0, NV>([&](auto ii) { a[IDX(ii)] = qacc[IDX(ii)]; });
        sfor<0, NV>([&](auto ii) { v[IDX(ii)] += h * a[IDX(ii)]; });
        integrate_pos<T>(q, v, h);
    }
}

}  // namespace ilqg
