// ilqg.cu — sm_100a kernels and the C ABI (include/ilqg_b200.h) of the iLQG hot path.
//
// FD linearisation (replaces /root/reference/src/mjderivative.cpp:43-255), all knots of all trajectories per call:
//   fd_center_kernel  : one thread per knot — the centre evaluation and its warm-up solves (:61-68); its product is the
//                       warm start every perturbed solve begins from (:75).  It also ranks the knot in its work class.
//   large batches     : fd_bin_kernel (work-class ordering of the knots), fd_velctrl_kernel (qvel + ctrl columns on ONE position
//                       stage per thread: the reference's mjSTAGE_POS / mjSTAGE_VEL skips, :92,124) and fd_qpos_kernel (qpos
//                       columns, full pipeline, :178).
//   small batches     : fd_perturb_kernel, one thread per perturbed evaluation of any column in a single launch.
//   Blocks of deriv are written in the reference layout (staged in shared memory for contiguous runs where that pays), to the
//   caller's buffer or to several (peer-GPU) destinations at once.
// Batched iLQR (ilqr.cuh), the warp-per-rollout engine for large trees (coop.cuh), the peer-memory barrier and the C ABI
// (include/ilqg_b200.h) follow.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <cmath>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "dyn.cuh"
#include "ilqr.cuh"
#include "coop.cuh"

namespace ilqg {

// ------------------------------------------------------------------ step cost on the device
template <class T>
DEV double cost_eval(const ilqg_cost& c, const double (&q)[T::NQ], const double (&v)[T::NV], const double (&u)[nz(T::NU)]) {
    // same term order as the host-side evaluation (no FMA contraction: the +eps forward difference
    // divides by 1e-6, so the cost must round like the caller's C++ cost function)
    double g = 0;
    sfor<0, T::NQ>([&](auto ii) { constexpr int i = IDX(ii); g = __dadd_rn(g, __dmul_rn(__dmul_rn(c.q2[i], q[i]), q[i])); g = __dadd_rn(g, __dmul_rn(c.q1[i], q[i])); });
    sfor<0, T::NV>([&](auto ii) { constexpr int i = IDX(ii); g = __dadd_rn(g, __dmul_rn(__dmul_rn(c.v2[i], v[i]), v[i])); g = __dadd_rn(g, __dmul_rn(c.v1[i], v[i])); });
    sfor<0, T::NU>([&](auto ii) { constexpr int i = IDX(ii); g = __dadd_rn(g, __dmul_rn(__dmul_rn(c.u2[i], u[i]), u[i])); g = __dadd_rn(g, __dmul_rn(c.u1[i], u[i])); });
    return g;
}

template <class T>
DEV void load_knot(int k, const double* qpos, const double* qvel, const double* ctrl, double (&q)[T::NQ], double (&v)[T::NV],
                   double (&u)[nz(T::NU)]) {
    sfor<0, T::NQ>([&](auto ii) { q[IDX(ii)] = qpos[(size_t)k * T::NQ + IDX(ii)]; });
    sfor<0, T::NV>([&](auto ii) { v[IDX(ii)] = qvel[(size_t)k * T::NV + IDX(ii)]; });
    sfor<0, T::NU>([&](auto ii) { u[IDX(ii)] = ctrl[(size_t)k * T::NU + IDX(ii)]; });
}

}  // namespace ilqg
#include "gsolve.cuh"   // (uses cost_eval / load_knot above)
namespace ilqg {

// ------------------------------------------------------------------ FD: centre
// Work classes of a batch (large batches only).  The cost of a knot's perturbed evaluations is set by the number of constraint
// rows its solves walk (0 in flight, 4 per contact point in stance, +1 per joint at its limit), so the centre kernel — which
// knows that number — ranks every knot inside the bucket of its row count, heaviest bucket first; fd_bin_kernel turns the
// ranks into a permutation and the column kernels take their knots through it.  Lanes of a warp and warps of a CTA then run
// the same solver trip counts (no lane waits for a neighbour in contact, no CTA for its one stance warp), and the heavy CTAs
// start first.  Placement only: every knot's arithmetic is unchanged.
constexpr int FD_NBUCKET = 64;
struct FdBins {
    int* cnt;            // [FD_NBUCKET] knots per bucket (zeroed before the centre kernel)
    unsigned int* key;   // [nknots] bucket << 24 | rank inside the bucket
    int* perm;           // [nknots] slot -> knot
    double* pos;         // [nknots][pos_record_doubles<T>()] the centre's position-stage products (xfer_pos_stage): with them the
                         // qvel / ctrl column kernel does not run a position stage of its own (mjSTAGE_POS skip across kernels), or NULL
    double* fac;         // [nknots][NT + 1] the centre solution's Newton factor and (as bits) its active set: the qvel / ctrl columns
                         // of the knot share M, J and D with the centre, so their Newton systems with that active set are this matrix
};

// rows of the heaviest knot of the CTA that starts at slot k0 (the centre kernel's work-class key: bucket = FD_NBUCKET - 1 - rows)
DEV int fd_cta_rows(const FdBins& bins, int k0) { return FD_NBUCKET - 1 - (int)(bins.key[bins.perm[k0]] >> 24); }

// (register caps for 12 / 16 resident warps per SM were measured on this kernel after the planar algebra: +16 % time both)
template <class T, bool EXPORT = false>
__global__ void __launch_bounds__(128) fd_center_kernel(const __grid_constant__ DevModel<T> m, int nknots, const double* __restrict__ qpos,
                                                        const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                        const double* __restrict__ warmstart, int niter, int nwarmup,
                                                        double* __restrict__ qacc_center, int* __restrict__ status, const FdBins bins,
                                                        int* __restrict__ diag) {
    // programmatic dependent launch: the single-launch column kernel that follows does not need this kernel's result before
    // its own position / velocity stages are done — let it start now and wait (griddepcontrol.wait) just before its solves
    asm volatile("griddepcontrol.launch_dependents;");
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < nknots;
    const int kk = valid ? k : nknots - 1;   // the tail's idle lanes evaluate a clamped knot, writes masked (warp-wide ranking below)
    double q[T::NQ], v[T::NV], u[nz(T::NU)], warm[T::NV], qacc[T::NV];
    load_knot<T>(kk, qpos, qvel, ctrl, q, v, u);
    sfor<0, T::NV>([&](auto ii) { warm[IDX(ii)] = warmstart ? warmstart[(size_t)kk * T::NV + IDX(ii)] : 0.0; });
    Work<T> w;
    const long long t0 = clock64();
    if constexpr (EXPORT) {   // stage by stage, and the position stage's products go to the knot's record for the qvel / ctrl columns
        PosStage<T> ps;
        build_pos<T, false, false>(m, q, ps, w);
        build_vel<T, false, false>(m, ps, v, w);
        finish_smooth<T>(m, u, w);
        xfer_pos_stage<false, T>(bins.pos + (size_t)kk * pos_record_doubles<T>(), ps, w);
    } else
        build_problem<T>(m, q, v, u, w);
    const long long t1 = clock64();
    int it_first = 0, it_all = 0;
    // The reference repeats the centre solve nwarmup times to polish the warm start (mjderivative.cpp:67-68).  A solve that left
    // through the exact-optimum test (dyn.cuh) already sits on the minimiser: repeating it from there changes nothing but
    // round-off, so the repetitions stop at the first such solve.
    constexpr int NTF = T::NV * (T::NV + 1) / 2;
    double* fout = bins.fac ? bins.fac + (size_t)kk * (NTF + 1) : nullptr;   // (idle lanes rewrite the clamped knot's block with the same values)
#pragma unroll 1
    for (int rep = 0; rep < nwarmup; rep++) {   // one copy of the solver (instruction footprint)
        solve<T>(m, w, warm, qacc, niter, 0.0, diag != nullptr, nullptr, ~0ull, 1, fout, fout ? reinterpret_cast<unsigned long long*>(fout + NTF) : nullptr);
        if (rep == 0) it_first = w.iters;
        it_all += w.iters;
        if (w.exact) break;
    }
    if (diag && valid) {   // ILQG_DIAG_* (include/ilqg_b200.h)
        const long long t2 = clock64();
        int na = 0;
        for (int r = 0; r < w.nefc; r++) na += w.rows.jar(r) < 0;
        int4* d = reinterpret_cast<int4*>(diag + (size_t)k * ILQG_DIAG_INTS);
        d[0] = make_int4(w.nefc, it_first, it_all, na);
        d[1] = make_int4((int)(t1 - t0), (int)(t2 - t1), w.ncon, 0);
    }
    if (bins.key) {
        const int ne = w.nefc < FD_NBUCKET - 1 ? w.nefc : FD_NBUCKET - 1;
        const int b = valid ? FD_NBUCKET - 1 - ne : FD_NBUCKET;   // heaviest first; idle lanes form a group of their own
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        int base = 0;
        if (valid && lane == leader) base = atomicAdd(&bins.cnt[b], __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid) bins.key[k] = ((unsigned)b << 24) | (unsigned)(base + __popc(peers & ((1u << lane) - 1)));
    }
    if (!valid) return;
    bool ok = true;
    sfor<0, T::NV>([&](auto ii) { qacc_center[(size_t)k * T::NV + IDX(ii)] = qacc[IDX(ii)]; ok = ok && isfinite(qacc[IDX(ii)]); });
    if (status) status[k] = ok ? 0 : ILQG_ERR_NONFINITE;
}

__global__ void __launch_bounds__(256) fd_bin_kernel(int nknots, const FdBins bins) {
    __shared__ int off[FD_NBUCKET];
    if (threadIdx.x < FD_NBUCKET) {
        int s = 0;
        for (int b = 0; b < (int)threadIdx.x; b++) s += bins.cnt[b];
        off[threadIdx.x] = s;
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nknots) return;
    const unsigned key = bins.key[k];
    bins.perm[off[key >> 24] + (int)(key & 0xffffffu)] = k;
}

// ------------------------------------------------------------------ FD: perturbed evaluations
// (FdDst — the destinations of the deriv blocks, one or several — is declared in dyn.cuh)

template <class T>
struct FdShape {
    static constexpr int NV = T::NV, NU = T::NU;
    static constexpr int NCOL = 2 * NV + NU;    // columns per knot: ctrl, qvel, qpos
    static constexpr int G = 2 * NCOL;          // lanes per knot
    static constexpr int KPW = 32 / G;          // knots per warp
    static constexpr int ND = NV * NCOL + NCOL; // deriv doubles per knot
    static constexpr int NJAC = NV * NCOL;
    static_assert(G <= 32, "thread-per-rollout FD kernel needs 2(2nv+nu) <= 32; larger models use the cooperative kernel");
};

template <class T>
__global__ void __launch_bounds__(256, 1) fd_perturb_kernel(const __grid_constant__ DevModel<T> m, int nknots,
                                                                 const double* __restrict__ qpos, const double* __restrict__ qvel,
                                                                 const double* __restrict__ ctrl, const double* __restrict__ qacc_center,
                                                                 const ilqg_cost* __restrict__ cost, double eps, int niter,
                                                                 const FdDst dst, int* __restrict__ status) {
    using S = FdShape<T>;
    constexpr int NV = T::NV, NU = T::NU, NQ = T::NQ, WARPS = 8;
    constexpr bool SYNC = true;   // stage barriers: the CTA's warps share instruction fetches
    __shared__ double stage[WARPS][S::KPW * S::ND];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * WARPS + wib;
    const int sub = lane / S::G, l = lane - sub * S::G;
    const int k = warp * S::KPW + sub;
    const bool valid = sub < S::KPW && k < nknots;
    const int col = l >> 1;
    const double se = (l & 1) ? -eps : eps;
    double qacc[NV];
    double dcost = 0;
    {
        // idle lanes (tail of the grid, lanes 30-31 of a hopper warp) evaluate a clamped knot with their writes masked, so
        // that every thread reaches the stage barriers
        const int kk = valid ? k : (k < nknots ? k : nknots - 1);
        double q[NQ], v[NV], u[nz(NU)], warm[NV];
        load_knot<T>(kk, qpos, qvel, ctrl, q, v, u);
        double c0 = 0;
        if (cost) c0 = cost_eval<T>(*cost, q, v, u);
        // perturb this lane's input (ctrl: mjderivative.cpp:85,99; qvel: :117,130; qpos: :164-169,187-192)
        sfor<0, NU>([&](auto ii) { if (col == IDX(ii)) u[IDX(ii)] += se; });
        sfor<0, NV>([&](auto ii) { if (col == NU + IDX(ii)) v[IDX(ii)] += se; });
        sfor<0, NV>([&](auto ii) {
            constexpr int i = IDX(ii), j = T::dof_jnt(i);
            if (col == NU + NV + i) {
                if constexpr (T::jnt_type(j) == ILQG_JNT_FREE && i >= T::jnt_dofadr(j) + 3) {
                    constexpr int a = i - T::jnt_dofadr(j) - 3;
                    quat_integrate(&q[T::jnt_qposadr(j) + 3], V3{a == 0 ? se : 0.0, a == 1 ? se : 0.0, a == 2 ? se : 0.0}, 1.0);
                } else
                    q[T::jnt_qposadr(j) + i - T::jnt_dofadr(j)] += se;
            }
        });
        if (cost && !(l & 1)) dcost = __ddiv_rn(__dsub_rn(cost_eval<T>(*cost, q, v, u), c0), eps);
        Work<T> w;
        build_problem<T, SYNC>(m, q, v, u, w);
        // everything up to here is independent of the centre kernel: launched as its programmatic dependent, this kernel overlaps
        // it and waits here for its completion (a no-op when launched the ordinary way).  Only the warm start comes from there
        // (mjderivative.cpp:75,91) — and the status words, written below.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        sfor<0, NV>([&](auto ii) { warm[IDX(ii)] = __ldcg(&qacc_center[(size_t)kk * NV + IDX(ii)]); });
        solve<T>(m, w, warm, qacc, niter, 0.0);
    }
    // central difference: the '+' lane (even) takes the '-' lane's result
    bool finite = true;
    const double inv2eps = 1.0 / (2 * eps);
    double* st = stage[wib] + (sub < S::KPW ? sub : 0) * S::ND;
    sfor<0, NV>([&](auto jj) {
        constexpr int j = IDX(jj);
        double other = __shfl_xor_sync(0xffffffffu, qacc[j], 1);
        double d = (qacc[j] - other) * inv2eps;
        finite = finite && isfinite(d);
        if (valid && !(l & 1)) {
            // reference layout: block of kind, element i + j*stride
            int off;
            if (col < NU) off = 2 * NV * NV + col + j * NU;
            else if (col < NU + NV) off = NV * NV + (col - NU) + j * NV;
            else off = (col - NU - NV) + j * NV;
            st[off] = d;
        }
    });
    if (valid && !(l & 1)) {
        // cost gradient entries: dg/dqpos[nv], dg/dqvel[nv], dg/dctrl[nu]
        int off;
        if (col < NU) off = S::NJAC + 2 * NV + col;
        else if (col < NU + NV) off = S::NJAC + NV + (col - NU);
        else off = S::NJAC + (col - NU - NV);
        st[off] = dcost;
        if (!finite && status) atomicExch(&status[k], ILQG_ERR_NONFINITE);
    }
    __syncwarp();
    // coalesced write-out: the warp's knots are contiguous in deriv
    const int k0 = warp * S::KPW;
    int nk = nknots - k0;
    if (nk > S::KPW) nk = S::KPW;
    if (nk > 0) {
        const int per = cost ? S::ND : S::NJAC;  // without a device cost the gradient entries stay untouched
        for (int d = 0; d < dst.n; d++) {
            double* out = dst.p[d] + (size_t)k0 * S::ND;
            if (per == S::ND) {
                for (int e = lane; e < nk * S::ND; e += 32) out[e] = stage[wib][e];
            } else {
                for (int kk = 0; kk < nk; kk++)
                    for (int e = lane; e < per; e += 32) out[kk * S::ND + e] = stage[wib][kk * S::ND + e];
            }
        }
    }
}

// ------------------------------------------------------------------ FD of a small batch in ONE launch
// A batch too small to fill the GPU (the T = 1000 horizon of BASELINE configs[4], the 125 knots a rank holds when that horizon is
// sharded over 8 GPUs, a single trajectory of 21 knots) is bound by the latency of the chain centre -> perturbed evaluation, not by
// throughput.  Here the knot's centre evaluation rides on a spare lane next to its G perturbed evaluations: all G + 1 problems are
// built side by side (one pass through the pipeline's instruction stream instead of two kernels' worth), the centre lane runs its
// solves (mjderivative.cpp:64-68) while the others wait, hands the warm start over by shuffle (:75,91), and the perturbed solves
// follow — one launch, no round trip through HBM, no second kernel start.  Same arithmetic per rollout as the two-launch path.
// (Round 1 measured this arrangement and dropped it: with three full centre solves per knot the lone centre lane cost more than the
// second launch.  With the repetitions stopping at the first exact solve the balance is the other way for small batches.)
template <class T>
struct FdFusedShape {
    static constexpr int NV = T::NV, NU = T::NU;
    static constexpr int NCOL = 2 * NV + NU;
    static constexpr int G = 2 * NCOL;          // perturbed evaluations per knot
    static constexpr int GL = G + 1;            // lanes per knot: + the centre
    static constexpr int KPW = 32 / GL;         // knots per warp
    static constexpr int ND = NV * NCOL + NCOL, NJAC = NV * NCOL;
    static constexpr bool OK = GL <= 32;
};

// COOP (a knot per warp, i.e. G + 1 > 16 lanes per knot — the hopper): the centre's solves are done by the whole warp (solve_coop,
// dyn.cuh) and leave the warm start on every lane.
template <class T, bool COOP>
__global__ void __launch_bounds__(256, 1) fd_fused_kernel(const __grid_constant__ DevModel<T> m, int nknots, const double* __restrict__ qpos,
                                                          const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                          const double* __restrict__ warmstart, const ilqg_cost* __restrict__ cost, double eps,
                                                          int niter, int nwarmup, const FdDst dst, double* __restrict__ qacc_center,
                                                          int* __restrict__ status, int* __restrict__ diag) {
    using S = FdFusedShape<T>;
    constexpr int NV = T::NV, NU = T::NU, NQ = T::NQ, WARPS = 8, KPW = S::KPW > 0 ? S::KPW : 1;
    constexpr bool COOPC = COOP && S::KPW == 1 && T::MAXEFC > 0;
    __shared__ double stage[WARPS][KPW * S::ND];
    __shared__ double csh[COOPC ? WARPS : 1][COOPC ? CoopSolveShape<T>::DOUBLES : 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * WARPS + wib;
    const int sub = lane / S::GL, l = lane - sub * S::GL;
    const int k = warp * KPW + sub;
    const bool valid = sub < KPW && k < nknots;
    const bool is_center = l == S::G;
    const int col = l >> 1;                     // the centre lane's "column" (NCOL) matches no input: it evaluates the knot itself
    const double se = (l & 1) ? -eps : eps;
    const int base = (sub < KPW ? sub : 0) * S::GL;   // first lane of this knot's group
    double qacc[NV], warm[NV];
    double dcost = 0;
    int it_first = 0, it_all = 0, nefc = 0, nact = 0, ncon = 0;
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    {
        // idle lanes (tail of the grid, spare lanes of a warp) evaluate a clamped knot with their writes masked, so that every
        // thread reaches the stage barriers
        const int kk = valid ? k : (k < nknots ? k : nknots - 1);
        double q[NQ], v[NV], u[nz(NU)];
        load_knot<T>(kk, qpos, qvel, ctrl, q, v, u);
        sfor<0, NV>([&](auto ii) { warm[IDX(ii)] = (is_center && warmstart) ? warmstart[(size_t)kk * NV + IDX(ii)] : 0.0; });
        double c0 = 0;
        if (cost) c0 = cost_eval<T>(*cost, q, v, u);
        sfor<0, NU>([&](auto ii) { if (col == IDX(ii)) u[IDX(ii)] += se; });
        sfor<0, NV>([&](auto ii) { if (col == NU + IDX(ii)) v[IDX(ii)] += se; });
        sfor<0, NV>([&](auto ii) {
            constexpr int i = IDX(ii), j = T::dof_jnt(i);
            if (col == NU + NV + i) {
                if constexpr (T::jnt_type(j) == ILQG_JNT_FREE && i >= T::jnt_dofadr(j) + 3) {
                    constexpr int a = i - T::jnt_dofadr(j) - 3;
                    quat_integrate(&q[T::jnt_qposadr(j) + 3], V3{a == 0 ? se : 0.0, a == 1 ? se : 0.0, a == 2 ? se : 0.0}, 1.0);
                } else
                    q[T::jnt_qposadr(j) + i - T::jnt_dofadr(j)] += se;
            }
        });
        if (cost && !(l & 1) && !is_center) dcost = __ddiv_rn(__dsub_rn(cost_eval<T>(*cost, q, v, u), c0), eps);
        Work<T> w;
        t0 = clock64();
        build_problem<T, true>(m, q, v, u, w);
        t1 = clock64();
        // phase 0: the centre lanes solve (repetitions stop at the first exact solve, see fd_center_kernel); phase 1: the warm start
        // goes to the knot's other lanes and they solve.  ONE copy of the solver for both (instruction footprint).
        double cq[NV];   // COOP: the centre's solution (every lane has it; the centre lane's result)
#pragma unroll 1
        for (int phase = COOPC ? 1 : 0; phase < 2; phase++) {
            if constexpr (COOPC) {
                // the warp solves the centre's problem together; afterwards warm = the centre's solution on every lane
                int ex = 0;
#pragma unroll 1
                for (int rep = 0; rep < nwarmup; rep++) {
                    int it = 0;
                    solve_coop<T>(m, w, is_center, S::G, csh[wib], warm, cq, niter, 0.0, it, ex, nact);
                    if (rep == 0) it_first = it;
                    it_all += it;
                    if (ex) break;
                }
                t2 = clock64();
            } else if (phase == 1) {
                t2 = clock64();
                sfor<0, NV>([&](auto jj) {
                    const double c = __shfl_sync(0xffffffffu, qacc[IDX(jj)], base + S::G);
                    if (!is_center) warm[IDX(jj)] = c;
                });
            }
            if ((phase == 0) == is_center) {
                const int reps = phase == 0 ? nwarmup : 1;
#pragma unroll 1
                for (int rep = 0; rep < reps; rep++) {
                    solve<T>(m, w, warm, qacc, niter, 0.0, phase == 0 && diag != nullptr);
                    if (phase == 0) { if (rep == 0) it_first = w.iters; it_all += w.iters; }
                    if (w.exact) break;
                }
            }
            __syncwarp();
        }
        t3 = clock64();
        if constexpr (COOPC) { if (is_center) sfor<0, NV>([&](auto ii) { qacc[IDX(ii)] = cq[IDX(ii)]; }); }
        if (is_center) {
            nefc = w.nefc;
            ncon = w.ncon;
            if constexpr (!COOPC) { if (diag) for (int r = 0; r < w.nefc; r++) nact += w.rows.jar(r) < 0; }
        }
    }
    if (valid && is_center) {
        bool ok = true;
        sfor<0, NV>([&](auto ii) { qacc_center[(size_t)k * NV + IDX(ii)] = qacc[IDX(ii)]; ok = ok && isfinite(qacc[IDX(ii)]); });
        if (status) status[k] = ok ? 0 : ILQG_ERR_NONFINITE;
        if (diag) {
            int4* d = reinterpret_cast<int4*>(diag + (size_t)k * ILQG_DIAG_INTS);
            d[0] = make_int4(nefc, it_first, it_all, nact);
            d[1] = make_int4((int)(t1 - t0), (int)(t2 - t1), ncon, (int)(t3 - t2));
        }
    }
    __syncwarp();   // the centre's status word is written before the columns may flag it
    // central difference: the '+' lane (even l) takes the '-' lane's result
    bool finite = true;
    const double inv2eps = 1.0 / (2 * eps);
    double* st = stage[wib] + (sub < KPW ? sub : 0) * S::ND;
    const int partner = is_center ? lane : base + (l ^ 1);
    sfor<0, NV>([&](auto jj) {
        constexpr int j = IDX(jj);
        double other = __shfl_sync(0xffffffffu, qacc[j], partner);
        double d = (qacc[j] - other) * inv2eps;
        finite = finite && isfinite(d);
        if (valid && !(l & 1) && !is_center) {
            int off;   // reference layout: block of kind, element i + j*stride
            if (col < NU) off = 2 * NV * NV + col + j * NU;
            else if (col < NU + NV) off = NV * NV + (col - NU) + j * NV;
            else off = (col - NU - NV) + j * NV;
            st[off] = d;
        }
    });
    if (valid && !(l & 1) && !is_center) {
        int off;   // cost gradient entries: dg/dqpos[nv], dg/dqvel[nv], dg/dctrl[nu]
        if (col < NU) off = S::NJAC + 2 * NV + col;
        else if (col < NU + NV) off = S::NJAC + NV + (col - NU);
        else off = S::NJAC + (col - NU - NV);
        st[off] = dcost;
        if (!finite && status) atomicExch(&status[k], ILQG_ERR_NONFINITE);
    }
    __syncwarp();
    // coalesced write-out: the warp's knots are contiguous in deriv
    const int k0 = warp * KPW;
    int nk = nknots - k0;
    if (nk > KPW) nk = KPW;
    if (nk > 0) {
        const int per = cost ? S::ND : S::NJAC;  // without a device cost the gradient entries stay untouched
        for (int d = 0; d < dst.n; d++) {
            double* out = dst.p[d] + (size_t)k0 * S::ND;
            if (per == S::ND) {
                for (int e = lane; e < nk * S::ND; e += 32) out[e] = stage[wib][e];
            } else {
                for (int kk = 0; kk < nk; kk++)
                    for (int e = lane; e < per; e += 32) out[kk * S::ND + e] = stage[wib][kk * S::ND + e];
            }
        }
    }
}

// ------------------------------------------------------------------ FD with stage skipping: two kernels after the centre
// The reference re-runs only the stages a perturbation can change: mj_forwardSkip(mjSTAGE_VEL) for ctrl columns,
// mj_forwardSkip(mjSTAGE_POS) for qvel columns, everything for qpos columns (mjderivative.cpp:92,124,178).  In SIMT the
// lanes of a warp must walk the same stages, so the columns are regrouped by kind:
//   fd_velctrl_kernel : GK = gcd(nv, nu) threads per knot; each runs the position stage ONCE on the unperturbed qpos and
//                       then loops over its nu/GK ctrl columns (actuation + solve only) and nv/GK qvel columns
//                       (velocity stage + solve), +eps then -eps, all lanes of the CTA in the same iteration kind.
//   fd_qpos_kernel    : one thread per perturbed evaluation of a qpos column (full pipeline), +/- lanes adjacent.
// Per knot that is (GK + 2nv) position stages instead of 2(2nv+nu).  A CTA owns a whole number of knots
// (floor(CTA size / lanes-per-knot)).  The qpos kernel stages its dq segments (36 doubles per hopper knot) in shared memory and
// writes contiguous runs; the qvel/ctrl kernel stores its entries directly (see there).
template <class T, int THREADS_ = 256>
struct FdSplit {
    static constexpr int NV = T::NV, NU = T::NU, NCOL = 2 * NV + NU;
    static constexpr int gcd_(int a, int b) { return b == 0 ? a : gcd_(b, a % b); }
#ifdef ILQG_VU_GK   // (compile-time A/B: threads per knot of the qvel/ctrl kernel; must divide nv and nu)
    static constexpr int GK = (NU > 0 && NV % ILQG_VU_GK == 0 && NU % ILQG_VU_GK == 0) ? ILQG_VU_GK : (NU > 0 ? gcd_(NV, NU) : 1);
#else
    static constexpr int GK = NU > 0 ? gcd_(NV, NU) : 1;
#endif
    static constexpr int CU = NU / GK, CV = NV / GK;      // ctrl / qvel columns per thread
    static constexpr int THREADS = THREADS_;
    static constexpr int KPC_VU = THREADS / GK;           // knots per CTA
    static constexpr int KPC_Q = THREADS / (2 * NV);
    static constexpr int NJAC = NV * NCOL, ND = NJAC + NCOL;
    static constexpr int SEG_Q = NV * NV, STG_Q = SEG_Q + NV;                     // dq | dg/dqpos
    static_assert(2 * NV <= THREADS, "thread-per-rollout FD kernels need 2 nv <= CTA size");
};

// `perm` (slot -> knot, from fd_bin_kernel) or NULL for the identity: which knot a CTA slot works on.  A knot's segment is
// contiguous in deriv either way; with the identity the CTA's segments also follow each other.
template <class T, bool SYNC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fd_velctrl_kernel(const __grid_constant__ DevModel<T> m, int nknots, const double* __restrict__ qpos,
                                                            const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                            const double* __restrict__ qacc_center, const ilqg_cost* __restrict__ cost,
                                                            double eps, int niter, const FdDst dst, int* __restrict__ status,
                                                            const int* __restrict__ perm, const FdBins bins, int skip_lo, int skip_hi) {
    if (skip_hi > skip_lo) {   // CTAs whose heaviest knot has skip_lo < rows <= skip_hi belong to fd_velctrl_shared_kernel
        const int rows = fd_cta_rows(bins, blockIdx.x * FdSplit<T, THREADS>::KPC_VU);
        if (rows > skip_lo && rows <= skip_hi) return;
    }
    // No shared-memory staging of the results here (the qpos kernel has it): each thread stores its own entries of the dv | du
    // blocks straight to deriv.  The stores are few (54 doubles per knot) and fire-and-forget, while the 43 KB a staged CTA would
    // take out of the SM's L1 hold the rollouts' constraint rows (local memory) — measured: 0.324 -> 0.305 ms.
    using S = FdSplit<T, THREADS>;
    constexpr int NV = T::NV, NU = T::NU, NQ = T::NQ, GK = S::GK;
    const int kl = threadIdx.x / GK, g = threadIdx.x - kl * GK;
    const int k0 = blockIdx.x * S::KPC_VU, slot = k0 + kl;
    const bool valid = kl < S::KPC_VU && slot < nknots;
    const int sc = slot < nknots ? slot : nknots - 1;   // idle lanes evaluate a clamped knot with their writes masked (stage barriers)
    const int kk = perm ? perm[sc] : sc;
    double q[NQ], v[NV], u[nz(NU)], center[NV];
    load_knot<T>(kk, qpos, qvel, ctrl, q, v, u);
    sfor<0, NV>([&](auto ii) { center[IDX(ii)] = qacc_center[(size_t)kk * NV + IDX(ii)]; });
    double c0 = 0;
    if (cost) c0 = cost_eval<T>(*cost, q, v, u);
    PosStage<T> ps;
    Work<T> w;
    build_pos<T, SYNC>(m, q, ps, w);
    const double inv2eps = 1.0 / (2 * eps);
    const size_t base = (size_t)kk * S::ND;
    double qplus[NV], dcost = 0;
    bool finite = true;
#pragma unroll 1
    for (int it = 0; it < 2 * (S::CU + S::CV); it++) {
        const int c = it >> 1;
        const bool is_vel = c >= S::CU;                       // uniform over the grid
        const int col = (is_vel ? c - S::CU : c) * GK + g;    // column within its kind
        const double se = (it & 1) ? -eps : eps;
        double vp[NV], up[nz(NU)], warm[NV], qacc[NV];
        sfor<0, NV>([&](auto ii) { vp[IDX(ii)] = v[IDX(ii)] + ((is_vel && col == IDX(ii)) ? se : 0.0); warm[IDX(ii)] = center[IDX(ii)]; });
        sfor<0, NU>([&](auto ii) { up[IDX(ii)] = u[IDX(ii)] + ((!is_vel && col == IDX(ii)) ? se : 0.0); });
        if (cost && !(it & 1)) dcost = __ddiv_rn(__dsub_rn(cost_eval<T>(*cost, q, vp, up), c0), eps);
        if (it == 0 || is_vel) build_vel<T, false>(m, ps, vp, w);   // ctrl columns keep the centre's velocity stage (mjSTAGE_VEL skip)
        finish_smooth<T>(m, up, w);
        solve<T>(m, w, warm, qacc, niter, 0.0);
        // (measured and rejected, round 2: reusing the centre's Newton factor here — FdBins::fac, as fd_velctrl_shared_kernel does —
        //  0.538 -> 0.552 ms, and the second inlined copy of the solver alone costs 0.465 -> 0.538 ms: instruction footprint)
        if (!(it & 1)) {
            sfor<0, NV>([&](auto jj) { qplus[IDX(jj)] = qacc[IDX(jj)]; });
        } else {
            sfor<0, NV>([&](auto jj) {
                constexpr int j = IDX(jj);
                double d = (qplus[j] - qacc[j]) * inv2eps;
                finite = finite && isfinite(d);
                // reference layout: dv block at nv^2 (element i + j nv), du block at 2 nv^2 (element i + j nu)
                const size_t off = base + NV * NV + (is_vel ? col + j * NV : NV * NV + col + j * NU);
                if (valid)
                    for (int dd = 0; dd < dst.n; dd++) dst.p[dd][off] = d;
            });
            if (valid && cost)   // without a device cost the gradient entries stay untouched
                for (int dd = 0; dd < dst.n; dd++) dst.p[dd][base + S::NJAC + NV + (is_vel ? col : NV + col)] = dcost;
        }
    }
    if (valid && !finite && status) atomicExch(&status[kk], ILQG_ERR_NONFINITE);
}

// The same columns with the knot's constraint rows in SHARED memory, one copy per knot.  The GK threads of a knot run the
// position stage on the same qpos: their J, D, B and k-term are identical, and in per-thread local memory a CTA of stance knots
// (85 knots x 3 copies x ~1 KB) overflows the L1 (ncu, round 1: 60 % hit rate, 128 MB of DRAM write-back per launch against
// 46 MB of results, long-scoreboard the top stall).  Here one thread of the group stores them, all read them (a broadcast),
// and only aref / jar / jv — which differ per column — are per thread, also in shared memory.  The centre's Newton factor
// comes along (FdBins::fac): a column's Newton iteration whose active set is the centre solution's skips the Hessian assembly
// and the 6 x 6 factorisation.  `cap` rows fit the carve-out; the batch is ordered heaviest knot first (work classes), so a
// CTA whose first knot has <= cap rows takes this kernel and the others the local-memory one above: both are launched over
// the whole grid and a CTA of the other class leaves at once.
template <class T, int THREADS>
struct FdVuShared {
    using S = FdSplit<T, THREADS>;
    static constexpr int KF = RowsShared<T>::KF, NT = T::NV * (T::NV + 1) / 2, KPC = S::KPC_VU;
    static constexpr size_t bytes(int cap) { return sizeof(double) * ((size_t)cap * (KF * KPC + 3 * THREADS) + (size_t)NT * KPC); }
    static constexpr int max_cap(size_t smem) {
        const long c = ((long)(smem / sizeof(double)) - (long)NT * KPC) / (KF * KPC + 3 * THREADS);
        return c < 0 ? 0 : (c > T::MAXEFC ? T::MAXEFC : (int)c);
    }
};

template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) fd_velctrl_shared_kernel(const __grid_constant__ DevModel<T> m, int nknots, const double* __restrict__ qpos,
                                                                  const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                                  const double* __restrict__ qacc_center, const ilqg_cost* __restrict__ cost,
                                                                  double eps, int niter, const FdDst dst, int* __restrict__ status,
                                                                  const FdBins bins, int cap_lo, int cap) {
    using S = FdSplit<T, THREADS>;
    using P = FdVuShared<T, THREADS>;
    constexpr int NV = T::NV, NU = T::NU, NQ = T::NQ, GK = S::GK, KPC = P::KPC, NT = P::NT;
    extern __shared__ __align__(16) double vu_smem[];
    const int k0 = blockIdx.x * KPC;
    {
        const int rows = fd_cta_rows(bins, k0);
        if (rows <= cap_lo || rows > cap) return;   // another launch's CTA (uniform over the block)
    }
    const int kl = threadIdx.x / GK, g = threadIdx.x - kl * GK;
    const int slot = k0 + kl;
    const bool valid = kl < KPC && slot < nknots;
    const int sc = slot < nknots ? slot : nknots - 1;   // idle lanes evaluate a clamped knot with their writes masked (stage barriers)
    const int kk = bins.perm[sc];
    const int kq = kl < KPC ? kl : KPC - 1;             // the CTA's spare thread reads the last knot's blocks and writes nothing
    double* sh_knot = vu_smem;
    double* sh_mine = sh_knot + (size_t)cap * P::KF * KPC;
    double* sh_fac = sh_mine + (size_t)cap * 3 * THREADS;
    double q[NQ], v[NV], u[nz(NU)], center[NV];
    load_knot<T>(kk, qpos, qvel, ctrl, q, v, u);
    sfor<0, NV>([&](auto ii) { center[IDX(ii)] = qacc_center[(size_t)kk * NV + IDX(ii)]; });
    const double* kfac = bins.fac + (size_t)kk * (NT + 1);
    if (kl < KPC)
        for (int e = g; e < NT; e += GK) sh_fac[e * KPC + kl] = kfac[e];
    const unsigned long long fmask = *reinterpret_cast<const unsigned long long*>(kfac + NT);
    double c0 = 0;
    if (cost) c0 = cost_eval<T>(*cost, q, v, u);
    PosStage<T> ps;
    Work<T, RowsShared<T>> w;
    w.rows.knot = sh_knot + kq;
    w.rows.mine = sh_mine + threadIdx.x;
    w.rows.kstride = KPC;
    w.rows.tstride = THREADS;
    w.rows.cap = cap;
    w.rows.writer = g == 0 && kl < KPC;
    build_pos<T, true>(m, q, ps, w);   // ends with a block barrier: the knot's rows (and the factors) are visible to its threads
    const double inv2eps = 1.0 / (2 * eps);
    const size_t base = (size_t)kk * S::ND;
    double qplus[NV], dcost = 0;
    bool finite = true;
#pragma unroll 1
    for (int it = 0; it < 2 * (S::CU + S::CV); it++) {
        const int c = it >> 1;
        const bool is_vel = c >= S::CU;                       // uniform over the grid
        const int col = (is_vel ? c - S::CU : c) * GK + g;    // column within its kind
        const double se = (it & 1) ? -eps : eps;
        double vp[NV], up[nz(NU)], warm[NV], qacc[NV];
        sfor<0, NV>([&](auto ii) { vp[IDX(ii)] = v[IDX(ii)] + ((is_vel && col == IDX(ii)) ? se : 0.0); warm[IDX(ii)] = center[IDX(ii)]; });
        sfor<0, NU>([&](auto ii) { up[IDX(ii)] = u[IDX(ii)] + ((!is_vel && col == IDX(ii)) ? se : 0.0); });
        if (cost && !(it & 1)) dcost = __ddiv_rn(__dsub_rn(cost_eval<T>(*cost, q, vp, up), c0), eps);
        if (it == 0 || is_vel) build_vel<T, false>(m, ps, vp, w);   // ctrl columns keep the centre's velocity stage (mjSTAGE_VEL skip)
        finish_smooth<T>(m, up, w);
        solve<T>(m, w, warm, qacc, niter, 0.0, false, sh_fac + kq, fmask, KPC);
        if (!(it & 1)) {
            sfor<0, NV>([&](auto jj) { qplus[IDX(jj)] = qacc[IDX(jj)]; });
        } else {
            sfor<0, NV>([&](auto jj) {
                constexpr int j = IDX(jj);
                double d = (qplus[j] - qacc[j]) * inv2eps;
                finite = finite && isfinite(d);
                const size_t off = base + NV * NV + (is_vel ? col + j * NV : NV * NV + col + j * NU);
                if (valid)
                    for (int dd = 0; dd < dst.n; dd++) dst.p[dd][off] = d;
            });
            if (valid && cost)   // without a device cost the gradient entries stay untouched
                for (int dd = 0; dd < dst.n; dd++) dst.p[dd][base + S::NJAC + NV + (is_vel ? col : NV + col)] = dcost;
        }
    }
    if (valid && !finite && status) atomicExch(&status[kk], ILQG_ERR_NONFINITE);
}

// The same columns WITHOUT a position stage of their own: the centre kernel has already run it on this qpos (the reference's
// mjSTAGE_POS skip, mjderivative.cpp:124, carried across kernels) and left its products in the knot's record (FdBins::pos).
// Per knot that is 13 position stages instead of 16, and the kernel loses the stage whose register pressure sets its occupancy.
template <class T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fd_velctrl_loaded_kernel(const __grid_constant__ DevModel<T> m, int nknots, const double* __restrict__ qpos,
                                                                     const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                                     const double* __restrict__ qacc_center, const ilqg_cost* __restrict__ cost,
                                                                     double eps, int niter, const FdDst dst, int* __restrict__ status,
                                                                     const FdBins bins) {
    using S = FdSplit<T, THREADS>;
    constexpr int NV = T::NV, NU = T::NU, NQ = T::NQ, GK = S::GK;
    const int kl = threadIdx.x / GK, g = threadIdx.x - kl * GK;
    const int slot = blockIdx.x * S::KPC_VU + kl;
    if (kl >= S::KPC_VU || slot >= nknots) return;   // no barriers in this kernel
    const int kk = bins.perm ? bins.perm[slot] : slot;
    double q[NQ], v[NV], u[nz(NU)], center[NV];
    load_knot<T>(kk, qpos, qvel, ctrl, q, v, u);
    sfor<0, NV>([&](auto ii) { center[IDX(ii)] = qacc_center[(size_t)kk * NV + IDX(ii)]; });
    double c0 = 0;
    if (cost) c0 = cost_eval<T>(*cost, q, v, u);
    PosStage<T> ps;
    Work<T> w;
    w.nefc = 0;
    xfer_pos_stage<true, T>(bins.pos + (size_t)kk * pos_record_doubles<T>(), ps, w);
    const double inv2eps = 1.0 / (2 * eps);
    const size_t base = (size_t)kk * S::ND;
    double qplus[NV], dcost = 0;
    bool finite = true;
#pragma unroll 1
    for (int it = 0; it < 2 * (S::CU + S::CV); it++) {
        const int c = it >> 1;
        const bool is_vel = c >= S::CU;                       // uniform over the grid
        const int col = (is_vel ? c - S::CU : c) * GK + g;    // column within its kind
        const double se = (it & 1) ? -eps : eps;
        double vp[NV], up[nz(NU)], warm[NV], qacc[NV];
        sfor<0, NV>([&](auto ii) { vp[IDX(ii)] = v[IDX(ii)] + ((is_vel && col == IDX(ii)) ? se : 0.0); warm[IDX(ii)] = center[IDX(ii)]; });
        sfor<0, NU>([&](auto ii) { up[IDX(ii)] = u[IDX(ii)] + ((!is_vel && col == IDX(ii)) ? se : 0.0); });
        if (cost && !(it & 1)) dcost = __ddiv_rn(__dsub_rn(cost_eval<T>(*cost, q, vp, up), c0), eps);
        if (it == 0 || is_vel) build_vel<T, false>(m, ps, vp, w);   // ctrl columns keep the centre's velocity stage (mjSTAGE_VEL skip)
        finish_smooth<T>(m, up, w);
        solve<T>(m, w, warm, qacc, niter, 0.0);
        if (!(it & 1)) {
            sfor<0, NV>([&](auto jj) { qplus[IDX(jj)] = qacc[IDX(jj)]; });
        } else {
            sfor<0, NV>([&](auto jj) {
                constexpr int j = IDX(jj);
                double d = (qplus[j] - qacc[j]) * inv2eps;
                finite = finite && isfinite(d);
                const size_t off = base + NV * NV + (is_vel ? col + j * NV : NV * NV + col + j * NU);
                for (int dd = 0; dd < dst.n; dd++) dst.p[dd][off] = d;
            });
            if (cost)   // without a device cost the gradient entries stay untouched
                for (int dd = 0; dd < dst.n; dd++) dst.p[dd][base + S::NJAC + NV + (is_vel ? col : NV + col)] = dcost;
        }
    }
    if (!finite && status) atomicExch(&status[kk], ILQG_ERR_NONFINITE);
}

// Experiment (ILQG_Q_MIX = W): the batch is ordered heaviest knot first, so the CTAs resident on an SM at the same time are all of one
// work class and the stance phase of the launch overflows every L1 at once.  With W > 0 the launch order alternates waves of W CTAs
// from the heavy end and from the light end of that order (a relabelling of blockIdx: heavy-designated launch indices take
// 0, 1, 2, ... and light-designated ones G - 1, G - 2, ...), so that a stance CTA shares its SM's L1 with a flight CTA.
DEV int fd_mix_block(int i, int G, int W) {
    const int pair = i / (2 * W), r = i - pair * 2 * W;
    return r < W ? pair * W + r : G - 1 - (pair * W + (r - W));
}

template <class T, bool SYNC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fd_qpos_kernel(const __grid_constant__ DevModel<T> m, int nknots, const double* __restrict__ qpos,
                                                         const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                         const double* __restrict__ qacc_center, const ilqg_cost* __restrict__ cost,
                                                         double eps, int niter, const FdDst dst, int* __restrict__ status,
                                                         const int* __restrict__ perm, int mix) {
    using S = FdSplit<T, THREADS>;
    constexpr int NV = T::NV, NU = T::NU, NQ = T::NQ, G = 2 * NV;
    __shared__ double stage[S::KPC_Q * S::STG_Q];
    __shared__ int knot_of[S::KPC_Q];
    const int kl = threadIdx.x / G, l = threadIdx.x - kl * G;
    const int bx = mix > 0 ? fd_mix_block(blockIdx.x, gridDim.x, mix) : blockIdx.x;
    const int k0 = bx * S::KPC_Q, slot = k0 + kl;
    const bool valid = kl < S::KPC_Q && slot < nknots;
    const int sc = slot < nknots ? slot : nknots - 1;
    const int kk = perm ? perm[sc] : sc;
    if (valid && l == 0) knot_of[kl] = kk;
    const int col = l >> 1;
    const double se = (l & 1) ? -eps : eps;
    double qacc[NV], dcost = 0;
    {
        double q[NQ], v[NV], u[nz(NU)], warm[NV];
        load_knot<T>(kk, qpos, qvel, ctrl, q, v, u);
        sfor<0, NV>([&](auto ii) { warm[IDX(ii)] = qacc_center[(size_t)kk * NV + IDX(ii)]; });
        double c0 = 0;
        if (cost) c0 = cost_eval<T>(*cost, q, v, u);
        // tangent-space perturbation of qpos (mjderivative.cpp:152-169,187-192)
        sfor<0, NV>([&](auto ii) {
            constexpr int i = IDX(ii), j = T::dof_jnt(i);
            if (col == i) {
                if constexpr (T::jnt_type(j) == ILQG_JNT_FREE && i >= T::jnt_dofadr(j) + 3) {
                    constexpr int a = i - T::jnt_dofadr(j) - 3;
                    quat_integrate(&q[T::jnt_qposadr(j) + 3], V3{a == 0 ? se : 0.0, a == 1 ? se : 0.0, a == 2 ? se : 0.0}, 1.0);
                } else
                    q[T::jnt_qposadr(j) + i - T::jnt_dofadr(j)] += se;
            }
        });
        if (cost && !(l & 1)) dcost = __ddiv_rn(__dsub_rn(cost_eval<T>(*cost, q, v, u), c0), eps);
        Work<T> w;
        build_problem<T, SYNC>(m, q, v, u, w);
        solve<T>(m, w, warm, qacc, niter, 0.0);
    }
    bool finite = true;
    const double inv2eps = 1.0 / (2 * eps);
    double* st = stage + (kl < S::KPC_Q ? kl : 0) * S::STG_Q;
    sfor<0, NV>([&](auto jj) {
        constexpr int j = IDX(jj);
        double other = __shfl_xor_sync(0xffffffffu, qacc[j], 1);   // G is even: the +/- pair never straddles a warp
        double d = (qacc[j] - other) * inv2eps;
        finite = finite && isfinite(d);
        if (valid && !(l & 1)) st[col + j * NV] = d;
    });
    if (valid && !(l & 1)) {
        st[S::SEG_Q + col] = dcost;
        if (!finite && status) atomicExch(&status[kk], ILQG_ERR_NONFINITE);
    }
    __syncthreads();
    int nk = nknots - k0;
    if (nk > S::KPC_Q) nk = S::KPC_Q;
    for (int e = threadIdx.x; e < nk * S::SEG_Q; e += blockDim.x) {
        const int kn = e / S::SEG_Q, off = e - kn * S::SEG_Q;
        const double val = stage[kn * S::STG_Q + off];
        for (int d = 0; d < dst.n; d++) dst.p[d][(size_t)knot_of[kn] * S::ND + off] = val;
    }
    if (cost)
        for (int e = threadIdx.x; e < nk * NV; e += blockDim.x) {
            const int kn = e / NV, off = e - kn * NV;
            const double val = stage[kn * S::STG_Q + S::SEG_Q + off];
            for (int d = 0; d < dst.n; d++) dst.p[d][(size_t)knot_of[kn] * S::ND + S::NJAC + off] = val;
        }
}

// ------------------------------------------------------------------ forward / step batches
template <class T>
__global__ void __launch_bounds__(128) forward_kernel(const __grid_constant__ DevModel<T> m, int n, const double* __restrict__ qpos,
                                                      const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                      double* __restrict__ warmstart, double* __restrict__ qacc_out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double q[T::NQ], v[T::NV], u[nz(T::NU)], warm[T::NV], qacc[T::NV];
    load_knot<T>(k, qpos, qvel, ctrl, q, v, u);
    sfor<0, T::NV>([&](auto ii) { warm[IDX(ii)] = warmstart ? warmstart[(size_t)k * T::NV + IDX(ii)] : 0.0; });
    Work<T> w;
    build_problem<T>(m, q, v, u, w);
    solve<T>(m, w, warm, qacc, m.iterations, m.tolerance);
    sfor<0, T::NV>([&](auto ii) {
        qacc_out[(size_t)k * T::NV + IDX(ii)] = qacc[IDX(ii)];
        if (warmstart) warmstart[(size_t)k * T::NV + IDX(ii)] = warm[IDX(ii)];
    });
}

template <class T>
__global__ void __launch_bounds__(128) step_kernel(const __grid_constant__ DevModel<T> m, int n, int nsteps, double* __restrict__ qpos,
                                                   double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                   double* __restrict__ warmstart, double* __restrict__ qacc_out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double q[T::NQ], v[T::NV], u[nz(T::NU)], warm[T::NV], qacc[T::NV];
    load_knot<T>(k, qpos, qvel, ctrl, q, v, u);
    sfor<0, T::NV>([&](auto ii) { warm[IDX(ii)] = warmstart ? warmstart[(size_t)k * T::NV + IDX(ii)] : 0.0; qacc[IDX(ii)] = 0; });
    Work<T> w;
    for (int s = 0; s < nsteps; s++) step<T>(m, w, q, v, u, warm, qacc);
    sfor<0, T::NQ>([&](auto ii) { qpos[(size_t)k * T::NQ + IDX(ii)] = q[IDX(ii)]; });
    sfor<0, T::NV>([&](auto ii) {
        qvel[(size_t)k * T::NV + IDX(ii)] = v[IDX(ii)];
        if (warmstart) warmstart[(size_t)k * T::NV + IDX(ii)] = warm[IDX(ii)];
        if (qacc_out) qacc_out[(size_t)k * T::NV + IDX(ii)] = qacc[IDX(ii)];
    });
}

template <class T, bool OK = (T::NQ == T::NV)>
struct IlqrLaunch {
    static cudaError_t rollout(const DevModel<T>&, const IlqrBuffers&, const ilqg_cost*, cudaStream_t, bool = false, double* = nullptr, int* = nullptr) {
        return cudaErrorNotSupported;
    }
    static cudaError_t accept(const IlqrBuffers&, int, double*, int*, cudaStream_t) { return cudaErrorNotSupported; }
    static cudaError_t backward(const IlqrBuffers&, double, cudaStream_t) { return cudaErrorNotSupported; }
};
template <class T>
struct IlqrLaunch<T, true> {
    static bool Jtrace_direct_ok(const IlqrBuffers& b, bool direct) { return direct && b.nalpha == 1; }
    static cudaError_t rollout(const DevModel<T>& dm, const IlqrBuffers& b, const ilqg_cost* cost, cudaStream_t s, bool direct = false,
                               double* Jtrace = nullptr, int* acc_trace = nullptr) {
        int n = b.ninst * b.nalpha;
        // A rollout is one long dependent chain per thread (T steps x integrator stages), latency-bound.  One warp per CTA, so that the
        // warps of a small batch spread over the SMs instead of sharing one SM's L1 four at a time: 1024 hopper problems x 6 alphas
        // (192 warps, rows in local memory) 0.953 -> 0.899 ms per batch iteration; the pendulum (no rows worth mentioning) is indifferent,
        // 0.2160 / 0.2166 ms.  Fewer rollouts per warp (16 / 8 lanes) were measured too: pendulum +1 %, hopper -1 % / -9 %.
        if (Jtrace_direct_ok(b, direct))
            ilqr_rollout_kernel<T, true><<<(n + 31) / 32, 32, 0, s>>>(dm, b, cost, Jtrace, acc_trace);
        else
            ilqr_rollout_kernel<T><<<(n + 31) / 32, 32, 0, s>>>(dm, b, cost);
        return cudaGetLastError();
    }
    static cudaError_t accept(const IlqrBuffers& b, int accept_always, double* Jtrace, int* acc_trace, cudaStream_t s) {
        ilqr_accept_kernel<T><<<(b.ninst + 127) / 128, 128, 0, s>>>(b, accept_always, Jtrace, acc_trace);
        const size_t T1 = (size_t)(b.N + 1) * b.ninst;
        ilqr_commit_kernel<T><<<(unsigned)((T1 + 127) / 128), 128, 0, s>>>(b);
        return cudaGetLastError();
    }
    static cudaError_t backward(const IlqrBuffers& b, double dt, cudaStream_t s) {
        constexpr int NX = 2 * T::NV;
        if constexpr (NX <= 4 && T::NU <= 2) {   // tiny models: one thread per instance, matrices in registers
            ilqr_backward_small_kernel<T::NV, T::NU><<<(b.ninst + 31) / 32, 32, 0, s>>>(b, dt);
            return cudaGetLastError();
        }
        constexpr int LANES = NX * NX >= 64 ? 32 : 8, GROUPS = 128 / LANES;
        size_t smem = sizeof(BackwardSmem<T::NV, T::NU, LANES>) * GROUPS;
        ilqr_backward_kernel<T::NV, T::NU, LANES, GROUPS><<<(b.ninst + GROUPS - 1) / GROUPS, LANES * GROUPS, smem, s>>>(b, dt);
        return cudaGetLastError();
    }
};

// ------------------------------------------------------------------ engines (one per compiled-in topology)
struct Engine {
    int fd_variant = -1;   // -1: chosen per call by batch size (ILQG_FD_VARIANT overrides)
    int fd_bins = 1;       // work-class ordering of the knots in the stage-skipping kernels (ILQG_FD_BINS=0 disables)
    int* fd_diag = nullptr;   // [nknots][ILQG_DIAG_INTS] per-knot diagnostics of the next fd() call (device), or NULL
    int fd_pdl = 1;           // single-launch column kernel as the centre kernel's programmatic dependent (ILQG_FD_PDL=0 disables)
    int ilqr_direct = 1;      // one-launch forward pass in the reference's mode (ILQG_ILQR_DIRECT=0: rollout + accept + commit kernels)
    // one-launch kernel: the centre's solves by the whole warp (solve_coop; ILQG_FD_COOP=1).  Measured on B200 (T = 1000 hopper horizon,
    // tools/prof_fused_diag.py): the centre's solve tail drops from 91 K to 43 K cycles, the one-launch pass from 84.8 to 72.5 us — and the
    // centre kernel + programmatic dependent column kernel stays ahead at 64 us; on a single trajectory (few rows) the staging costs
    // more than the rows' passes (27.0 against 20.9 us).  Off by default.
    int fd_coop = 0;
    // false: fd() keeps engine-owned scratch indexed by the knot's position in the call, so two fd() calls must not overlap
    // on different streams (the host-pointer pipeline then uses ONE compute stream)
    virtual bool fd_calls_may_overlap() const { return true; }
    virtual void set_fused_max(int) {}
    // knots one full wave of the single-launch column kernel covers (SMs x resident CTAs x knots per CTA), 0 = unknown: the block
    // size at which the FD of an iLQR sub-batch is one wave (ilqg_ilqr_s)
    virtual int fd_wave_knots() { return 0; }
    virtual void set_group(int /*max knots*/, int /*lanes per perturbed solve*/) {}
    virtual void set_vu_classes(const char*) {}
    virtual void set_q_minb(int) {}
    virtual void set_q_mix(int) {}
    virtual void set_vu_pos(int) {}
    // ints of scratch fd() wants for `nknots` knots (bucket counters, keys, permutation); 0 = none
    // (`batch`: the size of the whole batch a chunk belongs to — the kernel variant is chosen on it, so that a chunked host
    //  call runs the same kernels, and returns the same bits, as one device call over the batch)
    virtual size_t fd_scratch_ints(int nknots, int batch) const { (void)nknots; (void)batch; return 0; }
    virtual ~Engine() {}
    virtual int fd_launches() const { return 2; }
    virtual const char* name() const = 0;
    virtual cudaError_t fd(int nknots, const double* qpos, const double* qvel, const double* ctrl, const double* warm,
                           const ilqg_cost* cost_dev, const ilqg_fd_opts& o, const FdDst& dst, double* qacc_center, int* status,
                           int* scratch, int batch, cudaStream_t s, cudaEvent_t* ev) = 0;
    virtual cudaError_t forward(int n, const double* qpos, const double* qvel, const double* ctrl, double* warm, double* qacc,
                                cudaStream_t s) = 0;
    virtual cudaError_t step(int n, int nsteps, double* qpos, double* qvel, const double* ctrl, double* warm, double* qacc,
                             cudaStream_t s) = 0;
    // batched iLQR (ilqr.cuh); false when the model's state is not (qpos, qvel) with nq == nv (quirk Q9)
    virtual bool ilqr_supported() const = 0;
    virtual cudaError_t ilqr_rollout(const IlqrBuffers& b, const ilqg_cost* cost_dev, cudaStream_t s) = 0;
    // forward pass of the reference's own mode (one alpha, accepted unconditionally) in ONE launch: rollout straight over the nominal
    // plus the acceptance bookkeeping (ilqr_rollout_kernel<T, true>).  false: not available, use ilqr_rollout + ilqr_accept
    virtual bool ilqr_rollout_direct(const IlqrBuffers&, const ilqg_cost*, double*, int*, cudaStream_t, cudaError_t*) { return false; }
    virtual cudaError_t ilqr_accept(const IlqrBuffers& b, int accept_always, double* Jtrace, int* acc_trace, cudaStream_t s) = 0;
    virtual cudaError_t ilqr_backward(const IlqrBuffers& b, cudaStream_t s) = 0;
};

template <class T>
struct EngineT : Engine {
    DevModel<T> dm;
    const char* name() const override { return T::NAME; }
    int fd_launches() const override { return last_launches; }
    int last_launches = 3;
    // knots from which the stage-skipping split is chosen.  Measured on B200 (hopper, M knots/s, single launch vs split):
    // 16K 58.9 / 48.1, 21.5K 60.5 / 55.8, 28.7K 64.9 / 63.4, 43K 69.8 / 82.9, 86K 76.9 / 105 — the split's qvel/ctrl kernel has a
    // latency floor of 0.18 ms (a stance CTA's 6 sequential solves), the single launch none.
    // (a tree without collision pairs — the inverted pendulum — has no position stage worth sharing: at 86,016 knots the single-launch
    //  column kernel behind the centre's programmatic launch takes 0.054 ms against the split's 0.062)
    // Round 2, back-to-back passes (tools/prof_split_threshold.py; us, centre + single column kernel / split): SURVEY-8d mix (51 % contact)
    // 20,480 knots 347 / 364, 21,504 370 / 378, 24,576 440 / 401; stance-heavy batch 16,384 466 / 464, 21,504 600 / 524, 24,576 671 / 552.  The
    // crossover moves with the share of stance knots; from 21,504 knots (1024 problems x 21: the hopper iLQR workspace of the bench) the split
    // loses 2 % on the mixed batch and wins 13 % on the stance one.
    static constexpr int SPLIT_MIN = T::NPAIR > 0 ? 21504 : (1 << 22);
    // below this the batch cannot fill the GPU and the one-launch kernel (centre on a spare lane of its knot's warp) has the
    // shortest chain; above it the lone centre lane costs throughput (ILQG_FD_FUSED_MAX overrides; measured on B200)
    int fused_max = FdFusedShape<T>::OK ? 64 : 0;
    void set_fused_max(int n) override { fused_max = FdFusedShape<T>::OK ? n : 0; }
    void set_q_minb(int n) override { q_minb = n; }
    void set_q_mix(int n) override { q_mix = n; }
    void set_vu_pos(int n) override { vu_pos = n; }
    void set_vu_classes(const char* e) override {
        vu_nclass = 0;
        while (*e && vu_nclass < 4) {
            const int c = atoi(e);
            if (c > 0) vu_class[vu_nclass++] = c;
            while (*e && *e != ',') e++;
            if (*e == ',') e++;
        }
    }
    // build kernel + group-solve kernel (gsolve.cuh): several lanes per solve, for batches <= group_max knots (ILQG_FD_GROUP_MAX, or
    // ILQG_FD_VARIANT=4; ILQG_FD_GW = lanes per perturbed solve + 100 x resident CTAs per SM the registers are capped for).
    // Measured on B200 and NOT a gain: off by default (group_max = 0).
    int group_max = 0, group_gw = 4, group_minb = 2;
    static constexpr int GROUP_CAP = 8192;   // knots the record buffer is ever sized for (96 KB per hopper knot)
    double* d_rec = nullptr;
    size_t rec_knots = 0;
    void set_group(int n, int gw) override {
        if (n >= 0) group_max = FdRecord<T>::OK ? (n < GROUP_CAP ? n : GROUP_CAP) : 0;
        if (gw % 100 == 2 || gw % 100 == 4) group_gw = gw % 100;   // ILQG_FD_GW = lanes + 100 * (resident CTAs per SM the registers are capped for)
        if (gw >= 100) group_minb = gw / 100;
    }
    ~EngineT() override { if (d_rec) cudaFree(d_rec); }
    bool fd_calls_may_overlap() const override { return group_max == 0 && fd_variant != 4; }   // (the group-solve experiment owns one record buffer)
    int wave_knots = -1;
    int fd_wave_knots() override {
        if (wave_knots < 0) {
            int dev = 0, sms = 0, occ = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fd_perturb_kernel<T>, 256, 0) != cudaSuccess) occ = 0;
            // a column kernel that owns a whole SM per CTA (the hopper: 255 registers) leaves no room for another block's rollout beside
            // it: blocks then only shrink the FD batches (measured: 1024 hopper problems 1.07 M iterations/s in one block, 1.01 M in four)
            wave_knots = occ >= 2 ? sms * occ * 8 * FdShape<T>::KPW : 0;
        }
        return wave_knots;
    }
    int variant_for(int nknots) const {
        if (fd_variant >= 0) {
            if (fd_variant == 1 && !FdFusedShape<T>::OK) return 2;
            if (fd_variant == 4 && (!FdRecord<T>::OK || nknots > GROUP_CAP)) return 2;
            return fd_variant;
        }
        return nknots >= SPLIT_MIN ? 3 : (nknots <= fused_max ? 1 : (nknots <= group_max ? 4 : 2));
    }
    size_t fd_scratch_ints(int nknots, int batch) const override {
        constexpr int NTF = T::NV * (T::NV + 1) / 2;   // + the centre's Newton factor and active set per knot (FdBins::fac)
        if (!(variant_for(batch) >= 3 && fd_bins && nknots < (1 << 24))) return 0;
        return (size_t)FD_NBUCKET + 2 * (size_t)nknots * (NTF + 2) + (vu_pos ? 2 * (size_t)nknots * pos_record_doubles<T>() : 0);
    }
    // CTA shapes of the split kernels (measured on B200 after the planar algebra shrank the per-rollout state):
    //   qvel/ctrl: 256 threads, one CTA per SM at 255 registers (two CTAs at 128 registers spill the rows' neighbours: +70 % time;
    //              192 threads, one CTA per SM — a quarter less local-memory footprint in L1, a quarter fewer warps: +19 %);
    //   qpos     : 192 threads = 16 knots with no idle lane, two CTAs per SM at 168 registers (12 warps per SM: -7 % time;
    //              two 256-thread CTAs at 128 registers: +8 %).
#ifndef ILQG_VU_THREADS   // (compile-time A/B of the CTA shapes: -DILQG_VU_THREADS=... into a second library, loaded through ILQG_LIB)
#define ILQG_VU_THREADS 256
#endif
#ifndef ILQG_VU_MINB
#define ILQG_VU_MINB 1
#endif
#ifndef ILQG_Q_THREADS
#define ILQG_Q_THREADS 192
#endif
#ifndef ILQG_Q_MINBLOCKS
#define ILQG_Q_MINBLOCKS 2
#endif
    static constexpr int VU_THREADS = ILQG_VU_THREADS, VU_MINB = ILQG_VU_MINB, Q_THREADS = ILQG_Q_THREADS, Q_MINB = ILQG_Q_MINBLOCKS;
    // rows per knot the shared-memory qvel/ctrl kernel holds (ILQG_VU_CAP; 0 = local-memory kernel only).  Measured on B200: see DESIGN.md
    // Row-capacity classes of the shared-memory qvel/ctrl kernel, ascending (ILQG_VU_CLASSES="8,16"; "0" = local-memory kernel
    // only): one launch per class with a carve-out sized for it; knots without rows (flight) and knots above the last class
    // run the local-memory kernel, whose CTAs keep the whole L1.
    // MEASURED on B200 (round 2, tools/prof_split.py): not a gain.  Benchmark batch (23 % stance): local-memory kernel
    // (With classes the last bits of a knot at a class boundary are not repeatable from run to run: the CTA that straddles two classes
    //  goes to the kernel of its heaviest knot, and the ranks inside a class follow the centre kernel's execution order.)
    // 0.299 ms, one class of 16 rows 0.516 ms (each class is one more serialised launch with the 0.1 ms latency floor of a stance
    // CTA, and the carve-out takes the L1 that the rest of a rollout's local state lives in); a batch with 91 % of the knots in
    // stance: 0.779 ms local against 0.803-0.826 ms.  Off by default.
    int vu_class[4] = {0, 0, 0, 0};
    int vu_nclass = 0;
    bool vu_attr_set = false;
    int q_minb = Q_MINB;
    int q_mix = 0;   // ILQG_Q_MIX (fd_mix_block)
    // qvel / ctrl columns on the centre's position-stage products (ILQG_VU_POS: 0 = own position stage per thread, 1 = loaded, one
    // CTA per SM, 2 = loaded, two CTAs per SM at 128 registers)
    int vu_pos = 0;
    void launch_split(int nknots, const double* qpos, const double* qvel, const double* ctrl, const ilqg_cost* cost_dev, const ilqg_fd_opts& o,
                      const FdDst& dst, const double* qacc_center, int* status, const FdBins& bins, cudaStream_t s, cudaEvent_t* ev) {
        using PV = FdSplit<T, VU_THREADS>;
        using PQ = FdSplit<T, Q_THREADS>;
        using SH = FdVuShared<T, VU_THREADS>;
        const int* perm = bins.perm;
        const unsigned grid_vu = (nknots + PV::KPC_VU - 1) / PV::KPC_VU;
        if (bins.pos) {
            if (vu_pos == 2)
                fd_velctrl_loaded_kernel<T, VU_THREADS, 2><<<grid_vu, VU_THREADS, 0, s>>>(dm, nknots, qpos, qvel, ctrl, qacc_center, cost_dev, o.eps, o.niter,
                                                                                        dst, status, bins);
            else
                fd_velctrl_loaded_kernel<T, VU_THREADS, 1><<<grid_vu, VU_THREADS, 0, s>>>(dm, nknots, qpos, qvel, ctrl, qacc_center, cost_dev, o.eps, o.niter,
                                                                                        dst, status, bins);
            if (ev) cudaEventRecord(ev[3], s);
            launch_qpos(nknots, qpos, qvel, ctrl, cost_dev, o, dst, qacc_center, status, perm, s);
            return;
        }
        int lo = 0, hi = 0;   // classes cover lo < rows <= hi
        if constexpr (T::MAXEFC > 0) {
            if (bins.key && bins.fac) {
                const int maxc = SH::max_cap(227 * 1024);
                if (!vu_attr_set) {
                    cudaFuncSetAttribute(fd_velctrl_shared_kernel<T, VU_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH::bytes(maxc));
                    vu_attr_set = true;
                }
                for (int c = 0; c < vu_nclass; c++) {
                    const int cap = vu_class[c] < maxc ? vu_class[c] : maxc;
                    if (cap <= hi) continue;
                    fd_velctrl_shared_kernel<T, VU_THREADS><<<grid_vu, VU_THREADS, SH::bytes(cap), s>>>(dm, nknots, qpos, qvel, ctrl, qacc_center, cost_dev,
                                                                                                   o.eps, o.niter, dst, status, bins, hi, cap);
                    hi = cap;
                }
            }
        }
        // everything else (and the whole batch when it is not ordered by work class): rows in local memory
        // (the default L1 / shared-memory split is the best one for this kernel: forcing a larger shared carve-out shrinks the L1 that
        //  holds the rollouts' local-memory rows and costs up to 30 % — measured with cudaFuncAttributePreferredSharedMemoryCarveout)
        fd_velctrl_kernel<T, true, VU_THREADS, VU_MINB><<<grid_vu, VU_THREADS, 0, s>>>(
            dm, nknots, qpos, qvel, ctrl, qacc_center, cost_dev, o.eps, o.niter, dst, status, perm, bins, lo, hi);
        if (ev) cudaEventRecord(ev[3], s);
        launch_qpos(nknots, qpos, qvel, ctrl, cost_dev, o, dst, qacc_center, status, perm, s);
    }
    void launch_qpos(int nknots, const double* qpos, const double* qvel, const double* ctrl, const ilqg_cost* cost_dev, const ilqg_fd_opts& o,
                     const FdDst& dst, const double* qacc_center, int* status, const int* perm, cudaStream_t s) {
        using PQ = FdSplit<T, Q_THREADS>;
        if (q_minb == 1)   // experiment (ILQG_Q_MINB=1): one CTA per SM, no register cap — half the local-memory footprint per SM
            fd_qpos_kernel<T, true, Q_THREADS, 1><<<(nknots + PQ::KPC_Q - 1) / PQ::KPC_Q, Q_THREADS, 0, s>>>(
                dm, nknots, qpos, qvel, ctrl, qacc_center, cost_dev, o.eps, o.niter, dst, status, perm, perm ? q_mix : 0);
        else
        fd_qpos_kernel<T, true, Q_THREADS, Q_MINB><<<(nknots + PQ::KPC_Q - 1) / PQ::KPC_Q, Q_THREADS, 0, s>>>(
            dm, nknots, qpos, qvel, ctrl, qacc_center, cost_dev, o.eps, o.niter, dst, status, perm, perm ? q_mix : 0);
    }
    cudaError_t fd(int nknots, const double* qpos, const double* qvel, const double* ctrl, const double* warm, const ilqg_cost* cost_dev,
                   const ilqg_fd_opts& o, const FdDst& dst, double* qacc_center, int* status, int* scratch, int batch, cudaStream_t s,
                   cudaEvent_t* ev) override {
        using S = FdShape<T>;
        // variant 3 = stage-skipping split (fewest instructions: best once its qvel/ctrl kernel fills the GPU), 2 = one thread per
        // perturbed evaluation in a single launch (30x more threads per knot: lower latency for small batches); fd_variant -1 = by size
        const int variant = variant_for(batch > nknots ? batch : nknots);
        if (nknots <= 0) return cudaSuccess;
        FdBins bins{nullptr, nullptr, nullptr, nullptr, nullptr};
        if (scratch && fd_scratch_ints(nknots, batch > nknots ? batch : nknots)) {
            constexpr int NTF = T::NV * (T::NV + 1) / 2;
            bins.fac = vu_nclass > 0 ? reinterpret_cast<double*>(scratch) : nullptr;   // doubles first (the scratch base is 8-byte aligned)
            int* ints = scratch + 2 * (size_t)nknots * (NTF + 1);
            if (vu_pos) {
                bins.pos = reinterpret_cast<double*>(ints);
                ints += 2 * (size_t)nknots * pos_record_doubles<T>();
            }
            bins.cnt = ints;
            bins.key = (unsigned int*)(ints + FD_NBUCKET);
            bins.perm = ints + FD_NBUCKET + nknots;
            cudaMemsetAsync(bins.cnt, 0, FD_NBUCKET * sizeof(int), s);
        }
        last_launches = variant >= 3 ? (bins.key ? 4 + (T::MAXEFC > 0 ? vu_nclass : 0) : 3) : (variant == 1 ? 1 : 2);
        if (ev) cudaEventRecord(ev[0], s);
        if constexpr (FdFusedShape<T>::OK) {
            if (variant == 1) {   // small batch: one launch
                using F = FdFusedShape<T>;
                const int nw = (nknots + F::KPW - 1) / F::KPW;
                if (ev) { cudaEventRecord(ev[1], s); cudaEventRecord(ev[3], s); }
                if (fd_coop)
                    fd_fused_kernel<T, true><<<(nw + 7) / 8, 256, 0, s>>>(dm, nknots, qpos, qvel, ctrl, warm, cost_dev, o.eps, o.niter, o.nwarmup, dst,
                                                                          qacc_center, status, fd_diag);
                else
                    fd_fused_kernel<T, false><<<(nw + 7) / 8, 256, 0, s>>>(dm, nknots, qpos, qvel, ctrl, warm, cost_dev, o.eps, o.niter, o.nwarmup, dst,
                                                                           qacc_center, status, fd_diag);
                if (ev) cudaEventRecord(ev[2], s);
                return cudaGetLastError();
            }
        }
        if constexpr (FdRecord<T>::OK) {
            if (variant == 4) {   // latency-bound batch: build + group solve
                using R = FdRecord<T>;
                if (rec_knots < (size_t)nknots) {
                    if (d_rec) { cudaDeviceSynchronize(); cudaFree(d_rec); d_rec = nullptr; rec_knots = 0; }
                    size_t want = 1024;
                    while (want < (size_t)nknots) want *= 2;
                    cudaError_t ae = cudaMalloc(&d_rec, want * R::DOUBLES * sizeof(double));
                    if (ae != cudaSuccess) return ae;
                    rec_knots = want;
                }
                last_launches = 2;
                fd_build_kernel<T><<<(nknots + 7) / 8, 256, 0, s>>>(dm, nknots, qpos, qvel, ctrl, cost_dev, o.eps, d_rec);
                if (ev) { cudaEventRecord(ev[1], s); cudaEventRecord(ev[3], s); }
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(nknots);
                cfg.stream = s;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at;
                cfg.numAttrs = (ev || !fd_pdl) ? 0 : 1;
                const double* rc = d_rec;
                const int has_cost = cost_dev != nullptr;
                cudaError_t le = cudaErrorInvalidValue;
                auto go = [&](auto gw, auto mb) {
                    constexpr int GWc = decltype(gw)::value, MB = decltype(mb)::value;
                    cfg.blockDim = dim3(32 * FdSolveShape<T, GWc>::NW);
                    le = cudaLaunchKernelEx(&cfg, fd_solve_kernel<T, GWc, MB>, dm, nknots, rc, warm, has_cost, o.eps, o.niter, o.nwarmup, dst, qacc_center,
                                            status, fd_diag);
                };
                using I2 = std::integral_constant<int, 2>; using I4 = std::integral_constant<int, 4>; using I6 = std::integral_constant<int, 6>;
                using I8 = std::integral_constant<int, 8>; using I12 = std::integral_constant<int, 12>;
                if (group_gw == 2) { if (group_minb <= 4) go(I2{}, I4{}); else if (group_minb <= 8) go(I2{}, I8{}); else go(I2{}, I12{}); }
                else { if (group_minb <= 2) go(I4{}, I2{}); else if (group_minb <= 4) go(I4{}, I4{}); else go(I4{}, I6{}); }
                if (le != cudaSuccess) return le;
                if (ev) cudaEventRecord(ev[2], s);
                return cudaGetLastError();
            }
        }
        // (64- and 32-thread CTAs for finer-grained balancing of this kernel's 2.3 waves: no gain, measured.  Ordering this kernel's
        //  knots by the permutation the previous call on the same batch ended with — stance knots first, sharing warps — takes 21 us
        //  off it (SMs are busy 61 % of this kernel: ncu) but scrambles the ranks inside the buckets, which follow this kernel's
        //  execution order; the column kernels then lose the locality of neighbouring knots and give 17 us back.  Not kept.)
        if (bins.pos)
            fd_center_kernel<T, true><<<(nknots + 127) / 128, 128, 0, s>>>(dm, nknots, qpos, qvel, ctrl, warm, o.niter, o.nwarmup, qacc_center, status, bins, fd_diag);
        else
        fd_center_kernel<T><<<(nknots + 127) / 128, 128, 0, s>>>(dm, nknots, qpos, qvel, ctrl, warm, o.niter, o.nwarmup, qacc_center, status, bins, fd_diag);
        if (bins.key) fd_bin_kernel<<<(nknots + 255) / 256, 256, 0, s>>>(nknots, bins);
        if (ev) cudaEventRecord(ev[1], s);
        if (variant >= 3) {  // stage-skipping split: qvel/ctrl columns, then qpos columns
            launch_split(nknots, qpos, qvel, ctrl, cost_dev, o, dst, qacc_center, status, bins, s, ev);
            if (ev) cudaEventRecord(ev[2], s);
            return cudaGetLastError();
        }
        if (ev) cudaEventRecord(ev[3], s);
        const int nwarps = (nknots + S::KPW - 1) / S::KPW;
        {
            // programmatic dependent launch (see fd_center_kernel): legal only directly behind the centre kernel in the stream, so
            // not when profiling events sit between the two
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((nwarps + 7) / 8);
            cfg.blockDim = dim3(256);
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = (ev || !fd_pdl) ? 0 : 1;
            const double* qc = qacc_center;
            cudaError_t le = cudaLaunchKernelEx(&cfg, fd_perturb_kernel<T>, dm, nknots, qpos, qvel, ctrl, qc, cost_dev, o.eps, o.niter, dst, status);
            if (le != cudaSuccess) return le;
        }
        if (ev) cudaEventRecord(ev[2], s);
        return cudaGetLastError();
    }
    cudaError_t forward(int n, const double* qpos, const double* qvel, const double* ctrl, double* warm, double* qacc,
                        cudaStream_t s) override {
        if (n <= 0) return cudaSuccess;
        forward_kernel<T><<<(n + 127) / 128, 128, 0, s>>>(dm, n, qpos, qvel, ctrl, warm, qacc);
        return cudaGetLastError();
    }
    cudaError_t step(int n, int nsteps, double* qpos, double* qvel, const double* ctrl, double* warm, double* qacc, cudaStream_t s) override {
        if (n <= 0) return cudaSuccess;
        step_kernel<T><<<(n + 127) / 128, 128, 0, s>>>(dm, n, nsteps, qpos, qvel, ctrl, warm, qacc);
        return cudaGetLastError();
    }
    bool ilqr_supported() const override { return T::NQ == T::NV; }
    cudaError_t ilqr_rollout(const IlqrBuffers& b, const ilqg_cost* cost_dev, cudaStream_t s) override {
        return IlqrLaunch<T>::rollout(dm, b, cost_dev, s);
    }
    bool ilqr_rollout_direct(const IlqrBuffers& b, const ilqg_cost* cost_dev, double* Jtrace, int* acc_trace, cudaStream_t s, cudaError_t* e) override {
        if (T::NQ != T::NV || b.nalpha != 1 || !ilqr_direct) return false;
        *e = IlqrLaunch<T>::rollout(dm, b, cost_dev, s, true, Jtrace, acc_trace);
        return true;
    }
    cudaError_t ilqr_accept(const IlqrBuffers& b, int accept_always, double* Jtrace, int* acc_trace, cudaStream_t s) override {
        return IlqrLaunch<T>::accept(b, accept_always, Jtrace, acc_trace, s);
    }
    cudaError_t ilqr_backward(const IlqrBuffers& b, cudaStream_t s) override { return IlqrLaunch<T>::backward(b, dm.timestep, s); }
};

// generic warp-per-rollout engine (coop.cuh): any model of the subset, used when no topology instantiation matches
struct CoopEngine : Engine {
    GModel* d_g = nullptr;
    int cdbl = 0, cfull = 0, pdbl = 0;   // doubles of the core C-state / C-state with the centre's reusable products / private block
    int vc_warps = 1;                // warps per CTA of the qvel/ctrl kernel
    ilqg_model tab;
    double* d_cstate = nullptr;      // [chunk][cdbl]  the centre's position-stage products
    int* d_cand = nullptr;           // [chunk][COOP_MAXCAND + 1]  collision pairs near contact
    int* d_rowbound = nullptr;       // [chunk]  most rows an evaluation within +-eps of the knot can have
    // row-capacity classes of the qpos kernel: 8 / 7 / 6 rollouts per SM for the humanoid (measured on B200, 2048 knots, qpos
    // columns: one class of 72 rows 8.70 ms; classes 40|72 7.72 ms; 56|72 7.80 ms)
    static constexpr int NCAP = 3;
    int cap_class[NCAP] = {40, 56, COOP_MAXEFC};
    int cdbl_c[NCAP] = {0, 0, 0}, pdbl_c[NCAP] = {0, 0, 0};
    int chunk_cap = 0;
    static constexpr int MAX_CHUNK = 8192;   // knots per internal pass (bounds the scratch: 27 KB of C-state per humanoid knot)
    const char* name() const override { return "generic-warp-per-rollout"; }
    bool fd_calls_may_overlap() const override { return false; }   // d_cstate / d_cand / d_rowbound are indexed by position in the call
    int fd_launches() const override { return 2 + NCAP; }
    ~CoopEngine() override { cudaFree(d_g); cudaFree(d_cstate); cudaFree(d_cand); cudaFree(d_rowbound); }
    size_t warp_bytes() const { return (size_t)(cdbl + pdbl) * sizeof(double); }
    size_t center_bytes() const { return (size_t)(cfull + pdbl) * sizeof(double); }
    cudaError_t init(const ilqg_model& m) {
        GModel* hg = new GModel();
        if (!gmodel_from_tables(m, *hg)) { delete hg; return cudaErrorInvalidValue; }
        tab = m;
        cdbl = (int)coop_cstate_doubles(m, false);
        cfull = (int)coop_cstate_doubles(m, true);
        pdbl = (int)coop_priv_doubles(m);
        for (int c = 0; c < NCAP; c++) {
            cdbl_c[c] = (int)coop_cstate_doubles(m, false, cap_class[c], false);   // the qpos kernel's blocks carry no factor of M
            pdbl_c[c] = (int)coop_priv_doubles(m, cap_class[c]);
        }
        cudaError_t e = cudaMalloc(&d_g, sizeof(GModel));
        if (e == cudaSuccess) e = cudaMemcpy(d_g, hg, sizeof(GModel), cudaMemcpyHostToDevice);
        delete hg;
        if (e != cudaSuccess) return e;
        int dev = 0, maxblk = 0, maxsm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&maxblk, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        if (center_bytes() > (size_t)maxblk) return cudaErrorInvalidValue;
        // qvel/ctrl kernel: as many warps as let two CTAs share an SM (each CTA also pays 1 KB of system shared memory)
        const size_t cb = (size_t)cfull * sizeof(double), pb = (size_t)pdbl * sizeof(double);
        const size_t half = (size_t)maxsm / 2 - 1024;
        int w2 = half > cb ? (int)((half - cb) / pb) : 0;
        int w1 = (int)(((size_t)maxblk - cb) / pb);
        vc_warps = w2 >= 4 ? w2 : w1;
        if (vc_warps > 8) vc_warps = 8;   // __launch_bounds__(256, 2) of coop_velctrl_kernel
        if (vc_warps < 1) return cudaErrorInvalidValue;
        const int one = (int)warp_bytes(), vc = (int)(cb + vc_warps * pb);
        if ((e = cudaFuncSetAttribute(coop_center_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)center_bytes())) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(coop_qpos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, one)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(coop_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, one)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(coop_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, one)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(coop_velctrl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vc)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    cudaError_t ensure_scratch(int n) {
        if (n <= chunk_cap) return cudaSuccess;
        cudaFree(d_cstate); cudaFree(d_cand); cudaFree(d_rowbound);
        d_cstate = nullptr; d_cand = nullptr; d_rowbound = nullptr; chunk_cap = 0;
        cudaError_t e = cudaMalloc(&d_cstate, (size_t)n * cfull * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&d_cand, (size_t)n * (COOP_MAXCAND + 1) * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&d_rowbound, (size_t)n * sizeof(int));
        if (e == cudaSuccess) chunk_cap = n;
        return e;
    }
    cudaError_t fd(int nknots, const double* qpos, const double* qvel, const double* ctrl, const double* warm, const ilqg_cost* cost_dev,
                   const ilqg_fd_opts& o, const FdDst& dst, double* qacc_center, int* status, int* /*scratch*/, int /*batch*/, cudaStream_t s,
                   cudaEvent_t* ev) override {
        if (nknots <= 0) return cudaSuccess;
        const int nq = tab.nq, nv = tab.nv, nu = tab.nu, nd = nv * (2 * nv + nu) + 2 * nv + nu;
        const size_t one = warp_bytes(), vc = (size_t)cfull * sizeof(double) + (size_t)vc_warps * pdbl * sizeof(double);
        // a perturbation of eps moves a geom by eps x (lever arm): pairs farther than margin + slack from contact cannot become active
        const double slack = 2000.0 * o.eps;
        cudaError_t e = ensure_scratch(nknots < MAX_CHUNK ? nknots : MAX_CHUNK);
        if (e != cudaSuccess) return e;
        if (ev) cudaEventRecord(ev[0], s);
        for (int lo = 0; lo < nknots; lo += MAX_CHUNK) {
            const int n = nknots - lo < MAX_CHUNK ? nknots - lo : MAX_CHUNK;
            const double *q = qpos + (size_t)lo * nq, *v = qvel + (size_t)lo * nv, *u = ctrl + (size_t)lo * nu, *w = warm ? warm + (size_t)lo * nv : nullptr;
            double* qc = qacc_center + (size_t)lo * nv;
            FdDst dv = dst;   // this pass's knots in every destination
            for (int d = 0; d < dst.n; d++) dv.p[d] = dst.p[d] + (size_t)lo * nd;
            int* st = status ? status + lo : nullptr;
            coop_center_kernel<<<n, 32, center_bytes(), s>>>(d_g, n, q, v, u, w, o.niter, o.nwarmup, slack, cfull, pdbl, qc, st, d_cstate, d_cand,
                                                             d_rowbound, fd_diag ? fd_diag + (size_t)lo * ILQG_DIAG_INTS : nullptr);
            if (ev && lo == 0) cudaEventRecord(ev[1], s);
            coop_velctrl_kernel<<<n, vc_warps * 32, vc, s>>>(d_g, n, d_cstate, cost_dev, o.eps, o.niter, cfull, pdbl, dv, st);
            if (ev && lo == 0) cudaEventRecord(ev[3], s);
            // qpos columns, one launch per row-capacity class of the knots (a warp whose knot belongs to another class leaves at once)
            for (int c = 0; c < NCAP; c++) {
                const size_t bytes = (size_t)(cdbl_c[c] + pdbl_c[c]) * sizeof(double);
                coop_qpos_kernel<<<(unsigned)((long)n * nv), 32, bytes, s>>>(d_g, n, q, v, u, qc, d_cand, cost_dev, o.eps, o.niter, cdbl_c[c], pdbl_c[c],
                                                                            d_rowbound, c ? cap_class[c - 1] : -1, cap_class[c], dv, st);
            }
        }
        if (ev) cudaEventRecord(ev[2], s);
        return cudaGetLastError();
    }
    cudaError_t forward(int n, const double* qpos, const double* qvel, const double* ctrl, double* warm, double* qacc, cudaStream_t s) override {
        if (n <= 0) return cudaSuccess;
        coop_forward_kernel<<<n, 32, warp_bytes(), s>>>(d_g, n, qpos, qvel, ctrl, warm, qacc, cdbl, pdbl);
        return cudaGetLastError();
    }
    cudaError_t step(int n, int nsteps, double* qpos, double* qvel, const double* ctrl, double* warm, double* qacc, cudaStream_t s) override {
        if (n <= 0) return cudaSuccess;
        if (tab.integrator != ILQG_INT_EULER) return cudaErrorNotSupported;
        coop_step_kernel<<<n, 32, warp_bytes(), s>>>(d_g, n, nsteps, qpos, qvel, ctrl, warm, qacc, cdbl, pdbl);
        return cudaGetLastError();
    }
    // Batched iLQR for trees that only this engine runs (the humanoid: nq = 28 != nv = 27).  The reference's ILQR is undefined there
    // (quirk Q9); this is the opt-in tangent-space extension (state x = (q (-) q*, v - v*), corrected A/B layout): rollouts one warp
    // per (instance, alpha), the Riccati sweep one CTA per instance with every matrix in shared memory.  The backward kernel is
    // instantiated per (nv, nu): 27 x 21 here.
    using HumanoidSmem = BackwardSmem<27, 21, 256>;
    bool ilqr_attr_set = false;
    bool ilqr_supported() const override { return tab.nv == 27 && tab.nu == 21 && tab.integrator == ILQG_INT_EULER; }
    cudaError_t ilqr_rollout(const IlqrBuffers& b, const ilqg_cost* cost_dev, cudaStream_t s) override {
        coop_rollout_kernel<<<(unsigned)(b.ninst * b.nalpha), 32, warp_bytes(), s>>>(d_g, b, cost_dev, cdbl, pdbl);
        return cudaGetLastError();
    }
    cudaError_t ilqr_accept(const IlqrBuffers& b, int accept_always, double* Jtrace, int* acc_trace, cudaStream_t s) override {
        ilqr_accept_kernel<Topo_inverted_pendulum><<<(b.ninst + 127) / 128, 128, 0, s>>>(b, accept_always, Jtrace, acc_trace);   // (no model sizes in it)
        const size_t T1 = (size_t)(b.N + 1) * b.ninst;
        ilqr_commit_rt_kernel<<<(unsigned)((T1 + 127) / 128), 128, 0, s>>>(b, tab.nq, tab.nv, tab.nu);
        return cudaGetLastError();
    }
    cudaError_t ilqr_backward(const IlqrBuffers& b, cudaStream_t s) override {
        if (!ilqr_supported() || !b.cdiff) return cudaErrorNotSupported;
        if (!ilqr_attr_set) {
            cudaError_t e = cudaFuncSetAttribute(ilqr_backward_kernel<27, 21, 256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HumanoidSmem));
            if (e != cudaSuccess) return e;
            ilqr_attr_set = true;
        }
        const size_t items = (size_t)b.N * b.ninst;
        coop_cdiff_kernel<<<(unsigned)((items + 3) / 4), 128, 0, s>>>(d_g, b);
        ilqr_backward_kernel<27, 21, 256, 1><<<b.ninst, 256, sizeof(HumanoidSmem), s>>>(b, tab.timestep);
        return cudaGetLastError();
    }
};

static Engine* make_engine(const ilqg_model& m) {
    if (getenv("ILQG_FORCE_GENERIC")) {  // test hook: run a compiled-in topology through the generic engine
        auto c = std::make_unique<CoopEngine>();
        return c->init(m) == cudaSuccess ? c.release() : nullptr;
    }
#define ILQG_TRY(TOPO)                                          \
    {                                                           \
        auto e = std::make_unique<EngineT<TOPO>>();             \
        if (dev_model_from_tables<TOPO>(m, e->dm)) return e.release(); \
    }
    ILQG_FOR_EACH_TOPOLOGY(ILQG_TRY)
#undef ILQG_TRY
    if (getenv("ILQG_NO_GENERIC")) return nullptr;
    auto c = std::make_unique<CoopEngine>();
    if (c->init(m) == cudaSuccess) return c.release();
    return nullptr;
}

// ------------------------------------------------------------------ node-wide barrier through peer memory
// After a rank's FD kernels have stored their blocks into every peer's array, the ranks only need to know that everybody
// is done.  One warp: lane r release-stores this pass's epoch into slot [rank] of rank r's flag array (over NVLink for the
// peers), then acquire-spins on slot [r] of its own array.  Stream order + the system-scope release make the preceding
// kernels' peer stores visible before the flag.  Each rank runs on its own GPU, so all barrier kernels are co-resident;
// the spin is bounded (about 2 s) and reports a timeout instead of hanging the device.
struct PeerFlags {
    int* p[ILQG_MAX_PEERS];
};
__global__ void peer_barrier_kernel(const PeerFlags f, int nranks, int rank, int epoch, int* __restrict__ timed_out) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
    __threadfence_system();
    int* dst = f.p[r] + rank;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
    const int* mine = f.p[rank] + r;
    const long long t0 = clock64();
    for (;;) {
        int seen;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if (seen - epoch >= 0) break;
        if (clock64() - t0 > 4000000000LL) { *timed_out = 1; break; }
        __nanosleep(100);
    }
}

// ------------------------------------------------------------------ fp64 pipe peak (roofline denominator)
// MEASURED_PEAKS.json carries no fp64 figure, so bench.py measures one: register-resident DFMA chains,
// 16 independent accumulators per thread, every SM busy.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double b, double c) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace ilqg

// ------------------------------------------------------------------ iLQR results of a blocked workspace, packed for one copy
struct IlqrSubTable { int nsub; int i0[17]; };
// thread per instance: its first control (knot N of its block, time-major inside the block) and the last `kept` slots of its cost trace
// u*_n = the same control at every knot of a block's nominal ([T][ni][nu]): the state ILQR::ILQR starts from (ilqr.h:69-97)
__global__ void ilqr_broadcast_ctrl_kernel(double* __restrict__ nom_u, const double* __restrict__ ctrl, int T, int per) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per) return;
    const double c = ctrl ? ctrl[e] : 0.0;
    for (int t = 0; t < T; t++) nom_u[(size_t)t * per + e] = c;
}

__global__ void ilqr_pack_first_control_kernel(ilqg::IlqrBuffers b, IlqrSubTable tab, const double* __restrict__ Jtrace, int nu, int kept, int first_slot,
                                               double* __restrict__ u0, double* __restrict__ Jt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.ninst) return;
    int sb = 0;
    while (sb + 1 < tab.nsub && i >= tab.i0[sb + 1]) sb++;
    const size_t i0 = tab.i0[sb], ni = tab.i0[sb + 1] - tab.i0[sb], li = i - i0, T = (size_t)b.N + 1;
    const double* nu_blk = b.nom_u + T * i0 * nu;       // the block's nominal controls [T][ni][nu]
    for (int e = 0; e < nu; e++) u0[(size_t)i * nu + e] = nu_blk[((size_t)b.N * ni + li) * nu + e];
    const double* jt_blk = Jtrace + (size_t)b.trace_cap * i0;   // the block's trace [trace_cap][ni]
    for (int j = 0; j < kept; j++) Jt[(size_t)i * kept + j] = jt_blk[(size_t)((first_slot + j) % b.trace_cap) * ni + li];
}

// ==================================================================== C ABI
#define ILQG_HOST_MAXCHUNKS 32
#define ILQG_HOST_MAXCOMP 8
#include "host_copy.h"   // BounceCrew / CopyPool: the copy threads of the host-pointer FD call

struct ilqg_handle_s {
    void* stage_host = nullptr;      // pinned landing buffer of small packed results (ilqg_ilqr_get_first_control_*_host)
    size_t stage_host_bytes = 0;
    int device = 0;
    ilqg_model model;
    ilqg::Engine* eng = nullptr;
    std::string err;
    // scratch owned by the handle (grown on demand)
    double* d_center = nullptr; size_t center_cap = 0;
    int* d_bins = nullptr; size_t bins_cap = 0;   // work-class scratch of the FD kernels (ints)
    ilqg_cost* d_cost = nullptr;
    int* d_timeout = nullptr;  // set by a peer barrier that gave up waiting
    int* diag = nullptr;       // caller's device array for per-knot diagnostics of the *_dev FD calls (ilqg_fd_set_diag), or NULL
    // staging for the *_host entry points
    void* d_stage = nullptr; size_t stage_cap = 0;
    long launches = 0;  // kernels launched through this handle (bench.py reports it)
    int host_chunks = 0;  // > 0: forced chunk count of the host-pointer FD pipeline
    int host_comp_streams = 4;  // compute streams of that pipeline (ILQG_HOST_COMP)
    bool profiling = false;
    cudaStream_t pipe[4] = {nullptr, nullptr, nullptr, nullptr};  // upload / compute A / download / compute B streams of the *_host FD entry point
    int host_prio = 1;          // ILQG_HOST_PRIO (0: two plain streams, alternating): compute streams of that pipeline form a priority ladder (comp[])
    cudaStream_t comp[ILQG_HOST_MAXCOMP] = {};   // compute stream j has priority (greatest + j): chunk ci runs on comp[ci % ncomp]
    cudaEvent_t pipe_ev[3 * ILQG_HOST_MAXCHUNKS] = {};    // per chunk: uploaded, computed (2c, 2c + 1); downloaded (2 MAXCHUNKS + c)
    int* h_stat = nullptr; size_t hstat_cap = 0;          // pinned landing buffer of the status words
    double* h_bounce = nullptr; size_t bounce_cap = 0;    // pinned mirror of the staging block for PAGEABLE caller buffers (BounceCrew)
    CopyPool* copy_pool = nullptr;                        // its threads (created on first use)
    int host_threads = -1;                                // ILQG_HOST_THREADS: copy threads of that path (-1: by core count, 0: off)
    // opt-in (ilqg_set_host_pinning): large caller buffers of the *_host FD call are page-locked (cudaHostRegister) the first time
    // they are seen and stay so until the handle is destroyed — a calcMJDerivatives caller hands the same malloc'ed deriv array
    // every call, and a pageable destination is copied at a fifth of the pinned rate
    bool pin_host = false;
    struct Reg { const void* p; size_t bytes; };
    std::vector<Reg> regs;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // before centre, after centre, after the last FD kernel, between the two column kernels
};

static thread_local std::string g_create_err;

static int fail(ilqg_handle h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}
static int cuda_fail(ilqg_handle h, cudaError_t e, const char* what) {
    return fail(h, ILQG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(h, call)                                         \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return cuda_fail(h, e_, #call); \
    } while (0)

extern "C" {

void ilqg_fd_opts_default(ilqg_fd_opts* o) {
    if (!o) return;
    o->eps = 1e-6;
    o->niter = 30;
    o->nwarmup = 3;
}

int ilqg_deriv_size(const ilqg_model* m) { return m ? m->nv * (2 * m->nv + m->nu) + 2 * m->nv + m->nu : 0; }

const char* ilqg_last_error(ilqg_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int ilqg_create(const ilqg_model* m, int device, ilqg_handle* out) {
    if (!m || !out) return fail(nullptr, ILQG_ERR_ARG, "null argument");
    *out = nullptr;
    {
        char verr[256] = "";
        const int vrc = ilqg_model_validate(m, verr, sizeof verr);
        if (vrc) return fail(nullptr, vrc, verr);
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, ILQG_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, ILQG_ERR_ARG, "bad device index");
    CU(nullptr, cudaSetDevice(device));
    ilqg::Engine* eng = ilqg::make_engine(*m);
    if (!eng)
        return fail(nullptr, ILQG_ERR_UNSUPPORTED,
                    "no kernel instantiation matches this model's kinematic tree; add it with tools/gen_topology and rebuild");
    auto* h = new ilqg_handle_s();
    h->device = device;
    h->model = *m;
    h->eng = eng;
    if (const char* e = getenv("ILQG_FD_VARIANT")) eng->fd_variant = atoi(e);
    if (const char* e = getenv("ILQG_FD_BINS")) eng->fd_bins = atoi(e);
    if (const char* e = getenv("ILQG_FD_PDL")) eng->fd_pdl = atoi(e);
    if (const char* e = getenv("ILQG_FD_COOP")) eng->fd_coop = atoi(e);
    if (const char* e = getenv("ILQG_ILQR_DIRECT")) eng->ilqr_direct = atoi(e);
    {
        const char *gm = getenv("ILQG_FD_GROUP_MAX"), *gw = getenv("ILQG_FD_GW");
        if (gm || gw) eng->set_group(gm ? atoi(gm) : -1, gw ? atoi(gw) : 0);
    }
    if (const char* e = getenv("ILQG_FD_FUSED_MAX")) eng->set_fused_max(atoi(e));
    if (const char* e = getenv("ILQG_VU_CLASSES")) eng->set_vu_classes(e);
    if (const char* e = getenv("ILQG_Q_MINB")) eng->set_q_minb(atoi(e));
    if (const char* e = getenv("ILQG_Q_MIX")) eng->set_q_mix(atoi(e));
    if (const char* e = getenv("ILQG_VU_POS")) eng->set_vu_pos(atoi(e));
    if (const char* e = getenv("ILQG_HOST_CHUNKS")) h->host_chunks = atoi(e);
    if (const char* e = getenv("ILQG_HOST_COMP")) h->host_comp_streams = atoi(e);
    if (const char* e = getenv("ILQG_HOST_PRIO")) h->host_prio = atoi(e);
    if (const char* e = getenv("ILQG_HOST_THREADS")) h->host_threads = atoi(e);
    if (const char* e = getenv("ILQG_PIN_HOST")) h->pin_host = atoi(e) != 0;
    if (cudaMalloc(&h->d_cost, sizeof(ilqg_cost)) != cudaSuccess) {
        delete eng;
        delete h;
        return fail(nullptr, ILQG_ERR_CUDA, "cudaMalloc failed");
    }
    *out = h;
    return ILQG_OK;
}

int ilqg_destroy(ilqg_handle h) {
    if (!h) return ILQG_OK;
    cudaSetDevice(h->device);
    cudaFree(h->d_center);
    cudaFree(h->d_bins);
    cudaFree(h->d_cost);
    cudaFree(h->d_timeout);
    cudaFree(h->d_stage);
    for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 4; i++) if (h->pipe[i]) cudaStreamDestroy(h->pipe[i]);
    for (int i = 0; i < ILQG_HOST_MAXCOMP; i++) if (h->comp[i]) cudaStreamDestroy(h->comp[i]);
    for (int i = 0; i < 3 * ILQG_HOST_MAXCHUNKS; i++) if (h->pipe_ev[i]) cudaEventDestroy(h->pipe_ev[i]);
    if (h->h_stat) cudaFreeHost(h->h_stat);
    delete h->copy_pool;
    if (h->h_bounce) cudaFreeHost(h->h_bounce);
    if (h->stage_host) cudaFreeHost(h->stage_host);
    for (auto& r : h->regs) cudaHostUnregister(const_cast<void*>(r.p));
    delete h->eng;
    delete h;
    return ILQG_OK;
}

long ilqg_launch_count(ilqg_handle h) { return h ? h->launches : 0; }

int ilqg_set_profiling(ilqg_handle h, int on) {
    if (!h) return ILQG_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    if (on && !h->ev[0])
        for (int i = 0; i < 4; i++) CU(h, cudaEventCreate(&h->ev[i]));
    h->profiling = on != 0;
    return ILQG_OK;
}

int ilqg_fd_last_kernel_ms(ilqg_handle h, float* center_ms, float* perturb_ms) {
    if (!h || !h->ev[0] || !center_ms || !perturb_ms) return ILQG_ERR_ARG;
    CU(h, cudaEventSynchronize(h->ev[2]));
    CU(h, cudaEventElapsedTime(center_ms, h->ev[0], h->ev[1]));
    CU(h, cudaEventElapsedTime(perturb_ms, h->ev[1], h->ev[2]));
    return ILQG_OK;
}
int ilqg_fd_last_stage_ms(ilqg_handle h, float* center_ms, float* velctrl_ms, float* qpos_ms) {
    if (!h || !h->ev[0] || !center_ms || !velctrl_ms || !qpos_ms) return ILQG_ERR_ARG;
    CU(h, cudaEventSynchronize(h->ev[2]));
    CU(h, cudaEventElapsedTime(center_ms, h->ev[0], h->ev[1]));
    CU(h, cudaEventElapsedTime(velctrl_ms, h->ev[1], h->ev[3]));
    CU(h, cudaEventElapsedTime(qpos_ms, h->ev[3], h->ev[2]));
    return ILQG_OK;
}
const char* ilqg_engine_name(ilqg_handle h) { return h && h->eng ? h->eng->name() : ""; }

// diagnostic: sustained fp64 FMA throughput of `device` in TFLOP/s (FMA = 2 flops)
int ilqg_fp64_peak(int device, double* tflops) {
    if (!tflops) return ILQG_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return ILQG_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ILQG_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return ILQG_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        ilqg::fp64_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return ILQG_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 16 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return ILQG_OK;
}

static int ensure_center(ilqg_handle h, size_t n) {
    if (n <= h->center_cap) return ILQG_OK;
    cudaFree(h->d_center);
    h->d_center = nullptr;
    h->center_cap = 0;
    CU(h, cudaMalloc(&h->d_center, n * sizeof(double)));
    h->center_cap = n;
    return ILQG_OK;
}
static int ensure_bins(ilqg_handle h, size_t n) {
    if (n <= h->bins_cap) return ILQG_OK;
    cudaFree(h->d_bins);
    h->d_bins = nullptr;
    h->bins_cap = 0;
    CU(h, cudaMalloc(&h->d_bins, n * sizeof(int)));
    h->bins_cap = n;
    return ILQG_OK;
}
static int ensure_stage(ilqg_handle h, size_t bytes) {
    if (bytes <= h->stage_cap) return ILQG_OK;
    cudaFree(h->d_stage);
    h->d_stage = nullptr;
    h->stage_cap = 0;
    CU(h, cudaMalloc(&h->d_stage, bytes));
    h->stage_cap = bytes;
    return ILQG_OK;
}

// launch the two FD kernels; `dcost` is a DEVICE pointer (or NULL)
// `scratch`: the engine's work-class scratch for this call (fd_scratch_ints(nknots) ints), or NULL to use the handle's own
// (calls that overlap on different streams must bring their own)
static int fd_launch(ilqg_handle h, int nknots, const double* qpos, const double* qvel, const double* ctrl, const double* warmstart,
                     const ilqg_cost* dcost, const ilqg_fd_opts* opts, const ilqg::FdDst& dst, double* qacc_out, int* status, cudaStream_t s,
                     int* scratch = nullptr, int batch = 0, int* diag = nullptr) {
    ilqg_fd_opts o;
    ilqg_fd_opts_default(&o);
    if (opts) o = *opts;
    if (!(o.eps > 0) || o.niter < 0 || o.nwarmup < 1) return fail(h, ILQG_ERR_ARG, "bad FD options");
    double* center = qacc_out;
    if (!center) {
        int rc = ensure_center(h, (size_t)nknots * h->model.nv);
        if (rc) return rc;
        center = h->d_center;
    }
    if (batch < nknots) batch = nknots;
    if (!scratch) {
        const size_t want = h->eng->fd_scratch_ints(nknots, batch);
        if (want) {
            int rc = ensure_bins(h, want);
            if (rc) return rc;
            scratch = h->d_bins;
        }
    }
    h->eng->fd_diag = diag;
    cudaError_t fe = h->eng->fd(nknots, qpos, qvel, ctrl, warmstart, dcost, o, dst, center, status, scratch, batch, s, h->profiling ? h->ev : nullptr);
    h->eng->fd_diag = nullptr;
    CU(h, fe);
    h->launches += h->eng->fd_launches();
    return ILQG_OK;
}

int ilqg_fd_batch_dev(ilqg_handle h, int nknots, const double* qpos, const double* qvel, const double* ctrl, const double* warmstart,
                      const ilqg_cost* cost, const ilqg_fd_opts* opts, double* deriv, double* qacc_out, int* status, void* stream) {
    if (!h) return ILQG_ERR_ARG;
    if (nknots < 0 || (nknots > 0 && (!qpos || !qvel || !deriv || (h->model.nu > 0 && !ctrl)))) return fail(h, ILQG_ERR_ARG, "null buffer");
    if (nknots == 0) return ILQG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, cudaSetDevice(h->device));
    const ilqg_cost* dcost = nullptr;
    if (cost) {  // `cost` is a HOST struct even in the _dev flavour (it is a small parameter block, like opts)
        CU(h, cudaMemcpyAsync(h->d_cost, cost, sizeof(ilqg_cost), cudaMemcpyHostToDevice, s));
        dcost = h->d_cost;
    }
    ilqg::FdDst dst{};
    dst.p[0] = deriv;
    dst.n = 1;
    return fd_launch(h, nknots, qpos, qvel, ctrl, warmstart, dcost, opts, dst, qacc_out, status, s, nullptr, 0, h->diag);
}

// per-knot diagnostics of the centre evaluation (ILQG_DIAG_* in ilqg_b200.h), written by the following *_dev FD calls on this handle
int ilqg_fd_set_diag(ilqg_handle h, int* diag_dev) {
    if (!h) return ILQG_ERR_ARG;
    h->diag = diag_dev;
    return ILQG_OK;
}

// The same linearisation with the deriv blocks stored to `ndst` destinations at once (each dsts[i] points at the slot of knot 0 of
// this call in destination i).  Destinations may live on peer GPUs (ilqg_peer_open): the kernels' write-out then IS the gather.
int ilqg_fd_batch_dev_scatter(ilqg_handle h, int nknots, const double* qpos, const double* qvel, const double* ctrl, const double* warmstart,
                              const ilqg_cost* cost, const ilqg_fd_opts* opts, double* const* dsts, int ndst, double* qacc_out, int* status,
                              void* stream) {
    if (!h) return ILQG_ERR_ARG;
    if (nknots < 0 || ndst < 1 || ndst > ILQG_MAX_PEERS || !dsts || (nknots > 0 && (!qpos || !qvel || (h->model.nu > 0 && !ctrl))))
        return fail(h, ILQG_ERR_ARG, "bad argument");
    if (nknots == 0) return ILQG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, cudaSetDevice(h->device));
    const ilqg_cost* dcost = nullptr;
    if (cost) {
        CU(h, cudaMemcpyAsync(h->d_cost, cost, sizeof(ilqg_cost), cudaMemcpyHostToDevice, s));
        dcost = h->d_cost;
    }
    ilqg::FdDst dst{};
    for (int i = 0; i < ndst; i++) {
        if (!dsts[i]) return fail(h, ILQG_ERR_ARG, "null destination");
        dst.p[i] = dsts[i];
    }
    dst.n = ndst;
    return fd_launch(h, nknots, qpos, qvel, ctrl, warmstart, dcost, opts, dst, qacc_out, status, s, nullptr, 0, h->diag);
}

// ---- peer-visible device buffers (CUDA IPC): one per rank, opened by every other rank of the node
int ilqg_peer_alloc(ilqg_handle h, size_t bytes, void** dev_ptr, unsigned char* ipc_handle /* ILQG_IPC_HANDLE_BYTES */) {
    if (!h || !dev_ptr || !ipc_handle || bytes == 0) return ILQG_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    void* p = nullptr;
    CU(h, cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(h, e, "cudaIpcGetMemHandle"); }
    static_assert(sizeof(cudaIpcMemHandle_t) == ILQG_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(ipc_handle, &hd, sizeof(hd));
    CU(h, cudaMemset(p, 0, bytes));
    *dev_ptr = p;
    return ILQG_OK;
}
int ilqg_peer_open(ilqg_handle h, const unsigned char* ipc_handle, void** dev_ptr) {
    if (!h || !dev_ptr || !ipc_handle) return ILQG_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, ipc_handle, sizeof(hd));
    CU(h, cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return ILQG_OK;
}
int ilqg_peer_close(ilqg_handle h, void* dev_ptr) {
    if (!h || !dev_ptr) return ILQG_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaIpcCloseMemHandle(dev_ptr));
    return ILQG_OK;
}
// barrier over the ranks of a node: flags[r] = rank r's flag array (ILQG_MAX_PEERS ints, zero-initialised, own or peer-mapped);
// epoch must increase by one per call on every rank.  Asynchronous on `stream`; ilqg_peer_barrier_timed_out reports a lost peer.
int ilqg_peer_barrier(ilqg_handle h, int* const* flags, int nranks, int rank, int epoch, void* stream) {
    if (!h || !flags || nranks < 1 || nranks > ILQG_MAX_PEERS || rank < 0 || rank >= nranks) return ILQG_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    if (!h->d_timeout) {
        CU(h, cudaMalloc(&h->d_timeout, sizeof(int)));
        CU(h, cudaMemset(h->d_timeout, 0, sizeof(int)));
    }
    ilqg::PeerFlags f{};
    for (int r = 0; r < nranks; r++) {
        if (!flags[r]) return fail(h, ILQG_ERR_ARG, "null flag array");
        f.p[r] = flags[r];
    }
    ilqg::peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, nranks, rank, epoch, h->d_timeout);
    CU(h, cudaGetLastError());
    h->launches += 1;
    return ILQG_OK;
}
int ilqg_peer_barrier_timed_out(ilqg_handle h) {   // synchronises the device; reading clears the flag
    if (!h || !h->d_timeout) return 0;
    int v = 0;
    cudaSetDevice(h->device);
    if (cudaMemcpy(&v, h->d_timeout, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    if (v) cudaMemset(h->d_timeout, 0, sizeof(int));
    return v;
}
int ilqg_peer_free(ilqg_handle h, void* dev_ptr) {
    if (!h || !dev_ptr) return ILQG_ERR_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaFree(dev_ptr));
    return ILQG_OK;
}

// page-lock [p, p + bytes) once (no-op when it already is, or is too small to matter); failures leave the buffer pageable
static void pin_once(ilqg_handle h, const void* p, size_t bytes) {
    if (!h->pin_host || !p || bytes < (1u << 20)) return;
    for (size_t i = 0; i < h->regs.size(); i++) {
        auto& r = h->regs[i];
        if (r.p == p && r.bytes >= bytes) return;
        const char *a = (const char*)r.p, *b = (const char*)p;
        if (a < b + bytes && b < a + r.bytes) {   // overlaps an older registration (the caller re-allocated): drop that one
            cudaHostUnregister(const_cast<void*>(r.p));
            h->regs.erase(h->regs.begin() + i);
            i--;
        }
    }
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return;   // pinned or device memory already
    cudaGetLastError();
    if (h->regs.size() >= 16) {
        cudaHostUnregister(const_cast<void*>(h->regs.front().p));
        h->regs.erase(h->regs.begin());
    }
    if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) h->regs.push_back({p, bytes});
    else cudaGetLastError();
}

int ilqg_set_host_pinning(ilqg_handle h, int on) {
    if (!h) return ILQG_ERR_ARG;
    h->pin_host = on != 0;
    if (!on) {
        cudaSetDevice(h->device);
        for (auto& r : h->regs) cudaHostUnregister(const_cast<void*>(r.p));
        h->regs.clear();
    }
    return ILQG_OK;
}

int ilqg_fd_batch_host(ilqg_handle h, int nknots, const double* qpos, const double* qvel, const double* ctrl, const double* warmstart,
                       const ilqg_cost* cost, const ilqg_fd_opts* opts, double* deriv, double* qacc_out, int* status) {
    if (!h) return ILQG_ERR_ARG;
    if (nknots < 0 || (nknots > 0 && (!qpos || !qvel || !deriv || (h->model.nu > 0 && !ctrl)))) return fail(h, ILQG_ERR_ARG, "null buffer");
    if (nknots == 0) return ILQG_OK;
    CU(h, cudaSetDevice(h->device));
    const int nq = h->model.nq, nv = h->model.nv, nu = h->model.nu, nd = ilqg_deriv_size(&h->model);
    size_t n = (size_t)nknots;
    pin_once(h, deriv, n * nd * sizeof(double));
    pin_once(h, qpos, n * nq * sizeof(double));
    pin_once(h, qvel, n * nv * sizeof(double));
    pin_once(h, ctrl, n * nu * sizeof(double));
    pin_once(h, warmstart, n * nv * sizeof(double));
    pin_once(h, qacc_out, n * nv * sizeof(double));
    // one staging block: qpos | qvel | ctrl | warm | qacc | deriv | status | work-class scratch
    size_t off_q = 0, off_v = off_q + n * nq, off_u = off_v + n * nv, off_w = off_u + n * nu, off_a = off_w + n * nv,
           off_d = off_a + n * nv, ndbl = off_d + n * nd;
    // Pipeline over chunks of knots on streams BY FUNCTION — upload, compute (two, alternating), download — chained by events.
    // The chunks' kernels run (nearly) in chunk order, so the first download starts after one chunk's upload + kernels and the
    // copy engine then streams deriv back to back: the download is 5x the upload and, at PCIe rates, longer than the kernels
    // (86,016 hopper knots = 72 MB = 1.3 ms at the measured 57 GB/s).  Two compute streams let the next chunk's CTAs fill the
    // SMs the current chunk's last wave leaves idle.  Every chunk runs the kernel variant of the WHOLE batch, so the result is
    // bit-identical to one device call.  The compute streams form a PRIORITY LADDER (below): SURVEY-8d batch 1.85 -> 1.78 ms with
    // 8 chunks on 4 streams against two plain alternating streams.  Measured on B200 in round 1 (its lighter batch), 86,016 knots:
    // 1.73 ms with 4-6 chunks on two plain streams (one stream per chunk: 1.97;
    // one compute stream: 1.87; chunk sizes ramping up from 4096: 1.97 — small chunks of the split kernels are latency-bound;
    // kernels storing deriv straight into pinned host memory, no copy: 2.05 ms — the 432/288-byte segments of the reference
    // layout reach 35 GB/s over PCIe as SM stores against the copy engine's 57).
    size_t csize[ILQG_HOST_MAXCHUNKS];
    const bool overlap = h->eng->fd_calls_may_overlap();
    const size_t per_chunk = overlap && h->host_prio ? 10752 : 14336;
    size_t nchunks = n / per_chunk < 1 ? 1 : n / per_chunk, cmax = 0;
    if (h->host_chunks > 0) nchunks = (size_t)h->host_chunks;   // ILQG_HOST_CHUNKS (experiments)
    if (nchunks > ILQG_HOST_MAXCHUNKS) nchunks = ILQG_HOST_MAXCHUNKS;
    {
        const size_t c = (n + nchunks - 1) / nchunks;
        nchunks = 0;
        for (size_t lo = 0; lo < n; lo += c) csize[nchunks++] = lo + c <= n ? c : n - lo;
    }
    for (size_t i = 0; i < nchunks; i++) cmax = csize[i] > cmax ? csize[i] : cmax;
    const size_t scr = (h->eng->fd_scratch_ints((int)cmax, nknots) + 1) & ~(size_t)1;   // even: the scratch holds doubles too
    const size_t nstat = (n + 1) & ~(size_t)1;
    size_t bytes = ndbl * sizeof(double) + (nstat + ILQG_HOST_MAXCOMP * scr) * sizeof(int);
    int rc = ensure_stage(h, bytes);
    if (rc) return rc;
    double* b = (double*)h->d_stage;
    int* dstat = (int*)(b + ndbl);
    int* dscr = dstat + nstat;
    if (n > h->hstat_cap) {   // pinned landing buffer of the status words (the caller's array may be pageable)
        if (h->h_stat) cudaFreeHost(h->h_stat);
        h->h_stat = nullptr;
        h->hstat_cap = 0;
        CU(h, cudaHostAlloc((void**)&h->h_stat, n * sizeof(int), cudaHostAllocDefault));
        h->hstat_cap = n;
    }
    if (!h->pipe[0]) {
        for (int i = 0; i < 4; i++) CU(h, cudaStreamCreateWithFlags(&h->pipe[i], cudaStreamNonBlocking));
        for (int i = 0; i < 3 * ILQG_HOST_MAXCHUNKS; i++) CU(h, cudaEventCreateWithFlags(&h->pipe_ev[i], cudaEventDisableTiming));
    }
    // pageable caller buffers go through the pinned mirror (BounceCrew); h* = where the copy engines read / write
    BounceCrew crew;
    const double *hq = qpos, *hv = qvel, *hu = ctrl, *hw = warmstart;
    double *hd = deriv, *ha = qacc_out;
    {
        int K = h->host_threads;
        if (K < 0) {
            const unsigned hc = std::thread::hardware_concurrency();
            K = hc >= 16 ? 8 : (hc >= 8 ? 4 : (hc >= 4 ? 2 : 0));
        }
        if (K > 16) K = 16;
        auto pageable = [](const void* p) {
            if (!p) return false;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
            return at.type == cudaMemoryTypeUnregistered;
        };
        if (K > 0 && n * nd * sizeof(double) >= ((size_t)4 << 20)) {
            const bool pq = pageable(qpos), pv = pageable(qvel), pu = nu && pageable(ctrl), pw = pageable(warmstart), pd = pageable(deriv),
                       pa = pageable(qacc_out);
            if (pq || pv || pu || pw || pd || pa) {
                // (no pinned mirror or no threads to be had: the call still works, with the staging left to the driver)
                bool ok = true;
                if (ndbl > h->bounce_cap) {
                    if (h->h_bounce) cudaFreeHost(h->h_bounce);
                    h->h_bounce = nullptr;
                    h->bounce_cap = 0;
                    if (cudaHostAlloc((void**)&h->h_bounce, ndbl * sizeof(double), cudaHostAllocDefault) == cudaSuccess) h->bounce_cap = ndbl;
                    else { cudaGetLastError(); h->h_bounce = nullptr; ok = false; }
                }
                if (ok) {
                    double* m = h->h_bounce;
                    size_t lo = 0;
                    for (size_t ci = 0; ci < nchunks; lo += csize[ci], ci++) {
                        const size_t cn = csize[ci];
                        auto in = [&](bool on, size_t off, const double* src, size_t w) {
                            if (on) crew.in_jobs[ci].push_back({(char*)(m + off + lo * w), (const char*)(src + lo * w), cn * w * sizeof(double)});
                        };
                        in(pq, off_q, qpos, nq); in(pv, off_v, qvel, nv); in(pu, off_u, ctrl, nu); in(pw, off_w, warmstart, nv);
                        if (!cost) in(pd, off_d, deriv, nd);   // the caller's cost-gradient entries ride up with the block
                        if (pd) crew.out_jobs[ci].push_back({(char*)(deriv + lo * nd), (const char*)(m + off_d + lo * nd), cn * nd * sizeof(double)});
                        if (pa) crew.out_jobs[ci].push_back({(char*)(qacc_out + lo * nv), (const char*)(m + off_a + lo * nv), cn * nv * sizeof(double)});
                    }
                    crew.K = K;
                    crew.nchunks = (int)nchunks;
                    try {
                        if (!h->copy_pool) h->copy_pool = new CopyPool;
                        h->copy_pool->launch(&crew);
                    } catch (...) {   // thread creation failed: drop the pool (its destructor joins what did start) and the crew's work
                        delete h->copy_pool;
                        h->copy_pool = nullptr;
                        crew.K = 0;
                        crew.started = false;
                        ok = false;
                    }
                }
                if (ok) {
                    double* m = h->h_bounce;
                    if (pq) hq = m + off_q;
                    if (pv) hv = m + off_v;
                    if (pu) hu = m + off_u;
                    if (pw) hw = m + off_w;
                    if (pd) hd = m + off_d;
                    if (pa) ha = m + off_a;
                }
            }
        }
    }
    cudaStream_t up = h->pipe[0], down = h->pipe[2];
    int ncomp = overlap ? h->host_comp_streams : 1;
    if (ncomp < 1) ncomp = 1;
    if (ncomp > ILQG_HOST_MAXCOMP) ncomp = ILQG_HOST_MAXCOMP;
    const bool ladder = h->host_prio != 0 && ncomp > 1;
    if (ladder && !h->comp[0]) {
        int least = 0, greatest = 0;
        CU(h, cudaDeviceGetStreamPriorityRange(&least, &greatest));   // numerically lower = higher priority
        for (int j = 0; j < ILQG_HOST_MAXCOMP; j++) {
            const int pr = greatest + j < least ? greatest + j : least;
            CU(h, cudaStreamCreateWithPriority(&h->comp[j], cudaStreamNonBlocking, pr));
        }
    }
    const ilqg_cost* dcost = nullptr;
    if (cost) {
        CU(h, cudaMemcpyAsync(h->d_cost, cost, sizeof(ilqg_cost), cudaMemcpyHostToDevice, up));
        dcost = h->d_cost;
    }
    size_t lo = 0;
    for (int ci = 0; ci < (int)nchunks; lo += csize[ci], ci++) {
        const size_t cn = csize[ci];
        crew.inputs_of(ci);   // (pageable inputs: the chunk's slices are in the pinned mirror)
        CU(h, cudaMemcpyAsync(b + off_q + lo * nq, hq + lo * nq, cn * nq * sizeof(double), cudaMemcpyHostToDevice, up));
        CU(h, cudaMemcpyAsync(b + off_v + lo * nv, hv + lo * nv, cn * nv * sizeof(double), cudaMemcpyHostToDevice, up));
        if (nu) CU(h, cudaMemcpyAsync(b + off_u + lo * nu, hu + lo * nu, cn * nu * sizeof(double), cudaMemcpyHostToDevice, up));
        if (warmstart) CU(h, cudaMemcpyAsync(b + off_w + lo * nv, hw + lo * nv, cn * nv * sizeof(double), cudaMemcpyHostToDevice, up));
        else CU(h, cudaMemsetAsync(b + off_w + lo * nv, 0, cn * nv * sizeof(double), up));
        if (!cost)  // keep the caller's cost-gradient entries
            CU(h, cudaMemcpyAsync(b + off_d + lo * nd, hd + lo * nd, cn * nd * sizeof(double), cudaMemcpyHostToDevice, up));
        // compute stream of the chunk.  Two plain streams, alternating (default); or a priority ladder (ILQG_HOST_PRIO=1,
        // ILQG_HOST_COMP = streams): chunk ci on stream ci % ncomp, stream j at priority greatest + j — with as many streams as chunks the
        // CTA scheduler always prefers the EARLIEST chunk's pending CTAs (the chunk the download is waiting for) and later chunks
        // only fill the SMs it leaves idle.
        const int cslot = ncomp > 1 ? ci % ncomp : 0;
        cudaStream_t comp = ladder ? h->comp[cslot] : h->pipe[(cslot & 1) ? 3 : 1];
        CU(h, cudaEventRecord(h->pipe_ev[2 * ci], up));
        CU(h, cudaStreamWaitEvent(comp, h->pipe_ev[2 * ci], 0));
        ilqg::FdDst dst{};
        dst.p[0] = b + off_d + lo * nd;
        dst.n = 1;
        rc = fd_launch(h, (int)cn, b + off_q + lo * nq, b + off_v + lo * nv, b + off_u + lo * nu, b + off_w + lo * nv, dcost, opts, dst,
                       b + off_a + lo * nv, dstat + lo, comp, scr ? dscr + (ladder ? cslot : (cslot & 1)) * scr : nullptr, nknots);   // scratch per compute stream
        if (rc) return rc;
        CU(h, cudaEventRecord(h->pipe_ev[2 * ci + 1], comp));
        CU(h, cudaStreamWaitEvent(down, h->pipe_ev[2 * ci + 1], 0));
        CU(h, cudaMemcpyAsync(hd + lo * nd, b + off_d + lo * nd, cn * nd * sizeof(double), cudaMemcpyDeviceToHost, down));
        if (qacc_out) CU(h, cudaMemcpyAsync(ha + lo * nv, b + off_a + lo * nv, cn * nv * sizeof(double), cudaMemcpyDeviceToHost, down));
        CU(h, cudaMemcpyAsync(h->h_stat + lo, dstat + lo, cn * sizeof(int), cudaMemcpyDeviceToHost, down));
        if (crew.K) CU(h, cudaEventRecord(h->pipe_ev[2 * ILQG_HOST_MAXCHUNKS + ci], down));
    }
    if (crew.K) {   // hand every chunk to the crew as its download lands; the crew's last slice ends the call
        for (int ci = 0; ci < (int)nchunks; ci++) {
            CU(h, cudaEventSynchronize(h->pipe_ev[2 * ILQG_HOST_MAXCHUNKS + ci]));
            crew.landed(ci);
        }
        crew.finish();
    }
    CU(h, cudaStreamSynchronize(down));   // the last download follows every upload and every kernel
    int bad = 0;
    for (size_t i = 0; i < n; i++) {
        if (status) status[i] = h->h_stat[i];
        if (h->h_stat[i]) bad = 1;
    }
    return bad ? fail(h, ILQG_ERR_NONFINITE, "non-finite accelerations or exceeded contact capacity in at least one knot (see status[])") : ILQG_OK;
}

int ilqg_forward_batch_dev(ilqg_handle h, int n, const double* qpos, const double* qvel, const double* ctrl, double* warmstart,
                           double* qacc, void* stream) {
    if (!h) return ILQG_ERR_ARG;
    if (n < 0 || (n > 0 && (!qpos || !qvel || !qacc))) return fail(h, ILQG_ERR_ARG, "null buffer");
    CU(h, cudaSetDevice(h->device));
    CU(h, h->eng->forward(n, qpos, qvel, ctrl, warmstart, qacc, (cudaStream_t)stream));
    h->launches += 1;
    return ILQG_OK;
}

int ilqg_step_batch_dev(ilqg_handle h, int n, int nsteps, double* qpos, double* qvel, const double* ctrl, double* warmstart, double* qacc,
                        void* stream) {
    if (!h) return ILQG_ERR_ARG;
    if (n < 0 || nsteps < 0 || (n > 0 && (!qpos || !qvel))) return fail(h, ILQG_ERR_ARG, "null buffer");
    CU(h, cudaSetDevice(h->device));
    CU(h, h->eng->step(n, nsteps, qpos, qvel, ctrl, warmstart, qacc, (cudaStream_t)stream));
    h->launches += 1;
    return ILQG_OK;
}

static int state_host_call(ilqg_handle h, int n, int nsteps, bool stepping, double* qpos, double* qvel, const double* ctrl, double* warmstart,
                           double* qacc) {
    if (!h) return ILQG_ERR_ARG;
    if (n < 0 || (n > 0 && (!qpos || !qvel || (h->model.nu > 0 && !ctrl)))) return fail(h, ILQG_ERR_ARG, "null buffer");
    if (n == 0) return ILQG_OK;
    CU(h, cudaSetDevice(h->device));
    const int nq = h->model.nq, nv = h->model.nv, nu = h->model.nu;
    size_t N = (size_t)n;
    size_t off_q = 0, off_v = off_q + N * nq, off_u = off_v + N * nv, off_w = off_u + N * nu, off_a = off_w + N * nv, ndbl = off_a + N * nv;
    int rc = ensure_stage(h, ndbl * sizeof(double));
    if (rc) return rc;
    double* b = (double*)h->d_stage;
    cudaStream_t s = 0;
    CU(h, cudaMemcpyAsync(b + off_q, qpos, N * nq * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(h, cudaMemcpyAsync(b + off_v, qvel, N * nv * sizeof(double), cudaMemcpyHostToDevice, s));
    if (nu) CU(h, cudaMemcpyAsync(b + off_u, ctrl, N * nu * sizeof(double), cudaMemcpyHostToDevice, s));
    if (warmstart) CU(h, cudaMemcpyAsync(b + off_w, warmstart, N * nv * sizeof(double), cudaMemcpyHostToDevice, s));
    else CU(h, cudaMemsetAsync(b + off_w, 0, N * nv * sizeof(double), s));
    if (stepping) rc = ilqg_step_batch_dev(h, n, nsteps, b + off_q, b + off_v, b + off_u, b + off_w, b + off_a, s);
    else rc = ilqg_forward_batch_dev(h, n, b + off_q, b + off_v, b + off_u, b + off_w, b + off_a, s);
    if (rc) return rc;
    if (stepping) {
        CU(h, cudaMemcpyAsync(qpos, b + off_q, N * nq * sizeof(double), cudaMemcpyDeviceToHost, s));
        CU(h, cudaMemcpyAsync(qvel, b + off_v, N * nv * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    if (warmstart) CU(h, cudaMemcpyAsync(warmstart, b + off_w, N * nv * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (qacc) CU(h, cudaMemcpyAsync(qacc, b + off_a, N * nv * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    return ILQG_OK;
}

int ilqg_forward_batch_host(ilqg_handle h, int n, const double* qpos, const double* qvel, const double* ctrl, double* warmstart,
                            double* qacc) {
    if (n > 0 && !qacc) return h ? fail(h, ILQG_ERR_ARG, "null buffer") : ILQG_ERR_ARG;
    return state_host_call(h, n, 0, false, const_cast<double*>(qpos), const_cast<double*>(qvel), ctrl, warmstart, qacc);
}

int ilqg_step_batch_host(ilqg_handle h, int n, int nsteps, double* qpos, double* qvel, const double* ctrl, double* warmstart, double* qacc) {
    if (nsteps < 0) return h ? fail(h, ILQG_ERR_ARG, "negative nsteps") : ILQG_ERR_ARG;
    return state_host_call(h, n, nsteps, true, qpos, qvel, ctrl, warmstart, qacc);
}


// ------------------------------------------------------------------ batched iLQR workspace
struct ilqg_ilqr_s {
    ilqg_handle h = nullptr;
    ilqg::IlqrBuffers b{};
    ilqg_cost* d_cost = nullptr;
    bool has_cost = false;
    bool host_cost = false;
    double* d_Jtrace = nullptr;   // [trace_cap][ninst]
    int* d_acc_trace = nullptr;
    int trace_cap = 0, iters = 0;
    std::vector<void*> allocs;
    ilqg_fd_opts fd;
    // ilqg_ilqr_iterate as a CUDA graph: one iteration is 7-9 dependent launches of 0.04-0.15 ms kernels; replaying a captured
    // graph takes the host's launch cost (and, with 8 ranks on one host, its contention) off the chain
    cudaGraphExec_t graph = nullptr;
    int graph_niter = 0, graph_accept = 0, graph_warm = 0;   // what the graph was captured for; calls seen with that shape
    long graph_launches = 0;                                 // kernels one replay launches
    cudaStream_t cap_stream = nullptr;
    bool use_graph = true;
    void drop_graph() {
        if (graph) cudaGraphExecDestroy(graph);
        graph = nullptr;
        graph_warm = 0;
    }
    // Sub-batches.  Every phase of an iteration is ONE dependent chain per instance (the rollout: T x 4 evaluations on one warp per
    // scheduler; the Riccati sweep: T knots), so a phase takes the same time for 1024 instances as for 4096 and leaves most of the GPU idle.
    // The workspace is therefore laid out as nsub blocks of instances, each block time-major in itself (all arrays: block s starts at
    // per-instance size x sub_i0[s]), and ilqg_ilqr_iterate runs the blocks' iteration chains on nsub streams (a forked CUDA graph): while
    // one block is in its rollout the others are in FD or in the backward pass.  Measured on B200, 4096 pendulum problems: 0.268 -> 0.241 ms
    // per batch iteration with 4 blocks (tools/prof_ilqr_split.py).  Same arithmetic per instance; ILQG_ILQR_SUB overrides the count.
    static constexpr int MAXSUB = 16;
    int nsub = 1;
    int sub_i0[MAXSUB + 1] = {0};
    cudaStream_t sub_stream[MAXSUB] = {nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[MAXSUB] = {nullptr};
};

// the view of sub-batch s: same struct, pointers advanced to the block, ninst = the block's instances
static ilqg::IlqrBuffers ilqr_sub_view(const ilqg_ilqr_s* w, int s) {
    ilqg::IlqrBuffers b = w->b;
    if (w->nsub <= 1) return b;
    const ilqg_model& m = w->h->model;
    const size_t nq = m.nq, nv = m.nv, nu = m.nu, nx = 2 * nv, nd = (size_t)ilqg_deriv_size(&m), T = (size_t)b.N + 1, na = b.nalpha;
    const size_t i0 = w->sub_i0[s];
    b.ninst = w->sub_i0[s + 1] - w->sub_i0[s];
    b.nom_q += T * i0 * nq; b.nom_v += T * i0 * nv; b.nom_u += T * i0 * nu; b.nom_w += T * i0 * nv;
    b.init_q += i0 * nq; b.init_v += i0 * nv; b.init_w += i0 * nv;
    b.cand_q += na * T * i0 * nq; b.cand_v += na * T * i0 * nv; b.cand_u += na * T * i0 * nu; b.cand_w += na * T * i0 * nv; b.cand_J += na * i0;
    b.nom_J += i0; b.accepted += i0;
    b.K += T * i0 * nu * nx; b.k += T * i0 * nu; b.V += i0 * nx * nx; b.v += i0 * nx;
    b.deriv += T * i0 * nd;
    if (b.mu_i) b.mu_i += i0;
    if (b.cdiff) b.cdiff += T * i0 * nx;
    b.iter_dev += s;
    b.ticket += s;
    return b;
}
// run `f(s)` with the workspace narrowed to sub-batch s (w->b, the trace arrays), for every s
struct IlqrSubScope {
    ilqg_ilqr_s* w;
    ilqg::IlqrBuffers full;
    double* jt;
    int* at;
    int nsub;
    IlqrSubScope(ilqg_ilqr_s* w_, int s) : w(w_), full(w_->b), jt(w_->d_Jtrace), at(w_->d_acc_trace), nsub(w_->nsub) {
        w->b = ilqr_sub_view(w, s);
        w->d_Jtrace = jt + (size_t)w->trace_cap * w->sub_i0[s];
        w->d_acc_trace = at + (size_t)w->trace_cap * w->sub_i0[s];
        w->nsub = 1;
    }
    ~IlqrSubScope() { w->b = full; w->d_Jtrace = jt; w->d_acc_trace = at; w->nsub = nsub; }
};
// the origin stream forks into the sub-batch streams / joins them again (capturable)
static cudaError_t ilqr_fork(ilqg_ilqr_s* w, cudaStream_t s) {
    if (w->nsub <= 1) return cudaSuccess;
    cudaError_t e = cudaEventRecord(w->ev_fork, s);
    for (int i = 0; i < w->nsub && e == cudaSuccess; i++) e = cudaStreamWaitEvent(w->sub_stream[i], w->ev_fork, 0);
    return e;
}
static cudaError_t ilqr_join(ilqg_ilqr_s* w, cudaStream_t s) {
    if (w->nsub <= 1) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < w->nsub && e == cudaSuccess; i++) {
        e = cudaEventRecord(w->ev_join[i], w->sub_stream[i]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s, w->ev_join[i], 0);
    }
    return e;
}
static cudaStream_t ilqr_stream_of(ilqg_ilqr_s* w, int s, cudaStream_t origin) { return w->nsub <= 1 ? origin : w->sub_stream[s]; }

#define ILQR_ALLOC(w, ptr, count)                                                              \
    do {                                                                                       \
        void* p_ = nullptr;                                                                    \
        cudaError_t e_ = cudaMalloc(&p_, sizeof(*(ptr)) * (size_t)(count));                    \
        if (e_ != cudaSuccess) { ilqg_ilqr_destroy(w); return cuda_fail(h, e_, "cudaMalloc"); } \
        cudaMemset(p_, 0, sizeof(*(ptr)) * (size_t)(count));                                   \
        (w)->allocs.push_back(p_);                                                             \
        (ptr) = (decltype(ptr))p_;                                                             \
    } while (0)

int ilqg_ilqr_destroy(ilqg_ilqr w) {
    if (!w) return ILQG_OK;
    if (w->h) cudaSetDevice(w->h->device);
    w->drop_graph();
    if (w->cap_stream) cudaStreamDestroy(w->cap_stream);
    for (int i = 0; i < ilqg_ilqr_s::MAXSUB; i++) {
        if (w->sub_stream[i]) cudaStreamDestroy(w->sub_stream[i]);
        if (w->ev_join[i]) cudaEventDestroy(w->ev_join[i]);
    }
    if (w->ev_fork) cudaEventDestroy(w->ev_fork);
    for (void* p : w->allocs) cudaFree(p);
    delete w;
    return ILQG_OK;
}

int ilqg_ilqr_create(ilqg_handle h, int ninst, int N, int nalpha, const double* alphas, ilqg_ilqr* out) {
    if (!h || !out) return ILQG_ERR_ARG;
    *out = nullptr;
    if (ninst <= 0 || N < 1 || nalpha < 1 || nalpha > 64) return fail(h, ILQG_ERR_ARG, "bad iLQR sizes");
    if (!h->eng->ilqr_supported())
        return fail(h, ILQG_ERR_UNSUPPORTED, "no batched iLQR for this model: the thread-per-rollout engine needs nq == nv (the reference's state vector, "
                                             "SURVEY quirk Q9); the generic engine runs the humanoid (27 dofs, 21 actuators, Euler) in tangent coordinates");
    CU(h, cudaSetDevice(h->device));
    auto* w = new ilqg_ilqr_s();
    w->h = h;
    ilqg_fd_opts_default(&w->fd);
    const int nq = h->model.nq, nv = h->model.nv, nu = h->model.nu, nx = 2 * nv, nd = ilqg_deriv_size(&h->model);
    const size_t T = (size_t)N + 1, TI = T * ninst;
    auto& b = w->b;
    b.ninst = ninst; b.N = N; b.nalpha = nalpha; b.mu = 1000.0;  // ilqr.h:65
    b.corrected = 0;
    b.mu_i = nullptr; b.mu_factor = 1.0; b.mu_min = 1e-6; b.mu_max = 1e10;
    b.trace_cap = 256;
    if (const char* e = getenv("ILQG_ILQR_GRAPH")) w->use_graph = atoi(e) != 0;
    ILQR_ALLOC(w, b.nom_q, TI * nq); ILQR_ALLOC(w, b.nom_v, TI * nv); ILQR_ALLOC(w, b.nom_u, TI * nu); ILQR_ALLOC(w, b.nom_w, TI * nv);
    ILQR_ALLOC(w, b.init_q, (size_t)ninst * nq); ILQR_ALLOC(w, b.init_v, (size_t)ninst * nv); ILQR_ALLOC(w, b.init_w, (size_t)ninst * nv);
    ILQR_ALLOC(w, b.cand_q, nalpha * TI * nq); ILQR_ALLOC(w, b.cand_v, nalpha * TI * nv); ILQR_ALLOC(w, b.cand_u, nalpha * TI * nu);
    ILQR_ALLOC(w, b.cand_w, nalpha * TI * nv); ILQR_ALLOC(w, b.cand_J, (size_t)nalpha * ninst);
    ILQR_ALLOC(w, b.alphas, nalpha); ILQR_ALLOC(w, b.nom_J, ninst); ILQR_ALLOC(w, b.accepted, ninst);
    ILQR_ALLOC(w, b.K, TI * nu * nx); ILQR_ALLOC(w, b.k, TI * nu); ILQR_ALLOC(w, b.V, (size_t)ninst * nx * nx); ILQR_ALLOC(w, b.v, (size_t)ninst * nx);
    ILQR_ALLOC(w, b.deriv, TI * nd);
    b.cdiff = nullptr;
    if (nq != nv) {   // quaternions in qpos: the tangent-space extension (set_layout(1) is then mandatory, see ilqr_check_layout)
        ILQR_ALLOC(w, b.cdiff, TI * nx);
    }
    ILQR_ALLOC(w, w->d_cost, 1);
    w->trace_cap = 256;
    ILQR_ALLOC(w, w->d_Jtrace, (size_t)w->trace_cap * ninst); ILQR_ALLOC(w, w->d_acc_trace, (size_t)w->trace_cap * ninst);
    {
        // default: blocks whose FD is one wave of the column kernel, at least 256 instances each (measured on B200, tools/prof_ilqr_sub.py:
        // 4096 pendulum problems 15.4 M iterations/s in one block, 17.4 M in 6-8, 18.9-19.1 M in 12-14 = one wave of 7104 knots each)
        int want = 1;
        if (const int wave = h->eng->fd_wave_knots()) {
            want = (int)(((long long)ninst * (N + 1) + wave / 2) / wave);
            if (want > ninst / 256) want = ninst / 256;
        }
        if (const char* e = getenv("ILQG_ILQR_SUB")) want = atoi(e);
        if (!h->eng->fd_calls_may_overlap() || want < 1) want = 1;
        if (want > ilqg_ilqr_s::MAXSUB) want = ilqg_ilqr_s::MAXSUB;
        // block boundaries at multiples of 32 instances (warps of the rollout do not straddle blocks)
        int nsub = 0;
        w->sub_i0[0] = 0;
        for (int k = 1; k <= want; k++) {
            int hi = k == want ? ninst : (int)((long long)ninst * k / want) / 32 * 32;
            if (hi > w->sub_i0[nsub]) w->sub_i0[++nsub] = hi;
        }
        w->nsub = nsub;
        if (nsub > 1) {
            cudaError_t ce = cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming);
            for (int k = 0; k < nsub && ce == cudaSuccess; k++) {
                ce = cudaStreamCreateWithFlags(&w->sub_stream[k], cudaStreamNonBlocking);
                if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&w->ev_join[k], cudaEventDisableTiming);
            }
            if (ce != cudaSuccess) { ilqg_ilqr_destroy(w); return cuda_fail(h, ce, "cudaStreamCreate / cudaEventCreate"); }
        }
    }
    ILQR_ALLOC(w, b.iter_dev, 2 * ilqg_ilqr_s::MAXSUB);
    b.ticket = b.iter_dev + ilqg_ilqr_s::MAXSUB;
    std::vector<double> al(nalpha);
    for (int a = 0; a < nalpha; a++) al[a] = alphas ? alphas[a] : std::ldexp(1.0, -a);  // default ladder 1, 1/2, 1/4, ...
    CU(h, cudaMemcpy(b.alphas, al.data(), sizeof(double) * nalpha, cudaMemcpyHostToDevice));
    *out = w;
    return ILQG_OK;
}

int ilqg_ilqr_set_cost(ilqg_ilqr w, const ilqg_cost* cost) {
    if (!w) return ILQG_ERR_ARG;
    ilqg_handle h = w->h;
    CU(h, cudaSetDevice(h->device));
    if (cost) CU(h, cudaMemcpy(w->d_cost, cost, sizeof(ilqg_cost), cudaMemcpyHostToDevice));
    w->has_cost = true;
    w->drop_graph();   // the kernels' parameters are baked into a captured graph
    w->host_cost = cost == nullptr;  // NULL: the caller owns the cost (a host stepCostFn) and supplies the gradient rows itself
    return ILQG_OK;
}
// opt-in: read the deriv blocks as what they are (d qacc_j / d x_i at i + j*stride), i.e. undo quirk Q1/Q2.
// Default 0 = the reference's A/B, which is what parity is judged on.
int ilqg_ilqr_set_layout(ilqg_ilqr w, int corrected) {
    if (!w) return ILQG_ERR_ARG;
    w->b.corrected = corrected ? 1 : 0;
    w->drop_graph();
    return ILQG_OK;
}
int ilqg_ilqr_set_mu(ilqg_ilqr w, double mu) {
    if (!w) return ILQG_ERR_ARG;
    if (mu != w->b.mu) w->drop_graph();
    w->b.mu = mu;
    if (w->b.mu_i) {   // (re)start the schedule from this value
        std::vector<double> v((size_t)w->b.ninst, mu);
        CU(w->h, cudaSetDevice(w->h->device));
        CU(w->h, cudaMemcpy(w->b.mu_i, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice));
    }
    return ILQG_OK;
}
// opt-in: per-instance mu, divided by `factor` after an iteration whose ladder accepted a step and multiplied
// by it after a rejected one, clamped to [mu_min, mu_max].  factor <= 1 returns to the reference's constant mu.
int ilqg_ilqr_set_mu_schedule(ilqg_ilqr w, double factor, double mu_min, double mu_max) {
    if (!w || !(mu_min > 0) || !(mu_max >= mu_min)) return ILQG_ERR_ARG;
    ilqg_handle h = w->h;
    CU(h, cudaSetDevice(h->device));
    if (factor > 1.0 && !w->b.mu_i) {   // (a failed allocation leaves the workspace as it was: the caller still owns it)
        double* p = nullptr;
        CU(h, cudaMalloc(&p, sizeof(double) * (size_t)w->b.ninst));
        w->allocs.push_back(p);
        std::vector<double> v((size_t)w->b.ninst, w->b.mu);
        cudaError_t ce = cudaMemcpy(p, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) return cuda_fail(h, ce, "cudaMemcpy");
        w->b.mu_i = p;
    }
    w->b.mu_factor = factor; w->b.mu_min = mu_min; w->b.mu_max = mu_max;
    w->drop_graph();
    return ILQG_OK;
}

// setDInit (ilqr.h:110-113) for every instance; device pointers, instance-major
int ilqg_ilqr_set_state_dev(ilqg_ilqr w, const double* qpos, const double* qvel, const double* warm, void* stream) {
    if (!w || !qpos || !qvel) return ILQG_ERR_ARG;
    ilqg_handle h = w->h;
    cudaStream_t s = (cudaStream_t)stream;
    const int nq = h->model.nq, nv = h->model.nv, n = w->b.ninst;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaMemcpyAsync(w->b.init_q, qpos, sizeof(double) * (size_t)n * nq, cudaMemcpyDeviceToDevice, s));
    CU(h, cudaMemcpyAsync(w->b.init_v, qvel, sizeof(double) * (size_t)n * nv, cudaMemcpyDeviceToDevice, s));
    if (warm) CU(h, cudaMemcpyAsync(w->b.init_w, warm, sizeof(double) * (size_t)n * nv, cudaMemcpyDeviceToDevice, s));
    else CU(h, cudaMemsetAsync(w->b.init_w, 0, sizeof(double) * (size_t)n * nv, s));
    return ILQG_OK;
}

// ILQR constructor (ilqr.h:69-97): open-loop rollout of every instance under its initial control; K = k = 0 (quirk Q5)
int ilqg_ilqr_init_dev(ilqg_ilqr w, const double* qpos, const double* qvel, const double* ctrl, const double* warm, void* stream) {
    if (!w || !qpos || !qvel) return ILQG_ERR_ARG;
    ilqg_handle h = w->h;
    if (!w->has_cost) return fail(h, ILQG_ERR_ARG, "ilqg_ilqr_set_cost must be called first");
    cudaStream_t s = (cudaStream_t)stream;
    const int nq = h->model.nq, nv = h->model.nv, nu = h->model.nu, nx = 2 * nv, n = w->b.ninst;
    const size_t T = (size_t)w->b.N + 1, TI = T * n;
    int rc = ilqg_ilqr_set_state_dev(w, qpos, qvel, warm, stream);
    if (rc) return rc;
    auto& b = w->b;
    CU(h, cudaMemsetAsync(b.K, 0, sizeof(double) * TI * nu * nx, s));
    CU(h, cudaMemsetAsync(b.k, 0, sizeof(double) * TI * nu, s));
    CU(h, cudaMemsetAsync(b.nom_q, 0, sizeof(double) * TI * nq, s));
    CU(h, cudaMemsetAsync(b.nom_v, 0, sizeof(double) * TI * nv, s));
    CU(h, cudaMemsetAsync(b.iter_dev, 0, sizeof(int) * 2 * ilqg_ilqr_s::MAXSUB, s));   // iteration counters and tickets
    CU(h, ilqr_fork(w, s));
    for (int sb = 0; sb < w->nsub; sb++) {
        ilqg::IlqrBuffers one = ilqr_sub_view(w, sb);
        cudaStream_t ss = ilqr_stream_of(w, sb, s);
        const size_t i0 = w->sub_i0[sb], ni = one.ninst;
        if (nu > 0) {   // u*_n = the initial control at every knot (one launch per block: T copies per block were 1 ms of API calls at 12 blocks)
            const int per = (int)(ni * nu);
            ilqr_broadcast_ctrl_kernel<<<(per + 255) / 256, 256, 0, ss>>>(one.nom_u, ctrl ? ctrl + i0 * nu : nullptr, (int)T, per);
        }
        one.nalpha = 1;  // alphas[0] multiplies k = 0: any value gives the open-loop rollout
        one.mu_i = nullptr;  // the constructor's rollout is not a line-search outcome: the mu schedule does not move
        one.iter_dev = nullptr;
        CU(h, h->eng->ilqr_rollout(one, w->host_cost ? nullptr : w->d_cost, ss));
        CU(h, h->eng->ilqr_accept(one, 1, nullptr, nullptr, ss));
        h->launches += 3;
    }
    CU(h, ilqr_join(w, s));
    w->iters = 0;
    return ILQG_OK;
}

int ilqg_ilqr_get_host(ilqg_ilqr w, double* qpos, double* qvel, double* ctrl, double* K, double* k, double* V, double* v, double* Jtrace,
                       int* accepted);

// The three phases of ILQR::iterate (ilqr.h:179-186), each for the whole batch and asynchronous on `stream`.
// The reference assembles A/B through column-major views of the row-major deriv blocks and takes its state as 2 nv doubles at qpos
// (quirks Q1, Q9): with a quaternion in qpos neither is defined — only the corrected layout in tangent coordinates is.
static int ilqr_check_layout(ilqg_ilqr w) {
    if (w->b.cdiff && !w->b.corrected)
        return fail(w->h, ILQG_ERR_UNSUPPORTED, "this model has nq != nv: the reference's state vector and A/B views are undefined (quirk Q9); call ilqg_ilqr_set_layout(w, 1)");
    return ILQG_OK;
}
// (one sub-batch: `b` is its view, `s` its stream)
static int ilqr_forward_one(ilqg_ilqr w, const ilqg::IlqrBuffers& b, int sb, int accept_always, cudaStream_t s) {
    ilqg_handle h = w->h;
    const size_t toff = (size_t)w->trace_cap * w->sub_i0[sb];
    if (accept_always && b.nalpha == 1) {   // the reference's own mode: one launch (rollout over the nominal + the acceptance bookkeeping)
        cudaError_t de = cudaSuccess;
        if (h->eng->ilqr_rollout_direct(b, w->host_cost ? nullptr : w->d_cost, w->d_Jtrace + toff, w->d_acc_trace + toff, s, &de)) {
            CU(h, de);
            h->launches += 1;
            return ILQG_OK;
        }
    }
    CU(h, h->eng->ilqr_rollout(b, w->host_cost ? nullptr : w->d_cost, s));
    CU(h, h->eng->ilqr_accept(b, accept_always, w->d_Jtrace + toff, w->d_acc_trace + toff, s));   // trace slot: the device iteration counter
    h->launches += 3;
    return ILQG_OK;
}
// scratch of the FD calls of all sub-batches (sized once, before any stream forks or captures): centre accelerations and, for
// the stage-skipping split, the work-class bins — one slice per sub-batch
static int ilqr_fd_scratch(ilqg_ilqr w, size_t* bins_off /* [nsub + 1] */) {
    ilqg_handle h = w->h;
    const int T = w->b.N + 1;
    int rc = ensure_center(h, (size_t)T * w->b.ninst * h->model.nv);
    if (rc) return rc;
    bins_off[0] = 0;
    for (int sb = 0; sb < w->nsub; sb++) {
        const int nk = T * (w->sub_i0[sb + 1] - w->sub_i0[sb]);
        size_t want = h->eng->fd_scratch_ints(nk, nk);
        bins_off[sb + 1] = bins_off[sb] + ((want + 3) & ~(size_t)3);   // 16-byte aligned slices (the slice starts with doubles)
    }
    if (bins_off[w->nsub]) rc = ensure_bins(h, bins_off[w->nsub]);
    return rc;
}
static int ilqr_linearise_one(ilqg_ilqr w, const ilqg::IlqrBuffers& b, int sb, const size_t* bins_off, cudaStream_t s) {
    ilqg_handle h = w->h;
    const int nknots = (b.N + 1) * b.ninst;
    ilqg::FdDst dst{};
    dst.p[0] = b.deriv;
    dst.n = 1;
    int* scratch = bins_off[sb + 1] > bins_off[sb] ? h->d_bins + bins_off[sb] : nullptr;
    double* center = h->d_center + (size_t)(b.N + 1) * w->sub_i0[sb] * h->model.nv;
    CU(h, h->eng->fd(nknots, b.nom_q, b.nom_v, b.nom_u, b.nom_w, w->host_cost ? nullptr : w->d_cost, w->fd, dst, center, nullptr, scratch, nknots, s,
                     nullptr));
    h->launches += h->eng->fd_launches();
    return ILQG_OK;
}
static int ilqr_backward_one(ilqg_ilqr w, const ilqg::IlqrBuffers& b, cudaStream_t s) {
    ilqg_handle h = w->h;
    CU(h, h->eng->ilqr_backward(b, s));
    h->launches += 1;
    return ILQG_OK;
}
// niter x (forward, linearise, backward) of every sub-batch, each sub-batch's chain on its own stream between a fork and a join
static int ilqr_run_phases(ilqg_ilqr w, int niter, bool fwd, bool lin, bool bwd, int accept_always, void* stream) {
    ilqg_handle h = w->h;
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, cudaSetDevice(h->device));
    size_t bins_off[ilqg_ilqr_s::MAXSUB + 1] = {0};
    if (lin) { if (int rc = ilqr_fd_scratch(w, bins_off)) return rc; }
    CU(h, ilqr_fork(w, s));
    for (int sb = 0; sb < w->nsub; sb++) {
        const ilqg::IlqrBuffers b = ilqr_sub_view(w, sb);
        cudaStream_t ss = ilqr_stream_of(w, sb, s);
        for (int it = 0; it < niter; it++) {
            int rc;
            if (fwd && (rc = ilqr_forward_one(w, b, sb, accept_always, ss))) return rc;
            if (lin && (rc = ilqr_linearise_one(w, b, sb, bins_off, ss))) return rc;
            if (bwd && (rc = ilqr_backward_one(w, b, ss))) return rc;
        }
    }
    CU(h, ilqr_join(w, s));
    return ILQG_OK;
}
int ilqg_ilqr_forward(ilqg_ilqr w, int accept_always, void* stream) {   // forwardPass (+ A10 acceptance) + setDInit(dArray[N])
    if (!w) return ILQG_ERR_ARG;
    ilqg_handle h = w->h;
    if (!w->has_cost) return fail(h, ILQG_ERR_ARG, "ilqg_ilqr_set_cost must be called first");
    if (int lrc = ilqr_check_layout(w)) return lrc;
    if (int rc = ilqr_run_phases(w, 1, true, false, false, accept_always, stream)) return rc;
    w->iters++;
    return ILQG_OK;
}
int ilqg_ilqr_linearise(ilqg_ilqr w, void* stream) {                     // FD at every knot of every instance
    if (!w) return ILQG_ERR_ARG;
    return ilqr_run_phases(w, 1, false, true, false, 0, stream);
}
int ilqg_ilqr_backward(ilqg_ilqr w, void* stream) {                      // initV + backwardPass
    if (!w) return ILQG_ERR_ARG;
    if (int lrc = ilqr_check_layout(w)) return lrc;
    return ilqr_run_phases(w, 1, false, false, true, 0, stream);
}
static int ilqr_iterate_plain(ilqg_ilqr w, int niter, int accept_always, void* stream) {
    ilqg_handle h = w->h;
    if (!w->has_cost) return fail(h, ILQG_ERR_ARG, "ilqg_ilqr_set_cost must be called first");
    if (int lrc = ilqr_check_layout(w)) return lrc;
    if (int rc = ilqr_run_phases(w, niter, true, true, true, accept_always, stream)) return rc;
    w->iters += niter;
    return ILQG_OK;
}
int ilqg_ilqr_iterate(ilqg_ilqr w, int niter, int accept_always, void* stream) {
    if (!w || niter < 0) return ILQG_ERR_ARG;
    ilqg_handle h = w->h;
    if (w->host_cost) return fail(h, ILQG_ERR_ARG, "host-cost workspaces drive the phases themselves (cost rows come from the host)");
    if (niter == 0) return ILQG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    accept_always = accept_always ? 1 : 0;
    if (!w->use_graph || h->profiling) return ilqr_iterate_plain(w, niter, accept_always, stream);
    CU(h, cudaSetDevice(h->device));
    if (w->graph && (w->graph_niter != niter || w->graph_accept != accept_always)) w->drop_graph();
    if (!w->graph) {
        // The first call of a shape runs launch by launch (it also sizes the handle's scratch: no allocation may happen while a
        // stream is capturing); the second captures the same sequence on a private stream; from then on the graph is replayed.
        if (w->graph_niter != niter || w->graph_accept != accept_always || w->graph_warm == 0) {
            w->graph_niter = niter; w->graph_accept = accept_always; w->graph_warm = 1;
            return ilqr_iterate_plain(w, niter, accept_always, stream);
        }
        if (!w->cap_stream) CU(h, cudaStreamCreateWithFlags(&w->cap_stream, cudaStreamNonBlocking));
        const int iters0 = w->iters;
        const long launches0 = h->launches;
        const int pdl0 = h->eng->fd_pdl;
        h->eng->fd_pdl = 0;   // plain kernel-to-kernel edges inside the graph (programmatic edges were measured too: no difference, 0.2168 / 0.2162 ms)
        CU(h, cudaStreamBeginCapture(w->cap_stream, cudaStreamCaptureModeThreadLocal));
        int rc = ilqr_iterate_plain(w, niter, accept_always, w->cap_stream);
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(w->cap_stream, &g);
        h->eng->fd_pdl = pdl0;
        w->iters = iters0;            // nothing ran yet
        w->graph_launches = h->launches - launches0;
        h->launches = launches0;
        if (rc || ce != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            w->use_graph = false;     // capture is not possible here: stay with plain launches
            return ilqr_iterate_plain(w, niter, accept_always, stream);
        }
        ce = cudaGraphInstantiate(&w->graph, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) {
            w->graph = nullptr;
            cudaGetLastError();
            w->use_graph = false;
            return ilqr_iterate_plain(w, niter, accept_always, stream);
        }
    }
    CU(h, cudaGraphLaunch(w->graph, s));
    w->iters += niter;
    h->launches += w->graph_launches;
    return ILQG_OK;
}

// host-cost mode: the 2nv+nu cost-gradient entries of every knot's deriv block, instance-major [ninst][T][2nv+nu]
int ilqg_ilqr_put_cost_rows_host(ilqg_ilqr w, const double* rows) {
    if (!w || !rows) return ILQG_ERR_ARG;
    if (w->nsub > 1) {   // block by block (see ilqg_ilqr_s)
        const size_t per = (size_t)(w->b.N + 1) * (2 * w->h->model.nv + w->h->model.nu);
        for (int sb = 0, n = w->nsub; sb < n; sb++) {
            IlqrSubScope scope(w, sb);
            if (int rc = ilqg_ilqr_put_cost_rows_host(w, rows + per * scope.w->sub_i0[sb])) return rc;
        }
        return ILQG_OK;
    }
    ilqg_handle h = w->h;
    CU(h, cudaSetDevice(h->device));
    auto& b = w->b;
    const int nv = h->model.nv, nu = h->model.nu, nd = ilqg_deriv_size(&h->model), nr = 2 * nv + nu, T = b.N + 1, n = b.ninst;
    CU(h, cudaDeviceSynchronize());
    // one strided copy per instance: T rows of nr doubles into the tails of its T deriv blocks (time-major: block stride n * nd)
    for (int i = 0; i < n; i++)
        CU(h, cudaMemcpy2D(b.deriv + (size_t)i * nd + (nd - nr), sizeof(double) * (size_t)n * nd, rows + (size_t)i * T * nr, sizeof(double) * nr,
                           sizeof(double) * nr, (size_t)T, cudaMemcpyHostToDevice));
    return ILQG_OK;
}

// the knots of the nominal incl. warm starts, instance-major [ninst][T][.]
int ilqg_ilqr_get_knots_host(ilqg_ilqr w, double* qpos, double* qvel, double* ctrl, double* warm) {
    if (!w) return ILQG_ERR_ARG;
    if (w->nsub > 1) {
        const ilqg_model& m = w->h->model;
        const size_t T = (size_t)w->b.N + 1;
        for (int sb = 0, n = w->nsub; sb < n; sb++) {
            IlqrSubScope scope(w, sb);
            const size_t o = T * w->sub_i0[sb];
            if (int rc = ilqg_ilqr_get_knots_host(w, qpos ? qpos + o * m.nq : nullptr, qvel ? qvel + o * m.nv : nullptr, ctrl ? ctrl + o * m.nu : nullptr,
                                                  warm ? warm + o * m.nv : nullptr))
                return rc;
        }
        return ILQG_OK;
    }
    ilqg_handle h = w->h;
    int rc = ilqg_ilqr_get_host(w, qpos, qvel, ctrl, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc || !warm) return rc;
    auto& b = w->b;
    const int nv = h->model.nv, T = b.N + 1, n = b.ninst;
    std::vector<double> tmp((size_t)T * n * nv);
    CU(h, cudaMemcpy(tmp.data(), b.nom_w, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
    for (int t = 0; t < T; t++)
        for (int i = 0; i < n; i++) memcpy(warm + ((size_t)i * T + t) * nv, tmp.data() + ((size_t)t * n + i) * nv, sizeof(double) * nv);
    return ILQG_OK;
}

int ilqg_ilqr_iterations_done(ilqg_ilqr w) { return w ? w->iters : 0; }

// Results to the host, instance-major: traj [ninst][T][.], K [ninst][T][nu*nx], k [ninst][T][nu], V [ninst][nx*nx],
// v [ninst][nx], Jtrace / accepted [ninst][iters] (last min(iters, 256) iterations).  Any pointer may be NULL.
int ilqg_ilqr_get_host(ilqg_ilqr w, double* qpos, double* qvel, double* ctrl, double* K, double* k, double* V, double* v, double* Jtrace,
                       int* accepted) {
    if (!w) return ILQG_ERR_ARG;
    if (w->nsub > 1) {
        const ilqg_model& m = w->h->model;
        const size_t T = (size_t)w->b.N + 1, nx = 2 * (size_t)m.nv, kept = w->iters < w->trace_cap ? w->iters : w->trace_cap;
        for (int sb = 0, n = w->nsub; sb < n; sb++) {
            IlqrSubScope scope(w, sb);
            const size_t i0 = w->sub_i0[sb], o = T * i0;
            if (int rc = ilqg_ilqr_get_host(w, qpos ? qpos + o * m.nq : nullptr, qvel ? qvel + o * m.nv : nullptr, ctrl ? ctrl + o * m.nu : nullptr,
                                            K ? K + o * m.nu * nx : nullptr, k ? k + o * m.nu : nullptr, V ? V + i0 * nx * nx : nullptr,
                                            v ? v + i0 * nx : nullptr, Jtrace ? Jtrace + i0 * kept : nullptr, accepted ? accepted + i0 * kept : nullptr))
                return rc;
        }
        return ILQG_OK;
    }
    ilqg_handle h = w->h;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaDeviceSynchronize());
    auto& b = w->b;
    const int nq = h->model.nq, nv = h->model.nv, nu = h->model.nu, nx = 2 * nv, n = b.ninst, T = b.N + 1;
    auto fetch_tm = [&](const double* dev, double* host, int width) -> int {  // time-major device -> instance-major host
        if (!host) return ILQG_OK;
        std::vector<double> tmp((size_t)T * n * width);
        CU(h, cudaMemcpy(tmp.data(), dev, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
        for (int t = 0; t < T; t++)
            for (int i = 0; i < n; i++)
                memcpy(host + ((size_t)i * T + t) * width, tmp.data() + ((size_t)t * n + i) * width, sizeof(double) * width);
        return ILQG_OK;
    };
    int rc;
    if ((rc = fetch_tm(b.nom_q, qpos, nq))) return rc;
    if ((rc = fetch_tm(b.nom_v, qvel, nv))) return rc;
    if ((rc = fetch_tm(b.nom_u, ctrl, nu))) return rc;
    if ((rc = fetch_tm(b.K, K, nu * nx))) return rc;
    if ((rc = fetch_tm(b.k, k, nu))) return rc;
    if (V) CU(h, cudaMemcpy(V, b.V, sizeof(double) * (size_t)n * nx * nx, cudaMemcpyDeviceToHost));
    if (v) CU(h, cudaMemcpy(v, b.v, sizeof(double) * (size_t)n * nx, cudaMemcpyDeviceToHost));
    int kept = w->iters < w->trace_cap ? w->iters : w->trace_cap;
    if ((Jtrace || accepted) && kept > 0) {
        std::vector<double> tj((size_t)w->trace_cap * n);
        std::vector<int> ta((size_t)w->trace_cap * n);
        CU(h, cudaMemcpy(tj.data(), w->d_Jtrace, sizeof(double) * tj.size(), cudaMemcpyDeviceToHost));
        CU(h, cudaMemcpy(ta.data(), w->d_acc_trace, sizeof(int) * ta.size(), cudaMemcpyDeviceToHost));
        int first = w->iters - kept;
        for (int j = 0; j < kept; j++) {
            int slot = (first + j) % w->trace_cap;
            for (int i = 0; i < n; i++) {
                if (Jtrace) Jtrace[(size_t)i * kept + j] = tj[(size_t)slot * n + i];
                if (accepted) accepted[(size_t)i * kept + j] = ta[(size_t)slot * n + i];
            }
        }
    }
    return ILQG_OK;
}

// What one MPC step hands back to the plant (InvertedPendulum::forward, /root/reference/src/inverted_pendulum/inverted_pendulum.cpp:26):
// the first control of every problem, u0[ninst][nu] = dArray[N]->ctrl, and (optionally) the cost trace [ninst][kept] of the
// last kept = min(iterations, 256) iterations — without downloading the trajectories.
int ilqg_ilqr_get_first_control_last_host(ilqg_ilqr w, int nlast, double* u0, double* Jtrace) {
    if (!w || nlast < 0) return ILQG_ERR_ARG;
    // one gather kernel packs u0[ninst][nu] | Jtrace[ninst][kept] instance-major into the staging buffer (whatever the workspace's
    // block layout), one copy brings it down
    ilqg_handle h = w->h;
    CU(h, cudaSetDevice(h->device));
    int kept = w->iters < w->trace_cap ? w->iters : w->trace_cap;
    if (kept > nlast) kept = nlast;
    if (!Jtrace) kept = 0;
    const int nu = u0 ? h->model.nu : 0, n = w->b.ninst;
    const size_t tot = (size_t)n * (nu + kept);
    CU(h, cudaDeviceSynchronize());
    if (tot == 0) return ILQG_OK;
    if (int rc = ensure_stage(h, tot * sizeof(double))) return rc;
    IlqrSubTable tab{};
    tab.nsub = w->nsub;
    for (int k = 0; k <= w->nsub; k++) tab.i0[k] = w->sub_i0[k];
    if (w->nsub <= 1) { tab.nsub = 1; tab.i0[0] = 0; tab.i0[1] = n; }
    double* st = (double*)h->d_stage;
    ilqr_pack_first_control_kernel<<<(n + 127) / 128, 128>>>(w->b, tab, w->d_Jtrace, nu, kept, (w->iters - kept) % w->trace_cap, st, st + (size_t)n * nu);
    CU(h, cudaGetLastError());
    h->launches += 1;
    if (h->stage_host_bytes < tot * sizeof(double)) {   // pinned landing buffer, grown on demand
        if (h->stage_host) cudaFreeHost(h->stage_host);
        h->stage_host = nullptr; h->stage_host_bytes = 0;
        CU(h, cudaMallocHost(&h->stage_host, tot * sizeof(double)));
        h->stage_host_bytes = tot * sizeof(double);
    }
    double* tmp = (double*)h->stage_host;
    CU(h, cudaMemcpy(tmp, st, sizeof(double) * tot, cudaMemcpyDeviceToHost));
    if (nu) memcpy(u0, tmp, sizeof(double) * (size_t)n * nu);
    if (kept) memcpy(Jtrace, tmp + (size_t)n * nu, sizeof(double) * (size_t)n * kept);
    return ILQG_OK;
}
int ilqg_ilqr_get_first_control_host(ilqg_ilqr w, double* u0, double* Jtrace) {
    return w ? ilqg_ilqr_get_first_control_last_host(w, w->trace_cap, u0, Jtrace) : ILQG_ERR_ARG;
}

// host-pointer conveniences (copy in, call the device flavour)
static int ilqr_upload(ilqg_ilqr w, const double* qpos, const double* qvel, const double* ctrl, const double* warm, double** dq, double** dv,
                       double** du, double** dw) {
    ilqg_handle h = w->h;
    const int nq = h->model.nq, nv = h->model.nv, nu = h->model.nu, n = w->b.ninst;
    size_t tot = (size_t)n * (nq + 2 * nv + nu);
    int rc = ensure_stage(h, tot * sizeof(double));
    if (rc) return rc;
    double* base = (double*)h->d_stage;
    *dq = base; *dv = *dq + (size_t)n * nq; *du = *dv + (size_t)n * nv; *dw = *du + (size_t)n * nu;
    CU(h, cudaMemcpy(*dq, qpos, sizeof(double) * (size_t)n * nq, cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(*dv, qvel, sizeof(double) * (size_t)n * nv, cudaMemcpyHostToDevice));
    if (ctrl && nu) CU(h, cudaMemcpy(*du, ctrl, sizeof(double) * (size_t)n * nu, cudaMemcpyHostToDevice));
    else if (nu) CU(h, cudaMemset(*du, 0, sizeof(double) * (size_t)n * nu));
    if (warm) CU(h, cudaMemcpy(*dw, warm, sizeof(double) * (size_t)n * nv, cudaMemcpyHostToDevice));
    else CU(h, cudaMemset(*dw, 0, sizeof(double) * (size_t)n * nv));
    return ILQG_OK;
}
int ilqg_ilqr_init_host(ilqg_ilqr w, const double* qpos, const double* qvel, const double* ctrl, const double* warm) {
    if (!w || !qpos || !qvel) return ILQG_ERR_ARG;
    CU(w->h, cudaSetDevice(w->h->device));
    double *dq, *dv, *du, *dw;
    int rc = ilqr_upload(w, qpos, qvel, ctrl, warm, &dq, &dv, &du, &dw);
    if (rc) return rc;
    rc = ilqg_ilqr_init_dev(w, dq, dv, du, dw, nullptr);
    if (rc) return rc;
    CU(w->h, cudaDeviceSynchronize());
    return ILQG_OK;
}
int ilqg_ilqr_set_state_host(ilqg_ilqr w, const double* qpos, const double* qvel, const double* warm) {
    if (!w || !qpos || !qvel) return ILQG_ERR_ARG;
    CU(w->h, cudaSetDevice(w->h->device));
    double *dq, *dv, *du, *dw;
    int rc = ilqr_upload(w, qpos, qvel, nullptr, warm, &dq, &dv, &du, &dw);
    if (rc) return rc;
    rc = ilqg_ilqr_set_state_dev(w, dq, dv, dw, nullptr);
    if (rc) return rc;
    CU(w->h, cudaDeviceSynchronize());
    return ILQG_OK;
}

}  // extern "C"
