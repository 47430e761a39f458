// ilqr.cuh — batched iLQR on the GPU: closed-loop rollouts for all line-search step sizes at once,
// ladder-order acceptance, and the Riccati backward pass, for `ninst` independent problems.
//
// Replaces, for a whole batch and without leaving the device:
//   ILQR::forwardPass   /root/reference/inc/ilqr.h:116-130   -> ilqr_rollout_kernel (one thread per instance x alpha)
//   (A10 line search: absent from the reference; specification in oracle/mjo_ilqr.c) -> ilqr_accept_kernel
//   Differentiator::updateDerivatives  /root/reference/inc/differentiator.h:85-93 -> fused into the backward kernel
//   ILQR::initV / backwardPass  /root/reference/inc/ilqr.h:100-107,133-176 -> ilqr_backward_kernel (one lane group per instance)
// The FD linearisation of all T x ninst knots between them is fd_center_kernel / fd_perturb_kernel (ilqg.cu).
//
// Device layout is time-major so that the threads of a warp (consecutive instances) touch consecutive memory:
//   knot (n, i) of instance i lives at index n * ninst + i;  n = N is the initial knot, n = 0 the final one
//   (the reference's dArray indexing, ilqr.h:52).  K[n] is nu x nx column-major, as Eigen stores it.
#pragma once
#include "dyn.cuh"

namespace ilqg {

template <class T>
DEV double cost_eval(const ilqg_cost& c, const double (&q)[T::NQ], const double (&v)[T::NV], const double (&u)[nz(T::NU)]);

// bring [p, p + bytes) towards the SM (every 128-byte line the range touches)
DEV void prefetch_l1(const void* p, int bytes) {
    const char* c = static_cast<const char*>(p);
    for (int off = 0; off < bytes; off += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(c + off));
    if (bytes > 8 && (bytes & 127) != 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(c + bytes - 8));
}

struct IlqrBuffers {
    int ninst, N, nalpha;
    // nominal trajectory and the state every forward pass starts from (ILQR::d)
    double *nom_q, *nom_v, *nom_u, *nom_w;
    double *init_q, *init_v, *init_w;
    // candidates: [nalpha][T][ninst]
    double *cand_q, *cand_v, *cand_u, *cand_w, *cand_J;
    double *alphas;     // [nalpha]
    double *nom_J;      // [ninst]
    int* accepted;      // [ninst] index of the accepted alpha in the last iteration (-1: none)
    double *K, *k;      // [T][ninst][nu*nx], [T][ninst][nu]
    double *V, *v;      // [ninst][nx*nx] (column-major), [ninst][nx]
    double* deriv;      // [T][ninst][ND]
    double mu;          // Levenberg-Marquardt term (ilqr.h:65,166) when mu_i is NULL
    double* mu_i;       // [ninst] per-instance mu under the opt-in schedule, else NULL
    double mu_factor, mu_min, mu_max;
    int corrected;      // 0: A/B through the reference's column-major views of the row-major deriv blocks (quirk Q1); 1: transposed back
    double* cdiff;      // [T][ninst][nx] x*_{n-1} (-) x*_n in the tangent space, filled before the backward pass for models with
                        // quaternions (nq != nv: the opt-in extension beyond quirk Q9), else NULL: plain differences of qpos | qvel
    int* iter_dev;      // iterations done (device counter: the trace slot of a pass is taken from it, so that a captured CUDA graph
                        // of iterations can be replayed), or NULL for a pass that is not an iteration (the constructor's rollout)
    int trace_cap;      // slots of the cost / accepted-alpha traces
    int* ticket;        // CTAs of a direct-mode rollout that have finished (the last one advances iter_dev and clears it)
};

// ------------------------------------------------------------------ forward pass, all alphas at once
// DIRECT (the reference's own mode: one alpha, accepted unconditionally — ilqr.h:116-130 has no line search): the rollout IS the new
// nominal, so the snapshots go straight over the nominal's knots (a thread reads knot n's record before it writes knot n, and touches
// no other instance's data) and the bookkeeping of ilqr_accept_kernel / ilqr_commit_kernel — cost, trace slot, mu schedule, iteration
// counter — is done here: one launch instead of three per iteration.  The trace slot is read from the device counter at the start;
// the counter is advanced by the last CTA to finish (ticket), i.e. after every thread has read it.
template <class T, bool DIRECT = false>
__global__ void __launch_bounds__(128) ilqr_rollout_kernel(const __grid_constant__ DevModel<T> m, IlqrBuffers b, const ilqg_cost* __restrict__ cost,
                                                           double* __restrict__ Jtrace = nullptr, int* __restrict__ acc_trace = nullptr) {
    constexpr int NQ = T::NQ, NV = T::NV, NU = T::NU, NX = 2 * NV;
    static_assert(NQ == NV, "the reference's state vector is 2*nv doubles starting at qpos (quirk Q9)");
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int ninst = b.ninst;
    size_t slot = 0;
    if constexpr (DIRECT) slot = b.iter_dev ? (size_t)(*b.iter_dev % b.trace_cap) * ninst : 0;
    if (tid >= ninst * b.nalpha) return;   // (a CTA's first thread always has an instance: it does the ticket below)
    const int a = tid / ninst, i = tid - a * ninst;   // consecutive threads = consecutive instances
    const double alpha = b.alphas[a];
    double q[NQ], v[NV], u[nz(NU)], warm[NV], qacc[NV];
    sfor<0, NQ>([&](auto ii) { q[IDX(ii)] = b.init_q[(size_t)i * NQ + IDX(ii)]; });
    sfor<0, NV>([&](auto ii) { v[IDX(ii)] = b.init_v[(size_t)i * NV + IDX(ii)]; warm[IDX(ii)] = b.init_w[(size_t)i * NV + IDX(ii)]; qacc[IDX(ii)] = 0; });
    Work<T> w;
    double J = 0;
    const size_t T1 = (size_t)(b.N + 1) * ninst;
    // The rollout is one dependent chain per thread; the knot's feedback data (K, k, nominal) does not depend on the state, so
    // the NEXT knot's record is brought into L1 while the current mj_step runs instead of stalling the chain on an L2 round trip.
    // (A prefetch instruction, not loads into registers: ncu showed the compiler spilling the early-loaded record to local memory
    //  right behind the loads — the spill stores waited for the L2 round trip the loads were meant to hide, 14 % of the kernel's
    //  stall samples; with the prefetch 10 % sit on the first use instead.  Staging the record in shared memory with cp.async was
    //  also measured: pendulum forward pass 0.138 -> 0.137 ms, hopper 0.328 -> 0.348 ms — not kept.)
    double Kn[nz(NU) * NX], kn_[nz(NU)], xq[NQ], xv[NV], xu[nz(NU)];
    auto prefetch = [&](int n) {
        const size_t kn = (size_t)n * ninst + i;
        prefetch_l1(b.K + kn * NU * NX, NU * NX * 8);
        prefetch_l1(b.k + kn * NU, NU * 8);
        prefetch_l1(b.nom_u + kn * NU, NU * 8);
        prefetch_l1(b.nom_q + kn * NQ, NQ * 8);
        prefetch_l1(b.nom_v + kn * NV, NV * 8);
    };
    auto fetch = [&](int n) {
        const size_t kn = (size_t)n * ninst + i;
        sfor<0, NU * NX>([&](auto ee) { Kn[IDX(ee)] = b.K[kn * NU * NX + IDX(ee)]; });
        sfor<0, NU>([&](auto rr) { kn_[IDX(rr)] = b.k[kn * NU + IDX(rr)]; xu[IDX(rr)] = b.nom_u[kn * NU + IDX(rr)]; });
        sfor<0, NQ>([&](auto ii) { xq[IDX(ii)] = b.nom_q[kn * NQ + IDX(ii)]; });
        sfor<0, NV>([&](auto ii) { xv[IDX(ii)] = b.nom_v[kn * NV + IDX(ii)]; });
    };
    for (int n = b.N; n >= 0; n--) {
        const size_t kn = (size_t)n * ninst + i;
        fetch(n);
        // u = K[n] (x - x*_n) + alpha k[n] + u*_n      (ilqr.h:126; alpha = 1 there)
        double dx[NX];
        sfor<0, NV>([&](auto ii) {
            dx[IDX(ii)] = q[IDX(ii)] - xq[IDX(ii)];
            dx[NV + IDX(ii)] = v[IDX(ii)] - xv[IDX(ii)];
        });
        sfor<0, NU>([&](auto rr) {
            constexpr int r = IDX(rr);
            double s = 0;
            sfor<0, NX>([&](auto cc) { s += Kn[r + IDX(cc) * NU] * dx[IDX(cc)]; });
            u[r] = s + alpha * kn_[r] + xu[r];
        });
        if (n > 0) prefetch(n - 1);
        // snapshot the knot (cpMjData(dArray[n], d), ilqr.h:127) into this alpha's candidate (DIRECT: into the nominal itself)
        if constexpr (DIRECT) {
            sfor<0, NQ>([&](auto ii) { b.nom_q[kn * NQ + IDX(ii)] = q[IDX(ii)]; });
            sfor<0, NV>([&](auto ii) { b.nom_v[kn * NV + IDX(ii)] = v[IDX(ii)]; b.nom_w[kn * NV + IDX(ii)] = warm[IDX(ii)]; });
            sfor<0, NU>([&](auto ii) { b.nom_u[kn * NU + IDX(ii)] = u[IDX(ii)]; });
        } else {
            const size_t cn = (size_t)a * T1 + kn;
            sfor<0, NQ>([&](auto ii) { b.cand_q[cn * NQ + IDX(ii)] = q[IDX(ii)]; });
            sfor<0, NV>([&](auto ii) { b.cand_v[cn * NV + IDX(ii)] = v[IDX(ii)]; b.cand_w[cn * NV + IDX(ii)] = warm[IDX(ii)]; });
            sfor<0, NU>([&](auto ii) { b.cand_u[cn * NU + IDX(ii)] = u[IDX(ii)]; });
        }
        if (cost) J = __dadd_rn(J, cost_eval<T>(*cost, q, v, u));
        step<T>(m, w, q, v, u, warm, qacc);   // mj_step (ilqr.h:128)
    }
    if constexpr (DIRECT) {
        // what ilqr_accept_kernel does for accept_always (acc = 0) and ilqr_commit_kernel's counter; setDInit(dArray[N]) is a no-op here:
        // knot N of the new nominal is the state this pass started from
        b.nom_J[i] = J;
        if (b.mu_i && b.mu_factor > 1.0) {
            double mu = b.mu_i[i] / b.mu_factor;
            if (mu < b.mu_min) mu = b.mu_min;
            b.mu_i[i] = mu;
        }
        b.accepted[i] = 0;
        if (Jtrace) Jtrace[slot + i] = J;
        if (acc_trace) acc_trace[slot + i] = 0;
        __syncthreads();
        if (threadIdx.x == 0 && b.iter_dev) {
            __threadfence();
            if (atomicAdd(b.ticket, 1) == (int)gridDim.x - 1) { *b.ticket = 0; *b.iter_dev += 1; }
        }
    } else
        b.cand_J[(size_t)a * ninst + i] = J;
}

// ------------------------------------------------------------------ ladder-order acceptance
// The accepted alpha is the FIRST one in ladder order whose cost beats the nominal's — exactly what sequential backtracking
// would pick.  Two small launches: the decision per instance, then one thread per (knot, instance) copies the accepted
// candidate over the nominal (time-major layout: consecutive threads touch consecutive knots' records).
template <class T>
__global__ void ilqr_accept_kernel(IlqrBuffers b, int accept_always, double* __restrict__ Jtrace, int* __restrict__ acc_trace) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.ninst) return;
    const int ninst = b.ninst;
    int acc = -1;
    const double Jprev = b.nom_J[i];
    double J = Jprev;
    if (accept_always) { acc = 0; J = b.cand_J[i]; }
    else
        for (int a = 0; a < b.nalpha; a++) {
            double Ja = b.cand_J[(size_t)a * ninst + i];
            if (Ja < Jprev) { acc = a; J = Ja; break; }
        }
    if (acc >= 0) b.nom_J[i] = J;
    if (b.mu_i && b.mu_factor > 1.0) {   // opt-in schedule: relax after an accepted step, stiffen after a rejected ladder
        double mu = b.mu_i[i];
        if (acc >= 0) { mu = mu / b.mu_factor; if (mu < b.mu_min) mu = b.mu_min; }
        else { mu = mu * b.mu_factor; if (mu > b.mu_max) mu = b.mu_max; }
        b.mu_i[i] = mu;
    }
    b.accepted[i] = acc;
    const size_t slot = (Jtrace || acc_trace) && b.iter_dev ? (size_t)(*b.iter_dev % b.trace_cap) * ninst : 0;
    if (Jtrace) Jtrace[slot + i] = J;
    if (acc_trace) acc_trace[slot + i] = acc;
}

template <class T>
__global__ void ilqr_commit_kernel(IlqrBuffers b) {
    constexpr int NQ = T::NQ, NV = T::NV, NU = T::NU;
    const int ninst = b.ninst, Tn = b.N + 1;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, T1 = (size_t)Tn * ninst;
    if (t >= T1) return;
    const int n = (int)(t / ninst), i = (int)(t - (size_t)n * ninst);
    const int acc = b.accepted[i];
    const size_t kn = t;
    if (acc >= 0) {
        const size_t cn = (size_t)acc * T1 + kn;
        for (int c = 0; c < NQ; c++) b.nom_q[kn * NQ + c] = b.cand_q[cn * NQ + c];
        for (int c = 0; c < NV; c++) { b.nom_v[kn * NV + c] = b.cand_v[cn * NV + c]; b.nom_w[kn * NV + c] = b.cand_w[cn * NV + c]; }
        for (int c = 0; c < NU; c++) b.nom_u[kn * NU + c] = b.cand_u[cn * NU + c];
    }
    if (t == 0 && b.iter_dev) *b.iter_dev += 1;   // (the accept kernel, which reads the counter, has finished)
    if (n == b.N) {   // setDInit(dArray[N]) (ilqr.h:183): the next pass starts from the nominal's first knot
        for (int c = 0; c < NQ; c++) b.init_q[(size_t)i * NQ + c] = b.nom_q[kn * NQ + c];
        for (int c = 0; c < NV; c++) { b.init_v[(size_t)i * NV + c] = b.nom_v[kn * NV + c]; b.init_w[(size_t)i * NV + c] = b.nom_w[kn * NV + c]; }
    }
}

// the same copy for a model whose sizes are only known at run time (the warp-cooperative engine)
__global__ void ilqr_commit_rt_kernel(IlqrBuffers b, int NQ, int NV, int NU) {
    const int ninst = b.ninst, Tn = b.N + 1;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, T1 = (size_t)Tn * ninst;
    if (t >= T1) return;
    const int n = (int)(t / ninst), i = (int)(t - (size_t)n * ninst);
    const int acc = b.accepted[i];
    const size_t kn = t;
    if (acc >= 0) {
        const size_t cn = (size_t)acc * T1 + kn;
        for (int c = 0; c < NQ; c++) b.nom_q[kn * NQ + c] = b.cand_q[cn * NQ + c];
        for (int c = 0; c < NV; c++) { b.nom_v[kn * NV + c] = b.cand_v[cn * NV + c]; b.nom_w[kn * NV + c] = b.cand_w[cn * NV + c]; }
        for (int c = 0; c < NU; c++) b.nom_u[kn * NU + c] = b.cand_u[cn * NU + c];
    }
    if (t == 0 && b.iter_dev) *b.iter_dev += 1;
    if (n == b.N) {
        for (int c = 0; c < NQ; c++) b.init_q[(size_t)i * NQ + c] = b.nom_q[kn * NQ + c];
        for (int c = 0; c < NV; c++) { b.init_v[(size_t)i * NV + c] = b.nom_v[kn * NV + c]; b.init_w[(size_t)i * NV + c] = b.nom_w[kn * NV + c]; }
    }
}

// ------------------------------------------------------------------ backward pass
// LANES threads cooperate on one instance; all matrices live in shared memory, column-major like the reference's
// Eigen objects.  Every lane loops over the output elements it owns; the nu x nu solve runs redundantly per lane
// on a private copy (nu <= 3 here), a pivoted L D L^T like the reference's .ldlt() (ilqr.h:167).
template <int NV, int NU, int LANES>
struct BackwardSmem {
    static constexpr int NX = 2 * NV;
    double V[NX * NX], A[NX * NX], Acl[NX * NX], T1[NX * NX], Vn[NX * NX];
    double B[NX * NU], VB[NX * NU], K[NU * NX], S[NU * NU];
    double v[NX], q[NX], c[NX], w[NX], wV[NX], rK[NX], vn[NX], r[NU], k[NU];
    // a CTA-wide group (LANES > 32) factorises S once per knot (ldlt_factor_warp); lane groups solve on private copies
    double Lf[LANES > 32 ? NU * NU : 1], Df[LANES > 32 ? NU : 1];
    int perm[LANES > 32 ? NU : 1];
};

template <int N>
DEV void ldlt_solve_small(const double* S, double* x) {  // S: N x N column-major symmetric (any sign); x in/out
    double W[N * N], L[N * N], D[N];
    int perm[N];
#pragma unroll
    for (int i = 0; i < N * N; i++) { W[i] = S[i]; L[i] = 0; }
#pragma unroll
    for (int i = 0; i < N; i++) perm[i] = i;
    for (int j = 0; j < N; j++) {
        int p = j;
        for (int i = j + 1; i < N; i++) if (fabs(W[i + i * N]) > fabs(W[p + p * N])) p = i;
        if (p != j) {
            for (int c = 0; c < N; c++) { double t = W[j + c * N]; W[j + c * N] = W[p + c * N]; W[p + c * N] = t; }
            for (int r = 0; r < N; r++) { double t = W[r + j * N]; W[r + j * N] = W[r + p * N]; W[r + p * N] = t; }
            for (int c = 0; c < j; c++) { double t = L[j + c * N]; L[j + c * N] = L[p + c * N]; L[p + c * N] = t; }
            int t = perm[j]; perm[j] = perm[p]; perm[p] = t;
        }
        D[j] = W[j + j * N];
        L[j + j * N] = 1;
        for (int i = j + 1; i < N; i++) L[i + j * N] = W[i + j * N] / D[j];
        for (int r = j + 1; r < N; r++) for (int c = j + 1; c < N; c++) W[r + c * N] -= L[r + j * N] * D[j] * L[c + j * N];
    }
    double y[N];
    for (int i = 0; i < N; i++) y[i] = x[perm[i]];
    for (int i = 0; i < N; i++) for (int c = 0; c < i; c++) y[i] -= L[i + c * N] * y[c];
    for (int i = 0; i < N; i++) y[i] /= D[i];
    for (int i = N - 1; i >= 0; i--) for (int c = i + 1; c < N; c++) y[i] -= L[c + i * N] * y[c];
    for (int i = 0; i < N; i++) x[perm[i]] = y[i];
}

// The same pivoted L D L^T for a large nu (the humanoid's 21 x 21), done ONCE per knot by one warp in shared memory — same pivot
// rule (largest |diagonal|, first among equals), same order of operations per element as ldlt_solve_small — and applied to the
// nx + 1 right-hand sides by one lane each.  W: N x N work copy of S (destroyed), L: unit lower factor, D, perm: outputs.
template <int N>
DEV void ldlt_factor_warp(double* W, double* L, double* D, int* perm, int lane) {
    for (int e = lane; e < N * N; e += 32) L[e] = 0;
    for (int e = lane; e < N; e += 32) perm[e] = e;
    __syncwarp();
    for (int j = 0; j < N; j++) {
        double best = -1;
        int p = N;
        for (int i = j + lane; i < N; i += 32) { const double a = fabs(W[i + i * N]); if (a > best) { best = a; p = i; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int op = __shfl_xor_sync(0xffffffffu, p, o);
            if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
        }
        if (p >= N) p = j;   // (no comparable diagonal left: NaNs propagate, indices stay in range)
        if (p != j) {   // (uniform over the warp)
            for (int c = lane; c < N; c += 32) { const double t = W[j + c * N]; W[j + c * N] = W[p + c * N]; W[p + c * N] = t; }
            __syncwarp();
            for (int r = lane; r < N; r += 32) { const double t = W[r + j * N]; W[r + j * N] = W[r + p * N]; W[r + p * N] = t; }
            for (int c = lane; c < j; c += 32) { const double t = L[j + c * N]; L[j + c * N] = L[p + c * N]; L[p + c * N] = t; }
            if (lane == 0) { const int t = perm[j]; perm[j] = perm[p]; perm[p] = t; }
            __syncwarp();
        }
        const double d = W[j + j * N];
        if (lane == 0) { D[j] = d; L[j + j * N] = 1; }
        for (int i = j + 1 + lane; i < N; i += 32) L[i + j * N] = W[i + j * N] / d;
        __syncwarp();
        const int m = N - j - 1;
        for (int e = lane; e < m * m; e += 32) {
            const int r = j + 1 + e % m, c = j + 1 + e / m;
            W[r + c * N] -= L[r + j * N] * d * L[c + j * N];
        }
        __syncwarp();
    }
}

template <int N>
DEV void ldlt_apply(const double* L, const double* D, const int* perm, double* x, int stride) {  // x[a * stride], in/out
    double y[N];
#pragma unroll
    for (int i = 0; i < N; i++) y[i] = x[perm[i] * stride];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int c = 0; c < i; c++) y[i] -= L[i + c * N] * y[c];
#pragma unroll
    for (int i = 0; i < N; i++) y[i] /= D[i];
#pragma unroll
    for (int i = N - 1; i >= 0; i--)
#pragma unroll
        for (int c = i + 1; c < N; c++) y[i] -= L[c + i * N] * y[c];
#pragma unroll
    for (int i = 0; i < N; i++) x[perm[i] * stride] = y[i];
}

// Small state vectors (2 nv <= 4: the inverted pendulum): ONE THREAD per instance, every matrix in registers, all loops unrolled.
// Same algebra, same order of operations as the lane-group kernel below (which stays the path for larger nx); what goes is the
// dozen shared-memory phases and group barriers per knot that made a 4 x 4 Riccati step cost ~10 K cycles.  The next knot's
// deriv block and nominal are fetched while the current knot is processed.
template <int NV, int NU>
__global__ void __launch_bounds__(32) ilqr_backward_small_kernel(IlqrBuffers b, double dt) {
    constexpr int NX = 2 * NV, ND = NV * (2 * NV + NU) + 2 * NV + NU, NQ = NV;
    static_assert(NX <= 4 && NU <= 2, "register-resident Riccati step is for tiny models");
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.ninst) return;
    const int ninst = b.ninst;
    const double mu = b.mu_i ? b.mu_i[i] : b.mu;
#define CMX(M, r, c, rows) ((M)[(r) + (c) * (rows)])
    double V[NX * NX], v[NX];
    {   // initV (ilqr.h:100-107): v = dgdx at knot 0, V = v' v
        const double* d0 = b.deriv + ((size_t)0 * ninst + i) * ND + 2 * NV * NV + NV * NU;
#pragma unroll
        for (int e = 0; e < NX; e++) v[e] = d0[e];
#pragma unroll
        for (int e = 0; e < NX * NX; e++) V[e] = v[e % NX] * v[e / NX];
    }
    double dnext[ND], xn[NX], xp[NX];   // deriv block of knot n, nominal state of knots n and n-1
    auto fetch = [&](int n) {
        const size_t kn = (size_t)n * ninst + i;
        const double* d = b.deriv + kn * ND;
#pragma unroll
        for (int e = 0; e < ND; e++) dnext[e] = d[e];
#pragma unroll
        for (int e = 0; e < NV; e++) { xn[e] = b.nom_q[kn * NQ + e]; xn[NV + e] = b.nom_v[kn * NV + e]; }
    };
    {
        const size_t k0 = (size_t)0 * ninst + i;
#pragma unroll
        for (int e = 0; e < NV; e++) { xp[e] = b.nom_q[k0 * NQ + e]; xp[NV + e] = b.nom_v[k0 * NV + e]; }
    }
    fetch(1);
    for (int n = 1; n <= b.N; n++) {
        const size_t kn = (size_t)n * ninst + i;
        double deriv[ND], c[NX];
#pragma unroll
        for (int e = 0; e < ND; e++) deriv[e] = dnext[e];
#pragma unroll
        for (int e = 0; e < NX; e++) { c[e] = xp[e] - xn[e]; xp[e] = xn[e]; }
        if (n < b.N) fetch(n + 1);
        double Vn[NX * NX], A[NX * NX], B[NX * NU], q[NX], r[NU];
#pragma unroll
        for (int e = 0; e < NX * NX; e++) { const int rr = e % NX, cc = e / NX; Vn[e] = (CMX(V, rr, cc, NX) + CMX(V, cc, rr, NX)) / 2; }
#pragma unroll
        for (int e = 0; e < NX * NX; e++) {
            const int rr = e % NX, cc = e / NX;
            double a;
            if (rr < NV) a = (cc == rr ? 1.0 : 0.0) + (cc == NV + rr ? dt : 0.0);
            else if (cc < NV) a = deriv[b.corrected ? cc + (rr - NV) * NV : (rr - NV) + cc * NV] * dt;
            else a = ((rr - NV) == (cc - NV) ? 1.0 : 0.0) + deriv[NV * NV + (b.corrected ? (cc - NV) + (rr - NV) * NV : (rr - NV) + (cc - NV) * NV)] * dt;
            A[e] = a;
        }
#pragma unroll
        for (int e = 0; e < NX * NU; e++) {
            const int rr = e % NX, cc = e / NX;
            B[e] = rr < NV ? 0.0 : deriv[2 * NV * NV + (b.corrected ? cc + (rr - NV) * NU : (rr - NV) + cc * NV)] * dt;
        }
#pragma unroll
        for (int e = 0; e < NX; e++) q[e] = deriv[2 * NV * NV + NV * NU + e];
#pragma unroll
        for (int e = 0; e < NU; e++) r[e] = deriv[2 * NV * NV + NV * NU + 2 * NV + e];
#pragma unroll
        for (int e = 0; e < NX * NX; e++) V[e] = Vn[e] + ((e % NX) == (e / NX) ? mu : 0.0);   // ilqr.h:166; never removed (quirk Q3)
        double VB[NX * NU], T1[NX * NX], w[NX];
#pragma unroll
        for (int e = 0; e < NX * NU; e++) {
            const int rr = e % NX, cc = e / NX;
            double t = 0;
#pragma unroll
            for (int x = 0; x < NX; x++) t += CMX(V, rr, x, NX) * CMX(B, x, cc, NX);
            VB[e] = t;
        }
#pragma unroll
        for (int e = 0; e < NX * NX; e++) {
            const int rr = e % NX, cc = e / NX;
            double t = 0;
#pragma unroll
            for (int x = 0; x < NX; x++) t += CMX(V, rr, x, NX) * CMX(A, x, cc, NX);
            T1[e] = t;
        }
#pragma unroll
        for (int e = 0; e < NX; e++) {
            double t = 0;
#pragma unroll
            for (int x = 0; x < NX; x++) t += CMX(V, e, x, NX) * c[x];
            w[e] = v[e] + 2 * t;
        }
        double S[NU * NU], K[NU * NX], k[NU];
#pragma unroll
        for (int e = 0; e < NU * NU; e++) {
            const int a = e % NU, cc = e / NU;
            double t = 0;
#pragma unroll
            for (int x = 0; x < NX; x++) t += CMX(B, x, a, NX) * CMX(VB, x, cc, NX);
            S[e] = -2 * t - 2 * r[a] * r[cc];
        }
#pragma unroll
        for (int e = 0; e < NU * NX; e++) {
            const int a = e % NU, cc = e / NU;
            double t = 0;
#pragma unroll
            for (int x = 0; x < NX; x++) t += CMX(B, x, a, NX) * CMX(T1, x, cc, NX);
            K[e] = 2 * t;
        }
#pragma unroll
        for (int e = 0; e < NU; e++) {
            double t = 0;
#pragma unroll
            for (int x = 0; x < NX; x++) t += CMX(B, x, e, NX) * w[x];
            k[e] = t + r[e];
        }
#pragma unroll
        for (int cc = 0; cc < NX + 1; cc++) {   // K[n] = S^-1 rhsK (column by column), k[n] = S^-1 rhsk
            double x[NU];
#pragma unroll
            for (int a = 0; a < NU; a++) x[a] = cc < NX ? K[a + cc * NU] : k[a];
            ldlt_solve_small<NU>(S, x);
#pragma unroll
            for (int a = 0; a < NU; a++) { if (cc < NX) K[a + cc * NU] = x[a]; else k[a] = x[a]; }
        }
#pragma unroll
        for (int e = 0; e < NU * NX; e++) b.K[kn * NU * NX + e] = K[e];
#pragma unroll
        for (int e = 0; e < NU; e++) b.k[kn * NU + e] = k[e];
        double Acl[NX * NX], rK[NX];
#pragma unroll
        for (int e = 0; e < NX * NX; e++) {
            const int rr = e % NX, cc = e / NX;
            double t = A[e];
#pragma unroll
            for (int a = 0; a < NU; a++) t += CMX(B, rr, a, NX) * K[a + cc * NU];
            Acl[e] = t;
        }
#pragma unroll
        for (int e = 0; e < NX; e++) {
            double t = 0;
#pragma unroll
            for (int a = 0; a < NU; a++) t += r[a] * K[a + e * NU];
            rK[e] = t;
            double u = c[e];
#pragma unroll
            for (int a = 0; a < NU; a++) u += CMX(B, e, a, NX) * k[a];
            w[e] = u;
        }
#pragma unroll
        for (int e = 0; e < NX * NX; e++) {   // T1 = V Acl
            const int rr = e % NX, cc = e / NX;
            double t = 0;
#pragma unroll
            for (int x = 0; x < NX; x++) t += CMX(V, rr, x, NX) * CMX(Acl, x, cc, NX);
            T1[e] = t;
        }
#pragma unroll
        for (int e = 0; e < NX * NX; e++) {   // Vn = Acl' V Acl + q'q + (rK)'(rK)   (ilqr.h:173)
            const int rr = e % NX, cc = e / NX;
            double t = 0;
#pragma unroll
            for (int x = 0; x < NX; x++) t += CMX(Acl, x, rr, NX) * CMX(T1, x, cc, NX);
            Vn[e] = t + q[rr] * q[cc] + rK[rr] * rK[cc];
        }
        double wV[NX], vn[NX];
#pragma unroll
        for (int e = 0; e < NX * NX; e++) V[e] = Vn[e];
#pragma unroll
        for (int e = 0; e < NX; e++) {        // wV = (k'B' + c') V_new      (quirk Q4)
            double t = 0;
#pragma unroll
            for (int x = 0; x < NX; x++) t += w[x] * CMX(Vn, x, e, NX);
            wV[e] = t;
        }
        double kr = 0;
#pragma unroll
        for (int a = 0; a < NU; a++) kr += k[a] * r[a];
#pragma unroll
        for (int e = 0; e < NX; e++) {        // v = 2 wV Acl + v Acl + q + 2 (k'r') rK   (ilqr.h:174)
            double t1 = 0, t2 = 0;
#pragma unroll
            for (int x = 0; x < NX; x++) { t1 += wV[x] * CMX(Acl, x, e, NX); t2 += v[x] * CMX(Acl, x, e, NX); }
            vn[e] = 2 * t1 + t2 + q[e] + 2 * kr * rK[e];
        }
#pragma unroll
        for (int e = 0; e < NX; e++) v[e] = vn[e];
    }
#pragma unroll
    for (int e = 0; e < NX * NX; e++) b.V[(size_t)i * NX * NX + e] = V[e];
#pragma unroll
    for (int e = 0; e < NX; e++) b.v[(size_t)i * NX + e] = v[e];
#undef CMX
}

// (Round 2 also measured FOUR LANES per instance on this sweep — lane c owning column c of V / K, row c of the products with V, register
//  transposes and all-gathers by shuffle, every output element still one lane's sum in the same order: bit-identical results, a
//  quarter of the multiply-adds per lane, and SLOWER: 4096 pendulum problems 0.215 -> 0.248 ms per batch iteration.  The 32 shuffles
//  and the select chains per knot lengthen the dependent chain by more than the shorter sums save.  Not kept.)
template <int NV, int NU, int LANES, int GROUPS>
__global__ void __launch_bounds__(LANES * GROUPS) ilqr_backward_kernel(IlqrBuffers b, double dt) {
    constexpr int NX = 2 * NV, ND = NV * (2 * NV + NU) + 2 * NV + NU, NQ = NV;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef BackwardSmem<NV, NU, LANES> SM;
    SM* all = reinterpret_cast<SM*>(smem_raw);
    const int g = threadIdx.x / LANES, lane = threadIdx.x % LANES;
    const int i = blockIdx.x * GROUPS + g;
    const bool live = i < b.ninst;
    SM& s = all[g];
    const int ninst = b.ninst;
    const double mu = (b.mu_i && live) ? b.mu_i[i] : b.mu;
    // groups are whole warps or aligned sub-warps (sync the lanes of this group only) — or, for large state vectors, the whole
    // CTA works on one instance (LANES > 32, GROUPS = 1: block barrier)
    static_assert(LANES <= 32 || GROUPS == 1, "a group wider than a warp is the whole CTA");
    const unsigned gmask = LANES >= 32 ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x % 32) / LANES * LANES));
    auto gsync = [&]() { if constexpr (LANES > 32) __syncthreads(); else __syncwarp(gmask); };
#define CMX(M, r, c, rows) ((M)[(r) + (c) * (rows)])
    if (live) {
        // initV (ilqr.h:100-107): v = dgdx at knot 0, V = v' v
        const double* d0 = b.deriv + ((size_t)0 * ninst + i) * ND + 2 * NV * NV + NV * NU;
        for (int e = lane; e < NX; e += LANES) s.v[e] = d0[e];
    }
    gsync();
    if (live)
        for (int e = lane; e < NX * NX; e += LANES) s.V[e] = s.v[e % NX] * s.v[e / NX];
    gsync();
    for (int n = 1; n <= b.N; n++) {
        const size_t kn = (size_t)n * ninst + i, kp = (size_t)(n - 1) * ninst + i;
        if (live) {
            const double* deriv = b.deriv + kn * ND;
            // V <- (V + V')/2 into Vn, then back (ilqr.h:150)
            for (int e = lane; e < NX * NX; e += LANES) { int r = e % NX, c = e / NX; s.Vn[e] = (CMX(s.V, r, c, NX) + CMX(s.V, c, r, NX)) / 2; }
            // A, B through the column-major views of the row-major deriv blocks (differentiator.h:68-71,89-92; quirk Q1)
            for (int e = lane; e < NX * NX; e += LANES) {
                int r = e % NX, c = e / NX;
                double a;
                if (r < NV) a = (c == r ? 1.0 : 0.0) + (c == NV + r ? dt : 0.0);
                else if (c < NV) a = deriv[b.corrected ? c + (r - NV) * NV : (r - NV) + c * NV] * dt;
                else a = ((r - NV) == (c - NV) ? 1.0 : 0.0) + deriv[NV * NV + (b.corrected ? (c - NV) + (r - NV) * NV : (r - NV) + (c - NV) * NV)] * dt;
                s.A[e] = a;
            }
            for (int e = lane; e < NX * NU; e += LANES) {
                int r = e % NX, c = e / NX;
                s.B[e] = r < NV ? 0.0 : deriv[2 * NV * NV + (b.corrected ? c + (r - NV) * NU : (r - NV) + c * NV)] * dt;
            }
            for (int e = lane; e < NX; e += LANES) {
                s.q[e] = deriv[2 * NV * NV + NV * NU + e];
                if (b.cdiff) s.c[e] = b.cdiff[kn * NX + e];   // tangent-space difference (models with quaternions)
                else s.c[e] = e < NV ? b.nom_q[kp * NQ + e] - b.nom_q[kn * NQ + e] : b.nom_v[kp * NV + (e - NV)] - b.nom_v[kn * NV + (e - NV)];
            }
            for (int e = lane; e < NU; e += LANES) s.r[e] = deriv[2 * NV * NV + NV * NU + 2 * NV + e];
        }
        gsync();
        if (live)  // V = sym(V) + mu I   (ilqr.h:166; never removed — quirk Q3)
            for (int e = lane; e < NX * NX; e += LANES) s.V[e] = s.Vn[e] + ((e % NX) == (e / NX) ? mu : 0.0);
        gsync();
        if (live) {
            for (int e = lane; e < NX * NU; e += LANES) {  // VB = V B
                int r = e % NX, c = e / NX;
                double t = 0;
                for (int x = 0; x < NX; x++) t += CMX(s.V, r, x, NX) * CMX(s.B, x, c, NX);
                s.VB[e] = t;
            }
            for (int e = lane; e < NX * NX; e += LANES) {  // T1 = V A
                int r = e % NX, c = e / NX;
                double t = 0;
                for (int x = 0; x < NX; x++) t += CMX(s.V, r, x, NX) * CMX(s.A, x, c, NX);
                s.T1[e] = t;
            }
            for (int e = lane; e < NX; e += LANES) {       // w = v' + 2 V c
                double t = 0;
                for (int x = 0; x < NX; x++) t += CMX(s.V, e, x, NX) * s.c[x];
                s.w[e] = s.v[e] + 2 * t;
            }
        }
        gsync();
        if (live) {
            for (int e = lane; e < NU * NU; e += LANES) {  // S = -2 B'VB - 2 r'r
                int a = e % NU, c = e / NU;
                double t = 0;
                for (int x = 0; x < NX; x++) t += CMX(s.B, x, a, NX) * CMX(s.VB, x, c, NX);
                s.S[e] = -2 * t - 2 * s.r[a] * s.r[c];
            }
            for (int e = lane; e < NU * NX; e += LANES) {  // rhsK = 2 B' V A
                int a = e % NU, c = e / NU;
                double t = 0;
                for (int x = 0; x < NX; x++) t += CMX(s.B, x, a, NX) * CMX(s.T1, x, c, NX);
                s.K[e] = 2 * t;
            }
            for (int e = lane; e < NU; e += LANES) {       // rhsk = B'(v' + 2Vc) + r'
                double t = 0;
                for (int x = 0; x < NX; x++) t += CMX(s.B, x, e, NX) * s.w[x];
                s.k[e] = t + s.r[e];
            }
        }
        gsync();
        if constexpr (LANES > 32) {
            if (live && lane < 32) ldlt_factor_warp<NU>(s.S, s.Lf, s.Df, s.perm, lane);
            gsync();
            if (live)
                for (int c = lane; c < NX + 1; c += LANES) ldlt_apply<NU>(s.Lf, s.Df, s.perm, c < NX ? s.K + c * NU : s.k, 1);
        } else if (live) {  // K[n] = S^-1 rhsK (column by column), k[n] = S^-1 rhsk
            for (int c = lane; c < NX + 1; c += LANES) {
                double x[NU];
                if (c < NX) { for (int a = 0; a < NU; a++) x[a] = s.K[a + c * NU]; }
                else for (int a = 0; a < NU; a++) x[a] = s.k[a];
                ldlt_solve_small<NU>(s.S, x);
                if (c < NX) { for (int a = 0; a < NU; a++) s.K[a + c * NU] = x[a]; }
                else for (int a = 0; a < NU; a++) s.k[a] = x[a];
            }
        }
        gsync();
        if (live) {
            for (int e = lane; e < NU * NX; e += LANES) b.K[kn * NU * NX + e] = s.K[e];
            for (int e = lane; e < NU; e += LANES) b.k[kn * NU + e] = s.k[e];
            for (int e = lane; e < NX * NX; e += LANES) {  // Acl = A + B K
                int r = e % NX, c = e / NX;
                double t = s.A[e];
                for (int a = 0; a < NU; a++) t += CMX(s.B, r, a, NX) * s.K[a + c * NU];
                s.Acl[e] = t;
            }
            for (int e = lane; e < NX; e += LANES) {       // rK = r K ; wB = c + B k
                double t = 0;
                for (int a = 0; a < NU; a++) t += s.r[a] * s.K[a + e * NU];
                s.rK[e] = t;
                double u = s.c[e];
                for (int a = 0; a < NU; a++) u += CMX(s.B, e, a, NX) * s.k[a];
                s.w[e] = u;
            }
        }
        gsync();
        if (live)
            for (int e = lane; e < NX * NX; e += LANES) {  // T1 = V Acl
                int r = e % NX, c = e / NX;
                double t = 0;
                for (int x = 0; x < NX; x++) t += CMX(s.V, r, x, NX) * CMX(s.Acl, x, c, NX);
                s.T1[e] = t;
                // wide groups: Acl' into A (free from here on) so that the next product reads consecutive words per lane instead
                // of a stride of NX doubles (NX = 54: 4-way bank conflicts on every load)
                if constexpr (LANES > 32) CMX(s.A, c, r, NX) = s.Acl[e];
            }
        gsync();
        if (live)
            for (int e = lane; e < NX * NX; e += LANES) {  // Vn = Acl' V Acl + q'q + (rK)'(rK)   (ilqr.h:173)
                int r = e % NX, c = e / NX;
                double t = 0;
                if constexpr (LANES > 32) { for (int x = 0; x < NX; x++) t += CMX(s.A, r, x, NX) * CMX(s.T1, x, c, NX); }
                else for (int x = 0; x < NX; x++) t += CMX(s.Acl, x, r, NX) * CMX(s.T1, x, c, NX);
                s.Vn[e] = t + s.q[r] * s.q[c] + s.rK[r] * s.rK[c];
            }
        gsync();
        if (live) {
            for (int e = lane; e < NX * NX; e += LANES) s.V[e] = s.Vn[e];
            for (int e = lane; e < NX; e += LANES) {       // wV = (k'B' + c') V_new      (quirk Q4)
                double t = 0;
                for (int x = 0; x < NX; x++) t += s.w[x] * CMX(s.Vn, x, e, NX);
                s.wV[e] = t;
            }
        }
        gsync();
        if (live) {
            double kr = 0;
            for (int a = 0; a < NU; a++) kr += s.k[a] * s.r[a];
            for (int e = lane; e < NX; e += LANES) {       // v = 2 wV Acl + v Acl + q + 2 (k'r') rK   (ilqr.h:174)
                double t1 = 0, t2 = 0;
                for (int x = 0; x < NX; x++) { t1 += s.wV[x] * CMX(s.Acl, x, e, NX); t2 += s.v[x] * CMX(s.Acl, x, e, NX); }
                s.vn[e] = 2 * t1 + t2 + s.q[e] + 2 * kr * s.rK[e];
            }
        }
        gsync();
        if (live)
            for (int e = lane; e < NX; e += LANES) s.v[e] = s.vn[e];
        gsync();
    }
    if (live) {
        for (int e = lane; e < NX * NX; e += LANES) b.V[(size_t)i * NX * NX + e] = s.V[e];
        for (int e = lane; e < NX; e += LANES) b.v[(size_t)i * NX + e] = s.v[e];
    }
#undef CMX
}

}  // namespace ilqg
