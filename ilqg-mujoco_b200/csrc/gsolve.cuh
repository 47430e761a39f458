// gsolve.cuh — FD of a latency-bound batch (one horizon of <= a few thousand knots, a single trajectory) with SEVERAL LANES PER SOLVE.
//
// A pass over such a batch is one dependent chain: build the knot's problems -> centre solve(s) (mjderivative.cpp:64-68) -> the perturbed
// solves warm-started from it (:75,91).  Measured on the T = 1000 hopper horizon (tools/prof_fused_diag.py, B200): the builds take
// 21-31 K cycles, and then the SOLVES are the chain — a thread walking its <= 41 constraint rows in local memory spends ~1000 cycles per
// row and Newton iteration (every pass over the rows is a string of dependent local loads and fp64 FMAs with nothing to overlap), 16-19 K
// cycles per iteration in stance, and the slowest of a knot's 30 perturbed solves (2-4 iterations when the +-eps stencil moves the active
// set) sets the pass: 84 K cycles at the tail.  Lanes are plentiful in this regime (1000 knots x 31 evaluations = 1000 warps on 148 SMs),
// so the rows of ONE solve are spread over GW lanes:
//   fd_build_kernel   the pipeline up to the constraint problem, one thread per evaluation exactly as in the one-launch kernel (31 lanes
//                     per hopper knot, registers for the tree), but no solver: every lane leaves its problem — M, qfrc_smooth,
//                     qacc_smooth, the rows (J | D | aref) — in a lane-interleaved record in HBM / L2 ([element][lane]: coalesced stores);
//   fd_solve_kernel   a lean second kernel behind it (programmatic dependent launch), one CTA per knot: warp 0 solves the centre problem
//                     with all 32 lanes (row r on lane r), then the CTA's 4 warps solve the 30 perturbed problems GW = 4 lanes each,
//                     take the central differences (the +-eps problems of a column sit in neighbouring lane groups: one shuffle) and
//                     write the knot's deriv block out coalesced.
// solve_rows() below is solve() of dyn.cuh re-expressed for a lane group: the passes over the rows become one or a few rows per lane
// and a butterfly sum over the group (gradient, cost and the Hessian's row term in ONE round of reductions per Newton iteration, two sums
// and a ballot per line-search trial); the dense nv x nv part (products with M, Cholesky, triangular solves) is done by every lane on
// bit-identical operands, which keeps the lanes of a group on the same decisions without any broadcast.  Same algorithm and exits as
// solve() (active sets as bit masks, exact-optimum test, MuJoCo's termination rule); the sums associate differently, so results agree
// with the thread-per-solve kernels to round-off (1e-11 relative on the deriv blocks), not bit for bit.
// The groups of a warp run in LOCKSTEP (full-mask shuffles, finished groups idle): divergence between groups would serialise them.
#pragma once
#include "dyn.cuh"

namespace ilqg {

// ------------------------------------------------------------------ the problem record
// Per knot: NE elements x 32 lanes (lane l = evaluation l of the knot: 2 x column + sign, the centre last), element-major.
template <class T>
struct FdRecord {
    static constexpr int NV = T::NV, NU = T::NU, NT = NV * (NV + 1) / 2, ME = nz(T::MAXEFC);
    static constexpr int NCOL = 2 * NV + NU, G = 2 * NCOL, GL = G + 1;
    static constexpr bool OK = GL <= 32 && GL > 16 && T::MAXEFC > 0;   // a knot per warp in the build kernel, rows worth spreading
    static constexpr int E_NE = 0, E_DCOST = 1, E_M = 2, E_FS = E_M + NT, E_AS = E_FS + NV, E_ROWS = (E_AS + NV + 1) & ~1;
    static constexpr int RS = NV + 2;                      // J[nv], D, aref per row
    static constexpr int NE = E_ROWS + ME * RS;            // elements per evaluation
    static constexpr size_t DOUBLES = (size_t)NE * 32;     // per knot
};

template <int GW>
DEV double gsum(double x) {
#pragma unroll
    for (int o = GW / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
template <int GW>
DEV bool gall(bool p, int lane) {   // true on every lane of a group iff p holds on all of them (executed by the whole warp)
    const unsigned b = __ballot_sync(0xffffffffu, p);
    if constexpr (GW == 32) return b == 0xffffffffu;
    else {
        const unsigned gm = ((1u << GW) - 1u) << (lane & ~(GW - 1));
        return (b & gm) == gm;
    }
}

// rows of one lane: row r of the problem lives in slot r / GW of lane r % GW of the group
template <class T, int NSL>
struct LaneRows {
    double J[NSL][T::NV], D[NSL], jar[NSL], jv[NSL];
    int ns;   // rows this lane holds
};

// Newton solve by lane groups; called by all 32 lanes of the warp, converged.  `live`: the group has a problem (with rows) to solve.
// In: R.J / R.D, R.jar = -aref, M (NT doubles, the same for the lanes of a group; any address space), fs, as, warm (the warm start).
// Out (group-uniform): qacc = warm = the solution, iters, exact, and R.jar = the rows' residuals there.
template <class T, int GW, int NSL, class MP>
DEV void solve_rows(const DevModel<T>& m, bool live, LaneRows<T, NSL>& R, MP M, const double (&fs)[T::NV], const double (&as)[T::NV],
                    double (&warm)[T::NV], double (&qacc)[T::NV], int maxiter, double tol, int& iters, int& exact) {
    constexpr int NV = T::NV, NT = NV * (NV + 1) / 2;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool STATIC = NSL <= 2;   // slots in registers, loops unrolled
    const int lane = threadIdx.x & 31;
    const int nsm = STATIC ? NSL : __reduce_max_sync(FULL, live ? R.ns : 0);   // the warp's row loops run in step
#define GS_ROWS(s) _Pragma("unroll") for (int s = 0; s < (STATIC ? NSL : nsm); s++) if (s < R.ns)
#define GS_ROWS_DYN(s) _Pragma("unroll 1") for (int s = 0; s < nsm; s++) if (s < R.ns)
    const double scale = 1.0 / (m.meaninertia * (NV > 1 ? NV : 1));
    double Ma[NV], grad[NV], search[NV], Mv[NV];
    iters = 0;
    exact = 0;
    {
        // start from the better of the warm start and qacc_smooth (one pass evaluates both)
        double cw = 0, cs = 0;
        auto body = [&](int s) {
            double a = R.jar[s], b = a;
            sfor<0, NV>([&](auto ii) { a += R.J[s][IDX(ii)] * warm[IDX(ii)]; b += R.J[s][IDX(ii)] * as[IDX(ii)]; });
            R.jar[s] = a;
            R.jv[s] = b;
            if (a < 0) cw += 0.5 * R.D[s] * a * a;
            if (b < 0) cs += 0.5 * R.D[s] * b * b;
        };
        if constexpr (STATIC) { GS_ROWS(s) body(s); } else { GS_ROWS_DYN(s) body(s); }
        cw = gsum<GW>(cw);
        cs = gsum<GW>(cs);
        sfor<0, NV>([&](auto ii) {
            constexpr int i = IDX(ii);
            double s = 0;
            sfor<0, NV>([&](auto kk) { s += M[tri(i, IDX(kk))] * warm[IDX(kk)]; });
            Ma[i] = s;
            cw += 0.5 * (s - fs[i]) * (warm[i] - as[i]);
        });
        if (cw < cs) sfor<0, NV>([&](auto ii) { qacc[IDX(ii)] = warm[IDX(ii)]; });
        else {
            sfor<0, NV>([&](auto ii) { qacc[IDX(ii)] = as[IDX(ii)]; });
            sfor<0, NV>([&](auto ii) {
                constexpr int i = IDX(ii);
                double s = 0;
                sfor<0, NV>([&](auto kk) { s += M[tri(i, IDX(kk))] * qacc[IDX(kk)]; });
                Ma[i] = s;
            });
            if constexpr (STATIC) { GS_ROWS(s) R.jar[s] = R.jv[s]; } else { GS_ROWS_DYN(s) R.jar[s] = R.jv[s]; }
        }
    }
    double cost = 0, old = 0;
    int iter = 0;
    bool done = !live;
    for (int round = 0; round <= maxiter; round++) {
        // ---- cost, gradient, Hessian factor and Newton direction: one round of reductions over [fc | cost | row term of H]
        double red[NV + 1 + NT];
        sfor<0, NV + 1 + NT>([&](auto xx) { red[IDX(xx)] = 0; });
        unsigned act = 0;   // this lane's active slots
        {
            auto body = [&](int s) {
                const double jar = R.jar[s];
                if (jar < 0) {
                    act |= 1u << s;
                    const double D = R.D[s], f = -D * jar;
                    red[NV] += 0.5 * D * jar * jar;
                    sfor<0, NV>([&](auto ii) {
                        constexpr int i = IDX(ii);
                        const double Ji = R.J[s][i], t = D * Ji;
                        red[i] += Ji * f;
                        sfor<0, i + 1>([&](auto kk) { red[NV + 1 + tri(i, IDX(kk))] += t * R.J[s][IDX(kk)]; });
                    });
                }
            };
            if constexpr (STATIC) { GS_ROWS(s) body(s); } else { GS_ROWS_DYN(s) body(s); }
        }
        sfor<0, NV + 1 + NT>([&](auto xx) { red[IDX(xx)] = gsum<GW>(red[IDX(xx)]); });
        bool stop = false;
        {
            double H[NT], Lh[NT];
            sfor<0, NT>([&](auto tt) { H[IDX(tt)] = M[IDX(tt)] + red[NV + 1 + IDX(tt)]; });
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
            if (scale * (old - cost) < tol || scale * sqrt(gn) < tol) stop = true;
        }
        if (iter >= maxiter) stop = true;
        // ---- exact linesearch: root of the piecewise-linear derivative along `search`
        double g1 = 0, g2 = 0;
        sfor<0, NV>([&](auto ii) {
            constexpr int i = IDX(ii);
            double s = 0;
            sfor<0, NV>([&](auto kk) { s += M[tri(i, IDX(kk))] * search[IDX(kk)]; });
            Mv[i] = s;
        });
        sfor<0, NV>([&](auto ii) { constexpr int i = IDX(ii); g1 += search[i] * (Ma[i] - fs[i]); g2 += search[i] * Mv[i]; });
        double p1 = 0, p2 = 0;
        {
            auto body = [&](int s) {
                double x = 0;
                sfor<0, NV>([&](auto ii) { x += R.J[s][IDX(ii)] * search[IDX(ii)]; });
                R.jv[s] = x;
                if ((act >> s) & 1) {
                    const double t = R.D[s] * x;
                    p1 += t * R.jar[s];
                    p2 += t * x;
                }
            };
            if constexpr (STATIC) { GS_ROWS(s) body(s); } else { GS_ROWS_DYN(s) body(s); }
        }
        double d1 = g1 + gsum<GW>(p1), d2 = g2 + gsum<GW>(p2);
        if (d1 >= 0 || d2 < ILQG_MINVAL) stop = true;  // not a descent direction: converged to round-off
        done = done || stop;
        double alpha = 0, lo = 0, hi = CUDART_INF;
        unsigned cur = act, reached = act;
        bool lsdone = done;
        for (int it = 0; it < m.ls_iterations; it++) {
            if (__all_sync(FULL, lsdone)) break;
            if (d1 < 0) lo = alpha; else hi = alpha;
            double an = alpha - d1 / d2;
            if (!(an > lo && an < hi)) an = isinf(hi) ? 2 * alpha + 1 : 0.5 * (lo + hi);
            double q1 = 0, q2 = 0;
            unsigned mk = 0;
            auto body = [&](int s) {
                const double jv = R.jv[s], x = R.jar[s] + an * jv;
                if (x < 0) {
                    const double t = R.D[s] * jv;
                    q1 += t * x;
                    q2 += t * jv;
                    mk |= 1u << s;
                }
            };
            if constexpr (STATIC) { GS_ROWS(s) body(s); } else { GS_ROWS_DYN(s) body(s); }
            const double e1 = g1 + g2 * an + gsum<GW>(q1), e2 = g2 + gsum<GW>(q2);
            const bool same = gall<GW>(mk == cur, lane);   // the step stayed inside one linear piece of the derivative: `an` is its root
            if (!lsdone) {
                alpha = an;
                d1 = e1;
                d2 = e2;
                cur = mk;
                reached = mk;
                if (same || d1 == 0 || d2 < ILQG_MINVAL) lsdone = true;
            }
        }
        const bool ex = gall<GW>(reached == act, lane);  // exact optimum: the minimiser lies in the piece the Hessian was built for
        if (!done) {
            if (alpha == 0) done = true;
            else {
                sfor<0, NV>([&](auto ii) { constexpr int i = IDX(ii); qacc[i] += alpha * search[i]; Ma[i] += alpha * Mv[i]; });
                if constexpr (STATIC) { GS_ROWS(s) R.jar[s] += alpha * R.jv[s]; } else { GS_ROWS_DYN(s) R.jar[s] += alpha * R.jv[s]; }
                old = cost;
                iter++;
                if (ex) { exact = 1; done = true; }
            }
        }
        if (__all_sync(FULL, done)) break;
    }
#undef GS_ROWS
#undef GS_ROWS_DYN
    iters = iter;
    sfor<0, NV>([&](auto ii) { warm[IDX(ii)] = qacc[IDX(ii)]; });
}

// ------------------------------------------------------------------ kernel 1: build and export
// Same lane layout as the one-launch kernel: lane l < G of a knot's warp evaluates column l / 2 at +eps (l even) or -eps, lane G the knot.
template <class T>
__global__ void __launch_bounds__(256, 1) fd_build_kernel(const __grid_constant__ DevModel<T> m, int nknots, const double* __restrict__ qpos,
                                                          const double* __restrict__ qvel, const double* __restrict__ ctrl,
                                                          const ilqg_cost* __restrict__ cost, double eps, double* __restrict__ rec) {
    using S = FdRecord<T>;
    constexpr int NV = T::NV, NU = T::NU, NQ = T::NQ, WARPS = 8;
    asm volatile("griddepcontrol.launch_dependents;");   // the solve kernel may become resident now; it waits for this grid's records
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * WARPS + wib;
    const bool valid = k < nknots && lane < S::GL;
    const bool is_center = lane == S::G;
    const int col = lane >> 1;
    const double se = (lane & 1) ? -eps : eps;
    const int kk = k < nknots ? k : nknots - 1;   // idle warps / lanes evaluate a clamped knot, stores masked (stage barriers)
    double q[NQ], v[NV], u[nz(NU)];
    load_knot<T>(kk, qpos, qvel, ctrl, q, v, u);
    double c0 = 0, dcost = 0;
    if (cost) c0 = cost_eval<T>(*cost, q, v, u);
    if (!is_center) {   // perturb this lane's input (ctrl: mjderivative.cpp:85,99; qvel: :117,130; qpos: :164-169,187-192)
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
        if (cost && !(lane & 1)) dcost = __ddiv_rn(__dsub_rn(cost_eval<T>(*cost, q, v, u), c0), eps);
    }
    Work<T> w;
    build_problem<T, true>(m, q, v, u, w);
    if (!valid) return;
    double* r = rec + (size_t)k * S::DOUBLES + lane;   // element e at r[e * 32]
    __stcg(r + S::E_NE * 32, (double)w.nefc);
    __stcg(r + S::E_DCOST * 32, dcost);
    sfor<0, S::NT>([&](auto tt) { __stcg(r + (S::E_M + IDX(tt)) * 32, w.M[IDX(tt)]); });
    sfor<0, NV>([&](auto ii) { __stcg(r + (S::E_FS + IDX(ii)) * 32, w.fs[IDX(ii)]); __stcg(r + (S::E_AS + IDX(ii)) * 32, w.as[IDX(ii)]); });
    for (int e = 0; e < w.nefc; e++) {
        double* rr = r + (S::E_ROWS + e * S::RS) * 32;
        sfor<0, NV>([&](auto ii) { __stcg(rr + IDX(ii) * 32, w.rows.J(e, IDX(ii))); });
        __stcg(rr + NV * 32, w.rows.D(e));
        __stcg(rr + (NV + 1) * 32, w.rows.aref(e));
    }
}

// ------------------------------------------------------------------ kernel 2: solve, difference, write out
// One CTA of 32 * NW threads per knot; NW * 32 / GW >= G lane groups.
template <class T, int GW>
struct FdSolveShape {
    using S = FdRecord<T>;
    static constexpr int PPW = 32 / GW;                          // perturbed problems per warp
    static constexpr int NW = (S::G + PPW - 1) / PPW;            // warps per knot
    static constexpr int NSL = (S::ME + GW - 1) / GW;            // row slots per lane
    static constexpr int NSC = (S::ME + 31) / 32;                // ... of the centre solve (all 32 lanes)
    static_assert(PPW % 2 == 0, "the +-eps problems of a column must share a warp");
};

template <class T, int GW, int MINB>
__global__ void __launch_bounds__(32 * FdSolveShape<T, GW>::NW, MINB) fd_solve_kernel(const __grid_constant__ DevModel<T> m, int nknots,
                                                                                 const double* __restrict__ rec, const double* __restrict__ warmstart,
                                                                                 int has_cost, double eps, int niter, int nwarmup, const FdDst dst,
                                                                                 double* __restrict__ qacc_center, int* __restrict__ status,
                                                                                 int* __restrict__ diag) {
    using S = FdRecord<T>;
    using P = FdSolveShape<T, GW>;
    constexpr int NV = T::NV, NU = T::NU, NT = S::NT, ND = NV * S::NCOL + S::NCOL, NJAC = NV * S::NCOL;
    __shared__ double sM[P::NW][P::PPW][NT];   // mass matrices of the warp's problems (M of the centre problem: sM[0][0] during phase 0)
    __shared__ double sC[NV];                  // the centre's solution
    __shared__ double stage[ND];
    __shared__ int sflag;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, k = blockIdx.x;
    const long long t0 = clock64();
    if (threadIdx.x == 0) sflag = 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the records of fd_build_kernel
    const double* rk = rec + (size_t)k * S::DOUBLES;
    auto ld = [&](int e, int l) { return __ldcg(rk + (size_t)e * 32 + l); };
    // ---- phase 0: the centre problem (lane S::G of the record), all 32 lanes of warp 0: row r on lane r % 32, slot r / 32
    int it_first = 0, it_all = 0, nact = 0, ne_c = 0;
    long long t1 = 0, t2 = 0;
    if (wib == 0) {
        ne_c = (int)ld(S::E_NE, S::G);
        double fs[NV], as[NV], warm[NV], qa[NV];
        sfor<0, NV>([&](auto ii) {
            fs[IDX(ii)] = ld(S::E_FS + IDX(ii), S::G);
            as[IDX(ii)] = ld(S::E_AS + IDX(ii), S::G);
            warm[IDX(ii)] = warmstart ? warmstart[(size_t)k * NV + IDX(ii)] : 0.0;
        });
        if (lane < NT) sM[0][0][lane] = ld(S::E_M + lane, S::G);
        LaneRows<T, P::NSC> R;
        R.ns = 0;
#pragma unroll
        for (int s = 0; s < P::NSC; s++) {
            const int r = lane + 32 * s;
            const bool has = r < ne_c;
            if (has) R.ns = s + 1;
            const int e = S::E_ROWS + (has ? r : 0) * S::RS;
            sfor<0, NV>([&](auto ii) { const double x = ld(e + IDX(ii), S::G); R.J[s][IDX(ii)] = has ? x : 0.0; });
            const double D = ld(e + NV, S::G), ar = ld(e + NV + 1, S::G);
            R.D[s] = has ? D : 0.0;
            R.jar[s] = has ? -ar : 1.0;
            R.jv[s] = 0;
        }
        __syncwarp();
        t1 = clock64();
        if (ne_c == 0) sfor<0, NV>([&](auto ii) { qa[IDX(ii)] = as[IDX(ii)]; });
        else {
            // the reference repeats the centre solve nwarmup times (mjderivative.cpp:67-68); a solve that left through the exact-optimum
            // test sits on the minimiser, repeating it changes nothing but round-off (see fd_center_kernel)
#pragma unroll 1
            for (int rep = 0; rep < nwarmup; rep++) {
                int it = 0, ex = 0;
                if (rep > 0) {
#pragma unroll
                    for (int s = 0; s < P::NSC; s++) {   // residuals back to -aref: solve_rows adds J warm
                        const int r = lane + 32 * s;
                        R.jar[s] = r < ne_c ? -ld(S::E_ROWS + r * S::RS + NV + 1, S::G) : 1.0;
                    }
                }
                solve_rows<T, 32, P::NSC>(m, true, R, &sM[0][0][0], fs, as, warm, qa, niter, 0.0, it, ex);
                if (rep == 0) it_first = it;
                it_all += it;
                if (ex) break;
            }
#pragma unroll
            for (int s = 0; s < P::NSC; s++) nact += __popc(__ballot_sync(0xffffffffu, R.jar[s] < 0));
        }
        t2 = clock64();
        if (lane == 0) {
            bool ok = true;
            sfor<0, NV>([&](auto ii) { sC[IDX(ii)] = qa[IDX(ii)]; ok = ok && isfinite(qa[IDX(ii)]); });
            if (qacc_center) sfor<0, NV>([&](auto ii) { qacc_center[(size_t)k * NV + IDX(ii)] = qa[IDX(ii)]; });
            if (!ok) sflag = 1;
        }
    }
    __syncthreads();
    // ---- phase 1: the perturbed problems, GW lanes each; problem p = column p / 2 at +eps (p even) or -eps
    const int g = lane / GW, j = lane % GW, p = wib * P::PPW + g;
    const bool live_p = p < S::G;
    const int pp = live_p ? p : 0;
    const int ne = (int)ld(S::E_NE, pp);
    double fs[NV], as[NV], warm[NV], qa[NV];
    sfor<0, NV>([&](auto ii) { fs[IDX(ii)] = ld(S::E_FS + IDX(ii), pp); as[IDX(ii)] = ld(S::E_AS + IDX(ii), pp); warm[IDX(ii)] = sC[IDX(ii)]; });
    for (int t = j; t < NT; t += GW) sM[wib][g][t] = ld(S::E_M + t, pp);
    LaneRows<T, P::NSL> R;
    R.ns = 0;
#pragma unroll 1
    for (int s = 0; s < P::NSL; s++) {
        const int r = j + GW * s;
        if (r < ne) {
            R.ns = s + 1;
            const int e = S::E_ROWS + r * S::RS;
            sfor<0, NV>([&](auto ii) { R.J[s][IDX(ii)] = ld(e + IDX(ii), pp); });
            R.D[s] = ld(e + NV, pp);
            R.jar[s] = -ld(e + NV + 1, pp);
        }
    }
    __syncwarp();
    int it = 0, ex = 0;
    solve_rows<T, GW, P::NSL>(m, live_p && ne > 0, R, &sM[wib][g][0], fs, as, warm, qa, niter, 0.0, it, ex);
    if (ne == 0) sfor<0, NV>([&](auto ii) { qa[IDX(ii)] = as[IDX(ii)]; });
    const long long t3 = clock64();
    // ---- central differences: the '-' problem is the next group
    bool finite = true;
    const double inv2eps = 1.0 / (2 * eps);
    const int col = p >> 1;
    sfor<0, NV>([&](auto jj) {
        constexpr int jq = IDX(jj);
        const double other = __shfl_xor_sync(0xffffffffu, qa[jq], GW);
        const double d = (qa[jq] - other) * inv2eps;
        finite = finite && isfinite(d);
        if (live_p && !(p & 1) && j == 0) {
            int off;   // reference layout: block of kind, element i + j*stride
            if (col < NU) off = 2 * NV * NV + col + jq * NU;
            else if (col < NU + NV) off = NV * NV + (col - NU) + jq * NV;
            else off = (col - NU - NV) + jq * NV;
            stage[off] = d;
        }
    });
    if (live_p && !(p & 1) && j == 0) {
        int off;   // cost gradient entries: dg/dqpos[nv], dg/dqvel[nv], dg/dctrl[nu]
        if (col < NU) off = NJAC + 2 * NV + col;
        else if (col < NU + NV) off = NJAC + NV + (col - NU);
        else off = (col - NU - NV) + NJAC;
        stage[off] = ld(S::E_DCOST, p);
        if (!finite) sflag = 1;
    }
    __syncthreads();
    const int per = has_cost ? ND : NJAC;   // without a device cost the gradient entries stay untouched
    for (int d = 0; d < dst.n; d++) {
        double* out = dst.p[d] + (size_t)k * ND;
        for (int e = threadIdx.x; e < per; e += blockDim.x) out[e] = stage[e];
    }
    if (threadIdx.x == 0) {
        if (status) status[k] = sflag ? ILQG_ERR_NONFINITE : 0;
        if (diag) {   // ILQG_DIAG_* (include/ilqg_b200.h); cycles: loading the centre problem / its solves / until this warp's perturbed solves were done
            int4* dd = reinterpret_cast<int4*>(diag + (size_t)k * ILQG_DIAG_INTS);
            dd[0] = make_int4(ne_c, it_first, it_all, nact);
            dd[1] = make_int4((int)(t1 - t0), (int)(t2 - t1), 0, (int)(t3 - t2));
        }
    }
}

}  // namespace ilqg
