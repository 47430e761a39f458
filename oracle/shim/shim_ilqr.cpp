// shim_ilqr.cpp — CPU ORACLE (test infrastructure): drives the REFERENCE's own ILQR / Differentiator /
// InvertedPendulum classes (/root/reference/inc/ilqr.h, inc/differentiator.h,
// src/inverted_pendulum/inverted_pendulum.cpp — compiled verbatim against the MuJoCo and Eigen shims).
#include <fcntl.h>
#include <sched.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>

#include "mujoco/mujoco.h"
#include "ilqr.h"                                 // reference header
#include "inverted_pendulum/inverted_pendulum.h"  // reference header
#include "inverted_pendulum/cost.h"               // reference header (stepCost)

namespace {
struct Quiet {  // the reference prints two lines per knot from inside the Riccati loop (ilqr.h:146-147)
    int saved;
    Quiet() { fflush(stdout); saved = dup(1); int nul = open("/dev/null", O_WRONLY); dup2(nul, 1); close(nul); }
    ~Quiet() { std::cout.flush(); fflush(stdout); dup2(saved, 1); close(saved); }
};
struct Cpus {   // MAXTHREAD = 16 guard, see shim.cpp
    cpu_set_t saved; bool ok;
    explicit Cpus(int maxcpus) {
        ok = sched_getaffinity(0, sizeof(saved), &saved) == 0;
        if (!ok) return;
        cpu_set_t lim; CPU_ZERO(&lim);
        int n = 0;
        for (int c = 0; c < CPU_SETSIZE && n < maxcpus; c++) if (CPU_ISSET(c, &saved)) { CPU_SET(c, &lim); n++; }
        sched_setaffinity(0, sizeof(lim), &lim);
    }
    ~Cpus() { if (ok) sched_setaffinity(0, sizeof(saved), &saved); }
};
}  // namespace

extern "C" {

// The reference's MPC demo, headless (cmd/basic.cpp:155-164 minus GLFW): InvertedPendulum(m,d) — which does its own 10
// warm-up steps — then nmpc x forward().  Only ONE call per process is meaningful: ILQR::backwardPass binds
// function-local statics to the first instance (ilqr.h:137-140, quirk Q13).
// Outputs: state/ctrl after each forward() [nmpc x 5], and the solver's final nominal, gains and value model.
int ref_pendulum_mpc(const ilqg_model* tab, const double* qpos0, const double* qvel0, int nmpc, double* trace, double* nom_qpos,
                     double* nom_qvel, double* nom_ctrl, double* K, double* k, double* V, double* v) {
    static bool used = false;
    if (used) return -1;
    used = true;
    Quiet q;
    Cpus c(16);
    mjModel* m = shim_model_from_tables(tab);
    mjData* d = mj_makeData(m);
    if (qpos0) mju_copy(d->qpos, qpos0, m->nq);
    if (qvel0) mju_copy(d->qvel, qvel0, m->nv);
    InvertedPendulum ip(m, d);
    constexpr int N = InvertedPendulum::N, nv = InvertedPendulum::nv, nu = InvertedPendulum::nu;
    for (int s = 0; s < nmpc; s++) {
        ip.forward();
        if (trace) {
            trace[5 * s + 0] = d->qpos[0]; trace[5 * s + 1] = d->qpos[1];
            trace[5 * s + 2] = d->qvel[0]; trace[5 * s + 3] = d->qvel[1]; trace[5 * s + 4] = d->ctrl[0];
        }
    }
    auto* il = ip.iLQR;
    for (int n = 0; n <= N; n++) {
        if (nom_qpos) mju_copy(nom_qpos + n * nv, il->dArray[n]->qpos, nv);
        if (nom_qvel) mju_copy(nom_qvel + n * nv, il->dArray[n]->qvel, nv);
        if (nom_ctrl) mju_copy(nom_ctrl + n * nu, il->dArray[n]->ctrl, nu);
        if (K && n >= 1) mju_copy(K + n * nu * 2 * nv, il->K[n].data(), nu * 2 * nv);
        if (k && n >= 1) mju_copy(k + n * nu, il->k[n].data(), nu);
    }
    if (V) mju_copy(V, il->V->data(), 4 * nv * nv);
    if (v) mju_copy(v, il->v->data(), 2 * nv);
    return 0;
}

}  // extern "C"
