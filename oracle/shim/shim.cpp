// shim.cpp — CPU ORACLE (test infrastructure): the MuJoCo symbols of shim/mujoco/mujoco.h on top of
// the oracle engine, plus C entry points that drive the REFERENCE's own calcMJDerivatives
// (/root/reference/src/mjderivative.cpp:212, compiled verbatim into this library) for tests and for
// bench.py's `--impl reference` / cpu_baseline legs.
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include "mujoco/mujoco.h"
#include "mjderivative.h"  // the reference's own header (-I/root/reference/inc)
#include "../../include/ilqg_b200.h"

extern "C" {

mjModel* shim_model_from_tables(const ilqg_model* tab) {
    mjModel* m = (mjModel*)calloc(1, sizeof(mjModel));
    m->tab = *tab;
    m->nq = tab->nq; m->nv = tab->nv; m->nu = tab->nu; m->nbody = tab->nbody;
    m->dof_jntid = m->tab.dof_jntid;
    m->jnt_type = m->tab.jnt_type;
    m->jnt_qposadr = m->tab.jnt_qposadr;
    m->jnt_dofadr = m->tab.jnt_dofadr;
    m->opt.timestep = tab->timestep;
    m->opt.tolerance = tab->tolerance;
    m->opt.iterations = tab->iterations;
    return m;
}
void mj_deleteModel(mjModel* m) { free(m); }
mjData* mj_makeData(const mjModel* m) { return mjo_make_data(&m->tab); }
void mj_deleteData(mjData* d) { mjo_delete_data(d); }
mjtNum* mj_stackAlloc(mjData* d, int size) {
    mjtNum* p = d->stack + d->pstack;
    d->pstack += size;
    return p;
}
void mj_forwardSkip(const mjModel* m, mjData* d, int skipstage, int) {
    mjo_forward_skip(&m->tab, d, skipstage, m->opt.iterations, m->opt.tolerance);
}
void mj_forward(const mjModel* m, mjData* d) { mj_forwardSkip(m, d, mjSTAGE_NONE, 0); }
void mj_step(const mjModel* m, mjData* d) {
    ilqg_model t = m->tab;  // honour a mutated opt block
    t.iterations = m->opt.iterations; t.tolerance = m->opt.tolerance; t.timestep = m->opt.timestep;
    mjo_step(&t, d);
}
void mju_copy(mjtNum* res, const mjtNum* data, int n) { memcpy(res, data, sizeof(mjtNum) * n); }
void mju_quatIntegrate(mjtNum* quat, const mjtNum* vel, mjtNum scale) { mjo_quat_integrate(quat, vel, scale); }
void* mju_malloc(size_t size) { return malloc(size); }
void mju_free(void* p) { free(p); }

// ---- driving the reference's calcMJDerivatives --------------------------------------------
static const ilqg_cost* g_cost = nullptr;
static const ilqg_model* g_tab = nullptr;
static mjtNum quadCost(const mjData* d) {
    const ilqg_cost* c = g_cost;
    mjtNum g = 0;
    for (int i = 0; i < g_tab->nq; i++) { g += c->q2[i] * d->qpos[i] * d->qpos[i]; g += c->q1[i] * d->qpos[i]; }
    for (int i = 0; i < g_tab->nv; i++) { g += c->v2[i] * d->qvel[i] * d->qvel[i]; g += c->v1[i] * d->qvel[i]; }
    for (int i = 0; i < g_tab->nu; i++) { g += c->u2[i] * d->ctrl[i] * d->ctrl[i]; g += c->u1[i] * d->ctrl[i]; }
    return g;
}
static mjtNum zeroCost(const mjData*) { return 0; }

// The reference sizes its per-thread array with MAXTHREAD = 16 but launches omp_get_num_procs()
// workers (mjderivative.cpp:32,217,220 — SURVEY quirk Q7).  Restrict the affinity mask to at most
// `maxcpus` (<= 16) CPUs around the call so the unmodified code stays inside its array.
static int restrict_cpus(cpu_set_t* saved, int maxcpus) {
    if (sched_getaffinity(0, sizeof(*saved), saved) != 0) return -1;
    cpu_set_t lim;
    CPU_ZERO(&lim);
    int n = 0;
    for (int c = 0; c < CPU_SETSIZE && n < maxcpus; c++)
        if (CPU_ISSET(c, saved)) { CPU_SET(c, &lim); n++; }
    sched_setaffinity(0, sizeof(lim), &lim);
    return n;
}

// knots serially, as ILQR::backwardPass does (/root/reference/inc/ilqr.h:144-154).
// Returns the number of CPUs (= OpenMP workers) the reference used.
int ref_calc_derivatives_batch(const ilqg_model* tab, int nknots, const double* qpos, const double* qvel, const double* ctrl,
                               const double* warm, const ilqg_cost* cost, double* deriv, int maxcpus) {
    if (maxcpus <= 0 || maxcpus > 16) maxcpus = 16;
    cpu_set_t saved;
    int ncpu = restrict_cpus(&saved, maxcpus);
    mjModel* m = shim_model_from_tables(tab);
    mjData* d = mj_makeData(m);
    g_cost = cost; g_tab = tab;
    int nv = tab->nv, nu = tab->nu, nq = tab->nq, nd = nv * (2 * nv + nu) + 2 * nv + nu;
    for (int k = 0; k < nknots; k++) {
        mju_copy(d->qpos, qpos + (size_t)k * nq, nq);
        mju_copy(d->qvel, qvel + (size_t)k * nv, nv);
        mju_copy(d->ctrl, ctrl + (size_t)k * nu, nu);
        if (warm) mju_copy(d->qacc_warmstart, warm + (size_t)k * nv, nv);
        calcMJDerivatives(m, d, deriv + (size_t)k * nd, cost ? quadCost : zeroCost);
    }
    mj_deleteData(d);
    mj_deleteModel(m);
    if (ncpu > 0) sched_setaffinity(0, sizeof(saved), &saved);
    return ncpu;
}

}  // extern "C"
