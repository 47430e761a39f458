// host_api.cpp — host-language side of the drop-in: the MuJoCo-shaped entry points and the reference's free functions
// (calcMJDerivatives, cpMjData, forwardStep/forwardFrame, InvertedPendulum) implemented on the B200 C ABI.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "inverted_pendulum/cost.h"
#include "hopper/hopper.h"
#include "humanoid/humanoid.h"
#include "inverted_pendulum/inverted_pendulum.h"
#include "mjderivative.h"
#include "update.h"
#include "util.h"

static const double kEps = 1e-6;  // /root/reference/src/mjderivative.cpp:39

extern "C" {

// MuJoCo's error / warning hooks: a user handler (mju_user_error / mju_user_warning) replaces the default print-and-exit /
// print.  A handler that returns hands control back to the wrapper, which leaves without computing.
void (*mju_user_error)(const char*) = NULL;
void (*mju_user_warning)(const char*) = NULL;
void mju_error(const char* msg) {
    if (mju_user_error) { mju_user_error(msg); return; }
    fprintf(stderr, "ilqg-b200 ERROR: %s\n", msg);
    exit(1);
}
void mju_warning(const char* msg) {
    if (mju_user_warning) { mju_user_warning(msg); return; }
    fprintf(stderr, "ilqg-b200 WARNING: %s\n", msg);
}
// q <- normalize(q) * quat(axis = vel / |vel|, angle = scale |vel|): MuJoCo's mju_quatIntegrate, the tangent-space perturbation
// of /root/reference/src/mjderivative.cpp:168,191 (same operation order as the kernels' quat_integrate, csrc/dyn.cuh)
void mju_quatIntegrate(mjtNum* quat, const mjtNum* vel, mjtNum scale) {
    mjtNum n = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
    mjtNum ax[3] = {1, 0, 0};
    if (n < 1e-15) n = 0;
    else { ax[0] = vel[0] / n; ax[1] = vel[1] / n; ax[2] = vel[2] / n; }
    const mjtNum s = sin(0.5 * scale * n), c = cos(0.5 * scale * n);
    mjtNum qn = sqrt(quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3]);
    mjtNum q[4] = {1, 0, 0, 0};
    if (qn >= 1e-15) for (int i = 0; i < 4; i++) q[i] = quat[i] / qn;
    const mjtNum r[4] = {c, ax[0] * s, ax[1] * s, ax[2] * s};
    quat[0] = q[0] * r[0] - q[1] * r[1] - q[2] * r[2] - q[3] * r[3];
    quat[1] = q[0] * r[1] + q[1] * r[0] + q[2] * r[3] - q[3] * r[2];
    quat[2] = q[0] * r[2] - q[1] * r[3] + q[2] * r[0] + q[3] * r[1];
    quat[3] = q[0] * r[3] + q[1] * r[2] - q[2] * r[1] + q[3] * r[0];
}
void mju_copy(mjtNum* res, const mjtNum* data, int n) { memcpy(res, data, sizeof(mjtNum) * n); }
void* mju_malloc(size_t size) { return malloc(size); }
void mju_free(void* p) { free(p); }
int mj_activate(const char*) { return 1; }

mjModel* mj_loadXML(const char* filename, const void*, char* error, int error_sz) {
    mjModel* m = (mjModel*)calloc(1, sizeof(mjModel));
    int rc;
    size_t len = filename ? strlen(filename) : 0;
    if (len > 6 && !strcmp(filename + len - 6, ".ilqgm")) {
        rc = ilqg_model_load(filename, &m->tab);
        if (rc && error) snprintf(error, error_sz, "cannot load compiled model '%s' (%d)", filename, rc);
    } else
        rc = ilqg_compile_mjcf(filename, &m->tab, error, error_sz);
    if (rc) { free(m); return NULL; }
    const char* dev = getenv("ILQG_DEVICE");
    rc = ilqg_create(&m->tab, dev ? atoi(dev) : 0, &m->gpu);
    if (rc) {
        if (error) snprintf(error, error_sz, "GPU model creation failed (%d): %s", rc, ilqg_last_error(NULL));
        free(m);
        return NULL;
    }
    m->nq = m->tab.nq; m->nv = m->tab.nv; m->nu = m->tab.nu; m->nbody = m->tab.nbody;
    m->dof_jntid = m->tab.dof_jntid; m->jnt_type = m->tab.jnt_type;
    m->jnt_qposadr = m->tab.jnt_qposadr; m->jnt_dofadr = m->tab.jnt_dofadr;
    m->opt.timestep = m->tab.timestep; m->opt.tolerance = m->tab.tolerance; m->opt.iterations = m->tab.iterations;
    return m;
}
void mj_deleteModel(mjModel* m) {
    if (!m) return;
    ilqg_destroy(m->gpu);
    free(m);
}
mjData* mj_makeData(const mjModel* m) {
    mjData* d = (mjData*)calloc(1, sizeof(mjData));
    size_t n = (size_t)m->nq + 4 * (size_t)m->nv + m->nu + 6 * (size_t)m->nbody;
    mjtNum* b = (mjtNum*)calloc(n, sizeof(mjtNum));
    d->buffer = b;
    d->qpos = b; b += m->nq;
    d->qvel = b; b += m->nv;
    d->qacc_warmstart = b; b += m->nv;
    d->ctrl = b; b += m->nu;
    d->qfrc_applied = b; b += m->nv;
    d->xfrc_applied = b; b += 6 * m->nbody;
    d->qacc = b;
    mj_resetData(m, d);
    return d;
}
void mj_deleteData(mjData* d) {
    if (!d) return;
    free(d->buffer);
    free(d);
}
void mj_resetData(const mjModel* m, mjData* d) {
    size_t n = (size_t)m->nq + 4 * (size_t)m->nv + m->nu + 6 * (size_t)m->nbody;
    memset(d->buffer, 0, n * sizeof(mjtNum));
    memcpy(d->qpos, m->tab.qpos0, sizeof(mjtNum) * m->nq);
    d->time = 0;
}
static bool gpu_check(const mjModel* m, int rc, const char* what) {
    if (rc) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s failed (%d): %s", what, rc, ilqg_last_error(m->gpu));
        mju_error(buf);
        return false;
    }
    return true;
}
// A knot is what cpMjData copies (/root/reference/src/util.cpp:4-13), and that includes qfrc_applied / xfrc_applied, which
// MuJoCo adds to qfrc_smooth.  The GPU pipeline does not take them (the reference never sets them): a state that carries a
// non-zero applied force is refused here instead of being linearised without it.
int ilqg_host_check_applied(const mjModel* m, const mjData* d) {
    for (int i = 0; i < m->nv; i++) if (d->qfrc_applied[i] != 0) return ILQG_ERR_UNSUPPORTED;
    for (int i = 0; i < 6 * m->nbody; i++) if (d->xfrc_applied[i] != 0) return ILQG_ERR_UNSUPPORTED;
    return ILQG_OK;
}
static bool applied_ok(const mjModel* m, const mjData* d, const char* what) {
    if (ilqg_host_check_applied(m, d) == ILQG_OK) return true;
    char buf[256];
    snprintf(buf, sizeof buf, "%s: non-zero qfrc_applied / xfrc_applied are not supported by the GPU pipeline (%d)", what, ILQG_ERR_UNSUPPORTED);
    mju_error(buf);
    return false;
}
void mj_step(const mjModel* m, mjData* d) {
    if (!applied_ok(m, d, "mj_step")) return;
    if (!gpu_check(m, ilqg_step_batch_host(m->gpu, 1, 1, d->qpos, d->qvel, d->ctrl, d->qacc_warmstart, d->qacc), "mj_step")) return;
    d->time += m->opt.timestep;
}
void mj_forward(const mjModel* m, mjData* d) {
    if (!applied_ok(m, d, "mj_forward")) return;
    gpu_check(m, ilqg_forward_batch_host(m->gpu, 1, d->qpos, d->qvel, d->ctrl, d->qacc_warmstart, d->qacc), "mj_forward");
}

}  // extern "C"

// ---- /root/reference/src/util.cpp:4-13
void cpMjData(const mjModel* m, mjData* d_dest, const mjData* d_src) {
    d_dest->time = d_src->time;
    mju_copy(d_dest->qpos, d_src->qpos, m->nq);
    mju_copy(d_dest->qvel, d_src->qvel, m->nv);
    mju_copy(d_dest->qacc, d_src->qacc, m->nv);
    mju_copy(d_dest->qacc_warmstart, d_src->qacc_warmstart, m->nv);
    mju_copy(d_dest->qfrc_applied, d_src->qfrc_applied, m->nv);
    mju_copy(d_dest->xfrc_applied, d_src->xfrc_applied, 6 * m->nbody);
    mju_copy(d_dest->ctrl, d_src->ctrl, m->nu);
}

// ---- /root/reference/src/update.cpp:8-20
void forwardStep(mjModel* model, mjData* data) { mj_step(model, data); }
void forwardFrame(mjModel* model, mjData* data) {
    const mjtNum simstart = data->time;
    while (data->time - simstart < 1.0 / 60.0) forwardStep(model, data);
}

// ---- cost rows: forward differences with the caller's host function, same expressions and evaluation order as
//      /root/reference/src/mjderivative.cpp:72,88,120,174 (centre from dmain, perturbed from a private copy)
void calcCostGradientRows(const mjModel* m, const mjData* dmain, stepCostFn_t stepCostFn, mjtNum* rows) {
    const int nv = m->nv, nu = m->nu;
    mjData* d = mj_makeData(m);
    cpMjData(m, d, dmain);
    const mjtNum costCenter = stepCostFn(dmain);
    for (int i = 0; i < nv && i < nu; i++) {  // the reference's ctrl loop runs over min(nv, nu) columns (:81-82)
        d->ctrl[i] = dmain->ctrl[i] + kEps;
        rows[2 * nv + i] = (stepCostFn(d) - costCenter) / kEps;
        d->ctrl[i] = dmain->ctrl[i];
    }
    for (int i = 0; i < nv; i++) {
        d->qvel[i] = dmain->qvel[i] + kEps;
        rows[nv + i] = (stepCostFn(d) - costCenter) / kEps;
        d->qvel[i] = dmain->qvel[i];
    }
    for (int i = 0; i < nv; i++) {
        const int jid = m->dof_jntid[i];
        // quaternion address and position of the dof inside it (-1: a scalar joint), /root/reference/src/mjderivative.cpp:150-160
        int quatadr = -1, dofpos = 0;
        if (m->jnt_type[jid] == mjJNT_BALL) { quatadr = m->jnt_qposadr[jid]; dofpos = i - m->jnt_dofadr[jid]; }
        else if (m->jnt_type[jid] == mjJNT_FREE && i >= m->jnt_dofadr[jid] + 3) { quatadr = m->jnt_qposadr[jid] + 3; dofpos = i - m->jnt_dofadr[jid] - 3; }
        if (quatadr >= 0) {   // tangent-space perturbation (:163-168)
            mjtNum angvel[3] = {0, 0, 0};
            angvel[dofpos] = kEps;
            mju_quatIntegrate(d->qpos + quatadr, angvel, 1);
        } else
            d->qpos[m->jnt_qposadr[jid] + i - m->jnt_dofadr[jid]] += kEps;
        rows[i] = (stepCostFn(d) - costCenter) / kEps;   // (:174)
        mju_copy(d->qpos, dmain->qpos, m->nq);            // undo (:184)
    }
    mj_deleteData(d);
}

void calcMJDerivativesBatch(mjModel* m, mjData* const* dknots, int nknots, mjtNum* deriv, stepCostFn_t stepCostFn) {
    const int nq = m->nq, nv = m->nv, nu = m->nu, nd = ilqg_deriv_size(&m->tab), nr = 2 * nv + nu;
    std::vector<mjtNum> q((size_t)nknots * nq), v((size_t)nknots * nv), u((size_t)nknots * nu), w((size_t)nknots * nv);
    for (int k = 0; k < nknots; k++) {
        mju_copy(q.data() + (size_t)k * nq, dknots[k]->qpos, nq);
        mju_copy(v.data() + (size_t)k * nv, dknots[k]->qvel, nv);
        mju_copy(u.data() + (size_t)k * nu, dknots[k]->ctrl, nu);
        mju_copy(w.data() + (size_t)k * nv, dknots[k]->qacc_warmstart, nv);
    }
    for (int k = 0; k < nknots; k++)
        if (!applied_ok(m, dknots[k], "calcMJDerivatives")) return;
    std::vector<int> status((size_t)nknots, 0);
    int rc = ilqg_fd_batch_host(m->gpu, nknots, q.data(), v.data(), u.data(), w.data(), NULL, NULL, deriv, NULL, status.data());
    if (rc && rc != ILQG_ERR_NONFINITE && !gpu_check(m, rc, "calcMJDerivatives")) return;
    // per-knot flags: a knot that overflowed the kernels' row / contact capacity has no valid block (an error, as MuJoCo's
    // "nconmax / njmax too small"); a non-finite one is MuJoCo's "Nan, Inf or huge value in QACC" warning
    for (int k = 0; k < nknots; k++) {
        if (status[k] == ILQG_ERR_CAPACITY) {
            char buf[160];
            snprintf(buf, sizeof buf, "calcMJDerivatives: knot %d exceeds the constraint-row / contact capacity of the GPU kernels (%d)", k, status[k]);
            mju_error(buf);
            return;
        }
        if (status[k] == ILQG_ERR_NONFINITE) {
            char buf[160];
            snprintf(buf, sizeof buf, "calcMJDerivatives: Nan, Inf or huge value in the derivatives of knot %d", k);
            mju_warning(buf);
        }
    }
    if (stepCostFn)
        for (int k = 0; k < nknots; k++) calcCostGradientRows(m, dknots[k], stepCostFn, deriv + (size_t)k * nd + (nd - nr));
}

// drop-in for /root/reference/src/mjderivative.cpp:212 (one knot)
void calcMJDerivatives(mjModel* m, mjData* dmain, mjtNum* deriv, stepCostFn_t stepCostFn) {
    mjData* one[1] = {dmain};
    calcMJDerivativesBatch(m, one, 1, deriv, stepCostFn);
}

// ---- /root/reference/src/inverted_pendulum/inverted_pendulum.cpp:6-30
InvertedPendulum::InvertedPendulum(mjModel* m, mjData* d) : m(m), d(d) {
    stepCostFn = stepCost;
    for (auto i = 0; i < 10; i++) mj_step(m, d);
    iLQR = new ILQR<nv, nu, N>(m, d, stepCostFn);
    // the class's own step cost (cost.h) is a quadratic form: the ten iterations of forward() run as one device call
    // (ILQR::iterate(int)); ILQG_MIRROR_HOST_COST=1 keeps the reference's cadence, one iterate() with the host function at a time
    if (!getenv("ILQG_MIRROR_HOST_COST")) {
        ilqg_cost c;
        memset(&c, 0, sizeof c);
        c.q2[0] = 1.0; c.q2[1] = 10.0; c.v2[0] = 1.0; c.v2[1] = 10.0; c.u2[0] = 1.0;   // cost.h:7-17
        iLQR->setDeviceCost(c);
    }
}
void InvertedPendulum::forward() {
    iLQR->setDInit(d);
    if (iLQR->hasDeviceCost() && stepCostFn == stepCost) iLQR->iterate(maxIterUtilConvergence);   // (a caller that swapped stepCostFn gets its function)
    else for (int i = 0; i < maxIterUtilConvergence; i++) iLQR->iterate();
    mju_copy(d->ctrl, iLQR->dArray[N]->ctrl, nu);  // get first u
    mj_step(m, d);                                 // proceed simulation
}

// ---- hopper task (no reference equivalent: SURVEY 8f row 3)
ilqg_cost Hopper::hopperCost() {
    ilqg_cost c;
    memset(&c, 0, sizeof c);
    c.q2[1] = 5.0; c.q1[1] = -12.5;      // 5 (z - 1.25)^2 up to a constant
    c.q2[2] = 1.0;                       // torso pitch
    c.v1[0] = -1.0;                      // forward progress
    for (int i = 0; i < nv; i++) c.v2[i] = 0.05;
    for (int i = 0; i < nu; i++) c.u2[i] = 0.01;
    return c;
}
static void hopperCheck(mjModel* m, int rc, const char* what) {
    if (rc) {
        char buf[512];
        snprintf(buf, sizeof buf, "Hopper: %s failed (%d): %s", what, rc, ilqg_last_error(m->gpu));
        mju_error(buf);
    }
}
Hopper::Hopper(mjModel* m, mjData* d) : m(m), d(d) {
    cost = hopperCost();
    for (auto i = 0; i < 10; i++) mj_step(m, d);
    double alphas[nalpha];
    for (int a = 0; a < nalpha; a++) alphas[a] = 1.0 / (1 << a);
    hopperCheck(m, ilqg_ilqr_create(m->gpu, 1, N, nalpha, alphas, &ws), "ilqg_ilqr_create");
    hopperCheck(m, ilqg_ilqr_set_cost(ws, &cost), "ilqg_ilqr_set_cost");
    hopperCheck(m, ilqg_ilqr_set_layout(ws, 1), "ilqg_ilqr_set_layout");
    hopperCheck(m, ilqg_ilqr_set_mu_schedule(ws, 2.0, 1.0, 1e8), "ilqg_ilqr_set_mu_schedule");
    hopperCheck(m, ilqg_ilqr_init_host(ws, d->qpos, d->qvel, d->ctrl, d->qacc_warmstart), "ilqg_ilqr_init_host");
    for (int i = 0; i < maxIterUtilConvergence; i++) J[i] = 0;
}
Hopper::~Hopper() { ilqg_ilqr_destroy(ws); }
void Hopper::forward() {
    hopperCheck(m, ilqg_ilqr_set_state_host(ws, d->qpos, d->qvel, d->qacc_warmstart), "ilqg_ilqr_set_state_host");
    hopperCheck(m, ilqg_ilqr_iterate(ws, maxIterUtilConvergence, 0, NULL), "ilqg_ilqr_iterate");
    mjtNum u[(N + 1) * nu];
    const int done = ilqg_ilqr_iterations_done(ws);
    std::vector<mjtNum> Jt(done < 256 ? done : 256);
    hopperCheck(m, ilqg_ilqr_get_host(ws, NULL, NULL, u, NULL, NULL, NULL, NULL, Jt.data(), NULL), "ilqg_ilqr_get_host");
    for (int i = 0; i < maxIterUtilConvergence; i++) J[i] = Jt[Jt.size() - maxIterUtilConvergence + i];
    mju_copy(d->ctrl, u + N * nu, nu);   // get first u (knot N is the initial one, ilqr.h:52)
    mj_step(m, d);                       // proceed simulation
}

// ---- humanoid task (no reference equivalent: SURVEY 8f row 3; tangent-space iLQR beyond quirk Q9)
ilqg_cost Humanoid::humanoidCost() {
    ilqg_cost c;
    memset(&c, 0, sizeof c);
    c.q2[2] = 2.0; c.q1[2] = -5.2;       // 2 (z - 1.3)^2 up to a constant
    c.q2[4] = 1.0; c.q2[5] = 1.0;        // root quaternion x, y: upright torso
    for (int i = 0; i < nv; i++) c.v2[i] = 0.05;
    for (int i = 0; i < nu; i++) c.u2[i] = 0.02;
    return c;
}
static void humanoidCheck(mjModel* m, int rc, const char* what) {
    if (rc) {
        char buf[512];
        snprintf(buf, sizeof buf, "Humanoid: %s failed (%d): %s", what, rc, ilqg_last_error(m->gpu));
        mju_error(buf);
    }
}
Humanoid::Humanoid(mjModel* m, mjData* d) : m(m), d(d) {
    cost = humanoidCost();
    for (auto i = 0; i < 10; i++) mj_step(m, d);
    double alphas[nalpha];
    for (int a = 0; a < nalpha; a++) alphas[a] = 1.0 / (1 << a);
    humanoidCheck(m, ilqg_ilqr_create(m->gpu, 1, N, nalpha, alphas, &ws), "ilqg_ilqr_create");
    humanoidCheck(m, ilqg_ilqr_set_cost(ws, &cost), "ilqg_ilqr_set_cost");
    humanoidCheck(m, ilqg_ilqr_set_layout(ws, 1), "ilqg_ilqr_set_layout");
    humanoidCheck(m, ilqg_ilqr_set_mu_schedule(ws, 2.0, 1.0, 1e8), "ilqg_ilqr_set_mu_schedule");
    humanoidCheck(m, ilqg_ilqr_init_host(ws, d->qpos, d->qvel, d->ctrl, d->qacc_warmstart), "ilqg_ilqr_init_host");
    for (int i = 0; i < maxIterUtilConvergence; i++) J[i] = 0;
}
Humanoid::~Humanoid() { ilqg_ilqr_destroy(ws); }
void Humanoid::forward() {
    humanoidCheck(m, ilqg_ilqr_set_state_host(ws, d->qpos, d->qvel, d->qacc_warmstart), "ilqg_ilqr_set_state_host");
    humanoidCheck(m, ilqg_ilqr_iterate(ws, maxIterUtilConvergence, 0, NULL), "ilqg_ilqr_iterate");
    mjtNum u0[nu];
    const int done = ilqg_ilqr_iterations_done(ws);
    std::vector<mjtNum> Jt(done < 256 ? done : 256);
    humanoidCheck(m, ilqg_ilqr_get_first_control_host(ws, u0, Jt.data()), "ilqg_ilqr_get_first_control_host");
    for (int i = 0; i < maxIterUtilConvergence; i++) J[i] = Jt[Jt.size() - maxIterUtilConvergence + i];
    mju_copy(d->ctrl, u0, nu);           // get first u (knot N is the initial one, ilqr.h:52)
    mj_step(m, d);                       // proceed simulation
}

// ---- C entry points for tests (ctypes): the headless MPC demos and a Differentiator<6,3> probe
extern "C" int ilqg_host_humanoid_mpc(const char* model_path, const double* qpos0, const double* qvel0, int nmpc, double* trace /* [nmpc][76] */,
                                      double* Jtrace /* [nmpc][4] */) {
    char err[512] = "";
    mjModel* m = mj_loadXML(model_path, NULL, err, sizeof err);
    if (!m) { fprintf(stderr, "%s\n", err); return 1; }
    mjData* d = mj_makeData(m);
    if (qpos0) mju_copy(d->qpos, qpos0, m->nq);
    if (qvel0) mju_copy(d->qvel, qvel0, m->nv);
    {
        Humanoid hm(m, d);
        for (int s = 0; s < nmpc; s++) {
            hm.forward();
            if (trace) { mju_copy(trace + 76 * s, d->qpos, 28); mju_copy(trace + 76 * s + 28, d->qvel, 27); mju_copy(trace + 76 * s + 55, d->ctrl, 21); }
            if (Jtrace) mju_copy(Jtrace + Humanoid::maxIterUtilConvergence * s, hm.J, Humanoid::maxIterUtilConvergence);
        }
    }
    mj_deleteData(d);
    mj_deleteModel(m);
    return 0;
}

extern "C" int ilqg_host_hopper_mpc(const char* model_path, const double* qpos0, const double* qvel0, int nmpc, double* trace /* [nmpc][15] */,
                                    double* Jtrace /* [nmpc][10] */) {
    char err[512] = "";
    mjModel* m = mj_loadXML(model_path, NULL, err, sizeof err);
    if (!m) { fprintf(stderr, "%s\n", err); return 1; }
    mjData* d = mj_makeData(m);
    if (qpos0) mju_copy(d->qpos, qpos0, m->nq);
    if (qvel0) mju_copy(d->qvel, qvel0, m->nv);
    {
        Hopper hp(m, d);
        for (int s = 0; s < nmpc; s++) {
            hp.forward();
            if (trace) { mju_copy(trace + 15 * s, d->qpos, 6); mju_copy(trace + 15 * s + 6, d->qvel, 6); mju_copy(trace + 15 * s + 12, d->ctrl, 3); }
            if (Jtrace) mju_copy(Jtrace + Hopper::maxIterUtilConvergence * s, hp.J, Hopper::maxIterUtilConvergence);
        }
    }
    mj_deleteData(d);
    mj_deleteModel(m);
    return 0;
}

extern "C" int ilqg_host_pendulum_mpc(const char* model_path, const double* qpos0, const double* qvel0, int nmpc, double* trace,
                                      double* nom_qpos, double* nom_qvel, double* nom_ctrl, double* K, double* k, double* V, double* v) {
    char err[512] = "";
    mjModel* m = mj_loadXML(model_path, NULL, err, sizeof err);
    if (!m) { fprintf(stderr, "%s\n", err); return 1; }
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
    for (int n = 0; n <= N; n++) {
        if (nom_qpos) mju_copy(nom_qpos + n * nv, ip.iLQR->dArray[n]->qpos, nv);
        if (nom_qvel) mju_copy(nom_qvel + n * nv, ip.iLQR->dArray[n]->qvel, nv);
        if (nom_ctrl) mju_copy(nom_ctrl + n * nu, ip.iLQR->dArray[n]->ctrl, nu);
        if (K) mju_copy(K + n * nu * 2 * nv, ip.iLQR->K[n].data(), nu * 2 * nv);
        if (k) mju_copy(k + n * nu, ip.iLQR->k[n].data(), nu);
    }
    if (V) mju_copy(V, ip.iLQR->V->data(), 4 * nv * nv);
    if (v) mju_copy(v, ip.iLQR->v->data(), 2 * nv);
    delete ip.iLQR;
    mj_deleteData(d);
    mj_deleteModel(m);
    return 0;
}

static mjtNum hopperTestCost(const mjData* d) { return d->qpos[0]; }  // /root/reference/tst/test_derivatives.cpp:16-20

// the scenario of /root/reference/tst/test_derivatives.cpp:34-93 with Differentiator<6,3>: returns A (12x12), B (12x3), deriv and
// the one-step prediction residual the reference test prints
extern "C" int ilqg_host_hopper_differentiator(const char* model_path, int nsteps, double* A, double* B, double* deriv, double* residual) {
    char err[512] = "";
    mjModel* m = mj_loadXML(model_path, NULL, err, sizeof err);
    if (!m) { fprintf(stderr, "%s\n", err); return 1; }
    mjData* dStar = mj_makeData(m);
    for (int i = 0; i < nsteps; i++) mj_step(m, dStar);
    for (int i = 0; i < m->nu; i++) dStar->ctrl[i] -= 0.1;
    stepCostFn_t fn = hopperTestCost;
    Differentiator<6, 3> diff(m, dStar, fn);
    diff.updateDerivatives();
    if (A) mju_copy(A, diff.A->data(), 144);
    if (B) mju_copy(B, diff.B->data(), 36);
    if (deriv) mju_copy(deriv, diff.deriv, 105);
    if (residual) {
        mjData* d = mj_makeData(m);
        cpMjData(m, d, dStar);
        double xs[12], us[3];
        mju_copy(xs, dStar->qpos, 12);
        mju_copy(us, dStar->ctrl, 3);
        mj_step(m, dStar);
        for (int i = 0; i < 12; i++) d->qpos[i] += 1e-6;  // qpos and qvel are contiguous
        for (int i = 0; i < 3; i++) d->ctrl[i] += 1e-6;
        double x[12], u[3];
        mju_copy(x, d->qpos, 12);
        mju_copy(u, d->ctrl, 3);
        mj_step(m, d);
        for (int r = 0; r < 12; r++) {
            double p = dStar->qpos[r];
            for (int c = 0; c < 12; c++) p += (*diff.A)(r, c) * (x[c] - xs[c]);
            for (int c = 0; c < 3; c++) p += (*diff.B)(r, c) * (u[c] - us[c]);
            residual[r] = p - d->qpos[r];
        }
        mj_deleteData(d);
    }
    mj_deleteData(dStar);
    mj_deleteModel(m);
    return 0;
}

// ---- probes for the parity holes of the drop-in (tests/test_host_dropin_gpu.py)
// a host cost that depends on the root orientation (an uprightness term: 1 - z-axis of the torso dotted with the world z-axis)
static mjtNum uprightCost(const mjData* d) {
    const mjtNum* q = d->qpos + 3;
    return (1.0 - (1.0 - 2.0 * (q[1] * q[1] + q[2] * q[2]))) + 0.5 * d->qpos[2] * d->qpos[2] + 0.1 * d->qvel[4] + 0.01 * d->ctrl[2] * d->ctrl[2];
}
// calcMJDerivatives on a free-joint model with that cost: deriv out (ND doubles)
extern "C" int ilqg_host_freejoint_cost_rows(const char* model_path, const double* qpos, const double* qvel, const double* ctrl, double* deriv) {
    char err[512] = "";
    mjModel* m = mj_loadXML(model_path, NULL, err, sizeof err);
    if (!m) { fprintf(stderr, "%s\n", err); return 1; }
    mjData* d = mj_makeData(m);
    mju_copy(d->qpos, qpos, m->nq);
    mju_copy(d->qvel, qvel, m->nv);
    mju_copy(d->ctrl, ctrl, m->nu);
    calcMJDerivatives(m, d, deriv, uprightCost);
    mj_deleteData(d);
    mj_deleteModel(m);
    return 0;
}

static int g_probe_errors = 0;
static char g_probe_msg[256];
static void probeHandler(const char* msg) { g_probe_errors++; snprintf(g_probe_msg, sizeof g_probe_msg, "%s", msg); }
// a state with applied forces must be refused by every wrapper (and left untouched); returns the number of refusals, -1 on a
// setup failure; msg receives the last error text
extern "C" int ilqg_host_applied_force_probe(const char* model_path, int use_xfrc, char* msg, int msglen) {
    char err[512] = "";
    mjModel* m = mj_loadXML(model_path, NULL, err, sizeof err);
    if (!m) { fprintf(stderr, "%s\n", err); return -1; }
    mjData* d = mj_makeData(m);
    if (ilqg_host_check_applied(m, d) != ILQG_OK) return -1;
    if (use_xfrc) d->xfrc_applied[6 * (m->nbody - 1) + 2] = 3.0; else d->qfrc_applied[m->nv - 1] = -0.5;
    if (ilqg_host_check_applied(m, d) != ILQG_ERR_UNSUPPORTED) return -1;
    std::vector<mjtNum> deriv(ilqg_deriv_size(&m->tab), 123.0), q0(d->qpos, d->qpos + m->nq);
    g_probe_errors = 0;
    void (*old)(const char*) = mju_user_error;
    mju_user_error = probeHandler;
    calcMJDerivatives(m, d, deriv.data(), NULL);
    mj_step(m, d);
    mj_forward(m, d);
    mju_user_error = old;
    int n = g_probe_errors;
    for (size_t i = 0; i < deriv.size(); i++) if (deriv[i] != 123.0) n = -2;   // nothing may have been computed
    for (int i = 0; i < m->nq; i++) if (d->qpos[i] != q0[i]) n = -3;
    if (d->time != 0) n = -4;
    if (msg && msglen > 0) snprintf(msg, msglen, "%s", g_probe_msg);
    mj_deleteData(d);
    mj_deleteModel(m);
    return n;
}
