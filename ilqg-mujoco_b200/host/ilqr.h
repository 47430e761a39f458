// ilqr.h — mirror of /root/reference/inc/ilqr.h:14-188: same class template, public data (dArray, K, k, V, v, mu, d,
// differentiator) and methods (setDInit, forwardPass, initV, backwardPass, iterate).  Every heavy step runs on the GPU
// through a one-instance iLQR workspace (include/ilqg_b200.h):
//   forwardPass  -> ilqg_ilqr_forward   (all N+1 mj_steps in one launch instead of N+1 library calls)
//   backwardPass -> ilqg_ilqr_linearise (all N+1 knots in one FD call instead of one calcMJDerivatives per knot)
//                   + the caller's stepCostFn rows on the host + ilqg_ilqr_backward (Riccati on the device)
// Reference quirks kept: K/k zero before the first backward pass (Q5), mu never removed from V (Q3), nominal overwritten in
// place (Q12).  Not kept: the function-local statics of backwardPass (Q13) — several instances may coexist.
#pragma once
#include <cstdio>
#include <vector>

#include "differentiator.h"
#include "util.h"

template <int nv, int nu, int N>
class ILQR {
public:
    typedef ilqg::Mat<mjtNum, nu, 2 * nv> K_t;
    typedef ilqg::Mat<mjtNum, nu, 1> k_t;
    typedef ilqg::Mat<mjtNum, 2 * nv, 2 * nv> V_t;
    typedef ilqg::Mat<mjtNum, 1, 2 * nv> v_t;

    mjModel* m;
    mjData* d = NULL;
    Differentiator<nv, nu>* differentiator;
    mjData* dArray[N + 1];  // d[N]: init state, d[0]: landing state
    V_t* V;
    v_t* v;
    K_t K[N + 1];
    k_t k[N + 1];
    mjtNum mu = 1000.0;

    ILQR(mjModel* m, mjData* dmain, stepCostFn_t& stepCostFn) : m(m) {
        d = mj_makeData(m);
        setDInit(dmain);
        differentiator = new Differentiator<nv, nu>(m, d, stepCostFn);
        for (int n = 0; n <= N; n++) dArray[n] = mj_makeData(m);
        V = new V_t;
        v = new v_t;
        const double one = 1.0;
        check(ilqg_ilqr_create(m->gpu, 1, N, 1, &one, &ws), "ilqg_ilqr_create");
        check(ilqg_ilqr_set_cost(ws, NULL), "ilqg_ilqr_set_cost");  // host stepCostFn: rows are supplied from here
        // initial trajectory: open-loop rollout under dmain's control (ilqr.h:82-87)
        check(ilqg_ilqr_init_host(ws, d->qpos, d->qvel, d->ctrl, d->qacc_warmstart), "ilqg_ilqr_init_host");
        fetchKnots();
    }
    virtual ~ILQR() { ilqg_ilqr_destroy(ws); }

    virtual void initV() {}  // the terminal condition is part of ilqg_ilqr_backward (ilqr.h:100-107)

    void setDInit(mjData* dInit) { cpMjData(m, d, dInit); }

    void forwardPass() {  // apply control policies in K and k, starting from d
        check(ilqg_ilqr_set_state_host(ws, d->qpos, d->qvel, d->qacc_warmstart), "ilqg_ilqr_set_state_host");
        check(ilqg_ilqr_forward(ws, 1, NULL), "ilqg_ilqr_forward");
        fetchKnots();
    }

    void backwardPass() {  // calculate K and k
        ilqg_ilqr_set_mu(ws, mu);
        check(ilqg_ilqr_linearise(ws, NULL), "ilqg_ilqr_linearise");
        constexpr int NR = 2 * nv + nu;
        std::vector<mjtNum> rows((size_t)(N + 1) * NR);
        for (int n = 0; n <= N; n++) calcCostGradientRows(m, dArray[n], differentiator->stepCostFn, rows.data() + (size_t)n * NR);
        check(ilqg_ilqr_put_cost_rows_host(ws, rows.data()), "ilqg_ilqr_put_cost_rows_host");
        check(ilqg_ilqr_backward(ws, NULL), "ilqg_ilqr_backward");
        fetchGains();
    }

    void iterate() {
        forwardPass();
        setDInit(dArray[N]);
        backwardPass();
    }

    // Not in the reference: `niter` iterations in ONE device call.  iterate() above keeps every public member current after every
    // phase, as the reference's class does — four synchronous host round trips per iteration, because the caller's stepCostFn is a
    // host function.  A caller whose step cost is a quadratic form can hand it over as an ilqg_cost (setDeviceCost): the forward
    // differences of the cost are then taken on the device with the same arithmetic (bit-exact rows, tests/test_fd_gpu.py), the
    // niter x (forwardPass; setDInit(dArray[N]); backwardPass) chain replays as one CUDA graph, and the public members (dArray, d,
    // K, k, V, v) are brought up to date once, at the end — the state iterate() x niter would leave.
    void setDeviceCost(const ilqg_cost& c) {
        check(ilqg_ilqr_set_cost(ws, &c), "ilqg_ilqr_set_cost");
        deviceCost = true;
    }
    bool hasDeviceCost() const { return deviceCost; }
    void iterate(int niter) {
        if (!deviceCost) {
            for (int i = 0; i < niter; i++) iterate();
            return;
        }
        check(ilqg_ilqr_set_state_host(ws, d->qpos, d->qvel, d->qacc_warmstart), "ilqg_ilqr_set_state_host");
        ilqg_ilqr_set_mu(ws, mu);
        check(ilqg_ilqr_iterate(ws, niter, 1, NULL), "ilqg_ilqr_iterate");
        fetchKnots();
        setDInit(dArray[N]);
        fetchGains();
    }

private:
    ilqg_ilqr ws = NULL;
    bool deviceCost = false;
    void fetchGains() {
        std::vector<mjtNum> Kb((size_t)(N + 1) * nu * 2 * nv), kb((size_t)(N + 1) * nu);
        check(ilqg_ilqr_get_host(ws, NULL, NULL, NULL, Kb.data(), kb.data(), V->data(), v->data(), NULL, NULL), "ilqg_ilqr_get_host");
        for (int n = 0; n <= N; n++) {
            for (int e = 0; e < nu * 2 * nv; e++) K[n].data()[e] = Kb[(size_t)n * nu * 2 * nv + e];
            for (int e = 0; e < nu; e++) k[n].data()[e] = kb[(size_t)n * nu + e];
        }
    }
    void check(int rc, const char* what) {
        if (rc) {
            char buf[512];
            snprintf(buf, sizeof buf, "ILQR: %s failed (%d): %s", what, rc, ilqg_last_error(m->gpu));
            mju_error(buf);
        }
    }
    void fetchKnots() {
        std::vector<mjtNum> q((size_t)(N + 1) * nv), qv((size_t)(N + 1) * nv), u((size_t)(N + 1) * nu), w((size_t)(N + 1) * nv);
        check(ilqg_ilqr_get_knots_host(ws, q.data(), qv.data(), u.data(), w.data()), "ilqg_ilqr_get_knots_host");
        for (int n = 0; n <= N; n++) {
            mju_copy(dArray[n]->qpos, q.data() + (size_t)n * nv, nv);
            mju_copy(dArray[n]->qvel, qv.data() + (size_t)n * nv, nv);
            mju_copy(dArray[n]->ctrl, u.data() + (size_t)n * nu, nu);
            mju_copy(dArray[n]->qacc_warmstart, w.data() + (size_t)n * nv, nv);
        }
    }
};
