// differentiator.h — mirror of /root/reference/inc/differentiator.h:9-95 (same class name, template parameters, public
// members and methods).  updateDerivatives() = one GPU FD call for the bound knot + the reference's A/B assembly, including
// its layout quirk: the row-major deriv blocks are read through column-major views, and A/B model an explicit
// Euler step whatever the XML integrator is (Q2).
#pragma once
#include <new>

#include "mjderivative.h"
#include "smallmat.h"

template <int nv, int nu>
class Differentiator {
public:
    typedef ilqg::Mat<mjtNum, 2 * nv, 2 * nv> A_t;
    typedef ilqg::Mat<mjtNum, 2 * nv, nu> B_t;
    typedef ilqg::Mat<mjtNum, 2 * nv, 1> x_t;
    typedef ilqg::Mat<mjtNum, nu, 1> u_t;
    typedef ilqg::MatMap<mjtNum, nv, nv> dqdq_mt;
    typedef ilqg::MatMap<mjtNum, nv, nu> dqdu_mt;
    typedef ilqg::MatMap<mjtNum, 2 * nv, 1> x_mt;
    typedef ilqg::MatMap<mjtNum, nu, 1> u_mt;
    typedef ilqg::MatMap<mjtNum, 1, 2 * nv> q_mt;
    typedef ilqg::MatMap<mjtNum, 1, nu> r_mt;

    mjModel* m;
    mjData* d;
    mjtNum* deriv;
    dqdq_mt* dqaccdq;
    dqdq_mt* dqaccdqvel;
    dqdu_mt* dqaccdctrl;
    stepCostFn_t& stepCostFn;
    q_mt* dgdx;
    r_mt* dgdu;
    x_mt* x;
    u_mt* u;
    A_t* A;
    B_t* B;

    Differentiator(mjModel* m, mjData* d, stepCostFn_t& stepCostFn) : m(m), d(d), stepCostFn(stepCostFn) {
        deriv = (mjtNum*)mju_malloc((nv * (2 * nv + nu) + 2 * nv + nu) * sizeof(mjtNum));
        dqaccdq = new dqdq_mt(deriv);
        dqaccdqvel = new dqdq_mt(deriv + nv * nv);
        dqaccdctrl = new dqdu_mt(deriv + 2 * nv * nv);
        dgdx = new q_mt(deriv + 2 * nv * nv + nv * nu);
        dgdu = new r_mt(deriv + 2 * nv * nv + nv * nu + 2 * nv);
        x = new x_mt(d->qpos);
        u = new u_mt(d->ctrl);
        A = new A_t;
        B = new B_t;
        for (int i = 0; i < nv; i++) {
            (*A)(i, i) = 1;
            (*A)(i, nv + i) = m->opt.timestep;
        }
    }

    void setMJData(mjData* dStar) {
        d = dStar;
        new (x) x_mt(d->qpos);
        new (u) u_mt(d->ctrl);
    }

    // assemble A, B from a deriv block (the same arithmetic ilqr_backward_kernel does on the device)
    void assemble(const mjtNum* blk) {
        const mjtNum dt = m->opt.timestep;
        for (int r = 0; r < nv; r++)
            for (int c = 0; c < nv; c++) {
                (*A)(nv + r, c) = blk[r + c * nv] * dt;
                (*A)(nv + r, nv + c) = (r == c ? 1.0 : 0.0) + blk[nv * nv + r + c * nv] * dt;
            }
        for (int r = 0; r < nv; r++)
            for (int c = 0; c < nu; c++) (*B)(nv + r, c) = blk[2 * nv * nv + r + c * nv] * dt;
    }

    void updateDerivatives() {  // derivatives are taken at d
        calcMJDerivatives(m, d, deriv, stepCostFn);
        assemble(deriv);
    }
};
