// inverted_pendulum.h — mirror of /root/reference/inc/inverted_pendulum/inverted_pendulum.h:10-35
#pragma once
#include "ilqr.h"
#include "mjderivative.h"

class InvertedPendulum {
public:
    mjModel* m = NULL;
    mjData* d = NULL;
    static inline constexpr int nv = 2;
    static inline constexpr int nu = 1;
    static inline constexpr int N = 20;
    static inline constexpr int maxIterUtilConvergence = 10;
    ILQR<nv, nu, N>* iLQR;
    stepCostFn_t stepCostFn;

    InvertedPendulum(mjModel* m, mjData* d);
    void forward();
};
