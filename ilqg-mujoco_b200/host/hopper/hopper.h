// hopper.h — a task class for /root/reference/res/hopper.xml in the shape of the reference's InvertedPendulum
// (/root/reference/inc/inverted_pendulum/inverted_pendulum.h:10-35; MPC cadence /root/reference/src/inverted_pendulum/inverted_pendulum.cpp:6-30).
// The reference has no such class (README.md:31, "Extend to walking robots"); this is SURVEY.md 8(f) row 3.  It drives the batched
// GPU workspace (ilqg_ilqr_*, one instance) with the opt-in extensions, because the reference's own full-step iteration with its
// column-major-view A/B (quirk Q1) diverges on this model by the third iteration: corrected A/B layout, the backtracking
// ladder of row A10 and the mu schedule.
#pragma once
#include "mujoco/mujoco.h"

class Hopper {
public:
    mjModel* m = NULL;
    mjData* d = NULL;
    static inline constexpr int nv = 6;
    static inline constexpr int nu = 3;
    static inline constexpr int N = 20;
    static inline constexpr int maxIterUtilConvergence = 10;
    static inline constexpr int nalpha = 6;          // 1, 1/2, ... 1/32
    ilqg_ilqr ws = NULL;
    ilqg_cost cost;                                   // see hopperCost()
    mjtNum J[maxIterUtilConvergence];                 // trajectory cost after each iteration of the last forward()

    Hopper(mjModel* m, mjData* d);
    ~Hopper();
    void forward();                                   // one MPC step: maxIter iterations from d, apply the first control, mj_step
    static ilqg_cost hopperCost();                    // 5 z^2 - 12.5 z  (height near 1.25 m) + pitch^2 - x_dot + 0.05 |v|^2 + 0.01 |u|^2
};
