// humanoid.h — a task class for /root/reference/res/humanoid.xml in the shape of the reference's InvertedPendulum
// (/root/reference/inc/inverted_pendulum/inverted_pendulum.h:10-35; MPC cadence /root/reference/src/inverted_pendulum/inverted_pendulum.cpp:6-30).
// The reference has none — its README lists "Extend to walking robots" as a TODO (README.md:31) and its ILQR / Differentiator take
// "2 nv doubles at qpos" as the state, which is undefined with the humanoid's free-joint quaternion (nq = 28, nv = 27; SURVEY quirk
// Q9).  This class drives the batched GPU workspace in the opt-in TANGENT-SPACE mode: state difference through
// mju_subQuat for the root orientation, A/B from the FD blocks (which calcMJDerivatives already takes in tangent coordinates,
// /root/reference/src/mjderivative.cpp:152-169) in the corrected layout, the backtracking ladder of row A10 and the mu schedule.
#pragma once
#include "mujoco/mujoco.h"

class Humanoid {
public:
    mjModel* m = NULL;
    mjData* d = NULL;
    static inline constexpr int nq = 28;
    static inline constexpr int nv = 27;
    static inline constexpr int nu = 21;
    static inline constexpr int N = 10;
    static inline constexpr int maxIterUtilConvergence = 4;
    static inline constexpr int nalpha = 4;          // 1, 1/2, 1/4, 1/8
    ilqg_ilqr ws = NULL;
    ilqg_cost cost;                                   // see humanoidCost()
    mjtNum J[maxIterUtilConvergence];                 // trajectory cost after each iteration of the last forward()

    Humanoid(mjModel* m, mjData* d);
    ~Humanoid();
    void forward();                                   // one MPC step: maxIter iterations from d, apply the first control, mj_step
    static ilqg_cost humanoidCost();                  // 2 z^2 - 5.2 z (torso height near 1.3 m) + qx^2 + qy^2 (upright) + 0.05 |v|^2 + 0.02 |u|^2
};
