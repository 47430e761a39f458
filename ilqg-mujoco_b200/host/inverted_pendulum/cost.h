// cost.h — the step cost of /root/reference/inc/inverted_pendulum/cost.h:7-17: q0^2 + 10 q1^2 + v0^2 + 10 v1^2 + u0^2
#pragma once
#include "mujoco/mujoco.h"

inline mjtNum stepCost(const mjData* d) {
    const mjtNum wq[2] = {1.0, 10.0};
    mjtNum cost = 0;
    for (int i = 0; i < 2; i++) cost += wq[i] * d->qpos[i] * d->qpos[i];
    for (int i = 0; i < 2; i++) cost += wq[i] * d->qvel[i] * d->qvel[i];
    cost += 1.0 * d->ctrl[0] * d->ctrl[0];
    return cost;
}
