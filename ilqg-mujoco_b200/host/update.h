// update.h — /root/reference/inc/update.h:6-8
#pragma once
#include "mujoco/mujoco.h"

void forwardStep(mjModel* model, mjData* data);
void forwardFrame(mjModel* model, mjData* data);
