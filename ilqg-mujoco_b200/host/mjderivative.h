// mjderivative.h — same declarations as /root/reference/inc/mjderivative.h:5-7; the body (host_api.cpp) sends the knot
// to the GPU instead of forking OpenMP workers.
#pragma once
#include "mujoco/mujoco.h"

typedef mjtNum (*stepCostFn_t)(const mjData*);

// one knot (drop-in for the reference call)
void calcMJDerivatives(mjModel* m, mjData* dmain, mjtNum* deriv, stepCostFn_t stepCostFn);
// all knots of a trajectory in one GPU call: deriv[n] receives knot dknots[n]'s block (ND doubles each)
void calcMJDerivativesBatch(mjModel* m, mjData* const* dknots, int nknots, mjtNum* deriv, stepCostFn_t stepCostFn);
// the forward-difference cost rows of one knot, computed on the host with the caller's function exactly as
// /root/reference/src/mjderivative.cpp:72,88,120,174 does; rows = dg/dqpos[nv], dg/dqvel[nv], dg/dctrl[nu]
void calcCostGradientRows(const mjModel* m, const mjData* dmain, stepCostFn_t stepCostFn, mjtNum* rows);
