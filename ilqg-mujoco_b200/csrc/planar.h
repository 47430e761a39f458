// planar.h — host-side test shared by the topology generator (tools/gen_topology.cpp) and the model binding (dyn.cuh).
#pragma once
#include "../../include/ilqg_b200.h"

namespace ilqg {

// a tree that moves in the xz-plane — slides in the plane, hinges about y, every offset / frame / geom axis / inertia
// product / gravity component that would leave the plane exactly zero (tools/gen_topology.cpp sets T::PLANAR_Y by the same test)
inline bool model_is_planar_y(const ilqg_model& s) {
    if (s.gravity[1] != 0) return false;
    for (int b = 0; b < s.nbody; b++)
        if (s.body_pos[b][1] != 0 || s.body_ipos[b][1] != 0 || s.body_quat[b][1] != 0 || s.body_quat[b][3] != 0 || s.body_inertia[b][3] != 0 ||
            s.body_inertia[b][5] != 0)
            return false;
    for (int j = 0; j < s.njnt; j++) {
        if (s.jnt_pos[j][1] != 0) return false;
        if (s.jnt_type[j] == ILQG_JNT_SLIDE) { if (s.jnt_axis[j][1] != 0) return false; }
        else if (s.jnt_type[j] == ILQG_JNT_HINGE) { if (s.jnt_axis[j][0] != 0 || s.jnt_axis[j][2] != 0) return false; }
        else return false;
    }
    for (int g = 0; g < s.ngeom; g++) {
        const double* q = s.geom_quat[g];
        if (s.geom_pos[g][1] != 0 || 2 * (q[2] * q[3] - q[0] * q[1]) != 0) return false;   // y of the geom's local +z axis
    }
    return true;
}

}  // namespace ilqg
