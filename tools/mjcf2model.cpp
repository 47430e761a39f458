// mjcf2model — compile an MJCF file into the flat ilqg_model table file used by tests and bench.
// usage: mjcf2model in.xml out.ilqgm [--print]
#include <cstdio>
#include <cstring>
#include "ilqg_b200.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.xml out.ilqgm [--print]\n", argv[0]); return 2; }
    static ilqg_model m;
    char err[512];
    int rc = ilqg_compile_mjcf(argv[1], &m, err, sizeof err);
    if (rc) { fprintf(stderr, "error %d: %s\n", rc, err); return 1; }
    if (ilqg_model_save(argv[2], &m)) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    if (argc > 3 && !strcmp(argv[3], "--print")) {
        printf("nq %d nv %d nu %d nbody %d njnt %d ngeom %d npair %d dt %g int %d meaninertia %.10g\n", m.nq, m.nv, m.nu, m.nbody,
               m.njnt, m.ngeom, m.npair, m.timestep, m.integrator, m.meaninertia);
        for (int b = 1; b < m.nbody; b++)
            printf("body %d parent %d mass %.10g ipos %.6g %.6g %.6g I %.6g %.6g %.6g %.3g %.3g %.3g invw %.6g %.6g pos %.4g %.4g %.4g\n", b,
                   m.body_parentid[b], m.body_mass[b], m.body_ipos[b][0], m.body_ipos[b][1], m.body_ipos[b][2], m.body_inertia[b][0],
                   m.body_inertia[b][1], m.body_inertia[b][2], m.body_inertia[b][3], m.body_inertia[b][4], m.body_inertia[b][5],
                   m.body_invweight0[b][0], m.body_invweight0[b][1], m.body_pos[b][0], m.body_pos[b][1], m.body_pos[b][2]);
        for (int j = 0; j < m.njnt; j++)
            printf("jnt %d type %d body %d qadr %d dadr %d pos %.4g %.4g %.4g axis %.4g %.4g %.4g range %.4g %.4g lim %d stiff %g\n", j,
                   m.jnt_type[j], m.jnt_bodyid[j], m.jnt_qposadr[j], m.jnt_dofadr[j], m.jnt_pos[j][0], m.jnt_pos[j][1], m.jnt_pos[j][2],
                   m.jnt_axis[j][0], m.jnt_axis[j][1], m.jnt_axis[j][2], m.jnt_range[j][0], m.jnt_range[j][1], m.jnt_limited[j],
                   m.jnt_stiffness[j]);
        for (int d = 0; d < m.nv; d++)
            printf("dof %d body %d jnt %d parent %d arm %g damp %g invw %.6g\n", d, m.dof_bodyid[d], m.dof_jntid[d], m.dof_parentid[d],
                   m.dof_armature[d], m.dof_damping[d], m.dof_invweight0[d]);
        for (int p = 0; p < m.npair && p < 12; p++)
            printf("pair %d g %d %d condim %d margin %g mu %g solref %g %g solimp %g %g %g\n", p, m.pair_geom1[p], m.pair_geom2[p],
                   m.pair_condim[p], m.pair_margin[p], m.pair_friction[p], m.pair_solref[p][0], m.pair_solref[p][1], m.pair_solimp[p][0],
                   m.pair_solimp[p][1], m.pair_solimp[p][2]);
        for (int u = 0; u < m.nu; u++) printf("act %d dof %d gear %g lim %d [%g %g]\n", u, m.act_dofid[u], m.act_gear[u], m.act_ctrllimited[u], m.act_ctrlrange[u][0], m.act_ctrlrange[u][1]);
        printf("qpos0:"); for (int i = 0; i < m.nq; i++) printf(" %g", m.qpos0[i]); printf("\n");
    }
    return 0;
}
