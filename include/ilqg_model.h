/* ilqg_model.h — flat, fixed-capacity "compiled model" tables.
 *
 * This POD plays the role the reference gives to `mjModel*` (the object every hot-path
 * call receives: calcMJDerivatives(mjModel*, ...) at /root/reference/inc/mjderivative.h:7,
 * mj_step(m,d) at /root/reference/inc/ilqr.h:86,128).  MuJoCo's loader (`mj_loadXML`,
 * /root/reference/cmd/basic.cpp:123, /root/reference/tst/test_derivatives.cpp:34) is not
 * available, so `ilqg_compile_mjcf` (ilqg_b200.h) produces this struct from the MJCF subset
 * that /root/reference/res/{inverted_pendulum,hopper,humanoid}.xml use.
 *
 * The struct has no pointers: it can be fwrite()n as a fixture, memcpy'd across a C ABI,
 * passed to ctypes, and uploaded to the GPU verbatim.  Field names follow mjModel's.
 */
#ifndef ILQG_MODEL_H
#define ILQG_MODEL_H

#ifdef __cplusplus
extern "C" {
#endif

#define ILQG_MODEL_MAGIC 0x494c5147 /* "ILQG" */
#define ILQG_MODEL_VERSION 3

#define ILQG_MAXBODY 16
#define ILQG_MAXJNT 24
#define ILQG_MAXQ 32
#define ILQG_MAXV 32
#define ILQG_MAXU 24
#define ILQG_MAXGEOM 24
#define ILQG_MAXPAIR 176

/* joint types: numbering of MuJoCo's mjtJoint (used at /root/reference/src/mjderivative.cpp:152,157) */
enum { ILQG_JNT_FREE = 0, ILQG_JNT_BALL = 1, ILQG_JNT_SLIDE = 2, ILQG_JNT_HINGE = 3 };
/* geom types: numbering of mjtGeom for the primitives the three models use */
enum { ILQG_GEOM_PLANE = 0, ILQG_GEOM_SPHERE = 2, ILQG_GEOM_CAPSULE = 3 };
/* integrators (mjtIntegrator) */
enum { ILQG_INT_EULER = 0, ILQG_INT_RK4 = 1 };
/* stages of mj_forwardSkip (mjtStage; /root/reference/src/mjderivative.cpp:68,124,178) */
enum { ILQG_STAGE_NONE = 0, ILQG_STAGE_POS = 1, ILQG_STAGE_VEL = 2 };

typedef struct ilqg_model {
    int magic, version;
    /* sizes */
    int nq, nv, nu, nbody, njnt, ngeom, npair;
    int pad0;

    /* option block (mjOption) */
    double timestep;
    double gravity[3];
    double tolerance;      /* solver tolerance used by rollouts (FD pins 0) */
    double ls_tolerance;   /* linesearch relative-gradient tolerance */
    double impratio;
    int integrator;        /* ILQG_INT_* */
    int iterations;        /* solver iterations used by rollouts (FD pins 30) */
    int ls_iterations;
    int pad1;
    double meaninertia;    /* stat.meaninertia = trace(M(qpos0))/nv */

    /* bodies (body 0 = world) */
    int body_parentid[ILQG_MAXBODY];
    int body_rootid[ILQG_MAXBODY];
    int body_jntadr[ILQG_MAXBODY];
    int body_jntnum[ILQG_MAXBODY];
    int body_dofadr[ILQG_MAXBODY];
    int body_dofnum[ILQG_MAXBODY];
    double body_pos[ILQG_MAXBODY][3];     /* frame offset in parent frame */
    double body_quat[ILQG_MAXBODY][4];
    double body_mass[ILQG_MAXBODY];
    double body_ipos[ILQG_MAXBODY][3];    /* centre of mass in body frame */
    double body_inertia[ILQG_MAXBODY][6]; /* xx,yy,zz,xy,xz,yz about ipos, body-frame axes */
    double body_invweight0[ILQG_MAXBODY][2];

    /* joints */
    int jnt_type[ILQG_MAXJNT];
    int jnt_qposadr[ILQG_MAXJNT];
    int jnt_dofadr[ILQG_MAXJNT];
    int jnt_bodyid[ILQG_MAXJNT];
    int jnt_limited[ILQG_MAXJNT];
    double jnt_pos[ILQG_MAXJNT][3];       /* anchor in body frame */
    double jnt_axis[ILQG_MAXJNT][3];      /* unit axis in body frame */
    double jnt_range[ILQG_MAXJNT][2];     /* radians for hinges */
    double jnt_stiffness[ILQG_MAXJNT];
    double jnt_margin[ILQG_MAXJNT];
    double jnt_solref[ILQG_MAXJNT][2];
    double jnt_solimp[ILQG_MAXJNT][5];

    double qpos0[ILQG_MAXQ];
    double qpos_spring[ILQG_MAXQ];

    /* dofs */
    int dof_bodyid[ILQG_MAXV];
    int dof_jntid[ILQG_MAXV];
    int dof_parentid[ILQG_MAXV];
    double dof_armature[ILQG_MAXV];
    double dof_damping[ILQG_MAXV];
    double dof_invweight0[ILQG_MAXV];

    /* geoms */
    int geom_type[ILQG_MAXGEOM];
    int geom_bodyid[ILQG_MAXGEOM];
    int geom_contype[ILQG_MAXGEOM];
    int geom_conaffinity[ILQG_MAXGEOM];
    int geom_condim[ILQG_MAXGEOM];
    double geom_size[ILQG_MAXGEOM][3];
    double geom_pos[ILQG_MAXGEOM][3];     /* in body frame */
    double geom_quat[ILQG_MAXGEOM][4];
    double geom_friction[ILQG_MAXGEOM][3];
    double geom_margin[ILQG_MAXGEOM];
    double geom_gap[ILQG_MAXGEOM];
    double geom_solref[ILQG_MAXGEOM][2];
    double geom_solimp[ILQG_MAXGEOM][5];
    double geom_solmix[ILQG_MAXGEOM];

    /* collision candidates after contype/conaffinity, same-body and parent-child filtering;
       geom1 < geom2, plane (if any) is geom1; contact parameters already mixed */
    int pair_geom1[ILQG_MAXPAIR];
    int pair_geom2[ILQG_MAXPAIR];
    int pair_condim[ILQG_MAXPAIR];
    double pair_margin[ILQG_MAXPAIR];     /* includemargin = margin - gap */
    double pair_friction[ILQG_MAXPAIR];   /* sliding friction mu (tangent 1 == tangent 2) */
    double pair_solref[ILQG_MAXPAIR][2];
    double pair_solimp[ILQG_MAXPAIR][5];

    /* actuators: joint motors only (gain 1, no bias, no activation) */
    int act_dofid[ILQG_MAXU];
    int act_ctrllimited[ILQG_MAXU];
    double act_gear[ILQG_MAXU];
    double act_ctrlrange[ILQG_MAXU][2];
} ilqg_model;

#ifdef __cplusplus
}
#endif
#endif
