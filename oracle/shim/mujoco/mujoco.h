/* mujoco/mujoco.h SHIM — CPU ORACLE (test infrastructure).
 *
 * Declares exactly the MuJoCo 2.x symbols the reference's FD driver uses, backed
 * by the plain-C oracle in oracle/mjo_engine.c, so that /root/reference/src/mjderivative.cpp,
 * src/util.cpp and src/update.cpp compile VERBATIM from where they lie (oracle/Makefile target
 * `ref` -> oracle/_ref/libref_fd.so).  Nothing here is copied from MuJoCo or the reference.
 */
#ifndef SHIM_MUJOCO_H
#define SHIM_MUJOCO_H

#include <stddef.h>
#include "../../mjo.h"

typedef double mjtNum;

typedef enum { mjJNT_FREE = 0, mjJNT_BALL = 1, mjJNT_SLIDE = 2, mjJNT_HINGE = 3 } mjtJoint;
typedef enum { mjSTAGE_NONE = 0, mjSTAGE_POS = 1, mjSTAGE_VEL = 2, mjSTAGE_ACC = 3 } mjtStage;

typedef struct mjOption_ {
    mjtNum timestep;
    mjtNum tolerance;
    int iterations;
} mjOption;

typedef struct mjModel_ {
    int nq, nv, nu, nbody;
    int* dof_jntid;
    int* jnt_type;
    int* jnt_qposadr;
    int* jnt_dofadr;
    mjOption opt;
    ilqg_model tab; /* the compiled tables the oracle engine reads */
} mjModel;

/* mjData IS the oracle's data block: time, qpos, qvel, ctrl, qacc, qacc_warmstart,
   qfrc_applied, xfrc_applied are its leading members */
typedef struct mjo_data mjData;

#define mjMIN(a, b) (((a) < (b)) ? (a) : (b))
#define mjMAX(a, b) (((a) > (b)) ? (a) : (b))
#define mjMARKSTACK int _mark = d->pstack;
#define mjFREESTACK d->pstack = _mark;

#ifdef __cplusplus
extern "C" {
#endif
mjModel* shim_model_from_tables(const ilqg_model* tab);
void mj_deleteModel(mjModel* m);
mjData* mj_makeData(const mjModel* m);
void mj_deleteData(mjData* d);
mjtNum* mj_stackAlloc(mjData* d, int size);
void mj_forward(const mjModel* m, mjData* d);
void mj_forwardSkip(const mjModel* m, mjData* d, int skipstage, int skipsensor);
void mj_step(const mjModel* m, mjData* d);
void mju_copy(mjtNum* res, const mjtNum* data, int n);
void mju_quatIntegrate(mjtNum* quat, const mjtNum* vel, mjtNum scale);
void* mju_malloc(size_t size);
void mju_free(void* p);
#ifdef __cplusplus
}
#endif
#endif
