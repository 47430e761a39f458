// mujoco/mujoco.h — the slice of MuJoCo's C API that the reference's hot path touches, re-declared
// on top of the B200 C ABI (include/ilqg_b200.h).  `mjModel` carries the compiled tables and the GPU-resident model;
// `mj_step` / `mj_forward` run on the GPU (there is no CPU implementation behind them).
// Reference call sites: /root/reference/src/mjderivative.cpp, /root/reference/inc/ilqr.h:73-86,128,
// /root/reference/src/inverted_pendulum/inverted_pendulum.cpp:12-13,26-29, /root/reference/cmd/basic.cpp:119-128.
#pragma once
#include <stddef.h>

#include "ilqg_b200.h"

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
    mjOption opt;      // read-only here: the GPU model is compiled from `tab` when the model is loaded
    ilqg_model tab;    // flat tables (ilqg_compile_mjcf)
    ilqg_handle gpu;   // GPU-resident compiled model
} mjModel;

typedef struct mjData_ {
    mjtNum time;
    mjtNum* qpos;            // qpos[nq] is immediately followed by qvel[nv] (the reference's x-vector, differentiator.h:62)
    mjtNum* qvel;
    mjtNum* qacc_warmstart;
    mjtNum* ctrl;
    mjtNum* qfrc_applied;
    mjtNum* xfrc_applied;
    mjtNum* qacc;
    void* buffer;
} mjData;

#define mjMIN(a, b) (((a) < (b)) ? (a) : (b))
#define mjMAX(a, b) (((a) > (b)) ? (a) : (b))

#ifdef __cplusplus
extern "C" {
#endif
// mj_loadXML compiles the MJCF subset and uploads it to GPU `ILQG_DEVICE` (env, default 0); *.ilqgm files load precompiled tables
mjModel* mj_loadXML(const char* filename, const void* vfs, char* error, int error_sz);
void mj_deleteModel(mjModel* m);
mjData* mj_makeData(const mjModel* m);
void mj_deleteData(mjData* d);
void mj_resetData(const mjModel* m, mjData* d);
void mj_step(const mjModel* m, mjData* d);     // one step on the GPU (ilqg_step_batch_host, n = 1)
void mj_forward(const mjModel* m, mjData* d);  // qacc on the GPU (ilqg_forward_batch_host, n = 1)
int mj_activate(const char* filename);         // licence call of MuJoCo 2.0 (cmd/basic.cpp:119): no-op
void mju_copy(mjtNum* res, const mjtNum* data, int n);
void* mju_malloc(size_t size);
void mju_free(void* p);
void mju_error(const char* msg);               // prints and aborts, as MuJoCo's default handler does — unless mju_user_error is set
void mju_warning(const char* msg);
extern void (*mju_user_error)(const char*);    // MuJoCo's handler hooks
extern void (*mju_user_warning)(const char*);
void mju_quatIntegrate(mjtNum* quat, const mjtNum* vel, mjtNum scale);   // /root/reference/src/mjderivative.cpp:168,191
// ILQG_OK when the state carries no applied forces, else ILQG_ERR_UNSUPPORTED (qfrc_applied / xfrc_applied are part of the
// knot, /root/reference/src/util.cpp:10-11, but not of the GPU pipeline: mj_step / mj_forward / calcMJDerivatives refuse them)
int ilqg_host_check_applied(const struct mjModel_* m, const struct mjData_* d);
#ifdef __cplusplus
}
#endif
