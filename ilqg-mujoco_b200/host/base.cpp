// base.cpp — headless equivalent of the reference's ./bin/base (/root/reference/cmd/basic.cpp:109-196 minus GLFW/OpenGL):
// load the model, build InvertedPendulum, run the MPC loop, print one JSON line per step.
// usage: base model.xml|model.ilqgm [nsteps]
#include <cstdio>
#include <cstdlib>

#include "hopper/hopper.h"
#include "humanoid/humanoid.h"
#include "inverted_pendulum/inverted_pendulum.h"

int main(int argc, const char** argv) {
    if (argc < 2) { printf(" USAGE:  base modelfile [nsteps]\n"); return 0; }
    mj_activate("mjkey.txt");
    char error[1000] = "Could not load model";
    mjModel* m = mj_loadXML(argv[1], 0, error, 1000);
    if (!m) mju_error(error);
    mjData* d = mj_makeData(m);
    int nsteps = argc > 2 ? atoi(argv[2]) : 50;
    if (m->nv == Hopper::nv && m->nu == Hopper::nu) {   // res/hopper.xml: the hopper task (no reference equivalent)
        Hopper hopper(m, d);
        for (int s = 0; s < nsteps; s++) {
            hopper.forward();
            printf("{\"step\": %d, \"x\": %.9g, \"z\": %.9g, \"pitch\": %.9g, \"xdot\": %.9g, \"cost\": %.9g, \"ctrl\": [%.9g, %.9g, %.9g]}\n", s, d->qpos[0],
                   d->qpos[1], d->qpos[2], d->qvel[0], hopper.J[Hopper::maxIterUtilConvergence - 1], d->ctrl[0], d->ctrl[1], d->ctrl[2]);
        }
        mj_deleteData(d);
        mj_deleteModel(m);
        return 0;
    }
    if (m->nv == Humanoid::nv && m->nu == Humanoid::nu) {   // res/humanoid.xml: the humanoid task (no reference equivalent)
        Humanoid humanoid(m, d);
        for (int s = 0; s < nsteps; s++) {
            humanoid.forward();
            printf("{\"step\": %d, \"x\": %.9g, \"z\": %.9g, \"quat\": [%.9g, %.9g, %.9g, %.9g], \"cost\": %.9g}\n", s, d->qpos[0], d->qpos[2], d->qpos[3],
                   d->qpos[4], d->qpos[5], d->qpos[6], humanoid.J[Humanoid::maxIterUtilConvergence - 1]);
        }
        mj_deleteData(d);
        mj_deleteModel(m);
        return 0;
    }
    InvertedPendulum invertedPendulum(m, d);
    for (int s = 0; s < nsteps; s++) {
        invertedPendulum.forward();
        printf("{\"step\": %d, \"time\": %.6f, \"qpos\": [%.12g, %.12g], \"qvel\": [%.12g, %.12g], \"ctrl\": [%.12g]}\n", s, d->time, d->qpos[0],
               d->qpos[1], d->qvel[0], d->qvel[1], d->ctrl[0]);
    }
    mj_deleteData(d);
    mj_deleteModel(m);
    return 0;
}
