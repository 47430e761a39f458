// base.cpp — headless equivalent of the reference's ./bin/base (/root/reference/cmd/basic.cpp:109-196 minus GLFW/OpenGL):
// load the model, build InvertedPendulum, run the MPC loop, print one JSON line per step.
// usage: base model.xml|model.ilqgm [nsteps]
#include <cstdio>
#include <cstdlib>

#include "inverted_pendulum/inverted_pendulum.h"

int main(int argc, const char** argv) {
    if (argc < 2) { printf(" USAGE:  base modelfile [nsteps]\n"); return 0; }
    mj_activate("mjkey.txt");
    char error[1000] = "Could not load model";
    mjModel* m = mj_loadXML(argv[1], 0, error, 1000);
    if (!m) mju_error(error);
    mjData* d = mj_makeData(m);
    int nsteps = argc > 2 ? atoi(argv[2]) : 50;
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
