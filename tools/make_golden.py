#!/usr/bin/env python
"""Generate tests/golden/*.npz: seeded knot inputs and the CPU oracle's outputs for them.

The reference ships no golden vectors and MuJoCo is unavailable, so these fixtures pin the
ORACLE (regression) and give the GPU tests a device-independent target; the oracle itself is anchored by
tests/test_oracle_anchors.py.  Run from the repo root:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
from conftest import scenario_states  # noqa: E402

o = entry.load_oracle()
pkg = entry.load_package()
out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
COSTS = {"inverted_pendulum": dict(q2=[1, 10], v2=[1, 10], u2=[1]),   # /root/reference/inc/inverted_pendulum/cost.h:7-17
         "hopper": dict(q1=[1]),                                      # /root/reference/tst/test_derivatives.cpp:16-20
         "humanoid": dict(q1=[0, 0, 1])}
for name, n, roll in (("inverted_pendulum", 24, 10), ("hopper", 24, 120), ("humanoid", 6, 40)):
    om = o.Model(os.path.join(pkg.MODELS_DIR, name + ".ilqgm"))
    q, v, u, w = scenario_states(name, n, seed=7, oracle=o, om=om, roll=roll)
    if name == "hopper":  # the reference test's scenario: 500 passive steps from qpos0, then ctrl -= 0.1
        q0 = np.array([[0, 1.25, 0, 0, 0, 0.0]]); z = np.zeros((1, 6)); u0 = np.zeros((1, 3))
        qs, vs, ws, _ = o.step_batch(om, q0, z, u0, z.copy(), 500)
        q[0], v[0], w[0], u[0] = qs[0], vs[0], ws[0], -0.1
    cost = o.make_cost(**COSTS[name])
    deriv, qacc, _ = o.fd_batch(om, q, v, u, w, cost)
    q1, v1, w1, a1 = o.step_batch(om, q, v, u, w, 1)
    np.savez(os.path.join(out, f"fd_{name}.npz"), qpos=q, qvel=v, ctrl=u, warm=w, cost=cost, deriv=deriv, qacc=qacc,
             step_qpos=q1, step_qvel=v1, step_warm=w1, step_qacc=a1)
    print(name, "knots", n, "max|deriv|", np.abs(deriv).max())

# ---- outputs of the REFERENCE'S OWN classes (verbatim /root/reference sources compiled against the shims: oracle/_ref) for the
# pendulum MPC run of BASELINE configs[0]: closed-loop trace, final nominal, gains and value model.  /root/reference and oracle/_ref do
# not exist on the GPU box; these vectors let the GPU tests compare with the reference's own code in one hop.  One reference ILQR
# instance per process (function-local statics, ilqr.h:137-140 — quirk Q13), hence the subprocesses.
import json  # noqa: E402
import subprocess  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "libref_fd.so")
REF_SCRIPT = r"""
import sys, os, json
sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, ctypes as C, mjo_py as o
R = C.CDLL(os.path.join(%(root)r, "oracle", "_ref", "libref_fd.so"))
m = o.Model(os.path.join(%(root)r, "ilqg-mujoco_b200", "models", "inverted_pendulum.ilqgm"))
q0 = np.array(%(q0)r); v0 = np.array(%(v0)r); nmpc = %(nmpc)d; N = 20
a = dict(tr=np.zeros((nmpc, 5)), q=np.zeros((N + 1, 2)), v=np.zeros((N + 1, 2)), u=np.zeros((N + 1, 1)), K=np.zeros((N + 1, 4)),
         k=np.zeros((N + 1, 1)), V=np.zeros(16), vv=np.zeros(4))
assert R.ref_pendulum_mpc(m.ptr, o._p(q0), o._p(v0), nmpc, *[o._p(a[x]) for x in ("tr", "q", "v", "u", "K", "k", "V", "vv")]) == 0
print(json.dumps({k: x.tolist() for k, x in a.items()}))
"""
if os.path.exists(REF):
    cases = [([0.0, 0.0], [0.0, 0.0], 2), ([0.1, 0.2], [0.0, 0.0], 3), ([-0.3, -0.25], [0.2, -0.4], 2)]
    save = {"ncase": np.array(len(cases))}
    for c, (q0, v0, nmpc) in enumerate(cases):
        r = subprocess.run([sys.executable, "-c", REF_SCRIPT % dict(root=ROOT, q0=q0, v0=v0, nmpc=nmpc)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res = json.loads(r.stdout.strip().splitlines()[-1])
        save[f"c{c}_q0"] = np.array(q0); save[f"c{c}_v0"] = np.array(v0); save[f"c{c}_nmpc"] = np.array(nmpc)
        for k, x in res.items():
            save[f"c{c}_{k}"] = np.array(x)
        print("reference classes, pendulum MPC case", c, "final state", res["tr"][-1])
    np.savez(os.path.join(out, "mpc_inverted_pendulum_reference_classes.npz"), **save)
else:
    print("oracle/_ref not built: mpc_inverted_pendulum_reference_classes.npz left as it is")
