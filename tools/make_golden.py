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
