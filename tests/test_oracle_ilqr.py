"""The oracle's restatement of Differentiator / ILQR / InvertedPendulum against the REFERENCE's own classes
(/root/reference/inc/ilqr.h, inc/differentiator.h, src/inverted_pendulum/inverted_pendulum.cpp compiled verbatim
against the MuJoCo and Eigen shims, oracle/_ref).  BASELINE configs[0]: pendulum MPC at the default horizon."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

SCRIPT = r"""
import sys, os, json
sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, ctypes as C, mjo_py as o
L = o.lib(); R = C.CDLL(os.path.join(%(root)r, "oracle", "_ref", "libref_fd.so"))
m = o.Model(os.path.join(%(root)r, "ilqg-mujoco_b200", "models", "inverted_pendulum.ilqgm"))
q0 = np.array(%(q0)r); v0 = np.array(%(v0)r); nmpc = %(nmpc)d; N = 20
def bufs():
    return dict(tr=np.zeros((nmpc, 5)), q=np.zeros((N + 1, 2)), v=np.zeros((N + 1, 2)), u=np.zeros((N + 1, 1)), K=np.zeros((N + 1, 4)),
                k=np.zeros((N + 1, 1)), V=np.zeros(16), vv=np.zeros(4))
a = bufs(); b = bufs()
rc = R.ref_pendulum_mpc(m.ptr, o._p(q0), o._p(v0), nmpc, *[o._p(a[x]) for x in ("tr", "q", "v", "u", "K", "k", "V", "vv")])
assert rc == 0
cost = o.make_cost(q2=[1, 10], v2=[1, 10], u2=[1])
J = np.zeros((nmpc, 10))
L.mjo_mpc_run(m.ptr, o._p(q0), o._p(v0), 10, N, 10, nmpc, o._p(cost), None, 0, 1, o._p(b["tr"]), o._p(J), None,
              *[o._p(b[x]) for x in ("q", "v", "u", "K", "k", "V", "vv")])
res = {}
for key in a:
    x, y = a[key], b[key]
    if key in ("K", "k"): x, y = x[1:], y[1:]
    res[key] = float(np.abs(x - y).max() / max(1e-300, np.abs(x).max()))
res["finite"] = bool(np.isfinite(a["tr"]).all())
res["J0"] = J[0].tolist()
print(json.dumps(res))
"""


@pytest.mark.parametrize("q0,v0,nmpc", [([0.0, 0.0], [0.0, 0.0], 2), ([0.1, 0.2], [0.0, 0.0], 3), ([-0.3, -0.25], [0.2, -0.4], 2)])
def test_restated_mpc_equals_reference_classes(q0, v0, nmpc):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_fd.so")):
        pytest.skip("oracle/_ref not built")
    # one reference ILQR instance per process (function-local statics, ilqr.h:137-140 — quirk Q13)
    out = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT, q0=q0, v0=v0, nmpc=nmpc)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["finite"]
    assert res["tr"] < 1e-6 and res["q"] < 1e-6 and res["u"] < 1e-5
    for key in ("K", "k", "V", "vv"):
        assert res[key] < 1e-5, (key, res[key])
    assert res["J0"][-1] < res["J0"][0]          # ten iterations do reduce the cost from a perturbed start


@pytest.mark.parametrize("case", [0, 1, 2])
def test_restated_mpc_reproduces_the_reference_classes_golden(oracle, omodels, case):
    """The committed outputs of the reference's own classes (tests/golden/mpc_inverted_pendulum_reference_classes.npz, written by
    tools/make_golden.py where /root/reference exists) against the oracle's restated MPC loop — the same comparison as
    test_restated_mpc_equals_reference_classes, but one that also runs where oracle/_ref cannot be built."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "mpc_inverted_pendulum_reference_classes.npz"))
    om = omodels["inverted_pendulum"]
    N = 20
    q0, v0, nmpc = g[f"c{case}_q0"].copy(), g[f"c{case}_v0"].copy(), int(g[f"c{case}_nmpc"])
    b = dict(tr=np.zeros((nmpc, 5)), q=np.zeros((N + 1, 2)), v=np.zeros((N + 1, 2)), u=np.zeros((N + 1, 1)), K=np.zeros((N + 1, 4)),
             k=np.zeros((N + 1, 1)), V=np.zeros(16), vv=np.zeros(4))
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1])
    oracle.lib().mjo_mpc_run(om.ptr, oracle._p(q0), oracle._p(v0), 10, N, 10, nmpc, oracle._p(cost), None, 0, 1, oracle._p(b["tr"]), None, None,
                             *[oracle._p(b[x]) for x in ("q", "v", "u", "K", "k", "V", "vv")])
    for key in b:
        x, y = b[key], g[f"c{case}_{key}"].reshape(b[key].shape)
        if key in ("K", "k"):
            x, y = x[1:], y[1:]
        assert np.abs(x - y).max() <= (1e-5 if key in ("K", "k", "V", "vv", "u") else 1e-6) * max(1e-300, np.abs(y).max()), key


def test_linesearch_spec_reduces_to_reference_iterate(oracle, omodels):
    """A10: alphas = {1}, accept_always -> identical numbers to ILQR::iterate()."""
    om = omodels["inverted_pendulum"]
    rng = np.random.default_rng(1)
    q = rng.uniform(-.4, .4, (5, 2)); v = rng.normal(0, .3, (5, 2)); u = rng.uniform(-.2, .2, (5, 1)); w = np.zeros((5, 2))
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1])
    a = oracle.ilqr_run_batch(om, 20, 6, q, v, u, w, cost, alphas=None)
    b = oracle.ilqr_run_batch(om, 20, 6, q, v, u, w, cost, alphas=[1.0], accept_always=True)
    for key in ("J", "qpos", "ctrl", "K", "k", "V", "v"):
        assert np.array_equal(a[key], b[key]), key


def test_backtracking_is_monotone_and_ordered(oracle, omodels):
    om = omodels["inverted_pendulum"]
    rng = np.random.default_rng(2)
    q = rng.uniform(-.5, .5, (16, 2)); v = rng.normal(0, .5, (16, 2)); u = rng.uniform(-.2, .2, (16, 1)); w = np.zeros((16, 2))
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1])
    r = oracle.ilqr_run_batch(om, 20, 10, q, v, u, w, cost, alphas=[1, .5, .25, .125], accept_always=False, mu=10.0)
    J = r["J"]
    assert (np.diff(J, axis=1) <= 0).all()                      # accepted steps only ever decrease the cost
    assert ((r["accepted"] >= -1) & (r["accepted"] < 4)).all()
    rejected = r["accepted"] == -1
    assert (np.diff(J, axis=1)[rejected[:, 1:]] == 0).all()     # a rejected iteration keeps the nominal
