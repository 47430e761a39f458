"""Anchors of the oracle's integrators (mj_step, /root/reference/inc/ilqr.h:86,128 calls it for every knot of a rollout).

  * inverted pendulum (integrator="RK4", timestep 0.02): twenty-five steps against (a) the classical four-stage Runge-Kutta scheme written out
    here on the CLOSED-FORM cart-pole equations of tests/test_oracle_anchors.py (control held over the step) — same scheme, same step:
    agreement to round-off; (b) the exact trajectory of that ODE (scipy DOP853 at 1e-12): the fourth-order error, four orders of
    magnitude below what a first-order scheme leaves at this step size;
  * hopper in flight (integrator="Euler", timestep 0.002, joint damping 1): one step against the semi-implicit Euler update MuJoCo 2.x
    documents — (M + h diag(damping)) a' = M a, v' = v + h a', q' = q + h v' — assembled here from the oracle's own qM and qacc and the
    damping attributes of the XML text; and forty steps in one call against that update repeated."""
import math
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from test_oracle_anchors import capsule

RES = "/root/reference/res"


def cartpole_rhs():
    mcart, _ = capsule(0.1, 0.1)
    mp, Iyy = capsule(0.049, math.hypot(0.001, 0.6) / 2)
    cx, cz = 0.0005, 0.3

    def f(x, u):
        th, sd, thd = x[1], x[2], x[3]
        a = -cx * math.sin(th) + cz * math.cos(th); b = -cx * math.cos(th) - cz * math.sin(th)
        M = np.array([[mcart + mp, mp * a], [mp * a, mp * (cx * cx + cz * cz) + Iyy]])
        rhs = np.array([100 * u - sd - mp * b * thd * thd, -thd - mp * 9.81 * b])
        return np.concatenate([x[2:], np.linalg.solve(M, rhs)])
    return f


def test_pendulum_rk4_steps_against_the_closed_form_ode(oracle, omodels):
    from scipy.integrate import solve_ivp
    m = omodels["inverted_pendulum"]
    f = cartpole_rhs()
    h, nsteps = 0.02, 25
    rng = np.random.default_rng(5)
    worst_scheme = worst_exact = worst_euler = 0.0
    for trial in range(8):
        x0 = np.concatenate([rng.uniform(-0.2, 0.2, 1), rng.uniform(-0.2, 0.2, 1), rng.normal(0, 0.2, 2)])
        u = float(rng.uniform(-0.3, 0.3))
        q, v, _, _ = oracle.step_batch(m, x0[None, :2], x0[None, 2:], np.array([[u]]), np.zeros((1, 2)), nsteps)
        got = np.concatenate([q[0], v[0]])
        assert abs(got[0]) < 0.9 and abs(got[1]) < 1.4           # inside the joint ranges (the pole falls monotonically): no limit row was ever active
        x = x0.copy(); xe = x0.copy()
        for _ in range(nsteps):                                   # classical RK4, control held over the step; explicit Euler beside it
            k1 = f(x, u); k2 = f(x + 0.5 * h * k1, u); k3 = f(x + 0.5 * h * k2, u); k4 = f(x + h * k3, u)
            x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
            xe = xe + h * f(xe, u)
        exact = solve_ivp(lambda t, y: f(y, u), (0, h * nsteps), x0, method="DOP853", rtol=1e-12, atol=1e-14).y[:, -1]
        worst_scheme = max(worst_scheme, np.abs(got - x).max())
        worst_exact = max(worst_exact, np.abs(got - exact).max())
        worst_euler = max(worst_euler, np.abs(xe - exact).max())
    assert worst_scheme < 1e-11                                   # the same scheme on the same equations
    assert worst_exact < 1e-5 and worst_exact < 1e-3 * worst_euler


def hopper_damping():
    path = os.path.join(RES, "hopper.xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    dflt = float(root.find("default").find("joint").get("damping", "0"))
    return np.array([float(j.get("damping", dflt)) for j in root.iter("joint") if j.get("name")])


def test_hopper_euler_step_is_semi_implicit_in_the_damping(oracle, omodels):
    m = omodels["hopper"]
    damp = hopper_damping()
    assert damp.shape == (6,) and (damp[:3] == 0).all() and (damp[3:] == 1).all()
    h = 0.002
    rng = np.random.default_rng(6)
    for trial in range(6):
        q = np.array([0.0, rng.uniform(1.6, 2.0), rng.uniform(-0.3, 0.3), *rng.uniform(-0.4, 0.0, 2), rng.uniform(-0.3, 0.3)])   # in the air
        v = rng.normal(0, 1.0, 6); u = rng.uniform(-0.5, 0.5, 3)
        d = oracle.dump(m, q, v, u)
        assert d["nefc"] == 0
        a2 = np.linalg.solve(d["qM"] + h * np.diag(damp), d["qM"] @ d["qacc"])
        v2 = v + h * a2
        q2 = q + h * v2
        qo, vo, _, _ = oracle.step_batch(m, q[None], v[None], u[None], np.zeros((1, 6)), 1)
        assert np.allclose(vo[0], v2, rtol=0, atol=1e-13) and np.allclose(qo[0], q2, rtol=0, atol=1e-14)


def test_hopper_euler_trajectory_is_that_update_repeated(oracle, omodels):
    """Forty steps in one call = forty times the update above (each assembled from that state's own qM and qacc): the stepping loop adds
    nothing of its own (no sub-stepping, no other order of the position and velocity updates)."""
    m = omodels["hopper"]
    damp = hopper_damping()
    h, nsteps = 0.002, 40
    q = np.array([0.0, 3.0, 0.1, -1.0, -1.2, 0.0]); v = np.array([0.5, 1.0, 0.3, -1.5, 2.0, 0.8]); u = np.array([0.2, -0.1, 0.3])   # joints far from their limits
    qo, vo, _, _ = oracle.step_batch(m, q[None], v[None], u[None], np.zeros((1, 6)), nsteps)
    for _ in range(nsteps):
        d = oracle.dump(m, q, v, u)
        assert d["nefc"] == 0                                     # no contact, no joint at its limit
        v = v + h * np.linalg.solve(d["qM"] + h * np.diag(damp), d["qM"] @ d["qacc"])
        q = q + h * v
    assert np.allclose(qo[0], q, rtol=0, atol=1e-12) and np.allclose(vo[0], v, rtol=0, atol=1e-11)


# ------------------------------------------------------------------ quaternions: mju_quatIntegrate and the free joint's position update
def qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def q2m(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def quat_integrate(oracle, q, w, s):
    import ctypes as C
    out = np.array(q, np.float64)
    oracle.lib().mjo_quat_integrate(oracle._p(out), oracle._p(np.ascontiguousarray(w, np.float64)), C.c_double(s))
    return out


def test_quat_integrate_is_a_rotation_about_a_body_axis(oracle):
    """mju_quatIntegrate(q, w, s) (the FD perturbation of ball / free rotations, /root/reference/src/mjderivative.cpp:152-169, and
    mj_integratePos): q (x) [cos(|w| s / 2), sin(|w| s / 2) w / |w|] — in matrices R(q') = R(q) expm([w s]x), the angular velocity
    expressed in the BODY frame; unit norm kept; zero velocity is the identity."""
    from scipy.linalg import expm
    rng = np.random.default_rng(8)
    for _ in range(50):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w = rng.normal(size=3) * 10.0 ** rng.uniform(-6, 1); s = 10.0 ** rng.uniform(-6, 0)
        got = quat_integrate(oracle, q, w, s)
        ang = np.linalg.norm(w) * s
        want = qmul(q, np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / np.linalg.norm(w)]))
        assert np.allclose(got, want, rtol=0, atol=1e-14)
        assert abs(np.linalg.norm(got) - 1) < 1e-14
        W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) * s
        assert np.allclose(q2m(got), q2m(q) @ expm(W), rtol=0, atol=1e-12)
    q = np.array([0.5, -0.5, 0.5, 0.5])
    assert np.array_equal(quat_integrate(oracle, q, np.zeros(3), 1.0), q)


def test_humanoid_euler_step_with_its_free_joint(oracle, omodels, pkg):
    """One step of the humanoid (integrator Euler, timestep 0.005) from random moving states, contacts and joint limits included:
    (M + h diag(damping)) a' = M a with the damping attributes of the XML text, v' = v + h a', and mj_integratePos — the root's position
    by the world-frame linear velocity, its quaternion by mju_quatIntegrate with the body-frame angular velocity, scalar joints added."""
    path = os.path.join(RES, "humanoid.xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    dflt = float(root.find("default").find("joint").get("damping", "0"))
    damp = np.concatenate([np.zeros(6), [float(j.get("damping", dflt)) for j in root.iter("joint") if j.get("name")]])
    m = omodels["humanoid"]
    pm = pkg.Model.named("humanoid")
    assert damp.shape == (27,)
    h = 0.005
    rng = np.random.default_rng(9)
    rngs = pm.field("jnt_range").reshape(-1, 2)[:pm.njnt]
    for trial in range(6):
        q = pm.field("qpos0")[:28].copy()
        q[2] = rng.uniform(0.3, 1.6)
        w = rng.normal(0, 1.0, 3); ang = np.linalg.norm(w)
        q[3:7] = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])
        q[7:] = rng.uniform(rngs[1:, 0], rngs[1:, 1])
        v = rng.normal(0, 1.0, 27); u = rng.uniform(-0.4, 0.4, 21)
        d = oracle.dump(m, q, v, u, iterations=50, tolerance=1e-10)      # the XML's own solver settings, as mj_step uses them
        a2 = np.linalg.solve(d["qM"] + h * np.diag(damp), d["qM"] @ d["qacc"])
        v2 = v + h * a2
        q2 = q.copy()
        q2[:3] += h * v2[:3]
        q2[3:7] = quat_integrate(oracle, q[3:7], v2[3:6], h)
        q2[7:] += h * v2[6:]
        qo, vo, _, _ = oracle.step_batch(m, q[None], v[None], u[None], np.zeros((1, 27)), 1)
        assert np.allclose(vo[0], v2, rtol=0, atol=1e-9 * max(1.0, np.abs(v2).max())), trial
        assert np.allclose(qo[0], q2, rtol=0, atol=1e-11 * max(1.0, np.abs(v2).max())), trial
