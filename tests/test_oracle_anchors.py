"""What pins the CPU oracle (MuJoCo and reference golden vectors do not exist here — SURVEY.md §8c):
closed-form cart-pole dynamics, compiler-vs-CRBA mass matrices, energy conservation, KKT conditions of the
contact solve, the reference's own FD driver compiled verbatim, and committed golden vectors."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from conftest import ROOT, scenario_states

GOLD = os.path.join(ROOT, "tests", "golden")


def capsule(r, h):
    H = 2 * h
    mc, ms = 1000 * math.pi * r * r * H, 1000 * 4 / 3 * math.pi * r ** 3
    return mc + ms, mc * (3 * r * r + H * H) / 12 + ms * (0.4 * r * r + H * H / 4 + 3 * H * r / 8)


def test_cartpole_matches_closed_form(oracle, omodels):
    """qacc of res/inverted_pendulum.xml against Lagrangian cart-pole equations derived by hand."""
    m = omodels["inverted_pendulum"]
    mcart, _ = capsule(0.1, 0.1)
    mp, Iyy = capsule(0.049, math.hypot(0.001, 0.6) / 2)
    cx, cz = 0.0005, 0.3
    rng = np.random.default_rng(0)
    n = 256
    q = np.stack([rng.uniform(-.9, .9, n), rng.uniform(-1.5, 1.5, n)], 1)
    v = rng.normal(size=(n, 2)); u = rng.uniform(-3, 3, (n, 1))
    qacc, _ = oracle.forward_batch(m, q, v, u, np.zeros((n, 2)))
    for i in range(n):
        th, sd, thd = q[i, 1], v[i, 0], v[i, 1]
        a = -cx * math.sin(th) + cz * math.cos(th); b = -cx * math.cos(th) - cz * math.sin(th)
        M = np.array([[mcart + mp, mp * a], [mp * a, mp * (cx * cx + cz * cz) + Iyy]])
        rhs = np.array([100 * u[i, 0] - sd - mp * b * thd * thd, -thd - mp * 9.81 * b])
        ref = np.linalg.solve(M, rhs)
        assert np.allclose(qacc[i], ref, rtol=1e-12, atol=1e-12)


def mass_bias(oracle, m, q, v):
    M = np.zeros((m.nv, m.nv)); b = np.zeros(m.nv); e = C.c_double()
    oracle.lib().mjo_debug_mass_bias(m.ptr, oracle._p(np.ascontiguousarray(q)), oracle._p(np.ascontiguousarray(v)), oracle._p(M), oracle._p(b), C.byref(e))
    return M, b, e.value


@pytest.mark.parametrize("name", ["inverted_pendulum", "hopper", "humanoid"])
def test_crba_mass_matrix_agrees_with_compiler_constants(pkg, oracle, omodels, name):
    """dof_invweight0 / meaninertia were computed by the compiler from a dense Jacobian formulation;
    the oracle's CRBA at qpos0 must give the same numbers."""
    om = omodels[name]; pm = pkg.Model.named(name)
    q0 = pm.field("qpos0")[:om.nq].copy()
    M, _, _ = mass_bias(oracle, om, q0, np.zeros(om.nv))
    assert np.allclose(M, M.T, atol=1e-13)
    assert np.linalg.eigvalsh(M).min() > 0
    assert np.trace(M) / om.nv == pytest.approx(pm.field("meaninertia")[0], rel=1e-12)
    Minv = np.linalg.inv(M)
    inv0 = pm.field("dof_invweight0")[:om.nv]
    jt = pm.field("jnt_type")[:om.njnt]
    d = np.diag(Minv).copy()
    if jt[0] == 0:  # free joint: translational / rotational averages
        d[0:3] = d[0:3].mean(); d[3:6] = d[3:6].mean()
    assert np.allclose(d, inv0, rtol=1e-10)


@pytest.mark.parametrize("name", ["hopper", "humanoid"])
def test_energy_is_conserved_without_dissipation(pkg, oracle, name):
    """Strip damping, springs, limits and contacts; integrate free flight with RK4: kinetic + potential
    energy must stay constant — checks CRBA, RNE bias, quaternion integration together."""
    pm = pkg.Model.named(name).copy()
    pm.field("dof_damping")[:] = 0; pm.field("jnt_stiffness")[:] = 0; pm.field("jnt_limited")[:] = 0
    pm.field("npair")[0] = 0; pm.field("integrator")[0] = 1; pm.field("timestep")[0] = 5e-4
    path = os.path.join(GOLD, f"_tmp_{name}.ilqgm")
    pm.buf.tofile(path)
    try:
        om = oracle.Model(path)
        q, v, u, w = scenario_states(name, 4, seed=3)
        u[:] = 0
        if name == "hopper":
            q[:, 1] += 3.0
        e0 = [mass_bias(oracle, om, q[i], v[i])[2] for i in range(4)]
        q1, v1, _, _ = oracle.step_batch(om, q, v, u, w, 400)
        e1 = [mass_bias(oracle, om, q1[i], v1[i])[2] for i in range(4)]
        ke = [0.5 * v[i] @ mass_bias(oracle, om, q[i], v[i])[0] @ v[i] for i in range(4)]
        for a, b, k in zip(e0, e1, ke):
            assert abs(a - b) < 1e-8 * max(1.0, abs(k)), (a, b)
        assert np.abs(q1 - q).max() > 1e-2  # it really moved
    finally:
        os.remove(path)


def debug_forward(oracle, om, q, v, u, w, iters=30, tol=0.0):
    qacc = np.zeros(om.nv); info = (C.c_int * 4)(); kkt = C.c_double(); f = np.zeros(320); dist = np.zeros(96)
    oracle.lib().mjo_debug_forward(om.ptr, oracle._p(q.copy()), oracle._p(v.copy()), oracle._p(u.copy()), oracle._p(w.copy()), iters,
                                   C.c_double(tol), oracle._p(qacc), info, C.byref(kkt), oracle._p(f), oracle._p(dist))
    return qacc, list(info), kkt.value, f[:info[1]], dist[:info[0]]


def test_hopper_contact_solve_satisfies_kkt(oracle, omodels):
    om = omodels["hopper"]
    q = np.array([[0, 1.25, 0, 0, 0, 0.0]]); z = np.zeros((1, 6)); u = np.zeros((1, 3))
    q, v, w, _ = oracle.step_batch(om, q, z, u, z.copy(), 300)  # settles on its foot
    qacc, info, kkt, f, dist = debug_forward(oracle, om, q[0], v[0], u[0], w[0])
    ncon, nefc, iters, nact = info
    assert ncon == 2 and nefc == 8            # foot capsule: two end spheres x 4 pyramid facets
    assert (dist < 0.001).all()               # inside the margin
    assert (f >= 0).all() and kkt < 1e-9      # unilateral forces, stationarity
    # at rest the normal load carries the weight: sum of facet forces = m g (facet normals are the contact normal)
    weight = 9.81 * (3.665191429 + 4.057890511 + 2.781356696 + 5.31557477)
    assert f.sum() == pytest.approx(weight, rel=2e-2)
    assert np.abs(qacc).max() < 5.0


def test_pendulum_joint_limit_rows(oracle, omodels):
    om = omodels["inverted_pendulum"]
    q = np.array([1.02, 0.3]); v = np.array([0.5, 0.0]); u = np.array([0.0]); w = np.zeros(2)
    qacc, info, kkt, f, _ = debug_forward(oracle, om, q, v, u, w)
    assert info[1] == 1 and f[0] > 0 and kkt < 1e-9    # slider beyond +1: one active limit row pushing back
    qacc_free, info2, _, _, _ = debug_forward(oracle, om, np.array([0.5, 0.3]), v, u, w)
    assert info2[1] == 0
    assert qacc[0] < qacc_free[0]


def test_fd_against_analytic_control_jacobian(oracle, omodels):
    """Smooth pendulum: d qacc / d ctrl = M^-1 [gear, 0] exactly; the FD block must match to round-off."""
    om = omodels["inverted_pendulum"]
    q, v, u, w = scenario_states("inverted_pendulum", 16, seed=5)
    deriv, qacc, _ = oracle.fd_batch(om, q, v, u, w, None)
    for i in range(16):
        M, _, _ = mass_bias(oracle, om, q[i], v[i])
        col = np.linalg.solve(M, np.array([100.0, 0.0]))
        blk = deriv[i, 8:10]       # 2nv^2 + i + j*nu, nu = 1
        assert np.allclose(blk, col, rtol=1e-7, atol=1e-7)
        # d qacc / d qvel carries the joint damping: -M^-1 diag(1,1) plus Coriolis terms; symmetric part check
        assert np.isfinite(deriv[i]).all()


@pytest.mark.parametrize("name", ["inverted_pendulum", "hopper", "humanoid"])
def test_restated_fd_equals_reference_driver(oracle, omodels, name):
    """The reference's own calcMJDerivatives (/root/reference/src/mjderivative.cpp, compiled verbatim against the
    shim) and the oracle's restatement of it produce bit-identical deriv buffers."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_fd.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    R = C.CDLL(ref)
    om = omodels[name]
    n = 6 if name != "humanoid" else 2
    q, v, u, w = scenario_states(name, n, seed=11, oracle=oracle, om=om, roll=60 if name != "humanoid" else 20)
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1], q1=[0, 0, 3])
    d1, _, _ = oracle.fd_batch(om, q, v, u, w, cost)
    d2 = np.zeros_like(d1)
    ncpu = R.ref_calc_derivatives_batch(om.ptr, n, oracle._p(q), oracle._p(v), oracle._p(u), oracle._p(w), oracle._p(cost), oracle._p(d2), 16)
    assert 1 <= ncpu <= 16
    assert np.array_equal(d1, d2)


@pytest.mark.parametrize("name", ["inverted_pendulum", "hopper", "humanoid"])
def test_golden_deriv_is_what_the_reference_driver_outputs(oracle, omodels, name):
    """The committed golden `deriv` blocks (tests/golden/fd_<model>.npz — what the GPU tests are compared with on a box that has neither
    /root/reference nor oracle/_ref) are bit for bit what the reference's own calcMJDerivatives writes for those knots."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_fd.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    R = C.CDLL(ref)
    g = np.load(os.path.join(GOLD, f"fd_{name}.npz"))
    om = omodels[name]
    q, v, u, w, cost = (np.ascontiguousarray(g[k]) for k in ("qpos", "qvel", "ctrl", "warm", "cost"))
    d = np.zeros_like(g["deriv"])
    ncpu = R.ref_calc_derivatives_batch(om.ptr, q.shape[0], oracle._p(q), oracle._p(v), oracle._p(u), oracle._p(w), oracle._p(cost), oracle._p(d), 16)
    assert 1 <= ncpu <= 16
    assert np.array_equal(d, g["deriv"])


@pytest.mark.parametrize("name", ["inverted_pendulum", "hopper", "humanoid"])
def test_oracle_reproduces_golden_vectors(oracle, omodels, name):
    g = np.load(os.path.join(GOLD, f"fd_{name}.npz"))
    om = omodels[name]
    deriv, qacc, _ = oracle.fd_batch(om, g["qpos"], g["qvel"], g["ctrl"], g["warm"], g["cost"])
    scale = max(1.0, np.abs(g["deriv"]).max())
    assert np.abs(deriv - g["deriv"]).max() <= 1e-9 * scale
    assert np.allclose(qacc, g["qacc"], rtol=1e-10, atol=1e-10)
    q1, v1, w1, a1 = oracle.step_batch(om, g["qpos"], g["qvel"], g["ctrl"], g["warm"], 1)
    assert np.allclose(q1, g["step_qpos"], rtol=1e-12, atol=1e-12) and np.allclose(v1, g["step_qvel"], rtol=1e-10, atol=1e-10)
