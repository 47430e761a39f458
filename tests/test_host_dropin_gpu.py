"""The host-language mirror of the reference's classes (ilqg-mujoco_b200/host: calcMJDerivatives, Differentiator<nv,nu>,
ILQR<nv,nu,N>, InvertedPendulum, cpMjData, mj_step on the GPU) driven exactly like the reference's programs:
  - cmd/basic.cpp's MPC loop (BASELINE configs[0]) -> compared with the oracle's restated MPC run
    (which test_oracle_ilqr.py ties to the reference's own ilqr.h compiled verbatim)
  - tst/test_derivatives.cpp's Differentiator<6,3> scenario -> compared with the oracle's deriv and the reference's A/B assembly."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
PKG = os.path.join(ROOT, "ilqg-mujoco_b200")


@pytest.fixture(scope="module")
def hostlib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    path = os.path.join(PKG, "libilqg_host.so")
    assert os.path.exists(path), "libilqg_host.so missing: run __graft_entry__.build()"
    return C.CDLL(path)


@pytest.mark.parametrize("q0,v0,nmpc", [([0.0, 0.0], [0.0, 0.0], 2), ([0.1, 0.2], [0.0, 0.0], 3)])
def test_pendulum_mpc_matches_oracle(hostlib, oracle, omodels, q0, v0, nmpc):
    N = 20
    om = omodels["inverted_pendulum"]
    q0 = np.array(q0); v0 = np.array(v0)

    def bufs():
        return dict(tr=np.zeros((nmpc, 5)), q=np.zeros((N + 1, 2)), v=np.zeros((N + 1, 2)), u=np.zeros((N + 1, 1)), K=np.zeros((N + 1, 4)),
                    k=np.zeros((N + 1, 1)), V=np.zeros(16), vv=np.zeros(4))
    a, b = bufs(), bufs()
    path = os.path.join(PKG, "models", "inverted_pendulum.ilqgm").encode()
    rc = hostlib.ilqg_host_pendulum_mpc(path, oracle._p(q0), oracle._p(v0), nmpc, *[oracle._p(a[x]) for x in ("tr", "q", "v", "u", "K", "k", "V", "vv")])
    assert rc == 0
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1])
    oracle.lib().mjo_mpc_run(om.ptr, oracle._p(q0), oracle._p(v0), 10, N, 10, nmpc, oracle._p(cost), None, 0, 1, oracle._p(b["tr"]), None, None,
                             *[oracle._p(b[x]) for x in ("q", "v", "u", "K", "k", "V", "vv")])
    assert np.isfinite(a["tr"]).all()
    assert np.allclose(a["tr"], b["tr"], rtol=1e-6, atol=1e-7)           # closed-loop state / first control after each MPC step
    assert np.allclose(a["q"], b["q"], rtol=1e-6, atol=1e-7) and np.allclose(a["u"], b["u"], rtol=1e-5, atol=1e-6)
    for key in ("K", "k", "V", "vv"):
        x, y = (a[key][1:], b[key][1:]) if key in ("K", "k") else (a[key], b[key])
        assert np.abs(x - y).max() <= 1e-5 * np.abs(y).max(), key


@pytest.mark.parametrize("case", [0, 1, 2])
def test_pendulum_mpc_matches_the_reference_classes_golden(hostlib, oracle, case):
    """The same MPC run against tests/golden/mpc_inverted_pendulum_reference_classes.npz: what the REFERENCE'S OWN InvertedPendulum / ILQR /
    Differentiator / calcMJDerivatives (verbatim sources compiled against the shims, tools/make_golden.py) leave behind — closed-loop trace,
    final nominal, gains, value model.  GPU drop-in vs the reference's code in one hop (the physics under both is the restated MuJoCo
    subset); tolerances as in test_pendulum_mpc_matches_oracle."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "mpc_inverted_pendulum_reference_classes.npz"))
    N = 20
    q0, v0, nmpc = g[f"c{case}_q0"].copy(), g[f"c{case}_v0"].copy(), int(g[f"c{case}_nmpc"])
    a = dict(tr=np.zeros((nmpc, 5)), q=np.zeros((N + 1, 2)), v=np.zeros((N + 1, 2)), u=np.zeros((N + 1, 1)), K=np.zeros((N + 1, 4)),
             k=np.zeros((N + 1, 1)), V=np.zeros(16), vv=np.zeros(4))
    path = os.path.join(PKG, "models", "inverted_pendulum.ilqgm").encode()
    assert hostlib.ilqg_host_pendulum_mpc(path, oracle._p(q0), oracle._p(v0), nmpc, *[oracle._p(a[x]) for x in ("tr", "q", "v", "u", "K", "k", "V", "vv")]) == 0
    b = {k: g[f"c{case}_{k}"] for k in a}
    assert np.allclose(a["tr"], b["tr"], rtol=1e-6, atol=1e-7)
    assert np.allclose(a["q"], b["q"], rtol=1e-6, atol=1e-7) and np.allclose(a["u"], b["u"].reshape(a["u"].shape), rtol=1e-5, atol=1e-6)
    for key in ("K", "k", "V", "vv"):
        x, y = a[key], b[key].reshape(a[key].shape)
        if key in ("K", "k"):
            x, y = x[1:], y[1:]
        assert np.abs(x - y).max() <= 1e-5 * np.abs(y).max(), key


def test_pendulum_mpc_one_device_call_per_step_equals_the_reference_cadence(hostlib, oracle, monkeypatch):
    """InvertedPendulum::forward runs its ten iterations as ONE device call (ILQR::iterate(int): the class's own quadratic step cost
    handed over as an ilqg_cost, the chain replayed as a CUDA graph, public members synchronised once); ILQG_MIRROR_HOST_COST=1 keeps the
    reference's cadence — one iterate() at a time, cost rows from the host function, every member current after every phase.  Same
    arithmetic either way: closed-loop trace, nominal, gains and value model agree bit for bit."""
    N, nmpc = 20, 4
    q0 = np.array([0.1, 0.2]); v0 = np.array([0.0, 0.1])
    path = os.path.join(PKG, "models", "inverted_pendulum.ilqgm").encode()
    res = []
    for host_cost in (False, True):
        if host_cost:
            monkeypatch.setenv("ILQG_MIRROR_HOST_COST", "1")
        a = dict(tr=np.zeros((nmpc, 5)), q=np.zeros((N + 1, 2)), v=np.zeros((N + 1, 2)), u=np.zeros((N + 1, 1)), K=np.zeros((N + 1, 4)),
                 k=np.zeros((N + 1, 1)), V=np.zeros(16), vv=np.zeros(4))
        assert hostlib.ilqg_host_pendulum_mpc(path, oracle._p(q0), oracle._p(v0), nmpc, *[oracle._p(a[x]) for x in ("tr", "q", "v", "u", "K", "k", "V", "vv")]) == 0
        res.append(a)
    monkeypatch.delenv("ILQG_MIRROR_HOST_COST")
    for key in res[0]:
        assert np.array_equal(res[0][key], res[1][key]), key
    assert np.isfinite(res[0]["tr"]).all() and np.abs(res[0]["K"]).max() > 0


def test_hopper_differentiator_scenario(hostlib, oracle, omodels):
    om = omodels["hopper"]
    A = np.zeros((12, 12), order="F"); B = np.zeros((12, 3), order="F"); deriv = np.zeros(105); res = np.zeros(12)
    path = os.path.join(PKG, "models", "hopper.ilqgm").encode()
    rc = hostlib.ilqg_host_hopper_differentiator(path, 500, oracle._p(A), oracle._p(B), oracle._p(deriv), oracle._p(res))
    assert rc == 0
    q = np.array([[0, 1.25, 0, 0, 0, 0.0]]); z = np.zeros((1, 6)); u = np.zeros((1, 3))
    q, v, w, _ = oracle.step_batch(om, q, z, u, z.copy(), 500)
    d_ref, _, _ = oracle.fd_batch(om, q, v, u - 0.1, w, oracle.make_cost(q1=[1.0]))
    scale = max(1.0, np.abs(d_ref).max())
    assert np.abs(deriv - d_ref[0]).max() <= 1e-5 * scale      # 500 GPU steps vs 500 oracle steps, then FD
    dt = 0.002
    # the reference's assembly (differentiator.h:68-71,89-92): column-major views of the row-major blocks
    assert np.allclose(A[:6, :6], np.eye(6)) and np.allclose(A[:6, 6:], dt * np.eye(6))
    assert np.allclose(A[6:, :6], dt * deriv[:36].reshape(6, 6, order="F"))
    assert np.allclose(A[6:, 6:], np.eye(6) + dt * deriv[36:72].reshape(6, 6, order="F"))
    assert np.allclose(B[6:], dt * deriv[72:90].reshape(6, 3, order="F")) and np.allclose(B[:6], 0)
    assert np.isfinite(res).all()   # the reference's test only prints this residual (it is not small: SURVEY F4)


def test_headless_base_demo_runs():
    out = subprocess.run([os.path.join(PKG, "base"), os.path.join(PKG, "models", "inverted_pendulum.ilqgm"), "3"], capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 3
    import json
    rec = json.loads(lines[-1])
    assert rec["step"] == 2 and abs(rec["time"] - 13 * 0.02) < 1e-9 and np.isfinite(rec["ctrl"][0])


def test_hopper_task_mpc_matches_oracle(hostlib, oracle, omodels):
    """the Hopper task class (host/hopper/hopper.h) — MPC through contacts with the opt-in corrected A/B
    layout, the backtracking ladder and the mu schedule — against the oracle's restated MPC loop with the same switches."""
    om = omodels["hopper"]
    nmpc, N, niter = 3, 20, 10
    tr = np.zeros((nmpc, 15)); Jt = np.zeros((nmpc, niter))
    path = os.path.join(PKG, "models", "hopper.ilqgm").encode()
    # start in stance: the hopper dropped from its rest pose and settled on the ground (400 passive steps)
    z6 = np.zeros((1, 6))
    q0, v0, _, _ = oracle.step_batch(om, np.array([[0, 1.25, 0, 0, 0, 0.0]]), z6, np.zeros((1, 3)), z6.copy(), 400)
    q0, v0 = q0[0].copy(), v0[0].copy()
    assert hostlib.ilqg_host_hopper_mpc(path, oracle._p(q0), oracle._p(v0), nmpc, oracle._p(tr), oracle._p(Jt)) == 0
    cost = oracle.make_cost(q2=[0, 5, 1], q1=[0, -12.5], v1=[-1.0], v2=[0.05] * 6, u2=[0.01] * 3)
    al = np.array([0.5 ** a for a in range(6)])
    tr_o = np.zeros((nmpc, 15)); Jt_o = np.zeros((nmpc, niter)); acc_o = np.zeros((nmpc, niter), np.int32)
    L = oracle.lib()
    L.mjo_ilqr_set_corrected_layout(1)
    L.mjo_ilqr_set_mu_schedule(C.c_double(2.0), C.c_double(1.0), C.c_double(1e8))
    try:
        L.mjo_mpc_run(om.ptr, oracle._p(q0), oracle._p(v0), 10, N, niter, nmpc, oracle._p(cost), oracle._p(al), 6, 0, oracle._p(tr_o), oracle._p(Jt_o),
                      oracle._p(acc_o), None, None, None, None, None, None, None)
    finally:
        L.mjo_ilqr_set_corrected_layout(0)
        L.mjo_ilqr_set_mu_schedule(C.c_double(1.0), C.c_double(1e-6), C.c_double(1e10))
    assert np.isfinite(tr).all() and np.isfinite(Jt).all()
    assert (acc_o >= 0).sum() >= nmpc            # the ladder accepts steps: the extensions make iLQR usable on this model
    for J in (Jt, Jt_o):                         # the cost never increases within an MPC step, on either side
        assert (np.diff(J, axis=1) <= 1e-9 * np.abs(J[:, :-1])).all()
    # first MPC step: ten iterations through contacts agree to round-off, and so does the state the plant reaches
    assert np.allclose(Jt[0], Jt_o[0], rtol=1e-8, atol=1e-9)
    assert np.allclose(tr[0], tr_o[0], rtol=1e-6, atol=1e-8)
    # Later MPC steps run through stick/slip transitions where the ladder's discrete decisions flip under 1e-11 perturbations of
    # the initial state (checked on the oracle alone: the accepted-alpha sequence of step 1 changes), so only the first
    # iterations of step 1 — before the first marginal decision — are compared, and the overall progress.
    assert np.allclose(Jt[1, :5], Jt_o[1, :5], rtol=1e-8, atol=1e-9)
    assert Jt[-1, -1] < Jt[0, 0] - 1.0 and Jt_o[-1, -1] < Jt_o[0, 0] - 1.0


def test_host_cost_gradient_on_quaternion_dofs(hostlib, oracle, omodels):
    """/root/reference/src/mjderivative.cpp:161-174: the cost row of a free joint's rotational dofs comes from a tangent-space
    perturbation (mju_quatIntegrate) followed by the caller's stepCostFn — a host cost that depends on the root orientation
    (uprightness) must see it.  Checked against the same forward difference done in numpy."""
    om = omodels["humanoid"]
    rng = np.random.default_rng(3)
    q = np.zeros(28); q[2] = 1.3; q[3:7] = [0.9, 0.2, -0.3, 0.1]; q[3:7] /= np.linalg.norm(q[3:7]); q[7:] = rng.uniform(-0.1, 0.1, 21)
    v = rng.normal(0, 0.2, 27); u = rng.uniform(-0.3, 0.3, 21)
    deriv = np.zeros(2100)
    path = os.path.join(PKG, "models", "humanoid.ilqgm").encode()
    assert hostlib.ilqg_host_freejoint_cost_rows(path, oracle._p(q), oracle._p(v), oracle._p(u), oracle._p(deriv)) == 0

    def cost(q, v, u):
        return (1.0 - (1.0 - 2.0 * (q[4] ** 2 + q[5] ** 2))) + 0.5 * q[2] ** 2 + 0.1 * v[4] + 0.01 * u[2] ** 2

    def qmul(a, b):
        return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])
    eps, c0 = 1e-6, cost(q, v, u)
    rows = np.zeros(27 + 27 + 21)
    for i in range(27):
        qp = q.copy()
        if 3 <= i < 6:
            ax = np.zeros(3); ax[i - 3] = 1.0
            qp[3:7] = qmul(q[3:7] / np.linalg.norm(q[3:7]), np.concatenate([[np.cos(eps / 2)], ax * np.sin(eps / 2)]))
        else:
            qp[i if i < 3 else i + 1] += eps
        rows[i] = (cost(qp, v, u) - c0) / eps
    for i in range(27):
        vp = v.copy(); vp[i] += eps
        rows[27 + i] = (cost(q, vp, u) - c0) / eps
    for i in range(21):
        up = u.copy(); up[i] += eps
        rows[54 + i] = (cost(q, v, up) - c0) / eps
    got = deriv[-75:]
    assert np.abs(got[3:6]).max() > 1e-2            # the orientation rows are not zero (they were, silently, in round 1)
    assert np.allclose(got, rows, rtol=0, atol=2e-9)  # two evaluations of the same differences: round-off / eps
    # and the dynamics blocks of the same call are the oracle's
    d_ref, _, _ = oracle.fd_batch(om, q[None], v[None], u[None], np.zeros((1, 27)), None)
    from test_fd_gpu import assert_deriv_close
    both = np.concatenate([deriv[:-75], d_ref[0, -75:]])[None]
    assert_deriv_close(both, d_ref, 27, 21, tol=1e-5)


@pytest.mark.parametrize("use_xfrc", [0, 1])
def test_applied_forces_are_refused_not_ignored(hostlib, use_xfrc):
    """qfrc_applied / xfrc_applied are part of the knot (/root/reference/src/util.cpp:10-11) but not of the GPU pipeline:
    calcMJDerivatives, mj_step and mj_forward refuse a state that carries them (ILQG_ERR_UNSUPPORTED through mju_error) and
    compute nothing."""
    msg = C.create_string_buffer(256)
    path = os.path.join(PKG, "models", "hopper.ilqgm").encode()
    n = hostlib.ilqg_host_applied_force_probe(path, use_xfrc, msg, 256)
    assert n == 3, (n, msg.value)
    assert b"applied" in msg.value and b"(5)" in msg.value


def test_humanoid_task_mpc_matches_oracle(hostlib, oracle, omodels):
    """the Humanoid task class (host/humanoid/humanoid.h) — tangent-space iLQR on the warp-cooperative engine
    (nq = 28 != nv = 27: the reference's own ILQR is undefined here, quirk Q9) — against the oracle's restated MPC loop with
    the same extension: the first MPC step's cost trace and the state the plant reaches."""
    om = omodels["humanoid"]
    nmpc, N, niter = 2, 10, 4
    tr = np.zeros((nmpc, 76)); Jt = np.zeros((nmpc, niter))
    path = os.path.join(PKG, "models", "humanoid.ilqgm").encode()
    q0 = np.zeros(28); q0[2] = 1.45; q0[3] = 1.0; q0[3:7] += [0.0, 0.03, -0.02, 0.01]; q0[3:7] /= np.linalg.norm(q0[3:7])
    q0[7:] = np.linspace(-0.1, 0.1, 21)
    v0 = np.linspace(-0.2, 0.2, 27)
    assert hostlib.ilqg_host_humanoid_mpc(path, oracle._p(q0), oracle._p(v0), nmpc, oracle._p(tr), oracle._p(Jt)) == 0
    cost = oracle.make_cost(q2=[0, 0, 2.0, 0, 1, 1, 0], q1=[0, 0, -5.2], v2=[0.05] * 27, u2=[0.02] * 21)
    al = np.array([0.5 ** a for a in range(4)])
    tr_o = np.zeros((nmpc, 76)); Jt_o = np.zeros((nmpc, niter)); acc_o = np.zeros((nmpc, niter), np.int32)
    L = oracle.lib()
    L.mjo_ilqr_set_corrected_layout(1)
    L.mjo_ilqr_set_mu_schedule(C.c_double(2.0), C.c_double(1.0), C.c_double(1e8))
    try:
        L.mjo_mpc_run(om.ptr, oracle._p(q0), oracle._p(v0), 10, N, niter, nmpc, oracle._p(cost), oracle._p(al), 4, 0, oracle._p(tr_o), oracle._p(Jt_o),
                      oracle._p(acc_o), None, None, None, None, None, None, None)
    finally:
        L.mjo_ilqr_set_corrected_layout(0)
        L.mjo_ilqr_set_mu_schedule(C.c_double(1.0), C.c_double(1e-6), C.c_double(1e10))
    assert np.isfinite(tr).all() and np.isfinite(Jt).all()
    for J in (Jt, Jt_o):
        assert (np.diff(J, axis=1) <= 1e-9 * np.abs(J[:, :-1])).all()      # the ladder never accepts a worse trajectory
    assert np.allclose(Jt[0], Jt_o[0], rtol=1e-7, atol=1e-8)
    assert np.allclose(tr[0], tr_o[0], rtol=1e-6, atol=1e-7)
    assert np.allclose(np.linalg.norm(tr[:, 3:7], axis=1), 1.0, atol=1e-9)
