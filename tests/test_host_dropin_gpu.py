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
