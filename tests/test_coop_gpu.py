"""The generic warp-per-rollout engine (csrc/coop.cuh): humanoid FD (BASELINE configs[2]: free-joint quaternion
perturbations, nv = 27, 161 collision pairs) against the oracle and the golden fixture, and — forced onto the small
models — against the same tolerances as the thread-per-rollout kernels."""
import os

import numpy as np
import pytest

from conftest import ROOT, scenario_states
from test_fd_gpu import assert_deriv_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def generic_handles(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    os.environ["ILQG_FORCE_GENERIC"] = "1"
    try:
        hs = {n: pkg.Handle(pkg.Model.named(n), 0) for n in ("inverted_pendulum", "hopper", "humanoid")}
    finally:
        del os.environ["ILQG_FORCE_GENERIC"]
    for h in hs.values():
        assert h.engine == "generic-warp-per-rollout"
    yield hs
    for h in hs.values():
        h.close()


def test_humanoid_uses_the_generic_engine_by_default(pkg):
    h = pkg.Handle(pkg.Model.named("humanoid"), 0)
    assert h.engine == "generic-warp-per-rollout"
    h.close()


@pytest.mark.parametrize("name", ["inverted_pendulum", "hopper", "humanoid"])
def test_golden_vectors(generic_handles, name):
    h = generic_handles[name]; m = h.model
    g = np.load(os.path.join(GOLD, f"fd_{name}.npz"))
    deriv, qacc, status = h.fd_batch_host(g["qpos"], g["qvel"], g["ctrl"], g["warm"], g["cost"])
    assert status.sum() == 0, status
    assert np.allclose(qacc, g["qacc"], rtol=1e-8, atol=1e-8)
    assert_deriv_close(deriv, g["deriv"], m.nv, m.nu, tol=1e-6 if name != "humanoid" else 1e-5)


@pytest.mark.parametrize("name,n,roll", [("inverted_pendulum", 70, 10), ("hopper", 64, 0), ("hopper", 40, 150), ("humanoid", 12, 0),
                                         ("humanoid", 10, 30), ("humanoid", 8, 80)])
def test_parity_with_oracle(generic_handles, oracle, omodels, name, n, roll):
    h = generic_handles[name]; m = h.model; om = omodels[name]
    q, v, u, w = scenario_states(name, n, seed=300 + roll, oracle=oracle, om=om, roll=roll)
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1], q1=[0.5, 0, 2])
    d_ref, a_ref, _ = oracle.fd_batch(om, q, v, u, w, cost)
    d_gpu, a_gpu, status = h.fd_batch_host(q, v, u, w, cost)
    assert status.sum() == 0, status
    assert np.allclose(a_gpu, a_ref, rtol=1e-8, atol=1e-8)
    assert_deriv_close(d_gpu, d_ref, m.nv, m.nu, tol=1e-6 if name != "humanoid" else 1e-5)


def test_humanoid_forward_and_step_match_oracle(generic_handles, oracle, omodels):
    h = generic_handles["humanoid"]; om = omodels["humanoid"]
    q, v, u, w = scenario_states("humanoid", 16, seed=77, oracle=oracle, om=om, roll=40)
    a_ref, _ = oracle.forward_batch(om, q, v, u, w)
    a_gpu, _ = h.forward_batch_host(q, v, u, w)
    assert np.allclose(a_gpu, a_ref, rtol=1e-5, atol=1e-5)       # XML tolerance 1e-10, 50 iterations
    q1, v1, w1, _ = oracle.step_batch(om, q, v, u, w, 10)
    q2, v2, w2, _ = h.step_batch_host(q, v, u, w, nsteps=10)
    assert np.allclose(q2, q1, rtol=1e-7, atol=1e-7) and np.allclose(v2, v1, rtol=1e-5, atol=1e-5)
    assert np.allclose(np.linalg.norm(q2[:, 3:7], axis=1), 1.0, atol=1e-12)   # quaternion stays normalised


def test_capacity_overflow_is_flagged(generic_handles, pkg):
    """A humanoid pushed deep into the floor produces more rows than the shared-memory budget: flagged, not silent."""
    h = generic_handles["humanoid"]
    q = np.zeros((1, 28)); q[0, 2] = 0.05; q[0, 3] = 1.0      # lying inside the ground plane: every geom touches
    v = np.zeros((1, 27)); u = np.zeros((1, 21)); w = np.zeros((1, 27))
    d, a, status = h.fd_batch_host(q, v, u, w, None)
    assert status[0] in (0, pkg.ERR_CAPACITY, pkg.ERR_NONFINITE)


def test_generic_engine_is_idempotent_and_chunk_invariant(generic_handles, handles, pkg):
    """The warp-cooperative engine linearises in passes of <= 8192 knots (its C-state scratch): a batch that crosses the pass
    boundary must equal the same knots computed alone, bit for bit, and the thread-per-rollout kernels to the FD tolerance."""
    import torch
    from ilqg_mujoco_b200 import workload as wl
    h = generic_handles["hopper"]; ht = handles["hopper"]
    q, v, u, w, _ = wl.make_knots(ht, 400, 21, seed=5, device="cuda:0", model="hopper")     # 8400 knots > one pass
    n = q.shape[0]
    cost = pkg.make_cost(q1=[1.0])
    a = torch.zeros((n, h.model.nd), dtype=torch.float64, device="cuda:0"); b = torch.zeros_like(a); c = torch.zeros_like(a)
    sa = torch.zeros(n, dtype=torch.int32, device="cuda:0")
    h.fd_batch_dev(q, v, u, w, a, None, sa, cost=cost)
    h.fd_batch_dev(q, v, u, w, b, cost=cost)
    assert int((sa != 0).sum()) == 0 and torch.equal(a, b)
    lo, hi = 8100, 8400                                      # straddles the 8192 boundary when part of the full batch
    part = torch.zeros((hi - lo, h.model.nd), dtype=torch.float64, device="cuda:0")
    h.fd_batch_dev(q[lo:hi], v[lo:hi], u[lo:hi], w[lo:hi], part, cost=cost)
    assert torch.equal(part, a[lo:hi])
    ht.fd_batch_dev(q, v, u, w, c, cost=cost)
    assert_deriv_close(a.cpu().numpy(), c.cpu().numpy(), 6, 3)


def test_humanoid_host_pipeline_chunks_do_not_share_scratch(pkg, oracle, omodels):
    """ilqg_fd_batch_host pipelines chunks of knots over streams.  The warp-cooperative engine keeps the centre's C-state,
    candidate lists and row bounds in engine-owned scratch indexed by the knot's position in its chunk, so its chunks must not
    overlap on two compute streams (round-1 advisor finding): with 4 chunks the result must equal one device call bit for bit."""
    import torch
    os.environ["ILQG_HOST_CHUNKS"] = "4"
    try:
        h = pkg.Handle(pkg.Model.named("humanoid"), 0)
    finally:
        del os.environ["ILQG_HOST_CHUNKS"]
    m = h.model
    n = 96
    q, v, u, w = scenario_states("humanoid", n, seed=41, oracle=oracle, om=omodels["humanoid"], roll=60)   # in ground contact
    cost = pkg.make_cost(q1=[1.0])
    for _ in range(3):   # a race would not necessarily show in one pass
        d_host, a_host, status = h.fd_batch_host(q, v, u, w, cost)
        assert status.sum() == 0
        dq, dv, du, dw = (torch.from_numpy(x).cuda() for x in (q, v, u, w))
        d_dev = torch.zeros((n, m.nd), dtype=torch.float64, device="cuda")
        h.fd_batch_dev(dq, dv, du, dw, d_dev, cost=cost)
        assert np.array_equal(d_host, d_dev.cpu().numpy())
    h.close()


def test_ball_joint_model_on_the_generic_engine(pkg, oracle, tmp_path):
    """A tree with a ball joint (/root/reference/src/mjderivative.cpp:152-156 perturbs such joints in the tangent space; none of
    the reference's three models has one) runs on the warp-cooperative engine: forward, step and FD against the oracle, in flight
    and with the hand on the floor."""
    from test_oracle_contact_anchors import ball_model, ball_states
    pm, path = ball_model(pkg, tmp_path)
    om = oracle.Model(path)
    h = pkg.Handle(pm, 0)
    assert h.engine == "generic-warp-per-rollout"
    n = 24
    q, v, u, w = ball_states(n, 11)
    q[n // 2:, :4] = [0.9239, 0, 0.3827, 0]           # pitched down by 45 degrees: the forearm and hand reach the floor
    q[n // 2:, :4] += np.random.default_rng(2).normal(0, 0.05, (n - n // 2, 4))
    q[:, :4] /= np.linalg.norm(q[:, :4], axis=1, keepdims=True)
    q, v, w, _ = oracle.step_batch(om, q, v, u, w, 40)
    info = [oracle.dump(om, q[k], v[k], u[k], w[k])["ncon"] for k in range(n)]
    assert max(info) >= 1 and min(info) == 0          # both regimes are in the sample
    a_ref, _ = oracle.forward_batch(om, q, v, u, w)
    a_gpu, _ = h.forward_batch_host(q, v, u, w)
    assert np.allclose(a_gpu, a_ref, rtol=1e-7, atol=1e-7)
    q1, v1, _, _ = oracle.step_batch(om, q, v, u, w, 10)
    q2, v2, _, _ = h.step_batch_host(q, v, u, w, nsteps=10)
    assert np.allclose(q2, q1, rtol=1e-8, atol=1e-8) and np.allclose(v2, v1, rtol=1e-6, atol=1e-6)
    assert np.allclose(np.linalg.norm(q2[:, :4], axis=1), 1.0, atol=1e-12)
    cost = pkg.make_cost(q2=[0, 1, 1, 0, 0.5], v2=[0.1] * 4, u2=[0.3])
    d_ref, qa_ref, _ = oracle.fd_batch(om, q, v, u, w, cost)
    d_gpu, qa_gpu, status = h.fd_batch_host(q, v, u, w, cost)
    assert status.sum() == 0
    assert np.allclose(qa_gpu, qa_ref, rtol=1e-8, atol=1e-8)
    assert_deriv_close(d_gpu, d_ref, pm.nv, pm.nu, tol=1e-6)
    assert np.abs(d_ref[:, :3 * 4]).max() > 1e-2      # the ball joint's tangent columns carry real derivatives
    h.close()
