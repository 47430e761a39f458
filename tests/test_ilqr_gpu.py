"""Batched GPU iLQR (rollouts for all alphas, ladder-order acceptance, FD, Riccati) against the oracle's restatement of
ILQR<nv,nu,N> — which itself is checked against the reference's own ilqr.h compiled verbatim (test_oracle_ilqr.py).

Stated tolerances: cost traces 1e-8 relative and accepted alphas identical on the pendulum (N=20, 10 iterations);
gains K/k and the value model V/v 1e-6 relative (they pass through mu = 1000 conditioning and 1e-6 finite differences)."""
import numpy as np
import pytest

from conftest import scenario_states

pytestmark = pytest.mark.gpu
PEND_COST = dict(q2=[1, 10], v2=[1, 10], u2=[1])  # /root/reference/inc/inverted_pendulum/cost.h:7-17


def rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


def run_gpu(pkg, h, q, v, u, w, cost, N, niter, alphas, accept_always):
    il = pkg.Ilqr(h, q.shape[0], N, alphas)
    il.set_cost(cost)
    il.init_host(q, v, u, w)
    il.iterate(niter, accept_always)
    out = il.get()
    il.close()
    return out


def test_pendulum_reference_mode_matches_oracle(pkg, handles, oracle, omodels):
    h = handles["inverted_pendulum"]; om = omodels["inverted_pendulum"]
    q, v, u, w = scenario_states("inverted_pendulum", 64, seed=40)
    u *= 0.2
    cost = oracle.make_cost(**PEND_COST)
    ref = oracle.ilqr_run_batch(om, 20, 10, q, v, u, w, cost, alphas=None)
    out = run_gpu(pkg, h, q, v, u, w, cost, 20, 10, (1.0,), True)
    assert rel(out["J"], ref["J"]) < 1e-8
    assert np.allclose(out["qpos"], ref["qpos"], rtol=1e-7, atol=1e-8) and np.allclose(out["ctrl"], ref["ctrl"], rtol=1e-6, atol=1e-7)
    for key in ("K", "k", "V", "v"):
        assert rel(out[key][:, 1:] if key in ("K", "k") else out[key], ref[key][:, 1:] if key in ("K", "k") else ref[key]) < 1e-6, key
    assert (out["accepted"] == 0).all()


def test_pendulum_linesearch_accepts_same_alphas_as_sequential_backtracking(pkg, handles, oracle, omodels):
    h = handles["inverted_pendulum"]; om = omodels["inverted_pendulum"]
    q, v, u, w = scenario_states("inverted_pendulum", 96, seed=41)
    u *= 0.2
    cost = oracle.make_cost(**PEND_COST)
    alphas = [1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125]
    ref = oracle.ilqr_run_batch(om, 20, 10, q, v, u, w, cost, alphas=alphas, accept_always=False, mu=10.0)
    il = pkg.Ilqr(h, 96, 20, alphas)
    il.set_cost(cost); il.set_mu(10.0)
    il.init_host(q, v, u, w)
    il.iterate(10, accept_always=False)
    out = il.get(); il.close()
    assert np.array_equal(out["accepted"], ref["accepted"])
    assert len(np.unique(ref["accepted"])) > 1     # the ladder is really exercised (some rejections / partial steps)
    assert rel(out["J"], ref["J"]) < 1e-8
    # monotone: an accepted step never increases the cost
    assert (np.diff(out["J"], axis=1) <= 1e-12 * np.abs(out["J"][:, :-1])).all()


def test_hopper_ilqr_matches_oracle(pkg, handles, oracle, omodels):
    """nx = 12, nu = 3 exercises the warp-per-instance Riccati path and the scrambled B layout (quirk Q1).
    The reference's full-step iLQR is not stable on the hopper (its A/B are mis-assembled for nu = 3):
    the cost grows and diverges by the third iteration in the oracle too, so parity is checked on the first
    iteration's gains / value model and on the first two cost values."""
    h = handles["hopper"]; om = omodels["hopper"]
    q, v, u, w = scenario_states("hopper", 8, seed=42, oracle=oracle, om=om, roll=100)
    cost = oracle.make_cost(q2=[0, 5, 1, 0, 0, 0], q1=[-1.0], v2=[0.1] * 6, u2=[0.01] * 3)
    ref = oracle.ilqr_run_batch(om, 20, 1, q, v, u, w, cost, alphas=None)
    out = run_gpu(pkg, h, q, v, u, w, cost, 20, 1, (1.0,), True)
    assert rel(out["J"], ref["J"]) < 1e-9                      # open-loop pass
    assert np.allclose(out["qpos"], ref["qpos"], rtol=1e-8, atol=1e-9)
    for key in ("K", "k", "V", "v"):
        a_, b_ = (out[key][:, 1:], ref[key][:, 1:]) if key in ("K", "k") else (out[key], ref[key])
        assert rel(a_, b_) < 1e-5, (key, rel(a_, b_))
    ref2 = oracle.ilqr_run_batch(om, 20, 2, q, v, u, w, cost, alphas=None)
    out2 = run_gpu(pkg, h, q, v, u, w, cost, 20, 2, (1.0,), True)
    assert rel(out2["J"], ref2["J"]) < 1e-4                    # closed-loop pass through contacts with the gains above


def test_mpc_state_update(pkg, handles, oracle, omodels):
    """setDInit with a NEW state between iterations (the MPC cadence of inverted_pendulum.cpp:19-30)."""
    h = handles["inverted_pendulum"]
    q, v, u, w = scenario_states("inverted_pendulum", 4, seed=43)
    cost = oracle.make_cost(**PEND_COST)
    il = pkg.Ilqr(h, 4, 20, (1.0,))
    il.set_cost(cost)
    il.init_host(q, v, u * 0, w)
    il.iterate(10, True)
    a = il.get()
    q2, v2, w2, _ = h.step_batch_host(q, v, a["ctrl"][:, 20], w, nsteps=1)
    il.set_state_host(q2, v2, w2)
    il.iterate(1, True)
    b = il.get()
    il.close()
    assert np.allclose(b["qpos"][:, 20], q2) and np.allclose(b["qvel"][:, 20], v2)   # the pass started from the new state
    assert b["J"].shape == (4, 11)


def test_unsupported_and_bad_arguments(pkg, handles):
    import ctypes as C
    h = handles["hopper"]
    w = C.c_void_p()
    assert pkg.lib().ilqg_ilqr_create(h._h, 0, 20, 1, None, C.byref(w)) == pkg.ERR_ARG
    il = pkg.Ilqr(h, 2, 5, (1.0,))
    with pytest.raises(pkg.IlqgError):
        il.iterate(1)          # cost not set
    il.close()


def test_corrected_layout_mode_matches_oracle_and_differs_from_the_quirk(pkg, handles, oracle, omodels):
    """Opt-in: A/B assembled from deriv as d qacc_j / d x_i instead of through the reference's
    column-major views (quirk Q1).  Not the parity mode — but the oracle carries the same switch, so the GPU is still
    checked against a CPU statement of it; on the hopper (nu = 3 != nv) the two layouts give different gains."""
    h = handles["hopper"]; om = omodels["hopper"]
    q, v, u, w = scenario_states("hopper", 6, seed=43, oracle=oracle, om=om, roll=100)
    cost = oracle.make_cost(q2=[0, 5, 1, 0, 0, 0], q1=[-1.0], v2=[0.1] * 6, u2=[0.01] * 3)
    ref = oracle.ilqr_run_batch(om, 20, 1, q, v, u, w, cost, alphas=None, corrected=True)
    quirk = oracle.ilqr_run_batch(om, 20, 1, q, v, u, w, cost, alphas=None)
    il = pkg.Ilqr(h, 6, 20, (1.0,))
    il.set_cost(cost)
    il.set_layout(True)
    il.init_host(q, v, u, w)
    il.iterate(1, True)
    out = il.get()
    il.close()
    for key in ("K", "k", "V", "v"):
        a_, b_ = (out[key][:, 1:], ref[key][:, 1:]) if key in ("K", "k") else (out[key], ref[key])
        assert rel(a_, b_) < 1e-5, (key, rel(a_, b_))
    assert rel(quirk["K"][:, 1:], ref["K"][:, 1:]) > 1e-3      # the switch changes the matrices
    # pendulum, 10 iterations with the ladder: cost trace parity in corrected mode as well
    hp = handles["inverted_pendulum"]; omp_ = omodels["inverted_pendulum"]
    q, v, u, w = scenario_states("inverted_pendulum", 16, seed=7)
    u = u * 0
    pc = oracle.make_cost(**PEND_COST)
    al = tuple(0.5 ** a for a in range(6))
    ref = oracle.ilqr_run_batch(omp_, 20, 10, q, v, u, w, pc, alphas=al, accept_always=False, corrected=True)
    il = pkg.Ilqr(hp, 16, 20, al)
    il.set_cost(pc)
    il.set_layout(True)
    il.init_host(q, v, u, w)
    il.iterate(10, False)
    out = il.get()
    il.close()
    assert np.array_equal(out["accepted"], ref["accepted"])
    assert rel(out["J"], ref["J"]) < 1e-8


def test_mu_schedule_matches_oracle(pkg, handles, oracle, omodels):
    """Opt-in: per-instance Levenberg-Marquardt term, relaxed after an accepted line-search step and
    stiffened after a rejected ladder.  Same switch in the oracle; traces and accepted alphas must agree, and the schedule
    must change the iterates with respect to the reference's constant mu = 1000."""
    h = handles["inverted_pendulum"]; om = omodels["inverted_pendulum"]
    q, v, u, w = scenario_states("inverted_pendulum", 24, seed=21)
    u = u * 0
    cost = oracle.make_cost(**PEND_COST)
    al = tuple(0.5 ** a for a in range(5))
    sched = (2.0, 1.0, 1e6)
    ref = oracle.ilqr_run_batch(om, 20, 8, q, v, u, w, cost, alphas=al, accept_always=False, mu_schedule=sched)
    const = oracle.ilqr_run_batch(om, 20, 8, q, v, u, w, cost, alphas=al, accept_always=False)
    il = pkg.Ilqr(h, 24, 20, al)
    il.set_cost(cost)
    il.set_mu_schedule(*sched)
    il.init_host(q, v, u, w)
    il.iterate(8, False)
    out = il.get()
    il.close()
    assert np.array_equal(out["accepted"], ref["accepted"])
    assert rel(out["J"], ref["J"]) < 1e-8
    assert rel(ref["J"][:, -1], const["J"][:, -1]) > 1e-6      # the schedule is not a no-op


def test_humanoid_ilqr_in_tangent_coordinates_matches_oracle(pkg, oracle, omodels):
    """the reference's ILQR takes "2 nv doubles at qpos" as its state, which is undefined with the
    humanoid's quaternion (nq = 28, nv = 27).  The opt-in extension keeps the reference's algorithm (ilqr.h:116-186) in TANGENT
    coordinates — x (-) x* through mju_subQuat for the free joint, A/B from the FD blocks, which calcMJDerivatives already takes
    in those coordinates — on the warp-cooperative engine; the oracle carries the same extension (mjo_state_diff)."""
    h = pkg.Handle(pkg.Model.named("humanoid"), 0)
    om = omodels["humanoid"]
    n, N = 3, 6
    q, v, u, w = scenario_states("humanoid", n, seed=5)     # airborne: smooth dynamics, the parity is not at the mercy of contact flips
    u = u * 0
    cost = oracle.make_cost(q2=[0, 0, 2.0, 0, 1, 1, 0], q1=[0, 0, -5.2], v2=[0.05] * 27, u2=[0.02] * 21)   # height, uprightness, effort
    with pytest.raises(pkg.IlqgError):      # the reference's layout is refused for this model
        il = pkg.Ilqr(h, n, N, (1.0,))
        il.set_cost(cost)
        il.init_host(q, v, u, w)
        il.iterate(1, True)
    il.close()
    al = (1.0, 0.5, 0.25, 0.125)
    sched = (2.0, 1.0, 1e8)
    # one iteration: open-loop pass, then gains and value model of the 54-dimensional Riccati sweep
    ref1 = oracle.ilqr_run_batch(om, N, 1, q, v, u, w, cost, alphas=al, accept_always=False, corrected=True, mu_schedule=sched)
    il = pkg.Ilqr(h, n, N, al)
    il.set_cost(cost)
    il.set_layout(True)
    il.set_mu_schedule(*sched)
    il.init_host(q, v, u, w)
    il.iterate(1, False)
    out1 = il.get()
    assert rel(out1["J"], ref1["J"]) < 1e-9
    assert np.allclose(out1["qpos"], ref1["qpos"], rtol=1e-8, atol=1e-9)
    for key in ("K", "k", "V", "v"):
        a_, b_ = (out1[key][:, 1:], ref1[key][:, 1:]) if key in ("K", "k") else (out1[key], ref1[key])
        assert rel(a_, b_) < 1e-5, (key, rel(a_, b_))
    # several iterations with the ladder: cost trace and accepted alphas
    ref = oracle.ilqr_run_batch(om, N, 4, q, v, u, w, cost, alphas=al, accept_always=False, corrected=True, mu_schedule=sched)
    il.init_host(q, v, u, w)
    il.set_mu(1000.0)
    il.iterate(4, False)
    out = il.get()
    il.close()
    h.close()
    assert np.array_equal(out["accepted"][:, -4:], ref["accepted"])
    assert rel(out["J"][:, -4:], ref["J"]) < 1e-6
    assert (ref["accepted"] >= 0).any() and (np.diff(ref["J"], axis=1) <= 1e-12).all()   # the ladder accepts steps and the cost falls
    assert np.allclose(np.linalg.norm(out["qpos"][:, :, 3:7], axis=2), 1.0, atol=1e-9)


@pytest.mark.parametrize("name,n,nsub", [("inverted_pendulum", 100, 3), ("inverted_pendulum", 33, 2), ("hopper", 70, 2)])
def test_sub_batched_workspace_gives_the_same_bits(pkg, handles, oracle, omodels, name, n, nsub, monkeypatch):
    """ilqg_ilqr_iterate runs the iteration chains of blocks of instances on separate streams (ilqg_ilqr_s: sub-batches; by default
    four blocks from 1024 instances on).  Blocks change where an instance's data lives and when its kernels run, not its arithmetic:
    every result — trajectories, gains, value model, cost trace, accepted alphas, first controls, the knots' warm starts — equals the
    single-block workspace bit for bit, through the captured graph as well as launch by launch, with ragged blocks (boundaries at
    multiples of 32 instances), the line-search ladder and the per-instance mu schedule."""
    h = handles[name]
    q, v, u, w = scenario_states(name, n, seed=77, oracle=oracle, om=omodels[name], roll=140 if name == "hopper" else 0)
    u = u * 0.2
    m = h.model
    cost = oracle.make_cost(**PEND_COST) if name == "inverted_pendulum" else oracle.make_cost(q2=[0, 5, 1, 0.1, 0.1, 0.1], v2=[0.1] * 6, u2=[0.01] * 3)
    alphas = [1.0, 0.5, 0.25, 0.125]
    res = []
    for sub in (1, nsub):
        monkeypatch.setenv("ILQG_ILQR_SUB", str(sub))
        il = pkg.Ilqr(h, n, 12, alphas)
        monkeypatch.delenv("ILQG_ILQR_SUB")
        il.set_cost(cost)
        if name == "hopper":
            il.set_layout(1)
        il.set_mu(10.0)
        il.set_mu_schedule(2.0)
        il.init_host(q, v, u, w)
        for _ in range(4):                 # plain launches, graph capture, two replays
            il.iterate(2, accept_always=False)
        out = il.get()
        out["u0"], out["Jt"] = il.fetch_controls()
        kq = np.zeros((n, 13, m.nq)); kv = np.zeros((n, 13, m.nv)); ku = np.zeros((n, 13, m.nu)); kw = np.zeros((n, 13, m.nv))
        from ctypes import c_void_p
        h._check(pkg.lib().ilqg_ilqr_get_knots_host(il._w, kq.ctypes.data_as(c_void_p), kv.ctypes.data_as(c_void_p), ku.ctypes.data_as(c_void_p),
                                                   kw.ctypes.data_as(c_void_p)))
        out.update(kq=kq, kv=kv, ku=ku, kw=kw)
        il.set_state_host(q * 0.9, v)       # a later MPC step on the same workspace
        il.iterate(2, accept_always=False)
        out["u0b"], out["Jtb"] = il.fetch_controls()
        il.close()
        res.append(out)
    for key in res[0]:
        assert np.array_equal(res[0][key], res[1][key]), key
    assert np.array_equal(res[0]["kq"], res[0]["qpos"]) and len(np.unique(res[0]["accepted"])) > 1


@pytest.mark.parametrize("nsub", [1, 3])
def test_one_launch_forward_pass_equals_rollout_accept_commit(pkg, oracle, omodels, nsub, monkeypatch):
    """In the reference's own mode (one alpha, accepted unconditionally) the forward pass is ONE launch: the rollout writes straight over
    the nominal and does the acceptance bookkeeping itself (ilqr_rollout_kernel<T, true>: cost, trace slot from the device counter, mu
    schedule, the counter advanced by the last CTA to finish).  ILQG_ILQR_DIRECT=0 restores the three launches (rollout into the
    candidate buffers, accept, commit).  Same bits in every output, through plain launches and graph replays, for one block and for
    ragged blocks."""
    q, v, u, w = scenario_states("inverted_pendulum", 100, seed=91)
    u = u * 0.2
    cost = oracle.make_cost(**PEND_COST)
    res = []
    for direct in ("1", "0"):
        monkeypatch.setenv("ILQG_ILQR_DIRECT", direct)
        monkeypatch.setenv("ILQG_ILQR_SUB", str(nsub))
        h = pkg.Handle(pkg.Model.named("inverted_pendulum"), 0)
        il = pkg.Ilqr(h, 100, 20, (1.0,))
        monkeypatch.delenv("ILQG_ILQR_DIRECT"); monkeypatch.delenv("ILQG_ILQR_SUB")
        il.set_cost(cost)
        il.set_mu(50.0)
        il.set_mu_schedule(1.5)
        il.init_host(q, v, u, w)
        n_launch = []
        for _ in range(4):
            l0 = h.launches
            il.iterate(3, accept_always=True)
            n_launch.append(h.launches - l0)
        out = il.get()
        out["u0"], out["Jt"] = il.fetch_controls(last=5)
        out["launches"] = np.array(n_launch)
        il.close(); h.close()
        res.append(out)
    for key in res[0]:
        if key != "launches":     # (full steps without a line search: a few random starts diverge, identically on both sides)
            assert np.array_equal(res[0][key], res[1][key], equal_nan=res[0][key].dtype.kind == "f"), key
    assert (res[0]["launches"] < res[1]["launches"]).all()      # two launches fewer per block and iteration
    assert (res[0]["accepted"] == 0).all() and np.isfinite(res[0]["J"]).all(axis=1).mean() > 0.9
