"""The constraint solve of the oracle (mj_fwdConstraint with the solver pinned as /root/reference/src/mjderivative.cpp:241-246 pins it:
30 iterations, tolerance 0) at random folded, moving poses of all three models, judged by what the SOLUTION must satisfy rather than
by how it was reached.  The problem is the strictly convex

    min_a  1/2 (a - a_s)' M (a - a_s) + sum_r 1/2 D_r min(0, J_r a - aref_r)^2                    

so its minimiser is unique and characterised by stationarity, M (a - a_s) = J' f with f_r = -D_r min(0, J_r a - aref_r) >= 0.  Checked from
the dumped M, a_s, J, D, aref alone (numpy): the stationarity residual, the forces, that no random neighbour has a lower cost, and the
minimiser itself against an independent solve — Newton on the active set written out here in dense numpy, started from a_s instead of
the warm start.  tests/test_oracle_anchors.py checks the KKT residual at one resting hopper state; this covers limits, ground contacts
and self-collisions of hopper and humanoid and the pendulum's limits."""
import numpy as np
import pytest


def cost(a, M, a_s, J, D, aref):
    jar = J @ a - aref
    return 0.5 * (a - a_s) @ M @ (a - a_s) + 0.5 * float(D @ np.minimum(jar, 0.0) ** 2)


def active_set_newton(M, a_s, J, D, aref, iters=100):
    """Newton with an exact line search on the piecewise-quadratic cost: dense numpy, started at a_s"""
    a = a_s.copy()
    for _ in range(iters):
        jar = J @ a - aref
        act = jar < 0
        g = M @ (a - a_s) + J[act].T @ (D[act] * jar[act])
        H = M + J[act].T @ (D[act][:, None] * J[act])
        p = -np.linalg.solve(H, g)
        if np.abs(p).max() < 1e-14 * max(1.0, np.abs(a).max()):
            return a, True
        # exact minimiser of the convex piecewise-quadratic along p: bisection on the derivative (monotone in t)
        jp = J @ p
        def dphi(t):
            r = jar + t * jp
            return p @ (M @ (a + t * p - a_s)) + float((D * np.minimum(r, 0.0)) @ jp)
        lo, hi = 0.0, 1.0
        while dphi(hi) < 0 and hi < 1e6:
            hi *= 2
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            if dphi(mid) < 0:
                lo = mid
            else:
                hi = mid
        a = a + 0.5 * (lo + hi) * p
    return a, False


def random_state(name, m, rng):
    q = m.field("qpos0")[:m.nq].copy()
    rngs = m.field("jnt_range").reshape(-1, 2)[:m.njnt]
    if name == "humanoid":
        q[2] = rng.uniform(0.2, 0.8)
        w = rng.normal(0, 1.0, 3); ang = np.linalg.norm(w)
        q[3:7] = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])
        q[7:] = rng.uniform(rngs[1:, 0] - 0.01, rngs[1:, 1] + 0.01)
    elif name == "hopper":
        q[1] = rng.uniform(0.3, 1.2); q[2] = rng.uniform(-1.5, 1.5)
        q[3:] = rng.uniform(rngs[3:, 0] - 0.01, rngs[3:, 1] + 0.01)
    else:
        q[:] = [rng.choice([-1.0, 1.0]) * rng.uniform(0.9, 1.02), rng.choice([-1.0, 1.0]) * rng.uniform(1.4, 1.6)]   # near / beyond the limits
    return q, rng.normal(0, 1.0, m.nv), rng.uniform(-1, 1, m.nu)


@pytest.mark.parametrize("name", ["inverted_pendulum", "hopper", "humanoid"])
def test_constraint_solve_returns_the_unique_minimiser(oracle, omodels, pkg, name):
    om = omodels[name]
    m = pkg.Model.named(name)
    rng = np.random.default_rng(55)
    with_rows = worst_stat = worst_alt = 0
    for trial in range(60):
        q, v, u = random_state(name, m, rng)
        d = oracle.dump(om, q, v, u, iterations=30, tolerance=0.0)
        if d["nefc"] == 0:
            assert np.allclose(d["qacc"], d["qacc_smooth"], rtol=0, atol=1e-12)
            continue
        with_rows += 1
        M, a_s, a, J, D, aref = d["qM"], d["qacc_smooth"], d["qacc"], d["efc_J"], d["efc_D"], d["efc_aref"]
        jar = J @ a - aref
        f = -D * np.minimum(jar, 0.0)
        assert (f >= 0).all()
        assert np.allclose(d["efc_force"], f, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(f).max()))
        scale = max(1.0, np.abs(M @ a_s).max(), np.abs(J.T @ f).max())
        stat = np.abs(M @ (a - a_s) - J.T @ f).max() / scale
        worst_stat = max(worst_stat, stat)
        assert stat < 1e-9, (trial, stat)
        c0 = cost(a, M, a_s, J, D, aref)
        for _ in range(20):                                          # no neighbour is better
            da = rng.normal(0, 1.0, m.nv) * 10.0 ** rng.uniform(-6, 0) * max(1.0, np.abs(a).max())
            assert cost(a + da, M, a_s, J, D, aref) >= c0 - 1e-9 * max(1.0, abs(c0))
        alt, ok = active_set_newton(M, a_s, J, D, aref)
        assert ok, trial
        err = np.abs(alt - a).max() / max(1.0, np.abs(a).max())
        worst_alt = max(worst_alt, err)
        assert err < 1e-8, (trial, err)
    assert with_rows >= (10 if name == "inverted_pendulum" else 20)
