"""GPU parity of the FD linearisation (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerance (stated): per deriv block, max-abs difference <= 1e-6 * max(1, ||block||_inf).
The central difference divides by 2e-6, so this bounds the disagreement of the underlying accelerations
at ~1e-12 relative — round-off of two independent fp64 implementations (FMA contraction on the GPU, none
in the oracle).  Knots whose contact active set flips inside the +/-1e-6 stencil are the only legitimate
outliers; none occur in the seeded sets below."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, scenario_states

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 1e-6


def blocks(nv, nu):
    return {"dqacc/dqpos": slice(0, nv * nv), "dqacc/dqvel": slice(nv * nv, 2 * nv * nv), "dqacc/dctrl": slice(2 * nv * nv, 2 * nv * nv + nv * nu),
            "dcost": slice(2 * nv * nv + nv * nu, 2 * nv * nv + nv * nu + 2 * nv + nu)}


def assert_deriv_close(d_gpu, d_ref, nv, nu, tol=TOL):
    for name, sl in blocks(nv, nu).items():
        a, b = d_gpu[:, sl], d_ref[:, sl]
        scale = np.maximum(1.0, np.abs(b).max(axis=1))
        err = np.abs(a - b).max(axis=1) / scale
        assert err.max() <= tol, (name, err.max(), int(err.argmax()))


@pytest.mark.parametrize("name", ["inverted_pendulum", "hopper"])
def test_golden_vectors(handles, name):
    h = handles[name]; m = h.model
    g = np.load(os.path.join(GOLD, f"fd_{name}.npz"))
    deriv, qacc, status = h.fd_batch_host(g["qpos"], g["qvel"], g["ctrl"], g["warm"], g["cost"])
    assert status.sum() == 0
    assert_deriv_close(deriv, g["deriv"], m.nv, m.nu)
    assert np.allclose(qacc, g["qacc"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name,n,roll", [("inverted_pendulum", 257, 0), ("inverted_pendulum", 64, 25), ("hopper", 200, 0),
                                         ("hopper", 96, 150), ("hopper", 33, 400)])
def test_parity_with_oracle_on_seeded_states(handles, oracle, omodels, name, n, roll):
    h = handles[name]; m = h.model; om = omodels[name]
    q, v, u, w = scenario_states(name, n, seed=100 + roll, oracle=oracle, om=om, roll=roll)
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1], q1=[0.5])
    d_ref, a_ref, _ = oracle.fd_batch(om, q, v, u, w, cost)
    d_gpu, a_gpu, status = h.fd_batch_host(q, v, u, w, cost)
    assert status.sum() == 0
    assert_deriv_close(d_gpu, d_ref, m.nv, m.nu)
    assert np.allclose(a_gpu, a_ref, rtol=1e-9, atol=1e-9)


def test_reference_test_scenario(handles, oracle, omodels):
    """/root/reference/tst/test_derivatives.cpp:38-56: hopper, 500 passive steps, ctrl -= 0.1, linearise, cost = qpos[0]."""
    h = handles["hopper"]; om = omodels["hopper"]
    q = np.array([[0, 1.25, 0, 0, 0, 0.0]]); z = np.zeros((1, 6)); u = np.zeros((1, 3))
    q, v, w, _ = h.step_batch_host(q, z, u, z.copy(), nsteps=500)
    qo, vo, wo, _ = oracle.step_batch(om, np.array([[0, 1.25, 0, 0, 0, 0.0]]), z, u, z.copy(), 500)
    assert np.allclose(q, qo, atol=1e-7) and np.allclose(v, vo, atol=1e-6)  # 500 contact-rich steps, two implementations
    u = u - 0.1
    cost = oracle.make_cost(q1=[1.0])
    d_gpu, _, status = h.fd_batch_host(qo, vo, u, wo, cost)
    d_ref, _, _ = oracle.fd_batch(om, qo, vo, u, wo, cost)
    assert status[0] == 0
    assert_deriv_close(d_gpu, d_ref, 6, 3)
    assert d_gpu[0, -15] == pytest.approx(1.0, abs=1e-9)  # dg/dqpos[0] of cost = qpos[0]


def test_joint_limits_active(handles, oracle, omodels):
    h = handles["inverted_pendulum"]; om = omodels["inverted_pendulum"]
    q = np.array([[1.01, 0.2], [-1.02, -0.4], [0.3, 1.58], [0.999999, -1.5707], [1.0, 0.0]])
    v = np.array([[0.3, 0.1], [-0.2, 0.3], [0.0, 0.5], [0.1, -0.1], [0.0, 0.0]]); u = np.zeros((5, 1)) + 0.2
    w = np.zeros((5, 2))
    d_ref, _, _ = oracle.fd_batch(om, q, v, u, w, None)
    d_gpu, _, status = h.fd_batch_host(q, v, u, w, None)
    assert status.sum() == 0
    # rows 3 and 4 sit on a limit boundary: the active set flips inside the stencil, so only the smooth knots are compared tightly
    assert_deriv_close(d_gpu[:3], d_ref[:3], 2, 1)
    assert np.isfinite(d_gpu).all()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 95, 97])
def test_ragged_batch_sizes(handles, oracle, omodels, n):
    """Empty and ragged batches (3 pendulum knots share a warp; hopper uses one warp per knot)."""
    for name in ("inverted_pendulum", "hopper"):
        h = handles[name]; m = h.model; om = omodels[name]
        q, v, u, w = scenario_states(name, max(n, 1), seed=n)
        q, v, u, w = q[:n], v[:n], u[:n], w[:n]
        d_gpu, a_gpu, status = h.fd_batch_host(q, v, u, w, None)
        assert d_gpu.shape == (n, m.nd)
        if n:
            d_ref, _, _ = oracle.fd_batch(om, q, v, u, w, None)
            assert_deriv_close(d_gpu, d_ref, m.nv, m.nu)


def test_null_cost_leaves_gradient_entries_untouched(handles):
    h = handles["hopper"]; m = h.model
    q, v, u, w = scenario_states("hopper", 5, seed=1)
    deriv = np.full((5, m.nd), 123.25)
    d, _, _ = h.fd_batch_host(q, v, u, w, None, deriv=deriv)
    njac = m.nv * (2 * m.nv + m.nu)
    assert (d[:, njac:] == 123.25).all() and not (d[:, :njac] == 123.25).any()


def test_cost_gradient_is_bit_exact_vs_host_arithmetic(handles, oracle, omodels):
    """The forward-difference cost rows are computed with the caller's arithmetic (no FMA contraction)."""
    h = handles["inverted_pendulum"]; om = omodels["inverted_pendulum"]
    q, v, u, w = scenario_states("inverted_pendulum", 64, seed=9)
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1])
    d_ref, _, _ = oracle.fd_batch(om, q, v, u, w, cost)
    d_gpu, _, _ = h.fd_batch_host(q, v, u, w, cost)
    assert np.array_equal(d_gpu[:, 10:], d_ref[:, 10:])


def test_nonfinite_input_is_flagged_not_fatal(handles, pkg):
    h = handles["hopper"]
    q, v, u, w = scenario_states("hopper", 4, seed=2)
    q[2, 1] = np.nan
    d, _, status = h.fd_batch_host(q, v, u, w, None)
    assert status[2] == pkg.ERR_NONFINITE and status[0] == 0 and status[1] == 0 and status[3] == 0
    assert np.isfinite(d[[0, 1, 3]]).all()


def test_bad_arguments(handles, pkg):
    h = handles["hopper"]
    L = pkg.lib()
    assert L.ilqg_fd_batch_host(h._h, 4, None, None, None, None, None, None, None, None, None) == pkg.ERR_ARG
    opts = pkg.FdOpts(eps=-1.0, niter=30, nwarmup=3)
    q, v, u, w = scenario_states("hopper", 2, seed=2)
    with pytest.raises(pkg.IlqgError):
        h.fd_batch_host(q, v, u, w, None, opts=opts)


def test_full_size_properties(handles, pkg):
    """BASELINE configs[1] at full size (4096 x 21 knots): size-independent properties instead of the oracle.
    (1) all knots finite; (2) device-pointer and host-pointer paths agree bit for bit; (3) idempotence;
    (4) the control Jacobian of flight knots equals M^-1 * gear (linearity in ctrl): checked via
        qacc(u + du) - qacc(u) = B du through the forward kernel on a subsample."""
    import torch
    from ilqg_mujoco_b200 import workload as wl
    h = handles["hopper"]; m = h.model
    q, v, u, w, nbad = wl.make_knots(h, 4096, 21, seed=0, device="cuda:0")
    nk = q.shape[0]
    assert nk == 4096 * 21 and nbad == 0
    deriv = torch.zeros((nk, m.nd), dtype=torch.float64, device="cuda:0")
    qacc = torch.zeros((nk, m.nv), dtype=torch.float64, device="cuda:0")
    status = torch.zeros(nk, dtype=torch.int32, device="cuda:0")
    h.fd_batch_dev(q, v, u, w, deriv, qacc, status, cost=None)
    torch.cuda.synchronize()
    assert int(status.sum()) == 0 and bool(torch.isfinite(deriv[:, :90]).all())
    d2 = torch.zeros_like(deriv)
    h.fd_batch_dev(q, v, u, w, d2, qacc, status, cost=None)
    torch.cuda.synchronize()
    assert torch.equal(deriv, d2)                       # idempotent / deterministic
    dh, _, st = h.fd_batch_host(q.cpu().numpy(), v.cpu().numpy(), u.cpu().numpy(), w.cpu().numpy(), None)
    assert np.array_equal(dh[:, :90], deriv[:, :90].cpu().numpy())
    # linearity in ctrl inside the clamp range: finite control step predicted by the FD block
    idx = torch.arange(0, nk, 97, device="cuda:0")
    qs, vs, us, ws = q[idx].contiguous(), v[idx].contiguous(), (0.5 * u[idx]).contiguous(), w[idx].contiguous()
    n = qs.shape[0]
    dd = torch.zeros((n, m.nd), dtype=torch.float64, device="cuda:0"); a0 = torch.zeros((n, 6), dtype=torch.float64, device="cuda:0")
    h.fd_batch_dev(qs, vs, us, ws, dd, a0, None, cost=None)
    du = torch.full((n, 3), 1e-3, dtype=torch.float64, device="cuda:0")
    a1 = torch.zeros_like(a0); wtmp = a0.clone()
    h.forward_batch_dev(qs, vs, (us + du).contiguous(), wtmp, a1)
    torch.cuda.synchronize()
    B = dd[:, 72:90].reshape(n, 6, 3)                   # d qacc_j / d ctrl_i at i + j*nu
    pred = torch.einsum("nji,ni->nj", B, du)
    err = (a1 - a0 - pred).abs().max(dim=1).values / (1 + (a1 - a0).abs().max(dim=1).values)
    # contact knots are piecewise linear: allow the few whose active set changes under the 1e-3 step
    assert float(err.median()) < 1e-6 and float((err < 1e-4).double().mean()) > 0.9


@pytest.mark.parametrize("variant", ["1", "1:warp solves the centre", "2", "3", "4:4 lanes per solve", "4:2 lanes per solve"])
@pytest.mark.parametrize("name,n,roll", [("inverted_pendulum", 300, 10), ("hopper", 271, 150), ("hopper", 85, 0), ("hopper", 1, 200)])
def test_both_fd_kernel_variants(pkg, oracle, omodels, variant, name, n, roll):
    """The per-call choice between the one-launch kernel (variant 1: centre evaluation on a spare lane of its knot's warp, its solves
    done by that lane alone or, ILQG_FD_COOP=1, by the whole warp — solve_coop), the column kernel behind the centre kernel (variant
    2: one thread per perturbed evaluation) and the stage-skipping split (variant 3: fd_velctrl_kernel + fd_qpos_kernel) depends on
    the batch size; the round-2 experiment with several lanes per solve (variant 4: fd_build_kernel + fd_solve_kernel, gsolve.cuh) is
    off by default.  All must meet the same tolerance on the same inputs, including ragged tails (n not a multiple of the knots a
    CTA owns: 85 / 21 hopper); the pendulum has no variant 4 and falls back to 2."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    os.environ["ILQG_FD_VARIANT"] = variant[0]
    if variant.startswith("1:"):
        os.environ["ILQG_FD_COOP"] = "1"
    if variant.startswith("4:"):
        os.environ["ILQG_FD_GW"] = "204" if "4 lanes" in variant else "802"
    try:
        h = pkg.Handle(pkg.Model.named(name), 0)
    finally:
        for k in ("ILQG_FD_VARIANT", "ILQG_FD_COOP", "ILQG_FD_GW"):
            os.environ.pop(k, None)
    m = h.model; om = omodels[name]
    q, v, u, w = scenario_states(name, n, seed=900 + roll, oracle=oracle, om=om, roll=roll)
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1], q1=[0.5])
    d_ref, a_ref, _ = oracle.fd_batch(om, q, v, u, w, cost)
    d_gpu, a_gpu, status = h.fd_batch_host(q, v, u, w, cost)
    assert status.sum() == 0
    assert_deriv_close(d_gpu, d_ref, m.nv, m.nu)
    # without a device cost the caller's gradient entries must survive (host wrappers fill them)
    pre = np.full((n, m.nd), 7.0)
    d2, _, _ = h.fd_batch_host(q, v, u, w, None, deriv=pre.copy())
    nj = m.nv * (2 * m.nv + m.nu)
    assert np.array_equal(d2[:, nj:], pre[:, nj:]) and np.array_equal(d2[:, :nj], d_gpu[:, :nj])
    h.close()


def test_empty_and_single_knot_batches(handles, pkg):
    """nknots = 0 is a no-op that returns OK on every entry point (host and device flavours); nknots = 1 works."""
    import torch
    L = pkg.lib()
    for name in ("inverted_pendulum", "hopper"):
        h = handles[name]; m = h.model
        assert L.ilqg_fd_batch_host(h._h, 0, None, None, None, None, None, None, None, None, None) == pkg.OK
        assert L.ilqg_fd_batch_dev(h._h, 0, None, None, None, None, None, None, None, None, None, None) == pkg.OK
        assert L.ilqg_step_batch_host(h._h, 0, 5, None, None, None, None, None) == pkg.OK
        assert L.ilqg_forward_batch_dev(h._h, 0, None, None, None, None, None, None) == pkg.OK
        assert L.ilqg_fd_batch_host(h._h, -1, None, None, None, None, None, None, None, None, None) == pkg.ERR_ARG
        q = np.zeros((1, m.nq)); q[0, min(1, m.nq - 1)] = 1.2 if name == "hopper" else 0.1
        d, a, st = h.fd_batch_host(q, np.zeros((1, m.nv)), np.zeros((1, m.nu)), None, None)
        assert d.shape == (1, m.nd) and st[0] == 0 and np.isfinite(d[:, :m.nv * (2 * m.nv + m.nu)]).all()


def test_out_of_plane_variant_of_a_planar_tree_leaves_the_planar_kernels(pkg, oracle, tmp_path):
    """The compiled-in hopper topology carries PLANAR_Y (its kernels drop every out-of-plane component at compile time).
    The same tree with one geom moved out of the xz-plane must NOT bind to those kernels: it runs on the generic engine
    and still matches the oracle, which computes the dense algebra on the modified tables."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    m = pkg.Model.named("hopper").copy()
    gp = m.field("geom_pos").reshape(-1, 3)
    gp[4, 1] = 0.03                       # the foot capsule, 3 cm off the plane
    path = str(tmp_path / "hopper_offplane.ilqgm")
    assert pkg.lib().ilqg_model_save(path.encode(), m.ptr) == 0
    om = oracle.Model(path)
    h = pkg.Handle(m, 0)
    assert h.engine == "generic-warp-per-rollout"
    q, v, u, w = scenario_states("hopper", 48, seed=4242, oracle=oracle, om=om, roll=120)
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1])
    d_ref, _, _ = oracle.fd_batch(om, q, v, u, w, cost)
    d_gpu, _, status = h.fd_batch_host(q, v, u, w, cost)
    assert status.sum() == 0
    assert_deriv_close(d_gpu, d_ref, m.nv, m.nu)
    h.close()
    # and the untouched tables do bind to the planar instantiation
    h2 = pkg.Handle(pkg.Model.named("hopper"), 0)
    assert h2.engine == "hopper"
    h2.close()


@pytest.mark.parametrize("name", ["hopper", "humanoid"])
def test_centre_diagnostics_report_solver_iterations(pkg, handles, oracle, omodels, name):
    """ilqg_fd_set_diag: rows and Newton iterations of the centre evaluation (the reference pins 30 iterations / tolerance 0,
    mjderivative.cpp:241-242; the kernels' solver leaves at the exact minimiser — the caller can see how many ran)."""
    import torch
    h = handles[name] if name in handles else pkg.Handle(pkg.Model.named(name), 0)
    m = h.model; om = omodels[name]
    n = 64 if name == "hopper" else 12
    q, v, u, w = scenario_states(name, n, seed=17, oracle=oracle, om=om, roll=150 if name == "hopper" else 60)
    dq, dv, du, dw = (torch.from_numpy(x).cuda() for x in (q, v, u, w))
    deriv = torch.zeros((n, m.nd), dtype=torch.float64, device="cuda")
    diag = torch.full((n, 8), -1, dtype=torch.int32, device="cuda")
    h.fd_set_diag(diag)
    h.fd_batch_dev(dq, dv, du, dw, deriv)
    h.fd_set_diag(None)
    d = diag.cpu().numpy()
    info = np.zeros(4, np.int32); qa = np.zeros(om.nv)
    for k in range(n):
        oracle.lib().mjo_debug_forward(om.ptr, oracle._p(q[k]), oracle._p(v[k]), oracle._p(u[k]), oracle._p(w[k]), 30, C.c_double(0.0), oracle._p(qa),
                                       info.ctypes.data_as(C.c_void_p), None, None, None)
        assert d[k, 0] == info[1] and d[k, 6] == info[0]   # nefc, ncon  (info = ncon, nefc, iterations, active rows)
    assert (d[:, 0] > 0).any() and ((d[:, 1] >= 0) & (d[:, 1] <= 30)).all() and (d[:, 2] >= d[:, 1]).all()
    assert (d[d[:, 0] == 0, 2] == 0).all()             # no rows: no iterations
    assert (d[d[:, 0] > 0, 2] >= 1).all() and (d[:, 3] <= d[:, 0]).all()
    assert (d[:, 4] > 0).all() and (d[:, 5] >= 0).all()
    # switched off: the array is left alone
    diag.fill_(-7)
    h.fd_batch_dev(dq, dv, du, dw, deriv)
    torch.cuda.synchronize()
    assert int((diag != -7).sum()) == 0
    if name not in handles:
        h.close()


def test_host_pinning_of_caller_buffers_is_transparent(pkg, oracle, omodels):
    """ilqg_set_host_pinning: the host-pointer call page-locks the caller's (pageable) arrays on first use — same bits out, the
    registration is reused by the next call, released when switched off or when the handle is destroyed."""
    h = pkg.Handle(pkg.Model.named("hopper"), 0)
    n = 2100   # deriv = 1.76 MB: above the 1 MB threshold of the registration
    q, v, u, w = scenario_states("hopper", n, seed=23)
    cost = pkg.make_cost(q1=[1.0])
    d0, a0, s0 = h.fd_batch_host(q, v, u, w, cost)
    assert pkg.lib().ilqg_set_host_pinning(h._h, 1) == 0
    buf = np.zeros((n, h.model.nd))
    for _ in range(3):
        d1, a1, s1 = h.fd_batch_host(q, v, u, w, cost, deriv=buf)
        assert d1 is buf and np.array_equal(d1, d0) and np.array_equal(a1, a0)
    assert pkg.lib().ilqg_set_host_pinning(h._h, 0) == 0
    d2, _, _ = h.fd_batch_host(q, v, u, w, cost, deriv=buf)
    assert np.array_equal(d2, d0)
    pkg.lib().ilqg_set_host_pinning(h._h, 1)
    h.fd_batch_host(q, v, u, w, cost, deriv=buf)
    h.close()   # unregisters
    buf[:] = 1.0  # the array is ordinary memory again


@pytest.mark.parametrize("n", [0, 1, 21, 31, 33, 64, 65])
def test_tiny_and_empty_batches(pkg, handles, oracle, omodels, n):
    """Edge sizes of the small-batch kernels (one launch up to 64 knots, two overlapped launches above): empty, a single knot, one
    trajectory, warp-boundary counts — against the oracle, with the centre accelerations and status words."""
    h = handles["hopper"]; om = omodels["hopper"]
    q, v, u, w = scenario_states("hopper", max(n, 1), seed=77, oracle=oracle, om=om, roll=130)
    q, v, u, w = q[:n], v[:n], u[:n], w[:n]
    cost = oracle.make_cost(q2=[1, 10], v2=[1, 10], u2=[1], q1=[0.5])
    d_gpu, a_gpu, status = h.fd_batch_host(q, v, u, w, cost)
    assert d_gpu.shape == (n, 105) and status.shape == (n,)
    if n == 0:
        return
    d_ref, a_ref, _ = oracle.fd_batch(om, q, v, u, w, cost)
    assert status.sum() == 0
    assert_deriv_close(d_gpu, d_ref, 6, 3)
    assert np.allclose(a_gpu, a_ref, rtol=1e-9, atol=1e-9)
