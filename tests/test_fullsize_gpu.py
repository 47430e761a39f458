"""BASELINE.json's full size (configs[1]: 4096 trajectories x 21 hopper knots = 86,016 knots on one GPU) through
size-independent properties, plus an oracle spot check on a strided sample:
  * idempotence: the same call twice is bit-identical (no data races, no dependence on scheduling);
  * batch-split invariance: every knot's block is independent of which other knots share the launch — a slice computed alone
    (same kernel variant) is bit-identical to the slice of the full batch, the other kernel variant agrees to the FD tolerance;
  * permutation: permuting the knots permutes the blocks;
  * the host-pointer entry point (chunked copy/compute pipeline) returns the device-pointer result bit for bit."""
import os

import numpy as np
import pytest

from test_fd_gpu import assert_deriv_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ilqg_mujoco_b200 import workload as wl
    os.environ["ILQG_FD_VARIANT"] = "3"
    try:
        h = pkg.Handle(pkg.Model.named("hopper"), 0)
    finally:
        del os.environ["ILQG_FD_VARIANT"]
    q, v, u, w, _ = wl.make_knots(h, 4096, 21, seed=0, device="cuda:0", model="hopper")
    cost = pkg.make_cost(q1=[1.0])
    n = q.shape[0]
    deriv = torch.zeros((n, h.model.nd), dtype=torch.float64, device="cuda:0")
    status = torch.zeros(n, dtype=torch.int32, device="cuda:0")
    h.fd_batch_dev(q, v, u, w, deriv, None, status, cost=cost)
    torch.cuda.synchronize()
    yield dict(h=h, q=q, v=v, u=u, w=w, cost=cost, deriv=deriv, status=status, n=n)
    h.close()


def test_full_batch_is_finite_and_idempotent(full):
    import torch
    f = full
    assert f["n"] == 86016 and int((f["status"] != 0).sum()) == 0
    assert bool(torch.isfinite(f["deriv"]).all())
    again = torch.zeros_like(f["deriv"])
    f["h"].fd_batch_dev(f["q"], f["v"], f["u"], f["w"], again, cost=f["cost"])
    assert torch.equal(again, f["deriv"])


def test_batch_split_and_permutation_invariance(full, pkg):
    import torch
    f = full; h = f["h"]
    lo, hi = 30000, 30000 + 17011            # a ragged slice, still the split kernels (forced variant 3)
    part = torch.zeros((hi - lo, h.model.nd), dtype=torch.float64, device="cuda:0")
    h.fd_batch_dev(f["q"][lo:hi], f["v"][lo:hi], f["u"][lo:hi], f["w"][lo:hi], part, cost=f["cost"])
    assert torch.equal(part, f["deriv"][lo:hi])
    perm = torch.randperm(f["n"], generator=torch.Generator().manual_seed(3)).to("cuda:0")
    out = torch.zeros_like(f["deriv"])
    h.fd_batch_dev(f["q"][perm].contiguous(), f["v"][perm].contiguous(), f["u"][perm].contiguous(), f["w"][perm].contiguous(), out, cost=f["cost"])
    assert torch.equal(out, f["deriv"][perm])
    # the single-launch kernel (what small batches use) on a slice: same numbers to the FD tolerance
    os.environ["ILQG_FD_VARIANT"] = "2"
    try:
        h2 = pkg.Handle(pkg.Model.named("hopper"), 0)
    finally:
        del os.environ["ILQG_FD_VARIANT"]
    other = torch.zeros((4000, h.model.nd), dtype=torch.float64, device="cuda:0")
    h2.fd_batch_dev(f["q"][50000:54000], f["v"][50000:54000], f["u"][50000:54000], f["w"][50000:54000], other, cost=f["cost"])
    assert_deriv_close(other.cpu().numpy(), f["deriv"][50000:54000].cpu().numpy(), 6, 3)
    h2.close()


def test_host_entry_point_equals_device_path(full):
    f = full; h = f["h"]
    d, a, st = h.fd_batch_host(f["q"].cpu().numpy(), f["v"].cpu().numpy(), f["u"].cpu().numpy(), f["w"].cpu().numpy(), f["cost"])
    assert st.sum() == 0
    assert np.array_equal(d, f["deriv"].cpu().numpy())


def test_strided_sample_matches_oracle(full, oracle, omodels):
    f = full
    idx = np.arange(0, f["n"], 97)           # 887 knots spread over flight and stance trajectories
    import torch
    it = torch.from_numpy(idx).to("cuda:0")
    q, v, u, w = (f[k][it].cpu().numpy() for k in ("q", "v", "u", "w"))
    d_ref, _, _ = oracle.fd_batch(omodels["hopper"], q, v, u, w, f["cost"])
    assert_deriv_close(f["deriv"][it].cpu().numpy(), d_ref, 6, 3)


def test_work_class_ordering_changes_placement_only(full, pkg):
    """The split kernels take their knots through a permutation that groups them by the centre's constraint-row count
    (heaviest first).  Switching the ordering off (identity placement) must give the same bits for every knot."""
    import torch
    f = full
    os.environ["ILQG_FD_VARIANT"] = "3"
    os.environ["ILQG_FD_BINS"] = "0"
    try:
        h0 = pkg.Handle(pkg.Model.named("hopper"), 0)
    finally:
        del os.environ["ILQG_FD_VARIANT"], os.environ["ILQG_FD_BINS"]
    plain = torch.zeros_like(f["deriv"])
    st = torch.full((f["n"],), -1, dtype=torch.int32, device="cuda:0")
    h0.fd_batch_dev(f["q"], f["v"], f["u"], f["w"], plain, None, st, cost=f["cost"])
    torch.cuda.synchronize()
    assert torch.equal(plain, f["deriv"]) and torch.equal(st, f["status"])
    h0.close()


@pytest.mark.parametrize("n", [50011, 28673, 14337])
def test_host_pipeline_ragged_sizes(full, n):
    """The host-pointer call splits the batch into uniform chunks over an upload, several compute and a download stream; sizes that
    do not divide (ragged last chunk), that give an odd chunk count and that fall just above one chunk must return the bits of
    one device call over the same knots (forced split kernels: the chunks run the variant of the whole batch)."""
    import torch
    f = full; h = f["h"]
    sl = slice(1000, 1000 + n)
    dev = torch.zeros((n, h.model.nd), dtype=torch.float64, device="cuda:0")
    st = torch.zeros(n, dtype=torch.int32, device="cuda:0")
    h.fd_batch_dev(f["q"][sl].contiguous(), f["v"][sl].contiguous(), f["u"][sl].contiguous(), f["w"][sl].contiguous(), dev, None, st, cost=f["cost"])
    torch.cuda.synchronize()
    d, a, hs = h.fd_batch_host(f["q"][sl].cpu().numpy(), f["v"][sl].cpu().numpy(), f["u"][sl].cpu().numpy(), f["w"][sl].cpu().numpy(), f["cost"])
    assert np.array_equal(hs, st.cpu().numpy())
    assert np.array_equal(d, dev.cpu().numpy())
    assert np.array_equal(d, f["deriv"][sl].cpu().numpy())     # and of the full batch's slice


@pytest.mark.parametrize("env", [dict(ILQG_HOST_PRIO="0"), dict(ILQG_HOST_PRIO="0", ILQG_HOST_COMP="1"), dict(ILQG_HOST_CHUNKS="7", ILQG_HOST_COMP="3"),
                                 dict(ILQG_HOST_CHUNKS="12", ILQG_HOST_COMP="8"), dict(ILQG_HOST_CHUNKS="5", ILQG_HOST_COMP="2", ILQG_Q_MIX="37"),
                                 dict(ILQG_HOST_THREADS="0"), dict(ILQG_HOST_THREADS="3", ILQG_HOST_CHUNKS="6")])
def test_host_pipeline_stream_layouts(full, pkg, env):
    """The compute streams of the host-pointer call are a priority ladder by default (chunk c on stream c mod 4, stream j at priority
    greatest + j); ILQG_HOST_PRIO=0 is round 1's pair of plain alternating streams.  Every layout — more chunks than streams (two
    chunks share a stream and its work-class scratch), one stream, eight — must return the bits of one device call, pass after pass
    (a scratch slot shared by two chunks in flight would not necessarily show in one).  The last case also relabels the qpos
    kernel's CTAs (ILQG_Q_MIX): placement only."""
    f = full
    os.environ.update(dict(env, ILQG_FD_VARIANT="3"))
    try:
        h = pkg.Handle(pkg.Model.named("hopper"), 0)
    finally:
        for k in list(env) + ["ILQG_FD_VARIANT"]:
            del os.environ[k]
    n = 40000
    q, v, u, w = (f[k][:n].cpu().numpy() for k in ("q", "v", "u", "w"))
    ref = f["deriv"][:n].cpu().numpy()
    for _ in range(3):
        d, a, st = h.fd_batch_host(q, v, u, w, f["cost"])
        assert st.sum() == 0
        assert np.array_equal(d, ref)
    h.close()


def test_host_pipeline_pageable_and_pinned_buffers(full, pkg):
    """The arrays above are numpy arrays: PAGEABLE memory, which the call stages itself through a pinned mirror with a crew of copy
    threads (ILQG_HOST_THREADS=0: the driver's staging).  Same bits with page-locked caller buffers, with a mix of both, and without
    a device cost — then the caller's cost-gradient entries ride up with the deriv block and must come back untouched, and the
    Jacobian entries equal the ones computed with a cost."""
    import torch
    f = full; h = f["h"]; m = h.model
    n = 30011
    q, v, u, w = (f[k][:n].cpu() for k in ("q", "v", "u", "w"))
    ref = f["deriv"][:n].cpu().numpy()
    d_page, a_page, st = h.fd_batch_host(q.numpy(), v.numpy(), u.numpy(), w.numpy(), f["cost"])
    assert st.sum() == 0 and np.array_equal(d_page, ref)
    pin = [t.pin_memory() for t in (q, v, u, w)]
    d_pin = torch.zeros((n, m.nd), dtype=torch.float64).pin_memory()
    out, a_pin, st = h.fd_batch_host(*(t.numpy() for t in pin), f["cost"], deriv=d_pin.numpy())
    assert st.sum() == 0 and np.array_equal(d_pin.numpy(), ref) and np.array_equal(a_pin, a_page)
    d_mix, a_mix, st = h.fd_batch_host(pin[0].numpy(), v.numpy(), pin[2].numpy(), w.numpy(), f["cost"])   # pinned qpos / ctrl, pageable rest
    assert st.sum() == 0 and np.array_equal(d_mix, ref)
    njac = m.nv * (2 * m.nv + m.nu)
    mine = np.full((n, m.nd), 7.25)
    d_nc, _, st = h.fd_batch_host(q.numpy(), v.numpy(), u.numpy(), w.numpy(), None, deriv=mine)
    assert st.sum() == 0 and np.array_equal(d_nc[:, :njac], ref[:, :njac]) and (d_nc[:, njac:] == 7.25).all()


@pytest.mark.parametrize("env", [dict(ILQG_FD_VARIANT="1"), dict(ILQG_FD_VARIANT="1", ILQG_FD_COOP="1"), dict(ILQG_FD_VARIANT="2", ILQG_FD_PDL="0"), dict(ILQG_FD_VARIANT="2"),
                                 dict(ILQG_FD_VARIANT="3", ILQG_VU_CLASSES="8,16"), dict(ILQG_FD_VARIANT="3", ILQG_VU_POS="1"),
                                 dict(ILQG_FD_VARIANT="3", ILQG_VU_POS="2"), dict(ILQG_FD_VARIANT="3", ILQG_Q_MINB="1")])
def test_every_kernel_variant_gives_the_same_jacobians(full, pkg, env):
    """The kernel variants — the one-launch kernel for tiny batches, the two overlapped launches (with and without the programmatic
    dependent launch), the stage-skipping split, and the round-2 experiments that are off by default (shared-memory rows with the
    centre's Newton factor, position-stage records through HBM, one qpos CTA per SM) — differ in placement, launch
    structure and, for some, in the order of a few additions: same Jacobians to the FD tolerance on a slice of the full batch that
    mixes flight and stance, same status words, and idempotent."""
    import torch
    f = full
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        h = pkg.Handle(pkg.Model.named("hopper"), 0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    sl = slice(40000, 40000 + 30011)          # 30,011 knots: above the split threshold, ragged tail
    a = torch.zeros((30011, h.model.nd), dtype=torch.float64, device="cuda:0"); b = torch.zeros_like(a)
    st = torch.full((30011,), -1, dtype=torch.int32, device="cuda:0")
    args = [f[k][sl].contiguous() for k in ("q", "v", "u", "w")]
    h.fd_batch_dev(*args, a, None, st, cost=f["cost"])
    h.fd_batch_dev(*args, b, cost=f["cost"])
    torch.cuda.synchronize()
    if "ILQG_VU_CLASSES" in env:
        # a CTA that straddles two work classes is served by the kernel of its heaviest knot, and the ranks inside a class follow the
        # centre kernel's execution order: a knot at a class boundary may get either kernel from run to run — same value, other last bits
        assert float((a - b).abs().max()) <= 1e-12 * float(a.abs().max())
    else:
        diff = (a != b).any(dim=1)
        assert not bool(diff.any()), (int(diff.sum()), diff.nonzero()[:8].flatten().tolist(), float((a - b).abs().max()))
    assert int((st != 0).sum()) == 0
    assert_deriv_close(a.cpu().numpy(), f["deriv"][sl].cpu().numpy(), 6, 3, tol=1e-6)
    h.close()
