"""Knot-sharded FD with the gather fused into the kernels' write-out (ilqg_fd_batch_dev_scatter + CUDA IPC peer buffers).

1 GPU: several destinations on the same device receive identical blocks.  >= 2 GPUs (skipped otherwise): two ranks, NCCL
for the plumbing, each linearises half of the knots and stores its blocks into both ranks' arrays over NVLink; every rank
must end with exactly what a single GPU computes for the whole horizon."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, scenario_states

pytestmark = pytest.mark.gpu


def test_scatter_to_several_local_destinations(pkg, oracle, omodels):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    h = pkg.Handle(pkg.Model.named("hopper"), 0)
    m = h.model
    q, v, u, w = scenario_states("hopper", 53, seed=11, oracle=oracle, om=omodels["hopper"], roll=120)
    dq, dv, du, dw = (torch.from_numpy(a).cuda() for a in (q, v, u, w))
    cost = pkg.make_cost(q1=[1.0])
    ref = torch.zeros((53, m.nd), dtype=torch.float64, device="cuda")
    h.fd_batch_dev(dq, dv, du, dw, ref, cost=cost)
    from ilqg_mujoco_b200 import sharding
    bufs = [h.peer_alloc(60 * m.nd * 8)[0] for _ in range(3)]
    first = 5   # the range's blocks land at knot offset 5 of every destination
    h.fd_batch_dev_scatter(dq, dv, du, dw, [b + first * m.nd * 8 for b in bufs], cost=cost)
    torch.cuda.synchronize()
    for b in bufs:
        full = torch.as_tensor(sharding._DevArray(b, (60, m.nd)), device="cuda:0")
        assert torch.equal(full[first:first + 53], ref)
        assert float(full[:first].abs().sum()) == 0.0 and float(full[first + 53:].abs().sum()) == 0.0
    for b in bufs:
        h.peer_free(b)
    h.close()


def _rank(rank, world, port, T, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as entry
    pkg = entry.load_package()
    o = entry.load_oracle()
    from ilqg_mujoco_b200 import sharding
    om = o.Model(os.path.join(pkg.MODELS_DIR, "hopper.ilqgm"))
    q, v, u, w = scenario_states("hopper", T, seed=5, oracle=o, om=om, roll=100)
    dev = f"cuda:{rank}"
    dq, dv, du, dw = (torch.from_numpy(a).to(dev) for a in (q, v, u, w))
    h = pkg.Handle(pkg.Model.named("hopper"), rank)
    peer = sharding.PeerDeriv(h, T, h.model.nd)
    cost = pkg.make_cost(q1=[1.0])
    # repeated passes exercise the barrier's epochs and the two buffers; the inputs differ from pass to pass (the controls are
    # scaled) and rank 1 reads each pass's array late (a long kernel ahead of the copy), while rank 0 is already storing the
    # next pass — into the other buffer
    spin = torch.zeros(1 << 24, dtype=torch.float64, device=dev)
    snaps = []
    for p in range(6):
        full = sharding.fd_knot_sharded_peer(h, peer, dq, dv, du * (1.0 - 0.1 * (5 - p)), dw, cost=cost)
        assert full.data_ptr() == peer.bufs[p & 1].data_ptr()
        if rank == 1:
            for _ in range(20):
                spin.add_(1.0)
        snaps.append(full.clone())   # the reader, ordered on the stream before the next pass, as the contract asks
    torch.cuda.synchronize()
    for p in range(6):
        one = torch.zeros((T, h.model.nd), dtype=torch.float64, device=dev)
        h.fd_batch_dev(dq, dv, du * (1.0 - 0.1 * (5 - p)), dw, one, cost=cost)
        assert torch.equal(snaps[p], one), p
    peer.check()   # synchronises; raises on a barrier timeout
    np.save(os.path.join(out, f"peer_{rank}.npy"), full.cpu().numpy())
    if rank == 0:
        ref = torch.zeros((T, h.model.nd), dtype=torch.float64, device=dev)
        h.fd_batch_dev(dq, dv, du, dw, ref, cost=cost)
        np.save(os.path.join(out, "single.npy"), ref.cpu().numpy())
    peer.close()
    h.close()
    dist.destroy_process_group()


def test_two_ranks_store_into_each_others_arrays(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on one node")
    T = 101
    mp.spawn(_rank, args=(2, 29700 + os.getpid() % 200, T, str(tmp_path)), nprocs=2, join=True)
    single = np.load(tmp_path / "single.npy")
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"peer_{r}.npy"), single)


def test_generic_engine_scatters_too(pkg, oracle, omodels):
    """The warp-cooperative engine (humanoid) stores its blocks to several destinations as well (round 1 refused dst.n > 1)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ilqg_mujoco_b200 import sharding
    h = pkg.Handle(pkg.Model.named("humanoid"), 0)
    m = h.model
    n = 7
    q, v, u, w = scenario_states("humanoid", n, seed=19, oracle=oracle, om=omodels["humanoid"], roll=40)
    dq, dv, du, dw = (torch.from_numpy(a).cuda() for a in (q, v, u, w))
    cost = pkg.make_cost(q1=[1.0])
    ref = torch.zeros((n, m.nd), dtype=torch.float64, device="cuda")
    h.fd_batch_dev(dq, dv, du, dw, ref, cost=cost)
    bufs = [h.peer_alloc(10 * m.nd * 8)[0] for _ in range(2)]
    h.fd_batch_dev_scatter(dq, dv, du, dw, [b + 2 * m.nd * 8 for b in bufs], cost=cost)
    torch.cuda.synchronize()
    for b in bufs:
        full = torch.as_tensor(sharding._DevArray(b, (10, m.nd)), device="cuda:0")
        assert torch.equal(full[2:2 + n], ref)
        assert float(full[:2].abs().sum()) == 0.0 and float(full[2 + n:].abs().sum()) == 0.0
    for b in bufs:
        h.peer_free(b)
    h.close()
