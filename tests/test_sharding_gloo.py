"""Host-side logic of the N > 1 paths on CPU: world_size-2 gloo process groups, the oracle standing in for the GPU compute."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, scenario_states


def _worker(rank, world, port, T, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    pkg = entry.load_package()
    o = entry.load_oracle()
    from ilqg_mujoco_b200 import sharding
    om = o.Model(os.path.join(pkg.MODELS_DIR, "hopper.ilqgm"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    q, v, u, w = scenario_states("hopper", T, seed=5)   # identical on every rank (deterministic nominal)
    tq, tv, tu, tw = (torch.from_numpy(a) for a in (q, v, u, w))
    calls = []

    def compute(qq, vv, uu, ww):
        calls.append(qq.shape[0])
        d, _, _ = o.fd_batch(om, qq.numpy(), vv.numpy(), uu.numpy(), ww.numpy(), None, nthreads=1)
        return torch.from_numpy(d)

    full = sharding.fd_knot_sharded(compute, tq, tv, tu, tw, om.nd)
    # instance sharding + final gather
    lo, hi = sharding.shard_range(T, world, rank)
    mine = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)
    gathered = sharding.gather_results(mine, T)
    np.save(os.path.join(tmpdir, f"full_{rank}.npy"), full.numpy())
    np.save(os.path.join(tmpdir, f"gath_{rank}.npy"), gathered.numpy())
    np.save(os.path.join(tmpdir, f"calls_{rank}.npy"), np.array(calls))
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [21, 10])
def test_knot_sharded_fd_all_gather(tmp_path, oracle, omodels, T):
    world = 2
    port = 29500 + (os.getpid() % 2000) + T
    mp.spawn(_worker, args=(world, port, T, str(tmp_path)), nprocs=world, join=True)
    om = omodels["hopper"]
    q, v, u, w = scenario_states("hopper", T, seed=5)
    ref, _, _ = oracle.fd_batch(om, q, v, u, w, None, nthreads=1)
    for r in range(world):
        full = np.load(tmp_path / f"full_{r}.npy")
        assert full.shape == ref.shape and np.array_equal(full, ref)      # every rank ends with all blocks, in knot order
        g = np.load(tmp_path / f"gath_{r}.npy")
        assert np.array_equal(g[:, 0], np.arange(T))
    c0, c1 = np.load(tmp_path / "calls_0.npy"), np.load(tmp_path / "calls_1.npy")
    assert c0.sum() + c1.sum() == T and abs(int(c0.sum()) - int(c1.sum())) <= (T + 1) // 2   # work is split, not replicated


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.load_package()
    from ilqg_mujoco_b200 import sharding
    for n in (0, 1, 7, 4096, 1000):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                lo, hi = sharding.shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
            assert got == list(range(n))
            assert sharding.padded_count(n, world) * world >= n
