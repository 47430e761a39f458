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


class _HostPeerHandle:
    """Stand-in for pkg.Handle on a CPU-only box: 'peer-visible buffers' are POSIX shared-memory segments, the FD 'kernel' is the
    oracle writing each block to every destination address, the flag barrier is exercised through the same epoch protocol on the
    shared flags.  It lets the world_size-2 gloo test run the real PeerDeriv / fd_knot_sharded_peer host logic (slot addressing,
    destination order, first-knot offsets, epochs) without a GPU."""

    def __init__(self, oracle, om):
        from multiprocessing import shared_memory
        self._shm_mod = shared_memory
        self.o, self.om, self.segs = oracle, om, {}

    def _addr(self, shm):
        import ctypes
        return ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))

    def peer_alloc(self, nbytes):
        shm = self._shm_mod.SharedMemory(create=True, size=nbytes)
        shm.buf[:nbytes] = bytes(nbytes)
        a = self._addr(shm)
        self.segs[a] = shm
        return a, shm.name.encode().ljust(64, b"\0")

    def peer_open(self, handle_bytes):
        shm = self._shm_mod.SharedMemory(name=handle_bytes.rstrip(b"\0").decode())
        a = self._addr(shm)
        self.segs[a] = shm
        return a

    def peer_view(self, ptr, shape):
        import ctypes
        n = int(np.prod(shape))
        return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_double * n).from_address(ptr)).reshape(shape))

    def fd_batch_dev_scatter(self, q, v, u, w, dst_ptrs, cost=None, stream=None):
        import ctypes
        d, _, _ = self.o.fd_batch(self.om, q.numpy(), v.numpy(), u.numpy(), w.numpy(), cost, nthreads=1)
        d = np.ascontiguousarray(d)
        for p in dst_ptrs:
            ctypes.memmove(int(p), d.ctypes.data, d.nbytes)

    def peer_barrier(self, flag_ptrs, rank, epoch, stream=None):
        import ctypes, time
        for p in flag_ptrs:                       # release-store my epoch into slot [rank] of every rank's flag array
            ctypes.c_int.from_address(int(p) + 4 * rank).value = epoch
        mine = flag_ptrs[rank]
        t0 = time.time()
        for r in range(len(flag_ptrs)):           # acquire-spin on my own array
            while ctypes.c_int.from_address(int(mine) + 4 * r).value < epoch:
                assert time.time() - t0 < 60, "peer barrier timed out"
                time.sleep(0.001)

    def peer_close(self, ptr):
        self.segs.pop(ptr).close()

    def peer_free(self, ptr):
        shm = self.segs.pop(ptr)
        shm.close()
        shm.unlink()


def _peer_worker(rank, world, port, T, tmpdir):
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
    h = _HostPeerHandle(o, om)
    peer = sharding.PeerDeriv(h, T, om.nd)
    assert peer.scatter_ptrs(3)[0] == peer.own + 3 * om.nd * 8          # this rank's copy first, knot offset in bytes
    assert peer.scatter_ptrs(3, 1)[0] == peer.own + (T + 3) * om.nd * 8  # the second buffer follows the first
    import time
    for p in range(3):                                                    # several passes with DIFFERENT inputs: epochs 1, 2, 3
        q, v, u, w = scenario_states("hopper", T, seed=5 + p)
        ref, _, _ = o.fd_batch(om, q, v, u, w, None, nthreads=1)
        tq, tv, tu, tw = (torch.from_numpy(a) for a in (q, v, u, w))
        full = sharding.fd_knot_sharded_peer(h, peer, tq, tv, tu, tw)
        assert full.data_ptr() == peer.bufs[p & 1].data_ptr()
        if rank == 1:
            time.sleep(0.3)   # a slow reader: rank 0 is already storing the next pass — into the OTHER buffer
        assert np.array_equal(full.numpy(), ref)
    np.save(os.path.join(tmpdir, f"peer_{rank}.npy"), full.numpy().copy())
    assert peer.epoch == 3
    # gather to the backward-pass rank only: one destination per store
    assert peer.scatter_ptrs(2, 1, root=1) == [peer.ptrs[1] + (T + 2) * om.nd * 8]
    q, v, u, w = scenario_states("hopper", T, seed=9)
    full = sharding.fd_knot_sharded_peer(h, peer, *(torch.from_numpy(a) for a in (q, v, u, w)), root=0)
    if rank == 0:
        ref, _, _ = o.fd_batch(om, q, v, u, w, None, nthreads=1)
        assert np.array_equal(full.numpy(), ref)
    assert peer.epoch == 4
    dist.barrier()
    for r, p in enumerate(peer.ptrs):
        if r != rank:
            h.peer_close(p)
    dist.barrier()
    peer.full = None
    peer.bufs = None
    full = None
    h.peer_free(peer.own)
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [21, 10])
def test_knot_sharded_fd_peer_store_gather_host_logic(tmp_path, oracle, omodels, T):
    """The peer-store route of BASELINE config 5 (PeerDeriv + fd_knot_sharded_peer) with host-memory stand-ins for the CUDA-IPC
    buffers, the FD kernels and the flag barrier: every rank ends with all T blocks in knot order, each computed by one rank only."""
    world = 2
    port = 29600 + (os.getpid() % 2000) + T
    mp.spawn(_peer_worker, args=(world, port, T, str(tmp_path)), nprocs=world, join=True)
    om = omodels["hopper"]
    q, v, u, w = scenario_states("hopper", T, seed=7)   # the last pass's inputs
    ref, _, _ = oracle.fd_batch(om, q, v, u, w, None, nthreads=1)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"peer_{r}.npy"), ref)
