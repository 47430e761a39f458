"""Multi-GPU partitioning of the hot path.  One process per GPU (torchrun); `torch.distributed` is plumbing.

Two natural shardings, nothing else:
  * independent trajectories / iLQR instances: a contiguous slice per rank, NO data-path collective
    (`shard_range`, `gather_results` for the optional final gather of small results);
  * the knots of ONE long horizon (BASELINE config 5, hopper T = 1000): FD at knot n depends only on the nominal
    (x_n, u_n, warm_n) (/root/reference/src/mjderivative.cpp:61,72), so knots shard contiguously and one all-gather
    of the `deriv` blocks (840 B per hopper knot) brings them to the rank(s) running the sequential Riccati sweep.
The Riccati recursion and the rollout are sequential in n: replicas only.
"""
import torch
import torch.distributed as dist


def shard_range(n, world, rank):
    """Contiguous, balanced slice [lo, hi) of n units for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pick_ranks(T, world, knots_per_rank_min=1000):
    """How many of the node's `world` GPUs to shard a horizon of T knots over.  A pass over up to ~1000 hopper knots is bound by the
    latency of one centre + one perturbed evaluation (60-80 us on B200 whether a rank holds 125 knots or 1000) and sharding adds
    the flag barrier (~10 us).  Measured on one 8 x B200 node (tools/prof_t_sweep.py, round 2; us per pass at 1 / 2 / 4 / 8 ranks):
    T = 1000: 78 / 87 / 88 / 87;  2000: 107 / 89 / 88 / 90;  4000: 152 / 115 / 91 / 89;  8000: 244 / 162 / 120 / 96;
    16000: 435 / 256 / 167 / 126;  64000: 1171 / 717 / 472 / 290;  128000: 2238 / 1225 / 796 / 506.
    So doubling the ranks pays as long as every rank keeps at least ~1000 knots.  Returns a power of two <= world."""
    r = 1
    while r * 2 <= world and T // (r * 2) >= knots_per_rank_min:
        r *= 2
    return r


def padded_count(n, world):
    """Per-rank unit count after padding n to a multiple of world (all_gather needs equal contributions)."""
    return (n + world - 1) // world


def fd_knot_sharded(compute_fn, qpos, qvel, ctrl, warm, nd, group=None):
    """FD linearisation of one long trajectory, knots sharded over the ranks of `group`.

    compute_fn(qpos, qvel, ctrl, warm) -> deriv[nk_local, nd] runs the local slice (on the GPU:
    Handle.fd_batch_dev; in CPU tests: the oracle).  Every rank holds the full nominal (the reference's
    rollout is deterministic, so ranks either re-run it or receive a broadcast); returns the full
    deriv[T, nd] on every rank, knots in their original order.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    T = qpos.shape[0]
    per = padded_count(T, world)
    lo = min(rank * per, T)
    hi = min(lo + per, T)
    local = torch.zeros((per, nd), dtype=torch.float64, device=qpos.device)
    if hi > lo:
        local[: hi - lo] = compute_fn(qpos[lo:hi].contiguous(), qvel[lo:hi].contiguous(), ctrl[lo:hi].contiguous(), warm[lo:hi].contiguous())
    if world == 1:
        return local[:T]
    full = torch.empty((world * per, nd), dtype=torch.float64, device=qpos.device)
    dist.all_gather_into_tensor(full, local, group=group)
    return full[:T]


class _DevArray:
    """Expose a raw device allocation to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}


class PeerDeriv:
    """One deriv[T, nd] array per rank, each mapped into every other rank of the node with CUDA IPC, so that the FD kernels
    of a knot-sharded horizon store their blocks straight into all copies (ilqg_fd_batch_dev_scatter): the all-gather of
    SURVEY 8(e) rides in the kernels' write-out over NVLink instead of following them as a separate collective.

    The array is DOUBLE-BUFFERED by pass parity: the flag barrier that closes pass k only proves that every rank's stores of
    pass k have landed; a faster rank may already be storing pass k+1 while a slower one still reads pass k, so pass k+1 goes
    to the other buffer.  Buffer k & 1 is stored to again in pass k+2, which no rank can start before every rank has reached
    the barrier of pass k+1 — and a rank issues that barrier after its own readers of pass k PROVIDED THEY RUN ON THE SAME
    STREAM as the FD calls (or are otherwise ordered before the next `fd_knot_sharded_peer` call).  That is the contract."""

    def __init__(self, handle, T, nd, group=None):
        self.h, self.T, self.nd, self.group = handle, int(T), int(nd), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.buf_bytes = self.T * self.nd * 8
        self.flag_off = 2 * self.buf_bytes             # the barrier's flag array sits behind the two buffers (zero-initialised)
        self.epoch = 0
        self.own, hd = handle.peer_alloc(self.flag_off + 256)
        handles = [hd]
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, hd, group=group)
        self.ptrs = [self.own if r == self.rank else handle.peer_open(handles[r]) for r in range(self.world)]
        if hasattr(handle, "peer_view"):   # host-memory stand-in used by the CPU (gloo) tests of this logic
            both = handle.peer_view(self.own, (2, self.T, self.nd))
        else:
            both = torch.as_tensor(_DevArray(self.own, (2, self.T, self.nd)), device=f"cuda:{torch.cuda.current_device()}")
        self.bufs = [both[0], both[1]]
        self.full = self.bufs[0]

    def scatter_ptrs(self, first_knot, parity=0, root=None):
        """Destination addresses of knot `first_knot` in buffer `parity`: every rank's copy, this rank's first — or, with `root`,
        only the copy of the rank that runs the backward pass (the all-gather becomes a gather: 1/world of the NVLink traffic)."""
        order = [self.rank] + [r for r in range(self.world) if r != self.rank] if root is None else [int(root)]
        return [self.ptrs[r] + parity * self.buf_bytes + first_knot * self.nd * 8 for r in order]

    def check(self):
        """Synchronises; raises if a barrier gave up waiting for a peer (the arrays then hold a partial pass)."""
        if self.h.peer_barrier_timed_out():
            raise RuntimeError("ilqg peer barrier timed out: a rank of the node did not finish its FD pass; the gathered deriv array is incomplete")

    def barrier(self, stream=None):
        """Every rank's preceding kernels (and their peer stores) are complete and visible once this returns on the stream."""
        self.epoch += 1
        self.h.peer_barrier([p + self.flag_off for p in self.ptrs], self.rank, self.epoch, stream=stream)

    def close(self):
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)      # nobody unmaps while a peer may still be storing
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self.h.peer_close(p)
        self.full = None
        self.bufs = None
        self.h.peer_free(self.own)


def fd_knot_sharded_peer(handle, peer, qpos, qvel, ctrl, warm, cost=None, stream=None, root=None):
    """FD linearisation of one long trajectory, knots sharded over the ranks, blocks stored by the kernels into every
    rank's copy of the pass's buffer.  Returns that buffer (all T blocks, knot order; also `peer.full`), valid on `stream`
    once every rank's kernels have finished.  Consecutive passes alternate between two buffers; readers of the returned array
    must be ordered on `stream` before the next call (see PeerDeriv).  root = r: the blocks go to rank r's copy only (the rank
    that runs the Riccati sweep, /root/reference/inc/ilqr.h:144-175); the array returned on the other ranks is then not filled."""
    T = qpos.shape[0]
    lo, hi = shard_range(T, peer.world, peer.rank)
    parity = peer.epoch & 1
    if hi > lo:
        handle.fd_batch_dev_scatter(qpos[lo:hi], qvel[lo:hi], ctrl[lo:hi], warm[lo:hi], peer.scatter_ptrs(lo, parity, root), cost=cost, stream=stream)
    peer.full = peer.bufs[parity]
    if peer.world > 1:
        peer.barrier(stream=stream)   # a 1-warp kernel per rank exchanging flags through peer memory: no NCCL call on this path
    else:
        peer.epoch += 1
    return peer.full


def gather_results(local, n_total, group=None):
    """Optional final gather of small per-instance results (first control, cost) from instance-sharded ranks."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    per = padded_count(n_total, world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    full = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, pad, group=group)
    # ranks hold balanced contiguous slices (shard_range); drop each rank's padding
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, world, r)
        parts.append(full[r * per: r * per + (hi - lo)])
    return torch.cat(parts, 0)
