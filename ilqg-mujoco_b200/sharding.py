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
