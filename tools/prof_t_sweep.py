"""Knot-sharded hopper FD (BASELINE configs[4]) over the ranks of a torchrun launch, for a sweep of horizon lengths: from which
T do N GPUs beat one?  (sharding.pick_ranks encodes the answer.)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/prof_t_sweep.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import __graft_entry__ as e

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
pkg = e.load_package()
from ilqg_mujoco_b200 import sharding, workload as wl
model = pkg.Model.named("hopper")
h = pkg.Handle(model, local)
q, v, u, w, _ = wl.make_knots(h, 1, 1000, seed=0, device=dev, model="hopper")
stream = torch.cuda.current_stream().cuda_stream
rows = []
root = int(os.environ["ILQG_ROOT"]) if os.environ.get("ILQG_ROOT") else None   # gather to one rank instead of all
for T in (1000, 2000, 4000, 8000, 16000, 32000, 64000, 128000):
    rep = (T + 999) // 1000
    ql, vl, ul, wl_ = (x.repeat(rep, 1)[:T].contiguous() for x in (q, v, u, w))
    peer = sharding.PeerDeriv(h, T, model.nd)
    for _ in range(3):
        sharding.fd_knot_sharded_peer(h, peer, ql, vl, ul, wl_, stream=stream, root=root)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if T <= 16000 else 6
    e0.record()
    for _ in range(reps):
        sharding.fd_knot_sharded_peer(h, peer, ql, vl, ul, wl_, stream=stream, root=root)
    e1.record(); e1.synchronize()
    peer.check()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rows.append((T, float(t[0])))
    peer.close()
if rank == 0:
    print(f"ranks={world} root={root}: " + "  ".join(f"T={T}: {ms * 1e3:.0f} us {T / ms / 1e3:.1f} M/s" for T, ms in rows))
h.close()
if world > 1:
    dist.destroy_process_group()
