"""Experiment: FD kernels storing deriv straight into pinned host memory (zero-copy over PCIe) vs device memory + copy."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
model = pkg.Model.named("hopper")
h = pkg.Handle(model, 0)
q, v, u, w, _ = wl.make_knots(h, 4096, 21, seed=0, device="cuda:0", model="hopper")
n = q.shape[0]
cost = pkg.make_cost(q1=[1.0])
d_dev = torch.zeros((n, model.nd), dtype=torch.float64, device="cuda:0")
d_host = torch.zeros((n, model.nd), dtype=torch.float64).pin_memory()
qacc = torch.zeros((n, model.nv), dtype=torch.float64, device="cuda:0")
status = torch.zeros(n, dtype=torch.int32, device="cuda:0")
def run(deriv, reps=20):
    for _ in range(3): h.fd_batch_dev(q, v, u, w, deriv, qacc, status, cost=cost)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): h.fd_batch_dev(q, v, u, w, deriv, qacc, status, cost=cost)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print(f"device deriv: {run(d_dev):.3f} ms")
print(f"pinned-host deriv (zero-copy stores): {run(d_host):.3f} ms -> {n*model.nd*8/run(d_host)/1e6:.1f} GB/s")
h.fd_batch_dev(q, v, u, w, d_dev, qacc, status, cost=cost); torch.cuda.synchronize()
print("equal:", bool(torch.equal(d_dev.cpu(), d_host)))
