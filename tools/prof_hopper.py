"""Small hopper FD run for ncu (profiles/) and per-kernel event timings: n trajectories x 21 knots."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
name = sys.argv[3] if len(sys.argv) > 3 else "hopper"
model = pkg.Model.named(name)
h = pkg.Handle(model, 0)
q, v, u, w, nbad = wl.make_knots(h, ntraj, 21, seed=0, device="cuda:0", model=name)
n = q.shape[0]
deriv = torch.zeros((n, model.nd), dtype=torch.float64, device="cuda:0"); qacc = torch.zeros((n, model.nv), dtype=torch.float64, device="cuda:0")
status = torch.zeros(n, dtype=torch.int32, device="cuda:0")
cost = pkg.make_cost(q1=[1.0])
L = pkg.lib(); L.ilqg_set_profiling(h._h, 1)
for _ in range(3):
    h.fd_batch_dev(q, v, u, w, deriv, qacc, status, cost=cost)
torch.cuda.synchronize()
acc = np.zeros(3)
for _ in range(reps):
    h.fd_batch_dev(q, v, u, w, deriv, qacc, status, cost=cost)
    a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
    L.ilqg_fd_last_stage_ms(h._h, C.byref(a), C.byref(b), C.byref(c))
    acc += [a.value, b.value, c.value]
acc /= reps
print(f"{n} knots: centre {acc[0]:.4f} ms, velctrl {acc[1]:.4f} ms, qpos {acc[2]:.4f} ms -> {n / acc.sum() / 1e3:.2f} M knots/s; status ok {int((status == 0).sum())}")
