"""Small humanoid FD run for ncu (profiles/): 256 knots through the generic warp-per-rollout engine."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = pkg.Handle(pkg.Model.named("humanoid"), 0)
q, v, u, w, nbad = wl.humanoid_states(h, n, seed=0)
m = h.model
deriv = torch.zeros((n, m.nd), dtype=torch.float64, device="cuda:0"); qacc = torch.zeros((n, m.nv), dtype=torch.float64, device="cuda:0")
status = torch.zeros(n, dtype=torch.int32, device="cuda:0")
for _ in range(3):
    h.fd_batch_dev(q, v, u, w, deriv, qacc, status, cost=None)
torch.cuda.synchronize()
import ctypes as C
L = pkg.lib(); L.ilqg_set_profiling(h._h, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); h.fd_batch_dev(q, v, u, w, deriv, qacc, status, cost=None); e1.record(); e1.synchronize()
a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
L.ilqg_fd_last_stage_ms(h._h, C.byref(a), C.byref(b), C.byref(c))
ms = e0.elapsed_time(e1)
print(f"humanoid {n} knots: {ms:.3f} ms -> {n / ms:.1f} K knots/s (first chunk: centre {a.value:.3f}, velctrl {b.value:.3f}; rest incl. qpos {c.value:.3f}); status ok {int((status == 0).sum())}")
