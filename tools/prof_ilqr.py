"""Per-phase timing of the batched iLQR iteration (BASELINE configs[3]): rollout+accept / FD / Riccati, CUDA events."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
ninst = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nalpha = int(sys.argv[3]) if len(sys.argv) > 3 else 1
name = sys.argv[4] if len(sys.argv) > 4 else "inverted_pendulum"
dev = "cuda:0"
model = pkg.Model.named(name)
h = pkg.Handle(model, 0)
L = pkg.lib()
N = 20
if name == "humanoid":   # bench.py's bench_humanoid_ilqr: tangent-space extension, N = 10
    N = 10
    dq, dv, du, dw, _ = wl.humanoid_states(h, ninst, seed=50, device=dev)
    du = du * 0.0
    cost = pkg.make_cost(q2=[0, 0, 2.0, 0, 1, 1, 0], q1=[0, 0, -5.2], v2=[0.05] * 27, u2=[0.02] * 21)
else:
    if name == "inverted_pendulum":
        q, v, u, _ = wl.pendulum_initial_states(ninst, seed=100)
        cost = pkg.make_cost(q2=[1, 10], v2=[1, 10], u2=[1])
    else:
        q, v, u, _ = wl.hopper_initial_states(ninst, seed=100)
        cost = pkg.make_cost(q2=[0, 1, 1, 0, 0, 0], v2=[1] * 6, u2=[0.1] * 3)
    u = u * 0.0
    dq, dv, du = (torch.from_numpy(a).to(dev) for a in (q, v, u))
    dw = torch.zeros((ninst, model.nv), dtype=torch.float64, device=dev)
il = pkg.Ilqr(h, ninst, N, tuple(0.5 ** a for a in range(nalpha)))
il.set_cost(cost)
if name == "humanoid":
    il.set_layout(True)
    il.set_mu(1000.0)
s = torch.cuda.current_stream().cuda_stream
sp = C.c_void_p(s)
for rep in range(3):
    il.init_dev(dq, dv, du, dw, stream=s)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * niter + 1)]
    ev[0].record()
    for it in range(niter):
        L.ilqg_ilqr_forward(il._w, 1 if nalpha == 1 else 0, sp); ev[3 * it + 1].record()
        L.ilqg_ilqr_linearise(il._w, sp); ev[3 * it + 2].record()
        L.ilqg_ilqr_backward(il._w, sp); ev[3 * it + 3].record()
    torch.cuda.synchronize()
    ph = np.zeros(3)
    for it in range(niter):
        for j in range(3):
            ph[j] += ev[3 * it + j].elapsed_time(ev[3 * it + j + 1])
    tot = ev[0].elapsed_time(ev[-1])
    print(f"{name} ninst={ninst} nalpha={nalpha}: per iteration forward {ph[0]/niter:.4f} ms, FD {ph[1]/niter:.4f} ms, backward {ph[2]/niter:.4f} ms; "
          f"total {tot/niter:.4f} ms -> {ninst*niter/tot/1e3:.2f} M its/s")
