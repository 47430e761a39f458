"""Per-knot cycle counts of the one-launch FD kernel on the T = 1000 hopper horizon (ilqg_fd_set_diag): stages, centre solves,
perturbed solves; the slowest knots of each phase.   python tools/prof_fused_diag.py [T] [ILQG_FD_COOP]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ILQG_FD_VARIANT", "1")
if len(sys.argv) > 2:
    os.environ["ILQG_FD_COOP"] = sys.argv[2]
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
sys.argv = sys.argv[:2]
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
from ilqg_mujoco_b200 import workload as wl
dev = "cuda:0"
model = pkg.Model.named("hopper")
h = pkg.Handle(model, 0)
q, v, u, w, _ = wl.make_knots(h, 1, T, seed=0, device=dev, model="hopper")
tt = torch.arange(T, device=dev, dtype=torch.float64)[:, None] * 0.002
gen = torch.Generator(device="cpu").manual_seed(0)
amp = torch.rand((3, 3), generator=gen, dtype=torch.float64).to(dev) * 0.3
frq = (torch.rand((3, 3), generator=gen, dtype=torch.float64) * 6 + 1).to(dev)
u = sum(amp[i][None, :] * torch.sin(2 * np.pi * frq[i][None, :] * tt + i) for i in range(3)).contiguous()
qs, vs, ws = q[:1].clone(), v[:1].clone(), w[:1].clone()
Q, V, W = [], [], []
for t in range(T):
    Q.append(qs.clone()); V.append(vs.clone()); W.append(ws.clone())
    h.step_batch_dev(qs, vs, u[t:t + 1].contiguous(), ws, None, nsteps=1)
q, v, w = torch.cat(Q), torch.cat(V), torch.cat(W)
deriv = torch.zeros((T, model.nd), dtype=torch.float64, device=dev)
diag = torch.zeros((T, 8), dtype=torch.int32, device=dev)
for _ in range(3):
    h.fd_batch_dev(q, v, u, w, deriv)
h.fd_set_diag(diag)
h.fd_batch_dev(q, v, u, w, deriv)
torch.cuda.synchronize()
h.fd_set_diag(None)
d = diag.cpu().numpy()
print("coop =", os.environ.get("ILQG_FD_COOP", "1"))
for name, c in (("build", 4), ("centre solves", 5), ("perturbed solves", 7)):
    x = d[:, c]
    print(f"{name:18s} cycles: median {np.median(x):.0f}  p90 {np.percentile(x, 90):.0f}  p99 {np.percentile(x, 99):.0f}  max {x.max()}")
for name, c in (("centre solves", 5), ("perturbed solves", 7)):
    o = np.argsort(-d[:, c])[:8]
    print(f"slowest {name}: (knot, nefc, iters, nactive, ncon, build, centre, columns)")
    for k in o:
        print("   ", k, d[k, 0], d[k, 1], d[k, 3], d[k, 6], d[k, 4], d[k, 5], d[k, 7])
print("centre solve cycles by (nefc bucket, iterations):")
for lo, hi in ((0, 0), (1, 4), (5, 8), (9, 12), (13, 20), (21, 41)):
    for it in range(0, 6):
        sel = (d[:, 0] >= lo) & (d[:, 0] <= hi) & (d[:, 1] == it)
        if sel.any():
            print(f"   nefc {lo}-{hi} iters {it}: {int(sel.sum())} knots, centre median {np.median(d[sel, 5]):.0f}, columns median {np.median(d[sel, 7]):.0f}")
