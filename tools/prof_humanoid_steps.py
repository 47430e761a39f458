import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
h = pkg.Handle(pkg.Model.named("humanoid"), 0)
n = 1184
q, v, u, w, _ = wl.humanoid_states(h, n, seed=50)
m = h.model
deriv = torch.zeros((n, m.nd), dtype=torch.float64, device="cuda:0")
diag = torch.zeros((n, 8), dtype=torch.int32, device="cuda:0")
h.fd_set_diag(diag)
h.fd_batch_dev(q, v, u, w, deriv)
torch.cuda.synchronize()
d = diag.cpu().numpy()
print("nefc", np.percentile(d[:, 0], [0, 25, 50, 75, 100]))
print("first-solve iterations", dict(zip(*np.unique(d[:, 1], return_counts=True))))
print("build cycles median", np.median(d[:, 4]), "solve cycles median", np.median(d[:, 5]), "max", d[:, 5].max())
h.fd_set_diag(None)
for ns in (1, 11):
    qq, vv, ww = q.clone(), v.clone(), w.clone()
    h.step_batch_dev(qq, vv, u * 0, ww, None, nsteps=ns)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    qq, vv, ww = q.clone(), v.clone(), w.clone()
    e0.record(); h.step_batch_dev(qq, vv, u * 0, ww, None, nsteps=ns); e1.record(); e1.synchronize()
    print(f"step_batch n={n} nsteps={ns}: {e0.elapsed_time(e1):.3f} ms")
