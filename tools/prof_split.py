"""A/B of the large-batch (stage-skipping split) hopper FD kernels on the BASELINE configs[1] batch: per-kernel CUDA-event times
for several settings of the shared-memory row capacity of the qvel/ctrl kernel (ILQG_VU_CLASSES, e.g. "8,16"; 0 = rows in local memory).

    python tools/prof_split.py [ntraj] [reps] [caps...]
"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
ntraj = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
caps = sys.argv[3:] or ["0", "8", "16", "8,16"]
model = pkg.Model.named("hopper")
L = pkg.lib()
h0 = pkg.Handle(model, 0)
if os.environ.get("ILQG_WORKLOAD") == "stance":   # a hopper standing / bouncing on the ground: > 90 % of the knots in contact
    q, v, u, w, _ = wl.make_knots(h0, 1, 1000, seed=0, device="cuda:0", model="hopper")
    q, v, u, w = (x.repeat((ntraj * 21 + 999) // 1000, 1)[:ntraj * 21].contiguous() for x in (q, v, u, w))
elif os.environ.get("ILQG_WORKLOAD") == "r1":     # round 1's batch: 23 % of the knots in contact, constant control
    q, v, u, w, nbad = wl.make_knots(h0, ntraj, 21, seed=0, device="cuda:0", model="hopper")
else:                                              # the SURVEY 8d batch: half of the knots in contact, time-varying control
    q, v, u, w, nbad = wl.make_knots_8d(h0, ntraj, 21, seed=0, device="cuda:0")
n = q.shape[0]
cost = pkg.make_cost(q1=[1.0])
ref = None
for cap in caps:
    envs = dict(kv.split("=") for kv in cap.split(";") if "=" in kv)   # "ILQG_Q_MINB=1;ILQG_VU_CLASSES=8" or a bare class list
    if not envs:
        envs = {"ILQG_VU_CLASSES": cap}
    os.environ.update(envs)
    h = pkg.Handle(model, 0)
    for k in envs:
        del os.environ[k]
    deriv = torch.zeros((n, model.nd), dtype=torch.float64, device="cuda:0"); qacc = torch.zeros((n, model.nv), dtype=torch.float64, device="cuda:0")
    status = torch.zeros(n, dtype=torch.int32, device="cuda:0")
    L.ilqg_set_profiling(h._h, 1)
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
    if ref is None:
        ref = deriv.clone()
    err = float((deriv - ref).abs().max() / ref.abs().max())
    print(f"{cap:>28s}: {n} knots: centre {acc[0]:.4f} ms, velctrl {acc[1]:.4f} ms, qpos {acc[2]:.4f} ms -> {n / acc.sum() / 1e3:.2f} M knots/s; "
          f"status ok {int((status == 0).sum())}; max rel diff to first {err:.2e}")
    h.close()
