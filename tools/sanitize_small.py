"""Tiny invocations of every kernel family (a driver for compute-sanitizer where that tool is available; it is closed on the B200 pool):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
hopper FD through every kernel variant, pendulum FD + batched iLQR, humanoid FD (warp-cooperative engine, in contact) + tangent-space iLQR,
forward / step batches.  Prints one line per case; the sanitizer's summary is the result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
dev = "cuda:0"
cost = pkg.make_cost(q2=[1, 10], v2=[1, 10], u2=[1], q1=[0.5])


def handle(name, **env):
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return pkg.Handle(pkg.Model.named(name), 0)
    finally:
        for k in env:
            del os.environ[k]


h0 = handle("hopper")
q, v, u, w, _ = wl.make_knots_8d(h0, 3, 21, seed=3, device=dev)     # 63 knots, flight and stance
n = q.shape[0]
for env in ({}, dict(ILQG_FD_VARIANT=1), dict(ILQG_FD_VARIANT=1, ILQG_FD_COOP=1), dict(ILQG_FD_VARIANT=2), dict(ILQG_FD_VARIANT=3),
            dict(ILQG_FD_VARIANT=3, ILQG_VU_CLASSES="8,16"), dict(ILQG_FD_VARIANT=3, ILQG_VU_POS=1), dict(ILQG_FD_VARIANT=4, ILQG_FD_GW=204),
            dict(ILQG_FD_VARIANT=4, ILQG_FD_GW=802), dict(ILQG_FORCE_GENERIC=1)):
    h = handle("hopper", **env)
    d = torch.zeros((n, h.model.nd), dtype=torch.float64, device=dev)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    h.fd_batch_dev(q, v, u, w, d, None, st, cost=cost)
    torch.cuda.synchronize()
    print("hopper FD", env, "finite", bool(torch.isfinite(d).all()), "status", int(st.abs().sum()), flush=True)
    h.close()
qq, vv, ww = q.clone(), v.clone(), w.clone()
h0.step_batch_dev(qq, vv, u, ww, None, nsteps=3)
a = torch.zeros_like(v)
h0.forward_batch_dev(qq, vv, u, ww, a)
torch.cuda.synchronize()
print("hopper step/forward finite", bool(torch.isfinite(a).all()), flush=True)

hp = handle("inverted_pendulum")
ninst, N = 40, 20
x0 = wl.pendulum_initial_states(ninst, seed=1)
il = pkg.Ilqr(hp, ninst, N)
il.set_cost(pkg.make_cost(q2=[1, 10], v2=[1, 10], u2=[1]))
il.init_host(x0[0], x0[1], np.zeros((ninst, N, 1)))
il.iterate(3)
torch.cuda.synchronize()
print("pendulum iLQR iterations", il.iterations, flush=True)
il.close()

hh = handle("humanoid")
q, v, u, w, _ = wl.humanoid_states(hh, 6, seed=50)
d = torch.zeros((6, hh.model.nd), dtype=torch.float64, device=dev)
st = torch.zeros(6, dtype=torch.int32, device=dev)
hh.fd_batch_dev(q, v, u, w, d, None, st)
torch.cuda.synchronize()
print("humanoid FD finite", bool(torch.isfinite(d).all()), "status", int(st.abs().sum()), flush=True)
il = pkg.Ilqr(hh, 2, 4, alphas=(1.0, 0.5))
il.set_layout(1)
il.set_cost(pkg.make_cost(q2=[0.1] * 28, v2=[0.01] * 27, u2=[0.01] * 21))
il.init_host(q[:2].cpu().numpy(), v[:2].cpu().numpy(), np.zeros((2, 4, 21)))
il.iterate(2, accept_always=False)
torch.cuda.synchronize()
print("humanoid iLQR iterations", il.iterations, flush=True)
il.close()
print("done")
