"""End-to-end time of an MPC step of the 4096-pendulum iLQR line through the host-pointer calls, phase by phase, and what
continuing an optimisation (setDInit alone) does to the cost of a step.   python tools/prof_ilqr_e2e.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
model = pkg.Model.named("inverted_pendulum")
h = pkg.Handle(model, 0)
ninst, niter = 4096, 10
q, v, u, _ = wl.pendulum_initial_states(ninst, seed=100)
u = u * 0.0
cost = pkg.make_cost(q2=[1, 10], v2=[1, 10], u2=[1])
il = pkg.Ilqr(h, ninst, 20, (1.0,))
il.set_cost(cost)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    il.init_host(q, v, u, None); il.iterate(niter, accept_always=True, stream=stream); il.fetch_controls(last=niter)
torch.cuda.synchronize()
T = np.zeros(4)
R = 20
for _ in range(R):
    t0 = time.perf_counter(); il.init_host(q, v, u, None); torch.cuda.synchronize()
    t1 = time.perf_counter(); il.iterate(niter, accept_always=True, stream=stream)
    t2 = time.perf_counter(); torch.cuda.synchronize()
    t3 = time.perf_counter(); il.fetch_controls(last=niter)
    t4 = time.perf_counter()
    T += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
print("fresh problems every step, ms: init_host %.3f, iterate enqueue %.3f, iterate wait %.3f, fetch %.3f" % tuple(T / R * 1e3))
ts = []
for _ in range(12):
    t0 = time.perf_counter()
    il.init_host(q, v, u, None); il.iterate(niter, accept_always=True, stream=stream); il.fetch_controls(last=niter)
    ts.append((time.perf_counter() - t0) * 1e3)
print("fresh problems every step, per-step ms:", " ".join(f"{t:.2f}" for t in ts))
ts = []
il.init_host(q, v, u, None)
for _ in range(12):
    t0 = time.perf_counter()
    il.set_state_host(q, v); il.iterate(niter, accept_always=True, stream=stream); il.fetch_controls(last=niter)
    ts.append((time.perf_counter() - t0) * 1e3)
print("setDInit alone (the optimisation continues), per-step ms:", " ".join(f"{t:.2f}" for t in ts))
il.close(); h.close()
import bench
for reps in (3, 10):
    r = bench.bench_ilqr(pkg, 0, 4096, 10, reps, 1, 0, with_cpu=False)
    print(f"bench_ilqr reps={reps}: device {r['value']/1e6:.2f} M, e2e {r['e2e']['value']/1e6:.2f} M its/s")
