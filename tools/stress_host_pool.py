"""Stress of the host-pointer FD call's copy pool (BounceCrew / CopyPool, ilqg.cu): handles created and destroyed with different thread
counts, batch sizes on both sides of the 4 MB threshold, pageable / pinned / mixed caller buffers in random order, every result compared
bit for bit with one device call over the same knots.   python tools/stress_host_pool.py [calls]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 300
model = pkg.Model.named("hopper")
os.environ["ILQG_FD_VARIANT"] = "3"      # one kernel variant for every size: bit-identical results to compare
h = pkg.Handle(model, 0)
q, v, u, w, _ = wl.make_knots_8d(h, 2048, 21, seed=7, device="cuda:0")
N = q.shape[0]
cost = pkg.make_cost(q1=[1.0])
ref = torch.zeros((N, model.nd), dtype=torch.float64, device="cuda:0")
h.fd_batch_dev(q, v, u, w, ref, cost=cost)
torch.cuda.synchronize()
ref = ref.cpu().numpy()
hq, hv, hu, hw = (t.cpu() for t in (q, v, u, w))
pq, pv, pu, pw = (t.pin_memory() for t in (hq, hv, hu, hw))
rng = np.random.default_rng(0)
t0 = time.time()
bad = 0
for c in range(calls):
    if c % 25 == 0:
        h.close()
        os.environ["ILQG_HOST_THREADS"] = str(rng.choice([-1, 0, 1, 2, 3, 5, 8, 16]))
        os.environ["ILQG_HOST_CHUNKS"] = str(rng.choice([0, 1, 3, 8, 13]))
        h = pkg.Handle(model, 0)
    n = int(rng.choice([1, 100, 4000, 6000, 9000, 20000, N]))
    lo = int(rng.integers(0, N - n + 1))
    kind = rng.integers(0, 3)
    src = [(a[lo:lo + n].numpy() if (kind == 0 or (kind == 2 and i % 2)) else b[lo:lo + n].numpy()) for i, (a, b) in enumerate(((hq, pq), (hv, pv), (hu, pu), (hw, pw)))]
    d, a, st = h.fd_batch_host(*src, cost if rng.uniform() < 0.8 else None)
    njac = model.nv * (2 * model.nv + model.nu)
    if st.sum() != 0 or not np.array_equal(d[:, :njac], ref[lo:lo + n, :njac]):
        bad += 1
        print("MISMATCH at call", c, "n", n, "kind", kind, flush=True)
h.close()
print(f"{calls} calls in {time.time() - t0:.1f} s, {bad} mismatches")
sys.exit(1 if bad else 0)
