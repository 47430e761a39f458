"""What pipelining the phases of the batched iLQR iteration over sub-batches would give: the 4096 pendulum problems of BASELINE configs[3]
as S independent (handle, workspace) pairs of 4096 / S problems on S streams.  Every phase of an iteration is one dependent chain
(rollout: 84 evaluations on one warp per scheduler), so sub-batches in different phases overlap.   python tools/prof_ilqr_split.py [ninst]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
ninst = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
niter, reps = 10, 10
dev = "cuda:0"
model = pkg.Model.named("inverted_pendulum")
cost = pkg.make_cost(q2=[1, 10], v2=[1, 10], u2=[1])
q, v, u, _ = wl.pendulum_initial_states(ninst, seed=100)
u = u * 0.0
Slist = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (1, 2, 3, 4, 6, 8, 16)
for S in Slist:
    ni = ninst // S
    hs, ils, sts, args = [], [], [], []
    for s in range(S):
        h = pkg.Handle(model, 0)
        il = pkg.Ilqr(h, ni, 20, (1.0,))
        il.set_cost(cost)
        st = torch.cuda.Stream()
        sl = slice(s * ni, (s + 1) * ni)
        a = [torch.from_numpy(np.ascontiguousarray(x[sl])).to(dev) for x in (q, v, u)] + [torch.zeros((ni, 2), dtype=torch.float64, device=dev)]
        hs.append(h); ils.append(il); sts.append(st); args.append(a)
    times = []
    for r in range(reps + 3):
        for il, st, a in zip(ils, sts, args):
            il.init_dev(*a, stream=st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in sts:
            st.wait_stream(torch.cuda.current_stream())
        for il, st in zip(ils, sts):
            il.iterate(niter, accept_always=True, stream=st.cuda_stream)
        for st in sts:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        e1.synchronize()
        if r >= 3:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times)) / niter
    print(f"ILQG_ILQR_SUB={os.environ.get('ILQG_ILQR_SUB', 'default')} S={S:2d} workspaces of {ni}: {ms:.4f} ms per batch iteration -> {S * ni / ms / 1e3:.2f} M iterations/s", flush=True)
    for il in ils:
        il.close()
    for h in hs:
        h.close()
