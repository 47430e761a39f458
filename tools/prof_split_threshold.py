"""Where the stage-skipping split (variant 3) overtakes the centre + single column kernel pair (variant 2) on the SURVEY-8d hopper batch
and on a stance-heavy one: time per FD pass for a sweep of batch sizes.   python tools/prof_split_threshold.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
model = pkg.Model.named("hopper")
h0 = pkg.Handle(model, 0)
cost = pkg.make_cost(q1=[1.0])
Q = {"8d": wl.make_knots_8d(h0, 1400, 21, seed=0, device="cuda:0")[:4]}
qs, vs, us, ws, _ = wl.make_knots(h0, 1, 1000, seed=0, device="cuda:0", model="hopper")
Q["stance"] = tuple(x.repeat(30, 1).contiguous() for x in (qs, vs, us, ws))
h0.close()
for name, (q, v, u, w) in Q.items():
    for n in (8192, 12288, 16384, 20480, 21504, 24576, 28672):
        row = []
        for var in ("2", "3"):
            os.environ["ILQG_FD_VARIANT"] = var
            h = pkg.Handle(model, 0)
            del os.environ["ILQG_FD_VARIANT"]
            a = [x[:n].contiguous() for x in (q, v, u, w)]
            d = torch.zeros((n, model.nd), dtype=torch.float64, device="cuda:0")
            for _ in range(5):
                h.fd_batch_dev(*a, d, cost=cost)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                h.fd_batch_dev(*a, d, cost=cost)
            e1.record(); e1.synchronize()
            row.append(e0.elapsed_time(e1) / 30 * 1e3)
            h.close()
        print(f"{name:7s} n={n:6d}: centre + columns {row[0]:7.1f} us, split {row[1]:7.1f} us")
