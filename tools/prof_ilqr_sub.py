"""Sweep of the iLQR workspace's sub-batch count (ILQG_ILQR_SUB) on the bench's two iLQR configurations.
    python tools/prof_ilqr_sub.py [subs, comma separated]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
import bench
subs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "2", "3", "4", "6", "8", "12", "default"]
for sub in subs:
    if sub == "default":
        os.environ.pop("ILQG_ILQR_SUB", None)
    else:
        os.environ["ILQG_ILQR_SUB"] = sub
    r = bench.bench_ilqr(pkg, 0, 4096, 10, 5, 1, 0, False)
    r2 = bench.bench_hopper_ilqr(pkg, 0, 1024, 10, 3, 1, 0)
    print(f"ILQG_ILQR_SUB={sub:8s} pendulum 4096: {r['ms_per_batch_iteration']:.4f} ms/batch iteration = {r['value'] / 1e6:.2f} M its/s (e2e {r['e2e']['value'] / 1e6:.2f} M); "
          f"hopper 1024 x 6 alphas: {r2['ms_per_batch_iteration']:.4f} ms = {r2['value'] / 1e6:.3f} M its/s", flush=True)
