"""Is the rank-to-rank spread of the 8-GPU pendulum iLQR line data or host?  The bench draws every rank's 4096 starts from its own seed
(100 + rank); this runs the eight seeds one after the other on ONE GPU and prints the time of a batch iteration for each.
    python tools/prof_ilqr_seeds.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, __graft_entry__ as e
pkg = e.load_package()
for r in range(8):
    res = bench.bench_ilqr(pkg, 0, 4096, 10, 5, 1, r, with_cpu=False)
    print(f"seed {100 + r}: {res['ms_per_batch_iteration']:.4f} ms per batch iteration, {res['diverged_instances']} diverged instances")
