"""Host-pointer FD call (ilqg_fd_batch_host) timing for a few chunkings of its copy/compute pipeline.

    python tools/prof_e2e.py [chunks[:compute_streams[:priority_ladder]] ...]      (- = the default of that setting)
ILQG_WORKLOAD=r1: round 1's batch instead of the SURVEY-8d one."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl
model = pkg.Model.named("hopper")
h = pkg.Handle(model, 0)
if os.environ.get("ILQG_WORKLOAD") == "r1":
    q, v, u, w, _ = wl.make_knots(h, 4096, 21, seed=0, device="cuda:0", model="hopper")
else:
    q, v, u, w, _ = wl.make_knots_8d(h, 4096, 21, seed=0, device="cuda:0")
h.close()
nk = q.shape[0]
hq, hv, hu, hw = (t.cpu().pin_memory() for t in (q, v, u, w))
hd = torch.zeros((nk, model.nd), dtype=torch.float64).pin_memory()
ha = torch.zeros((nk, model.nv), dtype=torch.float64).pin_memory()
hs = torch.zeros(nk, dtype=torch.int32).pin_memory()
cost = pkg.make_cost(q1=[1.0]); L = pkg.lib()
def step():
    L.ilqg_fd_batch_host(h._h, nk, C.c_void_p(hq.data_ptr()), C.c_void_p(hv.data_ptr()), C.c_void_p(hu.data_ptr()), C.c_void_p(hw.data_ptr()),
                         cost.ctypes.data_as(C.c_void_p), None, C.c_void_p(hd.data_ptr()), C.c_void_p(ha.data_ptr()), C.c_void_p(hs.data_ptr()))
for spec in sys.argv[1:] or ["-"]:
    f = spec.split(":") + ["-", "-"]
    ch = f[0]
    for key, val in (("ILQG_HOST_CHUNKS", f[0]), ("ILQG_HOST_COMP", f[1]), ("ILQG_HOST_PRIO", f[2])):   # read when the handle is created
        os.environ.pop(key, None)
        if val != "-": os.environ[key] = val
    h = pkg.Handle(model, 0)
    for _ in range(3): step()
    t0 = time.perf_counter()
    for _ in range(30): step()
    dt = (time.perf_counter() - t0) / 30
    h.close()
    print(f"chunks:streams:ladder={spec}: {dt*1e3:.3f} ms/step -> {nk/dt/1e6:.2f} M knots/s e2e; D2H {nk*(model.nd+model.nv)*8/dt/1e9:.1f} GB/s")
# the same call with PAGEABLE caller buffers (numpy / malloc'ed memory, what a calcMJDerivatives-shaped caller has): the call's own
# pinned mirror + copy threads (ILQG_HOST_THREADS; 0 = leave the staging to the driver)
pq, pv, pu, pw = (t.cpu().numpy().copy() for t in (q, v, u, w))
pd = np.zeros((nk, model.nd)); pa = np.zeros((nk, model.nv)); ps = np.zeros(nk, dtype=np.int32)
def pstep():
    L.ilqg_fd_batch_host(h._h, nk, pq.ctypes.data_as(C.c_void_p), pv.ctypes.data_as(C.c_void_p), pu.ctypes.data_as(C.c_void_p), pw.ctypes.data_as(C.c_void_p),
                         cost.ctypes.data_as(C.c_void_p), None, pd.ctypes.data_as(C.c_void_p), pa.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p))
for key in ("ILQG_HOST_CHUNKS", "ILQG_HOST_COMP", "ILQG_HOST_PRIO"):
    os.environ.pop(key, None)
for th in os.environ.get("ILQG_E2E_THREADS", "-1,0,2,4,8,16").split(","):
    os.environ["ILQG_HOST_THREADS"] = th
    h = pkg.Handle(model, 0)
    for _ in range(3): pstep()
    t0 = time.perf_counter()
    for _ in range(20): pstep()
    dt = (time.perf_counter() - t0) / 20
    h.close()
    same = bool(np.array_equal(pd, hd.numpy()))
    print(f"pageable buffers, copy threads={th}: {dt*1e3:.3f} ms/step -> {nk/dt/1e6:.2f} M knots/s e2e; equals the pinned call's deriv: {same}")
os.environ.pop("ILQG_HOST_THREADS", None)
# raw PCIe rates for reference
d = torch.empty(nk * model.nd, dtype=torch.float64, device="cuda:0")
for name, fn in (("D2H", lambda: hd.view(-1).copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(hd.view(-1), non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"raw {name} of the deriv array: {nk*model.nd*8/dt/1e9:.1f} GB/s")
