"""Where the time of a small FD pass goes (BASELINE configs[4]: one hopper horizon of T = 1000 knots, and the T / 8 = 125 knots
a rank holds when the horizon is sharded over 8 GPUs): kernel times by CUDA events for a sweep of knot counts, the centre
evaluation's per-knot cycle counts and Newton iterations (ilqg_fd_set_diag), every kernel variant and the generic engine.

    python tools/prof_t1000.py [T]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as e

pkg = e.load_package()
from ilqg_mujoco_b200 import workload as wl

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = "cuda:0"
model = pkg.Model.named("hopper")


def horizon(h, T):
    """The nominal of bench.py's bench_t1000: config-2 initial state, smooth random control (3 sinusoids), rolled on the device."""
    q, v, u, w, _ = wl.make_knots(h, 1, T, seed=0, device=dev, model="hopper")
    tt = torch.arange(T, device=dev, dtype=torch.float64)[:, None] * 0.002
    gen = torch.Generator(device="cpu").manual_seed(0)
    amp = torch.rand((3, 3), generator=gen, dtype=torch.float64).to(dev) * 0.3
    frq = (torch.rand((3, 3), generator=gen, dtype=torch.float64) * 6 + 1).to(dev)
    u = sum(amp[i][None, :] * torch.sin(2 * np.pi * frq[i][None, :] * tt + i) for i in range(3)).contiguous()
    qs, vs, ws = q[:1].clone(), v[:1].clone(), w[:1].clone()
    Q, V, W = [], [], []
    for t in range(T):
        Q.append(qs.clone()); V.append(vs.clone()); W.append(ws.clone())
        h.step_batch_dev(qs, vs, u[t:t + 1].contiguous(), ws, None, nsteps=1)
    return torch.cat(Q), torch.cat(V), u, torch.cat(W)


def timed(h, q, v, u, w, reps=20, diag=None):
    """(per-kernel us with profiling events between the kernels, call us without them, deriv).  The events disable the
    programmatic dependent launch of the column kernel, so the two are measured in separate loops."""
    L = pkg.lib()
    n = q.shape[0]
    deriv = torch.zeros((n, model.nd), dtype=torch.float64, device=dev)
    if diag is not None:
        h.fd_set_diag(diag)
    L.ilqg_set_profiling(h._h, 0)
    for _ in range(3):
        h.fd_batch_dev(q, v, u, w, deriv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        h.fd_batch_dev(q, v, u, w, deriv)
    e1.record()
    e1.synchronize()
    wall = e0.elapsed_time(e1) / reps
    L.ilqg_set_profiling(h._h, 1)
    acc = np.zeros(3)
    for _ in range(reps):
        h.fd_batch_dev(q, v, u, w, deriv)
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        L.ilqg_fd_last_stage_ms(h._h, C.byref(a), C.byref(b), C.byref(c))
        acc += [a.value, b.value, c.value]
    L.ilqg_set_profiling(h._h, 0)
    h.fd_set_diag(None)
    return acc / reps * 1e3, wall * 1e3, deriv


def handle(**env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return pkg.Handle(model, 0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


h = handle()
q, v, u, w = horizon(h, T)
print(f"horizon of {T} knots built")
diag = torch.zeros((T, 8), dtype=torch.int32, device=dev)
k_us, wall_us, d_ref = timed(h, q, v, u, w, diag=diag)
d = diag.cpu().numpy()
print(f"[default engine={h.engine}] T={T}: centre {k_us[0]:.1f} us, velctrl {k_us[1]:.1f} us, qpos/perturb {k_us[2]:.1f} us, call {wall_us:.1f} us "
      f"-> {T / wall_us:.2f} M knots/s")
print("  nefc histogram:", dict(zip(*np.unique(d[:, 0], return_counts=True))))
print("  first-solve iterations:", dict(zip(*np.unique(d[:, 1], return_counts=True))))
print("  all-warm-up iterations:", dict(zip(*np.unique(d[:, 2], return_counts=True))))
for lab, sel in (("flight", d[:, 0] == 0), ("stance", d[:, 0] > 0)):
    if sel.any():
        print(f"  {lab}: {int(sel.sum())} knots; build cycles median {np.median(d[sel, 4]):.0f} max {d[sel, 4].max()}; "
              f"solve cycles median {np.median(d[sel, 5]):.0f} max {d[sel, 5].max()}")
print("  (cycles at ~1.9 GHz: 1000 cycles = 0.52 us)")

print("-- knot-count sweep (prefix of the horizon, tiled when longer): call us / M knots/s per kernel variant")
variants = (("centre+columns PDL", dict(ILQG_FD_VARIANT=2)), ("auto", {}),
            ("build+group solve 4 lanes/128r", dict(ILQG_FD_VARIANT=4, ILQG_FD_GW=404)), ("4 lanes/255r", dict(ILQG_FD_VARIANT=4, ILQG_FD_GW=204)),
            ("4 lanes/80r", dict(ILQG_FD_VARIANT=4, ILQG_FD_GW=604)), ("2 lanes/255r", dict(ILQG_FD_VARIANT=4, ILQG_FD_GW=402)),
            ("2 lanes/128r", dict(ILQG_FD_VARIANT=4, ILQG_FD_GW=802)), ("2 lanes/80r", dict(ILQG_FD_VARIANT=4, ILQG_FD_GW=1202)),
            ("1-launch, warp solves centre", dict(ILQG_FD_VARIANT=1, ILQG_FD_COOP=1)), ("1-launch, centre lane alone", dict(ILQG_FD_VARIANT=1)),
            ("centre+columns no PDL", dict(ILQG_FD_VARIANT=2, ILQG_FD_PDL=0)), ("split", dict(ILQG_FD_VARIANT=3)))
if os.environ.get("PROF_VARIANTS"):
    keep = [int(x) for x in os.environ["PROF_VARIANTS"].split(",")]
    variants = tuple(variants[i] for i in keep)
hs = [(lab, handle(**env)) for lab, env in variants]
print("  (third figure: max |difference| to the first variant's deriv blocks / max |block|)")
print("  n      " + "".join(f"{lab:>30s}" for lab, _ in hs))
for n in (21, 32, 125, 250, 500, 1000, 2000, 3000, 4000, 6000, 8000, 16000, 24000):
    reps = (n + T - 1) // T
    qq, vv, uu, ww = (x.repeat(reps, 1)[:n].contiguous() for x in (q, v, u, w))
    row = f"  {n:6d} "
    ref = None
    for lab, hh in hs:
        k_us, wall_us, dd = timed(hh, qq, vv, uu, ww)
        if ref is None:
            ref = dd
        err = float((dd[:, :90] - ref[:, :90]).abs().max() / ref[:, :90].abs().max())
        row += f"{wall_us:11.1f} us {n / wall_us:6.2f} M {err:7.0e}"
    print(row)
for _, hh in hs:
    hh.close()
hh = handle(ILQG_FORCE_GENERIC=1)
for n in (125, 1000):
    k_us, wall_us, dd = timed(hh, q[:n].contiguous(), v[:n].contiguous(), u[:n].contiguous(), w[:n].contiguous())
    err = float((dd[:, :90] - d_ref[:n, :90]).abs().max() / d_ref[:n, :90].abs().max())
    print(f"[generic warp-per-rollout engine] n={n}: centre {k_us[0]:.1f} us, velctrl {k_us[1]:.1f} us, qpos {k_us[2]:.1f} us, call {wall_us:.1f} us "
          f"-> {n / wall_us:.2f} M knots/s; max rel diff to default {err:.2e}")
hh.close()
h.close()
