"""One-off wide parity sweep against the CPU oracle (checker) on the benchmark generators' states: hopper (both FD kernel
variants) and humanoid.  Prints max / p99 block errors and the worst knots; written for profiles/r01_parity_sweep.md."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as e
pkg = e.load_package(); o = e.load_oracle()
from ilqg_mujoco_b200 import workload as wl


def block_err(d, r, nv, nu):
    nj = nv * (2 * nv + nu)
    out = []
    for lo, hi in ((0, nv * nv), (nv * nv, 2 * nv * nv), (2 * nv * nv, nj)):
        scale = np.maximum(1.0, np.abs(r[:, lo:hi]).max(axis=1))
        out.append(np.abs(d[:, lo:hi] - r[:, lo:hi]).max(axis=1) / scale)
    return np.max(out, axis=0)


for name, variant, ntraj, stride in (("hopper", "3", 4096, 4), ("hopper", "2", 4096, 16), ("humanoid", None, 0, 0)):
    if variant:
        os.environ["ILQG_FD_VARIANT"] = variant
    h = pkg.Handle(pkg.Model.named(name), 0)
    os.environ.pop("ILQG_FD_VARIANT", None)
    m = h.model; om = o.Model(os.path.join(pkg.MODELS_DIR, name + ".ilqgm"))
    if name == "hopper":
        q, v, u, w, _ = wl.make_knots(h, ntraj, 21, seed=0, device="cuda:0", model="hopper")
        idx = torch.arange(0, q.shape[0], stride, device="cuda:0")
        cost = pkg.make_cost(q1=[1.0])
    else:
        q, v, u, w, _ = wl.humanoid_states(h, 2048, seed=0, device="cuda:0")
        idx = torch.arange(0, q.shape[0], 4, device="cuda:0")
        cost = None
    n = q.shape[0]
    deriv = torch.zeros((n, m.nd), dtype=torch.float64, device="cuda:0"); st = torch.zeros(n, dtype=torch.int32, device="cuda:0")
    h.fd_batch_dev(q, v, u, w, deriv, None, st, cost=cost)
    sq, sv, su, sw = (t[idx].cpu().numpy() for t in (q, v, u, w))
    ref, _, _ = o.fd_batch(om, sq, sv, su, sw, cost, nthreads=0)
    err = block_err(deriv[idx].cpu().numpy(), ref, m.nv, m.nu)
    ok = st[idx].cpu().numpy() == 0
    worst = np.argsort(err)[-3:][::-1]
    print(f"{name} variant {variant or 'generic'}: {idx.numel()} knots of {n}; status ok {int(ok.sum())}; max rel block err {err[ok].max():.3e}, "
          f"p99 {np.quantile(err[ok], 0.99):.3e}, median {np.median(err[ok]):.3e}; knots above 1e-6: {int((err[ok] > 1e-6).sum())}; worst {[(int(idx[i]), float(err[i])) for i in worst]}")
    h.close()
