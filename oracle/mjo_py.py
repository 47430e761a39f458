"""ctypes view of the CPU ORACLE (oracle/libmjo.so).  Test infrastructure only:
import it from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MODEL_BYTES = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmjo.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libmjo.so missing: run `make -C oracle` (or __graft_entry__.build())")
        _LIB = C.CDLL(path)
        _LIB.mjo_make_data.restype = C.c_void_p
        _LIB.mjo_energy.restype = C.c_double
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Model:
    """Opaque ilqg_model table loaded from a .ilqgm file, plus the few sizes tests need."""

    def __init__(self, path):
        self.buf = np.fromfile(path, dtype=np.uint8)
        ints = self.buf[:40].view(np.int32)
        assert ints[0] == 0x494C5147, "bad model magic"
        self.nq, self.nv, self.nu, self.nbody, self.njnt, self.ngeom, self.npair = (int(x) for x in ints[2:9])
        self.nd = self.nv * (2 * self.nv + self.nu) + 2 * self.nv + self.nu
        self.timestep = float(self.buf[40:48].view(np.float64)[0])

    @property
    def ptr(self):
        return self.buf.ctypes.data_as(C.c_void_p)


def fd_batch(m, qpos, qvel, ctrl, warm, cost=None, eps=1e-6, niter=30, nwarmup=3, nthreads=0):
    n = qpos.shape[0]
    qpos = np.ascontiguousarray(qpos, np.float64); qvel = np.ascontiguousarray(qvel, np.float64)
    ctrl = np.ascontiguousarray(ctrl, np.float64); warm = np.ascontiguousarray(warm, np.float64)
    deriv = np.zeros((n, m.nd)); qacc = np.zeros((n, m.nv)); fl = C.c_double(0)
    lib().mjo_fd_batch_quad(m.ptr, n, _p(qpos), _p(qvel), _p(ctrl), _p(warm), _p(cost), C.c_double(eps), niter, nwarmup,
                            _p(deriv), _p(qacc), nthreads, C.byref(fl))
    return deriv, qacc, fl.value


def step_batch(m, qpos, qvel, ctrl, warm, nsteps, nthreads=0):
    qpos = np.array(qpos, np.float64, order="C"); qvel = np.array(qvel, np.float64, order="C")
    ctrl = np.ascontiguousarray(ctrl, np.float64); warm = np.array(warm, np.float64, order="C")
    n = qpos.shape[0]
    qacc = np.zeros((n, m.nv))
    lib().mjo_step_batch(m.ptr, n, nsteps, _p(qpos), _p(qvel), _p(ctrl), _p(warm), _p(qacc), nthreads)
    return qpos, qvel, warm, qacc


def forward_batch(m, qpos, qvel, ctrl, warm, nthreads=0):
    qpos = np.ascontiguousarray(qpos, np.float64); qvel = np.ascontiguousarray(qvel, np.float64)
    ctrl = np.ascontiguousarray(ctrl, np.float64); warm = np.array(warm, np.float64, order="C")
    n = qpos.shape[0]
    qacc = np.zeros((n, m.nv))
    lib().mjo_forward_batch(m.ptr, n, _p(qpos), _p(qvel), _p(ctrl), _p(warm), _p(qacc), nthreads)
    return qacc, warm


def make_cost(q2=(), q1=(), v2=(), v1=(), u2=(), u1=()):
    """ilqg_cost struct as a float64 array: q2[32] q1[32] v2[32] v1[32] u2[24] u1[24]."""
    c = np.zeros(32 * 4 + 24 * 2)
    for off, v in ((0, q2), (32, q1), (64, v2), (96, v1), (128, u2), (152, u1)):
        c[off:off + len(v)] = v
    return c


def ilqr_run_batch(m, N, niter, qpos0, qvel0, ctrl0, warm0, cost, alphas=None, accept_always=True, mu=1000.0, corrected=False, mu_schedule=None):
    """niter x iterate for every instance (reference cadence); alphas=None -> the reference's own iterate()."""
    n = qpos0.shape[0]
    T, nx = N + 1, 2 * m.nv
    q0, v0, u0 = (np.ascontiguousarray(a, np.float64) for a in (qpos0, qvel0, ctrl0))
    w0 = np.ascontiguousarray(warm0, np.float64) if warm0 is not None else None
    al = np.ascontiguousarray(alphas, np.float64) if alphas is not None else None
    out = dict(J=np.zeros((n, niter)), accepted=np.zeros((n, niter), np.int32), qpos=np.zeros((n, T, m.nq)), qvel=np.zeros((n, T, m.nv)),
               ctrl=np.zeros((n, T, m.nu)), K=np.zeros((n, T, m.nu * nx)), k=np.zeros((n, T, m.nu)), V=np.zeros((n, nx * nx)), v=np.zeros((n, nx)))
    lib().mjo_ilqr_set_corrected_layout(1 if corrected else 0)
    f, lo, hi = mu_schedule if mu_schedule else (1.0, 1e-6, 1e10)
    lib().mjo_ilqr_set_mu_schedule(C.c_double(f), C.c_double(lo), C.c_double(hi))
    lib().mjo_ilqr_run_batch(m.ptr, n, N, niter, _p(q0), _p(v0), _p(u0), _p(w0), _p(cost), _p(al), 0 if al is None else len(al),
                             1 if accept_always else 0, C.c_double(mu), _p(out["J"]), _p(out["accepted"]), _p(out["qpos"]), _p(out["qvel"]),
                             _p(out["ctrl"]), _p(out["K"]), _p(out["k"]), _p(out["V"]), _p(out["v"]), 0)
    lib().mjo_ilqr_set_corrected_layout(0)
    lib().mjo_ilqr_set_mu_schedule(C.c_double(1.0), C.c_double(1e-6), C.c_double(1e10))
    return out


def dump(m, qpos, qvel, ctrl, warm=None, iterations=30, tolerance=0.0):
    """Every intermediate of one forward evaluation, under mjData's names (mjo_debug_dump): the quantities the MuJoCo
    cross-check (tools/mujoco_fixtures.py) and the hand-derived contact anchors compare."""
    nv, nb = m.nv, m.nbody
    out = dict(xpos=np.zeros((nb, 3)), xquat=np.zeros((nb, 4)), xipos=np.zeros((nb, 3)), subtree_com=np.zeros((nb, 3)), cinert=np.zeros((nb, 10)),
               cdof=np.zeros((nv, 6)), qM=np.zeros((nv, nv)), qfrc_bias=np.zeros(nv), qfrc_passive=np.zeros(nv), qfrc_actuator=np.zeros(nv),
               qacc_smooth=np.zeros(nv), qacc=np.zeros(nv))
    counts = np.zeros(3, np.int32)
    con = np.zeros((96, 13)); con_geom = np.zeros((96, 2), np.int32); efc_J = np.zeros((320, nv)); efc = np.zeros((320, 7))
    q, v, u = (np.ascontiguousarray(a, np.float64) for a in (qpos, qvel, ctrl))
    w = np.ascontiguousarray(warm, np.float64) if warm is not None else None
    lib().mjo_debug_dump(m.ptr, _p(q), _p(v), _p(u), _p(w), int(iterations), C.c_double(tolerance), *[_p(out[k]) for k in
                         ("xpos", "xquat", "xipos", "subtree_com", "cinert", "cdof", "qM", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qacc_smooth", "qacc")],
                         counts.ctypes.data_as(C.c_void_p), _p(con), con_geom.ctypes.data_as(C.c_void_p), _p(efc_J), _p(efc))
    ncon, nefc = int(counts[0]), int(counts[1])
    out.update(ncon=ncon, nefc=nefc, solver_iter=int(counts[2]), contact_dist=con[:ncon, 0].copy(), contact_pos=con[:ncon, 1:4].copy(),
               contact_frame=con[:ncon, 4:13].copy(), contact_geom=con_geom[:ncon].copy(), efc_J=efc_J[:nefc].copy(), efc_pos=efc[:nefc, 0].copy(),
               efc_margin=efc[:nefc, 1].copy(), efc_diagApprox=efc[:nefc, 2].copy(), efc_R=efc[:nefc, 3].copy(), efc_D=efc[:nefc, 4].copy(),
               efc_aref=efc[:nefc, 5].copy(), efc_force=efc[:nefc, 6].copy())
    return out
