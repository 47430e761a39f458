"""ilqg-mujoco_b200 — Python view (ctypes) of the C ABI in include/ilqg_b200.h.

The product is the CUDA library `libilqg_b200.so` (sources in csrc/, built by
`__graft_entry__.build()`); this module only marshals pointers.  There is no CPU
implementation here: if the library is missing, or no CUDA device is present, calls raise.
The host-language mirror of the reference's classes (calcMJDerivatives / Differentiator /
ILQR / InvertedPendulum) is C++ and lives in host/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ILQG_LIB") or os.path.join(_HERE, "libilqg_b200.so")   # ILQG_LIB: A/B-test an alternative build
MODELS_DIR = os.path.join(_HERE, "models")

OK, ERR_ARG, ERR_MODEL, ERR_IO, ERR_CUDA, ERR_UNSUPPORTED, ERR_NONFINITE, ERR_CAPACITY = range(8)
MAXQ, MAXV, MAXU = 32, 32, 24

_lib = None


class IlqgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ilqg error {code}: {msg}")
        self.code = code


def lib():
    """Load libilqg_b200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the CUDA extension is the product; there is no fallback)")
        L = C.CDLL(LIB_PATH)
        L.ilqg_last_error.restype = C.c_char_p
        L.ilqg_last_error.argtypes = [C.c_void_p]
        L.ilqg_engine_name.restype = C.c_char_p
        L.ilqg_engine_name.argtypes = [C.c_void_p]
        L.ilqg_launch_count.restype = C.c_long
        L.ilqg_launch_count.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class FdOpts(C.Structure):
    _fields_ = [("eps", C.c_double), ("niter", C.c_int), ("nwarmup", C.c_int)]


def make_cost(q2=(), q1=(), v2=(), v1=(), u2=(), u1=()):
    """struct ilqg_cost as a float64 array: q2[32] q1[32] v2[32] v1[32] u2[24] u1[24]."""
    c = np.zeros(4 * MAXQ + 2 * MAXU)
    for off, v in ((0, q2), (32, q1), (64, v2), (96, v1), (128, u2), (152, u1)):
        c[off:off + len(v)] = v
    return c


class Model:
    """The flat `ilqg_model` tables (opaque bytes) plus the sizes Python needs."""

    def __init__(self, buf):
        self.buf = np.ascontiguousarray(buf, dtype=np.uint8)
        if self.buf.size != lib().ilqg_model_sizeof():
            raise ValueError(f"not an ilqg model table: {self.buf.size} bytes, expected {lib().ilqg_model_sizeof()}")
        ints = self.buf[:40].view(np.int32)
        if ints[0] != 0x494C5147:
            raise ValueError("not an ilqg model table")
        self.nq, self.nv, self.nu, self.nbody, self.njnt, self.ngeom, self.npair = (int(x) for x in ints[2:9])
        self.nd = self.nv * (2 * self.nv + self.nu) + 2 * self.nv + self.nu
        self.timestep = float(self.buf[40:48].view(np.float64)[0])

    def validate(self):
        """Structural check of the table (ilqg_model_validate): raises IlqgError with the offending field."""
        err = C.create_string_buffer(256)
        rc = lib().ilqg_model_validate(self.ptr, err, 256)
        if rc:
            raise IlqgError(rc, err.value.decode())

    @classmethod
    def load(cls, path):
        return cls(np.fromfile(path, dtype=np.uint8))

    @classmethod
    def named(cls, name):
        """One of the compiled copies of the reference's models (inverted_pendulum, hopper, humanoid)."""
        return cls.load(os.path.join(MODELS_DIR, name + ".ilqgm"))

    @classmethod
    def from_mjcf(cls, path):
        n = lib().ilqg_model_sizeof()
        buf = np.zeros(n, dtype=np.uint8)
        err = C.create_string_buffer(512)
        rc = lib().ilqg_compile_mjcf(path.encode(), buf.ctypes.data_as(C.c_void_p), err, 512)
        if rc:
            raise IlqgError(rc, err.value.decode())
        return cls(buf)

    @property
    def ptr(self):
        return self.buf.ctypes.data_as(C.c_void_p)

    def field(self, name):
        """Writable numpy view of a named table field (see ilqg_model_field)."""
        off, cnt, dbl = C.c_int(), C.c_int(), C.c_int()
        rc = lib().ilqg_model_field(name.encode(), C.byref(off), C.byref(cnt), C.byref(dbl))
        if rc:
            raise KeyError(name)
        n = cnt.value * (8 if dbl.value else 4)
        return self.buf[off.value:off.value + n].view(np.float64 if dbl.value else np.int32)

    def copy(self):
        return Model(self.buf.copy())


def _hp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _dp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Handle:
    """GPU-resident compiled model (ilqg_create)."""

    def __init__(self, model, device=0):
        self.model = model
        self._h = C.c_void_p()
        rc = lib().ilqg_create(model.ptr, int(device), C.byref(self._h))
        if rc:
            raise IlqgError(rc, lib().ilqg_last_error(None).decode())

    def close(self):
        if self._h:
            lib().ilqg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise IlqgError(rc, lib().ilqg_last_error(self._h).decode())

    @property
    def engine(self):
        return lib().ilqg_engine_name(self._h).decode()

    @property
    def launches(self):
        return int(lib().ilqg_launch_count(self._h))

    # ---- host-pointer flavour (numpy) ------------------------------------------------
    def fd_batch_host(self, qpos, qvel, ctrl, warm=None, cost=None, opts=None, deriv=None):
        m = self.model
        qpos = np.ascontiguousarray(qpos, np.float64).reshape(-1, m.nq)
        n = qpos.shape[0]
        qvel = np.ascontiguousarray(qvel, np.float64).reshape(n, m.nv)
        ctrl = np.ascontiguousarray(ctrl, np.float64).reshape(n, m.nu)
        warm = None if warm is None else np.ascontiguousarray(warm, np.float64).reshape(n, m.nv)
        if deriv is None:
            deriv = np.zeros((n, m.nd))
        qacc = np.zeros((n, m.nv))
        status = np.zeros(n, np.int32)
        rc = lib().ilqg_fd_batch_host(self._h, n, _hp(qpos), _hp(qvel), _hp(ctrl), _hp(warm), _hp(cost),
                                      C.byref(opts) if opts is not None else None, _hp(deriv), _hp(qacc), _hp(status))
        if rc and rc != ERR_NONFINITE:
            self._check(rc)
        return deriv, qacc, status

    def forward_batch_host(self, qpos, qvel, ctrl, warm=None):
        m = self.model
        qpos = np.ascontiguousarray(qpos, np.float64).reshape(-1, m.nq)
        n = qpos.shape[0]
        qvel = np.ascontiguousarray(qvel, np.float64).reshape(n, m.nv)
        ctrl = np.ascontiguousarray(ctrl, np.float64).reshape(n, m.nu)
        warm = np.zeros((n, m.nv)) if warm is None else np.array(warm, np.float64, order="C").reshape(n, m.nv)
        qacc = np.zeros((n, m.nv))
        self._check(lib().ilqg_forward_batch_host(self._h, n, _hp(qpos), _hp(qvel), _hp(ctrl), _hp(warm), _hp(qacc)))
        return qacc, warm

    def step_batch_host(self, qpos, qvel, ctrl, warm=None, nsteps=1):
        m = self.model
        qpos = np.array(qpos, np.float64, order="C").reshape(-1, m.nq)
        n = qpos.shape[0]
        qvel = np.array(qvel, np.float64, order="C").reshape(n, m.nv)
        ctrl = np.ascontiguousarray(ctrl, np.float64).reshape(n, m.nu)
        warm = np.zeros((n, m.nv)) if warm is None else np.array(warm, np.float64, order="C").reshape(n, m.nv)
        qacc = np.zeros((n, m.nv))
        self._check(lib().ilqg_step_batch_host(self._h, n, int(nsteps), _hp(qpos), _hp(qvel), _hp(ctrl), _hp(warm), _hp(qacc)))
        return qpos, qvel, warm, qacc

    # ---- device-pointer flavour (torch CUDA tensors, float64, contiguous) ---------------
    def fd_batch_dev(self, qpos, qvel, ctrl, warm, deriv, qacc=None, status=None, cost=None, opts=None, stream=None):
        n = qpos.shape[0]
        self._check(lib().ilqg_fd_batch_dev(self._h, n, _dp(qpos), _dp(qvel), _dp(ctrl), _dp(warm), _hp(cost),
                                            C.byref(opts) if opts is not None else None, _dp(deriv), _dp(qacc), _dp(status),
                                            C.c_void_p(stream) if stream else None))

    def fd_batch_dev_scatter(self, qpos, qvel, ctrl, warm, dst_ptrs, qacc=None, status=None, cost=None, opts=None, stream=None):
        """FD of a knot range with the deriv blocks stored to every pointer of `dst_ptrs` (ints: device addresses of this
        range's first block in each destination, local or peer-mapped) — ilqg_fd_batch_dev_scatter."""
        n = qpos.shape[0]
        arr = (C.c_void_p * len(dst_ptrs))(*[C.c_void_p(int(p)) for p in dst_ptrs])
        self._check(lib().ilqg_fd_batch_dev_scatter(self._h, n, _dp(qpos), _dp(qvel), _dp(ctrl), _dp(warm), _hp(cost),
                                                    C.byref(opts) if opts is not None else None, arr, len(dst_ptrs), _dp(qacc), _dp(status),
                                                    C.c_void_p(stream) if stream else None))

    # ---- peer-visible buffers (CUDA IPC) ---------------------------------------------------
    def peer_alloc(self, nbytes):
        ptr, hd = C.c_void_p(), (C.c_ubyte * 64)()
        self._check(lib().ilqg_peer_alloc(self._h, C.c_size_t(nbytes), C.byref(ptr), hd))
        return int(ptr.value), bytes(hd)

    def peer_open(self, handle_bytes):
        ptr = C.c_void_p()
        hd = (C.c_ubyte * 64).from_buffer_copy(handle_bytes)
        self._check(lib().ilqg_peer_open(self._h, hd, C.byref(ptr)))
        return int(ptr.value)

    def peer_barrier(self, flag_ptrs, rank, epoch, stream=None):
        arr = (C.c_void_p * len(flag_ptrs))(*[C.c_void_p(int(p)) for p in flag_ptrs])
        self._check(lib().ilqg_peer_barrier(self._h, arr, len(flag_ptrs), int(rank), int(epoch), C.c_void_p(stream) if stream else None))

    def fd_set_diag(self, diag):
        """Per-knot diagnostics of the centre evaluation of the following fd_batch_dev calls: int32 CUDA tensor [nknots, 8]
        (columns: nefc, iterations of the first solve, of all warm-up solves, active rows, cycles build, cycles solve) or None."""
        self._check(lib().ilqg_fd_set_diag(self._h, _dp(diag)))

    def peer_barrier_timed_out(self):
        return bool(lib().ilqg_peer_barrier_timed_out(self._h))

    def peer_close(self, ptr):
        self._check(lib().ilqg_peer_close(self._h, C.c_void_p(ptr)))

    def peer_free(self, ptr):
        self._check(lib().ilqg_peer_free(self._h, C.c_void_p(ptr)))

    def step_batch_dev(self, qpos, qvel, ctrl, warm, qacc=None, nsteps=1, stream=None):
        n = qpos.shape[0]
        self._check(lib().ilqg_step_batch_dev(self._h, n, int(nsteps), _dp(qpos), _dp(qvel), _dp(ctrl), _dp(warm), _dp(qacc),
                                              C.c_void_p(stream) if stream else None))

    def forward_batch_dev(self, qpos, qvel, ctrl, warm, qacc, stream=None):
        n = qpos.shape[0]
        self._check(lib().ilqg_forward_batch_dev(self._h, n, _dp(qpos), _dp(qvel), _dp(ctrl), _dp(warm), _dp(qacc),
                                                 C.c_void_p(stream) if stream else None))


class Ilqr:
    """Batched iLQR workspace on the GPU (ilqg_ilqr_*): the role of `ILQR<nv,nu,N>` for `ninst` problems."""

    def __init__(self, handle, ninst, N, alphas=(1.0,)):
        self.h, self.ninst, self.N = handle, int(ninst), int(N)
        al = np.ascontiguousarray(alphas, np.float64)
        self._w = C.c_void_p()
        handle._check(lib().ilqg_ilqr_create(handle._h, self.ninst, self.N, len(al), _hp(al), C.byref(self._w)))

    def close(self):
        if self._w:
            lib().ilqg_ilqr_destroy(self._w)
            self._w = C.c_void_p()

    def set_cost(self, cost):
        self.h._check(lib().ilqg_ilqr_set_cost(self._w, _hp(cost)))

    def set_layout(self, corrected):
        self.h._check(lib().ilqg_ilqr_set_layout(self._w, 1 if corrected else 0))

    def set_mu_schedule(self, factor, mu_min=1e-6, mu_max=1e10):
        self.h._check(lib().ilqg_ilqr_set_mu_schedule(self._w, C.c_double(factor), C.c_double(mu_min), C.c_double(mu_max)))

    def set_mu(self, mu):
        self.h._check(lib().ilqg_ilqr_set_mu(self._w, C.c_double(mu)))

    def init_host(self, qpos, qvel, ctrl, warm=None):
        m = self.h.model
        a = [np.ascontiguousarray(x, np.float64) if x is not None else None for x in (qpos, qvel, ctrl, warm)]
        assert a[0].shape == (self.ninst, m.nq) and a[1].shape == (self.ninst, m.nv)
        self.h._check(lib().ilqg_ilqr_init_host(self._w, _hp(a[0]), _hp(a[1]), _hp(a[2]), _hp(a[3])))

    def set_state_host(self, qpos, qvel, warm=None):
        a = [np.ascontiguousarray(x, np.float64) if x is not None else None for x in (qpos, qvel, warm)]
        self.h._check(lib().ilqg_ilqr_set_state_host(self._w, _hp(a[0]), _hp(a[1]), _hp(a[2])))

    def init_dev(self, qpos, qvel, ctrl, warm=None, stream=None):
        self.h._check(lib().ilqg_ilqr_init_dev(self._w, _dp(qpos), _dp(qvel), _dp(ctrl), _dp(warm), C.c_void_p(stream) if stream else None))

    def iterate(self, niter=1, accept_always=True, stream=None):
        self.h._check(lib().ilqg_ilqr_iterate(self._w, int(niter), 1 if accept_always else 0, C.c_void_p(stream) if stream else None))

    @property
    def iterations(self):
        return int(lib().ilqg_ilqr_iterations_done(self._w))

    def fetch_controls(self, last=None):
        """First control of every problem (dArray[N]->ctrl, what InvertedPendulum::forward applies) and the cost trace of the
        iterations run so far (`last`: only of the last that many): (u0[ninst, nu], J[ninst, kept]).  Synchronises."""
        m = self.h.model
        kept = min(self.iterations, 256) if last is None else min(self.iterations, 256, int(last))
        u0 = np.zeros((self.ninst, m.nu))
        J = np.zeros((self.ninst, kept))
        self.h._check(lib().ilqg_ilqr_get_first_control_last_host(self._w, kept if last is not None else 256, _hp(u0), _hp(J)))
        return u0, J

    def get(self):
        m = self.h.model
        n, T, nx = self.ninst, self.N + 1, 2 * m.nv
        kept = min(self.iterations, 256)
        out = dict(qpos=np.zeros((n, T, m.nq)), qvel=np.zeros((n, T, m.nv)), ctrl=np.zeros((n, T, m.nu)), K=np.zeros((n, T, m.nu * nx)),
                   k=np.zeros((n, T, m.nu)), V=np.zeros((n, nx * nx)), v=np.zeros((n, nx)), J=np.zeros((n, kept)),
                   accepted=np.zeros((n, kept), np.int32))
        self.h._check(lib().ilqg_ilqr_get_host(self._w, _hp(out["qpos"]), _hp(out["qvel"]), _hp(out["ctrl"]), _hp(out["K"]), _hp(out["k"]),
                                               _hp(out["V"]), _hp(out["v"]), _hp(out["J"]), _hp(out["accepted"])))
        return out
