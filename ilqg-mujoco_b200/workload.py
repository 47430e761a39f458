"""Synthetic workloads of the shapes BASELINE.json names, generated on the GPU through
the product's own step kernel (no oracle, no reference).  Deterministic for a given seed."""
import numpy as np
import torch


def hopper_initial_states(n, seed=0):
    """Config-2 generator: rootz ~ U[1.0,1.4], joint angles within 50 % of their range,
    qvel ~ N(0,0.5^2), ctrl ~ U[-1,1] (ranges: /root/reference/res/hopper.xml:18,21,24,32-34)."""
    rng = np.random.default_rng(seed)
    qpos = np.zeros((n, 6))
    qpos[:, 0] = rng.uniform(-0.1, 0.1, n)
    qpos[:, 1] = rng.uniform(1.0, 1.4, n)
    qpos[:, 2] = rng.uniform(-0.1, 0.1, n)
    lo = np.array([-150.0, -150.0, -45.0]) * np.pi / 180
    hi = np.array([0.0, 0.0, 45.0]) * np.pi / 180
    mid, half = 0.5 * (lo + hi), 0.5 * (hi - lo)
    qpos[:, 3:] = mid + rng.uniform(-0.5, 0.5, (n, 3)) * half
    qvel = rng.normal(0, 0.5, (n, 6))
    ctrl = rng.uniform(-1, 1, (n, 3))
    roll = rng.integers(0, 11, n) * 20  # pre-roll steps in {0,20,...,200}
    return qpos, qvel, ctrl, roll


def pendulum_initial_states(n, seed=0):
    """Config-4 generator: slider, hinge ~ U[-0.5,0.5], qvel ~ N(0,0.5^2)."""
    rng = np.random.default_rng(seed)
    qpos = rng.uniform(-0.5, 0.5, (n, 2))
    qvel = rng.normal(0, 0.5, (n, 2))
    ctrl = rng.uniform(-0.5, 0.5, (n, 1))
    return qpos, qvel, ctrl, np.zeros(n, dtype=np.int64)


def humanoid_states(handle, n, seed=0, device="cuda:0"):
    """Config-3 generator: torso z ~ U[1.2,1.5], unit quaternion perturbed by N(0,0.1^2) in the tangent space, joints within
    30 % of their range, qvel ~ N(0,0.3^2), ctrl ~ U[-0.4,0.4] (/root/reference/res/humanoid.xml:16), rolled forward
    0..300 steps on the device."""
    m = handle.model
    rng = np.random.default_rng(seed)
    rngs = m.field("jnt_range").reshape(-1, 2)[1:m.njnt]
    mid, half = 0.5 * (rngs[:, 0] + rngs[:, 1]), 0.5 * (rngs[:, 1] - rngs[:, 0])
    q = np.zeros((n, m.nq))
    q[:, 2] = rng.uniform(1.2, 1.5, n)
    w = rng.normal(0, 0.1, (n, 3))
    ang = np.linalg.norm(w, axis=1, keepdims=True)
    q[:, 3] = np.cos(ang[:, 0] / 2)
    q[:, 4:7] = w / np.maximum(ang, 1e-12) * np.sin(ang / 2)
    q[:, 7:] = mid + rng.uniform(-0.3, 0.3, (n, m.njnt - 1)) * half
    v = rng.normal(0, 0.3, (n, m.nv))
    u = rng.uniform(-0.4, 0.4, (n, m.nu))
    roll = rng.integers(0, 7, n) * 50
    order = np.argsort(roll, kind="stable")
    q, v, u, roll = q[order], v[order], u[order], roll[order]
    dq, dv, du = (torch.from_numpy(a).to(device) for a in (q, v, u))
    dw = torch.zeros((n, m.nv), dtype=torch.float64, device=device)
    for r in range(50, int(roll.max()) + 1, 50):
        start = int(np.searchsorted(roll, r, side="left"))
        if start < n:
            handle.step_batch_dev(dq[start:], dv[start:], du[start:], dw[start:], None, nsteps=50)
    torch.cuda.synchronize()
    bad = ~(torch.isfinite(dq).all(dim=1) & torch.isfinite(dv).all(dim=1) & torch.isfinite(dw).all(dim=1))
    nbad = int(bad.sum())
    if nbad:
        dq[bad] = torch.from_numpy(m.field("qpos0")[:m.nq].copy()).to(device)
        dv[bad] = 0.0
        dw[bad] = 0.0
    return dq.contiguous(), dv.contiguous(), du.contiguous(), dw.contiguous(), nbad


def make_knots(handle, ntraj, T, seed=0, device="cuda:0", model="hopper"):
    """ntraj trajectories x T knots: random initial states, pre-rolled 0..200 steps so that a good share
    of the knots is in ground contact, then T knots one mj_step apart under constant control.
    Returns device tensors (qpos[ntraj*T,nq], qvel, ctrl, warm) in trajectory-major order."""
    m = handle.model
    gen = hopper_initial_states if model == "hopper" else pendulum_initial_states
    qpos, qvel, ctrl, roll = gen(ntraj, seed)
    order = np.argsort(roll, kind="stable")
    qpos, qvel, ctrl, roll = qpos[order], qvel[order], ctrl[order], roll[order]
    dq = torch.from_numpy(qpos).to(device)
    dv = torch.from_numpy(qvel).to(device)
    du = torch.from_numpy(ctrl).to(device)
    dw = torch.zeros((ntraj, m.nv), dtype=torch.float64, device=device)
    # pre-roll: instances sorted by roll count; advance the tail that still has steps to go, 20 at a time
    for r in range(20, int(roll.max()) + 1, 20):
        start = int(np.searchsorted(roll, r, side="left"))
        if start < ntraj:
            handle.step_batch_dev(dq[start:], dv[start:], du[start:], dw[start:], None, nsteps=20)
    Q = torch.empty((ntraj, T, m.nq), dtype=torch.float64, device=device)
    V = torch.empty((ntraj, T, m.nv), dtype=torch.float64, device=device)
    W = torch.empty((ntraj, T, m.nv), dtype=torch.float64, device=device)
    for t in range(T):
        Q[:, t], V[:, t], W[:, t] = dq, dv, dw
        if t + 1 < T:
            handle.step_batch_dev(dq, dv, du, dw, None, nsteps=1)
    U = du[:, None, :].expand(ntraj, T, m.nu).contiguous()
    torch.cuda.synchronize()
    # a diverged pre-roll (non-finite state) would poison the benchmark: replace by the model's rest pose
    bad = ~(torch.isfinite(Q).all(dim=2) & torch.isfinite(V).all(dim=2) & torch.isfinite(W).all(dim=2))
    nbad = int(bad.sum())
    if nbad:
        Q[bad] = 0.0
        V[bad] = 0.0
        W[bad] = 0.0
        if model == "hopper":
            Q[bad, 1] = 1.25
    return (Q.reshape(-1, m.nq).contiguous(), V.reshape(-1, m.nv).contiguous(), U.reshape(-1, m.nu).contiguous(),
            W.reshape(-1, m.nv).contiguous(), nbad)


def make_knots_8d(handle, ntraj, T, seed=0, device="cuda:0", roll_step=20, roll_max=360, ctrl_amp=0.25):
    """The hopper batch SURVEY.md 8(d) config 2 specifies: the config-2 initial states pre-rolled long enough that about HALF of
    the knots are in ground contact (pre-roll uniform over {0, 20, ..., roll_max} steps; calibrated with the CPU oracle: 200 ->
    25 %, 300 -> 40 %, 400 -> 56 % of the knots with at least one contact), then T knots one mj_step apart under a TIME-VARYING
    control u_t = clip(u_0 + ctrl_amp sin(w t + phi), -1, 1) (per actuator: w ~ U[1, 8] Hz, phi ~ U[0, 2 pi]).
    Returns device tensors (qpos[ntraj*T,nq], qvel, ctrl, warm) in trajectory-major order and the count of replaced knots."""
    m = handle.model
    qpos, qvel, ctrl, _ = hopper_initial_states(ntraj, seed)
    rng = np.random.default_rng(seed + 7919)
    roll = rng.integers(0, roll_max // roll_step + 1, ntraj) * roll_step
    freq = rng.uniform(1.0, 8.0, (ntraj, m.nu)) * 2 * np.pi
    phase = rng.uniform(0, 2 * np.pi, (ntraj, m.nu))
    order = np.argsort(roll, kind="stable")
    qpos, qvel, ctrl, roll, freq, phase = qpos[order], qvel[order], ctrl[order], roll[order], freq[order], phase[order]
    dq, dv, du0 = (torch.from_numpy(a).to(device) for a in (qpos, qvel, ctrl))
    dfr, dph = torch.from_numpy(freq).to(device), torch.from_numpy(phase).to(device)
    dw = torch.zeros((ntraj, m.nv), dtype=torch.float64, device=device)
    for r in range(roll_step, int(roll.max()) + 1, roll_step):   # pre-roll under the constant u_0, the tail that still has steps to go
        start = int(np.searchsorted(roll, r, side="left"))
        if start < ntraj:
            handle.step_batch_dev(dq[start:], dv[start:], du0[start:], dw[start:], None, nsteps=roll_step)
    Q = torch.empty((ntraj, T, m.nq), dtype=torch.float64, device=device)
    V = torch.empty((ntraj, T, m.nv), dtype=torch.float64, device=device)
    W = torch.empty((ntraj, T, m.nv), dtype=torch.float64, device=device)
    U = torch.empty((ntraj, T, m.nu), dtype=torch.float64, device=device)
    for t in range(T):
        ut = torch.clamp(du0 + ctrl_amp * torch.sin(dfr * (t * m.timestep) + dph), -1.0, 1.0).contiguous()
        Q[:, t], V[:, t], W[:, t], U[:, t] = dq, dv, dw, ut
        if t + 1 < T:
            handle.step_batch_dev(dq, dv, ut, dw, None, nsteps=1)
    torch.cuda.synchronize()
    bad = ~(torch.isfinite(Q).all(dim=2) & torch.isfinite(V).all(dim=2) & torch.isfinite(W).all(dim=2))
    nbad = int(bad.sum())
    if nbad:
        Q[bad] = 0.0
        V[bad] = 0.0
        W[bad] = 0.0
        Q[bad, 1] = 1.25
    return (Q.reshape(-1, m.nq).contiguous(), V.reshape(-1, m.nv).contiguous(), U.reshape(-1, m.nu).contiguous(),
            W.reshape(-1, m.nv).contiguous(), nbad)
