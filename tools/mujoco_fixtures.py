#!/usr/bin/env python
"""Pin the physics to MuJoCo itself — the day a MuJoCo wheel is available.

The reference links libmujoco (/root/reference/Makefile:19, /root/reference/src/mjderivative.cpp:5); neither the library nor its
Python bindings exist in this image, so parity of oracle/mjo_engine.c with upstream MuJoCo is UNPINNED.  This script
closes that with one command on any machine that has `pip install mujoco` (2.1.2 <= version < 3 matches the reference's API use;
3.x is accepted with a warning — its defaults differ in places the script neutralises, see NEUTRALISE below) and the reference's
model files:

    python tools/mujoco_fixtures.py --res /path/to/iLQG-MuJoCo/res [--out tests/golden]

It loads res/{inverted_pendulum,hopper,humanoid}.xml with mj_loadXML, puts MuJoCo on the seeded states of tests/conftest.py
(scenario_states: the same generator the GPU parity tests use; rolled forward with mj_step so that contacts and warm starts are
live), runs mj_forward / mj_step / the reference's FD schedule (mjderivative.cpp:43-255 restated on the bindings: eps 1e-6,
iterations 30, tolerance 0, nwarmup 3, central differences, mju_quatIntegrate on quaternion dofs) and writes
tests/golden/mujoco_<model>.npz.  tests/test_mujoco_fixtures.py consumes the files when present and SKIPS LOUDLY otherwise.

What each fixture field falsifies (SURVEY.md Appendix A.2 — every formula of the restated pipeline has a field):

| fixture field (mjModel / mjData name)              | oracle formula it checks                                                        |
|---|---|
| body_mass, body_inertia, body_ipos (model)         | inertia from capsule / sphere geoms at density 1000, `coordinate="global"`, fromto |
| dof_invweight0, body_invweight0, stat_meaninertia  | compile-time constants behind efc_diagApprox -> R and the solver's scaling         |
| xpos, xquat, xipos, subtree_com                    | mj_kinematics (quaternion chain, joint anchors), mj_comPos                         |
| qM (mj_fullM)                                      | mj_crb + armature                                                                  |
| qfrc_bias, qfrc_passive, qfrc_actuator             | mj_rne(flg_acc=0), joint damping / springs, gear + ctrlrange clamp (ctrllimited)   |
| qacc_smooth                                        | M^-1 (passive - bias + actuator)                                                   |
| ncon, contact_dist / pos / frame / geom            | inclusion rule (margin, parent-child filter), plane-sphere / plane-capsule / sphere-sphere / sphere-capsule / capsule-capsule narrow phase, mju_makeFrame incl. the capsule-axis tangent hint; CONTACT ORDER (documented deviation: compared as a set) |
| nefc, efc_J                                        | limit rows (sign, side), contact Jacobian, pyramidal facets J_n +- mu J_t, row order |
| efc_pos, efc_margin, efc_diagApprox, efc_R, efc_D  | impedance d(r) of solimp, R = (1-d)/d diagApprox, pyramid R = 2 mu^2 R_first        |
| efc_aref                                           | K, B from solref (refsafe), aref = -B vel - K d (pos - margin)                      |
| efc_force, qacc                                    | the constraint problem's MINIMISER (unique: the solver path is free)    |
| step_qpos, step_qvel (1 and 20 steps)              | mj_Euler with implicit joint damping / mj_RungeKutta(4), mj_integratePos, quaternions |
| deriv                                              | the FD schedule on top of all of the above (warm-start handling, stage skipping)   |
| warm_after                                         | where MuJoCo saves qacc_warmstart (inside mj_fwdConstraint in 2.x)                  |
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODELS = {"inverted_pendulum": dict(n=24, roll=10), "hopper": dict(n=24, roll=150), "humanoid": dict(n=8, roll=60)}
EPS, NITER, NWARMUP = 1e-6, 30, 3   # /root/reference/src/mjderivative.cpp:37-39


def version(mujoco):
    """NEUTRALISE: options whose DEFAULTS moved between MuJoCo 2.x (what the reference was written against) and 3.x are forced
    back in load(): 3.x's <compiler autolimits="true"> would infer ctrllimited from the ctrlrange of res/inverted_pendulum.xml:7,
    which 2.x does not clamp."""
    ver = tuple(int(x) for x in mujoco.__version__.split(".")[:2])
    if ver >= (3, 0):
        print(f"WARNING: MuJoCo {mujoco.__version__} >= 3: the reference's API use brackets 2.1.2 <= v < 3; forcing 2.x semantics where defaults moved",
              file=sys.stderr)
    return ver


def load(mujoco, path, ver):
    xml = open(path).read()
    if ver >= (3, 0) and "autolimits" not in xml:
        if "<compiler" in xml:
            xml = xml.replace("<compiler", '<compiler autolimits="false"', 1)
        else:
            head = xml.index(">", xml.index("<mujoco")) + 1
            xml = xml[:head] + '\n  <compiler autolimits="false"/>' + xml[head:]
    m = mujoco.MjModel.from_xml_string(xml)
    m.opt.jacobian = mujoco.mjtJacobian.mjJAC_DENSE
    m.opt.solver = mujoco.mjtSolver.mjSOL_NEWTON
    m.opt.cone = mujoco.mjtCone.mjCONE_PYRAMIDAL
    return m


def reference_fd(mujoco, m, d_main):
    """calcMJDerivatives (/root/reference/src/mjderivative.cpp:43-255) on the bindings, single worker (all columns)."""
    nv, nu = m.nv, m.nu
    deriv = np.zeros(nv * (2 * nv + nu) + 2 * nv + nu)
    save = (m.opt.iterations, m.opt.tolerance)
    m.opt.iterations, m.opt.tolerance = NITER, 0.0
    d = mujoco.MjData(m)

    def cp(dst, src):   # cpMjData, /root/reference/src/util.cpp:4-13
        dst.time = src.time
        for f in ("qpos", "qvel", "qacc", "qacc_warmstart", "qfrc_applied", "xfrc_applied", "ctrl"):
            getattr(dst, f)[:] = getattr(src, f)
    cp(d, d_main)
    mujoco.mj_forward(m, d)
    for _ in range(1, NWARMUP):
        mujoco.mj_forwardSkip(m, d, mujoco.mjtStage.mjSTAGE_VEL, 1)
    warm = d.qacc_warmstart.copy()
    center = d.qacc.copy()

    def solve(stage):
        d.qacc_warmstart[:] = warm
        mujoco.mj_forwardSkip(m, d, stage, 1)
        return d.qacc.copy()
    S = mujoco.mjtStage
    for i in range(nu):
        d.ctrl[i] += EPS; plus = solve(S.mjSTAGE_VEL); d.ctrl[i] = d_main.ctrl[i]
        d.ctrl[i] -= EPS; minus = solve(S.mjSTAGE_VEL); d.ctrl[i] = d_main.ctrl[i]
        deriv[2 * nv * nv + i + np.arange(nv) * nu] = (plus - minus) / (2 * EPS)
    for i in range(nv):
        d.qvel[i] += EPS; plus = solve(S.mjSTAGE_POS); d.qvel[i] = d_main.qvel[i]
        d.qvel[i] -= EPS; minus = solve(S.mjSTAGE_POS); d.qvel[i] = d_main.qvel[i]
        deriv[nv * nv + i + np.arange(nv) * nv] = (plus - minus) / (2 * EPS)
    for i in range(nv):
        jid = m.dof_jntid[i]
        quatadr, dofpos = -1, 0
        if m.jnt_type[jid] == mujoco.mjtJoint.mjJNT_BALL:
            quatadr, dofpos = m.jnt_qposadr[jid], i - m.jnt_dofadr[jid]
        elif m.jnt_type[jid] == mujoco.mjtJoint.mjJNT_FREE and i >= m.jnt_dofadr[jid] + 3:
            quatadr, dofpos = m.jnt_qposadr[jid] + 3, i - m.jnt_dofadr[jid] - 3
        res = []
        for sgn in (1.0, -1.0):
            if quatadr >= 0:
                angvel = np.zeros(3); angvel[dofpos] = sgn * EPS
                quat = d.qpos[quatadr:quatadr + 4].copy()
                mujoco.mju_quatIntegrate(quat, angvel, 1.0)
                d.qpos[quatadr:quatadr + 4] = quat
            else:
                d.qpos[m.jnt_qposadr[jid] + i - m.jnt_dofadr[jid]] += sgn * EPS
            res.append(solve(S.mjSTAGE_NONE))
            d.qpos[:] = d_main.qpos
        deriv[i + np.arange(nv) * nv] = (res[0] - res[1]) / (2 * EPS)
    m.opt.iterations, m.opt.tolerance = save
    return deriv, center, warm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="/root/reference/res", help="directory with the reference's MJCF files")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    try:
        import mujoco
    except ImportError:
        print("mujoco is not importable here: `pip install 'mujoco<3'` on a machine with network access, then re-run.\n"
              "Until then tests/test_mujoco_fixtures.py skips and parity with upstream MuJoCo stays UNPINNED.", file=sys.stderr)
        return 2
    from conftest import scenario_states
    ver = version(mujoco)
    for name, cfg in MODELS.items():
        m = load(mujoco, os.path.join(args.res, name + ".xml"), ver)
        q0, v0, u0, _ = scenario_states(name, cfg["n"], seed=2024)
        rec = {k: [] for k in ("qpos", "qvel", "ctrl", "warm", "xpos", "xquat", "xipos", "subtree_com", "qM", "qfrc_bias", "qfrc_passive", "qfrc_actuator",
                               "qacc_smooth", "qacc", "warm_after", "ncon", "nefc", "contact_dist", "contact_pos", "contact_frame", "contact_geom",
                               "efc_J", "efc_pos", "efc_margin", "efc_diagApprox", "efc_R", "efc_D", "efc_aref", "efc_force", "step1_qpos", "step1_qvel",
                               "step20_qpos", "step20_qvel", "deriv", "fd_center_qacc")}
        for k in range(cfg["n"]):
            d = mujoco.MjData(m)
            d.qpos[:], d.qvel[:], d.ctrl[:] = q0[k], v0[k], u0[k]
            for _ in range(cfg["roll"]):
                mujoco.mj_step(m, d)
            if not (np.isfinite(d.qpos).all() and np.isfinite(d.qvel).all()):
                continue
            rec["qpos"].append(d.qpos.copy()); rec["qvel"].append(d.qvel.copy()); rec["ctrl"].append(d.ctrl.copy())
            rec["warm"].append(d.qacc_warmstart.copy())
            save = (m.opt.iterations, m.opt.tolerance)
            m.opt.iterations, m.opt.tolerance = NITER, 0.0
            mujoco.mj_forward(m, d)
            m.opt.iterations, m.opt.tolerance = save
            for f in ("xpos", "xquat", "xipos", "subtree_com", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qacc_smooth", "qacc"):
                rec[f].append(np.array(getattr(d, f)).copy())
            rec["warm_after"].append(d.qacc_warmstart.copy())
            M = np.zeros((m.nv, m.nv)); mujoco.mj_fullM(m, M, d.qM); rec["qM"].append(M)
            rec["ncon"].append(d.ncon); rec["nefc"].append(d.nefc)

            def pad(a, n, w):
                out = np.zeros((n, w)); a = np.asarray(a, np.float64).reshape(-1, w); out[:len(a)] = a; return out
            rec["contact_dist"].append(pad([c.dist for c in d.contact[:d.ncon]], 96, 1)[:, 0])
            rec["contact_pos"].append(pad([c.pos for c in d.contact[:d.ncon]], 96, 3))
            rec["contact_frame"].append(pad([c.frame for c in d.contact[:d.ncon]], 96, 9))
            rec["contact_geom"].append(pad([[c.geom1, c.geom2] for c in d.contact[:d.ncon]], 96, 2))
            rec["efc_J"].append(pad(np.array(d.efc_J).reshape(-1, m.nv)[:d.nefc], 320, m.nv))
            for f in ("efc_pos", "efc_margin", "efc_diagApprox", "efc_R", "efc_D", "efc_aref", "efc_force"):
                rec[f].append(pad(np.array(getattr(d, f))[:d.nefc], 320, 1)[:, 0])
            base = mujoco.MjData(m)
            base.qpos[:], base.qvel[:], base.ctrl[:], base.qacc_warmstart[:] = rec["qpos"][-1], rec["qvel"][-1], rec["ctrl"][-1], rec["warm"][-1]
            dv, qc, _ = reference_fd(mujoco, m, base)
            rec["deriv"].append(dv); rec["fd_center_qacc"].append(qc)
            s = mujoco.MjData(m)
            s.qpos[:], s.qvel[:], s.ctrl[:], s.qacc_warmstart[:] = rec["qpos"][-1], rec["qvel"][-1], rec["ctrl"][-1], rec["warm"][-1]
            mujoco.mj_step(m, s)
            rec["step1_qpos"].append(s.qpos.copy()); rec["step1_qvel"].append(s.qvel.copy())
            for _ in range(19):
                mujoco.mj_step(m, s)
            rec["step20_qpos"].append(s.qpos.copy()); rec["step20_qvel"].append(s.qvel.copy())
        out = {k: np.array(v) for k, v in rec.items()}
        out.update(mujoco_version=np.array(mujoco.__version__), body_mass=m.body_mass.copy(), body_inertia=m.body_inertia.copy(), body_ipos=m.body_ipos.copy(),
                   body_iquat=m.body_iquat.copy(), dof_invweight0=m.dof_invweight0.copy(), body_invweight0=m.body_invweight0.copy(),
                   stat_meaninertia=np.array(m.stat.meaninertia), geom_type=m.geom_type.copy(), geom_bodyid=m.geom_bodyid.copy(), timestep=np.array(m.opt.timestep))
        os.makedirs(args.out, exist_ok=True)
        path = os.path.join(args.out, f"mujoco_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"wrote {path}: {len(out['qpos'])} states, MuJoCo {mujoco.__version__}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
