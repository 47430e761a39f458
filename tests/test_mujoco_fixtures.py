"""Parity of the CPU oracle with MuJoCo ITSELF — consumed when tests/golden/mujoco_<model>.npz exist.

Those files are written by tools/mujoco_fixtures.py on a machine that has the `mujoco` package (this image does not: SURVEY.md F2);
the script's docstring lists, field by field, which formula of oracle/mjo_engine.c each fixture entry falsifies.  Without the
files every test here SKIPS with a message that says so: parity with upstream libmujoco is then UNPINNED, which is what
DESIGN.md section 5 and the oracle's header state."""
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
NAMES = ("inverted_pendulum", "hopper", "humanoid")


def fixture(name):
    path = os.path.join(GOLD, f"mujoco_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"PARITY WITH UPSTREAM MUJOCO UNPINNED: {path} is absent (no `mujoco` package in this image). "
                    "Run `python tools/mujoco_fixtures.py --res <iLQG-MuJoCo>/res` where MuJoCo 2.1.2 <= v < 3 is installed and commit the .npz files.")
    return np.load(path)


def close(a, b, rtol, atol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), f"{what}: max violation {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


@pytest.mark.parametrize("name", NAMES)
def test_compiled_constants_match_mujoco(pkg, name):
    """body masses / inertias from geoms, dof_invweight0 / body_invweight0 / meaninertia (csrc/mjcf_compile.cpp vs mj_loadXML)."""
    g = fixture(name)
    pm = pkg.Model.named(name)
    nb, nv = pm.nbody, pm.nv
    close(pm.field("body_mass")[:nb], g["body_mass"], 1e-9, 1e-12, "body_mass")
    close(pm.field("body_ipos").reshape(-1, 3)[:nb], g["body_ipos"], 1e-9, 1e-10, "body_ipos")
    # the tables keep the full inertia tensor in body axes; MuJoCo its principal moments (body_inertia) and frame (body_iquat)
    for b in range(1, nb):
        I = pm.field("body_inertia").reshape(-1, 6)[b]
        T = np.array([[I[0], I[3], I[4]], [I[3], I[1], I[5]], [I[4], I[5], I[2]]])
        close(np.sort(np.linalg.eigvalsh(T)), np.sort(g["body_inertia"][b]), 1e-8, 1e-12, f"principal inertia of body {b}")
    close(pm.field("dof_invweight0")[:nv], g["dof_invweight0"], 1e-8, 1e-12, "dof_invweight0")
    close(pm.field("body_invweight0").reshape(-1, 2)[:nb], g["body_invweight0"], 1e-8, 1e-12, "body_invweight0")
    close(pm.field("meaninertia")[0], g["stat_meaninertia"], 1e-9, 0, "stat.meaninertia")


@pytest.mark.parametrize("name", NAMES)
def test_forward_dynamics_intermediates_match_mujoco(oracle, omodels, name):
    g = fixture(name)
    om = omodels[name]
    for k in range(len(g["qpos"])):
        d = oracle.dump(om, g["qpos"][k], g["qvel"][k], g["ctrl"][k], g["warm"][k], 30, 0.0)
        for f, tol in (("xpos", 1e-10), ("xquat", 1e-10), ("xipos", 1e-10), ("subtree_com", 1e-10), ("qM", 1e-9), ("qfrc_bias", 1e-8), ("qfrc_passive", 1e-9),
                       ("qfrc_actuator", 1e-10), ("qacc_smooth", 1e-8)):
            ref = g[f][k]
            if f == "subtree_com":   # the oracle keeps the tree's com at the root body only
                ref = ref[1:2]; got = d[f][1:2]
            else:
                got = d[f]
            close(got, ref, tol, tol, f"{f} of state {k}")
        assert d["ncon"] == int(g["ncon"][k]), f"ncon of state {k}"
        assert d["nefc"] == int(g["nefc"][k]), f"nefc of state {k}"
        # contacts as a SET (MuJoCo's order within the broad phase is a documented deviation): match by geom pair, then by position
        nc = d["ncon"]
        def key(geom, pos):
            return sorted(range(nc), key=lambda c: (tuple(int(x) for x in geom[c]), tuple(np.round(pos[c], 6))))
        oo, mm = key(d["contact_geom"], d["contact_pos"]), key(g["contact_geom"][k][:nc], g["contact_pos"][k][:nc])
        close(d["contact_dist"][oo], g["contact_dist"][k][:nc][mm], 1e-9, 1e-11, f"contact dist of state {k}")
        close(d["contact_pos"][oo], g["contact_pos"][k][:nc][mm], 1e-9, 1e-10, f"contact pos of state {k}")
        close(d["contact_frame"][oo], g["contact_frame"][k][:nc][mm], 1e-8, 1e-9, f"contact frame of state {k}")
        # rows: limits first (same order), contact rows follow their contacts -> compare the multiset of (R, aref, pos) and J row norms
        ne = d["nefc"]
        for f, tol in (("efc_pos", 1e-9), ("efc_R", 1e-7), ("efc_D", 1e-7), ("efc_aref", 1e-7), ("efc_diagApprox", 1e-8)):
            close(np.sort(d[f]), np.sort(g[f][k][:ne]), tol, tol, f"{f} (sorted) of state {k}")
        close(np.sort(np.linalg.norm(d["efc_J"], axis=1)), np.sort(np.linalg.norm(g["efc_J"][k][:ne], axis=1)), 1e-8, 1e-9, f"efc_J row norms of state {k}")
        # the minimiser is unique whatever the solver path
        scale = max(1.0, np.abs(g["qacc"][k]).max())
        close(d["qacc"], g["qacc"][k], 0, 1e-7 * scale, f"qacc of state {k}")
        close(np.sort(d["efc_force"]), np.sort(g["efc_force"][k][:ne]), 1e-6, 1e-6 * max(1.0, np.abs(g["efc_force"][k]).max()), f"efc_force of state {k}")


@pytest.mark.parametrize("name", NAMES)
def test_step_and_fd_match_mujoco(oracle, omodels, name):
    g = fixture(name)
    om = omodels[name]
    q1, v1, _, _ = oracle.step_batch(om, g["qpos"], g["qvel"], g["ctrl"], g["warm"], 1)
    close(q1, g["step1_qpos"], 1e-9, 1e-10, "qpos after 1 mj_step")
    close(v1, g["step1_qvel"], 1e-7, 1e-8, "qvel after 1 mj_step")
    q20, v20, _, _ = oracle.step_batch(om, g["qpos"], g["qvel"], g["ctrl"], g["warm"], 20)
    close(q20, g["step20_qpos"], 1e-6, 1e-7, "qpos after 20 mj_step")
    deriv, qc, _ = oracle.fd_batch(om, g["qpos"], g["qvel"], g["ctrl"], g["warm"], None)
    nv, nu = om.nv, om.nu
    njac = nv * (2 * nv + nu)
    # SURVEY 8d tolerance for contact-rich models: 1e-4 relative Frobenius per knot at p99; 1e-6 on the smooth pendulum
    rel = np.linalg.norm(deriv[:, :njac] - g["deriv"][:, :njac], axis=1) / np.maximum(1.0, np.linalg.norm(g["deriv"][:, :njac], axis=1))
    assert np.percentile(rel, 99) <= (1e-6 if name == "inverted_pendulum" else 1e-4), rel
    close(qc, g["fd_center_qacc"], 0, 1e-7 * max(1.0, np.abs(g["fd_center_qacc"]).max()), "centre qacc after the warm-up solves")
