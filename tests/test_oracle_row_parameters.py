"""The soft-constraint parameters of EVERY row (joint limits, ground contacts, body-body pairs) at random folded, moving poses of hopper
and humanoid, recomputed from the attributes in the MJCF text by the formulas of SURVEY.md Appendix A.2 (MuJoCo 2.x's documented
solref / solimp model), with the inverse weights re-derived here instead of read from the compiled tables:

  impedance d(|pos - margin|) from solimp = (d0, dmax, width, midpoint, power)  [the humanoid's default solimp makes it vary with depth]
  K = 1 / (dmax^2 tc^2 dr^2),  B = 2 / (dmax tc),  tc = max(solref[0], 2 dt),  aref = -B (J qvel) - K d (pos - margin)
  limit row of dof i :  diagApprox = (M0^-1)_ii                                        R = (1 - d) / d * diagApprox
  contact, condim 1  :  diagApprox = tran_1 + tran_2                                   R = (1 - d) / d * diagApprox
  contact, pyramidal :  diagApprox = (tran_1 + tran_2)(1 + mu^2)                       R = 2 mu^2 (1 - d) / d * diagApprox     D = 1 / R
  tran_b = tr(J_b M0^-1 J_b') / 3 at qpos0, J_b = velocity Jacobian of body b's centre of mass — by numerical differentiation of the XML
  forward kinematics (tests/test_oracle_contact_rows.py); M0 = the oracle's qM at qpos0 (anchored as the Hessian of the kinetic energy in
  tests/test_oracle_first_principles.py).
tests/test_oracle_contact_anchors.py does this by hand for the hopper's foot at one state; this file covers every row the oracle
generates.  What it cannot settle is whether upstream MuJoCo uses exactly these formulas: that is the job of tools/mujoco_fixtures.py.
(The XML files are only present in the authoring container: the tests skip elsewhere.)"""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from test_oracle_contact_rows import fk_frames, geom_table, move
from test_oracle_first_principles import RES, vec, walk

SOLREF0, SOLIMP0 = (0.02, 1.0), (0.9, 0.95, 0.001, 0.5, 2.0)      # MuJoCo's defaults


def impedance(solimp, x):
    d0, dmax, width, mid, power = solimp
    if d0 == dmax:
        return d0
    x = abs(x) / width
    if x >= 1:
        return dmax
    y = x if power == 1 else (x ** power / mid ** (power - 1) if x <= mid else 1 - (1 - x) ** power / (1 - mid) ** (power - 1))
    return d0 + y * (dmax - d0)


def solver_attrs(elem, dflt):
    """(solref, solimp, margin) of a geom / joint element with its default class's values behind it"""
    def get(key, std):
        s = elem.get(key, dflt.get(key) if dflt is not None else None)
        if s is None:
            return std
        v = list(vec(s))
        return tuple(v + list(std[len(v):])) if isinstance(std, tuple) else v[0]
    return get("solref", SOLREF0), get("solimp", SOLIMP0), get("margin", 0.0)


@pytest.mark.parametrize("name", ["hopper", "humanoid"])
def test_every_row_s_parameters_from_the_mjcf_text(oracle, omodels, pkg, name):
    path = os.path.join(RES, name + ".xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    degrees = root.find("compiler").get("angle", "degree") == "degree"
    dt = float(root.find("option").get("timestep"))
    om = omodels[name]
    m = pkg.Model.named(name)
    gt = geom_table(root)
    dflt = root.find("default")
    gdflt, jdflt = dflt.find("geom"), dflt.find("joint")
    bodies = []
    for b in root.find("worldbody").findall("body"):
        walk(b, bodies)
    geoms = list(root.find("worldbody").findall("geom")) + [g for body in bodies for g in body.findall("geom")]
    gattr = [solver_attrs(g, gdflt) for g in geoms]
    # scalar joints in dof order with their ranges (radians) and solver attributes
    jinfo, di = {}, 0
    for body in bodies:
        for j in body:
            if j.tag == "freejoint":
                di += 6
            elif j.tag == "joint":
                lim = j.get("limited", jdflt.get("limited", "false") if jdflt is not None else "false") == "true"
                rg = vec(j.get("range", "0 0"))
                if degrees and j.get("type", "hinge") == "hinge":
                    rg = np.deg2rad(rg)
                jinfo[di] = (lim, rg, solver_attrs(j, jdflt))
                di += 1
    assert di == m.nv
    # inverse weights at qpos0, re-derived
    q0 = m.field("qpos0")[:m.nq].copy()
    d0 = oracle.dump(om, q0, np.zeros(m.nv), np.zeros(m.nu))
    Minv = np.linalg.inv(d0["qM"])
    fr0 = fk_frames(root, q0, degrees)
    h = 1e-6
    mv0 = [(fk_frames(root, move(name, q0, i, h), degrees), fk_frames(root, move(name, q0, i, -h), degrees)) for i in range(m.nv)]
    tran = [0.0]                                                   # the world does not move
    for b, (P, R) in enumerate(fr0):
        cl = R.T @ (d0["xipos"][b + 1] - P)
        J = np.stack([((mv0[i][0][b][0] + mv0[i][0][b][1] @ cl) - (mv0[i][1][b][0] + mv0[i][1][b][1] @ cl)) / (2 * h) for i in range(m.nv)], 1)
        tran.append(float(np.trace(J @ Minv @ J.T)) / 3)
    assert np.array(tran[1:]) == pytest.approx(m.field("body_invweight0").reshape(-1, 2)[1:m.nbody, 0], rel=2e-6)   # and the compiled tables agree

    rng = np.random.default_rng(44)
    rngs = m.field("jnt_range").reshape(-1, 2)[:m.njnt]
    seen = {"limit": 0, "ground": 0, "pair": 0}
    for trial in range(30):
        q = q0.copy()
        if name == "humanoid":
            q[2] = rng.uniform(0.2, 0.8)
            w = rng.normal(0, 1.0, 3); ang = np.linalg.norm(w)
            q[3:7] = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])
            q[7:] = rng.uniform(rngs[1:, 0], rngs[1:, 1])
            lo_, hi_, qs = rngs[1:, 0], rngs[1:, 1], q[7:]
        else:
            q[1] = rng.uniform(0.3, 1.2); q[2] = rng.uniform(-1.5, 1.5)
            q[3:] = rng.uniform(rngs[3:, 0], rngs[3:, 1])
            lo_, hi_, qs = rngs[3:, 0], rngs[3:, 1], q[3:]
        for k in range(len(qs)):                                   # a quarter of the joints up to 0.02 rad beyond one of their limits
            if rng.uniform() < 0.25:
                qs[k] = lo_[k] - rng.uniform(0, 0.02) if rng.uniform() < 0.5 else hi_[k] + rng.uniform(0, 0.02)
        v = rng.normal(0, 1.0, m.nv)
        d = oracle.dump(om, q, v, np.zeros(m.nu))
        nrows = [4 if -1 in (gt[a][0], gt[b_][0]) else 1 for a, b_ in d["contact_geom"]]
        nlim = d["nefc"] - sum(nrows)
        assert nlim >= 0

        def check(r, solref, solimp, margin, pos, diag, scale, what):
            tc = max(solref[0], 2 * dt)
            K = 1.0 / (solimp[1] ** 2 * tc ** 2 * solref[1] ** 2); B = 2.0 / (solimp[1] * tc)
            imp = impedance(solimp, pos - margin)
            R = scale * (1 - imp) / imp * diag
            assert d["efc_pos"][r] == pytest.approx(pos, abs=1e-10), what
            assert d["efc_margin"][r] == pytest.approx(margin, abs=1e-15), what
            assert d["efc_R"][r] == pytest.approx(R, rel=3e-6), what
            assert d["efc_D"][r] == pytest.approx(1 / R, rel=3e-6), what
            assert d["efc_aref"][r] == pytest.approx(-B * (d["efc_J"][r] @ v) - K * imp * (pos - margin), rel=1e-9, abs=1e-9), what

        for r in range(nlim):                                      # joint limits: one entry +-1 says which joint and which side
            Jr = d["efc_J"][r]
            i = int(np.abs(Jr).argmax())
            assert abs(Jr[i]) == 1.0 and np.count_nonzero(Jr) == 1
            lim, rg, (solref, solimp, margin) = jinfo[i]
            assert lim
            qi = q[i + 1] if name == "humanoid" else q[i]
            pos = qi - rg[0] if Jr[i] > 0 else rg[1] - qi          # lower limit pushes up (+1), upper pushes down (-1)
            assert pos < margin + 1e-12                            # the row exists because the joint is within its margin of the limit
            check(r, solref, solimp, margin, pos, Minv[i, i], 1.0, ("limit", trial, i))
            seen["limit"] += 1
        first = nlim + np.concatenate([[0], np.cumsum(nrows)[:-1]]).astype(int) if nrows else []
        for c, (g1, g2) in enumerate(d["contact_geom"]):
            (sr1, si1, mg1), (sr2, si2, mg2) = gattr[g1], gattr[g2]
            assert sr1 == sr2 and si1 == si2                       # every geom of these files carries its file's one parameter set: no mixing to do
            margin = max(mg1, mg2)
            diag = tran[gt[g1][0] + 1] + tran[gt[g2][0] + 1]
            if nrows[c] == 4:
                mu = max(gt[g1][1], gt[g2][1])
                for k in range(4):
                    check(first[c] + k, sr1, si1, margin, d["contact_dist"][c], diag * (1 + mu * mu), 2 * mu * mu, ("ground", trial, c, k))
                seen["ground"] += 1
            else:
                check(first[c], sr1, si1, margin, d["contact_dist"][c], diag, 1.0, ("pair", trial, c))
                seen["pair"] += 1
    assert seen["limit"] >= 10 and seen["ground"] >= 15 and seen["pair"] >= 3
