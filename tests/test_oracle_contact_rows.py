"""Independent anchor of the oracle's CONTACT ROWS: every ground contact of hopper and humanoid at random folded poses.

tests/test_oracle_contact_anchors.py re-derives the hopper foot's pyramid rows by hand at one state.  Here, for every plane contact the
oracle reports at 30 random poses of both models, the four pyramid rows of efc_J are rebuilt from nothing but the TEXT of the MJCF file:
  * the body frames by the forward kinematics of tests/test_oracle_first_principles.py (composed from the XML's body / joint records);
  * the contact point's velocity Jacobian by NUMERICAL differentiation of that kinematics: the point fixed in the geom's body that
    sits at the contact position, moved by +-h along every dof — in the tangent space for the humanoid's free joint (translations in
    the world frame, rotations about body axes: MuJoCo's convention for a free joint's qvel, the one mju_quatIntegrate implements);
  * friction = max of the two geoms' `friction` attributes (MuJoCo's rule; default 1), condim 3 from the floor;
  * rows J_n + mu J_t1, J_n - mu J_t1, J_n + mu J_t2, J_n - mu J_t2 in the contact frame the oracle reports (the frame itself is
    checked for normal = floor normal, orthonormality and handedness), joint-limit rows first;
  * for body-body pairs (self-collisions of the folded poses, condim 1) the single row n . (v_2 - v_1) at the contact position.
The oracle builds the same rows from cdof (motion axes about the tree's centre of mass): no shared code, no shared formula.
(The XML files are only present in the authoring container: the tests skip elsewhere.)"""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from test_oracle_first_principles import RES, num, quat_mat, rodrigues, vec, walk


def fk_frames(root, qpos, degrees):
    """(P, R) of every body at qpos from the XML's body and joint records (the kinematics of fk_world_coms without the mass
    properties: this one is called 2 nv times per pose)."""
    global_coords = root.find("compiler").get("coordinate", "local") == "global"
    out = []
    qi = [0]

    def visit(body, Pp, Rp, Pp0):
        pos = vec(body.get("pos", "0 0 0"))
        quat = vec(body.get("quat", "1 0 0 0"))
        if global_coords:
            P0 = pos.copy()
            P, R = Pp + Rp @ (P0 - Pp0), Rp.copy()
        else:
            P0 = None
            P, R = Pp + Rp @ pos, Rp @ quat_mat(quat)
        for j in [j for j in body if j.tag in ("joint", "freejoint")]:
            ty = "free" if j.tag == "freejoint" else j.get("type", "hinge")
            if ty == "free":
                P, R = qpos[qi[0]:qi[0] + 3].copy(), quat_mat(qpos[qi[0] + 3:qi[0] + 7])
                qi[0] += 7
                continue
            axis = vec(j.get("axis", "0 0 1"))
            jp = vec(j.get("pos", "0 0 0"))
            if global_coords:
                jp = jp - P0
            ref = num(j.get("ref", "0"))
            if ty == "hinge" and degrees:
                ref = np.deg2rad(ref)
            q = qpos[qi[0]] - ref
            qi[0] += 1
            if ty == "slide":
                P = P + R @ (axis / np.linalg.norm(axis)) * q
            else:
                anchor = P + R @ jp
                R = R @ rodrigues(axis, q)
                P = anchor - R @ jp
        out.append((P, R))
        for ch in body.findall("body"):
            visit(ch, P, R, P0 if global_coords else None)

    for b in root.find("worldbody").findall("body"):
        visit(b, np.zeros(3), np.eye(3), np.zeros(3))
    return out


def geom_table(root):
    """(body index or -1 for the world, friction) of every geom in the model's geom order"""
    out = [(-1, vec(g.get("friction", "1"))[0]) for g in root.find("worldbody").findall("geom")]
    bodies = []
    for b in root.find("worldbody").findall("body"):
        walk(b, bodies)
    for bi, body in enumerate(bodies):
        for g in body.findall("geom"):
            out.append((bi, vec(g.get("friction", "1"))[0]))
    return out


def qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def move(name, q, i, h):
    """qpos after a step h along dof i (tangent space)"""
    q = q.copy()
    if name == "humanoid":
        if i < 3:
            q[i] += h
        elif i < 6:
            e = np.zeros(3); e[i - 3] = 1.0
            q[3:7] = qmul(q[3:7], np.concatenate([[np.cos(h / 2)], np.sin(h / 2) * e]))     # rotation about a BODY axis
        else:
            q[i + 1] += h
    else:
        q[i] += h
    return q


@pytest.mark.parametrize("name", ["hopper", "humanoid"])
def test_plane_contact_rows_against_numerical_kinematics(oracle, omodels, pkg, name):
    path = os.path.join(RES, name + ".xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    degrees = root.find("compiler").get("angle", "degree") == "degree"
    om = omodels[name]
    m = pkg.Model.named(name)
    gt = geom_table(root)
    assert len(gt) == m.ngeom
    rng = np.random.default_rng(33)
    rngs = m.field("jnt_range").reshape(-1, 2)[:m.njnt]
    h = 1e-6
    checked = pairs = 0
    for trial in range(30):
        q = m.field("qpos0")[:m.nq].copy()
        if name == "humanoid":
            q[2] = rng.uniform(0.2, 0.8)
            w = rng.normal(0, 1.0, 3); ang = np.linalg.norm(w)
            q[3:7] = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])
            q[7:] = rng.uniform(rngs[1:, 0], rngs[1:, 1])
        else:
            q[1] = rng.uniform(0.3, 1.2); q[2] = rng.uniform(-1.5, 1.5)
            q[3:] = rng.uniform(rngs[3:, 0], rngs[3:, 1])
        d = oracle.dump(om, q, np.zeros(m.nv), np.zeros(m.nu))
        if d["ncon"] == 0:
            continue
        # joint-limit rows first, then the contacts in order: four pyramid rows where the floor is involved (condim = max of the two geoms':
        # 3 from the floor), one frictionless row for a body-body pair (condim 1 on every other geom of both files)
        nrows = [4 if -1 in (gt[a][0], gt[b_][0]) else 1 for a, b_ in d["contact_geom"]]
        nlim = d["nefc"] - sum(nrows)
        assert nlim >= 0
        first = nlim + np.concatenate([[0], np.cumsum(nrows)[:-1]]).astype(int)
        frames = fk_frames(root, q, degrees)
        moved = [(fk_frames(root, move(name, q, i, h), degrees), fk_frames(root, move(name, q, i, -h), degrees)) for i in range(m.nv)]
        for c in range(d["ncon"]):
            g1, g2 = d["contact_geom"][c]
            if gt[g1][0] != -1 and gt[g2][0] != -1:
                # body-body pair (folded limbs): ONE frictionless row = normal component of the velocity of geom 2's point relative to
                # geom 1's at the contact position, the normal pointing from geom 1 to geom 2
                n = d["contact_frame"][c][:3]
                assert np.linalg.norm(n) == pytest.approx(1.0, abs=1e-12)
                pos = d["contact_pos"][c]
                rel = np.zeros((3, m.nv))
                for gi, sg in ((g2, 1.0), (g1, -1.0)):
                    bb = gt[gi][0]
                    P, R = frames[bb]
                    pl = R.T @ (pos - P)
                    for i in range(m.nv):
                        (Pa, Ra), (Pb, Rb) = moved[i][0][bb], moved[i][1][bb]
                        rel[:, i] += sg * ((Pa + Ra @ pl) - (Pb + Rb @ pl)) / (2 * h)
                want1 = n @ rel
                got1 = d["efc_J"][first[c]]
                assert np.abs(got1 - want1).max() < 2e-8 * max(1.0, np.abs(want1).max()), (trial, c, "pair", np.abs(got1 - want1).max())
                pairs += 1
                continue
            sign = 1.0
            if gt[g1][0] != -1:                       # (the plane is geom 1 in MuJoCo's order; keep the other order honest anyway)
                g1, g2, sign = g2, g1, -1.0
            b = gt[g2][0]
            mu = max(gt[g1][1], gt[g2][1])
            pos = d["contact_pos"][c]
            fr = d["contact_frame"][c].reshape(3, 3)
            n, t1, t2 = fr
            assert n == pytest.approx([0, 0, 1], abs=1e-12)
            assert fr @ fr.T == pytest.approx(np.eye(3), abs=1e-12) and np.cross(n, t1) == pytest.approx(t2, abs=1e-12)
            P, R = frames[b]
            p_local = R.T @ (pos - P)                 # the point of the geom's body that sits at the contact position
            Jp = np.zeros((3, m.nv))
            for i in range(m.nv):
                (Pa, Ra), (Pb, Rb) = moved[i][0][b], moved[i][1][b]
                Jp[:, i] = ((Pa + Ra @ p_local) - (Pb + Rb @ p_local)) / (2 * h)
            Jp *= sign                                # velocity of geom 2's point relative to geom 1's (the floor does not move)
            jn, ja, jb = n @ Jp, t1 @ Jp, t2 @ Jp
            want = np.stack([jn + mu * ja, jn - mu * ja, jn + mu * jb, jn - mu * jb])
            got = d["efc_J"][first[c]: first[c] + 4]
            assert np.abs(got - want).max() < 2e-8 * max(1.0, np.abs(want).max()), (trial, c, np.abs(got - want).max())
            assert (d["efc_pos"][first[c]: first[c] + 4] == d["contact_dist"][c]).all()
            checked += 1
    assert checked >= (20 if name == "hopper" else 40)
    assert pairs >= 5            # self-collisions of the folded poses, both models
