"""Anchors of the compiled model tables and of the oracle's smooth dynamics that share NO code and no formula with either.

The model tables (ilqg-mujoco_b200/models/*.ilqgm) are read by the CUDA kernels AND by the CPU oracle, so an error in them — or in what
both sides believe MuJoCo computes from them — would be common-mode: GPU == oracle would stay green.  Everything below starts from the
TEXT of /root/reference/res/hopper.xml and humanoid.xml and from textbook mechanics:
  * every body's mass, centre of mass and inertia tensor by NUMERICAL QUADRATURE of a uniform density (1000 kg/m^3, MuJoCo's default) over
    its capsules and spheres (Gauss-Legendre in cylindrical / spherical coordinates: exact for these polynomial integrands, no
    closed-form capsule inertia involved; overlapping geoms count twice, as in MuJoCo) against body_mass / body_ipos / body_inertia;
  * a forward kinematics composed from the XML's body and joint records against the oracle's xpos / xquat / xipos / subtree_com;
  * the mass matrix as the Hessian of the kinetic energy sum_b 1/2 m |v_com|^2 + 1/2 w' I w (body velocities by numerical
    differentiation of that kinematics) plus armature, against the oracle's qM (CRBA);
  * the hopper's bias forces from Lagrange's equations on those energies against qfrc_bias (RNE);
  * passive and actuator forces from the XML's damping / stiffness / gear / ctrlrange attributes.
(The XML files are only present in the authoring container: the tests skip elsewhere.)"""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

RES = "/root/reference/res"
DENSITY = 1000.0
GL_X, GL_W = np.polynomial.legendre.leggauss(12)


def gl(a, b):
    return 0.5 * (b - a) * GL_X + 0.5 * (b + a), 0.5 * (b - a) * GL_W


def moments_capsule(radius, half):
    """mass, second-moment matrix int x x' dm of a capsule along z, centred at the origin: cylinder + two hemispheres."""
    m, S = 0.0, np.zeros((3, 3))
    th, wth = gl(0.0, 2 * np.pi)
    # cylinder: x = (r cos t, r sin t, z), dV = r dr dt dz
    r, wr = gl(0.0, radius)
    z, wz = gl(-half, half)
    for ri, wri in zip(r, wr):
        for ti, wti in zip(th, wth):
            for zi, wzi in zip(z, wz):
                x = np.array([ri * np.cos(ti), ri * np.sin(ti), zi])
                dm = DENSITY * ri * wri * wti * wzi
                m += dm
                S += dm * np.outer(x, x)
    # hemispheres: x = (s sin p cos t, s sin p sin t, +-(half + s cos p)), p in [0, pi/2], dV = s^2 sin p ds dp dt
    s, ws = gl(0.0, radius)
    p, wp = gl(0.0, np.pi / 2)
    for sign in (1.0, -1.0):
        for si, wsi in zip(s, ws):
            for pi_, wpi in zip(p, wp):
                for ti, wti in zip(th, wth):
                    x = np.array([si * np.sin(pi_) * np.cos(ti), si * np.sin(pi_) * np.sin(ti), sign * (half + si * np.cos(pi_))])
                    dm = DENSITY * si * si * np.sin(pi_) * wsi * wpi * wti
                    m += dm
                    S += dm * np.outer(x, x)
    return m, S


def frame_with_z(axis):
    z = axis / np.linalg.norm(axis)
    a = np.array([1.0, 0, 0]) if abs(z[0]) < 0.9 else np.array([0, 1.0, 0])
    x = np.cross(a, z); x /= np.linalg.norm(x)
    return np.stack([x, np.cross(z, x), z], axis=1)    # columns: the capsule frame's axes in the parent frame


def num(s):
    """MuJoCo reads numbers with strtod: '0.13/2' is 0.13 (hopper.xml's foot; a documented quirk)."""
    import re
    return float(re.match(r"\s*[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?", s).group(0))


def vec(s):
    return np.array([num(t) for t in s.split()])


_MOMENTS = {}


def body_moments(body):
    """(mass, first moment, second-moment matrix) of the body's own geoms in the coordinates the XML gives them in."""
    key = ET.tostring(body)[:4096]          # (memoised: the kinematics below asks for the same bodies thousands of times)
    if key not in _MOMENTS:
        _MOMENTS[key] = _body_moments(body)
    return _MOMENTS[key]


def _body_moments(body):
    M, F, S = 0.0, np.zeros(3), np.zeros((3, 3))
    for g in body.findall("geom"):
        ty = g.get("type", "sphere")
        size = vec(g.get("size"))
        if ty == "capsule":
            a, b = vec(g.get("fromto"))[:3], vec(g.get("fromto"))[3:]
            c, R = 0.5 * (a + b), frame_with_z(b - a)
            m, S0 = moments_capsule(size[0], 0.5 * np.linalg.norm(b - a))
        elif ty == "sphere":
            c, R = vec(g.get("pos", "0 0 0")), np.eye(3)
            m, S0 = moments_capsule(size[0], 0.0)
        else:
            raise AssertionError(ty)
        S0 = R @ S0 @ R.T
        M += m
        F += m * c
        S += S0 + m * np.outer(c, c)       # (the geom's own first moment about its centre is zero)
    return M, F, S


def walk(body, out):
    out.append(body)
    for ch in body.findall("body"):
        walk(ch, out)


@pytest.mark.parametrize("name", ["hopper", "humanoid"])
def test_body_mass_com_and_inertia_by_quadrature(pkg, name):
    path = os.path.join(RES, name + ".xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    global_coords = root.find("compiler").get("coordinate", "local") == "global"
    bodies = []
    for b in root.find("worldbody").findall("body"):
        walk(b, bodies)
    m = pkg.Model.named(name)
    assert m.nbody == len(bodies) + 1
    mass, ipos, inertia = m.field("body_mass"), m.field("body_ipos").reshape(-1, 3), m.field("body_inertia").reshape(-1, 6)
    for k, body in enumerate(bodies, start=1):
        M, F, S = body_moments(body)
        com = F / M
        Sc = S - M * np.outer(com, com)
        I = np.trace(Sc) * np.eye(3) - Sc
        if global_coords:      # geoms and body frames are given in world coordinates (no rotated body frame in this model)
            com = com - vec(body.get("pos"))
        assert mass[k] == pytest.approx(M, rel=1e-11), body.get("name")
        assert ipos[k] == pytest.approx(com, abs=1e-11), body.get("name")
        want = np.array([I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]])
        assert np.abs(inertia[k] - want).max() <= 1e-10 * np.abs(want).max(), (body.get("name"), inertia[k], want)


# ------------------------------------------------------------------ kinematics, from the MJCF text again
def quat_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def quat_mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rodrigues(axis, ang):
    a = axis / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def fk_world_coms(root, qpos, degrees):
    """World position, rotation and centre of mass of every body at qpos, by composing the XML's own body / joint records
    (MuJoCo's kinematics: a body starts at parent o (pos, quat); its joints act in order — a free joint sets the frame, a slide
    translates along its axis, a hinge turns the frame about its anchor)."""
    global_coords = root.find("compiler").get("coordinate", "local") == "global"
    out = []
    qi = [0]

    def visit(body, Pp, Rp, Pp0):
        # frame at the reference configuration
        pos = vec(body.get("pos", "0 0 0"))
        quat = vec(body.get("quat", "1 0 0 0"))
        if global_coords:
            P0, R0 = pos.copy(), np.eye(3)                   # given in world coordinates at qpos0 (no rotated frames in hopper.xml)
            P, R = Pp + Rp @ (P0 - Pp0), Rp.copy()           # the same offset, carried by the parent's current frame
        else:
            P0 = None
            P, R = Pp + Rp @ pos, Rp @ quat_mat(quat)
        joints = [j for j in body if j.tag in ("joint", "freejoint")]
        for j in joints:
            ty = "free" if j.tag == "freejoint" else j.get("type", "hinge")
            if ty == "free":
                P, R = qpos[qi[0]:qi[0] + 3].copy(), quat_mat(qpos[qi[0] + 3:qi[0] + 7])
                qi[0] += 7
                continue
            axis = vec(j.get("axis", "0 0 1"))
            jp = vec(j.get("pos", "0 0 0"))
            if global_coords:
                jp = jp - P0                                  # anchor in the body frame
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
        M, F, _ = body_moments(body)
        com_local = F / M - (P0 if global_coords else 0.0)
        out.append((P, R, P + R @ com_local, M))
        for ch in body.findall("body"):
            visit(ch, P, R, P0 if global_coords else None)

    for b in root.find("worldbody").findall("body"):
        visit(b, np.zeros(3), np.eye(3), np.zeros(3))
    return out


@pytest.mark.parametrize("name", ["hopper", "humanoid"])
def test_oracle_kinematics_against_the_mjcf_text(oracle, omodels, pkg, name):
    """The oracle's mj_kinematics / mj_comPos products (xpos, xipos, subtree_com of the whole tree) at random configurations against
    a forward kinematics composed here from the XML's body and joint records and the quadrature centres of mass above: pins
    body_pos / body_quat / jnt_pos / jnt_axis / qpos0 / body_ipos of the compiled tables and the kinematics that reads them."""
    path = os.path.join(RES, name + ".xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    degrees = root.find("compiler").get("angle", "degree") == "degree"
    om = omodels[name]
    m = pkg.Model.named(name)
    rng = np.random.default_rng(12)
    for trial in range(4):
        q = m.field("qpos0")[:m.nq].copy()
        if name == "humanoid":
            q[:3] += rng.uniform(-0.5, 0.5, 3)
            w = rng.normal(0, 0.6, 3); ang = np.linalg.norm(w)
            q[3:7] = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])
            q[7:] += rng.uniform(-0.7, 0.7, m.nq - 7)
        else:
            q += rng.uniform(-0.6, 0.6, m.nq)
        d = oracle.dump(om, q, np.zeros(m.nv), np.zeros(m.nu))
        fk = fk_world_coms(root, q, degrees)
        tot_m, tot_f = 0.0, np.zeros(3)
        for k, (P, R, com, M) in enumerate(fk, start=1):
            assert d["xpos"][k] == pytest.approx(P, abs=1e-12), (trial, k)
            assert quat_mat(d["xquat"][k]) == pytest.approx(R, abs=1e-12), (trial, k)
            assert d["xipos"][k] == pytest.approx(com, abs=1e-11), (trial, k)
            tot_m += M; tot_f += M * com
        assert d["subtree_com"][1] == pytest.approx(tot_f / tot_m, abs=1e-11)     # (body 1 roots the whole mechanism in both models)


# ------------------------------------------------------------------ mass matrix and bias forces from first principles
def body_inertias(root):
    """per body (XML order): inertia tensor about the centre of mass in the body frame, by quadrature"""
    bodies = []
    for b in root.find("worldbody").findall("body"):
        walk(b, bodies)
    out = []
    for body in bodies:
        M, F, S = body_moments(body)
        com = F / M
        Sc = S - M * np.outer(com, com)
        out.append(np.trace(Sc) * np.eye(3) - Sc)
    return out


def advance(name, q, v, eps):
    """the configuration reached from q along the tangent v (MuJoCo's mj_integratePos): free joint = world-frame linear velocity,
    body-frame angular velocity"""
    if name != "humanoid":
        return q + eps * v
    out = q.copy()
    out[:3] += eps * v[:3]
    w = eps * v[3:6]
    ang = np.linalg.norm(w)
    dq = np.array([1.0, 0, 0, 0]) if ang < 1e-300 else np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])
    out[3:7] = quat_mul(q[3:7] / np.linalg.norm(q[3:7]), dq)
    out[7:] += eps * v[6:]
    return out


def kinetic_and_potential(root, name, degrees, inertias, q, v, eps=1e-4):
    """T(q, v) = sum_b 1/2 m |d com / dt|^2 + 1/2 w' (R I R') w  with the body velocities taken by central differences of the forward
    kinematics above along the tangent v, and V(q) = sum_b m g z_com."""
    fp, fm, f0 = (fk_world_coms(root, advance(name, q, v, s * eps), degrees) for s in (1.0, -1.0, 0.0))
    T = V = 0.0
    for (Pp, Rp, cp, M), (Pm, Rm, cm, _), (P0, R0, c0, _), I in zip(fp, fm, f0, inertias):
        vc = (cp - cm) / (2 * eps)
        W = (Rp @ Rm.T - Rm @ Rp.T) / (4 * eps)
        w = np.array([W[2, 1], W[0, 2], W[1, 0]])
        T += 0.5 * M * vc @ vc + 0.5 * w @ (R0 @ I @ R0.T) @ w
        V += M * 9.81 * c0[2]
    return T, V


def armatures(root):
    dflt = root.find("default").find("joint")
    out = []
    bodies = []
    for b in root.find("worldbody").findall("body"):
        walk(b, bodies)
    for body in bodies:
        for j in body:
            if j.tag == "freejoint":
                out += [0.0] * 6
            elif j.tag == "joint":
                out.append(num(j.get("armature", dflt.get("armature", "0") if dflt is not None else "0")))
    return np.array(out)


def mass_matrix_from_energy(root, name, degrees, inertias, q, nv):
    """T is a quadratic form in v: M_ij = T(e_i + e_j) - T(e_i) - T(e_j), M_ii = 2 T(e_i)"""
    E = np.eye(nv)
    Ti = np.array([kinetic_and_potential(root, name, degrees, inertias, q, E[i])[0] for i in range(nv)])
    M = np.diag(2 * Ti)
    for i in range(nv):
        for j in range(i):
            M[i, j] = M[j, i] = kinetic_and_potential(root, name, degrees, inertias, q, E[i] + E[j])[0] - Ti[i] - Ti[j]
    return M


@pytest.mark.parametrize("name", ["hopper", "humanoid"])
def test_mass_matrix_against_the_kinetic_energy_of_the_mjcf_text(oracle, omodels, pkg, name):
    """qM of the oracle (composite rigid body algorithm on the compiled tables' spatial inertias and motion axes) against the Hessian of
    the kinetic energy, where the kinetic energy is the textbook sum over bodies — velocities by numerical differentiation of the forward
    kinematics composed from the XML, inertia tensors by quadrature — plus the joints' armature.  No spatial algebra, no table."""
    path = os.path.join(RES, name + ".xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    degrees = root.find("compiler").get("angle", "degree") == "degree"
    om = omodels[name]
    m = pkg.Model.named(name)
    inertias = body_inertias(root)
    rng = np.random.default_rng(5)
    q = m.field("qpos0")[:m.nq].copy()
    if name == "humanoid":
        w = rng.normal(0, 0.5, 3); ang = np.linalg.norm(w)
        q[3:7] = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])
        q[7:] += rng.uniform(-0.6, 0.6, m.nq - 7)
    else:
        q += rng.uniform(-0.6, 0.6, m.nq)
    d = oracle.dump(om, q, np.zeros(m.nv), np.zeros(m.nu))
    M = mass_matrix_from_energy(root, name, degrees, inertias, q, m.nv) + np.diag(armatures(root))
    assert np.abs(d["qM"] - M).max() <= 2e-7 * np.abs(M).max(), np.abs(d["qM"] - M).max()


def test_hopper_bias_forces_against_lagranges_equations(oracle, omodels, pkg):
    """qfrc_bias of the oracle (recursive Newton-Euler: Coriolis, centrifugal and gravity terms) against Lagrange's equations on the
    energies above:  bias_i = sum_j d(M v)_i / dq_j v_j - 1/2 v' dM/dq_i v + dV/dq_i,  with M(q) and V(q) differentiated numerically."""
    name = "hopper"
    path = os.path.join(RES, name + ".xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    om = omodels[name]
    m = pkg.Model.named(name)
    inertias = body_inertias(root)
    rng = np.random.default_rng(6)
    q = m.field("qpos0")[:m.nq].copy() + rng.uniform(-0.5, 0.5, m.nq)
    v = rng.normal(0, 1.5, m.nv)
    nv, h = m.nv, 1e-4
    Mq = lambda qq: mass_matrix_from_energy(root, name, True, inertias, qq, nv)       # (armature is constant: drops out of dM/dq)
    Vq = lambda qq: kinetic_and_potential(root, name, True, inertias, qq, np.zeros(nv))[1]
    dM = np.zeros((nv, nv, nv)); dV = np.zeros(nv)
    for k in range(nv):
        e = np.zeros(nv); e[k] = h
        dM[k] = (Mq(q + e) - Mq(q - e)) / (2 * h)
        dV[k] = (Vq(q + e) - Vq(q - e)) / (2 * h)
    bias = np.einsum("jik,k,j->i", dM, v, v) - 0.5 * np.einsum("ijk,j,k->i", dM, v, v) + dV
    d = oracle.dump(om, q, v, np.zeros(m.nu))
    assert np.abs(d["qfrc_bias"] - bias).max() <= 1e-5 * max(1.0, np.abs(bias).max()), (d["qfrc_bias"], bias)


@pytest.mark.parametrize("name", ["hopper", "humanoid"])
def test_passive_forces_against_the_mjcf_text(oracle, omodels, pkg, name):
    """qfrc_passive = -damping v - stiffness (q - springref), joint by joint, with damping / stiffness read from the XML (defaults class
    first, the joint's own attributes over it)."""
    path = os.path.join(RES, name + ".xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    dflt = root.find("default").find("joint")
    bodies = []
    for b in root.find("worldbody").findall("body"):
        walk(b, bodies)
    damp, stiff, qadr = [], [], []
    qi = 0
    for body in bodies:
        for j in body:
            if j.tag == "freejoint":
                damp += [0.0] * 6; stiff += [0.0] * 6; qadr += [-1] * 6; qi += 7
            elif j.tag == "joint":
                damp.append(num(j.get("damping", dflt.get("damping", "0"))))
                stiff.append(num(j.get("stiffness", dflt.get("stiffness", "0"))))
                qadr.append(qi); qi += 1
    om = omodels[name]
    m = pkg.Model.named(name)
    rng = np.random.default_rng(8)
    q = m.field("qpos0")[:m.nq].copy()
    q[7 if name == "humanoid" else 0:] += rng.uniform(-0.5, 0.5, m.nq - (7 if name == "humanoid" else 0))
    v = rng.normal(0, 1.0, m.nv)
    d = oracle.dump(om, q, v, np.zeros(m.nu))
    want = np.array([-damp[i] * v[i] - (stiff[i] * q[qadr[i]] if qadr[i] >= 0 else 0.0) for i in range(m.nv)])   # springref = 0 throughout
    assert d["qfrc_passive"] == pytest.approx(want, abs=1e-12)


@pytest.mark.parametrize("name", ["hopper", "humanoid"])
def test_actuator_forces_against_the_mjcf_text(oracle, omodels, pkg, name):
    """qfrc_actuator = gear * clamp(ctrl, ctrlrange) on the motor's joint, from the XML's <actuator> section (defaults class first)."""
    path = os.path.join(RES, name + ".xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    dflt = root.find("default").find("motor")
    bodies = []
    for b in root.find("worldbody").findall("body"):
        walk(b, bodies)
    dof_of, di = {}, 0
    for body in bodies:
        for j in body:
            if j.tag == "freejoint":
                di += 6
            elif j.tag == "joint":
                dof_of[j.get("name")] = di; di += 1
    om = omodels[name]
    m = pkg.Model.named(name)
    rng = np.random.default_rng(9)
    u = rng.uniform(-1.5, 1.5, m.nu)
    want = np.zeros(m.nv)
    motors = root.find("actuator").findall("motor")
    assert len(motors) == m.nu
    for a, mot in enumerate(motors):
        lim = (mot.get("ctrllimited", dflt.get("ctrllimited", "false")) == "true")
        lo, hi = vec(mot.get("ctrlrange", dflt.get("ctrlrange", "0 0")))
        c = min(max(u[a], lo), hi) if lim else u[a]
        want[dof_of[mot.get("joint")]] += num(mot.get("gear", "1")) * c
    d = oracle.dump(om, m.field("qpos0")[:m.nq].copy(), np.zeros(m.nv), u)
    assert d["qfrc_actuator"] == pytest.approx(want, abs=1e-12)


# ------------------------------------------------------------------ narrow phase: distances by brute force
def world_geoms(root, q, degrees):
    """every geom in world coordinates at q: ('plane',) | ('sphere', centre, r) | ('capsule', end a, end b, r), in the model's geom order
    (the worldbody's own geoms first, then body by body)"""
    global_coords = root.find("compiler").get("coordinate", "local") == "global"
    fk = fk_world_coms(root, q, degrees)
    bodies = []
    for b in root.find("worldbody").findall("body"):
        walk(b, bodies)
    out = [("plane",) for g in root.find("worldbody").findall("geom") if g.get("type") == "plane"]
    for body, (P, R, _, _) in zip(bodies, fk):
        P0 = vec(body.get("pos", "0 0 0")) if global_coords else np.zeros(3)
        for g in body.findall("geom"):
            size = vec(g.get("size"))
            if g.get("type", "sphere") == "capsule":
                ft = vec(g.get("fromto"))
                out.append(("capsule", P + R @ (ft[:3] - P0), P + R @ (ft[3:] - P0), size[0]))
            else:
                out.append(("sphere", P + R @ (vec(g.get("pos", "0 0 0")) - P0), size[0]))
    return out


def segment_points(g):
    return (g[1], g[1]) if g[0] == "sphere" else (g[1], g[2])


def brute_force_distance(g1, g2):
    """signed distance between two spheres / capsules: minimum over a dense grid of the two axis parameters, refined on a finer grid
    around the best cell (the squared distance is a convex quadratic in (s, t): no local minima to miss)"""
    (a1, b1), (a2, b2) = segment_points(g1), segment_points(g2)
    lo = np.zeros(2); hi = np.ones(2)
    best = None
    for _ in range(6):
        s = np.linspace(lo[0], hi[0], 41); t = np.linspace(lo[1], hi[1], 41)
        p1 = a1[None, :] + s[:, None] * (b1 - a1)[None, :]
        p2 = a2[None, :] + t[:, None] * (b2 - a2)[None, :]
        d = np.linalg.norm(p1[:, None, :] - p2[None, :, :], axis=2)
        i, j = np.unravel_index(d.argmin(), d.shape)
        best = (d[i, j], p1[i], p2[j])
        ws, wt = (hi[0] - lo[0]) / 40, (hi[1] - lo[1]) / 40
        lo = np.array([max(0.0, s[i] - ws), max(0.0, t[j] - wt)]); hi = np.array([min(1.0, s[i] + ws), min(1.0, t[j] + wt)])
    d, p1, p2 = best
    return d - g1[-1] - g2[-1], p1, p2


@pytest.mark.parametrize("name", ["hopper", "humanoid"])
def test_contact_distances_against_brute_force_geometry(oracle, omodels, pkg, name):
    """Every contact the oracle's narrow phase reports (plane-sphere, plane-capsule, sphere / capsule pairs) at random folded poses near
    the ground: its distance against geometry done the slow way on the geoms placed by the XML forward kinematics above — the end
    spheres' heights for a plane, a dense-grid minimum over both axis parameters for a pair — and its position / normal against the
    witness points.  Pins geom placement, the pair distances and MuJoCo's 'midpoint, normal from geom 1 to geom 2' convention as the
    oracle restates it."""
    path = os.path.join(RES, name + ".xml")
    if not os.path.exists(path):
        pytest.skip("reference MJCF files not present on this machine")
    root = ET.parse(path).getroot()
    degrees = root.find("compiler").get("angle", "degree") == "degree"
    om = omodels[name]
    m = pkg.Model.named(name)
    rng = np.random.default_rng(21)
    rngs = m.field("jnt_range").reshape(-1, 2)[:m.njnt]
    seen = {"plane": 0, "pair": 0}
    for trial in range(40):
        q = m.field("qpos0")[:m.nq].copy()
        if name == "humanoid":
            q[2] = rng.uniform(0.25, 0.9)
            w = rng.normal(0, 1.0, 3); ang = np.linalg.norm(w)
            q[3:7] = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / ang])
            q[7:] = rng.uniform(rngs[1:, 0], rngs[1:, 1])          # anywhere inside the joint ranges: folded limbs touch
        else:
            q[1] = rng.uniform(0.3, 1.2); q[2] = rng.uniform(-1.5, 1.5)
            q[3:] = rng.uniform(rngs[3:, 0], rngs[3:, 1])
        d = oracle.dump(om, q, np.zeros(m.nv), np.zeros(m.nu))
        geoms = world_geoms(root, q, degrees)
        assert len(geoms) == m.ngeom
        for c in range(d["ncon"]):
            g1, g2 = d["contact_geom"][c]
            dist, pos, nrm = d["contact_dist"][c], d["contact_pos"][c], d["contact_frame"][c][:3]
            A, B = geoms[g1], geoms[g2]
            if A[0] == "plane":
                ends = [B[1]] if B[0] == "sphere" else [B[1], B[2]]
                e = min(ends, key=lambda p: np.linalg.norm(p[:2] - pos[:2]))      # the end this contact belongs to
                assert dist == pytest.approx(e[2] - B[-1], abs=1e-10), (trial, c)
                assert nrm == pytest.approx([0, 0, 1], abs=1e-12)
                assert pos == pytest.approx([e[0], e[1], e[2] - B[-1] - 0.5 * dist], abs=1e-10)    # midway between the surfaces
                seen["plane"] += 1
            else:
                want, p1, p2 = brute_force_distance(A, B)
                ax1 = A[2] - A[1] if A[0] == "capsule" else None
                ax2 = B[2] - B[1] if B[0] == "capsule" else None
                if ax1 is not None and ax2 is not None and abs(ax1 @ ax2) > 0.999 * np.linalg.norm(ax1) * np.linalg.norm(ax2):
                    continue                                                       # (near-parallel axes: several contacts, other convention)
                assert dist == pytest.approx(want, abs=2e-7), (trial, c, g1, g2)
                if np.linalg.norm(p2 - p1) < 1e-3:
                    continue                                                       # (crossing axes: the direction between the witness points is noise)
                n = (p2 - p1) / np.linalg.norm(p2 - p1)
                assert nrm == pytest.approx(n, abs=2e-5), (trial, c)
                s1 = p1 + n * A[-1]; s2 = p2 - n * B[-1]
                assert pos == pytest.approx(0.5 * (s1 + s2), abs=2e-6), (trial, c)
                seen["pair"] += 1
    assert seen["plane"] > 10
    if name == "humanoid":
        assert seen["pair"] > 5
