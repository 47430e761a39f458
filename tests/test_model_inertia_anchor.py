"""An anchor for the MJCF compiler's `inertiafromgeom` that shares NO code and no formula with it.

The model tables (ilqg-mujoco_b200/models/*.ilqgm) are read by the CUDA kernels AND by the CPU oracle, so a wrong mass, centre of mass
or inertia tensor in them would be common-mode: GPU == oracle would stay green.  Published MuJoCo body masses pin the masses
(tests/test_mjcf_compile.py); this file pins the first and second moments as well: every body's mass, centre of mass and inertia tensor
are re-derived here by NUMERICAL QUADRATURE of a uniform density (1000 kg/m^3, MuJoCo's default) over the body's capsules and spheres
(Gauss-Legendre in cylindrical / spherical coordinates: exact for these polynomial integrands, no closed-form capsule inertia involved),
straight from the text of /root/reference/res/hopper.xml and humanoid.xml, and compared with the compiled tables.  Overlapping geoms
count twice, as in MuJoCo.  (The XML files are only present in the authoring container: the test skips elsewhere.)"""
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


def body_moments(body):
    """(mass, first moment, second-moment matrix) of the body's own geoms in the coordinates the XML gives them in."""
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
