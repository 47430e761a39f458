"""MJCF-subset compiler: the replacement of mj_loadXML (/root/reference/cmd/basic.cpp:123)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

CARTPOLE = """
<mujoco model="t">
  <compiler inertiafromgeom="true"/>
  <default><joint damping="1" limited="true"/><geom contype="0"/></default>
  <option gravity="0 0 -9.81" integrator="RK4" timestep="0.02"/>
  <worldbody>
    <body name="cart" pos="0 0 0">
      <joint axis="1 0 0" name="slider" range="-1 1" type="slide"/>
      <geom quat="0.707 0 0.707 0" size="0.1 0.1" type="capsule"/>
      <body name="pole" pos="0 0 0">
        <joint axis="0 1 0" name="hinge" range="-90 90" type="hinge"/>
        <geom fromto="0 0 0 0.001 0 0.6" size="0.049 0.3" type="capsule"/>
      </body>
    </body>
  </worldbody>
  <actuator><motor gear="100" joint="slider"/></actuator>
</mujoco>
"""


def compile_string(pkg, xml):
    buf = np.zeros(pkg.lib().ilqg_model_sizeof(), dtype=np.uint8)
    err = C.create_string_buffer(512)
    rc = pkg.lib().ilqg_compile_mjcf_string(xml.encode(), buf.ctypes.data_as(C.c_void_p), err, 512)
    return rc, err.value.decode(), buf


def capsule_mass(r, h):
    return 1000 * (math.pi * r * r * 2 * h + 4 / 3 * math.pi * r ** 3)


def test_cartpole_tables(pkg):
    rc, err, buf = compile_string(pkg, CARTPOLE)
    assert rc == 0, err
    m = pkg.Model(buf)
    assert (m.nq, m.nv, m.nu, m.nbody, m.njnt, m.npair) == (2, 2, 1, 3, 2, 0)
    assert m.timestep == 0.02 and m.field("integrator")[0] == 1
    mass = m.field("body_mass")
    assert mass[1] == pytest.approx(capsule_mass(0.1, 0.1), rel=1e-14)
    L = math.hypot(0.001, 0.6)
    assert mass[2] == pytest.approx(capsule_mass(0.049, L / 2), rel=1e-14)
    # hinge range is in degrees by default; slide range is not converted
    rng = m.field("jnt_range").reshape(-1, 2)
    assert rng[0].tolist() == [-1, 1]
    assert rng[1] == pytest.approx([-math.pi / 2, math.pi / 2])
    # MuJoCo's published body masses for this model (gym InvertedPendulum): 10.47197551, 5.01859164
    assert mass[1] == pytest.approx(10.47197551, abs=1e-8)
    assert mass[2] == pytest.approx(5.01859164, abs=1e-8)


def test_humanoid_masses_match_published_mujoco_values(pkg):
    # body_mass of the MuJoCo humanoid as reported by upstream builds (torso, lwaist, pelvis, thigh, shin)
    m = pkg.Model.named("humanoid")
    mass = m.field("body_mass")
    for idx, ref in ((1, 8.90746237), (2, 2.26194671), (3, 6.61619413), (4, 4.75175093), (5, 2.75569617)):
        assert mass[idx] == pytest.approx(ref, abs=1e-8)
    assert (m.nq, m.nv, m.nu, m.npair) == (28, 27, 21, 161)


def test_hopper_global_coordinates_and_pairs(pkg):
    m = pkg.Model.named("hopper")
    assert (m.nq, m.nv, m.nu, m.nbody, m.npair) == (6, 6, 3, 5, 7)
    g1 = m.field("pair_geom1")[:7].tolist(); g2 = m.field("pair_geom2")[:7].tolist()
    # floor against all four capsules (world-child is not filtered), then the three non-adjacent capsule pairs
    assert list(zip(g1, g2)) == [(0, 1), (0, 2), (0, 3), (0, 4), (1, 3), (1, 4), (2, 4)]
    assert m.field("pair_condim")[:7].tolist() == [3, 3, 3, 3, 1, 1, 1]
    assert m.field("pair_friction")[:4].tolist() == [1.0, 1.0, 1.0, 2.0]
    assert m.field("qpos0")[:6].tolist() == [0, 1.25, 0, 0, 0, 0]  # rootz ref=1.25
    # coordinate="global": thigh frame is 0.2 below the torso frame
    assert m.field("body_pos").reshape(-1, 3)[2] == pytest.approx([0, 0, -0.2])
    assert m.field("act_ctrllimited")[:3].tolist() == [1, 1, 1]
    assert m.field("dof_damping")[:6].tolist() == [0, 0, 0, 1, 1, 1]


@pytest.mark.parametrize("snippet,needle", [
    ('<geom type="box" size="1 1 1"/>', "unsupported geom type"),
    ('<geom type="sphere" size="1" condim="4"/>', "condim"),
])
def test_unsupported_features_are_rejected(pkg, snippet, needle):
    xml = CARTPOLE.replace('<body name="cart" pos="0 0 0">', '<body name="cart" pos="0 0 0">' + snippet)
    rc, err, _ = compile_string(pkg, xml)
    assert rc == pkg.ERR_MODEL and needle in err


def test_malformed_xml_and_missing_file(pkg):
    rc, err, _ = compile_string(pkg, "<mujoco><worldbody></mujoco>")
    assert rc == pkg.ERR_MODEL and "XML" in err
    buf = np.zeros(pkg.lib().ilqg_model_sizeof(), dtype=np.uint8)
    e = C.create_string_buffer(256)
    assert pkg.lib().ilqg_compile_mjcf(b"/nonexistent.xml", buf.ctypes.data_as(C.c_void_p), e, 256) == pkg.ERR_IO


@pytest.mark.skipif(not os.path.isdir("/root/reference/res"), reason="reference XML only exists in the authoring container")
@pytest.mark.parametrize("name", ["inverted_pendulum", "hopper", "humanoid"])
def test_committed_tables_are_what_the_compiler_produces(pkg, name):
    m = pkg.Model.from_mjcf(f"/root/reference/res/{name}.xml")
    ref = pkg.Model.named(name)
    assert np.array_equal(m.buf, ref.buf)
