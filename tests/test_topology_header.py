"""The compile-time topology header (csrc/topo_gen.h) must be what tools/gen_topology.cpp emits for the shipped models,
including the PLANAR_Y flag that selects the statically sparse algebra (csrc/planar.h, dyn.cuh Alg<true>)."""
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def gen(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("gen") / "gen_topology")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "gen_topology.cpp"),
                           os.path.join(ROOT, "ilqg-mujoco_b200", "csrc", "mjcf_compile.cpp"), "-o", exe])
    return exe


def _models(*names):
    return [f"{n}={os.path.join(ROOT, 'ilqg-mujoco_b200', 'models', n + '.ilqgm')}" for n in names]


def test_committed_header_is_current(gen):
    out = subprocess.check_output([gen] + _models("inverted_pendulum", "hopper"), text=True)
    assert out == open(os.path.join(ROOT, "ilqg-mujoco_b200", "csrc", "topo_gen.h")).read()


def test_planar_flag(gen):
    out = subprocess.check_output([gen] + _models("inverted_pendulum", "hopper", "humanoid"), text=True)
    flags = [line.split("=")[1].strip(" ;\n") for line in out.splitlines() if "PLANAR_Y" in line]
    assert flags == ["1", "1", "0"]      # slide + hinge about y in the plane; the humanoid has a free joint
