"""Every script under tools/ (profiling drivers, fixture generators, summaries), bench.py and __graft_entry__.py at least compiles, and
the pin-ready MuJoCo fixture generator answers --help without the `mujoco` package (it must stay one command away from use)."""
import glob
import os
import py_compile
import subprocess
import sys

from conftest import ROOT


def test_every_tool_compiles(tmp_path):
    files = sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    assert len(files) > 20
    for f in files:
        py_compile.compile(f, cfile=str(tmp_path / (os.path.basename(f) + "c")), doraise=True)


def test_mujoco_fixture_generator_is_one_command_away():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mujoco_fixtures.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "--res" in r.stdout
