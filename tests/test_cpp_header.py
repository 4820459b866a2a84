"""CPU: the C++ drop-in header compiles with plain g++ against libmplb.so and keeps the
reference's error behaviour (std::invalid_argument for an unloadable mesh)."""
import os
import subprocess

import mplambda_b200 as M
from conftest import ROOT, REFERENCE


def test_header_compiles_and_runs(tmp_path):
    exe = tmp_path / "scenario_smoke"
    lib = M.lib_path()
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "scenario_smoke.cpp"), "-o", str(exe), lib,
                           "-Wl,-rpath," + os.path.dirname(lib)])
    args = [str(exe)]
    res = os.path.join(REFERENCE, "resources", "se3")
    if os.path.isdir(res):
        args += [os.path.join(res, "Twistycool_env.dae"), os.path.join(res, "Twistycool_robot.dae")]
    out = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "invalid_argument: could not load mesh file" in out.stdout
