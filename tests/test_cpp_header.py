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


def _build_parity(tmp_path):
    from oracle import oracle_py
    so = oracle_py.build()
    exe = tmp_path / "parity_test"
    lib = M.lib_path()
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "cpp", "parity_test.cpp"),
                           "-o", str(exe), lib, so, "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.dirname(so)])
    return exe


def test_cpp_parity_test_compiles(tmp_path):
    exe = _build_parity(tmp_path)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2 and "usage" in out.stdout


import pytest  # noqa: E402


@pytest.mark.gpu
def test_cpp_parity_on_gpu(tmp_path, scenes):
    """The reference-style C++ test: drop-in scenario classes vs the oracle, same seeded inputs."""
    import numpy as np
    from mplambda_b200.workload import SE3_SCENES, SEED, fetch_configs, se3_poses
    exe = _build_parity(tmp_path)
    env, robot = scenes["twistycool"]
    sc = SE3_SCENES["twistycool"]
    P = se3_poses(20000, sc["min"], sc["max"], SEED + 31)
    s = M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], 0.1)
    A = P[s.check_states(P)][:400]
    B = se3_poses(len(A), sc["min"], sc["max"], SEED + 32)
    files = {"env": env, "robot": robot, "states": P, "from": A, "to": B, "autolab": scenes["autolab"][0],
             "fetch": fetch_configs(20000, SEED + 33)}
    paths = []
    for k in ("env", "robot", "states", "from", "to", "autolab", "fetch"):
        p = tmp_path / (k + ".bin")
        np.ascontiguousarray(files[k], dtype=np.float64).tofile(p)
        paths.append(str(p))
    out = subprocess.run([str(exe)] + paths, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "PASSED" in out.stdout, out.stdout + out.stderr
