"""Scalar-call throughput with T threads, with and without the combiner: python tools/scalar_calls.py [scene] [threads ...]"""
import os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1] if len(sys.argv) > 1 else "twistycool"
threads = [int(x) for x in sys.argv[2:]] or [1, 4, 16]
d = np.load(os.path.join(root, "tests/golden/scenes/%s.npz" % name)); sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P = se3_poses(4096, sc["min"], sc["max"], SEED)
A = np.ascontiguousarray(P[s.check_states(P)][:512]); B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
tmp = tempfile.mkdtemp()
paths = []
for k, a in (("env", d["env"]), ("robot", d["robot"]), ("states", P), ("from", A), ("to", B)):
    p = os.path.join(tmp, k + ".bin"); np.ascontiguousarray(a, np.float64).tofile(p); paths.append(p)
exe = os.path.join(tmp, "scalar_calls"); lib = M.lib_path()
subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-pthread", "-I", os.path.join(root, "include"),
                       os.path.join(root, "tools/scalar_calls.cpp"), "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)])
if os.environ.get("MPLB_SCALAR_VERIFY"):   # every verdict against the batched call's (itself checked against the oracle by the tests)
    for k, a in (("want_s", s.check_states(P)), ("want_e", s.check_edges(A, B))):
        p = os.path.join(tmp, k + ".bin"); np.ascontiguousarray(a, np.float64).tofile(p); paths.append(p)
    paths = paths[:5] + ["%THREADS%", "1.0", "%CO%"] + paths[5:]
for t in threads:
    for co in (0, 1):
        args = [x.replace("%THREADS%", str(t)).replace("%CO%", str(co)) for x in paths] if "%CO%" in paths else paths + [str(t), "1.0", str(co)]
        print(subprocess.run([exe] + args, capture_output=True, text=True).stdout, end="", flush=True)
