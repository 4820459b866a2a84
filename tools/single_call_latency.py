"""Latency of the scalar drop-in calls: isValid(q) = mplb_check_states(n=1), isValid(a,b) = mplb_check_edges(n=1)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
from mplambda_b200 import _capi
import ctypes as C

for name in sys.argv[1:] or ["twistycool", "cubicles"]:
    d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
    s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
    P = se3_poses(4096, sc["min"], sc["max"], SEED)
    ok = s.check_states(P)
    A = np.ascontiguousarray(P[ok][:256]); B = se3_poses(256, sc["min"], sc["max"], SEED + 1)
    v = np.zeros(1, np.uint8)
    L = _capi.lib(); h = s._h
    def t_states(reps=2000):
        t = time.perf_counter()
        for i in range(reps):
            L.mplb_check_states(h, P[i % 4096].ctypes.data_as(C.c_void_p), 1, v.ctypes.data_as(C.c_void_p))
        return (time.perf_counter() - t) / reps * 1e6
    def t_edges(reps=256):
        t = time.perf_counter()
        for i in range(reps):
            L.mplb_check_edges(h, A[i].ctypes.data_as(C.c_void_p), B[i].ctypes.data_as(C.c_void_p), 1, v.ctypes.data_as(C.c_void_p), None)
        return (time.perf_counter() - t) / reps * 1e6
    t_states(200); t_edges(32)
    print("%s: isValid(q) %.1f us/call, isValid(from,to) %.1f us/call (mean steps %.0f)" % (name, t_states(), t_edges(), s.edge_steps(A, B).mean()), flush=True)
