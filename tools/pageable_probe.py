"""Dev: host calls with pageable (plain numpy) vs pinned inputs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
name = "cubicles"
d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
n = 1 << 20
P = se3_poses(n, sc["min"], sc["max"], SEED)
hv = np.empty(n, np.uint8)
pp = torch.from_numpy(P).pin_memory(); pv = torch.empty(n, dtype=torch.uint8).pin_memory()
def t(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        a = time.perf_counter(); fn(); ts.append(time.perf_counter() - a)
    return min(ts) * 1e3
print("states 2^20: pageable %.3f ms, pinned %.3f ms" % (
    t(lambda: s.check_states_ptr(P.ctypes.data, n, hv.ctypes.data)), t(lambda: s.check_states_ptr(pp.data_ptr(), n, pv.data_ptr()))))
ok = s.check_states(P)
A = np.ascontiguousarray(P[ok][:65536]); B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
pA, pB = torch.from_numpy(A).pin_memory().numpy(), torch.from_numpy(B).pin_memory().numpy()
print("edges 65536: pageable %.3f ms, pinned %.3f ms" % (t(lambda: s.check_edges(A, B)), t(lambda: s.check_edges(pA, pB))))
