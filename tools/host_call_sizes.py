"""mplb_check_states / mplb_check_edges through host buffers over batch sizes (development): where the paths switch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
name = sys.argv[1] if len(sys.argv) > 1 else "cubicles"
d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P = se3_poses(1 << 18, sc["min"], sc["max"], SEED)
ok = s.check_states(P)
A = np.ascontiguousarray(P[ok][:1 << 16]); B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
def best(f, reps=7):
    f(); b = 1e9
    for _ in range(reps):
        t = time.perf_counter(); f(); b = min(b, time.perf_counter() - t)
    return b * 1e6
for n in (1, 64, 512, 1024, 2048, 2049, 4096, 8192, 16384, 32768, 65535, 65536, 131072):
    p, pp = P[:n], pin(P[:n])
    line = "states n=%6d: pageable %8.1f us, pinned %8.1f us" % (n, best(lambda: s.check_states(p)), best(lambda: s.check_states(pp)))
    if n <= len(A):
        a, b, pa, pb = A[:n], B[:n], pin(A[:n]), pin(B[:n])
        line += " | edges: pageable %8.1f us, pinned %8.1f us" % (best(lambda: s.check_edges(a, b), 5), best(lambda: s.check_edges(pa, pb), 5))
    print(line, flush=True)
