import sys, time, numpy as np
sys.path.insert(0, ".")
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
name = "twistycool"
d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P = se3_poses(200000, sc["min"], sc["max"], SEED)
keys = P[s.check_states(P)][:60000]
nn = M.Nigh(M.nn.SE3); nn.insert(keys)
q = se3_poses(2048, sc["min"], sc["max"], SEED + 5)
def t(fn, reps=20):
    fn(); a = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - a) / reps * 1e3
idx, dist = nn.nearest(q)
print("nearest 2048 x 60000: %.3f ms" % t(lambda: nn.nearest(q)))
print("check_edges 2048 (nearest -> sample): %.3f ms" % t(lambda: s.check_edges(keys[idx], q)))
print("insert 600 keys: %.3f ms" % t(lambda: nn.insert(keys[:600]), reps=5))
print("check_states 2048: %.3f ms" % t(lambda: s.check_states(q)))
