"""Small end-to-end case for compute-sanitizer (memcheck): every kernel family once, tiny sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.workload import *

d = np.load("tests/golden/scenes/twistycool.npz"); sc = SE3_SCENES["twistycool"]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P = se3_poses(3000, sc["min"], sc["max"], SEED)
v = s.check_states(P)
A = P[v][:200]; B = se3_poses(200, sc["min"], sc["max"], SEED + 1)
B[::2] = A[::2] + (B[::2] - A[::2]) * 0.01
for i in range(0, 200, 2):
    B[i, :4] /= np.linalg.norm(B[i, :4])
e = s.check_edges(A, B)
print("se3 states valid %.3f edges valid %.3f" % (v.mean(), e.mean()), s.stats())
d = np.load("tests/golden/scenes/apartment.npz"); sc = SE3_SCENES["apartment"]
s2 = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P2 = se3_poses(600, sc["min"], sc["max"], SEED)
print("apartment (wide entries) valid %.3f" % s2.check_states(P2).mean())
nn = M.Nigh(M.nn.SE3); nn.insert(P[:1500])
print("nn", nn.nearest(P[1500:1600])[0][:5], nn.knearest(P[1500:1540], 9)[0].shape)
env = np.load("tests/golden/scenes/autolab.npz")["env"]
f = M.FetchScenario(frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"]), env, checkResolution=0.01)
Q = fetch_configs(1500)
fv = f.check_states(Q)
QA = Q[fv][:20]; QB = QA + (Q[:20] - QA) * 0.05
print("fetch valid %.3f edges %s" % (fv.mean(), f.check_edges(QA, QB).astype(int).tolist()))
