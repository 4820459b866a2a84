"""Diagnostic (needs the -DMPLB_WARP_TIMES build): the slowest single Cubicles edge, and a 256-edge batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
name = "cubicles"; ne = 256
d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P = se3_poses(ne * 8, sc["min"], sc["max"], SEED)
A = P[s.check_states(P)][:ne]
B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
steps = s.edge_steps(A, B)
i = int(np.nonzero(steps == 7905)[0][0])
for rep in range(2):
    print("single edge %d" % i, file=sys.stderr, flush=True)
    s.check_edges(A[i:i + 1], B[i:i + 1])
for rep in range(2):
    print("256 edges", file=sys.stderr, flush=True)
    s.check_edges(A, B)
