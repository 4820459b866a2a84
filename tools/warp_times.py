"""Diagnostic (needs the -DMPLB_WARP_TIMES build, MPLB_LIB=...): device-resident pose batches of several sizes,
the library prints when the warps ran dry / finished."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses

name = sys.argv[1] if len(sys.argv) > 1 else "alpha15"
d = np.load("tests/golden/scenes/%s.npz" % name)
sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P = se3_poses(1 << 20, sc["min"], sc["max"], SEED)
hp = torch.from_numpy(P).pin_memory(); hv = torch.empty(1 << 20, dtype=torch.uint8).pin_memory()
for n in (1024, 1 << 14, 60000, 1 << 20):
    for rep in range(3):
        print("n=%d" % n, file=sys.stderr, flush=True)
        s.check_states_ptr(hp.data_ptr(), n, hv.data_ptr())
