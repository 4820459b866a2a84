"""Streamed vs copy-then-kernel host batches from pinned memory over batch sizes (development; MPLB_STREAM_MIN picks)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
name = sys.argv[1] if len(sys.argv) > 1 else "alpha15"
d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P = torch.from_numpy(se3_poses(1 << 20, sc["min"], sc["max"], SEED)).pin_memory()
V = torch.empty(1 << 20, dtype=torch.uint8).pin_memory()
for n in (65536, 98304, 131072, 196608, 262144, 524288, 1048576):
    for _ in range(3): s.check_states_ptr(P.data_ptr(), n, V.data_ptr())
    b = 1e9
    for _ in range(9):
        t = time.perf_counter(); s.check_states_ptr(P.data_ptr(), n, V.data_ptr()); b = min(b, time.perf_counter() - t)
    print("n=%8d: %8.1f us" % (n, b * 1e6), flush=True)
