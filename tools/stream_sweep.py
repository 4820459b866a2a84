"""e2e time of mplb_check_states (2^20 pinned poses) for the streaming knobs given in the environment."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses

name = sys.argv[1] if len(sys.argv) > 1 else "alpha15"
d = np.load("tests/golden/scenes/%s.npz" % name)
sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
n = 1 << 20
hp = [torch.from_numpy(se3_poses(n, sc["min"], sc["max"], SEED + i)).pin_memory() for i in range(3)]
hv = torch.empty(n, dtype=torch.uint8).pin_memory()
for i in range(4):
    s.check_states_ptr(hp[i % 3].data_ptr(), n, hv.data_ptr())
ts = []
for i in range(12):
    t = time.perf_counter()
    s.check_states_ptr(hp[i % 3].data_ptr(), n, hv.data_ptr())
    ts.append(time.perf_counter() - t)
ts.sort()
print("%s chunk=%s lanes=%s: median %.3f ms, best %.3f ms, stalls %d" % (
    name, os.environ.get("MPLB_STREAM_CHUNK", "default"), os.environ.get("MPLB_STREAM_LANES", "default"),
    ts[len(ts) // 2] * 1e3, ts[0] * 1e3, s.stats()["stream_stalls"]), flush=True)
