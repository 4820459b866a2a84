"""Diagnostic: consecutive streamed host batches (MPLB_TIMELINE=1 prints where the time goes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses

name = sys.argv[1] if len(sys.argv) > 1 else "alpha15"
d = np.load("tests/golden/scenes/%s.npz" % name)
sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
n = 1 << 20
P = se3_poses(n, sc["min"], sc["max"], SEED)
hp = torch.from_numpy(P).pin_memory(); hv = torch.empty(n, dtype=torch.uint8).pin_memory()
dbuf = torch.empty(n * 7, dtype=torch.float64, device="cuda")
def copy_alone():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dbuf.copy_(hp.view(-1), non_blocking=True); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
print("copy alone:", ["%.3f" % copy_alone() for _ in range(6)], flush=True)
for i in range(10):
    t = time.perf_counter(); s.check_states_ptr(hp.data_ptr(), n, hv.data_ptr()); dt = time.perf_counter() - t
    print("call %d: %.3f ms" % (i, dt * 1e3), file=sys.stderr, flush=True)
    if i == 5: time.sleep(0.5)
print("copy alone:", ["%.3f" % copy_alone() for _ in range(6)], flush=True)
