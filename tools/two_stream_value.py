"""Device-resident pose batches issued on one stream vs alternately on two (development): does the drain of one launch
overlap the start of the next?  Two scenario objects stand in for per-stream control blocks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
name = sys.argv[1] if len(sys.argv) > 1 else "alpha15"
d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
S = [M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1) for _ in range(2)]
n = 1 << 20
P = [torch.from_numpy(se3_poses(n, sc["min"], sc["max"], SEED + i)).cuda() for i in range(4)]
out = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
st = [torch.cuda.Stream(), torch.cuda.Stream()]
def run(two, K=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st[0]); st[1].wait_event(e0)
    for j in range(K):
        q = j % 2 if two else 0
        S[q].check_states_device(P[j % 4].data_ptr(), n, out[j % 4].data_ptr(), st[q].cuda_stream)
    ev = torch.cuda.Event(); ev.record(st[1]); st[0].wait_event(ev)
    e1.record(st[0]); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
for two in (0, 1, 0, 1):
    run(two, 4)
    print(name, "two streams" if two else "one stream ", "%.4f ms per 2^20 poses" % run(two), flush=True)
ref = S[0].check_states(P[0].cpu().numpy())
print("verdicts equal:", bool((out[0].cpu().numpy().astype(bool) == ref).all()))
