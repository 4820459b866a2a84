"""Diagnostic: latency of single-edge launches (which edges are the long poles of a small batch?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses

name = sys.argv[1] if len(sys.argv) > 1 else "cubicles"
ne = 256
d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P = se3_poses(ne * 8, sc["min"], sc["max"], SEED)
A = P[s.check_states(P)][:ne]
B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
steps = s.edge_steps(A, B)
dA, dB, dS = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(steps.astype(np.int32)).cuda()
out = torch.empty(ne, dtype=torch.uint8, device="cuda")
st_ = torch.cuda.current_stream().cuda_stream
rows = []
for i in range(ne):
    def run():
        s.check_edges_device(dA[i:].data_ptr(), dB[i:].data_ptr(), dS[i:].data_ptr(), 1, int(steps[i]), out.data_ptr(), st_)
    run(); torch.cuda.synchronize()
    s.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); run(); e1.record(); torch.cuda.synchronize()
    stt = s.stats()
    rows.append((e0.elapsed_time(e1) / 3 * 1e3, int(steps[i]), int(out[0]), stt["checks"] // 3, stt["bv_tests"] // 3, stt["tri_tests"] // 3, stt["exact_tests"] // 3))
rows = np.array(rows)
order = np.argsort(rows[:, 0])
print("single-edge latency us: median %.0f p90 %.0f max %.0f" % (np.median(rows[:, 0]), np.percentile(rows[:, 0], 90), rows[:, 0].max()))
print("slowest 8: (us, steps, valid, states checked, bv, tri, exact)")
for j in order[-8:]:
    print("  ", ["%.0f" % x for x in rows[j]])
print("valid ones: median us %.0f ; invalid: median %.0f" % (np.median(rows[rows[:, 2] == 1, 0]), np.median(rows[rows[:, 2] == 0, 0])))
