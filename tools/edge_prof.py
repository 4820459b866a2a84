"""Dev: a few device-resident edge-batch launches for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
name = sys.argv[1] if len(sys.argv) > 1 else "cubicles"
ne = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P = se3_poses(ne * 8, sc["min"], sc["max"], SEED)
A = P[s.check_states(P)][:ne]; B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
steps = s.edge_steps(A, B)
dA, dB, dS = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(steps.astype(np.int32)).cuda()
out = torch.empty(len(A), dtype=torch.uint8, device="cuda")
for _ in range(3):
    s.check_edges_device(dA.data_ptr(), dB.data_ptr(), dS.data_ptr(), len(A), int(steps.max()), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("valid", out.float().mean().item())
