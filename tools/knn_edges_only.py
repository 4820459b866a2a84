"""One index-named k-NN edge batch on Home (development; for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
cs = torch.cuda.current_stream()
name = sys.argv[1] if len(sys.argv) > 1 else "home"
d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
P = se3_poses(8 << 16, sc["min"], sc["max"], SEED); ok = s.check_states(P)
tree = P[ok][:3000]; new = P[ok][70000:74096]; k = 26
nn = M.Nigh(M.nn.SE3); nn.insert(tree)
d_new = torch.from_numpy(np.ascontiguousarray(new)).cuda()
d_i = torch.empty((len(new), k), dtype=torch.int64, device="cuda"); d_d = torch.empty((len(new), k), dtype=torch.float64, device="cuda")
d_v = torch.empty(len(new) * k, dtype=torch.uint8, device="cuda")
nn.knearest_device(d_new.data_ptr(), len(new), k, d_i.data_ptr(), d_d.data_ptr(), cs.cuda_stream)
for _ in range(3):
    s.check_edges_indexed_device(d_new.data_ptr(), len(new), None, nn.keys_device(), nn.size(), d_i.data_ptr(), len(new) * k, d_v.data_ptr(), from_div=k, stream=cs.cuda_stream)
torch.cuda.synchronize()
print(int(d_v.sum()), s.stats())
