"""Config 4 (Home: k-NN of new nodes + the motions to their neighbours) split into its device parts (development)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
d = np.load("tests/golden/scenes/home.npz"); sc = SE3_SCENES["home"]
s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
keys = se3_poses(1 << 14, sc["min"], sc["max"], SEED); qs = se3_poses(1 << 16, sc["min"], sc["max"], SEED + 2)
tree = keys[s.check_states(keys)]; new = qs[s.check_states(qs)][:4096]
k = int(math.ceil(math.e * (7.0 / 6.0) * math.log(len(tree) + 1)))
nn = M.Nigh(M.nn.SE3); nn.insert(tree)
d_new = torch.from_numpy(new).cuda()
d_i = torch.empty((len(new), k), dtype=torch.int64, device="cuda"); d_d = torch.empty((len(new), k), dtype=torch.float64, device="cuda")
d_v = torch.empty(len(new) * k, dtype=torch.uint8, device="cuda")
cs = torch.cuda.current_stream()
def knn(): nn.knearest_device(d_new.data_ptr(), len(new), k, d_i.data_ptr(), d_d.data_ptr(), cs.cuda_stream)
def edges(): s.check_edges_indexed_device(d_new.data_ptr(), len(new), None, nn.keys_device(), nn.size(), d_i.data_ptr(), len(new) * k, d_v.data_ptr(), from_div=k, stream=cs.cuda_stream)
def t(f, reps=10):
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cs)
    for _ in range(reps): f()
    e1.record(cs); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("tree %d nodes, %d new nodes, k = %d, %d edges" % (len(tree), len(new), k, len(new) * k))
print("k-NN            %.3f ms" % t(knn))
s.stats(reset=True)
print("indexed edges   %.3f ms" % t(edges))
st = s.stats()
steps = s.edge_steps(np.repeat(new, k, axis=0), tree[d_i.cpu().numpy().reshape(-1)])
print("edge steps: mean %.1f, max %d; states evaluated per call %.0f; valid %.3f" % (steps.mean(), steps.max(), st["checks"] / 12, d_v.float().mean().item()))
