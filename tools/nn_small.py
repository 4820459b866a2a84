"""k-NN on small trees: brute-force pipeline vs the grid index (MPLB_NN_GRID_MIN picks) (development)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import se3_poses, SE3_SCENES
sc = SE3_SCENES["home"]
cs = torch.cuda.current_stream()
for n in (256, 1024, 2807, 8192, 16000):
    for nq, k in ((2048, 1), (4096, 26)):
        keys = se3_poses(n, sc["min"], sc["max"], 5); qs = se3_poses(nq, sc["min"], sc["max"], 6)
        dq = torch.from_numpy(qs).cuda()
        di = torch.empty((nq, k), dtype=torch.int64, device="cuda"); dd = torch.empty((nq, k), dtype=torch.float64, device="cuda")
        nn = M.Nigh(M.nn.SE3); nn.insert(keys)
        for _ in range(3): nn.knearest_device(dq.data_ptr(), nq, k, di.data_ptr(), dd.data_ptr(), cs.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cs)
        for _ in range(10): nn.knearest_device(dq.data_ptr(), nq, k, di.data_ptr(), dd.data_ptr(), cs.cuda_stream)
        e1.record(cs); torch.cuda.synchronize()
        print("n %6d nq %5d k %2d: %.3f ms  indexed %d" % (n, nq, k, e0.elapsed_time(e1) / 10, nn.index_info()["indexed"]), flush=True)
