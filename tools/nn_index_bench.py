"""NN search through the index vs brute force (development): python tools/nn_index_bench.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import se3_poses, SE3_SCENES
sc = SE3_SCENES["home"]
for n, nq, k in ((16384, 16384, 32), (65536, 4096, 1), (65536, 4096, 26), (1000000, 1024, 45), (1000000, 65536, 1)):
    keys = se3_poses(n, sc["min"], sc["max"], 5); qs = se3_poses(nq, sc["min"], sc["max"], 6)
    dq = torch.from_numpy(qs).cuda()
    di = torch.empty((nq, k), dtype=torch.int64, device="cuda"); dd = torch.empty((nq, k), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for on in (True, False):
        nn = M.Nigh(M.nn.SE3); nn.set_index(on); nn.insert(keys)
        t0 = time.perf_counter(); nn.knearest_device(dq.data_ptr(), nq, k, di.data_ptr(), dd.data_ptr(), st); torch.cuda.synchronize()
        first = time.perf_counter() - t0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): nn.knearest_device(dq.data_ptr(), nq, k, di.data_ptr(), dd.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        t = time.perf_counter(); ih, dh = nn.knearest(qs, k); host = time.perf_counter() - t
        res[on] = (e0.elapsed_time(e1) / 5, first * 1e3, host * 1e3, di.cpu().numpy().copy(), nn.index_info())
    same = (res[True][3] == res[False][3]).all()
    print("n %7d nq %6d k %2d: index %.3f ms (first call incl. build %.2f ms, host call %.2f ms) brute %.3f ms (host call %.2f ms) same %s  %s" % (
        n, nq, k, res[True][0], res[True][1], res[True][2], res[False][0], res[False][2], same, res[True][4]), flush=True)
