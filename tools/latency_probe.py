"""Latency of small launches: one heavy pose (1891 / 549 FCL-order box tests), replicated 1..4096 times,
and light random batches.  Development tool for the engine's low-occupancy path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses

def timeit(s, dP, n, out, st, reps=20):
    for _ in range(3):
        s.check_states_device(dP.data_ptr(), n, out.data_ptr(), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        s.check_states_device(dP.data_ptr(), n, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

def main():
    name = "alpha15"
    d = np.load("tests/golden/scenes/%s.npz" % name)
    sc = SE3_SCENES[name]
    s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
    P = se3_poses(1 << 22, sc["min"], sc["max"], SEED)[:1 << 16]
    out = torch.empty(1 << 16, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for label, idx in (("heavy(1891)", 10388), ("heavy(1727,4tri)", 252), ("p99(549)", 29071), ("light", 0)):
        for n in (1, 2, 8, 32, 64, 512, 4096):
            Q = np.repeat(P[idx:idx + 1], n, 0)
            dQ = torch.from_numpy(np.ascontiguousarray(Q)).cuda()
            s.stats(reset=True)
            us = timeit(s, dQ, n, out, st)
            stt = s.stats()
            print("%-18s n=%5d: %8.1f us   (bv/pose %.0f, valid %d)" % (label, n, us, stt["bv_tests"] / 23 / n, int(out[0])), flush=True)
    for n in (1, 32, 1024, 8192):
        dQ = torch.from_numpy(np.ascontiguousarray(P[:n])).cuda()
        print("random            n=%5d: %8.1f us" % (n, timeit(s, dQ, n, out, st)), flush=True)

main()
