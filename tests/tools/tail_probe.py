"""How much of a pose-batch launch is tail?  time(n) for n = 2^17..2^22 device-resident poses, and the
per-pose work distribution (BV tests) from the oracle on a sample.  Development tool."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses

def main():
    names = sys.argv[1:] or ["alpha15", "twistycool"]
    for name in names:
        d = np.load("tests/golden/scenes/%s.npz" % name)
        sc = SE3_SCENES[name]
        s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
        N = 1 << 22
        P = se3_poses(N, sc["min"], sc["max"], SEED)
        dP = torch.from_numpy(P).cuda()
        out = torch.empty(N, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for lg in range(14, 23):
            n = 1 << lg
            for _ in range(3):
                s.check_states_device(dP.data_ptr(), n, out.data_ptr(), st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 10
            for _ in range(reps):
                s.check_states_device(dP.data_ptr(), n, out.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print("%s n=2^%d: %.3f ms  %.1f M/s" % (name, lg, ms, n / ms / 1e3), flush=True)
        if os.environ.get("HIST"):
            sys.path.insert(0, "oracle")
            import oracle_py as O
            orc = O.SE3Oracle(d["env"], d["robot"])
            bv = []
            for i in range(0, 1 << 16):
                c = O.Counters(); orc.check_states(P[i:i + 1], O.BVH, 1, c); bv.append(c.bv_tests)
            bv = np.array(bv)
            print(name, "per-pose BV tests: mean %.1f p50 %d p99 %d p99.9 %d max %d; share of total in top 0.1%%: %.3f" % (
                bv.mean(), np.median(bv), np.percentile(bv, 99), np.percentile(bv, 99.9), bv.max(),
                np.sort(bv)[-len(bv) // 1000:].sum() / bv.sum()))

main()
