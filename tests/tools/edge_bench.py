"""Dev timing of edge batches (config 3 of BASELINE.json): 65,536 seeded edges per scene."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
from oracle import oracle_py as O

ne = int(os.environ.get("NE", 65536))
for name in sys.argv[1:] or ["cubicles", "apartment", "twistycool"]:
    d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
    s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
    P = se3_poses(ne * 8, sc["min"], sc["max"], SEED)
    A = P[s.check_states(P)][:ne]
    B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
    s.check_edges(A[:256], B[:256])
    s.stats(reset=True)
    t = time.perf_counter(); v, chk = s.check_edges(A, B, return_states_checked=True); dt = time.perf_counter() - t
    st = s.stats(); steps = s.edge_steps(A, B)
    print("%s: %d edges, sum steps %.3g, states checked %.3g (%.1f%%), valid %.3f, %.2f ms -> %.2f M edges/s, %.2f G states/s; bv/state %.1f tri/state %.2f" % (
        name, len(A), steps.sum(), chk, 100 * chk / steps.sum(), v.mean(), dt * 1e3, len(A) / dt / 1e6, chk / dt / 1e9,
        st["bv_tests"] / chk, st["tri_tests"] / chk), flush=True)
    if os.environ.get("CPU"):
        orc = O.SE3Oracle(d["env"], d["robot"], check_resolution=0.1)
        m = 2048; c = O.Counters()
        t = time.perf_counter(); w = orc.check_edges(A[:m], B[:m], O.BVH, counters=c); dc = time.perf_counter() - t
        print("   CPU oracle %d edges: %.1f ms -> %.4f M edges/s (%d threads), states visited %.3g; mismatches %d" % (
            m, dc * 1e3, m / dc / 1e6, O.lib().orc_omp_max_threads(), c.checks, int((w != v[:m]).sum())), flush=True)
    if os.environ.get("DEV"):
        import torch
        t = time.perf_counter(); steps = s.edge_steps(A, B); tp = time.perf_counter() - t
        dA, dB, dS = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(steps.astype(np.int32)).cuda()
        out = torch.empty(len(A), dtype=torch.uint8, device="cuda")
        st_ = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            s.check_edges_device(dA.data_ptr(), dB.data_ptr(), dS.data_ptr(), len(A), int(steps.max()), out.data_ptr(), st_)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            s.check_edges_device(dA.data_ptr(), dB.data_ptr(), dS.data_ptr(), len(A), int(steps.max()), out.data_ptr(), st_)
        e1.record(); torch.cuda.synchronize()
        print("   host plan (steps) %.2f ms; device-resident kernel %.3f ms -> %.2f M edges/s; verdicts equal %s" % (
            tp * 1e3, e0.elapsed_time(e1) / 5, len(A) / (e0.elapsed_time(e1) / 5) / 1e3, bool((out.cpu().numpy().astype(bool) == v).all())), flush=True)
