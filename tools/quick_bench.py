"""Scratch timing harness used during development (not the contract bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses

def main():
    names = sys.argv[1:] or ["twistycool", "alpha15", "cubicles", "home", "apartment"]
    for name in names:
        d = np.load("tests/golden/scenes/%s.npz" % name)
        sc = SE3_SCENES[name]
        s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
        n = 1 << 20
        P = se3_poses(n, sc["min"], sc["max"], SEED)
        dP = torch.from_numpy(P).cuda()
        out = torch.empty(n, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            s.check_states_device(dP.data_ptr(), n, out.data_ptr(), st)
        torch.cuda.synchronize()
        s.stats(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5
        for _ in range(reps):
            s.check_states_device(dP.data_ptr(), n, out.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        stt = s.stats()
        print("%s: %.3f ms / 2^20 poses = %.1f M checks/s; valid %.3f; per check bv %.1f tri %.2f exact %.4f" % (
            name, ms, n / ms / 1e3, out.float().mean().item(), stt["bv_tests"] / reps / n, stt["tri_tests"] / reps / n,
            stt["exact_tests"] / reps / n), flush=True)
        hp = torch.from_numpy(P).pin_memory(); hv = torch.empty(n, dtype=torch.uint8).pin_memory()
        s.check_states_ptr(hp.data_ptr(), n, hv.data_ptr())
        t = time.perf_counter()
        for _ in range(3):
            s.check_states_ptr(hp.data_ptr(), n, hv.data_ptr())
        dt = (time.perf_counter() - t) / 3
        print("   e2e pinned host: %.3f ms = %.1f M checks/s" % (dt * 1e3, n / dt / 1e6), flush=True)
        # edges
        ne = 4096
        A = P[hv.numpy().astype(bool)][:ne]
        B = se3_poses(ne, sc["min"], sc["max"], SEED + 1)
        t = time.perf_counter(); v, chk = s.check_edges(A, B, return_states_checked=True); dt = time.perf_counter() - t
        t = time.perf_counter(); v, chk = s.check_edges(A, B, return_states_checked=True); dt = time.perf_counter() - t
        steps = s.edge_steps(A, B)
        print("   edges: %d edges, sum steps %d, states checked %d, valid %.3f, %.3f ms -> %.2f M edges/s" % (
            ne, steps.sum(), chk, v.mean(), dt * 1e3, ne / dt / 1e6), flush=True)

main()
