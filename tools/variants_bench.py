"""Device-resident pose-batch time of every libmplb variant under mplambda_b200/_variants (development only):
python tools/variants_bench.py [scene ...]   -- each variant runs in its own process (MPLB_LIB)."""
import glob, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scenes = sys.argv[1:] or ["alpha15", "twistycool"]
child = r'''
import sys, os
sys.path.insert(0, %r)
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
for name in sys.argv[1:]:
    d = np.load("tests/golden/scenes/%%s.npz" %% name); sc = SE3_SCENES[name]
    s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
    n = 1 << 20
    dP = torch.from_numpy(se3_poses(n, sc["min"], sc["max"], SEED)).cuda()
    out = torch.empty(n, dtype=torch.uint8, device="cuda"); st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): s.check_states_device(dP.data_ptr(), n, out.data_ptr(), st)
    torch.cuda.synchronize(); s.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): s.check_states_device(dP.data_ptr(), n, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10; t = s.stats()
    print("  %%-10s %%.3f ms  bv %%.1f tri %%.2f exact %%.3f  nvalid %%d" %% (name, ms, t["bv_tests"] / 10 / n, t["tri_tests"] / 10 / n, t["exact_tests"] / 10 / n, int(out.sum())), flush=True)
    if os.environ.get("VARIANT_EDGES"):
        ne = 1 << 16
        P = se3_poses(ne * 8, sc["min"], sc["max"], SEED)
        A = P[s.check_states(P)][:ne]; B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
        steps = s.edge_steps(A, B)
        dA, dB, dS = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(steps.astype(np.int32)).cuda()
        ov = torch.empty(len(A), dtype=torch.uint8, device="cuda")
        for _ in range(3): s.check_edges_device(dA.data_ptr(), dB.data_ptr(), dS.data_ptr(), len(A), int(steps.max()), ov.data_ptr(), st)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10): s.check_edges_device(dA.data_ptr(), dB.data_ptr(), dS.data_ptr(), len(A), int(steps.max()), ov.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        print("  %%-10s edges: %%d edges %%.3f ms, valid %%d" %% (name, len(A), e0.elapsed_time(e1) / 10, int(ov.sum())), flush=True)
''' % root
for lib in sorted(glob.glob(os.path.join(root, "mplambda_b200/_variants/libmplb_*.so"))):
    print(os.path.basename(lib), flush=True)
    env = dict(os.environ, MPLB_LIB=lib)
    r = subprocess.run([sys.executable, "-c", child] + scenes, env=env, cwd=root, capture_output=True, text=True)
    print(r.stdout + (r.stderr[-600:] if r.returncode else ""), flush=True)
