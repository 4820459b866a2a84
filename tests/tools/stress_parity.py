"""Wide parity sweep (GPU box): every scene, 2^20 poses and 65,536 edges per seed, all compared with the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
from oracle import oracle_py as O

thr = len(os.sched_getaffinity(0))
seeds = [int(x) for x in (os.environ.get("SEEDS", "0,1").split(","))]
tot = dict(poses=0, pose_mis=0, edges=0, edge_mis=0)
for name in sys.argv[1:] or ["twistycool", "alpha15", "cubicles", "home", "apartment"]:
    d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
    s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
    orc = O.SE3Oracle(d["env"], d["robot"], check_resolution=0.1)
    for k in seeds:
        n = 1 << 20
        P = se3_poses(n, sc["min"], sc["max"], SEED + 100 * k)
        t = time.perf_counter(); v = s.check_states(P); tg = time.perf_counter() - t
        t = time.perf_counter(); w = orc.check_states(P, O.BVH, threads=thr); tc = time.perf_counter() - t
        pm = int((v != w).sum())
        A = np.ascontiguousarray(P[v][:65536]); B = se3_poses(len(A), sc["min"], sc["max"], SEED + 100 * k + 1)
        t = time.perf_counter(); ev = s.check_edges(A, B); teg = time.perf_counter() - t
        t = time.perf_counter(); ew = orc.check_edges(A, B, O.BVH, threads=thr); tec = time.perf_counter() - t
        em = int((ev != ew).sum())
        tot["poses"] += n; tot["pose_mis"] += pm; tot["edges"] += len(A); tot["edge_mis"] += em
        print("%-10s seed %d: poses %d mismatches %d (gpu %.1f ms, cpu %.0f ms) | edges %d mismatches %d (gpu %.1f ms, cpu %.0f ms)" % (
            name, k, n, pm, tg * 1e3, tc * 1e3, len(A), em, teg * 1e3, tec * 1e3), flush=True)
print(tot)
