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

# ---- Fetch: 2^18 configurations per seed and 16,384 edges, both environment frames
from mplambda_b200.workload import FETCH_TASKS, fetch_configs, frame_from_option
env = np.load("tests/golden/scenes/autolab.npz")["env"]
ftot = dict(states=0, state_mis=0, edges=0, edge_mis=0)
for task in sorted(FETCH_TASKS):
    frame = frame_from_option(FETCH_TASKS[task]["env_frame"])
    f = M.FetchScenario(frame, env, checkResolution=0.01)
    forc = O.FetchOracle(env, frame, 0.01)
    for k in seeds:
        Q = fetch_configs(1 << 18, SEED + 100 * k)
        v = f.check_states(Q); w = forc.check_states(Q, threads=thr)
        A = Q[v][:16384]; B = fetch_configs(len(A), SEED + 100 * k + 1)
        ev = f.check_edges(A, B); ew = forc.check_edges(A, B, threads=thr)
        ftot["states"] += len(Q); ftot["state_mis"] += int((v != w).sum()); ftot["edges"] += len(A); ftot["edge_mis"] += int((ev != ew).sum())
        print("%s seed %d: states %d mismatches %d | edges %d mismatches %d" % (task, k, len(Q), int((v != w).sum()), len(A), int((ev != ew).sum())), flush=True)
print("fetch", ftot)
