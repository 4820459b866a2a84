"""Dev: Fetch edge batch (BASELINE config 5) timing + parity on a prefix."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.workload import *
from oracle import oracle_py as O
env = np.load("tests/golden/scenes/autolab.npz")["env"]
frame = frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"])
f = M.FetchScenario(frame, env, checkResolution=0.01)
Q = fetch_configs(1 << 16, SEED)
v = f.check_states(Q)
A = Q[v][:16384]; B = fetch_configs(len(A), SEED + 1)
f.check_edges(A[:256], B[:256])
f.stats(reset=True)
for _ in range(2):
    t = time.perf_counter(); ev, chk = f.check_edges(A, B, return_states_checked=True); dt = time.perf_counter() - t
print("fetch edges %d: %.2f ms -> %.2f M edges/s, valid %.3f, states evaluated %d of %d, launches %d" % (
    len(A), dt * 1e3, len(A) / dt / 1e6, ev.mean(), chk, f.edge_steps(A, B).sum(), f.stats()["launches"] // 2))
m = int(os.environ.get("M", 2048))
w = O.FetchOracle(env, frame, 0.01).check_edges(A[:m], B[:m], threads=len(os.sched_getaffinity(0)))
print("mismatches vs oracle on first %d: %d (valid %d)" % (m, (w != ev[:m]).sum(), w.sum()))
