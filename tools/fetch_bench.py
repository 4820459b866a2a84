"""Fetch state batches through the C ABI (development): python tools/fetch_bench.py  -> ms per 2^18 configurations"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mplambda_b200 as M
from mplambda_b200.workload import *
env = np.load("tests/golden/scenes/autolab.npz")["env"]
s = M.FetchScenario(frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"]), env, checkResolution=0.01)
Q = fetch_configs(1 << 18)
s.check_states(Q)
best = 1e9
for _ in range(5):
    t = time.perf_counter(); v = s.check_states(Q); best = min(best, time.perf_counter() - t)
print("fetch 2^18 configs %.3f ms = %.1f M checks/s, valid %.4f (sum %d)" % (best * 1e3, len(Q) / best / 1e6, v.mean(), v.sum()), s.stats())
A = Q[v][:1 << 14]; B = fetch_configs(len(A), SEED + 1)
s.check_edges(A[:256], B[:256])
ts = []
for _ in range(4):
    t = time.perf_counter(); ev = s.check_edges(A, B); ts.append((time.perf_counter() - t) * 1e3)
print("fetch 16384 edges %s ms valid %.4f (sum %d)" % (" ".join("%.2f" % x for x in ts), ev.mean(), ev.sum()))
