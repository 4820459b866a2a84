"""How a Fetch edge batch spends its time (development): launch groups and states evaluated per call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mplambda_b200 as M
from mplambda_b200.workload import *
env = np.load("tests/golden/scenes/autolab.npz")["env"]
s = M.FetchScenario(frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"]), env, checkResolution=0.01)
Q = fetch_configs(1 << 16); v = s.check_states(Q)
for ne in (256, 2048, 16384):
    A = Q[v][:ne]; B = fetch_configs(len(A), SEED + 1)
    s.check_edges(A, B); s.stats(reset=True)
    best = 1e9
    for _ in range(3):
        t = time.perf_counter(); ev, chk = s.check_edges(A, B, return_states_checked=True); best = min(best, time.perf_counter() - t)
    st = s.stats(reset=True)
    print("%6d edges: %.2f ms, launch groups per call %d, states evaluated %d (of %d interpolated), valid %d" % (
        len(A), best * 1e3, st["launches"] // 18, chk, int(s.edge_steps(A, B).sum()), ev.sum()), flush=True)
