import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mplambda_b200 as M
from mplambda_b200.workload import *
env = np.load("tests/golden/scenes/autolab.npz")["env"]
s = M.FetchScenario(frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"]), env, checkResolution=0.01)
Q = fetch_configs(1 << 16)
for _ in range(3):
    t = time.perf_counter(); v = s.check_states(Q); dt = time.perf_counter() - t
print("fetch 65536 configs %.2f ms valid %.3f" % (dt * 1e3, v.mean()), s.stats())
