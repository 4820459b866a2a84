"""Anytime planning: BatchedRRTStar (PCForest's loop, mplambda_b200/planner.py) for a fixed wall-clock budget with
the GPU scenario + GPU nearest neighbours vs the CPU oracle on all host cores: tree size and best path cost."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.planner import BatchedRRTStar
from mplambda_b200.workload import SE3_SCENES
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpu_backend import CpuNN, CpuScenario  # noqa: E402

budget = float(os.environ.get("BUDGET", 1.0))
threads = len(os.sched_getaffinity(0))
for name in sys.argv[1:] or ["cubicles", "twistycool", "home"]:
    d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
    row = {"scene": name, "budget_s": budget}
    for backend in ("gpu", "cpu"):
        if backend == "gpu":
            scen, nn = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1), M.Nigh(M.nn.SE3)
        else:
            scen, nn = CpuScenario(d["env"], d["robot"], sc, threads), CpuNN(threads)
        rrt = BatchedRRTStar(scen, nn, sc["start"], batch=1024, seed=1)
        rrt.solve(time_limit=budget)
        row[backend] = {"nodes": rrt.n, "iterations": rrt.stats["iterations"], "edges_checked": rrt.stats["edges_checked"],
                        "rewired": rrt.stats["rewired"], "best_cost": None if rrt.best_cost == float("inf") else round(rrt.best_cost, 1)}
    print(json.dumps(row), flush=True)
