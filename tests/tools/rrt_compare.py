"""RRT time-to-solution: the same batch-issuing planner (mplambda_b200/planner.py) with the GPU scenario +
GPU nearest neighbour vs the CPU oracle on all host cores as the validity/NN backend.  The secondary half
of BASELINE.json's metric ("RRT time-to-solution vs FCL CPU")."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.planner import BatchedRRT
from mplambda_b200.workload import SE3_SCENES
from oracle import oracle_py as O


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpu_backend import CpuNN, CpuScenario  # noqa: E402


def main():
    names = sys.argv[1:] or ["cubicles", "home", "alpha15", "twistycool", "apartment"]
    limit = float(os.environ.get("LIMIT", 30))
    threads = len(os.sched_getaffinity(0))
    for name in names:
        d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
        row = {"scene": name}
        for backend in ("gpu", "cpu"):
            times, ok = [], 0
            for seed in (1, 2, 3):
                if backend == "gpu":
                    scen = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
                    nn = M.Nigh(M.nn.SE3)
                else:
                    scen, nn = CpuScenario(d["env"], d["robot"], sc, threads), CpuNN(threads)
                rrt = BatchedRRT(scen, nn, sc["start"], batch=int(os.environ.get("BATCH", 2048)), seed=seed)
                path = rrt.solve(time_limit=limit)
                if path is not None:
                    ok += 1
                    times.append(rrt.stats["seconds"])
                last = rrt.stats
            row[backend] = {"solved": ok, "of": 3, "median_s": float(np.median(times)) if times else None, "last": last}
        print(json.dumps(row), flush=True)


main()
