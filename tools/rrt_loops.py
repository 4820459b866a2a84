"""Host-driven BatchedRRT against the device-resident loop (mplb_rrt_*): iterations, nodes, time per iteration (development)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.planner import BatchedRRT, DeviceRRT
from mplambda_b200.workload import SE3_SCENES
for name in sys.argv[1:] or ["cubicles", "twistycool"]:
    d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
    g = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
    for batch in (512, 2048, 8192):
        for seed in (1, 2):
            a = BatchedRRT(g, M.Nigh(M.nn.SE3), sc["start"], batch=batch, seed=seed); pa = a.solve(time_limit=30)
            b = DeviceRRT(g, M.Nigh(M.nn.SE3), sc["start"], batch=batch, seed=seed); pb = b.solve(time_limit=30)
            sa, sb = a.stats, b.stats
            print("%-10s batch %5d seed %d | host loop: %.4f s, %3d iterations (%.0f us each), %5d nodes | device loop: %.4f s, %3d iterations (%.0f us each), %5d nodes" % (
                name, batch, seed, sa["seconds"], sa["iterations"], 1e6 * sa["seconds"] / max(1, sa["iterations"]), a.n,
                sb["seconds"], sb["iterations"], 1e6 * sb["seconds"] / max(1, sb["iterations"]), sb["nodes"]), flush=True)
            b.close()
