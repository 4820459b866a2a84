"""Dev experiment: BV/tri test counts per check for valid vs colliding poses, GPU vs oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
from oracle import oracle_py as O
for name in sys.argv[1:] or ["twistycool", "alpha15"]:
    d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
    s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
    orc = O.SE3Oracle(d["env"], d["robot"])
    P = se3_poses(1 << 18, sc["min"], sc["max"], SEED)
    v = s.check_states(P)
    for label, sub in (("valid", P[v]), ("colliding", P[~v])):
        s.stats(reset=True); s.check_states(sub); st = s.stats()
        c = O.Counters(); orc.check_states(sub, O.BVH, counters=c)
        n = len(sub)
        print("%s %s n=%d  GPU bv %.1f tri %.2f exact %.3f | CPU bv %.1f tri %.2f" % (
            name, label, n, st["bv_tests"] / n, st["tri_tests"] / n, st["exact_tests"] / n, c.bv_tests / n, c.tri_tests / n))
