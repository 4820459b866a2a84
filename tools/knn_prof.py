"""Dev: k-NN timing (device side) for ncu launch lists."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
sc = SE3_SCENES["home"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
k = int(sys.argv[3]) if len(sys.argv) > 3 else 32
keys = se3_poses(n, sc["min"], sc["max"], SEED); qs = se3_poses(nq, sc["min"], sc["max"], SEED + 2)
nn = M.Nigh(M.nn.SE3); nn.insert(keys)
nn.knearest(qs[:256], k)
for _ in range(2):
    t = time.perf_counter(); nn.knearest(qs, k); dt = time.perf_counter() - t
print("n=%d nq=%d k=%d: %.2f ms, %.1f G pair distances/s" % (n, nq, k, dt * 1e3, n * nq / dt / 1e9))
t = time.perf_counter(); nn.nearest(qs); dt = time.perf_counter() - t
print("1-NN: %.2f ms, %.1f G pair distances/s" % (dt * 1e3, n * nq / dt / 1e9))
