"""Diagnostic: host-side breakdown of mplb_check_edges (MPLB_TIMELINE=1), pageable vs pinned inputs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mplambda_b200 as M
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses
name = sys.argv[1] if len(sys.argv) > 1 else "cubicles"
for ne in (4096, 65536):
    d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
    s = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1)
    P = se3_poses(ne * 8, sc["min"], sc["max"], SEED)
    A = np.ascontiguousarray(P[s.check_states(P)][:ne]); B = se3_poses(len(A), sc["min"], sc["max"], SEED + 1)
    for rep in range(3):
        t = time.perf_counter(); v = s.check_edges(A, B); dt = time.perf_counter() - t
        print("pageable numpy, %d edges: %.3f ms" % (len(A), dt * 1e3), file=sys.stderr, flush=True)
    pA, pB = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
    for rep in range(3):
        t = time.perf_counter(); v = s.check_edges(pA.numpy(), pB.numpy()); dt = time.perf_counter() - t
        print("pinned, %d edges: %.3f ms" % (len(A), dt * 1e3), file=sys.stderr, flush=True)
