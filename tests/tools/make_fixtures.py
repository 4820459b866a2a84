#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the GPU box only ever reads the generated files).

  tests/golden/scenes/<scene>.npz   triangle soups of the reference's benchmark meshes
        env, robot : float64 [n,3,3]   what load_mesh.hpp:232-293 hands to FCL (robot centred)
        center     : float64 [3]       the centroid subtracted from the robot
     produced by the numpy restatement oracle/collada_np.py; tests/test_collada.py checks the
     product's C++ loader against the same files.
  tests/golden/se3_verdicts.npz     known-answer verdicts from oracle O1 (FP64 brute-force all-pairs
        17-axis SAT, no hierarchy) for the first poses of each scene's seeded stream, plus the
        edge verdicts / step counts of short seeded edge sets.  PARITY UNPINNED: these pin the
        oracle against itself over time (and O2 and the CUDA path against O1); the reference
        holds no golden vector for this path.
  tests/golden/paths.npz            the reference's own weak fixtures: scripts/Twistycool.{2,3}.path
        (x y z qx qy qz qw per line) converted to the state layout.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.collada_np import load_collada  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses  # noqa: E402

REF = "/root/reference"
RES = os.path.join(REF, "resources")
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(os.path.join(OUT, "scenes"), exist_ok=True)
    scenes = {}
    for name, sc in SE3_SCENES.items():
        env = load_collada(os.path.join(RES, sc["env"]))
        robot, center = load_collada(os.path.join(RES, sc["robot"]), True, return_center=True)
        np.savez_compressed(os.path.join(OUT, "scenes", name + ".npz"), env=env, robot=robot, center=center)
        scenes[name] = (env, robot)
        print(name, env.shape, robot.shape, center)
    autolab = load_collada(os.path.join(RES, "AUTOLAB.dae"), False, True)
    np.savez_compressed(os.path.join(OUT, "scenes", "autolab.npz"), env=autolab)

    gold = {}
    for name, npose, nedge in [("twistycool", 4096, 64), ("cubicles", 2048, 32), ("home", 1024, 32),
                               ("alpha15", 96, 0)]:
        env, robot = scenes[name]
        sc = SE3_SCENES[name]
        orc = O.SE3Oracle(env, robot)
        P = se3_poses(npose, sc["min"], sc["max"], SEED)
        gold[name + "_valid"] = orc.check_states(P, O.BRUTE)
        if nedge:
            A = se3_poses(4 * nedge, sc["min"], sc["max"], SEED + 7)
            A = A[orc.check_states(A, O.BRUTE)][:nedge]
            B = se3_poses(len(A), sc["min"], sc["max"], SEED + 8)
            # half of the edges are short hops (more of them survive)
            for i in range(0, len(A), 2):
                B[i] = orc.interpolate(A[i], B[i], 0.004)
            gold[name + "_edge_from"] = A
            gold[name + "_edge_to"] = B
            gold[name + "_edge_steps"] = np.array([orc.edge_steps(a, b) for a, b in zip(A, B)], dtype=np.uint32)
            gold[name + "_edge_valid"] = orc.check_edges(A, B, O.BRUTE)
        print(name, "valid frac", gold[name + "_valid"].mean(),
              "edges valid", gold.get(name + "_edge_valid", np.zeros(1)).mean())
    np.savez_compressed(os.path.join(OUT, "se3_verdicts.npz"), **gold)

    paths = {}
    for f in ("Twistycool.2.path", "Twistycool.3.path"):
        rows = np.loadtxt(os.path.join(REF, "scripts", f))
        paths[f.replace(".", "_")] = np.concatenate([rows[:, 3:7], rows[:, 0:3]], axis=1)
    np.savez_compressed(os.path.join(OUT, "paths.npz"), **paths)

    # src/mpl_fetch.cpp:36-220: ten joint-space solution paths the reference's authors recorded
    import re
    src = open(os.path.join(REF, "src", "mpl_fetch.cpp")).read()
    fp = {}
    for m in re.finditer(r"std::vector<T> (makePath\d+)\(\) \{(.*?)return P;", src, re.S):
        rows = [[float(x) for x in r.split(",")] for r in re.findall(r"addPath\(P, ([^)]*)\)", m.group(2))]
        P = np.array(rows)
        if "std::reverse" in m.group(2):
            P = P[::-1].copy()
        fp[m.group(1)] = P
    np.savez_compressed(os.path.join(OUT, "fetch_paths.npz"), **fp)

    # Fetch known answers from the oracle for the seeded configuration stream (both task frames)
    from mplambda_b200.workload import FETCH_TASKS, fetch_configs, frame_from_option
    fg = {}
    Q = fetch_configs(4096, SEED)
    for task, cfg in FETCH_TASKS.items():
        orc = O.FetchOracle(autolab, frame_from_option(cfg["env_frame"]), 0.01)
        fg[task + "_valid"] = orc.check_states(Q)
        fg[task + "_why"] = np.array([orc.why(q) for q in Q[:1024]], dtype=np.int32)
        print(task, "valid frac", fg[task + "_valid"].mean())
    fk = np.array([O.FetchOracle.fk(q)[0] for q in Q[:64]])
    fg["fk_frames"] = fk
    np.savez_compressed(os.path.join(OUT, "fetch_verdicts.npz"), **fg)


if __name__ == "__main__":
    main()
