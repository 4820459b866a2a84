"""GPU: Fetch validity kernels (FK + MPR self collision + mesh-vs-shape) through the C ABI vs the
FP64 oracle.  Bar: bit-exact verdicts (the device runs the same FP64 operation sequence; only the FK's
sin/cos come from CUDA's libm)."""
import os

import numpy as np
import pytest

import mplambda_b200 as M
from mplambda_b200.workload import FETCH_TASKS, SEED, fetch_configs, frame_from_option
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def autolab(scenes):
    return scenes["autolab"][0]


@pytest.mark.parametrize("task", sorted(FETCH_TASKS))
def test_known_answers(task, autolab):
    g = np.load(os.path.join(GOLDEN, "fetch_verdicts.npz"))
    s = M.FetchScenario(frame_from_option(FETCH_TASKS[task]["env_frame"]), autolab, checkResolution=0.01)
    Q = fetch_configs(4096, SEED)
    got = s.check_states(Q)
    want = g[task + "_valid"]
    assert (got == want).all(), np.nonzero(got != want)[0][:10]
    assert s.isValid(FETCH_TASKS[task]["start"])
    for i in range(0, 40, 3):
        assert s.isValid(Q[i]) == bool(want[i])


def test_states_match_oracle(autolab, oracle):
    frame = frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"])
    s = M.FetchScenario(frame, autolab, checkResolution=0.01)
    Q = fetch_configs(200000, SEED + 3)
    got = s.check_states(Q)
    want = oracle.FetchOracle(autolab, frame, 0.01).check_states(Q)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, "%d verdicts differ, first %s" % (len(bad), bad[:10])
    assert 0.3 < want.mean() < 0.7


def test_recorded_paths_and_edges(autolab, oracle):
    frame = frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"])
    s = M.FetchScenario(frame, autolab, checkResolution=0.01)
    orc = oracle.FetchOracle(autolab, frame, 0.01)
    paths = np.load(os.path.join(GOLDEN, "fetch_paths.npz"))
    for name in paths.files:
        P = paths[name]
        assert (s.check_states(P) == orc.check_states(P)).all(), name
        assert (s.edge_steps(P[:-1], P[1:]) == [orc.edge_steps(a, b) for a, b in zip(P[:-1], P[1:])]).all()
        assert (s.check_edges(P[:-1], P[1:]) == orc.check_edges(P[:-1], P[1:])).all(), name
    P = paths["makePath010"]
    assert s.check_edges(P[:-1], P[1:]).all() and all(s.isValid(a, b) for a, b in zip(P[:-1], P[1:]))


def test_random_edges_match_oracle(autolab, oracle):
    frame = frame_from_option(FETCH_TASKS["fetch_task_002"]["env_frame"])
    s = M.FetchScenario(frame, autolab, checkResolution=0.01)
    orc = oracle.FetchOracle(autolab, frame, 0.01)
    Q = fetch_configs(3000, SEED + 9)
    Q = Q[s.check_states(Q)]
    n = 300
    A, B = Q[:n], Q[n:2 * n].copy()
    B[::2] = A[::2] + (B[::2] - A[::2]) * 0.03        # short hops: many fully valid edges
    got, checked = s.check_edges(A, B, return_states_checked=True)
    want = orc.check_edges(A, B)
    assert (got == want).all(), np.nonzero(got != want)[0][:10]
    assert 0.1 < want.mean() < 0.9 and checked > 0
    assert s.isValid(A[0], A[0])                       # steps == 0: only `to`


def test_empty_environment_only_self_collision(oracle):
    s = M.FetchScenario(np.hstack([np.eye(3), np.zeros((3, 1))]), np.zeros((0, 3, 3)), checkResolution=0.01)
    Q = fetch_configs(20000, SEED + 1)
    want = np.array([oracle.FetchOracle.self_collision(q) == 0 for q in Q])
    assert (s.check_states(Q) == want).all()


def test_edge_interval_certificates_at_extremes(autolab, oracle):
    """isValid(from,to) is decided through interval certificates: zero-length and few-step edges, an empty batch,
    a single edge, and a very fine resolution (~60 000 samples per edge) must still give the discrete verdict."""
    frame = frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"])
    f = M.FetchScenario(frame, autolab, checkResolution=0.01)
    orc = oracle.FetchOracle(autolab, frame, 0.01)
    Q = fetch_configs(4096, SEED)
    A = Q[f.check_states(Q)][:64]
    assert f.check_edges(A, A).all()                                         # steps == 0: only `to`
    B = A.copy()
    B[:, 3] += np.arange(64) % 6 * 0.01 * 0.999                               # 0 .. 5 steps
    assert (f.check_edges(A, B) == orc.check_edges(A, B)).all()
    assert f.check_edges(np.zeros((0, 8)), np.zeros((0, 8))).shape == (0,)
    C = fetch_configs(64, SEED + 9)
    assert (f.check_edges(A[:1], C[:1]) == orc.check_edges(A[:1], C[:1])).all()
    fine = M.FetchScenario(frame, autolab, checkResolution=0.0005)
    got, evaluated = fine.check_edges(A[:32], C[:32], return_states_checked=True)
    assert (got == oracle.FetchOracle(autolab, frame, 0.0005).check_edges(A[:32], C[:32])).all()
    assert evaluated < fine.edge_steps(A[:32], C[:32]).sum() // 10             # certified, not enumerated


def test_large_environment_uses_the_global_memory_traversal(scenes, oracle):
    """An environment whose hierarchy does not fit in shared memory (the Apartment mesh, 74 228 nodes, shrunk to
    room size around the robot) goes through the global-memory variant of the environment kernel."""
    env = scenes["apartment"][0] * 0.012
    env = env - env.reshape(-1, 3).mean(axis=0) + np.array([0.9, 0.0, 0.6])
    frame = np.hstack([np.eye(3), np.zeros((3, 1))])
    f = M.FetchScenario(frame, env, checkResolution=0.01)
    orc = oracle.FetchOracle(env, frame, 0.01)
    Q = fetch_configs(6000, SEED + 3)
    got = f.check_states(Q)
    want = orc.check_states(Q)
    assert (got == want).all(), np.nonzero(got != want)[0][:10]
    self_only = np.array([oracle.FetchOracle.self_collision(q) == 0 for q in Q[:2000]])
    assert (want[:2000] != self_only).any()                                   # the environment really decides some
    A = Q[got][:200]
    B = fetch_configs(len(A), SEED + 4)
    assert (f.check_edges(A, B) == orc.check_edges(A, B)).all()


def test_work_lists_overflow():
    """MPLB_FETCH_LIST_CAP=64: the pair and leaf lists between the culling and the exact kernels overflow at once -- pairs
    are then decided inside the culling kernel, (state, geometry) items whose leaves did not fit are done again by the
    exact redo pass: verdicts as the oracle's, states and edges."""
    import subprocess
    import sys
    from conftest import GOLDEN, ROOT
    code = ("import numpy as np, sys, os; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import mplambda_b200 as M\n"
            "from mplambda_b200.workload import FETCH_TASKS, SEED, fetch_configs, frame_from_option\n"
            "from oracle import oracle_py as O\n"
            "env = np.load(os.path.join(%r, 'scenes', 'autolab.npz'))['env']\n"
            "frame = frame_from_option(FETCH_TASKS['fetch_task_001']['env_frame'])\n"
            "f = M.FetchScenario(frame, env, checkResolution=0.01)\n"
            "orc = O.FetchOracle(env, frame, 0.01)\n"
            "Q = fetch_configs(30000, SEED + 5)\n"
            "v = f.check_states(Q); w = orc.check_states(Q)\n"
            "assert (v == w).all(), int((v != w).sum())\n"
            "A = Q[v][:300]; B = fetch_configs(len(A), SEED + 6)\n"
            "assert (f.check_edges(A, B) == orc.check_edges(A, B)).all()\nprint('ok')\n" % (ROOT, os.path.join(ROOT, "tests"), GOLDEN))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MPLB_FETCH_LIST_CAP="64"), capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout[-500:] + out.stderr[-2000:]
