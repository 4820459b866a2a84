"""The batch-issuing RRT (mplambda_b200/planner.py): planner logic on the CPU with the
oracle as backend, and on the GPU with the real scenario -- both must build the SAME tree, because every
verdict and every nearest neighbour is identical."""
import numpy as np
import pytest

from mplambda_b200.planner import BatchedRRT
from mplambda_b200.workload import SE3_SCENES
from cpu_backend import CpuNN, CpuScenario


def _check_path(path, sc, orc_scenario):
    assert path is not None
    np.testing.assert_array_equal(path[0], np.array(sc["start"]))
    assert orc_scenario.isGoal(path[-1])
    assert orc_scenario.check_states(path).all()
    assert orc_scenario.check_edges(path[:-1], path[1:]).all()


def test_cpu_backend_solves_cubicles(scenes):
    env, robot = scenes["cubicles"]
    sc = SE3_SCENES["cubicles"]
    scen = CpuScenario(env, robot, sc)
    rrt = BatchedRRT(scen, CpuNN(), sc["start"], batch=512, seed=3)
    path = rrt.solve(time_limit=120)
    _check_path(path, sc, scen)
    assert rrt.stats["nodes"] == rrt.n > 1 and rrt.stats["edges_checked"] > 0


def test_invalid_start_raises(scenes):
    env, robot = scenes["cubicles"]
    sc = SE3_SCENES["cubicles"]
    scen = CpuScenario(env, robot, sc)
    cand = scen.sample_batch(np.random.default_rng(0), 64)
    bad = cand[~scen.check_states(cand)][0]
    with pytest.raises(ValueError):                 # prrt.hpp:109-110
        BatchedRRT(scen, CpuNN(), bad)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cubicles", "home"])
def test_gpu_planner_matches_cpu_planner(name, scenes):
    import mplambda_b200 as M
    env, robot = scenes[name]
    sc = SE3_SCENES[name]
    g = M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], 0.1)
    rg = BatchedRRT(g, M.Nigh(M.nn.SE3), sc["start"], batch=1024, seed=5)
    pg = rg.solve(time_limit=120)
    c = CpuScenario(env, robot, sc)
    rc = BatchedRRT(c, CpuNN(), sc["start"], batch=1024, seed=5)
    pc = rc.solve(time_limit=300)
    _check_path(pg, sc, c)
    assert rg.n == rc.n and rg.stats["iterations"] == rc.stats["iterations"]
    np.testing.assert_array_equal(rg.nodes[:rg.n], rc.nodes[:rc.n])
    np.testing.assert_array_equal(pg, pc)


@pytest.mark.gpu
def test_knn_then_connect_matches_oracle(scenes, oracle):
    """BASELINE config 4 at test size: PCForest's addNodeNear (pcforest.hpp:512-566) batched -- the k nearest tree
    nodes of each new node, then the motion from the node to each of them.  Neighbours and verdicts bit-exact."""
    import math
    import mplambda_b200 as M
    from mplambda_b200.workload import SEED, se3_poses
    sc = SE3_SCENES["home"]
    env, robot = scenes["home"]
    s = M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], 0.1)
    pool = se3_poses(8192, sc["min"], sc["max"], SEED)
    tree = pool[s.check_states(pool)][:1024]
    cand = se3_poses(2048, sc["min"], sc["max"], SEED + 2)
    new = cand[s.check_states(cand)][:40]
    k = int(math.ceil(math.e * (7.0 / 6.0) * math.log(len(tree) + 1)))
    assert k == 22
    nn = M.Nigh(M.nn.SE3)
    nn.insert(tree)
    ki, kd = nn.knearest(new, k)
    wi, wd = oracle.nn_knearest(tree, new, k)
    assert (ki == wi).all() and (kd == wd).all()
    assert (np.diff(kd, axis=1) >= 0).all()                      # ascending: pcforest.hpp:523-524 relies on it
    frm = np.repeat(new, k, axis=0)
    to = tree[ki.reshape(-1)]
    got = s.check_edges(frm, to)
    want = oracle.SE3Oracle(env, robot, check_resolution=0.1).check_edges(frm, to)
    assert (got == want).all()
    assert 0 < got.sum() < len(got)


def _star_invariants(rrt, sc, orc_scenario):
    """Every edge chain is an existing, valid path whose length is the stored cost."""
    sp = orc_scenario.space()
    path = rrt.path()
    _check_path(path, sc, orc_scenario)
    length = sum(sp.distance(a, b) for a, b in zip(path[:-1], path[1:]))
    assert abs(length - rrt.best_cost) <= 1e-9 * max(1.0, length)
    cost = np.asarray(rrt.edge_cost)
    par = np.asarray(rrt.edge_parent)
    node = np.asarray(rrt.edge_node)
    child = np.nonzero(par >= 0)[0]
    assert (cost[child] > cost[par[child]]).all()                 # costs grow along every chain
    # a node's current edge is its cheapest one
    for n in np.random.default_rng(0).integers(0, rrt.n, 200):
        assert cost[rrt.node_edge[n]] == cost[node == n].min()


def test_cpu_rrt_star_improves_the_path(scenes):
    from mplambda_b200.planner import BatchedRRTStar
    env, robot = scenes["cubicles"]
    sc = SE3_SCENES["cubicles"]
    scen = CpuScenario(env, robot, sc)
    rrt = BatchedRRTStar(scen, CpuNN(), sc["start"], batch=256, seed=2)
    rrt.solve(time_limit=120, stop_at_first=True)
    assert rrt.solution >= 0
    first = rrt.best_cost
    rrt.solve(time_limit=120, max_iterations=rrt.stats["iterations"] + 6)
    assert rrt.best_cost <= first
    _star_invariants(rrt, sc, scen)


@pytest.mark.gpu
def test_gpu_rrt_star_matches_cpu_rrt_star(scenes):
    import mplambda_b200 as M
    from mplambda_b200.planner import BatchedRRTStar
    env, robot = scenes["cubicles"]
    sc = SE3_SCENES["cubicles"]
    g = M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], 0.1)
    rg = BatchedRRTStar(g, M.Nigh(M.nn.SE3), sc["start"], batch=512, seed=4)
    rg.solve(time_limit=300, stop_at_first=True)
    rg.solve(time_limit=300, max_iterations=rg.stats["iterations"] + 5)      # a few improving rounds after the first path
    c = CpuScenario(env, robot, sc)
    rc = BatchedRRTStar(c, CpuNN(), sc["start"], batch=512, seed=4)
    rc.solve(time_limit=900, max_iterations=rg.stats["iterations"])
    assert rg.n == rc.n and rg.stats["rewired"] == rc.stats["rewired"] and rg.stats["rewired"] > 0
    np.testing.assert_array_equal(rg.nodes[:rg.n], rc.nodes[:rc.n])
    np.testing.assert_array_equal(np.asarray(rg.edge_cost), np.asarray(rc.edge_cost))
    assert rg.best_cost == rc.best_cost < float("inf")
    _star_invariants(rg, sc, c)
