"""Batch-issuing RRT over the scenario interface: the caller-side change that lets the planner feed the
GPU.  The loop body is the reference's `Planner<Scenario, PRRT>::Thread::addSample`
(reference include/mpl/prrt.hpp:280-299) applied to B samples at once -- which is also what the reference's
OpenMP threads do concurrently against a shared tree (prrt.hpp:135-153):

    sample (goal-biased on a fraction of the batch, prrt.hpp:301-315)  ->  nearest(scale(q))  (prrt.hpp:66-68)
    d == 0 ? skip : isValid(q) and isValid(nearest, q)                 (prrt.hpp:291-295; steering distance is +inf)
    add node, solution as soon as a goal node is added                (prrt.hpp:60-64)

`isValid(from, to)` already begins with `isValid(to)` (se3_rigid_body_scenario.hpp:136), so one batched edge
call covers both checks.  The backend is any object with check_edges / isGoal / randomSample / sampleGoal
(the GPU scenario from mplambda_b200.scenario, or a CPU stand-in in tests and comparisons) and a nearest-
neighbour object with insert / nearest.
"""
import time

import numpy as np


def se3_random_batch(rng, n, lo, hi):
    """n states distributed as SE3RigidBodyScenario::randomSample (se3...:107-113, randomize.hpp:12-38)."""
    a, b, c = rng.uniform(0, 1, n), rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
    q = np.empty((n, 7))
    q[:, 0] = np.sqrt(1 - a) * np.sin(b)
    q[:, 1] = np.sqrt(1 - a) * np.cos(b)
    q[:, 2] = np.sqrt(a) * np.sin(c)
    q[:, 3] = np.sqrt(a) * np.cos(c)
    q[:, 4:7] = rng.uniform(lo, hi, (n, 3))
    return q


class BatchedRRT:
    """scenario: check_states / check_edges / isGoal / sampleGoal and either sample_batch(rng, n) or
    randomSample(rng); nn: insert / nearest.  All host work per iteration is vectorised numpy."""

    def __init__(self, scenario, nn, start, batch=2048, goal_bias=0.01, seed=1, capacity=1 << 16):
        self.scenario, self.nn = scenario, nn
        self.batch, self.goal_bias = int(batch), float(goal_bias)
        self.rng = np.random.default_rng(seed)
        start = np.asarray(start, dtype=np.float64)
        if not scenario.check_states(start[None])[0]:
            raise ValueError("start state is not valid")            # prrt.hpp:109-110
        self.dim = len(start)
        self.nodes = np.empty((capacity, self.dim))
        self.parent = np.empty(capacity, dtype=np.int64)
        self.nodes[0], self.parent[0], self.n = start, -1, 1
        nn.insert(self._scale(start[None]))
        self.solution = -1
        self.stats = dict(iterations=0, samples=0, edges_checked=0, nodes=1, seconds=0.0)

    def _scale(self, q):
        return np.ascontiguousarray(self.scenario.scale(q))

    def _draw(self):
        sc = self.scenario
        if hasattr(sc, "sample_batch"):
            q = sc.sample_batch(self.rng, self.batch)
        else:
            q = np.array([sc.randomSample(self.rng) for _ in range(self.batch)])
        known = np.zeros(self.batch, dtype=bool)
        for i in range(self.rng.binomial(self.batch, self.goal_bias)):   # prrt.hpp:306-312
            g = sc.sampleGoal(self.rng)
            if g is not None:
                q[i], known[i] = g, True
        return q, known

    def step(self):
        q, known = self._draw()
        idx, dist = self.nn.nearest(self._scale(q))
        cand = np.nonzero(dist > 0)[0]                                # prrt.hpp:283-284
        self.stats["samples"] += self.batch
        if len(cand) == 0:
            return
        ok = self.scenario.check_edges(self.nodes[idx[cand]], q[cand])
        self.stats["edges_checked"] += len(cand)
        new = cand[ok]
        if len(new) == 0:
            return
        if self.n + len(new) > len(self.nodes):
            grow = max(len(self.nodes), len(new))
            self.nodes = np.concatenate([self.nodes, np.empty((grow, self.dim))])
            self.parent = np.concatenate([self.parent, np.empty(grow, dtype=np.int64)])
        first = self.n
        self.nodes[first:first + len(new)] = q[new]
        self.parent[first:first + len(new)] = idx[new]
        self.n += len(new)
        self.nn.insert(self._scale(q[new]))
        goal = known[new]
        if not goal.any() and hasattr(self.scenario, "is_goal_batch"):
            goal = self.scenario.is_goal_batch(q[new])
        hit = np.nonzero(goal)[0]
        if len(hit):
            self.solution = first + int(hit[0])

    def solve(self, time_limit=60.0, max_iterations=10 ** 9):
        t0 = time.perf_counter()
        while self.solution < 0 and time.perf_counter() - t0 < time_limit and self.stats["iterations"] < max_iterations:
            self.step()
            self.stats["iterations"] += 1
        self.stats["seconds"] = time.perf_counter() - t0
        self.stats["nodes"] = self.n
        return self.path()

    def path(self):
        if self.solution < 0:
            return None
        out, i = [], self.solution
        while i >= 0:
            out.append(self.nodes[i].copy())
            i = int(self.parent[i])
        return np.asarray(out[::-1])


def path_cost(scenario, path):
    sp = scenario.space()
    return float(sum(sp.distance(a, b) for a, b in zip(path[:-1], path[1:])))


def solve_sharded(make_planner, exchange, time_limit=60.0, sync_every=4):
    """Independent planner instances, one per rank (the reference's lambdas: src/mpl_coordinator.cpp:591-600),
    exchanging their outcome with collectives instead of the coordinator's TCP relays: every `sync_every`
    iterations an all-reduce tells whether any shard has a solution (the Done broadcast,
    mpl_coordinator.cpp:538-551); then the cheapest path wins and is broadcast to all shards
    (gotPath, mpl_coordinator.cpp:573-589).  -> (cost, winning rank, path), identical on every rank.
    make_planner(rank) must return a BatchedRRT (seeded per rank); exchange is a sharding.SolutionExchange."""
    rrt = make_planner(exchange.rank)
    t0 = time.perf_counter()
    while True:
        for _ in range(sync_every):
            if rrt.solution < 0:
                rrt.step()
                rrt.stats["iterations"] += 1
        timed_out = time.perf_counter() - t0 > time_limit
        if exchange.any_done(rrt.solution >= 0 or timed_out):
            break
    rrt.stats["seconds"] = time.perf_counter() - t0
    rrt.stats["nodes"] = rrt.n
    path = rrt.path()
    cost = path_cost(rrt.scenario, path) if path is not None else float("inf")
    return exchange.publish(cost, path) + (rrt,)
