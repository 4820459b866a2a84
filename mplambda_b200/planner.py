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
        hit = np.nonzero(self._goals(q[new], known[new]))[0]
        if len(hit):
            self.solution = first + int(hit[0])

    def _goals(self, q_new, known_new):
        if not known_new.any() and hasattr(self.scenario, "is_goal_batch"):
            return self.scenario.is_goal_batch(q_new)
        return known_new

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


class DeviceRRT:
    """The same loop run on the device (libmplb's mplb_rrt_*: sampler, nearest neighbour, index-named motion check and
    node selection are kernels; the host reads a 24-byte status per iteration and the tree once at the end).
    scenario: an SE3RigidBodyScenario; the tree's nodes are the keys of `nn` (an empty SE3 Nigh)."""

    def __init__(self, scenario, nn, start, batch=2048, goal_bias=0.01, seed=1):
        import ctypes as C
        from . import _capi
        self._C, self._capi = C, _capi
        self.scenario, self.nn, self.batch = scenario, nn, int(batch)
        self._h = C.c_void_p()
        start = np.ascontiguousarray(start, dtype=np.float64).reshape(7)
        goal = np.ascontiguousarray(scenario.goal_, dtype=np.float64)
        rc = _capi.lib().mplb_rrt_create(scenario._h, nn._h, start.ctypes.data_as(_capi.dp), goal.ctypes.data_as(_capi.dp),
                                         scenario.goalRadius_, scenario.min_.ctypes.data_as(_capi.dp),
                                         scenario.max_.ctypes.data_as(_capi.dp), int(seed), self.batch, float(goal_bias),
                                         C.byref(self._h))
        if rc == _capi.ERR_INVALID:
            raise ValueError(_capi.lib().mplb_last_error().decode("utf-8", "replace"))   # prrt.hpp:109-110
        _capi.check(rc)
        self.solution = -1
        self.stats = dict(iterations=0, samples=0, edges_checked=0, nodes=1, seconds=0.0)
        self.nodes = self.parent = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._capi.lib().mplb_rrt_destroy(self._h)
            self._h = self._C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, time_limit=60.0, max_iterations=10 ** 9):
        st = self._capi.RrtStatus()
        self._capi.check(self._capi.lib().mplb_rrt_run(self._h, int(max_iterations), float(time_limit), self._C.byref(st)))
        self.stats.update(st.as_dict())
        self.solution = int(st.solution)
        self.n = int(st.nodes)
        return self.path()

    def tree(self):
        n = self._C.c_size_t()
        self._capi.check(self._capi.lib().mplb_rrt_tree(self._h, None, None, 0, self._C.byref(n)))
        nodes = np.empty((n.value, 7))
        parent = np.empty(n.value, dtype=np.int64)
        self._capi.check(self._capi.lib().mplb_rrt_tree(self._h, nodes.ctypes.data_as(self._C.c_void_p),
                                                        parent.ctypes.data_as(self._C.c_void_p), n.value, self._C.byref(n)))
        self.nodes, self.parent, self.n = nodes, parent, n.value
        return nodes, parent

    def samples(self, iteration):
        """(states [batch, 7], goal-sample flags [batch]) of iteration `iteration`, as the device loop draws them."""
        q = np.empty((self.batch, 7))
        g = np.empty(self.batch, dtype=np.uint8)
        self._capi.check(self._capi.lib().mplb_rrt_samples(self._h, int(iteration), q.ctypes.data_as(self._C.c_void_p),
                                                           g.ctypes.data_as(self._C.c_void_p)))
        return q, g.astype(bool)

    def path(self):
        if self.solution < 0:
            return None
        nodes, parent = self.tree()
        out, i = [], self.solution
        while i >= 0:
            out.append(nodes[i].copy())
            i = int(parent[i])
        return np.asarray(out[::-1])


class ReplayRRT(BatchedRRT):
    """BatchedRRT fed with the sample batches a DeviceRRT draws (iteration by iteration), on any backend: the tree it
    builds must be the device loop's tree (tests)."""

    def __init__(self, scenario, nn, start, device_rrt):
        super().__init__(scenario, nn, start, batch=device_rrt.batch, goal_bias=0.0, seed=0)
        self._src = device_rrt

    def _draw(self):
        return self._src.samples(self.stats["iterations"])

    def _goals(self, q_new, known_new):   # the device loop's rule: isGoal(q) (se3...:115-117); sampled goals satisfy it
        return known_new | np.array([bool(self.scenario.isGoal(x)) for x in q_new], dtype=bool)


class BatchedRRTStar(BatchedRRT):
    """The asymptotically optimal loop of the reference's second planner, `Planner<Scenario, PCForest>`
    (include/mpl/pcforest.hpp:487-567, one C-FOREST shard), applied to B samples at once:

        sample -> nearest -> isValid(nearest, q)                                  (addSample, pcforest.hpp:487-506)
        k nearest tree nodes, k = ceil(e (1 + 1/d) ln(n + 1))                     (pcforest.hpp:80-87,138)
        parent = the neighbour with the cheapest valid connection                 (pcforest.hpp:512-540)
        rewire every neighbour that gets cheaper through the new node             (pcforest.hpp:548-566)

    The reference collects parent candidates walking the neighbours in ascending distance up to the first one that is
    not cheaper than the nearest node, then tries them in increasing cost and stops at the first valid motion; here the
    same candidates of all new nodes are one `check_edges` batch (and all rewiring motions another), and the cheapest
    valid one wins -- the same choice.  (Candidates the reference never gets to test, because a cheaper one was valid,
    are tested here; an invalid one among them is then also skipped in the rewiring step, where the reference would find
    the reverse motion invalid.)  Samples of one batch do not see each other, like the reference's
    concurrent threads.  As in the reference an Edge is immutable (node, parent edge, length, path cost): rewiring a
    node gives it a new edge, its descendants keep pointing at the old one, so every stored cost is the length of a
    path that exists.  nn needs `knearest` besides insert / nearest; distances come from the nearest-neighbour
    structure (the space's metric).  All host work per iteration is vectorised numpy."""

    def __init__(self, scenario, nn, start, batch=1024, goal_bias=0.01, seed=1, capacity=1 << 16, dimension=6):
        super().__init__(scenario, nn, start, batch, goal_bias, seed, capacity)
        self.dimension = dimension
        # edges: node, parent edge (-1: root), path cost; node -> its current edge
        self._edge_node = np.zeros(capacity, dtype=np.int64)
        self._edge_parent = np.full(capacity, -1, dtype=np.int64)
        self._edge_cost = np.zeros(capacity)
        self.n_edges = 1
        self.node_edge = np.zeros(len(self.nodes), dtype=np.int64)
        self.is_goal = np.zeros(len(self.nodes), dtype=bool)
        self.best_edge, self.best_cost = -1, float("inf")
        self.stats["rewired"] = 0

    edge_node = property(lambda self: self._edge_node[:self.n_edges])
    edge_parent = property(lambda self: self._edge_parent[:self.n_edges])
    edge_cost = property(lambda self: self._edge_cost[:self.n_edges])

    def _cost(self, nodes):
        return self._edge_cost[self.node_edge[nodes]]

    def _new_edges(self, nodes, parent_edges, costs):
        m, first = len(nodes), self.n_edges
        if first + m > len(self._edge_cost):
            grow = max(len(self._edge_cost), m)
            self._edge_node = np.concatenate([self._edge_node, np.zeros(grow, dtype=np.int64)])
            self._edge_parent = np.concatenate([self._edge_parent, np.full(grow, -1, dtype=np.int64)])
            self._edge_cost = np.concatenate([self._edge_cost, np.zeros(grow)])
        self._edge_node[first:first + m], self._edge_parent[first:first + m] = nodes, parent_edges
        self._edge_cost[first:first + m] = costs
        self.n_edges += m
        return np.arange(first, first + m)

    def _better_goal(self, nodes):
        """updateSolution (pcforest.hpp:542-543): the cheapest goal node among `nodes` if it beats the incumbent."""
        nodes = nodes[self.is_goal[nodes]]
        if len(nodes):
            c = self._cost(nodes)
            i = int(np.argmin(c))
            if c[i] < self.best_cost:
                self.best_cost, self.solution = float(c[i]), int(nodes[i])
                self.best_edge = int(self.node_edge[nodes[i]])

    @staticmethod
    def _first_per_key(keys, costs):
        """indices of the cheapest entry of every distinct key (ties: the earliest entry)"""
        order = np.lexsort((costs, keys))
        k = keys[order]
        return order[np.concatenate([[True], k[1:] != k[:-1]])]

    def step(self):
        sc = self.scenario
        q, known = self._draw()
        near, dnear = self.nn.nearest(self._scale(q))
        cand = np.nonzero(dnear > 0)[0]
        self.stats["samples"] += self.batch
        if len(cand) == 0:
            return
        ok = sc.check_edges(self.nodes[near[cand]], q[cand])
        self.stats["edges_checked"] += len(cand)
        new = cand[ok]
        if len(new) == 0:
            return
        qn, m = q[new], len(new)
        k = int(np.ceil(np.e * (1.0 + 1.0 / self.dimension) * np.log(self.n + 1)))
        nbr, ndist = self.nn.knearest(self._scale(qn), k)                    # ascending distance, -1 padded
        have = nbr >= 0
        nbr_safe = np.where(have, nbr, 0)
        nbr_cost = np.where(have, self._cost(nbr_safe), np.inf)
        via = nbr_cost + ndist                                               # path cost through each neighbour
        parent_cost = self._cost(near[new]) + dnear[new]
        parent = near[new].copy()
        # ---- parent selection: the neighbours, in ascending distance, up to the first one that would not be cheaper
        #      than the nearest node (pcforest.hpp:519-526 stops there: `if (nbrPathCost >= parentCost) break`)
        rows, cols = np.nonzero(np.logical_and.accumulate(via < parent_cost[:, None], axis=1))
        rejected = np.zeros_like(have)
        if len(rows):
            v = sc.check_edges(self.nodes[nbr[rows, cols]], qn[rows])
            self.stats["edges_checked"] += len(rows)
            rejected[rows[~v], cols[~v]] = True                               # pcforest.hpp:527,555-556
            r, c = rows[v], cols[v]
            if len(r):
                best = self._first_per_key(r, via[r, c])
                parent_cost[r[best]], parent[r[best]] = via[r[best], c[best]], nbr[r[best], c[best]]
        # ---- insert
        if self.n + m > len(self.nodes):
            grow = max(len(self.nodes), m)
            self.nodes = np.concatenate([self.nodes, np.empty((grow, self.dim))])
            self.parent = np.concatenate([self.parent, np.empty(grow, dtype=np.int64)])
            self.node_edge = np.concatenate([self.node_edge, np.zeros(grow, dtype=np.int64)])
            self.is_goal = np.concatenate([self.is_goal, np.zeros(grow, dtype=bool)])
        first = self.n
        ids = np.arange(first, first + m)
        self.nodes[ids], self.parent[ids] = qn, parent
        goal = known[new].copy()
        if hasattr(sc, "is_goal_batch"):
            goal |= sc.is_goal_batch(qn)
        else:
            goal |= np.array([bool(sc.isGoal(x)) for x in qn])
        self.is_goal[ids] = goal
        self.node_edge[ids] = self._new_edges(ids, self.node_edge[parent], parent_cost)
        self.n += m
        self.nn.insert(self._scale(qn))
        self._better_goal(ids)
        # ---- rewiring: neighbours that get cheaper through the new node (not the ones already found unreachable)
        through = parent_cost[:, None] + ndist
        rows, cols = np.nonzero(have & ~rejected & (through < np.where(have, nbr_cost, -np.inf)))
        if len(rows):
            v = sc.check_edges(qn[rows], self.nodes[nbr[rows, cols]])
            self.stats["edges_checked"] += len(rows)
            r, c = rows[v], cols[v]
            if len(r):
                best = self._first_per_key(nbr[r, c], through[r, c])          # several new nodes, one neighbour: the cheapest
                node, cost, via_new = nbr[r[best], c[best]], through[r[best], c[best]], first + r[best]
                self.node_edge[node] = self._new_edges(node, self.node_edge[via_new], cost)   # setEdge, pcforest.hpp:569-581
                self.parent[node] = via_new
                self.stats["rewired"] += len(node)
                self._better_goal(node)

    def solve(self, time_limit=60.0, max_iterations=10 ** 9, stop_at_first=False):
        """Anytime: runs until the limit and keeps the cheapest path found (stop_at_first: return at the first one)."""
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < time_limit and self.stats["iterations"] < max_iterations:
            if stop_at_first and self.solution >= 0:
                break
            self.step()
            self.stats["iterations"] += 1
        self.stats["seconds"] = time.perf_counter() - t0
        self.stats["nodes"] = self.n
        return self.path()

    def path(self):
        if self.best_edge < 0:
            return None
        out, e = [], self.best_edge
        while e >= 0:                                                         # an edge chain is an existing path
            out.append(self.nodes[self._edge_node[e]].copy())
            e = int(self._edge_parent[e])
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
