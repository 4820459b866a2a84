"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path -- query sharding, verdict
all-gather, best-solution exchange and the done flag (the collectives bench.py and sharded planners
use with NCCL on GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from mplambda_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 4, 8):
            b = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    rng = np.random.default_rng(0)
    steps = rng.integers(0, 9000, size=5000)
    steps[:50] = 0
    for world in (1, 2, 4, 8):
        b = sharding.balanced_edge_bounds(steps, world)
        assert b[0][0] == 0 and b[-1][1] == len(steps)
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        loads = [int((steps[lo:hi] + 1).sum()) for lo, hi in b]
        assert max(loads) - min(loads) <= 2 * 9001
    assert sharding.balanced_edge_bounds([], 4) == [(0, 0)] * 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. sharded validity batch with a stand-in checker (the GPU entry point on a real box)
        states = np.arange(101 * 7, dtype=np.float64).reshape(101, 7)
        calls = []

        def check(shard):
            calls.append(len(shard))
            return shard[:, 0] % 3 == 0
        full = sharding.sharded_check(check, states)
        assert (full == (states[:, 0] % 3 == 0)).all()
        assert calls == [sharding.shard_bounds(101, rank, world)[1] - sharding.shard_bounds(101, rank, world)[0]]
        # 2. solution exchange: nobody solved yet
        ex = sharding.SolutionExchange(state_dim=7, max_waypoints=16)
        assert ex.publish(float("inf"), None) == (float("inf"), -1, None)
        # rank 1 has the cheaper path
        path = np.full((3 + rank, 7), float(rank))
        cost, who, best = ex.publish(10.0 - rank, path)
        assert (cost, who) == (9.0, 1) and best.shape == (4, 7) and (best == 1.0).all()
        # tie: lowest rank wins; only one rank has a solution: it wins
        cost, who, best = ex.publish(5.0, path)
        assert who == 0 and best.shape == (3, 7)
        cost, who, best = ex.publish(float("inf") if rank == 0 else 7.5, path)
        assert (cost, who) == (7.5, 1)
        # a path longer than max_waypoints is relayed, not an error on the winner with the others stuck in the broadcast
        long = np.arange(40 * 7, dtype=np.float64).reshape(40, 7)
        cost, who, best = ex.publish(1.0 if rank == 1 else 2.0, long)
        assert who == 1 and best.shape == (40, 7) and (best == long).all()
        # 3. done flag (replaces the coordinator's Done broadcast)
        assert ex.any_done(False) is False
        assert ex.any_done(rank == 1) is True
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res == {0: "ok", 1: "ok"}, res


def _planner_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from cpu_backend import CpuNN, CpuScenario
        from mplambda_b200.planner import BatchedRRT, solve_sharded
        from mplambda_b200.workload import SE3_SCENES
        d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scenes", "cubicles.npz"))
        sc = SE3_SCENES["cubicles"]
        scen = CpuScenario(d["env"], d["robot"], sc, threads=2)
        ex = sharding.SolutionExchange(state_dim=7, max_waypoints=64)
        cost, who, path, rrt = solve_sharded(lambda r: BatchedRRT(scen, CpuNN(2), sc["start"], batch=256, seed=100 + r),
                                             ex, time_limit=120, sync_every=2)
        ok = path is not None and scen.check_edges(path[:-1], path[1:]).all() and scen.isGoal(path[-1])
        q.put((rank, (round(cost, 9), who, path.tobytes() if path is not None else None, bool(ok))))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_planners_agree_on_the_winner():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_planner_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert isinstance(res[0], tuple) and isinstance(res[1], tuple), res
    assert res[0] == res[1]                  # same cost, same winner, same waypoints on both shards
    assert res[0][3] is True and res[0][1] in (0, 1)


def _bridge_worker(rank, world, port, coord_port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mplambda_b200 import bridge
        link = bridge.CoordinatorLink(77, "127.0.0.1", coord_port) if rank == 0 else None
        bx = bridge.BridgedExchange(sharding.SolutionExchange(state_dim=7), link)
        # rank 1 solves first: rank 0 forwards the group's best upstream (once)
        mine = np.full((3, 7), 1.0 + rank)
        cost, who, best = bx.publish(float("inf") if rank == 0 else 8.0, mine, solve_millis=40)
        assert (cost, who) == (8.0, 1)
        cost, who, best = bx.publish(float("inf") if rank == 0 else 8.0, mine, solve_millis=50)   # no improvement: not re-sent
        # the coordinator relays another lambda's cheaper path (the stand-in sends it after the first path it
        # receives): it joins the exchange as rank 0's candidate and wins on every rank
        import time
        for _ in range(100):
            cost, who, best = bx.publish(float("inf") if rank == 0 else 8.0, mine, solve_millis=60)
            if cost < 8.0:
                break
            time.sleep(0.02)
        assert (cost, who) == (2.5, 0) and best.shape == (2, 7) and (best[:, 4] == [4.0, 5.0]).all()
        assert bx.any_done(False) is False
        assert bx.any_done(rank == 1) is True        # rank 0 tells the coordinator
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_group_best_goes_upstream_and_foreign_paths_come_down():
    """BridgedExchange: two ranks over gloo, rank 0 linked to a stand-in coordinator speaking packet.hpp's format."""
    import threading
    from test_bridge import _coordinator
    server = socket.socket()
    server.bind(("127.0.0.1", 0))
    server.listen(1)
    log = []
    foreign = (2.5, np.array([[0, 0, 0, 1.0, 4, 0, 0], [0, 0, 0, 1.0, 5, 0, 0]]))
    t = threading.Thread(target=_coordinator, args=(server, log, foreign), daemon=True)
    t.start()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bridge_worker, args=(r, 2, port, server.getsockname()[1], q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    t.join(10)
    server.close()
    assert res == {0: "ok", 1: "ok"}, res
    kinds = [p[0] for p in log]
    assert kinds == ["hello", "path", "done"], kinds      # one Hello, the group's best once, Done
    assert log[0][1] == 77 and log[1][1] == 8.0 and log[1][2] == 40 and (log[1][3] == 2.0).all()
