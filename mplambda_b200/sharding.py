"""Multi-GPU plumbing for the validity path: one process per GPU, scene
replicated, query batches split by index, no collective on the check path.  The only exchange
step mirrors what the reference's coordinator relays between lambdas over TCP -- solution paths
to the other group members and the Done broadcast (reference src/mpl_coordinator.cpp:529-589,
include/mpl/comm.hpp:80-106) -- done here with torch.distributed collectives (NCCL over NVLink on
GPUs, gloo in the CPU tests).  torch is plumbing only; verdicts come from libmplb.so.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous split of n queries: rank g owns [g*n//G, (g+1)*n//G)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def balanced_edge_bounds(steps, world):
    """Contiguous split of an edge batch balanced by state count (prefix sum of `steps`, each edge
    also costing its `to` check) instead of edge count.  -> list of (lo, hi) per rank."""
    cost = np.asarray(steps, dtype=np.int64) + 1
    csum = np.concatenate([[0], np.cumsum(cost)])
    total = csum[-1]
    cuts = [0]
    for g in range(1, world):
        cuts.append(int(np.searchsorted(csum, total * g / world, side="left")))
    cuts.append(len(cost))
    cuts = np.maximum.accumulate(np.minimum(cuts, len(cost)))
    return [(int(cuts[g]), int(cuts[g + 1])) for g in range(world)]


def _device_for(group):
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def sharded_check(check_fn, states, group=None):
    """Every rank checks its contiguous shard with `check_fn(states_shard) -> bool[n]` and the verdicts
    are all-gathered, so each rank ends up with the full verdict vector (what a planner instance that
    issued the batch needs back)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    states = np.asarray(states)
    n = len(states)
    lo, hi = shard_bounds(n, rank, world)
    local = np.asarray(check_fn(states[lo:hi]), dtype=np.uint8)
    dev = _device_for(group)
    cap = -(-n // world) + 1
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    buf[: hi - lo] = torch.from_numpy(local).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    full = np.empty(n, dtype=bool)
    for g in range(world):
        glo, ghi = shard_bounds(n, g, world)
        full[glo:ghi] = out[g][: ghi - glo].cpu().numpy().astype(bool)
    return full


class SolutionExchange:
    """Best-solution exchange between planner shards (C-FOREST style): every rank contributes the cost
    of its current best path; the cheapest wins (ties: lowest rank) and its waypoints are broadcast.
    Replaces the coordinator's gotPath relay (mpl_coordinator.cpp:573-589) and, with `any_done`, the
    Done broadcast (mpl_coordinator.cpp:538-551)."""

    def __init__(self, state_dim, max_waypoints=256, group=None):
        self.group = group
        self.dim, self.cap = state_dim, max_waypoints   # cap: kept for callers; paths of any length are relayed
        self.dev = _device_for(group)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def publish(self, cost, path):
        """cost: float (inf when this rank has no solution); path: [m, dim] array or None.
        -> (best_cost, best_rank, best_path or None), identical on every rank.  Costs and path lengths travel
        together, so the broadcast buffer is sized from the winner's length on every rank: no rank can fail
        (and leave the others waiting in the broadcast) because of a long path."""
        p = None if path is None else np.asarray(path, dtype=np.float64).reshape(-1, self.dim)
        mine = torch.tensor([float(cost), 0.0 if p is None else float(len(p))], dtype=torch.float64, device=self.dev)
        both = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(both, mine, group=self.group)
        both = torch.stack(both).cpu().numpy()
        costs = both[:, 0]
        if not np.isfinite(costs).any():
            return float("inf"), -1, None
        best = int(np.argmin(costs))          # first minimum = lowest rank on ties
        m = int(both[best, 1])
        buf = torch.zeros(max(1, m * self.dim), dtype=torch.float64, device=self.dev)
        if self.rank == best and m:
            buf[:p.size] = torch.from_numpy(p.reshape(-1)).to(self.dev)
        src = best if self.group is None else dist.get_global_rank(self.group, best)
        dist.broadcast(buf, src=src, group=self.group)
        return float(costs[best]), best, buf[:m * self.dim].cpu().numpy().reshape(m, self.dim)

    def any_done(self, done):
        flag = torch.tensor([1 if done else 0], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        return bool(flag.item())
