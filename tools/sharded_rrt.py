"""Sharded planner instances, one per GPU, exchanging solutions over NCCL (run under torchrun)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import mplambda_b200 as M
from mplambda_b200 import sharding
from mplambda_b200.planner import BatchedRRT, solve_sharded
from mplambda_b200.workload import SE3_SCENES

rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for name in sys.argv[1:] or ["twistycool", "home"]:
    d = np.load("tests/golden/scenes/%s.npz" % name); sc = SE3_SCENES[name]
    scen = M.SE3RigidBodyScenario(d["env"], d["robot"], sc["goal"], sc["min"], sc["max"], 0.1, device=lr)
    ex = sharding.SolutionExchange(state_dim=7, max_waypoints=256)
    ex.any_done(False)   # warm the communicator up
    torch.cuda.synchronize(); dist.barrier()
    t = time.perf_counter()
    cost, who, path, rrt = solve_sharded(lambda r: BatchedRRT(scen, M.Nigh(M.nn.SE3, device=lr), sc["start"], batch=2048, seed=10 + r), ex, time_limit=60)
    dt = time.perf_counter() - t
    ok = path is not None and scen.check_edges(path[:-1], path[1:]).all()
    print(json.dumps({"scene": name, "rank": rank, "winner": who, "cost": cost, "waypoints": None if path is None else len(path),
                      "path_valid": bool(ok), "seconds": dt, "my_iterations": rrt.stats["iterations"], "my_nodes": rrt.n}), flush=True)
dist.destroy_process_group()
