"""GPU: the device-resident planner iteration (sampler, k-NN on device queries, index-named motion checks, the RRT loop
of mplb_rrt_*) against the oracle and against the host-buffer entry points.  Device memory comes from torch tensors
(plumbing only)."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import mplambda_b200 as M
from mplambda_b200.planner import DeviceRRT, ReplayRRT
from mplambda_b200.workload import (FETCH_JOINT_MAX, FETCH_JOINT_MIN, SE3_SCENES, SEED, fetch_configs, se3_poses)
from conftest import GOLDEN, ROOT
from cpu_backend import CpuNN, CpuScenario

pytestmark = pytest.mark.gpu


def make(scenes, name, res=0.1):
    env, robot = scenes[name]
    sc = SE3_SCENES[name]
    return M.SE3RigidBodyScenario(env, robot, sc["goal"], sc["min"], sc["max"], res)


def _ulps(a, b):
    return np.abs(a - b) / np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b))), 1e-300)


def test_device_sampler_names_the_host_stream(scenes):
    import torch
    s = make(scenes, "cubicles")
    sc = SE3_SCENES["cubicles"]
    n, first = 50000, 123457
    d = torch.empty((n, 7), dtype=torch.float64, device="cuda")
    s.sample_states_device(SEED + 5, first, n, d.data_ptr())
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    want = se3_poses(n, sc["min"], sc["max"], SEED + 5, start=first)
    np.testing.assert_array_equal(got[:, 4:], want[:, 4:])          # translations: correctly rounded operations only
    assert _ulps(got[:, :4], want[:, :4]).max() <= 4                # sin/cos: CUDA's against glibc's
    assert np.abs(np.linalg.norm(got[:, :4], axis=1) - 1).max() < 1e-12


def test_device_sampler_fetch_is_exact(scenes):
    import torch
    from mplambda_b200.workload import FETCH_TASKS, frame_from_option
    env = scenes["autolab"][0]
    f = M.FetchScenario(frame_from_option(FETCH_TASKS["fetch_task_001"]["env_frame"]), env, checkResolution=0.01)
    lo = np.maximum(FETCH_JOINT_MIN, -4 * math.pi)
    hi = np.minimum(FETCH_JOINT_MAX, 4 * math.pi)
    n = 20000
    d = torch.empty((n, 8), dtype=torch.float64, device="cuda")
    M._capi.check(M._capi.lib().mplb_sample_states_device(f._h, lo.ctypes.data_as(M._capi.dp), hi.ctypes.data_as(M._capi.dp),
                                                          SEED, 7, n, d.data_ptr(), None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d.cpu().numpy(), fetch_configs(n, SEED, start=7))


@pytest.mark.parametrize("name", ["cubicles", "twistycool", "apartment"])
def test_indexed_edges_match_oracle_and_host_path(name, scenes, oracle):
    import torch
    env, robot = scenes[name]
    sc = SE3_SCENES[name]
    s = make(scenes, name)
    pool = se3_poses(6000, sc["min"], sc["max"], SEED + 31)
    tree = pool[s.check_states(pool)][:1500]
    to = se3_poses(700, sc["min"], sc["max"], SEED + 32)
    d_tree, d_to = torch.from_numpy(tree).cuda(), torch.from_numpy(to).cuda()
    rng = np.random.default_rng(1)
    fi = rng.integers(0, len(tree), 2000)
    ti = rng.integers(0, len(to), 2000)
    got, checked = s.check_edges_indexed(d_tree.data_ptr(), len(tree), fi, d_to.data_ptr(), len(to), ti,
                                         return_states_checked=True)
    host = s.check_edges(tree[fi], to[ti])
    want = oracle.SE3Oracle(env, robot, check_resolution=0.1).check_edges(tree[fi], to[ti])
    assert (host == want).all()
    assert (got == want).all(), "%d verdicts differ" % int((got != want).sum())
    assert checked > 0 and 0 < got.sum() < len(got)
    # edge e = node e // k -> its own k targets (the addNodeNear pattern), indices out of range kill the edge
    k = 4
    ti2 = rng.integers(0, len(to), 200 * k)
    ti2[::17] = -1
    ti2[5::31] = len(to)
    got2 = s.check_edges_indexed(d_tree.data_ptr(), len(tree), None, d_to.data_ptr(), len(to), ti2, n=200 * k, from_div=k)
    ok = (ti2 >= 0) & (ti2 < len(to))
    want2 = np.zeros(200 * k, dtype=bool)
    want2[ok] = s.check_edges(np.repeat(tree[:200], k, axis=0)[ok], to[ti2[ok]])
    assert (got2 == want2).all()
    # index arrays and verdicts in device memory
    d_fi, d_ti = torch.from_numpy(fi).cuda(), torch.from_numpy(ti).cuda()
    d_v = torch.empty(len(fi), dtype=torch.uint8, device="cuda")
    s.check_edges_indexed_device(d_tree.data_ptr(), len(tree), d_fi.data_ptr(), d_to.data_ptr(), len(to), d_ti.data_ptr(),
                                 len(fi), d_v.data_ptr(), synchronous=True)
    assert (d_v.cpu().numpy().astype(bool) == want).all()
    d_v.zero_()
    s.check_edges_indexed_device(d_tree.data_ptr(), len(tree), d_fi.data_ptr(), d_to.data_ptr(), len(to), d_ti.data_ptr(),
                                 len(fi), d_v.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (d_v.cpu().numpy().astype(bool) == want).all()
    assert s.stats()["ambiguous_steps"] == 0


def test_ambiguous_step_counts_are_replanned_on_the_host(scenes):
    """The re-planning path (edges whose device-planned step count is within rounding of an integer) forced through
    the debug hooks, in a child process (the hooks are read once per process)."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, torch, mplambda_b200 as M\n"
            "from mplambda_b200.workload import SE3_SCENES, SEED, se3_poses\n"
            "d = np.load(%r); sc = SE3_SCENES['cubicles']\n"
            "s = M.SE3RigidBodyScenario(d['env'], d['robot'], sc['goal'], sc['min'], sc['max'], 0.1)\n"
            "P = se3_poses(3000, sc['min'], sc['max'], SEED + 3); tree = P[s.check_states(P)][:500]\n"
            "to = se3_poses(500, sc['min'], sc['max'], SEED + 4)\n"
            "dt, dq = torch.from_numpy(tree).cuda(), torch.from_numpy(to).cuda()\n"
            "rng = np.random.default_rng(0); fi = rng.integers(0, len(tree), 900); ti = rng.integers(0, 500, 900)\n"
            "got = s.check_edges_indexed(dt.data_ptr(), len(tree), fi, dq.data_ptr(), 500, ti)\n"
            "want = s.check_edges(tree[fi], to[ti])\n"
            "assert (got == want).all(); st = s.stats(); assert st['ambiguous_steps'] >= 1, st; print('ok', st['ambiguous_steps'])\n"
            % (ROOT, os.path.join(ROOT, "tests"), os.path.join(GOLDEN, "scenes", "cubicles.npz")))
    for extra in ({"MPLB_DEBUG_AMBIGUOUS_EVERY": "97"}, {"MPLB_DEBUG_AMBIGUOUS_EVERY": "97", "MPLB_DEBUG_REPLAN_ALL": "1"},
                  {"MPLB_DEBUG_AMBIGUOUS_EVERY": "3", "MPLB_DEBUG_REPLAN_ALL": "1"}):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **extra), capture_output=True, text=True,
                             timeout=600)
        assert out.returncode == 0 and "ok" in out.stdout, (extra, out.stderr[-2000:])


def test_knn_on_device_queries_and_device_insert(scenes, oracle):
    import torch
    sc = SE3_SCENES["home"]
    keys = se3_poses(5000, sc["min"], sc["max"], SEED + 61)
    q = se3_poses(777, sc["min"], sc["max"], SEED + 62)
    nn = M.Nigh(M.nn.SE3)
    nn.reserve(8192)
    nn.insert(keys[:2000])
    d_more = torch.from_numpy(keys[2000:]).cuda()
    assert nn.insert_device(d_more.data_ptr(), 3000) == 2000 and nn.size() == 5000
    k = 26
    wi, wd = oracle.nn_knearest(keys, q, k)
    hi, hd = nn.knearest(q, k)                       # host call after a device-side insert: mirror fetched on demand
    assert (hi == wi).all() and (hd == wd).all()
    d_q = torch.from_numpy(q).cuda()
    d_i = torch.empty((len(q), k), dtype=torch.int64, device="cuda")
    d_d = torch.empty((len(q), k), dtype=torch.float64, device="cuda")
    nn.knearest_device(d_q.data_ptr(), len(q), k, d_i.data_ptr(), d_d.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (d_i.cpu().numpy() == wi).all()
    assert _ulps(d_d.cpu().numpy(), wd).max() <= 4   # device-evaluated distances: CUDA's acos
    # the key pool is what index-named edges read: same rows as inserted
    np.testing.assert_array_equal(_d2h(nn.keys_device(), (5000, 7)), keys)


def _d2h(ptr, shape):
    import ctypes as C
    out = np.empty(shape, dtype=np.float64)
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), out.nbytes, 2) == 0
    return out


def test_knn_then_connect_on_the_device(scenes, oracle):
    """BASELINE config 4 (PCForest addNodeNear, pcforest.hpp:507-567) with nothing but indices moving: k nearest tree
    nodes of every new node by a device query, then the motions new node -> neighbour named by those indices."""
    import torch
    sc = SE3_SCENES["home"]
    env, robot = scenes["home"]
    s = make(scenes, "home")
    pool = se3_poses(8192, sc["min"], sc["max"], SEED)
    tree = pool[s.check_states(pool)][:1024]
    cand = se3_poses(2048, sc["min"], sc["max"], SEED + 2)
    new = cand[s.check_states(cand)][:64]
    k = int(math.ceil(math.e * (7.0 / 6.0) * math.log(len(tree) + 1)))
    nn = M.Nigh(M.nn.SE3)
    nn.insert(tree)
    d_new = torch.from_numpy(new).cuda()
    d_i = torch.empty((len(new), k), dtype=torch.int64, device="cuda")
    d_d = torch.empty((len(new), k), dtype=torch.float64, device="cuda")
    d_v = torch.empty(len(new) * k, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    nn.knearest_device(d_new.data_ptr(), len(new), k, d_i.data_ptr(), d_d.data_ptr(), st)
    s.check_edges_indexed_device(d_new.data_ptr(), len(new), None, nn.keys_device(), nn.size(), d_i.data_ptr(),
                                 len(new) * k, d_v.data_ptr(), from_div=k, stream=st)
    torch.cuda.synchronize()
    wi, _ = oracle.nn_knearest(tree, new, k)
    assert (d_i.cpu().numpy() == wi).all()
    want = oracle.SE3Oracle(env, robot, check_resolution=0.1).check_edges(np.repeat(new, k, axis=0), tree[wi.reshape(-1)])
    got = d_v.cpu().numpy().astype(bool)
    assert (got == want).all() and 0 < got.sum() < len(got)


@pytest.mark.parametrize("name,batch", [("cubicles", 1024), ("home", 512), ("twistycool", 2048)])
def test_device_rrt_builds_the_tree_of_the_cpu_loop(name, batch, scenes):
    """mplb_rrt_* (everything on the device) against the batch-issuing loop on the CPU oracle, fed with the same sample
    batches: same nodes, same parents, same path; the path is valid under the oracle."""
    env, robot = scenes[name]
    sc = SE3_SCENES[name]
    g = make(scenes, name)
    dev = DeviceRRT(g, M.Nigh(M.nn.SE3), sc["start"], batch=batch, goal_bias=0.01, seed=11)
    path = dev.solve(time_limit=120)
    assert path is not None and dev.solution >= 0
    nodes, parent = dev.tree()
    c = CpuScenario(env, robot, sc)
    cpu = ReplayRRT(c, CpuNN(), sc["start"], dev)
    cpath = cpu.solve(time_limit=600, max_iterations=dev.stats["iterations"])
    assert cpu.n == len(nodes) and cpu.stats["iterations"] == dev.stats["iterations"]
    np.testing.assert_array_equal(cpu.nodes[:cpu.n], nodes)
    np.testing.assert_array_equal(cpu.parent[:cpu.n], parent)
    assert cpu.solution == dev.solution
    np.testing.assert_array_equal(cpath, path)
    np.testing.assert_array_equal(path[0], np.array(sc["start"]))
    assert c.isGoal(path[-1]) and c.check_states(path).all() and c.check_edges(path[:-1], path[1:]).all()
    assert dev.stats["edges_checked"] == cpu.stats["edges_checked"]


def test_device_rrt_rejects_an_invalid_start(scenes):
    sc = SE3_SCENES["cubicles"]
    g = make(scenes, "cubicles")
    cand = se3_poses(256, sc["min"], sc["max"], SEED)
    bad = cand[~g.check_states(cand)][0]
    with pytest.raises(ValueError):                 # prrt.hpp:109-110
        DeviceRRT(g, M.Nigh(M.nn.SE3), bad)
