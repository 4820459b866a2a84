"""GPU: nearest-neighbour kernels (C ABI mplb_nn_*) vs the oracle's exact brute force.
Bar: indices bit-exact (ties to the lowest index), k-NN order identical; distances are
recomputed on the host with the CPU path's arithmetic, so they are bit-exact too."""
import numpy as np
import pytest

import mplambda_b200 as M
from mplambda_b200.workload import se3_poses

pytestmark = pytest.mark.gpu

LO, HI = [-300, -200, -100], [300, 200, 100]


@pytest.mark.parametrize("n,nq", [(1, 5), (7, 3), (1000, 257), (16384, 4096), (100003, 513)])
def test_se3_nearest_matches_oracle(n, nq, oracle):
    keys = se3_poses(n, LO, HI, 31)
    qs = se3_poses(nq, LO, HI, 32)
    nn = M.Nigh(M.nn.SE3)
    assert nn.size() == 0
    idx, dist = nn.nearest(qs[:2])
    assert (idx == -1).all() and np.isinf(dist).all()            # empty structure: no neighbour
    # insert in two batches (the planner inserts incrementally)
    assert nn.insert(keys[: n // 2]) == 0
    assert nn.insert(keys[n // 2:]) == n // 2
    assert nn.size() == n
    idx, dist = nn.nearest(qs)
    widx, wdist = oracle.nn_nearest(keys, qs)
    assert (idx == widx).all()
    assert (dist == wdist).all()


@pytest.mark.parametrize("n,nq,k", [(50, 20, 7), (5000, 300, 32), (20000, 1000, 45), (10, 4, 16)])
def test_se3_knearest_matches_oracle(n, nq, k, oracle):
    keys = se3_poses(n, LO, HI, 41)
    qs = se3_poses(nq, LO, HI, 42)
    nn = M.Nigh(M.nn.SE3)
    nn.insert(keys)
    idx, dist = nn.knearest(qs, k)
    widx, wdist = oracle.nn_knearest(keys, qs, k)
    assert (idx == widx).all()
    assert (dist == wdist).all()
    assert (np.diff(dist[:, :min(k, n)], axis=1) >= 0).all()      # ascending: pcforest.hpp:523-524 relies on it


def test_l1_scaled_keys_and_ties(oracle):
    rng = np.random.default_rng(3)
    scale = np.array([10, 4, 2, 1, 1, 1, 1, 1.0])                 # FetchScenario::scale (fetch_scenario.hpp:42-45)
    keys = rng.uniform(-2, 2, size=(3000, 8)) * scale
    keys = np.concatenate([keys, keys[:100]])                     # exact duplicates: ties go to the lowest index
    qs = np.concatenate([rng.uniform(-2, 2, size=(500, 8)) * scale, keys[:100]])
    nn = M.Nigh(M.nn.L1, dim=8)
    nn.insert(keys)
    idx, dist = nn.nearest(qs)
    widx, wdist = oracle.nn_nearest(keys, qs, metric="l1")
    assert (idx == widx).all() and (dist == wdist).all()
    assert (idx[-100:] == np.arange(100)).all() and (dist[-100:] == 0).all()
    kidx, kdist = nn.knearest(qs, 9)
    wkidx, wkdist = oracle.nn_knearest(keys, qs, 9, metric="l1")
    assert (kidx == wkidx).all() and (kdist == wkdist).all()


def test_device_resident_queries(oracle):
    import torch
    keys = se3_poses(30000, LO, HI, 51)
    qs = se3_poses(2000, LO, HI, 52)
    nn = M.Nigh(M.nn.SE3)
    nn.insert(keys)
    dq = torch.from_numpy(qs).cuda()
    di = torch.empty(len(qs), dtype=torch.int64, device="cuda")
    dd = torch.empty(len(qs), dtype=torch.float64, device="cuda")
    nn.nearest_device(dq.data_ptr(), len(qs), di.data_ptr(), dd.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    widx, wdist = oracle.nn_nearest(keys, qs)
    assert (di.cpu().numpy() == widx).all()
    np.testing.assert_allclose(dd.cpu().numpy(), wdist, rtol=4e-16, atol=0)


def test_bad_arguments():
    nn = M.Nigh(M.nn.SE3)
    nn.insert(se3_poses(10, LO, HI, 1))
    with pytest.raises(M.MplbError):
        nn.knearest(se3_poses(2, LO, HI, 2), 65)
    with pytest.raises(ValueError):
        nn.insert(np.zeros((3, 5)))


def test_non_unit_quaternions_fall_back_to_exact(oracle):
    keys = se3_poses(3000, LO, HI, 61)
    keys[::3, :4] *= 1.7                      # scaled quaternions: |q.q'| may exceed 1, the FP32 bound is off
    qs = se3_poses(300, LO, HI, 62)
    qs[::2, :4] *= 0.6
    nn = M.Nigh(M.nn.SE3)
    nn.insert(keys)
    idx, dist = nn.nearest(qs)
    widx, wdist = oracle.nn_nearest(keys, qs)
    assert (idx == widx).all() and (dist == wdist).all()


def test_clustered_keys_near_ties(oracle):
    # many keys within the FP32 margin of each other: the candidate pass must keep them all
    rng = np.random.default_rng(7)
    base = se3_poses(64, LO, HI, 63)
    keys = np.repeat(base, 50, axis=0)
    keys[:, 4:] += rng.normal(scale=1e-5, size=(len(keys), 3))
    qs = base + 0.0
    qs[:, 4:] += rng.normal(scale=1e-5, size=(len(qs), 3))
    nn = M.Nigh(M.nn.SE3)
    nn.insert(keys)
    idx, dist = nn.knearest(qs, 20)
    widx, wdist = oracle.nn_knearest(keys, qs, 20)
    assert (idx == widx).all() and (dist == wdist).all()


def test_knearest_small_sets_and_duplicates(oracle):
    """k larger than / equal to the number of keys, an empty query batch, odd k values around the sampled-threshold
    path, and many identical keys (ties to the lowest index)."""
    keys = se3_poses(37, LO, HI, 3)
    qs = se3_poses(9, LO, HI, 4)
    nn = M.Nigh(M.nn.SE3)
    nn.insert(keys)
    i, d = nn.knearest(qs, 50)
    wi, wd = oracle.nn_knearest(keys, qs, 50)
    assert (i == wi).all() and (d[:, :37] == wd[:, :37]).all() and (i[:, 37:] == -1).all() and np.isinf(d[:, 37:]).all()
    i, d = nn.knearest(qs, 37)
    wi, wd = oracle.nn_knearest(keys, qs, 37)
    assert (i == wi).all() and (d == wd).all()
    assert nn.knearest(qs[:0], 5)[0].shape == (0, 5)
    keys = se3_poses(5000, LO, HI, 5)
    nn2 = M.Nigh(M.nn.SE3)
    nn2.insert(keys)
    for k in (2, 3, 17, 64):
        i, d = nn2.knearest(qs, k)
        wi, wd = oracle.nn_knearest(keys, qs, k)
        assert (i == wi).all() and (d == wd).all(), k
    dup = np.repeat(keys[:10], 50, axis=0)
    nn3 = M.Nigh(M.nn.SE3)
    nn3.insert(dup)
    i, d = nn3.knearest(keys[:10], 20)
    wi, wd = oracle.nn_knearest(dup, keys[:10], 20)
    assert (i == wi).all() and (d == wd).all()


# ---- the spatial index (nn_grid.cu): same results as brute force, bit for bit -------------------------------------------

def _both(keys, batches=None):
    """the same keys in an indexed structure and in one kept on brute force"""
    a, b = M.Nigh(M.nn.SE3), M.Nigh(M.nn.SE3)
    b.set_index(False)
    for part in (batches or [keys]):
        a.insert(part)
        b.insert(part)
    return a, b


@pytest.mark.parametrize("k", [1, 8, 45, 64])
def test_index_equals_brute_force(k, oracle):
    n = 200000
    keys = se3_poses(n, LO, HI, 51)
    qs = np.concatenate([se3_poses(1500, LO, HI, 52),
                         se3_poses(300, [-900, -700, -500], [900, 700, 500], 53),   # queries far outside the keys' box
                         keys[:200]])                                              # queries that are keys
    a, b = _both(keys)
    ia, da = a.knearest(qs, k)
    ib, db = b.knearest(qs, k)
    assert a.index_info()["indexed"] == n and b.index_info()["indexed"] == 0
    assert (ia == ib).all() and (da == db).all()
    if k == 1:
        i1, d1 = a.nearest(qs)
        assert (i1 == ia[:, 0]).all() and (d1 == da[:, 0]).all()
    sub = slice(0, 40)
    wi, wd = oracle.nn_knearest(keys, qs[sub], k)
    assert (ia[sub] == wi).all() and (da[sub] == wd).all()
    # far fewer exact evaluations than keys x queries, and no query needed the slow selection
    info = a.index_info()
    assert info["exact_evaluations"] < 0.001 * n * len(qs) and info["slow_queries"] == 0, info


def test_index_with_inserts_in_between():
    """The planner's pattern: search, insert a few nodes, search again -- the keys inserted since the last build are
    searched by brute force and merged; the index is rebuilt when they outgrow a quarter of it."""
    keys = se3_poses(90000, LO, HI, 61)
    qs = se3_poses(700, LO, HI, 62)
    a, b = _both(keys[:30000])
    for lo, hi in ((30000, 30000), (30000, 33000), (33000, 36500), (36500, 60000), (60000, 90000)):
        if hi > lo:
            a.insert(keys[lo:hi]); b.insert(keys[lo:hi])
        for k in (1, 20):
            ia, da = a.knearest(qs, k)
            ib, db = b.knearest(qs, k)
            assert (ia == ib).all() and (da == db).all(), (lo, hi, k)
        info = a.index_info()
        assert 0 < info["indexed"] <= hi and hi - info["indexed"] <= max(4096, info["indexed"] // 4)


def test_index_duplicates_ties_and_odd_queries():
    keys = se3_poses(40000, LO, HI, 71)
    keys[1000:1700] = keys[999]                       # 701 copies of one key: more candidates than a warp's list holds
    keys[5000:5040, 4:] = keys[4999, 4:]              # same translation, different rotations
    qs = np.concatenate([keys[999:1001], keys[4999:5001], se3_poses(50, LO, HI, 72)])
    odd = se3_poses(6, LO, HI, 73)
    odd[:3, :4] *= 1.7                                # non-unit quaternions: no FP32 filtering, everything exact
    qs = np.concatenate([qs, odd])
    a, b = _both(keys)
    for k in (1, 64):
        ia, da = a.knearest(qs, k)
        ib, db = b.knearest(qs, k)
        assert (ia == ib).all() and (da == db).all()
    assert (a.knearest(qs[:1], 64)[0][0] == np.arange(999, 999 + 64)).all()   # ties in insertion order
    assert a.index_info()["slow_queries"] > 0


def test_index_needs_spread_out_translations():
    """All keys at (nearly) one translation: nothing to index, brute force answers."""
    keys = se3_poses(20000, LO, HI, 81)
    keys[:, 4:] = [1.0, 2.0, 3.0]
    qs = se3_poses(100, LO, HI, 82)
    a, b = _both(keys)
    ia, da = a.knearest(qs, 5)
    ib, db = b.knearest(qs, 5)
    assert a.index_info()["indexed"] == 0
    assert (ia == ib).all() and (da == db).all()


def test_index_device_queries():
    import torch
    keys = se3_poses(60000, LO, HI, 91)
    qs = se3_poses(512, LO, HI, 92)
    a, b = _both(keys)
    dq = torch.from_numpy(qs).cuda()
    k = 12
    di = torch.empty((len(qs), k), dtype=torch.int64, device="cuda")
    dd = torch.empty((len(qs), k), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    a.knearest_device(dq.data_ptr(), len(qs), k, di.data_ptr(), dd.data_ptr(), st)
    torch.cuda.synchronize()
    ib, db = b.knearest(qs, k)
    assert (di.cpu().numpy() == ib).all()
    assert np.allclose(dd.cpu().numpy(), db, rtol=0, atol=1e-9)   # device acos vs host libm: last bits


@pytest.mark.parametrize("k", [1, 9, 40])
def test_l1_index_equals_brute_force(k, oracle):
    """L1Space<S, 8> keys (the Fetch planners' scaled joint space): the grid lies over the first three coordinates."""
    rng = np.random.default_rng(5)
    scale = np.array([10, 4, 2, 1, 1, 1, 1, 1.0])                 # FetchScenario::scale (fetch_scenario.hpp:42-45)
    lo = np.array([0.0, -1.6, -1.2, -3.1, -2.2, -3.1, -2.1, -3.1]) * scale
    hi = np.array([0.38, 1.6, 1.5, 3.1, 2.2, 3.1, 2.1, 3.1]) * scale
    keys = rng.uniform(lo, hi, size=(80000, 8))
    keys[3000:3300] = keys[2999]                                  # duplicates: ties in insertion order
    qs = np.concatenate([rng.uniform(lo, hi, size=(600, 8)), rng.uniform(lo - 5, hi + 5, size=(100, 8)), keys[2999:3001]])
    a, b = M.Nigh(M.nn.L1, dim=8), M.Nigh(M.nn.L1, dim=8)
    b.set_index(False)
    a.insert(keys[:50000]); b.insert(keys[:50000])
    assert (a.knearest(qs, k)[0] == b.knearest(qs, k)[0]).all()    # builds the index
    a.insert(keys[50000:]); b.insert(keys[50000:])                # 30,000 more: rebuilt at the next search
    ia, da = a.knearest(qs, k)
    ib, db = b.knearest(qs, k)
    assert a.index_info()["indexed"] == len(keys) and b.index_info()["indexed"] == 0
    assert (ia == ib).all() and (da == db).all()
    wi, wd = oracle.nn_knearest(keys, qs[:30], k, metric="l1")
    assert (ia[:30] == wi).all() and (da[:30] == wd).all()
    if k >= 9:
        assert (ia[-2, :9] == np.arange(2999, 3008)).all()
