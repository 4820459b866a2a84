"""Test infrastructure: oracle-backed stand-ins with the scenario / nearest-neighbour interface, used to
run the batch-issuing planner on the CPU (tests, tests/tools/rrt_compare.py).  Never used by the product."""
import numpy as np

from oracle import oracle_py as O
from mplambda_b200.planner import se3_random_batch


class CpuScenario:
    def __init__(self, env, robot, sc, threads=0, resolution=0.1):
        self.orc = O.SE3Oracle(env, robot, check_resolution=resolution)
        self.goal_, self.min_, self.max_ = np.array(sc["goal"]), np.array(sc["min"]), np.array(sc["max"])
        self.threads = threads

    @staticmethod
    def scale(q):
        return q

    def check_states(self, s):
        return self.orc.check_states(np.atleast_2d(s), O.BVH, threads=self.threads)

    def check_edges(self, a, b):
        return self.orc.check_edges(a, b, O.BVH, threads=self.threads)

    def isGoal(self, q):
        return self.orc.distance(self.goal_, q) <= 0.1

    def sampleGoal(self, rng):
        return self.goal_.copy()

    def sample_batch(self, rng, n):
        return se3_random_batch(rng, n, self.min_, self.max_)


class CpuNN:
    def __init__(self, threads=0):
        self.keys, self.threads = np.zeros((0, 7)), threads

    def insert(self, k):
        self.keys = np.concatenate([self.keys, np.atleast_2d(k)])

    def nearest(self, q):
        return O.nn_nearest(self.keys, q, threads=self.threads)

    def knearest(self, q, k):
        return O.nn_knearest(self.keys, q, k, threads=self.threads)


class _Space:
    def __init__(self, orc):
        self.orc = orc

    def distance(self, a, b):
        return self.orc.distance(a, b)


def _space(self):
    return _Space(self.orc)


CpuScenario.space = _space
