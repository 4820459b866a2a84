"""ORACLE -- test infrastructure only: ctypes binding of oracle/_build/libmplb_oracle.so.

PARITY UNPINNED (see oracle/mplb_oracle.h): FCL/libccd/nigh/Eigen are not available, the C code
restates their published algorithms.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmplb_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("mplb_oracle.c", "mplb_oracle_fetch.inc", "mplb_oracle.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


class Counters(C.Structure):
    _fields_ = [("checks", C.c_uint64), ("bv_tests", C.c_uint64), ("tri_tests", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, u8p, i64p = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int64)
        cp = C.POINTER(Counters)
        L.orc_mesh_create.restype = C.c_void_p
        L.orc_mesh_create.argtypes = [dp, C.c_size_t]
        L.orc_mesh_destroy.argtypes = [C.c_void_p]
        L.orc_mesh_num_nodes.restype = C.c_size_t
        L.orc_mesh_num_nodes.argtypes = [C.c_void_p]
        L.orc_se3_collide_brute.argtypes = [C.c_void_p, C.c_void_p, dp]
        L.orc_se3_collide_bvh.argtypes = [C.c_void_p, C.c_void_p, dp, cp]
        L.orc_se3_check_states.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_size_t, u8p, C.c_int, C.c_int, cp]
        L.orc_se3_distance.restype = C.c_double
        L.orc_se3_distance.argtypes = [dp, dp, C.c_double, C.c_double]
        L.orc_se3_interpolate.argtypes = [dp, dp, C.c_double, dp]
        L.orc_se3_edge_steps.restype = C.c_uint64
        L.orc_se3_edge_steps.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double]
        L.orc_se3_check_edge.argtypes = [C.c_void_p, C.c_void_p, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, cp]
        L.orc_se3_check_edges.argtypes = [C.c_void_p, C.c_void_p, dp, dp, C.c_size_t, C.c_double, C.c_double, C.c_double, u8p, C.c_int, C.c_int, cp]
        L.orc_nn_nearest.argtypes = [dp, C.c_size_t, C.c_int, C.c_int, C.c_double, C.c_double, dp, C.c_size_t, i64p, dp, C.c_int]
        L.orc_nn_knearest.argtypes = [dp, C.c_size_t, C.c_int, C.c_int, C.c_double, C.c_double, dp, C.c_size_t, C.c_int, i64p, dp, C.c_int]
        L.orc_omp_max_threads.restype = C.c_int
        ip = C.POINTER(C.c_int)
        L.orc_fetch_fk.argtypes = [dp, dp, dp]
        L.orc_fetch_self_collision.argtypes = [dp]
        L.orc_fetch_is_valid.argtypes = [C.c_void_p, dp, dp, ip, cp]
        L.orc_fetch_check_states.argtypes = [C.c_void_p, dp, dp, C.c_size_t, u8p, C.c_int, cp]
        L.orc_fetch_distance.restype = C.c_double
        L.orc_fetch_distance.argtypes = [dp, dp]
        L.orc_fetch_edge_steps.restype = C.c_uint64
        L.orc_fetch_edge_steps.argtypes = [dp, dp, C.c_double]
        L.orc_fetch_check_edge.argtypes = [C.c_void_p, dp, dp, dp, C.c_double, cp]
        L.orc_fetch_check_edges.argtypes = [C.c_void_p, dp, dp, dp, C.c_size_t, C.c_double, u8p, C.c_int, cp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Mesh:
    def __init__(self, tris):
        t = _f64(tris, (-1, 3, 3))
        self.ntris = len(t)
        self.h = lib().orc_mesh_create(_dp(t), len(t))

    @property
    def nnodes(self):
        return lib().orc_mesh_num_nodes(self.h)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_mesh_destroy(self.h)
            self.h = None


BRUTE, BVH = 0, 1


class SE3Oracle:
    """FP64 restatement of SE3RigidBodyScenario's validity methods
    (se3_rigid_body_scenario.hpp:119-178)."""

    def __init__(self, env_tris, robot_tris, so3_weight=50.0, l2_weight=1.0, check_resolution=0.1):
        self.env, self.robot = Mesh(env_tris), Mesh(robot_tris)
        self.so3w, self.l2w = float(so3_weight), float(l2_weight)
        self.inv_step = 1.0 / check_resolution

    def check_states(self, states, mode=BVH, threads=0, counters=None):
        s = _f64(states, (-1, 7))
        out = np.empty(len(s), dtype=np.uint8)
        ctr = counters if counters is not None else Counters()
        lib().orc_se3_check_states(self.robot.h, self.env.h, _dp(s), len(s),
                                   out.ctypes.data_as(C.POINTER(C.c_uint8)), mode, threads, C.byref(ctr))
        return out.astype(bool)

    def check_edges(self, frm, to, mode=BVH, threads=0, counters=None):
        a, b = _f64(frm, (-1, 7)), _f64(to, (-1, 7))
        out = np.empty(len(a), dtype=np.uint8)
        ctr = counters if counters is not None else Counters()
        lib().orc_se3_check_edges(self.robot.h, self.env.h, _dp(a), _dp(b), len(a), self.so3w, self.l2w,
                                  self.inv_step, out.ctypes.data_as(C.POINTER(C.c_uint8)), mode, threads,
                                  C.byref(ctr))
        return out.astype(bool)

    def distance(self, a, b):
        return lib().orc_se3_distance(_dp(_f64(a)), _dp(_f64(b)), self.so3w, self.l2w)

    def edge_steps(self, a, b):
        return lib().orc_se3_edge_steps(_dp(_f64(a)), _dp(_f64(b)), self.so3w, self.l2w, self.inv_step)

    @staticmethod
    def interpolate(a, b, t):
        out = np.empty(7)
        lib().orc_se3_interpolate(_dp(_f64(a)), _dp(_f64(b)), float(t), _dp(out))
        return out


def nn_nearest(keys, queries, metric="se3", so3w=50.0, l2w=1.0, threads=0):
    keys, queries = _f64(keys), _f64(queries)
    dim = keys.shape[1]
    idx = np.empty(len(queries), dtype=np.int64)
    dist = np.empty(len(queries))
    lib().orc_nn_nearest(_dp(keys), len(keys), dim, 0 if metric == "se3" else 1, so3w, l2w, _dp(queries),
                         len(queries), idx.ctypes.data_as(C.POINTER(C.c_int64)), _dp(dist), threads)
    return idx, dist


def nn_knearest(keys, queries, k, metric="se3", so3w=50.0, l2w=1.0, threads=0):
    keys, queries = _f64(keys), _f64(queries)
    dim = keys.shape[1]
    idx = np.empty((len(queries), k), dtype=np.int64)
    dist = np.empty((len(queries), k))
    lib().orc_nn_knearest(_dp(keys), len(keys), dim, 0 if metric == "se3" else 1, so3w, l2w, _dp(queries),
                          len(queries), k, idx.ctypes.data_as(C.POINTER(C.c_int64)), _dp(dist), threads)
    return idx, dist


class FetchOracle:
    """FP64 restatement of FetchScenario's validity methods (fetch_scenario.hpp:105-173,
    fetch_robot.hpp:276-315,502-647).  env_frame: 3x4 [R|t] of the environment mesh."""

    def __init__(self, env_tris, env_frame, check_resolution=0.01):
        self.env = Mesh(env_tris)
        self.frame = _f64(env_frame, (12,))
        self.inv_step = 1.0 / check_resolution

    @staticmethod
    def fk(q):
        frames = np.empty((12, 12))
        gz = C.c_double()
        lib().orc_fetch_fk(_dp(_f64(q, (8,))), _dp(frames), C.byref(gz))
        return frames.reshape(12, 3, 4), gz.value

    @staticmethod
    def self_collision(q):
        return lib().orc_fetch_self_collision(_dp(_f64(q, (8,))))

    def why(self, q):
        w = C.c_int()
        lib().orc_fetch_is_valid(self.env.h, _dp(self.frame), _dp(_f64(q, (8,))), C.byref(w), None)
        return w.value

    def check_states(self, states, threads=0, counters=None):
        s = _f64(states, (-1, 8))
        out = np.empty(len(s), dtype=np.uint8)
        ctr = counters if counters is not None else Counters()
        lib().orc_fetch_check_states(self.env.h, _dp(self.frame), _dp(s), len(s),
                                     out.ctypes.data_as(C.POINTER(C.c_uint8)), threads, C.byref(ctr))
        return out.astype(bool)

    def check_edges(self, frm, to, threads=0, counters=None):
        a, b = _f64(frm, (-1, 8)), _f64(to, (-1, 8))
        out = np.empty(len(a), dtype=np.uint8)
        ctr = counters if counters is not None else Counters()
        lib().orc_fetch_check_edges(self.env.h, _dp(self.frame), _dp(a), _dp(b), len(a), self.inv_step,
                                    out.ctypes.data_as(C.POINTER(C.c_uint8)), threads, C.byref(ctr))
        return out.astype(bool)

    def edge_steps(self, a, b):
        return lib().orc_fetch_edge_steps(_dp(_f64(a, (8,))), _dp(_f64(b, (8,))), self.inv_step)
