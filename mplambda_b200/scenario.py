"""Host-side mirror of the reference's scenario interface over the C ABI.

`SE3RigidBodyScenario` keeps the names, argument meaning and error behaviour of
`mpl::demo::SE3RigidBodyScenario<S,50,1>` (reference include/mpl/demo/se3_rigid_body_scenario.hpp:15-179)
so that planner code written against the Scenario concept (prrt.hpp:52-76, pcforest.hpp:80-104)
reads the same; the batched methods (`check_states`, `check_edges`) are what the hot path adds
underneath the unchanged one-at-a-time calls.  All validity work happens in libmplb.so on the
GPU; nothing here computes a verdict.
"""
import ctypes as C
import math

import numpy as np

from . import _capi


def load_collada(path, shift_to_center=False, identity_root_transform=False):
    """MeshLoad::load (load_mesh.hpp:232-293) -> (triangles float64 [n,3,3], centre [3])."""
    L = _capi.lib()
    ptr = _capi.dp()
    n = C.c_size_t()
    center = (C.c_double * 3)()
    _capi.check(L.mplb_mesh_load_collada(str(path).encode(), int(shift_to_center), int(identity_root_transform),
                                         C.byref(ptr), C.byref(n), center))
    try:
        tris = np.ctypeslib.as_array(ptr, shape=(n.value * 9,)).copy().reshape(-1, 3, 3)
    finally:
        L.mplb_free(ptr)
    return tris, np.array(center[:])


def _f64(a, cols):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.ndim != 2 or a.shape[1] != cols:
        raise ValueError("expected [n,%d] states, got %s" % (cols, a.shape))
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class SE3Space:
    """nigh::SE3Space<S, so3weight, l2weight> distance (se3_rigid_body_scenario.hpp:18)."""

    def __init__(self, scenario):
        self._sc = scenario

    def distance(self, a, b):
        a, b = _f64(a, 7), _f64(b, 7)
        return _capi.lib().mplb_distance(self._sc._h, a.ctypes.data_as(_capi.dp), b.ctypes.data_as(_capi.dp))

    @staticmethod
    def dimensions():
        return 6


class SE3RigidBodyScenario:
    """State = (qx,qy,qz,qw,tx,ty,tz) float64, the Eigen coefficient order of
    std::tuple<Quaterniond, Vector3d> (se3...:18-19)."""

    multiGoal = True  # se3...:88

    def __init__(self, envMesh, robotMesh, goal, min, max, checkResolution=0.1, so3weight=50.0, l2weight=1.0,
                 device=0):
        L = _capi.lib()
        self._h = C.c_void_p()
        self.goal_ = np.asarray(goal, dtype=np.float64).reshape(7)
        self.min_ = np.asarray(min, dtype=np.float64).reshape(3)
        self.max_ = np.asarray(max, dtype=np.float64).reshape(3)
        self.goalRadius_ = 0.1  # se3...:39
        self.invStepSize_ = 1.0 / checkResolution  # se3...:63
        self.device = device
        if isinstance(envMesh, (str, bytes)) or hasattr(envMesh, "__fspath__"):
            rc = L.mplb_scene_create_se3_files(str(envMesh).encode(), str(robotMesh).encode(), so3weight, l2weight,
                                               checkResolution, device, C.byref(self._h))
        else:
            env = np.ascontiguousarray(envMesh, dtype=np.float64).reshape(-1, 3, 3)
            rob = np.ascontiguousarray(robotMesh, dtype=np.float64).reshape(-1, 3, 3)
            rc = L.mplb_scene_create_se3(env.ctypes.data_as(_capi.dp), len(env), rob.ctypes.data_as(_capi.dp),
                                         len(rob), so3weight, l2weight, checkResolution, device, C.byref(self._h))
        if rc in (_capi.ERR_IO, _capi.ERR_INVALID):
            # load_mesh.hpp:250-254 throws std::invalid_argument
            raise ValueError(L.mplb_last_error().decode("utf-8", "replace"))
        _capi.check(rc)
        self._space = SE3Space(self)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _capi.lib().mplb_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the Scenario concept ---------------------------------------------------------------
    @staticmethod
    def scale(q):
        return q  # se3...:90-92

    def space(self):
        return self._space

    @staticmethod
    def maxSteering():
        return math.inf  # se3...:98-100

    def sampleGoal(self, rng=None):
        return self.goal_.copy()  # se3...:102-105

    def randomSample(self, rng):
        """se3...:107-113 with randomize.hpp:12-38; rng is a numpy Generator."""
        a, b, c = rng.uniform(0, 1), rng.uniform(0, 2 * math.pi), rng.uniform(0, 2 * math.pi)
        q = np.empty(7)
        q[0] = math.sqrt(1 - a) * math.sin(b)
        q[1] = math.sqrt(1 - a) * math.cos(b)
        q[2] = math.sqrt(a) * math.sin(c)
        q[3] = math.sqrt(a) * math.cos(c)
        for i in range(3):
            q[4 + i] = rng.uniform(self.min_[i], self.max_[i])
        return q

    def isGoal(self, q):
        return self._space.distance(self.goal_, q) <= self.goalRadius_  # se3...:115-117

    def sample_batch(self, rng, n):
        """n independent randomSample draws, vectorised (host side, feeds the batched entry points)."""
        from .planner import se3_random_batch
        return se3_random_batch(rng, n, self.min_, self.max_)

    def isValid(self, q, to=None):
        """isValid(q) (se3...:119-126) or isValid(from, to) (se3...:134-178)."""
        if to is None:
            return bool(self.check_states(q)[0])
        return bool(self.check_edges(q, to)[0])

    # ---- batched entry points (what the GPU path adds) ---------------------------------------
    def check_states(self, states):
        s = _f64(states, 7)
        out = np.empty(len(s), dtype=np.uint8)
        _capi.check(_capi.lib().mplb_check_states(self._h, _ptr(s), len(s), _ptr(out)))
        return out.astype(bool)

    def check_edges(self, frm, to, return_states_checked=False):
        a, b = _f64(frm, 7), _f64(to, 7)
        if a.shape != b.shape:
            raise ValueError("from/to shape mismatch")
        out = np.empty(len(a), dtype=np.uint8)
        checked = C.c_uint64()
        # (states_checked is only asked for when wanted: scalar calls that ask for it do not share launches)
        _capi.check(_capi.lib().mplb_check_edges(self._h, _ptr(a), _ptr(b), len(a), _ptr(out),
                                                 C.byref(checked) if return_states_checked else None))
        if return_states_checked:
            return out.astype(bool), checked.value
        return out.astype(bool)

    def edge_steps(self, frm, to):
        a, b = _f64(frm, 7), _f64(to, 7)
        steps = np.empty(len(a), dtype=np.uint32)
        _capi.check(_capi.lib().mplb_edges_plan(self._h, _ptr(a), _ptr(b), len(a), _ptr(steps)))
        return steps

    # raw-pointer variants for device-resident batches (pointers from torch tensors' data_ptr())
    def check_states_ptr(self, host_states_ptr, n, host_valid_ptr):
        _capi.check(_capi.lib().mplb_check_states(self._h, host_states_ptr, n, host_valid_ptr))

    def check_states_device(self, d_states_ptr, n, d_valid_ptr, stream=0):
        """Enqueued on `stream` (a cudaStream_t as int), not synchronised.  May be issued on several streams at once;
        successive batches issued alternately on two streams overlap one launch's drain with the next one's start."""
        _capi.check(_capi.lib().mplb_check_states_device(self._h, d_states_ptr, n, d_valid_ptr, stream))

    def check_edges_device(self, d_from, d_to, d_steps, n, max_steps, d_valid, stream=0):
        _capi.check(_capi.lib().mplb_check_edges_device(self._h, d_from, d_to, d_steps, n, int(max_steps), d_valid,
                                                        stream))

    def sample_states_device(self, seed, first, n, d_states_ptr, stream=0):
        """randomSample for n states into device memory: draw first + i of the Philox stream keyed by `seed`
        (the stream of mplambda_b200.workload.se3_poses)."""
        _capi.check(_capi.lib().mplb_sample_states_device(self._h, self.min_.ctypes.data_as(_capi.dp),
                                                          self.max_.ctypes.data_as(_capi.dp), int(seed), int(first), int(n),
                                                          d_states_ptr, stream))

    def check_edges_indexed(self, d_from_pool, from_rows, from_idx, d_to_pool, to_rows, to_idx, n=None, from_div=1, to_div=1,
                            return_states_checked=False):
        """isValid(from, to) for edges named by index into device-resident state pools (device pointers, e.g.
        Nigh.keys_device()): edge e = from_pool[from_idx[e] or e // from_div] -> to_pool[to_idx[e] or e // to_div].
        from_idx / to_idx: int64 numpy arrays (copied in, 8 B per end point) or None."""
        def idx(a):
            if a is None:
                return None, None
            a = np.ascontiguousarray(a, dtype=np.int64)
            return a, _ptr(a)
        fi, fp = idx(from_idx)
        ti, tp = idx(to_idx)
        if n is None:
            n = len(fi) if fi is not None else len(ti)
        out = np.empty(n, dtype=np.uint8)
        checked = C.c_uint64()
        _capi.check(_capi.lib().mplb_check_edges_indexed(self._h, d_from_pool, from_rows, fp, from_div, d_to_pool, to_rows,
                                                         tp, to_div, n, 0, _ptr(out), C.byref(checked)))
        if return_states_checked:
            return out.astype(bool), checked.value
        return out.astype(bool)

    def check_edges_indexed_device(self, d_from_pool, from_rows, d_from_idx, d_to_pool, to_rows, d_to_idx, n, d_valid,
                                   from_div=1, to_div=1, stream=0, synchronous=False):
        """the same with index arrays and verdicts in device memory; synchronous=True resolves ambiguous step counts
        like the host call, otherwise the work is only enqueued on `stream`."""
        L = _capi.lib()
        if synchronous:
            _capi.check(L.mplb_check_edges_indexed(self._h, d_from_pool, from_rows, d_from_idx, from_div, d_to_pool, to_rows,
                                                   d_to_idx, to_div, n, 1, d_valid, None))
        else:
            _capi.check(L.mplb_check_edges_indexed_device(self._h, d_from_pool, from_rows, d_from_idx, from_div, d_to_pool,
                                                          to_rows, d_to_idx, to_div, n, d_valid, stream))

    def track_margins(self, on):
        """Diagnostic: record the smallest margin of the exact triangle-pair decisions (stats()["min_exact_margin"])."""
        _capi.check(_capi.lib().mplb_scene_track_margins(self._h, int(bool(on))))

    def set_coalescing(self, on):
        """Concurrent scalar isValid calls share launches (on by default; mplb_scene_set_coalescing)."""
        _capi.check(_capi.lib().mplb_scene_set_coalescing(self._h, int(bool(on))))

    def stats(self, reset=False):
        st = _capi.Stats()
        _capi.check(_capi.lib().mplb_scene_stats(self._h, C.byref(st), int(reset)))
        return st.as_dict()


class L1Space:
    """nigh::L1Space<S, 8> distance (fetch_scenario.hpp:19)."""

    def __init__(self, scenario):
        self._sc = scenario

    def distance(self, a, b):
        a, b = _f64(a, 8), _f64(b, 8)
        return _capi.lib().mplb_distance(self._sc._h, a.ctypes.data_as(_capi.dp), b.ctypes.data_as(_capi.dp))

    @staticmethod
    def dimensions():
        return 8


class FetchScenario:
    """Mirror of mpl::demo::FetchScenario<S> (reference include/mpl/demo/fetch_scenario.hpp:11-177) for
    the validity part of the Scenario concept.  State = 8 joint values (fetch_robot.hpp:80-99).
    Goal sampling / isGoal go through the IK solver in the reference (fetch_scenario.hpp:84-103): out of
    scope here."""

    multiGoal = True
    _SCALE = np.array([10.0, 4, 2, 1, 1, 1, 1, 1])   # fetch_scenario.hpp:42-45

    def __init__(self, envFrame, envMesh, goal=None, goalTol=None, checkResolution=0.01, device=0):
        L = _capi.lib()
        self._h = C.c_void_p()
        frame = np.ascontiguousarray(np.asarray(envFrame, dtype=np.float64).reshape(3, 4))
        self.envFrame_ = frame
        self.goal_, self.goalTol_ = goal, goalTol
        self.invStepSize_ = 1.0 / checkResolution
        if isinstance(envMesh, (str, bytes)) or hasattr(envMesh, "__fspath__"):
            rc = L.mplb_scene_create_fetch_file(str(envMesh).encode(), frame.ctypes.data_as(_capi.dp), checkResolution,
                                                device, C.byref(self._h))
        else:
            env = np.ascontiguousarray(envMesh, dtype=np.float64).reshape(-1, 3, 3)
            rc = L.mplb_scene_create_fetch(env.ctypes.data_as(_capi.dp), len(env), frame.ctypes.data_as(_capi.dp),
                                           checkResolution, device, C.byref(self._h))
        if rc in (_capi.ERR_IO, _capi.ERR_INVALID):
            raise ValueError(L.mplb_last_error().decode("utf-8", "replace"))
        _capi.check(rc)
        self._space = L1Space(self)

    close = SE3RigidBodyScenario.close
    __del__ = SE3RigidBodyScenario.__del__
    stats = SE3RigidBodyScenario.stats

    @classmethod
    def scale(cls, q):
        return np.asarray(q, dtype=np.float64) * cls._SCALE   # fetch_scenario.hpp:75-77

    def space(self):
        return self._space

    @staticmethod
    def maxSteering():
        return math.inf

    def randomSample(self, rng):
        """FetchRobot::randomConfig (fetch_robot.hpp:163-175)."""
        from .workload import FETCH_JOINT_MAX, FETCH_JOINT_MIN
        lo = np.maximum(FETCH_JOINT_MIN, -4 * math.pi)
        hi = np.minimum(FETCH_JOINT_MAX, 4 * math.pi)
        return np.array([rng.uniform(lo[i], hi[i]) for i in range(8)])

    def isValid(self, q, to=None):
        if to is None:
            return bool(self.check_states(q)[0])
        return bool(self.check_edges(q, to)[0])

    def check_states(self, states):
        s = _f64(states, 8)
        out = np.empty(len(s), dtype=np.uint8)
        _capi.check(_capi.lib().mplb_check_states(self._h, _ptr(s), len(s), _ptr(out)))
        return out.astype(bool)

    def check_edges(self, frm, to, return_states_checked=False):
        a, b = _f64(frm, 8), _f64(to, 8)
        if a.shape != b.shape:
            raise ValueError("from/to shape mismatch")
        out = np.empty(len(a), dtype=np.uint8)
        checked = C.c_uint64()
        # (states_checked is only asked for when wanted: scalar calls that ask for it do not share launches)
        _capi.check(_capi.lib().mplb_check_edges(self._h, _ptr(a), _ptr(b), len(a), _ptr(out),
                                                 C.byref(checked) if return_states_checked else None))
        if return_states_checked:
            return out.astype(bool), checked.value
        return out.astype(bool)

    def edge_steps(self, frm, to):
        a, b = _f64(frm, 8), _f64(to, 8)
        steps = np.empty(len(a), dtype=np.uint32)
        _capi.check(_capi.lib().mplb_edges_plan(self._h, _ptr(a), _ptr(b), len(a), _ptr(steps)))
        return steps
