"""ctypes binding of libmplb.so -- the same C ABI (include/mplb.h) a cgo/JNI/C++ caller binds.

The library is built in-tree by mplambda_b200/build.py (nvcc, sm_100a).  There is no Python or
CPU fallback: if the shared object is missing, loading raises.
"""
import ctypes as C
import os

from . import build as _build

OK, ERR_INVALID, ERR_IO, ERR_CUDA, ERR_NOMEM = 0, -1, -2, -3, -4


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("checks", "bv_tests", "tri_tests", "exact_tests", "launches",
                                           "env_tris", "robot_tris", "env_nodes", "robot_nodes",
                                           "stream_stalls", "streaming_off", "ambiguous_steps", "coalesced_calls")] + \
               [("min_exact_margin", C.c_double)]

    def as_dict(self):
        return {n: (float if n == "min_exact_margin" else int)(getattr(self, n)) for n, _ in self._fields_}


class RrtStatus(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("samples", C.c_uint64), ("edges_checked", C.c_uint64), ("nodes", C.c_uint64),
                ("solution", C.c_int64), ("seconds", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class MplbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmplb error %d: %s" % (code, msg))
        self.code = code


_lib = None

dp = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)

_SIGNATURES = {
    "mplb_last_error": (C.c_char_p, []),
    "mplb_device_count": (C.c_int, []),
    "mplb_mesh_load_collada": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(dp), C.POINTER(C.c_size_t), dp]),
    "mplb_free": (None, [C.c_void_p]),
    "mplb_scene_create_se3": (C.c_int, [dp, C.c_size_t, dp, C.c_size_t, C.c_double, C.c_double, C.c_double,
                                        C.c_int, C.POINTER(C.c_void_p)]),
    "mplb_scene_create_se3_files": (C.c_int, [C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_int,
                                              C.POINTER(C.c_void_p)]),
    "mplb_scene_create_fetch": (C.c_int, [dp, C.c_size_t, dp, C.c_double, C.c_int, C.POINTER(C.c_void_p)]),
    "mplb_scene_create_fetch_file": (C.c_int, [C.c_char_p, dp, C.c_double, C.c_int, C.POINTER(C.c_void_p)]),
    "mplb_scene_destroy": (None, [C.c_void_p]),
    "mplb_scene_state_dim": (C.c_int, [C.c_void_p]),
    "mplb_check_states": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mplb_check_edges": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, u64p]),
    "mplb_check_states_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "mplb_edges_plan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mplb_check_edges_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                          C.c_void_p, C.c_void_p]),
    "mplb_distance": (C.c_double, [C.c_void_p, dp, dp]),
    "mplb_scene_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats), C.c_int]),
    "mplb_scene_set_coalescing": (C.c_int, [C.c_void_p, C.c_int]),
    "mplb_scene_track_margins": (C.c_int, [C.c_void_p, C.c_int]),
    "mplb_nn_create": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_void_p)]),
    "mplb_nn_destroy": (None, [C.c_void_p]),
    "mplb_nn_insert": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mplb_nn_set_index": (C.c_int, [C.c_void_p, C.c_int]),
    "mplb_nn_index_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), u64p, u64p]),
    "mplb_nn_size": (C.c_size_t, [C.c_void_p]),
    "mplb_nn_nearest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "mplb_nn_knearest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "mplb_nn_nearest_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mplb_nn_knearest_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mplb_nn_insert_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mplb_nn_reserve": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mplb_nn_capacity": (C.c_size_t, [C.c_void_p]),
    "mplb_nn_keys_device": (C.c_void_p, [C.c_void_p]),
    "mplb_sample_states_device": (C.c_int, [C.c_void_p, dp, dp, C.c_uint64, C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p]),
    "mplb_check_edges_indexed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.c_uint32, C.c_size_t, C.c_int, C.c_void_p, u64p]),
    "mplb_check_edges_indexed_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p,
                                                  C.c_size_t, C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p, C.c_void_p]),
    "mplb_rrt_create": (C.c_int, [C.c_void_p, C.c_void_p, dp, dp, C.c_double, dp, dp, C.c_uint64, C.c_size_t, C.c_double,
                                  C.POINTER(C.c_void_p)]),
    "mplb_set_cache_dir": (C.c_int, [C.c_char_p]),
    "mplb_cache_stats": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "mplb_rrt_destroy": (None, [C.c_void_p]),
    "mplb_rrt_run": (C.c_int, [C.c_void_p, C.c_uint64, C.c_double, C.POINTER(RrtStatus)]),
    "mplb_rrt_tree": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "mplb_rrt_samples": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
}


def lib_path():
    return _build.LIB


def lib():
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError("libmplb.so is not built (%s); run `python -m mplambda_b200.build` -- "
                              "there is no fallback implementation" % path)
        L = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise MplbError(rc, lib().mplb_last_error().decode("utf-8", "replace"))
