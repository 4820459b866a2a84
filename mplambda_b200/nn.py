"""Host-side mirror of the nearest-neighbour structure the planners use
(`unc::robotics::nigh::Nigh<Node*, Space, NodeKey, Concurrent, Strategy>`; reference
include/mpl/prrt.hpp:42,61,67,115 and include/mpl/pcforest.hpp:51,81,86,107,193) over the C ABI.
Values are insertion indices (the planner maps them back to its Node*)."""
import ctypes as C

import numpy as np

from . import _capi

SE3, L1 = 0, 1


class Nigh:
    def __init__(self, metric=SE3, dim=7, so3weight=50.0, l2weight=1.0, device=0):
        self._h = C.c_void_p()
        self.dim = 7 if metric == SE3 else dim
        _capi.check(_capi.lib().mplb_nn_create(metric, dim, so3weight, l2weight, device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _capi.lib().mplb_nn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _keys(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        if a.ndim != 2 or a.shape[1] != self.dim:
            raise ValueError("expected [n,%d] keys, got %s" % (self.dim, a.shape))
        return a

    def insert(self, keys):
        """nn_.insert(node) (prrt.hpp:61); returns the index of the first inserted key."""
        k = self._keys(keys)
        first = self.size()
        _capi.check(_capi.lib().mplb_nn_insert(self._h, k.ctypes.data_as(C.c_void_p), len(k)))
        return first

    def size(self):
        return int(_capi.lib().mplb_nn_size(self._h))

    def nearest(self, queries):
        """nn_.nearest(key) (prrt.hpp:67) batched -> (index [nq] int64, distance [nq]); -1/inf when empty."""
        q = self._keys(queries)
        idx = np.empty(len(q), dtype=np.int64)
        dist = np.empty(len(q), dtype=np.float64)
        _capi.check(_capi.lib().mplb_nn_nearest(self._h, q.ctypes.data_as(C.c_void_p), len(q),
                                                idx.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p)))
        return idx, dist

    def knearest(self, queries, k):
        """nn_.nearest(nbh, key, k) (pcforest.hpp:84-87) batched -> ([nq,k] indices, [nq,k] distances),
        ascending by (distance, index), padded with -1 / inf."""
        q = self._keys(queries)
        idx = np.empty((len(q), k), dtype=np.int64)
        dist = np.empty((len(q), k), dtype=np.float64)
        _capi.check(_capi.lib().mplb_nn_knearest(self._h, q.ctypes.data_as(C.c_void_p), len(q), int(k),
                                                 idx.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p)))
        return idx, dist

    def nearest_device(self, d_queries_ptr, nq, d_idx_ptr, d_dist_ptr, stream=0):
        _capi.check(_capi.lib().mplb_nn_nearest_device(self._h, d_queries_ptr, nq, d_idx_ptr, d_dist_ptr, stream))

    def knearest_device(self, d_queries_ptr, nq, k, d_idx_ptr, d_dist_ptr, stream=0):
        """k-NN on device-resident queries: [nq, k] int64 / float64 device arrays, enqueued on `stream`."""
        _capi.check(_capi.lib().mplb_nn_knearest_device(self._h, d_queries_ptr, nq, int(k), d_idx_ptr, d_dist_ptr, stream))

    def insert_device(self, d_keys_ptr, n, stream=0):
        """keys that already sit in device memory; returns the index of the first inserted key."""
        first = self.size()
        _capi.check(_capi.lib().mplb_nn_insert_device(self._h, d_keys_ptr, n, stream))
        return first

    def reserve(self, capacity):
        _capi.check(_capi.lib().mplb_nn_reserve(self._h, int(capacity)))

    def capacity(self):
        return int(_capi.lib().mplb_nn_capacity(self._h))

    def keys_device(self):
        """device pointer of the key array [size, dim] float64 (the pool index-named edges read tree nodes from)."""
        return _capi.lib().mplb_nn_keys_device(self._h) or 0

    def set_index(self, on):
        """Spatial index on / off (mplb_nn_set_index; on by default from 16 384 SE(3) keys)."""
        _capi.check(_capi.lib().mplb_nn_set_index(self._h, int(bool(on))))

    def index_info(self):
        """{indexed keys, cells, FP64 evaluations of indexed searches, queries that took the slow exact selection}"""
        a, b = C.c_size_t(), C.c_size_t()
        e, s = C.c_uint64(), C.c_uint64()
        _capi.check(_capi.lib().mplb_nn_index_info(self._h, C.byref(a), C.byref(b), C.byref(e), C.byref(s)))
        return {"indexed": a.value, "cells": b.value, "exact_evaluations": e.value, "slow_queries": s.value}
