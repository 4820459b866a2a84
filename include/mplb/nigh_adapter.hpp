// nigh's nearest-neighbour interface, as the reference planners use it, over the mplb C ABI.
//
//   unc::robotics::nigh::Nigh<Node*, Space, NodeKey, Concurrent, NNStrategy> nn_;      prrt.hpp:42, pcforest.hpp:51
//   nn_.insert(node)                               prrt.hpp:61,  pcforest.hpp:107
//   nn_.nearest(Scenario::scale(q))  -> std::optional<std::pair<Node*, Distance>>      prrt.hpp:67,280  pcforest.hpp:81,488
//   nn_.nearest(nbh, Scenario::scale(q), k)   nbh: std::vector<std::tuple<Node*, Distance>>, ascending    pcforest.hpp:86
//   nn_.size()                                     prrt.hpp:115, pcforest.hpp:85,193
//
// mplb::NighAdapter<T, Space, KeyFn> offers exactly these members; the values live in a std::deque on the host, the keys
// (KeyFn()(value), flattened to doubles by Space::flatten) in an mplb_nn on the GPU.  Results equal nigh's on the same
// keys: exact metric, ties to the lowest insertion index (include/mplb.h).  insert and the queries may be called
// concurrently from the planner's OpenMP threads (the reference instantiates nigh with Concurrent): the value table is
// guarded by a shared mutex, the device structure serialises itself.
//
// include/mplb/shim/nigh/*.hpp declare unc::robotics::nigh::{SE3Space, L1Space, Concurrent, auto_strategy_t, Nigh} in
// terms of this adapter, so that the reference's prrt.hpp / pcforest.hpp compile UNCHANGED with
// `-I include/mplb/shim` in place of nigh's include directory.
#pragma once
#ifndef MPLB_NIGH_ADAPTER_HPP
#define MPLB_NIGH_ADAPTER_HPP

#include <cstddef>
#include <cstdint>
#include <deque>
#include <limits>
#include <mutex>
#include <optional>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../mplb.h"

namespace mplb {

// Space requirements: `using Type`, `using Distance`, static int metric() (0 SE3, 1 L1), static int keyDim(),
// static double so3Weight(), static double l2Weight(), static void flatten(const Type&, double*).
template <class T, class Space, class KeyFn>
class NighAdapter {
public:
    using Key = typename Space::Type;
    using Distance = typename Space::Distance;

private:
    mplb_nn *nn_ = nullptr;
    KeyFn keyFn_;
    std::deque<T> values_;            // insertion index -> value
    mutable std::shared_mutex mu_;

    [[noreturn]] static void raise(int rc) {
        const std::string msg = mplb_last_error();
        if (rc == MPLB_ERR_INVALID) throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
    }
    T valueAt(std::int64_t i) const {
        std::shared_lock<std::shared_mutex> g(mu_);
        return values_[(std::size_t)i];
    }

public:
    explicit NighAdapter(const Space & = Space(), const KeyFn &keyFn = KeyFn(), int device = 0) : keyFn_(keyFn) {
        const int rc = mplb_nn_create(Space::metric(), Space::keyDim(), Space::so3Weight(), Space::l2Weight(), device, &nn_);
        if (rc != MPLB_OK) raise(rc);
    }
    NighAdapter(const NighAdapter &) = delete;
    NighAdapter &operator=(const NighAdapter &) = delete;
    ~NighAdapter() { mplb_nn_destroy(nn_); }

    std::size_t size() const { return mplb_nn_size(nn_); }

    void insert(const T &value) {
        double k[16];
        Space::flatten(keyFn_(value), k);
        // index = position in the device structure: both are appended under the same lock
        std::unique_lock<std::shared_mutex> g(mu_);
        const int rc = mplb_nn_insert(nn_, k, 1);
        if (rc != MPLB_OK) raise(rc);
        values_.push_back(value);
    }

    std::optional<std::pair<T, Distance>> nearest(const Key &q) const {
        double k[16];
        Space::flatten(q, k);
        std::int64_t idx = -1;
        double dist = 0;
        const int rc = mplb_nn_nearest(nn_, k, 1, &idx, &dist);
        if (rc != MPLB_OK) raise(rc);
        if (idx < 0) return std::nullopt;
        return std::make_pair(valueAt(idx), static_cast<Distance>(dist));
    }

    // the k nearest within maxRadius, ascending by (distance, insertion index)
    template <class Tuple, class Alloc>
    void nearest(std::vector<Tuple, Alloc> &nbh, const Key &q, std::size_t k,
                 Distance maxRadius = std::numeric_limits<Distance>::infinity()) const {
        nbh.clear();
        k = std::min<std::size_t>(k, 64);   // (mplb_nn_knearest: k <= 64; PCForest's k = e(1 + 1/d) ln(n + 1) stays below 64 up to n = 5e8)
        if (k == 0) return;
        double key[16];
        Space::flatten(q, key);
        std::int64_t idx[64];
        double dist[64];
        const int rc = mplb_nn_knearest(nn_, key, 1, (int)k, idx, dist);
        if (rc != MPLB_OK) raise(rc);
        for (std::size_t i = 0; i < k && idx[i] >= 0; ++i) {
            if (!(static_cast<Distance>(dist[i]) <= maxRadius)) break;
            nbh.emplace_back(valueAt(idx[i]), static_cast<Distance>(dist[i]));
        }
    }

    mplb_nn *handle() const { return nn_; }
};

}  // namespace mplb

#endif
