// mplb/batched_rrt.hpp -- a batch-issuing RRT over the drop-in scenario classes.
//
// The reference's Planner<Scenario, PRRT> (include/mpl/prrt.hpp) issues one isValid per call from each OpenMP
// thread; the kernels behind include/mplb.h want thousands of states per launch.  This is the same algorithm
// with the loop body of Thread::addSample (prrt.hpp:280-299) applied to `batch` samples at once:
//
//     sample (goal with probability 1 %, prrt.hpp:301-315)  ->  nearest(scale(q))          (prrt.hpp:66-68)
//     d == 0 ? skip : isValid(nearest, q)   [isValid(from,to) starts with isValid(to), se3...:136]
//     add the survivors to the tree, solved as soon as a goal node is added                 (prrt.hpp:60-64)
//
// which is also what the reference's threads do concurrently against a shared tree (prrt.hpp:135-153; the
// steering distance of both scenarios is +inf, so there is no steering step).  Nearest neighbours come from an
// mplb_nn (exact, ties to the lowest index) instead of nigh.  Scenario needs: State, isValidBatch(from, to),
// isValid(q), randomSample(rng), sampleGoal(rng), isGoal(q), and the key helpers kKeyDim / key() / nnCreate().
#ifndef MPLB_BATCHED_RRT_HPP
#define MPLB_BATCHED_RRT_HPP

#include <chrono>
#include <cmath>
#include <cstdint>
#include <limits>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../mplb.h"

namespace mplb {

template <class Scenario, class RNG = std::mt19937_64>
class BatchedRRT {
public:
    using State = typename Scenario::State;

    struct Stats {
        std::size_t iterations = 0, samples = 0, edgesChecked = 0, nodes = 0;
        double seconds = 0;
    };

    // The scenario is constructed in place from `args`, like Planner<Scenario, Algorithm>(args...) (planner.hpp)
    template <class... Args>
    explicit BatchedRRT(std::size_t batch, std::uint64_t seed, int device, Args &&...args)
        : scenario_(std::forward<Args>(args)...), batch_(batch ? batch : 1), rng_(seed) {
        if (Scenario::nnCreate(device, &nn_) != MPLB_OK) throw std::runtime_error(mplb_last_error());
    }
    BatchedRRT(const BatchedRRT &) = delete;
    BatchedRRT &operator=(const BatchedRRT &) = delete;
    ~BatchedRRT() { mplb_nn_destroy(nn_); }

    Scenario &scenario() { return scenario_; }
    const Stats &stats() const { return stats_; }
    std::size_t size() const { return nodes_.size(); }
    bool isSolved() const { return solution_ >= 0; }

    // prrt.hpp:107-113: an invalid start is an error
    void addStart(const State &q) {
        if (!scenario_.isValid(q)) throw std::invalid_argument("start state is not valid");
        insert(std::vector<State>{q}, std::vector<std::int64_t>{-1});
    }

    // one batch: `batch` samples -> nearest -> edge checks -> insert
    void step() {
        std::vector<State> q(batch_);
        std::vector<std::uint8_t> isGoalSample(batch_, 0);
        std::uniform_real_distribution<double> u01(0, 1);
        for (std::size_t i = 0; i < batch_; ++i) {
            if (u01(rng_) < 0.01) {   // prrt.hpp:306-312
                if (auto g = scenario_.sampleGoal(rng_)) {
                    q[i] = *g;
                    isGoalSample[i] = 1;
                    continue;
                }
            }
            q[i] = scenario_.randomSample(rng_);
        }
        std::vector<double> keys(batch_ * Scenario::kKeyDim), dist(batch_);
        std::vector<std::int64_t> near(batch_);
        for (std::size_t i = 0; i < batch_; ++i) Scenario::key(q[i], &keys[i * Scenario::kKeyDim]);
        if (mplb_nn_nearest(nn_, keys.data(), batch_, near.data(), dist.data()) != MPLB_OK)
            throw std::runtime_error(mplb_last_error());
        stats_.samples += batch_;
        std::vector<State> from, to;
        std::vector<std::int64_t> parent;
        std::vector<std::uint8_t> goalFlag;
        for (std::size_t i = 0; i < batch_; ++i) {
            if (near[i] < 0 || !(dist[i] > 0)) continue;   // prrt.hpp:283-284
            from.push_back(nodes_[(std::size_t)near[i]]);
            to.push_back(q[i]);
            parent.push_back(near[i]);
            goalFlag.push_back(isGoalSample[i]);
        }
        if (to.empty()) return;
        const std::vector<std::uint8_t> ok = scenario_.isValidBatch(from, to);
        stats_.edgesChecked += to.size();
        std::vector<State> add;
        std::vector<std::int64_t> addParent;
        std::vector<std::uint8_t> addGoal;
        for (std::size_t i = 0; i < to.size(); ++i)
            if (ok[i]) {
                add.push_back(to[i]);
                addParent.push_back(parent[i]);
                addGoal.push_back(goalFlag[i]);
            }
        if (add.empty()) return;
        const std::size_t first = nodes_.size();
        insert(add, addParent);
        for (std::size_t i = 0; i < add.size() && solution_ < 0; ++i)
            if (addGoal[i] || scenario_.isGoal(add[i])) solution_ = (std::int64_t)(first + i);
    }

    // Planner::solve(doneFn) with a time limit (planner.hpp / prrt.hpp:135-153) -> solved?
    template <class Rep, class Period>
    bool solveFor(std::chrono::duration<Rep, Period> limit) {
        const auto t0 = std::chrono::steady_clock::now();
        while (!isSolved() && std::chrono::steady_clock::now() - t0 < limit) {
            step();
            ++stats_.iterations;
        }
        stats_.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats_.nodes = nodes_.size();
        return isSolved();
    }

    // start ... goal (prrt.hpp:217-229 walks goal -> start; reversed here)
    std::vector<State> solution() const {
        std::vector<State> path;
        for (std::int64_t i = solution_; i >= 0; i = parent_[(std::size_t)i]) path.push_back(nodes_[(std::size_t)i]);
        return std::vector<State>(path.rbegin(), path.rend());
    }

private:
    void insert(const std::vector<State> &qs, const std::vector<std::int64_t> &parents) {
        std::vector<double> keys(qs.size() * Scenario::kKeyDim);
        for (std::size_t i = 0; i < qs.size(); ++i) Scenario::key(qs[i], &keys[i * Scenario::kKeyDim]);
        if (mplb_nn_insert(nn_, keys.data(), qs.size()) < 0) throw std::runtime_error(mplb_last_error());
        nodes_.insert(nodes_.end(), qs.begin(), qs.end());
        parent_.insert(parent_.end(), parents.begin(), parents.end());
    }

    Scenario scenario_;
    mplb_nn *nn_ = nullptr;
    std::size_t batch_;
    RNG rng_;
    std::vector<State> nodes_;
    std::vector<std::int64_t> parent_;
    std::int64_t solution_ = -1;
    Stats stats_;
};

// The asymptotically optimal loop of the reference's second planner, Planner<Scenario, PCForest>
// (include/mpl/pcforest.hpp:487-567, one C-FOREST shard), applied to `batch` samples at once:
//
//     sample -> nearest -> isValid(nearest, q)                              (addSample, pcforest.hpp:487-506)
//     k nearest tree nodes, k = ceil(e (1 + 1/d) ln(n + 1))                 (pcforest.hpp:80-87,138)
//     parent = the neighbour with the cheapest valid connection             (pcforest.hpp:512-540)
//     rewire every neighbour that gets cheaper through the new node         (pcforest.hpp:548-566)
//
// The reference collects parent candidates walking the neighbours in ascending distance up to the first one that is not
// cheaper than the nearest node, tries them in increasing cost and stops at the first valid motion; here the same
// candidates of all new nodes are ONE isValidBatch call (all rewiring motions another) and the cheapest valid one wins
// -- the same choice.  Samples of one batch do not see each other, like the reference's concurrent threads.
// As in the reference an edge is immutable (node, parent edge, path cost): rewiring gives a node a new edge, its
// descendants keep the old chain, so every stored cost is the length of a path that exists.
// (mplambda_b200/planner.py BatchedRRTStar is the same loop; tests/test_planner.py checks it against the CPU oracle.)
template <class Scenario, class RNG = std::mt19937_64>
class BatchedRRTStar {
public:
    using State = typename Scenario::State;

    struct Stats {
        std::size_t iterations = 0, samples = 0, edgesChecked = 0, nodes = 0, rewired = 0;
        double seconds = 0;
    };

    template <class... Args>
    explicit BatchedRRTStar(std::size_t batch, std::uint64_t seed, int device, int dimension, Args &&...args)
        : scenario_(std::forward<Args>(args)...), batch_(batch ? batch : 1), dimension_(dimension), rng_(seed) {
        if (Scenario::nnCreate(device, &nn_) != MPLB_OK) throw std::runtime_error(mplb_last_error());
    }
    BatchedRRTStar(const BatchedRRTStar &) = delete;
    BatchedRRTStar &operator=(const BatchedRRTStar &) = delete;
    ~BatchedRRTStar() { mplb_nn_destroy(nn_); }

    Scenario &scenario() { return scenario_; }
    const Stats &stats() const { return stats_; }
    std::size_t size() const { return nodes_.size(); }
    bool isSolved() const { return bestEdge_ >= 0; }
    double bestCost() const { return bestCost_; }

    void addStart(const State &q) {
        if (!scenario_.isValid(q)) throw std::invalid_argument("start state is not valid");   // pcforest.hpp:476-477
        addNodes({q}, {-1}, {0.0}, {0});
    }

    void step() {
        const std::size_t K = Scenario::kKeyDim;
        std::vector<State> q(batch_);
        std::vector<std::uint8_t> known(batch_, 0);
        std::uniform_real_distribution<double> u01(0, 1);
        for (std::size_t i = 0; i < batch_; ++i) {
            if (u01(rng_) < 0.01) {
                if (auto g = scenario_.sampleGoal(rng_)) {
                    q[i] = *g;
                    known[i] = 1;
                    continue;
                }
            }
            q[i] = scenario_.randomSample(rng_);
        }
        std::vector<double> keys(batch_ * K), dist(batch_);
        std::vector<std::int64_t> near(batch_);
        for (std::size_t i = 0; i < batch_; ++i) Scenario::key(q[i], &keys[i * K]);
        check(mplb_nn_nearest(nn_, keys.data(), batch_, near.data(), dist.data()));
        stats_.samples += batch_;
        std::vector<State> from, to;
        std::vector<std::size_t> src;
        for (std::size_t i = 0; i < batch_; ++i)
            if (near[i] >= 0 && dist[i] > 0) {
                from.push_back(nodes_[(std::size_t)near[i]]);
                to.push_back(q[i]);
                src.push_back(i);
            }
        if (to.empty()) return;
        std::vector<std::uint8_t> ok = scenario_.isValidBatch(from, to);
        stats_.edgesChecked += to.size();
        std::vector<State> qn;
        std::vector<std::int64_t> parent;
        std::vector<double> parentCost, newKeys;
        std::vector<std::uint8_t> goal;
        for (std::size_t j = 0; j < to.size(); ++j)
            if (ok[j]) {
                const std::size_t i = src[j];
                qn.push_back(q[i]);
                parent.push_back(near[i]);
                parentCost.push_back(cost((std::size_t)near[i]) + dist[i]);
                goal.push_back(known[i] || scenario_.isGoal(q[i]));
                newKeys.insert(newKeys.end(), &keys[i * K], &keys[i * K] + K);
            }
        const std::size_t m = qn.size();
        if (m == 0) return;
        const int k = (int)std::ceil(std::exp(1.0) * (1.0 + 1.0 / dimension_) * std::log((double)nodes_.size() + 1.0));
        std::vector<std::int64_t> nbr(m * (std::size_t)k);
        std::vector<double> nd(m * (std::size_t)k);
        check(mplb_nn_knearest(nn_, newKeys.data(), m, k, nbr.data(), nd.data()));
        // ---- parent selection: the neighbours, in ascending distance, up to the first one that would not be cheaper
        //      than the nearest node (pcforest.hpp:519-526 stops there: `if (nbrPathCost >= parentCost) break`)
        std::vector<std::uint8_t> rejected(nbr.size(), 0);
        std::vector<std::size_t> at;
        from.clear();
        to.clear();
        for (std::size_t r = 0; r < m; ++r)
            for (int c = 0; c < k; ++c) {
                const std::size_t x = r * (std::size_t)k + c;
                if (!(nbr[x] >= 0 && cost((std::size_t)nbr[x]) + nd[x] < parentCost[r])) break;
                from.push_back(nodes_[(std::size_t)nbr[x]]);
                to.push_back(qn[r]);
                at.push_back(x);
            }
        if (!at.empty()) {
            ok = scenario_.isValidBatch(from, to);
            stats_.edgesChecked += at.size();
            for (std::size_t j = 0; j < at.size(); ++j) {
                const std::size_t x = at[j], r = x / (std::size_t)k;
                if (!ok[j]) {
                    rejected[x] = 1;   // pcforest.hpp:527,555-556
                } else if (cost((std::size_t)nbr[x]) + nd[x] < parentCost[r]) {
                    parentCost[r] = cost((std::size_t)nbr[x]) + nd[x];
                    parent[r] = nbr[x];
                }
            }
        }
        const std::size_t first = nodes_.size();
        addNodes(qn, parent, parentCost, goal);
        // ---- rewiring: neighbours that get cheaper through the new node
        at.clear();
        from.clear();
        to.clear();
        for (std::size_t r = 0; r < m; ++r)
            for (int c = 0; c < k; ++c) {
                const std::size_t x = r * (std::size_t)k + c;
                if (nbr[x] >= 0 && !rejected[x] && parentCost[r] + nd[x] < cost((std::size_t)nbr[x])) {
                    from.push_back(qn[r]);
                    to.push_back(nodes_[(std::size_t)nbr[x]]);
                    at.push_back(x);
                }
            }
        if (!at.empty()) {
            ok = scenario_.isValidBatch(from, to);
            stats_.edgesChecked += at.size();
            for (std::size_t j = 0; j < at.size(); ++j) {
                if (!ok[j]) continue;
                const std::size_t x = at[j], r = x / (std::size_t)k, node = (std::size_t)nbr[x];
                const double through = parentCost[r] + nd[x];
                if (through < cost(node)) {   // setEdge keeps the cheaper one, pcforest.hpp:569-581
                    nodeEdge_[node] = newEdge(node, nodeEdge_[first + r], through);
                    ++stats_.rewired;
                    if (isGoal_[node] && through < bestCost_) {
                        bestCost_ = through;
                        bestEdge_ = nodeEdge_[node];
                    }
                }
            }
        }
    }

    // anytime: runs for the whole limit (or `iterations` more steps) and keeps the cheapest path found
    template <class Rep, class Period>
    bool solveFor(std::chrono::duration<Rep, Period> limit, std::size_t iterations = ~(std::size_t)0, bool stopAtFirst = false) {
        const auto t0 = std::chrono::steady_clock::now();
        for (std::size_t it = 0; it < iterations && std::chrono::steady_clock::now() - t0 < limit; ++it) {
            if (stopAtFirst && isSolved()) break;
            step();
            ++stats_.iterations;
        }
        stats_.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        stats_.nodes = nodes_.size();
        return isSolved();
    }

    std::vector<State> solution() const {
        std::vector<State> path;
        for (std::int64_t e = bestEdge_; e >= 0; e = edgeParent_[(std::size_t)e]) path.push_back(nodes_[edgeNode_[(std::size_t)e]]);
        return std::vector<State>(path.rbegin(), path.rend());
    }

private:
    static void check(int rc) {
        if (rc != MPLB_OK) throw std::runtime_error(mplb_last_error());
    }
    double cost(std::size_t node) const { return edgeCost_[(std::size_t)nodeEdge_[node]]; }
    std::int64_t newEdge(std::size_t node, std::int64_t parentEdge, double c) {
        edgeNode_.push_back(node);
        edgeParent_.push_back(parentEdge);
        edgeCost_.push_back(c);
        return (std::int64_t)edgeCost_.size() - 1;
    }
    void addNodes(const std::vector<State> &qs, const std::vector<std::int64_t> &parents, const std::vector<double> &costs,
                  const std::vector<std::uint8_t> &goal) {
        std::vector<double> keys(qs.size() * Scenario::kKeyDim);
        for (std::size_t i = 0; i < qs.size(); ++i) Scenario::key(qs[i], &keys[i * Scenario::kKeyDim]);
        check(mplb_nn_insert(nn_, keys.data(), qs.size()));
        for (std::size_t i = 0; i < qs.size(); ++i) {
            const std::size_t node = nodes_.size();
            nodes_.push_back(qs[i]);
            isGoal_.push_back(goal[i]);
            nodeEdge_.push_back(newEdge(node, parents[i] < 0 ? -1 : nodeEdge_[(std::size_t)parents[i]], costs[i]));
            if (goal[i] && costs[i] < bestCost_) {   // updateSolution, pcforest.hpp:542-543
                bestCost_ = costs[i];
                bestEdge_ = nodeEdge_[node];
            }
        }
    }

    Scenario scenario_;
    mplb_nn *nn_ = nullptr;
    std::size_t batch_;
    int dimension_;
    RNG rng_;
    std::vector<State> nodes_;
    std::vector<std::uint8_t> isGoal_;
    std::vector<std::int64_t> nodeEdge_, edgeParent_;
    std::vector<std::size_t> edgeNode_;
    std::vector<double> edgeCost_;
    std::int64_t bestEdge_ = -1;
    double bestCost_ = std::numeric_limits<double>::infinity();
    Stats stats_;
};

}  // namespace mplb

#endif
