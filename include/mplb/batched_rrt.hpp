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
#include <cstdint>
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

}  // namespace mplb

#endif
