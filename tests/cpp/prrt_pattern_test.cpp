// The reference planners' use of nigh, spelled as they spell it, against the drop-in scenario + the nigh stand-in
// (include/mplb/shim): compiled with `-I include -I include/mplb/shim`.
//
//   prrt.hpp:35-42      using Concurrency = nigh::Concurrent; using NNStrategy = nigh::auto_strategy_t<Space, Concurrency>;
//                       nigh::Nigh<Node*, Space, NodeKey, Concurrency, NNStrategy> nn_;
//   prrt.hpp:59-68      nn_.insert(n); nn_.nearest(Scenario::scale(q))
//   prrt.hpp:280-299    auto [nNear, d] = planner.nearest(qRand).value(); isValid(qRand); isValid(nNear->state(), qRand)
//   pcforest.hpp:84-87  k = ceil(kRRG * log(nn_.size() + 1)); nn_.nearest(nbh, Scenario::scale(q), k)
//
// This is a single-threaded RRT written with those statements (the reference's own prrt.hpp needs Eigen, which this
// image does not have; tests/test_cpp_header.py compiles the real header too wherever Eigen and the reference tree exist).
//   prrt_pattern_test env.bin robot.bin start[7] goal[7] min[3] max[3] [seed] [max samples]
// prints "solved <waypoints> nodes <n> knn_ok <0|1>" and the path.
#include <nigh/auto_strategy.hpp>
#include <mplb/se3_rigid_body_scenario.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <random>
#include <string>
#include <vector>

using Scenario = mplb::demo::SE3RigidBodyScenario<double>;
using Space = typename Scenario::Space;
using State = typename Scenario::State;
using Distance = typename Scenario::Distance;
static_assert(std::is_same<Space, unc::robotics::nigh::SE3Space<double, 50, 1>>::value,
              "Scenario::Space must be nigh's SE3Space (se3_rigid_body_scenario.hpp:18)");

class Node {
    const Node *parent_;
    State state_;
public:
    Node(const Node *parent, const State &q) : parent_(parent), state_(q) {}
    const Node *parent() const { return parent_; }
    const State &state() const { return state_; }
};
struct NodeKey {
    State operator()(const Node *node) const { return Scenario::scale(node->state()); }
};
using Concurrency = unc::robotics::nigh::Concurrent;
using NNStrategy = unc::robotics::nigh::auto_strategy_t<Space, Concurrency>;
using NN = unc::robotics::nigh::Nigh<Node *, Space, NodeKey, Concurrency, NNStrategy>;
using Neighborhood = std::vector<std::tuple<Node *, Distance>>;   // pcforest.hpp:45

static std::vector<double> slurp(const char *path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::invalid_argument(std::string("cannot open ") + path);
    std::vector<double> v((size_t)f.tellg() / sizeof(double));
    f.seekg(0);
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(v.size() * sizeof(double)));
    return v;
}

int main(int argc, char **argv) {
    if (argc < 23) {
        std::puts("usage: prrt_pattern_test env.bin robot.bin start[7] goal[7] min[3] max[3] [seed] [max samples]");
        return 2;
    }
    try {
        const std::vector<double> env = slurp(argv[1]), robot = slurp(argv[2]);
        double v[20];
        for (int i = 0; i < 20; ++i) v[i] = std::atof(argv[3 + i]);
        const State start = Scenario::fromKey(v), goal = Scenario::fromKey(v + 7);
        Scenario scenario(env.data(), env.size() / 9, robot.data(), robot.size() / 9, goal, v + 14, v + 17, 0.1);
        std::mt19937_64 rng(argc > 23 ? std::atol(argv[23]) : 1);
        const long maxSamples = argc > 24 ? std::atol(argv[24]) : 200000;
        NN nn_;
        std::deque<Node> nodes_;
        const Node *solution_ = nullptr;
        if (!scenario.isValid(start)) throw std::invalid_argument("start state is not valid");   // prrt.hpp:109-110
        nn_.insert(&nodes_.emplace_back(nullptr, start));
        std::uniform_real_distribution<double> unif01;
        bool knnOk = true;
        Neighborhood nbh;
        for (long s = 0; s < maxSamples && !solution_; ++s) {
            bool knownGoal = unif01(rng) < 0.01;
            State qRand = knownGoal ? *scenario.sampleGoal(rng) : scenario.randomSample(rng);
            auto [nNear, d] = nn_.nearest(Scenario::scale(qRand)).value();
            if (d == 0) continue;
            if (!scenario.isValid(qRand)) continue;
            if (!scenario.isValid(nNear->state(), qRand)) continue;
            Node *nNew = &nodes_.emplace_back(nNear, qRand);
            // pcforest.hpp:84-87: the neighbourhood of the new node; its first entry must be the nearest found above
            const unsigned k = (unsigned)std::ceil(2.718281828459045 * (1 + 1.0 / 6) * std::log(double(nn_.size() + 1)));
            nn_.nearest(nbh, Scenario::scale(qRand), k);
            knnOk = knnOk && !nbh.empty() && std::get<0>(nbh[0]) == nNear && std::get<1>(nbh[0]) == d &&
                    nbh.size() == std::min<std::size_t>(k, nn_.size());
            for (std::size_t i = 1; i < nbh.size(); ++i) knnOk = knnOk && std::get<1>(nbh[i - 1]) <= std::get<1>(nbh[i]);
            knnOk = knnOk && scenario.space().distance(Scenario::scale(nNear->state()), Scenario::scale(qRand)) == d;
            nn_.insert(nNew);
            if (knownGoal || scenario.isGoal(nNew->state())) solution_ = nNew;
        }
        if (!solution_) {
            std::printf("unsolved nodes %zu\n", nn_.size());
            return 1;
        }
        std::vector<const Node *> path;
        for (const Node *n = solution_; n; n = n->parent()) path.push_back(n);
        std::printf("solved %zu nodes %zu knn_ok %d\n", path.size(), nn_.size(), knnOk ? 1 : 0);
        for (auto it = path.rbegin(); it != path.rend(); ++it) {
            double k[7];
            Scenario::key((*it)->state(), k);
            std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", k[0], k[1], k[2], k[3], k[4], k[5], k[6]);
        }
        return 0;
    } catch (const std::exception &e) {
        std::printf("error: %s\n", e.what());
        return 3;
    }
}
