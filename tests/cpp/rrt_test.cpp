// The C++ batch-issuing RRT (include/mplb/batched_rrt.hpp) on a scene given as raw triangle files.
//   rrt_test env.bin robot.bin  s0..s6  g0..g6  min0..2  max0..2  [batch] [seed] [seconds] [star]
// env.bin / robot.bin: float64 [n][3][3] (robot already centred).  Prints "solved <n>" and the path, one
// state per line (%.17g), or "unsolved".  tests/test_cpp_header.py checks the path against the oracle.
#include <mplb/batched_rrt.hpp>
#include <mplb/se3_rigid_body_scenario.hpp>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

using Scenario = mplb::demo::SE3RigidBodyScenario<double>;

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
        std::puts("usage: rrt_test env.bin robot.bin start[7] goal[7] min[3] max[3] [batch] [seed] [seconds]");
        return 2;
    }
    try {
        const std::vector<double> env = slurp(argv[1]), robot = slurp(argv[2]);
        double v[20];
        for (int i = 0; i < 20; ++i) v[i] = std::atof(argv[3 + i]);
        const Scenario::State start = Scenario::fromKey(v), goal = Scenario::fromKey(v + 7);
        const std::size_t batch = argc > 23 ? (std::size_t)std::atol(argv[23]) : 2048;
        const std::uint64_t seed = argc > 24 ? (std::uint64_t)std::atol(argv[24]) : 1;
        const double seconds = argc > 25 ? std::atof(argv[25]) : 30.0;
        const bool star = argc > 26 && std::string(argv[26]) == "star";
        if (star) {   // PCForest's loop: first path, then a few improving rounds
            mplb::BatchedRRTStar<Scenario> rrt(batch, seed, 0, 6, env.data(), env.size() / 9, robot.data(), robot.size() / 9, goal,
                                               v + 14, v + 17, 0.1);
            rrt.addStart(start);
            if (!rrt.solveFor(std::chrono::duration<double>(seconds), ~(std::size_t)0, true)) {
                std::puts("unsolved");
                return 1;
            }
            const double first = rrt.bestCost();
            rrt.solveFor(std::chrono::duration<double>(seconds), 6);
            const auto path = rrt.solution();
            const auto &st = rrt.stats();
            std::printf("solved %zu  iterations %zu samples %zu edges %zu nodes %zu rewired %zu first %.17g best %.17g\n", path.size(),
                        st.iterations, st.samples, st.edgesChecked, st.nodes, st.rewired, first, rrt.bestCost());
            for (const auto &q : path) {
                double k[7];
                Scenario::key(q, k);
                std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", k[0], k[1], k[2], k[3], k[4], k[5], k[6]);
            }
            return 0;
        }
        mplb::BatchedRRT<Scenario> rrt(batch, seed, 0, env.data(), env.size() / 9, robot.data(), robot.size() / 9, goal, v + 14,
                                       v + 17, 0.1);
        rrt.addStart(start);
        if (!rrt.solveFor(std::chrono::duration<double>(seconds))) {
            std::puts("unsolved");
            return 1;
        }
        const auto path = rrt.solution();
        const auto &st = rrt.stats();
        std::printf("solved %zu  iterations %zu samples %zu edges %zu nodes %zu seconds %.4f\n", path.size(), st.iterations,
                    st.samples, st.edgesChecked, st.nodes, st.seconds);
        for (const auto &q : path) {
            double k[7];
            Scenario::key(q, k);
            std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", k[0], k[1], k[2], k[3], k[4], k[5], k[6]);
        }
        return 0;
    } catch (const std::exception &e) {
        std::printf("error: %s\n", e.what());
        return 3;
    }
}
