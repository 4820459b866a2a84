// Aggregate throughput of the reference's scalar calls issued by T host threads at once -- isValid(q) =
// mplb_check_states(n = 1), isValid(from, to) = mplb_check_edges(n = 1) -- the way PRRT's OpenMP threads call the
// scenario (reference include/mpl/prrt.hpp:140-152, 280-299).  Built and driven by tools/scalar_calls.py.
//   argv: env.bin robot.bin states.bin from.bin to.bin threads seconds coalesce(0|1) [want_states.bin want_edges.bin]
// (raw float64 files; the optional last two hold the expected verdicts, 0.0 / 1.0: every call is checked against them)
#include <mplb.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <thread>
#include <vector>

static std::vector<double> slurp(const char *path) {
    std::ifstream f(path, std::ios::binary);
    f.seekg(0, std::ios::end);
    std::vector<double> v(f.tellg() / sizeof(double));
    f.seekg(0);
    f.read(reinterpret_cast<char *>(v.data()), v.size() * sizeof(double));
    return v;
}

int main(int argc, char **argv) {
    if (argc < 9) {
        std::printf("usage: scalar_calls env robot states from to threads seconds coalesce\n");
        return 2;
    }
    auto env = slurp(argv[1]), robot = slurp(argv[2]), states = slurp(argv[3]), from = slurp(argv[4]), to = slurp(argv[5]);
    const int threads = std::atoi(argv[6]);
    const double seconds = std::atof(argv[7]);
    const int coalesce = std::atoi(argv[8]);
    std::vector<double> wantS, wantE;
    if (argc >= 11) {
        wantS = slurp(argv[9]);
        wantE = slurp(argv[10]);
    }
    mplb_scene *sc = nullptr;
    if (mplb_scene_create_se3(env.data(), env.size() / 9, robot.data(), robot.size() / 9, 50, 1, 0.1, 0, &sc)) {
        std::printf("scene: %s\n", mplb_last_error());
        return 1;
    }
    mplb_scene_set_coalescing(sc, coalesce);
    const size_t ns = states.size() / 7, ne = from.size() / 7;
    for (int edges = 0; edges < 2; ++edges) {
        std::atomic<bool> stop{false};
        std::atomic<long long> calls{0}, invalid{0}, errors{0}, wrong{0};
        auto work = [&](int t) {
            long long mine = 0, bad = 0;
            uint8_t v = 0;
            for (size_t i = (size_t)t * 7919; !stop.load(std::memory_order_relaxed); ++i) {
                const int rc = edges ? mplb_check_edges(sc, &from[7 * (i % ne)], &to[7 * (i % ne)], 1, &v, nullptr)
                                     : mplb_check_states(sc, &states[7 * (i % ns)], 1, &v);
                if (rc) ++errors;
                const std::vector<double> &want = edges ? wantE : wantS;
                if (!want.empty() && (v != 0) != (want[i % (edges ? ne : ns)] != 0.0)) ++wrong;
                bad += !v;
                ++mine;
            }
            calls += mine;
            invalid += bad;
        };
        for (int warm = 0; warm < 2; ++warm) {   // the first round warms contexts, graphs and the combiner up
            stop = false;
            calls = 0; invalid = 0; wrong = 0;
            std::vector<std::thread> th;
            const auto t0 = std::chrono::steady_clock::now();
            for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
            std::this_thread::sleep_for(std::chrono::duration<double>(warm ? seconds : 0.2));
            stop = true;
            for (auto &x : th) x.join();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (warm) {
                mplb_stats_t st;
                mplb_scene_stats(sc, &st, 1);
                std::printf("%s threads %d coalesce %d: %.0f calls/s (%.2f us per call per thread), invalid %.3f, errors %lld, wrong %lld, launches per call %.3f, coalesced %llu\n",
                            edges ? "isValid(from,to)" : "isValid(q)", threads, coalesce, calls / dt, dt * threads / calls * 1e6,
                            (double)invalid / calls, (long long)errors, (long long)wrong, (double)st.launches / calls, (unsigned long long)st.coalesced_calls);
            } else {
                mplb_stats_t st;
                mplb_scene_stats(sc, &st, 1);
            }
        }
    }
    mplb_scene_destroy(sc);
    return 0;
}
