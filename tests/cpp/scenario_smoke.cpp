// Compiles the C++ drop-in header against libmplb.so.  Without a GPU the constructor must fail
// loudly (no CPU fallback); with one, argv[1..2] = env/robot .dae and a few calls are exercised.
#include <mplb/se3_rigid_body_scenario.hpp>
#include <cstdio>
#include <random>

using Scenario = mplb::demo::SE3RigidBodyScenario<double>;

int main(int argc, char **argv) {
    Scenario::State goal{};
    double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    try {
        Scenario bad("/no/such/env.dae", "/no/such/robot.dae", goal, lo, hi, 0.1);
        std::puts("FAIL: missing mesh accepted");
        return 1;
    } catch (const std::invalid_argument &e) {
        std::printf("invalid_argument: %s\n", e.what());
    }
    if (argc < 3) return 0;
    try {
        Scenario s(argv[1], argv[2], goal, lo, hi, 0.1);
        std::mt19937_64 rng(1);
        auto q = s.randomSample(rng);
        std::printf("isValid(q)=%d isValid(q,q)=%d calls=%u\n", int(s.isValid(q)), int(s.isValid(q, q)), s.calls());
    } catch (const std::runtime_error &e) {
        std::printf("runtime_error: %s\n", e.what());
        return mplb_device_count() > 0 ? 1 : 0;
    }
    return 0;
}
