// C++ parity test in the style of the reference's own tests (test/test.hpp: EXPECT_THAT(expr) == value,
// executables run by the build): the drop-in scenario classes (include/mplb/*.hpp -> libmplb.so, GPU) against
// the FP64 oracle (oracle/mplb_oracle.h, linked here as the checker) on the same seeded inputs.
//   argv: <env.bin> <robot.bin> <states.bin> <edges_from.bin> <edges_to.bin> <autolab.bin> <fetch_states.bin>
// every file is raw little-endian float64.
#include <mplb/fetch_scenario.hpp>
#include <mplb/se3_rigid_body_scenario.hpp>

#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

extern "C" {
#include "mplb_oracle.h"
}

struct Failed : std::runtime_error { using std::runtime_error::runtime_error; };
#define EXPECT_THAT(expr)                                                                      \
    do {                                                                                       \
        if (!(expr)) {                                                                         \
            std::ostringstream os;                                                             \
            os << __FILE__ << ":" << __LINE__ << ": expected " #expr;                          \
            throw Failed(os.str());                                                            \
        }                                                                                      \
    } while (0)

static std::vector<double> slurp(const char *path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Failed(std::string("cannot read ") + path);
    f.seekg(0, std::ios::end);
    std::vector<double> v(f.tellg() / sizeof(double));
    f.seekg(0);
    f.read(reinterpret_cast<char *>(v.data()), v.size() * sizeof(double));
    return v;
}

using SE3 = mplb::demo::SE3RigidBodyScenario<double>;
using Fetch = mplb::demo::FetchScenario<double>;

static SE3::State se3State(const double *p) {
    SE3::State q;
#ifdef MPLB_HAVE_EIGEN
    for (int i = 0; i < 4; ++i) std::get<0>(q).coeffs()[i] = p[i];
    for (int i = 0; i < 3; ++i) std::get<1>(q)[i] = p[4 + i];
#else
    for (int i = 0; i < 7; ++i) q[i] = p[i];
#endif
    return q;
}

static void test_se3(char **argv) {
    auto env = slurp(argv[1]), robot = slurp(argv[2]), states = slurp(argv[3]), ef = slurp(argv[4]), et = slurp(argv[5]);
    double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    SE3 s(env.data(), env.size() / 9, robot.data(), robot.size() / 9, se3State(states.data()), lo, hi, 0.1);
    orc_mesh *oe = orc_mesh_create(env.data(), env.size() / 9), *orob = orc_mesh_create(robot.data(), robot.size() / 9);
    const std::size_t n = states.size() / 7;
    std::vector<SE3::State> qs;
    for (std::size_t i = 0; i < n; ++i) qs.push_back(se3State(&states[7 * i]));
    auto got = s.isValidBatch(qs);
    std::vector<std::uint8_t> want(n);
    orc_se3_check_states(orob, oe, states.data(), n, want.data(), 1, 0, nullptr);
    std::size_t nvalid = 0;
    for (std::size_t i = 0; i < n; ++i) {
        EXPECT_THAT(got[i] == want[i]);
        nvalid += want[i];
    }
    EXPECT_THAT(nvalid > 0 && nvalid < n);
    for (std::size_t i = 0; i < 16; ++i) EXPECT_THAT(s.isValid(qs[i]) == (want[i] != 0));   // the reference's single call
    const std::size_t m = ef.size() / 7;
    std::vector<SE3::State> from, to;
    for (std::size_t i = 0; i < m; ++i) {
        from.push_back(se3State(&ef[7 * i]));
        to.push_back(se3State(&et[7 * i]));
    }
    auto gotE = s.isValidBatch(from, to);
    std::vector<std::uint8_t> wantE(m);
    orc_se3_check_edges(orob, oe, ef.data(), et.data(), m, 50, 1, 10.0, wantE.data(), 1, 0, nullptr);
    for (std::size_t i = 0; i < m; ++i) EXPECT_THAT(gotE[i] == wantE[i]);
    for (std::size_t i = 0; i < 8 && i < m; ++i) EXPECT_THAT(s.isValid(from[i], to[i]) == (wantE[i] != 0));
    EXPECT_THAT(std::abs(s.space().distance(from[0], to[0]) - orc_se3_distance(&ef[0], &et[0], 50, 1)) == 0);
    std::printf("se3: %zu states (%zu valid), %zu edges: identical verdicts, isValid calls %u\n", n, nvalid, m, s.calls());
    orc_mesh_destroy(oe);
    orc_mesh_destroy(orob);
}

static void test_fetch(char **argv) {
    auto env = slurp(argv[6]), states = slurp(argv[7]);
    // scripts/fetch_task_001.sh:33 --env-frame=0.57,-0.90,0.00,0,0,-PI/2 as the reference parses it (float-truncated)
    const double c = -4.371139000186241e-08, sn = -1.0;
    const double frame[12] = {c, -sn, 0, (double)0.57f, sn, c, 0, (double)-0.90f, 0, 0, 1, 0};
    Fetch::Frame f;
#ifdef MPLB_FETCH_HAVE_EIGEN
    f.setIdentity();
    for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) f.linear()(r, cc) = frame[4 * r + cc]; f.translation()[r] = frame[4 * r + 3]; }
#else
    for (int i = 0; i < 12; ++i) f[i] = frame[i];
#endif
    Fetch s(f, env.data(), env.size() / 9, 0.01);
    orc_mesh *oe = orc_mesh_create(env.data(), env.size() / 9);
    const std::size_t n = states.size() / 8;
    std::vector<Fetch::State> qs(n);
    for (std::size_t i = 0; i < n; ++i)
        for (int k = 0; k < 8; ++k) qs[i][k] = states[8 * i + k];
    auto got = s.isValidBatch(qs);
    std::vector<std::uint8_t> want(n);
    orc_fetch_check_states(oe, frame, states.data(), n, want.data(), 0, nullptr);
    std::size_t nvalid = 0;
    for (std::size_t i = 0; i < n; ++i) {
        EXPECT_THAT(got[i] == want[i]);
        nvalid += want[i];
    }
    EXPECT_THAT(nvalid > 0 && nvalid < n);
    for (std::size_t i = 0; i < 8; ++i) EXPECT_THAT(s.isValid(qs[i]) == (want[i] != 0));
    std::printf("fetch: %zu configurations (%zu valid): identical verdicts\n", n, nvalid);
    orc_mesh_destroy(oe);
}

int main(int argc, char **argv) {
    if (argc < 8) {
        std::puts("usage: parity_test env.bin robot.bin states.bin from.bin to.bin autolab.bin fetch_states.bin");
        return 2;
    }
    try {
        test_se3(argv);
        test_fetch(argv);
    } catch (const std::exception &e) {
        std::cout << "FAILED: " << e.what() << std::endl;
        return 1;
    }
    std::puts("PASSED");
    return 0;
}
