// O0: the reference's own SE3RigidBodyScenario -- real FCL, Eigen, nigh and assimp -- as the verdict source.
//
// TEST INFRASTRUCTURE (like everything under oracle/): nothing in the product links or runs this.  It exists for the
// one thing the in-repo restatement (oracle/mplb_oracle.c, "O1") cannot give: a two-sided pin of the oracle against
// FCL itself (SURVEY.md 8c "Consequence").  It cannot be built in the development container (no
// FCL / libccd / Eigen / assimp / nigh, no network); on a machine that has them,
//     python oracle/o0_fcl/build.py            -> oracle/_ref/o0_se3   (or a one-line reason why not)
//     python -m pytest tests/test_o0_fcl.py    -> O0 against O1 on the seeded pose and edge streams
// The reference header is compiled where it lies (MPLB_REFERENCE_ROOT, default /root/reference): no reference source is
// copied into this repository.
//
//   o0_se3 states <env.dae> <robot.dae> <resolution> <in.f64> <out.u8>     in: [n][7] qx qy qz qw tx ty tz
//   o0_se3 edges  <env.dae> <robot.dae> <resolution> <in.f64> <out.u8>     in: [n][14] from | to
// out: one byte per state / edge, 1 = isValid.
#include <mpl/demo/se3_rigid_body_scenario.hpp>   // reference include/mpl/demo/se3_rigid_body_scenario.hpp:14-180

#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

using S = double;
using Scenario = mpl::demo::SE3RigidBodyScenario<S>;
using State = Scenario::State;

static State stateOf(const double *p) {
    State q;
    // Eigen's coefficient order is x, y, z, w (the layout of include/mplb.h's SE3 states)
    std::get<Eigen::Quaternion<S>>(q).coeffs() << p[0], p[1], p[2], p[3];
    std::get<Eigen::Matrix<S, 3, 1>>(q) << p[4], p[5], p[6];
    return q;
}

int main(int argc, char **argv) {
    if (argc != 7 || (std::strcmp(argv[1], "states") && std::strcmp(argv[1], "edges"))) {
        std::fprintf(stderr, "usage: o0_se3 states|edges env.dae robot.dae resolution in.f64 out.u8\n");
        return 2;
    }
    const bool edges = !std::strcmp(argv[1], "edges");
    std::ifstream in(argv[5], std::ios::binary);
    if (!in) {
        std::fprintf(stderr, "cannot read %s\n", argv[5]);
        return 1;
    }
    in.seekg(0, std::ios::end);
    std::vector<double> v(in.tellg() / sizeof(double));
    in.seekg(0);
    in.read(reinterpret_cast<char *>(v.data()), v.size() * sizeof(double));
    const std::size_t width = edges ? 14 : 7, n = v.size() / width;
    Eigen::Matrix<S, 3, 1> lo, hi;   // sampling bounds: not used by isValid
    lo << -1, -1, -1;
    hi << 1, 1, 1;
    State goal = stateOf(std::vector<double>{0, 0, 0, 1, 0, 0, 0}.data());
    Scenario scenario(argv[2], argv[3], goal, lo, hi, std::stod(argv[4]));
    std::vector<unsigned char> out(n);
    for (std::size_t i = 0; i < n; ++i) {
        const double *p = &v[width * i];
        out[i] = edges ? scenario.isValid(stateOf(p), stateOf(p + 7)) : scenario.isValid(stateOf(p));
    }
    std::ofstream o(argv[6], std::ios::binary);
    o.write(reinterpret_cast<const char *>(out.data()), out.size());
    return o ? 0 : 1;
}
