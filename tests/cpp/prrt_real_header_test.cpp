// Instantiates the reference's OWN Planner<Scenario, PRRT> (include/mpl/prrt.hpp, unmodified, found through -I) on the
// drop-in scenario with the nigh stand-in on the include path.  Needs Eigen (prrt.hpp -> interpolate.hpp) and the
// reference tree: tests/test_cpp_header.py builds it only where both exist.
#include <mplb/se3_rigid_body_scenario.hpp>
#include <mpl/prrt.hpp>
#include <cstdio>

using Scenario = mplb::demo::SE3RigidBodyScenario<double>;
using Planner = mpl::Planner<Scenario, mpl::PRRT>;

int main(int argc, char **argv) {
    if (argc < 3) {
        std::puts("usage: prrt_real_header_test env.dae robot.dae");
        return 2;
    }
    Scenario::State start, goal;
    std::get<0>(start) = Eigen::Quaterniond(0, 0, 1, 0);   // (w, x, y, z): scripts/twistycool.sh start 0,1,0,0 in x,y,z,w order
    std::get<1>(start) = Eigen::Vector3d(270, 160, -200);
    std::get<0>(goal) = Eigen::Quaterniond(0, 0, 1, 0);
    std::get<1>(goal) = Eigen::Vector3d(270, 160, -400);
    Eigen::Vector3d min(53.46, -21.25, -476.86), max(402.96, 269.25, -91.0);
    Planner planner(std::string(argv[1]), std::string(argv[2]), goal, min, max, 0.1);
    planner.addStart(start);
    int budget = 200000;
    planner.solve([&] { return planner.isSolved() || --budget < 0; });
    std::printf("%s nodes %zu\n", planner.isSolved() ? "solved" : "unsolved", planner.size());
    return planner.isSolved() ? 0 : 1;
}
