// Drop-in C++ host side of the SE(3) rigid-body scenario over the mplb C ABI.
//
// Mirrors mpl::demo::SE3RigidBodyScenario<S, so3weight, l2weight>
// (reference include/mpl/demo/se3_rigid_body_scenario.hpp:15-179): same constructor arguments,
// same Scenario-concept members (scale, space, maxSteering, sampleGoal, randomSample, isGoal,
// isValid(q), isValid(from,to)).  Planner<Scenario, PRRT/PCForest> (prrt.hpp:35-76, pcforest.hpp:42-104)
// instantiate nigh on Scenario::Space: with nigh -- or include/mplb/shim, which resolves nigh's names to
// the GPU nearest-neighbour structure -- on the include path, Space is nigh's SE3Space and those planners
// compile against this class unchanged (tests/cpp/prrt_pattern_test.cpp).  The two isValid overloads forward to libmplb.so; the batched members are what the
// GPU path adds underneath.  With Eigen available State is the reference's
// std::tuple<Eigen::Quaternion<S>, Eigen::Matrix<S,3,1>>; without it a 7-double array in the same
// coefficient order (qx,qy,qz,qw,tx,ty,tz).
#pragma once
#ifndef MPLB_SE3_RIGID_BODY_SCENARIO_HPP
#define MPLB_SE3_RIGID_BODY_SCENARIO_HPP

#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <limits>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../mplb.h"

#if defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#include <tuple>
#define MPLB_HAVE_EIGEN 1
#endif
// nigh itself, or include/mplb/shim (the stand-in that routes nigh's names to the GPU structure), on the include path:
// Scenario::Space is then the very type the reference planners instantiate nigh with (prrt.hpp:23,35-42)
#if __has_include(<nigh/se3_space.hpp>)
#include <nigh/se3_space.hpp>
#define MPLB_HAVE_NIGH_SPACE 1
#endif
#endif

namespace mplb::demo {

namespace detail {
#ifdef MPLB_HAVE_EIGEN
template <class S>
using SE3State = std::tuple<Eigen::Quaternion<S>, Eigen::Matrix<S, 3, 1>>;
template <class S>
inline void pack(const SE3State<S> &q, double *o) {
    const auto &r = std::get<0>(q).coeffs();   // x,y,z,w
    const auto &t = std::get<1>(q);
    for (int i = 0; i < 4; ++i) o[i] = static_cast<double>(r[i]);
    for (int i = 0; i < 3; ++i) o[4 + i] = static_cast<double>(t[i]);
}
template <class S>
inline SE3State<S> unpack(const double *o) {
    SE3State<S> q;
    for (int i = 0; i < 4; ++i) std::get<0>(q).coeffs()[i] = static_cast<S>(o[i]);
    for (int i = 0; i < 3; ++i) std::get<1>(q)[i] = static_cast<S>(o[4 + i]);
    return q;
}
#else
template <class S>
using SE3State = std::array<S, 7>;
template <class S>
inline void pack(const SE3State<S> &q, double *o) {
    for (int i = 0; i < 7; ++i) o[i] = static_cast<double>(q[i]);
}
template <class S>
inline SE3State<S> unpack(const double *o) {
    SE3State<S> q;
    for (int i = 0; i < 7; ++i) q[i] = static_cast<S>(o[i]);
    return q;
}
#endif
}  // namespace detail

template <class S, std::intmax_t so3weight = 50, std::intmax_t l2weight = 1>
class SE3RigidBodyScenario {
public:
    using State = detail::SE3State<S>;
    using Distance = S;
#ifdef MPLB_HAVE_EIGEN
    using Bound = Eigen::Matrix<S, 3, 1>;
#else
    using Bound = std::array<S, 3>;
#endif

#ifdef MPLB_HAVE_NIGH_SPACE
    using Space = unc::robotics::nigh::SE3Space<S, so3weight, l2weight>;   // se3_rigid_body_scenario.hpp:18
    static_assert(std::is_same<typename Space::Type, State>::value, "nigh's SE3 state type differs from the drop-in's");
#else
    // nigh::SE3Space<S, so3weight, l2weight> as far as the planners use it (distance only)
    class Space {
        const mplb_scene *scene_ = nullptr;
        friend class SE3RigidBodyScenario;
    public:
        using Type = State;
        Distance distance(const State &a, const State &b) const {
            double pa[7], pb[7];
            detail::pack<S>(a, pa);
            detail::pack<S>(b, pb);
            return static_cast<Distance>(mplb_distance(scene_, pa, pb));
        }
        static constexpr unsigned dimensions() { return 6; }
    };
#endif

private:
    mplb_scene *scene_ = nullptr;
    Space space_;
    State goal_;
    Bound min_, max_;
    Distance goalRadius_{0.1};
    Distance invStepSize_;
    mutable std::atomic<unsigned> calls_{0};

    void bindSpace() {
#ifndef MPLB_HAVE_NIGH_SPACE
        space_.scene_ = scene_;
#endif
    }
    [[noreturn]] static void raise(int rc) {
        const std::string msg = mplb_last_error();
        if (rc == MPLB_ERR_IO || rc == MPLB_ERR_INVALID) throw std::invalid_argument(msg);  // load_mesh.hpp:250-254
        throw std::runtime_error(msg);
    }

public:
    template <class Min, class Max>
    SE3RigidBodyScenario(const std::string &envMesh, const std::string &robotMesh, const State &goal,
                         const Min &min, const Max &max, S checkResolution, int device = 0)
        : goal_(goal), invStepSize_(1 / checkResolution) {
        for (int i = 0; i < 3; ++i) {
            min_[i] = min[i];
            max_[i] = max[i];
        }
        int rc = mplb_scene_create_se3_files(envMesh.c_str(), robotMesh.c_str(), double(so3weight), double(l2weight),
                                             double(checkResolution), device, &scene_);
        if (rc != MPLB_OK) raise(rc);
        bindSpace();
    }
    // the same scene from explicit triangle soups (float64 [n][3][3], robot already centred): what
    // load(envMesh,false,false) / load(robotMesh,true,false) would have produced (se3...:58-59)
    template <class Min, class Max>
    SE3RigidBodyScenario(const double *envTris, std::size_t nEnv, const double *robotTris, std::size_t nRobot,
                         const State &goal, const Min &min, const Max &max, S checkResolution, int device = 0)
        : goal_(goal), invStepSize_(1 / checkResolution) {
        for (int i = 0; i < 3; ++i) {
            min_[i] = min[i];
            max_[i] = max[i];
        }
        int rc = mplb_scene_create_se3(envTris, nEnv, robotTris, nRobot, double(so3weight), double(l2weight),
                                       double(checkResolution), device, &scene_);
        if (rc != MPLB_OK) raise(rc);
        bindSpace();
    }
    SE3RigidBodyScenario(const SE3RigidBodyScenario &) = delete;
    SE3RigidBodyScenario &operator=(const SE3RigidBodyScenario &) = delete;
    SE3RigidBodyScenario(SE3RigidBodyScenario &&o) noexcept
        : scene_(o.scene_), space_(o.space_), goal_(o.goal_), min_(o.min_), max_(o.max_),
          goalRadius_(o.goalRadius_), invStepSize_(o.invStepSize_), calls_(o.calls_.load()) {
        o.scene_ = nullptr;
    }
    ~SE3RigidBodyScenario() { mplb_scene_destroy(scene_); }

    static constexpr bool multiGoal = true;
    static State scale(const State &q) { return q; }
    const Space &space() const { return space_; }
    Distance maxSteering() const { return std::numeric_limits<Distance>::infinity(); }
    unsigned calls() const { return calls_.load(); }

    template <class RNG>
    std::optional<State> sampleGoal(RNG &) { return goal_; }

    // se3...:107-113 with randomize.hpp:12-38
    template <class RNG>
    State randomSample(RNG &rng) {
        static constexpr S PI = S(3.1415926535897932384626433832795028841968L);
        std::uniform_real_distribution<S> dist01(0, 1), dist2pi(0, 2 * PI);
        S a = dist01(rng), b = dist2pi(rng), c = dist2pi(rng);
        double o[7];
        o[0] = std::sqrt(1 - a) * std::sin(b);
        o[1] = std::sqrt(1 - a) * std::cos(b);
        o[2] = std::sqrt(a) * std::sin(c);
        o[3] = std::sqrt(a) * std::cos(c);
        for (int i = 0; i < 3; ++i) {
            std::uniform_real_distribution<S> d(min_[i], max_[i]);
            o[4 + i] = d(rng);
        }
        return detail::unpack<S>(o);
    }

    bool isGoal(const State &q) const { return space_.distance(goal_, q) <= goalRadius_; }

    // se3...:119-126
    bool isValid(const State &q) const {
        ++calls_;
        double s[7];
        detail::pack<S>(q, s);
        std::uint8_t v = 0;
        int rc = mplb_check_states(scene_, s, 1, &v);
        if (rc != MPLB_OK) raise(rc);
        return v != 0;
    }

    // se3...:134-178 (one launch expands and checks every interpolated state of the edge)
    bool isValid(const State &from, const State &to) const {
        double a[7], b[7];
        detail::pack<S>(from, a);
        detail::pack<S>(to, b);
        std::uint8_t v = 0;
        std::uint64_t checked = 0;
        int rc = mplb_check_edges(scene_, a, b, 1, &v, &checked);
        if (rc != MPLB_OK) raise(rc);
        calls_ += static_cast<unsigned>(checked);
        return v != 0;
    }

    // ---- batched entry points (the data-parallel axis the GPU path adds) ----
    std::vector<std::uint8_t> isValidBatch(const std::vector<State> &qs) const {
        std::vector<double> s(7 * qs.size());
        for (std::size_t i = 0; i < qs.size(); ++i) detail::pack<S>(qs[i], &s[7 * i]);
        std::vector<std::uint8_t> v(qs.size());
        int rc = mplb_check_states(scene_, s.data(), qs.size(), v.data());
        if (rc != MPLB_OK) raise(rc);
        calls_ += static_cast<unsigned>(qs.size());
        return v;
    }
    std::vector<std::uint8_t> isValidBatch(const std::vector<State> &from, const std::vector<State> &to) const {
        if (from.size() != to.size()) throw std::invalid_argument("from/to size mismatch");
        std::vector<double> a(7 * from.size()), b(7 * to.size());
        for (std::size_t i = 0; i < from.size(); ++i) {
            detail::pack<S>(from[i], &a[7 * i]);
            detail::pack<S>(to[i], &b[7 * i]);
        }
        std::vector<std::uint8_t> v(from.size());
        std::uint64_t checked = 0;
        int rc = mplb_check_edges(scene_, a.data(), b.data(), from.size(), v.data(), &checked);
        if (rc != MPLB_OK) raise(rc);
        calls_ += static_cast<unsigned>(checked);
        return v;
    }
    mplb_scene *handle() const { return scene_; }

    // nearest-neighbour key of a state (scale(q) as doubles) and the matching mplb_nn_create arguments, for
    // planners that keep their nodes in an mplb_nn (include/mplb/batched_rrt.hpp)
    static constexpr int kKeyDim = 7;
    static void key(const State &q, double *out) { detail::pack<S>(scale(q), out); }
    static State fromKey(const double *k) { return detail::unpack<S>(k); }
    static int nnCreate(int device, mplb_nn **out) { return mplb_nn_create(0, 7, double(so3weight), double(l2weight), device, out); }
};

}  // namespace mplb::demo

#endif
