// Drop-in C++ host side of the Fetch scenario's validity interface over the mplb C ABI.
//
// Mirrors the validity part of mpl::demo::FetchScenario<S> (reference include/mpl/demo/fetch_scenario.hpp:11-177):
// same constructor arguments (envFrame, envMesh, goal, goalTol, checkResolution), isValid(q, report) and
// isValid(from, to, report), scale(), space().distance(), maxSteering(), randomSample().  Goal sampling and
// isGoal run the IK solver of FetchRobot in the reference (fetch_scenario.hpp:84-103, Eigen-only code, no
// collision work): they stay with the reference's header -- INTEGRATION.md shows the three-line patch that
// swaps only the isValid bodies.  Without Eigen, State is std::array<S,8> and Frame a row-major 3x4 array.
#pragma once
#ifndef MPLB_FETCH_SCENARIO_HPP
#define MPLB_FETCH_SCENARIO_HPP

#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../mplb.h"

#if defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#define MPLB_FETCH_HAVE_EIGEN 1
#endif
#endif

namespace mplb::demo {

template <class S>
class FetchScenario {
public:
    static constexpr int kDOF = 8;
#ifdef MPLB_FETCH_HAVE_EIGEN
    using State = Eigen::Matrix<S, kDOF, 1>;
    using Frame = Eigen::Transform<S, 3, Eigen::Isometry>;
#else
    using State = std::array<S, kDOF>;
    using Frame = std::array<S, 12>;   // row-major 3x4 [R|t]
#endif
    using Distance = S;

    class Space {   // nigh::L1Space<S, 8> as far as the planners use it
    public:
        using Type = State;
        Distance distance(const State &a, const State &b) const {
            Distance s = 0;
            for (int i = 0; i < kDOF; ++i) s += std::abs(a[i] - b[i]);
            return s;
        }
        static constexpr unsigned dimensions() { return kDOF; }
    };

private:
    mplb_scene *scene_ = nullptr;
    Space space_;

    static void frameToArray(const Frame &f, double out[12]) {
#ifdef MPLB_FETCH_HAVE_EIGEN
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) out[4 * r + c] = double(f.linear()(r, c));
            out[4 * r + 3] = double(f.translation()[r]);
        }
#else
        for (int i = 0; i < 12; ++i) out[i] = double(f[i]);
#endif
    }
    [[noreturn]] static void raise(int rc) {
        const std::string msg = mplb_last_error();
        if (rc == MPLB_ERR_IO || rc == MPLB_ERR_INVALID) throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
    }

public:
    // fetch_scenario.hpp:48-63 (goal / goalTol only matter to the IK-based goal handling)
    template <class Goal, class Tol>
    FetchScenario(const Frame &envFrame, const std::string &envMesh, const Goal &, const Tol &, S checkResolution = 0.01,
                  int device = 0) {
        double f[12];
        frameToArray(envFrame, f);
        int rc = mplb_scene_create_fetch_file(envMesh.c_str(), f, double(checkResolution), device, &scene_);
        if (rc != MPLB_OK) raise(rc);
    }
    FetchScenario(const Frame &envFrame, const double *envTris, std::size_t nEnv, S checkResolution = 0.01, int device = 0) {
        double f[12];
        frameToArray(envFrame, f);
        int rc = mplb_scene_create_fetch(envTris, nEnv, f, double(checkResolution), device, &scene_);
        if (rc != MPLB_OK) raise(rc);
    }
    FetchScenario(const FetchScenario &) = delete;
    FetchScenario &operator=(const FetchScenario &) = delete;
    ~FetchScenario() { mplb_scene_destroy(scene_); }

    static constexpr bool multiGoal = true;
    const Space &space() const { return space_; }
    Distance maxSteering() const { return std::numeric_limits<Distance>::infinity(); }

    // fetch_scenario.hpp:42-45,75-77
    static State scale(const State &q) {
        static const S k[kDOF] = {10, 4, 2, 1, 1, 1, 1, 1};
        State r = q;
        for (int i = 0; i < kDOF; ++i) r[i] = q[i] * k[i];
        return r;
    }

    // FetchRobot::randomConfig (fetch_robot.hpp:163-175)
    template <class RNG>
    State randomSample(RNG &rng) {
        static constexpr S PI = S(3.14159265358979323846264338327950288L);
        static const S lo[kDOF] = {0.0, -1.6056, -1.221, -4 * PI, -2.251, -4 * PI, -2.16, -4 * PI};
        static const S hi[kDOF] = {0.38615, 1.6056, 1.518, 4 * PI, 2.251, 4 * PI, 2.16, 4 * PI};
        State q;
        for (int i = 0; i < kDOF; ++i) {
            std::uniform_real_distribution<S> dist(lo[i], hi[i]);
            q[i] = dist(rng);
        }
        return q;
    }

    // fetch_scenario.hpp:105-118
    bool isValid(const State &q, bool /*report*/ = false) const {
        double s[kDOF];
        for (int i = 0; i < kDOF; ++i) s[i] = double(q[i]);
        std::uint8_t v = 0;
        int rc = mplb_check_states(scene_, s, 1, &v);
        if (rc != MPLB_OK) raise(rc);
        return v != 0;
    }
    // fetch_scenario.hpp:120-173
    bool isValid(const State &from, const State &to, bool /*report*/ = false) const {
        double a[kDOF], b[kDOF];
        for (int i = 0; i < kDOF; ++i) {
            a[i] = double(from[i]);
            b[i] = double(to[i]);
        }
        std::uint8_t v = 0;
        int rc = mplb_check_edges(scene_, a, b, 1, &v, nullptr);
        if (rc != MPLB_OK) raise(rc);
        return v != 0;
    }
    // batched entry points
    std::vector<std::uint8_t> isValidBatch(const std::vector<State> &qs) const {
        std::vector<double> s(kDOF * qs.size());
        for (std::size_t i = 0; i < qs.size(); ++i)
            for (int k = 0; k < kDOF; ++k) s[kDOF * i + k] = double(qs[i][k]);
        std::vector<std::uint8_t> v(qs.size());
        int rc = mplb_check_states(scene_, s.data(), qs.size(), v.data());
        if (rc != MPLB_OK) raise(rc);
        return v;
    }
    std::vector<std::uint8_t> isValidBatch(const std::vector<State> &from, const std::vector<State> &to) const {
        if (from.size() != to.size()) throw std::invalid_argument("from/to size mismatch");
        std::vector<double> a(kDOF * from.size()), b(kDOF * to.size());
        for (std::size_t i = 0; i < from.size(); ++i)
            for (int k = 0; k < kDOF; ++k) {
                a[kDOF * i + k] = double(from[i][k]);
                b[kDOF * i + k] = double(to[i][k]);
            }
        std::vector<std::uint8_t> v(from.size());
        int rc = mplb_check_edges(scene_, a.data(), b.data(), from.size(), v.data(), nullptr);
        if (rc != MPLB_OK) raise(rc);
        return v;
    }
    mplb_scene *handle() const { return scene_; }

    // nearest-neighbour key of a state (scale(q) as doubles) and the matching mplb_nn_create arguments
    static constexpr int kKeyDim = kDOF;
    static void key(const State &q, double *out) {
        const State r = scale(q);
        for (int i = 0; i < kDOF; ++i) out[i] = static_cast<double>(r[i]);
    }
    static int nnCreate(int device, mplb_nn **out) { return mplb_nn_create(1, kDOF, 0, 0, device, out); }
};

}  // namespace mplb::demo

#endif
