// Stand-in for <nigh/se3_space.hpp> (UNC-Robotics/nigh): the part of unc::robotics::nigh::SE3Space the reference uses
// (include/mpl/demo/se3_rigid_body_scenario.hpp:18-19,115-117; prrt.hpp:205-210; pcforest.hpp:99-102) -- the state type
// and the weighted SO(3) + L2 distance.  The distance is evaluated inline in mplb_distance's operation order (the same
// as the oracle's; compile without FMA contraction), so planner costs and nearest-neighbour distances agree bit for bit.
#pragma once
#ifndef MPLB_SHIM_NIGH_SE3_SPACE_HPP
#define MPLB_SHIM_NIGH_SE3_SPACE_HPP

#include <array>
#include <cmath>
#include <cstdint>
#include <tuple>

#if defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#define MPLB_SHIM_HAVE_EIGEN 1
#endif
#endif

namespace unc::robotics::nigh {

template <class S, std::intmax_t so3weight, std::intmax_t l2weight>
struct SE3Space {
#ifdef MPLB_SHIM_HAVE_EIGEN
    using Type = std::tuple<Eigen::Quaternion<S>, Eigen::Matrix<S, 3, 1>>;
    static void flatten(const Type &q, double *o) {
        const auto &r = std::get<0>(q).coeffs();   // x, y, z, w
        const auto &t = std::get<1>(q);
        for (int i = 0; i < 4; ++i) o[i] = static_cast<double>(r[i]);
        for (int i = 0; i < 3; ++i) o[4 + i] = static_cast<double>(t[i]);
    }
#else
    using Type = std::array<S, 7>;   // qx, qy, qz, qw, tx, ty, tz
    static void flatten(const Type &q, double *o) {
        for (int i = 0; i < 7; ++i) o[i] = static_cast<double>(q[i]);
    }
#endif
    using Distance = S;
    static constexpr unsigned dimensions() { return 6; }
    static constexpr int metric() { return 0; }
    static constexpr int keyDim() { return 7; }
    static constexpr double so3Weight() { return static_cast<double>(so3weight); }
    static constexpr double l2Weight() { return static_cast<double>(l2weight); }

    // so3weight * acos|qa . qb| + l2weight * ||ta - tb||, in the order libmplb and the oracle use
    Distance distance(const Type &a, const Type &b) const {
        double p[7], q[7];
        flatten(a, p);
        flatten(b, q);
        const double d = std::fabs(((p[0] * q[0] + p[1] * q[1]) + p[2] * q[2]) + p[3] * q[3]);
        const double so3 = d < 1.0 ? std::acos(d) : 0.0;
        const double dx = p[4] - q[4], dy = p[5] - q[5], dz = p[6] - q[6];
        return static_cast<Distance>(so3Weight() * so3 + l2Weight() * std::sqrt((dx * dx + dy * dy) + dz * dz));
    }
};

}  // namespace unc::robotics::nigh

#endif
