// Stand-in for <nigh/lp_space.hpp>: unc::robotics::nigh::L1Space<S, dim> as the Fetch scenario uses it
// (include/mpl/demo/fetch_scenario.hpp:19,75-77): the state type and the unweighted L1 distance.
#pragma once
#ifndef MPLB_SHIM_NIGH_LP_SPACE_HPP
#define MPLB_SHIM_NIGH_LP_SPACE_HPP

#include <array>
#include <cmath>

#if defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#ifndef MPLB_SHIM_HAVE_EIGEN
#define MPLB_SHIM_HAVE_EIGEN 1
#endif
#endif
#endif

namespace unc::robotics::nigh {

template <class S, int dim>
struct L1Space {
    static_assert(dim >= 1 && dim <= 16, "mplb_nn keys hold at most 16 coordinates");
#ifdef MPLB_SHIM_HAVE_EIGEN
    using Type = Eigen::Matrix<S, dim, 1>;
#else
    using Type = std::array<S, dim>;
#endif
    using Distance = S;
    static void flatten(const Type &q, double *o) {
        for (int i = 0; i < dim; ++i) o[i] = static_cast<double>(q[i]);
    }
    static constexpr unsigned dimensions() { return dim; }
    static constexpr int metric() { return 1; }
    static constexpr int keyDim() { return dim; }
    static constexpr double so3Weight() { return 0.0; }
    static constexpr double l2Weight() { return 1.0; }
    Distance distance(const Type &a, const Type &b) const {
        double s = 0;
        for (int i = 0; i < dim; ++i) s += std::fabs(static_cast<double>(a[i]) - static_cast<double>(b[i]));
        return static_cast<Distance>(s);
    }
};

}  // namespace unc::robotics::nigh

#endif
