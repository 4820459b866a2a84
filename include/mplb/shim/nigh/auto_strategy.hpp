// Stand-in for <nigh/auto_strategy.hpp>: the names prrt.hpp:35-42 and pcforest.hpp:42-51 spell --
//   using Concurrency = unc::robotics::nigh::Concurrent;
//   using NNStrategy  = unc::robotics::nigh::auto_strategy_t<Space, Concurrency>;
//   unc::robotics::nigh::Nigh<Node*, Space, NodeKey, Concurrency, NNStrategy> nn_;
// -- resolved to the GPU nearest-neighbour structure (mplb::NighAdapter over mplb_nn_*).  With this directory on the
// include path in place of nigh's, the reference planners compile unchanged and their nn_ lives on the device.
#pragma once
#ifndef MPLB_SHIM_NIGH_AUTO_STRATEGY_HPP
#define MPLB_SHIM_NIGH_AUTO_STRATEGY_HPP

#include "../../nigh_adapter.hpp"

namespace unc::robotics::nigh {

struct Concurrent {};
struct NoThreadSafety {};
struct MplbBruteForce {};   // the only strategy there is: the exact tiled scan of nn_kernels.cu

template <class Space, class Concurrency>
using auto_strategy_t = MplbBruteForce;

template <class T, class Space, class KeyFn, class Concurrency = Concurrent, class Strategy = MplbBruteForce>
using Nigh = ::mplb::NighAdapter<T, Space, KeyFn>;

}  // namespace unc::robotics::nigh

#endif
