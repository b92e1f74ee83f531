// TEST INFRASTRUCTURE (oracle).  `hls::` math for the ap_fixed emulation (shim_fixed/ap_fixed.h).  GIN and DGN need only
// signbit (*/src/util.h:24) and abs (DGN/src/load_inputs.cc:109, node_embedding.cc:146).  sqrt / recip exist because
// GIN/src/load_inputs.cc:130 calls them on a value it never reads again (dead code inherited from GCN); their results
// do not reach the output and are NOT bit-accurate to Vitis' fixed-point CORDIC.
#ifndef FLOWGNN_ORACLE_SHIM_FIXED_HLS_MATH_H
#define FLOWGNN_ORACLE_SHIM_FIXED_HLS_MATH_H

#include <cmath>

#include "ap_fixed.h"

namespace hls {
template <int W, int I>
inline bool signbit(const ap_fixed<W, I>& x) { return x.raw < 0; }
// hls::abs on ap_fixed<W,I> returns the same type: -x of the most negative value wraps back to itself
template <int W, int I>
inline ap_fixed<W, I> abs(const ap_fixed<W, I>& x) { return (x.raw < 0) ? ap_fixed<W, I>(-static_cast<double>(x)) : x; }
inline double sqrt(double x) { return std::sqrt(x); }
inline double recip(double x) { return 1.0 / x; }
}

#endif
