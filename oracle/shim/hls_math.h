// TEST INFRASTRUCTURE (oracle).  fp32 stand-ins for the `hls::` math the
// reference calls: signbit (*/src/util.h:24), sqrt/recip (GCN/src/load_inputs.cc:32,122),
// exp (GAT/src/message_passing.cc:128), log (PNA/src/load_inputs.cc:105),
// abs (DGN/src/load_inputs.cc:109).  Each takes and returns float so that the
// float-wrapper `ap_fixed` converts implicitly and unambiguously.
#ifndef FLOWGNN_ORACLE_SHIM_HLS_MATH_H
#define FLOWGNN_ORACLE_SHIM_HLS_MATH_H

#include <cmath>

namespace hls {
constexpr bool signbit(float x) { return x < 0.0f; }
inline float sqrt(float x) { return std::sqrt(x); }
inline float recip(float x) { return 1.0f / x; }
inline float exp(float x) { return std::exp(x); }
inline float log(float x) { return std::log(x); }
inline float abs(float x) { return std::fabs(x); }
}

#endif
