// TEST INFRASTRUCTURE (oracle).  Bit-level emulation of the subset of Xilinx `ap_fixed<16,I>` (default modes
// AP_TRN quantisation = round toward minus infinity, AP_WRAP overflow = keep the low 16 bits) that the reference's
// GIN and DGN kernel sources use (GIN/src/dcl.h:58-59 `ap_fixed<16,6>`, DGN/src/dcl.h:54-55 `ap_fixed<16,3>`).
// The Vitis HLS 2021.1 header itself is not vendored by the reference and is not in this image, so this
// is a restatement of its published semantics; there is no fixed-point golden output in the reference to pin it
// against ("parity unpinned" at the bit level; the product is held bit-exact to THIS emulation).
//
// How it works.  In Vitis every binary operator on ap_fixed values returns a WIDER exact type (a product of two
// <16,I> values is <32,2I>, a sum grows by one bit) and only the assignment back to a <16,I> variable quantises.
// Here a value converts implicitly to `double`, the builtin double operators evaluate the expression, and the
// constructor quantises: every intermediate of the reference's expressions is a dyadic rational with at most
// 2 x 13 fraction bits and magnitude below 2^22, which a double (53-bit significand) holds exactly, so
// "evaluate in double, then floor(x * 2^F) mod 2^16" equals "evaluate in the wide fixed type, then truncate and wrap".
// Division is NOT exact and is emulated separately (see operator/ below): Vitis computes the quotient with an integer
// division of the raw operands (ap_fixed_base::operator/: dividend = V << max(F2, 0); r.V = dividend sdiv divisor), which
// truncates TOWARD ZERO at the quotient type's last fraction bit.
#ifndef FLOWGNN_ORACLE_SHIM_FIXED_AP_FIXED_H
#define FLOWGNN_ORACLE_SHIM_FIXED_AP_FIXED_H

#include <cmath>
#include <cstdint>

template <int W, int I>
struct ap_fixed
{
    static_assert(W == 16, "the emulation stores 16-bit values only");
    static constexpr int width = W;
    static constexpr int iwidth = I;
    static constexpr int F = W - I;

    int16_t raw;

    static int16_t quantise(double x)
    {
        const double scaled = x * static_cast<double>(1 << F);                 // exact (power of two)
        long long q = static_cast<long long>(scaled);                          // toward zero ...
        if (static_cast<double>(q) > scaled) q--;                              // ... corrected to floor: AP_TRN
        return static_cast<int16_t>(static_cast<uint16_t>(q));                 // AP_WRAP
    }

    ap_fixed() = default;
    ap_fixed(double x) : raw(quantise(x)) {}
    ap_fixed(float x) : raw(quantise(static_cast<double>(x))) {}
    ap_fixed(int x) : raw(quantise(static_cast<double>(x))) {}

    operator double() const { return static_cast<double>(raw) * (1.0 / static_cast<double>(1 << F)); }

    ap_fixed& operator+=(double x) { raw = quantise(static_cast<double>(*this) + x); return *this; }
    ap_fixed& operator-=(double x) { raw = quantise(static_cast<double>(*this) - x); return *this; }
    ap_fixed& operator*=(double x) { raw = quantise(static_cast<double>(*this) * x); return *this; }
};

// ap_fixed<16,I> / int (GIN/src/finalize.cc:112, DGN/src/node_embedding.cc:145): the int is an ap_fixed<32,32>, so the
// quotient type keeps F fraction bits and r.V = raw sdiv n, C++ integer division (toward zero).  n == 0 is undefined
// in the reference (hardware divider); the emulation returns 0 so that the checker does not trap.
template <int W, int I>
inline double operator/(const ap_fixed<W, I>& a, int n)
{
    const int q = (n == 0) ? 0 : static_cast<int>(a.raw) / n;
    return std::ldexp(static_cast<double>(q), -ap_fixed<W, I>::F);
}

// wide expression / ap_fixed<16,I> (DGN/src/node_embedding.cc:146, the only call site): the numerator there is
// <16,I> - <16,I> * <16,I>, an exact value with 2F fraction bits; the divisor has F, so dividend = V << F and the
// quotient keeps 2F fraction bits, truncated toward zero.  Divisor 0 cannot happen at the call site (replaced by epsilon).
template <int W, int I>
inline double operator/(double num, const ap_fixed<W, I>& den)
{
    constexpr int F = ap_fixed<W, I>::F;
    const long long nraw = static_cast<long long>(std::ldexp(num, 2 * F));      // exact: num is a multiple of 2^-2F
    const long long q = (den.raw == 0) ? 0 : (nraw * (1ll << F)) / static_cast<long long>(den.raw);
    return std::ldexp(static_cast<double>(q), -2 * F);
}

static_assert(sizeof(ap_fixed<16, 6>) == 2, "ap_fixed emulation must be layout-compatible with int16");

#endif
