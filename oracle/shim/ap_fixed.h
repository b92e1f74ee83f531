// TEST INFRASTRUCTURE (oracle).  Float-typed stand-in for Xilinx `ap_fixed.h`,
// which the reference includes (GIN/src/dcl.h:13) but does not vendor.
//
// `ap_fixed<W,I>` here is a 4-byte wrapper around one IEEE fp32 value: no
// quantisation, no wrap.  With it the reference's kernel translation units
// compile unmodified and evaluate their exact statement order in fp32 -- the
// "fp32 csim flavour" that SURVEY.md section 8c defines as the parity oracle.
// Only `width`/`iwidth` are kept from the real type because the reference's
// util.h derives constants from them (GIN/src/util.h:27-32, PNA/src/util.h:34-46).
#ifndef FLOWGNN_ORACLE_SHIM_AP_FIXED_H
#define FLOWGNN_ORACLE_SHIM_AP_FIXED_H

template <int W, int I>
struct ap_fixed
{
    static constexpr int width = W;
    static constexpr int iwidth = I;

    float v;

    ap_fixed() = default;
    ap_fixed(float x) : v(x) {}
    ap_fixed(double x) : v(static_cast<float>(x)) {}
    ap_fixed(int x) : v(static_cast<float>(x)) {}

    operator float() const { return v; }

    ap_fixed& operator+=(float x) { v = v + x; return *this; }
    ap_fixed& operator-=(float x) { v = v - x; return *this; }
    ap_fixed& operator*=(float x) { v = v * x; return *this; }
    ap_fixed& operator/=(float x) { v = v / x; return *this; }
};

static_assert(sizeof(ap_fixed<16, 6>) == sizeof(float), "ap_fixed shim must be layout-compatible with float");

#endif
