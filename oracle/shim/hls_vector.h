// TEST INFRASTRUCTURE (oracle).  `hls::vector<T,N>` (GAT/src/dcl.h:14,61-69) as a
// plain, unaligned array with element-wise arithmetic.  sizeof == N*sizeof(T)
// so GAT's node_feature_t stays 36 bytes like the host's std::array<int,9>
// (GAT/src/host.h:34).
#ifndef FLOWGNN_ORACLE_SHIM_HLS_VECTOR_H
#define FLOWGNN_ORACLE_SHIM_HLS_VECTOR_H

#include <array>
#include <cstddef>

namespace hls {
template <typename T, std::size_t N>
struct vector
{
    T d[N];

    vector() = default;
    vector(const T& x) { for (std::size_t i = 0; i < N; i++) d[i] = x; }

    T& operator[](std::size_t i) { return d[i]; }
    const T& operator[](std::size_t i) const { return d[i]; }

#define FLOWGNN_SHIM_VEC_COMPOUND(OP) \
    vector& operator OP(const vector& o) { for (std::size_t i = 0; i < N; i++) d[i] OP o.d[i]; return *this; } \
    vector& operator OP(const T& o) { for (std::size_t i = 0; i < N; i++) d[i] OP o; return *this; }
    FLOWGNN_SHIM_VEC_COMPOUND(+=)
    FLOWGNN_SHIM_VEC_COMPOUND(-=)
    FLOWGNN_SHIM_VEC_COMPOUND(*=)
    FLOWGNN_SHIM_VEC_COMPOUND(/=)
#undef FLOWGNN_SHIM_VEC_COMPOUND
};

#define FLOWGNN_SHIM_VEC_BINARY(OP, COMPOUND) \
    template <typename T, std::size_t N> \
    vector<T, N> operator OP(const vector<T, N>& a, const vector<T, N>& b) { vector<T, N> r = a; r COMPOUND b; return r; } \
    template <typename T, std::size_t N> \
    vector<T, N> operator OP(const vector<T, N>& a, const T& b) { vector<T, N> r = a; r COMPOUND b; return r; } \
    template <typename T, std::size_t N> \
    vector<T, N> operator OP(const T& a, const vector<T, N>& b) { vector<T, N> r(a); r COMPOUND b; return r; }
FLOWGNN_SHIM_VEC_BINARY(+, +=)
FLOWGNN_SHIM_VEC_BINARY(-, -=)
FLOWGNN_SHIM_VEC_BINARY(*, *=)
FLOWGNN_SHIM_VEC_BINARY(/, /=)
#undef FLOWGNN_SHIM_VEC_BINARY
}

#endif
