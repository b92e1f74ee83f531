// Shared device/host helpers for the FlowGNN B200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace fg {

// ---- error plumbing (the library never exits; SURVEY.md 8b "Errors") -------------------------
void set_last_error(const std::string& msg);

#define FG_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::fg::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +      \
                                 " (" __FILE__ ":" + std::to_string(__LINE__) + ")");              \
            return static_cast<int>(_e);                                                           \
        }                                                                                          \
    } while (0)

#define FG_TRY(expr)                                                                               \
    do {                                                                                           \
        int _r = (expr);                                                                           \
        if (_r != 0) return _r;                                                                    \
    } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory, once per (kernel, device): function attributes are
// per device, and a process may hold one context per GPU (api.cu)
int opt_in_smem(const void* kernel, int bytes);

constexpr int FG_ERR_INVALID = 10001;      // bad argument
constexpr int FG_ERR_LIMIT = 10002;        // graph exceeds a documented limit
constexpr int FG_ERR_STATE = 10003;        // call order (no weights / no batch)

template <typename T>
constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// ---- feature vocabularies (GIN/src/load_inputs.cc:5, message_passing.cc:3) --------------------
constexpr int ND_FEATURE = 9;
constexpr int ND_FEATURE_TOTAL = 173;
constexpr int EDGE_ATTR = 3;
constexpr int ED_FEATURE_PER_LAYER = 13;
constexpr int ED_COMBOS = 60;              // 5 * 6 * 2 distinct bond-attribute triples

#ifdef __CUDACC__

__device__ __constant__ const int c_nd_feature_offsets[ND_FEATURE] = {0, 119, 123, 135, 147, 157, 163, 169, 171};

// ap_fixed_relu: signbit(x) ? 0 : x (GIN/src/util.h:20-25).  NaN must pass through,
// so this is a compare-select, not fmaxf.
__device__ __forceinline__ float relu_f(float x) { return (x < 0.0f) ? 0.0f : x; }

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming store: written once, read by the next launch -> do not keep in L1
__device__ __forceinline__ void stg_f4_stream(float* p, const float4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) --------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// global -> shared bulk copy, completion signalled on `bar` (bytes: multiple of 16, both sides 16-B aligned)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// order generic-proxy smem accesses before later async-proxy (TMA) accesses to the same bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- cp.async (LDGSTS) for the weight k-chunks --------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#endif  // __CUDACC__

}  // namespace fg
