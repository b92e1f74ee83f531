// Stand-alone probe for the tcgen05 building blocks in flowgnn_b200/csrc/tc.cuh (run on a B200):
//   D[128 x N] = A[128 x K] * W[N x K]^T with A in TMEM (bf16 hi/lo split), W stationary in shared memory
//   (bf16 hi/lo split, no-swizzle K-major canonical layout), three products hi*hi + lo*hi + hi*lo, fp32 accumulate.
// Checks the descriptor encodings and TMEM layouts against a float64 CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/tc_probe tools/tc_probe.cu && /tmp/tc_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../flowgnn_b200/csrc/tc.cuh"

namespace fg { void set_last_error(const std::string&) {} }
using namespace fg;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int N, int K, int PRODUCTS>
__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    unsigned char* w_hi = smem;
    unsigned char* w_lo = smem + (size_t)N * K * 2;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    for (int i = tid; i < N * K; i += 128)
    {
        const int n = i / K, k = i % K;
        const uint32_t s = tc::split_bf16(W[i]);
        const size_t off = tc::b_offset_bytes(n, k, N);
        *reinterpret_cast<uint16_t*>(w_hi + off) = (uint16_t)(s & 0xFFFF);
        *reinterpret_cast<uint16_t*>(w_lo + off) = (uint16_t)(s >> 16);
    }
    fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    constexpr int A_HI = 0, A_LO = K / 2, D_COL = 256;

    // row `tid` of A -> TMEM (packed bf16 pairs)
    for (int c = 0; c < K / 2; c += 8)
    {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            const uint32_t s0 = tc::split_bf16(A[tid * K + 2 * (c + j)]), s1 = tc::split_bf16(A[tid * K + 2 * (c + j) + 1]);
            hi[j] = tc::pack_hi(s0, s1);
            lo[j] = tc::pack_lo(s0, s1);
        }
        tc::st8(lane_base + A_HI + c, hi);
        tc::st8(lane_base + A_LO + c, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();

    if (tid == 0)
    {
        tc::fence_after_sync();
        const uint32_t idesc = tc::idesc_bf16(128, N);
        const uint32_t hi_addr = smem_u32(w_hi), lo_addr = smem_u32(w_lo);
        bool acc = false;
        for (int p = 0; p < PRODUCTS; p++)
        {
            const uint32_t a_col = (p == 1) ? A_LO : A_HI;
            const uint32_t b_addr = (p == 2) ? lo_addr : hi_addr;
            for (int j = 0; j < K / 16; j++)
            {
                const uint64_t bd = tc::smem_desc(b_addr + (uint32_t)(2 * j) * N * 16, N * 16, 128);
                tc::mma_ts(tbase + D_COL, tbase + a_col + 8 * j, bd, idesc, acc);
                acc = true;
            }
        }
        tc::commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c = 0; c < N; c += 16)
    {
        uint32_t r[16];
        tc::ld16(lane_base + D_COL + c, r);
        tc::wait_ld();
#pragma unroll
        for (int j = 0; j < 16; j++) D[tid * N + c + j] = __uint_as_float(r[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

static float bf16_rn(float x)
{
    uint32_t u; memcpy(&u, &x, 4);
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000u;
    float y; memcpy(&y, &u, 4); return y;
}

template <int N, int K, int PRODUCTS>
static int run(const char* name, float scale)
{
    std::vector<float> A(128 * K), W(N * K), D(128 * N, -1.f);
    srand(1234 + N + K);
    for (auto& x : A) x = scale * ((rand() / (float)RAND_MAX) * 2.f - 1.f);
    for (auto& x : W) x = (rand() / (float)RAND_MAX) * 1.5f - 0.75f;
    float *dA, *dW, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, D.size() * 4));
    const int smem = 2 * N * K * 2;
    CK(cudaFuncSetAttribute(probe_kernel<N, K, PRODUCTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<N, K, PRODUCTS><<<1, 128, smem>>>(dA, dW, dD);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    int bad_r = -1, bad_c = -1;
    for (int r = 0; r < 128; r++)
        for (int c = 0; c < N; c++)
        {
            double ref = 0;
            for (int k = 0; k < K; k++)
            {
                const double a = PRODUCTS == 1 ? bf16_rn(A[r * K + k]) : A[r * K + k];
                const double w = PRODUCTS == 1 ? bf16_rn(W[c * K + k]) : W[c * K + k];
                ref += a * w;
            }
            const double e = fabs(ref - D[r * N + c]);
            if (e > max_err) { max_err = e; bad_r = r; bad_c = c; }
            max_ref = fmax(max_ref, fabs(ref));
        }
    const double tol = (PRODUCTS == 1 ? 2e-5 : 5e-5) * max_ref;
    printf("%s N=%d K=%d products=%d: max |err| %.3e (max |ref| %.3e, rel %.2e) at (%d,%d) -> %s\n", name, N, K, PRODUCTS, max_err, max_ref,
           max_err / max_ref, bad_r, bad_c, max_err <= tol ? "OK" : "MISMATCH");
    if (max_err > tol)
        for (int c = 0; c < 8; c++) printf("   D[0][%d] = %g   D[1][%d] = %g  D[127][%d] = %g\n", c, D[c], c, D[N + c], c, D[127 * N + c]);
    cudaFree(dA); cudaFree(dW); cudaFree(dD);
    return max_err <= tol ? 0 : 1;
}

int main()
{
    int bad = 0;
    bad += run<208, 112, 1>("gemm1 bf16", 4.f);
    bad += run<112, 208, 1>("gemm2 bf16", 4.f);
    bad += run<208, 112, 3>("gemm1 3xbf16", 300.f);
    bad += run<112, 208, 3>("gemm2 3xbf16", 300.f);
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}
