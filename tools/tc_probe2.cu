// Probe: thread/register <-> (TMEM lane, column) mapping of tcgen05.st/ld .16x256b (run on a B200).
#include <cstdio>
#include <cstdlib>
#include "../flowgnn_b200/csrc/tc.cuh"
namespace fg { void set_last_error(const std::string&) {} }
using namespace fg;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out, uint32_t* out2)
{
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 64); tc::tmem_relinquish(); }
    tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    // zero 32 columns
    for (int c = 0; c < 32; c += 8) { const uint32_t z[8] = {0,0,0,0,0,0,0,0}; tc::st8(lane_base + c, z); }
    tc::wait_st();
    __syncwarp();
    // 16x256b.x1 to lanes [0,16) at column 0 and lanes [16,32) at column 8 ; x2 at lanes [0,16) column 16
    {
        uint32_t r0 = 0x1000 | (lane << 4) | 0, r1 = 0x1000 | (lane << 4) | 1, r2 = 0x1000 | (lane << 4) | 2, r3 = 0x1000 | (lane << 4) | 3;
        asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(lane_base + 0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
        r0 |= 0x2000; r1 |= 0x2000; r2 |= 0x2000; r3 |= 0x2000;
        asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(lane_base + (16u << 16) + 8), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
        uint32_t q[8];
        for (int i = 0; i < 8; i++) q[i] = 0x4000 | (lane << 4) | i;
        asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(lane_base + 16), "r"(q[0]), "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7]) : "memory");
    }
    tc::wait_st();
    __syncwarp();
    for (int c = 0; c < 32; c += 16)
    {
        uint32_t r[16];
        tc::ld16(lane_base + c, r);
        tc::wait_ld();
        for (int j = 0; j < 16; j++) out[tid * 32 + c + j] = r[j];
    }
    // and the matching load shape: read columns 0..7 of lanes 0..15 with 16x256b.x1
    {
        uint32_t a, b, c, d;
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(lane_base + 0) : "memory");
        tc::wait_ld();
        out2[tid * 4 + 0] = a; out2[tid * 4 + 1] = b; out2[tid * 4 + 2] = c; out2[tid * 4 + 3] = d;
    }
    tc::fence_before_sync(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 64);
}

int main()
{
    uint32_t *d, *d2; CK(cudaMalloc(&d, 128 * 32 * 4)); CK(cudaMalloc(&d2, 128 * 4 * 4));
    probe<<<1, 128>>>(d, d2); CK(cudaDeviceSynchronize());
    static uint32_t h[128 * 32], h2[128 * 4];
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2, d2, sizeof(h2), cudaMemcpyDeviceToHost));
    for (int warp = 0; warp < 2; warp++)
    {
        printf("warp %d: TMEM lane x column contents (hex: Txxr = pattern T, source lane xx, register r)\n", warp);
        for (int l = 0; l < 32; l++) { printf("lane %2d:", l); for (int c = 0; c < 32; c++) printf(" %04x", h[(warp * 32 + l) * 32 + c]); printf("\n"); }
    }
    printf("16x256b.x1 load of lanes 0..15 cols 0..7 (warp 0): thread -> 4 regs\n");
    for (int t = 0; t < 32; t++) printf("t%2d: %04x %04x %04x %04x\n", t, h2[t * 4], h2[t * 4 + 1], h2[t * 4 + 2], h2[t * 4 + 3]);
    return 0;
}
