// CTA-pair (tcgen05 cta_group::2) building blocks shared by the GIN layer kernels gin_fused.cu and gin_tc2.cu:
// cluster primitives, remote mbarrier arrivals, parked waits, pair MMA / commit / TMEM allocation wrappers, the
// bf16 hi/lo split, the z conversion epilogue.  See gin_tc2.cu for the design notes.
#pragma once

#include "internal.cuh"
#include "layers.cuh"
#include "tc.cuh"

namespace fg {
namespace pair {

constexpr unsigned FULL = 0xFFFFFFFFu;

// ---- cluster / pair primitives ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    // default semantics (release, CTA scope) as in CUTLASS' ClusterBarrier::arrive(cta_id): what the waiter consumes was
    // either written through the async proxy after fence.proxy.async or lives in tensor memory behind tcgen05 fences;
    // .release.cluster would add a GPU-scope MEMBAR to every arrival
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// mbarrier wait with a suspend-time hint: the thread is parked by the hardware until the phase completes (or the hint
// expires) instead of re-issuing try_wait in a tight loop -- the spinning warps of the other roles otherwise take a
// fifth of all issue slots (and of the power budget the kernel runs into)
__device__ __forceinline__ void mbar_wait_park(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (!done);
}
// wait with cluster-scope acquire: the arrivals may come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// One lane of a fully active warp (the lowest): the surrounding code -- descriptor arithmetic, loop control -- runs
// convergently on all 32 lanes and therefore in the uniform datapath, only the tcgen05 instruction itself is
// predicated.  Issuing from a branch on `lane == 0` instead makes ptxas wrap every MMA in an elect-and-retry loop and
// move the descriptors through R2UR: ~150 cycles per MMA, three times the time the tensor pipe needs to execute it.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// arrive on the barrier at the same shared-memory offset in both CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void commit2(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// D[tmem, 256 x N over the pair] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T, one K = 16 step
__device__ __forceinline__ void mma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// The same three instructions issued by the elected lane of a CONVERGED warp: elect.sync inside the asm block, the tcgen05
// instruction predicated on it.  The loop around them then runs on all 32 lanes (uniform datapath, descriptors in uniform
// registers) and ptxas emits one predicated UTCHMMA per MMA -- issued from a `lane == 0` branch every MMA is wrapped in an
// elect-and-retry loop with R2UR moves (~150 cycles per MMA, three times the time the tensor pipe needs to execute it).
__device__ __forceinline__ void commit2_elect(uint64_t* bar)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xFFFFFFFF;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void mma_ss2_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xFFFFFFFF;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_ts2_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xFFFFFFFF;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// relu that lets NaN through, like the reference's compare-select (GIN/src/util.h:20-25), in ONE instruction
__device__ __forceinline__ float relu_nan(float x)
{
    float y;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(0.0f));
    return y;
}

template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// two fp32 additions in one instruction (FADD2, sm_100): same IEEE round-to-nearest result per element
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    unsigned long long ua, ub, ud;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(ud));
    return d;
}

// (x0, x1) -> packed bf16 pairs: hi = rn(x), lo = rn(x - hi); element 0 in the low half
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16);
    const float r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
// the same with the two residuals in one FADD2 (used where register pairs are cheap: the epilogue warps; in the gather
// warps the pair alignment costs more registers than the 80 they have)
__device__ __forceinline__ void split2_p(float x0, float x1, uint32_t& hi, uint32_t& lo)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float2 r = add2(make_float2(x0, x1), make_float2(-__uint_as_float(hi << 16), -__uint_as_float(hi & 0xFFFF0000u)));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r.y), "f"(r.x));
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b)
{
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// 16-byte load that is not issued when `on` is false (reads as zero).  Plain C++ on purpose: ptxas turns this into
// "zero the quad, @p LDG into the same quad"; an inline-asm version made it load into a scratch quad and copy, i.e.
// wait for every load right after issuing it.
__device__ __forceinline__ float4 ldg_f4_if(const float* ptr, bool on)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) v = __ldg(reinterpret_cast<const float4*>(ptr));
    return v;
}

__device__ __forceinline__ void acc_edge(float4& m, const float4& t, const float4& h)
{
    m.x += relu_nan(t.x + h.x); m.y += relu_nan(t.y + h.y); m.z += relu_nan(t.z + h.z); m.w += relu_nan(t.w + h.w);
}

// a register copy the compiler cannot fold: the consumer of a prefetched value waits for its load HERE, once, and
// later uses of the copy carry no scoreboard dependency that would serialise them behind the feature-row loads
__device__ __forceinline__ int reg_copy(int x)
{
    int y;
    asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}


// z = relu(acc) for 16 accumulator columns of this thread's row (b1 is already in acc: the A tile carries a constant-1
// column k = 100 and W1 the bias in that column) -> bf16 hi/lo, written back in place
__device__ __forceinline__ void convert_regs(uint32_t zaddr, const uint32_t (&r)[16])
{
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; j++) split2_p(relu_nan(__uint_as_float(r[2 * j])), relu_nan(__uint_as_float(r[2 * j + 1])), hi[j], lo[j]);
    tc::st8(zaddr, hi);
    tc::st8(zaddr + 8, lo);
}
// chunks c, c + 2, ... < c_end of 16 columns each; the TMEM load of the next chunk is in flight while one is converted
__device__ __forceinline__ void convert_range(uint32_t zbase, int c, int c_end)
{
    uint32_t r0[16], r1[16];
    if (c >= c_end) return;
    tc::ld16(zbase + 16 * c, r0);
    while (true)
    {
        tc::wait_ld();
        if (c + 2 < c_end) tc::ld16(zbase + 16 * (c + 2), r1);
        convert_regs(zbase + 16 * c, r0);
        c += 2;
        if (c >= c_end) break;
        tc::wait_ld();
        if (c + 2 < c_end) tc::ld16(zbase + 16 * (c + 2), r0);
        convert_regs(zbase + 16 * c, r1);
        c += 2;
        if (c >= c_end) break;
    }
}

}  // namespace pair
}  // namespace fg
