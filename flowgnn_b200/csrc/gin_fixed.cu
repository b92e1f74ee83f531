// GIN / GIN-VN forward in the reference's OWN arithmetic: ap_fixed<16,6> (GIN/src/dcl.h:58-59), option "fixed_point".
//
// Every feature and weight is a 16-bit two's-complement integer with 10 fraction bits.  Vitis evaluates an expression
// in a wider exact type and quantises on assignment with the default modes AP_TRN (floor) and AP_WRAP (keep the low
// 16 bits), so, with raw integers (value = raw / 1024):
//   x = a + b            ->  raw = (a + b) mod 2^16
//   acc += a * w         ->  raw = (acc + floor(a * w / 1024)) mod 2^16          (linear.cc:41, node_embedding.cc:123-124,173)
//   relu(x)              ->  raw < 0 ? 0 : raw  on the WRAPPED value               (util.h:20-25)
//   x / n  (n an int)    ->  raw = trunc_toward_zero(raw / n)                      (finalize.cc:112; ap_fixed_base::operator/ is an
//                                                                                   integer division of the raw operands)
// Wrap-around addition is associative, so any summation order gives the reference's bits; only the per-product floor
// must be kept.  The node MLP therefore runs on the integer pipe, not on the tensor cores (a sum of individually floored
// products is not a matrix product), two instructions per multiply-accumulate (fixed.cuh::mac_floor):
//   IMAD    p = a * w + 2^30        raw operands; |a * w| <= 2^30, so p is a non-negative 32-bit number
//   LEA.HI  acc += p >> 10          logical shift = floor(a * w / 1024) + 2^20, and the 2^20 vanish mod 2^16
// (a single `mad.hi.s32` on pre-shifted operands computes the same and is what the pool / head kernel uses, but IMAD.HI
// issues at a quarter of the IMAD rate: 40 ms per 41,127-graph batch against the two-instruction form's time in DESIGN.md).
//
// Checked bit for bit against the reference's unmodified sources compiled over an ap_fixed emulation (tests/test_fixed_point.py).
#include "internal.cuh"
#include "fixed.cuh"

#include <algorithm>

namespace fg {

namespace {

constexpr int D = 100;               // EMB_DIM
constexpr int H = 200;               // MLP_1_OUT
constexpr int TM = 48;               // nodes per tile
constexpr int NT = 320;              // 300 compute threads (50 x 6 and 25 x 12 register tiles) + 20 idle
constexpr int W1_OFF = 0;                               // int32 [100][200]  raw w, k-major
constexpr int W2_OFF = W1_OFF + 4 * D * H;              // int32 [200][100]
constexpr int ACT_OFF = W2_OFF + 4 * H * D;             // int32 [TM][100]   raw a
constexpr int HID_OFF = ACT_OFF + 4 * D * TM;           // int32 [TM][200]   raw relu(z)
constexpr int SMEM_BYTES = HID_OFF + 4 * H * TM;        // 217,600

struct FixedLayerParams {
    const int16_t* h_in; int16_t* h_out;
    const int* in_ptr; const int* src; const uint8_t* code;
    const int16_t* ee;               // [60][100] this layer, the three tables of a bond triple already summed (mod 2^16)
    const int* w1; const int* b1;    // [100][200] raw << 6, [200] raw
    const int* w2; const int* b2;    // [200][100] raw << 6, [100] raw
    int num_nodes; int num_tiles; int relu_out;
};

// One layer: message passing (message_passing.cc:77-150) + the two-layer node MLP (node_embedding.cc:84-201).
// eps is never loaded by the reference kernel, so (1 + eps) * h == h exactly.
__global__ void __launch_bounds__(NT, 1) gin_fixed_layer_kernel(FixedLayerParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    int* w1s = reinterpret_cast<int*>(smem + W1_OFF);
    int* w2s = reinterpret_cast<int*>(smem + W2_OFF);
    int* acts = reinterpret_cast<int*>(smem + ACT_OFF);
    int* hids = reinterpret_cast<int*>(smem + HID_OFF);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int i = tid; i < D * H / 4; i += NT)
    {
        const int4 a = __ldg(reinterpret_cast<const int4*>(p.w1) + i), b = __ldg(reinterpret_cast<const int4*>(p.w2) + i);
        reinterpret_cast<int4*>(w1s)[i] = make_int4(a.x >> 6, a.y >> 6, a.z >> 6, a.w >> 6);          // the weight image holds raw << 6
        reinterpret_cast<int4*>(w2s)[i] = make_int4(b.x >> 6, b.y >> 6, b.z >> 6, b.w >> 6);
    }
    __syncthreads();

    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
    {
        const int n0 = tile * TM;
        const int rows = min(TM, p.num_nodes - n0);

        // ---- message passing: a warp per destination node, lanes 0..24 hold four columns each ----
        for (int r = wid; r < TM; r += NT / 32)
        {
            int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
            if (r < rows && lane < D / 4)
            {
                const int v = n0 + r;
                const int e0 = __ldg(p.in_ptr + v), e1 = __ldg(p.in_ptr + v + 1);
                for (int e = e0; e < e1; e++)
                {
                    const int u = __ldg(p.src + e);
                    const int c = __ldg(p.code + e);
                    const short4 hu = __ldg(reinterpret_cast<const short4*>(p.h_in + (size_t)u * D) + lane);
                    const short4 ee = __ldg(reinterpret_cast<const short4*>(p.ee + c * D) + lane);
                    m0 += relu16(ee.x + hu.x); m1 += relu16(ee.y + hu.y); m2 += relu16(ee.z + hu.z); m3 += relu16(ee.w + hu.w);
                }
                const short4 hv = __ldg(reinterpret_cast<const short4*>(p.h_in + (size_t)v * D) + lane);
                m0 += hv.x; m1 += hv.y; m2 += hv.z; m3 += hv.w;
            }
            if (lane < D / 4)
            {
                *reinterpret_cast<int4*>(acts + r * D + 4 * lane) = make_int4(wrap16(m0), wrap16(m1), wrap16(m2), wrap16(m3));
            }
        }
        __syncthreads();

        // ---- z = W1 a + b1, relu: thread (og, ng) owns outputs 4 og .. 4 og + 3 of nodes 8 ng .. 8 ng + 7 ----
        if (tid < 300)
        {
            const int og = tid % 50, ng = tid / 50;
            int acc[8][4];
            const int4 b = __ldg(reinterpret_cast<const int4*>(p.b1) + og);
#pragma unroll
            for (int i = 0; i < 8; i++) { acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.z; acc[i][3] = b.w; }
#pragma unroll 1
            for (int k = 0; k < D; k += 4)
            {
                int4 a[8];
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const int4*>(acts + (8 * ng + i) * D + k);
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
                {
                    const int4 w = *reinterpret_cast<const int4*>(w1s + (k + kk) * H + 4 * og);
#pragma unroll
                    for (int i = 0; i < 8; i++)
                    {
                        const int av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                        acc[i][0] = mac_floor(av, w.x, acc[i][0]); acc[i][1] = mac_floor(av, w.y, acc[i][1]);
                        acc[i][2] = mac_floor(av, w.z, acc[i][2]); acc[i][3] = mac_floor(av, w.w, acc[i][3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
                *reinterpret_cast<int4*>(hids + (8 * ng + i) * H + 4 * og) =
                    make_int4(relu16(acc[i][0]), relu16(acc[i][1]), relu16(acc[i][2]), relu16(acc[i][3]));
        }
        __syncthreads();

        // ---- h' = W2 relu(z) + b2 (+ relu): thread (og, ng) owns outputs 4 og .. 4 og + 3 of nodes 4 ng .. 4 ng + 3 ----
        if (tid < 300)
        {
            const int og = tid % 25, ng = tid / 25;
            int acc[4][4];
            const int4 b = __ldg(reinterpret_cast<const int4*>(p.b2) + og);
#pragma unroll
            for (int i = 0; i < 4; i++) { acc[i][0] = b.x; acc[i][1] = b.y; acc[i][2] = b.z; acc[i][3] = b.w; }
#pragma unroll 2
            for (int k = 0; k < H; k += 4)
            {
                int4 a[4];
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = *reinterpret_cast<const int4*>(hids + (4 * ng + i) * H + k);
#pragma unroll
                for (int kk = 0; kk < 4; kk++)
                {
                    const int4 w = *reinterpret_cast<const int4*>(w2s + (k + kk) * D + 4 * og);
#pragma unroll
                    for (int i = 0; i < 4; i++)
                    {
                        const int av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                        acc[i][0] = mac_floor(av, w.x, acc[i][0]); acc[i][1] = mac_floor(av, w.y, acc[i][1]);
                        acc[i][2] = mac_floor(av, w.z, acc[i][2]); acc[i][3] = mac_floor(av, w.w, acc[i][3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                const int r = 4 * ng + i;
                if (r < rows)
                {
                    short4 o;
                    if (p.relu_out) o = make_short4((short)relu16(acc[i][0]), (short)relu16(acc[i][1]), (short)relu16(acc[i][2]), (short)relu16(acc[i][3]));
                    else o = make_short4((short)acc[i][0], (short)acc[i][1], (short)acc[i][2], (short)acc[i][3]);
                    reinterpret_cast<short4*>(p.h_out + (size_t)(n0 + r) * D)[og] = o;
                }
            }
        }
        __syncthreads();
    }
}

// finalize: wrap-around sum over the graph's nodes, / num_of_nodes toward zero, Linear(100 -> 1)  (finalize.cc:36-115, linear.cc:11-49)
__global__ void __launch_bounds__(256) gin_fixed_pool_head_kernel(const int16_t* __restrict__ h, const int* __restrict__ node_off, const int* __restrict__ nn,
                                                                 const int* __restrict__ pred_w, const int* __restrict__ pred_b, float* __restrict__ out,
                                                                 int num_graphs)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int g = warp; g < num_graphs; g += nwarps)
    {
        const int n = __ldg(nn + g);
        const size_t base = (size_t)__ldg(node_off + g);
        int part = 0;
        if (lane < D / 4)
        {
            int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (int r = 0; r < n; r++)
            {
                const short4 x = __ldg(reinterpret_cast<const short4*>(h + (base + r) * D) + lane);
                s0 += x.x; s1 += x.y; s2 += x.z; s3 += x.w;
            }
            // an empty graph divides by zero in the reference; the oracle's emulation defines the quotient as 0
            const int q0 = n ? wrap16(s0) / n : 0, q1 = n ? wrap16(s1) / n : 0, q2 = n ? wrap16(s2) / n : 0, q3 = n ? wrap16(s3) / n : 0;
            const int4 w = __ldg(reinterpret_cast<const int4*>(pred_w) + lane);
            part = mad_hi(q0 << 16, w.x, 0) + mad_hi(q1 << 16, w.y, 0) + mad_hi(q2 << 16, w.z, 0) + mad_hi(q3 << 16, w.w, 0);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) out[g] = (float)wrap16(part + __ldg(pred_b)) * (1.0f / 1024.0f);
    }
}

}  // namespace

int gin_fixed_forward(DeviceBatch& b, const GinWeights& w, int sm_count, cudaStream_t s, int* launches)
{
    const long N = b.total_nodes;
    const int G = b.num_graphs;
    if (G == 0) return 0;
    FG_TRY(opt_in_smem(reinterpret_cast<const void*>(&gin_fixed_layer_kernel), SMEM_BYTES));
    FG_TRY(b.act[0].reserve(sizeof(int16_t) * (size_t)std::max<long>(N, 1) * D));
    FG_TRY(b.act[1].reserve(sizeof(int16_t) * (size_t)std::max<long>(N, 1) * D));
    int16_t* cur = b.act[0].as<int16_t>();
    int16_t* nxt = b.act[1].as<int16_t>();
    if (N > 0)
    {
        FixedEmbedOffsets off;
        const int concat[ND_FEATURE] = {0, 119, 123, 135, 147, 157, 163, 169, 171};     // GIN/src/load_inputs.cc:5
        for (int f = 0; f < ND_FEATURE; f++) off.off[f] = concat[f];
        FG_TRY(fixed_embed_launch(b.node_feature.as<int>(), w.fx_ne.as<int16_t>(), off, cur, N, sm_count, s));
        (*launches)++;
        const int tiles = (int)ceil_div<long>(N, TM);
        for (int l = 0; l < 5; l++)
        {
            FixedLayerParams p;
            p.h_in = cur; p.h_out = nxt;
            p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.code = b.code.as<uint8_t>();
            p.ee = w.fx_ee.as<int16_t>() + (size_t)l * ED_COMBOS * D;
            p.w1 = w.fx_w1.as<int>() + (size_t)l * D * H; p.b1 = w.fx_b1.as<int>() + l * H;
            p.w2 = w.fx_w2.as<int>() + (size_t)l * H * D; p.b2 = w.fx_b2.as<int>() + l * D;
            p.num_nodes = (int)N; p.num_tiles = tiles; p.relu_out = (l != 4);
            gin_fixed_layer_kernel<<<std::min(tiles, sm_count), NT, SMEM_BYTES, s>>>(p);
            FG_CUDA(cudaGetLastError());
            (*launches)++;
            std::swap(cur, nxt);
        }
    }
    gin_fixed_pool_head_kernel<<<std::min(ceil_div(G, 8), sm_count * 8), 256, 0, s>>>(cur, b.node_off.as<int>(), b.nums_of_nodes.as<int>(), w.fx_pw.as<int>(),
                                                                                      w.fx_pb.as<int>(), b.out.as<float>(), G);
    FG_CUDA(cudaGetLastError());
    (*launches)++;
    return 0;
}

}  // namespace fg
