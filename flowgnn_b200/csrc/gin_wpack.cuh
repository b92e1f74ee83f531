// Weight image of the GIN CTA-pair kernels (gin_tc2.cu, gin_fused.cu): per layer and cluster rank, half of the N rows of
// W1 / W2 as bf16 hi / lo blocks in the tcgen05 no-swizzle K-major canonical layout, biases as an extra K column.
// Packed on the host by gin_tc2_pack_layer (gin_tc2.cu), one bulk-TMA load per launch.
#pragma once

namespace fg {
namespace ginw {

constexpr int D = 100;
constexpr int Q = D / 4;
constexpr int TM = 128;                       // nodes per CTA tile (pair tile = 256)
constexpr int N1A = 112, N1B = 96, N1 = N1A + N1B;   // GEMM1 N halves (z columns), whole pair
// Experiment (-DFG_G1_SINGLE=1): GEMM1 as ONE N = 208 MMA per k-step and product instead of two N halves -- the A tile is
// read from shared memory 21 times per tile instead of 42 (86 KB less operand traffic), but the conversion of the first
// z half no longer overlaps the second half of GEMM1.  Measured: 0.467 ms per layer against 0.39 ms, so the split stays.
#ifndef FG_G1_SINGLE
#define FG_G1_SINGLE 0
#endif
constexpr bool G1_SINGLE = FG_G1_SINGLE != 0;
constexpr int LBO_W1 = (N1 / 2) * 16;           // single block: CTA r holds z columns 104 r .. 104 r + 103
constexpr int N2 = 128;                       // GEMM2 N (100 used)
constexpr int K1_STEPS = 7, K2_STEPS = 13;    // K = 16 per step
constexpr int K1_CHUNKS = 13;                 // stored 8-element K chunks of A and W1 (k < 104; chunk 13 reads as zero)
constexpr int K2_CHUNKS = 26;                 // stored K chunks of W2: k < 200 plus the bias column k = 200 (z column 200 == 1)

// per-CTA weight image (bytes): half of the N rows of every block
constexpr int LBO_W1A = (N1A / 2) * 16, LBO_W1B = (N1B / 2) * 16, LBO_W2 = (N2 / 2) * 16;
constexpr int W1A_BYTES = LBO_W1A * K1_CHUNKS, W1B_BYTES = LBO_W1B * K1_CHUNKS, W2_BYTES = LBO_W2 * K2_CHUNKS;
constexpr int OFF_W1A_HI = 0, OFF_W1A_LO = OFF_W1A_HI + W1A_BYTES, OFF_W1B_HI = OFF_W1A_LO + W1A_BYTES, OFF_W1B_LO = OFF_W1B_HI + W1B_BYTES,
              OFF_W2_HI = OFF_W1B_LO + W1B_BYTES, OFF_W2_LO = OFF_W2_HI + W2_BYTES;
constexpr int W_BYTES = OFF_W2_LO + W2_BYTES;           // 94,464


}  // namespace ginw
}  // namespace fg
