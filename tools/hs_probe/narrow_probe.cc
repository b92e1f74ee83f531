// Host-side narrowing throughput (flowgnn_b200/csrc/host_stage.h) by thread count: narrow_probe <threads>
#include "../../flowgnn_b200/csrc/host_stage.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
using namespace fg;
int main(int argc, char** argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const size_t N = 1034135, E = 2253418;                  // the bench batch: 82.6 MB of int32 words
    std::vector<int32_t> f(9 * N), e(2 * E), a(3 * E);
    for (size_t i = 0; i < f.size(); i++) f[i] = i % 119;
    for (size_t i = 0; i < e.size(); i++) e[i] = i % 60;
    for (size_t i = 0; i < a.size(); i++) a[i] = i % 2;
    const bool which[4] = {true, true, true, false};
    NarrowRun::Chunk ch;
    ch.plan.layout(N, E, which);
    ch.src[0] = f.data(); ch.src[1] = e.data(); ch.src[2] = a.data();
    const size_t bytes = NarrowRun::layout(&ch, 1);
    uint8_t* blk = static_cast<uint8_t*>(aligned_alloc(4096, (bytes + 4095) & ~size_t(4095)));
    HostPool pool(T);
    NarrowRun run;
    double best = 1e9;
    for (int r = 0; r < 8; r++)
    {
        const auto t0 = std::chrono::steady_clock::now();
        run.start(pool, &ch, 1, blk);
        bool ok[3];
        run.wait_chunk(pool, 0, ok);
        run.finish(pool);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < best) best = ms;
        if (!(ok[0] && ok[1] && ok[2])) { puts("range flags wrong"); return 1; }
    }
    for (size_t i = 0; i < f.size(); i++) if (blk[ch.plan.off_feat + i] != (uint8_t)f[i]) { puts("BAD feat"); return 1; }
    for (size_t i = 0; i < e.size(); i++) if (((uint16_t*)(blk + ch.plan.off_edge))[i] != (uint16_t)e[i]) { puts("BAD edge"); return 1; }
    for (size_t i = 0; i < a.size(); i++) if (blk[ch.plan.off_attr + i] != (uint8_t)a[i]) { puts("BAD attr"); return 1; }
    printf("threads %2d: %.3f ms for %.1f MB of int32 words = %.1f GB/s read (avx2 path %s)\n", T, best, 4e-6 * (f.size() + e.size() + a.size()),
           4e-6 * (f.size() + e.size() + a.size()) / best, getenv("FLOWGNN_B200_NO_AVX2") ? "off" : "on if the CPU has it");
    return 0;
}
