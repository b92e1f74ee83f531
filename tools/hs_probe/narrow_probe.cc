// Host-side narrowing throughput (flowgnn_b200/csrc/host_stage.h) by thread count: narrow_probe <threads>
// Three chunks graded 1 : 3 : 6 as in the entry points (node_feature, edge_list, edge_attr narrowed, node_eigen copied), every value checked.
#include "../../flowgnn_b200/csrc/host_stage.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
using namespace fg;
int main(int argc, char** argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const size_t N = 1034135, E = 2253418;                  // the bench batch: 82.6 MB of int32 words
    std::vector<int32_t> f(9 * N), e(2 * E), a(3 * E);
    std::vector<float> g(4 * N);
    for (size_t i = 0; i < f.size(); i++) f[i] = (int32_t)((i * 7) % 119);
    for (size_t i = 0; i < e.size(); i++) e[i] = (int32_t)((i * 13) % 60000);
    for (size_t i = 0; i < a.size(); i++) a[i] = (int32_t)(i % 6);
    for (size_t i = 0; i < g.size(); i++) g[i] = (float)i * 0.5f;
    const bool which[4] = {true, true, true, true};
    const size_t nb[4] = {0, N / 10, N / 10 * 4, N}, eb[4] = {0, E / 10, E / 10 * 4, E};
    NarrowRun::Chunk ch[3];
    for (int c = 0; c < 3; c++)
    {
        ch[c].plan.layout(nb[c + 1] - nb[c], eb[c + 1] - eb[c], which);
        ch[c].src[0] = f.data() + 9 * nb[c]; ch[c].src[1] = e.data() + 2 * eb[c]; ch[c].src[2] = a.data() + 3 * eb[c];
        ch[c].eig = g.data() + 4 * nb[c];
    }
    const size_t bytes = NarrowRun::layout(ch, 3);
    uint8_t* blk = static_cast<uint8_t*>(aligned_alloc(4096, (bytes + 4095) & ~size_t(4095)));
    HostPool pool(T);
    NarrowRun run;
    int firsts = 0;
    double best = 1e9;
    for (int r = 0; r < 8; r++)
    {
        const auto t0 = std::chrono::steady_clock::now();
        run.start(pool, ch, 3, blk, [&](int) { __atomic_fetch_add(&firsts, 1, __ATOMIC_RELAXED); });
        for (int c = 0; c < 3; c++)
        {
            bool ok[3];
            run.wait_chunk(pool, c, ok);
            if (!(ok[0] && ok[1] && ok[2])) { puts("range flags wrong"); return 1; }
        }
        run.finish(pool);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < best) best = ms;
    }
    if (firsts != 8 * 3) { puts("per-chunk first jobs did not all run"); return 1; }
    for (int c = 0; c < 3; c++)
    {
        const uint8_t* base = blk + run.chunk(c).base;
        const NarrowPlan& p = run.chunk(c).plan;
        for (size_t i = 0; i < p.n_feat; i++) if (base[p.off_feat + i] != (uint8_t)f[9 * nb[c] + i]) { puts("BAD feat"); return 1; }
        for (size_t i = 0; i < p.n_edge; i++) if (((const uint16_t*)(base + p.off_edge))[i] != (uint16_t)e[2 * eb[c] + i]) { puts("BAD edge"); return 1; }
        for (size_t i = 0; i < p.n_attr; i++) if (base[p.off_attr + i] != (uint8_t)a[3 * eb[c] + i]) { puts("BAD attr"); return 1; }
        if (std::memcmp(base + p.off_eig, g.data() + 4 * nb[c], 4 * p.n_eig) != 0) { puts("BAD eigen"); return 1; }
    }
    // a value that does not fit must clear its array's flag, and only that one
    e[2 * eb[1] + 5] = 70000;
    run.start(pool, ch, 3, blk);
    for (int c = 0; c < 3; c++)
    {
        bool ok[3];
        run.wait_chunk(pool, c, ok);
        if (ok[0] != true || ok[2] != true || ok[1] != (c != 1)) { puts("out-of-range value not reported for its chunk / array"); return 1; }
    }
    run.finish(pool);
    const double mb = 4e-6 * (f.size() + e.size() + a.size());
    printf("threads %2d: %.3f ms for %.1f MB of int32 words (+ %.1f MB of eigenvectors copied) = %.1f GB/s read (avx2 path %s)\n", T, best, mb, 16e-6 * N,
           (mb + 16e-6 * N) / best, getenv("FLOWGNN_B200_NO_AVX2") ? "off" : "on if the CPU has it");
    return 0;
}
