// Host thread pool + int32 -> u8 / u16 narrowing of the reference-layout inputs (see host_stage.h).  Plain C++, no CUDA.
#include "host_stage.h"

#include <sched.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

namespace fg {

int HostPool::local_ranks()
{
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) return std::max(1, std::atoi(e));
    return 1;
}

int HostPool::default_threads()
{
    if (const char* e = std::getenv("FLOWGNN_B200_HOST_THREADS")) return std::max(1, std::min(64, std::atoi(e)));
    int cores = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
    const int ranks = local_ranks();                        // one process per GPU (torchrun): share the cores
    return std::max(1, std::min(12, cores * 3 / 4 / ranks));
}

HostPool::HostPool(int threads)
{
    if (const char* e = std::getenv("FLOWGNN_B200_HOST_SPIN_US")) spin_us_ = std::max(0, std::min(100000, std::atoi(e)));
    for (int i = 1; i < threads; i++) workers_.emplace_back([this] { worker(); });
}

HostPool::~HostPool()
{
    {
        std::lock_guard<std::mutex> lock(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (std::thread& t : workers_) t.join();
}

// FLOWGNN_B200_HOST_SPIN_US > 0: an idle worker polls for work that long before it sleeps on the condition variable, so that calls
// following each other within milliseconds find the pool awake.  Off by default: measured on the 41k-graph GIN batch (2 ms of polling
// against none, 30 calls each, twice) 2.49 / 2.50 ms against 2.51 / 2.56 ms per call -- inside the run-to-run spread, not worth the cores.
void HostPool::worker()
{
    std::unique_lock<std::mutex> lock(mu_);
    for (;;)
    {
        if (!(stop_ || next_.load(std::memory_order_relaxed) < njobs_.load(std::memory_order_relaxed)) && spin_us_ > 0)
        {
            lock.unlock();
            const auto t0 = std::chrono::steady_clock::now();
            for (int k = 0;; k++)
            {
                if (stop_.load(std::memory_order_relaxed) || next_.load(std::memory_order_relaxed) < njobs_.load(std::memory_order_relaxed)) break;
#if defined(__x86_64__)
                _mm_pause();
#endif
                if ((k & 63) == 63 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us_)) break;
            }
            lock.lock();
        }
        cv_.wait(lock, [this] { return stop_ || next_ < njobs_; });
        if (stop_) return;
        const int j = next_++;
        lock.unlock();
        fn_(j);
        lock.lock();
        if (--pending_ == 0) cv_done_.notify_all();
    }
}

void HostPool::start(int njobs, std::function<void(int)> fn)
{
    std::lock_guard<std::mutex> lock(mu_);
    fn_ = std::move(fn);
    njobs_ = njobs; next_ = 0; pending_ = njobs;
    if (njobs > 0) cv_.notify_all();
}

bool HostPool::run_one()
{
    std::unique_lock<std::mutex> lock(mu_);
    if (next_ >= njobs_) return false;
    const int j = next_++;
    lock.unlock();
    fn_(j);
    lock.lock();
    if (--pending_ == 0) cv_done_.notify_all();
    return true;
}

void HostPool::finish()
{
    std::unique_lock<std::mutex> lock(mu_);
    while (next_ < njobs_)
    {
        const int j = next_++;
        lock.unlock();
        fn_(j);
        lock.lock();
        --pending_;
    }
    cv_done_.wait(lock, [this] { return pending_ == 0; });
    njobs_ = 0; next_ = 0;
}

// Portable loops (gcc vectorises them: mask + pack) and, on x86-64 with AVX2, explicit versions with non-temporal stores: the
// narrow block is written once and read next by the DMA engine, so it need not be read into the cache first (a sixth less memory
// traffic where the narrowing is what bounds the call, e.g. 785 MB of hep10k inputs per step).
static uint32_t narrow_u8_plain(const int32_t* __restrict src, uint8_t* __restrict dst, size_t n)
{
    uint32_t seen = 0;
    for (size_t i = 0; i < n; i++)
    {
        const uint32_t v = (uint32_t)src[i];
        seen |= v;
        dst[i] = (uint8_t)v;
    }
    return seen;
}

static uint32_t narrow_u16_plain(const int32_t* __restrict src, uint16_t* __restrict dst, size_t n)
{
    uint32_t seen = 0;
    for (size_t i = 0; i < n; i++)
    {
        const uint32_t v = (uint32_t)src[i];
        seen |= v;
        dst[i] = (uint16_t)v;
    }
    return seen;
}

#if defined(__x86_64__) && defined(__GNUC__)
#define FG_HAVE_AVX2_PATH 1
__attribute__((target("avx2"))) static uint32_t or_lanes(__m256i v)
{
    alignas(32) uint32_t w[8];
    _mm256_store_si256(reinterpret_cast<__m256i*>(w), v);
    return w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7];
}

__attribute__((target("avx2"))) static uint32_t narrow_u8_avx2(const int32_t* src, uint8_t* dst, size_t n)
{
    size_t i = 0;
    uint32_t seen = 0;
    for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31); i++) { const uint32_t v = (uint32_t)src[i]; seen |= v; dst[i] = (uint8_t)v; }
    __m256i acc = _mm256_setzero_si256();
    const __m256i mask = _mm256_set1_epi32(0xFF), order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    for (; i + 32 <= n; i += 32)
    {
        __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)), b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 8));
        __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 16)), d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 24));
        acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, d)));
        // per 128-bit half: a0-3 b0-3 | a4-7 b4-7, then a0-3 b0-3 c0-3 d0-3 | a4-7 b4-7 c4-7 d4-7; the dword permute restores the order
        const __m256i ab = _mm256_packus_epi32(_mm256_and_si256(a, mask), _mm256_and_si256(b, mask));
        const __m256i cd = _mm256_packus_epi32(_mm256_and_si256(c, mask), _mm256_and_si256(d, mask));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_permutevar8x32_epi32(_mm256_packus_epi16(ab, cd), order));
    }
    _mm_sfence();
    seen |= or_lanes(acc);
    for (; i < n; i++) { const uint32_t v = (uint32_t)src[i]; seen |= v; dst[i] = (uint8_t)v; }
    return seen;
}

__attribute__((target("avx2"))) static uint32_t narrow_u16_avx2(const int32_t* src, uint16_t* dst, size_t n)
{
    size_t i = 0;
    uint32_t seen = 0;
    for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31); i++) { const uint32_t v = (uint32_t)src[i]; seen |= v; dst[i] = (uint16_t)v; }
    __m256i acc = _mm256_setzero_si256();
    const __m256i mask = _mm256_set1_epi32(0xFFFF);
    for (; i + 16 <= n; i += 16)
    {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)), b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 8));
        acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
        const __m256i ab = _mm256_packus_epi32(_mm256_and_si256(a, mask), _mm256_and_si256(b, mask));      // a0-3 b0-3 | a4-7 b4-7
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_permute4x64_epi64(ab, 0xD8));
    }
    _mm_sfence();
    seen |= or_lanes(acc);
    for (; i < n; i++) { const uint32_t v = (uint32_t)src[i]; seen |= v; dst[i] = (uint16_t)v; }
    return seen;
}

static bool have_avx2()
{
    static const bool yes = __builtin_cpu_supports("avx2") && !std::getenv("FLOWGNN_B200_NO_AVX2");
    return yes;
}
#endif

uint32_t narrow_u8(const int32_t* src, uint8_t* dst, size_t n)
{
#ifdef FG_HAVE_AVX2_PATH
    if (have_avx2()) return narrow_u8_avx2(src, dst, n);
#endif
    return narrow_u8_plain(src, dst, n);
}

uint32_t narrow_u16(const int32_t* src, uint16_t* dst, size_t n)
{
#ifdef FG_HAVE_AVX2_PATH
    if (have_avx2()) return narrow_u16_avx2(src, dst, n);
#endif
    return narrow_u16_plain(src, dst, n);
}

size_t NarrowRun::layout(Chunk* chunks, int n)
{
    size_t at = 0;
    for (int c = 0; c < n; c++)
    {
        chunks[c].base = at;
        at += chunks[c].plan.bytes;
    }
    return at;
}

void NarrowRun::start(HostPool& pool, const Chunk* chunks, int n, uint8_t* block, std::function<void(int)> first)
{
    n_ = n; block_ = block;
    first_ = std::move(first);
    jobs_.clear();
    for (int c = 0; c < n; c++)
    {
        chunks_[c] = chunks[c];
        const size_t counts[3] = {chunks[c].plan.n_feat, chunks[c].plan.n_edge, chunks[c].plan.n_attr};
        int jobs = 0;
        if (first_) { jobs_.push_back({c, -1, 0, 0}); jobs++; }
        for (int a = 0; a < 3; a++)
        {
            seen_[c][a].store(0u, std::memory_order_relaxed);
            if (!chunks[c].src[a]) continue;
            for (size_t i0 = 0; i0 < counts[a]; i0 += SLICE) { jobs_.push_back({c, a, i0, std::min(SLICE, counts[a] - i0)}); jobs++; }
        }
        if (chunks[c].eig)
            for (size_t i0 = 0; i0 < chunks[c].plan.n_eig; i0 += SLICE) { jobs_.push_back({c, 3, i0, std::min(SLICE, chunks[c].plan.n_eig - i0)}); jobs++; }
        remaining_[c].store(jobs, std::memory_order_relaxed);
    }
    active_ = true;
    pool.start((int)jobs_.size(), [this](int j) {
        const Job& job = jobs_[(size_t)j];
        if (job.array < 0) { first_(job.chunk); remaining_[job.chunk].fetch_sub(1, std::memory_order_release); return; }
        const Chunk& ch = chunks_[job.chunk];
        uint8_t* base = block_ + ch.base;
        if (job.array == 3)
        {
            std::memcpy(base + ch.plan.off_eig + 4 * job.i0, ch.eig + job.i0, 4 * job.len);
            remaining_[job.chunk].fetch_sub(1, std::memory_order_release);
            return;
        }
        uint32_t seen;
        if (job.array == 0) seen = narrow_u8(ch.src[0] + job.i0, base + ch.plan.off_feat + job.i0, job.len);
        else if (job.array == 1) seen = narrow_u16(ch.src[1] + job.i0, reinterpret_cast<uint16_t*>(base + ch.plan.off_edge) + job.i0, job.len);
        else seen = narrow_u8(ch.src[2] + job.i0, base + ch.plan.off_attr + job.i0, job.len);
        seen_[job.chunk][job.array].fetch_or(seen, std::memory_order_relaxed);
        remaining_[job.chunk].fetch_sub(1, std::memory_order_release);
    });
}

void NarrowRun::wait_chunk(HostPool& pool, int ci, bool ok[3])
{
    while (remaining_[ci].load(std::memory_order_acquire) > 0)
        if (!pool.run_one()) std::this_thread::yield();     // the last slices are in other threads' hands
    ok[0] = chunks_[ci].src[0] && (seen_[ci][0].load(std::memory_order_relaxed) & ~0xFFu) == 0;
    ok[1] = chunks_[ci].src[1] && (seen_[ci][1].load(std::memory_order_relaxed) & ~0xFFFFu) == 0;
    ok[2] = chunks_[ci].src[2] && (seen_[ci][2].load(std::memory_order_relaxed) & ~0xFFu) == 0;
}

void NarrowRun::finish(HostPool& pool)
{
    if (!active_) return;
    pool.finish();
    active_ = false;
}

}  // namespace fg
