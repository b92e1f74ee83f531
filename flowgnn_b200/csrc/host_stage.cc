// Host thread pool + int32 -> u8 / u16 narrowing of the reference-layout inputs (see host_stage.h).  Plain C++, no CUDA.
#include "host_stage.h"

#include <sched.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace fg {

int HostPool::default_threads()
{
    if (const char* e = std::getenv("FLOWGNN_B200_HOST_THREADS")) return std::max(1, std::min(64, std::atoi(e)));
    int cores = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
    int ranks = 1;                                          // one process per GPU (torchrun): share the cores
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, std::atoi(e));
    return std::max(1, std::min(12, cores * 3 / 4 / ranks));
}

HostPool::HostPool(int threads)
{
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

void HostPool::worker()
{
    std::unique_lock<std::mutex> lock(mu_);
    for (;;)
    {
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

// The loops are written so that gcc vectorises them (mask + pack); the OR keeps the range check off the critical path.
uint32_t narrow_u8(const int32_t* __restrict src, uint8_t* __restrict dst, size_t n)
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

uint32_t narrow_u16(const int32_t* __restrict src, uint16_t* __restrict dst, size_t n)
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
