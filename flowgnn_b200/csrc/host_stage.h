// Host-side staging of the reference-layout inputs for the host-pointer entry points (api.cu::run_reference_entry).
//
// The reference's kernel ABI hands over int32 words for values that are tiny: nine categorical atom features per node
// (node_feature_t, GIN/src/dcl.h:62-64), two graph-local node ids per edge (edge_t, dcl.h:61; the reference's cap is
// MAX_NODE = 500) and three bond attributes per edge (edge_attr_t, dcl.h:65-67) -- 36 B per node and 20 B per edge over
// PCIe.  A small pool of host threads narrows them into a pinned block (u8 / u16 / u8: 9 B per node, 7 B per edge) while
// the previous chunk's kernels run; one copy moves the block and `unpack_inputs_kernel` (prep.cu) widens it into the
// int32 arrays every kernel downstream reads, so nothing else changes.  The caller's arrays may be ordinary pageable
// memory (what the reference's host allocates, common/includes/xcl2/xcl2.hpp:61-76): they are only read by the CPU.
// An array with a value outside the narrow range (negative, > 255 / > 65,535 -- an out-of-vocabulary input) is uploaded
// unchanged instead, so the device-side checks see exactly what the caller passed.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace fg {

// Persistent worker threads; run() hands out job indices 0..n-1, the calling thread works too.
class HostPool {
public:
    explicit HostPool(int threads);
    ~HostPool();
    HostPool(const HostPool&) = delete;
    HostPool& operator=(const HostPool&) = delete;
    int threads() const { return (int)workers_.size() + 1; }
    void start(int njobs, std::function<void(int)> fn);   // returns at once; workers begin
    bool run_one();                                        // the caller takes ONE job if there is one left
    void finish();                                         // the caller takes jobs too, then waits for the last one
    static int local_ranks();                              // LOCAL_WORLD_SIZE (one process per GPU under torchrun), else 1
    static int default_threads();                          // FLOWGNN_B200_HOST_THREADS, else min(12, 3/4 of the usable cores / ranks on this node)

private:
    void worker();
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, cv_done_;
    std::function<void(int)> fn_;
    std::atomic<int> njobs_{0}, next_{0};                  // written under mu_; read without it by workers in their spin phase
    int pending_ = 0;
    int spin_us_ = 0;                                      // FLOWGNN_B200_HOST_SPIN_US: how long an idle worker polls before it sleeps
    std::atomic<bool> stop_{false};
};

// dst[i] = (narrow)src[i]; returns the OR of all source words (bits outside the narrow type = value out of range)
uint32_t narrow_u8(const int32_t* src, uint8_t* dst, size_t n);
uint32_t narrow_u16(const int32_t* src, uint16_t* dst, size_t n);

// One chunk's narrowed inputs inside a pinned block: [feat u8 x 9N | pad16][edge u16 x 2E | pad16][attr u8 x 3E | pad16]
// [node_eigen f32 x 4N] (element counts are 0 for arrays that do not go through the block)
struct NarrowPlan {
    size_t n_feat = 0, n_edge = 0, n_attr = 0;             // element counts (9N, 2E, 3E or 0)
    size_t n_eig = 0;                                      // DGN node_eigen floats (4N or 0): copied as they are, so that a pageable array is read by the pool
    size_t off_feat = 0, off_edge = 0, off_attr = 0, off_eig = 0, bytes = 0;
    static size_t pad16(size_t b) { return (b + 15) & ~size_t(15); }
    // which[a]: array a (feat / edge_list / edge_attr / node_eigen) goes through the block; the others take no room
    void layout(size_t nodes, size_t edges, const bool which[4])
    {
        n_feat = which[0] ? 9 * nodes : 0; n_edge = which[1] ? 2 * edges : 0; n_attr = which[2] ? 3 * edges : 0; n_eig = which[3] ? 4 * nodes : 0;
        off_feat = 0;
        off_edge = pad16(n_feat);
        off_attr = off_edge + pad16(2 * n_edge);
        off_eig = off_attr + pad16(n_attr);
        bytes = off_eig + pad16(4 * n_eig);
    }
};

// One call's narrowing: the slices of ALL chunks are queued at once, in chunk order, so the pool works through them without
// pausing while the calling thread ships chunk after chunk as each one completes.
class NarrowRun {
public:
    static constexpr int MAX_CHUNKS = 16;
    struct Chunk {
        NarrowPlan plan;
        size_t base = 0;                                    // byte offset of the chunk's block inside the run's pinned block
        const int32_t* src[3] = {nullptr, nullptr, nullptr}; // feat / edge_list / edge_attr of this chunk (nullptr: not narrowed)
        const float* eig = nullptr;                         // node_eigen of this chunk (nullptr: not staged)
    };
    // total bytes of the pinned block for these chunks (fills Chunk::base)
    static size_t layout(Chunk* chunks, int n);
    // `first`: an extra job per chunk, queued ahead of the chunk's slices (the entry points compute the chunk's tile packing there)
    void start(HostPool& pool, const Chunk* chunks, int n, uint8_t* block, std::function<void(int)> first = nullptr);
    // the caller works on slices too until chunk `ci` is complete; ok[a] = every value of array a fits the narrow type
    void wait_chunk(HostPool& pool, int ci, bool ok[3]);
    void finish(HostPool& pool);                            // join everything (also on error paths)
    bool active() const { return active_; }
    const Chunk& chunk(int ci) const { return chunks_[ci]; }

private:
    static constexpr size_t SLICE = 64 * 1024;             // source words per job (256 KB read)
    struct Job { int chunk, array; size_t i0, len; };      // array -1: the chunk's `first` job
    std::function<void(int)> first_;
    std::vector<Job> jobs_;
    Chunk chunks_[MAX_CHUNKS];
    int n_ = 0;
    uint8_t* block_ = nullptr;
    std::atomic<int> remaining_[MAX_CHUNKS];
    std::atomic<uint32_t> seen_[MAX_CHUNKS][3];
    bool active_ = false;
};

}  // namespace fg
