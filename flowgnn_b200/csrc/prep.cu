// On-device `load_graph`: COO edge list -> per-batch CSR by destination.
//
// Reference: GIN/src/load_inputs.cc:87-172 (and GCN :77-166, PNA :48-131, DGN :28-112, GAT :87-166).
// The FPGA builds, per graph, four banked neighbour tables ordered by source node; a destination's
// messages therefore arrive in (source ascending, edge-list order).  Here the whole batch gets ONE
// CSR keyed by GLOBAL destination node id whose rows keep exactly that order, so the gather in the
// layer kernels adds a node's in-edges in the reference's order and is deterministic (no float
// atomics anywhere).  Derived per-edge / per-node tables follow the per-model variants:
//   GCN  norm = dis[u]*dis[v], dis = 1/sqrt(outdeg+1) for nodes that appear as a source, else 0
//        (GCN/src/load_inputs.cc:100-122,163)
//   DGN  eig_w = phi_u - phi_v with phi = eig[:,1]; per destination sum|eig_w| and sum eig_w
//        (DGN/src/load_inputs.cc:91-111)
//   all  out-degree per node (degree_table)
#include "internal.cuh"

#include <algorithm>

namespace fg {

namespace {

constexpr int SCAN_THREADS = 1024;

// Exclusive prefix sums of nums_of_nodes / nums_of_edges (the reference carries them as running
// offsets in its serial graph loop, GIN/src/GIN_compute.cc:44,96-97).  One block of 32 warps; every warp owns
// a contiguous segment of graphs: pass 1 sums it (coalesced, independent loads), the 32 segment totals are
// scanned, pass 2 re-reads the segment and writes warp-scanned offsets with a running carry.
__device__ __forceinline__ void scan_offsets_body(const int* __restrict__ nn, const int* __restrict__ ne, int* __restrict__ node_off,
                                                  int* __restrict__ edge_off, int num_graphs)
{
    __shared__ int2 seg_tot[SCAN_THREADS / 32];
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = ((num_graphs + SCAN_THREADS - 1) / SCAN_THREADS) * 32;      // graphs per warp, multiple of 32
    const int g0 = min(wid * per, num_graphs), g1 = min(g0 + per, num_graphs);
    int2 x = make_int2(0, 0);
#pragma unroll 4
    for (int g = g0 + lane; g < g1; g += 32) { x.x += __ldg(nn + g); x.y += __ldg(ne + g); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { x.x += __shfl_xor_sync(full, x.x, d); x.y += __shfl_xor_sync(full, x.y, d); }
    if (lane == 0) seg_tot[wid] = x;
    __syncthreads();
    if (wid == 0)
    {
        const int2 t = seg_tot[lane];
        int2 ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            int a = __shfl_up_sync(full, ti.x, d), b = __shfl_up_sync(full, ti.y, d);
            if (lane >= d) { ti.x += a; ti.y += b; }
        }
        seg_tot[lane] = make_int2(ti.x - t.x, ti.y - t.y);      // exclusive over segments
    }
    __syncthreads();
    int2 carry = seg_tot[wid];
#pragma unroll 4
    for (int base = g0; base < g1; base += 32)
    {
        const int g = base + lane;
        const int2 v = (g < g1) ? make_int2(__ldg(nn + g), __ldg(ne + g)) : make_int2(0, 0);
        int2 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            int a = __shfl_up_sync(full, incl.x, d), b = __shfl_up_sync(full, incl.y, d);
            if (lane >= d) { incl.x += a; incl.y += b; }
        }
        if (g < g1)
        {
            node_off[g] = carry.x + incl.x - v.x;
            edge_off[g] = carry.y + incl.y - v.y;
        }
        carry.x += __shfl_sync(full, incl.x, 31);
        carry.y += __shfl_sync(full, incl.y, 31);
    }
    if (g1 == num_graphs && g0 < g1 && lane == 0)
    {
        node_off[num_graphs] = carry.x;
        edge_off[num_graphs] = carry.y;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_offsets_kernel(const int* __restrict__ nn, const int* __restrict__ ne,
                                                                    int* __restrict__ node_off, int* __restrict__ edge_off,
                                                                    int num_graphs)
{
    scan_offsets_body(nn, ne, node_off, edge_off, num_graphs);
}

// Row descriptor of a node for the GIN tensor-core kernel (gin_tc2.cu): its first four in-edges, each packed as
// (source - node + 32768) | code << 16 (absent slots: the node itself, sentinel code 60), in-degree (capped at 255)
// in bits 24..31 of .x.  One 16-byte load replaces the in_ptr -> src/code pointer chase of the gather.
__device__ __forceinline__ int4 empty_row_desc()
{
    const int e = 32768 | (ED_COMBOS << 16);
    return make_int4(e, e, e, e);
}

constexpr int CSR_NCAP = 1024;     // nodes per graph supported by the warp-local shared-memory tables (reference cap: 500);
                                   // larger graphs use tables in global memory (big_tab), the same code, slower
constexpr int CSR_NCAP_SMALL = 128;
// The build is a chain of dependent memory round trips per graph (edges -> histogram -> sort scratch -> payload ->
// descriptors), so it lives on occupancy.  Two instantiations run back to back: graphs of up to 128 nodes (every
// molecule) take the small tables -- 1.5 KB of shared memory per warp, 64 warps per SM -- the others the 1,024-node
// tables (12 KB per warp, 16 warps per SM); each kernel skips the graphs of the other.

struct CsrParams {
    const int* nn; const int* ne; const int* node_off; const int* edge_off;      // where the graph's rows / in-edges are WRITTEN
    const int* node_in_off; const int* edge_in_off;                                // where its inputs are READ (caller order); the same arrays
                                                                                   // unless the graphs are re-ordered for tile packing (api.cu)
    int* node_map;                   // [N] caller-order node index of every (re-ordered) row, or nullptr
    const int* edge_list; const int* edge_attr; const float* node_eigen;
    int* in_ptr; int* src; uint8_t* code; float* edge_w; int* out_deg; float* node_w0; float* node_w1; int4* row_desc;
    int* sort_tmp; int* status;
    int* big_tab;                    // [3][N] histogram / cursor tables of graphs above CSR_NCAP nodes (nullptr: the batch has none)
    long total_nodes;
    int num_graphs; int flags; int has_attr;
};

// A graph the build rejects (status bit set; the call returns an error after the forward): every row empty, and the
// graph's edge slots -- which the row pointers of the neighbouring rows still span -- filled with defined values, so
// that the layer kernels that run before the status is read stay in bounds (sources = the graph's first node, or node 0
// for a graph without nodes; code 0; weight 0).
__device__ __forceinline__ void neutralise_graph(const CsrParams& p, int nb, int n, int eb, int e)
{
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < n; i += 32)
    {
        p.in_ptr[nb + i] = eb; p.out_deg[nb + i] = 0;
        if (p.row_desc) p.row_desc[nb + i] = empty_row_desc();
        if (p.flags & PREP_DGN_EIG) { p.node_w0[nb + i] = 0.f; p.node_w1[nb + i] = 0.f; }
    }
    const int safe = n > 0 ? nb : 0;
    for (int i = lane; i < e; i += 32)
    {
        p.src[eb + i] = safe;
        if (p.has_attr) p.code[eb + i] = 0;
        if (p.flags & (PREP_GCN_NORM | PREP_DGN_EIG)) p.edge_w[eb + i] = 0.f;
    }
}

// One warp per graph.  Two stable counting-sort passes (by source, then by destination) give the
// (destination, source, list-order) ordering; ranks inside a 32-edge chunk come from
// __match_any_sync, chunks are consumed in order, so the sort is stable and deterministic.
__device__ __forceinline__ void build_graph_csr(const CsrParams& p, int g, int ncap, int* deg, int* pu, int* pv)
{
    const int lane = threadIdx.x & 31;
    const int n = p.nn[g], e = p.ne[g];
    const int nb = p.node_off[g], eb = p.edge_off[g];
    const int nb_in = p.node_in_off[g], eb_in = p.edge_in_off[g];
    if (nb + n == p.total_nodes && n > 0 && lane == 0) p.in_ptr[nb + n] = eb + e;      // the graph that ends the (possibly re-ordered) batch
    if (p.node_map)
        for (int i = lane; i < n; i += 32) p.node_map[nb + i] = nb_in + i;
    if (n > ncap || n < 0 || e < 0)
    {
        if (lane == 0) atomicOr(p.status, 1);
        neutralise_graph(p, nb, n, eb, e);
        return;
    }
    const int2* edges = reinterpret_cast<const int2*>(p.edge_list) + eb_in;
    const unsigned full = 0xffffffffu;

    for (int i = lane; i < n; i += 32) { pu[i] = 0; pv[i] = 0; }
    __syncwarp();
    bool bad = false;
    for (int i = lane; i < e; i += 32)
    {
        const int2 uv = __ldg(edges + i);
        if ((unsigned)uv.x >= (unsigned)n || (unsigned)uv.y >= (unsigned)n) { bad = true; continue; }
        atomicAdd(&pu[uv.x], 1);
        atomicAdd(&pv[uv.y], 1);
    }
    if (__any_sync(full, bad))
    {
        if (lane == 0) atomicOr(p.status, 2);
        neutralise_graph(p, nb, n, eb, e);
        return;
    }
    __syncwarp();
    // exclusive scans of both histograms, 32 nodes at a time
    int carry_u = 0, carry_v = 0;
    for (int base = 0; base < n; base += 32)
    {
        const int i = base + lane;
        const int cu = (i < n) ? pu[i] : 0, cv = (i < n) ? pv[i] : 0;
        int su = cu, sv = cv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            int a = __shfl_up_sync(full, su, d), b = __shfl_up_sync(full, sv, d);
            if (lane >= d) { su += a; sv += b; }
        }
        if (i < n)
        {
            deg[i] = cu;
            pu[i] = carry_u + su - cu;
            pv[i] = carry_v + sv - cv;
            p.out_deg[nb + i] = cu;
            p.in_ptr[nb + i] = eb + carry_v + sv - cv;
        }
        carry_u += __shfl_sync(full, su, 31);
        carry_v += __shfl_sync(full, sv, 31);
    }
    __syncwarp();

    int* tmp = p.sort_tmp + eb;
    // pass 1: stable by source
    for (int base = 0; base < e; base += 32)
    {
        const int i = base + lane;
        const bool act = i < e;
        const unsigned mask = __ballot_sync(full, act);
        if (act)
        {
            const int u = __ldg(edges + i).x;
            const unsigned peers = __match_any_sync(mask, u);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            tmp[pu[u] + rank] = i;
            __syncwarp(mask);
            if (rank == __popc(peers) - 1) pu[u] += rank + 1;
        }
        __syncwarp();
    }
    // pass 2: stable by destination, emit the CSR payload
    for (int base = 0; base < e; base += 32)
    {
        const int k = base + lane;
        const bool act = k < e;
        const unsigned mask = __ballot_sync(full, act);
        if (act)
        {
            const int i = tmp[k];
            const int2 uv = __ldg(edges + i);
            const unsigned peers = __match_any_sync(mask, uv.y);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            const int pos = eb + pv[uv.y] + rank;
            p.src[pos] = nb + uv.x;
            if (p.has_attr)
            {
                const int* a = p.edge_attr + 3 * (size_t)(eb_in + i);
                const int a0 = __ldg(a), a1 = __ldg(a + 1), a2 = __ldg(a + 2);
                // bond features outside the vocabulary {5, 6, 2} (GIN/src/host_load.cc:5-6) would alias another triple
                // or index past the combined table: reject the batch (code 0 keeps the layer kernels in bounds)
                const bool oov = (unsigned)a0 >= 5u || (unsigned)a1 >= 6u || (unsigned)a2 >= 2u;
                if (oov) atomicOr(p.status, 4);
                p.code[pos] = oov ? (uint8_t)0 : (uint8_t)(a0 * 12 + a1 * 2 + a2);
            }
            if (p.flags & PREP_GCN_NORM)
            {
                // recip(sqrt(deg+1)); a node that never appears as a source keeps dis = 0
                const float du = (deg[uv.x] > 0) ? __frcp_rn(__fsqrt_rn((float)(deg[uv.x] + 1))) : 0.0f;
                const float dv = (deg[uv.y] > 0) ? __frcp_rn(__fsqrt_rn((float)(deg[uv.y] + 1))) : 0.0f;
                p.edge_w[pos] = __fmul_rn(du, dv);
            }
            else if (p.flags & PREP_DGN_EIG)
            {
                const float* eig = p.node_eigen + 4 * (size_t)nb_in;
                p.edge_w[pos] = __fsub_rn(__ldg(eig + 4 * uv.x + 1), __ldg(eig + 4 * uv.y + 1));
            }
            __syncwarp(mask);
            if (rank == __popc(peers) - 1) pv[uv.y] += rank + 1;
        }
        __syncwarp();
    }
    if (p.row_desc)
    {
        __threadfence_block();
        __syncwarp();
        for (int v = lane; v < n; v += 32)
        {
            const int end = eb + pv[v];
            const int beg = (v == 0) ? eb : eb + pv[v - 1];
            int4 d = empty_row_desc();
            int* dq = reinterpret_cast<int*>(&d);
            for (int q = 0; q < 4 && beg + q < end; q++)
            {
                const int rel = p.src[beg + q] - (nb + v);
                if (rel < -32768 || rel > 32767) atomicOr(p.status, 8);      // the descriptor holds 16-bit relative positions
                dq[q] = ((rel + 32768) & 0xFFFF) | ((p.has_attr ? (int)p.code[beg + q] : 0) << 16);
            }
            d.x |= min(end - beg, 255) << 24;
            p.row_desc[nb + v] = d;
        }
    }
    if (p.flags & PREP_DGN_EIG)
    {
        // per-destination sums over the (now sorted) in-edges; pv[v] ends at the end of row v
        __threadfence_block();
        __syncwarp();
        for (int v = lane; v < n; v += 32)
        {
            const int end = eb + pv[v];
            const int beg = (v == 0) ? eb : eb + pv[v - 1];
            float sa = 0.0f, sw = 0.0f;
            for (int k = beg; k < end; k++)
            {
                const float w = p.edge_w[k];
                sa = __fadd_rn(sa, fabsf(w));
                sw = __fadd_rn(sw, w);
            }
            p.node_w0[nb + v] = sa;
            p.node_w1[nb + v] = sw;
        }
    }
}

// small graphs: one warp per graph, 8 warps per block
__global__ void __launch_bounds__(8 * 32) build_csr_small_kernel(CsrParams p)
{
    __shared__ int s_tab[8][3][CSR_NCAP_SMALL];
    const int wid = threadIdx.x >> 5;
    const int g = blockIdx.x * 8 + wid;
    if (g >= p.num_graphs) return;
    const int n = p.nn[g], e = p.ne[g];
    if (!(n >= 0 && n <= CSR_NCAP_SMALL && e >= 0)) return;
    build_graph_csr(p, g, CSR_NCAP_SMALL, s_tab[wid][0], s_tab[wid][1], s_tab[wid][2]);
}

// the rest (more than 128 nodes, or invalid counts): a persistent grid whose warps scan 32 graphs at a time, so that a
// batch of molecules costs this kernel a microsecond
__global__ void __launch_bounds__(4 * 32) build_csr_large_kernel(CsrParams p)
{
    __shared__ int s_tab[4][3][CSR_NCAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warp = blockIdx.x * 4 + wid, nwarps = gridDim.x * 4;
    for (int g0 = warp * 32; g0 < p.num_graphs; g0 += nwarps * 32)
    {
        const int g = g0 + lane;
        bool mine = false;
        if (g < p.num_graphs)
        {
            const int n = p.nn[g], e = p.ne[g];
            mine = !(n >= 0 && n <= CSR_NCAP_SMALL && e >= 0);
        }
        unsigned todo = __ballot_sync(0xffffffffu, mine);
        while (todo)
        {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const int gk = g0 + k, nk = p.nn[gk];
            if (nk > CSR_NCAP && p.big_tab)
            {
                // a graph beyond the shared-memory tables: the same build on tables in global memory (node ranges of
                // different graphs are disjoint, so the batch-wide arrays are indexed by global node id)
                int* t = p.big_tab + p.node_off[gk];
                build_graph_csr(p, gk, 0x7fffffff, t, t + p.total_nodes, t + 2 * p.total_nodes);
            }
            else build_graph_csr(p, gk, CSR_NCAP, s_tab[wid][0], s_tab[wid][1], s_tab[wid][2]);
            __syncwarp();
        }
    }
}


// ---- graph-aligned tiles for the GIN layer kernel (gin_fused.cu) ---------------------------------------------------
// The layer kernel stages the feature rows of one tile (<= 128 consecutive nodes) in shared memory with ONE bulk copy
// and gathers the in-edges from that stage.  In-edge sources are nodes of the SAME graph, so tiles are packed from
// whole graphs: no gather ever leaves its stage.  A graph of more than 128 nodes is cut into 128-row tiles flagged
// "external" (bit 30 of .y): their gathers read source rows from global memory.
// One block: thread i packs a contiguous chunk of graphs greedily (a tile closes when the next graph does not fit, and at
// the end of the chunk), a block-wide scan of the per-chunk tile counts places the chunks' tiles, a second pass writes
// them.  tiles[t] = (first node, rows | ext << 30); tile_count[0] = number of tiles.
constexpr int TILE_ROWS = 128;
constexpr int PACK_THREADS = 1024;

template <bool WRITE>
__device__ __forceinline__ int pack_chunk(const int* __restrict__ nn, const int* node_off, int g0, int g1, int2* out)
{
    int count = 0, start = 0, rows = 0;
    for (int g = g0; g < g1; g++)
    {
        const int n = nn[g];
        if (n <= 0) continue;
        if (n > TILE_ROWS)
        {
            if (rows > 0) { if (WRITE) out[count] = make_int2(start, rows); count++; rows = 0; }
            const int nb = node_off[g];
            for (int r = 0; r < n; r += TILE_ROWS)
            {
                if (WRITE) out[count] = make_int2(nb + r, min(TILE_ROWS, n - r) | (1 << 30));
                count++;
            }
            continue;
        }
        if (rows + n > TILE_ROWS) { if (WRITE) out[count] = make_int2(start, rows); count++; rows = 0; }
        if (rows == 0) start = node_off[g];
        rows += n;
    }
    if (rows > 0) { if (WRITE) out[count] = make_int2(start, rows); count++; }
    return count;
}

__device__ __forceinline__ void pack_tiles_body(const int* __restrict__ nn, const int* node_off, int num_graphs, int2* __restrict__ tiles,
                                                int* __restrict__ tile_count)
{
    __shared__ int warp_tot[PACK_THREADS / 32];
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (num_graphs + PACK_THREADS - 1) / PACK_THREADS;
    const int g0 = min(tid * per, num_graphs), g1 = min(g0 + per, num_graphs);
    const int mine = pack_chunk<false>(nn, node_off, g0, g1, nullptr);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const int a = __shfl_up_sync(full, incl, d);
        if (lane >= d) incl += a;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0)
    {
        const int t = warp_tot[lane];
        int ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int a = __shfl_up_sync(full, ti, d);
            if (lane >= d) ti += a;
        }
        warp_tot[lane] = ti - t;
        if (lane == 31) tile_count[0] = ti;
    }
    __syncthreads();
    pack_chunk<true>(nn, node_off, g0, g1, tiles + warp_tot[wid] + incl - mine);
}

// offsets scan + tile packing as ONE launch: both are single-block kernels and the second only needs the first's node offsets
// (every launch of the prep chain costs the chunked host-pointer path ~10 us per chunk)
static_assert(PACK_THREADS == SCAN_THREADS, "scan_pack_kernel runs both bodies on one block");
__global__ void __launch_bounds__(PACK_THREADS) scan_pack_kernel(const int* __restrict__ nn, const int* __restrict__ ne, int* node_off,
                                                                 int* __restrict__ edge_off, int num_graphs, int2* __restrict__ tiles,
                                                                 int* __restrict__ tile_count)
{
    scan_offsets_body(nn, ne, node_off, edge_off, num_graphs);
    __syncthreads();                       // node_off is complete and visible to the block
    pack_tiles_body(nn, node_off, num_graphs, tiles, tile_count);
}

// Row descriptors of every tile re-ordered by in-degree (stable), the row's position inside the tile in bits 24..30 of .y.
// The layer kernel (gin_fused.cu) hands FOUR rows to every warp instruction of its gather and pads them to the longest
// of their in-edge lists: with the rows of a tile grouped by degree there is next to nothing to pad (average slots per
// row 2.2 instead of 3.0 on molecules).  One warp per tile, counting sort with __match_any_sync ranks.
__global__ void __launch_bounds__(256) sort_tile_rows_kernel(const int2* __restrict__ tiles, const int* __restrict__ tile_count,
                                                             const int4* __restrict__ row_desc, int4* __restrict__ sorted)
{
    __shared__ int s_off[8][32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ntiles = *tile_count;
    for (int t = blockIdx.x * 8 + wid; t < ntiles; t += gridDim.x * 8)
    {
        const int2 ti = tiles[t];
        const int start = ti.x, rows = ti.y & 0xFFFF;
        int4 d[TILE_ROWS / 32];
        int key[TILE_ROWS / 32];
        int cnt = 0;                                   // rows of key == lane
#pragma unroll
        for (int i = 0; i < TILE_ROWS / 32; i++)
        {
            const int r = lane + 32 * i;
            key[i] = -1;
            if (r < rows)
            {
                d[i] = row_desc[start + r];
                key[i] = min((int)((unsigned)d[i].x >> 24), 31);
            }
#pragma unroll
            for (int k = 0; k < 32; k++)
            {
                const unsigned m = __ballot_sync(full, key[i] == k);
                if (lane == k) cnt += __popc(m);
            }
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int a = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += a;
        }
        s_off[wid][lane] = incl - cnt;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < TILE_ROWS / 32; i++)
        {
            const int r = lane + 32 * i;
            const bool act = r < rows;
            const unsigned mask = __ballot_sync(full, act);
            if (act)
            {
                const unsigned peers = __match_any_sync(mask, key[i]);
                const int rank = __popc(peers & ((1u << lane) - 1u));
                const int pos = s_off[wid][key[i]] + rank;
                int4 o = d[i];
                o.y = (o.y & 0x00FFFFFF) | (r << 24);
                sorted[start + pos] = o;
                __syncwarp(mask);
                if (rank == __popc(peers) - 1) s_off[wid][key[i]] += rank + 1;
            }
            __syncwarp();
        }
    }
}

}  // namespace

// Small fills and device-to-device copies as KERNELS.  cudaMemsetAsync / cudaMemcpyAsync(D2D) are served by the copy engines, where
// they queue behind whatever upload is in flight: in the chunked entry points every such call inside a forward waited 0.25-0.35 ms
// for the next chunk's host-to-device copy (DGN: one memset per layer, +1 ms per chunk).
namespace {
__global__ void __launch_bounds__(256) zero_bytes_kernel(uint4* __restrict__ p, size_t n16)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = make_uint4(0u, 0u, 0u, 0u);
}
struct CopySeg { int* dst; const int* src; long n; };
__global__ void __launch_bounds__(256) copy_segments_kernel(CopySeg a, CopySeg b, CopySeg c, CopySeg d, int* zero_word)
{
    const long stride = (long)gridDim.x * blockDim.x, t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long i = t; i < a.n; i += stride) a.dst[i] = __ldg(a.src + i);
    for (long i = t; i < b.n; i += stride) b.dst[i] = __ldg(b.src + i);
    for (long i = t; i < c.n; i += stride) c.dst[i] = __ldg(c.src + i);
    for (long i = t; i < d.n; i += stride) d.dst[i] = __ldg(d.src + i);
    if (t == 0 && zero_word) *zero_word = 0;
}
}  // namespace

// zero [p, p + bytes): p from cudaMalloc (256-byte aligned), the buffer padded to a multiple of 16 bytes (DevBuf::reserve rounds to 256)
int zero_bytes_launch(void* p, size_t bytes, cudaStream_t stream)
{
    const size_t n16 = (bytes + 15) / 16;
    if (!n16) return 0;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>((n16 + 255) / 256, 148 * 4));
    zero_bytes_kernel<<<blocks, 256, 0, stream>>>(static_cast<uint4*>(p), n16);
    FG_CUDA(cudaGetLastError());
    return 0;
}


int prep_batch(DeviceBatch& b, int flags, cudaStream_t stream, int* launches, bool use_perm)
{
    const int G = b.num_graphs;
    if (G <= 0) return 0;
    FG_TRY(b.node_off.reserve(sizeof(int) * (size_t)(G + 1)));
    FG_TRY(b.edge_off.reserve(sizeof(int) * (size_t)(G + 1)));
    FG_TRY(b.in_ptr.reserve(sizeof(int) * (size_t)(b.total_nodes + 1)));
    FG_TRY(b.src.reserve(sizeof(int) * (size_t)(b.total_edges + 1)));
    FG_TRY(b.code.reserve((size_t)(b.total_edges + 16)));
    FG_TRY(b.out_deg.reserve(sizeof(int) * (size_t)(b.total_nodes + 1)));
    FG_TRY(b.sort_tmp.reserve(sizeof(int) * (size_t)(b.total_edges + 1)));
    FG_TRY(b.status.reserve(sizeof(int)));
    if (flags & PREP_ROW_DESC) FG_TRY(b.row_desc.reserve(sizeof(int4) * (size_t)(b.total_nodes + 1)));
    if (flags & (PREP_GCN_NORM | PREP_DGN_EIG)) FG_TRY(b.edge_w.reserve(sizeof(float) * (size_t)(b.total_edges + 1)));
    if (flags & PREP_DGN_EIG)
    {
        FG_TRY(b.node_w0.reserve(sizeof(float) * (size_t)(b.total_nodes + 1)));
        FG_TRY(b.node_w1.reserve(sizeof(float) * (size_t)(b.total_nodes + 1)));
    }
    // Re-ordered graphs (api.cu::pack_graphs, computed on the host at upload): the rows of graph g are written at node_off_perm[g],
    // its inputs are still read at the caller's offsets; the tile list came with the upload.  Everything downstream uses node_off.
    const bool perm = use_perm && b.has_perm && (flags & PREP_TILES);
    b.perm_active = perm;
    int nl = 1;
    int* scan_node = perm ? b.node_off_in.as<int>() : b.node_off.as<int>();
    int* scan_edge = perm ? b.edge_off_in.as<int>() : b.edge_off.as<int>();
    if (!perm) { FG_TRY(zero_bytes_launch(b.status.ptr, sizeof(int), stream)); nl++; }
    if (perm)
    {
        b.max_tiles = b.tiles_perm_count;
        FG_TRY(b.tiles.reserve(sizeof(int2) * (size_t)std::max<long>(b.max_tiles, 1)));
        FG_TRY(b.tile_count.reserve(sizeof(int)));
        // adopt what came with the upload (one launch, not four copy-engine copies and a memset) and clear the status word
        const CopySeg sa{b.node_off.as<int>(), b.node_off_perm.as<int>(), (long)G + 1}, sb{b.edge_off.as<int>(), b.edge_off_perm.as<int>(), (long)G + 1};
        const CopySeg sc{b.tiles.as<int>(), b.tiles_perm.as<int>(), 2 * b.max_tiles}, sd{b.tile_count.as<int>(), b.tiles_perm.as<int>() + 2 * b.max_tiles, 1};
        copy_segments_kernel<<<(int)std::max<long>(1, std::min<long>(((long)G + 256) / 256, 148 * 2)), 256, 0, stream>>>(sa, sb, sc, sd, b.status.as<int>());
        FG_CUDA(cudaGetLastError());
        nl = 1;                                             // (the caller-order offsets came with the upload too: no scan launch)
    }
    else if (flags & PREP_TILES)
    {
        // upper bound: a tile closes at most once per graph, per 128 rows of a large graph and per packing chunk
        b.max_tiles = (long)G + b.total_nodes / TILE_ROWS + PACK_THREADS + 1;
        FG_TRY(b.tiles.reserve(sizeof(int2) * (size_t)b.max_tiles));
        FG_TRY(b.tile_count.reserve(sizeof(int)));
        scan_pack_kernel<<<1, PACK_THREADS, 0, stream>>>(b.nums_of_nodes.as<int>(), b.nums_of_edges.as<int>(), b.node_off.as<int>(), b.edge_off.as<int>(), G,
                                                         b.tiles.as<int2>(), b.tile_count.as<int>());
    }
    else
        scan_offsets_kernel<<<1, SCAN_THREADS, 0, stream>>>(b.nums_of_nodes.as<int>(), b.nums_of_edges.as<int>(), b.node_off.as<int>(),
                                                             b.edge_off.as<int>(), G);
    FG_CUDA(cudaGetLastError());

    CsrParams p;
    p.nn = b.nums_of_nodes.as<int>(); p.ne = b.nums_of_edges.as<int>();
    p.node_off = b.node_off.as<int>(); p.edge_off = b.edge_off.as<int>();
    p.node_in_off = scan_node; p.edge_in_off = scan_edge;
    p.node_map = nullptr;                                   // written by node_map_launch (on the embedding's stream)
    p.edge_list = b.edge_list.as<int>(); p.edge_attr = b.edge_attr.as<int>(); p.node_eigen = b.node_eigen.as<float>();
    p.in_ptr = b.in_ptr.as<int>(); p.src = b.src.as<int>(); p.code = b.code.as<uint8_t>(); p.edge_w = b.edge_w.as<float>();
    p.out_deg = b.out_deg.as<int>(); p.node_w0 = b.node_w0.as<float>(); p.node_w1 = b.node_w1.as<float>();
    p.sort_tmp = b.sort_tmp.as<int>(); p.status = b.status.as<int>();
    p.big_tab = nullptr; p.total_nodes = b.total_nodes;
    if (b.max_graph_nodes > CSR_NCAP)
    {
        FG_TRY(b.big_tab.reserve(sizeof(int) * 3 * (size_t)(b.total_nodes + 1)));
        p.big_tab = b.big_tab.as<int>();
    }
    p.row_desc = (flags & PREP_ROW_DESC) ? b.row_desc.as<int4>() : nullptr;
    p.num_graphs = G; p.flags = flags; p.has_attr = b.has_attr ? 1 : 0;
    build_csr_small_kernel<<<ceil_div(G, 8), 8 * 32, 0, stream>>>(p);
    FG_CUDA(cudaGetLastError());
    nl++;
    if (b.max_graph_nodes > CSR_NCAP_SMALL)          // counts are validated on the host (upload_into): nothing else is left for this kernel
    {
        build_csr_large_kernel<<<std::max(1, std::min(ceil_div(G, 128), 148 * 4)), 4 * 32, 0, stream>>>(p);
        FG_CUDA(cudaGetLastError());
        nl++;
    }
    if ((flags & PREP_TILES) && (flags & PREP_ROW_DESC))
    {
        FG_TRY(b.row_desc_sorted.reserve(sizeof(int4) * (size_t)(b.total_nodes + 1)));
        sort_tile_rows_kernel<<<(int)std::max<long>(1, std::min<long>(ceil_div<long>(b.max_tiles, 8), 148 * 8)), 256, 0, stream>>>(
            b.tiles.as<int2>(), b.tile_count.as<int>(), b.row_desc.as<int4>(), b.row_desc_sorted.as<int4>());
        FG_CUDA(cudaGetLastError());
        nl++;
    }
    if (launches) *launches += nl;
    return 0;
}

namespace {
// row -> caller-order node index of a re-ordered batch: a warp per graph
__global__ void __launch_bounds__(256) node_map_kernel(const int* __restrict__ nn, const int* __restrict__ node_off, const int* __restrict__ node_in_off,
                                                       int* __restrict__ node_map, int num_graphs)
{
    const int lane = threadIdx.x & 31;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < num_graphs; g += (gridDim.x * blockDim.x) >> 5)
    {
        const int n = __ldg(nn + g), nb = __ldg(node_off + g), nb_in = __ldg(node_in_off + g);
        for (int i = lane; i < n; i += 32) node_map[nb + i] = nb_in + i;
    }
}
}  // namespace

// For the embedding kernels of a re-ordered batch.  Reads only what came with the upload (node_off_perm, node_off_in), so it can run on
// the embedding's stream next to the CSR build.
int node_map_launch(DeviceBatch& b, cudaStream_t stream)
{
    FG_TRY(b.node_map.reserve(sizeof(int) * (size_t)(b.total_nodes + 1)));
    node_map_kernel<<<std::max(1, std::min(ceil_div(b.num_graphs, 8), 148 * 8)), 256, 0, stream>>>(b.nums_of_nodes.as<int>(), b.node_off_perm.as<int>(),
                                                                                                  b.node_off_in.as<int>(), b.node_map.as<int>(), b.num_graphs);
    FG_CUDA(cudaGetLastError());
    return 0;
}

namespace {
// Widen the narrowed inputs of the host-pointer entry points (host_stage.h) into the int32 arrays of the caller layout:
// a thread takes four source values (one 4-byte / 8-byte load) and stores one int4, so a warp reads 128 / 256 contiguous
// bytes and writes 512.  The three regions of the block start at multiples of 16 bytes and are padded to them.
__global__ void __launch_bounds__(256) unpack_inputs_kernel(const uint8_t* __restrict__ block, size_t off_edge, size_t off_attr,
                                                            int32_t* __restrict__ feat, size_t n_feat, int32_t* __restrict__ edges, size_t n_edge,
                                                            int32_t* __restrict__ attr, size_t n_attr)
{
    const size_t ga = (n_feat + 3) / 4, gb = (n_edge + 3) / 4, gc = (n_attr + 3) / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ga + gb + gc; g += stride)
    {
        int4 v;
        int32_t* dst;
        size_t left;
        if (g >= ga && g < ga + gb)
        {
            const size_t k = g - ga;
            const uint2 w = __ldg(reinterpret_cast<const uint2*>(block + off_edge) + k);
            v = make_int4((int)(w.x & 0xFFFFu), (int)(w.x >> 16), (int)(w.y & 0xFFFFu), (int)(w.y >> 16));
            dst = edges + 4 * k; left = n_edge - 4 * k;
        }
        else
        {
            const bool a = g < ga;
            const size_t k = a ? g : g - ga - gb;
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(block + (a ? 0 : off_attr)) + k);
            v = make_int4((int)(w & 0xFFu), (int)((w >> 8) & 0xFFu), (int)((w >> 16) & 0xFFu), (int)(w >> 24));
            dst = (a ? feat : attr) + 4 * k; left = (a ? n_feat : n_attr) - 4 * k;
        }
        if (left >= 4) *reinterpret_cast<int4*>(dst) = v;
        else
        {
            dst[0] = v.x;
            if (left > 1) dst[1] = v.y;
            if (left > 2) dst[2] = v.z;
        }
    }
}
}  // namespace

// counts of 0 skip a region (that array was uploaded unchanged because it holds values outside the narrow range)
int unpack_inputs_launch(const uint8_t* block, size_t off_edge, size_t off_attr, int32_t* feat, size_t n_feat, int32_t* edges, size_t n_edge,
                         int32_t* attr, size_t n_attr, cudaStream_t stream)
{
    const size_t groups = (n_feat + 3) / 4 + (n_edge + 3) / 4 + (n_attr + 3) / 4;
    if (!groups) return 0;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>((groups + 255) / 256, 148 * 8));
    unpack_inputs_kernel<<<blocks, 256, 0, stream>>>(block, off_edge, off_attr, feat, n_feat, edges, n_edge, attr, n_attr);
    FG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fg
