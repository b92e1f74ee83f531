/* TEST INFRASTRUCTURE (oracle) -- see flowgnn_oracle.h.  NOT on the product path.
 *
 * Plain-C, fp32, re-entrant restatement of the reference's six kernels.  The reference is an
 * HLS dataflow design (streams, 4 message-passing PEs banked by v%4, ping/pong message
 * buffers); none of that structure is kept.  What IS kept is the order of every fp32
 * operation, so that this file and the reference's sources compiled with the float shim
 * (oracle/_ref, -O2 -ffp-contract=off) agree bit for bit:
 *
 *  - a destination's in-edges are accumulated in (source ascending, then edge-list order),
 *    because each PE walks a CSR-by-source (GIN/src/load_inputs.cc:140-171,
 *    message_passing.cc:110-147) and one destination lives in exactly one PE;
 *  - dense layers start from the bias and add products in dim_in order
 *    (GIN/src/node_embedding.cc:124-135, linear.cc:34-44);
 *  - the mean pool adds nodes in pairs (v, v+1) before folding them into the running sum
 *    (GIN/src/finalize.cc:46-113).
 */
#include "flowgnn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ND_FEATURE 9
#define ND_FEATURE_TOTAL 173
#define ED_PER_LAYER 13

static const int nd_feature_offsets[ND_FEATURE] = {0, 119, 123, 135, 147, 157, 163, 169, 171}; /* GIN/src/load_inputs.cc:5 */
static const int ed_feature_offsets[3] = {0, 5, 11};                                            /* GIN/src/message_passing.cc:3 */

/* ap_fixed_relu: hls::signbit(x) ? 0 : x  (GIN/src/util.h:20-25); NaN passes through. */
static inline float relu(float x) { return (x < 0.0f) ? 0.0f : x; }

static int max_i32(const int32_t* a, int n)
{
    int m = 0;
    for (int i = 0; i < n; i++) if (a[i] > m) m = a[i];
    return m;
}

/* ---- load_graph: COO -> CSR by source, stable in edge-list order ---------------------------
 * (GIN/src/load_inputs.cc:87-172).  `order[k]` lists edge indices grouped by source u
 * ascending; inside one u they keep edge-list order.  Walking `order` front to back therefore
 * visits, for any fixed destination v, its in-edges exactly as the v%4 PE does. */
typedef struct {
    int* out_deg;   /* [N]   degree_table */
    int* in_deg;    /* [N] */
    int* out_ptr;   /* [N+1] */
    int* order;     /* [E] */
} csr_t;

static void csr_alloc(csr_t* c, int max_n, int max_e)
{
    c->out_deg = (int*)malloc(sizeof(int) * (size_t)(max_n + 1));
    c->in_deg = (int*)malloc(sizeof(int) * (size_t)(max_n + 1));
    c->out_ptr = (int*)malloc(sizeof(int) * (size_t)(max_n + 2));
    c->order = (int*)malloc(sizeof(int) * (size_t)(max_e + 1));
}

static void csr_free(csr_t* c)
{
    free(c->out_deg); free(c->in_deg); free(c->out_ptr); free(c->order);
}

static void csr_build(csr_t* c, const int32_t* edges, int n, int e)
{
    for (int i = 0; i < n; i++) { c->out_deg[i] = 0; c->in_deg[i] = 0; }
    for (int i = 0; i < e; i++) { c->out_deg[edges[2 * i]]++; c->in_deg[edges[2 * i + 1]]++; }
    c->out_ptr[0] = 0;
    for (int i = 0; i < n; i++) c->out_ptr[i + 1] = c->out_ptr[i] + c->out_deg[i];
    int* cursor = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    memcpy(cursor, c->out_ptr, sizeof(int) * (size_t)n);
    for (int i = 0; i < e; i++) c->order[cursor[edges[2 * i]]++] = i;
    free(cursor);
}

/* ---- load_input_node_embeddings: h0_v = sum_f Table[off_f + x_vf], added in f order from 0
 * (GIN/src/load_inputs.cc:174-220; same in GCN :168-215 and PNA :133-179). */
static void embed_nodes(float* h, const int32_t* feat, const float* table, int n, int dim)
{
    for (int v = 0; v < n; v++)
        for (int d = 0; d < dim; d++)
        {
            float s = 0.0f;
            for (int f = 0; f < ND_FEATURE; f++)
                s += table[(size_t)(nd_feature_offsets[f] + feat[v * ND_FEATURE + f]) * dim + d];
            h[(size_t)v * dim + d] = s;
        }
}

/* ---- global_mean_pooling (GIN/src/finalize.cc:36-115; identical in GAT/PNA/DGN) ------------ */
static void mean_pool_pairs(const float* h, int n, int dim, float* h_graph)
{
    int num_iters = (n + 1) / 2 - 1;
    int tail_nodes = ((n - 1) % 2) + 1;
    for (int d = 0; d < dim; d++)
    {
        float sums = 0.0f;
        for (int i = 0; i < num_iters; i++)
        {
            float el = 0.0f;
            el += h[(size_t)(2 * i) * dim + d];
            el += h[(size_t)(2 * i + 1) * dim + d];
            if (i != 0) el += sums;
            sums = el;
        }
        float el = 0.0f;
        for (int k = 0; k < tail_nodes; k++) el += h[(size_t)(2 * num_iters + k) * dim + d];
        if (num_iters != 0) el += sums;
        h_graph[d] = el / n;
    }
}

/* ---- linear.cc ---------------------------------------------------------------------------- */
/* linear / linear_output_stationary: out = bias; out += in[i] * w[o][i]  (linear.cc:34-44, 82-92) */
static void linear_os(const float* in, const float* w, const float* b, float* out, int din, int dout, int do_relu)
{
    for (int o = 0; o < dout; o++)
    {
        float acc = b[o];
        for (int i = 0; i < din; i++) acc += in[i] * w[(size_t)o * din + i];
        if (do_relu && acc < 0.0f) acc = 0.0f;
        out[o] = acc;
    }
}

/* linear_input_stationary with PARALLEL=2: out = bias; per input pair: addend = 0 + in0*w0 + in1*w1;
 * out += addend  (linear.cc:115-149) */
static void linear_is2(const float* in, const float* w, const float* b, float* out, int din, int dout, int do_relu)
{
    for (int o = 0; o < dout; o++) out[o] = b[o];
    for (int base = 0; base < din; base += 2)
        for (int o = 0; o < dout; o++)
        {
            float addend = 0.0f;
            for (int k = 0; k < 2; k++)
                if (base + k < din) addend += in[base + k] * w[(size_t)o * din + base + k];
            out[o] += addend;
        }
    if (do_relu)
        for (int o = 0; o < dout; o++) if (out[o] < 0.0f) out[o] = 0.0f;
}

/* edge embedding: ((0 + T[a0]) + T[5+a1]) + T[11+a2]  (GIN/src/message_passing.cc:136-142) */
static inline float edge_embed(const float* ee_layer, const int32_t* attr, int d, int dim)
{
    float s = 0.0f;
    for (int f = 0; f < 3; f++) s += ee_layer[(size_t)(ed_feature_offsets[f] + attr[f]) * dim + d];
    return s;
}

/* ============================================================================================
 * GIN / GIN-VN  (GIN/src/GIN_compute.cc:44-98)
 * ========================================================================================== */
void oracle_GIN_compute_graphs(
    int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* reload_weights,
    float* out, const int32_t* node_feature_in, const int32_t* edge_list_in, const int32_t* edge_attr_in,
    const float* node_embedding_weight_in, const float* edge_embedding_weight_in,
    const float* node_mlp_1_weights, const float* node_mlp_1_bias,
    const float* node_mlp_2_weights, const float* node_mlp_2_bias,
    const float* graph_pred_weights_in, const float* graph_pred_bias_in)
{
    enum { D = 100, H = 200, L = 5 };
    int max_n = max_i32(nums_of_nodes, num_graphs), max_e = max_i32(nums_of_edges, num_graphs);
    csr_t c; csr_alloc(&c, max_n, max_e);
    float* h = (float*)malloc(sizeof(float) * (size_t)(max_n + 1) * D);
    float* m = (float*)malloc(sizeof(float) * (size_t)(max_n + 1) * D);
    float acc[H], h_graph[D];

    long nodes_offset = 0, edges_offset = 0;
    int weights_ndx = -1;
    for (int g = 0; g < num_graphs; g++)
    {
        int n = nums_of_nodes[g], e = nums_of_edges[g];
        if (reload_weights[g]) weights_ndx++;
        const float* ne_w = node_embedding_weight_in + (size_t)weights_ndx * ND_FEATURE_TOTAL * D;
        const float* ee_w = edge_embedding_weight_in + (size_t)weights_ndx * L * ED_PER_LAYER * D;
        const float* w1 = node_mlp_1_weights + (size_t)weights_ndx * L * H * D;
        const float* b1 = node_mlp_1_bias + (size_t)weights_ndx * L * H;
        const float* w2 = node_mlp_2_weights + (size_t)weights_ndx * L * D * H;
        const float* b2 = node_mlp_2_bias + (size_t)weights_ndx * L * D;
        const float* pw = graph_pred_weights_in + (size_t)weights_ndx * D;
        const float* pb = graph_pred_bias_in + (size_t)weights_ndx;
        const int32_t* edges = edge_list_in + 2 * edges_offset;
        const int32_t* attrs = edge_attr_in + 3 * edges_offset;

        csr_build(&c, edges, n, e);
        embed_nodes(h, node_feature_in + nodes_offset * ND_FEATURE, ne_w, n, D);

        for (int l = 0; l < L; l++)
        {
            /* message passing: m_v += relu(h_u + edge_embed)  (GIN/src/message_passing.cc:110-147) */
            memset(m, 0, sizeof(float) * (size_t)n * D);
            const float* ee_l = ee_w + (size_t)l * ED_PER_LAYER * D;
            for (int k = 0; k < e; k++)
            {
                int ei = c.order[k], u = edges[2 * ei], v = edges[2 * ei + 1];
                for (int d = 0; d < D; d++)
                {
                    float total = edge_embed(ee_l, attrs + 3 * ei, d, D) + h[(size_t)u * D + d];
                    m[(size_t)v * D + d] += relu(total);
                }
            }
            /* node transform: MLP 100 -> 200 -> 100; eps is never loaded, so (1+eps) == 1
             * (GIN/src/node_embedding.cc:117,124-135,165-191) */
            const float* w1l = w1 + (size_t)l * H * D; const float* b1l = b1 + (size_t)l * H;
            const float* w2l = w2 + (size_t)l * D * H; const float* b2l = b2 + (size_t)l * D;
            for (int v = 0; v < n; v++)
            {
                for (int i = 0; i < D; i++)
                {
                    float a = m[(size_t)v * D + i] + 1.0f * h[(size_t)v * D + i];
                    for (int o = 0; o < H; o++)
                    {
                        float addend = a * w1l[(size_t)o * D + i];
                        acc[o] = addend + ((i == 0) ? b1l[o] : acc[o]);
                    }
                }
                for (int o = 0; o < D; o++)
                {
                    float r = b2l[o];
                    for (int i = 0; i < H; i++) r += relu(acc[i]) * w2l[(size_t)o * H + i];
                    if (l != L - 1) r = relu(r);
                    h[(size_t)v * D + o] = r;
                }
            }
        }
        mean_pool_pairs(h, n, D, h_graph);
        linear_os(h_graph, pw, pb, out + g, D, 1, 0);                   /* GIN/src/finalize.cc:28-33 */

        nodes_offset += n; edges_offset += e;
    }
    free(h); free(m); csr_free(&c);
}

/* ============================================================================================
 * GCN  (GCN/src/GCN_compute.cc:50-102)
 * ========================================================================================== */
void oracle_GCN_compute_graphs(
    int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* reload_weights,
    float* out, const int32_t* node_feature_in, const int32_t* edge_list_in, const int32_t* edge_attr_in,
    const float* node_embedding_weight_in, const float* edge_embedding_weight_in,
    const float* convs_weight_in, const float* convs_bias_in, const float* convs_root_emb_weight_in,
    const float* bn_weight_in, const float* bn_bias_in, const float* bn_mean_in, const float* bn_var_in,
    const float* graph_pred_weights_in, const float* graph_pred_bias_in)
{
    enum { D = 100, L = 5 };
    int max_n = max_i32(nums_of_nodes, num_graphs), max_e = max_i32(nums_of_edges, num_graphs);
    csr_t c; csr_alloc(&c, max_n, max_e);
    float* h = (float*)malloc(sizeof(float) * (size_t)(max_n + 1) * D);
    float* m = (float*)malloc(sizeof(float) * (size_t)(max_n + 1) * D);
    float* dis = (float*)malloc(sizeof(float) * (size_t)(max_n + 1));
    float bn_sqrt_var[L * D], acc[D], h_graph[D];

    long nodes_offset = 0, edges_offset = 0;
    int weights_ndx = -1;
    for (int g = 0; g < num_graphs; g++)
    {
        int n = nums_of_nodes[g], e = nums_of_edges[g];
        if (reload_weights[g])
        {
            weights_ndx++;
            /* bn_sqrt_var = sqrt(var + 2^-10)  (GCN/src/load_inputs.cc:32, util.h:27-32) */
            const float* var = bn_var_in + (size_t)weights_ndx * L * D;
            for (int i = 0; i < L * D; i++) bn_sqrt_var[i] = sqrtf(var[i] + (float)(1.0 / (1 << 10)));
        }
        const float* ne_w = node_embedding_weight_in + (size_t)weights_ndx * ND_FEATURE_TOTAL * D;
        const float* ee_w = edge_embedding_weight_in + (size_t)weights_ndx * L * ED_PER_LAYER * D;
        const float* cw = convs_weight_in + (size_t)weights_ndx * L * D * D;
        const float* cb = convs_bias_in + (size_t)weights_ndx * L * D;
        const float* root = convs_root_emb_weight_in + (size_t)weights_ndx * L * D;
        const float* bnw = bn_weight_in + (size_t)weights_ndx * L * D;
        const float* bnb = bn_bias_in + (size_t)weights_ndx * L * D;
        const float* bnm = bn_mean_in + (size_t)weights_ndx * L * D;
        const float* pw = graph_pred_weights_in + (size_t)weights_ndx * D;
        const float* pb = graph_pred_bias_in + (size_t)weights_ndx;
        const int32_t* edges = edge_list_in + 2 * edges_offset;
        const int32_t* attrs = edge_attr_in + 3 * edges_offset;

        csr_build(&c, edges, n, e);
        /* degree_inv_sqrt is only ever written for nodes that appear as a source; others stay 0
         * (GCN/src/load_inputs.cc:100-122) */
        for (int i = 0; i < n; i++) dis[i] = (c.out_deg[i] > 0) ? 1.0f / sqrtf((float)(c.out_deg[i] + 1)) : 0.0f;
        embed_nodes(h, node_feature_in + nodes_offset * ND_FEATURE, ne_w, n, D);
        memset(m, 0, sizeof(float) * (size_t)n * D);

        for (int l = 0; l < L; l++)
        {
            /* node transform: finish layer l-1 (self term, BN, relu), then Linear_l
             * (GCN/src/node_embedding.cc:98-146) */
            for (int v = 0; v < n; v++)
            {
                for (int i = 0; i < D; i++)
                {
                    float activation;
                    if (l == 0) activation = h[(size_t)v * D + i];
                    else
                    {
                        int p = (l - 1) * D + i;
                        activation = m[(size_t)v * D + i] + relu(h[(size_t)v * D + i] + root[p]) / (c.out_deg[v] + 1);
                        activation = (activation - bnm[p]) / bn_sqrt_var[p] * bnw[p] + bnb[p];
                        activation = relu(activation);
                    }
                    for (int o = 0; o < D; o++)
                    {
                        float addend = activation * cw[((size_t)l * D + o) * D + i];
                        acc[o] = addend + ((i == 0) ? cb[l * D + o] : acc[o]);
                    }
                }
                memcpy(h + (size_t)v * D, acc, sizeof(acc));
            }
            /* message passing: m_v += norm * relu(p_u + edge_embed)  (GCN/src/message_passing.cc:141-170) */
            memset(m, 0, sizeof(float) * (size_t)n * D);
            const float* ee_l = ee_w + (size_t)l * ED_PER_LAYER * D;
            for (int k = 0; k < e; k++)
            {
                int ei = c.order[k], u = edges[2 * ei], v = edges[2 * ei + 1];
                float norm = dis[u] * dis[v];
                for (int d = 0; d < D; d++)
                {
                    float total = edge_embed(ee_l, attrs + 3 * ei, d, D) + h[(size_t)u * D + d];
                    m[(size_t)v * D + d] += norm * relu(total);
                }
            }
        }
        /* finalize: finish layer 4 without relu, mean over nodes in node order, Linear 100->1
         * (GCN/src/finalize.cc:39-115, linear_input_stationary) */
        for (int d = 0; d < D; d++)
        {
            int p = (L - 1) * D + d;
            float sums = 0.0f;
            for (int v = 0; v < n; v++)
            {
                float activation = m[(size_t)v * D + d];
                activation += relu(h[(size_t)v * D + d] + root[p]) / (c.out_deg[v] + 1);
                activation = (activation - bnm[p]) / bn_sqrt_var[p] * bnw[p] + bnb[p];
                sums += activation;
            }
            h_graph[d] = sums / n;
        }
        linear_is2(h_graph, pw, pb, out + g, D, 1, 0);

        nodes_offset += n; edges_offset += e;
    }
    free(h); free(m); free(dis); csr_free(&c);
}

/* ============================================================================================
 * PNA  (PNA/src/PNA_compute.cc:44-98)
 * ========================================================================================== */
void oracle_PNA_compute_graphs(
    int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* reload_weights,
    float* out, const int32_t* node_feature_in, const int32_t* edge_list_in,
    const float* node_embedding_weight_in, const float* node_conv_weights_in, const float* node_conv_bias_in,
    const float* graph_mlp_1_weights_in, const float* graph_mlp_1_bias_in,
    const float* graph_mlp_2_weights_in, const float* graph_mlp_2_bias_in,
    const float* graph_mlp_3_weights_in, const float* graph_mlp_3_bias_in,
    const float* avg_deg_in)
{
    enum { D = 80, L = 4, M1 = 40, M2 = 20, MEAN = 0, MIN = 1, MAX = 2, STD = 3 };
    /* ap_fixed_max / ap_fixed_min of <16,6> (PNA/src/util.h:34-46) */
    const float fm_max = (float)((1 << 5) - (1.0 / (1 << 10)));
    const float fm_min = (float)(-(1 << 5));
    int max_n = max_i32(nums_of_nodes, num_graphs), max_e = max_i32(nums_of_edges, num_graphs);
    csr_t c; csr_alloc(&c, max_n, max_e);
    float* h = (float*)malloc(sizeof(float) * (size_t)(max_n + 1) * D);
    float* msg = (float*)malloc(sizeof(float) * (size_t)(max_n + 1) * D * 4);   /* [v][d][aggr] */
    float acc[D], h_graph[D], o1[M1], o2[M2];

    long nodes_offset = 0, edges_offset = 0;
    int weights_ndx = -1;
    for (int g = 0; g < num_graphs; g++)
    {
        int n = nums_of_nodes[g], e = nums_of_edges[g];
        if (reload_weights[g]) weights_ndx++;
        const float* ne_w = node_embedding_weight_in + (size_t)weights_ndx * ND_FEATURE_TOTAL * D;
        const float* cw = node_conv_weights_in + (size_t)weights_ndx * L * D * 12 * D;   /* [l][out][scaler][aggr][in] */
        const float* cb = node_conv_bias_in + (size_t)weights_ndx * L * D;
        const float avg_deg = avg_deg_in[weights_ndx];
        const int32_t* edges = edge_list_in + 2 * edges_offset;

        csr_build(&c, edges, n, e);
        embed_nodes(h, node_feature_in + nodes_offset * ND_FEATURE, ne_w, n, D);

        for (int l = 0; l < L; l++)
        {
            /* message passing: four planes per (v, d)  (PNA/src/message_passing.cc:88-147) */
            for (size_t i = 0; i < (size_t)n * D; i++)
            {
                msg[4 * i + MEAN] = 0.0f; msg[4 * i + STD] = 0.0f; msg[4 * i + MIN] = fm_max; msg[4 * i + MAX] = fm_min;
            }
            for (int k = 0; k < e; k++)
            {
                int ei = c.order[k], u = edges[2 * ei], v = edges[2 * ei + 1];
                for (int d = 0; d < D; d++)
                {
                    float x = h[(size_t)u * D + d];
                    float* p = msg + 4 * ((size_t)v * D + d);
                    p[MEAN] += x;
                    p[STD] += x * x;
                    if (x < p[MIN]) p[MIN] = x;
                    if (x > p[MAX]) p[MAX] = x;
                }
            }
            /* node transform (PNA/src/node_embedding.cc:106-215) */
            for (int v = 0; v < n; v++)
            {
                int in_degree = c.in_deg[v];
                if (in_degree == 0) in_degree = 1;
                float log_degree = logf((float)(c.out_deg[v] + 1));           /* PNA/src/load_inputs.cc:105 */
                for (int i = 0; i < D; i++)
                {
                    const float* p = msg + 4 * ((size_t)v * D + i);
                    float sum = p[MEAN], sum_squares = p[STD], mn = p[MIN], mx = p[MAX];
                    float mean = sum / in_degree;
                    float stddev = sqrtf(relu((sum_squares / in_degree) - (mean * mean)));
                    float t = log_degree / avg_deg;
                    float scale = avg_deg / log_degree;
                    if (scale == 0) scale = 1;
                    for (int o = 0; o < D; o++)
                    {
                        const float* w = cw + (((size_t)l * D + o) * 12) * D + i;    /* w[(scaler*4+aggr)*D] */
#define W(s, a) w[(size_t)((s) * 4 + (a)) * D]
                        float addend =
                            ((mean * W(0, MEAN) + stddev * W(0, STD)) + (mn * W(0, MIN) + mx * W(0, MAX)))
                            + ((((mean * W(1, MEAN) + stddev * W(1, STD)) + (mn * W(1, MIN) + mx * W(1, MAX))) * t)
                               + (((mean * W(2, MEAN) + stddev * W(2, STD)) + (mn * W(2, MIN) + mx * W(2, MAX))) * scale));
#undef W
                        acc[o] = addend + ((i == 0) ? cb[l * D + o] : acc[o]);
                    }
                }
                for (int o = 0; o < D; o++) h[(size_t)v * D + o] = h[(size_t)v * D + o] + relu(acc[o]);
            }
        }
        /* finalize: mean pool, 80 -> 40 relu -> 20 relu -> 1  (PNA/src/finalize.cc:14-57) */
        mean_pool_pairs(h, n, D, h_graph);
        linear_os(h_graph, graph_mlp_1_weights_in + (size_t)weights_ndx * M1 * D, graph_mlp_1_bias_in + (size_t)weights_ndx * M1, o1, D, M1, 1);
        linear_is2(o1, graph_mlp_2_weights_in + (size_t)weights_ndx * M2 * M1, graph_mlp_2_bias_in + (size_t)weights_ndx * M2, o2, M1, M2, 1);
        linear_os(o2, graph_mlp_3_weights_in + (size_t)weights_ndx * M2, graph_mlp_3_bias_in + (size_t)weights_ndx, out + g, M2, 1, 0);

        nodes_offset += n; edges_offset += e;
    }
    free(h); free(msg); csr_free(&c);
}

/* ============================================================================================
 * DGN  (DGN/src/DGN_compute.cc:36-103)
 * ========================================================================================== */
void oracle_DGN_compute_graphs(
    int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* reload_weights,
    float* out, const int32_t* node_feature_in, const float* node_eigen_in, const int32_t* edge_list_in,
    const float* embedding_h_atom_embedding_list_weights_in,
    const float* layers_posttrans_fully_connected_0_linear_weight_in,
    const float* layers_posttrans_fully_connected_0_linear_bias_in,
    const float* MLP_layer_FC_layers_0_weight_in, const float* MLP_layer_FC_layers_0_bias_in,
    const float* MLP_layer_FC_layers_1_weight_in, const float* MLP_layer_FC_layers_1_bias_in,
    const float* MLP_layer_FC_layers_2_weight_in, const float* MLP_layer_FC_layers_2_bias_in)
{
    enum { D = 100, L = 4, F0 = 50, F1 = 25 };
    const float wt_eps = (float)(1.0 / (1 << 13));      /* ap_fixed_epsilon of <16,3> (DGN/src/dcl.h:54-55) */
    int max_n = max_i32(nums_of_nodes, num_graphs), max_e = max_i32(nums_of_edges, num_graphs);
    csr_t c; csr_alloc(&c, max_n, max_e);
    float* h = (float*)malloc(sizeof(float) * (size_t)(max_n + 1) * D);
    float* m0 = (float*)malloc(sizeof(float) * (size_t)(max_n + 1) * D);
    float* m1 = (float*)malloc(sizeof(float) * (size_t)(max_n + 1) * D);
    float* eig_w = (float*)malloc(sizeof(float) * (size_t)(max_e + 1));
    float* eig_abssums = (float*)malloc(sizeof(float) * (size_t)(max_n + 1));
    float* eigw_sums = (float*)malloc(sizeof(float) * (size_t)(max_n + 1));
    float acc[D], h_graph[D], o0[F0], o1[F1];

    long nodes_offset = 0, edges_offset = 0;
    int weights_ndx = -1;
    for (int g = 0; g < num_graphs; g++)
    {
        int n = nums_of_nodes[g], e = nums_of_edges[g];
        if (reload_weights[g]) weights_ndx++;
        const float* emb = embedding_h_atom_embedding_list_weights_in + (size_t)weights_ndx * 9 * 119 * D;
        const float* lw = layers_posttrans_fully_connected_0_linear_weight_in + (size_t)weights_ndx * L * D * 2 * D;
        const float* lb = layers_posttrans_fully_connected_0_linear_bias_in + (size_t)weights_ndx * L * D;
        const int32_t* edges = edge_list_in + 2 * edges_offset;
        const int32_t* feat = node_feature_in + nodes_offset * ND_FEATURE;
        const float* eig = node_eigen_in + nodes_offset * 4;

        csr_build(&c, edges, n, e);
        /* eigenvector weights, accumulated in edge-LIST order  (DGN/src/load_inputs.cc:91-111) */
        for (int i = 0; i < n; i++) { eig_abssums[i] = 0.0f; eigw_sums[i] = 0.0f; }
        for (int i = 0; i < e; i++)
        {
            int u = edges[2 * i], v = edges[2 * i + 1];
            float diff = eig[u * 4 + 1] - eig[v * 4 + 1];
            eig_w[i] = diff;
            eig_abssums[v] += fabsf(diff);
            eigw_sums[v] += diff;
        }
        /* embedding from nine separate [119][100] tables  (DGN/src/load_inputs.cc:114-168) */
        for (int v = 0; v < n; v++)
            for (int d = 0; d < D; d++)
            {
                float s = 0.0f;
                for (int f = 0; f < ND_FEATURE; f++) s += emb[((size_t)f * 119 + feat[v * ND_FEATURE + f]) * D + d];
                h[(size_t)v * D + d] = s;
            }

        for (int l = 0; l < L; l++)
        {
            /* message passing  (DGN/src/message_passing.cc:121-153) */
            memset(m0, 0, sizeof(float) * (size_t)n * D);
            memset(m1, 0, sizeof(float) * (size_t)n * D);
            for (int k = 0; k < e; k++)
            {
                int ei = c.order[k], u = edges[2 * ei], v = edges[2 * ei + 1];
                float w = eig_w[ei];
                for (int d = 0; d < D; d++)
                {
                    float x = h[(size_t)u * D + d];
                    m0[(size_t)v * D + d] += x;
                    m1[(size_t)v * D + d] += x * w;
                }
            }
            /* node transform  (DGN/src/node_embedding.cc:106-183) */
            for (int v = 0; v < n; v++)
            {
                float eig_abssum = eig_abssums[v];
                if (eig_abssum == 0.0) eig_abssum = wt_eps;
                for (int i = 0; i < D; i++)
                {
                    float h_el = h[(size_t)v * D + i];
                    float a1 = m0[(size_t)v * D + i] / c.out_deg[v];
                    float a2 = fabsf((m1[(size_t)v * D + i] - eigw_sums[v] * h_el) / eig_abssum);
                    for (int o = 0; o < D; o++)
                    {
                        const float* w = lw + ((size_t)l * D + o) * 2 * D;
                        float addend = a1 * w[i] + a2 * w[D + i];
                        acc[o] = addend + ((i == 0) ? lb[l * D + o] : acc[o]);
                    }
                }
                for (int o = 0; o < D; o++) h[(size_t)v * D + o] = h[(size_t)v * D + o] + relu(acc[o]);
            }
        }
        /* finalize: mean pool, 100 -> 50 relu -> 25 relu -> 1  (DGN/src/finalize.cc:14-53) */
        mean_pool_pairs(h, n, D, h_graph);
        linear_os(h_graph, MLP_layer_FC_layers_0_weight_in + (size_t)weights_ndx * F0 * D, MLP_layer_FC_layers_0_bias_in + (size_t)weights_ndx * F0, o0, D, F0, 1);
        linear_is2(o0, MLP_layer_FC_layers_1_weight_in + (size_t)weights_ndx * F1 * F0, MLP_layer_FC_layers_1_bias_in + (size_t)weights_ndx * F1, o1, F0, F1, 1);
        linear_os(o1, MLP_layer_FC_layers_2_weight_in + (size_t)weights_ndx * F1, MLP_layer_FC_layers_2_bias_in + (size_t)weights_ndx, out + g, F1, 1, 0);

        nodes_offset += n; edges_offset += e;
    }
    free(h); free(m0); free(m1); free(eig_w); free(eig_abssums); free(eigw_sums); csr_free(&c);
}

/* ============================================================================================
 * GAT  (GAT/src/GAT_compute.cc:47-108)
 * ========================================================================================== */
static int g_gat_node_offset_bug = 1;
void oracle_set_gat_node_offset_bug(int enabled) { g_gat_node_offset_bug = enabled; }

void oracle_GAT_compute_graphs(
    int num_graphs, const int32_t* nums_of_nodes, const int32_t* nums_of_edges, const int32_t* reload_weights,
    float* out, const int32_t* node_feature_in, const int32_t* edge_list_in,
    const float* scoring_fn_target_in, const float* scoring_fn_source_in,
    const float* linear_proj_weights_in, const float* skip_proj_weights_in,
    const float* graph_pred_weights_in, const float* graph_pred_bias_in)
{
    enum { NH = 4, F = 16, L = 5, PE = 4, HF = NH * F };
    /* feature layout follows the reference: [v][dim][head] */
    int max_n = max_i32(nums_of_nodes, num_graphs), max_e = max_i32(nums_of_edges, num_graphs);
    size_t nbuf = (size_t)(max_n + 1);
    float* hproj = (float*)malloc(sizeof(float) * nbuf * HF);
    float* hproj_next = (float*)malloc(sizeof(float) * nbuf * HF);
    float* o_prev = (float*)malloc(sizeof(float) * nbuf * HF);
    float* o_next = (float*)malloc(sizeof(float) * nbuf * HF);
    float* S = (float*)malloc(sizeof(float) * nbuf * NH);      /* scores_source: indexed by the aggregating node */
    float* T = (float*)malloc(sizeof(float) * nbuf * NH);      /* scores_target: indexed by the neighbour */
    float* S_next = (float*)malloc(sizeof(float) * nbuf * NH);
    float* T_next = (float*)malloc(sizeof(float) * nbuf * NH);
    float* emb = (float*)malloc(sizeof(float) * nbuf * F);
    /* in-neighbour lists banked by source PE (u % 4), self loop first  (GAT/src/load_inputs.cc:87-166) */
    int* deg_pe = (int*)malloc(sizeof(int) * nbuf * PE);
    int* ptr_pe = (int*)malloc(sizeof(int) * (nbuf * PE + 1));
    int* nbr = (int*)malloc(sizeof(int) * (size_t)(max_e + max_n + 1));
    int* cursor = (int*)malloc(sizeof(int) * nbuf * PE);
    float h_graph[F];

    long nodes_offset = 0, edges_offset = 0;
    int weights_ndx = -1;
    for (int g = 0; g < num_graphs; g++)
    {
        int n = nums_of_nodes[g], e = nums_of_edges[g];
        if (reload_weights[g]) weights_ndx++;
        const float* a_tgt = scoring_fn_target_in + (size_t)weights_ndx * L * NH * F;
        const float* a_src = scoring_fn_source_in + (size_t)weights_ndx * L * NH * F;
        const float* w_lin = linear_proj_weights_in + (size_t)weights_ndx * L * HF * HF;    /* [l][ho][do][hi][di] */
        const float* w_skip = skip_proj_weights_in + (size_t)weights_ndx * L * HF * HF;
        const float* pw = graph_pred_weights_in + (size_t)weights_ndx * F;
        const float* pb = graph_pred_bias_in + (size_t)weights_ndx;
        const int32_t* edges = edge_list_in + 2 * edges_offset;
        /* the reference passes node_feature_in without the per-graph offset */
        const int32_t* feat = node_feature_in + (g_gat_node_offset_bug ? 0 : nodes_offset * ND_FEATURE);

        /* load_graph: slot (pe, v) holds v's in-neighbours u with u%4 == pe, in edge-list order;
         * the self loop is the first entry of slot (v%4, v). Slots are laid out pe-major. */
        for (int i = 0; i < n * PE; i++) deg_pe[i] = 0;
        for (int v = 0; v < n; v++) deg_pe[(v % PE) * n + v] = 1;
        for (int i = 0; i < e; i++) deg_pe[(edges[2 * i] % PE) * n + edges[2 * i + 1]]++;
        ptr_pe[0] = 0;
        for (int i = 0; i < n * PE; i++) ptr_pe[i + 1] = ptr_pe[i] + deg_pe[i];
        for (int i = 0; i < n * PE; i++) cursor[i] = ptr_pe[i];
        for (int v = 0; v < n; v++) nbr[cursor[(v % PE) * n + v]++] = v;
        for (int i = 0; i < e; i++) { int u = edges[2 * i], v = edges[2 * i + 1]; nbr[cursor[(u % PE) * n + v]++] = u; }

        /* load_input_node_embeddings  (GAT/src/load_inputs.cc:168-227) */
        for (int v = 0; v < n; v++)
        {
            float* proj = hproj + (size_t)v * HF;
            float* op = o_prev + (size_t)v * HF;
            for (int i = 0; i < HF; i++) { proj[i] = 0.0f; op[i] = 0.0f; }
            for (int f = 0; f < ND_FEATURE; f++)
            {
                float x = (float)feat[v * ND_FEATURE + f];
                op[f * NH + 0] = x;
                for (int d = 0; d < F; d++)
                    for (int ho = 0; ho < NH; ho++)
                        proj[d * NH + ho] += x * w_lin[(((size_t)0 * NH + ho) * F + d) * HF + 0 * F + f];
            }
            for (int hd = 0; hd < NH; hd++) { S[v * NH + hd] = 0.0f; T[v * NH + hd] = 0.0f; }
            for (int d = 0; d < F; d++)
                for (int hd = 0; hd < NH; hd++)
                {
                    S[v * NH + hd] += proj[d * NH + hd] * a_src[(0 * NH + hd) * F + d];
                    T[v * NH + hd] += proj[d * NH + hd] * a_tgt[(0 * NH + hd) * F + d];
                }
        }

        for (int l = 0; l < L; l++)
        {
            for (int v = 0; v < n; v++)
            {
                /* message passing: per source bank, then banks added 0..3 and normalised
                 * (GAT/src/message_passing.cc:94-157; conv_layer.cc:158-177) */
                float msg[HF], score_sum[NH];
                for (int i = 0; i < HF; i++) msg[i] = 0.0f;
                for (int hd = 0; hd < NH; hd++) score_sum[hd] = 0.0f;
                for (int pe = 0; pe < PE; pe++)
                {
                    float part[HF], part_sum[NH];
                    for (int i = 0; i < HF; i++) part[i] = 0.0f;
                    for (int hd = 0; hd < NH; hd++) part_sum[hd] = 0.0f;
                    for (int k = ptr_pe[pe * n + v]; k < ptr_pe[pe * n + v + 1]; k++)
                    {
                        int u = nbr[k];
                        float w[NH];
                        for (int hd = 0; hd < NH; hd++)
                        {
                            float score = S[v * NH + hd] + T[u * NH + hd];
                            if (score < 0) score = score * 0.2f;
                            w[hd] = expf(score);
                            part_sum[hd] += w[hd];
                        }
                        for (int d = 0; d < F; d++)
                            for (int hd = 0; hd < NH; hd++)
                                part[d * NH + hd] += w[hd] * hproj[(size_t)u * HF + d * NH + hd];
                    }
                    for (int hd = 0; hd < NH; hd++) score_sum[hd] += part_sum[hd];
                    for (int i = 0; i < HF; i++) msg[i] += part[i];
                }
                for (int d = 0; d < F; d++)
                    for (int hd = 0; hd < NH; hd++) msg[d * NH + hd] /= score_sum[hd];

                const float* op = o_prev + (size_t)v * HF;
                if (l < L - 1)
                {
                    /* node transform: skip projection + ELU, next projection, next scores
                     * (GAT/src/node_embedding.cc:98-271) */
                    float accs[HF];
                    float* on = o_next + (size_t)v * HF;
                    for (int dout = 0; dout < F; dout++)
                    {
                        float o[NH];
                        for (int ho = 0; ho < NH; ho++) o[ho] = msg[dout * NH + ho];
                        for (int di = 0; di < F; di++)
                            for (int ho = 0; ho < NH; ho++)
                                for (int hi = 0; hi < NH; hi++)
                                    o[ho] += op[di * NH + hi] * w_skip[(((size_t)l * NH + ho) * F + dout) * HF + hi * F + di];
                        for (int ho = 0; ho < NH; ho++)
                        {
                            if (o[ho] <= 0) o[ho] = expf(o[ho]) - 1.0f;
                            on[dout * NH + ho] = o[ho];
                        }
                        for (int pd = 0; pd < F; pd++)
                        {
                            float acc[NH];
                            for (int ho = 0; ho < NH; ho++) acc[ho] = (dout != 0) ? accs[pd * NH + ho] : 0.0f;
                            for (int hi = 0; hi < NH; hi++)
                                for (int ho = 0; ho < NH; ho++)
                                    acc[ho] += o[hi] * w_lin[(((size_t)(l + 1) * NH + ho) * F + pd) * HF + hi * F + dout];
                            for (int ho = 0; ho < NH; ho++) accs[pd * NH + ho] = acc[ho];
                        }
                    }
                    float s_acc[NH], t_acc[NH];
                    for (int d = 0; d < F; d++)
                        for (int hd = 0; hd < NH; hd++)
                        {
                            float r = accs[d * NH + hd];
                            hproj_next[(size_t)v * HF + d * NH + hd] = r;
                            float s = 0.0f, t = 0.0f;
                            s += r * a_src[((l + 1) * NH + hd) * F + d];
                            t += r * a_tgt[((l + 1) * NH + hd) * F + d];
                            if (d != 0) { s += s_acc[hd]; t += t_acc[hd]; }
                            s_acc[hd] = s; t_acc[hd] = t;
                        }
                    for (int hd = 0; hd < NH; hd++) { S_next[v * NH + hd] = s_acc[hd]; T_next[v * NH + hd] = t_acc[hd]; }
                }
                else
                {
                    /* finalize: last skip-add, mean over heads  (GAT/src/finalize.cc:46-112) */
                    for (int d = 0; d < F; d++)
                    {
                        float of = 0.0f;
                        for (int hd = 0; hd < NH; hd++) of += msg[d * NH + hd];
                        for (int di = 0; di < F; di++)
                            for (int ho = 0; ho < NH; ho++)
                                for (int hi = 0; hi < NH; hi++)
                                    of += op[di * NH + hi] * w_skip[(((size_t)l * NH + ho) * F + d) * HF + hi * F + di];
                        emb[(size_t)v * F + d] = of / NH;
                    }
                }
            }
            float* tmp;
            tmp = hproj; hproj = hproj_next; hproj_next = tmp;
            tmp = o_prev; o_prev = o_next; o_next = tmp;
            tmp = S; S = S_next; S_next = tmp;
            tmp = T; T = T_next; T_next = tmp;
        }
        mean_pool_pairs(emb, n, F, h_graph);
        linear_os(h_graph, pw, pb, out + g, F, 1, 0);

        nodes_offset += n; edges_offset += e;
    }
    free(hproj); free(hproj_next); free(o_prev); free(o_next); free(S); free(T); free(S_next); free(T_next);
    free(emb); free(deg_pe); free(ptr_pe); free(nbr); free(cursor);
}
