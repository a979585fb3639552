// hub.cu -- per-pass precomputation that removes the walk's redundant work (sm_100a).
//
// The reference recomputes E.E^T + b for EVERY root (graph_gan.py:238) and re-derives the same
// root-step softmax for each of the root's sample_num walks (:260-262).  Two exact reuses:
//
//  * hub_score_kernel: all_score[u, v] for the adjacency of high-degree nodes u.  A score does
//    not depend on the root, only the candidate SET does (children of u in that root's tree),
//    so one pass over a hub's neighbour rows serves every walk that ever stands on u.
//  * root_cdf_kernel: the root step's candidate list is tree[root][1:] = all neighbours of the
//    root for every walk of that root, so its normalised CDF is built once per root and each
//    walk only draws u and inverts it (walk.cu: cdf_search).
//
// Both produce exactly the floats the on-demand path produces (same canonical arithmetic),
// so sampled indices are unchanged; tests run the walk with and without them.
#include "walk_common.cuh"

namespace gg {
namespace {

template <int CPL>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
hub_score_kernel(long long n_tiles, const int *__restrict__ tile_node, const long long *__restrict__ tile_begin,
                 int tile_edges, const long long *__restrict__ indptr, const int *__restrict__ adj,
                 const float *__restrict__ emb, const float *__restrict__ bias, int ld, float *__restrict__ edge_score) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = warp; t < n_tiles; t += nwarps) {
        const int u = tile_node[t];
        const long long e0 = tile_begin[t], a1 = indptr[u + 1];
        const int n = (int)((a1 - e0) < tile_edges ? (a1 - e0) : tile_edges);
        float4 c4[CPL];
        load_row<CPL>(emb, ld, u, lane & 7, c4);
        score_edges<CPL>(emb, bias, ld, c4, adj, e0, n, edge_score + e0, u, lane);
    }
}

template <int CPL>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
root_cdf_kernel(const __grid_constant__ gg_walk_desc d, float *__restrict__ root_sc, double *__restrict__ root_q) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long slot = warp; slot < d.n_roots; slot += nwarps) {
        const int root = d.roots[slot];
        const long long a0 = d.indptr[root], a1 = d.indptr[root + 1], o = d.rq_ptr[slot];
        const int n = (int)(a1 - a0);
        if (n == 0) continue;
        float *sc = root_sc + o;
        if (d.edge_score && n >= d.hub_threshold) {
            for (int i = lane; i < n; i += 32) sc[i] = __ldg(d.edge_score + a0 + i);
            __syncwarp();
        } else {
            float4 c4[CPL];
            load_row<CPL>(d.emb, d.ld, root, lane & 7, c4);
            score_edges<CPL>(d.emb, d.bias, d.ld, c4, d.adj, a0, n, sc, root, lane);
        }
        cdf_store(sc, n, root_q + o, lane);
    }
}

}  // namespace
}  // namespace gg

extern "C" int gg_hub_scores(int64_t n_tiles, const int32_t *tile_node, const int64_t *tile_begin, int32_t tile_edges,
                             const int64_t *indptr, const int32_t *adj, const float *emb, const float *bias, int32_t ld,
                             float *edge_score, void *stream) {
    if (n_tiles == 0) return 0;
    GG_REQUIRE(tile_node && tile_begin && indptr && adj && emb && bias && edge_score, "null pointer");
    GG_REQUIRE(tile_edges > 0 && ld > 0 && ld % 32 == 0, "bad tile_edges / ld");
    long long blocks = (n_tiles + gg::WARPS_PER_CTA - 1) / gg::WARPS_PER_CTA;
    const long long cap = (long long)gg::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define GG_LAUNCH(C)                                                                                           \
    gg::hub_score_kernel<C><<<(unsigned)blocks, gg::WARPS_PER_CTA * 32, 0, st>>>(                              \
        n_tiles, tile_node, (const long long *)tile_begin, tile_edges, (const long long *)indptr, adj, emb, bias, ld, \
        edge_score)
    switch (ld / 32) {
        case 1: GG_LAUNCH(1); break;
        case 2: GG_LAUNCH(2); break;
        case 4: GG_LAUNCH(4); break;
        case 8: GG_LAUNCH(8); break;
        default: gg::set_error("gg_hub_scores: unsupported ld %d (supported: 32, 64, 128, 256)", ld); return 2;
    }
#undef GG_LAUNCH
    return gg::check_cuda(cudaGetLastError(), "hub score kernel launch");
}

extern "C" int gg_root_cdf(const gg_walk_desc *dp, float *root_sc, double *root_q, void *stream) {
    GG_REQUIRE(dp, "null descriptor");
    if (dp->n_roots == 0) return 0;
    GG_REQUIRE(root_sc && root_q, "null pointer");
    const gg_walk_desc &d = *dp;
    GG_REQUIRE(d.roots && d.indptr && d.adj && d.emb && d.bias && d.rq_ptr, "null pointer in descriptor");
    GG_REQUIRE(d.ld > 0 && d.ld % 32 == 0, "ld must be a positive multiple of 32");
    if (d.n_roots == 0) return 0;
    long long blocks = (d.n_roots + gg::WARPS_PER_CTA - 1) / gg::WARPS_PER_CTA;
    const long long cap = (long long)gg::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    switch (d.ld / 32) {
        case 1: gg::root_cdf_kernel<1><<<(unsigned)blocks, gg::WARPS_PER_CTA * 32, 0, st>>>(d, root_sc, root_q); break;
        case 2: gg::root_cdf_kernel<2><<<(unsigned)blocks, gg::WARPS_PER_CTA * 32, 0, st>>>(d, root_sc, root_q); break;
        case 4: gg::root_cdf_kernel<4><<<(unsigned)blocks, gg::WARPS_PER_CTA * 32, 0, st>>>(d, root_sc, root_q); break;
        case 8: gg::root_cdf_kernel<8><<<(unsigned)blocks, gg::WARPS_PER_CTA * 32, 0, st>>>(d, root_sc, root_q); break;
        default: gg::set_error("gg_root_cdf: unsupported ld %d (supported: 32, 64, 128, 256)", d.ld); return 2;
    }
    return gg::check_cuda(cudaGetLastError(), "root cdf kernel launch");
}
