// bfs.cu -- BFS-tree construction for a batch of roots (sm_100a).
//
// Replaces GraphGAN.construct_trees (reference src/GraphGAN/graph_gan.py:84-108).  The
// reference stores, per root, a dict node -> [father, children...]: O(N) Python objects per
// root and O(N^2) overall, which cannot exist at N >= 1e5.  Here a tree is one int32 parent
// array; the children of `cur` are recovered during the walk as the adjacency entries whose
// father is `cur`, which reproduces the reference's list order exactly (graph_gan.py:102-105).
//
// The father of v must be the FIRST node, in the reference's FIFO order, that has v in its
// adjacency.  Level-synchronous formulation: number every (frontier node, adjacency slot)
// pair of a level consecutively in (frontier order, slot order) -- that number `q` is exactly
// the order in which the reference's loop would look at the edge -- and let every edge
// atomicMin its q into claim[v].  The minimum is the reference's discoverer, and the winners
// sorted by q (a stable compaction) are the next frontier in FIFO order.  q keeps growing
// across levels, so claim[v] < level_base  <=>  "v was discovered earlier": no separate
// visited array is needed and already-discovered nodes can never win again.
//
// One CTA owns one root at a time.  HBM/L2-bound integer work; no tensor cores.
#include "gg_common.cuh"

namespace gg {
namespace {

constexpr int BFS_THREADS = 1024;
constexpr int BFS_WARPS = BFS_THREADS / 32;

// exclusive scan of f(i), i in [0,n), into out[0..n]; returns total (block-wide, all threads).
template <typename F>
__device__ unsigned block_exclusive_scan(F f, unsigned *out, unsigned n, unsigned *s_warp, unsigned *s_carry) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) *s_carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < n; base += BFS_THREADS) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = (i < n) ? f(i) : 0u;
        unsigned x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned y = __shfl_up_sync(FULL, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned t = s_warp[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned y = __shfl_up_sync(FULL, t, off);
                if (lane >= off) t += y;
            }
            s_warp[lane] = t;
        }
        __syncthreads();
        const unsigned before = *s_carry + (wid ? s_warp[wid - 1] : 0u) + (x - v);
        if (i < n) out[i] = before;
        __syncthreads();
        if (threadIdx.x == BFS_THREADS - 1) *s_carry = before + v;
        __syncthreads();
    }
    const unsigned total = *s_carry;
    if (threadIdx.x == 0) out[n] = total;
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(BFS_THREADS, 1)
bfs_kernel(long long n_node, const long long *__restrict__ indptr, const int *__restrict__ adj, long long n_roots,
           const int *__restrict__ roots, int *__restrict__ parent, unsigned *__restrict__ scratch) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t N = (size_t)n_node;
    unsigned *claim = scratch + (size_t)blockIdx.x * (5 * N + 2);
    int *fa = reinterpret_cast<int *>(claim + N);
    int *fb = fa + N;
    unsigned *off = reinterpret_cast<unsigned *>(fb + N);  // [N+1]
    unsigned *base = off + N + 1;                           // [N+1]

    for (long long r = blockIdx.x; r < n_roots; r += gridDim.x) {
        const int root = roots[r];
        int *par = parent + (size_t)r * N;
        for (size_t i = threadIdx.x; i < N; i += BFS_THREADS) { claim[i] = 0xffffffffu; par[i] = -1; }
        __syncthreads();
        if (threadIdx.x == 0) { claim[root] = 0u; fa[0] = root; }
        __syncthreads();
        unsigned nf = 1, level_base = 1;
        int *cur = fa, *nxt = fb;
        while (nf > 0) {
            // A: q numbering = exclusive prefix of the frontier degrees
            const unsigned n_edges = block_exclusive_scan(
                [&](unsigned i) { const int u = cur[i]; return (unsigned)(indptr[u + 1] - indptr[u]); }, off, nf, s_warp,
                &s_carry);
            // B: every frontier edge claims its head with its visit order
            for (unsigned i = wid; i < nf; i += BFS_WARPS) {
                const int u = cur[i];
                const long long a0 = indptr[u];
                const unsigned dg = (unsigned)(indptr[u + 1] - a0), q0 = level_base + off[i];
                for (unsigned j = lane; j < dg; j += 32) atomicMin(claim + adj[a0 + j], q0 + j);
            }
            __syncthreads();
            // C: winners per frontier node
            for (unsigned i = wid; i < nf; i += BFS_WARPS) {
                const int u = cur[i];
                const long long a0 = indptr[u];
                const unsigned dg = (unsigned)(indptr[u + 1] - a0), q0 = level_base + off[i];
                unsigned c = 0;
                for (unsigned j0 = 0; j0 < dg; j0 += 32) {
                    const unsigned j = j0 + lane;
                    const bool win = (j < dg) && (__ldcg(claim + adj[a0 + j]) == q0 + j);
                    c += __popc(__ballot_sync(FULL, win));
                }
                if (lane == 0) base[i] = c;
            }
            __syncthreads();
            // D: stable compaction offsets
            const unsigned n_next = block_exclusive_scan([&](unsigned i) { return base[i]; }, base, nf, s_warp, &s_carry);
            // E: winners become children (father = u) and the next frontier, in q order
            for (unsigned i = wid; i < nf; i += BFS_WARPS) {
                const int u = cur[i];
                const long long a0 = indptr[u];
                const unsigned dg = (unsigned)(indptr[u + 1] - a0), q0 = level_base + off[i];
                unsigned o = base[i];
                for (unsigned j0 = 0; j0 < dg; j0 += 32) {
                    const unsigned j = j0 + lane;
                    int v = -1;
                    bool win = false;
                    if (j < dg) { v = adj[a0 + j]; win = (__ldcg(claim + v) == q0 + j); }
                    const unsigned mk = __ballot_sync(FULL, win);
                    if (win) { nxt[o + __popc(mk & ((1u << lane) - 1u))] = v; par[v] = u; }
                    o += __popc(mk);
                }
            }
            __syncthreads();
            level_base += n_edges;
            nf = n_next;
            int *t = cur; cur = nxt; nxt = t;
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace gg

extern "C" int gg_bfs_scratch_bytes(int64_t n_node, int64_t *bytes) {
    GG_REQUIRE(bytes && n_node >= 0, "bad arguments");
    *bytes = (int64_t)gg::sm_count() * (5 * n_node + 2) * 4;
    return 0;
}

extern "C" int gg_bfs_build(int64_t n_node, const int64_t *indptr, const int32_t *adj, int64_t n_roots,
                            const int32_t *roots, int32_t *parent, void *scratch, int64_t scratch_bytes, void *stream) {
    GG_REQUIRE(indptr && adj && roots && parent && scratch, "null pointer");
    if (n_roots == 0 || n_node == 0) return 0;
    const int64_t per_cta = (5 * n_node + 2) * 4;
    int64_t ctas = scratch_bytes / per_cta;
    if (ctas > gg::sm_count()) ctas = gg::sm_count();
    if (ctas > n_roots) ctas = n_roots;
    GG_REQUIRE(ctas >= 1, "scratch too small");
    gg::bfs_kernel<<<(unsigned)ctas, gg::BFS_THREADS, 0, (cudaStream_t)stream>>>(
        n_node, (const long long *)indptr, adj, n_roots, roots, parent, (unsigned *)scratch);
    return gg::check_cuda(cudaGetLastError(), "bfs kernel launch");
}
