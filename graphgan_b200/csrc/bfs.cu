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
// the order in which the reference's loop would look at the edge -- and let every edge whose
// head is still undiscovered atomicMin its q into claim[head].  The minimum is the reference's
// discoverer, and the winners in q order (a stable compaction) are the next frontier in FIFO
// order.  q keeps growing across levels, so a stale claim can never equal a current q.
//
// One 1024-thread CTA owns one root at a time.  Per level: (A) prefix sums over the frontier,
// (B) claim sweep, (C) winner sweep (records one win bit per frontier edge), (D) prefix sum of
// the winner counts, (E) scatter from the win bits.  Small frontier nodes (degree <= 32) are
// processed one per thread, large ones one per warp.  A visited bitmap in SHARED memory (N bits;
// global scratch when N > ~1.7M) filters already-discovered heads, so the only random global
// accesses are one atomicMin + one 4-byte read per edge into the NEXT level.
// HBM/L2-bound integer work; no tensor cores.
#include "gg_common.cuh"

namespace gg {
namespace {

constexpr int BFS_THREADS = 1024;
constexpr int BFS_WARPS = BFS_THREADS / 32;
constexpr long long BFS_SMEM_BITMAP_MAX_BYTES = 200 * 1024;

__host__ __device__ inline long long bfs_words_per_cta(long long n, long long nnz, bool bitmap_in_smem) {
    const long long bm = bitmap_in_smem ? 0 : (n + 31) / 32;
    return 8 * n + nnz / 32 + bm + 8;
}

// exclusive scan of f(i), i in [0,n), into out[0..n]; returns total (block-wide, all threads).
// Every warp scans one contiguous chunk with a running carry (coalesced, no block barrier inside the loop);
// the 32 chunk totals are scanned once; a second sweep adds the chunk offsets.  Three barriers in total --
// the frontier of a level can have ~N entries, and a barrier per 1024 elements used to dominate the build.
template <typename F>
__device__ unsigned block_exclusive_scan(F f, unsigned *out, unsigned n, unsigned *s_warp, unsigned *s_carry) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned chunk = ((n + BFS_WARPS - 1) / BFS_WARPS + 31u) & ~31u;
    const unsigned lo = wid * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
    unsigned carry = 0;
    for (unsigned base = lo; base < hi; base += 32) {
        const unsigned i = base + lane;
        const unsigned v = (i < hi) ? f(i) : 0u;
        unsigned x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned y = __shfl_up_sync(FULL, x, off);
            if (lane >= off) x += y;
        }
        if (i < hi) out[i] = carry + x - v;
        carry += __shfl_sync(FULL, x, 31);
    }
    if (lane == 0) s_warp[wid] = carry;
    __syncthreads();
    if (wid == 0) {
        const unsigned v = s_warp[lane];
        unsigned t = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned y = __shfl_up_sync(FULL, t, off);
            if (lane >= off) t += y;
        }
        s_warp[lane] = t - v;                 // exclusive offset of every chunk
        if (lane == 31) *s_carry = t;         // grand total
    }
    __syncthreads();
    const unsigned add = s_warp[wid];
    if (add)
        for (unsigned i = lo + lane; i < hi; i += 32) out[i] += add;
    const unsigned total = *s_carry;
    if (threadIdx.x == 0) out[n] = total;
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(BFS_THREADS, 1)
bfs_kernel(long long n_node, long long nnz, const long long *__restrict__ indptr, const int *__restrict__ adj,
           long long n_roots, const int *__restrict__ roots, int *__restrict__ parent, unsigned *__restrict__ scratch,
           int bitmap_in_smem) {
    extern __shared__ unsigned s_bitmap[];
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry, s_nbig;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t N = (size_t)n_node;
    const size_t bm_words = (N + 31) / 32;
    unsigned *claim = scratch + (size_t)blockIdx.x * (size_t)bfs_words_per_cta(n_node, nnz, bitmap_in_smem != 0);
    int *fa = reinterpret_cast<int *>(claim + N);
    int *fb = fa + N;
    unsigned *off = reinterpret_cast<unsigned *>(fb + N);   // [N+1] q numbering
    unsigned *base = off + N + 1;                            // [N+1] winners per frontier node / compaction offsets
    unsigned *woff = base + N + 1;                           // [N+1] first win-mask word of a frontier node
    unsigned *wmask = woff + N + 1;                          // [nnz/32 + N + 1]
    unsigned *big = wmask + nnz / 32 + N + 1;                // [N] frontier indices of nodes with degree > 32
    unsigned *bm = bitmap_in_smem ? s_bitmap : (big + N);

    for (long long r = blockIdx.x; r < n_roots; r += gridDim.x) {
        const int root = roots[r];
        int *par = parent + (size_t)r * N;
        for (size_t i = threadIdx.x; i < N; i += BFS_THREADS) { claim[i] = 0xffffffffu; par[i] = -1; }
        for (size_t i = threadIdx.x; i < bm_words; i += BFS_THREADS) bm[i] = 0u;
        __syncthreads();
        if (threadIdx.x == 0) { bm[root >> 5] |= 1u << (root & 31); fa[0] = root; }
        __syncthreads();
        unsigned nf = 1, level_base = 1;
        int *cur = fa, *nxt = fb;
        while (nf > 0) {
            // A: q numbering and win-mask word numbering = prefix sums over the frontier
            const unsigned n_edges = block_exclusive_scan(
                [&](unsigned i) { const int u = cur[i]; return (unsigned)(indptr[u + 1] - indptr[u]); }, off, nf, s_warp,
                &s_carry);
            block_exclusive_scan(
                [&](unsigned i) { const int u = cur[i]; return (unsigned)((indptr[u + 1] - indptr[u] + 31) >> 5); }, woff, nf,
                s_warp, &s_carry);
            // Work assignment (per sweep): a frontier node of degree <= 32 is handled by ONE THREAD (32 nodes
            // per warp advance in parallel -- most nodes of a power-law graph are small, and a warp per node would
            // spend its time on the dependent cur -> indptr -> adj latency chain); larger nodes are collected in
            // `big` during sweep B and handled by a whole warp, 128 edges in flight.
            if (threadIdx.x == 0) s_nbig = 0;
            __syncthreads();
            // B: every frontier edge with an undiscovered head claims it with its visit order
            for (unsigned i = threadIdx.x; i < nf; i += BFS_THREADS) {
                const int u = cur[i];
                const long long a0 = indptr[u];
                const unsigned dg = (unsigned)(indptr[u + 1] - a0), q0 = level_base + off[i];
                if (dg > 32) { big[atomicAdd(&s_nbig, 1u)] = i; continue; }
                for (unsigned j0 = 0; j0 < dg; j0 += 8) {   // 8 independent adjacency loads in flight per thread
                    int v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = (j0 + k < dg) ? __ldg(adj + a0 + j0 + k) : -1;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (v[k] >= 0 && !((bm[v[k] >> 5] >> (v[k] & 31)) & 1u)) atomicMin(claim + v[k], q0 + j0 + k);
                }
            }
            __syncthreads();
            const unsigned nbig = s_nbig;
            for (unsigned b = wid; b < nbig; b += BFS_WARPS) {
                const unsigned i = big[b];
                const int u = cur[i];
                const long long a0 = indptr[u];
                const unsigned dg = (unsigned)(indptr[u + 1] - a0), q0 = level_base + off[i];
                for (unsigned j0 = 0; j0 < dg; j0 += 128) {
                    int v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const unsigned j = j0 + 32 * k + lane; v[k] = (j < dg) ? __ldg(adj + a0 + j) : -1; }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (v[k] >= 0 && !((bm[v[k] >> 5] >> (v[k] & 31)) & 1u)) atomicMin(claim + v[k], q0 + j0 + 32 * k + lane);
                }
            }
            __syncthreads();
            // C: winners per frontier node, one win bit per edge
            for (unsigned i = threadIdx.x; i < nf; i += BFS_THREADS) {
                const int u = cur[i];
                const long long a0 = indptr[u];
                const unsigned dg = (unsigned)(indptr[u + 1] - a0), q0 = level_base + off[i];
                if (dg > 32) continue;
                unsigned mk = 0;
                for (unsigned j0 = 0; j0 < dg; j0 += 8) {
                    int v[8];
                    unsigned cl[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = (j0 + k < dg) ? __ldg(adj + a0 + j0 + k) : -1;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        cl[k] = (v[k] >= 0 && !((bm[v[k] >> 5] >> (v[k] & 31)) & 1u)) ? __ldcg(claim + v[k]) : 0u;   // q >= 1
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (cl[k] == q0 + j0 + k) mk |= 1u << (j0 + k);
                }
                if (dg) wmask[woff[i]] = mk;
                base[i] = __popc(mk);
            }
            for (unsigned b = wid; b < nbig; b += BFS_WARPS) {
                const unsigned i = big[b];
                const int u = cur[i];
                const long long a0 = indptr[u];
                const unsigned dg = (unsigned)(indptr[u + 1] - a0), q0 = level_base + off[i], w0 = woff[i];
                unsigned c = 0;
                for (unsigned j0 = 0; j0 < dg; j0 += 128) {
                    int v[4];
                    bool cand[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const unsigned j = j0 + 32 * k + lane; v[k] = (j < dg) ? __ldg(adj + a0 + j) : -1; }
#pragma unroll
                    for (int k = 0; k < 4; ++k) cand[k] = v[k] >= 0 && !((bm[v[k] >> 5] >> (v[k] & 31)) & 1u);
                    unsigned cl[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) cl[k] = cand[k] ? __ldcg(claim + v[k]) : 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned mk = __ballot_sync(FULL, cand[k] && cl[k] == q0 + j0 + 32 * k + lane);
                        if (j0 + 32 * k < dg) {
                            if (lane == 0) wmask[w0 + (j0 >> 5) + k] = mk;
                            c += __popc(mk);
                        }
                    }
                }
                if (lane == 0) base[i] = c;
            }
            __syncthreads();
            // D: stable compaction offsets
            const unsigned n_next = block_exclusive_scan([&](unsigned i) { return base[i]; }, base, nf, s_warp, &s_carry);
            // E: winners become children (father = u) and the next frontier, in q order
            for (unsigned i = threadIdx.x; i < nf; i += BFS_THREADS) {
                unsigned o = base[i];
                if (base[i + 1] == o) continue;   // no winner under this node
                const int u = cur[i];
                const long long a0 = indptr[u];
                const unsigned dg = (unsigned)(indptr[u + 1] - a0);
                if (dg > 32) continue;
                unsigned mk = wmask[woff[i]];
                while (mk) {
                    const int j = __ffs(mk) - 1;
                    mk &= mk - 1u;
                    const int v = __ldg(adj + a0 + j);
                    nxt[o++] = v;
                    par[v] = u;
                    atomicOr(bm + (v >> 5), 1u << (v & 31));
                }
            }
            for (unsigned b = wid; b < nbig; b += BFS_WARPS) {
                const unsigned i = big[b];
                unsigned o = base[i];
                if (base[i + 1] == o) continue;
                const int u = cur[i];
                const long long a0 = indptr[u];
                const unsigned dg = (unsigned)(indptr[u + 1] - a0), w0 = woff[i];
                for (unsigned j0 = 0; j0 < dg; j0 += 32) {
                    const unsigned mk = wmask[w0 + (j0 >> 5)];
                    if (mk == 0u) continue;
                    if ((mk >> lane) & 1u) {
                        const int v = __ldg(adj + a0 + j0 + lane);
                        nxt[o + __popc(mk & ((1u << lane) - 1u))] = v;
                        par[v] = u;
                        atomicOr(bm + (v >> 5), 1u << (v & 31));
                    }
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

extern "C" int gg_bfs_scratch_bytes(int64_t n_node, int64_t nnz, int64_t *bytes) {
    GG_REQUIRE(bytes && n_node >= 0 && nnz >= 0, "bad arguments");
    const bool in_smem = (n_node + 31) / 32 * 4 <= gg::BFS_SMEM_BITMAP_MAX_BYTES;
    *bytes = (int64_t)gg::sm_count() * gg::bfs_words_per_cta(n_node, nnz, in_smem) * 4;
    return 0;
}

extern "C" int gg_bfs_build(int64_t n_node, int64_t nnz, const int64_t *indptr, const int32_t *adj, int64_t n_roots,
                            const int32_t *roots, int32_t *parent, void *scratch, int64_t scratch_bytes, void *stream) {
    GG_REQUIRE(indptr && adj && roots && parent && scratch, "null pointer");
    GG_REQUIRE(nnz < 0xfffffff0ll, "too many edges for 32-bit visit numbers");
    if (n_roots == 0 || n_node == 0) return 0;
    const long long bm_bytes = (n_node + 31) / 32 * 4;
    const bool in_smem = bm_bytes <= gg::BFS_SMEM_BITMAP_MAX_BYTES;
    const int64_t per_cta = gg::bfs_words_per_cta(n_node, nnz, in_smem) * 4;
    int64_t ctas = scratch_bytes / per_cta;
    if (ctas > gg::sm_count()) ctas = gg::sm_count();
    if (ctas > n_roots) ctas = n_roots;
    GG_REQUIRE(ctas >= 1, "scratch too small");
    const size_t smem = in_smem ? (size_t)bm_bytes : 0;
    if (smem > 48 * 1024)
        GG_CHECK(cudaFuncSetAttribute(gg::bfs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gg::bfs_kernel<<<(unsigned)ctas, gg::BFS_THREADS, smem, (cudaStream_t)stream>>>(
        n_node, nnz, (const long long *)indptr, adj, n_roots, roots, parent, (unsigned *)scratch, in_smem ? 1 : 0);
    return gg::check_cuda(cudaGetLastError(), "bfs kernel launch");
}
