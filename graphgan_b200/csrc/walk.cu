// walk.cu -- K1, the graph-softmax walk sampler (sm_100a).
//
// Replaces GraphGAN.sample (reference src/GraphGAN/graph_gan.py:225-270) for a whole batch of
// roots: one warp owns one walk at a time (persistent CTAs pull walk ids from a global
// counter).  Per step the warp
//   1. enumerates the candidate list [father] + children(cur) from the walk CSR and the
//      root's BFS parent array (children = adjacency entries whose father is cur, in
//      adjacency order == the reference's list order, graph_gan.py:96,102-105);
//   2. scores the candidates on demand: generator.all_score[cur, cand] = e_cur . e_cand + b_cand
//      (generator.py:21) -- four 8-lane groups, each streaming one embedding row per
//      LDG.128 quartet (one full 128 B line per group per instruction);
//   3. softmax (utils.py:131-133) + float64 CDF + inverse-CDF draw (numpy legacy
//      RandomState.choice, called at graph_gan.py:262) with warp shuffles.
// The arithmetic is the canonical sequence of DESIGN.md section 3 == oracle/gg_oracle.c.
#include <string.h>

#include "walk_common.cuh"

namespace gg {
namespace {

struct Rng {
    int mode;
    uint32_t k0, k1, tag;
    const double *stream;
    long long n_stream;
    long long cursor;  // GG_RNG_STREAM only
    int exhausted;
    __device__ __forceinline__ double draw(uint32_t root, uint32_t walk, uint32_t step) {
        if (mode == GG_RNG_PHILOX) {
            uint32_t a, b;
            philox4x32_10(root, walk, step, tag, k0, k1, a, b);
            return u53(a, b);
        }
        if (cursor >= n_stream) { exhausted = 1; return 0.0; }
        return stream[cursor++];
    }
};

// children of `cur` among its walk-CSR entries [a0, a1): the set bits of the root's tree row `tb` (csrc/bfs.cu), in
// entry order == adjacency order == the reference's list order (graph_gan.py:102-105).  One coalesced load brings 32
// bitmap words (1024 entries); only words with a set bit touch adj[] / edge_score[], U of them in flight.
template <int U>
__device__ __forceinline__ void enumerate_children(const gg_walk_desc &d, const uint32_t *__restrict__ tb, long long a0,
                                                   long long a1, bool cached, int *ids, float *sc, int lane, int &n, float &m,
                                                   Stage &stg) {
    if (a1 <= a0) return;
    const unsigned lt = (1u << lane) - 1u;
    if (a1 - a0 <= 64) {
        // short list (the common case): adjacency entries and their bitmap words are loaded together -- the step's
        // dependent chain is indptr -> {bits, adj} -> rows instead of indptr -> bits -> adj -> rows
        const long long e0 = a0 + lane, e1 = a0 + 32 + lane;
        const bool in0 = e0 < a1, in1 = e1 < a1;
        const unsigned w0 = in0 ? __ldg(tb + (e0 >> 5)) : 0u, w1 = in1 ? __ldg(tb + (e1 >> 5)) : 0u;
        const int v0 = in0 ? __ldg(d.adj + e0) : -1, v1 = in1 ? __ldg(d.adj + e1) : -1;
        const float c0 = (cached && in0) ? __ldg(d.edge_score + e0) : 0.0f, c1 = (cached && in1) ? __ldg(d.edge_score + e1) : 0.0f;
        const bool s0 = in0 && ((w0 >> (e0 & 31)) & 1u), s1 = in1 && ((w1 >> (e1 & 31)) & 1u);
        const unsigned m0 = __ballot_sync(FULL, s0), m1 = __ballot_sync(FULL, s1);
        if (s0) {
            const int pos = n + __popc(m0 & lt);
            ids[pos] = v0;
            if (cached) { sc[pos] = c0; m = fmaxf(m, c0); }
        }
        n += __popc(m0);
        if (s1) {
            const int pos = n + __popc(m1 & lt);
            ids[pos] = v1;
            if (cached) { sc[pos] = c1; m = fmaxf(m, c1); }
        }
        n += __popc(m1);
        return;
    }
    const long long wfirst = a0 >> 5, wlast = (a1 - 1) >> 5;
    if (cached && stg.on && (a1 - a0 + 1) > SC_CAP) {
        // ---- hub list, TMA staged: the list is longer than the warp's shared score buffer, so its scores go to the
        // global scratch and the buffer is idle: each 512-entry block of adj[] and edge_score[] (contiguous, 128-B
        // aligned) is brought in by ONE elected lane with two cp.async.bulk copies completing on the warp's mbarrier --
        // one round trip per 512 entries instead of one per 128 -- while the lanes fetch the block's 16 bitmap words.
        int *s_adj = reinterpret_cast<int *>(stg.buf);
        float *s_cs = stg.buf + STAGE_ENTRIES;
        const long long a1r = (a1 + 3) & ~3ll;                      // copy sizes are multiples of 16 bytes (arrays are padded)
        for (long long eb = wfirst << 5; eb < a1; eb += STAGE_ENTRIES) {
            const unsigned bytes = (unsigned)(((eb + STAGE_ENTRIES < a1r) ? (long long)STAGE_ENTRIES : (a1r - eb)) * 4);
            if (lane == 0) {
                mbar_expect_tx(stg.bar, 2 * bytes);
                bulk_g2s(s_adj, d.adj + eb, bytes, stg.bar);
                bulk_g2s(s_cs, d.edge_score + eb, bytes, stg.bar);
            }
            const long long wi = (eb >> 5) + lane;
            unsigned word = (lane < STAGE_ENTRIES / 32 && wi <= wlast) ? __ldg(tb + wi) : 0u;
            if (wi == wfirst) word &= 0xffffffffu << (a0 & 31);
            if (wi == wlast && (a1 & 31)) word &= (1u << (a1 & 31)) - 1u;
            unsigned nz = __ballot_sync(FULL, word != 0u);
            mbar_wait(stg.bar, stg.phase);
            stg.phase ^= 1u;
            while (nz) {
                const int j = __ffs(nz) - 1;
                nz &= nz - 1u;
                const unsigned wv = __shfl_sync(FULL, word, j);
                if ((wv >> lane) & 1u) {
                    const int pos = n + __popc(wv & lt);
                    const float cs = s_cs[32 * j + lane];
                    ids[pos] = s_adj[32 * j + lane];
                    sc[pos] = cs;
                    m = fmaxf(m, cs);
                }
                n += __popc(wv);
            }
            __syncwarp();                                           // every lane is done with the block before it is overwritten
        }
        return;
    }
    for (long long wb = wfirst; wb <= wlast; wb += 32) {
        const long long wi = wb + lane;
        unsigned word = (wi <= wlast) ? __ldg(tb + wi) : 0u;
        if (wi == wfirst) word &= 0xffffffffu << (a0 & 31);
        if (wi == wlast && (a1 & 31)) word &= (1u << (a1 & 31)) - 1u;
        unsigned nz = __ballot_sync(FULL, word != 0u);
        while (nz) {
            int jw[U], v[U];
            unsigned wv[U];
            float cs[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                jw[k] = nz ? (__ffs(nz) - 1) : -1;
                nz &= nz - 1u;                              // 0 & 0xffffffff == 0: stays empty
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const unsigned x = __shfl_sync(FULL, word, jw[k] & 31);
                wv[k] = (jw[k] >= 0) ? x : 0u;
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const bool isc = (wv[k] >> lane) & 1u;
                const long long e = ((wb + jw[k]) << 5) + lane;
                v[k] = isc ? __ldg(d.adj + e) : -1;
                cs[k] = (cached && isc) ? __ldg(d.edge_score + e) : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                if (jw[k] < 0) break;                       // warp-uniform
                if ((wv[k] >> lane) & 1u) {
                    const int pos = n + __popc(wv[k] & lt);
                    ids[pos] = v[k];
                    if (cached) { sc[pos] = cs[k]; m = fmaxf(m, cs[k]); }
                }
                n += __popc(wv[k]);
            }
        }
    }
}

// Candidate list of `cur` in the tree of the root whose tree row is `tb` (graph_gan.py:250-259):
// [father] + children in adjacency order, with scores all_score[cur, cand] (generator.py:21) -- cached hub
// scores or the on-demand canonical dot -- and their max.  Warp-cooperative; results are warp-uniform.
template <int CPL, int U>
__device__ __forceinline__ void build_list(const gg_walk_desc &d, const uint32_t *__restrict__ tb, int cur, int prev,
                                           bool inc_father, int *s_ids, float *s_sc, int *g_ids, float *g_sc, int lane,
                                           int &n_out, float &m_out, int *&ids_out, float *&sc_out,
                                           unsigned long long &rows_gathered, unsigned int (&cyc)[7], Stage &stg) {
    const long long a0 = d.indptr[cur], a1 = d.indptr[cur + 1];
    const bool cached = d.edge_score && (a1 - a0) >= d.hub_threshold;  // scores precomputed per pass
    int *ids = (a1 - a0 + 1) <= ID_CAP ? s_ids : g_ids;
    float *sc = (a1 - a0 + 1) <= SC_CAP ? s_sc : g_sc;
    int n = 0;
    if (inc_father) { if (lane == 0) ids[0] = prev; n = 1; }
    float m = -INFINITY;   // running max of the cached scores (lane local)
    const long long t_e = clock64();
    // (16 tiles in flight for hub adjacency was measured: the extra registers spill and the kernel gets slower)
    enumerate_children<U>(d, tb, a0, a1, cached, ids, sc, lane, n, m, stg);
    __syncwarp();
    const long long t_s = clock64();
    cyc[0] += (unsigned int)(t_s - t_e);
    n_out = n; ids_out = ids; sc_out = sc; m_out = m;
    // n == 1: softmax = [1.0], cdf = [1.0], and 1.0 > u for every uniform u in [0, 1): the draw is index 0
    // whatever the score is, so neither the score nor the CDF is computed (leaves of the BFS tree: [father])
    if (n <= 1) return;
    if (!cached || inc_father) {
        float4 c4[CPL];
        load_row<CPL>(d.emb, d.ld, cur, lane & 7, c4);
        score_list<CPL>(d.emb, d.bias, d.ld, c4, ids, sc, cached ? 1 : n, cur, lane);
        rows_gathered += 1u + (unsigned)(cached ? 1 : n);
    }
    if (cached) {
        m = warp_max(m);
        if (inc_father) m = fmaxf(m, sc[0]);
    } else {
        m = list_max(sc, n, lane);
    }
    m_out = m;
    cyc[1] += (unsigned int)(clock64() - t_s);
}

// One complete walk, executed by a full warp.  Returns the status.
// a (root, depth-1 child) pair gets a shared CDF (step1_cdf_kernel) when at least this many walks picked it
#ifndef GG_S1_MIN_WALKS
#define GG_S1_MIN_WALKS 1
#endif
constexpr int S1_MIN_WALKS = GG_S1_MIN_WALKS;

// Does the pair (root slot, i-th neighbour) at `s1pos` get a shared CDF from step1_cdf_kernel?  With S1_MIN_WALKS = 1: every
// pair that was picked (measured: hub lists are long and the builder's queue starts the longest first; 2 -- pairs picked
// once stay inside their walk -- and "2, or 1 when the child is score-cached" were both slower with the flat steps).
__device__ __forceinline__ bool s1_is_shared(const gg_walk_desc &d, long long s1pos) {
    return __ldg(d.s1_cnt + s1pos) >= S1_MIN_WALKS;
}

// where a walk (re)starts: a fresh walk stands on its root; a walk handed over by the level-synchronous steps
// (flat_*_kernel below) continues from the node it reached (step == choices made so far)
struct WalkState {
    int cur, prev, step, fedge, suml;
};

template <int CPL>
__device__ __forceinline__ int walk_one(const gg_walk_desc &d, Rng &rng, int slot, uint32_t k, long long w,
                                        int *s_ids, float *s_sc, int *g_ids, float *g_sc, int lane,
                                        unsigned long long &raw_steps, unsigned long long &raw_suml,
                                        unsigned long long &overflow, unsigned long long &rows_gathered,
                                        unsigned int (&cyc)[7], Stage &stg, const WalkState *from = nullptr) {
    const int root = d.roots[slot];
    const uint32_t *tb = d.tree_bits + (size_t)slot * (size_t)d.tree_words;
    int cur = root, prev = -1, step = 0, fedge = -1, plen = 0;
    int steps = 0, suml = 0, status = GG_NOTRUN, sample = -1;
    int32_t *prow = (d.max_path > 0 && d.paths) ? d.paths + (size_t)w * (size_t)d.max_path : nullptr;
    if (from) {
        cur = from->cur; prev = from->prev; step = from->step; fedge = from->fedge; steps = from->step; suml = from->suml;
        plen = step + 1;
    } else {
        if (prow && lane == 0) prow[0] = cur;
        plen = 1;
    }
    const int steps_in = steps, suml_in = suml;

    const long long t_walk = clock64();
    for (;;) {
        const long long t_step = clock64();
        const long long a0 = d.indptr[cur], a1 = d.indptr[cur + 1];
        int n, idx, nxt;
        bool inc_father = false;
        long long s1pos = -1;     // slice of the depth-1 cache, when this (root, child) pair was picked by >= 2 walks
        if (step == 1 && d.s1_q) {
            s1pos = __ldg(d.rq_ptr + slot) + (fedge - d.indptr[root]);
            if (!s1_is_shared(d, s1pos)) s1pos = -1;
        }
        if (step == 0 && d.root_q) {
            // ---- root step from the per-root CDF (hub.cu: root_cdf_kernel): every walk of a root
            // sees the same candidate list tree[root][1:] and the same scores, so the softmax/CDF
            // is computed once per root per pass and each walk only inverts it.
            n = (int)(a1 - a0);
            if (n == 0) { status = GG_VOID; break; }  // graph_gan.py:252-253
            const double u = rng.draw((uint32_t)root, k, 0u);
            if (rng.exhausted) { status = GG_NOTRUN; break; }
            idx = d.first_idx ? __ldg(d.first_idx + w) : cdf_search(d.root_q + __ldg(d.rq_ptr + slot), n, u);
            nxt = __ldg(d.adj + a0 + idx);
        } else if (s1pos >= 0) {
            // ---- depth-1 step from the per-(root, child) CDF (step1_cdf_kernel): walks of a root that picked the
            // same child share one candidate list; it was built once, each walk only inverts it
            n = __ldg(d.s1_n + s1pos);
            if (n == 0) { status = GG_VOID; break; }  // graph_gan.py:255-257
            inc_father = !d.for_d && !((d.d1_bits[fedge >> 5] >> (fedge & 31)) & 1u);
            const double u = rng.draw((uint32_t)root, k, 1u);
            const long long o = __ldg(d.s1_ptr + s1pos);
            idx = (n == 1) ? 0 : cdf_search_raw(d.s1_q + o, n, u);
            nxt = __ldg(d.s1_ids + o + idx);
        } else {
            // ---- candidate list (graph_gan.py:250-259) + scores
            inc_father = step > 0;
            if (d.for_d && step == 1) inc_father = false;
            if (!d.for_d && step == 1 && ((d.d1_bits[fedge >> 5] >> (fedge & 31)) & 1u)) inc_father = false;
            int *ids; float *sc; float m;
            build_list<CPL, UNR>(d, tb, cur, prev, inc_father, s_ids, s_sc, g_ids, g_sc, lane, n, m, ids, sc, rows_gathered, cyc, stg);
            if (n == 0) { status = GG_VOID; break; }  // graph_gan.py:252-257

            // ---- softmax + inverse CDF (utils.py:131-133, np.random.choice at graph_gan.py:262)
            const long long t_c = clock64();
            const double u = rng.draw((uint32_t)root, k, (uint32_t)step);   // the stream mode consumes it regardless
            if (rng.exhausted) { status = GG_NOTRUN; break; }
            idx = (n == 1) ? 0 : choose_index(sc, n, m, u, lane, sc != s_sc ? reinterpret_cast<double *>(s_sc) : nullptr);
            nxt = ids[idx];
            __syncwarp();
            cyc[2] += (unsigned int)(clock64() - t_c);
        }
        cyc[step == 0 ? 3 : (step == 1 ? 4 : 5)] += (unsigned int)(clock64() - t_step);   // (dynamic index: keeps the counters in local memory, off the register budget)
        if (step == 0) fedge = (int)(a0 + idx);  // every walk-CSR neighbour of the root is its child
        if (prow && lane == 0 && plen < d.max_path) prow[plen] = nxt;
        ++plen;
        ++steps; suml += n;
        if (inc_father && idx == 0) { sample = cur; status = GG_DONE; break; }  // graph_gan.py:264-266
        prev = cur; cur = nxt; ++step;
    }
    cyc[6] += (unsigned int)(clock64() - t_walk);
    raw_steps += (unsigned)(steps - steps_in); raw_suml += (unsigned)(suml - suml_in);
    if (lane == 0) {
        d.samples[w] = sample;
        d.status[w] = status;
        d.first_edge[w] = fedge;
        d.wsteps[w] = steps;
        d.wsuml[w] = suml;
        if (d.path_len) d.path_len[w] = (status == GG_DONE) ? plen : 0;
    }
    if (status == GG_DONE && d.max_path > 0 && plen > d.max_path) overflow += 1;
    return status;
}

// ---------------------------------------------------------------- depth-1 reuse (Philox mode)
// root_step_kernel: the root step of every walk (one thread per walk inverts the root's CDF) and a count of
// the walks per (root, depth-1 child).  step1_cdf_kernel: for every pair picked by >= S1_MIN_WALKS walks, the
// child's candidate list / softmax / un-normalised CDF + total, built ONCE (walk_kernel then inverts it per walk;
// a pair picked once is cheaper inside its walk, which needs no CDF array).
__global__ void root_step_kernel(const __grid_constant__ gg_walk_desc d) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= d.n_walks) return;
    const int slot = __ldg(d.walk_slot + w);
    const int root = d.roots[slot];
    const uint32_t k = (uint32_t)(w - __ldg(d.walk_ptr + slot));
    const uint32_t k0 = (uint32_t)d.seed, k1 = (uint32_t)(d.seed >> 32);
    uint32_t a, b;
    if (d.update_ratio < 1.0) {
        philox4x32_10((uint32_t)root, 0xffffffffu, 0u, d.pass_tag, k0, k1, a, b);
        if (!(u53(a, b) < d.update_ratio)) { d.first_idx[w] = -2; return; }
    }
    const long long a0 = d.indptr[root];
    const int n = (int)(d.indptr[root + 1] - a0);
    if (n == 0) { d.first_idx[w] = -1; return; }
    philox4x32_10((uint32_t)root, k, 0u, d.pass_tag, k0, k1, a, b);
    const long long o = __ldg(d.rq_ptr + slot);
    const int idx = cdf_search(d.root_q + o, n, u53(a, b));
    d.first_idx[w] = idx;
    atomicAdd(d.s1_cnt + o + idx, 1);
}

constexpr int S1_SINGLES = 8192, S1_CHUNK = 16;

template <int CPL>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, WALK_MIN_CTAS) step1_cdf_kernel(const __grid_constant__ gg_walk_desc d) {
    extern __shared__ __align__(16) unsigned char walk_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *s_sc = reinterpret_cast<float *>(walk_smem + (size_t)wid * WALK_SMEM_PER_WARP);
    int *s_ids = reinterpret_cast<int *>(s_sc + SC_CAP);
    Stage stg;
    stg.buf = s_sc;
    stg.bar = reinterpret_cast<unsigned long long *>(s_ids + ID_CAP);
    stg.phase = 0u;
    stg.on = !d.no_tma && d.edge_score != nullptr;
    if (stg.on) {
        if (lane == 0) {
            mbar_init(stg.bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    const long long gw = (long long)blockIdx.x * WARPS_PER_CTA + wid;
    const long long nwarps = (long long)gridDim.x * WARPS_PER_CTA;
    int *g_ids = reinterpret_cast<int *>(d.scratch) + (size_t)gw * 2 * (size_t)d.max_cand;
    float *g_sc = reinterpret_cast<float *>(g_ids + d.max_cand);
    unsigned long long rows_gathered = 0;
    unsigned int cyc[7] = {0, 0, 0, 0, 0, 0, 0};
    // queue mode (s1_order): the first S1_SINGLES items are single pairs (the largest lists, one warp each, started
    // first); after them one item is a chunk of S1_CHUNK pairs (most pairs were never picked: one atomic per
    // chunk keeps the queue cheap).  Without s1_order: static striding.
    long long item = gw, sub = 0, sub_end = 0;
    const long long n_single = d.s1_nq < S1_SINGLES ? d.s1_nq : S1_SINGLES;
    const long long n_items = n_single + (d.s1_nq - n_single + S1_CHUNK - 1) / S1_CHUNK;
    for (;;) {
        long long pos;
        if (d.s1_order) {
            if (sub >= sub_end) {
                unsigned int it = 0;
                if (lane == 0) it = atomicAdd(d.work_counter, 1u);
                it = __shfl_sync(FULL, it, 0);
                if ((long long)it >= n_items) break;
                if ((long long)it < n_single) { sub = it; sub_end = sub + 1; }
                else { sub = n_single + ((long long)it - n_single) * S1_CHUNK; sub_end = sub + S1_CHUNK < d.s1_nq ? sub + S1_CHUNK : d.s1_nq; }
            }
            pos = __ldg(d.s1_order + sub);
            ++sub;
        } else {
            if (item >= d.s1_nq) break;
            pos = item;
            item += nwarps;
        }
        if (!s1_is_shared(d, pos)) continue;
        const int slot = __ldg(d.s1_slot + pos);
        const int root = d.roots[slot];
        const uint32_t *tb = d.tree_bits + (size_t)slot * (size_t)d.tree_words;
        const long long e = d.indptr[root] + (pos - __ldg(d.rq_ptr + slot));
        const int c = __ldg(d.adj + e);
        const bool inc_father = !d.for_d && !((d.d1_bits[e >> 5] >> (e & 31)) & 1u);   // graph_gan.py:258-259
        int n; float m; int *ids; float *sc;
        build_list<CPL, UNR_S1>(d, tb, c, root, inc_father, s_ids, s_sc, g_ids, g_sc, lane, n, m, ids, sc, rows_gathered, cyc, stg);
        if (lane == 0) d.s1_n[pos] = n;
        if (n == 0) continue;
        const long long o = __ldg(d.s1_ptr + pos);
        for (int i = lane; i < n; i += 32) d.s1_ids[o + i] = ids[i];
        if (n > 1) cdf_store_raw<UNR_S1>(sc, n, m, d.s1_q + o, lane);
        __syncwarp();
    }
    if (lane == 0 && rows_gathered) atomicAdd(d.counters + GG_CNT_ROWS_GATHERED, rows_gathered);
}

// ---------------------------------------------------------------- order-free (Philox) kernel
// FlatView: the buffers of the level-synchronous steps (see below), carved out of gg_walk_desc.flat_buf.
struct FlatView {
    int4 *list[2];         // [W] walks that execute step s next, as records (walk, node it stands on, node it came from,
                           //     root slot): list[s & 1] -- one 16-byte load gives a kernel everything about the walk
    int4 *tail;            // [W] walks the persistent kernel finishes (after the last level-synchronous step), same records
    int *hub;              // [W] per level: items (indices into the level's list) that stand on a score-cached node
    int *item_n;           // [W] per item: candidate-list length | father flag << 30 (0: nothing left to do for the item)
    int *pool_ids;         // [W * stride] per item: its candidate ids
    unsigned *ctr;         // counters: [0] tail length; level s: [1 + 4 s + {0: items, 1: hub items, 2: -, 3: work queue}]
    int stride;            // pool entries per item (>= hub_threshold: a node below the threshold has fewer neighbours)
    int steps;             // level-synchronous steps 1 .. steps
};
constexpr int FLAT_MAX_STEPS = 14;
constexpr int FLAT_CTR_WORDS = 1 + 4 * (FLAT_MAX_STEPS + 2);
#define GG_FCTR(fv, s, k) ((fv).ctr + 1 + 4 * (s) + (k))

template <int CPL>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, WALK_MIN_CTAS) walk_kernel(const __grid_constant__ gg_walk_desc d,
                                                                                 const FlatView fv, const int tail_mode) {
    extern __shared__ __align__(16) unsigned char walk_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *s_sc = reinterpret_cast<float *>(walk_smem + (size_t)wid * WALK_SMEM_PER_WARP);
    int *s_ids = reinterpret_cast<int *>(s_sc + SC_CAP);
    Stage stg;
    stg.buf = s_sc;
    stg.bar = reinterpret_cast<unsigned long long *>(s_ids + ID_CAP);
    stg.phase = 0u;
    stg.on = !d.no_tma && d.edge_score != nullptr;
    if (stg.on) {
        if (lane == 0) {
            mbar_init(stg.bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    const long long gw = (long long)blockIdx.x * WARPS_PER_CTA + wid;
    int *g_ids = reinterpret_cast<int *>(d.scratch) + (size_t)gw * 2 * (size_t)d.max_cand;
    float *g_sc = reinterpret_cast<float *>(g_ids + d.max_cand);
    Rng rng;
    rng.mode = GG_RNG_PHILOX; rng.k0 = (uint32_t)d.seed; rng.k1 = (uint32_t)(d.seed >> 32); rng.tag = d.pass_tag;
    rng.stream = nullptr; rng.n_stream = 0; rng.cursor = 0; rng.exhausted = 0;
    unsigned long long raw_steps = 0, raw_suml = 0, overflow = 0, rows_gathered = 0;
    unsigned int cyc[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool ratio_all = d.update_ratio >= 1.0;

    const long long n_items = tail_mode ? (long long)fv.ctr[0] : d.n_walks;
    for (;;) {
        unsigned int wi = 0;
        if (lane == 0) wi = atomicAdd(d.work_counter, 1u);
        wi = __shfl_sync(FULL, wi, 0);
        if ((long long)wi >= n_items) break;
        if (tail_mode) {
            // a walk handed over by the level-synchronous steps: continue where it stands
            const int4 rec = fv.tail[wi];
            const long long w = rec.x;
            const int slot = rec.w;
            WalkState from;
            from.cur = rec.y; from.prev = rec.z; from.step = d.wsteps[w]; from.fedge = d.first_edge[w]; from.suml = d.wsuml[w];
            walk_one<CPL>(d, rng, slot, (uint32_t)(w - __ldg(d.walk_ptr + slot)), w, s_ids, s_sc, g_ids, g_sc, lane, raw_steps,
                          raw_suml, overflow, rows_gathered, cyc, stg, &from);
            continue;
        }
        const long long w = d.walk_order ? (long long)__ldg(d.walk_order + wi) : (long long)wi;
        // walk -> root slot: last slot with walk_ptr[slot] <= w (table when the caller provides one)
        int slot;
        if (d.walk_slot) {
            slot = __ldg(d.walk_slot + w);
        } else {
            long long lo = 0, hi = d.n_roots;
            while (hi - lo > 1) {
                const long long mid = (lo + hi) >> 1;
                if (__ldg(d.walk_ptr + mid) <= w) lo = mid; else hi = mid;
            }
            slot = (int)lo;
        }
        const uint32_t k = (uint32_t)(w - __ldg(d.walk_ptr + slot));
        if (!ratio_all) {  // graph_gan.py:189/209: one draw per root
            uint32_t a, b;
            philox4x32_10((uint32_t)d.roots[slot], 0xffffffffu, 0u, rng.tag, rng.k0, rng.k1, a, b);
            if (!(u53(a, b) < d.update_ratio)) {
                if (lane == 0) {
                    d.samples[w] = -1; d.status[w] = GG_SKIPPED; d.first_edge[w] = -1; d.wsteps[w] = 0; d.wsuml[w] = 0;
                    if (d.path_len) d.path_len[w] = 0;
                }
                continue;
            }
        }
        walk_one<CPL>(d, rng, slot, k, w, s_ids, s_sc, g_ids, g_sc, lane, raw_steps, raw_suml, overflow, rows_gathered,
                      cyc, stg);
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 7; ++q) if (cyc[q]) atomicAdd(d.counters + GG_CNT_CYC_ENUM + q, (unsigned long long)cyc[q]);
        if (raw_steps) atomicAdd(d.counters + GG_CNT_RAW_STEPS, raw_steps);
        if (raw_suml) atomicAdd(d.counters + GG_CNT_RAW_SUML, raw_suml);
        if (overflow) atomicAdd(d.counters + GG_CNT_PATH_OVERFLOW, overflow);
        if (rows_gathered) atomicAdd(d.counters + GG_CNT_ROWS_GATHERED, rows_gathered);
    }
}

// ---------------------------------------------------------------- level-synchronous ("flat") steps
// The persistent kernel above gives every walk to one warp from root to leaf: per step the warp runs the dependent chain
// indptr -> tree bits / adjacency -> embedding rows -> softmax -> draw alone, so most of the time a resident warp has
// nothing in flight (ncu: 22 cycles per issued instruction, a third of them instruction-cache misses of 32 warps
// scattered over 130 KB of code).  Here all unfinished walks take step s TOGETHER, one phase per kernel:
//   flat_start_kernel   thread per walk: the root step (inverts the root's CDF) and, for (root, child) pairs picked by
//                       several walks, step 1 from the shared CDF (step1_cdf_kernel); survivors enter level 1 or 2
//   flat_enum_kernel    warp per unfinished walk: candidate list [father] + children(cur) from the tree bits into the
//                       item's slab of the id pool (short chain: indptr -> bits + adjacency); empty and single-candidate
//                       lists are finished on the spot; walks standing on a score-cached (hub) node go to the hub list
//   flat_choose_kernel  warp per item: on-demand scores (rows of the candidates: the only phase with row gathers, so
//                       every resident warp has 8 rows in flight almost all the time), softmax + CDF + draw, next
//                       node; hub items run the cached-list step of the persistent kernel (TMA-staged enumeration)
// and after `steps` levels the few walks still alive are finished by walk_kernel in tail mode.  Same arithmetic, same
// Philox counters (root, walk, step): bit-identical to the persistent kernel (tests: test_flat_steps_*).
__device__ __forceinline__ bool step_includes_father(const gg_walk_desc &d, int s, int fedge) {
    if (s == 0) return false;                              // graph_gan.py:250: the root has no father
    if (s == 1) {
        if (d.for_d) return false;                         // graph_gan.py:255-257
        return !((d.d1_bits[fedge >> 5] >> (fedge & 31)) & 1u);   // graph_gan.py:258-259: father entry removed by a D pass
    }
    return true;
}

// lane 0: the walk made its choice at step s over n candidates
__device__ __forceinline__ void flat_advance(const gg_walk_desc &d, const FlatView &fv, int s, long long w, int slot, int cur,
                                             int n, int idx, int nxt, bool inc_father, unsigned long long &overflow) {
    if (d.max_path > 0 && d.paths && s + 1 < d.max_path) d.paths[(size_t)w * (size_t)d.max_path + s + 1] = nxt;
    d.wsteps[w] = s + 1;
    d.wsuml[w] = d.wsuml[w] + n;
    if (inc_father && idx == 0) {                          // graph_gan.py:264-266: back to the father: cur is the sample
        d.samples[w] = cur; d.status[w] = GG_DONE;
        if (d.path_len) d.path_len[w] = s + 2;
        if (d.max_path > 0 && s + 2 > d.max_path) overflow += 1;
    } else {
        const int4 rec = make_int4((int)w, nxt, cur, slot);
        if (s < fv.steps) fv.list[(s + 1) & 1][atomicAdd(GG_FCTR(fv, s + 1, 0), 1u)] = rec;
        else fv.tail[atomicAdd(fv.ctr, 1u)] = rec;
    }
}
__device__ __forceinline__ void flat_void(const gg_walk_desc &d, int s, long long w) {   // lane 0; graph_gan.py:252-257
    d.samples[w] = -1; d.status[w] = GG_VOID; d.wsteps[w] = s;
    if (d.path_len) d.path_len[w] = 0;
}

__global__ void __launch_bounds__(256) flat_start_kernel(const __grid_constant__ gg_walk_desc d, const FlatView fv) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int dest = 0;                                          // 1 / 2: enters level 1 / 2 (or the tail when there is no such level)
    int4 rec = make_int4(0, 0, 0, 0);
    unsigned steps = 0, suml = 0, over = 0;
    if (w < d.n_walks) {
        const int slot = __ldg(d.walk_slot + w);
        const int root = d.roots[slot];
        const uint32_t k = (uint32_t)(w - __ldg(d.walk_ptr + slot));
        const int fi = d.first_idx[w];                      // root_step_kernel: -2 skipped (update_ratio), -1 isolated root
        int32_t *prow = (d.max_path > 0 && d.paths) ? d.paths + (size_t)w * (size_t)d.max_path : nullptr;
        int sample = -1, status = GG_NOTRUN, fedge = -1, ws = 0, wl = 0, plen = 0;
        if (fi == -2) {
            status = GG_SKIPPED;
        } else {
            if (prow) prow[0] = root;
            if (fi < 0) {
                status = GG_VOID;                           // graph_gan.py:252-253
            } else {
                const long long a0 = d.indptr[root];
                const int n0 = (int)(d.indptr[root + 1] - a0);
                fedge = (int)(a0 + fi);
                const int c = __ldg(d.adj + fedge);
                if (prow && 1 < d.max_path) prow[1] = c;
                ws = 1; wl = n0; steps = 1; suml = (unsigned)n0;
                const long long s1pos = __ldg(d.rq_ptr + slot) + fi;
                if (s1_is_shared(d, s1pos)) {
                    const int n = __ldg(d.s1_n + s1pos);
                    if (n == 0) {
                        status = GG_VOID;                   // graph_gan.py:255-257
                    } else {
                        const bool inc_father = step_includes_father(d, 1, fedge);
                        const long long o = __ldg(d.s1_ptr + s1pos);
                        uint32_t a, b;
                        philox4x32_10((uint32_t)root, k, 1u, d.pass_tag, (uint32_t)d.seed, (uint32_t)(d.seed >> 32), a, b);
                        const int idx = (n == 1) ? 0 : cdf_search_raw(d.s1_q + o, n, u53(a, b));
                        const int nxt = __ldg(d.s1_ids + o + idx);
                        if (prow && 2 < d.max_path) prow[2] = nxt;
                        ws = 2; wl = n0 + n; steps = 2; suml = (unsigned)(n0 + n);
                        if (inc_father && idx == 0) {
                            sample = c; status = GG_DONE; plen = 3;
                            if (d.max_path > 0 && 3 > d.max_path) over = 1;
                        } else {
                            rec = make_int4((int)w, nxt, c, slot); dest = 2;
                        }
                    }
                } else {
                    rec = make_int4((int)w, c, root, slot); dest = 1;
                }
            }
        }
        d.samples[w] = sample; d.status[w] = status; d.first_edge[w] = fedge; d.wsteps[w] = ws; d.wsuml[w] = wl;
        if (d.path_len) d.path_len[w] = plen;
    }
    // warp-aggregated appends (one atomic per warp and destination)
#pragma unroll
    for (int lv = 1; lv <= 2; ++lv) {
        const unsigned mk = __ballot_sync(FULL, dest == lv);
        if (!mk) continue;
        const int leader = __ffs(mk) - 1;
        int4 *list = (lv <= fv.steps) ? fv.list[lv & 1] : fv.tail;
        unsigned *cnt = (lv <= fv.steps) ? GG_FCTR(fv, lv, 0) : fv.ctr;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(cnt, (unsigned)__popc(mk));
        base = __shfl_sync(FULL, base, leader);
        if (dest == lv) list[base + __popc(mk & ((1u << lane) - 1u))] = rec;
    }
    steps = __reduce_add_sync(FULL, steps); suml = __reduce_add_sync(FULL, suml); over = __reduce_add_sync(FULL, over);
    if (lane == 0) {
        if (steps) atomicAdd(d.counters + GG_CNT_RAW_STEPS, (unsigned long long)steps);
        if (suml) atomicAdd(d.counters + GG_CNT_RAW_SUML, (unsigned long long)suml);
        if (over) atomicAdd(d.counters + GG_CNT_PATH_OVERFLOW, (unsigned long long)over);
    }
}

constexpr int FLAT_ENUM_WARPS = 8;
__global__ void __launch_bounds__(FLAT_ENUM_WARPS * 32, 6) flat_enum_kernel(const __grid_constant__ gg_walk_desc d,
                                                                            const FlatView fv, const int s) {
    const int lane = threadIdx.x & 31;
    const unsigned gw = blockIdx.x * FLAT_ENUM_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * FLAT_ENUM_WARPS;
    const int4 *A = fv.list[s & 1];
    const unsigned nA = *GG_FCTR(fv, s, 0);
    Stage stg;
    stg.buf = nullptr; stg.bar = nullptr; stg.phase = 0u; stg.on = false;
    unsigned long long raw_steps = 0, raw_suml = 0, overflow = 0;
    int4 rec_next = (gw < nA) ? A[gw] : make_int4(0, 0, 0, 0);
    for (unsigned i = gw; i < nA; i += nwarps) {
        const int4 rec = rec_next;
        if (i + nwarps < nA) rec_next = A[i + nwarps];     // the next item's record is in flight while this one is enumerated
        const long long w = rec.x;
        const int cur = rec.y, prev = rec.z, slot = rec.w;
        const long long a0 = d.indptr[cur], a1 = d.indptr[cur + 1];
        if (d.edge_score && (a1 - a0) >= d.hub_threshold) {          // score-cached node: the whole step runs in flat_choose_kernel
            if (lane == 0) { fv.hub[atomicAdd(GG_FCTR(fv, s, 1), 1u)] = (int)i; fv.item_n[i] = 0; }
            continue;
        }
        const bool inc_father = step_includes_father(d, s, s == 1 ? d.first_edge[w] : 0);
        const uint32_t *tb = d.tree_bits + (size_t)slot * (size_t)d.tree_words;
        int *ids = fv.pool_ids + (size_t)i * (size_t)fv.stride;
        int n = 0;
        if (inc_father) { if (lane == 0) ids[0] = prev; n = 1; }
        float m = 0.0f;
        enumerate_children<UNR>(d, tb, a0, a1, false, ids, nullptr, lane, n, m, stg);
        __syncwarp();
        if (n == 0) {
            if (lane == 0) { flat_void(d, s, w); fv.item_n[i] = 0; }
        } else if (n == 1) {
            // softmax = [1.0], cdf = [1.0] and 1.0 > u for every u in [0, 1): index 0, no score, no draw needed
            const int nxt = inc_father ? prev : ids[0];
            if (lane == 0) { flat_advance(d, fv, s, w, slot, cur, 1, 0, nxt, inc_father, overflow); fv.item_n[i] = 0; }
            raw_steps += 1; raw_suml += 1;
        } else if (lane == 0) {
            fv.item_n[i] = n | (inc_father ? (1 << 30) : 0);
        }
    }
    if (lane == 0) {
        if (raw_steps) atomicAdd(d.counters + GG_CNT_RAW_STEPS, raw_steps);
        if (raw_suml) atomicAdd(d.counters + GG_CNT_RAW_SUML, raw_suml);
        if (overflow) atomicAdd(d.counters + GG_CNT_PATH_OVERFLOW, overflow);
    }
}

template <int CPL>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, WALK_MIN_CTAS) flat_choose_kernel(const __grid_constant__ gg_walk_desc d,
                                                                                        const FlatView fv, const int s) {
    extern __shared__ __align__(16) unsigned char walk_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *s_sc = reinterpret_cast<float *>(walk_smem + (size_t)wid * WALK_SMEM_PER_WARP);
    int *s_ids = reinterpret_cast<int *>(s_sc + SC_CAP);
    Stage stg;
    stg.buf = s_sc;
    stg.bar = reinterpret_cast<unsigned long long *>(s_ids + ID_CAP);
    stg.phase = 0u;
    stg.on = !d.no_tma && d.edge_score != nullptr;
    if (stg.on) {
        if (lane == 0) {
            mbar_init(stg.bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    const long long gw = (long long)blockIdx.x * WARPS_PER_CTA + wid;
    int *g_ids = reinterpret_cast<int *>(d.scratch) + (size_t)gw * 2 * (size_t)d.max_cand;
    float *g_sc = reinterpret_cast<float *>(g_ids + d.max_cand);
    const int4 *A = fv.list[s & 1];
    const unsigned nA = *GG_FCTR(fv, s, 0), nH = *GG_FCTR(fv, s, 1);
    const uint32_t k0 = (uint32_t)d.seed, k1 = (uint32_t)(d.seed >> 32);
    unsigned long long raw_steps = 0, raw_suml = 0, overflow = 0, rows_gathered = 0;
    unsigned int cyc[7] = {0, 0, 0, 0, 0, 0, 0};
    // work queue: the hub items first, one per pull (they are the long ones); then the other items in chunks of
    // FLAT_CHUNK consecutive list positions per pull, the next item's record and list length in flight while the
    // current one is scored
    constexpr unsigned FLAT_CHUNK = 8;
    const unsigned n_pulls = nH + (nA + FLAT_CHUNK - 1) / FLAT_CHUNK;
    for (;;) {
        unsigned j = 0;
        if (lane == 0) j = atomicAdd(GG_FCTR(fv, s, 3), 1u);
        j = __shfl_sync(FULL, j, 0);
        if (j >= n_pulls) break;
        const bool hub_item = j < nH;
        unsigned i = hub_item ? (unsigned)fv.hub[j] : (j - nH) * FLAT_CHUNK;
        const unsigned i_end = hub_item ? i + 1 : ((i + FLAT_CHUNK < nA) ? i + FLAT_CHUNK : nA);
        int4 rec_next = A[i];
        int n_next = hub_item ? 0 : fv.item_n[i];
        for (; i < i_end; ++i) {
            const int4 rec = rec_next;
            const int nrec = n_next;
            if (i + 1 < i_end) { rec_next = A[i + 1]; n_next = fv.item_n[i + 1]; }
            if (!hub_item && (nrec & 0x3fffffff) < 2) continue;      // finished by flat_enum_kernel, or a hub item
            const long long w = rec.x;
            const int cur = rec.y, prev = rec.z, slot = rec.w;
            int n, idx, nxt;
            bool inc_father;
            uint32_t a, b;
            if (hub_item) {
                inc_father = step_includes_father(d, s, s == 1 ? d.first_edge[w] : 0);
                const uint32_t *tb = d.tree_bits + (size_t)slot * (size_t)d.tree_words;
                int *ids; float *sc; float m;
                build_list<CPL, UNR>(d, tb, cur, prev, inc_father, s_ids, s_sc, g_ids, g_sc, lane, n, m, ids, sc, rows_gathered, cyc, stg);
                if (n == 0) {
                    if (lane == 0) flat_void(d, s, w);
                    continue;
                }
                philox4x32_10((uint32_t)d.roots[slot], (uint32_t)(w - __ldg(d.walk_ptr + slot)), (uint32_t)s, d.pass_tag, k0, k1, a, b);
                idx = (n == 1) ? 0 : choose_index(sc, n, m, u53(a, b), lane, sc != s_sc ? reinterpret_cast<double *>(s_sc) : nullptr);
                nxt = ids[idx];
                __syncwarp();
            } else {
                n = nrec & 0x3fffffff;
                inc_father = (nrec >> 30) & 1;
                const int *ids = fv.pool_ids + (size_t)i * (size_t)fv.stride;
                float4 c4[CPL];
                load_row<CPL>(d.emb, d.ld, cur, lane & 7, c4);
                const int root = d.roots[slot];
                const uint32_t k = (uint32_t)(w - __ldg(d.walk_ptr + slot));
                score_list<CPL>(d.emb, d.bias, d.ld, c4, ids, s_sc, n, cur, lane);
                rows_gathered += 1u + (unsigned)n;
                const float m = list_max(s_sc, n, lane);
                philox4x32_10((uint32_t)root, k, (uint32_t)s, d.pass_tag, k0, k1, a, b);
                idx = choose_index(s_sc, n, m, u53(a, b), lane);
                nxt = ids[idx];
                __syncwarp();
            }
            if (lane == 0) {
                flat_advance(d, fv, s, w, slot, cur, n, idx, nxt, inc_father, overflow);
                raw_steps += 1; raw_suml += (unsigned)n;
            }
        }
    }
    if (lane == 0) {
        if (raw_steps) atomicAdd(d.counters + GG_CNT_RAW_STEPS, raw_steps);
        if (raw_suml) atomicAdd(d.counters + GG_CNT_RAW_SUML, raw_suml);
        if (overflow) atomicAdd(d.counters + GG_CNT_PATH_OVERFLOW, overflow);
        if (rows_gathered) atomicAdd(d.counters + GG_CNT_ROWS_GATHERED, rows_gathered);
    }
}

// the layout of desc.flat_buf (host): returns the bytes needed; fills `fv` when `buf` is given
size_t flat_layout(void *buf, long long n_walks, int hub_threshold, int steps, FlatView *fv) {
    const size_t W = (size_t)(n_walks > 0 ? n_walks : 1);
    const int stride = ((hub_threshold > 0 ? hub_threshold : 1) + 31) / 32 * 32;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *p = buf ? (void *)((unsigned char *)buf + off) : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    unsigned *ctr = (unsigned *)take(sizeof(unsigned) * FLAT_CTR_WORDS);
    int4 *l0 = (int4 *)take(16 * W), *l1 = (int4 *)take(16 * W), *tail = (int4 *)take(16 * W);
    int *hub = (int *)take(4 * W), *item_n = (int *)take(4 * W);
    int *pool = (int *)take(4 * W * (size_t)stride);
    if (fv) {
        fv->ctr = ctr; fv->list[0] = l0; fv->list[1] = l1; fv->tail = tail; fv->hub = hub;
        fv->item_n = item_n; fv->pool_ids = pool; fv->stride = stride; fv->steps = steps;
    }
    return off;
}

// ---------------------------------------------------------------- reference-order (stream) kernel
// One warp replays the reference's sequential consumption of a uniform stream: a draw per
// root (graph_gan.py:189/209), a draw per choice (:262), stop at a root's first void.
template <int CPL>
__global__ void __launch_bounds__(32) walk_stream_kernel(const __grid_constant__ gg_walk_desc d) {
    extern __shared__ __align__(16) unsigned char walk_smem[];
    float *s_sc = reinterpret_cast<float *>(walk_smem);
    int *s_ids = reinterpret_cast<int *>(s_sc + SC_CAP);
    const int lane = threadIdx.x;
    int *g_ids = reinterpret_cast<int *>(d.scratch);
    float *g_sc = reinterpret_cast<float *>(g_ids + d.max_cand);
    Rng rng;
    rng.mode = GG_RNG_STREAM; rng.k0 = rng.k1 = rng.tag = 0;
    rng.stream = d.stream; rng.n_stream = d.n_stream; rng.cursor = 0; rng.exhausted = 0;
    unsigned long long raw_steps = 0, raw_suml = 0, overflow = 0, rows_gathered = 0;
    unsigned int cyc[7] = {0, 0, 0, 0, 0, 0, 0};
    Stage stg;                                              // the replay kernel uses plain loads
    stg.buf = s_sc; stg.bar = nullptr; stg.phase = 0u; stg.on = false;
    for (long long slot = 0; slot < d.n_roots && !rng.exhausted; ++slot) {
        const long long w0 = d.walk_ptr[slot], w1 = d.walk_ptr[slot + 1];
        const double ur = rng.draw(0, 0, 0);
        const bool skip = !(ur < d.update_ratio);
        bool dead = skip || rng.exhausted;
        for (long long w = w0; w < w1; ++w) {
            if (dead) {
                if (lane == 0) {
                    d.samples[w] = -1; d.status[w] = skip ? GG_SKIPPED : GG_NOTRUN; d.first_edge[w] = -1;
                    d.wsteps[w] = 0; d.wsuml[w] = 0;
                    if (d.path_len) d.path_len[w] = 0;
                }
                continue;
            }
            const int st = walk_one<CPL>(d, rng, (int)slot, (uint32_t)(w - w0), w, s_ids, s_sc, g_ids, g_sc, lane,
                                         raw_steps, raw_suml, overflow, rows_gathered, cyc, stg);
            if (st != GG_DONE) dead = true;
        }
    }
    if (lane == 0) {
        atomicAdd(d.counters + GG_CNT_RAW_STEPS, raw_steps);
        atomicAdd(d.counters + GG_CNT_RAW_SUML, raw_suml);
        atomicAdd(d.counters + GG_CNT_PATH_OVERFLOW, overflow);
        atomicAdd(d.counters + GG_CNT_ROWS_GATHERED, rows_gathered);
        d.counters[GG_CNT_STREAM_USED] = (unsigned long long)rng.cursor + (rng.exhausted ? (1ull << 62) : 0ull);
    }
}

// ---------------------------------------------------------------- finalize (thread per walk)
// Three flat passes instead of one warp per root (a hub root has > 10 k walks: its warp was the launch's tail):
//   1. every walk that is not DONE lowers its root's "first bad walk" (atomicMin into root_ok, used as scratch)
//   2. every walk up to and including the first bad one adds its counters and (D mode) its father-removal bit;
//      the walks after it are blanked -- the reference never ran them (graph_gan.py:252-257 returned early)
//   3. per root: root_ok = no bad walk and at least one walk ("neg is not None and len(pos) != 0", graph_gan.py:192)
constexpr int FIN_NONE = 0x7f7f7f7f;                      // (the memset pattern root_ok is initialised with)

__device__ __forceinline__ long long walk_root_slot(const long long *__restrict__ walk_ptr, long long n_roots, long long w) {
    long long lo = 0, hi = n_roots;                         // last slot with walk_ptr[slot] <= w
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(walk_ptr + mid) <= w) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void finalize_mark_kernel(long long n_roots, const long long *__restrict__ walk_ptr, const int *__restrict__ status,
                                     int *root_ok) {
    const long long W = walk_ptr[n_roots];
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += (long long)gridDim.x * blockDim.x) {
        if (status[w] == GG_DONE) continue;
        const long long slot = walk_root_slot(walk_ptr, n_roots, w);
        atomicMin(root_ok + slot, (int)(w - walk_ptr[slot]));
    }
}

__global__ void finalize_apply_kernel(long long n_roots, const long long *__restrict__ walk_ptr, int for_d, int *samples,
                                      int *status, const int *__restrict__ first_edge, int *wsteps, int *wsuml, int *path_len,
                                      uint32_t *d1_bits, const int *__restrict__ root_ok, unsigned long long *counters) {
    const long long W = walk_ptr[n_roots];
    unsigned long long steps = 0, suml = 0;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += (long long)gridDim.x * blockDim.x) {
        const long long slot = walk_root_slot(walk_ptr, n_roots, w);
        const long long off = w - walk_ptr[slot];
        if (off <= (long long)root_ok[slot]) {
            steps += (unsigned)wsteps[w]; suml += (unsigned)wsuml[w];
            if (for_d && status[w] == GG_DONE) {
                const int fe = first_edge[w];
                if (fe >= 0) atomicOr(d1_bits + (fe >> 5), 1u << (fe & 31));
            }
        } else {
            if (status[w] != GG_SKIPPED) status[w] = GG_NOTRUN;
            samples[w] = -1; wsteps[w] = 0; wsuml[w] = 0;
            if (path_len) path_len[w] = 0;
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        steps += __shfl_xor_sync(FULL, steps, off);
        suml += __shfl_xor_sync(FULL, suml, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (steps) atomicAdd(counters + GG_CNT_STEPS, steps);
        if (suml) atomicAdd(counters + GG_CNT_SUML, suml);
    }
}

__global__ void finalize_roots_kernel(long long n_roots, const long long *__restrict__ walk_ptr, int *root_ok,
                                      unsigned long long *counters) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0, okr = 0;
    if (slot < n_roots) {
        const long long k = walk_ptr[slot + 1] - walk_ptr[slot];
        const bool ok = root_ok[slot] == FIN_NONE && k > 0;
        root_ok[slot] = ok ? 1 : 0;
        if (ok) { acc = (unsigned long long)k; okr = 1; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        acc += __shfl_xor_sync(FULL, acc, off);
        okr += __shfl_xor_sync(FULL, okr, off);
    }
    if ((threadIdx.x & 31) == 0 && okr) {
        atomicAdd(counters + GG_CNT_ACCEPTED, acc);
        atomicAdd(counters + GG_CNT_OK_ROOTS, okr);
    }
}

// ---------------------------------------------------------------- D rows
__global__ void row_count_kernel(long long n_roots, const long long *walk_ptr, const int *root_ok, long long *row_ptr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_roots) row_ptr[i] = root_ok[i] ? 2 * (walk_ptr[i + 1] - walk_ptr[i]) : 0;
}

// thread per walk: its positive row and its negative row (graph_gan.py:194-201)
__global__ void emit_rows_kernel(long long n_roots, const int *__restrict__ roots, const long long *__restrict__ walk_ptr,
                                 const long long *__restrict__ pos_indptr, const int *__restrict__ pos_flat,
                                 const int *__restrict__ root_ok, const int *__restrict__ samples,
                                 const long long *__restrict__ row_ptr, int *center, int *neighbor, int *label) {
    const long long W = walk_ptr[n_roots];
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < W; w += (long long)gridDim.x * blockDim.x) {
        const long long slot = walk_root_slot(walk_ptr, n_roots, w);
        if (!root_ok[slot]) continue;
        const int r = roots[slot];
        const long long w0 = walk_ptr[slot], k = walk_ptr[slot + 1] - w0, t = w - w0, o = row_ptr[slot];
        center[o + t] = r; neighbor[o + t] = pos_flat[pos_indptr[r] + t]; label[o + t] = 1;
        center[o + k + t] = r; neighbor[o + k + t] = samples[w]; label[o + k + t] = 0;
    }
}

int grid_ctas() { return sm_count() * WALK_MIN_CTAS; }

}  // namespace
}  // namespace gg

extern "C" int gg_walk_scratch_bytes(int32_t max_cand, int64_t *bytes) {
    GG_REQUIRE(bytes && max_cand > 0, "bad arguments");
    const int64_t warps = (int64_t)gg::grid_ctas() * gg::WARPS_PER_CTA;
    *bytes = warps * 2 * (int64_t)max_cand * 4;
    return 0;
}

extern "C" int gg_walk_flat_bytes(int64_t n_walks, int32_t hub_threshold, int32_t flat_steps, int64_t *bytes) {
    GG_REQUIRE(bytes && n_walks >= 0 && hub_threshold > 0 && flat_steps >= 0 && flat_steps <= gg::FLAT_MAX_STEPS, "bad arguments");
    *bytes = (int64_t)gg::flat_layout(nullptr, n_walks, hub_threshold, flat_steps, nullptr);
    return 0;
}

extern "C" int gg_walk_sample(const gg_walk_desc *dp, void *stream) {
    GG_REQUIRE(dp, "null descriptor");
    const gg_walk_desc &d = *dp;
    GG_REQUIRE(d.ld > 0 && d.ld % 32 == 0, "ld must be a positive multiple of 32");
    if (d.n_walks == 0 || d.n_roots == 0) return 0;   // nothing to do (empty batches carry null pointers)
    GG_REQUIRE(d.emb && d.bias && d.indptr && d.adj && d.roots && d.tree_bits && d.walk_ptr, "null graph/embedding pointer");
    GG_REQUIRE(d.tree_words > 0, "tree_words missing (gg_tree_words)");
    GG_REQUIRE(d.samples && d.status && d.first_edge && d.wsteps && d.wsuml && d.counters && d.work_counter,
               "null output pointer");
    GG_REQUIRE(d.for_d || d.d1_bits, "G mode needs d1_bits");
    GG_REQUIRE(d.n_walks < (1ll << 32), "too many walks in one call");
    GG_REQUIRE(d.max_cand > 0 && d.scratch, "scratch missing");
    GG_REQUIRE(!d.root_q || d.rq_ptr, "root_q needs rq_ptr");
    // optional groups: all of a group or none of it (a half-filled descriptor is an argument error, not a fault)
    GG_REQUIRE(d.max_path >= 0 && (d.max_path == 0 || (d.paths && d.path_len)), "max_path > 0 needs paths and path_len");
    GG_REQUIRE(d.phase_mask >= 0 && d.phase_mask <= 3, "phase_mask must be 0..3");
    GG_REQUIRE(d.update_ratio >= 0.0, "update_ratio must be >= 0");
    GG_REQUIRE(d.rng_mode == GG_RNG_PHILOX || d.rng_mode == GG_RNG_STREAM, "unknown rng_mode");
    {
        const bool any_s1 = d.s1_q || d.s1_ids || d.s1_cnt || d.s1_n || d.s1_ptr || d.s1_slot || d.first_idx || d.s1_order;
        GG_REQUIRE(!any_s1 || (d.s1_q && d.s1_ids && d.s1_cnt && d.s1_n && d.s1_ptr && d.s1_slot && d.first_idx && d.s1_nq > 0),
                   "depth-1 reuse: s1_q, s1_ids, s1_cnt, s1_n, s1_ptr, s1_slot, first_idx and s1_nq go together");
        GG_REQUIRE(!any_s1 || d.rng_mode == GG_RNG_PHILOX, "depth-1 reuse needs GG_RNG_PHILOX");
        GG_REQUIRE(!d.walk_order || d.rng_mode == GG_RNG_PHILOX, "walk_order needs GG_RNG_PHILOX");
    }
    GG_REQUIRE(!d.edge_score || (d.hub_threshold > 0 && d.hub_threshold < gg::SMEM_CAP), "hub_threshold out of range");
    cudaStream_t st = (cudaStream_t)stream;
    GG_CHECK(cudaMemsetAsync(d.work_counter, 0, sizeof(unsigned int), st));
    if (d.n_walks == 0 || d.n_roots == 0) return 0;
    const int cpl = d.ld / 32;
    if (d.rng_mode == GG_RNG_STREAM) {
        GG_REQUIRE(d.stream, "stream mode needs the uniform stream");
        GG_REQUIRE(d.scratch_bytes >= 2ll * d.max_cand * 4, "scratch too small");
        switch (cpl) {
#define GG_STREAM(C)                                                                                                  \
    GG_CHECK(cudaFuncSetAttribute(gg::walk_stream_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                  gg::WALK_SMEM_PER_WARP));                                                          \
    gg::walk_stream_kernel<C><<<1, 32, gg::WALK_SMEM_PER_WARP, st>>>(d)
            case 1: GG_STREAM(1); break;
            case 2: GG_STREAM(2); break;
            case 4: GG_STREAM(4); break;
            case 8: GG_STREAM(8); break;
#undef GG_STREAM
            default: gg::set_error("gg_walk_sample: unsupported ld %d (supported: 32, 64, 128, 256)", d.ld); return 2;
        }
    } else {
        const int ctas = gg::grid_ctas();
        GG_REQUIRE(d.scratch_bytes >= (int64_t)ctas * gg::WARPS_PER_CTA * 2 * d.max_cand * 4, "scratch too small");
        const int pm = d.phase_mask ? d.phase_mask : 3;   // 1 = depth-1 precompute, 2 = walk kernel (default both)
        if (d.s1_q && (pm & 1)) {   // depth-1 reuse: root steps + one CDF per (root, child) pair that occurs
            GG_REQUIRE(d.root_q && d.walk_slot && d.first_idx && d.s1_cnt && d.s1_n && d.s1_ptr && d.s1_ids && d.s1_slot,
                       "depth-1 reuse needs root_q, walk_slot and the s1_* buffers");
            GG_CHECK(cudaMemsetAsync(d.s1_cnt, 0, sizeof(int32_t) * (size_t)d.s1_nq, st));
            gg::root_step_kernel<<<(unsigned)((d.n_walks + 255) / 256), 256, 0, st>>>(d);
            GG_CHECK(cudaGetLastError());
#define GG_S1(C)                                                                                                      \
    GG_CHECK(cudaFuncSetAttribute(gg::step1_cdf_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                  gg::WARPS_PER_CTA * gg::WALK_SMEM_PER_WARP));                                      \
    gg::step1_cdf_kernel<C><<<ctas, gg::WARPS_PER_CTA * 32, gg::WARPS_PER_CTA * gg::WALK_SMEM_PER_WARP, st>>>(d)
            switch (cpl) {
                case 1: GG_S1(1); break;
                case 2: GG_S1(2); break;
                case 4: GG_S1(4); break;
                case 8: GG_S1(8); break;
                default: gg::set_error("gg_walk_sample: unsupported ld %d (supported: 32, 64, 128, 256)", d.ld); return 2;
            }
#undef GG_S1
            GG_CHECK(cudaGetLastError());
            GG_CHECK(cudaMemsetAsync(d.work_counter, 0, sizeof(unsigned int), st));   // the walk kernel's queue starts at 0
        }
        if (!(pm & 2)) return 0;
        gg::FlatView fv;
        memset(&fv, 0, sizeof(fv));
        int tail_mode = 0;
        if (d.flat_steps > 0) {
            // level-synchronous steps (flat_*_kernel), then the persistent kernel finishes what is left
            GG_REQUIRE(d.s1_q && d.edge_score && d.root_q && d.walk_slot && d.first_idx, "flat_steps needs the depth-1 reuse (s1_*, edge_score, root_q, walk_slot)");
            GG_REQUIRE(d.flat_steps <= gg::FLAT_MAX_STEPS, "flat_steps too large");
            GG_REQUIRE(d.flat_buf && d.flat_bytes >= (int64_t)gg::flat_layout(nullptr, d.n_walks, d.hub_threshold, d.flat_steps, nullptr),
                       "flat_buf too small (gg_walk_flat_bytes)");
            gg::flat_layout(d.flat_buf, d.n_walks, d.hub_threshold, d.flat_steps, &fv);
            GG_CHECK(cudaMemsetAsync(fv.ctr, 0, sizeof(unsigned) * gg::FLAT_CTR_WORDS, st));
            gg::flat_start_kernel<<<(unsigned)((d.n_walks + 255) / 256), 256, 0, st>>>(d, fv);
            GG_CHECK(cudaGetLastError());
            const int enum_ctas = gg::sm_count() * 8;
            for (int s = 1; s <= d.flat_steps; ++s) {
                gg::flat_enum_kernel<<<enum_ctas, gg::FLAT_ENUM_WARPS * 32, 0, st>>>(d, fv, s);
                GG_CHECK(cudaGetLastError());
                switch (cpl) {
#define GG_FLAT(C)                                                                                                    \
    GG_CHECK(cudaFuncSetAttribute(gg::flat_choose_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                  gg::WARPS_PER_CTA * gg::WALK_SMEM_PER_WARP));                                      \
    gg::flat_choose_kernel<C><<<ctas, gg::WARPS_PER_CTA * 32, gg::WARPS_PER_CTA * gg::WALK_SMEM_PER_WARP, st>>>(d, fv, s)
                    case 1: GG_FLAT(1); break;
                    case 2: GG_FLAT(2); break;
                    case 4: GG_FLAT(4); break;
                    case 8: GG_FLAT(8); break;
#undef GG_FLAT
                    default: gg::set_error("gg_walk_sample: unsupported ld %d (supported: 32, 64, 128, 256)", d.ld); return 2;
                }
                GG_CHECK(cudaGetLastError());
            }
            tail_mode = 1;
        }
        switch (cpl) {
#define GG_WALK(C)                                                                                                    \
    GG_CHECK(cudaFuncSetAttribute(gg::walk_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,                   \
                                  gg::WARPS_PER_CTA * gg::WALK_SMEM_PER_WARP));                                      \
    gg::walk_kernel<C><<<ctas, gg::WARPS_PER_CTA * 32, gg::WARPS_PER_CTA * gg::WALK_SMEM_PER_WARP, st>>>(d, fv, tail_mode)
            case 1: GG_WALK(1); break;
            case 2: GG_WALK(2); break;
            case 4: GG_WALK(4); break;
            case 8: GG_WALK(8); break;
#undef GG_WALK
            default: gg::set_error("gg_walk_sample: unsupported ld %d (supported: 32, 64, 128, 256)", d.ld); return 2;
        }
    }
    return gg::check_cuda(cudaGetLastError(), "walk kernel launch");
}

extern "C" int gg_walk_finalize(int64_t n_roots, const int64_t *walk_ptr, int32_t for_d, int32_t *samples,
                                int32_t *status, const int32_t *first_edge, int32_t *wsteps, int32_t *wsuml,
                                int32_t *path_len, uint32_t *d1_bits, int32_t *root_ok,
                                unsigned long long *counters, void *stream) {
    GG_REQUIRE(walk_ptr && samples && status && first_edge && wsteps && wsuml && root_ok && counters, "null pointer");
    GG_REQUIRE(!for_d || d1_bits, "D mode needs d1_bits");
    if (n_roots == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    const unsigned wgrid = (unsigned)(gg::sm_count() * 8);           // grid-stride over the walks (their number is on the device)
    GG_CHECK(cudaMemsetAsync(root_ok, 0x7f, sizeof(int32_t) * (size_t)n_roots, st));
    gg::finalize_mark_kernel<<<wgrid, threads, 0, st>>>(n_roots, (const long long *)walk_ptr, status, root_ok);
    GG_CHECK(cudaGetLastError());
    gg::finalize_apply_kernel<<<wgrid, threads, 0, st>>>(n_roots, (const long long *)walk_ptr, for_d, samples, status, first_edge,
                                                        wsteps, wsuml, path_len, d1_bits, root_ok, counters);
    GG_CHECK(cudaGetLastError());
    gg::finalize_roots_kernel<<<(unsigned)((n_roots + threads - 1) / threads), threads, 0, st>>>(
        n_roots, (const long long *)walk_ptr, root_ok, counters);
    return gg::check_cuda(cudaGetLastError(), "finalize kernel launch");
}

extern "C" int gg_emit_d_rows(int64_t n_roots, const int32_t *roots, const int64_t *walk_ptr,
                              const int64_t *pos_indptr, const int32_t *pos_flat, const int32_t *root_ok,
                              const int32_t *samples, int64_t *row_ptr, int32_t *center, int32_t *neighbor,
                              int32_t *label, int64_t *n_rows_out, void *stream) {
    GG_REQUIRE(row_ptr && n_rows_out, "null pointer");
    GG_REQUIRE(n_roots == 0 || (roots && walk_ptr && pos_indptr && pos_flat && root_ok && samples), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    if (n_roots > 0) {
        gg::row_count_kernel<<<(unsigned)((n_roots + threads - 1) / threads), threads, 0, st>>>(
            n_roots, (const long long *)walk_ptr, root_ok, (long long *)row_ptr);
        GG_CHECK(cudaGetLastError());
    }
    int rc = gg::launch_exclusive_scan_i64((long long *)row_ptr, n_roots, (long long *)n_rows_out, st);
    if (rc) return rc;
    if (n_roots == 0) return 0;
    GG_REQUIRE(center && neighbor && label, "null output pointer");
    gg::emit_rows_kernel<<<(unsigned)(gg::sm_count() * 8), threads, 0, st>>>(n_roots, roots, (const long long *)walk_ptr,
                                                               (const long long *)pos_indptr, pos_flat, root_ok,
                                                               samples, (const long long *)row_ptr, center, neighbor,
                                                               label);
    return gg::check_cuda(cudaGetLastError(), "emit rows launch");
}
