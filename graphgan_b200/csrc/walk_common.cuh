// walk_common.cuh -- building blocks shared by the walk sampler (walk.cu) and the per-pass
// precompute kernels (hub.cu): canonical scoring of a candidate list and the canonical
// softmax / CDF passes.  All of it is the arithmetic of DESIGN.md section 3.
#pragma once
#include "gg_common.cuh"

namespace gg {

// Tuning knobs (compile-time; tools/variants.py builds A/B libraries with -DGG_...): the walk kernels are latency
// bound, so the trade is occupancy (registers, shared memory per warp) against loads in flight per warp (unrolling)
// and against instruction footprint (the hot path must stay near the 32 KB L1.5 instruction cache).
#ifndef GG_SC_CAP
#define GG_SC_CAP 1024
#endif
#ifndef GG_UNR
#define GG_UNR 2
#endif
#ifndef GG_UNR_S1
#define GG_UNR_S1 8
#endif
#ifndef GG_WALK_MIN_CTAS
#define GG_WALK_MIN_CTAS 4
#endif
constexpr int WARPS_PER_CTA = 8;
constexpr int WALK_MIN_CTAS = GG_WALK_MIN_CTAS;   // CTAs per SM the walk kernels are compiled and launched for
constexpr int ID_CAP = 320;    // candidate ids per warp kept in shared memory (longer lists: global scratch)
constexpr int SC_CAP = GG_SC_CAP;   // candidate scores per warp kept in shared memory
constexpr int SMEM_CAP = ID_CAP;
constexpr int WALK_SMEM_PER_WARP = SC_CAP * 4 + ID_CAP * 4 + 16;   // + the warp's mbarrier (TMA staging of hub lists)
constexpr int STAGE_ENTRIES = 512;  // adjacency entries per bulk-copy block: 2 KB of ids + 2 KB of cached scores
static_assert(2 * STAGE_ENTRIES * 4 <= SC_CAP * 4, "the staging area is the warp's (idle) score buffer");

// Per-warp TMA staging state: `buf` = the warp's shared score buffer (free whenever a list is too long for it), `bar` =
// the warp's mbarrier, `phase` = its parity.  on = false: plain loads (stream-replay kernel, or desc.no_tma).
struct Stage {
    float *buf;
    unsigned long long *bar;
    unsigned phase;
    bool on;
};
constexpr int UNR = GG_UNR;    // tiles of 32 candidates in flight per pass iteration (walk kernel: short lists, many warps)
constexpr int UNR_S1 = GG_UNR_S1;   // same, for the per-pass list builders (step1_cdf_kernel: long lists)

// cur row in registers: lane (grp, g) holds float4 chunks g, g+8, ... (replicated over the 4 groups)
template <int CPL>
__device__ __forceinline__ void load_row(const float *__restrict__ emb, int ld, int node, int g, float4 (&c4)[CPL]) {
    const float *crow = emb + (size_t)node * (size_t)ld + 4 * g;
#pragma unroll
    for (int c = 0; c < CPL; ++c) c4[c] = ldg4(crow + 32 * c);
}

// sc[i] = dot(cur, emb[ids[i]]) + bias[ids[i]] for i in [0, n): all_score[cur, cand] (generator.py:21).
// Four 8-lane groups, two candidate rows in flight per group.
template <int CPL>
__device__ __forceinline__ void score_list(const float *__restrict__ emb, const float *__restrict__ bias, int ld,
                                           const float4 (&c4)[CPL], const int *ids, float *sc, int n, int fallback,
                                           int lane) {
    const int grp = lane >> 3, g = lane & 7;
    // (the ids of the NEXT eight candidates are fetched while this iteration's rows are in flight: the id -> row
    // dependency costs one exposed round trip per list instead of one per iteration)
    int ca_n = (grp < n) ? ids[grp] : fallback, cb_n = (4 + grp < n) ? ids[4 + grp] : fallback;
    for (int i0 = 0; i0 < n; i0 += 8) {
        const int ia = i0 + grp, ib = i0 + 4 + grp;
        const bool va = ia < n, vb = ib < n;
        const int ca = ca_n, cb = cb_n;
        ca_n = (ia + 8 < n) ? ids[ia + 8] : fallback;
        cb_n = (ib + 8 < n) ? ids[ib + 8] : fallback;
        const float *ra = emb + (size_t)ca * (size_t)ld + 4 * g;
        const float *rb = emb + (size_t)cb * (size_t)ld + 4 * g;
        float4 xa[CPL], xb[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) xa[c] = ldg4(ra + 32 * c);
#pragma unroll
        for (int c = 0; c < CPL; ++c) xb[c] = ldg4(rb + 32 * c);
        const float ba = __ldg(bias + ca), bb = __ldg(bias + cb);
        float sa = 0.0f, sb = 0.0f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) sa = fma4(c4[c], xa[c], sa);
#pragma unroll
        for (int c = 0; c < CPL; ++c) sb = fma4(c4[c], xb[c], sb);
        sa = group8_sum(sa);
        sb = group8_sum(sb);
        if (g == 0) {
            if (va) sc[ia] = __fadd_rn(sa, ba);
            if (vb) sc[ib] = __fadd_rn(sb, bb);
        }
    }
    __syncwarp();
}

// Same, for a contiguous run of adjacency entries (ids read straight from the CSR).
template <int CPL>
__device__ __forceinline__ void score_edges(const float *__restrict__ emb, const float *__restrict__ bias, int ld,
                                            const float4 (&c4)[CPL], const int *__restrict__ adj, long long e0,
                                            int n, float *out, int fallback, int lane) {
    const int grp = lane >> 3, g = lane & 7;
    int ca_n = (grp < n) ? __ldg(adj + e0 + grp) : fallback, cb_n = (4 + grp < n) ? __ldg(adj + e0 + 4 + grp) : fallback;
    for (int i0 = 0; i0 < n; i0 += 8) {
        const int ia = i0 + grp, ib = i0 + 4 + grp;
        const bool va = ia < n, vb = ib < n;
        const int ca = ca_n, cb = cb_n;          // (next iteration's ids in flight behind this iteration's rows, see score_list)
        ca_n = (ia + 8 < n) ? __ldg(adj + e0 + ia + 8) : fallback;
        cb_n = (ib + 8 < n) ? __ldg(adj + e0 + ib + 8) : fallback;
        const float *ra = emb + (size_t)ca * (size_t)ld + 4 * g;
        const float *rb = emb + (size_t)cb * (size_t)ld + 4 * g;
        float4 xa[CPL], xb[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) xa[c] = ldg4(ra + 32 * c);
#pragma unroll
        for (int c = 0; c < CPL; ++c) xb[c] = ldg4(rb + 32 * c);
        const float ba = __ldg(bias + ca), bb = __ldg(bias + cb);
        float sa = 0.0f, sb = 0.0f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) sa = fma4(c4[c], xa[c], sa);
#pragma unroll
        for (int c = 0; c < CPL; ++c) sb = fma4(c4[c], xb[c], sb);
        sa = group8_sum(sa);
        sb = group8_sum(sb);
        if (g == 0) {
            if (va) out[ia] = __fadd_rn(sa, ba);
            if (vb) out[ib] = __fadd_rn(sb, bb);
        }
    }
    __syncwarp();
}

// max over sc[0..n)
__device__ __forceinline__ float list_max(const float *sc, int n, int lane) {
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, sc[i]);
    return warp_max(m);
}

// softmax numerators in place (sc[i] <- e_i = exp_c(s_i - m)) and their canonical sum
// S = T_0 + T_1 + ... (tile sums by butterfly; 0 + T_0 == T_0 and S + 0 == S exactly, so empty
// tiles of the unrolled tail are harmless).  UNR tiles are in flight per iteration.
template <int U = UNR>
__device__ __forceinline__ float softmax_exp_sum(float *sc, int n, float m, int lane) {
    float S = 0.0f;
    for (int t0 = 0; t0 < n; t0 += 32 * U) {
        float x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = t0 + 32 * u + lane; x[u] = (i < n) ? sc[i] : 0.0f; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (t0 + 32 * u >= n) break;            // warp-uniform: no work for the empty tiles of a short list
            const int i = t0 + 32 * u + lane;
            float e = 0.0f;
            if (i < n) { e = exp_c(__fsub_rn(x[u], m)); sc[i] = e; }
            S = __fadd_rn(S, warp_sum_butterfly(e));
        }
    }
    __syncwarp();
    return S;
}

// total of the float64 CDF over p_i = e_i / S.  Lane (t & 31) also keeps the running total after tile t
// in car[t >> 5] (tiles 0..63), so that the draw can jump straight to the tile that contains it.
template <int U = UNR>
__device__ __forceinline__ double cdf_total(const float *sc, int n, float S, int lane, double (&car)[2]) {
    double total = 0.0;
    car[0] = car[1] = 0.0;
    for (int t0 = 0; t0 < n; t0 += 32 * U) {
        float e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = t0 + 32 * u + lane; e[u] = (i < n) ? sc[i] : 0.0f; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (t0 + 32 * u >= n) break;            // warp-uniform
            double x = (double)__fdiv_rn(e[u], S);   // 0 / S == 0 for the padding lanes
            x = warp_scan_ks(x, lane);
            total = __dadd_rn(total, __shfl_sync(FULL, x, 31));
            const int t = (t0 >> 5) + u;
            if (t < 64 && lane == (t & 31)) car[t >> 5] = total;
        }
    }
    return total;
}

// first i with cdf_i / total > u, scanning tiles [t_begin, ...) with `carry` = CDF before tile t_begin
__device__ __forceinline__ int cdf_pick_from(const float *sc, int n, float S, double total, double u, int lane,
                                             int t_begin, double carry) {
    for (int t0 = 32 * t_begin; t0 < n; t0 += 32) {
        const int i = t0 + lane;
        double x = (double)__fdiv_rn((i < n) ? sc[i] : 0.0f, S);
        x = warp_scan_ks(x, lane);
        const double q = __ddiv_rn(__dadd_rn(carry, x), total);
        const unsigned hit = __ballot_sync(FULL, (i < n) && (q > u));
        if (hit) return t0 + __ffs(hit) - 1;
        carry = __dadd_rn(carry, __shfl_sync(FULL, x, 31));
    }
    return n - 1;
}

// The CDF is non-decreasing, so the first index with q > u lies in the first tile whose LAST q exceeds u,
// and that last q is exactly car[t] / total (same operations as the linear scan performs).  All lanes return idx.
__device__ __forceinline__ int cdf_pick(const float *sc, int n, float S, double total, double u, int lane,
                                        const double (&car)[2]) {
    const int ntiles = (n + 31) >> 5;
    int t_hit = -1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int t = 32 * h + lane;
        const bool p = (t < ntiles) && (__ddiv_rn(car[h], total) > u);
        const unsigned mk = __ballot_sync(FULL, p);
        if (t_hit < 0 && mk) t_hit = 32 * h + __ffs(mk) - 1;
    }
    if (t_hit < 0) {   // beyond the 64 tracked tiles (n > 2048): linear scan from tile 64
        if (ntiles <= 64) return n - 1;
        const double c63 = __shfl_sync(FULL, car[1], 31);
        return cdf_pick_from(sc, n, S, total, u, lane, 64, c63);
    }
    double before = 0.0;
    if (t_hit > 0) {
        const int tp = t_hit - 1;
        const double lo = __shfl_sync(FULL, car[0], tp & 31), hi = __shfl_sync(FULL, car[1], tp & 31);
        before = (tp >> 5) ? hi : lo;
    }
    return cdf_pick_from(sc, n, S, total, u, lane, t_hit, before);
}

// softmax + CDF + draw over sc[0..n) given its max m (== ggo_choose).  All lanes return the index.
// Long lists (global scratch): running total after EVERY tile goes to `tiles` (the warp's idle shared score buffer
// viewed as doubles), so the draw can locate its tile among hundreds without a linear scan.
__device__ __forceinline__ double cdf_total_tiles(const float *sc, int n, float S, int lane, double *tiles) {
    constexpr int U = 8;
    double total = 0.0;
    for (int t0 = 0; t0 < n; t0 += 32 * U) {
        float e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = t0 + 32 * u + lane; e[u] = (i < n) ? sc[i] : 0.0f; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (t0 + 32 * u >= n) break;
            double x = (double)__fdiv_rn(e[u], S);
            x = warp_scan_ks(x, lane);
            total = __dadd_rn(total, __shfl_sync(FULL, x, 31));
            if (lane == 0) tiles[(t0 >> 5) + u] = total;
        }
    }
    __syncwarp();
    return total;
}
__device__ __forceinline__ int cdf_pick_tiles(const float *sc, int n, float S, double total, double u, int lane,
                                              const double *tiles) {
    const int ntiles = (n + 31) >> 5;
    for (int tb = 0; tb < ntiles; tb += 32) {
        const int t = tb + lane;
        const bool p = (t < ntiles) && (__ddiv_rn(tiles[t], total) > u);
        const unsigned mk = __ballot_sync(FULL, p);
        if (mk) {
            const int t_hit = tb + __ffs(mk) - 1;
            return cdf_pick_from(sc, n, S, total, u, lane, t_hit, t_hit > 0 ? tiles[t_hit - 1] : 0.0);
        }
    }
    return n - 1;
}

// lists longer than the shared score buffer (global scratch): rare, so kept out of line and modestly unrolled -- the
// walk kernel's instruction footprint is what its warps stall on otherwise (ncu: "no instruction")
static __device__ __noinline__ int choose_index_long(float *sc, int n, float m, double u, int lane, double *tiles) {
    float S;
    double car[2], total;
    if (tiles && ((n + 31) >> 5) <= SC_CAP / 2) {
        // per-tile running totals in the (idle) shared score buffer: the draw finds its tile among hundreds directly
        S = softmax_exp_sum<8>(sc, n, m, lane);
        total = cdf_total_tiles(sc, n, S, lane, tiles);
        return cdf_pick_tiles(sc, n, S, total, u, lane, tiles);
    }
    S = softmax_exp_sum<8>(sc, n, m, lane);
    total = cdf_total<8>(sc, n, S, lane, car);
    return cdf_pick(sc, n, S, total, u, lane, car);
}

__device__ __forceinline__ int choose_index(float *sc, int n, float m, double u, int lane, double *tiles = nullptr) {
    if (n <= 32) {
        // one tile (the common case): S = 0 + T_0 = T_0, total = 0 + scan_31 = scan_31 and the carry of the draw is 0, so
        // the canonical sequence collapses to ONE scan and ONE division per lane -- the same floats, bit for bit
        const float e = (lane < n) ? exp_c(__fsub_rn(sc[lane], m)) : 0.0f;
        const float S = warp_sum_butterfly(e);
        const double x = warp_scan_ks((double)__fdiv_rn(e, S), lane);
        const double total = __shfl_sync(FULL, x, 31);
        const unsigned hit = __ballot_sync(FULL, (lane < n) && (__ddiv_rn(x, total) > u));
        return hit ? __ffs(hit) - 1 : n - 1;
    }
    if (n > SC_CAP) return choose_index_long(sc, n, m, u, lane, tiles);
    double car[2];
    const float S = softmax_exp_sum(sc, n, m, lane);
    const double total = cdf_total(sc, n, S, lane, car);
    return cdf_pick(sc, n, S, total, u, lane, car);
}

// normalised CDF q_i = cdf_i / total written out (the array numpy's choice would searchsorted)
__device__ __forceinline__ void cdf_store(float *sc, int n, double *q_out, int lane) {
    const float m = list_max(sc, n, lane);
    const float S = softmax_exp_sum<UNR_S1>(sc, n, m, lane);
    double car[2];
    const double total = cdf_total<UNR_S1>(sc, n, S, lane, car);
    double carry = 0.0;
    for (int t0 = 0; t0 < n; t0 += 32) {
        const int i = t0 + lane;
        double x = (i < n) ? (double)__fdiv_rn(sc[i], S) : 0.0;
        x = warp_scan_ks(x, lane);
        if (i < n) q_out[i] = __ddiv_rn(__dadd_rn(carry, x), total);
        carry = __dadd_rn(carry, __shfl_sync(FULL, x, 31));
    }
}

// un-normalised CDF c_i = carry + scan(e/S)_i written out, the total after the n entries (c[n] = total).  One scan
// pass: the division by the total is left to the search, which performs it only on the entries it probes.
template <int U>
__device__ __forceinline__ void cdf_store_raw(float *sc, int n, float m, double *c_out, int lane) {
    const float S = softmax_exp_sum<U>(sc, n, m, lane);
    double carry = 0.0;
    for (int t0 = 0; t0 < n; t0 += 32 * U) {
        double x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = t0 + 32 * u + lane;
            x[u] = (i < n) ? (double)__fdiv_rn(sc[i], S) : 0.0;   // 0 / S == 0 for the padding lanes
        }
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = warp_scan_ks(x[u], lane);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (t0 + 32 * u >= n) break;               // warp-uniform
            const int i = t0 + 32 * u + lane;
            if (i < n) c_out[i] = __dadd_rn(carry, x[u]);
            carry = __dadd_rn(carry, __shfl_sync(FULL, x[u], 31));
        }
    }
    if (lane == 0) c_out[n] = carry;
}

// first i in [0,n) with c[i] / c[n] > u: the same comparison as cdf_pick / cdf_search, on the raw array
__device__ __forceinline__ int cdf_search_raw(const double *__restrict__ c, int n, double u) {
    const double total = __ldg(c + n);
    int lo = 0, hi = n - 1;  // invariant: answer in [lo, hi]  (c[n-1] == total, so c[n-1] / total == 1 > u)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ddiv_rn(__ldg(c + mid), total) > u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// first i in [0,n) with q[i] > u (q non-decreasing, q[n-1] == 1 > u): what the linear scan finds
__device__ __forceinline__ int cdf_search(const double *__restrict__ q, int n, double u) {
    int lo = 0, hi = n - 1;  // invariant: answer in [lo, hi]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(q + mid) > u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

}  // namespace gg
