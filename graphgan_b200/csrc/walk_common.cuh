// walk_common.cuh -- building blocks shared by the walk sampler (walk.cu) and the per-pass
// precompute kernels (hub.cu): canonical scoring of a candidate list and the canonical
// softmax / CDF passes.  All of it is the arithmetic of DESIGN.md section 3.
#pragma once
#include "gg_common.cuh"

namespace gg {

constexpr int WARPS_PER_CTA = 8;
constexpr int ID_CAP = 320;    // candidate ids per warp kept in shared memory (longer lists: global scratch)
constexpr int SC_CAP = 2048;   // candidate scores per warp kept in shared memory
constexpr int SMEM_CAP = ID_CAP;
constexpr int WALK_SMEM_PER_WARP = SC_CAP * 4 + ID_CAP * 4;
constexpr int UNR = 4;         // tiles of 32 candidates in flight per pass iteration

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
    for (int i0 = 0; i0 < n; i0 += 8) {
        const int ia = i0 + grp, ib = i0 + 4 + grp;
        const bool va = ia < n, vb = ib < n;
        const int ca = va ? ids[ia] : fallback, cb = vb ? ids[ib] : fallback;
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
    for (int i0 = 0; i0 < n; i0 += 8) {
        const int ia = i0 + grp, ib = i0 + 4 + grp;
        const bool va = ia < n, vb = ib < n;
        const int ca = va ? __ldg(adj + e0 + ia) : fallback, cb = vb ? __ldg(adj + e0 + ib) : fallback;
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
__device__ __forceinline__ float softmax_exp_sum(float *sc, int n, float m, int lane) {
    float S = 0.0f;
    for (int t0 = 0; t0 < n; t0 += 32 * UNR) {
        float x[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) { const int i = t0 + 32 * u + lane; x[u] = (i < n) ? sc[i] : 0.0f; }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int i = t0 + 32 * u + lane;
            float e = 0.0f;
            if (i < n) { e = exp_c(__fsub_rn(x[u], m)); sc[i] = e; }
            S = __fadd_rn(S, warp_sum_butterfly(e));
        }
    }
    __syncwarp();
    return S;
}

// total of the float64 CDF over p_i = e_i / S
__device__ __forceinline__ double cdf_total(const float *sc, int n, float S, int lane) {
    double total = 0.0;
    for (int t0 = 0; t0 < n; t0 += 32 * UNR) {
        float e[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) { const int i = t0 + 32 * u + lane; e[u] = (i < n) ? sc[i] : 0.0f; }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            double x = (double)__fdiv_rn(e[u], S);   // 0 / S == 0 for the padding lanes
            x = warp_scan_ks(x, lane);
            total = __dadd_rn(total, __shfl_sync(FULL, x, 31));
        }
    }
    return total;
}

// first i with cdf_i / total > u; all lanes return it
__device__ __forceinline__ int cdf_pick(const float *sc, int n, float S, double total, double u, int lane) {
    double carry = 0.0;
    for (int t0 = 0; t0 < n; t0 += 32 * UNR) {
        float e[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k) { const int i = t0 + 32 * k + lane; e[k] = (i < n) ? sc[i] : 0.0f; }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            const int i = t0 + 32 * k + lane;
            double x = (double)__fdiv_rn(e[k], S);
            x = warp_scan_ks(x, lane);
            const double q = __ddiv_rn(__dadd_rn(carry, x), total);
            const unsigned hit = __ballot_sync(FULL, (i < n) && (q > u));
            if (hit) return t0 + 32 * k + __ffs(hit) - 1;
            carry = __dadd_rn(carry, __shfl_sync(FULL, x, 31));
        }
    }
    return n - 1;
}

// softmax + CDF + draw over sc[0..n) given its max m (== ggo_choose).  All lanes return the index.
__device__ __forceinline__ int choose_index(float *sc, int n, float m, double u, int lane) {
    const float S = softmax_exp_sum(sc, n, m, lane);
    const double total = cdf_total(sc, n, S, lane);
    return cdf_pick(sc, n, S, total, u, lane);
}

// normalised CDF q_i = cdf_i / total written out (the array numpy's choice would searchsorted)
__device__ __forceinline__ void cdf_store(float *sc, int n, double *q_out, int lane) {
    const float m = list_max(sc, n, lane);
    const float S = softmax_exp_sum(sc, n, m, lane);
    const double total = cdf_total(sc, n, S, lane);
    double carry = 0.0;
    for (int t0 = 0; t0 < n; t0 += 32) {
        const int i = t0 + lane;
        double x = (i < n) ? (double)__fdiv_rn(sc[i], S) : 0.0;
        x = warp_scan_ks(x, lane);
        if (i < n) q_out[i] = __ddiv_rn(__dadd_rn(carry, x), total);
        carry = __dadd_rn(carry, __shfl_sync(FULL, x, 31));
    }
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
