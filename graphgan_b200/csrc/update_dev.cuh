// update_dev.cuh -- device bodies of K2 (mini-batch gradient) and K3 (TF1.8 dense Adam sweep), shared by the
// stand-alone kernels (pairs.cu, adam.cu) and the persistent step loop (steps.cu).  See pairs.cu / adam.cu for
// the reference semantics each one restates.
#pragma once
#include "gg_common.cuh"

namespace gg {

// ---------------------------------------------------------------- mini-batch gradient (1 CTA)
constexpr int GRAD_THREADS = 1024;

// COH = true: the parameters may have been written earlier in the SAME kernel by other SMs (persistent step
// loop): read them through L2 (ld.global.cg) instead of the non-coherent read-only path.
template <bool COH> __device__ __forceinline__ float4 row4(const float *p) {
    if constexpr (COH) return __ldcg(reinterpret_cast<const float4 *>(p));
    else return ldg4(p);
}
template <bool COH> __device__ __forceinline__ float ldf(const float *p) {
    if constexpr (COH) return __ldcg(p);
    else return __ldg(p);
}
template <bool COH>
__device__ __forceinline__ float group_dot_t(const float *a, const float *b, int ld, int g) {
    float s = 0.0f;
    for (int c = 4 * g; c < ld; c += 32) s = fma4(row4<COH>(a + c), row4<COH>(b + c), s);
    return group8_sum(s);
}

// One mini-batch gradient, executed by one CTA of GRAD_THREADS threads; smem: 5 * B ints.
template <bool COH>
__device__ __forceinline__ void pair_grad_body(int *smem, int mode, int B, int batch_total, const int *__restrict__ ni,
                                               const int *__restrict__ nj, const float *__restrict__ aux,
                                               const float *emb, const float *bias, int ld, float lambda, int *n_unique,
                                               int *uniq_ids, float *grad_rows, float *grad_bias, int *row_slot) {
    int *ids = smem;              // [2B]  entry -> row id (i-side entries first, then j-side)
    int *slot = ids + 2 * B;      // [2B]  entry -> unique slot
    float *delta = reinterpret_cast<float *>(slot + 2 * B);  // [B] dL/dscore_k
    __shared__ int s_warp[32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, grp = lane >> 3, g = lane & 7;
    const int E = 2 * B;
    for (int t = tid; t < E; t += GRAD_THREADS) ids[t] = (t < B) ? ni[t] : nj[t - B];
    // ---- forward: score and dL/dscore
    for (int k0 = wid * 4; k0 < B; k0 += (GRAD_THREADS / 32) * 4) {
        const int k = k0 + grp;
        const bool valid = k < B;
        const int i = valid ? ni[k] : 0, j = valid ? nj[k] : 0;
        float s = group_dot_t<COH>(emb + (size_t)i * ld, emb + (size_t)j * ld, ld, g);
        if (valid && g == 0) {
            s = __fadd_rn(s, ldf<COH>(bias + j));
            const float p = (float)(1.0 / (1.0 + exp(-(double)s)));   // sigmoid (B values: fp64 costs nothing)
            float d;
            if (mode == 0) {
                d = p - aux[k];                          // d/ds sigmoid_xent(label, s) = sigmoid(s) - label
            } else {
                // d/ds [-(1/B) r log(clip(p,1e-5,1))] = -(r/B)(1-p) where the clip passes (p >= 1e-5)
                d = (p >= 1e-5f) ? -(aux[k] / (float)batch_total) * (1.0f - p) : 0.0f;
            }
            delta[k] = d;
        }
    }
    __syncthreads();
    // ---- unique: first occurrence of every row id gets a slot, in entry order
    // phase A: 8 lanes per entry look for the earliest equal id (strided scan, min over the group)
    for (int t0 = 0; t0 < E; t0 += GRAD_THREADS / 8) {
        const int t = t0 + (tid >> 3);
        int f = E;
        if (t < E) {
            const int id = ids[t];
            for (int q = g; q < t; q += 8) if (ids[q] == id) { f = q; break; }
        }
        f = min(f, __shfl_xor_sync(FULL, f, 4));
        f = min(f, __shfl_xor_sync(FULL, f, 2));
        f = min(f, __shfl_xor_sync(FULL, f, 1));
        if (t < E && g == 0) slot[t] = min(f, t);
    }
    __syncthreads();
    // phase B: exclusive scan of the first-occurrence flags (one entry per thread per round)
    int is_first = 0, first_t = 0;
    int base_total = 0;
    for (int t0 = 0; t0 < E; t0 += GRAD_THREADS) {
        const int t = t0 + tid;
        is_first = 0; first_t = t;
        if (t < E) { first_t = slot[t]; is_first = (first_t == t); }   // (each thread rewrites only its own slot[t])
        int x = is_first;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(FULL, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int v = s_warp[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int y = __shfl_up_sync(FULL, v, off);
                if (lane >= off) v += y;
            }
            s_warp[lane] = v;
        }
        __syncthreads();
        const int excl = base_total + (wid ? s_warp[wid - 1] : 0) + x - is_first;
        if (t < E) slot[t] = is_first ? excl : -1 - first_t;  // non-first: remember where the first is
        if (t < E && is_first) { uniq_ids[excl] = ids[t]; row_slot[ids[t]] = excl; }
        base_total += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) { s_total = base_total; *n_unique = base_total; }
    __syncthreads();
    for (int t = tid; t < E; t += GRAD_THREADS) if (slot[t] < 0) { const int f = -1 - slot[t]; slot[t] = slot[f] < 0 ? -1 : slot[f]; }
    __syncthreads();
    const int U = s_total;
    // ---- segment sums: slot u accumulates its entries in entry order
    for (int u = wid; u < U; u += GRAD_THREADS / 32) {
        const int row = uniq_ids[u];
        const float *erow = emb + (size_t)row * ld;
        const float bself = ldf<COH>(bias + row);
        float gb = 0.0f;
        for (int c0 = 0; c0 < ld; c0 += 128) {
            const int c = c0 + 4 * lane;
            const bool on = c < ld;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 self = on ? row4<COH>(erow + c) : acc;
            for (int tb = 0; tb < E; tb += 32) {          // this slot's entries, found 32 at a time, in entry order
                unsigned m = __ballot_sync(FULL, tb + lane < E && slot[tb + lane] == u);
                while (m) {
                    const int t = tb + __ffs(m) - 1;
                    m &= m - 1;
                    const int k = (t < B) ? t : t - B;
                    const int other = (t < B) ? nj[k] : ni[k];
                    const float d = delta[k];
                    if (on) {
                        const float4 o = row4<COH>(emb + (size_t)other * ld + c);
                        // d(score)/d(this row) = other row;  l2 term: lambda * this row, once per occurrence.
                        // Explicit mul/mul/add/add (no fma contraction): the same op sequence as the IndexedSlices
                        // sum of the numpy oracle, so cancellation noise in near-zero coordinates stays comparable.
#define GG_ACC(f) acc.f = __fadd_rn(acc.f, __fadd_rn(__fmul_rn(d, o.f), __fmul_rn(lambda, self.f)))
                        GG_ACC(x); GG_ACC(y); GG_ACC(z); GG_ACC(w);
#undef GG_ACC
                    }
                    if (c0 == 0 && t >= B)   // bias gradient: j-side entries only; generator.py:28-29 has no bias l2
                        gb = __fadd_rn(gb, mode == 0 ? __fadd_rn(d, __fmul_rn(lambda, bself)) : d);
                }
            }
            if (on) *reinterpret_cast<float4 *>(grad_rows + (size_t)u * ld + c) = acc;
        }
        if (lane == 0) grad_bias[u] = gb;
    }
}


// The dense Adam sweep over all rows, executed by the whole grid (warp per row).  COH: the gradient slots were
// written earlier in the same kernel by another SM -> read them through L2.
template <bool COH>
__device__ __forceinline__ void adam_rows(long long n_node, int ld, float *emb, float *m_emb, float *v_emb, float *bias,
                                          float *m_bias, float *v_bias, const float *grad_rows, const float *grad_bias,
                                          int *row_slot, float lr_t, float b1, float b2, float eps) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
    for (long long row = warp; row < n_node; row += nwarps) {
        int slot = -1;
        if (lane == 0) slot = COH ? __ldcg(row_slot + row) : row_slot[row];
        slot = __shfl_sync(FULL, slot, 0);
        const size_t ro = (size_t)row * ld;
        for (int c = 4 * lane; c < ld; c += 128) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (slot >= 0) g = COH ? __ldcg(reinterpret_cast<const float4 *>(grad_rows + (size_t)slot * ld + c))
                                : *reinterpret_cast<const float4 *>(grad_rows + (size_t)slot * ld + c);
            float4 m = *reinterpret_cast<float4 *>(m_emb + ro + c);
            float4 v = *reinterpret_cast<float4 *>(v_emb + ro + c);
            float4 x = *reinterpret_cast<float4 *>(emb + ro + c);
// TF1.8 op order (assign m*b1; scatter_add (1-b1)*g; ... var -= lr*m/(sqrt(v)+eps)), no contraction
#define GG_ADAM1(f)                                                                                   \
    m.f = __fadd_rn(__fmul_rn(m.f, b1), __fmul_rn(omb1, g.f));                                        \
    v.f = __fadd_rn(__fmul_rn(v.f, b2), __fmul_rn(__fmul_rn(omb2, g.f), g.f));                        \
    x.f = __fsub_rn(x.f, __fdiv_rn(__fmul_rn(lr_t, m.f), __fadd_rn(__fsqrt_rn(v.f), eps)));
            GG_ADAM1(x) GG_ADAM1(y) GG_ADAM1(z) GG_ADAM1(w)
#undef GG_ADAM1
            *reinterpret_cast<float4 *>(m_emb + ro + c) = m;
            *reinterpret_cast<float4 *>(v_emb + ro + c) = v;
            *reinterpret_cast<float4 *>(emb + ro + c) = x;
        }
        if (lane == 0) {
            const float g = slot >= 0 ? (COH ? __ldcg(grad_bias + slot) : grad_bias[slot]) : 0.0f;
            const float m = __fadd_rn(__fmul_rn(m_bias[row], b1), __fmul_rn(omb1, g));
            const float v = __fadd_rn(__fmul_rn(v_bias[row], b2), __fmul_rn(__fmul_rn(omb2, g), g));
            m_bias[row] = m; v_bias[row] = v;
            bias[row] = __fsub_rn(bias[row], __fdiv_rn(__fmul_rn(lr_t, m), __fadd_rn(__fsqrt_rn(v), eps)));
            if (slot >= 0) row_slot[row] = -1;
        }
    }
}


}  // namespace gg
