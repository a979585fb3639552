// adam.cu -- K3: TF1.8 AdamOptimizer "sparse" apply, which is dense (sm_100a).
//
// generator.py:30-31 / discriminator.py:31-32 call tf.train.AdamOptimizer(lr).minimize(loss)
// on variables whose gradients are IndexedSlices.  TF 1.8's _apply_sparse_shared does
//     m <- m * beta1;  m[idx] += (1 - beta1) * g          (every row decays)
//     v <- v * beta2;  v[idx] += (1 - beta2) * g * g
//     var <- var - lr_t * m / (sqrt(v) + eps)              (every row moves)
// with lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) computed by the host wrapper.  So one
// 64-pair step streams all of E, m, v: 24 * N * ld bytes -- a pure HBM-bandwidth kernel.
// A warp sweeps 512-byte segments, two in flight, 32 warps per SM (update_dev.cuh: adam_rows; the IEEE div / sqrt
// chains of the update are what the warps wait on -- ncu: fixed-latency dependency stalls -- so occupancy beats
// deeper unrolling: 4 in flight at 112 registers ran at 45 % of DRAM peak); the row -> gradient-slot map
// written by gg_pair_grad tells whether the row has a gradient, and is reset here.
#include "update_dev.cuh"

namespace gg {
namespace {

__global__ void __launch_bounds__(256, 4) adam_kernel(long long n_node, int ld, float *__restrict__ emb,
                                                   float *__restrict__ m_emb, float *__restrict__ v_emb,
                                                   float *__restrict__ bias, float *__restrict__ m_bias,
                                                   float *__restrict__ v_bias, const float *__restrict__ grad_rows,
                                                   const float *__restrict__ grad_bias, int *__restrict__ row_slot,
                                                   float lr_t, float b1, float b2, float eps) {
    adam_rows<false, 2>(n_node, ld, emb, m_emb, v_emb, bias, m_bias, v_bias, grad_rows, grad_bias, row_slot, lr_t, b1, b2, eps);
}

}  // namespace
}  // namespace gg

extern "C" int gg_adam_apply(int64_t n_node, int32_t ld, float *emb, float *m_emb, float *v_emb, float *bias,
                             float *m_bias, float *v_bias, const int32_t *n_unique, const int32_t *uniq_ids,
                             const float *grad_rows, const float *grad_bias, int32_t *row_slot, float lr_t, float beta1,
                             float beta2, float eps, void *stream) {
    (void)n_unique; (void)uniq_ids;
    GG_REQUIRE(emb && m_emb && v_emb && bias && m_bias && v_bias && grad_rows && grad_bias && row_slot, "null pointer");
    GG_REQUIRE(ld == 32 || ld == 64 || ld == 128 || ld == 256, "ld must be 32, 64, 128 or 256 (row stride in floats)");
    if (n_node == 0) return 0;
    const int q = ld / 4;                                                       // float4 per row
    const long long nseg = q >= 32 ? n_node * (q / 32) : (n_node + 32 / q - 1) / (32 / q);   // 512-byte segments
    long long blocks = (nseg + 2 * 8 - 1) / (2 * 8);                            // 8 warps x 2 segments in flight
    const long long cap = (long long)gg::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    gg::adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n_node, ld, emb, m_emb, v_emb, bias, m_bias,
                                                                        v_bias, grad_rows, grad_bias, row_slot, lr_t,
                                                                        beta1, beta2, eps);
    return gg::check_cuda(cudaGetLastError(), "adam kernel launch");
}
