// steps.cu -- the reference's per-batch loop (graph_gan.py:149-157, 168-176) driven from C.
//
// One 64-pair step is two tiny launches (K2 pair_grad, K3 adam); at config C1 an epoch is ~230 k of them, so
// the host-language overhead per step (argument marshalling, tensor slicing) would dominate.  This entry point
// walks the caller's shuffled start list and enqueues every step on the stream; nothing synchronises.
#include <math.h>

#include "gg_common.cuh"

extern "C" int gg_train_steps(int32_t mode, int64_t n_rows, const int64_t *start_list, int64_t n_starts, int32_t batch_size,
                              const int32_t *node_id, const int32_t *node_neighbor_id, const float *aux, int64_t n_node,
                              int32_t ld, float *emb, float *m_emb, float *v_emb, float *bias, float *m_bias, float *v_bias,
                              float lambda, int32_t *n_unique, int32_t *uniq_ids, float *grad_rows, float *grad_bias,
                              int32_t *row_slot, float lr, float beta1, float beta2, float eps, float *beta1_power,
                              float *beta2_power, void *stream) {
    GG_REQUIRE(start_list && beta1_power && beta2_power, "null host pointer");
    GG_REQUIRE(batch_size > 0 && batch_size <= GG_MAX_BATCH, "batch size out of range");
    for (int64_t s = 0; s < n_starts; ++s) {
        const int64_t start = start_list[s];
        GG_REQUIRE(start >= 0 && start < n_rows, "start out of range");
        const int64_t end = start + batch_size < n_rows ? start + batch_size : n_rows;
        int rc = gg_pair_grad(mode, (int32_t)(end - start), 0, node_id + start, node_neighbor_id + start, aux + start, emb, bias,
                              ld, lambda, n_unique, uniq_ids, grad_rows, grad_bias, row_slot, stream);
        if (rc) return rc;
        // lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), fp32 step by step like the TF graph (and model.py)
        volatile float one_m_b2 = 1.0f - *beta2_power;
        volatile float root = sqrtf(one_m_b2);
        volatile float num = lr * root;
        volatile float den = 1.0f - *beta1_power;
        const float lr_t = num / den;
        rc = gg_adam_apply(n_node, ld, emb, m_emb, v_emb, bias, m_bias, v_bias, n_unique, uniq_ids, grad_rows, grad_bias,
                           row_slot, lr_t, beta1, beta2, eps, stream);
        if (rc) return rc;
        volatile float p1 = *beta1_power * beta1, p2 = *beta2_power * beta2;
        *beta1_power = p1;
        *beta2_power = p2;
    }
    return 0;
}
