"""oracle/updates.py -- TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's TF1.8 model
graphs (pair score, losses, gradients, AdamOptimizer).

  * Discriminator  -- src/GraphGAN/discriminator.py:10-34
  * Generator      -- src/GraphGAN/generator.py:10-31
  * Adam           -- tf.train.AdamOptimizer (TF 1.8, not in the reference tree).  Its sparse
    path (_apply_sparse_duplicate_indices -> _apply_sparse_shared) first sums duplicate indices
    (unique + unsorted_segment_sum), then DECAYS m and v OVER ALL ROWS, scatter-adds the new
    gradient terms and updates ALL rows of the variable:
        lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
        m <- beta1*m ; m[idx] += (1-beta1) g ; v <- beta2*v ; v[idx] += (1-beta2) g^2
        var <- var - lr_t * m / (sqrt(v) + eps)
    This semantic comes from the TF 1.8 sources, which no reference test pins: "parity unpinned".

Gradients are derived by hand from the losses (no autograd dependency) and cross-checked
against torch.autograd in tests/test_oracle.py.  Never imported by graphgan_b200/.
"""
import numpy as np

F = np.float32


def sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F)


class AdamState:
    """One tf.train.AdamOptimizer instance over (embedding_matrix, bias_vector)."""

    def __init__(self, lr, n, d, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = F(lr), F(beta1), F(beta2), F(eps)
        self.m_e, self.v_e = np.zeros((n, d), F), np.zeros((n, d), F)
        self.m_b, self.v_b = np.zeros(n, F), np.zeros(n, F)
        self.b1_pow, self.b2_pow = F(beta1), F(beta2)  # beta^t with t = 1 for the first step

    def lr_t(self):
        return F(self.lr * np.sqrt(F(1) - self.b2_pow) / (F(1) - self.b1_pow))

    def apply(self, emb, bias, rows, g_rows, g_bias):
        """rows: unique ids; g_rows [U, d], g_bias [U] already summed over duplicates."""
        lr_t = self.lr_t()
        for var, m, v, g in ((emb, self.m_e, self.v_e, g_rows), (bias, self.m_b, self.v_b, g_bias)):
            m *= self.b1
            v *= self.b2
            m[rows] += (F(1) - self.b1) * g
            v[rows] += (g * g) * (F(1) - self.b2)         # TF: (grad * grad) * (1 - beta2_t)
            var -= lr_t * m / (np.sqrt(v) + self.eps)
        self.b1_pow = F(self.b1_pow * self.b1)
        self.b2_pow = F(self.b2_pow * self.b2)


class PairModel:
    """Shared part of Generator / Discriminator: embedding_matrix [N,d] fp32 + bias_vector [N]."""

    def __init__(self, n_node, node_emd_init, lr, lam, bias_init=None):
        self.E = np.asarray(node_emd_init, np.float64).astype(F).copy()
        self.b = np.zeros(n_node, F) if bias_init is None else np.asarray(bias_init, F).copy()
        self.lam = F(lam)
        self.adam = AdamState(lr, *self.E.shape)

    def score(self, i, j):
        i, j = np.asarray(i, np.int64), np.asarray(j, np.int64)
        return (np.sum(self.E[i] * self.E[j], axis=1, dtype=F) + self.b[j]).astype(F)

    def _segment(self, i, j, delta, bias_l2):
        """IndexedSlices of the loss wrt embedding rows / bias, duplicates summed."""
        i, j = np.asarray(i, np.int64), np.asarray(j, np.int64)
        ids = np.concatenate([i, j])
        contrib = np.concatenate([delta[:, None] * self.E[j] + self.lam * self.E[i],
                                  delta[:, None] * self.E[i] + self.lam * self.E[j]]).astype(F)
        bcon = delta + (self.lam * self.b[j] if bias_l2 else F(0))
        # unique in first-occurrence order (the GPU kernel's slot order); the order is irrelevant to Adam
        _, first = np.unique(ids, return_index=True)
        rows = ids[np.sort(first)]
        pos = {int(r): k for k, r in enumerate(rows)}
        g_rows = np.zeros((rows.shape[0], self.E.shape[1]), F)
        g_bias = np.zeros(rows.shape[0], F)
        for t, r in enumerate(ids):
            g_rows[pos[int(r)]] += contrib[t]
        for k, r in enumerate(j):
            g_bias[pos[int(r)]] += bcon[k]
        return rows, g_rows, g_bias


class Discriminator(PairModel):
    def grads(self, node_id, node_neighbor_id, label):
        s = self.score(node_id, node_neighbor_id)
        delta = (sigmoid(s) - np.asarray(label, F)).astype(F)      # d/ds sigmoid_xent (discriminator.py:26-27)
        return self._segment(node_id, node_neighbor_id, delta, bias_l2=True)   # + l2 on e_j, e_i, b_j (:28-30)

    def loss(self, node_id, node_neighbor_id, label):
        s = self.score(node_id, node_neighbor_id).astype(np.float64)
        y = np.asarray(label, np.float64)
        i, j = np.asarray(node_id, np.int64), np.asarray(node_neighbor_id, np.int64)
        xent = np.maximum(s, 0) - s * y + np.log1p(np.exp(-np.abs(s)))
        l2 = 0.5 * (np.sum(self.E[j].astype(np.float64) ** 2) + np.sum(self.E[i].astype(np.float64) ** 2)
                    + np.sum(self.b[j].astype(np.float64) ** 2))
        return float(xent.sum() + float(self.lam) * l2)

    def d_updates(self, node_id, node_neighbor_id, label):          # discriminator.py:31-32
        self.adam.apply(self.E, self.b, *self.grads(node_id, node_neighbor_id, label))

    def reward(self, node_id, node_neighbor_id):                    # discriminator.py:33-34
        s = np.clip(self.score(node_id, node_neighbor_id), -10, 10).astype(F)
        return np.log(F(1) + np.exp(s)).astype(F)


class Generator(PairModel):
    def grads(self, node_id, node_neighbor_id, reward):
        s = self.score(node_id, node_neighbor_id)
        p = sigmoid(s)
        r = np.asarray(reward, F)
        B = F(len(r))
        # loss = -mean(log(clip(p, 1e-5, 1)) * r) (generator.py:26-28); clip passes grad where p >= 1e-5
        delta = np.where(p >= F(1e-5), -(r / B) * (F(1) - p), F(0)).astype(F)
        return self._segment(node_id, node_neighbor_id, delta, bias_l2=False)   # l2 on e_j, e_i only (:28-29)

    def loss(self, node_id, node_neighbor_id, reward):
        s = self.score(node_id, node_neighbor_id).astype(np.float64)
        p = np.clip(1.0 / (1.0 + np.exp(-s)), 1e-5, 1.0)
        i, j = np.asarray(node_id, np.int64), np.asarray(node_neighbor_id, np.int64)
        l2 = 0.5 * (np.sum(self.E[j].astype(np.float64) ** 2) + np.sum(self.E[i].astype(np.float64) ** 2))
        return float(-np.mean(np.log(p) * np.asarray(reward, np.float64)) + float(self.lam) * l2)

    def g_updates(self, node_id, node_neighbor_id, reward):          # generator.py:30-31
        self.adam.apply(self.E, self.b, *self.grads(node_id, node_neighbor_id, reward))

    def all_score(self):                                             # generator.py:21
        return (self.E @ self.E.T + self.b).astype(F)
