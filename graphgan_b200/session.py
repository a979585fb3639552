"""``Session.run(fetch, feed_dict)``: the reference's plugin boundary.

The reference crosses into TensorFlow at exactly five call sites (src/GraphGAN/graph_gan.py:154,
173, 220, 238, 298).  This shim keeps that call shape so reference-style code keeps working; each
fetch dispatches to the CUDA entry point that replaces the TF sub-graph.  Feeds are numpy arrays
or Python lists as in the reference (copied host->device here) or device tensors (no copy).
"""
import numpy as np

from .model import Fetch, Placeholder


class Session:
    def __init__(self, config=None):
        self.config = config

    def run(self, fetch, feed_dict=None):
        if isinstance(fetch, (list, tuple)):
            return [self.run(f, feed_dict) for f in fetch]
        if fetch is None or getattr(fetch, "kind", None) == "init_op":
            return None
        if not isinstance(fetch, Fetch) or isinstance(fetch, Placeholder):
            raise TypeError("cannot fetch %r" % (fetch,))
        m, feed = fetch.owner, (feed_dict or {})
        get = lambda ph: feed[ph] if ph in feed else _missing(ph)
        k = fetch.kind
        if k == "d_updates":       # graph_gan.py:154-157
            m.step(get(m.node_id), get(m.node_neighbor_id), get(m.label))
            return None
        if k == "g_updates":       # graph_gan.py:173-176
            m.step(get(m.node_id), get(m.node_neighbor_id), get(m.reward))
            return None
        if k == "reward":          # graph_gan.py:220-222
            return m.reward_pairs(get(m.node_id), get(m.node_neighbor_id)).cpu().numpy()
        if k == "all_score":       # graph_gan.py:238
            return m.all_score_matrix().cpu().numpy()
        if k == "embedding_matrix":  # graph_gan.py:298
            return m.embedding_numpy()
        if k == "bias_vector":
            return m.bias_t.cpu().numpy()
        # compatibility-only fetches: attributes the reference defines but never runs
        torch = m.torch
        i = m._dev_i32(get(m.node_id)).long() if k != "node_neighbor_embedding" and k != "bias" else None
        j = m._dev_i32(get(m.node_neighbor_id)).long() if k != "node_embedding" else None
        if k == "node_embedding":
            return m.emb[i, :m.n_emb].cpu().numpy()
        if k == "node_neighbor_embedding":
            return m.emb[j, :m.n_emb].cpu().numpy()
        if k == "bias":
            return m.bias_t[j].cpu().numpy()
        s = (m.emb[i] * m.emb[j]).sum(1) + m.bias_t[j]
        if k == "score":
            return s.cpu().numpy()
        if k == "score_clipped":
            return s.clamp(-10, 10).cpu().numpy()
        if k == "prob":
            return torch.sigmoid(s).clamp(1e-5, 1).cpu().numpy()
        if k == "loss":
            lam = float(m.lam)
            l2 = 0.5 * ((m.emb[j] ** 2).sum() + (m.emb[i] ** 2).sum())
            if m._step_mode == 0:
                y = m._dev_f32(get(m.label))
                xent = torch.nn.functional.binary_cross_entropy_with_logits(s, y, reduction="sum")
                return float(xent + lam * (l2 + 0.5 * (m.bias_t[j] ** 2).sum()))
            r = m._dev_f32(get(m.reward))
            return float(-(torch.log(torch.sigmoid(s).clamp(1e-5, 1)) * r).mean() + lam * l2)
        raise KeyError("unknown fetch %r" % (fetch,))

    def close(self):
        pass


def _missing(ph):
    raise ValueError("You must feed a value for placeholder %r" % (ph,))
