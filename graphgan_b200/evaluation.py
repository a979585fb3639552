"""Link-prediction sanity check (reference src/evaluation/link_prediction.py:10-38): score the test
positives and negatives by embedding dot product, threshold at the median, report accuracy.
``LinkPredictEval`` is the file-based form (text round trip, CPU) kept for compatibility; ``DeviceLinkPredictEval``
computes the same number from the device-resident embeddings (csrc/eval.cu) without the round trip."""
import numpy as np

from . import graph as G
from . import io


class LinkPredictEval:
    def __init__(self, embed_filename, test_filename, test_neg_filename, n_node, n_embed):
        self.test_filename, self.test_neg_filename = test_filename, test_neg_filename
        self.emd = io.read_embeddings(embed_filename, n_node=n_node, n_embed=n_embed)

    def eval_link_prediction(self):
        pos = G.read_edge_file(self.test_filename)
        neg = G.read_edge_file(self.test_neg_filename)
        edges = np.concatenate([pos, neg])
        score = np.einsum("ij,ij->i", self.emd[edges[:, 0]], self.emd[edges[:, 1]])
        pred = (score >= np.median(score)).astype(np.float64)
        truth = np.zeros(len(edges))
        truth[:len(edges) // 2] = 1
        return float(np.mean(pred == truth))


class DeviceLinkPredictEval:
    """Same quality line from a device-resident model: float64 row dots of the test positives then negatives
    (gg_pair_dot_f64), np.median threshold and accuracy on the device (gg_link_pred_acc).  The text round trip of an
    fp32 value is exact, so the operands equal those the file-based evaluation reads back."""

    _edge_cache = {}

    def __init__(self, model, test_filename, test_neg_filename):
        import torch
        self.model = model
        key = (test_filename, test_neg_filename, str(model.device))
        if key not in self._edge_cache:
            edges = np.concatenate([G.read_edge_file(test_filename), G.read_edge_file(test_neg_filename)])
            t = torch.as_tensor(np.ascontiguousarray(edges.astype(np.int32))).to(model.device)
            self._edge_cache[key] = (t[:, 0].contiguous(), t[:, 1].contiguous())
        self.a, self.b = self._edge_cache[key]

    def eval_link_prediction(self):
        import torch
        from . import _cabi
        from ._cabi import ptr
        m, lib = self.model, _cabi.lib()
        n = int(self.a.shape[0])
        score = torch.empty(n, dtype=torch.float64, device=m.device)
        out = torch.zeros(2, dtype=torch.float64, device=m.device)
        _cabi.check(lib.gg_pair_dot_f64(n, ptr(self.a), ptr(self.b), ptr(m.emb), m.ld, ptr(score), m._stream()), "gg_pair_dot_f64")
        _cabi.check(lib.gg_link_pred_acc(n, ptr(score), ptr(out), m._stream()), "gg_link_pred_acc")
        return float(out[0].item())
