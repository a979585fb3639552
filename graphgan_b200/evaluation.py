"""Link-prediction sanity check (reference src/evaluation/link_prediction.py:10-38): score the test
positives and negatives by embedding dot product, threshold at the median, report accuracy.
Out of the accelerated scope (tiny, CPU); kept so the trainer logs the same quality line."""
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
