"""``Generator`` with the reference's constructor and attribute surface (src/GraphGAN/generator.py:5-31).

Attributes that are TF tensors/ops in the reference are opaque ``Fetch`` handles here, evaluated by
``session.Session.run(fetch, feed_dict)``; the parameters live in HBM (model.PairModel).  The new
trainer calls the batched methods (``step``, ``all_score_matrix``) directly.
"""
from . import config
from .model import Fetch, PairModel, Placeholder


class Generator(PairModel):
    _step_mode = 1  # generator loss (generator.py:26-29)

    def __init__(self, n_node, node_emd_init, device=None):
        super().__init__(n_node, node_emd_init, lr=config.lr_gen, lam=config.lambda_gen, device=device)
        # generator.py:11-15
        self.embedding_matrix = Fetch(self, "embedding_matrix")
        self.bias_vector = Fetch(self, "bias_vector")
        # generator.py:17-19
        self.node_id = Placeholder(self, "node_id")
        self.node_neighbor_id = Placeholder(self, "node_neighbor_id")
        self.reward = Placeholder(self, "reward")
        # generator.py:21-26
        self.all_score = Fetch(self, "all_score")
        self.node_embedding = Fetch(self, "node_embedding")
        self.node_neighbor_embedding = Fetch(self, "node_neighbor_embedding")
        self.bias = Fetch(self, "bias")
        self.score = Fetch(self, "score")
        self.prob = Fetch(self, "prob")
        # generator.py:28-31
        self.loss = Fetch(self, "loss")
        self.g_updates = Fetch(self, "g_updates")

    def g_step(self, node_id, node_neighbor_id, reward):
        """sess.run(generator.g_updates, {node_id, node_neighbor_id, reward}) (graph_gan.py:173-176)."""
        self.step(node_id, node_neighbor_id, reward)
