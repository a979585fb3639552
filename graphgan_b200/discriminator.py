"""``Discriminator`` with the reference's constructor and attribute surface
(src/GraphGAN/discriminator.py:5-34).  See generator.py for the conventions."""
from . import config
from .model import Fetch, PairModel, Placeholder


class Discriminator(PairModel):
    _step_mode = 0  # discriminator loss (discriminator.py:26-30)

    def __init__(self, n_node, node_emd_init, device=None):
        super().__init__(n_node, node_emd_init, lr=config.lr_dis, lam=config.lambda_dis, device=device)
        # discriminator.py:11-15
        self.embedding_matrix = Fetch(self, "embedding_matrix")
        self.bias_vector = Fetch(self, "bias_vector")
        # discriminator.py:17-19
        self.node_id = Placeholder(self, "node_id")
        self.node_neighbor_id = Placeholder(self, "node_neighbor_id")
        self.label = Placeholder(self, "label")
        # discriminator.py:21-24
        self.node_embedding = Fetch(self, "node_embedding")
        self.node_neighbor_embedding = Fetch(self, "node_neighbor_embedding")
        self.bias = Fetch(self, "bias")
        # discriminator.py:26-34.  As in the reference, ``score`` is re-bound to the CLIPPED score after
        # the loss has been built (discriminator.py:33), so fetching ``score`` yields the clipped value.
        self.loss = Fetch(self, "loss")
        self.d_updates = Fetch(self, "d_updates")
        self.score = Fetch(self, "score_clipped")
        self.reward = Fetch(self, "reward")

    def d_step(self, node_id, node_neighbor_id, label):
        """sess.run(discriminator.d_updates, {node_id, node_neighbor_id, label}) (graph_gan.py:154-157)."""
        self.step(node_id, node_neighbor_id, label)
