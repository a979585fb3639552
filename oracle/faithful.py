"""oracle/faithful.py -- TEST INFRASTRUCTURE ONLY (tier "T0", the reference-faithful oracle).

A numpy restatement of the reference's sampling path, keeping its control flow, data
structures (dict-of-lists BFS trees that are mutated in place), candidate order and
consumption of the legacy ``RandomState`` stream:

  * tree construction          -- src/GraphGAN/graph_gan.py:84-108
  * ``sample``                 -- src/GraphGAN/graph_gan.py:225-270
  * ``prepare_data_for_d``     -- src/GraphGAN/graph_gan.py:182-202
  * ``prepare_data_for_g``     -- src/GraphGAN/graph_gan.py:204-223
  * window pairs               -- src/GraphGAN/graph_gan.py:272-291
  * softmax                    -- src/utils.py:131-133
  * ``generator.all_score``    -- src/GraphGAN/generator.py:21   (numpy stands in for TF1.8)
  * ``discriminator.reward``   -- src/GraphGAN/discriminator.py:21-24, 33-34

Pinned by tests/test_oracle.py against tests/golden/*.npz, which were produced by running the
reference's own Python (tests/golden/make_golden.py).  TensorFlow itself cannot run here, so
the dense arithmetic (sgemm summation order, exp) is numpy's: "parity unpinned" at that seam.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs
may import this module.  The product (graphgan_b200/) never does.

Score modes
  literal : all_score = E.E^T + b recomputed on EVERY sample() call (graph_gan.py:238)
  cached  : the same matrix computed once per pass (identical values, used for fixtures)
  lazy    : all_score[cur, cand] computed on demand -- the only form that exists at N >= 100k
"""
import collections

import numpy as np


def softmax(x):  # utils.py:131-133
    e = np.exp(x - np.max(x))
    return e / e.sum()


def build_trees(graph, roots):
    """graph_gan.py:84-108: root -> {node -> [father, child_0, child_1, ...]}."""
    trees = {}
    for root in roots:
        t = {root: [root]}
        seen = {root}
        queue = collections.deque([root])
        while queue:
            cur = queue.popleft()
            for nb in graph[cur]:
                if nb not in seen:
                    seen.add(nb)
                    t[cur].append(nb)
                    t[nb] = [cur]
                    queue.append(nb)
        trees[root] = t
    return trees


def node_pairs_from_path(path, window_size=2):
    """graph_gan.py:272-291."""
    body = path[:-1]
    out = []
    for i, center in enumerate(body):
        lo, hi = max(i - window_size, 0), min(i + window_size + 1, len(body))
        for j in range(lo, hi):
            if j != i:
                out.append([center, body[j]])
    return out


class ParentTrees:
    """Lazy stand-in for the dict trees at scales where O(N^2) dicts cannot exist.

    trees[root][cur] is synthesised from a parent array + the unique-neighbour adjacency;
    the in-place father removal (graph_gan.py:258-259) is kept as a per-root set."""

    def __init__(self, adj_lists, parent_of_root):
        self.adj = adj_lists          # node -> np.ndarray of unique neighbours, file order
        self.parent_of_root = parent_of_root  # root -> np.ndarray parent[N]
        self.removed = collections.defaultdict(set)

    def tree(self, root):
        return _LazyTree(self, root)


class _LazyTree:
    def __init__(self, owner, root):
        self.o, self.root = owner, root
        self.par = owner.parent_of_root[root]
        self.cache = {}

    def __getitem__(self, cur):
        lst = self.cache.get(cur)
        if lst is None:
            nb = self.o.adj[cur]
            kids = nb[self.par[nb] == cur].tolist()
            if cur == self.root:
                lst = [cur] + kids
            elif cur in self.o.removed[self.root]:
                lst = kids
            else:
                lst = [int(self.par[cur])] + kids
            lst = _TrackedList(lst, self, cur)
            self.cache[cur] = lst
        return lst


class _TrackedList(list):
    def __init__(self, it, tree, cur):
        super().__init__(it)
        self._tree, self._cur = tree, cur

    def remove(self, x):
        super().remove(x)
        self._tree.o.removed[self._tree.root].add(self._cur)


class Faithful:
    def __init__(self, graph, emb_g, bias_g=None, emb_d=None, bias_d=None, rng=None, score_mode="cached",
                 trees=None):
        self.graph = graph                      # list / dict: node -> list of neighbours (raw, file order)
        self.n_node = len(graph)
        # tf fp32 variable (generator.py:10-14); an fp32 array is taken as is
        self.E = emb_g if getattr(emb_g, "dtype", None) == np.float32 else np.asarray(emb_g, np.float64).astype(np.float32)
        self.b = np.zeros(self.n_node, np.float32) if bias_g is None else np.asarray(bias_g, np.float32)
        if emb_d is not None:
            self.Ed = np.asarray(emb_d, np.float64).astype(np.float32)
            self.bd = np.zeros(self.n_node, np.float32) if bias_d is None else np.asarray(bias_d, np.float32)
        self.rng = rng if rng is not None else np.random.RandomState(0)
        self.score_mode = score_mode
        self.trees = trees
        self._all = None
        self.stats = collections.Counter()

    # -- the TF fetch stand-ins ------------------------------------------------------------
    def all_score(self):
        return (self.E @ self.E.T + self.b).astype(np.float32)      # generator.py:21

    def reward(self, node_1, node_2):                                # discriminator.py:21-24, 33-34
        i, j = np.asarray(node_1, np.int64), np.asarray(node_2, np.int64)
        if i.size == 0:
            return np.zeros(0, np.float32)
        s = np.sum(self.Ed[i] * self.Ed[j], axis=1, dtype=np.float32) + self.bd[j]
        s = np.clip(s, -10, 10).astype(np.float32)
        return np.log(np.float32(1) + np.exp(s)).astype(np.float32)

    def invalidate(self):
        self._all = None

    # -- graph_gan.py:225-270 --------------------------------------------------------------
    def sample(self, root, tree, sample_num, for_d):
        if self.score_mode == "literal":
            all_score = self.all_score()
        elif self.score_mode == "cached":
            if self._all is None:
                self._all = self.all_score()
            all_score = self._all
        else:
            all_score = None
        samples, paths = [], []
        n = 0
        while len(samples) < sample_num:
            cur, prev = root, -1
            paths.append([cur])
            at_root = True
            while True:
                cand = tree[cur][1:] if at_root else tree[cur]
                at_root = False
                if len(cand) == 0:
                    return None, None
                if for_d:
                    if cand == [root]:
                        return None, None
                    if root in cand:
                        cand.remove(root)          # in place: mutates the cached tree
                if all_score is not None:
                    rel = all_score[cur, cand]
                else:
                    c = np.asarray(cand, np.int64)
                    rel = (self.E[c] @ self.E[cur] + self.b[c]).astype(np.float32)
                prob = softmax(rel)
                nxt = self.rng.choice(cand, size=1, p=prob)[0]
                self.stats["steps"] += 1
                self.stats["sum_l"] += len(cand)
                paths[n].append(nxt)
                if nxt == prev:
                    samples.append(cur)
                    break
                prev, cur = cur, nxt
            n += 1
        return samples, paths

    def _tree(self, root):
        return self.trees.tree(root) if isinstance(self.trees, ParentTrees) else self.trees[root]

    # -- graph_gan.py:182-202 --------------------------------------------------------------
    def prepare_data_for_d(self, roots=None, update_ratio=1):
        center, neighbor, labels = [], [], []
        for i in (range(self.n_node) if roots is None else roots):
            if self.rng.rand() < update_ratio:
                pos = self.graph[i]
                neg, _ = self.sample(i, self._tree(i), len(pos), for_d=True)
                if len(pos) != 0 and neg is not None:
                    center.extend([i] * len(pos)); neighbor.extend(pos); labels.extend([1] * len(pos))
                    center.extend([i] * len(pos)); neighbor.extend(neg); labels.extend([0] * len(neg))
                    self.stats["neg_edges"] += len(neg)
        return center, neighbor, labels

    # -- graph_gan.py:204-223 --------------------------------------------------------------
    def prepare_data_for_g(self, roots=None, n_sample_gen=20, window_size=2, update_ratio=1, with_paths=False):
        paths = []
        for i in (range(self.n_node) if roots is None else roots):
            if self.rng.rand() < update_ratio:
                _, p = self.sample(i, self._tree(i), n_sample_gen, for_d=False)
                if p is not None:
                    paths.extend(p)
        node_1, node_2 = [], []
        for p in paths:
            for a, b in node_pairs_from_path(p, window_size):
                node_1.append(a); node_2.append(b)
        reward = self.reward(node_1, node_2)
        if with_paths:
            return node_1, node_2, reward, paths
        return node_1, node_2, reward
