"""Embedding text I/O in the reference's formats (out of the accelerated path; kept so the
re-hosted trainer reads/writes the same files).

  read_embeddings   -- src/utils.py:57-67
  write_embeddings  -- src/GraphGAN/graph_gan.py:293-306
"""
import numpy as np


def read_embeddings(filename, n_node, n_embed):
    """First line is a header; every other line ``id v0 v1 ...``.  Rows absent from the file keep
    ``np.random.rand`` values (utils.py:63), drawn from the global numpy RNG like the reference."""
    emb = np.random.rand(n_node, n_embed)
    with open(filename, "r") as f:
        f.readline()
        for line in f:
            parts = line.split()
            if parts:
                emb[int(parts[0]), :] = [float(x) for x in parts[1:]]
    return emb


def write_embeddings(filename, matrix):
    """``N\\td`` header then ``id\\tv0\\tv1...`` per node (graph_gan.py:299-306)."""
    m = np.asarray(matrix)
    n, d = m.shape
    with open(filename, "w+") as f:
        f.write("%d\t%d\n" % (n, d))
        for i in range(n):
            f.write(str(i) + "\t" + "\t".join(str(x) for x in m[i].tolist()) + "\n")
