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
    """``N\\td`` header then ``id\\tv0\\tv1...`` per node (graph_gan.py:299-306); values are the shortest
    round-trip repr of the fp32 value widened to float64, exactly what the reference's str() writes."""
    m = np.asarray(matrix)
    n, d = m.shape
    with open(filename, "w+") as f:
        f.write("%d\t%d\n" % (n, d))
        for lo in range(0, n, 4096):
            rows = m[lo:lo + 4096].tolist()
            f.write("".join("%d\t%s\n" % (lo + k, "\t".join(map(repr, r))) for k, r in enumerate(rows)))


def write_embeddings_binary(filename, model):
    """Binary dump of a device-resident model: ``filename`` gets a 16-byte header (magic b"GGE1", int32 N, int32 d,
    int32 0) followed by N*d little-endian fp32 values, row-major -- the same numbers the text file holds, 4 bytes
    each instead of ~12 characters.  The padded rows are compacted on the device (gg_unpad_rows) and cross PCIe once."""
    import torch
    from . import _cabi
    from ._cabi import ptr
    n, d = model.n_node, model.n_emb
    dense = torch.empty((n, d), dtype=torch.float32, device=model.device)
    _cabi.check(_cabi.lib().gg_unpad_rows(n, model.ld, d, ptr(model.emb), ptr(dense), model._stream()), "gg_unpad_rows")
    host = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
    host.copy_(dense, non_blocking=True)
    torch.cuda.current_stream(model.device).synchronize()
    with open(filename, "wb") as f:
        f.write(b"GGE1" + np.asarray([n, d, 0], np.int32).tobytes())
        host.numpy().tofile(f)


def read_embeddings_binary(filename):
    with open(filename, "rb") as f:
        head = f.read(16)
        if head[:4] != b"GGE1":
            raise ValueError("%s is not a GGE1 embedding dump" % filename)
        n, d, _ = np.frombuffer(head[4:], np.int32)
        return np.fromfile(f, np.float32, int(n) * int(d)).reshape(int(n), int(d))
