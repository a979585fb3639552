"""Host driver of K1, the graph-softmax walk (csrc/walk.cu), and of the tree builder.

Batched replacement for ``GraphGAN.sample`` (reference src/GraphGAN/graph_gan.py:225-270):
instead of one Python call per root it runs every walk of a batch of roots in one kernel
launch.  Torch tensors are used purely as device-memory containers; all computation happens
in libgraphgan_b200.so through the C ABI (include/graphgan_b200.h).
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import ptr

RNG_PHILOX, RNG_STREAM = 0, 1
NOTRUN, DONE, VOID, SKIPPED = 0, 1, 2, 3
CNT = dict(steps=0, sum_l=1, accepted=2, ok_roots=3, path_overflow=4, raw_steps=5, raw_sum_l=6, stream_used=7)


def round_up(x, m):
    return (x + m - 1) // m * m


def pad_embedding(emb, device=None):
    """float64/32 [N, d] -> fp32 [N, ld] device tensor, ld = round_up(d, 32), zero padded
    (the tf fp32 variable of generator.py:11-14 in the HBM layout of DESIGN.md section 2)."""
    import torch
    e = emb.float() if isinstance(emb, torch.Tensor) else torch.as_tensor(np.asarray(emb, np.float64).astype(np.float32))
    n, d = e.shape
    ld = round_up(d, 32)
    out = torch.zeros((n, ld), dtype=torch.float32, device=device if device is not None else e.device)
    out[:, :d] = e.to(out.device)
    return out


class TreeBatch:
    """BFS trees of a batch of roots as parent arrays: ``trees[root]`` of graph_gan.py:84-108."""

    def __init__(self, roots, parent):
        self.roots = roots      # device int32 [R]
        self.parent = parent    # device int32 [R, N]


class WalkOutput:
    """Device-side results of one pass (everything stays on the GPU until asked for)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def counters_host(self):
        c = self.counters.cpu().numpy().astype(np.uint64)
        return {k: int(c[i]) for k, i in CNT.items()}


class WalkSampler:
    def __init__(self, graph):
        import torch
        self.torch = torch
        self.g = graph
        self.device = graph.device
        self.lib = _cabi.lib()
        self.max_cand = graph.max_deg + 1
        nbytes = C.c_int64(0)
        _cabi.check(self.lib.gg_walk_scratch_bytes(self.max_cand, C.byref(nbytes)), "gg_walk_scratch_bytes")
        self.scratch = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=self.device)
        self.work_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._bfs_scratch = None

    # ------------------------------------------------------------------ trees
    def build_trees(self, roots):
        """construct_trees (graph_gan.py:84-108) for ``roots`` on the GPU -> TreeBatch."""
        torch = self.torch
        roots_d = roots if isinstance(roots, torch.Tensor) else torch.as_tensor(np.asarray(roots, np.int32)).to(self.device)
        R, N = int(roots_d.shape[0]), self.g.n_node
        parent = torch.empty((R, N), dtype=torch.int32, device=self.device)
        if self._bfs_scratch is None:
            nbytes = C.c_int64(0)
            _cabi.check(self.lib.gg_bfs_scratch_bytes(N, C.byref(nbytes)), "gg_bfs_scratch_bytes")
            self._bfs_scratch = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _cabi.check(self.lib.gg_bfs_build(N, ptr(self.g.indptr), ptr(self.g.adj), R, ptr(roots_d), ptr(parent),
                                          ptr(self._bfs_scratch), self._bfs_scratch.numel(), st), "gg_bfs_build")
        return TreeBatch(roots_d, parent)

    # ------------------------------------------------------------------ K1
    def run(self, emb, bias, trees, sample_num, for_d, *, seed=0, pass_tag=0, update_ratio=1.0, max_path=0,
            rng_mode=RNG_PHILOX, stream=None, finalize=True):
        """All walks of one pass.  ``sample_num``: int (G mode, config.n_sample_gen) or device
        int64 [R] (D mode, len(graph[root]))."""
        torch = self.torch
        dev = self.device
        R = int(trees.roots.shape[0])
        if isinstance(sample_num, int):
            walk_ptr = torch.arange(0, (R + 1) * sample_num, max(sample_num, 1), dtype=torch.int64, device=dev)[:R + 1] \
                if sample_num > 0 else torch.zeros(R + 1, dtype=torch.int64, device=dev)
            W = R * sample_num
        else:
            walk_ptr = torch.zeros(R + 1, dtype=torch.int64, device=dev)
            walk_ptr[1:] = torch.cumsum(sample_num.to(torch.int64), 0)   # plumbing: prefix of sample_num
            W = int(walk_ptr[-1].item())
        i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
        out = WalkOutput(
            walk_ptr=walk_ptr, n_walks=W, n_roots=R, for_d=bool(for_d), max_path=max_path,
            samples=i32(max(W, 1)), status=i32(max(W, 1)), first_edge=i32(max(W, 1)), wsteps=i32(max(W, 1)),
            wsuml=i32(max(W, 1)), paths=i32(max(W, 1), max_path) if max_path > 0 else None,
            path_len=i32(max(W, 1)) if max_path > 0 else None, root_ok=torch.zeros(max(R, 1), dtype=torch.int32, device=dev),
            counters=torch.zeros(8, dtype=torch.int64, device=dev), roots=trees.roots)
        d = _cabi.WalkDesc()
        d.n_node, d.ld = self.g.n_node, int(emb.shape[1])
        d.emb, d.bias, d.indptr, d.adj = ptr(emb), ptr(bias), ptr(self.g.indptr), ptr(self.g.adj)
        d.n_roots, d.roots, d.parent, d.walk_ptr, d.n_walks = R, ptr(trees.roots), ptr(trees.parent), ptr(walk_ptr), W
        d.for_d, d.rng_mode, d.d1_bits = int(bool(for_d)), rng_mode, ptr(self.g.d1_bits)
        d.seed, d.pass_tag, d.max_path = seed, pass_tag, max_path
        d.stream, d.n_stream = (ptr(stream), int(stream.numel())) if stream is not None else (None, 0)
        d.update_ratio, d.max_cand = float(update_ratio), self.max_cand
        d.samples, d.status, d.first_edge, d.wsteps, d.wsuml = (ptr(out.samples), ptr(out.status), ptr(out.first_edge),
                                                                 ptr(out.wsteps), ptr(out.wsuml))
        d.paths, d.path_len, d.counters = ptr(out.paths), ptr(out.path_len), ptr(out.counters)
        d.scratch, d.scratch_bytes, d.work_counter = ptr(self.scratch), self.scratch.numel(), ptr(self.work_counter)
        assert emb.dtype == torch.float32 and emb.is_contiguous() and bias.dtype == torch.float32
        st = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(self.lib.gg_walk_sample(C.byref(d), st), "gg_walk_sample")
        if finalize:
            self.finalize(out)
        return out

    def finalize(self, out):
        st = self.torch.cuda.current_stream(self.device).cuda_stream
        _cabi.check(self.lib.gg_walk_finalize(out.n_roots, ptr(out.walk_ptr), int(out.for_d), ptr(out.samples),
                                              ptr(out.status), ptr(out.first_edge), ptr(out.wsteps), ptr(out.wsuml),
                                              ptr(out.path_len), ptr(self.g.d1_bits), ptr(out.root_ok),
                                              ptr(out.counters), st), "gg_walk_finalize")

    def emit_d_rows(self, out):
        """prepare_data_for_d's (center, neighbor, label) rows (graph_gan.py:192-201), on device."""
        torch = self.torch
        dev = self.device
        cap = 2 * out.n_walks
        row_ptr = torch.empty(out.n_roots + 1, dtype=torch.int64, device=dev)
        n_rows = torch.zeros(1, dtype=torch.int64, device=dev)
        center, neighbor, label = (torch.empty(max(cap, 1), dtype=torch.int32, device=dev) for _ in range(3))
        st = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(self.lib.gg_emit_d_rows(out.n_roots, ptr(out.roots), ptr(out.walk_ptr), ptr(self.g.raw_indptr),
                                            ptr(self.g.raw_adj), ptr(out.root_ok), ptr(out.samples), ptr(row_ptr),
                                            ptr(center), ptr(neighbor), ptr(label), ptr(n_rows), st), "gg_emit_d_rows")
        return center, neighbor, label, n_rows
