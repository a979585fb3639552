"""Host driver of K1, the graph-softmax walk (csrc/walk.cu), and of the tree builder.

Batched replacement for ``GraphGAN.sample`` (reference src/GraphGAN/graph_gan.py:225-270):
instead of one Python call per root it runs every walk of a batch of roots in one kernel
launch.  Torch tensors are used purely as device-memory containers; all computation happens
in libgraphgan_b200.so through the C ABI (include/graphgan_b200.h).
"""
import ctypes as C
import os

import numpy as np

from . import _cabi
from ._cabi import ptr

RNG_PHILOX, RNG_STREAM = 0, 1
NOTRUN, DONE, VOID, SKIPPED = 0, 1, 2, 3
CNT = dict(steps=0, sum_l=1, accepted=2, ok_roots=3, path_overflow=4, raw_steps=5, raw_sum_l=6, stream_used=7,
           rows_gathered=8, cyc_enum=9, cyc_score=10, cyc_choose=11, cyc_step0=12, cyc_step1=13, cyc_step2p=14, cyc_walk=15)


def round_up(x, m):
    return (x + m - 1) // m * m


def pad_embedding(emb, device=None):
    """float64/32 [N, d] -> fp32 [N, ld] device tensor, ld = 32 * 2^k >= d (32, 64, 128 or 256), zero padded
    (the tf fp32 variable of generator.py:11-14 in the HBM layout of DESIGN.md section 2)."""
    import torch
    e = emb.float() if isinstance(emb, torch.Tensor) else torch.as_tensor(np.asarray(emb, np.float64).astype(np.float32))
    n, d = e.shape
    if d > 256:
        raise ValueError("n_emb = %d is not supported (the kernels are instantiated for row strides 32, 64, 128, 256)" % d)
    ld = 32
    while ld < d:      # zero columns add exactly +0 to every canonical dot, so the amount of padding is invisible
        ld *= 2
    out = torch.zeros((n, ld), dtype=torch.float32, device=device if device is not None else e.device)
    out[:, :d] = e.to(out.device)
    return out


class TreeBatch:
    """BFS trees of a batch of roots (``trees[root]`` of graph_gan.py:84-108) as one bit per walk-CSR entry:
    bit e of row k <=> adj[e] is a child of entry e's source node in the tree of roots[k] (csrc/bfs.cu)."""

    def __init__(self, roots, tree_bits, graph=None):
        self.roots = roots            # device int32 [R]
        self.tree_bits = tree_bits    # device int32 [R, tree_words] (bit patterns)
        self.graph = graph

    def slice(self, lo, hi):
        return TreeBatch(self.roots[lo:hi].contiguous(), self.tree_bits[lo:hi].contiguous(), self.graph)

    def select(self, idx):
        return TreeBatch(self.roots[idx].contiguous(), self.tree_bits[idx].contiguous(), self.graph)

    def parent_arrays(self, rows=None):
        """int32 [R', N] parent arrays (root and unreachable nodes: -1) of all / the selected rows -- the form the
        oracle and host-side consumers use; expanded on the device by gg_tree_parent."""
        import torch
        g, lib = self.graph, _cabi.lib()
        bits = self.tree_bits if rows is None else self.tree_bits[rows].contiguous()
        roots = self.roots if rows is None else self.roots[rows].contiguous()
        R = int(bits.shape[0])
        out = torch.empty((R, g.n_node), dtype=torch.int32, device=bits.device)
        st = torch.cuda.current_stream(bits.device).cuda_stream
        for lo in range(0, R, 32768):
            hi = min(R, lo + 32768)
            _cabi.check(lib.gg_tree_parent(g.n_node, ptr(g.indptr), ptr(g.adj), hi - lo, ptr(roots[lo:hi]), ptr(bits[lo:hi]),
                                           int(bits.shape[1]), ptr(out[lo:hi]), st), "gg_tree_parent")
        return out


class WalkOutput:
    """Device-side results of one pass (everything stays on the GPU until asked for)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def counters_host(self):
        c = self.counters.cpu().numpy().astype(np.uint64)
        return {k: int(c[i]) for k, i in CNT.items()}


class WalkPlan:
    """Everything about a pass that does not depend on the embeddings: the walk list of a root batch
    (walk_ptr), the root-CDF offsets, and the output buffers (reused by every run with this plan)."""

    def __init__(self, sampler, trees, sample_num, for_d, max_path):
        import torch
        dev, g = sampler.device, sampler.g
        R = int(trees.roots.shape[0])
        self.trees, self.for_d, self.max_path, self.n_roots = trees, bool(for_d), int(max_path), R
        if isinstance(sample_num, int):
            self.walk_ptr = torch.arange(R + 1, dtype=torch.int64, device=dev) * sample_num
            self.n_walks = R * sample_num
        else:
            self.walk_ptr = torch.zeros(R + 1, dtype=torch.int64, device=dev)
            self.walk_ptr[1:] = torch.cumsum(sample_num.to(torch.int64), 0)      # plumbing: prefix of sample_num
            self.n_walks = int(self.walk_ptr[-1].item())
        nw = self.walk_ptr[1:] - self.walk_ptr[:-1]
        self.walk_slot = torch.repeat_interleave(torch.arange(R, dtype=torch.int32, device=dev), nw,
                                                 output_size=self.n_walks) if self.n_walks else None
        r = trees.roots.long()
        self.rq_ptr = torch.zeros(R + 1, dtype=torch.int64, device=dev)
        self.rq_ptr[1:] = torch.cumsum(g.indptr[r + 1] - g.indptr[r], 0)       # prefix of the roots' walk degrees
        self.nq = int(self.rq_ptr[-1].item())
        nq = max(self.nq, 1)
        self.root_q = torch.empty(nq, dtype=torch.float64, device=dev)
        self.root_sc = torch.empty(nq, dtype=torch.float32, device=dev)
        self._s1 = None
        self._order = None
        W = max(self.n_walks, 1)
        i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
        self.samples, self.status, self.first_edge, self.wsteps, self.wsuml = i32(W), i32(W), i32(W), i32(W), i32(W)
        self.paths = i32(W, max_path) if max_path > 0 else None
        self.path_len = i32(W) if max_path > 0 else None
        self.root_ok = torch.zeros(max(R, 1), dtype=torch.int32, device=dev)
        self.counters = torch.zeros(16, dtype=torch.int64, device=dev)
        self.row_ptr = torch.empty(R + 1, dtype=torch.int64, device=dev)
        self.n_rows = torch.zeros(1, dtype=torch.int64, device=dev)
        self.rows = [i32(max(2 * self.n_walks, 1)) for _ in range(3)] if for_d else None


    def start_order(self, sampler):
        """Order in which walk_kernel starts the walks: roots whose neighbourhood holds the largest hub first
        (a walk that steps onto a 10 k-neighbour node costs ~100x a median one; started last it would be the
        tail of the launch), walks of one root kept together (they share the tree row in L2).  Static per
        plan; plumbing only (one gather, one segment max, one sort)."""
        if self._order is None and self.n_walks > 0 and self.nq > 0:
            torch, g, dev = sampler.torch, sampler.g, sampler.device
            R = self.n_roots
            deg_r = self.rq_ptr[1:] - self.rq_ptr[:-1]
            slot = torch.repeat_interleave(torch.arange(R, dtype=torch.int64, device=dev), deg_r, output_size=self.nq)
            ent = g.indptr[self.trees.roots.long()[slot]] + (torch.arange(self.nq, dtype=torch.int64, device=dev) - self.rq_ptr[slot])
            c = g.adj[ent].long()
            key = torch.zeros(R, dtype=torch.int64, device=dev).scatter_reduce_(0, slot, g.indptr[c + 1] - g.indptr[c], "amax")
            perm = torch.argsort(key, descending=True, stable=True)
            nw = (self.walk_ptr[1:] - self.walk_ptr[:-1])[perm]
            first = torch.cumsum(nw, 0) - nw                                  # position of each root's first walk in the order
            base = torch.repeat_interleave(self.walk_ptr[:-1][perm] - first, nw, output_size=self.n_walks)
            self._order = (base + torch.arange(self.n_walks, dtype=torch.int64, device=dev)).to(torch.int32)
        return self._order

    def flat_buffer(self, sampler):
        """scratch of the level-synchronous steps (csrc/walk.cu: flat_*_kernel), sized by the library"""
        key = (sampler.flat_steps, sampler.hub_threshold)
        if getattr(self, "_flat_key", None) != key:
            nbytes = C.c_int64(0)
            _cabi.check(sampler.lib.gg_walk_flat_bytes(self.n_walks, sampler.hub_threshold, sampler.flat_steps, C.byref(nbytes)),
                        "gg_walk_flat_bytes")
            self._flat = sampler.torch.empty(max(nbytes.value, 16), dtype=sampler.torch.uint8, device=sampler.device)
            self._flat_key = key
        return self._flat

    def depth1_buffers(self, sampler):
        """Static layout of the depth-1 CDF cache (csrc/walk.cu: step1_cdf_kernel): one slice of degree(child) + 1
        entries per (root, neighbour) pair.  Built on first use (plumbing: gathers + one cumsum)."""
        if self._s1 is None:
            torch, g, dev = sampler.torch, sampler.g, sampler.device
            deg_r = self.rq_ptr[1:] - self.rq_ptr[:-1]
            slot = torch.repeat_interleave(torch.arange(self.n_roots, dtype=torch.int32, device=dev), deg_r, output_size=self.nq)
            r = self.trees.roots.long()[slot.long()]
            ent = g.indptr[r] + (torch.arange(self.nq, dtype=torch.int64, device=dev) - self.rq_ptr[slot.long()])
            c = g.adj[ent].long()
            ptr_ = torch.zeros(self.nq + 1, dtype=torch.int64, device=dev)
            ptr_[1:] = torch.cumsum(g.indptr[c + 1] - g.indptr[c] + 1, 0)
            total = max(int(ptr_[-1].item()), 1)
            i32 = lambda k: torch.empty(max(k, 1), dtype=torch.int32, device=dev)
            order = torch.argsort(g.indptr[c + 1] - g.indptr[c], descending=True, stable=True).to(torch.int32)
            self._s1 = dict(slot=slot, ptr=ptr_, cnt=i32(self.nq), n=i32(self.nq), ids=i32(total), order=order,
                            q=torch.empty(total, dtype=torch.float64, device=dev), first=i32(self.n_walks))
        return self._s1


class WalkSampler:
    def __init__(self, graph, hub_threshold=128, depth1=True, hub_first=True, tma=True):
        import torch
        self.torch = torch
        self.g = graph
        self.device = graph.device
        self.lib = _cabi.lib()
        self.max_cand = graph.max_deg + 1
        self.hub_threshold = int(hub_threshold)   # 0 disables both per-pass reuses (pure on-demand path)
        # one CDF per (root, depth-1 child) pair that occurs (needs reuse): the walks of a root that pick the same child
        # share its candidate list (5.5x fewer neighbour probes at step 1 on C3); the builder kernel pulls the pairs
        # from a queue, largest lists first (hub_first), because a 13.8k-entry hub list occupies one warp for ~0.5-1 ms
        self.depth1 = bool(depth1)
        self.tma = bool(tma)                      # cp.async.bulk staging of hub lists (csrc/walk.cu); False = plain loads (A/B)
        self.hub_first = bool(hub_first)          # start order of the walks (WalkPlan.start_order); results do not depend on it
        nbytes = C.c_int64(0)
        _cabi.check(self.lib.gg_walk_scratch_bytes(self.max_cand, C.byref(nbytes)), "gg_walk_scratch_bytes")
        self.scratch = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=self.device)
        self.work_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._bfs_scratch = None
        # tree builder: direction-optimising BFS (csrc/bfs.cu).  < 0: library default ratio, 0: top-down + small sorted
        # bottom-up levels only; the trees are identical in every mode (tests/test_walk_gpu.py)
        self.bfs_bottom_up_ratio = float(os.environ.get("GG_BFS_BU_RATIO", "-1"))
        self.bfs_flags = 0
        # level-synchronous walk steps (csrc/walk.cu: flat_*_kernel) for steps 1..flat_steps; 0 = persistent kernel only
        self.flat_steps = int(os.environ.get("GG_FLAT_STEPS", "4"))

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ trees
    def build_trees(self, roots):
        """construct_trees (graph_gan.py:84-108) for ``roots`` on the GPU -> TreeBatch."""
        torch = self.torch
        roots_d = roots if isinstance(roots, torch.Tensor) else torch.as_tensor(np.asarray(roots, np.int32)).to(self.device)
        R, N = int(roots_d.shape[0]), self.g.n_node
        nnz = int(self.g.adj.shape[0])
        words = C.c_int64(0)
        _cabi.check(self.lib.gg_tree_words(nnz, C.byref(words)), "gg_tree_words")
        tree_bits = torch.empty((R, words.value), dtype=torch.int32, device=self.device)
        if self._bfs_scratch is None:
            nbytes = C.c_int64(0)
            _cabi.check(self.lib.gg_bfs_scratch_bytes(N, int(self.g.adj.shape[0]), C.byref(nbytes)), "gg_bfs_scratch_bytes")
            self._bfs_scratch = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=self.device)
        rev = self.g.reverse_entries() if self.bfs_bottom_up_ratio != 0.0 else None
        _cabi.check(self.lib.gg_bfs_build_ex(N, nnz, ptr(self.g.indptr), ptr(self.g.adj), ptr(rev) if rev is not None else None,
                                             R, ptr(roots_d), ptr(tree_bits), words.value, ptr(self._bfs_scratch),
                                             self._bfs_scratch.numel(), float(self.bfs_bottom_up_ratio), int(self.bfs_flags),
                                             self._stream()),
                    "gg_bfs_build_ex")
        return TreeBatch(roots_d, tree_bits, self.g)

    # ------------------------------------------------------------------ K1
    def plan(self, trees, sample_num, for_d, max_path=0):
        return WalkPlan(self, trees, sample_num, for_d, max_path)

    def _desc(self, emb, bias, plan, *, seed, pass_tag, update_ratio, rng_mode, stream, reuse, phase_mask=0):
        torch = self.torch
        assert emb.dtype == torch.float32 and emb.is_contiguous() and bias.dtype == torch.float32
        t, d = plan.trees, _cabi.WalkDesc()
        d.n_node, d.ld = self.g.n_node, int(emb.shape[1])
        d.emb, d.bias, d.indptr, d.adj = ptr(emb), ptr(bias), ptr(self.g.indptr), ptr(self.g.adj)
        d.n_roots, d.roots, d.walk_ptr, d.n_walks = plan.n_roots, ptr(t.roots), ptr(plan.walk_ptr), plan.n_walks
        d.tree_bits, d.tree_words = ptr(t.tree_bits), int(t.tree_bits.shape[1])
        d.for_d, d.rng_mode, d.d1_bits = int(plan.for_d), rng_mode, ptr(self.g.d1_bits)
        d.seed, d.pass_tag, d.max_path = seed, pass_tag, plan.max_path
        d.stream, d.n_stream = (ptr(stream), int(stream.numel())) if stream is not None else (None, 0)
        d.update_ratio, d.max_cand, d.phase_mask = float(update_ratio), self.max_cand, int(phase_mask)
        d.no_tma = 0 if self.tma else 1
        d.samples, d.status, d.first_edge, d.wsteps, d.wsuml = (ptr(plan.samples), ptr(plan.status), ptr(plan.first_edge),
                                                                 ptr(plan.wsteps), ptr(plan.wsuml))
        d.paths, d.path_len, d.counters = ptr(plan.paths), ptr(plan.path_len), ptr(plan.counters)
        d.scratch, d.scratch_bytes, d.work_counter = ptr(self.scratch), self.scratch.numel(), ptr(self.work_counter)
        d.rq_ptr, d.walk_slot = ptr(plan.rq_ptr), ptr(plan.walk_slot)
        if self.hub_first and rng_mode == RNG_PHILOX:
            d.walk_order = ptr(plan.start_order(self))
        if reuse:
            self.g.hub_tiles(self.hub_threshold)
            d.edge_score, d.hub_threshold, d.root_q = ptr(self.g.edge_score), self.hub_threshold, ptr(plan.root_q)
            if self.depth1 and rng_mode == RNG_PHILOX and plan.nq > 0 and plan.n_walks > 0:
                b = plan.depth1_buffers(self)
                d.s1_nq, d.s1_slot, d.s1_ptr, d.s1_cnt, d.s1_n = plan.nq, ptr(b["slot"]), ptr(b["ptr"]), ptr(b["cnt"]), ptr(b["n"])
                d.s1_q, d.s1_ids, d.first_idx = ptr(b["q"]), ptr(b["ids"]), ptr(b["first"])
                if self.hub_first:
                    d.s1_order = ptr(b["order"])
                if self.flat_steps > 0:
                    buf = plan.flat_buffer(self)
                    d.flat_buf, d.flat_bytes, d.flat_steps = ptr(buf), buf.numel(), self.flat_steps
        return d

    def precompute(self, emb, bias, plan, desc=None):
        """Per-pass reuse (csrc/hub.cu): hub adjacency scores, then one CDF per root.  Depends on the
        embeddings, so it belongs to every pass; `run` calls it unless told otherwise."""
        d = desc if desc is not None else self._desc(emb, bias, plan, seed=0, pass_tag=0, update_ratio=1.0,
                                                     rng_mode=RNG_PHILOX, stream=None, reuse=True)
        tile_node, tile_begin, n_tiles, _ = self.g.hub_tiles(self.hub_threshold)
        st = self._stream()
        _cabi.check(self.lib.gg_hub_scores(n_tiles, ptr(tile_node), ptr(tile_begin), 256, ptr(self.g.indptr), ptr(self.g.adj),
                                           ptr(emb), ptr(bias), int(emb.shape[1]), ptr(self.g.edge_score), st), "gg_hub_scores")
        _cabi.check(self.lib.gg_root_cdf(C.byref(d), ptr(plan.root_sc), ptr(plan.root_q), st), "gg_root_cdf")

    def run(self, emb, bias, trees, sample_num, for_d, *, seed=0, pass_tag=0, update_ratio=1.0, max_path=0,
            rng_mode=RNG_PHILOX, stream=None, finalize=True, plan=None, reuse=None, precompute=True, phase_mask=0, zero_counters=True):
        """All walks of one pass.  ``sample_num``: int (G mode, config.n_sample_gen) or device int64 [R]
        (D mode, len(graph[root])).  ``reuse`` (default: hub_threshold > 0) turns the per-pass score / CDF
        reuse on; results are bit-identical either way."""
        if plan is None:
            plan = self.plan(trees, sample_num, for_d, max_path)
        reuse = (self.hub_threshold > 0) if reuse is None else bool(reuse)
        if zero_counters:
            plan.counters.zero_()
        d = self._desc(emb, bias, plan, seed=seed, pass_tag=pass_tag, update_ratio=update_ratio, rng_mode=rng_mode,
                       stream=stream, reuse=reuse, phase_mask=phase_mask)
        if reuse and precompute:
            self.precompute(emb, bias, plan, d)
        _cabi.check(self.lib.gg_walk_sample(C.byref(d), self._stream()), "gg_walk_sample")
        out = WalkOutput(walk_ptr=plan.walk_ptr, n_walks=plan.n_walks, n_roots=plan.n_roots, for_d=plan.for_d,
                         max_path=plan.max_path, samples=plan.samples, status=plan.status, first_edge=plan.first_edge,
                         wsteps=plan.wsteps, wsuml=plan.wsuml, paths=plan.paths, path_len=plan.path_len,
                         root_ok=plan.root_ok, counters=plan.counters, roots=trees.roots, plan=plan)
        if finalize:
            self.finalize(out)
        return out

    def finalize(self, out):
        _cabi.check(self.lib.gg_walk_finalize(out.n_roots, ptr(out.walk_ptr), int(out.for_d), ptr(out.samples),
                                              ptr(out.status), ptr(out.first_edge), ptr(out.wsteps), ptr(out.wsuml),
                                              ptr(out.path_len), ptr(self.g.d1_bits), ptr(out.root_ok),
                                              ptr(out.counters), self._stream()), "gg_walk_finalize")

    def emit_d_rows(self, out):
        """prepare_data_for_d's (center, neighbor, label) rows (graph_gan.py:192-201), on device."""
        p = out.plan
        center, neighbor, label = p.rows
        _cabi.check(self.lib.gg_emit_d_rows(out.n_roots, ptr(out.roots), ptr(out.walk_ptr), ptr(self.g.raw_indptr),
                                            ptr(self.g.raw_adj), ptr(out.root_ok), ptr(out.samples), ptr(p.row_ptr),
                                            ptr(center), ptr(neighbor), ptr(label), ptr(p.n_rows), self._stream()),
                    "gg_emit_d_rows")
        return center, neighbor, label, p.n_rows
