"""One process per GPU (torch.distributed / NCCL over NVLink): how the hot path shards.

SURVEY.md section 8e.  Sampling shards by ROOTS and needs no collective: every rank holds the replicated
graph and embeddings, takes a contiguous block of the root list, and -- because the walk RNG is Philox keyed
by (root, walk, step) -- produces exactly the rows a single GPU would produce for those roots.  Rows are
all-gathered so that every rank sees the same training set (the reference's lists, in root order).

Updates are data parallel: every rank scores its slice of the 64-pair batch (K2), ONE collective per step
exchanges the compact gradients (ids + summed rows; a few KB, latency bound, NVLS-friendly), every rank merges
them in the same rank-major order (gg_grad_merge) and applies the same K3 Adam sweep, so replicas stay
bit-identical without broadcasting parameters.
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import ptr


def block_range(n, rank, world):
    """Contiguous block of [0, n) owned by `rank` (sizes differ by at most one, rank order = item order)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_root_ranges(weights, world):
    """Contiguous split of the root list into `world` blocks of roughly equal total weight (work of a root
    ~ its degree times the candidate lists it meets); keeps root order, so the concatenation of the ranks'
    rows equals the single-GPU row order."""
    w = np.asarray(weights, np.float64)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [int(np.searchsorted(cum, cum[-1] * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, len(w)
    return [(min(cuts[r], cuts[r + 1]), cuts[r + 1]) for r in range(world)]


def all_gather_varlen(t, group=None):
    """Concatenate 1-D tensors of different lengths from every rank, in rank order (works on gloo and nccl)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes + [1])
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:k] for o, k in zip(outs, sizes)])


class DataParallelStep:
    """Data-parallel replacement for PairModel.step: same arguments (the WHOLE mini-batch, identical on
    every rank), one collective per step."""

    def __init__(self, model, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.model, self.group = torch, dist, model, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.lib = _cabi.lib()
        self._cap = None

    def _buffers(self, cap):
        torch, m = self.torch, self.model
        if self._cap != cap:
            nf = int(self.lib.gg_grad_buf_floats(cap, m.ld))
            self.local = torch.zeros(nf, dtype=torch.float32, device=m.device)
            self.gathered = torch.empty(self.world * nf, dtype=torch.float32, device=m.device)
            self.slot_tmp = None
            self._cap, self._nf = cap, nf
        return self.local, self.gathered

    def step(self, node_id, node_neighbor_id, aux):
        m = self.model
        i, j, a = m._dev_i32(node_id), m._dev_i32(node_neighbor_id), m._dev_f32(aux)
        B = int(i.shape[0])
        if B == 0:
            return
        lo, hi = block_range(B, self.rank, self.world)
        cap = 2 * (-(-B // self.world))
        local, gathered = self._buffers(cap)
        ld, st = m.ld, m._stream()
        base = local.data_ptr()
        rows_p, bias_p, ids_p, nu_p = base, base + 4 * cap * ld, base + 4 * (cap * ld + cap), base + 4 * (cap * ld + 2 * cap)
        if hi > lo:
            _cabi.check(self.lib.gg_pair_grad(m._step_mode, hi - lo, B, ptr(i[lo:hi]), ptr(j[lo:hi]), ptr(a[lo:hi]), ptr(m.emb),
                                              ptr(m.bias_t), ld, C.c_float(float(m.lam)), nu_p, ids_p, rows_p, bias_p,
                                              ptr(m.row_slot), st), "gg_pair_grad")
        else:
            local[cap * ld + 2 * cap:].zero_()     # n_unique = 0
        self.dist.all_gather_into_tensor(gathered, local, group=self.group)   # the step's only collective
        _cabi.check(self.lib.gg_grad_merge(self.world, cap, ld, ptr(gathered), ptr(m.n_unique), ptr(m.uniq_ids),
                                           ptr(m.grad_rows), ptr(m.grad_bias), ptr(m.row_slot), st), "gg_grad_merge")
        m.apply_adam()
