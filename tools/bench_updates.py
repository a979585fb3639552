"""Micro-benchmark of the optimizer step (K2 + K3): C-driven launches vs the persistent cooperative loop.

usage: python tools/bench_updates.py [n_node n_emb batch n_steps]   (defaults: config C1 shape)
Prints us/step for gg_train_steps, gg_train_loop, and for K2 / K3 launched alone; checks both loops agree bit for bit.
"""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from graphgan_b200 import _cabi                      # noqa: E402
from graphgan_b200._cabi import ptr                  # noqa: E402
from graphgan_b200.generator import Generator        # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3


def main():
    n, n_emb, B, S = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (5242, 50, 64, 20000)))
    rng = np.random.default_rng(0)
    init = (rng.standard_normal((n, n_emb)) * 0.1).astype(np.float32)
    P = B * 4096
    i = torch.as_tensor(rng.integers(0, n, P).astype(np.int32)).cuda()
    j = torch.as_tensor(rng.integers(0, n, P).astype(np.int32)).cuda()
    r = torch.as_tensor(rng.random(P).astype(np.float32)).cuda()
    starts = (rng.integers(0, P // B, S) * B).astype(np.int64)
    res = {}
    import os
    variants = [(False, None, None), ("two-barrier", None, None), (True, None, None)]
    for persistent, nt, ct in variants:
        for k, val in (("GG_LOOP_THREADS", nt), ("GG_LOOP_CTAS", ct)):
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(val)
        g = Generator(n, init)
        g.train_steps(i, j, r, starts[:64], B, persistent=persistent)          # warm-up
        us = timed(lambda: g.train_steps(i, j, r, starts, B, persistent=persistent))
        res[persistent] = g.emb.clone()
        print("%-15s %8.2f us/step  (%d steps, n=%d ld=%d B=%d)" % ({False: "gg_train_steps", True: "gg_train_fused"}.get(persistent, "gg_train_loop"), us / S, S, n, g.ld, B))
        if persistent:
            c = g.sync_words.cpu().numpy()
            print("   CTA 0 cycles per step: gradient %.0f, sweep %.0f, wait for the other CTAs %.0f" % tuple(c[2:5] / max(c[5], 1)))
    os.environ.pop("GG_LOOP_THREADS", None); os.environ.pop("GG_LOOP_CTAS", None)
    print("bit-identical:", all(bool(torch.equal(res[False], v)) for v in res.values()))
    g = Generator(n, init)
    lib = g.lib
    st = g._stream()
    K = min(S, 5000)

    def k2():
        for s in range(K):
            o = int(starts[s])
            lib.gg_pair_grad(1, B, 0, i.data_ptr() + 4 * o, j.data_ptr() + 4 * o, r.data_ptr() + 4 * o, ptr(g.emb), ptr(g.bias_t),
                             g.ld, C.c_float(1e-5), ptr(g.n_unique), ptr(g.uniq_ids), ptr(g.grad_rows), ptr(g.grad_bias),
                             ptr(g.row_slot), st)

    def k3():
        for s in range(K):
            lib.gg_adam_apply(g.n_node, g.ld, ptr(g.emb), ptr(g.m_emb), ptr(g.v_emb), ptr(g.bias_t), ptr(g.m_bias), ptr(g.v_bias),
                              ptr(g.n_unique), ptr(g.uniq_ids), ptr(g.grad_rows), ptr(g.grad_bias), ptr(g.row_slot),
                              C.c_float(1e-3), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), st)
    k2(); t0 = time.perf_counter(); us2 = timed(k2); w2 = (time.perf_counter() - t0) * 1e6
    k3(); t0 = time.perf_counter(); us3 = timed(k3); w3 = (time.perf_counter() - t0) * 1e6
    print("K2 alone %.2f us/launch (host loop %.2f), K3 alone %.2f us/launch (host loop %.2f)" % (us2 / K, w2 / K, us3 / K, w3 / K))
    gb = 6.0 * n * g.ld * 4 / 1e9
    print("K3 sweep bytes %.4f GB -> %.0f GB/s at the stand-alone time" % (gb, gb / (us3 / K * 1e-6)))


if __name__ == "__main__":
    main()
