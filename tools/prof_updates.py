"""Tiny driver for ncu captures of the step-loop kernels: python tools/prof_updates.py {fused|loop|steps} [n_steps [n_node n_emb]]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from graphgan_b200.generator import Generator        # noqa: E402

how = {"fused": True, "loop": "two-barrier", "steps": False}[sys.argv[1]]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
n, n_emb, B = (int(sys.argv[3]), int(sys.argv[4]), 64) if len(sys.argv) > 4 else (5242, 50, 64)
rng = np.random.default_rng(0)
init = (rng.standard_normal((n, n_emb)) * 0.1).astype(np.float32)
P = B * 4096
i = torch.as_tensor(rng.integers(0, n, P).astype(np.int32)).cuda()
j = torch.as_tensor(rng.integers(0, n, P).astype(np.int32)).cuda()
r = torch.as_tensor(rng.random(P).astype(np.float32)).cuda()
starts = (rng.integers(0, P // B, S) * B).astype(np.int64)
g = Generator(n, init)
g.train_steps(i, j, r, starts, B, persistent=how)
torch.cuda.synchronize()
print("done", how, S)
