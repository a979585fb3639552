"""Config C1 end to end: train GraphGAN on the CA-GrQc fixture with the reference's default hyper-parameters
through the re-hosted trainer, log link-prediction accuracy per epoch and wall time.  (The reference's numbers
for the same run: epoch-0 accuracy 0.7598 of the shipped pretrain embeddings; no throughput is published.)"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.golden import loader
from graphgan_b200 import config, graph as G, io
from graphgan_b200.graph_gan import GraphGAN

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
c = loader.load("cagrqc")
tmp = tempfile.mkdtemp()
def wr(name, e):
    p = os.path.join(tmp, name)
    with open(p, "w") as f:
        f.write("".join("%d\t%d\n" % (a, b) for a, b in e))
    return p
config.train_filename, config.test_filename = wr("train.txt", c.train_edges), wr("test.txt", c.test_edges)
config.test_neg_filename = wr("test_neg.txt", c.test_neg_edges)
pre = os.path.join(tmp, "pre.emb")
io.write_embeddings(pre, c.emb_g)
config.pretrain_emb_filename_d = config.pretrain_emb_filename_g = pre
config.emb_filenames = [os.path.join(tmp, "gen.emb"), os.path.join(tmp, "dis.emb")]
config.result_filename, config.model_log = os.path.join(tmp, "res.txt"), os.path.join(tmp, "log") + "/"
config.n_epochs = epochs
t0 = time.time()
gan = GraphGAN()
t1 = time.time()
gan.train()
import torch; torch.cuda.synchronize()
t2 = time.time()
print("setup %.1f s (incl. %d BFS trees), train %d epochs %.1f s (%.2f s/epoch)" % (t1 - t0, gan.n_node, epochs, t2 - t1, (t2 - t1) / max(epochs, 1)))
print("optimizer steps: gen %d dis %d" % (gan.generator.step_count, gan.discriminator.step_count))
print(open(config.result_filename).read())
