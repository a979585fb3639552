"""Debug: save -> load -> continue under sharding (torchrun, 2 GPUs): per-stage checksums of the uninterrupted run's
second epoch against the resumed run's epoch."""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def h(t):
    return hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:10]


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from graphgan_b200 import config, graph as G
    from graphgan_b200.graph_gan import GraphGAN
    from tests.golden import loader
    c = loader.load("rand1200")
    tmp = tempfile.mkdtemp()
    config.n_emb, config.n_epochs, config.n_epochs_dis, config.dis_interval = 50, 1, 1, 1
    config.n_epochs_gen, config.gen_interval, config.n_sample_gen, config.seed = 1, 1, 2, 9
    config.app = "none"
    config.emb_filenames = [os.path.join(tmp, "g%d.emb" % rank), os.path.join(tmp, "d%d.emb" % rank)]
    config.result_filename, config.model_log = os.path.join(tmp, "r%d.txt" % rank), "/tmp/dbg_ckpt/"
    hgc = G.HostGraph(c.train_edges, c.test_edges)
    log = {}

    def instrument(gan, name):
        log[name] = []
        L = log[name]
        od, og = gan.prepare_data_for_d, gan.prepare_data_for_g
        def pd(*a, **k):
            L.append(("pre_d G.emb", h(gan.generator.emb), "D.emb", h(gan.discriminator.emb), "bits", h(gan.device_graph.d1_bits),
                      "tag", gan.pass_counter, "b1p", float(gan.generator.beta1_power), float(gan.discriminator.beta1_power)))
            r = od(*a, **k)
            L.append(("d_rows", h(r[0]), h(r[1]), h(r[2]), int(r[0].shape[0])))
            return r
        def pg(*a, **k):
            L.append(("pre_g G.emb", h(gan.generator.emb), "D.emb", h(gan.discriminator.emb), "bits", h(gan.device_graph.d1_bits)))
            r = og(*a, **k)
            L.append(("g_pairs", h(r[0]), h(r[1]), h(r[2]), int(r[0].shape[0])))
            return r
        gan.prepare_data_for_d, gan.prepare_data_for_g = pd, pg

    gan = GraphGAN(host_graph=hgc, node_embed_init_d=c.emb_d, node_embed_init_g=c.emb_g)
    instrument(gan, "cont")
    config.n_epochs, config.save_steps = 2, 1
    gan.train()
    log["cont"].append(("end G.emb", h(gan.generator.emb), "D.emb", h(gan.discriminator.emb)))
    dist.barrier()
    config.n_epochs, config.load_model = 1, True
    gan2 = GraphGAN(host_graph=hgc, node_embed_init_d=c.emb_d, node_embed_init_g=c.emb_g)
    instrument(gan2, "resumed")
    gan2.train()
    log["resumed"].append(("end G.emb", h(gan2.generator.emb), "D.emb", h(gan2.discriminator.emb)))
    for r in range(world):
        dist.barrier()
        if r == rank:
            print("==== rank", rank)
            for name in ("cont", "resumed"):
                print("--", name)
                for e in log[name]:
                    print("  ", e)
            sys.stdout.flush()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
