"""Worker for the multi-process tests (launched by tests/test_dist.py through torch.distributed.run).

  mode "cpu": gloo, host logic only (no kernels): partitioning + variable-length all-gather.
  mode "gpu": nccl, one rank per GPU: root-sharded sampling is invariant, data-parallel updates match the
              single-GPU step and keep the replicas bit-identical.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(mode):
    import torch
    import torch.distributed as dist
    from graphgan_b200 import parallel
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode == "cpu":
        dist.init_process_group("gloo")
        # block ranges tile [0, n) in rank order
        for n in (0, 1, 7, 64, 65):
            rngs = [parallel.block_range(n, r, world) for r in range(world)]
            assert rngs[0][0] == 0 and rngs[-1][1] == n and all(rngs[k][1] == rngs[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in rngs) - min(b - a for a, b in rngs) <= 1
        w = np.random.RandomState(0).pareto(1.5, size=1000) + 1
        rr = parallel.balanced_root_ranges(w, world)
        assert rr[0][0] == 0 and rr[-1][1] == 1000 and all(rr[k][1] == rr[k + 1][0] for k in range(world - 1))
        tot = [w[a:b].sum() for a, b in rr]
        assert max(tot) <= w.sum() / world + w.max() + 1e-9
        # variable-length gather keeps rank order
        mine = torch.arange(rank * 100, rank * 100 + 3 + 2 * rank, dtype=torch.int32)
        got = parallel.all_gather_varlen(mine)
        want = torch.cat([torch.arange(r * 100, r * 100 + 3 + 2 * r, dtype=torch.int32) for r in range(world)])
        assert torch.equal(got, want)
        empty = parallel.all_gather_varlen(torch.zeros(0 if rank else 2, dtype=torch.float32))
        assert empty.shape[0] == 2
        dist.barrier()
        if rank == 0:
            print("DIST_CPU_OK")
        dist.destroy_process_group()
        return
    # ------------------------------------------------------------------ gpu
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from graphgan_b200 import graph as G, sampler as S, synth
    from graphgan_b200.discriminator import Discriminator
    from graphgan_b200.generator import Generator
    n, d = 4000, 128
    hg = G.HostGraph(synth.power_law(n, 12, seed=2), None, n_node=n)
    dg = G.DeviceGraph(hg, dev)
    smp = S.WalkSampler(dg)
    emb = S.pad_embedding(synth.embeddings(n, d, seed=3), dev)
    bias = torch.zeros(n, dtype=torch.float32, device=dev)
    roots = synth.pick_roots(hg.degrees(), 600, seed=4)
    # (1) root-sharded sampling == single-GPU sampling (Philox is keyed by root, walk, step)
    lo, hi = parallel.balanced_root_ranges(hg.degrees()[roots], world)[rank]
    mine = roots[lo:hi]
    t = smp.build_trees(mine)
    out = smp.run(emb, bias, t, dg.raw_deg[t.roots.long()], True, seed=11, pass_tag=1)
    c, nb, lb, k = smp.emit_d_rows(out)
    k = int(k.item())
    gc, gn, gl = (parallel.all_gather_varlen(x[:k]) for x in (c, nb, lb))
    if rank == 0:
        dg.reset_tree_mutations()
        tf = smp.build_trees(roots)
        of = smp.run(emb, bias, tf, dg.raw_deg[tf.roots.long()], True, seed=11, pass_tag=1)
        fc, fn, fl, fk = smp.emit_d_rows(of)
        fk = int(fk.item())
        assert fk == gc.shape[0] and torch.equal(fc[:fk], gc) and torch.equal(fn[:fk], gn) and torch.equal(fl[:fk], gl)
    # (2) data-parallel updates
    rs = np.random.RandomState(5)
    e0 = rs.normal(0, 0.5, size=(n, d))
    for cls, mode in ((Discriminator, 0), (Generator, 1)):
        single, repl = cls(n, e0, device=dev), cls(n, e0, device=dev)
        dp = parallel.DataParallelStep(repl)
        for step in range(5):
            B = (64, 64, 63, 5, 1)[step]
            i, j = rs.randint(0, n, B).astype(np.int32), rs.randint(0, n, B).astype(np.int32)
            if B > 4:
                i[1], j[1] = i[0], j[0]
            aux = ((rs.random_sample(B) < 0.5) if mode == 0 else rs.random_sample(B) * 3).astype(np.float32)
            single.step(i, j, aux)
            dp.step(i, j, aux)
            a, b = single.emb.double(), repl.emb.double()
            assert float((a - b).norm() / a.norm()) <= 1e-5, (cls.__name__, step)
            assert int((repl.row_slot != -1).sum()) == 0
        # replicas bit-identical across ranks
        rows = [torch.empty_like(repl.emb) for _ in range(world)]
        dist.all_gather(rows, repl.emb)
        assert all(torch.equal(rows[0], r) for r in rows[1:])
        # the C-side batch loop (gg_dp_train_steps: one collective per step, no host round trip) == step by step
        M, Bs = 300, 64
        ii, jj = rs.randint(0, n, M).astype(np.int32), rs.randint(0, n, M).astype(np.int32)
        ax = ((rs.random_sample(M) < 0.5) if mode == 0 else rs.random_sample(M) * 3).astype(np.float32)
        starts = list(range(0, M, Bs))
        rs.shuffle(starts)
        ra, rb = cls(n, e0, device=dev), cls(n, e0, device=dev)
        da, db = parallel.DataParallelStep(ra), parallel.DataParallelStep(rb)
        for s0 in starts:
            da.step(ii[s0:s0 + Bs], jj[s0:s0 + Bs], ax[s0:s0 + Bs])
        before = db.stats()["collectives_issued"]
        db.train_steps(ii, jj, ax, starts, Bs)
        for name in ("emb", "bias_t", "m_emb", "v_emb", "m_bias", "v_bias"):
            assert torch.equal(getattr(ra, name), getattr(rb, name)), name
        assert ra.beta1_power == rb.beta1_power and ra.step_count == rb.step_count
        st = db.stats()
        assert st["comm_nranks"] == world and st["collectives_issued"] - before == len(starts)
        # the peer-memory transport (gradient + exchange fused in one kernel, NVLink P2P stores) == the NCCL transport
        rc = cls(n, e0, device=dev)
        dc = parallel.DataParallelStep(rc, transport="p2p")
        dc.train_steps(ii, jj, ax, starts, Bs)
        for name in ("emb", "bias_t", "m_emb", "v_emb", "m_bias", "v_bias"):
            assert torch.equal(getattr(ra, name), getattr(rc, name)), ("p2p", name)
        for s0 in starts[:3]:
            dc.step(ii[s0:s0 + 5], jj[s0:s0 + 5], ax[s0:s0 + 5])      # short batches: some ranks own no row
            da.step(ii[s0:s0 + 5], jj[s0:s0 + 5], ax[s0:s0 + 5])
        assert torch.equal(ra.emb, rc.emb)
    # (3) the re-hosted trainer under torch.distributed: sharded sampling + data-parallel updates
    from graphgan_b200 import config
    from graphgan_b200.graph_gan import GraphGAN
    from tests.golden import loader
    import tempfile
    c = loader.load("rand1200")
    box = [tempfile.mkdtemp() if rank == 0 else None]       # ONE directory for all ranks: the checkpoint is written by rank 0
    dist.broadcast_object_list(box, src=0)                   # and read by every rank (part 4)
    tmp = box[0]
    config.n_emb, config.n_epochs, config.n_epochs_dis, config.dis_interval = 50, 1, 1, 1
    config.n_epochs_gen, config.gen_interval, config.n_sample_gen, config.seed = 1, 1, 2, 9
    config.app = "none"
    config.emb_filenames = [os.path.join(tmp, "g%d.emb" % rank), os.path.join(tmp, "d%d.emb" % rank)]
    config.result_filename, config.model_log = os.path.join(tmp, "r%d.txt" % rank), tmp + "/"
    hgc = G.HostGraph(c.train_edges, c.test_edges)
    gan = GraphGAN(host_graph=hgc, node_embed_init_d=c.emb_d, node_embed_init_g=c.emb_g)
    assert gan.world == world
    ce, ne, la = gan.prepare_data_for_d()
    if rank == 0:   # the same pass on one GPU (no process group in this instance)
        solo = GraphGAN.__new__(GraphGAN)
        solo.__dict__.update(gan.__dict__)
        solo.dist, solo.rank, solo.world, solo.trees, solo._tree_key = None, 0, 1, None, None
        solo.pass_counter = gan.pass_counter - 1
        solo.device_graph.reset_tree_mutations()
        se, sn, sl = solo.prepare_data_for_d()
        assert torch.equal(se, ce) and torch.equal(sn, ne) and torch.equal(sl, la)
    dist.barrier()
    gan.device_graph.reset_tree_mutations()
    gan.pass_counter = 0
    gan.trees, gan._tree_key = None, None
    config.n_epochs, config.save_steps = 2, 1
    gan.train()                       # saves (rank 0, after OR-reducing the removal bits) at the start of epoch 1
    rows = [torch.empty_like(gan.generator.emb) for _ in range(world)]
    dist.all_gather(rows, gan.generator.emb)
    assert all(torch.equal(rows[0], r) for r in rows[1:])
    dist.barrier()
    # (4) save -> load -> continue == uninterrupted, under sharding (every rank loads the same file)
    config.n_epochs, config.load_model = 1, True
    gan2 = GraphGAN(host_graph=hgc, node_embed_init_d=c.emb_d, node_embed_init_g=c.emb_g)
    gan2.train()
    for ma, mb in ((gan.generator, gan2.generator), (gan.discriminator, gan2.discriminator)):
        for name in ("emb", "bias_t", "m_emb", "v_emb"):
            assert torch.equal(getattr(ma, name), getattr(mb, name)), name
    dist.barrier()
    if rank == 0:
        print("DIST_GPU_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
