"""GPU-vs-oracle parity ON THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs C2, C3, C5-shape).

The fixtures of tests/test_walk_gpu.py are small graphs; these tests run the same bit-exact comparison (BFS trees,
sampled nodes, statuses, per-walk counters, father-removal bits, G paths) on the graphs bench.py times:

  C3  power-law N = 1M, avg-deg 20, n_emb = 128  -- 13 828-neighbour hubs, candidate lists beyond the shared score
      buffer under the real degree mix, the depth-1 CDF sharing with its queue, walk_order; the root set contains the
      top-degree node and neighbours of it;
  C2  Erdos-Renyi N = 100k, avg-deg 10, n_emb = 128;
  C5-shape  power-law N = 2M (global visited bitmap in the tree builder), avg-deg 8, n_emb = 256 (ld 256), with a
      path-like tail so that BFS depth >> 12 and generator walks get long (max_path_len = 64 must hold them).

Oracle = oracle/gg_oracle.c (T1, canonical arithmetic), reference semantics graph_gan.py:84-108, 182-270.
The per-root walk count is capped (the Philox draw of walk k does not depend on the count), so that the
scalar C oracle finishes in seconds even for the 13 828-walk root.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(hg, emb_h, roots, sample_cap, n_sample_gen, cuda_device, hub_threshold=128, max_path=64, seed=11, bias_h=None,
             flat_steps=0):
    import torch
    from graphgan_b200 import graph as G, sampler as S
    from oracle import canonical as can
    roots = np.asarray(roots, np.int32)
    dg = G.DeviceGraph(hg, cuda_device)
    smp = S.WalkSampler(dg, hub_threshold=hub_threshold)
    smp.flat_steps = flat_steps       # 0: persistent walk kernel; > 0: level-synchronous steps first (csrc/walk.cu: flat_*_kernel)
    trees = smp.build_trees(roots)
    par = trees.parent_arrays().cpu().numpy()
    want = can.bfs_parents(hg.indptr, hg.adj, roots)
    assert np.array_equal(par, want), "BFS trees differ from the FIFO-order oracle"
    del want
    emb = S.pad_embedding(emb_h, cuda_device)
    if bias_h is None:
        bias_h = np.random.RandomState(5).normal(0, 0.1, hg.n_node).astype(np.float32)
    bias = torch.as_tensor(bias_h).to(cuda_device)
    E = can.pad_rows(emb_h, int(emb.shape[1]))
    sn = np.minimum(hg.degrees()[roots], sample_cap).astype(np.int64)
    bits = np.zeros(dg.n_bit_words, np.uint32)
    stats = {}
    for for_d, tag, num in ((True, 21, sn), (False, 22, np.full(len(roots), n_sample_gen, np.int64))):
        ref = can.walk_pass(E, bias_h, hg.indptr, hg.adj, roots, par, num, for_d, bits, seed=seed, pass_tag=tag,
                            max_path=0 if for_d else max_path)
        out = smp.run(emb, bias, trees, torch.as_tensor(num).to(cuda_device) if for_d else int(n_sample_gen), for_d,
                      seed=seed, pass_tag=tag, max_path=0 if for_d else max_path)
        W = ref.samples.shape[0]
        assert np.array_equal(out.status.cpu().numpy()[:W], ref.status)
        assert np.array_equal(out.samples.cpu().numpy()[:W], ref.samples)
        assert np.array_equal(out.wsteps.cpu().numpy()[:W], ref.wsteps)
        assert np.array_equal(out.wsuml.cpu().numpy()[:W], ref.wsuml)
        assert np.array_equal(out.root_ok.cpu().numpy()[:len(roots)], ref.root_ok)
        assert np.array_equal(dg.d1_bits.cpu().numpy().view(np.uint32), bits)
        cnt = out.counters_host()
        assert (cnt["steps"], cnt["sum_l"]) == (ref.steps, ref.sum_l)
        assert cnt["path_overflow"] == ref.path_overflow
        if not for_d:
            assert np.array_equal(out.path_len.cpu().numpy()[:W], ref.path_len)
            gp, rp = out.paths.cpu().numpy(), ref.paths
            done = np.flatnonzero(ref.status == can.DONE)
            for w in done:
                assert np.array_equal(gp[w, :ref.path_len[w]], rp[w, :ref.path_len[w]])
        stats["d" if for_d else "g"] = dict(walks=W, steps=ref.steps, sum_l=ref.sum_l, max_l=ref.max_l,
                                            overflow=ref.path_overflow,
                                            max_path=int(ref.path_len.max()) if not for_d else 0)
    return stats


@pytest.mark.parametrize("flat_steps", [0, 4])
def test_c3_powerlaw_1m_with_hub_roots(flat_steps, cuda_device):
    """BASELINE.json configs[2]: the graph bench.py times.  Roots: the top-degree node, three of its neighbours, other
    hubs, and a spread of ordinary roots."""
    from graphgan_b200 import graph as G, synth
    n, d = 1_000_000, 128
    hg = G.HostGraph(synth.power_law(n, 20, seed=0), None, n_node=n)
    deg = np.diff(hg.indptr)
    top = int(np.argmax(deg))
    assert deg[top] > 10000
    nb = hg.adj[hg.indptr[top]:hg.indptr[top + 1]]
    rs = np.random.RandomState(3)
    ordinary = rs.choice(np.flatnonzero(hg.degrees() > 0), 56, replace=False)
    roots = np.unique(np.concatenate([[top, 1, 7, 300], nb[[0, len(nb) // 2, len(nb) - 1]], ordinary]))
    emb_h = synth.embeddings(n, d, seed=1)
    st = _compare(hg, emb_h, roots, sample_cap=40, n_sample_gen=20, cuda_device=cuda_device, flat_steps=flat_steps)
    assert st["d"]["max_l"] > 2048          # candidate lists longer than the shared-memory score buffer were walked
    assert st["g"]["overflow"] == 0


@pytest.mark.parametrize("flat_steps", [0, 1, 3])
def test_c2_erdos_renyi_100k(flat_steps, cuda_device):
    """BASELINE.json configs[1]: ER N = 100k, avg-deg 10, n_emb = 128; 256 roots with their full sample_num."""
    from graphgan_b200 import graph as G, synth
    n, d = 100_000, 128
    hg = G.HostGraph(synth.erdos_renyi(n, 10, seed=0), None, n_node=n)
    roots = synth.pick_roots(hg.degrees(), 256, seed=0)
    st = _compare(hg, synth.embeddings(n, d, seed=1), roots, sample_cap=1 << 30, n_sample_gen=20, cuda_device=cuda_device,
                  flat_steps=flat_steps)
    assert st["d"]["walks"] == int(hg.degrees()[roots].sum())
    assert st["g"]["overflow"] == 0


@pytest.mark.parametrize("flat_steps", [0, 14])
def test_c5_shape_deep_trees_ld256(flat_steps, cuda_device):
    """BASELINE.json configs[4] shape: avg-deg 8, n_emb = 256, deep BFS trees.  N = 2M (the tree builder's visited
    bitmap no longer fits in shared memory) plus a 44-node path hanging off a low-degree node, and roots at and next to
    the path's far end: BFS depth > 40.  The generator's biases grow by 8 per node towards the anchor, so walks that
    start at the far end run down the whole path (~45-50 nodes); config.max_path_len = 64 must hold them (a longer
    walk is a RuntimeError in GraphGAN.prepare_data_for_g)."""
    from graphgan_b200 import graph as G, synth
    n0, d, tail = 2_000_000, 256, 44
    base = synth.power_law(n0, 8, seed=2)
    n = n0 + tail
    anchor = n0 - 5                                        # a low-degree node of the power-law part
    chain = np.stack([np.concatenate([[anchor], np.arange(n0, n - 1)]), np.arange(n0, n)], 1)
    hg = G.HostGraph(np.concatenate([base, chain]), None, n_node=n)
    rs = np.random.RandomState(4)
    ordinary = rs.choice(np.flatnonzero(hg.degrees()[:n0] > 0), 20, replace=False)
    roots = np.unique(np.concatenate([[0, anchor, n - 1, n - 2, n0 + 3], ordinary]))
    bias_h = np.random.RandomState(5).normal(0, 0.1, n).astype(np.float32)
    bias_h[n0:] = 8.0 * (tail - 1 - np.arange(tail))
    st = _compare(hg, synth.embeddings(n, d, seed=6, sigma=0.35), roots, sample_cap=48, n_sample_gen=20,
                  cuda_device=cuda_device, bias_h=bias_h, flat_steps=flat_steps)
    assert st["g"]["overflow"] == 0 and st["d"]["overflow"] == 0
    assert 40 < st["g"]["max_path"] <= 64                  # deep walks really occurred, and fit
