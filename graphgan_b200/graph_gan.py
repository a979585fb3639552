"""Re-hosted trainer: the reference's ``GraphGAN`` class (src/GraphGAN/graph_gan.py:17-319) with the
same method names, whose bodies batch over roots and call the CUDA hot path.

What is accelerated: ``construct_trees`` (GPU BFS), ``sample`` /
``prepare_data_for_d`` / ``prepare_data_for_g`` (K1 + finalize + row/pair emission, all on device),
``discriminator.reward`` and the two update ops (K2 + K3).  What is only re-hosted (thin, not
accelerated): the epoch loop, the text dumps and the link-prediction check.

Differences from the reference that a caller can observe, all documented in DESIGN.md:
  * training rows are device int32/fp32 tensors, not Python lists (len() and slicing still work);
  * walks draw from Philox keyed by (config.seed, pass, root, walk, step) instead of the global
    MT19937 stream, so results do not depend on the number of GPUs or on root order;
  * BFS trees are parent arrays built on the GPU; there is no pickle cache (config.cache_filename
    is ignored) because the dict-of-lists form is O(N^2).
"""
import os

import numpy as np

from . import config
from . import evaluation as lp
from . import io
from .discriminator import Discriminator
from .generator import Generator
from .graph import DeviceGraph, HostGraph
from .sampler import WalkSampler
from .session import Session
from ._cabi import ptr
from . import _cabi
import ctypes as C


def _device():
    dev = config.device
    if "LOCAL_RANK" in os.environ and str(dev).startswith("cuda"):
        dev = "cuda:%d" % int(os.environ["LOCAL_RANK"])
    return dev


class GraphGAN(object):
    def __init__(self, host_graph=None, node_embed_init_d=None, node_embed_init_g=None):
        import torch
        self.torch = torch
        self.device = torch.device(_device())
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        print("reading graphs...")
        self.host_graph = host_graph if host_graph is not None else HostGraph.from_files(config.train_filename,
                                                                                         config.test_filename)
        self.n_node = self.host_graph.n_node
        self.graph = self.host_graph            # graph[i] == self.host_graph.neighbors(i)
        self.root_nodes = [i for i in range(self.n_node)]

        print("reading initial embeddings...")
        # graph_gan.py:24-29: the discriminator's file is read first (it consumes the global RNG first)
        self.node_embed_init_d = node_embed_init_d if node_embed_init_d is not None else io.read_embeddings(
            filename=config.pretrain_emb_filename_d, n_node=self.n_node, n_embed=config.n_emb)
        self.node_embed_init_g = node_embed_init_g if node_embed_init_g is not None else io.read_embeddings(
            filename=config.pretrain_emb_filename_g, n_node=self.n_node, n_embed=config.n_emb)

        self.device_graph = DeviceGraph(self.host_graph, self.device)
        self.sampler = WalkSampler(self.device_graph)
        self.lib = _cabi.lib()

        # BFS trees: resident for all roots when they fit (the reference's pickle cache, graph_gan.py:31-46)
        self.trees, self._tree_key = None, None
        import torch.distributed as _d
        if not (_d.is_available() and _d.is_initialized() and _d.get_world_size() > 1) and \
                self._tree_bytes(self.n_node) <= config.tree_cache_bytes:
            print("constructing BFS-trees...")
            self.trees = self.construct_trees(self.root_nodes)

        print("building GAN model...")
        self.discriminator = None
        self.generator = None
        self.build_generator()
        self.build_discriminator()
        self.sess = Session()
        self.shuffle_rng = np.random.RandomState(config.seed)
        self.pass_counter = 0
        self.last_counters = {}
        # one process per GPU: roots are sharded, rows all-gathered, updates data parallel (parallel.py)
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        if self.dist:
            from .parallel import DataParallelStep
            self._dp_d, self._dp_g = DataParallelStep(self.discriminator), DataParallelStep(self.generator)

    # ------------------------------------------------------------------ trees (graph_gan.py:63-108)
    def _tree_bytes(self, n_roots):
        """device bytes of the tree rows of `n_roots` roots: one bit per walk-CSR entry each (csrc/bfs.cu)"""
        return n_roots * 4 * ((int(self.host_graph.adj.shape[0]) + 31) // 32 + 1)

    def construct_trees(self, nodes):
        """BFS trees of ``nodes`` -> sampler.TreeBatch (parent arrays on the GPU)."""
        return self.sampler.build_trees(np.asarray(nodes, np.int32))

    def construct_trees_with_mp(self, nodes):
        """Kept for source compatibility (graph_gan.py:63-82); the GPU builder needs no process pool."""
        self.trees = self.construct_trees(nodes)

    def build_generator(self):
        self.generator = Generator(n_node=self.n_node, node_emd_init=self.node_embed_init_g, device=self.device)

    def build_discriminator(self):
        self.discriminator = Discriminator(n_node=self.n_node, node_emd_init=self.node_embed_init_d, device=self.device)

    # ------------------------------------------------------------------ root batching
    def _root_batches(self, roots):
        roots = np.asarray(roots, np.int32)
        if self.dist:   # this rank's contiguous, degree-balanced block of the root list (same for D and G passes)
            from .parallel import balanced_root_ranges
            lo, hi = balanced_root_ranges(self.host_graph.degrees()[roots] + 1, self.world)[self.rank]
            roots = roots[lo:hi]
            key = (lo, hi, hash(roots.tobytes()))           # the cached trees belong to exactly these roots
            if self.trees is not None and self._tree_key != key:
                self.trees = None
            if self.trees is None and self._tree_bytes(max(hi - lo, 1)) <= config.tree_cache_bytes:
                self.trees, self._tree_key = self.construct_trees(roots), key
            if self.trees is not None:
                yield self.trees
                return
        if self.trees is not None and roots.shape[0] == self.n_node and np.array_equal(roots, np.arange(self.n_node)):
            yield self.trees
            return
        for s in range(0, roots.shape[0], config.root_batch):
            yield self.construct_trees(roots[s:s + config.root_batch])

    def _next_tag(self):
        self.pass_counter += 1
        return self.pass_counter

    # ------------------------------------------------------------------ graph_gan.py:182-202
    def prepare_data_for_d(self, roots=None):
        """positive and negative samples for the discriminator -> (center_nodes, neighbor_nodes, labels)
        as device tensors (int32, int32, fp32), rows in the reference's order."""
        torch = self.torch
        tag = self._next_tag()
        cs, ns, ls = [], [], []
        tot = dict(steps=0, sum_l=0, accepted=0)
        for trees in self._root_batches(self.root_nodes if roots is None else roots):
            sample_num = self.device_graph.raw_deg[trees.roots.long()]
            out = self.sampler.run(self.generator.emb, self.generator.bias_t, trees, sample_num, True,
                                   seed=config.seed, pass_tag=tag, update_ratio=float(config.update_ratio))
            c, n, l, n_rows = self.sampler.emit_d_rows(out)
            k = int(n_rows.item())
            cs.append(c[:k]); ns.append(n[:k]); ls.append(l[:k])
            cnt = out.counters_host()
            for key in tot:
                tot[key] += cnt[key]
        self.last_counters = tot
        cat = lambda xs, dt: torch.cat(xs) if xs else torch.zeros(0, dtype=dt, device=self.device)
        c, n, l = cat(cs, torch.int32), cat(ns, torch.int32), cat(ls, torch.int32)
        if self.dist:   # every rank ends up with the full row lists, in root order
            from .parallel import all_gather_varlen
            c, n, l = all_gather_varlen(c), all_gather_varlen(n), all_gather_varlen(l)
        return c, n, l.float()

    # ------------------------------------------------------------------ graph_gan.py:204-223
    def prepare_data_for_g(self, roots=None):
        """sample nodes for the generator -> (node_1, node_2, reward) device tensors."""
        torch = self.torch
        tag = self._next_tag()
        n1s, n2s = [], []
        st = torch.cuda.current_stream(self.device).cuda_stream
        for trees in self._root_batches(self.root_nodes if roots is None else roots):
            out = self.sampler.run(self.generator.emb, self.generator.bias_t, trees, int(config.n_sample_gen), False,
                                   seed=config.seed, pass_tag=tag, update_ratio=float(config.update_ratio),
                                   max_path=config.max_path_len)
            if out.counters_host()["path_overflow"]:
                raise RuntimeError("a walk is longer than config.max_path_len=%d" % config.max_path_len)
            W = out.n_walks
            pair_ptr = torch.empty(W + 1, dtype=torch.int64, device=self.device)
            n_pairs = torch.zeros(1, dtype=torch.int64, device=self.device)
            # count first (capacity 0), then emit
            _cabi.check(self.lib.gg_window_pairs(W, ptr(out.paths), ptr(out.path_len), out.max_path, int(config.window_size),
                                                 ptr(pair_ptr), None, None, ptr(n_pairs), 0, st), "gg_window_pairs")
            m = int(n_pairs.item())
            n1 = torch.empty(max(m, 1), dtype=torch.int32, device=self.device)
            n2 = torch.empty(max(m, 1), dtype=torch.int32, device=self.device)
            _cabi.check(self.lib.gg_window_pairs(W, ptr(out.paths), ptr(out.path_len), out.max_path, int(config.window_size),
                                                 ptr(pair_ptr), ptr(n1), ptr(n2), ptr(n_pairs), m, st), "gg_window_pairs")
            n1s.append(n1[:m]); n2s.append(n2[:m])
        node_1 = torch.cat(n1s) if n1s else torch.zeros(0, dtype=torch.int32, device=self.device)
        node_2 = torch.cat(n2s) if n2s else torch.zeros(0, dtype=torch.int32, device=self.device)
        if self.dist:
            from .parallel import all_gather_varlen
            node_1, node_2 = all_gather_varlen(node_1), all_gather_varlen(node_2)
        reward = self.discriminator.reward_pairs(node_1, node_2)     # graph_gan.py:220-222, one fetch for all pairs
        return node_1, node_2, reward

    # ------------------------------------------------------------------ graph_gan.py:225-270 (single-root form)
    def sample(self, root, tree, sample_num, for_d):
        """Reference-shaped call: one root -> (samples, paths) as Python lists, or (None, None).
        ``tree`` may be a TreeBatch holding this root or anything else (then the tree is rebuilt)."""
        from .sampler import DONE, TreeBatch
        torch = self.torch
        if not (isinstance(tree, TreeBatch) and int(tree.roots.shape[0]) == 1 and int(tree.roots[0]) == root):
            tree = self.construct_trees([root])
        if sample_num == 0:
            return [], []
        out = self.sampler.run(self.generator.emb, self.generator.bias_t, tree, int(sample_num), bool(for_d),
                               seed=config.seed, pass_tag=self._next_tag(), max_path=config.max_path_len)
        if not int(out.root_ok[0]):
            return None, None
        samples = out.samples[:sample_num].cpu().tolist()
        plen, paths = out.path_len.cpu().numpy(), out.paths.cpu().numpy()
        return samples, [paths[w, :plen[w]].tolist() for w in range(sample_num)]

    @staticmethod
    def get_node_pairs_from_path(path):
        """path = [1, 0, 2, 4, 2], window_size = 2 -> [[1,0],[1,2],[0,1],[0,2],[0,4],[2,1],[2,0],[2,4],[4,0],[4,2]]
        (graph_gan.py:272-291; the batched device version is gg_window_pairs)."""
        body, w, pairs = path[:-1], config.window_size, []
        for pos, center in enumerate(body):
            for other in range(max(pos - w, 0), min(pos + w + 1, len(body))):
                if other != pos:
                    pairs.append([center, body[other]])
        return pairs

    # ------------------------------------------------------------------ graph_gan.py:122-180
    def train(self):
        torch = self.torch
        ckpt = os.path.join(config.model_log, "model.checkpoint.pt")
        if config.load_model and os.path.isfile(ckpt):
            print("loading the checkpoint: %s" % ckpt)
            self.load(ckpt)
        self.write_embeddings_to_file()
        self.evaluation(self)
        print("start training...")
        for epoch in range(config.n_epochs):
            print("epoch %d" % epoch)
            if epoch > 0 and epoch % config.save_steps == 0:
                self.save(ckpt)
            # D-steps
            center_nodes = neighbor_nodes = labels = None
            for d_epoch in range(config.n_epochs_dis):
                if d_epoch % config.dis_interval == 0:
                    center_nodes, neighbor_nodes, labels = self.prepare_data_for_d()
                train_size = len(center_nodes)
                start_list = list(range(0, train_size, config.batch_size_dis))
                self.shuffle_rng.shuffle(start_list)
                # the per-batch sess.run loop of graph_gan.py:152-157, enqueued from C (identical steps)
                if self.dist:
                    self._dp_d.train_steps(center_nodes, neighbor_nodes, labels, start_list, config.batch_size_dis)
                else:
                    self.discriminator.train_steps(center_nodes, neighbor_nodes, labels, start_list, config.batch_size_dis)
            # G-steps
            node_1 = node_2 = reward = None
            for g_epoch in range(config.n_epochs_gen):
                if g_epoch % config.gen_interval == 0:
                    node_1, node_2, reward = self.prepare_data_for_g()
                train_size = len(node_1)
                start_list = list(range(0, train_size, config.batch_size_gen))
                self.shuffle_rng.shuffle(start_list)
                if self.dist:
                    self._dp_g.train_steps(node_1, node_2, reward, start_list, config.batch_size_gen)
                else:
                    self.generator.train_steps(node_1, node_2, reward, start_list, config.batch_size_gen)   # graph_gan.py:171-176
            self.write_embeddings_to_file()
            self.evaluation(self)
        print("training completes")

    # ------------------------------------------------------------------ graph_gan.py:293-319
    def write_embeddings_to_file(self):
        if self.rank != 0:      # replicas are bit-identical; one writer
            return
        modes = [self.generator, self.discriminator]
        for i in range(2):
            os.makedirs(os.path.dirname(config.emb_filenames[i]) or ".", exist_ok=True)
            if config.binary_embeddings:      # [N, n_emb] fp32 straight from the device (the text form is minutes at N = 1M)
                io.write_embeddings_binary(config.emb_filenames[i] + ".f32", modes[i])
            if config.text_embeddings:
                io.write_embeddings(config.emb_filenames[i], self.sess.run(modes[i].embedding_matrix))

    @staticmethod
    def evaluation(self):
        results = []
        if getattr(self, "rank", 0) != 0:
            return results
        if config.app == "link_prediction":
            modes = [self.generator, self.discriminator]
            for i in range(2):
                if config.device_eval:        # from the device-resident embeddings (csrc/eval.cu): no text round trip
                    lpe = lp.DeviceLinkPredictEval(modes[i], config.test_filename, config.test_neg_filename)
                else:
                    lpe = lp.LinkPredictEval(config.emb_filenames[i], config.test_filename, config.test_neg_filename,
                                             self.n_node, config.n_emb)
                results.append(config.modes[i] + ":" + str(lpe.eval_link_prediction()) + "\n")
        os.makedirs(os.path.dirname(config.result_filename) or ".", exist_ok=True)
        with open(config.result_filename, mode="a+") as f:
            f.writelines(results)
        return results

    # ------------------------------------------------------------------ checkpoint (tf.train.Saver stand-in)
    def save(self, path):
        """Replicas are bit-identical, so rank 0 alone writes (to a temporary file, then an atomic rename).  The
        father-removal bits (graph_gan.py:258-259) are per root, i.e. per rank shard: they are OR-reduced over the
        ranks first, so the file holds the removals of every root."""
        torch = self.torch
        bits = self.device_graph.d1_bits.clone()
        if self.dist:       # OR over the ranks (NCCL has no bitwise reduction: gather, then OR locally)
            parts = [torch.empty_like(bits) for _ in range(self.world)]
            self.dist.all_gather(parts, bits)
            for p_ in parts:
                bits |= p_
        if self.rank == 0:
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            rs = self.shuffle_rng.get_state()
            state = {"generator": self.generator.state_dict(), "discriminator": self.discriminator.state_dict(),
                     "d1_bits": bits.cpu(), "pass_counter": torch.tensor(self.pass_counter),
                     "shuffle_rng": {"keys": torch.from_numpy(rs[1].astype(np.int64)), "pos": int(rs[2]),
                                     "has_gauss": int(rs[3]), "cached_gaussian": float(rs[4])}}
            tmp = "%s.tmp.%d" % (path, os.getpid())
            torch.save(state, tmp)
            os.replace(tmp, path)
        if self.dist:
            self.dist.barrier()

    def load(self, path):
        """Every rank reads the same file; the OR-ed removal bits only add bits for roots of other shards, which this
        rank's walks never look at."""
        sd = self.torch.load(path, weights_only=True, map_location="cpu")
        self.generator.load_state_dict(sd["generator"])
        self.discriminator.load_state_dict(sd["discriminator"])
        self.device_graph.d1_bits.copy_(sd["d1_bits"].to(self.device))
        self.pass_counter = int(sd["pass_counter"])
        r = sd["shuffle_rng"]
        self.shuffle_rng.set_state(("MT19937", r["keys"].numpy().astype(np.uint32), r["pos"], r["has_gauss"],
                                    r["cached_gaussian"]))
