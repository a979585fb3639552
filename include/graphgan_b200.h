/*
 * graphgan_b200.h -- C ABI of libgraphgan_b200.so, the B200 (sm_100a) implementation of
 * GraphGAN's scoring-and-sampling hot path.
 *
 * The reference (hwwang55/GraphGAN) has no native code and no FFI: its seam is the five
 * tf.Session.run(fetch, feed_dict) call sites of src/GraphGAN/graph_gan.py (154, 173, 220,
 * 238, 298).  Each entry point below states the reference code it replaces; the Python
 * binding a maintainer adds is shown in INTEGRATION.md (ctypes, graphgan_b200/_cabi.py).
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller (in the
 *     Python host: torch tensors used only as memory containers, passed as data_ptr()).
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it, no hidden
 *     synchronisation, no allocation.  Scratch is caller supplied (query *_scratch_bytes).
 *   - return 0 on success, non-zero on CUDA / argument error; gg_last_error() gives the
 *     message for the calling thread.  Nothing aborts.
 *   - embedding rows are fp32 [N, ld], zero padded, with ld = 32, 64, 128 or 256 (the smallest that holds n_emb):
 *     every entry point that takes ld rejects other values.
 *   - graph = two CSRs in the reference's adjacency-file order (src/utils.py:27-37):
 *       raw  : graph[i] as read (duplicates and self-loops kept) -> positives, sample_num
 *       walk : first occurrences only, self-loops dropped        -> BFS trees and walks
 */
#ifndef GRAPHGAN_B200_H
#define GRAPHGAN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_ABI_VERSION 3

/* walk status codes (per walk) */
enum { GG_NOTRUN = 0, GG_DONE = 1, GG_VOID = 2, GG_SKIPPED = 3 };
/* rng modes */
enum {
    GG_RNG_PHILOX = 0, /* Philox4x32-10, key=(seed), counter=(root, walk, step, pass_tag): order free */
    GG_RNG_STREAM = 1  /* caller supplied doubles consumed in the reference's sequential order
                          (np.random.rand at graph_gan.py:189/209, np.random.choice at :262) */
};
/* counters[] slots written by gg_walk_finalize (reference semantics: walks after a root's
 * first voiding walk do not exist) */
enum {
    GG_CNT_STEPS = 0, GG_CNT_SUML = 1, GG_CNT_ACCEPTED = 2, GG_CNT_OK_ROOTS = 3,
    GG_CNT_PATH_OVERFLOW = 4, GG_CNT_RAW_STEPS = 5, GG_CNT_RAW_SUML = 6, GG_CNT_STREAM_USED = 7,
    GG_CNT_ROWS_GATHERED = 8, /* embedding rows the walk kernel actually fetched (on-demand scores + cur rows) */
    /* warp-cycles (clock64, summed over warps) spent per phase of the walk kernel */
    GG_CNT_CYC_ENUM = 9, GG_CNT_CYC_SCORE = 10, GG_CNT_CYC_CHOOSE = 11, GG_CNT_CYC_STEP0 = 12,
    GG_CNT_CYC_STEP1 = 13, GG_CNT_CYC_STEP2P = 14, GG_CNT_CYC_WALK = 15,
    GG_CNT_SLOTS = 16
};

const char *gg_last_error(void);
int gg_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * K1: graph-softmax walk.  Replaces GraphGAN.sample (graph_gan.py:225-270) together with the
 * generator.all_score fetch at :238 (scores are computed on demand for the candidates only,
 * generator.py:21), utils.softmax (utils.py:131-133) and np.random.choice's inverse-CDF
 * draw (:262), for a whole batch of roots at once.
 * ------------------------------------------------------------------------------------------ */
typedef struct gg_walk_desc {
    int64_t n_node;
    int32_t ld;                 /* row stride in floats, multiple of 32 */
    const float *emb;           /* device [N, ld]  generator.embedding_matrix (generator.py:11-14) */
    const float *bias;          /* device [N]      generator.bias_vector      (generator.py:15)    */
    const int64_t *indptr;      /* device [N+1]    walk CSR */
    const int32_t *adj;         /* device [nnz]    */
    int64_t n_roots;
    const int32_t *roots;       /* device [R] root node ids (batch order = reference root order) */
    const uint32_t *tree_bits;  /* device [R, tree_words] BFS trees (gg_bfs_build): bit e of row k is set iff adj[e]
                                   is a child of entry e's source node in the tree of roots[k] */
    int64_t tree_words;         /* row stride of tree_bits in 32-bit words (gg_tree_words(nnz)) */
    const int64_t *walk_ptr;    /* device [R+1] exclusive prefix of per-root sample_num */
    int64_t n_walks;            /* = walk_ptr[R] */
    int32_t for_d;              /* graph_gan.py:225 `for_d` */
    int32_t rng_mode;
    uint32_t *d1_bits;          /* device bitset over walk-CSR entries: "father entry removed"
                                   (the in-place tree mutation of graph_gan.py:258-259).
                                   G mode reads it here; D mode sets it in gg_walk_finalize */
    uint64_t seed;
    uint32_t pass_tag;
    int32_t max_path;           /* row stride of paths (0: paths not recorded) */
    const double *stream;       /* device, GG_RNG_STREAM only */
    int64_t n_stream;
    double update_ratio;        /* config.update_ratio (graph_gan.py:189/209) */
    int32_t max_cand;           /* >= max walk-CSR degree + 1 */
    int32_t phase_mask;         /* 0 = whole call; 1 = only the depth-1 precompute (root steps + per-pair CDFs), 2 = only
                                   the walk kernel -- lets a profiler time the two stages of one pass separately */
    /* per-walk outputs, device [W] */
    int32_t *samples;           /* sampled node (graph_gan.py:265) or -1 */
    int32_t *status;
    int32_t *first_edge;        /* walk-CSR entry of the depth-1 node chosen at the root step */
    int32_t *wsteps;            /* choices made */
    int32_t *wsuml;             /* sum of candidate-list lengths */
    int32_t *paths;             /* device [W, max_path] or NULL */
    int32_t *path_len;          /* device [W] or NULL */
    unsigned long long *counters; /* device [GG_CNT_SLOTS] */
    void *scratch;              /* device, gg_walk_scratch_bytes(max_cand) */
    int64_t scratch_bytes;
    unsigned int *work_counter; /* device, 1 word, zeroed by the call */
    /* optional per-pass precomputation (identical results, far less traffic) */
    const float *edge_score;    /* device [nnz] all_score[u, adj[e]] for walk-CSR entries of nodes with
                                   degree >= hub_threshold (gg_hub_scores); NULL = always score on demand */
    const double *root_q;       /* device: per root, the normalised CDF of its root step (gg_root_cdf);
                                   NULL = compute the root step per walk */
    const int64_t *rq_ptr;      /* device [R+1] offsets into root_q (prefix of the roots' walk-CSR degrees) */
    int32_t hub_threshold;
    int32_t no_tma;             /* 1 = enumerate hub lists with plain loads instead of cp.async.bulk staging (A/B measurement;
                                   identical results) */
    const int32_t *walk_slot;   /* optional device [W]: root slot of every walk (saves a binary search per walk) */
    /* optional depth-1 reuse (GG_RNG_PHILOX, needs root_q + walk_slot): the walks of a root that pick the same
       depth-1 child share one candidate list.  gg_walk_sample first runs the root step of every walk and counts
       the walks per (root, child) pair, builds the candidate ids and the un-normalised canonical CDF (plus its
       total) of every pair picked by at least two walks, and the walk kernel then only inverts it (pairs picked
       once are handled inside their walk).  Indexing: pair (root slot k, i-th neighbour) <-> pos = rq_ptr[k] + i,
       pos < s1_nq. */
    int64_t s1_nq;              /* = rq_ptr[R] */
    const int32_t *s1_slot;     /* device [s1_nq] root slot of every pair */
    const int64_t *s1_ptr;      /* device [s1_nq+1] exclusive prefix of (degree(child) + 1): pool offsets */
    int32_t *s1_cnt;            /* device [s1_nq] scratch: walks per pair */
    int32_t *s1_n;              /* device [s1_nq] out: list length of a pair (0 = void) */
    double *s1_q;               /* device pool [s1_ptr[s1_nq]]: per pair c[0..n-1] = carry + scan(e/S), c[n] = total */
    int32_t *s1_ids;            /* device pool: candidate ids */
    int32_t *first_idx;         /* device [W] scratch: root-step choice of every walk */
    const int32_t *s1_order;    /* optional device [s1_nq]: the (root, child) pairs sorted by decreasing child degree; with it
                                 * step1_cdf_kernel pulls pairs from a queue (largest lists first) instead of striding */
    const int32_t *walk_order;  /* optional device [W]: the order in which walk_kernel starts the walks (a permutation of
                                 * 0..W-1, e.g. expensive roots first); results do not depend on it (Philox mode) */
    /* optional level-synchronous steps (needs the depth-1 reuse): all unfinished walks take step s together -- one
       kernel enumerates the candidate lists, one gathers the candidates' rows and draws -- for steps 1..flat_steps;
       the persistent kernel finishes the walks that are still alive after that.  Identical results. */
    void *flat_buf;             /* device scratch of gg_walk_flat_bytes(n_walks, hub_threshold, flat_steps) bytes */
    int64_t flat_bytes;
    int32_t flat_steps;         /* 0 = off (persistent kernel only); <= 14 */
    int32_t flat_reserved;
} gg_walk_desc;

/* all_score[u, v] = e_u.e_v + b_v (generator.py:21) for every walk-CSR entry (u -> v) of the listed hub
 * tiles: tile t covers entries [tile_begin[t], min(tile_begin[t] + tile_edges, indptr[tile_node[t]+1])).
 * Must be re-run whenever the generator's embeddings change (i.e. once per sampling pass). */
int gg_hub_scores(int64_t n_tiles, const int32_t *tile_node, const int64_t *tile_begin, int32_t tile_edges,
                  const int64_t *indptr, const int32_t *adj, const float *emb, const float *bias, int32_t ld,
                  float *edge_score, void *stream);
/* Root-step softmax + CDF, once per root per pass: root_q[rq_ptr[k] + i] = cdf_i / cdf_last over the
 * candidates tree[root][1:] (graph_gan.py:250,260-262).  Uses d->{roots, n_roots, indptr, adj, emb, bias,
 * ld, rq_ptr, edge_score, hub_threshold}.  root_sc: device float scratch of rq_ptr[R] entries. */
int gg_root_cdf(const gg_walk_desc *d, float *root_sc, double *root_q, void *stream);

int gg_walk_scratch_bytes(int32_t max_cand, int64_t *bytes);
int gg_walk_flat_bytes(int64_t n_walks, int32_t hub_threshold, int32_t flat_steps, int64_t *bytes);
int gg_walk_sample(const gg_walk_desc *d, void *stream);

/* Per root: find the first voiding walk (graph_gan.py:252-257 returns None for the WHOLE
 * root), blank the walks after it, set root_ok ("neg is not None and len(pos) != 0",
 * graph_gan.py:192), apply the father-removal bits of the surviving D walks, and reduce the
 * counters. */
int gg_walk_finalize(int64_t n_roots, const int64_t *walk_ptr, int32_t for_d, int32_t *samples,
                     int32_t *status, const int32_t *first_edge, int32_t *wsteps, int32_t *wsuml,
                     int32_t *path_len, uint32_t *d1_bits, int32_t *root_ok,
                     unsigned long long *counters, void *stream);

/* prepare_data_for_d's output rows (graph_gan.py:192-201): for every accepted root, in batch
 * order: [i]*k + [i]*k | pos + neg | 1*k + 0*k.  row_ptr: device [R+1] scratch/out (exclusive
 * scan of 2*len(pos) over accepted roots); n_rows_out: device int64. */
int gg_emit_d_rows(int64_t n_roots, const int32_t *roots, const int64_t *walk_ptr,
                   const int64_t *pos_indptr, const int32_t *pos_flat, const int32_t *root_ok,
                   const int32_t *samples, int64_t *row_ptr, int32_t *center, int32_t *neighbor,
                   int32_t *label, int64_t *n_rows_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Tree construction: GraphGAN.construct_trees (graph_gan.py:84-108) for a batch of roots.  A tree is one bit per
 * walk-CSR entry: bit e of row k is set iff adj[e] is a child of the source node of entry e in the tree of
 * roots[k] (first discoverer in the reference's FIFO / adjacency order).  tree_bits: device [R, tree_words],
 * tree_words = gg_tree_words(nnz) (= ceil(nnz / 32) + 1); rows are zeroed by the call.  The walk never needs a
 * father pointer (it only descends: the father is the previous node); gg_tree_parent expands rows into the
 * parent-array form (parent[root] = parent[unreachable] = -1) for tests and host-side consumers.
 * ------------------------------------------------------------------------------------------ */
int gg_tree_words(int64_t nnz, int64_t *words);
int gg_bfs_scratch_bytes(int64_t n_node, int64_t nnz, int64_t *bytes);
int gg_bfs_build(int64_t n_node, int64_t nnz, const int64_t *indptr, const int32_t *adj, int64_t n_roots,
                 const int32_t *roots, uint32_t *tree_bits, int64_t tree_words, void *scratch,
                 int64_t scratch_bytes, void *stream);
/* Direction-optimising form.  rev (device [nnz], gg_reverse_entries; static per graph) maps the entry (u -> v) to the
 * entry (v -> u) and lets a level run bottom-up: every undiscovered node picks the visited neighbour with the smallest
 * queue position, the tree row then yields the new nodes in FIFO order (csrc/bfs.cu).  A level runs bottom-up when
 * (adjacency entries of the undiscovered nodes + 4 * frontier nodes) < bottom_up_ratio * (adjacency entries of the
 * frontier); bottom_up_ratio < 0: library default, 0: never.  rev == NULL: plain gg_bfs_build.  The trees are the
 * same bit for bit in every mode.  gg_reverse_entries sets *n_missing (device) to the number of entries without a
 * reverse (rev = -1 there): a CSR with n_missing != 0 is not symmetric and must be built with rev == NULL. */
#define GG_BFS_NO_SORTED_BOTTOM_UP 1   /* flags: disable the small sort-based bottom-up levels (tests / A-B) */
int gg_reverse_entries(int64_t n_node, int64_t nnz, const int64_t *indptr, const int32_t *adj, int32_t *rev,
                       int32_t *n_missing, void *stream);
int gg_bfs_build_ex(int64_t n_node, int64_t nnz, const int64_t *indptr, const int32_t *adj, const int32_t *rev,
                    int64_t n_roots, const int32_t *roots, uint32_t *tree_bits, int64_t tree_words, void *scratch,
                    int64_t scratch_bytes, float bottom_up_ratio, int32_t flags, void *stream);
int gg_tree_parent(int64_t n_node, const int64_t *indptr, const int32_t *adj, int64_t n_roots,
                   const int32_t *roots, const uint32_t *tree_bits, int64_t tree_words, int32_t *parent,
                   void *stream);

/* ------------------------------------------------------------------------------------------
 * K2: pair scoring.  score_k = e_{i_k}.e_{j_k} + b_{j_k} (discriminator.py:21-24 /
 * generator.py:22-25).
 * ------------------------------------------------------------------------------------------ */
/* discriminator.reward (discriminator.py:33-34): log(1 + exp(clip(score, -10, 10))) */
int gg_pair_reward(int64_t n_pairs, const int32_t *node_id, const int32_t *node_neighbor_id,
                   const float *emb, const float *bias, int32_t ld, float *reward, void *stream);
/* generator.all_score rows (generator.py:21) for small N only (tests / source compat) */
int gg_all_score(int64_t n_node, const float *emb, const float *bias, int32_t ld, float *out,
                 void *stream);

/* Sparse gradient of one mini-batch (<= GG_MAX_BATCH pairs), duplicates summed in pair order
 * (TF1.8 AdamOptimizer._apply_sparse_duplicate_indices: unique + segment_sum):
 *   mode 0 = discriminator loss (discriminator.py:26-30), aux = label
 *   mode 1 = generator loss     (generator.py:26-29),     aux = reward
 * outputs: n_unique (device int32), uniq_ids[2B], grad_rows[2B, ld], grad_bias[2B],
 * row_slot: device [N] int32 map, must be all -1 on entry; set for touched rows.
 * batch_total: size of the whole mini-batch when n_pairs is one rank's slice of it (the generator
 * loss is a MEAN over the batch, generator.py:28); 0 = n_pairs. */
#define GG_MAX_BATCH 1024
int gg_pair_grad(int32_t mode, int32_t n_pairs, int32_t batch_total, const int32_t *node_id,
                 const int32_t *node_neighbor_id, const float *aux, const float *emb,
                 const float *bias, int32_t ld, float lambda, int32_t *n_unique, int32_t *uniq_ids,
                 float *grad_rows, float *grad_bias, int32_t *row_slot, void *stream);

/* Data-parallel step: merge the compact gradients of `world` ranks (each laid out as one buffer of
 * gg_grad_buf_floats(cap, ld) floats: rows[cap, ld] | bias[cap] | ids[cap] (int32 bits) | n_unique) into
 * the final unique/summed form, rank-major entry order -- every rank computes the identical result
 * from the all-gathered buffers, so replicas stay bit-identical.  Clears and re-sets row_slot. */
int64_t gg_grad_buf_floats(int32_t cap, int32_t ld);
int gg_grad_merge(int32_t world, int32_t cap, int32_t ld, const float *gathered, int32_t *n_unique,
                  int32_t *uniq_ids, float *grad_rows, float *grad_bias, int32_t *row_slot, void *stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel optimizer step with its collective inside the library (csrc/comm.cu).  N replicas of the
 * per-batch loops of graph_gan.py:149-157 / 168-176: every rank computes the gradient of ITS rows of the mini-batch
 * (gg_pair_grad with batch_total), ONE ncclAllGather exchanges the compact gradients over NVLink, every rank merges
 * them in rank-major order (gg_grad_merge) and applies the same Adam sweep -- replicas stay bit-identical.
 *   gg_comm_unique_id : ncclGetUniqueId (128 bytes; rank 0 creates it, the caller broadcasts it out of band)
 *   gg_comm_init      : ncclCommInitRank on the CURRENT device -> opaque handle
 *   gg_dp_step        : the whole step on `stream` (node_id / node_neighbor_id / aux: the WHOLE batch, device, identical
 *                       on all ranks; local_buf: gg_grad_buf_floats(cap, ld) floats; gathered_buf: world times that;
 *                       cap >= 2 * ceil(n_pairs / world))
 *   gg_dp_train_steps : gg_dp_step for every start of a (host) shuffled start list, enqueued from C; beta powers as in
 *                       gg_train_steps
 * NCCL is dlopen-ed at first use (the process's already-loaded libnccl.so.2 if there is one).
 * ------------------------------------------------------------------------------------------ */
int gg_comm_unique_id(void *id128);
int gg_comm_init(const void *id128, int32_t rank, int32_t world, void **comm_out);
int gg_comm_destroy(void *comm);
int gg_comm_info(void *comm, int32_t *rank, int32_t *world, int32_t *nccl_version, uint64_t *collectives);
/* Peer-memory transport for the same step (optional; NVLink P2P through CUDA IPC, one process per GPU): every rank
 * creates an exchange buffer (gg_comm_p2p_export -> 64-byte cudaIpcMemHandle_t), the caller all-gathers the handles,
 * gg_comm_p2p_connect maps the peers' buffers.  With it, gg_dp_step runs the gradient AND the exchange in one kernel
 * (stores into every peer's buffer + a release flag) and the merge kernel waits on the flags: no collective call per step.
 * capacity_floats >= world * gg_grad_buf_floats(cap, ld).  gg_comm_use_p2p switches between the two transports. */
int gg_comm_p2p_export(void *comm, int64_t capacity_floats, void *handle64);
int gg_comm_p2p_connect(void *comm, const void *all_handles);
int gg_comm_use_p2p(void *comm, int32_t on);
int gg_dp_step(void *comm, int32_t mode, int32_t n_pairs, const int32_t *node_id, const int32_t *node_neighbor_id,
               const float *aux, int64_t n_node, int32_t ld, float *emb, float *m_emb, float *v_emb, float *bias,
               float *m_bias, float *v_bias, float lambda, float *local_buf, float *gathered_buf, int32_t cap,
               int32_t *n_unique, int32_t *uniq_ids, float *grad_rows, float *grad_bias, int32_t *row_slot,
               float lr_t, float beta1, float beta2, float eps, void *stream);
int gg_dp_train_steps(void *comm, int32_t mode, int64_t n_rows, const int64_t *start_list, int64_t n_starts,
                      int32_t batch_size, const int32_t *node_id, const int32_t *node_neighbor_id, const float *aux,
                      int64_t n_node, int32_t ld, float *emb, float *m_emb, float *v_emb, float *bias, float *m_bias,
                      float *v_bias, float lambda, float *local_buf, float *gathered_buf, int32_t cap,
                      int32_t *n_unique, int32_t *uniq_ids, float *grad_rows, float *grad_bias, int32_t *row_slot,
                      float lr, float beta1, float beta2, float eps, float *beta1_power, float *beta2_power,
                      void *stream);

/* K3: TF1.8 AdamOptimizer sparse apply == dense decay (generator.py:30-31,
 * discriminator.py:31-32): m <- b1*m (+ (1-b1) g on touched rows), v likewise, then for ALL
 * rows var -= lr_t * m / (sqrt(v) + eps).  Resets row_slot to -1. */
/* Selects the kernel behind gg_adam_apply (identical results): "ldg" per-thread loads (default), "tma" / "tma256x2" /
 * "tma512x3" cp.async.bulk pipeline with a CTA barrier per tile, "ws16" / "ws8" warp-specialised cp.async.bulk pipeline.
 * The environment variable GG_ADAM_PATH sets the initial choice. */
int gg_set_adam_path(const char *name);
int gg_adam_apply(int64_t n_node, int32_t ld, float *emb, float *m_emb, float *v_emb, float *bias,
                  float *m_bias, float *v_bias, const int32_t *n_unique, const int32_t *uniq_ids,
                  const float *grad_rows, const float *grad_bias, int32_t *row_slot, float lr_t,
                  float beta1, float beta2, float eps, void *stream);

/* The inner training loop of graph_gan.py:149-157 / 168-176: for each start in start_list (host array, already
 * shuffled by the caller): one optimizer step on rows [start, min(start + batch_size, n_rows)) of the device
 * arrays node_id / node_neighbor_id / aux -- i.e. gg_pair_grad + gg_adam_apply per step, enqueued from C so that
 * the per-step cost is two kernel launches, not a round trip through the host language.  beta1_power/beta2_power:
 * host in/out, the AdamOptimizer's beta^t accumulators (fp32, multiplied once per step like TF's _finish). */
int gg_train_steps(int32_t mode, int64_t n_rows, const int64_t *start_list, int64_t n_starts, int32_t batch_size,
                   const int32_t *node_id, const int32_t *node_neighbor_id, const float *aux, int64_t n_node, int32_t ld,
                   float *emb, float *m_emb, float *v_emb, float *bias, float *m_bias, float *v_bias, float lambda,
                   int32_t *n_unique, int32_t *uniq_ids, float *grad_rows, float *grad_bias, int32_t *row_slot, float lr,
                   float beta1, float beta2, float eps, float *beta1_power, float *beta2_power, void *stream);

/* The same loop as gg_train_steps in ONE cooperative launch (persistent kernel; a ready flag and an arrival
 * counter order the gradient and the Adam sweep of every step).  start_list_dev is a DEVICE array; sync_words
 * is a device scratch of eight uint64 (zeroed by the call; [0..1] flag and counter, [2..5] diagnostic: CTA 0's
 * clock cycles in gradient / sweep / wait, and the number of steps).  Bit-identical results. */
int gg_train_loop(int32_t mode, int64_t n_rows, const int64_t *start_list_dev, int64_t n_starts, int32_t batch_size,
                  const int32_t *node_id, const int32_t *node_neighbor_id, const float *aux, int64_t n_node, int32_t ld,
                  float *emb, float *m_emb, float *v_emb, float *bias, float *m_bias, float *v_bias, float lambda,
                  int32_t *n_unique, int32_t *uniq_ids, float *grad_rows, float *grad_bias, int32_t *row_slot, float lr,
                  float beta1, float beta2, float eps, float *beta1_power, float *beta2_power, uint64_t *sync_words,
                  void *stream);

/* The same loop with ONE inter-CTA barrier per step (csrc/steps.cu: train_fused_kernel): every CTA rebuilds the
 * mini-batch's forward pass and entry lists, sweeps the rows it owns and accumulates their gradient on the fly;
 * parameters ping-pong between (emb, bias) and the caller's second buffers (emb2 [N, ld], bias2 [N]); the result is
 * always left in (emb, bias).  For graphs whose (E, m, v) stay in L2; returns an error when the per-CTA row table
 * does not fit in shared memory.  grad_rows / uniq_ids / row_slot are not used.  Bit-identical to gg_train_steps. */
int gg_train_fused(int32_t mode, int64_t n_rows, const int64_t *start_list_dev, int64_t n_starts, int32_t batch_size,
                   const int32_t *node_id, const int32_t *node_neighbor_id, const float *aux, int64_t n_node, int32_t ld,
                   float *emb, float *m_emb, float *v_emb, float *bias, float *m_bias, float *v_bias, float *emb2,
                   float *bias2, float lambda, float lr, float beta1, float beta2, float eps, float *beta1_power,
                   float *beta2_power, uint64_t *sync_words, void *stream);

/* ------------------------------------------------------------------------------------------
 * End-of-epoch dump and quality line on the device (csrc/eval.cu).  Replaces the text round trip of
 * write_embeddings_to_file (graph_gan.py:293-306) -> utils.read_embeddings (utils.py:57-67) ->
 * LinkPredictEval.eval_link_prediction (src/evaluation/link_prediction.py:19-38).
 *   gg_pair_dot_f64  : out[k] = float64 dot of rows node_id[k], node_neighbor_id[k] (np.dot on the re-read rows)
 *   gg_link_pred_acc : out2[0] = accuracy of (score >= np.median(score)) against labels [1]*(n/2) + [0]*(n - n/2),
 *                      out2[1] = the median; score / out2: device float64
 *   gg_unpad_rows    : dense [N, n_emb] fp32 copy of the padded [N, ld] rows (payload of the binary dump)
 * ------------------------------------------------------------------------------------------ */
int gg_pair_dot_f64(int64_t n_pairs, const int32_t *node_id, const int32_t *node_neighbor_id, const float *emb,
                    int32_t ld, double *out, void *stream);
int gg_link_pred_acc(int64_t n, const double *score, double *out2, void *stream);
int gg_unpad_rows(int64_t n_node, int32_t ld, int32_t n_emb, const float *emb, float *out, void *stream);

/* get_node_pairs_from_path (graph_gan.py:272-291) for a batch of recorded paths.
 * pair_ptr: device [W+1] (out, exclusive scan of per-path pair counts). */
int gg_window_pairs(int64_t n_walks, const int32_t *paths, const int32_t *path_len, int32_t max_path,
                    int32_t window, int64_t *pair_ptr, int32_t *node_1, int32_t *node_2,
                    int64_t *n_pairs_out, int64_t capacity, void *stream);

#ifdef __cplusplus
}
#endif
#endif
