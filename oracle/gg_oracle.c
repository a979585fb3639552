/*
 * gg_oracle.c -- TEST INFRASTRUCTURE ONLY.  See gg_oracle.h for the contract.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off matters: every multiply-add that is meant to be fused is written as
 * fmaf(); nothing else may be fused, or the CUDA kernels (which use __fmaf_rn/__fadd_rn
 * explicitly) would no longer execute the same operation sequence.
 */
#include "gg_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- Philox4x32-10 */
/* Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11); constants from
 * the paper. Checked against the Random123 known-answer vectors in tests/test_oracle.py. */
void ggo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Same 53-bit construction as MT19937's random_sample (what RandomState.choice draws at
 * graph_gan.py:262 and np.random.rand() at :189/:209). */
double ggo_u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

/* ---------------------------------------------------------------- canonical exp, x <= 0 */
static inline float bits_to_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f_to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

float ggo_exp(float x) {
    if (x < -86.0f) return 0.0f;
    const float MAGIC = 12582912.0f; /* 1.5 * 2^23: fmaf lands on an integer, ties-to-even */
    float t = fmaf(x, 1.44269504088896341f, MAGIC);
    float n = t - MAGIC;
    float r = fmaf(n, -0.693359375f, x);        /* Cody-Waite split of ln 2 */
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float e = fmaf(p, r2, r);
    e = e + 1.0f;
    int ni = (int)n; /* exact integer in [-125, 0] */
    return bits_to_f(f_to_bits(e) + ((uint32_t)ni << 23));
}

/* ---------------------------------------------------------------- canonical dot */
float ggo_dot(const float *a, const float *b, int ld) {
    float acc[8];
    int nchunk = ld / 4;
    for (int g = 0; g < 8; ++g) {
        float s = 0.0f;
        for (int c = g; c < nchunk; c += 8) {
            const float *pa = a + 4 * c, *pb = b + 4 * c;
            s = fmaf(pa[0], pb[0], s);
            s = fmaf(pa[1], pb[1], s);
            s = fmaf(pa[2], pb[2], s);
            s = fmaf(pa[3], pb[3], s);
        }
        acc[g] = s;
    }
    float t[8];
    for (int g = 0; g < 8; ++g) t[g] = acc[g] + acc[g ^ 4];
    for (int g = 0; g < 8; ++g) acc[g] = t[g] + t[g ^ 2];
    return acc[0] + acc[1];
}

/* ---------------------------------------------------------------- softmax + choice */
static inline void ks_scan32(double *x) {
    for (int off = 1; off < 32; off <<= 1)
        for (int l = 31; l >= off; --l) x[l] = x[l] + x[l - off];
}

int ggo_choose(float *sc, int n, double u) {
    float m = sc[0];
    for (int i = 1; i < n; ++i) m = fmaxf(m, sc[i]);
    int ntile = (n + 31) / 32;
    float S = 0.0f;
    for (int t = 0; t < ntile; ++t) {
        float v[32], w[32];
        for (int l = 0; l < 32; ++l) {
            int i = 32 * t + l;
            v[l] = 0.0f;
            if (i < n) { v[l] = ggo_exp(sc[i] - m); sc[i] = v[l]; }
        }
        for (int off = 16; off >= 1; off >>= 1) {
            for (int l = 0; l < 32; ++l) w[l] = v[l] + v[l ^ off];
            memcpy(v, w, sizeof(v));
        }
        S = (t == 0) ? v[0] : S + v[0];
    }
    double total = 0.0;
    for (int t = 0; t < ntile; ++t) {
        double x[32];
        for (int l = 0; l < 32; ++l) {
            int i = 32 * t + l;
            x[l] = (i < n) ? (double)(sc[i] / S) : 0.0;
        }
        ks_scan32(x);
        total = total + x[31];
    }
    double carry = 0.0;
    for (int t = 0; t < ntile; ++t) {
        double x[32];
        for (int l = 0; l < 32; ++l) {
            int i = 32 * t + l;
            x[l] = (i < n) ? (double)(sc[i] / S) : 0.0;
        }
        ks_scan32(x);
        for (int l = 0; l < 32; ++l) {
            int i = 32 * t + l;
            if (i >= n) break;
            double q = (carry + x[l]) / total;
            if (q > u) return i;
        }
        carry = carry + x[31];
    }
    return n - 1;
}

/* ---------------------------------------------------------------- BFS parent array */
int64_t ggo_bfs_parent(int64_t n, const int64_t *indptr, const int32_t *adj, int32_t root,
                       int32_t *parent, int32_t *queue) {
    /* graph_gan.py:93-107: FIFO queue, neighbours in adjacency order, first discoverer is
     * the father.  The root's own slot and unreachable nodes stay -1. */
    for (int64_t i = 0; i < n; ++i) parent[i] = -1;
    int64_t head = 0, tail = 0;
    queue[tail++] = root;
    parent[root] = root; /* temporarily "used" */
    while (head < tail) {
        int32_t cur = queue[head++];
        for (int64_t e = indptr[cur]; e < indptr[cur + 1]; ++e) {
            int32_t v = adj[e];
            if (parent[v] == -1) { parent[v] = cur; queue[tail++] = v; }
        }
    }
    parent[root] = -1;
    return tail;
}

/* ---------------------------------------------------------------- the walk pass */
int ggo_walk_pass(const ggo_walk_args *a) {
    const int64_t N = a->n_node;
    int64_t maxdeg = 0;
    for (int64_t i = 0; i < N; ++i) {
        int64_t dg = a->indptr[i + 1] - a->indptr[i];
        if (dg > maxdeg) maxdeg = dg;
    }
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(maxdeg + 2));
    int64_t *cedge = (int64_t *)malloc(sizeof(int64_t) * (size_t)(maxdeg + 2));
    float *sc = (float *)malloc(sizeof(float) * (size_t)(maxdeg + 2));
    if (!cand || !cedge || !sc) return -1;
    int64_t cursor = 0, tot_steps = 0, tot_suml = 0, overflow = 0, max_l = 0;
    const uint32_t key[2] = {(uint32_t)(a->seed & 0xffffffffu), (uint32_t)(a->seed >> 32)};
    int rc = 0;

    for (int64_t ri = 0; ri < a->n_roots && rc == 0; ++ri) {
        const int32_t root = a->roots[ri];
        const int32_t *par = a->parent + (size_t)ri * (size_t)N;
        const int64_t w0 = a->walk_ptr[ri], nw = a->walk_ptr[ri + 1] - w0;
        for (int64_t k = 0; k < nw; ++k) {
            a->status[w0 + k] = GGO_NOTRUN; a->samples[w0 + k] = -1; a->first_edge[w0 + k] = -1;
            a->wsteps[w0 + k] = 0; a->wsuml[w0 + k] = 0;
            if (a->path_len) a->path_len[w0 + k] = 0;
        }
        /* graph_gan.py:189 / :209 -- one uniform per root, always drawn */
        double ur;
        if (a->rng_mode == GGO_RNG_STREAM) {
            if (cursor >= a->n_stream) { rc = -2; break; }
            ur = a->stream[cursor++];
        } else {
            uint32_t ctr[4] = {(uint32_t)root, 0xffffffffu, 0u, a->pass_tag}, o[4];
            ggo_philox4x32_10(ctr, key, o);
            ur = ggo_u53(o[0], o[1]);
        }
        if (!(ur < a->update_ratio)) {
            for (int64_t k = 0; k < nw; ++k) a->status[w0 + k] = GGO_SKIPPED;
            a->root_ok[ri] = 0;
            continue;
        }
        int ok = 1;
        for (int64_t k = 0; k < nw && ok; ++k) {
            const int64_t w = w0 + k;
            int32_t cur = root, prev = -1, step = 0;
            int64_t fedge = -1;
            int32_t plen = 0;
            int32_t *prow = (a->max_path > 0) ? a->paths + (size_t)w * (size_t)a->max_path : 0;
            if (prow) { if (plen < a->max_path) prow[plen] = cur; }
            plen++;
            for (;;) {
                /* graph_gan.py:250-259 -- candidate list */
                int inc_father = step > 0;
                if (a->for_d && step == 1) inc_father = 0;      /* :258-259 root removed */
                if (!a->for_d && step == 1 && fedge >= 0 &&
                    ((a->d1_bits[fedge >> 5] >> (fedge & 31)) & 1u)) inc_father = 0; /* mutated tree */
                int32_t n = 0;
                if (inc_father) { cand[0] = prev; cedge[0] = -1; n = 1; }
                for (int64_t e = a->indptr[cur]; e < a->indptr[cur + 1]; ++e) {
                    int32_t v = a->adj[e];
                    if (par[v] == cur) { cand[n] = v; cedge[n] = e; ++n; }
                }
                if (n == 0) { /* :252-253 and :255-257 */
                    a->status[w] = GGO_VOID; ok = 0; break;
                }
                const float *ecur = a->emb + (size_t)cur * (size_t)a->ld;
                for (int32_t i = 0; i < n; ++i)
                    sc[i] = ggo_dot(ecur, a->emb + (size_t)cand[i] * (size_t)a->ld, a->ld) + a->bias[cand[i]];
                double u;
                if (a->rng_mode == GGO_RNG_STREAM) {
                    if (cursor >= a->n_stream) { rc = -2; ok = 0; break; }
                    u = a->stream[cursor++];
                } else {
                    uint32_t ctr[4] = {(uint32_t)root, (uint32_t)k, (uint32_t)step, a->pass_tag}, o[4];
                    ggo_philox4x32_10(ctr, key, o);
                    u = ggo_u53(o[0], o[1]);
                }
                int idx = ggo_choose(sc, n, u);
                int32_t nxt = cand[idx];
                if (step == 0) fedge = cedge[idx];
                if (prow && plen < a->max_path) prow[plen] = nxt;
                plen++;
                a->wsteps[w] += 1; a->wsuml[w] += n;
                tot_steps += 1; tot_suml += n; if (n > max_l) max_l = n;
                if (inc_father && idx == 0) { /* :264-266 next == previous */
                    a->samples[w] = cur; a->status[w] = GGO_DONE; break;
                }
                prev = cur; cur = nxt; ++step;
            }
            a->first_edge[w] = (int32_t)fedge;
            if (a->status[w] == GGO_DONE) {
                if (a->path_len) a->path_len[w] = plen;
                if (a->max_path > 0 && plen > a->max_path) overflow++;
                if (a->for_d && fedge >= 0) a->d1_bits[fedge >> 5] |= (1u << (fedge & 31));
            }
        }
        a->root_ok[ri] = (ok && nw > 0) ? 1 : 0;
    }
    if (a->counters) {
        a->counters[0] = tot_steps; a->counters[1] = tot_suml; a->counters[2] = cursor;
        a->counters[3] = overflow; a->counters[4] = max_l;
    }
    free(cand); free(cedge); free(sc);
    return rc;
}
