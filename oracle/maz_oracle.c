/*
 * maz_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C (C11, gcc) CPU restatement of MAZero's batched sampled-MCTS tree engine, written from the
 * behavioural spec, each function citing the reference file:line it follows.
 * It exists to check the CUDA path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it.  The product path (mazero_b200/) must never call into this file.
 *
 * PARITY PINNING: the reference's own tests pin shapes only (unit_test_mcts.py:203-210), so this
 * restatement is pinned instead against (i) the reference's own C++ compiled by oracle/ref_shim.cpp
 * (tests/test_oracle_vs_ref.py, bit-exact on every output, when oracle/_ref/libmazref.so is present),
 * and (ii) committed golden vectors generated from that compiled reference (tests/golden/, made by
 * tests/golden/make_golden.py).
 *
 * Third-party arithmetic restated here (not under /root/reference): libstdc++ 13.3 <random>
 *   std::mt19937                       (bits/random.h / random.tcc mersenne_twister_engine)
 *   std::discrete_distribution<int>    (bits/random.tcc:2657-2714)
 *   std::generate_canonical<double,53> (bits/random.tcc:3346-3381)
 * and glibc libm log/sqrt/ceil in their double overloads (what unqualified log(float)/sqrt(int)/
 * ceil(float) resolve to with only <cmath> included: cnode.cpp:313-314, utils.cpp:31).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no FMA contraction; the reference build has none).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_abi.h"

static _Thread_local char g_err[256];
const char *mazo_last_error(void) { return g_err; }
static int fail(const char *msg)
{
    snprintf(g_err, sizeof g_err, "%s", msg);
    return 1;
}

/* ------------------------------------------------------------------ MT19937 (std::mt19937) ------ */
typedef struct {
    uint32_t s[624];
    int p;
} mt_t;

static void mt_seed(mt_t *g, uint32_t seed)
{
    g->s[0] = seed;
    for (int i = 1; i < 624; ++i) g->s[i] = 1812433253u * (g->s[i - 1] ^ (g->s[i - 1] >> 30)) + (uint32_t)i;
    g->p = 624;
}

static uint32_t mt_next(mt_t *g)
{
    if (g->p >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->s[i] & 0x80000000u) | (g->s[(i + 1) % 624] & 0x7fffffffu);
            g->s[i] = g->s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->p = 0;
    }
    uint32_t z = g->s[g->p++];
    z ^= z >> 11;
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= z >> 18;
    return z;
}

/* std::generate_canonical<double,53>(mt19937): two 32-bit draws, low word first
 * (random.tcc:3362-3378): sum = x1 + x2*2^32 (rounded to double), / 2^64, clamp below 1. */
static double mt_canonical(mt_t *g)
{
    double sum = 0.0, tmp = 1.0;
    for (int k = 0; k < 2; ++k) {
        sum += (double)mt_next(g) * tmp;
        tmp *= 4294967296.0;
    }
    double r = sum / tmp;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

/* ------------------------------------------------------- sorted float multiset (std::multiset) -- */
typedef struct {
    float *v;
    int n, cap;
} fset;

static void fset_insert(fset *s, float x)
{
    if (s->n == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 4;
        s->v = (float *)realloc(s->v, sizeof(float) * (size_t)s->cap);
    }
    int i = s->n;
    while (i > 0 && s->v[i - 1] > x) {
        s->v[i] = s->v[i - 1];
        --i;
    }
    s->v[i] = x;
    s->n++;
}
static void fset_erase_at(fset *s, int i)
{
    memmove(s->v + i, s->v + i + 1, sizeof(float) * (size_t)(s->n - i - 1));
    s->n--;
}
static int fset_find(const fset *s, float x)
{
    for (int i = 0; i < s->n; ++i)
        if (s->v[i] == x) return i;
    return -1;
}

/* ------------------------------------------------ OS(lambda) statistic: SubTreeValueSet --------- */
/* utils.h:18-40, utils.cpp:20-77 */
typedef struct {
    float wsum, wtot;
    int nd;       /* number of depth slots */
    int cap;
    fset *big, *small;
    int *count;
    float *lam_pow;
} vset;

static int vset_update(vset *s, float key, int depth, float rho, float lam)
{
    if (s->nd <= depth) { /* utils.cpp:21-27: exactly one slot is appended */
        if (s->nd == s->cap) {
            int nc = s->cap ? 2 * s->cap : 4;
            s->big = (fset *)realloc(s->big, sizeof(fset) * (size_t)nc);
            s->small = (fset *)realloc(s->small, sizeof(fset) * (size_t)nc);
            s->count = (int *)realloc(s->count, sizeof(int) * (size_t)nc);
            s->lam_pow = (float *)realloc(s->lam_pow, sizeof(float) * (size_t)nc);
            s->cap = nc;
        }
        memset(&s->big[s->nd], 0, sizeof(fset));
        memset(&s->small[s->nd], 0, sizeof(fset));
        s->count[s->nd] = 0;
        s->lam_pow[s->nd] = (depth == 0) ? 1.0f : s->lam_pow[depth - 1] * lam;
        s->nd++;
    }
    if (depth >= s->nd) return fail("SubTreeValueSet::update: depth slot skipped (UB in the reference)");
    s->count[depth] += 1;

    fset *B = &s->big[depth], *S = &s->small[depth];
    const float lp = s->lam_pow[depth];
    int cur = B->n;
    /* utils.cpp:31: count(int) * (1 - quantile)(float) -> float; ceil in double */
    float prod = (float)s->count[depth] * (1 - rho);
    int lim = (int)ceil((double)prod);
    if (lim < 1) lim = 1;

    if (cur == lim) { /* utils.cpp:33-48 */
        float m = B->v[0];
        if (key < m) {
            fset_insert(S, key);
        } else {
            fset_insert(S, m);
            s->wsum -= lp * m;
            s->wtot -= lp;
            fset_erase_at(B, 0);
            fset_insert(B, key);
            s->wtot += lp;
            s->wsum += lp * key;
        }
    } else { /* utils.cpp:49-70 */
        if (cur + 1 != lim) return fail("SubTreeValueSet::update: cur_size+1!=size_lim.");
        if (S->n == 0) {
            fset_insert(B, key);
            s->wtot += lp;
            s->wsum += lp * key;
        } else {
            float M = S->v[S->n - 1];
            if (key > M) {
                fset_insert(B, key);
                s->wtot += lp;
                s->wsum += lp * key;
            } else {
                fset_insert(B, M);
                s->wtot += lp;
                s->wsum += lp * M;
                fset_erase_at(S, S->n - 1);
                fset_insert(S, key);
            }
        }
    }
    return 0;
}

static void vset_free(vset *s)
{
    for (int d = 0; d < s->nd; ++d) {
        free(s->big[d].v);
        free(s->small[d].v);
    }
    free(s->big);
    free(s->small);
    free(s->count);
    free(s->lam_pow);
    memset(s, 0, sizeof *s);
}

/* ---------------------------------------------------------------- nodes / trees ----------------- */
/* cnode.h:11-45 */
typedef struct {
    int visit, nchild, hidx, child0;
    float reward, pred_value, prior, pred_prob, beta, beta_hat;
    int is_root;
    vset sub;
    int *act; /* action of the edge leading to this node (agent_num ints), NULL for the root */
} node_t;

/* cnode.h:59-105 */
typedef struct {
    mt_t gen;
    int tot_nodes;
    node_t *pool;
    int *act_pool;
    fset minmax; /* CMinMaxStats::se, utils.h:42-53 */
    int search_len, res_idx;
    int *res_action;
    int *path; /* pool indices, path[0] = root */
} tree_t;

typedef struct {
    int B, N, A, K, S, P;
    float delta_lb, rho, lam;
    tree_t *trees;
} batch_t;

static float node_value(const node_t *n) /* cnode.cpp:42-56 */
{
    if (n->nchild <= 0) return 0.0f;
    return n->sub.wsum / n->sub.wtot;
}
static float node_qsa(const node_t *n, float discount) /* cnode.cpp:58-67 */
{
    return n->reward + discount * node_value(n);
}

static float minmax_normalize(const fset *se, float delta_lb, float v) /* utils.cpp:95-103 */
{
    if (se->n == 0) return v;
    float mx = se->v[se->n - 1], mn = se->v[0];
    float delta = mx - mn;
    return (v - mn) / (delta_lb > delta ? delta_lb : delta);
}

/* cnode.cpp:224-295 */
static int tree_expand(batch_t *b, tree_t *t, int node_i, int hidx, float reward, float value, const float *probs,
                       const float *beta, int K, float eps, const float *noises)
{
    const int N = b->N, A = b->A;
    node_t *nd = &t->pool[node_i];
    nd->hidx = hidx;
    nd->reward = reward;
    nd->pred_value = value;

    /* std::discrete_distribution per agent (random.tcc:2657-2678) */
    double *cp = (double *)malloc(sizeof(double) * (size_t)N * (size_t)A);
    for (int i = 0; i < N; ++i) {
        if (A < 2) continue; /* empty _M_cp: always returns 0, consumes no randomness */
        double sum = 0.0;
        for (int a = 0; a < A; ++a) sum += (double)beta[i * A + a];
        double run = 0.0;
        for (int a = 0; a < A; ++a) {
            double p = (double)beta[i * A + a] / sum;
            run = (a == 0) ? p : run + p;
            cp[i * A + a] = run;
        }
        cp[i * A + A - 1] = 1.0;
    }

    /* cnode.cpp:251-262: K joint draws, map keyed by the wrapped signed 64-bit hash */
    int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)K);
    float *cnt = (float *)malloc(sizeof(float) * (size_t)K);
    int *acts = (int *)malloc(sizeof(int) * (size_t)K * (size_t)N);
    int *tmp = (int *)malloc(sizeof(int) * (size_t)N);
    int nk = 0;
    for (int k = 0; k < K; ++k) {
        uint64_t key = 0;
        for (int i = 0; i < N; ++i) {
            int a = 0;
            if (A >= 2) {
                double u = mt_canonical(&t->gen);
                const double *c = cp + i * A; /* std::lower_bound: first c[a] with !(c[a] < u) */
                int lo = 0, hi = A;
                while (lo < hi) {
                    int mid = lo + (hi - lo) / 2;
                    if (c[mid] < u) lo = mid + 1; else hi = mid;
                }
                a = lo;
            }
            tmp[i] = a;
            key = key * 23333ull + (uint64_t)(int64_t)a; /* C long wraps for N >= 5 */
        }
        int64_t sk = (int64_t)key;
        int pos = 0;
        while (pos < nk && keys[pos] < sk) ++pos;
        if (pos < nk && keys[pos] == sk) {
            cnt[pos] += 1.0f;
        } else {
            memmove(keys + pos + 1, keys + pos, sizeof(int64_t) * (size_t)(nk - pos));
            memmove(cnt + pos + 1, cnt + pos, sizeof(float) * (size_t)(nk - pos));
            memmove(acts + (size_t)(pos + 1) * N, acts + (size_t)pos * N, sizeof(int) * (size_t)(nk - pos) * (size_t)N);
            keys[pos] = sk;
            cnt[pos] = 1.0f;
            ++nk;
        }
        memcpy(acts + (size_t)pos * N, tmp, sizeof(int) * (size_t)N); /* action_map[key] = last sample */
    }

    nd->nchild = nk;
    nd->child0 = t->tot_nodes;
    if (t->tot_nodes + nk > b->P) {
        free(cp); free(keys); free(cnt); free(acts); free(tmp);
        return fail("node pool exhausted");
    }
    for (int c = 0; c < nk; ++c) { /* cnode.cpp:268-294, ascending key order */
        const int *a = acts + (size_t)c * N;
        float betahat_prob = cnt[c] / K;
        float beta_prob = 1.0f, pred_prob = 1.0f, prior = 1.0f;
        for (int i = 0; i < N; ++i) {
            beta_prob *= beta[i * A + a[i]];
            pred_prob *= probs[i * A + a[i]];
            if (eps > 0) {
                float p = probs[i * A + a[i]] * (1 - eps) + noises[i * A + a[i]] * eps;
                prior *= p;
            } else {
                prior *= probs[i * A + a[i]];
            }
        }
        prior = prior * betahat_prob / beta_prob;
        int slot = t->tot_nodes++;
        node_t *ch = &t->pool[slot];
        memset(ch, 0, sizeof *ch);
        ch->hidx = -1;
        ch->prior = prior;
        ch->pred_prob = pred_prob;
        ch->beta = beta_prob;
        ch->beta_hat = betahat_prob;
        ch->act = t->act_pool + (size_t)slot * N;
        memcpy(ch->act, a, sizeof(int) * (size_t)N);
    }
    free(cp); free(keys); free(cnt); free(acts); free(tmp);
    return 0;
}

/* cnode.cpp:297-335 */
static float tree_ucb(const batch_t *b, const tree_t *t, const node_t *child, float parent_q, int total, float c_base,
                      float c_init, float discount)
{
    float pb_c;
    float ratio = ((float)total + c_base + 1) / c_base;
    pb_c = (float)(log((double)ratio) + (double)c_init);
    pb_c = (float)((double)pb_c * (sqrt((double)total) / (double)(child->visit + 1)));
    float prior_score = pb_c * child->prior;
    float value_score;
    if (child->visit == 0) value_score = 0;
    else value_score = node_qsa(child, discount) - parent_q;
    value_score = minmax_normalize(&t->minmax, b->delta_lb, value_score);
    if (value_score < 0) value_score = 0;
    if (value_score > 1) value_score = 1;
    return prior_score + value_score;
}

/* cnode.cpp:337-379 */
static int tree_select_child(const batch_t *b, tree_t *t, const node_t *nd, float c_base, float c_init, float discount,
                             float parent_q, int *list)
{
    float max_score = -1000000.0f;
    const float epsilon = 0.000001f;
    int nl = 0;
    for (int c = 0; c < nd->nchild; ++c) {
        float s = tree_ucb(b, t, &t->pool[nd->child0 + c], parent_q, nd->visit - 1, c_base, c_init, discount);
        if (max_score < s) {
            max_score = s;
            nl = 0;
            list[nl++] = c;
        } else if (s >= max_score - epsilon) {
            list[nl++] = c;
        }
    }
    int ci = 0;
    if (nl > 0) ci = list[mt_next(&t->gen) % (uint32_t)nl];
    return ci;
}

/* cnode.cpp:381-413 */
static void tree_select_path(const batch_t *b, tree_t *t, float c_base, float c_init, float discount, int *list)
{
    int ni = 0;
    t->search_len = 0;
    t->path[0] = 0;
    while (t->pool[ni].nchild > 0) {
        const node_t *nd = &t->pool[ni];
        int ci;
        if (nd->is_root && nd->visit <= nd->nchild) ci = nd->visit - 1;
        else ci = tree_select_child(b, t, nd, c_base, c_init, discount, nd->pred_value, list);
        ni = nd->child0 + ci;
        t->res_action = t->pool[ni].act;
        t->search_len += 1;
        t->path[t->search_len] = ni;
    }
    t->res_idx = t->pool[t->path[t->search_len - 1]].hidx;
}

static int minmax_remove(fset *se, float v) /* utils.cpp:83-88 */
{
    int i = fset_find(se, v);
    if (i < 0) return fail("CMinMaxStats::remove: value not found");
    fset_erase_at(se, i);
    return 0;
}

/* cnode.cpp:415-450 */
static int tree_backprop(batch_t *b, tree_t *t, float value, float discount)
{
    float G = value;
    const int len = t->search_len;
    for (int i = len; i >= 0; --i) {
        node_t *nd = &t->pool[t->path[i]];
        if (i != len && i != 0) {
            const node_t *fa = &t->pool[t->path[i - 1]];
            if (minmax_remove(&t->minmax, node_qsa(nd, discount) - fa->pred_value)) return 1;
        }
        nd->visit += 1;
        if (vset_update(&nd->sub, G, len - i, b->rho, b->lam)) return 1;
        if (i != 0) {
            const node_t *fa = &t->pool[t->path[i - 1]];
            fset_insert(&t->minmax, node_qsa(nd, discount) - fa->pred_value);
        }
        G = nd->reward + discount * G;
    }
    return 0;
}

/* ------------------------------------------------------------------------ batch ABI ------------- */
void *mazo_create(int B, int N, int A, int K, int S, float delta_lb, unsigned int seed, float rho, float lam)
{
    if (B <= 0 || N <= 0 || A <= 0 || K <= 0 || S < 0) {
        fail("mazo_create: bad dimensions");
        return NULL;
    }
    batch_t *b = (batch_t *)calloc(1, sizeof *b);
    b->B = B; b->N = N; b->A = A; b->K = K; b->S = S;
    b->P = K * (S + 2); /* cnode.cpp:562 */
    b->delta_lb = delta_lb; b->rho = rho; b->lam = lam;
    b->trees = (tree_t *)calloc((size_t)B, sizeof(tree_t));
    for (int i = 0; i < B; ++i) {
        tree_t *t = &b->trees[i];
        mt_seed(&t->gen, seed * 2333u + (unsigned int)i); /* cnode.cpp:574 */
        t->pool = (node_t *)calloc((size_t)b->P, sizeof(node_t));
        t->act_pool = (int *)calloc((size_t)b->P * (size_t)N, sizeof(int));
        t->path = (int *)calloc((size_t)S + 2, sizeof(int));
    }
    return b;
}

void mazo_destroy(void *h)
{
    batch_t *b = (batch_t *)h;
    if (!b) return;
    for (int i = 0; i < b->B; ++i) {
        tree_t *t = &b->trees[i];
        for (int j = 0; j < t->tot_nodes; ++j) vset_free(&t->pool[j].sub);
        free(t->pool); free(t->act_pool); free(t->path); free(t->minmax.v);
    }
    free(b->trees);
    free(b);
}

/* cnode.cpp:589-614 + 205-222 */
int mazo_prepare(void *h, const float *rewards, const float *values, const float *probs, const float *beta, int K,
                 float eps, const float *noises)
{
    batch_t *b = (batch_t *)h;
    const size_t NA = (size_t)b->N * (size_t)b->A;
    for (int i = 0; i < b->B; ++i) {
        tree_t *t = &b->trees[i];
        node_t *root = &t->pool[0];
        memset(root, 0, sizeof *root);
        root->hidx = -1;
        root->prior = root->pred_prob = root->beta = root->beta_hat = 1.0f;
        root->is_root = 1;
        t->tot_nodes = 1;
        if (tree_expand(b, t, 0, 0, rewards[i], values[i], probs + i * NA, beta + i * NA, K, eps, noises + i * NA)) return 1;
        root->visit += 1;
        if (vset_update(&root->sub, values[i], 0, b->rho, b->lam)) return 1;
    }
    return 0;
}

/* cnode.cpp:616-642 */
int mazo_batch_selection(void *h, float c_base, float c_init, float discount, int *idx_x, int *idx_y, int *act)
{
    batch_t *b = (batch_t *)h;
    int *list = (int *)malloc(sizeof(int) * (size_t)(b->K > 0 ? b->K : 1) * 4);
    for (int i = 0; i < b->B; ++i) {
        tree_t *t = &b->trees[i];
        tree_select_path(b, t, c_base, c_init, discount, list);
        idx_x[i] = t->res_idx;
        idx_y[i] = i;
        for (int j = 0; j < b->N; ++j) act[(size_t)i * b->N + j] = t->res_action[j];
    }
    free(list);
    return 0;
}

/* cnode.cpp:644-670 + 452-469 */
int mazo_batch_expansion_and_backup(void *h, int hidx, float discount, int K, const float *rewards, const float *values,
                                    const float *probs, const float *beta)
{
    batch_t *b = (batch_t *)h;
    const size_t NA = (size_t)b->N * (size_t)b->A;
    for (int i = 0; i < b->B; ++i) {
        tree_t *t = &b->trees[i];
        int leaf = t->path[t->search_len];
        if (tree_expand(b, t, leaf, hidx, rewards[i], values[i], probs + i * NA, beta + i * NA, K, 0.0f, NULL)) return 1;
        if (tree_backprop(b, t, values[i], discount)) return 1;
    }
    return 0;
}

int mazo_get_roots_values(void *h, float *out) /* cnode.cpp:672-679 */
{
    batch_t *b = (batch_t *)h;
    for (int i = 0; i < b->B; ++i) out[i] = node_value(&b->trees[i].pool[0]);
    return 0;
}

int mazo_get_roots_marginal_visit_count(void *h, int *out) /* cnode.cpp:682-690, 69-79 */
{
    batch_t *b = (batch_t *)h;
    const size_t NA = (size_t)b->N * (size_t)b->A;
    memset(out, 0, sizeof(int) * (size_t)b->B * NA);
    for (int i = 0; i < b->B; ++i) {
        const tree_t *t = &b->trees[i];
        const node_t *r = &t->pool[0];
        for (int c = 0; c < r->nchild; ++c) {
            const node_t *ch = &t->pool[r->child0 + c];
            for (int j = 0; j < b->N; ++j) out[i * NA + (size_t)j * b->A + ch->act[j]] += ch->visit;
        }
    }
    return 0;
}

int mazo_get_roots_marginal_priors(void *h, float *out) /* cnode.cpp:692-700, 81-91 */
{
    batch_t *b = (batch_t *)h;
    const size_t NA = (size_t)b->N * (size_t)b->A;
    memset(out, 0, sizeof(float) * (size_t)b->B * NA);
    for (int i = 0; i < b->B; ++i) {
        const tree_t *t = &b->trees[i];
        const node_t *r = &t->pool[0];
        for (int c = 0; c < r->nchild; ++c) {
            const node_t *ch = &t->pool[r->child0 + c];
            for (int j = 0; j < b->N; ++j) out[i * NA + (size_t)j * b->A + ch->act[j]] += ch->prior;
        }
    }
    return 0;
}

int mazo_get_roots_num_children(void *h, int *out) /* cnode.cpp:702-705 */
{
    batch_t *b = (batch_t *)h;
    for (int i = 0; i < b->B; ++i) out[i] = b->trees[i].pool[0].nchild;
    return 0;
}

/* cnode.cpp:93-171, 707-781 */
int mazo_readout(void *h, float discount, int k_pad, int *actions, int *visits, float *pred_probs, float *beta,
                 float *beta_hat, float *priors, float *imp_ratio, float *pred_values, float *mcts_values, float *rewards,
                 float *qvalues)
{
    batch_t *b = (batch_t *)h;
    for (int i = 0; i < b->B; ++i) {
        const tree_t *t = &b->trees[i];
        const node_t *r = &t->pool[0];
        if (r->nchild > k_pad) return fail("mazo_readout: k_pad smaller than num_children");
        for (int c = 0; c < r->nchild; ++c) {
            const node_t *ch = &t->pool[r->child0 + c];
            const size_t o = (size_t)i * k_pad + c;
            if (actions) for (int j = 0; j < b->N; ++j) actions[o * b->N + j] = ch->act[j];
            if (visits) visits[o] = ch->visit;
            if (pred_probs) pred_probs[o] = ch->pred_prob;
            if (beta) beta[o] = ch->beta;
            if (beta_hat) beta_hat[o] = ch->beta_hat;
            if (priors) priors[o] = ch->prior;
            if (imp_ratio) imp_ratio[o] = ch->beta_hat / ch->beta * ch->pred_prob;
            if (pred_values) pred_values[o] = ch->pred_value;
            if (mcts_values) mcts_values[o] = node_value(ch);
            if (rewards) rewards[o] = ch->reward;
            if (qvalues) qvalues[o] = node_qsa(ch, discount);
        }
    }
    return 0;
}

int mazo_stats(void *h, int *tot_nodes, int *last_search_len)
{
    batch_t *b = (batch_t *)h;
    for (int i = 0; i < b->B; ++i) {
        if (tot_nodes) tot_nodes[i] = b->trees[i].tot_nodes;
        if (last_search_len) last_search_len[i] = b->trees[i].search_len;
    }
    return 0;
}
