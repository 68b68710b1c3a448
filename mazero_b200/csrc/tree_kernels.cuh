// tree_kernels.cuh -- the sm_100a kernels of the batched sampled-MCTS tree engine.
// One warp per tree; blockDim.x = 32 * warps_per_block; trees are independent (no atomics).
#pragma once
#include "tree_step.cuh"

namespace maz {

__device__ __forceinline__ bool warp_tree(const TreeLayout &L, int &tree, int &lane)
{
    tree = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    lane = threadIdx.x & 31;
    return tree < L.B;
}

// ---- std::mt19937(seed) for every tree (cnode.cpp:574, 186-189) ----------------------------------------
// The seeding recurrence is serial per tree, so a lane owns a tree; a warp transposes its 32 trees' words
// through shared memory so that every global store is a full 128-byte line of ONE tree's state (the naive
// one-thread-per-tree store pattern touched a different line per lane: 74 us per 1024 trees, profiles/).
__global__ void __launch_bounds__(128) k_seed(TreeLayout L, char *arena, unsigned int seed_base)
{
    __shared__ uint32_t tile[4][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b0 = (blockIdx.x * 4 + warp) * 32;          // first tree of this warp
    if (b0 >= L.B) return;
    const int b = b0 + lane;
    uint32_t x = seed_base + (unsigned int)b;
    for (int c = 0; c < kMtN; c += 32) {                  // 20 chunks of 32 words (the last one is 16 words)
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int w = c + i;
            if (w > 0) x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)w;
            tile[warp][lane][i] = x;                      // row = tree, column = word
            if (w == kMtN - 1) break;
        }
        __syncwarp();
        const int nw = min(32, kMtN - c);
        for (int t = 0; t < 32; ++t) {                    // tree b0+t: lanes write its words c..c+nw-1 (one line)
            if (b0 + t < L.B && lane < nw)
                f_mt(L, arena + (size_t)(b0 + t) * L.slab_bytes)[c + lane] = tile[warp][t][lane];
        }
        __syncwarp();
    }
    if (b < L.B) {
        TreeHdr *h = f_hdr(arena + (size_t)b * L.slab_bytes);
        h->mt_pos = kMtN;
        h->err = 0;
    }
}

// ---- CTree_batch::prepare -> CTree::prepare (cnode.cpp:589-614, 205-222) ------------------------------
__global__ void k_prepare(TreeLayout L, char *arena, const float *__restrict__ lam_pow,
                          const float *__restrict__ rewards, const float *__restrict__ values,
                          const float *__restrict__ probs, const float *__restrict__ beta, int K, float eps,
                          const float *__restrict__ noises, int *g_err)
{
    extern __shared__ __align__(16) char smem[];
    int tree, lane;
    if (!warp_tree(L, tree, lane)) return;
    char *tb = arena + (size_t)tree * L.slab_bytes;
    TreeHdr *h = f_hdr(tb);
    const ExpandScratch sc = carve_scratch(smem + (size_t)(threadIdx.x >> 5) * tree_scratch_bytes(L.N, L.A, L.K, L.S), L.N, L.A);

    int tot_nodes = 1, n_expanded = 0, log_len = 0, err = 0;
    int mt_pos = h->mt_pos;
    if (lane == 0) {  // new (root) CNode(1,1,1,1,true,...)  cnode.cpp:217
        f_prior(L, tb)[0] = 1.0f;
        f_pred_prob(L, tb)[0] = 1.0f;
        f_beta(L, tb)[0] = 1.0f;
        f_beta_hat(L, tb)[0] = 1.0f;
        f_visit(L, tb)[0] = 0;
        f_qdelta(L, tb)[0] = 0.0f;   // (indexed by hidden-state index; the root's entry is never used)
    }
    const size_t NA = (size_t)L.N * L.A;
    const float value = values[tree];
    expand_node(L, tb, hot_from_slab(L, tb), tot_nodes, n_expanded, mt_pos, err, 0, 0, rewards[tree], value, probs + tree * NA,
                beta + tree * NA, K, eps, noises + tree * NA, sc, lane);
    float ws = 0.0f, wt = 0.0f;
    vs_update(L, tb, log_len, err, ws, wt, 0, 0, value, lam_pow, lane);  // root.subtree_info.update(value, 0)
    if (lane == 0) {
        f_visit(L, tb)[0] = 1;
        f_wsum(L, tb)[0] = ws;
        f_wtot(L, tb)[0] = wt;
        h->tot_nodes = tot_nodes;
        h->log_len = log_len;
        h->path_len = 0;
        h->mt_pos = mt_pos;
        h->mm_min = 0.0f;
        h->mm_max = 0.0f;
        h->mm_cnt = 0;
        h->n_expanded = n_expanded;
        h->err = err;
        h->sum_path_len = 0;
        if (err) *g_err = err;
    }
}

__global__ void k_select(TreeLayout L, char *arena, const float *__restrict__ logterm, const double *__restrict__ sqrtn,
                         int table_len, float discount, int *__restrict__ idx_x, int *__restrict__ idx_y,
                         int *__restrict__ act_out, int *g_err)
{
    int tree, lane;
    if (!warp_tree(L, tree, lane)) return;
    char *tb = arena + (size_t)tree * L.slab_bytes;
    select_path_device(L, tb, f_hdr(tb), hot_from_slab(L, tb), logterm, sqrtn, table_len, discount, tree, lane, idx_x, idx_y, act_out, g_err);
}

__global__ void k_expand_backup(TreeLayout L, char *arena, const float *__restrict__ lam_pow, int hidx, float discount,
                                int K, const float *__restrict__ rewards, const float *__restrict__ values,
                                const float *__restrict__ probs, const float *__restrict__ beta, int *g_err)
{
    extern __shared__ __align__(16) char smem[];
    int tree, lane;
    if (!warp_tree(L, tree, lane)) return;
    char *tb = arena + (size_t)tree * L.slab_bytes;
    const StepScratch scr = carve_step_scratch(smem + (size_t)(threadIdx.x >> 5) * tree_scratch_bytes(L.N, L.A, L.K, L.S), L.N, L.A, L.K, L.S);
    const size_t NA = (size_t)L.N * L.A;
    expand_backup_device(L, tb, f_hdr(tb), hot_from_slab(L, tb), lam_pow, hidx, discount, K, rewards + tree, values + tree, probs + tree * NA,
                         beta + tree * NA, scr, lane, g_err);
}

// ---- fused: expansion + backup of simulation s, then selection of simulation s+1 (one launch per simulation
// in the on-device search loop: the tree's header and hot nodes are still in L1) ------------------------------------
__global__ void k_expand_backup_select(TreeLayout L, char *arena, const float *__restrict__ lam_pow, int hidx, float discount,
                                       int K, const float *__restrict__ rewards, const float *__restrict__ values,
                                       const float *__restrict__ probs, const float *__restrict__ beta,
                                       const float *__restrict__ logterm, const double *__restrict__ sqrtn, int table_len,
                                       int *__restrict__ idx_x, int *__restrict__ idx_y, int *__restrict__ act_out, int *g_err)
{
    extern __shared__ __align__(16) char smem[];
    int tree, lane;
    if (!warp_tree(L, tree, lane)) return;
    char *tb = arena + (size_t)tree * L.slab_bytes;
    const StepScratch scr = carve_step_scratch(smem + (size_t)(threadIdx.x >> 5) * tree_scratch_bytes(L.N, L.A, L.K, L.S), L.N, L.A, L.K, L.S);
    const size_t NA = (size_t)L.N * L.A;
    expand_backup_device(L, tb, f_hdr(tb), hot_from_slab(L, tb), lam_pow, hidx, discount, K, rewards + tree, values + tree, probs + tree * NA,
                         beta + tree * NA, scr, lane, g_err, tree);
    __syncwarp();
    select_next_device(L, tb, f_hdr(tb), hot_from_slab(L, tb), logterm, sqrtn, table_len, discount, tree, lane, &scr, idx_x, idx_y, act_out, g_err);
    MAZ_TS(L, tree, lane, 6);
}

// ---- the same step with TWO warps per tree: expansion and back-propagation touch disjoint state (the expansion
// samples and creates children: RNG, node pool; the backup walks the path: value log, visit counts, q-deltas), so they
// run concurrently; the next selection needs both and follows a block barrier.  Same arithmetic, same results; the
// tree step is a chain of dependent instructions on a mostly idle SM, so the second warp is free. -----------------------
__global__ void k_expand_backup_select2(TreeLayout L, char *arena, const float *__restrict__ lam_pow, int hidx, float discount,
                                        int K, const float *__restrict__ rewards, const float *__restrict__ values,
                                        const float *__restrict__ probs, const float *__restrict__ beta,
                                        const float *__restrict__ logterm, const double *__restrict__ sqrtn, int table_len,
                                        int *__restrict__ idx_x, int *__restrict__ idx_y, int *__restrict__ act_out, int *g_err)
{
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = warp >> 1, role = warp & 1;              // role 0: expansion (+ selection), role 1: backup
    const int tree = blockIdx.x * (blockDim.x >> 6) + pair;
    const bool active = tree < L.B;
    const StepScratch scr = carve_step_scratch(smem + (size_t)pair * tree_scratch_bytes(L.N, L.A, L.K, L.S), L.N, L.A, L.K, L.S);
    char *tb = arena + (size_t)(active ? tree : 0) * L.slab_bytes;
    TreeHdr *h = f_hdr(tb);
    const size_t NA = (size_t)L.N * L.A;
    griddep_launch();
    // profiling only (TreeLayout::dbg_clock, >= 64 + 8*B entries): per tree, cycles of expansion / backup / selection + depths
    long long *const dbg = (L.dbg_clock && active && lane == 0) ? L.dbg_clock + 64 + (size_t)tree * 8 : nullptr;
    const long long t0 = dbg ? clock64() : 0;
    // both warps read the header BEFORE either of them may write it
    const int tot_nodes0 = h->tot_nodes, log_len0 = h->log_len, mt_pos0 = h->mt_pos, n_exp0 = h->n_expanded, err0 = h->err;
    const int len = h->path_len;
    __syncthreads();
    if (active && role == 0) {
        // ---------------- expansion (cnode.cpp:224-295) ----------------
        const ExpandScratch sc = scr.ex;
        int tot_nodes = tot_nodes0, n_expanded = n_exp0, mt_pos = mt_pos0, err = err0;
        const int leaf = f_path(L, tb)[len];
        const int n_draw = (L.A >= 2) ? 2 * K * L.N : 0;
        const bool draws_pre = n_draw > 0 && n_draw <= kMtChunk && mt_pos + n_draw <= kMtN;
        if (draws_pre) {
            const uint32_t *mt = f_mt(L, tb);
            for (int t = lane; t < n_draw; t += 32) sc.draws[t] = mt_temper(mt[mt_pos + t]);
            mt_pos += n_draw;
        }
        griddep_wait();
        expand_node(L, tb, hot_from_slab(L, tb), tot_nodes, n_expanded, mt_pos, err, leaf, hidx, __ldcg(rewards + tree), __ldcg(values + tree),
                    probs + tree * NA, beta + tree * NA, K, 0.0f, nullptr, sc, lane, draws_pre, -1, /*init_stats=*/false, /*depth=*/len);
        if (lane == 0) {
            h->tot_nodes = tot_nodes;
            h->mt_pos = mt_pos;
            h->n_expanded = n_expanded;
            h->err = err;
            if (err) *g_err = err;
        }
    } else if (active) {
        // ---------------- back_propagate (cnode.cpp:415-450): the shared backup_parallel ----------------
        int log_len = log_len0, err = 0;
        LogRegs lr;
        log_prefetch(L, tb, log_len0, lane, lr);
        griddep_wait();
        const float reward_in = __ldcg(rewards + tree), value = __ldcg(values + tree);
        backup_parallel(L, tb, hot_from_slab(L, tb), lam_pow, scr, lr, len, log_len0, n_exp0, reward_in, value, discount, lane, log_len, err);
        const int n_exp1 = n_exp0 + 1;
        float mn, mx;
        minmax_reduce(hot_from_slab(L, tb), n_exp1, lane, mn, mx);
        if (lane == 0) {
            h->log_len = log_len;
            h->mm_cnt = n_exp1 - 1;
            h->mm_min = mn;
            h->mm_max = mx;
            if (err) *g_err = err;
        }
    }
    if (dbg) dbg[role] = clock64() - t0;
    __syncthreads();   // expansion and backup of every tree of the block are complete and visible
    const long long t2 = dbg ? clock64() : 0;
    if (active && role == 0)
        select_next_device(L, tb, h, hot_from_slab(L, tb), logterm, sqrtn, table_len, discount, tree, lane, &scr, idx_x, idx_y, act_out, g_err);
    if (dbg && role == 0) {
        dbg[2] = clock64() - t2;
        dbg[3] = len;
        dbg[4] = h->path_len;
        dbg[5] = t2 - t0;
    }
}

// ---- readouts (cnode.cpp:69-171, 471-530, 672-781) ------------------------------------------------------
struct ReadoutPtrs {
    float *values;
    int *marg_visits;
    float *marg_priors;
    int *num_children;
    int *actions;
    int *visits;
    float *pred_probs, *beta, *beta_hat, *priors, *imp_ratio, *pred_values, *mcts_values, *rewards, *qvalues;
};

__global__ void k_readout(TreeLayout L, char *arena, float discount, ReadoutPtrs o)
{
    int tree, lane;
    if (!warp_tree(L, tree, lane)) return;
    char *tb = arena + (size_t)tree * L.slab_bytes;
    const int C = f_nchild(L, tb)[0];
    const int base = f_cbase(L, tb)[0];
    const int K = L.K, N = L.N, A = L.A;
    if (lane == 0) {
        if (o.values) o.values[tree] = __fdiv_rn(f_wsum(L, tb)[0], f_wtot(L, tb)[0]);
        if (o.num_children) o.num_children[tree] = C;
    }
    int cvis = 0;
    float prior = 0.0f;
    if (lane < K) {
        const size_t oi = (size_t)tree * K + lane;
        float pp = 0, b = 0, bh = 0, imp = 0, pv = 0, mv = 0, rew = 0, q = 0;
        if (lane < C) {
            const int cs = base + lane;
            prior = f_prior(L, tb)[cs];
            cvis = f_visit(L, tb)[cs];
            pp = f_pred_prob(L, tb)[cs];
            b = f_beta(L, tb)[cs];
            bh = f_beta_hat(L, tb)[cs];
            imp = __fmul_rn(__fdiv_rn(bh, b), pp);
            pv = f_pred_value(L, tb)[cs];
            rew = f_reward(L, tb)[cs];
            if (f_nchild(L, tb)[cs] > 0) mv = __fdiv_rn(f_wsum(L, tb)[cs], f_wtot(L, tb)[cs]);
            q = __fadd_rn(rew, __fmul_rn(discount, mv));
        }
        if (o.visits) o.visits[oi] = cvis;
        if (o.pred_probs) o.pred_probs[oi] = pp;
        if (o.beta) o.beta[oi] = b;
        if (o.beta_hat) o.beta_hat[oi] = bh;
        if (o.priors) o.priors[oi] = prior;
        if (o.imp_ratio) o.imp_ratio[oi] = imp;
        if (o.pred_values) o.pred_values[oi] = pv;
        if (o.mcts_values) o.mcts_values[oi] = mv;
        if (o.rewards) o.rewards[oi] = rew;
        if (o.qvalues) o.qvalues[oi] = q;
    }
    const uint8_t *act = f_actions(L, tb) + (size_t)base * N;
    if (o.actions) {
        for (int t = lane; t < K * N; t += 32) o.actions[(size_t)tree * K * N + t] = (t < C * N) ? (int)act[t] : 0;
    }
    if (o.marg_visits || o.marg_priors) {
        // scatter-add over root children in child order (float adds in that order, cnode.cpp:81-91)
        for (int t0 = 0; t0 < N * A; t0 += 32) {
            const int t = t0 + lane;
            const int j = t / A, a = t - j * A;
            int iv = 0;
            float fp = 0.0f;
            for (int c = 0; c < C; ++c) {
                const int vcc = __shfl_sync(MAZ_FULL, cvis, c);
                const float prc = __shfl_sync(MAZ_FULL, prior, c);
                if (t < N * A && act[(size_t)c * N + j] == a) {
                    iv += vcc;
                    fp = __fadd_rn(fp, prc);
                }
            }
            if (t < N * A) {
                if (o.marg_visits) o.marg_visits[(size_t)tree * N * A + t] = iv;
                if (o.marg_priors) o.marg_priors[(size_t)tree * N * A + t] = fp;
            }
        }
    }
}

// ---- statistics ------------------------------------------------------------------------------------------
__global__ void k_stats(TreeLayout L, char *arena, int *tot_nodes, int *last_len, unsigned long long *sums)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.B) return;
    const TreeHdr *h = f_hdr(arena + (size_t)b * L.slab_bytes);
    if (tot_nodes) tot_nodes[b] = h->tot_nodes;
    if (last_len) last_len[b] = h->path_len;
    atomicAdd(&sums[0], (unsigned long long)h->sum_path_len);
    atomicAdd(&sums[1], (unsigned long long)h->n_expanded);
}

}  // namespace maz
