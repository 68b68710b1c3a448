// tree_kernels.cuh -- the sm_100a kernels of the batched sampled-MCTS tree engine.
// One warp per tree; blockDim.x = 32 * warps_per_block; trees are independent (no atomics).
#pragma once
#include "tree_device.cuh"

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
    const ExpandScratch sc = carve_scratch(smem + (size_t)(threadIdx.x >> 5) * expand_scratch_bytes(L.N, L.A, L.K), L.N, L.A);

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
    expand_node(L, tb, tot_nodes, n_expanded, mt_pos, err, 0, 0, rewards[tree], value, probs + tree * NA,
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

// ---- CTree_batch::cbatch_selection -> CTree::select_path / select_child / ucb_score --------------------
// (cnode.cpp:616-642, 381-413, 337-379, 297-335)
// Latency-bound pointer chase: ONE memory round trip per tree level.  While the children of the current
// node are scored, each child's own header (num_children, child_base, visit, pred_value, hidden index) is
// already in its lane's registers, so descending is a shuffle.  The next few mt19937 outputs are
// prefetched at kernel start (one raw draw per select_child call).
__device__ __forceinline__ void select_path_device(const TreeLayout &L, char *tb, TreeHdr *h, const float *__restrict__ logterm,
                                                   const double *__restrict__ sqrtn, int table_len, float discount, int tree,
                                                   int lane, int *__restrict__ idx_x, int *__restrict__ idx_y,
                                                   int *__restrict__ act_out, int *g_err)
{
    uint16_t *path = f_path(L, tb);
    const auto nchild = f_nchild(L, tb);
    const auto cbase = f_cbase(L, tb);
    const auto visit = f_visit(L, tb);
    const auto pred_value = f_pred_value(L, tb);
    const auto hidx = f_hidx(L, tb);
    uint32_t *mt = f_mt(L, tb);

    // round trip 0: tree header, root header, prefetched random words
    int mt_pos = h->mt_pos;
    const float mn = h->mm_min, mx = h->mm_max;
    const int mmc = h->mm_cnt;
    const RecRegs root = rec_load(L, tb, 0);
    int C = rec_nchild(root), base = rec_cbase(root), vc = rec_visit(root);
    float pq = rec_pred_value(root);
    constexpr int kPre = 8;
    uint32_t pre = 0;                          // lane l < kPre holds raw state word mt_pos + l (if in this block)
    if (lane < kPre && mt_pos + lane < kMtN) pre = mt[mt_pos + lane];
    int pre_used = 0;                          // draws consumed from the prefetched words
    const int pre_avail = min(kPre, max(0, kMtN - mt_pos));

    int node = 0, parent_hidx = 0, node_hidx = 0, len = 0, err = 0;
    if (lane == 0) path[0] = 0;
    while (C > 0) {
        // round trip `len+1`: children fields + the children's own headers
        float prior = 0.f, rew = 0.f, ws = 0.f, wt = 1.f, cpq = 0.f;
        int cvis = 0, cC = 0, cbase_c = 0, chidx = -1;
        if (lane < C) {                       // the child's whole record: two 16-byte loads, ONE round trip per level
            const RecRegs r = rec_load(L, tb, base + lane);
            prior = rec_prior(r);
            cvis = rec_visit(r);
            rew = rec_reward(r);
            cC = rec_nchild(r);
            cbase_c = rec_cbase(r);
            cpq = rec_pred_value(r);
            chidx = rec_hidx(r);
            ws = rec_wsum(r);                 // (zero for never-visited children, unused then)
            wt = rec_wtot(r);
        }
        int ci;
        if (node == 0 && vc <= C) {
            ci = vc - 1;  // forced root round-robin, no RNG draw (cnode.cpp:398-399)
        } else {
            int n = vc - 1;
            if (n >= table_len) n = table_len - 1;
            float score = 0.0f;
            if (lane < C) {
                // pb_c = log((n + c_base + 1)/c_base) + c_init   [float <- double]   (host table)
                // pb_c *= sqrt(n) / (visit + 1)                   [float <- double product]
                const float pb_c = (L.pbc_dim > 0)
                                       ? L.pbc_table[(size_t)n * L.pbc_dim + min(cvis, L.pbc_dim - 1)]
                                       : (float)__dmul_rn((double)logterm[n], __ddiv_rn(sqrtn[n], (double)(cvis + 1)));
                const float prior_score = __fmul_rn(pb_c, prior);
                float v = 0.0f;
                if (cvis != 0) v = __fsub_rn(__fadd_rn(rew, __fmul_rn(discount, __fdiv_rn(ws, wt))), pq);
                if (mmc > 0) {  // CMinMaxStats::normalize (utils.cpp:95-103)
                    const float delta = __fsub_rn(mx, mn);
                    const float den = (L.delta_lb < delta) ? delta : L.delta_lb;
                    v = __fdiv_rn(__fsub_rn(v, mn), den);
                }
                if (v < 0.0f) v = 0.0f;
                if (v > 1.0f) v = 1.0f;
                score = __fadd_rn(prior_score, v);
            }
            // sequential epsilon-tie list of cnode.cpp:351-370, evaluated in parallel:
            // list = {first index of the maximum} U {later indices with score >= max - 1e-6f}
            const bool valid = (lane < C) && (score > -1000000.0f);
            const uint32_t o = valid ? f2ord(__fadd_rn(score, 0.0f)) : 0u;
            const uint32_t gmax = __reduce_max_sync(MAZ_FULL, o);
            const unsigned anyvalid = __ballot_sync(MAZ_FULL, valid);
            unsigned listmask;
            if (anyvalid) {
                const float M = ord2f(gmax);
                const unsigned ismax = __ballot_sync(MAZ_FULL, valid && score == M);
                const int istar = __ffs(ismax) - 1;
                const float thr = __fsub_rn(M, 0.000001f);
                listmask = __ballot_sync(MAZ_FULL, (lane < C) && (lane > istar) && (score >= thr)) | (1u << istar);
            } else {
                const float thr = __fsub_rn(-1000000.0f, 0.000001f);
                listmask = __ballot_sync(MAZ_FULL, (lane < C) && (score >= thr));
            }
            const int nl = __popc(listmask);
            ci = 0;
            if (nl > 0) {  // one raw draw even for a single candidate (cnode.cpp:373-377)
                uint32_t r;
                if (pre_used < pre_avail) {
                    r = mt_temper(__shfl_sync(MAZ_FULL, pre, pre_used));
                    ++pre_used;
                    ++mt_pos;
                } else {
                    r = mt_next(mt, mt_pos, lane);
                }
                uint32_t mrem = listmask;                   // drop the (r % nl) lowest candidates, take the next one
                for (uint32_t skip = r % (uint32_t)nl; skip > 0; --skip) mrem &= mrem - 1;
                ci = __ffs(mrem) - 1;
            }
        }
        // descend: the chosen child's header is in lane ci's registers
        parent_hidx = node_hidx;
        node = base + ci;
        node_hidx = __shfl_sync(MAZ_FULL, chidx, ci);
        C = __shfl_sync(MAZ_FULL, cC, ci);
        base = __shfl_sync(MAZ_FULL, cbase_c, ci);
        vc = __shfl_sync(MAZ_FULL, cvis, ci);
        pq = __shfl_sync(MAZ_FULL, cpq, ci);
        ++len;
        if (len > L.S + 1) {
            err = kErrPathOverflow;
            break;
        }
        if (lane == 0) path[len] = (uint16_t)node;
    }
    if (lane == 0) {
        h->path_len = len;
        h->mt_pos = mt_pos;
        h->sum_path_len += len;
        idx_x[tree] = parent_hidx;
        idx_y[tree] = tree;
        if (err) {
            h->err = err;
            *g_err = err;
        }
    }
    const uint8_t *act = f_actions(L, tb) + (size_t)node * L.N;
    for (int j = lane; j < L.N; j += 32) act_out[(size_t)tree * L.N + j] = act[j];
}

__global__ void k_select(TreeLayout L, char *arena, const float *__restrict__ logterm, const double *__restrict__ sqrtn,
                         int table_len, float discount, int *__restrict__ idx_x, int *__restrict__ idx_y,
                         int *__restrict__ act_out, int *g_err)
{
    int tree, lane;
    if (!warp_tree(L, tree, lane)) return;
    char *tb = arena + (size_t)tree * L.slab_bytes;
    select_path_device(L, tb, f_hdr(tb), logterm, sqrtn, table_len, discount, tree, lane, idx_x, idx_y, act_out, g_err);
}

// ---- CTree_batch::cbatch_expansion_and_backup -> expand_and_backprop / back_propagate ------------------
// (cnode.cpp:644-670, 452-469, 415-450)
// Latency plan: everything the backup needs is requested up front, in parallel with the expansion:
//   round trip 1: tree header;   round trip 2: path[], the value log (first 256 entries, in registers),
//   the mt19937 words of this expansion, beta / probs;   round trip 3: per-path-node fields (lane i <-> path[i]).
// The log snapshot stays valid for the whole backup: every path node owns a distinct (slot, depth) tag, and
// entries appended / flags flipped for one node never match another node's tag.
constexpr int kLogRegs = 8;   // 8 x 32 log entries cached in registers; longer logs: tail scanned from memory

__device__ __forceinline__ void expand_backup_device(const TreeLayout &L, char *tb, TreeHdr *h, const float *__restrict__ lam_pow,
                                                     int hidx, float discount, int K, const float *__restrict__ reward_ptr,
                                                     const float *__restrict__ value_ptr,
                                                     const float *__restrict__ probs, const float *__restrict__ beta,
                                                     const ExpandScratch &sc, int lane, int *g_err, int tree = -1)
{
    MAZ_TS(L, tree, lane, 0);
    griddep_launch();   // PDL: the next kernel (inference of the next simulation) may start its prologue now
    int tot_nodes = h->tot_nodes, log_len = h->log_len, mt_pos = h->mt_pos, n_expanded = h->n_expanded, err = h->err;
    const int len = h->path_len;
    const int log_len0 = log_len;
    const int leaf_eid = n_expanded;                  // expansion order the leaf is about to get
    const uint16_t *path = f_path(L, tb);
    const uint32_t *vk = f_vskey(L, tb);
    const float *vv = f_vsval(L, tb);
    const auto reward = f_reward(L, tb);
    const auto wsum = f_wsum(L, tb);
    const auto wtot = f_wtot(L, tb);

    MAZ_TS(L, tree, lane, 1);
    // ---- round trip 2: issue everything that depends only on the header -------------------------------------
    const bool fast = len < 32;                       // path fits one lane per node (else: per-node loads below)
    int my_slot = 0;
    if (fast && lane <= len) my_slot = path[lane];
    uint32_t lk[kLogRegs];
    float lv[kLogRegs];
#pragma unroll
    for (int c = 0; c < kLogRegs; ++c) {
        const int e = c * 32 + lane;
        lk[c] = (e < log_len0) ? vk[e] : 0xffffffffu;   // tag 0x7fffffff<<1 never matches (slot < 65536, depth < 32768)
        lv[c] = (e < log_len0) ? vv[e] : 0.0f;
    }
    const int n_draw = (L.A >= 2) ? 2 * K * L.N : 0;
    const bool draws_pre = n_draw > 0 && n_draw <= kMtChunk && mt_pos + n_draw <= kMtN;
    if (draws_pre) {
        const uint32_t *mt = f_mt(L, tb);
        for (int t = lane; t < n_draw; t += 32) sc.draws[t] = mt_temper(mt[mt_pos + t]);
        mt_pos += n_draw;
    }
    const float my_lp = (fast && lane <= len) ? lam_pow[lane] : 0.f;   // lam_pow[depth], depth = lane
    // ---- round trip 3: per-path-node fields, lane i <-> path[i] ------------------------------------------------
    const int leaf = fast ? __shfl_sync(MAZ_FULL, my_slot, len) : (int)path[len];
    float my_rew = 0.f, my_ws = 0.f, my_wt = 0.f, my_ppv = 0.f;
    int my_vis = 0, my_hidx = 0;
    if (fast && lane < len) {                         // the leaf (lane == len) is filled by the expansion below
        const RecRegs r = rec_load(L, tb, my_slot);
        my_rew = rec_reward(r);
        my_ws = rec_wsum(r);
        my_wt = rec_wtot(r);
        my_vis = rec_visit(r);
        my_hidx = rec_eid(r);
    }
    {
        const int prev = __shfl_up_sync(MAZ_FULL, my_slot, 1);
        if (fast && lane >= 1 && lane <= len) my_ppv = f_pred_value(L, tb)[prev];   // parent's pred_value
    }

    MAZ_TS(L, tree, lane, 2);
    // PDL: everything above only touched this tree's own state (written by the previous tree kernel, long
    // complete); the network outputs of THIS simulation are produced by the kernel we may be overlapping with.
    griddep_wait();
    const float reward_in = __ldcg(reward_ptr), value = __ldcg(value_ptr);   // L2 loads, see expand_node
    expand_node(L, tb, tot_nodes, n_expanded, mt_pos, err, leaf, hidx, reward_in, value, probs, beta, K, 0.0f, nullptr, sc,
                lane, draws_pre, tree);
    MAZ_TS(L, tree, lane, 3);

    // ---- back_propagate (cnode.cpp:415-450) -----------------------------------------------------------------------
    float *qd = f_qdelta(L, tb);                      // q-delta of the e-th expanded node (the CMinMaxStats entries)
    float G = value;
    for (int i = len; i >= 0; --i) {
        int slot, vis, nh;
        float rew, ws, wt, ppv;
        if (fast) {
            slot = __shfl_sync(MAZ_FULL, my_slot, i);
            rew = __shfl_sync(MAZ_FULL, my_rew, i);
            ws = __shfl_sync(MAZ_FULL, my_ws, i);
            wt = __shfl_sync(MAZ_FULL, my_wt, i);
            vis = __shfl_sync(MAZ_FULL, my_vis, i);
            nh = __shfl_sync(MAZ_FULL, my_hidx, i);
            ppv = __shfl_sync(MAZ_FULL, my_ppv, i);
        } else {
            slot = path[i];
            rew = reward[slot]; ws = wsum[slot]; wt = wtot[slot];
            vis = f_visit(L, tb)[slot];
            nh = f_eid(L, tb)[slot];
            ppv = (i > 0) ? f_pred_value(L, tb)[path[i - 1]] : 0.f;
        }
        if (i == len) { rew = reward_in; ws = 0.f; wt = 0.f; vis = 0; nh = leaf_eid; }   // freshly expanded leaf
        // (the reference removes the node's old q-delta from the min-max multiset here; in this layout the
        //  entry simply lives in qd[hidden index] and is overwritten below)
        const uint32_t tag = vs_tag(slot, len - i);
        VsScan r;
        vs_scan_init(r);
#pragma unroll
        for (int c = 0; c < kLogRegs; ++c) vs_scan_entry(r, tag, lk[c], lv[c], c * 32 + lane);
        for (int e = kLogRegs * 32 + lane; e < log_len0; e += 32) vs_scan_entry(r, tag, vk[e], vv[e], e);
        const float lp = fast ? __shfl_sync(MAZ_FULL, my_lp, len - i) : lam_pow[len - i];
        vs_apply(L, tb, log_len, err, ws, wt, tag, lp, G, lane, r);
        if (lane == 0) {
            f_visit(L, tb)[slot] = vis + 1;
            wsum[slot] = ws;
            wtot[slot] = wt;
            if (i != 0) qd[nh] = __fsub_rn(__fadd_rn(rew, __fmul_rn(discount, __fdiv_rn(ws, wt))), ppv);
        }
        G = __fadd_rn(rew, __fmul_rn(discount, G));
    }
    __syncwarp();
    MAZ_TS(L, tree, lane, 4);
    // CMinMaxStats min / max = reduction over the q-deltas of all visited (= expanded) non-root nodes
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int e = 1 + lane; e < n_expanded; e += 32) {
        const uint32_t o = f2ord(qd[e]);
        lo = min(lo, o);
        hi = max(hi, o);
    }
    lo = __reduce_min_sync(MAZ_FULL, lo);
    hi = __reduce_max_sync(MAZ_FULL, hi);
    if (lane == 0) {
        h->tot_nodes = tot_nodes;
        h->log_len = log_len;
        h->mt_pos = mt_pos;
        h->n_expanded = n_expanded;
        h->mm_cnt = n_expanded - 1;
        h->mm_min = ord2f(lo);
        h->mm_max = ord2f(hi);
        h->err = err;
        if (err) *g_err = err;
    }
    MAZ_TS(L, tree, lane, 5);
}

__global__ void k_expand_backup(TreeLayout L, char *arena, const float *__restrict__ lam_pow, int hidx, float discount,
                                int K, const float *__restrict__ rewards, const float *__restrict__ values,
                                const float *__restrict__ probs, const float *__restrict__ beta, int *g_err)
{
    extern __shared__ __align__(16) char smem[];
    int tree, lane;
    if (!warp_tree(L, tree, lane)) return;
    char *tb = arena + (size_t)tree * L.slab_bytes;
    const ExpandScratch sc = carve_scratch(smem + (size_t)(threadIdx.x >> 5) * expand_scratch_bytes(L.N, L.A, L.K), L.N, L.A);
    const size_t NA = (size_t)L.N * L.A;
    expand_backup_device(L, tb, f_hdr(tb), lam_pow, hidx, discount, K, rewards + tree, values + tree, probs + tree * NA,
                         beta + tree * NA, sc, lane, g_err);
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
    const ExpandScratch sc = carve_scratch(smem + (size_t)(threadIdx.x >> 5) * expand_scratch_bytes(L.N, L.A, L.K), L.N, L.A);
    const size_t NA = (size_t)L.N * L.A;
    expand_backup_device(L, tb, f_hdr(tb), lam_pow, hidx, discount, K, rewards + tree, values + tree, probs + tree * NA,
                         beta + tree * NA, sc, lane, g_err, tree);
    __syncwarp();
    select_path_device(L, tb, f_hdr(tb), logterm, sqrtn, table_len, discount, tree, lane, idx_x, idx_y, act_out, g_err);
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
        const ExpandScratch sc = carve_scratch(smem + (size_t)pair * expand_scratch_bytes(L.N, L.A, L.K), L.N, L.A);
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
        expand_node(L, tb, tot_nodes, n_expanded, mt_pos, err, leaf, hidx, __ldcg(rewards + tree), __ldcg(values + tree),
                    probs + tree * NA, beta + tree * NA, K, 0.0f, nullptr, sc, lane, draws_pre, -1, /*init_stats=*/false);
        if (lane == 0) {
            h->tot_nodes = tot_nodes;
            h->mt_pos = mt_pos;
            h->n_expanded = n_expanded;
            h->err = err;
            if (err) *g_err = err;
        }
    } else if (active) {
        // ---------------- back_propagate (cnode.cpp:415-450) ----------------
        int log_len = log_len0, err = 0;
        const uint16_t *path = f_path(L, tb);
        const uint32_t *vk = f_vskey(L, tb);
        const float *vv = f_vsval(L, tb);
        const auto reward = f_reward(L, tb);
    const auto wsum = f_wsum(L, tb);
    const auto wtot = f_wtot(L, tb);
        const bool fast = len < 32;
        int my_slot = 0;
        if (fast && lane <= len) my_slot = path[lane];
        uint32_t lk[kLogRegs];
        float lv[kLogRegs];
#pragma unroll
        for (int c = 0; c < kLogRegs; ++c) {
            const int e = c * 32 + lane;
            lk[c] = (e < log_len0) ? vk[e] : 0xffffffffu;
            lv[c] = (e < log_len0) ? vv[e] : 0.0f;
        }
        const float my_lp = (fast && lane <= len) ? lam_pow[lane] : 0.f;
        float my_rew = 0.f, my_ws = 0.f, my_wt = 0.f, my_ppv = 0.f;
        int my_vis = 0, my_eid = 0;
        if (fast && lane < len) {
            const RecRegs r = rec_load(L, tb, my_slot);
            my_rew = rec_reward(r);
            my_ws = rec_wsum(r);
            my_wt = rec_wtot(r);
            my_vis = rec_visit(r);
            my_eid = rec_eid(r);
        }
        {
            const int prev = __shfl_up_sync(MAZ_FULL, my_slot, 1);
            if (fast && lane >= 1 && lane <= len) my_ppv = f_pred_value(L, tb)[prev];
        }
        griddep_wait();
        const float reward_in = __ldcg(rewards + tree), value = __ldcg(values + tree);
        float *qd = f_qdelta(L, tb);
        float G = value;
        for (int i = len; i >= 0; --i) {
            int slot, vis, nh;
            float rew, ws, wt, ppv;
            if (fast) {
                slot = __shfl_sync(MAZ_FULL, my_slot, i);
                rew = __shfl_sync(MAZ_FULL, my_rew, i);
                ws = __shfl_sync(MAZ_FULL, my_ws, i);
                wt = __shfl_sync(MAZ_FULL, my_wt, i);
                vis = __shfl_sync(MAZ_FULL, my_vis, i);
                nh = __shfl_sync(MAZ_FULL, my_eid, i);
                ppv = __shfl_sync(MAZ_FULL, my_ppv, i);
            } else {
                slot = path[i];
                rew = reward[slot]; ws = wsum[slot]; wt = wtot[slot];
                vis = f_visit(L, tb)[slot];
                nh = f_eid(L, tb)[slot];
                ppv = (i > 0) ? f_pred_value(L, tb)[path[i - 1]] : 0.f;
            }
            if (i == len) { rew = reward_in; ws = 0.f; wt = 0.f; vis = 0; nh = n_exp0; }   // the leaf being expanded
            const uint32_t tag = vs_tag(slot, len - i);
            VsScan r;
            vs_scan_init(r);
#pragma unroll
            for (int c = 0; c < kLogRegs; ++c) vs_scan_entry(r, tag, lk[c], lv[c], c * 32 + lane);
            for (int e = kLogRegs * 32 + lane; e < log_len0; e += 32) vs_scan_entry(r, tag, vk[e], vv[e], e);
            const float lp = fast ? __shfl_sync(MAZ_FULL, my_lp, len - i) : lam_pow[len - i];
            vs_apply(L, tb, log_len, err, ws, wt, tag, lp, G, lane, r);
            if (lane == 0) {
                f_visit(L, tb)[slot] = vis + 1;
                wsum[slot] = ws;
                wtot[slot] = wt;
                if (i != 0) qd[nh] = __fsub_rn(__fadd_rn(rew, __fmul_rn(discount, __fdiv_rn(ws, wt))), ppv);
            }
            G = __fadd_rn(rew, __fmul_rn(discount, G));
        }
        __syncwarp();
        const int n_exp1 = n_exp0 + 1;
        uint32_t lo = 0xffffffffu, hi = 0u;
        for (int e = 1 + lane; e < n_exp1; e += 32) {
            const uint32_t o = f2ord(qd[e]);
            lo = min(lo, o);
            hi = max(hi, o);
        }
        lo = __reduce_min_sync(MAZ_FULL, lo);
        hi = __reduce_max_sync(MAZ_FULL, hi);
        if (lane == 0) {
            h->log_len = log_len;
            h->mm_cnt = n_exp1 - 1;
            h->mm_min = ord2f(lo);
            h->mm_max = ord2f(hi);
            if (err) *g_err = err;
        }
    }
    if (dbg) dbg[role] = clock64() - t0;
    __syncthreads();   // expansion and backup of every tree of the block are complete and visible
    const long long t2 = dbg ? clock64() : 0;
    if (active && role == 0)
        select_path_device(L, tb, h, logterm, sqrtn, table_len, discount, tree, lane, idx_x, idx_y, act_out, g_err);
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
