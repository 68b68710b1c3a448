// tree_step.cuh -- the per-tree step functions of the sampled-MCTS tree engine (one warp = one tree), shared by the
// stand-alone tree kernels (tree_kernels.cuh) and the persistent whole-search kernel (search_persist.cuh).
//
// Cost model that shaped this file (measured, B200, profiles/): a tree step is a chain of dependent instructions and L2 round
// trips run by ONE warp, and every simulation ends with the slowest tree of its group.  A deep, narrow tree used to cost
// ~2 k cycles per path level in the selection and ~1.2 k in the backup (57 k cycles per simulation at mean depth 12 against
// 20 k for a typical tree), so ONE deep tree set the pace of the whole search.  Both are depth-independent now:
//   * backup: one pass over the tree's value log, every entry read ONCE by one lane and scattered to its path level with
//     shared-memory atomics (an entry (slot, depth) can only belong to the path node at distance `depth` from the leaf);
//   * selection: when the last path was long, every EXPANDED node picks its child at once (one lane per node; the RNG word
//     a node would consume is known from its depth), then the path is a pointer chase through shared memory.
// Every float operation is the reference's, in the reference's order: results are bit-identical in either form.
#pragma once
#include "tree_device.cuh"

namespace maz {

// ---- per-warp shared-memory scratch of a tree step ---------------------------------------------------------------------
struct LevelAcc {                     // one (node, depth) set of SubTreeValueSet, accumulated over the value log
    unsigned int cnt, nbig;           // entries in the set / in its big half
    unsigned long long minb, maxs;    // (f2ord(value) << 32 | log position): min over big, max over small
};
struct StepScratch {
    ExpandScratch ex;                 // expansion staging (tree_device.cuh)
    LevelAcc *acc;                    // [S + 2] backup accumulators, one per path level
    uint32_t *chase;                  // [S + 2] wide selection: (chosen child slot << 16) | (child's expansion index + 1, 0 = leaf)
    uint16_t *pathc;                  // [S + 2] copy of the path in shared memory (backup's entry -> level lookup)
};
__host__ __device__ inline size_t tree_scratch_bytes(int N, int A, int K, int S)
{
    const size_t lv = (size_t)S + 2;
    size_t b = expand_scratch_bytes(N, A, K) + lv * sizeof(LevelAcc) + lv * 4 + ((lv * 2 + 15) & ~(size_t)15);
    return (b + 15) & ~(size_t)15;
}
__device__ __forceinline__ StepScratch carve_step_scratch(char *p, int N, int A, int K, int S)
{
    StepScratch s;
    s.ex = carve_scratch(p, N, A);
    p += expand_scratch_bytes(N, A, K);           // (multiple of 16)
    s.acc = reinterpret_cast<LevelAcc *>(p);
    p += (size_t)(S + 2) * sizeof(LevelAcc);      // 24-byte entries: stays 8-byte aligned
    s.chase = reinterpret_cast<uint32_t *>(p);
    p += (size_t)(S + 2) * 4;
    s.pathc = reinterpret_cast<uint16_t *>(p);
    return s;
}

// ---- CTree::ucb_score (cnode.cpp:297-335), one child -------------------------------------------------------------------------
// n = parent.visit_count - 1 (clamped to the host tables); mixed float / double arithmetic exactly as the reference's toolchain
// resolves it: the log term and the prior coefficient come from host-built tables.
__device__ __forceinline__ float ucb_score_device(const TreeLayout &L, const float *__restrict__ logterm, const double *__restrict__ sqrtn,
                                                  int n, float discount, float pq, float mn, float mx, int mmc, float prior, int cvis,
                                                  float rew, float ws, float wt)
{
    // pb_c = log((n + c_base + 1)/c_base) + c_init   [float <- double]   (host table)
    // pb_c *= sqrt(n) / (visit + 1)                   [float <- double product]
    const float pb_c = (L.pbc_dim > 0) ? L.pbc_table[(size_t)n * L.pbc_dim + min(cvis, L.pbc_dim - 1)]
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
    return __fadd_rn(prior_score, v);
}

// ---- CTree_batch::cbatch_selection -> CTree::select_path / select_child (cnode.cpp:616-642, 381-413, 337-379) ----------------
// Sequential form: latency-bound pointer chase, ONE memory round trip per tree level.  While the children of the current node are
// scored (one lane per child), each child's own header is already in its lane's registers, so descending is a shuffle.  The
// next few mt19937 outputs are prefetched at the start (one raw draw per select_child call).
__device__ __forceinline__ void select_path_device(const TreeLayout &L, char *tb, TreeHdr *h, const TreeHot &hot, const float *__restrict__ logterm,
                                                   const double *__restrict__ sqrtn, int table_len, float discount, int tree,
                                                   int lane, int *__restrict__ idx_x, int *__restrict__ idx_y,
                                                   int *__restrict__ act_out, int *g_err)
{
    uint16_t *path = hot.path;
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
            if (lane < C) score = ucb_score_device(L, logterm, sqrtn, n, discount, pq, mn, mx, mmc, prior, cvis, rew, ws, wt);
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
        if (idx_y) idx_y[tree] = tree;
        if (err) {
            h->err = err;
            *g_err = err;
        }
    }
    const uint8_t *act = f_actions(L, tb) + (size_t)node * L.N;
    for (int j = lane; j < L.N; j += 32) act_out[(size_t)tree * L.N + j] = act[j];
}

// Wide form: one lane per EXPANDED node.  The child a node would pick depends only on that node's children, the tree's min-max
// bounds and the raw RNG word its select_child call consumes -- and that word's index is the node's depth (the path holds
// exactly one node per depth, each consuming one word, except a root still in its forced round-robin: cnode.cpp:398-399,
// 373-377).  So all choices are made at once (children scored sequentially inside the lane: literally the reference's loop,
// cnode.cpp:351-370), and the path is then a pointer chase through a shared-memory table.  Returns false, having changed
// nothing, when the words needed straddle the end of the generator's block (the sequential form regenerates it on the way).
// (out of line: rare, large; the sequential form and the backup stay inlined in their kernels)
static __device__ __noinline__ bool select_path_wide(const TreeLayout &L, char *tb, TreeHdr *h, const TreeHot &hot, const float *__restrict__ logterm,
                                                 const double *__restrict__ sqrtn, int table_len, float discount, int tree, int lane,
                                                 const StepScratch &scr, int *__restrict__ idx_x, int *__restrict__ idx_y,
                                                 int *__restrict__ act_out, int *g_err)
{
    const int n_exp = h->n_expanded, mt_pos = h->mt_pos;
    if (mt_pos + n_exp > kMtN) return false;                 // (a path visits at most n_exp expanded nodes)
    const float mn = h->mm_min, mx = h->mm_max;
    const int mmc = h->mm_cnt;
    const uint32_t *mt = f_mt(L, tb);
    const uint16_t *expslot = hot.expslot;
    const uint16_t *depth = hot.depth;
    const RecRegs root = rec_load(L, tb, 0);
    const int forced = (rec_visit(root) <= rec_nchild(root)) ? 1 : 0;
    // NPL nodes per lane and pass, children in batches of CB: every load of a batch (CB x 32 bytes per lane and node) is in flight
    // before the first score is computed -- a pass is a handful of L2 round trips, not one per child.  (NPL = 2 spills registers
    // next to the inference stages' 160+; NPL = 1 with 5-child batches does not.)
    constexpr int NPL = 1, CB = 5;
#pragma unroll 1
    for (int e0 = 0; e0 < n_exp; e0 += 32 * NPL) {
        int slot[NPL], C[NPL], base[NPL], n[NPL], dep[NPL];
        float pq[NPL], max_score[NPL];
        unsigned listmask[NPL];
        uint32_t word[NPL];
        bool on[NPL], draw[NPL];
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
            const int e = e0 + 32 * u + lane;
            on[u] = e < n_exp;
            slot[u] = on[u] ? expslot[e] : 0;
            dep[u] = on[u] ? depth[e] : 0;
        }
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
            const int e = e0 + 32 * u + lane;
            const RecRegs nr = rec_load(L, tb, slot[u]);
            C[u] = on[u] ? rec_nchild(nr) : 0;
            base[u] = rec_cbase(nr);
            const int vc = rec_visit(nr);
            pq[u] = rec_pred_value(nr);
            draw[u] = on[u] && !(e == 0 && forced);
            word[u] = draw[u] ? mt[mt_pos + dep[u] - forced] : 0u;
            n[u] = min(vc - 1, table_len - 1);
            max_score[u] = -1000000.0f;                      // FLOAT_MIN (utils.h:11-12)
            listmask[u] = 0u;
            if (on[u] && !draw[u]) listmask[u] = 1u << (vc - 1);   // forced root round-robin (no draw): the only candidate
        }
        int Cmax = 0;
#pragma unroll
        for (int u = 0; u < NPL; ++u) Cmax = max(Cmax, draw[u] ? C[u] : 0);
#pragma unroll 1
        for (int c0 = 0; c0 < Cmax; c0 += CB) {
            RecRegs r[NPL][CB];
#pragma unroll
            for (int u = 0; u < NPL; ++u)
#pragma unroll
                for (int j = 0; j < CB; ++j)
                    if (draw[u] && c0 + j < C[u]) r[u][j] = rec_load(L, tb, base[u] + c0 + j);
#pragma unroll
            for (int u = 0; u < NPL; ++u)
#pragma unroll
                for (int j = 0; j < CB; ++j)
                    if (draw[u] && c0 + j < C[u]) {          // the reference's loop body (cnode.cpp:354-369), child c0 + j
                        const float sc = ucb_score_device(L, logterm, sqrtn, n[u], discount, pq[u], mn, mx, mmc, rec_prior(r[u][j]),
                                                          rec_visit(r[u][j]), rec_reward(r[u][j]), rec_wsum(r[u][j]), rec_wtot(r[u][j]));
                        if (max_score[u] < sc) {
                            max_score[u] = sc;
                            listmask[u] = 1u << (c0 + j);
                        } else if (sc >= __fsub_rn(max_score[u], 0.000001f)) {
                            listmask[u] |= 1u << (c0 + j);
                        }
                    }
        }
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
            if (!on[u]) continue;
            int ci = 0;
            const int nl = __popc(listmask[u]);
            if (nl > 0) {
                uint32_t mrem = listmask[u];
                if (draw[u])                                 // one raw word per select_child call, even for a single candidate
                    for (uint32_t skip = mt_temper(word[u]) % (uint32_t)nl; skip > 0; --skip) mrem &= mrem - 1;
                ci = __ffs(mrem) - 1;
            }
            const int cs = base[u] + ci;
            const RecRegs cr = rec_load(L, tb, cs);          // (just scored: L1)
            const int ceid = (rec_nchild(cr) > 0) ? rec_eid(cr) + 1 : 0;      // expanded() = num_children > 0
            scr.chase[e0 + 32 * u + lane] = ((uint32_t)cs << 16) | (uint32_t)ceid;
        }
    }
    __syncwarp();
    // the path: root -> chosen child -> ... until an unexpanded node
    uint16_t *path = hot.path;
    int e = 0, len = 0, node = 0, parent = 0, err = 0;
    if (lane == 0) path[0] = 0;
    while (true) {
        const uint32_t w = scr.chase[e];                     // (same address in every lane: broadcast)
        parent = node;
        node = (int)(w >> 16);
        ++len;
        if (len > L.S + 1) {
            err = kErrPathOverflow;
            break;
        }
        if (lane == 0) path[len] = (uint16_t)node;
        const int ne = (int)(w & 0xffffu);
        if (ne == 0) break;
        e = ne - 1;
    }
    if (lane == 0) {
        h->path_len = len;
        h->mt_pos = mt_pos + len - forced;                   // one word per select_child call
        h->sum_path_len += len;
        idx_x[tree] = f_hidx(L, tb)[parent];                 // hidden_state_index_x of the leaf's parent
        if (idx_y) idx_y[tree] = tree;
        if (err) {
            h->err = err;
            *g_err = err;
        }
    }
    const uint8_t *act = f_actions(L, tb) + (size_t)node * L.N;
    for (int j = lane; j < L.N; j += 32) act_out[(size_t)tree * L.N + j] = act[j];
    return true;
}

// the selection of the next simulation: wide when the last path was long (scr may be NULL: sequential only)
__device__ __forceinline__ void select_next_device(const TreeLayout &L, char *tb, TreeHdr *h, const TreeHot &hot, const float *__restrict__ logterm,
                                                   const double *__restrict__ sqrtn, int table_len, float discount, int tree,
                                                   int lane, const StepScratch *scr, int *__restrict__ idx_x, int *__restrict__ idx_y,
                                                   int *__restrict__ act_out, int *g_err)
{
    if (scr != nullptr && h->path_len >= L.wide_min_len &&
        select_path_wide(L, tb, h, hot, logterm, sqrtn, table_len, discount, tree, lane, *scr, idx_x, idx_y, act_out, g_err))
        return;
    select_path_device(L, tb, h, hot, logterm, sqrtn, table_len, discount, tree, lane, idx_x, idx_y, act_out, g_err);
}

// ---- CTree::back_propagate (cnode.cpp:415-450), shared by every kernel --------------------------------------------------
// The reference walks the path leaf -> root; per node: visit++, subtree_info.update(G, depth), q-delta entry, G = r + gamma G.
// Only the scalar recurrence of G is sequential.  The value-set update of a path node touches ONLY its own (slot, depth) set --
// entries appended / flags flipped for one node never match another node's tag, and the append positions are known up front
// (log_len0 + distance from the leaf) -- so all levels are updated in parallel, one lane per level, from per-level
// accumulators filled by ONE pass over the value log.
constexpr int kLogRegs = 8;   // 8 x 32 log entries requested up front (before the expansion); longer logs: tail read in the pass

struct LogRegs {
    uint32_t k[kLogRegs];
    float v[kLogRegs];
};
__device__ __forceinline__ void log_prefetch(const TreeLayout &L, char *tb, int log_len0, int lane, LogRegs &lr)
{
    const uint32_t *vk = f_vskey(L, tb);
    const float *vv = f_vsval(L, tb);
#pragma unroll
    for (int c = 0; c < kLogRegs; ++c) {
        const int e = c * 32 + lane;
        lr.k[c] = (e < log_len0) ? vk[e] : 0xffffffffu;   // depth 0x7fff never matches a path level
        lr.v[c] = (e < log_len0) ? vv[e] : 0.0f;
    }
}
// one log entry -> the accumulators of the path level it belongs to (if any)
__device__ __forceinline__ void backup_scatter_entry(const StepScratch &scr, int len, uint32_t k, float v, int e)
{
    const int i = len - (int)((k >> 1) & 0x7fffu);        // the path node at distance `depth` from the leaf
    if (i >= 0 && scr.pathc[i] == (uint16_t)(k >> 16)) {
        LevelAcc *a = scr.acc + i;
        const unsigned long long oe = ((unsigned long long)f2ord(v) << 32) | (unsigned int)e;
        atomicAdd(&a->cnt, 1u);
        if (k & 1u) {
            atomicAdd(&a->nbig, 1u);
            atomicMin(&a->minb, oe);
        } else {
            atomicMax(&a->maxs, oe);
        }
    }
}

// leaf -> root.  `len` = SearchResult::search_len, path[0..len] the path's node slots (global memory); `leaf_eid` = expansion
// order the leaf gets in this simulation; reward_in / value = this simulation's network outputs; `lr` = the first 256 entries of
// the log as they were BEFORE this backup.  Writes visit / wsum / wtot of every path node (the leaf's included), the q-delta
// entries and the log appends.  Warp-uniform log_len / err.
__device__ __forceinline__ void backup_parallel(const TreeLayout &L, char *tb, const TreeHot &hot, const float *__restrict__ lam_pow,
                                                const StepScratch &scr, const LogRegs &lr, int len, int log_len0, int leaf_eid,
                                                float reward_in, float value, float discount, int lane, int &log_len, int &err)
{
    uint32_t *vk = f_vskey(L, tb);
    float *vv = f_vsval(L, tb);
    float *qd = hot.qd;                               // q-delta of the e-th expanded node (the CMinMaxStats entries)
    const uint16_t *path = hot.path;
    // ---- pass over the log: every entry is looked at once -----------------------------------------------------------------
    for (int i = lane; i <= len; i += 32) {
        scr.pathc[i] = path[i];
        scr.acc[i].cnt = 0u; scr.acc[i].nbig = 0u; scr.acc[i].minb = ~0ull; scr.acc[i].maxs = 0ull;
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < kLogRegs; ++c) backup_scatter_entry(scr, len, lr.k[c], lr.v[c], c * 32 + lane);
    for (int e0 = kLogRegs * 32; e0 < log_len0; e0 += 128) {         // tail: four independent loads per lane in flight
        uint32_t k[4];
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + 32 * u + lane;
            k[u] = (e < log_len0) ? vk[e] : 0xffffffffu;
            v[u] = (e < log_len0) ? vv[e] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) backup_scatter_entry(scr, len, k[u], v[u], e0 + 32 * u + lane);
    }
    __syncwarp();
    // ---- one lane per path level, chunks of 32 levels, leaf first -----------------------------------------------------------
    float G = value;
#pragma unroll 1
    for (int base = len; base >= 0; base -= 32) {
        const int nlev = min(32, base + 1);
        const bool active = lane < nlev;
        const int i = base - lane;                    // path index of my level
        int slot = 0, vis = 0, nh = 0;
        float rew = 0.f, ws = 0.f, wt = 0.f, ppv = 0.f, lp = 0.f;
        if (active) {
            slot = scr.pathc[i];
            lp = lam_pow[len - i];
            if (i == len) {                           // the freshly expanded leaf
                rew = reward_in; nh = leaf_eid;
            } else {
                const RecRegs q = rec_load(L, tb, slot);
                rew = rec_reward(q); ws = rec_wsum(q); wt = rec_wtot(q); vis = rec_visit(q); nh = rec_eid(q);
            }
            if (i > 0) ppv = f_pred_value(L, tb)[scr.pathc[i - 1]];     // parent's pred_value
        }
        // G of every level: the reference's sequential recurrence, in its order
        float myG = 0.f;
        for (int l = 0; l < nlev; ++l) {
            const float r_l = __shfl_sync(MAZ_FULL, rew, l);
            if (lane == l) myG = G;
            G = __fadd_rn(r_l, __fmul_rn(discount, G));
        }
        // (the reference removes the node's old q-delta from the min-max multiset first; here the entry simply lives in
        //  qd[expansion order] and is overwritten below)
        int myerr = 0;
        if (active) {
            const LevelAcc a = scr.acc[i];
            const uint32_t tag = vs_tag(slot, len - i);
            const VsDecision d = vs_decide(L, (int)a.cnt, (int)a.nbig, (uint32_t)(a.minb >> 32), (int)(uint32_t)a.minb,
                                           (uint32_t)(a.maxs >> 32), (int)(uint32_t)a.maxs, tag, lp, myG, ws, wt);
            myerr = d.err;
            const int pos = log_len + lane;           // the reference appends leaf first
            if (pos >= L.L) {
                myerr = kErrLogOverflow;
            } else {
                if (d.flip_pos >= 0) vk[d.flip_pos] = d.flip_key;
                vk[pos] = tag | (d.append_big ? 1u : 0u);
                vv[pos] = myG;
            }
            f_visit(L, tb)[slot] = vis + 1;
            f_wsum(L, tb)[slot] = ws;
            f_wtot(L, tb)[slot] = wt;
            if (i != 0) qd[nh] = __fsub_rn(__fadd_rn(rew, __fmul_rn(discount, __fdiv_rn(ws, wt))), ppv);
        }
        const unsigned bad = __ballot_sync(MAZ_FULL, myerr != 0);
        if (bad) err = __shfl_sync(MAZ_FULL, myerr, __ffs(bad) - 1);
        log_len += nlev;
    }
    __syncwarp();
}

// CMinMaxStats min / max = reduction over the q-deltas of all visited (= expanded) non-root nodes [1, n_expanded)
__device__ __forceinline__ void minmax_reduce(const TreeHot &hot, int n_expanded, int lane, float &mn, float &mx)
{
    const float *qd = hot.qd;
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int e = 1 + lane; e < n_expanded; e += 32) {
        const uint32_t o = f2ord(qd[e]);
        lo = min(lo, o);
        hi = max(hi, o);
    }
    mn = ord2f(__reduce_min_sync(MAZ_FULL, lo));
    mx = ord2f(__reduce_max_sync(MAZ_FULL, hi));
}

// ---- CTree_batch::cbatch_expansion_and_backup -> expand_and_backprop (cnode.cpp:644-670, 452-469), ONE warp per tree ------
// Latency plan: everything that depends only on the tree's own state (header, log snapshot, the expansion's random words) is
// requested before the network outputs are needed.
template <bool kCg = true>
__device__ __forceinline__ void expand_backup_device(const TreeLayout &L, char *tb, TreeHdr *h, const TreeHot &hot, const float *__restrict__ lam_pow,
                                                     int hidx, float discount, int K, const float *__restrict__ reward_ptr,
                                                     const float *__restrict__ value_ptr,
                                                     const float *__restrict__ probs, const float *__restrict__ beta,
                                                     const StepScratch &scr, int lane, int *g_err, int tree = -1)
{
    MAZ_TS(L, tree, lane, 0);
    griddep_launch();   // PDL: the next kernel (inference of the next simulation) may start its prologue now
    int tot_nodes = h->tot_nodes, log_len = h->log_len, mt_pos = h->mt_pos, n_expanded = h->n_expanded, err = h->err;
    const int len = h->path_len;
    const int log_len0 = log_len;
    const int leaf_eid = n_expanded;                  // expansion order the leaf is about to get
    MAZ_TS(L, tree, lane, 1);
    LogRegs lr;
    log_prefetch(L, tb, log_len0, lane, lr);
    const int n_draw = (L.A >= 2) ? 2 * K * L.N : 0;
    const bool draws_pre = n_draw > 0 && n_draw <= kMtChunk && mt_pos + n_draw <= kMtN;
    if (draws_pre) {
        const uint32_t *mt = f_mt(L, tb);
        for (int t = lane; t < n_draw; t += 32) scr.ex.draws[t] = mt_temper(mt[mt_pos + t]);
        mt_pos += n_draw;
    }
    const int leaf = hot.path[len];
    MAZ_TS(L, tree, lane, 2);
    // PDL: everything above only touched this tree's own state (written by the previous tree kernel, long
    // complete); the network outputs of THIS simulation are produced by the kernel we may be overlapping with.
    griddep_wait();
    const float reward_in = ld_in<kCg>(reward_ptr), value = ld_in<kCg>(value_ptr);   // (L2 loads, see expand_node)
    expand_node<kCg>(L, tb, hot, tot_nodes, n_expanded, mt_pos, err, leaf, hidx, reward_in, value, probs, beta, K, 0.0f, nullptr, scr.ex,
                     lane, draws_pre, tree, /*init_stats=*/true, /*depth=*/len);
    MAZ_TS(L, tree, lane, 3);
    backup_parallel(L, tb, hot, lam_pow, scr, lr, len, log_len0, leaf_eid, reward_in, value, discount, lane, log_len, err);
    MAZ_TS(L, tree, lane, 4);
    float mn, mx;
    minmax_reduce(hot, n_expanded, lane, mn, mx);
    if (lane == 0) {
        h->tot_nodes = tot_nodes;
        h->log_len = log_len;
        h->mt_pos = mt_pos;
        h->n_expanded = n_expanded;
        h->mm_cnt = n_expanded - 1;
        h->mm_min = mn;
        h->mm_max = mx;
        h->err = err;
        if (err) *g_err = err;
    }
    MAZ_TS(L, tree, lane, 5);
}

}  // namespace maz
