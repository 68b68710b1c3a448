// tree_step.cuh -- the per-tree step functions of the sampled-MCTS tree engine (one warp = one tree), shared by the
// stand-alone tree kernels (tree_kernels.cuh) and the persistent whole-search kernel (search_persist.cuh).
#pragma once
#include "tree_device.cuh"

namespace maz {

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

// ---- CTree::back_propagate (cnode.cpp:415-450), shared by every kernel --------------------------------------------------
// The reference walks the path leaf -> root; per node: visit++, subtree_info.update(G, depth), q-delta entry, G = r + gamma G.
// Only the scalar recurrence of G is sequential.  The value-set update of a path node touches ONLY its own (slot, depth) set --
// entries appended / flags flipped for one node never match another node's tag, and the append positions are known up front
// (log_len0 + distance from the leaf) -- so all levels are processed IN PARALLEL: a group of 32 / pow2(levels) lanes per level
// scans a snapshot of the tree's value log for that level's set, combines, and its leader applies the reference's exact fp32
// update (vs_decide).  Cost: one pass over the log, whatever the depth (the sequential form cost ~1.2 k cycles per level and
// made the deepest tree of a batch the pace-setter of every simulation).
constexpr int kLogCache = 256;    // log entries staged in shared memory per tree; a longer log's tail is read from global memory
__host__ __device__ inline size_t backup_scratch_bytes() { return (size_t)kLogCache * 8; }

// per-warp shared-memory scratch of a tree step: the expansion's staging + the log cache
__host__ __device__ inline size_t tree_scratch_bytes(int N, int A, int K) { return expand_scratch_bytes(N, A, K) + backup_scratch_bytes(); }

struct LogCache {
    uint32_t *k;        // [kLogCache] keys   (shared memory)
    float *v;           // [kLogCache] values
};
__device__ __forceinline__ LogCache carve_log_cache(char *p)
{
    return LogCache{reinterpret_cast<uint32_t *>(p), reinterpret_cast<float *>(p + 4 * kLogCache)};
}
// stage entries [from, min(to, kLogCache)) of the tree's value log
__device__ __forceinline__ void log_cache_fill(const TreeLayout &L, char *tb, const LogCache &lc, int from, int to, int lane)
{
    const uint32_t *vk = f_vskey(L, tb);
    const float *vv = f_vsval(L, tb);
    const int n = min(to, kLogCache);
    for (int e = from + lane; e < n; e += 32) {
        lc.k[e] = vk[e];
        lc.v[e] = vv[e];
    }
}

// leaf -> root.  `len` = SearchResult::search_len, path[0..len] the path's node slots; `leaf_eid` = expansion order the leaf gets
// in this simulation; reward_in / value = this simulation's network outputs; the log cache holds entries [0, min(log_len0,
// kLogCache)) as they were BEFORE this backup.  Writes visit / wsum / wtot of every path node (the leaf's included), the
// q-delta entries and the log appends (global memory, and the cache when `keep_cache`).  Warp-uniform log_len / err.
__device__ __forceinline__ void backup_parallel(const TreeLayout &L, char *tb, const float *__restrict__ lam_pow, const LogCache &lc,
                                                const uint16_t *__restrict__ path, int len, int log_len0, int leaf_eid,
                                                float reward_in, float value, float discount, int lane, int &log_len, int &err,
                                                bool keep_cache = false)
{
    uint32_t *vk = f_vskey(L, tb);
    float *vv = f_vsval(L, tb);
    float *qd = f_qdelta(L, tb);                      // q-delta of the e-th expanded node (the CMinMaxStats entries)
    const int ncache = min(log_len0, kLogCache);
    float G = value;
#pragma unroll 1
    for (int base = len; base >= 0; base -= 32) {     // chunks of up to 32 path levels, leaf first
        const int nlev = min(32, base + 1);
        int gl = 1;
        while (gl < nlev) gl <<= 1;                   // levels rounded up to a power of two
        const int gs = 32 / gl;                       // lanes per level
        const int lev = lane / gs, sub = lane - lev * gs;
        const bool active = lev < nlev;
        const int i = base - lev;                     // path index of my level
        int slot = 0, vis = 0, nh = 0;
        float rew = 0.f, ws = 0.f, wt = 0.f, ppv = 0.f, lp = 0.f;
        if (active) {
            slot = path[i];
            lp = lam_pow[len - i];
            if (i == len) {                           // the freshly expanded leaf
                rew = reward_in; nh = leaf_eid;
            } else {
                const RecRegs q = rec_load(L, tb, slot);
                rew = rec_reward(q); ws = rec_wsum(q); wt = rec_wtot(q); vis = rec_visit(q); nh = rec_eid(q);
            }
            if (i > 0) ppv = f_pred_value(L, tb)[path[i - 1]];     // parent's pred_value
        }
        // G of every level: the reference's sequential recurrence, in its order
        float myG = 0.f;
        for (int l = 0; l < nlev; ++l) {
            const float r_l = __shfl_sync(MAZ_FULL, rew, l * gs);
            if (lev == l) myG = G;
            G = __fadd_rn(r_l, __fmul_rn(discount, G));
        }
        // (the reference removes the node's old q-delta from the min-max multiset first; here the entry simply lives in
        //  qd[expansion order] and is overwritten below)
        const uint32_t tag = vs_tag(slot, len - i);
        VsScan r;
        vs_scan_init(r);
        if (active) {
            for (int e = sub; e < ncache; e += gs) vs_scan_entry(r, tag, lc.k[e], lc.v[e], e);
            for (int e = kLogCache + sub; e < log_len0; e += gs) vs_scan_entry(r, tag, vk[e], vv[e], e);
        }
        for (int o = 1; o < gs; o <<= 1) {            // combine the group's partial scans
            const int ocnt = __shfl_xor_sync(MAZ_FULL, r.cnt, o), onbig = __shfl_xor_sync(MAZ_FULL, r.nbig, o);
            const uint32_t omin = __shfl_xor_sync(MAZ_FULL, r.minbig, o), omax = __shfl_xor_sync(MAZ_FULL, r.maxsmall, o);
            const int ominpos = __shfl_xor_sync(MAZ_FULL, r.minpos, o), omaxpos = __shfl_xor_sync(MAZ_FULL, r.maxpos, o);
            r.cnt += ocnt; r.nbig += onbig;
            if (ominpos >= 0 && (r.minpos < 0 || omin < r.minbig)) { r.minbig = omin; r.minpos = ominpos; }
            if (omaxpos >= 0 && (r.maxpos < 0 || omax > r.maxsmall)) { r.maxsmall = omax; r.maxpos = omaxpos; }
        }
        int myerr = 0;
        if (active && sub == 0) {
            const VsDecision d = vs_decide(L, r.cnt, r.nbig, r.minbig, r.minpos, r.maxsmall, r.maxpos, tag, lp, myG, ws, wt);
            myerr = d.err;
            const int pos = log_len + lev;            // the reference appends leaf first
            if (pos >= L.L) {
                myerr = kErrLogOverflow;
            } else {
                const uint32_t key = tag | (d.append_big ? 1u : 0u);
                if (d.flip_pos >= 0) vk[d.flip_pos] = d.flip_key;
                vk[pos] = key;
                vv[pos] = myG;
                if (keep_cache) {
                    if (d.flip_pos >= 0 && d.flip_pos < kLogCache) lc.k[d.flip_pos] = d.flip_key;
                    if (pos < kLogCache) { lc.k[pos] = key; lc.v[pos] = myG; }
                }
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
__device__ __forceinline__ void minmax_reduce(const TreeLayout &L, char *tb, int n_expanded, int lane, float &mn, float &mx)
{
    const float *qd = f_qdelta(L, tb);
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
__device__ __forceinline__ void expand_backup_device(const TreeLayout &L, char *tb, TreeHdr *h, const float *__restrict__ lam_pow,
                                                     int hidx, float discount, int K, const float *__restrict__ reward_ptr,
                                                     const float *__restrict__ value_ptr,
                                                     const float *__restrict__ probs, const float *__restrict__ beta,
                                                     const ExpandScratch &sc, const LogCache &lc, int lane, int *g_err, int tree = -1)
{
    MAZ_TS(L, tree, lane, 0);
    griddep_launch();   // PDL: the next kernel (inference of the next simulation) may start its prologue now
    int tot_nodes = h->tot_nodes, log_len = h->log_len, mt_pos = h->mt_pos, n_expanded = h->n_expanded, err = h->err;
    const int len = h->path_len;
    const int log_len0 = log_len;
    const int leaf_eid = n_expanded;                  // expansion order the leaf is about to get
    MAZ_TS(L, tree, lane, 1);
    log_cache_fill(L, tb, lc, 0, log_len0, lane);
    const int n_draw = (L.A >= 2) ? 2 * K * L.N : 0;
    const bool draws_pre = n_draw > 0 && n_draw <= kMtChunk && mt_pos + n_draw <= kMtN;
    if (draws_pre) {
        const uint32_t *mt = f_mt(L, tb);
        for (int t = lane; t < n_draw; t += 32) sc.draws[t] = mt_temper(mt[mt_pos + t]);
        mt_pos += n_draw;
    }
    const uint16_t *path = f_path(L, tb);
    const int leaf = path[len];
    MAZ_TS(L, tree, lane, 2);
    // PDL: everything above only touched this tree's own state (written by the previous tree kernel, long
    // complete); the network outputs of THIS simulation are produced by the kernel we may be overlapping with.
    griddep_wait();
    const float reward_in = __ldcg(reward_ptr), value = __ldcg(value_ptr);   // L2 loads, see expand_node
    expand_node(L, tb, tot_nodes, n_expanded, mt_pos, err, leaf, hidx, reward_in, value, probs, beta, K, 0.0f, nullptr, sc,
                lane, draws_pre, tree);
    MAZ_TS(L, tree, lane, 3);
    backup_parallel(L, tb, lam_pow, lc, path, len, log_len0, leaf_eid, reward_in, value, discount, lane, log_len, err);
    MAZ_TS(L, tree, lane, 4);
    float mn, mx;
    minmax_reduce(L, tb, n_expanded, lane, mn, mx);
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
