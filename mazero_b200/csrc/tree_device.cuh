// tree_device.cuh -- device-side building blocks of the sampled-MCTS tree engine (sm_100a).
//
// Execution model: ONE WARP OWNS ONE TREE for the whole kernel.  No atomics, no cross-warp sharing.
// Lanes are used for (a) the children of the node being scored / created, (b) the (sample k, agent i)
// pairs of the joint-action draw, (c) scanning the tree's value log, (d) the 624-word mt19937 twist.
//
// Bit-exactness: every fp32/fp64 operation that the reference performs is issued here with an explicit
// round-to-nearest intrinsic (__fmul_rn/__fadd_rn/__ddiv_rn ...) so that nvcc cannot contract a*b+c
// into an FMA (the reference is built for baseline x86-64: no FMA; core/mcts/ctree/setup.py:34).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tree_layout.h"

#include <cstddef>

namespace maz {

#define MAZ_FULL 0xffffffffu

// ---- slab field accessors --------------------------------------------------------------------------
#define MAZ_FIELD(type, name, off)                                                                      \
    __device__ __forceinline__ type *name(const TreeLayout &L, char *tb) { return reinterpret_cast<type *>(tb + L.off); }
MAZ_FIELD(uint32_t, f_mt, off_mt)
MAZ_FIELD(NodeRec, f_rec, off_rec)
MAZ_FIELD(float, f_pred_prob, off_pred_prob)
MAZ_FIELD(float, f_beta, off_beta)
MAZ_FIELD(float, f_beta_hat, off_beta_hat)
MAZ_FIELD(float, f_qdelta, off_qdelta)
MAZ_FIELD(uint8_t, f_actions, off_actions)
MAZ_FIELD(uint16_t, f_expslot, off_expslot)
MAZ_FIELD(uint16_t, f_path, off_path)
MAZ_FIELD(uint32_t, f_vskey, off_vskey)
MAZ_FIELD(float, f_vsval, off_vsval)
MAZ_FIELD(uint16_t, f_depth, off_depth)
#undef MAZ_FIELD

// one field of the node records, indexed by slot (the cold paths keep the array notation; the hot paths load whole records)
template <typename T, size_t OFF>
struct RecField {
    char *base;
    __device__ __forceinline__ T &operator[](int slot) const { return *reinterpret_cast<T *>(base + (size_t)slot * sizeof(NodeRec) + OFF); }
};
#define MAZ_RECFIELD(type, name, member)                                                    \
    __device__ __forceinline__ RecField<type, offsetof(NodeRec, member)> name(const TreeLayout &L, char *tb) \
    {                                                                                       \
        return RecField<type, offsetof(NodeRec, member)>{tb + L.off_rec};                   \
    }
MAZ_RECFIELD(float, f_prior, prior)
MAZ_RECFIELD(float, f_reward, reward)
MAZ_RECFIELD(float, f_pred_value, pred_value)
MAZ_RECFIELD(float, f_wsum, wsum)
MAZ_RECFIELD(float, f_wtot, wtot)
MAZ_RECFIELD(int, f_visit, visit)
MAZ_RECFIELD(uint16_t, f_nchild, nchild)
MAZ_RECFIELD(uint16_t, f_cbase, cbase)
MAZ_RECFIELD(int16_t, f_hidx, hidx)
MAZ_RECFIELD(uint16_t, f_eid, eid)
#undef MAZ_RECFIELD

// whole-record access: two 16-byte transactions
struct RecRegs { uint4 a, b; };
__device__ __forceinline__ RecRegs rec_load(const TreeLayout &L, char *tb, int slot)
{
    const uint4 *p = reinterpret_cast<const uint4 *>(tb + L.off_rec + (size_t)slot * sizeof(NodeRec));
    RecRegs r;
    r.a = p[0];
    r.b = p[1];
    return r;
}
__device__ __forceinline__ float rec_prior(const RecRegs &r) { return __uint_as_float(r.a.x); }
__device__ __forceinline__ float rec_reward(const RecRegs &r) { return __uint_as_float(r.a.y); }
__device__ __forceinline__ float rec_pred_value(const RecRegs &r) { return __uint_as_float(r.a.z); }
__device__ __forceinline__ float rec_wsum(const RecRegs &r) { return __uint_as_float(r.a.w); }
__device__ __forceinline__ float rec_wtot(const RecRegs &r) { return __uint_as_float(r.b.x); }
__device__ __forceinline__ int rec_visit(const RecRegs &r) { return (int)r.b.y; }
__device__ __forceinline__ int rec_nchild(const RecRegs &r) { return (int)(r.b.z & 0xffffu); }
__device__ __forceinline__ int rec_cbase(const RecRegs &r) { return (int)(r.b.z >> 16); }
__device__ __forceinline__ int rec_hidx(const RecRegs &r) { return (int)(int16_t)(r.b.w & 0xffffu); }
__device__ __forceinline__ int rec_eid(const RecRegs &r) { return (int)(r.b.w >> 16); }
// a freshly created child (CNode ctor, cnode.cpp:14-16): prior set, everything else zero, hidden index -1
__device__ __forceinline__ void rec_store_new_child(const TreeLayout &L, char *tb, int slot, float prior)
{
    uint4 *p = reinterpret_cast<uint4 *>(tb + L.off_rec + (size_t)slot * sizeof(NodeRec));
    p[0] = make_uint4(__float_as_uint(prior), 0u, 0u, 0u);      // prior, reward 0, pred_value 0, wsum 0
    p[1] = make_uint4(0u, 0u, 0u, 0x0000ffffu);                  // wtot 0, visit 0, nchild 0 | cbase 0, hidx -1 | eid 0
}

__device__ __forceinline__ TreeHdr *f_hdr(char *tb) { return reinterpret_cast<TreeHdr *>(tb); }

struct TreeHot;
// The tree's small, hot bookkeeping arrays.  The stand-alone kernels use them where they live in the slab (global memory); the
// persistent whole-search kernel keeps a copy of them -- and of the header -- in shared memory for the whole search, because the
// warp that owns the tree touches them in every phase of every simulation (each access was an L2 round trip on the critical path).
struct TreeHot {
    uint16_t *path;      // [S + 2] SearchResult::search_path (node slots)
    float *qd;           // [S + 2] q-delta entries of the expanded nodes (CMinMaxStats)
    uint16_t *expslot;   // [S + 2] slot of the e-th expanded node
    uint16_t *depth;     // [S + 2] its depth
};
__device__ __forceinline__ TreeHot hot_from_slab(const TreeLayout &L, char *tb)
{
    return TreeHot{f_path(L, tb), f_qdelta(L, tb), f_expslot(L, tb), f_depth(L, tb)};
}

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may start while its predecessor in the stream is still running; griddep_wait() blocks until the predecessor has
// completed and its writes are visible (a no-op without the attribute); griddep_launch() lets the successor start.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// profiling: phase timestamps of tree 0 (lane 0) when TreeLayout::dbg_clock is set
#define MAZ_TS(L, tree, lane, slot)                                                          \
    do {                                                                                     \
        if ((L).dbg_clock && (tree) == 0 && (lane) == 0) (L).dbg_clock[(slot)] = clock64(); \
    } while (0)

// order-preserving float <-> uint map (for redux.sync min/max on floats)
__device__ __forceinline__ uint32_t f2ord(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u)
{
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(b);
}

// ---- std::mt19937 -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mt_temper(uint32_t z)
{
    z ^= z >> 11;
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= z >> 18;
    return z;
}

// Regenerate the 624-word block in registers: lane l holds words 32c+l (c = 0..19).  Word i needs
// word i+1 (old) and word i+397 (old, i < 227) or i-227 (new, i >= 227): all reachable with shuffles
// whose source REGISTER index is a compile-time constant once the c-loop is unrolled.
// (not inlined: it runs once per 624 draws, and every inlined copy is ~4 KB of code on the hot path's instruction stream)
static __device__ __noinline__ void mt_regen(uint32_t *s, int lane)
{
    uint32_t r[20];
#pragma unroll
    for (int c = 0; c < 20; ++c) {
        int i = c * 32 + lane;
        r[c] = (i < kMtN) ? s[i] : 0u;
    }
#pragma unroll
    for (int c = 0; c < 20; ++c) {
        const int i = c * 32 + lane;
        // s[i+1]: same register one lane up, or lane 0 of the next register; i = 623 wraps to the NEW s[0]
        uint32_t nxt = __shfl_down_sync(MAZ_FULL, r[c], 1);
        uint32_t nxt_next = __shfl_sync(MAZ_FULL, r[(c < 19) ? c + 1 : 0], 0);
        if (lane == 31 || (c == 19 && lane == 15)) nxt = nxt_next;
        // s[i+397] = register c+12 (+1 when lane+13 wraps), lane (lane+13)&31      [i < 227]
        // s[i-227] = register c-8  (+1 when lane+29 wraps), lane (lane+29)&31      [i >= 227]
        uint32_t src = 0;
        if (c <= 7) {
            uint32_t a_lo = __shfl_sync(MAZ_FULL, r[(c + 12 <= 19) ? c + 12 : 19], (lane + 13) & 31);
            uint32_t a_hi = __shfl_sync(MAZ_FULL, r[(c + 13 <= 19) ? c + 13 : 19], (lane + 13) & 31);
            src = (lane < 19) ? a_lo : a_hi;
        }
        if (c >= 7) {
            uint32_t b_lo = __shfl_sync(MAZ_FULL, r[(c >= 8) ? c - 8 : 0], (lane + 29) & 31);
            uint32_t b_hi = __shfl_sync(MAZ_FULL, r[c - 7], (lane + 29) & 31);
            uint32_t b = (lane < 3) ? b_lo : b_hi;
            if (i >= 227) src = b;
        }
        uint32_t y = (r[c] & 0x80000000u) | (nxt & 0x7fffffffu);
        uint32_t v = src ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        if (i < kMtN) r[c] = v;
    }
#pragma unroll
    for (int c = 0; c < 20; ++c) {
        int i = c * 32 + lane;
        if (i < kMtN) s[i] = r[c];
    }
}

// Stage the next n (<= kMtChunk) tempered outputs of the tree's generator into out[] (shared memory).
// mt_pos is the warp-uniform stream position (kept in a register by the caller).
__device__ __forceinline__ void mt_fetch(uint32_t *s, int &mt_pos, uint32_t *out, int n, int lane)
{
    const int p = mt_pos;
    for (int t = lane; t < n; t += 32) {
        int q = p + t;
        if (q < kMtN) out[t] = mt_temper(s[q]);
    }
    if (p + n > kMtN) {
        __syncwarp();
        mt_regen(s, lane);
        __syncwarp();
        for (int t = lane; t < n; t += 32) {
            int q = p + t;
            if (q >= kMtN) out[t] = mt_temper(s[q - kMtN]);
        }
        mt_pos = p + n - kMtN;
    } else {
        mt_pos = p + n;
    }
    __syncwarp();
}

// One raw draw (std::mt19937::operator()), warp-uniform result.
__device__ __forceinline__ uint32_t mt_next(uint32_t *s, int &mt_pos, int lane)
{
    if (mt_pos >= kMtN) {
        __syncwarp();
        mt_regen(s, lane);
        __syncwarp();
        mt_pos = 0;
    }
    uint32_t z = mt_temper(s[mt_pos]);  // same address in every lane: one broadcast transaction
    mt_pos += 1;
    return z;
}

// std::generate_canonical<double,53>(mt19937) from two consecutive 32-bit draws (random.tcc:3362-3378)
__device__ __forceinline__ double mt_canonical(uint32_t x1, uint32_t x2)
{
    double sum = __dadd_rn((double)x1, __dmul_rn((double)x2, 4294967296.0));
    double r = __dmul_rn(sum, 5.42101086242752217e-20);  // exact: * 2^-64
    if (r >= 1.0) r = 0.99999999999999988897769753748;   // nextafter(1.0, 0.0)
    return r;
}

// ---- per-warp shared-memory scratch of the expansion ----------------------------------------------
struct ExpandScratch {
    double *cp;         // [N*A] cumulative probabilities (std::discrete_distribution::_M_cp)
    long long *keys;    // [32]
    float *beta;        // [N*A]
    float *probs;       // [N*A] policy probabilities (staged with beta: no dependent global loads later)
    uint32_t *draws;    // [kMtChunk]
    uint8_t *samp;      // [K*N] sampled per-agent actions
};
__host__ __device__ inline size_t expand_scratch_bytes(int N, int A, int K)
{
    size_t na = (size_t)N * A;
    size_t b = 8 * na + 8 * 32 + 4 * na + 4 * na + 4 * kMtChunk + (size_t)K * N;
    return (b + 15) & ~(size_t)15;
}
__device__ __forceinline__ ExpandScratch carve_scratch(char *p, int N, int A)
{
    ExpandScratch s;
    size_t na = (size_t)N * A;
    s.cp = reinterpret_cast<double *>(p);
    s.keys = reinterpret_cast<long long *>(p + 8 * na);
    s.beta = reinterpret_cast<float *>(p + 8 * na + 256);
    s.probs = reinterpret_cast<float *>(p + 8 * na + 256 + 4 * na);
    s.draws = reinterpret_cast<uint32_t *>(p + 8 * na + 256 + 8 * na);
    s.samp = reinterpret_cast<uint8_t *>(p + 8 * na + 256 + 8 * na + 4 * kMtChunk);
    return s;
}

// ---- CTree::expand (cnode.cpp:224-295) --------------------------------------------------------------
// Samples K joint actions from the factorised per-agent distribution beta, merges duplicates, creates
// the children in ascending (wrapped, signed 64-bit) key order in consecutive slots.  Returns the number
// of children.  `probs`, `beta`, `noises` point at this tree's (N,A) rows in global memory.
// kCg: read the network outputs with ld.global.cg (L2): they are produced by another kernel that, under programmatic dependent
// launch, may still be running when this one becomes resident (stale L1 lines of the re-used buffers).  false: plain generic loads
// (the persistent kernel hands them over in shared memory).
template <bool kCg = true>
__device__ __forceinline__ float ld_in(const float *p)
{
    if (kCg) return __ldcg(p);
    return *p;
}

template <bool kCg = true>
__device__ __forceinline__ int expand_node(const TreeLayout &L, char *tb, const TreeHot &hot, int &tot_nodes, int &n_expanded, int &mt_pos,
                                           int &err, int slot, int hidx, float reward, float value,
                                           const float *__restrict__ probs, const float *__restrict__ beta,
                                           int K, float eps, const float *__restrict__ noises,
                                           const ExpandScratch &sc, int lane, bool draws_prefetched = false, int dbg_tree = -1,
                                           bool init_stats = true, int depth = 0)
{
    const int N = L.N, A = L.A, NA = N * A;

    // .cg loads (L2 only): under programmatic dependent launch this kernel can already be resident on an SM whose
    // L1 still holds the previous simulation's lines of these (re-used) buffers
    for (int t = lane; t < NA; t += 32) {
        sc.beta[t] = ld_in<kCg>(beta + t);
        sc.probs[t] = ld_in<kCg>(probs + t);
    }
    __syncwarp();
    MAZ_TS(L, dbg_tree, lane, 11);

    if (A >= 2) {
        // std::discrete_distribution::param_type::_M_initialize (random.tcc:2657-2678).  The two running sums are
        // sequential fp64 chains per agent (their order is observable), but the A divisions p_a = w_a / sum are
        // independent: one lane per (agent, action) does them in parallel (an fp64 divide is ~300 cycles of latency).
        if (N <= 32) {
            double *sums = reinterpret_cast<double *>(sc.keys);   // 32 x 8 bytes, free until the hash keys are built
            if (lane < N) {                                 // pass 1: sum_i = accumulate(w_i, 0.0), lane i
                const float *bi = sc.beta + lane * A;
                double sum = 0.0;
                for (int a = 0; a < A; ++a) sum = __dadd_rn(sum, (double)bi[a]);
                sums[lane] = sum;
            }
            __syncwarp();
            for (int t = lane; t < NA; t += 32)             // pass 2: every quotient, in parallel
                sc.cp[t] = __ddiv_rn((double)sc.beta[t], sums[t / A]);
            __syncwarp();
            if (lane < N) {                                 // pass 3: partial_sum, lane i; last entry forced to 1.0
                double *ci = sc.cp + lane * A;
                double run = ci[0];
                for (int a = 1; a < A; ++a) {
                    run = __dadd_rn(run, ci[a]);
                    ci[a] = run;
                }
                ci[A - 1] = 1.0;
            }
        } else {
            for (int i = lane; i < N; i += 32) {
                const float *bi = sc.beta + i * A;
                double sum = 0.0;
                for (int a = 0; a < A; ++a) sum = __dadd_rn(sum, (double)bi[a]);
                double run = 0.0;
                for (int a = 0; a < A; ++a) {
                    double p = __ddiv_rn((double)bi[a], sum);
                    run = (a == 0) ? p : __dadd_rn(run, p);
                    sc.cp[i * A + a] = run;
                }
                sc.cp[i * A + A - 1] = 1.0;
            }
        }
        __syncwarp();
        MAZ_TS(L, dbg_tree, lane, 12);
        // K*N draws in (k major, agent minor) order (cnode.cpp:251-259); 2 raw outputs per draw
        const int KN = K * N;
        uint32_t *mt = f_mt(L, tb);
        for (int j0 = 0; j0 < KN; j0 += kMtChunk / 2) {
            const int np = min(kMtChunk / 2, KN - j0);
            if (!draws_prefetched) mt_fetch(mt, mt_pos, sc.draws, 2 * np, lane);   // else: staged by the caller
            for (int t = lane; t < np; t += 32) {
                const int j = j0 + t;
                const int i = j % N;
                const double u = mt_canonical(sc.draws[2 * t], sc.draws[2 * t + 1]);
                const double *c = sc.cp + i * A;
                int lo = 0, hi = A;  // std::lower_bound: first c[a] with !(c[a] < u)
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (c[mid] < u) lo = mid + 1; else hi = mid;
                }
                sc.samp[j] = (uint8_t)lo;
            }
            __syncwarp();
        }
    } else {
        // A < 2: empty _M_cp, operator() returns 0 and consumes no randomness (random.tcc:2696-2697)
        for (int t = lane; t < K * N; t += 32) sc.samp[t] = 0;
        __syncwarp();
    }

    MAZ_TS(L, dbg_tree, lane, 13);
    // hash key of sample k = lane (cnode.cpp:258): key = key*23333 + a, C `long` wrap-around
    unsigned long long ukey = 0;
    if (lane < K) {
        const uint8_t *sk = sc.samp + lane * N;
        for (int i = 0; i < N; ++i) ukey = ukey * 23333ull + (unsigned long long)sk[i];
        sc.keys[lane] = (long long)ukey;
    }
    __syncwarp();
    const long long key = (long long)ukey;
    // std::map semantics: count per distinct key, action vector of the LAST sample with that key,
    // children in ascending key order.
    // lanes >= K get a private key so that they never match a real sample
    const unsigned same = __match_any_sync(MAZ_FULL, (lane < K) ? key : (long long)(0x7fffffffffff0000ll + lane));
    const int cnt = __popc(same);
    const bool is_last = (lane < K) && (31 - __clz(same) == lane);
    const unsigned lastmask = __ballot_sync(MAZ_FULL, is_last);
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < K; ++j) {
        long long kj = sc.keys[j];
        if (((lastmask >> j) & 1u) && kj < key) ++rank;
    }
    const int C = __popc(lastmask);
    const int base = tot_nodes;
    MAZ_TS(L, dbg_tree, lane, 14);
    if (base + C > L.P) {
        err = kErrPoolExhausted;
        return 0;
    }

    if (is_last) {
        const uint8_t *sk = sc.samp + lane * N;
        const int cs = base + rank;
        const float betahat_prob = __fdiv_rn((float)cnt, (float)K);  // count / sampled_times
        float beta_prob = 1.0f, pred_prob = 1.0f, prior = 1.0f;
        const float ome = __fsub_rn(1.0f, eps);
        uint8_t *act = f_actions(L, tb) + (size_t)cs * N;
        for (int i = 0; i < N; ++i) {
            const int a = sk[i];
            const float pb = sc.beta[i * A + a];
            const float pp = sc.probs[i * A + a];
            beta_prob = __fmul_rn(beta_prob, pb);
            pred_prob = __fmul_rn(pred_prob, pp);
            if (eps > 0) {
                float p = __fadd_rn(__fmul_rn(pp, ome), __fmul_rn(ld_in<kCg>(noises + i * A + a), eps));
                prior = __fmul_rn(prior, p);
            } else {
                prior = __fmul_rn(prior, pp);
            }
            act[i] = (uint8_t)a;
        }
        prior = __fdiv_rn(__fmul_rn(prior, betahat_prob), beta_prob);
        rec_store_new_child(L, tb, cs, prior);   // visit 0, no children, hidden index -1, reward(0.), pred_value(0.)
        f_pred_prob(L, tb)[cs] = pred_prob;       // (cnode.cpp:14-16); the root readouts return these for never-expanded children
        f_beta(L, tb)[cs] = beta_prob;
        f_beta_hat(L, tb)[cs] = betahat_prob;
    }
    if (lane == 0) {
        f_hidx(L, tb)[slot] = (int16_t)hidx;
        f_reward(L, tb)[slot] = reward;
        f_pred_value(L, tb)[slot] = value;
        f_nchild(L, tb)[slot] = (uint16_t)C;
        f_cbase(L, tb)[slot] = (uint16_t)base;
        if (init_stats) {   // (two-warp kernel: the backup warp owns the leaf's statistics)
            f_wsum(L, tb)[slot] = 0.0f;
            f_wtot(L, tb)[slot] = 0.0f;
        }
        hot.expslot[n_expanded] = (uint16_t)slot;
        hot.depth[n_expanded] = (uint16_t)depth;
        f_eid(L, tb)[slot] = (uint16_t)n_expanded;   // expansion order: index of this node's q-delta entry
    }
    tot_nodes = base + C;
    n_expanded += 1;
    __syncwarp();
    MAZ_TS(L, dbg_tree, lane, 15);
    return C;
}

// ---- SubTreeValueSet::update (utils.cpp:20-71) on the per-tree value log -----------------------------
// Scan result for one (node slot, depth) set: sizes and the two order statistics the reference queries.
struct VsScan {
    int cnt, nbig;
    uint32_t minbig, maxsmall;   // order-preserving images (f2ord) of min(big) / max(small), per lane
    int minpos, maxpos;          // log positions holding them (-1: none in this lane)
};
__device__ __forceinline__ void vs_scan_init(VsScan &r)
{
    r.cnt = 0; r.nbig = 0; r.minbig = 0xffffffffu; r.maxsmall = 0u; r.minpos = -1; r.maxpos = -1;
}
__device__ __forceinline__ void vs_scan_entry(VsScan &r, uint32_t tag, uint32_t k, float v, int e)
{
    if ((k & ~1u) == tag) {
        const uint32_t o = f2ord(v);
        ++r.cnt;
        if (k & 1u) {
            ++r.nbig;
            if (o < r.minbig || r.minpos < 0) { r.minbig = o; r.minpos = e; }
        } else {
            if (o > r.maxsmall || r.maxpos < 0) { r.maxsmall = o; r.maxpos = e; }
        }
    }
}
// The decision of SubTreeValueSet::update (utils.cpp:28-70) for ONE (node, depth) set, given its sizes and the two order
// statistics the reference queries.  Pure per-thread arithmetic in the reference's exact fp32 operation order; wsum / wtot are
// the node's weighted_sum / tot_weight and are updated in place.  lp = lam_pow[depth] (utils.cpp:25-26).
struct VsDecision {
    bool append_big;     // the new value joins the big set (else the small set)
    int flip_pos;        // log entry whose big/small flag flips (-1: none)
    uint32_t flip_key;   // its new key: an entry of this set is exactly tag | flag, so the flip is a plain store (a
                         // read-modify-write of the log in global memory would stall the warp for an L2 round trip per level)
    int err;
};
__device__ __forceinline__ VsDecision vs_decide(const TreeLayout &L, int cnt, int nbig, uint32_t minbig, int minpos, uint32_t maxsmall,
                                                int maxpos, uint32_t tag, float lp, float key, float &wsum, float &wtot)
{
    VsDecision d;
    d.flip_pos = -1; d.flip_key = 0; d.err = 0;
    const int nsmall = cnt - nbig;
    // size_lim = max(1, (int)ceil(count * (1 - quantile)))  (utils.cpp:31; float product, ceil)
    int lim = __float2int_ru(__fmul_rn((float)(cnt + 1), L.one_minus_rho));
    if (lim < 1) lim = 1;
    if (nbig == lim) {       // utils.cpp:33-48
        const float m = ord2f(minbig);
        if (key < m) {
            d.append_big = false;
        } else {
            d.flip_pos = minpos;
            d.flip_key = tag;                     // min(big) moves to the small set
            wsum = __fsub_rn(wsum, __fmul_rn(lp, m));
            wtot = __fsub_rn(wtot, lp);
            d.append_big = true;
            wtot = __fadd_rn(wtot, lp);
            wsum = __fadd_rn(wsum, __fmul_rn(lp, key));
        }
    } else {                 // utils.cpp:49-70
        if (nbig + 1 != lim) d.err = kErrValueSetInvariant;
        if (nsmall == 0) {
            d.append_big = true;
            wtot = __fadd_rn(wtot, lp);
            wsum = __fadd_rn(wsum, __fmul_rn(lp, key));
        } else {
            const float M = ord2f(maxsmall);
            if (key > M) {
                d.append_big = true;
                wtot = __fadd_rn(wtot, lp);
                wsum = __fadd_rn(wsum, __fmul_rn(lp, key));
            } else {
                d.flip_pos = maxpos;
                d.flip_key = tag | 1u;            // max(small) moves to the big set
                wtot = __fadd_rn(wtot, lp);
                wsum = __fadd_rn(wsum, __fmul_rn(lp, M));
                d.append_big = false;
            }
        }
    }
    return d;
}

// Warp-wide variant: the scan of the set is spread over the 32 lanes (root update in k_prepare).
__device__ __forceinline__ void vs_apply(const TreeLayout &L, char *tb, int &log_len, int &err, float &wsum, float &wtot,
                                         uint32_t tag, float lp, float key, int lane, const VsScan &r)
{
    uint32_t *vk = f_vskey(L, tb);
    float *vv = f_vsval(L, tb);
    const int cnt = __reduce_add_sync(MAZ_FULL, r.cnt);
    const int nbig = __reduce_add_sync(MAZ_FULL, r.nbig);
    const uint32_t gmin = __reduce_min_sync(MAZ_FULL, r.minpos >= 0 ? r.minbig : 0xffffffffu);
    const uint32_t gmax = __reduce_max_sync(MAZ_FULL, r.maxpos >= 0 ? r.maxsmall : 0u);
    const unsigned who_min = __ballot_sync(MAZ_FULL, r.minpos >= 0 && r.minbig == gmin);
    const unsigned who_max = __ballot_sync(MAZ_FULL, r.maxpos >= 0 && r.maxsmall == gmax);
    const int minpos = __shfl_sync(MAZ_FULL, r.minpos, who_min ? __ffs(who_min) - 1 : 0);
    const int maxpos = __shfl_sync(MAZ_FULL, r.maxpos, who_max ? __ffs(who_max) - 1 : 0);
    const VsDecision d = vs_decide(L, cnt, nbig, gmin, minpos, gmax, maxpos, tag, lp, key, wsum, wtot);
    if (d.err) err = d.err;
    if (log_len >= L.L) {
        err = kErrLogOverflow;
        return;
    }
    if (lane == 0) {
        if (d.flip_pos >= 0) vk[d.flip_pos] = d.flip_key;
        vk[log_len] = tag | (d.append_big ? 1u : 0u);
        vv[log_len] = key;
    }
    log_len += 1;
}
__device__ __forceinline__ uint32_t vs_tag(int slot, int depth) { return ((uint32_t)slot << 16) | ((uint32_t)depth << 1); }

// scan the log in global memory (entries [from, log_len)) and apply
__device__ __forceinline__ void vs_update(const TreeLayout &L, char *tb, int &log_len, int &err, float &wsum, float &wtot,
                                          int slot, int depth, float key, const float *__restrict__ lam_pow, int lane)
{
    const uint32_t *vk = f_vskey(L, tb);
    const float *vv = f_vsval(L, tb);
    const uint32_t tag = vs_tag(slot, depth);
    VsScan r;
    vs_scan_init(r);
    for (int e = lane; e < log_len; e += 32) vs_scan_entry(r, tag, vk[e], vv[e], e);
    vs_apply(L, tb, log_len, err, wsum, wtot, tag, lam_pow[depth], key, lane, r);
    __syncwarp();
}

}  // namespace maz
