// search_persist.cuh -- ONE kernel per search: a CTA owns a group of roots for all S simulations (sm_100a).
//
// Replaces the per-simulation loop of SampledMCTS.batch_search (core/mcts/tree_search/mcts_sampled.py:114-172):
//     selection -> gather parent hidden -> recurrent_inference (+ prediction) -> softmax / beta -> expansion + backup
// The two-kernel loop (fused inference + tree step, 2 launches per simulation inside a CUDA graph) ends every simulation with
// the slowest of ALL trees and pays two kernel boundaries; here the trees of a CTA only wait for each other:
//   * the CTA's 8 compute warps run the fused recurrent_inference of its tile on the tensor cores (the stages of
//     infer_hmma.cuh, unchanged arithmetic), then each of the first `rpt` warps steps ONE tree (expansion, backup and the
//     next selection: tree_step.cuh, the same bit-exact code as the stand-alone tree kernels), then the next simulation;
//   * no grid-wide barrier, no host, no launch between simulations: CTAs drift apart freely;
//   * the producer warp keeps streaming the weight chunks through the shared-memory ring across simulations, so the first
//     chunks of simulation s+1 arrive during the tree step of simulation s;
//   * the expansion scratch of the tree warps aliases the activation tiles (idle during the tree step).
// Trees are independent (cnode.cpp:571-576), so the grouping changes nothing in the results: every readout is bit-identical
// to the two-kernel loop (tests/test_search_native_gpu.py).
#pragma once
#include "infer_hmma.cuh"
#include "tree_step.cuh"

namespace maz {
namespace persist {

using namespace hmma;

struct SearchParams {
    Desc d;                    // network parameters + buffers.  d.pool: hidden-state pool (S+1, B, N*H); d.next_hidden: pool slot 1;
                               // d.idx_x / actions / reward / value / probs / beta: the per-simulation exchange buffers
    TreeLayout L;              // tree arena layout (L.N = agents in the tree = d.Nt)
    char *arena;
    const float *lam_pow, *logterm;
    const double *sqrtn;
    int table_len;
    float discount;
    int K, S;
    int *g_err;
    int *idx_y;                // (B,) the reference's second selection output (= arange)
    int *greedy_w;             // (S+1, B, N) greedy actions of every expanded node (sequential-agent mode), or NULL
    int rec;                   // 1: the exchange buffers are (S, ...) arrays, simulation s uses slice s (parity replay)
    long long *tree_clock;     // profiling (NULL): [2 s + {0,1}] = CTA 0's inference / tree-step cycles of simulation s; then per CTA c at
                               // [2 S + 4 c + ...]: total SM cycles, total nanoseconds (%globaltimer), sum of inference cycles, sum of tree cycles
};

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// tree-step scratch: the activation tiles X, T, HH, the one-hot / policy-hidden tiles and the q|k|v region are idle then
constexpr uint32_t TREE_SCRATCH_OFF = OFF_X;
constexpr uint32_t TREE_SCRATCH_BYTES = OFF_STAT - OFF_X;

__global__ void __launch_bounds__(NTHREADS, 1) k_search_persistent(const __grid_constant__ SearchParams P)
{
    HSM_DECL;
    __shared__ __align__(8) uint64_t bar_full[NSLOT], bar_empty[NSLOT], bar_vec;
    const Desc &d = P.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], NCONS / 32); }
        mbar_init(&bar_vec, 1);
        mbar_fence_init();
    }
    setup_rows(d);
    __syncthreads();

    if (warp == NCONS / 32) {
        if (lane == 0) produce_weights(d, bar_full, bar_empty, &bar_vec, P.S);
        return;
    }
    // ==================================== compute warps =========================================================
    Ring ring{bar_full, bar_empty, 0};
    const TreeLayout &L = P.L;
    const int rpt = tile_rpt(d);
    const int tree = blockIdx.x * rpt + warp;
    const bool has_tree = warp < rpt && tree < d.B;
    char *tb = P.arena + (size_t)(has_tree ? tree : 0) * L.slab_bytes;
    const size_t B = (size_t)d.B, NA = (size_t)L.N * L.A;
    const StepScratch scr = carve_step_scratch(reinterpret_cast<char *>(hsm) + TREE_SCRATCH_OFF + (size_t)warp * tree_scratch_bytes(L.N, L.A, L.K, L.S),
                                               L.N, L.A, L.K, L.S);
    long long *clk = (P.tree_clock != nullptr && blockIdx.x == 0 && tid == 0) ? P.tree_clock : nullptr;
    long long *cta_clk = (P.tree_clock != nullptr && tid == 0) ? P.tree_clock + 2 * P.S + 4 * blockIdx.x : nullptr;
    const long long k0 = cta_clk ? clock64() : 0;
    const unsigned long long g0 = cta_clk ? globaltimer_ns() : 0;
    long long sum_inf = 0, sum_tree = 0;
    // per tree (after the 64 stage timestamps): cycles in expansion + backup, cycles in selection, sum of path lengths, max
    long long *tree_clk = (P.tree_clock != nullptr && has_tree && lane == 0)
                              ? P.tree_clock + 2 * P.S + 4 * (long long)gridDim.x + 64 + 4 * (long long)tree : nullptr;
    long long tc_eb = 0, tc_sel = 0, tc_len = 0, tc_max = 0;

    mbar_wait(&bar_vec, 0);          // parameters resident
    setup_head_copies(d);
    cta_sync();

    // Iteration s: [tree warps] expansion + backup of simulation s-1 (its network outputs are in the exchange buffers), then the
    // selection of simulation s;  [all compute warps] gather + recurrent_inference of simulation s.  One call site per tree
    // function: a single inlined copy of each (with the kernel's parameters as constant-bank operands).
    SimIo io;
#pragma unroll 1
    for (int s = 0; s <= P.S; ++s) {
        const long long t1 = cta_clk ? clock64() : 0;
        if (has_tree) {
            const long long q0 = tree_clk ? clock64() : 0;
            if (s > 0)
                expand_backup_device(L, tb, f_hdr(tb), P.lam_pow, s, P.discount, P.K, io.reward + tree, io.value + tree,
                                     io.probs + (size_t)tree * NA, io.beta + (size_t)tree * NA, scr, lane, P.g_err);
            __syncwarp();
            const long long q1 = tree_clk ? clock64() : 0;
            if (s < P.S) {
                const size_t rn = P.rec ? (size_t)s : 0;
                select_next_device(L, tb, f_hdr(tb), P.logterm, P.sqrtn, P.table_len, P.discount, tree, lane, &scr,
                                   const_cast<int *>(d.idx_x) + rn * B, P.idx_y, const_cast<int *>(d.actions) + rn * B * L.N, P.g_err);
            }
            if (tree_clk) {
                const long long q2 = clock64();
                const int pl = f_hdr(tb)->path_len;
                tc_eb += q1 - q0; tc_sel += q2 - q1; tc_len += pl; tc_max = pl > tc_max ? pl : tc_max;
            }
        }
        cta_sync();                  // idx_x / actions of simulation s are visible; the tree scratch (activation tiles) is free
        const long long t2 = cta_clk ? clock64() : 0;
        if (cta_clk) {
            sum_tree += t2 - t1;
            if (clk && s > 0) clk[2 * (s - 1) + 1] = t2 - t1;
        }
        if (s == P.S) break;
        const size_t ro = P.rec ? (size_t)s : 0;
        io.idx_x = d.idx_x + ro * B;
        io.actions = d.actions + ro * B * L.N;
        io.next_hidden = d.next_hidden + (size_t)s * B * (size_t)(d.N * H);
        io.reward = d.reward + ro * B;
        io.value = d.value + ro * B;
        io.probs = d.probs + ro * B * NA;
        io.beta = d.beta + ro * B * NA;
        io.greedy = P.greedy_w ? P.greedy_w + (size_t)(s + 1) * B * d.N : nullptr;
        io.logits_out = nullptr;
        stage_gather(d, io);
        cta_sync();
        // profiling: stage timestamps of CTA 0 in the middle simulation, after the per-CTA counters (64 entries)
        long long *sclk = (clk && s == P.S / 2) ? P.tree_clock + 2 * P.S + 4 * (long long)gridDim.x : nullptr;
        if (sclk) { sclk[0] = t2; sclk[1] = clock64(); }
        infer_stages(d, io, ring, sclk);
        cta_sync();                  // the tile's network outputs are visible to the tree warps; the activation tiles are free
        if (cta_clk) {
            const long long t3 = clock64();
            sum_inf += t3 - t2;
            if (clk) clk[2 * s] = t3 - t2;
        }
    }
    if (tree_clk) { tree_clk[0] = tc_eb; tree_clk[1] = tc_sel; tree_clk[2] = tc_len; tree_clk[3] = tc_max; }
    if (cta_clk) {
        cta_clk[0] = clock64() - k0;
        cta_clk[1] = (long long)(globaltimer_ns() - g0);
        cta_clk[2] = sum_inf;
        cta_clk[3] = sum_tree;
    }
}

}  // namespace persist
}  // namespace maz
