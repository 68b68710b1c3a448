// infer_hmma.cuh -- the SMALL-BATCH variant of the fused MAMuZeroNet.recurrent_inference kernel (sm_100a).
//
// Same computation and same descriptor as infer_fused.cuh (config/smac/model.py:562-574 + core/config.py:430-499 +
// mcts_sampled.py:158-161); different decomposition.  tcgen05.mma needs a tile of 128 token rows per CTA (M = 64 only
// fills 16 lanes of each TMEM quadrant), so the 3m configuration of BASELINE.json (1024 roots x 3 agents = 3072 rows)
// occupies 26 of the 148 SMs and every one of them walks the 25-stage dependency chain of its tile at
// ~3.6 us per stage (TMEM round trip, MMA-warp hand-off, commit -> wait latency): 90 us per launch, tensor pipe 2 % busy.
// The search is bound by that LATENCY (one launch per simulation), not by tensor throughput.  This kernel trades
// peak tensor rate for parallelism and a short chain:
//   * 32-row tiles (floor(32/N) whole roots)  -> 103 CTAs for 3m instead of 26;
//   * warp-level mma.sync.m16n8k16 (bf16 x bf16 -> fp32): accumulators and the fp32 residual stream stay in REGISTERS,
//     the epilogue of a stage runs on the fragments it just computed -- no TMEM load, no issuer warp, no commit;
//   * a stage costs one or two 256-thread barriers; the three output heads share two stages (21 instead of 25);
//   * weights stream from L2 through the same 3-slot bulk-TMA ring (a dedicated producer warp), row-major with an
//     8-element pad so that ldmatrix is bank-conflict free.
// Above ~150 tiles of 32 rows the tcgen05 kernel wins (weights are re-read from L2 once per CTA: 0.9 MB x tiles);
// maz_infer_recurrent picks by batch size (maz_infer.cu).
#pragma once
#include <cuda_bf16.h>

#include "../../include/maz_infer.h"
#include "umma.cuh"

namespace maz {
namespace hmma {

using namespace umma;

constexpr int H = 128, NHEAD = 8, HD = 16, GH = 64, PH = 32, SUP = 11, NLAYER = 3;
constexpr int NCHUNK = MAZ_INFER_NCHUNK, NSLOT = 3;
constexpr int TM = 32;                       // token rows per CTA
constexpr int NCONS = 256, NTHREADS = NCONS + 32;   // 8 compute warps + 1 producer warp
constexpr int PAD = 8;                       // bf16 elements of row padding (16 bytes: ldmatrix rows land in distinct banks)
constexpr int LDA = (H + PAD) * 2;           // bytes per activation row
constexpr int LDQ = (3 * H + PAD) * 2;       // bytes per q|k|v row
constexpr int LDP = (PH + PAD) * 2;          // bytes per policy-hidden row
constexpr int LDG = GH + 4;                  // floats per graph-sum row
constexpr uint32_t SLOT_BYTES = 128 * (H + PAD) * 2;
#ifndef MAZ_HMMA_PIECE
#define MAZ_HMMA_PIECE 4352
#endif
constexpr uint32_t PIECE = MAZ_HMMA_PIECE;   // bytes per bulk-copy request (16 weight rows)

using Desc = ::maz_infer_desc;

struct RowInfo { int root, agent, valid, r0; };   // r0 = first row of my root inside the tile

struct Smem {
    uint8_t *X, *T, *One, *Ph, *Q, *W;
    float *G, *L, *Stat, *P;
    RowInfo *Row;
};

__host__ __device__ inline size_t smem_bytes(int KA, int vec_floats)
{
    return 2 * (size_t)TM * LDA + (size_t)TM * (KA + PAD) * 2 + (size_t)TM * LDP + (size_t)TM * LDQ + 3 * TM * 4 * 8 +
           TM * sizeof(RowInfo) + NSLOT * (size_t)SLOT_BYTES + (size_t)vec_floats * 4 + 256;
}

// ---- tensor-core primitives ------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    // (no "memory" clobber: every shared-memory write that a ldmatrix reads is separated from it by a bar.sync / mbarrier
    //  wait, which are compiler barriers; a clobber here would spill the accumulators of by-reference fragments)
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b)
{
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&p);
}

// who am I inside the 8 compute warps: warp = nq*2 + mt; the warp owns rows [16*mt, +16) and, of a 128-wide stage,
// columns [32*nq, +32) as four n8 tiles.  Fragment element acc[nt][2*hf + j] = (row 16*mt + g + 8*hf, col n0 + 8*nt + 2*t + j).
struct Thr {
    int lane, warp, g, t, mt, nq, rA, rB;
    uint32_t a_lane, w_lane;      // per-lane byte offsets of the ldmatrix row addresses (A: rows x k; W: out-rows x k)
    __device__ __forceinline__ Thr()
    {
        const int tid = threadIdx.x;
        lane = tid & 31; warp = tid >> 5; g = lane >> 2; t = lane & 3; mt = warp & 1; nq = warp >> 1;
        rA = 16 * mt + g; rB = rA + 8;
        a_lane = (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8);      // row inside the m16 tile; + (lane>>4)*16 bytes of k
        w_lane = (uint32_t)((lane & 7) + (lane >> 4) * 8);            // out-row inside a pair of n8 tiles; + ((lane>>3)&1)*16 bytes
    }
};

// acc[NT] (+)= A[16 rows x K] * W[n0 .. n0+8*NT)[K]^T.   a_tile: shared address of the activation tile row 0 (row stride lda
// bytes); w_slot: shared address of the weight chunk, rows = output features (row stride ldw bytes).  NT even.
template <int NT>
__device__ __forceinline__ void gemm(float (&acc)[NT][4], const Thr &th, uint32_t a_tile, int lda, uint32_t w_slot, int ldw, int n0, int K)
{
    const uint32_t a_addr = a_tile + (uint32_t)(16 * th.mt + th.a_lane) * lda + (uint32_t)(th.lane >> 4) * 16u;
    const uint32_t w_addr = w_slot + (uint32_t)(n0 + th.w_lane) * ldw + (uint32_t)((th.lane >> 3) & 1) * 16u;
#pragma unroll 4
    for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t a[4];
        ldsm4(a_addr + 2u * k0, a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int nt = 0; nt < NT; nt += 2) {
            uint32_t b0, b1, b2, b3;
            ldsm4(w_addr + (uint32_t)(8 * nt) * ldw + 2u * k0, b0, b1, b2, b3);
            mma16816(acc[nt], a, b0, b1);
            mma16816(acc[nt + 1], a, b2, b3);
        }
    }
}

// ---- weight ring (consumer side) ---------------------------------------------------------------------------------
struct Ring {
    uint64_t *full, *empty;
    uint32_t slots;
    int c;
    __device__ __forceinline__ uint32_t acquire()
    {
        const int s = c % NSLOT;
        mbar_wait(&full[s], (c / NSLOT) & 1);
        return slots + (uint32_t)s * SLOT_BYTES;
    }
    __device__ __forceinline__ void release(int lane)
    {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[c % NSLOT]);
        ++c;
    }
};

__device__ __forceinline__ void cta_sync() { named_bar_sync(1, NCONS); }

template <int NT>
__device__ __forceinline__ void zero(float (&acc)[NT][4])
{
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
}

// acc += bias[c] (fp32 parameters in shared memory), c = column of the fragment element
__device__ __forceinline__ void add_bias(float (&acc)[4][4], const Thr &th, const float *b, int n0)
{
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const float2 v = *reinterpret_cast<const float2 *>(b + n0 + 8 * nt + 2 * th.t);
        acc[nt][0] += v.x; acc[nt][1] += v.y; acc[nt][2] += v.x; acc[nt][3] += v.y;
    }
}
// fragment -> bf16 activation tile (row stride ld bytes), columns n0 + ...
__device__ __forceinline__ void store_tile(uint8_t *tile, int ld, const Thr &th, int n0, const float (&acc)[4][4])
{
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int c = n0 + 8 * nt + 2 * th.t;
        *reinterpret_cast<uint32_t *>(tile + (size_t)th.rA * ld + 2 * c) = pack2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<uint32_t *>(tile + (size_t)th.rB * ld + 2 * c) = pack2(acc[nt][2], acc[nt][3]);
    }
}

// row statistics of a fragment with NT tiles per warp, combined over the 4 lanes of a quad and the 4 column warps
// through `stat` ([TM][4] float2).  Contains ONE cta_sync; afterwards (mean, rstd) of rows rA / rB over `width` columns.
template <int NT>
__device__ __forceinline__ void row_stats_write(const float (&v)[NT][4], const Thr &th, float *stat)
{
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s0 += v[nt][0] + v[nt][1]; q0 += v[nt][0] * v[nt][0] + v[nt][1] * v[nt][1];
        s1 += v[nt][2] + v[nt][3]; q1 += v[nt][2] * v[nt][2] + v[nt][3] * v[nt][3];
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    }
    if (th.t == 0) {
        *reinterpret_cast<float2 *>(stat + (th.rA * 4 + th.nq) * 2) = make_float2(s0, q0);
        *reinterpret_cast<float2 *>(stat + (th.rB * 4 + th.nq) * 2) = make_float2(s1, q1);
    }
}
__device__ __forceinline__ void row_stats_read(const float *stat, int row, float inv_width, float &mean, float &rstd)
{
    const float4 a = *reinterpret_cast<const float4 *>(stat + row * 8), b = *reinterpret_cast<const float4 *>(stat + row * 8 + 4);
    const float s = a.x + a.z + b.x + b.z, q = a.y + a.w + b.y + b.w;
    mean = s * inv_width;
    rstd = rsqrtf(fmaxf(q * inv_width - mean * mean, 0.f) + 1e-5f);
}

// LayerNorm(128) with affine on a fragment (bias already added); optional ReLU after.  One cta_sync inside.
__device__ __forceinline__ void layer_norm128(float (&x)[4][4], const Thr &th, float *stat, const float *gam, const float *bet, int n0,
                                              bool relu)
{
    row_stats_write<4>(x, th, stat);
    cta_sync();
    float mA, rA_, mB, rB_;
    row_stats_read(stat, th.rA, 1.f / H, mA, rA_);
    row_stats_read(stat, th.rB, 1.f / H, mB, rB_);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int c = n0 + 8 * nt + 2 * th.t;
        const float2 gg = *reinterpret_cast<const float2 *>(gam + c), bb = *reinterpret_cast<const float2 *>(bet + c);
        x[nt][0] = (x[nt][0] - mA) * rA_ * gg.x + bb.x; x[nt][1] = (x[nt][1] - mA) * rA_ * gg.y + bb.y;
        x[nt][2] = (x[nt][2] - mB) * rB_ * gg.x + bb.x; x[nt][3] = (x[nt][3] - mB) * rB_ * gg.y + bb.y;
        if (relu) {
            x[nt][0] = fmaxf(x[nt][0], 0.f); x[nt][1] = fmaxf(x[nt][1], 0.f);
            x[nt][2] = fmaxf(x[nt][2], 0.f); x[nt][3] = fmaxf(x[nt][3], 0.f);
        }
    }
}

// ---- stages (not inlined: the kernel is one long straight line, small code keeps the instruction cache warm) ----------
// fp32 pool rows -> bf16 tile T; one-hot joint action -> tile One.  Thread (row = tid/8, 16 columns each).
__device__ __noinline__ void stage_gather(const Desc &d, const Smem &sm, bool onehot)
{
    const int tid = threadIdx.x, r = tid >> 3, part = tid & 7;
    const RowInfo ri = sm.Row[r];
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int action = -1;
    if (ri.valid) {
        const int ix = d.idx_x ? __ldcg(d.idx_x + ri.root) : 0;
        const float *h = d.pool + ((size_t)ix * d.B + ri.root) * (size_t)(d.N * H) + (size_t)ri.agent * H + part * 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(h) + i);
            w[2 * i] = pack2(v.x, v.y);
            w[2 * i + 1] = pack2(v.z, v.w);
        }
        action = __ldcg(d.actions + (size_t)ri.root * d.N + ri.agent);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(sm.T + (size_t)r * LDA + part * 32);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    if (onehot && part * 8 < d.KA) {     // 8 columns per thread
        const int c0 = part * 8;
        const __nv_bfloat16 one = __float2bfloat16(1.f), zer = __float2bfloat16(0.f);
        __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(sm.One + (size_t)r * (d.KA + PAD) * 2) + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (c0 + i == action) ? one : zer;
    }
}

// x0 = relu(W_in [h | onehot] + b) + pos[agent]  -> residual fragment x, bf16 tile X
__device__ __noinline__ void stage_inproj(const Desc &d, const Smem &sm, Ring &ring, float (&xio)[4][4])
{
    const Thr th;
    const int n0 = 32 * th.nq, ldo = (d.KA + PAD) * 2;
    float x[4][4];                  // registers (xio lives in the caller's frame: local memory)
    zero<4>(x);
    uint32_t w = ring.acquire();
    gemm<4>(x, th, smem_u32(sm.T), LDA, w, LDA, n0, H);
    ring.release(th.lane);
    w = ring.acquire();
    gemm<4>(x, th, smem_u32(sm.One), ldo, w, ldo, n0, d.KA);
    ring.release(th.lane);
    add_bias(x, th, sm.P + d.o_bin, n0);
    const float *pA = sm.P + d.o_pos + sm.Row[th.rA].agent * H, *pB = sm.P + d.o_pos + sm.Row[th.rB].agent * H;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int c = n0 + 8 * nt + 2 * th.t;
        const float2 a = *reinterpret_cast<const float2 *>(pA + c), b = *reinterpret_cast<const float2 *>(pB + c);
        x[nt][0] = fmaxf(x[nt][0], 0.f) + a.x; x[nt][1] = fmaxf(x[nt][1], 0.f) + a.y;
        x[nt][2] = fmaxf(x[nt][2], 0.f) + b.x; x[nt][3] = fmaxf(x[nt][3], 0.f) + b.y;
    }
    store_tile(sm.X, LDA, th, n0, x);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) xio[i][j] = x[i][j];
    cta_sync();
}

// q | k | v = X W^T + b  -> bf16 rows in Q (stride LDQ)
__device__ __noinline__ void stage_qkv(const Desc &d, const Smem &sm, Ring &ring, const float *lv)
{
    const Thr th;
    const int n0 = 32 * th.nq;
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
        float acc[4][4];
        zero<4>(acc);
        const uint32_t w = ring.acquire();
        gemm<4>(acc, th, smem_u32(sm.X), LDA, w, LDA, n0, H);
        ring.release(th.lane);
        add_bias(acc, th, lv + which * H, n0);
        store_tile(sm.Q + which * (H * 2), LDQ, th, n0, acc);
    }
    cta_sync();
}

// scaled dot-product attention over the agents of my root; thread = (row, head).  Output (bf16) -> tile T.
__device__ __noinline__ void stage_attention(const Desc &d, const Smem &sm)
{
    const int tid = threadIdx.x, r = tid >> 3, hh = tid & 7;
    const RowInfo ri = sm.Row[r];
    const int N = d.N;
    float q[16], o[16];
    {
        const uint4 *qp = reinterpret_cast<const uint4 *>(sm.Q + (size_t)r * LDQ + hh * 32);
        const uint4 a = qp[0], b = qp[1];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            q[2 * i] = __uint_as_float(w[i] << 16) * 0.25f;               // 1/sqrt(head_dim)
            q[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u) * 0.25f;
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = 0.f;
    float m = -INFINITY, lsum = 0.f;
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
        const uint8_t *kr = sm.Q + (size_t)(ri.r0 + j) * LDQ + H * 2 + hh * 32;
        const uint4 k0 = reinterpret_cast<const uint4 *>(kr)[0], k1 = reinterpret_cast<const uint4 *>(kr)[1];
        const uint4 v0 = reinterpret_cast<const uint4 *>(kr + H * 2)[0], v1 = reinterpret_cast<const uint4 *>(kr + H * 2)[1];
        const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
        const uint32_t vw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += q[2 * i] * __uint_as_float(kw[i] << 16) + q[2 * i + 1] * __uint_as_float(kw[i] & 0xffff0000u);
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn), p = __expf(s - mn);
        lsum = lsum * corr + p;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            o[2 * i] = o[2 * i] * corr + p * __uint_as_float(vw[i] << 16);
            o[2 * i + 1] = o[2 * i + 1] * corr + p * __uint_as_float(vw[i] & 0xffff0000u);
        }
        m = mn;
    }
    const float inv = 1.f / lsum;
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = pack2(o[2 * i] * inv, o[2 * i + 1] * inv);
    uint4 *dst = reinterpret_cast<uint4 *>(sm.T + (size_t)r * LDA + hh * 32);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    cta_sync();
}

// x = LN(x + A W^T + b) * g + be   (post-LN residual block: out-proj / linear2) -> fragment x, bf16 tile X
__device__ __noinline__ void stage_residual_ln(const Smem &sm, Ring &ring, float (&xio)[4][4], const uint8_t *a_tile, const float *bias,
                                               const float *gam, const float *bet)
{
    const Thr th;
    const int n0 = 32 * th.nq;
    float x[4][4];                  // registers (xio lives in the caller's frame: local memory)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) x[i][j] = xio[i][j];
    const uint32_t w = ring.acquire();
    gemm<4>(x, th, smem_u32(a_tile), LDA, w, LDA, n0, H);
    ring.release(th.lane);
    add_bias(x, th, bias, n0);
    layer_norm128(x, th, sm.Stat, gam, bet, n0, false);
    store_tile(sm.X, LDA, th, n0, x);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) xio[i][j] = x[i][j];
    cta_sync();
}

// f = relu(X W1^T + b1) -> bf16 tile T
__device__ __noinline__ void stage_linear_relu(const Smem &sm, Ring &ring, const float *bias)
{
    const Thr th;
    const int n0 = 32 * th.nq;
    float acc[4][4];
    zero<4>(acc);
    const uint32_t w = ring.acquire();
    gemm<4>(acc, th, smem_u32(sm.X), LDA, w, LDA, n0, H);
    ring.release(th.lane);
    add_bias(acc, th, bias, n0);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        acc[nt][0] = fmaxf(acc[nt][0], 0.f); acc[nt][1] = fmaxf(acc[nt][1], 0.f);
        acc[nt][2] = fmaxf(acc[nt][2], 0.f); acc[nt][3] = fmaxf(acc[nt][3], 0.f);
    }
    store_tile(sm.T, LDA, th, n0, acc);
    cta_sync();
}

// fc_dynamic (model.py:262-268): [h | onehot | attn] -> Linear LN ReLU -> Linear LN ReLU -> Linear, + h; next_hidden to
// global (fp32) and to tile T (bf16)
__device__ __noinline__ void stage_dynamics(const Desc &d, const Smem &sm, Ring &ring)
{
    const Thr th;
    const int n0 = 32 * th.nq, ldo = (d.KA + PAD) * 2;
    const float *dv = sm.P + d.o_dyn;
    float acc[4][4];
    zero<4>(acc);
    uint32_t w = ring.acquire();
    gemm<4>(acc, th, smem_u32(sm.T), LDA, w, LDA, n0, H);          // h
    ring.release(th.lane);
    w = ring.acquire();
    gemm<4>(acc, th, smem_u32(sm.One), ldo, w, ldo, n0, d.KA);     // one-hot action
    ring.release(th.lane);
    w = ring.acquire();
    gemm<4>(acc, th, smem_u32(sm.X), LDA, w, LDA, n0, H);          // attention output
    ring.release(th.lane);
    add_bias(acc, th, dv, n0);
    layer_norm128(acc, th, sm.Stat, dv + 128, dv + 256, n0, true);
    store_tile(sm.T, LDA, th, n0, acc);                            // (every warp is past its MMAs: the LN barrier)
    cta_sync();
    zero<4>(acc);
    w = ring.acquire();
    gemm<4>(acc, th, smem_u32(sm.T), LDA, w, LDA, n0, H);
    ring.release(th.lane);
    add_bias(acc, th, dv + 384, n0);
    layer_norm128(acc, th, sm.Stat, dv + 512, dv + 640, n0, true);
    store_tile(sm.X, LDA, th, n0, acc);
    cta_sync();
    zero<4>(acc);
    w = ring.acquire();
    gemm<4>(acc, th, smem_u32(sm.X), LDA, w, LDA, n0, H);
    ring.release(th.lane);
    add_bias(acc, th, dv + 768, n0);
    // + h (fp32 residual straight from the pool), next_hidden out
    const RowInfo ra = sm.Row[th.rA], rb = sm.Row[th.rB];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const RowInfo &ri = hf ? rb : ra;
        if (ri.valid) {
            const int ix = d.idx_x ? __ldcg(d.idx_x + ri.root) : 0;
            const float *h = d.pool + ((size_t)ix * d.B + ri.root) * (size_t)(d.N * H) + (size_t)ri.agent * H;
            float *nh = d.next_hidden + (size_t)ri.root * (d.N * H) + (size_t)ri.agent * H;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int c = n0 + 8 * nt + 2 * th.t;
                const float2 hv = __ldcg(reinterpret_cast<const float2 *>(h + c));
                acc[nt][2 * hf] += hv.x;
                acc[nt][2 * hf + 1] += hv.y;
                *reinterpret_cast<float2 *>(nh + c) = make_float2(acc[nt][2 * hf], acc[nt][2 * hf + 1]);
            }
        }
    }
    store_tile(sm.T, LDA, th, n0, acc);
    cta_sync();
}

// One GraphNetNN layer (model.py:151-163) on a fragment of the stacked GEMM.  The stacked weight rows are interleaved on the
// host so that warp nq holds, for the 16 features f = 16*nq .. +15, tiles 0-1 = gc.lin(x)[f] and tiles 2-3 = nn(x)[f].
// y = LN(relu(sum_over_agents(gc + b_gc) + nn + b_nn)) (no affine).  The agent sum goes through `G` ([TM][LDG] fp32).
// Contains two cta_syncs.  Result in y[2][4] (tile, element) for features 16*nq + 8*tile + 2t + j.
__device__ __forceinline__ void gnn_layer(const float (&acc)[4][4], const Thr &th, const Smem &sm, float *G, float *stat, const float *b,
                                          int N, float (&y)[2][4])
{
    const int f0 = 16 * th.nq + 2 * th.t;
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
        const float2 bg = *reinterpret_cast<const float2 *>(b + f0 + 8 * tl);
        *reinterpret_cast<float2 *>(G + th.rA * LDG + f0 + 8 * tl) = make_float2(acc[tl][0] + bg.x, acc[tl][1] + bg.y);
        *reinterpret_cast<float2 *>(G + th.rB * LDG + f0 + 8 * tl) = make_float2(acc[tl][2] + bg.x, acc[tl][3] + bg.y);
    }
    cta_sync();
    const int r0A = sm.Row[th.rA].r0, r0B = sm.Row[th.rB].r0;
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
        float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            const float2 a = *reinterpret_cast<const float2 *>(G + (r0A + j) * LDG + f0 + 8 * tl);
            const float2 c = *reinterpret_cast<const float2 *>(G + (r0B + j) * LDG + f0 + 8 * tl);
            sa.x += a.x; sa.y += a.y; sb.x += c.x; sb.y += c.y;
        }
        const float2 bn = *reinterpret_cast<const float2 *>(b + GH + f0 + 8 * tl);
        y[tl][0] = fmaxf(sa.x + acc[2 + tl][0] + bn.x, 0.f); y[tl][1] = fmaxf(sa.y + acc[2 + tl][1] + bn.y, 0.f);
        y[tl][2] = fmaxf(sb.x + acc[2 + tl][2] + bn.x, 0.f); y[tl][3] = fmaxf(sb.y + acc[2 + tl][3] + bn.y, 0.f);
    }
    row_stats_write<2>(y, th, stat);
    cta_sync();
    float mA, sA, mB, sB;
    row_stats_read(stat, th.rA, 1.f / GH, mA, sA);
    row_stats_read(stat, th.rB, 1.f / GH, mB, sB);
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
        y[tl][0] = (y[tl][0] - mA) * sA; y[tl][1] = (y[tl][1] - mA) * sA;
        y[tl][2] = (y[tl][2] - mB) * sB; y[tl][3] = (y[tl][3] - mB) * sB;
    }
}

// heads, first layers, ONE stage: reward GNN layer 1 on [next_hidden | onehot], value GNN layer 1 and fc_policy.0 on
// next_hidden.  Outputs: X[:, 0:64] = reward features, X[:, 64:128] = value features, Ph = relu(LN(policy hidden)).
__device__ __noinline__ void stage_heads1(const Desc &d, const Smem &sm, Ring &ring)
{
    const Thr th;
    const int n0 = 32 * th.nq, ldo = (d.KA + PAD) * 2;
    float aR[4][4], aV[4][4], aP[2][4];
    zero<4>(aR); zero<4>(aV); zero<2>(aP);
    uint32_t w = ring.acquire();
    gemm<4>(aR, th, smem_u32(sm.T), LDA, w, LDA, n0, H);
    ring.release(th.lane);
    w = ring.acquire();
    gemm<4>(aR, th, smem_u32(sm.One), ldo, w, ldo, n0, d.KA);
    ring.release(th.lane);
    w = ring.acquire();
    gemm<4>(aV, th, smem_u32(sm.T), LDA, w, LDA, n0, H);
    ring.release(th.lane);
    w = ring.acquire();
    if (th.nq < 2) gemm<2>(aP, th, smem_u32(sm.T), LDA, w, LDA, 16 * th.nq, H);   // 32 policy-hidden columns: warps nq = 0, 1
    ring.release(th.lane);
    float yR[2][4], yV[2][4];
    gnn_layer(aR, th, sm, sm.G, sm.Stat, sm.P + d.o_rg, d.N, yR);
    gnn_layer(aV, th, sm, sm.G + TM * LDG, sm.Stat + TM * 8, sm.P + d.o_vg, d.N, yV);
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
        const int c = 16 * th.nq + 8 * tl + 2 * th.t;
        *reinterpret_cast<uint32_t *>(sm.X + (size_t)th.rA * LDA + 2 * c) = pack2(yR[tl][0], yR[tl][1]);
        *reinterpret_cast<uint32_t *>(sm.X + (size_t)th.rB * LDA + 2 * c) = pack2(yR[tl][2], yR[tl][3]);
        *reinterpret_cast<uint32_t *>(sm.X + (size_t)th.rA * LDA + 2 * (GH + c)) = pack2(yV[tl][0], yV[tl][1]);
        *reinterpret_cast<uint32_t *>(sm.X + (size_t)th.rB * LDA + 2 * (GH + c)) = pack2(yV[tl][2], yV[tl][3]);
    }
    // policy hidden: relu(LN(acc + b) * g + be) over 32 columns; warps nq < 2 own columns [16*nq, +16)
    const float *pv = sm.P + d.o_pol;
    float *statP = sm.Stat + 2 * TM * 8;
    if (th.nq < 2) {
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
            const float2 bb = *reinterpret_cast<const float2 *>(pv + 16 * th.nq + 8 * tl + 2 * th.t);
            aP[tl][0] += bb.x; aP[tl][1] += bb.y; aP[tl][2] += bb.x; aP[tl][3] += bb.y;
        }
    } else {
        zero<2>(aP);
    }
    row_stats_write<2>(aP, th, statP);       // (warps nq >= 2 contribute zeros)
    cta_sync();
    if (th.nq < 2) {
        float mA, sA, mB, sB;
        row_stats_read(statP, th.rA, 1.f / PH, mA, sA);
        row_stats_read(statP, th.rB, 1.f / PH, mB, sB);
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
            const int c = 16 * th.nq + 8 * tl + 2 * th.t;
            const float2 gg = *reinterpret_cast<const float2 *>(pv + PH + c), bb = *reinterpret_cast<const float2 *>(pv + 2 * PH + c);
            const float v0 = fmaxf((aP[tl][0] - mA) * sA * gg.x + bb.x, 0.f), v1 = fmaxf((aP[tl][1] - mA) * sA * gg.y + bb.y, 0.f);
            const float v2 = fmaxf((aP[tl][2] - mB) * sB * gg.x + bb.x, 0.f), v3 = fmaxf((aP[tl][3] - mB) * sB * gg.y + bb.y, 0.f);
            *reinterpret_cast<uint32_t *>(sm.Ph + (size_t)th.rA * LDP + 2 * c) = pack2(v0, v1);
            *reinterpret_cast<uint32_t *>(sm.Ph + (size_t)th.rB * LDP + 2 * c) = pack2(v2, v3);
        }
    }
    cta_sync();
}

// softmax . support -> inv_h (core/config.py:430-442, 463-499)
__device__ __forceinline__ float support_to_scalar(const float *lg)
{
    float m = lg[0];
#pragma unroll
    for (int k = 1; k < SUP; ++k) m = fmaxf(m, lg[k]);
    float s = 0.f, x = 0.f;
#pragma unroll
    for (int k = 0; k < SUP; ++k) {
        const float e = __expf(lg[k] - m);
        s += e;
        x += e * (float)(k - 5);
    }
    x /= s;
    const float eps = 0.001f;
    const float r = (sqrtf(1.f + 4.f * eps * (fabsf(x) + 1.f + eps)) - 1.f) / (2.f * eps);
    float out = r * r - 1.f;
    out = (x < 0.f) ? -out : out;
    if (!(out == out)) out = 0.f;
    if (fabsf(out) < eps) out = 0.f;
    return out;
}

// partial logits of the 64 -> 11 value/reward head over my 4 features per row, summed over the quad; lane t == 0 writes
// L[kind][row][nq][k]
__device__ __forceinline__ void head_partial(const float (&y)[2][4], const Thr &th, const float *V, float *L)
{
    const int f0 = 16 * th.nq + 2 * th.t;
#pragma unroll 1
    for (int k = 0; k < SUP; ++k) {
        const float2 w0 = *reinterpret_cast<const float2 *>(V + k * GH + f0), w1 = *reinterpret_cast<const float2 *>(V + k * GH + f0 + 8);
        float a = y[0][0] * w0.x + y[0][1] * w0.y + y[1][0] * w1.x + y[1][1] * w1.y;
        float b = y[0][2] * w0.x + y[0][3] * w0.y + y[1][2] * w1.x + y[1][3] * w1.y;
        a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
        b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2);
        if (th.t == 0) {
            L[(th.rA * 4 + th.nq) * 12 + k] = a;
            L[(th.rB * 4 + th.nq) * 12 + k] = b;
        }
    }
}

// heads, second layers + outputs, ONE stage: reward / value GNN layer 2 -> mean over agents -> 64->11 head -> scalar;
// fc_policy.3 -> softmax / beta / greedy (mcts_sampled.py:158-161).
__device__ __noinline__ void stage_heads2(const Desc &d, const Smem &sm, Ring &ring)
{
    const Thr th;
    const int n0 = 32 * th.nq, ldg = (GH + PAD) * 2, ldp = (PH + PAD) * 2;
    const int N = d.N, A = d.A, NAP = d.NAP;
    float aR[4][4], aV[4][4], aP[2][4];
    zero<4>(aR); zero<4>(aV); zero<2>(aP);
    uint32_t w = ring.acquire();
    gemm<4>(aR, th, smem_u32(sm.X), LDA, w, ldg, n0, GH);
    ring.release(th.lane);
    w = ring.acquire();
    gemm<4>(aV, th, smem_u32(sm.X) + GH * 2, LDA, w, ldg, n0, GH);
    ring.release(th.lane);
    w = ring.acquire();
    const int pc0 = 16 * th.nq;                                    // policy logit columns [16*nq, +16) when < NAP
    if (pc0 < NAP) gemm<2>(aP, th, smem_u32(sm.Ph), LDP, w, ldp, pc0, PH);
    ring.release(th.lane);
    float yR[2][4], yV[2][4];
    gnn_layer(aR, th, sm, sm.G, sm.Stat, sm.P + d.o_rg + 128, N, yR);
    gnn_layer(aV, th, sm, sm.G + TM * LDG, sm.Stat + TM * 8, sm.P + d.o_vg + 128, N, yV);
    cta_sync();                                                    // G is reused below as L / logits scratch
    float *L = sm.G;                                               // [2][TM][4][12]
    head_partial(yR, th, sm.P + d.o_rg + 256, L);
    head_partial(yV, th, sm.P + d.o_vg + 256, L + TM * 48);
    // policy logits (+ bias) -> fp32 scratch [TM][NAP] behind L
    float *PL = sm.G + 2 * TM * 48;
    if (pc0 < NAP) {
        const float *b2 = sm.P + d.o_pol + 96;
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
            const int c = pc0 + 8 * tl + 2 * th.t;
            const float2 bb = *reinterpret_cast<const float2 *>(b2 + c);
            *reinterpret_cast<float2 *>(PL + th.rA * NAP + c) = make_float2(aP[tl][0] + bb.x, aP[tl][1] + bb.y);
            *reinterpret_cast<float2 *>(PL + th.rB * NAP + c) = make_float2(aP[tl][2] + bb.x, aP[tl][3] + bb.y);
        }
    }
    cta_sync();
    const int tid = threadIdx.x;
    const int rpt = TM / N;
    // (root, kind, k): logit k = bias + mean over the root's rows of the sum over the 4 column warps
    float *LG = PL + TM * NAP;                                     // [rpt][2][12]
    for (int i = tid; i < rpt * 2 * SUP; i += NCONS) {
        const int rl = i / (2 * SUP), rem = i - rl * (2 * SUP), kind = rem / SUP, k = rem - kind * SUP;
        const float *src = L + kind * (TM * 48);
        float s = 0.f;
        for (int j = 0; j < N; ++j) {
            const float *p = src + ((rl * N + j) * 4) * 12 + k;
            s += p[0] + p[12] + p[24] + p[36];
        }
        const float *vb = sm.P + (kind ? d.o_vg : d.o_rg) + 256 + SUP * GH;
        LG[(rl * 2 + kind) * 12 + k] = s / (float)N + vb[k];
    }
    // policy outputs: 8 threads per row
    {
        const int r = tid >> 3, sub = tid & 7;
        const RowInfo ri = sm.Row[r];
        const float *lg = PL + r * NAP;
        float m = -INFINITY;
        int am = 0x7fffffff;
        for (int a = sub; a < A; a += 8) {
            const float v = lg[a];
            if (v > m) { m = v; am = a; }
        }
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, m, o);
            const int oa = __shfl_xor_sync(0xffffffffu, am, o);
            if (om > m || (om == m && oa < am)) { m = om; am = oa; }
        }
        const bool unit_tau = (d.inv_tau == 1.0f);
        float s = 0.f, sb = 0.f;
        for (int a = sub; a < A; a += 8) {
            const float e = __expf(lg[a] - m);
            s += e;
            sb += unit_tau ? e : __powf(e, d.inv_tau);
        }
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        if (ri.valid) {
            if (d.greedy && sub == 0) d.greedy[(size_t)ri.root * N + ri.agent] = am;
            const int ta = (d.cur < 0) ? ri.agent : (ri.agent == d.cur ? 0 : -1);
            const float invs = 1.f / s, invb = 1.f / sb;
            for (int a = sub; a < A; a += 8) {
                const float v = lg[a];
                if (d.logits_out) d.logits_out[((size_t)ri.root * N + ri.agent) * A + a] = v;
                if (ta >= 0) {
                    const float e = __expf(v - m);
                    const size_t o = ((size_t)ri.root * d.Nt + ta) * A + a;
                    d.probs[o] = e * invs;
                    d.beta[o] = (unit_tau ? e : __powf(e, d.inv_tau)) * invb;
                }
            }
        }
    }
    cta_sync();
    if (tid < rpt * 2) {
        const int rl = tid >> 1, kind = tid & 1;
        const int root = blockIdx.x * rpt + rl;
        if (root < d.B) {
            const float v = support_to_scalar(LG + (rl * 2 + kind) * 12);
            (kind ? d.value : d.reward)[root] = v;
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) k_recurrent_inference_small(const __grid_constant__ Desc d)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[NSLOT], bar_empty[NSLOT], bar_vec;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int N = d.N, KA = d.KA;
    Smem sm;
    uint8_t *p = smem;
    sm.X = p; p += TM * LDA;
    sm.T = p; p += TM * LDA;
    sm.One = p; p += TM * (KA + PAD) * 2;
    sm.Ph = p; p += TM * LDP;
    sm.Q = p; sm.G = reinterpret_cast<float *>(p); p += TM * LDQ;     // q|k|v rows (attention) / graph sums + logits (heads)
    sm.Stat = reinterpret_cast<float *>(p); p += 3 * TM * 4 * 8;
    sm.Row = reinterpret_cast<RowInfo *>(p); p += TM * sizeof(RowInfo);
    p = smem + ((p - smem + 127) / 128) * 128;
    sm.W = p; p += NSLOT * (size_t)SLOT_BYTES;
    sm.P = reinterpret_cast<float *>(p);
    sm.L = nullptr;

    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], NCONS / 32); }
        mbar_init(&bar_vec, 1);
        mbar_fence_init();
    }
    if (tid < TM) {
        const int rpt = TM / N, rl = tid / N, agent = tid - rl * N;
        const int root = blockIdx.x * rpt + rl;
        RowInfo ri;
        ri.valid = (rl < rpt) && (root < d.B);
        ri.root = root; ri.agent = (rl < rpt) ? agent : 0;
        ri.r0 = (rl < rpt) ? rl * N : 0;
        sm.Row[tid] = ri;
    }
    __syncthreads();

    if (warp == NCONS / 32) {
        // ================================ weight producer (one thread) =========================================
        if ((tid & 31) == 0) {
            mbar_expect_tx(&bar_vec, (uint32_t)d.vec_floats * 4u);
            bulk_g2s(sm.P, d.vec, (uint32_t)d.vec_floats * 4u, &bar_vec);
#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c) {
                const int s = c % NSLOT;
                if (c >= NSLOT) mbar_wait_backoff(&bar_empty[s], ((c / NSLOT) - 1) & 1);
                // one chunk = several concurrent bulk copies on the same mbarrier: a single 1-D bulk request streams at only
                // a few bytes per cycle (measured: 34 KB in ~4.6 us), independent requests overlap
                const uint32_t nbytes = (d.dbg_flags & 4) ? 16u : d.chunk_bytes[c];   // profiling: tiny copies
                const uint8_t *src = reinterpret_cast<const uint8_t *>(d.wpk) + d.chunk_off[c];
                uint8_t *dst = sm.W + (size_t)s * SLOT_BYTES;
                mbar_expect_tx(&bar_full[s], nbytes);
                for (uint32_t o = 0; o < nbytes; o += PIECE) {
                    const uint32_t n = (nbytes - o < PIECE) ? (nbytes - o) : PIECE;
                    bulk_g2s(dst + o, src + o, n, &bar_full[s]);
                }
            }
        }
        return;
    }
    // ==================================== compute warps =========================================================
    Ring ring{bar_full, bar_empty, smem_u32(sm.W), 0};
    int ts_n = 0;
    const bool ts_on = d.dbg_clock != nullptr && blockIdx.x == 0 && tid == 0;
#define TS() \
    if (ts_on && ts_n < 64) d.dbg_clock[ts_n++] = clock64();
    TS();
    stage_gather(d, sm, true);
    TS();
    mbar_wait(&bar_vec, 0);          // parameters resident
    cta_sync();
    TS();
    float x[4][4];
    stage_inproj(d, sm, ring, x);
    TS();
#pragma unroll 1
    for (int l = 0; l < NLAYER; ++l) {
        const float *lv = sm.P + d.o_layer + l * 1280;     // bq bk bv bo g1 be1 b1 b2 g2 be2
        stage_qkv(d, sm, ring, lv);
        TS();
        stage_attention(d, sm);
        TS();
        stage_residual_ln(sm, ring, x, sm.T, lv + 384, lv + 512, lv + 640);      // x = norm1(x + sa Wo^T + bo)
        TS();
        stage_linear_relu(sm, ring, lv + 768);                                    // f = relu(linear1(x))
        TS();
        stage_residual_ln(sm, ring, x, sm.T, lv + 896, lv + 1024, lv + 1152);    // x = norm2(x + linear2(f))
        TS();
    }
    stage_gather(d, sm, false);      // h (bf16) back into T; X holds the attention output
    cta_sync();
    TS();
    stage_dynamics(d, sm, ring);
    TS();
    stage_heads1(d, sm, ring);
    TS();
    stage_heads2(d, sm, ring);
    TS();
#undef TS
}

}  // namespace hmma
}  // namespace maz
