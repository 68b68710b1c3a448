// infer_fused.cuh -- ONE kernel for MAMuZeroNet.recurrent_inference inside the search (sm_100a).
//
// Reference computation: config/smac/model.py:562-574 (dynamics :251-282 with attention.py:27-43,
// prediction :335-373, GraphNetNN :136-174) + inverse support transform core/config.py:430-442,463-499
// + the driver's softmax / beta post-processing mcts_sampled.py:158-161.
//
// Tiling: a CTA of 128 threads owns a tile of 128 token rows (thread r == row r == TMEM lane r).  A token
// is one agent of one root; every warp holds floor(32/N) whole roots, so everything that mixes the
// agents of a root (attention over the agent axis, the all-ones-adjacency graph sums) stays inside a warp.
// All 30 weight matrices stream from L2 through a 3-slot shared-memory ring (1-D bulk TMA + mbarriers),
// pre-packed on the host in the tcgen05 operand layout; every GEMM is M=128 on the tensor cores with the
// fp32 accumulator in TMEM; the residual stream x lives in TMEM (columns 0..127) so `x + f(x)` is just an
// accumulating MMA.  Per-row work (bias, ReLU, LayerNorm, softmax, support transform) is thread-local.
#pragma once
#include "../../include/maz_infer.h"
#include "umma.cuh"

namespace maz {
namespace fused {

using namespace umma;

constexpr int H = 128;        // hidden per agent            (config/smac/__init__.py:15)
constexpr int NHEAD = 8;      // attention.py:36
constexpr int HD = 16;
constexpr int GH = 64;        // GNN hidden                  (model.py:213,300 defaults)
constexpr int PH = 32;        // fc_policy_layers = [32]
constexpr int SUP = 11;       // DiscreteSupport(-5, 5)
constexpr int NLAYER = 3;     // AttentionEncoder(3, ...)    (model.py:221)
constexpr int NCHUNK = 30;
constexpr int NSLOT = 3;

enum : uint32_t { TM_X = 0, TM_Q = 128, TM_K = 256, TM_V = 384, TM_ACC = 128 };

using Desc = ::maz_infer_desc;
static_assert(NCHUNK == MAZ_INFER_NCHUNK, "chunk count");

__host__ __device__ inline uint32_t slot_bytes(int KA) { return 128u * (128u + (uint32_t)KA) * 2u; }
__host__ __device__ inline size_t smem_bytes(int KA)
{
    return 2 * operand_bytes(128, 128) + operand_bytes(128, KA) + NSLOT * (size_t)slot_bytes(KA) + 1024;
}

// ---- weight-ring producer / MMA issuer state (thread 0 only) ---------------------------------------------
struct Issuer {
    const Desc *d;
    uint8_t *slots;
    uint32_t slot_sz;
    uint64_t *full, *empty;
    int next_load;

    __device__ __forceinline__ void load(int c)
    {
        const int s = c % NSLOT;
        if (c >= NSLOT) mbar_wait(&empty[s], ((c / NSLOT) - 1) & 1);
        mbar_expect_tx(&full[s], d->chunk_bytes[c]);
        bulk_g2s(slots + (size_t)s * slot_sz, reinterpret_cast<const uint8_t *>(d->wpk) + d->chunk_off[c], d->chunk_bytes[c], &full[s]);
    }
    __device__ __forceinline__ void ensure(int upto)
    {
        if (upto > NCHUNK - 1) upto = NCHUNK - 1;
        while (next_load <= upto) load(next_load++);
    }
    // wait until chunk c is resident; returns its shared-memory byte address
    __device__ __forceinline__ uint32_t acquire(int c)
    {
        ensure(c + 2);
        mbar_wait(&full[c % NSLOT], (c / NSLOT) & 1);
        tc_fence_after();
        return smem_u32(slots + (size_t)(c % NSLOT) * slot_sz);
    }
    // the slot may be refilled once every MMA issued so far has completed
    __device__ __forceinline__ void release(int c) { mma_commit(&empty[c % NSLOT]); }
};

// ---- small per-thread helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack2(float a, float b)
{
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&p);
}
// write 16 consecutive columns [c, c+16) of row `row` into a bf16 operand tile with row length K
__device__ __forceinline__ void store16(uint8_t *tile, int row, int c, int K, const float (&v)[16])
{
    uint4 a, b;
    a.x = pack2(v[0], v[1]); a.y = pack2(v[2], v[3]); a.z = pack2(v[4], v[5]); a.w = pack2(v[6], v[7]);
    b.x = pack2(v[8], v[9]); b.y = pack2(v[10], v[11]); b.z = pack2(v[12], v[13]); b.w = pack2(v[14], v[15]);
    *reinterpret_cast<uint4 *>(tile + chunk_off(row, c >> 3, K)) = a;
    *reinterpret_cast<uint4 *>(tile + chunk_off(row, (c >> 3) + 1, K)) = b;
}
__device__ __forceinline__ void add_vec16(float (&v)[16], const float *__restrict__ p)
{
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p + i));
        v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
    }
}

// inverse categorical transform of an 11-way support head (core/config.py:430-442, 463-499)
__device__ __forceinline__ float support_to_scalar(const float (&lg)[SUP])
{
    float m = lg[0];
#pragma unroll
    for (int t = 1; t < SUP; ++t) m = fmaxf(m, lg[t]);
    float s = 0.f, x = 0.f;
#pragma unroll
    for (int t = 0; t < SUP; ++t) {
        const float e = __expf(lg[t] - m);
        s += e;
        x += e * (float)(t - 5);
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

// y = LN(pre) over 128 TMEM columns starting at `tcol`, pre = acc + bias; writes fp32 to TMEM X (optional)
// and bf16 to `tile` (K=128).  relu_after: apply ReLU after the affine (mlp()), else plain (post-LN encoder).
__device__ __forceinline__ void ln128(uint32_t trow, uint32_t tcol, const float *__restrict__ bias, const float *__restrict__ g,
                                      const float *__restrict__ be, bool relu_after, bool store_x, uint8_t *tile, int row)
{
    float sum = 0.f, sq = 0.f;
#pragma unroll 1
    for (int c = 0; c < 128; c += 16) {
        float v[16];
        tmem_ld16(trow + tcol + c, v);
        add_vec16(v, bias + c);
#pragma unroll
        for (int i = 0; i < 16; ++i) { sum += v[i]; sq += v[i] * v[i]; }
    }
    const float mean = sum * (1.f / 128.f);
    const float rstd = rsqrtf(fmaxf(sq * (1.f / 128.f) - mean * mean, 0.f) + 1e-5f);
#pragma unroll 1
    for (int c = 0; c < 128; c += 16) {
        float v[16];
        tmem_ld16(trow + tcol + c, v);
        add_vec16(v, bias + c);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float y = (v[i] - mean) * rstd * __ldg(g + c + i) + __ldg(be + c + i);
            v[i] = relu_after ? fmaxf(y, 0.f) : y;
        }
        if (store_x) tmem_st16(trow + TM_X + c, v);
        store16(tile, row, c, 128, v);
    }
    if (store_x) tmem_st_wait();
}

// GraphNetNN layer (model.py:151-163) from one stacked GEMM: columns [0,64) = gc.lin(x), [64,128) = nn(x).
// out[i] = LN(relu(sum_over_agents(gc + b_gc) + nn + b_nn)), no affine.  Result in y[64] (registers).
__device__ __forceinline__ void gnn_layer(uint32_t trow, const float *__restrict__ bgc, const float *__restrict__ bnn, int N,
                                          int root_lane0, float (&y)[GH])
{
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int c = 0; c < GH; c += 16) {
        float gsum[16], g[16], nn[16];
        tmem_ld16(trow + TM_ACC + c, g);
        add_vec16(g, bgc + c);
#pragma unroll
        for (int i = 0; i < 16; ++i) gsum[i] = 0.f;
        for (int j = 0; j < N; ++j) {
#pragma unroll
            for (int i = 0; i < 16; ++i) gsum[i] += __shfl_sync(0xffffffffu, g[i], root_lane0 + j);
        }
        tmem_ld16(trow + TM_ACC + GH + c, nn);
        add_vec16(nn, bnn + c);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float t = fmaxf(gsum[i] + nn[i], 0.f);
            y[c + i] = t;
            sum += t;
            sq += t * t;
        }
    }
    const float mean = sum * (1.f / GH);
    const float rstd = rsqrtf(fmaxf(sq * (1.f / GH) - mean * mean, 0.f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < GH; ++i) y[i] = (y[i] - mean) * rstd;
}

// mean-pool over the agents of the root, then the 64 -> 11 head and the support transform
__device__ __forceinline__ float gnn_head(float (&o)[GH], int N, int root_lane0, const float *__restrict__ Vw,
                                          const float *__restrict__ Vb)
{
    const float invn = 1.f / (float)N;
#pragma unroll
    for (int i = 0; i < GH; ++i) {
        float s = 0.f;
        const float mine = o[i];
        for (int j = 0; j < N; ++j) s += __shfl_sync(0xffffffffu, mine, root_lane0 + j);
        o[i] = s * invn;
    }
    float lg[SUP];
#pragma unroll
    for (int t = 0; t < SUP; ++t) {
        float a = __ldg(Vb + t);
#pragma unroll
        for (int i = 0; i < GH; ++i) a += __ldg(Vw + t * GH + i) * o[i];
        lg[t] = a;
    }
    return support_to_scalar(lg);
}

__global__ void __launch_bounds__(128, 1) k_recurrent_inference(const __grid_constant__ Desc d)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[NSLOT], bar_empty[NSLOT], bar_mma;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = d.N, A = d.A, KA = d.KA;
    uint8_t *sX = smem;                               // 128 x 128 bf16
    uint8_t *sT = sX + operand_bytes(128, 128);       // 128 x 128 bf16
    uint8_t *sOne = sT + operand_bytes(128, 128);     // 128 x KA  bf16 (one-hot joint action)
    uint8_t *sW = sOne + operand_bytes(128, KA);      // weight ring
    const uint32_t slot_sz = slot_bytes(KA);

    // token <-> (root, agent)
    const int rpw = 32 / N;                           // whole roots per warp
    const int rl = lane / N, agent = lane - rl * N;
    const int root = (blockIdx.x * 4 + warp) * rpw + rl;
    const bool valid = (rl < rpw) && (root < d.B);
    const int root_lane0 = (rl < rpw) ? rl * N : 0;   // first lane of my root inside the warp
    const int row = tid;

    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        mbar_init(&bar_mma, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes

    Issuer is{&d, sW, slot_sz, bar_full, bar_empty, 0};
    if (tid == 0) is.ensure(NSLOT - 1);               // start streaming the first weight chunks
    uint32_t mma_phase = 0;
    const uint32_t aX = smem_u32(sX), aT = smem_u32(sT), aOne = smem_u32(sOne);

    // ---- gather the parent's hidden state and build the one-hot action operand ----------------------------
    const float *hrow = nullptr;
    int my_action = 0;
    if (valid) {
        const int ix = d.idx_x ? d.idx_x[root] : 0;
        hrow = d.pool + ((size_t)ix * d.B + root) * (size_t)(N * H) + (size_t)agent * H;
        my_action = d.actions[(size_t)root * N + agent];
    }
    auto gather_h = [&]() {
#pragma unroll 4
        for (int c = 0; c < H; c += 16) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                float4 t = valid ? *reinterpret_cast<const float4 *>(hrow + c + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
            }
            store16(sT, row, c, H, v);
        }
    };
    gather_h();
    for (int c = 0; c < KA; c += 16) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (valid && (c + i) == my_action) ? 1.f : 0.f;
        store16(sOne, row, c, KA, v);
    }

#define STAGE_SYNC()          \
    fence_proxy_async();      \
    tc_fence_before();        \
    __syncthreads();
#define WAIT_MMA()                 \
    mbar_wait(&bar_mma, mma_phase); \
    mma_phase ^= 1;                 \
    tc_fence_after();

    int c = 0;  // next weight chunk (all threads track it; only thread 0 uses it)
    const float *vec = d.vec;

    // ---- attention_stack[0..1]: x0 = relu(W_in [h | onehot] + b) (+ positional table) ----------------------
    STAGE_SYNC();
    if (tid == 0) {
        const uint32_t w = is.acquire(c);
        issue_gemm(tmem + TM_ACC, aT, H, 0, w, H + KA, 0, H, 128, false);
        issue_gemm(tmem + TM_ACC, aOne, KA, 0, w, H + KA, H, KA, 128, true);
        is.release(c);
        mma_commit(&bar_mma);
    }
    ++c;
    WAIT_MMA();
    {
        const float *pos = vec + d.o_pos + agent * H;
#pragma unroll 1
        for (int cc = 0; cc < H; cc += 16) {
            float v[16];
            tmem_ld16(trow + TM_ACC + cc, v);
            add_vec16(v, vec + d.o_bin + cc);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f) + __ldg(pos + cc + i);
            tmem_st16(trow + TM_X + cc, v);
            store16(sX, row, cc, H, v);
        }
        tmem_st_wait();
    }

    // ---- 3 x post-LN TransformerEncoderLayer over the agent axis (attention.py:36-43) ------------------------
#pragma unroll 1
    for (int l = 0; l < NLAYER; ++l) {
        const float *lv = vec + d.o_layer + l * 1280;
        const float *bq = lv, *bk = lv + 128, *bv = lv + 256, *bo = lv + 384, *g1 = lv + 512, *be1 = lv + 640;
        const float *b1 = lv + 768, *b2 = lv + 896, *g2 = lv + 1024, *be2 = lv + 1152;
        STAGE_SYNC();
        if (tid == 0) {
            for (int p = 0; p < 3; ++p) {  // q, k, v projections into three TMEM regions
                const uint32_t w = is.acquire(c + p);
                issue_gemm(tmem + TM_Q + 128 * p, aX, H, 0, w, H, 0, H, 128, false);
                is.release(c + p);
            }
            mma_commit(&bar_mma);
        }
        c += 3;
        WAIT_MMA();
        // scaled dot-product attention, one head at a time, online softmax over the N agents of my root
#pragma unroll 1
        for (int hh = 0; hh < NHEAD; ++hh) {
            float q[16], k[16], v[16], o[16];
            tmem_ld16(trow + TM_Q + hh * HD, q);
            tmem_ld16(trow + TM_K + hh * HD, k);
            tmem_ld16(trow + TM_V + hh * HD, v);
            add_vec16(q, bq + hh * HD);
            add_vec16(k, bk + hh * HD);
            add_vec16(v, bv + hh * HD);
#pragma unroll
            for (int i = 0; i < 16; ++i) { q[i] *= 0.25f; o[i] = 0.f; }
            float m = -INFINITY, lsum = 0.f;
            for (int j = 0; j < N; ++j) {
                const int src = root_lane0 + j;
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) s += q[i] * __shfl_sync(0xffffffffu, k[i], src);
                const float mn = fmaxf(m, s);
                const float corr = __expf(m - mn), p = __expf(s - mn);
                lsum = lsum * corr + p;
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = o[i] * corr + p * __shfl_sync(0xffffffffu, v[i], src);
                m = mn;
            }
            const float inv = 1.f / lsum;
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] *= inv;
            store16(sT, row, hh * HD, H, o);
        }
        STAGE_SYNC();
        if (tid == 0) {  // x += sa Wo^T   (residual add = accumulate into the TMEM-resident stream)
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_X, aT, H, 0, w, H, 0, H, 128, true);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
        ln128(trow, TM_X, bo, g1, be1, false, true, sX, row);   // norm1
        STAGE_SYNC();
        if (tid == 0) {
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_ACC, aX, H, 0, w, H, 0, H, 128, false);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
#pragma unroll 1
        for (int cc = 0; cc < H; cc += 16) {  // f = relu(x W1^T + b1)
            float v[16];
            tmem_ld16(trow + TM_ACC + cc, v);
            add_vec16(v, b1 + cc);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            store16(sT, row, cc, H, v);
        }
        STAGE_SYNC();
        if (tid == 0) {  // x += f W2^T
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_X, aT, H, 0, w, H, 0, H, 128, true);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
        ln128(trow, TM_X, b2, g2, be2, false, true, sX, row);   // norm2
    }

    // ---- fc_dynamic on [h | onehot | attn]: Linear-LN-ReLU, Linear-LN-ReLU, Linear; residual (model.py:262-268)
    {
        const float *dv = vec + d.o_dyn;
        gather_h();   // h (bf16) back into sT; sX holds the attention output
        STAGE_SYNC();
        if (tid == 0) {
            uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_ACC, aT, H, 0, w, H, 0, H, 128, false);
            is.release(c);
            w = is.acquire(c + 1);
            issue_gemm(tmem + TM_ACC, aOne, KA, 0, w, KA, 0, KA, 128, true);
            is.release(c + 1);
            w = is.acquire(c + 2);
            issue_gemm(tmem + TM_ACC, aX, H, 0, w, H, 0, H, 128, true);
            is.release(c + 2);
            mma_commit(&bar_mma);
        }
        c += 3;
        WAIT_MMA();
        ln128(trow, TM_ACC, dv, dv + 128, dv + 256, true, false, sT, row);
        STAGE_SYNC();
        if (tid == 0) {
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_ACC, aT, H, 0, w, H, 0, H, 128, false);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
        ln128(trow, TM_ACC, dv + 384, dv + 512, dv + 640, true, false, sX, row);
        STAGE_SYNC();
        if (tid == 0) {
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_ACC, aX, H, 0, w, H, 0, H, 128, false);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
        float *nh = valid ? d.next_hidden + (size_t)root * (N * H) + (size_t)agent * H : nullptr;
#pragma unroll 1
        for (int cc = 0; cc < H; cc += 16) {  // next_hidden = update + hidden  (fp32 residual from the pool)
            float v[16];
            tmem_ld16(trow + TM_ACC + cc, v);
            add_vec16(v, dv + 768 + cc);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(hrow + cc + i);
                    v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
                    *reinterpret_cast<float4 *>(nh + cc + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
            store16(sT, row, cc, H, v);
        }
    }

    // ---- reward head: GraphNetNN on [next_hidden | onehot] (model.py:270-277) ----------------------------------
    {
        const float *rv = vec + d.o_rg;
        float y[GH];
        STAGE_SYNC();
        if (tid == 0) {
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_ACC, aT, H, 0, w, H + KA, 0, H, 128, false);
            issue_gemm(tmem + TM_ACC, aOne, KA, 0, w, H + KA, H, KA, 128, true);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
        gnn_layer(trow, rv, rv + 64, N, root_lane0, y);
#pragma unroll
        for (int cc = 0; cc < GH; cc += 16) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = y[cc + i];
            store16(sX, row, cc, GH, v);
        }
        STAGE_SYNC();
        if (tid == 0) {
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_ACC, aX, GH, 0, w, GH, 0, GH, 128, false);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
        gnn_layer(trow, rv + 128, rv + 192, N, root_lane0, y);
        const float r = gnn_head(y, N, root_lane0, rv + 256, rv + 256 + SUP * GH);
        if (valid && agent == 0) d.reward[root] = r;
    }
    // ---- value head: GraphNetNN on next_hidden (model.py:359) ---------------------------------------------------
    {
        const float *vv = vec + d.o_vg;
        float y[GH];
        STAGE_SYNC();
        if (tid == 0) {
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_ACC, aT, H, 0, w, H, 0, H, 128, false);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
        gnn_layer(trow, vv, vv + 64, N, root_lane0, y);
#pragma unroll
        for (int cc = 0; cc < GH; cc += 16) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = y[cc + i];
            store16(sX, row, cc, GH, v);
        }
        STAGE_SYNC();
        if (tid == 0) {
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_ACC, aX, GH, 0, w, GH, 0, GH, 128, false);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
        gnn_layer(trow, vv + 128, vv + 192, N, root_lane0, y);
        const float val = gnn_head(y, N, root_lane0, vv + 256, vv + 256 + SUP * GH);
        if (valid && agent == 0) d.value[root] = val;
    }
    // ---- policy head: Linear(128,32)-LN-ReLU-Linear(32,A) per agent, then the driver's softmax / beta --------------
    {
        const float *pv = vec + d.o_pol;
        STAGE_SYNC();
        if (tid == 0) {
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_ACC, aT, H, 0, w, H, 0, H, PH, false);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
        {
            float p[PH];
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int cc = 0; cc < PH; cc += 16) {
                float v[16];
                tmem_ld16(trow + TM_ACC + cc, v);
                add_vec16(v, pv + cc);
#pragma unroll
                for (int i = 0; i < 16; ++i) { p[cc + i] = v[i]; sum += v[i]; sq += v[i] * v[i]; }
            }
            const float mean = sum * (1.f / PH);
            const float rstd = rsqrtf(fmaxf(sq * (1.f / PH) - mean * mean, 0.f) + 1e-5f);
#pragma unroll
            for (int cc = 0; cc < PH; cc += 16) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    v[i] = fmaxf((p[cc + i] - mean) * rstd * __ldg(pv + 32 + cc + i) + __ldg(pv + 64 + cc + i), 0.f);
                store16(sX, row, cc, PH, v);
            }
        }
        STAGE_SYNC();
        if (tid == 0) {
            const uint32_t w = is.acquire(c);
            issue_gemm(tmem + TM_ACC, aX, PH, 0, w, PH, 0, PH, d.NAP, false);
            is.release(c);
            mma_commit(&bar_mma);
        }
        ++c;
        WAIT_MMA();
        float lg[48];
#pragma unroll
        for (int cc = 0; cc < 48; cc += 16) {
            float v[16];
            if (cc < d.NAP) {
                tmem_ld16(trow + TM_ACC + cc, v);
                add_vec16(v, pv + 96 + cc);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) lg[cc + i] = (cc < d.NAP && cc + i < A) ? v[i] : -INFINITY;
        }
        if (valid) {
            float m = -INFINITY;
            int am = 0;
#pragma unroll
            for (int a = 0; a < 48; ++a)
                if (lg[a] > m) { m = lg[a]; am = a; }
            if (d.greedy) d.greedy[(size_t)root * N + agent] = am;
            if (d.logits_out) {
#pragma unroll
                for (int a = 0; a < 48; ++a)
                    if (a < A) d.logits_out[((size_t)root * N + agent) * A + a] = lg[a];
            }
            const int ta = (d.cur < 0) ? agent : (agent == d.cur ? 0 : -1);
            if (ta >= 0) {
                float s = 0.f, sb = 0.f;
                float pr[48];
#pragma unroll
                for (int a = 0; a < 48; ++a) { pr[a] = __expf(lg[a] - m); s += pr[a]; }
                const float invs = 1.f / s;
                const bool unit_tau = (d.inv_tau == 1.0f);
#pragma unroll
                for (int a = 0; a < 48; ++a) {
                    pr[a] *= invs;
                    sb += unit_tau ? pr[a] : ((a < A) ? __powf(pr[a], d.inv_tau) : 0.f);
                }
                const float invb = 1.f / sb;
                float *po = d.probs + ((size_t)root * d.Nt + ta) * A;
                float *bo_ = d.beta + ((size_t)root * d.Nt + ta) * A;
#pragma unroll
                for (int a = 0; a < 48; ++a)
                    if (a < A) {
                        po[a] = pr[a];
                        bo_[a] = (unit_tau ? pr[a] : __powf(pr[a], d.inv_tau)) * invb;
                    }
            }
        }
    }
#undef STAGE_SYNC
#undef WAIT_MMA
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace fused
}  // namespace maz
