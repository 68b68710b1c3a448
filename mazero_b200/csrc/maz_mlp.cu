// maz_mlp.cu -- recurrent_inference of the reference's MLP-family network (the matrix-game MAMuZeroNet,
// config/matrix/model.py:86-166,320-368) for the on-device search: plain fp32 SIMT, one CTA per root.
//
// The network is a handful of tiny dense layers over the CONCATENATED state of all agents (2 agents x 64 for the
// matrix games, BASELINE configs[0]: 16 roots): there is no tile a tensor core could fill, so this is a
// latency-oriented kernel -- activations live in shared memory, weights are read pre-transposed ([in][out], one
// coalesced request per input feature across the output lanes) and stay L2-resident (< 100 KB), reductions are
// warp shuffles.  It produces exactly what the fused SMAC kernel produces (infer_fused.cuh): next hidden state in
// the pool, reward / value scalars after the inverse support transform (core/config.py:430-442,463-499), softmax
// and beta of the tree agents (mcts_sampled.py:158-161), greedy actions of every agent (mcts_sampled.py:137-145).
#include <cuda_runtime.h>

#include <string>

#include "../../include/maz_infer.h"

namespace maz { int set_last_error(int code, const std::string &msg); }

namespace {

constexpr int NT = 128;                 // threads per CTA
constexpr int MAXW = MAZ_MLP_MAXWIDTH;  // widest activation vector

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum, result broadcast to every thread; `red` = NT/32 + 1 floats of shared scratch
__device__ __forceinline__ float block_sum(float v, float *red)
{
    v = warp_sum(v);
    __syncthreads();                                  // scratch free
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += red[w];
    return t;
}
__device__ __forceinline__ float block_max(float v, float *red)
{
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) t = fmaxf(t, red[w]);
    return t;
}

// y[0..out) = act(W x + b).  x, y in shared memory (distinct buffers).  When the layer is narrow (out <= NT/2)
// the input range is split over NT/out thread groups and the partial sums are combined in group order.
__device__ void dense(const maz_mlp_layer &L, const float *x, float *y, float *part, float *red)
{
    const int tid = threadIdx.x, in = L.in, out = L.out;
    const float *__restrict__ wt = L.wt;
    int groups = 1;
    if (out <= NT / 2) {
        groups = NT / out;
        if (groups > 8) groups = 8;
    }
    if (groups == 1) {
        for (int j = tid; j < out; j += NT) {
            float acc = __ldg(L.b + j);
#pragma unroll 4
            for (int i = 0; i < in; ++i) acc = fmaf(x[i], __ldg(wt + (size_t)i * out + j), acc);
            y[j] = acc;
        }
    } else {
        const int g = tid / out, j = tid - g * out;
        if (g < groups) {
            const int per = (in + groups - 1) / groups, i0 = g * per, i1 = min(in, i0 + per);
            float acc = 0.f;
#pragma unroll 4
            for (int i = i0; i < i1; ++i) acc = fmaf(x[i], __ldg(wt + (size_t)i * out + j), acc);
            part[g * out + j] = acc;
        }
        __syncthreads();
        if (tid < out) {
            float acc = __ldg(L.b + tid);
            for (int q = 0; q < groups; ++q) acc += part[q * out + tid];
            y[tid] = acc;
        }
    }
    __syncthreads();
    if (L.kind == MAZ_MLP_LINEAR) return;
    // LayerNorm over the `out` features (biased variance, eps 1e-5: torch.nn.LayerNorm), with the ReLU before
    // (matrix mlp(): Linear, ReLU, LayerNorm -- config/matrix/model.py:44-46) or after (SMAC mlp(): Linear,
    // LayerNorm, ReLU -- config/smac/model.py:52-58)
    const bool relu_first = (L.kind == MAZ_MLP_RELU_LN);
    float s = 0.f;
    for (int j = tid; j < out; j += NT) {
        float v = y[j];
        if (relu_first) v = fmaxf(v, 0.f);
        s += v;
    }
    const float mean = block_sum(s, red) / (float)out;
    float q = 0.f;
    for (int j = tid; j < out; j += NT) {
        float v = y[j];
        if (relu_first) v = fmaxf(v, 0.f);
        v -= mean;
        q += v * v;
    }
    const float rstd = rsqrtf(block_sum(q, red) / (float)out + 1e-5f);
    for (int j = tid; j < out; j += NT) {
        float v = y[j];
        if (relu_first) v = fmaxf(v, 0.f);
        v = (v - mean) * rstd * __ldg(L.ln_w + j) + __ldg(L.ln_b + j);
        if (!relu_first) v = fmaxf(v, 0.f);
        y[j] = v;
    }
    __syncthreads();
}

// runs the layers of `net` on `x` (length net.l[0].in); returns the buffer holding the result
__device__ float *run_net(const maz_mlp_net &net, float *x, float *a, float *b, float *part, float *red)
{
    float *src = x, *dst = a;
    for (int l = 0; l < net.n; ++l) {
        dense(net.l[l], src, dst, part, red);
        src = dst;
        dst = (dst == a) ? b : a;
    }
    return src;
}

// softmax(logits) . support -> inv_h   (core/config.py:430-442, 463-499); a support of size 1 is the scalar head
// of use_vectorization=False (core/config.py:378-381,488,495): the network output is the value itself
__device__ float inverse_support(const float *logits, int size, int smin, float *red)
{
    const int tid = threadIdx.x;
    if (size == 1) return logits[0];
    float m = -INFINITY;
    for (int j = tid; j < size; j += NT) m = fmaxf(m, logits[j]);
    m = block_max(m, red);
    float se = 0.f, sx = 0.f;
    for (int j = tid; j < size; j += NT) {
        const float e = expf(logits[j] - m);
        se += e;
        sx += e * (float)(smin + j);
    }
    se = block_sum(se, red);
    sx = block_sum(sx, red);
    const float x = sx / se, eps = 0.001f;
    const float r = (sqrtf(1.f + 4.f * eps * (fabsf(x) + 1.f + eps)) - 1.f) / (2.f * eps);
    float o = r * r - 1.f;
    if (x < 0.f) o = -o;
    if (o != o) o = 0.f;
    if (fabsf(o) < eps) o = 0.f;
    return o;
}

__global__ void __launch_bounds__(NT) k_mlp_recurrent(const maz_mlp_desc d)
{
    __shared__ float s_in[MAXW], s_a[MAXW], s_b[MAXW], s_state[MAXW], s_part[8 * (NT / 2)], s_red[NT / 32 + 1];
    const int tid = threadIdx.x, b = blockIdx.x;
    const int D = d.N * d.H, NA = d.N * d.A;
    const size_t prow = ((size_t)(d.idx_x ? d.idx_x[b] : 0) * d.B + b) * D;

    // input of the dynamics net: [hidden | one-hot joint action]   (config/matrix/model.py:334-343, 118-121)
    for (int i = tid; i < D; i += NT) s_in[i] = d.pool[prow + i];
    for (int i = tid; i < NA; i += NT) {
        const int n = i / d.A;
        int act;
        if (d.greedy_pool == nullptr || d.cur < 0) act = d.actions[b * d.N + n];
        else if (n == d.cur) act = d.actions[b];                      // sequential-agent mode (mcts_sampled.py:116-147)
        else if (n < d.cur) act = d.factor ? d.factor[b * d.N + n] : 0;
        else act = d.greedy_pool[((size_t)(d.idx_x ? d.idx_x[b] : 0) * d.B + b) * d.N + n];
        s_in[D + i] = (act == i - n * d.A) ? 1.f : 0.f;
    }
    __syncthreads();
    float *dyn = run_net(d.dyn, s_in, s_a, s_b, s_part, s_red);
    for (int i = tid; i < D; i += NT) {
        const float v = dyn[i] + s_in[i];                          // state += pre_state (:122)
        s_state[i] = v;
        d.next_hidden[(size_t)b * D + i] = v;
    }
    __syncthreads();
    for (int i = tid; i < D; i += NT) s_in[i] = s_state[i];       // reward input: [state | one-hot] (:124)
    __syncthreads();
    {
        float *r = run_net(d.rew, s_in, s_a, s_b, s_part, s_red);
        const float v = inverse_support(r, d.reward_support_size, d.reward_support_min, s_red);
        if (tid == 0) d.reward[b] = v;
    }
    __syncthreads();
    {
        float *r = run_net(d.val, s_state, s_a, s_b, s_part, s_red);
        const float v = inverse_support(r, d.value_support_size, d.value_support_min, s_red);
        if (tid == 0) d.value[b] = v;
    }
    __syncthreads();
    // per-agent policy head on that agent's slice of the state (config/matrix/model.py:165)
    const int lane = tid & 31, warp = tid >> 5;
    for (int n = 0; n < d.N; ++n) {
        float *lg = run_net(d.pol, s_state + n * d.H, s_a, s_b, s_part, s_red);
        if (warp == 0) {
            float m = -INFINITY;
            int am = 0x7fffffff;
            for (int a = lane; a < d.A; a += 32) {
                const float v = lg[a];
                if (d.logits_out) d.logits_out[((size_t)b * d.N + n) * d.A + a] = v;
                if (v > m) { m = v; am = a; }
            }
            // first maximal index (np.argmax, mcts_sampled.py:145)
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, m, o);
                const int oa = __shfl_xor_sync(0xffffffffu, am, o);
                if (om > m || (om == m && oa < am)) { m = om; am = oa; }
            }
            if (d.greedy && lane == 0) d.greedy[b * d.N + n] = am;
            const int slot = (d.cur < 0) ? n : (n == d.cur ? 0 : -1);
            if (slot >= 0) {
                float se = 0.f;
                for (int a = lane; a < d.A; a += 32) se += expf(lg[a] - m);
                se = warp_sum(se);
                float sb = 0.f;
                for (int a = lane; a < d.A; a += 32) {
                    const float p = expf(lg[a] - m) / se;
                    sb += (d.inv_tau == 1.f) ? p : powf(p, d.inv_tau);
                }
                sb = warp_sum(sb);
                for (int a = lane; a < d.A; a += 32) {
                    const float p = expf(lg[a] - m) / se;
                    const size_t o = ((size_t)b * d.Nt + slot) * d.A + a;
                    d.probs[o] = p;
                    d.beta[o] = ((d.inv_tau == 1.f) ? p : powf(p, d.inv_tau)) / sb;
                }
            }
        }
        __syncthreads();
    }
}

bool net_ok(const maz_mlp_net &n, int in, int out)
{
    if (n.n < 1 || n.n > MAZ_MLP_MAXLAYERS) return false;
    int w = in;
    for (int l = 0; l < n.n; ++l) {
        const maz_mlp_layer &L = n.l[l];
        if (L.in != w || L.out < 1 || L.out > MAXW || !L.wt || !L.b) return false;
        if (L.kind != MAZ_MLP_LINEAR && (!L.ln_w || !L.ln_b)) return false;
        if (L.kind < 0 || L.kind > 2) return false;
        w = L.out;
    }
    return out < 0 || w == out;
}

}   // namespace

extern "C" int maz_mlp_recurrent(const maz_mlp_desc *d, void *stream)
{
    if (!d) return maz::set_last_error(1, "maz_mlp_recurrent: NULL descriptor");
    if (d->B <= 0 || d->N <= 0 || d->A <= 0 || d->H <= 0 || d->N * d->H + d->N * d->A > MAXW || d->A > 255)
        return maz::set_last_error(3, "maz_mlp_recurrent: unsupported shape (agents*hidden + agents*actions <= 640)");
    if (!d->pool || !d->actions || !d->next_hidden || !d->reward || !d->value || !d->probs || !d->beta)
        return maz::set_last_error(1, "maz_mlp_recurrent: NULL tensor");
    if (d->Nt != (d->cur < 0 ? d->N : 1) || d->cur >= d->N) return maz::set_last_error(1, "maz_mlp_recurrent: Nt / cur mismatch");
    const int D = d->N * d->H, NA = d->N * d->A;
    if (!net_ok(d->dyn, D + NA, D) || !net_ok(d->rew, D + NA, d->reward_support_size) ||
        !net_ok(d->val, D, d->value_support_size) || !net_ok(d->pol, d->H, d->A))
        return maz::set_last_error(3, "maz_mlp_recurrent: layer chain does not match the network shape");
    k_mlp_recurrent<<<(unsigned)d->B, NT, 0, static_cast<cudaStream_t>(stream)>>>(*d);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return maz::set_last_error(2, std::string("k_mlp_recurrent: ") + cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Generic row-wise MLP (the representation network of `initial_inference`: config/smac/model.py:176-195, 470-489;
// config/matrix/model.py:54-83, 323-329): y[r] = net(x[r]) for `rows` independent rows (one agent's observation each).
// RPC rows per CTA share every weight read; warp r normalises row r.
namespace {

constexpr int RPC = 4;     // rows per CTA (= warps per CTA)

__device__ void dense_rows(const maz_mlp_layer &L, const float (*x)[MAXW], float (*y)[MAXW])
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, in = L.in, out = L.out;
    if (L.kind != MAZ_MLP_LN_ONLY) {
        const float *__restrict__ wt = L.wt;
        for (int j = tid; j < out; j += NT) {
            float acc[RPC];
            const float b = __ldg(L.b + j);
#pragma unroll
            for (int r = 0; r < RPC; ++r) acc[r] = b;
#pragma unroll 4
            for (int i = 0; i < in; ++i) {
                const float w = __ldg(wt + (size_t)i * out + j);
#pragma unroll
                for (int r = 0; r < RPC; ++r) acc[r] = fmaf(x[r][i], w, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < RPC; ++r) y[r][j] = acc[r];
        }
    } else {
        for (int j = tid; j < out; j += NT)
#pragma unroll
            for (int r = 0; r < RPC; ++r) y[r][j] = x[r][j];
    }
    __syncthreads();
    if (L.kind == MAZ_MLP_LINEAR) return;
    // warp `warp` owns row `warp`
    const bool relu_first = (L.kind == MAZ_MLP_RELU_LN);
    float *row = y[warp];
    float s = 0.f;
    for (int j = lane; j < out; j += 32) {
        float v = row[j];
        if (relu_first) v = fmaxf(v, 0.f);
        s += v;
    }
    const float mean = warp_sum(s) / (float)out;
    float q = 0.f;
    for (int j = lane; j < out; j += 32) {
        float v = row[j];
        if (relu_first) v = fmaxf(v, 0.f);
        v -= mean;
        q += v * v;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)out + 1e-5f);
    for (int j = lane; j < out; j += 32) {
        float v = row[j];
        if (relu_first) v = fmaxf(v, 0.f);
        v = (v - mean) * rstd * __ldg(L.ln_w + j) + __ldg(L.ln_b + j);
        if (L.kind == MAZ_MLP_LN_RELU) v = fmaxf(v, 0.f);
        row[j] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NT) k_mlp_rows(const maz_mlp_net net, const float *__restrict__ x, int rows, float *__restrict__ y)
{
    static_assert(NT == 32 * RPC, "one warp per row");
    __shared__ float s_a[RPC][MAXW], s_b[RPC][MAXW];
    const int tid = threadIdx.x, r0 = blockIdx.x * RPC;
    const int in0 = net.l[0].in;
    for (int r = 0; r < RPC; ++r)
        for (int i = tid; i < in0; i += NT) s_a[r][i] = (r0 + r < rows) ? x[(size_t)(r0 + r) * in0 + i] : 0.f;
    __syncthreads();
    float (*src)[MAXW] = s_a, (*dst)[MAXW] = s_b;
    for (int l = 0; l < net.n; ++l) {
        dense_rows(net.l[l], src, dst);
        float (*t)[MAXW] = src;
        src = dst;
        dst = t;
    }
    const int outn = net.l[net.n - 1].out;
    for (int r = 0; r < RPC; ++r)
        if (r0 + r < rows)
            for (int j = tid; j < outn; j += NT) y[(size_t)(r0 + r) * outn + j] = src[r][j];
}

}   // namespace

extern "C" int maz_mlp_forward(const maz_mlp_net *net, const float *x, int rows, float *y, void *stream)
{
    if (!net || !x || !y || rows <= 0) return maz::set_last_error(1, "maz_mlp_forward: bad arguments");
    if (net->n < 1 || net->n > MAZ_MLP_MAXLAYERS) return maz::set_last_error(3, "maz_mlp_forward: 1..6 layers");
    int w = net->l[0].in;
    if (w < 1 || w > MAXW) return maz::set_last_error(3, "maz_mlp_forward: input wider than 640");
    for (int l = 0; l < net->n; ++l) {
        const maz_mlp_layer &L = net->l[l];
        if (L.in != w || L.out < 1 || L.out > MAXW || L.kind < 0 || L.kind > MAZ_MLP_LN_ONLY)
            return maz::set_last_error(3, "maz_mlp_forward: layer chain does not match");
        if (L.kind == MAZ_MLP_LN_ONLY ? (L.in != L.out) : (!L.wt || !L.b)) return maz::set_last_error(1, "maz_mlp_forward: NULL weights");
        if (L.kind != MAZ_MLP_LINEAR && (!L.ln_w || !L.ln_b)) return maz::set_last_error(1, "maz_mlp_forward: NULL LayerNorm affine");
        w = L.out;
    }
    k_mlp_rows<<<(unsigned)((rows + RPC - 1) / RPC), NT, 0, static_cast<cudaStream_t>(stream)>>>(*net, x, rows, y);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return maz::set_last_error(2, std::string("k_mlp_rows: ") + cudaGetErrorString(e));
    return 0;
}
