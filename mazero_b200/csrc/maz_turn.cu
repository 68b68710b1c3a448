// maz_turn.cu -- the steps either side of the search loop as device kernels (include/maz_turn.h):
// root preparation (mcts_sampled.py:57-106), an on-device Dirichlet sampler, and the workers' per-agent turn logic
// (selfplay_worker.py:230-257, reanalyze_worker.py:298-327).  All are tiny, latency-oriented kernels: one warp per
// (root, agent) row for the softmax-like work, one thread per root for the scalar turn logic.  Their point is not
// their own speed but that NOTHING returns to the host between the N per-agent searches of an environment step.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/maz_turn.h"

namespace maz { int set_last_error(int code, const std::string &msg); }

namespace {

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// ---------------------------------------------------------------------------------------------------------------
// k_root_prepare: one warp per (root, agent) row of the logits.  Rows of tree agents produce probs / beta / noise;
// every row produces its greedy action.
constexpr int RP_WARPS = 4;
constexpr int RP_MAXA = 256;

__global__ void __launch_bounds__(RP_WARPS * 32) k_root_prepare(const float *__restrict__ logits, const float *__restrict__ legal,
                                                                const float *__restrict__ noises_in, int B, int NA, int A, int cur,
                                                                float noise_eps, float inv_tau, float *__restrict__ probs,
                                                                float *__restrict__ beta, float *__restrict__ noises_out,
                                                                int *__restrict__ greedy)
{
    __shared__ float s_p[RP_WARPS][RP_MAXA], s_n[RP_WARPS][RP_MAXA];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row = blockIdx.x * RP_WARPS + w;          // (b, n) over all agents
    if (row >= B * NA) return;
    const int b = row / NA, n = row - b * NA;
    const float *lg = logits + (size_t)row * A;

    float m = -INFINITY;
    int am = 0x7fffffff;
    for (int a = lane; a < A; a += 32) {
        const float v = lg[a];
        if (v > m) { m = v; am = a; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {                       // first maximal index (np.argmax)
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    if (greedy && lane == 0) greedy[row] = am;

    const int slot = (cur < 0) ? n : (n == cur ? 0 : -1);
    if (slot < 0) return;
    const int Nt = (cur < 0) ? NA : 1;
    const size_t o = ((size_t)b * Nt + slot) * A;
    const float *mk = legal ? legal + (size_t)row * A : nullptr;
    float *P = s_p[w], *Nz = s_n[w];

    // softmax (mcts_sampled.py:64-65)
    float se = 0.f;
    for (int a = lane; a < A; a += 32) {
        const float e = expf(lg[a] - m);
        P[a] = e;
        se += e;
    }
    se = warp_sum(se);
    float sp = 0.f, sn = 0.f;
    for (int a = lane; a < A; a += 32) {
        float p = P[a] / se, z = noises_in[o + a];
        if (mk) {                                        // :73-83  (the += of a float64 product rounds once, to float32)
            const float k = mk[a];
            p = (float)((double)(p * k) + (double)k * 1e-4);
            z = (float)((double)(z * k) + (double)k * 1e-4);
        }
        P[a] = p;
        Nz[a] = z;
        sp += p;
        sn += z;
    }
    sp = warp_sum(sp);
    sn = warp_sum(sn);
    float sb = 0.f;
    const float one_m = 1.f - noise_eps;
    for (int a = lane; a < A; a += 32) {
        float p = P[a], z = Nz[a];
        if (mk) {
            p = p / sp;
            z = z / sn;
        }
        probs[o + a] = p;
        noises_out[o + a] = z;
        float bt = p * one_m + z * noise_eps;            // :93-100
        if (inv_tau != 1.f) bt = powf(bt, inv_tau);
        if (mk) bt *= mk[a];
        P[a] = bt;
        sb += bt;
    }
    sb = warp_sum(sb);
    for (int a = lane; a < A; a += 32) beta[o + a] = P[a] / sb;
}

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter-based, no state to carry between launches.
struct Philox {
    uint32_t c[4], k[2];
    __device__ Philox(unsigned long long seed, unsigned long long stream, uint32_t ctr)
    {
        k[0] = (uint32_t)seed; k[1] = (uint32_t)(seed >> 32);
        c[0] = ctr; c[1] = 0; c[2] = (uint32_t)stream; c[3] = (uint32_t)(stream >> 32);
    }
    __device__ void next(uint32_t out[4])
    {
        uint32_t x0 = c[0], x1 = c[1], x2 = c[2], x3 = c[3], k0 = k[0], k1 = k[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            x0 = hi1 ^ x1 ^ k0; x1 = lo1; x2 = hi0 ^ x3 ^ k1; x3 = lo0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
        ++c[0];
    }
};
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.f / 16777216.f); }   // (0,1)

// Gamma(alpha, 1), alpha > 0: Marsaglia & Tsang (2000) on alpha+1, then * U^(1/alpha)
__device__ float gamma_sample(Philox &g, float alpha)
{
    const float a1 = alpha + 1.f, d = a1 - 1.f / 3.f, c = rsqrtf(9.f * d);
    uint32_t r[4];
    for (int attempt = 0; attempt < 64; ++attempt) {
        g.next(r);
        const float u1 = u01(r[0]), u2 = u01(r[1]), u3 = u01(r[2]), u4 = u01(r[3]);
        const float x = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);          // Box-Muller
        float v = 1.f + c * x;
        if (v <= 0.f) continue;
        v = v * v * v;
        if (logf(u3) < 0.5f * x * x + d - d * v + d * logf(v)) return d * v * powf(u4, 1.f / alpha);
    }
    return d;   // unreachable in practice (acceptance > 95 % per attempt)
}

__global__ void __launch_bounds__(128) k_dirichlet(float *__restrict__ out, int rows, int A, float alpha, unsigned long long seed)
{
    const int lane = threadIdx.x & 31, row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    float s = 0.f;
    for (int a = lane; a < A; a += 32) {
        Philox g(seed, (unsigned long long)row * 256ull + (unsigned)a, 0u);
        const float v = fmaxf(gamma_sample(g, alpha), 1e-30f);
        out[(size_t)row * A + a] = v;
        s += v;
    }
    s = warp_sum(s);
    for (int a = lane; a < A; a += 32) out[(size_t)row * A + a] /= s;
}

// ---------------------------------------------------------------------------------------------------------------
// k_agent_turn: one thread per root; everything in float64 exactly as numpy / Python floats do it on the host.
__device__ __forceinline__ double powd(double x, double e)
{
    if (e == 1.0) return x;                  // temperature 1 (x ** 1.0 is exact)
    if (e == 2.0) return x * x;              // temperature 0.5 / 0.25: exact for counts < 2^13
    if (e == 4.0) { const double y = x * x; return y * y; }
    return pow(x, e);
}

__global__ void __launch_bounds__(128) k_agent_turn(int mode, int B, int NA, int A, int K, int agent,
                                                    const int *__restrict__ num_children, const int *__restrict__ s_actions,
                                                    const int *__restrict__ s_visits, const int *__restrict__ m_visits,
                                                    const float *__restrict__ legal, double inv_t, const double *__restrict__ uniforms,
                                                    float eps, const float *__restrict__ eps_u, const int *__restrict__ rand_act,
                                                    int *__restrict__ actions, double *__restrict__ dist, double *__restrict__ prob_prod,
                                                    double *__restrict__ entropy)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int *mv = m_visits + (size_t)b * A;
    const float *la = legal ? legal + ((size_t)b * NA + agent) * A : nullptr;
    long long tot = 0;
    for (int a = 0; a < A; ++a) tot += mv[a];
    double *dd = dist + ((size_t)b * NA + agent) * A;
    // visit distribution of this agent's search (reanalyze_worker.py:319-327; selfplay_worker.py:283-286)
    if (tot > 0) {
        for (int a = 0; a < A; ++a) dd[a] = (double)mv[a] / (double)tot;
    } else {
        double nl = 0.0;
        for (int a = 0; a < A; ++a) nl += la ? (double)la[a] : 1.0;
        for (int a = 0; a < A; ++a) dd[a] = (nl > 0.0) ? (la ? (double)la[a] : 1.0) / nl : (a == 0 ? 1.0 : 0.0);
    }
    int action = 0;
    const int C = num_children[b];
    if (mode == MAZ_TURN_GREEDY) {
        // np.argmax(marginal_visits * legal_actions): first maximum (reanalyze_worker.py:309)
        double best = -1.0;
        for (int a = 0; a < A; ++a) {
            const double v = (double)mv[a] * (la ? (double)la[a] : 1.0);
            if (v > best) { best = v; action = a; }
        }
    } else {
        // select_action (core/utils.py:289-319) + np_random.choice (cdf.searchsorted(u, side='right'))
        const int *cnt = s_visits + (size_t)b * K;
        double total = 0.0;
        for (int i = 0; i < C; ++i) total += powd((double)cnt[i], inv_t);
        double run = 0.0, last = 0.0;
        for (int i = 0; i < C; ++i) last += powd((double)cnt[i], inv_t) / total;     // cdf[-1]
        const double u = uniforms[b];
        int pos = 0;
        double ent_sum = 0.0, psum = 0.0;
        for (int i = 0; i < C; ++i) psum += powd((double)cnt[i], inv_t) / total;
        for (int i = 0; i < C; ++i) {
            const double p = powd((double)cnt[i], inv_t) / total;
            run += p;
            if (run / last <= u) pos = i + 1;
            const double q = p / psum;                                                 // scipy.stats.entropy renormalises
            if (q > 0.0) ent_sum -= q * log(q);
        }
        if (pos > C - 1) pos = C - 1;
        action = (C > 0) ? s_actions[(size_t)b * K + pos] : 0;
        if (entropy) entropy[(size_t)b * NA + agent] = ent_sum / log(2.0);
        // eps_greedy_action (core/utils.py:322-334) with injected randoms
        if (eps_u && rand_act && eps_u[b] < eps) action = rand_act[b];
    }
    actions[(size_t)b * NA + agent] = action;
    const double pa = dd[action];
    prob_prod[b] = (agent == 0) ? 1.0 * pa : prob_prod[b] * pa;                       // reanalyze_worker.py:334-345
}

}   // namespace

static int launch_err(const char *name)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return maz::set_last_error(2, std::string(name) + ": " + cudaGetErrorString(e));
    return 0;
}

extern "C" int maz_root_prepare_dev(const float *logits, const float *legal, const float *noises_in, int B, int n_agents, int A,
                                    int cur, float noise_eps, float inv_tau, float *probs, float *beta, float *noises_out,
                                    int *greedy, void *stream)
{
    if (!logits || !noises_in || !probs || !beta || !noises_out) return maz::set_last_error(1, "maz_root_prepare_dev: NULL tensor");
    if (B <= 0 || n_agents <= 0 || A <= 0 || cur >= n_agents) return maz::set_last_error(1, "maz_root_prepare_dev: bad shape");
    if (A > RP_MAXA) return maz::set_last_error(3, "maz_root_prepare_dev: action_space_size > 256");
    const int rows = B * n_agents;
    k_root_prepare<<<(rows + RP_WARPS - 1) / RP_WARPS, RP_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, legal, noises_in, B, n_agents, A, cur, noise_eps, inv_tau, probs, beta, noises_out, greedy);
    return launch_err("k_root_prepare");
}

extern "C" int maz_dirichlet_dev(float *out, int rows, int A, float alpha, unsigned long long seed, void *stream)
{
    if (!out || rows <= 0 || A <= 0 || A > 256 || !(alpha > 0.f)) return maz::set_last_error(1, "maz_dirichlet_dev: bad arguments");
    k_dirichlet<<<(rows + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(out, rows, A, alpha, seed);
    return launch_err("k_dirichlet");
}

extern "C" int maz_agent_turn_dev(int mode, int B, int n_agents, int A, int K, int agent, const int *num_children,
                                  const int *sampled_actions, const int *sampled_visit_count, const int *marginal_visit_count,
                                  const float *legal, double inv_temperature, const double *uniforms, float greedy_epsilon,
                                  const float *eps_u, const int *random_action, int *actions, double *policy_dist,
                                  double *prob_prod, double *entropy, void *stream)
{
    if (mode != MAZ_TURN_GREEDY && mode != MAZ_TURN_SAMPLE) return maz::set_last_error(1, "maz_agent_turn_dev: bad mode");
    if (B <= 0 || n_agents <= 0 || A <= 0 || K <= 0 || agent < 0 || agent >= n_agents)
        return maz::set_last_error(1, "maz_agent_turn_dev: bad shape");
    if (!num_children || !sampled_actions || !sampled_visit_count || !marginal_visit_count || !actions || !policy_dist || !prob_prod)
        return maz::set_last_error(1, "maz_agent_turn_dev: NULL tensor");
    if (mode == MAZ_TURN_SAMPLE && !uniforms) return maz::set_last_error(1, "maz_agent_turn_dev: SAMPLE needs uniforms");
    k_agent_turn<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        mode, B, n_agents, A, K, agent, num_children, sampled_actions, sampled_visit_count, marginal_visit_count, legal,
        inv_temperature, uniforms, greedy_epsilon, eps_u, random_action, actions, policy_dist, prob_prod, entropy);
    return launch_err("k_agent_turn");
}
