/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Builds the REFERENCE's own tree engine into oracle/_ref/libmazref.so.  The two reference translation
 * units are #included from where they lie under /root/reference (the Makefile passes -I$(REFERENCE));
 * nothing of the reference is copied into this repository.  This mirrors what the reference's Cython
 * binding does (ctree.pxd:5-9 textually includes utils.cpp and cnode.cpp) with a flat C ABI instead of
 * a Python extension type, so that the compiled library can travel to the GPU box (no Cython, no
 * Python objects) and be driven by ctypes.
 *
 * Exception convention: the reference throws std::runtime_error from my_assert (utils.cpp:8-18); the
 * Cython layer turns those into Python RuntimeError (ctree.pxd:14-20).  Here they become return code 1
 * + mazref_last_error().
 */
#include "core/mcts/ctree/common_lib/utils.cpp"
#include "core/mcts/ctree/ctree_sampled/lib/cnode.cpp"

#include <string>
#include <stdexcept>

#include "oracle_abi.h"

static thread_local std::string g_err;

#define REF_GUARD(stmt)                          \
    try {                                        \
        stmt;                                    \
        return 0;                                \
    } catch (const std::exception &e) {          \
        g_err = e.what() ? e.what() : "unknown"; \
        return 1;                                \
    } catch (...) {                              \
        g_err = "unknown C++ exception";         \
        return 1;                                \
    }

extern "C" {

const char *mazref_last_error(void) { return g_err.c_str(); }

void *mazref_create(int root_num, int agent_num, int action_space_size, int sampled_times, int simulation_num,
                    float delta_lb, unsigned int seed, float rho, float lam)
{
    try {
        return new tree::CTree_batch(root_num, agent_num, action_space_size, sampled_times, simulation_num,
                                     delta_lb, seed, rho, lam);
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void mazref_destroy(void *h) { delete static_cast<tree::CTree_batch *>(h); }

int mazref_prepare(void *h, const float *rewards, const float *values, const float *probs, const float *beta,
                   int sampled_times, float noise_eps, const float *noises)
{
    auto *t = static_cast<tree::CTree_batch *>(h);
    REF_GUARD(t->prepare(const_cast<float *>(rewards), const_cast<float *>(values), const_cast<float *>(probs),
                         const_cast<float *>(beta), sampled_times, noise_eps, const_cast<float *>(noises)))
}

int mazref_batch_selection(void *h, float c_base, float c_init, float discount, int *idx_x, int *idx_y, int *act)
{
    auto *t = static_cast<tree::CTree_batch *>(h);
    REF_GUARD(t->cbatch_selection(c_base, c_init, discount, idx_x, idx_y, act))
}

int mazref_batch_expansion_and_backup(void *h, int hidx, float discount, int sampled_times, const float *rewards,
                                      const float *values, const float *probs, const float *beta)
{
    auto *t = static_cast<tree::CTree_batch *>(h);
    REF_GUARD(t->cbatch_expansion_and_backup(hidx, discount, sampled_times, const_cast<float *>(rewards),
                                             const_cast<float *>(values), const_cast<float *>(probs),
                                             const_cast<float *>(beta)))
}

int mazref_get_roots_values(void *h, float *out)
{
    auto *t = static_cast<tree::CTree_batch *>(h);
    REF_GUARD(t->get_roots_values(out))
}

int mazref_get_roots_marginal_visit_count(void *h, int *out)
{
    auto *t = static_cast<tree::CTree_batch *>(h);
    REF_GUARD(t->get_roots_marginal_visit_count(out))
}

int mazref_get_roots_marginal_priors(void *h, float *out)
{
    auto *t = static_cast<tree::CTree_batch *>(h);
    REF_GUARD(t->get_roots_marginal_priors(out))
}

int mazref_get_roots_num_children(void *h, int *out)
{
    auto *t = static_cast<tree::CTree_batch *>(h);
    for (int i = 0; i < t->root_num; ++i) out[i] = t->get_num_children_of_root(i);
    return 0;
}

int mazref_readout(void *h, float discount, int k_pad, int *actions, int *visits, float *pred_probs, float *beta,
                   float *beta_hat, float *priors, float *imp_ratio, float *pred_values, float *mcts_values,
                   float *rewards, float *qvalues)
{
    auto *t = static_cast<tree::CTree_batch *>(h);
    try {
        const int n = t->agent_num;
        for (int i = 0; i < t->root_num; ++i) {
            const int c = t->get_num_children_of_root(i);
            if (c > k_pad) throw std::runtime_error("mazref_readout: k_pad smaller than num_children");
            const size_t o = (size_t)i * k_pad;
            /* the same 11 per-root getters cytree.pyx:111-241 loops over */
            if (actions) t->get_root_sampled_actions(i, actions + o * n);
            if (visits) t->get_root_sampled_visit_count(i, visits + o);
            if (pred_probs) t->get_root_sampled_pred_probs(i, pred_probs + o);
            if (beta) t->get_root_sampled_beta(i, beta + o);
            if (beta_hat) t->get_root_sampled_beta_hat(i, beta_hat + o);
            if (priors) t->get_root_sampled_priors(i, priors + o);
            if (imp_ratio) t->get_root_sampled_imp_ratio(i, imp_ratio + o);
            if (pred_values) t->get_root_sampled_pred_values(i, pred_values + o);
            if (mcts_values) t->get_root_sampled_mcts_values(i, mcts_values + o);
            if (rewards) t->get_root_sampled_rewards(i, rewards + o);
            if (qvalues) t->get_root_sampled_qvalues(i, qvalues + o, discount);
        }
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

int mazref_stats(void *h, int *tot_nodes, int *last_search_len)
{
    auto *t = static_cast<tree::CTree_batch *>(h);
    for (int i = 0; i < t->root_num; ++i) {
        if (tot_nodes) tot_nodes[i] = t->trees[i].tot_nodes;
        if (last_search_len) last_search_len[i] = t->trees[i].result.search_len;
    }
    return 0;
}

} /* extern "C" */
