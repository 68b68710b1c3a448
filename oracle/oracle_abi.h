/*
 * oracle_abi.h -- TEST INFRASTRUCTURE, not product code.
 *
 * One host-pointer C ABI shared by the two CPU checkers of the batched sampled-MCTS tree:
 *   - oracle/ref_shim.cpp  -> oracle/_ref/libmazref.so   (prefix mazref_): the reference's OWN C++
 *     (core/mcts/ctree/ctree_sampled/lib/cnode.cpp + core/mcts/ctree/common_lib/utils.cpp) compiled
 *     where it lies under /root/reference;
 *   - oracle/maz_oracle.c  -> oracle/libmazoracle.so     (prefix mazo_): a plain-C restatement.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * either library.  The product (mazero_b200/) never does.
 *
 * Every function mirrors one method of the reference's CTree_batch (cnode.h:107-141) /
 * cytree.Tree_batch (cytree.pyx:7-247).  Arrays are C-contiguous fp32/int32, (B,N,A) row-major.
 * Functions return 0 on success, non-zero on error (message via <pfx>_last_error()).
 */
#ifndef MAZ_ORACLE_ABI_H
#define MAZ_ORACLE_ABI_H

#ifdef __cplusplus
extern "C" {
#endif

#define MAZ_ORACLE_DECLARE(PFX)                                                                         \
    void *PFX##_create(int root_num, int agent_num, int action_space_size, int sampled_times,           \
                       int simulation_num, float tree_value_stat_delta_lb, unsigned int random_seed,     \
                       float rho, float lam);                                                            \
    void PFX##_destroy(void *h);                                                                         \
    int PFX##_prepare(void *h, const float *rewards, const float *values, const float *policy_probs,     \
                      const float *beta, int sampled_times, float noise_eps, const float *noises);       \
    int PFX##_batch_selection(void *h, float pb_c_base, float pb_c_init, float discount, int *idx_x,      \
                              int *idx_y, int *act);                                                      \
    int PFX##_batch_expansion_and_backup(void *h, int hidden_state_index_x, float discount,               \
                                         int sampled_times, const float *rewards, const float *values,    \
                                         const float *policy_probs, const float *beta);                   \
    int PFX##_get_roots_values(void *h, float *out);                                                      \
    int PFX##_get_roots_marginal_visit_count(void *h, int *out);                                          \
    int PFX##_get_roots_marginal_priors(void *h, float *out);                                             \
    int PFX##_get_roots_num_children(void *h, int *out);                                                  \
    /* padded per-root-child readout: arrays are (B,K_pad[,N]); rows >= num_children[b] untouched */      \
    int PFX##_readout(void *h, float discount, int k_pad, int *actions, int *visits, float *pred_probs,   \
                      float *beta, float *beta_hat, float *priors, float *imp_ratio, float *pred_values,  \
                      float *mcts_values, float *rewards, float *qvalues);                                \
    /* per-tree statistics: nodes allocated so far, path depth of the last selection */                   \
    int PFX##_stats(void *h, int *tot_nodes, int *last_search_len);                                       \
    const char *PFX##_last_error(void);

MAZ_ORACLE_DECLARE(mazref)
MAZ_ORACLE_DECLARE(mazo)

#ifdef __cplusplus
}
#endif
#endif
