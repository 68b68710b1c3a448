/*
 * maz_tree.h -- C ABI of libmaz_b200.so: the B200 (sm_100a) batched sampled-MCTS tree engine.
 *
 * Drop-in boundary for MAZero's native tree object.  Every entry point replaces one method of the
 * reference's C++ `tree::CTree_batch` (core/mcts/ctree/ctree_sampled/lib/cnode.h:107-141), i.e. exactly
 * what its Cython binding `cytree.Tree_batch` binds (core/mcts/ctree/ctree_sampled/ctree.pxd:11-37,
 * cytree.pyx:7-247).  Plain pointers and sizes only; no torch / Python types.
 *
 * Conventions
 *   - all arrays are C-contiguous float32 / int32; (B,N,A) is row-major p[(b*N+n)*A+a]
 *     (reference: common_lib/utils.cpp:107-116);
 *   - functions return MAZ_OK (0) or a MAZ_ERR_* code; the message is in maz_last_error() (thread-local).
 *     The reference throws std::runtime_error -> Python RuntimeError (utils.cpp:8-18, ctree.pxd:14-20);
 *     a binding should raise RuntimeError(maz_last_error()) on a non-zero return;
 *   - functions WITHOUT a suffix take HOST pointers (what the reference's binding passes) and are
 *     synchronous: outputs are valid on return;
 *   - functions with the `_dev` suffix take DEVICE pointers, only enqueue work on the handle's stream
 *     (maz_tree_set_stream) and never synchronise: they can be captured into a CUDA graph.  Device-side
 *     invariant failures are reported at the next synchronous call or by maz_tree_check();
 *   - a handle owns one HBM arena (structure-of-arrays node pools, one slab per tree, one warp per tree);
 *     there is no global state: several handles / processes can share a GPU.
 *   - there is NO CPU fallback: every call fails with MAZ_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MAZ_TREE_H
#define MAZ_TREE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAZ_ABI_VERSION 1

enum {
    MAZ_OK = 0,
    MAZ_ERR_INVALID = 1,     /* bad argument                                                    */
    MAZ_ERR_CUDA = 2,        /* CUDA runtime failure (message has cudaGetErrorString)            */
    MAZ_ERR_UNSUPPORTED = 3, /* shape outside the kernels' limits (K > 32, A > 255, ...)         */
    MAZ_ERR_DEVICE = 4       /* device-side invariant failed (the reference's my_assert sites)   */
};

typedef struct maz_tree maz_tree;

int maz_abi_version(void);
const char *maz_last_error(void);

/* ---- construction -------------------------------------------------------------------------------
 * replaces CTree_batch::CTree_batch (cnode.cpp:553-577) / Tree_batch.__cinit__ (cytree.pyx:11-16).
 * Tree i uses std::mt19937 seed `random_seed*2333 + root_index_offset + i` (cnode.cpp:574; the
 * offset is 0 in the reference and is the rank's first global root index when roots are sharded). */
int maz_tree_create(maz_tree **out, int root_num, int agent_num, int action_space_size, int sampled_times,
                    int simulation_num, float tree_value_stat_delta_lb, unsigned int random_seed, float rho,
                    float lam);
int maz_tree_create_ex(maz_tree **out, int root_num, int agent_num, int action_space_size, int sampled_times,
                       int simulation_num, float tree_value_stat_delta_lb, unsigned int random_seed, float rho,
                       float lam, int device, unsigned int root_index_offset);
/* replaces CTree_batch::~CTree_batch (cnode.cpp:579-587) / Tree_batch.__dealloc__ (cytree.pyx:18-19) */
void maz_tree_destroy(maz_tree *t);
/* re-arm an existing arena for a new search with a new seed (same shapes): what constructing a fresh
 * Tree_batch per batch_search does in the reference (mcts_sampled.py:89), without the allocation. */
int maz_tree_reset(maz_tree *t, unsigned int random_seed, float tree_value_stat_delta_lb, float rho, float lam,
                   unsigned int root_index_offset);
/* stream (cudaStream_t) all work of this handle is enqueued on; NULL = the legacy default stream. */
int maz_tree_set_stream(maz_tree *t, void *cuda_stream);
/* synchronise the stream and surface device-side invariant failures (MAZ_ERR_DEVICE). */
int maz_tree_check(maz_tree *t);
/* upload the pUCT tables for (pb_c_base, pb_c_init) ahead of time (needed before graph capture of
 * maz_tree_batch_selection_dev; done lazily by the synchronous entry points). */
int maz_tree_set_puct(maz_tree *t, float pb_c_base, float pb_c_init);

/* ---- the four step entry points (host pointers) -------------------------------------------------
 * replaces CTree_batch::prepare (cnode.cpp:589-614) / Tree_batch.prepare (cytree.pyx:21-47) */
int maz_tree_prepare(maz_tree *t, const float *rewards, const float *values, const float *policy_probs,
                     const float *beta, int sampled_times, float noise_eps, const float *noises);
/* replaces CTree_batch::cbatch_selection (cnode.cpp:616-642) / Tree_batch.batch_selection
 * (cytree.pyx:50-67).  idx_x (B,), idx_y (B,), act (B,N). */
int maz_tree_batch_selection(maz_tree *t, float pb_c_base, float pb_c_init, float discount, int *idx_x,
                             int *idx_y, int *act);
/* replaces CTree_batch::cbatch_expansion_and_backup (cnode.cpp:644-670) /
 * Tree_batch.batch_expansion_and_backup (cytree.pyx:69-91) */
int maz_tree_batch_expansion_and_backup(maz_tree *t, int hidden_state_index_x, float discount, int sampled_times,
                                        const float *rewards, const float *values, const float *policy_probs,
                                        const float *beta);

/* ---- the same three, device pointers, asynchronous on the handle's stream ------------------------ */
int maz_tree_prepare_dev(maz_tree *t, const float *rewards, const float *values, const float *policy_probs,
                         const float *beta, int sampled_times, float noise_eps, const float *noises);
int maz_tree_batch_selection_dev(maz_tree *t, float pb_c_base, float pb_c_init, float discount, int *idx_x,
                                 int *idx_y, int *act);
int maz_tree_batch_expansion_and_backup_dev(maz_tree *t, int hidden_state_index_x, float discount,
                                            int sampled_times, const float *rewards, const float *values,
                                            const float *policy_probs, const float *beta);

/* fused step for the on-device search loop: cbatch_expansion_and_backup of simulation s followed by
 * cbatch_selection of simulation s+1 in ONE launch (cnode.cpp:644-670 then 616-642). */
int maz_tree_expansion_backup_selection_dev(maz_tree *t, int hidden_state_index_x, float discount, int sampled_times,
                                            const float *rewards, const float *values, const float *policy_probs,
                                            const float *beta, float pb_c_base, float pb_c_init, int *idx_x, int *idx_y,
                                            int *act);

/* ---- readouts -------------------------------------------------------------------------------------
 * replaces CTree_batch::get_roots_values / get_roots_marginal_visit_count / get_roots_marginal_priors
 * (cnode.cpp:672-700) and get_num_children_of_root (cnode.cpp:702-705). */
int maz_tree_get_roots_values(maz_tree *t, float *out /* (B,) */);
int maz_tree_get_roots_marginal_visit_count(maz_tree *t, int *out /* (B,N,A) */);
int maz_tree_get_roots_marginal_priors(maz_tree *t, float *out /* (B,N,A) */);
int maz_tree_get_roots_num_children(maz_tree *t, int *out /* (B,) */);
/* replaces the 11 per-root getters get_root_sampled_* (cnode.cpp:707-781) that cytree.pyx:111-241
 * loops over root by root: ONE call fills padded (B,K[,N]) arrays, K = sampled_times of the handle;
 * rows c >= num_children[b] are zero.  Any output pointer may be NULL.
 * imp_ratio = beta_hat/beta*pred_prob (cnode.cpp:137); qvalues = reward + discount*value (cnode.cpp:66). */
int maz_tree_readout(maz_tree *t, float discount, float *values /* (B,) */, int *marginal_visit_count /* (B,N,A) */,
                     float *marginal_priors /* (B,N,A) */, int *num_children /* (B,) */, int *actions /* (B,K,N) */,
                     int *visit_count /* (B,K) */, float *pred_probs, float *beta, float *beta_hat, float *priors,
                     float *imp_ratio, float *pred_values, float *mcts_values, float *rewards, float *qvalues);
int maz_tree_readout_dev(maz_tree *t, float discount, float *values, int *marginal_visit_count,
                         float *marginal_priors, int *num_children, int *actions, int *visit_count, float *pred_probs,
                         float *beta, float *beta_hat, float *priors, float *imp_ratio, float *pred_values,
                         float *mcts_values, float *rewards, float *qvalues);

/* ---- statistics (synchronous) ---------------------------------------------------------------------
 * per tree: nodes allocated (CTree::tot_nodes), depth of the last selection (SearchResult::search_len),
 * sum of selection depths so far -- the d-bar / C-bar inputs of the roofline byte formula. */
int maz_tree_stats(maz_tree *t, int *tot_nodes /* (B,) */, int *last_search_len /* (B,) */,
                   long long *sum_search_len /* scalar */, long long *sum_expanded /* scalar */);
/* profiling: SM-cycle timestamps of tree 0's phases in the next kernels are written to `dev_clock64` (device
 * pointer to 64 long longs), NULL switches it off */
int maz_tree_set_debug_clock(maz_tree *t, long long *dev_clock64);
/* bytes of HBM held by the handle's arena */
size_t maz_tree_arena_bytes(const maz_tree *t);

#ifdef __cplusplus
}
#endif
#endif /* MAZ_TREE_H */
