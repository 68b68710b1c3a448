/*
 * maz_search.h -- C ABI of the WHOLE search (libmaz_b200.so): what a host that binds the library calls instead of the body
 * of `SampledMCTS.batch_search` (core/mcts/tree_search/mcts_sampled.py:51-200).
 *
 * One call = root preparation (:57-106) -> Tree_batch construction + prepare (:89,106) -> `num_simulations` times
 * [batch_selection, gather parent hidden, prediction / recurrent_inference, softmax / beta, batch_expansion_and_backup]
 * (:114-172) -> the 13 root readouts (:176-191), all on the device, nothing returning to the host in between.
 * The host keeps what the reference draws from ITS random generator: the Dirichlet noise (:68) and the tree seed (:89).
 *
 * Two execution strategies behind the same call, chosen per handle (maz_search_strategy()):
 *   MAZ_SEARCH_PERSISTENT  one kernel per search: a CTA owns a group of roots for all simulations (fused inference on the
 *                          tensor cores + one warp per tree; csrc/search_persist.cuh).  Small / medium batches.
 *   MAZ_SEARCH_GRAPH       a CUDA graph of 2 kernels per simulation (fused inference, tree step), built inside the library
 *                          with the CUDA runtime and cached per set of search constants.  Large batches, MLP networks.
 * Results are identical bit for bit (trees are independent; same device code).
 *
 * Conventions as in maz_tree.h: MAZ_OK / MAZ_ERR_*, message in maz_last_error(); arrays C-contiguous float32 / int32;
 * `_dev` = device pointers, asynchronous on the handle's stream, never synchronises; without suffix = host pointers,
 * synchronous.  No CPU fallback.
 */
#ifndef MAZ_SEARCH_H
#define MAZ_SEARCH_H

#include <stddef.h>

#include "maz_infer.h"
#include "maz_tree.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct maz_search maz_search;

enum { MAZ_NET_SMAC = 0, MAZ_NET_MLP = 1 };
enum { MAZ_SEARCH_AUTO = 0, MAZ_SEARCH_PERSISTENT = 1, MAZ_SEARCH_GRAPH = 2 };

/* Shape of the problem and the network.  The descriptors carry the WEIGHT half (packed weights, parameter vector, chunk
 * tables, offsets, architecture sizes); their buffer pointers are ignored -- the handle owns every buffer of the loop. */
typedef struct maz_search_config {
    int B, N, A, K, S;          /* roots, agents of the network, actions per agent, sampled_times, num_simulations */
    int joint;                  /* 1: tree agent_num = N (upstream joint mode); 0: the fork's sequential-agent mode (agent_num = 1,
                                   mcts_sampled.py:53,89): one handle serves every current_agent_idx */
    int hidden;                 /* hidden-state size per agent */
    int device;
    int strategy;               /* MAZ_SEARCH_AUTO / _PERSISTENT / _GRAPH */
    int net_kind;               /* MAZ_NET_SMAC: smac_tc + smac_small;  MAZ_NET_MLP: mlp */
    const maz_infer_desc *smac_tc;      /* weights packed for the tcgen05 kernel (maz_infer_recurrent) */
    const maz_infer_desc *smac_small;   /* the same weights packed for the small-batch / persistent kernel */
    const maz_mlp_desc *mlp;
    float *pool;                /* optional caller-owned hidden-state pool (S+1, B, N*hidden) on the device; NULL = owned */
} maz_search_config;

/* replaces `cytree.Tree_batch(...)` + the `hidden_states_pool` list (mcts_sampled.py:86,89): allocated once, reused by every
 * search of this shape (the reference allocates both per batch_search call). */
int maz_search_create(maz_search **out, const maz_search_config *cfg);
void maz_search_destroy(maz_search *s);
int maz_search_set_stream(maz_search *s, void *cuda_stream);
int maz_search_strategy(const maz_search *s);           /* the resolved strategy */
size_t maz_search_device_bytes(const maz_search *s);    /* HBM held by the handle (arena + pool + buffers) */
float *maz_search_pool(maz_search *s);                  /* device pointer of the hidden-state pool; slot 0 = the roots */
maz_tree *maz_search_tree(maz_search *s);               /* the handle's tree batch (statistics, debugging) */
/* device pointers of the prepared root arrays of the last search: exactly what Tree_batch.prepare received (:106) */
int maz_search_root_arrays(maz_search *s, const float **probs, const float **beta, const float **noises);

/* The 13 readouts (mcts_sampled.py:176-191) as padded arrays: see maz_tree_readout(). */
typedef struct maz_search_readout {
    float *values;               /* (B,) */
    int *marginal_visit_count;   /* (B,Nt,A) */
    float *marginal_priors;      /* (B,Nt,A) */
    int *num_children;           /* (B,) */
    int *actions;                /* (B,K,Nt) */
    int *visit_count;            /* (B,K) */
    float *pred_probs, *beta, *beta_hat, *priors, *imp_ratio, *pred_values, *mcts_values, *rewards, *qvalues;   /* (B,K) */
} maz_search_readout;

/* One search = the arguments of batch_search (mcts_sampled.py:34-46) + the config attributes it reads (:51-53,89,114) + the
 * two host-RNG draws (:68,89). */
typedef struct maz_search_call {
    int cur;                     /* current_agent_idx, or -1 in joint mode */
    unsigned int seed;           /* np_random.choice(256)   (:89) */
    unsigned int root_index_offset;   /* first global root index of this shard (0 in the reference) */
    float noise_eps;             /* root_exploration_fraction, 0 when add_noise is False (:69-70) */
    float tau;                   /* sampled_tau */
    float pb_c_base, pb_c_init, discount, delta_lb, rho, lam;
    const float *root_hidden;    /* (B, N*hidden) network_output.hidden_state; NULL = already in pool slot 0 */
    const float *rewards;        /* (B,) network_output.reward */
    const float *values;         /* (B,) network_output.value */
    const float *logits;         /* (B,N,A) network_output.policy_logits */
    const float *legal;          /* (B,N,A) float 0/1, or NULL */
    const float *noise;          /* (B,Nt,A) np_random.dirichlet draws (:68) */
    const int *factor;           /* (B,N) actions of agents < cur (sequential mode), or NULL = zeros (:116-120) */
    maz_search_readout out;
} maz_search_call;

int maz_search_run_dev(maz_search *s, const maz_search_call *c);   /* device pointers; asynchronous */
int maz_search_run(maz_search *s, const maz_search_call *c);       /* host pointers; synchronous; one H2D + one D2H copy */
/* synchronise the stream and surface device-side invariant failures */
int maz_search_check(maz_search *s);

/* Parity instrumentation: when set (device pointers, all non-NULL), simulation i of the NEXT searches stores what the
 * reference's loop would pass between its steps: the selection outputs and the injected network outputs.  NULL rewards
 * switches it off.  rewards / values (S,B); probs / beta (S,B,Nt,A); idx_x (S,B); actions (S,B,Nt). */
int maz_search_set_record(maz_search *s, float *rewards, float *values, float *probs, float *beta, int *idx_x, int *actions);
/* profiling: SM-cycle counters of the persistent kernel, or NULL: [2 i + {0,1}] = CTA 0's inference / tree-step cycles of
 * simulation i; then for every CTA c, at [2 S + 4 c ...]: total cycles, total nanoseconds, inference cycles, tree cycles */
int maz_search_set_debug_clock(maz_search *s, long long *dev_clock);
/* measurement: when on, CUDA events bracket the SIMULATION LOOP of every search (the persistent kernel / the graph launch,
 * without root preparation, tree construction and readout); maz_search_loop_ms waits for the last one and returns its
 * duration in milliseconds. */
int maz_search_set_timing(maz_search *s, int on);
int maz_search_loop_ms(maz_search *s, float *ms);
/* roots per CTA of the persistent kernel (0 for the graph strategy) */
int maz_search_roots_per_cta(const maz_search *s);

#ifdef __cplusplus
}
#endif
#endif /* MAZ_SEARCH_H */
