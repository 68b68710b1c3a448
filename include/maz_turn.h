/*
 * maz_turn.h -- C ABI of the steps either side of the search loop (libmaz_b200.so), the SURVEY 8(f) rows:
 * the driver's root preparation and the workers' per-agent turn logic, as device kernels, so that the N
 * sequential per-agent searches of one environment step run back to back on the GPU with the `factor`
 * hand-off on the device and ONE host synchronisation at the end.
 *
 * Device pointers only, asynchronous on the given stream (cudaStream_t), same error convention as
 * maz_tree.h (MAZ_OK / MAZ_ERR_*, maz_last_error()).  (B,N,A) arrays are row-major.
 */
#ifndef MAZ_TURN_H
#define MAZ_TURN_H

#ifdef __cplusplus
extern "C" {
#endif

/* Root preparation of one search: replaces the numpy block of SampledMCTS.batch_search
 * (core/mcts/tree_search/mcts_sampled.py:57-106):
 *   probs = softmax(logits[:, agent])                                     :64-65
 *   legal mask:  p *= m; p += m*1e-4; p /= sum(p); same for the noise      :73-83
 *   beta  = (p*(1-eps) + noise*eps) ** inv_tau, * m, / sum                 :93-100
 * and the greedy actions argmax_a logits of EVERY agent (what model.prediction(root) gives the
 * sequential-agent mode, :137-145).
 *   logits (B,n_agents,A); legal (B,n_agents,A) float 0/1 or NULL; noises_in (B,Nt,A) raw Dirichlet draws;
 *   cur = agent index of a sequential-agent search (Nt = 1) or -1 for the joint mode (Nt = n_agents);
 *   outputs probs / beta / noises_out (B,Nt,A) -- exactly the arrays Tree_batch.prepare takes --
 *   and greedy (B,n_agents) int32 (may be NULL). */
int maz_root_prepare_dev(const float *logits, const float *legal, const float *noises_in, int B, int n_agents,
                         int A, int cur, float noise_eps, float inv_tau, float *probs, float *beta, float *noises_out,
                         int *greedy, void *cuda_stream);

/* Dirichlet(alpha) exploration noise drawn ON THE DEVICE (opt-in; the reference draws it with numpy,
 * mcts_sampled.py:68): counter-based Philox4x32-10 keyed by (seed, row), Gamma(alpha) by Marsaglia-Tsang with the
 * alpha+1 boost, normalised per row.  out (rows, A).  Same distribution as the reference's, not the same stream. */
int maz_dirichlet_dev(float *out, int rows, int A, float alpha, unsigned long long seed, void *cuda_stream);

/* One agent's turn after its search.  Reads the padded readouts of that search (maz_tree_readout_dev with
 * agent_num = 1) and writes the agent's action into column `agent` of `actions` (B,n_agents) -- the `factor` of
 * the next agent's search.
 *   MAZ_TURN_GREEDY   reanalyze_worker.py:298-327: action = argmax_a(marginal_visits * legal) (first maximum);
 *   MAZ_TURN_SAMPLE   selfplay_worker.py:230-257: select_action (core/utils.py:289-319: counts**(1/T), normalised
 *                     in float64, np_random.choice -> cdf.searchsorted(u, 'right') with the INJECTED uniform u[b]),
 *                     then eps_greedy_action (core/utils.py:322-334) with INJECTED eps_u[b] and random_action[b];
 * both modes: policy_dist[b,agent,:] = marginal_visits / sum (float64; reanalyze_worker.py:319-320,
 * selfplay_worker.py:283-286), prob_prod[b] (float64) *= policy_dist[b,agent,action] (set to that for agent 0)
 * (reanalyze_worker.py:334-345), entropy[b,agent] = base-2 entropy of the visit distribution (SAMPLE only). */
enum { MAZ_TURN_GREEDY = 0, MAZ_TURN_SAMPLE = 1 };
int maz_agent_turn_dev(int mode, int B, int n_agents, int A, int K, int agent,
                       const int *num_children /* (B,) */, const int *sampled_actions /* (B,K,1) */,
                       const int *sampled_visit_count /* (B,K) */, const int *marginal_visit_count /* (B,1,A) */,
                       const float *legal /* (B,n_agents,A) or NULL */, double inv_temperature,
                       const double *uniforms /* (B,) SAMPLE */, float greedy_epsilon, const float *eps_u /* (B,) or NULL */,
                       const int *random_action /* (B,) or NULL */, int *actions /* (B,n_agents) */,
                       double *policy_dist /* (B,n_agents,A) */, double *prob_prod /* (B,) */,
                       double *entropy /* (B,n_agents) or NULL */, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MAZ_TURN_H */
