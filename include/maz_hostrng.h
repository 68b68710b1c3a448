/*
 * maz_hostrng.h -- host-side helper of the search driver (libmaz_b200.so): the exploration-noise draw
 *   noises = np_random.dirichlet([alpha] * A, rows)          core/mcts/tree_search/mcts_sampled.py:68
 * reproduced BIT FOR BIT (same values, same final generator state) but evaluated in parallel.
 *
 * numpy's legacy RandomState draws each Gamma(alpha < 1) variate with a rejection loop whose every attempt
 * consumes exactly two doubles = four MT19937 words (numpy/random/src/legacy/legacy-distributions.c:
 * legacy_standard_gamma, shape < 1 branch; legacy_standard_exponential = -log(1 - U); mt19937_next_double =
 * (a>>5, b>>6) -> (a*2^26 + b) / 2^53).  Attempt j therefore depends only on words 4j..4j+3 of the stream, the
 * k-th variate is the k-th ACCEPTED attempt, and the attempts can be evaluated independently on all host cores;
 * only the (cheap) word generation and the in-order compaction are serial.  Rows are normalised exactly as
 * RandomState.dirichlet does (numpy/random/mtrand.pyx: acc += g sequentially, g * (1/acc)).
 * At 1024 roots x 3 agents x 9 actions this draw is the largest host cost of one search (1.1 ms of 6.9 ms).
 */
#ifndef MAZ_HOSTRNG_H
#define MAZ_HOSTRNG_H

#ifdef __cplusplus
extern "C" {
#endif

/* key[624], *pos: the MT19937 state of the RandomState (np_random.get_state()[1:3]); updated in place to the state
 * after the draw.  alpha in (0, 1).  out: rows*A float32 (= the float64 result cast, `.astype(np.float32)`), or
 * out64: rows*A float64; either may be NULL.  threads <= 0: all hardware threads (at most 16).
 * Returns MAZ_OK, or MAZ_ERR_UNSUPPORTED for alpha >= 1 (caller uses numpy itself). */
int maz_legacy_dirichlet(unsigned int *key, int *pos, double alpha, int rows, int A, float *out, double *out64,
                         int threads);

#ifdef __cplusplus
}
#endif
#endif /* MAZ_HOSTRNG_H */
