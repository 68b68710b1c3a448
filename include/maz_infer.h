/*
 * maz_infer.h -- C ABI of the fused network forward used inside the search (libmaz_b200.so).
 *
 * Replaces, for the search only, the reference's per-simulation model calls
 *   model.recurrent_inference(hidden, action)   config/smac/model.py:562-574
 *   model.prediction(hidden)                    config/smac/model.py:494-501
 * (called from core/mcts/tree_search/mcts_sampled.py:137,151) with one sm_100a kernel whose GEMMs run on
 * the tcgen05 tensor cores in bf16 with fp32 accumulation in TMEM.  Device pointers only; asynchronous on
 * the given stream.  Same error convention as maz_tree.h (maz_last_error()).
 */
#ifndef MAZ_INFER_H
#define MAZ_INFER_H

#ifdef __cplusplus
extern "C" {
#endif

/* Self-test of the tensor-core GEMM stage: out[128 x n] (fp32) = bf16(a[128 x k]) * w_packed[n x k]^T.
 * `w_packed` is bf16 in the kernels' shared-memory operand layout (see mazero_b200/csrc/umma.cuh):
 * W.view(n/8, 8, k/8, 8).permute(0, 2, 1, 3).  n, k multiples of 16, n <= 256, k <= 512. */
int maz_dbg_umma_gemm(const float *a, const void *w_packed, float *out, int n, int k, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MAZ_INFER_H */
