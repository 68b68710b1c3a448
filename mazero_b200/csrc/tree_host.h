// tree_host.h -- the host-side handle of one batch of trees (private to libmaz_b200.so: shared by maz_tree.cu, which owns it,
// and maz_search.cu, which launches the persistent whole-search kernel on its arena).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "tree_layout.h"

struct maz_tree {
    using TreeLayout = maz::TreeLayout;
    TreeLayout L{};
    int device = 0;
    cudaStream_t stream = nullptr;
    char *arena = nullptr;
    size_t arena_bytes = 0;
    float *d_lam_pow = nullptr;   // lam_pow[d], d = 0..S+1 (utils.cpp:25-26 running fp32 product)
    float *d_logterm = nullptr;   // (float)(log((n + c_base + 1)/c_base) + c_init), n = 0..S+1
    double *d_sqrtn = nullptr;    // sqrt((double)n)
    float *d_pbc = nullptr;       // pb_c[n][visit] (cnode.cpp:313-314 evaluated on the host for every (n, visit))
    int table_len = 0;
    bool puct_set = false;
    float c_base = 0, c_init = 0;
    int *d_err = nullptr;
    unsigned long long *d_sums = nullptr;
    unsigned int seed = 0, root_offset = 0;
    int wpb = 1;                  // warps (= trees) per block
    size_t scratch_per_warp = 0;
    // staging for the host-pointer entry points (allocated on first use)
    float *s_rewards = nullptr, *s_values = nullptr, *s_probs = nullptr, *s_beta = nullptr, *s_noises = nullptr;
    int *s_idx = nullptr;         // idx_x | idx_y | act   (B*(2+N))
    char *s_readout = nullptr;    // device mirror of all readout arrays
    bool prepared = false;
};


namespace maz {
int set_last_error(int code, const std::string &msg);
const char *device_error_message(int code);
}
