// tree_layout.h -- HBM layout of one batch of search trees (shared by host and device code).
//
// One SLAB per tree, slabs back to back in one arena.  The fields of a node that selection, backup and expansion touch
// together (prior, reward, pred_value, weighted sum / total weight, visit count, children count / first child, hidden
// index, expansion order) form ONE 32-byte record per node slot (`NodeRec`): the children of a node occupy consecutive
// slots, as in the reference's pool (cnode.cpp:290-293), so a tree level is two 16-byte loads per child lane and three
// cache lines per ten children (the first layout, one array per field, cost nine scattered 4-byte loads per child and
// level, and the selection turned out to be bound by the NUMBER of load instructions).  The fields only the root readouts
// need (pred_prob, beta, beta_hat) and the joint actions stay separate arrays over the P = K*(S+2) slots.
//
// The reference's per-node SubTreeValueSet (utils.h:18-40: per-depth multisets big/small) becomes a
// per-TREE append-only value log: entry = (node slot, relative depth, in-big flag | value).  Only
// min(big) and max(small) of one (node, depth) set are ever queried (utils.cpp:34,59), which a warp
// gets by scanning the log with a shuffle reduction; moving a value between big and small flips one
// flag bit.  weighted_sum / tot_weight stay per node and are updated with the reference's exact fp32
// operation order.
//
// The reference's CMinMaxStats multiset (utils.h:42-53) holds exactly one entry per visited non-root
// node (cnode.cpp:431-446): here it is the array `qdelta`, indexed by the node's expansion order (`eid`, the
// e-th expanded node of the tree), reduced at the end of each backup; selection reads the cached (min,max).
#pragma once
#include <stdint.h>

namespace maz {

constexpr int kMaxSampledTimes = 32;   // children of a node are scored by one warp
constexpr int kMtN = 624;              // std::mt19937 state words
constexpr int kMtChunk = 128;          // raw draws staged per fetch (per warp, in shared memory)

struct TreeHdr {              // 64 bytes at the start of every slab
    int tot_nodes;            // CTree::tot_nodes
    int log_len;              // entries in the value log
    int path_len;             // SearchResult::search_len of the last selection
    int mt_pos;               // position in the mt19937 output block (624 = regenerate first)
    float mm_min, mm_max;     // CMinMaxStats: min / max of qdelta over visited non-root nodes
    int mm_cnt;               // 0 => empty set => normalize() is the identity (utils.cpp:97-98)
    int n_expanded;           // number of expanded nodes (root included) = next hidden_state_index_x
    int err;                  // first device-side invariant failure of this tree
    int sum_path_len;         // statistics: sum of path_len over selections
    int pad[6];
};
static_assert(sizeof(TreeHdr) == 64, "TreeHdr must be 64 bytes");

struct alignas(16) NodeRec {   // 32 bytes
    float prior;              // CNode::prior
    float reward;             // CNode::reward
    float pred_value;         // CNode::pred_value
    float wsum;               // SubTreeValueSet::weighted_sum
    float wtot;               // SubTreeValueSet::tot_weight
    int visit;                // CNode::visit_count
    uint16_t nchild, cbase;   // children count, first child slot
    int16_t hidx;             // hidden_state_index_x (-1: not expanded)
    uint16_t eid;             // expansion order = index of the node's q-delta entry
};
static_assert(sizeof(NodeRec) == 32, "NodeRec must be 32 bytes");

struct TreeLayout {
    int B, N, A, K, S;
    int P;        // node slots per tree = K*(S+2)   (cnode.cpp:562)
    int L;        // value-log capacity = 1 + S*(S+3)/2 + slack (worst case: one chain)
    float delta_lb, one_minus_rho, lam;
    unsigned long long slab_bytes;
    long long *dbg_clock;   // profiling only (NULL): SM-cycle timestamps of tree 0's phases, 64 entries
    const float *pbc_table; // pb_c[n][visit] = (float)((double)logterm[n] * (sqrt((double)n) / (double)(visit + 1))), host-built
    int pbc_dim;            // table is pbc_dim x pbc_dim (0: compute on the device with fp64)
    int wide_min_len;       // selection: use the all-expanded-nodes-in-parallel form when the previous path was at least this long
                            // (0 = always when possible, large = never); both forms give identical results
    // byte offsets inside a slab (all multiples of 128)
    unsigned off_mt, off_rec, off_pred_prob, off_beta, off_beta_hat, off_qdelta, off_actions, off_expslot, off_path, off_vskey,
        off_vsval, off_depth;
};

// device error codes stored in TreeHdr::err / the handle's global error word
enum : int {
    kErrNone = 0,
    kErrValueSetInvariant = 1,   // utils.cpp:50 "cur_size+1!=size_lim"
    kErrPoolExhausted = 2,
    kErrLogOverflow = 3,
    kErrPathOverflow = 4,
};

}  // namespace maz
