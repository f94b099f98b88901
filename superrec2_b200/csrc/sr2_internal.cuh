// sr2_internal.cuh -- shared declarations of libsr2 (not part of the C ABI).
//
// Data layout in HBM:
//   species tree  : int32 arrays indexed by BFS rank (parent, child0, depth,
//                   preorder entry/exit times) + the list of internal nodes in
//                   BFS order with per-level offsets; the i-th internal node of
//                   that list has children 2i+1 and 2i+2.
//   DP tables     : one row of `pitch` int32 per INTERNAL object node ("row"),
//                   rows ordered by wave (= height of the node), so one wave
//                   reads rows written by earlier launches and writes a
//                   contiguous block.  Leaf rows are never stored: a leaf row is
//                   0 at its species and INF elsewhere and is synthesised on load.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/sr2.h"

#define SR2_NO_TAG 0xffffffffu

// Device-side invariant checks, compiled in with -DSR2_DEBUG_CHECKS (scripts/debug_checks.sh builds
// libsr2_debug.so and runs the GPU tests against it).  compute-sanitizer is closed on this pool, so
// index bounds and ordering invariants of the kernels are asserted by hand instead; a violated check
// fails the launch loudly (cudaErrorAssert -> SR2_ERR_CUDA).
#ifdef SR2_DEBUG_CHECKS
#include <assert.h>
#define SR2_DCHECK(cond) assert(cond)
#else
#define SR2_DCHECK(cond) ((void)0)
#endif

typedef unsigned long long u64;

// Device view of the species tree (all pointers are device memory).
struct DevSpecies {
    int S;           // number of species nodes
    int pitch;       // row pitch of the DP tables (int32 elements, multiple of 32)
    int D;           // number of levels (max depth + 1)
    int n_internal;  // number of internal species nodes
    const int32_t* child0;      // [S]  rank of the first child, -1 for leaves
    const int32_t* depth;       // [S]
    const int32_t* tin;         // [S]  preorder entry time
    const int32_t* tout;        // [S]  preorder exit time (exclusive)
    const int32_t* parent;      // [S]
    const int32_t* int_list;    // [n_internal] BFS ranks of the internal nodes, BFS order
    const int32_t* int_lvl_off; // [D+1] offsets into int_list per depth
    // packed per-species words for the 32-bit-key kernels (rank_bits = ceil log2 S,
    // depth_bits = ceil log2 D):
    const uint32_t* keylow;     // [S]  rank << depth_bits | depth  (low part of a 32-bit key)
    const uint32_t* meta;       // [S]  (child0 + 1) << 16 | depth  (0 in the high half for leaves)
    int rank_bits, depth_bits;
    // sweep schedule for the warp-per-row kernels: n_steps rows of 32 words, step k word l =
    // (first child rank c) << 16 | (rank p) of the internal species handled by lane l, or
    // 0xffffffff when the lane idles; steps are in top-down level order and one level never
    // shares a step with another (levels wider than 32 take several steps)
    const uint32_t* sweep_sched;
    int n_steps;
    // dtl_wave_flat2_kernel: the same schedule as byte offsets (child pair << 16 | parent) into a
    // row's shared-memory array of flat_rs 8-byte elements (element y at index y + 1; behind the
    // pitch: index pitch+2,+3 = an always-infinite pair, +4,+5 / +6 = dummies for idle lanes);
    // flat_rs == 0: offsets do not fit 16 bits.  keylow_p / cmeta: [pitch] padded per-species words,
    // cmeta = (byte offset of the child pair, or of the infinite pair for leaves) << 16 | depth
    const uint32_t* flat_sched;
    int flat_rs;
    const uint32_t* keylow_p;
    const uint32_t* cmeta;
    // [S] x 4 words: tin, tout, (child0 + 1) << 16 | depth, tin of the second child (or tout):
    // everything the closed-form cherry cells need about a species in one 16-byte load
    const uint4* span4;
    // anc_tab[s * anc_pitch + d] = ancestor of species s at depth d (d <= depth[s]); null when the
    // table would be too large (S x D > 2^23 entries)
    const uint16_t* anc_tab;
    int anc_pitch;
    const uint8_t* lca_depth;  // [S * S] depth of lca(a, b); null for large species trees
    // pathmask[s * mask_stride + w]: bit set of the species on the root path of s (s included), one
    // bit per BFS rank, mask_words = pitch / 32.  The finite cells of a DP row are a subset of the
    // union of the root paths of the leaf species below its object node.
    const uint32_t* pathmask;
    int mask_words;
    int mask_stride;  // words between two bit sets: mask_words rounded up to a multiple of 4 (uint4 access)
};

struct DtlProblem;
struct SuperProblem;

// Plan of a range of polytomy resolutions whose DP rows are shared (sr2_resolve.cu builds it,
// sr2_superdtl.cu consumes it).  Two resolutions contain the same binary subtree below a joint
// whenever they agree on the digits of the original nodes under it, so the table has one row
// ("slot") per (original node v, variant of v's subtree, joint of v's arrangement) instead of one
// per (resolution, internal node).  A variant of v is numbered var = (r / stride[v]) % count[v].
struct SharedPlan {
    // the multifurcating tree (original nodes), as the device-side unranking needs it
    int n = 0;                                  // original nodes
    int Gb = 0;                                 // nodes of one resolved (binary) tree
    int64_t first = 0, count = 0;               // resolution numbers [first, first + count)
    std::vector<int32_t> arity, base;           // [n] children / first id of the node's block in a resolved tree
    std::vector<int64_t> ways, stride, cnt;     // [n] own arrangements, digit stride, variants of the subtree
    std::vector<int64_t> var_first, raw_base;   // [n] first variant in the plan, first raw slot of the node
    std::vector<int32_t> child_of;              // [n_edges] children, concatenated (child_off below)
    std::vector<int32_t> child_off;             // [n+1]
    std::vector<int32_t> arr_off;               // [n] offset of the node's arrangement table in arr
    // arr[arr_off[v] + (j * (k-1) + q) * 2 + side]: child of joint q (preorder rank inside the block) in
    // arrangement j: >= 0 item (child number), < 0 joint -1-q'
    std::vector<int32_t> arr;
    // arr_pre[arr_pre_off[v] + j * (2k - 1) + i]: preorder position, relative to the root of v's block, of joint i
    // (i < k - 1) or of the root of item i - (k - 1)'s subtree in arrangement j.  The subtree of an original
    // node has the same size in every resolution, so these only depend on (v, j); with parent[] / child_pos[]
    // a node finds its absolute preorder position by walking up the ORIGINAL tree, without any pass over
    // the resolved tree.
    std::vector<int32_t> arr_pre, arr_pre_off, parent, child_pos;
    // slots, ordered by height: wave h = [wave_off[h], wave_off[h+1]), h = 1..H
    int64_t n_slots = 0;
    std::vector<int64_t> wave_off;
    std::vector<int32_t> perm;                  // [n_slots] raw slot -> slot
    std::vector<int32_t> slot_cl, slot_cr;      // [n_slots] table descriptors: child slot, or -1 - leaf species
    std::vector<int32_t> dag_left, dag_right;   // [Gb + n_slots] children as dag ids (leaf: its id in a resolved
                                                // tree; slot s: Gb + s), -1 for leaves
    std::vector<int32_t> leaf_species;          // [Gb] species of the leaves of a resolved tree, -1 elsewhere
    std::vector<uint32_t> leaf_bits;            // [Gb * W]
    // scratch of the planner (raw slots before the sort by height); kept with the plan so that a caller
    // that plans range after range reuses the memory instead of faulting in fresh pages every time
    std::vector<int32_t> raw_l, raw_r;
    std::vector<uint16_t> height;
};
int sr2_superdtl_upload_shared(sr2_ctx* ctx, const SharedPlan& plan, int32_t n_words, const int32_t costs[5],
                               uint32_t flags);

// Grow-only buffer pools owned by the context: repeated uploads reuse device buffers and
// pinned host staging instead of paying cudaMalloc/cudaFree and first-touch page faults.
struct Sr2Pool {
    void* ptr = nullptr;
    size_t cap = 0;
};
enum Sr2DevPool {
    DP_ROWS_SLAB, DP_ROOT_ROW, DP_ROOT_NODE, DP_NODE_PRE, DP_NODE_INST,
    DP_DTL_T, DP_DTL_TAG, DP_DTL_C, DP_DTL_MINCOST, DP_DTL_ROOTSP, DP_DTL_MAP, DP_DTL_MASK, DP_DTL_LISTS, DP_DTL_COUNTS, DP_DTL_CHAIN, DP_DTL_ROWCNT, DP_DTL_ROWINFO, DP_DTL_CLASSES, DP_DTL_DESC,
    DP_S_LEFT, DP_S_RIGHT, DP_S_NODEOFF, DP_S_BITS, DP_S_PRESENT, DP_S_OUTSIDE, DP_S_GAIN, DP_S_L, DP_S_FLAGS,
    DP_S_T, DP_S_TAG, DP_S_DEC, DP_S_ANC, DP_S_TAU, DP_S_TAUMIN, DP_S_COST, DP_S_MINCOST, DP_S_ROOTSP, DP_S_MAP,
    DP_S_KIND, DP_S_SYN, DP_S_BEST, DP_S_ALLOWED, DP_S_EV, DP_S_CHAIN,
    DP_S_NODESLOT, DP_S_NODEDAG, DP_S_NODEPRE, DP_S_NODEINST, DP_S_PLAN32, DP_S_PLAN64,
    DP_RAW, DP_RAW_HEIGHT, DP_RAW_NODEROW, DP_RAW_INSTBASE, DP_RAW_COUNTS, DP_RAW_NODEOFF,
    DP_SPECIES_SLAB, DP_SPECIES_ANC, DP_RAW16, DP_DTL_MAP16, DP_COUNT
};
enum Sr2PinPool { HP_ROWS_SLAB, HP_ROOT_ROW, HP_ROOT_NODE, HP_NODE_PRE, HP_NODE_INST, HP_RESULTS, HP_RAW, HP_COUNTS, HP_RESULTS16, HP_COUNT };
enum Sr2HostPool { MP_HEIGHT, MP_NODE_ROW, MP_CNT, MP_COUNT };

struct sr2_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;    // device-to-host result copies of pipelined runs (lazy)
    cudaStream_t stream2 = nullptr;        // second compute stream of pipelined runs (lazy)
    cudaStream_t fill_stream = nullptr;    // device-side bucketing, second phase (lazy)
    cudaStream_t part_stream[2] = {nullptr, nullptr};  // path sets / row lists of the waves, ahead of the waves
    cudaStream_t dense_stream[2] = {nullptr, nullptr}; // the dense rows of a wave, next to its sparse rows
    cudaStream_t tier_streams[2][8] = {};              // the sparse tiers of a wave after the first (SR2_DTL_SPLIT_TIERS)
    std::vector<cudaEvent_t> split_events[2];
    size_t split_used[2] = {0, 0};
    std::vector<cudaEvent_t> part_events[2];           // "lists of wave k are ready", per table set
    cudaStream_t upload_stream = nullptr;  // host-to-device row copies of pipelined runs (lazy)
    std::vector<cudaEvent_t> chunk_events; // "rows of chunk i are on the device"
    // per-chunk events of the dtl solves, pooled for the life of the context (creating and
    // destroying events per call showed up as multi-millisecond stalls): waves begin / end,
    // chunk finished (timing enabled), results copied back (no timing)
    std::vector<cudaEvent_t> dtl_ev_w0, dtl_ev_w1, dtl_ev_done, dtl_ev_copied;
    // free device memory as last reported by the driver; refreshed only after a pool grew
    size_t mem_free_cached = 0;
    bool mem_dirty = true;
    int policy = SR2_POLICY_ANY;   // sr2_set_policy: ALL keeps the whole table of later uploads
    bool species_valid = false;   // the device copy of the species tree matches h_parent / h_child0
    std::string err;
    int sm_count = 148;
    int host_threads = 1;  // OpenMP threads of the host-side bucketing (see sr2_create)
    size_t smem_optin = 0;

    // species tree (host copies + one device slab)
    int S = 0, pitch = 0, D = 0, n_internal = 0;
    std::vector<int32_t> h_parent, h_child0, h_depth, h_tin, h_tout, h_int_list, h_int_lvl_off;
    int32_t* d_species_slab = nullptr;
    uint16_t* d_anc_tab = nullptr;
    DevSpecies dsp{};

    Sr2Pool dev[DP_COUNT], pin[HP_COUNT], host[MP_COUNT];

    DtlProblem* dtl = nullptr;
    SuperProblem* super = nullptr;
    sr2_timing timing{};
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

int sr2_fail(sr2_ctx* ctx, int code, const std::string& msg);
int sr2_cuda_fail(sr2_ctx* ctx, cudaError_t e, const char* what);
int sr2_dev_reserve(sr2_ctx* ctx, int id, size_t bytes, void** out);   // device memory
int sr2_pin_reserve(sr2_ctx* ctx, int id, size_t bytes, void** out);   // pinned host memory
int sr2_host_reserve(sr2_ctx* ctx, int id, size_t bytes, void** out);  // plain host scratch
size_t sr2_dev_pooled(const sr2_ctx* ctx, int id);
int sr2_mem_free(sr2_ctx* ctx, size_t* out_free);  // cudaMemGetInfo, cached until a pool grows
void sr2_free_dtl(sr2_ctx* ctx);
void sr2_free_super(sr2_ctx* ctx);

#define SR2_CUDA(ctx, call)                                        \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return sr2_cuda_fail(ctx, e__, #call); \
    } while (0)

// ---------------------------------------------------------------------------
// 64-bit arg-min keys.  Lexicographic order (value, BFS rank[, kind]) is the
// reference's "first candidate that reaches the minimum wins"
// (utils/dynamic_programming.py:228-238): candidates are offered in BFS order.
// The low 16 bits carry the depth of the species so that the loss term
// loss * distance needs no second look-up; they never decide a comparison
// because the rank above them is unique.
//   dtl      : value[63:32] | rank[31:16]          | depth[15:0]
//   superdtl : value[63:32] | rank[31:16] | kind[15] | depth[14:0]
// ---------------------------------------------------------------------------
#define SR2_KEY_NONE 0xffffffffffffffffull

__device__ __forceinline__ u64 key_make(int value, int rank, int depth) {
    return ((u64)(unsigned)value << 32) | ((u64)(unsigned)rank << 16) | (unsigned)depth;
}
__device__ __forceinline__ int key_value(u64 k) { return (int)(k >> 32); }
__device__ __forceinline__ int key_rank(u64 k) { return (int)((k >> 16) & 0xffffu); }
__device__ __forceinline__ int key_depth(u64 k) { return (int)(k & 0xffffu); }
__device__ __forceinline__ u64 key_min(u64 a, u64 b) { return a < b ? a : b; }
