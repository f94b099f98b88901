/* sr2.h -- C ABI of libsr2, the B200 (sm_100a) reconciliation DP library.
 *
 * The reference (UdeM-LBIT/superrec2) is pure Python and has no FFI; this is
 * the new seam that its two compute entry points bind through ctypes
 * (see INTEGRATION.md for the reference-side stub):
 *
 *   sr2_dtl_*       replaces  reconcile_thl
 *                   /root/reference/src/superrec2/compute/reconciliation.py:266-294
 *                   (table fill :192-225, decode :228-263, selection :284-294)
 *   sr2_superdtl_*  replaces  usreconcile_extended_uspfs
 *                   /root/reference/src/superrec2/compute/unordered_super_reconciliation.py:523-542
 *                   (gain/LCA sets :85-136, table :147-358, decode :361-446,
 *                   selection over resolutions :449-497)
 *   sr2_resolutions_* replaces ReconciliationInput.binarize / utils.trees.binarize
 *                   /root/reference/src/superrec2/model/reconciliation.py:161-190,
 *                   /root/reference/src/superrec2/utils/trees.py:377-446
 *
 * Conventions
 *   - plain pointers and sizes only; all buffers are caller-owned HOST memory
 *     unless a function says "device";
 *   - every function returns 0 on success and a negative sr2_status on error,
 *     never throws; sr2_last_error() gives the message of the last failure on
 *     that context;
 *   - species nodes are indexed by LEVEL-ORDER (BFS) RANK of the species tree,
 *     which is the reference's candidate order (ete3 traverse() default) and
 *     therefore its tie-breaking order; the root is 0 and the two children of
 *     an internal node are consecutive ranks (child1 == child0 + 1);
 *   - object nodes of one instance are numbered 0..G-1 with the root at 0 and
 *     every parent before its children (e.g. preorder); left/right are -1 for
 *     leaves; instances of a batch are concatenated, node_off[B+1] gives the
 *     first node of each instance and child indices are instance-local;
 *   - costs[5] = {speciation, duplication, transfer, full loss, segmental loss},
 *     the order of get_default_cost()
 *     (/root/reference/src/superrec2/model/reconciliation.py:65-73);
 *   - SR2_INF marks "no entry" (the reference's infinity.inf).
 */
#ifndef SR2_H
#define SR2_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SR2_INF 0x3fffffff

typedef enum sr2_status {
    SR2_OK = 0,
    SR2_ERR_ARG = -1,      /* malformed input (message says what) */
    SR2_ERR_CUDA = -2,     /* a CUDA call failed (no CPU fallback exists) */
    SR2_ERR_RANGE = -3,    /* costs/sizes exceed what int32 tables can hold */
    SR2_ERR_NOMEM = -4,    /* device or host allocation failed */
    SR2_ERR_STATE = -5     /* call order (e.g. solve before upload) */
} sr2_status;

typedef struct sr2_ctx sr2_ctx;

/* flags for sr2_*_upload / sr2_*_run */
#define SR2_KEEP_TABLES 1u   /* keep every DP row so that sr2_*_tables can return them (tests) */

/* ---- context ------------------------------------------------------------ */
/* One context per (device, stream). Fails with SR2_ERR_CUDA when no sm_100
 * class GPU is usable: there is no CPU path behind this ABI. */
int sr2_create(int device, sr2_ctx** out);
void sr2_destroy(sr2_ctx* ctx);
const char* sr2_last_error(const sr2_ctx* ctx);
/* Launch on an existing CUDA stream (cudaStream_t passed as void*), e.g. the
 * current torch stream, so that caller-side CUDA events bracket the kernels. */
int sr2_set_stream(sr2_ctx* ctx, void* cuda_stream);
int sr2_synchronize(sr2_ctx* ctx);
/* sizeof-style introspection used by the tests */
int sr2_version(void);

/* ---- retention policy ------------------------------------------------------
 * RetentionPolicy of the reference (utils/dynamic_programming.py:89-101), the second argument of both
 * compute entry points (compute/reconciliation.py:266-269, unordered_super_reconciliation.py:523-526).
 *   SR2_POLICY_ANY (default): one solution per instance -- the one the reference returns under ANY
 *       (first candidate that reaches the minimum, level-order tie-breaking); only the rows the
 *       traceback needs stay on the device.
 *   SR2_POLICY_ALL: the values of the DP table do not depend on the policy (they are minima), the SET of
 *       optimal tags per cell does.  Under ALL every later upload keeps its whole table (as with
 *       SR2_KEEP_TABLES) so that sr2_*_tables can serve the enumeration of all optimal solutions, which
 *       the binding runs on the host (their number is exponential and each becomes a host object);
 *       sr2_*_results still return the ANY solution.  Applies to uploads made after the call. */
#define SR2_POLICY_ANY 0
#define SR2_POLICY_ALL 1
int sr2_set_policy(sr2_ctx* ctx, int policy);

/* ---- species tree --------------------------------------------------------
 * parent[0] = -1; child0/child1 = -1 for leaves.  Replaces the per-input
 * LowestCommonAncestor oracle (/root/reference/src/superrec2/utils/trees.py:31-142):
 * depth, ancestor tests and distances become table look-ups on the device. */
int sr2_set_species(sr2_ctx* ctx, int32_t n_species,
                    const int32_t* parent, const int32_t* child0, const int32_t* child1);

/* ---- dtl (reconcile_thl) -------------------------------------------------
 * upload : copy a batch to the device and bucket its internal nodes by height
 * solve  : wavefront DP + selection of the returned root + traceback, all on
 *          the device, inputs and outputs resident in HBM
 * results: copy min cost, chosen root species and the full mapping back
 * run    : upload + solve + results in one call (host buffers in and out)
 *
 * out_min_cost[b]    cost() of the returned solution (SR2_INF if none)
 * out_root_species[b] BFS rank of the species of the root (-1 if none)
 * out_map_species[node_off[b] + u]  species rank of node u (-1 if none)
 */
int sr2_dtl_upload(sr2_ctx* ctx, int32_t n_instances, const int64_t* node_off,
                   const int32_t* left, const int32_t* right, const int32_t* leaf_species,
                   const int32_t costs[5], uint32_t flags);
int sr2_dtl_solve(sr2_ctx* ctx);
int sr2_dtl_results(sr2_ctx* ctx, int32_t* out_min_cost, int32_t* out_root_species,
                    int32_t* out_map_species);
int sr2_dtl_run(sr2_ctx* ctx, int32_t n_instances, const int64_t* node_off,
                const int32_t* left, const int32_t* right, const int32_t* leaf_species,
                const int32_t costs[5], uint32_t flags,
                int32_t* out_min_cost, int32_t* out_root_species, int32_t* out_map_species);
/* sr2_dtl_run with 16-bit node and species numbers, for batches whose trees have at most 32,767 nodes
 * over at most 32,767 species nodes (the common case: BASELINE config #2 has 999 and 399): the three
 * input arrays and the two per-node / per-instance species results cross PCIe at half the size and are
 * widened / narrowed on the device.  Same results as sr2_dtl_run; SR2_ERR_RANGE when a tree or the
 * species tree is too large. */
int sr2_dtl_run16(sr2_ctx* ctx, int32_t n_instances, const int64_t* node_off,
                  const int16_t* left, const int16_t* right, const int16_t* leaf_species,
                  const int32_t costs[5], uint32_t flags,
                  int32_t* out_min_cost, int16_t* out_root_species, int16_t* out_map_species);
/* Test hook (needs SR2_KEEP_TABLES): full table of one instance.
 * out_value[u * n_species + x] = T[u][x] (leaf rows included), SR2_INF if absent;
 * out_tag[(u * n_species + x) * 2 + {0,1}] = species rank chosen for the
 * left/right child (MappingInfo, reconciliation.py:47-54), -1 if absent. */
int sr2_dtl_tables(sr2_ctx* ctx, int32_t instance, int32_t* out_value, int32_t* out_tag);

/* ---- superdtl (usreconcile_extended_uspfs) -----------------------------------
 * One instance = one BINARY object tree with leaf syntenies (a batch of independent
 * super-reconciliations, or the enumerated resolutions of one multifurcating tree).
 * Syntenies are bitsets of n_words 32-bit words per node: bit i = the i-th gene
 * family of the problem in sort_synteny order
 * (/root/reference/src/superrec2/model/synteny.py:21-33); leaf_bits[(node)*n_words + w],
 * rows of internal nodes are ignored.
 *
 * Per instance the device computes the gain and LCA sets (:85-136), fills the
 * (object, species, kind) table (:147-358), decodes the solution from every root
 * species in level order reproducing the reference's in-place synteny unions
 * (:361-446, SURVEY.md Hazard A), evaluates SuperReconciliationOutput.cost()
 * (model/reconciliation.py:511-554) on each decode and keeps the first strict
 * minimum (:474-495).
 *
 * index_base: global number of instance 0 (the resolution index when the batch is a
 * shard of an enumeration); sr2_superdtl_best returns the lexicographic minimum of
 * (cost, index_base + instance, root species) over the batch -- the reference's
 * running minimum over resolutions (:454) -- and its packed 64-bit key
 * (cost << 40 | instance << 16 | root) for a cross-GPU min-reduce.
 *
 * out_map_kind[node]: 0 = LCA synteny, 1 = INHERIT (SyntenyAssignment, :37-45).
 * out_synteny_bits[node*n_words + w]: the synteny the reference reports for the node.
 */
int sr2_superdtl_upload(sr2_ctx* ctx, int32_t n_instances, const int64_t* node_off,
                        const int32_t* left, const int32_t* right, const int32_t* leaf_species,
                        int32_t n_words, const uint32_t* leaf_bits, const int32_t costs[5],
                        int64_t index_base, uint32_t flags);
/* Optional, between upload and solve: restrict every object node to one species
 * (node_species[node], -1 = any species) -- the `allowed_species` argument of
 * _compute_uspfs_table (compute/unordered_super_reconciliation.py:300-358); with the LCA
 * mapping this is usreconcile_base_uspfs (:500-520).  NULL lifts the restriction. */
int sr2_superdtl_set_allowed(sr2_ctx* ctx, const int32_t* node_species);
int sr2_superdtl_solve(sr2_ctx* ctx);
int sr2_superdtl_results(sr2_ctx* ctx, int32_t* out_min_cost, int32_t* out_root_species,
                         int32_t* out_map_species, int32_t* out_map_kind, uint32_t* out_synteny_bits);
int sr2_superdtl_best(sr2_ctx* ctx, int32_t* out_min_cost, int64_t* out_instance,
                      int32_t* out_root_species, uint64_t* out_key);
int sr2_superdtl_result_one(sr2_ctx* ctx, int32_t instance, int32_t* out_map_species,
                            int32_t* out_map_kind, uint32_t* out_synteny_bits);
int sr2_superdtl_run(sr2_ctx* ctx, int32_t n_instances, const int64_t* node_off,
                     const int32_t* left, const int32_t* right, const int32_t* leaf_species,
                     int32_t n_words, const uint32_t* leaf_bits, const int32_t costs[5], uint32_t flags,
                     int32_t* out_min_cost, int32_t* out_root_species, int32_t* out_map_species,
                     int32_t* out_map_kind, uint32_t* out_synteny_bits);
/* Test hook (needs SR2_KEEP_TABLES).  out_value[(u*S + x)*2 + kind];
 * out_tag[((u*S + x)*2 + kind)*4 + {0..3}] = (left species, left kind, right species,
 * right kind) = ChildrenAssignment (:55-59), -1 if absent; out_gain/out_lca[u*n_words + w]. */
int sr2_superdtl_tables(sr2_ctx* ctx, int32_t instance, int32_t* out_value, int32_t* out_tag,
                        uint32_t* out_gain, uint32_t* out_lca);

/* ---- polytomy resolutions (ReconciliationInput.binarize) -------------------------
 * The multifurcating object tree is given in CSR form: nodes 0..n_nodes-1 with the root
 * at 0 and parents before children, children of v = child_idx[child_off[v] .. child_off[v+1])
 * in order (0 or 2..16 children).  Resolution numbers follow the reference's enumeration
 * order (model/reconciliation.py:161-190, utils/trees.py:377-446).
 *
 * sr2_resolutions_count : number of resolutions (capped at 2^40) and nodes of a resolved tree
 * sr2_resolutions_tree  : resolution `index` as flat left/right arrays in the library's
 *                         numbering; out_origin[id] = original node of leaves and of the roots
 *                         of resolved polytomies, -1 for the joints created by the resolution
 * sr2_resolutions_upload: unrank resolutions [first, first+count) and upload them as a
 *                         superdtl batch with index_base = first (then sr2_superdtl_solve,
 *                         sr2_superdtl_best, ...).  leaf_species / leaf_bits are indexed by the
 *                         nodes of the multifurcating tree (internal entries ignored). */
int sr2_resolutions_count(int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx,
                          int64_t* out_count, int32_t* out_binary_nodes);
int sr2_resolutions_tree(int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx,
                         int64_t index, int32_t* out_left, int32_t* out_right, int32_t* out_origin);
int sr2_resolutions_upload(sr2_ctx* ctx, int32_t n_nodes, const int32_t* child_off,
                           const int32_t* child_idx, const int32_t* leaf_species, int32_t n_words,
                           const uint32_t* leaf_bits, const int32_t costs[5], int64_t first,
                           int64_t count, uint32_t flags);

/* ---- timing of the last solve (CUDA events on the context's stream) ------ */
typedef struct sr2_timing {
    float upload_ms;     /* host prep + H2D of the last upload */
    float solve_ms;      /* whole solve */
    float waves_ms;      /* the wave kernels only (the dominant kernel) */
    float results_ms;    /* D2H of the last results call */
    int32_t n_waves;     /* wave kernel launches in the last solve */
    int32_t n_launches;  /* all kernel launches in the last solve */
    int64_t n_cells;     /* DP cells (or cell-kinds) evaluated by the last solve */
} sr2_timing;
int sr2_last_timing(const sr2_ctx* ctx, sr2_timing* out);

#ifdef __cplusplus
}
#endif
#endif /* SR2_H */
