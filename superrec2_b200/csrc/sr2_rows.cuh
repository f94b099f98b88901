// sr2_rows.cuh -- wave layout of a batch of object trees (shared by dtl and superdtl).
//
// Every INTERNAL object node becomes a "row" (one DP row of |S| cells).  Rows are
// ordered by (chunk, height, instance, node) so that one wave = one height = one
// contiguous block of rows, written by one launch and read by later launches.
// A chunk is a range of instances whose rows fit the device-memory budget; chunks
// are solved one after the other in the same table buffers.
#pragma once
#include "sr2_internal.cuh"

struct RowChunk {
    int32_t inst_begin = 0, inst_end = 0;
    int64_t row_begin = 0, row_end = 0;  // global row range of the chunk
    std::vector<int64_t> wave_off;       // wave h (1-based height) = rows [wave_off[h], wave_off[h+1])
    int height() const { return (int)wave_off.size() - 2; }
};

struct RowLayout {
    int32_t B = 0;
    int64_t N = 0, R = 0;
    std::vector<int64_t> node_off;  // [B+1]
    std::vector<RowChunk> chunks;
    int64_t max_chunk_rows = 0;
    int max_height = 0;
    int64_t max_nodes = 0;          // largest instance
    // device (pooled in the context, not owned): per row child descriptors and node ids
    int32_t* d_slab = nullptr;
    int32_t *row_cl = nullptr, *row_cr = nullptr;      // >=0: child's row inside the chunk; <0: -1-leaf species
    int32_t *row_node = nullptr, *row_left = nullptr, *row_right = nullptr;  // global node ids
    int64_t* d_root_row = nullptr;   // [B] global row of the root, or -1-leaf_species for a one-node tree
    int32_t* d_root_node = nullptr;  // [B] global node id of the root
    // optional
    int32_t* d_node_pre = nullptr;   // [N] preorder position of each node inside its instance
    int32_t* d_node_inst = nullptr;  // [N] instance of each node
    std::vector<int32_t> h_node_row; // [N] global row of a node, -1 for leaves (kept on request)
    std::vector<int32_t> h_leaf;     // [N] leaf species (kept on request)

    void release();  // forgets the pooled pointers (the context keeps the memory)
};

// Optional callbacks of sr2_build_rows, so that a caller can overlap the host-side bucketing of
// the next chunk with device work on the previous one: on_plan runs once chunks, sizes
// (R, max_chunk_rows, max_nodes) and the device row arrays are known; on_chunk(i) runs after
// chunk i's rows were bucketed and their host-to-device copies were enqueued on ctx->stream.
struct RowHooks {
    void* user = nullptr;
    int (*on_plan)(void* user) = nullptr;
    int (*on_chunk)(void* user, int chunk) = nullptr;
    // sr2_build_rows_device only: the caller's raw arrays as 16-bit integers (sr2_dtl_run16); they
    // cross PCIe at half the size and are widened on the device.  Null: the 32-bit arguments are used.
    const int16_t *left16 = nullptr, *right16 = nullptr, *leaf16 = nullptr;
};

// Splits the batch into chunks of at most `rows_budget` rows, then chunk by chunk validates the
// trees (ABI numbering rules), computes heights, fills the row arrays and copies them to the
// device (asynchronously; the call returns after a stream synchronize).
int sr2_build_rows(sr2_ctx* ctx, const char* who, int32_t B, const int64_t* node_off,
                   const int32_t* left, const int32_t* right, const int32_t* leaf_species,
                   int64_t rows_budget, bool keep_host, bool want_preorder, RowLayout* out,
                   const RowHooks* hooks = nullptr, int64_t first_chunk_rows = 0);

// The same result as sr2_build_rows, with the bucketing done ON THE DEVICE (dtl pipelined runs):
// the raw child / leaf arrays of a chunk are copied to the device (directly when the caller's arrays
// are pinned, else through pinned staging), one thread per tree validates it, computes heights and
// per-height counts (rows_height_kernel), the host turns the chunk's counts into wave offsets, and a
// second kernel hands out rows bottom-up (rows_fill_kernel).  The host only plans chunks and waits
// for a few hundred bytes per chunk.  Returns SR2_ROWS_FALLBACK (> 0, nothing launched for the
// offending chunk) when a tree is higher than the per-thread tables allow; the caller then uses
// sr2_build_rows.  hooks->on_chunk is required.
#define SR2_ROWS_FALLBACK 1
int sr2_build_rows_device(sr2_ctx* ctx, const char* who, int32_t B, const int64_t* node_off,
                          const int32_t* left, const int32_t* right, const int32_t* leaf_species,
                          int64_t rows_budget, RowLayout* out, const RowHooks* hooks, int64_t first_chunk_rows);
