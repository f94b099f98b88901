// sr2_rows.cu -- host-side bucketing of a batch into waves (see sr2_rows.cuh).
// Replaces the postorder loop of the reference's table drivers
// (/root/reference/src/superrec2/compute/reconciliation.py:202-223,
//  compute/unordered_super_reconciliation.py:329-356): nodes of equal height are
// independent, so they are grouped into one launch.
#include <algorithm>
#include <cstring>
#include <omp.h>

#include "sr2_rows.cuh"

void RowLayout::release() {
    d_slab = nullptr;
    row_cl = row_cr = row_node = row_left = row_right = nullptr;
    d_root_row = nullptr;
    d_root_node = nullptr;
    d_node_pre = nullptr;
    d_node_inst = nullptr;
}

int sr2_build_rows(sr2_ctx* ctx, const char* who, int32_t B, const int64_t* node_off,
                   const int32_t* left, const int32_t* right, const int32_t* leaf_species,
                   int64_t rows_budget, bool keep_host, bool want_preorder, RowLayout* out) {
    const std::string tag = std::string(who) + ": ";
    const int S = ctx->S;
    const int64_t N = B ? node_off[B] : 0;
    if (N >= (int64_t)1 << 31)
        return sr2_fail(ctx, SR2_ERR_RANGE, tag + "more than 2^31 object nodes in one batch");
    out->B = B;
    out->N = N;
    out->node_off.assign(node_off, node_off + B + 1);

    // ---- pass 1: validate, heights ------------------------------------------------
    uint16_t* height = nullptr;
    {
        void* ptr = nullptr;
        int rc = sr2_host_reserve(ctx, MP_HEIGHT, (size_t)std::max<int64_t>(N, 1) * sizeof(uint16_t), &ptr);
        if (rc) return rc;
        height = (uint16_t*)ptr;
    }
    int bad_inst = -1;
    std::string bad_msg;
    int Hmax = 0;
    int64_t maxG = 0;
#pragma omp parallel for num_threads(ctx->host_threads) schedule(dynamic, 64) reduction(max : Hmax) reduction(max : maxG)
    for (int b = 0; b < B; ++b) {
        int64_t off = node_off[b];
        int64_t G = node_off[b + 1] - off;
        const char* why = nullptr;
        if (G < 1) why = "empty object tree";
        std::vector<uint8_t> refs;
        if (!why) refs.assign(G, 0);
        for (int64_t u = G - 1; u >= 0 && !why; --u) {
            int l = left[off + u], r = right[off + u];
            if (l < 0 && r < 0) {
                int s = leaf_species[off + u];
                if (s < 0 || s >= S) why = "leaf species out of range";
                height[off + u] = 0;
            } else if (l <= u || r <= u || l >= G || r >= G || l == r) {
                why = "children must be two distinct nodes numbered after their parent";
            } else {
                if (++refs[l] > 1 || ++refs[r] > 1) why = "a node has two parents";
                int h = 1 + std::max(height[off + l], height[off + r]);
                if (h > 65535) why = "object tree higher than 65535";
                height[off + u] = (uint16_t)h;
            }
        }
        if (!why)
            for (int64_t u = 1; u < G; ++u)
                if (refs[u] != 1) {
                    why = "a node other than node 0 has no parent";
                    break;
                }
        if (why) {
#pragma omp critical
            if (bad_inst < 0 || b < bad_inst) {
                bad_inst = b;
                bad_msg = why;
            }
            continue;
        }
        Hmax = std::max(Hmax, (int)height[off]);
        maxG = std::max(maxG, G);
    }
    if (bad_inst >= 0)
        return sr2_fail(ctx, SR2_ERR_ARG, tag + "instance " + std::to_string(bad_inst) + ": " + bad_msg);
    out->max_height = Hmax;
    out->max_nodes = maxG;

    // ---- chunks ------------------------------------------------------------------------
    out->chunks.clear();
    {
        RowChunk c;
        int64_t rows = 0, total_rows = 0;
        for (int b = 0; b < B; ++b) {
            int64_t mine = (node_off[b + 1] - node_off[b] - 1) / 2;
            if (mine > rows_budget)
                return sr2_fail(ctx, SR2_ERR_NOMEM, tag + "one instance alone exceeds device memory");
            if (rows + mine > rows_budget) {
                c.inst_end = b;
                c.row_end = total_rows;
                out->chunks.push_back(c);
                c = RowChunk();
                c.inst_begin = b;
                c.row_begin = total_rows;
                rows = 0;
            }
            rows += mine;
            total_rows += mine;
        }
        c.inst_end = B;
        c.row_end = total_rows;
        out->chunks.push_back(c);
        out->R = total_rows;
    }
    const int64_t R = out->R;

    // ---- pass 2: rows in (chunk, height, instance, node) order ---------------------------
    // staging in pinned host memory (reused across uploads): full-speed H2D, no page faults
    const size_t slab_elems = (size_t)std::max<int64_t>(R, 1) * 5;
    int32_t* h_cl = nullptr;
    int64_t* h_root_row = nullptr;
    int32_t* h_root_node = nullptr;
    int32_t* node_row = nullptr;
    {
        void* ptr = nullptr;
        int rc = sr2_pin_reserve(ctx, HP_ROWS_SLAB, slab_elems * sizeof(int32_t), &ptr);
        if (rc) return rc;
        h_cl = (int32_t*)ptr;
        rc = sr2_pin_reserve(ctx, HP_ROOT_ROW, (size_t)std::max(B, 1) * sizeof(int64_t), &ptr);
        if (rc) return rc;
        h_root_row = (int64_t*)ptr;
        rc = sr2_pin_reserve(ctx, HP_ROOT_NODE, (size_t)std::max(B, 1) * sizeof(int32_t), &ptr);
        if (rc) return rc;
        h_root_node = (int32_t*)ptr;
        rc = sr2_host_reserve(ctx, MP_NODE_ROW, (size_t)std::max<int64_t>(N, 1) * sizeof(int32_t), &ptr);
        if (rc) return rc;
        node_row = (int32_t*)ptr;
    }
    int32_t* h_cr = h_cl + R;
    int32_t* h_node = h_cr + R;
    int32_t* h_left = h_node + R;
    int32_t* h_right = h_left + R;
    const int W = Hmax + 1;
    out->max_chunk_rows = 0;
    for (auto& c : out->chunks) {
        int nb = c.inst_end - c.inst_begin;
        int32_t* cnt = nullptr;
        {
            void* ptr = nullptr;
            int rc = sr2_host_reserve(ctx, MP_CNT, (size_t)std::max(nb, 1) * W * sizeof(int32_t), &ptr);
            if (rc) return rc;
            cnt = (int32_t*)ptr;
        }
#pragma omp parallel for num_threads(ctx->host_threads) schedule(static)
        for (int b = c.inst_begin; b < c.inst_end; ++b) {
            int64_t off = node_off[b], G = node_off[b + 1] - off;
            int32_t* mine = cnt + (size_t)(b - c.inst_begin) * W;
            for (int h = 0; h < W; ++h) mine[h] = 0;
            for (int64_t u = 0; u < G; ++u) mine[height[off + u]]++;
        }
        c.wave_off.assign(W + 1, 0);
        int64_t run = c.row_begin;
        c.wave_off[0] = run;
        for (int h = 1; h < W; ++h) {
            c.wave_off[h] = run;
            for (int i = 0; i < nb; ++i) {
                int32_t k = cnt[(size_t)i * W + h];
                cnt[(size_t)i * W + h] = (int32_t)run;
                run += k;
            }
        }
        c.wave_off[W] = run;
        out->max_chunk_rows = std::max(out->max_chunk_rows, c.row_end - c.row_begin);
#pragma omp parallel for num_threads(ctx->host_threads) schedule(dynamic, 64)
        for (int b = c.inst_begin; b < c.inst_end; ++b) {
            int64_t off = node_off[b], G = node_off[b + 1] - off;
            int32_t* cursor = cnt + (size_t)(b - c.inst_begin) * W;
            for (int64_t u = 0; u < G; ++u) {
                int h = height[off + u];
                node_row[off + u] = h > 0 ? cursor[h]++ : -1;
            }
            for (int64_t u = 0; u < G; ++u) {
                int64_t gr = node_row[off + u];
                if (gr < 0) continue;
                int l = left[off + u], r = right[off + u];
                int64_t rl = node_row[off + l], rr = node_row[off + r];
                h_cl[gr] = rl >= 0 ? (int32_t)(rl - c.row_begin) : -1 - leaf_species[off + l];
                h_cr[gr] = rr >= 0 ? (int32_t)(rr - c.row_begin) : -1 - leaf_species[off + r];
                h_node[gr] = (int32_t)(off + u);
                h_left[gr] = (int32_t)(off + l);
                h_right[gr] = (int32_t)(off + r);
            }
            h_root_node[b] = (int32_t)off;
            h_root_row[b] = node_row[off] >= 0 ? node_row[off] : -1 - (int64_t)leaf_species[off];
        }
    }

    // ---- device copies (pooled buffers) ---------------------------------------------------
    out->release();
    {
        void* ptr = nullptr;
        int rc = sr2_dev_reserve(ctx, DP_ROWS_SLAB, slab_elems * sizeof(int32_t), &ptr);
        if (rc) return rc;
        out->d_slab = (int32_t*)ptr;
        rc = sr2_dev_reserve(ctx, DP_ROOT_ROW, (size_t)std::max(B, 1) * sizeof(int64_t), &ptr);
        if (rc) return rc;
        out->d_root_row = (int64_t*)ptr;
        rc = sr2_dev_reserve(ctx, DP_ROOT_NODE, (size_t)std::max(B, 1) * sizeof(int32_t), &ptr);
        if (rc) return rc;
        out->d_root_node = (int32_t*)ptr;
    }
    out->row_cl = out->d_slab;
    out->row_cr = out->row_cl + R;
    out->row_node = out->row_cr + R;
    out->row_left = out->row_node + R;
    out->row_right = out->row_left + R;
    SR2_CUDA(ctx, cudaMemcpyAsync(out->d_slab, h_cl, slab_elems * sizeof(int32_t), cudaMemcpyHostToDevice,
                                  ctx->stream));
    SR2_CUDA(ctx, cudaMemcpyAsync(out->d_root_row, h_root_row, (size_t)std::max(B, 1) * sizeof(int64_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
    SR2_CUDA(ctx, cudaMemcpyAsync(out->d_root_node, h_root_node, (size_t)std::max(B, 1) * sizeof(int32_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
    if (want_preorder) {
        int32_t *pre = nullptr, *inst = nullptr;
        {
            void* ptr = nullptr;
            size_t bytes = (size_t)std::max<int64_t>(N, 1) * sizeof(int32_t);
            int rc = sr2_pin_reserve(ctx, HP_NODE_PRE, bytes, &ptr);
            if (rc) return rc;
            pre = (int32_t*)ptr;
            rc = sr2_pin_reserve(ctx, HP_NODE_INST, bytes, &ptr);
            if (rc) return rc;
            inst = (int32_t*)ptr;
            rc = sr2_dev_reserve(ctx, DP_NODE_PRE, bytes, &ptr);
            if (rc) return rc;
            out->d_node_pre = (int32_t*)ptr;
            rc = sr2_dev_reserve(ctx, DP_NODE_INST, bytes, &ptr);
            if (rc) return rc;
            out->d_node_inst = (int32_t*)ptr;
        }
#pragma omp parallel for num_threads(ctx->host_threads) schedule(dynamic, 64)
        for (int b = 0; b < B; ++b) {
            int64_t off = node_off[b], G = node_off[b + 1] - off;
            std::vector<int32_t> stack;
            stack.push_back(0);
            int32_t clock = 0;
            while (!stack.empty()) {
                int32_t u = stack.back();
                stack.pop_back();
                pre[off + u] = clock++;
                inst[off + u] = b;
                if (left[off + u] >= 0) {
                    stack.push_back(right[off + u]);
                    stack.push_back(left[off + u]);
                }
            }
            
        }
        SR2_CUDA(ctx, cudaMemcpyAsync(out->d_node_pre, pre, (size_t)std::max<int64_t>(N, 1) * sizeof(int32_t),
                                      cudaMemcpyHostToDevice, ctx->stream));
        SR2_CUDA(ctx, cudaMemcpyAsync(out->d_node_inst, inst, (size_t)std::max<int64_t>(N, 1) * sizeof(int32_t),
                                      cudaMemcpyHostToDevice, ctx->stream));
    }
    SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (keep_host) {
        out->h_node_row.assign(node_row, node_row + N);
        out->h_leaf.assign(leaf_species, leaf_species + N);
    }
    return SR2_OK;
}
