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
    cudaFree(d_slab);
    cudaFree(d_root_row);
    cudaFree(d_root_node);
    cudaFree(d_node_pre);
    cudaFree(d_node_inst);
    d_slab = nullptr;
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
    std::vector<uint16_t> height(N);
    int bad_inst = -1;
    std::string bad_msg;
    int Hmax = 0;
    int64_t maxG = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(max : Hmax) reduction(max : maxG)
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
    std::vector<int32_t> slab((size_t)std::max<int64_t>(R, 1) * 5);
    int32_t* h_cl = slab.data();
    int32_t* h_cr = h_cl + R;
    int32_t* h_node = h_cr + R;
    int32_t* h_left = h_node + R;
    int32_t* h_right = h_left + R;
    std::vector<int64_t> h_root_row(std::max(B, 1));
    std::vector<int32_t> h_root_node(std::max(B, 1));
    std::vector<int64_t> node_row(N, -1);
    const int W = Hmax + 1;
    out->max_chunk_rows = 0;
    for (auto& c : out->chunks) {
        int nb = c.inst_end - c.inst_begin;
        std::vector<int64_t> cnt((size_t)std::max(nb, 1) * W, 0);
#pragma omp parallel for schedule(dynamic, 64)
        for (int b = c.inst_begin; b < c.inst_end; ++b) {
            int64_t off = node_off[b], G = node_off[b + 1] - off;
            int64_t* mine = cnt.data() + (size_t)(b - c.inst_begin) * W;
            for (int64_t u = 0; u < G; ++u) mine[height[off + u]]++;
        }
        c.wave_off.assign(W + 1, 0);
        int64_t run = c.row_begin;
        c.wave_off[0] = run;
        for (int h = 1; h < W; ++h) {
            c.wave_off[h] = run;
            for (int i = 0; i < nb; ++i) {
                int64_t k = cnt[(size_t)i * W + h];
                cnt[(size_t)i * W + h] = run;
                run += k;
            }
        }
        c.wave_off[W] = run;
        out->max_chunk_rows = std::max(out->max_chunk_rows, c.row_end - c.row_begin);
#pragma omp parallel for schedule(dynamic, 64)
        for (int b = c.inst_begin; b < c.inst_end; ++b) {
            int64_t off = node_off[b], G = node_off[b + 1] - off;
            int64_t* cursor = cnt.data() + (size_t)(b - c.inst_begin) * W;
            for (int64_t u = 0; u < G; ++u) {
                int h = height[off + u];
                if (h > 0) node_row[off + u] = cursor[h]++;
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

    // ---- device copies -----------------------------------------------------------------
    out->release();
    SR2_CUDA(ctx, cudaMalloc(&out->d_slab, slab.size() * sizeof(int32_t)));
    out->row_cl = out->d_slab;
    out->row_cr = out->row_cl + R;
    out->row_node = out->row_cr + R;
    out->row_left = out->row_node + R;
    out->row_right = out->row_left + R;
    SR2_CUDA(ctx, cudaMalloc(&out->d_root_row, h_root_row.size() * sizeof(int64_t)));
    SR2_CUDA(ctx, cudaMalloc(&out->d_root_node, h_root_node.size() * sizeof(int32_t)));
    SR2_CUDA(ctx, cudaMemcpyAsync(out->d_slab, slab.data(), slab.size() * sizeof(int32_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
    SR2_CUDA(ctx, cudaMemcpyAsync(out->d_root_row, h_root_row.data(), h_root_row.size() * sizeof(int64_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
    SR2_CUDA(ctx, cudaMemcpyAsync(out->d_root_node, h_root_node.data(), h_root_node.size() * sizeof(int32_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int32_t> pre, inst;
    if (want_preorder) {
        pre.assign(std::max<int64_t>(N, 1), 0);
        inst.assign(std::max<int64_t>(N, 1), 0);
#pragma omp parallel for schedule(dynamic, 64)
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
        SR2_CUDA(ctx, cudaMalloc(&out->d_node_pre, pre.size() * sizeof(int32_t)));
        SR2_CUDA(ctx, cudaMalloc(&out->d_node_inst, inst.size() * sizeof(int32_t)));
        SR2_CUDA(ctx, cudaMemcpyAsync(out->d_node_pre, pre.data(), pre.size() * sizeof(int32_t),
                                      cudaMemcpyHostToDevice, ctx->stream));
        SR2_CUDA(ctx, cudaMemcpyAsync(out->d_node_inst, inst.data(), inst.size() * sizeof(int32_t),
                                      cudaMemcpyHostToDevice, ctx->stream));
    }
    SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (keep_host) {
        out->h_node_row.swap(node_row);
        out->h_leaf.assign(leaf_species, leaf_species + N);
    }
    return SR2_OK;
}
