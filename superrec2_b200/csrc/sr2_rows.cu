// sr2_rows.cu -- host-side bucketing of a batch into waves (see sr2_rows.cuh).
// Replaces the postorder loop of the reference's table drivers
// (/root/reference/src/superrec2/compute/reconciliation.py:202-223,
//  compute/unordered_super_reconciliation.py:329-356): nodes of equal height are
// independent, so they are grouped into one launch.
#include <algorithm>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <omp.h>

#include "sr2_rows.cuh"

static double rows_now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

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
                   int64_t rows_budget, bool keep_host, bool want_preorder, RowLayout* out,
                   const RowHooks* hooks, int64_t first_chunk_rows) {
    const std::string tag = std::string(who) + ": ";
    const int S = ctx->S;
    const int64_t N = B ? node_off[B] : 0;
    if (N >= (int64_t)1 << 31)
        return sr2_fail(ctx, SR2_ERR_RANGE, tag + "more than 2^31 object nodes in one batch");
    out->B = B;
    out->N = N;
    out->node_off.assign(node_off, node_off + B + 1);

    // ---- plan: chunks of instances from the node counts alone (a binary tree of G nodes has
    //      (G - 1) / 2 internal nodes; the trees themselves are validated chunk by chunk) ------
    out->chunks.clear();
    int64_t maxG = 0;
    {
        RowChunk c;
        int64_t rows = 0, total_rows = 0;
        for (int b = 0; b < B; ++b) {
            int64_t G = node_off[b + 1] - node_off[b];
            if (G < 1) return sr2_fail(ctx, SR2_ERR_ARG, tag + "instance " + std::to_string(b) + ": empty object tree");
            maxG = std::max(maxG, G);
            int64_t mine = (G - 1) / 2;
            if (mine > rows_budget)
                return sr2_fail(ctx, SR2_ERR_NOMEM, tag + "one instance alone exceeds device memory");
            // a pipelined caller starts the device early with a half-sized first chunk
            const int64_t budget = (first_chunk_rows > 0 && out->chunks.empty()) ? std::min(first_chunk_rows, rows_budget)
                                                                                  : rows_budget;
            if (rows + mine > budget && rows > 0) {
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
    out->max_nodes = maxG;
    out->max_chunk_rows = 0;
    for (auto& c : out->chunks) out->max_chunk_rows = std::max(out->max_chunk_rows, c.row_end - c.row_begin);

    // ---- buffers: pinned staging (reused across uploads: full-speed H2D, no page faults),
    //      host scratch, pooled device arrays -----------------------------------------------
    const size_t slab_elems = (size_t)std::max<int64_t>(R, 1) * 5;
    uint16_t* height = nullptr;
    int32_t* h_cl = nullptr;
    int64_t* h_root_row = nullptr;
    int32_t* h_root_node = nullptr;
    out->release();
    {
        void* ptr = nullptr;
        int rc = sr2_host_reserve(ctx, MP_HEIGHT, (size_t)std::max<int64_t>(N, 1) * sizeof(uint16_t), &ptr);
        if (rc) return rc;
        height = (uint16_t*)ptr;
        rc = sr2_pin_reserve(ctx, HP_ROWS_SLAB, slab_elems * sizeof(int32_t), &ptr);
        if (rc) return rc;
        h_cl = (int32_t*)ptr;
        rc = sr2_pin_reserve(ctx, HP_ROOT_ROW, (size_t)std::max(B, 1) * sizeof(int64_t), &ptr);
        if (rc) return rc;
        h_root_row = (int64_t*)ptr;
        rc = sr2_pin_reserve(ctx, HP_ROOT_NODE, (size_t)std::max(B, 1) * sizeof(int32_t), &ptr);
        if (rc) return rc;
        h_root_node = (int32_t*)ptr;
        rc = sr2_dev_reserve(ctx, DP_ROWS_SLAB, slab_elems * sizeof(int32_t), &ptr);
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
    int32_t* h_arr[5] = {h_cl, h_cl + R, h_cl + 2 * R, h_cl + 3 * R, h_cl + 4 * R};
    int32_t* d_arr[5] = {out->row_cl, out->row_cr, out->row_node, out->row_left, out->row_right};
    int32_t *h_cr = h_arr[1], *h_node = h_arr[2], *h_left = h_arr[3], *h_right = h_arr[4];
    if (hooks && hooks->on_plan) {
        int rc = hooks->on_plan(hooks->user);
        if (rc) return rc;
    }

    // ---- chunk by chunk: validate + heights, rows in (height, instance, node) order, H2D ----
    // One parallel region per chunk.  Every thread owns a contiguous block of instances (balanced
    // by node count) and (1) validates its trees and counts its internal nodes per height, then
    // after a prefix sum over (height, thread) (2) hands out rows bottom-up, so that the rows
    // of a node's children are known when the node itself is written.
    out->max_height = 0;
    const bool trace = getenv("SR2_TRACE") != nullptr;
    const double trace_t0 = rows_now_ms();
    const bool async_h2d = hooks && hooks->on_chunk;
    if (async_h2d) {
        if (!ctx->upload_stream)
            SR2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking));
        while (ctx->chunk_events.size() < out->chunks.size()) {
            cudaEvent_t e;
            SR2_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->chunk_events.push_back(e);
        }
    }
    if (keep_host) out->h_node_row.assign((size_t)N, -1);
    const int T = std::max(1, ctx->host_threads);
    std::vector<std::vector<int32_t>> counts(T);   // [thread][height] internal nodes, then row cursors
    std::vector<int> first_inst(T + 1);
    for (size_t ci = 0; ci < out->chunks.size(); ++ci) {
        RowChunk& c = out->chunks[ci];
        const int nb = c.inst_end - c.inst_begin;
        double tr0 = rows_now_ms(), tr2 = 0, tr3 = 0;
        {   // instance blocks of about equal node counts
            const int64_t n0 = node_off[c.inst_begin], n1 = node_off[c.inst_end];
            int b = c.inst_begin;
            for (int t = 0; t < T; ++t) {
                first_inst[t] = b;
                int64_t target = n0 + (n1 - n0) * (t + 1) / T;
                while (b < c.inst_end && node_off[b + 1] <= target) ++b;
            }
            first_inst[T] = c.inst_end;
            first_inst[T - 1] = std::min(first_inst[T - 1], c.inst_end);
        }
        int bad_inst = -1;
        const char* bad_msg = nullptr;
        int Hmax = 0;
        c.wave_off.clear();
#pragma omp parallel num_threads(T)
        {
            const int t = omp_get_thread_num();
            std::vector<int32_t>& cnt = counts[t];
            cnt.assign(64, 0);
            std::vector<uint8_t> refs;
            std::vector<int32_t> local_row;
            // (1) validate, heights, counts
            if (t < T)
                for (int b = first_inst[t]; b < first_inst[t + 1]; ++b) {
                    const int64_t off = node_off[b], G = node_off[b + 1] - off;
                    const int32_t *L = left + off, *Rt = right + off, *Sp = leaf_species + off;
                    uint16_t* Ht = height + off;
                    const char* why = nullptr;
                    refs.assign((size_t)G, 0);
                    for (int64_t u = G - 1; u >= 0; --u) {
                        const int l = L[u], r = Rt[u];
                        if (l < 0 && r < 0) {
                            const int sp = Sp[u];
                            if (sp < 0 || sp >= S) { why = "leaf species out of range"; break; }
                            Ht[u] = 0;
                        } else if (l <= u || r <= u || l >= G || r >= G || l == r) {
                            why = "children must be two distinct nodes numbered after their parent";
                            break;
                        } else {
                            if (++refs[l] > 1 || ++refs[r] > 1) { why = "a node has two parents"; break; }
                            const int h = 1 + std::max(Ht[l], Ht[r]);
                            if (h > 65535) { why = "object tree higher than 65535"; break; }
                            Ht[u] = (uint16_t)h;
                            if ((size_t)h >= cnt.size()) cnt.resize(std::max<size_t>(2 * cnt.size(), (size_t)h + 1), 0);
                            cnt[h]++;
                        }
                    }
                    if (!why)
                        for (int64_t u = 1; u < G; ++u)
                            if (refs[u] != 1) { why = "a node other than node 0 has no parent"; break; }
                    if (why) {
#pragma omp critical
                        if (bad_inst < 0 || b < bad_inst) {
                            bad_inst = b;
                            bad_msg = why;
                        }
                        break;
                    }
                }
#pragma omp barrier
#pragma omp single
            {
                if (bad_inst < 0) {
                    size_t W = 1;
                    for (int k = 0; k < T; ++k) {
                        size_t top = counts[k].size();
                        while (top > 1 && counts[k][top - 1] == 0) --top;
                        W = std::max(W, top);
                    }
                    for (int k = 0; k < T; ++k) counts[k].resize(W, 0);
                    Hmax = (int)W - 1;
                    c.wave_off.assign(W + 1, 0);
                    int64_t run = c.row_begin;
                    c.wave_off[0] = run;
                    for (size_t h = 1; h < W; ++h) {
                        c.wave_off[h] = run;
                        for (int k = 0; k < T; ++k) {
                            int32_t n = counts[k][h];
                            counts[k][h] = (int32_t)run;
                            run += n;
                        }
                    }
                    c.wave_off[W] = run;
                }
            }   // implicit barrier
            // (2) rows, bottom-up inside every tree
            if (bad_inst < 0)
                for (int b = first_inst[t]; b < first_inst[t + 1]; ++b) {
                    const int64_t off = node_off[b], G = node_off[b + 1] - off;
                    const int32_t *L = left + off, *Rt = right + off, *Sp = leaf_species + off;
                    const uint16_t* Ht = height + off;
                    local_row.resize((size_t)G);
                    for (int64_t u = G - 1; u >= 0; --u) {
                        const int h = Ht[u];
                        if (h == 0) {
                            local_row[u] = -1 - Sp[u];  // a leaf as a child descriptor
                            continue;
                        }
                        const int32_t gr = cnt[h]++;
                        const int l = L[u], r = Rt[u];
                        const int32_t rl = local_row[l], rr = local_row[r];
                        // descriptors: >= 0 the child's row inside the chunk, < 0 = -1 - leaf species
                        h_cl[gr] = rl >= 0 ? rl - (int32_t)c.row_begin : rl;
                        h_cr[gr] = rr >= 0 ? rr - (int32_t)c.row_begin : rr;
                        h_node[gr] = (int32_t)(off + u);
                        h_left[gr] = (int32_t)(off + l);
                        h_right[gr] = (int32_t)(off + r);
                        local_row[u] = gr;
                    }
                    h_root_node[b] = (int32_t)off;
                    h_root_row[b] = (int64_t)local_row[0];  // a one-node tree: -1 - leaf species
                    if (keep_host)
                        for (int64_t u = 0; u < G; ++u)
                            out->h_node_row[off + u] = Ht[u] ? local_row[u] : -1;
                }
        }
        if (bad_inst >= 0) {
            cudaStreamSynchronize(ctx->stream);  // earlier chunks may be in flight
            if (ctx->upload_stream) cudaStreamSynchronize(ctx->upload_stream);
            return sr2_fail(ctx, SR2_ERR_ARG, tag + "instance " + std::to_string(bad_inst) + ": " + bad_msg);
        }
        out->max_height = std::max(out->max_height, Hmax);
        tr2 = rows_now_ms();
        // this chunk's slices of the row arrays go to the device while the next chunk is prepared
        // (on their own stream when the caller overlaps device work with the bucketing)
        cudaStream_t up = async_h2d ? ctx->upload_stream : ctx->stream;
        const size_t n_rows = (size_t)(c.row_end - c.row_begin);
        if (n_rows)
            for (int k = 0; k < 5; ++k)
                SR2_CUDA(ctx, cudaMemcpyAsync(d_arr[k] + c.row_begin, h_arr[k] + c.row_begin, n_rows * sizeof(int32_t),
                                              cudaMemcpyHostToDevice, up));
        if (nb) {
            SR2_CUDA(ctx, cudaMemcpyAsync(out->d_root_row + c.inst_begin, h_root_row + c.inst_begin,
                                          (size_t)nb * sizeof(int64_t), cudaMemcpyHostToDevice, up));
            SR2_CUDA(ctx, cudaMemcpyAsync(out->d_root_node + c.inst_begin, h_root_node + c.inst_begin,
                                          (size_t)nb * sizeof(int32_t), cudaMemcpyHostToDevice, up));
        }
        // on_chunk must make the stream it computes on wait for ctx->chunk_events[ci]
        if (async_h2d) SR2_CUDA(ctx, cudaEventRecord(ctx->chunk_events[ci], up));
        tr3 = rows_now_ms();
        if (hooks && hooks->on_chunk) {
            int rc = hooks->on_chunk(hooks->user, (int)ci);
            if (rc) return rc;
        }
        if (trace)
            fprintf(stderr, "[sr2] chunk %zu: %d instances, host: bucket %.2f .. %.2f ms, h2d enqueued %.2f, hook done %.2f ms\n",
                    ci, nb, tr0 - trace_t0, tr2 - trace_t0, tr3 - trace_t0, rows_now_ms() - trace_t0);
    }

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
            int64_t off = node_off[b];
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
    if (async_h2d) SR2_CUDA(ctx, cudaStreamSynchronize(ctx->upload_stream));
    if (keep_host) out->h_leaf.assign(leaf_species, leaf_species + N);
    return SR2_OK;
}
