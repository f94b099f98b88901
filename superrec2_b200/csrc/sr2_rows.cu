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
    // small batches stay on the calling thread: waking an OpenMP team costs more than the work, and
    // under a busy host (other runtimes' teams spinning) it costs milliseconds (profiles/README.md r3h)
    const int T = N < 65536 ? 1 : std::max(1, ctx->host_threads);
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
#pragma omp parallel for num_threads(ctx->host_threads) schedule(dynamic, 64) if (N >= 65536)
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

// ---------------------------------------------------------------------------
// bucketing on the device
// ---------------------------------------------------------------------------
#define ROWS_HCAP 128  // heights a thread can count (higher trees: host path)

#define ROWS_UNKNOWN 0xffffu

// One CTA per tree.  The tree's child arrays are staged in shared memory, validated (ABI numbering
// rules; every node but the root has exactly one parent), heights are found by relaxation (a node
// becomes known once both children are; at most `height` rounds over the tree), the CTA's
// per-height counts reserve the tree's share of every wave with one atomicAdd per height.
// status[0] = first bad instance (atomicMin, INT_MAX = none), status[1] = an error code,
// status[2] = 1 if some tree is higher than ROWS_HCAP - 1.
__global__ void __launch_bounds__(128)
rows_height_kernel(int inst_begin, const int64_t* __restrict__ node_off, const int32_t* __restrict__ left,
                   const int32_t* __restrict__ right, const int32_t* __restrict__ leaf_species, int S,
                   uint16_t* __restrict__ height, int32_t* __restrict__ inst_base,
                   unsigned* __restrict__ wave_count, int* __restrict__ status) {
    extern __shared__ int32_t rows_smem[];
    __shared__ int cnt[ROWS_HCAP];
    __shared__ int pending, bad;
    const int b = inst_begin + blockIdx.x;
    const int64_t off = node_off[b];
    const int G = (int)(node_off[b + 1] - off);
    int32_t* L = rows_smem;
    int32_t* R = L + G;
    int32_t* refs = R + G;
    uint16_t* H = reinterpret_cast<uint16_t*>(refs + G);
    for (int h = threadIdx.x; h < ROWS_HCAP; h += blockDim.x) cnt[h] = 0;
    if (threadIdx.x == 0) bad = 0;
    for (int u = threadIdx.x; u < G; u += blockDim.x) refs[u] = 0;
    __syncthreads();
    for (int u = threadIdx.x; u < G; u += blockDim.x) {
        const int l = left[off + u], r = right[off + u];
        L[u] = l;
        R[u] = r;
        if (l < 0 && r < 0) {
            const int sp = leaf_species[off + u];
            if (sp < 0 || sp >= S) atomicMax(&bad, 1);
            H[u] = 0;
        } else if (l <= u || r <= u || l >= G || r >= G || l == r) {
            atomicMax(&bad, 2);
            H[u] = 0;
        } else {
            if (atomicAdd(&refs[l], 1) > 0 || atomicAdd(&refs[r], 1) > 0) atomicMax(&bad, 3);
            H[u] = ROWS_UNKNOWN;
        }
    }
    __syncthreads();
    if (!bad)
        for (int u = 1 + threadIdx.x; u < G; u += blockDim.x)
            if (refs[u] != 1) atomicMax(&bad, 4);
    __syncthreads();
    if (bad) {
        if (threadIdx.x == 0 && atomicMin(&status[0], b) > b) status[1] = bad;
        return;
    }
    // heights
    for (;;) {
        if (threadIdx.x == 0) pending = 0;
        __syncthreads();
        int mine = 0;
        for (int u = threadIdx.x; u < G; u += blockDim.x) {
            if (H[u] != ROWS_UNKNOWN) continue;
            const unsigned hl = H[L[u]], hr = H[R[u]];
            if (hl != ROWS_UNKNOWN && hr != ROWS_UNKNOWN) {
                const unsigned h = 1 + max(hl, hr);
                if (h >= ROWS_HCAP) mine = 2;
                H[u] = (uint16_t)min(h, (unsigned)ROWS_HCAP - 1);
            } else {
                mine = max(mine, 1);
            }
        }
        if (mine) atomicMax(&pending, mine);
        __syncthreads();
        const int p = pending;
        __syncthreads();
        if (p == 2) {
            if (threadIdx.x == 0) status[2] = 1;
            return;
        }
        if (p == 0) break;
    }
    for (int u = threadIdx.x; u < G; u += blockDim.x) {
        const int h = H[u];
        height[off + u] = (uint16_t)h;
        if (h) atomicAdd(&cnt[h], 1);
    }
    __syncthreads();
    for (int h = 1 + threadIdx.x; h < ROWS_HCAP; h += blockDim.x)
        if (cnt[h]) inst_base[(size_t)blockIdx.x * ROWS_HCAP + h] = (int32_t)atomicAdd(&wave_count[h], (unsigned)cnt[h]);
}

// wave_off[h] = chunk_row_begin + rows of the waves below h (exclusive prefix of the chunk's counts)
__global__ void rows_scan_kernel(const unsigned* __restrict__ wave_count, int64_t chunk_row_begin,
                                 int64_t* __restrict__ wave_off) {
    if (threadIdx.x == 0) {
        int64_t run = chunk_row_begin;
        wave_off[0] = run;
        for (int h = 1; h <= ROWS_HCAP; ++h) {
            wave_off[h] = run;
            if (h < ROWS_HCAP) run += wave_count[h];
        }
    }
}

// One CTA per tree: every internal node takes the next row of its tree's share of its wave, then
// the five row arrays are written (children's rows are all known after the first pass).
__global__ void __launch_bounds__(128)
rows_fill_kernel(int inst_begin, const int64_t* __restrict__ node_off, const int32_t* __restrict__ left,
                 const int32_t* __restrict__ right, const int32_t* __restrict__ leaf_species,
                 const uint16_t* __restrict__ height, const int32_t* __restrict__ inst_base,
                 const int64_t* __restrict__ wave_off, int64_t chunk_row_begin, int32_t* __restrict__ row_cl,
                 int32_t* __restrict__ row_cr, int32_t* __restrict__ row_node, int32_t* __restrict__ row_left,
                 int32_t* __restrict__ row_right, int64_t* __restrict__ root_row, int32_t* __restrict__ root_node,
                 const int* __restrict__ status) {
    extern __shared__ int32_t rows_smem[];
    __shared__ int cursor[ROWS_HCAP];
    if (status[0] != 0x7fffffff || status[2]) return;  // heights of this chunk are not all there
    const int b = inst_begin + blockIdx.x;
    const int64_t off = node_off[b];
    const int G = (int)(node_off[b + 1] - off);
    int32_t* node_row = rows_smem;
    for (int h = threadIdx.x; h < ROWS_HCAP; h += blockDim.x)
        cursor[h] = h ? (int)wave_off[h] + inst_base[(size_t)blockIdx.x * ROWS_HCAP + h] : 0;
    __syncthreads();
    for (int u = threadIdx.x; u < G; u += blockDim.x) {
        const int h = height[off + u];
        node_row[u] = h ? atomicAdd(&cursor[h], 1) : -1 - leaf_species[off + u];
    }
    __syncthreads();
    for (int u = threadIdx.x; u < G; u += blockDim.x) {
        const int gr = node_row[u];
        if (gr < 0) continue;
        const int l = left[off + u], r = right[off + u];
        const int rl = node_row[l], rr = node_row[r];
        row_cl[gr] = rl >= 0 ? rl - (int)chunk_row_begin : rl;
        row_cr[gr] = rr >= 0 ? rr - (int)chunk_row_begin : rr;
        row_node[gr] = (int32_t)(off + u);
        row_left[gr] = (int32_t)(off + l);
        row_right[gr] = (int32_t)(off + r);
    }
    if (threadIdx.x == 0) {
        root_node[b] = (int32_t)off;
        root_row[b] = (int64_t)node_row[0];
    }
}

// 16-bit raw arrays (sr2_dtl_run16) to the 32-bit arrays the bucketing kernels read
__global__ void rows_widen_kernel(const int16_t* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

static bool rows_is_pinned(const void* ptr) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

int sr2_build_rows_device(sr2_ctx* ctx, const char* who, int32_t B, const int64_t* node_off,
                          const int32_t* left, const int32_t* right, const int32_t* leaf_species,
                          int64_t rows_budget, RowLayout* out, const RowHooks* hooks, int64_t first_chunk_rows) {
    const std::string tag = std::string(who) + ": ";
    const int S = ctx->S;
    const double trace_begin = rows_now_ms();
    const int64_t N = B ? node_off[B] : 0;
    if (N >= (int64_t)1 << 31)
        return sr2_fail(ctx, SR2_ERR_RANGE, tag + "more than 2^31 object nodes in one batch");
    out->B = B;
    out->N = N;
    out->node_off.assign(node_off, node_off + B + 1);
    // ---- plan (as in sr2_build_rows) -------------------------------------------------------
    out->chunks.clear();
    int64_t maxG = 0, max_chunk_nodes = 0;
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
    const size_t K = out->chunks.size();
    out->max_nodes = maxG;
    out->max_chunk_rows = 0;
    int max_chunk_inst = 0;
    for (auto& c : out->chunks) {
        out->max_chunk_rows = std::max(out->max_chunk_rows, c.row_end - c.row_begin);
        max_chunk_nodes = std::max(max_chunk_nodes, node_off[c.inst_end] - node_off[c.inst_begin]);
        max_chunk_inst = std::max(max_chunk_inst, c.inst_end - c.inst_begin);
    }
    // ---- buffers -------------------------------------------------------------------------------
    const size_t slab_elems = (size_t)std::max<int64_t>(R, 1) * 5;
    out->release();
    int32_t *d_raw = nullptr, *d_inst_base = nullptr;
    int64_t* d_wave_off = nullptr;  // [K][ROWS_HCAP + 1]
    uint16_t* d_height = nullptr;
    unsigned* d_counts = nullptr;   // [K][ROWS_HCAP] wave counts, then [K][4] status
    int64_t* d_node_off = nullptr;
    unsigned* h_counts = nullptr;
    const size_t counts_words = K * (ROWS_HCAP + 4);
    {
        void* ptr = nullptr;
        int rc;
#define ROWS_RESERVE(call, var, type)  \
        if ((rc = (call))) return rc;  \
        var = (type)ptr;
        ROWS_RESERVE(sr2_dev_reserve(ctx, DP_ROWS_SLAB, slab_elems * sizeof(int32_t), &ptr), out->d_slab, int32_t*)
        ROWS_RESERVE(sr2_dev_reserve(ctx, DP_ROOT_ROW, (size_t)std::max(B, 1) * sizeof(int64_t), &ptr), out->d_root_row, int64_t*)
        ROWS_RESERVE(sr2_dev_reserve(ctx, DP_ROOT_NODE, (size_t)std::max(B, 1) * sizeof(int32_t), &ptr), out->d_root_node, int32_t*)
        // raw arrays, heights and node rows of every chunk (chunks overlap in time)
        ROWS_RESERVE(sr2_dev_reserve(ctx, DP_RAW, (size_t)std::max<int64_t>(N, 1) * 3 * sizeof(int32_t), &ptr), d_raw, int32_t*)
        ROWS_RESERVE(sr2_dev_reserve(ctx, DP_RAW_HEIGHT, (size_t)std::max<int64_t>(N, 1) * sizeof(uint16_t), &ptr), d_height, uint16_t*)
        ROWS_RESERVE(sr2_dev_reserve(ctx, DP_RAW_NODEROW, K * (ROWS_HCAP + 1) * sizeof(int64_t), &ptr), d_wave_off, int64_t*)
        ROWS_RESERVE(sr2_dev_reserve(ctx, DP_RAW_INSTBASE, (size_t)std::max(B, 1) * ROWS_HCAP * sizeof(int32_t), &ptr), d_inst_base, int32_t*)
        ROWS_RESERVE(sr2_dev_reserve(ctx, DP_RAW_COUNTS, counts_words * sizeof(unsigned), &ptr), d_counts, unsigned*)
        ROWS_RESERVE(sr2_dev_reserve(ctx, DP_RAW_NODEOFF, (size_t)(B + 1) * sizeof(int64_t), &ptr), d_node_off, int64_t*)
        ROWS_RESERVE(sr2_pin_reserve(ctx, HP_COUNTS, counts_words * sizeof(unsigned), &ptr), h_counts, unsigned*)
#undef ROWS_RESERVE
    }
    out->row_cl = out->d_slab;
    out->row_cr = out->row_cl + R;
    out->row_node = out->row_cr + R;
    out->row_left = out->row_node + R;
    out->row_right = out->row_left + R;
    int32_t *d_left = d_raw, *d_right = d_raw + N, *d_leaf = d_raw + 2 * N;
    // one CTA per tree with the tree in shared memory
    const size_t smem_height = (size_t)maxG * 14 + 16, smem_fill = (size_t)maxG * 4 + 16;
    if (smem_height > ctx->smem_optin) return SR2_ROWS_FALLBACK;
    SR2_CUDA(ctx, cudaFuncSetAttribute(rows_height_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_height));
    SR2_CUDA(ctx, cudaFuncSetAttribute(rows_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fill));
    // bucketing kernels are small and latency-bound: they must not queue behind the wave kernels'
    // CTAs, which fill every SM
    {
        int least = 0, greatest = 0;
        SR2_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));
        for (cudaStream_t* s : {&ctx->upload_stream, &ctx->fill_stream})
            if (!*s) SR2_CUDA(ctx, cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, greatest));
    }
    while (ctx->chunk_events.size() < 2 * K) {
        cudaEvent_t e;
        SR2_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_events.push_back(e);
    }
    if (hooks->on_plan) {
        int rc = hooks->on_plan(hooks->user);
        if (rc) return rc;
    }
    // the caller's arrays go to the device directly when they are pinned, else through pinned staging
    const bool narrow = hooks->left16 != nullptr;   // 16-bit raw arrays: half the bytes over PCIe
    const size_t elem = narrow ? sizeof(int16_t) : sizeof(int32_t);
    const char* src_all[3] = {narrow ? (const char*)hooks->left16 : (const char*)left,
                              narrow ? (const char*)hooks->right16 : (const char*)right,
                              narrow ? (const char*)hooks->leaf16 : (const char*)leaf_species};
    const bool direct = rows_is_pinned(src_all[0]) && rows_is_pinned(src_all[1]) && rows_is_pinned(src_all[2]);
    char* stage = nullptr;
    if (!direct) {
        void* ptr = nullptr;
        int rc = sr2_pin_reserve(ctx, HP_RAW, (size_t)std::max<int64_t>(N, 1) * 3 * elem, &ptr);
        if (rc) return rc;
        stage = (char*)ptr;
    }
    int16_t* d_raw16 = nullptr;
    if (narrow) {
        void* ptr = nullptr;
        int rc = sr2_dev_reserve(ctx, DP_RAW16, (size_t)std::max<int64_t>(N, 1) * 3 * sizeof(int16_t), &ptr);
        if (rc) return rc;
        d_raw16 = (int16_t*)ptr;
    }
    cudaStream_t up = ctx->upload_stream;
    const bool trace = getenv("SR2_TRACE") != nullptr;
    const double trace_t0 = rows_now_ms();
    if (trace) fprintf(stderr, "[sr2] device bucketing: plan + buffers + on_plan took %.2f ms, inputs %s\n",
                       trace_t0 - trace_begin, direct ? "pinned" : "pageable (staged)");
    SR2_CUDA(ctx, cudaMemcpyAsync(d_node_off, node_off, (size_t)(B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, up));
    SR2_CUDA(ctx, cudaMemsetAsync(d_counts, 0, counts_words * sizeof(unsigned), up));
    // ---- phase 1 of every chunk: raw arrays up, heights + counts, counts back --------------------
    for (size_t ci = 0; ci < K; ++ci) {
        const RowChunk& c = out->chunks[ci];
        const int nb = c.inst_end - c.inst_begin;
        const int64_t n0 = node_off[c.inst_begin], nn = node_off[c.inst_end] - n0;
        int32_t* dst[3] = {d_left, d_right, d_leaf};
        for (int k = 0; k < 3; ++k) {
            const char* from = src_all[k] + (size_t)n0 * elem;
            if (!direct) {
                char* to = stage + ((size_t)k * N + n0) * elem;
                const int64_t block = 1 << 16, n_blocks = (nn + block - 1) / block;
#pragma omp parallel for num_threads(ctx->host_threads) schedule(static) if (n_blocks > 1)
                for (int64_t q = 0; q < n_blocks; ++q) {
                    int64_t lo = q * block, hi = std::min(nn, lo + block);
                    memcpy(to + lo * elem, from + lo * elem, (size_t)(hi - lo) * elem);
                }
                from = to;
            }
            if (!narrow) {
                SR2_CUDA(ctx, cudaMemcpyAsync(dst[k] + n0, from, (size_t)nn * elem, cudaMemcpyHostToDevice, up));
            } else if (nn > 0) {
                int16_t* d16 = d_raw16 + (size_t)k * N + n0;
                SR2_CUDA(ctx, cudaMemcpyAsync(d16, from, (size_t)nn * elem, cudaMemcpyHostToDevice, up));
                rows_widen_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, up>>>(d16, dst[k] + n0, nn);
            }
        }
        unsigned* counts = d_counts + ci * (ROWS_HCAP + 4);
        int* status = (int*)(counts + ROWS_HCAP);
        const int init_status[4] = {0x7fffffff, 0, 0, 0};
        memcpy(h_counts + ci * (ROWS_HCAP + 4) + ROWS_HCAP, init_status, sizeof(init_status));
        SR2_CUDA(ctx, cudaMemcpyAsync(status, h_counts + ci * (ROWS_HCAP + 4) + ROWS_HCAP, sizeof(init_status),
                                      cudaMemcpyHostToDevice, up));
        if (nb > 0) {
            rows_height_kernel<<<nb, 128, smem_height, up>>>(c.inst_begin, d_node_off, d_left, d_right, d_leaf, S, d_height,
                                                             d_inst_base + (size_t)c.inst_begin * ROWS_HCAP, counts,
                                                             status);
            rows_scan_kernel<<<1, 32, 0, up>>>(counts, c.row_begin, d_wave_off + ci * (ROWS_HCAP + 1));
        }
        SR2_CUDA(ctx, cudaGetLastError());
        SR2_CUDA(ctx, cudaMemcpyAsync(h_counts + ci * (ROWS_HCAP + 4), counts, (ROWS_HCAP + 4) * sizeof(unsigned),
                                      cudaMemcpyDeviceToHost, up));
        SR2_CUDA(ctx, cudaEventRecord(ctx->chunk_events[K + ci], up));
        // rows: no host round trip needed (a malformed or too-high tree leaves garbage rows behind,
        // but the host looks at the chunk's status before it launches anything that reads them)
        if (nb > 0)
            rows_fill_kernel<<<nb, 128, smem_fill, up>>>(
                c.inst_begin, d_node_off, d_left, d_right, d_leaf, d_height, d_inst_base + (size_t)c.inst_begin * ROWS_HCAP,
                d_wave_off + ci * (ROWS_HCAP + 1), c.row_begin, out->row_cl, out->row_cr, out->row_node, out->row_left,
                out->row_right, out->d_root_row, out->d_root_node, status);
        SR2_CUDA(ctx, cudaGetLastError());
        SR2_CUDA(ctx, cudaEventRecord(ctx->chunk_events[ci], up));
    }
    if (trace) fprintf(stderr, "[sr2] device bucketing: phase 1 of %zu chunks enqueued at %.2f ms\n", K,
                       rows_now_ms() - trace_t0);
    // ---- phase 2: wave offsets on the host, rows on the device, the caller's work --------------------
    out->max_height = 0;
    for (size_t ci = 0; ci < K; ++ci) {
        RowChunk& c = out->chunks[ci];
        const int nb = c.inst_end - c.inst_begin;
        SR2_CUDA(ctx, cudaEventSynchronize(ctx->chunk_events[K + ci]));
        const unsigned* counts = h_counts + ci * (ROWS_HCAP + 4);
        const int* status = (const int*)(counts + ROWS_HCAP);
        if (status[2] || status[0] != 0x7fffffff) {
            cudaStreamSynchronize(up);
            cudaStreamSynchronize(ctx->stream);
            if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
            if (status[0] != 0x7fffffff) {
                static const char* const why[] = {"", "leaf species out of range",
                                                  "children must be two distinct nodes numbered after their parent",
                                                  "a node has two parents", "a node other than node 0 has no parent"};
                int code = status[1] >= 1 && status[1] <= 4 ? status[1] : 2;
                return sr2_fail(ctx, SR2_ERR_ARG, tag + "instance " + std::to_string(status[0]) + ": " + why[code]);
            }
            return SR2_ROWS_FALLBACK;
        }
        int H = 0;
        for (int h = 1; h < ROWS_HCAP; ++h)
            if (counts[h]) H = h;
        c.wave_off.assign(H + 2, 0);
        int64_t run = c.row_begin;
        c.wave_off[0] = run;
        for (int h = 1; h <= H; ++h) {
            c.wave_off[h] = run;
            run += counts[h];
        }
        c.wave_off[H + 1] = run;
        if (run != c.row_end) {
            cudaStreamSynchronize(up);
            return sr2_fail(ctx, SR2_ERR_ARG, tag + "internal error: row count of a chunk does not match its trees");
        }
        out->max_height = std::max(out->max_height, H);
        int rc = hooks->on_chunk(hooks->user, (int)ci);
        if (rc) return rc;
        if (trace)
            fprintf(stderr, "[sr2] chunk %zu: %d instances bucketed on the device, enqueued at %.2f ms\n", ci, nb,
                    rows_now_ms() - trace_t0);
    }
    SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SR2_CUDA(ctx, cudaStreamSynchronize(up));
    if (trace) fprintf(stderr, "[sr2] device bucketing: streams idle at %.2f ms\n", rows_now_ms() - trace_t0);
    return SR2_OK;
}
