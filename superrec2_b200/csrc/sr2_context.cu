// sr2_context.cu -- context, error reporting and the species-tree layout.
//
// The species tree replaces the reference's LowestCommonAncestor oracle
// (/root/reference/src/superrec2/utils/trees.py:31-142): instead of an Euler
// tour + sparse table queried per candidate, every node gets its depth and a
// preorder [tin, tout) interval once, and the DP kernels only use look-ups.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "sr2_internal.cuh"

int sr2_fail(sr2_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

int sr2_cuda_fail(sr2_ctx* ctx, cudaError_t e, const char* what) {
    std::string msg = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    cudaGetLastError();  // clear the sticky flag where possible
    return sr2_fail(ctx, SR2_ERR_CUDA, msg);
}

int sr2_dev_reserve(sr2_ctx* ctx, int id, size_t bytes, void** out) {
    Sr2Pool& pool = ctx->dev[id];
    bytes = std::max<size_t>(bytes, 256);
    if (pool.cap < bytes) {
        ctx->mem_dirty = true;
        if (pool.ptr) cudaFree(pool.ptr);
        pool.ptr = nullptr;
        pool.cap = 0;
        // a little head-room: batches of similar size reuse the buffer instead of reallocating it
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&pool.ptr, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&pool.ptr, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return sr2_fail(ctx, SR2_ERR_NOMEM, std::string("device allocation of ") + std::to_string(bytes) +
                                                    " bytes failed: " + cudaGetErrorString(e));
        }
        pool.cap = want;
    }
    *out = pool.ptr;
    return SR2_OK;
}

int sr2_pin_reserve(sr2_ctx* ctx, int id, size_t bytes, void** out) {
    Sr2Pool& pool = ctx->pin[id];
    bytes = std::max<size_t>(bytes, 256);
    if (pool.cap < bytes) {
        if (pool.ptr) cudaFreeHost(pool.ptr);
        pool.ptr = nullptr;
        pool.cap = 0;
        size_t want = bytes + bytes / 8;  // a little head-room: batches of similar size reuse it
        cudaError_t e = cudaHostAlloc(&pool.ptr, want, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return sr2_fail(ctx, SR2_ERR_NOMEM, std::string("pinned host allocation failed: ") + cudaGetErrorString(e));
        }
        pool.cap = want;
    }
    *out = pool.ptr;
    return SR2_OK;
}

int sr2_host_reserve(sr2_ctx* ctx, int id, size_t bytes, void** out) {
    Sr2Pool& pool = ctx->host[id];
    bytes = std::max<size_t>(bytes, 256);
    if (pool.cap < bytes) {
        free(pool.ptr);
        size_t want = bytes + bytes / 8;
        pool.ptr = malloc(want);
        pool.cap = pool.ptr ? want : 0;
        if (!pool.ptr) return sr2_fail(ctx, SR2_ERR_NOMEM, "host allocation failed");
    }
    *out = pool.ptr;
    return SR2_OK;
}

size_t sr2_dev_pooled(const sr2_ctx* ctx, int id) { return ctx->dev[id].cap; }

int sr2_mem_free(sr2_ctx* ctx, size_t* out_free) {
    if (ctx->mem_dirty) {
        size_t free_b = 0, total_b = 0;
        SR2_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
        ctx->mem_free_cached = free_b;
        ctx->mem_dirty = false;
    }
    *out_free = ctx->mem_free_cached;
    return SR2_OK;
}

static std::string g_create_error;

extern "C" int sr2_version(void) { return 100; }

extern "C" int sr2_create(int device, sr2_ctx** out) {
    if (!out) return SR2_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device available (") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                         "); libsr2 has no CPU path";
        cudaGetLastError();
        return SR2_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        g_create_error = "device index out of range";
        return SR2_ERR_ARG;
    }
    sr2_ctx* ctx = new sr2_ctx();
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        delete ctx;
        return SR2_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        delete ctx;
        return SR2_ERR_CUDA;
    }
    if (prop.major < 10) {
        g_create_error = "libsr2 is built for sm_100a (B200) only; found sm_" +
                         std::to_string(prop.major) + std::to_string(prop.minor);
        delete ctx;
        return SR2_ERR_CUDA;
    }
    // Host threads for bucketing/unranking.  Launchers such as torchrun export
    // OMP_NUM_THREADS=1, which would serialise the host side of every upload, so the library
    // sizes its own team: SR2_HOST_THREADS, else the cores of this box divided by the ranks
    // sharing it (LOCAL_WORLD_SIZE).
    {
        int cores = (int)std::thread::hardware_concurrency();
        if (cores < 1) cores = 1;
        int ranks = 1;
        if (const char* lws = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(lws));
        ctx->host_threads = std::max(1, cores / ranks);
        if (const char* t = getenv("SR2_HOST_THREADS")) ctx->host_threads = std::max(1, atoi(t));
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
        delete ctx;
        return SR2_ERR_CUDA;
    }
    ctx->own_stream = true;
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    *out = ctx;
    return SR2_OK;
}

extern "C" void sr2_destroy(sr2_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    sr2_free_dtl(ctx);
    sr2_free_super(ctx);
    for (auto& pool : ctx->dev)
        if (pool.ptr) cudaFree(pool.ptr);
    for (auto& pool : ctx->pin)
        if (pool.ptr) cudaFreeHost(pool.ptr);
    for (auto& pool : ctx->host) free(pool.ptr);
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->fill_stream) cudaStreamDestroy(ctx->fill_stream);
    for (int k = 0; k < 2; ++k) {
        if (ctx->part_stream[k]) cudaStreamDestroy(ctx->part_stream[k]);
        for (cudaEvent_t e : ctx->part_events[k]) cudaEventDestroy(e);
        if (ctx->dense_stream[k]) cudaStreamDestroy(ctx->dense_stream[k]);
        for (int j = 0; j < 8; ++j)
            if (ctx->tier_streams[k][j]) cudaStreamDestroy(ctx->tier_streams[k][j]);
        for (cudaEvent_t e : ctx->split_events[k]) cudaEventDestroy(e);
    }
    for (cudaEvent_t e : ctx->chunk_events) cudaEventDestroy(e);
    for (auto* list : {&ctx->dtl_ev_w0, &ctx->dtl_ev_w1, &ctx->dtl_ev_done, &ctx->dtl_ev_copied})
        for (cudaEvent_t e : *list) cudaEventDestroy(e);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* sr2_last_error(const sr2_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int sr2_set_stream(sr2_ctx* ctx, void* cuda_stream) {
    if (!ctx) return SR2_ERR_ARG;
    if (ctx->own_stream && ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return SR2_OK;
}

extern "C" int sr2_synchronize(sr2_ctx* ctx) {
    if (!ctx) return SR2_ERR_ARG;
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SR2_OK;
}

extern "C" int sr2_last_timing(const sr2_ctx* ctx, sr2_timing* out) {
    if (!ctx || !out) return SR2_ERR_ARG;
    *out = ctx->timing;
    return SR2_OK;
}

extern "C" int sr2_set_policy(sr2_ctx* ctx, int policy) {
    if (!ctx) return SR2_ERR_ARG;
    if (policy != SR2_POLICY_ANY && policy != SR2_POLICY_ALL)
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_set_policy: policy must be SR2_POLICY_ANY or SR2_POLICY_ALL");
    ctx->policy = policy;
    return SR2_OK;
}

extern "C" int sr2_set_species(sr2_ctx* ctx, int32_t S, const int32_t* parent,
                               const int32_t* child0, const int32_t* child1) {
    if (!ctx) return SR2_ERR_ARG;
    if (S < 1 || !parent || !child0 || !child1)
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_set_species: empty species tree or null array");
    if (S > 65535)
        return sr2_fail(ctx, SR2_ERR_RANGE,
                        "sr2_set_species: more than 65535 species nodes (arg-min keys hold 16-bit ranks)");
    if (parent[0] != -1)
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_set_species: node 0 must be the root (parent -1)");
    // BFS numbering <=> the i-th internal node (in rank order) has children 2i+1, 2i+2.
    std::vector<int32_t> int_list;
    int next = 1;
    for (int y = 0; y < S; ++y) {
        bool leaf = child0[y] < 0;
        if (leaf != (child1[y] < 0))
            return sr2_fail(ctx, SR2_ERR_ARG, "sr2_set_species: node " + std::to_string(y) +
                                                  " has exactly one child (tree must be binary)");
        if (leaf) continue;
        if (child0[y] != next || child1[y] != next + 1 || next + 1 >= S)
            return sr2_fail(ctx, SR2_ERR_ARG,
                            "sr2_set_species: nodes are not numbered in level order at node " +
                                std::to_string(y));
        if (parent[next] != y || parent[next + 1] != y)
            return sr2_fail(ctx, SR2_ERR_ARG, "sr2_set_species: parent[] disagrees with child arrays at node " +
                                                  std::to_string(y));
        int_list.push_back(y);
        next += 2;
    }
    if (next != S)
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_set_species: child arrays do not cover all nodes");

    // The same tree again (the Python entry points set the species before every upload, and the winner
    // of an enumeration is re-solved on a re-parsed copy of the same tree): everything on the device is
    // still valid.  cudaMalloc / cudaFree are kept off this path altogether -- with peer access enabled
    // (NCCL, one process per GPU) a single allocation was seen to take 200 ms (profiles/README.md r3k).
    if (ctx->S == S && ctx->species_valid && ctx->h_parent.size() == (size_t)S &&
        memcmp(ctx->h_parent.data(), parent, sizeof(int32_t) * S) == 0 &&
        memcmp(ctx->h_child0.data(), child0, sizeof(int32_t) * S) == 0)
        return SR2_OK;
    ctx->species_valid = false;
    ctx->S = S;
    ctx->pitch = (S + 31) / 32 * 32;
    ctx->n_internal = (int)int_list.size();
    ctx->h_parent.assign(parent, parent + S);
    ctx->h_child0.assign(child0, child0 + S);
    ctx->h_depth.assign(S, 0);
    for (int y = 1; y < S; ++y) ctx->h_depth[y] = ctx->h_depth[parent[y]] + 1;
    ctx->D = ctx->h_depth[S - 1] + 1;
    if (ctx->D > 32767)
        return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_set_species: species tree deeper than 32767 levels");
    // preorder entry/exit times, iteratively
    ctx->h_tin.assign(S, 0);
    ctx->h_tout.assign(S, 0);
    {
        std::vector<int32_t> size(S, 1);
        for (int y = S - 1; y >= 1; --y) size[parent[y]] += size[y];
        ctx->h_tin[0] = 0;
        for (int y = 0; y < S; ++y) {
            ctx->h_tout[y] = ctx->h_tin[y] + size[y];
            if (child0[y] >= 0) {
                ctx->h_tin[child0[y]] = ctx->h_tin[y] + 1;
                ctx->h_tin[child1[y]] = ctx->h_tin[y] + 1 + size[child0[y]];
            }
        }
    }
    ctx->h_int_list = int_list;
    ctx->h_int_lvl_off.assign(ctx->D + 1, 0);
    for (int y : int_list) ctx->h_int_lvl_off[ctx->h_depth[y] + 1]++;
    for (int d = 0; d < ctx->D; ++d) ctx->h_int_lvl_off[d + 1] += ctx->h_int_lvl_off[d];

    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    sr2_free_dtl(ctx);
    sr2_free_super(ctx);
    SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the pooled slab below may still be read
    size_t n_int = int_list.size();
    // sweep schedule (see DevSpecies::sweep_sched)
    std::vector<int32_t> sched;
    for (int d = 0; d + 1 < ctx->D; ++d) {
        int b = ctx->h_int_lvl_off[d], e = ctx->h_int_lvl_off[d + 1];
        for (int i = b; i < e; i += 32)
            for (int l = 0; l < 32; ++l) {
                int k = i + l;
                sched.push_back(k < e ? (int32_t)(((uint32_t)(2 * k + 1) << 16) | (uint32_t)int_list[k])
                                      : (int32_t)0xffffffffu);
            }
    }
    // the same schedule for dtl_wave_flat2_kernel: byte offsets into a row's M array
    // (element y at index y + 1, 8 bytes each); idle lanes work on dummy slots behind the row
    const int flat_rs = ctx->pitch + 8;
    std::vector<int32_t> sched2(std::max<size_t>(sched.size(), 32 * 32));  // padded to 32 steps with idle words
    for (size_t i = 0; i < sched2.size(); ++i) {
        uint32_t w = i < sched.size() ? (uint32_t)sched[i] : 0xffffffffu;
        if (w == 0xffffffffu) {
            sched2[i] = 0;  // idle lane
            continue;
        }
        uint32_t p_idx = (w & 0xffffu) + 1, c_idx = (w >> 16) + 1;
        sched2[i] = (int32_t)((c_idx * 8) << 16 | (p_idx * 8));
    }
    size_t total = (size_t)S * 7 + (n_int ? n_int : 1) + (size_t)ctx->D + 1 + std::max<size_t>(sched.size(), 32) + sched2.size() +
                   (size_t)S * 4 + 2 * (size_t)ctx->pitch + 24 + (size_t)S * ((ctx->pitch / 32 + 3) / 4 * 4);
    {
        void* ptr = nullptr;
        int rc = sr2_dev_reserve(ctx, DP_SPECIES_SLAB, total * sizeof(int32_t), &ptr);
        if (rc) return rc;
        ctx->d_species_slab = (int32_t*)ptr;
    }
    std::vector<int32_t> slab(total, 0);
    int32_t* h = slab.data();
    int32_t* d = ctx->d_species_slab;
    size_t off = 0;
    auto put = [&](const std::vector<int32_t>& v, size_t n) {
        if (!v.empty()) memcpy(h + off, v.data(), v.size() * sizeof(int32_t));
        const int32_t* p = d + off;
        off += n;
        return p;
    };
    ctx->dsp.S = S;
    ctx->dsp.pitch = ctx->pitch;
    ctx->dsp.D = ctx->D;
    ctx->dsp.n_internal = ctx->n_internal;
    ctx->dsp.child0 = put(ctx->h_child0, S);
    ctx->dsp.depth = put(ctx->h_depth, S);
    ctx->dsp.tin = put(ctx->h_tin, S);
    ctx->dsp.tout = put(ctx->h_tout, S);
    ctx->dsp.parent = put(ctx->h_parent, S);
    ctx->dsp.int_list = put(ctx->h_int_list, n_int ? n_int : 1);
    ctx->dsp.int_lvl_off = put(ctx->h_int_lvl_off, (size_t)ctx->D + 1);
    {
        int rank_bits = 1, depth_bits = 1;
        while ((1 << rank_bits) < S) ++rank_bits;
        while ((1 << depth_bits) < ctx->D) ++depth_bits;
        ctx->dsp.rank_bits = rank_bits;
        ctx->dsp.depth_bits = depth_bits;
        std::vector<int32_t> keylow(S), meta(S);
        for (int y = 0; y < S; ++y) {
            keylow[y] = (int32_t)(((uint32_t)y << depth_bits) | (uint32_t)ctx->h_depth[y]);
            meta[y] = (int32_t)(((uint32_t)(child0[y] + 1) << 16) | (uint32_t)ctx->h_depth[y]);
        }
        ctx->dsp.keylow = (const uint32_t*)put(keylow, S);
        ctx->dsp.meta = (const uint32_t*)put(meta, S);
        ctx->dsp.n_steps = (int)(sched.size() / 32);
        ctx->dsp.sweep_sched = (const uint32_t*)put(sched, std::max<size_t>(sched.size(), 32));
        ctx->dsp.flat_rs = (size_t)flat_rs * 8 * 2 <= 65535 ? flat_rs : 0;  // 16-bit byte offsets
        ctx->dsp.flat_sched = (const uint32_t*)put(sched2, sched2.size());
        // per-species words padded to the row pitch (the flat2 kernel runs over whole pitches)
        std::vector<int32_t> keylow_p(ctx->pitch, 0), cmeta(ctx->pitch, 0);
        for (int y = 0; y < ctx->pitch; ++y) {
            int depth = y < S ? ctx->h_depth[y] : 0;
            int pair_idx = (y < S && child0[y] >= 0) ? child0[y] + 1 : ctx->pitch + 2;  // leaves: the infinite pair
            keylow_p[y] = (int32_t)(((uint32_t)y << depth_bits) | (uint32_t)depth);
            cmeta[y] = (int32_t)(((uint32_t)(pair_idx * 8) << 16) | (uint32_t)depth);
        }
        ctx->dsp.keylow_p = (const uint32_t*)put(keylow_p, ctx->pitch);
        ctx->dsp.cmeta = (const uint32_t*)put(cmeta, ctx->pitch);
        {
            // rows of mask_stride words (a multiple of 4, zero padded): read and written as uint4
            const int mw = ctx->pitch / 32, ms = (mw + 3) / 4 * 4;
            std::vector<int32_t> pathmask((size_t)S * ms, 0);
            for (int y = 0; y < S; ++y) {
                int32_t* row = pathmask.data() + (size_t)y * ms;
                if (y > 0) memcpy(row, pathmask.data() + (size_t)parent[y] * ms, sizeof(int32_t) * ms);
                row[y >> 5] |= (int32_t)(1u << (y & 31));
            }
            ctx->dsp.mask_words = mw;
            ctx->dsp.mask_stride = ms;
            off = (off + 3) / 4 * 4;  // 16-byte alignment (the slab itself is 256-byte aligned)
            ctx->dsp.pathmask = (const uint32_t*)put(pathmask, pathmask.size());
        }
        off = (off + 3) / 4 * 4;  // 16-byte alignment (the slab itself is 256-byte aligned)
        std::vector<int32_t> span(4 * (size_t)S);
        for (int y = 0; y < S; ++y) {
            span[4 * y + 0] = ctx->h_tin[y];
            span[4 * y + 1] = ctx->h_tout[y];
            span[4 * y + 2] = meta[y];
            span[4 * y + 3] = child0[y] >= 0 ? ctx->h_tin[child0[y] + 1] : ctx->h_tout[y];
        }
        ctx->dsp.span4 = (const uint4*)put(span, 4 * (size_t)S);
    }
    SR2_CUDA(ctx, cudaMemcpyAsync(d, h, total * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    // ancestor table for the sparse cherry kernel
    ctx->d_anc_tab = nullptr;
    ctx->dsp.anc_tab = nullptr;
    ctx->dsp.lca_depth = nullptr;
    ctx->dsp.anc_pitch = (ctx->D + 1) & ~1;
    std::vector<uint16_t> anc;
    if ((size_t)S * ctx->dsp.anc_pitch <= ((size_t)1 << 23)) {
        anc.assign((size_t)S * ctx->dsp.anc_pitch, 0);
        for (int y = 0; y < S; ++y) {
            uint16_t* row = anc.data() + (size_t)y * ctx->dsp.anc_pitch;
            if (y > 0) memcpy(row, anc.data() + (size_t)parent[y] * ctx->dsp.anc_pitch, sizeof(uint16_t) * ctx->h_depth[y]);
            row[ctx->h_depth[y]] = (uint16_t)y;
        }
        // depth of lca(a, b) for small species trees (S x S bytes behind the ancestor table): lets the
        // consumers of a cherry row evaluate its closed form themselves (sr2_dtl.cu, "virtual cherries")
        std::vector<uint8_t> lcad;
        if (S <= 1024 && ctx->D <= 255) {
            lcad.resize((size_t)S * S);
            const int ap = ctx->dsp.anc_pitch;
            for (int a = 0; a < S; ++a)
                for (int b = a; b < S; ++b) {
                    const uint16_t *ra = anc.data() + (size_t)a * ap, *rb = anc.data() + (size_t)b * ap;
                    const int dmin = std::min(ctx->h_depth[a], ctx->h_depth[b]);
                    int d = 0;
                    while (d <= dmin && ra[d] == rb[d]) ++d;
                    lcad[(size_t)a * S + b] = lcad[(size_t)b * S + a] = (uint8_t)(d - 1);
                }
        }
        {
            void* ptr = nullptr;
            int rc = sr2_dev_reserve(ctx, DP_SPECIES_ANC, anc.size() * sizeof(uint16_t) + lcad.size(), &ptr);
            if (rc) return rc;
            ctx->d_anc_tab = (uint16_t*)ptr;
        }
        SR2_CUDA(ctx, cudaMemcpyAsync(ctx->d_anc_tab, anc.data(), anc.size() * sizeof(uint16_t), cudaMemcpyHostToDevice,
                                      ctx->stream));
        ctx->dsp.anc_tab = ctx->d_anc_tab;
        if (!lcad.empty()) {
            uint8_t* d_lcad = reinterpret_cast<uint8_t*>(ctx->d_anc_tab + anc.size());
            SR2_CUDA(ctx, cudaMemcpyAsync(d_lcad, lcad.data(), lcad.size(), cudaMemcpyHostToDevice, ctx->stream));
            SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // lcad is a local
            ctx->dsp.lca_depth = d_lcad;
        }
    }
    SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->species_valid = true;
    return SR2_OK;
}
