// sr2_dtl.cu -- the `dtl` (Tofigh-Hallett-Lagergren) reconciliation DP on sm_100a.
//
// Replaces /root/reference/src/superrec2/compute/reconciliation.py:
//   _compute_thl_try_speciation            :60-113   (candidates 1-2 below)
//   _compute_thl_try_duplication_transfer  :116-189  (candidates 3-5 below)
//   _compute_thl_table                     :192-225  (the wavefront driver)
//   _decode_thl_table + reconcile_thl      :228-294  (selection + traceback)
//
// The reference evaluates every (object node u, species x) cell by scanning all
// species (O(|G| |S|^2)).  Here one launch handles one HEIGHT of the object
// trees of a whole batch; each object node is one "row" of |S| cells owned by a
// warp (or by a CTA for large |S|).  Per row the two child rows are staged in
// shared memory as 64-bit keys (value, BFS rank, depth) and two sweeps over the
// BFS-indexed species tree give every cell what it needs in O(1):
//   up-sweep   M[y] = min over subtree(y)            (SUBMIN, reverse BFS levels)
//   down-sweep P[x] = min over species incomparable  (SEPMIN, forward BFS levels)
//              to x  = min(P[parent x], M[sibling x])
// Lexicographic key order == the reference's first-candidate-wins rule
// (utils/dynamic_programming.py:228-238) because candidates are offered in BFS
// order.  The loss term is added AFTER the arg-min, as the reference does
// (reconciliation.py:97-108,153-183), so cell values depend on that tie-break.
//
// What runs by default for a batch whose species tree fits the batched kernels (|S| <= ~480,
// costs that fit 32-bit keys); DESIGN.md section 4 has the measurements behind each piece:
//   wave 1 (cherries)      not launched: readers evaluate the closed form (cherry_form / the per-row constants
//                          of dtl_wave_thread_kernel)
//   side stream            dtl_partition_kernel / dtl_rowmask_kernel: root-path bit set of every row, per wave the
//                          lists of rows by path-set size (bins grouped into the tiers of the thread kernel, and
//                          the dense rows), per row the record the thread kernel starts from; dtl_desc_kernel
//   big waves              dtl_wave_thread_kernel (rows with <= 64 path species: one THREAD per row, tiers of
//                          29 / 42 / 52 / 64 species for config #2, rows stored compact: value | species << 22 and the tag per
//                          path species, nothing for the infinite cells) and dtl_wave_flat2_kernel<NS, false>
//                          (a warp per dense row; rows of up to 3/8 pitch path species stored compact too), on
//                          up to five streams
//   small top waves        dtl_wave_flat2_kernel<NS, true>: one launch, rows chained by done flags
//   selection + traceback  dtl_select_trace_kernel: one CTA per tree
// Everything else (dtl_wave_sparse_kernel -- 4 lanes per sparse row, dense storage: the path when the tables are
// kept for the caller --, dtl_wave_kernel -- a warp or a CTA per row with 64-bit keys, BASELINE config #3 runs on
// it --, dtl_wave_inplace_kernel, dtl_wave_flat_kernel, dtl_cherry_*_kernel, dtl_select_kernel,
// dtl_traceback_kernel, dtl_decoded_cost_kernel) is the general path: large species trees, large
// costs, tables kept for the caller, charged speciations, cross-checks in the tests.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <omp.h>

#include "sr2_rows.cuh"

// ---------------------------------------------------------------------------
// device-resident problem
// ---------------------------------------------------------------------------
#define SPARSE_MAXM_DEFAULT 57
struct DtlCosts;
struct Key32;
struct DtlProblem {
    RowLayout rows;
    uint32_t flags = 0;
    int32_t costs[5] = {0, 0, 0, 0, 0};
    bool need_decoded_cost = false;
    // device
    int32_t* d_T = nullptr;          // [max_chunk_rows * pitch]
    uint32_t* d_tag = nullptr;       // [max_chunk_rows * pitch]
    int32_t* d_C = nullptr;          // decoded cost rows (only if need_decoded_cost)
    int32_t* d_min_cost = nullptr;   // [B]
    int32_t* d_root_species = nullptr;  // [B]
    int32_t* d_map = nullptr;        // [N]
    bool solved = false;
    bool prepared = false;           // buffers reserved, kernels configured
    bool wait_rows = false;          // sr2_dtl_run: row arrays arrive chunk by chunk on the upload stream
    bool two_streams = false;        // pipelined run: chunks alternate between two streams / table sets
    cudaStream_t cur_stream = nullptr;  // the chunk being enqueued: its stream and tables
    int32_t* cur_T = nullptr;
    uint32_t* cur_tag = nullptr;
    int32_t* cur_C = nullptr;
    uint32_t* d_mask = nullptr;      // [rows * pitch/32] root-path bit sets of the rows (flat2 only)
    uint32_t* cur_mask = nullptr;
    // Row lists of every wave (rebuilt by every solve, on a side stream ahead of the wave kernels):
    // d_lists[cls * R + row] = rows of class cls (0 / 1 = the two sparse tiers, 2 = dense) of the wave that starts
    // at `row` (wave-local indices), d_counts[(chunk * 65536 + wave) * 4 + cls] = their lengths.
    int* d_lists = nullptr;
    int* d_counts = nullptr;
    int cur_chunk = 0;
    int cur_set = 0;                 // table set / stream pair of the chunk being enqueued
    bool split_waves = false;        // dense rows of a wave on a second stream, next to the sparse rows
    bool split_tiers = false;        // ... and the second sparse tier on a third one
    cudaEvent_t cur_part_event = nullptr;  // set by dtl_enqueue_chunk for the wave being launched
    int64_t list_pitch = 0;          // = R
    int sparse_maxm = SPARSE_MAXM_DEFAULT;
    int sparse_small = 0;            // rows with at most this many path species take the first tier (0 = one tier)
    bool sparse_thread = false;      // the sparse tiers run on dtl_wave_thread_kernel (a thread per row, compact rows)
    uint8_t* d_rowcnt = nullptr;     // [rows] per table set: entries of a compact row, 0 = dense
    uint8_t* cur_rowcnt = nullptr;
    uint4* d_rowinfo = nullptr;      // [rows][2] per table set: dtl_partition_kernel's record of a listed row
    uint4* cur_rowinfo = nullptr;
    int4* d_desc = nullptr;          // [rows] per table set: (left child, right child, node of the left child, of the right child),
    int4* cur_desc = nullptr;        // the four row arrays packed for dtl_select_trace_kernel (dtl_desc_kernel)
    bool cur_compact = false;        // chunk being enqueued: its sparse rows are stored compact
    // classes of dtl_partition_kernel: table [0] without compact rows (two sparse tiers + dense), table [1] with
    // (bins of path-set sizes grouped into the tiers of dtl_wave_thread_kernel + dense); device copies of 1024 bytes
    uint8_t* d_class_of = nullptr;
    std::vector<uint8_t> class_table;   // host copy of what d_class_of holds
    int n_classes[2] = {3, 3};
    int n_tiers = 0;                 // dtl_wave_thread_kernel launches per wave
    int compact_max = 0;             // listed rows with at most this many path species are stored compact
    bool decouple_dense = true;      // the dense rows of a wave are not waited for by the sparse rows of the next
    cudaEvent_t pending_dense = nullptr;   // last dense launch the chunk's stream has not joined yet
    int tier_maxm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int tier_cls0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // first class of a tier; [n_tiers] = the dense class
    int n_waves = 0, n_launch = 0;
    // wave launch configuration (fixed per upload)
    int spad = 0, wave_warps = 0;
    size_t wave_smem = 0;
    bool cta_per_row = false;
    int wave_threads = 512;          // CTA size of the one-CTA-per-row kernel
    bool wave_inplace = false;       // the two-row form of the CTA-per-row kernel (very large species trees)
    // batched wave kernels (32-bit keys)
    bool simd_ok = false;
    int rank_bits = 0, depth_bits = 0;
    int64_t max_finite = 0;
    int64_t simd_min_rows = 0;
    int simd_thresh = 0;
    bool flat = true;
    int flat_rs = 0;
    size_t flat_smem = 0;
    bool flat2 = false;          // the default batched kernel when its preconditions hold
    size_t flat2_smem = 0;
    int flat2_rows = 8;
    int flat2_low_waves = 5;     // waves up to this height launch an eighth of the dense-list CTAs (measured: 0 / 4 / 6 / 99 -> 6.70 / 6.59 / 6.59 / 6.66 ms)
    int flat2_sh = 0, flat2_thresh = 0;
    void (*flat2_fn)(DevSpecies, const int32_t*, const int32_t*, int64_t, int64_t, int, int32_t*, uint32_t*, DtlCosts,
                     Key32, const uint32_t*, int, int, int, const int*, const int*, int*, int, const uint8_t*, int) = nullptr;
    // The small top waves of a chunk run as ONE launch (flat2, chained): a row starts when the "done"
    // flags of its child rows are up, instead of one latency-bound launch per wave.
    void (*flat2_chain_fn)(DevSpecies, const int32_t*, const int32_t*, int64_t, int64_t, int, int32_t*, uint32_t*, DtlCosts,
                           Key32, const uint32_t*, int, int, int, const int*, const int*, int*, int, const uint8_t*, int) = nullptr;
    int cur_cherry_end = 0;          // chunk being enqueued: its wave-1 rows [0, cherry_end) are virtual (0 = materialised)
    size_t trace_smem_set = 0;       // dynamic shared memory opted in for dtl_select_trace_kernel
    int64_t chain_rows = 0;          // waves with fewer rows than this join the chained launch (0 = off)
    int* d_chain_sync = nullptr;     // per table set: [0] CTA ticket, [1 + i] done flag of chained row i
    int flat2_pf_near = 0, flat2_pf_far = 0, sparse_pf_near = 0, sparse_pf_far = 0;
    int flat2_sparse_max = 0;        // rows with at most this many root-path species take the sparse cells path
};

void sr2_free_dtl(sr2_ctx* ctx) {
    DtlProblem* p = ctx->dtl;
    if (!p) return;
    delete p;  // device buffers live in the context's pools
    ctx->dtl = nullptr;
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
struct DtlCosts {
    int spe, dup, hgt, loss;
};

template <int TPR>
__device__ __forceinline__ void group_sync() {
    if (TPR == 32)
        __syncwarp();
    else
        __syncthreads();
}

// One candidate of a cell: children placed at the species held by keys a (left
// child) and b (right child); `extra` is event cost + loss * (sum of distances).
__device__ __forceinline__ void dtl_offer(u64 a, u64 b, int extra, int& best, unsigned& tag) {
    unsigned va = (unsigned)key_value(a), vb = (unsigned)key_value(b);
    if (va < (unsigned)SR2_INF && vb < (unsigned)SR2_INF) {
        int v = (int)(va + vb) + extra;
        if (v < best) {
            best = v;
            tag = (unsigned)key_rank(a) | ((unsigned)key_rank(b) << 16);
        }
    }
}

// Cell (u, x).  Ml/Mr: SUBMIN keys at x; Pl/Pr: SEPMIN keys at x (unused for the
// root); Sl/Sr: slot arrays, still holding SUBMIN at the children of x.
// Candidate order = reference update order (reconciliation.py:110-113,185-189).
__device__ __forceinline__ void dtl_cell(int x, int dx, int cx, bool is_root, u64 Ml, u64 Mr,
                                         u64 Pl, u64 Pr, const u64* Sl, const u64* Sr,
                                         const DtlCosts& cs, int& best, unsigned& tag) {
    best = SR2_INF;
    tag = SR2_NO_TAG;
    if (cx >= 0) {
        u64 l0 = Sl[cx], l1 = Sl[cx + 1], r0 = Sr[cx], r1 = Sr[cx + 1];
        // 1. speciation, left child below x's first child, right below the second
        dtl_offer(l0, r1, cs.loss * (key_depth(l0) + key_depth(r1) - 2 * dx - 2), best, tag);
        // 2. speciation, the other way round
        dtl_offer(l1, r0, cs.loss * (key_depth(l1) + key_depth(r0) - 2 * dx - 2), best, tag);
    }
    // 3. duplication: both children anywhere below (or at) x
    dtl_offer(Ml, Mr, cs.dup + cs.loss * (key_depth(Ml) + key_depth(Mr) - 2 * dx), best, tag);
    if (!is_root) {
        // 4. transfer of the left child to a species incomparable to x
        dtl_offer(Pl, Mr, cs.hgt + cs.loss * (key_depth(Mr) - dx), best, tag);
        // 5. transfer of the right child
        dtl_offer(Ml, Pr, cs.hgt + cs.loss * (key_depth(Ml) - dx), best, tag);
    }
}

// Species topology staged once per CTA behind the key rows (see superdtl_wave_kernel): per-species
// (child0 + 1) << 16 | depth, level offsets of the internal species, their level-order ranks.
__host__ __device__ __forceinline__ size_t dtl_topology_bytes(int S, int D, int n_int) {
    return ((size_t)S + (size_t)D + 1) * 4 + (((size_t)n_int + 1) & ~(size_t)1) * 2;
}

// TPR = threads per row: 32 (one warp per row, several rows per CTA) or the CTA size (one row per
// CTA, for species trees too large for a warp's share of shared memory: BASELINE config #3).
// Four key rows per object node -- SUBMIN of the left / right child (Ml, Mr), SEPMIN (Pl, Pr) -- and
// three phases: up-sweep of Ml / Mr, top-down sweep filling Pl / Pr, then the cells element-wise.
// The sweeps take one step per species level; their work items are (internal species, key row) so
// that narrow levels still use the lanes, and nothing in the level loops touches global memory
// (ncu r3b on the superdtl twin of this kernel: the per-level __ldg of the first form cost more
// than the arithmetic).
// CHAIN (one CTA per row only): every wave from the second one on in ONE launch.  CTAs take rows by ticket
// (row order = height order), a CTA waits until the "done" flags of its two child rows are up, reads them
// with L2-only loads and raises its own flag behind a fence.  A row can only wait for rows with smaller
// tickets, which already run, so the launch cannot deadlock; and a row starts as soon as ITS children are
// done instead of when the whole wave below is (a single instance has a handful of rows per top wave).
template <int TPR, bool CHAIN = false>
__global__ void __launch_bounds__(TPR == 32 ? 256 : TPR)
dtl_wave_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl,
                const int32_t* __restrict__ row_cr, int64_t row_begin, int64_t chunk_row0,
                int n_rows, int32_t* __restrict__ T, uint32_t* __restrict__ tags, DtlCosts cs,
                int spad, int* __restrict__ chain_sync = nullptr) {
    extern __shared__ u64 smem[];
    __shared__ int s_ticket;
    const int S = sp.S;
    int group, lane, groups_per_cta;
    if (TPR == 32) {
        group = threadIdx.x >> 5;
        lane = threadIdx.x & 31;
        groups_per_cta = blockDim.x >> 5;
    } else {
        group = 0;
        lane = threadIdx.x;
        groups_per_cta = 1;
    }
    uint32_t* s_meta = reinterpret_cast<uint32_t*>(smem + (size_t)groups_per_cta * 4 * spad);
    int32_t* s_lvl = reinterpret_cast<int32_t*>(s_meta + S);
    uint16_t* s_int = reinterpret_cast<uint16_t*>(s_lvl + sp.D + 1);
    for (int i = threadIdx.x; i < S; i += blockDim.x) s_meta[i] = __ldg(sp.meta + i);
    for (int i = threadIdx.x; i <= sp.D; i += blockDim.x) s_lvl[i] = __ldg(sp.int_lvl_off + i);
    for (int i = threadIdx.x; i < sp.n_internal; i += blockDim.x) s_int[i] = (uint16_t)__ldg(sp.int_list + i);
    if (CHAIN && threadIdx.x == 0) s_ticket = atomicAdd(chain_sync, 1);
    __syncthreads();
    int local = CHAIN ? s_ticket : blockIdx.x * groups_per_cta + group;
    if (local >= n_rows) return;  // uniform per group
    const int64_t gr = row_begin + local;              // global row
    const size_t slot = (size_t)(gr - chunk_row0);     // row inside the chunk's tables
    u64* base = smem + (size_t)group * 4 * spad;       // Ml, Mr, Pl, Pr
    u64* Ml = base;
    u64* Mr = base + spad;
    const int cl = row_cl[gr], cr = row_cr[gr];
    SR2_DCHECK(cl < (int)slot && cr < (int)slot && cl >= -S && cr >= -S);
    if (CHAIN) {
        const int first_slot = (int)(row_begin - chunk_row0);
        const int c = threadIdx.x == 0 ? cl : cr;
        if (threadIdx.x < 2 && c >= first_slot) {
            if (c - first_slot >= local) __trap();   // rows are sorted by height: a child never follows its parent
            const int* flag = chain_sync + 1 + (c - first_slot);
            int up;
            for (unsigned spins = 0;; ++spins) {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(up) : "l"(flag) : "memory");
                if (up) break;
                if (spins > (1u << 24)) __trap();
                __nanosleep(100);
            }
        }
        __syncthreads();
    }
    const int32_t* __restrict__ Tl = cl >= 0 ? T + (size_t)cl * sp.pitch : nullptr;
    const int32_t* __restrict__ Tr = cr >= 0 ? T + (size_t)cr * sp.pitch : nullptr;

    // stage the two child rows as keys (CHAIN: written by other CTAs of this launch, so past L1)
    for (int y = lane; y < S; y += TPR) {
        const int d = (int)(s_meta[y] & 0xffffu);
        const int vl = Tl ? (CHAIN ? __ldcg(Tl + y) : Tl[y]) : (y == -1 - cl ? 0 : SR2_INF);
        const int vr = Tr ? (CHAIN ? __ldcg(Tr + y) : Tr[y]) : (y == -1 - cr ? 0 : SR2_INF);
        Ml[y] = key_make(vl, y, d);
        Mr[y] = key_make(vr, y, d);
    }
    group_sync<TPR>();

    // 1. up-sweep: M[y] <- min over subtree(y); the deepest level has leaves only
    for (int d = sp.D - 2; d >= 0; --d) {
        const int b = s_lvl[d], items = (s_lvl[d + 1] - b) * 2;
        for (int it = lane; it < items; it += TPR) {
            const int i = b + (it >> 1);
            const int p = s_int[i], c = 2 * i + 1;
            u64* arr = base + (size_t)(it & 1) * spad;
            arr[p] = key_min(arr[p], key_min(arr[c], arr[c + 1]));
        }
        group_sync<TPR>();
    }

    // 2. SEPMIN: nothing is incomparable to the root; P[child] = min(P[parent], M[sibling])
    if (lane < 2) base[(size_t)(2 + lane) * spad] = SR2_KEY_NONE;
    group_sync<TPR>();
    for (int d = 0; d < sp.D - 1; ++d) {
        const int b = s_lvl[d], items = (s_lvl[d + 1] - b) * 2;
        for (int it = lane; it < items; it += TPR) {
            const int i = b + (it >> 1);
            const int p = s_int[i], c = 2 * i + 1;
            const u64* M = base + (size_t)(it & 1) * spad;
            u64* P = base + (size_t)(2 + (it & 1)) * spad;
            const u64 Pp = P[p];
            P[c] = key_min(Pp, M[c + 1]);
            P[c + 1] = key_min(Pp, M[c]);
        }
        group_sync<TPR>();
    }

    // 3. the cells
    int32_t* __restrict__ Tout = T + slot * sp.pitch;
    uint32_t* __restrict__ Gout = tags + slot * sp.pitch;
    const u64* Pl = base + (size_t)2 * spad;
    const u64* Pr = base + (size_t)3 * spad;
    for (int x = lane; x < S; x += TPR) {
        const uint32_t m = s_meta[x];
        int best;
        unsigned tag;
        dtl_cell(x, (int)(m & 0xffffu), (int)(m >> 16) - 1, x == 0, Ml[x], Mr[x], Pl[x], Pr[x], Ml, Mr, cs, best,
                 tag);
        Tout[x] = best;
        Gout[x] = tag;
    }
    if (CHAIN) {   // the row is complete: every thread's stores, then the flag
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            int* flag = chain_sync + 1 + local;
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
        }
    }
}


// The first form of the kernel above, kept for species trees whose four key rows do not fit one CTA's
// shared memory (more than ~5,900 species nodes): two key rows, SEPMIN written in place over SUBMIN and
// the cells computed inside the top-down sweep.
template <int TPR>
__global__ void __launch_bounds__(TPR == 32 ? 256 : TPR)
dtl_wave_inplace_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl,
                const int32_t* __restrict__ row_cr, int64_t row_begin, int64_t chunk_row0,
                int n_rows, int32_t* __restrict__ T, uint32_t* __restrict__ tags, DtlCosts cs,
                int spad) {
    extern __shared__ u64 smem[];
    const int S = sp.S;
    int group, lane, groups_per_cta;
    if (TPR == 32) {
        group = threadIdx.x >> 5;
        lane = threadIdx.x & 31;
        groups_per_cta = blockDim.x >> 5;
    } else {
        group = 0;
        lane = threadIdx.x;
        groups_per_cta = 1;
    }
    int local = blockIdx.x * groups_per_cta + group;
    if (local >= n_rows) return;  // uniform per group
    const int64_t gr = row_begin + local;              // global row
    const size_t slot = (size_t)(gr - chunk_row0);     // row inside the chunk's tables
    u64* Sl = smem + (size_t)group * 2 * spad;
    u64* Sr = Sl + spad;
    const int cl = row_cl[gr], cr = row_cr[gr];
    const int32_t* __restrict__ Tl = cl >= 0 ? T + (size_t)cl * sp.pitch : nullptr;
    const int32_t* __restrict__ Tr = cr >= 0 ? T + (size_t)cr * sp.pitch : nullptr;

    // stage the two child rows as keys
    for (int y = lane; y < S; y += TPR) {
        int d = __ldg(sp.depth + y);
        int vl = Tl ? Tl[y] : (y == -1 - cl ? 0 : SR2_INF);
        int vr = Tr ? Tr[y] : (y == -1 - cr ? 0 : SR2_INF);
        Sl[y] = key_make(vl, y, d);
        Sr[y] = key_make(vr, y, d);
    }
    group_sync<TPR>();

    // up-sweep: slot[y] <- min over subtree(y); deepest level has leaves only
    for (int d = sp.D - 2; d >= 0; --d) {
        int b = __ldg(sp.int_lvl_off + d), e = __ldg(sp.int_lvl_off + d + 1);
        for (int i = b + lane; i < e; i += TPR) {
            int p = __ldg(sp.int_list + i), c = 2 * i + 1;
            Sl[p] = key_min(Sl[p], key_min(Sl[c], Sl[c + 1]));
            Sr[p] = key_min(Sr[p], key_min(Sr[c], Sr[c + 1]));
        }
        group_sync<TPR>();
    }

    int32_t* __restrict__ Tout = T + slot * sp.pitch;
    uint32_t* __restrict__ Gout = tags + slot * sp.pitch;

    // root cell; afterwards slot[0] holds P[root] = "nothing is incomparable to the root"
    if (lane == 0) {
        int best;
        unsigned tag;
        dtl_cell(0, 0, __ldg(sp.child0), true, Sl[0], Sr[0], SR2_KEY_NONE, SR2_KEY_NONE, Sl, Sr, cs,
                 best, tag);
        Tout[0] = best;
        Gout[0] = tag;
        Sl[0] = SR2_KEY_NONE;
        Sr[0] = SR2_KEY_NONE;
    }
    group_sync<TPR>();

    // down-sweep: each internal species p finishes the cells of its two children
    // and leaves their P keys in their slots for the next level.
    for (int d = 0; d < sp.D - 1; ++d) {
        int b = __ldg(sp.int_lvl_off + d), e = __ldg(sp.int_lvl_off + d + 1);
        for (int i = b + lane; i < e; i += TPR) {
            int p = __ldg(sp.int_list + i), c = 2 * i + 1;
            u64 Ppl = Sl[p], Ppr = Sr[p];
            u64 m0l = Sl[c], m1l = Sl[c + 1], m0r = Sr[c], m1r = Sr[c + 1];
            u64 P0l = key_min(Ppl, m1l), P1l = key_min(Ppl, m0l);
            u64 P0r = key_min(Ppr, m1r), P1r = key_min(Ppr, m0r);
            int best0, best1;
            unsigned tag0, tag1;
            dtl_cell(c, d + 1, __ldg(sp.child0 + c), false, m0l, m0r, P0l, P0r, Sl, Sr, cs, best0, tag0);
            dtl_cell(c + 1, d + 1, __ldg(sp.child0 + c + 1), false, m1l, m1r, P1l, P1r, Sl, Sr, cs, best1,
                     tag1);
            Tout[c] = best0;
            Tout[c + 1] = best1;
            Gout[c] = tag0;
            Gout[c + 1] = tag1;
            Sl[c] = P0l;
            Sl[c + 1] = P1l;
            Sr[c] = P0r;
            Sr[c + 1] = P1r;
        }
        group_sync<TPR>();
    }
}


// ---------------------------------------------------------------------------
// Wave 1: rows whose two children are leaves ("cherries").  Every cell has a closed
// form: with the children at species a and b, each SUBMIN/SEPMIN key is either
// (0, a|b) or infinite, so the five candidates reduce to interval tests.  One warp
// per row, lanes over species, fully coalesced stores.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dtl_cherry_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                  int64_t row_begin, int64_t chunk_row0, int n_rows, int32_t* __restrict__ T,
                  uint32_t* __restrict__ tags, DtlCosts cs, uint32_t* __restrict__ rowmask) {
    int local = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (local >= n_rows) return;
    const int64_t gr = row_begin + local;
    const size_t slot = (size_t)(gr - chunk_row0);
    const int a = -1 - row_cl[gr], b = -1 - row_cr[gr];
    const int ta = __ldg(sp.tin + a), oa = __ldg(sp.tout + a), da = __ldg(sp.depth + a);
    const int tb = __ldg(sp.tin + b), ob = __ldg(sp.tout + b), db = __ldg(sp.depth + b);
    int32_t* __restrict__ Tout = T + slot * sp.pitch;
    if (rowmask && lane < sp.mask_words)
        rowmask[slot * sp.mask_stride + lane] = __ldg(sp.pathmask + (size_t)a * sp.mask_stride + lane) |
                                                __ldg(sp.pathmask + (size_t)b * sp.mask_stride + lane);
    for (int x = lane; x < sp.S; x += 32) {
        const uint4 sx = __ldg(sp.span4 + x);
        const int tx = (int)sx.x, ox = (int)sx.y, dx = (int)(sx.z & 0xffffu), t1 = (int)sx.w;
        const bool internal = (sx.z >> 16) != 0;
        bool in_a = tx <= ta && ta < ox, in_b = tx <= tb && tb < ox;
        int best = SR2_INF;
        // candidates 1 / 2: subtree(first child) = [tx+1, t1), subtree(second child) = [t1, ox)
        if (internal && in_a && in_b && ta > tx && tb > tx && ((ta < t1) != (tb < t1)))
            best = cs.loss * (da + db - 2 * dx - 2);
        if (in_a && in_b) {  // 3. duplication
            int v = cs.dup + cs.loss * (da + db - 2 * dx);
            if (v < best) best = v;
        }
        bool sep_a = !in_a && !(ta <= tx && tx < oa), sep_b = !in_b && !(tb <= tx && tx < ob);
        if (sep_a && in_b) {  // 4. left child transferred
            int v = cs.hgt + cs.loss * (db - dx);
            if (v < best) best = v;
        }
        if (in_a && sep_b) {  // 5. right child transferred
            int v = cs.hgt + cs.loss * (da - dx);
            if (v < best) best = v;
        }
        Tout[x] = best;
        // no tag row: every finite cell of a cherry has the tag (a, b), which its readers rebuild
        // from the child descriptors (dtl_traceback_kernel, dtl_decoded_cost_kernel, sr2_dtl_tables)
    }
}

// Wave 1, sparse form.  A cherry's row is infinite except on the two root paths of its leaf
// species a and b (at most 2 x depth cells of |S|):
//   x above both (depth d <= depth of lca(a, b)):  duplication  dup + loss (da + db - 2d),
//                                                   at the lca itself (if neither a nor b) the speciation
//                                                   loss (da + db - 2d - 2), offered first
//   x above a only:  right child transferred        hgt + loss (da - d)
//   x above b only:  left child transferred         hgt + loss (db - d)
// (the same values as dtl_cherry_kernel's interval tests, which are kept for species trees whose
// ancestor table would be too large).  One warp per row: 128-bit stores of the infinite row, then
// the path cells from the ancestor table.  HBM-write-bound instead of ALU-bound.
__global__ void __launch_bounds__(256)
dtl_cherry_sparse_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                         int64_t row_begin, int64_t chunk_row0, int n_rows, int32_t* __restrict__ T, DtlCosts cs,
                         uint32_t* __restrict__ rowmask) {
    int local = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (local >= n_rows) return;
    const int64_t gr = row_begin + local;
    const size_t slot = (size_t)(gr - chunk_row0);
    const int a = -1 - row_cl[gr], b = -1 - row_cr[gr];
    const int da = __ldg(sp.depth + a), db = __ldg(sp.depth + b);
    int32_t* __restrict__ Tout = T + slot * sp.pitch;
    uint4* __restrict__ row4 = reinterpret_cast<uint4*>(Tout);
    const uint4 inf4 = make_uint4(SR2_INF, SR2_INF, SR2_INF, SR2_INF);
    for (int i = lane; i < (sp.pitch >> 2); i += 32) row4[i] = inf4;
    if (rowmask && lane < sp.mask_words)
        rowmask[slot * sp.mask_stride + lane] = __ldg(sp.pathmask + (size_t)a * sp.mask_stride + lane) |
                                                __ldg(sp.pathmask + (size_t)b * sp.mask_stride + lane);
    const uint16_t* __restrict__ pa = sp.anc_tab + (size_t)a * sp.anc_pitch;
    const uint16_t* __restrict__ pb = sp.anc_tab + (size_t)b * sp.anc_pitch;
    // depths at which the two root paths coincide: 0 .. common - 1
    int common = 0;
    const int dmin = min(da, db);
    for (int d0 = 0; d0 <= dmin; d0 += 32) {
        const int d = d0 + lane;
        const bool eq = d <= dmin && pa[d] == pb[d];
        common += __popc(__ballot_sync(0xffffffffu, eq));
    }
    __syncwarp();  // the infinite row is written before its finite cells
    // dl = depth of lca(a, b).  "Leaf species" may be internal nodes of the species tree (the
    // reference maps leaves to any named node), so one of a, b can be the lca itself: then there is
    // no speciation (both children must sit strictly below x) and no transfer below the lca on the
    // other one's path (the species left behind would be an ancestor of x, not separated from it).
    const int dl = common - 1;
    const bool a_below = da > dl, b_below = db > dl;
    for (int d = lane; d <= da; d += 32) {
        int v;
        if (d <= dl)
            v = (d == dl && a_below && b_below) ? cs.loss * (da + db - 2 * d - 2) : cs.dup + cs.loss * (da + db - 2 * d);
        else if (b_below)
            v = cs.hgt + cs.loss * (da - d);
        else
            continue;
        Tout[pa[d]] = v;
    }
    if (a_below)
        for (int d = common + lane; d <= db; d += 32) Tout[pb[d]] = cs.hgt + cs.loss * (db - d);
}

// ---------------------------------------------------------------------------
// 32-bit arg-min keys of the batched wave kernels:
//   value[31:sh] | rank[sh-1:db] | depth[db-1:0]
// which the host only selects when every finite table value fits the value field.
// ---------------------------------------------------------------------------
// compact rows (dtl_wave_thread_kernel): entry = value | species << TH_VAL_BITS, TH_VAL_MASK = infinite
#define TH_VAL_BITS 22
#define TH_VAL_MASK 0x3fffffu
struct Key32 {
    int sh, db;            // value shift (>= rank bits + depth bits), depth bits
    unsigned rm, dm;       // rank mask (after >> db), depth mask
    unsigned inf32;        // the infinite value: top bit of the value field
    int thresh;            // sums below this are finite (inf32 minus the largest correction)
};

// ---------------------------------------------------------------------------
// "Virtual cherries".  The row of a cherry (both children leaves, at species a and b) has a closed
// form (dtl_cherry_sparse_kernel), so the batched kernels that read a cherry child evaluate it
// themselves instead of reading a row from HBM, and wave 1 is not launched at all: no 1.6 KB row
// written per cherry, no gather from it.  A child row slot below `cherry_end` (the chunk's wave-1
// rows come first) is a cherry; its leaf species are its own descriptors.
// ---------------------------------------------------------------------------
struct CherryForm {
    const uint16_t* pa;  // root paths of a and b by depth
    const uint16_t* pb;
    int da, db, dl;      // depths of a, b and lca(a, b)
    bool a_below, b_below;
};
__device__ __forceinline__ CherryForm cherry_form(const DevSpecies& sp, int a, int b) {
    CherryForm f;
    f.pa = sp.anc_tab + (size_t)a * sp.anc_pitch;
    f.pb = sp.anc_tab + (size_t)b * sp.anc_pitch;
    f.da = __ldg(sp.depth + a);
    f.db = __ldg(sp.depth + b);
    f.dl = (int)__ldg(sp.lca_depth + (size_t)a * sp.S + b);
    f.a_below = f.da > f.dl;
    f.b_below = f.db > f.dl;
    return f;
}
// The finite cells of the cherry sit on the two root paths: value of the cell at the depth-d species of
// a's path (which = 0, d <= da) or of b's path below the lca (which = 1, dl < d <= db); 0xffffffff =
// no candidate (see dtl_cherry_sparse_kernel for the cases).  Readers walk the paths by depth, so the
// ancestor-table loads of a lane are independent of each other.
__device__ __forceinline__ unsigned cherry_path_value(const CherryForm& f, int which, int d, const DtlCosts& cs) {
    if (which == 0) {
        if (d <= f.dl)
            return (unsigned)((d == f.dl && f.a_below && f.b_below) ? cs.loss * (f.da + f.db - 2 * d - 2)
                                                                     : cs.dup + cs.loss * (f.da + f.db - 2 * d));
        return f.b_below ? (unsigned)(cs.hgt + cs.loss * (f.da - d)) : 0xffffffffu;
    }
    return (d > f.dl && f.a_below) ? (unsigned)(cs.hgt + cs.loss * (f.db - d)) : 0xffffffffu;
}

// ---------------------------------------------------------------------------
// Batched waves, first form ("flat", kept as a cross-check of flat2): one warp per row, 32-bit keys, and BOTH sweeps
// materialised in shared memory before any cell is evaluated:
//   ML/MR = SUBMIN of the two child rows (up-sweep, in place)
//   PL/PR = SEPMIN of the two child rows (top-down sweep into two more arrays)
// The cells are then an element-wise pass with lanes over consecutive species: every lane
// is busy, child-row loads and the T / tag stores are coalesced, and the level structure
// of the species tree only costs the two cheap sweeps.  Arrays are shifted by one word so
// that the keys of the two children of a species (ranks 2i+1, 2i+2) are one aligned
// 64-bit shared-memory load.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dtl_wave_flat_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                     int64_t row_begin, int64_t chunk_row0, int n_rows, int32_t* __restrict__ T,
                     uint32_t* __restrict__ tags, DtlCosts cs, int rs, Key32 kp) {
    extern __shared__ uint2 smem64[];
    const int S = sp.S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int local = blockIdx.x * nwarps + warp;
    if (local >= n_rows) return;
    const int64_t gr = row_begin + local;
    const size_t slot = (size_t)(gr - chunk_row0);
    // Per row two arrays of (left-child key, right-child key) pairs: M (SUBMIN) and P (SEPMIN).
    // Element y lives at index y + 1, so the pair of children (2i+1, 2i+2) of a species is one
    // aligned 16-byte word.
    uint2* M = smem64 + (size_t)warp * 2 * rs + 1;
    uint2* P = M + rs;
    const int cl = row_cl[gr], cr = row_cr[gr];
    const unsigned inf32 = kp.inf32, vmul = 1u << kp.sh;

    // ---- load both child rows as keys: key = min(value, inf) * 2^sh + (rank, depth) ------------
    // All global loads of a 128-species chunk (both children + the shared low key words, up to
    // 12 per lane) are issued before any is consumed; a leaf child's row (0 at one species,
    // infinite elsewhere) is synthesised without touching memory.
    {
        const int32_t* __restrict__ Tl = T + (size_t)(cl >= 0 ? cl : 0) * sp.pitch;
        const int32_t* __restrict__ Tr = T + (size_t)(cr >= 0 ? cr : 0) * sp.pitch;
        const bool il = cl >= 0, ir = cr >= 0;
        const int al = -1 - cl, ar = -1 - cr;
#pragma unroll 1
        for (int y0 = 0; y0 < S; y0 += 128) {
            unsigned vl[4], vr[4], low[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int y = y0 + k * 32 + lane;
                const bool ok = y < S;
                low[k] = ok ? __ldg(sp.keylow + y) : 0u;
                vl[k] = (ok && il) ? (unsigned)Tl[y] : (y == al ? 0u : inf32);
                vr[k] = (ok && ir) ? (unsigned)Tr[y] : (y == ar ? 0u : inf32);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int y = y0 + k * 32 + lane;
                if (y < S) M[y] = make_uint2(min(vl[k], inf32) * vmul + low[k], min(vr[k], inf32) * vmul + low[k]);
            }
        }
    }
    __syncwarp();

    // ---- up-sweep: SUBMIN in place, following the precomputed schedule bottom-up; the next
    //      step's schedule word is fetched while the current step runs -----------------------------
    {
        const uint32_t* __restrict__ sched = sp.sweep_sched + lane;
        unsigned w = sp.n_steps > 0 ? __ldg(sched + (sp.n_steps - 1) * 32) : 0xffffffffu;
#pragma unroll 1
        for (int step = sp.n_steps - 1; step >= 0; --step) {
            const unsigned next = step > 0 ? __ldg(sched + (step - 1) * 32) : 0xffffffffu;
            if (w != 0xffffffffu) {
                const int p = (int)(w & 0xffffu), c = (int)(w >> 16);
                const uint4 k = *reinterpret_cast<const uint4*>(M + c);  // (l,r) of c and of c+1
                const uint2 own = M[p];
                M[p] = make_uint2(min(own.x, min(k.x, k.z)), min(own.y, min(k.y, k.w)));
            }
            w = next;
            __syncwarp();
        }
    }

    // ---- top-down sweep: SEPMIN into P ------------------------------------------------------------
    if (lane == 0) P[0] = make_uint2(0xffffffffu, 0xffffffffu);  // nothing is incomparable to the root
    __syncwarp();
    {
        const uint32_t* __restrict__ sched = sp.sweep_sched + lane;
        unsigned w = sp.n_steps > 0 ? __ldg(sched) : 0xffffffffu;
#pragma unroll 1
        for (int step = 0; step < sp.n_steps; ++step) {
            const unsigned next = step + 1 < sp.n_steps ? __ldg(sched + (step + 1) * 32) : 0xffffffffu;
            if (w != 0xffffffffu) {
                const int p = (int)(w & 0xffffu), c = (int)(w >> 16);
                const uint4 k = *reinterpret_cast<const uint4*>(M + c);
                const uint2 pp = P[p];
                // incomparable(c) = incomparable(p) + subtree(c+1), and the other way round
                *reinterpret_cast<uint4*>(P + c) =
                    make_uint4(min(pp.x, k.z), min(pp.y, k.w), min(pp.x, k.x), min(pp.y, k.y));
            }
            w = next;
            __syncwarp();
        }
    }

    // ---- cells: element-wise over species -------------------------------------------------------------
    int32_t* __restrict__ Tout = T + slot * sp.pitch;
    uint32_t* __restrict__ Gout = tags + slot * sp.pitch;
    const int loss = cs.loss, NONE = 0x7fffffff;
    const int sh = kp.sh, db = kp.db, thresh = kp.thresh;
    const unsigned dm = kp.dm, rm = kp.rm;
#pragma unroll 1
    for (int x = lane; x < S; x += 32) {
        const unsigned meta = __ldg(sp.meta + x);
        const int dx = (int)(meta & 0xffffu), cx = (int)(meta >> 16) - 1;
        const uint2 m = M[x], pq = P[x];
        const int qMl = (int)(m.x >> sh) + loss * (int)(m.x & dm), qMr = (int)(m.y >> sh) + loss * (int)(m.y & dm);
        int w1 = NONE, w2 = NONE, w4 = NONE, w5 = NONE;
        uint4 k = make_uint4(0, 0, 0, 0);  // (l0, r0, l1, r1): keys of the two children of x
        if (cx >= 0) {
            k = *reinterpret_cast<const uint4*>(M + cx);
            const int b12 = -2 * loss * (dx + 1);
            const int ql0 = (int)(k.x >> sh) + loss * (int)(k.x & dm), qr0 = (int)(k.y >> sh) + loss * (int)(k.y & dm);
            const int ql1 = (int)(k.z >> sh) + loss * (int)(k.z & dm), qr1 = (int)(k.w >> sh) + loss * (int)(k.w & dm);
            w1 = ((ql0 + qr1 + b12) << 3) | 0;  // 1. speciation, left child below x0, right below x1
            w2 = ((ql1 + qr0 + b12) << 3) | 1;  // 2. speciation, the other way round
        }
        const int w3 = ((qMl + qMr + cs.dup - 2 * loss * dx) << 3) | 2;  // 3. duplication
        if (x != 0) {
            w4 = (((int)(pq.x >> sh) + qMr + cs.hgt - loss * dx) << 3) | 3;  // 4. transfer, left child away
            w5 = ((qMl + (int)(pq.y >> sh) + cs.hgt - loss * dx) << 3) | 4;  // 5. transfer, right child away
        }
        const int w = min(min(min(w1, w2), w3), min(w4, w5));  // smallest value, then first candidate
        const int best = w >> 3, which = w & 7;
        unsigned ka = m.x, kb = m.y;
        ka = which == 0 ? k.x : ka;
        ka = which == 1 ? k.z : ka;
        ka = which == 3 ? pq.x : ka;
        kb = which == 0 ? k.w : kb;
        kb = which == 1 ? k.y : kb;
        kb = which == 4 ? pq.y : kb;
        const bool finite = best < thresh;
        Tout[x] = finite ? best : SR2_INF;
        Gout[x] = finite ? (((ka >> db) & rm) | (((kb >> db) & rm) << 16)) : SR2_NO_TAG;
    }
}

// ---------------------------------------------------------------------------
// Batched waves, the form in use ("flat2"): the flat kernel with its three phases put on an
// instruction diet (the flat kernel is issue-bound, profiles/README.md r1j):
//   * everything runs over whole row pitches (multiples of 32 species) -- no bounds tests;
//     pad cells hold garbage that nothing reads;
//   * child rows become keys with ONE multiply-add: value * 2^sh + (rank, depth).  The stored
//     infinity 0x3fffffff wraps to an all-ones value field, which is the infinite key;
//   * the sweep schedule (DevSpecies::flat_sched) lives in registers as absolute shared-memory
//     addresses (child pair << 16 | parent, CTA shared memory < 64 KB), so a sweep step is
//     2 address ops + 2 loads + min + store; idle lanes work on dummy slots, not on predicates;
//   * cells: per-species words come from shared memory; leaves point their "child pair" at an
//     always-infinite pair and the root's SEPMIN is the infinite key, so there is no branch.
// Row layout in shared memory: M[0..rs) and P[0..rs) of (left key, right key) pairs, species
// y at index y + 1 so that the children (2i+1, 2i+2) of a species are one aligned 16-byte word.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint2 lds64(unsigned a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(unsigned a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(unsigned a, unsigned x, unsigned y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts128(unsigned a, unsigned x, unsigned y, unsigned z, unsigned w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

#define FLAT2_NS 32   // sweep steps held in registers
#define FLAT2_NCH 16  // row chunks of 32 species held in registers (pitch <= 512)

// Child rows -> keys.  Every global load of the row (up to 2 x n_chunks per lane) is issued
// before the first one is consumed: one DRAM round trip per row.
template <bool IL, bool IR, bool CG = false>
__device__ __forceinline__ void flat2_load(const int32_t* __restrict__ Tl, const int32_t* __restrict__ Tr,
                                           const uint32_t* __restrict__ keylow, uint2* __restrict__ M,
                                           int n_chunks, unsigned vmul, unsigned infkey) {
    unsigned vl[FLAT2_NCH], vr[FLAT2_NCH];
#pragma unroll
    for (int j = 0; j < FLAT2_NCH; ++j) {
        vl[j] = vr[j] = 0u;
        if (j < n_chunks) {
            // CG: the row may have been written by another SM during this launch (chained rows)
            if (IL) vl[j] = (unsigned)(CG ? __ldcg(Tl + j * 32) : Tl[j * 32]);
            if (IR) vr[j] = (unsigned)(CG ? __ldcg(Tr + j * 32) : Tr[j * 32]);
        }
    }
#pragma unroll
    for (int j = 0; j < FLAT2_NCH; ++j) {
        if (j < n_chunks) {
            const unsigned low = __ldg(keylow + j * 32);
            const unsigned a = IL ? vl[j] * vmul + low : infkey + low;
            const unsigned b = IR ? vr[j] * vmul + low : infkey + low;
            M[j * 32] = make_uint2(a, b);
        }
    }
}

// value + loss * depth of a key, on the FMA pipe where possible (the ALU pipe is the busy one):
// hi32(k * 2^(32-sh)) = k >> sh
__device__ __forceinline__ int flat2_q(unsigned k, unsigned hi_mul, unsigned dm, int loss) {
    const int ld = loss * (int)(k & dm);
    unsigned q;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(q) : "r"(k), "r"(hi_mul), "r"((unsigned)ld));
    return (int)q;
}
__device__ __forceinline__ int flat2_pack(int sum, int idx) {  // sum * 8 + idx, as an IMAD
    int w;
    asm("mad.lo.s32 %0, %1, 8, %2;" : "=r"(w) : "r"(sum), "r"(idx));
    return w;
}

// CHAIN = false: the rows of one wave's "dense" list.  CHAIN = true: rows [row_begin, row_begin + n_rows) of
// SEVERAL consecutive waves in one launch; a row waits for the done flags of its child rows inside the
// range (chain_sync[1 + i]) and raises its own after its cells are written.  CTAs take their 8 rows by
// ticket (chain_sync[0]), so every row a CTA can wait for belongs to a CTA that is already running.
template <int NS, bool CHAIN>
__global__ void __launch_bounds__(256, 4)
dtl_wave_flat2_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                      int64_t row_begin, int64_t chunk_row0, int n_rows, int32_t* __restrict__ T,
                      uint32_t* __restrict__ tags, DtlCosts cs, Key32 kp, const uint32_t* __restrict__ rowmask,
                      int sparse_max, int pf_near, int pf_far, const int* __restrict__ row_list,
                      const int* __restrict__ list_count, int* __restrict__ chain_sync, int cherry_end,
                      const uint8_t* __restrict__ row_cnt, int compact_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int cta_ticket;
    const int pitch = sp.pitch, rs = sp.flat_rs, n_chunks = pitch >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int local = 0;
    if (CHAIN) {
        if (threadIdx.x == 0) cta_ticket = atomicAdd(chain_sync, 1);
        __syncthreads();  // the only CTA-wide barrier: the warps are independent from here on
        local = cta_ticket * (blockDim.x >> 5) + warp;
        if (local >= n_rows) return;
    }
    // CHAIN = false: the rows of this launch are the wave's "dense" list written by dtl_partition_kernel.  The
    // grid may be smaller than the list (low waves have few dense rows and are launched with a fraction of the
    // wave's rows): a warp then strides over the list.
    for (int pos = blockIdx.x * (blockDim.x >> 5) + warp;; pos += gridDim.x * (blockDim.x >> 5)) {
    if (!CHAIN) {
        if (pos >= *list_count) return;
        local = row_list[pos];
    }
    const int64_t gr = row_begin + local;
    const size_t slot = (size_t)(gr - chunk_row0);
    uint2* Mrow = reinterpret_cast<uint2*>(smem_raw) + (size_t)warp * 2 * rs;
    uint2* M = Mrow + 1;
    uint2* P = M + rs;
    const unsigned Mb = (unsigned)__cvta_generic_to_shared(Mrow);
    const unsigned pofs = (unsigned)rs * 8u;

    // ---- child rows -> keys ------------------------------------------------------------------
    const int cl = row_cl[gr], cr = row_cr[gr];
    SR2_DCHECK(local >= 0 && local < n_rows && cl < (int)slot && cr < (int)slot);
    SR2_DCHECK(cl >= -sp.S && cr >= -sp.S);      // a leaf child is -1 - species
    const unsigned vmul = 1u << kp.sh, infkey = 0xffffffffu << kp.sh;
    if (CHAIN) {
        // child rows inside the chained range: wait for their flags (lanes 0 and 1 poll one child each)
        const int first_slot = (int)(row_begin - chunk_row0);
        const int c = lane == 0 ? cl : cr;
        if (lane < 2 && c >= first_slot) {
            // Rows are handed out by ticket in row order and rows are sorted by height (sr2_build_rows*), so a
            // child inside the chained range always belongs to a warp that already runs: the wait cannot
            // deadlock.  A row layout that breaks the invariant (a child at or behind its parent) must not
            // hang the GPU: it traps, and the host reports the failed launch (SR2_ERR_CUDA).
            if (c - first_slot >= local) __trap();
            const int* flag = chain_sync + 1 + (c - first_slot);
            int up;
            for (unsigned spins = 0;; ++spins) {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(up) : "l"(flag) : "memory");
                if (up) break;
                if (spins > (1u << 24)) __trap();   // seconds of waiting: something upstream failed
                __nanosleep(100);
            }
        }
        __syncwarp();
    }
    {
        const int32_t* __restrict__ Tl = T + (size_t)(cl >= 0 ? cl : 0) * pitch + lane;
        const int32_t* __restrict__ Tr = T + (size_t)(cr >= 0 ? cr : 0) * pitch + lane;
        const uint32_t* __restrict__ kl = sp.keylow_p + lane;
        // rows to read: internal children that are not virtual cherries (slots below cherry_end) and not
        // compact rows (row_cnt != 0: the entries of a sparse row written by dtl_wave_thread_kernel)
        const int ccl = (row_cnt && cl >= cherry_end) ? (int)row_cnt[cl] : 0;
        const int ccr = (row_cnt && cr >= cherry_end) ? (int)row_cnt[cr] : 0;
        const bool il = cl >= cherry_end && ccl == 0, ir = cr >= cherry_end && ccr == 0;
        if (il && ir) flat2_load<true, true, CHAIN>(Tl, Tr, kl, M + lane, n_chunks, vmul, infkey);
        else if (il) flat2_load<true, false, CHAIN>(Tl, Tr, kl, M + lane, n_chunks, vmul, infkey);
        else if (ir) flat2_load<false, true, CHAIN>(Tl, Tr, kl, M + lane, n_chunks, vmul, infkey);
        else flat2_load<false, false, CHAIN>(Tl, Tr, kl, M + lane, n_chunks, vmul, infkey);
        // a virtual cherry child: its closed form along the two root paths, lanes over depths
        if (cherry_end > 0 || row_cnt) __syncwarp();  // M of the whole row is written
        // a compact child: its entries are scattered over the (infinite) key row
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const int cc = side ? ccr : ccl;
            if (cc == 0) continue;
            const uint32_t* __restrict__ Tc = reinterpret_cast<const uint32_t*>(T) + (size_t)(side ? cr : cl) * pitch;
            for (int e = lane; e < cc; e += 32) {
                const unsigned v = Tc[e], y = v >> TH_VAL_BITS, val = v & TH_VAL_MASK;
                SR2_DCHECK((int)y < sp.S);
                const unsigned key = (val == TH_VAL_MASK ? infkey : val * vmul) + __ldg(sp.keylow_p + y);
                if (side) M[y].y = key;
                else M[y].x = key;
            }
        }
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const int c = side ? cr : cl;
            if (c < 0 || c >= cherry_end) continue;
            const CherryForm f = cherry_form(sp, -1 - row_cl[chunk_row0 + c], -1 - row_cr[chunk_row0 + c]);
            for (int which = 0; which < 2; ++which) {
                const uint16_t* path = which ? f.pb : f.pa;
                const int last = which ? f.db : f.da;
                for (int d = (which ? f.dl + 1 : 0) + lane; d <= last; d += 32) {
                    const int y = (int)path[d];
                    const unsigned v = cherry_path_value(f, which, d, cs);
                    if (v != 0xffffffffu) {
                        const unsigned key = v * vmul + __ldg(sp.keylow_p + y);
                        if (side) M[y].y = key;
                        else M[y].x = key;
                    }
                }
            }
        }
    }
    // The kernel is latency-bound (two dependent DRAM trips per row: descriptors, then child rows).
    // Pull the descriptors of a row far ahead into L2, and read those of a nearer row (prefetched by
    // an earlier CTA) to pull ITS child rows into L2; the prefetches are issued before the sweeps.
    int cl2 = -1, cr2 = -1;
    if (!CHAIN) {
        const int n_list = *list_count;
        if (pf_far > 0 && pos + pf_far < n_list && lane < 2)
            asm volatile("prefetch.global.L2 [%0];" ::"l"((lane ? row_cr : row_cl) + row_begin + row_list[pos + pf_far]));
        if (pf_near > 0 && pos + pf_near < n_list) {
            const int64_t g2 = row_begin + row_list[pos + pf_near];
            cl2 = row_cl[g2];
            cr2 = row_cr[g2];
        }
    }
    // the row's root-path bit set (dtl_partition_kernel): one 32-species word per lane
    const unsigned mword = lane < n_chunks ? rowmask[slot * sp.mask_stride + lane] : 0u;
    // sweep schedule -> registers, as absolute shared-memory addresses of this warp's row
    // (the table is padded to NS steps with idle words)
    unsigned sw[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const unsigned w = __ldg(sp.flat_sched + i * 32 + lane);
        sw[i] = w ? w + Mb * 0x10001u : 0u;  // 0 = this lane idles in step i
    }
    __syncwarp();
    if (lane == 0) {
        // a leaf child: 0 at its own species
        if (cl < 0) M[-1 - cl].x = __ldg(sp.keylow_p + (-1 - cl));
        if (cr < 0) M[-1 - cr].y = __ldg(sp.keylow_p + (-1 - cr));
        const uint2 none = make_uint2(0xffffffffu, 0xffffffffu);
        Mrow[pitch + 2] = none;  // the "children" of a leaf species
        Mrow[pitch + 3] = none;
        P[0] = none;             // nothing is incomparable to the root
    }
    __syncwarp();

    if (lane * 32 < pitch) {  // one 128-byte line per lane
        if (cl2 >= cherry_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(T + (size_t)cl2 * pitch + lane * 32));
        if (cr2 >= cherry_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(T + (size_t)cr2 * pitch + lane * 32));
    }
    // ---- up-sweep: SUBMIN in place, deepest level first ------------------------------------------
#pragma unroll
    for (int i = NS - 1; i >= 0; --i) {
        // idle lanes stay out of the shared-memory pipe (the kernel's busiest unit): a level of w
        // species costs ceil(w / 8) wavefronts for the child pairs instead of 4
        if (sw[i]) {
            const unsigned pa = sw[i] & 0xffffu, ca = sw[i] >> 16;
            const uint4 k = lds128(ca);
            const uint2 own = lds64(pa);
            sts64(pa, min(own.x, min(k.x, k.z)), min(own.y, min(k.y, k.w)));
        }
        __syncwarp();
    }
    // ---- top-down sweep: SEPMIN into P ------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        if (sw[i]) {
            const unsigned pa = sw[i] & 0xffffu, ca = sw[i] >> 16;
            const uint4 k = lds128(ca);
            const uint2 pp = lds64(pa + pofs);
            // incomparable(c) = incomparable(p) + subtree(c+1), and the other way round
            sts128(ca + pofs, min(pp.x, k.z), min(pp.y, k.w), min(pp.x, k.x), min(pp.y, k.y));
        }
        __syncwarp();
    }

    // ---- cells ---------------------------------------------------------------------------------------
    int32_t* __restrict__ Trow = T + slot * pitch;
    uint32_t* __restrict__ Grow = tags + slot * pitch;
    const unsigned char* Mbytes = reinterpret_cast<const unsigned char*>(Mrow);
    const int loss = cs.loss, db = kp.db, thresh = kp.thresh;
    const unsigned dm = kp.dm, rm = kp.rm, hi_mul = 1u << (32 - kp.sh);
    const int c12 = -2 * loss, c3 = cs.dup, c45 = cs.hgt;
    // one cell: species x, its word (child pair offset << 16 | depth); writes value and tag at entry `out` of the
    // row: the species itself, or, for a compact row, the position of x in the path set (value | x << 22)
    auto cell = [&](int x, unsigned mx, int out, bool compact) {
        const uint2 m = M[x], pq = P[x];
        const uint4 k = *reinterpret_cast<const uint4*>(Mbytes + (mx >> 16));  // (l0, r0, l1, r1)
        const int nld = -loss * (int)(mx & 0xffffu);
        const int qMl = flat2_q(m.x, hi_mul, dm, loss), qMr = flat2_q(m.y, hi_mul, dm, loss);
        const int ql0 = flat2_q(k.x, hi_mul, dm, loss), qr0 = flat2_q(k.y, hi_mul, dm, loss);
        const int ql1 = flat2_q(k.z, hi_mul, dm, loss), qr1 = flat2_q(k.w, hi_mul, dm, loss);
        const int pl = (int)__umulhi(pq.x, hi_mul), pr = (int)__umulhi(pq.y, hi_mul);
        const int b12 = 2 * nld + c12, b3 = 2 * nld + c3, b45 = nld + c45;
        const int w1 = flat2_pack(ql0 + qr1 + b12, 0);  // 1. speciation, left below x0, right below x1
        const int w2 = flat2_pack(ql1 + qr0 + b12, 1);  // 2. the other way round
        const int w3 = flat2_pack(qMl + qMr + b3, 2);   // 3. duplication
        const int w4 = flat2_pack(pl + qMr + b45, 3);   // 4. transfer, left child away
        const int w5 = flat2_pack(qMl + pr + b45, 4);   // 5. transfer, right child away
        const int w = min(min(min(w1, w2), w3), min(w4, w5));  // smallest value, then first candidate
        const int best = w >> 3, which = w & 7;
        unsigned ka = m.x, kb = m.y;
        ka = which == 0 ? k.x : ka;
        ka = which == 1 ? k.z : ka;
        ka = which == 3 ? pq.x : ka;
        kb = which == 0 ? k.w : kb;
        kb = which == 1 ? k.y : kb;
        kb = which == 4 ? pq.y : kb;
        const bool finite = best < thresh;
        Trow[out] = compact ? (int)((finite ? (unsigned)best : TH_VAL_MASK) | ((unsigned)x << TH_VAL_BITS)) : (finite ? best : SR2_INF);
        Grow[out] = finite ? (((ka >> db) & rm) | (((kb >> db) & rm) << 16)) : SR2_NO_TAG;
    };
    // A cell can only be finite on the root path of a leaf species below the object node.  Rows whose
    // path set is small (most rows: low in the object trees) write the infinite row with 128-bit
    // stores and evaluate the cells of the set only, ceil(m / 32) passes instead of pitch / 32; cells
    // outside the set get no tag (readers look at the value first).
    const int cnt = __popc(mword);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const int m_set = __shfl_sync(0xffffffffu, incl, 31);
    // With compact rows (dtl_partition_kernel set row_cnt[slot] = m for the rows of its lists with m <= compact_max)
    // the row is just its m entries in BFS order, written by consecutive lanes: no infinite row, no scattered
    // 4-byte stores.
    const bool compact = !CHAIN && m_set <= compact_max;
    SR2_DCHECK(!row_cnt || CHAIN || (int)row_cnt[slot] == (compact ? m_set : 0));
    if (m_set <= sparse_max || compact) {
        if (!compact) {
            uint4* __restrict__ row4 = reinterpret_cast<uint4*>(Trow);
            const uint4 inf4 = make_uint4(SR2_INF, SR2_INF, SR2_INF, SR2_INF);
            for (int i = lane; i < (pitch >> 2); i += 32) row4[i] = inf4;
            __syncwarp();
        }
        const int excl = incl - cnt;
        for (int i0 = 0; i0 < m_set; i0 += 32) {
            const int i = min(i0 + lane, m_set - 1);
            // the word that holds the i-th species of the set: the number of words that end at or before i
            // (binary search over the non-decreasing prefix counts; lanes past n_chunks hold m_set > i)
            int w = 0;
#pragma unroll
            for (int step = FLAT2_NCH / 2; step; step >>= 1)
                if (__shfl_sync(0xffffffffu, incl, w + step - 1) <= i) w += step;
            const unsigned word = __shfl_sync(0xffffffffu, mword, w);
            const int before = __shfl_sync(0xffffffffu, excl, w);
            const int x = w * 32 + (int)__fns(word, 0, i - before + 1);
            if (i0 + lane < m_set) cell(x, __ldg(sp.cmeta + x), compact ? i0 + lane : x, compact);
        }
    } else {
        // per-species words of the cells this lane evaluates
        unsigned meta[FLAT2_NCH];
#pragma unroll
        for (int j = 0; j < FLAT2_NCH; ++j) meta[j] = j < n_chunks ? __ldg(sp.cmeta + j * 32 + lane) : 0u;
#pragma unroll
        for (int j = 0; j < FLAT2_NCH; ++j)
            if (j < n_chunks) cell(j * 32 + lane, meta[j], j * 32 + lane, false);
    }
    if (CHAIN) {  // the row (values and tags of every lane) is visible before its flag
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            int* flag = chain_sync + 1 + local;
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
        }
        return;
    }
    __syncwarp();  // the warp's shared-memory row is reused by its next list entry
    }
}

// ---------------------------------------------------------------------------
// Sparse rows.  A cell can only be finite on the root path of a leaf species below its object node
// (88 % of config #2's cells are infinite), so low rows of the object trees hold a few dozen live
// cells out of |S|.  dtl_partition_kernel computes every row's root-path bit set and splits the wave
// into rows with at most SPARSE_MAXM path species ("sparse") and the rest ("dense", flat2).
// dtl_wave_sparse_kernel gives a sparse row 4 lanes instead of a warp and works on the path set only:
//   elements   = the path species in BFS order (non-decreasing depth), position e in [0, m)
//   Mc / Pc    = SUBMIN / SEPMIN (left key, right key) pairs per element, in shared memory
//   up-sweep   = one step per depth, deepest first: atomicMin of an element into its parent's slot
//   down-sweep = one step per depth from the top: P[e] = min(P[parent], M[sibling] if on a path)
//   cells      = element-wise, the same candidates and packing as flat2
// 8 rows per warp, 1.5 KB of shared memory per row instead of 6.8 KB: 4 x as many rows in flight
// and every lane busy where the dense sweeps leave 90 % of the lanes idle.
// ---------------------------------------------------------------------------
#define SPARSE_MAXM 57        // largest path set of a "sparse" row: 1,376 bytes of shared memory per row, 5 CTAs (20 warps) per SM
#define SPARSE_MAXM_SMALL 38  // ... of the first tier: 928 bytes per row, 7 CTAs (28 warps) per SM
// (measured, small / large tier -> ms per step: 44 / 64 -> 6.57, 48 / 64 -> 6.52, 52 / 64 -> 6.76, 48 / 57 -> 6.43, 38 / 57 -> 6.39)
// shared-memory words per row for a tier of MAXM species: Mc 2 MAXM, Pc 2 MAXM, el MAXM, mk + pre (later the
// links) MAXM, padded to 8 or 24 mod 32 so that the 4 rows of a half-warp start 8 banks apart
#define SPARSE_ROW_WORDS(MAXM) ((MAXM) == 38 ? 232 : 344)

// 4 lanes per row, one uint4 (128 root-path bits) per lane and child: 64 rows per CTA in flight.
// class_of[size of the path set] = the row's class; the last class holds the "dense" rows (flat2).  Without
// compact rows: 0 / 1 = the two tiers of dtl_wave_sparse_kernel.  With compact rows (dtl_wave_thread_kernel)
// the classes below the last are narrow bins of path-set sizes, so that the 32 rows of a warp of that kernel
// have sets of nearly the same size; rows of at most compact_max species are stored compact.
// lists[cls * list_pitch + i], counts[cls].
#define DTL_MAX_CLASSES 32
__global__ void __launch_bounds__(256)
dtl_partition_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                     int64_t row_begin, int64_t chunk_row0, int n_rows, uint32_t* __restrict__ rowmask,
                     const uint8_t* __restrict__ class_of, int compact_max, int* __restrict__ lists, int list_pitch,
                     int* __restrict__ counts, int cherry_rows, uint8_t* __restrict__ row_cnt,
                     uint4* __restrict__ row_info) {
    __shared__ int cta_cnt[DTL_MAX_CLASSES], cta_base[DTL_MAX_CLASSES];
    const int sub = threadIdx.x & 3;
    const int local = blockIdx.x * 64 + (threadIdx.x >> 2);
    const int ms = sp.mask_stride, quads = ms >> 2;
    if (threadIdx.x < DTL_MAX_CLASSES) cta_cnt[threadIdx.x] = 0;
    __syncthreads();
    int cls = -1, my_off = 0, cnt = 0;
    const bool valid = local < n_rows;
    if (valid) {
        const int64_t gr = row_begin + local;
        const int cl = row_cl[gr], cr = row_cr[gr];
        // a child's set: its row's (internal child), the root path of its species (leaf child), or -- for a
        // cherry child, a slot below cherry_rows, whose own set is never stored -- the paths of its two leaves
        const uint4* a4 = reinterpret_cast<const uint4*>(cl >= 0 ? rowmask + (size_t)cl * ms : sp.pathmask + (size_t)(-1 - cl) * ms);
        const uint4* b4 = reinterpret_cast<const uint4*>(cr >= 0 ? rowmask + (size_t)cr * ms : sp.pathmask + (size_t)(-1 - cr) * ms);
        const uint4* a5 = a4;
        const uint4* b5 = b4;
        // leaf species behind the two children: of a leaf child (one), of a cherry child (two)
        int la = cl < 0 ? -1 - cl : -1, lb = la, ra = cr < 0 ? -1 - cr : -1, rb = ra;
        if (cl >= 0 && cl < cherry_rows) {
            la = -1 - row_cl[chunk_row0 + cl];
            lb = -1 - row_cr[chunk_row0 + cl];
            a4 = reinterpret_cast<const uint4*>(sp.pathmask + (size_t)la * ms);
            a5 = reinterpret_cast<const uint4*>(sp.pathmask + (size_t)lb * ms);
        }
        if (cr >= 0 && cr < cherry_rows) {
            ra = -1 - row_cl[chunk_row0 + cr];
            rb = -1 - row_cr[chunk_row0 + cr];
            b4 = reinterpret_cast<const uint4*>(sp.pathmask + (size_t)ra * ms);
            b5 = reinterpret_cast<const uint4*>(sp.pathmask + (size_t)rb * ms);
        }
        // dtl_wave_thread_kernel's view of the row in two 16-byte words: child descriptors, preorder times of
        // the four leaf species, and per side depth(a) | depth(b) << 8 | depth(lca(a, b)) << 16 (a leaf or cherry
        // child) or entries of the child's compact row << 24 (a stored child)
        if (row_info && sub < 2) {
            const int a = sub ? ra : la, b = sub ? rb : lb;
            unsigned ta = 0u, tb = 0u, dpack = 0u;
            if (a >= 0) {
                ta = (unsigned)__ldg(sp.tin + a);
                tb = (unsigned)__ldg(sp.tin + b);
                dpack = (unsigned)__ldg(sp.depth + a) | ((unsigned)__ldg(sp.depth + b) << 8) |
                        ((unsigned)__ldg(sp.lca_depth + (size_t)a * sp.S + b) << 16);
            } else if (row_cnt) {
                // a stored child: the size of its compact row (0 = dense), written by this kernel's launch for the
                // child's wave (same stream) -- saves dtl_wave_thread_kernel a dependent trip to memory
                dpack = (unsigned)row_cnt[sub ? cr : cl] << 24;
            }
            row_info[(size_t)(gr - chunk_row0) * 2 + sub] = make_uint4((unsigned)(sub ? cr : cl), ta, tb, dpack);
        }
        uint4* o4 = reinterpret_cast<uint4*>(rowmask + (size_t)(gr - chunk_row0) * ms);
        for (int w = sub; w < quads; w += 4) {
            const uint4 a = a4[w], b = b4[w], a2 = a5[w], b2 = b5[w];
            const uint4 o = make_uint4(a.x | b.x | a2.x | b2.x, a.y | b.y | a2.y | b2.y, a.z | b.z | a2.z | b2.z,
                                       a.w | b.w | a2.w | b2.w);
            o4[w] = o;
            cnt += __popc(o.x) + __popc(o.y) + __popc(o.z) + __popc(o.w);
        }
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    if (valid) {
        cls = (int)class_of[min(cnt, sp.S)];
        if (sub == 0) my_off = atomicAdd(&cta_cnt[cls], 1);
        // compact rows (dtl_wave_thread_kernel): the size of the row's path set; 0 = a dense row
        if (sub == 0 && row_cnt) row_cnt[row_begin + local - chunk_row0] = (uint8_t)(cnt <= compact_max ? cnt : 0);
    }
    __syncthreads();
    if (threadIdx.x < DTL_MAX_CLASSES)
        cta_base[threadIdx.x] = cta_cnt[threadIdx.x] ? atomicAdd(&counts[threadIdx.x], cta_cnt[threadIdx.x]) : 0;
    __syncthreads();
    if (cls >= 0 && sub == 0) lists[(size_t)cls * list_pitch + cta_base[cls] + my_off] = local;
}

template <int MAXM>
__global__ void __launch_bounds__(128)
dtl_wave_sparse_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                       int64_t row_begin, int64_t chunk_row0, int32_t* __restrict__ T, uint32_t* __restrict__ tags,
                       DtlCosts cs, Key32 kp, const uint32_t* __restrict__ rowmask, const int* __restrict__ row_list,
                       const int* __restrict__ list_count, int pf_near, int pf_far, int cherry_end) {
    extern __shared__ uint32_t sp_smem[];  // [4 warps][8 rows][SPARSE_ROW_WORDS(MAXM)]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, sub = lane & 3;
    const int n_list = *list_count;
    const int first = (blockIdx.x * 4 + warp) * 8;
    if (first >= n_list) return;
    const bool valid = first + g < n_list;
    const int pitch = sp.pitch, mw = sp.mask_words, ms = sp.mask_stride;
    uint32_t* base = sp_smem + (size_t)(warp * 8 + g) * SPARSE_ROW_WORDS(MAXM);
    uint2* Mc = reinterpret_cast<uint2*>(base);
    uint2* Pc = Mc + MAXM;
    uint32_t* el = base + 4 * MAXM;  // species | depth << 16
    uint32_t* mk = el + MAXM;        // the row's path set (<= 16 words)
    uint32_t* pre = mk + 16;         // pre[w] = path species in words < w; pre[mw] = m (<= 17 words)
    int64_t gr = 0;
    size_t slot = 0;
    int cl = -1, cr = -1, m = 0;
    int2 cha = make_int2(-1, -1), chb = make_int2(-1, -1);
    if (valid) {
        gr = row_begin + row_list[first + g];
        slot = (size_t)(gr - chunk_row0);
        cl = row_cl[gr];
        cr = row_cr[gr];
        // a virtual cherry child: its two leaf species (its own descriptors) instead of a row to read
        if (cl >= 0 && cl < cherry_end) cha = make_int2(-1 - row_cl[chunk_row0 + cl], -1 - row_cr[chunk_row0 + cl]);
        if (cr >= 0 && cr < cherry_end) chb = make_int2(-1 - row_cl[chunk_row0 + cr], -1 - row_cr[chunk_row0 + cr]);
        // the row's own path set (dtl_partition_kernel); the children's rows are infinite outside
        // their own sets, so their values are simply gathered at every species of this set
        // (word loads: one uint4 per lane measured 10 % slower for the whole kernel)
        for (int w = sub; w < mw; w += 4) mk[w] = rowmask[slot * ms + w];
    }
    // latency: list entry -> descriptors / path set -> child values are three dependent DRAM trips.
    // Rows further down the list are pulled into L2 ahead of the CTAs that will work on them: the
    // descriptors and path set of a far row, the child rows of a nearer one.
    int r_far = -1, r_near = -1;
    if (first + g + pf_far < n_list && pf_far > 0) r_far = row_list[first + g + pf_far];
    if (first + g + pf_near < n_list && pf_near > 0) r_near = row_list[first + g + pf_near];
    __syncwarp();
    if (valid && sub == 0) {
        int run = 0;
        for (int w = 0; w < mw; ++w) {
            pre[w] = run;
            run += __popc(mk[w]);
        }
        pre[mw] = run;
    }
    __syncwarp();
    if (valid) m = (int)pre[mw];
    SR2_DCHECK(m <= MAXM);                       // dtl_partition_kernel only lists rows that fit the tier
    SR2_DCHECK(!valid || (cl < (int)slot && cr < (int)slot));   // children are rows of earlier waves
    // position of a species of the path set
    auto pos_of = [&](int s) { return (int)pre[s >> 5] + __popc(mk[s >> 5] & ((1u << (s & 31)) - 1u)); };
    auto on_path = [&](int s) { return (mk[s >> 5] >> (s & 31)) & 1u; };

    // ---- elements: the path species in BFS order ------------------------------------------------------
    {
        int w = 0;
        for (int e = sub; e < m; e += 4) {
            while ((int)pre[w + 1] <= e) ++w;
            const int s = w * 32 + (int)__fns(mk[w], 0, e - (int)pre[w] + 1);
            SR2_DCHECK(s >= 0 && s < sp.S && w < mw);
            const unsigned low = __ldg(sp.keylow_p + s);
            el[e] = (unsigned)s | ((low & kp.dm) << 16);
            // positions of the parent, the sibling and the two children inside the path set (0xff = not
            // on a path); parked in Pc until the bit sets are no longer needed
            const int par = s ? __ldg(sp.parent + s) : 0, c0 = __ldg(sp.child0 + s);
            const int sib = s ? ((s & 1) ? s + 1 : s - 1) : 0;
            unsigned link = (unsigned)pos_of(par);
            link |= (s && on_path(sib) ? (unsigned)pos_of(sib) : 0xffu) << 8;
            link |= (c0 >= 0 && on_path(c0) ? (unsigned)pos_of(c0) : 0xffu) << 16;
            link |= (c0 >= 0 && on_path(c0 + 1) ? (unsigned)pos_of(c0 + 1) : 0xffu) << 24;
            SR2_DCHECK(on_path(par) && (int)(link & 0xffu) <= e);   // root paths are closed under "parent"
            Pc[e].x = link;
        }
    }
    int cl_near = -1, cr_near = -1;
    if (r_near >= 0) {
        cl_near = row_cl[row_begin + r_near];
        cr_near = row_cr[row_begin + r_near];
    }
    if (r_far >= 0) {
        const int64_t g_far = row_begin + r_far;
        if (sub == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(row_cl + g_far));
        if (sub == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(row_cr + g_far));
        if (sub == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowmask + (size_t)(g_far - chunk_row0) * ms));
    }
    // ---- child values -> keys; the gathers of four elements per lane are in flight together ----------
    const unsigned vmul = 1u << kp.sh, infkey = 0xffffffffu << kp.sh;
    for (int e0 = sub; e0 < m; e0 += 16) {
        unsigned va[4], vb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = e0 + 4 * q;
            va[q] = vb[q] = 0xffffffffu;  // "not on the child's paths": infinite
            if (e < m) {
                const int s = (int)(el[e] & 0xffffu);
                if (cl >= cherry_end)
                    va[q] = (unsigned)T[(size_t)cl * pitch + s];
                else if (s == -1 - cl)
                    va[q] = 0u;
                if (cr >= cherry_end)
                    vb[q] = (unsigned)T[(size_t)cr * pitch + s];
                else if (s == -1 - cr)
                    vb[q] = 0u;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = e0 + 4 * q;
            if (e < m) {
                const int s = (int)(el[e] & 0xffffu);
                const unsigned low = ((unsigned)s << kp.db) | (el[e] >> 16);
                // a stored infinity (0x3fffffff) wraps to the all-ones value field, like the sentinel
                Mc[e] = make_uint2(va[q] == 0xffffffffu ? infkey + low : va[q] * vmul + low,
                                   vb[q] == 0xffffffffu ? infkey + low : vb[q] * vmul + low);
            }
        }
    }
    // virtual cherry children (left infinite above): the closed form along the two root paths, lanes
    // over depths; a path species is always in this row's set
    __syncwarp();  // Mc of the whole row is written
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        const int2 ch = side ? chb : cha;
        if (ch.x < 0) continue;
        const CherryForm f = cherry_form(sp, ch.x, ch.y);
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const uint16_t* path = which ? f.pb : f.pa;
            const int last = which ? f.db : f.da, d_first = (which ? f.dl + 1 : 0) + sub;
            for (int d0 = d_first; d0 <= last; d0 += 16) {
                int sd[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) sd[q] = d0 + 4 * q <= last ? (int)path[d0 + 4 * q] : -1;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int d = d0 + 4 * q, y = sd[q];
                    if (y < 0) continue;
                    const unsigned v = cherry_path_value(f, which, d, cs);
                    if (v == 0xffffffffu) continue;
                    const unsigned key = v * vmul + (((unsigned)y << kp.db) | (unsigned)d);
                    const int e = pos_of(y);
                    if (side) Mc[e].y = key;
                    else Mc[e].x = key;
                }
            }
        }
    }
    __syncwarp();
    uint32_t* lk = mk;  // [MAXM] links, over the bit set and its prefix counts (dead from here on)
    for (int e = sub; e < m; e += 4) lk[e] = Pc[e].x;
    // the infinite value rows of the warp's 8 rows, all lanes together (before any finite cell)
    {
        const uint4 inf4 = make_uint4(SR2_INF, SR2_INF, SR2_INF, SR2_INF);
        for (int r = 0; r < 8; ++r) {
            const unsigned long long row_slot = __shfl_sync(0xffffffffu, (unsigned long long)slot, r * 4);
            const int row_ok = __shfl_sync(0xffffffffu, (int)valid, r * 4);
            if (!row_ok) continue;
            uint4* row4 = reinterpret_cast<uint4*>(T + (size_t)row_slot * pitch);
            for (int i = lane; i < (pitch >> 2); i += 32) row4[i] = inf4;
        }
    }
    __syncwarp();
    const int D = sp.D;
    // ---- up-sweep: SUBMIN, one step per depth, deepest first ---------------------------------------
    {
        int e = m > sub ? sub + ((m - 1 - sub) >> 2) * 4 : -1;  // my last element
        for (int d = D - 1; d >= 1; --d) {
            while (e >= 0 && (int)(el[e] >> 16) == d) {
                const int pp = (int)(lk[e] & 0xffu);
                const uint2 k = Mc[e];
                atomicMin(&Mc[pp].x, k.x);
                atomicMin(&Mc[pp].y, k.y);
                e -= 4;
            }
            __syncwarp();
        }
    }
    // ---- down-sweep: SEPMIN, one step per depth from the top -----------------------------------------
    {
        int e = sub;
        if (valid && sub == 0) {
            Pc[0] = make_uint2(0xffffffffu, 0xffffffffu);  // element 0 is the root: nothing is incomparable
            e = 4;
        }
        __syncwarp();
        for (int d = 1; d < D; ++d) {
            while (e < m && (int)(el[e] >> 16) == d) {
                const unsigned link = lk[e];
                const int pp = (int)(link & 0xffu), sp_ = (int)((link >> 8) & 0xffu);
                uint2 q = Pc[pp];
                if (sp_ != 0xff) {  // the sibling (ranks 2i+1, 2i+2 are siblings) is on a path
                    const uint2 ms = Mc[sp_];
                    q.x = min(q.x, ms.x);
                    q.y = min(q.y, ms.y);
                }
                Pc[e] = q;
                e += 4;
            }
            __syncwarp();
        }
    }
    for (int line = sub; line * 32 < pitch; line += 4) {  // 128-byte lines of the two child rows
        if (cl_near >= cherry_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(T + (size_t)cl_near * pitch + line * 32));
        if (cr_near >= cherry_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(T + (size_t)cr_near * pitch + line * 32));
    }
    // ... and the descriptors of its virtual cherry children
    if (cl_near >= 0 && cl_near < cherry_end && sub < 2)
        asm volatile("prefetch.global.L2 [%0];" ::"l"((sub ? row_cr : row_cl) + chunk_row0 + cl_near));
    if (cr_near >= 0 && cr_near < cherry_end && sub >= 2)
        asm volatile("prefetch.global.L2 [%0];" ::"l"((sub == 3 ? row_cr : row_cl) + chunk_row0 + cr_near));
    // ---- cells: the path species only ---------------------------------------------------------------
    if (!valid) return;
    int32_t* __restrict__ Trow = T + slot * pitch;
    uint32_t* __restrict__ Grow = tags + slot * pitch;
    const int loss = cs.loss, db = kp.db, thresh = kp.thresh;
    const unsigned dm = kp.dm, rm = kp.rm, hi_mul = 1u << (32 - kp.sh);
    const int c12 = -2 * loss, c3 = cs.dup, c45 = cs.hgt;
    for (int e = sub; e < m; e += 4) {
        const int x = (int)(el[e] & 0xffffu), dx = (int)(el[e] >> 16);
        const uint2 mm = Mc[e], pq = Pc[e];
        uint4 k = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);  // (l0, r0, l1, r1)
        const int p0 = (int)((lk[e] >> 16) & 0xffu), p1 = (int)(lk[e] >> 24);
        if (p0 != 0xff) {
            const uint2 t = Mc[p0];
            k.x = t.x, k.y = t.y;
        }
        if (p1 != 0xff) {
            const uint2 t = Mc[p1];
            k.z = t.x, k.w = t.y;
        }
        const int nld = -loss * dx;
        const int qMl = flat2_q(mm.x, hi_mul, dm, loss), qMr = flat2_q(mm.y, hi_mul, dm, loss);
        const int ql0 = flat2_q(k.x, hi_mul, dm, loss), qr0 = flat2_q(k.y, hi_mul, dm, loss);
        const int ql1 = flat2_q(k.z, hi_mul, dm, loss), qr1 = flat2_q(k.w, hi_mul, dm, loss);
        const int pl = (int)__umulhi(pq.x, hi_mul), pr = (int)__umulhi(pq.y, hi_mul);
        const int b12 = 2 * nld + c12, b3 = 2 * nld + c3, b45 = nld + c45;
        const int w1 = flat2_pack(ql0 + qr1 + b12, 0);  // 1. speciation, left below x0, right below x1
        const int w2 = flat2_pack(ql1 + qr0 + b12, 1);  // 2. the other way round
        const int w3 = flat2_pack(qMl + qMr + b3, 2);   // 3. duplication
        const int w4 = flat2_pack(pl + qMr + b45, 3);   // 4. transfer, left child away
        const int w5 = flat2_pack(qMl + pr + b45, 4);   // 5. transfer, right child away
        const int w = min(min(min(w1, w2), w3), min(w4, w5));  // smallest value, then first candidate
        const int best = w >> 3, which = w & 7;
        unsigned ka = mm.x, kb = mm.y;
        ka = which == 0 ? k.x : ka;
        ka = which == 1 ? k.z : ka;
        ka = which == 3 ? pq.x : ka;
        kb = which == 0 ? k.w : kb;
        kb = which == 1 ? k.y : kb;
        kb = which == 4 ? pq.y : kb;
        const bool finite = best < thresh;
        Trow[x] = finite ? best : SR2_INF;
        Grow[x] = finite ? (((ka >> db) & rm) | (((kb >> db) & rm) << 16)) : SR2_NO_TAG;
    }
}

// ---------------------------------------------------------------------------
// Sparse rows, one THREAD per row, stored COMPACT.  The 4-lanes-per-row kernel above is bound by
// instruction latency (a warp of 8 rows issues ~4,900 instructions, one every ~12 cycles, at 28 warps per SM)
// and by its stores: a dense row is 1.6 KB of infinity plus m scattered 4-byte cells and tags, every one a
// partial-sector write (measured with the stores compiled out: 1.7 -> 0.63 ms for wave 2).  Here
//   * a lane owns a whole row and walks its path set serially, so a warp-instruction works on 32 rows; the
//     row's state sits in shared memory as columns (word i of thread t at [i][t]: any per-thread index is
//     conflict-free), 12 bytes per path species:
//       el[e]  species | depth << 9 | parent position << 14 | first-child / has-sibling << 22 | which of its
//              children are on a path << 26 | "on the root path of leaf la / lb / ra / rb" << 28..31
//              (elements are in BFS order: the children of the elements follow one another, so the position of
//              an element's first child is a running count in a forward pass and needs no field)
//       Mc[e]  (left key, right key): child values -> SUBMIN (up-sweep: e = m-1 .. 1 folds into its parent)
//              -> SEPMIN, in place: siblings are adjacent in BFS order, an element's SUBMIN is last needed by
//              the sibling after it (kept in registers for one step), its SEPMIN by its children further on;
//   * the row is written compact: entry e of the row's slot in T is value | species << 22 (TH_VAL_MASK = the
//     infinite value), entry e of its tag row the usual tag, e < m = row_cnt[slot] in BFS order -- nothing for
//     the infinite cells.  A lane writes four entries of its row at a time, 16 bytes of values and 16 of tags (the
//     two halves of a sector follow each other within a few dozen instructions and merge in L2; a per-warp staging
//     tile that made full 32-byte sectors of them cost more instructions than it saved: 4.73 -> 4.66 ms per step);
//   * readers: this kernel (a child's path set is a subset of its parent's, in the same order: a merge),
//     dtl_wave_flat2_kernel (scatters a compact child into its dense key row), dtl_select_trace_kernel (the
//     position of a species = number of path species before it, from the row's bit set).
// The path set is generated top-down from the row's bit set (children of element p are appended in BFS
// order, so the positions of parents / children / siblings fall out of the generation).  la, lb / ra, rb are
// the leaf species of a leaf child (la only) or of a virtual cherry child: "species on the root path of
// leaf a" is decided during the generation from preorder times (a sits below the first child of s iff
// tin[a] < tin[second child]), so leaf rows and the closed form of cherry rows need no look-up per element.
// Where the time went and what was done about it (profiles/README.md, r7a-r7c): the per-element look-up of the
// species word (an L2 trip each) -> the table is staged per CTA, 4 bytes per species; the entries of the stored
// children read one by one inside the merge -> everything the row reads (record, bit set, children) is requested
// up front, the bit set parked in the staging tile, the children's entries in the key slots (left .x, right .y)
// and merged BACKWARDS in place (entry q of a child matches an element e >= q).
// Only used with virtual cherries (every stored child of a sparse row is then a compact row).
// ---------------------------------------------------------------------------
#ifndef THREAD_ROWS
#define THREAD_ROWS 64   // rows (threads) per CTA
#endif
#define TH_S_MASK 511u
#define TH_D_SHIFT 9
#define TH_PP_SHIFT 14   // 8 bits: a path set of the thread kernel has at most THREAD_MAXM species
#define TH_PP_MASK 255u
#define TH_FIRST_BIT 22  // the first of its parent's children that are on a path
#define TH_SIB_BIT 23    // its sibling is on a path too (then it is the element next to it)
#define TH_C0_BIT 26
#define TH_C1_BIT 27
#define THREAD_MAXM 128   // largest tier (row_cnt is a byte, compact rows reach 255 entries)
#define TH_FLAG_SHIFT 28
// shared-memory words of a CTA: Mc [MAXM + 1] pairs, el [MAXM] and the bit set [MW = mask words] per thread
#define THREAD_SMEM_WORDS(MAXM, MW) ((3 * (MAXM) + 2 + (MW)) * THREAD_ROWS)
// ... plus the species table of the generation, one word per species: (first child + 1) << 16 | preorder time of the second child
#define THREAD_SMEM_BYTES(MAXM, MW, S) (4 * (size_t)(THREAD_SMEM_WORDS(MAXM, MW) + (((S) + 3) & ~3)))

// The rows of this launch: the classes [cls0, cls0 + n_cls) of the wave's lists, one after the other (a "tier":
// the classes whose sets have at most maxm species); the shared memory of the launch is sized for maxm.
// STORED = false: the rows of wave 2, whose children are leaves and (virtual) cherries only -- no child rows to load or merge.
template <bool STORED>
__global__ void __launch_bounds__(THREAD_ROWS)
dtl_wave_thread_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                       int64_t row_begin, int64_t chunk_row0, int32_t* __restrict__ T, uint32_t* __restrict__ tags,
                       DtlCosts cs, Key32 kp, const uint32_t* __restrict__ rowmask, const int* __restrict__ lists,
                       int list_pitch, const int* __restrict__ counts, int cls0, int n_cls, int MAXM, int cherry_end,
                       const uint8_t* __restrict__ row_cnt, const uint4* __restrict__ row_info) {
    extern __shared__ __align__(16) uint32_t th_smem[];
    const int t = threadIdx.x, lane = t & 31;
    const int idx = blockIdx.x * THREAD_ROWS + t;
    // the class this thread's list position falls into
    int n_list = 0, my_cls = -1, my_off = 0;
    for (int b = 0; b < n_cls; ++b) {
        const int c = counts[cls0 + b];
        if (idx >= n_list && idx < n_list + c) my_cls = cls0 + b, my_off = idx - n_list;
        n_list += c;
    }
    if (blockIdx.x * THREAD_ROWS >= n_list) return;  // the whole CTA is beyond the lists
    const bool valid = idx < n_list;
    const int pitch = sp.pitch, ms = sp.mask_stride;
    // the row's record is requested before the species table is staged (one trip to memory instead of two)
    size_t slot = 0;
    uint4 il = make_uint4(0u, 0u, 0u, 0u), ir = il;
    if (valid) {
        slot = (size_t)(row_begin - chunk_row0) + (size_t)lists[(size_t)my_cls * list_pitch + my_off];
        il = row_info[slot * 2];
        ir = row_info[slot * 2 + 1];
    }
    // The generation below looks up every path species (first child, preorder time of the second child: which
    // child a leaf species sits under).  From global memory that is a dependent trip to L2 per element (the
    // carve-out leaves no L1 to speak of): 27 % of the kernel's stall samples in the r7a capture.  The table is
    // staged per CTA instead, 4 bytes per species.
    uint32_t* spt = th_smem + THREAD_SMEM_WORDS(MAXM, sp.mask_words);
    for (int y = t; y < sp.S; y += THREAD_ROWS) {
        const uint2 z = __ldg(reinterpret_cast<const uint2*>(sp.span4 + y) + 1);  // (child0 + 1) << 16 | depth, tin of the second child
        spt[y] = (z.x & 0xffff0000u) | z.y;
    }
    __syncthreads();
    if (idx - lane >= n_list) return;  // the whole warp is beyond the lists
    uint2* Mc = reinterpret_cast<uint2*>(th_smem) + t;            // Mc[e * THREAD_ROWS]
    uint32_t* el = th_smem + 2 * (MAXM + 1) * THREAD_ROWS + t;    // el[e * THREAD_ROWS]
    // the row's bit set, shifted down by one bit: mask_words words per thread, [word][thread].  In BFS numbering the
    // children of a species are 2i + 1 and 2i + 2, so after the shift their two bits sit in one word at an even
    // position: one look-up answers "which of the children are on a path".
    uint32_t* bits = th_smem + (3 * MAXM + 2) * THREAD_ROWS + t;
    auto children_on_path = [&](int c0) {   // bit 0: c0, bit 1: c0 + 1
        SR2_DCHECK((c0 & 1) == 1);
        return (bits[((c0 - 1) >> 5) * THREAD_ROWS] >> ((c0 - 1) & 31)) & 3u;
    };
    int cl = -1, cr = -1, m = 0;
    // per side: 0 = leaf at species a, 1 = virtual cherry (a, b), 2 = stored (compact) row
    int mode_l = 0, mode_r = 0, m_l = 0, m_r = 0;
    // closed forms of the children that are leaves or virtual cherries (dtl_partition_kernel's record of the row):
    // depths of the leaf species a, b and of lca(a, b) per side, preorder times of the four species
    int l_da = 0, l_db = 0, l_dl = 0, r_da = 0, r_db = 0, r_dl = 0;
    if (valid) {
        cl = (int)il.x;
        cr = (int)ir.x;
        SR2_DCHECK(cl == row_cl[chunk_row0 + slot] && cr == row_cr[chunk_row0 + slot]);
        mode_l = cl < 0 ? 0 : (cl < cherry_end ? 1 : 2);
        mode_r = cr < 0 ? 0 : (cr < cherry_end ? 1 : 2);
        SR2_DCHECK(STORED || (mode_l != 2 && mode_r != 2));
        if (STORED && mode_l == 2) m_l = (int)(il.w >> 24);
        if (STORED && mode_r == 2) m_r = (int)(ir.w >> 24);
        SR2_DCHECK(mode_l != 2 || (m_l > 0 && m_l == (int)row_cnt[cl]));  // a child's path set is a subset of this
        SR2_DCHECK(mode_r != 2 || (m_r > 0 && m_r == (int)row_cnt[cr]));  // row's: it is a compact row
        // Everything the row reads from memory is requested here, in one go: its bit set and the entries of its
        // stored children (a subsequence of this row's elements, so they fit the key slots Mc[0 .. m), left child
        // in .x, right child in .y; the merge below walks backwards and overwrites them in place).  The r7b capture
        // had 23 % of the kernel's stall samples at the per-entry loads of the merge.
        const uint4* m4 = reinterpret_cast<const uint4*>(rowmask + slot * ms);
        const uint4* Tl4 = reinterpret_cast<const uint4*>(T + (size_t)(mode_l == 2 ? cl : 0) * pitch);
        const uint4* Tr4 = reinterpret_cast<const uint4*>(T + (size_t)(mode_r == 2 ? cr : 0) * pitch);
        uint4 mk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) mk[q] = q < (ms >> 2) ? m4[q] : make_uint4(0u, 0u, 0u, 0u);
        const int n4 = (max(m_l, m_r) + 3) >> 2;
        // (four quads per child in flight: a row of 64 entries costs 4 round trips, not 16)
        for (int j0 = 0; j0 < n4; j0 += 4) {
            uint4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                // (a quad may reach past the child's entries, never past its row: m <= 64 < pitch)
                a[u] = b[u] = make_uint4(0u, 0u, 0u, 0u);
                if (4 * (j0 + u) < m_l) a[u] = Tl4[j0 + u];
                if (4 * (j0 + u) < m_r) b[u] = Tr4[j0 + u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                // (Mc has MAXM + 1 slots; the last quad of a child with nearly MAXM entries must not spill into el)
                if (j < n4) {
                    Mc[(4 * j) * THREAD_ROWS] = make_uint2(a[u].x, b[u].x);
                    if (4 * j + 1 <= MAXM) Mc[(4 * j + 1) * THREAD_ROWS] = make_uint2(a[u].y, b[u].y);
                    if (4 * j + 2 <= MAXM) Mc[(4 * j + 2) * THREAD_ROWS] = make_uint2(a[u].z, b[u].z);
                    if (4 * j + 3 <= MAXM) Mc[(4 * j + 3) * THREAD_ROWS] = make_uint2(a[u].w, b[u].w);
                }
            }
        }
        const int mw = sp.mask_words;   // (the words past it are padding: zero)
        {
            const unsigned wd[17] = {mk[0].x, mk[0].y, mk[0].z, mk[0].w, mk[1].x, mk[1].y, mk[1].z, mk[1].w, mk[2].x,
                                     mk[2].y, mk[2].z, mk[2].w, mk[3].x, mk[3].y, mk[3].z, mk[3].w, 0u};
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < mw) bits[j * THREAD_ROWS] = __funnelshift_r(wd[j], j + 1 < mw ? wd[j + 1] : 0u, 1);
        }
        const int t_la = (int)il.y, t_lb = (int)il.z, t_ra = (int)ir.y, t_rb = (int)ir.z;
        l_da = (int)(il.w & 0xffu), l_db = (int)((il.w >> 8) & 0xffu), l_dl = (int)((il.w >> 16) & 0xffu);
        r_da = (int)(ir.w & 0xffu), r_db = (int)((ir.w >> 8) & 0xffu), r_dl = (int)((ir.w >> 16) & 0xffu);
        // ---- elements: top-down over the path set ----------------------------------------------------
        // the root: species 0, depth 0, on every root path; its "parent" is slot MAXM, which holds the infinite pair
        // during the cells pass (nothing is incomparable to the root)
        el[0] = (0xfu << TH_FLAG_SHIFT) | ((unsigned)MAXM << TH_PP_SHIFT);
        m = 1;
        for (int p = 0; p < m; ++p) {
            const unsigned w = el[p * THREAD_ROWS];
            const unsigned z = spt[w & TH_S_MASK];
            const int d = (int)((w >> TH_D_SHIFT) & 31u);
            const int c0 = (int)(z >> 16) - 1;
            if (c0 < 0) continue;
            const unsigned pair = children_on_path(c0), b0 = pair & 1u, b1 = pair >> 1;
            // (neither child on a path: an object leaf sits at this internal species)
            SR2_DCHECK(m + (int)(b0 + b1) <= MAXM);  // dtl_partition_kernel only lists rows that fit the tier
            el[p * THREAD_ROWS] = w | (b0 << TH_C0_BIT) | (b1 << TH_C1_BIT);
            const int split = (int)(z & 0xffffu);
            const unsigned first = (unsigned)(t_la < split) | ((unsigned)(t_lb < split) << 1) |
                                   ((unsigned)(t_ra < split) << 2) | ((unsigned)(t_rb < split) << 3);
            const unsigned flags = w >> TH_FLAG_SHIFT;
            const unsigned base = ((unsigned)(d + 1) << TH_D_SHIFT) | ((unsigned)p << TH_PP_SHIFT);
            const unsigned w0 = (unsigned)c0 | base | ((flags & first) << TH_FLAG_SHIFT) | (1u << TH_FIRST_BIT) | (b1 << TH_SIB_BIT);
            const unsigned w1 = (unsigned)(c0 + 1) | base | ((flags & ~first) << TH_FLAG_SHIFT) | ((b0 ^ 1u) << TH_FIRST_BIT) |
                                (b0 << TH_SIB_BIT);
            if (b0) el[(m++) * THREAD_ROWS] = w0;
            if (b1) el[(m++) * THREAD_ROWS] = w1;
        }
        SR2_DCHECK(m == (int)row_cnt[slot]);
    }
    // ---- child values -> keys ----------------------------------------------------------------------------
    {
        const unsigned vmul = 1u << kp.sh, infkey = 0xffffffffu << kp.sh;
        const int db = kp.db;
        // the entries of the stored children sit in Mc[0 .. m_l).x / Mc[0 .. m_r).y: walked backwards, like the
        // elements, entry q of a child matches an element e >= q, so writing the keys of element e never hits an
        // entry that is still to be read (the current entries are held in registers)
        int q_l = m_l - 1, q_r = m_r - 1;
        unsigned next_l = q_l >= 0 ? Mc[q_l * THREAD_ROWS].x : 0xffffffffu;
        unsigned next_r = q_r >= 0 ? Mc[q_r * THREAD_ROWS].y : 0xffffffffu;
        // Closed form of a cherry (a, b) at depth d of a's path: above the lca a duplication, at the lca a
        // speciation when both leaves sit below it, below the lca a transfer of b (if b is below the lca);
        // on b's path below the lca a transfer of a.  A leaf child is a cherry with a == b whose only finite
        // cell is the one at the lca (= a), 0.  All of it as selects on per-row constants (the lanes of a warp hold
        // rows of every kind); BIG stands for an infinite value: anything >= TH_VAL_MASK becomes the infinite key.
        // A flag says "on a's root path" down to a itself only (a leaf may sit at an internal species, whose
        // first-child chain inherits the flag): hence the depth tests.  On b's path but not on a's = below the lca.
        const int loss = cs.loss;
        const unsigned BIG = 0x40000000u;
        struct Side {
            unsigned k_top, k_lca, k_a, k_b;
            int dl, da, db;
        };
        auto make_side = [&](int mode, int da, int db_, int dl) {
            Side z;
            const bool a_below = da > dl, b_below = db_ > dl;
            const unsigned dup_top = (unsigned)(cs.dup + loss * (da + db_));                  // - 2 loss d
            z.dl = dl, z.da = da, z.db = db_;
            z.k_top = mode == 1 ? dup_top : BIG;
            z.k_lca = mode == 1 ? ((a_below && b_below) ? (unsigned)(loss * (da + db_ - 2)) : dup_top)
                                : (mode == 0 ? (unsigned)(2 * loss * dl) : BIG);                // - 2 loss d, at d == dl
            z.k_a = (mode == 1 && b_below) ? (unsigned)(cs.hgt + loss * da) : BIG;             // - loss d, a's path below the lca
            z.k_b = (mode == 1 && a_below) ? (unsigned)(cs.hgt + loss * db_) : BIG;            // - loss d, b's path below the lca
            return z;
        };
        const Side zl = make_side(mode_l, l_da, l_db, l_dl), zr = make_side(mode_r, r_da, r_db, r_dl);
        auto closed = [&](const Side& z, unsigned fa, unsigned fb, int d, unsigned ld) {
            const unsigned top = (d == z.dl ? z.k_lca : z.k_top) - 2u * ld;
            const unsigned va = d <= z.dl ? top : z.k_a - ld;
            const unsigned vb = z.k_b - ld;
            const bool on_a = fa && d <= z.da, on_b = fb && d <= z.db;
            return on_a ? va : (on_b ? vb : BIG);
        };
        for (int e = m - 1; e >= 0; --e) {
            const unsigned w = el[e * THREAD_ROWS];
            const unsigned s = w & TH_S_MASK;
            const int d = (int)((w >> TH_D_SHIFT) & 31u);
            const unsigned f = w >> TH_FLAG_SHIFT;
            const unsigned ld = (unsigned)(loss * d);
            unsigned va = closed(zl, f & 1u, f & 2u, d, ld), vb = closed(zr, f & 4u, f & 8u, d, ld);
            if (STORED && (next_l >> TH_VAL_BITS) == s) {   // (0xffffffff >> 22 = 1023 is no species: rank_bits <= 9)
                va = next_l & TH_VAL_MASK;
                q_l -= 1;
                next_l = q_l >= 0 ? Mc[q_l * THREAD_ROWS].x : 0xffffffffu;
            }
            if (STORED && (next_r >> TH_VAL_BITS) == s) {
                vb = next_r & TH_VAL_MASK;
                q_r -= 1;
                next_r = q_r >= 0 ? Mc[q_r * THREAD_ROWS].y : 0xffffffffu;
            }
            const unsigned low = (s << db) | (unsigned)d;
            Mc[e * THREAD_ROWS] = make_uint2(va >= TH_VAL_MASK ? infkey + low : va * vmul + low,
                                             vb >= TH_VAL_MASK ? infkey + low : vb * vmul + low);
        }
        SR2_DCHECK(q_l == -1 && q_r == -1);  // every entry of a child found its element
    }
    // ---- up-sweep: SUBMIN ---------------------------------------------------------------------------------
    for (int e = m - 1; e >= 1; --e) {
        const int pp = (int)((el[e * THREAD_ROWS] >> TH_PP_SHIFT) & TH_PP_MASK);
        const uint2 k = Mc[e * THREAD_ROWS], q = Mc[pp * THREAD_ROWS];
        Mc[pp * THREAD_ROWS] = make_uint2(min(q.x, k.x), min(q.y, k.y));
    }
    // ---- top-down: SEPMIN in place, and the cells; the warp walks in step so that the results of 8 elements
    //      x 32 rows leave through the staging tile as full sectors -------------------------------------------
    const int loss = cs.loss, db = kp.db, thresh = kp.thresh;
    const unsigned dm = kp.dm, rm = kp.rm, hi_mul = 1u << (32 - kp.sh);
    const int c12 = -2 * loss, c3 = cs.dup, c45 = cs.hgt;
    const uint2 inf2 = make_uint2(0xffffffffu, 0xffffffffu);
    // Results leave as 16-byte stores, four entries of the lane's own row at a time (the two halves of a sector
    // are written a few dozen instructions apart and merge in L2): no staging tile, no flush.
    uint32_t* __restrict__ Tq = reinterpret_cast<uint32_t*>(T) + slot * (size_t)pitch;
    uint32_t* __restrict__ Gq = tags + slot * (size_t)pitch;
    uint2 prevM = inf2;  // SUBMIN of the element before (its slot already holds its SEPMIN)
    unsigned w = m > 0 ? el[0] : 0u;
    int fc = 1;  // position of the first on-path child of the element at hand: a running count in BFS order
    Mc[MAXM * THREAD_ROWS] = inf2;  // the root's "parent"
    for (int e0 = 0; e0 < m; e0 += 4) {
        unsigned vq[4] = {0u, 0u, 0u, 0u}, gq[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u;
            if (e < m) {
                const unsigned wn = el[(e + 1) * THREAD_ROWS];  // (past m, or past el into the bit set: unused)
                const int pp = (int)((w >> TH_PP_SHIFT) & TH_PP_MASK);
                const uint2 Pp = Mc[pp * THREAD_ROWS], mm = Mc[e * THREAD_ROWS], Mn = Mc[(e + 1) * THREAD_ROWS];
                // a sibling on a path is the next element (this one is the first child) or the previous one
                const bool both = ((w >> TH_SIB_BIT) & 1u) != 0u;
                const bool is_first = ((w >> TH_FIRST_BIT) & 1u) != 0u;
                uint2 Ms = is_first ? Mn : prevM;
                if (!both) Ms = inf2;
                const uint2 pq = make_uint2(min(Pp.x, Ms.x), min(Pp.y, Ms.y));
                // the cell
                const unsigned x = w & TH_S_MASK;
                const int dx = (int)((w >> TH_D_SHIFT) & 31u);
                const unsigned h0 = (w >> TH_C0_BIT) & 1u, h1 = (w >> TH_C1_BIT) & 1u;
                uint2 k0 = inf2, k1 = inf2;  // SUBMIN at the two children of x
                if (h0) k0 = Mc[fc * THREAD_ROWS];
                if (h1) k1 = Mc[(fc + (int)h0) * THREAD_ROWS];
                fc += (int)(h0 + h1);
                const int nld = -loss * dx;
                const int qMl = flat2_q(mm.x, hi_mul, dm, loss), qMr = flat2_q(mm.y, hi_mul, dm, loss);
                const int ql0 = flat2_q(k0.x, hi_mul, dm, loss), qr0 = flat2_q(k0.y, hi_mul, dm, loss);
                const int ql1 = flat2_q(k1.x, hi_mul, dm, loss), qr1 = flat2_q(k1.y, hi_mul, dm, loss);
                const int pl = (int)__umulhi(pq.x, hi_mul), pr = (int)__umulhi(pq.y, hi_mul);
                const int b12 = 2 * nld + c12, b3 = 2 * nld + c3, b45 = nld + c45;
                const int w1 = flat2_pack(ql0 + qr1 + b12, 0);  // 1. speciation, left below x0, right below x1
                const int w2 = flat2_pack(ql1 + qr0 + b12, 1);  // 2. the other way round
                const int w3 = flat2_pack(qMl + qMr + b3, 2);   // 3. duplication
                const int w4 = flat2_pack(pl + qMr + b45, 3);   // 4. transfer, left child away
                const int w5 = flat2_pack(qMl + pr + b45, 4);   // 5. transfer, right child away
                const int wb = min(min(min(w1, w2), w3), min(w4, w5));  // smallest value, then first candidate
                const int best = wb >> 3, which = wb & 7;
                unsigned ka = mm.x, kb = mm.y;
                ka = which == 0 ? k0.x : ka;
                ka = which == 1 ? k1.x : ka;
                ka = which == 3 ? pq.x : ka;
                kb = which == 0 ? k1.y : kb;
                kb = which == 1 ? k0.y : kb;
                kb = which == 4 ? pq.y : kb;
                const bool finite = best < thresh;
                SR2_DCHECK(!finite || (unsigned)best < TH_VAL_MASK);
                vq[u] = (finite ? (unsigned)best : TH_VAL_MASK) | (x << TH_VAL_BITS);
                gq[u] = finite ? (((ka >> db) & rm) | (((kb >> db) & rm) << 16)) : SR2_NO_TAG;
                Mc[e * THREAD_ROWS] = pq;
                prevM = mm;
                w = wn;
            }
        }
        // (entries past the row's size, up to the quad's end, are never read)
        *reinterpret_cast<uint4*>(Tq + e0) = make_uint4(vq[0], vq[1], vq[2], vq[3]);
        *reinterpret_cast<uint4*>(Gq + e0) = make_uint4(gq[0], gq[1], gq[2], gq[3]);
    }
}

// Root-path bit set of every row of a wave computed by a kernel that does not maintain it itself:
// mask(row) = mask(left child) | mask(right child), a leaf child contributing its species' path.
__global__ void dtl_rowmask_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                                   int64_t row_begin, int64_t chunk_row0, int n_rows, uint32_t* __restrict__ rowmask,
                                   int cherry_rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ms = sp.mask_stride, quads = ms >> 2;  // one uint4 per thread
    if (i >= (int64_t)n_rows * quads) return;
    const int64_t gr = row_begin + i / quads;
    const int w = (int)(i % quads);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int c = side ? row_cr[gr] : row_cl[gr];
        if (c >= cherry_rows) {
            const uint4 a = reinterpret_cast<const uint4*>(rowmask + (size_t)c * ms)[w];
            o = make_uint4(o.x | a.x, o.y | a.y, o.z | a.z, o.w | a.w);
        } else {  // a leaf, or a cherry (whose set is not stored): root paths of the species
            const int s0 = c < 0 ? -1 - c : -1 - row_cl[chunk_row0 + c];
            const int s1 = c < 0 ? s0 : -1 - row_cr[chunk_row0 + c];
            const uint4 a = reinterpret_cast<const uint4*>(sp.pathmask + (size_t)s0 * ms)[w];
            const uint4 b = reinterpret_cast<const uint4*>(sp.pathmask + (size_t)s1 * ms)[w];
            o = make_uint4(o.x | a.x | b.x, o.y | a.y | b.y, o.z | a.z | b.z, o.w | a.w | b.w);
        }
    }
    reinterpret_cast<uint4*>(rowmask + (size_t)(gr - chunk_row0) * ms)[w] = o;
}

// the instantiation whose register-resident schedule is the shortest that holds n_steps
typedef void (*Flat2Fn)(DevSpecies, const int32_t*, const int32_t*, int64_t, int64_t, int, int32_t*, uint32_t*,
                        DtlCosts, Key32, const uint32_t*, int, int, int, const int*, const int*, int*, int, const uint8_t*, int);
template <bool CHAIN>
static Flat2Fn flat2_pick(int n_steps) {
    if (n_steps <= 12) return dtl_wave_flat2_kernel<12, CHAIN>;
    if (n_steps <= 16) return dtl_wave_flat2_kernel<16, CHAIN>;
    if (n_steps <= 20) return dtl_wave_flat2_kernel<20, CHAIN>;
    if (n_steps <= 24) return dtl_wave_flat2_kernel<24, CHAIN>;
    if (n_steps <= 28) return dtl_wave_flat2_kernel<28, CHAIN>;
    return dtl_wave_flat2_kernel<32, CHAIN>;
}

// Event cost of node (x; children at a, b) as ReconciliationOutput.cost() sees it
// (/root/reference/src/superrec2/model/reconciliation.py:273-365).  SR2_INF = INVALID.
__device__ __forceinline__ int dtl_event_cost(const DevSpecies& sp, int x, int a, int b,
                                              const DtlCosts& cs) {
    int tx = __ldg(sp.tin + x), ox = __ldg(sp.tout + x), dx = __ldg(sp.depth + x);
    int ta = __ldg(sp.tin + a), oa = __ldg(sp.tout + a), da = __ldg(sp.depth + a);
    int tb = __ldg(sp.tin + b), ob = __ldg(sp.tout + b), db = __ldg(sp.depth + b);
    bool x_anc_a = tx <= ta && ta < ox, x_anc_b = tx <= tb && tb < ox;
    bool a_sanc_x = a != x && ta <= tx && tx < oa, b_sanc_x = b != x && tb <= tx && tx < ob;
    if (a_sanc_x || b_sanc_x) return SR2_INF;
    if (x_anc_a && x_anc_b) {
        // speciation iff a and b sit below different children of x
        int c = __ldg(sp.child0 + x);
        bool spec = false;
        if (c >= 0 && a != x && b != x) {
            int t1 = __ldg(sp.tin + c + 1);  // subtree(c) = [tx+1, t1), subtree(c+1) = [t1, ox)
            spec = (ta < t1) != (tb < t1);
        }
        int dist = da + db - 2 * dx;
        return spec ? cs.spe + cs.loss * (dist - 2) : cs.dup + cs.loss * dist;
    }
    if (x_anc_a) return cs.hgt + cs.loss * (da - dx);
    if (x_anc_b) return cs.hgt + cs.loss * (db - dx);
    return SR2_INF;
}

// Decoded-cost rows: C[u][x] = cost() of the solution decoded from (u, x)
//   = event(x, sl, sr) + C[l][sl] + C[r][sr]      with (sl, sr) = tag[u][x].
// Only needed when cost() can differ from the table value (speciation cost != 0,
// see DESIGN.md "selection"); one warp per row.
__global__ void __launch_bounds__(256)
dtl_decoded_cost_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl,
                        const int32_t* __restrict__ row_cr, int64_t row_begin, int64_t chunk_row0,
                        int n_rows, const uint32_t* __restrict__ tags, int32_t* __restrict__ C,
                        DtlCosts cs, const int32_t* __restrict__ T) {
    int local = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (local >= n_rows) return;
    const int64_t gr = row_begin + local;
    const size_t slot = (size_t)(gr - chunk_row0);
    const int cl = row_cl[gr], cr = row_cr[gr];
    const bool cherry = cl < 0 && cr < 0;  // height-1 rows store no tags
    for (int x = lane; x < sp.S; x += 32) {
        // an infinite cell has no tag (and its tag word may never have been written)
        const bool finite = T[slot * sp.pitch + x] < SR2_INF;
        unsigned tag = !finite ? SR2_NO_TAG
                       : cherry ? (unsigned)(-1 - cl) | ((unsigned)(-1 - cr) << 16)
                                : tags[slot * sp.pitch + x];
        int out = SR2_INF;
        if (tag != SR2_NO_TAG) {
            int a = tag & 0xffffu, b = tag >> 16;
            int ca = cl >= 0 ? C[(size_t)cl * sp.pitch + a] : (a == -1 - cl ? 0 : SR2_INF);
            int cb = cr >= 0 ? C[(size_t)cr * sp.pitch + b] : (b == -1 - cr ? 0 : SR2_INF);
            int ev = dtl_event_cost(sp, x, a, b, cs);
            if (ca < SR2_INF && cb < SR2_INF && ev < SR2_INF) out = ev + ca + cb;
        }
        C[slot * sp.pitch + x] = out;
    }
}

// Selection (reconcile_thl:284-294): roots are tried in BFS order and the first
// strict minimum of cost() wins.  One warp per instance.
__global__ void __launch_bounds__(256)
dtl_select_kernel(DevSpecies sp, const int64_t* __restrict__ root_row,
                  const int32_t* __restrict__ root_node, int inst_begin, int n_inst,
                  int64_t chunk_row0, const int32_t* __restrict__ T, const int32_t* __restrict__ C,
                  int32_t* __restrict__ min_cost, int32_t* __restrict__ root_species,
                  int32_t* __restrict__ map) {
    int local = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (local >= n_inst) return;
    int b = inst_begin + local;
    int64_t gr = root_row[b];
    if (gr < 0) {  // single-leaf object tree: the leaf sits at its own species at cost 0
        if (lane == 0) {
            int s = (int)(-1 - gr);
            min_cost[b] = 0;
            root_species[b] = s;
            map[root_node[b]] = s;
        }
        return;
    }
    size_t slot = (size_t)(gr - chunk_row0);
    u64 best = SR2_KEY_NONE;
    for (int x = lane; x < sp.S; x += 32) {
        int t = T[slot * sp.pitch + x];
        if (t < SR2_INF) {  // roots without an entry are skipped (Hazard B)
            int c = C ? C[slot * sp.pitch + x] : t;
            best = key_min(best, ((u64)(unsigned)c << 32) | (unsigned)x);
        }
    }
    for (int off = 16; off; off >>= 1) best = key_min(best, __shfl_xor_sync(0xffffffffu, best, off));
    if (lane == 0) {
        bool none = best == SR2_KEY_NONE;
        int x = none ? -1 : (int)(best & 0xffffffffu);
        min_cost[b] = none ? SR2_INF : (int)(best >> 32);
        root_species[b] = x;
        map[root_node[b]] = x;
    }
}

// Traceback, one height per launch from the top down: a node of height h reads its
// own species (set by its parent, which is higher) and assigns its two children.
__global__ void __launch_bounds__(256)
dtl_traceback_kernel(DevSpecies sp, const int32_t* __restrict__ row_node,
                     const int32_t* __restrict__ row_left, const int32_t* __restrict__ row_right,
                     int64_t row_begin, int64_t chunk_row0, int n_rows,
                     const uint32_t* __restrict__ tags, int32_t* __restrict__ map,
                     const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                     const int32_t* __restrict__ T) {
    int local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= n_rows) return;
    int64_t gr = row_begin + local;
    size_t slot = (size_t)(gr - chunk_row0);
    int x = map[row_node[gr]];
    int a = -1, b = -1;
    const int cl = row_cl[gr], cr = row_cr[gr];
    if (x >= 0) {
        if (cl < 0 && cr < 0) {
            // a cherry stores no tags: its parent only ever points at a finite cell, whose tag is
            // the two leaf species
            if (T[slot * sp.pitch + x] < SR2_INF) a = -1 - cl, b = -1 - cr;
        } else if (T[slot * sp.pitch + x] < SR2_INF) {  // infinite cells have no (defined) tag
            unsigned tag = tags[slot * sp.pitch + x];
            if (tag != SR2_NO_TAG) {
                a = tag & 0xffffu;
                b = tag >> 16;
            }
        }
    }
    map[row_left[gr]] = a;
    map[row_right[gr]] = b;
}

// The traceback touches every row once, at random: read from four row arrays that is four 32-byte sectors per
// node for 16 bytes of use (the kernel moves ~190 B per node and runs at the random-access rate of the HBM:
// 0.94 GB in 0.36 ms for config #2).  This packs them, one 16-byte record per row (side stream, once per solve).
__global__ void __launch_bounds__(256)
dtl_desc_kernel(const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                const int32_t* __restrict__ row_left, const int32_t* __restrict__ row_right, int64_t chunk_row0,
                int64_t n_rows, int4* __restrict__ desc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int64_t gr = chunk_row0 + i;
    desc[i] = make_int4(row_cl[gr], row_cr[gr], row_left[gr], row_right[gr]);
}

// Selection and traceback of a whole object tree in one CTA (the default): after the root is chosen
// as in dtl_select_kernel, the tree is walked breadth-first with the frontier -- (row, species) pairs of
// the internal nodes of one depth -- in shared memory, so a level costs one round trip to the tables
// and a CTA barrier instead of a launch over the rows of one height of every tree.
// Shared memory: 2 lists of `cap` pairs; cap >= the internal nodes of one depth (<= leaves / 2).
__global__ void __launch_bounds__(1024)
dtl_select_trace_kernel(DevSpecies sp, const int64_t* __restrict__ root_row, const int32_t* __restrict__ root_node,
                        int inst_begin, int64_t chunk_row0, const int32_t* __restrict__ T,
                        const int32_t* __restrict__ C, const uint32_t* __restrict__ tags,
                        const int32_t* __restrict__ row_left, const int32_t* __restrict__ row_right,
                        const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr, int cap, int check_T,
                        int cherry_end, DtlCosts cs, int32_t* __restrict__ min_cost, int32_t* __restrict__ root_species,
                        int32_t* __restrict__ map, const uint8_t* __restrict__ row_cnt,
                        const uint32_t* __restrict__ rowmask, const int4* __restrict__ desc) {
    extern __shared__ int2 trace_lists[];  // [2][cap] of (row slot, species)
    __shared__ u64 warp_best[32];
    __shared__ int n_list[2];
    const int b = inst_begin + blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int64_t gr0 = root_row[b];
    if (gr0 < 0) {  // single-leaf object tree: the leaf sits at its own species at cost 0
        if (threadIdx.x == 0) {
            const int s = (int)(-1 - gr0);
            min_cost[b] = 0;
            root_species[b] = s;
            map[root_node[b]] = s;
        }
        return;
    }
    const int pitch = sp.pitch;
    const size_t slot0 = (size_t)(gr0 - chunk_row0);
    u64 best = SR2_KEY_NONE;
    if ((int64_t)slot0 < cherry_end) {  // a two-leaf tree whose row was never written (virtual cherries)
        const CherryForm f = cherry_form(sp, -1 - row_cl[gr0], -1 - row_cr[gr0]);
        for (int which = 0; which < 2; ++which) {
            const uint16_t* path = which ? f.pb : f.pa;
            const int last = which ? f.db : f.da;
            for (int d = (which ? f.dl + 1 : 0) + threadIdx.x; d <= last; d += blockDim.x) {
                const unsigned t = cherry_path_value(f, which, d, cs);
                if (t != 0xffffffffu) best = key_min(best, ((u64)t << 32) | (unsigned)path[d]);
            }
        }
    } else if (row_cnt && row_cnt[slot0]) {  // a compact root row (small trees): entries = value | species << 22
        const int m0 = (int)row_cnt[slot0];
        for (int e = threadIdx.x; e < m0; e += blockDim.x) {
            const unsigned v = (unsigned)T[slot0 * pitch + e], t = v & TH_VAL_MASK;
            if (t != TH_VAL_MASK) best = key_min(best, ((u64)t << 32) | (v >> TH_VAL_BITS));
        }
    } else {
        for (int x = threadIdx.x; x < sp.S; x += blockDim.x) {
            const int t = T[slot0 * pitch + x];
            if (t < SR2_INF) {  // roots without an entry are skipped (Hazard B)
                const int c = C ? C[slot0 * pitch + x] : t;
                best = key_min(best, ((u64)(unsigned)c << 32) | (unsigned)x);
            }
        }
    }
    for (int off = 16; off; off >>= 1) best = key_min(best, __shfl_xor_sync(0xffffffffu, best, off));
    if (lane == 0) warp_best[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < n_warps; ++w) best = key_min(best, warp_best[w]);
        const bool none = best == SR2_KEY_NONE;
        const int x = none ? -1 : (int)(best & 0xffffffffu);
        min_cost[b] = none ? SR2_INF : (int)(best >> 32);
        root_species[b] = x;
        map[root_node[b]] = x;
        trace_lists[0] = make_int2((int)slot0, x);
        n_list[0] = 1;
        n_list[1] = 0;
    }
    __syncthreads();
    int cur = 0;
    for (;;) {
        const int n = min(n_list[cur], cap);
        if (n == 0) break;
        const int2* from = trace_lists + (size_t)cur * cap;
        int2* to = trace_lists + (size_t)(cur ^ 1) * cap;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int slot = from[i].x, x = from[i].y;
            const int64_t gr = chunk_row0 + slot;
            int cl, cr, node_l, node_r;
            if (desc) {
                const int4 dd = desc[slot];
                cl = dd.x, cr = dd.y, node_l = dd.z, node_r = dd.w;
            } else {
                cl = row_cl[gr], cr = row_cr[gr], node_l = row_left[gr], node_r = row_right[gr];
            }
            int a = -1, c = -1;
            // infinite cells have no (defined) tag; a cherry stores none at all: its parent only ever
            // points at a finite cell, whose tag is the two leaf species
            // (a parent's tag only ever points at a finite cell of its child: the root is finite, and a
            // finite value is a sum of the finite child values at the tagged species)
            if (x >= 0) {
                if (cl < 0 && cr < 0) {
                    a = -1 - cl;
                    c = -1 - cr;
                } else if (row_cnt && row_cnt[slot]) {
                    // a compact row: the entry of species x is the number of path species before it
                    const uint32_t* mk = rowmask + (size_t)slot * sp.mask_stride;
                    const int xw = x >> 5;
                    int pos = 0;
                    const unsigned below = (1u << (x & 31)) - 1u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {  // at most 16 words; the loads are independent
                        if (q * 4 > xw) continue;
                        const uint4 v = reinterpret_cast<const uint4*>(mk)[q];
                        const unsigned wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int wi = q * 4 + i;
                            pos += wi < xw ? __popc(wd[i]) : (wi == xw ? __popc(wd[i] & below) : 0);
                        }
                    }
                    SR2_DCHECK(pos < (int)row_cnt[slot]);
                    SR2_DCHECK(((unsigned)T[(size_t)slot * pitch + pos] >> TH_VAL_BITS) == (unsigned)x);
                    if (!check_T || ((unsigned)T[(size_t)slot * pitch + pos] & TH_VAL_MASK) != TH_VAL_MASK) {
                        const unsigned tag = tags[(size_t)slot * pitch + pos];
                        if (tag != SR2_NO_TAG) {
                            a = tag & 0xffffu;
                            c = tag >> 16;
                        }
                    }
                } else if (!check_T || T[(size_t)slot * pitch + x] < SR2_INF) {
                    const unsigned tag = tags[(size_t)slot * pitch + x];
                    if (tag != SR2_NO_TAG) {
                        a = tag & 0xffffu;
                        c = tag >> 16;
                    }
                }
            }
            map[node_l] = a;
            map[node_r] = c;
            // (validated full binary trees never exceed cap; the guard keeps a bad tree out of other memory)
            if (cl >= 0) {
                const int k = atomicAdd(&n_list[cur ^ 1], 1);
                if (k < cap) to[k] = make_int2(cl, a);
            }
            if (cr >= 0) {
                const int k = atomicAdd(&n_list[cur ^ 1], 1);
                if (k < cap) to[k] = make_int2(cr, c);
            }
            // what the next level reads first about the two children, pulled towards the SM while this level
            // finishes: descriptors, and for compact rows the size and the bit set (the entry's position)
            if (row_cnt) {
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const int ch = side ? cr : cl, xs = side ? c : a;
                    if (ch < 0 || xs < 0) continue;
                    if (desc) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(desc + ch));
                    } else {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(row_cl + chunk_row0 + ch));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(row_cr + chunk_row0 + ch));
                    }
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(row_cnt + ch));
                    const uint32_t* mk2 = rowmask + (size_t)ch * sp.mask_stride;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(mk2));
                    if ((xs >> 5) >= 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(mk2 + 8));
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) n_list[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int sr2_check_costs(sr2_ctx* ctx, const char* who, const int32_t costs[5], int64_t max_nodes,
                    int64_t* out_bound) {
    for (int k = 0; k < 5; ++k)
        if (costs[k] < 0 || costs[k] > (1 << 20))
            return sr2_fail(ctx, SR2_ERR_RANGE, std::string(who) + ": event costs must be integers in [0, 2^20]");
    // finite table values are bounded by (#internal nodes) * (largest event term)
    int64_t ev = std::max<int64_t>({costs[0], costs[1], costs[2]}) + 2 * (int64_t)costs[3] * ctx->D +
                 2 * (int64_t)costs[4];
    if (max_nodes * ev >= (1 << 29))
        return sr2_fail(ctx, SR2_ERR_RANGE, std::string(who) + ": costs x tree size may overflow the int32 DP table");
    if (out_bound) *out_bound = max_nodes * ev;
    return SR2_OK;
}

// Tiers of dtl_wave_thread_kernel and the classes of dtl_partition_kernel (DtlProblem::tier_*, n_classes);
// `table` (2 x 1024 bytes: class of a path-set size without / with compact rows) is filled when given.
static void dtl_thread_tiers(sr2_ctx* ctx, DtlProblem* p, std::vector<uint8_t>* table) {
    // Default tiers: the largest path sets that still fit 8 / 6 / 5 / 4 CTAs of the thread kernel on an SM (228 KB
    // of shared memory, 1 KB of it reserved per CTA): 29 / 42 / 52 / 64 species for config #2 (399 species: 13 words
    // of bit set per row, 1.6 KB of species table per CTA).  Below 4 CTAs per SM the kernel loses to flat2.
    std::vector<int> tiers;
    for (int k : {8, 6, 5, 4}) {
        const long fixed = 1024 + (long)THREAD_SMEM_BYTES(0, ctx->dsp.mask_words, ctx->S);
        const long fit = ((long)(228 * 1024) / k - fixed) / (long)(3 * THREAD_ROWS * sizeof(uint32_t));
        tiers.push_back((int)std::max<long>(8, std::min<long>(64, fit)));
    }
    if (const char* tv = getenv("SR2_THREAD_TIERS")) {
        tiers.clear();
        for (const char* q = tv; *q && tiers.size() < 8;) {
            tiers.push_back(std::max(1, std::min(THREAD_MAXM, atoi(q))));
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        if (tiers.empty()) tiers.push_back(64);
    }
    std::sort(tiers.begin(), tiers.end());
    tiers.erase(std::unique(tiers.begin(), tiers.end()), tiers.end());
    const int nt = (int)tiers.size();
    int cls = 0, k = 0;
    p->n_tiers = nt;
    p->tier_cls0[0] = 0;
    for (int m = 0; m < 1024; ++m) {
        if (table)
            (*table)[m] = (uint8_t)(p->sparse_maxm <= 0 ? 2 : (m <= p->sparse_small ? 0 : (m <= p->sparse_maxm ? 1 : 2)));
        if (k >= nt) {
            if (table) (*table)[1024 + m] = (uint8_t)p->tier_cls0[nt];
            continue;
        }
        if (table) (*table)[1024 + m] = (uint8_t)cls;
        if (m == tiers[k]) {
            p->tier_maxm[k] = tiers[k];
            p->tier_cls0[++k] = ++cls;
        } else if (m >= 16 && (m <= 40 ? m % 2 == 0 : m % 4 == 0) && cls + (nt - k) + 2 < DTL_MAX_CLASSES) {
            ++cls;
        }
    }
    p->n_classes[0] = 3;
    p->n_classes[1] = p->tier_cls0[nt] + 1;
}

// Everything of an upload that needs the plan of the batch (chunks, sizes) but not its rows:
// cost range check, table buffers.  Runs as the on_plan hook of sr2_build_rows.
static int dtl_prepare(void* user) {
    sr2_ctx* ctx = (sr2_ctx*)user;
    DtlProblem* p = ctx->dtl;
    const size_t B = p->rows.B, N = p->rows.N;
    int rc = sr2_check_costs(ctx, "sr2_dtl_upload", p->costs, p->rows.max_nodes, &p->max_finite);
    if (rc) return rc;
    if ((p->flags & SR2_KEEP_TABLES) && p->rows.chunks.size() > 1)
        return sr2_fail(ctx, SR2_ERR_NOMEM, "sr2_dtl_upload: SR2_KEEP_TABLES needs the batch to fit in one chunk");
    size_t table_elems = (size_t)std::max<int64_t>(p->rows.max_chunk_rows, 1) * ctx->pitch;
    if (p->two_streams) {
        table_elems *= 2;
        if (!ctx->stream2) SR2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    }
    void* ptr = nullptr;
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_MINCOST, std::max<size_t>(B, 1) * sizeof(int32_t), &ptr))) return rc;
    p->d_min_cost = (int32_t*)ptr;
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_ROOTSP, std::max<size_t>(B, 1) * sizeof(int32_t), &ptr))) return rc;
    p->d_root_species = (int32_t*)ptr;
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_MAP, std::max<size_t>(N, 1) * sizeof(int32_t), &ptr))) return rc;
    p->d_map = (int32_t*)ptr;
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_T, table_elems * sizeof(int32_t), &ptr))) return rc;
    p->d_T = (int32_t*)ptr;
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_TAG, table_elems * sizeof(uint32_t), &ptr))) return rc;
    p->d_tag = (uint32_t*)ptr;
    if (p->need_decoded_cost) {
        if ((rc = sr2_dev_reserve(ctx, DP_DTL_C, table_elems * sizeof(int32_t), &ptr))) return rc;
        p->d_C = (int32_t*)ptr;
    }
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_MASK, table_elems / ctx->pitch * ctx->dsp.mask_stride * sizeof(uint32_t) + 256, &ptr))) return rc;
    p->d_mask = (uint32_t*)ptr;
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_ROWCNT, table_elems / ctx->pitch + 256, &ptr))) return rc;
    p->d_rowcnt = (uint8_t*)ptr;
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_ROWINFO, (table_elems / ctx->pitch) * 2 * sizeof(uint4) + 256, &ptr))) return rc;
    p->d_rowinfo = (uint4*)ptr;
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_DESC, (table_elems / ctx->pitch) * sizeof(int4) + 256, &ptr))) return rc;
    p->d_desc = (int4*)ptr;
    p->list_pitch = std::max<int64_t>(p->rows.R, 1);
    dtl_thread_tiers(ctx, p, nullptr);
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_LISTS, (size_t)std::max(p->n_classes[0], p->n_classes[1]) * p->list_pitch * sizeof(int), &ptr))) return rc;
    p->d_lists = (int*)ptr;
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_COUNTS, p->rows.chunks.size() * DTL_MAX_CLASSES * 65536 * sizeof(int), &ptr))) return rc;
    p->d_counts = (int*)ptr;
    if ((rc = sr2_dev_reserve(ctx, DP_DTL_CHAIN, (size_t)2 * (std::max<int64_t>(p->rows.max_chunk_rows, 1) + 1) * sizeof(int), &ptr)))
        return rc;
    p->d_chain_sync = (int*)ptr;
    for (auto* list : {&ctx->dtl_ev_w0, &ctx->dtl_ev_w1, &ctx->dtl_ev_done})
        while (list->size() < p->rows.chunks.size()) {
            cudaEvent_t e;
            SR2_CUDA(ctx, cudaEventCreate(&e));
            list->push_back(e);
        }
    while (ctx->dtl_ev_copied.size() < p->rows.chunks.size()) {
        cudaEvent_t e;
        SR2_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->dtl_ev_copied.push_back(e);
    }
    p->prepared = true;
    return SR2_OK;
}

// pipeline_chunks > 1: split the batch into about that many instance chunks even when it would fit
// in one, so that sr2_dtl_run can solve chunk k on the device while the host buckets chunk k+1.
static int dtl_upload_impl(sr2_ctx* ctx, int32_t B, const int64_t* node_off, const int32_t* left,
                           const int32_t* right, const int32_t* leaf_species, const int32_t costs[5],
                           uint32_t flags, int pipeline_chunks, const RowHooks* hooks) {
    if (!ctx) return SR2_ERR_ARG;
    if (ctx->policy == SR2_POLICY_ALL) flags |= SR2_KEEP_TABLES;  // sr2_set_policy
    double t0 = now_ms();
    if (ctx->S == 0) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_dtl_upload: call sr2_set_species first");
    const bool raw16 = hooks && hooks->left16;
    if (B < 0 || !node_off || !costs || (B > 0 && !raw16 && (!left || !right || !leaf_species)) ||
        (B > 0 && raw16 && (!hooks->right16 || !hooks->leaf16)))
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_dtl_upload: null array");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    sr2_free_dtl(ctx);
    const int64_t N = B ? node_off[B] : 0;

    DtlProblem* p = new DtlProblem();
    ctx->dtl = p;
    p->flags = flags;
    memcpy(p->costs, costs, sizeof(p->costs));
    // cost() == table value unless speciations are charged: the table never adds the
    // speciation cost (reconciliation.py:97-108) but cost() does (DESIGN.md "selection").
    p->need_decoded_cost = costs[0] != 0;

    size_t free_b = 0;
    {
        int rc_mem = sr2_mem_free(ctx, &free_b);
        if (rc_mem) return rc_mem;
    }
    // buffers this context already holds are reused, so they count as available
    for (int id : {DP_ROWS_SLAB, DP_DTL_T, DP_DTL_TAG, DP_DTL_C, DP_DTL_MAP, DP_DTL_MASK}) free_b += sr2_dev_pooled(ctx, id);
    size_t per_row = (size_t)ctx->pitch * (p->need_decoded_cost ? 12 : 8) + ctx->pitch / 8 + 16 + 4 * DTL_MAX_CLASSES + 40;  // tables, path set, lists, records and flags
    size_t fixed = (size_t)N * 4 * 7 + (size_t)B * 32 + (64u << 20);
    if (free_b < fixed + per_row * 4)
        return sr2_fail(ctx, SR2_ERR_NOMEM, "sr2_dtl_upload: batch does not fit in device memory");
    int64_t rows_budget = (int64_t)(((free_b - fixed) / 10 * 9) / per_row);
    int64_t first_chunk_rows = 0;
    const bool resident_split = !hooks && pipeline_chunks > 1;
    p->two_streams = pipeline_chunks > 1 && !(flags & SR2_KEEP_TABLES) && ((hooks && hooks->on_chunk) || resident_split);
    if (p->two_streams) rows_budget /= 2;  // two sets of tables
    if (pipeline_chunks > 1 && !(flags & SR2_KEEP_TABLES)) {
        int64_t rows_est = (N - B) / 2;
        int64_t per_chunk = (rows_est + pipeline_chunks - 1) / pipeline_chunks;
        // a chunk should still fill the machine's big-wave kernels
        per_chunk = std::max<int64_t>(per_chunk, 64 * 8 * (int64_t)ctx->sm_count);
        rows_budget = std::min(rows_budget, per_chunk);
        int first_div = resident_split ? 1 : 2;
        if (const char* fd = getenv("SR2_PIPELINE_FIRST")) first_div = std::max(1, atoi(fd));
        first_chunk_rows = rows_budget / first_div;
    }
    RowHooks plan_only;
    plan_only.user = ctx;
    plan_only.on_plan = dtl_prepare;
    // pipelined runs of big batches bucket their rows on the device (one thread per tree)
    if (p->two_streams && B >= 1024 && hooks) {
        const char* dev_rows = getenv("SR2_DEVICE_ROWS");
        if (!(dev_rows && dev_rows[0] == '0')) {
            size_t raw = (size_t)N * 22 + (size_t)B * 520;  // raw arrays, heights, node rows, per-instance bases
            int64_t budget = rows_budget;
            if (free_b > fixed + raw + per_row * 8) {
                budget = std::min(budget, (int64_t)(((free_b - fixed - raw) / 10 * 9) / per_row / 2));
                if (getenv("SR2_TRACE")) fprintf(stderr, "[sr2] dtl upload: %.2f ms before the row builder\n", now_ms() - t0);
                int rc = sr2_build_rows_device(ctx, "sr2_dtl_run", B, node_off, left, right, leaf_species, budget,
                                               &p->rows, hooks, first_chunk_rows);
                if (rc != SR2_ROWS_FALLBACK) {
                    if (!rc) ctx->timing.upload_ms = (float)(now_ms() - t0);
                    return rc;
                }
                p->prepared = false;  // a tree too high for the device tables: bucket on the host
            }
        }
    }
    std::vector<int32_t> wide;   // sr2_dtl_run16 off the device path (small batches, very high trees)
    if (hooks && hooks->left16) {
        wide.resize((size_t)N * 3);
        const int16_t* from[3] = {hooks->left16, hooks->right16, hooks->leaf16};
        for (int k = 0; k < 3; ++k)
            for (int64_t i = 0; i < N; ++i) wide[(size_t)k * N + i] = from[k][i];
        left = wide.data();
        right = wide.data() + N;
        leaf_species = wide.data() + 2 * N;
    }
    int rc = sr2_build_rows(ctx, "sr2_dtl_upload", B, node_off, left, right, leaf_species, rows_budget,
                            (flags & SR2_KEEP_TABLES) != 0, false, &p->rows, hooks ? hooks : &plan_only,
                            first_chunk_rows);
    if (rc) return rc;
    ctx->timing.upload_ms = (float)(now_ms() - t0);
    return SR2_OK;
}

extern "C" int sr2_dtl_upload(sr2_ctx* ctx, int32_t B, const int64_t* node_off,
                              const int32_t* left, const int32_t* right,
                              const int32_t* leaf_species, const int32_t costs[5], uint32_t flags) {
    // SR2_DTL_SOLVE_CHUNKS=n: a resident batch in n instance chunks that alternate between two streams and table
    // sets, so that the latency-bound top waves and the traceback of one chunk overlap the big waves of the next
    int chunks = 1;
    if (const char* sc = getenv("SR2_DTL_SOLVE_CHUNKS")) chunks = std::max(1, atoi(sc));
    return dtl_upload_impl(ctx, B, node_off, left, right, leaf_species, costs, flags, chunks, nullptr);
}

static int dtl_configure_waves(sr2_ctx* ctx, DtlProblem* p) {
    const int S = ctx->S;
    p->spad = (S + 1) & ~1;
    const size_t per_row = (size_t)4 * p->spad * sizeof(u64);   // Ml, Mr, Pl, Pr
    const size_t topology = dtl_topology_bytes(S, ctx->D, ctx->n_internal);
    if (S <= 1024) {
        p->cta_per_row = false;
        p->wave_warps = (int)std::max<size_t>(1, std::min<size_t>(8, (ctx->smem_optin - 1024 - topology) / per_row));
        p->wave_smem = per_row * p->wave_warps + topology;
        SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
    } else {
        p->cta_per_row = true;
        p->wave_warps = 16;
        p->wave_inplace = per_row + topology > ctx->smem_optin;
        if (p->wave_inplace) {
            if (per_row / 2 > ctx->smem_optin)
                return sr2_fail(ctx, SR2_ERR_RANGE, "dtl: species tree too large for one CTA's shared memory");
            p->wave_smem = per_row / 2;
            SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_inplace_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)p->wave_smem));
            return SR2_OK;
        }
        p->wave_smem = per_row + topology;
        p->wave_threads = S > 2048 ? 1024 : 512;
        if (const char* t = getenv("SR2_DTL_CTA_THREADS")) p->wave_threads = atoi(t) >= 1024 ? 1024 : 512;
        SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
        SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
        SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_kernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
        SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
    }
    return SR2_OK;
}

static int dtl_configure_simd(sr2_ctx* ctx, DtlProblem* p) {
    const int S = ctx->S;
    p->rank_bits = ctx->dsp.rank_bits;
    p->depth_bits = ctx->dsp.depth_bits;
    int value_bits = 32 - p->rank_bits - p->depth_bits;
    // finite sums stay below inf32 minus the largest correction a candidate subtracts
    int64_t inf32 = value_bits >= 2 ? (int64_t)1 << (value_bits - 1) : 0;
    int64_t slack = 4 * (int64_t)p->costs[3] * (ctx->D + 2);
    const char* off = getenv("SR2_DTL_NO_SIMD");
    p->simd_ok = !(off && off[0] == '1') && S >= 2 && value_bits >= 8 && value_bits <= 27 &&
                 p->max_finite + slack < inf32 - slack;
    p->simd_thresh = (int)(inf32 - slack);
    // waves below this size that are not part of the chained launch keep the row-per-warp kernel with
    // 64-bit keys; the threshold can be overridden for tests (SR2_DTL_SIMD_MIN_ROWS)
    p->simd_min_rows = 2 * (int64_t)ctx->sm_count;  // measured: e2e 11.5 -> 11.0 ms against 16 x SMs, value unchanged
    if (const char* min_rows = getenv("SR2_DTL_SIMD_MIN_ROWS")) p->simd_min_rows = atoll(min_rows);
    // "flat": the first batched form, selected with SR2_DTL_VARIANT=flat (a cross-check of flat2 in the tests)
    p->flat = false;
    if (const char* v = getenv("SR2_DTL_VARIANT")) p->flat = strcmp(v, "flat") == 0;
    p->flat_rs = (S + 2 + 15) / 16 * 16 + 2;   // uint2 elements per array; even (aligned uint4 pairs)
    p->flat_smem = (size_t)8 * 2 * p->flat_rs * sizeof(uint2);
    if (p->flat_smem > ctx->smem_optin || !p->simd_ok) p->flat = false;
    if (p->flat)
        SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->flat_smem));
    // flat2 (default): needs the schedule in FLAT2_NS registers, 16-bit shared-memory addresses and
    // at most 26 value bits (two infinite keys plus corrections, times 8, must fit an int)
    {
        const char* v = getenv("SR2_DTL_VARIANT");
        int vb = std::min(32 - p->rank_bits - p->depth_bits, 26);
        int64_t inf = vb >= 2 ? (int64_t)1 << (vb - 1) : 0;
        p->flat2_sh = 32 - vb;
        p->flat2_thresh = (int)(inf - slack);
        p->flat2_rows = 8;  // rows (warps) per CTA
        if (const char* rpc = getenv("SR2_FLAT2_ROWS")) p->flat2_rows = std::max(1, std::min(8, atoi(rpc)));
        p->flat2_smem = (size_t)p->flat2_rows * 2 * ctx->dsp.flat_rs * sizeof(uint2);
        p->flat2 = (!v || strcmp(v, "flat2") == 0) && !(off && off[0] == '1') && S >= 2 && vb >= 8 &&
                   p->max_finite + slack < inf - slack && ctx->dsp.flat_rs > 0 && ctx->dsp.n_steps <= FLAT2_NS && ctx->pitch <= 32 * FLAT2_NCH &&
                   p->flat2_smem <= std::min<size_t>(ctx->smem_optin, 65536 - 2048);  // 16-bit addresses: system + static words come first
        if (p->flat2) {
            p->flat2_fn = flat2_pick<false>(ctx->dsp.n_steps);
            p->flat2_chain_fn = flat2_pick<true>(ctx->dsp.n_steps);
            // waves below 4 x the resident rows of the kernel are latency-bound as launches of their own
            if (const char* lw = getenv("SR2_FLAT2_LOW_WAVES")) p->flat2_low_waves = atoi(lw);
            p->split_waves = true;  // measured on config #2: 7.33 -> 7.15 ms per step
            if (const char* sw = getenv("SR2_DTL_SPLIT_WAVES")) p->split_waves = sw[0] != '0';
            p->split_tiers = true;  // measured on config #2: 6.34 -> 6.27 ms per step (an almost empty tier's CTAs no longer sit in front of the other)
            if (const char* st = getenv("SR2_DTL_SPLIT_TIERS")) p->split_tiers = st[0] != '0';
            p->chain_rows = 4 * 8 * 4 * (int64_t)ctx->sm_count;
            if (const char* cr = getenv("SR2_DTL_CHAIN_ROWS")) p->chain_rows = atoll(cr);
            SR2_CUDA(ctx, cudaFuncSetAttribute((const void*)p->flat2_chain_fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)p->flat2_smem));
            // ceil(m / 32) sparse passes cost about as much as 2 dense chunks each (locating a lane's species);
            // measured on config #2 (pitch 416): 104 / 144 / 176 / 208 / 256 -> 6.84 / 6.78 / 6.79 / 6.80 / 6.86 ms
            p->flat2_sparse_max = std::min(S, ctx->pitch * 3 / 8);
            if (const char* sm = getenv("SR2_FLAT2_SPARSE_MAX")) p->flat2_sparse_max = atoi(sm);
            p->sparse_maxm = ctx->dsp.mask_words <= 16 ? SPARSE_MAXM : 0;
            SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_sparse_kernel<SPARSE_MAXM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(4 * 8 * SPARSE_ROW_WORDS(SPARSE_MAXM) * sizeof(uint32_t))));
            SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_sparse_kernel<SPARSE_MAXM_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(4 * 8 * SPARSE_ROW_WORDS(SPARSE_MAXM_SMALL) * sizeof(uint32_t))));
            SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_sparse_kernel<SPARSE_MAXM_SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            p->sparse_small = SPARSE_MAXM_SMALL;
            // a thread per sparse row (dtl_wave_thread_kernel) when species and depth fit its element word
            p->sparse_thread = p->rank_bits <= 9 && p->depth_bits <= 5 && ctx->dsp.lca_depth && ctx->dsp.span4;
            if (const char* th = getenv("SR2_SPARSE_THREAD")) p->sparse_thread = p->sparse_thread && th[0] != '0';
            if (const char* sm = getenv("SR2_SPARSE_SMALL")) p->sparse_small = std::min(atoi(sm), SPARSE_MAXM_SMALL);
            if (const char* sm = getenv("SR2_SPARSE_MAXM")) p->sparse_maxm = std::min(atoi(sm), SPARSE_MAXM);
            // Tiers of the thread kernel: path-set limits, one launch (and shared-memory size) each, at most 64
            // species (6-bit positions).  Classes of the partition = bins of path-set sizes (width 2 up to 40
            // species, 4 above) cut at the tier limits; the class after the last tier is the dense one.
            {
                std::vector<uint8_t> table(2048, 0);
                dtl_thread_tiers(ctx, p, &table);
                void* ptr = nullptr;
                int rc = sr2_dev_reserve(ctx, DP_DTL_CLASSES, table.size(), &ptr);
                if (rc) return rc;
                if (p->d_class_of != (uint8_t*)ptr || p->class_table != table) {  // (every solve configures: upload once)
                    p->d_class_of = (uint8_t*)ptr;
                    p->class_table = table;
                    SR2_CUDA(ctx, cudaMemcpyAsync(p->d_class_of, p->class_table.data(), table.size(), cudaMemcpyHostToDevice, ctx->stream));
                }
                const int th_smem_max = (int)THREAD_SMEM_BYTES(p->tier_maxm[p->n_tiers - 1], ctx->dsp.mask_words, ctx->S);
                SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_thread_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, th_smem_max));
                SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_thread_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_thread_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, th_smem_max));
                SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_wave_thread_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                // compact rows: everything the thread kernel writes, and the rows of flat2's sparse cells pass (row_cnt is a byte)
                p->compact_max = std::max(p->tier_maxm[p->n_tiers - 1], std::min(p->flat2_sparse_max, 255));
                if (const char* dd = getenv("SR2_DTL_DECOUPLE_DENSE")) p->decouple_dense = dd[0] != '0';
                if (const char* cm = getenv("SR2_COMPACT_MAX")) p->compact_max = std::max(p->tier_maxm[p->n_tiers - 1], std::min(atoi(cm), 255));
            }
            const int resident = 8 * 4 * ctx->sm_count;  // rows of the CTAs that run at one time
            p->flat2_pf_near = resident;
            p->flat2_pf_far = 3 * resident;
            const int sparse_resident = 32 * 4 * ctx->sm_count;  // rows of the sparse kernel's resident CTAs
            p->sparse_pf_near = sparse_resident / 2;   // measured: 8.07 / 7.80 / 8.09 ms per step for 0 / 0.5 / 1 x resident
            p->sparse_pf_far = 3 * p->sparse_pf_near;
            if (const char* pf = getenv("SR2_SPARSE_PREFETCH")) {
                p->sparse_pf_near = atoi(pf) * sparse_resident / 4;
                p->sparse_pf_far = 3 * p->sparse_pf_near;
            }
            if (const char* pf = getenv("SR2_FLAT2_PREFETCH")) {
                p->flat2_pf_near = atoi(pf) * resident / 4;
                p->flat2_pf_far = 3 * p->flat2_pf_near;
            }
            SR2_CUDA(ctx, cudaFuncSetAttribute((const void*)p->flat2_fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)p->flat2_smem));
        }
    }
    return SR2_OK;
}

static int dtl_launch_wave(sr2_ctx* ctx, DtlProblem* p, const RowChunk& c, int h, int64_t row_begin,
                           int64_t n_rows, const DtlCosts& cs) {
    if (n_rows <= 0) return SR2_OK;
    if (h == 1 && ctx->dsp.anc_tab && !getenv("SR2_DTL_DENSE_CHERRIES")) {
        dtl_cherry_sparse_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, p->cur_stream>>>(
            ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, (int)n_rows, p->cur_T, cs,
            nullptr);
        SR2_CUDA(ctx, cudaGetLastError());
        return SR2_OK;
    }
    if (h == 1) {  // both children of a height-1 node are leaves
        dtl_cherry_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, p->cur_stream>>>(
            ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, (int)n_rows, p->cur_T, p->cur_tag, cs,
            nullptr);
        SR2_CUDA(ctx, cudaGetLastError());
        return SR2_OK;
    }
    if (p->pending_dense && !(p->flat2 && n_rows >= p->simd_min_rows)) {
        SR2_CUDA(ctx, cudaStreamWaitEvent(p->cur_stream, p->pending_dense, 0));
        p->pending_dense = nullptr;
    }
    if (p->flat2 && n_rows >= p->simd_min_rows) {
        Key32 kp;
        kp.db = p->depth_bits;
        kp.sh = p->flat2_sh;
        kp.rm = (1u << p->rank_bits) - 1;
        kp.dm = (1u << p->depth_bits) - 1;
        kp.inf32 = 1u << (32 - kp.sh - 1);
        kp.thresh = p->flat2_thresh;
        // path sets + the wave's sparse / dense row lists, then one launch per list (grids sized for
        // the whole wave: CTAs beyond a list's length return at once)
        int* counts = p->d_counts + ((size_t)p->cur_chunk * 65536 + h) * DTL_MAX_CLASSES;
        int* lists = p->d_lists + row_begin;
        if (p->cur_part_event) SR2_CUDA(ctx, cudaStreamWaitEvent(p->cur_stream, p->cur_part_event, 0));
        // the dense list runs next to the sparse one on a stream of its own (both depend on the
        // previous wave only): the tail of one launch overlaps the other
        cudaStream_t ds = p->cur_stream, ts[8];
        cudaEvent_t join = nullptr, join_tier[8];
        for (int k = 0; k < 8; ++k) ts[k] = p->cur_stream, join_tier[k] = nullptr;
        const int n_sparse = p->sparse_maxm <= 0 ? 0 : (p->cur_compact ? p->n_tiers : (p->sparse_small > 0 ? 2 : 1));
        if (p->split_waves && p->sparse_maxm > 0) {
            const int set = p->cur_set;
            if (!ctx->dense_stream[set]) SR2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->dense_stream[set], cudaStreamNonBlocking));
            auto& pool = ctx->split_events[set];
            while (pool.size() < ctx->split_used[set] + 10) {
                cudaEvent_t e;
                SR2_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                pool.push_back(e);
            }
            cudaEvent_t fork = pool[ctx->split_used[set]++];
            join = pool[ctx->split_used[set]++];
            ds = ctx->dense_stream[set];
            SR2_CUDA(ctx, cudaEventRecord(fork, p->cur_stream));  // the previous wave (and the lists) are done
            SR2_CUDA(ctx, cudaStreamWaitEvent(ds, fork, 0));
            if (p->split_tiers) {  // the sparse tiers after the first on streams of their own as well
                for (int k = 1; k < n_sparse; ++k) {
                    cudaStream_t& tk = ctx->tier_streams[set][k - 1];
                    if (!tk) SR2_CUDA(ctx, cudaStreamCreateWithFlags(&tk, cudaStreamNonBlocking));
                    join_tier[k] = pool[ctx->split_used[set]++];
                    ts[k] = tk;
                    SR2_CUDA(ctx, cudaStreamWaitEvent(tk, fork, 0));
                }
            }
        }
        int dense_cls = 2;
        if (p->sparse_maxm > 0 && p->cur_compact) {  // the sparse rows, a thread per row, compact rows: one launch per tier
            // (grids sized for the whole wave: the warps beyond a tier's lists return at once)
            const unsigned ctas = (unsigned)((n_rows + THREAD_ROWS - 1) / THREAD_ROWS);
            for (int k = 0; k < p->n_tiers; ++k) {
                ++p->n_launch;
                // (wave 2 with virtual cherries: no row has a stored child)
                auto fn = (h == 2 && p->cur_cherry_end > 0) ? dtl_wave_thread_kernel<false> : dtl_wave_thread_kernel<true>;
                fn<<<ctas, THREAD_ROWS, THREAD_SMEM_BYTES(p->tier_maxm[k], ctx->dsp.mask_words, ctx->S), ts[k]>>>(
                    ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, p->cur_T, p->cur_tag, cs, kp,
                    p->cur_mask, lists, (int)p->list_pitch, counts, p->tier_cls0[k], p->tier_cls0[k + 1] - p->tier_cls0[k],
                    p->tier_maxm[k], p->cur_cherry_end, p->cur_rowcnt, p->cur_rowinfo);
            }
            --p->n_launch;  // (the caller counts one launch for this wave)
            dense_cls = p->tier_cls0[p->n_tiers];
        } else if (p->sparse_maxm > 0) {  // the two sparse tiers (the caller counts one launch for this wave)
            if (p->sparse_small > 0) {
                ++p->n_launch;
                dtl_wave_sparse_kernel<SPARSE_MAXM_SMALL><<<(unsigned)((n_rows + 31) / 32), 128,
                                                            4 * 8 * SPARSE_ROW_WORDS(SPARSE_MAXM_SMALL) * sizeof(uint32_t), p->cur_stream>>>(
                    ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, p->cur_T, p->cur_tag, cs, kp,
                    p->cur_mask, lists, counts, p->sparse_pf_near * 7 / 4, p->sparse_pf_far * 7 / 4, p->cur_cherry_end);
            }
            ++p->n_launch;
            dtl_wave_sparse_kernel<SPARSE_MAXM><<<(unsigned)((n_rows + 31) / 32), 128,
                                                  4 * 8 * SPARSE_ROW_WORDS(SPARSE_MAXM) * sizeof(uint32_t), ts[p->sparse_small > 0 ? 1 : 0]>>>(
                ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, p->cur_T, p->cur_tag, cs, kp,
                p->cur_mask, lists + p->list_pitch, counts + 1, p->sparse_pf_near * 5 / 4, p->sparse_pf_far * 5 / 4, p->cur_cherry_end);
        }
        // low waves hold few dense rows: a fraction of the wave's CTAs (the kernel strides if that is too few)
        int64_t dense_ctas = (n_rows + p->flat2_rows - 1) / p->flat2_rows;
        if (h <= p->flat2_low_waves) dense_ctas = std::max<int64_t>((dense_ctas + 7) / 8, std::min<int64_t>(dense_ctas, 4 * ctx->sm_count));
        p->flat2_fn<<<(unsigned)dense_ctas, 32 * p->flat2_rows, p->flat2_smem, ds>>>(
            ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, (int)n_rows, p->cur_T, p->cur_tag, cs, kp,
            p->cur_mask, p->flat2_sparse_max, p->flat2_pf_near, p->flat2_pf_far, lists + (size_t)dense_cls * p->list_pitch,
            counts + dense_cls, nullptr, p->cur_cherry_end, p->cur_compact ? p->cur_rowcnt : nullptr,
            p->cur_compact ? p->compact_max : 0);
        if (join) {
            SR2_CUDA(ctx, cudaEventRecord(join, ds));
            // Compact rows: a row of the thread kernel only ever reads rows of the thread kernel (its children's
            // path sets are subsets of its own), so the sparse rows of the next wave need not wait for the dense
            // rows of this one: the dense stream runs as a pipeline of its own behind the sparse one (it waits for
            // each wave's fork event) and joins before anything else reads the tables.
            if (p->cur_compact && p->decouple_dense) p->pending_dense = join;
            else SR2_CUDA(ctx, cudaStreamWaitEvent(p->cur_stream, join, 0));
        }
        for (int k = 1; k < 8; ++k)
            if (join_tier[k]) {
                SR2_CUDA(ctx, cudaEventRecord(join_tier[k], ts[k]));
                SR2_CUDA(ctx, cudaStreamWaitEvent(p->cur_stream, join_tier[k], 0));
            }
        SR2_CUDA(ctx, cudaGetLastError());
        return SR2_OK;
    }
    if (p->flat && n_rows >= p->simd_min_rows) {
        Key32 kp;
        kp.db = p->depth_bits;
        kp.sh = p->rank_bits + p->depth_bits;
        kp.rm = (1u << p->rank_bits) - 1;
        kp.dm = (1u << p->depth_bits) - 1;
        kp.inf32 = 1u << (32 - kp.sh - 1);
        kp.thresh = p->simd_thresh;
        dtl_wave_flat_kernel<<<(unsigned)((n_rows + 7) / 8), 256, p->flat_smem, p->cur_stream>>>(
            ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, (int)n_rows, p->cur_T, p->cur_tag, cs,
            p->flat_rs, kp);
        SR2_CUDA(ctx, cudaGetLastError());
        return SR2_OK;
    }
    if (!p->cta_per_row) {
        int64_t grid = (n_rows + p->wave_warps - 1) / p->wave_warps;
        dtl_wave_kernel<32><<<(unsigned)grid, p->wave_warps * 32, p->wave_smem, p->cur_stream>>>(
            ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, (int)n_rows, p->cur_T, p->cur_tag, cs,
            p->spad);
    } else {
        if (p->wave_inplace)
            dtl_wave_inplace_kernel<512><<<(unsigned)n_rows, 512, p->wave_smem, p->cur_stream>>>(
            ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, (int)n_rows, p->cur_T, p->cur_tag, cs,
            p->spad);
        else if (p->wave_threads == 1024)
            dtl_wave_kernel<1024><<<(unsigned)n_rows, 1024, p->wave_smem, p->cur_stream>>>(
                ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, (int)n_rows, p->cur_T, p->cur_tag, cs,
                p->spad);
        else
            dtl_wave_kernel<512><<<(unsigned)n_rows, 512, p->wave_smem, p->cur_stream>>>(
                ctx->dsp, p->rows.row_cl, p->rows.row_cr, row_begin, c.row_begin, (int)n_rows, p->cur_T, p->cur_tag, cs,
                p->spad);
    }
    SR2_CUDA(ctx, cudaGetLastError());
    return SR2_OK;
}

static int dtl_configure(sr2_ctx* ctx, DtlProblem* p) {
    int rc = dtl_configure_waves(ctx, p);
    if (rc) return rc;
    return dtl_configure_simd(ctx, p);
}

// All device work of one chunk, enqueued on the context's stream: one wave launch per height,
// (decoded cost rows), selection, one traceback launch per height from the top down.
static int dtl_enqueue_chunk(sr2_ctx* ctx, DtlProblem* p, int ci) {
    // pipelined runs alternate between two streams and two sets of tables, so that the
    // latency-bound top waves and the traceback of one chunk overlap the big waves of the next
    const int set = p->two_streams ? (ci & 1) : 0;
    const size_t table_elems = (size_t)std::max<int64_t>(p->rows.max_chunk_rows, 1) * ctx->pitch;
    p->cur_stream = set ? ctx->stream2 : ctx->stream;
    p->cur_T = p->d_T + set * table_elems;
    p->cur_tag = p->d_tag + set * table_elems;
    p->cur_C = p->d_C ? p->d_C + set * table_elems : nullptr;
    p->cur_mask = p->d_mask + set * (table_elems / ctx->pitch * ctx->dsp.mask_stride);
    p->cur_rowcnt = p->d_rowcnt + set * (table_elems / ctx->pitch);
    p->cur_rowinfo = p->d_rowinfo + set * (table_elems / ctx->pitch) * 2;
    p->cur_desc = p->d_desc + set * (table_elems / ctx->pitch);
    p->cur_chunk = ci;
    p->cur_set = set;
    p->pending_dense = nullptr;
    ctx->split_used[set] = 0;  // (re-recording an event does not disturb waits already enqueued on it)
    cudaStream_t st = p->cur_stream;
    const RowChunk& c = p->rows.chunks[ci];
    if (p->wait_rows) SR2_CUDA(ctx, cudaStreamWaitEvent(st, ctx->chunk_events[ci], 0));  // its rows' H2D copies
    DtlCosts cs{p->costs[0], p->costs[1], p->costs[2], p->costs[3]};
    const int H = (int)c.wave_off.size() - 2;  // highest wave
    int n_inst = c.inst_end - c.inst_begin;
    // one CTA per tree walks it top-down; frontier capacity from the largest tree of the chunk
    int64_t max_nodes = 1;
    for (int b = c.inst_begin; b < c.inst_end; ++b)
        max_nodes = std::max<int64_t>(max_nodes, p->rows.node_off[b + 1] - p->rows.node_off[b]);
    const int64_t cap = (max_nodes + 1) / 4 + 1;  // internal nodes of one depth <= leaves / 2
    const size_t trace_smem = (size_t)2 * cap * sizeof(int2);
    const bool per_tree = trace_smem + 1024 <= ctx->smem_optin && !getenv("SR2_DTL_TRACE_BY_WAVE");
    // Virtual cherries: wave 1 is not launched and its rows are never written when every reader of
    // a cherry row is a kernel that evaluates the closed form itself (sparse / flat2 waves, the
    // per-tree selection + traceback) and nobody asks for the tables.
    bool virt = p->flat2 && per_tree && ctx->dsp.anc_tab && ctx->dsp.lca_depth && !(p->flags & SR2_KEEP_TABLES) &&
                !p->need_decoded_cost && !getenv("SR2_DTL_DENSE_CHERRIES") && !getenv("SR2_DTL_REAL_CHERRIES");
    // Path sets and sparse / dense row lists of every wave do not depend on DP values: they are
    // computed on a side stream, wave after wave, ahead of the wave kernels, which wait for the
    // event of their wave.
    const bool chain = p->flat2;
    cudaEvent_t chain_tail = nullptr;
    // the first wave of the chained launch: waves shrink with height, so every wave from the first
    // one below the threshold on is small (H + 1 = no chained launch)
    int h_chain = H + 1;
    // the generic one-CTA-per-row kernel (large species trees): every wave from the second one on is chained
    const bool gen_chain = !p->flat2 && !p->flat && p->cta_per_row && !p->wave_inplace && H >= 2 &&
                           !getenv("SR2_DTL_NO_GENERIC_CHAIN");
    if (gen_chain) h_chain = 2;
    if (p->flat2 && p->chain_rows > 0)
        for (int h = 2; h <= H; ++h)
            if (c.wave_off[h + 1] - c.wave_off[h] < p->chain_rows) {
                h_chain = h;
                break;
            }
    for (int h = 2; h <= H && h < h_chain; ++h) {
        const int64_t n = c.wave_off[h + 1] - c.wave_off[h];
        if (n > 0 && n < p->simd_min_rows) virt = false;  // a wave of the row-per-warp kernel reads real rows
    }
    if (H < 1) virt = false;  // one-node trees only: no rows at all
    p->cur_cherry_end = virt ? (int)(c.wave_off[2] - c.row_begin) : 0;
    // compact sparse rows need every reader to know the layout: the thread kernel itself (its stored children are
    // then compact rows), flat2 and the per-tree traceback; nobody else may read the tables
    p->cur_compact = virt && p->sparse_thread && p->sparse_maxm > 0 && p->max_finite < (int64_t)TH_VAL_MASK - 16;
    std::vector<cudaEvent_t>& part_events = ctx->part_events[set];
    cudaStream_t ps = nullptr;
    if (chain) {
        if (!ctx->part_stream[set]) SR2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->part_stream[set], cudaStreamNonBlocking));
        ps = ctx->part_stream[set];
        if (p->wait_rows) SR2_CUDA(ctx, cudaStreamWaitEvent(ps, ctx->chunk_events[ci], 0));
        // the table set's path sets may still be read by the chunk that used the set before
        if (ci >= (p->two_streams ? 2 : 1)) SR2_CUDA(ctx, cudaStreamWaitEvent(ps, ctx->dtl_ev_done[ci - (p->two_streams ? 2 : 1)], 0));
        SR2_CUDA(ctx, cudaEventRecord(ctx->dtl_ev_w0[ci], st));   // also orders the side stream after earlier
        SR2_CUDA(ctx, cudaStreamWaitEvent(ps, ctx->dtl_ev_w0[ci], 0));  // work of this stream (uploads, solves)
        SR2_CUDA(ctx, cudaMemsetAsync(p->d_counts + (size_t)ci * DTL_MAX_CLASSES * 65536, 0, (size_t)DTL_MAX_CLASSES * (H + 2) * sizeof(int), ps));
        // rows of waves that are not listed (the chained top waves) are dense
        if (p->cur_compact) SR2_CUDA(ctx, cudaMemsetAsync(p->cur_rowcnt, 0, (size_t)(c.row_end - c.row_begin), ps));
    }
    // (wave 1 has no stored sets: a cherry's set is the two root paths of its leaves, which its parent's
    // kernel reads from the species tables)
    const int cherry_rows = H >= 1 ? (int)(c.wave_off[2] - c.row_begin) : 0;
    size_t events_used = 0;
    int sets_done = 1;  // path sets (and lists) are enqueued for the waves up to this height
    // enqueue the side-stream work of the waves up to `upto`; returns the event of wave h's lists through *ev
    auto enqueue_sets = [&](int upto, int want, cudaEvent_t* ev) -> int {
        for (int h = sets_done + 1; h <= std::min(upto, H); ++h) {
            sets_done = h;
            const int64_t n = c.wave_off[h + 1] - c.wave_off[h];
            if (n <= 0) continue;
            const bool listed = h < h_chain && n >= p->simd_min_rows;
            if (listed) {
                dtl_partition_kernel<<<(unsigned)((n + 63) / 64), 256, 0, ps>>>(
                    ctx->dsp, p->rows.row_cl, p->rows.row_cr, c.wave_off[h], c.row_begin, (int)n, p->cur_mask,
                    p->d_class_of + (p->cur_compact ? 1024 : 0), p->cur_compact ? p->compact_max : 0,
                    p->d_lists + c.wave_off[h], (int)p->list_pitch,
                    p->d_counts + ((size_t)ci * 65536 + h) * DTL_MAX_CLASSES, cherry_rows, p->cur_compact ? p->cur_rowcnt : nullptr,
                    p->cur_compact ? p->cur_rowinfo : nullptr);
                if (part_events.size() <= events_used) {
                    cudaEvent_t e;
                    SR2_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    part_events.push_back(e);
                }
                SR2_CUDA(ctx, cudaEventRecord(part_events[events_used], ps));
                if (h == want && ev) *ev = part_events[events_used];
                ++events_used;
                ++p->n_launch;
            } else {
                const int64_t quads = n * (ctx->dsp.mask_stride / 4);
                dtl_rowmask_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, ps>>>(
                    ctx->dsp, p->rows.row_cl, p->rows.row_cr, c.wave_off[h], c.row_begin, (int)n, p->cur_mask, cherry_rows);
                ++p->n_launch;
            }
        }
        return SR2_OK;
    };
    SR2_CUDA(ctx, cudaEventRecord(ctx->dtl_ev_w0[ci], st));
    // The side stream stays two waves ahead of the wave kernels; enqueueing all of it first would keep
    // the first wave kernel waiting for ~30 launches of host time.
    std::vector<cudaEvent_t> list_event(H + 2, nullptr);
    for (int h = virt ? 2 : 1; h <= H && h < h_chain; ++h) {
        int64_t n = c.wave_off[h + 1] - c.wave_off[h];
        if (chain) {
            for (int a2 = sets_done + 1; a2 <= std::min(h + 2, H); ++a2) {
                cudaEvent_t e = nullptr;
                int rc = enqueue_sets(a2, a2, &e);
                if (rc) return rc;
                list_event[a2] = e;
            }
        }
        if (n <= 0) continue;
        p->cur_part_event = (chain && h > 1 && n >= p->simd_min_rows) ? list_event[h] : nullptr;
        int rc = dtl_launch_wave(ctx, p, c, h, c.wave_off[h], n, cs);
        if (rc) return rc;
        ++p->n_waves;
        ++p->n_launch;
    }
    if (p->pending_dense) {  // the dense pipeline joins the chunk's stream
        SR2_CUDA(ctx, cudaStreamWaitEvent(st, p->pending_dense, 0));
        p->pending_dense = nullptr;
    }
    if (chain) {
        // the rest of the side stream (path sets of the chained top waves) joins the chunk's stream
        int rc = enqueue_sets(H, -1, nullptr);
        if (rc) return rc;
        if (part_events.size() <= events_used) {
            cudaEvent_t e;
            SR2_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            part_events.push_back(e);
        }
        SR2_CUDA(ctx, cudaEventRecord(part_events[events_used], ps));
        chain_tail = part_events[events_used];
        SR2_CUDA(ctx, cudaGetLastError());
    }
    if (h_chain <= H && gen_chain) {
        const int64_t first = c.wave_off[h_chain], n = c.wave_off[H + 1] - first;
        int* sync = p->d_chain_sync + (size_t)set * (std::max<int64_t>(p->rows.max_chunk_rows, 1) + 1);
        SR2_CUDA(ctx, cudaMemsetAsync(sync, 0, (size_t)(n + 1) * sizeof(int), st));
        if (p->wave_threads == 1024)
            dtl_wave_kernel<1024, true><<<(unsigned)n, 1024, p->wave_smem, st>>>(
                ctx->dsp, p->rows.row_cl, p->rows.row_cr, first, c.row_begin, (int)n, p->cur_T, p->cur_tag, cs, p->spad, sync);
        else
            dtl_wave_kernel<512, true><<<(unsigned)n, 512, p->wave_smem, st>>>(
                ctx->dsp, p->rows.row_cl, p->rows.row_cr, first, c.row_begin, (int)n, p->cur_T, p->cur_tag, cs, p->spad, sync);
        SR2_CUDA(ctx, cudaGetLastError());
        ++p->n_waves;
        ++p->n_launch;
    } else if (h_chain <= H) {
        const int64_t first = c.wave_off[h_chain], n = c.wave_off[H + 1] - first;
        int* sync = p->d_chain_sync + (size_t)set * (std::max<int64_t>(p->rows.max_chunk_rows, 1) + 1);
        SR2_CUDA(ctx, cudaMemsetAsync(sync, 0, (size_t)(n + 1) * sizeof(int), st));
        if (chain_tail) {  // the path sets of all chained rows
            SR2_CUDA(ctx, cudaStreamWaitEvent(st, chain_tail, 0));
            chain_tail = nullptr;
        }
        Key32 kp;
        kp.db = p->depth_bits;
        kp.sh = p->flat2_sh;
        kp.rm = (1u << p->rank_bits) - 1;
        kp.dm = (1u << p->depth_bits) - 1;
        kp.inf32 = 1u << (32 - kp.sh - 1);
        kp.thresh = p->flat2_thresh;
        p->flat2_chain_fn<<<(unsigned)((n + p->flat2_rows - 1) / p->flat2_rows), 32 * p->flat2_rows, p->flat2_smem, st>>>(
            ctx->dsp, p->rows.row_cl, p->rows.row_cr, first, c.row_begin, (int)n, p->cur_T, p->cur_tag, cs, kp,
            p->cur_mask, p->flat2_sparse_max, 0, 0, nullptr, nullptr, sync, p->cur_cherry_end,
            p->cur_compact ? p->cur_rowcnt : nullptr, 0);
        SR2_CUDA(ctx, cudaGetLastError());
        ++p->n_waves;
        ++p->n_launch;
    }
    SR2_CUDA(ctx, cudaEventRecord(ctx->dtl_ev_w1[ci], st));
    if (p->need_decoded_cost) {
        for (int h = 1; h <= H; ++h) {
            int64_t n = c.wave_off[h + 1] - c.wave_off[h];
            if (n <= 0) continue;
            dtl_decoded_cost_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(
                ctx->dsp, p->rows.row_cl, p->rows.row_cr, c.wave_off[h], c.row_begin, (int)n, p->cur_tag, p->cur_C, cs, p->cur_T);
            ++p->n_launch;
        }
    }
    // the packed row records of the traceback: on the side stream, behind the path sets (it has run dry by now)
    const bool use_desc = ps && per_tree && n_inst > 0 && c.row_end > c.row_begin && !getenv("SR2_DTL_NO_DESC");
    if (use_desc) {
        const int64_t n = c.row_end - c.row_begin;
        dtl_desc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ps>>>(p->rows.row_cl, p->rows.row_cr, p->rows.row_left,
                                                                    p->rows.row_right, c.row_begin, n, p->cur_desc);
        if (part_events.size() <= events_used) {
            cudaEvent_t e;
            SR2_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            part_events.push_back(e);
        }
        SR2_CUDA(ctx, cudaEventRecord(part_events[events_used], ps));
        SR2_CUDA(ctx, cudaStreamWaitEvent(st, part_events[events_used], 0));
        ++events_used;
        ++p->n_launch;
    }
    if (n_inst > 0 && per_tree) {
        int threads = 32;
        // measured on config #2 (cap 251; step time with compact rows): 32 / 64 / 128 / 256 threads -> 5.13 / 5.08 / 5.20 / 5.41 ms
        while (threads < 1024 && threads * 4 < cap) threads *= 2;
        if (const char* t = getenv("SR2_DTL_TRACE_THREADS")) threads = std::max(32, std::min(1024, atoi(t) / 32 * 32));
        if (trace_smem > 48 * 1024 && trace_smem > p->trace_smem_set) {
            SR2_CUDA(ctx, cudaFuncSetAttribute(dtl_select_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)trace_smem));
            p->trace_smem_set = trace_smem;
        }
        dtl_select_trace_kernel<<<n_inst, threads, trace_smem, st>>>(
            ctx->dsp, p->rows.d_root_row, p->rows.d_root_node, c.inst_begin, c.row_begin, p->cur_T,
            p->need_decoded_cost ? p->cur_C : nullptr, p->cur_tag, p->rows.row_left, p->rows.row_right,
            p->rows.row_cl, p->rows.row_cr, (int)cap, getenv("SR2_DTL_TRACE_CHECK") ? 1 : 0, p->cur_cherry_end, cs,
            p->d_min_cost, p->d_root_species, p->d_map, p->cur_compact ? p->cur_rowcnt : nullptr, p->cur_mask,
            use_desc ? p->cur_desc : nullptr);
        ++p->n_launch;
    } else if (n_inst > 0) {
        dtl_select_kernel<<<(n_inst + 7) / 8, 256, 0, st>>>(
            ctx->dsp, p->rows.d_root_row, p->rows.d_root_node, c.inst_begin, n_inst, c.row_begin, p->cur_T,
            p->need_decoded_cost ? p->cur_C : nullptr, p->d_min_cost, p->d_root_species, p->d_map);
        ++p->n_launch;
    }
    for (int h = H; h >= 1 && !per_tree; --h) {
        int64_t n = c.wave_off[h + 1] - c.wave_off[h];
        if (n <= 0) continue;
        dtl_traceback_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
            ctx->dsp, p->rows.row_node, p->rows.row_left, p->rows.row_right, c.wave_off[h], c.row_begin, (int)n,
            p->cur_tag, p->d_map, p->rows.row_cl, p->rows.row_cr, p->cur_T);
        ++p->n_launch;
    }
    SR2_CUDA(ctx, cudaGetLastError());
    if (chain_tail) SR2_CUDA(ctx, cudaStreamWaitEvent(st, chain_tail, 0));
    SR2_CUDA(ctx, cudaEventRecord(ctx->dtl_ev_done[ci], st));
    return SR2_OK;
}

// after a stream synchronize: timings of the chunks that ran
static void dtl_collect_timing(sr2_ctx* ctx, DtlProblem* p) {
    float waves_ms = 0.f;
    for (size_t ci = 0; ci < p->rows.chunks.size(); ++ci) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->dtl_ev_w0[ci], ctx->dtl_ev_w1[ci]);
        waves_ms += ms;
    }
    cudaEventElapsedTime(&ctx->timing.solve_ms, ctx->ev[0], ctx->ev[1]);
    if (getenv("SR2_TRACE"))
        for (size_t ci = 0; ci < p->rows.chunks.size(); ++ci) {
            float a = 0.f, b = 0.f, c = 0.f;
            cudaEventElapsedTime(&a, ctx->ev[0], ctx->dtl_ev_w0[ci]);
            cudaEventElapsedTime(&b, ctx->ev[0], ctx->dtl_ev_w1[ci]);
            cudaEventElapsedTime(&c, ctx->ev[0], ctx->dtl_ev_done[ci]);
            fprintf(stderr, "[sr2] chunk %zu on the device: waves %.2f .. %.2f ms, done %.2f ms\n", ci, a, b, c);
        }
    ctx->timing.waves_ms = waves_ms;
    ctx->timing.n_waves = p->n_waves;
    ctx->timing.n_launches = p->n_launch;
    ctx->timing.n_cells = p->rows.R * (int64_t)ctx->S;
}

extern "C" int sr2_dtl_solve(sr2_ctx* ctx) {
    if (!ctx) return SR2_ERR_ARG;
    DtlProblem* p = ctx->dtl;
    if (!p || !p->prepared) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_dtl_solve: nothing uploaded");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = dtl_configure(ctx, p);
    if (rc) return rc;
    p->n_waves = p->n_launch = 0;
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    for (size_t ci = 0; ci < p->rows.chunks.size(); ++ci)
        if ((rc = dtl_enqueue_chunk(ctx, p, (int)ci))) return rc;
    if (p->two_streams)  // chunks on the second stream
        for (size_t ci = 1; ci < p->rows.chunks.size(); ci += 2)
            SR2_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->dtl_ev_done[ci], 0));
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    SR2_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    dtl_collect_timing(ctx, p);
    p->solved = true;
    return SR2_OK;
}

extern "C" int sr2_dtl_results(sr2_ctx* ctx, int32_t* out_min_cost, int32_t* out_root_species,
                               int32_t* out_map_species) {
    if (!ctx) return SR2_ERR_ARG;
    DtlProblem* p = ctx->dtl;
    if (!p || !p->solved) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_dtl_results: solve first");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    double t0 = now_ms();
    const size_t B = p->rows.B, N = p->rows.N;
    // D2H into pinned staging (full PCIe speed), then a parallel copy into the caller's buffers
    void* ptr = nullptr;
    int rc = sr2_pin_reserve(ctx, HP_RESULTS, (2 * B + N + 16) * sizeof(int32_t), &ptr);
    if (rc) return rc;
    int32_t* stage = (int32_t*)ptr;
    if (out_min_cost)
        SR2_CUDA(ctx, cudaMemcpyAsync(stage, p->d_min_cost, B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_root_species)
        SR2_CUDA(ctx, cudaMemcpyAsync(stage + B, p->d_root_species, B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_map_species)
        SR2_CUDA(ctx, cudaMemcpyAsync(stage + 2 * B, p->d_map, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_min_cost) memcpy(out_min_cost, stage, B * 4);
    if (out_root_species) memcpy(out_root_species, stage + B, B * 4);
    if (out_map_species) {
        const int64_t block = 1 << 18;
        const int64_t n_blocks = ((int64_t)N + block - 1) / block;
#pragma omp parallel for num_threads(ctx->host_threads) schedule(static) if (n_blocks > 1)
        for (int64_t k = 0; k < n_blocks; ++k) {
            int64_t lo = k * block, hi = std::min<int64_t>((int64_t)N, lo + block);
            memcpy(out_map_species + lo, stage + 2 * B + lo, (size_t)(hi - lo) * 4);
        }
    }
    ctx->timing.results_ms = (float)(now_ms() - t0);
    return SR2_OK;
}

// sr2_dtl_run = upload + solve + results, pipelined over instance chunks: while the device solves
// chunk k (stream-ordered behind its own H2D copies) the host validates and buckets chunk k+1,
// and a second stream copies chunk k's results into pinned staging.
struct DtlRunState {
    sr2_ctx* ctx;
    int32_t* stage;  // pinned: [B] min cost, [B] root species, [N] mapping
    int32_t *out_cost, *out_root, *out_map;
    bool started;
    bool direct;     // the caller's result arrays are pinned: copy straight into them
    int delivered;   // chunks whose results already sit in the caller's buffers
    // sr2_dtl_run16: root species and mapping leave the device as 16-bit integers
    int16_t *out_root16 = nullptr, *out_map16 = nullptr;
    int16_t *stage16 = nullptr, *d_narrow = nullptr;   // pinned [B + N]; device [B + N]
};

// results as 16-bit integers before they cross PCIe (species ranks fit: S <= 32767 is checked)
__global__ void dtl_narrow_kernel(const int32_t* __restrict__ in, int16_t* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int16_t)in[i];
}

// staging -> caller's buffers for chunks [rs->delivered, upto) whose copies have finished
// (wait == false: stop at the first chunk that is still in flight)
static int dtl_deliver(DtlRunState* rs, int upto, bool wait) {
    sr2_ctx* ctx = rs->ctx;
    DtlProblem* p = ctx->dtl;
    const size_t B = p->rows.B;
    while (rs->delivered < upto) {
        cudaEvent_t e = ctx->dtl_ev_copied[rs->delivered];
        if (wait) SR2_CUDA(ctx, cudaEventSynchronize(e));
        else if (cudaEventQuery(e) != cudaSuccess) break;
        const RowChunk& c = p->rows.chunks[rs->delivered];
        const size_t b0 = c.inst_begin, nb = c.inst_end - c.inst_begin;
        const int64_t n0 = p->rows.node_off[c.inst_begin], nn = p->rows.node_off[c.inst_end] - n0;
        if (rs->direct) {
            ++rs->delivered;
            continue;
        }
        if (rs->out_cost) memcpy(rs->out_cost + b0, rs->stage + b0, nb * 4);
        if (rs->out_root16) memcpy(rs->out_root16 + b0, rs->stage16 + b0, nb * 2);
        if (rs->out_map16) {
            const int64_t block = 1 << 18;
            const int64_t n_blocks = (nn + block - 1) / block;
#pragma omp parallel for num_threads(ctx->host_threads) schedule(static) if (n_blocks > 1)
            for (int64_t k = 0; k < n_blocks; ++k) {
                int64_t lo = n0 + k * block, hi = std::min<int64_t>(n0 + nn, lo + block);
                memcpy(rs->out_map16 + lo, rs->stage16 + B + lo, (size_t)(hi - lo) * 2);
            }
        }
        if (rs->out_root) memcpy(rs->out_root + b0, rs->stage + B + b0, nb * 4);
        if (rs->out_map) {
            const int64_t block = 1 << 17;
            const int64_t n_blocks = (nn + block - 1) / block;
#pragma omp parallel for num_threads(ctx->host_threads) schedule(static) if (n_blocks > 1)
            for (int64_t k = 0; k < n_blocks; ++k) {
                int64_t lo = n0 + k * block, hi = std::min<int64_t>(n0 + nn, lo + block);
                memcpy(rs->out_map + lo, rs->stage + 2 * B + lo, (size_t)(hi - lo) * 4);
            }
        }
        ++rs->delivered;
    }
    return SR2_OK;
}

static int dtl_run_on_plan(void* user) {
    DtlRunState* rs = (DtlRunState*)user;
    sr2_ctx* ctx = rs->ctx;
    if (rs->started) {  // a second plan: the device-side bucketing handed over to the host-side one
        cudaStreamSynchronize(ctx->copy_stream);
        rs->delivered = 0;
    }
    int rc = dtl_prepare(ctx);
    if (rc) return rc;
    DtlProblem* p = ctx->dtl;
    if ((rc = dtl_configure(ctx, p))) return rc;
    void* ptr = nullptr;
    if ((rc = sr2_pin_reserve(ctx, HP_RESULTS, (2 * (size_t)p->rows.B + (size_t)p->rows.N + 16) * sizeof(int32_t), &ptr)))
        return rc;
    rs->stage = (int32_t*)ptr;
    {
        auto pinned = [](const void* q) {
            if (!q) return true;
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, q) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            return attr.type == cudaMemoryTypeHost;
        };
        rs->direct = pinned(rs->out_cost) && pinned(rs->out_root) && pinned(rs->out_map) &&
                     pinned(rs->out_root16) && pinned(rs->out_map16);
    }
    if (rs->out_root16 || rs->out_map16) {
        const size_t n16 = (size_t)p->rows.B + (size_t)p->rows.N + 16;
        if ((rc = sr2_pin_reserve(ctx, HP_RESULTS16, n16 * sizeof(int16_t), &ptr))) return rc;
        rs->stage16 = (int16_t*)ptr;
        if ((rc = sr2_dev_reserve(ctx, DP_DTL_MAP16, n16 * sizeof(int16_t), &ptr))) return rc;
        rs->d_narrow = (int16_t*)ptr;
    }
    if (!ctx->copy_stream) SR2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    p->n_waves = p->n_launch = 0;
    p->wait_rows = true;
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    rs->started = true;
    return SR2_OK;
}

static int dtl_run_on_chunk(void* user, int ci) {
    DtlRunState* rs = (DtlRunState*)user;
    sr2_ctx* ctx = rs->ctx;
    DtlProblem* p = ctx->dtl;
    int rc = dtl_enqueue_chunk(ctx, p, ci);
    if (rc) return rc;
    const RowChunk& c = p->rows.chunks[ci];
    const size_t B = p->rows.B;
    const size_t b0 = c.inst_begin, nb = c.inst_end - c.inst_begin;
    const size_t n0 = p->rows.node_off[c.inst_begin], nn = p->rows.node_off[c.inst_end] - n0;
    cudaStream_t cs = ctx->copy_stream;
    SR2_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->dtl_ev_done[ci], 0));
    int32_t* to_cost = rs->direct ? rs->out_cost : rs->stage;
    int32_t* to_root = rs->direct ? rs->out_root : rs->stage + B;
    int32_t* to_map = rs->direct ? rs->out_map : rs->stage + 2 * B;
    if (rs->out_cost && nb)
        SR2_CUDA(ctx, cudaMemcpyAsync(to_cost + b0, p->d_min_cost + b0, nb * 4, cudaMemcpyDeviceToHost, cs));
    if (rs->out_root && nb)
        SR2_CUDA(ctx, cudaMemcpyAsync(to_root + b0, p->d_root_species + b0, nb * 4, cudaMemcpyDeviceToHost, cs));
    if (rs->out_map && nn)
        SR2_CUDA(ctx, cudaMemcpyAsync(to_map + n0, p->d_map + n0, nn * 4, cudaMemcpyDeviceToHost, cs));
    if (rs->out_root16 && nb) {
        int16_t* to = rs->direct ? rs->out_root16 : rs->stage16;
        dtl_narrow_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, cs>>>(p->d_root_species + b0, rs->d_narrow + b0, (int64_t)nb);
        SR2_CUDA(ctx, cudaMemcpyAsync(to + b0, rs->d_narrow + b0, nb * 2, cudaMemcpyDeviceToHost, cs));
    }
    if (rs->out_map16 && nn) {
        int16_t* to = rs->direct ? rs->out_map16 : rs->stage16 + B;
        dtl_narrow_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, cs>>>(p->d_map + n0, rs->d_narrow + B + n0, (int64_t)nn);
        SR2_CUDA(ctx, cudaMemcpyAsync(to + n0, rs->d_narrow + B + n0, nn * 2, cudaMemcpyDeviceToHost, cs));
    }
    SR2_CUDA(ctx, cudaGetLastError());
    SR2_CUDA(ctx, cudaEventRecord(ctx->dtl_ev_copied[ci], cs));
    return dtl_deliver(rs, ci, false);  // earlier chunks that are back already
}

static int dtl_run_impl(sr2_ctx* ctx, int32_t B, const int64_t* node_off, const int32_t* left, const int32_t* right,
                        const int32_t* leaf_species, const int16_t* const raw16[3], const int32_t costs[5],
                        uint32_t flags, int32_t* out_min_cost, int32_t* out_root_species, int32_t* out_map_species,
                        int16_t* out_root16, int16_t* out_map16) {
    if (!ctx) return SR2_ERR_ARG;
    int pipeline_chunks = 2;  // + a half-sized first chunk: quarter, half, quarter of the batch (measured: 8.74 ms against 8.84 for 3)
    if (const char* v = getenv("SR2_PIPELINE_CHUNKS")) pipeline_chunks = std::max(1, atoi(v));
    DtlRunState rs{ctx, nullptr, out_min_cost, out_root_species, out_map_species, false, false, 0};
    rs.out_root16 = out_root16;
    rs.out_map16 = out_map16;
    RowHooks hooks;
    hooks.user = &rs;
    hooks.on_plan = dtl_run_on_plan;
    hooks.on_chunk = dtl_run_on_chunk;
    if (raw16) {
        hooks.left16 = raw16[0];
        hooks.right16 = raw16[1];
        hooks.leaf16 = raw16[2];
    }
    double t0 = now_ms();
    int rc = dtl_upload_impl(ctx, B, node_off, left, right, leaf_species, costs, flags, pipeline_chunks, &hooks);
    double t1 = now_ms();  // the host side is done; sr2_build_rows waited for the compute stream
    if (getenv("SR2_TRACE")) fprintf(stderr, "[sr2] dtl run: upload_impl returned after %.2f ms\n", t1 - t0);
    DtlProblem* p = ctx->dtl;
    if (!rc) rc = dtl_deliver(&rs, (int)p->rows.chunks.size(), true);
    if (rs.started) {
        cudaStreamSynchronize(ctx->stream);
        if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
        cudaStreamSynchronize(ctx->copy_stream);
    }
    if (rc) return rc;
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    SR2_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    dtl_collect_timing(ctx, p);
    p->solved = true;
    p->wait_rows = false;
    if (getenv("SR2_TRACE")) fprintf(stderr, "[sr2] dtl run: done after %.2f ms\n", now_ms() - t0);
    ctx->timing.upload_ms = (float)(t1 - t0);         // bucketing with the device work overlapped
    ctx->timing.results_ms = (float)(now_ms() - t1);  // what was left of staging -> caller's buffers
    return SR2_OK;
}

extern "C" int sr2_dtl_run(sr2_ctx* ctx, int32_t B, const int64_t* node_off, const int32_t* left,
                           const int32_t* right, const int32_t* leaf_species, const int32_t costs[5],
                           uint32_t flags, int32_t* out_min_cost, int32_t* out_root_species,
                           int32_t* out_map_species) {
    return dtl_run_impl(ctx, B, node_off, left, right, leaf_species, nullptr, costs, flags, out_min_cost,
                        out_root_species, out_map_species, nullptr, nullptr);
}

extern "C" int sr2_dtl_run16(sr2_ctx* ctx, int32_t B, const int64_t* node_off, const int16_t* left,
                             const int16_t* right, const int16_t* leaf_species, const int32_t costs[5],
                             uint32_t flags, int32_t* out_min_cost, int16_t* out_root_species,
                             int16_t* out_map_species) {
    if (!ctx) return SR2_ERR_ARG;
    if (ctx->S > 32767)
        return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_dtl_run16: more than 32767 species nodes (use sr2_dtl_run)");
    if (B < 0 || !node_off) return sr2_fail(ctx, SR2_ERR_ARG, "sr2_dtl_run16: null array");
    for (int b = 0; b < B; ++b)
        if (node_off[b + 1] - node_off[b] > 32767)
            return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_dtl_run16: instance " + std::to_string(b) +
                                                    " has more than 32767 nodes (use sr2_dtl_run)");
    const int16_t* raw16[3] = {left, right, leaf_species};
    return dtl_run_impl(ctx, B, node_off, nullptr, nullptr, nullptr, raw16, costs, flags, out_min_cost, nullptr, nullptr,
                        out_root_species, out_map_species);
}

extern "C" int sr2_dtl_tables(sr2_ctx* ctx, int32_t instance, int32_t* out_value, int32_t* out_tag) {
    if (!ctx) return SR2_ERR_ARG;
    DtlProblem* p = ctx->dtl;
    if (!p || !p->solved) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_dtl_tables: solve first");
    if (!(p->flags & SR2_KEEP_TABLES))
        return sr2_fail(ctx, SR2_ERR_STATE, "sr2_dtl_tables: upload with SR2_KEEP_TABLES");
    if (instance < 0 || instance >= p->rows.B) return sr2_fail(ctx, SR2_ERR_ARG, "sr2_dtl_tables: bad instance");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    const int S = ctx->S, pitch = ctx->pitch;
    int64_t off = p->rows.node_off[instance], G = p->rows.node_off[instance + 1] - off;
    std::vector<int32_t> rowT(pitch);
    std::vector<uint32_t> rowG(pitch);
    for (int64_t u = 0; u < G; ++u) {
        int64_t gr = p->rows.h_node_row[off + u];
        int32_t* ov = out_value ? out_value + (size_t)u * S : nullptr;
        int32_t* ot = out_tag ? out_tag + (size_t)u * S * 2 : nullptr;
        if (gr < 0) {
            for (int x = 0; x < S; ++x) {
                if (ov) ov[x] = x == p->rows.h_leaf[off + u] ? 0 : SR2_INF;
                if (ot) ot[2 * x] = ot[2 * x + 1] = -1;
            }
            continue;
        }
        SR2_CUDA(ctx, cudaMemcpy(rowT.data(), p->d_T + (size_t)gr * pitch, (size_t)pitch * 4, cudaMemcpyDeviceToHost));
        SR2_CUDA(ctx, cudaMemcpy(rowG.data(), p->d_tag + (size_t)gr * pitch, (size_t)pitch * 4, cudaMemcpyDeviceToHost));
        int32_t desc[2];
        SR2_CUDA(ctx, cudaMemcpy(&desc[0], p->rows.row_cl + gr, 4, cudaMemcpyDeviceToHost));
        SR2_CUDA(ctx, cudaMemcpy(&desc[1], p->rows.row_cr + gr, 4, cudaMemcpyDeviceToHost));
        if (desc[0] < 0 && desc[1] < 0)  // a cherry stores no tags (dtl_cherry_kernel)
            for (int x = 0; x < S; ++x)
                rowG[x] = rowT[x] < SR2_INF ? (uint32_t)(-1 - desc[0]) | ((uint32_t)(-1 - desc[1]) << 16) : SR2_NO_TAG;
        for (int x = 0; x < S; ++x) {
            if (ov) ov[x] = rowT[x];
            if (ot) {
                bool none = rowG[x] == SR2_NO_TAG || rowT[x] >= SR2_INF;  // infinite cells carry no tag
                ot[2 * x] = none ? -1 : (int32_t)(rowG[x] & 0xffffu);
                ot[2 * x + 1] = none ? -1 : (int32_t)(rowG[x] >> 16);
            }
        }
    }
    return SR2_OK;
}
