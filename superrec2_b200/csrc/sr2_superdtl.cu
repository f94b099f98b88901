// sr2_superdtl.cu -- the `superdtl` (unordered super-reconciliation) DP on sm_100a.
//
// Replaces /root/reference/src/superrec2/compute/unordered_super_reconciliation.py:
//   _compute_gain_sets / _compute_lca_sets   :85-136   (bitset kernels syn_*)
//   _compute_uspfs_entry                     :147-297  (superdtl_wave_kernel)
//   _compute_uspfs_table                     :300-358  (the wavefront driver)
//   _decode_uspfs_table                      :361-446  (sdecode_*, incl. the in-place
//                                                        set union of :386-391)
//   _uspfs selection                         :449-497  (scost_kernel, sselect_kernel,
//                                                        sbest_kernel)
// and SuperReconciliationOutput.cost() (model/reconciliation.py:511-554), the
// objective the reference actually minimises when it picks the returned solution.
//
// Table state per (object node u, species x): two values, kind LCA (0) and kind
// INHERIT (1).  For each child i and each parent kind K three rows of keys are
// built from the child's two value rows:
//   cons = "conserved" (subtree of x, synteny kept)     -> SUBMIN
//   seg  = "segment"   (subtree of x, segment lost)      -> SUBMIN
//   sep  = "separate"  (species incomparable to x)       -> SEPMIN
// cons/seg carry floss * depth so that a subtree minimum is valid for every
// ancestor x; the cell subtracts floss * depth[x] again ("rel").  The `left` /
// `right` choices of the reference are cons at the two children of x.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>

#include "sr2_rows.cuh"

int sr2_check_costs(sr2_ctx* ctx, const char* who, const int32_t costs[5], int64_t max_nodes,
                    int64_t* out_bound);

struct SuperProblem {
    RowLayout rows;
    uint32_t flags = 0;
    int32_t costs[5] = {0, 0, 0, 0, 0};
    int W = 1;                 // 32-bit words per synteny bitset
    int64_t index_base = 0;    // global number of instance 0 (resolution sharding)
    // device
    int32_t *d_left = nullptr, *d_right = nullptr;       // [N] global node ids, -1 for leaves
    int64_t* d_node_off = nullptr;                       // [B+1]
    uint32_t *d_bits = nullptr, *d_present = nullptr, *d_outside = nullptr, *d_gain = nullptr,
             *d_L = nullptr;                             // [N*W]
    uint8_t* d_row_flags = nullptr;                      // [R]
    int32_t* d_T = nullptr;                              // [rows*2*pitch]
    uint32_t* d_tag = nullptr;                           // [rows*2*pitch]
    uint32_t* d_ev = nullptr;                            // [rows*2*pitch] shared rows: super_event_word of every cell
    uint16_t *d_dec = nullptr, *d_anc = nullptr;         // [N*S] (species<<1|kind), local index of the set owner
    uint32_t *d_tau = nullptr, *d_taumin = nullptr;      // [N*W*32], [N]
    int32_t* d_cost = nullptr;                           // [B*S] cost() of the decode from each root
    int32_t *d_min_cost = nullptr, *d_root_species = nullptr;  // [B]
    int32_t *d_map = nullptr, *d_kind = nullptr;         // [N]
    uint32_t* d_syn = nullptr;                           // [N*W]
    unsigned long long* d_best = nullptr;                // packed (cost, instance, root)
    int32_t* d_allowed = nullptr;                        // [N] the only species a node may map to, -1 = any
    bool solved = false;
    int spad = 0, wave_warps = 0, cta_threads = 256;
    size_t wave_smem = 0;
    bool cta_per_row = false;
    // shared-row mode (resolutions of one multifurcating tree, SharedPlan): the table has one row per
    // distinct subtree ("slot"); per-node arrays still exist for every resolved tree
    bool shared = false;
    bool fused = false;             // the last solve used sfused_kernel
    bool materialised = true;       // d_map / d_kind / d_syn hold the last solve's per-node results
    int Gb = 0;                     // nodes of one resolved tree
    int64_t n_slots = 0;
    std::vector<int64_t> slot_wave_off;
    int32_t *d_node_slot = nullptr, *d_node_dag = nullptr;   // [N] table row / dag id of a node
    int32_t *d_node_pre = nullptr, *d_node_inst = nullptr;   // [N]
    int32_t *d_slot_cl = nullptr, *d_slot_cr = nullptr;      // [n_slots] table descriptors
    int32_t *d_dag_left = nullptr, *d_dag_right = nullptr;   // [Gb + n_slots]
    int32_t* d_slot_node = nullptr;                          // [n_slots] Gb + slot
    // The resolved trees themselves (d_left ... d_node_inst and the per-(node, root) arrays) are only
    // materialised when something needs them: the compact fused kernel unranks its tree itself.
    bool trees_ready = true;
    struct DevPlanStore* plan_store = nullptr;               // device plan of the upload (owned)
    ~SuperProblem();
};

void sr2_free_super(sr2_ctx* ctx) {
    SuperProblem* p = ctx->super;
    if (!p) return;
    delete p;  // device buffers live in the context's pools
    ctx->super = nullptr;
}

struct SuperCosts {
    int spe, dup, hgt, floss, sloss;
};

#define DEC_NONE 0xffffu
// The packed best key is cost << 40 | instance << 16 | root; costs stay below 2^23 so that the key
// also stays below 2^63 (it crosses GPUs as a signed 64-bit all-reduce(MIN)).  Uploads whose cost()
// bound (nodes x largest event term) reaches the limit are refused instead of truncated.
#define SUPER_KEY_COST_LIMIT ((int64_t)1 << 23)

// ---------------------------------------------------------------------------
// synteny bitsets (gain sets and LCA sets)
// ---------------------------------------------------------------------------
// present[u] = families carried by some leaf below u (starts as a copy of the leaf bits)
__global__ void syn_present_wave(const int32_t* __restrict__ row_node, const int32_t* __restrict__ row_left,
                                 const int32_t* __restrict__ row_right, int64_t row_begin, int n_rows, int W,
                                 uint32_t* __restrict__ present) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * W) return;
    int64_t gr = row_begin + i / W;
    int w = i % W;
    present[(size_t)row_node[gr] * W + w] =
        present[(size_t)row_left[gr] * W + w] | present[(size_t)row_right[gr] * W + w];
}

// outside[u] = families carried by some leaf NOT below u (top-down; zero at the roots)
__global__ void syn_outside_wave(const int32_t* __restrict__ row_node, const int32_t* __restrict__ row_left,
                                 const int32_t* __restrict__ row_right, int64_t row_begin, int n_rows, int W,
                                 const uint32_t* __restrict__ present, uint32_t* __restrict__ outside) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * W) return;
    int64_t gr = row_begin + i / W;
    int w = i % W;
    uint32_t mine = outside[(size_t)row_node[gr] * W + w];
    size_t l = (size_t)row_left[gr] * W + w, r = (size_t)row_right[gr] * W + w;
    outside[l] = mine | present[r];
    outside[r] = mine | present[l];
}

// A family is gained at the LCA of the leaves that carry it (:98-107): the node where it
// is "complete" (present and not outside) while complete in neither child.
__global__ void syn_gain_kernel(const int32_t* __restrict__ left, const int32_t* __restrict__ right, int64_t N,
                                int W, const uint32_t* __restrict__ present,
                                const uint32_t* __restrict__ outside, uint32_t* __restrict__ gain) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * W) return;
    int64_t u = i / W;
    int w = (int)(i % W);
    uint32_t full = present[i] & ~outside[i];
    int l = left[u];
    if (l >= 0) {
        size_t li = (size_t)l * W + w, ri = (size_t)right[u] * W + w;
        full &= ~((present[li] & ~outside[li]) | (present[ri] & ~outside[ri]));
    }
    gain[i] = full;
}

// L[u] = (L[l] | L[r]) - (gain[l] | gain[r]) (:126-134); also the two subset flags that
// _compute_uspfs_entry tests per child (:182).  L starts as a copy of the leaf bits.
__global__ void syn_lca_wave(const int32_t* __restrict__ row_node, const int32_t* __restrict__ row_left,
                             const int32_t* __restrict__ row_right, int64_t row_begin, int n_rows, int W,
                             const uint32_t* __restrict__ gain, uint32_t* __restrict__ L,
                             uint8_t* __restrict__ row_flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    int64_t gr = row_begin + i;
    size_t u = (size_t)row_node[gr] * W, l = (size_t)row_left[gr] * W, r = (size_t)row_right[gr] * W;
    bool sub0 = true, sub1 = true;
    for (int w = 0; w < W; ++w) {
        uint32_t Ll = L[l + w], Lr = L[r + w];
        uint32_t mine = (Ll | Lr) & ~(gain[l + w] | gain[r + w]);
        L[u + w] = mine;
        sub0 &= (mine & ~Ll) == 0;
        sub1 &= (mine & ~Lr) == 0;
    }
    row_flags[gr] = (uint8_t)((sub0 ? 1 : 0) | (sub1 ? 2 : 0));
}

// All four passes above for ONE instance in one launch (one CTA, a barrier per wave): a single tree has
// nothing to fill the machine with, so its 3 H + 1 launches were pure launch latency (config #4: 61
// launches, 0.2 ms of a 0.86 ms solve).  Rows of wave h = [off[h], off[h+1]).
#define SYN_SINGLE_MAXH 120
struct SynWaves {
    int H;
    int32_t off[SYN_SINGLE_MAXH + 2];
};
__global__ void __launch_bounds__(1024)
syn_single_kernel(SynWaves wo, const int32_t* __restrict__ row_node, const int32_t* __restrict__ row_left,
                  const int32_t* __restrict__ row_right, const int32_t* __restrict__ left,
                  const int32_t* __restrict__ right, int N, int W, const uint32_t* __restrict__ bits,
                  uint32_t* __restrict__ present, uint32_t* __restrict__ outside, uint32_t* __restrict__ gain,
                  uint32_t* __restrict__ L, uint8_t* __restrict__ row_flags) {
    const int t = threadIdx.x, T = blockDim.x;
    for (int i = t; i < N * W; i += T) {
        const uint32_t b = bits[i];
        present[i] = b;
        L[i] = b;
        outside[i] = 0;
    }
    __syncthreads();
    for (int h = 1; h <= wo.H; ++h) {   // syn_present_wave
        const int r0 = wo.off[h], n = wo.off[h + 1] - r0;
        for (int i = t; i < n * W; i += T) {
            const int gr = r0 + i / W, w = i % W;
            present[(size_t)row_node[gr] * W + w] =
                present[(size_t)row_left[gr] * W + w] | present[(size_t)row_right[gr] * W + w];
        }
        __syncthreads();
    }
    for (int h = wo.H; h >= 1; --h) {   // syn_outside_wave
        const int r0 = wo.off[h], n = wo.off[h + 1] - r0;
        for (int i = t; i < n * W; i += T) {
            const int gr = r0 + i / W, w = i % W;
            const uint32_t mine = outside[(size_t)row_node[gr] * W + w];
            const size_t l = (size_t)row_left[gr] * W + w, r = (size_t)row_right[gr] * W + w;
            outside[l] = mine | present[r];
            outside[r] = mine | present[l];
        }
        __syncthreads();
    }
    for (int i = t; i < N * W; i += T) {   // syn_gain_kernel
        const int u = i / W, w = i % W;
        uint32_t full = present[i] & ~outside[i];
        const int l = left[u];
        if (l >= 0) {
            const size_t li = (size_t)l * W + w, ri = (size_t)right[u] * W + w;
            full &= ~((present[li] & ~outside[li]) | (present[ri] & ~outside[ri]));
        }
        gain[i] = full;
    }
    __syncthreads();
    for (int h = 1; h <= wo.H; ++h) {   // syn_lca_wave
        const int r0 = wo.off[h], n = wo.off[h + 1] - r0;
        for (int i = t; i < n; i += T) {
            const int gr = r0 + i;
            const size_t u = (size_t)row_node[gr] * W, l = (size_t)row_left[gr] * W, r = (size_t)row_right[gr] * W;
            bool sub0 = true, sub1 = true;
            for (int w = 0; w < W; ++w) {
                const uint32_t Ll = L[l + w], Lr = L[r + w];
                const uint32_t mine = (Ll | Lr) & ~(gain[l + w] | gain[r + w]);
                L[u + w] = mine;
                sub0 &= (mine & ~Ll) == 0;
                sub1 &= (mine & ~Lr) == 0;
            }
            row_flags[gr] = (uint8_t)((sub0 ? 1 : 0) | (sub1 ? 2 : 0));
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// the table
// ---------------------------------------------------------------------------
// superdtl keys: value[63:32] | rank[31:16] | kind[15] | depth[14:0]; order (value, rank, kind)
__device__ __forceinline__ u64 skey(int value, int rank, int kind, int depth) {
    return ((u64)(unsigned)value << 32) | ((u64)(unsigned)rank << 16) | ((unsigned)kind << 15) | (unsigned)depth;
}
__device__ __forceinline__ int sat_add(int a, int b) { return (a >= SR2_INF || b >= SR2_INF) ? SR2_INF : a + b; }

template <int TPR>
__device__ __forceinline__ void sgroup_sync() {
    if (TPR == 32)
        __syncwarp();
    else
        __syncthreads();
}

// one candidate: children at keys a / b; da, db = depth the cons/seg value is relative to,
// or -1 for a "separate" key (raw value, no distance term).
struct SuperPick {
    int best;       // value of the cell so far
    unsigned tag;   // (species << 1 | kind) of the left child, the same of the right child << 16
    int which;      // candidate that won (0..5, the order of :289-297)
    u64 ka, kb;     // its two keys (value | rank | kind | depth)
};

__device__ __forceinline__ void super_offer(u64 a, int da, u64 b, int db, int event, int floss, int which,
                                            SuperPick& pick) {
    unsigned va = (unsigned)(a >> 32), vb = (unsigned)(b >> 32);
    if (va < (unsigned)SR2_INF && vb < (unsigned)SR2_INF) {
        int v = event + (int)va + (int)vb - floss * ((da < 0 ? 0 : da) + (db < 0 ? 0 : db));
        if (v < pick.best) {
            pick.best = v;
            pick.tag = (unsigned)((a >> 15) & 0xffffu) | ((unsigned)((b >> 15) & 0xffffu) << 16);
            pick.which = which;
            pick.ka = a;
            pick.kb = b;
        }
    }
}

// arrays of one row in shared memory: [child i][parent kind K][cons, seg, sep]
#define SARR(base, spad, i, K, t) ((base) + (size_t)((((i) * 2 + (K)) * 3) + (t)) * (spad))
enum { T_CONS = 0, T_SEG = 1, T_SEP = 2 };

// Cell (u, x, K).  P0/P1: SEPMIN keys of child 0/1 at x (none at the root).  Candidate order = :289-297.
__device__ __forceinline__ void super_cell(const u64* base, int spad, int K, int x, int dx, int cx, u64 P0, u64 P1,
                                           const SuperCosts& cs, SuperPick& pick) {
    const u64* C0 = SARR(base, spad, 0, K, T_CONS);
    const u64* C1 = SARR(base, spad, 1, K, T_CONS);
    const u64* G0 = SARR(base, spad, 0, K, T_SEG);
    const u64* G1 = SARR(base, spad, 1, K, T_SEG);
    pick.best = SR2_INF;
    pick.tag = SR2_NO_TAG;
    pick.which = 0;
    pick.ka = pick.kb = 0;
    if (cx >= 0) {
        super_offer(C0[cx], dx + 1, C1[cx + 1], dx + 1, cs.spe, cs.floss, 0, pick);  // left x right
        super_offer(C0[cx + 1], dx + 1, C1[cx], dx + 1, cs.spe, cs.floss, 1, pick);  // right x left
    }
    u64 c0 = C0[x], c1 = C1[x];
    super_offer(c0, dx, G1[x], dx, cs.dup, cs.floss, 2, pick);  // conserved x segment
    super_offer(G0[x], dx, c1, dx, cs.dup, cs.floss, 3, pick);  // segment x conserved
    super_offer(c0, dx, P1, -1, cs.hgt, cs.floss, 4, pick);     // conserved x separate
    super_offer(P0, -1, c1, dx, cs.hgt, cs.floss, 5, pick);     // separate x conserved
}

// What SuperReconciliationOutput.cost() (model/reconciliation.py:273-365, 511-543) charges for a
// node decoded through this cell, as far as it only depends on the cell: the event term
// (node_event of (x, a, c) + its loss distances) << 2 | which segmental-loss terms the labelling adds:
//   0 speciation: lc + rc    1 duplication: min(lc, rc)    2 transfer, left child kept: lc
//   3 transfer, right child kept: rc.
// The candidate that won fixes the relation of a and c to x: 0-1 sit below different children of x
// (speciation), 2-3 both below x (speciation iff neither is x and they sit below different
// children, else duplication -- cost() classifies the decoded triple, not the candidate), 4 keeps the
// left child below x, 5 the right one (its left child is incomparable with x, never above it).
enum { EV_CLASS_SPE = 0, EV_CLASS_DUP = 1, EV_CLASS_HGT_L = 2, EV_CLASS_HGT_R = 3 };
__device__ __forceinline__ uint32_t super_event_word(const DevSpecies& sp, int x, int dx, int cx,
                                                     const SuperPick& pick, const SuperCosts& cs) {
    const int a = (int)((pick.ka >> 16) & 0xffffu), c = (int)((pick.kb >> 16) & 0xffffu);
    const int da = (int)(pick.ka & 0x7fffu), dc = (int)(pick.kb & 0x7fffu);
    int cost, cls;
    if (pick.which >= 4) {
        cls = pick.which == 4 ? EV_CLASS_HGT_L : EV_CLASS_HGT_R;
        cost = cs.hgt + cs.floss * ((pick.which == 4 ? da : dc) - dx);
    } else {
        bool spec = pick.which < 2;
        if (!spec && cx >= 0 && a != x && c != x) {
            const int t1 = __ldg(sp.tin + cx + 1);
            spec = (__ldg(sp.tin + a) < t1) != (__ldg(sp.tin + c) < t1);
        }
        cls = spec ? EV_CLASS_SPE : EV_CLASS_DUP;
        cost = (spec ? cs.spe - 2 * cs.floss : cs.dup) + cs.floss * (da + dc - 2 * dx);
    }
    return ((uint32_t)cost << 2) | (uint32_t)cls;
}

// Topology of the species tree staged once per CTA behind the key rows, so that the level loops of
// the sweeps never wait for global memory (ncu r3b: two dependent __ldg per level made a wave of
// config #4 cost 56 us, two thirds of it at the level barriers):
//   s_meta[S]     (child0 + 1) << 16 | depth      (DevSpecies::meta)
//   s_lvl[D + 1]  offsets into the list of internal nodes per depth
//   s_int[n_int]  level-order rank of the i-th internal node (its children are 2i+1, 2i+2)
__host__ __device__ __forceinline__ size_t super_topology_bytes(int S, int D, int n_int) {
    return ((size_t)S + (size_t)D + 1) * 4 + (((size_t)n_int + 1) & ~(size_t)1) * 2;
}

// One row = both kinds of one object node.  Three phases over the 12 key rows of the node's children:
//   1. up-sweep of all rows (SUBMIN), one step per species level;
//   2. top-down sweep of the four "separate" rows only (SEPMIN, in place);
//   3. the cells, element-wise over the species with every lane busy: 6 candidates per kind.
// TPR = 32: one warp per row (small species trees); TPR = 256: one CTA per row.
// CHAIN (one CTA per row only): all waves in ONE launch, rows handed out by ticket in height order, a row
// waits for the "done" flags of its child rows and reads them past L1 (see dtl_wave_kernel in sr2_dtl.cu).
template <int TPR, bool CHAIN = false>
__global__ void __launch_bounds__(TPR == 32 ? 256 : TPR)
superdtl_wave_kernel(DevSpecies sp, const int32_t* __restrict__ row_cl, const int32_t* __restrict__ row_cr,
                     const uint8_t* __restrict__ row_flags, int64_t row_begin, int64_t chunk_row0, int n_rows,
                     int32_t* __restrict__ T, uint32_t* __restrict__ tags, SuperCosts cs, int spad,
                     const int32_t* __restrict__ row_node, const int32_t* __restrict__ allowed,
                     uint32_t* __restrict__ evs, int* __restrict__ chain_sync = nullptr) {
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
    uint32_t* s_meta = reinterpret_cast<uint32_t*>(smem + (size_t)groups_per_cta * 12 * spad);
    int32_t* s_lvl = reinterpret_cast<int32_t*>(s_meta + S);
    uint16_t* s_int = reinterpret_cast<uint16_t*>(s_lvl + sp.D + 1);
    for (int i = threadIdx.x; i < S; i += blockDim.x) s_meta[i] = __ldg(sp.meta + i);
    for (int i = threadIdx.x; i <= sp.D; i += blockDim.x) s_lvl[i] = __ldg(sp.int_lvl_off + i);
    for (int i = threadIdx.x; i < sp.n_internal; i += blockDim.x) s_int[i] = (uint16_t)__ldg(sp.int_list + i);
    if (CHAIN && threadIdx.x == 0) s_ticket = atomicAdd(chain_sync, 1);
    __syncthreads();
    int local = CHAIN ? s_ticket : blockIdx.x * groups_per_cta + group;
    if (local >= n_rows) return;
    const int64_t gr = row_begin + local;
    const size_t slot = (size_t)(gr - chunk_row0);
    u64* base = smem + (size_t)group * 12 * spad;
    if (CHAIN) {
        const int first_slot = (int)(row_begin - chunk_row0);
        const int c = threadIdx.x == 0 ? row_cl[gr] : row_cr[gr];
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
    const int flags = row_flags[gr];
    SR2_DCHECK(row_cl[gr] < (int)slot && row_cr[gr] < (int)slot && row_cl[gr] >= -S && row_cr[gr] >= -S);
    SR2_DCHECK(flags >= 0 && flags < 4);
    // allowed_species of _compute_uspfs_table (:340-356): every species (superdtl) or one species per
    // object node (base_uspfs); cells of other species are never created
    const int only = allowed ? allowed[row_node[gr]] : -1;

    // stage the child rows as the 12 key rows
    for (int i = 0; i < 2; ++i) {
        const int c = i == 0 ? row_cl[gr] : row_cr[gr];
        const bool sub = (flags >> i) & 1;               // L[u] <= L[child]  (:182)
        const int ll = sub ? 0 : cs.sloss;               // lca_lca_dist
        const int li = sub ? SR2_INF : 0;                // lca_inh_dist
        const int32_t* __restrict__ TA = c >= 0 ? T + ((size_t)c * 2 + 0) * sp.pitch : nullptr;
        const int32_t* __restrict__ TB = c >= 0 ? T + ((size_t)c * 2 + 1) * sp.pitch : nullptr;
        for (int y = lane; y < S; y += TPR) {
            int d = (int)(s_meta[y] & 0xffffu);
            int Araw = TA ? (CHAIN ? __ldcg(TA + y) : TA[y]) : (y == -1 - c ? 0 : SR2_INF);
            int Braw = TB ? (CHAIN ? __ldcg(TB + y) : TB[y]) : SR2_INF;
            int A = sat_add(Araw, cs.floss * d), B = sat_add(Braw, cs.floss * d);
            // parent kind INHERIT: (sloss, 0, 0)
            SARR(base, spad, i, 1, T_CONS)[y] = key_min(skey(sat_add(A, cs.sloss), y, 0, d), skey(B, y, 1, d));
            SARR(base, spad, i, 1, T_SEG)[y] = key_min(skey(A, y, 0, d), skey(B, y, 1, d));
            SARR(base, spad, i, 1, T_SEP)[y] = key_min(skey(Araw, y, 0, d), skey(Braw, y, 1, d));
            // parent kind LCA: (lca_lca_dist, lca_inh_dist, lca_inh_dist)
            SARR(base, spad, i, 0, T_CONS)[y] =
                key_min(skey(sat_add(A, ll), y, 0, d), skey(sat_add(B, li), y, 1, d));
            SARR(base, spad, i, 0, T_SEG)[y] = key_min(skey(A, y, 0, d), skey(sat_add(B, li), y, 1, d));
            SARR(base, spad, i, 0, T_SEP)[y] = key_min(skey(Araw, y, 0, d), skey(sat_add(Braw, li), y, 1, d));
        }
    }
    sgroup_sync<TPR>();

    // 1. up-sweep of all 12 rows.  A work item is (internal species of the level, key row): the levels of a
    // species tree are narrow (config #5: 3.6 internal species per level on average), so lanes over species
    // alone left 7 of 8 lanes idle while the busy ones walked the 12 rows one after the other.
    for (int d = sp.D - 2; d >= 0; --d) {
        const int b = s_lvl[d], items = (s_lvl[d + 1] - b) * 12;
        for (int it = lane; it < items; it += TPR) {
            const int i = b + it / 12, a = it % 12;
            const int p = s_int[i], c = 2 * i + 1;
            SR2_DCHECK(i < sp.n_internal && p < S && c + 1 < S && p < c);
            u64* arr = base + (size_t)a * spad;
            arr[p] = key_min(arr[p], key_min(arr[c], arr[c + 1]));
        }
        sgroup_sync<TPR>();
    }

    // 2. "separate" rows: SEPMIN[child] = min(SEPMIN[parent], SUBMIN[sibling]), nothing at the root
    if (lane < 4) SARR(base, spad, lane >> 1, lane & 1, T_SEP)[0] = SR2_KEY_NONE;
    sgroup_sync<TPR>();
    for (int d = 0; d < sp.D - 1; ++d) {
        const int b = s_lvl[d], items = (s_lvl[d + 1] - b) * 4;
        for (int it = lane; it < items; it += TPR) {
            const int i = b + (it >> 2), a = it & 3;
            const int p = s_int[i], c = 2 * i + 1;
            u64* E = SARR(base, spad, a >> 1, a & 1, T_SEP);
            const u64 Pp = E[p], m0 = E[c], m1 = E[c + 1];
            E[c] = key_min(Pp, m1);
            E[c + 1] = key_min(Pp, m0);
        }
        sgroup_sync<TPR>();
    }

    // 3. the cells
    int32_t* __restrict__ Tout0 = T + (slot * 2 + 0) * sp.pitch;
    int32_t* __restrict__ Tout1 = T + (slot * 2 + 1) * sp.pitch;
    uint32_t* __restrict__ Gout0 = tags + (slot * 2 + 0) * sp.pitch;
    uint32_t* __restrict__ Gout1 = tags + (slot * 2 + 1) * sp.pitch;
    for (int x = lane; x < S; x += TPR) {
        const uint32_t m = s_meta[x];
        const int dx = (int)(m & 0xffffu), cx = (int)(m >> 16) - 1;
        const bool skip = only >= 0 && only != x;
#pragma unroll
        for (int K = 0; K < 2; ++K) {
            SuperPick pick;
            super_cell(base, spad, K, x, dx, cx, SARR(base, spad, 0, K, T_SEP)[x], SARR(base, spad, 1, K, T_SEP)[x],
                       cs, pick);
            const bool none = skip || pick.tag == SR2_NO_TAG;
            (K ? Tout1 : Tout0)[x] = skip ? SR2_INF : pick.best;
            (K ? Gout1 : Gout0)[x] = none ? SR2_NO_TAG : pick.tag;
            // the fused decode kernel reads what cost() charges for the cell next to its tag
            if (evs) evs[(slot * 2 + K) * sp.pitch + x] = none ? 0u : super_event_word(sp, x, dx, cx, pick, cs);
        }
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

// ---------------------------------------------------------------------------
// decode from every root species, Hazard-A time stamps, cost(), selection
// ---------------------------------------------------------------------------
// dec[node*S + k] = (species << 1 | kind) of `node` in the decode that starts at root
// species k; anc[node*S + k] = instance-local index of the node whose lca-set the node
// labels itself with (itself for kind LCA, else its nearest LCA-kind ancestor).
__global__ void sdecode_root_kernel(DevSpecies sp, const int64_t* __restrict__ root_row,
                                    const int32_t* __restrict__ root_node, int inst_begin, int n_inst,
                                    int64_t chunk_row0, const uint32_t* __restrict__ tags,
                                    uint16_t* __restrict__ dec, uint16_t* __restrict__ anc) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_inst * sp.S) return;
    int b = inst_begin + (int)(i / sp.S), k = (int)(i % sp.S);
    int64_t gr = root_row[b];
    bool ok;
    if (gr < 0)
        ok = k == (int)(-1 - gr);  // one-node tree: only its own species has an entry (:393-404)
    else
        ok = tags[((size_t)(gr - chunk_row0) * 2 + 0) * sp.pitch + k] != SR2_NO_TAG;  // kind LCA (:487)
    size_t at = (size_t)root_node[b] * sp.S + k;
    dec[at] = ok ? (uint16_t)(k << 1) : (uint16_t)DEC_NONE;
    anc[at] = 0;
}

__global__ void sdecode_wave_kernel(DevSpecies sp, const int32_t* __restrict__ row_node,
                                    const int32_t* __restrict__ row_left, const int32_t* __restrict__ row_right,
                                    const int64_t* __restrict__ node_off, const int32_t* __restrict__ node_inst,
                                    const int32_t* __restrict__ node_pre, int64_t row_begin, int64_t chunk_row0,
                                    int n_rows, const uint32_t* __restrict__ tags, int W,
                                    const uint32_t* __restrict__ gain, uint16_t* __restrict__ dec,
                                    uint16_t* __restrict__ anc, uint32_t* __restrict__ tau,
                                    uint32_t* __restrict__ taumin) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_rows * sp.S) return;
    int64_t gr = row_begin + i / sp.S;
    int k = (int)(i % sp.S);
    size_t slot = (size_t)(gr - chunk_row0);
    int node = row_node[gr], l = row_left[gr], r = row_right[gr];
    unsigned e = dec[(size_t)node * sp.S + k];
    unsigned el = DEC_NONE, er = DEC_NONE;
    unsigned al = 0, ar = 0;
    if (e != DEC_NONE) {
        int x = e >> 1, kd = e & 1;
        unsigned tag = tags[(slot * 2 + kd) * sp.pitch + x];
        if (tag != SR2_NO_TAG) {
            el = tag & 0xffffu;
            er = tag >> 16;
            int64_t off = node_off[node_inst[node]];
            unsigned mine = anc[(size_t)node * sp.S + k];
            al = (el & 1) ? mine : (unsigned)(l - off);
            ar = (er & 1) ? mine : (unsigned)(r - off);
            // kind INHERIT: the node's gains are merged, in place, into the set of its owner at
            // time (root k, preorder position) and stay there for every later decode (:390)
#pragma unroll 1
            for (int side = 0; side < 2; ++side) {
                unsigned ec = side ? er : el;
                if (!(ec & 1)) continue;
                int child = side ? r : l;
                unsigned owner = side ? ar : al;
                unsigned stamp = ((unsigned)k << 16) | (unsigned)node_pre[child];
                bool any = false;
                for (int w = 0; w < W; ++w) {
                    uint32_t g = gain[(size_t)child * W + w];
                    while (g) {
                        int f = __ffs(g) - 1;
                        g &= g - 1;
                        atomicMin(&tau[((size_t)(off + owner) * W + w) * 32 + f], stamp);
                        any = true;
                    }
                }
                if (any) atomicMin(&taumin[off + owner], stamp);
            }
        }
    }
    dec[(size_t)l * sp.S + k] = (uint16_t)el;
    dec[(size_t)r * sp.S + k] = (uint16_t)er;
    anc[(size_t)l * sp.S + k] = (uint16_t)al;
    anc[(size_t)r * sp.S + k] = (uint16_t)ar;
}

// the synteny a node is labelled with at time `stamp`: clean lca-set of its owner plus
// every family merged into that set at or before `stamp`
__device__ __forceinline__ uint32_t syn_word(const uint32_t* __restrict__ L, size_t L_index,
                                             const uint32_t* __restrict__ tau,
                                             const uint32_t* __restrict__ taumin, size_t owner, int W, int w,
                                             unsigned stamp) {
    uint32_t bits = L[L_index * W + w];
    if (taumin[owner] <= stamp) {
        const uint32_t* t = tau + (owner * W + w) * 32;
        for (int f = 0; f < 32; ++f)
            if (t[f] <= stamp) bits |= 1u << f;
    }
    return bits;
}

__device__ __forceinline__ bool sp_anc(const DevSpecies& sp, int a, int b) {  // a is b or above b
    int ta = __ldg(sp.tin + a), tb = __ldg(sp.tin + b);
    return ta <= tb && tb < __ldg(sp.tout + a);
}

// cost() of every decode (model/reconciliation.py:511-554): thread = (chunk of SCOST_NODES consecutive
// nodes, root species k) adds the terms of its internal nodes to cost[instance * S + k] (zeroed by the
// caller).  One thread per root walking the whole tree took 1.5 ms for ONE 1,999-node instance.
#define SCOST_NODES 16
__global__ void __launch_bounds__(128)
scost_kernel(DevSpecies sp, const int64_t* __restrict__ node_off, const int32_t* __restrict__ left,
             const int32_t* __restrict__ right, const int32_t* __restrict__ node_pre,
             const int32_t* __restrict__ node_inst, int64_t n_nodes, int W, const uint32_t* __restrict__ L,
             const uint32_t* __restrict__ tau, const uint32_t* __restrict__ taumin,
             const uint16_t* __restrict__ dec, const uint16_t* __restrict__ anc, SuperCosts cs,
             int32_t* __restrict__ cost, const int32_t* __restrict__ node_dag) {
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= sp.S) return;
    const int64_t u0 = (int64_t)blockIdx.x * SCOST_NODES, u1 = min(u0 + (int64_t)SCOST_NODES, n_nodes);
    int cur = -1, total = 0;
    bool valid = false;
    int64_t off = 0;
    for (int64_t u = u0; u < u1; ++u) {
        int l = left[u];
        if (l < 0) continue;
        const int b = node_inst[u];
        if (b != cur) {
            if (valid && total) atomicAdd(&cost[(size_t)cur * sp.S + k], total);
            cur = b;
            total = 0;
            off = node_off[b];
            valid = dec[(size_t)off * sp.S + k] != DEC_NONE;  // roots without an entry decode nothing
        }
        if (!valid) continue;
        int r = right[u];
        unsigned eu = dec[(size_t)u * sp.S + k], el = dec[(size_t)l * sp.S + k], er = dec[(size_t)r * sp.S + k];
        int p = eu >> 1, a = el >> 1, c = er >> 1;
        SR2_DCHECK(l > u && r > u && r < n_nodes && eu != DEC_NONE && el != DEC_NONE && er != DEC_NONE);
        SR2_DCHECK(p < sp.S && a < sp.S && c < sp.S);
        int dp = __ldg(sp.depth + p), da = __ldg(sp.depth + a), dc = __ldg(sp.depth + c);
        bool pa = sp_anc(sp, p, a), pc = sp_anc(sp, p, c);
        // labelling cost (model/reconciliation.py:511-543)
        size_t ou = (size_t)off + anc[(size_t)u * sp.S + k], ol = (size_t)off + anc[(size_t)l * sp.S + k],
               orr = (size_t)off + anc[(size_t)r * sp.S + k];
        unsigned su = ((unsigned)k << 16) | (unsigned)node_pre[u], sl = ((unsigned)k << 16) | (unsigned)node_pre[l],
                 sr = ((unsigned)k << 16) | (unsigned)node_pre[r];
        bool in_l = true, in_r = true;
        // shared-row mode: the clean lca-sets live per dag node, the time stamps per tree node
        size_t Lu = node_dag ? (size_t)node_dag[ou] : ou, Ll = node_dag ? (size_t)node_dag[ol] : ol,
               Lr = node_dag ? (size_t)node_dag[orr] : orr;
        for (int w = 0; w < W; ++w) {
            uint32_t mine = syn_word(L, Lu, tau, taumin, ou, W, w, su);
            in_l &= (mine & ~syn_word(L, Ll, tau, taumin, ol, W, w, sl)) == 0;
            in_r &= (mine & ~syn_word(L, Lr, tau, taumin, orr, W, w, sr)) == 0;
        }
        int lc = in_l ? 0 : cs.sloss, rc = in_r ? 0 : cs.sloss;
        // event (model/reconciliation.py:273-365); decodes of finite entries are never INVALID
        if (pa && pc) {
            int cx = __ldg(sp.child0 + p);
            bool spec = false;
            if (cx >= 0 && a != p && c != p) {
                int t1 = __ldg(sp.tin + cx + 1);
                spec = (__ldg(sp.tin + a) < t1) != (__ldg(sp.tin + c) < t1);
            }
            if (spec)
                total += cs.spe + cs.floss * (da + dc - 2 * dp - 2) + lc + rc;
            else
                total += cs.dup + cs.floss * (da + dc - 2 * dp) + min(lc, rc);
        } else if (pa) {
            total += cs.hgt + cs.floss * (da - dp) + lc;  // left child kept (comparable)
        } else {
            // transfer with the right child kept; the left child's species is comparable
            // with p only if it is an ancestor of p, which a decode never produces
            total += cs.hgt + cs.floss * (dc - dp) + (sp_anc(sp, a, p) ? lc : rc);
        }
    }
    if (valid && total) atomicAdd(&cost[(size_t)cur * sp.S + k], total);
}

// first strict minimum over the root species in level order (_uspfs :474-495), one warp per instance
__global__ void __launch_bounds__(256)
sselect_kernel(DevSpecies sp, int inst_begin, int n_inst, const int32_t* __restrict__ cost,
               const uint16_t* __restrict__ dec, const int64_t* __restrict__ node_off,
               int32_t* __restrict__ min_cost, int32_t* __restrict__ root_species) {
    int local = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (local >= n_inst) return;
    int b = inst_begin + local;
    u64 best = SR2_KEY_NONE;
    for (int k = lane; k < sp.S; k += 32)
        if (dec[(size_t)node_off[b] * sp.S + k] != DEC_NONE)
            best = key_min(best, ((u64)(unsigned)cost[(size_t)b * sp.S + k] << 32) | (unsigned)k);
    for (int off = 16; off; off >>= 1) best = key_min(best, __shfl_xor_sync(0xffffffffu, best, off));
    if (lane == 0) {
        bool none = best == SR2_KEY_NONE;
        min_cost[b] = none ? SR2_INF : (int)(best >> 32);
        root_species[b] = none ? -1 : (int)(best & 0xffffffffu);
    }
}

// species, kind and synteny of every node in the decode chosen for its instance
__global__ void smaterialise_kernel(DevSpecies sp, int64_t node_begin, int64_t n_nodes,
                                    const int32_t* __restrict__ node_inst, const int64_t* __restrict__ node_off,
                                    const int32_t* __restrict__ node_pre, const int32_t* __restrict__ root_species,
                                    int W, const uint32_t* __restrict__ L, const uint32_t* __restrict__ tau,
                                    const uint32_t* __restrict__ taumin, const uint16_t* __restrict__ dec,
                                    const uint16_t* __restrict__ anc, int32_t* __restrict__ map,
                                    int32_t* __restrict__ kind, uint32_t* __restrict__ syn,
                                    const int32_t* __restrict__ node_dag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    int64_t u = node_begin + i;
    int b = node_inst[u];
    int k = root_species[b];
    if (k < 0) {
        map[u] = -1;
        kind[u] = -1;
        for (int w = 0; w < W; ++w) syn[(size_t)u * W + w] = 0;
        return;
    }
    unsigned e = dec[(size_t)u * sp.S + k];
    map[u] = e >> 1;
    kind[u] = e & 1;
    size_t owner = (size_t)node_off[b] + anc[(size_t)u * sp.S + k];
    unsigned stamp = ((unsigned)k << 16) | (unsigned)node_pre[u];
    size_t L_index = node_dag ? (size_t)node_dag[owner] : owner;
    for (int w = 0; w < W; ++w) syn[(size_t)u * W + w] = syn_word(L, L_index, tau, taumin, owner, W, w, stamp);
}

// ---------------------------------------------------------------------------
// shared-row mode: resolved trees on the device, decode by walking each tree in id order
// ---------------------------------------------------------------------------
struct DevPlan {
    int n, Gb;
    int64_t first;
    const int32_t *arity, *base, *child_off, *child_of, *arr_off, *arr, *perm;
    const int32_t *arr_pre_off, *arr_pre, *parent, *child_pos;   // see SharedPlan::arr_pre
    const int64_t *ways, *stride, *cnt, *var_first, *raw_base;
};

struct DevPlanStore {
    DevPlan plan;
};
SuperProblem::~SuperProblem() { delete plan_store; }

// One original node v of resolution r: its block of k - 1 joints (local ids inside the resolved tree).
// out_left / out_right / out_slot / out_dag are indexed by local id.
// (r / stride) % cnt for resolution numbers below 2^24 (checked at upload) in 32-bit arithmetic: 64-bit
// divisions are function calls on the device and every node of every tree does several of them
__device__ __forceinline__ int64_t unrank_digit(int64_t r, int64_t stride, int64_t cnt) {
    if (stride > r) return 0;
    const uint32_t q = (uint32_t)r / (uint32_t)stride;
    return cnt > (int64_t)q ? (int64_t)q : (int64_t)(q % (uint32_t)cnt);
}

__device__ __forceinline__ void unrank_node(const DevPlan& P, int64_t r, int v, int32_t* out_left, int32_t* out_right,
                                            int32_t* out_slot, int32_t* out_dag, int32_t* out_pre, int32_t id_offset) {
    const int k = P.arity[v];
    const int id0 = P.base[v];
    // preorder position of the root of v's block: up the original tree, adding where each ancestor's
    // arrangement puts the item we came from
    int abs_pre = 0;
    for (int cur = v; cur != 0;) {
        const int par = P.parent[cur], kp = P.arity[par];
        const int64_t jp = (int64_t)((uint32_t)unrank_digit(r, P.stride[par], P.cnt[par]) % (uint32_t)P.ways[par]);
        abs_pre += P.arr_pre[P.arr_pre_off[par] + jp * (2 * kp - 1) + (kp - 1) + P.child_pos[cur]];
        cur = par;
    }
    if (k == 0) {
        out_left[id0] = out_right[id0] = -1;
        out_slot[id0] = -1;
        out_dag[id0] = id0;
        out_pre[id0] = abs_pre;
        return;
    }
    const int64_t var = unrank_digit(r, P.stride[v], P.cnt[v]);
    const int64_t j = (int64_t)((uint32_t)var % (uint32_t)P.ways[v]);
    const int32_t* rel = P.arr_pre + P.arr_pre_off[v] + j * (2 * k - 1);
    const int64_t raw0 = P.raw_base[v] + (var - P.var_first[v]) * (k - 1);
    const int32_t* codes = P.arr + P.arr_off[v] + j * (k - 1) * 2;
    const int32_t* kids = P.child_of + P.child_off[v];
    for (int q = 0; q < k - 1; ++q) {
        const int cl = codes[2 * q], cr = codes[2 * q + 1];
        out_left[id0 + q] = id_offset + (cl >= 0 ? P.base[kids[cl]] : id0 + (-1 - cl));
        out_right[id0 + q] = id_offset + (cr >= 0 ? P.base[kids[cr]] : id0 + (-1 - cr));
        const int sl = P.perm[raw0 + q];
        out_slot[id0 + q] = sl;
        out_dag[id0 + q] = P.Gb + sl;
        out_pre[id0 + q] = abs_pre + rel[q];
    }
}

// Resolution first + i of the plan, as global-id child arrays, table row and dag id of every node,
// preorder positions and instance numbers.  One thread per resolution.
__global__ void resolve_unrank_kernel(DevPlan P, int64_t count, int32_t* __restrict__ left,
                                      int32_t* __restrict__ right, int32_t* __restrict__ node_slot,
                                      int32_t* __restrict__ node_dag, int32_t* __restrict__ node_pre,
                                      int32_t* __restrict__ node_inst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t r = P.first + i;
    const int64_t off = i * P.Gb;
    for (int v = 0; v < P.n; ++v) {
        const int k = P.arity[v];
        const int id0 = P.base[v];
        const int64_t g0 = off + id0;
        if (k == 0) {
            left[g0] = right[g0] = -1;
            node_slot[g0] = -1;
            node_dag[g0] = id0;
            continue;
        }
        const int64_t var = (r / P.stride[v]) % P.cnt[v];
        const int64_t j = var % P.ways[v];
        const int64_t raw0 = P.raw_base[v] + (var - P.var_first[v]) * (k - 1);
        const int32_t* codes = P.arr + P.arr_off[v] + j * (k - 1) * 2;
        const int32_t* kids = P.child_of + P.child_off[v];
        for (int q = 0; q < k - 1; ++q) {
            int cl = codes[2 * q], cr = codes[2 * q + 1];
            left[g0 + q] = (int32_t)(off + (cl >= 0 ? P.base[kids[cl]] : id0 + (-1 - cl)));
            right[g0 + q] = (int32_t)(off + (cr >= 0 ? P.base[kids[cr]] : id0 + (-1 - cr)));
            int s = P.perm[raw0 + q];
            node_slot[g0 + q] = s;
            node_dag[g0 + q] = P.Gb + s;
        }
    }
    // preorder positions from subtree sizes (children have larger ids than their parent)
    for (int u = P.Gb - 1; u >= 0; --u) {
        int l = left[off + u];
        node_inst[off + u] = l < 0 ? 1 : 1 + node_inst[l] + node_inst[right[off + u]];
    }
    node_pre[off] = 0;
    for (int u = 0; u < P.Gb; ++u) {
        int l = left[off + u];
        if (l < 0) continue;
        int mine = node_pre[off + u];
        node_pre[l] = mine + 1;
        node_pre[right[off + u]] = mine + 1 + node_inst[l];
    }
    for (int u = 0; u < P.Gb; ++u) node_inst[off + u] = (int32_t)i;
}

// sdecode_root_kernel + sdecode_wave_kernel for uniform trees whose nodes are numbered parents
// first: thread (instance b, root species k) walks the tree in id order.
__global__ void sdecode_walk_kernel(DevSpecies sp, int n_inst, int Gb, const int32_t* __restrict__ left,
                                    const int32_t* __restrict__ right, const int32_t* __restrict__ node_slot,
                                    const int32_t* __restrict__ node_dag, const int32_t* __restrict__ node_pre,
                                    const uint32_t* __restrict__ tags, int W, const uint32_t* __restrict__ gain,
                                    uint16_t* __restrict__ dec, uint16_t* __restrict__ anc,
                                    uint32_t* __restrict__ tau, uint32_t* __restrict__ taumin) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_inst * sp.S) return;
    const int b = (int)(i / sp.S), k = (int)(i % sp.S);
    const int64_t off = (int64_t)b * Gb;
    {
        bool ok = tags[((size_t)node_slot[off] * 2 + 0) * sp.pitch + k] != SR2_NO_TAG;  // kind LCA (:487)
        dec[(size_t)off * sp.S + k] = ok ? (uint16_t)(k << 1) : (uint16_t)DEC_NONE;
        anc[(size_t)off * sp.S + k] = 0;
    }
    for (int u = 0; u < Gb; ++u) {
        const int64_t node = off + u;
        const int l = left[node];
        if (l < 0) continue;
        const int r = right[node];
        unsigned e = dec[(size_t)node * sp.S + k];
        unsigned el = DEC_NONE, er = DEC_NONE;
        unsigned al = 0, ar = 0;
        if (e != DEC_NONE) {
            int x = e >> 1, kd = e & 1;
            unsigned tag = tags[((size_t)node_slot[node] * 2 + kd) * sp.pitch + x];
            if (tag != SR2_NO_TAG) {
                el = tag & 0xffffu;
                er = tag >> 16;
                unsigned mine = anc[(size_t)node * sp.S + k];
                al = (el & 1) ? mine : (unsigned)(l - off);
                ar = (er & 1) ? mine : (unsigned)(r - off);
                // kind INHERIT: see sdecode_wave_kernel
#pragma unroll 1
                for (int side = 0; side < 2; ++side) {
                    unsigned ec = side ? er : el;
                    if (!(ec & 1)) continue;
                    int child = side ? r : l;
                    unsigned owner = side ? ar : al;
                    unsigned stamp = ((unsigned)k << 16) | (unsigned)node_pre[child];
                    bool any = false;
                    for (int w = 0; w < W; ++w) {
                        uint32_t g = gain[(size_t)node_dag[child] * W + w];
                        while (g) {
                            int f = __ffs(g) - 1;
                            g &= g - 1;
                            atomicMin(&tau[((size_t)(off + owner) * W + w) * 32 + f], stamp);
                            any = true;
                        }
                    }
                    if (any) atomicMin(&taumin[off + owner], stamp);
                }
            }
        }
        dec[(size_t)l * sp.S + k] = (uint16_t)el;
        dec[(size_t)r * sp.S + k] = (uint16_t)er;
        anc[(size_t)l * sp.S + k] = (uint16_t)al;
        anc[(size_t)r * sp.S + k] = (uint16_t)ar;
    }
}

// Decode + Hazard-A time stamps + cost() + selection of ONE resolved tree per CTA, all state in
// shared memory: thread k decodes from root species k (sdecode_walk_kernel), the CTA synchronises
// (the in-place synteny unions only travel between the decodes of one tree), thread k evaluates
// cost() of its decode (scost_kernel), the CTA keeps the first strict minimum in level order
// (sselect_kernel) and offers it to the batch's packed best key (sbest_kernel).  Nothing per
// (node, root) goes to global memory.
__global__ void sfused_kernel(DevSpecies sp, int Gb, const int32_t* __restrict__ left,
                              const int32_t* __restrict__ right, const int32_t* __restrict__ node_slot,
                              const int32_t* __restrict__ node_dag, const int32_t* __restrict__ node_pre,
                              const uint32_t* __restrict__ tags, int W, const uint32_t* __restrict__ gain,
                              const uint32_t* __restrict__ L, SuperCosts cs, int64_t index_base,
                              int32_t* __restrict__ min_cost, int32_t* __restrict__ root_species,
                              unsigned long long* __restrict__ best_key) {
    extern __shared__ uint32_t fused_smem[];
    const int S = sp.S;
    const int b = blockIdx.x, k = threadIdx.x;
    const int64_t off = (int64_t)b * Gb;
    uint32_t* tau = fused_smem;                       // [Gb * W * 32]
    uint32_t* taumin = tau + (size_t)Gb * W * 32;     // [Gb]
    uint32_t* taumask = taumin + Gb;                  // [Gb * W] families ever merged into a set
    // the tree itself: children (local ids, -1 for leaves), table row, preorder position, clean
    // lca-set and gain set of every node
    int32_t* sL = reinterpret_cast<int32_t*>(taumask + (size_t)Gb * W);  // [Gb] left child
    int32_t* sR = sL + Gb;                            // [Gb] right child
    int32_t* sSlot = sR + Gb;                         // [Gb]
    int32_t* sPre = sSlot + Gb;                       // [Gb]
    uint32_t* sLca = reinterpret_cast<uint32_t*>(sPre + Gb);  // [Gb * W]
    uint32_t* sGain = sLca + (size_t)Gb * W;          // [Gb * W]
    uint16_t* dec = reinterpret_cast<uint16_t*>(sGain + (size_t)Gb * W);  // [Gb * S]
    uint16_t* anc = dec + (size_t)Gb * S;             // [Gb * S]
    __shared__ unsigned long long warp_best[32];
    for (int i = threadIdx.x; i < Gb * W * 32 + Gb; i += blockDim.x) tau[i] = 0xffffffffu;
    for (int u = threadIdx.x; u < Gb; u += blockDim.x) {
        const int lg = left[off + u];
        sL[u] = lg < 0 ? -1 : lg - (int)off;
        sR[u] = lg < 0 ? -1 : right[off + u] - (int)off;
        sSlot[u] = node_slot[off + u];
        sPre[u] = node_pre[off + u];
        const size_t dag = (size_t)node_dag[off + u] * W;
        for (int w = 0; w < W; ++w) {
            sLca[u * W + w] = L[dag + w];
            sGain[u * W + w] = gain[dag + w];
        }
    }
    __syncthreads();
    if (k < S) {
        bool ok = tags[((size_t)sSlot[0] * 2 + 0) * sp.pitch + k] != SR2_NO_TAG;  // kind LCA (:487)
        dec[k] = ok ? (uint16_t)(k << 1) : (uint16_t)DEC_NONE;
        anc[k] = 0;
        const int32_t* Slot = sSlot;
        for (int u = 0; u < Gb; ++u) {
            const int l = sL[u];
            if (l < 0) continue;
            const int r = sR[u];
            unsigned e = dec[u * S + k];
            unsigned el = DEC_NONE, er = DEC_NONE;
            unsigned al = 0, ar = 0;
            if (e != DEC_NONE) {
                int x = e >> 1, kd = e & 1;
                unsigned tag = tags[((size_t)Slot[u] * 2 + kd) * sp.pitch + x];
                if (tag != SR2_NO_TAG) {
                    el = tag & 0xffffu;
                    er = tag >> 16;
                    unsigned mine = anc[u * S + k];
                    al = (el & 1) ? mine : (unsigned)l;
                    ar = (er & 1) ? mine : (unsigned)r;
#pragma unroll 1
                    for (int side = 0; side < 2; ++side) {
                        unsigned ec = side ? er : el;
                        if (!(ec & 1)) continue;
                        int child = side ? r : l;
                        unsigned owner = side ? ar : al;
                        unsigned stamp = ((unsigned)k << 16) | (unsigned)sPre[child];
                        bool any = false;
                        for (int w = 0; w < W; ++w) {
                            uint32_t g = sGain[child * W + w];
                            while (g) {
                                int f = __ffs(g) - 1;
                                g &= g - 1;
                                atomicMin(&tau[((size_t)owner * W + w) * 32 + f], stamp);
                                any = true;
                            }
                        }
                        if (any) atomicMin(&taumin[owner], stamp);
                    }
                }
            }
            dec[l * S + k] = (uint16_t)el;
            dec[r * S + k] = (uint16_t)er;
            anc[l * S + k] = (uint16_t)al;
            anc[r * S + k] = (uint16_t)ar;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Gb * W; i += blockDim.x) {
        uint32_t mask = 0;
        if (taumin[i / W] != 0xffffffffu)
            for (int f = 0; f < 32; ++f) mask |= (tau[(size_t)i * 32 + f] != 0xffffffffu ? 1u : 0u) << f;
        taumask[i] = mask;
    }
    __syncthreads();
    u64 key = SR2_KEY_NONE;
    if (k < S && dec[k] != DEC_NONE) {
        int total = 0;
        const int32_t* Pre = sPre;
        const unsigned kk = (unsigned)k << 16;
        auto word = [&](int owner, int w, unsigned stamp) {
            uint32_t bits = sLca[owner * W + w];
            if (taumin[owner] <= stamp) {
                const uint32_t* t = tau + ((size_t)owner * W + w) * 32;
                uint32_t cand = taumask[owner * W + w] & ~bits;  // only families that were merged at all
                while (cand) {
                    int f = __ffs(cand) - 1;
                    cand &= cand - 1;
                    if (t[f] <= stamp) bits |= 1u << f;
                }
            }
            return bits;
        };
        for (int u = 0; u < Gb; ++u) {
            const int l = sL[u];
            if (l < 0) continue;
            const int r = sR[u];
            const unsigned eu = dec[u * S + k], el = dec[l * S + k], er = dec[r * S + k];
            const int p = eu >> 1, a = el >> 1, c = er >> 1;
            // (tin, tout, (child0 + 1) << 16 | depth, tin of the second child) of the three species
            const uint4 gp = __ldg(sp.span4 + p), ga = __ldg(sp.span4 + a), gc = __ldg(sp.span4 + c);
            const int dp = (int)(gp.z & 0xffffu), da = (int)(ga.z & 0xffffu), dc = (int)(gc.z & 0xffffu);
            const bool pa = gp.x <= ga.x && ga.x < gp.y, pc = gp.x <= gc.x && gc.x < gp.y;
            const int ou = anc[u * S + k], ol = anc[l * S + k], orr = anc[r * S + k];
            bool in_l = true, in_r = true;
            for (int w = 0; w < W; ++w) {
                uint32_t mine = word(ou, w, kk | (unsigned)Pre[u]);
                in_l &= (mine & ~word(ol, w, kk | (unsigned)Pre[l])) == 0;
                in_r &= (mine & ~word(orr, w, kk | (unsigned)Pre[r])) == 0;
            }
            const int lc = in_l ? 0 : cs.sloss, rc = in_r ? 0 : cs.sloss;
            if (pa && pc) {
                // speciation iff a and c sit below different children of p
                const bool spec = (gp.z >> 16) != 0 && a != p && c != p && ((ga.x < gp.w) != (gc.x < gp.w));
                if (spec)
                    total += cs.spe + cs.floss * (da + dc - 2 * dp - 2) + lc + rc;
                else
                    total += cs.dup + cs.floss * (da + dc - 2 * dp) + min(lc, rc);
            } else if (pa) {
                total += cs.hgt + cs.floss * (da - dp) + lc;
            } else {
                const bool a_above_p = ga.x <= gp.x && gp.x < ga.y;
                total += cs.hgt + cs.floss * (dc - dp) + (a_above_p ? lc : rc);
            }
        }
        key = ((u64)(unsigned)total << 32) | (unsigned)k;
    }
    for (int o = 16; o; o >>= 1) key = key_min(key, __shfl_xor_sync(0xffffffffu, key, o));
    if ((threadIdx.x & 31) == 0) warp_best[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x < 32) {
        key = threadIdx.x < (blockDim.x >> 5) ? warp_best[threadIdx.x] : SR2_KEY_NONE;
        for (int o = 16; o; o >>= 1) key = key_min(key, __shfl_xor_sync(0xffffffffu, key, o));
        if (threadIdx.x == 0) {
            bool none = key == SR2_KEY_NONE;
            int cost = none ? SR2_INF : (int)(key >> 32), root = none ? -1 : (int)(key & 0xffffffffu);
            min_cost[b] = cost;
            root_species[b] = root;
            if (!none)
                atomicMin(best_key, ((u64)(unsigned)cost << 40) | ((u64)(index_base + b) << 16) | (unsigned)root);
        }
    }
}

// Compact form of sfused_kernel for species trees of at most 127 nodes and resolved trees of at
// most 254 nodes (BASELINE config #5: 59 and 65).  Same result, a third of the instructions and
// half the shared memory (ncu r3b: sfused_kernel issued 10.6 k warp-instructions per tree at 18
// warps per SM):
//   * what cost() charges for a node beyond the segmental losses is a function of the table cell the
//     node was decoded through, so the wave kernel stores it next to the tag (super_event_word) and
//     the decode adds it up: no species look-ups, no event classification per (tree, root, node);
//   * per (node, root) state is ONE 16-bit word: owner << 8 | species << 1 | kind while the node
//     waits to be decoded, then owner << 8 | class << 2 | kinds of the two children for the labelling;
//   * a child labelled INHERIT is labelled from the same, only growing, set as its parent at a later
//     time, so its segmental-loss term is 0 and only LCA-kind children are compared;
//   * the time stamps of the in-place unions (SURVEY.md Hazard A) are kept per family: thread k
//     notes into whose set it merged family f (a family is gained at one node, so at most once per
//     decode), then one pass finds the first root that merged f into each set.  No atomics, 8-bit
//     entries; tau[a][f] = (first root, preorder position of the gain node of f).
struct FusedLayout {
    int lca, gain, pmask, slot;     // 32-bit word offsets
    int ent;                        // byte offsets from here on
    int left, right, pre, fam, owner, first, inner;
    int bytes;
};
__host__ __device__ __forceinline__ FusedLayout fused_layout(int Gb, int S, int W) {
    FusedLayout f;
    const int F = 32 * W, Gb4 = (Gb + 3) & ~3;
    f.lca = 0;
    f.gain = f.lca + Gb * W;
    f.pmask = f.gain + Gb * W;
    f.slot = f.pmask + Gb * W;
    f.ent = (f.slot + Gb) * 4;
    {
        const int ent_bytes = ((Gb * S + 1) & ~1) * 2, scratch_bytes = 16 * Gb;   // unranking scratch lives here first
        f.left = f.ent + ((ent_bytes > scratch_bytes ? ent_bytes : scratch_bytes) + 3) / 4 * 4;
    }
    f.right = f.left + Gb4;
    f.pre = f.right + Gb4;
    f.fam = f.pre + Gb4;
    f.owner = f.fam + F;
    f.first = f.owner + ((F * S + 3) & ~3);
    f.inner = f.first + Gb * F;
    f.bytes = f.inner + Gb4;
    return f;
}

template <int WT>   // WT = 1: one synteny word (up to 32 families), loops over words disappear; WT = 0: W words
__global__ void __launch_bounds__(128)
sfused_compact_kernel(DevSpecies sp, DevPlan plan, int Gb, const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                      const int32_t* __restrict__ node_slot, const int32_t* __restrict__ node_dag,
                      const int32_t* __restrict__ node_pre, const uint32_t* __restrict__ tags,
                      const uint32_t* __restrict__ evs, int W_runtime, const uint32_t* __restrict__ gain,
                      const uint32_t* __restrict__ L, SuperCosts cs, int64_t index_base,
                      int32_t* __restrict__ min_cost, int32_t* __restrict__ root_species,
                      unsigned long long* __restrict__ best_key) {
    extern __shared__ uint32_t fsm[];
    const int W = WT ? WT : W_runtime;
    const int S = sp.S, F = 32 * W;
    const int b = blockIdx.x, k = threadIdx.x;
    const int64_t off = (int64_t)b * Gb;
    const FusedLayout lay = fused_layout(Gb, S, W);
    uint32_t* sLca = fsm + lay.lca;
    uint32_t* sGain = fsm + lay.gain;
    uint32_t* sPmask = fsm + lay.pmask;
    const int32_t* sSlot = reinterpret_cast<int32_t*>(fsm + lay.slot);
    uint8_t* bytes = reinterpret_cast<uint8_t*>(fsm);
    uint16_t* ent = reinterpret_cast<uint16_t*>(bytes + lay.ent);
    uint8_t* sL = bytes + lay.left;
    uint8_t* sR = bytes + lay.right;
    uint8_t* sPre = bytes + lay.pre;
    uint8_t* sFam = bytes + lay.fam;
    uint8_t* sOwner = bytes + lay.owner;
    uint8_t* sFirst = bytes + lay.first;
    uint8_t* sInt = bytes + lay.inner;
    __shared__ unsigned long long warp_best[4];
    __shared__ int s_n_int, s_any_merge;

    // ---- the tree, and "nothing merged yet" ---------------------------------------------------
    for (int i = threadIdx.x; i < (lay.inner - lay.owner) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(sOwner)[i] = 0xffffffffu;           // sOwner and sFirst
    // The tree of this resolution.  left == nullptr: nobody materialised the resolved trees (the usual case);
    // the CTA unranks its own from the plan, threads over the ORIGINAL nodes (children, table row, dag id and
    // preorder position of every node of the node's block), into scratch that the decode state will overwrite.
    int32_t* t_left = reinterpret_cast<int32_t*>(ent);          // [Gb] each (the ent region holds >= 16 Gb bytes)
    int32_t* t_right = t_left + Gb;
    int32_t* t_dag = t_right + Gb;
    int32_t* t_pre = t_dag + Gb;
    const bool unrank_here = left == nullptr;
    if (unrank_here) {
        int32_t* t_slot = reinterpret_cast<int32_t*>(fsm + lay.slot);
        for (int v = threadIdx.x; v < plan.n; v += blockDim.x)
            unrank_node(plan, plan.first + b, v, t_left, t_right, t_slot, t_dag, t_pre, 0);
        __syncthreads();
    }
    for (int u = threadIdx.x; u < Gb; u += blockDim.x) {
        const int lg = unrank_here ? t_left[u] : left[off + u];
        const int rg = unrank_here ? t_right[u] : (lg < 0 ? -1 : right[off + u]);
        const int base_id = unrank_here ? 0 : (int)off;
        const uint8_t l8 = lg < 0 ? 0xff : (uint8_t)(lg - base_id), r8 = lg < 0 ? 0xff : (uint8_t)(rg - base_id);
        const uint8_t pre8 = (uint8_t)(unrank_here ? t_pre[u] : node_pre[off + u]);
        const size_t dag = (size_t)(unrank_here ? t_dag[u] : node_dag[off + u]) * W;
        if (!unrank_here) reinterpret_cast<int32_t*>(fsm + lay.slot)[u] = node_slot[off + u];
        sL[u] = l8;
        sR[u] = r8;
        sPre[u] = pre8;
        for (int w = 0; w < W; ++w) {
            sLca[u * W + w] = L[dag + w];
            uint32_t g = gain[dag + w];
            sGain[u * W + w] = g;
            sPmask[u * W + w] = 0;
            while (g) {                                                   // a family is gained at one node
                const int f = __ffs(g) - 1;
                g &= g - 1;
                sFam[w * 32 + f] = (uint8_t)u;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // ids of the internal nodes, in id order (parents before children)
        int count = 0;
        for (int base = 0; base < Gb; base += 32) {
            const int u = base + threadIdx.x;
            const bool inner = u < Gb && sL[u] != 0xff;
            const unsigned m = __ballot_sync(0xffffffffu, inner);
            if (inner) sInt[count + __popc(m & ((1u << threadIdx.x) - 1))] = (uint8_t)u;
            count += __popc(m);
        }
        if (threadIdx.x == 0) {
            s_n_int = count;
            s_any_merge = 0;
        }
    }
    __syncthreads();
    const int n_int = s_n_int;

    // ---- decode from root species k (thread k), event terms added on the way -------------------
    // Next to the event terms the decode adds the segmental-loss terms under the assumption that NO set of
    // this tree is ever polluted (no INHERIT node with a non-empty gain set in any root's decode): then every
    // label is the clean LCA set of its owner and needs no time stamps.  That is the common case (families
    // shared by most leaves are gained at the root, which is never INHERIT: config #5 has no pollution at
    // all); the exact passes below only run for trees where some thread merged something.
    bool ok = false;
    int total = 0, clean_labels = 0;
    if (k < S) {
        ok = Gb == 1 ? false : tags[((size_t)sSlot[0] * 2 + 0) * sp.pitch + k] != SR2_NO_TAG;  // kind LCA (:487)
        if (ok) {
            ent[k] = (uint16_t)(k << 1);
            for (int j = 0; j < n_int; ++j) {
                const int u = sInt[j];
                const unsigned e = ent[u * S + k];
                const unsigned own = e >> 8;
                // (the table has fewer than 2^31 cells: checked by the launcher)
                const unsigned at = (unsigned)sSlot[u] * 2u * (unsigned)sp.pitch + (e & 1) * (unsigned)sp.pitch + ((e >> 1) & 0x7f);
                const unsigned tag = tags[at], ev = evs[at];
                const unsigned el = tag & 0xffffu, er = tag >> 16;
                const int l = sL[u], r = sR[u];
                // a tag only ever points at finite child cells; children follow their parent in id order
                SR2_DCHECK(tag != SR2_NO_TAG && (int)(el >> 1) < S && (int)(er >> 1) < S);
                SR2_DCHECK(u < Gb && l > u && r > u && l < Gb && r < Gb && own < (unsigned)Gb);
                ent[l * S + k] = (uint16_t)((((el & 1) ? own : (unsigned)l) << 8) | el);
                ent[r * S + k] = (uint16_t)((((er & 1) ? own : (unsigned)r) << 8) | er);
                ent[u * S + k] = (uint16_t)((own << 8) | ((ev & 3) << 2) | ((el & 1) << 1) | (er & 1));
                total += (int)(ev >> 2);
                {   // clean labels: only LCA-kind children are compared (see the header of this kernel)
                    const int cls = (int)(ev & 3);
                    bool need_l = !(el & 1) && cls != EV_CLASS_HGT_R, need_r = !(er & 1) && cls != EV_CLASS_HGT_L;
                    if (cls == EV_CLASS_DUP && ((el | er) & 1)) need_l = need_r = false;
                    if (need_l || need_r) {
                        bool in_l = true, in_r = true;
                        for (int w = 0; w < W; ++w) {
                            const uint32_t mine = sLca[own * W + w];
                            if (need_l) in_l &= (mine & ~sLca[l * W + w]) == 0;
                            if (need_r) in_r &= (mine & ~sLca[r * W + w]) == 0;
                        }
                        const int lc = need_l && !in_l ? cs.sloss : 0, rc = need_r && !in_r ? cs.sloss : 0;
                        clean_labels += cls == EV_CLASS_DUP ? min(lc, rc) : lc + rc;
                    }
                }
                // kind INHERIT: the child's gains go, in place, into the set of its owner (:390)
                if ((el | er) & 1) {
                    for (int w = 0; w < W; ++w) {
                        uint32_t g = ((el & 1) ? sGain[l * W + w] : 0u) | ((er & 1) ? sGain[r * W + w] : 0u);
                        if (g) s_any_merge = 1;
                        while (g) {
                            const int f = __ffs(g) - 1;
                            g &= g - 1;
                            SR2_DCHECK(sOwner[(w * 32 + f) * S + k] == 0xff);   // a family is gained at ONE node
                            sOwner[(w * 32 + f) * S + k] = (uint8_t)own;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    const bool polluted = s_any_merge != 0;   // uniform: read after the barrier
    // ---- first root that merged family f into the set of node a ---------------------------------
    for (int f = threadIdx.x; f < F && polluted; f += blockDim.x) {
        for (int kk = 0; kk < S; ++kk) {
            const int a = sOwner[f * S + kk];
            if (a != 0xff && sFirst[a * F + f] == 0xff) {
                sFirst[a * F + f] = (uint8_t)kk;
                atomicOr(&sPmask[a * W + (f >> 5)], 1u << (f & 31));
            }
        }
    }
    if (polluted) __syncthreads();
    // ---- segmental-loss terms of cost() on the labels the reference would report ----------------
    u64 key = SR2_KEY_NONE;
    if (ok && !polluted) key = ((u64)(unsigned)(total + clean_labels) << 32) | (unsigned)k;
    if (ok && polluted) {
        auto word = [&](int a, int w, int pre_n) {
            uint32_t bits = sLca[a * W + w];
            uint32_t cand = sPmask[a * W + w] & ~bits;
            while (cand) {
                const int f = __ffs(cand) - 1;
                cand &= cand - 1;
                const int k1 = sFirst[a * F + w * 32 + f];
                if (k1 < k || (k1 == k && sPre[sFam[w * 32 + f]] <= pre_n)) bits |= 1u << f;
            }
            return bits;
        };
        for (int j = 0; j < n_int; ++j) {
            const int u = sInt[j];
            const unsigned rec = ent[u * S + k];
            const int cls = (rec >> 2) & 3;
            bool need_l = !(rec & 2) && cls != EV_CLASS_HGT_R, need_r = !(rec & 1) && cls != EV_CLASS_HGT_L;
            if (cls == EV_CLASS_DUP && (rec & 3)) need_l = need_r = false;  // min(lc, rc) with a zero term
            if (!(need_l || need_r)) continue;
            const int own = rec >> 8, l = sL[u], r = sR[u];
            bool in_l = true, in_r = true;
            for (int w = 0; w < W; ++w) {
                const uint32_t mine = word(own, w, sPre[u]);
                if (need_l) in_l &= (mine & ~word(l, w, sPre[l])) == 0;
                if (need_r) in_r &= (mine & ~word(r, w, sPre[r])) == 0;
            }
            const int lc = need_l && !in_l ? cs.sloss : 0, rc = need_r && !in_r ? cs.sloss : 0;
            total += cls == EV_CLASS_DUP ? min(lc, rc) : lc + rc;
        }
        key = ((u64)(unsigned)total << 32) | (unsigned)k;
    }
    for (int o = 16; o; o >>= 1) key = key_min(key, __shfl_xor_sync(0xffffffffu, key, o));
    if ((threadIdx.x & 31) == 0) warp_best[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) key = key_min(key, warp_best[w]);
        const bool none = key == SR2_KEY_NONE;
        const int cost = none ? SR2_INF : (int)(key >> 32), root = none ? -1 : (int)(key & 0xffffffffu);
        min_cost[b] = cost;
        root_species[b] = root;
        if (!none) atomicMin(best_key, ((u64)(unsigned)cost << 40) | ((u64)(index_base + b) << 16) | (unsigned)root);
    }
}

__global__ void iota64_kernel(int64_t* __restrict__ out, int64_t n, int64_t step) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i * step;
}

__global__ void iota_kernel(int32_t* __restrict__ out, int64_t n, int32_t start) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = start + (int32_t)i;
}

// lexicographic minimum of (cost, global instance number, root species) over the batch:
// the reference's running Entry(MIN, ANY) over resolutions (_uspfs :454, :481-495)
__global__ void sbest_kernel(int n_inst, int64_t index_base, const int32_t* __restrict__ min_cost,
                             const int32_t* __restrict__ root_species, unsigned long long* __restrict__ best) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    u64 key = SR2_KEY_NONE;
    if (i < n_inst && root_species[i] >= 0)
        key = ((u64)(unsigned)min_cost[i] << 40) | ((u64)(index_base + i) << 16) | (unsigned)root_species[i];
    for (int off = 16; off; off >>= 1) key = key_min(key, __shfl_xor_sync(0xffffffffu, key, off));
    if ((threadIdx.x & 31) == 0 && key != SR2_KEY_NONE) atomicMin(best, key);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static double snow_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

extern "C" int sr2_superdtl_upload(sr2_ctx* ctx, int32_t B, const int64_t* node_off, const int32_t* left,
                                   const int32_t* right, const int32_t* leaf_species, int32_t n_words,
                                   const uint32_t* leaf_bits, const int32_t costs[5], int64_t index_base,
                                   uint32_t flags) {
    if (!ctx) return SR2_ERR_ARG;
    if (ctx->policy == SR2_POLICY_ALL) flags |= SR2_KEEP_TABLES;  // sr2_set_policy
    double t0 = snow_ms();
    if (ctx->S == 0) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_superdtl_upload: call sr2_set_species first");
    if (B < 0 || !node_off || !costs || n_words < 1 || (B > 0 && (!left || !right || !leaf_species || !leaf_bits)))
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_superdtl_upload: null array or n_words < 1");
    if (ctx->S > 32767)
        return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_superdtl_upload: more than 32767 species nodes");
    if (index_base < 0 || index_base + B >= (1 << 24))
        return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_superdtl_upload: instance numbers must stay below 2^24");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    sr2_free_super(ctx);
    const int64_t N = B ? node_off[B] : 0;
    const int S = ctx->S, W = n_words;
    SuperProblem* p = new SuperProblem();
    ctx->super = p;
    p->flags = flags;
    p->W = W;
    p->index_base = index_base;
    memcpy(p->costs, costs, sizeof(p->costs));

    size_t free_b = 0;
    {
        int rc_mem = sr2_mem_free(ctx, &free_b);
        if (rc_mem) return rc_mem;
    }
    for (int id = DP_S_LEFT; id <= DP_S_BEST; ++id) free_b += sr2_dev_pooled(ctx, id);
    free_b += sr2_dev_pooled(ctx, DP_ROWS_SLAB);
    size_t per_row = (size_t)ctx->pitch * 16 + 1;
    size_t fixed = (size_t)N * (4 * 9 + (size_t)W * 4 * 6 + (size_t)W * 128 + (size_t)S * 4) +
                   (size_t)B * ((size_t)S * 4 + 32) + (64u << 20);
    if (free_b < fixed + per_row * 4)
        return sr2_fail(ctx, SR2_ERR_NOMEM, "sr2_superdtl_upload: batch does not fit in device memory");
    int64_t rows_budget = (int64_t)(((free_b - fixed) / 10 * 9) / per_row);
    int rc = sr2_build_rows(ctx, "sr2_superdtl_upload", B, node_off, left, right, leaf_species, rows_budget,
                            (flags & SR2_KEEP_TABLES) != 0, true, &p->rows);
    if (rc) return rc;
    if (p->rows.chunks.size() > 1)
        return sr2_fail(ctx, SR2_ERR_NOMEM,
                        "sr2_superdtl_upload: batch does not fit in device memory (split it into smaller batches)");
    if (p->rows.max_nodes > 65535)
        return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_superdtl_upload: more than 65535 nodes in one object tree");
    int64_t cost_bound = 0;
    rc = sr2_check_costs(ctx, "sr2_superdtl_upload", costs, p->rows.max_nodes, &cost_bound);
    if (rc) return rc;
    if (cost_bound >= SUPER_KEY_COST_LIMIT)
        return sr2_fail(ctx, SR2_ERR_RANGE,
                        "sr2_superdtl_upload: costs x tree size may overflow the 23-bit cost field of the packed "
                        "(cost, instance, root) key");

    const int64_t R = p->rows.R;
    // global-id child arrays for the per-instance kernels
    std::vector<int32_t> gl(std::max<int64_t>(N, 1)), grt(std::max<int64_t>(N, 1));
#pragma omp parallel for num_threads(ctx->host_threads) schedule(static) if (N >= 65536)
    for (int b = 0; b < B; ++b)
        for (int64_t u = node_off[b]; u < node_off[b + 1]; ++u) {
            gl[u] = left[u] < 0 ? -1 : (int32_t)(node_off[b] + left[u]);
            grt[u] = right[u] < 0 ? -1 : (int32_t)(node_off[b] + right[u]);
        }
    size_t nw = (size_t)N * W;
#define SR2_RESERVE(id, field, bytes)                                   \
    do {                                                                \
        void* ptr__ = nullptr;                                          \
        int rc__ = sr2_dev_reserve(ctx, id, (bytes), &ptr__);           \
        if (rc__) return rc__;                                          \
        p->field = (decltype(p->field))ptr__;                           \
    } while (0)
    SR2_RESERVE(DP_S_LEFT, d_left, (size_t)N * 4);
    SR2_RESERVE(DP_S_RIGHT, d_right, (size_t)N * 4);
    SR2_RESERVE(DP_S_NODEOFF, d_node_off, (size_t)(B + 1) * 8);
    SR2_RESERVE(DP_S_BITS, d_bits, nw * 4);
    SR2_RESERVE(DP_S_PRESENT, d_present, nw * 4);
    SR2_RESERVE(DP_S_OUTSIDE, d_outside, nw * 4);
    SR2_RESERVE(DP_S_GAIN, d_gain, nw * 4);
    SR2_RESERVE(DP_S_L, d_L, nw * 4);
    SR2_RESERVE(DP_S_FLAGS, d_row_flags, (size_t)R);
    SR2_RESERVE(DP_S_T, d_T, (size_t)R * 2 * ctx->pitch * 4);
    SR2_RESERVE(DP_S_TAG, d_tag, (size_t)R * 2 * ctx->pitch * 4);
    SR2_RESERVE(DP_S_DEC, d_dec, (size_t)N * S * 2);
    SR2_RESERVE(DP_S_ANC, d_anc, (size_t)N * S * 2);
    SR2_RESERVE(DP_S_TAU, d_tau, nw * 32 * 4);
    SR2_RESERVE(DP_S_TAUMIN, d_taumin, (size_t)N * 4);
    SR2_RESERVE(DP_S_COST, d_cost, (size_t)B * S * 4);
    SR2_RESERVE(DP_S_MINCOST, d_min_cost, (size_t)B * 4);
    SR2_RESERVE(DP_S_ROOTSP, d_root_species, (size_t)B * 4);
    SR2_RESERVE(DP_S_MAP, d_map, (size_t)N * 4);
    SR2_RESERVE(DP_S_KIND, d_kind, (size_t)N * 4);
    SR2_RESERVE(DP_S_SYN, d_syn, nw * 4);
    SR2_RESERVE(DP_S_BEST, d_best, 8);
#undef SR2_RESERVE
    cudaStream_t st = ctx->stream;
    SR2_CUDA(ctx, cudaMemcpyAsync(p->d_left, gl.data(), (size_t)N * 4, cudaMemcpyHostToDevice, st));
    SR2_CUDA(ctx, cudaMemcpyAsync(p->d_right, grt.data(), (size_t)N * 4, cudaMemcpyHostToDevice, st));
    SR2_CUDA(ctx, cudaMemcpyAsync(p->d_node_off, node_off, (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, st));
    SR2_CUDA(ctx, cudaMemcpyAsync(p->d_bits, leaf_bits, nw * 4, cudaMemcpyHostToDevice, st));
    SR2_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->timing.upload_ms = (float)(snow_ms() - t0);
    return SR2_OK;
}

// Can the compact fused kernel (which unranks its own tree) serve this upload?
static bool super_compact_eligible(const sr2_ctx* ctx, int Gb, int S, int W) {
    const char* unfused = getenv("SR2_SUPERDTL_UNFUSED");
    const char* wide = getenv("SR2_SUPERDTL_FUSED_WIDE");
    if ((unfused && unfused[0] == '1') || (wide && wide[0] == '1')) return false;
    return S <= 127 && Gb >= 2 && Gb <= 254 && (size_t)fused_layout(Gb, S, W).bytes <= ctx->smem_optin;
}
static bool super_compact_table_fits(const sr2_ctx* ctx, int64_t n_slots) {   // 32-bit cell offsets in the kernel
    return n_slots * 2 * (int64_t)ctx->pitch < ((int64_t)1 << 31);
}

// The resolved trees of a shared-row upload as arrays in global memory (child ids, table row, dag id,
// preorder position, instance of every node) plus the per-(node, root) buffers of the unfused decode.
// Only built when something reads them: the general fused kernel, the unfused kernels, per-node results.
static int super_materialise_trees(sr2_ctx* ctx, SuperProblem* p) {
    if (p->trees_ready) return SR2_OK;
    const int S = ctx->S, W = p->W, Gb = p->Gb;
    const int64_t B = p->rows.B, N = B * Gb;
#define SR2_RESERVE(id, field, bytes)                                   \
    do {                                                                \
        void* ptr__ = nullptr;                                          \
        int rc__ = sr2_dev_reserve(ctx, id, (bytes), &ptr__);           \
        if (rc__) return rc__;                                          \
        p->field = (decltype(p->field))ptr__;                           \
    } while (0)
    SR2_RESERVE(DP_S_LEFT, d_left, (size_t)N * 4);
    SR2_RESERVE(DP_S_RIGHT, d_right, (size_t)N * 4);
    SR2_RESERVE(DP_S_NODESLOT, d_node_slot, (size_t)N * 4);
    SR2_RESERVE(DP_S_NODEDAG, d_node_dag, (size_t)N * 4);
    SR2_RESERVE(DP_S_NODEPRE, d_node_pre, (size_t)N * 4);
    SR2_RESERVE(DP_S_NODEINST, d_node_inst, (size_t)N * 4);
    SR2_RESERVE(DP_S_DEC, d_dec, (size_t)N * S * 2);
    SR2_RESERVE(DP_S_ANC, d_anc, (size_t)N * S * 2);
    SR2_RESERVE(DP_S_TAU, d_tau, (size_t)N * W * 32 * 4);
    SR2_RESERVE(DP_S_TAUMIN, d_taumin, (size_t)N * 4);
    SR2_RESERVE(DP_S_COST, d_cost, (size_t)B * S * 4);
    SR2_RESERVE(DP_S_MAP, d_map, (size_t)N * 4);
    SR2_RESERVE(DP_S_KIND, d_kind, (size_t)N * 4);
    SR2_RESERVE(DP_S_SYN, d_syn, (size_t)N * W * 4);
#undef SR2_RESERVE
    if (B > 0)
        resolve_unrank_kernel<<<(unsigned)((B + 127) / 128), 128, 0, ctx->stream>>>(
            p->plan_store->plan, B, p->d_left, p->d_right, p->d_node_slot, p->d_node_dag, p->d_node_pre, p->d_node_inst);
    SR2_CUDA(ctx, cudaGetLastError());
    p->trees_ready = true;
    return SR2_OK;
}

// Shared-row upload (SharedPlan, sr2_internal.cuh): plan arrays to the device, resolved trees by
// resolve_unrank_kernel.  Replaces the per-resolution host unranking + bucketing of
// sr2_resolutions_upload; results are read through the same calls as a plain batch.
int sr2_superdtl_upload_shared(sr2_ctx* ctx, const SharedPlan& plan, int32_t n_words, const int32_t costs[5],
                               uint32_t flags) {
    double t0 = snow_ms();
    if (ctx->S == 0) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_resolutions_upload: call sr2_set_species first");
    if (ctx->S > 32767)
        return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_resolutions_upload: more than 32767 species nodes");
    if (plan.first + plan.count >= (1 << 24))
        return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_resolutions_upload: instance numbers must stay below 2^24");
    if (plan.Gb > 65535)
        return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_resolutions_upload: more than 65535 nodes in one object tree");
    if (flags & SR2_KEEP_TABLES)
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_resolutions_upload: SR2_KEEP_TABLES is not available for shared rows "
                                          "(set SR2_RESOLUTIONS_UNSHARED=1)");
    int64_t cost_bound = 0;
    int rc = sr2_check_costs(ctx, "sr2_resolutions_upload", costs, plan.Gb, &cost_bound);
    if (rc) return rc;
    if (cost_bound >= SUPER_KEY_COST_LIMIT)
        return sr2_fail(ctx, SR2_ERR_RANGE,
                        "sr2_resolutions_upload: costs x tree size may overflow the 23-bit cost field of the packed "
                        "(cost, instance, root) key");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    sr2_free_super(ctx);
    const int S = ctx->S, W = n_words, Gb = plan.Gb;
    const int64_t B = plan.count, N = B * Gb, NS = plan.n_slots, D = Gb + NS;
    SuperProblem* p = new SuperProblem();
    ctx->super = p;
    p->flags = flags;
    p->W = W;
    p->index_base = plan.first;
    p->shared = true;
    p->Gb = Gb;
    p->n_slots = NS;
    p->slot_wave_off = plan.wave_off;
    memcpy(p->costs, costs, sizeof(p->costs));
    p->rows.B = (int32_t)B;
    p->rows.N = N;
    p->rows.R = NS;
    p->rows.node_off.clear();   // uniform trees: instance i starts at node i * Gb (sr2_superdtl_result_one)
    p->rows.max_nodes = Gb;

    // two device slabs of plan arrays; every array is copied to its place on its own (no packing into a
    // temporary: for the 893,025-resolution range packing 25 MB took longer than the solve)
    size_t n32 = 0, n64 = 0;
    auto place32 = [&](const std::vector<int32_t>& v) {
        size_t at = n32;
        n32 += (v.size() + 3) & ~(size_t)3;
        return at;
    };
    auto place64 = [&](const std::vector<int64_t>& v) {
        size_t at = n64;
        n64 += (v.size() + 1) & ~(size_t)1;
        return at;
    };
    const size_t o_arity = place32(plan.arity), o_base = place32(plan.base), o_choff = place32(plan.child_off),
                 o_chof = place32(plan.child_of), o_arroff = place32(plan.arr_off), o_arr = place32(plan.arr),
                 o_perm = place32(plan.perm), o_scl = place32(plan.slot_cl), o_scr = place32(plan.slot_cr),
                 o_dl = place32(plan.dag_left), o_dr = place32(plan.dag_right), o_apo = place32(plan.arr_pre_off),
                 o_ap = place32(plan.arr_pre), o_par = place32(plan.parent), o_cpos = place32(plan.child_pos);
    const size_t o_ways = place64(plan.ways), o_stride = place64(plan.stride), o_cnt = place64(plan.cnt),
                 o_vf = place64(plan.var_first), o_rb = place64(plan.raw_base);
    const size_t o_noff = n64;       // node offsets: i * Gb, written by a kernel
    n64 += (size_t)B + 2;
    size_t free_b = 0;
    {
        int rc_mem = sr2_mem_free(ctx, &free_b);
        if (rc_mem) return rc_mem;
    }
    for (int id = DP_S_LEFT; id < DP_COUNT; ++id) free_b += sr2_dev_pooled(ctx, id);
    const bool compact = super_compact_eligible(ctx, Gb, S, W) && super_compact_table_fits(ctx, NS);
    size_t need = (compact ? 0 : (size_t)N * (4 * 8 + (size_t)W * 4 + (size_t)W * 128 + (size_t)S * 4)) + (size_t)D * W * 4 * 5 +
                  (size_t)NS * ((size_t)ctx->pitch * 24 + 8) + (size_t)B * ((size_t)S * 4 + 16) + n32 * 4 +
                  n64 * 8 + (64u << 20);
    if (free_b < need)
        return sr2_fail(ctx, SR2_ERR_NOMEM, "sr2_resolutions_upload: range does not fit in device memory");
    size_t dw = (size_t)D * W;
#define SR2_RESERVE(id, field, bytes)                                   \
    do {                                                                \
        void* ptr__ = nullptr;                                          \
        int rc__ = sr2_dev_reserve(ctx, id, (bytes), &ptr__);           \
        if (rc__) return rc__;                                          \
        p->field = (decltype(p->field))ptr__;                           \
    } while (0)
    int32_t* d32 = nullptr;
    int64_t* d64 = nullptr;
    {
        void* ptr = nullptr;
        if ((rc = sr2_dev_reserve(ctx, DP_S_PLAN32, (n32 + NS) * 4, &ptr))) return rc;
        d32 = (int32_t*)ptr;
        if ((rc = sr2_dev_reserve(ctx, DP_S_PLAN64, n64 * 8, &ptr))) return rc;
        d64 = (int64_t*)ptr;
    }
    SR2_RESERVE(DP_S_BITS, d_bits, dw * 4);
    SR2_RESERVE(DP_S_PRESENT, d_present, dw * 4);
    SR2_RESERVE(DP_S_OUTSIDE, d_outside, dw * 4);
    SR2_RESERVE(DP_S_GAIN, d_gain, dw * 4);
    SR2_RESERVE(DP_S_L, d_L, dw * 4);
    SR2_RESERVE(DP_S_FLAGS, d_row_flags, (size_t)std::max<int64_t>(NS, 1));
    SR2_RESERVE(DP_S_T, d_T, (size_t)std::max<int64_t>(NS, 1) * 2 * ctx->pitch * 4);
    SR2_RESERVE(DP_S_TAG, d_tag, (size_t)std::max<int64_t>(NS, 1) * 2 * ctx->pitch * 4);
    SR2_RESERVE(DP_S_EV, d_ev, (size_t)std::max<int64_t>(NS, 1) * 2 * ctx->pitch * 4);
    SR2_RESERVE(DP_S_MINCOST, d_min_cost, (size_t)B * 4);
    SR2_RESERVE(DP_S_ROOTSP, d_root_species, (size_t)B * 4);
    SR2_RESERVE(DP_S_BEST, d_best, 8);
#undef SR2_RESERVE
    cudaStream_t st = ctx->stream;
    const double t_packed = snow_ms();
    {
        const std::pair<size_t, const std::vector<int32_t>*> a32[] = {
            {o_arity, &plan.arity}, {o_base, &plan.base},   {o_choff, &plan.child_off}, {o_chof, &plan.child_of},
            {o_arroff, &plan.arr_off}, {o_arr, &plan.arr}, {o_perm, &plan.perm},       {o_scl, &plan.slot_cl},
            {o_scr, &plan.slot_cr},  {o_dl, &plan.dag_left}, {o_dr, &plan.dag_right}, {o_apo, &plan.arr_pre_off},
            {o_ap, &plan.arr_pre},   {o_par, &plan.parent},  {o_cpos, &plan.child_pos}};
        for (const auto& a : a32)
            if (!a.second->empty())
                SR2_CUDA(ctx, cudaMemcpyAsync(d32 + a.first, a.second->data(), a.second->size() * 4, cudaMemcpyHostToDevice, st));
        const std::pair<size_t, const std::vector<int64_t>*> a64[] = {
            {o_ways, &plan.ways}, {o_stride, &plan.stride}, {o_cnt, &plan.cnt}, {o_vf, &plan.var_first}, {o_rb, &plan.raw_base}};
        for (const auto& a : a64)
            if (!a.second->empty())
                SR2_CUDA(ctx, cudaMemcpyAsync(d64 + a.first, a.second->data(), a.second->size() * 8, cudaMemcpyHostToDevice, st));
        iota64_kernel<<<(unsigned)((B + 1 + 255) / 256), 256, 0, st>>>(d64 + o_noff, B + 1, (int64_t)Gb);
    }
    // leaf bits of the dag: the leaves of a resolved tree sit at the same ids in every resolution
    SR2_CUDA(ctx, cudaMemsetAsync(p->d_bits, 0, std::max<size_t>(dw * 4, 4), st));
    SR2_CUDA(ctx, cudaMemcpyAsync(p->d_bits, plan.leaf_bits.data(), (size_t)Gb * W * 4, cudaMemcpyHostToDevice, st));
    p->d_slot_cl = d32 + o_scl;
    p->d_slot_cr = d32 + o_scr;
    p->d_dag_left = d32 + o_dl;
    p->d_dag_right = d32 + o_dr;
    p->d_slot_node = d32 + n32;
    p->d_node_off = d64 + o_noff;
    if (NS > 0) iota_kernel<<<(unsigned)((NS + 255) / 256), 256, 0, st>>>(p->d_slot_node, NS, Gb);
    DevPlan P;
    P.n = plan.n;
    P.Gb = Gb;
    P.first = plan.first;
    P.arity = d32 + o_arity;
    P.base = d32 + o_base;
    P.child_off = d32 + o_choff;
    P.child_of = d32 + o_chof;
    P.arr_off = d32 + o_arroff;
    P.arr = d32 + o_arr;
    P.perm = d32 + o_perm;
    P.arr_pre_off = d32 + o_apo;
    P.arr_pre = d32 + o_ap;
    P.parent = d32 + o_par;
    P.child_pos = d32 + o_cpos;
    P.ways = d64 + o_ways;
    P.stride = d64 + o_stride;
    P.cnt = d64 + o_cnt;
    P.var_first = d64 + o_vf;
    P.raw_base = d64 + o_rb;
    p->plan_store = new DevPlanStore{P};
    p->trees_ready = false;
    if (!compact) {   // the kernels that will run read the resolved trees from global memory
        int rc_trees = super_materialise_trees(ctx, p);
        if (rc_trees) return rc_trees;
    }
    SR2_CUDA(ctx, cudaGetLastError());
    const double t_enqueued = snow_ms();
    SR2_CUDA(ctx, cudaStreamSynchronize(st));  // the plan's vectors may be rewritten once this returns
    ctx->timing.upload_ms = (float)(snow_ms() - t0);
    if (getenv("SR2_TRACE"))
        fprintf(stderr, "[sr2] shared upload: packed + reserved %.2f ms, enqueued %.2f ms, device done %.2f ms (%zu + %zu plan bytes)\n",
                t_packed - t0, t_enqueued - t0, snow_ms() - t0, n32 * 4, n64 * 8);
    return SR2_OK;
}

extern "C" int sr2_superdtl_set_allowed(sr2_ctx* ctx, const int32_t* node_species) {
    if (!ctx) return SR2_ERR_ARG;
    SuperProblem* p = ctx->super;
    if (!p) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_superdtl_set_allowed: nothing uploaded");
    if (p->shared)
        return sr2_fail(ctx, SR2_ERR_STATE, "sr2_superdtl_set_allowed: not available for resolution batches");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!node_species) {
        p->d_allowed = nullptr;
        return SR2_OK;
    }
    const size_t N = (size_t)p->rows.N;
    for (size_t u = 0; u < N; ++u)
        if (node_species[u] < -1 || node_species[u] >= ctx->S)
            return sr2_fail(ctx, SR2_ERR_ARG, "sr2_superdtl_set_allowed: species out of range");
    void* ptr = nullptr;
    int rc = sr2_dev_reserve(ctx, DP_S_ALLOWED, std::max<size_t>(N, 1) * 4, &ptr);
    if (rc) return rc;
    p->d_allowed = (int32_t*)ptr;
    SR2_CUDA(ctx, cudaMemcpyAsync(p->d_allowed, node_species, N * 4, cudaMemcpyHostToDevice, ctx->stream));
    SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    p->solved = false;
    return SR2_OK;
}

#define SUPER_WAVE_CTA(threads, grid, smem, stream, ...)                                            \
    do {                                                                                            \
        if ((threads) == 1024) superdtl_wave_kernel<1024><<<grid, 1024, smem, stream>>>(__VA_ARGS__); \
        else if ((threads) == 512) superdtl_wave_kernel<512><<<grid, 512, smem, stream>>>(__VA_ARGS__); \
        else superdtl_wave_kernel<256><<<grid, 256, smem, stream>>>(__VA_ARGS__);                    \
    } while (0)

static int super_configure(sr2_ctx* ctx, SuperProblem* p) {
    const int S = ctx->S;
    p->spad = (S + 1) & ~1;
    const size_t per_row = (size_t)12 * p->spad * sizeof(u64);
    const size_t topology = super_topology_bytes(S, ctx->D, ctx->n_internal);
    if (per_row + topology > ctx->smem_optin)
        return sr2_fail(ctx, SR2_ERR_RANGE, "superdtl: species tree too large for one CTA's shared memory");
    if (S <= 256) {
        p->cta_per_row = false;
        p->wave_warps = (int)std::max<size_t>(1, std::min<size_t>(8, (ctx->smem_optin - 1024 - topology) / per_row));
        p->wave_smem = per_row * p->wave_warps + topology;
        SR2_CUDA(ctx, cudaFuncSetAttribute(superdtl_wave_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
    } else {
        p->cta_per_row = true;
        p->wave_smem = per_row + topology;
        // one CTA per row: the staging and the cells are throughput-bound inside the CTA (a wave of a single
        // instance is a handful of CTAs on an empty machine), so large species trees get more threads
        p->cta_threads = S > 384 ? 512 : 256;   // config #4 (S = 999): 256 / 512 / 1024 threads -> 0.434 / 0.395 / 0.413 ms of waves
        if (const char* t = getenv("SR2_SUPERDTL_CTA_THREADS")) p->cta_threads = atoi(t) >= 1024 ? 1024 : (atoi(t) >= 512 ? 512 : 256);
        SR2_CUDA(ctx, cudaFuncSetAttribute(superdtl_wave_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
        SR2_CUDA(ctx, cudaFuncSetAttribute(superdtl_wave_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
        SR2_CUDA(ctx, cudaFuncSetAttribute(superdtl_wave_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
        SR2_CUDA(ctx, cudaFuncSetAttribute(superdtl_wave_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
        SR2_CUDA(ctx, cudaFuncSetAttribute(superdtl_wave_kernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
        SR2_CUDA(ctx, cudaFuncSetAttribute(superdtl_wave_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)p->wave_smem));
    }
    return SR2_OK;
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)std::max<int64_t>(1, (n + block - 1) / block); }

// Shared-row mode, per (node, root) state in global memory: decode, time stamps, cost(), selection
// and the mapping / kind / synteny of every node of every instance.  Runs in the solve when the
// fused kernel does not fit, and on demand when a caller asks for per-node results.
static int super_decode_shared(sr2_ctx* ctx, SuperProblem* p, int* n_launch) {
    {
        int rc_trees = super_materialise_trees(ctx, p);
        if (rc_trees) return rc_trees;
    }
    cudaStream_t st = ctx->stream;
    const int W = p->W, S = ctx->S, Gb = p->Gb;
    const int64_t N = p->rows.N;
    const int B = p->rows.B;
    SuperCosts cs{p->costs[0], p->costs[1], p->costs[2], p->costs[3], p->costs[4]};
    SR2_CUDA(ctx, cudaMemsetAsync(p->d_tau, 0xff, std::max<size_t>((size_t)N * W * 128, 4), st));
    SR2_CUDA(ctx, cudaMemsetAsync(p->d_taumin, 0xff, std::max<size_t>((size_t)N * 4, 4), st));
    sdecode_walk_kernel<<<grid_for((int64_t)B * S, 128), 128, 0, st>>>(
        ctx->dsp, B, Gb, p->d_left, p->d_right, p->d_node_slot, p->d_node_dag, p->d_node_pre, p->d_tag, W,
        p->d_gain, p->d_dec, p->d_anc, p->d_tau, p->d_taumin);
    SR2_CUDA(ctx, cudaMemsetAsync(p->d_cost, 0, std::max<size_t>((size_t)B * S * 4, 4), st));
    scost_kernel<<<dim3(grid_for(N, SCOST_NODES), grid_for(S, 128)), 128, 0, st>>>(
        ctx->dsp, p->d_node_off, p->d_left, p->d_right, p->d_node_pre, p->d_node_inst, N, W, p->d_L, p->d_tau,
        p->d_taumin, p->d_dec, p->d_anc, cs, p->d_cost, p->d_node_dag);
    sselect_kernel<<<grid_for(B, 8), 256, 0, st>>>(ctx->dsp, 0, B, p->d_cost, p->d_dec, p->d_node_off,
                                                   p->d_min_cost, p->d_root_species);
    smaterialise_kernel<<<grid_for(N, 256), 256, 0, st>>>(
        ctx->dsp, 0, N, p->d_node_inst, p->d_node_off, p->d_node_pre, p->d_root_species, W, p->d_L, p->d_tau,
        p->d_taumin, p->d_dec, p->d_anc, p->d_map, p->d_kind, p->d_syn, p->d_node_dag);
    SR2_CUDA(ctx, cudaGetLastError());
    if (n_launch) *n_launch += 4;
    p->materialised = true;
    return SR2_OK;
}

// Shared-row solve: synteny sets and the table over the distinct subtrees (dag), then decode,
// cost() and selection per resolved tree.
static int super_solve_shared(sr2_ctx* ctx, SuperProblem* p) {
    cudaStream_t st = ctx->stream;
    const int W = p->W, S = ctx->S, Gb = p->Gb;
    const int64_t NS = p->n_slots, D = Gb + NS;
    const int B = p->rows.B;
    const std::vector<int64_t>& wo = p->slot_wave_off;
    const int H = (int)wo.size() - 2;
    SuperCosts cs{p->costs[0], p->costs[1], p->costs[2], p->costs[3], p->costs[4]};
    int n_launch = 0, n_waves = 0;
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    size_t dw = (size_t)D * W;
    SR2_CUDA(ctx, cudaMemcpyAsync(p->d_present, p->d_bits, dw * 4, cudaMemcpyDeviceToDevice, st));
    SR2_CUDA(ctx, cudaMemcpyAsync(p->d_L, p->d_bits, dw * 4, cudaMemcpyDeviceToDevice, st));
    SR2_CUDA(ctx, cudaMemsetAsync(p->d_outside, 0, std::max<size_t>(dw * 4, 4), st));
    SR2_CUDA(ctx, cudaMemsetAsync(p->d_best, 0xff, 8, st));
    const int32_t* s_left = p->d_dag_left + Gb;    // indexed by slot
    const int32_t* s_right = p->d_dag_right + Gb;
    for (int h = 1; h <= H; ++h) {
        int64_t n = wo[h + 1] - wo[h];
        if (n <= 0) continue;
        syn_present_wave<<<grid_for(n * W, 256), 256, 0, st>>>(p->d_slot_node, s_left, s_right, wo[h], (int)n, W,
                                                               p->d_present);
        ++n_launch;
    }
    // a shared subtree has several parents, all of which give it the same "outside" set (the
    // leaves outside a subtree do not depend on how they are arranged)
    for (int h = H; h >= 1; --h) {
        int64_t n = wo[h + 1] - wo[h];
        if (n <= 0) continue;
        syn_outside_wave<<<grid_for(n * W, 256), 256, 0, st>>>(p->d_slot_node, s_left, s_right, wo[h], (int)n, W,
                                                               p->d_present, p->d_outside);
        ++n_launch;
    }
    syn_gain_kernel<<<grid_for(D * W, 256), 256, 0, st>>>(p->d_dag_left, p->d_dag_right, D, W, p->d_present,
                                                          p->d_outside, p->d_gain);
    ++n_launch;
    for (int h = 1; h <= H; ++h) {
        int64_t n = wo[h + 1] - wo[h];
        if (n <= 0) continue;
        syn_lca_wave<<<grid_for(n, 256), 256, 0, st>>>(p->d_slot_node, s_left, s_right, wo[h], (int)n, W, p->d_gain,
                                                       p->d_L, p->d_row_flags);
        ++n_launch;
    }
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    for (int h = 1; h <= H; ++h) {
        int64_t n = wo[h + 1] - wo[h];
        if (n <= 0) continue;
        if (!p->cta_per_row) {
            superdtl_wave_kernel<32><<<grid_for(n, p->wave_warps), p->wave_warps * 32, p->wave_smem, st>>>(
                ctx->dsp, p->d_slot_cl, p->d_slot_cr, p->d_row_flags, wo[h], 0, (int)n, p->d_T, p->d_tag, cs, p->spad,
                nullptr, nullptr, p->d_ev);
        } else {
            SUPER_WAVE_CTA(p->cta_threads, (unsigned)n, p->wave_smem, st, ctx->dsp, p->d_slot_cl, p->d_slot_cr,
                           p->d_row_flags, wo[h], 0, (int)n, p->d_T, p->d_tag, cs, p->spad, nullptr, nullptr, p->d_ev);
        }
        ++n_launch;
        ++n_waves;
    }
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    SR2_CUDA(ctx, cudaGetLastError());
    if (B > 0) {
        // one CTA per resolved tree with everything per (node, root) in shared memory, when it fits;
        // mappings / syntenies of every instance are then produced on demand (super_materialise)
        const size_t fused_smem = ((size_t)Gb * W * 35 + Gb * 5) * 4 + (size_t)Gb * S * 4;
        const char* unfused = getenv("SR2_SUPERDTL_UNFUSED");
        const bool compact = super_compact_eligible(ctx, Gb, S, W) &&   // (SR2_SUPERDTL_FUSED_WIDE=1: the general kernel)
                             super_compact_table_fits(ctx, NS);
        p->fused = compact || (!(unfused && unfused[0] == '1') && S <= 1024 && fused_smem <= ctx->smem_optin);
        const FusedLayout lay = fused_layout(Gb, S, W);
        if (compact) {
            auto kernel = W == 1 ? sfused_compact_kernel<1> : sfused_compact_kernel<0>;
            SR2_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.bytes));
            const bool in_memory = p->trees_ready;   // someone asked for per-node results earlier
            kernel<<<B, (S + 31) / 32 * 32, lay.bytes, st>>>(
                ctx->dsp, p->plan_store->plan, Gb, in_memory ? p->d_left : nullptr, p->d_right, p->d_node_slot,
                p->d_node_dag, p->d_node_pre, p->d_tag, p->d_ev, W, p->d_gain, p->d_L, cs, p->index_base,
                p->d_min_cost, p->d_root_species, p->d_best);
            n_launch += 1;
            p->materialised = false;
        } else if (p->fused) {
            int rc_trees = super_materialise_trees(ctx, p);
            if (rc_trees) return rc_trees;
            SR2_CUDA(ctx, cudaFuncSetAttribute(sfused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)fused_smem));
            sfused_kernel<<<B, (S + 31) / 32 * 32, fused_smem, st>>>(
                ctx->dsp, Gb, p->d_left, p->d_right, p->d_node_slot, p->d_node_dag, p->d_node_pre, p->d_tag, W,
                p->d_gain, p->d_L, cs, p->index_base, p->d_min_cost, p->d_root_species, p->d_best);
            n_launch += 1;
            p->materialised = false;
        } else {
            int rc = super_decode_shared(ctx, p, &n_launch);
            if (rc) return rc;
            sbest_kernel<<<grid_for(B, 256), 256, 0, st>>>(B, p->index_base, p->d_min_cost, p->d_root_species,
                                                           p->d_best);
            n_launch += 1;
        }
    }
    SR2_CUDA(ctx, cudaGetLastError());
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    SR2_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    cudaEventElapsedTime(&ctx->timing.waves_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->timing.solve_ms, ctx->ev[0], ctx->ev[1]);
    ctx->timing.n_waves = n_waves;
    ctx->timing.n_launches = n_launch;
    ctx->timing.n_cells = NS * (int64_t)S * 2;  // cell-kinds actually evaluated
    p->solved = true;
    return SR2_OK;
}

extern "C" int sr2_superdtl_solve(sr2_ctx* ctx) {
    if (!ctx) return SR2_ERR_ARG;
    SuperProblem* p = ctx->super;
    if (!p) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_superdtl_solve: nothing uploaded");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = super_configure(ctx, p);
    if (rc) return rc;
    if (p->shared) return super_solve_shared(ctx, p);
    cudaStream_t st = ctx->stream;
    const RowLayout& rows = p->rows;
    const RowChunk& c = rows.chunks[0];
    const int H = c.height(), W = p->W, S = ctx->S;
    const int64_t N = rows.N;
    const int B = rows.B;
    SuperCosts cs{p->costs[0], p->costs[1], p->costs[2], p->costs[3], p->costs[4]};
    int n_launch = 0, n_waves = 0;
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    size_t nw = (size_t)N * W;
    // ---- gain sets and lca sets ------------------------------------------------------
    const bool syn_single = B == 1 && H >= 1 && H <= SYN_SINGLE_MAXH && N * W < (1 << 22);
    if (!syn_single) {
        SR2_CUDA(ctx, cudaMemcpyAsync(p->d_present, p->d_bits, nw * 4, cudaMemcpyDeviceToDevice, st));
        SR2_CUDA(ctx, cudaMemcpyAsync(p->d_L, p->d_bits, nw * 4, cudaMemcpyDeviceToDevice, st));
        SR2_CUDA(ctx, cudaMemsetAsync(p->d_outside, 0, std::max<size_t>(nw * 4, 4), st));
    }
    SR2_CUDA(ctx, cudaMemsetAsync(p->d_tau, 0xff, std::max<size_t>(nw * 128, 4), st));
    SR2_CUDA(ctx, cudaMemsetAsync(p->d_taumin, 0xff, std::max<size_t>((size_t)N * 4, 4), st));
    SR2_CUDA(ctx, cudaMemsetAsync(p->d_best, 0xff, 8, st));
    if (syn_single) {
        SynWaves wo;
        wo.H = H;
        for (int h = 0; h <= H + 1; ++h) wo.off[h] = (int32_t)c.wave_off[h];
        syn_single_kernel<<<1, 1024, 0, st>>>(wo, rows.row_node, rows.row_left, rows.row_right, p->d_left, p->d_right,
                                              (int)N, W, p->d_bits, p->d_present, p->d_outside, p->d_gain, p->d_L,
                                              p->d_row_flags);
        ++n_launch;
    }
    for (int h = 1; h <= H && !syn_single; ++h) {
        int64_t n = c.wave_off[h + 1] - c.wave_off[h];
        if (n <= 0) continue;
        syn_present_wave<<<grid_for(n * W, 256), 256, 0, st>>>(rows.row_node, rows.row_left, rows.row_right,
                                                               c.wave_off[h], (int)n, W, p->d_present);
        ++n_launch;
    }
    for (int h = H; h >= 1 && !syn_single; --h) {
        int64_t n = c.wave_off[h + 1] - c.wave_off[h];
        if (n <= 0) continue;
        syn_outside_wave<<<grid_for(n * W, 256), 256, 0, st>>>(rows.row_node, rows.row_left, rows.row_right,
                                                               c.wave_off[h], (int)n, W, p->d_present, p->d_outside);
        ++n_launch;
    }
    if (N > 0 && !syn_single) {
        syn_gain_kernel<<<grid_for(N * W, 256), 256, 0, st>>>(p->d_left, p->d_right, N, W, p->d_present,
                                                              p->d_outside, p->d_gain);
        ++n_launch;
    }
    for (int h = 1; h <= H && !syn_single; ++h) {
        int64_t n = c.wave_off[h + 1] - c.wave_off[h];
        if (n <= 0) continue;
        syn_lca_wave<<<grid_for(n, 256), 256, 0, st>>>(rows.row_node, rows.row_left, rows.row_right, c.wave_off[h],
                                                       (int)n, W, p->d_gain, p->d_L, p->d_row_flags);
        ++n_launch;
    }
    // ---- the table ---------------------------------------------------------------------
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    const bool chain_waves = p->cta_per_row && H >= 2 && !getenv("SR2_SUPERDTL_NO_CHAIN");
    if (chain_waves) {
        // one CTA per row: every wave in one launch, rows chained by done flags
        const int64_t first = c.wave_off[1], n = c.wave_off[H + 1] - first;
        void* ptr = nullptr;
        int rc_sync = sr2_dev_reserve(ctx, DP_S_CHAIN, (size_t)(n + 1) * sizeof(int), &ptr);
        if (rc_sync) return rc_sync;
        int* sync = (int*)ptr;
        SR2_CUDA(ctx, cudaMemsetAsync(sync, 0, (size_t)(n + 1) * sizeof(int), st));
        if (p->cta_threads == 1024)
            superdtl_wave_kernel<1024, true><<<(unsigned)n, 1024, p->wave_smem, st>>>(
                ctx->dsp, rows.row_cl, rows.row_cr, p->d_row_flags, first, c.row_begin, (int)n, p->d_T, p->d_tag, cs,
                p->spad, rows.row_node, p->d_allowed, nullptr, sync);
        else if (p->cta_threads == 512)
            superdtl_wave_kernel<512, true><<<(unsigned)n, 512, p->wave_smem, st>>>(
                ctx->dsp, rows.row_cl, rows.row_cr, p->d_row_flags, first, c.row_begin, (int)n, p->d_T, p->d_tag, cs,
                p->spad, rows.row_node, p->d_allowed, nullptr, sync);
        else
            superdtl_wave_kernel<256, true><<<(unsigned)n, 256, p->wave_smem, st>>>(
                ctx->dsp, rows.row_cl, rows.row_cr, p->d_row_flags, first, c.row_begin, (int)n, p->d_T, p->d_tag, cs,
                p->spad, rows.row_node, p->d_allowed, nullptr, sync);
        ++n_launch;
        ++n_waves;
    }
    for (int h = 1; h <= H && !chain_waves; ++h) {
        int64_t n = c.wave_off[h + 1] - c.wave_off[h];
        if (n <= 0) continue;
        if (!p->cta_per_row) {
            superdtl_wave_kernel<32><<<grid_for(n, p->wave_warps), p->wave_warps * 32, p->wave_smem, st>>>(
                ctx->dsp, rows.row_cl, rows.row_cr, p->d_row_flags, c.wave_off[h], c.row_begin, (int)n, p->d_T,
                p->d_tag, cs, p->spad, rows.row_node, p->d_allowed, nullptr);
        } else {
            SUPER_WAVE_CTA(p->cta_threads, (unsigned)n, p->wave_smem, st, ctx->dsp, rows.row_cl, rows.row_cr,
                           p->d_row_flags, c.wave_off[h], c.row_begin, (int)n, p->d_T, p->d_tag, cs, p->spad,
                           rows.row_node, p->d_allowed, nullptr);
        }
        ++n_launch;
        ++n_waves;
    }
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    SR2_CUDA(ctx, cudaGetLastError());
    // ---- decode every root, cost(), selection --------------------------------------------
    if (B > 0) {
        sdecode_root_kernel<<<grid_for((int64_t)B * S, 256), 256, 0, st>>>(
            ctx->dsp, rows.d_root_row, rows.d_root_node, 0, B, c.row_begin, p->d_tag, p->d_dec, p->d_anc);
        ++n_launch;
        for (int h = H; h >= 1; --h) {
            int64_t n = c.wave_off[h + 1] - c.wave_off[h];
            if (n <= 0) continue;
            sdecode_wave_kernel<<<grid_for(n * S, 256), 256, 0, st>>>(
                ctx->dsp, rows.row_node, rows.row_left, rows.row_right, p->d_node_off, rows.d_node_inst,
                rows.d_node_pre, c.wave_off[h], c.row_begin, (int)n, p->d_tag, W, p->d_gain, p->d_dec, p->d_anc,
                p->d_tau, p->d_taumin);
            ++n_launch;
        }
        SR2_CUDA(ctx, cudaMemsetAsync(p->d_cost, 0, std::max<size_t>((size_t)B * S * 4, 4), st));
        scost_kernel<<<dim3(grid_for(N, SCOST_NODES), grid_for(S, 128)), 128, 0, st>>>(
            ctx->dsp, p->d_node_off, p->d_left, p->d_right, rows.d_node_pre, rows.d_node_inst, N, W, p->d_L,
            p->d_tau, p->d_taumin, p->d_dec, p->d_anc, cs, p->d_cost, nullptr);
        sselect_kernel<<<grid_for(B, 8), 256, 0, st>>>(ctx->dsp, 0, B, p->d_cost, p->d_dec, p->d_node_off,
                                                       p->d_min_cost, p->d_root_species);
        smaterialise_kernel<<<grid_for(N, 256), 256, 0, st>>>(
            ctx->dsp, 0, N, rows.d_node_inst, p->d_node_off, rows.d_node_pre, p->d_root_species, W, p->d_L,
            p->d_tau, p->d_taumin, p->d_dec, p->d_anc, p->d_map, p->d_kind, p->d_syn, nullptr);
        sbest_kernel<<<grid_for(B, 256), 256, 0, st>>>(B, p->index_base, p->d_min_cost, p->d_root_species,
                                                       p->d_best);
        n_launch += 4;
    }
    SR2_CUDA(ctx, cudaGetLastError());
    SR2_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    SR2_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    cudaEventElapsedTime(&ctx->timing.waves_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->timing.solve_ms, ctx->ev[0], ctx->ev[1]);
    ctx->timing.n_waves = n_waves;
    ctx->timing.n_launches = n_launch;
    ctx->timing.n_cells = rows.R * (int64_t)S * 2;
    p->solved = true;
    return SR2_OK;
}

extern "C" int sr2_superdtl_results(sr2_ctx* ctx, int32_t* out_min_cost, int32_t* out_root_species,
                                    int32_t* out_map_species, int32_t* out_map_kind, uint32_t* out_synteny_bits) {
    if (!ctx) return SR2_ERR_ARG;
    SuperProblem* p = ctx->super;
    if (!p || !p->solved) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_superdtl_results: solve first");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    double t0 = snow_ms();
    if (!p->materialised && (out_map_species || out_map_kind || out_synteny_bits)) {
        int rc = super_decode_shared(ctx, p, nullptr);
        if (rc) return rc;
    }
    cudaStream_t st = ctx->stream;
    size_t B = p->rows.B, N = p->rows.N;
    if (out_min_cost) SR2_CUDA(ctx, cudaMemcpyAsync(out_min_cost, p->d_min_cost, B * 4, cudaMemcpyDeviceToHost, st));
    if (out_root_species)
        SR2_CUDA(ctx, cudaMemcpyAsync(out_root_species, p->d_root_species, B * 4, cudaMemcpyDeviceToHost, st));
    if (out_map_species) SR2_CUDA(ctx, cudaMemcpyAsync(out_map_species, p->d_map, N * 4, cudaMemcpyDeviceToHost, st));
    if (out_map_kind) SR2_CUDA(ctx, cudaMemcpyAsync(out_map_kind, p->d_kind, N * 4, cudaMemcpyDeviceToHost, st));
    if (out_synteny_bits)
        SR2_CUDA(ctx, cudaMemcpyAsync(out_synteny_bits, p->d_syn, N * p->W * 4, cudaMemcpyDeviceToHost, st));
    SR2_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->timing.results_ms = (float)(snow_ms() - t0);
    return SR2_OK;
}

extern "C" int sr2_superdtl_best(sr2_ctx* ctx, int32_t* out_min_cost, int64_t* out_instance,
                                 int32_t* out_root_species, uint64_t* out_key) {
    if (!ctx) return SR2_ERR_ARG;
    SuperProblem* p = ctx->super;
    if (!p || !p->solved) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_superdtl_best: solve first");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long key = 0;
    SR2_CUDA(ctx, cudaMemcpyAsync(&key, p->d_best, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bool none = key == SR2_KEY_NONE;
    if (out_key) *out_key = key;
    if (out_min_cost) *out_min_cost = none ? SR2_INF : (int32_t)(key >> 40);
    if (out_instance) *out_instance = none ? -1 : (int64_t)((key >> 16) & 0xffffffu);
    if (out_root_species) *out_root_species = none ? -1 : (int32_t)(key & 0xffffu);
    return SR2_OK;
}

extern "C" int sr2_superdtl_result_one(sr2_ctx* ctx, int32_t instance, int32_t* out_map_species,
                                       int32_t* out_map_kind, uint32_t* out_synteny_bits) {
    if (!ctx) return SR2_ERR_ARG;
    SuperProblem* p = ctx->super;
    if (!p || !p->solved) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_superdtl_result_one: solve first");
    if (instance < 0 || instance >= p->rows.B)
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_superdtl_result_one: bad instance");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!p->materialised) {
        int rc = super_decode_shared(ctx, p, nullptr);
        if (rc) return rc;
    }
    cudaStream_t st = ctx->stream;
    size_t off = p->shared ? (size_t)instance * p->Gb : (size_t)p->rows.node_off[instance];
    size_t G = p->shared ? (size_t)p->Gb : (size_t)p->rows.node_off[instance + 1] - off;
    if (out_map_species)
        SR2_CUDA(ctx, cudaMemcpyAsync(out_map_species, p->d_map + off, G * 4, cudaMemcpyDeviceToHost, st));
    if (out_map_kind) SR2_CUDA(ctx, cudaMemcpyAsync(out_map_kind, p->d_kind + off, G * 4, cudaMemcpyDeviceToHost, st));
    if (out_synteny_bits)
        SR2_CUDA(ctx, cudaMemcpyAsync(out_synteny_bits, p->d_syn + off * p->W, G * p->W * 4, cudaMemcpyDeviceToHost, st));
    SR2_CUDA(ctx, cudaStreamSynchronize(st));
    return SR2_OK;
}

extern "C" int sr2_superdtl_run(sr2_ctx* ctx, int32_t B, const int64_t* node_off, const int32_t* left,
                                const int32_t* right, const int32_t* leaf_species, int32_t n_words,
                                const uint32_t* leaf_bits, const int32_t costs[5], uint32_t flags,
                                int32_t* out_min_cost, int32_t* out_root_species, int32_t* out_map_species,
                                int32_t* out_map_kind, uint32_t* out_synteny_bits) {
    int rc = sr2_superdtl_upload(ctx, B, node_off, left, right, leaf_species, n_words, leaf_bits, costs, 0, flags);
    if (rc) return rc;
    rc = sr2_superdtl_solve(ctx);
    if (rc) return rc;
    return sr2_superdtl_results(ctx, out_min_cost, out_root_species, out_map_species, out_map_kind,
                                out_synteny_bits);
}

extern "C" int sr2_superdtl_tables(sr2_ctx* ctx, int32_t instance, int32_t* out_value, int32_t* out_tag,
                                   uint32_t* out_gain, uint32_t* out_lca) {
    if (!ctx) return SR2_ERR_ARG;
    SuperProblem* p = ctx->super;
    if (!p || !p->solved) return sr2_fail(ctx, SR2_ERR_STATE, "sr2_superdtl_tables: solve first");
    if (!(p->flags & SR2_KEEP_TABLES))
        return sr2_fail(ctx, SR2_ERR_STATE, "sr2_superdtl_tables: upload with SR2_KEEP_TABLES");
    if (instance < 0 || instance >= p->rows.B) return sr2_fail(ctx, SR2_ERR_ARG, "sr2_superdtl_tables: bad instance");
    SR2_CUDA(ctx, cudaSetDevice(ctx->device));
    SR2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int S = ctx->S, pitch = ctx->pitch, W = p->W;
    int64_t off = p->rows.node_off[instance], G = p->rows.node_off[instance + 1] - off;
    if (out_gain) SR2_CUDA(ctx, cudaMemcpy(out_gain, p->d_gain + off * W, (size_t)G * W * 4, cudaMemcpyDeviceToHost));
    if (out_lca) SR2_CUDA(ctx, cudaMemcpy(out_lca, p->d_L + off * W, (size_t)G * W * 4, cudaMemcpyDeviceToHost));
    std::vector<int32_t> rowT(2 * pitch);
    std::vector<uint32_t> rowG(2 * pitch);
    for (int64_t u = 0; u < G && (out_value || out_tag); ++u) {
        int64_t gr = p->rows.h_node_row[off + u];
        int32_t* ov = out_value ? out_value + (size_t)u * S * 2 : nullptr;
        int32_t* ot = out_tag ? out_tag + (size_t)u * S * 8 : nullptr;
        if (gr < 0) {
            for (int x = 0; x < S; ++x)
                for (int k = 0; k < 2; ++k) {
                    if (ov) ov[x * 2 + k] = (k == 0 && x == p->rows.h_leaf[off + u]) ? 0 : SR2_INF;
                    if (ot)
                        for (int j = 0; j < 4; ++j) ot[(x * 2 + k) * 4 + j] = -1;
                }
            continue;
        }
        SR2_CUDA(ctx, cudaMemcpy(rowT.data(), p->d_T + (size_t)gr * 2 * pitch, (size_t)pitch * 8, cudaMemcpyDeviceToHost));
        SR2_CUDA(ctx, cudaMemcpy(rowG.data(), p->d_tag + (size_t)gr * 2 * pitch, (size_t)pitch * 8, cudaMemcpyDeviceToHost));
        for (int x = 0; x < S; ++x)
            for (int k = 0; k < 2; ++k) {
                if (ov) ov[x * 2 + k] = rowT[k * pitch + x];
                if (ot) {
                    uint32_t t = rowG[k * pitch + x];
                    bool none = t == SR2_NO_TAG;
                    int32_t* o = ot + (x * 2 + k) * 4;
                    o[0] = none ? -1 : (int32_t)((t & 0xffffu) >> 1);
                    o[1] = none ? -1 : (int32_t)(t & 1u);
                    o[2] = none ? -1 : (int32_t)((t >> 16) >> 1);
                    o[3] = none ? -1 : (int32_t)((t >> 16) & 1u);
                }
            }
    }
    return SR2_OK;
}
