// sr2_resolve.cu -- polytomy resolutions by unranking.
//
// Replaces ReconciliationInput.binarize (/root/reference/src/superrec2/model/reconciliation.py:161-190)
// and utils.trees.binarize / arrange_leaves / graft (/root/reference/src/superrec2/utils/trees.py:377-446).
// The reference materialises the full list of resolutions (twice); here resolution number r
// is built directly from r, so any contiguous range of numbers can be handed to any GPU.
//
// Order (checked against the reference's enumeration in
// tests/golden/binarize.json.gz): for a node with children c1..ck the resolutions of its subtree
// are numbered with the children's choices slowest (c1 slowest) and the node's own arrangement
// fastest; arrangement j of k items is a mixed-radix number whose fastest digit (radix 2k-3) is
// the preorder position at which item 0 is grafted into the arrangement of items 1..k-1, items
// being opaque, the new joint taking (item, displaced subtree) as (left, right).
//
// Numbering of a resolved tree: the subtree of original node v owns a contiguous block of
// 2*leaves(v)-1 ids; its k-1 joints come first (block root first, then preorder of the
// arrangement), then the blocks of its children in order.  Parents precede children.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <omp.h>

#include "sr2_internal.cuh"

extern "C" int sr2_superdtl_upload(sr2_ctx* ctx, int32_t B, const int64_t* node_off, const int32_t* left,
                                   const int32_t* right, const int32_t* leaf_species, int32_t n_words,
                                   const uint32_t* leaf_bits, const int32_t costs[5], int64_t index_base,
                                   uint32_t flags);

namespace {

const int MAX_ARITY = 16;
const int64_t COUNT_CAP = (int64_t)1 << 40;

struct PolyTree {
    int n = 0;
    std::vector<int32_t> child_off, child_idx;
    std::vector<int64_t> count;    // resolutions of the subtree (capped)
    std::vector<int64_t> ways;     // (2k-3)!! of the node itself
    std::vector<int64_t> stride;   // the node's arrangement digit: (r / stride) % ways
    std::vector<int32_t> leaves;   // leaves below
    std::vector<int32_t> base;     // first id of the node's block in a resolved tree
    int64_t total = 1;
    int n_binary = 0;              // nodes of a resolved tree
    std::string error;
};

int64_t mul_cap(int64_t a, int64_t b) {
    if (a == 0 || b == 0) return 0;
    if (a > COUNT_CAP / b) return COUNT_CAP;
    return std::min(a * b, COUNT_CAP);
}

bool analyse(int n, const int32_t* child_off, const int32_t* child_idx, PolyTree* t) {
    t->n = n;
    if (n < 1 || !child_off) {
        t->error = "empty tree";
        return false;
    }
    t->child_off.assign(child_off, child_off + n + 1);
    int n_edges = child_off[n];
    if (child_off[0] != 0 || n_edges != n - 1 || (n_edges > 0 && !child_idx)) {
        t->error = "child_off must start at 0 and list n-1 children in total";
        return false;
    }
    t->child_idx.assign(child_idx, child_idx + n_edges);
    std::vector<uint8_t> seen(n, 0);
    for (int v = 0; v < n; ++v) {
        int k = child_off[v + 1] - child_off[v];
        if (k < 0 || k == 1 || k > MAX_ARITY) {
            t->error = "node " + std::to_string(v) + " has " + std::to_string(k) + " children (allowed: 0 or 2.." +
                       std::to_string(MAX_ARITY) + ")";
            return false;
        }
        for (int e = child_off[v]; e < child_off[v + 1]; ++e) {
            int c = child_idx[e];
            if (c <= v || c >= n || seen[c]++) {
                t->error = "children must be numbered after their parent and have one parent";
                return false;
            }
        }
    }
    t->count.assign(n, 1);
    t->ways.assign(n, 1);
    t->leaves.assign(n, 0);
    for (int v = n - 1; v >= 0; --v) {
        int k = child_off[v + 1] - child_off[v];
        if (k == 0) {
            t->leaves[v] = 1;
            continue;
        }
        int64_t ways = 1;
        for (int m = 3; m <= k; ++m) ways = mul_cap(ways, 2 * m - 3);
        t->ways[v] = ways;
        int64_t total = ways;
        for (int e = child_off[v]; e < child_off[v + 1]; ++e) {
            total = mul_cap(total, t->count[child_idx[e]]);
            t->leaves[v] += t->leaves[child_idx[e]];
        }
        t->count[v] = total;
    }
    t->total = t->count[0];
    t->n_binary = 2 * t->leaves[0] - 1;
    // digit strides (top-down) and block bases
    t->stride.assign(n, 1);
    t->base.assign(n, 0);
    for (int v = 0; v < n; ++v) {
        int k = child_off[v + 1] - child_off[v];
        if (k == 0) continue;
        int64_t s = mul_cap(t->stride[v], t->ways[v]);
        for (int e = child_off[v + 1] - 1; e >= child_off[v]; --e) {  // last child varies fastest
            int c = child_idx[e];
            t->stride[c] = s;
            s = mul_cap(s, t->count[c]);
        }
        int32_t next = t->base[v] + (k - 1);
        for (int e = child_off[v]; e < child_off[v + 1]; ++e) {
            int c = child_idx[e];
            t->base[c] = next;
            next += 2 * t->leaves[c] - 1;
        }
    }
    return true;
}

// Arrangement j of k items (utils/trees.py:377-446 order, see the header of this file): the k-1
// joints in preorder, joint q having children code_l[q], code_r[q] where a code >= 0 is an item
// and a code < 0 is joint -1-q'.  Joint 0 is the root of the arrangement.
void arrange(int k, int64_t j, int* code_l, int* code_r) {
    int jl[MAX_ARITY], jr[MAX_ARITY];  // children of joint t (creation order), joints encoded as -1-t
    int n_joint = 0;
    int root = k - 1;  // item k-1
    int digit[MAX_ARITY];
    for (int m = 0; m < k - 1; ++m) {
        int radix = 2 * (k - 1 - m) - 1;
        digit[m] = (int)(j % radix);
        j /= radix;
    }
    for (int m = k - 2; m >= 0; --m) {
        // find the node at preorder position digit[m] and the link that points to it
        int pos = digit[m];
        int stack[2 * MAX_ARITY], from[2 * MAX_ARITY], side[2 * MAX_ARITY];
        int top = 0;
        stack[top] = root, from[top] = -1, side[top] = 0, ++top;
        int target = root, tfrom = -1, tside = 0;
        while (top) {
            --top;
            int node = stack[top], f = from[top], s = side[top];
            if (pos == 0) {
                target = node, tfrom = f, tside = s;
                break;
            }
            --pos;
            if (node < 0) {
                int q = -1 - node;
                stack[top] = jr[q], from[top] = q, side[top] = 1, ++top;
                stack[top] = jl[q], from[top] = q, side[top] = 0, ++top;
            }
        }
        int q = n_joint++;
        jl[q] = m;        // the grafted item goes to the left
        jr[q] = target;   // the displaced subtree to the right
        if (tfrom < 0)
            root = -1 - q;
        else if (tside == 0)
            jl[tfrom] = -1 - q;
        else
            jr[tfrom] = -1 - q;
    }
    // preorder ranks of the joints
    int rank_of[MAX_ARITY];
    {
        int stack[2 * MAX_ARITY], top = 0, next = 0;
        stack[top++] = root;
        while (top) {
            int node = stack[--top];
            if (node >= 0) continue;
            int q = -1 - node;
            rank_of[q] = next++;
            stack[top++] = jr[q];
            stack[top++] = jl[q];
        }
    }
    for (int q = 0; q < n_joint; ++q) {
        code_l[rank_of[q]] = jl[q] >= 0 ? jl[q] : -1 - rank_of[-1 - jl[q]];
        code_r[rank_of[q]] = jr[q] >= 0 ? jr[q] : -1 - rank_of[-1 - jr[q]];
    }
}

// Writes resolution r into left/right (size n_binary) and, optionally, origin[id] = original
// node of a block root / leaf, -1 for the other joints.
void unrank(const PolyTree& t, int64_t r, int32_t* left, int32_t* right, int32_t* origin) {
    for (int v = 0; v < t.n; ++v) {
        int k = t.child_off[v + 1] - t.child_off[v];
        int32_t root_id = t.base[v];
        if (origin) origin[root_id] = v;
        if (k == 0) {
            left[root_id] = right[root_id] = -1;
            continue;
        }
        int64_t j = (r / t.stride[v]) % t.ways[v];
        int code_l[MAX_ARITY], code_r[MAX_ARITY];
        arrange(k, j, code_l, code_r);
        auto resolve = [&](int code) {
            return code >= 0 ? t.base[t.child_idx[t.child_off[v] + code]] : root_id + (-1 - code);
        };
        for (int q = 0; q < k - 1; ++q) {
            left[root_id + q] = resolve(code_l[q]);
            right[root_id + q] = resolve(code_r[q]);
            if (origin && q) origin[root_id + q] = -1;
        }
    }
}

// The shared-row plan of resolutions [first, first + count) (SharedPlan in sr2_internal.cuh).
bool plan_shared(const PolyTree& t, int64_t first, int64_t count, const int32_t* leaf_species, int W,
                 const uint32_t* leaf_bits, int n_threads, SharedPlan* out, std::string* error) {
    const int n = t.n;
    SharedPlan& P = *out;
    P.n = n;
    P.Gb = t.n_binary;
    P.first = first;
    P.count = count;
    P.arity.resize(n);
    P.base = t.base;
    P.ways = t.ways;
    P.stride = t.stride;
    P.cnt = t.count;
    P.child_off = t.child_off;
    P.child_of = t.child_idx;
    P.var_first.assign(n, 0);
    P.raw_base.assign(n, -1);
    P.arr_off.assign(n, 0);
    P.arr.clear();
    P.arr_pre.clear();
    P.arr_pre_off.assign(n, 0);
    P.parent.assign(n, -1);
    P.child_pos.assign(n, 0);
    for (int v = 0; v < n; ++v)
        for (int e = t.child_off[v]; e < t.child_off[v + 1]; ++e) {
            P.parent[t.child_idx[e]] = v;
            P.child_pos[t.child_idx[e]] = e - t.child_off[v];
        }
    std::vector<int64_t> n_var(n, 0);
    for (int v = 0; v < n; ++v) {
        int k = t.child_off[v + 1] - t.child_off[v];
        P.arity[v] = k;
        if (k == 0) continue;
        // a node whose subtree carries every digit sees exactly the variants of the range
        bool full = t.count[v] == t.total;
        P.var_first[v] = full ? first : 0;
        n_var[v] = full ? count : t.count[v];
        if (t.ways[v] * (int64_t)(k - 1) * 2 + (int64_t)P.arr.size() > ((int64_t)1 << 27)) {
            *error = "arrangement tables too large";
            return false;
        }
        P.arr_off[v] = (int32_t)P.arr.size();
        P.arr.resize(P.arr.size() + (size_t)t.ways[v] * (k - 1) * 2);
        P.arr_pre_off[v] = (int32_t)P.arr_pre.size();
        P.arr_pre.resize(P.arr_pre.size() + (size_t)t.ways[v] * (2 * k - 1));
        for (int64_t j = 0; j < t.ways[v]; ++j) {
            int code_l[MAX_ARITY], code_r[MAX_ARITY];
            arrange(k, j, code_l, code_r);
            for (int q = 0; q < k - 1; ++q) {
                P.arr[P.arr_off[v] + (j * (k - 1) + q) * 2 + 0] = code_l[q];
                P.arr[P.arr_off[v] + (j * (k - 1) + q) * 2 + 1] = code_r[q];
            }
            // preorder walk of the block: a joint counts 1, an item the whole (resolution-independent) size
            // of its resolved subtree
            int32_t* rel = P.arr_pre.data() + P.arr_pre_off[v] + j * (2 * k - 1);
            int stack[2 * MAX_ARITY], top = 0, pos = 0;
            stack[top++] = -1;  // joint 0
            while (top) {
                const int code = stack[--top];
                if (code >= 0) {
                    rel[(k - 1) + code] = pos;
                    pos += 2 * t.leaves[t.child_idx[t.child_off[v] + code]] - 1;
                } else {
                    const int q = -1 - code;
                    rel[q] = pos++;
                    stack[top++] = code_r[q];
                    stack[top++] = code_l[q];
                }
            }
        }
    }
    int64_t n_raw = 0;
    for (int v = n - 1; v >= 0; --v) {
        if (P.arity[v] == 0) continue;
        P.raw_base[v] = n_raw;
        n_raw += n_var[v] * (P.arity[v] - 1);
    }
    if (n_raw >= ((int64_t)1 << 31) - P.Gb) {
        *error = "too many distinct subtrees for one batch";
        return false;
    }
    P.n_slots = n_raw;
    // raw slots: children (as raw slots or leaves) and heights, children's blocks first
    std::vector<int32_t>&raw_l = P.raw_l, &raw_r = P.raw_r;  // >= 0 raw slot, < 0: -1 - leaf's original node
    std::vector<uint16_t>& height = P.height;
    raw_l.resize((size_t)n_raw);
    raw_r.resize((size_t)n_raw);
    height.resize((size_t)n_raw);
    for (int v = n - 1; v >= 0; --v) {
        const int k = P.arity[v];
        if (k == 0) continue;
        const int32_t* arr = P.arr.data() + P.arr_off[v];
        const int64_t rb = P.raw_base[v];
#pragma omp parallel for num_threads(n_threads) schedule(static) if (n_var[v] >= 4096)
        for (int64_t i = 0; i < n_var[v]; ++i) {
            const int64_t var = P.var_first[v] + i;
            const int64_t j = var % t.ways[v];
            const int64_t mine = rb + i * (k - 1);
            for (int q = k - 2; q >= 0; --q) {  // a joint's child joints come after it in preorder
                int32_t desc[2];
                int h = 0;
                for (int side = 0; side < 2; ++side) {
                    int code = arr[(j * (k - 1) + q) * 2 + side];
                    if (code < 0) {
                        desc[side] = (int32_t)(mine + (-1 - code));
                    } else {
                        int c = t.child_idx[t.child_off[v] + code];
                        if (P.arity[c] == 0) {
                            desc[side] = -1 - c;
                        } else {
                            int64_t var_c = (var / (t.stride[c] / t.stride[v])) % t.count[c];
                            desc[side] = (int32_t)(P.raw_base[c] + (var_c - P.var_first[c]) * (P.arity[c] - 1));
                        }
                    }
                    if (desc[side] >= 0) h = std::max(h, (int)height[desc[side]]);
                }
                raw_l[mine + q] = desc[0];
                raw_r[mine + q] = desc[1];
                height[mine + q] = (uint16_t)(h + 1);
            }
        }
    }
    // slots = raw slots sorted by height (counting sort, stable)
    int H = 0;
#pragma omp parallel for num_threads(n_threads) schedule(static) reduction(max : H) if (n_raw >= 65536)
    for (int64_t s = 0; s < n_raw; ++s) H = std::max(H, (int)height[s]);
    P.wave_off.assign(H + 2, 0);
    for (int64_t s = 0; s < n_raw; ++s) P.wave_off[height[s] + 1]++;
    for (int h = 0; h <= H; ++h) P.wave_off[h + 1] += P.wave_off[h];
    P.perm.resize((size_t)n_raw);
    {
        std::vector<int64_t> cursor(P.wave_off.begin(), P.wave_off.end() - 1);
        for (int64_t s = 0; s < n_raw; ++s) P.perm[s] = (int32_t)cursor[height[s]]++;
    }
    P.slot_cl.resize((size_t)n_raw);
    P.slot_cr.resize((size_t)n_raw);
    P.dag_left.assign((size_t)(P.Gb + n_raw), -1);
    P.dag_right.assign((size_t)(P.Gb + n_raw), -1);
    P.leaf_species.assign(P.Gb, -1);
    P.leaf_bits.assign((size_t)P.Gb * W, 0);
    for (int v = 0; v < n; ++v)
        if (P.arity[v] == 0) {
            P.leaf_species[t.base[v]] = leaf_species[v];
            memcpy(&P.leaf_bits[(size_t)t.base[v] * W], leaf_bits + (size_t)v * W, sizeof(uint32_t) * W);
        }
#pragma omp parallel for num_threads(n_threads) schedule(static) if (n_raw >= 16384)
    for (int64_t s = 0; s < n_raw; ++s) {
        const int32_t slot = P.perm[s];
        for (int side = 0; side < 2; ++side) {
            int32_t d = side ? raw_r[s] : raw_l[s];
            int32_t table, dag;
            if (d >= 0) {
                table = P.perm[d];
                dag = P.Gb + P.perm[d];
            } else {
                int c = -1 - d;
                table = -1 - leaf_species[c];
                dag = t.base[c];
            }
            (side ? P.slot_cr : P.slot_cl)[slot] = table;
            (side ? P.dag_right : P.dag_left)[P.Gb + slot] = dag;
        }
    }
    return true;
}

}  // namespace

extern "C" int sr2_resolutions_count(int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx,
                                     int64_t* out_count, int32_t* out_binary_nodes) {
    PolyTree t;
    if (!analyse(n_nodes, child_off, child_idx, &t)) return SR2_ERR_ARG;
    if (out_count) *out_count = t.total;
    if (out_binary_nodes) *out_binary_nodes = t.n_binary;
    return SR2_OK;
}

extern "C" int sr2_resolutions_tree(int32_t n_nodes, const int32_t* child_off, const int32_t* child_idx,
                                    int64_t index, int32_t* out_left, int32_t* out_right, int32_t* out_origin) {
    PolyTree t;
    if (!analyse(n_nodes, child_off, child_idx, &t)) return SR2_ERR_ARG;
    if (index < 0 || index >= t.total || !out_left || !out_right) return SR2_ERR_ARG;
    unrank(t, index, out_left, out_right, out_origin);
    return SR2_OK;
}

extern "C" int sr2_resolutions_upload(sr2_ctx* ctx, int32_t n_nodes, const int32_t* child_off,
                                      const int32_t* child_idx, const int32_t* leaf_species, int32_t n_words,
                                      const uint32_t* leaf_bits, const int32_t costs[5], int64_t first,
                                      int64_t count, uint32_t flags) {
    if (!ctx) return SR2_ERR_ARG;
    PolyTree t;
    if (!analyse(n_nodes, child_off, child_idx, &t))
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_resolutions_upload: " + t.error);
    if (!leaf_species || !leaf_bits || n_words < 1)
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_resolutions_upload: null array or n_words < 1");
    if (t.total >= (1 << 24))
        return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_resolutions_upload: 2^24 or more resolutions");
    if (first < 0 || count < 0 || first + count > t.total)
        return sr2_fail(ctx, SR2_ERR_ARG, "sr2_resolutions_upload: resolution range out of bounds");
    const int Gb = t.n_binary, W = n_words;
    if ((int64_t)Gb * count >= ((int64_t)1 << 31))
        return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_resolutions_upload: range too large for one batch");
    for (int v = 0; v < n_nodes; ++v)
        if (child_off[v + 1] == child_off[v] && (leaf_species[v] < 0 || leaf_species[v] >= ctx->S))
            return sr2_fail(ctx, SR2_ERR_ARG, "sr2_resolutions_upload: leaf species out of range");
    const char* unshared = getenv("SR2_RESOLUTIONS_UNSHARED");
    if (!(unshared && unshared[0] == '1') && count > 0 && Gb > 1) {
        // one table row per distinct subtree, resolved trees built on the device
        // one plan object per calling thread, reused: its vectors keep their capacity between calls
        static thread_local SharedPlan plan;
        std::string error;
        const auto t0 = std::chrono::steady_clock::now();
        if (!plan_shared(t, first, count, leaf_species, W, leaf_bits, ctx->host_threads, &plan, &error))
            return sr2_fail(ctx, SR2_ERR_RANGE, "sr2_resolutions_upload: " + error);
        if (getenv("SR2_TRACE"))
            fprintf(stderr, "[sr2] resolutions upload: host plan of %lld slots took %.2f ms\n", (long long)plan.n_slots,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        return sr2_superdtl_upload_shared(ctx, plan, W, costs, flags);
    }
    std::vector<int64_t> node_off(count + 1);
    for (int64_t i = 0; i <= count; ++i) node_off[i] = i * Gb;
    size_t N = (size_t)Gb * count;
    std::vector<int32_t> left(std::max<size_t>(N, 1)), right(std::max<size_t>(N, 1)),
        leaf(std::max<size_t>(N, 1), -1);
    std::vector<uint32_t> bits(std::max<size_t>(N * W, 1), 0);
    // leaves sit at the same ids in every resolution
    std::vector<int32_t> tmpl_leaf(Gb, -1);
    std::vector<uint32_t> tmpl_bits((size_t)Gb * W, 0);
    for (int v = 0; v < n_nodes; ++v)
        if (child_off[v + 1] == child_off[v]) {
            tmpl_leaf[t.base[v]] = leaf_species[v];
            memcpy(&tmpl_bits[(size_t)t.base[v] * W], leaf_bits + (size_t)v * W, sizeof(uint32_t) * W);
        }
#pragma omp parallel for num_threads(ctx->host_threads) schedule(static) if (count * Gb >= 65536)
    for (int64_t i = 0; i < count; ++i) {
        unrank(t, first + i, left.data() + i * Gb, right.data() + i * Gb, nullptr);
        memcpy(leaf.data() + i * Gb, tmpl_leaf.data(), sizeof(int32_t) * Gb);
        memcpy(bits.data() + (size_t)i * Gb * W, tmpl_bits.data(), sizeof(uint32_t) * (size_t)Gb * W);
    }
    return sr2_superdtl_upload(ctx, (int32_t)count, node_off.data(), left.data(), right.data(), leaf.data(), W,
                               bits.data(), costs, first, flags);
}
