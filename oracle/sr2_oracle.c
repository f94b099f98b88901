/* sr2_oracle.c -- CPU restatement of the superrec2 reconciliation DP.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product
 * (superrec2_b200/) never does and has no CPU path.
 *
 * Parity pinning: this file is checked (tests/test_oracle_golden.py) against
 * fixtures produced by running the UNMODIFIED reference in the build container
 * (oracle/make_golden.py, with stand-ins for its two absent third-party
 * packages): the reference's own test vectors, data/example.out.json, and
 * several hundred random instances compared cell by cell (value and tag).
 *
 * It deliberately keeps the reference's O(|G| |S|^2) structure -- every cell
 * scans the species in level order and feeds Entry-like accumulators -- so it
 * is an independent check of the O(|G| |S|) sweep formulation used on the GPU.
 *
 * Array conventions are those of include/sr2.h: species indexed by level-order
 * (BFS) rank, object nodes numbered parents-before-children with the root at 0.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define INF 0x3fffffff
#define ERR_KEYERROR (-100) /* the reference would raise KeyError (SURVEY Hazard B) */

/* ---- Entry(MIN, ANY): /root/reference/src/superrec2/utils/dynamic_programming.py:215-238 */
typedef struct {
    int value;   /* INF = infinity.inf */
    int has_tag; /* len(_infos) != 0 */
    int a, b;    /* the info tag (one or two species / assignments) */
} entry_t;

static void entry_init(entry_t* e) {
    e->value = INF;
    e->has_tag = 0;
    e->a = e->b = -1;
}

/* Entry.update with one candidate (ANY retention) */
static void entry_update(entry_t* e, int value, int a, int b) {
    if (value > INF) value = INF;
    if (e->value == value) {
        if (!e->has_tag) { /* "is_any and not self._infos" */
            e->has_tag = 1;
            e->a = a;
            e->b = b;
        }
    }
    if (e->value > value) {
        e->has_tag = 1;
        e->a = a;
        e->b = b;
        e->value = value;
    }
}

static int sat_add(int x, int y) { return (x >= INF || y >= INF) ? INF : x + y; }

/* ---- species tree helpers (LowestCommonAncestor, utils/trees.py:31-142) -------- */
typedef struct {
    int S;
    const int *parent, *c0, *c1;
    int *depth, *tin, *tout;
} species_t;

static int species_init(species_t* sp, int S, const int* parent, const int* c0, const int* c1) {
    sp->S = S;
    sp->parent = parent;
    sp->c0 = c0;
    sp->c1 = c1;
    sp->depth = (int*)malloc(sizeof(int) * S);
    sp->tin = (int*)malloc(sizeof(int) * S);
    sp->tout = (int*)malloc(sizeof(int) * S);
    int* size = (int*)malloc(sizeof(int) * S);
    if (!sp->depth || !sp->tin || !sp->tout || !size) return -1;
    sp->depth[0] = 0;
    for (int y = 1; y < S; ++y) sp->depth[y] = sp->depth[parent[y]] + 1;
    for (int y = 0; y < S; ++y) size[y] = 1;
    for (int y = S - 1; y >= 1; --y) size[parent[y]] += size[y];
    sp->tin[0] = 0;
    for (int y = 0; y < S; ++y) {
        sp->tout[y] = sp->tin[y] + size[y];
        if (c0[y] >= 0) {
            sp->tin[c0[y]] = sp->tin[y] + 1;
            sp->tin[c1[y]] = sp->tin[y] + 1 + size[c0[y]];
        }
    }
    free(size);
    return 0;
}

static void species_free(species_t* sp) {
    free(sp->depth);
    free(sp->tin);
    free(sp->tout);
}

/* is_ancestor_of(a, b): a is b or an ancestor of b (utils/trees.py:81-92) */
static int anc(const species_t* sp, int a, int b) { return sp->tin[a] <= sp->tin[b] && sp->tin[b] < sp->tout[a]; }

static int lca(const species_t* sp, int a, int b) {
    while (!anc(sp, a, b)) a = sp->parent[a];
    return a;
}

/* distance(a, b) (utils/trees.py:134-142) */
static int dist(const species_t* sp, int a, int b) {
    return sp->depth[a] + sp->depth[b] - 2 * sp->depth[lca(sp, a, b)];
}

/* ---- node_event / cost(): model/reconciliation.py:273-369 ----------------------- */
enum { EV_LEAF, EV_INVALID, EV_SPE, EV_DUP, EV_HGT };

static int node_event(const species_t* sp, int p, int a, int b) {
    if ((a != p && anc(sp, a, p)) || (b != p && anc(sp, b, p))) return EV_INVALID;
    if (anc(sp, p, a) && anc(sp, p, b)) {
        int comparable = anc(sp, a, b) || anc(sp, b, a);
        return (p == lca(sp, a, b) && !comparable) ? EV_SPE : EV_DUP;
    }
    if (anc(sp, p, a) || anc(sp, p, b)) return EV_HGT;
    return EV_INVALID;
}

/* cost of the reconciliation `map` (species per object node); INF if invalid.
 * Iterative form of _cost_rec (:319-365). */
static int rec_cost(const species_t* sp, int G, const int* left, const int* right, const int* leaf_species,
                    const int* map, const int* costs) {
    long long total = 0;
    for (int u = 0; u < G; ++u) {
        if (left[u] < 0) {
            if (map[u] != leaf_species[u]) return INF;
            continue;
        }
        int p = map[u], a = map[left[u]], b = map[right[u]];
        int ev = node_event(sp, p, a, b);
        if (ev == EV_INVALID) return INF;
        int dl = dist(sp, p, a), dr = dist(sp, p, b);
        if (ev == EV_SPE)
            total += costs[0] + (long long)costs[3] * (dl + dr - 2);
        else if (ev == EV_DUP)
            total += costs[1] + (long long)costs[3] * (dl + dr);
        else
            total += costs[2] + (long long)costs[3] * (anc(sp, p, a) ? dl : dr);
    }
    return total >= INF ? INF : (int)total;
}

/* =================================================================================
 * dtl = reconcile_thl, /root/reference/src/superrec2/compute/reconciliation.py:60-294
 * ================================================================================= */

/* EntryProxy.update (utils/dynamic_programming.py:442-454): the entry is only created
 * when at least one candidate is finite, then ALL candidates are offered in order. */
static void proxy_update(entry_t* cell, int n, const int* val, const int* a, const int* b) {
    int any = 0;
    for (int i = 0; i < n; ++i)
        if (val[i] < INF) any = 1;
    if (!any) return;
    for (int i = 0; i < n; ++i) entry_update(cell, val[i], a[i], b[i]);
}

/* out_value[u*S+x], out_tag[(u*S+x)*2+{0,1}] may be NULL.
 * strict != 0: return ERR_KEYERROR where the reference raises (a root entry that was
 * never created); strict == 0: such roots are skipped (the product's documented choice).
 * Returns 0, ERR_KEYERROR or -1 (allocation). */
int oracle_dtl(int S, const int* parent, const int* c0, const int* c1, int G, const int* left,
               const int* right, const int* leaf_species, const int* costs, int strict, int* out_value,
               int* out_tag, int* out_min_cost, int* out_root_species, int* out_map) {
    species_t sp;
    if (species_init(&sp, S, parent, c0, c1)) return -1;
    const int dup = costs[1], hgt = costs[2], loss = costs[3];
    entry_t* T = (entry_t*)malloc(sizeof(entry_t) * (size_t)G * S);
    int* map = (int*)malloc(sizeof(int) * G);
    if (!T || !map) return -1;
    for (size_t i = 0; i < (size_t)G * S; ++i) entry_init(&T[i]);

    /* _compute_thl_table (:192-225): object nodes in postorder = descending index */
    for (int u = G - 1; u >= 0; --u) {
        if (left[u] < 0) {
            entry_update(&T[(size_t)u * S + leaf_species[u]], 0, -1, -1); /* Candidate(0): no tag */
            T[(size_t)u * S + leaf_species[u]].has_tag = 0;
            continue;
        }
        const entry_t* Tl = T + (size_t)left[u] * S;
        const entry_t* Tr = T + (size_t)right[u] * S;
        for (int x = S - 1; x >= 0; --x) { /* any order: cells of one node are independent */
            entry_t* cell = &T[(size_t)u * S + x];
            int val[3], ta[3], tb[3];
            if (c0[x] >= 0) {
                /* _compute_thl_try_speciation (:60-113) */
                entry_t ltl, rtl, ltr, rtr;
                entry_init(&ltl), entry_init(&rtl), entry_init(&ltr), entry_init(&rtr);
                for (int s = 0; s < S; ++s) { /* level order restricted to a subtree */
                    if (anc(&sp, c0[x], s)) {
                        entry_update(&ltl, Tl[s].value, s, -1);
                        entry_update(&rtl, Tr[s].value, s, -1);
                    }
                }
                for (int s = 0; s < S; ++s) {
                    if (anc(&sp, c1[x], s)) {
                        entry_update(&ltr, Tl[s].value, s, -1);
                        entry_update(&rtr, Tr[s].value, s, -1);
                    }
                }
                int n = 0;
                if (ltl.has_tag && rtr.has_tag) { /* min_ltl.combine(min_rtr, spe_combinator) */
                    val[n] = sat_add(sat_add(ltl.value, rtr.value),
                                     loss * (dist(&sp, x, ltl.a) + dist(&sp, x, rtr.a) - 2));
                    ta[n] = ltl.a, tb[n] = rtr.a, ++n;
                }
                if (ltr.has_tag && rtl.has_tag) { /* min_ltr.combine(min_rtl, spe_combinator) */
                    val[n] = sat_add(sat_add(ltr.value, rtl.value),
                                     loss * (dist(&sp, x, ltr.a) + dist(&sp, x, rtl.a) - 2));
                    ta[n] = ltr.a, tb[n] = rtl.a, ++n;
                }
                proxy_update(cell, n, val, ta, tb);
            }
            /* _compute_thl_try_duplication_transfer (:116-189) */
            entry_t ltc, lts, rtc, rts;
            entry_init(&ltc), entry_init(&lts), entry_init(&rtc), entry_init(&rts);
            for (int s = 0; s < S; ++s) {
                if (anc(&sp, x, s)) {
                    entry_update(&ltc, Tl[s].value, s, -1);
                    entry_update(&rtc, Tr[s].value, s, -1);
                } else if (!anc(&sp, s, x)) {
                    entry_update(&lts, Tl[s].value, s, -1);
                    entry_update(&rts, Tr[s].value, s, -1);
                }
            }
            int n = 0;
            if (ltc.has_tag && rtc.has_tag) { /* duplication */
                val[n] = sat_add(sat_add(dup, sat_add(ltc.value, rtc.value)),
                                 loss * (dist(&sp, x, ltc.a) + dist(&sp, x, rtc.a)));
                ta[n] = ltc.a, tb[n] = rtc.a, ++n;
            }
            if (lts.has_tag && rtc.has_tag) { /* hgt_r_combinator: left child transferred */
                val[n] = sat_add(sat_add(hgt, sat_add(lts.value, rtc.value)), loss * dist(&sp, x, rtc.a));
                ta[n] = lts.a, tb[n] = rtc.a, ++n;
            }
            if (ltc.has_tag && rts.has_tag) { /* hgt_l_combinator: right child transferred */
                val[n] = sat_add(sat_add(hgt, sat_add(ltc.value, rts.value)), loss * dist(&sp, x, ltc.a));
                ta[n] = ltc.a, tb[n] = rts.a, ++n;
            }
            proxy_update(cell, n, val, ta, tb);
        }
    }

    if (out_value || out_tag)
        for (size_t i = 0; i < (size_t)G * S; ++i) {
            if (out_value) out_value[i] = T[i].value;
            if (out_tag) {
                out_tag[2 * i] = T[i].has_tag ? T[i].a : -1;
                out_tag[2 * i + 1] = T[i].has_tag ? T[i].b : -1;
            }
        }

    /* reconcile_thl (:281-294): decode from every root species in level order, keep the
     * first strict minimum of cost() (Entry(MIN, ANY) semantics, including "inf == inf
     * with no tag yet keeps the first one"). */
    int rc = 0;
    entry_t results;
    entry_init(&results);
    int* best_map = (int*)malloc(sizeof(int) * G);
    if (!best_map) return -1;
    for (int x = 0; x < S && rc == 0; ++x) {
        const entry_t* root = &T[x];
        if (!root->has_tag) {
            /* _decode_thl_table (:234-236) yields a mapping of the root alone */
            if (left[0] >= 0) { /* cost() then indexes the missing children */
                if (strict) rc = ERR_KEYERROR;
                continue;
            }
            map[0] = x;
        } else {
            map[0] = x;
            for (int u = 0; u < G; ++u) { /* preorder-compatible: parents first */
                if (left[u] < 0) continue;
                const entry_t* e = &T[(size_t)u * S + map[u]];
                map[left[u]] = e->a;
                map[right[u]] = e->b;
            }
        }
        int c = rec_cost(&sp, G, left, right, leaf_species, map, costs);
        int before_tag = results.has_tag, before_val = results.value;
        entry_update(&results, c, x, -1);
        if (results.has_tag && (!before_tag || results.value < before_val))
            memcpy(best_map, map, sizeof(int) * G);
    }
    if (rc == 0) {
        if (results.has_tag) {
            *out_min_cost = results.value;
            *out_root_species = results.a;
            if (out_map) memcpy(out_map, best_map, sizeof(int) * G);
        } else {
            *out_min_cost = INF;
            *out_root_species = -1;
            if (out_map)
                for (int u = 0; u < G; ++u) out_map[u] = -1;
        }
    }
    free(best_map);
    free(map);
    free(T);
    species_free(&sp);
    return rc;
}

/* Batch driver used by bench.py's CPU baseline: independent instances over OpenMP
 * threads (the reference itself is single-threaded; instances are independent calls). */
int oracle_dtl_batch(int S, const int* parent, const int* c0, const int* c1, int B, const int64_t* node_off,
                     const int* left, const int* right, const int* leaf_species, const int* costs,
                     int* out_min_cost, int* out_root_species, int* out_map) {
    int rc_all = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int64_t off = node_off[b];
        int G = (int)(node_off[b + 1] - off);
        int rc = oracle_dtl(S, parent, c0, c1, G, left + off, right + off, leaf_species + off, costs, 0, NULL,
                            NULL, out_min_cost + b, out_root_species + b, out_map ? out_map + off : NULL);
        if (rc) {
#pragma omp critical
            rc_all = rc;
        }
    }
    return rc_all;
}
