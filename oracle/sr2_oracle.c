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
#ifdef _OPENMP
#include <omp.h>
#endif

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
        /* any order: cells of one node are independent (threads only matter for the full-size
         * fixtures of oracle/make_fullsize_golden.py; inside oracle_dtl_batch this region is nested
         * and therefore serial) */
#pragma omp parallel for schedule(dynamic, 16) if ((long long)S * S > 4000000)
        for (int x = S - 1; x >= 0; --x) {
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

/* Threads of the batch driver below (bench.py: launchers such as torchrun export OMP_NUM_THREADS=1,
 * and the CPU arm is to use all host threads). */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
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

/* =================================================================================
 * superdtl = usreconcile_extended_uspfs,
 * /root/reference/src/superrec2/compute/unordered_super_reconciliation.py:85-542
 * One call = one binary resolution.  Syntenies are bitsets of W 32-bit words; bit i
 * is the i-th family in sort_synteny order (model/synteny.py:21-33).
 * ================================================================================= */
enum { K_LCA = 0, K_INH = 1 };
enum { M_LEFT, M_RIGHT, M_CONSERVED, M_SEGMENT, M_SEPARATE, M_COUNT }; /* MappingChoices :62-79 */

static int bits_subset(const uint32_t* a, const uint32_t* b, int W) { /* a <= b */
    for (int w = 0; w < W; ++w)
        if (a[w] & ~b[w]) return 0;
    return 1;
}

/* two-candidate update: (value, (species, kind)) */
static void upd2(entry_t* e, int v_lca, int v_inh, int s) {
    entry_update(e, v_lca, s, K_LCA);
    entry_update(e, v_inh, s, K_INH);
}

typedef struct {
    int value, has_tag;
    int sl, kl, sr, kr;
} cell_t;

static void cell_offer(cell_t* c, int value, const entry_t* l, const entry_t* r) {
    /* Entry.update (ANY) on a cell whose tag is a ChildrenAssignment */
    if (value > INF) value = INF;
    if (c->value == value) {
        if (!c->has_tag) {
            c->has_tag = 1;
            c->sl = l->a, c->kl = l->b, c->sr = r->a, c->kr = r->b;
        }
    }
    if (c->value > value) {
        c->has_tag = 1;
        c->sl = l->a, c->kl = l->b, c->sr = r->a, c->kr = r->b;
        c->value = value;
    }
}

/* out_gain/out_lca: [G*W]; out_value: [G*S*2]; out_tag: [G*S*2*4] = (sl,kl,sr,kr) or -1;
 * out_map/out_kind: [G]; out_syn: [G*W] (decoded, i.e. possibly polluted, syntenies). */
int oracle_superdtl(int S, const int* parent, const int* c0, const int* c1, int G, const int* left,
                    const int* right, const int* leaf_species, int W, const uint32_t* leaf_bits,
                    const int* costs, uint32_t* out_gain, uint32_t* out_lca, int* out_value, int* out_tag,
                    int* out_min_cost, int* out_root_species, int* out_map, int* out_kind,
                    uint32_t* out_syn) {
    species_t sp;
    if (species_init(&sp, S, parent, c0, c1)) return -1;
    const int spe = costs[0], dup = costs[1], hgt = costs[2], floss = costs[3], sloss = costs[4];

    /* object-tree parents and depths (for the object LCA of _compute_gain_sets :96) */
    int* opar = (int*)malloc(sizeof(int) * G);
    int* odep = (int*)malloc(sizeof(int) * G);
    uint32_t* gain = (uint32_t*)calloc((size_t)G * W, sizeof(uint32_t));
    uint32_t* L = (uint32_t*)calloc((size_t)G * W, sizeof(uint32_t));
    cell_t* T = (cell_t*)malloc(sizeof(cell_t) * (size_t)G * S * 2);
    if (!opar || !odep || !gain || !L || !T) return -1;
    opar[0] = -1;
    odep[0] = 0;
    for (int u = 0; u < G; ++u)
        if (left[u] >= 0) {
            opar[left[u]] = opar[right[u]] = u;
            odep[left[u]] = odep[right[u]] = odep[u] + 1;
        }
    /* _compute_gain_sets (:85-109): each family is gained at the LCA of the leaves carrying it */
    for (int f = 0; f < W * 32; ++f) {
        int node = -1;
        for (int u = 0; u < G; ++u) {
            if (left[u] >= 0 || !(leaf_bits[(size_t)u * W + f / 32] >> (f % 32) & 1)) continue;
            if (node < 0) {
                node = u;
                continue;
            }
            int a = node, b = u;
            while (a != b) {
                if (odep[a] >= odep[b])
                    a = opar[a];
                else
                    b = opar[b];
            }
            node = a;
        }
        if (node >= 0) gain[(size_t)node * W + f / 32] |= 1u << (f % 32);
    }
    /* _compute_lca_sets (:112-136), postorder = descending index */
    for (int u = G - 1; u >= 0; --u) {
        for (int w = 0; w < W; ++w) {
            if (left[u] < 0)
                L[(size_t)u * W + w] = leaf_bits[(size_t)u * W + w];
            else
                L[(size_t)u * W + w] = (L[(size_t)left[u] * W + w] | L[(size_t)right[u] * W + w]) &
                                       ~(gain[(size_t)left[u] * W + w] | gain[(size_t)right[u] * W + w]);
        }
    }
    if (out_gain) memcpy(out_gain, gain, sizeof(uint32_t) * (size_t)G * W);
    if (out_lca) memcpy(out_lca, L, sizeof(uint32_t) * (size_t)G * W);

    /* _compute_uspfs_table (:300-358) */
    for (size_t i = 0; i < (size_t)G * S * 2; ++i) {
        T[i].value = INF;
        T[i].has_tag = 0;
        T[i].sl = T[i].kl = T[i].sr = T[i].kr = -1;
    }
#define CELL(u, x, k) T[((size_t)(u) * S + (x)) * 2 + (k)]
    for (int u = G - 1; u >= 0; --u) {
        if (left[u] < 0) {
            CELL(u, leaf_species[u], K_LCA).value = 0; /* Candidate(0): value only */
            continue;
        }
#pragma omp parallel for schedule(dynamic, 16) if ((long long)S * S > 250000)
        for (int x = 0; x < S; ++x) {
            /* _compute_uspfs_entry (:147-297) */
            entry_t sub[2][2][M_COUNT];
            for (int i = 0; i < 2; ++i)
                for (int k = 0; k < 2; ++k)
                    for (int m = 0; m < M_COUNT; ++m) entry_init(&sub[i][k][m]);
            for (int i = 0; i < 2; ++i) {
                int child = i == 0 ? left[u] : right[u];
                int lca_lca_dist, lca_inh_dist;
                if (bits_subset(L + (size_t)u * W, L + (size_t)child * W, W)) {
                    lca_lca_dist = 0;
                    lca_inh_dist = INF;
                } else {
                    lca_lca_dist = sloss;
                    lca_inh_dist = 0;
                }
                for (int s = 0; s < S; ++s) { /* level order */
                    int lca_cost = CELL(child, s, K_LCA).value;
                    int inh_cost = CELL(child, s, K_INH).value;
                    if (anc(&sp, x, s)) {
                        int above = dist(&sp, x, s) * floss;
                        upd2(&sub[i][K_INH][M_CONSERVED], sat_add(sat_add(above, lca_cost), sloss),
                             sat_add(above, inh_cost), s);
                        upd2(&sub[i][K_LCA][M_CONSERVED], sat_add(sat_add(above, lca_cost), lca_lca_dist),
                             sat_add(sat_add(above, inh_cost), lca_inh_dist), s);
                        upd2(&sub[i][K_INH][M_SEGMENT], sat_add(above, lca_cost), sat_add(above, inh_cost), s);
                        upd2(&sub[i][K_LCA][M_SEGMENT], sat_add(above, lca_cost),
                             sat_add(sat_add(above, inh_cost), lca_inh_dist), s);
                        if (c0[x] >= 0) {
                            int sd = above - floss;
                            int which = anc(&sp, c0[x], s) ? M_LEFT : (anc(&sp, c1[x], s) ? M_RIGHT : -1);
                            if (which >= 0) {
                                upd2(&sub[i][K_INH][which], sat_add(sat_add(sd, lca_cost), sloss),
                                     sat_add(sd, inh_cost), s);
                                upd2(&sub[i][K_LCA][which], sat_add(sat_add(sd, lca_cost), lca_lca_dist),
                                     sat_add(sat_add(sd, inh_cost), lca_inh_dist), s);
                            }
                        }
                    } else if (!anc(&sp, s, x)) {
                        upd2(&sub[i][K_INH][M_SEPARATE], lca_cost, inh_cost, s);
                        upd2(&sub[i][K_LCA][M_SEPARATE], lca_cost, sat_add(inh_cost, lca_inh_dist), s);
                    }
                }
            }
            for (int k = 0; k < 2; ++k) { /* :289-297 */
                static const int pair[6][2] = {{M_LEFT, M_RIGHT},       {M_RIGHT, M_LEFT},
                                               {M_CONSERVED, M_SEGMENT}, {M_SEGMENT, M_CONSERVED},
                                               {M_CONSERVED, M_SEPARATE}, {M_SEPARATE, M_CONSERVED}};
                const int ev[6] = {spe, spe, dup, dup, hgt, hgt};
                int val[6], n = 0, any = 0;
                const entry_t *el[6], *er[6];
                for (int c = 0; c < 6; ++c) {
                    const entry_t* a = &sub[0][k][pair[c][0]];
                    const entry_t* b = &sub[1][k][pair[c][1]];
                    if (!a->has_tag || !b->has_tag) continue; /* combine() of an empty entry */
                    val[n] = sat_add(ev[c], sat_add(a->value, b->value));
                    el[n] = a, er[n] = b;
                    if (val[n] < INF) any = 1;
                    ++n;
                }
                if (any) /* EntryProxy.update */
                    for (int c = 0; c < n; ++c) cell_offer(&CELL(u, x, k), val[c], el[c], er[c]);
            }
        }
    }
    if (out_value || out_tag)
        for (size_t i = 0; i < (size_t)G * S * 2; ++i) {
            if (out_value) out_value[i] = T[i].value;
            if (out_tag) {
                out_tag[4 * i + 0] = T[i].has_tag ? T[i].sl : -1;
                out_tag[4 * i + 1] = T[i].has_tag ? T[i].kl : -1;
                out_tag[4 * i + 2] = T[i].has_tag ? T[i].sr : -1;
                out_tag[4 * i + 3] = T[i].has_tag ? T[i].kr : -1;
            }
        }

    /* _uspfs (:474-497) + _decode_uspfs_table (:361-446): decode every root species in level
     * order with kind LCA; L is LIVE -- INHERIT nodes union their gain set into the set of
     * their nearest LCA-kind ancestor IN PLACE (:386-391), and that persists across roots. */
    int* map = (int*)malloc(sizeof(int) * G);
    int* kind = (int*)malloc(sizeof(int) * G);
    int* ancref = (int*)malloc(sizeof(int) * G);
    int* stack = (int*)malloc(sizeof(int) * (G + 1));
    uint32_t* syn = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)G * W);
    entry_t results;
    entry_init(&results);
    int rc = 0;
    for (int x = 0; x < S; ++x) {
        const cell_t* root = &CELL(0, x, K_LCA);
        if (left[0] >= 0 ? !root->has_tag : root->value >= INF) continue; /* yields nothing */
        map[0] = x;
        kind[0] = K_LCA;
        ancref[0] = 0;
        int top = 0, ok = 1;
        stack[top++] = 0;
        while (top) {
            int n = stack[--top];
            if (kind[n] == K_LCA) {
                ancref[n] = n;
            } else {
                for (int w = 0; w < W; ++w) L[(size_t)ancref[n] * W + w] |= gain[(size_t)n * W + w];
            }
            memcpy(syn + (size_t)n * W, L + (size_t)ancref[n] * W, sizeof(uint32_t) * W);
            if (left[n] < 0) {
                if (CELL(n, map[n], kind[n]).value >= INF) ok = 0;
                continue;
            }
            const cell_t* c = &CELL(n, map[n], kind[n]);
            if (!c->has_tag) {
                ok = 0;
                continue;
            }
            map[left[n]] = c->sl, kind[left[n]] = c->kl, ancref[left[n]] = ancref[n];
            map[right[n]] = c->sr, kind[right[n]] = c->kr, ancref[right[n]] = ancref[n];
            stack[top++] = right[n];
            stack[top++] = left[n];
        }
        if (!ok) { /* cannot happen from a root that has an entry */
            rc = -2;
            break;
        }
        /* SuperReconciliationOutput.cost() (model/reconciliation.py:511-554) */
        int c = rec_cost(&sp, G, left, right, leaf_species, map, costs);
        if (c < INF) {
            long long lab = 0;
            for (int n = 0; n < G; ++n) {
                if (left[n] < 0) continue;
                int lc = bits_subset(syn + (size_t)n * W, syn + (size_t)left[n] * W, W) ? 0 : sloss;
                int rcst = bits_subset(syn + (size_t)n * W, syn + (size_t)right[n] * W, W) ? 0 : sloss;
                int ev = node_event(&sp, map[n], map[left[n]], map[right[n]]);
                if (ev == EV_SPE)
                    lab += lc + rcst;
                else if (ev == EV_DUP)
                    lab += lc < rcst ? lc : rcst;
                else
                    lab += (anc(&sp, map[n], map[left[n]]) || anc(&sp, map[left[n]], map[n])) ? lc : rcst;
            }
            c = (c + lab >= INF) ? INF : (int)(c + lab);
        }
        int before_tag = results.has_tag, before_val = results.value;
        entry_update(&results, c, x, -1);
        if (results.has_tag && (!before_tag || results.value < before_val)) {
            if (out_map) memcpy(out_map, map, sizeof(int) * G);
            if (out_kind) memcpy(out_kind, kind, sizeof(int) * G);
            if (out_syn) memcpy(out_syn, syn, sizeof(uint32_t) * (size_t)G * W);
        }
    }
#undef CELL
    if (results.has_tag) {
        *out_min_cost = results.value;
        *out_root_species = results.a;
    } else {
        *out_min_cost = INF;
        *out_root_species = -1;
    }
    free(map), free(kind), free(ancref), free(stack), free(syn);
    free(opar), free(odep), free(gain), free(L), free(T);
    species_free(&sp);
    return rc;
}

/* Batch driver for bench.py's CPU baseline of the resolution workloads and for
 * oracle/make_fullsize_golden.py: independent binary instances (concatenated, instance-local child
 * indices, as in include/sr2.h) over OpenMP threads; results only (no tables). */
int oracle_superdtl_batch(int S, const int* parent, const int* c0, const int* c1, int B, const int64_t* node_off,
                          const int* left, const int* right, const int* leaf_species, int W,
                          const uint32_t* leaf_bits, const int* costs, int* out_min_cost,
                          int* out_root_species, int* out_map, int* out_kind, uint32_t* out_syn) {
    int rc_all = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int64_t off = node_off[b];
        int G = (int)(node_off[b + 1] - off);
        int rc = oracle_superdtl(S, parent, c0, c1, G, left + off, right + off, leaf_species + off, W,
                                 leaf_bits + off * W, costs, NULL, NULL, NULL, NULL, out_min_cost + b,
                                 out_root_species + b, out_map ? out_map + off : NULL,
                                 out_kind ? out_kind + off : NULL, out_syn ? out_syn + off * W : NULL);
        if (rc) {
#pragma omp critical
            rc_all = rc;
        }
    }
    return rc_all;
}
