/*
 * wide.c -- binary flattened tree (flatten.c) -> 4-wide, 8-bit quantised, 64-byte nodes.
 *
 * Why: an incoherent ray fetches one node per step from a different cache line than its warp
 * neighbours, so the ray-buffer kernel is bound by dependent L2 round trips and by L1 wavefronts
 * per step (ncu, profiles/).  Folding two binary levels into one 64-byte node halves the steps at
 * the same bytes per step.  This is legal because interior boxes only cull (the
 * reference result = best over LEAVES whose own exact box + sphere tests pass), so interior
 * boxes may be replaced by any boxes that CONTAIN them.  Leaves keep their exact box:
 *   regular leaf  (box bit-identical to {c - r, c + r}, bvh.c:26-35; every sphere_count==1 leaf):
 *                 ref = ~sphere_index, the kernel recomputes the box from the float4 sphere;
 *   special leaf  (sphere_count > 1 (Q2), sphere_count == 0 (Q3, carries its parent's box), or a
 *                 foreign tree whose leaf box is not the sphere box): ref = ~(SPECIAL | k) into a
 *                 table of 32-byte records {lo.xyz, hi.x, hi.y, hi.z, sphere index, 0};
 *   leaves that can never report a hit (sphere index == n: the reference's one-past-the-end
 *                 read, defined here as a miss; absent sibling of a synthetic root) are dropped.
 *
 * Node layout (16 x uint32, 64-byte aligned):
 *   w0..w2   p.xyz  (float)  origin of the node's quantisation grid; NaN = axis not quantisable,
 *                            the axis then accepts everything (NaN drops out of FMNMX)
 *   w3       s.x  (float)  grid cell size of the x axis, an exact power of two (s.y, s.z: w14, w15)
 *   w4..w6   qlo_x, qlo_y, qlo_z   one byte per child (child c in bits 8c..8c+7)
 *   w7..w9   qhi_x, qhi_y, qhi_z
 *   w10..w13 ref[4]  >= 0 wide node index; < 0 leaf (see above); SAHBVH_REF_NONE = no child
 *   w14,w15  s.y, s.z
 * Decoded plane of byte q = p + (1024 + q)*s (the kernel widens the bytes through fp16 1024 + q,
 * see sah_bytes1024), and, in exact arithmetic,
 *   p + (1024 + qlo)*s <= lo - s/512   and   p + (1024 + qhi)*s >= hi + s/512
 * (checked here per child and axis in long double; an axis that fails the check is marked
 * not quantisable).  The s/512 margin absorbs the rounding of the kernel's two FMAs per plane
 * (<= 2^-15 s, see traverse_core.cuh), so no run-time widening is needed.
 *
 * Every node also gets a FULL-PRECISION twin (w->full, 128 bytes = four 32-byte child records in the binary
 * tree's own format, boxes and leaf refs unchanged) with the same numbering: the packet kernel tests one child
 * box per lane, so it gains the halved depth without paying for a decode.
 *
 * Nodes are numbered breadth first (the upper levels and the up to four children of a node are adjacent).
 * Children are chosen by the usual surface-area greedy collapse: start from the two children of
 * a binary node and repeatedly open the inner child with the largest area until there are 4.
 */
#include "sahbvh_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float lo[3], hi[3]; int32_t ref; } rec_t;

typedef struct {
    const flat_bvh *f;
    wide_bvh *w;
    int64_t cap_nodes, cap_special;
    int rc;
    int max_depth;
} ctx_t;

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static void load_rec(const flat_bvh *f, int32_t node, int child, rec_t *r)
{
    const float *q = f->nodes + (size_t)node * 16 + child * 8;
    r->lo[0] = q[0]; r->lo[1] = q[1]; r->lo[2] = q[2];
    r->hi[0] = q[3]; r->hi[1] = q[4]; r->hi[2] = q[5];
    r->ref = (int32_t)f2u(q[6]);
}

static float rec_area(const rec_t *r)
{
    float dx = r->hi[0] - r->lo[0], dy = r->hi[1] - r->lo[1], dz = r->hi[2] - r->lo[2];
    float a = dx * dy + dy * dz + dz * dx;
    return a == a ? a : INFINITY; /* NaN boxes are opened first */
}

/* leaf record -> wide leaf ref (or SAHBVH_REF_NONE when it can never hit) */
static int32_t leaf_ref(ctx_t *c, const rec_t *r)
{
    const flat_bvh *f = c->f;
    int32_t idx = ~r->ref;
    if (r->ref == SAHBVH_REF_NONE || (uint32_t)idx >= (uint32_t)f->info.n_spheres) return SAHBVH_REF_NONE;
    const float *s = f->spheres4 + (size_t)idx * 4;
    float lo[3] = {s[0] - s[3], s[1] - s[3], s[2] - s[3]}, hi[3] = {s[0] + s[3], s[1] + s[3], s[2] + s[3]};
    if (memcmp(lo, r->lo, 12) == 0 && memcmp(hi, r->hi, 12) == 0) return ~idx;
    wide_bvh *w = c->w;
    if (w->n_special == c->cap_special) {
        int64_t ncap = c->cap_special ? c->cap_special * 2 : 1024;
        float *ns = (float *)realloc(w->special, (size_t)ncap * 32);
        if (!ns) { c->rc = SAHBVH_E_NOMEM; return SAHBVH_REF_NONE; }
        w->special = ns;
        c->cap_special = ncap;
    }
    float *q = w->special + (size_t)w->n_special * 8;
    q[0] = r->lo[0]; q[1] = r->lo[1]; q[2] = r->lo[2]; q[3] = r->hi[0];
    q[4] = r->hi[1]; q[5] = r->hi[2]; q[6] = u2f((uint32_t)idx); q[7] = 0.0f;
    return ~(int32_t)(SAHBVH_WIDE_SPECIAL | (uint32_t)w->n_special++);
}

/* quantise one axis of up to 4 child boxes; returns 0 and fills p/e/qlo/qhi, or -1 (axis ignored) */
static int quantise_axis(const rec_t *ch, int n, int a, float *p_out, int *e_out, uint8_t qlo[4], uint8_t qhi[4])
{
    double mn = INFINITY, mx = -INFINITY, maxabs = 0.0;
    for (int k = 0; k < n; k++) {
        double lo = ch[k].lo[a], hi = ch[k].hi[a];
        if (!isfinite(lo) || !isfinite(hi) || lo > hi) return -1;
        if (lo < mn) mn = lo;
        if (hi > mx) mx = hi;
        if (fabs(lo) > maxabs) maxabs = fabs(lo);
        if (fabs(hi) > maxabs) maxabs = fabs(hi);
    }
    if (n == 0) return -1;
    double ext = mx - mn;
    /* smallest grid with 253 cells >= extent, but never finer than 8 ulp of the coordinates
     * (p is a float: its own rounding must stay well inside one cell) */
    int e = -60;
    if (ext > 0.0) {
        int ee;
        frexp(ext / 253.0, &ee); /* ext/253 = m * 2^ee, m in [0.5,1) -> 2^ee >= ext/253 */
        if (ee > e) e = ee;
    }
    if (maxabs > 0.0) {
        int ea;
        frexp(maxabs, &ea);      /* maxabs < 2^ea, ulp(float) <= 2^(ea-24) */
        if (ea - 24 + 3 > e) e = ea - 24 + 3;
    }
    if (e > 27) return -1;
    double s = ldexp(1.0, e);
    float p = (float)(mn - 1024.5 * s);
    while ((double)p > mn - 1024.25 * s) p = nextafterf(p, -INFINITY);
    if (!isfinite(p)) return -1;
    for (int k = 0; k < n; k++) {
        double l = floor(((double)ch[k].lo[a] - (double)p) / s - 1024.0 - 1.0 / 256.0);
        double h = ceil(((double)ch[k].hi[a] - (double)p) / s - 1024.0 + 1.0 / 256.0);
        if (l < 0.0 || h > 255.0) return -1;
        /* containment with margin s/512, in long double */
        long double dl = (long double)p + (1024.0L + (long double)l) * (long double)s;
        long double dh = (long double)p + (1024.0L + (long double)h) * (long double)s;
        long double m = (long double)s / 512.0L;
        if (!(dl <= (long double)ch[k].lo[a] - m) || !(dh >= (long double)ch[k].hi[a] + m)) return -1;
        qlo[k] = (uint8_t)l;
        qhi[k] = (uint8_t)h;
    }
    *p_out = p;
    *e_out = e;
    return 0;
}

/* make room for node id `need - 1` (both node arrays grow together) */
static int reserve_nodes(ctx_t *c, int64_t need)
{
    wide_bvh *w = c->w;
    if (need <= c->cap_nodes) return SAHBVH_OK;
    int64_t ncap = c->cap_nodes * 2 > need ? c->cap_nodes * 2 : need;
    uint32_t *nn = (uint32_t *)aligned_alloc(64, (size_t)ncap * 64);
    float *nf = (float *)aligned_alloc(128, (size_t)ncap * 128);
    if (!nn || !nf) { free(nn); free(nf); return SAHBVH_E_NOMEM; }
    const int64_t keep = w->n_wide < c->cap_nodes ? w->n_wide : c->cap_nodes; /* ids beyond the old capacity are not written yet */
    memcpy(nn, w->nodes, (size_t)keep * 64);
    memcpy(nf, w->full, (size_t)keep * 128);
    free(w->nodes);
    free(w->full);
    w->nodes = nn;
    w->full = nf;
    c->cap_nodes = ncap;
    return SAHBVH_OK;
}

/* Node `me` (already numbered) from binary node b2.  Inner children are numbered n_wide, n_wide+1, ... in slot
 * order and their binary nodes are written to child_b2[] for the caller's queue: the tree comes out in
 * BREADTH-FIRST order (the device collapse in gpu_builder.cu numbers level by level and must produce the same
 * bytes; it also puts the upper levels and the four children of a node next to each other). */
static int emit(ctx_t *c, int32_t b2, int32_t me, int32_t child_b2[4], int *n_child)
{
    rec_t ch[4];
    int n = 2;
    load_rec(c->f, b2, 0, &ch[0]);
    load_rec(c->f, b2, 1, &ch[1]);
    while (n < 4) {
        int best = -1;
        float best_area = -1.0f;
        for (int k = 0; k < n; k++)
            if (ch[k].ref >= 0) {
                float a = rec_area(&ch[k]);
                if (a > best_area) { best_area = a; best = k; }
            }
        if (best < 0) break;
        int32_t open = ch[best].ref;
        for (int k = n; k > best + 1; k--) ch[k] = ch[k - 1]; /* keep left-to-right order */
        load_rec(c->f, open, 0, &ch[best]);
        load_rec(c->f, open, 1, &ch[best + 1]);
        n++;
    }
    /* full-precision twin of this node for the packet kernel: the same (up to) four children as 32-byte
     * records {lo.xyz, hi.x}{hi.y, hi.z, ref, 0} exactly as in the binary tree (leaf refs and boxes unchanged) */
    float full[32];
    for (int k = 0; k < 4; k++) {
        float *q = full + 8 * k;
        if (k < n) {
            q[0] = ch[k].lo[0]; q[1] = ch[k].lo[1]; q[2] = ch[k].lo[2]; q[3] = ch[k].hi[0];
            q[4] = ch[k].hi[1]; q[5] = ch[k].hi[2]; q[6] = u2f((uint32_t)ch[k].ref); q[7] = 0.0f;
        } else {
            q[0] = q[1] = q[2] = INFINITY; q[3] = q[4] = q[5] = -INFINITY;
            q[6] = u2f((uint32_t)SAHBVH_REF_NONE); q[7] = 0.0f;
        }
    }
    /* number the inner children (slot order), resolve leaf refs, drop children that can never hit */
    *n_child = 0;
    int m = 0;
    for (int k = 0; k < n; k++) {
        if (ch[k].ref >= 0) {
            const int32_t id = c->w->n_wide++;
            child_b2[(*n_child)++] = ch[k].ref;
            ch[k].ref = id;
            full[8 * k + 6] = u2f((uint32_t)id);
        } else {
            int32_t lr = leaf_ref(c, &ch[k]);
            if (lr == SAHBVH_REF_NONE) continue;
            ch[k].ref = lr;
        }
        ch[m++] = ch[k];
    }
    if (c->rc != SAHBVH_OK) return c->rc;
    n = m;

    uint32_t node[16];
    memset(node, 0, sizeof node);
    static const int s_word[3] = {3, 14, 15};
    for (int a = 0; a < 3; a++) {
        float p = NAN;
        int e = 0;
        uint8_t qlo[4] = {255, 255, 255, 255}, qhi[4] = {0, 0, 0, 0};
        if (quantise_axis(ch, n, a, &p, &e, qlo, qhi) != 0) {
            p = NAN;
            e = 0;
        }
        node[a] = f2u(p);
        node[s_word[a]] = f2u(ldexpf(1.0f, e));
        for (int k = 0; k < 4; k++) {
            node[4 + a] |= (uint32_t)(k < n ? qlo[k] : 255) << (8 * k);
            node[7 + a] |= (uint32_t)(k < n ? qhi[k] : 0) << (8 * k);
        }
    }
    for (int k = 0; k < 4; k++) node[10 + k] = k < n ? (uint32_t)ch[k].ref : (uint32_t)SAHBVH_REF_NONE;
    int rc = reserve_nodes(c, c->w->n_wide);
    if (rc != SAHBVH_OK) return rc;
    memcpy(c->w->nodes + (size_t)me * 16, node, 64);
    memcpy(c->w->full + (size_t)me * 32, full, 128);
    return SAHBVH_OK;
}

int wide_bvh_from_flat(wide_bvh *w, const flat_bvh *f)
{
    memset(w, 0, sizeof *w);
    if (!f || !f->nodes || f->info.n_inner <= 0) {
        sahbvh_set_error("wide: empty tree");
        return SAHBVH_E_ARG;
    }
    ctx_t c;
    memset(&c, 0, sizeof c);
    c.f = f;
    c.w = w;
    c.cap_nodes = (int64_t)f->info.n_inner / 2 + 16;
    c.rc = SAHBVH_OK;
    w->nodes = (uint32_t *)aligned_alloc(64, (size_t)c.cap_nodes * 64);
    w->full = (float *)aligned_alloc(128, (size_t)c.cap_nodes * 128);
    if (!w->nodes || !w->full) { wide_bvh_release(w); return SAHBVH_E_NOMEM; }
    {
        /* breadth-first: queue entry = (binary node, depth); wide id = position in the queue */
        int64_t qcap = 1024, qn = 1;
        int32_t *q_b2 = (int32_t *)malloc((size_t)qcap * sizeof(int32_t));
        int32_t *q_depth = (int32_t *)malloc((size_t)qcap * sizeof(int32_t));
        if (!q_b2 || !q_depth) { free(q_b2); free(q_depth); wide_bvh_release(w); return SAHBVH_E_NOMEM; }
        q_b2[0] = 0;
        q_depth[0] = 0;
        w->n_wide = 1;
        for (int64_t i = 0; i < qn && c.rc == SAHBVH_OK; i++) {
            int32_t kids[4];
            int nk = 0;
            if (q_depth[i] > c.max_depth) c.max_depth = q_depth[i];
            int rc = emit(&c, q_b2[i], (int32_t)i, kids, &nk);
            if (rc != SAHBVH_OK) { c.rc = rc; break; }
            if (qn + nk > qcap) {
                qcap *= 2;
                int32_t *a = (int32_t *)realloc(q_b2, (size_t)qcap * sizeof(int32_t));
                int32_t *b = a ? (int32_t *)realloc(q_depth, (size_t)qcap * sizeof(int32_t)) : NULL;
                if (a) q_b2 = a;
                if (b) q_depth = b;
                if (!a || !b) { c.rc = SAHBVH_E_NOMEM; break; }
            }
            for (int k = 0; k < nk; k++) { q_b2[qn] = kids[k]; q_depth[qn] = q_depth[i] + 1; qn++; }
        }
        free(q_b2);
        free(q_depth);
    }
    if (c.rc != SAHBVH_OK) {
        wide_bvh_release(w);
        return c.rc;
    }
    if (3 * (c.max_depth + 1) + 4 > SAHBVH_WIDE_STACK_MAX) {
        sahbvh_set_error("wide: tree depth %d exceeds the traversal stack", c.max_depth);
        wide_bvh_release(w);
        return SAHBVH_E_DEPTH;
    }
    w->max_depth = c.max_depth;
    if (!w->special) { /* keep a valid pointer for the upload */
        w->special = (float *)calloc(8, sizeof(float));
        if (!w->special) { wide_bvh_release(w); return SAHBVH_E_NOMEM; }
    }
    return SAHBVH_OK;
}

void wide_bvh_release(wide_bvh *w)
{
    free(w->nodes);
    free(w->full);
    free(w->special);
    w->nodes = NULL;
    w->full = NULL;
    w->special = NULL;
    w->n_wide = w->n_special = 0;
}
