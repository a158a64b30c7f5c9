/*
 * sah_builder.c -- host SAH builder with the SAME topology, bounds and in-place sphere
 * permutation as the reference's build_bvh_node (bvh.c:117-209), computed differently:
 *
 *   reference : for each of the 21 candidate planes (3 axes x 7 planes on the NODE bounds,
 *               bvh.c:143-159) one full sweep of the range (evaluate_sah, bvh.c:59-97).
 *   here      : ONE sweep per node.  The 7 planes of an axis are non-decreasing, so
 *               `center < plane_k` is monotone in k and every sphere falls into one of 8 bins
 *               per axis; a bin keeps a count and the union of its spheres' boxes.  The left
 *               set of plane k is bins[0..k), the right set bins[k..8).  Counts are integers
 *               and box unions are fmin/fmax (exact, order independent), so the 21 costs
 *               0.125f + (NL*SA_L + NR*SA_R) come out bit-identical, including the
 *               0*inf = NaN of an empty side and the strict-< first-best rule.
 *   fallback  : if a node's planes are not finite/monotone (non-finite input), the per-plane
 *               sweep is used so that behaviour stays defined by the same expressions.
 *
 * The partition replays the reference's swap sequence (bvh.c:172-201) because the final array
 * order defines the primitive index and the tie rule.
 * Output: preorder array tree (sah_tree); sah_tree_to_pointers() materialises the
 * reference-layout BVHNode tree, one malloc per node, for build_bvh_node().
 */
#include "sahbvh_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float lo[3], hi[3]; } box3;

static inline void box_reset(box3 *b)
{
    b->lo[0] = b->lo[1] = b->lo[2] = INFINITY;
    b->hi[0] = b->hi[1] = b->hi[2] = -INFINITY;
}
static inline void box_merge(box3 *b, const box3 *o)
{
    for (int a = 0; a < 3; a++) {
        b->lo[a] = fminf(b->lo[a], o->lo[a]);
        b->hi[a] = fmaxf(b->hi[a], o->hi[a]);
    }
}
static inline void sphere_box(const Sphere *s, box3 *b)
{
    float r = s->radius;
    b->lo[0] = s->center.x - r; b->lo[1] = s->center.y - r; b->lo[2] = s->center.z - r;
    b->hi[0] = s->center.x + r; b->hi[1] = s->center.y + r; b->hi[2] = s->center.z + r;
}
static inline float box_area(const box3 *b)
{
    float dx = b->hi[0] - b->lo[0], dy = b->hi[1] - b->lo[1], dz = b->hi[2] - b->lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

int sah_tree_init(sah_tree *t, int64_t cap)
{
    if (cap < 16) cap = 16;
    t->nodes = (sah_node *)malloc((size_t)cap * sizeof(sah_node));
    t->used = 0;
    t->cap = t->nodes ? cap : 0;
    return t->nodes ? SAHBVH_OK : SAHBVH_E_NOMEM;
}
void sah_tree_free(sah_tree *t)
{
    free(t->nodes);
    t->nodes = NULL;
    t->used = t->cap = 0;
}
static int64_t tree_alloc(sah_tree *t)
{
    if (t->used == t->cap) {
        int64_t ncap = t->cap * 2;
        sah_node *nn = (sah_node *)realloc(t->nodes, (size_t)ncap * sizeof(sah_node));
        if (!nn) return -1;
        t->nodes = nn;
        t->cap = ncap;
    }
    return t->used++;
}

typedef struct { int count; box3 box; } sah_bin;

/* returns node id, or -1 on allocation failure */
static int64_t build_range(sah_tree *t, Sphere *s, int start, int end, int depth)
{
    int64_t me = tree_alloc(t);
    if (me < 0) return -1;

    box3 nb;
    box_reset(&nb);
    for (int i = start; i < end; i++) {
        box3 sb;
        sphere_box(&s[i], &sb);
        box_merge(&nb, &sb);
    }
    {
        sah_node *nd = &t->nodes[me];
        memcpy(nd->lo, nb.lo, sizeof nb.lo);
        memcpy(nd->hi, nb.hi, sizeof nb.hi);
    }
    int n = end - start;
    if (n <= 1 || depth >= SAHBVH_LEAF_DEPTH) {
        sah_node *nd = &t->nodes[me];
        nd->left = nd->right = -1;
        nd->sphere = start;
        nd->count = n;
        return me;
    }

    /* candidate planes (bvh.c:146-159): lo + (k/8)*(hi-lo), k = 1..7 */
    float plane[3][7];
    int regular = 1;
    for (int a = 0; a < 3; a++) {
        float ext = nb.hi[a] - nb.lo[a];
        for (int k = 1; k < 8; k++) {
            float p = nb.lo[a] + ((float)k / 8.0f) * ext;
            plane[a][k - 1] = p;
            if (!isfinite(p) || (k > 1 && p < plane[a][k - 2])) regular = 0;
        }
    }

    float best_cost = INFINITY, best_split = 0.0f; /* bvh.c:139-141 */
    int best_axis = 0;

    if (regular) {
        sah_bin bins[3][8];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 8; b++) { bins[a][b].count = 0; box_reset(&bins[a][b].box); }
        for (int i = start; i < end; i++) {
            box3 sb;
            sphere_box(&s[i], &sb);
            const float *c = &s[i].center.x;
            for (int a = 0; a < 3; a++) {
                const float *p = plane[a];
                float ca = c[a];
                /* bin = number of planes NOT strictly above the centre; monotone planes make the
                 * set {k : centre < plane_k} an upper range, so the sum is that range's start */
                int b = !(ca < p[0]) + !(ca < p[1]) + !(ca < p[2]) + !(ca < p[3]) +
                        !(ca < p[4]) + !(ca < p[5]) + !(ca < p[6]);
                bins[a][b].count++;
                box_merge(&bins[a][b].box, &sb);
            }
        }
        for (int a = 0; a < 3; a++) {
            /* suffix unions: right[k] = bins[k..8) */
            box3 right[9];
            int rcount[9];
            box_reset(&right[8]);
            rcount[8] = 0;
            for (int b = 7; b >= 0; b--) {
                right[b] = right[b + 1];
                box_merge(&right[b], &bins[a][b].box);
                rcount[b] = rcount[b + 1] + bins[a][b].count;
            }
            box3 left;
            box_reset(&left);
            int lcount = 0;
            for (int k = 1; k < 8; k++) {
                box_merge(&left, &bins[a][k - 1].box);
                lcount += bins[a][k - 1].count;
                float cost = 0.125f + ((float)lcount * box_area(&left) + (float)rcount[k] * box_area(&right[k]));
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = plane[a][k - 1]; }
            }
        }
    } else {
        for (int a = 0; a < 3; a++)
            for (int k = 1; k < 8; k++) {
                float cost = evaluate_sah(s, start, end, a, plane[a][k - 1]);
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = plane[a][k - 1]; }
            }
    }

    /* in-place partition, same swap sequence as bvh.c:172-201 */
    int mid = start;
    for (int i = start; i < end; i++) {
        if ((&s[i].center.x)[best_axis] < best_split) {
            if (i != mid) { Sphere tmp = s[i]; s[i] = s[mid]; s[mid] = tmp; }
            mid++;
        }
    }

    int64_t l = build_range(t, s, start, mid, depth + 1);
    if (l < 0) return -1;
    int64_t r = build_range(t, s, mid, end, depth + 1);
    if (r < 0) return -1;
    sah_node *nd = &t->nodes[me];
    nd->left = (int32_t)l;
    nd->right = (int32_t)r;
    nd->sphere = -1;
    nd->count = 0;
    return me;
}

int sah_build(sah_tree *t, Sphere *spheres, int start, int end, int depth)
{
    t->used = 0;
    if (build_range(t, spheres, start, end, depth) < 0) {
        sahbvh_set_error("sah_build: out of memory after %lld nodes", (long long)t->used);
        return SAHBVH_E_NOMEM;
    }
    return SAHBVH_OK;
}

/* reference-layout pointer tree: one malloc per node so free_bvh (ours or the reference's,
 * benchmark.c:81-88) can release it node by node */
BVHNode *sah_tree_to_pointers(const sah_tree *t, Sphere *spheres)
{
    if (t->used <= 0) return NULL;
    BVHNode **ptr = (BVHNode **)malloc((size_t)t->used * sizeof(BVHNode *));
    if (!ptr) return NULL;
    for (int64_t i = 0; i < t->used; i++) {
        ptr[i] = (BVHNode *)malloc(sizeof(BVHNode));
        if (!ptr[i]) {
            for (int64_t j = 0; j < i; j++) free(ptr[j]);
            free(ptr);
            return NULL;
        }
    }
    for (int64_t i = 0; i < t->used; i++) {
        const sah_node *n = &t->nodes[i];
        BVHNode *p = ptr[i];
        p->bounds.min.x = n->lo[0]; p->bounds.min.y = n->lo[1]; p->bounds.min.z = n->lo[2];
        p->bounds.max.x = n->hi[0]; p->bounds.max.y = n->hi[1]; p->bounds.max.z = n->hi[2];
        if (n->sphere >= 0) {
            p->left = p->right = NULL;
            p->sphere = &spheres[n->sphere];
            p->sphere_count = n->count;
        } else {
            p->left = ptr[n->left];
            p->right = ptr[n->right];
            p->sphere = NULL;
            p->sphere_count = 0;
        }
    }
    BVHNode *root = ptr[0];
    free(ptr);
    return root;
}

BVHNode *build_bvh_node(Sphere *spheres, int start, int end, int depth)
{
    sah_tree t;
    int64_t guess = 2 * (int64_t)(end > start ? end - start : 1) + 16;
    if (sah_tree_init(&t, guess) != SAHBVH_OK) return NULL;
    BVHNode *root = NULL;
    if (sah_build(&t, spheres, start, end, depth) == SAHBVH_OK) root = sah_tree_to_pointers(&t, spheres);
    sah_tree_free(&t);
    return root;
}

void free_bvh(BVHNode *node)
{
    if (!node) return;
    sahbvh_scalar_forget(node);
    /* explicit stack, grows on demand */
    size_t cap = 128, sp = 0;
    BVHNode **stk = (BVHNode **)malloc(cap * sizeof(BVHNode *));
    if (!stk) return;
    stk[sp++] = node;
    while (sp > 0) {
        BVHNode *n = stk[--sp];
        if (sp + 2 > cap) {
            cap *= 2;
            BVHNode **ns = (BVHNode **)realloc(stk, cap * sizeof(BVHNode *));
            if (!ns) break;
            stk = ns;
        }
        if (n->left) stk[sp++] = n->left;
        if (n->right) stk[sp++] = n->right;
        free(n);
    }
    free(stk);
}
