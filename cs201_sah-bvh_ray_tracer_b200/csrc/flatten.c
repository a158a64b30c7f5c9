/*
 * flatten.c -- pointer tree (reference layout, bvh.h:12-18) -> preorder array tree ->
 * device-ready flattened form (see sahbvh_internal.h for the 64-byte node layout).
 *
 * What the flattened tree must preserve:
 *   - the set of leaves, each with the box the reference tests before the sphere
 *     (node->bounds of the leaf) and the ONE sphere it tests (spheres[start]);
 *   - the post-build sphere order (index = hit.object - spheres, defines tie breaking);
 *   - reachability of sphere_count==0 leaves (their own box always passes; the parent's does not).
 * Interior boxes only cull, so the GPU is free to test them conservatively.
 */
#include "sahbvh_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- pointer tree -> preorder array ----------------------------------------------------- */
int sah_tree_from_pointers(sah_tree *t, const BVHNode *root, const Sphere *base, int n)
{
    if (!root || (!base && n > 0) || n < 0) {
        sahbvh_set_error("flatten: NULL root/spheres or negative count");
        return SAHBVH_E_ARG;
    }
    typedef struct { const BVHNode *node; int64_t parent; int is_right; } item;
    size_t cap = 256, sp = 0;
    item *stk = (item *)malloc(cap * sizeof(item));
    if (!stk) return SAHBVH_E_NOMEM;
    t->used = 0;
    stk[sp++] = (item){root, -1, 0};
    int rc = SAHBVH_OK;
    while (sp > 0) {
        item it = stk[--sp];
        if (t->used == t->cap) {
            int64_t ncap = t->cap * 2;
            sah_node *nn = (sah_node *)realloc(t->nodes, (size_t)ncap * sizeof(sah_node));
            if (!nn) { rc = SAHBVH_E_NOMEM; break; }
            t->nodes = nn;
            t->cap = ncap;
        }
        int64_t me = t->used++;
        sah_node *nd = &t->nodes[me];
        const BVHNode *p = it.node;
        nd->lo[0] = p->bounds.min.x; nd->lo[1] = p->bounds.min.y; nd->lo[2] = p->bounds.min.z;
        nd->hi[0] = p->bounds.max.x; nd->hi[1] = p->bounds.max.y; nd->hi[2] = p->bounds.max.z;
        if (it.parent >= 0) {
            if (it.is_right) t->nodes[it.parent].right = (int32_t)me;
            else t->nodes[it.parent].left = (int32_t)me;
        }
        if (p->sphere != NULL) { /* hit.c:98 leaf criterion */
            ptrdiff_t byte_off = (const char *)p->sphere - (const char *)base;
            if (byte_off < 0 || byte_off % (ptrdiff_t)sizeof(Sphere) != 0 ||
                byte_off / (ptrdiff_t)sizeof(Sphere) > n) {
                sahbvh_set_error("flatten: leaf sphere pointer does not alias the sphere array (offset %lld, n=%d)",
                                 (long long)byte_off, n);
                rc = SAHBVH_E_ARG;
                break;
            }
            nd->left = nd->right = -1;
            nd->sphere = (int32_t)(byte_off / (ptrdiff_t)sizeof(Sphere));
            nd->count = p->sphere_count;
            continue;
        }
        if (!p->left || !p->right) {
            sahbvh_set_error("flatten: inner node with a NULL child");
            rc = SAHBVH_E_ARG;
            break;
        }
        nd->left = nd->right = -1;
        nd->sphere = -1;
        nd->count = 0;
        if (sp + 2 > cap) {
            cap *= 2;
            item *ns = (item *)realloc(stk, cap * sizeof(item));
            if (!ns) { rc = SAHBVH_E_NOMEM; break; }
            stk = ns;
        }
        /* push right first so the left subtree is numbered first (preorder) */
        stk[sp++] = (item){p->right, me, 1};
        stk[sp++] = (item){p->left, me, 0};
    }
    free(stk);
    return rc;
}

/* ---- preorder array -> 64-byte inner nodes + float4 spheres ------------------------------ */
static inline int32_t float_bits(float f)
{
    int32_t i;
    memcpy(&i, &f, 4);
    return i;
}
static inline float bits_float(int32_t i)
{
    float f;
    memcpy(&f, &i, 4);
    return f;
}

static void child_box(const sah_tree *t, const sah_node *parent, int32_t child, float lo[3], float hi[3])
{
    const sah_node *c = &t->nodes[child];
    if (c->sphere >= 0 && c->count == 0) { /* Q3: empty leaf is reachable iff the parent's box passes */
        memcpy(lo, parent->lo, 12);
        memcpy(hi, parent->hi, 12);
    } else {
        memcpy(lo, c->lo, 12);
        memcpy(hi, c->hi, 12);
    }
}

int flat_bvh_from_tree(flat_bvh *f, const sah_tree *t, const Sphere *spheres, int n)
{
    memset(f, 0, sizeof *f);
    if (t->used <= 0 || n < 0) {
        sahbvh_set_error("flatten: empty tree");
        return SAHBVH_E_ARG;
    }
    int64_t total = t->used;
    /* inner ids in preorder among inner nodes */
    int32_t *inner_id = (int32_t *)malloc((size_t)total * sizeof(int32_t));
    int32_t *depth = (int32_t *)malloc((size_t)total * sizeof(int32_t));
    if (!inner_id || !depth) { free(inner_id); free(depth); return SAHBVH_E_NOMEM; }
    int64_t n_inner = 0, n_leaf = 0, multi = 0, empty = 0;
    int max_depth = 0;
    depth[0] = 0;
    for (int64_t i = 0; i < total; i++) {
        const sah_node *nd = &t->nodes[i];
        if (nd->sphere >= 0) {
            inner_id[i] = -1;
            n_leaf++;
            if (nd->count > 1) multi++;
            if (nd->count == 0) empty++;
            if (nd->sphere > n) {
                sahbvh_set_error("flatten: leaf sphere index %d beyond n=%d", nd->sphere, n);
                free(inner_id); free(depth);
                return SAHBVH_E_ARG;
            }
        } else {
            inner_id[i] = (int32_t)n_inner++;
            depth[nd->left] = depth[i] + 1;   /* preorder: children come after their parent */
            depth[nd->right] = depth[i] + 1;
        }
        if (depth[i] > max_depth) max_depth = depth[i];
    }
    int synthetic_root = (t->nodes[0].sphere >= 0);
    int64_t n_out = synthetic_root ? 1 : n_inner;
    if (max_depth + 2 > SAHBVH_STACK_MAX) {
        sahbvh_set_error("flatten: tree depth %d exceeds the traversal stack (%d)", max_depth, SAHBVH_STACK_MAX);
        free(inner_id); free(depth);
        return SAHBVH_E_DEPTH;
    }

    f->nodes = (float *)aligned_alloc(64, (size_t)n_out * 64);
    f->spheres4 = (float *)aligned_alloc(64, (((size_t)n + 1) * 16 + 63) / 64 * 64);
    f->colors = (uint32_t *)malloc(((size_t)n + 1) * sizeof(uint32_t));
    if (!f->nodes || !f->spheres4 || !f->colors) {
        free(inner_id); free(depth);
        flat_bvh_release(f);
        return SAHBVH_E_NOMEM;
    }

    if (synthetic_root) {
        /* the whole tree is one leaf: wrap it as the left child of a root whose right child is absent */
        const sah_node *leaf = &t->nodes[0];
        float *w = f->nodes;
        float lo[3], hi[3];
        memcpy(lo, leaf->lo, 12);
        memcpy(hi, leaf->hi, 12);
        w[0] = lo[0]; w[1] = lo[1]; w[2] = lo[2]; w[3] = hi[0];
        w[4] = hi[1]; w[5] = hi[2]; w[6] = bits_float(~leaf->sphere); w[7] = 0.0f;
        w[8] = INFINITY; w[9] = INFINITY; w[10] = INFINITY; w[11] = -INFINITY;
        w[12] = -INFINITY; w[13] = -INFINITY; w[14] = bits_float(SAHBVH_REF_NONE); w[15] = 0.0f;
    } else {
        for (int64_t i = 0; i < total; i++) {
            const sah_node *nd = &t->nodes[i];
            if (nd->sphere >= 0) continue;
            float *w = f->nodes + (size_t)inner_id[i] * 16;
            float llo[3], lhi[3], rlo[3], rhi[3];
            child_box(t, nd, nd->left, llo, lhi);
            child_box(t, nd, nd->right, rlo, rhi);
            const sah_node *l = &t->nodes[nd->left], *r = &t->nodes[nd->right];
            int32_t lref = l->sphere >= 0 ? ~l->sphere : inner_id[nd->left];
            int32_t rref = r->sphere >= 0 ? ~r->sphere : inner_id[nd->right];
            w[0] = llo[0]; w[1] = llo[1]; w[2] = llo[2]; w[3] = lhi[0];
            w[4] = lhi[1]; w[5] = lhi[2]; w[6] = bits_float(lref); w[7] = 0.0f;
            w[8] = rlo[0]; w[9] = rlo[1]; w[10] = rlo[2]; w[11] = rhi[0];
            w[12] = rhi[1]; w[13] = rhi[2]; w[14] = bits_float(rref); w[15] = 0.0f;
        }
    }
    (void)float_bits;

    float rmax = 0.0f;
    float slo[3] = {INFINITY, INFINITY, INFINITY}, shi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; i++) {
        const Sphere *s = &spheres[i];
        float *q = f->spheres4 + (size_t)i * 4;
        q[0] = s->center.x; q[1] = s->center.y; q[2] = s->center.z; q[3] = s->radius;
        memcpy(&f->colors[i], &s->color, 4);
        if (s->radius < 0.0f) {
            /* bvh.c:26-35 builds an INVERTED box for a negative radius: the slab test (hit.c:49-82, min/max per axis)
             * still treats the leaf as [c-|r|, c+|r|], but the ancestors' bounds no longer contain that interval, so the
             * reference's recursion can cull a sphere whose own leaf test passes.  The leaf-only equivalence this library
             * rests on needs containment; refuse loudly instead of returning different hits. */
            sahbvh_set_error("flatten: sphere %d has a negative radius (%g): not supported, the reference's BVH result for it "
                             "depends on ancestor boxes that do not contain it", i, (double)s->radius);
            free(inner_id); free(depth);
            flat_bvh_release(f);
            return SAHBVH_E_ARG;
        }
        float ar = fabsf(s->radius);
        if (ar > rmax) rmax = ar;
        const float *c = &s->center.x;
        for (int a = 0; a < 3; a++) {
            slo[a] = fminf(slo[a], c[a] - ar);
            shi[a] = fmaxf(shi[a], c[a] + ar);
        }
    }
    {
        float *q = f->spheres4 + (size_t)n * 4; /* index n: the reference's out-of-bounds read, never a hit */
        q[0] = q[1] = q[2] = q[3] = NAN;
        f->colors[n] = 0;
    }

    f->info.n_spheres = n;
    f->info.n_inner = (int32_t)n_out;
    f->info.n_leaf_refs = (int32_t)n_leaf;
    f->info.max_depth = max_depth + (synthetic_root ? 1 : 0);
    f->info.multi_leaves = (int32_t)multi;
    f->info.empty_leaves = (int32_t)empty;
    f->info.max_radius = rmax;
    memcpy(f->info.scene_lo, slo, 12);
    memcpy(f->info.scene_hi, shi, 12);
    f->info.device_bytes = (int64_t)n_out * 64 + ((int64_t)n + 1) * 20;
    free(inner_id);
    free(depth);
    return SAHBVH_OK;
}

void flat_bvh_release(flat_bvh *f)
{
    free(f->nodes);
    free(f->spheres4);
    free(f->colors);
    f->nodes = NULL;
    f->spheres4 = NULL;
    f->colors = NULL;
}
