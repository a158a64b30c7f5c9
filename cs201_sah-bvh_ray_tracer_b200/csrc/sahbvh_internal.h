/* sahbvh_internal.h -- shared between the host C sources and the CUDA translation unit. */
#ifndef SAHBVH_INTERNAL_H
#define SAHBVH_INTERNAL_H

#include "sahbvh_gpu.h"
#include "sahbvh_ref_api.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Intermediate array tree, preorder (root = 0).  One entry per reference BVHNode. */
typedef struct sah_node {
    float lo[3], hi[3]; /* node->bounds (bvh.c:120-125) */
    int32_t left, right; /* preorder ids; -1 on leaves */
    int32_t sphere;      /* leaf: index of spheres[start] -- the only sphere hit.c:98-99 tests; -1 on inner */
    int32_t count;       /* leaf: sphere_count */
} sah_node;

typedef struct sah_tree {
    sah_node *nodes;
    int64_t used, cap;
} sah_tree;

/* Flattened, device-ready form.
 *
 * Inner node = 64 bytes, 64-byte aligned = two 32-byte CHILD RECORDS (left, right), each
 * two 16-byte words so that any kernel fetches a child with aligned 128-bit loads:
 *   c0 = { lo.x, lo.y, lo.z, hi.x }
 *   c1 = { hi.y, hi.z, ref (int32 bits), 0 }
 * child ref >= 0 : index of another inner node;  ref < 0 : leaf, sphere index = ~ref.
 * A leaf's inline box is the reference leaf's node->bounds, except for sphere_count==0
 * leaves (inverted-infinite bounds) which carry the PARENT's bounds: the reference
 * reaches such a leaf iff the parent's own box test passed, and its own box always passes.
 * Sphere index == n_spheres encodes the reference's one-past-the-end read (never hit);
 * SAHBVH_REF_NONE marks the absent sibling of a synthetic root (tree that is a single leaf).
 */
#define SAHBVH_REF_NONE ((int32_t)0x80000000)
#define SAHBVH_STACK_MAX 64

typedef struct flat_bvh {
    float *nodes;     /* n_inner * 16 floats (w3 holds int32 bit patterns) */
    float *spheres4;  /* (n_spheres + 1) * 4 : cx, cy, cz, r ; entry n is a NaN sentinel */
    uint32_t *colors; /* n_spheres RGBA8 (kept for the shading row that comes next) */
    gpu_bvh_info info;
} flat_bvh;

/* 4-wide, 8-bit quantised form used by the ray-buffer kernel (layout and rules: wide.c) */
#define SAHBVH_WIDE_SPECIAL 0x40000000u
#define SAHBVH_WIDE_STACK_MAX 192

typedef struct wide_bvh {
    uint32_t *nodes;   /* n_wide * 16 words, 64-byte aligned */
    float *full;       /* n_wide * 32 floats: the same nodes with four full-precision 32-byte child records */
    float *special;    /* n_special * 8 floats: lo.xyz, hi.x, hi.y, hi.z, sphere index (bits), 0 */
    int32_t n_wide, n_special;
    int32_t max_depth;
} wide_bvh;

void sahbvh_set_error(const char *fmt, ...);

/* sah_builder.c */
int sah_tree_init(sah_tree *t, int64_t cap);
void sah_tree_free(sah_tree *t);
int sah_build(sah_tree *t, Sphere *spheres, int start, int end, int depth);
BVHNode *sah_tree_to_pointers(const sah_tree *t, Sphere *spheres);

/* flatten.c */
int sah_tree_from_pointers(sah_tree *t, const BVHNode *root, const Sphere *base, int n);
int flat_bvh_from_tree(flat_bvh *f, const sah_tree *t, const Sphere *spheres, int n);
void flat_bvh_release(flat_bvh *f);

/* gpu_builder.cu: same tree and permutation as sah_build, built on the device (falls back to sah_build for
 * tiny ranges and when a degenerate input outgrows its node pool) */
int sah_build_gpu(sah_tree *t, Sphere *spheres, int start, int end, int depth, int device);

/* everything gpu_bvh needs, built and left on the device by gpu_builder.cu (device pointers, owned by the caller) */
typedef struct gpu_bvh_parts {
    float *d_nodes;       /* binary flattened nodes, 64 B each */
    float *d_spheres4;    /* (n_spheres + 1) float4 */
    uint32_t *d_colors;   /* (n_spheres + 1) RGBA8 */
    uint32_t *d_wide;     /* 4-wide quantised nodes, 64 B each */
    float *d_wide_full;   /* their full-precision twins, 128 B each */
    float *d_special;     /* special-leaf records, 32 B each */
    int32_t n_wide, n_special;
    gpu_bvh_info info;
} gpu_bvh_parts;
int gpu_build_parts(Sphere *spheres, int n_total, int start, int end, int depth, int device, gpu_bvh_parts *parts);
void gpu_bvh_parts_release(gpu_bvh_parts *parts);

/* wide.c */
int wide_bvh_from_flat(wide_bvh *w, const flat_bvh *f);
void wide_bvh_release(wide_bvh *w);

/* gpu_api.cu */
int gpu_bvh_from_flat(const flat_bvh *f, int device, gpu_bvh **out);
int sahbvh_scalar_trace(const BVHNode *root, const Ray *ray, HitRecord *out); /* ray_bvh_intersect backend */
void sahbvh_scalar_forget(const BVHNode *root);

#ifdef __cplusplus
}
#endif
#endif
