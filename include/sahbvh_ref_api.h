/*
 * sahbvh_ref_api.h -- the reference's own C function surface for the closest-hit path, exported
 * by libsahbvh_b200.so under the SAME names, argument meaning and ownership rules, so a
 * renderer.c / benchmark.c style driver links against this library instead of the
 * reference's bvh.c / hit.c / ray.c / camera.c / sphere.c objects.
 *
 * Every traversal entry point runs on the GPU (sm_100a kernels).  There is no CPU traversal
 * in this library: when no CUDA device is usable the batch calls return a negative status and
 * the by-value calls (which have no error channel in the reference's signature) print the
 * reason to stderr and abort().
 */
#ifndef SAHBVH_REF_API_H
#define SAHBVH_REF_API_H

#include "sahbvh_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- vec3 helpers (include/Custom/vec3.h:10-16, src/vec3.c:17-44,64-69) ----------------- */
Vec3 vec3_sub(Vec3 a, Vec3 b);
Vec3 vec3_normalize(Vec3 a);      /* vec3.c:21-24: double sqrt of a float sum, zero guard */
float vec3_dot(Vec3 a, Vec3 b);
Vec3 vec3_add(Vec3 a, Vec3 b);
Vec3 vec3_multiply(Vec3 v, float t);
Vec3 vec3_cross(Vec3 a, Vec3 b);
Vec3 vec3_random(float min, float max);   /* vec3.c:64-69: three rand() */

/* ---- acceleration structure (include/Custom/bvh.h:21-26, src/bvh.c) ------------------- */

/* bvh.c:19-24 */
AABB create_empty_aabb(void);
/* bvh.c:26-35 */
AABB create_aabb_from_sphere(Sphere *sphere);
/* bvh.c:37-46 */
AABB combine_aabb(AABB a, AABB b);
/* bvh.c:48-57 */
float get_aabb_surface_area(AABB box);
/* bvh.c:59-97  one candidate plane, full sweep (kept for API completeness; the builder below
 * evaluates all 21 planes of a node in ONE pass over 3x8 bins, with identical costs). */
float evaluate_sah(Sphere *spheres, int start, int end, int axis, float split);
/* bvh.c:117-209  Same tree, same in-place permutation of spheres[start,end), one malloc per
 * node (so the reference's free_bvh works on it too).  Leaf iff n<=1 || depth>=40. */
BVHNode *build_bvh_node(Sphere *spheres, int start, int end, int depth);
/* benchmark.c:81-88 (declared in include/Custom/benchmark.h:7).  Also drops the GPU copy
 * that ray_bvh_intersect() cached for this root. */
void free_bvh(BVHNode *node);
/* Drops the GPU copy ray_bvh_intersect() cached for `root` without freeing the tree: for drivers that keep their
 * own free_bvh (benchmark.c:81-88) -- call it first.  (A stale entry is also detected by a fingerprint of the top
 * of the tree on the next call, so forgetting this costs one re-upload, not wrong hits.) */
void sahbvh_scalar_forget(const BVHNode *root);

/* ---- intersection (include/Custom/hit.h:16-18, src/hit.c) ------------------------------ */

/* hit.c:19-39  scalar ray/sphere primitive (host arithmetic, bit-exact). */
HitRecord ray_sphere_intersect(Ray ray, Sphere *sphere);
/* hit.c:49-82  scalar slab test (host arithmetic, bit-exact incl. the d==0 rule). */
int ray_aabb_intersect(Ray ray, AABB box);
/* hit.c:91-109  closest hit through the BVH.  Runs on the GPU: the first call for a given
 * root flattens + uploads the tree (cached per root until free_bvh); each call then traces one
 * ray.  Use gpu_trace_closest_hit() (sahbvh_gpu.h) for batches -- this signature exists so
 * existing callers (renderer.c:31, benchmark.c:242) keep compiling and get identical records. */
HitRecord ray_bvh_intersect(Ray ray, BVHNode *node);

/* ---- rays and camera (include/Custom/ray.h:11, camera.h:17) --------------------------- */

/* ray.c:17-32 with the reference's compile-time 800x600 aspect (constants.h:7-8). */
Ray get_camera_ray(Camera *camera, float u, float v);
/* Same arithmetic with WIDTH/HEIGHT as run-time parameters. */
Ray get_camera_ray_wh(const Camera *camera, float u, float v, int width, int height);
/* camera.c:10-18 */
void camera_update(Camera *camera);
/* main.c:362-365: the primary ray of pixel (x, y) of a width x height frame (u, v, flipped v, get_camera_ray). */
Ray sahbvh_pixel_ray(const Camera *cam, int width, int height, int x, int y);
/* hit.h:8-14 on demand: the record ray_bvh_intersect returned for `ray`, rebuilt from the index of the sphere it
 * hit (gpu_trace_primary_prims / gpu_trace_closest_hit_prims return 4 bytes per ray); prim < 0 = the all-zero
 * miss record.  Pure host arithmetic (ray_sphere_intersect on the winner), bit-identical to the traced record. */
HitRecord sahbvh_hit_from_prim(Ray ray, Sphere *spheres, int32_t prim);

/* ---- primitives / samplers (include/Custom/sphere.h:14-19, src/sphere.c) --------------- */

/* sphere.c:14-16 */
float random_float(float min, float max);
/* sphere.c:19-24 (returns a NORMALISED direction) */
Vec3 random_in_unit_sphere(void);
/* sphere.c:26-32 */
Vec3 random_on_hemisphere(Vec3 normal);
/* sphere.c:34-41 */
Sphere create_benchmark_sphere(Vec3 center);
/* sphere.c:43-50 */
Sphere create_sphere(Vec3 center, float radius);
/* sphere.c:52-59 */
Sphere create_random_sphere(void);

/* ---- scene / ray recipes of the reference's drivers (inline code there, functions here) ---- */

/* benchmark.c:306-314: n spheres, centres U[-500,500)^3, r=0.5 (6 rand() each). */
void sahbvh_fill_benchmark_scene(int n, Sphere *out);
/* main.c:218-221: n x create_random_sphere(). */
void sahbvh_fill_random_scene(int n, Sphere *out);
/* benchmark.c:232-240: origin (0,0,0), direction normalize(U[-1,1]^3) (3 rand() each). */
void sahbvh_fill_benchmark_rays(int64_t n, Ray *out);
/* main.c:203-211: the default camera literal. */
void sahbvh_default_camera(Camera *cam);
/* renderer.c:51-54: one bounce ray {hit.point, random_on_hemisphere(hit.normal)} per hit, in
 * order, skipping misses; returns the number written.  `hits` are 32-byte gpu_hit records. */
int64_t sahbvh_fill_bounce_rays(const void *hits, int64_t n, Ray *out);

#ifdef __cplusplus
}
#endif
#endif /* SAHBVH_REF_API_H */
