/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, array-based) of the reference's closest-hit path.  It exists so
 * that parity tests, __graft_entry__.smoke() and bench.py's cpu_baseline leg have a checker
 * that travels to the GPU box (where /root/reference does not exist).  The product library
 * (libsahbvh_b200.so) never includes, links or calls anything in this directory.
 *
 * Parity pin: tests/test_oracle_pin.py checks every function here bit-for-bit against the
 * reference's own sources compiled by oracle/Makefile (oracle/_ref/libref*.so) and against
 * the committed fixtures under tests/golden/ (generated from that same compiled reference by
 * tests/golden/make_golden.py).  The reference has no tests of its own.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

typedef struct { float x, y, z; } orc_vec3;                     /* vec3.h:3-7      12 B */
typedef struct { orc_vec3 origin, direction; } orc_ray;         /* ray.h:5-8       24 B */
typedef struct { orc_vec3 center; float radius; uint8_t rgba[4]; } orc_sphere; /* sphere.h:7-11 20 B */
typedef struct {                                                /* camera.h:5-14   64 B */
    orc_vec3 position, forward, right, up;
    float yaw, pitch, fov;
    int move;
} orc_camera;
typedef struct {                                                /* 32 B, == gpu_hit */
    float t;
    int32_t prim;
    float px, py, pz;
    float nx, ny, nz;
} orc_hit;
typedef struct {                                                /* preorder array tree */
    float lo[3], hi[3];
    int32_t left, right;   /* preorder ids, -1 on leaves */
    int32_t sphere;        /* leaf: index of the ONE sphere the traversal tests; -1 on inner nodes */
    int32_t count;         /* leaf: sphere_count (0, 1 or >1) */
} orc_node;

typedef struct {
    int64_t nodes, leaves, multi_leaves, empty_leaves;
    int max_depth;
} orc_stats;

/* scenes / rays (glibc rand() stream) */
void orc_srand(unsigned seed);
void orc_scene_benchmark(int n, orc_sphere *out);
void orc_scene_random(int n, orc_sphere *out);
void orc_rays_benchmark(int64_t n, orc_ray *out);
void orc_default_camera(orc_camera *cam);
void orc_camera_update(orc_camera *cam);
orc_ray orc_camera_ray(const orc_camera *cam, float u, float v, int width, int height);
void orc_primary_rays(const orc_camera *cam, int width, int height, int y0, int y1, orc_ray *out);
int64_t orc_bounce_rays(const orc_hit *hits, int64_t n, orc_ray *out);

/* build: permutes spheres[start,end) in place; returns node count (preorder array), or -1 */
int64_t orc_build(orc_sphere *spheres, int start, int end, int depth, orc_node *out, int64_t cap);
void orc_tree_stats(const orc_node *nodes, int depth0, orc_stats *out);

/* scalar queries */
int orc_ray_aabb(orc_ray ray, const float lo[3], const float hi[3]);
int orc_ray_sphere(orc_ray ray, const orc_sphere *s, int32_t index, orc_hit *out);
void orc_intersect(const orc_node *nodes, const orc_sphere *spheres, int n_spheres, orc_ray ray, orc_hit *out);

/* batch loops (pthreads over rays); return wall seconds of the loop */
double orc_trace_bvh(const orc_node *nodes, const orc_sphere *spheres, int n_spheres,
                     const orc_ray *rays, int64_t n, orc_hit *out, int threads);
double orc_trace_brute(const orc_sphere *spheres, int n_spheres,
                       const orc_ray *rays, int64_t n, orc_hit *out, int threads);

/* work counters of a nearest-first, best-t-culled walk of the same binary tree:
 * out[0] = inner nodes visited, out[1] = sphere tests, out[2] = max stack depth, summed/maxed over rays.
 * Results are checked against orc_intersect by the tests (the walk returns identical hits). */
void orc_walk_counters(const orc_node *nodes, const orc_sphere *spheres, int n_spheres,
                       const orc_ray *rays, int64_t n, int64_t out[3], orc_hit *hits_or_null);

/* path loop + accumulation (renderer.c:21-76, main.c:356-405); serial rand() stream */
void orc_render(const orc_node *nodes, const orc_sphere *spheres, int n_spheres, const orc_camera *cam,
                int width, int height, int y0, int y1, int max_depth, int use_bvh, orc_vec3 *samples, uint8_t *rgba);
void orc_accumulate(float *accum, int frames, const uint8_t *rgba, int64_t n_pixels, uint8_t *avg);

#endif
