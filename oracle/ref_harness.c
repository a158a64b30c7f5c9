/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * A thin batch driver around the UNMODIFIED reference sources.  The recipe in
 * oracle/Makefile compiles /root/reference/src/{vec3,camera,sphere,ray,bvh,hit,renderer,
 * benchmark}.c where they lie, plus this file, into oracle/_ref/libref.so.  Nothing of the
 * reference is copied into this repository; this file only *calls* the reference's
 * functions (build_bvh_node, ray_bvh_intersect, ray_sphere_intersect, get_camera_ray,
 * create_benchmark_sphere, create_random_sphere, random_on_hemisphere, free_bvh) in loops so
 * that Python tests and bench.py can drive them through ctypes.
 *
 * The loops mirror the reference's drivers:
 *   - scene / ray recipes   : benchmark.c:297-314 and :232-240, main.c:203-221
 *   - per-pixel ray mapping : main.c:356-365
 *   - closest-hit loops     : benchmark.c:230-245 (BVH), renderer.c:33-44 (brute force)
 *   - bounce rays           : renderer.c:51-54
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <stddef.h>
#include <math.h>
#include <time.h>
#include <pthread.h>

#include "Custom/vec3.h"
#include "Custom/camera.h"
#include "Custom/ray.h"
#include "Custom/sphere.h"
#include "Custom/bvh.h"
#include "Custom/hit.h"
#include "Custom/benchmark.h"
#include "Custom/constants.h"

/* 32-byte hit record shared by the oracle, this harness and the CUDA path (include/sahbvh_gpu.h). */
typedef struct {
    float t;
    int32_t prim;          /* hit.object - spheres (post-build index), -1 on a miss */
    float px, py, pz;
    float nx, ny, nz;
} refh_hit;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---- layout report: the product headers must be layout-compatible ---- */
void refh_layout(int *out /* 24 ints */)
{
    int k = 0;
    out[k++] = (int)sizeof(Vec3);
    out[k++] = (int)sizeof(Ray);
    out[k++] = (int)sizeof(Camera);
    out[k++] = (int)sizeof(Sphere);
    out[k++] = (int)offsetof(Sphere, radius);
    out[k++] = (int)offsetof(Sphere, color);
    out[k++] = (int)sizeof(AABB);
    out[k++] = (int)sizeof(BVHNode);
    out[k++] = (int)offsetof(BVHNode, left);
    out[k++] = (int)offsetof(BVHNode, right);
    out[k++] = (int)offsetof(BVHNode, sphere);
    out[k++] = (int)offsetof(BVHNode, sphere_count);
    out[k++] = (int)sizeof(HitRecord);
    out[k++] = (int)offsetof(HitRecord, point);
    out[k++] = (int)offsetof(HitRecord, normal);
    out[k++] = (int)offsetof(HitRecord, hit_something);
    out[k++] = (int)offsetof(HitRecord, object);
    out[k++] = WIDTH;
    out[k++] = HEIGHT;
    out[k++] = MAX_DEPTH;
    while (k < 24) out[k++] = 0;
}

float refh_epsilon(void) { return EPSILON; }

/* ---- scenes (rand() stream; caller decides whether to reseed) ---- */
void refh_srand(unsigned seed) { srand(seed); }

/* benchmark.c:306-314 : centres U[-500,500)^3, r = 0.5, 6 rand() per sphere */
void refh_scene_benchmark(int n, Sphere *out)
{
    float world_size = 1000.0f;
    for (int j = 0; j < n; j++) {
        Vec3 center = {
            (float)rand() / RAND_MAX * world_size - world_size / 2,
            (float)rand() / RAND_MAX * world_size - world_size / 2,
            (float)rand() / RAND_MAX * world_size - world_size / 2};
        out[j] = create_benchmark_sphere(center);
    }
}

/* main.c:218-221 */
void refh_scene_random(int n, Sphere *out)
{
    for (int i = 0; i < n; i++) out[i] = create_random_sphere();
}

/* benchmark.c:232-240 : origin 0, direction normalize(U[-1,1]^3) */
void refh_rays_benchmark(int64_t n, Ray *out)
{
    for (int64_t i = 0; i < n; i++) {
        Vec3 dir = {
            (float)rand() / RAND_MAX * 2 - 1,
            (float)rand() / RAND_MAX * 2 - 1,
            (float)rand() / RAND_MAX * 2 - 1};
        dir = vec3_normalize(dir);
        Ray ray = {{0, 0, 0}, dir};
        out[i] = ray;
    }
}

/* main.c:203-211 default camera literal */
void refh_default_camera(Camera *cam)
{
    Camera c = {
        .position = {0, 4, 50},
        .forward = {0, 0, -1},
        .right = {1, 0, 0},
        .up = {0, 1, 0},
        .yaw = -M_PI,
        .pitch = 0,
        .fov = 45.0f,
        .move = 0};
    *cam = c;
}

/* main.c:356-365 : row-major y*WIDTH+x, WIDTH/HEIGHT are this build's compile-time constants */
void refh_primary_rays(Camera *cam, Ray *out)
{
    float aspect_ratio = (float)WIDTH / (float)HEIGHT;
    for (int y = 0; y < HEIGHT; y++) {
        for (int x = 0; x < WIDTH; x++) {
            float u = ((float)x / WIDTH - 0.5f) * aspect_ratio;
            float v = (float)y / HEIGHT - 0.5f;
            out[(size_t)y * WIDTH + x] = get_camera_ray(cam, u, -v);
        }
    }
}

/* one pixel (for spot checks at big resolutions) */
void refh_primary_ray_at(Camera *cam, int x, int y, Ray *out)
{
    float aspect_ratio = (float)WIDTH / (float)HEIGHT;
    float u = ((float)x / WIDTH - 0.5f) * aspect_ratio;
    float v = (float)y / HEIGHT - 0.5f;
    *out = get_camera_ray(cam, u, -v);
}

/* renderer.c:51-54 : bounce ray from a hit; serial rand() stream */
int64_t refh_bounce_rays(const refh_hit *hits, int64_t n, Ray *out)
{
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        if (hits[i].prim < 0) continue;
        Vec3 nrm = {hits[i].nx, hits[i].ny, hits[i].nz};
        Vec3 p = {hits[i].px, hits[i].py, hits[i].pz};
        Vec3 d = random_on_hemisphere(nrm);
        Ray r = {p, d};
        out[m++] = r;
    }
    return m;
}

/* ---- build / free ---- */
BVHNode *refh_build(Sphere *spheres, int start, int end, int depth, double *seconds)
{
    double t0 = now_s();
    BVHNode *root = build_bvh_node(spheres, start, end, depth);
    if (seconds) *seconds = now_s() - t0;
    return root;
}

void refh_free(BVHNode *root) { free_bvh(root); }

/* ---- tree statistics and export (preorder) ---- */
typedef struct {
    int64_t nodes, leaves, multi_leaves, empty_leaves;
    int max_depth;
} refh_stats;

static void stats_rec(const BVHNode *n, int depth, refh_stats *s)
{
    s->nodes++;
    if (depth > s->max_depth) s->max_depth = depth;
    if (n->sphere != NULL) {
        s->leaves++;
        if (n->sphere_count > 1) s->multi_leaves++;
        if (n->sphere_count == 0) s->empty_leaves++;
        return;
    }
    stats_rec(n->left, depth + 1, s);
    stats_rec(n->right, depth + 1, s);
}

void refh_tree_stats(const BVHNode *root, int depth0, refh_stats *out)
{
    memset(out, 0, sizeof(*out));
    stats_rec(root, depth0, out);
}

/* preorder export: bounds[6], left, right (preorder ids, -1 for leaves), sphere index, count */
typedef struct {
    float bounds[6];
    int32_t left, right;
    int32_t sphere;   /* index into the sphere array, -1 for inner nodes */
    int32_t count;
} refh_node;

static int64_t export_rec(const BVHNode *n, const Sphere *base, refh_node *out, int64_t *next)
{
    int64_t me = (*next)++;
    refh_node *o = &out[me];
    o->bounds[0] = n->bounds.min.x; o->bounds[1] = n->bounds.min.y; o->bounds[2] = n->bounds.min.z;
    o->bounds[3] = n->bounds.max.x; o->bounds[4] = n->bounds.max.y; o->bounds[5] = n->bounds.max.z;
    o->count = n->sphere_count;
    if (n->sphere != NULL) {
        o->left = o->right = -1;
        o->sphere = (int32_t)(n->sphere - base);
        return me;
    }
    o->sphere = -1;
    int64_t l = export_rec(n->left, base, out, next);
    int64_t r = export_rec(n->right, base, out, next);
    out[me].left = (int32_t)l;
    out[me].right = (int32_t)r;
    return me;
}

int64_t refh_export_tree(const BVHNode *root, const Sphere *base, refh_node *out)
{
    int64_t next = 0;
    export_rec(root, base, out, &next);
    return next;
}

/* ---- closest-hit loops ---- */
static void pack_hit(const HitRecord *h, const Sphere *base, refh_hit *o)
{
    if (h->hit_something) {
        o->t = h->t;
        o->prim = (int32_t)(h->object - base);
        o->px = h->point.x; o->py = h->point.y; o->pz = h->point.z;
        o->nx = h->normal.x; o->ny = h->normal.y; o->nz = h->normal.z;
    } else {
        memset(o, 0, sizeof(*o));
        o->prim = -1;
    }
}

typedef struct {
    BVHNode *root;
    Sphere *base;
    int n_spheres;
    const Ray *rays;
    refh_hit *out;
    int64_t lo, hi;
    int brute;
} job_t;

static void *job_main(void *arg)
{
    job_t *j = (job_t *)arg;
    for (int64_t i = j->lo; i < j->hi; i++) {
        HitRecord h;
        if (!j->brute) {
            h = ray_bvh_intersect(j->rays[i], j->root);            /* benchmark.c:242, renderer.c:31 */
        } else {
            HitRecord closest = {0};                               /* renderer.c:26-44 */
            closest.t = INFINITY;
            for (int s = 0; s < j->n_spheres; s++) {
                HitRecord hit = ray_sphere_intersect(j->rays[i], &j->base[s]);
                if (hit.hit_something && hit.t < closest.t) closest = hit;
            }
            h = closest;
        }
        if (j->out) pack_hit(&h, j->base, &j->out[i]);
    }
    return NULL;
}

static double run_jobs(job_t proto, int64_t n, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    job_t jobs[256];
    double t0 = now_s();
    for (int k = 0; k < threads; k++) {
        jobs[k] = proto;
        jobs[k].lo = n * k / threads;
        jobs[k].hi = n * (k + 1) / threads;
        if (threads == 1) job_main(&jobs[k]);
        else pthread_create(&tid[k], NULL, job_main, &jobs[k]);
    }
    if (threads > 1)
        for (int k = 0; k < threads; k++) pthread_join(tid[k], NULL);
    return now_s() - t0;
}

/* BVH closest hit over a ray array; returns wall seconds of the loop only */
double refh_trace_bvh(BVHNode *root, Sphere *base, const Ray *rays, int64_t n, refh_hit *out, int threads)
{
    job_t p = {root, base, 0, rays, out, 0, 0, 0};
    return run_jobs(p, n, threads);
}

/* brute-force closest hit (renderer.c:33-44 semantics: ties -> lowest index) */
double refh_trace_brute(Sphere *base, int n_spheres, const Ray *rays, int64_t n, refh_hit *out, int threads)
{
    job_t p = {NULL, base, n_spheres, rays, out, 0, 0, 1};
    return run_jobs(p, n, threads);
}

/* the two timed loops exactly as the reference runs them (clock(), rand() inside) */
double refh_benchmark_no_bvh(Sphere *spheres, int n, int rays) { return benchmark_no_bvh(spheres, n, rays); }
double refh_benchmark_with_bvh(BVHNode *root, int n, int rays) { return benchmark_with_bvh(root, n, rays); }

/* ---- the reference's own path loop: main.c:356-366 calling trace_ray (renderer.c:21) with MAX_DEPTH, at this
 * build's compile-time WIDTH x HEIGHT; rgba = SDL_Color bytes per pixel, row-major y*WIDTH+x ---- */
#include "Custom/renderer.h"
void refh_render(Camera *cam, Sphere *spheres, int n_spheres, BVHNode *root, unsigned char *rgba)
{
    float aspect_ratio = (float)WIDTH / (float)HEIGHT;
    for (int y = 0; y < HEIGHT; y++) {
        for (int x = 0; x < WIDTH; x++) {
            float u = ((float)x / WIDTH - 0.5f) * aspect_ratio;
            float v = (float)y / HEIGHT - 0.5f;
            Ray ray = get_camera_ray(cam, u, -v);
            SDL_Color color = trace_ray(ray, spheres, n_spheres, MAX_DEPTH, root);
            unsigned char *o = rgba + 4 * ((size_t)y * WIDTH + x);
            o[0] = color.r; o[1] = color.g; o[2] = color.b; o[3] = color.a;
        }
    }
}

/* main.c:368-405 verbatim on flat arrays (accum[pixel][3]); first != 0 is the camera.move branch */
void refh_accumulate(float *accum, int *frames, const unsigned char *rgba, int first, unsigned char *shown)
{
    if (first) *frames = 1; else (*frames)++;
    int accumulated_frames = *frames;
    for (size_t p = 0; p < (size_t)WIDTH * HEIGHT; p++) {
        SDL_Color color = {rgba[4 * p], rgba[4 * p + 1], rgba[4 * p + 2], rgba[4 * p + 3]};
        float *a = accum + 3 * p;
        if (first) {
            a[0] = (float)color.r / 255.0f; a[1] = (float)color.g / 255.0f; a[2] = (float)color.b / 255.0f;
            shown[4 * p] = color.r; shown[4 * p + 1] = color.g; shown[4 * p + 2] = color.b; shown[4 * p + 3] = color.a;
        } else {
            a[0] += (float)color.r / 255.0f; a[1] += (float)color.g / 255.0f; a[2] += (float)color.b / 255.0f;
            SDL_Color avg_color = {
                (Uint8)(fmin(a[0] / accumulated_frames * 255.0f, 255.0f)),
                (Uint8)(fmin(a[1] / accumulated_frames * 255.0f, 255.0f)),
                (Uint8)(fmin(a[2] / accumulated_frames * 255.0f, 255.0f)),
                255};
            shown[4 * p] = avg_color.r; shown[4 * p + 1] = avg_color.g; shown[4 * p + 2] = avg_color.b; shown[4 * p + 3] = avg_color.a;
        }
    }
}
