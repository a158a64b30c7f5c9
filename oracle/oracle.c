/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference's SAH-BVH build + closest-hit traversal, written
 * against an array tree instead of malloc'd pointer nodes.  Every function cites the reference
 * lines whose arithmetic it restates; float evaluation order is the reference's
 * (left to right, no FMA contraction: build with -ffp-contract=off, never -ffast-math).
 *
 * PARITY PIN: checked bit-for-bit against the reference's own sources compiled by
 * oracle/Makefile (oracle/_ref/libref.so) in tests/test_oracle_pin.py, and against the
 * committed golden fixtures in tests/golden/ generated from that compiled reference.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_EPSILON 0.000001f /* constants.h:6 */
#define ORC_LEAF_DEPTH 40     /* bvh.c:131 (the comment there says 20, the code says 40) */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------ vec3.c:17-44 */
static inline orc_vec3 v_sub(orc_vec3 a, orc_vec3 b) { orc_vec3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline orc_vec3 v_add(orc_vec3 a, orc_vec3 b) { orc_vec3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static inline orc_vec3 v_scale(orc_vec3 a, float s) { orc_vec3 r = {a.x * s, a.y * s, a.z * s}; return r; }
static inline float v_dot(orc_vec3 a, orc_vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline orc_vec3 v_cross(orc_vec3 a, orc_vec3 b)
{
    orc_vec3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
/* vec3.c:21-24: length through the double sqrt of a float sum, rounded to float; zero guard */
static inline orc_vec3 v_unit(orc_vec3 a)
{
    float len = (float)sqrt((double)(a.x * a.x + a.y * a.y + a.z * a.z));
    orc_vec3 zero = {0, 0, 0};
    if (len != 0) {
        orc_vec3 r = {a.x / len, a.y / len, a.z / len};
        return r;
    }
    return zero;
}

/* ------------------------------------------------------------------ RNG recipes */
void orc_srand(unsigned seed) { srand(seed); }

/* sphere.c:14-16 */
static float rnd_range(float lo, float hi) { return lo + ((float)rand() / RAND_MAX) * (hi - lo); }

/* benchmark.c:306-314 + sphere.c:34-41 : 3 rand() for the centre then 3 for the colour */
void orc_scene_benchmark(int n, orc_sphere *out)
{
    const float world = 1000.0f;
    for (int j = 0; j < n; j++) {
        float cx = (float)rand() / RAND_MAX * world - world / 2;
        float cy = (float)rand() / RAND_MAX * world - world / 2;
        float cz = (float)rand() / RAND_MAX * world - world / 2;
        orc_sphere s;
        s.center.x = cx; s.center.y = cy; s.center.z = cz;
        s.radius = 0.5f;
        s.rgba[0] = (uint8_t)(rand() % 256);
        s.rgba[1] = (uint8_t)(rand() % 256);
        s.rgba[2] = (uint8_t)(rand() % 256);
        s.rgba[3] = 255;
        out[j] = s;
    }
}

/* sphere.c:52-59 : x in [-40,40], y in [-20,20], z in [-10,5], r in [0.5,5]; 7 rand() */
void orc_scene_random(int n, orc_sphere *out)
{
    for (int i = 0; i < n; i++) {
        orc_sphere s;
        s.center.x = rnd_range(-40.0f, 40.0f);
        s.center.y = rnd_range(-20.0f, 20.0f);
        s.center.z = rnd_range(-10.0f, 5.0f);
        s.radius = rnd_range(0.5f, 5.0f);
        s.rgba[0] = (uint8_t)(rand() % 256);
        s.rgba[1] = (uint8_t)(rand() % 256);
        s.rgba[2] = (uint8_t)(rand() % 256);
        s.rgba[3] = 255;
        out[i] = s;
    }
}

/* benchmark.c:232-240 */
void orc_rays_benchmark(int64_t n, orc_ray *out)
{
    for (int64_t i = 0; i < n; i++) {
        orc_vec3 d;
        d.x = (float)rand() / RAND_MAX * 2 - 1;
        d.y = (float)rand() / RAND_MAX * 2 - 1;
        d.z = (float)rand() / RAND_MAX * 2 - 1;
        orc_ray r;
        r.origin.x = r.origin.y = r.origin.z = 0;
        r.direction = v_unit(d);
        out[i] = r;
    }
}

/* main.c:203-211 */
void orc_default_camera(orc_camera *cam)
{
    memset(cam, 0, sizeof(*cam));
    cam->position.y = 4; cam->position.z = 50;
    cam->forward.z = -1;
    cam->right.x = 1;
    cam->up.y = 1;
    cam->yaw = (float)-M_PI;
    cam->pitch = 0;
    cam->fov = 45.0f;
    cam->move = 0;
}

/* camera.c:10-18 : double sin/cos rounded to float on assignment */
void orc_camera_update(orc_camera *cam)
{
    cam->forward.x = (float)(cos((double)cam->pitch) * sin((double)cam->yaw));
    cam->forward.y = (float)sin((double)cam->pitch);
    cam->forward.z = (float)(cos((double)cam->pitch) * cos((double)cam->yaw));
    cam->forward = v_unit(cam->forward);
    orc_vec3 world_up = {0, 1, 0};
    cam->right = v_unit(v_cross(cam->forward, world_up));
    cam->up = v_unit(v_cross(cam->right, cam->forward));
}

/* ray.c:17-32 with WIDTH/HEIGHT as run-time parameters */
orc_ray orc_camera_ray(const orc_camera *cam, float u, float v, int width, int height)
{
    float aspect = (float)width / (float)height;
    float fov_rad = (float)((double)cam->fov * (M_PI / (double)180.0f));
    float half_h = (float)tan((double)(fov_rad / 2.0f));
    float half_w = aspect * half_h;
    orc_vec3 horizontal = v_scale(cam->right, 2.0f * half_w);
    orc_vec3 vertical = v_scale(cam->up, 2.0f * half_h);
    orc_vec3 d = cam->forward;
    d = v_add(d, v_scale(horizontal, u));
    d = v_add(d, v_scale(vertical, v));
    orc_ray r;
    r.origin = cam->position;
    r.direction = v_unit(d);
    return r;
}

/* main.c:356-365 : u uses aspect once more, v is flipped; rows [y0,y1), row-major output */
void orc_primary_rays(const orc_camera *cam, int width, int height, int y0, int y1, orc_ray *out)
{
    float aspect = (float)width / (float)height;
    for (int y = y0; y < y1; y++) {
        for (int x = 0; x < width; x++) {
            float u = ((float)x / (float)width - 0.5f) * aspect;
            float v = (float)y / (float)height - 0.5f;
            out[(size_t)(y - y0) * (size_t)width + (size_t)x] = orc_camera_ray(cam, u, -v, width, height);
        }
    }
}

/* sphere.c:19-32 + vec3.c:64-69 */
static orc_vec3 rnd_unit_dir(void)
{
    for (;;) {
        orc_vec3 p;
        p.x = -1.0f + (float)rand() / RAND_MAX * (1.0f - -1.0f);
        p.y = -1.0f + (float)rand() / RAND_MAX * (1.0f - -1.0f);
        p.z = -1.0f + (float)rand() / RAND_MAX * (1.0f - -1.0f);
        float q = v_dot(p, p);
        if (q < 1 && q != 0.0f) return v_unit(p);
    }
}

/* renderer.c:51-54 : origin = un-offset hit point, direction on the normal's hemisphere */
int64_t orc_bounce_rays(const orc_hit *hits, int64_t n, orc_ray *out)
{
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        if (hits[i].prim < 0) continue;
        orc_vec3 nrm = {hits[i].nx, hits[i].ny, hits[i].nz};
        orc_vec3 d = rnd_unit_dir();
        if (!((double)v_dot(d, nrm) > 0.0)) d = v_scale(d, -1.0f);
        orc_ray r;
        r.origin.x = hits[i].px; r.origin.y = hits[i].py; r.origin.z = hits[i].pz;
        r.direction = d;
        out[m++] = r;
    }
    return m;
}

/* ------------------------------------------------------------------ bvh.c:19-57 */
typedef struct { float lo[3], hi[3]; } box_t;

static box_t box_empty(void)
{
    box_t b = {{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
    return b;
}
static void box_grow_sphere(box_t *b, const orc_sphere *s)
{
    const float *c = &s->center.x;
    for (int a = 0; a < 3; a++) {
        float lo = c[a] - s->radius, hi = c[a] + s->radius;
        b->lo[a] = fminf(b->lo[a], lo);   /* fmin/fmax on floats are exact: == double versions */
        b->hi[a] = fmaxf(b->hi[a], hi);
    }
}
static float box_area(const box_t *b)
{
    float dx = b->hi[0] - b->lo[0], dy = b->hi[1] - b->lo[1], dz = b->hi[2] - b->lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

/* bvh.c:59-97 : one plane, full sweep; empty side -> 0*inf = NaN */
static float sah_cost(const orc_sphere *s, int start, int end, int axis, float split)
{
    int nl = 0, nr = 0;
    box_t bl = box_empty(), br = box_empty();
    for (int i = start; i < end; i++) {
        float c = (&s[i].center.x)[axis];
        if (c < split) { nl++; box_grow_sphere(&bl, &s[i]); }
        else           { nr++; box_grow_sphere(&br, &s[i]); }
    }
    float al = box_area(&bl), ar = box_area(&br);
    return 0.125f + ((float)nl * al + (float)nr * ar);
}

typedef struct { orc_node *out; int64_t cap, used; int overflow; } build_ctx;

/* bvh.c:117-209 */
static int64_t build_rec(build_ctx *ctx, orc_sphere *s, int start, int end, int depth)
{
    if (ctx->used >= ctx->cap) { ctx->overflow = 1; return -1; }
    int64_t me = ctx->used++;
    box_t b = box_empty();
    for (int i = start; i < end; i++) box_grow_sphere(&b, &s[i]);
    orc_node *nd = &ctx->out[me];
    memcpy(nd->lo, b.lo, sizeof b.lo);
    memcpy(nd->hi, b.hi, sizeof b.hi);
    int n = end - start;
    if (n <= 1 || depth >= ORC_LEAF_DEPTH) {          /* :131-137 leaf: only s[start] is ever tested */
        nd->left = nd->right = -1;
        nd->sphere = start;
        nd->count = n;
        return me;
    }
    float best = INFINITY, best_split = 0;            /* :139-141 fallback axis 0 / split 0.0f */
    int best_axis = 0;
    for (int axis = 0; axis < 3; axis++) {
        for (int i = 1; i < 8; i++) {
            float split = b.lo[axis] + ((float)i / 8.0f) * (b.hi[axis] - b.lo[axis]);   /* :146-159 */
            float cost = sah_cost(s, start, end, axis, split);
            if (cost < best) { best = cost; best_axis = axis; best_split = split; }    /* :163 strict */
        }
    }
    int mid = start;                                   /* :172-201 swap-forward partition */
    for (int i = start; i < end; i++) {
        if ((&s[i].center.x)[best_axis] < best_split) {
            orc_sphere tmp = s[i]; s[i] = s[mid]; s[mid] = tmp;
            mid++;
        }
    }
    int64_t l = build_rec(ctx, s, start, mid, depth + 1);
    int64_t r = build_rec(ctx, s, mid, end, depth + 1);
    nd = &ctx->out[me];
    nd->left = (int32_t)l;
    nd->right = (int32_t)r;
    nd->sphere = -1;
    nd->count = 0;
    return me;
}

int64_t orc_build(orc_sphere *spheres, int start, int end, int depth, orc_node *out, int64_t cap)
{
    build_ctx ctx = {out, cap, 0, 0};
    build_rec(&ctx, spheres, start, end, depth);
    return ctx.overflow ? -1 : ctx.used;
}

static void stats_rec(const orc_node *nodes, int64_t i, int depth, orc_stats *s)
{
    const orc_node *n = &nodes[i];
    s->nodes++;
    if (depth > s->max_depth) s->max_depth = depth;
    if (n->sphere >= 0) {
        s->leaves++;
        if (n->count > 1) s->multi_leaves++;
        if (n->count == 0) s->empty_leaves++;
        return;
    }
    stats_rec(nodes, n->left, depth + 1, s);
    stats_rec(nodes, n->right, depth + 1, s);
}
void orc_tree_stats(const orc_node *nodes, int depth0, orc_stats *out)
{
    memset(out, 0, sizeof *out);
    stats_rec(nodes, 0, depth0, out);
}

/* ------------------------------------------------------------------ hit.c:49-82 */
int orc_ray_aabb(orc_ray ray, const float lo[3], const float hi[3])
{
    const float *o = &ray.origin.x, *d = &ray.direction.x;
    float t1[3], t2[3];
    for (int a = 0; a < 3; a++) {
        if (d[a] == 0.0f) { t1[a] = -INFINITY; t2[a] = INFINITY; }   /* regardless of origin (Q7) */
        else { t1[a] = (lo[a] - o[a]) / d[a]; t2[a] = (hi[a] - o[a]) / d[a]; }
    }
    float tmin = fmaxf(fminf(t1[0], t2[0]), fmaxf(fminf(t1[1], t2[1]), fminf(t1[2], t2[2])));
    float tmax = fminf(fmaxf(t1[0], t2[0]), fminf(fmaxf(t1[1], t2[1]), fmaxf(t1[2], t2[2])));
    return tmax >= tmin && tmax > ORC_EPSILON;
}

/* ------------------------------------------------------------------ hit.c:19-39 */
int orc_ray_sphere(orc_ray ray, const orc_sphere *s, int32_t index, orc_hit *out)
{
    orc_vec3 oc = v_sub(ray.origin, s->center);
    float a = v_dot(ray.direction, ray.direction);
    float b = 2.0f * v_dot(oc, ray.direction);
    float c = v_dot(oc, oc) - s->radius * s->radius;
    float disc = b * b - 4 * a * c;
    if (disc > 0) {
        /* :28 evaluated in double (sqrt(double), double subtract and divide), rounded once (Q6) */
        float t = (float)(((double)(-b) - sqrt((double)disc)) / (double)(2.0f * a));
        if (t > ORC_EPSILON) {                          /* near root only (Q5) */
            orc_vec3 p = v_add(ray.origin, v_scale(ray.direction, t));
            orc_vec3 nrm = v_unit(v_sub(p, s->center));
            out->t = t; out->prim = index;
            out->px = p.x; out->py = p.y; out->pz = p.z;
            out->nx = nrm.x; out->ny = nrm.y; out->nz = nrm.z;
            return 1;
        }
    }
    return 0;
}

static const orc_hit MISS = {0.0f, -1, 0, 0, 0, 0, 0, 0};   /* hit.c:20,92 all-zero record (Q8) */

/* ------------------------------------------------------------------ hit.c:91-109 */
static int isect_rec(const orc_node *nodes, const orc_sphere *sp, int n_sp, int64_t i, orc_ray ray, orc_hit *out)
{
    const orc_node *n = &nodes[i];
    if (!orc_ray_aabb(ray, n->lo, n->hi)) return 0;
    if (n->sphere >= 0) {
        /* leaf: exactly one sphere, spheres[start] (Q2).  start == n_spheres is the reference's
         * one-past-the-end read (Q3, undefined behaviour): defined here as a miss. */
        if (n->sphere >= n_sp) return 0;
        return orc_ray_sphere(ray, &sp[n->sphere], n->sphere, out);
    }
    orc_hit l, r;
    int hl = isect_rec(nodes, sp, n_sp, n->left, ray, &l);
    int hr = isect_rec(nodes, sp, n_sp, n->right, ray, &r);
    if (!hl) { if (hr) *out = r; return hr; }
    if (!hr) { *out = l; return 1; }
    *out = (l.t < r.t) ? l : r;                          /* :108 ties -> right (Q1) */
    return 1;
}

void orc_intersect(const orc_node *nodes, const orc_sphere *spheres, int n_spheres, orc_ray ray, orc_hit *out)
{
    orc_hit h;
    if (isect_rec(nodes, spheres, n_spheres, 0, ray, &h)) *out = h;
    else *out = MISS;
}

/* ------------------------------------------------------------------ batch loops */
typedef struct {
    const orc_node *nodes; const orc_sphere *sp; int n_sp;
    const orc_ray *rays; orc_hit *out; int64_t lo, hi; int brute;
} job_t;

static void *job_main(void *arg)
{
    job_t *j = (job_t *)arg;
    for (int64_t i = j->lo; i < j->hi; i++) {
        orc_hit h;
        if (!j->brute) {
            orc_intersect(j->nodes, j->sp, j->n_sp, j->rays[i], &h);
        } else {                                         /* renderer.c:33-44: ties -> lowest index */
            h = MISS;
            float best = INFINITY;
            for (int s = 0; s < j->n_sp; s++) {
                orc_hit c;
                if (orc_ray_sphere(j->rays[i], &j->sp[s], s, &c) && c.t < best) { best = c.t; h = c; }
            }
        }
        if (j->out) j->out[i] = h;
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

double orc_trace_bvh(const orc_node *nodes, const orc_sphere *spheres, int n_spheres,
                     const orc_ray *rays, int64_t n, orc_hit *out, int threads)
{
    job_t p = {nodes, spheres, n_spheres, rays, out, 0, 0, 0};
    return run_jobs(p, n, threads);
}

double orc_trace_brute(const orc_sphere *spheres, int n_spheres,
                       const orc_ray *rays, int64_t n, orc_hit *out, int threads)
{
    job_t p = {NULL, spheres, n_spheres, rays, out, 0, 0, 1};
    return run_jobs(p, n, threads);
}

/* ------------------------------------------------------------------ ordered + culled walk
 * Nearest-child-first, skip a subtree when its entry distance exceeds the best hit so far.
 * Interior tests only cull; the reference test is applied at every node exactly as hit.c does,
 * so the result equals orc_intersect's (asserted by tests).  Counts the work a per-ray GPU
 * traversal of the same binary tree performs. */
static int slab_entry(orc_ray ray, const float lo[3], const float hi[3], float *entry)
{
    const float *o = &ray.origin.x, *d = &ray.direction.x;
    float t1[3], t2[3];
    for (int a = 0; a < 3; a++) {
        if (d[a] == 0.0f) { t1[a] = -INFINITY; t2[a] = INFINITY; }
        else { t1[a] = (lo[a] - o[a]) / d[a]; t2[a] = (hi[a] - o[a]) / d[a]; }
    }
    float tmin = fmaxf(fminf(t1[0], t2[0]), fmaxf(fminf(t1[1], t2[1]), fminf(t1[2], t2[2])));
    float tmax = fminf(fmaxf(t1[0], t2[0]), fminf(fmaxf(t1[1], t2[1]), fmaxf(t1[2], t2[2])));
    *entry = tmin;
    return tmax >= tmin && tmax > ORC_EPSILON;
}

void orc_walk_counters(const orc_node *nodes, const orc_sphere *spheres, int n_spheres,
                       const orc_ray *rays, int64_t n, int64_t out[3], orc_hit *hits)
{
    int64_t inner = 0, tests = 0, max_sp = 0;
    for (int64_t r = 0; r < n; r++) {
        orc_ray ray = rays[r];
        orc_hit best = MISS;
        int have = 0;
        /* generous slack: the computed sphere t can sit slightly before its box entry */
        int32_t stack[128];
        float sentry[128];
        int sp = 0;
        float e0;
        if (slab_entry(ray, nodes[0].lo, nodes[0].hi, &e0)) { stack[sp] = 0; sentry[sp] = e0; sp++; }
        while (sp > 0) {
            sp--;
            int32_t i = stack[sp];
            float e = sentry[sp];
            if (have && e > best.t * 1.01f + 1.0f) continue;
            const orc_node *nd = &nodes[i];
            if (nd->sphere >= 0) {
                if (nd->sphere >= n_spheres) continue;
                tests++;
                orc_hit c;
                if (orc_ray_sphere(ray, &spheres[nd->sphere], nd->sphere, &c)) {
                    if (!have || c.t < best.t || (c.t == best.t && c.prim > best.prim)) { best = c; have = 1; }
                }
                continue;
            }
            inner++;
            float el, er;
            int hl = slab_entry(ray, nodes[nd->left].lo, nodes[nd->left].hi, &el);
            int hr = slab_entry(ray, nodes[nd->right].lo, nodes[nd->right].hi, &er);
            if (hl && hr) {
                if (el <= er) { stack[sp] = nd->right; sentry[sp++] = er; stack[sp] = nd->left; sentry[sp++] = el; }
                else          { stack[sp] = nd->left; sentry[sp++] = el; stack[sp] = nd->right; sentry[sp++] = er; }
            } else if (hl) { stack[sp] = nd->left; sentry[sp++] = el; }
            else if (hr)   { stack[sp] = nd->right; sentry[sp++] = er; }
            if (sp > max_sp) max_sp = sp;
        }
        if (hits) hits[r] = best;
    }
    out[0] = inner; out[1] = tests; out[2] = max_sp;
}

/* ------------------------------------------------------------------ renderer.c:21-76 + main.c:356-405
 * trace_ray as a loop (the recursion's arithmetic in the recursion's order): per pixel, row-major, drawing
 * from the serial rand() stream exactly when the reference does (one random_on_hemisphere per hit, before the
 * recursive call).  samples (optional) records the unit vectors random_in_unit_sphere() returned,
 * [pixel][max_depth] (unused slots zero), so that a parallel implementation can be fed the same draws. */
static void sky_rgba(float dir_y, uint8_t *c)
{
    float t = 0.5f * (dir_y + 1.0f);
    float r = (1.0f - t) * 255 + t * 128, g = (1.0f - t) * 255 + t * 178;
    c[0] = (uint8_t)(int)r; c[1] = (uint8_t)(int)g; c[2] = 255; c[3] = 255;
}

void orc_render(const orc_node *nodes, const orc_sphere *spheres, int n_spheres, const orc_camera *cam,
                int width, int height, int y0, int y1, int max_depth, int use_bvh, orc_vec3 *samples, uint8_t *rgba)
{
    float aspect = (float)width / (float)height;
    for (int y = y0; y < y1; y++) {
        for (int x = 0; x < width; x++) {
            size_t pix = (size_t)(y - y0) * (size_t)width + (size_t)x;
            float u = ((float)x / (float)width - 0.5f) * aspect;
            float v = (float)y / (float)height - 0.5f;
            orc_ray ray = orc_camera_ray(cam, u, -v, width, height);
            uint8_t base[64][3], c[4] = {0, 0, 0, 255};
            int k = 0;
            if (samples) memset(samples + pix * (size_t)max_depth, 0, (size_t)max_depth * sizeof(orc_vec3));
            for (int depth = max_depth;; depth--) {
                if (depth <= 0) { c[0] = c[1] = c[2] = 0; c[3] = 255; break; }      /* renderer.c:23-24 */
                orc_hit h;
                if (use_bvh) orc_intersect(nodes, spheres, n_spheres, ray, &h);
                else {
                    h = MISS;
                    float best = INFINITY;
                    for (int i = 0; i < n_spheres; i++) {                            /* renderer.c:36-43 */
                        orc_hit t;
                        if (orc_ray_sphere(ray, &spheres[i], i, &t) && t.t < best) { best = t.t; h = t; }
                    }
                }
                if (h.prim < 0) { sky_rgba(ray.direction.y, c); break; }             /* renderer.c:66-71 */
                base[k][0] = spheres[h.prim].rgba[0]; base[k][1] = spheres[h.prim].rgba[1]; base[k][2] = spheres[h.prim].rgba[2];
                orc_vec3 nrm = {h.nx, h.ny, h.nz};
                orc_vec3 d = rnd_unit_dir();                                         /* renderer.c:51, sphere.c:26-32 */
                if (samples) samples[pix * (size_t)max_depth + (size_t)k] = d;
                if (!((double)v_dot(d, nrm) > 0.0)) d = v_scale(d, -1.0f);
                ray.origin.x = h.px; ray.origin.y = h.py; ray.origin.z = h.pz;       /* renderer.c:54 */
                ray.direction = d;
                k++;
            }
            for (int j = k - 1; j >= 0; j--)                                         /* renderer.c:56-58 */
                for (int ch = 0; ch < 3; ch++) c[ch] = (uint8_t)(int)((double)base[j][ch] + 0.5 * (double)c[ch]);
            c[3] = 255;
            memcpy(rgba + 4 * pix, c, 4);
        }
    }
}

/* main.c:368-405: frames = count including this frame; frames == 1 restarts */
void orc_accumulate(float *accum, int frames, const uint8_t *rgba, int64_t n_pixels, uint8_t *avg)
{
    for (int64_t p = 0; p < n_pixels; p++) {
        for (int ch = 0; ch < 3; ch++) {
            float v = (float)rgba[4 * p + ch] / 255.0f;
            if (frames <= 1) {
                accum[3 * p + ch] = v;
                avg[4 * p + ch] = rgba[4 * p + ch];
            } else {
                accum[3 * p + ch] += v;
                avg[4 * p + ch] = (uint8_t)(int)(fmin(accum[3 * p + ch] / frames * 255.0f, 255.0f));
            }
        }
        avg[4 * p + 3] = 255;
    }
}
