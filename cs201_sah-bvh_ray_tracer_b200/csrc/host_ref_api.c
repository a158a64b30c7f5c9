/*
 * host_ref_api.c -- host-side scalar half of the reference's function surface
 * (include/sahbvh_ref_api.h): vec3 helpers, camera, ray generation, sphere constructors and
 * samplers, AABB helpers and the two scalar intersection primitives.  These are the
 * functions callers use to PREPARE inputs (scenes, cameras, rays); all BVH traversal goes to
 * the GPU (gpu_api.cu).  Compiled with -ffp-contract=off: every product and sum below is
 * rounded separately, as in the reference's -O0 x86-64 build.
 */
#include "sahbvh_internal.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ error channel */
static _Thread_local char g_err[512] = "";

void sahbvh_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *sahbvh_last_error(void) { return g_err; }
const char *sahbvh_version(void) { return "sahbvh_b200 0.1 (sm_100a)"; }

/* ------------------------------------------------------------------ vec3 (vec3.c:17-44) */
Vec3 vec3_sub(Vec3 a, Vec3 b)
{
    Vec3 r;
    r.x = a.x - b.x; r.y = a.y - b.y; r.z = a.z - b.z;
    return r;
}
Vec3 vec3_add(Vec3 a, Vec3 b)
{
    Vec3 r;
    r.x = a.x + b.x; r.y = a.y + b.y; r.z = a.z + b.z;
    return r;
}
Vec3 vec3_multiply(Vec3 v, float t)
{
    Vec3 r;
    r.x = v.x * t; r.y = v.y * t; r.z = v.z * t;
    return r;
}
float vec3_dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
Vec3 vec3_cross(Vec3 a, Vec3 b)
{
    Vec3 r;
    r.x = a.y * b.z - a.z * b.y;
    r.y = a.z * b.x - a.x * b.z;
    r.z = a.x * b.y - a.y * b.x;
    return r;
}
/* vec3.c:21-24: the float sum of squares goes through the DOUBLE sqrt and is rounded back
 * (== correctly rounded sqrtf); zero-length input yields the zero vector. */
Vec3 vec3_normalize(Vec3 a)
{
    float sq = a.x * a.x + a.y * a.y + a.z * a.z;
    float len = (float)sqrt((double)sq);
    Vec3 r = {0.0f, 0.0f, 0.0f};
    if (len != 0) {
        r.x = a.x / len; r.y = a.y / len; r.z = a.z / len;
    }
    return r;
}
/* vec3.c:64-69 */
Vec3 vec3_random(float min, float max)
{
    Vec3 r;
    r.x = min + (float)rand() / RAND_MAX * (max - min);
    r.y = min + (float)rand() / RAND_MAX * (max - min);
    r.z = min + (float)rand() / RAND_MAX * (max - min);
    return r;
}

/* ------------------------------------------------------------------ camera (camera.c:10-18) */
void camera_update(Camera *camera)
{
    double cp = cos((double)camera->pitch), sp = sin((double)camera->pitch);
    double cy = cos((double)camera->yaw), sy = sin((double)camera->yaw);
    Vec3 f;
    f.x = (float)(cp * sy);
    f.y = (float)sp;
    f.z = (float)(cp * cy);
    camera->forward = vec3_normalize(f);
    Vec3 world_up = {0.0f, 1.0f, 0.0f};
    camera->right = vec3_normalize(vec3_cross(camera->forward, world_up));
    camera->up = vec3_normalize(vec3_cross(camera->right, camera->forward));
}

/* ------------------------------------------------------------------ ray generation (ray.c:17-32) */
Ray get_camera_ray_wh(const Camera *camera, float u, float v, int width, int height)
{
    float aspect = (float)width / (float)height;
    float fov_rad = (float)((double)camera->fov * (M_PI / (double)180.0f)); /* double product, rounded on store */
    float half_h = (float)tan((double)(fov_rad / 2.0f));
    float half_w = aspect * half_h;
    Vec3 horizontal = vec3_multiply(camera->right, 2.0f * half_w);
    Vec3 vertical = vec3_multiply(camera->up, 2.0f * half_h);
    Vec3 d = camera->forward;
    d = vec3_add(d, vec3_multiply(horizontal, u));
    d = vec3_add(d, vec3_multiply(vertical, v));
    Ray r;
    r.origin = camera->position;
    r.direction = vec3_normalize(d);
    return r;
}
Ray get_camera_ray(Camera *camera, float u, float v)
{
    return get_camera_ray_wh(camera, u, v, SAHBVH_DEFAULT_WIDTH, SAHBVH_DEFAULT_HEIGHT);
}

/* ------------------------------------------------------------------ spheres / samplers (sphere.c) */
float random_float(float min, float max) { return min + ((float)rand() / RAND_MAX) * (max - min); }

Vec3 random_in_unit_sphere(void)
{
    for (;;) {
        Vec3 p = vec3_random(-1.0f, 1.0f);
        float q = vec3_dot(p, p);
        if (q < 1 && q != 0.0f) return vec3_normalize(p);
    }
}
Vec3 random_on_hemisphere(Vec3 normal)
{
    Vec3 d = random_in_unit_sphere();
    if ((double)vec3_dot(d, normal) > 0.0) return d;
    return vec3_multiply(d, -1.0f);
}
static SDL_Color random_color(void)
{
    SDL_Color c;
    c.r = (uint8_t)(rand() % 256);
    c.g = (uint8_t)(rand() % 256);
    c.b = (uint8_t)(rand() % 256);
    c.a = 255;
    return c;
}
Sphere create_benchmark_sphere(Vec3 center)
{
    Sphere s;
    s.center = center;
    s.radius = 0.5f;
    s.color = random_color();
    return s;
}
Sphere create_sphere(Vec3 center, float radius)
{
    Sphere s;
    memset(&s, 0, sizeof s);
    s.center = center;
    s.radius = radius;
    return s;
}
Sphere create_random_sphere(void)
{
    Sphere s;
    s.center.x = random_float(-40.0f, 40.0f);
    s.center.y = random_float(-20.0f, 20.0f);
    s.center.z = random_float(-10.0f, 5.0f);
    s.radius = random_float(0.5f, 5.0f);
    s.color = random_color();
    return s;
}

/* ------------------------------------------------------------------ driver recipes */
void sahbvh_fill_benchmark_scene(int n, Sphere *out)
{
    const float world_size = 1000.0f; /* benchmark.c:298 */
    for (int j = 0; j < n; j++) {
        Vec3 c;
        c.x = (float)rand() / RAND_MAX * world_size - world_size / 2;
        c.y = (float)rand() / RAND_MAX * world_size - world_size / 2;
        c.z = (float)rand() / RAND_MAX * world_size - world_size / 2;
        out[j] = create_benchmark_sphere(c);
    }
}
void sahbvh_fill_random_scene(int n, Sphere *out)
{
    for (int i = 0; i < n; i++) out[i] = create_random_sphere();
}
void sahbvh_fill_benchmark_rays(int64_t n, Ray *out)
{
    for (int64_t i = 0; i < n; i++) {
        Vec3 d;
        d.x = (float)rand() / RAND_MAX * 2 - 1;
        d.y = (float)rand() / RAND_MAX * 2 - 1;
        d.z = (float)rand() / RAND_MAX * 2 - 1;
        out[i].origin.x = out[i].origin.y = out[i].origin.z = 0.0f;
        out[i].direction = vec3_normalize(d);
    }
}
void sahbvh_default_camera(Camera *cam)
{
    memset(cam, 0, sizeof *cam);
    cam->position.y = 4.0f; cam->position.z = 50.0f;
    cam->forward.z = -1.0f;
    cam->right.x = 1.0f;
    cam->up.y = 1.0f;
    cam->yaw = (float)-M_PI;
    cam->fov = 45.0f;
}
int64_t sahbvh_fill_bounce_rays(const void *hits, int64_t n, Ray *out)
{
    const gpu_hit *h = (const gpu_hit *)hits;
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        if (h[i].prim < 0) continue;
        Vec3 nrm = {h[i].nx, h[i].ny, h[i].nz};
        out[m].origin.x = h[i].px; out[m].origin.y = h[i].py; out[m].origin.z = h[i].pz;
        out[m].direction = random_on_hemisphere(nrm);
        m++;
    }
    return m;
}

/* ------------------------------------------------------------------ AABB helpers (bvh.c:19-57) */
AABB create_empty_aabb(void)
{
    AABB b;
    b.min.x = b.min.y = b.min.z = INFINITY;
    b.max.x = b.max.y = b.max.z = -INFINITY;
    return b;
}
AABB create_aabb_from_sphere(Sphere *sphere)
{
    AABB b;
    float r = sphere->radius;
    b.min.x = sphere->center.x - r; b.min.y = sphere->center.y - r; b.min.z = sphere->center.z - r;
    b.max.x = sphere->center.x + r; b.max.y = sphere->center.y + r; b.max.z = sphere->center.z + r;
    return b;
}
AABB combine_aabb(AABB a, AABB b)
{
    AABB r;
    r.min.x = fminf(a.min.x, b.min.x); r.min.y = fminf(a.min.y, b.min.y); r.min.z = fminf(a.min.z, b.min.z);
    r.max.x = fmaxf(a.max.x, b.max.x); r.max.y = fmaxf(a.max.y, b.max.y); r.max.z = fmaxf(a.max.z, b.max.z);
    return r;
}
float get_aabb_surface_area(AABB box)
{
    float dx = box.max.x - box.min.x, dy = box.max.y - box.min.y, dz = box.max.z - box.min.z;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}
float evaluate_sah(Sphere *spheres, int start, int end, int axis, float split)
{
    int nl = 0, nr = 0;
    AABB bl = create_empty_aabb(), br = create_empty_aabb();
    for (int i = start; i < end; i++) {
        float c = (&spheres[i].center.x)[axis];
        if (c < split) { nl++; bl = combine_aabb(bl, create_aabb_from_sphere(&spheres[i])); }
        else           { nr++; br = combine_aabb(br, create_aabb_from_sphere(&spheres[i])); }
    }
    return 0.125f + ((float)nl * get_aabb_surface_area(bl) + (float)nr * get_aabb_surface_area(br));
}

/* ------------------------------------------------------------------ scalar primitives (hit.c:19-82) */
HitRecord ray_sphere_intersect(Ray ray, Sphere *sphere)
{
    HitRecord rec;
    memset(&rec, 0, sizeof rec);
    Vec3 oc = vec3_sub(ray.origin, sphere->center);
    float a = vec3_dot(ray.direction, ray.direction);
    float b = 2.0f * vec3_dot(oc, ray.direction);
    float c = vec3_dot(oc, oc) - sphere->radius * sphere->radius;
    float disc = b * b - 4 * a * c;
    if (!(disc > 0)) return rec;
    float t = (float)(((double)(-b) - sqrt((double)disc)) / (double)(2.0f * a)); /* hit.c:28 in double */
    if (!(t > SAHBVH_EPSILON)) return rec;
    rec.hit_something = 1;
    rec.t = t;
    rec.point = vec3_add(ray.origin, vec3_multiply(ray.direction, t));
    rec.normal = vec3_normalize(vec3_sub(rec.point, sphere->center));
    rec.object = sphere;
    return rec;
}

int ray_aabb_intersect(Ray ray, AABB box)
{
    const float *o = &ray.origin.x, *d = &ray.direction.x, *lo = &box.min.x, *hi = &box.max.x;
    float near_t = -INFINITY, far_t = INFINITY;
    for (int a = 0; a < 3; a++) {
        float t1, t2;
        if (d[a] == 0.0f) { t1 = -INFINITY; t2 = INFINITY; }      /* hit.c:54-56 regardless of origin */
        else { t1 = (lo[a] - o[a]) / d[a]; t2 = (hi[a] - o[a]) / d[a]; }
        near_t = fmaxf(near_t, fminf(t1, t2));
        far_t = fminf(far_t, fmaxf(t1, t2));
    }
    return far_t >= near_t && far_t > SAHBVH_EPSILON;
}

/* hit.c:91-109 -- GPU-backed; see gpu_api.cu:sahbvh_scalar_trace */
HitRecord ray_bvh_intersect(Ray ray, BVHNode *node)
{
    HitRecord rec;
    memset(&rec, 0, sizeof rec);
    if (sahbvh_scalar_trace(node, &ray, &rec) != SAHBVH_OK) {
        fprintf(stderr, "sahbvh_b200: ray_bvh_intersect needs a CUDA device and has no CPU path: %s\n",
                sahbvh_last_error());
        abort();
    }
    return rec;
}

/* ------------------------------------------------------------------ lazily derived hit fields */
/* main.c:362-365 for one pixel of a width x height frame */
Ray sahbvh_pixel_ray(const Camera *cam, int width, int height, int x, int y)
{
    float aspect = (float)width / (float)height;
    float u = ((float)x / (float)width - 0.5f) * aspect;
    float v = (float)y / (float)height - 0.5f;
    return get_camera_ray_wh(cam, u, -v, width, height);
}

/* The HitRecord (hit.h:8-14) ray_bvh_intersect returned for `ray`, given only the index of the sphere it hit:
 * t, point and normal are pure functions of the ray and that sphere (hit.c:19-39 evaluated once more on the
 * winner), so a caller that keeps the 4-byte id frame derives the other fields here, on demand, bit for bit.
 * prim < 0 (miss) gives the all-zero record of hit.c:92. */
HitRecord sahbvh_hit_from_prim(Ray ray, Sphere *spheres, int32_t prim)
{
    HitRecord rec;
    memset(&rec, 0, sizeof rec);
    if (prim < 0 || !spheres) return rec;
    return ray_sphere_intersect(ray, &spheres[prim]);
}

/* ------------------------------------------------------------------ frame output (in place of SDL_RenderDrawPoint) */
int sahbvh_write_ppm(const char *path, const uint8_t *rgba, int width, int height)
{
    if (!path || !rgba || width <= 0 || height <= 0) {
        sahbvh_set_error("sahbvh_write_ppm: bad arguments");
        return SAHBVH_E_ARG;
    }
    FILE *f = fopen(path, "wb");
    if (!f) {
        sahbvh_set_error("sahbvh_write_ppm: cannot open %s", path);
        return SAHBVH_E_ARG;
    }
    fprintf(f, "P6\n%d %d\n255\n", width, height);
    for (size_t i = 0; i < (size_t)width * (size_t)height; i++) fwrite(rgba + 4 * i, 1, 3, f);
    fclose(f);
    return SAHBVH_OK;
}
