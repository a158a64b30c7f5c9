/*
 * sahbvh_types.h -- value types of the closest-hit path, layout-compatible with the reference's
 * headers (x86-64 SysV; sizes/offsets asserted below and pinned against the compiled
 * reference by tests/test_layout.py):
 *
 *   Vec3      12 B   include/Custom/vec3.h:3-7
 *   Ray       24 B   include/Custom/ray.h:5-8
 *   Camera    64 B   include/Custom/camera.h:5-14
 *   Sphere    20 B   include/Custom/sphere.h:7-11   (center@0, radius@12, color@16)
 *   AABB      24 B   include/Custom/bvh.h:7-10
 *   BVHNode   56 B   include/Custom/bvh.h:12-18     (bounds@0,left@24,right@32,sphere@40,sphere_count@48)
 *   HitRecord 40 B   include/Custom/hit.h:8-14      (t@0,point@4,normal@16,hit_something@28,object@32)
 *
 * A caller that already includes the reference's own headers defines
 * SAHBVH_USE_REFERENCE_HEADERS before including this file; the types then come from
 * "Custom/hit.h" (which pulls vec3/ray/camera/sphere/bvh) and only the asserts remain.
 */
#ifndef SAHBVH_TYPES_H
#define SAHBVH_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef SAHBVH_USE_REFERENCE_HEADERS
#include "Custom/hit.h"
#else

typedef struct Vec3 {
    float x;
    float y;
    float z;
} Vec3;

typedef struct {
    Vec3 origin;
    Vec3 direction;
} Ray;

typedef struct {
    Vec3 position;
    Vec3 forward;
    Vec3 right;
    Vec3 up;
    float yaw;
    float pitch;
    float fov;
    int move;
} Camera;

/* The only thing the path needs from SDL is this 4-byte colour type (SDL_pixels.h:327-333). */
#ifndef SDL_pixels_h_
typedef struct SDL_Color {
    uint8_t r;
    uint8_t g;
    uint8_t b;
    uint8_t a;
} SDL_Color;
#endif

typedef struct {
    Vec3 center;
    float radius;
    SDL_Color color;
} Sphere;

typedef struct AABB {
    Vec3 min;
    Vec3 max;
} AABB;

typedef struct BVHNode {
    AABB bounds;
    struct BVHNode *left;
    struct BVHNode *right;
    Sphere *sphere;
    int sphere_count;
} BVHNode;

typedef struct {
    float t;
    Vec3 point;
    Vec3 normal;
    int hit_something;
    Sphere *object;
} HitRecord;

#endif /* SAHBVH_USE_REFERENCE_HEADERS */

/* include/Custom/constants.h:5-8 */
#define SAHBVH_EPSILON 0.000001f
#define SAHBVH_DEFAULT_WIDTH 800
#define SAHBVH_DEFAULT_HEIGHT 600
#define SAHBVH_MAX_DEPTH 5
#define SAHBVH_LEAF_DEPTH 40 /* bvh.c:131 */

#if defined(__cplusplus)
#define SAHBVH_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define SAHBVH_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
SAHBVH_STATIC_ASSERT(sizeof(Vec3) == 12, "Vec3 layout");
SAHBVH_STATIC_ASSERT(sizeof(Ray) == 24, "Ray layout");
SAHBVH_STATIC_ASSERT(sizeof(Camera) == 64, "Camera layout");
SAHBVH_STATIC_ASSERT(sizeof(Sphere) == 20 && offsetof(Sphere, radius) == 12 && offsetof(Sphere, color) == 16, "Sphere layout");
SAHBVH_STATIC_ASSERT(sizeof(AABB) == 24, "AABB layout");
SAHBVH_STATIC_ASSERT(sizeof(BVHNode) == 56 && offsetof(BVHNode, left) == 24 && offsetof(BVHNode, right) == 32 &&
                         offsetof(BVHNode, sphere) == 40 && offsetof(BVHNode, sphere_count) == 48, "BVHNode layout");
SAHBVH_STATIC_ASSERT(sizeof(HitRecord) == 40 && offsetof(HitRecord, point) == 4 && offsetof(HitRecord, normal) == 16 &&
                         offsetof(HitRecord, hit_something) == 28 && offsetof(HitRecord, object) == 32, "HitRecord layout");

#endif /* SAHBVH_TYPES_H */
