/*
 * traverse_core.cuh -- per-ray closest-hit arithmetic shared by every kernel in sahbvh_kernels.cu.
 *
 * Result contract (what makes the output identical to the reference):
 *   hit(ray) = best over all LEAVES whose reference box test passes (hit.c:49-82, true
 *   divisions, d==0 axis accepts everything) and whose reference sphere test passes
 *   (hit.c:19-39: float disc, DOUBLE t, t > EPSILON), best = smaller t, equal t -> larger
 *   post-build sphere index (hit.c:108 applied inductively, Q1).
 * Interior boxes only cull (a leaf's exact test passing implies every ancestor's passes, Q7),
 * so inner nodes use a cheaper CONSERVATIVE slab test:
 *
 *   t = fma(plane, inv_d, -o*inv_d -+ e)                1 FFMA per plane
 *   |t - t_ref| <= 4u|t_ref| + u|o*inv_d|               (u = 2^-24; the product plane*inv_d is
 *                                                        exact inside the FMA, only o*inv_d and
 *                                                        inv_d itself are rounded)
 *   e      =  2^-20 |o*inv_d| per axis, subtracted from the NEAR plane's addend and added to the
 *             FAR plane's (>= 15x the absolute part of the bound after its own rounding)
 *   pass  <=>  max(tnear, 0) <= min(tfar*(1+2^-20), best_s)
 *   best_s =  best_t*(1+2^-9) + 2*r_max/|d|              cull slack: the reference's float
 *             discriminant can move the computed t by <= 1.0e-3*|oc|/|d| and a float-noise
 *             "hit" of a far sphere can sit <= sqrt(2)*r/|d| before its own box entry.
 *   Axes with |d| < 1e-30 (incl. the reference's d==0 rule) or a non-finite o*inv_d are
 *   ignored (NaN operands drop out of FMNMX), i.e. they accept everything, like hit.c:54-56.
 *
 * The same source compiles for the host (tests/host_emul) so the logic can be checked against
 * the oracle without a GPU; that build is test-only and is not part of libsahbvh_b200.so.
 */
#ifndef SAHBVH_TRAVERSE_CORE_CUH
#define SAHBVH_TRAVERSE_CORE_CUH

#include <math.h>
#include <stdint.h>
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

#if defined(__CUDACC__)
#define SAH_HD __host__ __device__ __forceinline__
#else
#define SAH_HD static inline
#endif

#if defined(__CUDA_ARCH__)
/* explicit round-to-nearest, never contracted into FMA */
#define F_ADD(a, b) __fadd_rn((a), (b))
#define F_SUB(a, b) __fsub_rn((a), (b))
#define F_MUL(a, b) __fmul_rn((a), (b))
#define F_DIV(a, b) __fdiv_rn((a), (b))
#define F_SQRT(a) __fsqrt_rn((a))
#define D_SUB(a, b) __dsub_rn((a), (b))
#define D_DIV(a, b) __ddiv_rn((a), (b))
#define D_SQRT(a) __dsqrt_rn((a))
#define F_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
/* host build: compiled with -ffp-contract=off */
#define F_ADD(a, b) ((float)((a) + (b)))
#define F_SUB(a, b) ((float)((a) - (b)))
#define F_MUL(a, b) ((float)((a) * (b)))
#define F_DIV(a, b) ((float)((a) / (b)))
#define F_SQRT(a) ((float)sqrt((double)(a)))
#define D_SUB(a, b) ((double)(a) - (double)(b))
#define D_DIV(a, b) ((double)(a) / (double)(b))
#define D_SQRT(a) sqrt((double)(a))
#define F_FMA(a, b, c) fmaf((a), (b), (c))
#endif

#define SAH_EPSILON 0.000001f               /* constants.h:6 */
#define SAH_REL_BOX 1.00000095367431640625f /* 1 + 2^-20 */
#define SAH_REL_CULL 1.001953125f           /* 1 + 2^-9  */
#define SAH_ABS_SCALE 9.5367431640625e-07f  /* 2^-20 */
#define SAH_TINY_DIR 1e-30f
#define SAH_STACK_MAX 64
#define SAH_REF_NONE ((int32_t)0x80000000)

struct SahHit {  /* == gpu_hit */
    float t;
    int32_t prim;
    float px, py, pz;
    float nx, ny, nz;
};

struct SahRay {
    float ox, oy, oz;
    float dx, dy, dz;
    float ix, iy, iz;      /* 1/d; 2^100 on a zero/tiny axis (geometric mode); NaN on an ignored axis */
    float nx, ny, nz;      /* addend of the LO plane: -(o*inv)            [geometric: -(o+pad)*2^100] */
    float mx, my, mz;      /* addend of the HI plane: same as n*           [geometric: -(o-pad)*2^100] */
    float a;               /* d.d in the reference's order (hit.c:22) */
    float cull_abs;        /* 2*r_max/|d| */
};

/* per-tree constants the ray set-up needs */
struct SahScene {
    int n_spheres;
    float r_max;
    float lo[3], hi[3];    /* bounds of all spheres */
    int geometric_zero_axes; /* 1 unless the tree has sphere_count==0 leaves (see sah_axis_setup) */
};

struct SahCounters {
    unsigned long long inner, spheres, exact, max_stack, max_visits;
};

SAH_HD float sah_nan() {
#if defined(__CUDA_ARCH__)
    return __int_as_float(0x7fffffff);
#else
    return NAN;
#endif
}

/*
 * One axis of the conservative interior slab test:  t_lo = fma(lo, inv, n_lo), t_hi = fma(hi, inv, n_hi).
 *
 *  |d| >= 1e-30 : inv = 1/d, n_lo = n_hi = -(o*inv)  (see the error bound in the file header).
 *  |d| <  1e-30 (including the reference's d == 0 rule, hit.c:54-56, which accepts EVERY box on
 *    that axis): accepting everything makes such a ray walk the whole tree (the centre pixel of
 *    the default camera has dx == dy == 0).  Instead the axis is tested geometrically with a pad:
 *        accept  <=>  lo - pad <= o <= hi + pad,      inv = 2^100, n_lo = -(o+pad)*2^100, n_hi = -(o-pad)*2^100
 *    (products by 2^100 are exact, so the signs of t_lo/t_hi are exact).  This never culls a leaf
 *    the reference would report: with d_a == 0 the reference's float discriminant obeys
 *    disc <= 4a(r^2 - oc_a^2) + 68u*a*|oc|^2, so a hit needs |o_a - c_a| < r + 1.0e-3*|oc|, i.e.
 *    o_a inside the leaf's slab padded by 1.0e-3*|oc| <= pad = 2^-9 * (distance from the origin to
 *    the farthest corner of the scene bounds).  For 0 < |d_a| < 1e-30 the reference's own leaf
 *    test is exact containment, a subset.  The argument needs the tested sphere to lie inside
 *    every ancestor box, which fails for sphere_count == 0 leaves (they alias a sphere owned by
 *    another subtree): trees that contain such leaves keep the accept-everything rule.
 *  Non-finite addends (|o| ~ 1e8 and beyond): the axis is ignored (NaN drops out of FMNMX).
 */
SAH_HD void sah_axis_setup(float o, float d, float pad, bool geometric, float &inv, float &n_lo, float &n_hi, float &slack_src) {
    const float big = 1.2676506002282294e30f; /* 2^100 */
    slack_src = 0.0f;
    if (fabsf(d) >= SAH_TINY_DIR) {
        float i = F_DIV(1.0f, d);
        float n = -F_MUL(o, i);
        bool ok = fabsf(n) <= 3.0e38f;
        /* the absolute error u|o*inv| of this axis' plane distances is covered HERE, per axis, by moving the
         * near plane's addend down and the far plane's up by 2^-20 |o*inv| (lo is the near plane iff inv >= 0).
         * A single slack max_a|o_a*inv_a| on the final comparison would let one tiny direction component
         * (|d_a| ~ 1e-6 gives |o*inv| ~ 1e8) switch off the culling of the other two axes: such rays were
         * measured at 35x the average node count, and a handful of them set the tail of every launch. */
        float e = F_MUL(fabsf(n), SAH_ABS_SCALE);
        bool lo_near = i >= 0.0f;
        inv = ok ? i : sah_nan();
        n_lo = ok ? (lo_near ? F_SUB(n, e) : F_ADD(n, e)) : sah_nan();
        n_hi = ok ? (lo_near ? F_ADD(n, e) : F_SUB(n, e)) : sah_nan();
    } else {
        float a = -F_MUL(F_ADD(o, pad), big), b = -F_MUL(F_SUB(o, pad), big);
        bool ok = geometric && fabsf(a) <= 3.0e38f && fabsf(b) <= 3.0e38f;
        inv = ok ? big : sah_nan();
        n_lo = ok ? a : sah_nan();
        n_hi = ok ? b : sah_nan();
    }
}

SAH_HD void sah_ray_setup(SahRay &r, float ox, float oy, float oz, float dx, float dy, float dz, const SahScene &sc) {
    r.ox = ox; r.oy = oy; r.oz = oz;
    r.dx = dx; r.dy = dy; r.dz = dz;
    /* pad for geometric zero axes: 2^-9 * distance to the farthest scene corner (+ a relative crumb) */
    float ex = fmaxf(fabsf(ox - sc.lo[0]), fabsf(ox - sc.hi[0]));
    float ey = fmaxf(fabsf(oy - sc.lo[1]), fabsf(oy - sc.hi[1]));
    float ez = fmaxf(fabsf(oz - sc.lo[2]), fabsf(oz - sc.hi[2]));
    float far2 = ex * ex + ey * ey + ez * ez;
    float pad = sqrtf(far2) * 0.001953125f + sc.r_max * 0.001953125f;
    bool geo = sc.geometric_zero_axes != 0 && pad < 3.0e38f;
    float sx, sy, sz;
    sah_axis_setup(ox, dx, pad, geo, r.ix, r.nx, r.mx, sx);
    sah_axis_setup(oy, dy, pad, geo, r.iy, r.ny, r.my, sy);
    sah_axis_setup(oz, dz, pad, geo, r.iz, r.nz, r.mz, sz);
    r.a = F_ADD(F_ADD(F_MUL(dx, dx), F_MUL(dy, dy)), F_MUL(dz, dz));
    (void)sx; (void)sy; (void)sz;
    /* 2*r_max/|d|; |d| from the same float sum (exactness irrelevant: slack only) */
    float len = sqrtf(r.a);
    float per_d = (len > 0.0f) ? (2.0f * sc.r_max) / len : INFINITY;
    r.cull_abs = per_d;
}

/* conservative slab test of one box; returns entry distance through tnear */
SAH_HD bool sah_box_conservative(const SahRay &r, float lox, float loy, float loz, float hix, float hiy, float hiz,
                                 float best_s, float &tnear) {
    float t1x = F_FMA(lox, r.ix, r.nx), t2x = F_FMA(hix, r.ix, r.mx);
    float t1y = F_FMA(loy, r.iy, r.ny), t2y = F_FMA(hiy, r.iy, r.my);
    float t1z = F_FMA(loz, r.iz, r.nz), t2z = F_FMA(hiz, r.iz, r.mz);
    float tn = fmaxf(fmaxf(fminf(t1x, t2x), fminf(t1y, t2y)), fminf(t1z, t2z));
    float tf = fminf(fminf(fmaxf(t1x, t2x), fmaxf(t1y, t2y)), fmaxf(t1z, t2z));
    tn = fmaxf(tn, 0.0f);
    tf = fminf(F_MUL(tf, SAH_REL_BOX), best_s);
    tnear = tn;
    return tn <= tf;
}

/* the reference's slab test, bit for bit (hit.c:49-82) */
SAH_HD bool sah_box_exact(const SahRay &r, float lox, float loy, float loz, float hix, float hiy, float hiz) {
    const float inf = INFINITY;
    float t1x, t2x, t1y, t2y, t1z, t2z;
    if (r.dx == 0.0f) { t1x = -inf; t2x = inf; }
    else { t1x = F_DIV(F_SUB(lox, r.ox), r.dx); t2x = F_DIV(F_SUB(hix, r.ox), r.dx); }
    if (r.dy == 0.0f) { t1y = -inf; t2y = inf; }
    else { t1y = F_DIV(F_SUB(loy, r.oy), r.dy); t2y = F_DIV(F_SUB(hiy, r.oy), r.dy); }
    if (r.dz == 0.0f) { t1z = -inf; t2z = inf; }
    else { t1z = F_DIV(F_SUB(loz, r.oz), r.dz); t2z = F_DIV(F_SUB(hiz, r.oz), r.dz); }
    float tmin = fmaxf(fminf(t1x, t2x), fmaxf(fminf(t1y, t2y), fminf(t1z, t2z)));
    float tmax = fminf(fmaxf(t1x, t2x), fminf(fmaxf(t1y, t2y), fmaxf(t1z, t2z)));
    return tmax >= tmin && tmax > SAH_EPSILON;
}

/* sah_box_exact with a shortcut for the common case (device only; the RESULT is sah_box_exact's, always).
 * The six rounded subtractions q = plane - o are the reference's own.  Each quotient t = fl(q / d) is then
 * approximated by a = fl(q * r), r = rcp.approx(d):  |r*d - 1| <= 2^-23, one rounding in a, one in t, so
 * a / t is in 1 +- 2^-22 whenever a and t are normal numbers.  With N = max_a min(a1,a2), F = min_a max(a1,a2):
 *     F (1 - 2^-18) >= max(N, 0)   and   F (1 - 2^-18) > EPSILON
 * imply  tmax >= tmin  and  tmax > EPSILON  for the reference's own t's (the margin is 16x the error bound; values
 * below EPSILON -- the only place denormals could matter -- never take the shortcut, a NaN product fails both
 * comparisons and an infinite one is excluded explicitly).  Anything else, including every zero direction component (r = inf), runs the
 * six true divisions.  A would-be winner's chord through its sphere's box is ~1e3 times wider than the margin, so
 * about 99 % of the calls stop after 3 MUFU + 6 FMUL instead of six IEEE division sequences. */
#if defined(__CUDA_ARCH__)
SAH_HD bool sah_box_exact_fast(const SahRay &r, float lox, float loy, float loz, float hix, float hiy, float hiz) {
    float rx, ry, rz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(r.dx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(r.dy));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rz) : "f"(r.dz));
    const float a1x = F_MUL(F_SUB(lox, r.ox), rx), a2x = F_MUL(F_SUB(hix, r.ox), rx);
    const float a1y = F_MUL(F_SUB(loy, r.oy), ry), a2y = F_MUL(F_SUB(hiy, r.oy), ry);
    const float a1z = F_MUL(F_SUB(loz, r.oz), rz), a2z = F_MUL(F_SUB(hiz, r.oz), rz);
    const float n = fmaxf(fmaxf(fminf(a1x, a2x), fminf(a1y, a2y)), fmaxf(fminf(a1z, a2z), 0.0f));
    const float f = F_MUL(fminf(fminf(fmaxf(a1x, a2x), fmaxf(a1y, a2y)), fmaxf(a1z, a2z)), 0.999996185302734375f);
    const bool finite = fabsf(rx) <= 3.0e38f && fabsf(ry) <= 3.0e38f && fabsf(rz) <= 3.0e38f;   /* no zero / denormal direction axis */
    /* f < 3e38: an overflowed product (inf >= inf would pass) never takes the shortcut either */
    if (finite && f >= n && f > SAH_EPSILON && f < 3.0e38f) return true;
    return sah_box_exact(r, lox, loy, loz, hix, hiy, hiz);
}
#else
SAH_HD bool sah_box_exact_fast(const SahRay &r, float lox, float loy, float loz, float hix, float hiy, float hiz) {
    return sah_box_exact(r, lox, loy, loz, hix, hiy, hiz);
}
#endif

/* the reference's sphere test up to t (hit.c:21-29); returns true and t when accepted */
SAH_HD bool sah_sphere_t(const SahRay &r, float cx, float cy, float cz, float rad, float &t_out) {
    float ocx = F_SUB(r.ox, cx), ocy = F_SUB(r.oy, cy), ocz = F_SUB(r.oz, cz);
    float b = F_MUL(2.0f, F_ADD(F_ADD(F_MUL(ocx, r.dx), F_MUL(ocy, r.dy)), F_MUL(ocz, r.dz)));
    float c = F_SUB(F_ADD(F_ADD(F_MUL(ocx, ocx), F_MUL(ocy, ocy)), F_MUL(ocz, ocz)), F_MUL(rad, rad));
    float disc = F_SUB(F_MUL(b, b), F_MUL(F_MUL(4.0f, r.a), c));
    if (!(disc > 0.0f)) return false;
    double num = D_SUB((double)(-b), D_SQRT((double)disc));
    double den = (double)F_MUL(2.0f, r.a);
    float t = (float)D_DIV(num, den);
    if (!(t > SAH_EPSILON)) return false;
    t_out = t;
    return true;
}

/* point and normal of an accepted hit (hit.c:32-33, vec3.c:21-24) */
SAH_HD void sah_finish_hit(const SahRay &r, float t, int32_t prim, float cx, float cy, float cz, SahHit &h) {
    float px = F_ADD(r.ox, F_MUL(r.dx, t));
    float py = F_ADD(r.oy, F_MUL(r.dy, t));
    float pz = F_ADD(r.oz, F_MUL(r.dz, t));
    float vx = F_SUB(px, cx), vy = F_SUB(py, cy), vz = F_SUB(pz, cz);
    float len = F_SQRT(F_ADD(F_ADD(F_MUL(vx, vx), F_MUL(vy, vy)), F_MUL(vz, vz)));
    h.t = t;
    h.prim = prim;
    h.px = px; h.py = py; h.pz = pz;
    if (len != 0.0f) { h.nx = F_DIV(vx, len); h.ny = F_DIV(vy, len); h.nz = F_DIV(vz, len); }
    else { h.nx = 0.0f; h.ny = 0.0f; h.nz = 0.0f; }
}

SAH_HD void sah_miss(SahHit &h) {
    h.t = 0.0f; h.prim = -1;
    h.px = h.py = h.pz = 0.0f;
    h.nx = h.ny = h.nz = 0.0f;
}

struct SahF4 { float x, y, z, w; };

#if defined(__CUDA_ARCH__)
SAH_HD SahF4 sah_ld4(const float *p) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    SahF4 r = {v.x, v.y, v.z, v.w};
    return r;
}
SAH_HD int32_t sah_f2i(float f) { return __float_as_int(f); }
#else
SAH_HD SahF4 sah_ld4(const float *p) {
    SahF4 r = {p[0], p[1], p[2], p[3]};
    return r;
}
SAH_HD int32_t sah_f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
#endif

/* leaf candidate: sphere test first (cheap reject), exact box test only for a would-be winner */
template <bool COUNT>
SAH_HD void sah_leaf(const SahRay &r, const float *__restrict__ spheres4, int n_spheres, int32_t ref,
                     float lox, float loy, float loz, float hix, float hiy, float hiz,
                     float &best_t, int32_t &best_prim, float &best_s, SahCounters *cnt) {
    int32_t idx = ~ref;
    if ((uint32_t)idx >= (uint32_t)n_spheres) return;   /* one-past-the-end leaf (Q3) / absent child */
    SahF4 s = sah_ld4(spheres4 + (size_t)idx * 4);
    if (COUNT) cnt->spheres++;
    float t;
    if (!sah_sphere_t(r, s.x, s.y, s.z, s.w, t)) return;
    bool better = (best_prim < 0) || (t < best_t) || (t == best_t && idx > best_prim);
    if (!better) return;
    if (COUNT) cnt->exact++;
    if (!sah_box_exact_fast(r, lox, loy, loz, hix, hiy, hiz)) return;
    best_t = t;
    best_prim = idx;
    best_s = F_FMA(t, SAH_REL_CULL, r.cull_abs);
}

/*
 * One ray through the flattened tree.  `stack` is caller-provided storage of SAH_STACK_MAX ints
 * with element stride `stride` (1 for a private array, blockDim for a shared-memory column).
 */
#define SAH_NODE_DONE (-1)

/* One node visit of one ray: the node's two child records are already in registers.
 * Returns the next inner node to visit, or SAH_NODE_DONE when the ray's stack is empty. */
template <bool COUNT>
SAH_HD int32_t sah_visit(const SahRay &r, const SahF4 &l0, const SahF4 &l1, const SahF4 &r0, const SahF4 &r1,
                         const float *__restrict__ spheres4, int n_spheres, int32_t *stack, int stride, int &sp,
                         float &best_t, int32_t &best_prim, float &best_s, SahCounters *cnt) {
    if (COUNT) cnt->inner++;
    int32_t lref = sah_f2i(l1.z), rref = sah_f2i(r1.z);
    float tl, tr;
    bool hl = sah_box_conservative(r, l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, best_s, tl);
    bool hr = sah_box_conservative(r, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, best_s, tr);
    if (hl && lref < 0) {
        sah_leaf<COUNT>(r, spheres4, n_spheres, lref, l0.x, l0.y, l0.z, l0.w, l1.x, l1.y,
                        best_t, best_prim, best_s, cnt);
        hl = false;
    }
    if (hr && rref < 0) {
        /* re-check against a best_s the left leaf may just have tightened */
        if (tr <= best_s)
            sah_leaf<COUNT>(r, spheres4, n_spheres, rref, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y,
                            best_t, best_prim, best_s, cnt);
        hr = false;
    }
    if (hl && hr) {
        bool left_first = tl <= tr;
        stack[sp * stride] = left_first ? rref : lref;
        sp++;
        if (COUNT && (unsigned long long)sp > cnt->max_stack) cnt->max_stack = (unsigned long long)sp;
        return left_first ? lref : rref;
    }
    if (hl) return lref;
    if (hr) return rref;
    if (sp == 0) return SAH_NODE_DONE;
    sp--;
    return stack[sp * stride];
}

template <bool COUNT>
SAH_HD void sah_trace_one(const float *__restrict__ nodes, const float *__restrict__ spheres4, int n_spheres,
                          const SahRay &r, int32_t *stack, int stride, SahHit &out, SahCounters *cnt) {
    float best_t = 0.0f, best_s = INFINITY;
    int32_t best_prim = -1;
    int sp = 0;
    int32_t node = 0;
    while (node != SAH_NODE_DONE) {
        const float *w = nodes + (size_t)node * 16;
        /* two 32-byte child records: {lo.xyz, hi.x} {hi.y, hi.z, ref, 0} */
        SahF4 l0 = sah_ld4(w), l1 = sah_ld4(w + 4), r0 = sah_ld4(w + 8), r1 = sah_ld4(w + 12);
        node = sah_visit<COUNT>(r, l0, l1, r0, r1, spheres4, n_spheres, stack, stride, sp, best_t, best_prim, best_s, cnt);
    }
    if (best_prim >= 0) {
        SahF4 s = sah_ld4(spheres4 + (size_t)best_prim * 4);
        sah_finish_hit(r, best_t, best_prim, s.x, s.y, s.z, out);
    } else {
        sah_miss(out);
    }
}

/* ---------------------------------------------------------------------------------------------
 * 4-wide quantised nodes (layout and containment guarantee: wide.c).
 *
 * Decoded plane of byte q:  p + (1024 + q) * s   (p, s from the node; s a power of two).
 * Per node and axis:   s_inv = s * inv                 exact (power of two)
 *                      A_lo  = fma(p, inv, n_lo),  A_hi = fma(p, inv, n_hi)
 * per child plane:     t     = fma(1024 + q, s_inv, A)  1024 + q exact in fp16 and fp32
 * Against the true t of the DECODED plane the error is
 *     <= 4u|t| + u|o*inv|          (as for sah_box_conservative: covered by SAH_REL_BOX and the per-axis addend slack)
 *      + u|A| + u|(1024+q)*s_inv|  (the intermediate A can be as large as |t| + 1279|s_inv|)
 *     <= ... + 2^-13 |s_inv|,
 * i.e. 2^-13 of a grid cell in t-space; the decoded box contains the true child box with a
 * margin of 2^-9 cell on every side, so a box the reference accepts is never rejected.  Zero /
 * tiny direction axes use the geometric mode of sah_axis_setup unchanged (inv = 2^100 > 0, two
 * addends); an axis whose p is NaN (not quantisable) yields NaN planes, which drop out of
 * FMNMX: the axis accepts everything.
 * The near/far plane of an axis is picked by the sign of inv once per node (monotone rounding
 * keeps t_near <= t_far for qlo <= qhi), which is equivalent to the min/max form.
 * --------------------------------------------------------------------------------------------- */
#if defined(__CUDA_ARCH__)
#define SAH_UNROLL _Pragma("unroll")
#else
#define SAH_UNROLL
#endif
#define SAH_WIDE_SPECIAL 0x40000000u
#define SAH_WIDE_STACK_MAX 192
#define SAH_KEY_NONE 0xffffffffu

/* traversal stack: the first `fast_depth` entries in `fast` (element stride `stride`: a private
 * array on the host, a bank-conflict-free shared-memory column on the device), overflow in `slow` */
struct SahStack {
    int32_t *fast;
    int stride;
    int fast_depth;
    int32_t *slow;
};
SAH_HD void sah_push(const SahStack &k, int &sp, int32_t v) {
    if (sp < k.fast_depth) k.fast[sp * k.stride] = v;
    else k.slow[sp - k.fast_depth] = v;
    sp++;
}
SAH_HD int32_t sah_pop(const SahStack &k, int &sp) {
    sp--;
    return sp < k.fast_depth ? k.fast[sp * k.stride] : k.slow[sp - k.fast_depth];
}

#if defined(__CUDA_ARCH__)
SAH_HD float sah_u2f(uint32_t u) { return __uint_as_float(u); }
SAH_HD uint32_t sah_f2u(float f) { return __float_as_uint(f); }
#else
SAH_HD float sah_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
SAH_HD uint32_t sah_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#endif
/* the four bytes of w as floats 1024 + b.  Device: two PRMTs build the fp16 pairs 0x64bb
 * (= 1024 + b exactly) and HADD2.F32 widens them on the FMA pipe -- I2F.U8 runs on the
 * quarter-rate XU pipe, which the ncu profile showed as a co-bottleneck of the wide kernel. */
#if defined(__CUDA_ARCH__)
SAH_HD void sah_bytes1024(uint32_t w, float (&f)[4]) {
    const uint32_t lo = __byte_perm(w, 0x64646464u, 0x4140), hi = __byte_perm(w, 0x64646464u, 0x4342);
    const __half2 hl = *reinterpret_cast<const __half2 *>(&lo), hh = *reinterpret_cast<const __half2 *>(&hi);
    f[0] = __low2float(hl); f[1] = __high2float(hl); f[2] = __low2float(hh); f[3] = __high2float(hh);
}
#else
SAH_HD void sah_bytes1024(uint32_t w, float (&f)[4]) {
    for (int k = 0; k < 4; k++) f[k] = 1024.0f + (float)((w >> (8 * k)) & 0xffu);
}
#endif

/* leaf of a wide node: regular (box recomputed from the sphere, bvh.c:26-35) or special (table) */
template <bool COUNT>
SAH_HD void sah_wide_leaf(const SahRay &r, const float *__restrict__ spheres4, const float *__restrict__ special,
                          int n_spheres, int n_special, int32_t ref,
                          float &best_t, int32_t &best_prim, float &best_s, SahCounters *cnt) {
    const uint32_t code = (uint32_t)~ref;
    int32_t idx;
    float lox, loy, loz, hix, hiy, hiz;
    SahF4 s;
    if (code & SAH_WIDE_SPECIAL) {
        const uint32_t k = code & ~SAH_WIDE_SPECIAL;
        if (k >= (uint32_t)n_special) return;
        const SahF4 a = sah_ld4(special + (size_t)k * 8), b = sah_ld4(special + (size_t)k * 8 + 4);
        idx = sah_f2i(b.z);
        if ((uint32_t)idx >= (uint32_t)n_spheres) return;
        lox = a.x; loy = a.y; loz = a.z; hix = a.w; hiy = b.x; hiz = b.y;
        s = sah_ld4(spheres4 + (size_t)idx * 4);
    } else {
        idx = (int32_t)code;
        if ((uint32_t)idx >= (uint32_t)n_spheres) return;
        s = sah_ld4(spheres4 + (size_t)idx * 4);
        lox = F_SUB(s.x, s.w); loy = F_SUB(s.y, s.w); loz = F_SUB(s.z, s.w);
        hix = F_ADD(s.x, s.w); hiy = F_ADD(s.y, s.w); hiz = F_ADD(s.z, s.w);
    }
    if (COUNT) cnt->spheres++;
    float t;
    if (!sah_sphere_t(r, s.x, s.y, s.z, s.w, t)) return;
    const bool better = (best_prim < 0) || (t < best_t) || (t == best_t && idx > best_prim);
    if (!better) return;
    if (COUNT) cnt->exact++;
    if (!sah_box_exact_fast(r, lox, loy, loz, hix, hiy, hiz)) return;
    best_t = t;
    best_prim = idx;
    best_s = F_FMA(t, SAH_REL_CULL, r.cull_abs);
}

#define SAH_CE(ka, ra, kb, rb)                                   \
    do {                                                         \
        const bool sw_ = (kb) < (ka);                            \
        const uint32_t k0_ = sw_ ? (kb) : (ka), k1_ = sw_ ? (ka) : (kb); \
        const int32_t r0_ = sw_ ? (rb) : (ra), r1_ = sw_ ? (ra) : (rb);  \
        (ka) = k0_; (kb) = k1_; (ra) = r0_; (rb) = r1_;          \
    } while (0)

/* Put a ray set up by sah_ray_setup into "wide form": n* becomes the addend of the NEAR plane of
 * its axis and m* the addend of the FAR plane (they differ only on geometric zero axes, where
 * inv = 2^100 > 0 and near = lo already).  sah_wide_visit expects this form. */
SAH_HD void sah_ray_to_wide(SahRay &r) {
    if (!(r.ix >= 0.0f)) { const float t = r.nx; r.nx = r.mx; r.mx = t; }
    if (!(r.iy >= 0.0f)) { const float t = r.ny; r.ny = r.my; r.my = t; }
    if (!(r.iz >= 0.0f)) { const float t = r.nz; r.nz = r.mz; r.mz = t; }
}

/* One wide-node visit of one ray (in wide form); w[] = the 16 words of node `self`.  Returns the
 * next wide node or SAH_NODE_DONE when the ray's stack is empty.
 *
 * Leaf children that pass the box test are NOT tested here: their refs are parked in the ray's two
 * pending slots (pend0/pend1, SAH_REF_NONE = free) and the caller runs sah_wide_leaf on them
 * later, for many rays at once (a warp runs the divergent sphere code with several lanes active
 * instead of one).  Deferral is safe because the result is a best-over-leaves: it only delays the
 * tightening of best_s.  If both slots are taken when leaf child c needs one, the visit sets
 * stall_at = c and returns `self` unchanged; the caller drains the slots and repeats the visit,
 * which then parks only children >= c (so a node with 3 or 4 hit leaves makes progress).
 * stall_at is -1 on entry and exit otherwise. */
template <bool COUNT>
SAH_HD int32_t sah_wide_visit(const SahRay &r, const uint32_t (&w)[16], int32_t self, const SahStack &stk, int &sp,
                              float best_s, int32_t &pend0, int32_t &pend1, int &stall_at, SahCounters *cnt) {
    const float sx = F_MUL(sah_u2f(w[3]), r.ix), sy = F_MUL(sah_u2f(w[14]), r.iy), sz = F_MUL(sah_u2f(w[15]), r.iz);
    const float px = sah_u2f(w[0]), py = sah_u2f(w[1]), pz = sah_u2f(w[2]);
    const float anx = F_FMA(px, r.ix, r.nx), afx = F_FMA(px, r.ix, r.mx);
    const float any_ = F_FMA(py, r.iy, r.ny), afy = F_FMA(py, r.iy, r.my);
    const float anz = F_FMA(pz, r.iz, r.nz), afz = F_FMA(pz, r.iz, r.mz);
    const bool posx = r.ix >= 0.0f, posy = r.iy >= 0.0f, posz = r.iz >= 0.0f;
    const uint32_t qnx = posx ? w[4] : w[7], qfx = posx ? w[7] : w[4];
    const uint32_t qny = posy ? w[5] : w[8], qfy = posy ? w[8] : w[5];
    const uint32_t qnz = posz ? w[6] : w[9], qfz = posz ? w[9] : w[6];

    float fnx[4], ffx[4], fny[4], ffy[4], fnz[4], ffz[4];
    sah_bytes1024(qnx, fnx); sah_bytes1024(qfx, ffx);
    sah_bytes1024(qny, fny); sah_bytes1024(qfy, ffy);
    sah_bytes1024(qnz, fnz); sah_bytes1024(qfz, ffz);
    float tn[4];
    bool hit[4];
    int32_t ref[4];
    int32_t leaf_bits = 0; /* sign bit set <=> some hit child is a leaf */
SAH_UNROLL
    for (int c = 0; c < 4; c++) {
        ref[c] = (int32_t)w[10 + c];
        const float tnx = F_FMA(fnx[c], sx, anx), tfx = F_FMA(ffx[c], sx, afx);
        const float tny = F_FMA(fny[c], sy, any_), tfy = F_FMA(ffy[c], sy, afy);
        const float tnz = F_FMA(fnz[c], sz, anz), tfz = F_FMA(ffz[c], sz, afz);
        tn[c] = fmaxf(fmaxf(fmaxf(tnx, tny), tnz), 0.0f);
        const float f_ = fminf(F_MUL(fminf(fminf(tfx, tfy), tfz), SAH_REL_BOX), best_s);
        /* an absent child (SAH_REF_NONE, inverted grid box) can pass only when every axis is
         * ignored; it then looks like a special leaf with an out-of-range index and is skipped */
        hit[c] = tn[c] <= f_;
        leaf_bits |= hit[c] ? ref[c] : 0;
    }
    const int from = stall_at;
    stall_at = -1;
    if (leaf_bits < 0) {
        int over = -1;
SAH_UNROLL
        for (int c = 0; c < 4; c++)
            if (hit[c] && ref[c] < 0 && c >= from && over < 0) {
                if (pend0 == SAH_REF_NONE) pend0 = ref[c];
                else if (pend1 == SAH_REF_NONE) pend1 = ref[c];
                else over = c;
            }
        stall_at = over;
        if (over >= 0) return self;
    }
    if (COUNT) cnt->inner++;
    /* inner children, nearest first: tn >= 0, so its bit pattern orders like its value (the order among equal
     * entry distances only affects the visiting order, never the result) */
    uint32_t k0, k1, k2, k3;
    {
        uint32_t k[4];
SAH_UNROLL
        for (int c = 0; c < 4; c++) {
            const bool inner_hit = hit[c] & (ref[c] >= 0);
            k[c] = inner_hit ? sah_f2u(tn[c]) : SAH_KEY_NONE;
        }
        k0 = k[0]; k1 = k[1]; k2 = k[2]; k3 = k[3];
    }
    int32_t r0 = ref[0], r1 = ref[1], r2 = ref[2], r3 = ref[3];
    SAH_CE(k0, r0, k1, r1);
    SAH_CE(k2, r2, k3, r3);
    SAH_CE(k0, r0, k2, r2);   /* k0 = nearest */
    SAH_CE(k1, r1, k3, r3);   /* k3 = farthest; k1, k2 stay in slot-pair order (only matters with 3+ hits) */
    if (k0 == SAH_KEY_NONE) {
        if (sp == 0) return SAH_NODE_DONE;
        return sah_pop(stk, sp);
    }
    if (sp + 3 <= stk.fast_depth) {
        /* common case, branch free: the up to three farther children go to sp.. in far-to-near
         * order; a slot written for an absent child is overwritten by the next store or push */
        int32_t *top = stk.fast + sp * stk.stride;
        int a = (k3 != SAH_KEY_NONE) ? 1 : 0;
        top[0] = r3;
        top[a * stk.stride] = r2;
        a += (k2 != SAH_KEY_NONE) ? 1 : 0;
        top[a * stk.stride] = r1;
        a += (k1 != SAH_KEY_NONE) ? 1 : 0;
        sp += a;
    } else {
        if (k3 != SAH_KEY_NONE) sah_push(stk, sp, r3);
        if (k2 != SAH_KEY_NONE) sah_push(stk, sp, r2);
        if (k1 != SAH_KEY_NONE) sah_push(stk, sp, r1);
    }
    if (COUNT && (unsigned long long)sp > cnt->max_stack) cnt->max_stack = (unsigned long long)sp;
    return r0;
}

/* run the deferred leaf tests of one ray and free its pending slots */
template <bool COUNT>
SAH_HD void sah_wide_drain(const SahRay &r, const float *__restrict__ spheres4, const float *__restrict__ special,
                           int n_spheres, int n_special, int32_t &pend0, int32_t &pend1,
                           float &best_t, int32_t &best_prim, float &best_s, SahCounters *cnt) {
    if (pend0 != SAH_REF_NONE) sah_wide_leaf<COUNT>(r, spheres4, special, n_spheres, n_special, pend0, best_t, best_prim, best_s, cnt);
    if (pend1 != SAH_REF_NONE) sah_wide_leaf<COUNT>(r, spheres4, special, n_spheres, n_special, pend1, best_t, best_prim, best_s, cnt);
    pend0 = SAH_REF_NONE;
    pend1 = SAH_REF_NONE;
}

#if !defined(__CUDA_ARCH__)
/* host-side loop over sah_wide_visit (tests/host_emul only; the device loop lives in the kernel).
 * `drain_every` = how many visits may pass before the pending leaves are tested: 1 tests them
 * right away, larger values exercise the deferral (and the stall path) the kernel relies on. */
template <bool COUNT>
SAH_HD void sah_wide_trace_one(const uint32_t *nodes, const float *spheres4, const float *special,
                               int n_spheres, int n_special, const SahRay &r, const SahStack &stk, int drain_every,
                               SahHit &out, SahCounters *cnt) {
    float best_t = 0.0f, best_s = INFINITY;
    int32_t best_prim = -1;
    int sp = 0;
    int32_t node = 0, pend0 = SAH_REF_NONE, pend1 = SAH_REF_NONE;
    int stall_at = -1;
    SahRay rw = r;
    sah_ray_to_wide(rw);
    long long step = 0;
    while (node != SAH_NODE_DONE) {
        uint32_t w[16];
        memcpy(w, nodes + (size_t)node * 16, 64);
        node = sah_wide_visit<COUNT>(rw, w, node, stk, sp, best_s, pend0, pend1, stall_at, cnt);
        if (stall_at >= 0 || ++step % drain_every == 0)
            sah_wide_drain<COUNT>(rw, spheres4, special, n_spheres, n_special, pend0, pend1, best_t, best_prim, best_s, cnt);
    }
    sah_wide_drain<COUNT>(rw, spheres4, special, n_spheres, n_special, pend0, pend1, best_t, best_prim, best_s, cnt);
    if (best_prim >= 0) {
        SahF4 s = sah_ld4(spheres4 + (size_t)best_prim * 4);
        sah_finish_hit(r, best_t, best_prim, s.x, s.y, s.z, out);
    } else {
        sah_miss(out);
    }
}
#endif

/* primary-ray constants computed once on the host (ray.c:18-24) */
struct SahCamera {
    float px, py, pz;          /* camera->position */
    float fx, fy, fz;          /* camera->forward */
    float hx, hy, hz;          /* horizontal = right * (2*half_w) */
    float vx, vy, vz;          /* vertical   = up * (2*half_h) */
    float aspect;              /* (float)W/(float)H */
    float fw, fh;              /* (float)W, (float)H */
};

/* main.c:362-363: the two pixel coordinates the camera ray is built from (v already negated, main.c:365) */
SAH_HD float sah_pixel_u(const SahCamera &c, int x) { return F_MUL(F_SUB(F_DIV((float)x, c.fw), 0.5f), c.aspect); }
SAH_HD float sah_pixel_v(const SahCamera &c, int y) { return -F_SUB(F_DIV((float)y, c.fh), 0.5f); }

/* ray.c:26-31 for (u, v): direction only (origin = position) */
SAH_HD void sah_primary_dir_uv(const SahCamera &c, float u, float v, float &dx, float &dy, float &dz) {
    float ax = F_ADD(c.fx, F_MUL(c.hx, u)), ay = F_ADD(c.fy, F_MUL(c.hy, u)), az = F_ADD(c.fz, F_MUL(c.hz, u));
    ax = F_ADD(ax, F_MUL(c.vx, v)); ay = F_ADD(ay, F_MUL(c.vy, v)); az = F_ADD(az, F_MUL(c.vz, v));
    float len = F_SQRT(F_ADD(F_ADD(F_MUL(ax, ax), F_MUL(ay, ay)), F_MUL(az, az)));
    if (len != 0.0f) { dx = F_DIV(ax, len); dy = F_DIV(ay, len); dz = F_DIV(az, len); }
    else { dx = 0.0f; dy = 0.0f; dz = 0.0f; }
}

/* main.c:362-365 + ray.c:26-31 for pixel (x,y): direction only (origin = position) */
SAH_HD void sah_primary_dir(const SahCamera &c, int x, int y, float &dx, float &dy, float &dz) {
    sah_primary_dir_uv(c, sah_pixel_u(c, x), sah_pixel_v(c, y), dx, dy, dz);
}

/* ---------------------------------------------------------------------------------------------
 * Packet (shared-origin) culling: one conservative box test for ALL rays of a pixel tile.
 *
 * The tile's rays share the origin o and their directions lie in the per-axis intervals
 * [dmin_a, dmax_a] (exact min/max over the tile's actual float directions).  For a box,
 *     LB_a <= near_a(ray)   and   UB_a >= far_a(ray)    for every ray of the tile,
 * hence  max_a LB_a > min_a UB_a  ==>  every ray has tmin > tmax  ==>  the reference's slab
 * test (hit.c:49-82) fails for every ray: the box is culled for the whole tile.  Per axis:
 *   all d >=  1e-30 : q/d is monotone in d, so near/far extremes sit at the interval ends:
 *                     LB = min(qn*ia, qn*ib), UB = max(qf*ia, qf*ib), ia = 1/dmax, ib = 1/dmin,
 *                     qn = lo - o, qf = hi - o  (the reference's own first rounding step)
 *   all d <= -1e-30 : same with the planes swapped
 *   straddling 0    : (may contain rays with d == 0, for which the reference accepts any slab)
 *                     padded geometric test as in sah_axis_setup: origin inside [lo-pad, hi+pad]
 *                     -> no constraint; below -> LB = (lo-pad-o)/dmax; above -> LB = (hi+pad-o)/dmin
 *   mode 3          : axis ignored (trees with sphere_count==0 leaves, non-finite input)
 * Products carry <= 2 ulp relative error against the reference's true divisions; the caller
 * compares with a 2^-19 relative slack.
 */
struct SahFrustumAxis {
    int mode;      /* 0 positive, 1 negative, 2 straddling, 3 ignored */
    float o;
    float ia, ib;  /* inverse-direction bounds (see above) */
    float pad;
};

SAH_HD void sah_frustum_axis_setup(SahFrustumAxis &A, float o, float dmin, float dmax, float pad, bool geometric) {
    A.o = o;
    A.pad = pad;
    A.ia = 0.0f;
    A.ib = 0.0f;
    if (!(dmin <= dmax) || !(fabsf(o) <= 3.0e38f)) { A.mode = 3; return; }
    if (dmin >= SAH_TINY_DIR) {
        A.mode = 0; A.ia = F_DIV(1.0f, dmax); A.ib = F_DIV(1.0f, dmin);
    } else if (dmax <= -SAH_TINY_DIR) {
        A.mode = 1; A.ia = F_DIV(1.0f, dmax); A.ib = F_DIV(1.0f, dmin);   /* ia <= ib in both modes (sah_frustum_axis_simple) */
    } else if (geometric && pad < 3.0e38f) {
        A.mode = 2;
        A.ia = (dmax >= SAH_TINY_DIR) ? F_DIV(1.0f, dmax) : INFINITY;     /* fastest approach from below */
        A.ib = (dmin <= -SAH_TINY_DIR) ? F_DIV(1.0f, dmin) : -INFINITY;   /* fastest approach from above */
    } else {
        A.mode = 3;
    }
}

SAH_HD void sah_frustum_axis(const SahFrustumAxis &A, float lo, float hi, float &LB, float &UB) {
    if (A.mode <= 1) {
        float qn = F_SUB(A.mode == 0 ? lo : hi, A.o), qf = F_SUB(A.mode == 0 ? hi : lo, A.o);
        LB = fminf(F_MUL(qn, A.ia), F_MUL(qn, A.ib));
        UB = fmaxf(F_MUL(qf, A.ia), F_MUL(qf, A.ib));
    } else if (A.mode == 2) {
        float qlo = F_SUB(F_SUB(lo, A.pad), A.o), qhi = F_SUB(F_ADD(hi, A.pad), A.o);
        UB = INFINITY;
        if (qlo > 0.0f) LB = F_MUL(qlo, A.ia);
        else if (qhi < 0.0f) LB = F_MUL(qhi, A.ib);
        else LB = -INFINITY;
        if (!(LB == LB)) LB = -INFINITY;   /* inf*0 cannot occur (q != 0 here); belt and braces */
    } else {
        LB = -INFINITY;
        UB = INFINITY;
    }
}

#define SAH_REL_PKT 1.0000019073486328125f /* 1 + 2^-19 */

/* true if the box may be hit by some ray of the tile with entry <= best_s; tnear = packet entry bound */
SAH_HD bool sah_frustum_box(const SahFrustumAxis &X, const SahFrustumAxis &Y, const SahFrustumAxis &Z,
                            float lox, float loy, float loz, float hix, float hiy, float hiz, float best_s, float &tnear) {
    float lbx, ubx, lby, uby, lbz, ubz;
    sah_frustum_axis(X, lox, hix, lbx, ubx);
    sah_frustum_axis(Y, loy, hiy, lby, uby);
    sah_frustum_axis(Z, loz, hiz, lbz, ubz);
    float tn = fmaxf(fmaxf(fmaxf(lbx, lby), lbz), 0.0f);
    float tf = fminf(fminf(ubx, uby), ubz);
    tnear = tn;
    return tn <= F_FMA(tf, SAH_REL_PKT, 1e-37f) && tn <= best_s;
}

/* sah_frustum_box for a tile whose three axes are all in mode 0 or 1 (the common case: no direction interval
 * straddles zero).  Same arithmetic as sah_frustum_axis for those modes, without its per-axis mode branches. */
SAH_HD bool sah_frustum_simple(const SahFrustumAxis &X, const SahFrustumAxis &Y, const SahFrustumAxis &Z) {
    return X.mode <= 1 && Y.mode <= 1 && Z.mode <= 1;
}

SAH_HD void sah_frustum_axis_simple(const SahFrustumAxis &A, float lo, float hi, float &LB, float &UB) {
    const bool pos = A.mode == 0;
    const float qn = F_SUB(pos ? lo : hi, A.o), qf = F_SUB(pos ? hi : lo, A.o);
#if defined(SAH_FRUSTUM_SELECT)
    /* ia <= ib in both modes (mode 0: 1/dmax <= 1/dmin, both positive; mode 1: 1/dmax <= 1/dmin, both negative), and
     * rounding is monotone, so min(fl(q*ia), fl(q*ib)) = fl(q * (q >= 0 ? ia : ib)) and max(...) = fl(q * (q >= 0 ? ib : ia)) */
    LB = F_MUL(qn, qn >= 0.0f ? A.ia : A.ib);
    UB = F_MUL(qf, qf >= 0.0f ? A.ib : A.ia);
#else
    LB = fminf(F_MUL(qn, A.ia), F_MUL(qn, A.ib));
    UB = fmaxf(F_MUL(qf, A.ia), F_MUL(qf, A.ib));
#endif
}

SAH_HD bool sah_frustum_box_simple(const SahFrustumAxis &X, const SahFrustumAxis &Y, const SahFrustumAxis &Z,
                                   float lox, float loy, float loz, float hix, float hiy, float hiz, float best_s, float &tnear) {
    float lbx, ubx, lby, uby, lbz, ubz;
    sah_frustum_axis_simple(X, lox, hix, lbx, ubx);
    sah_frustum_axis_simple(Y, loy, hiy, lby, uby);
    sah_frustum_axis_simple(Z, loz, hiz, lbz, ubz);
    float tn = fmaxf(fmaxf(fmaxf(lbx, lby), lbz), 0.0f);
    float tf = fminf(fminf(ubx, uby), ubz);
    tnear = tn;
    return tn <= F_FMA(tf, SAH_REL_PKT, 1e-37f) && tn <= best_s;
}

#endif /* SAHBVH_TRAVERSE_CORE_CUH */
