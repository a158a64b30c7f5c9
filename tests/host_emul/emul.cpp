/*
 * tests/host_emul/emul.cpp -- TEST-ONLY host build of the per-ray traversal core.
 *
 * Compiles csrc/traverse_core.cuh (the exact source the CUDA kernels use) with g++ so that the
 * flattener, the conservative interior test, the tie rule and the primary-ray arithmetic can be
 * checked against the oracle in this GPU-less container.  It is NOT part of
 * libsahbvh_b200.so, exports only emul_* symbols, and nothing in the product links it.
 */
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <thread>
#include <algorithm>

#include "sahbvh_internal.h"
#include "traverse_core.cuh"

/* the product's scalar GPU hooks are not available in this build */
extern "C" int sahbvh_scalar_trace(const BVHNode *, const Ray *, HitRecord *)
{
    sahbvh_set_error("host emulation build has no GPU");
    return SAHBVH_E_CUDA;
}
extern "C" void sahbvh_scalar_forget(const BVHNode *) {}

struct emul_bvh {
    flat_bvh f;
    wide_bvh w;   /* built lazily by emul_wide_* */
    bool has_wide;
};

static emul_bvh *wrap(sah_tree *t, const Sphere *spheres, int n, int rc)
{
    emul_bvh *e = nullptr;
    if (rc == SAHBVH_OK) {
        e = (emul_bvh *)calloc(1, sizeof(emul_bvh));
        rc = flat_bvh_from_tree(&e->f, t, spheres, n);
        if (rc != SAHBVH_OK) { free(e); e = nullptr; }
    }
    sah_tree_free(t);
    return e;
}

extern "C" emul_bvh *emul_from_pointers(const Sphere *permuted, int n, const BVHNode *root)
{
    sah_tree t;
    if (sah_tree_init(&t, 2 * (int64_t)(n > 0 ? n : 1) + 16) != SAHBVH_OK) return nullptr;
    return wrap(&t, permuted, n, sah_tree_from_pointers(&t, root, permuted, n));
}

extern "C" emul_bvh *emul_build(Sphere *spheres, int n_total, int start, int end, int depth)
{
    sah_tree t;
    if (sah_tree_init(&t, 2 * (int64_t)(end - start + 1) + 16) != SAHBVH_OK) return nullptr;
    return wrap(&t, spheres, n_total, sah_build(&t, spheres, start, end, depth));
}

/* preorder export of the product builder's tree (compared with the oracle/reference tree) */
extern "C" int64_t emul_build_tree(Sphere *spheres, int start, int end, int depth, sah_node *out, int64_t cap)
{
    sah_tree t;
    if (sah_tree_init(&t, 2 * (int64_t)(end - start + 1) + 16) != SAHBVH_OK) return -1;
    int64_t used = -1;
    if (sah_build(&t, spheres, start, end, depth) == SAHBVH_OK && t.used <= cap) {
        memcpy(out, t.nodes, (size_t)t.used * sizeof(sah_node));
        used = t.used;
    }
    sah_tree_free(&t);
    return used;
}

extern "C" void emul_free(emul_bvh *e)
{
    if (!e) return;
    flat_bvh_release(&e->f);
    if (e->has_wide) wide_bvh_release(&e->w);
    free(e);
}

extern "C" void emul_info(const emul_bvh *e, gpu_bvh_info *info) { *info = e->f.info; }

static void trace_range(const emul_bvh *e, const Ray *rays, int64_t lo, int64_t hi, gpu_hit *out, SahCounters *cnt)
{
    int32_t stack[SAH_STACK_MAX];
    SahScene sc;
    sc.n_spheres = e->f.info.n_spheres;
    sc.r_max = e->f.info.max_radius;
    for (int a = 0; a < 3; a++) { sc.lo[a] = e->f.info.scene_lo[a]; sc.hi[a] = e->f.info.scene_hi[a]; }
    sc.geometric_zero_axes = e->f.info.empty_leaves == 0;
    for (int64_t i = lo; i < hi; i++) {
        SahRay r;
        sah_ray_setup(r, rays[i].origin.x, rays[i].origin.y, rays[i].origin.z,
                      rays[i].direction.x, rays[i].direction.y, rays[i].direction.z, sc);
        SahHit h;
        sah_trace_one<true>(e->f.nodes, e->f.spheres4, e->f.info.n_spheres, r, stack, 1, h, cnt);
        memcpy(&out[i], &h, sizeof h);
    }
}

extern "C" void emul_trace(const emul_bvh *e, const Ray *rays, int64_t n, gpu_hit *out, gpu_trace_stats *stats, int threads)
{
    if (threads < 1) threads = 1;
    std::vector<SahCounters> cnt((size_t)threads, SahCounters{0, 0, 0, 0, 0});
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; k++)
        pool.emplace_back(trace_range, e, rays, n * k / threads, n * (k + 1) / threads, out, &cnt[(size_t)k]);
    for (auto &t : pool) t.join();
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->rays = (uint64_t)n;
        for (auto &c : cnt) {
            stats->inner_visits += c.inner;
            stats->sphere_tests += c.spheres;
            stats->exact_box_tests += c.exact;
            if (c.max_stack > stats->max_stack) stats->max_stack = c.max_stack;
        }
        for (int64_t i = 0; i < n; i++) stats->hits += out[i].prim >= 0;
    }
}

static void fill_camera(const Camera *cam, int width, int height, SahCamera *c)
{
    float aspect = (float)width / (float)height;
    float fov_rad = (float)((double)cam->fov * (M_PI / (double)180.0f));
    float half_h = (float)tan((double)(fov_rad / 2.0f));
    float half_w = aspect * half_h;
    Vec3 horizontal = vec3_multiply(cam->right, 2.0f * half_w);
    Vec3 vertical = vec3_multiply(cam->up, 2.0f * half_h);
    c->px = cam->position.x; c->py = cam->position.y; c->pz = cam->position.z;
    c->fx = cam->forward.x; c->fy = cam->forward.y; c->fz = cam->forward.z;
    c->hx = horizontal.x; c->hy = horizontal.y; c->hz = horizontal.z;
    c->vx = vertical.x; c->vy = vertical.y; c->vz = vertical.z;
    c->aspect = aspect; c->fw = (float)width; c->fh = (float)height;
}

extern "C" void emul_primary_rays(const Camera *cam, int width, int height, int y0, int y1, Ray *out)
{
    SahCamera c;
    fill_camera(cam, width, height, &c);
    for (int y = y0; y < y1; y++)
        for (int x = 0; x < width; x++) {
            Ray *r = &out[(size_t)(y - y0) * (size_t)width + (size_t)x];
            r->origin = cam->position;
            sah_primary_dir(c, x, y, r->direction.x, r->direction.y, r->direction.z);
        }
}

/* ---- packet (tile frustum) algorithm, serial restatement of k_trace_primary_packet's logic:
 * same frustum set-up, same conservative box test, same candidate processing and packet-wide
 * cull bound; the warp mechanics (ballots, batch-of-16 pops) are not emulated. ---- */
static SahScene scene_of(const emul_bvh *e)
{
    SahScene sc;
    sc.n_spheres = e->f.info.n_spheres;
    sc.r_max = e->f.info.max_radius;
    for (int a = 0; a < 3; a++) { sc.lo[a] = e->f.info.scene_lo[a]; sc.hi[a] = e->f.info.scene_hi[a]; }
    sc.geometric_zero_axes = e->f.info.empty_leaves == 0;
    return sc;
}

/* sub_w, sub_h > 0: the MACRO-tile variant (k_trace_primary_macro): the tile_w x tile_h tile is traversed once with
 * its own frustum, and a candidate leaf is tested only by the rays of the sub_w x sub_h sub-tiles whose own frustum
 * (with the sub-tile's own cull bound) its box passes. */
extern "C" void emul_trace_primary_packet(const emul_bvh *e, const Camera *cam, int width, int height, int y0, int y1,
                                          int tile_w, int tile_h, int sub_w, int sub_h, gpu_hit *out, gpu_trace_stats *stats)
{
    struct Sub { SahFrustumAxis x, y, z; float best_s; std::vector<int> rays; };
    SahCamera c;
    fill_camera(cam, width, height, &c);
    const SahScene sc = scene_of(e);
    const float *nodes = e->f.nodes, *spheres4 = e->f.spheres4;
    unsigned long long inner = 0, tests = 0, exact = 0, hits = 0;
    std::vector<int32_t> stack;
    for (int ty = y0; ty < y1; ty += tile_h)
        for (int tx = 0; tx < width; tx += tile_w) {
            const int nx = std::min(tile_w, width - tx), ny = std::min(tile_h, y1 - ty), nr = nx * ny;
            std::vector<float> dx(nr), dy(nr), dz(nr), ra(nr), bt(nr, 0.0f);
            std::vector<int32_t> bp(nr, -1);
            float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, min_a = INFINITY;
            for (int i = 0; i < nr; i++) {
                sah_primary_dir(c, tx + i % nx, ty + i / nx, dx[i], dy[i], dz[i]);
                ra[i] = F_ADD(F_ADD(F_MUL(dx[i], dx[i]), F_MUL(dy[i], dy[i])), F_MUL(dz[i], dz[i]));
                mn[0] = fminf(mn[0], dx[i]); mx[0] = fmaxf(mx[0], dx[i]);
                mn[1] = fminf(mn[1], dy[i]); mx[1] = fmaxf(mx[1], dy[i]);
                mn[2] = fminf(mn[2], dz[i]); mx[2] = fmaxf(mx[2], dz[i]);
                min_a = fminf(min_a, ra[i]);
            }
            float ex = fmaxf(fabsf(c.px - sc.lo[0]), fabsf(c.px - sc.hi[0]));
            float ey = fmaxf(fabsf(c.py - sc.lo[1]), fabsf(c.py - sc.hi[1]));
            float ez = fmaxf(fabsf(c.pz - sc.lo[2]), fabsf(c.pz - sc.hi[2]));
            float pad = sqrtf(ex * ex + ey * ey + ez * ez) * 0.001953125f + sc.r_max * 0.001953125f;
            bool geo = sc.geometric_zero_axes != 0;
            SahFrustumAxis FX, FY, FZ;
            sah_frustum_axis_setup(FX, c.px, mn[0], mx[0], pad, geo);
            sah_frustum_axis_setup(FY, c.py, mn[1], mx[1], pad, geo);
            sah_frustum_axis_setup(FZ, c.pz, mn[2], mx[2], pad, geo);
            float min_len = sqrtf(min_a);
            const float cull_abs = (min_len > 0.0f) ? (2.0f * sc.r_max) / min_len : INFINITY;
            float pkt_best_s = INFINITY;
            std::vector<Sub> subs;
            if (sub_w > 0 && sub_h > 0)
                for (int sy = 0; sy < ny; sy += sub_h)
                    for (int sx = 0; sx < nx; sx += sub_w) {
                        Sub sb;
                        float smn[3] = {INFINITY, INFINITY, INFINITY}, smx[3] = {-INFINITY, -INFINITY, -INFINITY};
                        for (int yy = sy; yy < std::min(sy + sub_h, ny); yy++)
                            for (int xx = sx; xx < std::min(sx + sub_w, nx); xx++) {
                                const int i = yy * nx + xx;
                                sb.rays.push_back(i);
                                smn[0] = fminf(smn[0], dx[i]); smx[0] = fmaxf(smx[0], dx[i]);
                                smn[1] = fminf(smn[1], dy[i]); smx[1] = fmaxf(smx[1], dy[i]);
                                smn[2] = fminf(smn[2], dz[i]); smx[2] = fmaxf(smx[2], dz[i]);
                            }
                        sah_frustum_axis_setup(sb.x, c.px, smn[0], smx[0], pad, geo);
                        sah_frustum_axis_setup(sb.y, c.py, smn[1], smx[1], pad, geo);
                        sah_frustum_axis_setup(sb.z, c.pz, smn[2], smx[2], pad, geo);
                        sb.best_s = INFINITY;
                        subs.push_back(sb);
                    }
            stack.clear();
            stack.push_back(0);
            while (!stack.empty()) {
                int32_t nd = stack.back();
                stack.pop_back();
                inner++;
                float tn[2];
                bool pass[2];
                int32_t ref[2];
                const float *w = nodes + (size_t)nd * 16;
                for (int ch = 0; ch < 2; ch++) {
                    const float *q = w + ch * 8;
                    ref[ch] = sah_f2i(q[6]);
                    pass[ch] = sah_frustum_box(FX, FY, FZ, q[0], q[1], q[2], q[3], q[4], q[5], pkt_best_s, tn[ch]);
                }
                bool changed = false;
                for (int ch = 0; ch < 2; ch++) {
                    if (!pass[ch] || ref[ch] >= 0) continue;
                    int32_t sidx = ~ref[ch];
                    if ((uint32_t)sidx >= (uint32_t)sc.n_spheres) continue;
                    const float *s = spheres4 + (size_t)sidx * 4, *q = w + ch * 8;
                    std::vector<char> take(nr, subs.empty() ? 1 : 0);
                    for (Sub &sb : subs) {
                        float tq;
                        if (sah_frustum_box(sb.x, sb.y, sb.z, q[0], q[1], q[2], q[3], q[4], q[5], sb.best_s, tq))
                            for (int i : sb.rays) take[i] = 1;
                    }
                    for (int i = 0; i < nr; i++) {
                        if (!take[i]) continue;
                        SahRay r;
                        r.ox = c.px; r.oy = c.py; r.oz = c.pz;
                        r.dx = dx[i]; r.dy = dy[i]; r.dz = dz[i];
                        r.a = ra[i];
                        float t;
                        tests++;
                        if (!sah_sphere_t(r, s[0], s[1], s[2], s[3], t)) continue;
                        if (!(bp[i] < 0 || t < bt[i] || (t == bt[i] && sidx > bp[i]))) continue;
                        exact++;
                        if (!sah_box_exact(r, q[0], q[1], q[2], q[3], q[4], q[5])) continue;
                        bt[i] = t;
                        bp[i] = sidx;
                        changed = true;
                    }
                }
                if (changed) {
                    float wmax = 0.0f;
                    for (int i = 0; i < nr; i++)
                        wmax = fmaxf(wmax, bp[i] >= 0 ? F_FMA(bt[i], SAH_REL_CULL, cull_abs) : INFINITY);
                    pkt_best_s = wmax;
                    for (Sub &sb : subs) {
                        float wq = 0.0f;
                        for (int i : sb.rays) wq = fmaxf(wq, bp[i] >= 0 ? F_FMA(bt[i], SAH_REL_CULL, cull_abs) : INFINITY);
                        sb.best_s = wq;
                    }
                }
                bool pl = pass[0] && ref[0] >= 0, pr = pass[1] && ref[1] >= 0;
                if (pl && pr) {
                    bool left_near = tn[0] <= tn[1];
                    stack.push_back(left_near ? ref[1] : ref[0]);
                    stack.push_back(left_near ? ref[0] : ref[1]);
                } else if (pl) stack.push_back(ref[0]);
                else if (pr) stack.push_back(ref[1]);
            }
            for (int i = 0; i < nr; i++) {
                SahHit h;
                if (bp[i] >= 0) {
                    SahRay r;
                    r.ox = c.px; r.oy = c.py; r.oz = c.pz;
                    r.dx = dx[i]; r.dy = dy[i]; r.dz = dz[i];
                    const float *s = spheres4 + (size_t)bp[i] * 4;
                    sah_finish_hit(r, bt[i], bp[i], s[0], s[1], s[2], h);
                    hits++;
                } else sah_miss(h);
                memcpy(&out[(size_t)(ty + i / nx - y0) * (size_t)width + (size_t)(tx + i % nx)], &h, sizeof h);
            }
        }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->rays = (uint64_t)(y1 - y0) * (uint64_t)width;
        stats->inner_visits = inner;
        stats->sphere_tests = tests;
        stats->exact_box_tests = exact;
        stats->hits = hits;
    }
}

/* ---- 4-wide quantised tree (csrc/wide.c) + sah_wide_visit, the arithmetic of k_trace_rays_wide ---- */
static int ensure_wide(emul_bvh *e)
{
    if (e->has_wide) return SAHBVH_OK;
    int rc = wide_bvh_from_flat(&e->w, &e->f);
    e->has_wide = rc == SAHBVH_OK;
    return rc;
}

extern "C" int emul_wide_info(emul_bvh *e, int32_t *out3)
{
    int rc = ensure_wide(e);
    if (rc != SAHBVH_OK) return rc;
    out3[0] = e->w.n_wide;
    out3[1] = e->w.n_special;
    out3[2] = e->w.max_depth;
    return SAHBVH_OK;
}

extern "C" int emul_wide_copy(emul_bvh *e, uint32_t *nodes, float *special)
{
    int rc = ensure_wide(e);
    if (rc != SAHBVH_OK) return rc;
    memcpy(nodes, e->w.nodes, (size_t)e->w.n_wide * 64);
    if (e->w.n_special) memcpy(special, e->w.special, (size_t)e->w.n_special * 32);
    return SAHBVH_OK;
}

static void wide_range(const emul_bvh *e, const Ray *rays, int64_t lo, int64_t hi, gpu_hit *out, SahCounters *cnt, int fast_depth,
                       int drain_every)
{
    int32_t fast[SAH_WIDE_STACK_MAX], slow[SAH_WIDE_STACK_MAX];
    const SahStack stk = {fast, 1, fast_depth, slow};
    const SahScene sc = scene_of(e);
    for (int64_t i = lo; i < hi; i++) {
        SahRay r;
        sah_ray_setup(r, rays[i].origin.x, rays[i].origin.y, rays[i].origin.z,
                      rays[i].direction.x, rays[i].direction.y, rays[i].direction.z, sc);
        SahHit h;
        const unsigned long long before = cnt->inner;
        sah_wide_trace_one<true>(e->w.nodes, e->f.spheres4, e->w.special, e->f.info.n_spheres, e->w.n_special, r, stk, drain_every, h, cnt);
        if (cnt->inner - before > cnt->max_visits) cnt->max_visits = cnt->inner - before;
        memcpy(&out[i], &h, sizeof h);
    }
}

extern "C" int emul_wide_trace(emul_bvh *e, const Ray *rays, int64_t n, gpu_hit *out, gpu_trace_stats *stats, int threads,
                               int fast_depth, int drain_every)
{
    int rc = ensure_wide(e);
    if (rc != SAHBVH_OK) return rc;
    if (threads < 1) threads = 1;
    std::vector<SahCounters> cnt((size_t)threads, SahCounters{0, 0, 0, 0, 0});
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; k++)
        pool.emplace_back(wide_range, e, rays, n * k / threads, n * (k + 1) / threads, out, &cnt[(size_t)k], fast_depth,
                          drain_every < 1 ? 1 : drain_every);
    for (auto &t : pool) t.join();
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->rays = (uint64_t)n;
        for (auto &c : cnt) {
            stats->inner_visits += c.inner;
            stats->sphere_tests += c.spheres;
            stats->exact_box_tests += c.exact;
            if (c.max_stack > stats->max_stack) stats->max_stack = c.max_stack;
            if (c.max_visits > stats->max_visits) stats->max_visits = c.max_visits;
        }
        for (int64_t i = 0; i < n; i++) stats->hits += out[i].prim >= 0;
    }
    return SAHBVH_OK;
}

/* ---- soundness of the interior (culling) tests, checked directly --------------------------------------------
 * The design rests on one implication: whenever the REFERENCE's slab test (sah_box_exact, hit.c:49-82) accepts a
 * box for a ray, every interior test the kernels use for that box accepts it too (with no best-t bound).  The
 * parity tests only see a violation when the wrongly culled box held the winning hit; these checkers test the
 * implication itself for every (ray, box) pair.  For a ray with a zero / tiny direction component the reference
 * accepts every slab on that axis while the kernels use the geometric rule, whose guarantee is
 * narrower: no LEAF the reference could report -- own box accepted AND sphere test passed -- is rejected (the rule is
 * monotone in the box, so the leaf's ancestors follow).  Such rays are checked against the leaves only. */
static bool ray_is_regular(const Ray &q)
{
    return fabsf(q.direction.x) >= SAH_TINY_DIR && fabsf(q.direction.y) >= SAH_TINY_DIR && fabsf(q.direction.z) >= SAH_TINY_DIR;
}

struct cull_counts { unsigned long long pairs, accepted, bad_binary, bad_wide; };

/* for rays with a zero / tiny axis: does leaf `sidx` pass the reference's sphere test (hit.c:19-39)? */
static bool leaf_reportable(const emul_bvh *e, const SahRay &r, int32_t sidx)
{
    if ((uint32_t)sidx >= (uint32_t)e->f.info.n_spheres) return false;
    const float *s = e->f.spheres4 + (size_t)sidx * 4;
    float t;
    return sah_sphere_t(r, s[0], s[1], s[2], s[3], t);
}

static int32_t wide_leaf_sphere(const emul_bvh *e, int32_t ref)
{
    const uint32_t code = (uint32_t)~ref;
    if (!(code & SAH_WIDE_SPECIAL)) return (int32_t)code;
    const uint32_t k = code & ~SAH_WIDE_SPECIAL;
    if (k >= (uint32_t)e->w.n_special) return -1;
    return sah_f2i(e->w.special[(size_t)k * 8 + 6]);
}

static void cull_range(const emul_bvh *e, const Ray *rays, int64_t lo, int64_t hi, int node_step, cull_counts *out)
{
    const SahScene sc = scene_of(e);
    int32_t fast[SAH_WIDE_STACK_MAX], slow[SAH_WIDE_STACK_MAX];
    const SahStack stk = {fast, 1, SAH_WIDE_STACK_MAX, slow};
    cull_counts c = {0, 0, 0, 0};
    for (int64_t i = lo; i < hi; i++) {
        const bool regular = ray_is_regular(rays[i]);
        SahRay r;
        sah_ray_setup(r, rays[i].origin.x, rays[i].origin.y, rays[i].origin.z,
                      rays[i].direction.x, rays[i].direction.y, rays[i].direction.z, sc);
        SahRay rw = r;
        sah_ray_to_wide(rw);
        /* binary flattened tree: both child records of every node_step-th node */
        for (int64_t nd = i % node_step; nd < e->f.info.n_inner; nd += node_step)
            for (int ch = 0; ch < 2; ch++) {
                const float *q = e->f.nodes + (size_t)nd * 16 + ch * 8;
                const int32_t ref = sah_f2i(q[6]);
                if (ref == SAHBVH_REF_NONE || (!regular && ref >= 0)) continue;
                c.pairs++;
                if (!sah_box_exact(r, q[0], q[1], q[2], q[3], q[4], q[5])) continue;
                if (!regular && !leaf_reportable(e, r, ~ref)) continue;
                c.accepted++;
                float tn;
                if (!sah_box_conservative(r, q[0], q[1], q[2], q[3], q[4], q[5], INFINITY, tn)) c.bad_binary++;
            }
        /* 4-wide quantised nodes: a child whose full-precision box the reference accepts must come out of
         * sah_wide_visit (as the next node, on the stack, or parked as a pending leaf) */
        for (int64_t nd = i % node_step; nd < e->w.n_wide; nd += node_step) {
            uint32_t w[16];
            memcpy(w, e->w.nodes + (size_t)nd * 16, 64);
            bool want[4];
            int n_want = 0;
            for (int ch = 0; ch < 4; ch++) {
                const float *q = e->w.full + (size_t)nd * 32 + ch * 8;
                const int32_t ref = (int32_t)w[10 + ch];
                want[ch] = ref != SAHBVH_REF_NONE && (regular || ref < 0) && sah_box_exact(r, q[0], q[1], q[2], q[3], q[4], q[5]) &&
                           (regular || leaf_reportable(e, r, wide_leaf_sphere(e, ref)));
                n_want += want[ch];
            }
            if (!n_want) continue;
            bool got[4] = {false, false, false, false};
            int stall_at = -1;
            for (int round = 0; round < 4; round++) {
                int sp = 0;
                int32_t pend0 = SAH_REF_NONE, pend1 = SAH_REF_NONE;
                const int32_t next = sah_wide_visit<false>(rw, w, (int32_t)nd, stk, sp, INFINITY, pend0, pend1, stall_at, nullptr);
                for (int ch = 0; ch < 4; ch++) {
                    const int32_t ref = (int32_t)w[10 + ch];
                    if (ref == SAHBVH_REF_NONE) continue;
                    bool seen = ref == pend0 || ref == pend1 || (stall_at < 0 && ref >= 0 && ref == next);
                    for (int k = 0; k < sp; k++) seen = seen || fast[k] == ref;
                    got[ch] = got[ch] || seen;
                }
                if (stall_at < 0) break; /* a stalled visit parks the remaining leaves when it is repeated */
            }
            for (int ch = 0; ch < 4; ch++)
                if (want[ch] && !got[ch]) c.bad_wide++;
        }
    }
    *out = c;
}

extern "C" int emul_check_culling(emul_bvh *e, const Ray *rays, int64_t n, int node_step, int threads, unsigned long long *out4)
{
    int rc = ensure_wide(e);
    if (rc != SAHBVH_OK) return rc;
    if (threads < 1) threads = 1;
    if (node_step < 1) node_step = 1;
    std::vector<cull_counts> cnt((size_t)threads);
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; k++)
        pool.emplace_back(cull_range, e, rays, n * k / threads, n * (k + 1) / threads, node_step, &cnt[(size_t)k]);
    for (auto &t : pool) t.join();
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    for (auto &c : cnt) { out4[0] += c.pairs; out4[1] += c.accepted; out4[2] += c.bad_binary; out4[3] += c.bad_wide; }
    return SAHBVH_OK;
}

/* packet kernel: if the reference accepts a box for ANY ray of a pixel tile, the tile's frustum test
 * (sah_frustum_box, and sah_frustum_box_simple where it applies) must accept the box */
extern "C" int emul_check_frustum(emul_bvh *e, const Camera *cam, int width, int height, int tile_w, int tile_h, int tile_step,
                                  int node_step, unsigned long long *out4)
{
    int rc = ensure_wide(e);
    if (rc != SAHBVH_OK) return rc;
    SahCamera c;
    fill_camera(cam, width, height, &c);
    const SahScene sc = scene_of(e);
    unsigned long long pairs = 0, accepted = 0, bad = 0, bad_simple = 0;
    int tile_no = 0;
    for (int ty = 0; ty < height; ty += tile_h)
        for (int tx = 0; tx < width; tx += tile_w, tile_no++) {
            if (tile_no % tile_step) continue;
            const int nx = std::min(tile_w, width - tx), ny = std::min(tile_h, height - ty), nr = nx * ny;
            std::vector<SahRay> rr((size_t)nr);
            std::vector<char> regular((size_t)nr);
            float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int i = 0; i < nr; i++) {
                float dx, dy, dz;
                sah_primary_dir(c, tx + i % nx, ty + i / nx, dx, dy, dz);
                sah_ray_setup(rr[(size_t)i], c.px, c.py, c.pz, dx, dy, dz, sc);
                regular[(size_t)i] = fabsf(dx) >= SAH_TINY_DIR && fabsf(dy) >= SAH_TINY_DIR && fabsf(dz) >= SAH_TINY_DIR;
                mn[0] = fminf(mn[0], dx); mx[0] = fmaxf(mx[0], dx);
                mn[1] = fminf(mn[1], dy); mx[1] = fmaxf(mx[1], dy);
                mn[2] = fminf(mn[2], dz); mx[2] = fmaxf(mx[2], dz);
            }
            float ex = fmaxf(fabsf(c.px - sc.lo[0]), fabsf(c.px - sc.hi[0]));
            float ey = fmaxf(fabsf(c.py - sc.lo[1]), fabsf(c.py - sc.hi[1]));
            float ez = fmaxf(fabsf(c.pz - sc.lo[2]), fabsf(c.pz - sc.hi[2]));
            float pad = sqrtf(ex * ex + ey * ey + ez * ez) * 0.001953125f + sc.r_max * 0.001953125f;
            SahFrustumAxis FX, FY, FZ;
            sah_frustum_axis_setup(FX, c.px, mn[0], mx[0], pad, sc.geometric_zero_axes != 0);
            sah_frustum_axis_setup(FY, c.py, mn[1], mx[1], pad, sc.geometric_zero_axes != 0);
            sah_frustum_axis_setup(FZ, c.pz, mn[2], mx[2], pad, sc.geometric_zero_axes != 0);
            const bool simple = sah_frustum_simple(FX, FY, FZ);
            for (int64_t nd = tile_no % node_step; nd < e->w.n_wide; nd += node_step)
                for (int ch = 0; ch < 4; ch++) {
                    const float *q = e->w.full + (size_t)nd * 32 + ch * 8;
                    if (sah_f2i(q[6]) == SAHBVH_REF_NONE) continue;
                    pairs++;
                    /* a ray with a zero / tiny axis counts only for a leaf it could report (see above) */
                    const int32_t ref = sah_f2i(q[6]);
                    bool any = false;
                    for (int i = 0; i < nr && !any; i++)
                        any = sah_box_exact(rr[(size_t)i], q[0], q[1], q[2], q[3], q[4], q[5]) &&
                              (regular[(size_t)i] || (ref < 0 && leaf_reportable(e, rr[(size_t)i], ~ref)));
                    if (!any) continue;
                    accepted++;
                    float tn;
                    if (!sah_frustum_box(FX, FY, FZ, q[0], q[1], q[2], q[3], q[4], q[5], INFINITY, tn)) bad++;
                    if (simple && !sah_frustum_box_simple(FX, FY, FZ, q[0], q[1], q[2], q[3], q[4], q[5], INFINITY, tn)) bad_simple++;
                }
        }
    out4[0] = pairs; out4[1] = accepted; out4[2] = bad; out4[3] = bad_simple;
    return SAHBVH_OK;
}

/* ---- best-t culling never loses the winner -------------------------------------------------------------------
 * best_s = best_t (1 + 2^-9) + 2 r_max / |d| must leave every box on the root-to-leaf path of the reference's
 * winner W acceptable, whatever was found before W.  Any earlier hit has t >= t_W, so the tightest bound a
 * schedule can hold when it reaches W's path is best_s(t_W): the path is checked against exactly that bound
 * (binary FMA test and 4-wide visit).  The winner is found the reference's way: exhaustive descent through the
 * boxes its own slab test accepts (hit.c:91-109). */
struct path_step { int32_t node; int child; };

struct winner_search {
    const emul_bvh *e;
    const SahRay *r;
    bool wide;
    float best_t;
    int32_t best_prim;
    std::vector<path_step> cur;
    std::vector<std::vector<path_step>> best_paths;   /* every leaf that yields the winning record: an empty leaf (Q3)
                                                         aliases a sphere that also has a leaf of its own */
};

static void winner_descend(winner_search &ws, int32_t node)
{
    const int fan = ws.wide ? 4 : 2;
    const float *rec = ws.wide ? ws.e->w.full + (size_t)node * 32 : ws.e->f.nodes + (size_t)node * 16;
    for (int ch = 0; ch < fan; ch++) {
        const float *q = rec + ch * 8;
        const int32_t ref = sah_f2i(q[6]);
        if (ref == SAHBVH_REF_NONE) continue;
        if (!sah_box_exact(*ws.r, q[0], q[1], q[2], q[3], q[4], q[5])) continue;
        ws.cur.push_back({node, ch});
        if (ref >= 0) {
            /* the twin keeps the BINARY tree's refs; the wide child index lives in the quantised node */
            winner_descend(ws, ws.wide ? (int32_t)ws.e->w.nodes[(size_t)node * 16 + 10 + ch] : ref);
        } else {
            const int32_t idx = ~ref;
            float t;
            if ((uint32_t)idx < (uint32_t)ws.e->f.info.n_spheres) {
                const float *s = ws.e->f.spheres4 + (size_t)idx * 4;
                if (sah_sphere_t(*ws.r, s[0], s[1], s[2], s[3], t)) {
                    if (ws.best_prim < 0 || t < ws.best_t || (t == ws.best_t && idx > ws.best_prim)) {
                        ws.best_t = t;
                        ws.best_prim = idx;
                        ws.best_paths.clear();
                        ws.best_paths.push_back(ws.cur);
                    } else if (t == ws.best_t && idx == ws.best_prim) {
                        ws.best_paths.push_back(ws.cur);
                    }
                }
            }
        }
        ws.cur.pop_back();
    }
}

struct bestt_counts { unsigned long long winners, steps, bad_binary, bad_wide, disagree, bad_frustum; };

static void bestt_range(const emul_bvh *e, const Ray *rays, int64_t lo, int64_t hi, bestt_counts *out)
{
    const SahScene sc = scene_of(e);
    int32_t fast[SAH_WIDE_STACK_MAX], slow[SAH_WIDE_STACK_MAX];
    const SahStack stk = {fast, 1, SAH_WIDE_STACK_MAX, slow};
    bestt_counts c = {0, 0, 0, 0, 0, 0};
    for (int64_t i = lo; i < hi; i++) {
        SahRay r;
        sah_ray_setup(r, rays[i].origin.x, rays[i].origin.y, rays[i].origin.z,
                      rays[i].direction.x, rays[i].direction.y, rays[i].direction.z, sc);
        SahRay rw = r;
        sah_ray_to_wide(rw);
        /* the packet kernel's test for the one-ray tile {r}: a real tile's entry bound can only be smaller and its
         * cull bound (max over the tile's rays) only larger, so passing here implies passing in any tile with r */
        SahFrustumAxis FX, FY, FZ;
        {
            float ex = fmaxf(fabsf(r.ox - sc.lo[0]), fabsf(r.ox - sc.hi[0]));
            float ey = fmaxf(fabsf(r.oy - sc.lo[1]), fabsf(r.oy - sc.hi[1]));
            float ez = fmaxf(fabsf(r.oz - sc.lo[2]), fabsf(r.oz - sc.hi[2]));
            float pad = sqrtf(ex * ex + ey * ey + ez * ez) * 0.001953125f + sc.r_max * 0.001953125f;
            sah_frustum_axis_setup(FX, r.ox, r.dx, r.dx, pad, sc.geometric_zero_axes != 0);
            sah_frustum_axis_setup(FY, r.oy, r.dy, r.dy, pad, sc.geometric_zero_axes != 0);
            sah_frustum_axis_setup(FZ, r.oz, r.dz, r.dz, pad, sc.geometric_zero_axes != 0);
        }
        winner_search wb = {e, &r, false, 0.0f, -1, {}, {}};
        winner_descend(wb, 0);
        winner_search ww = {e, &r, true, 0.0f, -1, {}, {}};
        winner_descend(ww, 0);
        /* the record is safe if ONE of the leaves that produce it survives */
        if (wb.best_prim != ww.best_prim || (wb.best_prim >= 0 && wb.best_t != ww.best_t)) c.disagree++;
        if (wb.best_prim < 0) continue;
        c.winners++;
        const float best_s = F_FMA(wb.best_t, SAH_REL_CULL, r.cull_abs);
        bool any_ok = false, any_ok_f = false;
        for (const auto &path : wb.best_paths) {
            bool ok = true, ok_f = true;
            for (const path_step &p : path) {
                const float *q = e->f.nodes + (size_t)p.node * 16 + p.child * 8;
                float tn;
                c.steps++;
                ok = ok && sah_box_conservative(r, q[0], q[1], q[2], q[3], q[4], q[5], best_s, tn);
                ok_f = ok_f && sah_frustum_box(FX, FY, FZ, q[0], q[1], q[2], q[3], q[4], q[5], best_s, tn);
                if (sah_frustum_simple(FX, FY, FZ))
                    ok_f = ok_f && sah_frustum_box_simple(FX, FY, FZ, q[0], q[1], q[2], q[3], q[4], q[5], best_s, tn);
            }
            any_ok = any_ok || ok;
            any_ok_f = any_ok_f || ok_f;
        }
        if (!any_ok) c.bad_binary++;
        if (!any_ok_f) c.bad_frustum++;
        any_ok = false;
        for (const auto &path : ww.best_paths) {
            bool ok = true;
            for (const path_step &p : path) {
                uint32_t w[16];
                memcpy(w, e->w.nodes + (size_t)p.node * 16, 64);
                const int32_t want = (int32_t)w[10 + p.child];
                bool got = false;
                int stall_at = -1;
                for (int round = 0; round < 4 && !got && want != SAHBVH_REF_NONE; round++) {   /* NONE: the leaf was dropped */
                    int sp = 0;
                    int32_t pend0 = SAH_REF_NONE, pend1 = SAH_REF_NONE;
                    const int32_t next = sah_wide_visit<false>(rw, w, p.node, stk, sp, best_s, pend0, pend1, stall_at, nullptr);
                    got = want == pend0 || want == pend1 || (stall_at < 0 && want >= 0 && want == next);
                    for (int k = 0; k < sp; k++) got = got || fast[k] == want;
                    if (stall_at < 0) break;
                }
                c.steps++;
                ok = ok && got;
            }
            any_ok = any_ok || ok;
        }
        if (!any_ok) c.bad_wide++;
    }
    *out = c;
}

extern "C" int emul_check_best_t(emul_bvh *e, const Ray *rays, int64_t n, int threads, unsigned long long *out6)
{
    int rc = ensure_wide(e);
    if (rc != SAHBVH_OK) return rc;
    if (threads < 1) threads = 1;
    std::vector<bestt_counts> cnt((size_t)threads);
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; k++)
        pool.emplace_back(bestt_range, e, rays, n * k / threads, n * (k + 1) / threads, &cnt[(size_t)k]);
    for (auto &t : pool) t.join();
    for (int k = 0; k < 6; k++) out6[k] = 0;
    for (auto &c : cnt) {
        out6[0] += c.winners; out6[1] += c.steps; out6[2] += c.bad_binary; out6[3] += c.bad_wide; out6[4] += c.disagree;
        out6[5] += c.bad_frustum;
    }
    return SAHBVH_OK;
}
