/*
 * sahbvh_kernels.cu -- sm_100a kernels + the C-ABI of include/sahbvh_gpu.h.
 *
 *   K1 k_trace_rays     ray buffer -> hit buffer        replaces the loop benchmark.c:230-245 /
 *                                                       renderer.c:29-32 over ray_bvh_intersect (hit.c:91)
 *   K2 k_trace_primary  fused camera rays + closest hit replaces main.c:356-366 (get_camera_ray ray.c:17 +
 *                                                       the BVH branch of trace_ray renderer.c:29-32)
 *   K2' k_primary_rays  ray generation only             parity aid for ray.c:17-32
 *   K3 k_brute_force    all-spheres closest hit         renderer.c:33-44 (ties -> lowest index)
 *
 * FP32 CUDA-core work (box slabs, sphere quadratics); tensor cores are not applicable.
 * Built with -fmad=false; the parity-critical arithmetic additionally uses explicit _rn
 * intrinsics (traverse_core.cuh), the conservative interior test uses explicit FMAs.
 */
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "sahbvh_internal.h"
#include "traverse_core.cuh"
#include "gpu_handle.cuh"

static_assert(sizeof(SahHit) == sizeof(gpu_hit), "hit record");
static_assert(SAH_STACK_MAX == SAHBVH_STACK_MAX, "stack depth");

#define TRACE_THREADS 128
#ifndef SAHBVH_HOST_DIRECT_DEFAULT
#define SAHBVH_HOST_DIRECT_DEFAULT 1 /* measured on the 1 M-sphere 4K frame: 0.674 ms direct vs 0.726 ms staged in 1 Mi-pixel bands */
#endif
#define PRIMARY_THREADS 256 /* 32 x 8 pixels per CTA, one 8x4 tile per warp */

std::atomic<unsigned long long> g_launches{0};
/* the process-wide default can be pinned from the environment (SAHBVH_PRIMARY_MODE=0..3): profiler runs use it, because
 * under ncu every launch is serialised and replayed and the AUTO tuner's event times mean nothing */
static int initial_primary_mode()
{
    const char *e = getenv("SAHBVH_PRIMARY_MODE");
    const int v = e ? atoi(e) : SAHBVH_PRIMARY_AUTO;
    return (e && v >= SAHBVH_PRIMARY_PER_RAY && v <= SAHBVH_PRIMARY_AUTO) ? v : SAHBVH_PRIMARY_AUTO;
}
static std::atomic<int> g_primary_mode{initial_primary_mode()};
static std::atomic<int> g_ray_mode{SAHBVH_RAYS_WIDE};
static std::atomic<int> g_build_mode{SAHBVH_BUILD_DEVICE};
static std::atomic<int> g_keep_binary{0};

extern "C" int sahbvh_set_keep_binary_tree(int keep)
{
    g_keep_binary.store(keep ? 1 : 0);
    return SAHBVH_OK;
}

int sahbvh_need_binary_tree(gpu_bvh *b)
{
    if (b->d_nodes) return SAHBVH_OK;
    sahbvh_set_error("this handle was created without the 64-byte binary tree (only SAHBVH_PRIMARY_PER_RAY, SAHBVH_RAYS_COOP and "
                     "SAHBVH_RAYS_ONE_SHOT read it): call sahbvh_set_keep_binary_tree(1) before gpu_bvh_upload / gpu_bvh_build");
    return SAHBVH_E_ARG;
}

extern "C" int sahbvh_set_build_mode(int mode)
{
    if (mode != SAHBVH_BUILD_HOST && mode != SAHBVH_BUILD_DEVICE) return SAHBVH_E_ARG;
    g_build_mode.store(mode);
    return SAHBVH_OK;
}

extern "C" int sahbvh_set_ray_mode(int mode)
{
    if (mode != SAHBVH_RAYS_ONE_SHOT && mode != SAHBVH_RAYS_COOP && mode != SAHBVH_RAYS_WIDE) return SAHBVH_E_ARG;
    g_ray_mode.store(mode);
    return SAHBVH_OK;
}

extern "C" int sahbvh_set_primary_mode(int mode)
{
    if (mode != SAHBVH_PRIMARY_PER_RAY && mode != SAHBVH_PRIMARY_PACKET && mode != SAHBVH_PRIMARY_MACRO && mode != SAHBVH_PRIMARY_AUTO) return SAHBVH_E_ARG;
    g_primary_mode.store(mode);
    return SAHBVH_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* kernels                                                                                      */
/* ------------------------------------------------------------------------------------------ */

__device__ __forceinline__ void store_hit(gpu_hit *out, const SahHit &h)
{
    float4 a = make_float4(h.t, __int_as_float(h.prim), h.px, h.py);
    float4 b = make_float4(h.pz, h.nx, h.ny, h.nz);
    float4 *o = reinterpret_cast<float4 *>(out);
    o[0] = a;
    o[1] = b;
}

__device__ __forceinline__ void flush_counters(const SahCounters &c, unsigned long long hit, unsigned long long *g)
{
    unsigned long long v[4] = {c.inner, c.spheres, c.exact, hit};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&g[k], x);
    }
    unsigned long long m = c.max_stack;
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long y = __shfl_xor_sync(0xffffffffu, m, o);
        m = y > m ? y : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(&g[4], m);
}

template <bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS)
k_trace_rays(const float *__restrict__ nodes, const float *__restrict__ spheres4, SahScene sc,
             const float *__restrict__ rays, long long n, gpu_hit *__restrict__ out, unsigned long long *counters)
{
    int32_t stack[SAH_STACK_MAX];
    SahCounters cnt = {0, 0, 0, 0};
    unsigned long long hits = 0;
    const long long stride = (long long)gridDim.x * TRACE_THREADS;
    for (long long i = (long long)blockIdx.x * TRACE_THREADS + threadIdx.x; i < n; i += stride) {
        const float *rp = rays + i * 6;
        SahRay r;
        sah_ray_setup(r, __ldg(rp), __ldg(rp + 1), __ldg(rp + 2), __ldg(rp + 3), __ldg(rp + 4), __ldg(rp + 5), sc);
        SahHit h;
        sah_trace_one<COUNT>(nodes, spheres4, sc.n_spheres, r, stack, 1, h, &cnt);
        store_hit(out + i, h);
        if (COUNT) hits += (h.prim >= 0);
    }
    if (COUNT) flush_counters(cnt, hits, counters);
}

/*
 * K1c: ray buffer -> hit buffer, persistent and warp-synchronous (default for gpu_trace_closest_hit*).
 *
 *  - cooperative node fetch: an incoherent warp visits 32 different 64-byte nodes per step.  Loaded
 *    per lane (4 x LDG.128 at a 64-byte stride) every instruction touches 32 cache lines = 32 L1
 *    wavefronts, 128 per step -- the measured bound of k_trace_rays on bounce rays.  Here 4 lanes
 *    fetch the 4 x 16 bytes of ONE node in the same instruction (8 lines per instruction, 32 per
 *    step), park them in shared memory (XOR-swizzled so both the 128-bit stores and the 128-bit
 *    loads are bank-conflict free) and every lane reads its own node back.
 *  - ray refill: a lane whose ray finished stores its record and takes the next ray index from the
 *    warp's chunk (one atomicAdd per 128 rays), so lanes do not idle while the slowest ray of a warp
 *    finishes.
 * Per-ray arithmetic is sah_visit(), identical to the one-shot kernels.
 */
#define COOP_WARPS 4
#define COOP_CHUNK 128
#ifndef COOP_MIN_BLOCKS
#define COOP_MIN_BLOCKS 8 /* 64 registers -> 32 resident warps per SM: the kernel is latency bound on node fetches */
#endif

template <bool COUNT>
__global__ void __launch_bounds__(COOP_WARPS * 32, COOP_MIN_BLOCKS)
k_trace_rays_coop(const float *__restrict__ nodes, const float *__restrict__ spheres4, SahScene sc,
                  const float *__restrict__ rays, long long n, gpu_hit *__restrict__ out,
                  unsigned long long *next_ray, unsigned long long *counters)
{
    __shared__ float4 s_nodes[COOP_WARPS][128];
    int32_t stack[SAH_STACK_MAX];
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 *sm = s_nodes[warp];
    const float4 *nodes4 = reinterpret_cast<const float4 *>(nodes);
    SahCounters cnt = {0, 0, 0, 0};
    unsigned long long hits = 0;

    SahRay r;
    sah_ray_setup(r, 0, 0, 0, 0, 0, 0, sc);
    float best_t = 0.0f, best_s = INFINITY;
    int32_t best_prim = -1, node = SAH_NODE_DONE;
    int sp = 0;
    long long ray_idx = -1;
    long long chunk_next = 0, chunk_end = 0; /* warp-uniform */
    bool more = true;                        /* warp-uniform: the global pool may still have rays */
    const int c4 = lane & 3;

    for (;;) {
        unsigned idle = __ballot_sync(full, node == SAH_NODE_DONE);
        if (idle) {
            if (chunk_next >= chunk_end && more) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(next_ray, (unsigned long long)COOP_CHUNK);
                base = __shfl_sync(full, base, 0);
                if ((long long)base >= n) {
                    more = false;
                } else {
                    chunk_next = (long long)base;
                    chunk_end = (long long)base + COOP_CHUNK < n ? (long long)base + COOP_CHUNK : n;
                }
            }
            const long long avail = chunk_end - chunk_next;
            if (avail > 0) {
                const int need = __popc(idle);
                const int take = (long long)need < avail ? need : (int)avail;
                const int my_rank = __popc(idle & ((1u << lane) - 1u));
                if (node == SAH_NODE_DONE && my_rank < take) {
                    ray_idx = chunk_next + my_rank;
                    const float *rp = rays + ray_idx * 6;
                    sah_ray_setup(r, __ldg(rp), __ldg(rp + 1), __ldg(rp + 2), __ldg(rp + 3), __ldg(rp + 4), __ldg(rp + 5), sc);
                    best_t = 0.0f; best_s = INFINITY; best_prim = -1;
                    sp = 0;
                    node = 0;
                }
                chunk_next += take;
                idle = __ballot_sync(full, node == SAH_NODE_DONE);
            }
            if (idle == full) {
                if (!more && chunk_next >= chunk_end) break;
                continue;
            }
        }
        /* ---- cooperative fetch: lane L brings 16 bytes (word c4) of the node owned by lane (L>>2)+8i ---- */
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int owner = (lane >> 2) + 8 * i;
            const int32_t onode = __shfl_sync(full, node, owner);
            if (onode != SAH_NODE_DONE)
                sm[c4 * 32 + (owner ^ (2 * c4))] = __ldg(nodes4 + (size_t)onode * 4 + c4);
        }
        __syncwarp();
        if (node != SAH_NODE_DONE) {
            const float4 a0 = sm[lane], a1 = sm[32 + (lane ^ 2)], b0 = sm[64 + (lane ^ 4)], b1 = sm[96 + (lane ^ 6)];
            const SahF4 l0 = {a0.x, a0.y, a0.z, a0.w}, l1 = {a1.x, a1.y, a1.z, a1.w};
            const SahF4 r0 = {b0.x, b0.y, b0.z, b0.w}, r1 = {b1.x, b1.y, b1.z, b1.w};
            node = sah_visit<COUNT>(r, l0, l1, r0, r1, spheres4, sc.n_spheres, stack, 1, sp, best_t, best_prim, best_s, &cnt);
            if (node == SAH_NODE_DONE) {
                SahHit h;
                if (best_prim >= 0) {
                    const float4 s = __ldg(reinterpret_cast<const float4 *>(spheres4) + best_prim);
                    sah_finish_hit(r, best_t, best_prim, s.x, s.y, s.z, h);
                    if (COUNT) hits++;
                } else {
                    sah_miss(h);
                }
                store_hit(out + ray_idx, h);
            }
        }
        __syncwarp();
    }
    if (COUNT) flush_counters(cnt, hits, counters);
}

/*
 * K1w: ray buffer -> hit buffer over the 4-WIDE QUANTISED tree (default for gpu_trace_closest_hit*).
 *
 * Same contract as K1/K1c (leaf tests are the reference's own, see traverse_core.cuh); what changes
 * is everything that bounded K1c in the ncu profile (profiles/r1_coop_bounce_1M_ncu_summary.txt):
 *  - half the dependent node fetches per ray (two binary levels per 64-byte node, wide.c);
 *  - the node is fetched by its own lane with two 256-bit loads (LDG.E.256, sm_100): 64 L1
 *    wavefronts per warp step, the same as the 4-lanes-per-node staging through shared memory,
 *    without its SHFL/STS/LDS instructions;
 *  - the traversal stack lives in a bank-conflict-free shared-memory column (16 entries, overflow
 *    in local memory): the local-memory stack of K1c missed L1 and paid an L2 round trip per pop;
 *  - finished lanes park until WIDE_REFILL_MIN of them are idle, then store their records and
 *    take new rays together, so the divergent finish/set-up code runs with several lanes active.
 */
#define WIDE_WARPS 4
#define WIDE_CHUNK 128
#define WIDE_FAST_STACK 16
#ifndef WIDE_MIN_BLOCKS
#define WIDE_MIN_BLOCKS 8 /* 64 registers, no spills; measured 16.97 ms vs 17.46 ms at 7 (config 4, 2^24 bounce rays) */
#endif
#ifndef WIDE_REFILL_MIN
#define WIDE_REFILL_MIN 4
#endif
#ifndef WIDE_TOP_NODES
#define WIDE_TOP_NODES 0 /* > 0: that many top nodes in shared memory (experiment, see the kernel) */
#endif

__device__ __forceinline__ void ld_wide_node(const uint32_t *__restrict__ nodes, int32_t node, uint32_t (&w)[16])
{
    const uint32_t *p = nodes + (size_t)node * 16;
#ifndef WIDE_LD
#define WIDE_LD "ld.global.nc.v8.b32"
#endif
    asm volatile(WIDE_LD " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
    asm volatile(WIDE_LD " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                 : "l"(p + 8));
}

template <bool COUNT>
__global__ void __launch_bounds__(WIDE_WARPS * 32, WIDE_MIN_BLOCKS)
k_trace_rays_wide(const uint32_t *__restrict__ nodes, int n_wide, const float *__restrict__ spheres4, const float *__restrict__ special,
                  int n_special, SahScene sc, const float *__restrict__ rays, long long n, gpu_hit *__restrict__ out,
                  int32_t *__restrict__ prims_out, unsigned long long *next_ray, unsigned long long *counters)
{
    /* out != NULL: 32-byte records; else prims_out receives the 4-byte primitive id per ray (point and normal are
     * never computed: sahbvh_hit_from_prim / gpu_hits_from_prims derive them on demand) */
    __shared__ int32_t s_stack[WIDE_FAST_STACK][WIDE_WARPS * 32];
#if WIDE_TOP_NODES > 0
    /* experiment (north_star "top BVH levels in shared memory"): the first WIDE_TOP_NODES nodes of the breadth-first
     * numbering (3 wide levels = 85 nodes = 5.3 KB) are read from shared memory instead of L1/L2; measured in
     * DESIGN.md section 4, off by default */
    __shared__ uint4 s_top[WIDE_TOP_NODES * 4];
    const int top_n = n_wide < WIDE_TOP_NODES ? n_wide : WIDE_TOP_NODES;
    for (int i = threadIdx.x; i < top_n * 4; i += WIDE_WARPS * 32) s_top[i] = __ldg(reinterpret_cast<const uint4 *>(nodes) + i);
    __syncthreads();
#endif
    int32_t slow[SAH_WIDE_STACK_MAX - WIDE_FAST_STACK];
    const SahStack stk = {&s_stack[0][threadIdx.x], WIDE_WARPS * 32, WIDE_FAST_STACK, slow};
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    SahCounters cnt = {0, 0, 0, 0};
    unsigned long long hits = 0;

    SahRay r;
    sah_ray_setup(r, 0, 0, 0, 0, 0, 0, sc);
    float best_t = 0.0f, best_s = INFINITY;
    int32_t best_prim = -1, node = SAH_NODE_DONE;
    int32_t pend0 = SAH_REF_NONE, pend1 = SAH_REF_NONE; /* parked leaf refs (sah_wide_visit) */
    int stall_at = -1;                                   /* child index a stalled visit resumes from (sah_wide_visit) */
    int sp = 0;
    long long ray_idx = -1;                  /* >= 0: this lane owns a ray (in flight, or finished and not yet stored) */
    long long chunk_next = 0, chunk_end = 0; /* warp-uniform */
    bool more = true;                        /* warp-uniform: the global pool may still have rays */
    unsigned dead = 0;                       /* warp-uniform: lanes that will get no more rays */
    unsigned long long visits0 = 0;          /* COUNT: cnt.inner when this lane's ray started */
    int thresh = WIDE_REFILL_MIN;            /* warp-uniform: waiting lanes that trigger a service round */
    static_assert(SAH_NODE_DONE == -1, "lane state encoding");

    for (;;) {
        /* node >= 0: traversing; SAH_NODE_DONE (-1): no ray / finished; <= -2: stalled at node -2-node
         * (both pending slots full).  Lanes that do not traverse wait for a service round. */
        const unsigned act = __ballot_sync(full, node >= 0);
        if (__popc(act | dead) <= 32 - thresh) {
            /* ---- service round, for the whole warp at once ----
             * 1. deferred leaf tests of every lane (also of the lanes still traversing) */
            sah_wide_drain<COUNT>(r, spheres4, special, sc.n_spheres, n_special, pend0, pend1, best_t, best_prim, best_s, &cnt);
            if (node < SAH_NODE_DONE) node = -2 - node;
            /* 2. finished rays: store the record */
            if (node == SAH_NODE_DONE && ray_idx >= 0) {
                if (out) {
                    SahHit h;
                    if (best_prim >= 0) {
                        const float4 s = __ldg(reinterpret_cast<const float4 *>(spheres4) + best_prim);
                        sah_finish_hit(r, best_t, best_prim, s.x, s.y, s.z, h);
                    } else {
                        sah_miss(h);
                    }
                    store_hit(out + ray_idx, h);
                } else {
                    prims_out[ray_idx] = best_prim;
                }
                if (COUNT && best_prim >= 0) hits++;
                ray_idx = -1;
                if (COUNT) atomicMax(&counters[5], cnt.inner - visits0);
            }
            /* 3. free lanes take new rays */
            const unsigned idle = __ballot_sync(full, node == SAH_NODE_DONE);
            if (idle && chunk_next >= chunk_end && more) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(next_ray, (unsigned long long)WIDE_CHUNK);
                base = __shfl_sync(full, base, 0);
                if ((long long)base >= n) {
                    more = false;
                } else {
                    chunk_next = (long long)base;
                    chunk_end = (long long)base + WIDE_CHUNK < n ? (long long)base + WIDE_CHUNK : n;
                }
            }
            const long long avail = chunk_end - chunk_next;
            if (idle && avail > 0) {
                const int need = __popc(idle);
                const int take = (long long)need < avail ? need : (int)avail;
                const int my_rank = __popc(idle & ((1u << lane) - 1u));
                if (node == SAH_NODE_DONE && my_rank < take) {
                    ray_idx = chunk_next + my_rank;
                    const float *rp = rays + ray_idx * 6;
                    sah_ray_setup(r, __ldg(rp), __ldg(rp + 1), __ldg(rp + 2), __ldg(rp + 3), __ldg(rp + 4), __ldg(rp + 5), sc);
                    sah_ray_to_wide(r);
                    best_t = 0.0f; best_s = INFINITY; best_prim = -1;
                    sp = 0;
                    node = 0;
                    if (COUNT) visits0 = cnt.inner;
                }
                chunk_next += take;
            }
            if (!more && chunk_next >= chunk_end) {
                /* pool exhausted: lanes without a ray are out; flush every finished ray at once */
                dead = __ballot_sync(full, node == SAH_NODE_DONE);
                thresh = 1;
                if (dead == full) break;
            }
        }
        if (node >= 0) {
            uint32_t w[16];
#if WIDE_TOP_NODES > 0
            if (node < top_n) {
                const uint4 *q = s_top + node * 4;
                const uint4 a = q[0], b4 = q[1], c4 = q[2], d4 = q[3];
                w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b4.x; w[5] = b4.y; w[6] = b4.z; w[7] = b4.w;
                w[8] = c4.x; w[9] = c4.y; w[10] = c4.z; w[11] = c4.w; w[12] = d4.x; w[13] = d4.y; w[14] = d4.z; w[15] = d4.w;
            } else
#endif
            ld_wide_node(nodes, node, w);
            node = sah_wide_visit<COUNT>(r, w, node, stk, sp, best_s, pend0, pend1, stall_at, &cnt);
            if (stall_at >= 0) node = -2 - node;
        }
    }
    if (COUNT) flush_counters(cnt, hits, counters);
}

template <bool COUNT>
__global__ void __launch_bounds__(PRIMARY_THREADS)
k_trace_primary(const float *__restrict__ nodes, const float *__restrict__ spheres4, SahScene sc,
                SahCamera cam, int width, int y0, int y1, gpu_hit *__restrict__ out, float2 *__restrict__ ids,
                int ids_mode, unsigned long long *counters)
{
    int32_t stack[SAH_STACK_MAX];
    SahCounters cnt = {0, 0, 0, 0};
    unsigned long long hits = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int y = y0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (x < width && y < y1) {
        float dx, dy, dz;
        sah_primary_dir(cam, x, y, dx, dy, dz);
        SahRay r;
        sah_ray_setup(r, cam.px, cam.py, cam.pz, dx, dy, dz, sc);
        SahHit h;
        sah_trace_one<COUNT>(nodes, spheres4, sc.n_spheres, r, stack, 1, h, &cnt);
        const size_t pix = (size_t)(y - y0) * (size_t)width + (size_t)x;
        if (out) store_hit(out + pix, h);
        if (ids_mode == SAH_IDS_PRIM) reinterpret_cast<int32_t *>(ids)[pix] = h.prim;
        else if (ids_mode == SAH_IDS_TPRIM) ids[pix] = make_float2(h.t, __int_as_float(h.prim));
        if (COUNT) hits += (h.prim >= 0);
    }
    if (COUNT) flush_counters(cnt, hits, counters);
}

/*
 * Tuning notes (1 M spheres, 3840x2160, B200): 8 blocks of 4 warps per SM at <= 64 registers is the sweet spot --
 * 9 blocks (56 registers) measures the same, 10 (48 registers, spills) and 2- or 8-warp blocks are slower; a
 * persistent grid pulling tiles from an atomic counter was 6-20 % SLOWER than one tile per warp (the tile loop costs
 * registers and adjacent tiles no longer run on the same SM at the same time), so the grid covers the frame.
 *
 * K2p: packet traversal for primary rays.  One WARP owns a 16x8 pixel tile (4 rays per lane,
 * shared origin).  Interior culling is done ONCE PER TILE over the 4-wide tree (full-precision
 * 128-byte nodes, wide.c): the lanes test the 32 child boxes of up
 * to 8 popped nodes in parallel against the tile's direction-interval frustum
 * (sah_frustum_box), surviving inner children are pushed back in near-last order, surviving
 * leaves become candidates.  Every ray of the tile then runs the reference-exact sphere test
 * (and, for a would-be winner, the reference-exact leaf box test) on each candidate, so the
 * result is the same "best over leaves whose own tests pass" as the per-ray kernel.
 */
#ifndef PKT_WARPS
#define PKT_WARPS 4        /* the block's 4 warps own a 2x2 group of tiles = 32 x 16 pixels (their frusta overlap more than a 4x1 strip's) */
#endif
#define PKT_GROUP_ROWS (PKT_WARPS / 2)   /* a block = 2 tiles across (one 32-pixel = 128-byte id row) x PKT_GROUP_ROWS tiles down */
static_assert(PKT_WARPS == 2 || PKT_WARPS == 4 || PKT_WARPS == 8, "PKT_WARPS");
#ifndef PKT_WARP_OUT
#define PKT_WARP_OUT 0     /* 1: every warp stores its own tile (64-byte id segments, no block barrier); tuning experiment */
#endif
#ifndef PKT_STACK
#define PKT_STACK 512
#endif
#define PKT_K 4            /* rays per lane: tile = 16 x (2*PKT_K) pixels; 16x12 and 16x16 tiles measured slower */
#define PKT_TILE_H (2 * PKT_K)
#ifndef PKT_UNROLL
#define PKT_UNROLL 4 /* per-ray loops: rolled (1) measures the same within 1%; the state is in shared memory either way */
#endif
static constexpr int kPktUnroll = PKT_UNROLL;
#ifndef PKT_MIN_BLOCKS
#define PKT_MIN_BLOCKS 8
#endif
/* Stack discipline (per-warp shared-memory stack of PKT_STACK refs, unchecked pushes): while top <= dfs_from a
 * step pops up to 8 entries and pushes at most 32 (net +24); above dfs_from it pops ONE entry (plain DFS: net +3
 * per wide level, and the tree has at most `depth` levels).  With dfs_from = PKT_STACK - 32 - 3*(depth + 2) the
 * stack never holds more than dfs_from + 24 + 3*depth <= PKT_STACK - 14 entries, whatever the tree looks like
 * (depth <= 62: flatten.c refuses deeper trees).  sahbvh_packet_dfs_from() computes it at launch. */
static std::atomic<int> g_pkt_dfs_override{0};
/* diagnostic: force the threshold (0 = derive it from the tree depth); tests use it to run whole frames in DFS mode
 * and to check the stack bound dfs_from + 24 + 3*depth against gpu_trace_stats.max_stack */
extern "C" int sahbvh_set_packet_dfs_from(int v)
{
    if (v < 0 || v > PKT_STACK - 32 - 3 * (SAHBVH_STACK_MAX + 2)) return SAHBVH_E_ARG;
    g_pkt_dfs_override.store(v);
    return SAHBVH_OK;
}
static int sahbvh_packet_dfs_from(int tree_depth, int stack_entries = PKT_STACK)
{
    int d = tree_depth < 1 ? 1 : tree_depth;
    if (d > SAHBVH_STACK_MAX) d = SAHBVH_STACK_MAX;
    int v = stack_entries - 32 - 3 * (d + 2);
    v = v < 8 ? 8 : v;
    const int forced = g_pkt_dfs_override.load();
    return forced > 0 && forced < v ? forced : v;   /* the diagnostic override can only lower the threshold */
}
static_assert(PKT_STACK - 32 - 3 * (SAHBVH_STACK_MAX + 2) >= 8, "PKT_STACK too small for the deepest admissible tree");

/* warp-wide min / max of non-NaN floats with ONE redux.sync each: the bit pattern of a float, with the magnitude bits
 * flipped when the sign is set, orders as a signed integer exactly like the float (-0 < +0, infinities included) */
__device__ __forceinline__ int float_order_key(float v)
{
    const int i = __float_as_int(v);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float warp_min(float v)
{
    const int k = __reduce_min_sync(0xffffffffu, float_order_key(v));
    return __int_as_float(k ^ ((k >> 31) & 0x7fffffff));
}
__device__ __forceinline__ float warp_max(float v)
{
    const int k = __reduce_max_sync(0xffffffffu, float_order_key(v));
    return __int_as_float(k ^ ((k >> 31) & 0x7fffffff));
}
/* same for values known to be >= +0 (or +inf): the raw bits already order as unsigned integers */
__device__ __forceinline__ float warp_max_nonneg(float v)
{
    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}

/* 32-byte record with ONE 256-bit store (sm_100 st.global.v8): a warp writing a 32-pixel row emits 1 KiB contiguous */
__device__ __forceinline__ void store_hit256(gpu_hit *out, const SahHit &h)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out), "r"(__float_as_uint(h.t)), "r"((uint32_t)h.prim),
                 "r"(__float_as_uint(h.px)), "r"(__float_as_uint(h.py)), "r"(__float_as_uint(h.pz)), "r"(__float_as_uint(h.nx)),
                 "r"(__float_as_uint(h.ny)), "r"(__float_as_uint(h.nz))
                 : "memory");
}

/* RECORDS: write the 32-byte hit records to `out` (else `out` is not touched and point/normal are never computed).
 * ids_mode: SAH_IDS_NONE / SAH_IDS_PRIM (int32 per pixel) / SAH_IDS_TPRIM ({float t, int32 prim} per pixel) into `ids`. */
template <bool COUNT, bool RECORDS>
__global__ void __launch_bounds__(PKT_WARPS * 32, PKT_MIN_BLOCKS)
k_trace_primary_packet(const float *__restrict__ nodes, const float *__restrict__ spheres4, SahScene sc,
                       SahCamera cam, int width, int y0, int y1, gpu_hit *__restrict__ out, float2 *__restrict__ ids,
                       int ids_mode, int dfs_from, unsigned long long *counters)
{
    /* The traversal loop is a chain of dependent node fetches, so it needs resident warps more than registers:
     * the per-ray state (direction, d.d, best t, best index of the lane's 4 rays) lives in shared memory -- it is
     * only touched when a candidate leaf is tested (about 5 times per tile) -- which leaves 62 registers per
     * thread and 8 CTAs (32 warps) per SM instead of 5. */
    __shared__ int32_t s_stack[PKT_WARPS][PKT_STACK];
    __shared__ float s_ray[PKT_WARPS][PKT_K][4][32];   /* dx, dy, dz, d.d */
    __shared__ float s_bt[PKT_WARPS][PKT_K][32];
    __shared__ int32_t s_bp[PKT_WARPS][PKT_K][32];
    __shared__ float s_tt[PKT_WARPS][PKT_K][32];       /* t of a would-be winner between the sphere and the box test */
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (width + 15) >> 4, tiles_y = (y1 - y0 + PKT_TILE_H - 1) / PKT_TILE_H;
    SahCounters cnt = {0, 0, 0, 0};
    unsigned long long hits = 0;
    const unsigned groups_x = (unsigned)(tiles_x + 1) >> 1;
    const int gy = (int)(blockIdx.x / groups_x), gx = (int)(blockIdx.x % groups_x);
    const int ty = gy * PKT_GROUP_ROWS + (warp >> 1), tx = gx * 2 + (warp & 1);
    if (tx < tiles_x && ty < tiles_y) {
        int32_t *stack = s_stack[warp];
        /* ---- this lane's 4 rays: pixel (tx*16 + (k&1)*8 + lane%8, y0 + ty*8 + (k>>1)*4 + lane/8) ---- */
        const int px0 = tx * 16 + (lane & 7), py0 = y0 + ty * PKT_TILE_H + (lane >> 3);
        unsigned valid = 0;
        float lminx = INFINITY, lminy = INFINITY, lminz = INFINITY, lmaxx = -INFINITY, lmaxy = -INFINITY, lmaxz = -INFINITY;
        float lmin_a = INFINITY;
        /* the lane's rays use 2 distinct columns and PKT_K/2 distinct rows: one division each (main.c:362-363) */
        const float u0 = sah_pixel_u(cam, px0), u1 = sah_pixel_u(cam, px0 + 8);
        float vrow[PKT_K / 2];
#pragma unroll
        for (int q = 0; q < PKT_K / 2; q++) vrow[q] = sah_pixel_v(cam, py0 + q * 4);
#pragma unroll kPktUnroll
        for (int k = 0; k < PKT_K; k++) {
            const int px = px0 + (k & 1) * 8, py = py0 + (k >> 1) * 4;
            float dx, dy, dz;
            sah_primary_dir_uv(cam, (k & 1) ? u1 : u0, vrow[k >> 1], dx, dy, dz);
            const float ra = F_ADD(F_ADD(F_MUL(dx, dx), F_MUL(dy, dy)), F_MUL(dz, dz));
            s_ray[warp][k][0][lane] = dx; s_ray[warp][k][1][lane] = dy; s_ray[warp][k][2][lane] = dz; s_ray[warp][k][3][lane] = ra;
            s_bt[warp][k][lane] = 0.0f;
            s_bp[warp][k][lane] = -1;
            if (px < width && py < y1) {
                valid |= 1u << k;
                lminx = fminf(lminx, dx); lmaxx = fmaxf(lmaxx, dx);
                lminy = fminf(lminy, dy); lmaxy = fmaxf(lmaxy, dy);
                lminz = fminf(lminz, dz); lmaxz = fmaxf(lmaxz, dz);
                lmin_a = fminf(lmin_a, ra);
            }
        }
        lminx = warp_min(lminx); lminy = warp_min(lminy); lminz = warp_min(lminz);
        lmaxx = warp_max(lmaxx); lmaxy = warp_max(lmaxy); lmaxz = warp_max(lmaxz);
        lmin_a = warp_min(lmin_a);
        /* ---- tile frustum (warp-uniform) ---- */
        float ex = fmaxf(fabsf(cam.px - sc.lo[0]), fabsf(cam.px - sc.hi[0]));
        float ey = fmaxf(fabsf(cam.py - sc.lo[1]), fabsf(cam.py - sc.hi[1]));
        float ez = fmaxf(fabsf(cam.pz - sc.lo[2]), fabsf(cam.pz - sc.hi[2]));
        float pad = sqrtf(ex * ex + ey * ey + ez * ez) * 0.001953125f + sc.r_max * 0.001953125f;
        bool geo = sc.geometric_zero_axes != 0;
        SahFrustumAxis FX, FY, FZ;
        sah_frustum_axis_setup(FX, cam.px, lminx, lmaxx, pad, geo);
        sah_frustum_axis_setup(FY, cam.py, lminy, lmaxy, pad, geo);
        sah_frustum_axis_setup(FZ, cam.pz, lminz, lmaxz, pad, geo);
        float min_len = sqrtf(lmin_a);
        const float cull_abs = (min_len > 0.0f) ? (2.0f * sc.r_max) / min_len : INFINITY;
        float pkt_best_s = INFINITY;
        const bool simple = sah_frustum_simple(FX, FY, FZ);   /* warp-uniform */

        int top = 1;
        if (lane == 0) stack[0] = 0;
        __syncwarp();
        while (top > 0) {
            int npop = top < 8 ? top : 8;
            if (top > dfs_from) npop = 1; /* nearly full: plain DFS, growth <= 3 per level (see sahbvh_packet_dfs_from) */
            const int j = lane >> 2, c = lane & 3;
            const bool active = j < npop;
            int32_t nd = active ? stack[top - 1 - j] : 0;
            __syncwarp();
            top -= npop;
            if (COUNT && lane == 0) cnt.inner += (unsigned long long)npop;
            bool pass = false;
            float tn = INFINITY;
            int32_t ref = SAH_REF_NONE;
            if (active) {
                /* 128-byte 4-wide node (wide.c, full-precision twin): lane (j, c) tests child c of popped node j */
                const float4 *rec = reinterpret_cast<const float4 *>(nodes + (size_t)nd * 32 + c * 8);
                const float4 c0 = __ldg(rec), c1 = __ldg(rec + 1);
                ref = __float_as_int(c1.z);
                /* an absent child (SAH_REF_NONE) carries an inverted box: it fails the test unless every axis is ignored */
                pass = simple ? sah_frustum_box_simple(FX, FY, FZ, c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, pkt_best_s, tn)
                              : sah_frustum_box(FX, FY, FZ, c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, pkt_best_s, tn);
                pass = pass && ref != SAH_REF_NONE;
            }
            const bool is_leaf = ref < 0;
            const bool push = pass && !is_leaf;
            const bool cand = pass && is_leaf && (uint32_t)(~ref) < (uint32_t)sc.n_spheres;
            /* stack order: children of deeper stack entries first, of the top entry last; within a node far first,
             * near last (rank among the four siblings by packet entry distance, ties by slot) */
            /* tn >= 0 or +inf, so its bit pattern orders like the value; the two low bits carry the slot */
            const uint32_t key = (__float_as_uint(tn) & 0x7ffffffcu) | (uint32_t)c;
            int rank = 0;
#pragma unroll
            for (int x = 1; x < 4; x++) rank += (__shfl_xor_sync(full, key, x) < key) ? 1 : 0;
            const int order = (7 - j) * 4 + (3 - rank);
            unsigned ord_mask = __reduce_or_sync(full, push ? (1u << order) : 0u);
            if (push) stack[top + __popc(ord_mask & ((1u << order) - 1u))] = ref;
            top += __popc(ord_mask);
            if (COUNT && (unsigned long long)top > cnt.max_stack) cnt.max_stack = (unsigned long long)top;
            unsigned m = __ballot_sync(full, cand);
            __syncwarp();
            /* ---- candidates: every ray of the tile runs the reference's own tests ---- */
            bool changed = false;
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int32_t sidx = ~__shfl_sync(full, ref, src);
                const float4 s = __ldg(reinterpret_cast<const float4 *>(spheres4) + sidx);
                unsigned needk = 0;
#pragma unroll kPktUnroll
                for (int k = 0; k < PKT_K; k++) {
                    if (valid & (1u << k)) {
                        SahRay r;
                        r.ox = cam.px; r.oy = cam.py; r.oz = cam.pz;
                        r.dx = s_ray[warp][k][0][lane]; r.dy = s_ray[warp][k][1][lane]; r.dz = s_ray[warp][k][2][lane];
                        r.a = s_ray[warp][k][3][lane];
                        if (COUNT) cnt.spheres++;
                        float t;
                        if (sah_sphere_t(r, s.x, s.y, s.z, s.w, t)) {
                            const float bt = s_bt[warp][k][lane];
                            const int32_t bp = s_bp[warp][k][lane];
                            if (bp < 0 || t < bt || (t == bt && sidx > bp)) {
                                needk |= 1u << k;
                                s_tt[warp][k][lane] = t;
                            }
                        }
                    }
                }
                if (__any_sync(full, needk != 0)) {
                    /* the candidate's own box, re-read from its node (an L1 hit, same address for every lane) */
                    const float4 *brec = reinterpret_cast<const float4 *>(nodes + (size_t)__shfl_sync(full, nd, src) * 32 + (src & 3) * 8);
                    const float4 b0 = __ldg(brec), b1 = __ldg(brec + 1);
                    const float blx = b0.x, bly = b0.y, blz = b0.z, bhx = b0.w, bhy = b1.x, bhz = b1.y;
#pragma unroll kPktUnroll
                    for (int k = 0; k < PKT_K; k++) {
                        if (needk & (1u << k)) {
                            SahRay r;
                            r.ox = cam.px; r.oy = cam.py; r.oz = cam.pz;
                            r.dx = s_ray[warp][k][0][lane]; r.dy = s_ray[warp][k][1][lane]; r.dz = s_ray[warp][k][2][lane];
                            if (COUNT) cnt.exact++;
                            if (sah_box_exact_fast(r, blx, bly, blz, bhx, bhy, bhz)) {
                                s_bt[warp][k][lane] = s_tt[warp][k][lane];
                                s_bp[warp][k][lane] = sidx;
                                changed = true;
                            }
                        }
                    }
                }
            }
            if (__any_sync(full, changed)) {
                float w = 0.0f; /* worst (largest) cull bound over this lane's valid rays */
#pragma unroll
                for (int k = 0; k < PKT_K; k++)
                    if (valid & (1u << k))
                        w = fmaxf(w, s_bp[warp][k][lane] >= 0 ? F_FMA(s_bt[warp][k][lane], SAH_REL_CULL, cull_abs) : INFINITY);
                pkt_best_s = warp_max_nonneg(w);   /* w >= 0: t > 1e-6, cull_abs >= 0 */
            }
        }
    }
    /* ---- results: the block's 32 x (8 PKT_GROUP_ROWS) pixels leave through shared memory, one 32-pixel ROW per warp
     * instruction (128 contiguous bytes of ids, 1 KiB of records), whichever warp traced them -- the ids may be going to
     * another GPU's frame over NVLink or to pinned host memory over PCIe, where 32-byte segments waste the link. ---- */
#if PKT_WARP_OUT
    __syncwarp();
    const int out_rows = PKT_TILE_H / 2;               /* a warp instruction covers two 16-pixel rows of the warp's own tile */
#else
    __syncthreads();
    const int out_rows = PKT_TILE_H * PKT_GROUP_ROWS / PKT_WARPS;
#endif
#pragma unroll 1
    for (int q = 0; q < out_rows; q++) {
#if PKT_WARP_OUT
        const int r = (warp >> 1) * PKT_TILE_H + q * 2 + (lane >> 4); /* row of the block */
        const int bx = gx * 32 + (warp & 1) * 16 + (lane & 15);
        const int lx = (warp & 1) * 16 + (lane & 15);                /* column of the block */
#else
        const int r = warp * out_rows + q;
        const int bx = gx * 32 + lane;
        const int lx = lane;
#endif
        const int py = y0 + gy * (PKT_TILE_H * PKT_GROUP_ROWS) + r;
        if (py >= y1 || bx >= width) continue;
        /* pixel (lx, r) of the block was ray k of lane l2 of warp w2 (tile layout above) */
        const int w2 = (r >> 3) * 2 + (lx >> 4);
        const int k = (((r & 7) >> 2) << 1) | ((lx >> 3) & 1);
        const int l2 = ((r & 3) << 3) | (lx & 7);
        const int32_t bp = s_bp[w2][k][l2];
        const float bt = s_bt[w2][k][l2];
        const size_t pix = (size_t)(py - y0) * (size_t)width + (size_t)bx;
        if (ids_mode == SAH_IDS_PRIM) reinterpret_cast<int32_t *>(ids)[pix] = bp;
        else if (ids_mode == SAH_IDS_TPRIM) ids[pix] = make_float2(bp >= 0 ? bt : 0.0f, __int_as_float(bp));
        if (RECORDS) {
            SahHit h;
            if (bp >= 0) {
                SahRay ray;
                ray.ox = cam.px; ray.oy = cam.py; ray.oz = cam.pz;
                ray.dx = s_ray[w2][k][0][l2]; ray.dy = s_ray[w2][k][1][l2]; ray.dz = s_ray[w2][k][2][l2];
                const float4 s = __ldg(reinterpret_cast<const float4 *>(spheres4) + bp);
                sah_finish_hit(ray, bt, bp, s.x, s.y, s.z, h);
            } else {
                sah_miss(h);
            }
            store_hit256(out + pix, h);
        }
        if (COUNT) hits += (bp >= 0);
    }
    if (COUNT) flush_counters(cnt, hits, counters);
}

/*
 * K2m: packet traversal over MACRO tiles.  One warp (= one block) owns 32 x 8 pixels = S = 2 of K2p's 16 x 8 tiles side
 * by side, i.e. 8 rays per lane (the template also takes S = 1 and S = 4; 2 ships).  The interior culling -- half of
 * K2p's instructions -- runs ONCE per macro tile against the macro frustum.  A candidate leaf is then tested against the
 * frusta of the tile's FILTER UNITS -- the eight 8 x 4-pixel quadrants (QUAD: quadrant j is ray slot j of all 32 lanes;
 * lane j owns its frustum and tests it, one ballot), or the S sub-tiles -- each with its own cull bound, and a lane runs
 * the reference's sphere / leaf-box tests only for its ray slots whose unit passed: 2.3 sphere tests per ray on the
 * headline frame instead of K2p's 5.2.  Per ray the exact tests are the same as in K2p (best over leaves whose own tests
 * pass), so the records are identical; what is shared wider, and filtered finer, is only the conservative culling.
 * A macro tile is 32 pixels wide, so the warp's own rows are whole 128-byte id rows (1 KiB of records): results
 * leave through the warp's shared-memory state without a block barrier.
 */
#ifndef PKT_MACRO_MIN_BLOCKS
#define PKT_MACRO_MIN_BLOCKS 8 /* x 4 / PKT_MACRO_WARPS1 blocks: 64 registers and 6.3 KB of shared memory per warp = 31 warps per SM */
#endif
#ifndef PKT_MACRO_STACK
#define PKT_MACRO_STACK 256
#endif
#ifndef PKT_PREFETCH
#define PKT_PREFETCH 0 /* L1 prefetch of the next step's nodes before a candidate phase: measured 0.3785 ms against 0.3729 ms without */
#endif
#ifndef PKT_QUAD
#define PKT_QUAD true /* filter candidates per 8 x 4-pixel quadrant (8 units per macro tile) instead of per 16 x 8 sub-tile (2 units) */
#endif
#ifndef PKT_RAY_UNROLL
#define PKT_RAY_UNROLL 1
#endif
static constexpr int kPktRayUnroll = PKT_RAY_UNROLL;
static_assert(PKT_MACRO_STACK - 32 - 3 * (SAHBVH_STACK_MAX + 2) >= 8, "PKT_MACRO_STACK too small for the deepest admissible tree");
#ifndef PKT_MACRO_WARPS1
#define PKT_MACRO_WARPS1 1 /* warps per block: a macro tile per BLOCK (32 blocks per SM).  Nothing in the kernel is shared between
                             the warps of a block, and a block's slot is only free again when its slowest tile ends: 1 / 2 / 4
                             warps per block measured 0.362 / 0.365 / 0.374 ms on the 1 M-sphere 4K frame */
#endif
/* one filter unit (a sub-tile, or an 8 x 4-pixel quadrant): its frustum without the fields every unit shares (origin,
 * pad), and its own cull bound (max over its rays of best_s; -1 = no valid ray) */
struct PktSubFrustum {
    int mode[3];
    float ia[3], ib[3];
    float best_s;
};
__device__ __forceinline__ void pkt_sub_store(PktSubFrustum &d, const SahFrustumAxis &x, const SahFrustumAxis &y, const SahFrustumAxis &z, float best_s)
{
    d.mode[0] = x.mode; d.mode[1] = y.mode; d.mode[2] = z.mode;
    d.ia[0] = x.ia; d.ia[1] = y.ia; d.ia[2] = z.ia;
    d.ib[0] = x.ib; d.ib[1] = y.ib; d.ib[2] = z.ib;
    d.best_s = best_s;
}
__device__ __forceinline__ void pkt_sub_load(const PktSubFrustum &d, const SahCamera &cam, float pad, SahFrustumAxis &x, SahFrustumAxis &y, SahFrustumAxis &z)
{
    x.mode = d.mode[0]; x.o = cam.px; x.ia = d.ia[0]; x.ib = d.ib[0]; x.pad = pad;
    y.mode = d.mode[1]; y.o = cam.py; y.ia = d.ia[1]; y.ib = d.ib[1]; y.pad = pad;
    z.mode = d.mode[2]; z.o = cam.pz; z.ia = d.ia[2]; z.ib = d.ib[2]; z.pad = pad;
}

template <bool COUNT, bool RECORDS, int S, bool QUAD>
__global__ void __launch_bounds__(PKT_MACRO_WARPS1 * 32, PKT_MACRO_MIN_BLOCKS * 4 / PKT_MACRO_WARPS1)
k_trace_primary_macro(const float *__restrict__ nodes, const float *__restrict__ spheres4, SahScene sc,
                      SahCamera cam, int width, int y0, int y1, gpu_hit *__restrict__ out, float2 *__restrict__ ids,
                      int ids_mode, int dfs_from, unsigned long long *counters)
{
    static_assert(S == 1 || S == 2 || S == 4, "sub-tiles per macro tile");
    constexpr int SX = S >= 2 ? 2 : 1, SY = S / SX;   /* sub-tiles across / down; S = 1 is K2p's own 16 x 8 tile, one per block */
    constexpr int MW = PKT_MACRO_WARPS1;
    constexpr int TW = 16 * SX;          /* pixel columns of a macro tile */
    constexpr int MROWS = 8 * SY;        /* pixel rows of a macro tile */
    constexpr int NF = QUAD ? 4 * S : S; /* filter units: the 8 x 4-pixel quadrants (ray slot j of every lane) or the 16 x 8 sub-tiles */
    __shared__ int32_t s_stack[MW][PKT_MACRO_STACK];
    __shared__ float s_ray[MW][S][PKT_K][3][32];   /* dx, dy, dz (d.d is recomputed: shared memory decides the occupancy) */
    __shared__ float s_bt[MW][S][PKT_K][32];
    __shared__ int32_t s_bp[MW][S][PKT_K][32];
    __shared__ PktSubFrustum s_sub[MW][NF];
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mtiles_x = (width + TW - 1) / TW, mtiles_y = (y1 - y0 + MROWS - 1) / MROWS;
    SahCounters cnt = {0, 0, 0, 0};
    unsigned long long hits = 0;
    const long long mt = (long long)blockIdx.x * MW + warp;   /* macro tiles in row-major order */
    if (mt >= (long long)mtiles_x * mtiles_y) return;                      /* no block barrier in this kernel */
    const int ty = (int)(mt / mtiles_x), tx = (int)(mt % mtiles_x);
    int32_t *stack = s_stack[warp];
    /* ---- the lane's 4 S rays: sub-tile q = (qy, qx), ray k as in K2p:
     *      pixel (tx*32 + qx*16 + (k&1)*8 + lane%8,  y0 + ty*MROWS + qy*8 + (k>>1)*4 + lane/8) ---- */
    const int px0 = tx * TW + (lane & 7), py0 = y0 + ty * MROWS + (lane >> 3);
    unsigned valid = 0; /* bit q*4+k */
    /* the lane's rays use 4 distinct columns: one division each (main.c:362); the loops below stay ROLLED -- the
     * IEEE division / square-root sequences of one ray set-up, one sphere test and one leaf-box test are each
     * instantiated once, which keeps the kernel inside the instruction cache (ncu: no_instruction stalls) */
    const float u0 = sah_pixel_u(cam, px0), u1 = sah_pixel_u(cam, px0 + 8);
    const float u2 = SX == 2 ? sah_pixel_u(cam, px0 + 16) : 0.0f, u3 = SX == 2 ? sah_pixel_u(cam, px0 + 24) : 0.0f;
    float mminx = INFINITY, mminy = INFINITY, mminz = INFINITY, mmaxx = -INFINITY, mmaxy = -INFINITY, mmaxz = -INFINITY;
    float mmin_a = INFINITY;
    float ex = fmaxf(fabsf(cam.px - sc.lo[0]), fabsf(cam.px - sc.hi[0]));
    float ey = fmaxf(fabsf(cam.py - sc.lo[1]), fabsf(cam.py - sc.hi[1]));
    float ez = fmaxf(fabsf(cam.pz - sc.lo[2]), fabsf(cam.pz - sc.hi[2]));
    const float pad = sqrtf(ex * ex + ey * ey + ez * ez) * 0.001953125f + sc.r_max * 0.001953125f;
    const bool geo = sc.geometric_zero_axes != 0;
    bool sub_simple_acc = true;
    /* QUAD: lane j keeps the direction bounds of quadrant j (ray slot j of all lanes) */
    float fminx = INFINITY, fminy = INFINITY, fminz = INFINITY, fmaxx = -INFINITY, fmaxy = -INFINITY, fmaxz = -INFINITY;
    bool fany = false;
#pragma unroll 1
    for (int q = 0; q < S; q++) {
        const int qx = q % SX, qy = q / SX;
        float lminx = INFINITY, lminy = INFINITY, lminz = INFINITY, lmaxx = -INFINITY, lmaxy = -INFINITY, lmaxz = -INFINITY;
        float lmin_a = INFINITY;
#pragma unroll 1
        for (int k = 0; k < PKT_K; k++) {
            const int px = px0 + qx * 16 + (k & 1) * 8, py = py0 + qy * 8 + (k >> 1) * 4;
            float dx, dy, dz;
            const float u = qx ? ((k & 1) ? u3 : u2) : ((k & 1) ? u1 : u0);
            sah_primary_dir_uv(cam, u, sah_pixel_v(cam, py), dx, dy, dz);
            const float ra = F_ADD(F_ADD(F_MUL(dx, dx), F_MUL(dy, dy)), F_MUL(dz, dz));
            s_ray[warp][q][k][0][lane] = dx; s_ray[warp][q][k][1][lane] = dy; s_ray[warp][q][k][2][lane] = dz;
            s_bt[warp][q][k][lane] = 0.0f;
            s_bp[warp][q][k][lane] = -1;
            const bool v = px < width && py < y1;
            if (v) {
                valid |= 1u << (q * 4 + k);
                lmin_a = fminf(lmin_a, ra);
            }
            if (QUAD) {
                /* the quadrant's bounds: one reduction per component over this ray slot of all lanes */
                const float a0 = warp_min(v ? dx : INFINITY), a1 = warp_max(v ? dx : -INFINITY);
                const float b0 = warp_min(v ? dy : INFINITY), b1 = warp_max(v ? dy : -INFINITY);
                const float c0 = warp_min(v ? dz : INFINITY), c1 = warp_max(v ? dz : -INFINITY);
                const bool any = __any_sync(full, v);
                if (lane == q * 4 + k) { fminx = a0; fmaxx = a1; fminy = b0; fmaxy = b1; fminz = c0; fmaxz = c1; fany = any; }
                lminx = fminf(lminx, a0); lmaxx = fmaxf(lmaxx, a1);
                lminy = fminf(lminy, b0); lmaxy = fmaxf(lmaxy, b1);
                lminz = fminf(lminz, c0); lmaxz = fmaxf(lmaxz, c1);
            } else if (v) {
                lminx = fminf(lminx, dx); lmaxx = fmaxf(lmaxx, dx);
                lminy = fminf(lminy, dy); lmaxy = fmaxf(lmaxy, dy);
                lminz = fminf(lminz, dz); lmaxz = fmaxf(lmaxz, dz);
            }
        }
        lmin_a = warp_min(lmin_a);
        if (!QUAD) {
            lminx = warp_min(lminx); lminy = warp_min(lminy); lminz = warp_min(lminz);
            lmaxx = warp_max(lmaxx); lmaxy = warp_max(lmaxy); lmaxz = warp_max(lmaxz);
            SahFrustumAxis ax, ay, az;
            sah_frustum_axis_setup(ax, cam.px, lminx, lmaxx, pad, geo);
            sah_frustum_axis_setup(ay, cam.py, lminy, lmaxy, pad, geo);
            sah_frustum_axis_setup(az, cam.pz, lminz, lmaxz, pad, geo);
            /* an empty sub-tile (off the frame) has mode-3 axes, which accept every box: a negative bound keeps it out */
            const float bs = __any_sync(full, (valid >> (q * 4)) & 15u) ? INFINITY : -1.0f;
            if (lane == 0) pkt_sub_store(s_sub[warp][q], ax, ay, az, bs);
            sub_simple_acc = sub_simple_acc && (bs < 0.0f || sah_frustum_simple(ax, ay, az));
        }
        mminx = fminf(mminx, lminx); mmaxx = fmaxf(mmaxx, lmaxx);
        mminy = fminf(mminy, lminy); mmaxy = fmaxf(mmaxy, lmaxy);
        mminz = fminf(mminz, lminz); mmaxz = fmaxf(mmaxz, lmaxz);
        mmin_a = fminf(mmin_a, lmin_a);
    }
    if (QUAD) {
        /* lane j < NF sets up quadrant j's frustum (one instance of the set-up code for all of them) */
        SahFrustumAxis ax, ay, az;
        sah_frustum_axis_setup(ax, cam.px, fminx, fmaxx, pad, geo);
        sah_frustum_axis_setup(ay, cam.py, fminy, fmaxy, pad, geo);
        sah_frustum_axis_setup(az, cam.pz, fminz, fmaxz, pad, geo);
        const float bs = fany ? INFINITY : -1.0f;
        if (lane < NF) pkt_sub_store(s_sub[warp][lane], ax, ay, az, bs);
        sub_simple_acc = __all_sync(full, lane >= NF || bs < 0.0f || sah_frustum_simple(ax, ay, az));
    }
    const bool sub_simple = sub_simple_acc; /* warp-uniform: every (non-empty) filter unit's frustum has the branch-free form */
    /* ---- macro frustum (warp-uniform) ---- */
    SahFrustumAxis FX, FY, FZ;
    sah_frustum_axis_setup(FX, cam.px, mminx, mmaxx, pad, geo);
    sah_frustum_axis_setup(FY, cam.py, mminy, mmaxy, pad, geo);
    sah_frustum_axis_setup(FZ, cam.pz, mminz, mmaxz, pad, geo);
    const float min_len = sqrtf(mmin_a);
    const float cull_abs = (min_len > 0.0f) ? (2.0f * sc.r_max) / min_len : INFINITY;
    float pkt_best_s = INFINITY;
    const bool simple = sah_frustum_simple(FX, FY, FZ);

    int top = 1;
    if (lane == 0) stack[0] = 0;
    __syncwarp();
    while (top > 0) {
        int npop = top < 8 ? top : 8;
        if (top > dfs_from) npop = 1;
        const int j = lane >> 2, c = lane & 3;
        const bool active = j < npop;
        int32_t nd = active ? stack[top - 1 - j] : 0;
        __syncwarp();
        top -= npop;
        if (COUNT && lane == 0) cnt.inner += (unsigned long long)npop;
        bool pass = false;
        float tn = INFINITY;
        int32_t ref = SAH_REF_NONE;
        if (active) {
            const float4 *rec = reinterpret_cast<const float4 *>(nodes + (size_t)nd * 32 + c * 8);
            const float4 c0 = __ldg(rec), c1 = __ldg(rec + 1);
            ref = __float_as_int(c1.z);
            pass = simple ? sah_frustum_box_simple(FX, FY, FZ, c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, pkt_best_s, tn)
                          : sah_frustum_box(FX, FY, FZ, c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, pkt_best_s, tn);
            pass = pass && ref != SAH_REF_NONE;
        }
        const bool is_leaf = ref < 0;
        const bool push = pass && !is_leaf;
        const bool cand = pass && is_leaf && (uint32_t)(~ref) < (uint32_t)sc.n_spheres;
        const uint32_t key = (__float_as_uint(tn) & 0x7ffffffcu) | (uint32_t)c;
        int rank = 0;
#pragma unroll
        for (int x = 1; x < 4; x++) rank += (__shfl_xor_sync(full, key, x) < key) ? 1 : 0;
        const int order = (7 - j) * 4 + (3 - rank);
        unsigned ord_mask = __reduce_or_sync(full, push ? (1u << order) : 0u);
        if (push) stack[top + __popc(ord_mask & ((1u << order) - 1u))] = ref;
        top += __popc(ord_mask);
        if (COUNT && (unsigned long long)top > cnt.max_stack) cnt.max_stack = (unsigned long long)top;
        unsigned m = __ballot_sync(full, cand);
        __syncwarp();
#if PKT_PREFETCH
        if (m && top > 0) {
            /* the candidate phase below is long and does not touch the stack: start the next step's node fetches now
             * (the entries the next iteration pops are already in place), so that its loads hit L1 */
            const int np = top > dfs_from ? 1 : (top < 8 ? top : 8);
            if (j < np) {
                const float *pn = nodes + (size_t)stack[top - 1 - j] * 32 + c * 8;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pn));
            }
        }
#endif
        /* ---- candidates: sub-tile filter, then the reference's own tests for the rays of the sub-tiles it can touch ---- */
        bool any_changed = false;
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int32_t sidx = ~__shfl_sync(full, ref, src);
            const float4 s = __ldg(reinterpret_cast<const float4 *>(spheres4) + sidx);
            /* the candidate's own box, re-read from its node (an L1 hit, same address for every lane) */
            const float4 *brec = reinterpret_cast<const float4 *>(nodes + (size_t)__shfl_sync(full, nd, src) * 32 + (src & 3) * 8);
            const float4 b0 = __ldg(brec), b1 = __ldg(brec + 1);
            const float blx = b0.x, bly = b0.y, blz = b0.z, bhx = b0.w, bhy = b1.x, bhz = b1.y;
            unsigned fm = 1u;
            if (NF > 1) { /* compile-time: a one-tile block has nothing to filter (its frustum IS the traversal's) */
                bool sub_pass = false;
                if (lane < NF) {
                    SahFrustumAxis ax, ay, az;
                    pkt_sub_load(s_sub[warp][lane], cam, pad, ax, ay, az);
                    const float bs = s_sub[warp][lane].best_s;
                    float tq;
                    sub_pass = sub_simple ? sah_frustum_box_simple(ax, ay, az, blx, bly, blz, bhx, bhy, bhz, bs, tq)
                                          : sah_frustum_box(ax, ay, az, blx, bly, blz, bhx, bhy, bhz, bs, tq);
                }
                fm = __ballot_sync(full, sub_pass) & ((1u << NF) - 1u);
            }
            /* the lane's rays to test: every ray slot of a passing sub-tile, or the slots of the passing quadrants */
            unsigned vq = valid;
            if (QUAD) {
                vq &= fm;
            } else {
                unsigned em = 0;
#pragma unroll
                for (int q = 0; q < S; q++) em |= ((fm >> q) & 1u) ? (15u << (q * 4)) : 0u;
                vq &= em;
            }
            /* the reference's sphere test and, for a would-be winner, its own leaf-box test, ray by ray (the box is
             * already here for the filter, so there is no reason to batch the box tests as K2p does) */
            unsigned chg = 0;
            const float *rq = &s_ray[warp][0][0][0][lane];   /* ray slot j: rq[j*96], rq[j*96+32], rq[j*96+64] */
            float *btq = &s_bt[warp][0][0][lane];            /* ray slot j: btq[j*32] */
            int32_t *bpq = &s_bp[warp][0][0][lane];
#pragma unroll kPktRayUnroll
            for (; vq; vq &= vq - 1u) {
                const int jr = __ffs(vq) - 1;
                SahRay r;
                r.ox = cam.px; r.oy = cam.py; r.oz = cam.pz;
                r.dx = rq[jr * 96]; r.dy = rq[jr * 96 + 32]; r.dz = rq[jr * 96 + 64];
                r.a = F_ADD(F_ADD(F_MUL(r.dx, r.dx), F_MUL(r.dy, r.dy)), F_MUL(r.dz, r.dz));
                if (COUNT) cnt.spheres++;
                float t;
                if (sah_sphere_t(r, s.x, s.y, s.z, s.w, t)) {
                    const float bt = btq[jr * 32];
                    const int32_t bp = bpq[jr * 32];
                    if (bp < 0 || t < bt || (t == bt && sidx > bp)) {
                        if (COUNT) cnt.exact++;
                        if (sah_box_exact_fast(r, blx, bly, blz, bhx, bhy, bhz)) {
                            btq[jr * 32] = t;
                            bpq[jr * 32] = sidx;
                            chg |= 1u << jr;
                        }
                    }
                }
            }
            /* new cull bounds of the filter units that changed */
            unsigned cm = __reduce_or_sync(full, chg);
            if (!QUAD) { /* ray slots -> sub-tiles */
                unsigned um = 0;
#pragma unroll
                for (int q = 0; q < S; q++) um |= ((cm >> (q * 4)) & 15u) ? (1u << q) : 0u;
                cm = um;
            }
            while (cm) {
                const int un = __ffs(cm) - 1;
                cm &= cm - 1;
                float w = 0.0f; /* worst (largest) cull bound over this lane's valid rays of the unit */
                if (QUAD) {
                    if ((valid >> un) & 1u) w = bpq[un * 32] >= 0 ? F_FMA(btq[un * 32], SAH_REL_CULL, cull_abs) : INFINITY;
                } else {
#pragma unroll
                    for (int k = 0; k < PKT_K; k++)
                        if (valid & (1u << (un * 4 + k)))
                            w = fmaxf(w, bpq[(un * 4 + k) * 32] >= 0 ? F_FMA(btq[(un * 4 + k) * 32], SAH_REL_CULL, cull_abs) : INFINITY);
                }
                w = warp_max_nonneg(w);
                if (lane == 0) s_sub[warp][un].best_s = w;
                any_changed = true;
            }
            __syncwarp(); /* the next candidate's filter reads the bounds just written */
        }
        if (any_changed) { /* warp-uniform: the traversal's bound is the largest unit bound (an empty unit's -1 counts as 0) */
            pkt_best_s = warp_max_nonneg(lane < NF ? fmaxf(s_sub[warp][lane].best_s, 0.0f) : 0.0f);
        }
    }
    /* ---- results: the warp's own rows (32 pixels = one 128-byte id row per store; a 16-pixel tile stores two rows) ---- */
    __syncwarp();
    constexpr int RPS = 32 / TW; /* rows per warp store */
#pragma unroll 1
    for (int r0 = 0; r0 < MROWS; r0 += RPS) {
        const int r = r0 + (RPS == 2 ? (lane >> 4) : 0), lx = RPS == 2 ? (lane & 15) : lane;
        const int py = y0 + ty * MROWS + r, bx = tx * TW + lx;
        if (py >= y1 || bx >= width) continue;
        const int q = (r >> 3) * SX + (lx >> 4);
        const int k = (((r & 7) >> 2) << 1) | ((lx >> 3) & 1);
        const int l2 = ((r & 3) << 3) | (lx & 7);
        const int32_t bp = s_bp[warp][q][k][l2];
        const float bt = s_bt[warp][q][k][l2];
        const size_t pix = (size_t)(py - y0) * (size_t)width + (size_t)bx;
        if (ids_mode == SAH_IDS_PRIM) reinterpret_cast<int32_t *>(ids)[pix] = bp;
        else if (ids_mode == SAH_IDS_TPRIM) ids[pix] = make_float2(bp >= 0 ? bt : 0.0f, __int_as_float(bp));
        if (RECORDS) {
            SahHit h;
            if (bp >= 0) {
                SahRay ray;
                ray.ox = cam.px; ray.oy = cam.py; ray.oz = cam.pz;
                ray.dx = s_ray[warp][q][k][0][l2]; ray.dy = s_ray[warp][q][k][1][l2]; ray.dz = s_ray[warp][q][k][2][l2];
                const float4 s = __ldg(reinterpret_cast<const float4 *>(spheres4) + bp);
                sah_finish_hit(ray, bt, bp, s.x, s.y, s.z, h);
            } else {
                sah_miss(h);
            }
            store_hit256(out + pix, h);
        }
        if (COUNT) hits += (bp >= 0);
    }
    if (COUNT) flush_counters(cnt, hits, counters);
}

__global__ void __launch_bounds__(PRIMARY_THREADS)
k_primary_rays(SahCamera cam, int width, int y0, int y1, float *__restrict__ out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int y = y0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (x < width && y < y1) {
        float dx, dy, dz;
        sah_primary_dir(cam, x, y, dx, dy, dz);
        float *o = out + ((size_t)(y - y0) * (size_t)width + (size_t)x) * 6;
        o[0] = cam.px; o[1] = cam.py; o[2] = cam.pz;
        o[3] = dx; o[4] = dy; o[5] = dz;
    }
}

/* renderer.c:33-44: closest over ALL spheres, strict < so the lowest index wins ties */
__global__ void __launch_bounds__(TRACE_THREADS)
k_brute_force(const float *__restrict__ spheres4, SahScene sc, const float *__restrict__ rays, long long n,
              gpu_hit *__restrict__ out)
{
    __shared__ float4 tile[TRACE_THREADS];
    const int n_spheres = sc.n_spheres;
    long long i = (long long)blockIdx.x * TRACE_THREADS + threadIdx.x;
    bool live = i < n;
    SahRay r;
    if (live) {
        const float *rp = rays + i * 6;
        sah_ray_setup(r, rp[0], rp[1], rp[2], rp[3], rp[4], rp[5], sc);
    } else {
        sah_ray_setup(r, 0, 0, 0, 0, 0, 0, sc);
    }
    float best_t = INFINITY;
    int best = -1;
    for (int base = 0; base < n_spheres; base += TRACE_THREADS) {
        int j = base + threadIdx.x;
        tile[threadIdx.x] = j < n_spheres ? reinterpret_cast<const float4 *>(spheres4)[j]
                                          : make_float4(0, 0, 0, __int_as_float(0x7fffffff));
        __syncthreads();
        int m = min(TRACE_THREADS, n_spheres - base);
        if (live) {
            for (int k = 0; k < m; k++) {
                float4 s = tile[k];
                float t;
                if (sah_sphere_t(r, s.x, s.y, s.z, s.w, t) && t < best_t) { best_t = t; best = base + k; }
            }
        }
        __syncthreads();
    }
    if (live) {
        SahHit h;
        if (best >= 0) {
            float4 s = reinterpret_cast<const float4 *>(spheres4)[best];
            sah_finish_hit(r, best_t, best, s.x, s.y, s.z, h);
        } else {
            sah_miss(h);
        }
        store_hit(out + i, h);
    }
}

/*
 * K4: benchmark.c's two counting loops.
 *  k_any_hit_flags  benchmark_no_bvh (benchmark.c:172-223): every ray against EVERY sphere, no early
 *                   exit, "hit" = ray_sphere_intersect(...).hit_something (hit.c:27-29); a ray counts
 *                   once (any-hit counting, not closest hit).  grid.y splits the spheres so
 *                   that 10 000 rays still fill the device; per-ray flags are OR-ed in global memory.
 *  k_count_flags / k_count_prims  the final `intersections++` reductions (flags, or prim >= 0 of a
 *                   hit buffer for benchmark_with_bvh, benchmark.c:225-254).
 */
#define ANYHIT_SPHERES_PER_BLOCK 4096

__global__ void __launch_bounds__(TRACE_THREADS)
k_any_hit_flags(const float *__restrict__ spheres4, SahScene sc, const float *__restrict__ rays, long long n,
                unsigned int *__restrict__ flags)
{
    __shared__ float4 tile[TRACE_THREADS];
    const long long i = (long long)blockIdx.x * TRACE_THREADS + threadIdx.x;
    const bool live = i < n;
    SahRay r;
    if (live) {
        const float *rp = rays + i * 6;
        sah_ray_setup(r, rp[0], rp[1], rp[2], rp[3], rp[4], rp[5], sc);
    } else {
        sah_ray_setup(r, 0, 0, 0, 0, 0, 0, sc);
    }
    const int s0 = blockIdx.y * ANYHIT_SPHERES_PER_BLOCK;
    const int s1 = min(sc.n_spheres, s0 + ANYHIT_SPHERES_PER_BLOCK);
    bool any = false;
    for (int base = s0; base < s1; base += TRACE_THREADS) {
        const int j = base + threadIdx.x;
        tile[threadIdx.x] = j < s1 ? reinterpret_cast<const float4 *>(spheres4)[j]
                                   : make_float4(0, 0, 0, __int_as_float(0x7fffffff));
        __syncthreads();
        const int m = min(TRACE_THREADS, s1 - base);
        if (live) {
            for (int k = 0; k < m; k++) {
                const float4 s = tile[k];
                float t;
                any |= sah_sphere_t(r, s.x, s.y, s.z, s.w, t);
            }
        }
        __syncthreads();
    }
    if (live && any) atomicOr(&flags[i], 1u);
}

__global__ void __launch_bounds__(256) k_count_flags(const unsigned int *__restrict__ flags, long long n, unsigned long long *count)
{
    unsigned long long c = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) c += flags[i] != 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

__global__ void __launch_bounds__(256) k_count_prims(const gpu_hit *__restrict__ hits, long long n, unsigned long long *count)
{
    unsigned long long c = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) c += hits[i].prim >= 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

/* HBM-streaming reduction (the frame consumer of the multi-GPU loop runs it on 4 bytes per pixel): 128-bit loads,
 * four in flight per thread; the unaligned head and the tail go through the scalar path */
__global__ void __launch_bounds__(256) k_count_prim_ids(const int32_t *__restrict__ prims, long long n, unsigned long long *count)
{
    unsigned c = 0;
    const long long tid = (long long)blockIdx.x * 256 + threadIdx.x, nthr = (long long)gridDim.x * 256;
    long long head = (long long)((16 - ((uintptr_t)prims & 15)) & 15) / 4;
    if (head > n) head = n;
    const int4 *v = reinterpret_cast<const int4 *>(prims + head);
    const long long nv = (n - head) / 4;
    long long i = tid;
    for (; i + 3 * nthr < nv; i += 4 * nthr) {
        const int4 a = __ldcs(v + i), b = __ldcs(v + i + nthr), d = __ldcs(v + i + 2 * nthr), e = __ldcs(v + i + 3 * nthr);
        c += (a.x >= 0) + (a.y >= 0) + (a.z >= 0) + (a.w >= 0) + (b.x >= 0) + (b.y >= 0) + (b.z >= 0) + (b.w >= 0);
        c += (d.x >= 0) + (d.y >= 0) + (d.z >= 0) + (d.w >= 0) + (e.x >= 0) + (e.y >= 0) + (e.z >= 0) + (e.w >= 0);
    }
    for (; i < nv; i += nthr) {
        const int4 a = __ldcs(v + i);
        c += (a.x >= 0) + (a.y >= 0) + (a.z >= 0) + (a.w >= 0);
    }
    if (tid < head) c += prims[tid] >= 0;
    const long long tail0 = head + nv * 4;
    if (tail0 + tid < n) c += prims[tail0 + tid] >= 0;   /* < 4 elements */
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
}

/* ------------------------------------------------------------------------------------------ */
/* handle                                                                                       */
/* ------------------------------------------------------------------------------------------ */

extern "C" int sahbvh_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" uint64_t sahbvh_kernel_launches(void) { return (uint64_t)g_launches.load(); }

extern "C" int gpu_bvh_from_flat(const flat_bvh *f, int device, gpu_bvh **out)
{
    if (!f || !out) {
        sahbvh_set_error("gpu_bvh_from_flat: NULL argument");
        return SAHBVH_E_ARG;
    }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        sahbvh_set_error("no usable CUDA device (%s); this library has no CPU traversal path",
                         e == cudaSuccess ? "device count is 0" : cu_errstr(e));
        cudaGetLastError();
        return SAHBVH_E_CUDA;
    }
    if (device < 0) CU_TRY(cudaGetDevice(&device));
    if (device >= ndev) {
        sahbvh_set_error("device %d out of range (%d devices)", device, ndev);
        return SAHBVH_E_ARG;
    }
    DeviceGuard guard(device);
    if (!guard.ok) {
        sahbvh_set_error("cudaSetDevice(%d) failed", device);
        return SAHBVH_E_CUDA;
    }
    gpu_bvh *b = (gpu_bvh *)calloc(1, sizeof(gpu_bvh));
    if (!b) return SAHBVH_E_NOMEM;
    b->device = device;
    b->info = f->info;
    b->info.device = device;
    wide_bvh wb;
    int wrc = wide_bvh_from_flat(&wb, f);
    if (wrc != SAHBVH_OK) {
        free(b);
        return wrc;
    }
    b->n_wide = wb.n_wide;
    b->n_special = wb.n_special;
    b->info.n_wide = wb.n_wide;
    const size_t wide_bytes = (size_t)wb.n_wide * 64, special_bytes = (size_t)(wb.n_special > 0 ? wb.n_special : 1) * 32;
    const bool keep_binary = g_keep_binary.load() != 0;
    size_t node_bytes = keep_binary ? (size_t)f->info.n_inner * 64 : 0;
    size_t sph_bytes = ((size_t)f->info.n_spheres + 1) * 16;
    size_t col_bytes = ((size_t)f->info.n_spheres + 1) * 4;
    cudaError_t err = cudaSuccess;
    do {
        if (keep_binary && (err = cudaMalloc(&b->d_nodes, node_bytes)) != cudaSuccess) break;
        if ((err = cudaMalloc(&b->d_spheres4, sph_bytes)) != cudaSuccess) break;
        if ((err = cudaMalloc(&b->d_colors, col_bytes)) != cudaSuccess) break;
        if ((err = cudaMalloc(&b->d_counters, 72 * sizeof(unsigned long long))) != cudaSuccess) break;
        if ((err = cudaMalloc(&b->d_wide, wide_bytes)) != cudaSuccess) break;
        if ((err = cudaMalloc(&b->d_special, special_bytes)) != cudaSuccess) break;
        if ((err = cudaMalloc(&b->d_wide_full, wide_bytes * 2)) != cudaSuccess) break;
        if ((err = cudaMemcpy(b->d_wide_full, wb.full, wide_bytes * 2, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((err = cudaMemcpy(b->d_wide, wb.nodes, wide_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((err = cudaMemcpy(b->d_special, wb.special, special_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (keep_binary && (err = cudaMemcpy(b->d_nodes, f->nodes, node_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((err = cudaMemcpy(b->d_spheres4, f->spheres4, sph_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((err = cudaMemcpy(b->d_colors, f->colors, col_bytes, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((err = cudaStreamCreateWithFlags(&b->streams[0], cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((err = cudaStreamCreateWithFlags(&b->streams[1], cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((err = cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) break;
    } while (0);
    wide_bvh_release(&wb);
    if (err != cudaSuccess) {
        sahbvh_set_error("gpu_bvh upload failed: %s", cu_errstr(err));
        gpu_bvh_free(b);
        return SAHBVH_E_CUDA;
    }
    b->info.device_bytes = (int64_t)(node_bytes + sph_bytes + col_bytes + 3 * wide_bytes + special_bytes);
    *out = b;
    return SAHBVH_OK;
}

/* handle around arrays the device builder left on the GPU (takes ownership of the pointers) */
static int gpu_bvh_from_parts(gpu_bvh_parts *p, int device, gpu_bvh **out)
{
    *out = nullptr;
    if (device < 0) CU_TRY(cudaGetDevice(&device));
    DeviceGuard guard(device);
    if (!guard.ok) {
        sahbvh_set_error("cudaSetDevice(%d) failed", device);
        return SAHBVH_E_CUDA;
    }
    gpu_bvh *b = (gpu_bvh *)calloc(1, sizeof(gpu_bvh));
    if (!b) return SAHBVH_E_NOMEM;
    b->device = device;
    b->info = p->info;
    b->info.device = device;
    b->d_nodes = p->d_nodes; b->d_spheres4 = p->d_spheres4; b->d_colors = p->d_colors;
    if (b->d_nodes && g_keep_binary.load() == 0) {
        /* the binary tree was the collapse's input; the default kernels never read it */
        cudaFree(b->d_nodes);
        b->d_nodes = nullptr;
        b->info.device_bytes -= (int64_t)b->info.n_inner * 64;
    }
    b->d_wide = p->d_wide; b->d_wide_full = p->d_wide_full; b->d_special = p->d_special;
    b->n_wide = p->n_wide; b->n_special = p->n_special;
    memset(p, 0, sizeof *p);
    cudaError_t err = cudaSuccess;
    do {
        if ((err = cudaMalloc(&b->d_counters, 72 * sizeof(unsigned long long))) != cudaSuccess) break;
        if ((err = cudaStreamCreateWithFlags(&b->streams[0], cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((err = cudaStreamCreateWithFlags(&b->streams[1], cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((err = cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) break;
    } while (0);
    if (err != cudaSuccess) {
        sahbvh_set_error("gpu_bvh set-up failed: %s", cu_errstr(err));
        gpu_bvh_free(b);
        return SAHBVH_E_CUDA;
    }
    *out = b;
    return SAHBVH_OK;
}

/* Diagnostic download of the handle's device arrays (tests compare the device-built arrays with the host-built
 * ones byte for byte).  what: 0 binary nodes, 1 spheres4, 2 colours, 3 wide nodes, 4 wide full-precision twins,
 * 5 special-leaf records.  Returns the array's size in bytes; copies min(size, cap) bytes when out != NULL. */
extern "C" int64_t gpu_bvh_download(const gpu_bvh *b, int what, void *out, int64_t cap)
{
    if (!b) return SAHBVH_E_ARG;
    const void *src = nullptr;
    int64_t bytes = 0;
    switch (what) {
    case 0: src = b->d_nodes; bytes = b->d_nodes ? (int64_t)b->info.n_inner * 64 : 0; break;
    case 1: src = b->d_spheres4; bytes = ((int64_t)b->info.n_spheres + 1) * 16; break;
    case 2: src = b->d_colors; bytes = ((int64_t)b->info.n_spheres + 1) * 4; break;
    case 3: src = b->d_wide; bytes = (int64_t)b->n_wide * 64; break;
    case 4: src = b->d_wide_full; bytes = (int64_t)b->n_wide * 128; break;
    case 5: src = b->d_special; bytes = (int64_t)b->n_special * 32; break;
    default: sahbvh_set_error("gpu_bvh_download: unknown array %d", what); return SAHBVH_E_ARG;
    }
    if (out && bytes > 0) {
        DeviceGuard guard(b->device);
        const int64_t n = bytes < cap ? bytes : cap;
        if (n > 0 && cudaMemcpy(out, src, (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) {
            sahbvh_set_error("gpu_bvh_download: copy failed");
            return SAHBVH_E_CUDA;
        }
    }
    return bytes;
}

extern "C" int gpu_bvh_upload(const Sphere *permuted, int n, const BVHNode *root, int device, gpu_bvh **out)
{
    if (!out) {
        sahbvh_set_error("gpu_bvh_upload: NULL out");
        return SAHBVH_E_ARG;
    }
    *out = nullptr;
    sah_tree t;
    int rc = sah_tree_init(&t, 2 * (int64_t)(n > 0 ? n : 1) + 16);
    if (rc != SAHBVH_OK) return rc;
    rc = sah_tree_from_pointers(&t, root, permuted, n);
    flat_bvh f;
    memset(&f, 0, sizeof f);
    if (rc == SAHBVH_OK) rc = flat_bvh_from_tree(&f, &t, permuted, n);
    if (rc == SAHBVH_OK) rc = gpu_bvh_from_flat(&f, device, out);
    flat_bvh_release(&f);
    sah_tree_free(&t);
    return rc;
}

extern "C" int gpu_bvh_build(Sphere *spheres, int n_total, int start, int end, int depth, int device, gpu_bvh **out)
{
    if (!out || !spheres || start < 0 || end < start || end > n_total) {
        sahbvh_set_error("gpu_bvh_build: bad arguments");
        return SAHBVH_E_ARG;
    }
    *out = nullptr;
    if (g_build_mode.load() == SAHBVH_BUILD_DEVICE && end - start >= 4096) {
        /* build, flatten and collapse on the device: only the permuted spheres come back */
        if (sahbvh_device_count() == 0) {
            sahbvh_set_error("no usable CUDA device; this library has no CPU traversal path");
            return SAHBVH_E_CUDA;
        }
        gpu_bvh_parts parts;
        int prc = gpu_build_parts(spheres, n_total, start, end, depth, device, &parts);
        if (prc == SAHBVH_OK) {
            prc = gpu_bvh_from_parts(&parts, device, out);
            if (prc != SAHBVH_OK) gpu_bvh_parts_release(&parts);
            return prc;
        }
        if (prc != 1) return prc;
        /* 1: input the device path does not take (degenerate tree, depth limits): the host path below decides */
    }
    sah_tree t;
    int rc = sah_tree_init(&t, 2 * (int64_t)(end - start + 1) + 16);
    if (rc != SAHBVH_OK) return rc;
    rc = sah_build(&t, spheres, start, end, depth);
    flat_bvh f;
    memset(&f, 0, sizeof f);
    if (rc == SAHBVH_OK) rc = flat_bvh_from_tree(&f, &t, spheres, n_total);
    if (rc == SAHBVH_OK) rc = gpu_bvh_from_flat(&f, device, out);
    flat_bvh_release(&f);
    sah_tree_free(&t);
    return rc;
}

static_assert(sizeof(sahbvh_tree_node) == sizeof(sah_node), "public tree node == internal tree node");

extern "C" int64_t sahbvh_build_tree(Sphere *spheres, int start, int end, int depth, int mode, int device,
                                     sahbvh_tree_node *out, int64_t cap)
{
    if (!spheres || !out || start < 0 || end < start || cap < 0 || (mode != SAHBVH_BUILD_HOST && mode != SAHBVH_BUILD_DEVICE)) {
        sahbvh_set_error("sahbvh_build_tree: bad arguments");
        return SAHBVH_E_ARG;
    }
    if (mode == SAHBVH_BUILD_DEVICE && sahbvh_device_count() == 0) {
        sahbvh_set_error("no usable CUDA device");
        return SAHBVH_E_CUDA;
    }
    sah_tree t;
    int rc = sah_tree_init(&t, 2 * (int64_t)(end - start + 1) + 16);
    if (rc != SAHBVH_OK) return rc;
    rc = mode == SAHBVH_BUILD_DEVICE ? sah_build_gpu(&t, spheres, start, end, depth, device) : sah_build(&t, spheres, start, end, depth);
    int64_t used = rc == SAHBVH_OK ? t.used : (int64_t)rc;
    if (rc == SAHBVH_OK && used <= cap) memcpy(out, t.nodes, (size_t)used * sizeof(sah_node));
    sah_tree_free(&t);
    return used;
}

extern "C" void gpu_bvh_free(gpu_bvh *b)
{
    if (!b) return;
    DeviceGuard guard(b->device);
    for (int k = 0; k < 2; k++) {
        if (b->streams[k]) cudaStreamDestroy(b->streams[k]);
        if (b->d_in[k]) cudaFree(b->d_in[k]);
        if (b->d_out[k]) cudaFree(b->d_out[k]);
    }
    if (b->d_nodes) cudaFree(b->d_nodes);
    if (b->d_spheres4) cudaFree(b->d_spheres4);
    if (b->d_colors) cudaFree(b->d_colors);
    if (b->d_counters) cudaFree(b->d_counters);
    if (b->d_wide) cudaFree(b->d_wide);
    if (b->d_special) cudaFree(b->d_special);
    if (b->d_wide_full) cudaFree(b->d_wide_full);
    if (b->d_render) cudaFree(b->d_render);
    for (int i = 0; i < SAHBVH_TUNE_SLOTS; i++)
        for (int k = 0; k < 2; k++)
            for (int sidx = 0; sidx < 2; sidx++)
                for (int e = 0; e < 2; e++)
                    if (b->tune[i].ev[k][sidx][e]) cudaEventDestroy(b->tune[i].ev[k][sidx][e]);
    free(b);
}

extern "C" int gpu_bvh_get_info(const gpu_bvh *b, gpu_bvh_info *info)
{
    if (!b || !info) {
        sahbvh_set_error("gpu_bvh_get_info: NULL argument");
        return SAHBVH_E_ARG;
    }
    *info = b->info;
    return SAHBVH_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* launches                                                                                     */
/* ------------------------------------------------------------------------------------------ */

SahScene sahbvh_make_scene(const gpu_bvh_info &info)
{
    SahScene sc;
    sc.n_spheres = info.n_spheres;
    sc.r_max = info.max_radius;
    for (int a = 0; a < 3; a++) {
        sc.lo[a] = info.scene_lo[a];
        sc.hi[a] = info.scene_hi[a];
    }
    sc.geometric_zero_axes = info.empty_leaves == 0;
    return sc;
}

/* north_star (2) "top BVH levels in ... L2-persisting windows": the wide trees are numbered breadth first, so their
 * top levels are a PREFIX of the node array.  With sahbvh_set_l2_persist_bytes(B > 0) the trace launches run with an
 * access-policy window over the first B bytes of the tree they read (hit ratio 1, persisting; everything else
 * streaming) and the device sets aside as much L2.  Off by default: measured in DESIGN.md section 4 (it only matters
 * when the tree is larger than the 126 MB L2, i.e. at 10 M spheres). */
static std::atomic<size_t> g_l2_persist{0};
extern "C" int sahbvh_set_l2_persist_bytes(uint64_t bytes)
{
    g_l2_persist.store((size_t)bytes);
    return SAHBVH_OK;
}
struct L2Window {
    cudaStream_t st;
    bool on;
    L2Window(gpu_bvh *b, cudaStream_t s, const void *base, size_t tree_bytes) : st(s), on(false)
    {
        size_t want = g_l2_persist.load();
        if (want == 0 || !base) {
            if (b->l2_limit_applied) { /* switched off again: give the set-aside back */
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
                cudaCtxResetPersistingL2Cache();
                b->l2_limit_applied = 0;
            }
            return;
        }
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, b->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, b->device);
        if (want > tree_bytes) want = tree_bytes;
        if (want > (size_t)max_persist) want = (size_t)max_persist;
        if (want > (size_t)max_window) want = (size_t)max_window;
        if (want == 0) return;
        if (b->l2_limit_applied != want) {
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { (void)cudaGetLastError(); return; }
            b->l2_limit_applied = want;
        }
        cudaStreamAttrValue v;
        memset(&v, 0, sizeof v);
        v.accessPolicyWindow.base_ptr = const_cast<void *>(base);
        v.accessPolicyWindow.num_bytes = want;
        v.accessPolicyWindow.hitRatio = 1.0f;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) { (void)cudaGetLastError(); return; }
        on = true;
    }
    ~L2Window()
    {
        if (!on) return;
        cudaStreamAttrValue v;
        memset(&v, 0, sizeof v);
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v); /* num_bytes 0: later launches of the stream are unaffected */
    }
};

static int trace_grid(const gpu_bvh *b, long long n)
{
    long long blocks = (n + TRACE_THREADS - 1) / TRACE_THREADS;
    long long cap = (long long)b->sm_count * 64; /* grid-stride beyond that */
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int sahbvh_launch_rays(gpu_bvh *b, const void *d_rays, long long n, void *d_out, cudaStream_t st, bool count, void *d_prims)
{
    if (n == 0) return SAHBVH_OK;
    if (!d_out && (!d_prims || count || g_ray_mode.load() != SAHBVH_RAYS_WIDE)) {
        sahbvh_set_error("ray-buffer launch without a record buffer: the 4-byte output needs SAHBVH_RAYS_WIDE");
        return SAHBVH_E_ARG;
    }
    const SahScene sc = sahbvh_make_scene(b->info);
    if (g_ray_mode.load() == SAHBVH_RAYS_WIDE) {
        /* persistent grid over the 4-wide tree: every resident warp pulls 128-ray chunks */
        if (b->wide_blocks_per_sm == 0) {
            int nb = 0;
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_rays_wide<false>, WIDE_WARPS * 32, 0));
            b->wide_blocks_per_sm = nb > 0 ? nb : 1;
        }
        long long want = (n + WIDE_CHUNK - 1) / WIDE_CHUNK;
        long long blocks = (want + WIDE_WARPS - 1) / WIDE_WARPS;
        long long cap = (long long)b->sm_count * b->wide_blocks_per_sm;
        if (blocks > cap) blocks = cap;
        unsigned long long *ctr = b->d_counters + 8 + (b->next_counter++ & 63);
        CU_TRY(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st));
        L2Window win(b, st, b->d_wide, (size_t)b->n_wide * 64);
        if (count)
            k_trace_rays_wide<true><<<(unsigned)blocks, WIDE_WARPS * 32, 0, st>>>(b->d_wide, b->n_wide, b->d_spheres4, b->d_special, b->n_special, sc,
                                                                                 (const float *)d_rays, n, (gpu_hit *)d_out, nullptr, ctr, b->d_counters);
        else
            k_trace_rays_wide<false><<<(unsigned)blocks, WIDE_WARPS * 32, 0, st>>>(b->d_wide, b->n_wide, b->d_spheres4, b->d_special, b->n_special, sc,
                                                                                  (const float *)d_rays, n, (gpu_hit *)d_out,
                                                                                  d_out ? nullptr : (int32_t *)d_prims, ctr, nullptr);
    } else if (g_ray_mode.load() == SAHBVH_RAYS_COOP) {
        /* persistent grid: every resident warp pulls 128-ray chunks from a per-launch counter */
        int brc = sahbvh_need_binary_tree(b);
        if (brc != SAHBVH_OK) return brc;
        if (b->coop_blocks_per_sm == 0) {
            int nb = 0;
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_rays_coop<false>, COOP_WARPS * 32, 0));
            b->coop_blocks_per_sm = nb > 0 ? nb : 1;
        }
        long long want = (n + COOP_CHUNK - 1) / COOP_CHUNK;               /* one chunk per warp at least */
        long long blocks = (want + COOP_WARPS - 1) / COOP_WARPS;
        long long cap = (long long)b->sm_count * b->coop_blocks_per_sm;
        if (blocks > cap) blocks = cap;
        unsigned long long *ctr = b->d_counters + 8 + (b->next_counter++ & 63);
        CU_TRY(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st));
        if (count)
            k_trace_rays_coop<true><<<(unsigned)blocks, COOP_WARPS * 32, 0, st>>>(b->d_nodes, b->d_spheres4, sc, (const float *)d_rays,
                                                                                 n, (gpu_hit *)d_out, ctr, b->d_counters);
        else
            k_trace_rays_coop<false><<<(unsigned)blocks, COOP_WARPS * 32, 0, st>>>(b->d_nodes, b->d_spheres4, sc, (const float *)d_rays,
                                                                                  n, (gpu_hit *)d_out, ctr, nullptr);
    } else {
        int brc = sahbvh_need_binary_tree(b);
        if (brc != SAHBVH_OK) return brc;
        int grid = trace_grid(b, n);
        if (count)
            k_trace_rays<true><<<grid, TRACE_THREADS, 0, st>>>(b->d_nodes, b->d_spheres4, sc,
                                                              (const float *)d_rays, n, (gpu_hit *)d_out, b->d_counters);
        else
            k_trace_rays<false><<<grid, TRACE_THREADS, 0, st>>>(b->d_nodes, b->d_spheres4, sc,
                                                               (const float *)d_rays, n, (gpu_hit *)d_out, nullptr);
    }
    g_launches++;
    CU_TRY(cudaGetLastError());
    return SAHBVH_OK;
}

void sahbvh_make_camera(const Camera *cam, int width, int height, SahCamera *c)
{
    /* ray.c:18-24, evaluated once per frame on the host with the reference's double/float mix */
    Camera tmp = *cam;
    float aspect = (float)width / (float)height;
    float fov_rad = (float)((double)tmp.fov * (M_PI / (double)180.0f));
    float half_h = (float)tan((double)(fov_rad / 2.0f));
    float half_w = aspect * half_h;
    Vec3 horizontal = vec3_multiply(tmp.right, 2.0f * half_w);
    Vec3 vertical = vec3_multiply(tmp.up, 2.0f * half_h);
    c->px = tmp.position.x; c->py = tmp.position.y; c->pz = tmp.position.z;
    c->fx = tmp.forward.x; c->fy = tmp.forward.y; c->fz = tmp.forward.z;
    c->hx = horizontal.x; c->hy = horizontal.y; c->hz = horizontal.z;
    c->vx = vertical.x; c->vy = vertical.y; c->vz = vertical.z;
    c->aspect = aspect;
    c->fw = (float)width;
    c->fh = (float)height;
}

/* SAHBVH_PRIMARY_AUTO.  K2p (16 x 8 tiles) and K2m (32 x 8 macro tiles with a sub-tile filter) return identical
 * records; which is faster depends on the scene and the frame (K2m: 7-19 % faster on the 5 K - 1 M sphere 4K frames, K2p
 * 12 % faster on the 10 M one).  Per handle and frame shape the first launches
 * alternate between the two with an event pair around each; once two samples of each have completed the faster one
 * (minimum of its samples) is used from then on.  Nothing synchronises: the events are only queried.  Instrumented
 * (counting) launches use the decision when there is one and K2p otherwise. */
static PrimaryTune *primary_tune_find(gpu_bvh *b, int width, int rows, bool create)
{
    for (int i = 0; i < b->n_tune; i++)
        if (b->tune[i].width == width && b->tune[i].rows == rows) return &b->tune[i];
    if (!create) return nullptr;
    PrimaryTune *t = &b->tune[b->n_tune < SAHBVH_TUNE_SLOTS ? b->n_tune++ : (b->tune_evict++ % SAHBVH_TUNE_SLOTS)];
    t->width = width;
    t->rows = rows;
    t->decided = -1;
    t->launches = 0;
    return t;
}

static int primary_tune_pick(gpu_bvh *b, int width, int rows, bool count, cudaStream_t st, cudaEvent_t *start, cudaEvent_t *stop)
{
    /* a stream that is being captured into a graph cannot be sampled (the events would become graph nodes and never
     * report a time): such launches use what is known and leave the tuner alone */
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    const bool capturing = cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone;
    if (capturing) (void)cudaGetLastError();
    PrimaryTune *t = primary_tune_find(b, width, rows, !count && !capturing);
    if (!t) return SAHBVH_PRIMARY_PACKET;
    if (t->decided >= 0 || count || capturing) return t->decided >= 0 ? t->decided : SAHBVH_PRIMARY_PACKET;
    const int n = t->launches++;
    if (n >= 6) {
        /* samples 1 and 2 of each kernel (launches 2..5; 0 and 1 were warm-up): decide once all of them have completed */
        float ms[2] = {1e30f, 1e30f};
        bool ready = true;
        for (int k = 0; k < 2 && ready; k++)
            for (int sidx = 0; sidx < 2 && ready; sidx++) {
                float v = 0.0f;
                const cudaError_t e = cudaEventElapsedTime(&v, t->ev[k][sidx][0], t->ev[k][sidx][1]);
                if (e == cudaSuccess) ms[k] = v < ms[k] ? v : ms[k];
                else { (void)cudaGetLastError(); ready = false; }
            }
        if (ready) t->decided = ms[1] < ms[0] ? SAHBVH_PRIMARY_MACRO : SAHBVH_PRIMARY_PACKET;
        else if (n >= 64) t->decided = SAHBVH_PRIMARY_PACKET; /* events never completed (capturing stream?): stop trying */
        return t->decided >= 0 ? t->decided : SAHBVH_PRIMARY_PACKET;
    }
    const int k = n & 1;
    if (n >= 2) {
        const int sidx = (n >> 1) - 1;
        for (int e = 0; e < 2; e++)
            if (!t->ev[k][sidx][e] && cudaEventCreate(&t->ev[k][sidx][e]) != cudaSuccess) {
                (void)cudaGetLastError();
                t->decided = SAHBVH_PRIMARY_PACKET; /* no events, no tuning */
                return t->decided;
            }
        *start = t->ev[k][sidx][0];
        *stop = t->ev[k][sidx][1];
    }
    return k ? SAHBVH_PRIMARY_MACRO : SAHBVH_PRIMARY_PACKET;
}

/* 0 per-ray, 1 K2p, 2 K2m: the kernel gpu_trace_primary* uses for this handle and frame shape right now; -1 while
 * SAHBVH_PRIMARY_AUTO is still sampling */
extern "C" int gpu_primary_kernel_in_use(const gpu_bvh *b, int width, int rows)
{
    if (!b) return SAHBVH_E_ARG;
    const int mode = g_primary_mode.load();
    if (mode != SAHBVH_PRIMARY_AUTO) return mode;
    const PrimaryTune *t = primary_tune_find(const_cast<gpu_bvh *>(b), width, rows, false);
    return t ? t->decided : -1;
}

int sahbvh_launch_primary(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, void *d_out,
                          cudaStream_t st, bool count, void *d_ids, int ids_mode)
{
    if (y1 <= y0 || width <= 0) return SAHBVH_OK;
    if (!d_ids) ids_mode = SAH_IDS_NONE;
    if (!d_out && (ids_mode == SAH_IDS_NONE || count)) {
        sahbvh_set_error("primary launch without an output buffer");
        return SAHBVH_E_ARG;
    }
    SahCamera c;
    sahbvh_make_camera(cam, width, height, &c);
    gpu_hit *out = (gpu_hit *)d_out;
    float2 *ids = (float2 *)d_ids;
    const SahScene sc = sahbvh_make_scene(b->info);
    /* which kernel: fixed by the mode, or -- SAHBVH_PRIMARY_AUTO -- the faster of the two packet kernels for this
     * handle and frame shape, found by timing the first launches (primary_tune_pick) */
    int kernel = g_primary_mode.load();
    cudaEvent_t tune_start = nullptr, tune_stop = nullptr;
    if (kernel == SAHBVH_PRIMARY_AUTO) kernel = primary_tune_pick(b, width, y1 - y0, count, st, &tune_start, &tune_stop);
    if (tune_start) cudaEventRecord(tune_start, st);
    if (kernel == SAHBVH_PRIMARY_PACKET || kernel == SAHBVH_PRIMARY_MACRO) {
        /* one CTA per 32 x 16 pixel group (2 x 2 tiles of 16 x 8) */
        const long long groups = (long long)((width + 31) / 32) * ((y1 - y0 + PKT_GROUP_ROWS * PKT_TILE_H - 1) / (PKT_GROUP_ROWS * PKT_TILE_H));
        if (groups >= (1ll << 31) - 1) {
            sahbvh_set_error("frame too large for one launch (%lld tile groups): trace it in row bands", groups);
            return SAHBVH_E_ARG;
        }
        const unsigned grid = (unsigned)groups;
        const int dfs_from = sahbvh_packet_dfs_from(b->info.max_depth);
        L2Window win(b, st, b->d_wide_full, (size_t)b->n_wide * 128);
        if (kernel == SAHBVH_PRIMARY_MACRO) {
            /* macro tiles: one warp per 32 x 8 pixels (S = 2 sub-tiles), PKT_MACRO_WARPS1 of them per block in row-major order.
             * The template also instantiates for S = 1 (one 16 x 8 tile per block: 0.418 ms on the headline frame against
             * K2p's 0.391) and S = 4 (32 x 16: 0.408 ms, shared memory limits it to 18 warps per SM); only S = 2 ships. */
            const int dfs_from = sahbvh_packet_dfs_from(b->info.max_depth, PKT_MACRO_STACK);
            const long long mtiles = (long long)((width + 31) / 32) * ((y1 - y0 + 7) / 8);
            const int mw = PKT_MACRO_WARPS1;
            const unsigned mgrid = (unsigned)((mtiles + mw - 1) / mw);
            if (count)
                k_trace_primary_macro<true, true, 2, PKT_QUAD><<<mgrid, mw * 32, 0, st>>>(b->d_wide_full, b->d_spheres4, sc, c, width, y0, y1,
                                                                               out, ids, ids_mode, dfs_from, b->d_counters);
            else if (out)
                k_trace_primary_macro<false, true, 2, PKT_QUAD><<<mgrid, mw * 32, 0, st>>>(b->d_wide_full, b->d_spheres4, sc, c, width, y0, y1,
                                                                                out, ids, ids_mode, dfs_from, nullptr);
            else
                k_trace_primary_macro<false, false, 2, PKT_QUAD><<<mgrid, mw * 32, 0, st>>>(b->d_wide_full, b->d_spheres4, sc, c, width, y0, y1,
                                                                                 nullptr, ids, ids_mode, dfs_from, nullptr);
        } else if (count)
            k_trace_primary_packet<true, true><<<grid, PKT_WARPS * 32, 0, st>>>(b->d_wide_full, b->d_spheres4, sc, c, width, y0, y1,
                                                                               out, ids, ids_mode, dfs_from, b->d_counters);
        else if (out)
            k_trace_primary_packet<false, true><<<grid, PKT_WARPS * 32, 0, st>>>(b->d_wide_full, b->d_spheres4, sc, c, width, y0, y1,
                                                                                out, ids, ids_mode, dfs_from, nullptr);
        else
            k_trace_primary_packet<false, false><<<grid, PKT_WARPS * 32, 0, st>>>(b->d_wide_full, b->d_spheres4, sc, c, width, y0, y1,
                                                                                 nullptr, ids, ids_mode, dfs_from, nullptr);
    } else {
        int rc = sahbvh_need_binary_tree(b);
        if (rc != SAHBVH_OK) return rc;
        dim3 grid((width + 31) / 32, (y1 - y0 + 7) / 8);
        if (count)
            k_trace_primary<true><<<grid, PRIMARY_THREADS, 0, st>>>(b->d_nodes, b->d_spheres4, sc, c, width, y0, y1,
                                                                   out, ids, ids_mode, b->d_counters);
        else
            k_trace_primary<false><<<grid, PRIMARY_THREADS, 0, st>>>(b->d_nodes, b->d_spheres4, sc, c, width, y0, y1,
                                                                    out, ids, ids_mode, nullptr);
    }
    if (tune_stop) cudaEventRecord(tune_stop, st);
    g_launches++;
    CU_TRY(cudaGetLastError());
    return SAHBVH_OK;
}

static int check_frame(const gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, const void *out)
{
    if (!b || !cam || !out || width <= 0 || height <= 0 || y0 < 0 || y1 < y0 || y1 > height) {
        sahbvh_set_error("bad frame arguments (width=%d height=%d rows=[%d,%d))", width, height, y0, y1);
        return SAHBVH_E_ARG;
    }
    return SAHBVH_OK;
}

int sahbvh_ensure_staging(gpu_bvh *b, size_t in_bytes, size_t out_bytes)
{
    if (in_bytes > b->in_cap) {
        for (int k = 0; k < 2; k++) {
            if (b->d_in[k]) cudaFree(b->d_in[k]);
            b->d_in[k] = nullptr;
        }
        b->in_cap = 0;
        for (int k = 0; k < 2; k++) CU_TRY(cudaMalloc(&b->d_in[k], in_bytes));
        b->in_cap = in_bytes;
    }
    if (out_bytes > b->out_cap) {
        for (int k = 0; k < 2; k++) {
            if (b->d_out[k]) cudaFree(b->d_out[k]);
            b->d_out[k] = nullptr;
        }
        b->out_cap = 0;
        for (int k = 0; k < 2; k++) CU_TRY(cudaMalloc(&b->d_out[k], out_bytes));
        b->out_cap = out_bytes;
    }
    return SAHBVH_OK;
}

#define HOST_CHUNK_RAYS (1ll << 22) /* 4 Mi rays: 96 MiB in, 128 MiB out per chunk (measured best of 1/2/4/8 Mi on 2^24 bounce rays) */
static long long host_chunk_rays()
{
    static long long v = 0;
    if (!v) {
        const char *e = getenv("SAHBVH_HOST_CHUNK_RAYS"); /* tuning aid */
        v = e ? atoll(e) : HOST_CHUNK_RAYS;
        if (v < 1024) v = HOST_CHUNK_RAYS;
    }
    return v;
}

extern "C" int gpu_trace_closest_hit_dev(gpu_bvh *b, const void *d_rays, int64_t n, void *d_out, void *stream)
{
    if (!b || n < 0 || (n > 0 && (!d_rays || !d_out))) {
        sahbvh_set_error("gpu_trace_closest_hit_dev: bad arguments");
        return SAHBVH_E_ARG;
    }
    DeviceGuard guard(b->device);
    return sahbvh_launch_rays(b, d_rays, n, d_out, (cudaStream_t)stream, false);
}

/* error exit of a host-buffer call whose earlier chunks may still be copying from / into the caller's arrays: let
 * them finish before the error returns (the caller may free the arrays right away) */
static int fail_after_drain(gpu_bvh *b, int rc)
{
    cudaStreamSynchronize(b->streams[0]);
    cudaStreamSynchronize(b->streams[1]);
    (void)cudaGetLastError();
    return rc;
}
#define CU_TRY_DRAIN(b, expr)                                                                     \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            sahbvh_set_error("%s failed: %s (%s:%d)", #expr, cu_errstr(_e), __FILE__, __LINE__); \
            return fail_after_drain((b), SAHBVH_E_CUDA);                                          \
        }                                                                                         \
    } while (0)

static int trace_rays_host(gpu_bvh *b, const Ray *rays, int64_t n, void *out, bool prims_only)
{
    if (!b || n < 0 || (n > 0 && (!rays || !out))) {
        sahbvh_set_error("gpu_trace_closest_hit: bad arguments");
        return SAHBVH_E_ARG;
    }
    if (n == 0) return SAHBVH_OK;
    DeviceGuard guard(b->device);
    const size_t rec = prims_only ? sizeof(int32_t) : sizeof(gpu_hit);
    long long chunk = n < host_chunk_rays() ? n : host_chunk_rays();
    int rc = sahbvh_ensure_staging(b, (size_t)chunk * sizeof(Ray), (size_t)chunk * rec);
    if (rc != SAHBVH_OK) return rc;
    int k = 0;
    for (long long off = 0; off < n; off += chunk, k ^= 1) {
        long long m = n - off < chunk ? n - off : chunk;
        cudaStream_t st = b->streams[k];
        CU_TRY_DRAIN(b, cudaMemcpyAsync(b->d_in[k], rays + off, (size_t)m * sizeof(Ray), cudaMemcpyHostToDevice, st));
        rc = prims_only ? sahbvh_launch_rays(b, b->d_in[k], m, nullptr, st, false, b->d_out[k])
                        : sahbvh_launch_rays(b, b->d_in[k], m, b->d_out[k], st, false);
        if (rc != SAHBVH_OK) return fail_after_drain(b, rc);
        CU_TRY_DRAIN(b, cudaMemcpyAsync((char *)out + (size_t)off * rec, b->d_out[k], (size_t)m * rec, cudaMemcpyDeviceToHost, st));
        /* the staging pair k is reused two chunks later: the stream orders it */
    }
    CU_TRY(cudaStreamSynchronize(b->streams[0]));
    CU_TRY(cudaStreamSynchronize(b->streams[1]));
    return SAHBVH_OK;
}

extern "C" int gpu_trace_closest_hit(gpu_bvh *b, const Ray *rays, int64_t n, gpu_hit *out)
{
    return trace_rays_host(b, rays, n, out, false);
}

extern "C" int gpu_trace_closest_hit_prims(gpu_bvh *b, const Ray *rays, int64_t n, int32_t *prims)
{
    return trace_rays_host(b, rays, n, prims, true);
}

extern "C" int gpu_trace_closest_hit_dev_prims(gpu_bvh *b, const void *d_rays, int64_t n, void *d_prims, void *stream)
{
    if (!b || n < 0 || (n > 0 && (!d_rays || !d_prims))) {
        sahbvh_set_error("gpu_trace_closest_hit_dev_prims: bad arguments");
        return SAHBVH_E_ARG;
    }
    DeviceGuard guard(b->device);
    return sahbvh_launch_rays(b, d_rays, n, nullptr, (cudaStream_t)stream, false, d_prims);
}

extern "C" int gpu_trace_primary_dev(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1,
                                     void *d_out, void *stream)
{
    int rc = check_frame(b, cam, width, height, y0, y1, d_out);
    if (rc != SAHBVH_OK) return rc;
    DeviceGuard guard(b->device);
    return sahbvh_launch_primary(b, cam, width, height, y0, y1, d_out, (cudaStream_t)stream, false, nullptr, 0);
}

extern "C" int gpu_trace_primary_dev_ids(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1,
                                         void *d_out, void *d_ids, void *stream)
{
    int rc = check_frame(b, cam, width, height, y0, y1, d_out ? d_out : d_ids);
    if (rc != SAHBVH_OK) return rc;
    DeviceGuard guard(b->device);
    return sahbvh_launch_primary(b, cam, width, height, y0, y1, d_out, (cudaStream_t)stream, false, d_ids, SAH_IDS_TPRIM);
}

extern "C" int gpu_trace_primary_dev_prims(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1,
                                           void *d_out, void *d_prims, void *stream)
{
    int rc = check_frame(b, cam, width, height, y0, y1, d_out ? d_out : d_prims);
    if (rc != SAHBVH_OK) return rc;
    DeviceGuard guard(b->device);
    return sahbvh_launch_primary(b, cam, width, height, y0, y1, d_out, (cudaStream_t)stream, false, d_prims, SAH_IDS_PRIM);
}

/* prim ids of image rows -> full hit records: t, point and normal are pure functions of the pixel's ray and the
 * hit sphere (hit.c:19-39 evaluated once more on the winner), so the result is bit-identical to the traced record */
__global__ void __launch_bounds__(PRIMARY_THREADS)
k_hits_from_prims(const float *__restrict__ spheres4, SahScene sc, SahCamera cam, int width, int y0, int y1,
                  const int32_t *__restrict__ prims, gpu_hit *__restrict__ out)
{
    const long long pix = (long long)blockIdx.x * PRIMARY_THREADS + threadIdx.x;
    if (pix >= (long long)(y1 - y0) * width) return;
    SahHit h;
    sah_miss(h);
    const int32_t prim = prims[pix];
    if (prim >= 0 && prim < sc.n_spheres) {
        float dx, dy, dz;
        sah_primary_dir(cam, (int)(pix % width), y0 + (int)(pix / width), dx, dy, dz);
        SahRay r;
        r.ox = cam.px; r.oy = cam.py; r.oz = cam.pz;
        r.dx = dx; r.dy = dy; r.dz = dz;
        r.a = F_ADD(F_ADD(F_MUL(dx, dx), F_MUL(dy, dy)), F_MUL(dz, dz));
        const float4 s = __ldg(reinterpret_cast<const float4 *>(spheres4) + prim);
        float t;
        if (sah_sphere_t(r, s.x, s.y, s.z, s.w, t)) sah_finish_hit(r, t, prim, s.x, s.y, s.z, h);
    }
    store_hit(out + pix, h);
}

extern "C" int gpu_hits_from_prims_dev(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1,
                                       const void *d_prims, void *d_out, void *stream)
{
    int rc = check_frame(b, cam, width, height, y0, y1, d_out);
    if (rc != SAHBVH_OK) return rc;
    if (!d_prims) {
        sahbvh_set_error("gpu_hits_from_prims_dev: NULL prims");
        return SAHBVH_E_ARG;
    }
    if (y1 == y0) return SAHBVH_OK;
    DeviceGuard guard(b->device);
    SahCamera c;
    sahbvh_make_camera(cam, width, height, &c);
    const long long n = (long long)(y1 - y0) * width;
    k_hits_from_prims<<<(unsigned)((n + PRIMARY_THREADS - 1) / PRIMARY_THREADS), PRIMARY_THREADS, 0, (cudaStream_t)stream>>>(
        b->d_spheres4, sahbvh_make_scene(b->info), c, width, y0, y1, (const int32_t *)d_prims, (gpu_hit *)d_out);
    g_launches++;
    CU_TRY(cudaGetLastError());
    return SAHBVH_OK;
}

/* ---- host-buffer frames.  fmt: what goes back over PCIe per pixel ---- */

static long long env_ll(const char *name, long long dflt, long long lo)
{
    const char *e = getenv(name);
    if (!e) return dflt;
    const long long v = atoll(e);
    return v < lo ? dflt : v;
}

/* SAHBVH_HOST_DIRECT (tuning aid): 1 = when the caller's buffer is page-locked, the kernel stores its rows straight
 * into it through the host mapping (no staging buffer, no copy engine); 0 = staged bands + cudaMemcpyAsync. */
static int host_direct_mode()
{
    static int v = -1;
    if (v < 0) v = (int)env_ll("SAHBVH_HOST_DIRECT", SAHBVH_HOST_DIRECT_DEFAULT, 0);
    return v;
}

/* device-visible alias of a page-locked host buffer, or NULL (pageable memory / no unified addressing) */
static void *mapped_alias(const void *host_ptr)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return at.devicePointer;
}

/* enqueue only: kernels and copies are in flight on the handle's two streams when this returns;
 * sahbvh_primary_host_wait() blocks until the caller's buffer is complete.  (Split so that one host thread can
 * drive several devices' bands concurrently, multi_gpu.cu.) */
int sahbvh_primary_host_enqueue(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, void *out, int fmt)
{
    int rc = check_frame(b, cam, width, height, y0, y1, out);
    if (rc != SAHBVH_OK) return rc;
    if (y1 == y0) return SAHBVH_OK;
    DeviceGuard guard(b->device);
    const size_t px_bytes = fmt == HOSTFMT_RECORDS ? sizeof(gpu_hit) : fmt == HOSTFMT_PRIMS ? 4 : 8;
    const int ids_mode = fmt == HOSTFMT_PRIMS ? SAH_IDS_PRIM : fmt == HOSTFMT_TPRIM ? SAH_IDS_TPRIM : SAH_IDS_NONE;
    if (host_direct_mode() && fmt != HOSTFMT_RECORDS) {
        void *alias = mapped_alias(out);
        if (alias) /* one launch for the whole frame; its 128/256-byte row stores are the PCIe writes */
            return sahbvh_launch_primary(b, cam, width, height, y0, y1, nullptr, b->streams[0], false, alias, ids_mode);
    }
    /* row bands, alternating between two streams, so the D2H copy of one band overlaps the next band's kernel:
     * ~4 Mi pixels per band for 32-byte records (128 MiB), ~1 Mi pixels for the compact formats (4-8 MiB) */
    const long long band_px = fmt == HOSTFMT_RECORDS ? HOST_CHUNK_RAYS : env_ll("SAHBVH_PRIMARY_BAND_PIXELS", 1ll << 20, 4096);
    int band = (int)(band_px / width);
    band = band < 16 ? 16 : (band / 16) * 16;         /* whole 32 x 16 pixel CTA groups */
    if (band > y1 - y0) band = ((y1 - y0 + 15) / 16) * 16;
    rc = sahbvh_ensure_staging(b, 0, (size_t)band * (size_t)width * px_bytes);
    if (rc != SAHBVH_OK) return rc;
    int k = 0;
    for (int ya = y0; ya < y1; ya += band, k ^= 1) {
        const int yb = ya + band < y1 ? ya + band : y1;
        cudaStream_t st = b->streams[k];
        if (fmt == HOSTFMT_RECORDS) rc = sahbvh_launch_primary(b, cam, width, height, ya, yb, b->d_out[k], st, false, nullptr, SAH_IDS_NONE);
        else rc = sahbvh_launch_primary(b, cam, width, height, ya, yb, nullptr, st, false, b->d_out[k], ids_mode);
        if (rc == SAHBVH_OK) {
            const cudaError_t e = cudaMemcpyAsync((char *)out + (size_t)(ya - y0) * (size_t)width * px_bytes, b->d_out[k],
                                                  (size_t)(yb - ya) * (size_t)width * px_bytes, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) {
                sahbvh_set_error("gpu_trace_primary: copy to the host buffer failed: %s", cu_errstr(e));
                rc = SAHBVH_E_CUDA;
            }
        }
        if (rc != SAHBVH_OK) return fail_after_drain(b, rc);
    }
    return SAHBVH_OK;
}

int sahbvh_primary_host_wait(gpu_bvh *b)
{
    DeviceGuard guard(b->device);
    CU_TRY(cudaStreamSynchronize(b->streams[0]));
    CU_TRY(cudaStreamSynchronize(b->streams[1]));
    return SAHBVH_OK;
}

static int trace_primary_host(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, void *out, int fmt)
{
    int rc = sahbvh_primary_host_enqueue(b, cam, width, height, y0, y1, out, fmt);
    if (rc != SAHBVH_OK || !b || y1 == y0) return rc;
    return sahbvh_primary_host_wait(b);
}

extern "C" int gpu_trace_primary(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, gpu_hit *out)
{
    return trace_primary_host(b, cam, width, height, y0, y1, out, HOSTFMT_RECORDS);
}

extern "C" int gpu_trace_primary_prims(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, int32_t *prims)
{
    return trace_primary_host(b, cam, width, height, y0, y1, prims, HOSTFMT_PRIMS);
}

extern "C" int gpu_trace_primary_tprim(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, gpu_tprim *out)
{
    return trace_primary_host(b, cam, width, height, y0, y1, out, HOSTFMT_TPRIM);
}

/* host id frame -> host records, on demand (H2D 4 B/pixel, k_hits_from_prims, D2H 32 B/pixel, banded over two streams) */
extern "C" int gpu_hits_from_prims(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1,
                                   const int32_t *prims, gpu_hit *out)
{
    int rc = check_frame(b, cam, width, height, y0, y1, out);
    if (rc != SAHBVH_OK) return rc;
    if (!prims) {
        sahbvh_set_error("gpu_hits_from_prims: NULL prims");
        return SAHBVH_E_ARG;
    }
    if (y1 == y0) return SAHBVH_OK;
    DeviceGuard guard(b->device);
    int band = (int)(HOST_CHUNK_RAYS / width);
    band = band < 1 ? 1 : band;
    if (band > y1 - y0) band = y1 - y0;
    rc = sahbvh_ensure_staging(b, (size_t)band * (size_t)width * 4, (size_t)band * (size_t)width * sizeof(gpu_hit));
    if (rc != SAHBVH_OK) return rc;
    int k = 0;
    for (int ya = y0; ya < y1; ya += band, k ^= 1) {
        const int yb = ya + band < y1 ? ya + band : y1;
        const size_t off = (size_t)(ya - y0) * (size_t)width, npx = (size_t)(yb - ya) * (size_t)width;
        cudaStream_t st = b->streams[k];
        CU_TRY_DRAIN(b, cudaMemcpyAsync(b->d_in[k], prims + off, npx * 4, cudaMemcpyHostToDevice, st));
        rc = gpu_hits_from_prims_dev(b, cam, width, height, ya, yb, b->d_in[k], b->d_out[k], st);
        if (rc != SAHBVH_OK) return fail_after_drain(b, rc);
        CU_TRY_DRAIN(b, cudaMemcpyAsync(out + off, b->d_out[k], npx * sizeof(gpu_hit), cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaStreamSynchronize(b->streams[0]));
    CU_TRY(cudaStreamSynchronize(b->streams[1]));
    return SAHBVH_OK;
}

extern "C" int gpu_generate_primary_rays(const Camera *cam, int width, int height, int y0, int y1, Ray *out, int device)
{
    if (!cam || !out || width <= 0 || height <= 0 || y0 < 0 || y1 < y0 || y1 > height) {
        sahbvh_set_error("gpu_generate_primary_rays: bad arguments");
        return SAHBVH_E_ARG;
    }
    if (y1 == y0) return SAHBVH_OK;
    int ndev = sahbvh_device_count();
    if (ndev == 0) {
        sahbvh_set_error("no usable CUDA device");
        return SAHBVH_E_CUDA;
    }
    if (device < 0) CU_TRY(cudaGetDevice(&device));
    DeviceGuard guard(device);
    SahCamera c;
    sahbvh_make_camera(cam, width, height, &c);
    size_t bytes = (size_t)(y1 - y0) * (size_t)width * sizeof(Ray);
    float *d = nullptr;
    CU_TRY(cudaMalloc(&d, bytes));
    dim3 grid((width + 31) / 32, (y1 - y0 + 7) / 8);
    k_primary_rays<<<grid, PRIMARY_THREADS>>>(c, width, y0, y1, d);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d, bytes, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) {
        sahbvh_set_error("gpu_generate_primary_rays: %s", cu_errstr(e));
        return SAHBVH_E_CUDA;
    }
    return SAHBVH_OK;
}

extern "C" int gpu_trace_brute_force(gpu_bvh *b, const Ray *rays, int64_t n, gpu_hit *out)
{
    if (!b || n < 0 || (n > 0 && (!rays || !out))) {
        sahbvh_set_error("gpu_trace_brute_force: bad arguments");
        return SAHBVH_E_ARG;
    }
    if (n == 0) return SAHBVH_OK;
    DeviceGuard guard(b->device);
    int rc = sahbvh_ensure_staging(b, (size_t)n * sizeof(Ray), (size_t)n * sizeof(gpu_hit));
    if (rc != SAHBVH_OK) return rc;
    cudaStream_t st = b->streams[0];
    CU_TRY(cudaMemcpyAsync(b->d_in[0], rays, (size_t)n * sizeof(Ray), cudaMemcpyHostToDevice, st));
    int grid = (int)((n + TRACE_THREADS - 1) / TRACE_THREADS);
    k_brute_force<<<grid, TRACE_THREADS, 0, st>>>(b->d_spheres4, sahbvh_make_scene(b->info), (const float *)b->d_in[0], n,
                                                 (gpu_hit *)b->d_out[0]);
    g_launches++;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out, b->d_out[0], (size_t)n * sizeof(gpu_hit), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return SAHBVH_OK;
}

/* benchmark.c:172-223 on the device: rays in, number of rays that hit ANY of the handle's spheres out */
extern "C" int gpu_benchmark_no_bvh(gpu_bvh *b, const Ray *rays, int64_t n, int64_t *intersections)
{
    if (!b || n < 0 || !intersections || (n > 0 && !rays)) {
        sahbvh_set_error("gpu_benchmark_no_bvh: bad arguments");
        return SAHBVH_E_ARG;
    }
    *intersections = 0;
    if (n == 0 || b->info.n_spheres == 0) return SAHBVH_OK;
    DeviceGuard guard(b->device);
    int rc = sahbvh_ensure_staging(b, (size_t)n * sizeof(Ray), (size_t)n * sizeof(gpu_hit));
    if (rc != SAHBVH_OK) return rc;
    cudaStream_t st = b->streams[0];
    unsigned int *flags = (unsigned int *)b->d_out[0];
    unsigned long long *cnt = b->d_counters + 7;
    CU_TRY(cudaMemcpyAsync(b->d_in[0], rays, (size_t)n * sizeof(Ray), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(flags, 0, (size_t)n * sizeof(unsigned int), st));
    CU_TRY(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), st));
    dim3 grid((unsigned)((n + TRACE_THREADS - 1) / TRACE_THREADS),
              (unsigned)((b->info.n_spheres + ANYHIT_SPHERES_PER_BLOCK - 1) / ANYHIT_SPHERES_PER_BLOCK));
    k_any_hit_flags<<<grid, TRACE_THREADS, 0, st>>>(b->d_spheres4, sahbvh_make_scene(b->info), (const float *)b->d_in[0], n, flags);
    int cblocks = (int)((n + 255) / 256);
    if (cblocks > b->sm_count * 8) cblocks = b->sm_count * 8;
    k_count_flags<<<cblocks, 256, 0, st>>>(flags, n, cnt);
    g_launches += 2;
    CU_TRY(cudaGetLastError());
    unsigned long long h = 0;
    CU_TRY(cudaMemcpyAsync(&h, cnt, sizeof h, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    *intersections = (int64_t)h;
    return SAHBVH_OK;
}

/* benchmark.c:225-254 on the device: rays in, number of rays whose closest hit exists out (8 bytes back) */
extern "C" int gpu_benchmark_with_bvh(gpu_bvh *b, const Ray *rays, int64_t n, int64_t *intersections)
{
    if (!b || n < 0 || !intersections || (n > 0 && !rays)) {
        sahbvh_set_error("gpu_benchmark_with_bvh: bad arguments");
        return SAHBVH_E_ARG;
    }
    *intersections = 0;
    if (n == 0) return SAHBVH_OK;
    DeviceGuard guard(b->device);
    long long chunk = n < HOST_CHUNK_RAYS ? n : HOST_CHUNK_RAYS;
    int rc = sahbvh_ensure_staging(b, (size_t)chunk * sizeof(Ray), (size_t)chunk * sizeof(gpu_hit));
    if (rc != SAHBVH_OK) return rc;
    unsigned long long *cnt = b->d_counters + 7;
    CU_TRY(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), b->streams[0]));
    CU_TRY(cudaStreamSynchronize(b->streams[0]));
    int k = 0;
    for (long long off = 0; off < n; off += chunk, k ^= 1) {
        long long m = n - off < chunk ? n - off : chunk;
        cudaStream_t st = b->streams[k];
        CU_TRY(cudaMemcpyAsync(b->d_in[k], rays + off, (size_t)m * sizeof(Ray), cudaMemcpyHostToDevice, st));
        /* benchmark.c:242-244 only counts hit_something: the 4-wide kernel writes 4-byte ids, no records */
        const bool ids_only = g_ray_mode.load() == SAHBVH_RAYS_WIDE;
        rc = ids_only ? sahbvh_launch_rays(b, b->d_in[k], m, nullptr, st, false, b->d_out[k])
                      : sahbvh_launch_rays(b, b->d_in[k], m, b->d_out[k], st, false);
        if (rc != SAHBVH_OK) return rc;
        int cblocks = (int)((m + 255) / 256);
        if (cblocks > b->sm_count * 8) cblocks = b->sm_count * 8;
        if (ids_only) k_count_prim_ids<<<cblocks, 256, 0, st>>>((const int32_t *)b->d_out[k], m, cnt);
        else k_count_prims<<<cblocks, 256, 0, st>>>((const gpu_hit *)b->d_out[k], m, cnt);
        g_launches++;
        CU_TRY(cudaGetLastError());
    }
    CU_TRY(cudaStreamSynchronize(b->streams[0]));
    CU_TRY(cudaStreamSynchronize(b->streams[1]));
    unsigned long long h = 0;
    CU_TRY(cudaMemcpy(&h, cnt, sizeof h, cudaMemcpyDeviceToHost));
    *intersections = (int64_t)h;
    return SAHBVH_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* peer framebuffer: the trace kernel of every rank stores its band straight into ONE GPU's frame  */
/* ------------------------------------------------------------------------------------------ */
/* page-locked host memory for callers without the CUDA headers (see sahbvh_gpu.h) */
static int host_mem_ready(const char *who, bool ok_args)
{
    if (!ok_args) {
        sahbvh_set_error("%s: bad arguments", who);
        return SAHBVH_E_ARG;
    }
    if (sahbvh_device_count() == 0) {
        sahbvh_set_error("no usable CUDA device");
        return SAHBVH_E_CUDA;
    }
    return SAHBVH_OK;
}

extern "C" int sahbvh_host_alloc(size_t bytes, void **ptr)
{
    if (ptr) *ptr = nullptr;
    int rc = host_mem_ready("sahbvh_host_alloc", ptr != nullptr && bytes > 0);
    if (rc != SAHBVH_OK) return rc;
    CU_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return SAHBVH_OK;
}

extern "C" int sahbvh_host_free(void *ptr)
{
    if (!ptr) return SAHBVH_OK;
    CU_TRY(cudaFreeHost(ptr));
    return SAHBVH_OK;
}

extern "C" int sahbvh_host_pin(void *ptr, size_t bytes)
{
    int rc = host_mem_ready("sahbvh_host_pin", ptr != nullptr && bytes > 0);
    if (rc != SAHBVH_OK) return rc;
    CU_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return SAHBVH_OK;
}

extern "C" int sahbvh_host_unpin(void *ptr)
{
    if (!ptr) return SAHBVH_OK;
    CU_TRY(cudaHostUnregister(ptr));
    return SAHBVH_OK;
}

static_assert(sizeof(sahbvh_ipc_handle) == sizeof(cudaIpcMemHandle_t), "IPC handle size");

extern "C" int sahbvh_peer_alloc(int device, size_t bytes, void **d_ptr, sahbvh_ipc_handle *handle)
{
    if (!d_ptr || !handle || bytes == 0) {
        sahbvh_set_error("sahbvh_peer_alloc: bad arguments");
        return SAHBVH_E_ARG;
    }
    *d_ptr = nullptr;
    if (sahbvh_device_count() == 0) {
        sahbvh_set_error("no usable CUDA device");
        return SAHBVH_E_CUDA;
    }
    if (device < 0) CU_TRY(cudaGetDevice(&device));
    DeviceGuard guard(device);
    void *p = nullptr;
    CU_TRY(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        sahbvh_set_error("cudaIpcGetMemHandle failed: %s", cu_errstr(e));
        return SAHBVH_E_CUDA;
    }
    memcpy(handle, &h, sizeof h);
    *d_ptr = p;
    return SAHBVH_OK;
}

extern "C" int sahbvh_peer_open(int device, const sahbvh_ipc_handle *handle, void **d_ptr)
{
    if (!d_ptr || !handle) {
        sahbvh_set_error("sahbvh_peer_open: bad arguments");
        return SAHBVH_E_ARG;
    }
    *d_ptr = nullptr;
    if (device < 0) CU_TRY(cudaGetDevice(&device));
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CU_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SAHBVH_OK;
}

/* "frame k of this rank is in place": ordered after the rank's trace kernel on the same stream, the store travels
 * through the same peer mapping as the frame data (system-scope fence first) */
__global__ void k_peer_signal(unsigned long long *flag, unsigned long long value)
{
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(flag) = value;
}

extern "C" int sahbvh_peer_signal(void *d_flag, uint64_t value, void *stream)
{
    if (!d_flag) {
        sahbvh_set_error("sahbvh_peer_signal: NULL flag");
        return SAHBVH_E_ARG;
    }
    k_peer_signal<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long *)d_flag, (unsigned long long)value);
    g_launches++;
    CU_TRY(cudaGetLastError());
    return SAHBVH_OK;
}

/* Device-side wait: the stream does not advance until every one of the n 64-bit flags is >= value (flags written by
 * other GPUs' sahbvh_peer_signal through peer memory, or read from another GPU's memory through a mapping).  This
 * is the frame loop's back-pressure: the owner of a frame buffer consumes frame k only when every band has landed,
 * and a producer reuses a buffer only when the owner has signalled that it was consumed -- without a host round trip.
 * A flag that never arrives would hang the stream, so the wait gives up after timeout_ms and stores 1 into *status
 * (a 32-bit word, may be NULL) instead. */
__global__ void k_peer_wait(const unsigned long long *flags, int n, unsigned long long value, unsigned long long timeout_ns,
                            unsigned int *status)
{
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < n; i++) {
        const volatile unsigned long long *f = flags + i;
        while (*f < value) {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                if (status) *status = 1u;
                return;
            }
        }
    }
    __threadfence_system();
}

extern "C" int sahbvh_peer_wait(const void *d_flags, int n, uint64_t value, uint32_t timeout_ms, void *d_status, void *stream)
{
    if (!d_flags || n <= 0) {
        sahbvh_set_error("sahbvh_peer_wait: bad arguments");
        return SAHBVH_E_ARG;
    }
    k_peer_wait<<<1, 1, 0, (cudaStream_t)stream>>>((const unsigned long long *)d_flags, n, (unsigned long long)value,
                                                   (unsigned long long)(timeout_ms ? timeout_ms : 2000u) * 1000000ull,
                                                   (unsigned int *)d_status);
    g_launches++;
    CU_TRY(cudaGetLastError());
    return SAHBVH_OK;
}

/* benchmark.c:242-244 on an id buffer already on the device: *d_count (uint64) += number of ids >= 0 */
extern "C" int gpu_count_prims_dev(const void *d_prims, int64_t n, void *d_count, void *stream)
{
    if (!d_prims || !d_count || n < 0) {
        sahbvh_set_error("gpu_count_prims_dev: bad arguments");
        return SAHBVH_E_ARG;
    }
    if (n == 0) return SAHBVH_OK;
    long long blocks = (n / 16 + 255) / 256;   /* 16 ids per thread and pass */
    if (blocks > 148 * 4) blocks = 148 * 4;     /* a light grid: the consumer shares the root GPU with the next frame's trace kernel */
    if (blocks < 1) blocks = 1;
    k_count_prim_ids<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const int32_t *)d_prims, n, (unsigned long long *)d_count);
    g_launches++;
    CU_TRY(cudaGetLastError());
    return SAHBVH_OK;
}

/* copies n 64-bit flags to the host (blocking on `stream`): the owner checks that every rank has signalled frame k */
extern "C" int sahbvh_peer_read_flags(const void *d_flags, int n, uint64_t *out, void *stream)
{
    if (!d_flags || !out || n <= 0) {
        sahbvh_set_error("sahbvh_peer_read_flags: bad arguments");
        return SAHBVH_E_ARG;
    }
    CU_TRY(cudaMemcpyAsync(out, d_flags, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return SAHBVH_OK;
}

extern "C" int sahbvh_peer_close(int device, void *d_ptr)
{
    if (!d_ptr) return SAHBVH_OK;
    if (device < 0) CU_TRY(cudaGetDevice(&device));
    DeviceGuard guard(device);
    CU_TRY(cudaIpcCloseMemHandle(d_ptr));
    return SAHBVH_OK;
}

extern "C" int sahbvh_peer_free(int device, void *d_ptr)
{
    if (!d_ptr) return SAHBVH_OK;
    if (device < 0) CU_TRY(cudaGetDevice(&device));
    DeviceGuard guard(device);
    CU_TRY(cudaFree(d_ptr));
    return SAHBVH_OK;
}

static int read_counters(gpu_bvh *b, uint64_t rays, gpu_trace_stats *stats)
{
    unsigned long long h[8];
    CU_TRY(cudaMemcpy(h, b->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    stats->rays = rays;
    stats->inner_visits = h[0];
    stats->sphere_tests = h[1];
    stats->exact_box_tests = h[2];
    stats->hits = h[3];
    stats->max_stack = h[4];
    stats->max_visits = h[5];
    return SAHBVH_OK;
}

extern "C" int gpu_trace_closest_hit_stats(gpu_bvh *b, const void *d_rays, int64_t n, void *d_out, gpu_trace_stats *stats)
{
    if (!b || !stats || n < 0 || (n > 0 && (!d_rays || !d_out))) {
        sahbvh_set_error("gpu_trace_closest_hit_stats: bad arguments");
        return SAHBVH_E_ARG;
    }
    DeviceGuard guard(b->device);
    CU_TRY(cudaMemset(b->d_counters, 0, 8 * sizeof(unsigned long long)));
    int rc = sahbvh_launch_rays(b, d_rays, n, d_out, 0, true);
    if (rc != SAHBVH_OK) return rc;
    CU_TRY(cudaDeviceSynchronize());
    return read_counters(b, (uint64_t)n, stats);
}

extern "C" int gpu_trace_primary_stats(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1,
                                       void *d_out, gpu_trace_stats *stats)
{
    int rc = check_frame(b, cam, width, height, y0, y1, d_out);
    if (rc != SAHBVH_OK) return rc;
    if (!stats) return SAHBVH_E_ARG;
    DeviceGuard guard(b->device);
    CU_TRY(cudaMemset(b->d_counters, 0, 8 * sizeof(unsigned long long)));
    rc = sahbvh_launch_primary(b, cam, width, height, y0, y1, d_out, 0, true, nullptr, 0);
    if (rc != SAHBVH_OK) return rc;
    CU_TRY(cudaDeviceSynchronize());
    return read_counters(b, (uint64_t)(y1 - y0) * (uint64_t)width, stats);
}

/* ------------------------------------------------------------------------------------------ */
/* roofline denominators                                                                        */
/* ------------------------------------------------------------------------------------------ */

template <bool MIX>
__global__ void __launch_bounds__(256) k_peak_fp32(float *sink, int iters, float seed)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 0.001f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
            if (MIX) {
                a4 = fminf(a4, a0); a5 = fmaxf(a5, a1); a6 = fminf(a6, a2); a7 = fmaxf(a7, a3);
            } else {
                a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
            }
        }
    }
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 12345.678f) sink[0] = r;
}

__global__ void __launch_bounds__(256) k_peak_read(const float4 *__restrict__ buf, size_t n4, int repeats, float *sink)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < repeats; r++) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v = __ldcg(buf + i); /* L2-level load: bypass L1 so the 48 MiB sweep measures L2, not L1 */
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 12345.678f) sink[0] = acc;
}

extern "C" int sahbvh_measure_peaks(int device, sahbvh_peaks *out)
{
    if (!out) return SAHBVH_E_ARG;
    memset(out, 0, sizeof *out);
    if (sahbvh_device_count() == 0) {
        sahbvh_set_error("no usable CUDA device");
        return SAHBVH_E_CUDA;
    }
    if (device < 0) CU_TRY(cudaGetDevice(&device));
    DeviceGuard guard(device);
    int sms = 0;
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    out->sm_count = sms;
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    float *sink = nullptr;
    CU_TRY(cudaMalloc(&sink, 256));
    const int blocks = sms * 8, iters = 4096;
    float ms = 0.f;
    for (int mix = 0; mix < 2; mix++) {
        double best = 0;
        for (int rep = 0; rep < 4; rep++) {
            CU_TRY(cudaEventRecord(e0));
            if (mix) k_peak_fp32<true><<<blocks, 256>>>(sink, iters, 1.0f);
            else k_peak_fp32<false><<<blocks, 256>>>(sink, iters, 1.0f);
            g_launches++;
            CU_TRY(cudaEventRecord(e1));
            CU_TRY(cudaEventSynchronize(e1));
            CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
            double rate = (double)blocks * 256.0 * iters * 64.0 / (ms * 1e-3);
            if (rep > 0 && rate > best) best = rate;
        }
        if (mix) out->fp32_mix_inst_per_s = best;
        else out->ffma_inst_per_s = best;
    }
    for (int big = 0; big < 2; big++) {
        size_t bytes = big ? ((size_t)2 << 30) : ((size_t)48 << 20);
        int repeats = big ? 1 : 32;
        float4 *buf = nullptr;
        if (cudaMalloc(&buf, bytes) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        cudaMemset(buf, 0, bytes);
        double best = 0;
        for (int rep = 0; rep < 4; rep++) {
            CU_TRY(cudaEventRecord(e0));
            k_peak_read<<<sms * 16, 256>>>(buf, bytes / 16, repeats, sink);
            g_launches++;
            CU_TRY(cudaEventRecord(e1));
            CU_TRY(cudaEventSynchronize(e1));
            CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
            double gbs = (double)bytes * repeats / (ms * 1e-3) / 1e9;
            if (rep > 0 && gbs > best) best = gbs;
        }
        if (big) out->hbm_read_gbs = best;
        else out->l2_read_gbs = best;
        cudaFree(buf);
    }
    cudaFree(sink);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return SAHBVH_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* scalar drop-in backend: ray_bvh_intersect(Ray, BVHNode*) (hit.c:91)                          */
/* ------------------------------------------------------------------------------------------ */

struct ScalarEntry {
    const BVHNode *root;
    uint64_t fingerprint;
    const Sphere *base;
    gpu_bvh *bvh;
};

/* The cache is keyed by the root's ADDRESS, and only this library's own free_bvh drops an entry.  A driver that
 * frees the tree with its own free_bvh (benchmark.c:81-88 wins at link time when src/benchmark.c stays in the
 * build) and rebuilds can get the same address back from malloc for a different tree.  Every call therefore
 * re-checks a fingerprint of the top of the tree -- the first 64 nodes in breadth-first order: bounds, child and
 * sphere pointers, sphere_count (FNV-1a) -- and re-uploads on a mismatch. */
static uint64_t scalar_fingerprint(const BVHNode *root)
{
    const BVHNode *queue[64];
    int head = 0, tail = 0;
    queue[tail++] = root;
    uint64_t h = 1469598103934665603ull;
    while (head < tail) {
        const BVHNode *p = queue[head++];
        unsigned char rec[sizeof(AABB) + 3 * sizeof(void *) + sizeof(int)];
        memcpy(rec, &p->bounds, sizeof(AABB));
        memcpy(rec + sizeof(AABB), &p->left, sizeof(void *));
        memcpy(rec + sizeof(AABB) + sizeof(void *), &p->right, sizeof(void *));
        memcpy(rec + sizeof(AABB) + 2 * sizeof(void *), &p->sphere, sizeof(void *));
        memcpy(rec + sizeof(AABB) + 3 * sizeof(void *), &p->sphere_count, sizeof(int));
        for (size_t i = 0; i < sizeof rec; i++) h = (h ^ rec[i]) * 1099511628211ull;
        if (!p->sphere && p->left && p->right) {
            if (tail < 64) queue[tail++] = p->left;
            if (tail < 64) queue[tail++] = p->right;
        }
    }
    return h;
}
static std::mutex g_scalar_mu;
static std::vector<ScalarEntry> g_scalar;

/* The scalar signature does not carry the sphere array; recover [base, base+n) from the leaves. */
static int scalar_span(const BVHNode *root, const Sphere **base, int *n)
{
    const Sphere *lo = nullptr, *hi = nullptr, *lo_any = nullptr;
    std::vector<const BVHNode *> stk;
    stk.push_back(root);
    while (!stk.empty()) {
        const BVHNode *p = stk.back();
        stk.pop_back();
        if (p->sphere) {
            if (!lo_any || p->sphere < lo_any) lo_any = p->sphere;
            if (p->sphere_count > 0) {
                if (!lo || p->sphere < lo) lo = p->sphere;
                const Sphere *last = p->sphere + (p->sphere_count - 1);
                if (!hi || last > hi) hi = last;
            }
            continue;
        }
        if (!p->left || !p->right) {
            sahbvh_set_error("ray_bvh_intersect: inner node with a NULL child");
            return SAHBVH_E_ARG;
        }
        stk.push_back(p->left);
        stk.push_back(p->right);
    }
    if (!lo) { /* no sphere at all: every leaf is empty */
        *base = lo_any;
        *n = 0;
        return SAHBVH_OK;
    }
    if (lo_any < lo) lo = lo_any;
    *base = lo;
    *n = (int)(hi - lo) + 1;
    return SAHBVH_OK;
}

extern "C" int sahbvh_scalar_trace(const BVHNode *root, const Ray *ray, HitRecord *out)
{
    if (!root || !ray || !out) {
        sahbvh_set_error("ray_bvh_intersect: NULL argument");
        return SAHBVH_E_ARG;
    }
    std::lock_guard<std::mutex> lock(g_scalar_mu);
    const uint64_t fp = scalar_fingerprint(root);
    ScalarEntry *e = nullptr;
    for (size_t i = 0; i < g_scalar.size(); i++) {
        if (g_scalar[i].root != root) continue;
        if (g_scalar[i].fingerprint == fp) {
            e = &g_scalar[i];
        } else { /* same address, different tree: the old one was freed behind our back */
            gpu_bvh_free(g_scalar[i].bvh);
            g_scalar.erase(g_scalar.begin() + (long)i);
        }
        break;
    }
    if (!e) {
        const Sphere *base = nullptr;
        int n = 0;
        int rc = scalar_span(root, &base, &n);
        if (rc != SAHBVH_OK) return rc;
        /* leaves that point past the recovered span (empty leaves at the array end, Q3) are
         * clamped to the one-past-the-end sentinel by the flattener's index check below */
        sah_tree t;
        rc = sah_tree_init(&t, 2 * (int64_t)(n > 0 ? n : 1) + 16);
        if (rc != SAHBVH_OK) return rc;
        rc = sah_tree_from_pointers(&t, root, base, 0x7ffffff0);
        if (rc == SAHBVH_OK) {
            for (int64_t i = 0; i < t.used; i++)
                if (t.nodes[i].sphere > n) t.nodes[i].sphere = n;
        }
        flat_bvh f;
        memset(&f, 0, sizeof f);
        gpu_bvh *bvh = nullptr;
        if (rc == SAHBVH_OK) rc = flat_bvh_from_tree(&f, &t, base, n);
        if (rc == SAHBVH_OK) rc = gpu_bvh_from_flat(&f, -1, &bvh);
        flat_bvh_release(&f);
        sah_tree_free(&t);
        if (rc != SAHBVH_OK) return rc;
        g_scalar.push_back({root, fp, base, bvh});
        e = &g_scalar.back();
    }
    gpu_hit h;
    int rc = gpu_trace_closest_hit(e->bvh, ray, 1, &h);
    if (rc != SAHBVH_OK) return rc;
    memset(out, 0, sizeof *out);
    if (h.prim >= 0) {
        out->t = h.t;
        out->point.x = h.px; out->point.y = h.py; out->point.z = h.pz;
        out->normal.x = h.nx; out->normal.y = h.ny; out->normal.z = h.nz;
        out->hit_something = 1;
        out->object = const_cast<Sphere *>(e->base) + h.prim;
    }
    return SAHBVH_OK;
}

extern "C" void sahbvh_scalar_forget(const BVHNode *root)
{
    std::lock_guard<std::mutex> lock(g_scalar_mu);
    for (size_t i = 0; i < g_scalar.size(); i++) {
        if (g_scalar[i].root == root) {
            gpu_bvh_free(g_scalar[i].bvh);
            g_scalar.erase(g_scalar.begin() + (long)i);
            return;
        }
    }
}
