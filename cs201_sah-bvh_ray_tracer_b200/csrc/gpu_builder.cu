/*
 * gpu_builder.cu -- SAH builder on the device with the SAME tree, bounds and in-place sphere
 * permutation as the reference's build_bvh_node (bvh.c:117-209).  SURVEY 8(f) row 1 / north_star (4).
 *
 * The host builder (sah_builder.c) already restates the 21 sweeps per node as ONE binned sweep
 * (8 bins per axis; counts are integers and box unions are fmin/fmax, so the 21 costs are
 * bit-identical).  Here the same restatement runs level by level over all "large" nodes at once:
 *
 *   per level   k_prepare   planes lo + (k/8)(hi-lo) of every active node (bvh.c:146-159), bins reset
 *               k_bin       one thread per sphere: 3 bin indices, count + box union per bin
 *                           (shared-memory bins for the block's first node, global atomics otherwise;
 *                           floats as order-preserving uint32 so atomicMin/Max give fmin/fmax)
 *               k_split     one thread per node: the 21 costs 0.125f + (NL*SA_L + NR*SA_R) in the
 *                           reference's order, strict '<' keeps the first best (bvh.c:161-169);
 *                           child bounds = unions of the chosen bins; children classified
 *               k_flags + scan + k_lpos + k_scatter   the in-place partition of bvh.c:172-201,
 *                           reproduced EXACTLY (see below), out of place between two buffers
 *   afterwards  k_small     one thread per subtree of <= GB_SMALL spheres (and per node the level
 *                           pass cannot handle: non-monotone/non-finite planes, all 21 costs NaN):
 *                           the host builder's loop verbatim, including the real swap sequence
 *
 * Exact partition order.  bvh.c:172-201 scans i = start..end and, for a left element, swaps a[i] with
 * a[mid++].  Left elements therefore keep their order (the k-th one lands on start+k-1), and a right
 * element sitting at position q stays there until mid reaches q, when it jumps to the ORIGINAL position
 * of the left element that is being placed on q.  With Lpos[start+k-1] = original position of the k-th
 * left element the final position of a right element is  q <- Lpos[q] while q < mid.  The jump map is
 * injective, so all chains together touch every position at most once.  The final order matters only
 * where leaves hold several spheres -- it defines which sphere is spheres[start] -- but it
 * is reproduced everywhere, so the permuted array is byte-identical to the reference's.
 *
 * Node ids come out in allocation order; k_gb_sizes (bottom-up subtree sizes: the second child to
 * arrive at a parent continues upwards) and k_gb_emit (preorder id = sum over the path to the root
 * of 1, plus the left sibling's subtree size where the path goes through a right child) renumber
 * them to the preorder the host builder produces, and k_gb_permute applies the sphere permutation,
 * so the host only copies the finished arrays.
 * Built with -fmad=false like the rest of the library.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <vector>

#include "sahbvh_internal.h"

#define GB_SMALL 64          /* subtrees of at most this many spheres are finished by one thread */
#define GB_LEAF_DEPTH 40     /* bvh.c:131 */
#define GB_THREADS 256

struct GbWork {
    int32_t node, start, end, depth;
    int32_t mode;            /* 0 = level-parallel, 1 = handed to the serial pass */
    int32_t axis, mid;
    float split;
    int32_t child_work[2];   /* next-level work index of the left/right child, -1 = not a large node */
    int32_t regular;
    float plane[3][7];
};
struct GbBins {
    uint32_t cnt[3][8];
    uint32_t lo[3][8][3];
    uint32_t hi[3][8][3];
};
struct GbSmall { int32_t node, start, end, depth; };
struct GbCounters { int32_t nodes, next_work, small, overflow, max_depth, pad[3]; };

/* error text for sahbvh_last_error(); clears the runtime's non-sticky last-error slot (see gpu_handle.cuh) */
static inline const char *cu_errstr(cudaError_t e)
{
    (void)cudaGetLastError();
    return cudaGetErrorString(e);
}

#define GB_TRY(expr)                                                                               \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            sahbvh_set_error("gpu builder: %s failed: %s", #expr, cu_errstr(_e));        \
            rc = SAHBVH_E_CUDA;                                                                    \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)

/* order-preserving float <-> uint32 (NaNs are never encoded: callers skip them, like fmin/fmax) */
__device__ __forceinline__ uint32_t f_enc(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f_dec(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
#define ENC_POS_INF 0xff800000u /* f_enc(+inf) */
#define ENC_NEG_INF 0x007fffffu /* f_enc(-inf) */

struct GbBox { float lo[3], hi[3]; };
__device__ __forceinline__ void gb_reset(GbBox &b)
{
    b.lo[0] = b.lo[1] = b.lo[2] = INFINITY;
    b.hi[0] = b.hi[1] = b.hi[2] = -INFINITY;
}
__device__ __forceinline__ void gb_merge(GbBox &b, const GbBox &o)
{
#pragma unroll
    for (int a = 0; a < 3; a++) {
        b.lo[a] = fminf(b.lo[a], o.lo[a]);
        b.hi[a] = fmaxf(b.hi[a], o.hi[a]);
    }
}
__device__ __forceinline__ void gb_sphere_box(const float4 s, GbBox &b)
{ /* bvh.c:26-35 */
    b.lo[0] = s.x - s.w; b.lo[1] = s.y - s.w; b.lo[2] = s.z - s.w;
    b.hi[0] = s.x + s.w; b.hi[1] = s.y + s.w; b.hi[2] = s.z + s.w;
}
__device__ __forceinline__ float gb_area(const GbBox &b)
{ /* bvh.c:48-57 */
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}
__device__ __forceinline__ float gb_axis(const float4 s, int a) { return a == 0 ? s.x : (a == 1 ? s.y : s.z); }

/* bvh.c:146-159: returns 1 when the 7 planes of every axis are finite and non-decreasing */
__device__ int gb_planes(const GbBox &nb, float plane[3][7])
{
    int regular = 1;
    for (int a = 0; a < 3; a++) {
        float ext = nb.hi[a] - nb.lo[a];
        for (int k = 1; k < 8; k++) {
            float p = nb.lo[a] + ((float)k / 8.0f) * ext;
            plane[a][k - 1] = p;
            if (!isfinite(p) || (k > 1 && p < plane[a][k - 2])) regular = 0;
        }
    }
    return regular;
}
__device__ __forceinline__ int gb_bin(const float p[7], float c)
{
    return !(c < p[0]) + !(c < p[1]) + !(c < p[2]) + !(c < p[3]) + !(c < p[4]) + !(c < p[5]) + !(c < p[6]);
}

/* the 21 costs of one node from its bins, in the reference's order; returns lcount of the winner */
struct GbChoice { float cost, split; int axis, k; int lcount; GbBox left, right; };
__device__ void gb_choose(const int cnt[3][8], const GbBox box[3][8], const float plane[3][7], GbChoice &c)
{
    c.cost = INFINITY; c.split = 0.0f; c.axis = 0; c.k = 0; c.lcount = 0; /* bvh.c:139-141 */
    gb_reset(c.left); gb_reset(c.right);
    for (int a = 0; a < 3; a++) {
        GbBox right[9];
        int rcount[9];
        gb_reset(right[8]);
        rcount[8] = 0;
        for (int b = 7; b >= 0; b--) {
            right[b] = right[b + 1];
            gb_merge(right[b], box[a][b]);
            rcount[b] = rcount[b + 1] + cnt[a][b];
        }
        GbBox left;
        gb_reset(left);
        int lcount = 0;
        for (int k = 1; k < 8; k++) {
            gb_merge(left, box[a][k - 1]);
            lcount += cnt[a][k - 1];
            float cost = 0.125f + ((float)lcount * gb_area(left) + (float)rcount[k] * gb_area(right[k]));
            if (cost < c.cost) {
                c.cost = cost; c.axis = a; c.k = k; c.split = plane[a][k - 1]; c.lcount = lcount;
                c.left = left; c.right = right[k];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* level-parallel pass                                                                          */
/* ------------------------------------------------------------------------------------------ */
__global__ void k_gb_init(const Sphere *__restrict__ in, int m, float4 *sph, int32_t *ids, int32_t *node_of, uint32_t *root_enc)
{
    int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i < 6 && blockIdx.x == 0) root_enc[i] = i < 3 ? ENC_POS_INF : ENC_NEG_INF;
    if (i >= m) return;
    const Sphere s = in[i];
    sph[i] = make_float4(s.center.x, s.center.y, s.center.z, s.radius);
    ids[i] = i;
    node_of[i] = 0;
}

__global__ void k_gb_root_bounds(const float4 *__restrict__ sph, int m, uint32_t *root_enc)
{
    __shared__ uint32_t s_enc[6];
    if (threadIdx.x < 6) s_enc[threadIdx.x] = threadIdx.x < 3 ? ENC_POS_INF : ENC_NEG_INF;
    __syncthreads();
    int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i < m) {
        GbBox b;
        gb_sphere_box(sph[i], b);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (b.lo[a] == b.lo[a]) atomicMin(&s_enc[a], f_enc(b.lo[a]));
            if (b.hi[a] == b.hi[a]) atomicMax(&s_enc[3 + a], f_enc(b.hi[a]));
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&root_enc[threadIdx.x], s_enc[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&root_enc[threadIdx.x], s_enc[threadIdx.x]);
}

__global__ void k_gb_root_finish(const uint32_t *root_enc, int m, int depth, sah_node *nodes, GbWork *work, GbCounters *ctr,
                                 int32_t *parent)
{
    parent[0] = -1;
    sah_node &n = nodes[0];
    for (int a = 0; a < 3; a++) {
        n.lo[a] = f_dec(root_enc[a]);
        n.hi[a] = f_dec(root_enc[3 + a]);
    }
    n.left = n.right = -1;
    n.sphere = -1;
    n.count = 0;
    work[0].node = 0; work[0].start = 0; work[0].end = m; work[0].depth = depth; work[0].mode = 0;
    ctr->nodes = 1; ctr->next_work = 0; ctr->small = 0; ctr->overflow = 0; ctr->max_depth = 0;
}

__global__ void k_gb_prepare(GbWork *work, int n_work, const sah_node *__restrict__ nodes, GbBins *bins)
{
    int w = blockIdx.x * 64 + threadIdx.x;
    if (w >= n_work) return;
    GbWork &k = work[w];
    GbBox nb;
    for (int a = 0; a < 3; a++) { nb.lo[a] = nodes[k.node].lo[a]; nb.hi[a] = nodes[k.node].hi[a]; }
    float plane[3][7];
    k.regular = gb_planes(nb, plane);
    for (int a = 0; a < 3; a++)
        for (int j = 0; j < 7; j++) k.plane[a][j] = plane[a][j];
    k.mode = k.regular ? 0 : 1;
    k.child_work[0] = k.child_work[1] = -1;
    GbBins &b = bins[w];
    for (int a = 0; a < 3; a++)
        for (int j = 0; j < 8; j++) {
            b.cnt[a][j] = 0;
            for (int x = 0; x < 3; x++) { b.lo[a][j][x] = ENC_POS_INF; b.hi[a][j][x] = ENC_NEG_INF; }
        }
}

__device__ __forceinline__ void gb_bin_update(uint32_t *cnt, uint32_t *lo, uint32_t *hi, const GbBox &sb)
{
    atomicAdd(cnt, 1u);
#pragma unroll
    for (int x = 0; x < 3; x++) {
        if (sb.lo[x] == sb.lo[x]) atomicMin(&lo[x], f_enc(sb.lo[x]));
        if (sb.hi[x] == sb.hi[x]) atomicMax(&hi[x], f_enc(sb.hi[x]));
    }
}

__global__ void __launch_bounds__(GB_THREADS)
k_gb_bin(const float4 *__restrict__ sph, const int32_t *__restrict__ node_of, int m, const GbWork *__restrict__ work, GbBins *bins)
{
    __shared__ GbBins s_bins;
    __shared__ int s_w0;
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (threadIdx.x == 0) s_w0 = node_of[blockIdx.x * GB_THREADS];
    uint32_t *sb32 = reinterpret_cast<uint32_t *>(&s_bins);
    for (int j = threadIdx.x; j < (int)(sizeof(GbBins) / 4); j += GB_THREADS)
        sb32[j] = j < 24 ? 0u : (j < 24 + 72 ? ENC_POS_INF : ENC_NEG_INF);
    __syncthreads();
    const int w0 = s_w0;
    const bool w0_live = w0 >= 0 && work[w0].mode == 0;
    if (i < m) {
        const int w = node_of[i];
        if (w >= 0 && work[w].mode == 0) {
            const float4 s = sph[i];
            GbBox sb;
            gb_sphere_box(s, sb);
            GbBins &dst = (w == w0) ? s_bins : bins[w];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const int b = gb_bin(work[w].plane[a], gb_axis(s, a));
                gb_bin_update(&dst.cnt[a][b], dst.lo[a][b], dst.hi[a][b], sb);
            }
        }
    }
    __syncthreads();
    if (w0_live) {
        uint32_t *g32 = reinterpret_cast<uint32_t *>(&bins[w0]);
        for (int j = threadIdx.x; j < (int)(sizeof(GbBins) / 4); j += GB_THREADS) {
            const uint32_t v = sb32[j];
            if (j < 24) { if (v) atomicAdd(&g32[j], v); }
            else if (j < 24 + 72) { if (v != ENC_POS_INF) atomicMin(&g32[j], v); }
            else { if (v != ENC_NEG_INF) atomicMax(&g32[j], v); }
        }
    }
}

__device__ void gb_write_leaf(sah_node &n, const GbBox &b, int sphere, int count)
{
    for (int a = 0; a < 3; a++) { n.lo[a] = b.lo[a]; n.hi[a] = b.hi[a]; }
    n.left = n.right = -1;
    n.sphere = sphere;
    n.count = count;
}

__global__ void k_gb_split(GbWork *work, int n_work, const GbBins *__restrict__ bins, sah_node *nodes, int node_cap,
                           GbWork *next, GbSmall *small, GbCounters *ctr, int base_index, int32_t *parent)
{
    int w = blockIdx.x * 64 + threadIdx.x;
    if (w >= n_work) return;
    GbWork &k = work[w];
    if (k.mode == 0) {
        int cnt[3][8];
        GbBox box[3][8];
        float plane[3][7];
        const GbBins &b = bins[w];
        for (int a = 0; a < 3; a++) {
            for (int j = 0; j < 8; j++) {
                cnt[a][j] = (int)b.cnt[a][j];
                for (int x = 0; x < 3; x++) { box[a][j].lo[x] = f_dec(b.lo[a][j][x]); box[a][j].hi[x] = f_dec(b.hi[a][j][x]); }
            }
            for (int j = 0; j < 7; j++) plane[a][j] = k.plane[a][j];
        }
        GbChoice c;
        gb_choose(cnt, box, plane, c);
        if (!(c.cost < INFINITY)) {
            k.mode = 1; /* all 21 costs NaN/inf: the x < 0.0f fallback of bvh.c:139-141 needs a real sweep */
        } else {
            const int base = atomicAdd(&ctr->nodes, 2);
            if (base + 2 > node_cap) { ctr->overflow = 1; k.mode = 2; return; }
            k.axis = c.axis; k.split = c.split; k.mid = k.start + c.lcount;
            sah_node &me = nodes[k.node];
            me.left = base; me.right = base + 1; me.sphere = -1; me.count = 0;
            parent[base] = parent[base + 1] = k.node;
            for (int side = 0; side < 2; side++) {
                const int cs = side ? k.mid : k.start, ce = side ? k.end : k.mid, cn = ce - cs;
                const GbBox &cb = side ? c.right : c.left;
                sah_node &ch = nodes[base + side];
                if (cn <= 1 || k.depth + 1 >= GB_LEAF_DEPTH) {
                    gb_write_leaf(ch, cb, base_index + cs, cn);
                } else {
                    for (int a = 0; a < 3; a++) { ch.lo[a] = cb.lo[a]; ch.hi[a] = cb.hi[a]; }
                    ch.left = ch.right = -1; ch.sphere = -1; ch.count = 0;
                    if (cn <= GB_SMALL) {
                        const int si = atomicAdd(&ctr->small, 1);
                        small[si].node = base + side; small[si].start = cs; small[si].end = ce; small[si].depth = k.depth + 1;
                    } else {
                        const int ni = atomicAdd(&ctr->next_work, 1);
                        next[ni].node = base + side; next[ni].start = cs; next[ni].end = ce; next[ni].depth = k.depth + 1;
                        next[ni].mode = 0;
                        k.child_work[side] = ni;
                    }
                }
            }
        }
    }
    if (k.mode == 1) { /* serial pass takes the whole subtree */
        const int si = atomicAdd(&ctr->small, 1);
        small[si].node = k.node; small[si].start = k.start; small[si].end = k.end; small[si].depth = k.depth;
    }
}

/* flags + per-block exclusive scan (1024 positions per block) */
#define GB_SCAN_ITEMS 4
__global__ void __launch_bounds__(GB_THREADS)
k_gb_flags_scan(const float4 *__restrict__ sph, const int32_t *__restrict__ node_of, int m, const GbWork *__restrict__ work,
                uint8_t *flags, uint32_t *scan, uint32_t *block_sums)
{
    __shared__ uint32_t s_warp[GB_THREADS / 32];
    const int base = (blockIdx.x * GB_THREADS + threadIdx.x) * GB_SCAN_ITEMS;
    uint32_t f[GB_SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int j = 0; j < GB_SCAN_ITEMS; j++) {
        const int i = base + j;
        uint32_t v = 0;
        if (i < m) {
            const int w = node_of[i];
            if (w >= 0 && work[w].mode == 0) v = gb_axis(sph[i], work[w].axis) < work[w].split ? 1u : 0u;
            flags[i] = (uint8_t)v;
        }
        f[j] = v;
        sum += v;
    }
    uint32_t incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = lane < GB_THREADS / 32 ? s_warp[lane] : 0, inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane < GB_THREADS / 32) s_warp[lane] = inc - v;
        if (lane == GB_THREADS / 32 - 1) block_sums[blockIdx.x] = inc;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + incl - sum;
#pragma unroll
    for (int j = 0; j < GB_SCAN_ITEMS; j++) {
        const int i = base + j;
        if (i < m) scan[i] = run;
        run += f[j];
    }
}

__global__ void __launch_bounds__(1024) k_gb_scan_sums(uint32_t *block_sums, int n_blocks)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        uint32_t v = i < n_blocks ? block_sums[i] : 0, incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t x = s_warp[lane], inc = x;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            s_warp[lane] = inc - x;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (i < n_blocks) block_sums[i] = carry + s_warp[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
}

/* left elements: final position, and Lpos[final] = original position */
__global__ void __launch_bounds__(GB_THREADS)
k_gb_lpos(const int32_t *__restrict__ node_of, int m, const GbWork *__restrict__ work, const uint8_t *__restrict__ flags,
          const uint32_t *__restrict__ scan, const uint32_t *__restrict__ block_sums, int32_t *Lpos)
{
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i >= m || !flags[i]) return;
    const GbWork &k = work[node_of[i]];
    const uint32_t mine = scan[i] + block_sums[i / (GB_THREADS * GB_SCAN_ITEMS)];
    const uint32_t first = scan[k.start] + block_sums[k.start / (GB_THREADS * GB_SCAN_ITEMS)];
    Lpos[k.start + (int)(mine - first)] = i;
}

__global__ void __launch_bounds__(GB_THREADS)
k_gb_scatter(const float4 *__restrict__ sph, const int32_t *__restrict__ ids, const int32_t *__restrict__ node_of, int m,
             const GbWork *__restrict__ work, const uint8_t *__restrict__ flags, const uint32_t *__restrict__ scan,
             const uint32_t *__restrict__ block_sums, const int32_t *__restrict__ Lpos,
             float4 *sph_out, int32_t *ids_out, int32_t *node_of_out)
{
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i >= m) return;
    const int w = node_of[i];
    int dest = i, next_w = -1;
    if (w >= 0 && work[w].mode == 0) {
        const GbWork &k = work[w];
        if (flags[i]) {
            const uint32_t mine = scan[i] + block_sums[i / (GB_THREADS * GB_SCAN_ITEMS)];
            const uint32_t first = scan[k.start] + block_sums[k.start / (GB_THREADS * GB_SCAN_ITEMS)];
            dest = k.start + (int)(mine - first);
        } else {
            int q = i;
            while (q < k.mid) q = Lpos[q];
            dest = q;
        }
        next_w = k.child_work[dest < k.mid ? 0 : 1];
    }
    sph_out[dest] = sph[i];
    ids_out[dest] = ids[i];
    node_of_out[dest] = next_w;
}

/* ------------------------------------------------------------------------------------------ */
/* serial pass: one thread per small subtree (sah_builder.c's build_range, iterative)            */
/* ------------------------------------------------------------------------------------------ */
__device__ float gb_eval_sweep(const float4 *sph, int start, int end, int axis, float split)
{ /* evaluate_sah, bvh.c:59-97 */
    int lc = 0, rc = 0;
    GbBox lb, rb;
    gb_reset(lb); gb_reset(rb);
    for (int i = start; i < end; i++) {
        GbBox sb;
        gb_sphere_box(sph[i], sb);
        if (gb_axis(sph[i], axis) < split) { lc++; gb_merge(lb, sb); }
        else { rc++; gb_merge(rb, sb); }
    }
    return 0.125f + ((float)lc * gb_area(lb) + (float)rc * gb_area(rb));
}

__global__ void __launch_bounds__(32)
k_gb_small(const GbSmall *__restrict__ small, int n_small, float4 *sph, int32_t *ids, sah_node *nodes, int node_cap,
           GbCounters *ctr, int base_index, int32_t *parent)
{
    const int t = blockIdx.x * 32 + threadIdx.x;
    if (t >= n_small) return;
    GbSmall stack[GB_LEAF_DEPTH + 4];
    int sp = 0;
    stack[sp++] = small[t];
    while (sp > 0) {
        const GbSmall job = stack[--sp];
        GbBox nb;
        gb_reset(nb);
        for (int i = job.start; i < job.end; i++) {
            GbBox sb;
            gb_sphere_box(sph[i], sb);
            gb_merge(nb, sb);
        }
        const int n = job.end - job.start;
        sah_node &me = nodes[job.node];
        if (n <= 1 || job.depth >= GB_LEAF_DEPTH) {
            gb_write_leaf(me, nb, base_index + job.start, n);
            continue;
        }
        for (int a = 0; a < 3; a++) { me.lo[a] = nb.lo[a]; me.hi[a] = nb.hi[a]; }
        float plane[3][7];
        const int regular = gb_planes(nb, plane);
        float best_cost = INFINITY, best_split = 0.0f;
        int best_axis = 0;
        if (regular) {
            int cnt[3][8];
            GbBox box[3][8];
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 8; b++) { cnt[a][b] = 0; gb_reset(box[a][b]); }
            for (int i = job.start; i < job.end; i++) {
                const float4 s = sph[i];
                GbBox sb;
                gb_sphere_box(s, sb);
                for (int a = 0; a < 3; a++) {
                    const int b = gb_bin(plane[a], gb_axis(s, a));
                    cnt[a][b]++;
                    gb_merge(box[a][b], sb);
                }
            }
            GbChoice c;
            gb_choose(cnt, box, plane, c);
            best_cost = c.cost; best_split = c.split; best_axis = c.axis;
        } else {
            for (int a = 0; a < 3; a++)
                for (int k = 1; k < 8; k++) {
                    const float cost = gb_eval_sweep(sph, job.start, job.end, a, plane[a][k - 1]);
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = plane[a][k - 1]; }
                }
        }
        (void)best_cost;
        int mid = job.start; /* bvh.c:172-201, the reference's own swap sequence */
        for (int i = job.start; i < job.end; i++) {
            if (gb_axis(sph[i], best_axis) < best_split) {
                if (i != mid) {
                    const float4 ts = sph[i]; sph[i] = sph[mid]; sph[mid] = ts;
                    const int32_t ti = ids[i]; ids[i] = ids[mid]; ids[mid] = ti;
                }
                mid++;
            }
        }
        const int base = atomicAdd(&ctr->nodes, 2);
        if (base + 2 > node_cap) { ctr->overflow = 1; return; }
        me.left = base; me.right = base + 1; me.sphere = -1; me.count = 0;
        parent[base] = parent[base + 1] = job.node;
        GbSmall r = {base + 1, mid, job.end, job.depth + 1}, l = {base, job.start, mid, job.depth + 1};
        stack[sp++] = r;
        stack[sp++] = l;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* renumbering to preorder + sphere permutation                                                  */
/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(GB_THREADS)
k_gb_sizes(const sah_node *__restrict__ nodes, int n_nodes, const int32_t *__restrict__ parent, int32_t *visit, int32_t *size)
{
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i >= n_nodes || nodes[i].sphere < 0) return; /* leaves start the walk */
    size[i] = 1;
    int c = i;
    for (;;) {
        const int p = parent[c];
        if (p < 0) break;
        __threadfence();
        if (atomicAdd(&visit[p], 1) == 0) break; /* the sibling subtree is not finished yet */
        __threadfence();
        size[p] = 1 + ((volatile int32_t *)size)[nodes[p].left] + ((volatile int32_t *)size)[nodes[p].right];
        c = p;
    }
}

__global__ void __launch_bounds__(GB_THREADS)
k_gb_emit(const sah_node *__restrict__ nodes, int n_nodes, const int32_t *__restrict__ parent, const int32_t *__restrict__ size,
          sah_node *out, GbCounters *ctr)
{
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i >= n_nodes) return;
    int pre = 0, c = i, levels = 0;
    for (;;) {
        const int p = parent[c];
        if (p < 0) break;
        pre += 1;
        levels++;
        if (nodes[p].right == c) pre += size[nodes[p].left];
        c = p;
    }
    if (levels > ctr->max_depth) atomicMax(&ctr->max_depth, levels);
    sah_node nd = nodes[i];
    if (nd.sphere < 0) {
        const int ls = size[nd.left];
        nd.left = pre + 1;
        nd.right = pre + 1 + ls;
    }
    out[pre] = nd;
}

__global__ void __launch_bounds__(GB_THREADS)
k_gb_permute(const Sphere *__restrict__ in, const int32_t *__restrict__ ids, int m, Sphere *out)
{
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i >= m) return;
    out[i] = in[ids[i]];
}

/* ------------------------------------------------------------------------------------------ */
/* flatten (flatten.c) and 4-wide collapse (wide.c) on the device: same bytes as the host code    */
/* ------------------------------------------------------------------------------------------ */
struct FlStats { int32_t n_leaf, multi, empty, bad_index, max_depth; uint32_t rmax_bits; uint32_t lo[3], hi[3]; int32_t n_special, pad; };

/* generic exclusive scan of uint32 (1024 items per block), three launches */
__global__ void __launch_bounds__(GB_THREADS) k_scan32_block(const uint32_t *__restrict__ in, int n, uint32_t *out, uint32_t *block_sums)
{
    __shared__ uint32_t s_warp[GB_THREADS / 32];
    const int base = (blockIdx.x * GB_THREADS + threadIdx.x) * GB_SCAN_ITEMS;
    uint32_t f[GB_SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int j = 0; j < GB_SCAN_ITEMS; j++) {
        f[j] = base + j < n ? in[base + j] : 0u;
        sum += f[j];
    }
    uint32_t incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = lane < GB_THREADS / 32 ? s_warp[lane] : 0, inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane < GB_THREADS / 32) s_warp[lane] = inc - v;
        if (lane == GB_THREADS / 32 - 1) block_sums[blockIdx.x] = inc;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + incl - sum;
#pragma unroll
    for (int j = 0; j < GB_SCAN_ITEMS; j++) {
        if (base + j < n) out[base + j] = run;
        run += f[j];
    }
}
__global__ void __launch_bounds__(GB_THREADS) k_scan32_add(uint32_t *out, int n, const uint32_t *__restrict__ block_sums, uint32_t *total)
{
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i < n) out[i] += block_sums[i / (GB_THREADS * GB_SCAN_ITEMS)];
    if (i == 0 && total) {
        const int last = (n - 1) / (GB_THREADS * GB_SCAN_ITEMS);
        (void)last;
    }
}
/* out = exclusive scan of in; *total (device) = sum of all; in may alias nothing */
static void scan32(const uint32_t *d_in, int n, uint32_t *d_out, uint32_t *d_sums)
{
    const int sb = (n + GB_THREADS * GB_SCAN_ITEMS - 1) / (GB_THREADS * GB_SCAN_ITEMS);
    k_scan32_block<<<sb, GB_THREADS>>>(d_in, n, d_out, d_sums);
    k_gb_scan_sums<<<1, 1024>>>(d_sums, sb);
    k_scan32_add<<<(n + GB_THREADS - 1) / GB_THREADS, GB_THREADS>>>(d_out, n, d_sums, nullptr);
}

__global__ void __launch_bounds__(GB_THREADS)
k_fl_flags(const sah_node *__restrict__ pre, int n_nodes, int n_total, uint32_t *flags, FlStats *st)
{
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i >= n_nodes) return;
    const sah_node nd = pre[i];
    const bool inner = nd.sphere < 0;
    flags[i] = inner ? 1u : 0u;
    if (!inner) {
        atomicAdd(&st->n_leaf, 1);
        if (nd.count > 1) atomicAdd(&st->multi, 1);
        if (nd.count == 0) atomicAdd(&st->empty, 1);
        if (nd.sphere > n_total) st->bad_index = 1;
    }
}

/* flatten.c: inner node -> two 32-byte child records; an empty leaf carries its parent's bounds (Q3) */
__global__ void __launch_bounds__(GB_THREADS)
k_fl_nodes(const sah_node *__restrict__ pre, int n_nodes, const uint32_t *__restrict__ inner_id, float *flat)
{
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i >= n_nodes) return;
    const sah_node nd = pre[i];
    if (nd.sphere >= 0) return;
    float *w = flat + (size_t)inner_id[i] * 16;
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int ci = side ? nd.right : nd.left;
        const sah_node ch = pre[ci];
        const bool empty_leaf = ch.sphere >= 0 && ch.count == 0;
        const float *lo = empty_leaf ? nd.lo : ch.lo, *hi = empty_leaf ? nd.hi : ch.hi;
        const int32_t ref = ch.sphere >= 0 ? ~ch.sphere : (int32_t)inner_id[ci];
        float *q = w + 8 * side;
        q[0] = lo[0]; q[1] = lo[1]; q[2] = lo[2]; q[3] = hi[0];
        q[4] = hi[1]; q[5] = hi[2]; q[6] = __int_as_float(ref); q[7] = 0.0f;
    }
}

/* spheres4 / colours of the WHOLE array (indices are global), max |r| and the bounds of all spheres */
__global__ void __launch_bounds__(GB_THREADS)
k_fl_spheres(const Sphere *__restrict__ all, int n_total, float *spheres4, uint32_t *colors, FlStats *st)
{
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i > n_total) return;
    if (i == n_total) { /* the reference's one-past-the-end read: never a hit */
        const float nan = __int_as_float(0x7fc00000);
        reinterpret_cast<float4 *>(spheres4)[i] = make_float4(nan, nan, nan, nan);
        colors[i] = 0;
        return;
    }
    const Sphere s = all[i];
    reinterpret_cast<float4 *>(spheres4)[i] = make_float4(s.center.x, s.center.y, s.center.z, s.radius);
    uint32_t c;
    memcpy(&c, &s.color, 4);
    colors[i] = c;
    if (s.radius < 0.0f) st->bad_index = 1;                          /* the host flattener reports it (flatten.c) */
    const float ar = fabsf(s.radius);
    if (ar == ar) atomicMax(&st->rmax_bits, __float_as_uint(ar));   /* ar >= 0: bit order == value order */
    const float cc[3] = {s.center.x, s.center.y, s.center.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float lo = cc[a] - ar, hi = cc[a] + ar;
        if (lo == lo) atomicMin(&st->lo[a], f_enc(lo));
        if (hi == hi) atomicMax(&st->hi[a], f_enc(hi));
    }
}

/* ---- wide collapse, one level per launch pair ---- */
struct WdRec { float lo[3], hi[3]; int32_t ref; };

__device__ __forceinline__ void wd_load(const float *__restrict__ flat, int32_t node, int child, WdRec &r)
{
    const float4 *q = reinterpret_cast<const float4 *>(flat + (size_t)node * 16 + child * 8);
    const float4 a = q[0], b = q[1];
    r.lo[0] = a.x; r.lo[1] = a.y; r.lo[2] = a.z; r.hi[0] = a.w; r.hi[1] = b.x; r.hi[2] = b.y;
    r.ref = __float_as_int(b.z);
}
__device__ __forceinline__ float wd_area(const WdRec &r)
{
    const float dx = r.hi[0] - r.lo[0], dy = r.hi[1] - r.lo[1], dz = r.hi[2] - r.lo[2];
    const float a = dx * dy + dy * dz + dz * dx;
    return a == a ? a : INFINITY;
}
/* the greedy collapse of wide.c: up to four children of binary node b2, left-to-right */
__device__ int wd_collapse(const float *__restrict__ flat, int32_t b2, WdRec ch[4])
{
    int n = 2;
    wd_load(flat, b2, 0, ch[0]);
    wd_load(flat, b2, 1, ch[1]);
    while (n < 4) {
        int best = -1;
        float best_area = -1.0f;
        for (int k = 0; k < n; k++)
            if (ch[k].ref >= 0) {
                const float a = wd_area(ch[k]);
                if (a > best_area) { best_area = a; best = k; }
            }
        if (best < 0) break;
        const int32_t open = ch[best].ref;
        for (int k = n; k > best + 1; k--) ch[k] = ch[k - 1];
        wd_load(flat, open, 0, ch[best]);
        wd_load(flat, open, 1, ch[best + 1]);
        n++;
    }
    return n;
}
/* leaf classification of wide.c: 0 = never hits (dropped), 1 = regular, 2 = special */
__device__ int wd_leaf_class(const WdRec &r, const float *__restrict__ spheres4, int n_spheres)
{
    const int32_t idx = ~r.ref;
    if (r.ref == SAHBVH_REF_NONE || (uint32_t)idx >= (uint32_t)n_spheres) return 0;
    const float4 s = reinterpret_cast<const float4 *>(spheres4)[idx];
    const float lo[3] = {s.x - s.w, s.y - s.w, s.z - s.w}, hi[3] = {s.x + s.w, s.y + s.w, s.z + s.w};
    bool same = true;
    for (int a = 0; a < 3; a++)
        same = same && __float_as_uint(lo[a]) == __float_as_uint(r.lo[a]) && __float_as_uint(hi[a]) == __float_as_uint(r.hi[a]);
    return same ? 1 : 2;
}

__global__ void __launch_bounds__(128)
k_wd_count(const float *__restrict__ flat, const float *__restrict__ spheres4, int n_spheres,
           const int32_t *__restrict__ frontier, int n_front, uint32_t *cnt_inner, uint32_t *cnt_special)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n_front) return;
    WdRec ch[4];
    const int n = wd_collapse(flat, frontier[i], ch);
    uint32_t ni = 0, ns = 0;
    for (int k = 0; k < n; k++) {
        if (ch[k].ref >= 0) ni++;
        else if (wd_leaf_class(ch[k], spheres4, n_spheres) == 2) ns++;
    }
    cnt_inner[i] = ni;
    cnt_special[i] = ns;
}

/* quantise_axis of wide.c, in double (the host's long-double containment check is re-evaluated in double:
 * the margins are 2^20 times wider than double rounding) */
__device__ int wd_quantise_axis(const WdRec *ch, int n, int a, float &p_out, int &e_out, uint8_t qlo[4], uint8_t qhi[4])
{
    double mn = INFINITY, mx = -INFINITY, maxabs = 0.0;
    for (int k = 0; k < n; k++) {
        const double lo = ch[k].lo[a], hi = ch[k].hi[a];
        if (!isfinite(lo) || !isfinite(hi) || lo > hi) return -1;
        if (lo < mn) mn = lo;
        if (hi > mx) mx = hi;
        if (fabs(lo) > maxabs) maxabs = fabs(lo);
        if (fabs(hi) > maxabs) maxabs = fabs(hi);
    }
    if (n == 0) return -1;
    const double ext = mx - mn;
    int e = -60;
    if (ext > 0.0) {
        int ee;
        frexp(ext / 253.0, &ee);
        if (ee > e) e = ee;
    }
    if (maxabs > 0.0) {
        int ea;
        frexp(maxabs, &ea);
        if (ea - 24 + 3 > e) e = ea - 24 + 3;
    }
    if (e > 27) return -1;
    const double s = ldexp(1.0, e);
    float p = __double2float_rn(mn - 1024.5 * s);
    while ((double)p > mn - 1024.25 * s) p = nextafterf(p, -INFINITY);
    if (!isfinite(p)) return -1;
    for (int k = 0; k < n; k++) {
        const double l = floor(((double)ch[k].lo[a] - (double)p) / s - 1024.0 - 1.0 / 256.0);
        const double h = ceil(((double)ch[k].hi[a] - (double)p) / s - 1024.0 + 1.0 / 256.0);
        if (l < 0.0 || h > 255.0) return -1;
        const double dl = (double)p + (1024.0 + l) * s, dh = (double)p + (1024.0 + h) * s, m = s / 512.0;
        if (!(dl <= (double)ch[k].lo[a] - m) || !(dh >= (double)ch[k].hi[a] + m)) return -1;
        qlo[k] = (uint8_t)l;
        qhi[k] = (uint8_t)h;
    }
    p_out = p;
    e_out = e;
    return 0;
}

__global__ void __launch_bounds__(128)
k_wd_emit(const float *__restrict__ flat, const float *__restrict__ spheres4, int n_spheres,
          const int32_t *__restrict__ frontier, int n_front, int first_id, const uint32_t *__restrict__ off_inner,
          const uint32_t *__restrict__ off_special, int next_first_id, int special_base,
          uint32_t *wide, float *full, float *special, int32_t *next_frontier)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n_front) return;
    const int32_t me = first_id + i;
    WdRec ch[4];
    int n = wd_collapse(flat, frontier[i], ch);
    float fl[32];
    for (int k = 0; k < 4; k++) {
        float *q = fl + 8 * k;
        if (k < n) {
            q[0] = ch[k].lo[0]; q[1] = ch[k].lo[1]; q[2] = ch[k].lo[2]; q[3] = ch[k].hi[0];
            q[4] = ch[k].hi[1]; q[5] = ch[k].hi[2]; q[6] = __int_as_float(ch[k].ref); q[7] = 0.0f;
        } else {
            q[0] = q[1] = q[2] = INFINITY; q[3] = q[4] = q[5] = -INFINITY;
            q[6] = __int_as_float(SAHBVH_REF_NONE); q[7] = 0.0f;
        }
    }
    int ni = 0, ns = 0, m = 0;
    for (int k = 0; k < n; k++) {
        if (ch[k].ref >= 0) {
            const int32_t id = next_first_id + (int32_t)off_inner[i] + ni;
            next_frontier[off_inner[i] + ni] = ch[k].ref;
            ni++;
            ch[k].ref = id;
            fl[8 * k + 6] = __int_as_float(id);
        } else {
            const int cls = wd_leaf_class(ch[k], spheres4, n_spheres);
            if (cls == 0) continue;
            if (cls == 2) {
                const int32_t sk = special_base + (int32_t)off_special[i] + ns;
                ns++;
                float *q = special + (size_t)sk * 8;
                q[0] = ch[k].lo[0]; q[1] = ch[k].lo[1]; q[2] = ch[k].lo[2]; q[3] = ch[k].hi[0];
                q[4] = ch[k].hi[1]; q[5] = ch[k].hi[2]; q[6] = __int_as_float(~ch[k].ref); q[7] = 0.0f;
                ch[k].ref = ~(int32_t)(SAHBVH_WIDE_SPECIAL | (uint32_t)sk);
            }
        }
        ch[m++] = ch[k];
    }
    n = m;
    uint32_t node[16];
    for (int k = 0; k < 16; k++) node[k] = 0;
    const int s_word[3] = {3, 14, 15};
    for (int a = 0; a < 3; a++) {
        float p = __int_as_float(0x7fc00000);
        int e = 0;
        uint8_t qlo[4] = {255, 255, 255, 255}, qhi[4] = {0, 0, 0, 0};
        if (wd_quantise_axis(ch, n, a, p, e, qlo, qhi) != 0) {
            p = __int_as_float(0x7fc00000);
            e = 0;
        }
        node[a] = __float_as_uint(p);
        node[s_word[a]] = __float_as_uint(ldexpf(1.0f, e));
        for (int k = 0; k < 4; k++) {
            node[4 + a] |= (uint32_t)(k < n ? qlo[k] : 255) << (8 * k);
            node[7 + a] |= (uint32_t)(k < n ? qhi[k] : 0) << (8 * k);
        }
    }
    for (int k = 0; k < 4; k++) node[10 + k] = k < n ? (uint32_t)ch[k].ref : (uint32_t)SAHBVH_REF_NONE;
    uint4 *wo = reinterpret_cast<uint4 *>(wide + (size_t)me * 16);
    for (int k = 0; k < 4; k++) wo[k] = make_uint4(node[4 * k], node[4 * k + 1], node[4 * k + 2], node[4 * k + 3]);
    float4 *fo = reinterpret_cast<float4 *>(full + (size_t)me * 32);
    for (int k = 0; k < 8; k++) fo[k] = make_float4(fl[4 * k], fl[4 * k + 1], fl[4 * k + 2], fl[4 * k + 3]);
}

/* ------------------------------------------------------------------------------------------ */
/* host driver                                                                                  */
/* ------------------------------------------------------------------------------------------ */
static int flatten_and_collapse(const sah_node *d_pre, int n_nodes, int tree_depth, const Sphere *d_all, int n_total,
                                gpu_bvh_parts *parts, bool timing);

/* t != NULL: tree to the host (preorder sah_tree).  parts != NULL: tree stays on the device and is flattened and
 * collapsed there; only the permuted spheres go back.  Returns 1 when the host path has to take over. */
static int gb_build(sah_tree *t, Sphere *spheres, int n_total, int start, int end, int depth, int device, gpu_bvh_parts *parts)
{
    const int m = end - start;
    if (m <= GB_SMALL * 4 || depth >= GB_LEAF_DEPTH) return 1; /* nothing to parallelise */
    int rc = SAHBVH_OK;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    if (device >= 0 && device != prev_dev) cudaSetDevice(device);

    const int node_cap = 2 * m + (1 << 16) + m / 4;
    const int max_work = m / GB_SMALL + 2, max_small = m / 2 + 2;
    const int blocks = (m + GB_THREADS - 1) / GB_THREADS;
    const int scan_blocks = (m + GB_THREADS * GB_SCAN_ITEMS - 1) / (GB_THREADS * GB_SCAN_ITEMS);
    char *pool = nullptr;
    Sphere *d_in = nullptr, *d_out_sph = nullptr;
    float4 *d_sph[2] = {nullptr, nullptr};
    int32_t *d_ids[2] = {nullptr, nullptr}, *d_node_of[2] = {nullptr, nullptr}, *d_lpos = nullptr;
    int32_t *d_parent = nullptr, *d_visit = nullptr, *d_size = nullptr;
    uint8_t *d_flags = nullptr;
    uint32_t *d_scan = nullptr, *d_sums = nullptr, *d_root = nullptr;
    sah_node *d_nodes = nullptr, *d_pre = nullptr;
    GbWork *d_work[2] = {nullptr, nullptr};
    GbBins *d_bins = nullptr;
    GbSmall *d_small = nullptr;
    GbCounters *d_ctr = nullptr;
    GbCounters h_ctr;
    size_t pool_bytes = 0;
    /* one allocation, carved into 256-byte aligned pieces (pass 0 sizes it, pass 1 hands the pieces out) */
    auto carve = [&](void **ptr, size_t bytes, int pass) {
        if (pass) *ptr = pool + pool_bytes;
        pool_bytes += (bytes + 255) / 256 * 256;
    };
    int cur = 0, n_work = 1, wcur = 0, levels = 0;
    const bool timing = getenv("SAHBVH_BUILD_TIMING") != nullptr; /* stage times on stderr (development aid) */
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        cudaDeviceSynchronize();
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[gpu builder] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };

    for (int pass = 0; pass < 2; pass++) {
        if (pass) GB_TRY(cudaMalloc((void **)&pool, pool_bytes));
        pool_bytes = 0;
        carve((void **)&d_in, (size_t)m * sizeof(Sphere), pass);
        carve((void **)&d_out_sph, (size_t)m * sizeof(Sphere), pass);
        for (int k = 0; k < 2; k++) {
            carve((void **)&d_sph[k], (size_t)m * sizeof(float4), pass);
            carve((void **)&d_ids[k], (size_t)m * sizeof(int32_t), pass);
            carve((void **)&d_node_of[k], (size_t)m * sizeof(int32_t), pass);
            carve((void **)&d_work[k], (size_t)max_work * sizeof(GbWork), pass);
        }
        carve((void **)&d_lpos, (size_t)m * sizeof(int32_t), pass);
        carve((void **)&d_flags, (size_t)m, pass);
        carve((void **)&d_scan, (size_t)m * sizeof(uint32_t), pass);
        carve((void **)&d_sums, (size_t)scan_blocks * sizeof(uint32_t), pass);
        carve((void **)&d_root, 8 * sizeof(uint32_t), pass);
        carve((void **)&d_nodes, (size_t)node_cap * sizeof(sah_node), pass);
        carve((void **)&d_pre, (size_t)node_cap * sizeof(sah_node), pass);
        carve((void **)&d_parent, (size_t)node_cap * sizeof(int32_t), pass);
        carve((void **)&d_visit, (size_t)node_cap * sizeof(int32_t), pass);
        carve((void **)&d_size, (size_t)node_cap * sizeof(int32_t), pass);
        carve((void **)&d_bins, (size_t)max_work * sizeof(GbBins), pass);
        carve((void **)&d_small, (size_t)max_small * sizeof(GbSmall), pass);
        carve((void **)&d_ctr, sizeof(GbCounters), pass);
    }
    lap("cudaMalloc");
    GB_TRY(cudaMemcpy(d_in, spheres + start, (size_t)m * sizeof(Sphere), cudaMemcpyHostToDevice));
    lap("H2D spheres");

    k_gb_init<<<blocks, GB_THREADS>>>(d_in, m, d_sph[0], d_ids[0], d_node_of[0], d_root);
    k_gb_root_bounds<<<blocks, GB_THREADS>>>(d_sph[0], m, d_root);
    k_gb_root_finish<<<1, 1>>>(d_root, m, depth, d_nodes, d_work[0], d_ctr, d_parent);
    GB_TRY(cudaGetLastError());

    while (n_work > 0) {
        const int wb = (n_work + 63) / 64;
        k_gb_prepare<<<wb, 64>>>(d_work[wcur], n_work, d_nodes, d_bins);
        k_gb_bin<<<blocks, GB_THREADS>>>(d_sph[cur], d_node_of[cur], m, d_work[wcur], d_bins);
        k_gb_split<<<wb, 64>>>(d_work[wcur], n_work, d_bins, d_nodes, node_cap, d_work[wcur ^ 1], d_small, d_ctr, start, d_parent);
        k_gb_flags_scan<<<scan_blocks, GB_THREADS>>>(d_sph[cur], d_node_of[cur], m, d_work[wcur], d_flags, d_scan, d_sums);
        k_gb_scan_sums<<<1, 1024>>>(d_sums, scan_blocks);
        k_gb_lpos<<<blocks, GB_THREADS>>>(d_node_of[cur], m, d_work[wcur], d_flags, d_scan, d_sums, d_lpos);
        k_gb_scatter<<<blocks, GB_THREADS>>>(d_sph[cur], d_ids[cur], d_node_of[cur], m, d_work[wcur], d_flags, d_scan, d_sums,
                                             d_lpos, d_sph[cur ^ 1], d_ids[cur ^ 1], d_node_of[cur ^ 1]);
        GB_TRY(cudaGetLastError());
        GB_TRY(cudaMemcpy(&h_ctr, d_ctr, sizeof h_ctr, cudaMemcpyDeviceToHost));
        if (h_ctr.overflow) break;
        n_work = h_ctr.next_work;
        if (n_work > max_work) { h_ctr.overflow = 1; break; }
        GB_TRY(cudaMemset(&d_ctr->next_work, 0, sizeof(int32_t)));
        cur ^= 1;
        wcur ^= 1;
        levels++;
    }
    lap("level-parallel pass");
    if (!h_ctr.overflow && h_ctr.small > 0) {
        if (h_ctr.small > max_small) h_ctr.overflow = 1;
        else {
            k_gb_small<<<(h_ctr.small + 31) / 32, 32>>>(d_small, h_ctr.small, d_sph[cur], d_ids[cur], d_nodes, node_cap, d_ctr, start, d_parent);
            GB_TRY(cudaGetLastError());
            GB_TRY(cudaMemcpy(&h_ctr, d_ctr, sizeof h_ctr, cudaMemcpyDeviceToHost));
        }
    }
    lap("serial pass (small subtrees)");
    if (timing) fprintf(stderr, "[gpu builder] m=%d levels=%d small=%d nodes=%d\n", m, levels, h_ctr.small, h_ctr.nodes);
    if (h_ctr.overflow || h_ctr.nodes > node_cap) {
        rc = 1; /* degenerate input outgrew the node pool: the host builder takes over below */
        goto done;
    }
    {
        const int nb = (h_ctr.nodes + GB_THREADS - 1) / GB_THREADS;
        GB_TRY(cudaMemset(d_visit, 0, (size_t)h_ctr.nodes * sizeof(int32_t)));
        k_gb_sizes<<<nb, GB_THREADS>>>(d_nodes, h_ctr.nodes, d_parent, d_visit, d_size);
        k_gb_emit<<<nb, GB_THREADS>>>(d_nodes, h_ctr.nodes, d_parent, d_size, d_pre, d_ctr);
        k_gb_permute<<<blocks, GB_THREADS>>>(d_in, d_ids[cur], m, d_out_sph);
        GB_TRY(cudaGetLastError());
    }
    lap("preorder renumbering + sphere permutation");
    if (parts) {
        GB_TRY(cudaMemcpy(&h_ctr, d_ctr, sizeof h_ctr, cudaMemcpyDeviceToHost));
        Sphere *d_all = d_out_sph;
        Sphere *d_all_owned = nullptr;
        if (start != 0 || end != n_total) { /* indices are global: the kernels need the whole array */
            GB_TRY(cudaMalloc((void **)&d_all_owned, (size_t)n_total * sizeof(Sphere)));
            cudaError_t e = cudaMemcpy(d_all_owned, spheres, (size_t)n_total * sizeof(Sphere), cudaMemcpyHostToDevice);
            if (e == cudaSuccess)
                e = cudaMemcpy(d_all_owned + start, d_out_sph, (size_t)m * sizeof(Sphere), cudaMemcpyDeviceToDevice);
            if (e != cudaSuccess) { cudaFree(d_all_owned); sahbvh_set_error("gpu builder: %s", cu_errstr(e)); rc = SAHBVH_E_CUDA; goto done; }
            d_all = d_all_owned;
        }
        rc = flatten_and_collapse(d_pre, h_ctr.nodes, h_ctr.max_depth, d_all, n_total, parts, timing);
        cudaFree(d_all_owned);
        lap("flatten + wide collapse (device)");
        /* the caller's array is permuted only once nothing can fail any more: on an error, or when the host
         * pipeline takes over (rc == 1), it still holds the input order the host builder expects */
        if (rc == SAHBVH_OK) {
            cudaError_t e = cudaMemcpy(spheres + start, d_out_sph, (size_t)m * sizeof(Sphere), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) {
                sahbvh_set_error("gpu builder: %s", cu_errstr(e));
                gpu_bvh_parts_release(parts);
                rc = SAHBVH_E_CUDA;
            }
            lap("D2H permuted spheres");
        }
        goto done;
    }
    if (t->cap < h_ctr.nodes) {
        sah_node *nn = (sah_node *)realloc(t->nodes, (size_t)h_ctr.nodes * sizeof(sah_node));
        if (!nn) { rc = SAHBVH_E_NOMEM; goto done; }
        t->nodes = nn;
        t->cap = h_ctr.nodes;
    }
    GB_TRY(cudaMemcpy(t->nodes, d_pre, (size_t)h_ctr.nodes * sizeof(sah_node), cudaMemcpyDeviceToHost));
    t->used = h_ctr.nodes;
    GB_TRY(cudaMemcpy(spheres + start, d_out_sph, (size_t)m * sizeof(Sphere), cudaMemcpyDeviceToHost));
    lap("D2H tree + permuted spheres");
done:
    cudaFree(pool);
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return rc;
}

extern "C" int sah_build_gpu(sah_tree *t, Sphere *spheres, int start, int end, int depth, int device)
{
    int rc = gb_build(t, spheres, end, start, end, depth, device, nullptr);
    if (rc == 1) return sah_build(t, spheres, start, end, depth);
    return rc;
}

extern "C" void gpu_bvh_parts_release(gpu_bvh_parts *p)
{
    if (!p) return;
    cudaFree(p->d_nodes); cudaFree(p->d_spheres4); cudaFree(p->d_colors);
    cudaFree(p->d_wide); cudaFree(p->d_wide_full); cudaFree(p->d_special);
    memset(p, 0, sizeof *p);
}

/* build + flatten + collapse, everything on the device; 1 = caller must use the host path instead */
extern "C" int gpu_build_parts(Sphere *spheres, int n_total, int start, int end, int depth, int device, gpu_bvh_parts *parts)
{
    memset(parts, 0, sizeof *parts);
    int rc = gb_build(nullptr, spheres, n_total, start, end, depth, device, parts);
    if (rc != SAHBVH_OK) gpu_bvh_parts_release(parts);
    return rc;
}

#define FC_TRY(expr)                                                                               \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            sahbvh_set_error("gpu flatten: %s failed: %s", #expr, cu_errstr(_e));        \
            rc = SAHBVH_E_CUDA;                                                                    \
            goto out;                                                                              \
        }                                                                                          \
    } while (0)

static int flatten_and_collapse(const sah_node *d_pre, int n_nodes, int tree_depth, const Sphere *d_all, int n_total,
                                gpu_bvh_parts *parts, bool timing)
{
    int rc = SAHBVH_OK;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        cudaDeviceSynchronize();
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[gpu flatten] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    };
    uint32_t *d_flags = nullptr, *d_inner_id = nullptr, *d_sums = nullptr;
    uint32_t *d_cnt_i = nullptr, *d_cnt_s = nullptr, *d_off_i = nullptr, *d_off_s = nullptr, *d_wide_tmp = nullptr;
    int32_t *d_front[2] = {nullptr, nullptr};
    float *d_full_tmp = nullptr, *d_special_tmp = nullptr;
    FlStats *d_st = nullptr;
    FlStats h_st;
    uint32_t last[2];
    int n_inner = 0, n_front = 1, first_id = 0, special_base = 0, levels = 0, cur = 0;
    const int nb = (n_nodes + GB_THREADS - 1) / GB_THREADS;
    const int sums_cap = (n_nodes + GB_THREADS * GB_SCAN_ITEMS - 1) / (GB_THREADS * GB_SCAN_ITEMS) + 1;

    memset(&h_st, 0, sizeof h_st);
    for (int a = 0; a < 3; a++) { h_st.lo[a] = ENC_POS_INF; h_st.hi[a] = ENC_NEG_INF; }
    FC_TRY(cudaMalloc((void **)&d_st, sizeof(FlStats)));
    FC_TRY(cudaMemcpy(d_st, &h_st, sizeof h_st, cudaMemcpyHostToDevice));
    FC_TRY(cudaMalloc((void **)&d_flags, (size_t)n_nodes * 4));
    FC_TRY(cudaMalloc((void **)&d_inner_id, (size_t)n_nodes * 4));
    FC_TRY(cudaMalloc((void **)&d_sums, (size_t)sums_cap * 4));
    k_fl_flags<<<nb, GB_THREADS>>>(d_pre, n_nodes, n_total, d_flags, d_st);
    scan32(d_flags, n_nodes, d_inner_id, d_sums);
    FC_TRY(cudaGetLastError());
    FC_TRY(cudaMemcpy(&last[0], d_inner_id + (n_nodes - 1), 4, cudaMemcpyDeviceToHost));
    FC_TRY(cudaMemcpy(&last[1], d_flags + (n_nodes - 1), 4, cudaMemcpyDeviceToHost));
    n_inner = (int)(last[0] + last[1]);
    FC_TRY(cudaMemcpy(&h_st, d_st, sizeof h_st, cudaMemcpyDeviceToHost));
    lap("flags + scan");
    if (h_st.bad_index || tree_depth + 2 > SAHBVH_STACK_MAX || n_inner <= 0 || 3 * (tree_depth + 1) + 4 > SAHBVH_WIDE_STACK_MAX) {
        rc = 1; /* the host flattener reports these cases */
        goto out;
    }
    /* ---- flatten ---- */
    FC_TRY(cudaMalloc((void **)&parts->d_nodes, (size_t)n_inner * 64));
    FC_TRY(cudaMalloc((void **)&parts->d_spheres4, ((size_t)n_total + 1) * 16));
    FC_TRY(cudaMalloc((void **)&parts->d_colors, ((size_t)n_total + 1) * 4));
    k_fl_nodes<<<nb, GB_THREADS>>>(d_pre, n_nodes, d_inner_id, parts->d_nodes);
    k_fl_spheres<<<(n_total + 1 + GB_THREADS - 1) / GB_THREADS, GB_THREADS>>>(d_all, n_total, parts->d_spheres4, parts->d_colors, d_st);
    FC_TRY(cudaGetLastError());
    lap("flatten (alloc + kernels)");
    /* ---- wide collapse, level by level (breadth-first numbering, like wide.c) ---- */
    FC_TRY(cudaMalloc((void **)&d_front[0], (size_t)n_inner * 4));
    FC_TRY(cudaMalloc((void **)&d_front[1], (size_t)n_inner * 4));
    FC_TRY(cudaMalloc((void **)&d_cnt_i, (size_t)n_inner * 4));
    FC_TRY(cudaMalloc((void **)&d_cnt_s, (size_t)n_inner * 4));
    FC_TRY(cudaMalloc((void **)&d_off_i, (size_t)n_inner * 4));
    FC_TRY(cudaMalloc((void **)&d_off_s, (size_t)n_inner * 4));
    FC_TRY(cudaMalloc((void **)&d_wide_tmp, (size_t)n_inner * 64));
    FC_TRY(cudaMalloc((void **)&d_full_tmp, (size_t)n_inner * 128));
    FC_TRY(cudaMalloc((void **)&d_special_tmp, ((size_t)h_st.n_leaf + 1) * 32));
    FC_TRY(cudaMemset(d_front[0], 0, 4));
    lap("collapse: temp alloc");
    while (n_front > 0) {
        const int fb = (n_front + 127) / 128;
        uint32_t tot[4];
        k_wd_count<<<fb, 128>>>(parts->d_nodes, parts->d_spheres4, n_total, d_front[cur], n_front, d_cnt_i, d_cnt_s);
        scan32(d_cnt_i, n_front, d_off_i, d_sums);
        scan32(d_cnt_s, n_front, d_off_s, d_sums);
        FC_TRY(cudaGetLastError());
        FC_TRY(cudaMemcpy(&tot[0], d_off_i + (n_front - 1), 4, cudaMemcpyDeviceToHost));
        FC_TRY(cudaMemcpy(&tot[1], d_cnt_i + (n_front - 1), 4, cudaMemcpyDeviceToHost));
        FC_TRY(cudaMemcpy(&tot[2], d_off_s + (n_front - 1), 4, cudaMemcpyDeviceToHost));
        FC_TRY(cudaMemcpy(&tot[3], d_cnt_s + (n_front - 1), 4, cudaMemcpyDeviceToHost));
        const int next_front = (int)(tot[0] + tot[1]), n_spec = (int)(tot[2] + tot[3]);
        if (first_id + n_front + next_front > n_inner || special_base + n_spec > h_st.n_leaf + 1) { rc = 1; goto out; }
        k_wd_emit<<<fb, 128>>>(parts->d_nodes, parts->d_spheres4, n_total, d_front[cur], n_front, first_id, d_off_i, d_off_s,
                               first_id + n_front, special_base, d_wide_tmp, d_full_tmp, d_special_tmp, d_front[cur ^ 1]);
        FC_TRY(cudaGetLastError());
        first_id += n_front;
        special_base += n_spec;
        n_front = next_front;
        cur ^= 1;
        levels++;
    }
    lap("collapse: levels");
    parts->n_wide = first_id;
    parts->n_special = special_base;
    FC_TRY(cudaMalloc((void **)&parts->d_wide, (size_t)first_id * 64));
    FC_TRY(cudaMalloc((void **)&parts->d_wide_full, (size_t)first_id * 128));
    FC_TRY(cudaMalloc((void **)&parts->d_special, (size_t)(special_base > 0 ? special_base : 1) * 32));
    FC_TRY(cudaMemcpy(parts->d_wide, d_wide_tmp, (size_t)first_id * 64, cudaMemcpyDeviceToDevice));
    FC_TRY(cudaMemcpy(parts->d_wide_full, d_full_tmp, (size_t)first_id * 128, cudaMemcpyDeviceToDevice));
    if (special_base > 0) FC_TRY(cudaMemcpy(parts->d_special, d_special_tmp, (size_t)special_base * 32, cudaMemcpyDeviceToDevice));
    else FC_TRY(cudaMemset(parts->d_special, 0, 32));
    FC_TRY(cudaMemcpy(&h_st, d_st, sizeof h_st, cudaMemcpyDeviceToHost));
    if (h_st.bad_index) { rc = 1; goto out; }
    {
        gpu_bvh_info &info = parts->info;
        memset(&info, 0, sizeof info);
        info.n_spheres = n_total;
        info.n_inner = n_inner;
        info.n_leaf_refs = h_st.n_leaf;
        info.max_depth = tree_depth;
        info.multi_leaves = h_st.multi;
        info.empty_leaves = h_st.empty;
        info.n_wide = first_id;
        uint32_t rb = h_st.rmax_bits;
        memcpy(&info.max_radius, &rb, 4);
        for (int a = 0; a < 3; a++) {
            const uint32_t lo = h_st.lo[a], hi = h_st.hi[a];
            const uint32_t lob = (lo & 0x80000000u) ? (lo & 0x7fffffffu) : ~lo, hib = (hi & 0x80000000u) ? (hi & 0x7fffffffu) : ~hi;
            memcpy(&info.scene_lo[a], &lob, 4);
            memcpy(&info.scene_hi[a], &hib, 4);
        }
        info.device_bytes = (int64_t)n_inner * 64 + ((int64_t)n_total + 1) * 20 + (int64_t)first_id * 192 + (int64_t)(special_base > 0 ? special_base : 1) * 32;
    }
    lap("collapse: right-sized copies");
    (void)levels;
out:
    cudaFree(d_st); cudaFree(d_flags); cudaFree(d_inner_id); cudaFree(d_sums);
    cudaFree(d_front[0]); cudaFree(d_front[1]); cudaFree(d_cnt_i); cudaFree(d_cnt_s); cudaFree(d_off_i); cudaFree(d_off_s);
    cudaFree(d_wide_tmp); cudaFree(d_full_tmp); cudaFree(d_special_tmp);
    return rc;
}
