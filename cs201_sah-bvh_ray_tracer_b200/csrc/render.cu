/*
 * render.cu -- the path loop around the closest-hit kernels: SURVEY 8(f) rows 2 and 3.
 *
 *   trace_ray (renderer.c:21-76) is a recursion of depth <= MAX_DEPTH: closest hit, on a hit draw
 *   random_on_hemisphere(normal) (sphere.c:26-32), recurse from the UN-OFFSET hit point
 *   (renderer.c:54, Q13) and blend  final = (Uint8)(base + 0.5 * reflected)  per channel
 *   (renderer.c:56-58, evaluated in double; sums above 255 wrap through the integer conversion);
 *   a miss returns the sky gradient of the ray's direction.y (renderer.c:66-71), exhausted depth
 *   returns black (renderer.c:23-24).
 *
 * Here the recursion is a wavefront: depth 0 is the fused primary kernel, every further depth is one
 * launch of the ray-buffer kernel over the compacted rays that are still alive, k_shade records the
 * hit sphere's colour per pixel and depth and emits the next rays, and k_resolve folds the colours
 * back from the last depth to the first -- the same arithmetic in the same order as the recursion.
 *
 * Random directions: the reference draws them from the serial rand() stream, pixel after pixel, which
 * no parallel schedule can reproduce.  gpu_render_frame therefore takes them in one of two ways:
 *   samples != NULL  the unit vectors random_in_unit_sphere() returned, [pixel][depth] -- with the
 *                    reference's own samples the frame is bit-identical to the reference's (parity mode);
 *   samples == NULL  a counter-based generator keyed by (seed, pixel, depth, attempt) runs the same
 *                    rejection loop (sphere.c:19-24) with the same float expressions (vec3.c:64-69).
 *
 * gpu_accumulate* is the running mean of main.c:368-405.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "gpu_handle.cuh"

#define RN_THREADS 256

struct RenderBufs {
    gpu_hit *hits;
    float *rays[2];
    int32_t *pix[2];
    uint32_t *base;    /* [pixel][max_depth] RGBA8 of the sphere hit at that depth */
    uint8_t *nbase;    /* hits recorded per pixel */
    uint32_t *term;    /* colour returned by the end of the path: sky or black */
    unsigned int *count;
};

__device__ __forceinline__ uint32_t pack_rgba(int r, int g, int b, int a)
{
    return (uint32_t)(r & 0xff) | ((uint32_t)(g & 0xff) << 8) | ((uint32_t)(b & 0xff) << 16) | ((uint32_t)(a & 0xff) << 24);
}

/* renderer.c:66-71 */
__device__ __forceinline__ uint32_t sky_color(float dir_y)
{
    const float t = F_MUL(0.5f, F_ADD(dir_y, 1.0f));
    const float r = F_ADD(F_MUL(F_SUB(1.0f, t), 255.0f), F_MUL(t, 128.0f));
    const float g = F_ADD(F_MUL(F_SUB(1.0f, t), 255.0f), F_MUL(t, 178.0f));
    return pack_rgba((int)r, (int)g, 255, 255);
}

/* counter-based stand-in for rand(): 31 uniform bits per call, keyed by (seed, pixel, depth, draw) */
__device__ __forceinline__ uint32_t rng31(uint64_t seed, uint64_t pixel, uint32_t depth, uint32_t draw)
{
    uint64_t z = seed + 0x9e3779b97f4a7c15ull * (pixel * 64ull + depth) + 0xbf58476d1ce4e5b9ull * (uint64_t)(draw + 1);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    return (uint32_t)(z >> 33);
}

/* sphere.c:19-24 with vec3.c:64-69's expressions: min + (float)r / RAND_MAX * (max - min), min = -1, max = 1 */
__device__ void random_in_unit_sphere_dev(uint64_t seed, uint64_t pixel, uint32_t depth, float &x, float &y, float &z)
{
    const float rmax = 2147483648.0f; /* (float)RAND_MAX */
    for (uint32_t attempt = 0;; attempt++) {
        const float px = F_ADD(-1.0f, F_MUL(F_DIV((float)rng31(seed, pixel, depth, 3 * attempt + 0), rmax), 2.0f));
        const float py = F_ADD(-1.0f, F_MUL(F_DIV((float)rng31(seed, pixel, depth, 3 * attempt + 1), rmax), 2.0f));
        const float pz = F_ADD(-1.0f, F_MUL(F_DIV((float)rng31(seed, pixel, depth, 3 * attempt + 2), rmax), 2.0f));
        const float q = F_ADD(F_ADD(F_MUL(px, px), F_MUL(py, py)), F_MUL(pz, pz));
        if (q < 1.0f && q != 0.0f) {
            const float len = F_SQRT(q); /* vec3.c:21-24: sqrt in double of the float sum == correctly rounded sqrtf */
            x = F_DIV(px, len); y = F_DIV(py, len); z = F_DIV(pz, len);
            return;
        }
        if (attempt > 1000) { x = 0.0f; y = 1.0f; z = 0.0f; return; } /* unreachable in practice (p = 0.52 per attempt) */
    }
}

template <bool PRIMARY>
__global__ void __launch_bounds__(RN_THREADS)
k_shade(const gpu_hit *__restrict__ hits, long long n_rays, const int32_t *__restrict__ pix_in, const float *__restrict__ rays_in,
        SahCamera cam, int width, int y0, const uint32_t *__restrict__ colors, int k, int max_depth,
        const float *__restrict__ samples, unsigned long long seed, uint32_t *base, uint8_t *nbase, uint32_t *term,
        float *rays_out, int32_t *pix_out, unsigned int *out_count)
{
    const long long i = (long long)blockIdx.x * RN_THREADS + threadIdx.x;
    bool emit = false;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0;
    long long p = 0;
    if (i < n_rays) {
        p = PRIMARY ? i : (long long)pix_in[i];
        const float4 h0 = reinterpret_cast<const float4 *>(hits)[2 * i], h1 = reinterpret_cast<const float4 *>(hits)[2 * i + 1];
        const int prim = __float_as_int(h0.y);
        if (prim < 0) {
            float ray_dy;
            if (PRIMARY) {
                float ax, az;
                sah_primary_dir(cam, (int)(i % width), y0 + (int)(i / width), ax, ray_dy, az);
            } else {
                ray_dy = rays_in[i * 6 + 4];
            }
            term[p] = sky_color(ray_dy);
            if (PRIMARY) nbase[p] = 0;
        } else {
            base[p * max_depth + k] = colors[prim] | 0xff000000u;
            nbase[p] = (uint8_t)(k + 1);
            if (k + 1 < max_depth) {
                float sx, sy, sz;
                if (samples) {
                    const float *s = samples + (p * max_depth + k) * 3;
                    sx = s[0]; sy = s[1]; sz = s[2];
                } else {
                    /* keyed by the pixel's index in the WHOLE frame: a frame rendered in row bands (several devices) equals the
                     * frame rendered in one call */
                    random_in_unit_sphere_dev(seed, (uint64_t)p + (uint64_t)y0 * (uint64_t)width, (uint32_t)k, sx, sy, sz);
                }
                const float nx = h1.y, ny = h1.z, nz = h1.w;
                const float dot = F_ADD(F_ADD(F_MUL(sx, nx), F_MUL(sy, ny)), F_MUL(sz, nz)); /* vec3.c:26-28 */
                if (!(dot > 0.0f)) { sx = F_MUL(sx, -1.0f); sy = F_MUL(sy, -1.0f); sz = F_MUL(sz, -1.0f); } /* sphere.c:28-31 */
                ox = h0.z; oy = h0.w; oz = h1.x;
                dx = sx; dy = sy; dz = sz;
            }
            if (k + 1 < max_depth) emit = true;
            else term[p] = pack_rgba(0, 0, 0, 255); /* renderer.c:23-24: depth exhausted */
        }
    }
    /* warp-aggregated append of the surviving rays */
    const unsigned m = __ballot_sync(0xffffffffu, emit);
    if (m) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned basei = 0;
        if (lane == leader) basei = atomicAdd(out_count, (unsigned)__popc(m));
        basei = __shfl_sync(0xffffffffu, basei, leader);
        if (emit) {
            const unsigned o = basei + (unsigned)__popc(m & ((1u << lane) - 1u));
            float *r = rays_out + (size_t)o * 6;
            r[0] = ox; r[1] = oy; r[2] = oz; r[3] = dx; r[4] = dy; r[5] = dz;
            pix_out[o] = (int32_t)p;
        }
    }
}

/* renderer.c:56-60 folded from the last hit back to the first */
__global__ void __launch_bounds__(RN_THREADS)
k_resolve(const uint32_t *__restrict__ base, const uint8_t *__restrict__ nbase, const uint32_t *__restrict__ term, long long n_pixels,
          int max_depth, uint32_t *rgba)
{
    const long long p = (long long)blockIdx.x * RN_THREADS + threadIdx.x;
    if (p >= n_pixels) return;
    uint32_t c = term[p];
    int r = c & 0xff, g = (c >> 8) & 0xff, b = (c >> 16) & 0xff;
    for (int j = (int)nbase[p] - 1; j >= 0; j--) {
        const uint32_t bc = base[p * max_depth + j];
        r = (int)((double)(int)(bc & 0xff) + 0.5 * (double)r) & 0xff;
        g = (int)((double)(int)((bc >> 8) & 0xff) + 0.5 * (double)g) & 0xff;
        b = (int)((double)(int)((bc >> 16) & 0xff) + 0.5 * (double)b) & 0xff;
    }
    rgba[p] = pack_rgba(r, g, b, 255);
}

__global__ void __launch_bounds__(RN_THREADS) k_fill(uint32_t *out, long long n, uint32_t v)
{
    const long long p = (long long)blockIdx.x * RN_THREADS + threadIdx.x;
    if (p < n) out[p] = v;
}

/* main.c:368-405 */
__global__ void __launch_bounds__(RN_THREADS)
k_accumulate(float *accum, int frames, const uint32_t *__restrict__ rgba, long long n_pixels, uint32_t *avg)
{
    const long long p = (long long)blockIdx.x * RN_THREADS + threadIdx.x;
    if (p >= n_pixels) return;
    const uint32_t c = rgba[p];
    int out[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float v = F_DIV((float)((c >> (8 * ch)) & 0xff), 255.0f);
        if (frames <= 1) {
            accum[p * 3 + ch] = v;                       /* main.c:368-372 */
            out[ch] = (int)((c >> (8 * ch)) & 0xff);     /* the first frame shows the colour itself */
        } else {
            const float a = F_ADD(accum[p * 3 + ch], v); /* main.c:392-394 */
            accum[p * 3 + ch] = a;
            const double m = fmin((double)F_MUL(F_DIV(a, (float)frames), 255.0f), (double)255.0f); /* main.c:396-399 */
            out[ch] = (int)m & 0xff;
        }
    }
    avg[p] = pack_rgba(out[0], out[1], out[2], 255);
}

/* ------------------------------------------------------------------------------------------ */

static int render_bufs(gpu_bvh *b, long long n_pixels, int max_depth, RenderBufs *rb)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_hits = take((size_t)n_pixels * sizeof(gpu_hit));
    const size_t o_r0 = take((size_t)n_pixels * 24), o_r1 = take((size_t)n_pixels * 24);
    const size_t o_p0 = take((size_t)n_pixels * 4), o_p1 = take((size_t)n_pixels * 4);
    const size_t o_base = take((size_t)n_pixels * 4 * (size_t)max_depth);
    const size_t o_nb = take((size_t)n_pixels), o_term = take((size_t)n_pixels * 4), o_cnt = take(256);
    if (off > b->render_cap) {
        if (b->d_render) cudaFree(b->d_render);
        b->d_render = nullptr;
        b->render_cap = 0;
        CU_TRY(cudaMalloc(&b->d_render, off));
        b->render_cap = off;
    }
    char *base = (char *)b->d_render;
    rb->hits = (gpu_hit *)(base + o_hits);
    rb->rays[0] = (float *)(base + o_r0); rb->rays[1] = (float *)(base + o_r1);
    rb->pix[0] = (int32_t *)(base + o_p0); rb->pix[1] = (int32_t *)(base + o_p1);
    rb->base = (uint32_t *)(base + o_base);
    rb->nbase = (uint8_t *)(base + o_nb);
    rb->term = (uint32_t *)(base + o_term);
    rb->count = (unsigned int *)(base + o_cnt);
    return SAHBVH_OK;
}

extern "C" int gpu_render_frame_dev(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, int max_depth,
                                    uint64_t seed, const void *d_samples, void *d_rgba, void *stream)
{
    if (!b || !cam || !d_rgba || width <= 0 || height <= 0 || y0 < 0 || y1 < y0 || y1 > height || max_depth < 0 || max_depth > 64) {
        sahbvh_set_error("gpu_render_frame: bad arguments");
        return SAHBVH_E_ARG;
    }
    const long long n_pixels = (long long)(y1 - y0) * width;
    if (n_pixels == 0) return SAHBVH_OK;
    DeviceGuard guard(b->device);
    if (!guard.ok) {
        sahbvh_set_error("gpu_render_frame: cudaSetDevice(%d) failed", b->device);
        return SAHBVH_E_CUDA;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = SAHBVH_OK;
    do {
        if (max_depth == 0) { /* renderer.c:23-24: black, alpha 255 */
            k_fill<<<(unsigned)((n_pixels + RN_THREADS - 1) / RN_THREADS), RN_THREADS, 0, st>>>((uint32_t *)d_rgba, n_pixels, 0xff000000u);
            g_launches++;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { sahbvh_set_error("gpu_render_frame: %s", cu_errstr(e)); rc = SAHBVH_E_CUDA; }
            break;
        }
        RenderBufs rb;
        rc = render_bufs(b, n_pixels, max_depth, &rb);
        if (rc != SAHBVH_OK) break;
        SahCamera c;
        sahbvh_make_camera(cam, width, height, &c);
        rc = sahbvh_launch_primary(b, cam, width, height, y0, y1, rb.hits, st, false, nullptr, SAH_IDS_NONE);
        if (rc != SAHBVH_OK) break;
        long long n_rays = n_pixels;
        int cur = 0;
        for (int k = 0; k < max_depth && n_rays > 0; k++) {
            cudaError_t e = cudaMemsetAsync(rb.count, 0, sizeof(unsigned int), st);
            const unsigned grid = (unsigned)((n_rays + RN_THREADS - 1) / RN_THREADS);
            if (k == 0)
                k_shade<true><<<grid, RN_THREADS, 0, st>>>(rb.hits, n_rays, nullptr, nullptr, c, width, y0, b->d_colors, k, max_depth,
                                                          (const float *)d_samples, (unsigned long long)seed, rb.base, rb.nbase,
                                                          rb.term, rb.rays[cur ^ 1], rb.pix[cur ^ 1], rb.count);
            else
                k_shade<false><<<grid, RN_THREADS, 0, st>>>(rb.hits, n_rays, rb.pix[cur], rb.rays[cur], c, width, y0, b->d_colors, k,
                                                           max_depth, (const float *)d_samples, (unsigned long long)seed, rb.base,
                                                           rb.nbase, rb.term, rb.rays[cur ^ 1], rb.pix[cur ^ 1], rb.count);
            g_launches++;
            unsigned int h_count = 0;
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(&h_count, rb.count, sizeof h_count, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { sahbvh_set_error("gpu_render_frame: %s", cu_errstr(e)); rc = SAHBVH_E_CUDA; break; }
            cur ^= 1;
            n_rays = (long long)h_count;
            if (n_rays > 0 && k + 1 < max_depth) {
                rc = sahbvh_launch_rays(b, rb.rays[cur], n_rays, rb.hits, st, false);
                if (rc != SAHBVH_OK) break;
            }
        }
        if (rc != SAHBVH_OK) break;
        k_resolve<<<(unsigned)((n_pixels + RN_THREADS - 1) / RN_THREADS), RN_THREADS, 0, st>>>(rb.base, rb.nbase, rb.term, n_pixels,
                                                                                               max_depth, (uint32_t *)d_rgba);
        g_launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { sahbvh_set_error("gpu_render_frame: %s", cu_errstr(e)); rc = SAHBVH_E_CUDA; }
    } while (0);
    return rc;
}

extern "C" int gpu_render_frame(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, int max_depth,
                                uint64_t seed, const Vec3 *samples, uint8_t *rgba_out)
{
    /* everything is validated before the first allocation (the _dev call repeats the same checks) */
    if (!b || !cam || !rgba_out || width <= 0 || height <= 0 || y0 < 0 || y1 < y0 || y1 > height || max_depth < 0 || max_depth > 64) {
        sahbvh_set_error("gpu_render_frame: bad arguments");
        return SAHBVH_E_ARG;
    }
    const long long n_pixels = (long long)(y1 - y0) * width;
    if (n_pixels == 0) return SAHBVH_OK;
    DeviceGuard guard(b->device);
    if (!guard.ok) {
        sahbvh_set_error("gpu_render_frame: cudaSetDevice(%d) failed", b->device);
        return SAHBVH_E_CUDA;
    }
    /* frame and samples go through the handle's staging buffers (grown on demand, kept between calls): a cudaMalloc /
     * cudaFree pair per frame costs more than the frame's D2H copy */
    const size_t sample_bytes = samples && max_depth > 0 ? (size_t)n_pixels * (size_t)max_depth * sizeof(Vec3) : 0;
    int rc = sahbvh_ensure_staging(b, sample_bytes, (size_t)n_pixels * 4);
    if (rc != SAHBVH_OK) return rc;
    void *d_samples = sample_bytes ? b->d_in[0] : nullptr, *d_rgba = b->d_out[0];
    if (sample_bytes) CU_TRY(cudaMemcpyAsync(d_samples, samples, sample_bytes, cudaMemcpyHostToDevice, b->streams[0]));
    rc = gpu_render_frame_dev(b, cam, width, height, y0, y1, max_depth, seed, d_samples, d_rgba, b->streams[0]);
    if (rc != SAHBVH_OK) return rc;
    CU_TRY(cudaMemcpyAsync(rgba_out, d_rgba, (size_t)n_pixels * 4, cudaMemcpyDeviceToHost, b->streams[0]));
    CU_TRY(cudaStreamSynchronize(b->streams[0]));
    return SAHBVH_OK;
}

extern "C" int gpu_accumulate_dev(void *d_accum, int frames, const void *d_rgba, int64_t n_pixels, void *d_avg_rgba, void *stream)
{
    if (!d_accum || !d_rgba || !d_avg_rgba || frames < 1 || n_pixels < 0) {
        sahbvh_set_error("gpu_accumulate: bad arguments");
        return SAHBVH_E_ARG;
    }
    if (n_pixels == 0) return SAHBVH_OK;
    k_accumulate<<<(unsigned)((n_pixels + RN_THREADS - 1) / RN_THREADS), RN_THREADS, 0, (cudaStream_t)stream>>>(
        (float *)d_accum, frames, (const uint32_t *)d_rgba, n_pixels, (uint32_t *)d_avg_rgba);
    g_launches++;
    CU_TRY(cudaGetLastError());
    return SAHBVH_OK;
}

extern "C" int gpu_accumulate(float *accum, int frames, const uint8_t *rgba, int64_t n_pixels, uint8_t *avg_rgba)
{
    if (!accum || !rgba || !avg_rgba || frames < 1 || n_pixels < 0) {
        sahbvh_set_error("gpu_accumulate: bad arguments");
        return SAHBVH_E_ARG;
    }
    if (n_pixels == 0) return SAHBVH_OK;
    if (sahbvh_device_count() == 0) {
        sahbvh_set_error("no usable CUDA device");
        return SAHBVH_E_CUDA;
    }
    float *d_acc = nullptr;
    uint32_t *d_rgba = nullptr, *d_avg = nullptr;
    int rc = SAHBVH_OK;
    cudaError_t e = cudaMalloc(&d_acc, (size_t)n_pixels * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_rgba, (size_t)n_pixels * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_avg, (size_t)n_pixels * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_acc, accum, (size_t)n_pixels * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_rgba, rgba, (size_t)n_pixels * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) rc = gpu_accumulate_dev(d_acc, frames, d_rgba, n_pixels, d_avg, nullptr);
    if (e == cudaSuccess && rc == SAHBVH_OK) e = cudaMemcpy(accum, d_acc, (size_t)n_pixels * 12, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == SAHBVH_OK) e = cudaMemcpy(avg_rgba, d_avg, (size_t)n_pixels * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_acc); cudaFree(d_rgba); cudaFree(d_avg);
    if (e != cudaSuccess) {
        sahbvh_set_error("gpu_accumulate: %s", cu_errstr(e));
        return SAHBVH_E_CUDA;
    }
    return rc;
}
