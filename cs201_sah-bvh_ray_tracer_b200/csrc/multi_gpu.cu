/*
 * multi_gpu.cu -- one process, N devices, behind the C-ABI (SURVEY 8e / 8b "flatten + upload + replicate").
 *
 * The reference's two drivers are single C loops over pixels (main.c:356-366) and rays (benchmark.c:230-245).
 * Rays are independent and the tree is read-only, so the loop shards: the BVH is REPLICATED on every device
 * (gpu_bvh_replicate: device-to-device copies of the flattened arrays over NVLink, nothing is rebuilt) and every
 * device takes a contiguous band of image rows / range of rays.  No collective runs anywhere: the frame is
 * assembled by the bands' own stores -- into the caller's host buffer (one D2H stream per device), or into ONE
 * device's frame buffer through peer memory (the trace kernel's row stores travel over NVLink while it computes;
 * CUDA events are the completion flags the root stream waits on).
 *
 * A gpu_bvh_set owns one worker thread per device, so the N devices are driven concurrently (a single host
 * thread would serialise ~10 us of launch calls per device, which is the whole kernel time of a 4K band on 8 GPUs).
 */
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "gpu_handle.cuh"

/* ---- sharding rules (also used by the one-process-per-GPU harness, bench.py) ---- */
extern "C" void sahbvh_shard_range(int64_t n, int part, int n_parts, int64_t *lo, int64_t *hi)
{
    if (n_parts < 1) n_parts = 1;
    *lo = n * part / n_parts;
    *hi = n * (part + 1) / n_parts;
}

/* row band [y0, y1) of part `part`: whole 16-row CTA groups (the packet kernel's 32 x 16 pixel block), so no block
 * straddles two devices; the last band takes the remainder */
extern "C" void sahbvh_shard_rows(int height, int part, int n_parts, int *y0, int *y1)
{
    const int align = 16;
    if (n_parts < 1) n_parts = 1;
    const long long blocks = (height + align - 1) / align;
    const long long b0 = blocks * part / n_parts, b1 = blocks * (part + 1) / n_parts;
    *y0 = (int)(b0 * align < height ? b0 * align : height);
    *y1 = (int)(b1 * align < height ? b1 * align : height);
}

/* Uneven bands for a frame loop whose ROOT (part 0) also consumes the assembled frame: the root's band is
 * root_share (0 < root_share <= 1) of an equal share, the other parts split the rest evenly; same 16-row alignment.
 * With t = time of an equal band and c = what consuming one frame costs the root, root_share = 1 - (c/t)(N-1)/N
 * makes every device finish a frame at the same time (bench.py calibrates it in its warm-up). */
extern "C" void sahbvh_shard_rows_weighted(int height, int part, int n_parts, double root_share, int *y0, int *y1)
{
    const int align = 16;
    if (n_parts < 2 || !(root_share > 0.0) || root_share >= 1.0) {
        sahbvh_shard_rows(height, part, n_parts, y0, y1);
        return;
    }
    const long long blocks = (height + align - 1) / align;
    long long b_root = (long long)((double)blocks / n_parts * root_share + 0.5);
    if (b_root < 1) b_root = blocks >= n_parts ? 1 : 0;
    if (b_root > blocks) b_root = blocks;
    const long long rest = blocks - b_root;
    long long b0, b1;
    if (part == 0) {
        b0 = 0;
        b1 = b_root;
    } else {
        b0 = b_root + rest * (part - 1) / (n_parts - 1);
        b1 = b_root + rest * part / (n_parts - 1);
    }
    *y0 = (int)(b0 * align < height ? b0 * align : height);
    *y1 = (int)(b1 * align < height ? b1 * align : height);
}

/* ---- worker: one thread pinned to one device of the set ---- */
struct SetWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
    std::string err;

    void loop()
    {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            std::function<int()> f = job;
            lk.unlock();
            int r = f();
            std::string e = r == SAHBVH_OK ? std::string() : std::string(sahbvh_last_error());
            lk.lock();
            rc = r;
            err = e;
            has_job = false;
            done = true;
            cv.notify_all();
        }
    }
};

struct gpu_bvh_set {
    int n;
    std::vector<gpu_bvh *> h;
    std::vector<int> dev;
    std::vector<cudaEvent_t> ev;
    std::vector<SetWorker *> w;      /* w[0] unused: the caller's thread drives device 0 */
    std::vector<char> peer_to_root;  /* device i can store into devices[0]'s memory */
};

/* run fn(i) for every member concurrently; returns the first error (text copied to the caller's thread) */
static int set_run(gpu_bvh_set *s, const std::function<int(int)> &fn)
{
    for (int i = 1; i < s->n; i++) {
        SetWorker *w = s->w[i];
        std::lock_guard<std::mutex> lk(w->mu);
        w->job = [fn, i] { return fn(i); };
        w->done = false;
        w->has_job = true;
        w->cv.notify_all();
    }
    int rc = fn(0);
    std::string err = rc == SAHBVH_OK ? std::string() : std::string(sahbvh_last_error());
    for (int i = 1; i < s->n; i++) {
        SetWorker *w = s->w[i];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
        if (rc == SAHBVH_OK && w->rc != SAHBVH_OK) {
            rc = w->rc;
            err = w->err;
        }
    }
    if (rc != SAHBVH_OK) sahbvh_set_error("device set: %s", err.c_str());
    return rc;
}

/* copy of a handle's device arrays on another device (cudaMemcpyPeer: NVLink when peers, staged otherwise) */
static int clone_handle(const gpu_bvh *src, int device, gpu_bvh **out)
{
    *out = nullptr;
    DeviceGuard guard(device);
    if (!guard.ok) {
        sahbvh_set_error("cudaSetDevice(%d) failed", device);
        return SAHBVH_E_CUDA;
    }
    gpu_bvh *b = (gpu_bvh *)calloc(1, sizeof(gpu_bvh));
    if (!b) return SAHBVH_E_NOMEM;
    b->device = device;
    b->info = src->info;
    b->info.device = device;
    b->n_wide = src->n_wide;
    b->n_special = src->n_special;
    const size_t n1 = (size_t)src->info.n_spheres + 1;
    struct Part { void **dst; const void *from; size_t bytes; } parts[] = {
        {(void **)&b->d_nodes, src->d_nodes, src->d_nodes ? (size_t)src->info.n_inner * 64 : 0},
        {(void **)&b->d_spheres4, src->d_spheres4, n1 * 16},
        {(void **)&b->d_colors, src->d_colors, n1 * 4},
        {(void **)&b->d_wide, src->d_wide, (size_t)src->n_wide * 64},
        {(void **)&b->d_wide_full, src->d_wide_full, (size_t)src->n_wide * 128},
        {(void **)&b->d_special, src->d_special, (size_t)(src->n_special > 0 ? src->n_special : 1) * 32},
    };
    cudaError_t err = cudaSuccess;
    do {
        if ((err = cudaStreamCreateWithFlags(&b->streams[0], cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((err = cudaStreamCreateWithFlags(&b->streams[1], cudaStreamNonBlocking)) != cudaSuccess) break;
        for (const Part &p : parts) {
            if (!p.from || p.bytes == 0) continue;
            if ((err = cudaMalloc(p.dst, p.bytes)) != cudaSuccess) break;
            if ((err = cudaMemcpyPeerAsync(*p.dst, device, p.from, src->device, p.bytes, b->streams[0])) != cudaSuccess) break;
        }
        if (err != cudaSuccess) break;
        if ((err = cudaMalloc(&b->d_counters, 72 * sizeof(unsigned long long))) != cudaSuccess) break;
        if ((err = cudaMemsetAsync(b->d_counters, 0, 72 * sizeof(unsigned long long), b->streams[0])) != cudaSuccess) break;
        if ((err = cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) break;
        if ((err = cudaStreamSynchronize(b->streams[0])) != cudaSuccess) break;
    } while (0);
    if (err != cudaSuccess) {
        sahbvh_set_error("gpu_bvh_replicate: device %d: %s", device, cu_errstr(err));
        gpu_bvh_free(b);
        return SAHBVH_E_CUDA;
    }
    *out = b;
    return SAHBVH_OK;
}

extern "C" void gpu_bvh_set_free(gpu_bvh_set *s)
{
    if (!s) return;
    for (size_t i = 1; i < s->w.size(); i++) {
        SetWorker *w = s->w[i];
        if (!w) continue;
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->quit = true;
            w->cv.notify_all();
        }
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (size_t i = 0; i < s->h.size(); i++) {
        if (s->ev[i]) {
            DeviceGuard guard(s->dev[i]);
            cudaEventDestroy(s->ev[i]);
        }
        gpu_bvh_free(s->h[i]);
    }
    delete s;
}

extern "C" int gpu_bvh_replicate(const gpu_bvh *src, const int *devices, int n_devices, gpu_bvh_set **out)
{
    if (!src || !out || n_devices < 1 || n_devices > 64) {
        sahbvh_set_error("gpu_bvh_replicate: bad arguments");
        return SAHBVH_E_ARG;
    }
    *out = nullptr;
    const int ndev = sahbvh_device_count();
    if (ndev == 0) {
        sahbvh_set_error("no usable CUDA device; this library has no CPU traversal path");
        return SAHBVH_E_CUDA;
    }
    gpu_bvh_set *s = new gpu_bvh_set();
    s->n = n_devices;
    s->h.assign(n_devices, nullptr);
    s->dev.assign(n_devices, 0);
    s->ev.assign(n_devices, nullptr);
    s->w.assign(n_devices, nullptr);
    s->peer_to_root.assign(n_devices, 0);
    for (int i = 0; i < n_devices; i++) {
        const int d = devices ? devices[i] : i;           /* devices == NULL: ordinals 0 .. n-1 */
        if (d < 0 || d >= ndev) {
            sahbvh_set_error("gpu_bvh_replicate: device %d out of range (%d devices)", d, ndev);
            gpu_bvh_set_free(s);
            return SAHBVH_E_ARG;
        }
        s->dev[i] = d;
    }
    int rc = SAHBVH_OK;
    for (int i = 0; i < n_devices && rc == SAHBVH_OK; i++) {
        rc = clone_handle(src, s->dev[i], &s->h[i]);
        if (rc != SAHBVH_OK) break;
        DeviceGuard guard(s->dev[i]);
        if (cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            sahbvh_set_error("gpu_bvh_replicate: cudaEventCreate failed on device %d", s->dev[i]);
            rc = SAHBVH_E_CUDA;
            break;
        }
        /* peer stores into the root's frame buffer (gpu_trace_primary_prims_multi_dev) */
        if (s->dev[i] == s->dev[0]) {
            s->peer_to_root[i] = 1;
        } else {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, s->dev[i], s->dev[0]) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(s->dev[0], 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) s->peer_to_root[i] = 1;
            }
            cudaGetLastError();
        }
    }
    if (rc != SAHBVH_OK) {
        gpu_bvh_set_free(s);
        return rc;
    }
    for (int i = 1; i < n_devices; i++) {
        s->w[i] = new SetWorker();
        SetWorker *w = s->w[i];
        w->th = std::thread([w] { w->loop(); });
    }
    *out = s;
    return SAHBVH_OK;
}

extern "C" int gpu_bvh_build_replicated(Sphere *spheres, int n_total, int start, int end, int depth,
                                        const int *devices, int n_devices, gpu_bvh_set **out)
{
    if (!out || n_devices < 1) {
        sahbvh_set_error("gpu_bvh_build_replicated: bad arguments");
        return SAHBVH_E_ARG;
    }
    *out = nullptr;
    gpu_bvh *first = nullptr;
    int rc = gpu_bvh_build(spheres, n_total, start, end, depth, devices ? devices[0] : 0, &first);
    if (rc != SAHBVH_OK) return rc;
    rc = gpu_bvh_replicate(first, devices, n_devices, out);
    gpu_bvh_free(first);
    return rc;
}

extern "C" int gpu_bvh_set_size(const gpu_bvh_set *s) { return s ? s->n : 0; }

extern "C" gpu_bvh *gpu_bvh_set_get(gpu_bvh_set *s, int i)
{
    if (!s || i < 0 || i >= s->n) return nullptr;
    return s->h[i];
}

static int check_set_frame(const gpu_bvh_set *s, const Camera *cam, int width, int height, const void *out)
{
    if (!s || !cam || !out || width <= 0 || height <= 0) {
        sahbvh_set_error("bad frame arguments for a device set (width=%d height=%d)", width, height);
        return SAHBVH_E_ARG;
    }
    return SAHBVH_OK;
}

/* every device traces its row band and copies it into the caller's host frame */
static int primary_multi_host(gpu_bvh_set *s, const Camera *cam, int width, int height, void *out, int fmt, size_t px_bytes)
{
    int rc = check_set_frame(s, cam, width, height, out);
    if (rc != SAHBVH_OK) return rc;
    return set_run(s, [=](int i) -> int {
        int y0, y1;
        sahbvh_shard_rows(height, i, s->n, &y0, &y1);
        if (y1 <= y0) return SAHBVH_OK;
        int r = sahbvh_primary_host_enqueue(s->h[i], cam, width, height, y0, y1, (char *)out + (size_t)y0 * (size_t)width * px_bytes, fmt);
        if (r != SAHBVH_OK) return r;
        return sahbvh_primary_host_wait(s->h[i]);
    });
}

extern "C" int gpu_trace_primary_multi(gpu_bvh_set *s, const Camera *cam, int width, int height, gpu_hit *out)
{
    return primary_multi_host(s, cam, width, height, out, HOSTFMT_RECORDS, sizeof(gpu_hit));
}

extern "C" int gpu_trace_primary_prims_multi(gpu_bvh_set *s, const Camera *cam, int width, int height, int32_t *prims)
{
    return primary_multi_host(s, cam, width, height, prims, HOSTFMT_PRIMS, 4);
}

/* The id frame assembled in ONE device's memory: d_prims lives on the set's first device; every other device's trace
 * kernel stores its band's rows straight into it through peer memory.  Asynchronous: on return the bands are in
 * flight; `root_stream` (a stream of the first device, NULL = its default stream) has been made to wait for every
 * band's completion event, so work enqueued on it afterwards (gpu_hits_from_prims_dev, a copy, a consumer kernel)
 * sees the whole frame. */
extern "C" int gpu_trace_primary_prims_multi_dev(gpu_bvh_set *s, const Camera *cam, int width, int height, void *d_prims,
                                                 void *root_stream)
{
    int rc = check_set_frame(s, cam, width, height, d_prims);
    if (rc != SAHBVH_OK) return rc;
    for (int i = 0; i < s->n; i++)
        if (!s->peer_to_root[i]) {
            sahbvh_set_error("device %d cannot store into device %d's memory (no peer access): use gpu_trace_primary_prims_multi", s->dev[i],
                             s->dev[0]);
            return SAHBVH_E_CUDA;
        }
    rc = set_run(s, [=](int i) -> int {
        int y0, y1;
        sahbvh_shard_rows(height, i, s->n, &y0, &y1);
        gpu_bvh *b = s->h[i];
        DeviceGuard guard(b->device);
        if (y1 > y0) {
            int r = sahbvh_launch_primary(b, cam, width, height, y0, y1, nullptr, b->streams[0], false,
                                          (char *)d_prims + (size_t)y0 * (size_t)width * 4, SAH_IDS_PRIM);
            if (r != SAHBVH_OK) return r;
        }
        CU_TRY(cudaEventRecord(s->ev[i], b->streams[0]));
        return SAHBVH_OK;
    });
    if (rc != SAHBVH_OK) return rc;
    DeviceGuard guard(s->dev[0]);
    for (int i = 0; i < s->n; i++) CU_TRY(cudaStreamWaitEvent((cudaStream_t)root_stream, s->ev[i], 0));
    return SAHBVH_OK;
}

/* ray buffers: contiguous ranges, host buffers */
extern "C" int gpu_trace_closest_hit_multi(gpu_bvh_set *s, const Ray *rays, int64_t n, gpu_hit *out)
{
    if (!s || n < 0 || (n > 0 && (!rays || !out))) {
        sahbvh_set_error("gpu_trace_closest_hit_multi: bad arguments");
        return SAHBVH_E_ARG;
    }
    return set_run(s, [=](int i) -> int {
        int64_t lo, hi;
        sahbvh_shard_range(n, i, s->n, &lo, &hi);
        return hi > lo ? gpu_trace_closest_hit(s->h[i], rays + lo, hi - lo, out + lo) : SAHBVH_OK;
    });
}

extern "C" int gpu_trace_closest_hit_prims_multi(gpu_bvh_set *s, const Ray *rays, int64_t n, int32_t *prims)
{
    if (!s || n < 0 || (n > 0 && (!rays || !prims))) {
        sahbvh_set_error("gpu_trace_closest_hit_prims_multi: bad arguments");
        return SAHBVH_E_ARG;
    }
    return set_run(s, [=](int i) -> int {
        int64_t lo, hi;
        sahbvh_shard_range(n, i, s->n, &lo, &hi);
        return hi > lo ? gpu_trace_closest_hit_prims(s->h[i], rays + lo, hi - lo, prims + lo) : SAHBVH_OK;
    });
}

/* benchmark.c:225-254 over N devices: the count is the sum of the ranges' counts */
extern "C" int gpu_benchmark_with_bvh_multi(gpu_bvh_set *s, const Ray *rays, int64_t n, int64_t *intersections)
{
    if (!s || n < 0 || !intersections || (n > 0 && !rays)) {
        sahbvh_set_error("gpu_benchmark_with_bvh_multi: bad arguments");
        return SAHBVH_E_ARG;
    }
    std::vector<int64_t> part((size_t)s->n, 0);
    int64_t *pp = part.data();
    int rc = set_run(s, [=](int i) -> int {
        int64_t lo, hi;
        sahbvh_shard_range(n, i, s->n, &lo, &hi);
        return hi > lo ? gpu_benchmark_with_bvh(s->h[i], rays + lo, hi - lo, pp + i) : SAHBVH_OK;
    });
    *intersections = 0;
    for (int i = 0; i < s->n; i++) *intersections += part[(size_t)i];
    return rc;
}

/* the path loop (renderer.c:21-76) per row band; samples (may be NULL) are indexed like the whole frame's */
extern "C" int gpu_render_frame_multi(gpu_bvh_set *s, const Camera *cam, int width, int height, int max_depth, uint64_t seed,
                                      const Vec3 *samples, uint8_t *rgba_out)
{
    int rc = check_set_frame(s, cam, width, height, rgba_out);
    if (rc != SAHBVH_OK) return rc;
    return set_run(s, [=](int i) -> int {
        int y0, y1;
        sahbvh_shard_rows(height, i, s->n, &y0, &y1);
        if (y1 <= y0) return SAHBVH_OK;
        const size_t off = (size_t)y0 * (size_t)width;
        return gpu_render_frame(s->h[i], cam, width, height, y0, y1, max_depth, seed,
                                     samples ? samples + off * (size_t)(max_depth > 0 ? max_depth : 0) : nullptr, rgba_out + off * 4);
    });
}
