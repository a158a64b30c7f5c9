/*
 * sahbvh_gpu.h -- the batch C-ABI of libsahbvh_b200.so (plain pointers and sizes only).
 *
 * These are the entry points a maintainer of the reference binds in place of the scalar loops
 *   for ray: ray_bvh_intersect(ray, root)            (benchmark.c:230-245, renderer.c:29-32)
 *   for y,x: get_camera_ray(...) + trace              (main.c:356-366)
 * See INTEGRATION.md for the reference-side patch.
 *
 * Status codes: 0 = ok, negative = error (SAHBVH_E_*); sahbvh_last_error() returns a
 * thread-local description.  Nothing here aborts and nothing falls back to the CPU: without a
 * usable CUDA device every call returns SAHBVH_E_CUDA.
 */
#ifndef SAHBVH_GPU_H
#define SAHBVH_GPU_H

#include "sahbvh_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SAHBVH_OK 0
#define SAHBVH_E_ARG (-1)      /* bad argument (NULL, negative size, pointer outside the sphere array) */
#define SAHBVH_E_CUDA (-2)     /* CUDA runtime error / no device */
#define SAHBVH_E_NOMEM (-3)    /* host allocation failed */
#define SAHBVH_E_DEPTH (-4)    /* tree deeper than the traversal stack supports */

/* 32-byte hit record.  Replaces HitRecord (hit.h:8-14) for batches: `prim` is
 * hit.object - spheres (the POST-BUILD index) instead of a pointer;
 * a miss is {t=0, prim=-1, point=0, normal=0} (hit.c:20,92 all-zero record, Q8). */
typedef struct gpu_hit {
    float t;
    int32_t prim;
    float px, py, pz;
    float nx, ny, nz;
} gpu_hit;
SAHBVH_STATIC_ASSERT(sizeof(gpu_hit) == 32, "gpu_hit layout");

/* opaque: flattened tree + SoA spheres resident on ONE device.  A handle is NOT thread-safe: it owns per-launch
 * ray cursors, counters and staging buffers, so at most one host thread may be inside a call on a given handle
 * at a time, and the asynchronous *_dev calls must be issued in order from one thread (use one handle per thread
 * or per device -- gpu_bvh_replicate -- for concurrency).  Different handles are independent. */
typedef struct gpu_bvh gpu_bvh;

typedef struct gpu_bvh_info {
    int32_t device;
    int32_t n_spheres;
    int32_t n_inner;         /* flattened 64-byte inner nodes */
    int32_t n_leaf_refs;     /* leaves folded into child references */
    int32_t max_depth;       /* levels below the root */
    int32_t multi_leaves;    /* leaves with sphere_count > 1 (Q2) */
    int32_t empty_leaves;    /* leaves with sphere_count == 0 (Q3) */
    int32_t n_wide;          /* 64-byte nodes of the 4-wide quantised tree (ray-buffer kernel) */
    float max_radius;
    float scene_lo[3], scene_hi[3];
    int64_t device_bytes;
} gpu_bvh_info;

/* work counters of one traced batch (instrumented kernel variant; not the timed path) */
typedef struct gpu_trace_stats {
    uint64_t rays;
    uint64_t inner_visits;   /* 64-byte nodes fetched            (V_i of SURVEY 8d) */
    uint64_t sphere_tests;   /* 16-byte spheres fetched + tested (V_s of SURVEY 8d) */
    uint64_t exact_box_tests;
    uint64_t hits;
    uint64_t max_stack;
    uint64_t max_visits;     /* most nodes any single ray fetched (ray-buffer kernel over the wide tree; 0 elsewhere) */
} gpu_trace_stats;

const char *sahbvh_last_error(void);
int sahbvh_device_count(void);
const char *sahbvh_version(void);

/* Flatten the pointer tree produced by build_bvh_node (reference's or ours) and upload it.
 *   permuted : the sphere array AFTER the build permuted it (bvh.c:172-201), n entries
 *   root     : tree whose leaf `sphere` pointers alias `permuted` (bvh.c:133)
 *   device   : CUDA ordinal, or -1 for the current device
 * The handle owns copies; host arrays may be freed after the call. */
int gpu_bvh_upload(const Sphere *permuted, int n, const BVHNode *root, int device, gpu_bvh **out);

/* Build (binned SAH, identical topology, bounds and permutation to bvh.c:117-209) directly into
 * the flattened form, skipping the per-node mallocs, then upload.  Permutes spheres[start,end)
 * in place exactly like build_bvh_node; n_total is the array length (indices are global).
 * Ranges of >= 4096 spheres are built ON THE DEVICE (level-parallel binning, exact replay of the
 * reference's swap order), smaller ones by the host builder; sahbvh_set_build_mode selects. */
int gpu_bvh_build(Sphere *spheres, int n_total, int start, int end, int depth, int device, gpu_bvh **out);
#define SAHBVH_BUILD_HOST 0
#define SAHBVH_BUILD_DEVICE 1
int sahbvh_set_build_mode(int mode);

/* The built tree as a preorder array (root = 0), one entry per reference BVHNode, for callers that want
 * the tree itself rather than a device handle (and for checking the device builder against bvh.c):
 * runs the selected builder (mode = SAHBVH_BUILD_HOST / SAHBVH_BUILD_DEVICE) over spheres[start,end),
 * permutes the array exactly like build_bvh_node and writes up to `cap` nodes.  Returns the node
 * count (> cap means the array was too small and nothing was written) or a negative error code. */
typedef struct sahbvh_tree_node {
    float lo[3], hi[3];   /* node->bounds (bvh.c:120-125) */
    int32_t left, right;  /* preorder ids of the children; -1 on leaves */
    int32_t sphere;       /* leaf: index of spheres[start], the only sphere hit.c:98-99 tests; -1 on inner nodes */
    int32_t count;        /* leaf: sphere_count */
} sahbvh_tree_node;
int64_t sahbvh_build_tree(Sphere *spheres, int start, int end, int depth, int mode, int device,
                          sahbvh_tree_node *out, int64_t cap);

void gpu_bvh_free(gpu_bvh *bvh);
/* Diagnostic copy of the handle's device arrays to the host.  what: 0 binary 64-byte nodes, 1 float4 spheres,
 * 2 RGBA8 colours, 3 4-wide quantised nodes (64 B), 4 their full-precision twins (128 B), 5 special-leaf records
 * (32 B).  Returns the array's size in bytes (negative = error) and copies min(size, cap) bytes when out != NULL. */
int64_t gpu_bvh_download(const gpu_bvh *bvh, int what, void *out, int64_t cap);
int gpu_bvh_get_info(const gpu_bvh *bvh, gpu_bvh_info *info);

/* Closest hit for a ray buffer.  HOST buffers: H2D copy of rays, kernel, D2H copy of hits,
 * chunked over two streams.  Replaces the loop at benchmark.c:230-245 / renderer.c:29-32. */
int gpu_trace_closest_hit(gpu_bvh *bvh, const Ray *rays, int64_t n, gpu_hit *out);
/* Same with DEVICE buffers on `stream` (a cudaStream_t, NULL = default); asynchronous. */
int gpu_trace_closest_hit_dev(gpu_bvh *bvh, const void *d_rays, int64_t n, void *d_out, void *stream);
/* Compact output: only `hit.object - spheres` (hit.h:13) per ray, 4 bytes instead of 32 (-1 = miss).  The other
 * HitRecord fields are pure functions of the ray and that sphere: sahbvh_hit_from_prim (sahbvh_ref_api.h) derives
 * them on the host for the rays that need them, bit for bit.  Over PCIe this moves 24 + 4 bytes per ray, not 24 + 32. */
int gpu_trace_closest_hit_prims(gpu_bvh *bvh, const Ray *rays, int64_t n, int32_t *prims);
int gpu_trace_closest_hit_dev_prims(gpu_bvh *bvh, const void *d_rays, int64_t n, void *d_prims, void *stream);

/* Fused ray generation + closest hit for image rows [y0,y1) of a width x height frame:
 * pixel (x,y) -> u=((float)x/W-0.5f)*aspect, v=(float)y/H-0.5f, get_camera_ray(cam,u,-v)
 * (main.c:356-365 + ray.c:17-32, bit-exact), hit written at out[(y-y0)*width + x]. */
int gpu_trace_primary(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1, gpu_hit *out);
/* The same frame with compact per-pixel output into a HOST buffer -- what renderer.c:29-58 / benchmark.c:242-244
 * actually consume is the hit OBJECT (hit.h:13): 4 bytes per pixel (`prims`: post-build sphere index, -1 = miss)
 * or 8 bytes ({t, prim}).  PCIe carries 4 or 8 bytes per ray instead of 32; the remaining HitRecord fields are
 * derived on demand, bit-identical to the traced record:
 *   one pixel, on the host : sahbvh_hit_from_prim(sahbvh_pixel_ray(cam, W, H, x, y), spheres, prim)
 *   a whole frame          : gpu_hits_from_prims (host buffers) / gpu_hits_from_prims_dev (device buffers)
 * Page-locked output buffers (sahbvh_host_alloc / sahbvh_host_pin) run at PCIe speed. */
typedef struct gpu_tprim {
    float t;
    int32_t prim;
} gpu_tprim;
int gpu_trace_primary_prims(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1, int32_t *prims);
int gpu_trace_primary_tprim(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1, gpu_tprim *out);
int gpu_hits_from_prims(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1,
                        const int32_t *prims, gpu_hit *out);
int gpu_trace_primary_dev(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1,
                          void *d_out, void *stream);
/* Same, and additionally writes the compact 8-byte "ID buffer" record {float t, int32 prim} per pixel
 * to d_ids (may be NULL).  Multi-GPU frame assembly gathers THIS buffer (point and normal are pure
 * functions of the pixel's ray, t and the replicated sphere array): the kernel writes a rank's band
 * straight into its slot of the all-gather buffer, so no pack/copy kernel sits between them. */
int gpu_trace_primary_dev_ids(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1,
                              void *d_out, void *d_ids, void *stream);
/* The 4-byte variant: d_prims receives only the int32 primitive id per pixel (-1 = miss).  t, point and normal
 * are pure functions of the pixel's ray and the hit sphere, so gpu_hits_from_prims_dev rebuilds the full,
 * bit-identical 32-byte records from the ids on whichever GPU needs them: the multi-GPU frame assembly
 * gathers 4 bytes per pixel over NVLink instead of 32. */
/* d_out may be NULL in both *_dev_ids and *_dev_prims: then no 32-byte records are written (or computed) at all. */
int gpu_trace_primary_dev_prims(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1,
                                void *d_out, void *d_prims, void *stream);
int gpu_hits_from_prims_dev(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1,
                            const void *d_prims, void *d_out, void *stream);
/* Multi-GPU frame assembly without a separate collective: one GPU (the root) allocates the frame's id buffer
 * with sahbvh_peer_alloc and hands the 64-byte handle to the other processes (any byte transport); they map it
 * with sahbvh_peer_open and pass  mapped + band offset  as d_prims to gpu_trace_primary_dev_prims.  The trace
 * kernel's own 4-byte stores then travel over NVLink / NVSwitch peer memory while it computes, and the frame
 * is complete on the root when the ranks' kernels are (one small cross-rank barrier per frame). */
typedef struct sahbvh_ipc_handle { unsigned char bytes[64]; } sahbvh_ipc_handle;
int sahbvh_peer_alloc(int device, size_t bytes, void **d_ptr, sahbvh_ipc_handle *handle);
int sahbvh_peer_open(int device, const sahbvh_ipc_handle *handle, void **d_ptr);
/* Completion flags in the same peer memory: after its trace launch a rank stores the frame number into ITS 64-bit
 * slot of a flag array owned by the root (a one-thread kernel ordered on the same stream, system-scope fence before
 * the store); the root reads the slots when it wants to consume a frame.  No collective runs in the frame loop. */
int sahbvh_peer_signal(void *d_flag, uint64_t value, void *stream);
int sahbvh_peer_read_flags(const void *d_flags, int n, uint64_t *out, void *stream);
/* Device-side wait (no host round trip): `stream` does not advance until all n 64-bit flags at d_flags are >= value.
 * With sahbvh_peer_signal this is the frame loop's back-pressure -- the owner consumes frame k once every band's
 * flag says k, a producer reuses a frame buffer once the owner's "consumed" flag allows it.  Gives up after
 * timeout_ms (0 = 2000) and stores 1 into the 32-bit word *d_status (may be NULL) instead of hanging the stream.
 * The flags waited for must be written by ANOTHER GPU or by the host: two streams of one device may share a hardware
 * queue, so a wait ahead of the same device's signalling kernel is not guaranteed to make progress -- order work of
 * one device with cudaEvents (bench.py's FrameGather and gpu_trace_primary_prims_multi_dev do). */
int sahbvh_peer_wait(const void *d_flags, int n, uint64_t value, uint32_t timeout_ms, void *d_status, void *stream);
/* benchmark.c:242-244 over an id buffer that is already on the device: *d_count (uint64) += #(ids >= 0).  Asynchronous. */
int gpu_count_prims_dev(const void *d_prims, int64_t n, void *d_count, void *stream);
int sahbvh_peer_close(int device, void *d_ptr);   /* for pointers from sahbvh_peer_open */
int sahbvh_peer_free(int device, void *d_ptr);    /* for pointers from sahbvh_peer_alloc */
/* Page-locked host buffers for the host-pointer entry points (gpu_trace_closest_hit, gpu_trace_primary, ...):
 * their copies run at PCIe speed (~54 GB/s measured) only when the caller's ray / hit arrays are page-locked;
 * with malloc'ed memory the driver stages every copy through its own bounce buffer.  A C caller without the CUDA
 * headers gets page-locked memory here: sahbvh_host_alloc/free for new buffers, sahbvh_host_pin/unpin around an
 * existing array (e.g. the static frame buffer of main.c:356-366) for as long as it is passed to the library. */
int sahbvh_host_alloc(size_t bytes, void **ptr);
int sahbvh_host_free(void *ptr);
int sahbvh_host_pin(void *ptr, size_t bytes);
int sahbvh_host_unpin(void *ptr);
/* Which kernel gpu_trace_primary* launches (process-wide; all give identical hit records):
 *   SAHBVH_PRIMARY_PACKET  one warp per 16x8 pixel tile, interior culling once per tile
 *   SAHBVH_PRIMARY_PER_RAY one thread per pixel, per-ray stack traversal (the ray-buffer kernel's core) */
#define SAHBVH_PRIMARY_PER_RAY 0
#define SAHBVH_PRIMARY_PACKET 1
/*   SAHBVH_PRIMARY_MACRO   one warp per 32x8 pixel macro tile: interior culling once per 256 rays, a candidate leaf is
 *                          first tested against the frusta of the tile's eight 8x4-pixel quadrants (one lane each) and
 *                          only the rays of the quadrants it can touch run the exact tests
 *   SAHBVH_PRIMARY_AUTO    (default) PACKET or MACRO, whichever the first launches on this handle and frame shape
 *                          measure as faster (event pairs around the launches, queried later; nothing synchronises) */
#define SAHBVH_PRIMARY_MACRO 2
#define SAHBVH_PRIMARY_AUTO 3
int sahbvh_set_primary_mode(int mode);   /* the environment variable SAHBVH_PRIMARY_MODE=0..3 sets the initial mode */
/* the kernel in use for this handle and frame shape (rows = y1 - y0): one of the modes above, or -1 while
 * SAHBVH_PRIMARY_AUTO is still sampling (its first six launches of that shape) */
int gpu_primary_kernel_in_use(const gpu_bvh *bvh, int width, int rows);
/* Which kernel gpu_trace_closest_hit* launches (identical hit records):
 *   SAHBVH_RAYS_WIDE     (default) persistent warps over the 4-wide, 8-bit quantised tree (64-byte
 *                        nodes, two binary levels per fetch), shared-memory stack, batched ray refill
 *   SAHBVH_RAYS_COOP     persistent warps over the binary tree, cooperative 4-lanes-per-node fetch
 *   SAHBVH_RAYS_ONE_SHOT one thread per ray over the binary tree, grid-stride */
#define SAHBVH_RAYS_ONE_SHOT 0
#define SAHBVH_RAYS_COOP 1
#define SAHBVH_RAYS_WIDE 2
int sahbvh_set_ray_mode(int mode);
/* The 64-byte binary tree is read only by SAHBVH_PRIMARY_PER_RAY, SAHBVH_RAYS_COOP and SAHBVH_RAYS_ONE_SHOT.
 * Handles created while keep == 0 (the default) do not hold it (64 MB per million spheres); those modes then
 * return SAHBVH_E_ARG on such a handle.  Set keep = 1 BEFORE gpu_bvh_upload / gpu_bvh_build to use them. */
int sahbvh_set_keep_binary_tree(int keep);
/* Top-of-tree residency (process-wide, default 0 = off).  The wide trees are numbered breadth first, so their top
 * levels are a prefix of the node array: with bytes > 0 every trace launch runs with a persisting-L2 access-policy
 * window over the first `bytes` of the tree it reads (clamped to the tree, to cudaDevAttrMaxPersistingL2CacheSize and
 * to cudaDevAttrMaxAccessPolicyWindowSize) and the rest of its traffic marked streaming.  Only worth it when the tree
 * is larger than the L2 (10 M spheres: 320 MB of quantised nodes); results never change. */
int sahbvh_set_l2_persist_bytes(uint64_t bytes);
/* Ray generation only (parity aid for ray.c): writes Ray records for rows [y0,y1). */
int gpu_generate_primary_rays(const Camera *cam, int width, int height, int y0, int y1, Ray *out, int device);

/* Brute-force closest hit over all n spheres (renderer.c:33-44 semantics: ties -> LOWEST index).
 * Parity aid for small n. */
int gpu_trace_brute_force(gpu_bvh *bvh, const Ray *rays, int64_t n, gpu_hit *out);

/* benchmark.c's two counting loops on the device, rays supplied by the caller:
 *   gpu_benchmark_no_bvh   = benchmark_no_bvh  (benchmark.c:172-223): every ray against EVERY sphere of the
 *                            handle, no early exit; *intersections = rays with at least one hit_something
 *   gpu_benchmark_with_bvh = benchmark_with_bvh (benchmark.c:225-254): *intersections = rays whose
 *                            ray_bvh_intersect record has hit_something (only 8 bytes come back) */
int gpu_benchmark_no_bvh(gpu_bvh *bvh, const Ray *rays, int64_t n, int64_t *intersections);
int gpu_benchmark_with_bvh(gpu_bvh *bvh, const Ray *rays, int64_t n, int64_t *intersections);

/* The path loop around the closest hit: main.c:356-366 + trace_ray renderer.c:21-76 for
 * image rows [y0,y1), one sample per pixel, as a wavefront: fused primary rays, then one ray-buffer launch per
 * bounce depth over the rays still alive, colour blend (Uint8)(base + 0.5 * reflected) folded back in the
 * recursion's order, sky gradient on a miss, black when max_depth (MAX_DEPTH = 5, constants.h:5) is exhausted.
 * Output: SDL_Color-layout RGBA8, 4 bytes per pixel.
 *   samples != NULL : [pixel][max_depth] unit vectors used as random_in_unit_sphere()'s return values
 *                     (sphere.c:19-24).  The reference draws them from the serial rand() stream; fed with
 *                     the reference's own samples the frame is bit-identical to the reference's.
 *   samples == NULL : counter-based generator keyed by (seed, pixel, depth); same rejection loop. */
int gpu_render_frame(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1, int max_depth,
                     uint64_t seed, const Vec3 *samples, uint8_t *rgba_out);
int gpu_render_frame_dev(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1, int max_depth,
                         uint64_t seed, const void *d_samples, void *d_rgba, void *stream);
/* Running mean of main.c:368-405: frames = number of frames including this one; frames == 1 restarts
 * (accum = colour/255, shown colour = the frame's own), otherwise accum += colour/255 and the shown colour
 * is (Uint8)fmin(accum / frames * 255.0f, 255.0f).  accum: 3 floats per pixel. */
int gpu_accumulate(float *accum, int frames, const uint8_t *rgba, int64_t n_pixels, uint8_t *avg_rgba);
int gpu_accumulate_dev(void *d_accum, int frames, const void *d_rgba, int64_t n_pixels, void *d_avg_rgba, void *stream);
/* Binary PPM (P6) of an RGBA8 frame, in place of SDL_RenderDrawPoint (main.c:376,402). 0 = ok. */
int sahbvh_write_ppm(const char *path, const uint8_t *rgba, int width, int height);

/* ---- one process, N devices (multi_gpu.cu).  The reference's drivers are single loops over pixels (main.c:356-366)
 * and rays (benchmark.c:230-245); rays are independent and the tree is read-only, so the loop shards: the BVH is
 * replicated on every device and each device takes a contiguous band of rows / range of rays.  No collective: the
 * bands' own stores assemble the frame, in the caller's host buffer or in one device's memory (peer stores). ---- */
typedef struct gpu_bvh_set gpu_bvh_set; /* N replicas of one tree, one worker thread per device; NOT thread-safe */
/* Device-to-device copies of src's flattened arrays (NVLink between peers; nothing is rebuilt).  devices == NULL
 * means ordinals 0 .. n_devices-1; an ordinal may repeat (several bands on one device).  src stays valid and
 * independent.  devices[0] is the set's ROOT: the device that owns an assembled frame (gpu_trace_primary_prims_multi_dev). */
int gpu_bvh_replicate(const gpu_bvh *src, const int *devices, int n_devices, gpu_bvh_set **out);
/* gpu_bvh_build on devices[0], then gpu_bvh_replicate: build_bvh_node's contract (spheres permuted in place). */
int gpu_bvh_build_replicated(Sphere *spheres, int n_total, int start, int end, int depth,
                             const int *devices, int n_devices, gpu_bvh_set **out);
void gpu_bvh_set_free(gpu_bvh_set *set);
int gpu_bvh_set_size(const gpu_bvh_set *set);
gpu_bvh *gpu_bvh_set_get(gpu_bvh_set *set, int i);   /* borrowed: the set keeps ownership */
/* the sharding rules: contiguous ray ranges; row bands of whole 16-row blocks (last band takes the remainder) */
void sahbvh_shard_range(int64_t n, int part, int n_parts, int64_t *lo, int64_t *hi);
void sahbvh_shard_rows(int height, int part, int n_parts, int *y0, int *y1);
/* uneven bands when part 0 (the root) also consumes the assembled frame: its band is root_share (0..1] of an equal
 * share, the others split the rest; root_share = 1 - (c/t)(N-1)/N for consumer cost c and equal-band time t */
void sahbvh_shard_rows_weighted(int height, int part, int n_parts, double root_share, int *y0, int *y1);
/* main.c:356-366 over N devices: the whole width x height frame into a HOST buffer (32-byte records / 4-byte ids). */
int gpu_trace_primary_multi(gpu_bvh_set *set, const Camera *cam, int width, int height, gpu_hit *out);
int gpu_trace_primary_prims_multi(gpu_bvh_set *set, const Camera *cam, int width, int height, int32_t *prims);
/* The id frame assembled in the ROOT device's memory: every device's trace kernel stores its band's rows into
 * d_prims (memory of devices[0]) through peer access, over NVLink, while it computes.  Asynchronous; root_stream (a
 * stream of the root device, NULL = its default stream) is made to wait for every band's completion event, so
 * whatever is enqueued on it afterwards (gpu_hits_from_prims_dev, gpu_count_prims_dev, a copy) sees the whole frame. */
int gpu_trace_primary_prims_multi_dev(gpu_bvh_set *set, const Camera *cam, int width, int height, void *d_prims,
                                      void *root_stream);
/* benchmark.c:230-245 over N devices: contiguous ray ranges, host buffers. */
int gpu_trace_closest_hit_multi(gpu_bvh_set *set, const Ray *rays, int64_t n, gpu_hit *out);
int gpu_trace_closest_hit_prims_multi(gpu_bvh_set *set, const Ray *rays, int64_t n, int32_t *prims);
int gpu_benchmark_with_bvh_multi(gpu_bvh_set *set, const Ray *rays, int64_t n, int64_t *intersections);
/* gpu_render_frame per row band; the counter-based draws are keyed by the pixel's index in the whole frame, so the
 * frame does not depend on the number of devices.  samples (may be NULL): [pixel of the frame][max_depth]. */
int gpu_render_frame_multi(gpu_bvh_set *set, const Camera *cam, int width, int height, int max_depth, uint64_t seed,
                           const Vec3 *samples, uint8_t *rgba_out);

/* Instrumented trace (counts work per ray; slower; same hits). */
int gpu_trace_closest_hit_stats(gpu_bvh *bvh, const void *d_rays, int64_t n, void *d_out, gpu_trace_stats *stats);
int gpu_trace_primary_stats(gpu_bvh *bvh, const Camera *cam, int width, int height, int y0, int y1,
                            void *d_out, gpu_trace_stats *stats);

/* Roofline denominators measured on the device in use (the L2 bandwidth and the FP32
 * issue rate are not in MEASURED_PEAKS.json).  Micro-kernels, ~50 ms in total. */
typedef struct sahbvh_peaks {
    double fp32_mix_inst_per_s;  /* thread-instructions/s of an FFMA+FMNMX mix (the box-test mix), all SMs */
    double ffma_inst_per_s;      /* thread-instructions/s of pure FFMA (x2 = FLOP/s) */
    double l2_read_gbs;          /* float4 read sweep over a 48 MiB buffer (L2 resident) */
    double hbm_read_gbs;         /* float4 read sweep over a 2 GiB buffer */
    double sm_count;
} sahbvh_peaks;
int sahbvh_measure_peaks(int device, sahbvh_peaks *out);

/* Number of kernel launches this library has issued in this process (bench.py's gpu_launches). */
uint64_t sahbvh_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* SAHBVH_GPU_H */
