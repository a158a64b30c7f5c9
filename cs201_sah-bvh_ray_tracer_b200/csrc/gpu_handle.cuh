/* gpu_handle.cuh -- the device handle and the launch helpers shared by the CUDA translation units. */
#ifndef SAHBVH_GPU_HANDLE_CUH
#define SAHBVH_GPU_HANDLE_CUH

#include <cuda_runtime.h>

#include <atomic>

#include "sahbvh_internal.h"
#include "traverse_core.cuh"

/* SAHBVH_PRIMARY_AUTO: per frame shape, which of the two packet kernels is faster on this handle (sahbvh_kernels.cu) */
#define SAHBVH_TUNE_SLOTS 8
struct PrimaryTune {
    int width, rows;          /* key */
    int decided;              /* -1 while sampling, else SAHBVH_PRIMARY_PACKET / SAHBVH_PRIMARY_MACRO */
    int launches;
    cudaEvent_t ev[2][2][2];  /* [kernel][sample][start, stop] */
};

struct gpu_bvh {
    int device;
    float *d_nodes;
    float *d_spheres4;
    uint32_t *d_colors;
    uint32_t *d_wide;   /* 4-wide quantised nodes (wide.c), 64 B each */
    float *d_special;   /* special-leaf records, 32 B each */
    float *d_wide_full; /* full-precision twins of the wide nodes, 128 B each (packet kernel) */
    int n_wide, n_special;
    int wide_blocks_per_sm;
    unsigned long long *d_counters; /* [0..7] work counters, [8..71] per-launch ray cursors of the persistent kernel */
    unsigned next_counter;
    int coop_blocks_per_sm;
    gpu_bvh_info info;
    cudaStream_t streams[2];
    /* staging for the host-buffer entry points (grown on demand) */
    void *d_in[2];
    void *d_out[2];
    size_t in_cap, out_cap;
    int sm_count;
    /* path-loop work buffers (render.cu), one allocation grown on demand */
    void *d_render;
    size_t render_cap;
    PrimaryTune tune[SAHBVH_TUNE_SLOTS];
    int n_tune;
    unsigned tune_evict;
    size_t l2_limit_applied; /* persisting-L2 set-aside this handle asked the device for (sahbvh_set_l2_persist_bytes) */
};

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(false)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        if (dev != prev && cudaSetDevice(dev) != cudaSuccess) return;
        ok = true;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};


extern std::atomic<unsigned long long> g_launches;

/* text of a runtime error that is being reported through sahbvh_last_error(); also clears the runtime's
 * (non-sticky) last-error slot so that the next launch check does not report the same failure again */
static inline const char *cu_errstr(cudaError_t e)
{
    (void)cudaGetLastError();
    return cudaGetErrorString(e);
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            sahbvh_set_error("%s failed: %s (%s:%d)", #expr, cu_errstr(_e), __FILE__, __LINE__); \
            return SAHBVH_E_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

SahScene sahbvh_make_scene(const gpu_bvh_info &info);
void sahbvh_make_camera(const Camera *cam, int width, int height, SahCamera *c);
/* d_out != NULL: 32-byte records; d_out == NULL: 4-byte primitive ids into d_prims (4-wide kernel only) */
int sahbvh_launch_rays(gpu_bvh *b, const void *d_rays, long long n, void *d_out, cudaStream_t st, bool count, void *d_prims = nullptr);
int sahbvh_launch_primary(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, void *d_out,
                          cudaStream_t st, bool count, void *d_ids, int ids_mode);
/* what the primary kernels write to d_ids */
#define SAH_IDS_NONE 0
#define SAH_IDS_PRIM 1   /* int32 primitive id per pixel (-1 = miss) */
#define SAH_IDS_TPRIM 2  /* {float t, int32 prim} per pixel */
/* host-buffer frames: what goes back over PCIe per pixel */
#define HOSTFMT_RECORDS 0 /* gpu_hit, 32 B */
#define HOSTFMT_PRIMS 1   /* int32 primitive id, 4 B */
#define HOSTFMT_TPRIM 2   /* {float t, int32 prim}, 8 B */
int sahbvh_primary_host_enqueue(gpu_bvh *b, const Camera *cam, int width, int height, int y0, int y1, void *out, int fmt);
int sahbvh_primary_host_wait(gpu_bvh *b);
/* staging pairs d_in[0..1] / d_out[0..1] of the host-buffer entry points, grown on demand */
int sahbvh_ensure_staging(gpu_bvh *b, size_t in_bytes, size_t out_bytes);
extern "C" int sahbvh_set_packet_dfs_from(int v); /* diagnostic, see sahbvh_kernels.cu */
/* the 64-byte binary tree is only read by the selectable binary-tree kernels: it is built from the 4-wide twin on
 * first use instead of being kept in every handle */
int sahbvh_need_binary_tree(gpu_bvh *b);

#endif
