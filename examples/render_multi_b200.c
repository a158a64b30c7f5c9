/*
 * examples/render_multi_b200.c -- the reference's render driver (main.c:203-228 scene + camera + BVH,
 * main.c:356-366 per-pixel loop, main.c:368-405 accumulation) on N GPUs of one box, in plain C on the C-ABI of
 * libsahbvh_b200.so.  The per-pixel loop
 *
 *     for y, x:  ray = get_camera_ray(&camera, u, -v);  color = trace_ray(ray, spheres, n, 0, bvh);
 *
 * becomes one call per frame; the BVH is built once (device builder, same tree and permutation as build_bvh_node),
 * replicated on every device, and every device takes a band of rows.
 *
 *   gcc -O2 -Iinclude examples/render_multi_b200.c -Lcs201_sah-bvh_ray_tracer_b200 -lsahbvh_b200 \
 *       -Wl,-rpath,$PWD/cs201_sah-bvh_ray_tracer_b200 -lm -o render_multi_b200
 *   ./render_multi_b200 [bands] [n_spheres] [width] [height] [spp] [out.ppm]      (default: all devices, 20, 800, 600, 4)
 *
 * `bands` may exceed the number of devices (ordinals repeat): the frame does not depend on it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sahbvh_gpu.h"
#include "sahbvh_ref_api.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define TRY(call)                                                            \
    do {                                                                     \
        if ((call) != SAHBVH_OK) {                                           \
            fprintf(stderr, "%s: %s\n", #call, sahbvh_last_error());         \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(int argc, char **argv)
{
    const int ndev = sahbvh_device_count();
    if (ndev < 1) {
        fprintf(stderr, "no usable CUDA device (this path has no CPU fallback)\n");
        return 1;
    }
    int bands = argc > 1 ? atoi(argv[1]) : ndev;
    const int n = argc > 2 ? atoi(argv[2]) : 20;
    const int width = argc > 3 ? atoi(argv[3]) : 800, height = argc > 4 ? atoi(argv[4]) : 600;
    const int spp = argc > 5 ? atoi(argv[5]) : 4;
    const char *out_path = argc > 6 ? argv[6] : "frame.ppm";
    if (bands < 1 || bands > 64 || n < 1 || width < 1 || height < 1 || spp < 1) {
        fprintf(stderr, "bad arguments\n");
        return 1;
    }
    int devices[64];
    for (int i = 0; i < bands; i++) devices[i] = i % ndev;

    /* main.c:203-221: camera literal, srand, scene (benchmark.c's distribution for the large scenes) */
    Camera camera;
    sahbvh_default_camera(&camera);
    srand(1);
    Sphere *spheres = malloc(((size_t)n + 1) * sizeof(Sphere));
    if (n <= 5000) sahbvh_fill_random_scene(n, spheres);
    else sahbvh_fill_benchmark_scene(n, spheres);

    /* main.c:225: bvh = build_bvh_node(spheres, 0, n, 0) -- here straight into the flattened device form, N replicas */
    double t0 = now_s();
    gpu_bvh_set *set = NULL;
    TRY(gpu_bvh_build_replicated(spheres, n, 0, n, 0, devices, bands, &set));
    printf("BVH over %d spheres built and replicated on %d band(s) / %d device(s) in %.3f s\n", n, bands, ndev, now_s() - t0);

    const size_t npx = (size_t)width * (size_t)height;
    int32_t *prims = NULL;
    uint8_t *rgba = NULL, *shown = NULL;
    TRY(sahbvh_host_alloc(npx * sizeof(int32_t), (void **)&prims));   /* page-locked: the copies run at PCIe speed */
    TRY(sahbvh_host_alloc(npx * 4, (void **)&rgba));
    shown = malloc(npx * 4);
    float *accum = calloc(npx * 3, sizeof(float));

    /* closest hit of every primary ray: the hit OBJECT per pixel (hit.h:13), 4 bytes */
    TRY(gpu_trace_primary_prims_multi(set, &camera, width, height, prims));   /* warm-up: staging buffers, module load */
    t0 = now_s();
    TRY(gpu_trace_primary_prims_multi(set, &camera, width, height, prims));
    double dt = now_s() - t0;
    size_t hits = 0;
    for (size_t i = 0; i < npx; i++) hits += prims[i] >= 0;
    printf("primary frame %dx%d: %zu of %zu rays hit, %.3f ms, %.2f Mrays/s end to end\n", width, height, hits, npx, dt * 1e3,
           (double)npx / dt / 1e6);
    /* the other HitRecord fields of any pixel, on demand (bit-identical to ray_bvh_intersect's record) */
    for (size_t i = npx / 2; i < npx; i++) {
        if (prims[i] < 0) continue;
        const int x = (int)(i % (size_t)width), y = (int)(i / (size_t)width);
        HitRecord rec = sahbvh_hit_from_prim(sahbvh_pixel_ray(&camera, width, height, x, y), spheres, prims[i]);
        printf("pixel (%d,%d): sphere %d  t=%.6f  point=(%.4f %.4f %.4f)  normal=(%.4f %.4f %.4f)\n", x, y, prims[i], rec.t,
               rec.point.x, rec.point.y, rec.point.z, rec.normal.x, rec.normal.y, rec.normal.z);
        break;
    }

    /* main.c:356-405: spp frames of trace_ray folded by the running mean */
    t0 = now_s();
    for (int f = 1; f <= spp; f++) {
        TRY(gpu_render_frame_multi(set, &camera, width, height, 5 /* MAX_DEPTH, constants.h:5 */, 1000u + (unsigned)f, NULL, rgba));
        TRY(gpu_accumulate(accum, f, rgba, (int64_t)npx, shown));
    }
    dt = now_s() - t0;
    printf("%d accumulated frame(s) of the path loop: %.3f ms per frame\n", spp, dt * 1e3 / spp);
    TRY(sahbvh_write_ppm(out_path, shown, width, height));
    unsigned long long sum = 0;
    for (size_t i = 0; i < npx * 4; i++) sum = sum * 1099511628211ull + shown[i];
    printf("frame checksum %016llx -> %s\n", sum, out_path);

    gpu_bvh_set_free(set);
    sahbvh_host_free(prims);
    sahbvh_host_free(rgba);
    free(shown);
    free(accum);
    free(spheres);
    return 0;
}
