/*
 * examples/benchmark_b200.c -- benchmark.c's run_benchmark_with_plotting() (benchmark.c:283-331) with its two timed
 * loops replaced by the batch C-ABI of libsahbvh_b200.so.  Everything else follows the reference line by line:
 * the scene recipe (benchmark.c:306-314), build_bvh_node(spheres, 0, n-1, 20) (benchmark.c:317), the ray recipe
 * (benchmark.c:232-240, drawn from the same rand() stream, first for the no-BVH loop, then for the BVH loop), the
 * printed lines and the "%d %f %f" rows of benchmark_data.txt (benchmark.c:160-170).
 *
 *   gcc -O2 -Iinclude examples/benchmark_b200.c -Lcs201_sah-bvh_ray_tracer_b200 -lsahbvh_b200 \
 *       -Wl,-rpath,$PWD/cs201_sah-bvh_ray_tracer_b200 -lm -o benchmark_b200
 *   ./benchmark_b200 [seed] [first_n] [last_n] [step]        (default: 1 5000 50000 5000; the reference seeds with time(NULL))
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "sahbvh_gpu.h"
#include "sahbvh_ref_api.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void fill_rays(Ray *rays, int num_rays)
{ /* benchmark.c:232-240 */
    for (int i = 0; i < num_rays; i++) {
        Vec3 dir = {(float)rand() / RAND_MAX * 2 - 1, (float)rand() / RAND_MAX * 2 - 1, (float)rand() / RAND_MAX * 2 - 1};
        dir = vec3_normalize(dir);
        Ray ray = {{0, 0, 0}, dir};
        rays[i] = ray;
    }
}

int main(int argc, char **argv)
{
    unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
    int first = argc > 2 ? atoi(argv[2]) : 5000, last = argc > 3 ? atoi(argv[3]) : 50000, step = argc > 4 ? atoi(argv[4]) : 5000;
    const int num_rays = 10000;
    const float world_size = 1000.0f;
    remove("benchmark_data.txt");
    srand(seed);
    Ray *rays = malloc((size_t)num_rays * sizeof(Ray));
    for (int num_spheres = first; num_spheres <= last; num_spheres += step) {
        printf("Testing with %d spheres:\n", num_spheres);
        Sphere *spheres = malloc((size_t)num_spheres * sizeof(Sphere));
        for (int j = 0; j < num_spheres; j++) {
            Vec3 center = {(float)rand() / RAND_MAX * world_size - world_size / 2,
                           (float)rand() / RAND_MAX * world_size - world_size / 2,
                           (float)rand() / RAND_MAX * world_size - world_size / 2};
            spheres[j] = create_benchmark_sphere(center);
        }
        BVHNode *root = build_bvh_node(spheres, 0, num_spheres - 1, 20);
        gpu_bvh *g = NULL;
        if (gpu_bvh_upload(spheres, num_spheres, root, -1, &g) != SAHBVH_OK) {
            fprintf(stderr, "gpu_bvh_upload: %s\n", sahbvh_last_error());
            return 1;
        }
        int64_t hits_no_bvh = 0, hits_with_bvh = 0;
        double t0 = now_s();                                   /* benchmark_no_bvh, benchmark.c:172-223 */
        fill_rays(rays, num_rays);
        if (gpu_benchmark_no_bvh(g, rays, num_rays, &hits_no_bvh) != SAHBVH_OK) {
            fprintf(stderr, "gpu_benchmark_no_bvh: %s\n", sahbvh_last_error());
            return 1;
        }
        double time_no_bvh = now_s() - t0;
        printf("No BVH:\nTime: %f seconds\nIntersection tests: %lld\nIntersections found: %lld\n\n", time_no_bvh,
               (long long)num_rays * num_spheres, (long long)hits_no_bvh);
        t0 = now_s();                                          /* benchmark_with_bvh, benchmark.c:225-254 */
        fill_rays(rays, num_rays);
        if (gpu_benchmark_with_bvh(g, rays, num_rays, &hits_with_bvh) != SAHBVH_OK) {
            fprintf(stderr, "gpu_benchmark_with_bvh: %s\n", sahbvh_last_error());
            return 1;
        }
        double time_with_bvh = now_s() - t0;
        printf("With BVH:\nTime: %f seconds\nIntersections found: %lld\n\n", time_with_bvh, (long long)hits_with_bvh);
        FILE *file = fopen("benchmark_data.txt", "a");         /* save_benchmark_data, benchmark.c:160-170 */
        if (file) {
            fprintf(file, "%d %f %f\n", num_spheres, time_no_bvh, time_with_bvh);
            fclose(file);
        }
        gpu_bvh_free(g);
        free_bvh(root);
        free(spheres);
        printf("----------------------------------------\n");
    }
    free(rays);
    return 0;
}
