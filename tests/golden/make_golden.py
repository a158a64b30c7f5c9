#!/usr/bin/env python
"""Generate the golden fixtures from the REFERENCE ITSELF (oracle/_ref/libref*.so = the
unmodified /root/reference sources compiled by oracle/Makefile).

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Writes tests/golden/known_answers.json (statistics + SHA-256 of every intermediate: scene
bytes, exported tree, rays, hit records) and tests/golden/vectors.npz (small input/output
vectors).  The reference has no tests or golden vectors of its own, so these
outputs of the compiled reference are what pins the oracle and, through it, the CUDA path.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle.oracle_py import RefLib  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    R = RefLib()
    R4 = RefLib("3840x2160")
    answers = {"generator": "tests/golden/make_golden.py", "reference_flags": "-O2 -ffp-contract=off", "cases": {}}
    vec = {}

    def record(name, scene, n, build_args, rays, stats, hits, keep=None):
        case = {
            "n_spheres": n, "build": list(build_args), "tree": stats,
            "n_rays": int(len(rays)), "hits": int((hits["prim"] >= 0).sum()),
            "sha_scene_post_build": sha(scene[:n]), "sha_rays": sha(rays), "sha_hits": sha(hits),
        }
        answers["cases"][name] = case
        if keep is not None:
            idx = keep
            vec[f"{name}__idx"] = idx.astype(np.int64)
            vec[f"{name}__rays"] = rays[idx]
            vec[f"{name}__hits"] = hits[idx]
            if n <= 5000:
                vec[f"{name}__spheres"] = scene[:n].copy()
        print(name, case["tree"], "hits", case["hits"])

    # --- benchmark.c as shipped: srand, scene, build(0,n-1,20), rays from the SAME stream
    for n, nrays in ((5000, 10000), (50000, 10000)):
        R.srand(1)
        s = R.scene_benchmark(n)
        root, _ = R.build(s, 0, n - 1, 20)
        st = R.tree_stats(root, 20)
        answers.setdefault("tree_sha", {})[f"bench_{n}"] = sha(R.export_tree(root, s, st["nodes"]))
        rays = R.rays_benchmark(nrays)
        hits, _ = R.trace_bvh(root, s, rays)
        keep = np.arange(nrays) if n == 5000 else np.nonzero(hits["prim"] >= 0)[0]
        record(f"bench_{n}_asshipped", s, n, (0, n - 1, 20), rays, st, hits, keep)
        R.free(root)

    # --- main.c-style builds (0,n,0) on the benchmark distribution
    for n, nrays in ((100000, 100000), (1000000, 100000)):
        R.srand(1)
        s = R.scene_benchmark(n)
        root, sec = R.build(s, 0, n, 0)
        st = R.tree_stats(root, 0)
        answers.setdefault("tree_sha", {})[f"bench_{n}_full"] = sha(R.export_tree(root, s, st["nodes"]))
        rays = R.rays_benchmark(nrays)
        hits, _ = R.trace_bvh(root, s, rays)
        keep = np.arange(0, nrays, 25)
        record(f"bench_{n}_full", s, n, (0, n, 0), rays, st, hits, keep)
        if n == 100000:
            # 4K primary frame sample (every 997th pixel) through the 3840x2160 build of ray.c
            cam = R4.default_camera()
            pr = R4.primary_rays(cam)
            idx = np.arange(0, len(pr), 997)
            ph, _ = R.trace_bvh(root, s, pr[idx])
            answers["cases"]["bench_100000_primary4k_sample"] = {
                "n_rays": int(len(idx)), "stride": 997, "hits": int((ph["prim"] >= 0).sum()),
                "sha_rays_full_frame": sha(pr), "sha_hits": sha(ph)}
            vec["bench_100000_primary4k_sample__hits"] = ph
            # bounce rays off the 800x600 frame
            cam8 = R.default_camera()
            p8 = R.primary_rays(cam8)
            h8, _ = R.trace_bvh(root, s, p8)
            R.srand(5)
            br = R.bounce_rays(h8)
            bh, _ = R.trace_bvh(root, s, br)
            answers["cases"]["bench_100000_bounce"] = {
                "n_rays": int(len(br)), "hits": int((bh["prim"] >= 0).sum()),
                "sha_primary_hits_800x600": sha(h8), "sha_rays": sha(br), "sha_hits": sha(bh)}
            vec["bench_100000_bounce__rays"] = br[:4096]
            vec["bench_100000_bounce__hits"] = bh[:4096]
        print(f"  ref build {sec:.1f}s")
        R.free(root)

    # --- default renderer scene, full 800x600 frame; degenerate seeds 649/1164 (Q2/Q3)
    for seed in (1, 7, 649, 1164):
        R.srand(seed)
        s = R.scene_random(20)
        root, _ = R.build(s, 0, 20, 0)
        st = R.tree_stats(root, 0)
        answers.setdefault("tree_sha", {})[f"default_{seed}"] = sha(R.export_tree(root, s, st["nodes"]))
        cam = R.default_camera()
        rays = R.primary_rays(cam)
        hits, _ = R.trace_bvh(root, s, rays)
        brute, _ = R.trace_brute(s, 20, rays)
        keep = np.arange(0, len(rays), 97)
        record(f"default_seed{seed}", s, 20, (0, 20, 0), rays, st, hits, keep)
        answers["cases"][f"default_seed{seed}"]["brute_object_mismatches"] = int((brute["prim"] != hits["prim"]).sum())
        answers["cases"][f"default_seed{seed}"]["sha_brute_hits"] = sha(brute)
        R.free(root)

    # --- 5000 varying-radius spheres: depth-40 multi-sphere leaves (Q2)
    R.srand(3)
    s = R.scene_random(5000)
    root, _ = R.build(s, 0, 5000, 0)
    st = R.tree_stats(root, 0)
    cam = R.default_camera()
    rays = R.primary_rays(cam)[::7].copy()
    hits, _ = R.trace_bvh(root, s, rays)
    record("random5000_seed3", s, 5000, (0, 5000, 0), rays, st, hits, np.arange(0, len(rays), 16))
    R.free(root)

    # --- moved camera (camera.c:10-18)
    cam = R.default_camera()
    cam["yaw"] = np.float32(-2.6)
    cam["pitch"] = np.float32(0.15)
    R.camera_update(cam)
    answers["camera_update"] = {"yaw": -2.6, "pitch": 0.15, "sha_camera": sha(cam), "sha_rays_800x600": sha(R.primary_rays(cam))}
    vec["camera_update__cam"] = cam

    # --- the path loop: main.c:356-366 over trace_ray (renderer.c:21-76), MAX_DEPTH 5, 800x600, and the
    #     accumulation of main.c:368-405 over three frames
    R.srand(1)
    s = R.scene_random(20)
    root, _ = R.build(s, 0, 20, 0)
    cam = R.default_camera()
    frames = []
    for rep in range(3):
        R.srand(100 + rep)
        frames.append(R.render(cam, s, 20, root if rep != 1 else None))
    acc = np.zeros((800 * 600, 3), np.float32)
    nfr = np.zeros(1, np.int32)
    shown = [R.accumulate(acc, nfr, f, first=(k == 0)) for k, f in enumerate(frames)]
    answers["render_default_seed1"] = {
        "scene": "srand(1), 20 x create_random_sphere, build(0,20,0)", "frames": "srand(100+k) before frame k; frame 1 brute force (bvh = NULL)",
        "sha_frames": [sha(f) for f in frames], "sha_shown": [sha(x) for x in shown], "sha_accum": sha(acc),
        "distinct_colours_frame0": int(len(np.unique(frames[0], axis=0)))}
    R.free(root)

    answers["layout"] = R.layout[:20]
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(answers, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "vectors.npz"), **vec)
    print("wrote", os.path.join(HERE, "known_answers.json"), "and vectors.npz",
          os.path.getsize(os.path.join(HERE, "vectors.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
