#!/usr/bin/env python
"""Golden fixtures for the 10 M-sphere scene (BASELINE config 5; reference sweep list benchmark.c:296),
generated from the REFERENCE ITSELF (oracle/_ref/libref*.so = the unmodified /root/reference sources).

Run in the build container (where /root/reference exists); the reference's own build_bvh_node takes
about three minutes for 10 M spheres on one core:

    python tests/golden/make_golden_10m.py

Writes tests/golden/vectors_10m.npz (rays + the reference's hit records) and
tests/golden/known_answers_10m.json (tree statistics, SHA-256 of the post-build sphere array, of the
exported tree and of every ray / hit array):

  * bench rays   : 20 000 rays of benchmark.c:232-240 drawn from the rand() stream right after the scene
  * primary 8K   : 20 000 pixels of the 7680x4320 frame (main.c:356-365 + ray.c:17-32 compiled at that size)
  * bounce       : 20 000 diffuse-bounce rays (renderer.c:51-54) off the hits of those pixels, srand(5)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle.oracle_py import RefLib  # noqa: E402

N = 10_000_000
N_RAYS = 20_000


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    R = RefLib()
    R8 = RefLib("7680x4320")
    t0 = time.time()
    R.srand(1)
    s = R.scene_benchmark(N)
    sha_scene_in = sha(s[:N])
    bench_rays = R.rays_benchmark(N_RAYS)          # the same rand() stream, right after the scene (benchmark.c order)
    root, sec = R.build(s, 0, N, 0)
    print(f"reference build_bvh_node(0,{N},0): {sec:.1f} s", flush=True)
    st = R.tree_stats(root, 0)
    tree = R.export_tree(root, s, st["nodes"])
    out = {"generator": "tests/golden/make_golden_10m.py", "reference_flags": "-O2 -ffp-contract=off",
           "n_spheres": N, "build": [0, N, 0], "tree": st, "reference_build_seconds": sec,
           "sha_scene_pre_build": sha_scene_in, "sha_scene_post_build": sha(s[:N]), "sha_tree": sha(tree), "cases": {}}
    del tree
    vec = {}

    def record(name, rays, hits, extra=None):
        out["cases"][name] = {"n_rays": int(len(rays)), "hits": int((hits["prim"] >= 0).sum()),
                              "sha_rays": sha(rays), "sha_hits": sha(hits), **(extra or {})}
        vec[f"{name}__rays"] = rays
        vec[f"{name}__hits"] = hits
        print(name, out["cases"][name]["hits"], "hits of", len(rays), flush=True)

    hb, _ = R.trace_bvh(root, s, bench_rays)
    record("bench_rays", bench_rays, hb)

    cam = R8.default_camera()
    pr = R8.primary_rays(cam)
    rng = np.random.RandomState(10)
    idx = np.sort(rng.choice(len(pr), N_RAYS, replace=False)).astype(np.int64)
    hp, _ = R.trace_bvh(root, s, pr[idx])
    vec["primary8k__idx"] = idx
    record("primary8k", pr[idx].copy(), hp, {"width": 7680, "height": 4320, "sha_rays_full_frame": sha(pr)})

    R.srand(5)
    br = R.bounce_rays(hp)
    while len(br) < N_RAYS:                        # hit rate < 1: draw again off the same hits (fresh directions)
        br = np.concatenate([br, R.bounce_rays(hp)])
    br = br[:N_RAYS].copy()
    hbnc, _ = R.trace_bvh(root, s, br)
    record("bounce", br, hbnc)
    R.free(root)

    out["seconds_total"] = time.time() - t0
    with open(os.path.join(HERE, "known_answers_10m.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "vectors_10m.npz"), **vec)
    print("wrote known_answers_10m.json and vectors_10m.npz", os.path.getsize(os.path.join(HERE, "vectors_10m.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
