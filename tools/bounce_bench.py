"""Development harness (GPU): time the ray-buffer kernels on config 4 (1 M spheres, 2^24 diffuse-bounce rays)
and check that every kernel variant returns the same hit records.  Not part of the product or the tests."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spheres", type=int, default=1000000)
    ap.add_argument("--rays", type=int, default=1 << 24)
    ap.add_argument("--modes", default="2,1,0")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--kind", default="bounce", choices=["bounce", "bench", "primary"])
    ap.add_argument("--stride", type=int, default=1, help="generate rays*stride rays and keep every stride-th")
    ap.add_argument("--sort", default="", choices=["", "morton", "morton_dir", "shuffle"], help="reorder the rays on the host before tracing")
    ap.add_argument("--l2-mb", default="", help="comma list of persisting-L2 window sizes (MiB) to time after the default run (sahbvh_set_l2_persist_bytes)")
    ap.add_argument("--tile", default="", help="WxH: take the primary hits in pixel-tile order (e.g. 8x4) instead of row-major")
    args = ap.parse_args()
    import torch
    pkg = g.load_pkg()
    pkg.srand(1)
    s = pkg.benchmark_scene(args.spheres)
    t0 = time.time()
    bvh = pkg.GpuBvh.build(s, args.spheres, 0, args.spheres, 0)
    print("build+upload s", round(time.time() - t0, 2), bvh.info(), flush=True)
    cam = pkg.default_camera()
    W, H = 3840, 2160
    buf = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    bvh.trace_primary_dev(cam, W, H, 0, H, buf.data_ptr(), stream)
    torch.cuda.synchronize()
    if args.kind == "bounce":
        hits = buf.cpu().numpy().view(pkg.HIT_DT).reshape(-1)
        if args.tile:
            tw, th = [int(v) for v in args.tile.split("x")]
            idx = np.arange(W * H).reshape(H // th, th, W // tw, tw).transpose(0, 2, 1, 3).reshape(-1)
            hits = np.ascontiguousarray(hits[idx])
        pkg.srand(2)
        parts, have = [], 0
        while have < args.rays * args.stride:
            r = pkg.bounce_rays(hits)
            parts.append(r)
            have += len(r)
        rays = np.ascontiguousarray(np.concatenate(parts)[:args.rays * args.stride:args.stride])
    elif args.kind == "bench":
        pkg.srand(3)
        rays = pkg.benchmark_rays(args.rays)
    else:
        rays = pkg.generate_primary_rays(cam, W, H)
    if args.sort:
        o = rays["o"].astype(np.float64)
        if args.sort == "shuffle":
            order = np.random.default_rng(0).permutation(len(rays))
        else:
            lo, hi = o.min(axis=0), o.max(axis=0)
            q = np.clip(((o - lo) / (hi - lo + 1e-9) * 1023).astype(np.uint64), 0, 1023)

            def spread(v):
                v = (v | (v << 16)) & 0x030000FF
                v = (v | (v << 8)) & 0x0300F00F
                v = (v | (v << 4)) & 0x030C30C3
                v = (v | (v << 2)) & 0x09249249
                return v
            key = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
            if args.sort == "morton_dir":
                d = rays["d"]
                octant = ((d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << 1) | ((d[:, 2] < 0).astype(np.uint64) << 2))
                key = (octant << 30) | key
            order = np.argsort(key, kind="stable")
        rays = np.ascontiguousarray(rays[order])
    n = len(rays)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 6)).cuda()
    ref = None
    for mode in [int(m) for m in args.modes.split(",")]:
        pkg.set_ray_mode(mode)
        d_hits = torch.zeros((n, 8), dtype=torch.float32, device="cuda")
        st = bvh.trace_closest_hit_stats(d_rays.data_ptr(), n, d_hits.data_ptr())
        for _ in range(2):
            bvh.trace_closest_hit_dev(d_rays.data_ptr(), n, d_hits.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            bvh.trace_closest_hit_dev(d_rays.data_ptr(), n, d_hits.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        h = d_hits.cpu().numpy()
        same = None
        if ref is None:
            ref = h
        else:
            same = bool((ref.view(np.uint32) == h.view(np.uint32)).all())
        print(f"mode {mode}: {ms:8.3f} ms  {n / ms / 1e6:7.3f} Grays/s  nodes/ray {st['inner_visits'] / n:6.1f} "
              f"spheres/ray {st['sphere_tests'] / n:5.2f} max_stack {st['max_stack']} max_visits {st['max_visits']} same_as_first {same} sha {__import__('hashlib').sha256(h.tobytes()).hexdigest()[:12]}", flush=True)
        for mb in [int(v) for v in args.l2_mb.split(",") if v]:
            pkg.set_l2_persist_bytes(mb << 20)
            for _ in range(2):
                bvh.trace_closest_hit_dev(d_rays.data_ptr(), n, d_hits.data_ptr(), stream)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                bvh.trace_closest_hit_dev(d_rays.data_ptr(), n, d_hits.data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            same_l2 = bool((d_hits.cpu().numpy().view(np.uint32) == ref.view(np.uint32)).all())
            print(f"mode {mode} with a {mb} MiB persisting-L2 window over the top of the tree: {e0.elapsed_time(e1) / args.steps:8.3f} ms  same {same_l2}", flush=True)
        pkg.set_l2_persist_bytes(0)


if __name__ == "__main__":
    main()
