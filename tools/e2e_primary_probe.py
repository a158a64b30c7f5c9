"""Development harness (GPU): the headline frame end to end through the host-buffer C-ABI calls, every output format,
staged bands vs kernel stores straight into the pinned buffer (SAHBVH_HOST_DIRECT), plus the kernel alone per format."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import __graft_entry__ as g
    import torch
    pkg = g.load_pkg()
    n, W, H = 1000000, 3840, 2160
    pkg.srand(1)
    s = pkg.benchmark_scene(n)
    bvh = pkg.GpuBvh.build(s, n, 0, n, 0)
    cam = pkg.default_camera()
    recs = torch.empty((W * H, 8), dtype=torch.float32).pin_memory()
    prims = torch.empty((W * H,), dtype=torch.int32).pin_memory()
    tprim = torch.empty((W * H, 2), dtype=torch.float32).pin_memory()
    tag = f"direct={os.environ.get('SAHBVH_HOST_DIRECT', '0')} band={os.environ.get('SAHBVH_PRIMARY_BAND_PIXELS', 'default')}"

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    for name, fn in (("records 32B", lambda: bvh.trace_primary_ptr(cam, W, H, 0, H, recs.data_ptr())),
                     ("prims 4B", lambda: bvh.trace_primary_prims_ptr(cam, W, H, 0, H, prims.data_ptr())),
                     ("tprim 8B", lambda: bvh.trace_primary_tprim_ptr(cam, W, H, 0, H, tprim.data_ptr()))):
        dt = timed(fn)
        print(f"[{tag}] e2e {name}: {dt * 1e3:.3f} ms  {W * H / dt / 1e9:.2f} Grays/s", flush=True)
    ok = bool((prims.numpy() == recs.numpy().view(np.int32)[:, 1]).all())
    print(f"[{tag}] ids == records.prim: {ok}", flush=True)
    if os.environ.get("SAHBVH_HOST_DIRECT", "0") == "0" and "SAHBVH_PRIMARY_BAND_PIXELS" not in os.environ:
        d_rec = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
        d_ids = torch.empty((W * H, 2), dtype=torch.float32, device="cuda")
        flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for name, fn in (("records", lambda: bvh.trace_primary_dev(cam, W, H, 0, H, d_rec.data_ptr(), st)),
                         ("records+prims", lambda: bvh.trace_primary_dev_prims(cam, W, H, 0, H, d_rec.data_ptr(), d_ids.data_ptr(), st)),
                         ("prims only", lambda: bvh.trace_primary_dev_prims(cam, W, H, 0, H, 0, d_ids.data_ptr(), st))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ms = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            print(f"kernel {name}: {np.mean(ms):.4f} ms (min {np.min(ms):.4f})", flush=True)
        # the render-frame and counting outputs
        rgba = torch.empty((W * H,), dtype=torch.int32).pin_memory()
        dt = timed(lambda: bvh.render_frame_ptr(cam, W, H, 0, H, 5, 1, rgba.data_ptr()), 5)
        print(f"e2e gpu_render_frame RGBA8 host: {dt * 1e3:.3f} ms  {W * H / dt / 1e9:.3f} Gpaths/s", flush=True)
    sys.exit(0)

for env in ({}, {"SAHBVH_HOST_DIRECT": "1"}, {"SAHBVH_PRIMARY_BAND_PIXELS": str(1 << 19)}, {"SAHBVH_PRIMARY_BAND_PIXELS": str(1 << 21)},
            {"SAHBVH_PRIMARY_BAND_PIXELS": str(1 << 22)}):
    e = dict(os.environ)
    e.update(env)
    subprocess.run([sys.executable, __file__, "child"], env=e, check=False)
