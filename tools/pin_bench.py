"""Development harness (GPU): gpu_trace_primary into a malloc'ed (numpy) frame buffer, as is and after sahbvh_host_pin."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_pkg()
n = 1000000
pkg.srand(1)
s = pkg.benchmark_scene(n)
bvh = pkg.GpuBvh.build(s, n, 0, n, 0)
cam = pkg.default_camera()
W, H = 3840, 2160
frame = np.zeros(W * H, pkg.HIT_DT)
for label in ("pageable", "pinned with sahbvh_host_pin"):
    if label.startswith("pinned"):
        t0 = time.perf_counter()
        pkg.host_pin(frame)
        print(f"sahbvh_host_pin of {frame.nbytes >> 20} MiB: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
    for _ in range(2):
        bvh.trace_primary(cam, W, H, out=frame)
    t = []
    for _ in range(8):
        t0 = time.perf_counter()
        bvh.trace_primary(cam, W, H, out=frame)
        t.append(time.perf_counter() - t0)
    print(f"{label}: {np.mean(t) * 1e3:.2f} ms per 4K frame  {W * H / np.mean(t) / 1e9:.2f} Grays/s", flush=True)
pkg.host_unpin(frame)
