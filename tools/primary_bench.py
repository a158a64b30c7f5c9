"""Development harness (GPU): time the fused primary kernels on the headline frame (1 M spheres, 3840x2160)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

import torch
pkg = g.load_pkg()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pkg.set_keep_binary_tree(True)
pkg.srand(1)
s = pkg.benchmark_scene(n)
bvh = pkg.GpuBvh.build(s, n, 0, n, 0)
cam = pkg.default_camera()
W, H = 3840, 2160
buf = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
ref = None
for mode in (pkg.PRIMARY_PACKET, pkg.PRIMARY_MACRO, pkg.PRIMARY_PER_RAY):
    pkg.set_primary_mode(mode)
    for _ in range(3):
        bvh.trace_primary_dev(cam, W, H, 0, H, buf.data_ptr(), stream)
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bvh.trace_primary_dev(cam, W, H, 0, H, buf.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    h = buf.cpu().numpy()
    same = None if ref is None else bool((ref.view(np.uint32) == h.view(np.uint32)).all())
    if ref is None:
        ref = h
    print(f"mode {mode}: {np.mean(ms):.4f} ms (min {np.min(ms):.4f})  {W * H / np.mean(ms) / 1e6:.2f} Grays/s  same_as_first {same}", flush=True)

pkg.set_primary_mode(int(os.environ.get('PRIMARY_MODE', pkg.PRIMARY_MACRO)))
ids = torch.empty((W * H,), dtype=torch.int32, device="cuda")
for _ in range(3):
    bvh.trace_primary_dev_prims(cam, W, H, 0, H, 0, ids.data_ptr(), stream)
torch.cuda.synchronize()
ms = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bvh.trace_primary_dev_prims(cam, W, H, 0, H, 0, ids.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print(f"ids only: {np.mean(ms):.4f} ms (min {np.min(ms):.4f})  same ids {bool((ids.cpu().numpy() == ref.view(np.int32).reshape(-1, 8)[:, 1]).all())}", flush=True)
