"""Development harness (GPU): both packet kernels (16x8 tiles, 32x8 macro tiles) across scenes, and what
SAHBVH_PRIMARY_AUTO settles on."""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
import torch
pkg = g.load_pkg()
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def t(bvh, cam, W, H, recs, mode):
    pkg.set_primary_mode(mode)
    buf = torch.empty((W * H, 8), dtype=torch.float32, device="cuda") if recs else None
    ids = torch.empty((W * H,), dtype=torch.int32, device="cuda")
    fn = (lambda: bvh.trace_primary_dev(cam, W, H, 0, H, buf.data_ptr(), st)) if recs else (lambda: bvh.trace_primary_dev_prims(cam, W, H, 0, H, 0, ids.data_ptr(), st))
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(10):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms))
cam = pkg.default_camera()


def row(name, b, W, H):
    a = [t(b, cam, W, H, True, m) for m in (pkg.PRIMARY_PACKET, pkg.PRIMARY_MACRO)]
    c = [t(b, cam, W, H, False, m) for m in (pkg.PRIMARY_PACKET, pkg.PRIMARY_MACRO)]
    pkg.set_primary_mode(pkg.PRIMARY_AUTO)
    buf = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    for _ in range(12):
        b.trace_primary_dev(cam, W, H, 0, H, buf.data_ptr(), st)
        torch.cuda.synchronize()
    print(f"{name:32s} records: packet {a[0]:7.4f} macro {a[1]:7.4f} ms   ids: packet {c[0]:7.4f} macro {c[1]:7.4f} ms   auto -> {b.primary_kernel_in_use(W, H)}", flush=True)


pkg.srand(1); s = pkg.random_scene(20); b = pkg.GpuBvh.build(s, 20, 0, 20, 0)
row("20 spheres 800x600", b, 800, 600); b.close()
for n in (5000, 50000, 100000, 1000000, 10000000):
    pkg.srand(1); s = pkg.benchmark_scene(n); b = pkg.GpuBvh.build(s, n, 0, n, 0)
    row(f"{n} spheres 3840x2160", b, 3840, 2160)
    if n == 1000000: row(f"{n} spheres 3840x272 band", b, 3840, 272)
    if n == 10000000: row(f"{n} spheres 7680x4320", b, 7680, 4320)
    b.close()
