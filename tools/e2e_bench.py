"""Development harness (GPU): gpu_trace_closest_hit end to end with pinned host buffers (config 4 rays)."""
import os, sys, time
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
import torch
pkg = g.load_pkg()
pkg.srand(1)
n = 1000000
s = pkg.benchmark_scene(n)
bvh = pkg.GpuBvh.build(s, n, 0, n, 0)
cam = pkg.default_camera()
W, H = 3840, 2160
hits = bvh.trace_primary(cam, W, H)
pkg.srand(2)
parts, have = [], 0
while have < 1 << 24:
    r = pkg.bounce_rays(hits); parts.append(r); have += len(r)
rays = np.concatenate(parts)[:1 << 24]
h_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 6)).pin_memory()
h_hits = torch.empty((len(rays), 8), dtype=torch.float32).pin_memory()
bvh.trace_closest_hit_ptr(h_rays.data_ptr(), len(rays), h_hits.data_ptr())
t0 = time.perf_counter()
for _ in range(3):
    bvh.trace_closest_hit_ptr(h_rays.data_ptr(), len(rays), h_hits.data_ptr())
dt = (time.perf_counter() - t0) / 3
print(f"chunk {os.environ.get('SAHBVH_HOST_CHUNK_RAYS','default')}: {dt*1e3:.2f} ms  {len(rays)/dt/1e9:.3f} Grays/s", flush=True)
