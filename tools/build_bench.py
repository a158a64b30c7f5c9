"""Development harness (GPU): host vs device SAH builder wall-clock (sahbvh_build_tree) and gpu_bvh_build."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_pkg()
for n in [int(a) for a in (sys.argv[1:] or ["100000", "1000000"])]:
    pkg.srand(1)
    s0 = pkg.benchmark_scene(n)
    row = {"n": n}
    for name, mode in (("device", pkg.BUILD_DEVICE), ("host", pkg.BUILD_HOST)):
        if name == "host" and n > 2000000:
            continue
        best = 1e9
        for rep in range(3 if name == "device" else 1):
            s = s0.copy()
            t0 = time.perf_counter()
            tree = pkg.build_tree(s, 0, n, 0, mode=mode)
            best = min(best, time.perf_counter() - t0)
        row[name + "_tree_s"] = round(best, 4)
        row["nodes"] = len(tree)
        pkg.set_build_mode(mode)
        s = s0.copy()
        t0 = time.perf_counter()
        bvh = pkg.GpuBvh.build(s, n, 0, n, 0)
        row[name + "_gpu_bvh_build_s"] = round(time.perf_counter() - t0, 4)
        bvh.close()
    print(row, flush=True)
