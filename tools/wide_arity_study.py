"""Development study (CPU only): how many wide-node visits would a K-wide greedy collapse of the reference's binary
tree cost per diffuse-bounce ray?  visits_K(ray) = number of K-wide node roots whose binary box the ray pierces
(child boxes lie inside their parents', so a node is reached iff its own box is pierced; best-t culling is ignored,
which the K = 4 row checks against the kernel's own counter: 103-106 visits per ray at 1 M spheres).
Usage: python tools/wide_arity_study.py [n_spheres] [n_rays]"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import HostEmul  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402


def collapse_roots(nodes, K):
    """flags[i] = binary node i is the root of a K-wide node (greedy: open the inner slot with the largest area)."""
    lo, hi = nodes["lo"].astype(np.float64), nodes["hi"].astype(np.float64)
    d = hi - lo
    area = d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]
    left, right = nodes["left"], nodes["right"]
    inner = left >= 0
    flags = np.zeros(len(nodes), bool)
    fill = []
    todo = [0]
    while todo:
        r = todo.pop()
        flags[r] = True
        slots = [int(left[r]), int(right[r])]
        while len(slots) < K:
            cand = [(area[s], k) for k, s in enumerate(slots) if inner[s]]
            if not cand:
                break
            _, k = max(cand)
            s = slots.pop(k)
            slots[k:k] = [int(left[s]), int(right[s])]
        fill.append(len(slots))
        todo.extend(s for s in slots if inner[s])
    return flags, float(np.mean(fill))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    o, em = Oracle(), HostEmul("wide")
    o.srand(1)
    s = o.scene_benchmark(n)
    nodes = em.build_tree(s, 0, n, 0)
    h = em.build(s.copy(), n, 0, n, 0)
    cam = o.default_camera()
    hits, _ = em.trace(h, o.primary_rays(cam, 640, 360), threads=8)
    o.srand(2)
    rays = o.bounce_rays(hits)
    rays = rays[np.random.default_rng(0).choice(len(rays), m, replace=False)]
    _, st = em.trace(h, rays, threads=8, drain_every=1)
    print(f"{n} spheres, {len(nodes)} binary nodes, {m} bounce rays; 4-wide kernel core counts {st['inner_visits'] / m:.1f} visits per ray")
    lo, hi = nodes["lo"].astype(np.float64), nodes["hi"].astype(np.float64)
    inner = nodes["left"] >= 0
    pierced = np.zeros((m, len(nodes)), bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i, r in enumerate(rays):
            oo, dd = r["o"].astype(np.float64), r["d"].astype(np.float64)
            t1, t2 = (lo - oo) / dd, (hi - oo) / dd
            tn, tf = np.minimum(t1, t2).max(axis=1), np.maximum(t1, t2).min(axis=1)
            pierced[i] = (tf >= np.maximum(tn, 0.0)) & inner
    print(f"binary inner nodes pierced per ray: {pierced.sum(axis=1).mean():.1f}")
    for K in (2, 4, 6, 8, 12, 16):
        flags, fill = collapse_roots(nodes, K)
        v = pierced[:, flags].sum(axis=1).mean()
        print(f"K = {K:2d}: {int(flags.sum()):8d} wide nodes, {fill:5.2f} children per node, {v:6.1f} visits per ray, "
              f"{v * fill:7.1f} child box tests per ray")


if __name__ == "__main__":
    main()
