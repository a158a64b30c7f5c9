"""The one implication the whole design rests on, tested directly (CPU, host build of traverse_core.cuh):

    the reference's slab test (hit.c:49-82) accepts box B for ray R
        ==>  every interior test the kernels apply to B for R accepts it
             (per-ray FMA test of the binary tree, byte-quantised 4-wide node test, per-tile frustum test).

Interior tests only cull, so this is exactly what makes "best over the leaves whose own exact tests pass" equal to
the reference's recursion.  The parity tests notice a violation only if the culled box held the winning
hit; here every (ray, box) pair is checked.  For rays with a zero / tiny direction component the kernels use the
geometric rule of DESIGN.md section 2; its guarantee -- no leaf the reference could report is rejected -- is checked
on the leaves (axis_rays)."""
import numpy as np
import pytest

from common import HostEmul, same
from test_fuzz import odd_rays, odd_scene


@pytest.fixture(scope="module")
def em():
    return HostEmul("wide")


def grazing_rays(rng, nodes, m):
    """rays through corners, edges and face points of randomly chosen node boxes (tmin ~ tmax: the cases in which
    one rounding decides the reference's test), from origins near and far, plus copies nudged by a few ulps"""
    from common import RAY_DT
    pick = nodes[rng.integers(0, len(nodes), m)]
    lo, hi = pick["lo"].astype(np.float64), pick["hi"].astype(np.float64)
    ok = np.isfinite(lo).all(axis=1) & np.isfinite(hi).all(axis=1)
    lo, hi = lo[ok], hi[ok]
    m = len(lo)
    sel = rng.integers(0, 3, size=(m, 3))                     # per axis: 0 = lo plane, 1 = hi plane, 2 = inside
    inside = lo + (hi - lo) * rng.random((m, 3))
    target = np.where(sel == 0, lo, np.where(sel == 1, hi, inside))
    ext = np.maximum((hi - lo).max(axis=1, keepdims=True), 1e-3)
    origin = target + rng.normal(size=(m, 3)) * ext * rng.choice([0.5, 3.0, 50.0, 3000.0], size=(m, 1))
    d = (target - origin) * rng.choice([1.0, 1e-3, 40.0], size=(m, 1))
    rays = np.zeros(2 * m, RAY_DT)
    rays["o"][:m] = origin.astype(np.float32)
    rays["d"][:m] = d.astype(np.float32)
    rays["o"][m:] = rays["o"][:m]
    nudged = rays["d"][:m].copy().view(np.int32) + rng.integers(-3, 4, size=(m, 3)).astype(np.int32)
    rays["d"][m:] = nudged.view(np.float32)
    return rays


def axis_rays(rng, s, n, m):
    """rays with one or two zero / tiny direction components that pass through (or just beside) spheres: the
    geometric rule for such axes must keep every leaf the reference could report"""
    from common import RAY_DT
    pick = rng.integers(0, n, m)
    c = s["c"][pick].astype(np.float64)
    rad = np.abs(s["r"][pick].astype(np.float64))[:, None]
    zero = rng.random((m, 3)) < 0.45
    zero[zero.all(axis=1), 0] = False                                     # keep one free axis
    zero[~zero.any(axis=1), 2] = True                                     # and at least one zero axis
    off_zero = rng.uniform(-1.2, 1.2, size=(m, 3)) * rad                  # inside, grazing or just outside the slab
    off_free = rng.normal(size=(m, 3)) * rng.choice([2.0, 30.0, 2000.0], size=(m, 1))
    o = c + np.where(zero, off_zero, off_free)
    d = np.where(zero, rng.choice([0.0, -0.0, 1e-35, -1e-41], size=(m, 3)), -off_free * rng.choice([1.0, 0.01], size=(m, 1)))
    rays = np.zeros(m, RAY_DT)
    rays["o"] = o.astype(np.float32)
    rays["d"] = d.astype(np.float32)
    return rays


def _rays_for(oracle, em, h, n_bench=1500):
    cam = oracle.default_camera()
    prim = oracle.primary_rays(cam, 160, 90)
    hits, _ = em.trace(h, prim, threads=8)
    oracle.srand(7)
    parts = [oracle.rays_benchmark(n_bench), prim[::7]]
    if (hits["prim"] >= 0).any():
        parts.append(oracle.bounce_rays(hits)[:n_bench])
    return np.concatenate(parts), cam


@pytest.mark.parametrize("kind,n,build", [("benchmark", 20000, (0, 20000, 0)), ("benchmark", 20000, (0, 19999, 20)),
                                          ("random", 3000, (0, 3000, 0)), ("random", 20, (0, 20, 0))])
def test_interior_tests_never_reject_what_the_reference_accepts(oracle, em, kind, n, build):
    oracle.srand(3)
    s = oracle.scene_benchmark(n) if kind == "benchmark" else oracle.scene_random(n)
    h = em.build(s, n, *build)
    rays, cam = _rays_for(oracle, em, h)
    nodes = em.build_tree(s.copy(), *build)
    r = em.check_culling(h, rays, node_step=1 if n <= 3000 else 3)
    assert r["accepted"] > 100 and r["bad_binary"] == 0 and r["bad_wide"] == 0, r
    r = em.check_culling(h, grazing_rays(np.random.default_rng(n), nodes, 3000))
    assert r["accepted"] > 100 and r["bad_binary"] == 0 and r["bad_wide"] == 0, r
    r = em.check_culling(h, axis_rays(np.random.default_rng(n + 1), s, n, 4000))
    assert r["accepted"] > 100 and r["bad_binary"] == 0 and r["bad_wide"] == 0, r
    f = em.check_frustum(h, cam, 320, 180, tile_step=1 if n <= 3000 else 5, node_step=1 if n <= 3000 else 3)
    assert f["accepted"] > 100 and f["bad"] == 0 and f["bad_simple"] == 0, f
    # the macro-tile kernel culls with the frustum of a 32 x 8 pixel tile (and filters leaves with its 16 x 8 halves, above)
    f = em.check_frustum(h, cam, 320, 180, tile_w=32, tile_h=8, tile_step=1 if n <= 3000 else 3, node_step=1 if n <= 3000 else 3)
    assert f["accepted"] > 100 and f["bad"] == 0 and f["bad_simple"] == 0, f
    em.free(h)


@pytest.mark.parametrize("seed,n", [(2, 17), (3, 64), (4, 300), (5, 1500), (6, 5000), (8, 40)])
def test_interior_tests_on_odd_scenes(oracle, em, seed, n):
    """the fuzzer's scenes: zero / tiny / huge radii, lattices, flat slabs; unnormalised, centre-aimed and far rays"""
    rng = np.random.default_rng(seed)
    s = odd_scene(rng, n)
    depth0 = int(rng.choice([0, 0, 20, 36]))
    h = em.build(s, n, 0, n, depth0)
    rays = np.concatenate([odd_rays(rng, s, n, 6000), grazing_rays(rng, em.build_tree(s.copy(), 0, n, depth0), 3000),
                           axis_rays(rng, s, n, 3000)])
    r = em.check_culling(h, rays)
    assert r["accepted"] > 100 and r["bad_binary"] == 0 and r["bad_wide"] == 0, r
    cam = oracle.default_camera()
    cam["position"] = tuple(float(v) for v in (s["c"][:n].mean(axis=0) + np.array([0.3, 0.2, 3.0 * (np.abs(s["c"][:n]).max() + 1)])))
    f = em.check_frustum(h, cam, 320, 200)
    assert f["bad"] == 0 and f["bad_simple"] == 0, f
    f = em.check_frustum(h, cam, 320, 200, tile_w=32, tile_h=8)
    assert f["bad"] == 0 and f["bad_simple"] == 0, f
    em.free(h)


def limb_rays(rng, s, n, m):
    """rays that graze spheres from near and far: at the limb the reference's FLOAT discriminant (hit.c:21-25) reports
    hits whose t can precede the sphere's own box entry -- the case the 2 r_max / |d| term of the cull bound is for"""
    from common import RAY_DT
    pick = rng.integers(0, n, m)
    c = s["c"][pick].astype(np.float64)
    rad = np.abs(s["r"][pick].astype(np.float64))[:, None]
    u = rng.normal(size=(m, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    w = np.cross(u, rng.normal(size=(m, 3)))
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    dist = rng.choice([3.0, 100.0, 3000.0, 60000.0], size=(m, 1))
    eps = rng.choice([0.0, 1e-6, -1e-6, 1e-4, -1e-4, 3e-3, -3e-3, -0.3], size=(m, 1))
    o = c - u * dist + w * rad * (1.0 + eps)
    rays = np.zeros(m, RAY_DT)
    rays["o"] = o.astype(np.float32)
    rays["d"] = (u * rng.choice([1.0, 1e-3, 25.0], size=(m, 1))).astype(np.float32)
    return rays


@pytest.mark.parametrize("kind,n,build", [("benchmark", 20000, (0, 20000, 0)), ("benchmark", 5000, (0, 4999, 20)),
                                          ("random", 3000, (0, 3000, 0)), ("random", 200, (0, 200, 0)), ("odd", 1500, (0, 1500, 0))])
def test_best_t_bound_never_culls_the_winners_path(oracle, em, kind, n, build):
    """best_s = best_t (1 + 2^-9) + 2 r_max / |d|: every box on the path to the reference's winner passes the
    interior tests under the bound of the winner's own t (the tightest bound any schedule can hold there)."""
    rng = np.random.default_rng(n)
    oracle.srand(5)
    s = odd_scene(rng, n) if kind == "odd" else (oracle.scene_benchmark(n) if kind == "benchmark" else oracle.scene_random(n))
    h = em.build(s, n, *build)
    rays, _ = _rays_for(oracle, em, h, 3000)
    rays = np.concatenate([rays, limb_rays(rng, s, n, 20000), axis_rays(rng, s, n, 3000), odd_rays(rng, s, n, 3000)])
    r = em.check_best_t(h, rays)
    assert r["winners"] > 1000 and r["bad_binary"] == 0 and r["bad_wide"] == 0 and r["bad_frustum"] == 0 and r["disagree"] == 0, r
    # and the winner found this way is the oracle's
    want, _ = oracle.trace_bvh(oracle.build(s.copy(), *build), s, n, rays[:5000])
    got, _ = em.trace(h, rays[:5000])
    assert same(got, want)
    em.free(h)

