"""Host-side logic of the product library, CPU only:

* struct layouts and the exported C-ABI symbol set (no compute calls: there is no GPU here);
* the binned SAH builder == the oracle's tree, bounds and sphere permutation;
* the flattener + the per-ray traversal core (csrc/traverse_core.cuh compiled for the host by
  tests/host_emul -- test-only) == the oracle's exhaustive recursion, including the
  reference's quirks (ties, depth-40 leaves, empty leaves, zero direction components);
* loud failure without a CUDA device (no CPU fallback).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import (HIT_DT, NODE_DT, RAY_DT, SPHERE_DT, ROOT, PKG_DIR, _ptr, assert_hits_equal, same, sha)


# ---------------------------------------------------------------- layout + ABI
def test_layout_matches_reference(golden, pkg):
    answers, _ = golden
    lay = answers["layout"]
    # Vec3, Ray, Camera, Sphere(+radius,color offsets), AABB, BVHNode(+4 offsets), HitRecord(+4 offsets), W, H, MAX_DEPTH
    assert lay[:17] == [12, 24, 64, 20, 12, 16, 24, 56, 24, 32, 40, 48, 40, 4, 16, 28, 32]
    assert lay[17:20] == [800, 600, 5]
    assert pkg.RAY_DT.itemsize == 24 and pkg.SPHERE_DT.itemsize == 20
    assert pkg.HIT_DT.itemsize == 32 and pkg.CAMERA_DT.itemsize == 64
    # include/sahbvh_types.h carries the same numbers as static asserts; compile-check it as C and C++
    import subprocess
    for comp, std in (("gcc", "-std=c11"), ("g++", "-std=c++17")):
        subprocess.check_call([comp, std, "-fsyntax-only", "-x", "c" if comp == "gcc" else "c++",
                               "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "include", "sahbvh_gpu.h")])


def test_layout_live(reflib):
    assert reflib.layout[:17] == [12, 24, 64, 20, 12, 16, 24, 56, 24, 32, 40, 48, 40, 4, 16, 28, 32]


def test_library_exports_every_declared_symbol(pkg):
    declared = set()
    for hdr in ("sahbvh_gpu.h", "sahbvh_ref_api.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text):
            declared.add(m.group(1))
    declared -= {"SAHBVH_STATIC_ASSERT", "sizeof", "offsetof"}
    assert declared, "no declarations parsed"
    assert declared == set(pkg.EXPORTED_SYMBOLS), declared ^ set(pkg.EXPORTED_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(pkg.lib, name), f"{name} not exported by libsahbvh_b200.so"


def test_product_does_not_link_oracle(pkg):
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg.LIB_PATH]).decode()
    assert "orc_" not in out and "refh_" not in out and "emul_" not in out
    srcs = os.listdir(os.path.join(PKG_DIR, "csrc"))
    for f in srcs + ["../__init__.py"]:
        text = open(os.path.join(PKG_DIR, "csrc", f)).read()
        assert "oracle" not in text.replace("the oracle", "").replace("oracle's", "").lower() or f == "traverse_core.cuh", f


def test_no_gpu_fails_loudly(pkg):
    if pkg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    pkg.srand(1)
    s = pkg.benchmark_scene(100)
    with pytest.raises(pkg.SahBvhError) as e:
        pkg.GpuBvh.build(s, 100, 0, 100, 0)
    assert "no CPU traversal" in str(e.value) or "CUDA" in str(e.value)


# ---------------------------------------------------------------- host helpers == oracle
def test_scene_and_ray_recipes_match_oracle(pkg, oracle):
    for n in (1, 20, 3000):
        pkg.srand(3)
        a = pkg.benchmark_scene(n)
        ra = pkg.benchmark_rays(100)
        oracle.srand(3)
        b = oracle.scene_benchmark(n)
        rb = oracle.rays_benchmark(100)
        assert same(a, b) and same(ra, rb)
        pkg.srand(4)
        a = pkg.random_scene(n)
        oracle.srand(4)
        b = oracle.scene_random(n)
        assert same(a, b)
    assert same(pkg.default_camera(), oracle.default_camera())
    cam = pkg.default_camera()
    pkg.lib.sahbvh_default_camera(_ptr(cam))
    assert same(cam, oracle.default_camera())
    cam["yaw"] = np.float32(-2.6)
    cam["pitch"] = np.float32(0.15)
    cam2 = cam.copy()
    pkg.lib.camera_update(_ptr(cam))
    oracle.camera_update(cam2)
    assert same(cam, cam2)


def test_build_bvh_node_pointer_tree_matches_oracle(pkg, oracle, emul):
    """build_bvh_node (product, binned) -> pointer tree -> flatten -> same hits; same permutation."""
    for n, build in ((5000, (0, 4999, 20)), (3000, (0, 3000, 0)), (1, (0, 1, 0)), (2, (0, 2, 0))):
        oracle.srand(21)
        so = oracle.scene_benchmark(n)
        sp = so.copy()
        nodes = oracle.build(so, *build)
        root = pkg.build_bvh_node(sp, *build)
        assert same(sp, so)
        h = emul.from_pointers(sp, n, root)
        rays = oracle.rays_benchmark(4000)
        want, _ = oracle.trace_bvh(nodes, so, n, rays)
        got, _ = emul.trace(h, rays)
        assert_hits_equal(got, want, f"pointer tree n={n}")
        emul.free(h)
        pkg.free_bvh(root)


def test_reference_built_tree_flattens(reflib, oracle, emul):
    """gpu_bvh_upload's input contract: a tree built by the REFERENCE's build_bvh_node."""
    reflib.srand(8)
    s = reflib.scene_random(300)
    root, _ = reflib.build(s, 0, 300, 0)
    rays = oracle.primary_rays(oracle.default_camera(), 800, 600)[::13].copy()
    want, _ = reflib.trace_bvh(root, s, rays)
    h = emul.from_pointers(s, 300, root)
    got, _ = emul.trace(h, rays)
    assert_hits_equal(got, want, "reference-built tree")
    emul.free(h)
    reflib.free(root)


@pytest.mark.parametrize("n,build,kind", [
    (0, (0, 0, 0), "bench"), (1, (0, 1, 0), "bench"), (2, (0, 2, 0), "bench"), (3, (0, 3, 0), "bench"),
    (20, (0, 20, 0), "bench"), (5000, (0, 5000, 0), "bench"), (50000, (0, 49999, 20), "bench"),
    (100000, (0, 100000, 0), "bench"), (20, (0, 20, 0), "random"), (200, (0, 200, 0), "random"),
    (5000, (0, 5000, 0), "random"), (64, (10, 50, 0), "random"), (64, (0, 64, 39), "random"), (64, (0, 64, 40), "random"),
])
def test_builder_and_traversal_match_oracle(oracle, emul, n, build, kind):
    oracle.srand(1 + n)
    so = oracle.scene_benchmark(n) if kind == "bench" else oracle.scene_random(n)
    sp, sp2 = so.copy(), so.copy()
    nodes = oracle.build(so, *build)
    tree = emul.build_tree(sp, *build)
    assert same(tree, nodes), "tree (bounds / topology / leaf sphere / count)"
    assert same(sp, so), "sphere permutation"
    h = emul.build(sp2, n, *build)
    if kind == "bench":
        rays = oracle.rays_benchmark(6000)
    else:
        rays = oracle.primary_rays(oracle.default_camera(), 800, 600)[::37].copy()
    want, _ = oracle.trace_bvh(nodes, so, n, rays)
    got, st = emul.trace(h, rays)
    assert_hits_equal(got, want, f"{kind} n={n} build={build}")
    assert st["max_stack"] < 64
    emul.free(h)


@pytest.mark.parametrize("seed", [649, 1164])
def test_degenerate_seeds(oracle, emul, golden, seed):
    """SURVEY Q2/Q3 regression fixtures: empty leaves + a depth-40 two-sphere leaf."""
    answers, vec = golden
    oracle.srand(seed)
    s = oracle.scene_random(20)
    h = emul.build(s, 20, 0, 20, 0)
    assert same(s[:20], vec[f"default_seed{seed}__spheres"])
    rays = oracle.primary_rays(oracle.default_camera(), 800, 600)
    got, _ = emul.trace(h, rays)
    assert sha(got) == answers["cases"][f"default_seed{seed}"]["sha_hits"]
    emul.free(h)


def _scene(spheres):
    s = np.zeros(len(spheres) + 1, SPHERE_DT)
    for i, (c, r) in enumerate(spheres):
        s[i]["c"] = c
        s[i]["r"] = r
    return s


def _rays(lst):
    r = np.zeros(len(lst), RAY_DT)
    for i, (o, d) in enumerate(lst):
        r[i]["o"] = o
        r[i]["d"] = d
    return r


def test_ties_go_to_larger_index(oracle, emul):
    """hit.c:108: equal t -> right child -> larger post-build index (Q1); brute force is the opposite."""
    s = _scene([((0, 0, -10), 1.0)] * 6 + [((3, 0, -10), 1.0)] * 3)
    so = s.copy()
    nodes = oracle.build(so, 0, 9, 0)
    h = emul.build(s, 9, 0, 9, 0)
    assert same(s, so)
    rays = _rays([((0, 0, 0), (0, 0, -1)), ((3, 0, 0), (0, 0, -1)), ((0.5, 0.2, 5), (0, 0, -1))])
    want, _ = oracle.trace_bvh(nodes, so, 9, rays)
    got, _ = emul.trace(h, rays)
    assert_hits_equal(got, want, "ties")
    brute, _ = oracle.trace_brute(so, 9, rays)
    assert (want["prim"] >= brute["prim"]).all() and (want["t"] == brute["t"]).all()
    emul.free(h)


def test_zero_and_tiny_direction_components(oracle, emul):
    """hit.c:54-56: a zero direction component accepts the slab regardless of the origin (Q7)."""
    oracle.srand(2)
    s = oracle.scene_random(400)
    so = s.copy()
    nodes = oracle.build(so, 0, 400, 0)
    h = emul.build(s, 400, 0, 400, 0)
    dirs = [(0, 0, -1), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, -1), (0, 1, -1), (1, 1, 0), (0, 0, 0),
            (1e-38, 0, -1), (1e-31, 1e-20, -1), (-0.0, 0.0, -1), (1e-45, 1e-45, -1), (3e-39, 1, 1e-39)]
    lst = []
    for ox in (-30.0, 0.0, 12.5):
        for oy in (-10.0, 4.0):
            for oz in (50.0, 0.0, -3.0):
                for d in dirs:
                    lst.append(((ox, oy, oz), d))
    rays = _rays(lst)
    want, _ = oracle.trace_bvh(nodes, so, 400, rays)
    got, _ = emul.trace(h, rays)
    assert_hits_equal(got, want, "zero/tiny direction")
    assert (want["prim"] >= 0).sum() > 10
    emul.free(h)


def test_origin_inside_and_far_and_unnormalised(oracle, emul):
    oracle.srand(5)
    s = oracle.scene_random(300)
    so = s.copy()
    nodes = oracle.build(so, 0, 300, 0)
    h = emul.build(s, 300, 0, 300, 0)
    rng = np.random.default_rng(0)
    n = 6000
    rays = np.zeros(n, RAY_DT)
    # origins: sphere centres (inside), far away, on surfaces; directions: unnormalised, scaled 1e-3..1e3
    c = so["c"][rng.integers(0, 300, n)]
    rays["o"] = c + rng.normal(size=(n, 3)).astype(np.float32) * rng.choice([0.0, 0.5, 5.0, 3000.0], size=(n, 1)).astype(np.float32)
    rays["d"] = rng.normal(size=(n, 3)).astype(np.float32) * rng.choice([1e-3, 1.0, 1e3], size=(n, 1)).astype(np.float32)
    want, _ = oracle.trace_bvh(nodes, so, 300, rays)
    got, _ = emul.trace(h, rays)
    assert_hits_equal(got, want, "inside/far/unnormalised")
    emul.free(h)


def test_bounce_rays_start_on_surfaces(oracle, emul):
    """renderer.c:54: bounce origins are un-offset hit points (Q13) -- self hits excluded only by t>1e-6."""
    oracle.srand(1)
    s = oracle.scene_random(2000)
    so = s.copy()
    nodes = oracle.build(so, 0, 2000, 0)
    h = emul.build(s, 2000, 0, 2000, 0)
    pr = oracle.primary_rays(oracle.default_camera(), 800, 600)[::5].copy()
    ph, _ = oracle.trace_bvh(nodes, so, 2000, pr)
    oracle.srand(77)
    br = oracle.bounce_rays(ph)
    want, _ = oracle.trace_bvh(nodes, so, 2000, br)
    got, _ = emul.trace(h, br)
    assert_hits_equal(got, want, "bounce")
    assert (want["prim"] >= 0).mean() > 0.2
    emul.free(h)


def test_emulated_primary_rays_match_oracle(oracle, emul):
    cam = oracle.default_camera()
    for w, hgt in ((800, 600), (3840, 2160), (1023, 517), (7, 3)):
        assert same(emul.primary_rays(cam, w, hgt), oracle.primary_rays(cam, w, hgt))
    cam["yaw"] = np.float32(-2.2)
    cam["pitch"] = np.float32(-0.3)
    oracle.camera_update(cam)
    assert same(emul.primary_rays(cam, 640, 360, 8, 64), oracle.primary_rays(cam, 640, 360, 8, 64))


def test_scalar_primitives_match_oracle(pkg, oracle):
    """ray_sphere_intersect / ray_aabb_intersect by value through the reference's signatures."""
    class V(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    class RayS(C.Structure):
        _fields_ = [("o", V), ("d", V)]

    class Box(C.Structure):
        _fields_ = [("lo", V), ("hi", V)]

    class Sph(C.Structure):
        _fields_ = [("c", V), ("r", C.c_float), ("rgba", C.c_uint8 * 4)]

    class Rec(C.Structure):
        _fields_ = [("t", C.c_float), ("p", V), ("n", V), ("hit", C.c_int), ("obj", C.c_void_p)]

    pkg.lib.ray_sphere_intersect.restype = Rec
    pkg.lib.ray_sphere_intersect.argtypes = [RayS, C.POINTER(Sph)]
    pkg.lib.ray_aabb_intersect.restype = C.c_int
    pkg.lib.ray_aabb_intersect.argtypes = [RayS, Box]
    oracle.lib.orc_ray_aabb.argtypes = [RayS, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3)]
    rng = np.random.default_rng(1)
    agree = 0
    for i in range(3000):
        o = rng.normal(size=3) * 20
        d = rng.normal(size=3) * rng.choice([0, 1, 1, 1], size=3)
        lo = rng.normal(size=3) * 10
        hi = lo + np.abs(rng.normal(size=3)) * 8
        ray = RayS(V(*o), V(*d))
        a = pkg.lib.ray_aabb_intersect(ray, Box(V(*lo), V(*hi)))
        b = oracle.lib.orc_ray_aabb(ray, (C.c_float * 3)(*lo), (C.c_float * 3)(*hi))
        assert a == b
        c = rng.normal(size=3) * 10
        sph = Sph(V(*c), float(abs(rng.normal()) * 6), (C.c_uint8 * 4)(1, 2, 3, 255))
        if i % 2:   # aim at the sphere so that a good share of the cases are hits
            aim = (c - o) + rng.normal(size=3) * 2
            ray = RayS(V(*o), V(*aim))
        rec = pkg.lib.ray_sphere_intersect(ray, C.byref(sph))
        s1 = np.zeros(1, SPHERE_DT)
        s1["c"] = (sph.c.x, sph.c.y, sph.c.z)
        s1["r"] = sph.r
        r1 = np.zeros(1, RAY_DT)
        r1["o"] = (ray.o.x, ray.o.y, ray.o.z)
        r1["d"] = (ray.d.x, ray.d.y, ray.d.z)
        want, _ = oracle.trace_brute(s1, 1, r1, threads=1)
        assert rec.hit == int(want["prim"][0] >= 0)
        if rec.hit:
            agree += 1
            got = np.array([rec.t, rec.p.x, rec.p.y, rec.p.z, rec.n.x, rec.n.y, rec.n.z], np.float32)
            w = np.array([want["t"][0], *want["p"][0], *want["n"][0]], np.float32)
            assert got.tobytes() == w.tobytes()
    assert agree > 100


def test_lazy_hit_record_from_primitive_id(pkg, oracle):
    """hit.h:8-14 on demand: sahbvh_pixel_ray == main.c:362-365 + ray.c:17-32 for any W x H, and
    sahbvh_hit_from_prim(ray, spheres, prim) == the record the traversal returned for that ray (host arithmetic)."""
    oracle.srand(7)
    s = oracle.scene_random(20)
    nodes = oracle.build(s, 0, 20, 0)
    cam = oracle.default_camera()
    for W, H in ((800, 600), (333, 97)):
        rays = oracle.primary_rays(cam, W, H)
        want, _ = oracle.trace_bvh(nodes, s, 20, rays)
        rng = np.random.default_rng(W)
        picks = list(rng.integers(0, W * H, 200)) + list(np.nonzero(want["prim"] >= 0)[0][:200])
        nhit = 0
        for p in picks:
            x, y = int(p % W), int(p // W)
            pkg.lib.sahbvh_pixel_ray.restype = pkg._Ray
            pkg.lib.sahbvh_pixel_ray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
            r = pkg.lib.sahbvh_pixel_ray(cam.ctypes.data, W, H, x, y)
            got_ray = np.array([r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z], np.float32)
            assert got_ray.tobytes() == rays[p].tobytes()
            rec = pkg.hit_from_prim(cam, W, H, x, y, s, int(want["prim"][p]))
            assert rec.hit == int(want["prim"][p] >= 0)
            got = np.array([rec.t, rec.p.x, rec.p.y, rec.p.z, rec.n.x, rec.n.y, rec.n.z], np.float32)
            assert got.tobytes() == np.array([want["t"][p], *want["p"][p], *want["n"][p]], np.float32).tobytes()
            if rec.hit:
                nhit += 1
                assert rec.obj == s.ctypes.data + 20 * int(want["prim"][p])
            else:
                assert rec.obj is None
        assert nhit >= 100


def test_write_ppm(pkg, tmp_path):
    rgba = (np.arange(7 * 5 * 4) % 251).astype(np.uint8).reshape(-1, 4)
    path = str(tmp_path / "frame.ppm")
    pkg.write_ppm(path, rgba, 7, 5)
    data = open(path, "rb").read()
    assert data.startswith(b"P6\n7 5\n255\n")
    assert data[len(b"P6\n7 5\n255\n"):] == rgba[:, :3].tobytes()


def _compile_example(tmp_path, name="benchmark_b200"):
    import subprocess
    exe = str(tmp_path / name)
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", name + ".c"), "-L", PKG_DIR, "-lsahbvh_b200",
                           "-Wl,-rpath," + PKG_DIR, "-lm", "-o", exe])
    return exe


def test_c_example_compiles_and_links(pkg, tmp_path):
    """examples/benchmark_b200.c = benchmark.c's driver loop on the C-ABI: plain C, the public headers only."""
    import subprocess
    exe = _compile_example(tmp_path)
    if pkg.device_count() == 0:     # no GPU here: the binary must refuse loudly instead of falling back to a CPU path
        out = subprocess.run([exe, "1", "5000", "5000", "5000"], capture_output=True, text=True, cwd=str(tmp_path))
        assert out.returncode != 0 and "no usable CUDA device" in out.stderr
    exe = _compile_example(tmp_path, "render_multi_b200")     # main.c's render driver on N GPUs (gpu_bvh_set)
    if pkg.device_count() == 0:
        out = subprocess.run([exe], capture_output=True, text=True, cwd=str(tmp_path))
        assert out.returncode != 0 and "no usable CUDA device" in out.stderr


def test_public_headers_are_strict_c99_and_cxx17(tmp_path):
    """include/*.h are the drop-in boundary: they must be usable from the reference's plain C and from C++ hosts."""
    import subprocess
    src = '#include "sahbvh_gpu.h"\n#include "sahbvh_ref_api.h"\nint main(void) { return 0; }\n'
    inc = os.path.join(ROOT, "include")
    c = tmp_path / "hdr.c"
    c.write_text(src)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", str(c)])
    cpp = tmp_path / "hdr.cpp"
    cpp.write_text(src)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", str(cpp)])


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the reference's own CPU path through oracle/_ref, or the oracle port) prints the
    contract's JSON line without touching CUDA; a small scene keeps it to a few seconds here."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--spheres", "20000",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "closest_hit_grays_per_s" and d["unit"] == "Grays/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "Grays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "rays per step" in cb["sample"]
    assert "workload" in d["config"] and "20000" in d["config"]["workload"]


@pytest.mark.parametrize("kind,n", [("benchmark", 20000), ("default", 20), ("odd", 700)])
def test_packet_and_macro_tile_algorithm_on_the_host(oracle, kind, n):
    """The packet kernels' algorithm restated on the host (tests/host_emul, emul_trace_primary_packet over the same
    traverse_core.cuh arithmetic): tiles culled once by their frustum (K2p: 16x8), and the macro-tile variant (K2m:
    32x8 tiles, a candidate leaf tested only by the rays of the sub-tiles / quadrants whose own frustum and cull bound it
    passes) give the oracle's records for every tiling, and the filter removes sphere tests without adding node visits."""
    from common import HostEmul
    em = HostEmul("binary")
    if kind == "benchmark":
        oracle.srand(1)
        s = oracle.scene_benchmark(n)
    elif kind == "default":
        oracle.srand(7)
        s = oracle.scene_random(n)
    else:
        from test_fuzz import odd_scene
        s = odd_scene(np.random.default_rng(3), n)
    nodes = oracle.build(s, 0, n, 0)
    h = em.build(s.copy(), n, 0, n, 0)
    cam = oracle.default_camera()
    if kind == "odd":
        cam["position"] = tuple(float(v) for v in (s["c"][:n].mean(axis=0) + np.array([0.3, 0.2, 3.0 * (np.abs(s["c"][:n]).max() + 1)])))
    W, H = 333, 203                                   # ragged: partial tiles on the right and at the bottom
    want, _ = oracle.trace_bvh(nodes, s, n, oracle.primary_rays(cam, W, H))
    stats = {}
    for tiling in ((16, 8, 0, 0), (32, 8, 16, 8), (32, 8, 8, 4), (32, 16, 8, 4), (33, 7, 5, 3)):
        got, st = em.trace_primary_packet(h, cam, W, H, *tiling)
        assert_hits_equal(got, want, f"{kind}: tiling {tiling}")
        stats[tiling] = st
    assert stats[(32, 8, 8, 4)]["inner_visits"] == stats[(32, 8, 16, 8)]["inner_visits"]        # the filter never changes the traversal ...
    assert stats[(32, 8, 8, 4)]["sphere_tests"] <= stats[(32, 8, 16, 8)]["sphere_tests"]        # ... it only removes sphere tests
    assert stats[(32, 8, 16, 8)]["inner_visits"] < stats[(16, 8, 0, 0)]["inner_visits"] or kind == "default"
    em.free(h)
