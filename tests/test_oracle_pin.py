"""Pin the oracle (oracle/oracle.c) -- CPU only.

1. against the committed golden fixtures (generated from the compiled reference by
   tests/golden/make_golden.py) -- always runs, also on the GPU box;
2. bit-for-bit against the compiled reference itself (oracle/_ref/libref.so) when it is
   available (in the build container, or as the prebuilt .so that travels).
"""
import numpy as np
import pytest

from common import same, sha, assert_hits_equal


def _bench_case(oracle, n, build, nrays):
    oracle.srand(1)
    s = oracle.scene_benchmark(n)
    nodes = oracle.build(s, *build)
    rays = oracle.rays_benchmark(nrays)
    hits, _ = oracle.trace_bvh(nodes, s, n, rays)
    return s, nodes, rays, hits


@pytest.mark.parametrize("name,n,build,nrays", [
    ("bench_5000_asshipped", 5000, (0, 4999, 20), 10000),
    ("bench_50000_asshipped", 50000, (0, 49999, 20), 10000),
    ("bench_100000_full", 100000, (0, 100000, 0), 100000),
])
def test_golden_benchmark(oracle, golden, name, n, build, nrays):
    answers, vec = golden
    case = answers["cases"][name]
    s, nodes, rays, hits = _bench_case(oracle, n, build, nrays)
    assert oracle.tree_stats(nodes, build[2]) == case["tree"]
    assert sha(s[:n]) == case["sha_scene_post_build"]
    assert sha(nodes) == answers["tree_sha"][name.replace("_asshipped", "")]
    assert sha(rays) == case["sha_rays"]
    assert int((hits["prim"] >= 0).sum()) == case["hits"]
    assert sha(hits) == case["sha_hits"]
    idx = vec[f"{name}__idx"]
    assert same(rays[idx], vec[f"{name}__rays"])
    assert_hits_equal(hits[idx], vec[f"{name}__hits"], name)


@pytest.mark.parametrize("seed", [1, 7, 649, 1164])
def test_golden_default_scene(oracle, golden, seed):
    answers, vec = golden
    case = answers["cases"][f"default_seed{seed}"]
    oracle.srand(seed)
    s = oracle.scene_random(20)
    nodes = oracle.build(s, 0, 20, 0)
    assert oracle.tree_stats(nodes, 0) == case["tree"]
    assert sha(nodes) == answers["tree_sha"][f"default_{seed}"]
    assert same(s[:20], vec[f"default_seed{seed}__spheres"])
    cam = oracle.default_camera()
    rays = oracle.primary_rays(cam, 800, 600)
    assert sha(rays) == case["sha_rays"]
    hits, _ = oracle.trace_bvh(nodes, s, 20, rays)
    assert int((hits["prim"] >= 0).sum()) == case["hits"]
    assert sha(hits) == case["sha_hits"]
    brute, _ = oracle.trace_brute(s, 20, rays)
    assert sha(brute) == case["sha_brute_hits"]
    assert int((brute["prim"] != hits["prim"]).sum()) == case["brute_object_mismatches"]


def test_golden_multi_sphere_leaves(oracle, golden):
    answers, vec = golden
    case = answers["cases"]["random5000_seed3"]
    oracle.srand(3)
    s = oracle.scene_random(5000)
    nodes = oracle.build(s, 0, 5000, 0)
    st = oracle.tree_stats(nodes, 0)
    assert st == case["tree"] and st["multi_leaves"] > 0 and st["empty_leaves"] > 0
    rays = oracle.primary_rays(oracle.default_camera(), 800, 600)[::7].copy()
    hits, _ = oracle.trace_bvh(nodes, s, 5000, rays)
    assert sha(hits) == case["sha_hits"]


def test_golden_4k_rays_bounce_and_camera(oracle, golden):
    answers, vec = golden
    cam = oracle.default_camera()
    pr = oracle.primary_rays(cam, 3840, 2160)
    assert sha(pr) == answers["cases"]["bench_100000_primary4k_sample"]["sha_rays_full_frame"]
    oracle.srand(1)
    s = oracle.scene_benchmark(100000)
    nodes = oracle.build(s, 0, 100000, 0)
    ph, _ = oracle.trace_bvh(nodes, s, 100000, pr[::997])
    assert sha(ph) == answers["cases"]["bench_100000_primary4k_sample"]["sha_hits"]
    case = answers["cases"]["bench_100000_bounce"]
    h8, _ = oracle.trace_bvh(nodes, s, 100000, oracle.primary_rays(cam, 800, 600))
    assert sha(h8) == case["sha_primary_hits_800x600"]
    oracle.srand(5)
    br = oracle.bounce_rays(h8)
    assert sha(br) == case["sha_rays"]
    bh, _ = oracle.trace_bvh(nodes, s, 100000, br)
    assert sha(bh) == case["sha_hits"]
    # camera_update (camera.c:10-18)
    cam2 = oracle.default_camera()
    cam2["yaw"] = np.float32(-2.6)
    cam2["pitch"] = np.float32(0.15)
    oracle.camera_update(cam2)
    assert sha(cam2) == answers["camera_update"]["sha_camera"]
    assert sha(oracle.primary_rays(cam2, 800, 600)) == answers["camera_update"]["sha_rays_800x600"]


def test_walk_counters_return_reference_hits(oracle):
    oracle.srand(1)
    s = oracle.scene_benchmark(20000)
    nodes = oracle.build(s, 0, 20000, 0)
    rays = oracle.rays_benchmark(5000)
    want, _ = oracle.trace_bvh(nodes, s, 20000, rays)
    cnt, got = oracle.walk_counters(nodes, s, 20000, rays, True)
    assert_hits_equal(got, want, "ordered walk")
    assert cnt["inner"] > 0 and cnt["max_stack"] < 64


# ---------------------------------------------------------------- live against the compiled reference
def test_oracle_equals_reference_live(oracle, reflib):
    for n, build in ((1, (0, 1, 0)), (2, (0, 2, 0)), (777, (0, 777, 0)), (20000, (0, 19999, 20))):
        reflib.srand(11 + n)
        sr = reflib.scene_benchmark(n)
        oracle.srand(11 + n)
        so = oracle.scene_benchmark(n)
        assert same(sr, so)
        root, _ = reflib.build(sr, *build)
        nodes = oracle.build(so, *build)
        st = reflib.tree_stats(root, build[2])
        assert st == oracle.tree_stats(nodes, build[2])
        assert same(reflib.export_tree(root, sr, st["nodes"]), nodes)
        assert same(sr, so)
        reflib.srand(5)
        rr = reflib.rays_benchmark(3000)
        oracle.srand(5)
        ro = oracle.rays_benchmark(3000)
        assert same(rr, ro)
        hr, _ = reflib.trace_bvh(root, sr, rr)
        ho, _ = oracle.trace_bvh(nodes, so, n, ro)
        assert_hits_equal(ho, hr, f"n={n}")
        if n <= 777:
            br, _ = reflib.trace_brute(sr, n, rr)
            bo, _ = oracle.trace_brute(so, n, ro)
            assert_hits_equal(bo, br, f"brute n={n}")
        reflib.free(root)


def test_oracle_random_scenes_equal_reference_live(oracle, reflib):
    cam = reflib.default_camera()
    assert same(cam, oracle.default_camera())
    pr = reflib.primary_rays(cam)
    assert same(pr, oracle.primary_rays(cam, 800, 600))
    for seed, n in ((2, 20), (649, 20), (4, 200), (5, 3000)):
        reflib.srand(seed)
        sr = reflib.scene_random(n)
        oracle.srand(seed)
        so = oracle.scene_random(n)
        assert same(sr, so)
        root, _ = reflib.build(sr, 0, n, 0)
        nodes = oracle.build(so, 0, n, 0)
        st = reflib.tree_stats(root, 0)
        assert same(reflib.export_tree(root, sr, st["nodes"]), nodes) and same(sr, so)
        rays = pr[::11].copy()
        hr, _ = reflib.trace_bvh(root, sr, rays)
        ho, _ = oracle.trace_bvh(nodes, so, n, rays)
        assert_hits_equal(ho, hr, f"seed={seed}")
        reflib.srand(99)
        br = reflib.bounce_rays(hr)
        oracle.srand(99)
        bo = oracle.bounce_rays(ho)
        assert same(br, bo)
        reflib.free(root)


def test_oracle_path_loop_equals_reference_trace_ray(oracle, reflib):
    """renderer.c:21-76 + main.c:356-366: the oracle's loop form of trace_ray == the reference's recursion, same
    rand() stream, 800x600, MAX_DEPTH 5, BVH and brute-force branches; and the accumulation of main.c:368-405."""
    cam = oracle.default_camera()
    for seed, n in ((1, 20), (7, 60)):
        reflib.srand(seed)
        s = reflib.scene_random(n)
        so = s.copy()
        root, _ = reflib.build(s, 0, n, 0)
        nodes = oracle.build(so, 0, n, 0)
        assert same(s, so)
        frames_ref, frames = [], []
        for rep in range(3):
            reflib.srand(100 + rep)
            frames_ref.append(reflib.render(cam, s, n, root if rep != 1 else None))
            oracle.srand(100 + rep)
            got, samples = oracle.render(nodes, so, n, cam, 800, 600, use_bvh=(rep != 1))
            frames.append(got)
            assert same(got, frames_ref[-1]), f"seed {seed} frame {rep}"
            assert (np.abs(np.linalg.norm(samples[samples.any(axis=2)], axis=1) - 1) < 1e-6).all()
        assert len(np.unique(frames[0], axis=0)) > 50 and (frames[0] != frames[2]).any()
        # accumulation
        acc_r = np.zeros((800 * 600, 3), np.float32)
        acc_o = np.zeros((800 * 600, 3), np.float32)
        nfr = np.zeros(1, np.int32)
        for k, f in enumerate(frames):
            shown_r = reflib.accumulate(acc_r, nfr, f, first=(k == 0))
            shown_o = oracle.accumulate(acc_o, k + 1, f)
            assert same(shown_r, shown_o) and same(acc_r, acc_o) and int(nfr[0]) == k + 1
        reflib.free(root)


def test_golden_path_loop(oracle, golden):
    """renderer.c:21-76 + main.c:368-405 against the committed fixture (three 800x600 frames of the default scene)."""
    answers, _ = golden
    g = answers["render_default_seed1"]
    oracle.srand(1)
    s = oracle.scene_random(20)
    nodes = oracle.build(s, 0, 20, 0)
    cam = oracle.default_camera()
    acc = np.zeros((800 * 600, 3), np.float32)
    for k in range(3):
        oracle.srand(100 + k)
        frame, _ = oracle.render(nodes, s, 20, cam, 800, 600, use_bvh=(k != 1), want_samples=False)
        assert sha(frame) == g["sha_frames"][k]
        assert sha(oracle.accumulate(acc, k + 1, frame)) == g["sha_shown"][k]
    assert sha(acc) == g["sha_accum"]


@pytest.mark.parametrize("seed,n", [(1, 2), (2, 17), (3, 64), (4, 300), (5, 1500), (6, 5000), (8, 40), (21, 700)])
def test_oracle_equals_reference_on_odd_scenes(oracle, reflib, seed, n):
    """The fuzzer's scenes (tests/test_fuzz.py: zero / tiny / huge radii, lattices of coincident centres, flat slabs,
    axis-parallel and unnormalised rays) through the COMPILED REFERENCE: the oracle the product is fuzzed against must
    itself be the reference on them.  Negative radii are added here: the product refuses them (flatten.c), the
    oracle still has to reproduce what the reference does with them."""
    from test_fuzz import odd_rays, odd_scene
    rng = np.random.default_rng(seed)
    s = odd_scene(rng, n)
    if seed % 2 == 0:
        s["r"][rng.integers(0, n, max(1, n // 20))] *= -1.0
    depth0 = int(rng.choice([0, 0, 20, 36]))
    sr, so = s.copy(), s.copy()
    root, _ = reflib.build(sr, 0, n, depth0)
    nodes = oracle.build(so, 0, n, depth0)
    st = reflib.tree_stats(root, depth0)
    assert st == oracle.tree_stats(nodes, depth0)
    assert same(reflib.export_tree(root, sr, st["nodes"]), nodes) and same(sr, so)
    rays = odd_rays(rng, so, n, 4000)
    hr, _ = reflib.trace_bvh(root, sr, rays)
    ho, _ = oracle.trace_bvh(nodes, so, n, rays)
    assert_hits_equal(ho, hr, f"odd scene seed {seed} n {n} depth0 {depth0}")
    reflib.free(root)
