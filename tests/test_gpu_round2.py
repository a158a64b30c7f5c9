"""GPU parity tests added in round 2 (run on the B200 box with -m gpu), all through the C-ABI:

* compact outputs (4-byte ids, {t, prim}) and the records derived from them on demand (hit.h:8-14 semantics:
  object index + lazily derived fields) against the oracle;
* the packet kernel's stack discipline on deep trees and coarse frames (one tile sees the whole scene);
* the scalar ray_bvh_intersect cache after a foreign free + address reuse;
* handles without the 64-byte binary tree (the default);
* one process, N devices (gpu_bvh_set) -- sharded frames and ray buffers equal the single-device result;
* reference-generated goldens for the 10 M-sphere scene (tests/golden/make_golden_10m.py);
* the 1 M-sphere 4K frame the bench times: packet == macro-tile == per-ray == ray-buffer over the FULL frame, and a
  200 K-pixel sample against the oracle;
* the macro-tile kernel (SAHBVH_PRIMARY_MACRO) wherever the packet kernel is checked, SAHBVH_PRIMARY_AUTO's sampling
  and settling, random cameras with odd frame sizes, the persisting-L2 window option, the id-count reduction.
"""
import ctypes as C
import json
import os
import time

import numpy as np
import pytest

from common import set_keep_binary, HIT_DT, RAY_DT, assert_hits_equal, same, sha

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gpu(pkg):
    if pkg.device_count() < 1:
        pytest.fail("no CUDA device: the closest-hit path has no CPU fallback")
    set_keep_binary(pkg, False)          # the product default: handles hold only what the default kernels read
    yield pkg
    set_keep_binary(pkg, False)


@pytest.fixture(autouse=True)
def _default_handles(gpu):
    """again per test (other modules' tests may run in between)"""
    set_keep_binary(gpu, False)
    yield
    set_keep_binary(gpu, False)


@pytest.fixture(scope="module")
def scene100k(gpu, oracle):
    oracle.srand(1)
    so = oracle.scene_benchmark(100000)
    nodes = oracle.build(so, 0, 100000, 0)
    gpu.srand(1)
    s = gpu.benchmark_scene(100000)
    bvh = gpu.GpuBvh.build(s, 100000, 0, 100000, 0)
    assert same(s, so)
    yield bvh, s, nodes
    bvh.close()


def test_compact_primary_outputs_and_lazy_records(gpu, oracle, scene100k):
    """gpu_trace_primary_prims / _tprim == the traced records' fields; gpu_hits_from_prims and the scalar
    sahbvh_hit_from_prim rebuild the oracle's 32-byte records bit for bit."""
    bvh, s, nodes = scene100k
    cam = gpu.default_camera()
    W, H = 1000, 700                                    # not a multiple of the 32 x 16 CTA group: ragged edges
    want, _ = oracle.trace_bvh(nodes, s, 100000, oracle.primary_rays(cam, W, H))
    full = bvh.trace_primary(cam, W, H)
    assert_hits_equal(full, want, "records")
    prims = bvh.trace_primary_prims(cam, W, H)
    assert same(prims, want["prim"])
    tp = bvh.trace_primary_tprim(cam, W, H)
    assert same(tp["prim"], want["prim"]) and same(tp["t"], want["t"])
    # a band that starts off the block grid, pinned and pageable output buffers
    band = bvh.trace_primary_prims(cam, W, H, 37, 455)
    assert same(band, want["prim"][37 * W:455 * W])
    pinned = np.empty(W * H, np.int32)
    gpu.host_pin(pinned)
    try:
        bvh.trace_primary_prims(cam, W, H, out=pinned)
        assert same(pinned, want["prim"])
    finally:
        gpu.host_unpin(pinned)
    rebuilt = bvh.hits_from_prims(cam, W, H, prims)
    assert_hits_equal(rebuilt, want, "records rebuilt from ids")
    # hit.h:8-14 on demand, one pixel at a time on the host
    rng = np.random.default_rng(5)
    hit_px = np.nonzero(want["prim"] >= 0)[0]
    for p in list(rng.choice(hit_px, 40, replace=False)) + list(rng.integers(0, W * H, 20)):
        rec = gpu.hit_from_prim(cam, W, H, int(p % W), int(p // W), s, int(prims[p]))
        w = want[p]
        assert rec.hit == int(w["prim"] >= 0)
        got = np.array([rec.t, rec.p.x, rec.p.y, rec.p.z, rec.n.x, rec.n.y, rec.n.z], np.float32)
        assert got.tobytes() == np.array([w["t"], *w["p"], *w["n"]], np.float32).tobytes()
        if rec.hit:
            assert rec.obj == s.ctypes.data + 20 * int(w["prim"])


def test_compact_ray_buffer_output_and_count(gpu, oracle, scene100k):
    bvh, s, nodes = scene100k
    oracle.srand(21)
    rays = oracle.rays_benchmark(30011)
    want, _ = oracle.trace_bvh(nodes, s, 100000, rays)
    prims = bvh.trace_closest_hit_prims(rays)
    assert same(prims, want["prim"])
    assert bvh.benchmark_with_bvh(rays) == int((want["prim"] >= 0).sum())
    assert len(bvh.trace_closest_hit_prims(rays[:0])) == 0


def test_default_handle_has_no_binary_tree(gpu, scene100k):
    """64 MB per million spheres less per handle; the binary-tree kernels refuse cleanly instead of reading NULL."""
    bvh, s, _ = scene100k
    info = bvh.info()
    assert len(bvh.download(0)) == 0
    assert info["device_bytes"] == (info["n_spheres"] + 1) * 20 + info["n_wide"] * 192 + max(len(bvh.download(5)), 32)
    cam = gpu.default_camera()
    try:
        gpu.set_primary_mode(gpu.PRIMARY_PER_RAY)
        with pytest.raises(gpu.SahBvhError, match="binary tree"):
            bvh.trace_primary(cam, 64, 32)
        gpu.set_primary_mode(gpu.PRIMARY_DEFAULT)
        gpu.set_ray_mode(gpu.RAYS_COOP)
        with pytest.raises(gpu.SahBvhError, match="binary tree"):
            bvh.trace_closest_hit(gpu.benchmark_rays(8))
    finally:
        gpu.set_primary_mode(gpu.PRIMARY_DEFAULT)
        gpu.set_ray_mode(gpu.RAYS_WIDE)
    assert (bvh.trace_primary(cam, 64, 32)["prim"] >= -1).all()     # the handle still works


def _spine_scene(n):
    """A scale-invariant progression (ratio 9 per sphere, radius 5 % of the distance) spanning 52 decades: the
    reference's surface-area products overflow, all 21 costs turn NaN and the fallback split (bvh.c:139-141) chains
    empty leaves (Q3) down to the depth limit -- with depth0 = -20 a 60-level tree, the deepest flatten.c admits."""
    from oracle.oracle_py import SPHERE_DT
    s = np.zeros(n + 1, SPHERE_DT)
    x = 1e-20 * 9.0 ** np.arange(n, dtype=np.float64)
    s["c"][:n, 0] = x
    s["c"][:n, 1] = 0.001 * x
    s["c"][:n, 2] = -2.0 * x
    s["r"][:n] = 0.05 * x
    return s


@pytest.mark.parametrize("kind", ["random5000", "spine_depth60", "bench10m"])
def test_packet_stack_on_deep_trees_and_coarse_frames(gpu, oracle, kind):
    """A coarse frame puts the whole scene into one tile's frustum: the packet kernel's shared-memory stack sees
    its widest frontier.  On deep trees (binary depth 40-62) the unchecked pushes used to be able to run past the
    512-entry stack; sahbvh_packet_dfs_from now derives the DFS threshold from the tree depth.  Checked against
    the per-ray kernel (and the oracle where it can build the tree)."""
    set_keep_binary(gpu, True)
    try:
        nodes = None
        if kind == "random5000":
            oracle.srand(3)
            so = oracle.scene_random(5000)
            s = so.copy()
            nodes = oracle.build(so, 0, 5000, 0)
            bvh = gpu.GpuBvh.build(s, 5000, 0, 5000, 0)
            n = 5000
            assert bvh.info()["max_depth"] >= 39
        elif kind == "spine_depth60":
            n = 56
            so = _spine_scene(n)
            s = so.copy()
            nodes = oracle.build(so, 0, n, -20)          # depth limit 40 counted from -20: up to 60 levels
            bvh = gpu.GpuBvh.build(s, n, 0, n, -20)
            assert bvh.info()["max_depth"] >= 59, bvh.info()
        else:
            n = 10_000_000
            gpu.srand(1)
            s = gpu.benchmark_scene(n)
            bvh = gpu.GpuBvh.build(s, n, 0, n, 0)
        cam = gpu.default_camera()
        for W, H in ((16, 8), (64, 36), (33, 17), (320, 200)):
            gpu.set_primary_mode(gpu.PRIMARY_PACKET)
            a = bvh.trace_primary(cam, W, H)
            gpu.set_primary_mode(gpu.PRIMARY_MACRO)
            assert_hits_equal(bvh.trace_primary(cam, W, H), a, f"{kind} {W}x{H}: macro-tile vs packet")
            gpu.set_primary_mode(gpu.PRIMARY_PER_RAY)
            b = bvh.trace_primary(cam, W, H)
            assert_hits_equal(a, b, f"{kind} {W}x{H}: packet vs per-ray")
            if nodes is not None:
                want, _ = oracle.trace_bvh(nodes, so, n, oracle.primary_rays(cam, W, H))
                assert_hits_equal(a, want, f"{kind} {W}x{H}: packet vs oracle")
        bvh.close()
    finally:
        gpu.set_primary_mode(gpu.PRIMARY_DEFAULT)
        set_keep_binary(gpu, False)


def test_packet_stack_bound_and_forced_dfs_mode(gpu, oracle, scene100k):
    """The stack discipline itself: with the DFS threshold forced down to 8 / 40 entries whole frames run in
    single-pop mode (same records), and the instrumented kernel's max_stack never exceeds the bound the launch
    relies on, dfs_from + 24 + 3 * depth."""
    import torch
    bvh, s, nodes = scene100k
    cam = gpu.default_camera()
    W, H = 256, 144
    want, _ = oracle.trace_bvh(nodes, s, 100000, oracle.primary_rays(cam, W, H))
    buf = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    depth = bvh.info()["max_depth"]
    try:
        for mode, stack_entries in ((gpu.PRIMARY_PACKET, 512), (gpu.PRIMARY_MACRO, 256)):     # PKT_STACK / PKT_MACRO_STACK
            gpu.set_primary_mode(mode)
            for forced in (8, 40, 0):
                assert gpu.lib.sahbvh_set_packet_dfs_from(C.c_int(forced)) == 0
                assert_hits_equal(bvh.trace_primary(cam, W, H), want, f"mode {mode} dfs_from={forced}")
                st = bvh.trace_primary_stats(cam, W, H, 0, H, buf.data_ptr())
                dfs_from = forced if forced else stack_entries - 32 - 3 * (depth + 2)
                assert 0 < st["max_stack"] <= dfs_from + 24 + 3 * depth, (mode, forced, st)
                if forced == 8:
                    assert st["max_stack"] > 8            # the frame really ran in single-pop mode above the threshold
    finally:
        gpu.lib.sahbvh_set_packet_dfs_from(C.c_int(0))
        gpu.set_primary_mode(gpu.PRIMARY_DEFAULT)


def test_scalar_cache_survives_foreign_free_and_address_reuse(gpu, oracle):
    """ray_bvh_intersect caches the uploaded tree by root ADDRESS.  A driver that frees the tree with its
    own free_bvh (benchmark.c:81-88) and rebuilds can get the same address for a different tree.  Emulated exactly:
    the root node of tree B is copied over the root node of tree A, so the cached address now heads tree B."""
    class V(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    class RayS(C.Structure):
        _fields_ = [("o", V), ("d", V)]

    class Rec(C.Structure):
        _fields_ = [("t", C.c_float), ("p", V), ("n", V), ("hit", C.c_int), ("obj", C.c_void_p)]

    gpu.lib.ray_bvh_intersect.restype = Rec
    gpu.lib.ray_bvh_intersect.argtypes = [RayS, C.c_void_p]
    cam = gpu.default_camera()
    rays = oracle.primary_rays(cam, 800, 600)[::3001]
    scenes = []
    for seed in (7, 8):
        oracle.srand(seed)
        so = oracle.scene_random(20)
        s = so.copy()
        nodes = oracle.build(so, 0, 20, 0)
        root = gpu.build_bvh_node(s, 0, 20, 0)
        want, _ = oracle.trace_bvh(nodes, so, 20, rays)
        scenes.append((s, root, want))
    (sa, root_a, want_a), (sb, root_b, want_b) = scenes
    assert not same(want_a["prim"], want_b["prim"])

    def trace(root):
        return np.array([gpu.lib.ray_bvh_intersect(RayS(V(*r["o"]), V(*r["d"])), C.c_void_p(root)).t for r in rays], np.float32)

    assert same(trace(root_a), want_a["t"])                 # caches tree A under root_a's address
    saved = C.string_at(root_a, 56)
    C.memmove(root_a, root_b, 56)                           # the same address now is the root of tree B
    try:
        got = trace(root_a)
        assert same(got, want_b["t"]), "stale cache entry: rays were traced against the freed tree"
        rec = gpu.lib.ray_bvh_intersect(RayS(V(*rays[0]["o"]), V(*rays[0]["d"])), C.c_void_p(root_a))
        if rec.hit:
            assert sb.ctypes.data <= rec.obj < sb.ctypes.data + 20 * 20
    finally:
        C.memmove(root_a, saved, 56)
    assert same(trace(root_a), want_a["t"])                 # and back again
    gpu.free_bvh(root_a)
    gpu.free_bvh(root_b)


def test_device_set_one_process_n_devices(gpu, oracle, scene100k):
    """gpu_bvh_replicate + the *_multi entry points: bands / ranges over N replicas (one per visible device, and
    several on one device when the box has a single GPU) reproduce the single-handle results bit for bit."""
    bvh, s, nodes = scene100k
    ndev = gpu.device_count()
    for devices in ([0], [0, 0, 0], list(range(ndev)) if ndev > 1 else [0, 0]):
        st = gpu.GpuBvhSet.replicate(bvh, devices)
        assert st.size() == len(devices)
        assert st.member_info(len(devices) - 1)["n_wide"] == bvh.info()["n_wide"]
        cam = gpu.default_camera()
        for (W, H) in ((640, 360), (333, 97)):
            want = bvh.trace_primary(cam, W, H)
            assert_hits_equal(st.trace_primary(cam, W, H), want, f"multi records {devices} {W}x{H}")
            assert same(st.trace_primary_prims(cam, W, H), want["prim"])
        # frame assembled in the root device's memory through peer stores, then consumed there
        import torch
        W, H = 1280, 720
        want = bvh.trace_primary(cam, W, H)
        with torch.cuda.device(devices[0]):
            d_prims = torch.full((W * H,), -7, dtype=torch.int32, device="cuda")
            d_hits = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
            d_count = torch.zeros(1, dtype=torch.int64, device="cuda")
            stream = torch.cuda.current_stream().cuda_stream
            st.trace_primary_prims_dev(cam, W, H, d_prims.data_ptr(), stream)
            bvh_root = gpu.GpuBvh(C.c_void_p(gpu.lib.gpu_bvh_set_get(st._h, 0)))
            bvh_root.hits_from_prims_dev(cam, W, H, 0, H, d_prims.data_ptr(), d_hits.data_ptr(), stream)
            gpu.count_prims_dev(d_prims.data_ptr(), W * H, d_count.data_ptr(), stream)
            bvh_root._h = None                                   # borrowed handle: the set owns it
            torch.cuda.synchronize()
            assert same(d_prims.cpu().numpy(), want["prim"])
            assert_hits_equal(d_hits.cpu().numpy().view(HIT_DT).reshape(-1), want, "records rebuilt on the root")
            assert int(d_count.item()) == int((want["prim"] >= 0).sum())
        oracle.srand(33)
        rays = oracle.rays_benchmark(10007)
        wr, _ = oracle.trace_bvh(nodes, s, 100000, rays)
        assert_hits_equal(st.trace_closest_hit(rays), wr, "multi ray buffer")
        assert same(st.trace_closest_hit_prims(rays), wr["prim"])
        assert st.benchmark_with_bvh(rays) == int((wr["prim"] >= 0).sum())
        # the path loop: banded frame == single-call frame (draws keyed by the frame-wide pixel index)
        a = bvh.render_frame(cam, 320, 200, seed=5)
        assert same(st.render_frame(cam, 320, 200, seed=5), a)
        st.close()


def test_c_example_render_driver_on_a_device_set(gpu, tmp_path):
    """examples/render_multi_b200.c: main.c's render driver in plain C on the multi-device C-ABI.  The frame must not
    depend on the number of bands (1, 3, and one per device)."""
    import subprocess
    from test_host_logic import _compile_example
    exe = _compile_example(tmp_path, "render_multi_b200")
    sums = []
    for bands in (1, 3, max(gpu.device_count(), 2)):
        out = subprocess.run([exe, str(bands), "5000", "640", "360", "2", str(tmp_path / f"f{bands}.ppm")], capture_output=True, text=True,
                             timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = out.stdout.splitlines()
        sums.append([ln for ln in lines if ln.startswith(("frame checksum", "pixel (", "primary frame"))])
        assert len(sums[-1]) == 3 and "rays hit" in sums[-1][0]
        sums[-1][0] = sums[-1][0].split(",")[0]                     # drop the timing part
        sums[-1][2] = sums[-1][2].split("->")[0]
    assert sums[0] == sums[1] == sums[2], sums
    assert (tmp_path / "f1.ppm").read_bytes() == (tmp_path / "f3.ppm").read_bytes()


def test_peer_wait_and_signal_on_one_device(gpu):
    """sahbvh_peer_wait: the stream advances once the flags are there; a flag that never arrives ends in the
    timeout status, not in a hung stream.  The blocking case is released by HOST stores into pinned flags: two
    streams of one device may share a hardware queue, so a spin-wait must never depend on a kernel of another
    stream of the same device (bench.py orders same-device work with events for that reason)."""
    import torch
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    main, side = torch.cuda.Stream(), torch.cuda.Stream()
    # (1) flags signalled by kernels (sahbvh_peer_signal) before the wait is enqueued: passes at once
    flags = torch.zeros(4, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for r in range(4):
        gpu.peer_signal(flags.data_ptr() + 8 * r, 3, side.cuda_stream)
    side.synchronize()
    gpu.peer_wait(flags.data_ptr(), 4, 3, 4000, status.data_ptr(), main.cuda_stream)
    torch.cuda.synchronize()
    assert int(status.item()) == 0 and flags.tolist() == [3, 3, 3, 3]
    # (2) the wait blocks the stream until the last flag arrives (host writes into page-locked memory)
    hflags = torch.zeros(4, dtype=torch.int64).pin_memory()
    gpu.peer_wait(hflags.data_ptr(), 4, 7, 20000, status.data_ptr(), main.cuda_stream)
    done = torch.cuda.Event()
    done.record(main)
    hflags[:3] = 7
    time.sleep(0.2)
    assert not done.query()
    hflags[3] = 9                                                                       # >= value releases it
    main.synchronize()
    assert done.query() and int(status.item()) == 0
    # (3) never signalled: 50 ms timeout
    gpu.peer_wait(flags.data_ptr(), 4, 4, 50, status.data_ptr(), main.cuda_stream)
    torch.cuda.synchronize()
    assert int(status.item()) == 1


def test_random_cameras_and_odd_frames(gpu, oracle, scene100k):
    """The packet kernel's tile frusta depend on the camera: positions inside and outside the scene, every yaw / pitch
    quadrant (axis-aligned views put zero direction components on whole pixel rows and columns), narrow and wide
    fields of view, frames from one pixel to ragged non-multiples of the 16 x 8 tile -- records equal the oracle's."""
    bvh, s, nodes = scene100k
    rng = np.random.default_rng(2024)
    cams = []
    for k in range(7):
        cam = oracle.default_camera()
        cam["position"] = tuple(float(v) for v in (rng.uniform(-600, 600, 3) if k % 2 else rng.uniform(-80, 80, 3)))
        cam["yaw"] = np.float32([0.0, np.pi / 2, -np.pi, 1.0, -2.5, 0.3, 3.0][k])
        cam["pitch"] = np.float32([0.0, 0.0, 0.7, -1.2, 0.05, 1.5, -0.4][k])
        cam["fov"] = np.float32([45.0, 90.0, 5.0, 120.0, 60.0, 30.0, 170.0][k])
        cams.append(oracle.camera_update(cam))
    sizes = [(333, 211), (64, 40), (1, 1), (15, 7), (17, 130)]
    for k, cam in enumerate(cams):
        for W, H in sizes[k % 2::2] if k else sizes:
            want, _ = oracle.trace_bvh(nodes, s, 100000, oracle.primary_rays(cam, W, H))
            assert_hits_equal(bvh.trace_primary(cam, W, H), want, f"camera {k} frame {W}x{H}")
            assert same(bvh.trace_primary_prims(cam, W, H), want["prim"])


def test_primary_auto_mode_samples_then_settles(gpu, oracle, scene100k):
    """SAHBVH_PRIMARY_AUTO: the first launches of a frame shape alternate between the two packet kernels (identical
    records whichever runs), then the handle settles on one of them for that shape; other shapes are tuned separately
    and counting launches never disturb the sampling."""
    bvh0, s, nodes = scene100k
    cam = gpu.default_camera()
    W, H = 640, 360
    want, _ = oracle.trace_bvh(nodes, s, 100000, oracle.primary_rays(cam, W, H))
    bvh = gpu.GpuBvh.build(s.copy(), 100000, 0, 100000, 0)      # a fresh handle: no tuning history
    try:
        gpu.set_primary_mode(gpu.PRIMARY_AUTO)
        assert bvh.primary_kernel_in_use(W, H) == -1
        seen = set()
        for k in range(12):
            assert_hits_equal(bvh.trace_primary(cam, W, H), want, f"auto mode, launch {k}")
            seen.add(bvh.primary_kernel_in_use(W, H))
        assert bvh.primary_kernel_in_use(W, H) in (gpu.PRIMARY_PACKET, gpu.PRIMARY_MACRO)   # settled (host calls synchronise)
        assert -1 in seen                                                                   # ... after a sampling phase
        assert bvh.primary_kernel_in_use(W, 48) == -1                                       # another shape: its own entry
        assert same(bvh.trace_primary_prims(cam, W, H, 100, 148), want["prim"][100 * W:148 * W])
        for mode in (gpu.PRIMARY_PACKET, gpu.PRIMARY_MACRO, gpu.PRIMARY_PER_RAY):
            gpu.set_primary_mode(mode)
            assert bvh.primary_kernel_in_use(W, H) == mode
    finally:
        gpu.set_primary_mode(gpu.PRIMARY_DEFAULT)
        bvh.close()


def test_l2_persisting_window_option_changes_nothing_but_residency(gpu, oracle, scene100k):
    """sahbvh_set_l2_persist_bytes: trace launches run with an access-policy window over the top of the breadth-first
    node array; records stay bit-identical, and switching it off again restores the default launches."""
    bvh, so, nodes = scene100k
    n = 100000
    oracle.srand(5)
    rays = oracle.rays_benchmark(30000)
    want, _ = oracle.trace_bvh(nodes, so, n, rays)
    cam = gpu.default_camera()
    try:
        base_prim = bvh.trace_primary_prims(cam, 640, 360)
        for mb in (1, 64, 4096):                                   # smaller, about equal to and far larger than the tree
            gpu.set_l2_persist_bytes(mb << 20)
            assert_hits_equal(bvh.trace_closest_hit(rays), want, f"L2 window {mb} MiB")
            assert same(bvh.trace_primary_prims(cam, 640, 360), base_prim)
        gpu.set_l2_persist_bytes(0)
        assert_hits_equal(bvh.trace_closest_hit(rays), want, "L2 window off again")
    finally:
        gpu.set_l2_persist_bytes(0)


def test_count_prims_dev_alignment_and_tails(gpu):
    """gpu_count_prims_dev (the frame consumer's reduction) reads 128-bit words: unaligned starts, short arrays and
    tails must count exactly like benchmark.c:242-244's loop."""
    import torch
    rng = np.random.default_rng(5)
    host = rng.integers(-1, 5, size=100003).astype(np.int32)
    base = torch.from_numpy(host).cuda()
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for off in (0, 1, 2, 3, 5):
        for n in (0, 1, 3, 4, 15, 16, 17, 1000, 65536, 100003 - off):
            cnt.zero_()
            gpu.count_prims_dev(base.data_ptr() + 4 * off, n, cnt.data_ptr(), 0)
            torch.cuda.synchronize()
            assert int(cnt.item()) == int((host[off:off + n] >= 0).sum()), (off, n)


# ---------------------------------------------------------------- the frame the bench times
def test_1m_4k_full_frame_kernels_agree_and_oracle_sample(gpu, oracle):
    """The headline workload itself: 1 M spheres, 3840 x 2160.  packet == per-ray == ray-buffer kernel over the
    FULL frame (8.3 M records, bit-exact), ids-only == records, and 200 K pixels against the oracle."""
    set_keep_binary(gpu, True)
    try:
        n, W, H = 1_000_000, 3840, 2160
        oracle.srand(1)
        so = oracle.scene_benchmark(n)
        nodes = oracle.build(so, 0, n, 0)
        gpu.srand(1)
        s = gpu.benchmark_scene(n)
        bvh = gpu.GpuBvh.build(s, n, 0, n, 0)
        assert same(s, so)
        cam = gpu.default_camera()
        gpu.set_primary_mode(gpu.PRIMARY_PACKET)
        a = bvh.trace_primary(cam, W, H)
        prims = bvh.trace_primary_prims(cam, W, H)
        assert same(prims, a["prim"])
        gpu.set_primary_mode(gpu.PRIMARY_MACRO)
        assert_hits_equal(bvh.trace_primary(cam, W, H), a, "1M 4K: macro-tile vs packet, full frame")
        assert same(bvh.trace_primary_prims(cam, W, H), prims)
        gpu.set_primary_mode(gpu.PRIMARY_PER_RAY)
        b = bvh.trace_primary(cam, W, H)
        gpu.set_primary_mode(gpu.PRIMARY_DEFAULT)
        assert_hits_equal(a, b, "1M 4K: packet vs per-ray, full frame")
        rays = gpu.generate_primary_rays(cam, W, H)
        assert same(rays, oracle.primary_rays(cam, W, H))
        c = bvh.trace_closest_hit(rays)
        assert_hits_equal(a, c, "1M 4K: fused packet vs ray-buffer kernel, full frame")
        rng = np.random.default_rng(2)
        idx = np.unique(np.concatenate([rng.integers(0, W * H, 200_000), (H // 2) * W + np.arange(W), np.arange(H) * W + W // 2]))
        assert len(idx) >= 200_000
        want, _ = oracle.trace_bvh(nodes, so, n, rays[idx])
        assert_hits_equal(a[idx], want, "1M 4K: 200 K-pixel sample vs oracle")
        # config 4 recipe: diffuse-bounce rays off the frame, 50 K against the oracle
        gpu.srand(9)
        br = gpu.bounce_rays(a[::83])
        bh = bvh.trace_closest_hit(br)
        want, _ = oracle.trace_bvh(nodes, so, n, br[:50_000])
        assert_hits_equal(bh[:50_000], want, "1M bounce: 50 K rays vs oracle")
        assert same(bvh.trace_closest_hit_prims(br), bh["prim"])
        bvh.close()
    finally:
        gpu.set_primary_mode(gpu.PRIMARY_DEFAULT)
        set_keep_binary(gpu, False)


# ---------------------------------------------------------------- 10 M spheres, pinned by the reference
def test_10m_reference_goldens(gpu):
    """tests/golden/make_golden_10m.py ran the REFERENCE (158 s build_bvh_node over 10 M spheres) and recorded its
    hits for 20 000 benchmark rays, 20 000 pixels of the 7680 x 4320 frame and 20 000 diffuse-bounce rays, plus the
    SHA-256 of its permuted sphere array and exported tree.  Device builder, flattener and every default kernel are
    checked against them bit for bit."""
    with open(os.path.join(HERE, "golden", "known_answers_10m.json")) as f:
        ans = json.load(f)
    vec = np.load(os.path.join(HERE, "golden", "vectors_10m.npz"))
    n = ans["n_spheres"]
    gpu.srand(1)
    s = gpu.benchmark_scene(n)
    assert sha(s[:n]) == ans["sha_scene_pre_build"]
    bench_rays = gpu.benchmark_rays(len(vec["bench_rays__rays"]))       # same rand() stream, right after the scene
    assert same(bench_rays, vec["bench_rays__rays"])
    tree = gpu.build_tree(s.copy(), 0, n, 0, mode=gpu.BUILD_DEVICE)
    assert len(tree) == ans["tree"]["nodes"] and sha(tree) == ans["sha_tree"], "device-built tree != reference tree"
    del tree
    bvh = gpu.GpuBvh.build(s, n, 0, n, 0)
    assert sha(s[:n]) == ans["sha_scene_post_build"], "permutation differs from the reference's"
    info = bvh.info()
    assert info["max_depth"] == ans["tree"]["max_depth"] and info["multi_leaves"] == ans["tree"]["multi_leaves"]
    hb = bvh.trace_closest_hit(bench_rays)
    assert_hits_equal(hb, vec["bench_rays__hits"].view(HIT_DT).reshape(-1), "10M benchmark rays vs reference")
    assert sha(hb) == ans["cases"]["bench_rays"]["sha_hits"]
    cam = gpu.default_camera()
    W, H = ans["cases"]["primary8k"]["width"], ans["cases"]["primary8k"]["height"]
    idx = vec["primary8k__idx"]
    want = vec["primary8k__hits"].view(HIT_DT).reshape(-1)
    rows = np.unique(idx // W)
    got = np.empty(len(idx), HIT_DT)
    # the frame in 16-row bands that contain sampled pixels (the full 8K frame is 1 GB of records)
    for y0 in np.unique(rows // 16) * 16:
        y1 = min(int(y0) + 16, H)
        band = bvh.trace_primary(cam, W, H, int(y0), y1)
        sel = np.nonzero((idx // W >= y0) & (idx // W < y1))[0]
        got[sel] = band[idx[sel] - int(y0) * W]
    assert_hits_equal(got, want, "10M 8K primary pixels vs reference")
    hp = bvh.trace_closest_hit(vec["primary8k__rays"].view(RAY_DT).reshape(-1))
    assert_hits_equal(hp, want, "10M 8K primary rays through the ray-buffer kernel vs reference")
    hbn = bvh.trace_closest_hit(vec["bounce__rays"].view(RAY_DT).reshape(-1))
    assert_hits_equal(hbn, vec["bounce__hits"].view(HIT_DT).reshape(-1), "10M bounce rays vs reference")
    bvh.close()
