"""The 4-wide quantised tree of the ray-buffer kernel (csrc/wide.c + sah_wide_visit in
csrc/traverse_core.cuh), CPU only, through the test-only host build of the same sources:

* every decoded child box contains the reference node's box (interior boxes may only grow);
* the leaf set is preserved: regular leaves by sphere index, special leaves (sphere_count != 1,
  SURVEY Q2/Q3) with their exact box in the side table;
* hit records are independent of WHEN the parked leaf tests run (the kernel defers them to a
  warp-wide service round) and of the shared/local split of the traversal stack.
"""
import numpy as np
import pytest

from common import HostEmul, RAY_DT, SPHERE_DT, assert_hits_equal

SPECIAL = 0x40000000
NONE = -0x80000000


@pytest.fixture(scope="module")
def wide():
    return HostEmul("wide")


def decode(nodes):
    """-> per node: lo[4,3], hi[4,3] (float64), ref[4] (int32)"""
    p = nodes[:, 0:3].view(np.float32).astype(np.float64)
    s = nodes[:, [3, 14, 15]].view(np.float32).astype(np.float64)
    ref = nodes[:, 10:14].view(np.int32)
    lo = np.zeros((len(nodes), 4, 3))
    hi = np.zeros((len(nodes), 4, 3))
    for a in range(3):
        for c in range(4):
            ql = (nodes[:, 4 + a] >> (8 * c)) & 0xff
            qh = (nodes[:, 7 + a] >> (8 * c)) & 0xff
            lo[:, c, a] = p[:, a] + (1024.0 + ql) * s[:, a]
            hi[:, c, a] = p[:, a] + (1024.0 + qh) * s[:, a]
    return lo, hi, ref, p, s


@pytest.mark.parametrize("kind,n,build", [("bench", 20000, (0, 20000, 0)), ("bench", 30000, (0, 29999, 20)),
                                           ("random", 5000, (0, 5000, 0)), ("random", 20, (0, 20, 0))])
def test_decoded_boxes_contain_the_spheres(oracle, wide, kind, n, build):
    oracle.srand(11)
    s = oracle.scene_benchmark(n) if kind == "bench" else oracle.scene_random(n)
    h = wide.build(s, n, *build)
    nodes, special = wide.wide_arrays(h)
    info = wide.wide_info(h)
    lo, hi, ref, p, cell = decode(nodes)
    assert np.isfinite(p).all(), "an axis fell back to 'not quantisable' on a finite scene"
    # every regular leaf's decoded box contains its sphere box with the promised margin (cell/512)
    c = s["c"].astype(np.float64)
    r = s["r"].astype(np.float64)
    seen = np.zeros(n, bool)
    for k in range(4):
        rk = ref[:, k]
        leaf = (rk < 0) & (rk != NONE)
        code = ~rk[leaf]
        reg = (code & SPECIAL) == 0
        idx = code[reg]
        assert (idx < n).all()
        seen[idx] = True
        m = cell[leaf][reg] / 512.0
        assert (lo[leaf][reg][:, k, :] <= (c[idx] - r[idx][:, None]) - m).all()
        assert (hi[leaf][reg][:, k, :] >= (c[idx] + r[idx][:, None]) + m).all()
        sp = code[~reg] & ~SPECIAL
        assert (sp < info["n_special"]).all()
        if len(sp):
            rec = special[sp]
            sidx = rec[:, 6].view(np.int32)
            assert ((sidx >= 0) & (sidx < n)).all()
            seen[sidx] = True
            assert (lo[leaf][~reg][:, k, :] <= rec[:, 0:3]).all()
            assert (hi[leaf][~reg][:, k, :] >= rec[:, 3:6]).all()
    # inner children: the child's own grid lies inside the parent's decoded box for that child
    for k in range(4):
        rk = ref[:, k]
        inner = rk >= 0
        ch = rk[inner]
        valid = (ref[ch] != NONE)
        clo = np.where(valid[:, :, None], lo[ch], np.inf).min(axis=1)
        chi = np.where(valid[:, :, None], hi[ch], -np.inf).max(axis=1)
        # child's decoded boxes may stick out by at most its own quantisation (2 cells); the TRUE boxes nest
        tol = 2.0 * cell[ch]
        assert (clo >= lo[inner][:, k, :] - tol).all() and (chi <= hi[inner][:, k, :] + tol).all()
    built = np.arange(build[0], build[1])
    multi = info["n_special"] > 0
    if not multi:
        assert seen[built].all(), "a sphere of the built range is not referenced by any leaf"
    wide.free(h)


def test_result_does_not_depend_on_drain_schedule_or_stack_split(oracle, wide):
    oracle.srand(5)
    s = oracle.scene_random(3000)
    so = s.copy()
    nodes = oracle.build(so, 0, 3000, 0)
    h = wide.build(s, 3000, 0, 3000, 0)
    rays = oracle.primary_rays(oracle.default_camera(), 800, 600)[::11].copy()
    ph, _ = oracle.trace_bvh(nodes, so, 3000, rays)
    oracle.srand(9)
    rays = np.concatenate([rays, oracle.bounce_rays(ph)])
    want, _ = oracle.trace_bvh(nodes, so, 3000, rays)
    visits = {}
    for drain_every in (1, 2, 7, 1000000):
        for fast_depth in (0, 3, 16):
            got, st = wide.trace(h, rays, fast_depth=fast_depth, drain_every=drain_every)
            assert_hits_equal(got, want, f"drain_every={drain_every} fast_depth={fast_depth}")
            visits[drain_every] = st["inner_visits"]
    assert visits[1] <= visits[1000000]   # deferring only delays culling
    wide.free(h)


def test_many_leaves_in_one_node_stall_and_resume(oracle, wide):
    """Coincident spheres: nodes whose 4 children are all hit leaves overflow the two pending slots."""
    s = np.zeros(13, SPHERE_DT)
    s["c"][:12] = (0, 0, -10)
    s["r"][:12] = 1.0
    s["c"][4:8] = (0.25, 0, -10)
    so = s.copy()
    nodes = oracle.build(so, 0, 12, 0)
    h = wide.build(s, 12, 0, 12, 0)
    rays = np.zeros(3, RAY_DT)
    rays["o"] = [(0, 0, 0), (0.25, 0, 0), (0.1, 0.2, 5)]
    rays["d"] = [(0, 0, -1), (0, 0, -1), (0, 0, -1)]
    want, _ = oracle.trace_bvh(nodes, so, 12, rays)
    for drain_every in (1, 3, 1000000):
        got, _ = wide.trace(h, rays, drain_every=drain_every, threads=1)
        assert_hits_equal(got, want, f"coincident spheres, drain_every={drain_every}")
    assert (want["prim"] >= 0).all()
    wide.free(h)


@pytest.mark.parametrize("seed", [649, 1164])
def test_special_leaves_of_degenerate_trees(oracle, wide, golden, seed):
    answers, _ = golden
    oracle.srand(seed)
    s = oracle.scene_random(20)
    h = wide.build(s, 20, 0, 20, 0)
    info = wide.wide_info(h)
    case = answers["cases"][f"default_seed{seed}"]
    assert info["n_special"] >= case["tree"]["multi_leaves"]
    rays = oracle.primary_rays(oracle.default_camera(), 800, 600)
    for drain_every in (1, 5):
        got, _ = wide.trace(h, rays, drain_every=drain_every)
        from common import sha
        assert sha(got) == case["sha_hits"]
    wide.free(h)


def test_tiny_direction_component_does_not_disable_culling(oracle):
    """A ray with one direction component of 1e-5..1e-25 has |o*inv| ~ 1e7..1e27 on that axis.  The absolute slack
    for that term must stay on that axis (it is folded into the near/far plane addends): applied to the whole
    comparison it would accept every box within ~|o*inv| * 2^-20 and the ray would walk most of the tree."""
    from common import HostEmul
    oracle.srand(1)
    n = 60000
    s = oracle.scene_benchmark(n)
    so = s.copy()
    nodes = oracle.build(so, 0, n, 0)
    rng = np.random.default_rng(5)
    m = 4000
    rays = np.zeros(m, RAY_DT)
    rays["o"] = rng.uniform(-450, 450, size=(m, 3)).astype(np.float32)
    d = rng.normal(size=(m, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    axis = rng.integers(0, 3, m)
    tiny = rng.choice([1e-4, 1e-5, 1e-6, 1e-8, 1e-12, 1e-20, 1e-25, 1e-29], m) * rng.choice([-1.0, 1.0], m)
    d[np.arange(m), axis] = tiny
    rays["d"] = d.astype(np.float32)
    ref_rays = rays.copy()
    ref_rays["d"] = rng.normal(size=(m, 3)).astype(np.float32)      # ordinary directions from the same origins
    want, _ = oracle.trace_bvh(nodes, so, n, rays)
    for tree in ("binary", "wide"):
        e = HostEmul(tree)
        h = e.build(s.copy(), n, 0, n, 0)
        got, st = e.trace(h, rays)
        assert_hits_equal(got, want, f"tiny direction component ({tree})")
        _, st_ref = e.trace(h, ref_rays)
        assert st["inner_visits"] < 2.0 * st_ref["inner_visits"], (tree, st, st_ref)
        if tree == "wide":
            assert st["max_visits"] < 6 * st_ref["max_visits"], (st, st_ref)
        e.free(h)
    assert (want["prim"] >= 0).sum() > 20
