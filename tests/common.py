"""Shared helpers for the tests (test infrastructure only)."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG_DIR = os.path.join(ROOT, "cs201_sah-bvh_ray_tracer_b200")
from oracle.oracle_py import HIT_DT, NODE_DT, RAY_DT, SPHERE_DT, CAMERA_DT, _ptr  # noqa: E402,F401


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def same(a, b):
    return np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()


def mismatching_rows(a, b):
    a = np.ascontiguousarray(a).view(np.uint8).reshape(len(a), -1)
    b = np.ascontiguousarray(b).view(np.uint8).reshape(len(b), -1)
    return np.nonzero((a != b).any(axis=1))[0]


def assert_hits_equal(got, want, what=""):
    """Bit-exact comparison of 32-byte hit records (integer index AND float payload)."""
    assert len(got) == len(want)
    bad = mismatching_rows(got, want)
    if len(bad):
        i = int(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {len(got)} hit records differ; first at {i}: got {got[i]} want {want[i]}")


class HostEmul:
    """tests/host_emul/libsahbvh_host_emul.so: csrc/traverse_core.cuh compiled for the host (test-only)."""

    def __init__(self, tree="binary"):
        self.tree = tree   # "binary": 64-byte binary nodes (sah_trace_one); "wide": 4-wide quantised nodes (sah_wide_trace_one)
        so = os.path.join(ROOT, "tests", "host_emul", "libsahbvh_host_emul.so")
        srcs = [os.path.join(ROOT, "tests", "host_emul", "emul.cpp"),
                os.path.join(PKG_DIR, "csrc", "traverse_core.cuh"),
                os.path.join(PKG_DIR, "csrc", "flatten.c"),
                os.path.join(PKG_DIR, "csrc", "wide.c"),
                os.path.join(PKG_DIR, "csrc", "sah_builder.c"),
                os.path.join(PKG_DIR, "csrc", "host_ref_api.c")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-s", "-C", PKG_DIR, "emul"])
        self.lib = C.CDLL(so)
        self.lib.emul_build.restype = C.c_void_p
        self.lib.emul_from_pointers.restype = C.c_void_p
        self.lib.emul_build_tree.restype = C.c_int64

    def build_tree(self, spheres, start, end, depth):
        cap = 64 * max(end - start, 1) + 1024
        out = np.zeros(cap, NODE_DT)
        used = self.lib.emul_build_tree(_ptr(spheres), C.c_int(start), C.c_int(end), C.c_int(depth), _ptr(out), C.c_int64(cap))
        assert used > 0
        return out[:used].copy()

    def build(self, spheres, n_total, start, end, depth):
        h = self.lib.emul_build(_ptr(spheres), C.c_int(n_total), C.c_int(start), C.c_int(end), C.c_int(depth))
        if not h:
            self.lib.sahbvh_last_error.restype = C.c_char_p
            raise RuntimeError(self.lib.sahbvh_last_error().decode())
        return h

    def from_pointers(self, spheres, n, root):
        h = self.lib.emul_from_pointers(_ptr(spheres), C.c_int(n), C.c_void_p(root))
        assert h
        return h

    def free(self, h):
        self.lib.emul_free(C.c_void_p(h))

    def wide_info(self, h):
        out = (C.c_int32 * 3)()
        assert self.lib.emul_wide_info(C.c_void_p(h), out) == 0
        return {"n_wide": out[0], "n_special": out[1], "max_depth": out[2]}

    def wide_arrays(self, h):
        info = self.wide_info(h)
        nodes = np.zeros((info["n_wide"], 16), np.uint32)
        special = np.zeros((max(info["n_special"], 1), 8), np.float32)
        assert self.lib.emul_wide_copy(C.c_void_p(h), _ptr(nodes), _ptr(special)) == 0
        return nodes, special[:info["n_special"]]

    def trace(self, h, rays, threads=8, tree=None, fast_depth=16, drain_every=7):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), HIT_DT)
        st = (C.c_uint64 * 7)()
        if (tree or self.tree) == "wide":
            rc = self.lib.emul_wide_trace(C.c_void_p(h), _ptr(rays), C.c_int64(len(rays)), _ptr(hits), st, C.c_int(threads),
                                          C.c_int(fast_depth), C.c_int(drain_every))
            assert rc == 0
        else:
            self.lib.emul_trace(C.c_void_p(h), _ptr(rays), C.c_int64(len(rays)), _ptr(hits), st, C.c_int(threads))
        stats = dict(zip(("rays", "inner_visits", "sphere_tests", "exact_box_tests", "hits", "max_stack", "max_visits"), map(int, st)))
        return hits, stats

    def check_culling(self, h, rays, node_step=1, threads=8):
        """(ray, box) pairs the reference's slab test accepts but an interior test rejects -- must be zero."""
        rays = np.ascontiguousarray(rays)
        out = (C.c_uint64 * 4)()
        rc = self.lib.emul_check_culling(C.c_void_p(h), _ptr(rays), C.c_int64(len(rays)), C.c_int(node_step), C.c_int(threads), out)
        assert rc == 0
        return dict(zip(("pairs", "accepted", "bad_binary", "bad_wide"), map(int, out)))

    def check_best_t(self, h, rays, threads=8):
        """boxes on the path to the reference's winner that the best-t bound of the winner itself would cull -- must be zero"""
        rays = np.ascontiguousarray(rays)
        out = (C.c_uint64 * 6)()
        rc = self.lib.emul_check_best_t(C.c_void_p(h), _ptr(rays), C.c_int64(len(rays)), C.c_int(threads), out)
        assert rc == 0
        return dict(zip(("winners", "steps", "bad_binary", "bad_wide", "disagree", "bad_frustum"), map(int, out)))

    def check_frustum(self, h, cam, w, hgt, tile_w=16, tile_h=8, tile_step=1, node_step=1):
        """(tile, box) pairs where the reference accepts the box for some ray of the tile but the tile's frustum
        test rejects it -- must be zero."""
        out = (C.c_uint64 * 4)()
        rc = self.lib.emul_check_frustum(C.c_void_p(h), _ptr(cam), C.c_int(w), C.c_int(hgt), C.c_int(tile_w), C.c_int(tile_h),
                                         C.c_int(tile_step), C.c_int(node_step), out)
        assert rc == 0
        return dict(zip(("pairs", "accepted", "bad", "bad_simple"), map(int, out)))

    def trace_primary_packet(self, h, cam, w, hgt, tile_w=16, tile_h=8, sub_w=0, sub_h=0, y0=0, y1=None):
        """host restatement of the packet kernels' algorithm: tile_w x tile_h tiles culled by their frustum; with
        sub_w, sub_h > 0 the macro-tile variant (candidates filtered by sub-tile frusta with their own cull bounds)"""
        y1 = hgt if y1 is None else y1
        hits = np.zeros((y1 - y0) * w, HIT_DT)
        st = (C.c_uint64 * 7)()
        self.lib.emul_trace_primary_packet(C.c_void_p(h), _ptr(cam), C.c_int(w), C.c_int(hgt), C.c_int(y0), C.c_int(y1),
                                           C.c_int(tile_w), C.c_int(tile_h), C.c_int(sub_w), C.c_int(sub_h), _ptr(hits), st)
        stats = dict(zip(("rays", "inner_visits", "sphere_tests", "exact_box_tests", "hits", "max_stack", "max_visits"), map(int, st)))
        return hits, stats

    def primary_rays(self, cam, w, h, y0=0, y1=None):
        y1 = h if y1 is None else y1
        out = np.zeros((y1 - y0) * w, RAY_DT)
        self.lib.emul_primary_rays(_ptr(cam), C.c_int(w), C.c_int(h), C.c_int(y0), C.c_int(y1), _ptr(out))
        return out


# The library's "keep the 64-byte binary tree in new handles" switch is process-wide and has no getter; the GPU test
# modules want different values and pytest may interleave them (session-scoped parametrised fixtures reorder tests),
# so every change goes through here and a module can restore what it found.
_KEEP_BINARY = [False]


def set_keep_binary(pkg, keep):
    pkg.set_keep_binary_tree(bool(keep))
    _KEEP_BINARY[0] = bool(keep)


def keep_binary_state():
    return _KEEP_BINARY[0]
