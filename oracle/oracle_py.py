"""ctypes bindings for the CHECKERS -- test infrastructure only.

* ``Oracle``  : oracle/liboracle.so  (our CPU restatement, oracle.c; travels to the GPU box)
* ``RefLib``  : oracle/_ref/libref*.so (the unmodified reference sources compiled by
                oracle/Makefile; built in the container where /root/reference exists, the
                prebuilt .so travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"

# numpy views of the 12/20/24/32/40/64-byte records (layouts pinned by tests/test_layout.py)
RAY_DT = np.dtype([("o", np.float32, 3), ("d", np.float32, 3)])                       # 24 B
SPHERE_DT = np.dtype([("c", np.float32, 3), ("r", np.float32), ("rgba", np.uint8, 4)])  # 20 B
HIT_DT = np.dtype([("t", np.float32), ("prim", np.int32), ("p", np.float32, 3), ("n", np.float32, 3)])  # 32 B
NODE_DT = np.dtype([("lo", np.float32, 3), ("hi", np.float32, 3), ("left", np.int32),
                    ("right", np.int32), ("sphere", np.int32), ("count", np.int32)])   # 40 B
CAMERA_DT = np.dtype([("position", np.float32, 3), ("forward", np.float32, 3), ("right", np.float32, 3),
                      ("up", np.float32, 3), ("yaw", np.float32), ("pitch", np.float32),
                      ("fov", np.float32), ("move", np.int32)])                        # 64 B
assert RAY_DT.itemsize == 24 and SPHERE_DT.itemsize == 20 and HIT_DT.itemsize == 32
assert NODE_DT.itemsize == 40 and CAMERA_DT.itemsize == 64


class Stats(C.Structure):
    _fields_ = [("nodes", C.c_int64), ("leaves", C.c_int64), ("multi_leaves", C.c_int64),
                ("empty_leaves", C.c_int64), ("max_depth", C.c_int)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def build_oracle(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = [os.path.join(HERE, f) for f in ("oracle.c", "oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return so


def build_ref(res: str | None = None):
    """Compile the reference where it lies (only possible where /root/reference exists)."""
    name = "libref.so" if not res else f"libref_{res}.so"
    so = os.path.join(HERE, "_ref", name)
    if os.path.isdir(os.path.join(REF_ROOT, "src")):
        harness = os.path.join(HERE, "ref_harness.c")
        if not os.path.exists(so) or os.path.getmtime(harness) > os.path.getmtime(so):
            cmd = ["make", "-s", "-C", HERE, "ref"] + ([f"RES={res}"] if res else [])
            subprocess.check_call(cmd)
    return so if os.path.exists(so) else None


def ref_available(res: str | None = None) -> bool:
    return build_ref(res) is not None


class Oracle:
    """oracle/liboracle.so"""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_trace_bvh.restype = C.c_double
        L.orc_trace_brute.restype = C.c_double
        L.orc_build.restype = C.c_int64
        L.orc_bounce_rays.restype = C.c_int64

    def srand(self, seed):
        self.lib.orc_srand(C.c_uint(seed))

    def scene_benchmark(self, n, pad=1):
        s = np.zeros(n + pad, SPHERE_DT)
        self.lib.orc_scene_benchmark(C.c_int(n), _ptr(s))
        return s

    def scene_random(self, n, pad=1):
        s = np.zeros(n + pad, SPHERE_DT)
        self.lib.orc_scene_random(C.c_int(n), _ptr(s))
        return s

    def rays_benchmark(self, n):
        r = np.zeros(n, RAY_DT)
        self.lib.orc_rays_benchmark(C.c_int64(n), _ptr(r))
        return r

    def default_camera(self):
        cam = np.zeros(1, CAMERA_DT)
        self.lib.orc_default_camera(_ptr(cam))
        return cam

    def camera_update(self, cam):
        self.lib.orc_camera_update(_ptr(cam))
        return cam

    def primary_rays(self, cam, w, h, y0=0, y1=None):
        y1 = h if y1 is None else y1
        r = np.zeros((y1 - y0) * w, RAY_DT)
        self.lib.orc_primary_rays(_ptr(cam), C.c_int(w), C.c_int(h), C.c_int(y0), C.c_int(y1), _ptr(r))
        return r

    def bounce_rays(self, hits):
        hits = np.ascontiguousarray(hits)
        out = np.zeros(len(hits), RAY_DT)
        m = self.lib.orc_bounce_rays(_ptr(hits), C.c_int64(len(hits)), _ptr(out))
        return out[:m].copy()

    def build(self, spheres, start, end, depth=0):
        """Permutes ``spheres`` in place; returns the preorder node array."""
        cap = max(4, 2 * max(end - start, 1) + 4 + 2 * 64)
        backup = spheres.copy()
        while True:
            nodes = np.zeros(cap, NODE_DT)
            n = self.lib.orc_build(_ptr(spheres), C.c_int(start), C.c_int(end), C.c_int(depth), _ptr(nodes), C.c_int64(cap))
            if n >= 0:
                return nodes[:n].copy()
            spheres[:] = backup      # degenerate trees (Q2/Q3 chains) can exceed 2n nodes: retry larger
            cap *= 4

    def tree_stats(self, nodes, depth0=0):
        st = Stats()
        self.lib.orc_tree_stats(_ptr(nodes), C.c_int(depth0), C.byref(st))
        return st.as_dict()

    def trace_bvh(self, nodes, spheres, n_spheres, rays, threads=0):
        threads = threads or (os.cpu_count() or 1)
        rays = np.ascontiguousarray(rays)
        out = np.zeros(len(rays), HIT_DT)
        sec = self.lib.orc_trace_bvh(_ptr(nodes), _ptr(spheres), C.c_int(n_spheres), _ptr(rays),
                                     C.c_int64(len(rays)), _ptr(out), C.c_int(threads))
        return out, sec

    def trace_brute(self, spheres, n_spheres, rays, threads=0):
        threads = threads or (os.cpu_count() or 1)
        rays = np.ascontiguousarray(rays)
        out = np.zeros(len(rays), HIT_DT)
        sec = self.lib.orc_trace_brute(_ptr(spheres), C.c_int(n_spheres), _ptr(rays),
                                       C.c_int64(len(rays)), _ptr(out), C.c_int(threads))
        return out, sec

    def render(self, nodes, spheres, n_spheres, cam, w, h, y0=0, y1=None, max_depth=5, use_bvh=True, want_samples=True):
        """renderer.c:21-76 over rows [y0,y1): RGBA8 [pixels,4] and the random_in_unit_sphere() draws [pixels,max_depth,3]."""
        y1 = h if y1 is None else y1
        npix = (y1 - y0) * w
        rgba = np.zeros((npix, 4), np.uint8)
        samples = np.zeros((npix, max(max_depth, 1), 3), np.float32) if want_samples else None
        self.lib.orc_render(_ptr(nodes) if nodes is not None else None, _ptr(spheres), C.c_int(n_spheres), _ptr(cam),
                            C.c_int(w), C.c_int(h), C.c_int(y0), C.c_int(y1), C.c_int(max_depth), C.c_int(1 if use_bvh else 0),
                            _ptr(samples) if want_samples else None, _ptr(rgba))
        return rgba, (samples[:, :max_depth] if want_samples else None)

    def accumulate(self, accum, frames, rgba):
        avg = np.zeros_like(rgba)
        self.lib.orc_accumulate(_ptr(accum), C.c_int(frames), _ptr(np.ascontiguousarray(rgba)), C.c_int64(len(rgba)), _ptr(avg))
        return avg

    def walk_counters(self, nodes, spheres, n_spheres, rays, want_hits=False):
        rays = np.ascontiguousarray(rays)
        cnt = (C.c_int64 * 3)()
        hits = np.zeros(len(rays), HIT_DT) if want_hits else None
        self.lib.orc_walk_counters(_ptr(nodes), _ptr(spheres), C.c_int(n_spheres), _ptr(rays),
                                   C.c_int64(len(rays)), cnt, _ptr(hits) if want_hits else None)
        return {"inner": int(cnt[0]), "sphere_tests": int(cnt[1]), "max_stack": int(cnt[2])}, hits


class RefLib:
    """oracle/_ref/libref*.so -- the reference itself (pointer tree, compile-time WIDTH/HEIGHT)."""

    def __init__(self, res: str | None = None):
        so = build_ref(res)
        if so is None:
            raise FileNotFoundError("reference library not built (needs /root/reference or a prebuilt oracle/_ref)")
        self.lib = C.CDLL(so)
        L = self.lib
        L.refh_build.restype = C.c_void_p
        L.refh_trace_bvh.restype = C.c_double
        L.refh_trace_brute.restype = C.c_double
        L.refh_export_tree.restype = C.c_int64
        L.refh_bounce_rays.restype = C.c_int64
        L.refh_epsilon.restype = C.c_float
        L.refh_benchmark_no_bvh.restype = C.c_double
        L.refh_benchmark_with_bvh.restype = C.c_double
        lay = (C.c_int * 24)()
        L.refh_layout(lay)
        self.layout = list(lay)
        self.width, self.height = self.layout[17], self.layout[18]

    def srand(self, seed):
        self.lib.refh_srand(C.c_uint(seed))

    def scene_benchmark(self, n, pad=1):
        s = np.zeros(n + pad, SPHERE_DT)
        self.lib.refh_scene_benchmark(C.c_int(n), _ptr(s))
        return s

    def scene_random(self, n, pad=1):
        s = np.zeros(n + pad, SPHERE_DT)
        self.lib.refh_scene_random(C.c_int(n), _ptr(s))
        return s

    def rays_benchmark(self, n):
        r = np.zeros(n, RAY_DT)
        self.lib.refh_rays_benchmark(C.c_int64(n), _ptr(r))
        return r

    def default_camera(self):
        cam = np.zeros(1, CAMERA_DT)
        self.lib.refh_default_camera(_ptr(cam))
        return cam

    def camera_update(self, cam):
        self.lib.camera_update(_ptr(cam))
        return cam

    def primary_rays(self, cam):
        r = np.zeros(self.width * self.height, RAY_DT)
        self.lib.refh_primary_rays(_ptr(cam), _ptr(r))
        return r

    def bounce_rays(self, hits):
        hits = np.ascontiguousarray(hits)
        out = np.zeros(len(hits), RAY_DT)
        m = self.lib.refh_bounce_rays(_ptr(hits), C.c_int64(len(hits)), _ptr(out))
        return out[:m].copy()

    def build(self, spheres, start, end, depth=0):
        sec = C.c_double()
        root = self.lib.refh_build(_ptr(spheres), C.c_int(start), C.c_int(end), C.c_int(depth), C.byref(sec))
        return root, sec.value

    def free(self, root):
        self.lib.refh_free(C.c_void_p(root))

    def tree_stats(self, root, depth0=0):
        st = Stats()
        self.lib.refh_tree_stats(C.c_void_p(root), C.c_int(depth0), C.byref(st))
        return st.as_dict()

    def export_tree(self, root, spheres, n_nodes):
        nodes = np.zeros(n_nodes, NODE_DT)
        n = self.lib.refh_export_tree(C.c_void_p(root), _ptr(spheres), _ptr(nodes))
        assert n == n_nodes
        return nodes

    def trace_bvh(self, root, spheres, rays, threads=0):
        threads = threads or (os.cpu_count() or 1)
        rays = np.ascontiguousarray(rays)
        out = np.zeros(len(rays), HIT_DT)
        sec = self.lib.refh_trace_bvh(C.c_void_p(root), _ptr(spheres), _ptr(rays), C.c_int64(len(rays)),
                                      _ptr(out), C.c_int(threads))
        return out, sec

    def render(self, cam, spheres, n_spheres, root):
        """The reference's own main.c:356-366 loop over trace_ray at this build's WIDTH x HEIGHT; root=None -> brute force."""
        w, h = self.layout[17], self.layout[18]
        rgba = np.zeros((w * h, 4), np.uint8)
        self.lib.refh_render(_ptr(cam), _ptr(spheres), C.c_int(n_spheres), C.c_void_p(root) if root else None, _ptr(rgba))
        return rgba

    def accumulate(self, accum, frames, rgba, first):
        """main.c:368-405; frames is a 1-element int32 array updated in place."""
        shown = np.zeros_like(rgba)
        self.lib.refh_accumulate(_ptr(accum), _ptr(frames), _ptr(np.ascontiguousarray(rgba)), C.c_int(1 if first else 0), _ptr(shown))
        return shown

    def trace_brute(self, spheres, n_spheres, rays, threads=0):
        threads = threads or (os.cpu_count() or 1)
        rays = np.ascontiguousarray(rays)
        out = np.zeros(len(rays), HIT_DT)
        sec = self.lib.refh_trace_brute(_ptr(spheres), C.c_int(n_spheres), _ptr(rays), C.c_int64(len(rays)),
                                        _ptr(out), C.c_int(threads))
        return out, sec
