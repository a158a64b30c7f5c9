"""ctypes loader for libsahbvh_b200.so -- the B200 closest-hit path of the SAH-BVH ray tracer.

The product is the C/CUDA shared library (``csrc/``, C-ABI in ``include/sahbvh_gpu.h`` and the
reference-named functions in ``include/sahbvh_ref_api.h``).  This module only binds it for the
Python test / bench harness: numpy record dtypes that match the reference's struct layouts and
thin wrappers with the same names and argument meaning as the C entry points.

There is no CPU traversal here and no fallback: if the shared library is missing the import
fails, and if no CUDA device is usable every trace call raises ``SahBvhError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SAHBVH_LIB") or os.path.join(_HERE, "libsahbvh_b200.so")  # override: kernel-tuning builds only

RAY_DT = np.dtype([("o", np.float32, 3), ("d", np.float32, 3)])                         # Ray, ray.h:5-8
SPHERE_DT = np.dtype([("c", np.float32, 3), ("r", np.float32), ("rgba", np.uint8, 4)])   # Sphere, sphere.h:7-11
HIT_DT = np.dtype([("t", np.float32), ("prim", np.int32), ("p", np.float32, 3), ("n", np.float32, 3)])  # gpu_hit
CAMERA_DT = np.dtype([("position", np.float32, 3), ("forward", np.float32, 3), ("right", np.float32, 3),
                      ("up", np.float32, 3), ("yaw", np.float32), ("pitch", np.float32),
                      ("fov", np.float32), ("move", np.int32)])                          # Camera, camera.h:5-14

EXPORTED_SYMBOLS = [
    # include/sahbvh_gpu.h
    "sahbvh_last_error", "sahbvh_device_count", "sahbvh_version", "sahbvh_kernel_launches",
    "gpu_bvh_upload", "gpu_bvh_build", "gpu_bvh_free", "gpu_bvh_get_info",
    "gpu_trace_closest_hit", "gpu_trace_closest_hit_dev", "gpu_trace_primary", "gpu_trace_primary_dev",
    "gpu_generate_primary_rays", "gpu_trace_brute_force", "gpu_trace_closest_hit_stats", "gpu_trace_primary_stats",
    "sahbvh_measure_peaks", "sahbvh_set_primary_mode", "gpu_trace_primary_dev_ids", "sahbvh_set_ray_mode",
    "gpu_benchmark_no_bvh", "gpu_benchmark_with_bvh", "sahbvh_set_build_mode", "sahbvh_build_tree", "gpu_bvh_download",
    "sahbvh_peer_alloc", "sahbvh_peer_open", "sahbvh_peer_close", "sahbvh_peer_free", "sahbvh_peer_signal", "sahbvh_peer_read_flags",
    "gpu_trace_primary_dev_prims", "gpu_hits_from_prims_dev", "gpu_render_frame", "gpu_render_frame_dev", "gpu_accumulate", "gpu_accumulate_dev", "sahbvh_write_ppm",
    "sahbvh_host_alloc", "sahbvh_host_free", "sahbvh_host_pin", "sahbvh_host_unpin",
    "gpu_trace_closest_hit_prims", "gpu_trace_closest_hit_dev_prims", "gpu_trace_primary_prims", "gpu_trace_primary_tprim",
    "gpu_hits_from_prims", "gpu_primary_kernel_in_use", "sahbvh_set_keep_binary_tree", "sahbvh_set_l2_persist_bytes", "sahbvh_peer_wait", "gpu_count_prims_dev",
    "gpu_bvh_replicate", "gpu_bvh_build_replicated", "gpu_bvh_set_free", "gpu_bvh_set_size", "gpu_bvh_set_get",
    "sahbvh_shard_range", "sahbvh_shard_rows", "sahbvh_shard_rows_weighted", "gpu_trace_primary_multi", "gpu_trace_primary_prims_multi",
    "gpu_trace_primary_prims_multi_dev", "gpu_trace_closest_hit_multi", "gpu_trace_closest_hit_prims_multi",
    "gpu_benchmark_with_bvh_multi", "gpu_render_frame_multi",
    # include/sahbvh_ref_api.h
    "vec3_sub", "vec3_normalize", "vec3_dot", "vec3_add", "vec3_multiply", "vec3_cross", "vec3_random",
    "create_empty_aabb", "create_aabb_from_sphere", "combine_aabb", "get_aabb_surface_area", "evaluate_sah",
    "build_bvh_node", "free_bvh", "ray_sphere_intersect", "ray_aabb_intersect", "ray_bvh_intersect",
    "get_camera_ray", "get_camera_ray_wh", "camera_update", "sahbvh_pixel_ray", "sahbvh_hit_from_prim", "sahbvh_scalar_forget",
    "random_float", "random_in_unit_sphere", "random_on_hemisphere",
    "create_benchmark_sphere", "create_sphere", "create_random_sphere",
    "sahbvh_fill_benchmark_scene", "sahbvh_fill_random_scene", "sahbvh_fill_benchmark_rays",
    "sahbvh_default_camera", "sahbvh_fill_bounce_rays",
]


class SahBvhError(RuntimeError):
    pass


class BvhInfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_spheres", C.c_int32), ("n_inner", C.c_int32),
                ("n_leaf_refs", C.c_int32), ("max_depth", C.c_int32), ("multi_leaves", C.c_int32),
                ("empty_leaves", C.c_int32), ("n_wide", C.c_int32), ("max_radius", C.c_float),
                ("scene_lo", C.c_float * 3), ("scene_hi", C.c_float * 3), ("device_bytes", C.c_int64)]

    def as_dict(self):
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d


class TraceStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("inner_visits", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("exact_box_tests", C.c_uint64), ("hits", C.c_uint64), ("max_stack", C.c_uint64),
                ("max_visits", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C cs201_sah-bvh_ray_tracer_b200`). There is no Python/CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.sahbvh_last_error.restype = C.c_char_p
    lib.sahbvh_version.restype = C.c_char_p
    lib.sahbvh_kernel_launches.restype = C.c_uint64
    lib.build_bvh_node.restype = C.c_void_p
    lib.free_bvh.restype = None
    lib.gpu_bvh_free.restype = None
    lib.gpu_bvh_set_free.restype = None
    lib.gpu_bvh_set_get.restype = C.c_void_p
    lib.sahbvh_shard_rows.restype = None
    lib.sahbvh_shard_rows_weighted.restype = None
    lib.sahbvh_shard_range.restype = None
    return lib


lib = _load()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc, what):
    if rc != 0:
        raise SahBvhError(f"{what} failed ({rc}): {lib.sahbvh_last_error().decode()}")


def device_count() -> int:
    return int(lib.sahbvh_device_count())


def kernel_launches() -> int:
    return int(lib.sahbvh_kernel_launches())


class Peaks(C.Structure):
    _fields_ = [("fp32_mix_inst_per_s", C.c_double), ("ffma_inst_per_s", C.c_double),
                ("l2_read_gbs", C.c_double), ("hbm_read_gbs", C.c_double), ("sm_count", C.c_double)]


def measure_peaks(device: int = -1) -> dict:
    p = Peaks()
    _check(lib.sahbvh_measure_peaks(C.c_int(device), C.byref(p)), "sahbvh_measure_peaks")
    return {k: float(getattr(p, k)) for k, _ in p._fields_}


PRIMARY_PER_RAY, PRIMARY_PACKET, PRIMARY_MACRO, PRIMARY_AUTO = 0, 1, 2, 3
PRIMARY_DEFAULT = PRIMARY_AUTO


RAYS_ONE_SHOT, RAYS_COOP, RAYS_WIDE = 0, 1, 2


def set_ray_mode(mode: int):
    _check(lib.sahbvh_set_ray_mode(C.c_int(mode)), "sahbvh_set_ray_mode")


BUILD_HOST, BUILD_DEVICE = 0, 1


def set_build_mode(mode: int):
    _check(lib.sahbvh_set_build_mode(C.c_int(mode)), "sahbvh_set_build_mode")


TREE_NODE_DT = np.dtype([("lo", np.float32, 3), ("hi", np.float32, 3), ("left", np.int32), ("right", np.int32),
                         ("sphere", np.int32), ("count", np.int32)])    # sahbvh_tree_node


def build_tree(spheres: np.ndarray, start: int, end: int, depth: int = 0, mode: int = BUILD_DEVICE, device: int = -1) -> np.ndarray:
    """sahbvh_build_tree: preorder array tree; permutes ``spheres`` in place like build_bvh_node."""
    lib.sahbvh_build_tree.restype = C.c_int64
    cap = 2 * max(end - start, 1) + 1024      # exact for trees without empty leaves; retried below otherwise
    backup = spheres.copy()
    for _ in range(2):
        out = np.empty(cap, TREE_NODE_DT)
        n = lib.sahbvh_build_tree(_ptr(spheres), C.c_int(start), C.c_int(end), C.c_int(depth), C.c_int(mode), C.c_int(device),
                                  _ptr(out), C.c_int64(cap))
        if n < 0:
            raise SahBvhError(f"sahbvh_build_tree failed ({n}): {lib.sahbvh_last_error().decode()}")
        if n <= cap:
            return out[:n].copy()
        cap = int(n)
        spheres[:] = backup                    # the build permuted the array; start again from the input order
    raise SahBvhError("sahbvh_build_tree: node count changed between two runs")


def set_keep_binary_tree(keep: bool):
    """Handles created afterwards keep (or drop) the 64-byte binary tree the PER_RAY / COOP / ONE_SHOT kernels read."""
    _check(lib.sahbvh_set_keep_binary_tree(C.c_int(1 if keep else 0)), "sahbvh_set_keep_binary_tree")


TPRIM_DT = np.dtype([("t", np.float32), ("prim", np.int32)])    # gpu_tprim


def set_l2_persist_bytes(nbytes: int):
    """persisting-L2 window over the top `nbytes` of the wide tree for every trace launch (0 = off)"""
    _check(lib.sahbvh_set_l2_persist_bytes(C.c_uint64(nbytes)), "sahbvh_set_l2_persist_bytes")


def set_primary_mode(mode: int):
    _check(lib.sahbvh_set_primary_mode(C.c_int(mode)), "sahbvh_set_primary_mode")


def default_camera() -> np.ndarray:
    """main.c:203-211"""
    cam = np.zeros(1, CAMERA_DT)
    cam["position"] = (0, 4, 50)
    cam["forward"] = (0, 0, -1)
    cam["right"] = (1, 0, 0)
    cam["up"] = (0, 1, 0)
    cam["yaw"] = np.float32(-np.pi)
    cam["fov"] = 45.0
    return cam


def srand(seed: int):
    C.CDLL(None).srand(C.c_uint(seed))


def benchmark_scene(n: int) -> np.ndarray:
    """benchmark.c:306-314 through the library's create_benchmark_sphere (glibc rand() stream)."""
    s = np.zeros(n + 1, SPHERE_DT)
    lib.sahbvh_fill_benchmark_scene(C.c_int(n), _ptr(s))
    return s


def random_scene(n: int) -> np.ndarray:
    """main.c:218-221 through the library's create_random_sphere."""
    s = np.zeros(n + 1, SPHERE_DT)
    lib.sahbvh_fill_random_scene(C.c_int(n), _ptr(s))
    return s


def benchmark_rays(n: int) -> np.ndarray:
    """benchmark.c:232-240"""
    r = np.zeros(n, RAY_DT)
    lib.sahbvh_fill_benchmark_rays(C.c_int64(n), _ptr(r))
    return r


def bounce_rays(hits: np.ndarray) -> np.ndarray:
    """renderer.c:51-54"""
    hits = np.ascontiguousarray(hits)
    out = np.zeros(len(hits), RAY_DT)
    lib.sahbvh_fill_bounce_rays.restype = C.c_int64
    m = lib.sahbvh_fill_bounce_rays(_ptr(hits), C.c_int64(len(hits)), _ptr(out))
    return out[:m].copy()


def build_bvh_node(spheres: np.ndarray, start: int, end: int, depth: int = 0) -> int:
    """bvh.c:117 signature; returns the BVHNode* as an integer. Permutes ``spheres`` in place."""
    root = lib.build_bvh_node(_ptr(spheres), C.c_int(start), C.c_int(end), C.c_int(depth))
    if not root:
        raise SahBvhError("build_bvh_node returned NULL")
    return root


def free_bvh(root: int):
    lib.free_bvh(C.c_void_p(root))


class GpuBvh:
    """Owner of one ``gpu_bvh*`` (flattened tree + SoA spheres on one device)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def upload(cls, spheres: np.ndarray, n: int, root: int, device: int = -1) -> "GpuBvh":
        h = C.c_void_p()
        _check(lib.gpu_bvh_upload(_ptr(spheres), C.c_int(n), C.c_void_p(root), C.c_int(device), C.byref(h)),
               "gpu_bvh_upload")
        return cls(h)

    @classmethod
    def build(cls, spheres: np.ndarray, n_total: int, start: int, end: int, depth: int = 0, device: int = -1) -> "GpuBvh":
        h = C.c_void_p()
        _check(lib.gpu_bvh_build(_ptr(spheres), C.c_int(n_total), C.c_int(start), C.c_int(end), C.c_int(depth),
                                 C.c_int(device), C.byref(h)), "gpu_bvh_build")
        return cls(h)

    def close(self):
        if self._h:
            lib.gpu_bvh_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def download(self, what: int) -> np.ndarray:
        """gpu_bvh_download: raw bytes of one device array (0 nodes, 1 spheres4, 2 colours, 3 wide, 4 wide full, 5 special)."""
        lib.gpu_bvh_download.restype = C.c_int64
        n = lib.gpu_bvh_download(self._h, C.c_int(what), None, C.c_int64(0))
        if n < 0:
            raise SahBvhError(f"gpu_bvh_download failed ({n}): {lib.sahbvh_last_error().decode()}")
        out = np.empty(n, np.uint8)
        if n:
            lib.gpu_bvh_download(self._h, C.c_int(what), _ptr(out), C.c_int64(n))
        return out

    def info(self) -> dict:
        i = BvhInfo()
        _check(lib.gpu_bvh_get_info(self._h, C.byref(i)), "gpu_bvh_get_info")
        return i.as_dict()

    # ---- host-buffer entry points (H2D + kernel + D2H inside the call) ----
    def trace_closest_hit(self, rays: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DT
        if out is None:
            out = np.empty(len(rays), HIT_DT)
        _check(lib.gpu_trace_closest_hit(self._h, _ptr(rays), C.c_int64(len(rays)), _ptr(out)), "gpu_trace_closest_hit")
        return out

    def trace_closest_hit_ptr(self, rays_ptr: int, n: int, out_ptr: int):
        _check(lib.gpu_trace_closest_hit(self._h, C.c_void_p(rays_ptr), C.c_int64(n), C.c_void_p(out_ptr)),
               "gpu_trace_closest_hit")

    def trace_primary(self, cam: np.ndarray, width: int, height: int, y0: int = 0, y1: int | None = None,
                      out: np.ndarray | None = None) -> np.ndarray:
        y1 = height if y1 is None else y1
        if out is None:
            out = np.empty((y1 - y0) * width, HIT_DT)
        _check(lib.gpu_trace_primary(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                     _ptr(out)), "gpu_trace_primary")
        return out

    def trace_primary_ptr(self, cam: np.ndarray, width: int, height: int, y0: int, y1: int, out_ptr: int):
        _check(lib.gpu_trace_primary(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                     C.c_void_p(out_ptr)), "gpu_trace_primary")

    def trace_primary_prims(self, cam: np.ndarray, width: int, height: int, y0: int = 0, y1: int | None = None,
                            out: np.ndarray | None = None) -> np.ndarray:
        """gpu_trace_primary_prims: 4-byte primitive id per pixel into a host buffer (-1 = miss)."""
        y1 = height if y1 is None else y1
        if out is None:
            out = np.empty((y1 - y0) * width, np.int32)
        _check(lib.gpu_trace_primary_prims(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                           _ptr(out)), "gpu_trace_primary_prims")
        return out

    def trace_primary_prims_ptr(self, cam: np.ndarray, width: int, height: int, y0: int, y1: int, out_ptr: int):
        _check(lib.gpu_trace_primary_prims(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                           C.c_void_p(out_ptr)), "gpu_trace_primary_prims")

    def trace_primary_tprim(self, cam: np.ndarray, width: int, height: int, y0: int = 0, y1: int | None = None,
                            out: np.ndarray | None = None) -> np.ndarray:
        """gpu_trace_primary_tprim: {t, prim} per pixel (8 bytes) into a host buffer."""
        y1 = height if y1 is None else y1
        if out is None:
            out = np.empty((y1 - y0) * width, TPRIM_DT)
        _check(lib.gpu_trace_primary_tprim(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                           _ptr(out)), "gpu_trace_primary_tprim")
        return out

    def trace_primary_tprim_ptr(self, cam: np.ndarray, width: int, height: int, y0: int, y1: int, out_ptr: int):
        _check(lib.gpu_trace_primary_tprim(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                           C.c_void_p(out_ptr)), "gpu_trace_primary_tprim")

    def hits_from_prims(self, cam: np.ndarray, width: int, height: int, prims: np.ndarray, y0: int = 0,
                        y1: int | None = None) -> np.ndarray:
        """gpu_hits_from_prims: the full records of an id frame, rebuilt on demand (host buffers)."""
        y1 = height if y1 is None else y1
        prims = np.ascontiguousarray(prims, np.int32)
        assert len(prims) == (y1 - y0) * width
        out = np.empty(len(prims), HIT_DT)
        _check(lib.gpu_hits_from_prims(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                       _ptr(prims), _ptr(out)), "gpu_hits_from_prims")
        return out

    def trace_closest_hit_prims(self, rays: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """gpu_trace_closest_hit_prims: 4-byte primitive id per ray (-1 = miss)."""
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DT
        if out is None:
            out = np.empty(len(rays), np.int32)
        _check(lib.gpu_trace_closest_hit_prims(self._h, _ptr(rays), C.c_int64(len(rays)), _ptr(out)), "gpu_trace_closest_hit_prims")
        return out

    def trace_closest_hit_prims_ptr(self, rays_ptr: int, n: int, out_ptr: int):
        _check(lib.gpu_trace_closest_hit_prims(self._h, C.c_void_p(rays_ptr), C.c_int64(n), C.c_void_p(out_ptr)),
               "gpu_trace_closest_hit_prims")

    def trace_closest_hit_dev_prims(self, d_rays: int, n: int, d_prims: int, stream: int = 0):
        _check(lib.gpu_trace_closest_hit_dev_prims(self._h, C.c_void_p(d_rays), C.c_int64(n), C.c_void_p(d_prims),
                                                   C.c_void_p(stream)), "gpu_trace_closest_hit_dev_prims")

    def benchmark_with_bvh_ptr(self, rays_ptr: int, n: int) -> int:
        out = C.c_int64(0)
        _check(lib.gpu_benchmark_with_bvh(self._h, C.c_void_p(rays_ptr), C.c_int64(n), C.byref(out)), "gpu_benchmark_with_bvh")
        return int(out.value)

    def render_frame_ptr(self, cam: np.ndarray, width: int, height: int, y0: int, y1: int, max_depth: int, seed: int, out_ptr: int):
        _check(lib.gpu_render_frame(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                    C.c_int(max_depth), C.c_uint64(seed), None, C.c_void_p(out_ptr)), "gpu_render_frame")

    def trace_brute_force(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays)
        out = np.empty(len(rays), HIT_DT)
        _check(lib.gpu_trace_brute_force(self._h, _ptr(rays), C.c_int64(len(rays)), _ptr(out)), "gpu_trace_brute_force")
        return out

    def benchmark_no_bvh(self, rays: np.ndarray) -> int:
        """benchmark.c:172-223: number of rays that hit any sphere (every sphere tested, no BVH)."""
        rays = np.ascontiguousarray(rays)
        out = C.c_int64(0)
        _check(lib.gpu_benchmark_no_bvh(self._h, _ptr(rays), C.c_int64(len(rays)), C.byref(out)), "gpu_benchmark_no_bvh")
        return int(out.value)

    def benchmark_with_bvh(self, rays: np.ndarray) -> int:
        """benchmark.c:225-254: number of rays whose closest hit exists."""
        rays = np.ascontiguousarray(rays)
        out = C.c_int64(0)
        _check(lib.gpu_benchmark_with_bvh(self._h, _ptr(rays), C.c_int64(len(rays)), C.byref(out)), "gpu_benchmark_with_bvh")
        return int(out.value)

    def render_frame(self, cam: np.ndarray, width: int, height: int, y0: int = 0, y1: int | None = None, max_depth: int = 5,
                     seed: int = 0, samples: np.ndarray | None = None) -> np.ndarray:
        """gpu_render_frame: RGBA8 rows [y0,y1); ``samples`` = float32 [pixels, max_depth, 3] (parity mode) or None."""
        y1 = height if y1 is None else y1
        out = np.empty(((y1 - y0) * width, 4), np.uint8)
        sp = None
        if samples is not None:
            sp = np.ascontiguousarray(samples, np.float32)
            assert sp.shape == ((y1 - y0) * width, max_depth, 3)
        _check(lib.gpu_render_frame(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                    C.c_int(max_depth), C.c_uint64(seed), _ptr(sp) if sp is not None else None, _ptr(out)),
               "gpu_render_frame")
        return out

    def render_frame_dev(self, cam: np.ndarray, width: int, height: int, y0: int, y1: int, max_depth: int, seed: int,
                         d_rgba: int, stream: int = 0):
        _check(lib.gpu_render_frame_dev(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                        C.c_int(max_depth), C.c_uint64(seed), None, C.c_void_p(d_rgba), C.c_void_p(stream)),
               "gpu_render_frame_dev")

    # ---- device-pointer entry points (asynchronous on `stream`) ----
    def trace_closest_hit_dev(self, d_rays: int, n: int, d_out: int, stream: int = 0):
        _check(lib.gpu_trace_closest_hit_dev(self._h, C.c_void_p(d_rays), C.c_int64(n), C.c_void_p(d_out),
                                             C.c_void_p(stream)), "gpu_trace_closest_hit_dev")

    def trace_primary_dev(self, cam: np.ndarray, width: int, height: int, y0: int, y1: int, d_out: int, stream: int = 0):
        _check(lib.gpu_trace_primary_dev(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                         C.c_void_p(d_out), C.c_void_p(stream)), "gpu_trace_primary_dev")

    def trace_primary_dev_ids(self, cam: np.ndarray, width: int, height: int, y0: int, y1: int, d_out: int, d_ids: int,
                              stream: int = 0):
        """Full 32-byte records to d_out and the compact {t, prim} ID buffer to d_ids (the all-gather slot)."""
        _check(lib.gpu_trace_primary_dev_ids(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                             C.c_void_p(d_out), C.c_void_p(d_ids), C.c_void_p(stream)),
               "gpu_trace_primary_dev_ids")

    def trace_primary_dev_prims(self, cam: np.ndarray, width: int, height: int, y0: int, y1: int, d_out: int, d_prims: int,
                                stream: int = 0):
        """Full records to d_out (0 = none: ids only) and the 4-byte primitive-id buffer to d_prims."""
        _check(lib.gpu_trace_primary_dev_prims(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                               C.c_void_p(d_out) if d_out else None, C.c_void_p(d_prims), C.c_void_p(stream)),
               "gpu_trace_primary_dev_prims")

    def hits_from_prims_dev(self, cam: np.ndarray, width: int, height: int, y0: int, y1: int, d_prims: int, d_out: int,
                            stream: int = 0):
        _check(lib.gpu_hits_from_prims_dev(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                           C.c_void_p(d_prims), C.c_void_p(d_out), C.c_void_p(stream)),
               "gpu_hits_from_prims_dev")

    def primary_kernel_in_use(self, width: int, rows: int) -> int:
        """PRIMARY_* id of the kernel this handle uses for frames of this shape; -1 while PRIMARY_AUTO is sampling"""
        return int(lib.gpu_primary_kernel_in_use(self._h, C.c_int(width), C.c_int(rows)))

    def trace_closest_hit_stats(self, d_rays: int, n: int, d_out: int) -> dict:
        st = TraceStats()
        _check(lib.gpu_trace_closest_hit_stats(self._h, C.c_void_p(d_rays), C.c_int64(n), C.c_void_p(d_out), C.byref(st)),
               "gpu_trace_closest_hit_stats")
        return st.as_dict()

    def trace_primary_stats(self, cam: np.ndarray, width: int, height: int, y0: int, y1: int, d_out: int) -> dict:
        st = TraceStats()
        _check(lib.gpu_trace_primary_stats(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                           C.c_void_p(d_out), C.byref(st)), "gpu_trace_primary_stats")
        return st.as_dict()


def host_pin(arr) -> None:
    """sahbvh_host_pin: page-lock a numpy array in place so the host-pointer entry points copy at PCIe speed."""
    _check(lib.sahbvh_host_pin(C.c_void_p(arr.ctypes.data), C.c_size_t(arr.nbytes)), "sahbvh_host_pin")


def host_unpin(arr) -> None:
    _check(lib.sahbvh_host_unpin(C.c_void_p(arr.ctypes.data)), "sahbvh_host_unpin")


def peer_alloc(nbytes: int, device: int = -1) -> tuple[int, bytes]:
    """sahbvh_peer_alloc: (device pointer, 64-byte IPC handle) of a buffer other processes' kernels can store into."""
    p = C.c_void_p()
    h = (C.c_ubyte * 64)()
    _check(lib.sahbvh_peer_alloc(C.c_int(device), C.c_size_t(nbytes), C.byref(p), h), "sahbvh_peer_alloc")
    return int(p.value), bytes(h)


def peer_open(handle: bytes, device: int = -1) -> int:
    p = C.c_void_p()
    h = (C.c_ubyte * 64).from_buffer_copy(handle)
    _check(lib.sahbvh_peer_open(C.c_int(device), h, C.byref(p)), "sahbvh_peer_open")
    return int(p.value)


def peer_signal(flag_ptr: int, value: int, stream: int = 0):
    _check(lib.sahbvh_peer_signal(C.c_void_p(flag_ptr), C.c_uint64(value), C.c_void_p(stream)), "sahbvh_peer_signal")


def peer_read_flags(flags_ptr: int, n: int, stream: int = 0) -> list[int]:
    out = (C.c_uint64 * n)()
    _check(lib.sahbvh_peer_read_flags(C.c_void_p(flags_ptr), C.c_int(n), out, C.c_void_p(stream)), "sahbvh_peer_read_flags")
    return [int(v) for v in out]


def peer_close(ptr: int, device: int = -1):
    _check(lib.sahbvh_peer_close(C.c_int(device), C.c_void_p(ptr)), "sahbvh_peer_close")


def peer_free(ptr: int, device: int = -1):
    _check(lib.sahbvh_peer_free(C.c_int(device), C.c_void_p(ptr)), "sahbvh_peer_free")


def accumulate(accum: np.ndarray, frames: int, rgba: np.ndarray) -> np.ndarray:
    """gpu_accumulate (main.c:368-405): updates ``accum`` (float32 [pixels,3]) in place, returns the shown RGBA8 frame."""
    assert accum.dtype == np.float32 and accum.flags.c_contiguous
    rgba = np.ascontiguousarray(rgba, np.uint8)
    avg = np.empty_like(rgba)
    _check(lib.gpu_accumulate(_ptr(accum), C.c_int(frames), _ptr(rgba), C.c_int64(len(rgba)), _ptr(avg)), "gpu_accumulate")
    return avg


def write_ppm(path: str, rgba: np.ndarray, width: int, height: int):
    _check(lib.sahbvh_write_ppm(path.encode(), _ptr(np.ascontiguousarray(rgba, np.uint8)), C.c_int(width), C.c_int(height)),
           "sahbvh_write_ppm")


def generate_primary_rays(cam: np.ndarray, width: int, height: int, y0: int = 0, y1: int | None = None,
                          device: int = -1) -> np.ndarray:
    y1 = height if y1 is None else y1
    out = np.empty((y1 - y0) * width, RAY_DT)
    _check(lib.gpu_generate_primary_rays(_ptr(cam), C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                         _ptr(out), C.c_int(device)), "gpu_generate_primary_rays")
    return out


# ---- ray / row sharding: contiguous, no data-path collective; the rule lives in the C library ----
def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous ray range [lo, hi) of part ``rank`` out of ``world`` (sizes differ by at most 1)."""
    lo, hi = C.c_int64(), C.c_int64()
    lib.sahbvh_shard_range(C.c_int64(n), C.c_int(rank), C.c_int(world), C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


def shard_rows_weighted(height: int, rank: int, world: int, root_share: float) -> tuple[int, int]:
    """row band of `rank` when rank 0 takes root_share of an equal share (it also consumes the frame)"""
    y0, y1 = C.c_int(), C.c_int()
    lib.sahbvh_shard_rows_weighted(C.c_int(height), C.c_int(rank), C.c_int(world), C.c_double(root_share), C.byref(y0), C.byref(y1))
    return y0.value, y1.value


def shard_rows(height: int, rank: int, world: int) -> tuple[int, int]:
    """Row band [y0, y1) of a frame: whole 16-row blocks (the packet kernel's CTA group), so no block straddles two parts."""
    y0, y1 = C.c_int(), C.c_int()
    lib.sahbvh_shard_rows(C.c_int(height), C.c_int(rank), C.c_int(world), C.byref(y0), C.byref(y1))
    return int(y0.value), int(y1.value)


def peer_wait(flags_ptr: int, n: int, value: int, timeout_ms: int = 0, status_ptr: int = 0, stream: int = 0):
    _check(lib.sahbvh_peer_wait(C.c_void_p(flags_ptr), C.c_int(n), C.c_uint64(value), C.c_uint32(timeout_ms),
                                C.c_void_p(status_ptr) if status_ptr else None, C.c_void_p(stream)), "sahbvh_peer_wait")


def count_prims_dev(d_prims: int, n: int, d_count: int, stream: int = 0):
    _check(lib.gpu_count_prims_dev(C.c_void_p(d_prims), C.c_int64(n), C.c_void_p(d_count), C.c_void_p(stream)), "gpu_count_prims_dev")


HITRECORD_DT = np.dtype([("t", np.float32), ("point", np.float32, 3), ("normal", np.float32, 3), ("hit_something", np.int32),
                         ("object", np.uint64)])    # HitRecord, hit.h:8-14 (40 bytes)


class _Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class _Ray(C.Structure):
    _fields_ = [("o", _Vec3), ("d", _Vec3)]


class _HitRecord(C.Structure):
    _fields_ = [("t", C.c_float), ("p", _Vec3), ("n", _Vec3), ("hit", C.c_int), ("obj", C.c_void_p)]


def hit_from_prim(cam: np.ndarray, width: int, height: int, x: int, y: int, spheres: np.ndarray, prim: int):
    """sahbvh_hit_from_prim(sahbvh_pixel_ray(cam, W, H, x, y), spheres, prim): the HitRecord of one pixel, derived on the host."""
    lib.sahbvh_pixel_ray.restype = _Ray
    lib.sahbvh_pixel_ray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.sahbvh_hit_from_prim.restype = _HitRecord
    lib.sahbvh_hit_from_prim.argtypes = [_Ray, C.c_void_p, C.c_int32]
    ray = lib.sahbvh_pixel_ray(cam.ctypes.data, width, height, x, y)
    return lib.sahbvh_hit_from_prim(ray, spheres.ctypes.data, prim)


class GpuBvhSet:
    """Owner of one ``gpu_bvh_set*``: N replicas of a tree in ONE process (gpu_bvh_replicate)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def replicate(cls, bvh: GpuBvh, devices: list[int]) -> "GpuBvhSet":
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        _check(lib.gpu_bvh_replicate(bvh._h, arr, C.c_int(len(devices)), C.byref(h)), "gpu_bvh_replicate")
        return cls(h)

    @classmethod
    def build(cls, spheres: np.ndarray, n_total: int, start: int, end: int, depth: int, devices: list[int]) -> "GpuBvhSet":
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        _check(lib.gpu_bvh_build_replicated(_ptr(spheres), C.c_int(n_total), C.c_int(start), C.c_int(end), C.c_int(depth),
                                            arr, C.c_int(len(devices)), C.byref(h)), "gpu_bvh_build_replicated")
        return cls(h)

    def close(self):
        if self._h:
            lib.gpu_bvh_set_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self) -> int:
        return int(lib.gpu_bvh_set_size(self._h))

    def member_info(self, i: int) -> dict:
        info = BvhInfo()
        _check(lib.gpu_bvh_get_info(C.c_void_p(lib.gpu_bvh_set_get(self._h, C.c_int(i))), C.byref(info)), "gpu_bvh_get_info")
        return info.as_dict()

    def trace_primary(self, cam, width, height, out=None):
        if out is None:
            out = np.empty(width * height, HIT_DT)
        _check(lib.gpu_trace_primary_multi(self._h, _ptr(cam), C.c_int(width), C.c_int(height), _ptr(out)), "gpu_trace_primary_multi")
        return out

    def trace_primary_prims(self, cam, width, height, out=None):
        if out is None:
            out = np.empty(width * height, np.int32)
        _check(lib.gpu_trace_primary_prims_multi(self._h, _ptr(cam), C.c_int(width), C.c_int(height), _ptr(out)),
               "gpu_trace_primary_prims_multi")
        return out

    def trace_primary_prims_ptr(self, cam, width, height, out_ptr: int):
        _check(lib.gpu_trace_primary_prims_multi(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_void_p(out_ptr)),
               "gpu_trace_primary_prims_multi")

    def trace_primary_prims_dev(self, cam, width, height, d_prims: int, root_stream: int = 0):
        _check(lib.gpu_trace_primary_prims_multi_dev(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_void_p(d_prims),
                                                     C.c_void_p(root_stream)), "gpu_trace_primary_prims_multi_dev")

    def trace_closest_hit(self, rays, out=None):
        rays = np.ascontiguousarray(rays)
        if out is None:
            out = np.empty(len(rays), HIT_DT)
        _check(lib.gpu_trace_closest_hit_multi(self._h, _ptr(rays), C.c_int64(len(rays)), _ptr(out)), "gpu_trace_closest_hit_multi")
        return out

    def trace_closest_hit_prims(self, rays, out=None):
        rays = np.ascontiguousarray(rays)
        if out is None:
            out = np.empty(len(rays), np.int32)
        _check(lib.gpu_trace_closest_hit_prims_multi(self._h, _ptr(rays), C.c_int64(len(rays)), _ptr(out)),
               "gpu_trace_closest_hit_prims_multi")
        return out

    def benchmark_with_bvh(self, rays) -> int:
        rays = np.ascontiguousarray(rays)
        out = C.c_int64(0)
        _check(lib.gpu_benchmark_with_bvh_multi(self._h, _ptr(rays), C.c_int64(len(rays)), C.byref(out)), "gpu_benchmark_with_bvh_multi")
        return int(out.value)

    def render_frame(self, cam, width, height, max_depth=5, seed=0, samples=None, out=None):
        if out is None:
            out = np.empty((width * height, 4), np.uint8)
        sp = None
        if samples is not None:
            sp = np.ascontiguousarray(samples, np.float32)
            assert sp.shape == (width * height, max_depth, 3)
        _check(lib.gpu_render_frame_multi(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(max_depth), C.c_uint64(seed),
                                          _ptr(sp) if sp is not None else None, _ptr(out)), "gpu_render_frame_multi")
        return out

    def render_frame_ptr(self, cam, width, height, max_depth, seed, out_ptr: int):
        _check(lib.gpu_render_frame_multi(self._h, _ptr(cam), C.c_int(width), C.c_int(height), C.c_int(max_depth), C.c_uint64(seed),
                                          None, C.c_void_p(out_ptr)), "gpu_render_frame_multi")
