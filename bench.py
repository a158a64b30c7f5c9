#!/usr/bin/env python
"""bench.py -- closest-hit throughput of the SAH-BVH path on B200 (contract in the task brief).

Headline workload (config.workload): the scene north_star's target is quoted on --
1 M random spheres (benchmark.c:306-314 recipe, srand(1)), SAH BVH `build(0,n,0)`,
3840x2160 primary rays of the default camera (main.c:203-211, :356-365), one closest-hit
query per pixel.  One "step" = one full frame through the fused ray-generation + traversal
kernel, 32-byte hit records written to HBM.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-sweep] [--no-cpu] [--config5]

value   frame resident in HBM (records 32 B/ray), CUDA events around the K steps, L2 flushed before every step.
e2e     the same frame through the host-buffer C-ABI call a driver makes (gpu_trace_primary_prims: the hit OBJECT
        per pixel, 4 B -- hit.h:13; the other HitRecord fields are derived on demand, bit-identical), wall clock
        around the call, D2H inside; `e2e_variants` lists the other output formats (32-byte records, {t,prim},
        the path loop's RGBA8 frame, benchmark.c's count).
sweep   N=1: the other single-GPU configurations of BASELINE.json.
multi   N>1: launched under torchrun, one rank per GPU; the BVH is replicated (every rank builds the same tree from
        the same seed), rows are sharded in contiguous bands and the frame's 4-byte id buffer is assembled on rank 0
        by the trace kernels' own stores over NVLink peer memory.  Rank 0 CONSUMES every frame inside the timed
        loop (waits for the bands' flags on the device, counts the hits like benchmark.c:242-244, signals
        "consumed"), and a rank reuses a frame buffer only after that signal (double buffering with back-pressure).
        Weak scaling: the frame grows with sqrt(N) per dimension at CONSTANT PIXEL PITCH (the field of view widens
        with the frame), so the rays of one GPU are as coherent at N=8 as at N=1.  Extra rows: fixed-frame strong
        scaling vs sphere count (config 3), config 5 (10 M spheres, 8K x 16 spp, framebuffer gathered), and the
        one-process N-device C-ABI path (gpu_bvh_replicate + gpu_trace_primary_prims_multi).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SPHERES = 1_000_000
WIDTH, HEIGHT = 3840, 2160
FOV0 = 45.0
METRIC = "closest_hit_grays_per_s"
UNIT = "Grays/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-sweep", action="store_true", help="skip the other configurations (N=1 sweep, N>1 extra rows)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--spheres", type=int, default=N_SPHERES)
    ap.add_argument("--config5", action="store_true",
                    help="N=1: also run BASELINE config 5 on this GPU's share (10 M spheres, 8K frame, 1 primary + 15 bounce batches)")
    ap.add_argument("--no-config5", action="store_true", help="N>1: skip the 10 M-sphere 8K x 16 spp row")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The timed region is a few milliseconds long, so the
    sampler polls NVML from a thread every ~2 ms (nvidia-smi -lms cannot go below ~100 ms); it falls back to one
    nvidia-smi query taken right after the timed steps when NVML is unavailable."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(mhz), int(mask)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        if self.h is None:
            return self._smi_once()
        self.stop_flag = True
        self.thread.join(timeout=2)
        if not self.samples:
            return self._smi_once()
        reasons = set()
        for _, mask in self.samples:
            for bit, name in self.REASONS.items():
                if mask & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median([m for m, _ in self.samples])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(self.samples), "how": "NVML polled every ~2 ms during the timed steps"}

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            f = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = [nm for k, nm in enumerate(names) if f[2 + k].lower().startswith("active")]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": reasons, "samples": 1,
                    "how": "one nvidia-smi query right after the timed steps (NVML polling unavailable)"}
        except Exception as e:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"no samples: {e}"]}


# ------------------------------------------------------------------------------------------------
# the workload's frame
# ------------------------------------------------------------------------------------------------
def frame_dims(world):
    """Weak scaling at constant pixel pitch: W x H grows with sqrt(N) per dimension and the vertical field of view
    widens so that tan(fov/2) / H stays what it is for the 4K frame at 45 degrees -- the image-plane distance
    between neighbouring rays, hence the coherence inside a 16 x 8 tile, does not depend on N.  W is a multiple of
    the CTA group width (32), H of 16 * N (equal, block-aligned row bands).  Returns (W, H, fov in degrees)."""
    scale = math.sqrt(world)
    W = int(round(WIDTH * scale / 32.0)) * 32
    H = int(round(HEIGHT * scale / (16.0 * world))) * 16 * world
    fov = FOV0 if world == 1 else 2.0 * math.degrees(math.atan(math.tan(math.radians(FOV0) / 2.0) * H / HEIGHT))
    return W, H, fov


def workload_text(n_spheres, W, H, fov):
    cam = "the default camera (main.c:203-211)" + ("" if abs(fov - FOV0) < 1e-9 else f" with fov {fov:.3f} deg (constant pixel pitch)")
    return (f"{n_spheres} random spheres (benchmark.c:306-314 recipe, srand 1), SAH BVH build(0,n,0), "
            f"{W}x{H} primary rays of {cam}, closest hit, 32 B hit record per ray")


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(n_spheres, width, height, fov, row_step, threads, steps=1, warmup=0):
    """Times the reference's ray_bvh_intersect loop (oracle/_ref = the unmodified reference sources; falls back to
    the oracle port if the prebuilt library is absent) over a bounded SAMPLE of the workload: every `row_step`-th
    image row of the primary frame."""
    from oracle.oracle_py import Oracle, RefLib, ref_available
    orc = Oracle()
    use_ref = ref_available()     # rays come from the oracle's run-time W/H ray generator (pinned bit-exact to ray.c)
    cam = orc.default_camera()
    cam["fov"] = np.float32(fov)
    rows = list(range(0, height, row_step))
    rays = np.concatenate([orc.primary_rays(cam, width, height, y, y + 1) for y in rows])
    t_build0 = time.time()
    if use_ref:
        R = RefLib()
        R.srand(1)
        s = R.scene_benchmark(n_spheres)
        root, _ = R.build(s, 0, n_spheres, 0)
        kind = "reference"

        def run():
            _, sec = R.trace_bvh(root, s, rays, threads)
            return sec
    else:
        orc.srand(1)
        s = orc.scene_benchmark(n_spheres)
        nodes = orc.build(s, 0, n_spheres, 0)
        kind = "port"

        def run():
            _, sec = orc.trace_bvh(nodes, s, n_spheres, rays, threads)
            return sec
    build_s = time.time() - t_build0
    for _ in range(warmup):
        run()
    secs = [run() for _ in range(steps)]
    sec = float(np.mean(secs))
    return {"value": len(rays) / sec / 1e9, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": (f"SAMPLED: every {row_step}th row of the" if row_step > 1 else "the WHOLE") +
                      f" {width}x{height} primary frame = {len(rays)} rays per step, "
                      f"{n_spheres} spheres, ray_bvh_intersect loop only (build {build_s:.1f} s excluded); rate-normalised",
            "seconds_per_step": sec, "rays_per_step": len(rays)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    W, H, fov = frame_dims(args.gpus)
    # ~2 s per step at 0.015 Mrays/s/thread (measured: 8 threads on 8 vCPU, 1 M spheres)
    row_step = max(1, int(round(H * W / (0.015e6 * threads * 2.0))))
    r = cpu_reference_run(args.spheres, W, H, fov, row_step, threads, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.spheres, W, H, fov), "sample": r["sample"]},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def time_steps(torch, fn, steps, warmup, flush):
    """W warm-up + K timed steps; per-step CUDA events on the current stream, L2 flushed between
    steps (outside the timed events)."""
    for _ in range(warmup):
        flush()
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    return float(np.sum(ms)), ms


def wall(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


KERNEL_NAMES = {0: "k_trace_primary", 1: "k_trace_primary_packet", 2: "k_trace_primary_macro"}


def settle_primary(torch, bvh, cam, W, H, y0, y1, d_out, stream, max_launches=24):
    """SAHBVH_PRIMARY_AUTO samples both packet kernels on the first launches of a frame shape; a steady-state measurement
    starts after the handle has settled.  Returns the kernel id in use (1 = 16x8 tiles, 2 = 32x8 macro tiles)."""
    for _ in range(max_launches):
        k = bvh.primary_kernel_in_use(W, y1 - y0)
        if k >= 0:
            return k
        bvh.trace_primary_dev(cam, W, H, y0, y1, d_out, stream)
        torch.cuda.synchronize()
    return bvh.primary_kernel_in_use(W, y1 - y0)


def algorithmic(stats, generated, with_input):
    """bytes/ray = 64 V_i + 16 V_s + B_in + 32 ; FLOP/ray = 48 V_i + 25 V_s + 23 [hit] + 26 [generated]."""
    rays = max(stats["rays"], 1)
    vi, vs, hit = stats["inner_visits"] / rays, stats["sphere_tests"] / rays, stats["hits"] / rays
    return {"V_i": vi, "V_s": vs, "hit_rate": hit,
            "bytes_per_ray": 64 * vi + 16 * vs + (24 if with_input else 0) + 32,
            "flop_per_ray": 48 * vi + 25 * vs + 23 * hit + (26 if generated else 0)}


class RawDeviceArray:
    """int32 view of a raw device pointer (sahbvh_peer_alloc memory) for torch.as_tensor"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


class Env:
    """torch / dist / package plumbing shared by the legs of the B200 arm."""

    def __init__(self, args):
        import torch
        import __graft_entry__ as g
        self.torch = torch
        self.pkg = g.load_pkg()
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist_on = self.world > 1
        if self.pkg.device_count() < 1:
            raise SystemExit("bench.py: no CUDA device; the closest-hit path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dist = None
        self.host_group = None
        if self.dist_on:
            import torch.distributed as dist
            self.dist = dist
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.host_group = dist.new_group(backend="gloo")    # host-side barriers that leave the GPUs idle
        self.main_stream = torch.cuda.current_stream()
        self.stream = self.main_stream.cuda_stream
        self.flush_buf = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def flush(self):
        self.flush_buf.fill_(1)

    def fence(self):
        self.torch.cuda.synchronize()
        if self.dist_on:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        if self.dist_on:
            self.dist.barrier(group=self.host_group)

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        if self.dist_on:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([int(x)], dtype=self.torch.int64, device="cuda")
        if self.dist_on:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())


class FrameGather:
    """N > 1: the frame's 4-byte primitive-id buffer lives on rank 0 (sahbvh_peer_alloc), every other process maps it
    through CUDA IPC and passes `mapped + band offset` to its trace kernel, whose row stores travel over NVLink while
    it computes.  Flags in the same peer memory: slot r = last frame rank r finished, slot `world` = last frame rank 0
    CONSUMED.  Per frame k: a rank waits on the device for consumed >= k-2 (the buffer it is about to overwrite),
    traces, signals; rank 0's consumer stream waits for all bands of frame k, counts the frame's hits
    (benchmark.c:242-244's reduction) and signals consumed = k.  No collective and no host round trip in the loop."""

    def __init__(self, env, max_pixels):
        pkg, dist, torch = env.pkg, env.dist, env.torch
        self.env = env
        self.ok = False
        self.frame_no = 0
        self.max_pixels = max_pixels
        world, rank, local = env.world, env.rank, env.local
        handles = [None] * 4
        owned = None
        ok = 0 if os.environ.get("SAHBVH_BENCH_NO_IPC") else 1      # test hook: exercise the no-assembly fallback
        try:
            if rank == 0:
                owned = [pkg.peer_alloc(max_pixels * 4, local) for _ in range(2)] + [pkg.peer_alloc(8 * (world + 1), local),
                                                                                    pkg.peer_alloc(max_pixels * 4, local)]
                handles = [h for _, h in owned]
        except Exception:
            ok = 0
        dist.broadcast_object_list(handles, src=0)
        self.ptrs = None
        try:
            if rank == 0:
                self.ptrs = [p for p, _ in owned] if owned else None
            elif handles[0] is not None:
                self.ptrs = [pkg.peer_open(h, local) for h in handles]
            if self.ptrs is None:
                ok = 0
        except Exception:
            ok = 0
        tok = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(tok, op=dist.ReduceOp.MIN)
        self.ok = int(tok.item()) == 1
        self.status = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.count = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.cons = torch.cuda.Stream()
        self.cons_done = [None, None]
        self.own_done = None
        self.rows_table = {}       # calibrate(): measured, load-balanced bands per frame height
        self.calibration = None
        if self.ok:
            self.flags = self.ptrs[2]
            self.consumed = self.flags + 8 * world
            if rank == 0:
                for r in range(world + 1):
                    pkg.peer_signal(self.flags + 8 * r, 0, env.stream)
                torch.cuda.synchronize()
            dist.barrier()

    # Ordering rules.  A device-side spin wait (sahbvh_peer_wait) may only depend on ANOTHER GPU: two streams of one
    # device can share a hardware queue, so a spinning kernel ahead of the kernel it waits for would dead-lock until
    # its timeout.  Rank 0's own band -> consumer and consumer -> rank 0's next frame are therefore ordered by CUDA
    # events; flags carry only cross-GPU edges (bands of ranks 1.. -> consumer, consumed -> ranks 1..).
    def wait_reusable(self, k, timeout_ms):
        """before overwriting buffer k&1: frame k-2 must have been consumed"""
        if k <= 2:
            return
        if self.env.rank == 0:
            self.env.main_stream.wait_event(self.cons_done[k & 1])
        else:
            self.env.pkg.peer_wait(self.consumed, 1, k - 2, timeout_ms, self.status.data_ptr(), self.env.stream)

    def band_done(self, k):
        """this rank's band of frame k is in place (enqueued behind the kernel that writes it)"""
        if self.env.rank == 0:
            self.own_done = self.env.torch.cuda.Event()
            self.own_done.record(self.env.main_stream)
        else:
            self.env.pkg.peer_signal(self.flags + 8 * self.env.rank, k, self.env.stream)

    def consumer_begin(self, k, timeout_ms):
        """rank 0: the consumer stream advances once every band of frame k has landed"""
        self.cons.wait_event(self.own_done)
        if self.env.world > 1:
            self.env.pkg.peer_wait(self.flags + 8, self.env.world - 1, k, timeout_ms, self.status.data_ptr(), self.cons.cuda_stream)

    def consumer_end(self, k):
        self.env.pkg.peer_signal(self.consumed, k, self.cons.cuda_stream)
        ev = self.env.torch.cuda.Event()
        ev.record(self.cons)
        self.cons_done[k & 1] = ev

    def calibrate(self, bvh, cam, W, H, d_records, reps=4, rounds=2):
        """Load balance by measurement.  Bands of equal height do not cost the same (the frame's rows see different parts
        of the scene) and rank 0 also consumes every frame (c = event time of the consumer's pass over the whole frame).
        Two rounds of: every rank times its current band (t_r), the times are shared, and the band heights are rescaled
        to h_r * (T - c [r = 0]) / t_r with T = (sum t_r + c) / N, in whole 16-row blocks."""
        env, pkg, torch = self.env, self.env.pkg, self.env.torch
        world, rank = env.world, env.rank
        buf = self.ptrs[0]
        blocks = (H + 15) // 16
        bounds = [blocks * r // world for r in range(world + 1)]
        c = 0.0
        if rank == 0:
            c = time_steps(torch, lambda: pkg.count_prims_dev(buf, W * H, self.count.data_ptr(), env.stream), reps, 2, env.flush)[0] / reps
        history = []
        for _ in range(rounds):
            y0, y1 = bounds[rank] * 16, min(bounds[rank + 1] * 16, H)
            settle_primary(torch, bvh, cam, W, H, y0, y1, d_records, env.stream)
            fn = lambda: bvh.trace_primary_dev_prims(cam, W, H, y0, y1, d_records, buf + y0 * W * 4, env.stream)
            env.fence()
            t = time_steps(torch, fn, reps, 2, env.flush)[0] / reps
            env.fence()
            ts = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(world)]
            env.dist.all_gather(ts, torch.tensor([t, c], dtype=torch.float64, device="cuda"))
            t_all = [float(v[0].item()) for v in ts]
            c_all = float(ts[0][1].item())
            history.append({"rows": [(bounds[r] * 16, min(bounds[r + 1] * 16, H)) for r in range(world)], "band_ms": t_all})
            target = (sum(t_all) + c_all) / world
            want = [(bounds[r + 1] - bounds[r]) * max(target - (c_all if r == 0 else 0.0), 0.25 * target) / max(t_all[r], 1e-6) for r in range(world)]
            scale = blocks / sum(want)
            acc, nb = 0.0, [0]
            for r in range(world):
                acc += want[r] * scale
                nb.append(max(nb[-1] + 1, min(blocks - (world - 1 - r), int(round(acc)))))
            nb[-1] = blocks
            bounds = nb
        self.rows_table = {H: [(bounds[r] * 16, min(bounds[r + 1] * 16, H)) for r in range(world)]}
        self.calibration = {"consumer_ms": c_all, "rounds": history, "rows": self.rows_table[H]}
        y0, y1 = self.rows_table[H][rank]
        settle_primary(torch, bvh, cam, W, H, y0, y1, d_records, env.stream)
        self.count.zero_()
        env.fence()
        return self.calibration

    def band(self, H, balanced, rank=None):
        rank = self.env.rank if rank is None else rank
        if balanced and H in self.rows_table:
            return self.rows_table[H][rank]
        return self.env.pkg.shard_rows(H, rank, self.env.world)

    def frames(self, bvh, cam, W, H, k_steps, d_records, flush=True, timeout_ms=10000, balanced=False):
        """K frames of rows sharded over the ranks (balanced: with the calibrated shorter band for rank 0); returns
        (event0, event1) around them on the main stream."""
        env, pkg, torch = self.env, self.env.pkg, self.env.torch
        assert W * H <= self.max_pixels
        y0, y1 = self.band(H, balanced)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k_steps):
            if flush:
                env.flush()
            self.frame_no += 1
            k = self.frame_no
            buf = self.ptrs[k & 1]
            self.wait_reusable(k, timeout_ms)
            bvh.trace_primary_dev_prims(cam, W, H, y0, y1, d_records, buf + y0 * W * 4, env.stream)
            self.band_done(k)
            if env.rank == 0:
                self.consumer_begin(k, timeout_ms)
                pkg.count_prims_dev(buf, W * H, self.count.data_ptr(), self.cons.cuda_stream)
                self.consumer_end(k)
        if env.rank == 0:
            env.main_stream.wait_stream(self.cons)
        e1.record()
        return e0, e1

    def timeouts(self):
        return self.env.sum_over_ranks(int(self.status.item()))

    def close(self):
        env = self.env
        if not self.ok:
            # set-up failed on some rank: every rank knows (ok is all-reduced); release what this rank holds, no barriers
            for p in (self.ptrs or []):
                try:
                    (env.pkg.peer_free if env.rank == 0 else env.pkg.peer_close)(p, env.local)
                except Exception:
                    pass
            self.ptrs = None
            return
        if not self.ptrs:
            return
        env.torch.cuda.synchronize()
        env.dist.barrier()
        if env.rank != 0:
            for p in self.ptrs:
                env.pkg.peer_close(p, env.local)
        env.dist.barrier()
        if env.rank == 0:
            for p in self.ptrs:
                env.pkg.peer_free(p, env.local)
        self.ptrs = None


def traffic_record(kernel, W, H, n_spheres):
    """Per-launch ncu numbers of a kernel on one configuration (profiles/ncu_traffic.json, keyed by config): DRAM
    bytes and warp instructions of ONE launch.  The workload is deterministic, so the instruction count of a launch
    is a property of (kernel build, scene, frame) and carries over to the timed launches; the DRAM bytes were
    captured WITHOUT the L2 flush the timed steps use (the tree stays in L2 between profiler replays) and are
    labelled so."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        rec = json.load(f).get(kernel, {})
    cfg = rec.get("config") or {"width": WIDTH, "height": HEIGHT, "n_spheres": N_SPHERES}
    if (cfg.get("width"), cfg.get("height"), cfg.get("n_spheres")) != (W, H, n_spheres):
        return None
    return rec


def run_b200(args):
    env = Env(args)
    torch, pkg, dist = env.torch, env.pkg, env.dist
    rank, local, world, dist_on, stream = env.rank, env.local, env.world, env.dist_on, env.stream
    hbm_peak, peak_src, sm_max = measured_peaks()

    # ---- scene + BVH (replicated per GPU: every rank builds the same tree from the same seed)
    t0 = time.time()
    pkg.srand(1)
    spheres = pkg.benchmark_scene(args.spheres)
    bvh = pkg.GpuBvh.build(spheres, args.spheres, 0, args.spheres, 0, device=local)
    build_s = time.time() - t0
    info = bvh.info()
    cam = pkg.default_camera()
    W, H, fov = frame_dims(world)
    cam["fov"] = np.float32(fov)
    y0, y1 = pkg.shard_rows(H, rank, world)
    rows = y1 - y0
    assert rows * world == H
    n_local = rows * W
    total_rays = W * H
    # room for the largest band of the load-balanced frame loop (the ranks other than the root take a little more than 1/N)
    hits_local = torch.empty((max(n_local + n_local // 4 + 16 * W, WIDTH * HEIGHT // world + 16 * WIDTH), 8), dtype=torch.float32, device="cuda")

    gather = None
    if dist_on:
        gather = FrameGather(env, max(total_rays, 7680 * 4320))
        if not gather.ok:
            # no CUDA IPC peer mapping on this box: the bands are still traced and timed, but nothing assembles or
            # consumes the frame; the line says so (config.parallelism, frame_consumer) instead of losing the run
            if rank == 0:
                print("bench.py: CUDA IPC peer mapping of the frame buffer is unavailable; timing the bands without frame assembly", file=sys.stderr)
            gather.close()
            gather = None
    assembled = gather is not None

    def run_steps(k_steps):
        if assembled:
            return gather.frames(bvh, cam, W, H, k_steps, hits_local.data_ptr(), balanced=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k_steps):
            env.flush()
            bvh.trace_primary_dev(cam, W, H, y0, y1, hits_local.data_ptr(), stream)
        e1.record()
        return e0, e1

    kernel_id = settle_primary(torch, bvh, cam, W, H, y0, y1, hits_local.data_ptr(), stream)
    if gather is not None:
        gather.calibrate(bvh, cam, W, H, hits_local.data_ptr())
    run_steps(args.warmup)
    env.fence()
    if gather is not None:
        gather.count.zero_()
    clocks = ClockSampler(local)
    clocks.start()
    env.fence()      # barrier + synchronize directly before the timed steps: the sampler's start-up must not skew the ranks' start
    launches0 = pkg.kernel_launches()
    e0, e1 = run_steps(args.steps)
    env.fence()
    launches = pkg.kernel_launches() - launches0
    clk = clocks.stop()
    my_ms = e0.elapsed_time(e1) / args.steps
    ms_per_step = env.max_over_ranks(my_ms)
    value = total_rays / (ms_per_step * 1e-3) / 1e9
    per_rank_ms = [my_ms]
    if dist_on:
        t_all = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(t_all, torch.tensor([my_ms], dtype=torch.float64, device="cuda"))
        per_rank_ms = [float(t.item()) for t in t_all]
    fy0, fy1 = gather.band(H, True) if gather is not None else (y0, y1)       # the band this rank traced in the timed loop
    local_hits = int((hits_local[:(fy1 - fy0) * W, 1].view(torch.int32) >= 0).sum().item())
    frame_hits = env.sum_over_ranks(local_hits)
    consumer = None
    if dist_on and not assembled:
        consumer = {"what": "NOT RUN: no CUDA IPC peer mapping on this box, the bands were traced without frame assembly", "ok": None}
    if assembled:
        consumed_total = int(gather.count.item()) if rank == 0 else 0
        consumer = {"what": "rank 0 consumes every assembled frame inside the timed loop: device-side wait for all bands' flags, hit "
                            "count over the frame's ids (benchmark.c:242-244), 'consumed' flag; producers wait for it before "
                            "reusing a frame buffer (double buffering with back-pressure)",
                    "hits_counted": consumed_total, "hits_expected": frame_hits * args.steps,
                    "ok": (consumed_total == frame_hits * args.steps) if rank == 0 else None,
                    "wait_timeouts": gather.timeouts(), "load_balance": gather.calibration, "ms_per_step_by_rank": per_rank_ms}

    # ---- the dominant kernel alone (per-launch events, L2 flushed between launches): records and ids-only
    kn = max(3, min(args.steps, 10))
    k_ms = env.max_over_ranks(time_steps(torch, lambda: bvh.trace_primary_dev(cam, W, H, y0, y1, hits_local.data_ptr(), stream), kn, 1, env.flush)[0] / kn)
    d_ids = torch.empty((n_local,), dtype=torch.int32, device="cuda")
    k_ms_ids = env.max_over_ranks(time_steps(torch, lambda: bvh.trace_primary_dev_prims(cam, W, H, y0, y1, 0, d_ids.data_ptr(), stream), kn, 1, env.flush)[0] / kn)

    # ---- e2e through the host-buffer C-ABI calls (pinned host buffers; copies inside the timed region; wall clock)
    e2e_steps = max(3, min(args.steps, 10))
    h_prims = torch.empty((n_local,), dtype=torch.int32).pin_memory()
    h_recs = torch.empty((n_local, 8), dtype=torch.float32).pin_memory()
    h_tp = torch.empty((n_local, 2), dtype=torch.float32).pin_memory()
    h_rgba = torch.empty((n_local,), dtype=torch.int32).pin_memory()

    def e2e_of(fn, reps):
        env.fence()
        return env.max_over_ranks(wall(fn, reps))

    s_prims = e2e_of(lambda: bvh.trace_primary_prims_ptr(cam, W, H, y0, y1, h_prims.data_ptr()), e2e_steps)
    s_recs = e2e_of(lambda: bvh.trace_primary_ptr(cam, W, H, y0, y1, h_recs.data_ptr()), e2e_steps)
    s_tp = e2e_of(lambda: bvh.trace_primary_tprim_ptr(cam, W, H, y0, y1, h_tp.data_ptr()), e2e_steps)
    s_rgba = e2e_of(lambda: bvh.render_frame_ptr(cam, W, H, y0, y1, 5, 1, h_rgba.data_ptr()), max(3, e2e_steps // 2))
    # the ids that came back are the traced records' ids, and the records derived from them on demand are the traced records
    ids_ok = bool((h_prims.numpy() == h_recs.numpy().view(np.int32)[:, 1]).all())
    m = min(n_local, 1 << 18)
    rebuilt = bvh.hits_from_prims(cam, W, H, h_prims.numpy()[:(m // W) * W], y0, y0 + m // W)
    ids_ok = ids_ok and rebuilt.tobytes() == h_recs.numpy()[:(m // W) * W].tobytes()
    ids_ok = env.sum_over_ranks(0 if ids_ok else 1) == 0
    e2e = {"value": total_rays / s_prims / 1e9, "unit": UNIT, "h2d_bytes_per_step": 64 * world, "d2h_bytes_per_step": 4 * total_rays,
           "api": "gpu_trace_primary_prims (pinned host buffer per rank): hit.object per pixel (hit.h:13), 4 B; t / point / normal "
                  "derived on demand by sahbvh_hit_from_prim / gpu_hits_from_prims, bit-identical to the traced records "
                  f"(checked on this run: {'pass' if ids_ok else 'FAIL'}); the only input is the 64-byte Camera",
           "ms_per_step": s_prims * 1e3}
    e2e_variants = [
        {"output": "32-byte hit records (gpu_trace_primary)", "value": total_rays / s_recs / 1e9, "unit": UNIT,
         "ms_per_step": s_recs * 1e3, "d2h_bytes_per_step": 32 * total_rays, "pcie_gbs": 32 * n_local / s_recs / 1e9},
        {"output": "4-byte hit object index (gpu_trace_primary_prims)", "value": total_rays / s_prims / 1e9, "unit": UNIT,
         "ms_per_step": s_prims * 1e3, "d2h_bytes_per_step": 4 * total_rays, "pcie_gbs": 4 * n_local / s_prims / 1e9},
        {"output": "8-byte {t, prim} (gpu_trace_primary_tprim)", "value": total_rays / s_tp / 1e9, "unit": UNIT,
         "ms_per_step": s_tp * 1e3, "d2h_bytes_per_step": 8 * total_rays, "pcie_gbs": 8 * n_local / s_tp / 1e9},
        {"output": "RGBA8 frame of the path loop, trace_ray renderer.c:21-76, MAX_DEPTH 5, 1 spp (gpu_render_frame)",
         "value": total_rays / s_rgba / 1e9, "unit": "Gpaths/s", "ms_per_step": s_rgba * 1e3, "d2h_bytes_per_step": 4 * total_rays},
    ]

    # ---- N > 1: extra rows (all ranks take part; rank 0 reports)
    multi_rows = []
    if assembled and not args.no_sweep:
        multi_rows = run_multi_rows(env, gather, bvh, spheres, hits_local)

    if rank != 0:
        if gather is not None:
            gather.close()
        if dist_on:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- work counters (instrumented kernels, untimed) and the roofline of the dominant kernel
    stats = bvh.trace_primary_stats(cam, W, H, y0, y1, hits_local.data_ptr())
    tiles = ((W + 15) // 16) * ((rows + 7) // 8)
    pkg.set_keep_binary_tree(True)       # the per-ray work model of SURVEY 8(d) is counted by the per-ray kernel over the binary tree
    s2 = spheres.copy()
    bvh_bin = pkg.GpuBvh.build(s2, args.spheres, 0, args.spheres, 0, device=local)     # same tree (the build is idempotent on a built array)
    pkg.set_keep_binary_tree(False)
    pkg.set_primary_mode(pkg.PRIMARY_PER_RAY)
    stats_ray = bvh_bin.trace_primary_stats(cam, W, H, y0, y1, hits_local.data_ptr())
    pkg.set_primary_mode(pkg.PRIMARY_DEFAULT)
    bvh_bin.close()
    model = algorithmic(stats_ray, generated=True, with_input=False)      # SURVEY 8(d) per-ray figure
    peaks = pkg.measure_peaks(local)
    k_s = k_ms * 1e-3
    compulsory_bytes = 32 * n_local + info["n_wide"] * 128 + (info["n_spheres"] + 1) * 16
    touched_bytes = 128 * stats["inner_visits"] + 16 * stats["sphere_tests"] + 32 * n_local   # 128-byte 4-wide nodes
    kernel_id = bvh.primary_kernel_in_use(W, rows) if bvh.primary_kernel_in_use(W, rows) >= 0 else kernel_id
    kname = KERNEL_NAMES.get(kernel_id, "k_trace_primary_packet")
    tiles = ((W + 31) // 32) * ((rows + 7) // 8) if kernel_id == 2 else tiles     # macro tiles of 32 x 8 pixels
    rec = traffic_record(kname, W, rows, args.spheres) if world == 1 else None
    scaled_from = None
    if rec is None and args.spheres == N_SPHERES:
        # N > 1 (or another frame of the same scene): the weak-scaled frame keeps the pixel pitch, so a band's tiles do the
        # same work as the headline frame's; the instruction count is the headline capture scaled by the tile count
        base = traffic_record(kname, WIDTH, HEIGHT, N_SPHERES)
        if base and base.get("warp_inst_per_launch"):
            base_tiles = (((WIDTH + 31) // 32) if kernel_id == 2 else ((WIDTH + 15) // 16)) * ((HEIGHT + 7) // 8)
            rec = dict(base)
            rec["warp_inst_per_launch"] = base["warp_inst_per_launch"] * tiles / base_tiles
            rec["dram_bytes_per_launch"] = None
            scaled_from = f"ESTIMATE: the ncu count of the {WIDTH}x{HEIGHT} capture x {tiles}/{base_tiles} tiles (same scene, same pixel pitch)"
    sm_mhz = clk.get("sm_mhz") or sm_max
    slots = 4.0 * peaks["sm_count"] * sm_mhz * 1e6          # warp-instruction issue slots per second: 4 schedulers per SM
    if rec and rec.get("warp_inst_per_launch"):
        inst = float(rec["warp_inst_per_launch"])
        roofline = {
            "kernel": kname, "bound": "fp32_issue",
            "achieved": inst / k_s / 1e9, "peak": slots / 1e9, "unit": "G warp-inst/s", "frac": inst / k_s / slots,
            "traffic": rec.get("dram_bytes_per_launch"),
            "traffic_note": "static: ncu dram__bytes_read+write of one launch of this configuration, captured WITHOUT the L2 "
                            "flush the timed steps use (profiles/ncu_traffic.json)",
            "what": "instruction issue is what bounds the kernel (SURVEY 8d: min(L2, FP32 issue); ncu: issue_active highest "
                    "utilisation): warp instructions of one launch (ncu smsp__inst_executed.sum of this exact configuration -- "
                    "the workload is deterministic) / kernel time measured in this run / (4 schedulers x SMs x SM clock sampled "
                    "in this run)",
            "warp_inst_per_launch": inst, "ncu_issue_active_pct": rec.get("issue_active_pct_ncu"), "inst_source": rec.get("inst_source"),
        }
        if scaled_from:
            roofline["inst_scaled"] = scaled_from
            roofline["traffic_note"] = "no DRAM capture for this frame size"
    else:
        roofline = {
            "kernel": kname, "bound": "hbm",
            "achieved": compulsory_bytes / k_s / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": compulsory_bytes / k_s / 1e9 / hbm_peak,
            "traffic": None,
            "what": "no ncu instruction count is recorded for this configuration (profiles/ncu_traffic.json holds the headline "
                    "one), so the line falls back to the compulsory-HBM fraction; the kernel is instruction-issue bound"}
    roofline.update({
        "kernel_ms": k_ms, "kernel_ms_ids_only": k_ms_ids, "peak_source": peak_src,
        "hbm_compulsory_frac": compulsory_bytes / k_s / 1e9 / hbm_peak, "hbm_compulsory_bytes_per_launch": compulsory_bytes,
        "l2_touched_frac": touched_bytes / k_s / 1e9 / peaks["l2_read_gbs"] if peaks["l2_read_gbs"] else None,
        "l2_touched_bytes_per_launch": touched_bytes, "l2_read_peak_gbs": peaks["l2_read_gbs"],
        "tile_work": {"tile": "32x8-pixel macro tile (candidates filtered per 8x4-pixel quadrant)" if kernel_id == 2 else "16x8-pixel tile",
                      "tiles": tiles, "nodes_popped_per_tile": stats["inner_visits"] / tiles,
                      "sphere_tests_per_ray": stats["sphere_tests"] / n_local,
                      "exact_leaf_box_tests_per_ray": stats["exact_box_tests"] / n_local, "max_stack": stats["max_stack"]},
        "survey_8d_per_ray_model": {
            **model,
            "equivalent_gbs": model["bytes_per_ray"] * n_local / k_s / 1e9,
            "equivalent_tflops": model["flop_per_ray"] * n_local / k_s / 1e12,
            "note": "bytes/flops a one-ray-per-thread traversal of the same tree performs for these rays (counted by the "
                    "instrumented per-ray kernel); the packet kernel does this culling once per 128-ray tile, so the "
                    "equivalent rates exceed the hardware peaks -- the gain is algorithmic"},
        "measured_peaks": peaks,
    })

    sweep = []
    if not args.no_sweep and not dist_on:
        sweep = run_sweep(env, bvh, spheres, peaks, clk)

    cpu = None
    if not dist_on and not args.no_cpu:
        threads = os.cpu_count() or 1
        row_step = max(1, int(round(HEIGHT * WIDTH / (0.04e6 * threads * 15.0))))
        try:
            r = cpu_reference_run(args.spheres, WIDTH, HEIGHT, FOV0, row_step, threads)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker must never take the product's number down
            cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"failed: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.spheres, W, H, fov),
                   "rays_per_step": total_rays, "rays_per_gpu": n_local, "bvh": info, "build_seconds": build_s,
                   "l2": "160 MiB device memset before every timed step (inside the timed region, ~25 us)",
                   "timing": "one CUDA-event pair around the K steps on the launch stream, max over ranks",
                   "weak_scaling": "frame grows with sqrt(N) per dimension at constant pixel pitch (fov widened with the frame)",
                   "parallelism": f"rays sharded in {world} contiguous row band(s), BVH replicated" +
                                  (", frame assembly fused into the trace kernel: every rank stores its band's 4-byte primitive ids straight "
                                   "into rank 0's frame buffer over NVLink peer memory (CUDA IPC mapping) as 128-byte row segments and a "
                                   "completion flag behind it; rank 0 consumes every frame and the ranks wait for its 'consumed' flag "
                                   "before reusing a buffer; no collective inside the frame loop" if assembled else
                                   (", NO frame assembly (CUDA IPC peer mapping unavailable on this box)" if dist_on else ""))},
        "clocks": clk,
        "e2e": e2e,
        "e2e_variants": e2e_variants,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "sweep": sweep,
    }
    if consumer is not None:
        line["frame_consumer"] = consumer
        line["frame_hits"] = frame_hits
    if multi_rows:
        line["multi_gpu_rows"] = multi_rows
    print(json.dumps(line), flush=True)
    if gather is not None:
        gather.close()
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# N > 1: what BASELINE.json names beyond the headline line
# ------------------------------------------------------------------------------------------------
def run_multi_rows(env, gather, bvh1m, spheres1m, hits_local):
    """Rows measured under --gpus N (every rank calls this; the values are max-over-ranks, rank 0 keeps them):
      * fixed 3840x2160 frame, row bands over the N ranks, ids assembled on rank 0 with the consumer in the loop, for
        5 K / 50 K / 100 K (BASELINE config 3) / 1 M spheres: strong scaling vs sphere count;
      * BASELINE config 5: 10 M spheres, 7680x4320, 16 samples per pixel = 16 accumulated frames of the path loop
        (main.c:368-405 running mean), the averaged RGBA8 framebuffer gathered on rank 0 and copied to the host there;
      * the one-process N-device C-ABI path (gpu_bvh_replicate + gpu_trace_primary_prims_multi), driven by rank 0
        alone while the other processes wait on a host barrier."""
    torch, pkg, dist, args = env.torch, env.pkg, env.dist, env.args
    rank, local, world, stream = env.rank, env.local, env.world, env.stream
    cam = pkg.default_camera()
    rows_out = []
    steps, warm = max(5, min(args.steps, 10)), 3
    W, H = WIDTH, HEIGHT
    y0, y1 = pkg.shard_rows(H, rank, world)

    def strong_row(n, bvh):
        settle_primary(torch, bvh, cam, W, H, y0, y1, hits_local.data_ptr(), stream)
        gather.frames(bvh, cam, W, H, warm, hits_local.data_ptr())
        env.fence()
        gather.count.zero_()
        e0, e1 = gather.frames(bvh, cam, W, H, steps, hits_local.data_ptr())
        env.fence()
        ms = env.max_over_ranks(e0.elapsed_time(e1)) / steps
        kms = env.max_over_ranks(time_steps(torch, lambda: bvh.trace_primary_dev(cam, W, H, y0, y1, hits_local.data_ptr(), stream),
                                            5, 1, env.flush)[0] / 5)
        hits = env.sum_over_ranks(int((hits_local[:(y1 - y0) * W, 1].view(torch.int32) >= 0).sum().item()))
        counted = int(gather.count.item()) if rank == 0 else 0
        return {"workload": f"strong scaling: {n} random spheres, fixed {W}x{H} primary frame in {world} row bands, ids assembled on rank 0, consumer in the loop",
                "n_spheres": n, "rays": W * H, "ms_per_step": ms, "value": W * H / (ms * 1e-3) / 1e9, "unit": UNIT,
                "kernel_ms_slowest_band": kms, "consumer_ok": counted == hits * steps if rank == 0 else None}

    for n in (5000, 50000, 100000):
        pkg.srand(1)
        s = pkg.benchmark_scene(n)
        b = pkg.GpuBvh.build(s, n, 0, n, 0, device=local)
        rows_out.append(strong_row(n, b))
        b.close()
    rows_out.append(strong_row(args.spheres, bvh1m))

    # ---- one process, N devices (rank 0 only; the other ranks' GPUs are idle and theirs to use)
    env.fence()
    if rank == 0:
        try:
            st = pkg.GpuBvhSet.replicate(bvh1m, list(range(world)))
            h_ids = torch.empty((W * H,), dtype=torch.int32).pin_memory()
            sec = wall(lambda: st.trace_primary_prims_ptr(cam, W, H, h_ids.data_ptr()), steps)
            ref = bvh1m.trace_primary_prims(cam, W, H)
            row = {"workload": f"one process, {world} devices through the C-ABI (gpu_bvh_replicate + gpu_trace_primary_prims_multi): "
                               f"{args.spheres} spheres, fixed {W}x{H} frame in {world} row bands, 4-byte ids into ONE pinned host frame",
                   "rays": W * H, "ms_per_step": sec * 1e3, "value": W * H / sec / 1e9, "unit": UNIT, "e2e": True,
                   "equals_single_device_frame": bool((h_ids.numpy() == ref).all())}
            Wb, Hb, fovb = frame_dims(world)
            camb = pkg.default_camera()
            camb["fov"] = np.float32(fovb)
            h_big = torch.empty((Wb * Hb,), dtype=torch.int32).pin_memory()
            secb = wall(lambda: st.trace_primary_prims_ptr(camb, Wb, Hb, h_big.data_ptr()), steps)
            row["weak_frame"] = {"frame": f"{Wb}x{Hb}", "rays": Wb * Hb, "ms_per_step": secb * 1e3, "value": Wb * Hb / secb / 1e9}
            d_frame = torch.empty((Wb * Hb,), dtype=torch.int32, device="cuda")
            d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")

            def dev_frame():
                st.trace_primary_prims_dev(camb, Wb, Hb, d_frame.data_ptr(), stream)
                pkg.count_prims_dev(d_frame.data_ptr(), Wb * Hb, d_cnt.data_ptr(), stream)

            for _ in range(3):
                dev_frame()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(steps):
                dev_frame()
            ev1.record()
            torch.cuda.synchronize()
            msd = ev0.elapsed_time(ev1) / steps
            row["weak_frame_assembled_on_device0"] = {
                "api": "gpu_trace_primary_prims_multi_dev (peer stores into device 0's frame, events as completion flags) + gpu_count_prims_dev as the consumer",
                "ms_per_step": msd, "value": Wb * Hb / (msd * 1e-3) / 1e9,
                "ids_equal_host_frame": bool((d_frame.cpu().numpy() == h_big.numpy()).all())}
            st.close()
            rows_out.append(row)
        except Exception as e:
            rows_out.append({"workload": "one process, N devices through the C-ABI", "error": str(e)})
    env.host_barrier()

    if not args.no_config5:
        rows_out.append(config5_multi(env, gather))
    return rows_out if rank == 0 else []


def config5_multi(env, gather):
    """BASELINE config 5 under --gpus N: 10 M spheres (replicated), 7680x4320, 16 spp.  The reference has no pixel
    jitter: 16 samples per pixel are 16 frames of trace_ray (renderer.c:21-76, random bounce directions) folded by
    main.c:368-405's running mean.  Every rank renders and accumulates ITS row band; on the 16th sample the averaged
    RGBA8 band is written straight into rank 0's framebuffer (peer memory) and flagged; rank 0 waits for all bands
    on the device and copies the assembled frame to pinned host memory (the display step)."""
    torch, pkg, dist = env.torch, env.pkg, env.dist
    rank, local, world, stream = env.rank, env.local, env.world, env.stream
    n, W, H, spp = 10_000_000, 7680, 4320, 16
    pkg.srand(1)
    s = pkg.benchmark_scene(n)
    t0 = time.perf_counter()
    bvh = pkg.GpuBvh.build(s, n, 0, n, 0, device=local)
    build_s = env.max_over_ranks(time.perf_counter() - t0)
    cam = pkg.default_camera()
    y0, y1 = pkg.shard_rows(H, rank, world)
    npx = (y1 - y0) * W
    d_rgba = torch.empty((npx,), dtype=torch.int32, device="cuda")
    d_acc = torch.zeros((npx, 3), dtype=torch.float32, device="cuda")
    d_avg = torch.empty((npx,), dtype=torch.int32, device="cuda")
    fb = gather.ptrs[3]                     # rank 0's RGBA8 frame (peer mapped)
    h_frame = torch.empty((W * H,), dtype=torch.int32).pin_memory() if rank == 0 else None
    fb_tensor = torch.as_tensor(RawDeviceArray(fb, W * H), device="cuda") if rank == 0 else None
    frames = 2

    def one_frame():
        gather.frame_no += 1
        k = gather.frame_no
        for sidx in range(spp):
            bvh.render_frame_dev(cam, W, H, y0, y1, 5, 1000 + sidx, d_rgba.data_ptr(), stream)
            last = sidx == spp - 1
            if last and rank != 0 and k > 1:      # the single RGBA8 frame on rank 0 is rewritten only once frame k-1 was read
                pkg.peer_wait(gather.consumed, 1, k - 1, 60000, gather.status.data_ptr(), stream)
            rc = pkg.lib.gpu_accumulate_dev(C.c_void_p(d_acc.data_ptr()), C.c_int(sidx + 1), C.c_void_p(d_rgba.data_ptr()), C.c_int64(npx),
                                            C.c_void_p((fb + y0 * W * 4) if last else d_avg.data_ptr()), C.c_void_p(stream))
            assert rc == 0
        gather.band_done(k)
        if rank == 0:
            gather.consumer_begin(k, 60000)
            with torch.cuda.stream(gather.cons):
                h_frame.copy_(fb_tensor, non_blocking=True)
            gather.consumer_end(k)
            env.main_stream.wait_stream(gather.cons)

    one_frame()
    env.fence()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(frames):
        one_frame()
    e1.record()
    env.fence()
    ms = env.max_over_ranks(e0.elapsed_time(e1)) / frames
    # primary-only rate of the same tree and frame (closest-hit Grays/s at 10 M spheres)
    hits = torch.empty((npx, 8), dtype=torch.float32, device="cuda")
    settle_primary(torch, bvh, cam, W, H, y0, y1, hits.data_ptr(), stream)
    pms = env.max_over_ranks(time_steps(torch, lambda: bvh.trace_primary_dev(cam, W, H, y0, y1, hits.data_ptr(), stream), 5, 2, env.flush)[0] / 5)
    info = bvh.info()
    bvh.close()
    alpha_ok = None
    if rank == 0:
        alpha_ok = bool(((h_frame.numpy().view(np.uint32) >> 24) == 255).all())
    return {"workload": f"config 5: {n} random spheres replicated on {world} GPUs, {W}x{H} x {spp} spp = {spp} accumulated frames of the path loop "
                        "(trace_ray MAX_DEPTH 5 + main.c:368-405 running mean), row bands, averaged RGBA8 framebuffer gathered on rank 0 "
                        "through peer stores and copied to the host there",
            "pixels": W * H, "samples_per_pixel": spp, "ms_per_frame": ms, "value": W * H * spp / (ms * 1e-3) / 1e9, "unit": "Gpaths/s",
            "primary_only": {"ms_slowest_band": pms, "value": W * H / (pms * 1e-3) / 1e9, "unit": UNIT},
            "build_seconds_slowest_rank": build_s, "bvh": info, "frame_complete_on_host": alpha_ok}


# ------------------------------------------------------------------------------------------------
# N = 1: the other configurations of BASELINE.json
# ------------------------------------------------------------------------------------------------
def run_sweep(env, bvh1m, spheres1m, peaks, clk):
    """Other single-GPU configurations of BASELINE.json, same timing method."""
    torch, pkg, args, stream = env.torch, env.pkg, env.args, env.stream
    out = []
    cam = pkg.default_camera()
    steps, warm = max(5, min(args.steps, 10)), 3
    slots = 4.0 * peaks["sm_count"] * (clk.get("sm_mhz") or 1965.0) * 1e6

    def primary_row(name, bvh, W, H, n):
        buf = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
        fn = lambda: bvh.trace_primary_dev(cam, W, H, 0, H, buf.data_ptr(), stream)
        kid = settle_primary(torch, bvh, cam, W, H, 0, H, buf.data_ptr(), stream)
        total_ms, _ = time_steps(torch, fn, steps, warm, env.flush)
        ms = total_ms / steps
        st = bvh.trace_primary_stats(cam, W, H, 0, H, buf.data_ptr())
        info = bvh.info()
        tiles = (((W + 31) // 32) if kid == 2 else ((W + 15) // 16)) * ((H + 7) // 8)
        touched = 128 * st["inner_visits"] + 16 * st["sphere_tests"] + 32 * W * H
        compulsory = 32 * W * H + info["n_wide"] * 128 + (info["n_spheres"] + 1) * 16
        h_ids = torch.empty((W * H,), dtype=torch.int32).pin_memory()
        sec = wall(lambda: bvh.trace_primary_prims_ptr(cam, W, H, 0, H, h_ids.data_ptr()), steps)
        out.append({"workload": name, "n_spheres": n, "rays": W * H, "ms_per_step": ms, "value": W * H / (ms * 1e-3) / 1e9, "unit": UNIT,
                    "kernel": KERNEL_NAMES.get(kid, "?") + (" (one warp per 32x8-pixel macro tile" if kid == 2 else " (one warp per 16x8-pixel tile") +
                              ": the work is per TILE, not per ray; picked by SAHBVH_PRIMARY_AUTO)",
                    "per_tile": {"tiles": tiles, "wide_nodes_popped": st["inner_visits"] / tiles, "sphere_tests_per_ray": st["sphere_tests"] / (W * H),
                                 "hit_rate": st["hits"] / (W * H)},
                    "l2_touched_frac": touched / (ms * 1e-3) / 1e9 / peaks["l2_read_gbs"] if peaks["l2_read_gbs"] else None,
                    "hbm_compulsory_frac": compulsory / (ms * 1e-3) / 1e9 / measured_peaks()[0],
                    "device_bytes": info["device_bytes"],
                    "e2e": {"value": W * H / sec / 1e9, "unit": UNIT, "h2d_bytes_per_step": 64, "d2h_bytes_per_step": 4 * W * H,
                            "api": "gpu_trace_primary_prims (pinned host buffer)"}})

    # config 1: benchmark.c as the repo runs it (srand(1) instead of time(NULL)): n = 5 000 .. 50 000, centres
    # U[-500,500)^3, r = 0.5, build(0, n-1, 20) (benchmark.c:317), 10 000 rays for benchmark_no_bvh and the NEXT
    # 10 000 of the same rand() stream for benchmark_with_bvh; rows in benchmark_data.txt's "%d %f %f" format.
    out.append(config1_as_shipped(pkg, args))

    # config 2: default renderer scene, 800x600
    pkg.srand(1)
    s20 = pkg.random_scene(20)
    b20 = pkg.GpuBvh.build(s20, 20, 0, 20, 0)
    primary_row("default renderer scene: 20 random spheres, 800x600 primary", b20, 800, 600, 20)
    b20.close()

    # config 3 and the sphere-count sweep on the 4K frame (benchmark.c:296's commented list, up to 1 M here)
    for n in (5000, 50000, 100000):
        pkg.srand(1)
        s = pkg.benchmark_scene(n)
        b = pkg.GpuBvh.build(s, n, 0, n, 0)
        primary_row(f"{n} random spheres, 3840x2160 primary" + (" (BASELINE config 3 on one GPU)" if n == 100000 else ""), b, WIDTH, HEIGHT, n)
        b.close()

    # config 4: 1 M spheres, 2^24 incoherent diffuse-bounce rays (renderer.c:51-54 recipe from the 4K frame's hits)
    n_b = 1 << 24
    buf = torch.empty((WIDTH * HEIGHT, 8), dtype=torch.float32, device="cuda")
    bvh1m.trace_primary_dev(cam, WIDTH, HEIGHT, 0, HEIGHT, buf.data_ptr(), stream)
    torch.cuda.synchronize()
    hits = buf.cpu().numpy().view(pkg.HIT_DT).reshape(-1)
    pkg.srand(2)
    parts, have = [], 0
    while have < n_b:
        r = pkg.bounce_rays(hits)
        parts.append(r)
        have += len(r)
    rays = np.concatenate(parts)[:n_b]
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 6)).cuda()
    d_hits = torch.empty((n_b, 8), dtype=torch.float32, device="cuda")
    st = bvh1m.trace_closest_hit_stats(d_rays.data_ptr(), n_b, d_hits.data_ptr())
    total_ms, _ = time_steps(torch, lambda: bvh1m.trace_closest_hit_dev(d_rays.data_ptr(), n_b, d_hits.data_ptr(), stream), steps, warm, env.flush)
    ms = total_ms / steps
    rays_n = max(st["rays"], 1)
    wi, vs = st["inner_visits"] / rays_n, st["sphere_tests"] / rays_n
    bytes_per_ray = 64 * wi + 16 * vs + 24 + 32
    # what bounds this kernel (ncu, profiles/): the L1 data pipe -- one 64-byte node per lane per visit is two 32-byte
    # sectors in two different LDG.256 wavefronts; the pipe retires one wavefront per cycle per SM
    wavefronts = 2.0 * st["inner_visits"] + 2.0 * st["sphere_tests"]
    wf_peak = peaks["sm_count"] * (clk.get("sm_mhz") or 1965.0) * 1e6
    row4 = {"workload": "1M random spheres, 2^24 incoherent diffuse-bounce rays (ray buffer in HBM)", "n_spheres": args.spheres, "rays": n_b,
            "ms_per_step": ms, "value": n_b / (ms * 1e-3) / 1e9, "unit": UNIT, "kernel": "k_trace_rays_wide",
            "per_ray": {"wide_nodes_visited": wi, "sphere_tests": vs, "hit_rate": st["hits"] / rays_n, "max_visits_one_ray": st["max_visits"],
                        "bytes_per_ray": bytes_per_ray},
            "l2_frac": bytes_per_ray * n_b / (ms * 1e-3) / 1e9 / peaks["l2_read_gbs"] if peaks["l2_read_gbs"] else None,
            "l1_wavefront_frac": wavefronts / (ms * 1e-3) / wf_peak,
            "bound": "L1 data-pipe wavefronts (ncu l1tex__data_pipe_lsu_wavefronts, profiles/): 2 per node visit per lane, 1 per cycle per SM"}
    out.append(row4)
    # e2e of the ray-buffer API with pinned host buffers: records (24 B in + 32 B out), ids (24 + 4), count (24 in, 8 B out)
    h_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 6)).pin_memory()
    h_hits = torch.empty((n_b, 8), dtype=torch.float32).pin_memory()
    h_ids = torch.empty((n_b,), dtype=torch.int32).pin_memory()
    dt = wall(lambda: bvh1m.trace_closest_hit_ptr(h_rays.data_ptr(), n_b, h_hits.data_ptr()), 3, 1)
    dti = wall(lambda: bvh1m.trace_closest_hit_prims_ptr(h_rays.data_ptr(), n_b, h_ids.data_ptr()), 3, 1)
    cnt = [0]

    def count_call():
        cnt[0] = bvh1m.benchmark_with_bvh_ptr(h_rays.data_ptr(), n_b)

    dtc = wall(count_call, 3, 1)
    row4["e2e"] = {"value": n_b / dti / 1e9, "unit": UNIT, "h2d_bytes_per_step": 24 * n_b, "d2h_bytes_per_step": 4 * n_b,
                   "api": "gpu_trace_closest_hit_prims (pinned host buffers)"}
    row4["e2e_variants"] = [
        {"output": "32-byte records (gpu_trace_closest_hit)", "value": n_b / dt / 1e9, "d2h_bytes_per_step": 32 * n_b},
        {"output": "4-byte ids (gpu_trace_closest_hit_prims)", "value": n_b / dti / 1e9, "d2h_bytes_per_step": 4 * n_b,
         "ids_equal_records": bool((h_ids.numpy() == h_hits.numpy().view(np.int32)[:, 1]).all())},
        {"output": "benchmark.c:242-244's count (gpu_benchmark_with_bvh)", "value": n_b / dtc / 1e9, "d2h_bytes_per_step": 8,
         "count": cnt[0], "count_equals_ids": cnt[0] == int((h_ids.numpy() >= 0).sum())}]
    del h_rays, h_hits, h_ids, d_rays, d_hits
    # the whole path loop (trace_ray, MAX_DEPTH 5, 1 sample per pixel) on the 1 M-sphere 4K frame,
    # counter-based random draws, RGBA8 frame left in HBM
    d_rgba = torch.empty((WIDTH * HEIGHT,), dtype=torch.int32, device="cuda")
    fn = lambda: bvh1m.render_frame_dev(cam, WIDTH, HEIGHT, 0, HEIGHT, 5, 1, d_rgba.data_ptr(), stream)
    total_ms, _ = time_steps(torch, fn, steps, warm, env.flush)
    ms = total_ms / steps
    out.append({"workload": "path loop (renderer.c:21-76): 1M random spheres, 3840x2160, MAX_DEPTH 5, 1 spp, RGBA8 out",
                "pixels": WIDTH * HEIGHT, "ms_per_step": ms, "value": WIDTH * HEIGHT / (ms * 1e-3) / 1e9, "unit": "Gpaths/s"})
    if args.config5:
        out.append(config5_10m(torch, pkg, args, stream))
    return out


def config5_10m(torch, pkg, args, stream):
    """BASELINE config 5 on ONE GPU: 10 M spheres (tree HBM resident, replicated per GPU), 7680x4320 frame; 16 spp =
    1 primary batch + 15 diffuse-bounce batches per pixel.  One bounce batch is timed and counted 15 times."""
    n = 10_000_000
    W, H = 7680, 4320
    pkg.srand(1)
    s = pkg.benchmark_scene(n)
    t0 = time.perf_counter()
    bvh = pkg.GpuBvh.build(s, n, 0, n, 0)
    build_s = time.perf_counter() - t0
    cam = pkg.default_camera()
    buf = torch.empty((W * H, 8), dtype=torch.float32, device="cuda")
    steps = 5

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    settle_primary(torch, bvh, cam, W, H, 0, H, buf.data_ptr(), stream)
    ms_p = timed(lambda: bvh.trace_primary_dev(cam, W, H, 0, H, buf.data_ptr(), stream))
    st_p = bvh.trace_primary_stats(cam, W, H, 0, H, buf.data_ptr())
    hits = buf.cpu().numpy().view(pkg.HIT_DT).reshape(-1)
    pkg.srand(2)
    rays = pkg.bounce_rays(hits)
    while len(rays) < W * H:
        rays = np.concatenate([rays, pkg.bounce_rays(hits)])
    rays = rays[:W * H]
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 6)).cuda()
    ms_b = timed(lambda: bvh.trace_closest_hit_dev(d_rays.data_ptr(), W * H, buf.data_ptr(), stream))
    st_b = bvh.trace_closest_hit_stats(d_rays.data_ptr(), W * H, buf.data_ptr())
    info = bvh.info()
    bvh.close()
    total_ms = ms_p + 15 * ms_b
    return {"workload": "config 5 on one GPU: 10M random spheres, 7680x4320, 16 spp = 1 primary + 15 bounce batches",
            "rays": 16 * W * H, "ms_per_step": total_ms, "value": 16 * W * H / (total_ms * 1e-3) / 1e9, "unit": UNIT,
            "primary": {"ms": ms_p, "value": W * H / (ms_p * 1e-3) / 1e9, "wide_nodes_popped_per_tile": st_p["inner_visits"] / (((W + 15) // 16) * ((H + 7) // 8)),
                        "hit_rate": st_p["hits"] / max(st_p["rays"], 1)},
            "bounce": {"ms": ms_b, "value": W * H / (ms_b * 1e-3) / 1e9, "wide_nodes_per_ray": st_b["inner_visits"] / max(st_b["rays"], 1),
                       "sphere_tests_per_ray": st_b["sphere_tests"] / max(st_b["rays"], 1), "hit_rate": st_b["hits"] / max(st_b["rays"], 1)},
            "build_and_upload_s_host": build_s, "bvh": info}


def config1_as_shipped(pkg, args):
    """SURVEY 8(d) config 1.  GPU seconds are wall-clock around the host-buffer C-ABI calls
    (gpu_benchmark_no_bvh / gpu_benchmark_with_bvh: H2D of the rays, kernels, 8 bytes back).  The CPU
    columns time the reference's own loops (oracle/_ref, else the oracle port) on the SAME ray arrays,
    1 core as the repo runs; brute force on a ray subset, scaled (10 000 x 50 000 tests take ~20 s)."""
    rows = []
    cpu = None
    if not args.no_cpu:
        from oracle.oracle_py import Oracle, RefLib, ref_available
        cpu = RefLib() if ref_available() else Oracle()
    pkg.srand(1)
    n_rays = 10000
    for n in range(5000, 50001, 5000):
        s = pkg.benchmark_scene(n)
        s_unbuilt = s.copy()
        t0 = time.perf_counter()
        bvh = pkg.GpuBvh.build(s, n, 0, n - 1, 20)
        build_s = time.perf_counter() - t0
        rays_a, rays_b = pkg.benchmark_rays(n_rays), pkg.benchmark_rays(n_rays)
        bvh.benchmark_no_bvh(rays_a[:64]); bvh.benchmark_with_bvh(rays_b[:64])   # staging buffers, module load
        t0 = time.perf_counter(); hits_a = bvh.benchmark_no_bvh(rays_a); t_a = time.perf_counter() - t0
        t0 = time.perf_counter(); hits_b = bvh.benchmark_with_bvh(rays_b); t_b = time.perf_counter() - t0
        row = {"n_spheres": n, "gpu_no_bvh_s": t_a, "gpu_with_bvh_s": t_b, "intersections_no_bvh": hits_a,
               "intersections_with_bvh": hits_b, "gpu_build_and_upload_s": build_s,
               "multi_leaves": bvh.info()["multi_leaves"]}
        bvh.close()
        if cpu is not None and n in (5000, 25000, 50000):
            sub = 500
            if hasattr(cpu, "export_tree"):          # compiled reference
                root, _ = cpu.build(s_unbuilt, 0, n - 1, 20)
                hb, sec_b = cpu.trace_bvh(root, s_unbuilt, rays_b, 1)
                ha, sec_a = cpu.trace_brute(s_unbuilt, n, rays_a[:sub], 1)
                cpu.free(root)
                kind = "reference"
            else:
                nodes = cpu.build(s_unbuilt, 0, n - 1, 20)
                hb, sec_b = cpu.trace_bvh(nodes, s_unbuilt, n, rays_b, 1)
                ha, sec_a = cpu.trace_brute(s_unbuilt, n, rays_a[:sub], 1)
                kind = "port"
            row.update({"cpu_with_bvh_s": sec_b, "cpu_no_bvh_s": sec_a * n_rays / sub, "cpu_kind": kind, "cpu_cores": 1,
                        "cpu_no_bvh_sample": f"{sub} of {n_rays} rays, scaled",
                        "cpu_intersections_with_bvh": int((hb["prim"] >= 0).sum())})
        rows.append(row)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "benchmark_data.txt"), "w") as f:
            for r in rows:
                f.write("%d %f %f\n" % (r["n_spheres"], r["gpu_no_bvh_s"], r["gpu_with_bvh_s"]))   # benchmark.c:168
    except OSError:
        pass
    return {"workload": "benchmark.c as shipped: n = 5000..50000, build(0,n-1,20), 10000 + 10000 rays, seconds per call "
                        "(benchmark_data.txt columns: n, no-BVH s, with-BVH s)", "rows": rows}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
