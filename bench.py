#!/usr/bin/env python
"""bench.py -- closest-hit throughput of the SAH-BVH path on B200 (contract in the task brief).

Headline workload (config.workload): the scene north_star's target is quoted on --
1 M random spheres (benchmark.c:306-314 recipe, srand(1)), SAH BVH `build(0,n,0)`,
3840x2160 primary rays of the default camera (main.c:203-211, :356-365), one closest-hit
query per pixel.  One "step" = one full frame through the fused ray-generation + traversal
kernel, hit records (32 B/ray) written to HBM.  `sweep` adds the other single-GPU
configurations (benchmark.c as shipped, default 20-sphere scene 800x600, 100 K spheres 4K primary,
1 M spheres incoherent diffuse-bounce rays, the depth-5 path loop; --config5: 10 M spheres, 8K x 16 spp)
measured the same way.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-sweep] [--no-cpu] [--config5]

N > 1: launched under torchrun, one rank per GPU; the BVH is replicated (every rank builds the
same tree from the same seed), rows are sharded in contiguous bands, every rank traces ITS band
of a full frame per step ("weak": per-GPU work fixed = one full 4K frame's worth of rows per
rank, the frame grows with N) and the frame's primitive-id buffer is assembled on rank 0 by the
trace kernels themselves (peer stores over NVLink; NCCL all_gather as the fallback).
"""
from __future__ import annotations

import argparse
import json
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
METRIC = "closest_hit_grays_per_s"
UNIT = "Grays/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--spheres", type=int, default=N_SPHERES)
    ap.add_argument("--config5", action="store_true",
                    help="also run BASELINE config 5 on this GPU's share: 10 M spheres, 8K frame, primary + bounce batches "
                         "(host build ~40 s; not part of the default run)")
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
    nvidia-smi query taken while a kernel loop is still in flight when NVML is unavailable."""
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
# reference arm: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(n_spheres, width, height, row_step, threads, steps=1, warmup=0):
    """Times the reference's ray_bvh_intersect loop (oracle/_ref = the unmodified reference
    sources; falls back to the oracle port if the prebuilt library is absent) over a bounded
    sample of the headline workload: every `row_step`-th image row of the 4K primary frame."""
    from oracle.oracle_py import Oracle, RefLib, ref_available
    orc = Oracle()
    use_ref = ref_available()     # rays come from the oracle's run-time W/H ray generator (pinned bit-exact to ray.c)
    cam = orc.default_camera()
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
            "sample": f"every {row_step}th row of the {width}x{height} primary frame = {len(rays)} rays per step, "
                      f"{n_spheres} spheres, ray_bvh_intersect loop only (build {build_s:.1f} s excluded)",
            "seconds_per_step": sec, "rays_per_step": len(rays)}


def frame_dims(world):
    """Weak scaling: the frame grows with sqrt(N) in each dimension (same field of view and aspect,
    denser sampling: 4K at N=1, 8K at N=4); W is a multiple of the CTA tile width (32), H of 8*N so
    that the N row bands are equal and tile aligned."""
    scale = float(np.sqrt(world))
    return int(round(WIDTH * scale / 32.0)) * 32, int(round(HEIGHT * scale / (8.0 * world))) * 8 * world


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    W, H = frame_dims(args.gpus)
    # ~2 s per step at 0.015 Mrays/s/thread (measured: 8 threads on 8 vCPU, 1 M spheres)
    row_step = max(1, int(round(H * W / (0.015e6 * threads * 2.0))))
    r = cpu_reference_run(args.spheres, W, H, row_step, threads, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.spheres} random spheres (benchmark.c:306-314 recipe, srand 1), SAH BVH build(0,n,0), "
                               f"{W}x{H} primary rays of the default camera, closest hit, 32 B hit record per ray",
                   "sample": r["sample"]},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def time_steps(torch, fn, steps, warmup, flush, dist_on):
    """W warm-up + K timed steps; per-step CUDA events on the current stream, L2 flushed between
    steps (outside the timed events); barrier + synchronize on both sides."""
    for _ in range(warmup):
        flush()
        fn()
    torch.cuda.synchronize()
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
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
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    return float(np.sum(ms)), ms


def algorithmic(stats, generated, with_input):
    """bytes/ray = 64 V_i + 16 V_s + B_in + 32 ; FLOP/ray = 48 V_i + 25 V_s + 23 [hit] + 26 [generated]."""
    rays = max(stats["rays"], 1)
    vi, vs, hit = stats["inner_visits"] / rays, stats["sphere_tests"] / rays, stats["hits"] / rays
    return {"V_i": vi, "V_s": vs, "hit_rate": hit,
            "bytes_per_ray": 64 * vi + 16 * vs + (24 if with_input else 0) + 32,
            "flop_per_ray": 48 * vi + 25 * vs + 23 * hit + (26 if generated else 0)}


def run_b200(args):
    import torch
    import __graft_entry__ as g
    pkg = g.load_pkg()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist_on = world > 1
    if pkg.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the closest-hit path has no CPU fallback")
    torch.cuda.set_device(local)
    if dist_on:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    main_stream = torch.cuda.current_stream()
    stream = main_stream.cuda_stream
    hbm_peak, peak_src, sm_max = measured_peaks()

    flush_buf = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def flush():
        flush_buf.fill_(1)

    # ---- scene + BVH (replicated per GPU: every rank builds the same tree from the same seed)
    t0 = time.time()
    pkg.srand(1)
    spheres = pkg.benchmark_scene(args.spheres)
    bvh = pkg.GpuBvh.build(spheres, args.spheres, 0, args.spheres, 0, device=local)
    build_s = time.time() - t0
    info = bvh.info()
    cam = pkg.default_camera()

    W, H = frame_dims(world)     # every rank traces one contiguous row band of ~8.3 M rays
    y0, y1 = pkg.shard_rows(H, rank, world)
    rows = y1 - y0
    assert rows * world == H
    n_local = rows * W
    total_rays = W * H
    hits_local = torch.empty((n_local, 8), dtype=torch.float32, device="cuda")          # this rank's 32-byte records
    # frame assembly (N > 1).  Default: FUSED -- rank 0 owns the frame's 4-byte primitive-id buffer (t, point, normal are
    # pure functions of the pixel's ray and the hit sphere: gpu_hits_from_prims_dev rebuilds bit-identical records from
    # the ids), the other ranks map it through CUDA IPC and every rank's trace kernel stores its band straight into it
    # over NVLink peer memory while it computes; an 8-byte all_reduce per frame on a side stream is the "frame complete"
    # barrier.  Fallback (peer mapping unavailable): NCCL all_gather of the id buffer, double-buffered under the next kernel.
    comm = torch.cuda.Stream() if dist_on else None
    ids, fb_ptrs, fb_owned, assembly = None, None, None, None
    if dist_on:
        ok = 1
        handles = [None, None, None, None]
        try:
            if rank == 0:
                fb_owned = ([pkg.peer_alloc(world * n_local * 4, local) for _ in range(2)] + [pkg.peer_alloc(8 * world, local)] +
                            [pkg.peer_alloc(world * n_local * 4, local)])      # 2 id frames, flags, RGBA8 frame of the path loop
                handles = [h for _, h in fb_owned]
        except Exception:
            ok = 0
        dist.broadcast_object_list(handles, src=0)
        try:
            if rank == 0:
                fb_ptrs = [p for p, _ in fb_owned] if fb_owned else None
            elif handles[0] is not None:
                fb_ptrs = [pkg.peer_open(h, local) for h in handles]
            if fb_ptrs is None:
                ok = 0
        except Exception:
            ok = 0
        tok = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(tok, op=dist.ReduceOp.MIN)
        if int(tok.item()) == 1:
            assembly = "peer"
            flags_ptr = fb_ptrs[2]           # world x uint64 on rank 0: slot r = last frame rank r has finished
            if rank == 0:
                for r in range(world):
                    pkg.peer_signal(flags_ptr + 8 * r, 0, stream)
                torch.cuda.synchronize()
            dist.barrier()
        else:
            assembly = "all_gather"
            ids = [torch.empty((world, n_local), dtype=torch.int32, device="cuda") for _ in range(2)]

    frame_no = [0]

    def run_steps(k_steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gather_done = [None, None]
        e0.record()
        for k in range(k_steps):
            flush()
            if dist_on and assembly == "peer":
                # the kernel stores the band into rank 0's frame; a one-thread kernel behind it stores "frame k done" into
                # this rank's flag slot on rank 0.  No collective in the loop: the fence after the loop is the barrier.
                frame_no[0] += 1
                bvh.trace_primary_dev_prims(cam, W, H, y0, y1, hits_local.data_ptr(), fb_ptrs[k & 1] + rank * n_local * 4, stream)
                pkg.peer_signal(flags_ptr + 8 * rank, frame_no[0], stream)
            elif dist_on:
                if gather_done[k & 1] is not None:
                    main_stream.wait_event(gather_done[k & 1])
                buf = ids[k & 1]
                bvh.trace_primary_dev_prims(cam, W, H, y0, y1, hits_local.data_ptr(), buf[rank].data_ptr(), stream)
                ek = torch.cuda.Event()
                ek.record(main_stream)
                with torch.cuda.stream(comm):
                    comm.wait_event(ek)
                    dist.all_gather_into_tensor(buf.view(world * n_local), buf[rank])
                    gd = torch.cuda.Event()
                    gd.record(comm)
                    gather_done[k & 1] = gd
            else:
                bvh.trace_primary_dev(cam, W, H, y0, y1, hits_local.data_ptr(), stream)
        if dist_on:
            main_stream.wait_stream(comm)
        e1.record()
        return e0, e1

    def fence():
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(args.warmup)
    fence()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = pkg.kernel_launches()
    e0, e1 = run_steps(args.steps)
    fence()
    launches = pkg.kernel_launches() - launches0
    clk = clocks.stop()
    assembled_ok = None
    render_multi = None
    if dist_on:
        # config-5 style frame assembly of the full path loop: every rank renders its band (trace_ray, MAX_DEPTH 5,
        # 1 spp, counter-based draws) and the RGBA8 framebuffer (4 B/pixel) is assembled on rank 0; same timing method
        rgba = torch.empty((world, n_local), dtype=torch.int32, device="cuda") if assembly != "peer" else None

        def render_step():
            if assembly == "peer":      # k_resolve stores the band's RGBA8 pixels straight into rank 0's frame
                frame_no[0] += 1
                bvh.render_frame_dev(cam, W, H, y0, y1, 5, 1, fb_ptrs[3] + rank * n_local * 4, stream)
                pkg.peer_signal(flags_ptr + 8 * rank, frame_no[0], stream)
            else:
                bvh.render_frame_dev(cam, W, H, y0, y1, 5, 1, rgba[rank].data_ptr(), stream)
                dist.all_gather_into_tensor(rgba.view(world * n_local), rgba[rank])

        for _ in range(2):
            render_step()
        fence()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r_steps = max(3, min(args.steps, 5))
        r0.record()
        for _ in range(r_steps):
            render_step()
        r1.record()
        fence()
        tr = torch.tensor([r0.elapsed_time(r1) / r_steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        render_multi = {"workload": f"path loop (renderer.c:21-76), {W}x{H}, MAX_DEPTH 5, 1 spp, row bands, RGBA8 frame assembled on rank 0 " +
                                    ("by the resolve kernels' peer stores" if assembly == "peer" else "by NCCL all_gather"),
                        "ms_per_frame": float(tr.item()), "value": total_rays / (float(tr.item()) * 1e-3) / 1e9, "unit": "Gpaths/s"}
    if dist_on:
        # untimed check of the assembled frame: the records rebuilt from the assembled ids of another rank's band must
        # equal what that rank traced (a 64 K-pixel sample per rank, exchanged through one all_gather)
        m = min(n_local, 1 << 16)
        sample = [torch.empty((m, 8), dtype=torch.float32, device="cuda") for _ in range(world)]
        dist.all_gather(sample, hits_local[:m].contiguous())
        last = (args.steps - 1) & 1
        rebuilt = torch.empty((n_local, 8), dtype=torch.float32, device="cuda")
        ok = True
        if assembly == "peer" and rank == 0:      # every rank has signalled the last frame through the peer mapping
            ok = pkg.peer_read_flags(flags_ptr, world, stream) == [frame_no[0]] * world
        checks = range(1, world) if (assembly == "peer" and rank == 0) else ([] if assembly == "peer" else [(rank + 1) % world])
        for r in checks:
            ry0, ry1 = pkg.shard_rows(H, r, world)
            src = fb_ptrs[last] + r * n_local * 4 if assembly == "peer" else ids[last][r].data_ptr()
            bvh.hits_from_prims_dev(cam, W, H, ry0, ry1, src, rebuilt.data_ptr(), stream)
            torch.cuda.synchronize()
            ok = ok and bool((sample[r].view(torch.int32) == rebuilt[:m].view(torch.int32)).all().item())
        tok = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(tok, op=dist.ReduceOp.MIN)
        assembled_ok = bool(tok.item())
    total_ms = e0.elapsed_time(e1)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dist_on:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = total_rays / (ms_per_step * 1e-3) / 1e9

    # ---- dominant kernel alone (per-launch events, L2 flushed between launches), for the roofline
    def kernel_only():
        bvh.trace_primary_dev(cam, W, H, y0, y1, hits_local.data_ptr(), stream)

    kn = max(3, min(args.steps, 10))
    k_ms, _ = time_steps(torch, kernel_only, kn, 1, flush, False)
    k_ms /= kn

    # ---- e2e through the host-buffer C-ABI call (pinned host output; D2H inside the timed region)
    host_out = torch.empty((n_local, 8), dtype=torch.float32).pin_memory()

    def e2e_step():
        bvh.trace_primary_ptr(cam, W, H, y0, y1, host_out.data_ptr())    # returns after the last D2H chunk landed

    for _ in range(2):
        e2e_step()
    fence()
    e2e_steps = max(3, min(args.steps, 10))
    t_e0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t_e0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist_on:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = total_rays / float(te.item()) / 1e9

    if rank != 0:
        if dist_on:
            if assembly == "peer":
                torch.cuda.synchronize()
                for p in fb_ptrs:
                    pkg.peer_close(p, local)
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- work counters (instrumented kernels, untimed)
    stats = bvh.trace_primary_stats(cam, W, H, y0, y1, hits_local.data_ptr())
    tiles = ((W + 15) // 16) * ((rows + 7) // 8)
    pkg.set_primary_mode(pkg.PRIMARY_PER_RAY)
    stats_ray = bvh.trace_primary_stats(cam, W, H, y0, y1, hits_local.data_ptr())
    pkg.set_primary_mode(pkg.PRIMARY_PACKET)
    model = algorithmic(stats_ray, generated=True, with_input=False)      # SURVEY 8(d) per-ray figure
    peaks = pkg.measure_peaks(local)
    k_s = k_ms * 1e-3
    compulsory_bytes = 32 * n_local + info["n_wide"] * 128 + (info["n_spheres"] + 1) * 16
    touched_bytes = 128 * stats["inner_visits"] + 16 * stats["sphere_tests"] + 32 * n_local   # 128-byte 4-wide nodes
    traffic, issue = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            rec = json.load(f).get("k_trace_primary_packet", {})
        traffic = rec.get("dram_bytes_per_launch")
        if rec.get("warp_inst_per_launch") and W == WIDTH and H == HEIGHT and args.spheres == N_SPHERES:
            # the kernel's real bound (FP32 issue): warp instructions of one launch (ncu) over the issue slots
            # of the timed launch, 4 schedulers x SM count x SM clock
            slots = 4.0 * peaks["sm_count"] * (clk.get("sm_mhz") or sm_max) * 1e6
            issue = {"warp_inst_per_launch": rec["warp_inst_per_launch"], "issue_slots_per_s": slots,
                     "frac": rec["warp_inst_per_launch"] / k_s / slots, "ncu_issue_active_pct": rec.get("issue_active_pct_ncu"),
                     "source": rec.get("inst_source")}
    roofline = {
        "kernel": "k_trace_primary_packet", "bound": "hbm",
        "achieved": compulsory_bytes / k_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
        "frac": compulsory_bytes / k_s / 1e9 / hbm_peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": compulsory_bytes,
        "what": "compulsory HBM bytes of one launch: 32 B hit record per ray written once + the flattened tree "
                "(128 B per 4-wide node, 16 B/sphere) read once; the ncu profile shows the kernel is instruction-issue bound, "
                "not HBM bound (profiles/)",
        "kernel_ms": k_ms,
        "issue_bound": issue,
        "tile_work": {"tiles": tiles, "nodes_popped_per_tile": stats["inner_visits"] / tiles,
                      "sphere_tests_per_ray": stats["sphere_tests"] / n_local,
                      "exact_leaf_box_tests_per_ray": stats["exact_box_tests"] / n_local,
                      "touched_bytes_per_launch": touched_bytes,
                      "touched_gbs": touched_bytes / k_s / 1e9, "l2_read_peak_gbs": peaks["l2_read_gbs"]},
        "survey_8d_per_ray_model": {
            **model,
            "equivalent_gbs": model["bytes_per_ray"] * n_local / k_s / 1e9,
            "equivalent_tflops": model["flop_per_ray"] * n_local / k_s / 1e12,
            "note": "bytes/flops a one-ray-per-thread traversal of the same tree performs for these rays (counted by the "
                    "instrumented per-ray kernel); the packet kernel does this culling once per 128-ray tile, so the "
                    "equivalent rates exceed the hardware peaks -- the gain is algorithmic"},
        "measured_peaks": peaks,
    }

    sweep = []
    if not args.no_sweep and not dist_on:
        sweep = run_sweep(torch, pkg, bvh, spheres, args, flush, stream, peaks)

    cpu = None
    if not dist_on and not args.no_cpu:
        threads = os.cpu_count() or 1
        row_step = max(1, int(round(HEIGHT * WIDTH / (0.04e6 * threads * 15.0))))
        try:
            r = cpu_reference_run(args.spheres, WIDTH, HEIGHT, row_step, threads)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker must never take the product's number down
            cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"failed: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.spheres} random spheres (benchmark.c:306-314 recipe, srand 1), SAH BVH build(0,n,0), "
                               f"{W}x{H} primary rays of the default camera, closest hit, 32 B hit record per ray",
                   "rays_per_step": total_rays, "rays_per_gpu": n_local, "bvh": info, "build_seconds": build_s,
                   "l2": "160 MiB device memset before every timed step (inside the timed region, ~25 us)",
                   "timing": "one CUDA-event pair around the K steps on the launch stream, max over ranks",
                   "parallelism": f"rays sharded in {world} contiguous row band(s), BVH replicated" +
                                  ((", frame assembly fused into the trace kernel: every rank stores its band's 4-byte primitive ids straight "
                                    "into rank 0's frame buffer over NVLink peer memory (CUDA IPC mapping) and a completion flag behind it; "
                                    "no collective inside the frame loop" if assembly == "peer" else
                                    ", NCCL all_gather of the frame's primitive-id buffer (4 B/pixel) on a side stream, double-buffered "
                                    "so gather k overlaps kernel k+1") if dist_on else "")},
        "clocks": clk,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 64 * world, "d2h_bytes_per_step": 32 * total_rays,
                "api": "gpu_trace_primary (pinned host output buffer per rank); the only input is the 64-byte Camera"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "sweep": sweep,
    }
    if render_multi is not None:
        line["render_path"] = render_multi
    if assembled_ok is not None:
        line["config"]["assembled_frame_check"] = ("records rebuilt from the assembled ids == the owning rank's traced records "
                                                   "(bit-exact, 64 K-pixel sample per rank): " + ("pass" if assembled_ok else "FAIL"))
    print(json.dumps(line), flush=True)
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()


def run_sweep(torch, pkg, bvh1m, spheres1m, args, flush, stream, peaks):
    """Other single-GPU configurations of BASELINE.json, same timing method."""
    out = []
    cam = pkg.default_camera()
    steps, warm = max(5, min(args.steps, 10)), 3

    def one(name, bvh, fn, n_rays, stats, generated, with_input, extra=None):
        total_ms, _ = time_steps(torch, fn, steps, warm, flush, False)
        ms = total_ms / steps
        alg = algorithmic(stats, generated, with_input)
        rec = {"workload": name, "rays": n_rays, "ms_per_step": ms, "value": n_rays / (ms * 1e-3) / 1e9, "unit": UNIT,
               "per_ray": alg,
               "l2_frac": alg["bytes_per_ray"] * n_rays / (ms * 1e-3) / 1e9 / peaks["l2_read_gbs"] if peaks["l2_read_gbs"] else None,
               "fp32_issue_frac": alg["flop_per_ray"] * n_rays / (ms * 1e-3) / peaks["fp32_mix_inst_per_s"] if peaks["fp32_mix_inst_per_s"] else None}
        if extra:
            rec.update(extra)
        out.append(rec)

    # config 1: benchmark.c as the repo runs it (srand(1) instead of time(NULL)): n = 5 000 .. 50 000, centres
    # U[-500,500)^3, r = 0.5, build(0, n-1, 20) (benchmark.c:317), 10 000 rays for benchmark_no_bvh and the NEXT
    # 10 000 of the same rand() stream for benchmark_with_bvh; rows in benchmark_data.txt's "%d %f %f" format.
    out.append(config1_as_shipped(pkg, args))

    # config 2: default renderer scene, 800x600
    pkg.srand(1)
    s20 = pkg.random_scene(20)
    b20 = pkg.GpuBvh.build(s20, 20, 0, 20, 0)
    buf = torch.empty((WIDTH * HEIGHT, 8), dtype=torch.float32, device="cuda")
    st = b20.trace_primary_stats(cam, 800, 600, 0, 600, buf.data_ptr())
    one("default renderer scene: 20 random spheres, 800x600 primary", b20,
        lambda: b20.trace_primary_dev(cam, 800, 600, 0, 600, buf.data_ptr(), stream), 480000, st, True, False)
    b20.close()

    # config 3: 100 K spheres, 4K primary
    pkg.srand(1)
    s100 = pkg.benchmark_scene(100000)
    b100 = pkg.GpuBvh.build(s100, 100000, 0, 100000, 0)
    st = b100.trace_primary_stats(cam, WIDTH, HEIGHT, 0, HEIGHT, buf.data_ptr())
    one("100K random spheres, 3840x2160 primary", b100,
        lambda: b100.trace_primary_dev(cam, WIDTH, HEIGHT, 0, HEIGHT, buf.data_ptr(), stream), WIDTH * HEIGHT, st, True, False)
    b100.close()

    # config 4: 1 M spheres, 2^24 incoherent diffuse-bounce rays (renderer.c:51-54 recipe from the 4K frame's hits)
    n_b = 1 << 24
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
    one("1M random spheres, 2^24 incoherent diffuse-bounce rays (ray buffer in HBM)", bvh1m,
        lambda: bvh1m.trace_closest_hit_dev(d_rays.data_ptr(), n_b, d_hits.data_ptr(), stream), n_b, st, False, True)
    # e2e of the ray-buffer API with pinned host buffers (H2D 24 B/ray + D2H 32 B/ray inside the call)
    h_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 6)).pin_memory()
    h_hits = torch.empty((n_b, 8), dtype=torch.float32).pin_memory()
    bvh1m.trace_closest_hit_ptr(h_rays.data_ptr(), n_b, h_hits.data_ptr())
    t0 = time.perf_counter()
    for _ in range(3):
        bvh1m.trace_closest_hit_ptr(h_rays.data_ptr(), n_b, h_hits.data_ptr())
    dt = (time.perf_counter() - t0) / 3
    out[-1]["e2e"] = {"value": n_b / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 24 * n_b, "d2h_bytes_per_step": 32 * n_b,
                      "api": "gpu_trace_closest_hit (pinned host buffers)"}
    # the whole path loop (trace_ray, MAX_DEPTH 5, 1 sample per pixel) on the 1 M-sphere 4K frame,
    # counter-based random draws, RGBA8 frame left in HBM; the ray count comes from the per-depth survivors
    d_rgba = torch.empty((WIDTH * HEIGHT,), dtype=torch.int32, device="cuda")
    fn = lambda: bvh1m.render_frame_dev(cam, WIDTH, HEIGHT, 0, HEIGHT, 5, 1, d_rgba.data_ptr(), stream)
    total_ms, _ = time_steps(torch, fn, steps, warm, flush, False)
    ms = total_ms / steps
    out.append({"workload": "path loop (renderer.c:21-76): 1M random spheres, 3840x2160, MAX_DEPTH 5, 1 spp, RGBA8 out",
                "pixels": WIDTH * HEIGHT, "ms_per_step": ms, "value": WIDTH * HEIGHT / (ms * 1e-3) / 1e9, "unit": "Gpaths/s"})
    if args.config5:
        out.append(config5_10m(torch, pkg, args, stream))
    return out


def config5_10m(torch, pkg, args, stream):
    """BASELINE config 5 on ONE GPU: 10 M spheres (tree HBM resident: 640 MB binary + 300 MB wide + 160 MB
    spheres, replicated per GPU), 7680x4320 frame; 16 spp = 1 primary batch + 15 diffuse-bounce batches per
    pixel (the reference has no pixel jitter).  One bounce batch is timed and counted 15 times."""
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
            "primary": {"ms": ms_p, "value": W * H / (ms_p * 1e-3) / 1e9, "nodes_per_tile_or_ray": st_p["inner_visits"] / max(st_p["rays"], 1),
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
