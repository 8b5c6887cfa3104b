#!/usr/bin/env python
"""bench.py -- frames/sec at 1920x1080 (palette-index framebuffer, bit-exact vs the oracle).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU oracle on the host cores

Workload (BASELINE.json configs[1]): synthetic E1M1-scale level (seed 1; no real doom1.wad exists in the
environment -- set B2D_IWAD=/path/doom1.wad to use one), 1000-pose fly-through, 1920x1080.  One "step" is
one pass of the hot path (BSP walk + raster) over the 1000-pose batch.  `value` is device-resident
throughput (poses already in HBM, frames written to HBM); `e2e` goes through b2d_render with pinned HOST
buffers -- host poses in, host frames out, both copies inside the timed region.  Multi-GPU: independent pose
blocks per rank (weak scaling: every rank renders the 1000-pose fly-through, rotated by rank), no data-path collective;
the optional frame all-gather is timed separately under "allgather" and never blended into `value`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WIDTH, HEIGHT = 1920, 1080
POSES_PER_STEP = 1000
METRIC = "frames/sec at 1920x1080 (palette-index bit-exact)"
UNIT = "frames/s"


def load_scene(b2d):
    from rust_doom_b200 import synthwad
    iwad = os.environ.get("B2D_IWAD")
    if iwad:
        arch = b2d.Archive.open(iwad)
        name = "%s level 0 (%s)" % (os.path.basename(iwad), arch.level_name(0))
    else:
        arch = b2d.Archive.from_bytes(synthwad.build_iwad(1, ("E1M1",)))
        name = "synthetic SYN_E1M1 (seed 1, E1M1-scale)"
    return b2d.Scene(arch, 0), name


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per raster launch from the committed ncu capture (profiles/roofline.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline.json")) as f:
            return json.load(f).get("raster_dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:] if len(sm) > 2 else sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(blob, poses, threads, steps, warmup, sample):
    """Oracle (the only CPU implementation of this path that exists) on `threads` host cores.
    Each step renders `sample` poses of the workload at 1920x1080."""
    from oracle import render
    view = render.make_view(WIDTH, HEIGHT)
    idx = np.linspace(0, len(poses) - 1, sample).astype(int)
    sub = np.ascontiguousarray(poses[idx])
    for _ in range(warmup):
        render.render(blob, view, sub[:max(threads, 1)], threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        render.render(blob, view, sub, threads=threads)
    dt = time.perf_counter() - t0
    return steps * sample / dt, dt / steps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2d", choices=["b2d", "reference"])
    ap.add_argument("--poses", type=int, default=POSES_PER_STEP)
    ap.add_argument("--cpu-sample", type=int, default=0, help="poses per CPU-baseline step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--pipeline", action="store_true",
                    help="b2d_walk_device / b2d_raster_device on two streams (the walk of batch k+1 under the raster of "
                         "batch k: +1.6 %% frames/s, but the kernels then time each other) instead of one "
                         "b2d_render_device call per step")
    ap.add_argument("--rgba", action="store_true", help="also materialise RGBA8 frames in HBM (5 B/pixel; not the headline config)")
    ap.add_argument("--gather-frames", type=int, default=128, help="frames per rank in the separate all-gather timing (N>1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return 0
        import rust_doom_b200 as b2d
        from rust_doom_b200 import poses as P
        scene, scene_name = load_scene(b2d)
        poses = P.flythrough_poses(scene, args.poses, 2)
        sample = args.cpu_sample or max(cores, min(4 * cores, 64))
        fps, ms = cpu_reference(scene.blob, poses, cores, args.steps, min(args.warmup, 1), sample)
        cb = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
              "sample": "%d of the %d fly-through poses per step at %dx%d, OpenMP over poses" % (sample, args.poses, WIDTH, HEIGHT)}
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s, %d-pose fly-through, %dx%d; CPU oracle (the reference has no CPU "
                                   "implementation of this path: it rasterises in OpenGL)" % (scene_name, args.poses, WIDTH, HEIGHT),
                       "poses_per_step": sample},
            "cpu_baseline": cb,
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    # ------------------------------------------------------------------ this repo's CUDA path
    import torch
    import torch.distributed as dist

    import rust_doom_b200 as b2d
    from rust_doom_b200 import poses as P

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    scene, scene_name = load_scene(b2d)
    n = args.poses
    # every rank renders the same 1000-pose fly-through, cyclically rotated by rank: identical work per GPU
    # (clean weak-scaling efficiency) while no two ranks are on the same pose at the same time
    poses_np = np.roll(P.flythrough_poses(scene, n, 2), -(rank * n // max(world, 1)))
    view = b2d.make_view(WIDTH, HEIGHT)
    r = b2d.Renderer(scene, view, device=local_rank, max_batch=n)
    npix = WIDTH * HEIGHT
    d_poses = torch.from_numpy(poses_np.view(np.int32).reshape(-1, 4).copy()).to(dev)
    d_index = torch.empty((n, HEIGHT, WIDTH), dtype=torch.uint8, device=dev)
    d_rgba = torch.empty((n, HEIGHT, WIDTH), dtype=torch.int32, device=dev) if args.rgba else None
    stream = torch.cuda.current_stream().cuda_stream

    # Default: a step = one b2d_render_device call (BSP walk, then raster, one stream).  With --pipeline a step =
    # the raster of this step's batch + the BSP walk of the next step's batch as separate calls
    # (b2d_walk_device / b2d_raster_device) on two streams, ordered by events inside the library; every step still
    # does one walk and one raster.
    pipelined = args.pipeline
    walk_stream = torch.cuda.Stream(device=dev, priority=-1) if pipelined else None
    pending = [r.walk_device(d_poses.data_ptr(), n, walk_stream.cuda_stream)] if pipelined else None

    def step():
        if pipelined:
            r.raster_device(pending[0], d_index.data_ptr(), d_rgba.data_ptr() if args.rgba else 0, stream)
            pending[0] = r.walk_device(d_poses.data_ptr(), n, walk_stream.cuda_stream)
        else:
            r.render_device(d_poses.data_ptr(), n, d_index.data_ptr(), d_rgba.data_ptr() if args.rgba else 0, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    args.warmup = max(args.warmup, 3)          # timing hygiene: at least 3 warm-up steps
    for _ in range(args.warmup):
        step()
    barrier()
    # parity spot check inside the run: one frame of this batch against the oracle
    if rank == 0:
        from oracle import render as orender
        probe = n // 2
        ofb = orender.render(scene.blob, orender.make_view(WIDTH, HEIGHT), poses_np[probe:probe + 1])
        if not np.array_equal(d_index[probe].cpu().numpy(), ofb[0]):
            raise SystemExit("parity check failed: GPU frame differs from the oracle")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = r.launch_count
    r.profile(True)
    r.profile_read()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    if pipelined:                              # the walk issued by the last step belongs to a step that never comes
        r.raster_device(pending[0], d_index.data_ptr(), d_rgba.data_ptr() if args.rgba else 0, stream)
        torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    walk_ms, raster_ms, batches = r.profile_read()
    r.profile(False)
    launches = r.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n * args.steps / (ms_total / 1e3)

    # roofline of the dominant kernel (raster): algorithmic bytes = W*H*1 per frame (index FB written once)
    peak, peak_src = measured_peak()
    alg_bytes = float(n) * npix * (5 if args.rgba else 1)
    raster_avg_ms = raster_ms / max(batches, 1)
    achieved = alg_bytes / (raster_avg_ms / 1e3) / 1e9 if raster_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if args.rgba else ncu_traffic(), "kernel": "b2d_raster_kernel<%s>" % ("rgba" if args.rgba else "index"),
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": raster_avg_ms,
                "walk_avg_launch_ms": walk_ms / max(batches, 1), "peak_source": peak_src,
                "note": "index-only output (no RGBA materialised); the kernel is issue/latency bound at 32 warps/SM, not HBM bound (DESIGN.md 5/6, profiles/README.md)"}

    # ------------------------------------------------------------------ end to end (host buffers)
    e2e = None
    if not args.no_e2e:
        e2e_n = n
        r2 = b2d.Renderer(scene, view, device=local_rank, max_batch=min(125, e2e_n))
        h_poses = torch.from_numpy(poses_np.view(np.int32).reshape(-1, 4).copy()).pin_memory()
        h_index = torch.empty((e2e_n, HEIGHT, WIDTH), dtype=torch.uint8).pin_memory()
        r2.render_ptr(h_poses.data_ptr(), e2e_n, h_index.data_ptr())        # warm-up (allocations)
        e2e_steps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r2.render_ptr(h_poses.data_ptr(), e2e_n, h_index.data_ptr())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        launches += 0
        e2e = {"value": world * e2e_n * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(poses_np.nbytes), "d2h_bytes_per_step": int(e2e_n * npix),
               "steps": e2e_steps, "api": "b2d_render (pinned host poses in, pinned host frames out, double-buffered D2H)"}
        if rank == 0 and not np.array_equal(h_index[n // 2].numpy(), d_index[n // 2].cpu().numpy()):
            raise SystemExit("e2e path disagrees with the device path")
        del r2

    # ------------------------------------------------------------------ optional frame all-gather (N>1), separate
    allgather = None
    if world > 1 and args.gather_frames > 0:
        from rust_doom_b200.parallel import all_gather_frames
        g = min(args.gather_frames, n)
        out = torch.empty((g * world, HEIGHT, WIDTH), dtype=torch.uint8, device=dev)
        all_gather_frames(d_index[:g], g * world, chunk_frames=64, out=out)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        all_gather_frames(d_index[:g], g * world, chunk_frames=64, out=out)
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        allgather = {"frames": g * world, "ms": float(tg.item()), "frames_per_s": g * world / (float(tg.item()) / 1e3),
                     "note": "NCCL all-gather of finished index frames, NVLink-bound, NOT part of `value`"}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample or max(cores, min(4 * cores, 64))
        reps = 1
        fps, ms = cpu_reference(scene.blob, poses_np, cores, reps, 1, sample)
        # keep the CPU leg near 10-30 s: repeat if it was very quick
        if ms < 3000:
            reps = int(min(100, max(1, 10000 // max(ms, 1))))
            fps, ms = cpu_reference(scene.blob, poses_np, cores, reps, 0, sample)
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d of the %d poses x %d passes at %dx%d, OpenMP over poses (oracle/b2d_oracle.c)" % (sample, n, reps, WIDTH, HEIGHT)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s, %d-pose fly-through per GPU, %dx%d, %s" % (scene_name, n, WIDTH, HEIGHT, "index + RGBA8 framebuffers" if args.rgba else "index framebuffer only"),
                       "poses_per_step_per_gpu": n, "segs": scene.info.n_segs, "subsectors": scene.info.n_ssectors,
                       "parallelism": "pose-sharded x%d, no data-path collective" % world,
                       "step": ("raster of this batch + BSP walk of the next batch, two streams "
                                "(b2d_walk_device / b2d_raster_device)") if pipelined
                               else "BSP walk then raster of one batch, one stream (b2d_render_device)",
                       "l2": "each step writes %.2f GB of frames per GPU (>> 126 MB L2); the scene (%.0f KB) is legitimately cache-resident"
                             % (n * npix / 1e9, scene.info.blob_bytes / 1024.0)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "allgather": allgather}))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
