"""`b2d` command line, mirroring rs_doom's flags (reference src/main.rs:17-80,89-124):

    python -m rust_doom_b200.cli --iwad doom1.wad --level 0 --resolution 1920x1080 [--fov 65]
                                 [--poses N] [--dump frame.ppm] [--device 0]
    python -m rust_doom_b200.cli --iwad doom1.wad list-levels
    python -m rust_doom_b200.cli --iwad doom1.wad check

`list-levels` prints "<index> <name>" per level (main.rs:116-121); `check` loads and compiles every level
(the reference's smoke test, main.rs:99-115 / game/src/game.rs:118-129) without needing a GPU.  Without
--iwad a synthetic IWAD is generated (no WAD ships with either project).  Unlike the reference, --fov is
honoured (its value is parsed but never read there: main.rs:131 vs game/src/game.rs:72)."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np


def encode_ppm(rgb: np.ndarray) -> bytes:
    """Binary PPM (P6) of an (H, W, 3) uint8 image."""
    h, w, _ = rgb.shape
    return b"P6\n%d %d\n255\n" % (w, h) + np.ascontiguousarray(rgb).tobytes()


def encode_png(rgb: np.ndarray) -> bytes:
    """Minimal PNG (8-bit RGB, one IDAT, filter 0) of an (H, W, 3) uint8 image."""
    import struct
    import zlib
    h, w, _ = rgb.shape
    raw = np.zeros((h, 1 + 3 * w), dtype=np.uint8)
    raw[:, 1:] = np.ascontiguousarray(rgb).reshape(h, 3 * w)

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
            + chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)) + chunk(b"IEND", b""))


def rgba_to_rgb(rgba_frame: np.ndarray) -> np.ndarray:
    """(H, W) uint32 R | G<<8 | B<<16 | A<<24 (the library's RGBA8) -> (H, W, 3) uint8."""
    h, w = rgba_frame.shape
    return rgba_frame.view(np.uint8).reshape(h, w, 4)[:, :, :3]


def _main_sharded(b2d, scene, view, poses, args, w, h, world) -> int:
    """Under torchrun: every rank renders its contiguous block of the poses on its own GPU (no data-path
    collective); with --stream the finished index frames travel to rank 0 in pose order (parallel.
    write_frames_in_order), which applies the palette and writes the PPM stream."""
    import torch
    import torch.distributed as dist
    from rust_doom_b200 import parallel

    rank, local_rank = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        s, e, per = parallel.shard_bounds(len(poses), rank, world)
        mine = poses[s:e]
        r = b2d.Renderer(scene, view, device=local_rank, max_batch=max(1, min(per, 256)))
        local = torch.empty((len(mine), h, w), dtype=torch.uint8, device="cuda")
        t0 = time.perf_counter()
        for c0 in range(0, len(mine), r.max_batch):
            c1 = min(len(mine), c0 + r.max_batch)
            dp = torch.from_numpy(mine[c0:c1].view(np.int32).reshape(-1, 4).copy()).cuda()
            r.render_device(dp.data_ptr(), c1 - c0, local[c0:c1].data_ptr())
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            dt = time.perf_counter() - t0
            print("rendered %d frame(s) %dx%d on %d GPUs in %.2f ms" % (len(poses), w, h, world, dt * 1e3))
        if args.stream or args.dump:
            pal = scene.palette_rgb()
            out = open(args.stream, "wb") if (args.stream and rank == 0) else None

            def sink(frames, first):
                if out is not None:
                    for f in frames:
                        out.write(encode_ppm(pal[f]))
                if args.dump and first == 0:
                    with open(args.dump, "wb") as fh:
                        fh.write(encode_png(pal[frames[0]]) if args.dump.lower().endswith(".png") else encode_ppm(pal[frames[0]]))

            parallel.write_frames_in_order(local, len(poses), sink, chunk_frames=32)
            if out is not None:
                out.close()
        return 0
    finally:
        dist.destroy_process_group()


def main(argv=None) -> int:
    import rust_doom_b200 as b2d
    from rust_doom_b200 import poses as P
    from rust_doom_b200 import synthwad

    ap = argparse.ArgumentParser(prog="b2d")
    ap.add_argument("-i", "--iwad", default=None, help="initial WAD file (default: generated synthetic IWAD)")
    ap.add_argument("-m", "--metadata", default=None, help="accepted for compatibility; the sky table is built in")
    ap.add_argument("-r", "--resolution", default="1280x720")
    ap.add_argument("-l", "--level", type=int, default=0)
    ap.add_argument("-f", "--fov", type=float, default=65.0)
    ap.add_argument("--poses", type=int, default=1, help="1 = spawn pose, N>1 = N-pose fly-through")
    ap.add_argument("--dump", default=None, help="write the first frame (.png, otherwise binary PPM)")
    ap.add_argument("--stream", default=None,
                    help="write every frame, in order, as concatenated binary PPMs (e.g. ffmpeg -f image2pipe -i FILE)")
    ap.add_argument("--tics-per-frame", type=int, default=0,
                    help="advance the level time by this many tics (1/35 s) per frame: animated flats / walls, "
                         "scrolling walls, light effects")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("command", nargs="?", choices=["check", "list-levels"], default=None)
    args = ap.parse_args(argv)

    try:
        w, h = (int(v) for v in args.resolution.lower().split("x"))
    except ValueError:
        print("resolution format is WIDTHxHEIGHT", file=sys.stderr)
        return 2
    try:
        arch = b2d.Archive.open(args.iwad) if args.iwad else b2d.Archive.from_bytes(synthwad.build_iwad(1, synthwad.E1_MAPS[:3]))
        if args.command == "list-levels":
            for i, name in enumerate(arch.level_names()):
                print("%3d %8s" % (i, name))
            return 0
        if args.command == "check":
            for i, name in enumerate(arch.level_names()):
                sc = b2d.Scene(arch, i)
                print("Level %d (%s): %d segs, %d subsectors, %d sectors, %d textures: ok"
                      % (i, name, sc.info.n_segs, sc.info.n_ssectors, sc.info.n_sectors, sc.info.n_textures))
            return 0
        scene = b2d.Scene(arch, args.level)
        view = b2d.make_view(w, h, args.fov)
        if args.poses <= 1:
            poses = scene.start_pose if scene.start_pose is not None else P.random_poses(scene, 1, 1)
        else:
            poses = P.flythrough_poses(scene, args.poses, 2)
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world > 1:
            return _main_sharded(b2d, scene, view, poses, args, w, h, world)
        r = b2d.Renderer(scene, view, device=args.device, max_batch=min(len(poses), 256))
        t0 = time.perf_counter()
        if args.tics_per_frame > 0:          # time is a per-batch input: one batch per frame
            rgba = np.empty((len(poses), h, w), dtype=np.uint32)
            for i in range(len(poses)):
                r.set_time(i * args.tics_per_frame)
                rgba[i] = r.render(poses[i:i + 1], rgba=True)[1][0]
        else:
            rgba = r.render(poses, rgba=True)[1]
        dt = time.perf_counter() - t0
        print("rendered %d frame(s) %dx%d in %.2f ms (%.0f frames/s end to end)" % (len(poses), w, h, dt * 1e3, len(poses) / dt))
        if args.dump:
            rgb = rgba_to_rgb(rgba[0])
            with open(args.dump, "wb") as f:
                f.write(encode_png(rgb) if args.dump.lower().endswith(".png") else encode_ppm(rgb))
        if args.stream:
            with open(args.stream, "wb") as f:
                for i in range(len(rgba)):
                    f.write(encode_ppm(rgba_to_rgb(rgba[i])))
        return 0
    except b2d.B2dError as e:
        print("Fatal error: %s" % e, file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
