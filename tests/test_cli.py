"""Host-side CLI pieces (src/main.rs:17-124 mirror): flag parsing, level listing, image writers."""
import struct
import zlib

import numpy as np
import pytest

from rust_doom_b200 import cli


def test_list_levels_and_check_on_synthetic_iwad(capsys):
    assert cli.main(["list-levels"]) == 0
    out = capsys.readouterr().out.split()
    assert out[:2] == ["0", "E1M1"]
    assert cli.main(["check"]) == 0
    assert "Level 0 (E1M1)" in capsys.readouterr().out


def test_bad_resolution_is_an_argument_error(capsys):
    assert cli.main(["--resolution", "1920by1080", "list-levels"]) == 2
    assert "WIDTHxHEIGHT" in capsys.readouterr().err


def test_missing_iwad_reports_fatal_error(capsys):
    assert cli.main(["--iwad", "/nonexistent/doom1.wad", "list-levels"]) == 1
    assert "Fatal error" in capsys.readouterr().err


def test_ppm_and_png_writers_roundtrip():
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    rgba = (rgb[:, :, 0].astype(np.uint32) | (rgb[:, :, 1].astype(np.uint32) << 8)
            | (rgb[:, :, 2].astype(np.uint32) << 16) | np.uint32(0xFF000000))
    assert np.array_equal(cli.rgba_to_rgb(rgba), rgb)
    ppm = cli.encode_ppm(rgb)
    assert ppm.startswith(b"P6\n7 5\n255\n") and ppm[-105:] == rgb.tobytes()
    png = cli.encode_png(rgb)
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    # parse the chunks back
    pos, chunks = 8, {}
    while pos < len(png):
        n, tag = struct.unpack(">I4s", png[pos:pos + 8])
        data = png[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + data) & 0xFFFFFFFF
        chunks[tag] = data
        pos += 12 + n
    assert struct.unpack(">IIBBBBB", chunks[b"IHDR"]) == (7, 5, 8, 2, 0, 0, 0)
    raw = np.frombuffer(zlib.decompress(chunks[b"IDAT"]), np.uint8).reshape(5, 1 + 21)
    assert (raw[:, 0] == 0).all() and np.array_equal(raw[:, 1:].reshape(5, 7, 3), rgb)


@pytest.mark.gpu
def test_cli_renders_dump_and_stream(tmp_path, capsys):
    png, stream = tmp_path / "f.png", tmp_path / "s.ppm"
    assert cli.main(["--resolution", "320x200", "--poses", "3", "--tics-per-frame", "8",
                     "--dump", str(png), "--stream", str(stream)]) == 0
    assert "rendered 3 frame(s) 320x200" in capsys.readouterr().out
    assert png.read_bytes()[:8] == b"\x89PNG\r\n\x1a\n"
    assert len(stream.read_bytes()) == 3 * (len(b"P6\n320 200\n255\n") + 320 * 200 * 3)


def test_scene_palette_matches_oracle_blob(b2d, product_scene, oracle_scene):
    from oracle import scene as S
    pal = product_scene.palette_rgb()
    h = S.header(oracle_scene)
    want = np.frombuffer(oracle_scene, dtype="<u4", count=256, offset=h[S.H_OFF_PALETTE])
    assert pal.shape == (256, 3) and pal.dtype == np.uint8
    assert np.array_equal(pal[:, 0], want & 0xFF) and np.array_equal(pal[:, 2], (want >> 16) & 0xFF)


def _b2d_binary():
    import os
    from rust_doom_b200 import build as B
    B.build()
    return os.path.join(os.path.dirname(B.OUT), "b2d")


def test_compiled_cli_on_the_c_abi(tmp_path):
    """rust-doom_b200/csrc/b2d_cli.cpp (C++ on include/b2d.h only): list-levels / check / error behaviour of
    src/main.rs:89-124, and no CPU rendering path."""
    import subprocess
    from rust_doom_b200 import synthwad
    wad = tmp_path / "syn.wad"
    wad.write_bytes(synthwad.build_iwad(1, ("E1M1", "E1M2")))
    exe = _b2d_binary()
    out = subprocess.run([exe, "--iwad", str(wad), "list-levels"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["0", "E1M1", "1", "E1M2"]
    out = subprocess.run([exe, "-i", str(wad), "check"], capture_output=True, text=True)
    assert out.returncode == 0 and "Level 1 (E1M2)" in out.stdout and out.stdout.count(": ok") == 2
    out = subprocess.run([exe, "-i", str(tmp_path / "missing.wad"), "list-levels"], capture_output=True, text=True)
    assert out.returncode == 1 and "Fatal error" in out.stderr
    bad = tmp_path / "bad.wad"
    bad.write_bytes(b"PWAD" + bytes(8))
    out = subprocess.run([exe, "-i", str(bad), "check"], capture_output=True, text=True)
    assert out.returncode == 1 and "Fatal error" in out.stderr
    out = subprocess.run([exe, "-i", str(wad), "-r", "320by200"], capture_output=True, text=True)
    assert out.returncode == 2


@pytest.mark.gpu
def test_compiled_cli_renders_the_same_frames_as_the_python_mirror(tmp_path, b2d):
    import subprocess
    from rust_doom_b200 import synthwad
    data = synthwad.build_iwad(1, ("E1M1",), cfg=synthwad.SynthConfig(mid_pct=20, thing_pct=30, anim=True))
    wad = tmp_path / "syn.wad"
    wad.write_bytes(data)
    stream = tmp_path / "s.ppm"
    out = subprocess.run([_b2d_binary(), "-i", str(wad), "-r", "320x200", "--poses", "4", "--tics", "40",
                          "--stream", str(stream)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    sc = b2d.Scene(b2d.Archive.from_bytes(data), 0)
    poses = np.repeat(sc.start_pose, 4)
    poses["angle"] = (poses["angle"].astype(np.uint64) + (np.arange(4, dtype=np.uint64) << np.uint64(32)) // np.uint64(4)).astype(np.uint32)
    r = b2d.Renderer(sc, b2d.make_view(320, 200), max_batch=4)
    r.set_time(40)
    rgba = r.render(poses, rgba=True)[1]
    want = b"".join(cli.encode_ppm(cli.rgba_to_rgb(rgba[i])) for i in range(4))
    assert stream.read_bytes() == want


@pytest.mark.gpu
def test_compiled_cli_sharded_world1(tmp_path):
    """The compiled front end drives b2d_render_sharded (communicator from a unique-id file, checksum consumer,
    b2d_renderer_status) through nothing but include/b2d.h; a one-rank run must gather every frame."""
    import os
    import subprocess
    from rust_doom_b200 import build, synthwad
    wad = tmp_path / "t.wad"
    wad.write_bytes(synthwad.build_iwad(1, ("E1M1",)))
    exe = build.build_cli()
    res = subprocess.run([exe, "--iwad", str(wad), "--resolution", "320x200", "--poses", "10", "--world", "1", "--rank", "0",
                          "--chunk", "4", "--id-file", str(tmp_path / "id")], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ))
    assert res.returncode == 0, res.stdout + res.stderr
    assert "rank 0/1: 10 frames gathered in 3 chunk(s)" in res.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["nccl", "ce"])
def test_compiled_cli_sharded_two_ranks(tmp_path, transport):
    """Two processes, two GPUs (skipped on a one-GPU box): the ranks gather each other's frames -- over ncclAllGather and
    over the copy-engine transport -- and both end with the checksum a single rank computes for the whole pose list."""
    import os
    import re
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from rust_doom_b200 import build, synthwad
    wad = tmp_path / "t.wad"
    wad.write_bytes(synthwad.build_iwad(1, ("E1M1",)))
    exe = build.build_cli()
    env = dict(os.environ)
    env["B2D_GATHER"] = "nccl" if transport == "nccl" else "ce"          # ce (copy engines over CUDA IPC) is the default
    base = [exe, "--iwad", str(wad), "--resolution", "640x400", "--poses", "22", "--chunk", "4"]
    one = subprocess.run(base + ["--world", "1", "--rank", "0", "--id-file", str(tmp_path / "id1")], capture_output=True, text=True, timeout=300)
    assert one.returncode == 0, one.stdout + one.stderr
    want = re.search(r"checksum ([0-9a-f]{8})", one.stdout).group(1)
    procs = [subprocess.Popen(base + ["--world", "2", "--rank", str(r), "--id-file", str(tmp_path / "id2")], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    # 22 poses on 2 ranks = 11 per rank; the single-rank run gathers the same 22 frames in the same order
    got = [re.search(r"checksum ([0-9a-f]{8})", o).group(1) for o in outs]
    assert got == [want, want], (got, want, outs)
