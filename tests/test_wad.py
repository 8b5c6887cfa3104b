"""WAD layer: the reference's own known-answer vectors (wad/src/name.rs:168-189) against both the oracle
restatement and the product (through the C ABI), plus container / picture edge cases."""
import struct

import numpy as np
import pytest

from oracle import wad as W

NAME_OK = [
    (b"", b"\0\0\0\0\0\0\0\0"), (b"\0", b"\0\0\0\0\0\0\0\0"), (b"\x001234567", b"\0\0\0\0\0\0\0\0"),
    (b"A", b"A\0\0\0\0\0\0\0"), (b"1234567", b"1234567\0"), (b"12345678", b"12345678"),
    (b"123\x005678", b"123\0\0\0\0\0"), (b"SKY1", b"SKY1\0\0\0\0"), (b"-", b"-\0\0\0\0\0\0\0"),
    (b"_", b"_\0\0\0\0\0\0\0"),
]
NAME_BAD = [b"123456789", b"1234\xfb", b"\xff123", b"$$ASDF_", b"123456789\0"]


@pytest.mark.parametrize("value,expect", NAME_OK)
def test_wad_name_vectors_oracle(value, expect):
    assert W.wad_name(value) == expect


@pytest.mark.parametrize("value", NAME_BAD)
def test_wad_name_rejects_oracle(value):
    with pytest.raises(W.WadError):
        W.wad_name(value)


@pytest.mark.parametrize("value,expect", NAME_OK)
def test_wad_name_vectors_product(b2d, value, expect):
    assert b2d.wad_name(value) == expect


@pytest.mark.parametrize("value", NAME_BAD)
def test_wad_name_rejects_product(b2d, value):
    with pytest.raises(b2d.B2dError) as e:
        b2d.wad_name(value)
    assert e.value.code == b2d.ERR_CORRUPT_WAD


def test_lowercase_is_uppercased(b2d):
    assert W.wad_name(b"sky1") == b"SKY1\0\0\0\0"
    assert b2d.wad_name(b"sky1") == b"SKY1\0\0\0\0"


def test_record_sizes():
    assert (W.THING.itemsize, W.VERTEX.itemsize, W.LINEDEF.itemsize, W.SIDEDEF.itemsize, W.SECTOR.itemsize,
            W.SUBSECTOR.itemsize, W.SEG.itemsize, W.NODE.itemsize) == (10, 4, 14, 30, 26, 4, 12, 28)


def test_archive_levels(b2d, synth_wad):
    oa = W.Archive(synth_wad)
    pa = b2d.Archive.from_bytes(synth_wad)
    assert oa.num_levels() == pa.num_levels() == 2
    assert [oa.level_name(i).rstrip(b"\0").decode() for i in range(2)] == pa.level_names() == ["E1M1", "E1M2"]


def test_pwad_rejected(b2d, synth_wad):
    bad = b"PWAD" + synth_wad[4:]
    with pytest.raises(W.WadError):
        W.Archive(bad)
    with pytest.raises(b2d.B2dError) as e:
        b2d.Archive.from_bytes(bad)
    assert e.value.code == b2d.ERR_CORRUPT_WAD


def test_truncated_and_empty(b2d):
    for data in (b"", b"IWAD", b"IWAD" + struct.pack("<ii", 5, 1000)):
        with pytest.raises(W.WadError):
            W.Archive(data)
        with pytest.raises(b2d.B2dError):
            b2d.Archive.from_bytes(data)


def test_invalid_directory_name_fails_open(b2d, synth_wad):
    # name.rs:41-75 + archive.rs:82-83: one bad byte in any directory name fails the whole open
    ident, n, off = struct.unpack_from("<4sii", synth_wad, 0)
    bad = bytearray(synth_wad)
    bad[off + 16 * 3 + 8] = ord("$")
    with pytest.raises(W.WadError):
        W.Archive(bytes(bad))
    with pytest.raises(b2d.B2dError):
        b2d.Archive.from_bytes(bytes(bad))


def test_missing_file_is_io_error(b2d):
    with pytest.raises(b2d.B2dError) as e:
        b2d.Archive.open("/nonexistent/doom1.wad")
    assert e.value.code == b2d.ERR_IO


def test_level_index_out_of_range(b2d, synth_wad):
    a = b2d.Archive.from_bytes(synth_wad)
    with pytest.raises(b2d.B2dError):
        b2d.Scene(a, 7)
    with pytest.raises(W.WadError):
        W.Level(W.Archive(synth_wad), 7)


def test_bad_lump_size_is_corrupt(b2d, synth_wad):
    # archive.rs:178-181: size must be a non-zero multiple of the record size
    ident, n, off = struct.unpack_from("<4sii", synth_wad, 0)
    oa = W.Archive(synth_wad)
    idx = oa.levels[0] + 4          # VERTEXES
    bad = bytearray(synth_wad)
    pos, size = struct.unpack_from("<ii", bad, off + 16 * idx)
    struct.pack_into("<i", bad, off + 16 * idx + 4, size - 1)
    with pytest.raises(W.WadError):
        W.Level(W.Archive(bytes(bad)), 0)
    with pytest.raises(b2d.B2dError) as e:
        b2d.Scene(b2d.Archive.from_bytes(bytes(bad)), 0)
    assert e.value.code == b2d.ERR_CORRUPT_WAD


def test_picture_roundtrip_and_posts():
    from rust_doom_b200 import synthwad
    img = np.full((20, 5), -1, dtype=np.int16)
    img[2:7, 0] = 10
    img[9:12, 0] = 11            # two posts in one column
    img[:, 2] = 200              # full column
    img[19, 4] = 7               # single pixel at the bottom
    buf = synthwad.encode_picture(img, 3, -4)
    px, xo, yo = W.decode_picture(buf)
    assert (xo, yo) == (3, -4) and px.shape == (20, 5)
    expect = np.where(img < 0, 0xFFFF, img).astype(np.uint16)
    assert np.array_equal(px, expect)
    with pytest.raises(W.WadError):
        W.decode_picture(buf[:30])


def test_blit_clipping_and_mask():
    dest = np.full((8, 8), W.TRANSPARENT_NEW, dtype=np.uint16)
    src = np.arange(16, dtype=np.uint16).reshape(4, 4)
    src[1, 1] = 0xFFFF
    W.blit(dest, src, -2, -1, True)       # clipped top-left, opaque copy incl. the transparent texel
    assert dest[0, 0] == src[1, 2] and dest[0, 1] == src[1, 3] and dest[2, 1] == src[3, 3]
    d2 = np.zeros((8, 8), dtype=np.uint16)
    W.blit(d2, src, 6, 6, False)          # clipped bottom-right, masked
    assert d2[6, 6] == src[0, 0] and d2[7, 7] == 0      # src[1,1] is transparent -> dest kept
    d3 = np.zeros((4, 4), dtype=np.uint16)
    W.blit(d3, src, 4, 0, False)          # fully out of bounds
    W.blit(d3, src, -9, 0, False)
    assert not d3.any()


def _encode_tall_picture(img):
    """Picture lump for an image taller than 254 rows, DeePsea style: a post whose topdelta is not above the previous
    post's is relative to it (runs of at most 100 rows; the second post of a column at absolute row 300 is written as
    topdelta 300 - 200 = 100 <= 200)."""
    import struct
    h, w = img.shape
    cols = []
    for x in range(w):
        out = bytearray()
        last = -1
        for y0 in range(0, h, 100):
            run = bytes(int(v) & 0xFF for v in img[y0:min(h, y0 + 100), x])
            if y0 <= 254 and y0 > last:
                top = y0                     # absolute while it fits and increases
            else:
                top = y0 - last              # relative to the previous post
                assert 0 < top <= last and top < 255
            out += bytes([top, len(run), 0]) + run + b"\0"
            last = y0
        out += b"\xff"
        cols.append(bytes(out))
    offs, pos = [], 8 + 4 * w
    for c in cols:
        offs.append(pos)
        pos += len(c)
    return struct.pack("<HHhh", w, h, 0, 0) + struct.pack("<%dI" % w, *offs) + b"".join(cols)


def test_pwad_overlay_and_tall_patches(b2d):
    """IWAD + PWAD (wad/src/archive.rs:69-72 rejects PWADs): a PWAD replaces E1M1, adds a level, overrides a
    flat between FF_START/FF_END, brings its own PNAMES / TEXTURE1 with a 64x420 texture built from a tall patch
    (relative posts).  Product and oracle loaders agree: level list, compiled scenes byte for byte, decoded tall texture."""
    import struct
    from oracle import scene as S
    from rust_doom_b200 import synthwad
    iwad = synthwad.build_iwad(1, ("E1M1", "E1M2"))
    donor = W.Archive(synthwad.build_iwad(7, ("E1M1",)))
    donor2 = W.Archive(synthwad.build_iwad(8, ("E1M1",)))

    def level_lumps(arch, new_name, rename=None):
        m = arch.levels[0]
        out = [(new_name, b"")]
        for k in range(1, 11):
            name, _, _ = arch.lumps[m + k]
            if name in (b"THINGS\0\0", b"LINEDEFS", b"SIDEDEFS", b"VERTEXES", b"SEGS\0\0\0\0", b"SSECTORS", b"NODES\0\0\0", b"SECTORS\0",
                        b"REJECT\0\0", b"BLOCKMAP"):
                data = arch.read(m + k)
                if rename and name == b"SIDEDEFS":
                    for old, new in rename.items():
                        data = data.replace(old, new)
                out.append((name.rstrip(b"\0").decode(), data))
        return out

    rng = np.random.default_rng(5)
    tall = rng.integers(1, 250, (420, 64)).astype(np.int16)
    base = W.Archive(iwad)
    pn = base.read(base.required(b"PNAMES"))
    npn = struct.unpack_from("<I", pn, 0)[0]
    pnames = struct.pack("<I", npn + 1) + pn[4:] + b"TALLP\0\0\0"
    t1 = base.read(base.required(b"TEXTURE1"))
    ntex = struct.unpack_from("<I", t1, 0)[0]
    offs = list(struct.unpack_from("<%dI" % ntex, t1, 4))
    bodies = t1[4 + 4 * ntex:]
    newtex = struct.pack("<8sIHHIH", b"TALLTEX\0", 0, 64, 420, 0, 1) + struct.pack("<hhHHH", 0, 0, npn, 1, 0)
    offs = [o + 4 for o in offs] + [4 + 4 * (ntex + 1) + len(bodies)]
    texture1 = struct.pack("<I", ntex + 1) + struct.pack("<%dI" % (ntex + 1), *offs) + bodies + newtex
    new_floor = bytes(rng.integers(0, 256, 4096, dtype=np.uint8))
    lumps = (level_lumps(donor, "E1M1", rename={b"BRICK1\0\0": b"TALLTEX\0"}) + level_lumps(donor2, "E1M9")
             + [("PNAMES", pnames), ("TEXTURE1", texture1), ("PP_START", b""), ("TALLP", _encode_tall_picture(tall)), ("PP_END", b""),
                ("FF_START", b""), ("FLOOR1", new_floor), ("FF_END", b"")])
    pwad = synthwad.assemble_wad(lumps, ident=b"PWAD")

    oa = W.Archive(iwad, overlays=(pwad,))
    pa = b2d.Archive.from_bytes(iwad, overlays=(pwad,))
    assert [n.rstrip(b"\0").decode() for n in (oa.level_name(i) for i in range(oa.num_levels()))] == ["E1M1", "E1M2", "E1M9"]
    assert pa.level_names() == ["E1M1", "E1M2", "E1M9"]
    otd = W.TextureDirectory(oa)
    assert np.array_equal(otd.textures[b"TALLTEX\0"], tall.astype(np.uint16))       # tall patch decoded, relative posts and all
    assert otd.flats[b"FLOOR1\0\0"] == new_floor
    uses_tall = False
    for lvl in range(3):
        ob = S.compile_scene(oa, otd, lvl)
        assert b2d.Scene(pa, lvl).blob == ob, "level %d" % lvl
        if lvl == 0:
            uses_tall = any(int(t[2]) == 420 for t in S.section(ob, "textures"))
    assert uses_tall, "the replaced E1M1 does not use the tall texture"
    # E1M1 is the PWAD's level now (E1M2 is still the IWAD's own)
    assert b2d.Scene(pa, 0).info.n_segs == b2d.Scene(b2d.Archive.from_bytes(synthwad.build_iwad(7, ("E1M1",))), 0).info.n_segs
    # an overlay must be a PWAD, the base an IWAD
    with pytest.raises(b2d.B2dError):
        b2d.Archive.from_bytes(iwad, overlays=(iwad,))
    with pytest.raises(W.WadError):
        W.Archive(iwad, overlays=(iwad,))
    with pytest.raises(b2d.B2dError):
        b2d.Archive.from_bytes(pwad)
