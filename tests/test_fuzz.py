"""Corrupt-WAD tolerance (SURVEY.md 5 'failure detection'): random byte corruption of a valid IWAD must end in
a clean error or a valid scene -- never a crash -- and whenever both the oracle's and the product's loaders
accept the file they must compile the same scene.  Runs in a subprocess so that a native crash is a test
failure, not the end of the pytest run."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import sys, struct
sys.path.insert(0, %(root)r)
import numpy as np
import rust_doom_b200 as b2d
from rust_doom_b200 import synthwad
from oracle import wad as W, scene as S
base = synthwad.build_iwad(1, ("E1M1",), cfg=synthwad.SynthConfig(gx=4, gy=3, origin=(-512, -384)))
rng = synthwad.SplitMix64(%(seed)d)
ident, nl, off = struct.unpack_from("<4sii", base, 0)
agree = both_ok = disagree = dyn_ok = moved_ok = 0
for it in range(%(iters)d):
    data = bytearray(base)
    mode = rng.below(4)
    if mode == 0:      # flip bytes anywhere
        for _ in range(1 + rng.below(8)):
            data[rng.below(len(data))] = rng.below(256)
    elif mode == 1:    # corrupt a directory entry (pos/size/name)
        e = off + 16 * rng.below(nl)
        for _ in range(1 + rng.below(3)):
            data[e + rng.below(16)] = rng.below(256)
    elif mode == 2:    # corrupt inside a level lump
        oa = W.Archive(base)
        idx = oa.levels[0] + 1 + rng.below(8)
        name, pos, size = oa.lumps[idx]
        if size:
            for _ in range(1 + rng.below(6)):
                data[pos + rng.below(size)] = rng.below(256)
    else:              # truncate
        data = data[:rng.below(len(data))]
    data = bytes(data)
    ob = pb = None
    try:
        oa = W.Archive(data); ob = S.compile_scene(oa, W.TextureDirectory(oa), 0)
    except W.WadError:
        pass
    try:
        pa = b2d.Archive.from_bytes(data); pb = b2d.Scene(pa, 0).blob
    except b2d.B2dError as e:
        assert e.code in (b2d.ERR_CORRUPT_WAD, b2d.ERR_IO, b2d.ERR_INVALID_ARG), e
    if (ob is None) == (pb is None):
        agree += 1
        if ob is not None:
            both_ok += 1
            assert ob == pb, "iteration %%d: both loaders accept the file but compile different scenes" %% it
            # the same corrupted level with a random dynamic-sector list and a random state: errors are fine, crashes and
            # disagreements are not (DESIGN.md C16: host-side table re-derivation on untrusted lumps)
            nsec = S.header(ob)[S.H_NSECTORS]
            dyn, mv = [], []
            for _ in range(1 + rng.below(4)):
                sec = rng.below(nsec + 2) - 1                # now and then out of range
                r4 = [rng.below(400) - 200 for _ in range(4)]
                dyn.append((sec, r4[0], r4[1], r4[2], r4[3]))
                mv.append((sec, rng.below(100) - 50, rng.below(100) - 50))
            od = pd = None
            try:
                od = S.compile_scene(oa, W.TextureDirectory(oa), 0, dynamic=dyn)
            except W.WadError:
                pass
            try:
                psc = b2d.Scene(pa, 0, dynamic=dyn); pd = psc.blob
            except b2d.B2dError as e:
                assert e.code in (b2d.ERR_CORRUPT_WAD, b2d.ERR_INVALID_ARG), e
            assert (od is None) == (pd is None) and od == pd, "iteration %%d: dynamic scene" %% it
            if od is not None:
                dyn_ok += 1
                om = pm = None
                try:
                    om = S.apply_moves(od, mv)
                except ValueError:
                    pass
                try:
                    pm = psc.tables_at(0, mv)
                except b2d.B2dError as e:
                    assert e.code == b2d.ERR_INVALID_ARG, e
                assert (om is None) == (pm is None), "iteration %%d: moves accepted by one side only" %% it
                if om is not None:
                    moved_ok += 1
                    h = S.header(om)
                    seg0 = h[S.H_NTEX] * 32 + h[S.H_NSECTORS] * 32
                    assert pm[seg0:seg0 + h[S.H_NSEGS] * 64] == om[h[S.H_OFF_SEGS]:h[S.H_OFF_SEGS] + h[S.H_NSEGS] * 64], "iteration %%d: moved segs" %% it
    else:
        disagree += 1
print("RESULT", agree, both_ok, disagree, dyn_ok, moved_ok)
'''


def _run(seed, iters):
    code = WORKER % dict(root=ROOT, seed=seed, iters=iters)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, "loader crashed or asserted:\n" + p.stdout[-2000:] + p.stderr[-4000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT")][-1]
    return [int(v) for v in line.split()[1:]]


def test_fuzzed_wads_never_crash_and_agree():
    agree, both_ok, disagree, dyn_ok, moved_ok = _run(1234, 300)
    assert dyn_ok > 5 and moved_ok > 0, (dyn_ok, moved_ok)      # the dynamic-sector paths were exercised too
    assert both_ok > 20            # many corruptions are benign (texture bytes, unused lumps)
    assert disagree == 0, "oracle and product disagree on accept/reject: %d" % disagree
