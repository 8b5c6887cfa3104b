#!/usr/bin/env python3
"""Random parity campaign without a GPU: the product's maths and fetch paths executed on the CPU (tests/hostcheck)
against the oracle, on random generated levels (all content kinds), random resolutions, fields of view and level
times.  Not part of the test-suite (minutes); the last run (1200 cases, 0 mismatches) is quoted in DESIGN.md.
usage: python tools/campaign.py [cases]"""
import sys, ctypes, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rust_doom_b200 as b2d
from rust_doom_b200 import synthwad
from oracle import render, scene as oscene, wad as owad
from tests.refcheck import moves as MV
from tests.conftest import sample_poses
import subprocess, os
so=os.path.join(ROOT,'tests','hostcheck','libb2d_hostcheck.so')
subprocess.check_call(["g++","-O2","-std=c++17","-fPIC","-shared","-o",so,os.path.join(ROOT,"tests","hostcheck","hostcheck.cpp")])
lib=ctypes.CDLL(so)
def host(blob, view, poses, tics, moves=()):
    n=len(poses); fb=np.empty((n,view.height,view.width),np.uint8); counts=np.zeros(n,np.int32)
    buf=(ctypes.c_char*len(blob)).from_buffer_copy(blob); poses=np.ascontiguousarray(poses)
    mv=np.array([tuple(int(v) for v in m) for m in moves],dtype=np.int32).reshape(-1,3)
    rc=lib.hostcheck_render_m(ctypes.c_void_p(ctypes.addressof(buf)), ctypes.byref(view), ctypes.c_void_p(poses.ctypes.data), n, ctypes.c_void_p(fb.ctypes.data), ctypes.c_void_p(counts.ctypes.data), None, 0, ctypes.c_uint32(tics), ctypes.c_void_p(mv.ctypes.data if len(mv) else None), len(mv))
    assert rc==0
    return fb
rng=np.random.default_rng(int(os.environ.get("B2D_CAMPAIGN_SEED", "12345")))
bad=0; t0=time.time()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 200):
    seed=int(rng.integers(100,100000))
    cfg=synthwad.SynthConfig(mid_pct=int(rng.integers(0,50)), thing_pct=int(rng.integers(0,60)), anim=bool(rng.integers(0,2)), odd_tex=bool(rng.integers(0,2)),
                             rock_pct=int(rng.integers(5,30)), sky_pct=int(rng.integers(0,40)), door_pct=int(rng.integers(5,40)))
    name=["E1M1","E2M3","MAP05","MAP15","MAP25"][int(rng.integers(0,5))]
    data=synthwad.build_iwad(seed,(name,),cfg=cfg)
    dyn,moves,oblob=(),(),None
    if it%3==2:      # every third case: some sectors declared dynamic and moved (C16); the oracle renders its own scene
        oa=owad.Archive(data); level=owad.Level(oa,0)
        dyn=MV.declare(level,seed,10); moves=MV.state(level,dyn,seed+1,hole_free=False)
        oblob=oscene.apply_moves(oscene.compile_scene(oa,owad.TextureDirectory(oa),0,dynamic=dyn),moves)
    sc=b2d.Scene(b2d.Archive.from_bytes(data),0,dynamic=dyn)
    w=int(rng.integers(40,1300)); h=int(rng.integers(30,900)); tics=int(rng.integers(0,1<<32)) if rng.integers(0,2) else 0
    fov=float(rng.uniform(40,110))
    poses=sample_poses(b2d, sc, 4, seed)
    ov=render.make_view(w,h,fov); pv=b2d.make_view(w,h,fov)
    o=render.render(oblob if oblob is not None else sc.blob, ov, poses, threads=4, tics=tics); hfb=host(sc.blob, pv, poses, tics, moves)
    if not np.array_equal(o,hfb):
        bad+=1; print("MISMATCH", seed, cfg, name, w, h, tics, fov, moves, int((o!=hfb).sum()))
print("done", it+1, "bad", bad, "time", round(time.time()-t0,1))
