"""ORACLE -- test infrastructure.  A scene handle built from the oracle's own WAD loader and scene compiler, with the
same duck type as rust_doom_b200.Scene (blob, sector_at, start_pose, info.n_segs ...), so that the pose generators
(rust-doom_b200/poses.py) and bench.py's `--impl reference` arm run without loading libb2d.so at all."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Tuple

import numpy as np

from . import scene as S
from . import wad as W
from .render import POSE


class OracleScene:
    def __init__(self, wad_bytes: bytes, level_index: int = 0):
        self.archive = W.Archive(wad_bytes)
        self.textures = W.TextureDirectory(self.archive)
        self.level = W.Level(self.archive, level_index)
        self.blob = S.compile_scene(self.archive, self.textures, level_index)
        h = S.header(self.blob)
        self.info = SimpleNamespace(n_segs=h[S.H_NSEGS], n_ssectors=h[S.H_NSSECTORS], n_sectors=h[S.H_NSECTORS],
                                    n_sprites=h[S.H_NSPRITES], n_masked_mids=h[S.H_NMIDS], blob_bytes=len(self.blob),
                                    has_start=h[S.H_HAS_START])
        self._hdr = h

    def sector_at(self, x: float, y: float) -> Tuple[int, int, int]:
        sec = S.sector_at(self.level, x, y)
        if sec < 0:
            return -1, 0, 0
        s = self.level.sectors[sec]
        return sec, int(s["floor"]), int(s["ceil"])

    @property
    def start_pose(self) -> Optional[np.ndarray]:
        h = self._hdr
        if not h[S.H_HAS_START]:
            return None

        def s32(v):
            return v - (1 << 32) if v & 0x80000000 else v
        p = np.zeros(1, dtype=POSE)
        p["x"], p["y"], p["z"] = s32(h[S.H_START_X]) * 65536, s32(h[S.H_START_Y]) * 65536, s32(h[S.H_START_Z]) * 65536
        p["angle"] = ((h[S.H_START_ANGLE] << 32) // 360) & 0xFFFFFFFF
        return p
