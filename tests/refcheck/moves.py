"""Random sector movements for the parity tests: which sectors are declared dynamic (with the height ranges a
LevelAnalysis would have found) and one state of them.  `reference_shows_no_hole` tells whether the reference's meshes
cover every opening in that state: it pre-extends lower quads over the floor ranges and one-sided walls over the sector's
whole range, but builds an upper quad only where the back ceiling is lower at rest and does not extend it
(wad/src/visitor.rs:791-807), so a ceiling that drops below a neighbour's, or a ceiling that rises above the top of the
upper quad in front of it, opens a hole there -- the one place where the column renderer (which closes it with the
upper texture) and the reference differ; the ray-caster comparison only uses states without such holes."""
import numpy as np

from oracle import wad as W


def two_sided_segs(level):
    """(front sector, back sector, back ceiling is sky, seg index) of every two-sided seg"""
    out = []
    sec_bytes = level.sectors.tobytes()
    for i, sg in enumerate(level.segs):
        side = level.seg_sidedef_index(sg)
        back_side = level.seg_back_sidedef_index(sg)
        if side < 0 or back_side < 0:
            continue
        f, b = int(level.sidedefs[side]["sector"]), int(level.sidedefs[back_side]["sector"])
        if f >= len(level.sectors) or b >= len(level.sectors):
            continue
        out.append((f, b, W.is_sky_flat(W.wad_name(sec_bytes[b * 26 + 12:b * 26 + 20])), i))
    return out


def reference_shows_no_hole(level, moves):
    df = {int(m[0]): int(m[1]) for m in moves}
    dc = {int(m[0]): int(m[2]) for m in moves}
    secs = level.sectors
    for f, b, back_sky, _ in two_sided_segs(level):
        if back_sky:
            continue
        fc, bc = int(secs[f]["ceil"]), int(secs[b]["ceil"])
        fc1, bc1 = fc + dc.get(f, 0), bc + dc.get(b, 0)
        if bc1 < fc1:                                   # an upper piece is visible in this state
            if not bc < fc:                             # ... but the reference never built the quad
                return False
            if fc + dc.get(b, 0) < fc1:                 # ... or the quad (moving with the back ceiling) ends below the ceiling
                return False
    return True


def declare(level, seed, n_sectors=6):
    """n_sectors random sectors with +-48 unit height ranges (sky ceilings stay put: an open-air sector has none to move)"""
    rng = np.random.default_rng(seed)
    secs = level.sectors
    sec_bytes = secs.tobytes()
    dynamic = []
    for s in rng.choice(len(secs), size=min(n_sectors, len(secs)), replace=False):
        f0, c0 = int(secs[s]["floor"]), int(secs[s]["ceil"])
        fmin, fmax = f0 - int(rng.integers(0, 49)), f0 + int(rng.integers(0, 49))
        cmin, cmax = c0 - int(rng.integers(0, 49)), c0 + int(rng.integers(0, 49))
        if W.is_sky_flat(W.wad_name(sec_bytes[s * 26 + 12:s * 26 + 20])):
            cmin = cmax = c0
        dynamic.append((int(s), fmin, fmax, cmin, cmax))
    return dynamic


def state(level, dynamic, seed, hole_free=True):
    """One random state of the declared sectors.  Floors move freely inside their range; a ceiling keeps the first of 20
    random heights that opens no hole in the reference (any height if not hole_free), else stays."""
    rng = np.random.default_rng(seed)
    secs = level.sectors
    moves = []
    for (s, fmin, fmax, cmin, cmax) in dynamic:
        f0, c0 = int(secs[s]["floor"]), int(secs[s]["ceil"])
        move = (s, 0, 0)
        for _ in range(20):
            f1, c1 = int(rng.integers(fmin, fmax + 1)), int(rng.integers(cmin, cmax + 1))
            cand = (s, min(f1, c1) - f0, c1 - c0)
            if not hole_free or reference_shows_no_hole(level, moves + [cand]):
                move = cand
                break
        moves.append(move)
    return moves


def pick(level, seed, n_sectors=6):
    """-> (dynamic, moves): declare(seed) and one hole-free state of it"""
    dynamic = declare(level, seed, n_sectors)
    moves = state(level, dynamic, seed + 1000)
    assert any(m[1] or m[2] for m in moves)
    return dynamic, moves
