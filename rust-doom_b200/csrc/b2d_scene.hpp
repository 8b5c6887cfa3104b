// Scene compiler: raw level lumps + texture directory -> one flat "B2DS" blob that is uploaded to
// HBM as-is and indexed by the kernels (layout: DESIGN.md "Scene blob"; records below).
//
// What is pre-resolved here is the reference's geometry emitter, per seg instead of per triangle:
//   which wall pieces a seg has, their texture, pegging anchor and offsets   wad/src/visitor.rs:711-937
//   fake contrast and the static light byte   visitor.rs:887-901, wad/src/light.rs:27-115,
//                                              game/src/lights.rs:14-30
//   flats / sky flats per sector               visitor.rs:939-985
//   sky texture of the level                   wad/src/meta.rs:156-172, assets/meta/doom.toml:29-68
//   player-1 start                             visitor.rs:1010-1060, game/src/level.rs:757-762
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "b2d_wad.hpp"

namespace b2d {

constexpr uint32_t kSceneMagic = 0x53443242u;   // "B2DS"
constexpr uint32_t kSceneVersion = 6;
constexpr uint32_t kLeaf = 0x80000000u;

enum HeaderField : int {
    H_MAGIC, H_VERSION, H_TOTAL, H_NVERTS, H_NNODES, H_NSSECTORS, H_NSEGS, H_NSECTORS, H_NTEX, H_NFLATS,
    H_OFF_VERTS, H_OFF_NODES, H_OFF_SSECTORS, H_OFF_SEGS, H_OFF_SECTORS, H_OFF_TEX, H_OFF_TEXELS,
    H_TEXEL_BYTES, H_OFF_FLATS, H_OFF_COLORMAP, H_OFF_PALETTE, H_ROOT, H_SKY_TEX, H_START_X, H_START_Y,
    H_START_Z, H_START_ANGLE, H_HAS_START, H_MIN_H, H_MAX_H, H_NMIDS, H_OFF_MIDS, H_NSPRITES, H_OFF_SPRITES,
    H_NANIM, H_OFF_ANIM, H_OFF_FLAT_ANIM, H_OFF_LIGHTS, H_NDYN, H_OFF_DYN, H_OFF_SEGDYN, H_COUNT = 64
};

// 64-byte records; all int32.
struct NodeRec { int32_t x, y, dx, dy, rbox[4], lbox[4]; uint32_t child[2]; int32_t pad[2]; };  // child[0]=right
struct SSectorRec { int32_t first_seg, num_segs, sector, sprites; };   // sprites = first | count << 24
// decoration thing: billboard of its sprite image's size, centred on (x,y), bottom edge at `low`
// (floor, or ceiling - height for hanging things; visitor.rs:1062-1137), lit by the sector light
struct SpriteRec { int32_t x, y, low, tex, light, sector, hanging, pad; };
struct SegRec {
    int32_t v1, v2, front, flags;
    int32_t uoff, len_q12;
    int32_t texA, tA, hA;      // upper (two-sided) or full-height middle (one-sided)
    int32_t texB, tB, hB;      // lower
    int32_t light, otop, obot, mid;   // mid: index into the mids section, -1 = no masked middle texture
};
// masked two-sided middle texture (visitor.rs:808-836,875-919): vertical extent [low, high) and the
// texture row at `high` (pegging and y offset folded in)
struct MidRec { int32_t tex, t_high, low, high, pad[4]; };
struct SectorRec { int32_t floor, ceil, floor_flat, ceil_flat, light, pad[3]; };
// mask_off = ~0u: opaque.  anim_nk = n | k << 16: frame k of an n-frame animation whose ids are anim[anim_first..+n)
struct TexRec { uint32_t texel_off, w, h, hmagic, hbias, mask_off, anim_first, anim_nk; };
struct FlatAnimRec { int32_t anim_first, anim_nk; };
// sector light effect (wad/src/light.rs:5-25): kind 0 none, 1 glow, 2 random, 3 alternate
struct LightRec { uint32_t kind; float level, alt, speed, duration, sync; uint32_t pad[2]; };
constexpr uint32_t kLightNone = 0, kLightGlow = 1, kLightRandom = 2, kLightAlternate = 3;
// what scene_at_state needs beyond the seg record: back sector (-1 = one-sided) and pegging / sky bits
struct SegDynRec { int32_t back, bits; };
constexpr int32_t kSegDynUnpegLower = 1, kSegDynBackSky = 2;
// a sector that may move, with the height ranges the host's LevelAnalysis found (visitor.rs:146-245)
struct DynRec { int32_t sector, floor_min, floor_max, ceil_min, ceil_max, pad[3]; };
// one state of a moving sector: offsets in map units relative to the heights in the level lumps
struct SectorMove { int32_t sector, floor_offset, ceil_offset; };
static_assert(sizeof(SegDynRec) == 8 && sizeof(DynRec) == 32 && sizeof(SectorMove) == 12, "record layout");
static_assert(sizeof(NodeRec) == 64 && sizeof(SegRec) == 64 && sizeof(SectorRec) == 32 &&
              sizeof(TexRec) == 32 && sizeof(SSectorRec) == 16 && sizeof(MidRec) == 32 &&
              sizeof(SpriteRec) == 32 && sizeof(LightRec) == 32, "record layout");

constexpr int32_t kSegTwoSided = 1, kSegScroll = 2, kSegInvalid = 0x80;   // scroll: special 0x30 (visitor.rs:922)
constexpr int32_t kFlatSky = -1, kFlatMissing = -2, kTexNone = -1;

// Level time (static.vert:23-39, visitor.rs:922).  `tics` counts 1/35 s.  Every frame name of an
// n-frame animation group is bound to the group's first frame (tex.rs:260, 302-306), so an animated image shows
// group frame (tics/8) mod n whichever frame name the map uses (frame 0 at tic 0: the tables are resolved at
// renderer creation too); walls of a scrolling line (special 0x30) advance their texture column by one texel per tic.  The kernels never see time: the three small tables that
// depend on it (texture records, sector flats, seg column offsets) are re-derived from the blob here and
// re-uploaded when the time changes.  Sector light effects (C15) ride on the same mechanism: the light bytes of
// effect sectors, their segs and their sprites are re-evaluated.  Outputs hold H_NTEX / H_NSECTORS / H_NSEGS /
// H_NSPRITES records.
// Light level of a sector at `tics` (game/src/lights.rs:26-66), float32 operation by operation (volatile keeps
// every intermediate a rounded float32: no x87 excess precision, no fused multiply-add).  The sine of the random
// effect's hash is the correctly rounded float32 sine: double-precision sine, rounded once.
inline uint8_t light_byte_at(const LightRec &L, uint32_t tics) {
    volatile float time = (float)tics / 35.0f;
    volatile float v = L.level;
    auto fract = [](float x) -> float { volatile float f = std::floor(x); volatile float r = x - f; return r; };
    if (L.kind == kLightGlow) {
        volatile float scale = L.level - L.alt;
        volatile float ts = time * L.speed;
        volatile float phase = ts / scale;
        volatile float d = 0.5f - fract(phase);
        volatile float a = std::fabs(d);
        volatile float a2 = a * 2.0f;
        volatile float a3 = a2 * scale;
        v = a3 + L.alt;
    } else if (L.kind == kLightRandom) {
        volatile float ts = time * L.speed;
        volatile float t = std::floor(ts);
        volatile float t1 = t / 1000.0f;
        volatile float s1 = L.sync + t1;
        volatile float s2 = s1 * 12.9898f;
        volatile float s3 = L.sync * 78.233f;
        volatile float arg = s2 + s3;
        volatile float sn = (float)std::sin((double)arg);
        volatile float n1 = sn * 43758.547f;
        volatile float n2 = 1.0f + n1;
        v = fract(n2) < L.duration ? L.alt : L.level;
    } else if (L.kind == kLightAlternate) {
        volatile float ts = time * L.speed;
        volatile float s3 = L.sync * 3.5435f;
        volatile float ph = ts + s3;
        v = fract(ph) < L.duration ? L.alt : L.level;
    }
    if (v < 0.0f) v = 0.0f; else if (v > 1.0f) v = 1.0f;
    volatile float scaled = v * 255.0f;
    return scaled >= 0.0f ? (uint8_t)(int)scaled : 0;
}

inline bool scene_is_timed(const uint8_t *blob) {
    const uint32_t *h = reinterpret_cast<const uint32_t *>(blob);
    if (h[H_NANIM] > 0 || h[H_NDYN] > 0) return true;      // sectors that may move: their tables are state too
    const LightRec *lights = reinterpret_cast<const LightRec *>(blob + h[H_OFF_LIGHTS]);
    for (uint32_t i = 0; i < h[H_NSECTORS]; i++)
        if (lights[i].kind != kLightNone) return true;
    const SegRec *segs = reinterpret_cast<const SegRec *>(blob + h[H_OFF_SEGS]);
    for (uint32_t i = 0; i < h[H_NSEGS]; i++)
        if (segs[i].flags & kSegScroll) return true;
    return false;
}

// `floor_off` / `ceil_off`: nullptr, or one offset per sector (map units) -- the state of the moving sectors.
// The reference attaches every wall quad, flat and decoration to the floor or the ceiling object of a sector and
// translates it rigidly with that object (visitor.rs:733-836 object_id, 957-983, 1106-1121; game/src/level.rs:201-245):
// one-sided wall and masked middle texture -> own floor if the line is lower-unpegged, else own ceiling; upper piece ->
// back ceiling; lower piece -> back floor; decoration -> floor, or ceiling if it hangs.  The anchors of the pre-resolved
// pieces move accordingly and the opening of a two-sided seg follows the moved heights (what the depth test leaves
// visible of the reference's pre-extended quads).  `mids_out` holds H_NMIDS records.
inline void scene_at_time(const uint8_t *blob, uint32_t tics, TexRec *tex_out, SectorRec *sectors_out, SegRec *segs_out,
                          SpriteRec *sprites_out, MidRec *mids_out = nullptr, const int32_t *floor_off = nullptr,
                          const int32_t *ceil_off = nullptr) {
    const uint32_t *h = reinterpret_cast<const uint32_t *>(blob);
    const TexRec *tex = reinterpret_cast<const TexRec *>(blob + h[H_OFF_TEX]);
    const SectorRec *sectors = reinterpret_cast<const SectorRec *>(blob + h[H_OFF_SECTORS]);
    const SegRec *segs = reinterpret_cast<const SegRec *>(blob + h[H_OFF_SEGS]);
    const int32_t *anim = reinterpret_cast<const int32_t *>(blob + h[H_OFF_ANIM]);
    const FlatAnimRec *fa = reinterpret_cast<const FlatAnimRec *>(blob + h[H_OFF_FLAT_ANIM]);
    const uint32_t ntex = h[H_NTEX], nflats = h[H_NFLATS], nanim = h[H_NANIM];
    auto now = [&](int64_t first, uint32_t nk, int32_t self) -> int32_t {
        const uint32_t n = nk & 0xFFFFu;
        if (n < 2 || first < 0 || first + n > nanim) return self;
        return anim[first + (int64_t)((tics >> 3) % n)];
    };
    for (uint32_t i = 0; i < ntex; i++) {
        const int32_t j = now((int32_t)tex[i].anim_first, tex[i].anim_nk, (int32_t)i);
        tex_out[i] = (j >= 0 && (uint32_t)j < ntex) ? tex[j] : tex[i];
    }
    auto flat_now = [&](int32_t f) -> int32_t {
        if (f < 0 || (uint32_t)f >= nflats) return f;
        const int32_t j = now(fa[f].anim_first, (uint32_t)fa[f].anim_nk, f);
        return (j >= 0 && (uint32_t)j < nflats) ? j : f;
    };
    const LightRec *lights = reinterpret_cast<const LightRec *>(blob + h[H_OFF_LIGHTS]);
    const SpriteRec *sprites = reinterpret_cast<const SpriteRec *>(blob + h[H_OFF_SPRITES]);
    const uint32_t nsect = h[H_NSECTORS];
    for (uint32_t i = 0; i < nsect; i++) {
        sectors_out[i] = sectors[i];
        sectors_out[i].floor_flat = flat_now(sectors[i].floor_flat);
        sectors_out[i].ceil_flat = flat_now(sectors[i].ceil_flat);
        if (lights[i].kind != kLightNone) sectors_out[i].light = light_byte_at(lights[i], tics);
    }
    // walls and sprites of a sector with a light effect carry the sector's light (no fake contrast, visitor.rs:889)
    for (uint32_t i = 0; i < h[H_NSEGS]; i++) {
        segs_out[i] = segs[i];
        if (segs[i].flags & kSegInvalid) continue;
        if (segs[i].flags & kSegScroll) segs_out[i].uoff = segs[i].uoff + (int32_t)(tics & 0xFFFFFFu);
        const uint32_t f = (uint32_t)segs[i].front;
        if (f < nsect && lights[f].kind != kLightNone) segs_out[i].light = sectors_out[f].light;
    }
    for (uint32_t i = 0; i < h[H_NSPRITES]; i++) {
        sprites_out[i] = sprites[i];
        const uint32_t f = (uint32_t)sprites[i].sector;
        if (f < nsect && lights[f].kind != kLightNone) sprites_out[i].light = sectors_out[f].light;
    }
    const MidRec *mids = reinterpret_cast<const MidRec *>(blob + h[H_OFF_MIDS]);
    if (mids_out)
        for (uint32_t i = 0; i < h[H_NMIDS]; i++) mids_out[i] = mids[i];
    if (!floor_off || !ceil_off) return;
    // ---- moving sectors
    const SegDynRec *segdyn = reinterpret_cast<const SegDynRec *>(blob + h[H_OFF_SEGDYN]);
    for (uint32_t i = 0; i < nsect; i++) {
        sectors_out[i].floor += floor_off[i];
        sectors_out[i].ceil += ceil_off[i];
    }
    for (uint32_t i = 0; i < h[H_NSEGS]; i++) {
        SegRec &S = segs_out[i];
        if (S.flags & kSegInvalid) continue;
        const uint32_t f = (uint32_t)S.front;
        if (f >= nsect) continue;
        const int32_t own = (segdyn[i].bits & kSegDynUnpegLower) ? floor_off[f] : ceil_off[f];
        if (!(S.flags & kSegTwoSided)) {
            S.hA += own;
            S.otop = sectors_out[f].ceil;
            S.obot = sectors_out[f].floor;
            continue;
        }
        const uint32_t b = (uint32_t)segdyn[i].back;
        if (b >= nsect) continue;
        S.hA += ceil_off[b];
        S.hB += floor_off[b];
        const int32_t ff = sectors_out[f].floor, fc = sectors_out[f].ceil, bf = sectors_out[b].floor, bc = sectors_out[b].ceil;
        S.otop = (bc < fc && !(segdyn[i].bits & kSegDynBackSky)) ? bc : fc;
        S.obot = bf > ff ? bf : ff;
        if (mids_out && S.mid >= 0 && (uint32_t)S.mid < h[H_NMIDS]) {
            mids_out[S.mid].low += own;
            mids_out[S.mid].high += own;
        }
    }
    for (uint32_t i = 0; i < h[H_NSPRITES]; i++) {
        const uint32_t f = (uint32_t)sprites_out[i].sector;
        if (f < nsect) sprites_out[i].low += sprites_out[i].hanging ? ceil_off[f] : floor_off[f];
    }
}

// Checks a list of sector moves against the scene's declared dynamic sectors and expands it to one floor and one ceiling
// offset per sector.  Returns nullptr, or the reason the list is rejected.
inline const char *expand_moves(const uint8_t *blob, const SectorMove *moves, size_t n, std::vector<int32_t> &floor_off,
                                std::vector<int32_t> &ceil_off) {
    const uint32_t *h = reinterpret_cast<const uint32_t *>(blob);
    const SectorRec *sectors = reinterpret_cast<const SectorRec *>(blob + h[H_OFF_SECTORS]);
    const DynRec *dyn = reinterpret_cast<const DynRec *>(blob + h[H_OFF_DYN]);
    floor_off.assign(h[H_NSECTORS], 0);
    ceil_off.assign(h[H_NSECTORS], 0);
    for (size_t k = 0; k < n; k++) {
        const DynRec *d = nullptr;
        for (uint32_t i = 0; i < h[H_NDYN]; i++)
            if (dyn[i].sector == moves[k].sector) { d = &dyn[i]; break; }
        if (!d) return "a moved sector was not declared dynamic when the scene was created";
        const int64_t f1 = (int64_t)sectors[d->sector].floor + moves[k].floor_offset;
        const int64_t c1 = (int64_t)sectors[d->sector].ceil + moves[k].ceil_offset;
        if (f1 < d->floor_min || f1 > d->floor_max || c1 < d->ceil_min || c1 > d->ceil_max || f1 > c1)
            return "a sector is moved outside its declared height range";
        floor_off[(size_t)d->sector] = moves[k].floor_offset;
        ceil_off[(size_t)d->sector] = moves[k].ceil_offset;
    }
    return nullptr;
}

std::vector<uint8_t> compile_scene(const Archive &wad, const TextureDirectory &tex, int level_index,
                                   const std::vector<DynRec> &dynamic = {});
std::vector<uint8_t> compile_scene(const Level &level, const TextureDirectory &tex, const std::vector<DynRec> &dynamic = {});

// LevelWalker::sector_at on the raw level (visitor.rs:1028-1060); -1 if outside.
int sector_at(const Level &level, double x, double y, int *subsector_out = nullptr);

// (light >> 3)/31 (+-2/31, clamped) * 255 truncated to u8, in the reference's float32 arithmetic.
uint8_t light_byte(int16_t light, int contrast);

}  // namespace b2d
