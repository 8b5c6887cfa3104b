// WAD container, record decoding, picture decoding and texture composition (host, C++17).
//
// Data semantics follow the reference's `wad/` crate (cristicbz/rust-doom):
//   archive + directory + level detection   wad/src/archive.rs:62-106
//   typed little-endian records             wad/src/types.rs:20-169, archive.rs:172-190
//   lump names                              wad/src/name.rs:41-75
//   level lump offsets                      wad/src/level.rs:13-20
//   picture (patch) format + blit           wad/src/image.rs:39-252
//   PNAMES / TEXTUREx / flats / sprites     wad/src/tex.rs:53-107,358-410,475-606
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace b2d {

struct WadError : std::runtime_error {
    int code;
    WadError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
constexpr int kErrCorrupt = -1, kErrIo = -2, kErrArg = -4;

using Name = std::array<uint8_t, 8>;
Name make_name(const uint8_t *bytes, size_t size);   // WadName::from_bytes; throws WadError
Name make_name(const char *literal);
std::string name_str(const Name &n);
inline bool is_untextured(const Name &n) { return n[0] == '-' && n[1] == 0; }   // util.rs:4-6
bool is_sky_flat(const Name &n);                                                 // util.rs:8-10

struct NameHash {
    size_t operator()(const Name &n) const {
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) v = v * 131 + n[i];
        return (size_t)v;
    }
};

struct Lump { Name name; int64_t pos; int64_t size; };

struct Thing { int16_t x, y, angle; uint16_t type, flags; };
struct Vertex { int16_t x, y; };
struct Linedef { uint16_t v1, v2, flags, special, tag; int16_t right, left; };
struct Sidedef { int16_t xoff, yoff; Name upper, lower, middle; uint16_t sector; };
struct Sector { int16_t floor, ceil; Name floor_tex, ceil_tex; int16_t light; uint16_t type, tag; };
struct Subsector { uint16_t num_segs, first_seg; };
struct Seg { uint16_t v1, v2, angle, linedef, direction, offset; };
struct Node { int16_t x, y, dx, dy; int16_t rbox[4], lbox[4]; uint16_t right, left; };

class Archive {
public:
    // files[0] is the IWAD (the reference opens nothing else: archive.rs:69-72); further files are PWADs applied in order,
    // Doom-engine style: their lumps are appended to the directory, so a later lump of the same name wins every by-name
    // lookup (PNAMES, TEXTUREx, patches, PLAYPAL ...), a level of an existing name replaces that level in place, new
    // level names are appended, and flats / sprites between a PWAD's own FF_START..FF_END / SS_START..SS_END (or
    // F_START / S_START) markers are added to the IWAD's.
    explicit Archive(std::vector<uint8_t> data);
    explicit Archive(std::vector<std::vector<uint8_t>> files);
    static Archive open(const std::string &path);
    static Archive open(const std::vector<std::string> &paths);
    // lump index ranges [first, last) between the start / end markers of every file that has them (the IWAD must)
    std::vector<std::pair<int, int>> marker_ranges(const char *start, const char *start2, const char *end, const char *end2) const;

    int num_levels() const { return (int)levels_.size(); }
    const Name &level_name(int level_index) const;
    int level_lump_index(int level_index) const;
    int find(const Name &name) const;              // -1 if absent; the later duplicate wins
    int require(const char *name) const;           // throws CorruptWad if absent
    const Lump &lump(int index) const;
    size_t num_lumps() const { return lumps_.size(); }
    const uint8_t *lump_data(int index) const;     // nullptr for virtual (size 0) lumps

private:
    void add_file(const std::vector<uint8_t> &file, bool iwad);
    std::vector<uint8_t> data_;
    std::vector<std::pair<int, int>> files_;       // [first lump, one past last lump) per file
    std::vector<Lump> lumps_;
    std::unordered_map<Name, int, NameHash> index_;
    std::vector<int> levels_;
};

struct RawLump { const uint8_t *data = nullptr; size_t size = 0; };

struct Level {
    Name name;
    std::vector<Thing> things;
    std::vector<Linedef> linedefs;
    std::vector<Sidedef> sidedefs;
    std::vector<Vertex> vertices;
    std::vector<Seg> segs;
    std::vector<Subsector> subsectors;
    std::vector<Node> nodes;
    std::vector<Sector> sectors;

    static Level load(const Archive &wad, int level_index);
    // the eight level lumps as raw bytes, in marker order: THINGS, LINEDEFS, SIDEDEFS, VERTEXES, SEGS, SSECTORS, NODES,
    // SECTORS (wad/src/level.rs:13-31) -- what a host that has already parsed the WAD hands over
    static Level from_lumps(const Name &name, const RawLump lumps[8]);
    int seg_sidedef(const Seg &s) const;          // -1 if none (level.rs:101-109)
    int seg_back_sidedef(const Seg &s) const;     // level.rs:111-119
    int sector_min_light(int sector_id) const;    // level.rs:163-182
};

// Row-major 16-bit image; a pixel with a non-zero high byte is transparent (image.rs:11-17).
struct Image {
    int w = 0, h = 0, xoff = 0, yoff = 0;
    std::vector<uint16_t> px;
    static Image blank(int w, int h);                        // filled 0xff00 (image.rs:30)
    static Image decode(const uint8_t *buf, size_t size);    // throws WadError (image.rs:39-169)
    void blit(const Image &src, int ox, int oy, bool ignore_transparency);   // image.rs:171-252
};

struct TextureDirectory {
    std::vector<std::array<uint8_t, 768>> palettes;
    std::vector<std::array<uint8_t, 256>> colormaps;
    std::vector<std::pair<Name, int>> patches;     // PNAMES order; second = index into images or -1
    std::vector<Image> patch_images;
    std::unordered_map<Name, int, NameHash> texture_index;   // name -> textures[] (later wins)
    std::vector<Image> textures;
    std::unordered_map<Name, int, NameHash> flat_index;      // name -> flat lump index in archive (or into own_flats)
    std::vector<std::array<uint8_t, 4096>> own_flats;        // flats handed over by the caller (wad == nullptr)
    const Archive *wad = nullptr;

    static TextureDirectory load(const Archive &wad);
    const Image *texture(const Name &n) const;
    const uint8_t *flat(const Name &n) const;      // 4096 bytes or nullptr
};

}  // namespace b2d
