// See b2d_wad.hpp for the reference citations.
#include "b2d_wad.hpp"

#include <cstdio>
#include <cstring>

namespace b2d {

namespace {

inline uint16_t rd_u16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline int16_t rd_i16(const uint8_t *p) { return (int16_t)rd_u16(p); }
inline uint32_t rd_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
inline int32_t rd_i32(const uint8_t *p) { return (int32_t)rd_u32(p); }

[[noreturn]] void corrupt(const std::string &m) { throw WadError(kErrCorrupt, m); }

bool valid_name_byte(uint8_t b) {
    return (b >= 'A' && b <= 'Z') || (b >= '0' && b <= '9') || b == '_' || b == '-' || b == '[' ||
           b == ']' || b == '%' || b == '\\';
}

}  // namespace

Name make_name(const uint8_t *bytes, size_t size) {
    Name out{};
    bool nulled = false;
    size_t n = size < 8 ? size : 8;
    for (size_t i = 0; i < n; i++) {
        uint8_t b = bytes[i];
        if (b >= 0x80) corrupt("invalid byte in wad name");
        if (b >= 'a' && b <= 'z') b = (uint8_t)(b - 32);
        if (b == 0) { nulled = true; break; }
        if (!valid_name_byte(b)) corrupt("invalid byte in wad name");
        out[i] = b;
    }
    if (!nulled && size > 8) corrupt("wad name too long");
    return out;
}

Name make_name(const char *literal) {
    return make_name(reinterpret_cast<const uint8_t *>(literal), std::strlen(literal));
}

std::string name_str(const Name &n) {
    std::string s;
    for (uint8_t c : n) {
        if (!c) break;
        s.push_back((char)c);
    }
    return s;
}

bool is_sky_flat(const Name &n) {
    static const Name sky = make_name("F_SKY1");
    return n == sky;
}

// ------------------------------------------------------------------------------------ Archive
Archive::Archive(std::vector<uint8_t> data) { add_file(data, true); }

Archive::Archive(std::vector<std::vector<uint8_t>> files) {
    if (files.empty()) corrupt("no wad file");
    for (size_t i = 0; i < files.size(); i++) add_file(files[i], i == 0);
}

void Archive::add_file(const std::vector<uint8_t> &file, bool iwad) {
    if (file.size() < 12) corrupt("bad wad header");
    if (std::memcmp(file.data(), iwad ? "IWAD" : "PWAD", 4) != 0)
        corrupt(iwad ? "bad wad header identifier (IWAD required)" : "bad wad header identifier (PWAD required for an overlay)");
    int32_t num = rd_i32(&file[4]);
    int32_t table = rd_i32(&file[8]);
    if (num < 0 || table < 0 || (uint64_t)table + 16ull * (uint64_t)num > file.size())
        corrupt("lump info table out of bounds");
    const int64_t base = (int64_t)data_.size();                  // lump positions become offsets into the concatenation
    data_.insert(data_.end(), file.begin(), file.end());
    static const Name things = make_name("THINGS");
    const int first = (int)lumps_.size();
    lumps_.reserve(lumps_.size() + (size_t)num);
    for (int32_t i = 0; i < num; i++) {
        const uint8_t *e = &file[(size_t)table + 16u * (size_t)i];
        Lump l;
        l.pos = rd_i32(e);
        l.size = rd_i32(e + 4);
        // a lump must lie inside its own file (a position past it would alias the next file of the concatenation)
        if (l.size > 0 && (l.pos < 0 || (uint64_t)l.pos + (uint64_t)l.size > file.size())) l.pos = -1;
        else if (l.pos >= 0) l.pos += base;
        l.name = make_name(e + 8, 8);       // an invalid name fails the whole open (name.rs:132-139)
        index_[l.name] = (int)lumps_.size();
        lumps_.push_back(l);
        if (l.name == things) {
            if (i == 0) corrupt("THINGS lump without level marker");
            const int marker = first + i - 1;
            bool replaced = false;
            for (int &lv : levels_)
                if (lumps_[(size_t)lv].name == lumps_[(size_t)marker].name) { lv = marker; replaced = true; }
            if (!replaced) levels_.push_back(marker);
        }
    }
    files_.emplace_back(first, (int)lumps_.size());
}

std::vector<std::pair<int, int>> Archive::marker_ranges(const char *start, const char *start2, const char *end,
                                                       const char *end2) const {
    const Name s1 = make_name(start), s2 = make_name(start2), e1 = make_name(end), e2 = make_name(end2);
    std::vector<std::pair<int, int>> out;
    for (size_t f = 0; f < files_.size(); f++) {
        int a = -1, b = -1;
        for (int i = files_[f].first; i < files_[f].second; i++) {
            const Name &n = lumps_[(size_t)i].name;
            if (n == s1 || n == s2) a = i;                       // the later duplicate wins, as a by-name lookup would
            if (n == e1 || n == e2) b = i;
        }
        if (f == 0 && (a < 0 || b < 0)) corrupt(std::string("missing required lump ") + (a < 0 ? start : end));
        if (a >= 0 && b >= 0) out.emplace_back(a, b);
    }
    return out;
}

namespace {
std::vector<uint8_t> read_file(const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw WadError(kErrIo, "cannot open wad file '" + path + "'");
    std::vector<uint8_t> data;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + n);
    bool bad = std::ferror(f) != 0;
    std::fclose(f);
    if (bad) throw WadError(kErrIo, "error reading wad file '" + path + "'");
    return data;
}
}  // namespace

Archive Archive::open(const std::string &path) { return Archive(read_file(path)); }

Archive Archive::open(const std::vector<std::string> &paths) {
    std::vector<std::vector<uint8_t>> files;
    for (const std::string &p : paths) files.push_back(read_file(p));
    return Archive(std::move(files));
}

const Name &Archive::level_name(int level_index) const { return lumps_.at((size_t)level_lump_index(level_index)).name; }

int Archive::level_lump_index(int level_index) const {
    if (level_index < 0 || level_index >= (int)levels_.size()) corrupt("no such level index");
    return levels_[(size_t)level_index];
}

int Archive::find(const Name &name) const {
    auto it = index_.find(name);
    return it == index_.end() ? -1 : it->second;
}

int Archive::require(const char *name) const {
    int i = find(make_name(name));
    if (i < 0) corrupt(std::string("missing required lump ") + name);
    return i;
}

const Lump &Archive::lump(int index) const {
    if (index < 0 || index >= (int)lumps_.size()) corrupt("missing required lump index");
    return lumps_[(size_t)index];
}

const uint8_t *Archive::lump_data(int index) const {
    const Lump &l = lump(index);
    if (l.size == 0) return nullptr;
    if (l.pos < 0 || l.size < 0 || (uint64_t)l.pos + (uint64_t)l.size > data_.size())
        corrupt("lump '" + name_str(l.name) + "' out of file bounds");
    return &data_[(size_t)l.pos];
}

// ------------------------------------------------------------------------------------ Level
namespace {

template <typename T, size_t kSize, typename Fn>
std::vector<T> decode_vec(const RawLump &l, const char *what, Fn decode_one) {
    if (l.size == 0 || l.size % kSize != 0 || !l.data)
        corrupt(std::string("bad lump size for '") + what + "'");        // archive.rs:178-181
    const uint8_t *p = l.data;
    size_t n = l.size / kSize;
    std::vector<T> out(n);
    for (size_t i = 0; i < n; i++) out[i] = decode_one(p + i * kSize);
    return out;
}

}  // namespace

Level Level::load(const Archive &wad, int level_index) {
    int start = wad.level_lump_index(level_index);
    RawLump lumps[8];
    for (int k = 0; k < 8; k++) {               // THINGS .. SECTORS at fixed offsets from the marker (level.rs:13-20)
        const Lump &l = wad.lump(start + 1 + k);
        lumps[k].size = l.size > 0 ? (size_t)l.size : 0;
        lumps[k].data = wad.lump_data(start + 1 + k);
    }
    return from_lumps(wad.lump(start).name, lumps);
}

Level Level::from_lumps(const Name &name, const RawLump lumps[8]) {
    Level lv;
    lv.name = name;
    lv.things = decode_vec<Thing, 10>(lumps[0], "THINGS", [](const uint8_t *p) {
        return Thing{rd_i16(p), rd_i16(p + 2), rd_i16(p + 4), rd_u16(p + 6), rd_u16(p + 8)};
    });
    lv.linedefs = decode_vec<Linedef, 14>(lumps[1], "LINEDEFS", [](const uint8_t *p) {
        return Linedef{rd_u16(p), rd_u16(p + 2), rd_u16(p + 4), rd_u16(p + 6), rd_u16(p + 8),
                       rd_i16(p + 10), rd_i16(p + 12)};
    });
    lv.sidedefs = decode_vec<Sidedef, 30>(lumps[2], "SIDEDEFS", [](const uint8_t *p) {
        Sidedef s;
        s.xoff = rd_i16(p); s.yoff = rd_i16(p + 2);
        s.upper = make_name(p + 4, 8); s.lower = make_name(p + 12, 8); s.middle = make_name(p + 20, 8);
        s.sector = rd_u16(p + 28);
        return s;
    });
    lv.vertices = decode_vec<Vertex, 4>(lumps[3], "VERTEXES", [](const uint8_t *p) {
        return Vertex{rd_i16(p), rd_i16(p + 2)};
    });
    lv.segs = decode_vec<Seg, 12>(lumps[4], "SEGS", [](const uint8_t *p) {
        return Seg{rd_u16(p), rd_u16(p + 2), rd_u16(p + 4), rd_u16(p + 6), rd_u16(p + 8), rd_u16(p + 10)};
    });
    lv.subsectors = decode_vec<Subsector, 4>(lumps[5], "SSECTORS", [](const uint8_t *p) {
        return Subsector{rd_u16(p), rd_u16(p + 2)};
    });
    lv.nodes = decode_vec<Node, 28>(lumps[6], "NODES", [](const uint8_t *p) {
        Node n;
        n.x = rd_i16(p); n.y = rd_i16(p + 2); n.dx = rd_i16(p + 4); n.dy = rd_i16(p + 6);
        for (int k = 0; k < 4; k++) { n.rbox[k] = rd_i16(p + 8 + 2 * k); n.lbox[k] = rd_i16(p + 16 + 2 * k); }
        n.right = rd_u16(p + 24); n.left = rd_u16(p + 26);
        return n;
    });
    lv.sectors = decode_vec<Sector, 26>(lumps[7], "SECTORS", [](const uint8_t *p) {
        Sector s;
        s.floor = rd_i16(p); s.ceil = rd_i16(p + 2);
        s.floor_tex = make_name(p + 4, 8); s.ceil_tex = make_name(p + 12, 8);
        s.light = rd_i16(p + 20); s.type = rd_u16(p + 22); s.tag = rd_u16(p + 24);
        return s;
    });
    return lv;
}

int Level::seg_sidedef(const Seg &s) const {
    if (s.linedef >= linedefs.size()) return -1;
    const Linedef &l = linedefs[s.linedef];
    int idx = s.direction == 0 ? l.right : l.left;
    return (idx >= 0 && idx < (int)sidedefs.size()) ? idx : -1;
}

int Level::seg_back_sidedef(const Seg &s) const {
    if (s.linedef >= linedefs.size()) return -1;
    const Linedef &l = linedefs[s.linedef];
    int idx = s.direction == 1 ? l.right : l.left;
    return (idx >= 0 && idx < (int)sidedefs.size()) ? idx : -1;
}

int Level::sector_min_light(int sector_id) const {
    int m = sectors[(size_t)sector_id].light;
    for (const Linedef &l : linedefs) {
        if (l.right < 0 || l.left < 0 || l.right >= (int)sidedefs.size() || l.left >= (int)sidedefs.size())
            continue;
        int a = sidedefs[(size_t)l.right].sector, b = sidedefs[(size_t)l.left].sector;
        int other = -1;
        if (a == sector_id && b != sector_id) other = b;
        else if (b == sector_id && a != sector_id) other = a;
        if (other >= 0 && other < (int)sectors.size() && sectors[(size_t)other].light < m)
            m = sectors[(size_t)other].light;
    }
    return m;
}

// ------------------------------------------------------------------------------------ Image
Image Image::blank(int w, int h) {
    if (w < 0 || h < 0 || w > 4096 || h > 4096) corrupt("image too large");
    Image im;
    im.w = w; im.h = h;
    im.px.assign((size_t)w * (size_t)h, 0xff00);
    return im;
}

Image Image::decode(const uint8_t *buf, size_t size) {
    if (!buf || size < 8) corrupt("image missing header");
    Image im;
    im.w = rd_u16(buf); im.h = rd_u16(buf + 2);
    if (im.w > 4096 || im.h > 4096) corrupt("image too large");
    im.xoff = rd_i16(buf + 4); im.yoff = rd_i16(buf + 6);
    if (size < 8 + 4 * (size_t)im.w) corrupt("unfinished image column directory");
    im.px.assign((size_t)im.w * (size_t)im.h, 0xffff);
    for (int x = 0; x < im.w; x++) {
        size_t p = rd_u32(buf + 8 + 4 * (size_t)x);
        if (p >= size) corrupt("invalid image column offset");
        int last_row = -1;
        for (;;) {
            if (p >= size) corrupt("unfinished image column");
            unsigned row = buf[p++];
            if (row == 255) break;
            // tall patches (DeePsea convention, used by PWAD textures taller than 254 rows): a post whose topdelta does not
            // exceed the previous post's is relative to it.  Posts of a stock patch are strictly increasing, so nothing
            // changes for the pictures the reference (image.rs:73-157, absolute offsets only) can represent.
            if (last_row >= 0 && (int)row <= last_row) row += (unsigned)last_row;
            last_row = (int)row;
            if (p >= size) corrupt("missing image run length");
            unsigned len = buf[p++];
            if ((int)(row + len) > im.h) corrupt("image run too big");
            if (p >= size) corrupt("image missing padding byte 1");
            p++;
            if (size - p < len) corrupt("image source underrun");
            for (unsigned k = 0; k < len; k++) im.px[(size_t)(row + k) * (size_t)im.w + (size_t)x] = buf[p + k];
            p += len;
            if (p >= size) corrupt("image missing padding byte 2");
            p++;
        }
    }
    return im;
}

void Image::blit(const Image &src, int ox, int oy, bool ignore_transparency) {
    if (ox >= w || oy >= h) return;
    int y0 = oy < 0 ? -oy : 0, x0 = ox < 0 ? -ox : 0;
    int y1 = h > src.h + oy ? src.h : h - oy;
    int x1 = w > src.w + ox ? src.w : w - ox;
    if (x1 <= x0 || y1 <= y0) return;   // fully off the top/left edge
    for (int y = y0; y < y1; y++) {
        const uint16_t *s = &src.px[(size_t)y * (size_t)src.w];
        uint16_t *d = &px[(size_t)(y + oy) * (size_t)w + (size_t)ox];
        for (int x = x0; x < x1; x++) {
            uint16_t v = s[x];
            if (ignore_transparency || !(v & 0x8000)) d[x] = v;
        }
    }
}

// ------------------------------------------------------------------------------------ textures
const Image *TextureDirectory::texture(const Name &n) const {
    auto it = texture_index.find(n);
    return it == texture_index.end() ? nullptr : &textures[(size_t)it->second];
}

const uint8_t *TextureDirectory::flat(const Name &n) const {
    auto it = flat_index.find(n);
    if (it == flat_index.end()) return nullptr;
    if (!wad) return (size_t)it->second < own_flats.size() ? own_flats[(size_t)it->second].data() : nullptr;
    if (wad->lump(it->second).size < 4096) return nullptr;
    return wad->lump_data(it->second);
}

TextureDirectory TextureDirectory::load(const Archive &wad) {
    TextureDirectory td;
    td.wad = &wad;
    {   // PLAYPAL / COLORMAP blobs (tex.rs:57-58, archive.rs:206-228)
        int ip = wad.require("PLAYPAL");
        const Lump &lp = wad.lump(ip);
        if (lp.size <= 0 || lp.size % 768) corrupt("bad PLAYPAL size");
        const uint8_t *p = wad.lump_data(ip);
        td.palettes.resize((size_t)lp.size / 768);
        for (size_t i = 0; i < td.palettes.size(); i++) std::memcpy(td.palettes[i].data(), p + 768 * i, 768);
        int ic = wad.require("COLORMAP");
        const Lump &lc = wad.lump(ic);
        if (lc.size <= 0 || lc.size % 256) corrupt("bad COLORMAP size");
        const uint8_t *c = wad.lump_data(ic);
        td.colormaps.resize((size_t)lc.size / 256);
        for (size_t i = 0; i < td.colormaps.size(); i++) std::memcpy(td.colormaps[i].data(), c + 256 * i, 256);
    }
    {   // PNAMES (tex.rs:358-410): unreadable / missing patches are kept as holes
        int ipn = wad.require("PNAMES");
        const Lump &l = wad.lump(ipn);
        const uint8_t *p = wad.lump_data(ipn);
        if (!p || l.size < 4) corrupt("missing number of patches in PNAMES");
        uint32_t n = rd_u32(p);
        for (uint32_t i = 0; i < n; i++) {
            if (4 + 8 * (uint64_t)(i + 1) > (uint64_t)l.size) break;
            Name nm;
            try { nm = make_name(p + 4 + 8 * (size_t)i, 8); } catch (const WadError &) { continue; }
            int li = wad.find(nm);
            int img = -1;
            if (li >= 0) {
                const uint8_t *bytes = wad.lump_data(li);      // a read failure fails the load (tex.rs:384-385)
                try {
                    Image im = Image::decode(bytes, (size_t)wad.lump(li).size);
                    img = (int)td.patch_images.size();
                    td.patch_images.push_back(std::move(im));
                } catch (const WadError &) { img = -1; }
            }
            td.patches.emplace_back(nm, img);
        }
    }
    for (const char *lump_name : {"TEXTURE1", "TEXTURE2"}) {   // tex.rs:499-592
        int it = wad.find(make_name(lump_name));
        if (it < 0) continue;
        const Lump &l = wad.lump(it);
        const uint8_t *p = wad.lump_data(it);
        size_t size = (size_t)l.size;
        if (!p || size < 4) corrupt("missing number of textures");
        uint32_t n = rd_u32(p);
        if (4ull * n >= size - 4) corrupt("textures lump too small for offsets");
        size_t decoded_px = 0;                 // a 22-byte entry can declare 4096x4096: cap what a tiny lump can make us allocate
        for (uint32_t i = 0; i < n; i++) {
            size_t off = rd_u32(p + 4 + 4 * (size_t)i);
            if (off >= size) corrupt("textures lump too small for offsets");
            if (off + 22 > size) continue;
            Name nm;
            try { nm = make_name(p + off, 8); } catch (const WadError &) { continue; }
            int w = rd_u16(p + off + 12), h = rd_u16(p + off + 14);
            int np = rd_u16(p + off + 20);
            if (w > 4096 || h > 4096) continue;
            decoded_px += (size_t)w * (size_t)h;
            if (decoded_px > ((size_t)1 << 28)) corrupt("composed textures exceed 256 Mpixel");
            Image img = Image::blank(w, h);
            size_t q = off + 22;
            for (int k = 0; k < np; k++) {
                if (q + 10 > size) break;
                int ox = rd_i16(p + q), oy = rd_i16(p + q + 2);
                unsigned pi = rd_u16(p + q + 4);
                q += 10;
                if (oy <= 0) oy = 0;                                 // tex.rs:560-567
                if (pi < td.patches.size() && td.patches[pi].second >= 0)
                    img.blit(td.patch_images[(size_t)td.patches[pi].second], ox, oy, k == 0);
            }
            auto found = td.texture_index.find(nm);
            if (found != td.texture_index.end()) td.textures[(size_t)found->second] = std::move(img);
            else { td.texture_index[nm] = (int)td.textures.size(); td.textures.push_back(std::move(img)); }
        }
    }
    {   // flats (tex.rs:594-606)
        for (const auto &rg : wad.marker_ranges("F_START", "FF_START", "F_END", "FF_END"))
            for (int i = rg.first; i < rg.second; i++)
                if (wad.lump(i).size != 0) {
                    (void)wad.lump_data(i);                    // read_bytes()? : out-of-file lumps fail the load
                    td.flat_index[wad.lump(i).name] = i;       // a PWAD's flat of the same name wins
                }
    }
    {   // sprites share the texture name space and may shadow a texture (tex.rs:475-497)
        for (const auto &rg : wad.marker_ranges("S_START", "SS_START", "S_END", "SS_END"))
        for (int i = rg.first + 1; i < rg.second; i++) {
            const uint8_t *bytes = wad.lump_data(i);           // read failure propagates (tex.rs:484-485)
            try {
                Image im = Image::decode(bytes, (size_t)wad.lump(i).size);
                const Name &nm = wad.lump(i).name;
                auto found = td.texture_index.find(nm);
                if (found != td.texture_index.end()) td.textures[(size_t)found->second] = std::move(im);
                else { td.texture_index[nm] = (int)td.textures.size(); td.textures.push_back(std::move(im)); }
            } catch (const WadError &) { continue; }
        }
    }
    return td;
}

}  // namespace b2d
