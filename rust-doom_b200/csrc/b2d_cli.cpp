// `b2d` -- command line front end on the C ABI (include/b2d.h), mirroring rs_doom's flags
// (reference src/main.rs:17-80): -i/--iwad, -l/--level, -r/--resolution WxH, -f/--fov, and the sub-commands
// `list-levels` (main.rs:116-121) and `check` (main.rs:99-115).  Rendering options are ours: --poses N turns the
// camera N steps around the spawn point, --tics T sets the level time, --dump FILE writes the first frame and
// --stream FILE all frames as binary PPM.  This is the compiled-code host side of the boundary: it links
// libb2d.so and uses nothing but the header; every frame comes from the CUDA kernels (no CPU path).
// Multi-GPU (one process per GPU): --rank R --world N --id-file PATH --chunk C renders the pose list through
// b2d_render_sharded -- rank 0 writes the NCCL unique id to PATH, the others read it -- with the frame all-gather on, and
// prints one checksum line per rank over all gathered frames (every rank must print the same value).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <unistd.h>

#include "../../include/b2d.h"

namespace {

int fail(const char *what) {
    std::fprintf(stderr, "Fatal error: %s: %s\n", what, b2d_last_error());
    return 1;
}

void write_ppm(std::FILE *f, const uint32_t *rgba, int w, int h) {
    std::fprintf(f, "P6\n%d %d\n255\n", w, h);
    std::vector<uint8_t> row((size_t)w * 3);
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            const uint32_t p = rgba[(size_t)y * w + x];
            row[3 * x] = (uint8_t)p; row[3 * x + 1] = (uint8_t)(p >> 8); row[3 * x + 2] = (uint8_t)(p >> 16);
        }
        std::fwrite(row.data(), 1, row.size(), f);
    }
}

// b2d_render_sharded consumer: per-frame checksums of every gathered chunk into a device table (the table lives in device memory
// obtained through b2d_device_alloc: the CLI itself links no CUDA library)
struct ShardSink {
    uint32_t *d_sums;      // world x per, device
    size_t per, npix;
};
void on_chunk(void *user, int, size_t first, size_t cnt, const uint8_t *d_frames, int ranks, void *stream) {
    ShardSink *s = static_cast<ShardSink *>(user);
    for (int q = 0; q < ranks; q++)
        b2d_frame_checksums_device(d_frames + (size_t)q * cnt * s->npix, cnt, s->npix, s->d_sums + (size_t)q * s->per + first, stream);
}

}  // namespace

int main(int argc, char **argv) {
    std::string iwad, dump, stream, command, id_file;
    int level = 0, width = 1280, height = 720, nposes = 1, rank = 0, world = 0, chunk = 16;
    double fov = 65.0;
    unsigned long tics = 0;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", name); std::exit(2); }
            return argv[++i];
        };
        if (a == "-i" || a == "--iwad") iwad = next("--iwad");
        else if (a == "-m" || a == "--metadata") next("--metadata");           // accepted; the sky table is built in
        else if (a == "-l" || a == "--level") level = std::atoi(next("--level"));
        else if (a == "-f" || a == "--fov") fov = std::atof(next("--fov"));
        else if (a == "-r" || a == "--resolution") {
            if (std::sscanf(next("--resolution"), "%dx%d", &width, &height) != 2) {
                std::fprintf(stderr, "resolution format is WIDTHxHEIGHT\n");
                return 2;
            }
        } else if (a == "--poses") nposes = std::atoi(next("--poses"));
        else if (a == "--tics") tics = std::strtoul(next("--tics"), nullptr, 10);
        else if (a == "--dump") dump = next("--dump");
        else if (a == "--stream") stream = next("--stream");
        else if (a == "--rank") rank = std::atoi(next("--rank"));
        else if (a == "--world") world = std::atoi(next("--world"));
        else if (a == "--chunk") chunk = std::atoi(next("--chunk"));
        else if (a == "--id-file") id_file = next("--id-file");
        else if (a == "list-levels" || a == "check") command = a;
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (iwad.empty()) { std::fprintf(stderr, "--iwad FILE is required\n"); return 2; }
    if (nposes < 1) nposes = 1;

    b2d_archive *arch = nullptr;
    if (b2d_archive_open(iwad.c_str(), &arch) != B2D_OK) return fail("open");
    const int nlevels = b2d_archive_num_levels(arch);
    if (command == "list-levels") {
        for (int i = 0; i < nlevels; i++) {
            char name[9] = {0};
            b2d_archive_level_name(arch, i, name);
            std::printf("%3d %8s\n", i, name);
        }
        b2d_archive_close(arch);
        return 0;
    }
    if (command == "check") {
        for (int i = 0; i < nlevels; i++) {
            char name[9] = {0};
            b2d_archive_level_name(arch, i, name);
            b2d_scene *sc = nullptr;
            if (b2d_scene_create(arch, i, &sc) != B2D_OK) { b2d_archive_close(arch); return fail(name); }
            b2d_scene_info info;
            b2d_scene_info_get(sc, &info);
            std::printf("Level %d (%s): %d segs, %d subsectors, %d sectors, %d textures: ok\n", i, name, info.n_segs,
                        info.n_ssectors, info.n_sectors, info.n_textures);
            b2d_scene_destroy(sc);
        }
        b2d_archive_close(arch);
        return 0;
    }

    b2d_scene *sc = nullptr;
    if (b2d_scene_create(arch, level, &sc) != B2D_OK) { b2d_archive_close(arch); return fail("level"); }
    b2d_scene_info info;
    b2d_scene_info_get(sc, &info);
    if (!info.has_start) { std::fprintf(stderr, "Fatal error: the level has no player-1 start\n"); return 1; }
    b2d_view view;
    if (b2d_view_init(&view, width, height, fov) != B2D_OK) return fail("view");
    b2d_renderer *r = nullptr;
    if (b2d_renderer_create(sc, &view, world > 0 ? rank : 0, nposes < 64 ? nposes : 64, &r) != B2D_OK) return fail("renderer");
    if (b2d_renderer_set_time(r, (uint32_t)tics) != B2D_OK) return fail("time");

    std::vector<b2d_pose> poses((size_t)nposes, info.start);
    for (int i = 0; i < nposes; i++)          // look around from the spawn point
        poses[(size_t)i].angle = info.start.angle + (uint32_t)(((uint64_t)i << 32) / (uint64_t)nposes);
    const size_t npix = (size_t)width * height;
    if (world > 0) {
        // ---- sharded: every rank runs this with the same pose list
        uint8_t id[B2D_COMM_ID_BYTES];
        if (id_file.empty()) { std::fprintf(stderr, "--id-file PATH is required with --world\n"); return 2; }
        if (rank == 0) {
            if (b2d_comm_unique_id(id) != B2D_OK) return fail("unique id");
            const std::string tmp = id_file + ".tmp";
            std::FILE *f = std::fopen(tmp.c_str(), "wb");
            if (!f || std::fwrite(id, 1, sizeof id, f) != sizeof id) { std::perror(tmp.c_str()); return 1; }
            std::fclose(f);
            std::rename(tmp.c_str(), id_file.c_str());
        } else {
            std::FILE *f = nullptr;
            for (int tries = 0; tries < 600 && !(f = std::fopen(id_file.c_str(), "rb")); tries++) usleep(100000);
            if (!f || std::fread(id, 1, sizeof id, f) != sizeof id) { std::fprintf(stderr, "cannot read %s\n", id_file.c_str()); return 1; }
            std::fclose(f);
        }
        b2d_comm *comm = nullptr;
        if (b2d_comm_create(id, rank, world, rank, &comm) != B2D_OK) return fail("communicator");
        const size_t per = ((size_t)nposes + (size_t)world - 1) / (size_t)world;
        ShardSink sink{nullptr, per, npix};
        if (b2d_device_alloc(rank, sizeof(uint32_t) * per * (size_t)world, reinterpret_cast<void **>(&sink.d_sums)) != B2D_OK) return fail("device memory");
        b2d_sharded_stats st;
        if (b2d_render_sharded(r, comm, poses.data(), (size_t)nposes, (size_t)chunk, B2D_SHARD_RENDER_GATHER, on_chunk, &sink, &st) != B2D_OK)
            return fail("sharded render");
        int32_t bits = 0;
        if (b2d_renderer_status(r, &bits) != B2D_OK) return fail("status");
        if (bits) { std::fprintf(stderr, "Fatal error: frames incomplete (status %d)\n", bits); return 1; }
        std::vector<uint32_t> sums(per * (size_t)world);
        if (b2d_device_download(rank, sums.data(), sink.d_sums, sums.size() * sizeof(uint32_t)) != B2D_OK) return fail("download");
        uint32_t all = 0;
        for (size_t i = 0; i < sums.size(); i++) all = all * 31u + sums[i];
        std::printf("rank %d/%d: %lld frames gathered in %lld chunk(s), %.3f ms, checksum %08x, buffers %s\n", rank, world,
                    (long long)st.frames_gathered, (long long)st.chunks, st.total_ms, all, st.registration);
        b2d_device_free(rank, sink.d_sums);
        b2d_comm_destroy(comm);
        b2d_renderer_destroy(r);
        b2d_scene_destroy(sc);
        b2d_archive_close(arch);
        return 0;
    }
    std::vector<uint8_t> index(npix * (size_t)nposes);
    std::vector<uint32_t> rgba(npix * (size_t)nposes);
    if (b2d_render(r, poses.data(), (size_t)nposes, index.data(), rgba.data()) != B2D_OK) return fail("render");
    std::printf("rendered %d frame(s) %dx%d\n", nposes, width, height);
    if (!dump.empty()) {
        std::FILE *f = std::fopen(dump.c_str(), "wb");
        if (!f) { std::perror(dump.c_str()); return 1; }
        write_ppm(f, rgba.data(), width, height);
        std::fclose(f);
    }
    if (!stream.empty()) {
        std::FILE *f = std::fopen(stream.c_str(), "wb");
        if (!f) { std::perror(stream.c_str()); return 1; }
        for (int i = 0; i < nposes; i++) write_ppm(f, rgba.data() + npix * (size_t)i, width, height);
        std::fclose(f);
    }
    b2d_renderer_destroy(r);
    b2d_scene_destroy(sc);
    b2d_archive_close(arch);
    return 0;
}
