// sm_100a kernels of the Doom-WAD software renderer.
//
//   b2d_walk_kernel    : one CTA per frame.  Thread-parallel view transform of all vertices, per-seg
//                        projection setup and per-node-child bounding-box ranges into shared memory,
//                        then (warp 0) a front-to-back BSP walk with a warp-wide solid-column bitmask
//                        (ballot / lane-striped words) that emits the compact seg worklist.
//   b2d_raster_kernel  : one warp per (frame, 32-column strip); lane = screen column.  Consumes the
//                        worklist front to back, keeps the per-column clip window in registers,
//                        draws wall columns, floor/ceiling spans and sky from texel planes that already
//                        carry the light->colormap lookup; every pixel is written exactly once.
//   b2d_prelight_*     : build those planes once per renderer (32 light rows x texels / flats).
//   b2d_palette_kernel : index -> RGBA8 with the 256-entry palette in shared memory, 128-bit I/O.
//
// There is no dense contraction anywhere on this path, so no tensor-core (tcgen05) work: the
// kernels are integer/LSU/latency bound and are tuned against the HBM write roofline.
#include "b2d_kernels.cuh"

#include <cstdlib>

namespace b2d {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;

// texel / table loads of the raster: read-only path; B2D_LOAD_EL (A/B, profiles/README.md) asks L1 to keep them
#if defined(B2D_LOAD_L2EL)
// A/B: ask L2 to keep the pre-lit planes (a 2 GB/ms write stream passes through the same L2)
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
#endif
__device__ __forceinline__ uint32_t tex_ld(const uint8_t *p) {
#if defined(B2D_LOAD_L2EL)
    uint32_t v;
    asm("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(l2_keep_policy()));
    return v;
#elif defined(B2D_LOAD_EL)
    uint32_t v;
    asm("ld.global.nc.L1::evict_last.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ uint32_t tex_ld(const uint32_t *p) {
#if defined(B2D_LOAD_L2EL)
    uint32_t v;
    asm("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(l2_keep_policy()));
    return v;
#elif defined(B2D_LOAD_EL)
    uint32_t v;
    asm("ld.global.nc.L1::evict_last.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
constexpr int kMaskWords = 128;      // up to 4096 columns
constexpr int kStackDepth = 128;

// packed per-seg / per-box visibility word in shared memory
constexpr uint32_t kVisBit = 1u << 25, kSolidBit = 1u << 24;
__device__ __forceinline__ uint32_t pack_range(int lo, int hi, uint32_t bits) {
    return (uint32_t)lo | ((uint32_t)hi << 12) | bits;
}
__device__ __forceinline__ int range_lo(uint32_t r) { return (int)(r & 0xFFFu); }
__device__ __forceinline__ int range_hi(uint32_t r) { return (int)((r >> 12) & 0xFFFu); }

// bits of columns [lo, hi] that fall into mask word w
__device__ __forceinline__ uint32_t word_bits(int w, int lo, int hi) {
    int a = max(lo, w * 32), b = min(hi, w * 32 + 31);
    if (a > b) return 0u;
    uint32_t m = 0xFFFFFFFFu >> (31 - (b - w * 32));
    return m & (0xFFFFFFFFu << (a - w * 32));
}

// warp-wide: is any column of [lo, hi] still open?  (lane-striped words + ballot)
__device__ __forceinline__ bool range_open(const uint32_t *mask, int lane, int lo, int hi, int passes) {
    bool open = false;
#pragma unroll
    for (int k = 0; k < kMaskWords / 32; k++) {
        if (k >= passes) break;              // only ceil(W / 1024) lane passes carry columns
        int w = lane + 32 * k;
        uint32_t bits = word_bits(w, lo, hi);
        if (bits & ~mask[w]) open = true;
    }
    return __any_sync(kFull, open);
}

// single lane: same test over its own range (used for per-seg culling)
__device__ __forceinline__ bool lane_range_open(const uint32_t *mask, int lo, int hi) {
    for (int w = lo >> 5; w <= (hi >> 5); w++)
        if (word_bits(w, lo, hi) & ~mask[w]) return true;
    return false;
}

__device__ __forceinline__ void mark_solid(uint32_t *mask, int lane, int lo, int hi, int passes) {
#pragma unroll
    for (int k = 0; k < kMaskWords / 32; k++) {
        if (k >= passes) break;
        int w = lane + 32 * k;
        uint32_t bits = word_bits(w, lo, hi);
        if (bits) mask[w] |= bits;
    }
}

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

struct WalkSmem {
    size_t off_tx, off_tz, off_segr, off_boxr, off_node, off_ssec, off_sprr, off_sprz, off_mask, off_stack, off_list, total;
};
__host__ __device__ inline WalkSmem walk_layout(int nverts, int nsegs, int nnodes, int nss, int nsprites) {
    WalkSmem L;
    size_t o = 0;
    L.off_tx = o; o = align16(o + 4 * (size_t)nverts);
    L.off_tz = o; o = align16(o + 4 * (size_t)nverts);
    L.off_segr = o; o = align16(o + 4 * (size_t)nsegs);
    L.off_boxr = o; o = align16(o + 8 * (size_t)nnodes);
    L.off_node = o; o = align16(o + 32 * (size_t)nnodes);     // {x,y,dx,dy,rchild,lchild,-,-} per node
    L.off_ssec = o; o = align16(o + 16 * (size_t)nss);        // SSectorRec copies
    L.off_sprr = o; o = align16(o + 4 * (size_t)nsprites);    // packed column range per sprite
    L.off_sprz = o; o = align16(o + 4 * (size_t)nsprites);    // view depth per sprite (ordering inside a subsector)
    L.off_mask = o; o = align16(o + 4 * kMaskWords);
    L.off_stack = o; o = align16(o + 4 * kStackDepth);
    L.off_list = o; o = align16(o + 2 * ((size_t)nsegs + (size_t)nsprites));
    L.total = o;
    return L;
}

// 1-D bulk copy global -> shared memory through the TMA unit (cp.async.bulk, SASS UBLKCP), completion on an mbarrier.
// B2D_WALK_NO_BULK (build variant, A/B in profiles/README.md) stages the same bytes with a thread loop instead.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {      // bounded: a lost copy must not hang the GPU
    for (int spin = 0; spin < (1 << 24); spin++) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// Kernel 1: BSP walk -> worklist
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 7)     // 7 CTAs/SM (72 registers): 1036 frames resident on 148 SMs
b2d_walk_kernel(const __grid_constant__ DeviceScene sc, const __grid_constant__ View vw, const Pose *__restrict__ poses, int n,
                FrameConst *__restrict__ frames, SegFrame *__restrict__ work, int stride) {
    // One CTA per frame.  The per-frame setup (steps 1-3) and the worklist records (step 5) are data-parallel and
    // use all 128 threads; the traversal itself (step 4) is sequential and runs in warp 0 with the lanes working
    // on the segs of a subsector / the words of the column mask.  The kernel is latency-bound (one frame = one
    // dependent chain), so spreading the parallel phases over four warps shortens the chain directly.
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const WalkSmem L = walk_layout(sc.nverts, sc.nsegs, sc.nnodes, sc.nss, sc.nsprites);
    uint8_t *base = smem;
    __shared__ int s_count, s_status;
    __shared__ __align__(8) uint64_t s_bar;
    int32_t *tx = reinterpret_cast<int32_t *>(base + L.off_tx);
    int32_t *tz = reinterpret_cast<int32_t *>(base + L.off_tz);
    uint32_t *segr = reinterpret_cast<uint32_t *>(base + L.off_segr);
    uint32_t *boxr = reinterpret_cast<uint32_t *>(base + L.off_boxr);
    int4 *node_s = reinterpret_cast<int4 *>(base + L.off_node);       // traversal reads shared memory, not L2
    int4 *ssec_s = reinterpret_cast<int4 *>(base + L.off_ssec);
    uint32_t *sprr = reinterpret_cast<uint32_t *>(base + L.off_sprr);
    int32_t *sprz = reinterpret_cast<int32_t *>(base + L.off_sprz);
    uint32_t *mask = reinterpret_cast<uint32_t *>(base + L.off_mask);
    uint32_t *stack = reinterpret_cast<uint32_t *>(base + L.off_stack);
    uint16_t *list = reinterpret_cast<uint16_t *>(base + L.off_list);

    // One frame per CTA when the grid has a CTA per frame; a smaller (persistent) grid loops: that is the form for running
    // under another batch's raster -- a CTA per SM keeps the walk's footprint at 1/8 of the register file while it takes
    // a few frame latencies, all hidden behind the raster (launch_walk, B2D_TUNE bit 2).
    // The traversal tables (node partition lines + children, subsector records) do not depend on the frame: one bulk copy
    // per CTA brings them into shared memory while the threads are busy with steps 1-3 of the first frame.
    const uint32_t static_bytes = 32u * (uint32_t)sc.nnodes + 16u * (uint32_t)sc.nss;
    bool static_pending = false;
#if !defined(B2D_WALK_NO_BULK)
    if (static_bytes) {
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (tid == 0) bulk_copy_g2s(node_s, sc.walk_static, static_bytes, &s_bar);
        static_pending = true;
    }
#else
    for (uint32_t i = tid; i < static_bytes / 16u; i += nthr) node_s[i] = reinterpret_cast<const int4 *>(sc.walk_static)[i];
#endif
    for (int frame = blockIdx.x; frame < n; frame += gridDim.x) {
    FrameConst fc;
    frame_setup(poses[frame], fc);

    // 1. all vertices into view space (lane-parallel)
    for (int i = tid; i < sc.nverts; i += nthr) {
        int32_t vx = sc.verts[2 * i], vy = sc.verts[2 * i + 1];
        int32_t a, b;
        to_view(fc, vx, vy, a, b);
        tx[i] = a; tz[i] = b;
    }
    __syncthreads();

    // 2. per-seg exact column interval + static/solid flags (lane-parallel, 64-bit setup)
    for (int i = tid; i < sc.nsegs; i += nthr) {
        const SegRec &S = sc.segs[i];
        uint32_t packed = 0;
        int32_t flags = S.flags;
        if (!(flags & kSegInvalid)) {
            SegFrame sf;
            const bool closes = !(flags & kSegTwoSided) || S.otop <= S.obot;   // every column it covers closes
            if (seg_frame_setup(vw, tx[S.v1], tz[S.v1], tx[S.v2], tz[S.v2], sf, closes)) {
                bool solid = closes && (sf.flags & kSegFrameNoSkip);
                packed = pack_range(sf.xlo, sf.xhi, kVisBit | (solid ? kSolidBit : 0u));
            }
        }
        segr[i] = packed;
    }
    // 3. conservative column range of both child boxes of every node
    for (int i = tid; i < 2 * sc.nnodes; i += nthr) {
        const NodeRec &N = sc.nodes[i >> 1];
        const int32_t *box = (i & 1) ? N.lbox : N.rbox;
        int32_t b4[4] = {box[0], box[1], box[2], box[3]};
        int lo, hi;
        boxr[i] = box_range(fc, vw, b4, lo, hi) ? pack_range(lo, hi, kVisBit) : 0u;
    }
    if (static_pending) {                        // first frame of this CTA: the bulk copy has had steps 1-3 to land
        if (!mbar_wait(&s_bar, 0)) __trap();
        static_pending = false;
    }
    for (int i = tid; i < sc.nsprites; i += nthr) {          // decoration sprites: exact column interval
        const SpriteRec &P = sc.sprites[i];
        SpriteFrame sp;
        uint32_t packed = 0;
        sp.cz = 0;
        if (P.tex >= 0 && P.tex < sc.ntex && sprite_setup(fc, vw, P.x, P.y, (int32_t)sc.tex[P.tex].w, sp))
            packed = pack_range(sp.lo, sp.hi, kVisBit);
        sprr[i] = packed;
        sprz[i] = (int32_t)sp.cz;
    }
    // solid-column mask: columns >= W start out solid
    for (int w = tid; w < kMaskWords; w += nthr) mask[w] = ~word_bits(w, 0, vw.W - 1);
    if (tid == 0) stack[0] = sc.root;
    __syncthreads();

    // 4. front-to-back traversal in warp 0 (control flow is warp-uniform)
    const int passes = (vw.W + 1023) / 1024;
    int sp = warp == 0 ? 1 : 0, count = 0, status = 0;
    int budget = 2 * (sc.nnodes + sc.nss) + 64;      // a corrupt BSP with a cycle must not hang the GPU
    while (sp > 0) {
        if (--budget < 0) { status |= 4; break; }
        uint32_t child = stack[--sp];
        __syncwarp();
        if (child & kLeaf) {
            uint32_t id = child & 0x7FFFFFFFu;
            if (id >= (uint32_t)sc.nss) continue;
            const int4 ssv = ssec_s[id];
            SSectorRec ss;
            ss.first_seg = ssv.x; ss.num_segs = ssv.y; ss.sector = ssv.z; ss.sprites = ssv.w;
            if (ss.sector < 0) continue;
            {   // the subsector's decoration sprites come first: they stand in front of its far segs.  Among themselves
                // nearest first (drawn back to front, the nearer billboard ends up on top); ties keep the stored order.
                // A lane's slot = the number of visible sprites of the subsector that sort before its own.
                const int sfirst = ss.sprites & 0xFFFFFF;
                int scnt = (ss.sprites >> 24) & 0xFF;
                if (sfirst + scnt > sc.nsprites) scnt = sc.nsprites > sfirst ? sc.nsprites - sfirst : 0;
                int nvis = 0;
                for (int k0 = 0; k0 < scnt; k0 += 32) {
                    const int k = k0 + lane, pi = sfirst + k;
                    const uint32_t r = k < scnt ? sprr[pi] : 0u;
                    const bool vis = (r & kVisBit) && lane_range_open(mask, range_lo(r), range_hi(r));
                    int rank = 0, total = 0;
                    const int32_t myz = vis ? sprz[pi] : 0;
                    for (int j0 = 0; j0 < scnt; j0 += 32) {                 // warp-uniform loop over all sprites of the subsector
                        const int j = j0 + lane, pj = sfirst + j;
                        const uint32_t rj = j < scnt ? sprr[pj] : 0u;
                        const bool vj = (rj & kVisBit) && lane_range_open(mask, range_lo(rj), range_hi(rj));
                        const int32_t zj = vj ? sprz[pj] : 0;
                        unsigned mj = __ballot_sync(kFull, vj);
                        total += __popc(mj);
                        while (mj) {
                            const int b = __ffs(mj) - 1;
                            mj &= mj - 1;
                            const int32_t zb = __shfl_sync(kFull, zj, b);
                            const int jb = j0 + b;
                            if (vis && (zb < myz || (zb == myz && jb < k))) rank++;
                        }
                    }
                    const int pos = count + rank;
                    if (vis && pos < sc.nsegs + sc.nsprites) list[pos] = (uint16_t)(sc.nsegs + pi);
                    nvis = total;
                }
                count = min(count + nvis, sc.nsegs + sc.nsprites);
                __syncwarp();
            }
            for (int k0 = 0; k0 < ss.num_segs; k0 += 32) {
                int k = k0 + lane;
                int si = ss.first_seg + k;
                uint32_t r = k < ss.num_segs ? segr[si] : 0u;
                bool vis = (r & kVisBit) && lane_range_open(mask, range_lo(r), range_hi(r));
                unsigned m = __ballot_sync(kFull, vis);
                int pos = count + __popc(m & ((1u << lane) - 1u));
                if (vis && pos < sc.nsegs + sc.nsprites) list[pos] = (uint16_t)si;
                count = min(count + __popc(m), sc.nsegs + sc.nsprites);
                unsigned sm = __ballot_sync(kFull, vis && (r & kSolidBit));
                __syncwarp();
                while (sm) {
                    int j = __ffs(sm) - 1;
                    sm &= sm - 1;
                    uint32_t rj = __shfl_sync(kFull, r, j);
                    mark_solid(mask, lane, range_lo(rj), range_hi(rj), passes);
                }
                __syncwarp();
            }
            if (!range_open(mask, lane, 0, vw.W - 1, passes)) break;     // every column is closed
        } else {
            if (child >= (uint32_t)sc.nnodes) continue;
            const int4 nl = node_s[2 * child], nc = node_s[2 * child + 1];
            int side = node_side(fc.pose, nl.x, nl.y, nl.z, nl.w);   // 1: left child is near
            uint32_t near_c = (uint32_t)(side ? nc.y : nc.x), far_c = (uint32_t)(side ? nc.x : nc.y);
            uint32_t rn = boxr[2 * child + side], rf = boxr[2 * child + (side ^ 1)];
            bool far_vis = (rf & kVisBit) && range_open(mask, lane, range_lo(rf), range_hi(rf), passes);
            bool near_vis = (rn & kVisBit) && range_open(mask, lane, range_lo(rn), range_hi(rn), passes);
            int need = (far_vis ? 1 : 0) + (near_vis ? 1 : 0);
            if (sp + need > kStackDepth) { status = 1; break; }
            if (lane == 0) {
                int p = sp;
                if (far_vis) stack[p++] = far_c;
                if (near_vis) stack[p++] = near_c;
            }
            sp += need;
            __syncwarp();
        }
    }
    __syncwarp();
    if (warp == 0 && lane == 0) {
        if (count > stride) { count = stride; status |= 2; }
        s_count = count; s_status = status;
    }
    __syncthreads();
    count = s_count; status = s_status;

    // 5. worklist records (thread-parallel): the projection coefficients of each emitted seg
    for (int k = tid; k < count; k += nthr) {
        int si = list[k];
        SegFrame sf;
        if (si >= sc.nsegs) {                         // sprite entry: seg = -1 - sprite index, (cx, cz) in Nc/Nx
            const int pi = si - sc.nsegs;
            const SpriteRec &P = sc.sprites[pi];
            SpriteFrame sp;
            sprite_setup(fc, vw, P.x, P.y, (int32_t)sc.tex[P.tex].w, sp);
            sf.Nc = sp.cx; sf.Nx = sp.cz; sf.Dc = 0; sf.Dx = 0; sf.Dmax = 0; sf.Rm = 0; sf.sh = 0; sf.e = 0;
            sf.seg = -1 - pi; sf.xlo = (int16_t)sp.lo; sf.xhi = (int16_t)sp.hi; sf.flags = 0; sf.pad = 0;
            work[(size_t)frame * stride + k] = sf;
            continue;
        }
        const SegRec &S = sc.segs[si];
        seg_frame_setup(vw, tx[S.v1], tz[S.v1], tx[S.v2], tz[S.v2], sf, false);
        sf.seg = si;
        work[(size_t)frame * stride + k] = sf;
    }
    if (tid == 0) {
        if (status) atomicOr(sc.status_flag, status);
        fc.count = count;
        fc.status = status;
#pragma unroll
        for (int i = 0; i < 6; i++) fc.pad[i] = 0;
        frames[frame] = fc;
    }
    __syncthreads();                              // the next frame of this CTA reuses the shared-memory tables
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel 2: raster
// ------------------------------------------------------------------------------------------------
// shared-memory accesses by 32-bit shared-window address: keeps the per-pixel address arithmetic to one
// add (the generic-pointer form re-derives the window base for every access)
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

struct RasterCtx {
    const DeviceScene *sc;
    uint32_t pal_s;            // ... of the palette (rgba only)
    uint2 *rowz;               // this warp's 32-entry staging of per-row plane constants {depth Q8, plane offset / 64}
    PlaneDir dir;              // direction of this lane's column ray (Q18), for the flat texel
    uint8_t *fb;               // &index_fb[frame][0][x]
    uint32_t *rgba;            // &rgba_fb[frame][0][x] or nullptr
    int W, H, x, lane;
    uint32_t skycol;
};

template <bool kRgba>
__device__ __forceinline__ void put_px(const RasterCtx &c, uint8_t *p8, uint32_t *p32, bool on, uint32_t v) {
    if (on) {
#if defined(B2D_STORE_CS)
        __stcs(p8, (uint8_t)v);       // A/B (profiles/README.md): streaming store, keeps frame bytes from displacing texels
#elif defined(B2D_STORE_WT)
        __stwt(p8, (uint8_t)v);
#else
        *p8 = (uint8_t)v;
#endif
        if (kRgba) *p32 = lds_u32(c.pal_s + 4u * v);
    }
}

__host__ __device__ __forceinline__ bool tex_interleaved(const TexRec &T) { return b2d::tex_interleaved(T.h, T.texel_off); }

// Rows are produced in batches of kBatch: all texel loads first, then all colormap lookups, then the stores.
// Issuing the independent loads back to back keeps kBatch of them in flight per warp (the one-row-at-a-time
// form serialises load -> lookup -> store and leaves the LSU idle while each warp waits).  `full` (warp-uniform)
// says every lane owns all rows of the batch, so the stores need no predicate.
constexpr int kBatch = 8;
#if defined(B2D_FLAT_BATCH)
constexpr int kFlatBatch = B2D_FLAT_BATCH;     // A/B: gathers in flight per warp in the flat loop
#else
constexpr int kFlatBatch = 8;
#endif

template <bool kRgba, int kW, int kBatch = ::b2d::kBatch>
__device__ __forceinline__ void store_batch(const RasterCtx &c, uint8_t *p8, uint32_t *p32, const uint32_t (&v)[kBatch],
                                            int y, int ya, int yb, bool full) {
    const int Wc = kW ? kW : c.W;
    if (full) {
#pragma unroll
        for (int k = 0; k < kBatch; k++) put_px<kRgba>(c, p8 + (size_t)k * Wc, kRgba ? p32 + (size_t)k * Wc : nullptr, true, v[k]);
    } else {
        const uint32_t m = row_mask(y, ya, yb, kBatch);      // one bit test per row instead of two compares
#pragma unroll
        for (int k = 0; k < kBatch; k++)
            put_px<kRgba>(c, p8 + (size_t)k * Wc, kRgba ? p32 + (size_t)k * Wc : nullptr, (m >> k) & 1u, v[k]);
    }
}

// rows [ya, yb) of this lane's column := void (index 0); lanes with ya >= yb idle
template <bool kRgba, int kW>
__device__ __forceinline__ void fill_void_warp(const RasterCtx &c, int ya, int yb) {
    bool act = ya < yb;
    int y0 = __reduce_min_sync(kFull, act ? ya : 0x7FFFFFFF);
    int y1 = __reduce_max_sync(kFull, act ? yb : 0);
    if (y0 >= y1) return;
    const int Wc = kW ? kW : c.W;
    uint8_t *p8 = c.fb + (size_t)y0 * Wc;
    uint32_t *p32 = kRgba ? c.rgba + (size_t)y0 * Wc : nullptr;
#pragma unroll 1
    for (int y = y0; y < y1; y++, p8 += Wc, p32 += Wc) put_px<kRgba>(c, p8, p32, y >= ya && y < yb, 0u);
}

template <bool kRgba, int kW>
__device__ __forceinline__ void draw_sky_warp(const RasterCtx &c, int ya, int yb) {
    const DeviceScene &sc = *c.sc;
    if (sc.sky_tex < 0) { fill_void_warp<kRgba, kW>(c, ya, yb); return; }
    const TexRec T = sc.tex[sc.sky_tex];
    const bool inter = tex_interleaved(T);
    const uint8_t *px = sc.lit_texels + T.texel_off + (inter ? 4u * c.skycol : c.skycol);   // light row 0
    const uint32_t w4 = 4u * T.w;
    bool act = ya < yb;
    int y0 = __reduce_min_sync(kFull, act ? ya : 0x7FFFFFFF);
    int y1 = __reduce_max_sync(kFull, act ? yb : 0);
    if (y0 >= y1) return;
    const int Wc = kW ? kW : c.W;
    uint8_t *p8 = c.fb + (size_t)y0 * Wc;
    uint32_t *p32 = kRgba ? c.rgba + (size_t)y0 * Wc : nullptr;
#pragma unroll 2
    for (int y = y0; y < y1; y++, p8 += Wc, p32 += Wc) {
        const uint32_t r = sc.skyrow[y];                                   // warp-uniform table entry
        put_px<kRgba>(c, p8, p32, y >= ya && y < yb, tex_ld(px + (inter ? (r >> 2) * w4 + (r & 3u) : r * T.w)));
    }
}

template <bool kRgba, int kW>
__device__ __forceinline__ void draw_plane_warp(const RasterCtx &c, const FrameConst &fc, const View &vw,
                                                int ya, int yb, int32_t h, int32_t flat, int lightb,
                                                bool visible) {
    const DeviceScene &sc = *c.sc;
    if (!__any_sync(kFull, ya < yb)) return;
    if (!visible) { fill_void_warp<kRgba, kW>(c, ya, yb); return; }
    if (flat == kFlatSky) { draw_sky_warp<kRgba, kW>(c, ya, yb); return; }
    if (flat < 0 || flat >= sc.nflats) { fill_void_warp<kRgba, kW>(c, ya, yb); return; }
    // the pre-lit flats start at a multiple of 4 GiB (b2d_api.cu alloc_aligned_4g): a texel's address is {high word,
    // 32-bit offset} -- no 64-bit add per pixel
    const uint64_t px_hi = reinterpret_cast<uint64_t>(sc.lit_flats) & 0xFFFFFFFF00000000ull;   // low word is zero: tell the compiler
    const uint32_t habs = plane_habs(h, fc.pose.z);
    bool act = ya < yb;
    int y0 = __reduce_min_sync(kFull, act ? ya : 0x7FFFFFFF);
    int y1 = __reduce_max_sync(kFull, act ? yb : 0);
    const int full_lo = __reduce_max_sync(kFull, act ? ya : 0x7FFFFFFF);
    const int full_hi = __reduce_min_sync(kFull, act ? yb : 0);
    const int Wc = kW ? kW : c.W;
    uint8_t *p8 = c.fb + (size_t)y0 * Wc;
    uint32_t *p32 = kRgba ? c.rgba + (size_t)y0 * Wc : nullptr;
    const uint32_t bu = (uint32_t)fc.pose.x << 10, bv = (uint32_t)fc.pose.y << 10;
    const uint32_t ax = (uint32_t)c.dir.ax, ay = (uint32_t)c.dir.ay;
    for (int yc = y0; yc < y1; yc += 32) {
        // lane j prepares the constants of row yc+j (64-bit maths, once per row per warp): view depth and plane offset
        int yy = yc + c.lane;
        if (yy < y1) {
            const PlaneRow pr = plane_row(habs, sc.yslope[yy]);
            // plane + flat offset / 64: the texel offset is then two instructions, LEA.HI (cm6 + (U >> 26)) and a funnel
            // shift ((.) << 6 | V >> 26)
            c.rowz[c.lane] = make_uint2(pr.z8q, (sc.lit_flat_stride >> 6) * (uint32_t)light_row(lightb, pr.z8) + 64u * (uint32_t)flat);
        }
        __syncwarp();
        const int rows = min(32, y1 - yc);
        int j = 0;
        for (; j + kFlatBatch <= rows; j += kFlatBatch, p8 += (size_t)kFlatBatch * Wc, p32 += (size_t)kFlatBatch * Wc) {
            uint32_t v[kFlatBatch];
#pragma unroll
            for (int k = 0; k < kFlatBatch; k++) {
                const uint2 rz = c.rowz[j + k];                               // shared-memory broadcast: one 64-bit word per row
                v[k] = tex_ld(reinterpret_cast<const uint8_t *>(px_hi | flat_offset(rz.y, bu + rz.x * ax, bv + rz.x * ay)));   // always in bounds
            }
            const int y = yc + j;
            store_batch<kRgba, kW, kFlatBatch>(c, p8, p32, v, y, ya, yb, y >= full_lo && y + kFlatBatch <= full_hi);
        }
        for (; j < rows; j++, p8 += Wc, p32 += Wc) {
            const uint2 rz = c.rowz[j];
            const int y = yc + j;
            put_px<kRgba>(c, p8, p32, y >= ya && y < yb,
                          tex_ld(reinterpret_cast<const uint8_t *>(px_hi | flat_offset(rz.y, bu + rz.x * ax, bv + rz.x * ay))));
        }
        __syncwarp();
    }
}

// Magnified wall columns (every lane's texture step <= kWallFast8 / kWallFast16, b2d_math.cuh): R rows per batch from
// two aligned word loads, the row quad tracked incrementally -- per pixel one IMAD + one shift (byte index), one PRMT, one store.
template <bool kRgba, int kW, int R>
__device__ __forceinline__ void wall_fast_loop(const RasterCtx &c, const uint8_t *plq, uint32_t w4, uint32_t nq, uint32_t q,
                                               uint32_t acc, uint32_t ts29, int y0, int y1, int ya, int yb,
                                               int full_lo, int full_hi) {
    const int Wc = kW ? kW : c.W;
    uint8_t *p8 = c.fb + (size_t)y0 * Wc;
    uint32_t *p32 = kRgba ? c.rgba + (size_t)y0 * Wc : nullptr;
    asm("" : "+l"(plq));       // keep the column's plane pointer whole: one IMAD.WIDE per load instead of re-adding the base
#pragma unroll 1
    for (int y = y0; y < y1; y += R, p8 += (size_t)R * Wc, p32 += (size_t)R * Wc) {
        const uint32_t q1 = q + 1u == nq ? 0u : q + 1u;
        const uint32_t w0 = tex_ld(reinterpret_cast<const uint32_t *>(plq + (size_t)q * w4));
        const uint32_t w1 = tex_ld(reinterpret_cast<const uint32_t *>(plq + (size_t)q1 * w4));
        if (y >= full_lo && y + R <= full_hi) {
#pragma unroll
            for (int k = 0; k < R; k++) {
                uint32_t v = pick_byte(w0, w1, wall_sel(acc, ts29, (uint32_t)k));
                if (kRgba) v &= 0xFFu;
                put_px<kRgba>(c, p8 + (size_t)k * Wc, kRgba ? p32 + (size_t)k * Wc : nullptr, true, v);
            }
        } else {
            const uint32_t m = row_mask(y, ya, yb, R);
#pragma unroll
            for (int k = 0; k < R; k++) {
                uint32_t v = pick_byte(w0, w1, wall_sel(acc, ts29, (uint32_t)k));
                if (kRgba) v &= 0xFFu;
                put_px<kRgba>(c, p8 + (size_t)k * Wc, kRgba ? p32 + (size_t)k * Wc : nullptr, (m >> k) & 1u, v);
            }
        }
        wall_advance(acc, q, ts29, (uint32_t)R, nq);
    }
}

template <bool kRgba, int kW>
__device__ __forceinline__ void draw_wall_warp(const RasterCtx &c, const FrameConst &fc, int ya, int yb,
                                               int32_t tex, int32_t tA, int32_t hA, int32_t ucol,
                                               int32_t iscale, int row) {
    const DeviceScene &sc = *c.sc;
    if (!__any_sync(kFull, ya < yb)) return;
    if (tex < 0 || tex >= sc.ntex) { fill_void_warp<kRgba, kW>(c, ya, yb); return; }
    const TexRec T = sc.tex[tex];
    bool act = ya < yb;
    const uint32_t col = (uint32_t)floormod32(ucol, (int32_t)T.w);
    const uint8_t *pl = sc.lit_texels + (size_t)row * sc.lit_texel_stride + T.texel_off;   // this lane's light plane
    const uint32_t tstep = (uint32_t)(iscale >> 4);
    int y0 = __reduce_min_sync(kFull, act ? ya : 0x7FFFFFFF);
    int y1 = __reduce_max_sync(kFull, act ? yb : 0);
    const int full_lo = __reduce_max_sync(kFull, act ? ya : 0x7FFFFFFF);
    const int full_hi = __reduce_min_sync(kFull, act ? yb : 0);
    uint32_t t = (uint32_t)wall_tbase(tA, hA, fc.pose.z, c.H, iscale) + (uint32_t)y0 * tstep;
    const int Wc = kW ? kW : c.W;
    uint8_t *p8 = c.fb + (size_t)y0 * Wc;
    uint32_t *p32 = kRgba ? c.rgba + (size_t)y0 * Wc : nullptr;
    // Texel loads are unconditional (t is a bounded linear function of y for every lane, so the row index
    // is always inside the texture); only the store is predicated.  That keeps the loops branch-free.
    if (tex_interleaved(T) && T.h >= 8u && !(sc.tune & 1u)) {
        // magnified columns: the whole piece runs on the incremental path when every lane qualifies (inactive lanes step 0)
        const uint32_t ts = act ? tstep : 0u;
        const bool f16 = !(sc.tune & 2u) && __all_sync(kFull, ts <= kWallFast16);
        if (f16 || __all_sync(kFull, ts <= kWallFast8)) {
            const uint32_t tt = act ? t : 0u;
            const uint32_t r0 = wall_row((int32_t)tt, T.h, T.hmagic, T.hbias);
            const uint8_t *plq = pl + 4u * col;
            if (f16) wall_fast_loop<kRgba, kW, 16>(c, plq, 4u * T.w, T.h >> 2, r0 >> 2, wall_acc29(tt, r0), ts << 13, y0, y1, ya, yb, full_lo, full_hi);
            else wall_fast_loop<kRgba, kW, 8>(c, plq, 4u * T.w, T.h >> 2, r0 >> 2, wall_acc29(tt, r0), ts << 13, y0, y1, ya, yb, full_lo, full_hi);
            return;
        }
    }
    if (tex_interleaved(T)) {
        // 4-row interleaved plane.  Rows of a batch: u_k = asr(t + k*tstep, 16), texture row = floormod(u_k, h).
        // With r0 the row of the first pixel, pixel k reads byte (r0 & 3) + u_k - u_0 of the 8 bytes made of row
        // quad r0 >> 2 and the next one (wrapping at h, a multiple of 4).  If that byte index stays below 8 for
        // every lane -- the column is magnified -- two aligned word loads serve the whole batch.
        const uint32_t colb = 4u * col, w4 = 4u * T.w;
        for (int y = y0; y < y1; y += kBatch, p8 += (size_t)kBatch * Wc, p32 += (size_t)kBatch * Wc, t += (uint32_t)kBatch * tstep) {
            const uint32_t r0 = wall_row((int32_t)t, T.h, T.hmagic, T.hbias);
            // acc = t with its integer part replaced by r0 & 3: byte index of pixel k = (acc + k*tstep) >> 16
            const uint32_t acc = wall_acc(t, r0);
            const uint32_t b7 = (acc + 7u * tstep) >> 16;
            uint32_t v[kBatch];
            if (__all_sync(kFull, b7 < 8u)) {
                const uint32_t q0 = r0 >> 2, q1 = next_quad(q0, T.h);
                const uint32_t w0 = tex_ld(reinterpret_cast<const uint32_t *>(pl + (q0 * w4 + colb)));
                const uint32_t w1 = tex_ld(reinterpret_cast<const uint32_t *>(pl + (q1 * w4 + colb)));
#pragma unroll
                for (int k = 0; k < kBatch; k++) {
                    v[k] = pick_byte(w0, w1, (acc + (uint32_t)k * tstep) >> 16);
                    if (kRgba) v[k] &= 0xFFu;
                }
            } else {
                // less magnified: re-anchor after 4 rows -- two word pairs serve the batch as long as each half
                // stays inside two row quads (up to ~1.3 texture rows per screen row); byte loads otherwise
                const uint32_t t4 = t + 4u * tstep;
                const uint32_t r4 = wall_row((int32_t)t4, T.h, T.hmagic, T.hbias);
                const uint32_t acc4 = wall_acc(t4, r4);
                if (__all_sync(kFull, ((acc + 3u * tstep) >> 16) < 8u && ((acc4 + 3u * tstep) >> 16) < 8u)) {
                    const uint32_t q0 = r0 >> 2, q4 = r4 >> 2;
                    const uint32_t a0 = tex_ld(reinterpret_cast<const uint32_t *>(pl + (q0 * w4 + colb)));
                    const uint32_t a1 = tex_ld(reinterpret_cast<const uint32_t *>(pl + (next_quad(q0, T.h) * w4 + colb)));
                    const uint32_t b0 = tex_ld(reinterpret_cast<const uint32_t *>(pl + (q4 * w4 + colb)));
                    const uint32_t b1 = tex_ld(reinterpret_cast<const uint32_t *>(pl + (next_quad(q4, T.h) * w4 + colb)));
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        v[k] = pick_byte(a0, a1, (acc + (uint32_t)k * tstep) >> 16);
                        v[k + 4] = pick_byte(b0, b1, (acc4 + (uint32_t)k * tstep) >> 16);
                        if (kRgba) { v[k] &= 0xFFu; v[k + 4] &= 0xFFu; }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kBatch; k++) {
                        const uint32_t rk = wall_row((int32_t)(t + (uint32_t)k * tstep), T.h, T.hmagic, T.hbias);
                        v[k] = tex_ld(pl + ((rk >> 2) * w4 + colb + (rk & 3u)));
                    }
                }
            }
            // rows past y1 belong to no lane: the predicated form drops them, so there is no tail loop
            store_batch<kRgba, kW>(c, p8, p32, v, y, ya, yb, y >= full_lo && y + kBatch <= full_hi);
        }
    } else {
        // heights that are not a multiple of 4 (patches used as textures, odd PWAD content): rare, kept small
        const uint8_t *px = pl + col;
#pragma unroll 1
        for (int y = y0; y < y1; y++, p8 += Wc, p32 += Wc, t += tstep)
            put_px<kRgba>(c, p8, p32, y >= ya && y < yb, tex_ld(px + wall_row((int32_t)t, T.h, T.hmagic, T.hbias) * T.w));
    }
}

// Deferred masked entries (33 words: worklist index + one packed window per lane) live in a per-launch arena of
// kMaskedChunk-entry chunks handed out by an atomic counter; a warp keeps the ids of its chunks in shared memory.  Memory
// follows what frames actually defer (a few entries per strip) instead of strips x cap x batch.
__device__ __forceinline__ uint32_t *masked_entry(const DeviceScene &sc, const uint32_t *chunks, int e) {
    return sc.masked_list + ((size_t)chunks[e / kMaskedChunk] * kMaskedChunk + (size_t)(e % kMaskedChunk)) * 33u;
}
// next entry of this warp, or nullptr when the strip's cap or the arena is exhausted (status bit 8: frames incomplete)
__device__ __forceinline__ uint32_t *masked_push(const DeviceScene &sc, uint32_t *chunks, int &mcount, int lane) {
    bool ok = mcount < sc.masked_cap;
    if (ok && mcount % kMaskedChunk == 0) {
        uint32_t id = 0;
        if (lane == 0) id = atomicAdd(sc.masked_counter, 1u);
        id = __shfl_sync(kFull, id, 0);
        ok = id < sc.masked_chunks;
        if (ok && lane == 0) chunks[mcount / kMaskedChunk] = id;
        __syncwarp();
    }
    if (!ok) {
        if (lane == 0) atomicOr(sc.status_flag, 8);
        return nullptr;
    }
    return masked_entry(sc, chunks, mcount++);
}

// Back-to-front pass over the masked middle textures this strip deferred during the solid pass.  Each entry
// holds the worklist index and, per lane, the clip window [ya, yb) that was open behind the seg when the
// front-to-back walk reached it (the per-column silhouette of everything nearer).  Texels whose opacity plane
// is 0 leave the pixel as the solid pass drew it (static.frag:21-22).  Kept out of line so that the
// register allocation of the solid pass is not affected.
template <bool kRgba, int kW>
__device__ __noinline__ void masked_pass(const DeviceScene &sc, const View &vw, int32_t pose_z, uint8_t *fb,
                                         uint32_t *rgba, uint32_t pal_s, int x, int lane,
                                         const SegFrame *wl, const uint32_t *chunks, int count) {
    // everything arrives by value (or points at kernel parameters): taking the address of the solid pass's
    // register-resident context would force it into local memory
    RasterCtx c;
    c.sc = &sc; c.pal_s = pal_s; c.fb = fb; c.rgba = rgba;
    c.W = vw.W; c.H = vw.H; c.x = x; c.lane = lane;
    const int Wc = kW ? kW : c.W;
    for (int e = count - 1; e >= 0; e--) {
        const uint32_t *ml = masked_entry(sc, chunks, e);
        const uint32_t k = ml[0];
        const uint32_t packed = ml[1 + c.lane];
        int ya = (int)(packed & 0xFFFFu), yb = (int)(packed >> 16);
        const SegFrame sf = wl[k];
        int32_t tex, tA, hA, ucol = 0, iscale = 1, row = 0;
        if (sf.seg < 0) {
            // decoration sprite (billboard at constant depth): everything but the column is warp-uniform
            const SpriteRec P = sc.sprites[-1 - sf.seg];
            if (P.tex < 0 || P.tex >= sc.ntex) continue;
            const int32_t sw = (int32_t)sc.tex[P.tex].w, sh = (int32_t)sc.tex[P.tex].h;
            SpriteFrame sp;
            sp.cx = sf.Nc; sp.cz = sf.Nx;
            int64_t scale = ((int64_t)vw.FY2 << 25) / sp.cz;
            const int64_t cap = (int64_t)vw.FY2 << 17;
            if (scale > cap) scale = cap;
            iscale = (int32_t)clampv<int64_t>(((int64_t)1 << 38) / scale, 1, 1 << 23);
            int64_t z8 = ((int64_t)iscale * vw.FY2) >> 18;
            row = light_row_sprite(P.light, z8 > 65535 ? 65535 : (int32_t)z8);
            tex = P.tex; tA = 0; hA = P.low + sh;
            if (ya < yb) {
                ya = max(ya, yrow(P.low + sh, (int32_t)scale, pose_z, c.H));
                yb = min(yb, yrow(P.low, (int32_t)scale, pose_z, c.H));
                ucol = sprite_column(sp, vw, c.x, sw);
            }
        } else {
            const SegRec S = sc.segs[sf.seg];
            if (S.mid < 0 || S.mid >= sc.nmids) continue;
            const MidRec M = sc.mids[S.mid];
            tex = M.tex; tA = M.t_high; hA = M.high;
            ColumnEval ce = {0u, 1, 1, 0};
            if (ya < yb && column_eval(sf, vw, c.x, ce)) {
                ya = max(ya, yrow(M.high, ce.scale, pose_z, c.H));
                yb = min(yb, yrow(M.low, ce.scale, pose_z, c.H));
                ucol = S.uoff + (int32_t)(((uint64_t)ce.s24 * (uint32_t)S.len_q12) >> 36);
                iscale = ce.iscale;
                row = light_row(S.light, ce.z8);
            } else {
                ya = yb = 0;
            }
        }
        if (tex < 0 || tex >= sc.ntex) continue;
        const TexRec T = sc.tex[tex];
        const bool act = ya < yb;
        int y0 = __reduce_min_sync(kFull, act ? ya : 0x7FFFFFFF);
        int y1 = __reduce_max_sync(kFull, act ? yb : 0);
        if (y0 >= y1) continue;
        const uint32_t col = (uint32_t)floormod32(ucol, (int32_t)T.w);
        // colour from this lane's pre-lit plane, opacity from plane 32 (same layout, written by the pre-light kernel
        // for textures with holes)
        const bool inter = tex_interleaved(T);
        const bool has_mask = T.mask_off != 0xFFFFFFFFu;
        const uint8_t *pl = sc.lit_texels + (size_t)row * sc.lit_texel_stride + T.texel_off;
        const uint8_t *pm = sc.lit_texels + (size_t)32 * sc.lit_texel_stride + T.texel_off;
        const uint32_t tstep = (uint32_t)(iscale >> 4);
        uint32_t t = (uint32_t)wall_tbase(tA, hA, pose_z, c.H, iscale) + (uint32_t)y0 * tstep;
        uint8_t *p8 = c.fb + (size_t)y0 * Wc;
        uint32_t *p32 = kRgba ? c.rgba + (size_t)y0 * Wc : nullptr;
        const uint32_t len = act ? (uint32_t)(yb - ya) : 0u;   // the clipped window can be inverted (ya > yb): no rows then
        if (inter) {
            // batches of 8 rows as in draw_wall_warp: magnified columns (sprites nearly always are) read two colour
            // words and two opacity words per batch
            const uint32_t colb = 4u * col, w4 = 4u * T.w;
#pragma unroll 1
            for (int y = y0; y < y1; y += kBatch, p8 += (size_t)kBatch * Wc, p32 += (size_t)kBatch * Wc, t += (uint32_t)kBatch * tstep) {
                const uint32_t r0 = wall_row((int32_t)t, T.h, T.hmagic, T.hbias);
                const uint32_t acc = wall_acc(t, r0);
                const uint32_t d = (uint32_t)(y - ya);
                uint32_t v[kBatch], o[kBatch];
                if (__all_sync(kFull, ((acc + 7u * tstep) >> 16) < 8u)) {
                    const uint32_t o0 = (r0 >> 2) * w4 + colb, o1 = next_quad(r0 >> 2, T.h) * w4 + colb;
                    const uint32_t w0 = tex_ld(reinterpret_cast<const uint32_t *>(pl + o0));
                    const uint32_t w1 = tex_ld(reinterpret_cast<const uint32_t *>(pl + o1));
                    uint32_t m0 = 0x01010101u, m1 = 0x01010101u;
                    if (has_mask) {
                        m0 = tex_ld(reinterpret_cast<const uint32_t *>(pm + o0));
                        m1 = tex_ld(reinterpret_cast<const uint32_t *>(pm + o1));
                    }
#pragma unroll
                    for (int k = 0; k < kBatch; k++) {
                        const uint32_t bk = (acc + (uint32_t)k * tstep) >> 16;
                        v[k] = pick_byte(w0, w1, bk) & 0xFFu;
                        o[k] = pick_byte(m0, m1, bk) & 0xFFu;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kBatch; k++) {
                        const uint32_t rk = wall_row((int32_t)(t + (uint32_t)k * tstep), T.h, T.hmagic, T.hbias);
                        const uint32_t off = (rk >> 2) * w4 + colb + (rk & 3u);
                        v[k] = tex_ld(pl + off);
                        o[k] = has_mask ? (uint32_t)tex_ld(pm + off) : 1u;
                    }
                }
#pragma unroll
                for (int k = 0; k < kBatch; k++)
                    put_px<kRgba>(c, p8 + (size_t)k * Wc, kRgba ? p32 + (size_t)k * Wc : nullptr, d + (uint32_t)k < len && o[k] != 0u, v[k]);
            }
        } else {
#pragma unroll 1
            for (int y = y0; y < y1; y++, p8 += Wc, p32 += Wc, t += tstep) {
                const uint32_t off = wall_row((int32_t)t, T.h, T.hmagic, T.hbias) * T.w + col;
                const bool on = (uint32_t)(y - ya) < len && (!has_mask || tex_ld(pm + off) != 0);
                put_px<kRgba>(c, p8, p32, on, tex_ld(pl + off));
            }
        }
    }
}

template <bool kRgba, int kMinBlocks, int kW, int kWarps, bool kMasked>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
b2d_raster_kernel(const __grid_constant__ DeviceScene sc, const __grid_constant__ View vw, const FrameConst *__restrict__ frames,
                  const SegFrame *__restrict__ work, int stride, int n, int strips,
                  uint8_t *__restrict__ index_fb, uint32_t *__restrict__ rgba_fb) {
    __shared__ uint32_t s_pal[kRgba ? 256 : 1];
    __shared__ uint2 s_rowz[kWarps][32];
    __shared__ uint32_t s_chunks[kMasked ? kWarps : 1][kMasked ? kMaskedCapMax / kMaskedChunk : 1];
    if (kRgba) {   // the palette into shared memory (colours come pre-lit from global memory: no colormap here)
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_pal[i] = sc.palette[i];
        __syncthreads();
    }

    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= (long long)n * strips) return;
    const int frame = (int)(gw / strips), strip = (int)(gw % strips);
    const int W = kW ? kW : vw.W, H = vw.H;
    const int x0 = strip * 32, x = x0 + lane;
    const bool inside = x < W;

    const FrameConst fc = frames[frame];
    RasterCtx c;
    c.sc = &sc;
    c.pal_s = (uint32_t)__cvta_generic_to_shared(s_pal);
    c.rowz = s_rowz[threadIdx.x >> 5];
    c.dir = plane_dir(fc, vw, inside ? x : 0, sc.invF);
    c.fb = index_fb + (size_t)frame * W * H + (inside ? x : 0);
    c.rgba = kRgba ? rgba_fb + (size_t)frame * W * H + (inside ? x : 0) : nullptr;
    c.W = W; c.H = H; c.x = x; c.lane = lane;
    c.skycol = 0;
    if (sc.sky_tex >= 0 && inside) c.skycol = umulhi32(sky_u32(x, vw, fc.pose.angle), sc.tex[sc.sky_tex].w);

    int ct = 0, cb = inside ? H : 0;              // open window [ct, cb) of this lane's column
    uint32_t *chunks = s_chunks[kMasked ? (threadIdx.x >> 5) : 0];
    const bool defer = kMasked && sc.masked_list != nullptr;
    int mcount = 0;
    const SegFrame *wl = work + (size_t)frame * stride;
    const int count = fc.count;
    bool done = false;

    for (int k0 = 0; k0 < count && !done; k0 += 32) {
        int k = k0 + lane;
        bool overlap = false;
        if (k < count) {
            int xlo = wl[k].xlo, xhi = wl[k].xhi;
            overlap = xhi >= x0 && xlo <= x0 + 31;
        }
        unsigned m = __ballot_sync(kFull, overlap);
        while (m) {
            int j = __ffs(m) - 1;
            m &= m - 1;
            if (!__any_sync(kFull, ct < cb)) { done = true; break; }
            const SegFrame sf = wl[k0 + j];                              // warp-uniform 64 B
            bool in = inside && ct < cb && x >= sf.xlo && x <= sf.xhi;
            if (!__any_sync(kFull, in)) continue;
            if (sf.seg < 0) {
                // decoration sprite: remember the windows open right now; it is drawn in the masked pass
                if (defer) {
                    uint32_t *ml = masked_push(sc, chunks, mcount, lane);
                    if (ml) {
                        if (lane == 0) ml[0] = (uint32_t)(k0 + j);
                        ml[1 + lane] = in ? ((uint32_t)ct | ((uint32_t)cb << 16)) : 0u;
                    }
                }
                continue;
            }
            ColumnEval ce = {0u, 1, 1, 0};
            bool ok = in && column_eval(sf, vw, x, ce);
            if (!__any_sync(kFull, ok)) continue;

            const SegRec S = sc.segs[sf.seg];
            const SectorRec SF = sc.sectors[S.front];
            const int32_t fcl = SF.ceil, ffl = SF.floor;
            const bool two = S.flags & kSegTwoSided;
            const bool ceil_vis = ((int64_t)fcl << 16) > fc.pose.z || SF.ceil_flat == kFlatSky;
            const bool floor_vis = ((int64_t)ffl << 16) < fc.pose.z || SF.floor_flat == kFlatSky;

            // per-lane rows: y1 wall top, y2 end of upper, y3 start of lower, y4 start of floor
            int y1 = ct, y2 = ct, y3 = ct, y4 = ct, yend = ct, row = 0;
            int32_t ucol = 0;
            if (ok) {
                row = light_row(S.light, ce.z8);
                ucol = S.uoff + (int32_t)(((uint64_t)ce.s24 * (uint32_t)S.len_q12) >> 36);
                int yfc = yrow(fcl, ce.scale, fc.pose.z, H), yff = yrow(ffl, ce.scale, fc.pose.z, H);
                y1 = clampv(yfc, ct, cb);
                if (!two) {
                    y2 = clampv(yff, y1, cb);       // one-sided: [y1,y2) is the middle texture
                    y3 = y2; y4 = y2;
                } else {
                    int yot = yrow(S.otop, ce.scale, fc.pose.z, H), yob = yrow(S.obot, ce.scale, fc.pose.z, H);
                    y2 = clampv(yot, y1, cb);
                    y3 = clampv(yob, y2, cb);
                    y4 = clampv(yff, y3, cb);
                }
                yend = cb;
            }
            // The two flat spans (ceiling [ct, y1), floor [y4, cb)) and the two wall pieces go through ONE inlined
            // copy of each span routine (loops kept rolled): the routines are large, and the kernel's speed depends
            // on its hot code staying resident in the instruction cache.
#pragma unroll 1
            for (int pz = 0; pz < 2; pz++) {
                const bool top = pz == 0;
                draw_plane_warp<kRgba, kW>(c, fc, vw, ok ? (top ? ct : y4) : 0, ok ? (top ? y1 : yend) : 0,
                                                    top ? fcl : ffl, top ? SF.ceil_flat : SF.floor_flat, SF.light,
                                                    top ? ceil_vis : floor_vis);
            }
#pragma unroll 1
            for (int pw = 0; pw < 2; pw++) {
                const bool upper = pw == 0;            // A: upper (two-sided) or the one-sided middle; B: lower
                if (upper ? (two && !(S.otop < fcl)) : !(two && S.obot > ffl)) continue;
                draw_wall_warp<kRgba, kW>(c, fc, ok ? (upper ? y1 : y3) : 0, ok ? (upper ? y2 : y4) : 0,
                                                   upper ? S.texA : S.texB, upper ? S.tA : S.tB, upper ? S.hA : S.hB,
                                                   ucol, ce.iscale, row);
            }
            if (ok) {
                if (!two || y2 >= y3) { ct = H; cb = 0; }
                else { ct = y2; cb = y3; }
            }
            if (defer && two && S.mid >= 0) {
                // defer the masked middle texture: remember the window that is open behind this seg
                const bool keep = ok && y2 < y3;
                if (__any_sync(kFull, keep)) {
                    uint32_t *ml = masked_push(sc, chunks, mcount, lane);
                    if (ml) {
                        if (lane == 0) ml[0] = (uint32_t)(k0 + j);
                        ml[1 + lane] = keep ? ((uint32_t)y2 | ((uint32_t)y3 << 16)) : 0u;
                    }
                }
            }
        }
    }
    // whatever is still open is void
    fill_void_warp<kRgba, kW>(c, inside ? ct : 0, inside ? cb : 0);
    if (kMasked && mcount > 0) {
        __syncwarp();
        masked_pass<kRgba, kW>(sc, vw, fc.pose.z, c.fb, c.rgba, c.pal_s, x, lane, wl, chunks, mcount);
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel 3: palette LUT
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
b2d_palette_kernel(const uint32_t *__restrict__ palette, const uint8_t *__restrict__ index,
                   uint32_t *__restrict__ rgba, size_t n_pixels) {
    __shared__ uint32_t s_pal[256];
    s_pal[threadIdx.x] = palette[threadIdx.x];
    __syncthreads();
    // 128-bit path only for 16-byte aligned buffers (a caller may pass &index_fb[i*W*H] with W*H % 16 != 0)
    const bool aligned = ((reinterpret_cast<uintptr_t>(index) | reinterpret_cast<uintptr_t>(rgba)) & 15) == 0;
    const size_t nvec = aligned ? n_pixels / 16 : 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        uint4 in = __ldcs(reinterpret_cast<const uint4 *>(index) + i);
        uint32_t wds[4] = {in.x, in.y, in.z, in.w};
        uint4 *out = reinterpret_cast<uint4 *>(rgba) + 4 * i;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint4 o;
            o.x = s_pal[wds[q] & 0xFF];
            o.y = s_pal[(wds[q] >> 8) & 0xFF];
            o.z = s_pal[(wds[q] >> 16) & 0xFF];
            o.w = s_pal[wds[q] >> 24];
            __stcs(out + q, o);
        }
    }
    // tail (n_pixels not a multiple of 16)
    for (size_t p = nvec * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pixels; p += stride)
        rgba[p] = s_pal[index[p]];
}

}  // namespace

// ------------------------------------------------------------------------------------------------
size_t walk_smem_per_warp(const DeviceScene &sc) { return walk_layout(sc.nverts, sc.nsegs, sc.nnodes, sc.nss, sc.nsprites).total; }

cudaError_t launch_walk(const DeviceScene &sc, const View &vw, const Pose *d_poses, int n,
                        FrameConst *d_frames, SegFrame *d_work, int stride, cudaStream_t stream, bool background) {
    if (n <= 0) return cudaSuccess;
    const size_t smem = walk_smem_per_warp(sc);          // one frame per CTA
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {   // per device and cheap: set it on every launch that needs the opt-in
        cudaError_t e = cudaFuncSetAttribute(b2d_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // background (b2d_walk_device: the walk of the NEXT batch, meant to run under another batch's raster): a persistent grid
    // of one CTA per SM.  It takes n/148 frame latencies instead of one, but holds 1/8 of the register file instead of
    // 7/8, so the raster keeps 28 of its 32 warps per SM while the walk hides behind it (measured: profiles/README.md).
    // B2D_TUNE bit 2 (4) turns it off for A/B.
    const int blocks = (background && !(sc.tune & 4u) && n > 148) ? 148 : n, warps = 4;
    b2d_walk_kernel<<<blocks, warps * 32, smem, stream>>>(sc, vw, d_poses, n, d_frames, d_work, stride);
    return cudaGetLastError();
}

cudaError_t launch_raster(const DeviceScene &sc, const View &vw, const FrameConst *d_frames,
                          const SegFrame *d_work, int stride, int n, uint8_t *d_index_fb,
                          uint32_t *d_rgba, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int strips = (vw.W + 31) / 32;
    // Launch shape (tuned on B200, profiles/README.md): ONE warp per CTA and 64 registers/thread, i.e. 32 CTAs =
    // 32 warps resident per SM.  Strips differ a lot in cost; with several warps per CTA the finished warps' slots
    // stay empty until the slowest warp of the CTA is done (1-warp CTAs: +10 % over 2 or 4, 8 and 16 lose more).
    // The frame width is a compile-time constant for the benchmark resolutions (immediate store offsets).
#if defined(B2D_RASTER_WARPS)         // A/B: more resident warps per SM at fewer registers per thread
    constexpr int kWarps = B2D_RASTER_WARPS, kMinBlocks = B2D_RASTER_MINBLOCKS;
#else
    constexpr int kWarps = 1, kMinBlocks = 32;
#endif
    const long long total_warps = (long long)n * strips;
    const int nblocks = (int)((total_warps + kWarps - 1) / kWarps);
    static const bool generic_w = getenv("B2D_RASTER_GENERIC_W") != nullptr;   // A/B knob for profiles/README.md
    const bool w1920 = vw.W == 1920 && !generic_w;
    // B2D_CARVEOUT=<percent of the SM's L1/shared array given to shared memory> (A/B): the raster uses ~300 B of it per warp
    static const int carve = getenv("B2D_CARVEOUT") ? atoi(getenv("B2D_CARVEOUT")) : -1;
#define B2D_RASTER_GO(RGBA, KW) do { \
    if ((sc.nmids > 0 || sc.nsprites > 0) && sc.masked_list) { \
        static const bool once = carve >= 0 && cudaFuncSetAttribute(b2d_raster_kernel<RGBA, kMinBlocks, KW, kWarps, true>, \
                                                                     cudaFuncAttributePreferredSharedMemoryCarveout, carve) == cudaSuccess; \
        (void)once; \
        b2d_raster_kernel<RGBA, kMinBlocks, KW, kWarps, true><<<nblocks, kWarps * 32, 0, stream>>>( \
            sc, vw, d_frames, d_work, stride, n, strips, d_index_fb, d_rgba); \
    } else { \
        static const bool once = carve >= 0 && cudaFuncSetAttribute(b2d_raster_kernel<RGBA, kMinBlocks, KW, kWarps, false>, \
                                                                     cudaFuncAttributePreferredSharedMemoryCarveout, carve) == cudaSuccess; \
        (void)once; \
        b2d_raster_kernel<RGBA, kMinBlocks, KW, kWarps, false><<<nblocks, kWarps * 32, 0, stream>>>( \
            sc, vw, d_frames, d_work, stride, n, strips, d_index_fb, d_rgba); } } while (0)
    if (d_rgba) { if (w1920) B2D_RASTER_GO(true, 1920); else B2D_RASTER_GO(true, 0); }
    else {
        const bool w3840 = vw.W == 3840 && !generic_w;      // BASELINE.json's 4K configuration (index frames only)
        if (w1920) B2D_RASTER_GO(false, 1920); else if (w3840) B2D_RASTER_GO(false, 3840); else B2D_RASTER_GO(false, 0);
    }
#undef B2D_RASTER_GO
    return cudaGetLastError();
}

namespace {
__global__ void __launch_bounds__(256)
b2d_prelight_kernel(const uint8_t *__restrict__ colormap, const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                    size_t n, size_t stride) {
    __shared__ uint8_t cm[32 * 256];
    for (int i = threadIdx.x; i < 32 * 256; i += blockDim.x) cm[i] = colormap[i];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t t = src[i];
#pragma unroll 4
        for (int r = 0; r < 32; r++) dst[(size_t)r * stride + i] = cm[r * 256 + t];
    }
}
}  // namespace

namespace {
// one CTA per texture: pre-lit copies in the layout tex_interleaved() selects
__global__ void __launch_bounds__(256)
b2d_prelight_tex_kernel(const uint8_t *__restrict__ colormap, const uint8_t *__restrict__ texels,
                        const TexRec *__restrict__ tex, int ntex, uint8_t *__restrict__ dst, size_t stride) {
    __shared__ uint8_t cm[32 * 256];
    for (int i = threadIdx.x; i < 32 * 256; i += blockDim.x) cm[i] = colormap[i];
    __syncthreads();
    for (int ti = blockIdx.x; ti < ntex; ti += gridDim.x) {
        const TexRec T = tex[ti];
        const bool inter = tex_interleaved(T);
        const uint32_t n = T.w * T.h;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t row = i / T.w, col = i - row * T.w;
            const uint32_t o = lit_index(inter, T.w, row, col);
            const uint32_t t = texels[T.texel_off + i];
            for (int r = 0; r < 32; r++) dst[(size_t)r * stride + T.texel_off + o] = cm[r * 256 + t];
            if (T.mask_off != 0xFFFFFFFFu) dst[(size_t)32 * stride + T.texel_off + o] = texels[T.mask_off + i];   // opacity
        }
    }
}
}  // namespace

cudaError_t launch_prelight_textures(const uint8_t *d_colormap, const uint8_t *d_texels, const TexRec *d_tex, int ntex,
                                     uint8_t *d_dst, size_t stride, cudaStream_t stream) {
    if (ntex <= 0) return cudaSuccess;
    b2d_prelight_tex_kernel<<<ntex < 148 * 8 ? ntex : 148 * 8, 256, 0, stream>>>(d_colormap, d_texels, d_tex, ntex, d_dst, stride);
    return cudaGetLastError();
}

cudaError_t launch_prelight(const uint8_t *d_colormap, const uint8_t *d_src, uint8_t *d_dst, size_t n, size_t stride,
                            cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    b2d_prelight_kernel<<<blocks, 256, 0, stream>>>(d_colormap, d_src, d_dst, n, stride);
    return cudaGetLastError();
}

cudaError_t launch_palette(const uint32_t *d_palette, const uint8_t *d_index, uint32_t *d_rgba,
                           size_t n_pixels, cudaStream_t stream) {
    if (n_pixels == 0) return cudaSuccess;
    size_t nvec = n_pixels / 16 + 1;
    int blocks = (int)((nvec + 255) / 256);
    const int cap = 148 * 16;
    if (blocks > cap) blocks = cap;
    b2d_palette_kernel<<<blocks, 256, 0, stream>>>(d_palette, d_index, d_rgba, n_pixels);
    return cudaGetLastError();
}

}  // namespace b2d
