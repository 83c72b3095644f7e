// boxblur_seg.cuh — shared by boxblur_seg_{h,v,ct}.cu: sm_100a "segment" kernels for vszip.BoxBlur on 8- and 16-bit integer clips.
//
// Integer BoxBlur has a closed form per pass (boxblur_seg_core.h; src/filters/boxblur_runtime.zig:10-41), so a line
// does not have to be walked by one thread.  Here a thread owns a segment of 60 consecutive samples of a line, keeps
// it in registers across ALL passes of the axis, and per pass only exchanges the r samples either side of the segment:
//
//   hseg_kernel    rows.  A warp owns a row (32 lanes x 60 samples = 1920; narrower planes pack 2 or 4 rows per warp).
//                  Rows are staged by the TMA engine: one cp.async.bulk (global -> shared, mbarrier completion) per row
//                  into a per-warp double buffer, one cp.async.bulk (shared -> global) per finished row.  No thread
//                  issues a global load or store; halo exchange is warp-local (shared memory + __syncwarp).
//   vseg_kernel    columns.  A CTA owns a 64-column strip of a plane; lane = pair of columns (one 32-bit word per row),
//                  warp = 60-row segment.  Global access is coalesced by construction (128 bytes per warp and row).
//   ctfused_kernel the comptime path (hradius == vradius <= 22, one pass each; src/filters/boxblur_comptime.zig:10-159)
//                  in ONE read and ONE write of the plane: a CTA streams down a band of rows keeping exact column sums in
//                  registers (8 columns per thread); the rows entering and leaving the window arrive through a TMA-fed
//                  shared-memory ring; the rounded means of 8 (16) rows go to shared memory and each warp then runs the
//                  horizontal closed form on one (two) of them.
//
// The radius is a template parameter (1..22: every comptime radius); other radii, float clips and planes wider
// than 1920 / taller than 1080 keep the streaming kernels of boxblur_kernels.cu.  8-bit clips run the same kernels: bytes are widened
// after the TMA load and narrowed before the store (template parameter U8).
#pragma once

#include <algorithm>

#include "boxblur_seg_core.h"
#include "common.h"
#include "filter.h"

namespace vsz {

namespace {

using namespace seg;

// --------------------------------------------------------------------------- job descriptors
struct SegPlane {
    size_t src_off, dst_off;
    int src_pitch, dst_pitch;
    int w, h;
    int cta_begin;
    int S;       // segments (threads) along the blurred axis
    int G;       // H: lanes per row (8 / 16 / 32)
    int rowbuf;  // H: bytes between staged rows
    int per_cta; // H: row groups per CTA; ctfused: rows per band
};
struct SegJob {
    const char* src;
    char* dst;
    size_t src_fs, dst_fs;
    int nplanes, ctas_per_frame, passes;
    uint32_t inv, inv2;
    SegPlane pl[3];
};

__device__ __forceinline__ const SegPlane& seg_plane(const SegJob& b, int cta, int& local) {
    int k = b.nplanes - 1;
    while (k > 0 && cta < b.pl[k].cta_begin) --k;
    local = cta - b.pl[k].cta_begin;
    return b.pl[k];
}

// --------------------------------------------------------------------------- TMA (bulk async copy) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
// L2 eviction priorities: rows that will be read again soon are kept (evict_last), data that is dead after this access
// leaves first (evict_first)
__device__ __forceinline__ uint64_t l2_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void st_global_hint(void* g, const uint4& v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(g), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

#ifndef VSZ_SEG_ALL_DP
#define VSZ_SEG_ALL_DP false
#endif
constexpr bool kAllDp = VSZ_SEG_ALL_DP;
#ifndef VSZ_SEG_HCHAINS
#define VSZ_SEG_HCHAINS 2
#endif
constexpr int kHChains = VSZ_SEG_HCHAINS;  // independent running sums per thread in the H pass

// One warp-synchronous H pass over the row(s) of a lane group.  `row` = the group's staged row (sample 0 at row + PAD);
// e[] = the lane's registers with the pass input (own part), o[] receives the results (own part).
//   first_pass: the input is in the staged row (put there by the TMA engine or by the comptime V phase), else in e[] and is
//   published to the row here.  Mirrors: when the row ends exactly at a segment end the first and last lane write them from
//   registers; ragged rows copy them inside the staged row and the last lane reloads its span (which contains mirrored samples).
template <int R>
__device__ __forceinline__ void hseg_pass(uint32_t (&e)[HGeom<R>::NW], uint32_t (&o)[HGeom<R>::NW], uint16_t* row, int n, int sg, int G, int S,
                                          bool first_pass, bool act, int lane, uint32_t inv, uint32_t inv2) {
    using Gm = HGeom<R>;
    uint16_t* own = row + Gm::PAD + L * sg;
    const int lig = lane & (G - 1);
    if (first_pass) h_load_own<R>(e, own);
    else if (act) h_store_own<R>(e, own);
    if (L * S == n) {
        if (lig == 0) h_write_left_pad<R>(e, row + Gm::PAD);
        if (lig == S - 1) h_write_right_pad<R>(e, row + Gm::PAD + n);
        __syncwarp();
    } else {
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 3; ++t) {  // r <= 22 < 3 * 8 lanes
            const int k = lig + t * G;
            if (k < R) {
                row[Gm::PAD - 1 - k] = row[Gm::PAD + k];
                row[Gm::PAD + n + k] = row[Gm::PAD + n - 1 - k];
            }
        }
        __syncwarp();
        if (lig == S - 1) h_load_own<R>(e, own);
    }
    h_load_halos<R>(e, own);
    const uint32_t W = h_window<R>(e);
    uint32_t C = line_const(W, inv, inv2);
    C = __shfl_sync(0xffffffffu, C, lane & ~(G - 1));  // the group's first lane holds the window at position 0
    h_slide<R, kHChains, kAllDp>(e, o, W, C, inv2);
    __syncwarp();  // every lane has read its halos: the staged row may be overwritten
}

// --------------------------------------------------------------------------- host side
constexpr int kMaxSmem = 227 * 1024;

void axis_consts(int r, uint32_t& inv, uint32_t& inv2) {
    const uint64_t v = ((1ull << 32) + (uint64_t)r) / (uint64_t)(2 * r + 1);
    inv = (uint32_t)v;
    inv2 = (uint32_t)(v >> 16);
}

SegJob base_job(const FrameLayout& l, const bool mask[3], const char* src, size_t sfs, char* dst, size_t dfs, int r, int passes) {
    SegJob j{};
    j.src = src; j.dst = dst; j.src_fs = sfs; j.dst_fs = dfs; j.passes = passes;
    axis_consts(r, j.inv, j.inv2);
    int k = 0;
    for (int p = 0; p < l.nplanes; ++p) {
        if (!mask[p]) continue;
        SegPlane& s = j.pl[k++];
        s.src_off = s.dst_off = l.pl[p].offset;
        s.src_pitch = s.dst_pitch = l.pl[p].pitch;
        s.w = l.pl[p].w; s.h = l.pl[p].h;
    }
    j.nplanes = k;
    return j;
}

int lanes_per_row(int S) { return S > 16 ? 32 : (S > 8 ? 16 : 8); }

// bytes between staged rows: a multiple of 16, and for 8-lane groups 16 words off a multiple of 32 words so that the two
// groups of a half-warp fall on disjoint banks
int rowbuf_bytes(int samples, int G) {
    int words = (samples + 1) / 2;
    words = (words + 3) & ~3;
    if (G == 8) while ((words & 31) != 16) words += 4;
    return words * 4;
}

template <class K>
int launch_frames(K kern, SegJob job, int count, int threads, size_t smem, cudaStream_t st) {
    VSZ_CUDA(allow_max_dynamic_smem(kern));
    for (int f0 = 0; f0 < count; f0 += 65535) {
        const int nf = std::min(65535, count - f0);
        SegJob j = job;
        j.src += (size_t)f0 * job.src_fs; j.dst += (size_t)f0 * job.dst_fs;
        kern<<<dim3(nf, job.ctas_per_frame), threads, smem, st>>>(j);  // frame index on grid.x (see boxblur_kernels.cu, launch_v)
        count_launch();
    }
    VSZ_CUDA(cudaGetLastError());
    return 0;
}

// Planes of different shapes want different CTA shapes (V: warps = segments per column; comptime: columns per thread), so a
// job is launched once per group of equally shaped planes.
template <class F>
int for_each_shape(const SegJob& job, F launch) {
    bool done[3] = {false, false, false};
    for (int k = 0; k < job.nplanes; ++k) {
        if (done[k]) continue;
        SegJob sub = job;
        sub.nplanes = 0;
        for (int q = k; q < job.nplanes; ++q)
            if (!done[q] && job.pl[q].w == job.pl[k].w && job.pl[q].h == job.pl[k].h) { sub.pl[sub.nplanes++] = job.pl[q]; done[q] = true; }
        const int rc = launch(sub);
        if (rc) return rc;
    }
    return 0;
}

#ifdef VSZ_SEG_DEV13  // A/B builds (scripts/build_variant.sh): one radius, seconds to compile
#define VSZ_SEG_RADII(X) X(13)
#else
#define VSZ_SEG_RADII(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20) X(21) X(22)
#endif

}  // namespace

}  // namespace vsz
