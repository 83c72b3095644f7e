// CPU lane-by-lane simulation of the segment BoxBlur kernels (vapoursynth_zip_b200/csrc/boxblur_seg_kernels.cu).
// TEST INFRASTRUCTURE: it runs the product's own per-thread arithmetic (boxblur_seg_core.h) with the kernels'
// orchestration (publish, mirror pads, halo loads, slide) replayed sequentially, so the index logic can be checked
// against the oracle in the GPU-less container (tests/test_seg_sim.py).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "boxblur_seg_core.h"

using namespace vsz::seg;

namespace {

struct Axis { uint32_t inv, inv2; };
Axis axis(int r) {
    const uint64_t inv = ((1ull << 32) + (uint64_t)r) / (uint64_t)(2 * r + 1);
    return {(uint32_t)inv, (uint32_t)(inv >> 16)};
}

// ---- H: one row, `passes` passes, the way a group of lanes runs it
template <int R>
void h_row(const uint16_t* src, uint16_t* dst, int n, int passes, int chains) {
    using G = HGeom<R>;
    const Axis ax = axis(R);
    const int S = (n + L - 1) / L;
    std::vector<uint16_t> buf(G::row_samples(S) + 8, 0xdead);
    uint16_t* row = buf.data() + G::PAD;
    // the bulk copy brings round16(n*2) bytes; model the garbage tail
    memcpy(row, src, (size_t)n * 2);
    std::vector<std::vector<uint32_t>> ext(S, std::vector<uint32_t>(G::NW, 0xabababab));
    for (int p = 0; p < passes; ++p) {
        if (p > 0) for (int s = 0; s < S; ++s) h_store_own<R>(*reinterpret_cast<uint32_t(*)[G::NW]>(ext[s].data()), row + L * s);
        if (L * S == n) {  // pads from the registers of the first / last lane
            if (p == 0) { h_load_own<R>(*reinterpret_cast<uint32_t(*)[G::NW]>(ext[0].data()), row); h_load_own<R>(*reinterpret_cast<uint32_t(*)[G::NW]>(ext[S - 1].data()), row + L * (S - 1)); }
            h_write_left_pad<R>(*reinterpret_cast<uint32_t(*)[G::NW]>(ext[0].data()), row);
            h_write_right_pad<R>(*reinterpret_cast<uint32_t(*)[G::NW]>(ext[S - 1].data()), row + n);
        } else {
            for (int k = 0; k < R; ++k) { row[-1 - k] = row[k]; row[n + k] = row[n - 1 - k]; }
        }
        uint32_t C = 0;
        std::vector<uint32_t> W(S);
        for (int s = 0; s < S; ++s) {
            auto& e = *reinterpret_cast<uint32_t(*)[G::NW]>(ext[s].data());
            if (p == 0 || (s == S - 1 && L * S > n)) h_load_own<R>(e, row + L * s);
            h_load_halos<R>(e, row + L * s);
            W[s] = h_window<R>(e);
            if (s == 0) C = line_const(W[0], ax.inv, ax.inv2);
        }
        for (int s = 0; s < S; ++s) {
            auto& e = *reinterpret_cast<uint32_t(*)[G::NW]>(ext[s].data());
            uint32_t out[G::NW];
            if (chains == 1) h_slide<R, 1, false>(e, out, W[s], C, ax.inv2); else if (chains == 2) h_slide<R, 2, true>(e, out, W[s], C, ax.inv2); else h_slide<R, 3, true>(e, out, W[s], C, ax.inv2);
            for (int k = 0; k < LW; ++k) e[G::HW + k] = out[G::HW + k];
        }
    }
    for (int s = 0; s < S; ++s) h_store_own<R>(*reinterpret_cast<uint32_t(*)[G::NW]>(ext[s].data()), row + L * s);
    memcpy(dst, row, (size_t)n * 2);
}

// ---- V: one pair of columns (a 32-bit word per row), `passes` passes, the way the S threads of a column run it
template <int R, int L>
void v_colpair(const uint32_t* src, ptrdiff_t sstride_w, uint32_t* dst, ptrdiff_t dstride_w, int n, int passes, bool all_dp) {
    using G = VGeom<R, L>;
    const Axis ax = axis(R);
    const int S = (n + L - 1) / L;
    const bool exact = (n % L) == 0;
    std::vector<uint32_t> tile(R + L * S + R + 1, 0xdeadbeef);  // row y at tile[R + y]
    std::vector<std::vector<uint32_t>> ext(S, std::vector<uint32_t>(G::NW, 0xabababab));
    for (int s = 0; s < S; ++s)
        for (int i = 0; i < L; ++i) ext[s][R + i] = src[(ptrdiff_t)std::min(L * s + i, n - 1) * sstride_w];
    for (int p = 0; p < passes; ++p) {
        uint32_t Wl, Wh;
        v_window0<R, L>(*reinterpret_cast<uint32_t(*)[G::NW]>(ext[0].data()), Wl, Wh);
        const uint32_t Cl = line_const(Wl, ax.inv, ax.inv2), Ch = line_const(Wh, ax.inv, ax.inv2);
        for (int s = 0; s < S; ++s) {
            if (exact) {
                for (int j = 0; j < R; ++j) { tile[R + L * s + j] = ext[s][R + j]; tile[R + L * s + L - R + j] = ext[s][R + L - R + j]; }
            } else {
                for (int i = 0; i < L; ++i) tile[R + L * s + i] = ext[s][R + i];
            }
        }
        if (!exact) for (int k = 0; k < R; ++k) { tile[R - 1 - k] = tile[R + k]; tile[R + n + k] = tile[R + n - 1 - k]; }
        for (int s = 0; s < S; ++s) {
            auto& e = ext[s];
            if (exact && s == 0) { for (int k = 0; k < R; ++k) e[R - 1 - k] = e[R + k]; }
            else for (int j = 0; j < R; ++j) e[j] = tile[L * s + j];
            if (exact && s == S - 1) { for (int k = 0; k < R; ++k) e[R + L + k] = e[R + L - 1 - k]; }
            else for (int j = 0; j < R; ++j) e[R + L + j] = tile[R + L * s + L + j];
            if (!exact && s == S - 1) for (int i = 0; i < L; ++i) e[R + i] = tile[R + L * s + i];
        }
        for (int s = 0; s < S; ++s) {
            auto& e = *reinterpret_cast<uint32_t(*)[G::NW]>(ext[s].data());
            uint32_t out[L];
            v_window<R, L>(e, Wl, Wh);
            if (all_dp) v_slide<R, L, true>(e, out, Wl, Wh, Cl, Ch, ax.inv2); else v_slide<R, L, false>(e, out, Wl, Wh, Cl, Ch, ax.inv2);
            for (int i = 0; i < L; ++i) e[R + i] = out[i];
        }
    }
    for (int s = 0; s < S; ++s)
        for (int i = 0; i < L && L * s + i < n; ++i) dst[(ptrdiff_t)(L * s + i) * dstride_w] = ext[s][R + i];
}

// ---- comptime path: bands of rows, exact column sums stepped with ct_add_row / ct_sub_row, rounded means, one H pass
template <int R>
void ct_plane(const uint16_t* src, uint16_t* dst, int w, int h, int band_rows) {
    const Axis ax = axis(R);
    std::vector<uint32_t> col(w);
    std::vector<uint16_t> tmp(w);
    for (int y0 = 0; y0 < h; y0 += band_rows) {
        const int y1 = std::min(h, y0 + band_rows);
        std::fill(col.begin(), col.end(), 0u);
        for (int k = 0; k <= 2 * R; ++k) {
            const uint16_t* row = src + (size_t)ct_tap_row(y0, k, R, h) * w;
            for (int x = 0; x < w; ++x) col[x] += row[x];
        }
        for (int y = y0; y < y1; ++y) {
            for (int x = 0; x < w; ++x) tmp[x] = (uint16_t)ct_mean(col[x], ax.inv);
            h_row<R>(tmp.data(), dst + (size_t)y * w, w, 1, 2);
            const uint16_t* a = src + (size_t)std::min(ct_add_row(y, R, h), h - 1) * w;
            const uint16_t* b = src + (size_t)std::min(ct_sub_row(y, R), h - 1) * w;
            for (int x = 0; x < w; ++x) col[x] = col[x] + a[x] - b[x];
        }
    }
}

template <int R>
int run(int what, const uint16_t* src, uint16_t* dst, int w, int h, int passes, int opt) {
    if (what == 0) {
        for (int y = 0; y < h; ++y) h_row<R>(src + (size_t)y * w, dst + (size_t)y * w, w, passes, 1 + opt % 3);
    } else if (what == 1) {
        // columns are paired the way the kernel pairs them; an odd last column is paired with a zero column
        const int wp = (w + 1) / 2 * 2;
        std::vector<uint16_t> a((size_t)wp * h, 0), b((size_t)wp * h, 0);
        for (int y = 0; y < h; ++y) memcpy(&a[(size_t)y * wp], src + (size_t)y * w, (size_t)w * 2);
        for (int c = 0; c < wp / 2; ++c) {
            const uint32_t* ap = reinterpret_cast<const uint32_t*>(a.data()) + c;
            uint32_t* bp = reinterpret_cast<uint32_t*>(b.data()) + c;
            if (h % 90 == 0) v_colpair<R, 90>(ap, wp / 2, bp, wp / 2, h, passes, opt != 0);  // the kernel's choice of segment length
            else v_colpair<R, 60>(ap, wp / 2, bp, wp / 2, h, passes, opt != 0);
        }
        for (int y = 0; y < h; ++y) memcpy(dst + (size_t)y * w, &b[(size_t)y * wp], (size_t)w * 2);
    } else {
        ct_plane<R>(src, dst, w, h, opt > 0 ? opt : h);
    }
    return 0;
}

}  // namespace

// what: 0 = H passes, 1 = V passes, 2 = comptime path (opt = rows per band).  Returns -1 for an unsupported radius.
extern "C" int seg_sim_u16(int what, const uint16_t* src, uint16_t* dst, int w, int h, int r, int passes, int opt) {
    switch (r) {
#define CASE(R) case R: return run<R>(what, src, dst, w, h, passes, opt);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15)
        CASE(16) CASE(17) CASE(18) CASE(19) CASE(20) CASE(21) CASE(22)
#undef CASE
    }
    return -1;
}
