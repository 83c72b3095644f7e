// vszip_oracle.cpp — CPU restatement of the vszip 19.0.0 hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  This file is the parity oracle for the
// CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` leg may load it.  The product (vapoursynth_zip_b200/) never
// links, imports or calls anything under oracle/.
//
// Parity status: PINNED.  The reference (Zig + VapourSynth) cannot be built in
// this image (no zig, no vapoursynth, no network), so the oracle is pinned by
// the reference's own golden vectors: tests/test_oracle_goldens.py rebuilds the
// reference's fixture clips (oracle/fixture.cpp) and checks the values below
// against tests/golden/*.json, which are verbatim excerpts of
// /root/reference/tests/goldens/{boxblur,bilateral,planeminmax,planeaverage}.json.
//
// Written from the arithmetic description in SURVEY.md §8a / Appendix A, not by
// transliterating the Zig sources.  Each function cites the reference file:line
// whose behaviour it restates.  All float arithmetic is IEEE f32 with separate
// multiply and add (build with -ffp-contract=off; never -ffast-math).

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <type_traits>
#include <vector>

namespace {

typedef _Float16 f16;

enum SampleType { ST_U8 = 0, ST_U16 = 1, ST_F16 = 2, ST_F32 = 3, ST_U32 = 4 };  // ST_U32: Limiter only

template <class T> struct is_flt : std::integral_constant<bool, std::is_same<T, f16>::value || std::is_same<T, float>::value> {};

template <class T> inline const T* row_ptr(const void* base, ptrdiff_t stride_bytes, int y) {
    return reinterpret_cast<const T*>(static_cast<const char*>(base) + stride_bytes * y);
}
template <class T> inline T* row_ptr(void* base, ptrdiff_t stride_bytes, int y) {
    return reinterpret_cast<T*>(static_cast<char*>(base) + stride_bytes * y);
}

// ---------------------------------------------------------------------------
// Index maps
// ---------------------------------------------------------------------------

// SYM: edge-repeating mirror.  Used by every runtime-path blur and the comptime
// integer horizontal pass (src/filters/boxblur_runtime.zig:24-40,
// src/filters/boxblur_comptime.zig:141-158).
inline int sym_index(int p, int n) {
    if (p < 0) return -p - 1;
    if (p >= n) return 2 * n - 1 - p;
    return p;
}

// R101q: window tap k (0..2r) around position i.  Reflect-101 at the low edge,
// "reflect about the current position" at the high edge
// (src/filters/boxblur_comptime.zig:50-70 and :201-262).
inline int r101q_index(int i, int k, int r, int n) {
    if (k < r) {
        const int need = r - k;  // distance of this tap below i
        return (i < need) ? std::min(need - i, n - 1) : i - need;
    }
    const int over = k - r;       // distance of this tap above i
    const int room = n - 1 - i;   // samples available above i
    return (room < over) ? i - std::min(over - room, i) : i + over;
}

// ---------------------------------------------------------------------------
// BoxBlur, runtime path: one line, one pass
// ---------------------------------------------------------------------------

// Integer line pass in closed form (equivalent to the running sum of
// src/filters/boxblur_runtime.zig:10-41): with W_x the SYM-mirrored window sum,
// out[x] = (S0 + inv2*(W_x - W_0)) >> 16, S0 = (W_0*inv + 2^31) >> 16.
template <class T>
void rt_line_int(const T* s, ptrdiff_t ss, T* d, ptrdiff_t ds, int n, int r) {
    const uint64_t k = 2ull * (uint64_t)r + 1;
    const uint64_t inv = ((1ull << 32) + (uint64_t)r) / k;
    const uint64_t inv2 = inv >> 16;
    std::vector<uint64_t> pre((size_t)n + 2 * (size_t)r + 1);
    pre[0] = 0;
    for (int i = 0; i < n + 2 * r; ++i) pre[i + 1] = pre[i] + (uint64_t)s[(ptrdiff_t)sym_index(i - r, n) * ss];
    const uint64_t w0 = pre[2 * r + 1] - pre[0];
    const uint64_t s0 = (w0 * inv + (1ull << 31)) >> 16;
    for (int x = 0; x < n; ++x) {
        const uint64_t wx = pre[x + 2 * r + 1] - pre[x];
        const uint64_t acc = s0 + inv2 * wx - inv2 * w0;  // modular u64; true value is non-negative
        d[(ptrdiff_t)x * ds] = (T)(acc >> 16);
    }
}

// Float line pass: strictly sequential f32 recurrence
// (src/filters/boxblur_runtime.zig:43-79).
template <class T>
void rt_line_float(const T* s, ptrdiff_t ss, T* d, ptrdiff_t ds, int n, int r) {
    const float div = 1.0f / (float)(2 * r + 1);
    float sum = (float)s[(ptrdiff_t)r * ss];
    for (int x = 0; x < r; ++x) {
        const float v = (float)s[(ptrdiff_t)x * ss];
        sum = sum + v * 2.0f;
    }
    sum = sum * div;
    for (int x = 0; x < n; ++x) {
        int ia, ib;
        if (x <= r) { ia = r + x; ib = r - x; }
        else if (x < n - r) { ia = r + x; ib = x - r - 1; }
        else { ia = 2 * n - r - x - 1; ib = x - r - 1; }
        const float a = (float)s[(ptrdiff_t)ia * ss];
        const float b = (float)s[(ptrdiff_t)ib * ss];
        const float delta = a - b;
        sum = sum + delta * div;
        d[(ptrdiff_t)x * ds] = (T)sum;
    }
}

template <class T>
void rt_line(const T* s, ptrdiff_t ss, T* d, ptrdiff_t ds, int n, int r) {
    if (is_flt<T>::value) rt_line_float<T>(s, ss, d, ds, n, r);
    else rt_line_int<T>(s, ss, d, ds, n, r);
}

// `passes` applications along one line; every pass re-quantises to T
// (src/filters/boxblur_runtime.zig:81-119).
template <class T>
void rt_line_passes(const T* s, ptrdiff_t ss, T* d, ptrdiff_t ds, int n, int r, int passes) {
    std::vector<T> a((size_t)n), b((size_t)n);
    for (int x = 0; x < n; ++x) a[x] = s[(ptrdiff_t)x * ss];
    for (int p = 0; p < passes; ++p) {
        rt_line<T>(a.data(), 1, b.data(), 1, n, r);
        a.swap(b);
    }
    for (int x = 0; x < n; ++x) d[(ptrdiff_t)x * ds] = a[x];
}

// Runtime path for one plane: all H passes, then all V passes, whatever route
// the reference's getFrame takes (src/vapoursynth/boxblur.zig:93-112; the fused
// and swept variants are documented bit-identical, boxblur_runtime.zig:147-152,
// :276-282).
template <class T>
void boxblur_rt_plane(const void* src, ptrdiff_t sst, void* dst, ptrdiff_t dstt, int w, int h,
                      int hr, int hp, int vr, int vp) {
    const bool hb = hr > 0 && hp > 0;
    const bool vb = vr > 0 && vp > 0;
    std::vector<T> mid((size_t)w * (size_t)h);
    for (int y = 0; y < h; ++y) {
        const T* sp = row_ptr<T>(src, sst, y);
        T* mp = mid.data() + (size_t)y * w;
        if (hb) rt_line_passes<T>(sp, 1, mp, 1, w, hr, hp);
        else std::memcpy(mp, sp, sizeof(T) * (size_t)w);
    }
    if (vb) {
        std::vector<T> col_out((size_t)h);
        for (int x = 0; x < w; ++x) {
            rt_line_passes<T>(mid.data() + x, w, col_out.data(), 1, h, vr, vp);
            for (int y = 0; y < h; ++y) row_ptr<T>(dst, dstt, y)[x] = col_out[y];
        }
    } else {
        for (int y = 0; y < h; ++y) std::memcpy(row_ptr<T>(dst, dstt, y), mid.data() + (size_t)y * w, sizeof(T) * (size_t)w);
    }
}

// ---------------------------------------------------------------------------
// BoxBlur, comptime path (hradius == vradius in 1..22, single pass): V then H
// ---------------------------------------------------------------------------

// Integer: exact R101q column sums -> rounded mean stored as T -> SYM closed
// form along the row (src/filters/boxblur_comptime.zig:10-38, :111-159).
template <class T>
void boxblur_ct_plane_int(const void* src, ptrdiff_t sst, void* dst, ptrdiff_t dstt, int w, int h, int r) {
    const uint64_t k = 2ull * (uint64_t)r + 1;
    const uint64_t inv = ((1ull << 32) + (uint64_t)r) / k;
    std::vector<T> tmp((size_t)w);
    std::vector<uint32_t> col((size_t)w);
    for (int i = 0; i < h; ++i) {
        std::fill(col.begin(), col.end(), 0u);
        for (int t = 0; t <= 2 * r; ++t) {
            const T* sp = row_ptr<T>(src, sst, r101q_index(i, t, r, h));
            for (int j = 0; j < w; ++j) col[j] += (uint32_t)sp[j];
        }
        for (int j = 0; j < w; ++j) tmp[j] = (T)(((uint64_t)col[j] * inv + (1ull << 31)) >> 32);
        rt_line_int<T>(tmp.data(), 1, row_ptr<T>(dst, dstt, i), 1, w, r);
    }
}

// Float: direct (non-running) sums in tap order, acc = acc + div*v, R101q in
// both directions, narrowed to T after each direction
// (src/filters/boxblur_comptime.zig:39-44, :161-263).
template <class T>
void boxblur_ct_plane_float(const void* src, ptrdiff_t sst, void* dst, ptrdiff_t dstt, int w, int h, int r) {
    const float div = 1.0f / (float)(2 * r + 1);
    std::vector<T> tmp((size_t)w);
    for (int i = 0; i < h; ++i) {
        for (int j = 0; j < w; ++j) {
            float acc = 0.0f;
            for (int t = 0; t <= 2 * r; ++t) {
                const float v = (float)row_ptr<T>(src, sst, r101q_index(i, t, r, h))[j];
                const float m = div * v;
                acc = acc + m;
            }
            tmp[j] = (T)acc;
        }
        T* dp = row_ptr<T>(dst, dstt, i);
        for (int j = 0; j < w; ++j) {
            float acc = 0.0f;
            for (int t = 0; t <= 2 * r; ++t) {
                const float v = (float)tmp[r101q_index(j, t, r, w)];
                const float m = div * v;
                acc = acc + m;
            }
            dp[j] = (T)acc;
        }
    }
}

template <class T>
void boxblur_plane_t(const void* src, ptrdiff_t sst, void* dst, ptrdiff_t dstt, int w, int h,
                     int hr, int hp, int vr, int vp) {
    // dispatch rule: src/vapoursynth/boxblur.zig:188
    const bool use_rt = (hr != vr) || (hr > 22) || (hp > 1) || (vp > 1);
    if (use_rt) boxblur_rt_plane<T>(src, sst, dst, dstt, w, h, hr, hp, vr, vp);
    else if (is_flt<T>::value) boxblur_ct_plane_float<T>(src, sst, dst, dstt, w, h, hr);
    else boxblur_ct_plane_int<T>(src, sst, dst, dstt, w, h, hr);
}

// ---------------------------------------------------------------------------
// Bilateral, algorithm 2
// ---------------------------------------------------------------------------

// src/filters/bilateral.zig:306-314
void spatial_lut(std::vector<float>& gs, int upper, double sigma_s) {
    gs.resize((size_t)upper * upper);
    for (int y = 0; y < upper; ++y)
        for (int x = 0; x < upper; ++x)
            gs[(size_t)y * upper + x] = (float)std::exp((double)(x * x + y * y) / (sigma_s * sigma_s * -2.0));
}

// src/filters/bilateral.zig:316-334
void range_lut(std::vector<float>& gr, int len, double range, double sigma_r) {
    gr.resize((size_t)len);
    const double lim = std::min(range, sigma_r * 8.0 * range + 0.5);
    const uint32_t upper = (uint32_t)std::trunc(lim);
    const double norm = std::sqrt(2.0 * M_PI) * sigma_r;
    uint32_t i = 0;
    for (; i <= upper && i < (uint32_t)len; ++i) {
        const double y = (double)i / range;
        const double x = y / sigma_r;
        gr[i] = (float)(std::exp(x * x / -2.0) / norm);
    }
    const float tail = gr[std::min<uint32_t>(upper, (uint32_t)len - 1)];
    for (; i < (uint32_t)len; ++i) gr[i] = tail;
}

template <class T> inline uint32_t range_index(T a, T b) {
    if (is_flt<T>::value) {
        // difference rounded in T, then widened (src/filters/bilateral.zig:15-22)
        const T dt = (T)(a - b);
        float ad = std::fabs((float)dt);
        const float m = std::min(1.0f, ad);
        const float scaled = m * 65535.0f;
        const float biased = scaled + 0.5f;
        return (uint32_t)std::trunc(biased);
    }
    return a > b ? (uint32_t)(a - b) : (uint32_t)(b - a);
}

template <class T> inline T bilateral_finalize(float sum, float wsum, float peak) {
    const float q = sum / wsum;
    if (is_flt<T>::value) return (T)q;
    const float b = q + 0.5f;
    const float c = std::min(std::max(b, 0.0f), peak);
    return (T)std::trunc(c);
}

// One plane of the "truncated" bilateral with replicate edges
// (src/filters/bilateral.zig:178-304).  Interior pixels never clamp, so a
// single clamped formulation covers the SIMD interior, the scalar remainder
// and the four edge bands.
template <class T>
void bilateral_plane_t(const void* src, ptrdiff_t sst, const void* ref, ptrdiff_t rst, void* dst, ptrdiff_t dstt,
                       int w, int h, const float* gs, const float* gr, int radius, int step, float peak) {
    const int r2 = radius + 1;
    const float w0 = gs[0] * gr[0];
    for (int y = 0; y < h; ++y) {
        T* dp = row_ptr<T>(dst, dstt, y);
        const T* sc = row_ptr<T>(src, sst, y);
        const T* rc = row_ptr<T>(ref, rst, y);
        for (int x = 0; x < w; ++x) {
            const T cx = rc[x];
            float wsum = w0;
            float sum = (float)sc[x] * wsum;
            for (int yy = 1; yy < r2; yy += step) {
                const int ya = std::max(y - yy, 0), yb = std::min(y + yy, h - 1);
                const T* sa = row_ptr<T>(src, sst, ya);
                const T* sb = row_ptr<T>(src, sst, yb);
                const T* ra = row_ptr<T>(ref, rst, ya);
                const T* rb = row_ptr<T>(ref, rst, yb);
                for (int xx = 1; xx < r2; xx += step) {
                    const int xp = std::min(x + xx, w - 1), xm = std::max(x - xx, 0);
                    const float sw = gs[yy * r2 + xx];
                    const float g1 = gr[range_index<T>(cx, ra[xp])];
                    const float g2 = gr[range_index<T>(cx, rb[xp])];
                    const float g3 = gr[range_index<T>(cx, ra[xm])];
                    const float g4 = gr[range_index<T>(cx, rb[xm])];
                    float gsum = g1 + g2; gsum = gsum + g3; gsum = gsum + g4;
                    const float wterm = sw * gsum;
                    wsum = wsum + wterm;
                    const float p1 = (float)sa[xp] * g1, p2 = (float)sb[xp] * g2;
                    const float p3 = (float)sa[xm] * g3, p4 = (float)sb[xm] * g4;
                    float psum = p1 + p2; psum = psum + p3; psum = psum + p4;
                    const float sterm = sw * psum;
                    sum = sum + sterm;
                }
            }
            dp[x] = bilateral_finalize<T>(sum, wsum, peak);
        }
    }
}

// ---------------------------------------------------------------------------
// Bilateral, algorithm 1 (PBFIC, "Real-Time O(1) Bilateral Filtering")
// ---------------------------------------------------------------------------

// src/filters/bilateral.zig:336-348 (f64 -> f32)
void recursive_gaussian_params(double sigma, float* b, float* b1, float* b2, float* b3) {
    const double q = sigma < 2.5 ? (3.97156 - 4.14554 * std::sqrt(1 - 0.26891 * sigma)) : 0.98711 * sigma - 0.96330;
    const double den = 1.57825 + 2.44413 * q + 1.4281 * q * q + 0.422205 * q * q * q;
    const double n1 = 2.44413 * q + 2.85619 * q * q + 1.26661 * q * q * q;
    const double n2 = -(1.4281 * q * q + 1.26661 * q * q * q);
    const double n3 = 0.422205 * q * q * q;
    *b = (float)(1 - (n1 + n2 + n3) / den);
    *b1 = (float)(n1 / den);
    *b2 = (float)(n2 / den);
    *b3 = (float)(n3 / den);
}

// p0 = ((b*x + b1*p1) + b2*p2) + b3*p3, separate multiplies and adds (strict float mode)
inline float iir_tap(float b, float x, float b1, float p1, float b2, float p2, float b3, float p3) {
    float acc = b * x;
    acc = acc + b1 * p1;
    acc = acc + b2 * p2;
    acc = acc + b3 * p3;
    return acc;
}

// In-place horizontal pass, src/filters/bilateral.zig:396-431: forward with the history initialised to the first
// sample, backward with the history initialised to the (forward-filtered) last sample, which itself is kept.
void recursive_gaussian_h(float* img, int w, int h, float b, float b1, float b2, float b3) {
    for (int j = 0; j < h; ++j) {
        float* row = img + (size_t)j * w;
        float p1 = row[0], p2 = p1, p3 = p1;
        for (int i = 1; i < w; ++i) {
            const float p0 = iir_tap(b, row[i], b1, p1, b2, p2, b3, p3);
            p3 = p2; p2 = p1; p1 = p0;
            row[i] = p0;
        }
        if (w == 1) continue;
        p1 = row[w - 1]; p2 = p1; p3 = p1;
        for (int i = w - 2; i >= 0; --i) {
            const float p0 = iir_tap(b, row[i], b1, p1, b2, p2, b3, p3);
            p3 = p2; p2 = p1; p1 = p0;
            row[i] = p0;
        }
    }
}

// In-place vertical pass, src/filters/bilateral.zig:350-394: the history rows are clamped row indices, so row 0
// (forward) and row h-1 (backward) are filtered against themselves (the right-hand side is read before the store).
void recursive_gaussian_v(float* img, int w, int h, float b, float b1, float b2, float b3) {
    for (int j = 0; j < h; ++j) {
        float* x0 = img + (size_t)j * w;
        const float* x1 = j < 1 ? x0 : x0 - w;
        const float* x2 = j < 2 ? x1 : x1 - w;
        const float* x3 = j < 3 ? x2 : x2 - w;
        for (int i = 0; i < w; ++i) x0[i] = iir_tap(b, x0[i], b1, x1[i], b2, x2[i], b3, x3[i]);
    }
    for (int j = h - 1; j >= 0; --j) {
        float* x0 = img + (size_t)j * w;
        const float* x1 = j >= h - 1 ? x0 : x0 + w;
        const float* x2 = j >= h - 2 ? x1 : x1 + w;
        const float* x3 = j >= h - 3 ? x2 : x2 + w;
        for (int i = 0; i < w; ++i) x0[i] = iir_tap(b, x0[i], b1, x1[i], b2, x2[i], b3, x3[i]);
    }
}

// src/filters/bilateral.zig:91-171
template <class T>
void pbfic_plane_t(const void* src, ptrdiff_t sst, const void* ref, ptrdiff_t rst, void* dst, ptrdiff_t dstt, int w, int h,
                   const float* gr, double sigma_s, int num, float peak) {
    std::vector<T> pk((size_t)num);
    if (is_flt<T>::value) {
        const T denom = (T)(float)(num - 1);
        for (int k = 0; k < num; ++k) pk[k] = (T)((T)(float)k / denom);  // the division is done in T
    } else {
        const float numf = (float)num;
        for (int k = 0; k < num; ++k) {
            float v = peak * (float)k;
            v = v / (numf - 1.0f);
            v = v + 0.5f;
            const float hi = (float)std::numeric_limits<T>::max();  // lossyCast: saturating truncation
            pk[k] = std::isnan(v) ? (T)0 : v >= hi ? std::numeric_limits<T>::max() : v <= 0.0f ? (T)0 : (T)v;
        }
    }
    float b, b1, b2, b3;
    recursive_gaussian_params(sigma_s, &b, &b1, &b2, &b3);
    const size_t n = (size_t)w * h;
    std::vector<float> levels((size_t)num * n), wk(n), jk(n);
    for (int k = 0; k < num; ++k) {
        for (int y = 0; y < h; ++y) {
            const T* sp = row_ptr<T>(src, sst, y);
            const T* rp = row_ptr<T>(ref, rst, y);
            for (int x = 0; x < w; ++x) {
                const float wv = gr[range_index<T>(pk[k], rp[x])];
                wk[(size_t)y * w + x] = wv;
                jk[(size_t)y * w + x] = wv * (float)sp[x];
            }
        }
        recursive_gaussian_h(wk.data(), w, h, b, b1, b2, b3);
        recursive_gaussian_v(wk.data(), w, h, b, b1, b2, b3);
        recursive_gaussian_h(jk.data(), w, h, b, b1, b2, b3);
        recursive_gaussian_v(jk.data(), w, h, b, b1, b2, b3);
        float* lv = levels.data() + (size_t)k * n;
        for (size_t i = 0; i < n; ++i) lv[i] = (wk[i] == 0.0f) ? 0.0f : jk[i] / wk[i];
    }
    for (int y = 0; y < h; ++y) {
        const T* rp = row_ptr<T>(ref, rst, y);
        T* dp = row_ptr<T>(dst, dstt, y);
        for (int x = 0; x < w; ++x) {
            int k = 0;
            while (k < num - 2) {
                if (rp[x] < pk[k + 1] && rp[x] >= pk[k]) break;
                ++k;
            }
            const float rf = (float)rp[x], p0f = (float)pk[k], p1f = (float)pk[k + 1];
            const size_t i = (size_t)y * w + x;
            const float lo = levels[(size_t)k * n + i], hi = levels[(size_t)(k + 1) * n + i];
            const float t0 = (p1f - rf) * lo, t1 = (rf - p0f) * hi;
            const float vf = (t0 + t1) / (p1f - p0f);
            dp[x] = bilateral_finalize<T>(vf, 1.0f, peak);
        }
    }
}

// ---------------------------------------------------------------------------
// Limiter (src/vapoursynth/limiter.zig:24-93): dst = min(max(lo, src), hi), bounds in the sample type
// ---------------------------------------------------------------------------
template <class T>
void limiter_plane_t(const void* src, ptrdiff_t sst, void* dst, ptrdiff_t dstt, int w, int h, double lo, double hi) {
    const T l = is_flt<T>::value ? (T)(float)lo : (T)(uint32_t)lo, u = is_flt<T>::value ? (T)(float)hi : (T)(uint32_t)hi;
    for (int y = 0; y < h; ++y) {
        const T* sp = row_ptr<T>(src, sst, y);
        T* dp = row_ptr<T>(dst, dstt, y);
        for (int x = 0; x < w; ++x) {
            if (is_flt<T>::value) dp[x] = (T)std::fmin(std::fmax((float)l, (float)sp[x]), (float)u);  // @max/@min: NaN yields the other operand
            else dp[x] = std::min(std::max(l, sp[x]), u);
        }
    }
}

// ---------------------------------------------------------------------------
// LimitFilter (src/filters/limit_filter.zig:3-34): soft limit of flt towards src, judged on |flt - ref| in f32
// ---------------------------------------------------------------------------
template <class T>
void limitfilter_plane_t(const void* flt, ptrdiff_t fst, const void* src, ptrdiff_t sst, const void* ref, ptrdiff_t rst, void* dst,
                         ptrdiff_t dstt, int w, int h, float dark_thr, float bright_thr, float elast) {
    for (int y = 0; y < h; ++y) {
        const T* fp = row_ptr<T>(flt, fst, y);
        const T* sp = row_ptr<T>(src, sst, y);
        const T* rp = row_ptr<T>(ref, rst, y);
        T* dp = row_ptr<T>(dst, dstt, y);
        for (int x = 0; x < w; ++x) {
            const float ff = (float)fp[x], sf = (float)sp[x], rf = (float)rp[x];
            const float d = ff - rf;
            const float ad = std::fabs(d);
            const float t1 = d > 0.0f ? bright_thr : dark_thr;
            const float t2 = t1 * elast;
            float o;
            if (ad <= t1) o = ff;                 // inside the threshold: keep the filtered sample
            else if (ad >= t2) o = sf;            // beyond thr * elast: back to the source
            else {
                const float num = (ff - sf) * (t2 - ad);
                o = sf + num / (t2 - t1);
            }
            if (is_flt<T>::value) dp[x] = (T)o;
            else dp[x] = (T)(int32_t)std::trunc(o + 0.5f);
        }
    }
}

// ---------------------------------------------------------------------------
// AdaptiveBinarize (src/vapoursynth/adaptive_binarize.zig:48-60): 255 where clip2 - clip >= c, else 0 (8-bit only)
// ---------------------------------------------------------------------------
void adaptive_binarize_plane(const uint8_t* a, ptrdiff_t ast, const uint8_t* b, ptrdiff_t bst, uint8_t* dst, ptrdiff_t dstt, int w, int h, int c) {
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) dst[dstt * y + x] = ((int)b[bst * y + x] - (int)a[ast * y + x] >= c) ? 255 : 0;
}

// ---------------------------------------------------------------------------
// PlaneMinMax / PlaneAverage
// ---------------------------------------------------------------------------

inline uint32_t sat_u16_from_float(float f) {  // std.math.lossyCast(u16, f)
    if (std::isnan(f)) return 0;
    if (f >= 65535.0f) return 65535;
    if (f <= 0.0f) return 0;
    return (uint32_t)f;  // truncation
}

template <class T> inline uint32_t hist_bin(T v) {
    if (is_flt<T>::value) {
        const float f = (float)v;
        const float s = f * 65535.0f;
        const float b = s + 0.5f;
        return sat_u16_from_float(b);
    }
    return (uint32_t)v;
}

template <class T> inline double abs_diff_f64(T a, T b) {
    if (is_flt<T>::value) {
        const T d = (T)(a - b);  // rounded in T (src/filters/planeminmax.zig:28)
        return (double)std::fabs((float)d);
    }
    return std::fabs((double)a - (double)b);
}

struct MinMaxOut { int64_t imin, imax; double fmin, fmax, diff; };

// src/filters/planeminmax.zig:11-70 (threshold scan) and :80-133 (no threshold)
template <class T>
void planeminmax_t(const void* src, ptrdiff_t sst, const void* ref, ptrdiff_t rst, int w, int h, int bits,
                   float minthr, float maxthr, MinMaxOut* out) {
    const bool flt = is_flt<T>::value;
    const uint32_t hist_size = flt ? 65536u : (1u << bits);
    const uint32_t peak = hist_size - 1;
    const double total = (double)((uint32_t)w * (uint32_t)h);
    double diffacc = 0.0;
    const bool no_thr = (minthr == 0.0f) && (maxthr == 0.0f);  // src/vapoursynth/planeminmax.zig:131
    if (no_thr) {
        float fmn = std::numeric_limits<float>::infinity(), fmx = -std::numeric_limits<float>::infinity();
        int64_t imn = flt ? 0 : (int64_t)std::numeric_limits<T>::max(), imx = 0;
        for (int y = 0; y < h; ++y) {
            const T* sp = row_ptr<T>(src, sst, y);
            const T* rp = ref ? row_ptr<T>(ref, rst, y) : nullptr;
            for (int x = 0; x < w; ++x) {
                if (rp) diffacc += abs_diff_f64<T>(sp[x], rp[x]);
                if (flt) { fmn = std::min(fmn, (float)sp[x]); fmx = std::max(fmx, (float)sp[x]); }
                else { imn = std::min<int64_t>(imn, (int64_t)sp[x]); imx = std::max<int64_t>(imx, (int64_t)sp[x]); }
            }
        }
        out->imin = imn; out->imax = imx; out->fmin = fmn; out->fmax = fmx;
    } else {
        std::vector<uint32_t> hist(65536, 0u);
        for (int y = 0; y < h; ++y) {
            const T* sp = row_ptr<T>(src, sst, y);
            const T* rp = ref ? row_ptr<T>(ref, rst, y) : nullptr;
            for (int x = 0; x < w; ++x) {
                hist[hist_bin<T>(sp[x])] += 1;
                if (rp) diffacc += abs_diff_f64<T>(sp[x], rp[x]);
            }
        }
        const uint32_t tmin = (uint32_t)std::trunc(total * (double)minthr);
        const uint32_t tmax = (uint32_t)std::trunc(total * (double)maxthr);
        uint32_t lo = peak, hi = 0, count = 0;
        for (uint32_t u = 0; u < hist_size; ++u) {
            count += hist[u];
            if (count > tmin) { lo = u; break; }
        }
        count = 0;
        for (int64_t i = (int64_t)peak; i >= 0; --i) {
            count += hist[(size_t)i];
            if (count > tmax) { hi = (uint32_t)i; break; }
        }
        out->imin = lo; out->imax = hi;
        out->fmin = (double)((float)lo / 65535.0f);
        out->fmax = (double)((float)hi / 65535.0f);
    }
    if (ref) {
        double d = diffacc / total;
        if (!flt) d = d / (double)(float)peak;
        out->diff = d;
    } else {
        out->diff = 0.0;
    }
}

struct AverageOut { double avg, diff; };

// src/filters/planeaverage.zig:16-84; peak = f32(2^bits - 1)
// (src/vapoursynth/planeaverage.zig:112).
template <class T>
void planeaverage_t(const void* src, ptrdiff_t sst, const void* ref, ptrdiff_t rst, int w, int h, int bits,
                    const int32_t* exclude, int nex, AverageOut* out) {
    const bool flt = is_flt<T>::value;
    const float peak = (float)((1ull << bits) - 1ull);
    const uint32_t all = (uint32_t)w * (uint32_t)h;
    uint32_t total = all;
    uint64_t iacc = 0, idiff = 0;
    double facc = 0.0, fdiff = 0.0;
    std::vector<float> exf((size_t)nex);
    for (int i = 0; i < nex; ++i) exf[i] = (float)exclude[i];
    for (int y = 0; y < h; ++y) {
        const T* sp = row_ptr<T>(src, sst, y);
        const T* rp = ref ? row_ptr<T>(ref, rst, y) : nullptr;
        for (int x = 0; x < w; ++x) {
            bool found = false;
            for (int i = 0; i < nex && !found; ++i) {
                if (flt) found = ((float)sp[x] == exf[i]);
                else found = ((int64_t)sp[x] == (int64_t)exclude[i]);
            }
            if (found) total -= 1;
            else if (flt) facc += (double)(float)sp[x];
            else iacc += (uint64_t)sp[x];
            if (rp) {
                if (flt) { const T d = sp[x] > rp[x] ? (T)(sp[x] - rp[x]) : (T)(rp[x] - sp[x]); fdiff += (double)(float)d; }
                else idiff += (uint64_t)(sp[x] > rp[x] ? sp[x] - rp[x] : rp[x] - sp[x]);
            }
        }
    }
    const double totalf = (double)total;
    if (total == 0) out->avg = 0.0;
    else if (flt) out->avg = facc / totalf;
    else out->avg = (double)iacc / totalf / (double)peak;
    if (ref) out->diff = flt ? fdiff / (double)all : (double)idiff / (double)all / (double)peak;
    else out->diff = 0.0;
}

}  // namespace

// ===========================================================================
// C entry points (ctypes-friendly)
// ===========================================================================

extern "C" {

// Per-plane BoxBlur with the reference's CT/RT dispatch.  Returns 0 on success.
int vso_boxblur_plane(int st, const void* src, ptrdiff_t sstride, void* dst, ptrdiff_t dstride, int w, int h,
                      int hradius, int hpasses, int vradius, int vpasses) {
    switch (st) {
        case ST_U8: boxblur_plane_t<uint8_t>(src, sstride, dst, dstride, w, h, hradius, hpasses, vradius, vpasses); return 0;
        case ST_U16: boxblur_plane_t<uint16_t>(src, sstride, dst, dstride, w, h, hradius, hpasses, vradius, vpasses); return 0;
        case ST_F16: boxblur_plane_t<f16>(src, sstride, dst, dstride, w, h, hradius, hpasses, vradius, vpasses); return 0;
        case ST_F32: boxblur_plane_t<float>(src, sstride, dst, dstride, w, h, hradius, hpasses, vradius, vpasses); return 0;
    }
    return -1;
}

// Bilateral parameter derivation (src/vapoursynth/bilateral.zig:104-199).
// n_* = number of user-supplied elements (0 = not given); planes_n < 0 = not given.
// Returns 0, or a negative code for the validation errors of that function.
struct vso_bilateral_params {
    double sigmaS[3], sigmaR[3];
    int process[3], algorithm[3];
    unsigned pbfic_num[3], radius[3], samples[3], step[3];
    float peak;
    int hist_len;
};

int vso_bilateral_derive(int is_yuv, int sample_is_float, int bits, int ssw, int ssh, int num_planes,
                         const double* sigmaS, int n_s, const double* sigmaR, int n_r,
                         const int* planes, int planes_n, const int* algorithm, int n_a,
                         const int* pbfic, int n_p, vso_bilateral_params* o) {
    o->hist_len = sample_is_float ? 65536 : (1 << bits);
    o->peak = (float)(o->hist_len - 1);
    for (int i = 0; i < 3; ++i) {
        if (i < n_s) o->sigmaS[i] = sigmaS[i];
        else if (i == 0) o->sigmaS[0] = 3.0;
        else if (i == 1 && is_yuv && ssh != 0 && ssw != 0)
            o->sigmaS[1] = o->sigmaS[0] / std::sqrt((double)((1u << ssh) * (1u << ssw)));
        else o->sigmaS[i] = o->sigmaS[i - 1];
        if (o->sigmaS[i] < 0) return -1;
    }
    for (int i = 0; i < 3; ++i) {
        o->sigmaR[i] = i < n_r ? sigmaR[i] : (i == 0 ? 0.02 : o->sigmaR[i - 1]);
        if (o->sigmaR[i] < 0) return -2;
        o->algorithm[i] = i < n_a ? algorithm[i] : (i == 0 ? 0 : o->algorithm[i - 1]);
        if (o->algorithm[i] < 0 || o->algorithm[i] > 2) return -3;
        const int pn = i < n_p ? pbfic[i] : (i == 0 ? 0 : (int)o->pbfic_num[i - 1]);
        if (pn < 0 || pn > 256) return -4;
        o->pbfic_num[i] = (unsigned)pn;
    }
    for (int i = 0; i < 3; ++i) o->process[i] = planes_n < 0 ? 1 : 0;
    for (int i = 0; i < planes_n; ++i) {
        if (planes[i] < 0 || planes[i] >= num_planes) return -5;
        if (o->process[planes[i]]) return -6;
        o->process[planes[i]] = 1;
    }
    for (int i = 0; i < 3; ++i)
        if (o->sigmaS[i] == 0 || o->sigmaR[i] == 0) o->process[i] = 0;
    for (int i = 0; i < 3; ++i)
        if (o->pbfic_num[i] == 1) return -7;
    for (int i = 0; i < 3; ++i) {
        if (o->process[i] && o->pbfic_num[i] == 0) {
            if (o->sigmaR[i] >= 0.08) o->pbfic_num[i] = 4;
            else if (o->sigmaR[i] >= 0.015) o->pbfic_num[i] = std::min(16u, (unsigned)std::trunc(4 * 0.08 / o->sigmaR[i] + 0.5));
            else o->pbfic_num[i] = std::min(32u, (unsigned)std::trunc(16 * 0.015 / o->sigmaR[i] + 0.5));
            if (i > 0 && is_yuv && (o->pbfic_num[i] % 2 == 0) && o->pbfic_num[i] < 256) o->pbfic_num[i] += 1;
        }
    }
    for (int i = 0; i < 3; ++i) {
        o->radius[i] = o->samples[i] = o->step[i] = 0;
        if (!o->process[i]) continue;
        const int orad = std::max((int)std::trunc(o->sigmaS[i] * 2 + 0.5), 1);
        o->step[i] = orad < 4 ? 1 : (orad < 8 ? 2 : 3);
        o->samples[i] = 1;
        o->radius[i] = 1 + (o->samples[i] - 1) * o->step[i];
        while (orad * 2 > (int)o->radius[i] * 3) {
            o->samples[i] += 1;
            o->radius[i] = 1 + (o->samples[i] - 1) * o->step[i];
            if ((int)o->radius[i] >= orad && o->samples[i] > 2) {
                o->samples[i] -= 1;
                o->radius[i] = 1 + (o->samples[i] - 1) * o->step[i];
                break;
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        if (o->process[i] && o->algorithm[i] <= 0) {
            if (o->step[i] == 1) o->algorithm[i] = 2;
            else if (o->sigmaR[i] < 0.08 && o->samples[i] < 5) o->algorithm[i] = 2;
            else if (4 * o->samples[i] * o->samples[i] <= 15 * o->pbfic_num[i]) o->algorithm[i] = 2;
            else o->algorithm[i] = 1;
        }
    }
    return 0;
}

// Fills gs (length (radius+1)^2) and gr (length hist_len).
void vso_bilateral_luts(double sigmaS, double sigmaR, int radius, int hist_len, float* gs, float* gr) {
    std::vector<float> a, b;
    spatial_lut(a, radius + 1, sigmaS);
    range_lut(b, hist_len, (double)(float)(hist_len - 1), sigmaR);
    std::memcpy(gs, a.data(), a.size() * sizeof(float));
    std::memcpy(gr, b.data(), b.size() * sizeof(float));
}

// Algorithm-2 bilateral on one plane.  `ref` may equal `src` (non-joint).
int vso_bilateral_plane(int st, const void* src, ptrdiff_t sstride, const void* ref, ptrdiff_t rstride,
                        void* dst, ptrdiff_t dstride, int w, int h, double sigmaS, double sigmaR,
                        int radius, int step, int hist_len) {
    std::vector<float> gs, gr;
    spatial_lut(gs, radius + 1, sigmaS);
    const float peak = (float)(hist_len - 1);
    range_lut(gr, hist_len, (double)peak, sigmaR);
    switch (st) {
        case ST_U8: bilateral_plane_t<uint8_t>(src, sstride, ref, rstride, dst, dstride, w, h, gs.data(), gr.data(), radius, step, peak); return 0;
        case ST_U16: bilateral_plane_t<uint16_t>(src, sstride, ref, rstride, dst, dstride, w, h, gs.data(), gr.data(), radius, step, peak); return 0;
        case ST_F16: bilateral_plane_t<f16>(src, sstride, ref, rstride, dst, dstride, w, h, gs.data(), gr.data(), radius, step, peak); return 0;
        case ST_F32: bilateral_plane_t<float>(src, sstride, ref, rstride, dst, dstride, w, h, gs.data(), gr.data(), radius, step, peak); return 0;
    }
    return -1;
}

// Algorithm-1 (PBFIC) bilateral on one plane.  `ref` may equal `src` (non-joint).
int vso_bilateral_pbfic_plane(int st, const void* src, ptrdiff_t sstride, const void* ref, ptrdiff_t rstride,
                              void* dst, ptrdiff_t dstride, int w, int h, double sigmaS, double sigmaR,
                              int pbfic_num, int hist_len) {
    if (pbfic_num < 2) return -2;
    std::vector<float> gr;
    const float peak = (float)(hist_len - 1);
    range_lut(gr, hist_len, (double)peak, sigmaR);
    switch (st) {
        case ST_U8: pbfic_plane_t<uint8_t>(src, sstride, ref, rstride, dst, dstride, w, h, gr.data(), sigmaS, pbfic_num, peak); return 0;
        case ST_U16: pbfic_plane_t<uint16_t>(src, sstride, ref, rstride, dst, dstride, w, h, gr.data(), sigmaS, pbfic_num, peak); return 0;
        case ST_F16: pbfic_plane_t<f16>(src, sstride, ref, rstride, dst, dstride, w, h, gr.data(), sigmaS, pbfic_num, peak); return 0;
        case ST_F32: pbfic_plane_t<float>(src, sstride, ref, rstride, dst, dstride, w, h, gr.data(), sigmaS, pbfic_num, peak); return 0;
    }
    return -1;
}

// src/filters/bilateral.zig:336-348, exposed for the host-logic tests
void vso_recursive_gaussian_params(double sigma, float* out4) { recursive_gaussian_params(sigma, out4, out4 + 1, out4 + 2, out4 + 3); }

int vso_limiter_plane(int st, const void* src, ptrdiff_t sstride, void* dst, ptrdiff_t dstride, int w, int h, double lo, double hi) {
    switch (st) {
        case ST_U8: limiter_plane_t<uint8_t>(src, sstride, dst, dstride, w, h, lo, hi); return 0;
        case ST_U16: limiter_plane_t<uint16_t>(src, sstride, dst, dstride, w, h, lo, hi); return 0;
        case ST_F16: limiter_plane_t<f16>(src, sstride, dst, dstride, w, h, lo, hi); return 0;
        case ST_F32: limiter_plane_t<float>(src, sstride, dst, dstride, w, h, lo, hi); return 0;
        case ST_U32: limiter_plane_t<uint32_t>(src, sstride, dst, dstride, w, h, lo, hi); return 0;
    }
    return -1;
}

int vso_limitfilter_plane(int st, const void* flt, ptrdiff_t fst, const void* src, ptrdiff_t sst, const void* ref, ptrdiff_t rst, void* dst,
                          ptrdiff_t dstt, int w, int h, float dark_thr, float bright_thr, float elast) {
    switch (st) {
        case ST_U8: limitfilter_plane_t<uint8_t>(flt, fst, src, sst, ref, rst, dst, dstt, w, h, dark_thr, bright_thr, elast); return 0;
        case ST_U16: limitfilter_plane_t<uint16_t>(flt, fst, src, sst, ref, rst, dst, dstt, w, h, dark_thr, bright_thr, elast); return 0;
        case ST_F16: limitfilter_plane_t<f16>(flt, fst, src, sst, ref, rst, dst, dstt, w, h, dark_thr, bright_thr, elast); return 0;
        case ST_F32: limitfilter_plane_t<float>(flt, fst, src, sst, ref, rst, dst, dstt, w, h, dark_thr, bright_thr, elast); return 0;
    }
    return -1;
}

int vso_adaptive_binarize_plane(const void* a, ptrdiff_t ast, const void* b, ptrdiff_t bst, void* dst, ptrdiff_t dstt, int w, int h, int c) {
    adaptive_binarize_plane((const uint8_t*)a, ast, (const uint8_t*)b, bst, (uint8_t*)dst, dstt, w, h, c);
    return 0;
}

struct vso_minmax_out { long long imin, imax; double fmin, fmax, diff; };

int vso_planeminmax_plane(int st, int bits, const void* src, ptrdiff_t sstride, const void* ref, ptrdiff_t rstride,
                          int w, int h, float minthr, float maxthr, vso_minmax_out* out) {
    MinMaxOut o{};
    switch (st) {
        case ST_U8: planeminmax_t<uint8_t>(src, sstride, ref, rstride, w, h, bits, minthr, maxthr, &o); break;
        case ST_U16: planeminmax_t<uint16_t>(src, sstride, ref, rstride, w, h, bits, minthr, maxthr, &o); break;
        case ST_F16: planeminmax_t<f16>(src, sstride, ref, rstride, w, h, bits, minthr, maxthr, &o); break;
        case ST_F32: planeminmax_t<float>(src, sstride, ref, rstride, w, h, bits, minthr, maxthr, &o); break;
        default: return -1;
    }
    out->imin = o.imin; out->imax = o.imax; out->fmin = o.fmin; out->fmax = o.fmax; out->diff = o.diff;
    return 0;
}

struct vso_average_out { double avg, diff; };

int vso_planeaverage_plane(int st, int bits, const void* src, ptrdiff_t sstride, const void* ref, ptrdiff_t rstride,
                           int w, int h, const int32_t* exclude, int nex, vso_average_out* out) {
    AverageOut o{};
    switch (st) {
        case ST_U8: planeaverage_t<uint8_t>(src, sstride, ref, rstride, w, h, bits, exclude, nex, &o); break;
        case ST_U16: planeaverage_t<uint16_t>(src, sstride, ref, rstride, w, h, bits, exclude, nex, &o); break;
        case ST_F16: planeaverage_t<f16>(src, sstride, ref, rstride, w, h, bits, exclude, nex, &o); break;
        case ST_F32: planeaverage_t<float>(src, sstride, ref, rstride, w, h, bits, exclude, nex, &o); break;
        default: return -1;
    }
    out->avg = o.avg; out->diff = o.diff;
    return 0;
}

// std.PlaneStats as used by the reference's golden store (tests/golden.py:106-121):
// avg normalised by peak for integer formats; f16 measured after exact widening.
void vso_plane_stats(int st, int bits, const void* src, ptrdiff_t sstride, int w, int h, double* avg, double* mn, double* mx) {
    double acc = 0.0, lo = std::numeric_limits<double>::infinity(), hi = -lo;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            double v = 0;
            switch (st) {
                case ST_U8: v = row_ptr<uint8_t>(src, sstride, y)[x]; break;
                case ST_U16: v = row_ptr<uint16_t>(src, sstride, y)[x]; break;
                case ST_F16: v = (double)(float)row_ptr<f16>(src, sstride, y)[x]; break;
                case ST_F32: v = (double)row_ptr<float>(src, sstride, y)[x]; break;
            }
            acc += v; lo = std::min(lo, v); hi = std::max(hi, v);
        }
    const double n = (double)w * (double)h;
    *avg = (st == ST_U8 || st == ST_U16) ? acc / n / (double)((1u << bits) - 1u) : acc / n;
    *mn = lo; *mx = hi;
}

}  // extern "C"
