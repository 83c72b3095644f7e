// fixture.cpp — rebuilds the reference test suite's input clips without
// VapourSynth/zimg.  TEST INFRASTRUCTURE ONLY (see vszip_oracle.cpp header).
//
// The reference's fixtures (tests/conftest.py:72-121) are
//   ImageRead(tests/image.png) -> Crop to 640x320 RGB24
//   -> resize.Bilinear(format=..., matrix=1)            (zimg)
// SURVEY.md §4 restates zimg's arithmetic for this input exactly: f32 with
// single-rounded fma, BT.709 coefficients rounded to f32, limited-range
// quantisation with round-half-even, and a (1/8,3/8,3/8,1/8) vertical /
// (1/4,1/2,1/4) left-sited horizontal bilinear pair for 4:2:0 chroma.
// The restatement is validated by the golden-key tests: a fixture error would
// break every key.

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace {
typedef _Float16 f16;

struct Coef { float y[3], u[3], v[3]; };

Coef bt709() {
    const double kr = 0.2126, kb = 0.0722, kg = 1.0 - kr - kb;
    Coef c;
    c.y[0] = (float)kr; c.y[1] = (float)kg; c.y[2] = (float)kb;
    c.u[0] = (float)(-kr / (2.0 - 2.0 * kb)); c.u[1] = (float)(-kg / (2.0 - 2.0 * kb)); c.u[2] = 0.5f;
    c.v[0] = 0.5f; c.v[1] = (float)(-kg / (2.0 - 2.0 * kr)); c.v[2] = (float)(-kb / (2.0 - 2.0 * kr));
    return c;
}

inline float mix(const float* k, float r, float g, float b) {
    return fmaf(b, k[2], fmaf(g, k[1], k[0] * r));
}

inline int mirror_edge(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }
}  // namespace

extern "C" {

// rgb8: 3 planes of w*h bytes (R, G, B), tightly packed.  rgbs: 3 planes f32.
void vsf_rgb8_to_rgbs(const uint8_t* rgb8, size_t count, float* rgbs) {
    const float k = (float)(1.0 / 255.0);
    for (size_t i = 0; i < count; ++i) rgbs[i] = (float)rgb8[i] * k;
}

// plane: 0 = Y, 1 = U, 2 = V (full resolution, f32), from planar RGBS.
void vsf_rgbs_to_yuv_plane(const float* r, const float* g, const float* b, size_t count, int plane, float* out) {
    const Coef c = bt709();
    const float* k = plane == 0 ? c.y : (plane == 1 ? c.u : c.v);
    for (size_t i = 0; i < count; ++i) out[i] = mix(k, r[i], g[i], b[i]);
}

// 4:2:0 chroma down-sampling of a full-res f32 plane (w, h even): vertical
// first, then horizontal.  out is (w/2) x (h/2).
void vsf_chroma_420(const float* in, int w, int h, float* out) {
    const int h2 = h / 2, w2 = w / 2;
    std::vector<float> vv((size_t)w * h2);
    for (int i = 0; i < h2; ++i) {
        const float* a = in + (size_t)mirror_edge(2 * i - 1, h) * w;
        const float* b = in + (size_t)mirror_edge(2 * i, h) * w;
        const float* c = in + (size_t)mirror_edge(2 * i + 1, h) * w;
        const float* d = in + (size_t)mirror_edge(2 * i + 2, h) * w;
        float* o = vv.data() + (size_t)i * w;
        for (int x = 0; x < w; ++x) {
            const float odd = fmaf(c[x], 0.375f, 0.125f * a[x]);
            const float even = fmaf(d[x], 0.125f, 0.375f * b[x]);
            o[x] = odd + even;
        }
    }
    for (int i = 0; i < h2; ++i) {
        const float* row = vv.data() + (size_t)i * w;
        float* o = out + (size_t)i * w2;
        for (int j = 0; j < w2; ++j) {
            const float l = row[mirror_edge(2 * j - 1, w)];
            const float m = row[mirror_edge(2 * j, w)];
            const float r = row[mirror_edge(2 * j + 1, w)];
            o[j] = fmaf(r, 0.25f, 0.25f * l) + 0.5f * m;
        }
    }
}

// limited-range quantisation; chroma != 0 selects the 224/128 scale.
// out16 receives the value for any depth in 8..16; caller narrows for 8-bit.
void vsf_quantise(const float* in, size_t count, int bits, int chroma, uint16_t* out16) {
    const float scale = (float)((chroma ? 224 : 219) << (bits - 8));
    const float offset = (float)((chroma ? 128 : 16) << (bits - 8));
    const float hi = (float)((1 << bits) - 1);
    for (size_t i = 0; i < count; ++i) {
        float v = nearbyintf(fmaf(in[i], scale, offset));
        v = v < 0.f ? 0.f : (v > hi ? hi : v);
        out16[i] = (uint16_t)v;
    }
}

void vsf_f32_to_f16(const float* in, size_t count, uint16_t* out_bits) {
    for (size_t i = 0; i < count; ++i) {
        f16 hv = (f16)in[i];
        uint16_t u;
        __builtin_memcpy(&u, &hv, 2);
        out_bits[i] = u;
    }
}

}  // extern "C"
