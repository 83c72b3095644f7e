// pbfic_kernels.cu — sm_100a kernels for vszip.Bilateral algorithm 1 (PBFIC, "Real-Time O(1) Bilateral Filtering").
//
// Semantics restated from the reference (src/filters/bilateral.zig:91-171, 336-431): for each of `num` range levels
// pk[k] build Wk = gr[|pk - ref|] and Jk = Wk*src, smooth both with a 3-tap recursive Gaussian (rows forward then
// backward, then columns forward then backward, each tap ((b*x + b1*p1) + b2*p2) + b3*p3 in f32 with separate
// multiplies and adds), take Jk/Wk as the level image and interpolate linearly between the two levels that
// bracket ref.  Every line runs the reference's own operation sequence, so results are bit-identical.
//
// Design: levels are processed one after the other so only four f32 images per plane live in HBM (W, J and a
// ping-pong pair of level images), whatever `num` is; a batch is cut into chunks of frames so that scratch stays
// bounded.  Per level and chunk two launches:
//   pbfic_h_kernel  one warp owns 16 rows and walks them in 32-column tiles: the tile of src/ref is loaded
//                   coalesced, turned into W/J on the fly (range LUT gathered from L2), transposed through shared
//                   memory (odd pitch: conflict-free both ways) so that lane = (image, row) for the recursion, and
//                   written back coalesced; then the same walk right-to-left over the forward result.
//   pbfic_v_kernel  one thread owns one column (coalesced by construction), forward then backward; the backward
//                   sweep emits the level image and, for samples bracketed by (level-1, level), the output.
// Bound: HBM/L2 traffic of the f32 intermediates (~60 B per sample and level) and the dependent 7-flop tap chains.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "filter.h"

namespace vsz {

struct PbficJob {
    const char* src; const char* ref; char* dst;
    size_t src_fs, ref_fs, dst_fs;   // frame strides (bytes); plane offsets already applied to the bases
    int src_pitch, ref_pitch, dst_pitch;
    int w, h;
    float* W; float* J; float* Lprev; float* Lcur;
    size_t img_fs;  // floats per frame in each scratch image
    int fpitch;     // floats per scratch row
    const float* gr;
    unsigned int gr_top;       // hist_len - 1 (a sample above the clip's peak must not index past the LUT)
    float b, b1, b2, b3;
    int level, num;
    float pk_f;                // this level as f32 (exactly representable in T)
    float lo_f;                // level-1 as f32
    float first_f, last2_f;    // pk[0], pk[num-2]: the union of the intervals before the last one
    float peak;
};

template <typename T> struct PTr { static constexpr bool flt = false; };
template <> struct PTr<__half> { static constexpr bool flt = true; };
template <> struct PTr<float> { static constexpr bool flt = true; };

template <typename T> __device__ __forceinline__ float to_f32(T v) { return (float)v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// rangeIndex(T, pk, ref) (src/filters/bilateral.zig:15-22): integers |a-b|; floats trunc(min(1,|a-b|)*65535 + 0.5)
// with the subtraction rounded in T.  a and b are T values widened exactly to f32.
template <typename T> __device__ __forceinline__ unsigned int range_index(float a, float b) {
    float d = __fsub_rn(a, b);
    if constexpr (PTr<T>::flt) {
        if constexpr (sizeof(T) == 2) d = __half2float(__float2half_rn(d));
        const float m = fminf(1.0f, fabsf(d));
        return (unsigned int)truncf(__fadd_rn(__fmul_rn(m, 65535.0f), 0.5f));
    } else {
        return (unsigned int)fabsf(d);
    }
}

// p0 = ((b*x + b1*p1) + b2*p2) + b3*p3
__device__ __forceinline__ float tap(const PbficJob& j, float x, float p1, float p2, float p3) {
    float acc = __fmul_rn(j.b, x);
    acc = __fadd_rn(acc, __fmul_rn(j.b1, p1));
    acc = __fadd_rn(acc, __fmul_rn(j.b2, p2));
    acc = __fadd_rn(acc, __fmul_rn(j.b3, p3));
    return acc;
}

// --------------------------------------------------------------------------- rows (build + forward + backward)
// One warp owns HR = 16 rows: lanes 0-15 run the W recursions of those rows, lanes 16-31 the J recursions, so a warp
// carries 32 independent chains while the grid has twice the warps of a 32-row layout (the kernel is latency-bound).
// The global loads of the next tile are issued before the current tile's recursion and parked in registers.
template <typename T>
__global__ void __launch_bounds__(32) pbfic_h_kernel(const PbficJob j) {
    constexpr int HR = 16, TP = HR + 1;
    __shared__ float tile[2 * 32 * TP + 16];
    float* tw = tile;
    float* tj = tile + 32 * TP + 16;  // +16 floats: the two half-warps then read disjoint banks
    const int lane = threadIdx.x, half = lane >> 4, r = lane & 15;
    float* mine = half ? tj : tw;
    const int row0 = blockIdx.x * HR, frame = blockIdx.y;
    const int nrows = min(HR, j.h - row0);
    const char* src = j.src + (size_t)frame * j.src_fs + (size_t)row0 * j.src_pitch;
    const char* ref = j.ref + (size_t)frame * j.ref_fs + (size_t)row0 * j.ref_pitch;
    float* W = j.W + (size_t)frame * j.img_fs + (size_t)row0 * j.fpitch;
    float* J = j.J + (size_t)frame * j.img_fs + (size_t)row0 * j.fpitch;
    const int nch = (j.w + 31) / 32;
    float p1 = 0.f, p2 = 0.f, p3 = 0.f;
    float pa[HR], pb[HR];  // prefetched tile: (src, ref) going forward, (W, J) going backward

    auto fetch_fwd = [&](int c) {
        const int x = c * 32 + lane;
        if (c < nch && x < j.w) {
#pragma unroll
            for (int rr = 0; rr < HR; ++rr) {
                if (rr < nrows) {
                    pa[rr] = to_f32<T>(reinterpret_cast<const T*>(src + (size_t)rr * j.src_pitch)[x]);
                    pb[rr] = to_f32<T>(reinterpret_cast<const T*>(ref + (size_t)rr * j.ref_pitch)[x]);
                }
            }
        }
    };
    auto fetch_bwd = [&](int c) {
        const int x = c * 32 + lane;
        if (c >= 0 && x < j.w) {
#pragma unroll
            for (int rr = 0; rr < HR; ++rr) {
                if (rr < nrows) { pa[rr] = W[(size_t)rr * j.fpitch + x]; pb[rr] = J[(size_t)rr * j.fpitch + x]; }
            }
        }
    };
    auto store_tile = [&](int x0, int ncols) {
        if (lane < ncols) {
#pragma unroll
            for (int rr = 0; rr < HR; ++rr) {
                if (rr < nrows) {
                    W[(size_t)rr * j.fpitch + x0 + lane] = tw[lane * TP + rr];
                    J[(size_t)rr * j.fpitch + x0 + lane] = tj[lane * TP + rr];
                }
            }
        }
    };

    // ---- forward, W/J built on the fly (src/filters/bilateral.zig:131-140, 396-414)
    fetch_fwd(0);
    for (int c = 0; c < nch; ++c) {
        const int x0 = c * 32, ncols = min(32, j.w - x0);
        if (lane < ncols) {
#pragma unroll
            for (int rr = 0; rr < HR; ++rr) {
                if (rr < nrows) {
                    const float wv = __ldg(j.gr + min(range_index<T>(j.pk_f, pb[rr]), j.gr_top));
                    tw[lane * TP + rr] = wv;
                    tj[lane * TP + rr] = __fmul_rn(wv, pa[rr]);
                }
            }
        }
        fetch_fwd(c + 1);
        __syncwarp();
        if (r < nrows) {
#pragma unroll 4
            for (int i = 0; i < ncols; ++i) {
                const float x = mine[i * TP + r];
                if (x0 + i == 0) {  // the first sample passes through and seeds the history
                    p1 = p2 = p3 = x;
                } else {
                    const float p0 = tap(j, x, p1, p2, p3);
                    p3 = p2; p2 = p1; p1 = p0;
                    mine[i * TP + r] = p0;
                }
            }
        }
        __syncwarp();
        store_tile(x0, ncols);
        __syncwarp();
    }
    // ---- backward over the forward result (:416-430); every lane re-reads only what it stored itself
    fetch_bwd(nch - 1);
    for (int c = nch - 1; c >= 0; --c) {
        const int x0 = c * 32, ncols = min(32, j.w - x0);
        if (lane < ncols) {
#pragma unroll
            for (int rr = 0; rr < HR; ++rr) {
                if (rr < nrows) { tw[lane * TP + rr] = pa[rr]; tj[lane * TP + rr] = pb[rr]; }
            }
        }
        fetch_bwd(c - 1);
        __syncwarp();
        if (r < nrows) {
#pragma unroll 4
            for (int i = ncols - 1; i >= 0; --i) {
                const float x = mine[i * TP + r];
                if (x0 + i == j.w - 1) {  // the last sample keeps its forward value and seeds the history
                    p1 = p2 = p3 = x;
                } else {
                    const float p0 = tap(j, x, p1, p2, p3);
                    p3 = p2; p2 = p1; p1 = p0;
                    mine[i * TP + r] = p0;
                }
            }
        }
        __syncwarp();
        store_tile(x0, ncols);
        __syncwarp();
    }
}

// --------------------------------------------------------------------------- columns + level image + output
template <typename T> __device__ __forceinline__ T finalize(float vf, float peak) {
    if constexpr (std::is_same<T, float>::value) return vf;
    else if constexpr (std::is_same<T, __half>::value) return __float2half_rn(vf);
    else {
        const float c = fminf(fmaxf(__fadd_rn(vf, 0.5f), 0.0f), peak);  // a NaN becomes 0 like the reference's clamp + cast
        return (T)truncf(c);
    }
}

template <typename T>
__global__ void __launch_bounds__(128) pbfic_v_kernel(const PbficJob j) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= j.w) return;
    const int frame = blockIdx.y, fp = j.fpitch, h = j.h;
    float* W = j.W + (size_t)frame * j.img_fs + x;
    float* J = j.J + (size_t)frame * j.img_fs + x;
    const float* Lp = j.Lprev + (size_t)frame * j.img_fs + x;
    float* Lc = j.Lcur + (size_t)frame * j.img_fs + x;
    const char* ref = j.ref + (size_t)frame * j.ref_fs + (size_t)x * sizeof(T);
    char* dst = j.dst + (size_t)frame * j.dst_fs + (size_t)x * sizeof(T);

    // ---- forward (src/filters/bilateral.zig:355-373): row 0 is filtered against itself
    float w1, w2, w3, j1, j2, j3;
    {
        const float xw = W[0], xj = J[0];
        const float pw = tap(j, xw, xw, xw, xw), pj = tap(j, xj, xj, xj, xj);
        W[0] = pw; J[0] = pj;
        w1 = w2 = w3 = pw; j1 = j2 = j3 = pj;
    }
    constexpr int U = 8;
    int y = 1;
    for (; y + U <= h; y += U) {
        float xw[U], xj[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { xw[u] = W[(size_t)(y + u) * fp]; xj[u] = J[(size_t)(y + u) * fp]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float pw = tap(j, xw[u], w1, w2, w3), pj = tap(j, xj[u], j1, j2, j3);
            w3 = w2; w2 = w1; w1 = pw; j3 = j2; j2 = j1; j1 = pj;
            W[(size_t)(y + u) * fp] = pw; J[(size_t)(y + u) * fp] = pj;
        }
    }
    for (; y < h; ++y) {
        const float pw = tap(j, W[(size_t)y * fp], w1, w2, w3), pj = tap(j, J[(size_t)y * fp], j1, j2, j3);
        w3 = w2; w2 = w1; w1 = pw; j3 = j2; j2 = j1; j1 = pj;
        W[(size_t)y * fp] = pw; J[(size_t)y * fp] = pj;
    }

    // ---- backward (:375-393) fused with the level image (:146-151) and the interpolation (:154-169)
    const bool last_level = (j.level == j.num - 1), interp = (j.level >= 1);
    // rf / lo are fetched by the caller together with the row's W/J so that no load sits inside the dependent chain
    auto emit = [&](int yy, float wv, float jv, float rf, float lo) {
        const float L = (wv == 0.0f) ? 0.0f : __fdiv_rn(jv, wv);
        if (!last_level) Lc[(size_t)yy * fp] = L;
        if (interp) {
            // bracket index k = level-1: the first k < num-2 with pk[k] <= ref < pk[k+1], else num-2
            bool mine = (rf >= j.lo_f) && (rf < j.pk_f);
            if (last_level) mine = !((rf >= j.first_f) && (rf < j.last2_f));
            if (mine) {
                const float t0 = __fmul_rn(__fsub_rn(j.pk_f, rf), lo), t1 = __fmul_rn(__fsub_rn(rf, j.lo_f), L);
                const float vf = __fdiv_rn(__fadd_rn(t0, t1), __fsub_rn(j.pk_f, j.lo_f));
                *reinterpret_cast<T*>(dst + (size_t)yy * j.dst_pitch) = finalize<T>(vf, j.peak);
            }
        }
    };
    auto ld_ref = [&](int yy) { return interp ? to_f32<T>(*reinterpret_cast<const T*>(ref + (size_t)yy * j.ref_pitch)) : 0.0f; };
    auto ld_lo = [&](int yy) { return interp ? Lp[(size_t)yy * fp] : 0.0f; };
    {
        const float xw = w1, xj = j1;  // forward value of row h-1 (still in registers), filtered against itself
        const float pw = tap(j, xw, xw, xw, xw), pj = tap(j, xj, xj, xj, xj);
        w1 = w2 = w3 = pw; j1 = j2 = j3 = pj;
        emit(h - 1, pw, pj, ld_ref(h - 1), ld_lo(h - 1));
    }
    y = h - 2;
    for (; y - (U - 1) >= 0; y -= U) {
        float xw[U], xj[U], rf[U], lo[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            xw[u] = W[(size_t)(y - u) * fp]; xj[u] = J[(size_t)(y - u) * fp];
            rf[u] = ld_ref(y - u); lo[u] = ld_lo(y - u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float pw = tap(j, xw[u], w1, w2, w3), pj = tap(j, xj[u], j1, j2, j3);
            w3 = w2; w2 = w1; w1 = pw; j3 = j2; j2 = j1; j1 = pj;
            emit(y - u, pw, pj, rf[u], lo[u]);
        }
    }
    for (; y >= 0; --y) {
        const float pw = tap(j, W[(size_t)y * fp], w1, w2, w3), pj = tap(j, J[(size_t)y * fp], j1, j2, j3);
        w3 = w2; w2 = w1; w1 = pw; j3 = j2; j2 = j1; j1 = pj;
        emit(y, pw, pj, ld_ref(y), ld_lo(y));
    }
}

// =========================================================================== host launcher
// src/filters/bilateral.zig:336-348 (f64 math, results rounded to f32)
static void recursive_gaussian_params(double sigma, float* b, float* b1, float* b2, float* b3) {
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

// Level values (src/filters/bilateral.zig:98-109) as f32 numbers that are exactly representable in T.
static std::vector<float> pbfic_levels(SampleKind kind, int num, float peak) {
    std::vector<float> pk((size_t)num);
    for (int k = 0; k < num; ++k) {
        if (kind == K_F32) {
            pk[k] = (float)k / (float)(num - 1);
        } else if (kind == K_F16) {
            // the division happens in f16; computing it in f32 and rounding once more is exact (24 >= 2*11 + 2)
            pk[k] = __half2float(__float2half_rn((float)k / (float)(num - 1)));
        } else {
            float v = peak * (float)k;
            v = v / ((float)num - 1.0f);
            v = v + 0.5f;
            const float hi = kind == K_U8 ? 255.0f : 65535.0f;  // lossyCast: saturating truncation
            pk[k] = std::isnan(v) ? 0.0f : v >= hi ? hi : v <= 0.0f ? 0.0f : std::trunc(v);
        }
    }
    return pk;
}

size_t pbfic_scratch_bytes_per_frame(int w, int h) {
    const size_t fpitch = ((size_t)w + 31) / 32 * 32;
    return 4 * fpitch * (size_t)h * sizeof(float);
}

template <typename T>
static int launch_pbfic_t(PbficJob j, int count, const std::vector<float>& pk, cudaStream_t st) {
    const size_t per_frame = 4 * j.img_fs * sizeof(float);
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)count, ((size_t)1 << 30) / per_frame));
    AsyncScratch scratch_mem;
    VSZ_CUDA(scratch_mem.alloc(per_frame * (size_t)chunk, st));
    float* scratch = reinterpret_cast<float*>(scratch_mem.p);
    const size_t img_all = j.img_fs * (size_t)chunk;
    const char* src0 = j.src; const char* ref0 = j.ref; char* dst0 = j.dst;
    for (int f0 = 0; f0 < count; f0 += chunk) {
        const int nf = std::min(chunk, count - f0);
        j.src = src0 + (size_t)f0 * j.src_fs; j.ref = ref0 + (size_t)f0 * j.ref_fs; j.dst = dst0 + (size_t)f0 * j.dst_fs;
        j.W = scratch; j.J = scratch + img_all;
        float* L[2] = {scratch + 2 * img_all, scratch + 3 * img_all};
        for (int k = 0; k < j.num; ++k) {
            j.level = k;
            j.pk_f = pk[k];
            j.lo_f = k ? pk[k - 1] : 0.0f;
            j.Lcur = L[k & 1]; j.Lprev = L[(k & 1) ^ 1];
            pbfic_h_kernel<T><<<dim3((j.h + 15) / 16, nf), 32, 0, st>>>(j);
            pbfic_v_kernel<T><<<dim3((j.w + 127) / 128, nf), 128, 0, st>>>(j);
            count_launch(2);
        }
    }
    VSZ_CUDA(cudaGetLastError());
    return 0;
}

int run_pbfic(const FrameLayout& l, int plane, const char* src, size_t src_fs, const char* ref, size_t ref_fs, char* dst, size_t dst_fs,
              int count, const float* gr_dev, int hist_len, double sigmaS, int num, float peak, cudaStream_t st) {
    if (count <= 0) return 0;
    if (num < 2) { set_error("Bilateral: PBFICnum must be at least 2"); return -2; }
    const PlaneGeom& g = l.pl[plane];
    PbficJob j{};
    j.src = src + g.offset; j.ref = (ref ? ref : src) + g.offset; j.dst = dst + g.offset;
    j.src_fs = src_fs; j.ref_fs = ref ? ref_fs : src_fs; j.dst_fs = dst_fs;
    j.src_pitch = j.ref_pitch = j.dst_pitch = g.pitch;
    j.w = g.w; j.h = g.h;
    j.fpitch = (g.w + 31) / 32 * 32;
    j.img_fs = (size_t)j.fpitch * g.h;
    j.gr = gr_dev; j.gr_top = (unsigned)hist_len - 1u;
    recursive_gaussian_params(sigmaS, &j.b, &j.b1, &j.b2, &j.b3);
    j.num = num; j.peak = peak;
    const std::vector<float> pk = pbfic_levels(l.kind, num, peak);
    j.first_f = pk[0]; j.last2_f = pk[(size_t)num - 2];
    switch (l.kind) {
        case K_U8: return launch_pbfic_t<uint8_t>(j, count, pk, st);
        case K_U16: return launch_pbfic_t<uint16_t>(j, count, pk, st);
        case K_F16: return launch_pbfic_t<__half>(j, count, pk, st);
        case K_F32: return launch_pbfic_t<float>(j, count, pk, st);
    }
    return -1;
}

}  // namespace vsz
