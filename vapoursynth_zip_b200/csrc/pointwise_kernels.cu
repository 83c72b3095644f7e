// pointwise_kernels.cu — sm_100a kernels for the pointwise neighbours of the hot path.
//
// vszip.Limiter (src/vapoursynth/limiter.zig:24-93): dst = min(max(lo, src), hi) per processed plane, bounds per plane
// in the sample type.  One read + one write per sample: HBM-bound by construction.  A CTA streams 4 rows of a plane
// with 16-byte vectors (4 independent loads in flight per thread), two 16-bit samples per VIMNMX.U16x2.
#include <cuda_fp16.h>

#include "filter.h"

namespace vsz {

struct LimiterParams {
    uint32_t lo_w[3], hi_w[3];  // bounds replicated into a 32-bit word of samples (u8 x4, u16 x2, f16 x2, f32 x1)
};

static constexpr int LNT = 256, LROWS = 4, LGROUPS = 4;  // a CTA streams LGROUPS groups of LROWS rows

template <typename T> __device__ __forceinline__ uint32_t clamp_word(uint32_t x, uint32_t lo, uint32_t hi);
template <> __device__ __forceinline__ uint32_t clamp_word<uint8_t>(uint32_t x, uint32_t lo, uint32_t hi) { return __vminu4(__vmaxu4(lo, x), hi); }
template <> __device__ __forceinline__ uint32_t clamp_word<uint16_t>(uint32_t x, uint32_t lo, uint32_t hi) { return __vminu2(__vmaxu2(lo, x), hi); }
template <> __device__ __forceinline__ uint32_t clamp_word<__half>(uint32_t x, uint32_t lo, uint32_t hi) {
    // @max / @min return the other operand when one is NaN, like __hmax2 / __hmin2
    const __half2 r = __hmin2(__hmax2(*reinterpret_cast<const __half2*>(&lo), *reinterpret_cast<const __half2*>(&x)), *reinterpret_cast<const __half2*>(&hi));
    return *reinterpret_cast<const uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t clamp_word<float>(uint32_t x, uint32_t lo, uint32_t hi) {
    return __float_as_uint(fminf(fmaxf(__uint_as_float(lo), __uint_as_float(x)), __uint_as_float(hi)));
}

template <typename T>
__global__ void __launch_bounds__(LNT) limiter_kernel(const BatchJob job, const LimiterParams prm) {
    int k = job.nplanes - 1;
    while (k > 0 && (int)blockIdx.x < job.pl[k].cta_begin) --k;
    const PlaneJob& pj = job.pl[k];
    const int local = (int)blockIdx.x - pj.cta_begin;
    const int yc = local * (LROWS * LGROUPS);
    const char* src = job.src + (size_t)blockIdx.y * job.src_fs + pj.src_off;
    char* dst = job.dst + (size_t)blockIdx.y * job.dst_fs + pj.dst_off;
    const uint32_t lo = prm.lo_w[pj.aux], hi = prm.hi_w[pj.aux];
    const int row_bytes = pj.w * (int)sizeof(T);
    const int nvec = row_bytes / 16;
    for (int y0 = yc; y0 < min(yc + LROWS * LGROUPS, pj.h); y0 += LROWS) {
    for (int v = threadIdx.x; v < nvec; v += LNT) {
        uint4 x[LROWS];
#pragma unroll
        for (int r = 0; r < LROWS; ++r) {
            const int y = min(y0 + r, pj.h - 1);
            x[r] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)y * pj.src_pitch) + v);
        }
#pragma unroll
        for (int r = 0; r < LROWS; ++r) {
            if (y0 + r < pj.h) {  // (a `break` here would keep the loop from unrolling and serialise the loads)
                uint4 o;
                o.x = clamp_word<T>(x[r].x, lo, hi); o.y = clamp_word<T>(x[r].y, lo, hi);
                o.z = clamp_word<T>(x[r].z, lo, hi); o.w = clamp_word<T>(x[r].w, lo, hi);
                reinterpret_cast<uint4*>(dst + (size_t)(y0 + r) * pj.dst_pitch)[v] = o;
            }
        }
    }
    // row tails (< 16 bytes): one sample per thread
    const int x0 = nvec * (16 / (int)sizeof(T)) + (int)threadIdx.x;
    if (x0 < pj.w) {
        constexpr int EPW = 4 / (int)sizeof(T);
        for (int r = 0; r < LROWS && y0 + r < pj.h; ++r) {
            const T s = reinterpret_cast<const T*>(src + (size_t)(y0 + r) * pj.src_pitch)[x0];
            uint32_t wv = 0u;
            reinterpret_cast<T*>(&wv)[0] = s;
            if (EPW > 1) { for (int e = 1; e < EPW; ++e) reinterpret_cast<T*>(&wv)[e] = s; }
            const uint32_t o = clamp_word<T>(wv, lo, hi);
            reinterpret_cast<T*>(dst + (size_t)(y0 + r) * pj.dst_pitch)[x0] = reinterpret_cast<const T*>(&o)[0];
        }
    }
    }
}

template <typename T>
static int launch_limiter_t(const BatchJob& job, const LimiterParams& prm, int count, cudaStream_t st) {
    for (int f0 = 0; f0 < count; f0 += 65535) {
        const int nf = std::min(65535, count - f0);
        BatchJob j = job;
        j.src += (size_t)f0 * job.src_fs; j.dst += (size_t)f0 * job.dst_fs;
        limiter_kernel<T><<<dim3(job.ctas_per_frame, nf), LNT, 0, st>>>(j, prm);
        count_launch();
    }
    VSZ_CUDA(cudaGetLastError());
    return 0;
}

// lo/hi: per plane bounds already rounded to the sample type (integers as exact doubles)
int run_limiter(const FrameLayout& l, const bool mask[3], const char* src, size_t src_fs, char* dst, size_t dst_fs, int count,
                const double lo[3], const double hi[3], cudaStream_t st) {
    if (count <= 0) return 0;
    BatchJob job = make_batch(l, mask, src, src_fs, nullptr, 0, dst, dst_fs, [](int, int h) { return (h + LROWS * LGROUPS - 1) / (LROWS * LGROUPS); });
    if (job.ctas_per_frame == 0) return 0;
    LimiterParams prm{};
    for (int p = 0; p < 3; ++p) {
        uint32_t a = 0, b = 0;
        switch (l.kind) {
            case K_U8: a = (uint32_t)lo[p] * 0x01010101u; b = (uint32_t)hi[p] * 0x01010101u; break;
            case K_U16: a = (uint32_t)lo[p] * 0x00010001u; b = (uint32_t)hi[p] * 0x00010001u; break;
            case K_F16: {
                const __half ha = __float2half_rn((float)lo[p]), hb = __float2half_rn((float)hi[p]);
                a = (uint32_t)(*reinterpret_cast<const uint16_t*>(&ha)) * 0x00010001u;
                b = (uint32_t)(*reinterpret_cast<const uint16_t*>(&hb)) * 0x00010001u;
                break;
            }
            case K_F32: {
                const float fa = (float)lo[p], fb = (float)hi[p];
                a = *reinterpret_cast<const uint32_t*>(&fa); b = *reinterpret_cast<const uint32_t*>(&fb);
                break;
            }
        }
        prm.lo_w[p] = a; prm.hi_w[p] = b;
    }
    switch (l.kind) {
        case K_U8: return launch_limiter_t<uint8_t>(job, prm, count, st);
        case K_U16: return launch_limiter_t<uint16_t>(job, prm, count, st);
        case K_F16: return launch_limiter_t<__half>(job, prm, count, st);
        case K_F32: return launch_limiter_t<float>(job, prm, count, st);
    }
    return -1;
}

}  // namespace vsz
