// filters.cu — host side of the four filters: argument parsing/validation with the reference's
// defaults and error strings (the twin of src/vapoursynth/{boxblur,bilateral,planeminmax,planeaverage}.zig
// create callbacks), the synchronous getFrame-style entry points (host buffers, staged through the
// per-request slot) and the batched device-resident entry points.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>

#include "filter.h"

using namespace vsz;

namespace {

// ZAPI getValue(u32/i32, ...) narrows the VapourSynth int64 with a saturating cast.
uint32_t sat_u32(int64_t v) { return v < 0 ? 0u : (v > 0xffffffffll ? 0xffffffffu : (uint32_t)v); }
int32_t sat_i32(int64_t v) { return v < INT32_MIN ? INT32_MIN : (v > INT32_MAX ? INT32_MAX : (int32_t)v); }

// hz.mapGetPlanes (src/helper.zig:128-158): an absent key keeps the filter's default mask.
bool parse_planes(const int64_t* planes, int n, int num_planes, const char* name, bool process[3]) {
    if (n < 0 || planes == nullptr) return true;
    process[0] = process[1] = process[2] = false;
    for (int i = 0; i < n; ++i) {
        const int64_t e = planes[i];
        if (e < 0 || e >= num_planes) { set_error("%s: plane index out of range", name); return false; }
        if (process[e]) { set_error("%s: plane specified twice.", name); return false; }
        process[e] = true;
    }
    return true;
}

// hz.compareNodes(..., .BIGGER_THAN, ...) (src/helper.zig:166-215)
bool compare_nodes(const vszip_video_info& a, const vszip_video_info& b, const char* name, bool same_len = false) {
    if (a.width != b.width || a.height != b.height) { set_error("%s: all input clips must have the same width and height.", name); return false; }
    if (a.color_family != b.color_family) { set_error("%s: all input clips must have the same color family.", name); return false; }
    if (a.sub_sampling_w != b.sub_sampling_w || a.sub_sampling_h != b.sub_sampling_h) { set_error("%s: all input clips must have the same subsampling.", name); return false; }
    if (a.bits_per_sample != b.bits_per_sample) { set_error("%s: all input clips must have the same bit depth.", name); return false; }
    if (same_len) {  // .SAME_LEN
        if (a.num_frames != b.num_frames) { set_error("%s: all input clips must have the same length.", name); return false; }
        return true;
    }
    if (a.num_frames > b.num_frames) { set_error("%s: second clip has less frames than input clip.", name); return false; }
    return true;
}

std::string fmt_num(double v) {  // Zig's {d}: shortest decimal that round-trips, never an exponent
    char buf[400];
    if (v == std::floor(v) && std::fabs(v) < 1e15) {
        snprintf(buf, sizeof buf, "%.0f", v);
        return buf;
    }
    for (int prec = 1; prec <= 17; ++prec) {
        snprintf(buf, sizeof buf, "%.*g", prec, v);
        if (strtod(buf, nullptr) == v) break;
    }
    if (strchr(buf, 'e')) snprintf(buf, sizeof buf, "%.17f", v);
    return buf;
}

std::string fmt_num_f32(float v) {  // {d} of an f32: shortest decimal that round-trips as f32
    char buf[400];
    if (v == std::floor(v) && std::fabs(v) < 1e15f) {
        snprintf(buf, sizeof buf, "%.0f", (double)v);
        return buf;
    }
    for (int prec = 1; prec <= 9; ++prec) {
        snprintf(buf, sizeof buf, "%.*g", prec, (double)v);
        if (strtof(buf, nullptr) == v) break;
    }
    if (strchr(buf, 'e')) snprintf(buf, sizeof buf, "%.9f", (double)v);
    return buf;
}

// hz.getArray(f32, ...) as LimitFilter uses it (src/helper.zig:340-404): values narrowed to f32 before the range tests
bool get_array_f32(const double* src, int n, float def, float lo, float hi, const char* key, const char* name, float out[3]) {
    if (n > 3) { set_error("%s: %s has too many elements (got %d, max 3).", name, key, n); return false; }
    for (int i = 0; i < 3; ++i) {
        if (i < n) out[i] = (float)src[i];
        else if (i == 0) out[i] = def;
        else out[i] = out[i - 1];
        if (out[i] < lo) { set_error("%s: %s value %s is below minimum %s.", name, key, fmt_num_f32(out[i]).c_str(), fmt_num_f32(lo).c_str()); return false; }
        if (out[i] > hi) { set_error("%s: %s value %s is above maximum %s.", name, key, fmt_num_f32(out[i]).c_str(), fmt_num_f32(hi).c_str()); return false; }
    }
    return true;
}

// hz.scaleValue(value, node, .{}) (src/helper.zig:312-338): an 8-bit integer-scale value brought to the clip's depth, f32
float scale_value8(float value, const vszip_video_info& vi, bool limited) {
    if (vi.bits_per_sample == 8) return value;
    const bool flt = vi.sample_type == VSZIP_ST_FLOAT;
    const int bits = vi.bits_per_sample, sh = bits - 8;
    const float full_peak = (float)((1ll << bits) - 1);
    const float in_peak = limited ? 235.0f : 255.0f, in_low = limited ? 16.0f : 0.0f;
    const float out_peak = flt ? 1.0f : (limited ? (float)(235ll << sh) : full_peak);
    const float out_low = flt ? 0.0f : (limited ? (float)(16ll << sh) : 0.0f);
    const float scale = (out_peak - out_low) / (in_peak - in_low);
    float v = value * scale;
    if (!flt) v = std::fmax(std::fmin(std::round(v), full_peak), 0.0f);
    return v;
}

// hz.getArray (src/helper.zig:340-404)
template <class T, class S>
bool get_array(const S* src, int n, T def, double lo, double hi, const char* key, const char* name, T out[3], T (*conv)(S)) {
    if (n > 3) { set_error("%s: %s has too many elements (got %d, max 3).", name, key, n); return false; }
    for (int i = 0; i < 3; ++i) {
        if (i < n) out[i] = conv(src[i]);
        else if (i == 0) out[i] = def;
        else out[i] = out[i - 1];
        if ((double)out[i] < lo) { set_error("%s: %s value %s is below minimum %s.", name, key, fmt_num((double)out[i]).c_str(), fmt_num(lo).c_str()); return false; }
        if ((double)out[i] > hi) { set_error("%s: %s value %s is above maximum %s.", name, key, fmt_num((double)out[i]).c_str(), fmt_num(hi).c_str()); return false; }
    }
    return true;
}

bool basic_vi_ok(const vszip_video_info* vi, const char* name) {
    if (!vi || vi->width <= 0 || vi->height <= 0 || vi->num_planes < 1 || vi->num_planes > 3) {
        set_error("%s: invalid video info", name);
        return false;
    }
    return true;
}

struct SlotGuard {
    DeviceCtx* d;
    Slot* s;
    explicit SlotGuard(DeviceCtx* dd) : d(dd), s(dd->acquire()) {}
    ~SlotGuard() { d->release(s); }
};

DeviceCtx* route(int32_t n, const char* name) {
    DeviceCtx* d = device_for_frame(n);
    if (!d) set_error("%s: vszip_cuda_init has not been called (no GPU context; there is no CPU fallback)", name);
    return d;
}

bool same_clip_shape(const vszip_dev_clip* a, const vszip_filter* f, const char* name) {
    if (!a || a->vi.width != f->vi.width || a->vi.height != f->vi.height || a->layout.kind != f->sample ||
        a->vi.num_planes != f->vi.num_planes || a->vi.sub_sampling_w != f->vi.sub_sampling_w ||
        a->vi.sub_sampling_h != f->vi.sub_sampling_h || a->vi.bits_per_sample != f->vi.bits_per_sample ||
        a->vi.sample_type != f->vi.sample_type) {
        set_error("%s: device clip does not match the format the filter was created for", name);
        return false;
    }
    return true;
}

bool range_ok(const vszip_dev_clip* c, int first, int count, const char* name) {
    if (first < 0 || count < 0 || first + count > c->num_frames) { set_error("%s: frame range out of bounds", name); return false; }
    return true;
}

}  // namespace

extern "C" {

void vszip_filter_free(vszip_filter* f) {
    if (!f) return;
    for (size_t d = 0; d < f->gr_dev.size(); ++d) {
        DeviceCtx* ctx = device_ctx((int)d);
        if (ctx) cudaSetDevice(ctx->ordinal);
        for (float* p : f->gr_dev[d]) if (p && cudaFree(p) != cudaSuccess) set_error("vszip_filter_free: cudaFree failed (%s)", cudaGetErrorString(cudaGetLastError()));
        for (float* p : f->gs_dev[d]) if (p && cudaFree(p) != cudaSuccess) set_error("vszip_filter_free: cudaFree failed (%s)", cudaGetErrorString(cudaGetLastError()));
    }
    for (size_t d = 0; d < f->exclude_i_dev.size(); ++d) {
        DeviceCtx* ctx = device_ctx((int)d);
        if (ctx) cudaSetDevice(ctx->ordinal);
        if (f->exclude_i_dev[d]) cudaFree(f->exclude_i_dev[d]);
        if (f->exclude_f_dev[d]) cudaFree(f->exclude_f_dev[d]);
    }
    delete f;
}

int vszip_filter_planes(const vszip_filter* f, int32_t process[3]) {
    for (int i = 0; i < 3; ++i) process[i] = (i < f->vi.num_planes && f->process[i]) ? 1 : 0;
    return 0;
}

// =========================================================================== BoxBlur
vszip_filter* vszip_boxblur_create(const vszip_video_info* vi, const vszip_boxblur_args* a) {
    static const char* name = "BoxBlur";
    if (!basic_vi_ok(vi, name)) return nullptr;
    SampleKind kind;
    if (!select_kind(*vi, name, false, &kind)) return nullptr;
    bool process[3] = {true, true, true};
    if (!parse_planes(a->planes, a->num_planes, vi->num_planes, name, process)) return nullptr;
    const uint32_t hr = a->has_hradius ? sat_u32(a->hradius) : 1u;
    const uint32_t vr = a->has_vradius ? sat_u32(a->vradius) : 1u;
    const int32_t hp = a->has_hpasses ? sat_i32(a->hpasses) : 1;
    const int32_t vp = a->has_vpasses ? sat_i32(a->vpasses) : 1;
    const bool vblur = vr > 0 && vp > 0, hblur = hr > 0 && hp > 0;
    if (!vblur && !hblur) { set_error("BoxBlur: nothing to be performed"); return nullptr; }
    // the comptime kernel blurs both axes with `hradius` whenever it is selected (boxblur.zig:188),
    // so the size checks must hold for both axes there
    const bool use_rt = (hr != vr) || (hr > 22) || (hp > 1) || (vp > 1);
    for (int p = 0; p < vi->num_planes; ++p) {
        if (!process[p]) continue;
        const uint32_t pw = (uint32_t)vi->width >> (p ? vi->sub_sampling_w : 0);
        const uint32_t ph = (uint32_t)vi->height >> (p ? vi->sub_sampling_h : 0);
        if ((hblur || !use_rt) && (uint64_t)hr * 2 >= pw) { set_error("BoxBlur: hradius too large; 2*hradius must be < the (smallest processed) plane width."); return nullptr; }
        if ((vblur || !use_rt) && (uint64_t)vr * 2 >= ph) { set_error("BoxBlur: vradius too large; 2*vradius must be < the (smallest processed) plane height."); return nullptr; }
    }
    vszip_filter* f = new vszip_filter();
    f->kind = F_BOXBLUR;
    f->vi = *vi;
    f->sample = kind;
    f->layout = make_layout(*vi, kind);
    for (int i = 0; i < 3; ++i) f->process[i] = process[i] && i < vi->num_planes;
    f->has_ref = false;
    f->hradius = hr; f->vradius = vr; f->hpasses = hp; f->vpasses = vp;
    return f;
}

int vszip_boxblur_get_frame(const vszip_filter* f, int32_t n, const vszip_frame* src, vszip_frame* dst) {
    if (!f || f->kind != F_BOXBLUR) { set_error("BoxBlur: bad filter handle"); return -1; }
    DeviceCtx* d = route(n, "BoxBlur");
    if (!d) return -1;
    SlotGuard g(d);
    Slot* s = g.s;
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    const size_t bytes = f->layout.frame_stride;
    if (slot_reserve(d, s, 0, bytes) || slot_reserve(d, s, 2, bytes)) return -1;
    if (stage_in(s, 0, f->layout, src, f->process)) return -1;
    int rc = run_boxblur(f->layout, f->process, s->dev[0], 0, s->dev[2], 0, 1, (int)f->hradius, f->hpasses, (int)f->vradius, f->vpasses, s->stream);
    if (rc) return rc;
    bool direct[3];
    if (stage_out_begin(s, f->layout, dst, f->process, direct)) return -1;
    VSZ_CUDA(cudaStreamSynchronize(s->stream));
    stage_out_finish(s, f->layout, dst, f->process, direct);
    return 0;
}

int vszip_boxblur_device(const vszip_filter* f, const vszip_dev_clip* src, vszip_dev_clip* dst, int32_t first, int32_t count, void* stream) {
    static const char* name = "BoxBlur";
    if (!f || f->kind != F_BOXBLUR) { set_error("BoxBlur: bad filter handle"); return -1; }
    if (!same_clip_shape(src, f, name) || !same_clip_shape(dst, f, name) || !range_ok(src, first, count, name) || !range_ok(dst, first, count, name)) return -1;
    if (src->device_index != dst->device_index) { set_error("BoxBlur: clips live on different devices"); return -1; }
    DeviceCtx* d = device_ctx(src->device_index);
    if (!d) { set_error("BoxBlur: library not initialised"); return -1; }
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->batch_stream;
    const size_t fs = src->layout.frame_stride;
    return run_boxblur(f->layout, f->process, src->base + (size_t)first * fs, fs, dst->base + (size_t)first * fs, fs, count,
                       (int)f->hradius, f->hpasses, (int)f->vradius, f->vpasses, st);
}

// =========================================================================== Bilateral
vszip_filter* vszip_bilateral_create(const vszip_video_info* vi, const vszip_video_info* ref_vi, const vszip_bilateral_args* a) {
    static const char* name = "Bilateral";
    if (!basic_vi_ok(vi, name)) return nullptr;
    SampleKind kind;
    if (!select_kind(*vi, name, false, &kind)) return nullptr;
    vszip_filter* f = new vszip_filter();
    auto fail = [&]() { vszip_filter_free(f); return (vszip_filter*)nullptr; };
    f->kind = F_BILATERAL;
    f->vi = *vi;
    f->sample = kind;
    f->layout = make_layout(*vi, kind);
    const bool yuv = vi->color_family == VSZIP_CF_YUV;
    f->hist_len = vi->sample_type == VSZIP_ST_INTEGER ? (1 << vi->bits_per_sample) : 65536;  // hz.getHistLen
    f->peak = (float)(f->hist_len - 1);

    double sigmaS[3], sigmaR[3];
    int32_t algorithm[3];
    uint32_t pbfic[3];
    for (int i = 0; i < 3; ++i) {  // bilateral.zig:104-125
        if (i < a->num_sigmaS) sigmaS[i] = a->sigmaS[i];
        else if (i == 0) sigmaS[0] = 3.0;
        else if (i == 1 && yuv && vi->sub_sampling_h != 0 && vi->sub_sampling_w != 0)
            sigmaS[1] = sigmaS[0] / std::sqrt((double)((1u << vi->sub_sampling_h) * (1u << vi->sub_sampling_w)));
        else sigmaS[i] = sigmaS[i - 1];
        if (sigmaS[i] < 0) { set_error("Bilateral: Invalid \"sigmaS\" assigned, must be non-negative float number"); return fail(); }
    }
    if (!get_array<double, double>(a->sigmaR, a->num_sigmaR, 0.02, 0.0, std::numeric_limits<double>::max(), "sigmaR", name, sigmaR, [](double v) { return v; })) return fail();
    if (!get_array<int32_t, int64_t>(a->algorithm, a->num_algorithm, 0, 0, 2, "algorithm", name, algorithm, [](int64_t v) { return sat_i32(v); })) return fail();
    if (!get_array<uint32_t, int64_t>(a->PBFICnum, a->num_PBFICnum, 0u, 0, 256, "PBFICnum", name, pbfic, [](int64_t v) { return sat_u32(v); })) return fail();
    bool process[3] = {true, true, true};
    if (!parse_planes(a->planes, a->num_planes, vi->num_planes, name, process)) return fail();
    for (int i = 0; i < 3; ++i)
        if (sigmaS[i] == 0 || sigmaR[i] == 0) process[i] = false;
    for (int i = 0; i < 3; ++i)
        if (pbfic[i] == 1) { set_error("Bilateral: Invalid \"PBFICnum\" assigned, must be integer ranges in [0,256] except 1"); return fail(); }
    for (int i = 0; i < 3; ++i) {  // bilateral.zig:147-162
        if (process[i] && pbfic[i] == 0) {
            if (sigmaR[i] >= 0.08) pbfic[i] = 4;
            else if (sigmaR[i] >= 0.015) pbfic[i] = std::min(16u, (uint32_t)std::trunc(4 * 0.08 / sigmaR[i] + 0.5));
            else pbfic[i] = std::min(32u, (uint32_t)std::trunc(16 * 0.015 / sigmaR[i] + 0.5));
            if (i > 0 && yuv && (pbfic[i] % 2 == 0) && pbfic[i] < 256) pbfic[i] += 1;
        }
    }
    for (int i = 0; i < 3; ++i) {  // bilateral.zig:164-199
        BilateralPlane& b = f->bl[i];
        b.sigmaS = sigmaS[i]; b.sigmaR = sigmaR[i]; b.algorithm = algorithm[i]; b.pbfic = pbfic[i];
        if (!process[i]) continue;
        const int orad = std::max((int)std::trunc(sigmaS[i] * 2 + 0.5), 1);
        b.step = orad < 4 ? 1 : (orad < 8 ? 2 : 3);
        b.samples = 1;
        b.radius = 1 + (b.samples - 1) * b.step;
        while (orad * 2 > (int)b.radius * 3) {
            b.samples += 1;
            b.radius = 1 + (b.samples - 1) * b.step;
            if ((int)b.radius >= orad && b.samples > 2) {
                b.samples -= 1;
                b.radius = 1 + (b.samples - 1) * b.step;
                break;
            }
        }
        if (b.algorithm <= 0)
            b.algorithm = (b.step == 1) ? 2 : ((sigmaR[i] < 0.08 && b.samples < 5) ? 2 : ((4 * b.samples * b.samples <= 15 * b.pbfic) ? 2 : 1));
    }
    for (int i = 0; i < vi->num_planes; ++i) {  // bilateral.zig:201-214
        if (process[i] && f->bl[i].algorithm == 2) {
            const uint32_t pw = (uint32_t)vi->width >> (i ? vi->sub_sampling_w : 0);
            const uint32_t ph = (uint32_t)vi->height >> (i ? vi->sub_sampling_h : 0);
            if (pw <= 2 * f->bl[i].radius || ph <= 2 * f->bl[i].radius) {
                set_error("Bilateral: plane too small for the spatial radius derived from sigmaS; lower sigmaS or use a larger clip.");
                return fail();
            }
        }
    }
    if (ref_vi && !compare_nodes(*vi, *ref_vi, name)) return fail();
    for (int i = 0; i < 3; ++i) f->process[i] = process[i] && i < vi->num_planes;
    f->has_ref = ref_vi != nullptr;
    // LUTs (src/filters/bilateral.zig:306-334), f64 math rounded to f32, uploaded to every device
    f->gs_host.resize(3); f->gr_host.resize(3);
    for (int i = 0; i < vi->num_planes; ++i) {
        if (!f->process[i]) continue;
        BilateralPlane& b = f->bl[i];
        const int upper = (int)b.radius + 1;
        f->gs_host[i].resize((size_t)upper * upper);
        for (int y = 0; y < upper; ++y)
            for (int x = 0; x < upper; ++x)
                f->gs_host[i][(size_t)y * upper + x] = (float)std::exp((double)(x * x + y * y) / (b.sigmaS * b.sigmaS * -2.0));
        const double range = (double)f->peak;
        const uint32_t top = (uint32_t)std::trunc(std::min(range, b.sigmaR * 8.0 * range + 0.5));
        const double norm = std::sqrt(2.0 * M_PI) * b.sigmaR;
        std::vector<float>& gr = f->gr_host[i];
        gr.resize((size_t)f->hist_len);
        uint32_t j = 0;
        for (; j <= top && j < (uint32_t)f->hist_len; ++j) {
            const double xx = ((double)j / range) / b.sigmaR;
            gr[j] = (float)(std::exp(xx * xx / -2.0) / norm);
        }
        const float tail = gr[std::min<uint32_t>(top, (uint32_t)f->hist_len - 1)];
        for (; j < (uint32_t)f->hist_len; ++j) gr[j] = tail;
        b.lut_len = (int)std::min<uint32_t>(top + 1, (uint32_t)f->hist_len);
        b.exact = (b.algorithm == 1) ? 1 : bilateral_weights_exact(b.lut_len);  // PBFIC always gathers the full LUT
    }
    return f;
}

int vszip_bilateral_get_info(const vszip_filter* f, vszip_bilateral_info* o) {
    if (!f || f->kind != F_BILATERAL) { set_error("Bilateral: bad filter handle"); return -1; }
    for (int i = 0; i < 3; ++i) {
        const BilateralPlane& b = f->bl[i];
        o->sigmaS[i] = b.sigmaS; o->sigmaR[i] = b.sigmaR; o->process[i] = f->process[i]; o->algorithm[i] = b.algorithm;
        o->PBFICnum[i] = b.pbfic; o->radius[i] = b.radius; o->samples[i] = b.samples; o->step[i] = b.step; o->exact_lut[i] = b.exact;
    }
    return 0;
}

// The weight tables are uploaded to a device the first time the instance runs there (create itself
// stays GPU-free so argument validation works anywhere).
static int bilateral_upload(const vszip_filter* cf, int dev_index) {
    vszip_filter* f = const_cast<vszip_filter*>(cf);
    std::lock_guard<std::mutex> lk(f->lut_mu);
    if (f->gr_dev.size() < (size_t)num_devices()) {
        f->gr_dev.resize(num_devices(), std::vector<float*>(3, nullptr));
        f->gs_dev.resize(num_devices(), std::vector<float*>(3, nullptr));
    }
    bool uploaded = false;
    for (int i = 0; i < f->vi.num_planes; ++i) {
        if (!f->process[i] || f->gr_dev[dev_index][i]) continue;
        const size_t gsb = f->gs_host[i].size() * sizeof(float), grb = f->gr_host[i].size() * sizeof(float);
        float *gs = nullptr, *gr = nullptr;
        // the pointers are published only after the copies have landed: the kernels run on non-blocking streams, which
        // are not ordered after a pageable cudaMemcpy on the default stream by themselves
        if (cudaMalloc((void**)&gs, gsb) != cudaSuccess || cudaMalloc((void**)&gr, grb) != cudaSuccess ||
            cudaMemcpy(gs, f->gs_host[i].data(), gsb, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(gr, f->gr_host[i].data(), grb, cudaMemcpyHostToDevice) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            set_error("Bilateral: uploading the weight tables failed (%s)", cudaGetErrorString(cudaGetLastError()));
            cudaFree(gs); cudaFree(gr);
            return -1;
        }
        f->gs_dev[dev_index][i] = gs;
        f->gr_dev[dev_index][i] = gr;
        uploaded = true;
    }
    (void)uploaded;
    return 0;
}

static BilateralLaunch bilateral_launch(const vszip_filter* f, int dev_index) {
    BilateralLaunch bp{};
    bp.peak = f->peak;
    for (int i = 0; i < 3; ++i) {
        bp.gs[i] = f->gs_dev[dev_index][i]; bp.gr[i] = f->gr_dev[dev_index][i];
        bp.radius[i] = (int)f->bl[i].radius; bp.step[i] = (int)f->bl[i].step; bp.lut_len[i] = f->bl[i].lut_len;
        if (f->process[i]) {
            const double s = (double)f->peak * f->bl[i].sigmaR;
            bp.c2[i] = (float)(-1.4426950408889634 / (2.0 * s * s));
            bp.cnorm[i] = (float)(1.0 / (std::sqrt(2.0 * M_PI) * f->bl[i].sigmaR));
        }
    }
    return bp;
}

// Algorithm 2 planes go through the tiled kernel in one launch; algorithm 1 (PBFIC) planes one plane at a time.
static int bilateral_run(const vszip_filter* f, int dev_index, const char* src, size_t sfs, const char* ref, size_t rfs, char* dst,
                         size_t dfs, int count, cudaStream_t st) {
    bool mask2[3] = {false, false, false}, any2 = false;
    for (int i = 0; i < 3; ++i) { mask2[i] = f->process[i] && f->bl[i].algorithm == 2; any2 = any2 || mask2[i]; }
    if (any2) {
        const BilateralLaunch bp = bilateral_launch(f, dev_index);
        const int rc = run_bilateral(f->layout, mask2, src, sfs, ref, rfs, dst, dfs, count, bp, st);
        if (rc) return rc;
    }
    for (int i = 0; i < 3; ++i) {
        if (!f->process[i] || f->bl[i].algorithm != 1) continue;
        const int rc = run_pbfic(f->layout, i, src, sfs, ref, rfs, dst, dfs, count, f->gr_dev[dev_index][i], f->hist_len, f->bl[i].sigmaS,
                                 (int)f->bl[i].pbfic, f->peak, st);
        if (rc) return rc;
    }
    return 0;
}

static int device_index_of(DeviceCtx* d) {
    for (int i = 0; i < num_devices(); ++i) if (device_ctx(i) == d) return i;
    return 0;
}

int vszip_bilateral_get_frame(const vszip_filter* f, int32_t n, const vszip_frame* src, const vszip_frame* ref, vszip_frame* dst) {
    if (!f || f->kind != F_BILATERAL) { set_error("Bilateral: bad filter handle"); return -1; }
    if (f->has_ref != (ref != nullptr)) { set_error("Bilateral: ref frame presence does not match the filter instance"); return -1; }
    DeviceCtx* d = route(n, "Bilateral");
    if (!d) return -1;
    SlotGuard g(d);
    Slot* s = g.s;
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    const size_t bytes = f->layout.frame_stride;
    if (slot_reserve(d, s, 0, bytes) || slot_reserve(d, s, 2, bytes) || (ref && slot_reserve(d, s, 1, bytes))) return -1;
    if (stage_in(s, 0, f->layout, src, f->process)) return -1;
    if (ref && stage_in(s, 1, f->layout, ref, f->process)) return -1;
    if (bilateral_upload(f, device_index_of(d))) return -1;
    int rc = bilateral_run(f, device_index_of(d), s->dev[0], 0, ref ? s->dev[1] : nullptr, 0, s->dev[2], 0, 1, s->stream);
    if (rc) return rc;
    bool direct[3];
    if (stage_out_begin(s, f->layout, dst, f->process, direct)) return -1;
    VSZ_CUDA(cudaStreamSynchronize(s->stream));
    stage_out_finish(s, f->layout, dst, f->process, direct);
    return 0;
}

int vszip_bilateral_device(const vszip_filter* f, const vszip_dev_clip* src, const vszip_dev_clip* ref, vszip_dev_clip* dst,
                           int32_t first, int32_t count, void* stream) {
    static const char* name = "Bilateral";
    if (!f || f->kind != F_BILATERAL) { set_error("Bilateral: bad filter handle"); return -1; }
    if (f->has_ref != (ref != nullptr)) { set_error("Bilateral: ref clip presence does not match the filter instance"); return -1; }
    if (!same_clip_shape(src, f, name) || !same_clip_shape(dst, f, name) || (ref && !same_clip_shape(ref, f, name))) return -1;
    if (!range_ok(src, first, count, name) || !range_ok(dst, first, count, name) || (ref && !range_ok(ref, first, count, name))) return -1;
    if (src->device_index != dst->device_index || (ref && ref->device_index != src->device_index)) { set_error("Bilateral: clips live on different devices"); return -1; }
    DeviceCtx* d = device_ctx(src->device_index);
    if (!d) { set_error("Bilateral: library not initialised"); return -1; }
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->batch_stream;
    const size_t fs = src->layout.frame_stride;
    if (bilateral_upload(f, src->device_index)) return -1;
    return bilateral_run(f, src->device_index, src->base + (size_t)first * fs, fs, ref ? ref->base + (size_t)first * fs : nullptr, fs,
                         dst->base + (size_t)first * fs, fs, count, st);
}

// =========================================================================== PlaneMinMax
vszip_filter* vszip_planeminmax_create(const vszip_video_info* vi, const vszip_video_info* clipb_vi, const vszip_planeminmax_args* a) {
    static const char* name = "PlaneMinMax";
    if (!basic_vi_ok(vi, name)) return nullptr;
    SampleKind kind;
    if (!select_kind(*vi, name, false, &kind)) return nullptr;
    if (clipb_vi && !compare_nodes(*vi, *clipb_vi, name)) return nullptr;
    bool process[3] = {true, false, false};
    if (!parse_planes(a->planes, a->num_planes, vi->num_planes, name, process)) return nullptr;
    const uint32_t hist_size = vi->sample_type == VSZIP_ST_FLOAT ? 65536u : (1u << vi->bits_per_sample);
    // getThr (planeminmax.zig:174-192): maxthr is read first
    const float maxthr = a->has_maxthr ? (float)a->maxthr : 0.0f;
    if (maxthr < 0 || maxthr > 1) { set_error("PlaneMinMax: maxthr should be a float between 0.0 and 1.0"); return nullptr; }
    const float minthr = a->has_minthr ? (float)a->minthr : 0.0f;
    if (minthr < 0 || minthr > 1) { set_error("PlaneMinMax: minthr should be a float between 0.0 and 1.0"); return nullptr; }
    const bool no_thr = maxthr == 0 && minthr == 0;
    const bool do_chroma = process[1] || process[2];
    if (do_chroma && !no_thr && vi->color_family == VSZIP_CF_YUV && vi->sample_type == VSZIP_ST_FLOAT) {
        set_error("PlaneMinMax: you can't use maxthr/minthr with float chroma, use planes=[0] or maxthr/minthr=0");
        return nullptr;
    }
    vszip_filter* f = new vszip_filter();
    f->kind = F_PLANEMINMAX;
    f->vi = *vi;
    f->sample = kind;
    f->layout = make_layout(*vi, kind);
    for (int i = 0; i < 3; ++i) f->process[i] = process[i] && i < vi->num_planes;
    f->has_ref = clipb_vi != nullptr;
    f->minthr = minthr; f->maxthr = maxthr; f->hist_size = hist_size; f->no_thr = no_thr;
    return f;
}

static void minmax_finalize(const vszip_filter* f, const StatsRaw* raw, vszip_minmax_props* out) {
    const bool flt = f->sample == K_F16 || f->sample == K_F32;
    memset(out, 0, sizeof *out);
    out->is_float = flt; out->has_diff = f->has_ref;
    const float peakf = (float)(f->hist_size - 1);  // planeminmax.zig:118-119
    int k = 0;
    for (int p = 0; p < f->layout.nplanes; ++p) {
        if (!f->process[p]) continue;
        const StatsRaw& r = raw[k];
        out->plane[k] = p;
        if (f->no_thr) {
            out->imin[k] = r.bin_min; out->imax[k] = r.bin_max;
            out->fmin[k] = (double)r.fmin; out->fmax[k] = (double)r.fmax;
        } else {
            out->imin[k] = r.bin_min; out->imax[k] = r.bin_max;
            out->fmin[k] = (double)((float)r.bin_min / 65535.0f);  // src/filters/planeminmax.zig:62-63
            out->fmax[k] = (double)((float)r.bin_max / 65535.0f);
        }
        if (f->has_ref) {
            const double total = (double)((uint32_t)f->layout.pl[p].w * (uint32_t)f->layout.pl[p].h);
            out->diff[k] = flt ? r.fdiff / total : (double)r.idiff / total / (double)peakf;
        }
        ++k;
    }
    out->count = k;
}

static int processed_planes(const vszip_filter* f) {
    int k = 0;
    for (int p = 0; p < f->layout.nplanes; ++p) k += f->process[p] ? 1 : 0;
    return k;
}

int vszip_planeminmax_get_frame(const vszip_filter* f, int32_t n, const vszip_frame* a, const vszip_frame* b, vszip_minmax_props* out) {
    if (!f || f->kind != F_PLANEMINMAX) { set_error("PlaneMinMax: bad filter handle"); return -1; }
    if (f->has_ref != (b != nullptr)) { set_error("PlaneMinMax: clipb frame presence does not match the filter instance"); return -1; }
    DeviceCtx* d = route(n, "PlaneMinMax");
    if (!d) return -1;
    SlotGuard g(d);
    Slot* s = g.s;
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    const size_t bytes = f->layout.frame_stride;
    const int np = processed_planes(f);
    const size_t scratch = stats_scratch_bytes(1, np) + 256;
    if (slot_reserve(d, s, 0, bytes) || (b && slot_reserve(d, s, 1, bytes)) || slot_reserve(d, s, 2, scratch)) return -1;
    if (stage_in(s, 0, f->layout, a, f->process)) return -1;
    if (b && stage_in(s, 1, f->layout, b, f->process)) return -1;
    StatsRaw* raw_dev = (StatsRaw*)s->dev_small;
    int rc = run_planeminmax(f->layout, f->process, s->dev[0], 0, b ? s->dev[1] : nullptr, 0, 1, f->no_thr, f->minthr, f->maxthr,
                             f->hist_size, s->dev[2], raw_dev, s->stream);
    if (rc) return rc;
    VSZ_CUDA(cudaMemcpyAsync(s->pin_small, raw_dev, sizeof(StatsRaw) * np, cudaMemcpyDeviceToHost, s->stream));
    VSZ_CUDA(cudaStreamSynchronize(s->stream));
    minmax_finalize(f, (const StatsRaw*)s->pin_small, out);
    return 0;
}

int vszip_planeminmax_device(const vszip_filter* f, const vszip_dev_clip* a, const vszip_dev_clip* b, int32_t first, int32_t count,
                             vszip_minmax_props* out, void* stream) {
    static const char* name = "PlaneMinMax";
    if (!f || f->kind != F_PLANEMINMAX) { set_error("PlaneMinMax: bad filter handle"); return -1; }
    if (f->has_ref != (b != nullptr)) { set_error("PlaneMinMax: clipb presence does not match the filter instance"); return -1; }
    if (!same_clip_shape(a, f, name) || (b && !same_clip_shape(b, f, name)) || !range_ok(a, first, count, name) || (b && !range_ok(b, first, count, name))) return -1;
    if (b && b->device_index != a->device_index) { set_error("%s: clips live on different devices", name); return -1; }
    DeviceCtx* d = device_ctx(a->device_index);
    if (!d) { set_error("PlaneMinMax: library not initialised"); return -1; }
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->batch_stream;
    const int np = processed_planes(f);
    if (count == 0 || np == 0) return 0;
    const size_t fs = a->layout.frame_stride;
    AsyncScratch scratch_mem;
    const size_t sb = stats_scratch_bytes(count, np), rb = sizeof(StatsRaw) * (size_t)count * np;
    VSZ_CUDA(scratch_mem.alloc(sb + rb, st));
    char* scratch = scratch_mem.p;
    StatsRaw* raw_dev = (StatsRaw*)(scratch + sb);
    int rc = run_planeminmax(f->layout, f->process, a->base + (size_t)first * fs, fs, b ? b->base + (size_t)first * fs : nullptr, fs, count,
                             f->no_thr, f->minthr, f->maxthr, f->hist_size, scratch, raw_dev, st);
    std::vector<StatsRaw> raw((size_t)count * np);
    if (!rc && out) {
        VSZ_CUDA(cudaMemcpyAsync(raw.data(), raw_dev, rb, cudaMemcpyDeviceToHost, st));
        VSZ_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < count; ++i) minmax_finalize(f, raw.data() + (size_t)i * np, out + i);
    }
    return rc;
}

// =========================================================================== PlaneAverage
vszip_filter* vszip_planeaverage_create(const vszip_video_info* vi, const vszip_video_info* clipb_vi, const vszip_planeaverage_args* a) {
    static const char* name = "PlaneAverage";
    if (!basic_vi_ok(vi, name)) return nullptr;
    if (a->num_exclude < 0 || (a->num_exclude > 0 && !a->exclude)) {  // VapourSynth rejects a missing non-opt argument itself
        set_error("PlaneAverage: argument exclude is required");
        return nullptr;
    }
    const bool u32 = vi->sample_type == VSZIP_ST_INTEGER && vi->bytes_per_sample == 4;
    SampleKind kind = K_U16;
    if (!u32 && !select_kind(*vi, name, true, &kind)) return nullptr;
    if (clipb_vi && !compare_nodes(*vi, *clipb_vi, name)) return nullptr;
    bool process[3] = {true, false, false};
    if (!parse_planes(a->planes, a->num_planes, vi->num_planes, name, process)) return nullptr;
    if (u32) { set_error("PlaneAverage: exclude is not supported for 32-bit integer clips."); return nullptr; }
    vszip_filter* f = new vszip_filter();
    f->kind = F_PLANEAVERAGE;
    f->vi = *vi;
    f->sample = kind;
    f->layout = make_layout(*vi, kind);
    for (int i = 0; i < 3; ++i) f->process[i] = process[i] && i < vi->num_planes;
    f->has_ref = clipb_vi != nullptr;
    f->avg_peak = (float)((1ull << vi->bits_per_sample) - 1ull);  // planeaverage.zig:112
    for (int i = 0; i < a->num_exclude; ++i) {
        f->exclude_f.push_back((float)a->exclude[i]);           // planeaverage.zig:122-125
        f->exclude_i.push_back(sat_i32(a->exclude[i]));         // math.lossyCast(i32, ..)
    }
    return f;
}

// Exclude lists longer than the 16 entries that travel in the kernel arguments are uploaded to a device the first
// time the instance runs there.
static int average_upload(const vszip_filter* cf, int dev_index, const int32_t** xi, const float** xf) {
    *xi = nullptr; *xf = nullptr;
    if (cf->exclude_i.size() <= 16) return 0;
    vszip_filter* f = const_cast<vszip_filter*>(cf);
    std::lock_guard<std::mutex> lk(f->lut_mu);
    if (f->exclude_i_dev.size() < (size_t)num_devices()) { f->exclude_i_dev.resize(num_devices(), nullptr); f->exclude_f_dev.resize(num_devices(), nullptr); }
    if (!f->exclude_i_dev[dev_index]) {
        const size_t n = f->exclude_i.size();
        int32_t* di = nullptr;
        float* df = nullptr;
        // published only once the copies have landed (the reductions run on non-blocking streams)
        if (cudaMalloc((void**)&di, n * sizeof(int32_t)) != cudaSuccess || cudaMalloc((void**)&df, n * sizeof(float)) != cudaSuccess ||
            cudaMemcpy(di, f->exclude_i.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(df, f->exclude_f.data(), n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            set_error("PlaneAverage: uploading the exclude list failed (%s)", cudaGetErrorString(cudaGetLastError()));
            cudaFree(di); cudaFree(df);
            return -1;
        }
        f->exclude_i_dev[dev_index] = di;
        f->exclude_f_dev[dev_index] = df;
    }
    *xi = f->exclude_i_dev[dev_index]; *xf = f->exclude_f_dev[dev_index];
    return 0;
}

static void average_finalize(const vszip_filter* f, const StatsRaw* raw, vszip_average_props* out) {
    const bool flt = f->sample == K_F16 || f->sample == K_F32;
    memset(out, 0, sizeof *out);
    out->has_diff = f->has_ref;
    int k = 0;
    for (int p = 0; p < f->layout.nplanes; ++p) {
        if (!f->process[p]) continue;
        const StatsRaw& r = raw[k];
        out->plane[k] = p;
        const uint32_t all = (uint32_t)f->layout.pl[p].w * (uint32_t)f->layout.pl[p].h;
        const double total = (double)(all - r.excluded);
        // result() (src/filters/planeaverage.zig:16-24)
        if (total == 0) out->avg[k] = 0.0;
        else if (flt) out->avg[k] = r.fsum / total;
        else out->avg[k] = (double)r.isum / total / (double)f->avg_peak;
        if (f->has_ref) out->diff[k] = flt ? r.fdiff / (double)all : (double)r.idiff / (double)all / (double)f->avg_peak;
        ++k;
    }
    out->count = k;
}

int vszip_planeaverage_get_frame(const vszip_filter* f, int32_t n, const vszip_frame* a, const vszip_frame* b, vszip_average_props* out) {
    if (!f || f->kind != F_PLANEAVERAGE) { set_error("PlaneAverage: bad filter handle"); return -1; }
    if (f->has_ref != (b != nullptr)) { set_error("PlaneAverage: clipb frame presence does not match the filter instance"); return -1; }
    DeviceCtx* d = route(n, "PlaneAverage");
    if (!d) return -1;
    SlotGuard g(d);
    Slot* s = g.s;
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    const size_t bytes = f->layout.frame_stride;
    const int np = processed_planes(f);
    const size_t scratch = stats_scratch_bytes(1, np) + 256;
    if (slot_reserve(d, s, 0, bytes) || (b && slot_reserve(d, s, 1, bytes)) || slot_reserve(d, s, 2, scratch)) return -1;
    if (stage_in(s, 0, f->layout, a, f->process)) return -1;
    if (b && stage_in(s, 1, f->layout, b, f->process)) return -1;
    StatsRaw* raw_dev = (StatsRaw*)s->dev_small;
    const int32_t* xi; const float* xf;
    if (average_upload(f, device_index_of(d), &xi, &xf)) return -1;
    int rc = run_planeaverage(f->layout, f->process, s->dev[0], 0, b ? s->dev[1] : nullptr, 0, 1, f->exclude_i.data(), f->exclude_f.data(),
                              (int)f->exclude_i.size(), xi, xf, s->dev[2], raw_dev, s->stream);
    if (rc) return rc;
    VSZ_CUDA(cudaMemcpyAsync(s->pin_small, raw_dev, sizeof(StatsRaw) * np, cudaMemcpyDeviceToHost, s->stream));
    VSZ_CUDA(cudaStreamSynchronize(s->stream));
    average_finalize(f, (const StatsRaw*)s->pin_small, out);
    return 0;
}

int vszip_planeaverage_device(const vszip_filter* f, const vszip_dev_clip* a, const vszip_dev_clip* b, int32_t first, int32_t count,
                              vszip_average_props* out, void* stream) {
    static const char* name = "PlaneAverage";
    if (!f || f->kind != F_PLANEAVERAGE) { set_error("PlaneAverage: bad filter handle"); return -1; }
    if (f->has_ref != (b != nullptr)) { set_error("PlaneAverage: clipb presence does not match the filter instance"); return -1; }
    if (!same_clip_shape(a, f, name) || (b && !same_clip_shape(b, f, name)) || !range_ok(a, first, count, name) || (b && !range_ok(b, first, count, name))) return -1;
    if (b && b->device_index != a->device_index) { set_error("%s: clips live on different devices", name); return -1; }
    DeviceCtx* d = device_ctx(a->device_index);
    if (!d) { set_error("PlaneAverage: library not initialised"); return -1; }
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->batch_stream;
    const int np = processed_planes(f);
    if (count == 0 || np == 0) return 0;
    const size_t fs = a->layout.frame_stride;
    AsyncScratch scratch_mem;
    const size_t sb = stats_scratch_bytes(count, np), rb = sizeof(StatsRaw) * (size_t)count * np;
    VSZ_CUDA(scratch_mem.alloc(sb + rb, st));
    char* scratch = scratch_mem.p;
    StatsRaw* raw_dev = (StatsRaw*)(scratch + sb);
    const int32_t* xi; const float* xf;
    if (average_upload(f, a->device_index, &xi, &xf)) return -1;
    int rc = run_planeaverage(f->layout, f->process, a->base + (size_t)first * fs, fs, b ? b->base + (size_t)first * fs : nullptr, fs, count,
                              f->exclude_i.data(), f->exclude_f.data(), (int)f->exclude_i.size(), xi, xf, scratch, raw_dev, st);
    std::vector<StatsRaw> raw((size_t)count * np);
    if (!rc && out) {
        VSZ_CUDA(cudaMemcpyAsync(raw.data(), raw_dev, rb, cudaMemcpyDeviceToHost, st));
        VSZ_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < count; ++i) average_finalize(f, raw.data() + (size_t)i * np, out + i);
    }
    return rc;
}

// =========================================================================== PlaneMinMax + PlaneAverage from one read
int vszip_planestats_device(const vszip_filter* fm, const vszip_filter* fa, const vszip_dev_clip* a, int32_t first, int32_t count,
                            vszip_minmax_props* mm_out, vszip_average_props* avg_out, int32_t* fused_out, void* stream) {
    static const char* name = "PlaneStats";
    if (fused_out) *fused_out = 0;
    if (!fm || fm->kind != F_PLANEMINMAX || !fa || fa->kind != F_PLANEAVERAGE) { set_error("PlaneStats: a PlaneMinMax and a PlaneAverage handle are required"); return -1; }
    if (fm->has_ref || fa->has_ref) { set_error("PlaneStats: filters created with clipb cannot be combined"); return -1; }
    if (!same_clip_shape(a, fm, name) || !same_clip_shape(a, fa, name) || !range_ok(a, first, count, name)) return -1;
    DeviceCtx* d = device_ctx(a->device_index);
    if (!d) { set_error("PlaneStats: library not initialised"); return -1; }
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->batch_stream;
    const int npm = processed_planes(fm), npa = processed_planes(fa);
    if (count == 0) return 0;
    const size_t fs = a->layout.frame_stride;
    const char* base = a->base + (size_t)first * fs;
    const bool same_planes = fm->process[0] == fa->process[0] && fm->process[1] == fa->process[1] && fm->process[2] == fa->process[2];
    const size_t sbm = stats_scratch_bytes(count, std::max(npm, 1)), sba = stats_scratch_bytes(count, std::max(npa, 1));
    const size_t rbm = sizeof(StatsRaw) * (size_t)count * npm, rba = sizeof(StatsRaw) * (size_t)count * npa;
    const size_t rbm_al = (rbm + 255) & ~(size_t)255;
    AsyncScratch scratch_mem;
    VSZ_CUDA(scratch_mem.alloc(sbm + sba + rbm_al + rba + 256, st));
    char* scratch = scratch_mem.p;
    StatsRaw* raw_m = (StatsRaw*)(scratch + sbm + sba);
    StatsRaw* raw_a = (StatsRaw*)(scratch + sbm + sba + rbm_al);
    int rc = 1;
    if (same_planes && npm > 0 && fa->exclude_i.size() <= 16)
        rc = run_planestats_fused(fm->layout, fm->process, base, fs, count, fm->no_thr, fm->minthr, fm->maxthr, fm->hist_size,
                                  fa->exclude_i.data(), fa->exclude_f.data(), (int)fa->exclude_i.size(), scratch, raw_m, raw_a, st);
    if (rc == 0 && fused_out) *fused_out = 1;
    if (rc == 1) {  // not eligible: the two reductions one after the other (two reads)
        rc = 0;
        if (npm) rc = run_planeminmax(fm->layout, fm->process, base, fs, nullptr, fs, count, fm->no_thr, fm->minthr, fm->maxthr, fm->hist_size, scratch, raw_m, st);
        const int32_t* xi; const float* xf;
        if (!rc && npa && average_upload(fa, a->device_index, &xi, &xf)) rc = -1;
        if (!rc && npa) rc = run_planeaverage(fa->layout, fa->process, base, fs, nullptr, fs, count, fa->exclude_i.data(), fa->exclude_f.data(),
                                               (int)fa->exclude_i.size(), xi, xf, scratch + sbm, raw_a, st);
    }
    if (!rc && (mm_out || avg_out)) {
        std::vector<StatsRaw> hm((size_t)count * npm), ha((size_t)count * npa);
        if (mm_out && npm) VSZ_CUDA(cudaMemcpyAsync(hm.data(), raw_m, rbm, cudaMemcpyDeviceToHost, st));
        if (avg_out && npa) VSZ_CUDA(cudaMemcpyAsync(ha.data(), raw_a, rba, cudaMemcpyDeviceToHost, st));
        VSZ_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < count; ++i) {
            if (mm_out) minmax_finalize(fm, hm.data() + (size_t)i * npm, mm_out + i);
            if (avg_out) average_finalize(fa, ha.data() + (size_t)i * npa, avg_out + i);
        }
    }
    return rc;
}

// =========================================================================== Limiter
// Range tables of src/filters/limiter.zig:66-91: [lo|hi][plane] for full / tv-range YUV / tv-range RGB at 8 bits;
// deeper integer formats are the 8-bit values shifted left by (bits - 8), full range is (1 << bits) - 1.
static void limiter_table(const vszip_video_info& vi, bool tv_range, bool yuv, double lo[3], double hi[3]) {
    if (vi.sample_type == VSZIP_ST_FLOAT) {  // floats ignore tv_range (limiter.zig:53-54 of the filter file)
        for (int p = 0; p < 3; ++p) { lo[p] = (yuv && p > 0) ? -0.5 : 0.0; hi[p] = (yuv && p > 0) ? 0.5 : 1.0; }
        return;
    }
    const int sh = vi.bits_per_sample - 8;  // up to 24 (32-bit integer clips: full32 / yuv32 / rgb32 of src/filters/limiter.zig:72-91)
    for (int p = 0; p < 3; ++p) {
        if (!tv_range) { lo[p] = 0.0; hi[p] = (double)((1ull << vi.bits_per_sample) - 1ull); }
        else { lo[p] = (double)(16ull << sh); hi[p] = (double)(((yuv && p > 0) ? 240ull : 235ull) << sh); }
    }
}

vszip_filter* vszip_limiter_create(const vszip_video_info* vi, const vszip_limiter_args* a) {
    static const char* name = "Limiter";
    if (!basic_vi_ok(vi, name)) return nullptr;
    const int np = vi->num_planes;
    const bool is_int = vi->sample_type == VSZIP_ST_INTEGER;
    const double peak = is_int ? (double)(float)((1ll << vi->bits_per_sample) - 1) : 1.0;  // getPeakValue(.., false, .FULL) is f32
    bool process[3] = {true, true, true};
    if (!parse_planes(a->planes, a->num_planes, np, name, process)) return nullptr;
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    const bool has_min = a->num_min >= 0, has_max = a->num_max >= 0;
    if (has_min) {  // limiter.zig:121-150
        if (a->num_min != np) { set_error("Limiter: min array must have the same number of elements as planes."); return nullptr; }
        for (int i = 0; i < np; ++i) {
            if (is_int) {
                const double t = std::trunc(a->min[i]);
                if (t < 0) { set_error("Limiter: min value must be greater than or equal to 0."); return nullptr; }
                if (a->min[i] > peak) { set_error("Limiter: min value must be less than or equal to peak value."); return nullptr; }
                lo[i] = t;
            } else {
                lo[i] = (double)(float)a->min[i];
            }
        }
    }
    if (has_max) {  // limiter.zig:152-181 (peak is tested before the sign here)
        if (a->num_max != np) { set_error("Limiter: max array must have the same number of elements as planes."); return nullptr; }
        for (int i = 0; i < np; ++i) {
            if (is_int) {
                const double t = std::trunc(a->max[i]);
                if (a->max[i] > peak) { set_error("Limiter: max value must be less than or equal to peak value."); return nullptr; }
                if (t < 0) { set_error("Limiter: max value must be greater than or equal to 0."); return nullptr; }
                hi[i] = t;
            } else {
                hi[i] = (double)(float)a->max[i];
            }
        }
    }
    if (has_min && !has_max) { set_error("Limiter: min array is set but max array is not."); return nullptr; }
    if (!has_min && has_max) { set_error("Limiter: max array is set but min array is not."); return nullptr; }
    if (has_min) {
        for (int p = 0; p < np; ++p)
            if (lo[p] > hi[p]) { set_error("Limiter: min value must be less than or equal to max value."); return nullptr; }
    }
    // BPSType.select (src/helper.zig:25-56)
    if (is_int) {
        const int b = vi->bits_per_sample;
        if (!(b == 8 || b == 9 || b == 10 || b == 12 || b == 14 || b == 16 || b == 32)) { set_error("Limiter: not supported Int format."); return nullptr; }
    } else if (!(vi->bits_per_sample == 16 || vi->bits_per_sample == 32)) {
        set_error("Limiter: not supported Float format.");
        return nullptr;
    }
    SampleKind kind;
    if (!select_kind(*vi, name, true, &kind)) return nullptr;
    if (kind == K_U32)  // the f32 peak of a 32-bit clip is 2^32: a bound that passed the check above may be one past the largest sample
        for (int i = 0; i < np; ++i) { lo[i] = std::min(lo[i], 4294967295.0); hi[i] = std::min(hi[i], 4294967295.0); }
    const bool tv_range = a->has_tv_range && a->tv_range != 0, maskf = a->has_mask && a->mask != 0;
    const bool yuv = vi->color_family == VSZIP_CF_YUV && !maskf;
    if (!has_min) limiter_table(*vi, tv_range, yuv, lo, hi);
    vszip_filter* f = new vszip_filter();
    f->kind = F_LIMITER;
    f->vi = *vi;
    f->sample = kind;
    f->layout = make_layout(*vi, kind);
    for (int i = 0; i < 3; ++i) { f->process[i] = process[i] && i < np; f->lim_lo[i] = lo[i]; f->lim_hi[i] = hi[i]; }
    f->has_ref = false;
    return f;
}

int vszip_limiter_get_frame(const vszip_filter* f, int32_t n, const vszip_frame* src, vszip_frame* dst) {
    if (!f || f->kind != F_LIMITER) { set_error("Limiter: bad filter handle"); return -1; }
    DeviceCtx* d = route(n, "Limiter");
    if (!d) return -1;
    SlotGuard g(d);
    Slot* s = g.s;
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    const size_t bytes = f->layout.frame_stride;
    if (slot_reserve(d, s, 0, bytes) || slot_reserve(d, s, 2, bytes)) return -1;
    if (stage_in(s, 0, f->layout, src, f->process)) return -1;
    int rc = run_limiter(f->layout, f->process, s->dev[0], 0, s->dev[2], 0, 1, f->lim_lo, f->lim_hi, s->stream);
    if (rc) return rc;
    bool direct[3];
    if (stage_out_begin(s, f->layout, dst, f->process, direct)) return -1;
    VSZ_CUDA(cudaStreamSynchronize(s->stream));
    stage_out_finish(s, f->layout, dst, f->process, direct);
    return 0;
}

int vszip_limiter_device(const vszip_filter* f, const vszip_dev_clip* src, vszip_dev_clip* dst, int32_t first, int32_t count, void* stream) {
    static const char* name = "Limiter";
    if (!f || f->kind != F_LIMITER) { set_error("Limiter: bad filter handle"); return -1; }
    if (!same_clip_shape(src, f, name) || !same_clip_shape(dst, f, name) || !range_ok(src, first, count, name) || !range_ok(dst, first, count, name)) return -1;
    if (src->device_index != dst->device_index) { set_error("Limiter: clips live on different devices"); return -1; }
    DeviceCtx* d = device_ctx(src->device_index);
    if (!d) { set_error("Limiter: library not initialised"); return -1; }
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->batch_stream;
    const size_t fs = src->layout.frame_stride;
    return run_limiter(f->layout, f->process, src->base + (size_t)first * fs, fs, dst->base + (size_t)first * fs, fs, count, f->lim_lo, f->lim_hi, st);
}

// =========================================================================== LimitFilter
vszip_filter* vszip_limitfilter_create(const vszip_video_info* flt_vi, const vszip_video_info* src_vi, const vszip_video_info* ref_vi,
                                       const vszip_limitfilter_args* a) {
    static const char* name = "LimitFilter";
    if (!basic_vi_ok(flt_vi, name)) return nullptr;
    if (!src_vi) { set_error("LimitFilter: the src clip is required"); return nullptr; }
    static const vszip_limitfilter_args none = {nullptr, 0, nullptr, 0, nullptr, 0, nullptr, -1, -1};
    if (!a) a = &none;
    SampleKind kind;
    if (!select_kind(*flt_vi, name, false, &kind)) return nullptr;                     // limit_filter.zig:101
    if (!compare_nodes(*flt_vi, *src_vi, name, true)) return nullptr;                  // :107-108, SAME_LEN
    if (ref_vi && !compare_nodes(*flt_vi, *ref_vi, name, true)) return nullptr;
    bool process[3] = {true, true, true};
    if (!parse_planes(a->planes, a->num_planes, flt_vi->num_planes, name, process)) return nullptr;
    float dark[3], bright[3], elast[3];
    if (!get_array_f32(a->dark_thr, a->num_dark_thr, 1.0f, 0.0f, 255.0f, "dark_thr", name, dark)) return nullptr;
    if (!get_array_f32(a->bright_thr, a->num_bright_thr, 1.0f, 0.0f, 255.0f, "bright_thr", name, bright)) return nullptr;
    if (!get_array_f32(a->elast, a->num_elast, 2.0f, 0.0f, 65535.0f, "elast", name, elast)) return nullptr;
    // getColorRange (src/helper.zig:259-276) when the frame carries no _ColorRange
    const bool limited = a->color_range < 0 ? flt_vi->color_family != VSZIP_CF_RGB : a->color_range == 1;
    vszip_filter* f = new vszip_filter();
    f->kind = F_LIMITFILTER;
    f->vi = *flt_vi;
    f->sample = kind;
    f->layout = make_layout(*flt_vi, kind);
    for (int i = 0; i < 3; ++i) {
        f->process[i] = process[i] && i < flt_vi->num_planes;
        f->lf_dark[i] = scale_value8(dark[i], *flt_vi, limited);                       // :114-118
        f->lf_bright[i] = scale_value8(bright[i], *flt_vi, limited);
        f->lf_elast[i] = elast[i];
    }
    f->lf_has_ref = ref_vi != nullptr;
    f->has_ref = true;  // multi-input: not fusable into a linear chain
    return f;
}

int vszip_limitfilter_get_info(const vszip_filter* f, float dark_thr[3], float bright_thr[3], float elast[3]) {
    if (!f || f->kind != F_LIMITFILTER) { set_error("LimitFilter: bad filter handle"); return -1; }
    for (int i = 0; i < 3; ++i) { dark_thr[i] = f->lf_dark[i]; bright_thr[i] = f->lf_bright[i]; elast[i] = f->lf_elast[i]; }
    return 0;
}

int vszip_limitfilter_get_frame(const vszip_filter* f, int32_t n, const vszip_frame* flt, const vszip_frame* src, const vszip_frame* ref,
                                vszip_frame* dst) {
    if (!f || f->kind != F_LIMITFILTER) { set_error("LimitFilter: bad filter handle"); return -1; }
    if (!flt || !src || !dst) { set_error("LimitFilter: flt, src and dst frames are required"); return -1; }
    if (f->lf_has_ref != (ref != nullptr)) { set_error("LimitFilter: ref frame presence does not match the filter instance"); return -1; }
    DeviceCtx* d = route(n, "LimitFilter");
    if (!d) return -1;
    SlotGuard g(d);
    Slot* s = g.s;
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    const size_t bytes = f->layout.frame_stride;
    if (slot_reserve(d, s, 0, bytes) || slot_reserve(d, s, 1, bytes) || slot_reserve(d, s, 2, bytes) || (ref && slot_reserve(d, s, 3, bytes))) return -1;
    if (stage_in(s, 0, f->layout, flt, f->process) || stage_in(s, 1, f->layout, src, f->process)) return -1;
    if (ref && stage_in(s, 3, f->layout, ref, f->process)) return -1;
    int rc = run_limitfilter(f->layout, f->process, s->dev[0], 0, s->dev[1], 0, ref ? s->dev[3] : nullptr, 0, s->dev[2], 0, 1, f->lf_dark,
                             f->lf_bright, f->lf_elast, s->stream);
    if (rc) return rc;
    bool direct[3];
    if (stage_out_begin(s, f->layout, dst, f->process, direct)) return -1;
    VSZ_CUDA(cudaStreamSynchronize(s->stream));
    stage_out_finish(s, f->layout, dst, f->process, direct);
    return 0;
}

int vszip_limitfilter_device(const vszip_filter* f, const vszip_dev_clip* flt, const vszip_dev_clip* src, const vszip_dev_clip* ref,
                             vszip_dev_clip* dst, int32_t first, int32_t count, void* stream) {
    static const char* name = "LimitFilter";
    if (!f || f->kind != F_LIMITFILTER) { set_error("LimitFilter: bad filter handle"); return -1; }
    if (f->lf_has_ref != (ref != nullptr)) { set_error("LimitFilter: ref clip presence does not match the filter instance"); return -1; }
    if (!same_clip_shape(flt, f, name) || !same_clip_shape(src, f, name) || !same_clip_shape(dst, f, name) || (ref && !same_clip_shape(ref, f, name))) return -1;
    if (!range_ok(flt, first, count, name) || !range_ok(src, first, count, name) || !range_ok(dst, first, count, name) || (ref && !range_ok(ref, first, count, name))) return -1;
    if (flt->device_index != dst->device_index || src->device_index != dst->device_index || (ref && ref->device_index != dst->device_index)) {
        set_error("LimitFilter: clips live on different devices");
        return -1;
    }
    DeviceCtx* d = device_ctx(dst->device_index);
    if (!d) { set_error("LimitFilter: library not initialised"); return -1; }
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->batch_stream;
    const size_t fs = dst->layout.frame_stride, off = (size_t)first * fs;
    return run_limitfilter(f->layout, f->process, flt->base + off, fs, src->base + off, fs, ref ? ref->base + off : nullptr, fs, dst->base + off, fs,
                           count, f->lf_dark, f->lf_bright, f->lf_elast, st);
}

// =========================================================================== AdaptiveBinarize
vszip_filter* vszip_adaptivebinarize_create(const vszip_video_info* vi, const vszip_video_info* clip2_vi, const vszip_adaptivebinarize_args* a) {
    static const char* name = "AdaptiveBinarize";
    if (!basic_vi_ok(vi, name)) return nullptr;
    if (!clip2_vi) { set_error("AdaptiveBinarize: clip2 is required"); return nullptr; }
    if (!compare_nodes(*vi, *clip2_vi, name)) return nullptr;                          // adaptive_binarize.zig:88-89, BIGGER_THAN
    if (vi->sample_type != VSZIP_ST_INTEGER || vi->bits_per_sample != 8) { set_error("AdaptiveBinarize: only 8 bit int format supported."); return nullptr; }
    const int64_t c = a && a->has_c ? (int64_t)sat_i32(a->c) : 3;                      // getValue(i32, "c") orelse 3
    vszip_filter* f = new vszip_filter();
    f->kind = F_ADAPTIVEBINARIZE;
    f->vi = *vi;
    f->sample = K_U8;
    f->layout = make_layout(*vi, K_U8);
    for (int i = 0; i < 3; ++i) f->process[i] = i < vi->num_planes;
    f->ab_c = (int)std::min<int64_t>(std::max<int64_t>(c, -256), 256);               // :97-99
    f->has_ref = true;
    return f;
}

int vszip_adaptivebinarize_get_frame(const vszip_filter* f, int32_t n, const vszip_frame* clip, const vszip_frame* clip2, vszip_frame* dst) {
    if (!f || f->kind != F_ADAPTIVEBINARIZE) { set_error("AdaptiveBinarize: bad filter handle"); return -1; }
    if (!clip || !clip2 || !dst) { set_error("AdaptiveBinarize: clip, clip2 and dst frames are required"); return -1; }
    DeviceCtx* d = route(n, "AdaptiveBinarize");
    if (!d) return -1;
    SlotGuard g(d);
    Slot* s = g.s;
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    const size_t bytes = f->layout.frame_stride;
    if (slot_reserve(d, s, 0, bytes) || slot_reserve(d, s, 1, bytes) || slot_reserve(d, s, 2, bytes)) return -1;
    if (stage_in(s, 0, f->layout, clip, f->process) || stage_in(s, 1, f->layout, clip2, f->process)) return -1;
    int rc = run_adaptivebinarize(f->layout, s->dev[0], 0, s->dev[1], 0, s->dev[2], 0, 1, f->ab_c, s->stream);
    if (rc) return rc;
    bool direct[3];
    if (stage_out_begin(s, f->layout, dst, f->process, direct)) return -1;
    VSZ_CUDA(cudaStreamSynchronize(s->stream));
    stage_out_finish(s, f->layout, dst, f->process, direct);
    return 0;
}

int vszip_adaptivebinarize_device(const vszip_filter* f, const vszip_dev_clip* clip, const vszip_dev_clip* clip2, vszip_dev_clip* dst,
                                  int32_t first, int32_t count, void* stream) {
    static const char* name = "AdaptiveBinarize";
    if (!f || f->kind != F_ADAPTIVEBINARIZE) { set_error("AdaptiveBinarize: bad filter handle"); return -1; }
    if (!same_clip_shape(clip, f, name) || !same_clip_shape(clip2, f, name) || !same_clip_shape(dst, f, name)) return -1;
    if (!range_ok(clip, first, count, name) || !range_ok(clip2, first, count, name) || !range_ok(dst, first, count, name)) return -1;
    if (clip->device_index != dst->device_index || clip2->device_index != dst->device_index) { set_error("AdaptiveBinarize: clips live on different devices"); return -1; }
    DeviceCtx* d = device_ctx(dst->device_index);
    if (!d) { set_error("AdaptiveBinarize: library not initialised"); return -1; }
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->batch_stream;
    const size_t fs = dst->layout.frame_stride, off = (size_t)first * fs;
    return run_adaptivebinarize(f->layout, clip->base + off, fs, clip2->base + off, fs, dst->base + off, fs, count, f->ab_c, st);
}

// =========================================================================== fused chains
struct vszip_chain {
    std::vector<const vszip_filter*> fl;
    vsz::FrameLayout layout;
    bool written[3];
    int npixel;
};

vszip_chain* vszip_chain_create(const vszip_filter* const* filters, int32_t count) {
    if (!filters || count < 1 || count > 16) { set_error("chain: between 1 and 16 filters are required"); return nullptr; }
    const vszip_filter* f0 = filters[0];
    for (int i = 0; i < count; ++i) {
        const vszip_filter* f = filters[i];
        if (!f) { set_error("chain: filter %d is NULL", i); return nullptr; }
        // LimitFilter (without ref) and AdaptiveBinarize take the chain's SOURCE frame as their second clip ("diamond" elements)
        const bool diamond = (f->kind == F_LIMITFILTER && !f->lf_has_ref) || f->kind == F_ADAPTIVEBINARIZE;
        if (f->has_ref && !diamond) { set_error("chain: filter %d takes a second clip (ref/clipb); only single-input filters can be fused", i); return nullptr; }
        const vszip_video_info &a = f0->vi, &b = f->vi;
        if (a.width != b.width || a.height != b.height || a.color_family != b.color_family || a.sample_type != b.sample_type ||
            a.bits_per_sample != b.bits_per_sample || a.sub_sampling_w != b.sub_sampling_w || a.sub_sampling_h != b.sub_sampling_h ||
            a.num_planes != b.num_planes) {
            set_error("chain: filter %d was created for a different video format than filter 0", i);
            return nullptr;
        }
    }
    vszip_chain* c = new vszip_chain();
    c->fl.assign(filters, filters + count);
    c->layout = f0->layout;
    c->npixel = 0;
    for (int p = 0; p < 3; ++p) c->written[p] = false;
    for (const vszip_filter* f : c->fl) {
        if (f->kind == F_BOXBLUR || f->kind == F_BILATERAL || f->kind == F_LIMITER || f->kind == F_LIMITFILTER || f->kind == F_ADAPTIVEBINARIZE) {
            ++c->npixel;
            for (int p = 0; p < 3; ++p) c->written[p] = c->written[p] || f->process[p];
        }
    }
    return c;
}

void vszip_chain_free(vszip_chain* c) { delete c; }

int vszip_chain_planes(const vszip_chain* c, int32_t written[3]) {
    if (!c) { set_error("chain: bad handle"); return -1; }
    for (int p = 0; p < 3; ++p) written[p] = c->written[p] ? 1 : 0;
    return 0;
}

int vszip_chain_get_frame(const vszip_chain* c, int32_t n, const vszip_frame* src, vszip_frame* dst, void* const* props_out) {
    if (!c) { set_error("chain: bad handle"); return -1; }
    if (c->npixel > 0 && !dst) { set_error("chain: dst is required when the chain holds a pixel filter"); return -1; }
    DeviceCtx* d = route(n, "chain");
    if (!d) return -1;
    SlotGuard g(d);
    Slot* s = g.s;
    VSZ_CUDA(cudaSetDevice(d->ordinal));
    const FrameLayout& l = c->layout;
    const size_t bytes = l.frame_stride;
    if (slot_reserve(d, s, 0, bytes) || slot_reserve(d, s, 2, bytes) || (c->npixel > 1 && slot_reserve(d, s, 1, bytes))) return -1;
    bool all[3] = {l.nplanes > 0, l.nplanes > 1, l.nplanes > 2};
    if (stage_in(s, 0, l, src, all)) return -1;  // later filters may read planes that earlier ones do not process
    const int dev_index = device_index_of(d);
    int cur = 0, pixel_seen = 0, stats_seen = 0;
    for (size_t i = 0; i < c->fl.size(); ++i) {
        const vszip_filter* f = c->fl[i];
        if (f->kind == F_BOXBLUR || f->kind == F_BILATERAL || f->kind == F_LIMITER || f->kind == F_LIMITFILTER || f->kind == F_ADAPTIVEBINARIZE) {
            // the last pixel filter must land in dev[2] (the D2H source), the ones before alternate 1 / 2;
            // dev[0] is never written, so the source frame stays available to the diamond elements
            const int nxt = ((c->npixel - 1 - pixel_seen) % 2 == 0) ? 2 : 1;
            ++pixel_seen;
            for (int p = 0; p < l.nplanes; ++p) {  // planes this filter passes through
                if (f->process[p]) continue;
                VSZ_CUDA(cudaMemcpyAsync(s->dev[nxt] + l.pl[p].offset, s->dev[cur] + l.pl[p].offset, (size_t)l.pl[p].pitch * l.pl[p].h,
                                         cudaMemcpyDeviceToDevice, s->stream));
            }
            int rc;
            if (f->kind == F_BOXBLUR) {
                rc = run_boxblur(l, f->process, s->dev[cur], 0, s->dev[nxt], 0, 1, (int)f->hradius, f->hpasses, (int)f->vradius, f->vpasses, s->stream);
            } else if (f->kind == F_LIMITER) {
                rc = run_limiter(l, f->process, s->dev[cur], 0, s->dev[nxt], 0, 1, f->lim_lo, f->lim_hi, s->stream);
            } else if (f->kind == F_LIMITFILTER) {       // flt = the chain's current value, src = the chain's source frame
                rc = run_limitfilter(l, f->process, s->dev[cur], 0, s->dev[0], 0, nullptr, 0, s->dev[nxt], 0, 1, f->lf_dark, f->lf_bright, f->lf_elast, s->stream);
            } else if (f->kind == F_ADAPTIVEBINARIZE) {  // clip = the chain's source frame, clip2 = the chain's current value
                rc = run_adaptivebinarize(l, s->dev[0], 0, s->dev[cur], 0, s->dev[nxt], 0, 1, f->ab_c, s->stream);
            } else {
                if (bilateral_upload(f, dev_index)) return -1;
                rc = bilateral_run(f, dev_index, s->dev[cur], 0, nullptr, 0, s->dev[nxt], 0, 1, s->stream);
            }
            if (rc) return rc;
            cur = nxt;
        } else {
            if (!props_out || !props_out[i]) { set_error("chain: props_out[%d] is required for a PlaneMinMax/PlaneAverage element", (int)i); return -1; }
            const int np = processed_planes(f);
            if (np == 0) { ++stats_seen; continue; }
            const size_t sb = stats_scratch_bytes(1, np);
            AsyncScratch stats_mem;
            VSZ_CUDA(stats_mem.alloc(sb, s->stream));
            char* stats_scratch = stats_mem.p;
            StatsRaw* raw_dev = (StatsRaw*)s->dev_small + (size_t)stats_seen * 3;
            int rc;
            if (f->kind == F_PLANEMINMAX) {
                rc = run_planeminmax(l, f->process, s->dev[cur], 0, nullptr, 0, 1, f->no_thr, f->minthr, f->maxthr, f->hist_size, stats_scratch,
                                     raw_dev, s->stream);
            } else {
                const int32_t* xi; const float* xf;
                if (average_upload(f, dev_index, &xi, &xf)) return -1;
                rc = run_planeaverage(l, f->process, s->dev[cur], 0, nullptr, 0, 1, f->exclude_i.data(), f->exclude_f.data(),
                                      (int)f->exclude_i.size(), xi, xf, stats_scratch, raw_dev, s->stream);
            }
            if (rc) return rc;
            VSZ_CUDA(cudaMemcpyAsync((StatsRaw*)s->pin_small + (size_t)stats_seen * 3, raw_dev, sizeof(StatsRaw) * np, cudaMemcpyDeviceToHost, s->stream));
            ++stats_seen;
        }
    }
    bool direct[3] = {false, false, false};
    if (c->npixel > 0 && stage_out_begin(s, l, dst, c->written, direct)) return -1;
    VSZ_CUDA(cudaStreamSynchronize(s->stream));
    if (c->npixel > 0) stage_out_finish(s, l, dst, c->written, direct);
    stats_seen = 0;
    for (size_t i = 0; i < c->fl.size(); ++i) {
        const vszip_filter* f = c->fl[i];
        if (f->kind != F_PLANEMINMAX && f->kind != F_PLANEAVERAGE) continue;
        const StatsRaw* raw = (const StatsRaw*)s->pin_small + (size_t)stats_seen * 3;
        if (processed_planes(f) == 0) {
            if (f->kind == F_PLANEMINMAX) ((vszip_minmax_props*)props_out[i])->count = 0;
            else ((vszip_average_props*)props_out[i])->count = 0;
        } else if (f->kind == F_PLANEMINMAX) {
            minmax_finalize(f, raw, (vszip_minmax_props*)props_out[i]);
        } else {
            average_finalize(f, raw, (vszip_average_props*)props_out[i]);
        }
        ++stats_seen;
    }
    return 0;
}

}  // extern "C"
