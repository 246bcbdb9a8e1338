// Input ingest and sample formatting either side of the denoiser path.
//   dmu_ingest_u8      decoded image bytes -> ToTensor -> Normalize (-> q_sample) in one pass, fp32 NCHW out
//   dmu_image_grid_u8  fp32 samples -> torchvision make_grid -> save_image's 8-bit quantisation, HWC bytes out
// Both are byte/elementwise work bounded by launch latency at the reference's sizes (a 128 x 3 x 32 x 32 batch is
// 393 KB of bytes in, 1.5 MB of floats out); the point is that the host never touches pixels and the host->device
// copy carries 1 byte per value instead of 4.  Arithmetic uses _rn intrinsics operation by operation so the results are
// bit-identical to the torch expressions they replace.
#include "common.cuh"

namespace dmu {

static inline int pipe_grid(int64_t items, int threads) {
    int64_t blocks = (items + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// One thread per 4 consecutive elements of the [B, C, H*W] output (hw % 4 == 0) or per element otherwise.
template <bool VEC>
__global__ void __launch_bounds__(256) ingest_u8_kernel(const uint8_t* __restrict__ img, int hwc, const float* __restrict__ mean,
                                                        const float* __restrict__ stdv, const float* __restrict__ noise,
                                                        const int64_t* __restrict__ t, const float* __restrict__ acp,
                                                        float* __restrict__ x0_out, float* __restrict__ xt_out, int64_t batch,
                                                        int channels, int64_t hw) {
    constexpr int V = VEC ? 4 : 1;
    const int64_t plane = hw, image = (int64_t)channels * hw;
    const int64_t total = batch * image / V;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g * V;              // NCHW element index of the first value
        const int64_t b = i / image;
        const int64_t r = i - b * image;
        const int c = (int)(r / plane);
        const int64_t p = r - (int64_t)c * plane;
        float v[V];
        if (hwc) {
            const uint8_t* src = img + b * image + p * channels + c;
#pragma unroll
            for (int k = 0; k < V; ++k) v[k] = (float)src[(int64_t)k * channels];
        } else {
            if constexpr (VEC) {
                const uchar4 u = *reinterpret_cast<const uchar4*>(img + i);
                v[0] = (float)u.x; v[1] = (float)u.y; v[2] = (float)u.z; v[3] = (float)u.w;
            } else {
                v[0] = (float)img[i];
            }
        }
        const float m = mean ? mean[c] : 0.f, s = stdv ? stdv[c] : 1.f;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            v[k] = __fdiv_rn(v[k], 255.f);                                    // ToTensor: .to(float32).div(255)
            if (mean) v[k] = __fsub_rn(v[k], m);                              // Normalize: .sub_(mean)
            if (stdv) v[k] = __fdiv_rn(v[k], s);                              //            .div_(std)
        }
        if (x0_out) {
            if constexpr (VEC) *reinterpret_cast<float4*>(x0_out + i) = make_float4(v[0], v[1], v[2], v[3]);
            else x0_out[i] = v[0];
        }
        if (xt_out) {
            const float a = acp[t[b]];
            const float ca = __fsqrt_rn(a), cs = __fsqrt_rn(__fsub_rn(1.f, a));   // models/ddpm.py:292-295
            float n[V];
            if constexpr (VEC) {
                const float4 q = *reinterpret_cast<const float4*>(noise + i);
                n[0] = q.x; n[1] = q.y; n[2] = q.z; n[3] = q.w;
            } else {
                n[0] = noise[i];
            }
#pragma unroll
            for (int k = 0; k < V; ++k) v[k] = __fadd_rn(__fmul_rn(ca, v[k]), __fmul_rn(cs, n[k]));
            if constexpr (VEC) *reinterpret_cast<float4*>(xt_out + i) = make_float4(v[0], v[1], v[2], v[3]);
            else xt_out[i] = v[0];
        }
    }
}

struct GridGeom {
    int64_t n, period, stride_mod, stride_div;
    int c, h, w, xmaps, ymaps, pad, cg;
    int64_t hg, wg;
};

// torchvision.utils.save_image: grid.mul(255).add_(0.5).clamp_(0, 255).to(uint8)
__device__ __forceinline__ uint8_t quantise(float v) {
    float q = __fadd_rn(__fmul_rn(v, 255.f), 0.5f);
    q = fminf(fmaxf(q, 0.f), 255.f);
    return (uint8_t)q;
}

// One thread per grid pixel: which cell, which image, then C plane reads and Cg byte writes.
// NORM: make_grid(normalize=True, value_range=(lo, hi)) first maps every image value to (clamp(v, lo, hi) - lo) / den,
// den = max(hi - lo, 1e-5) (torchvision's norm_ip); the padding value is not normalised.
template <bool NORM>
__global__ void __launch_bounds__(256) image_grid_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, GridGeom G,
                                                            float pad_value, float lo, float hi, float den) {
    const int64_t total = G.hg * G.wg;
    const uint8_t padq = quantise(pad_value);
    const int cell_h = G.h + G.pad, cell_w = G.w + G.pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gy = i / G.wg, gx = i - gy * G.wg;
        // image k occupies rows [cy*cell_h + pad, (cy+1)*cell_h) and the matching columns; everything else is padding
        const int64_t cy = gy / cell_h, cx = gx / cell_w;
        const int iy = (int)(gy - cy * cell_h) - G.pad, ix = (int)(gx - cx * cell_w) - G.pad;
        const int64_t k = cy * G.xmaps + cx;
        uint8_t* dst = out + i * G.cg;
        if (iy < 0 || ix < 0 || cy >= G.ymaps || cx >= G.xmaps || k >= G.n) {
            for (int c = 0; c < G.cg; ++c) dst[c] = padq;
            continue;
        }
        const float* src = x + (k % G.period) * G.stride_mod + (k / G.period) * G.stride_div + (int64_t)iy * G.w + ix;
        for (int c = 0; c < G.cg; ++c) {
            float v = src[(int64_t)(G.c == 1 ? 0 : c) * G.h * G.w];
            if constexpr (NORM) v = __fdiv_rn(__fsub_rn(fminf(fmaxf(v, lo), hi), lo), den);
            dst[c] = quantise(v);
        }
    }
}

static bool grid_geometry(int64_t n, int c, int h, int w, int nrow, int pad, GridGeom& G) {
    if (n < 1 || c < 1 || h < 1 || w < 1 || nrow < 1 || pad < 0) return false;
    G.n = n; G.c = c; G.h = h; G.w = w;
    G.cg = c == 1 ? 3 : c;                       // single-channel images are replicated to 3 channels
    if (n == 1) {                                // make_grid returns the lone image itself: no border
        G.pad = 0; G.xmaps = 1; G.ymaps = 1; G.hg = h; G.wg = w;
        return true;
    }
    G.pad = pad;
    G.xmaps = (int)(n < nrow ? n : nrow);
    G.ymaps = (int)((n + G.xmaps - 1) / G.xmaps);
    G.hg = (int64_t)(h + pad) * G.ymaps + pad;
    G.wg = (int64_t)(w + pad) * G.xmaps + pad;
    return true;
}

}  // namespace dmu

using namespace dmu;

extern "C" {

int dmu_ingest_u8(const uint8_t* img, int32_t hwc, const float* mean, const float* stdv, const float* noise, const int64_t* t,
                  const float* acp, float* x0_out, float* xt_out, int64_t batch, int32_t channels, int64_t hw,
                  dmu_stream_t stream) {
    DMU_REQUIRE(batch >= 0 && channels >= 0 && hw >= 0, "dmu_ingest_u8: negative size");
    if (batch == 0 || channels == 0 || hw == 0) return 0;   // empty batch: nothing to do (pointers may be NULL)
    DMU_REQUIRE(img, "dmu_ingest_u8: null image pointer");
    DMU_REQUIRE(x0_out || xt_out, "dmu_ingest_u8: no output requested");
    DMU_REQUIRE(!xt_out || (noise && t && acp), "dmu_ingest_u8: xt_out needs noise, t and alphas_cumprod");
    auto al = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
    const bool vec = (hw % 4) == 0 && al(x0_out, 16) && al(xt_out, 16) && al(noise, 16) && (hwc || al(img, 4));
    const int64_t groups = batch * channels * hw / (vec ? 4 : 1);
    const int grid = pipe_grid(groups, 256);
    if (vec)
        ingest_u8_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(img, hwc, mean, stdv, noise, t, acp, x0_out, xt_out, batch,
                                                                   channels, hw);
    else
        ingest_u8_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(img, hwc, mean, stdv, noise, t, acp, x0_out, xt_out, batch,
                                                                    channels, hw);
    return check_launch("dmu_ingest_u8");
}

int dmu_image_grid_shape(int64_t n_images, int32_t channels, int32_t height, int32_t width, int32_t nrow, int32_t padding,
                         int64_t* grid_h, int64_t* grid_w, int32_t* grid_c) {
    GridGeom G;
    DMU_REQUIRE(grid_geometry(n_images, channels, height, width, nrow, padding, G), "dmu_image_grid_shape: bad geometry");
    if (grid_h) *grid_h = G.hg;
    if (grid_w) *grid_w = G.wg;
    if (grid_c) *grid_c = G.cg;
    return 0;
}

int dmu_image_grid_u8(const float* x, int64_t n_images, int64_t period, int64_t stride_mod, int64_t stride_div, int32_t channels,
                      int32_t height, int32_t width, int32_t nrow, int32_t padding, float pad_value, uint8_t* out,
                      dmu_stream_t stream) {
    GridGeom G;
    DMU_REQUIRE(x && out, "dmu_image_grid_u8: null pointer");
    DMU_REQUIRE(grid_geometry(n_images, channels, height, width, nrow, padding, G), "dmu_image_grid_u8: bad geometry");
    DMU_REQUIRE(period >= 1 && stride_mod >= 0 && stride_div >= 0, "dmu_image_grid_u8: bad image addressing");
    G.period = period; G.stride_mod = stride_mod; G.stride_div = stride_div;
    image_grid_u8_kernel<false><<<pipe_grid(G.hg * G.wg, 256), 256, 0, as_stream(stream)>>>(x, out, G, pad_value, 0.f, 0.f, 1.f);
    return check_launch("dmu_image_grid_u8");
}

int dmu_image_grid_range_u8(const float* x, int64_t n_images, int64_t period, int64_t stride_mod, int64_t stride_div,
                            int32_t channels, int32_t height, int32_t width, int32_t nrow, int32_t padding, float pad_value,
                            double range_lo, double range_hi, uint8_t* out, dmu_stream_t stream) {
    GridGeom G;
    DMU_REQUIRE(x && out, "dmu_image_grid_range_u8: null pointer");
    DMU_REQUIRE(grid_geometry(n_images, channels, height, width, nrow, padding, G), "dmu_image_grid_range_u8: bad geometry");
    DMU_REQUIRE(period >= 1 && stride_mod >= 0 && stride_div >= 0, "dmu_image_grid_range_u8: bad image addressing");
    DMU_REQUIRE(range_lo <= range_hi, "dmu_image_grid_range_u8: empty value range");
    G.period = period; G.stride_mod = stride_mod; G.stride_div = stride_div;
    // torchvision: python doubles; max(high - low, 1e-5) is taken in double and each scalar becomes a float operand
    const double span = range_hi - range_lo;
    const float den = (float)(span > 1e-5 ? span : 1e-5);
    image_grid_u8_kernel<true><<<pipe_grid(G.hg * G.wg, 256), 256, 0, as_stream(stream)>>>(x, out, G, pad_value, (float)range_lo,
                                                                                          (float)range_hi, den);
    return check_launch("dmu_image_grid_range_u8");
}

}  // extern "C"
