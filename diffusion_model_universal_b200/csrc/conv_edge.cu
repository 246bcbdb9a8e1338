// The network's 3-channel boundary layers (stem conv 3->C, head conv C->3 and their gradients).
//
// K = 27 (or N = 3) is far too thin for the tensor-core tiles and the generic SIMT GEMM wastes most of its 16-wide
// K steps on them; at 32x32 x batch 128 these five launches were 0.74 ms of a 6.5 ms step.  They are HBM-/issue-bound
// elementwise-like passes over the one wide (C-channel, NHWC) tensor involved, so each gets a direct kernel:
//   narrow_in_kernel   few-channel input (any strides, e.g. NCHW fp32)  -> wide NHWC output     (stem fprop, head dgrad)
//   narrow_out_kernel  wide NHWC input -> few-channel output (any strides)                      (head fprop)
//   narrow_wgrad_kernel  sum_pixels wide[pix, c] * narrow[pix -/+ tap, j]                        (stem wgrad, head wgrad)
// stride 1 only; the transposed gather (dgrad) is the same arithmetic with mirrored taps.
#include "common.cuh"

namespace dmu {
namespace tc {   // conv_edge_tc.cu
int edge_wgrad_tc_supported(const dmu_tensor4* wide, const dmu_tensor4* narrow, int N, int H, int W, int Cw, int Cn);
int edge_wgrad_tc_launch(const dmu_tensor4* wide, const dmu_tensor4* narrow, int N, int H, int W, int Cw, int Cn, int sgn, float* out,
                         int64_t o_c, int64_t o_t, int64_t o_j, float* dbias, float* dbias_n, cudaStream_t stream);
}  // namespace tc
namespace edge {

constexpr int kMaxNarrow = 4;    // channels on the thin side
constexpr int kMaxTapsE = 25;    // up to 5x5

// ------------------------------------------------------------------------------------------------ narrow in -> wide out
// thread = 2 horizontally adjacent output pixels x 8 output channels; weights staged as fp32 [tap*Ck + k][Cj] in smem
template <typename TY>
__global__ void __launch_bounds__(256) narrow_in_kernel(dmu_conv_params P) {
    extern __shared__ float s_w[];   // [K][Cj]
    const int K = P.R * P.S * P.Ck;
    for (int i = threadIdx.x; i < K * P.Cj; i += blockDim.x) {
        const int j = i % P.Cj, k = i / P.Cj;
        const int tap = k / P.Ck, kc = k % P.Ck;
        s_w[i] = ld_as_float(P.w, (int64_t)j * P.w_sn + (int64_t)kc * P.w_sk + (int64_t)tap * P.w_st, P.w_dtype);
    }
    __syncthreads();
    const int groups = P.Cj >> 3;                 // 8-channel groups per pixel
    const int Wp = (P.Wo + 1) >> 1;               // pixel pairs per row
    const int64_t items = (int64_t)P.N * P.Ho * Wp * groups;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(it % groups);
        const int64_t pp = it / groups;
        const int wp = (int)(pp % Wp), ho = (int)((pp / Wp) % P.Ho), n = (int)(pp / ((int64_t)Wp * P.Ho));
        const int wo0 = wp * 2;
        const bool two = wo0 + 1 < P.Wo;
        float acc0[8], acc1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
        for (int r = 0; r < P.R; ++r) {
            const int hi = P.gather == 0 ? ho - P.pad + r : ho + P.pad - r;
            if (hi < 0 || hi >= P.Hi) continue;
            for (int s = 0; s < P.S; ++s) {
                const int wi0 = P.gather == 0 ? wo0 - P.pad + s : wo0 + P.pad - s;
                const int wi1 = wi0 + 1;
                const bool ok0 = wi0 >= 0 && wi0 < P.Wi, ok1 = two && wi1 >= 0 && wi1 < P.Wi;
                const float* wrow = s_w + (size_t)((r * P.S + s) * P.Ck) * P.Cj + g * 8;
                const int64_t base = (int64_t)n * P.x.sn + (int64_t)hi * P.x.sh;
                for (int kc = 0; kc < P.Ck; ++kc) {
                    const float x0 = ok0 ? ld_as_float(P.x.ptr, base + (int64_t)wi0 * P.x.sw + (int64_t)kc * P.x.sc, P.x.dtype) : 0.f;
                    const float x1 = ok1 ? ld_as_float(P.x.ptr, base + (int64_t)wi1 * P.x.sw + (int64_t)kc * P.x.sc, P.x.dtype) : 0.f;
                    const float4 wa = *reinterpret_cast<const float4*>(wrow + (size_t)kc * P.Cj);
                    const float4 wb = *reinterpret_cast<const float4*>(wrow + (size_t)kc * P.Cj + 4);
                    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) { acc0[i] = fmaf(x0, wv[i], acc0[i]); acc1[i] = fmaf(x1, wv[i], acc1[i]); }
                }
            }
        }
        const int j0 = g * 8;
        float add[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) add[i] = (P.bias ? P.bias[j0 + i] : 0.f) + (P.temb ? P.temb[(int64_t)n * P.temb_pitch + j0 + i] : 0.f);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (q == 1 && !two) break;
            float* acc = q == 0 ? acc0 : acc1;
            const int wo = wo0 + q;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += add[i];
            if (P.res.ptr) {
                float rr[8];
                load_vec<TY>(reinterpret_cast<const TY*>(P.res.ptr) + (int64_t)n * P.res.sn + (int64_t)ho * P.res.sh + (int64_t)wo * P.res.sw + j0, rr);
                if constexpr (sizeof(TY) == 4)
                    load_vec<TY>(reinterpret_cast<const TY*>(P.res.ptr) + (int64_t)n * P.res.sn + (int64_t)ho * P.res.sh + (int64_t)wo * P.res.sw + j0 + 4, rr + 4);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += rr[i];
            }
            TY* yp = reinterpret_cast<TY*>(P.y.ptr) + (int64_t)n * P.y.sn + (int64_t)ho * P.y.sh + (int64_t)wo * P.y.sw + j0;
            store_vec<TY>(yp, acc);
            if constexpr (sizeof(TY) == 4) store_vec<TY>(yp + 4, acc + 4);
        }
    }
}

// ------------------------------------------------------------------------------------------------ narrow in -> wide out, v2
// 3x3 / pad 1 / stride 1 only (stem fprop, head dgrad).  CTA = one band of rows of one image: the few-channel input band
// (+halo) and the filter bank are staged in shared memory as fp32, thread = 4 consecutive output pixels x 8 output channels
// (a 6-wide input window per (filter row, channel) serves 3 taps x 4 pixels; weights are 2 LDS.128 per tap), so the inner loop
// is ~30 instructions per output instead of a latency-bound global gather.  bf16 or fp32 NHWC output, + bias.
constexpr int kBandV2 = 4;
template <typename TY>
__global__ void __launch_bounds__(256) narrow_in_v2_kernel(dmu_conv_params P) {
    extern __shared__ float smem_f[];
    const int K = 9 * P.Ck;
    const int PW = P.Wi + 2;
    float* s_w = smem_f;                          // [K][Cj]
    float* s_x = smem_f + (size_t)K * P.Cj;       // [kBandV2 + 2][Ck][PW]
    for (int i = threadIdx.x; i < K * P.Cj; i += blockDim.x) {
        const int j = i % P.Cj, k = i / P.Cj;
        const int tap = k / P.Ck, kc = k % P.Ck;
        s_w[i] = ld_as_float(P.w, (int64_t)j * P.w_sn + (int64_t)kc * P.w_sk + (int64_t)tap * P.w_st, P.w_dtype);
    }
    const int groups = P.Cj >> 3;
    const int g = threadIdx.x % groups, pl = threadIdx.x / groups, lanes = blockDim.x / groups;
    const int QW = (P.Wo + 3) >> 2;
    const int bands_per_img = (P.Ho + kBandV2 - 1) / kBandV2;
    const int nbands = P.N * bands_per_img;
    float bj[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bj[i] = P.bias ? P.bias[g * 8 + i] : 0.f;
    for (int band = blockIdx.x; band < nbands; band += gridDim.x) {
        const int n = band / bands_per_img, h0 = (band % bands_per_img) * kBandV2;
        const int rows = min(kBandV2, P.Ho - h0);
        __syncthreads();
        for (int i = threadIdx.x; i < (kBandV2 + 2) * P.Ck * PW; i += blockDim.x) {
            const int pw = i % PW, kc = (i / PW) % P.Ck, rr = i / (PW * P.Ck);
            const int hh = h0 + rr - 1, ww = pw - 1;
            float v = 0.f;
            if (hh >= 0 && hh < P.Hi && ww >= 0 && ww < P.Wi)
                v = ld_as_float(P.x.ptr, (int64_t)n * P.x.sn + (int64_t)hh * P.x.sh + (int64_t)ww * P.x.sw + (int64_t)kc * P.x.sc, P.x.dtype);
            s_x[i] = v;
        }
        __syncthreads();
        if (pl >= lanes) continue;
        for (int q = pl; q < rows * QW; q += lanes) {
            const int hl = q / QW, w0 = (q % QW) * 4;
            float acc[4][8];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[p][i] = bj[i];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int dr = P.gather ? 2 - r : r;
                for (int kc = 0; kc < P.Ck; ++kc) {
                    const float* xr = s_x + ((size_t)(hl + dr) * P.Ck + kc) * PW + w0;
                    float xw[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) xw[i] = (w0 + i < PW) ? xr[i] : 0.f;
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int ds = P.gather ? 2 - t : t;
                        const float* wr = s_w + (size_t)((r * 3 + t) * P.Ck + kc) * P.Cj + g * 8;
                        const float4 wa = *reinterpret_cast<const float4*>(wr), wb = *reinterpret_cast<const float4*>(wr + 4);
                        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                        for (int p = 0; p < 4; ++p)
#pragma unroll
                            for (int i = 0; i < 8; ++i) acc[p][i] = fmaf(xw[p + ds], wv[i], acc[p][i]);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (w0 + p >= P.Wo) break;
                TY* yp = reinterpret_cast<TY*>(P.y.ptr) + (int64_t)n * P.y.sn + (int64_t)(h0 + hl) * P.y.sh + (int64_t)(w0 + p) * P.y.sw + g * 8;
                store_vec<TY>(yp, acc[p]);
                if constexpr (sizeof(TY) == 4) store_vec<TY>(yp + 4, acc[p] + 4);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ wide in -> narrow out
// 8 threads per output pixel, each owning 8 of every 64 input channels (one 16-byte load per tap and 64-channel chunk, all
// independent), then a 3-step shuffle reduction; weights fp32 [j][tap][k] in smem.
template <typename TX>
__global__ void __launch_bounds__(256) narrow_out_kernel(dmu_conv_params P) {
    extern __shared__ float s_w[];   // [Cj][taps*Ck]
    const int taps = P.R * P.S, K = taps * P.Ck;
    for (int i = threadIdx.x; i < K * P.Cj; i += blockDim.x) {
        const int j = i / K, k = i % K;
        const int tap = k / P.Ck, kc = k % P.Ck;
        s_w[i] = ld_as_float(P.w, (int64_t)j * P.w_sn + (int64_t)kc * P.w_sk + (int64_t)tap * P.w_st, P.w_dtype);
    }
    __syncthreads();
    const int sub = threadIdx.x & 7;
    const int64_t M = (int64_t)P.N * P.Ho * P.Wo;
    const int64_t m_stride = (int64_t)gridDim.x * (blockDim.x >> 3);
    for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); m < ((M + 31) / 32) * 32; m += m_stride) {
        const bool live = m < M;
        const int64_t mm = live ? m : 0;
        const int wo = (int)(mm % P.Wo), ho = (int)((mm / P.Wo) % P.Ho), n = (int)(mm / ((int64_t)P.Wo * P.Ho));
        float acc[kMaxNarrow] = {0.f, 0.f, 0.f, 0.f};
        if (live) {
            for (int r = 0; r < P.R; ++r) {
                const int hi = P.gather == 0 ? ho - P.pad + r : ho + P.pad - r;
                if (hi < 0 || hi >= P.Hi) continue;
                for (int s = 0; s < P.S; ++s) {
                    const int wi = P.gather == 0 ? wo - P.pad + s : wo + P.pad - s;
                    if (wi < 0 || wi >= P.Wi) continue;
                    const TX* xp = reinterpret_cast<const TX*>(P.x.ptr) + (int64_t)n * P.x.sn + (int64_t)hi * P.x.sh + (int64_t)wi * P.x.sw;
                    const int kbase = (r * P.S + s) * P.Ck;
                    for (int c0 = sub * 8; c0 < P.Ck; c0 += 64) {
                        float xv[8];
                        load_vec<TX>(xp + c0, xv);
                        if constexpr (sizeof(TX) == 4) load_vec<TX>(xp + c0 + 4, xv + 4);
#pragma unroll
                        for (int j = 0; j < kMaxNarrow; ++j) {
                            if (j < P.Cj) {
                                const float4 wa = *reinterpret_cast<const float4*>(s_w + (size_t)j * K + kbase + c0);
                                const float4 wb = *reinterpret_cast<const float4*>(s_w + (size_t)j * K + kbase + c0 + 4);
                                acc[j] = fmaf(xv[0], wa.x, acc[j]); acc[j] = fmaf(xv[1], wa.y, acc[j]);
                                acc[j] = fmaf(xv[2], wa.z, acc[j]); acc[j] = fmaf(xv[3], wa.w, acc[j]);
                                acc[j] = fmaf(xv[4], wb.x, acc[j]); acc[j] = fmaf(xv[5], wb.y, acc[j]);
                                acc[j] = fmaf(xv[6], wb.z, acc[j]); acc[j] = fmaf(xv[7], wb.w, acc[j]);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kMaxNarrow; ++j) {
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
        }
        if (live && sub < P.Cj) {
            const int j = sub;
            float v = acc[0];
#pragma unroll
            for (int q = 1; q < kMaxNarrow; ++q) v = (j == q) ? acc[q] : v;
            if (P.bias) v += P.bias[j];
            if (P.temb) v += P.temb[(int64_t)n * P.temb_pitch + j];
            const int64_t po = (int64_t)n * P.res.sn + (int64_t)ho * P.res.sh + (int64_t)wo * P.res.sw + (int64_t)j * P.res.sc;
            if (P.res.ptr) v += ld_as_float(P.res.ptr, po, P.res.dtype);
            st_from_float(P.y.ptr, (int64_t)n * P.y.sn + (int64_t)ho * P.y.sh + (int64_t)wo * P.y.sw + (int64_t)j * P.y.sc, P.y.dtype, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------ narrow x wide wgrad
// out[c][tap][j] (strides o_c, o_t, o_j) += sum_pix wide[pix, c] * narrow[pix + sgn*(tap - pad), j]
// CTA = one band of kBand rows of one image; the narrow tensor's halo patch sits in smem.  Thread = (8-channel group g,
// filter row r) x pixel lane: S*Cn*8 accumulators in registers, reduced over the pixel lanes through smem at the end of the
// CTA's bands, then one atomic per output.
struct NarrowWgradArgs {
    dmu_tensor4 wide, narrow;
    float* out; int64_t o_c, o_t, o_j;
    float* dbias;        // optional: [Cw] += sum_pix wide  (sgn < 0 form: bias of the conv whose output gradient is `wide`)
    float* dbias_n;      // optional: [Cn] += sum_pix narrow
    int N, H, W, Cw, Cn, R, S, pad, sgn;
};
constexpr int kBand = 4;

template <typename TW, int S, int CN>
__global__ void __launch_bounds__(256) narrow_wgrad_kernel(NarrowWgradArgs A) {
    extern __shared__ float smem[];
    const int groups = A.Cw >> 3;
    const int roles = groups * A.R;                       // (g, r) pairs
    const int lanes = blockDim.x / roles;                 // pixel lanes
    const int role = threadIdx.x % roles, pl = threadIdx.x / roles;
    const int g = role % groups, r = role / groups;
    const bool active = pl < lanes;
    const int PW = A.W + 2 * A.pad;                       // halo patch width
    const int PH = kBand + 2 * A.pad;
    float* s_n = smem;                                    // [PH][PW][CN]
    float acc[S * CN][8];
    float bsum[8];
#pragma unroll
    for (int i = 0; i < S * CN; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) bsum[k] = 0.f;
    float nsum = 0.f;                                     // threads 0..CN-1: running sum of the narrow tensor (dbias_n)

    const int bands_per_img = (A.H + kBand - 1) / kBand;
    const int nbands = A.N * bands_per_img;
    for (int band = blockIdx.x; band < nbands; band += gridDim.x) {
        const int n = band / bands_per_img, h0 = (band % bands_per_img) * kBand;
        const int rows = min(kBand, A.H - h0);
        __syncthreads();
        for (int i = threadIdx.x; i < PH * PW * CN; i += blockDim.x) {
            const int j = i % CN, pw = (i / CN) % PW, phh = i / (CN * PW);
            const int hh = h0 + phh - A.pad, ww = pw - A.pad;
            float v = 0.f;
            if (hh >= 0 && hh < A.H && ww >= 0 && ww < A.W)
                v = ld_as_float(A.narrow.ptr, (int64_t)n * A.narrow.sn + (int64_t)hh * A.narrow.sh + (int64_t)ww * A.narrow.sw + (int64_t)j * A.narrow.sc, A.narrow.dtype);
            s_n[i] = v;
        }
        __syncthreads();
        if (A.dbias_n && threadIdx.x < CN) {
            for (int phh = A.pad; phh < A.pad + rows; ++phh)
                for (int pw = A.pad; pw < A.pad + A.W; ++pw) nsum += s_n[(phh * PW + pw) * CN + threadIdx.x];
        }
        if (!active) continue;
        const int npix = rows * A.W;
        const TW* wb = reinterpret_cast<const TW*>(A.wide.ptr) + (int64_t)n * A.wide.sn + g * 8;
        for (int p = pl; p < npix; p += lanes) {
            const int hl = p / A.W, w = p % A.W;
            float xv[8];
            const TW* xp = wb + (int64_t)(h0 + hl) * A.wide.sh + (int64_t)w * A.wide.sw;
            load_vec<TW>(xp, xv);
            if constexpr (sizeof(TW) == 4) load_vec<TW>(xp + 4, xv + 4);
            if (r == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) bsum[k] += xv[k];
            }
            // narrow pixel for tap (r, s): (h + sgn*(r-pad), w + sgn*(s-pad)) -> patch coords (+pad)
            const int ph_ = hl + A.pad + A.sgn * (r - A.pad);
            const float* nrow = s_n + (size_t)(ph_ * PW) * CN;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int pw_ = w + A.pad + A.sgn * (s - A.pad);
#pragma unroll
                for (int j = 0; j < CN; ++j) {
                    const float nv = nrow[pw_ * CN + j];
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[s * CN + j][k] = fmaf(xv[k], nv, acc[s * CN + j][k]);
                }
            }
        }
    }
    // ---- reduce over pixel lanes: smem [lanes][roles][S*CN*8]
    __syncthreads();
    float* s_red = smem;
    const int per = S * CN * 8;
    if (active) {
#pragma unroll
        for (int i = 0; i < S * CN; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) s_red[((size_t)pl * roles + role) * per + i * 8 + k] = acc[i][k];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < roles * per; o += blockDim.x) {
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += s_red[(size_t)l * roles * per + o];
        const int ro = o / per, idx = o % per;
        const int gg = ro % groups, rr = ro / groups;
        const int s = (idx / 8) / CN, j = (idx / 8) % CN, k = idx % 8;
        const int c = gg * 8 + k, tap = rr * S + s;
        atomicAdd(&A.out[(int64_t)c * A.o_c + (int64_t)tap * A.o_t + (int64_t)j * A.o_j], t);
    }
    if (A.dbias) {
        __syncthreads();
        if (active && r == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) s_red[(size_t)pl * A.Cw + g * 8 + k] = bsum[k];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < A.Cw; c += blockDim.x) {
            float t = 0.f;
            for (int l = 0; l < lanes; ++l) t += s_red[(size_t)l * A.Cw + c];
            atomicAdd(&A.dbias[c], t);
        }
    }
    if (A.dbias_n && threadIdx.x < CN) atomicAdd(&A.dbias_n[threadIdx.x], nsum);
}

static size_t narrow_wgrad_smem(const NarrowWgradArgs& A, int threads) {
    const int roles = (A.Cw >> 3) * A.R, lanes = threads / roles;
    const size_t patch = (size_t)(kBand + 2 * A.pad) * (A.W + 2 * A.pad) * A.Cn;
    const size_t red = (size_t)lanes * roles * A.S * A.Cn * 8;
    const size_t bias = (size_t)lanes * A.Cw;
    size_t m = patch > red ? patch : red;
    if (bias > m) m = bias;
    return m * sizeof(float);
}

}  // namespace edge
}  // namespace dmu

using namespace dmu;

extern "C" {

int dmu_conv2d_edge_supported(const dmu_conv_params* p) {
    if (!p || p->stride != 1 || p->R * p->S > edge::kMaxTapsE) return 0;
    const bool y_wide = p->y.sc == 1 && p->Cj % 8 == 0 && p->Cj >= 16 && p->y.sw % 8 == 0 && p->y.sh % 8 == 0 && p->y.sn % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(p->y.ptr) & 15) == 0;
    if (p->Ck <= edge::kMaxNarrow && y_wide) {   // narrow in -> wide out
        if (p->res.ptr && !(p->res.sc == 1 && p->res.dtype == p->y.dtype && p->res.sw % 8 == 0 && p->res.sh % 8 == 0 && p->res.sn % 8 == 0 &&
                            (reinterpret_cast<uintptr_t>(p->res.ptr) & 15) == 0))
            return 0;
        return (size_t)p->R * p->S * p->Ck * p->Cj * sizeof(float) <= 96 * 1024;
    }
    const bool x_wide = p->x.sc == 1 && p->Ck % 64 == 0 && p->x.sw % 8 == 0 && p->x.sh % 8 == 0 && p->x.sn % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(p->x.ptr) & 15) == 0;
    if (p->Cj <= edge::kMaxNarrow && x_wide)     // wide in -> narrow out
        return (size_t)p->R * p->S * p->Ck * p->Cj * sizeof(float) <= 96 * 1024;
    return 0;
}

int dmu_conv2d_edge(const dmu_conv_params* p, dmu_stream_t stream) {
    const size_t smem = (size_t)p->R * p->S * p->Ck * p->Cj * sizeof(float);
    const int64_t M = (int64_t)p->N * p->Ho * p->Wo;
    if (p->Ck <= edge::kMaxNarrow && p->R == 3 && p->S == 3 && p->pad == 1 && !p->temb && !p->res.ptr && p->Cj <= 256 && p->Hi == p->Ho && p->Wi == p->Wo) {
        const size_t smem2 = ((size_t)9 * p->Ck * p->Cj + (size_t)(edge::kBandV2 + 2) * p->Ck * (p->Wi + 2)) * sizeof(float);
        if (smem2 <= 160 * 1024) {
            const int nbands = p->N * ((p->Ho + edge::kBandV2 - 1) / edge::kBandV2);
            int grid = nbands < 4 * sm_count() ? nbands : 4 * sm_count();
            if (p->y.dtype == DMU_BF16) {
                if (smem2 > 48 * 1024) cudaFuncSetAttribute(edge::narrow_in_v2_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
                edge::narrow_in_v2_kernel<__nv_bfloat16><<<grid, 256, smem2, as_stream(stream)>>>(*p);
            } else {
                if (smem2 > 48 * 1024) cudaFuncSetAttribute(edge::narrow_in_v2_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
                edge::narrow_in_v2_kernel<float><<<grid, 256, smem2, as_stream(stream)>>>(*p);
            }
            return check_launch("dmu_conv2d/narrow_in_v2");
        }
    }
    if (p->Ck <= edge::kMaxNarrow) {
        const int64_t items = (int64_t)p->N * p->Ho * ((p->Wo + 1) / 2) * (p->Cj / 8);
        int grid = (int)((items + 255) / 256);
        if (grid > sm_count() * 8) grid = sm_count() * 8;
        if (p->y.dtype == DMU_BF16) {
            if (smem > 48 * 1024) cudaFuncSetAttribute(edge::narrow_in_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            edge::narrow_in_kernel<__nv_bfloat16><<<grid, 256, smem, as_stream(stream)>>>(*p);
        } else {
            if (smem > 48 * 1024) cudaFuncSetAttribute(edge::narrow_in_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            edge::narrow_in_kernel<float><<<grid, 256, smem, as_stream(stream)>>>(*p);
        }
        return check_launch("dmu_conv2d/narrow_in");
    }
    int grid = (int)((M + 31) / 32);
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    if (p->x.dtype == DMU_BF16) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(edge::narrow_out_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        edge::narrow_out_kernel<__nv_bfloat16><<<grid, 256, smem, as_stream(stream)>>>(*p);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(edge::narrow_out_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        edge::narrow_out_kernel<float><<<grid, 256, smem, as_stream(stream)>>>(*p);
    }
    return check_launch("dmu_conv2d/narrow_out");
}

// dW[a][tap][b] += sum P[pix, a] * Q[pix - pad + tap, b]  with one of P, Q at most 4 channels wide, 3x3, stride 1
int dmu_wgrad_edge_supported(const dmu_wgrad_params* p) {
    if (!p || p->stride != 1 || p->R != 3 || p->S != 3 || p->pad != 1 || p->Hp != p->Hq || p->Wp != p->Wq) return 0;
    const dmu_tensor4& wide = p->Ca <= edge::kMaxNarrow ? p->q : p->p;
    const int Cw = p->Ca <= edge::kMaxNarrow ? p->Cb : p->Ca, Cn = p->Ca <= edge::kMaxNarrow ? p->Ca : p->Cb;
    if (Cn > edge::kMaxNarrow || Cn != 3 || Cw % 8 != 0 || Cw < 8 || (Cw / 8) * p->R > 256) return 0;
    if (wide.sc != 1 || wide.sw % 8 != 0 || wide.sh % 8 != 0 || wide.sn % 8 != 0 || (reinterpret_cast<uintptr_t>(wide.ptr) & 15) != 0) return 0;
    return 1;
}

int dmu_wgrad_edge(const dmu_wgrad_params* p, dmu_stream_t stream) {
    edge::NarrowWgradArgs A;
    const bool p_narrow = p->Ca <= edge::kMaxNarrow;
    A.N = p->N; A.H = p->Hp; A.W = p->Wp; A.R = p->R; A.S = p->S; A.pad = p->pad;
    A.out = p->dw;
    if (!p_narrow) {   // wide = P (unshifted), narrow = Q shifted by +(tap - pad):  out[c=a][tap][j=b]
        A.wide = p->p; A.narrow = p->q; A.Cw = p->Ca; A.Cn = p->Cb; A.sgn = 1;
        A.o_c = p->dw_sa; A.o_j = p->dw_sb; A.o_t = p->dw_st;
        A.dbias = p->dbias; A.dbias_n = nullptr;
    } else {           // wide = Q; substitute pix' = pix + tap - pad: narrow = P shifted by -(tap - pad):  out[c=b][tap][j=a]
        A.wide = p->q; A.narrow = p->p; A.Cw = p->Cb; A.Cn = p->Ca; A.sgn = -1;
        A.o_c = p->dw_sb; A.o_j = p->dw_sa; A.o_t = p->dw_st;
        A.dbias = nullptr; A.dbias_n = p->dbias;
    }
    // opt-in tensor-core form (conv_edge_tc.cu, DMU_EDGE_WGRAD_TC=1): same contract, pixels as the contraction axis
    if (A.R == 3 && A.S == 3 && A.pad == 1 && tc::edge_wgrad_tc_supported(&A.wide, &A.narrow, A.N, A.H, A.W, A.Cw, A.Cn))
        return tc::edge_wgrad_tc_launch(&A.wide, &A.narrow, A.N, A.H, A.W, A.Cw, A.Cn, A.sgn, A.out, A.o_c, A.o_t, A.o_j, A.dbias,
                                        A.dbias_n, as_stream(stream));
    const int threads = 256;
    const size_t smem = edge::narrow_wgrad_smem(A, threads);
    DMU_REQUIRE(smem <= 200 * 1024, "dmu_conv2d_wgrad/edge: tile does not fit shared memory");
    const int nbands = A.N * ((A.H + edge::kBand - 1) / edge::kBand);
    int grid = nbands < 2 * sm_count() ? nbands : 2 * sm_count();
    if (A.wide.dtype == DMU_BF16) {
        cudaFuncSetAttribute(edge::narrow_wgrad_kernel<__nv_bfloat16, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        edge::narrow_wgrad_kernel<__nv_bfloat16, 3, 3><<<grid, threads, smem, as_stream(stream)>>>(A);
    } else {
        cudaFuncSetAttribute(edge::narrow_wgrad_kernel<float, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        edge::narrow_wgrad_kernel<float, 3, 3><<<grid, threads, smem, as_stream(stream)>>>(A);
    }
    return check_launch("dmu_conv2d_wgrad/edge");
}

}  // extern "C"
