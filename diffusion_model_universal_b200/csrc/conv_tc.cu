// tcgen05 / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulators in TMEM).
//
// fprop / dgrad  (dmu_conv2d, impl 2):   D[pixel, j] = sum_{tap, k} X[pixel shifted by tap, k] * W[j][tap][k]
//   * M = 128 output pixels per CTA = one TMA box (64 channels x BW x BH x BN) of the NHWC input per (tap, 64-channel chunk);
//     the box lands in shared memory as 128 rows x 128 bytes with the 128-byte swizzle, which IS the K-major UMMA
//     operand layout: no im2col buffer, padding comes from TMA out-of-bounds zero fill (signed box coordinates).
//   * stride-2 convolutions read one of four parity sub-lattices of the input (four tensor maps with doubled strides);
//     transposed stride-2 convolutions run as four output-parity phases of 2x2 taps (blockIdx.z) - no zero insertion.
//   * N = 64 or 128 output channels per CTA; weights [J][taps][K] come through a 2-D tensor map.
//   * one elected thread issues TMA, one issues tcgen05.mma; an mbarrier ring of kStages connects them; the
//     accumulator is read back with tcgen05.ld and the epilogue fuses + bias + temb[n, j] + residual -> bf16 NHWC.
//
// wgrad (dmu_conv2d_wgrad, impl 2):  dW[a][tap][b] += sum_pixels P[pixel, a] * Q[pixel shifted by tap, b]
//   * the contraction runs over pixels, so both operands are MN-major (channels contiguous): the same TMA boxes
//     (64 channels x 64 pixels) are consumed through MN-major UMMA descriptors; M = two (tap, 64-channel) blocks of Q,
//     N = up to 128 channels of P; split over pixel ranges (blockIdx.z), fp32 atomics into the OIHW gradient.
#include <stdlib.h>
#include <string.h>
#include <type_traits>

#include "tc_common.cuh"

namespace dmu {
namespace tc {

template <int V> using IntC = std::integral_constant<int, V>;

// ------------------------------------------------------------------------------------------------ host helpers
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        cudaGetLastError();
    }
    return fn;
}

int make_map_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems, const uint32_t* box,
                  const char* what) {
    EncodeTiledFn enc = encode_tiled_fn();
    DMU_REQUIRE(enc, "%s: cuTensorMapEncodeTiled is not available from this driver", what);
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstr[i - 1] = strides_elems[i] * 2;
    }
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DMU_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed (CUresult %d; dims %llu %llu %llu %llu, box %u %u %u %u)", what, (int)r,
                (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return 0;
}

static inline int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool nhwc_bf16_ok(const dmu_tensor4& t) {
    return t.dtype == DMU_BF16 && t.sc == 1 && aligned16(t.ptr) && t.sw % 8 == 0 && t.sh % 8 == 0 && t.sn % 8 == 0;
}

constexpr int kMaxTaps = 16;
constexpr int kMaxPhases = 4;

struct Tap { int dh, dw, map, wk; };
struct Phase { int tap0, ntaps, oph, opw, TH, TW; };
struct Maps { CUtensorMap a[4]; CUtensorMap b; };

// pixel-box geometry shared by both kernels: `pix` pixels per box
struct Box { int BN, BH, BW, tiles_n, tiles_h, tiles_w; };
static Box make_box(int N, int TH, int TW, int pix) {
    Box b;
    b.BW = pow2_ceil(TW) < pix ? pow2_ceil(TW) : pix;
    b.BH = pow2_ceil(TH) < pix / b.BW ? pow2_ceil(TH) : pix / b.BW;
    b.BN = pix / (b.BW * b.BH);
    if (b.BN > N) b.BN = N;
    b.tiles_w = (TW + b.BW - 1) / b.BW;
    b.tiles_h = (TH + b.BH - 1) / b.BH;
    b.tiles_n = (N + b.BN - 1) / b.BN;
    return b;
}

// Tap table + tensor maps of a gather-0 ("convolution") read pattern: q = p*stride - pad + (r, s).
// Tap (r, s) reads parity sub-lattice ((r-pad) mod stride, (s-pad) mod stride) at box shift (floor((r-pad)/stride), ...).
static int build_gather0(const dmu_tensor4& x, int N, int Hi, int Wi, int C, int R, int S, int stride, int pad, const Box& b, Tap* taps,
                         Maps* maps, int* map_h, int* map_w, const char* what) {
    bool used[4] = {false, false, false, false};
    for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) {
            const int ah = floordiv(r - pad, stride), aw = floordiv(s - pad, stride);
            const int ph = (r - pad) - ah * stride, pw = (s - pad) - aw * stride;
            Tap& t = taps[r * S + s];
            t.dh = ah; t.dw = aw; t.map = ph * stride + pw; t.wk = (r * S + s) * C;
            used[t.map] = true;
        }
    for (int m = 0; m < stride * stride; ++m) {
        map_h[m] = map_w[m] = 0;
        if (!used[m]) continue;
        const int ph = m / stride, pw = m % stride;
        if (ph >= Hi || pw >= Wi) {  // empty sub-lattice (e.g. 1x1 input, stride 2): every tap on it is out of bounds
            used[m] = false;
            continue;
        }
        const int Hm = (Hi - ph + stride - 1) / stride, Wm = (Wi - pw + stride - 1) / stride;
        map_h[m] = Hm; map_w[m] = Wm;
        const uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wm, (uint64_t)Hm, (uint64_t)N};
        const uint64_t str[4] = {1, (uint64_t)x.sw * stride, (uint64_t)x.sh * stride, (uint64_t)x.sn};
        const uint32_t box[4] = {64, (uint32_t)b.BW, (uint32_t)b.BH, (uint32_t)b.BN};
        const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(x.ptr) + (int64_t)ph * x.sh + (int64_t)pw * x.sw;
        if (maps) {      // nullptr: geometry only
            if (int rc = make_map_bf16(&maps->a[m], base, 4, dims, str, box, what)) return rc;
        }
    }
    return 0;
}

// GroupNorm fused into the epilogue (dmu_conv_params.gn_fuse), device view.  The tile of a CTA then holds whole images
// (tiles_h == tiles_w == 1) and whole groups (64 % cpg == 0), so the statistics never leave the CTA.
struct GnEpi {
    int mode;                 // 0 none, 1 forward (normalise this launch's output), 2 backward (the accumulator is dy of the norm)
    int G, cpg, silu, C;
    float eps, cnt;           // cnt = cpg * H * W
    const float* gamma; const float* beta;
    float* sums;              // [N][G][2] raw (sum, sum of squares): written in mode 1, read in mode 2
    __nv_bfloat16* a; int64_t a_sn, a_sh, a_sw;              // mode 1: normalised (+SiLU) output
    const __nv_bfloat16* x; int64_t x_sn, x_sh, x_sw;        // mode 2: input of the norm
    __nv_bfloat16* dx; int64_t dx_sn, dx_sh, dx_sw;          // mode 2: result
    const __nv_bfloat16* add0; int64_t a0_sn, a0_sh, a0_sw;
    const __nv_bfloat16* add1; int64_t a1_sn, a1_sh, a1_sw;
    float* red;               // mode 2: [tiles][C][2] per-tile channel sums (sum du, sum du * xhat)
    int epi_off;              // byte offset from the ring base of a region that is NOT part of the ring (written while the pipeline runs)
};

// ================================================================================================ fprop / dgrad kernel
struct ConvArgs {
    Tap taps[kMaxTaps];
    Phase phases[kMaxPhases];
    int map_h[4], map_w[4];      // extents of each A tensor map (to skip taps whose whole box is out of bounds)
    int N, Ho, Wo, Ck, Cj, os;
    int BN, BH, BW, tiles_h, tiles_w;
    __nv_bfloat16* y; int64_t y_sn, y_sh, y_sw;
    const __nv_bfloat16* res; int64_t r_sn, r_sh, r_sw;
    const float* bias;
    const float* temb; int64_t temb_pitch;
    // split-K: gridDim.z = phases * splits, the `splits` CTAs of one output tile form a thread-block cluster; partial tiles
    // go to `ws` (fp32 [slot][split][128][NT], plain stores), and after a cluster barrier every CTA folds and finishes
    // 128/splits rows of the tile.  No atomics (REDG tops out near 0.2 G floats/us chip-wide) and no counters.
    int splits;
    float* ws;
    long long* dbg;   // development aid (dmu_debug_set_buffer): per-CTA clock64 stamps of the pipeline phases
    int prefetch;     // epilogue operands fetched while the pipeline runs (DMU_EPI_PREFETCH=0 turns it off: A/B aid)
    int prefetch_w;   // filter boxes prefetched into L2 before griddepcontrol.wait (DMU_W_PREFETCH=0: A/B aid)
    int stages;       // ring depth actually used (<= ConvCfg::kStages): sub-wave launches leave room for a weight-gradient CTA
    int stage_bytes, a_off;   // stage stride and offset of the filter tile: a 64-pixel box only reserves 8 KB for the pixels
    int kps, kb_bytes;   // k-blocks per ring stage and bytes of one k-block (pixels + filters)
    int m64;             // 64-pixel tile computed by M = 64 MMAs
    int pix256;          // GroupNorm launches: a 256-pixel box (one 16x16 image) computed as two M = 128 accumulators
    GnEpi gn;
};

constexpr int kBtImgs = 8;
constexpr int kMaxSplitCtas = 160;      // CTAs of a split launch: about one wave
constexpr int64_t kSplitWsBytes = (int64_t)kMaxSplitCtas * 128 * 128 * 4;

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}


// ---------------------------------------------------------------------------- GroupNorm epilogue helpers (GnEpi)
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void unpack_bf16x8(const uint4& r, float* out) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
}
// The GroupNorm epilogue runs on the four warps of one CTA: every loop below is latency-bound, so the activations are
// branch-free (a select, not a jump: eight independent elements stay interleaved) and the loops carry 4-8 independent chains.
// One-MUFU sigmoid, sigmoid(u) = 0.5 tanh(u / 2) + 0.5 - the arithmetic of norm.cu's bf16 kernels (sigmoid_fast), so that the fused
// and the stand-alone GroupNorm agree; tanh.approx.f32 is ~2^-11 relative, below the 2^-9 of the bf16 store that follows.
__device__ __forceinline__ float gn_sigmoid(float u) {
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * u));
    return fmaf(0.5f, th, 0.5f);
}
__device__ __forceinline__ float gn_act(float u, int silu) {
    const float v = u * gn_sigmoid(u);
    return silu ? v : u;
}
// du = dy * act'(u)
__device__ __forceinline__ float gn_act_grad(float u, float dy, int silu) {
    const float sg = gn_sigmoid(u);
    const float r = dy * (sg * fmaf(u, 1.f - sg, 1.f));
    return silu ? r : dy;
}
// DEEP = 0: 2 CTAs per SM with a short ring (grids of several waves: the co-resident CTA hides the pipeline bubbles);
// DEEP = 1: 1 CTA per SM with as many stages as shared memory holds.  One TMA round trip is ~1 us, so a CTA moves at most
// (stages x stage bytes) per us: layers whose whole grid is under one wave (the <= 8x8 stages) are bound by exactly that.
template <int NT, int DEEP>
struct ConvCfg {
    static constexpr int kStages = DEEP ? (NT == 64 ? 8 : 6) : (NT == 64 ? 4 : 3);
    static constexpr int kABytes = 128 * 128;     // 128 pixels x 64 bf16
    static constexpr int kBBytes = NT * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kSmem = kStages * kStageBytes + 1024;   // + alignment slack
};

// GN = 1: the instantiation with the fused GroupNorm epilogues (GnEpi); the plain kernels stay lean (registers / code size)
// The GN = 1 kernels run 256 threads: warps 4-7 idle through the pipeline and then take the upper 32 channels of every tile row in the
// epilogue (a warp reads the TMEM lane quarter warp % 4), which halves every latency-bound phase of the norm.
template <int NT, int DEEP, int GN, int TB2 = 0>      // TB2: GroupNorm launch on a 256-pixel box (two accumulators, ConvArgs::pix256)
__device__ __forceinline__ void conv_tc_body(const Maps& maps, const ConvArgs& P) {
    using Cfg = ConvCfg<NT, DEEP>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_bar;
    __shared__ uint32_t s_tmem, s_issued;
    __shared__ __align__(16) float s_bt[kBtImgs * NT];     // bias + temb[n] of the (<= kBtImgs) images this tile touches
    __shared__ __align__(16) float s_gb[GN ? 2 * NT : 4];  // GroupNorm gamma | beta of this CTA's channels
    __shared__ float s_x[GN ? 8 * 32 : 1];                 // GroupNorm: (image, group) sums handed between the two warps of a 64-row image
    __shared__ float s_red[GN ? 8 * 64 : 1];               // GroupNorm backward: per-warp channel sums

    const int split = (int)blockIdx.z % P.splits;
    const Phase ph = P.phases[blockIdx.z / P.splits];
    const int tile = blockIdx.x;
    const int tw0 = (tile % P.tiles_w) * P.BW;
    const int th0 = ((tile / P.tiles_w) % P.tiles_h) * P.BH;
    const int n0 = (tile / (P.tiles_w * P.tiles_h)) * P.BN;
    if (th0 >= ph.TH || tw0 >= ph.TW) {          // phase smaller than the grid's tile space (whole CTA leaves)
        pdl_wait();
        return;
    }
    const int j0 = blockIdx.y * NT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = P.Ck >> 6;

    pdl_trigger();
    long long* dbg = P.dbg ? P.dbg + 8 * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&acc_bar, 2);
        fence_mbar_init();
        s_issued = 0;
        tma_prefetch_desc(&maps.b);
    }
    if (warp == 1) tmem_alloc(&s_tmem, NT << TB2);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (P.prefetch_w && warp == 0 && elect_one()) {
        // the filter boxes of this CTA's first k-blocks go to L2 while the previous kernel still runs (they are parameters;
        // after a forward + backward sweep over the 640 MB activation arena they would otherwise come from HBM)
        const int chunks_ = P.Ck >> 6;
        int n = 0;
        for (int ti = 0; ti < ph.ntaps && n < kStages; ++ti) {
            const Tap t = P.taps[ph.tap0 + ti];
            const int h = th0 + t.dh, w = tw0 + t.dw;
            if (!(h < P.map_h[t.map] && h + P.BH > 0 && w < P.map_w[t.map] && w + P.BW > 0)) continue;
            for (int c = 0; c < chunks_ && n < kStages; ++c, ++n) tma_prefetch_l2_2d(&maps.b, t.wk + c * 64, j0);
        }
    }
    __syncwarp();
    pdl_wait();      // everything above overlapped the previous kernel's tail; its outputs are visible from here on
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();

    // Epilogue operands are fetched NOW, while the pipeline below runs: this thread's residual row goes to registers and
    // (bias + temb[n]) of the tile's images to shared memory, so that after the last MMA only TMEM loads, adds and the
    // stores remain (they used to be four dependent rounds of L2 loads, ~2 us of a ~8 us sub-wave launch).
    // M = 128: accumulator row r sits in TMEM lane r.  M = 64 (the 64-pixel tiles): row r sits in lane (r / 16) * 32 + r % 16
    // (scripts/probes/umma_m64_layout.cu), i.e. lanes 0-15 of every warp's quadrant hold 16 consecutive rows.
    const int row = P.m64 ? (lane < 16 ? warp * 16 + lane : 128) : (int)(threadIdx.x & 127);
    const int c0 = GN ? (int)(threadIdx.x >> 7) * 32 : 0;       // GN: this thread's 32-channel half of the row
    const int nstage = (int)blockDim.x - 64;                    // threads of warps >= 2 (staging helpers)
    const int wl = row % P.BW, hl = (row / P.BW) % P.BH, nl = row / (P.BW * P.BH);
    const int n = n0 + nl, th = th0 + hl, tw = tw0 + wl;
    const int ho = th * P.os + ph.oph, wo = tw * P.os + ph.opw;
    const bool valid = nl < P.BN && n < P.N && th < ph.TH && tw < ph.TW && ho < P.Ho && wo < P.Wo;
    const bool stage_bt = P.prefetch && P.splits == 1 && P.BN <= kBtImgs && (P.bias || P.temb);
    const __nv_bfloat16* rp = (P.res && valid && P.splits == 1) ? P.res + (int64_t)n * P.r_sn + (int64_t)ho * P.r_sh + (int64_t)wo * P.r_sw + j0 : nullptr;
    uint4 rpre[NT / 8];
    // GN, 256-pixel box (one 16x16 image): this thread's second row, r + 128, lives in the second accumulator
    int ho2 = 0, wo2 = 0;
    bool valid2 = false;
    uint4 rpre2[GN ? 4 : 1];
    if constexpr (GN) {
        // backward mode: the same registers hold this row of the norm's input x (a dgrad has no residual)
        if (P.gn.mode == 2) rp = valid ? P.gn.x + (int64_t)n * P.gn.x_sn + (int64_t)ho * P.gn.x_sh + (int64_t)wo * P.gn.x_sw + j0 : nullptr;
#pragma unroll
        for (int i = 0; i < NT / 8; ++i) rpre[i] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < 4; ++i) rpre2[i] = make_uint4(0u, 0u, 0u, 0u);
        if (rp) {      // this thread's 32 channels only
#pragma unroll
            for (int i = 0; i < 4; ++i) rpre[i] = *reinterpret_cast<const uint4*>(rp + c0 + i * 8);
        }
        if constexpr (TB2 != 0) {
            const int row2 = row + 128;
            const int th2 = th0 + (row2 / P.BW) % P.BH, tw2 = tw0 + row2 % P.BW;
            ho2 = th2 * P.os + ph.oph; wo2 = tw2 * P.os + ph.opw;
            valid2 = n < P.N && th2 < ph.TH && tw2 < ph.TW && ho2 < P.Ho && wo2 < P.Wo;
            const __nv_bfloat16* rp2 = nullptr;
            if (valid2 && P.gn.mode == 2) rp2 = P.gn.x + (int64_t)n * P.gn.x_sn + (int64_t)ho2 * P.gn.x_sh + (int64_t)wo2 * P.gn.x_sw + j0;
            else if (valid2 && P.res) rp2 = P.res + (int64_t)n * P.r_sn + (int64_t)ho2 * P.r_sh + (int64_t)wo2 * P.r_sw + j0;
            if (rp2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rpre2[i] = *reinterpret_cast<const uint4*>(rp2 + c0 + i * 8);
            }
        }
    } else if (rp && P.prefetch) {
#pragma unroll
        for (int i = 0; i < NT / 8; ++i) rpre[i] = *reinterpret_cast<const uint4*>(rp + i * 8);
    }
    if constexpr (GN) {
        if (warp >= 2) {
            for (int i = threadIdx.x - 64; i < NT; i += nstage) {
                s_gb[i] = __ldg(P.gn.gamma + j0 + i);
                s_gb[NT + i] = __ldg(P.gn.beta + j0 + i);
            }
            if (P.gn.mode == 2) {      // (mean, rstd) of every (image, group) of the tile, same arithmetic as norm.cu's stage_affine
                float* s_mr = reinterpret_cast<float*>(smem + P.gn.epi_off);
                const int GT = NT / P.gn.cpg;
                for (int i = threadIdx.x - 64; i < P.BN * GT; i += nstage) {
                    const int img = i / GT, g = i - img * GT;
                    float mean = 0.f, rstd = 0.f;
                    if (n0 + img < P.N) {
                        const float* sp = P.gn.sums + ((int64_t)(n0 + img) * P.gn.G + j0 / P.gn.cpg + g) * 2;
                        const float inv_cnt = 1.f / P.gn.cnt;
                        mean = sp[0] * inv_cnt;
                        rstd = rsqrtf(fmaxf(fmaf(sp[1], inv_cnt, -mean * mean), 0.f) + P.gn.eps);
                    }
                    s_mr[img * (2 * GT + 1) + 2 * g] = mean; s_mr[img * (2 * GT + 1) + 2 * g + 1] = rstd;     // odd image stride: no bank conflicts
                }
            }
        }
    }
    if (stage_bt && warp >= 2) {
        for (int i = threadIdx.x - 64; i < P.BN * NT; i += nstage) {
            const int img = i / NT, c = i % NT;
            float v = P.bias ? __ldg(P.bias + j0 + c) : 0.f;
            if (P.temb && n0 + img < P.N) v += __ldg(P.temb + (int64_t)(n0 + img) * P.temb_pitch + j0 + c);
            s_bt[i] = v;
        }
    }

    auto tap_live = [&](const Tap& t) {   // does the shifted box intersect its sub-lattice at all?
        const int h = th0 + t.dh, w = tw0 + t.dw;
        return h < P.map_h[t.map] && h + P.BH > 0 && w < P.map_w[t.map] && w + P.BW > 0;
    };

    // this CTA's share of the (tap, 64-channel chunk) k-blocks
    const int kb_total = ph.ntaps * chunks;
    const int kb_per = (kb_total + P.splits - 1) / P.splits;
    const int kb_lo = split * kb_per, kb_hi = min(kb_total, kb_lo + kb_per);

    // Each role is ONE elected lane of a converged warp that runs its whole loop inside a single elect_one() region (see
    // tc_common.cuh; measured with scripts/probes/umma_rate.cu: re-entering the region, a whole-warp barrier wait and a
    // __syncwarp per k-block cost ~105 clk per MMA against 48-65 for the loop below - these k-loops are issue-bound).
    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        // A ring stage holds P.kps consecutive k-blocks (sub-wave launches: 2): one barrier wait and one commit per stage is
        // what the issuer pays, and its loop is issue-bound (~105 clk per MMA with one k-block per stage, see above).
        if (elect_one()) {
            const uint32_t a_bytes = (uint32_t)(P.BN * P.BH * P.BW) * 128u;
            int st = 0, par = 1, slot = 0;
            uint32_t tx = 0;
            for (int ti = 0; ti < ph.ntaps; ++ti) {
                const Tap t = P.taps[ph.tap0 + ti];
                if (!tap_live(t)) continue;
                for (int c = 0; c < chunks; ++c) {
                    const int kb = ti * chunks + c;
                    if (kb < kb_lo || kb >= kb_hi) continue;
                    if (slot == 0) mbar_wait(&empty_bar[st], par);
                    uint8_t* sa = smem + st * P.stage_bytes + slot * P.kb_bytes;
                    tma_load_4d(sa, &maps.a[t.map], &full_bar[st], c * 64, tw0 + t.dw, th0 + t.dh, n0);
                    tma_load_2d(sa + P.a_off, &maps.b, &full_bar[st], t.wk + c * 64, j0);
                    tx += a_bytes + Cfg::kBBytes;
                    if (++slot == P.kps) {
                        // posted after the loads: the barrier's pending arrival keeps the phase open, the transaction count may
                        // run negative in between
                        mbar_arrive_expect_tx(&full_bar[st], tx);
                        slot = 0; tx = 0;
                        if (++st == P.stages) { st = 0; par ^= 1; }
                    }
                }
            }
            if (slot) mbar_arrive_expect_tx(&full_bar[st], tx);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        if (elect_one()) {
            // a 64-pixel tile runs as an M = 64 MMA: 32 instead of 48 clk at N = 64 (the SS-mode MMA is paced by its operand
            // reads from shared memory, (M + N) x 32 B at 128 B/clk), and no read of rows the TMA box never wrote
            const uint32_t idesc = P.m64 ? umma_idesc_bf16(64, NT, 0, 0) : umma_idesc_bf16(128, NT, 0, 0);
            const uint32_t smem0 = smem_u32(smem);
            int nkb = 0;                                  // live k-blocks of this CTA
            for (int ti = 0; ti < ph.ntaps; ++ti) {
                if (!tap_live(P.taps[ph.tap0 + ti])) continue;
                for (int c = 0; c < chunks; ++c) {
                    const int kb = ti * chunks + c;
                    nkb += (kb >= kb_lo && kb < kb_hi) ? 1 : 0;
                }
            }
            int st = 0, par = 0;
            long long waited = 0;
            // descriptors advance by plain 64-bit adds on the encoded start address (16-byte units; the ring stays far below
            // the 256 KB the 14-bit field covers): the issue loop is latency-bound, every scalar instruction in it counts
            const uint64_t d_ring = smem_desc_sw128(smem0, 16, 1024);
            const uint64_t stage16 = (uint64_t)(P.stage_bytes >> 4), kb16 = (uint64_t)(P.kb_bytes >> 4), aoff16 = (uint64_t)(P.a_off >> 4);
            uint64_t d_stage = d_ring;
            for (int g = 0; g < nkb; g += P.kps) {
                const int cnt = min(P.kps, nkb - g);
                const long long w0 = dbg ? clock64() : 0;
                mbar_wait(&full_bar[st], par);
                tc_fence_after();
                if (dbg) { waited += clock64() - w0; if (g == 0) dbg[2] = clock64(); }      // first stage landed
                uint64_t da = d_stage;
#pragma unroll 1
                for (int j = 0; j < cnt; ++j, da += kb16) {
                    const uint64_t db = da + aoff16;
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // 4 x K=16 inside the 128-byte swizzle row: +32 B per step
                        umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, (g | j | k) != 0);
                    if constexpr (TB2 != 0) {     // rows 128..255 of the box (16 KB further on) into the second accumulator
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(tmem + NT, da + 1024 + 2 * k, db + 2 * k, idesc, (g | j | k) != 0);
                    }
                }
                umma_commit(&empty_bar[st]);
                d_stage += stage16;
                if (++st == P.stages) { st = 0; par ^= 1; d_stage = d_ring; }
            }
            if (dbg) { dbg[3] = clock64(); dbg[6] = nkb; dbg[7] = waited; }   // last MMA issued; cycles spent waiting for operands
            s_issued = (uint32_t)nkb;
            umma_commit(&acc_bar);      // arrival 1: all MMAs retired
            mbar_arrive(&acc_bar);      // arrival 2: release-publishes s_issued to the epilogue threads
        }
        __syncwarp();
    }
    __syncwarp();

    // ---------------------------------------------------- epilogue: all 4 warps, thread = one output pixel (TMEM lane)
    if (stage_bt || GN) __syncthreads();     // s_bt (and the GroupNorm staging) written by warps 2-3 while the pipeline ran
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    if (dbg && threadIdx.x == 0) dbg[4] = clock64();   // accumulator complete
    const bool have_acc = *reinterpret_cast<volatile uint32_t*>(&s_issued) != 0;
    if constexpr (GN) {
        // ------------------------------------------------ fused GroupNorm epilogue (splits == 1, whole images per tile)
        // The norm runs out of REGISTERS: a thread holds 32 channels (= 32 >> SH whole groups) of one tile row, the rows of one
        // image are HWt consecutive TMEM lanes (HWt = BH * BW, a power of two), so the (image, group) sums are a butterfly over
        // min(HWt, 32) adjacent lanes plus one exchange between the warps of an image when it spans several, and nothing is
        // parked in shared memory.  (The first version staged the tile in shared memory and ran three 256-thread phases with
        // two barriers: 6.7k clk forward / 10.2k clk backward per launch against 2.2k for the plain epilogue,
        // scripts/gn_epi_timeline.py.)  A 256-pixel box (16x16 images: P.pix256) is two accumulators of 128 rows; a thread then
        // owns rows r and r + 128, sums over both and reads the accumulators a second time for the result instead of keeping
        // both rows in registers.  One instantiation per channels-per-group value (2, 4, 8, 16, 32), chosen at run time.
        const GnEpi& Gn = P.gn;
        constexpr int TB = 1 + TB2;
        const int HWt = P.BH * P.BW;
        const int seg = HWt < 32 ? HWt : 32;
        const int wpi = HWt <= 32 ? 1 : (HWt == 64 ? 2 : 4);      // warps one image spans inside a 128-row accumulator
        const int silu = Gn.silu;
        const float inv_cnt = 1.f / Gn.cnt;
        const float* gam = s_gb + c0;
        const float* bet = s_gb + NT + c0;
        float v[32];
        // accumulator `t` of this thread's row -> v (zeros for rows outside the problem)
        auto load_acc = [&](int t, bool ok) {
            tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * NT + c0), v);
            tmem_ld_wait();
            if (!have_acc || !ok) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
        };
        // sum over the rows of an image: lanes of the warp first (all 32 lanes take part; rows outside the problem carry zeros),
        // then the warps the image spans
        auto image_allreduce = [&](auto& a) {
            constexpr int NV = sizeof(a) / sizeof(float);
            for (int o = seg >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int i = 0; i < NV; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
            }
            if (wpi > 1) {
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < NV; ++i) s_x[warp * 32 + i] = a[i];
                }
                __syncthreads();
                const float* sp = s_x + (warp & ~(wpi - 1)) * 32;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    float s = sp[i] + sp[32 + i];
                    if (wpi == 4) s += sp[64 + i] + sp[96 + i];
                    a[i] = s;
                }
            }
        };
        // One accumulator (TB == 1): a single step forms the sums, reduces them and writes the result from the same registers.
        // Two (TB == 2): steps 0, 1 form the sums of rows r, r + 128, step 1 ends with the reduction, steps 2, 3 read the
        // accumulators again (same arithmetic, same values) and write the results.  Every piece of code below exists once.
        constexpr int nsteps = TB == 1 ? 1 : 4;
        if (Gn.mode == 1) {
            // ---- forward: y = conv (+ bias + temb + residual), a = act(GroupNorm(y))
            auto fwd = [&](auto shc) {
                constexpr int SH = decltype(shc)::value, CPG = 1 << SH, NG = 32 >> SH;
                float st[2 * NG];
#pragma unroll
                for (int i = 0; i < 2 * NG; ++i) st[i] = 0.f;
#pragma unroll 1
                for (int s = 0; s < nsteps; ++s) {
                    const int t = s & 1;
                    const bool second = s >= 2, vt = t ? valid2 : valid;
                    const int hot = t ? ho2 : ho, wot = t ? wo2 : wo;
                    load_acc(t, vt);
                    if (vt) {
                        if (stage_bt) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 b = *reinterpret_cast<const float4*>(&s_bt[nl * NT + c0 + i]);
                                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                            }
                        } else {
                            if (P.bias) {
#pragma unroll
                                for (int i = 0; i < 32; i += 4) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + j0 + c0 + i));
                                    v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                                }
                            }
                            if (P.temb) {
                                const float* tp = P.temb + (int64_t)n * P.temb_pitch + j0 + c0;
#pragma unroll
                                for (int i = 0; i < 32; i += 4) {
                                    const float4 b = __ldg(reinterpret_cast<const float4*>(tp + i));
                                    v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                                }
                            }
                        }
                        if (P.res) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float r8[8];
                                unpack_bf16x8(t ? rpre2[i] : rpre[i], r8);
#pragma unroll
                                for (int k = 0; k < 8; ++k) v[i * 8 + k] += r8[k];
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]);
                        if (!second) {
                            __nv_bfloat16* yp1 = P.y + (int64_t)n * P.y_sn + (int64_t)hot * P.y_sh + (int64_t)wot * P.y_sw + j0 + c0;
#pragma unroll
                            for (int i = 0; i < 32; i += 8) store_vec<__nv_bfloat16>(yp1 + i, v + i);
                        }
                    }
                    if (!second) {
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
                            float su = 0.f, q = 0.f;
#pragma unroll
                            for (int k = 0; k < CPG; ++k) { const float x = v[g * CPG + k]; su += x; q = fmaf(x, x, q); }
                            st[2 * g] += su; st[2 * g + 1] += q;
                        }
                    }
                    if (s == TB - 1) {
                        image_allreduce(st);
                        // raw (sum, sum of squares) for the backward: the image's first row writes (of the first warp when it spans several)
                        if (valid && (lane & (seg - 1)) == 0 && (warp & (wpi - 1)) == 0) {
                            float2* sp = reinterpret_cast<float2*>(Gn.sums + ((int64_t)n * Gn.G + ((j0 + c0) >> SH)) * 2);
#pragma unroll
                            for (int g = 0; g < NG; ++g) sp[g] = make_float2(st[2 * g], st[2 * g + 1]);
                        }
                    }
                    if ((TB == 1 || second) && vt) {
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
                            const float mean = st[2 * g] * inv_cnt;
                            const float rstd = rsqrtf(fmaxf(fmaf(st[2 * g + 1], inv_cnt, -mean * mean), 0.f) + Gn.eps);
#pragma unroll
                            for (int k = 0; k < CPG; ++k) {
                                const int i = g * CPG + k;
                                const float sc = rstd * gam[i];
                                v[i] = gn_act(fmaf(v[i], sc, bet[i] - mean * sc), silu);
                            }
                        }
                        __nv_bfloat16* ap = Gn.a + (int64_t)n * Gn.a_sn + (int64_t)hot * Gn.a_sh + (int64_t)wot * Gn.a_sw + j0 + c0;
#pragma unroll
                        for (int i = 0; i < 32; i += 8) store_vec<__nv_bfloat16>(ap + i, v + i);
                    }
                }
            };
            switch (Gn.cpg) {
                case 2: fwd(IntC<1>{}); break;
                case 4: fwd(IntC<2>{}); break;
                case 8: fwd(IntC<3>{}); break;
                case 16: fwd(IntC<4>{}); break;
                default: fwd(IntC<5>{}); break;
            }
        } else {
            // ---- backward: the accumulator is dy of the norm's output; rpre holds this thread's part of row x (zeros, like dy, for rows
            //      outside the problem);  du = dy act'(u),  dx = rstd (gamma du - A - xhat B) + add0 + add1  with
            //      A = mean_group(gamma du), B = mean_group(gamma du xhat);  per-tile channel sums (sum du, sum du xhat) for dgamma / dbeta
            float xh[32];
            const int64_t px = (int64_t)j0 + c0;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            uint4 r0[4], r1[4];      // the addends fly while the sums are formed
            auto load_adds = [&](bool ok, int hot, int wot) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { r0[i] = z; r1[i] = z; }
                if (ok && Gn.add0) {
                    const __nv_bfloat16* a0p = Gn.add0 + (int64_t)n * Gn.a0_sn + (int64_t)hot * Gn.a0_sh + (int64_t)wot * Gn.a0_sw + px;
#pragma unroll
                    for (int i = 0; i < 4; ++i) r0[i] = *reinterpret_cast<const uint4*>(a0p + i * 8);
                }
                if (ok && Gn.add1) {
                    const __nv_bfloat16* a1p = Gn.add1 + (int64_t)n * Gn.a1_sn + (int64_t)hot * Gn.a1_sh + (int64_t)wot * Gn.a1_sw + px;
#pragma unroll
                    for (int i = 0; i < 4; ++i) r1[i] = *reinterpret_cast<const uint4*>(a1p + i * 8);
                }
            };
            float w0 = 0.f, w1 = 0.f;      // this lane's channel (c0 + lane): sum du, sum du xhat over the warp's rows
            auto bwd = [&](auto shc) {
                constexpr int SH = decltype(shc)::value, CPG = 1 << SH, NG = 32 >> SH;
                const int IS = 2 * (NT >> SH) + 1;
                const float* mr = reinterpret_cast<const float*>(smem + Gn.epi_off) + (valid ? nl : 0) * IS + 2 * (c0 >> SH);
                float ab[2 * NG];
#pragma unroll
                for (int i = 0; i < 2 * NG; ++i) ab[i] = 0.f;
                load_adds(valid, ho, wo);
#pragma unroll 1
                for (int s = 0; s < nsteps; ++s) {
                    const int t = s & 1;
                    const bool second = s >= 2, vt = t ? valid2 : valid;
                    const int hot = t ? ho2 : ho, wot = t ? wo2 : wo;
                    if (second && t == 1) load_adds(vt, hot, wot);
                    // v <- du, xh <- xhat of accumulator t
                    load_acc(t, vt);
#pragma unroll
                    for (int i = 0; i < 4; ++i) unpack_bf16x8(t ? rpre2[i] : rpre[i], xh + i * 8);
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const float mean = mr[2 * g], rstd = mr[2 * g + 1];
                        float A = 0.f, B = 0.f;
#pragma unroll
                        for (int k = 0; k < CPG; ++k) {
                            const int i = g * CPG + k;
                            const float d = xh[i] - mean;
                            const float du = gn_act_grad(fmaf(d, rstd * gam[i], bet[i]), v[i], silu);
                            const float xn = d * rstd;
                            const float gd = gam[i] * du;
                            v[i] = du; xh[i] = xn;
                            A += gd; B = fmaf(gd, xn, B);
                        }
                        if (!second) { ab[2 * g] += A; ab[2 * g + 1] += B; }
                    }
                    if (s == TB - 1) image_allreduce(ab);
                    if (TB == 1 || second) {
                        // per-tile channel sums over the warp's 32 rows: a transposing butterfly - at offset o a lane keeps the half of
                        // its values its bit o selects and adds the partner's copy of that half - leaves lane L with channel c0 + L
                        {
                            float w[64];
#pragma unroll
                            for (int i = 0; i < 32; ++i) { w[2 * i] = v[i]; w[2 * i + 1] = v[i] * xh[i]; }
#pragma unroll
                            for (int half = 32; half >= 2; half >>= 1) {
                                const bool up = (lane & (half >> 1)) != 0;
#pragma unroll
                                for (int i = 0; i < half; ++i) {
                                    const float keep = up ? w[half + i] : w[i];
                                    const float send = up ? w[i] : w[half + i];
                                    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, half >> 1);
                                }
                            }
                            w0 += w[0]; w1 += w[1];
                        }
                        if (vt) {
#pragma unroll
                            for (int g = 0; g < NG; ++g) {
                                const float rstd = mr[2 * g + 1], A = ab[2 * g] * inv_cnt, B = ab[2 * g + 1] * inv_cnt;
#pragma unroll
                                for (int k = 0; k < CPG; ++k) {
                                    const int i = g * CPG + k;
                                    v[i] = rstd * (gam[i] * v[i] - A - xh[i] * B);
                                }
                            }
                            __nv_bfloat16* dxp = Gn.dx + (int64_t)n * Gn.dx_sn + (int64_t)hot * Gn.dx_sh + (int64_t)wot * Gn.dx_sw + px;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float t0[8], t1[8];
                                unpack_bf16x8(r0[i], t0);
                                unpack_bf16x8(r1[i], t1);
#pragma unroll
                                for (int k = 0; k < 8; ++k) v[i * 8 + k] += t0[k] + t1[k];
                                store_vec<__nv_bfloat16>(dxp + i * 8, v + i * 8);
                            }
                        }
                    }
                }
            };
            switch (Gn.cpg) {
                case 2: bwd(IntC<1>{}); break;
                case 4: bwd(IntC<2>{}); break;
                case 8: bwd(IntC<3>{}); break;
                case 16: bwd(IntC<4>{}); break;
                default: bwd(IntC<5>{}); break;
            }
            s_red[warp * 64 + 2 * lane] = w0; s_red[warp * 64 + 2 * lane + 1] = w1;
            __syncthreads();
            if (threadIdx.x < 2 * NT) {       // (channel, which sum): the four warps that share a channel half
                const int e = threadIdx.x, hf = e >> 6, r = e & 63;
                const float* sp = s_red + hf * 4 * 64 + r;
                Gn.red[((int64_t)tile * Gn.C + j0) * 2 + e] = (sp[0] + sp[64]) + (sp[128] + sp[192]);
            }
        }
        if (dbg && threadIdx.x == 0) dbg[5] = clock64();   // epilogue stores issued
        tc_fence_before();
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem, NT << TB2);
        return;
    }
    if (P.splits > 1) {
        // ---- split-K: park the partial tile, cluster barrier, then fold + finish this CTA's share of the rows
        const int slot = ((int)(blockIdx.z / P.splits) * (int)gridDim.y + (int)blockIdx.y) * (int)gridDim.x + tile;
        float* tile_ws = P.ws + (size_t)slot * P.splits * 128 * NT;
        float* mine = tile_ws + ((size_t)split * 128 + row) * NT;
#pragma unroll 1
        for (int c = 0; c < NT; c += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 o = have_acc ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                __stcg(reinterpret_cast<float4*>(mine + c + i), o);
            }
        }
        __threadfence();
        cluster_sync_all();
        constexpr int kGroups = NT / 8;               // 8-column groups per row
        constexpr int kRowsPar = 128 / kGroups;       // rows handled at once by the 128 threads
        const int rows_mine = 128 / P.splits;
        const int cg = threadIdx.x % kGroups, rsub = threadIdx.x / kGroups;
        for (int r0 = 0; r0 < rows_mine; r0 += kRowsPar) {
            const int rr = split * rows_mine + r0 + rsub;
            if (r0 + rsub >= rows_mine) break;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int sp = 0; sp < P.splits; ++sp) {
                const float* src = tile_ws + ((size_t)sp * 128 + rr) * NT + cg * 8;
                const float4 a4 = __ldcg(reinterpret_cast<const float4*>(src)), b4 = __ldcg(reinterpret_cast<const float4*>(src + 4));
                v[0] += a4.x; v[1] += a4.y; v[2] += a4.z; v[3] += a4.w; v[4] += b4.x; v[5] += b4.y; v[6] += b4.z; v[7] += b4.w;
            }
            const int wl2 = rr % P.BW, hl2 = (rr / P.BW) % P.BH, nl2 = rr / (P.BW * P.BH);
            const int n2 = n0 + nl2, th2 = th0 + hl2, tw2 = tw0 + wl2;
            const int ho2 = th2 * P.os + ph.oph, wo2 = tw2 * P.os + ph.opw;
            if (!(nl2 < P.BN && n2 < P.N && th2 < ph.TH && tw2 < ph.TW && ho2 < P.Ho && wo2 < P.Wo)) continue;
            const int jc = j0 + cg * 8;
            if (P.bias) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(P.bias + jc)), b4 = __ldg(reinterpret_cast<const float4*>(P.bias + jc + 4));
                v[0] += a4.x; v[1] += a4.y; v[2] += a4.z; v[3] += a4.w; v[4] += b4.x; v[5] += b4.y; v[6] += b4.z; v[7] += b4.w;
            }
            if (P.temb) {
                const float* tp2 = P.temb + (int64_t)n2 * P.temb_pitch + jc;
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(tp2)), b4 = __ldg(reinterpret_cast<const float4*>(tp2 + 4));
                v[0] += a4.x; v[1] += a4.y; v[2] += a4.z; v[3] += a4.w; v[4] += b4.x; v[5] += b4.y; v[6] += b4.z; v[7] += b4.w;
            }
            if (P.res) {
                float r8[8];
                load_vec<__nv_bfloat16>(P.res + (int64_t)n2 * P.r_sn + (int64_t)ho2 * P.r_sh + (int64_t)wo2 * P.r_sw + jc, r8);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] += r8[k];
            }
            store_vec<__nv_bfloat16>(P.y + (int64_t)n2 * P.y_sn + (int64_t)ho2 * P.y_sh + (int64_t)wo2 * P.y_sw + jc, v);
        }
        tc_fence_before();
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem, NT);
        return;
    }
    __nv_bfloat16* yp = P.y + (int64_t)n * P.y_sn + (int64_t)ho * P.y_sh + (int64_t)wo * P.y_sw + j0;
    const float* tp = P.temb ? P.temb + (int64_t)n * P.temb_pitch + j0 : nullptr;
#pragma unroll
    for (int c = 0; c < NT; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
        if (valid) {
            if (!have_acc) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            if (stage_bt) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(&s_bt[nl * NT + c + i]);
                    v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                }
            } else {
                if (P.bias) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + j0 + c + i));
                        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                    }
                }
                if (tp) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(tp + c + i));
                        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                    }
                }
            }
            if (rp && !P.prefetch) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) rpre[(c + i) >> 3] = *reinterpret_cast<const uint4*>(rp + c + i);
            }
            if (rp) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rpre[(c + i) >> 3]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); v[i + 2 * k] += f.x; v[i + 2 * k + 1] += f.y; }
                }
            }
#pragma unroll
            for (int i = 0; i < 32; i += 8) store_vec<__nv_bfloat16>(yp + c + i, v + i);
        }
    }
    if (dbg && threadIdx.x == 0) dbg[5] = clock64();   // epilogue stores issued
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, NT);
}

template <int NT, int DEEP>
__global__ void __launch_bounds__(128) conv_tc_kernel(const __grid_constant__ Maps maps, const __grid_constant__ ConvArgs P) {
    conv_tc_body<NT, DEEP, 0>(maps, P);
}
// The GroupNorm variants run 256 threads; the register cap keeps a weight-gradient CTA of the backward's other lane (128 threads
// x 56 registers) resident on the same SM next to one of these.
template <int DEEP, int TB2>
__global__ void __maxnreg__(216) conv_tc_gn_kernel(const __grid_constant__ Maps maps, const __grid_constant__ ConvArgs P) {
    conv_tc_body<64, DEEP, 1, TB2>(maps, P);
}

long long* g_debug_buffer = nullptr;

// conv_halo.cu
int halo_supported(const dmu_conv_params* p, int force);
int halo_stats_supported(const dmu_conv_params* p);
int halo_launch(const dmu_conv_params* p, cudaStream_t stream);
// conv_halo.cu: 4x4 stride-2 transposed gather (ConvTranspose2d upsampling, input gradient of the 4x4 stride-2 downsampling conv)
int halo_t_supported(const dmu_conv_params* p, int force);
int halo_t_launch(const dmu_conv_params* p, cudaStream_t stream);
// conv_halo.cu: 4x4 stride-2 convolution (learned downsampling, input gradient of the ConvTranspose2d upsampling)
int halo_s_supported(const dmu_conv_params* p, int force);
int halo_s_launch(const dmu_conv_params* p, cudaStream_t stream);
// conv_stem.cu: few-channel input (stem fprop, head dgrad)
int stem_supported(const dmu_conv_params* p);
int stem_launch(const dmu_conv_params* p, cudaStream_t stream);

static bool halo_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DMU_HALO");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// C -> (<= 4) channel 3x3 head conv written as fp32 through arbitrary strides: runs on the halo kernel (the weight box rows
// past Cj are TMA zeros), reading the wide input once instead of the SIMT kernel's latency-bound gather
static int narrow_head_supported(const dmu_conv_params* p) {
    if (!p || !p->x.ptr || !p->y.ptr || !p->w || p->Cj > 4 || p->Cj < 1) return 0;
    if (p->w_dtype != DMU_BF16 || !nhwc_bf16_ok(p->x) || p->y.dtype != DMU_F32 || p->res.ptr || p->temb || p->gather != 0) return 0;
    if (p->Ck != 64 || p->w_sk != 1 || p->w_st != p->Ck || p->w_sn != (int64_t)9 * p->Ck || !aligned16(p->w)) return 0;
    if (encode_tiled_fn() == nullptr || !halo_enabled()) return 0;
    return halo_supported(p, 0);
}

static int conv_gn_fuse_tiles(const dmu_conv_params* p);
static int conv_supported(const dmu_conv_params* p) {
    if (!p || !p->x.ptr || !p->y.ptr || !p->w) return 0;
    if (p->gn_fuse_mode) return conv_gn_fuse_tiles(p) > 0 ? 1 : 0;
    if (p->gn_coef) {   // fused GroupNorm: the halo kernel only (wide bf16 NHWC input, 3x3 stride 1, >= 8x8)
        if (!nhwc_bf16_ok(p->x) || p->Ck % 64 != 0 || !halo_supported(p, 1) || encode_tiled_fn() == nullptr) return 0;
        if (p->a_out.ptr && !nhwc_bf16_ok(p->a_out)) return 0;
    }
    if (narrow_head_supported(p)) return 1;
    if (p->impl != 4 && p->impl != 5 && stem_supported(p)) return 1;
    if (p->w_dtype != DMU_BF16 || !nhwc_bf16_ok(p->x) || !nhwc_bf16_ok(p->y)) return 0;
    if (p->res.ptr && !nhwc_bf16_ok(p->res)) return 0;
    if (p->Ck % 64 != 0 || p->Cj % 64 != 0) return 0;
    if (p->w_sk != 1 || (p->R * p->S > 1 && p->w_st != p->Ck) || p->w_sn != (int64_t)p->R * p->S * p->Ck || !aligned16(p->w)) return 0;
    if (p->stride < 1 || p->stride > 2 || p->R * p->S > kMaxTaps) return 0;
    if (p->bias && !aligned16(p->bias)) return 0;
    if (p->temb && (!aligned16(p->temb) || p->temb_pitch % 4 != 0)) return 0;
    if (encode_tiled_fn() == nullptr) return 0;
    return 1;
}

// Geometry of one per-tap launch: everything conv_launch needs that does not depend on the tensor maps' contents.  With
// maps == nullptr nothing is encoded (dmu_conv2d_gn_fuse_supported asks "what would this launch look like" with pointers that
// may not be mapped yet).
struct ConvGeom { int NT, pix, nph; Box b; dim3 grid; bool deep; size_t smem; int gn_tiles; };

static bool gn_tensor_ok(const dmu_tensor4& t) {
    return t.dtype == DMU_BF16 && t.sc == 1 && aligned16(t.ptr) && t.sw % 8 == 0 && t.sh % 8 == 0 && t.sn % 8 == 0;
}

static int conv_prepare(const dmu_conv_params* p, Maps* maps, ConvArgs& A, ConvGeom& G) {
    memset(&A, 0, sizeof(A));
    const dmu_gn_params* gn = p->gn_fuse_mode ? reinterpret_cast<const dmu_gn_params*>(p->gn_fuse) : nullptr;
    DMU_REQUIRE(p->gn_fuse_mode == 0 || (gn && (p->gn_fuse_mode == 1 || p->gn_fuse_mode == 2)), "dmu_conv2d/tc: gn_fuse_mode %d without a valid gn_fuse", p->gn_fuse_mode);
    const int st = p->stride;
    int nph = 1, THmax = p->Ho, TWmax = p->Wo;
    if (p->gather == 1) {
        nph = st * st;
        THmax = (p->Ho + st - 1) / st;
        TWmax = (p->Wo + st - 1) / st;
    }
    // Tile shape.  A sub-wave layer (the <= 8x8 stages) is bound by ONE CTA's serial k-loop, which ingests ~64 B/clk from L2:
    // 16 KB of activations + 16 KB of filters per k-block at 128 pixels x 128 channels.  While the CTAs still fit one wave
    // the tile is therefore halved along channels (NT = 64) and then along pixels (a 64-pixel TMA box; the MMA still reads 128
    // rows, the upper 64 accumulator rows are never stored): 4x the CTAs, half the bytes and k-loop time per CTA.
    int NT = (p->Cj % 128 == 0) ? 128 : 64;
    int pix = 128;
    const int img_rows = pow2_ceil(THmax) * pow2_ceil(TWmax);      // tile rows one whole image occupies
    if (gn) NT = 64;        // the fused GroupNorm epilogues exist for 64-channel tiles (register budget of two passes over the tile)
    {
        static const int small_tiles = [] { const char* e = getenv("DMU_SMALL_TILES"); return e ? atoi(e) : 1; }();    // A/B aid
        static const int limit_pct = [] { const char* e = getenv("DMU_SMALL_TILES_LIMIT"); return e ? atoi(e) : 100; }();
        const int cap = sm_count() * limit_pct / 100;
        const Box b0 = make_box(p->N, THmax, TWmax, 128);
        int ctas = b0.tiles_n * b0.tiles_h * b0.tiles_w * (p->Cj / NT) * nph;
        if (small_tiles && NT == 128 && ctas * 2 <= cap) { NT = 64; ctas *= 2; }
        if (small_tiles && ctas * 2 <= cap && (int64_t)p->N * THmax * TWmax > 64 && !(gn && img_rows > 64)) { pix = 64; ctas *= 2; }
        if (small_tiles >= 2 && pix == 64 && ctas * 2 <= cap && (int64_t)p->N * THmax * TWmax > 32 && !gn) pix = 32;
        // GroupNorm epilogue on a 16x16 image: one 256-pixel box per CTA (two M = 128 accumulators), so that the tile still holds
        // the whole image.  Sub-wave grids only: per SM it is the same k-loop traffic as two co-resident 128-pixel CTAs.
        static const int pix256 = [] { const char* e = getenv("DMU_GN_PIX256"); return e ? atoi(e) : 1; }();       // A/B aid
        if (gn && pix256 && img_rows == 256 && THmax == 16 && TWmax == 16 && nph == 1 && st == 1 &&
            p->N * (p->Cj / 64) <= sm_count())
            pix = 256;
    }
    const Box b = make_box(p->N, THmax, TWmax, pix);
    if (p->gather == 0) {
        A.os = 1;
        A.phases[0] = Phase{0, p->R * p->S, 0, 0, p->Ho, p->Wo};
        if (int rc = build_gather0(p->x, p->N, p->Hi, p->Wi, p->Ck, p->R, p->S, st, p->pad, b, A.taps, maps, A.map_h, A.map_w, "dmu_conv2d/tc"))
            return rc;
    } else {
        // transposed gather: output parity class (oph, opw) uses the taps with (oph + pad - r) divisible by stride
        A.os = st;
        int nt = 0;
        for (int oph = 0; oph < st; ++oph)
            for (int opw = 0; opw < st; ++opw) {
                Phase& ph = A.phases[oph * st + opw];
                ph.tap0 = nt; ph.oph = oph; ph.opw = opw;
                ph.TH = p->Ho > oph ? (p->Ho - oph + st - 1) / st : 0;
                ph.TW = p->Wo > opw ? (p->Wo - opw + st - 1) / st : 0;
                for (int r = 0; r < p->R; ++r)
                    for (int s = 0; s < p->S; ++s) {
                        const int eh = oph + p->pad - r, ew = opw + p->pad - s;
                        if (eh - floordiv(eh, st) * st != 0 || ew - floordiv(ew, st) * st != 0) continue;
                        DMU_REQUIRE(nt < kMaxTaps, "dmu_conv2d/tc: too many taps");
                        A.taps[nt++] = Tap{floordiv(eh, st), floordiv(ew, st), 0, (r * p->S + s) * p->Ck};
                    }
                ph.ntaps = nt - ph.tap0;
            }
        A.map_h[0] = p->Hi; A.map_w[0] = p->Wi;
        if (maps) {
            const uint64_t dims[4] = {(uint64_t)p->Ck, (uint64_t)p->Wi, (uint64_t)p->Hi, (uint64_t)p->N};
            const uint64_t str[4] = {1, (uint64_t)p->x.sw, (uint64_t)p->x.sh, (uint64_t)p->x.sn};
            const uint32_t box[4] = {64, (uint32_t)b.BW, (uint32_t)b.BH, (uint32_t)b.BN};
            if (int rc = make_map_bf16(&maps->a[0], p->x.ptr, 4, dims, str, box, "dmu_conv2d/tc")) return rc;
        }
    }
    if (maps) {
        const uint64_t dims[2] = {(uint64_t)p->R * p->S * p->Ck, (uint64_t)p->Cj};
        const uint64_t str[2] = {1, (uint64_t)p->w_sn};
        const uint32_t box[2] = {64, (uint32_t)NT};
        if (int rc = make_map_bf16(&maps->b, p->w, 2, dims, str, box, "dmu_conv2d/tc weights")) return rc;
    }
    A.N = p->N; A.Ho = p->Ho; A.Wo = p->Wo; A.Ck = p->Ck; A.Cj = p->Cj;
    A.BN = b.BN; A.BH = b.BH; A.BW = b.BW; A.tiles_h = b.tiles_h; A.tiles_w = b.tiles_w;
    A.y = reinterpret_cast<__nv_bfloat16*>(p->y.ptr); A.y_sn = p->y.sn; A.y_sh = p->y.sh; A.y_sw = p->y.sw;
    A.res = reinterpret_cast<const __nv_bfloat16*>(p->res.ptr); A.r_sn = p->res.sn; A.r_sh = p->res.sh; A.r_sw = p->res.sw;
    A.bias = p->bias; A.temb = p->temb; A.temb_pitch = p->temb_pitch;
    A.dbg = g_debug_buffer;
    {
        static int pf = -1;
        if (pf < 0) { const char* e = getenv("DMU_EPI_PREFETCH"); pf = (e && e[0] == '0') ? 0 : 1; }
        A.prefetch = pf;
        static int pw = -1;
        if (pw < 0) { const char* e = getenv("DMU_W_PREFETCH"); pw = (e && e[0] == '1') ? 1 : 0; }   // measured: no gain, off by default
        A.prefetch_w = pw;
    }
    // split-K for layers with few output tiles and a long contraction (the <= 4x4 stages: K up to 4608, 1-16 tiles):
    // the splits of one tile are a thread-block cluster along z (co-scheduled by hardware, so the in-kernel barrier is safe)
    const int slots = b.tiles_n * b.tiles_h * b.tiles_w * (p->Cj / NT) * nph;
    // k-blocks that actually run in the first tile of each phase (taps whose whole box is padding are skipped in-kernel);
    // measured on B200: a launch has a ~9 us floor, so splitting pays only from ~40 live k-blocks per CTA upwards
    int kb_min = 1 << 30;
    for (int i = 0; i < nph; ++i) {
        int live = 0;
        for (int t = 0; t < A.phases[i].ntaps; ++t) {
            const Tap& tp = A.taps[A.phases[i].tap0 + t];
            if (tp.dh < A.map_h[tp.map] && tp.dh + b.BH > 0 && tp.dw < A.map_w[tp.map] && tp.dw + b.BW > 0) ++live;
        }
        const int kb = live * (p->Ck / 64);
        if (kb < kb_min) kb_min = kb;
    }
    A.splits = 1;
    static int split_min = -1, split_per = -1;
    if (split_min < 0) {
        const char* e1 = getenv("DMU_SPLITK_MIN"); const char* e2 = getenv("DMU_SPLITK_PER");
        split_min = e1 ? atoi(e1) : 40; split_per = e2 ? atoi(e2) : 12;
    }
    if (p->workspace && p->workspace_bytes >= kSplitWsBytes && slots * 2 <= sm_count() && kb_min >= split_min) {
        int sp = 8;                                               // portable cluster limit
        while (sp > 1 && (slots * sp > kMaxSplitCtas || slots * sp > sm_count() + sm_count() / 4 || kb_min / sp < split_per)) sp >>= 1;
        if (sp >= 2) {
            A.splits = sp;
            A.ws = reinterpret_cast<float*>(p->workspace);
        }
    }
    G.grid = dim3(b.tiles_n * b.tiles_h * b.tiles_w, p->Cj / NT, nph * A.splits);
    G.deep = (int)(G.grid.x * G.grid.y * G.grid.z) <= sm_count();
    const bool deep = G.deep;
    // Ring geometry.  One TMA round trip is ~2000 clk here (scripts/conv_timeline.py), so a CTA's k-loop moves at most
    // (bytes in flight) / 2000 clk: the sub-wave (1 CTA per SM) variants keep ~128 KB in flight - compact stages for the
    // 64-pixel box, 8 KB + NT x 128 B each - which still leaves room on the SM for a weight-gradient CTA of the backward's other
    // lane (3 x 24-32 KB, see wgrad_launch) instead of the chain waiting for one to retire.  DMU_CONV_RING_KB: A/B aid.
    static const int ring_kb = [] { const char* e = getenv("DMU_CONV_RING_KB"); return e ? atoi(e) : 128; }();
    static const int kps_env = [] { const char* e = getenv("DMU_CONV_KPS"); return e ? atoi(e) : 2; }();
    // measured: the k-loop of these launches is issue-bound (~100 clk per MMA whatever its shape), M = 64 buys nothing: opt-in
    static const int m64_env = [] { const char* e = getenv("DMU_CONV_M64"); return e ? atoi(e) : 0; }();
    A.m64 = (pix == 64 && A.splits == 1 && m64_env && !gn) ? 1 : 0;
    A.pix256 = pix == 256 ? 1 : 0;
    A.a_off = pix * 128;
    A.kb_bytes = A.a_off + NT * 128;
    A.kps = (deep && pix <= 128) ? (kps_env < 1 ? 1 : kps_env) : 1;
    A.stage_bytes = A.kps * A.kb_bytes;
    const int max_stages = NT == 64 ? (deep ? ConvCfg<64, 1>::kStages : ConvCfg<64, 0>::kStages) : (deep ? ConvCfg<128, 1>::kStages : ConvCfg<128, 0>::kStages);
    A.stages = max_stages;
    // mode-2 GroupNorm statistics are staged while the pipeline runs, in a region behind the ring that comes out of the ring's budget
    int gn_pre = 0;
    if (gn && p->gn_fuse_mode == 2) gn_pre = (b.BN * (2 * (NT / (gn->C / gn->G)) + 1) * 4 + 1023) / 1024 * 1024;
    if (deep) {
        A.stages = (ring_kb * 1024 - gn_pre) / A.stage_bytes;
        if (A.stages > max_stages) A.stages = max_stages;
        if (A.stages < 2) A.stages = 2;
    }
    // + 1 KB alignment slack + the rows past a 64-pixel box that the M = 128 MMA of the last stage still reads
    const size_t ring = (size_t)A.stages * A.stage_bytes + (A.a_off < 128 * 128 ? 128 * 128 - A.a_off : 0);
    G.smem = ring + 1024 + gn_pre;
    G.NT = NT; G.pix = pix; G.nph = nph; G.b = b; G.gn_tiles = 0;
    if (gn) {
        // ---- can this launch take the GroupNorm in its epilogue?
        const int cpg = gn->G > 0 ? gn->C / gn->G : 0;
        const bool shape_ok = gn->C == p->Cj && gn->G > 0 && gn->C % gn->G == 0 && gn->N == p->N && gn->H == p->Ho && gn->W == p->Wo &&
                              (cpg == 2 || cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) && p->Cj % 64 == 0;
        const bool tile_ok = nph == 1 && A.os == 1 && b.tiles_h == 1 && b.tiles_w == 1 && A.splits == 1;
        bool ok = shape_ok && tile_ok && gn->gamma && gn->beta && gn->sums;
        if (ok && p->gn_fuse_mode == 1) ok = gn_tensor_ok(gn->y);
        if (ok && p->gn_fuse_mode == 2)
            ok = gn_tensor_ok(gn->x) && gn_tensor_ok(gn->dx) && gn->red && !p->res.ptr && !p->bias && !p->temb &&
                 (!gn->add0.ptr || gn_tensor_ok(gn->add0)) && (!gn->add1.ptr || gn_tensor_ok(gn->add1));
        if (!ok) return -1;      // not an error by itself: dmu_conv2d_gn_fuse_supported reports 0, dmu_conv2d fails loudly
        GnEpi& E = A.gn;
        E.mode = p->gn_fuse_mode; E.G = gn->G; E.cpg = cpg; E.silu = gn->silu; E.C = gn->C;
        E.eps = gn->eps; E.cnt = (float)cpg * (float)p->Ho * (float)p->Wo;
        E.gamma = gn->gamma; E.beta = gn->beta; E.sums = gn->sums;
        E.a = reinterpret_cast<__nv_bfloat16*>(gn->y.ptr); E.a_sn = gn->y.sn; E.a_sh = gn->y.sh; E.a_sw = gn->y.sw;
        E.x = reinterpret_cast<const __nv_bfloat16*>(gn->x.ptr); E.x_sn = gn->x.sn; E.x_sh = gn->x.sh; E.x_sw = gn->x.sw;
        E.dx = reinterpret_cast<__nv_bfloat16*>(gn->dx.ptr); E.dx_sn = gn->dx.sn; E.dx_sh = gn->dx.sh; E.dx_sw = gn->dx.sw;
        E.add0 = reinterpret_cast<const __nv_bfloat16*>(gn->add0.ptr); E.a0_sn = gn->add0.sn; E.a0_sh = gn->add0.sh; E.a0_sw = gn->add0.sw;
        E.add1 = reinterpret_cast<const __nv_bfloat16*>(gn->add1.ptr); E.a1_sn = gn->add1.sn; E.a1_sh = gn->add1.sh; E.a1_sw = gn->add1.sw;
        E.red = gn->red;
        E.epi_off = (int)ring;
        G.gn_tiles = b.tiles_n;
    }
    return 0;
}

static int conv_launch(const dmu_conv_params* p, cudaStream_t stream) {
    // 3x3 stride-1 layers of 8x8 pixels and up: the persistent halo kernel (each input pixel fetched once per CTA);
    // impl 4 forces the per-tap kernel below
    if (p->gn_fuse_mode == 3) return halo_launch(p, stream);      // statistics of the output in the halo kernel's epilogue
    if (!p->gn_fuse_mode) {
        if (p->impl == 5) {
            if (nhwc_bf16_ok(p->x) && nhwc_bf16_ok(p->y) && p->w_dtype == DMU_BF16 && aligned16(p->w) && halo_t_supported(p, 1)) return halo_t_launch(p, stream);
            if (nhwc_bf16_ok(p->x) && nhwc_bf16_ok(p->y) && p->w_dtype == DMU_BF16 && aligned16(p->w) && halo_s_supported(p, 1)) return halo_s_launch(p, stream);
            DMU_REQUIRE(halo_supported(p, 1), "dmu_conv2d: impl=halo requested for an unsupported shape (3x3 stride 1 pad 1 of >= 8x8, or 4x4 stride 2 pad 1 with 64 input channels)");
            return halo_launch(p, stream);
        }
        if (narrow_head_supported(p)) return halo_launch(p, stream);
        if (p->gn_coef) return halo_launch(p, stream);
        if (p->impl != 4 && stem_supported(p)) return stem_launch(p, stream);
        if (p->impl != 4 && halo_enabled() && halo_supported(p, 0)) return halo_launch(p, stream);
        if (p->impl != 4 && halo_enabled() && halo_t_supported(p, 0)) return halo_t_launch(p, stream);
        if (p->impl != 4 && halo_enabled() && halo_s_supported(p, 0)) return halo_s_launch(p, stream);
    }
    Maps maps;
    ConvArgs A;
    ConvGeom G;
    if (int rc = conv_prepare(p, &maps, A, G)) {
        if (rc == -1) return fail("dmu_conv2d/tc: this launch cannot take the GroupNorm of gn_fuse in its epilogue (ask dmu_conv2d_gn_fuse_supported first)");
        return rc;
    }
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv_tc_kernel<64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<64, 0>::kSmem);
        cudaFuncSetAttribute(conv_tc_kernel<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<128, 0>::kSmem);
        cudaFuncSetAttribute(conv_tc_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<64, 1>::kSmem);
        cudaFuncSetAttribute(conv_tc_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<128, 1>::kSmem);
        // the GroupNorm variants never exceed ring (<= 128 KB, its statistics region included) + slack: conv_prepare sizes them
        cudaFuncSetAttribute(conv_tc_gn_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(conv_tc_gn_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(conv_tc_gn_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (cudaError_t ae = cudaGetLastError(); ae != cudaSuccess) return fail("dmu_conv2d/tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(ae));
        attr_done = true;
    }
    const dim3 cluster(1, 1, (unsigned)A.splits);
    const dim3 grid = G.grid;
    const size_t smem = G.smem;
    DMU_REQUIRE(!A.gn.mode || smem <= 160 * 1024, "dmu_conv2d/tc: GroupNorm epilogue launch needs %zu bytes of shared memory", smem);
    const bool deep = G.deep;
    cudaError_t e;
    if (A.gn.mode && A.pix256) {
        DMU_REQUIRE(deep, "dmu_conv2d/tc: 256-pixel GroupNorm launch outside a sub-wave grid");
        e = launch_pdl(conv_tc_gn_kernel<1, 1>, grid, dim3(256), smem, stream, cluster, maps, A);
    } else if (A.gn.mode) e = deep ? launch_pdl(conv_tc_gn_kernel<1, 0>, grid, dim3(256), smem, stream, cluster, maps, A)
                                   : launch_pdl(conv_tc_gn_kernel<0, 0>, grid, dim3(256), smem, stream, cluster, maps, A);
    else if (G.NT == 64) e = deep ? launch_pdl(conv_tc_kernel<64, 1>, grid, dim3(128), smem, stream, cluster, maps, A)
                                  : launch_pdl(conv_tc_kernel<64, 0>, grid, dim3(128), smem, stream, cluster, maps, A);
    else e = deep ? launch_pdl(conv_tc_kernel<128, 1>, grid, dim3(128), smem, stream, cluster, maps, A)
                  : launch_pdl(conv_tc_kernel<128, 0>, grid, dim3(128), smem, stream, cluster, maps, A);
    if (e != cudaSuccess) return fail("dmu_conv2d/tc: launch failed: %s", cudaGetErrorString(e));
    return check_launch("dmu_conv2d/tc");
}

// what dmu_conv2d_gn_fuse_supported answers: would dmu_conv2d run this layer on the per-tap kernel with the norm in its epilogue?
static int conv_gn_fuse_tiles(const dmu_conv_params* p) {
    if (p && p->gn_fuse && p->gn_fuse_mode == 3) {
        if (p->impl == 1 || p->impl == 3 || p->impl == 4 || !halo_enabled()) return 0;
        if (p->w_dtype != DMU_BF16 || !nhwc_bf16_ok(p->x) || !nhwc_bf16_ok(p->y) || (p->res.ptr && !nhwc_bf16_ok(p->res))) return 0;
        if (p->Ck % 64 != 0 || p->Cj % 64 != 0 || encode_tiled_fn() == nullptr) return 0;
        if (p->w_sk != 1 || p->w_st != p->Ck || p->w_sn != (int64_t)9 * p->Ck || !aligned16(p->w)) return 0;
        return halo_stats_supported(p) ? 1 : 0;
    }
    if (!p || !p->gn_fuse || (p->gn_fuse_mode != 1 && p->gn_fuse_mode != 2)) return 0;
    if (p->impl == 1 || p->impl == 3 || p->impl == 5 || p->gn_coef) return 0;
    if (p->w_dtype != DMU_BF16 || !nhwc_bf16_ok(p->x) || !nhwc_bf16_ok(p->y)) return 0;
    if (p->res.ptr && !nhwc_bf16_ok(p->res)) return 0;
    if (p->Ck % 64 != 0 || p->Cj % 64 != 0) return 0;
    if (p->w_sk != 1 || (p->R * p->S > 1 && p->w_st != p->Ck) || p->w_sn != (int64_t)p->R * p->S * p->Ck || !aligned16(p->w)) return 0;
    if (p->stride < 1 || p->stride > 2 || p->R * p->S > kMaxTaps) return 0;
    if (p->bias && !aligned16(p->bias)) return 0;
    if (p->temb && (!aligned16(p->temb) || p->temb_pitch % 4 != 0)) return 0;
    if (encode_tiled_fn() == nullptr) return 0;
    // the layers the halo kernel would take by its own heuristics keep their stand-alone GroupNorm
    dmu_conv_params q = *p;
    q.gn_fuse = nullptr; q.gn_fuse_mode = 0;
    if (q.impl != 4 && halo_enabled() && halo_supported(&q, 0)) return 0;
    ConvArgs A;
    ConvGeom G;
    if (conv_prepare(p, nullptr, A, G) != 0) return 0;
    return G.gn_tiles;
}

// ================================================================================================ wgrad kernel
struct WgradArgs {
    Tap taps[kMaxTaps];
    int map_h[4], map_w[4];
    int N, Hp, Wp, Ca, Cb, RS;
    int BN, BH, BW, tiles_h, tiles_w, tiles_total, tiles_per_split;
    int units;                    // RS * Cb/64 (tap, 64-channel chunk) row blocks; two per CTA
    int stages;                   // ring depth actually used (<= WgradCfg::kStages)
    float* dw; int64_t dw_sa, dw_sb, dw_st;
};

template <int NT>
struct WgradCfg {
    static constexpr int kStages = NT == 64 ? 8 : 6;   // 1 CTA per SM: 8 x 24 KB / 6 x 32 KB in flight
    static constexpr int kBlk = 64 * 128;         // 64 pixels x 64 bf16 channels
    static constexpr int kABytes = 2 * kBlk;
    static constexpr int kBBytes = (NT / 64) * kBlk;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kSmem = kStages * kStageBytes + 1024;
};

template <int NT>
__global__ void __launch_bounds__(128) wgrad_tc_kernel(const __grid_constant__ Maps maps, const __grid_constant__ WgradArgs P) {
    using Cfg = WgradCfg<NT>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_bar;
    __shared__ uint32_t s_tmem, s_issued;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = P.Cb >> 6;
    const int u0 = blockIdx.x * 2, u1 = u0 + 1;
    const bool has1 = u1 < P.units;
    const Tap t0 = P.taps[u0 / chunks];
    const Tap t1 = P.taps[(has1 ? u1 : u0) / chunks];
    const int cb0 = (u0 % chunks) * 64, cb1 = ((has1 ? u1 : u0) % chunks) * 64;
    const int a0 = blockIdx.y * NT;
    const int tile_lo = blockIdx.z * P.tiles_per_split;
    const int tile_hi = min(P.tiles_total, tile_lo + P.tiles_per_split);

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&acc_bar, 2);
        fence_mbar_init();
        s_issued = 0;
        tma_prefetch_desc(&maps.b);
    }
    if (warp == 1) tmem_alloc(&s_tmem, NT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    pdl_wait();

    auto live = [&](const Tap& t, int th0, int tw0) {
        const int h = th0 + t.dh, w = tw0 + t.dw;
        return h < P.map_h[t.map] && h + P.BH > 0 && w < P.map_w[t.map] && w + P.BW > 0;
    };

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t blk_bytes = (uint32_t)(P.BN * P.BH * P.BW) * 128u;
            int st = 0, par = 1;
            // tile coordinates advance incrementally (no divisions in the issue loops: they are latency chains of one lane)
            int iw = tile_lo % P.tiles_w, ih = (tile_lo / P.tiles_w) % P.tiles_h, in_ = tile_lo / (P.tiles_w * P.tiles_h);
            for (int tile = tile_lo; tile < tile_hi; ++tile) {
                const int tw0 = iw * P.BW, th0 = ih * P.BH, n0 = in_ * P.BN;
                if (++iw == P.tiles_w) { iw = 0; if (++ih == P.tiles_h) { ih = 0; ++in_; } }
                const bool l0 = live(t0, th0, tw0), l1 = has1 && live(t1, th0, tw0);
                if (!l0 && !l1) continue;
                mbar_wait(&empty_bar[st], par);
                uint8_t* sa = smem + st * Cfg::kStageBytes;
                mbar_arrive_expect_tx(&full_bar[st], blk_bytes * (2 + NT / 64));
                // a dead tap still issues its (fully out-of-bounds, zero-filled) load so the block holds zeros, not stale data
                tma_load_4d(sa, &maps.a[t0.map], &full_bar[st], cb0, tw0 + t0.dw, th0 + t0.dh, n0);
                tma_load_4d(sa + Cfg::kBlk, &maps.a[t1.map], &full_bar[st], cb1, tw0 + t1.dw, th0 + t1.dh, n0);
#pragma unroll
                for (int q = 0; q < NT / 64; ++q)
                    tma_load_4d(sa + Cfg::kABytes + q * Cfg::kBlk, &maps.b, &full_bar[st], a0 + q * 64, tw0, th0, n0);
                if (++st == P.stages) { st = 0; par ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, NT, 1, 1);   // both operands MN-major (channels contiguous, K = pixels)
            const int ksteps = (P.BN * P.BH * P.BW + 15) >> 4;             // K = 16 pixels per MMA; a box holds <= 64 pixels
            const uint32_t smem0 = smem_u32(smem);
            int it = 0, st = 0, par = 0;
            int iw = tile_lo % P.tiles_w, ih = (tile_lo / P.tiles_w) % P.tiles_h;
            const uint64_t d_ring = smem_desc_sw128(smem0, Cfg::kBlk, 1024);
            uint64_t da = d_ring;
            for (int tile = tile_lo; tile < tile_hi; ++tile) {
                const int tw0 = iw * P.BW, th0 = ih * P.BH;
                if (++iw == P.tiles_w) { iw = 0; if (++ih == P.tiles_h) ih = 0; }
                const bool l0 = live(t0, th0, tw0), l1 = has1 && live(t1, th0, tw0);
                if (!l0 && !l1) continue;
                mbar_wait(&full_bar[st], par);
                tc_fence_after();
                const uint64_t db = da + (uint64_t)(Cfg::kABytes >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)   // K = 16 pixels = 16 rows of 128 B = 2048 B per step
                    if (k < ksteps) umma_bf16(tmem, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (it | k) != 0);
                umma_commit(&empty_bar[st]);
                ++it;
                da += (uint64_t)(Cfg::kStageBytes >> 4);
                if (++st == P.stages) { st = 0; par ^= 1; da = d_ring; }
            }
            s_issued = (uint32_t)it;
            umma_commit(&acc_bar);      // arrival 1: all MMAs retired
            mbar_arrive(&acc_bar);      // arrival 2: release-publishes s_issued to the epilogue threads
        }
        __syncwarp();
    }
    __syncwarp();

    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    const bool have_acc = *reinterpret_cast<volatile uint32_t*>(&s_issued) != 0;
    // thread = accumulator row m = (block, channel b); columns = channels a of P
    const int blk = threadIdx.x >> 6, bl = threadIdx.x & 63;
    const bool row_ok = have_acc && (blk == 0 || has1);
    const int tap = (blk == 0 ? u0 : u1) / chunks;
    const int b = (blk == 0 ? cb0 : cb1) + bl;
    float* dwp = P.dw + (int64_t)b * P.dw_sb + (int64_t)tap * P.dw_st;
#pragma unroll 1
    for (int c = 0; c < NT; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(dwp + (int64_t)(a0 + c + i) * P.dw_sa, v[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, NT);
}

static int wgrad_supported(const dmu_wgrad_params* p) {
    if (!p || !p->p.ptr || !p->q.ptr || !p->dw) return 0;
    if (!nhwc_bf16_ok(p->p) || !nhwc_bf16_ok(p->q)) return 0;
    if (p->Ca % 64 != 0 || p->Cb % 64 != 0) return 0;
    if (p->stride < 1 || p->stride > 2 || p->R * p->S > kMaxTaps) return 0;
    if (encode_tiled_fn() == nullptr) return 0;
    return 1;
}

// conv_wgrad_halo.cu: 3x3 stride-1 layers with many position tiles (both operands read once per tap group)
int wgrad_halo_supported(const dmu_wgrad_params* p, int force);
int wgrad_halo_launch(const dmu_wgrad_params* p, cudaStream_t stream);

static int wgrad_launch(const dmu_wgrad_params* p, cudaStream_t stream) {
    if (wgrad_halo_supported(p, p->impl == 5 ? 1 : 0)) return wgrad_halo_launch(p, stream);
    Maps maps;
    WgradArgs A;
    memset(&A, 0, sizeof(A));
    Box b = make_box(p->N, p->Hp, p->Wp, 64);
    // the UMMA K step is 16 pixels: the box must hold a multiple of 16 pixels (tiny batches of tiny images: pad the image
    // count of the box, TMA zero-fills the images past N)
    if (b.BH * b.BW < 16) {
        const int mult = 16 / (b.BH * b.BW);
        b.BN = (b.BN + mult - 1) / mult * mult;
        b.tiles_n = (p->N + b.BN - 1) / b.BN;
    }
    if (int rc = build_gather0(p->q, p->N, p->Hq, p->Wq, p->Cb, p->R, p->S, p->stride, p->pad, b, A.taps, &maps, A.map_h, A.map_w, "dmu_conv2d_wgrad/tc"))
        return rc;
    {
        const uint64_t dims[4] = {(uint64_t)p->Ca, (uint64_t)p->Wp, (uint64_t)p->Hp, (uint64_t)p->N};
        const uint64_t str[4] = {1, (uint64_t)p->p.sw, (uint64_t)p->p.sh, (uint64_t)p->p.sn};
        const uint32_t box[4] = {64, (uint32_t)b.BW, (uint32_t)b.BH, (uint32_t)b.BN};
        if (int rc = make_map_bf16(&maps.b, p->p.ptr, 4, dims, str, box, "dmu_conv2d_wgrad/tc P")) return rc;
    }
    const int NT = (p->Ca % 128 == 0) ? 128 : 64;
    A.N = p->N; A.Hp = p->Hp; A.Wp = p->Wp; A.Ca = p->Ca; A.Cb = p->Cb; A.RS = p->R * p->S;
    A.BN = b.BN; A.BH = b.BH; A.BW = b.BW; A.tiles_h = b.tiles_h; A.tiles_w = b.tiles_w;
    A.tiles_total = b.tiles_n * b.tiles_h * b.tiles_w;
    A.units = A.RS * (p->Cb / 64);
    A.dw = p->dw; A.dw_sa = p->dw_sa; A.dw_sb = p->dw_sb; A.dw_st = p->dw_st;
    const int mt = (A.units + 1) / 2, ntl = p->Ca / NT;
    // Grid size: the weight gradients run on the side lane of the backward graph, next to the dgrad / GroupNorm chain.  One wave
    // of CTAs with a shallow ring (see wg_cap below) shares each SM with a CTA of the chain; two waves of 193 KB CTAs held every
    // SM and the chain's launches queued behind them (30.7k -> 34.8k img/s with the two changes together).
    static int target_ctas = -1;
    if (target_ctas < 0) { const char* e = getenv("DMU_WGRAD_CTAS"); target_ctas = e ? atoi(e) : sm_count(); }
    int splits = (target_ctas + mt * ntl - 1) / (mt * ntl);
    if (splits > A.tiles_total) splits = A.tiles_total;
    if (splits < 1) splits = 1;
    A.tiles_per_split = (A.tiles_total + splits - 1) / splits;
    splits = (A.tiles_total + A.tiles_per_split - 1) / A.tiles_per_split;
    dim3 grid(mt, ntl, splits);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradCfg<64>::kSmem);
        cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradCfg<128>::kSmem);
        attr_done = true;
    }
    // Measured (B=128, 32x32 step): three stages (72-96 KB) instead of 8 / 6 (193 KB) cost the weight gradients nothing - they are
    // L2-bound, not latency-bound - and let a CTA of the dgrad / GroupNorm chain share the SM: backward 2.18 -> 2.03 ms.
    static const int wg_cap = [] { const char* e = getenv("DMU_WGRAD_STAGES"); return e ? atoi(e) : 3; }();
    const int full = NT == 64 ? WgradCfg<64>::kStages : WgradCfg<128>::kStages;
    A.stages = wg_cap < full ? (wg_cap < 2 ? 2 : wg_cap) : full;
    const size_t smem = (size_t)A.stages * (NT == 64 ? WgradCfg<64>::kStageBytes : WgradCfg<128>::kStageBytes) + 1024;
    cudaError_t e = NT == 64 ? launch_pdl(wgrad_tc_kernel<64>, grid, dim3(128), smem, stream, dim3(1, 1, 1), maps, A)
                             : launch_pdl(wgrad_tc_kernel<128>, grid, dim3(128), smem, stream, dim3(1, 1, 1), maps, A);
    if (e != cudaSuccess) return fail("dmu_conv2d_wgrad/tc: launch failed: %s", cudaGetErrorString(e));
    if (int rc = check_launch("dmu_conv2d_wgrad/tc")) return rc;
    if (p->dbias) return dmu_colsum(&p->p, p->N, p->Hp, p->Wp, p->Ca, nullptr, 0, p->dbias, 1.0f, (dmu_stream_t)stream);
    return 0;
}

}  // namespace tc
}  // namespace dmu

using namespace dmu;

extern "C" {
int64_t dmu_conv2d_workspace_bytes(void) { return tc::kSplitWsBytes; }
// development aid, not part of include/dmu_b200.h: int64 device buffer (8 per CTA) receiving clock64 stamps of conv_tc_kernel
void dmu_debug_set_buffer(void* p) { tc::g_debug_buffer = reinterpret_cast<long long*>(p); }
int dmu_conv2d_tc_supported(const dmu_conv_params* p) { return tc::conv_supported(p); }
int dmu_conv2d_gn_supported(const dmu_conv_params* p) {
    if (!p || !p->x.ptr || !p->y.ptr || !p->w || p->impl == 1 || p->impl == 3 || p->impl == 4) return 0;
    if (!tc::nhwc_bf16_ok(p->x) || p->Ck % 64 != 0 || tc::encode_tiled_fn() == nullptr || !tc::halo_enabled()) return 0;
    if (!tc::conv_supported(p)) return 0;
    return tc::halo_supported(p, p->impl == 5 ? 1 : 0);
}
int dmu_conv2d_tc(const dmu_conv_params* p, dmu_stream_t stream) { return tc::conv_launch(p, as_stream(stream)); }
int dmu_conv2d_gn_fuse_supported(const dmu_conv_params* p) { return tc::conv_gn_fuse_tiles(p); }
int dmu_wgrad_tc_supported(const dmu_wgrad_params* p) { return tc::wgrad_supported(p); }
int dmu_wgrad_tc(const dmu_wgrad_params* p, dmu_stream_t stream) { return tc::wgrad_launch(p, as_stream(stream)); }
}
