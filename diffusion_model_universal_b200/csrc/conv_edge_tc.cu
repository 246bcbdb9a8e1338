// Weight gradients of the two 3-channel boundary layers on the tensor cores (stem 3 -> C, head C -> 3; models/ddpm.py:49,90):
//   out[c][tap][j] += sum_pix wide[pix, c] * narrow[pix + sgn * (tap - pad), j]         (same contract as narrow_wgrad_kernel)
// Default for the bf16 path since round 2 (DMU_EDGE_WGRAD_TC=0 falls back to the SIMT kernel in conv_edge.cu): measured inside the
// training step it is not faster alone (29 - 59 us next to the other lane's kernels) but it runs 128 threads x 56 registers,
// where the SIMT kernel's two 256-thread x 125-register CTAs per SM fill the register file for ~45 us and stall everything queued
// beside it - the embedding backward at the end of the step, the first dgrads at its start (48.4k vs 47.7k img/s).
//
// Shape: the pixels are the contraction axis.  Per 64-pixel k-block
//   B operand  = wide[64 pixels][64 channels] straight from NHWC by one TMA box (MN-major, as the per-tap wgrad kernel uses it),
//   A operand  = the im2col rows of the narrow tensor, [m = tap * Cn + j][64 pixels] K-major SWIZZLE_128B, written by the CTA's
//                threads: rows 0 .. 9 Cn - 1 hold the bf16 value, rows 9 Cn + 1 .. 18 Cn the bf16 remainder of an fp32 narrow
//                tensor (x - bf16(x): the pair carries ~16 mantissa bits, so the network input / dL/d(eps) are not rounded to
//                bf16 here; both rows are added into the same output), row 9 Cn holds ones, so D[9 Cn][c] = sum_pix wide[pix, c]
//                = the bias gradient of the conv whose output gradient is `wide`, for free; the remaining rows stay zero,
//   D[128 x 64] += A * B with four M = 128, N = 64, K = 16 tcgen05.mma.
// Only descriptor / instruction configurations that other kernels of this library already run are used: K-major M = 128 A
// (conv_stem.cu), MN-major N = 64 B (wgrad_tc_kernel).  Each CTA reduces a contiguous range of k-blocks and ends with one pass of
// (18 * Cn + 1) x 64 atomics.  wide is read once (16.8 MB at B = 128, 32 x 32, C = 64: 2.6 us of HBM).
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace dmu {
namespace tc {

struct EdgeWgradMaps { CUtensorMap wide; };

struct EdgeWgradArgs {
    dmu_tensor4 narrow;
    float* out; int64_t o_c, o_t, o_j;
    float* dbias;        // [Cw] += sum_pix wide      (ones row)
    float* dbias_n;      // [Cn] += sum_pix narrow    (fp32, by the gathering threads)
    int H, W, Cn, sgn;
    int pixels, kblocks, per_cta;
};

constexpr int kEwA = 128 * 128;      // im2col tile: 128 rows x 64 bf16
constexpr int kEwB = 64 * 128;       // wide tile: 64 pixels x 64 bf16
constexpr int kEwStage = kEwA + kEwB;
constexpr int kEwStages = 2;

__device__ __forceinline__ uint32_t ew_sw128_off(int row, int k) {
    return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1)));
}

template <typename TN>
__global__ void __launch_bounds__(128) edge_wgrad_tc_kernel(const __grid_constant__ EdgeWgradMaps maps, const EdgeWgradArgs P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[kEwStages], empty_bar[kEwStages], acc_bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.y * 64;
    const int blk_lo = blockIdx.x * P.per_cta;
    const int blk_hi = min(P.kblocks, blk_lo + P.per_cta);
    const int nblk = blk_hi - blk_lo;                      // >= 1 by construction of the grid
    const int rows = 9 * P.Cn;                             // im2col rows; row `rows` = ones

    if (threadIdx.x == 0) {
        for (int i = 0; i < kEwStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
        tma_prefetch_desc(&maps.wide);
    }
    if (warp == 1) tmem_alloc(&s_tmem, 64);
    // both im2col tiles start as zeros: rows past `rows` are never written again
    for (int s = 0; s < kEwStages; ++s)
        for (int i = threadIdx.x; i < kEwA / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + s * kEwStage)[i] = make_uint4(0u, 0u, 0u, 0u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    const int p = threadIdx.x & 63, half = threadIdx.x >> 6;   // pixel of the k-block; taps half, half + 2, ...
    const int HW = P.H * P.W;
    const TN* np = reinterpret_cast<const TN*>(P.narrow.ptr);
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);     // A K-major, B MN-major
    float nsum[4] = {0.f, 0.f, 0.f, 0.f};

    // (measured: fetching the narrow values one k-block ahead does not shorten the launch and doubles its registers - the loop is
    // bound by the barrier + MMA round trip per k-block, not by the gather)
    for (int i = 0; i < nblk; ++i) {
        const int st = i & 1;
        uint8_t* s_a = smem + st * kEwStage;
        uint8_t* s_b = s_a + kEwA;
        if (i >= kEwStages) mbar_wait(&empty_bar[st], (uint32_t)(((i >> 1) - 1) & 1));     // MMAs of block i - 2 have read this stage
        const int pix0 = (blk_lo + i) * 64;
        if (warp == 0) {
            if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[st], kEwB);
                tma_load_2d(s_b, &maps.wide, &full_bar[st], c0, pix0);               // rows past the last pixel arrive as zeros
            }
            __syncwarp();
        }
        // ---- im2col rows of this thread's pixel
        const int gp = pix0 + p;
        const bool valid = gp < P.pixels;
        const int n = gp / HW, hw = gp - n * HW;
        const int h = hw / P.W, w = hw - h * P.W;
        const int64_t base = (int64_t)n * P.narrow.sn + (int64_t)h * P.narrow.sh + (int64_t)w * P.narrow.sw;
        for (int t = half; t < 9; t += 2) {
            const int dr = P.sgn * (t / 3 - 1), ds = P.sgn * (t % 3 - 1);
            const bool ok = valid && h + dr >= 0 && h + dr < P.H && w + ds >= 0 && w + ds < P.W;
            const int64_t off = base + (int64_t)dr * P.narrow.sh + (int64_t)ds * P.narrow.sw;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (ok && j < P.Cn) ? (float)np[off + (int64_t)j * P.narrow.sc] : 0.f;
            if (t == 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) nsum[j] += v[j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < P.Cn) {
                    const __nv_bfloat16 hi = __float2bfloat16_rn(v[j]);
                    *reinterpret_cast<__nv_bfloat16*>(s_a + ew_sw128_off(t * P.Cn + j, p)) = hi;
                    if constexpr (sizeof(TN) == 4)
                        *reinterpret_cast<__nv_bfloat16*>(s_a + ew_sw128_off(rows + 1 + t * P.Cn + j, p)) = __float2bfloat16_rn(v[j] - __bfloat162float(hi));
                }
        }
        if (half == 1) *reinterpret_cast<__nv_bfloat16*>(s_a + ew_sw128_off(rows, p)) = __float2bfloat16_rn(valid ? 1.f : 0.f);
        fence_proxy_async();       // generic-proxy writes of the im2col tile -> visible to the tensor core's async proxy
        __syncthreads();
        if (warp == 0) {
            if (elect_one()) {
                mbar_wait(&full_bar[st], (uint32_t)((i >> 1) & 1));
                tc_fence_after();
                const uint64_t da = smem_desc_sw128(smem_u32(s_a), 16, 1024);        // K-major: 32 B per K = 16 step
                const uint64_t db = smem_desc_sw128(smem_u32(s_b), kEwB, 1024);      // MN-major: 16 pixel rows = 2048 B per step
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(128 * k), idesc, (i | k) != 0);
                umma_commit(&empty_bar[st]);
                if (i == nblk - 1) umma_commit(&acc_bar);
            }
            __syncwarp();
        }
    }
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    // ---- epilogue: accumulator row m sits in TMEM lane m; rows 0 .. 2 * rows live in the quadrants of warps 0 and 1 (2 * rows + 1 <= 64)
    if (warp < 2) {
        const int L = warp * 32 + lane;
        const int m = L < rows ? L : ((L > rows && L <= 2 * rows) ? L - rows - 1 : -1);     // value row | remainder row of the same (tap, j)
#pragma unroll 1
        for (int c = 0; c < 64; c += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            tmem_ld_wait();
            if (m >= 0) {
                const int tap = m / P.Cn, j = m - tap * P.Cn;
                float* o = P.out + (int64_t)tap * P.o_t + (int64_t)j * P.o_j;
#pragma unroll
                for (int q = 0; q < 32; ++q) atomicAdd(o + (int64_t)(c0 + c + q) * P.o_c, v[q]);
            } else if (L == rows && P.dbias) {
#pragma unroll
                for (int q = 0; q < 32; ++q) atomicAdd(P.dbias + c0 + c + q, v[q]);
            }
        }
    }
    if (P.dbias_n && blockIdx.y == 0 && half == 0) {        // warps 0 and 1 hold the centre tap
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = nsum[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && j < P.Cn) atomicAdd(P.dbias_n + j, s);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 64);
}

static bool edge_wgrad_tc_enabled() {
    static const int v = [] { const char* e = getenv("DMU_EDGE_WGRAD_TC"); return e ? atoi(e) : 1; }();
    return v != 0;
}

// wide: bf16 NHWC, pixel-contiguous (one 2-D map), Cw % 64 == 0; narrow: 1..3 channels (18 * Cn + 1 <= 64), fp32 or bf16, any strides
int edge_wgrad_tc_supported(const dmu_tensor4* wide, const dmu_tensor4* narrow, int N, int H, int W, int Cw, int Cn) {
    if (!edge_wgrad_tc_enabled() || !wide || !narrow || !wide->ptr || !narrow->ptr) return 0;
    if (Cn < 1 || 18 * Cn + 1 > 64 || Cw % 64 != 0 || Cw < 64) return 0;
    if (wide->dtype != DMU_BF16 || wide->sc != 1 || wide->sw % 8 || (reinterpret_cast<uintptr_t>(wide->ptr) & 15)) return 0;
    if (wide->sh != (int64_t)W * wide->sw || wide->sn != (int64_t)H * wide->sh) return 0;
    if (narrow->dtype != DMU_F32 && narrow->dtype != DMU_BF16) return 0;
    if ((int64_t)N * H * W >= (1ll << 31) - 256 || encode_tiled_fn() == nullptr) return 0;
    return 1;
}

int edge_wgrad_tc_launch(const dmu_tensor4* wide, const dmu_tensor4* narrow, int N, int H, int W, int Cw, int Cn, int sgn, float* out,
                         int64_t o_c, int64_t o_t, int64_t o_j, float* dbias, float* dbias_n, cudaStream_t stream) {
    EdgeWgradArgs A;
    memset(&A, 0, sizeof(A));
    A.narrow = *narrow;
    A.out = out; A.o_c = o_c; A.o_t = o_t; A.o_j = o_j;
    A.dbias = dbias; A.dbias_n = dbias_n;
    A.H = H; A.W = W; A.Cn = Cn; A.sgn = sgn;
    A.pixels = N * H * W;
    A.kblocks = (A.pixels + 63) / 64;
    EdgeWgradMaps maps;
    {
        const uint64_t dims[2] = {(uint64_t)Cw, (uint64_t)A.pixels};
        const uint64_t str[2] = {1, (uint64_t)wide->sw};
        const uint32_t box[2] = {64, 64};
        if (int rc = make_map_bf16(&maps.wide, wide->ptr, 2, dims, str, box, "dmu_conv2d_wgrad/edge_tc wide")) return rc;
    }
    static const int per_sm = [] { const char* e = getenv("DMU_EDGE_WGRAD_CTAS"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
    const int ny = Cw / 64;
    int gx = per_sm * sm_count() / ny;
    if (gx < 1) gx = 1;
    if (gx > A.kblocks) gx = A.kblocks;
    A.per_cta = (A.kblocks + gx - 1) / gx;
    gx = (A.kblocks + A.per_cta - 1) / A.per_cta;          // every CTA owns at least one k-block
    const int smem = kEwStages * kEwStage + 1024;
    if (narrow->dtype == DMU_F32) {
        static const cudaError_t attr = cudaFuncSetAttribute(edge_wgrad_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEwStages * kEwStage + 1024);
        (void)attr;
        edge_wgrad_tc_kernel<float><<<dim3((unsigned)gx, ny), 128, smem, stream>>>(maps, A);
    } else {
        static const cudaError_t attr = cudaFuncSetAttribute(edge_wgrad_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEwStages * kEwStage + 1024);
        (void)attr;
        edge_wgrad_tc_kernel<__nv_bfloat16><<<dim3((unsigned)gx, ny), 128, smem, stream>>>(maps, A);
    }
    return check_launch("dmu_conv2d_wgrad/edge_tc");
}

}  // namespace tc
}  // namespace dmu
