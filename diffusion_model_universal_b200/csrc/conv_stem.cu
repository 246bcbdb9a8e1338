// The 3-channel boundary layers on the tensor cores: stem conv 3 -> C (fprop) and head conv C -> 3 (dgrad, the same
// arithmetic with mirrored taps), 3x3 / stride 1 / pad 1, reference call sites models/ddpm.py:49,90.
//
// K = 9 * Ck <= 64 is too thin for a TMA-fed pipeline, and the SIMT kernel in conv_edge.cu is issue-bound (33 us for a
// 16.8 MB output that HBM writes in 2.6 us).  Here the 128 threads of a CTA build the im2col tile of 128 consecutive output
// pixels directly in shared memory in the K-major SWIZZLE_128B operand layout (one row of <= 64 bf16 per pixel, read from
// the few-channel input through its strides: NCHW fp32 at the API boundary), the filter bank sits next to it, one elected
// lane issues ceil(K / 16) tcgen05.mma (M = 128, N = 64), and the epilogue adds the bias and writes bf16 NHWC rows.  Several
// CTAs share an SM, so one CTA's gather latency hides behind the others' stores.  The input is rounded to bf16 like every
// other activation of the bf16 path.
#include <string.h>

#include "tc_common.cuh"

namespace dmu {
namespace tc {

struct StemArgs {
    dmu_tensor4 x;            // few-channel input, any strides / dtype
    const void* w; int64_t w_sn, w_sk, w_st; int w_dtype;
    const float* bias;
    __nv_bfloat16* y; int64_t y_sn, y_sh, y_sw;
    int N, H, W, Ck, Cj, flip, ksteps;
    int64_t tiles;
};

// byte offset of element (row, k) in a 128-row K-major SWIZZLE_128B tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
    return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1)));
}

__global__ void __launch_bounds__(128) stem_tc_kernel(const StemArgs P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_a = smem;                    // 128 pixels x 64 bf16
    uint8_t* s_b = smem + 128 * 128;        // 64 output channels x 64 bf16
    __shared__ __align__(8) uint64_t acc_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[64];
    const int warp = threadIdx.x >> 5;
    const int j0 = blockIdx.y * 64;
    const int K = 9 * P.Ck;

    pdl_trigger();
    if (threadIdx.x == 0) { mbar_init(&acc_bar, 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(&s_tmem, 64);
    // zero both operand tiles once (columns >= K stay zero for the whole kernel), then the filter bank: parameters, not
    // produced by the previous launch, so all of this overlaps its tail
    for (int i = threadIdx.x; i < (128 + 64) * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < 64) s_bias[threadIdx.x] = P.bias ? P.bias[j0 + threadIdx.x] : 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * K; i += blockDim.x) {
        const int j = i / K, k = i % K;
        const int tap = k / P.Ck, c = k % P.Ck;
        const float v = ld_as_float(P.w, (int64_t)(j0 + j) * P.w_sn + (int64_t)c * P.w_sk + (int64_t)tap * P.w_st, P.w_dtype);
        *reinterpret_cast<__nv_bfloat16*>(s_b + sw128_off(j, k)) = __float2bfloat16_rn(v);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    pdl_wait();

    const int row = threadIdx.x;
    const int64_t HW = (int64_t)P.H * P.W;
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        const int64_t m = tile * 128 + row;
        const bool valid = m < (int64_t)P.N * HW;
        const int n = (int)(m / HW);
        const int hw = (int)(m - (int64_t)n * HW);
        const int h = hw / P.W, w = hw - h * P.W;
        // ---- im2col row of this pixel: taps x channels, out-of-image taps are zero
        if (valid) {
            for (int c = 0; c < P.Ck; ++c) {
                float v[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int r = t / 3, s = t % 3;
                    const int hh = h + (P.flip ? 1 - r : r - 1), ww = w + (P.flip ? 1 - s : s - 1);
                    v[t] = (hh >= 0 && hh < P.H && ww >= 0 && ww < P.W)
                               ? ld_as_float(P.x.ptr, (int64_t)n * P.x.sn + (int64_t)hh * P.x.sh + (int64_t)ww * P.x.sw + (int64_t)c * P.x.sc, P.x.dtype)
                               : 0.f;
                }
#pragma unroll
                for (int t = 0; t < 9; ++t) *reinterpret_cast<__nv_bfloat16*>(s_a + sw128_off(row, t * P.Ck + c)) = __float2bfloat16_rn(v[t]);
            }
        }
        fence_proxy_async();       // generic-proxy writes of the operand tile -> visible to the tensor core's async proxy
        __syncthreads();
        if (warp == 0) {
            if (elect_one()) {
                const uint64_t da = smem_desc_sw128(smem_u32(s_a), 16, 1024), db = smem_desc_sw128(smem_u32(s_b), 16, 1024);
                for (int k = 0; k < P.ksteps; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
                umma_commit(&acc_bar);
            }
            __syncwarp();
        }
        mbar_wait(&acc_bar, phase);
        phase ^= 1;
        tc_fence_after();
        __nv_bfloat16* yp = P.y + (int64_t)n * P.y_sn + (int64_t)h * P.y_sh + (int64_t)w * P.y_sw + j0;
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += s_bias[c + i];
#pragma unroll
                for (int i = 0; i < 32; i += 8) store_vec<__nv_bfloat16>(yp + c + i, v + i);
            }
        }
        tc_fence_before();
        __syncthreads();           // accumulator drained and operand tile free before the next tile overwrites them
        tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 64);
}

int stem_supported(const dmu_conv_params* p) {
    if (!p || !p->x.ptr || !p->y.ptr || !p->w) return 0;
    if (p->R != 3 || p->S != 3 || p->stride != 1 || p->pad != 1 || p->Hi != p->Ho || p->Wi != p->Wo) return 0;
    if (p->Ck < 1 || 9 * p->Ck > 64 || p->Cj % 64 != 0) return 0;
    if (p->temb || p->res.ptr) return 0;
    if (p->y.dtype != DMU_BF16 || p->y.sc != 1 || p->y.sw % 8 || p->y.sh % 8 || p->y.sn % 8 || (reinterpret_cast<uintptr_t>(p->y.ptr) & 15)) return 0;
    return 1;
}

int stem_launch(const dmu_conv_params* p, cudaStream_t stream) {
    StemArgs A;
    memset(&A, 0, sizeof(A));
    A.x = p->x;
    A.w = p->w; A.w_sn = p->w_sn; A.w_sk = p->w_sk; A.w_st = p->w_st; A.w_dtype = p->w_dtype;
    A.bias = p->bias;
    A.y = reinterpret_cast<__nv_bfloat16*>(p->y.ptr); A.y_sn = p->y.sn; A.y_sh = p->y.sh; A.y_sw = p->y.sw;
    A.N = p->N; A.H = p->Hi; A.W = p->Wi; A.Ck = p->Ck; A.Cj = p->Cj; A.flip = p->gather;
    A.ksteps = (9 * p->Ck + 15) / 16;
    A.tiles = ((int64_t)p->N * p->Hi * p->Wi + 127) / 128;
    const int smem = (128 + 64) * 128 + 1024;
    const int ny = p->Cj / 64;
    int64_t gx = (int64_t)4 * sm_count() / ny;          // ~4 co-resident CTAs per SM hide each other's gather latency
    if (gx > A.tiles) gx = A.tiles;
    if (gx < 1) gx = 1;
    cudaError_t e = launch_pdl(stem_tc_kernel, dim3((unsigned)gx, ny), dim3(128), (size_t)smem, stream, dim3(1, 1, 1), A);
    if (e != cudaSuccess) return fail("dmu_conv2d/stem: launch failed: %s", cudaGetErrorString(e));
    return check_launch("dmu_conv2d/stem");
}

}  // namespace tc
}  // namespace dmu
