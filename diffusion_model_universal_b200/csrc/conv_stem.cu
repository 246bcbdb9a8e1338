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

struct StemMaps { CUtensorMap y; };

struct StemArgs {
    dmu_tensor4 x;            // few-channel input, any strides / dtype
    const void* w; int64_t w_sn, w_sk, w_st; int w_dtype;
    const float* bias;
    __nv_bfloat16* y; int64_t y_sn, y_sh, y_sw;
    int N, H, W, Ck, Cj, flip, ksteps;
    int tiles, pixels;
};

// byte offset of element (row, k) in a 128-row K-major SWIZZLE_128B tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
    return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1)));
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(128) stem_tc_kernel(const __grid_constant__ StemMaps maps, const StemArgs P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_a = smem;                    // 128 pixels x 64 bf16
    uint8_t* s_b = smem + 128 * 128;        // 64 output channels x 64 bf16
    uint8_t* s_o = s_b + 64 * 128;          // output tile, 128 pixels x 64 bf16, SWIZZLE_128B (what the TMA store reads)
    __shared__ __align__(8) uint64_t acc_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[64];
    const int warp = threadIdx.x >> 5;
    const int j0 = blockIdx.y * 64;
    const int K = 9 * P.Ck;

    pdl_trigger();
    if (threadIdx.x == 0) { mbar_init(&acc_bar, 1); fence_mbar_init(); tma_prefetch_desc(&maps.y); }
    if (warp == 1) tmem_alloc(&s_tmem, 64);
    // zero both operand tiles once (columns >= K stay zero for the whole kernel), then the filter bank: parameters, not
    // produced by the previous launch, so all of this overlaps its tail
    for (int i = threadIdx.x; i < (128 + 64) * 128 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);   // s_a, s_b
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
    const int HW = P.H * P.W;
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        const int m = tile * 128 + row;
        const bool valid = m < P.pixels;
        const int n = m / HW;
        const int hw = m - n * HW;
        const int h = hw / P.W, w = hw - h * P.W;
        // ---- im2col row of this pixel: taps x channels, out-of-image taps are zero.  The dtype switch sits OUTSIDE the loads
        //      (with it inside, every load was its own branch and the 27 of them were serialised: ~16k clk per tile).
        if (valid) {
            const int64_t base = (int64_t)n * P.x.sn + (int64_t)h * P.x.sh + (int64_t)w * P.x.sw;
            bool ok[9];
            int64_t off[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int dr = P.flip ? 1 - t / 3 : t / 3 - 1, ds = P.flip ? 1 - t % 3 : t % 3 - 1;
                ok[t] = h + dr >= 0 && h + dr < P.H && w + ds >= 0 && w + ds < P.W;
                off[t] = base + (int64_t)dr * P.x.sh + (int64_t)ds * P.x.sw;
            }
            if (P.x.dtype == DMU_F32) {
                const float* xp = reinterpret_cast<const float*>(P.x.ptr);
                for (int c = 0; c < P.Ck; ++c) {
                    float v[9];
#pragma unroll
                    for (int t = 0; t < 9; ++t) v[t] = ok[t] ? __ldg(xp + off[t] + (int64_t)c * P.x.sc) : 0.f;
#pragma unroll
                    for (int t = 0; t < 9; ++t) *reinterpret_cast<__nv_bfloat16*>(s_a + sw128_off(row, t * P.Ck + c)) = __float2bfloat16_rn(v[t]);
                }
            } else {
                const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(P.x.ptr);
                for (int c = 0; c < P.Ck; ++c) {
                    __nv_bfloat16 v[9];
#pragma unroll
                    for (int t = 0; t < 9; ++t) v[t] = ok[t] ? xp[off[t] + (int64_t)c * P.x.sc] : __float2bfloat16_rn(0.f);
#pragma unroll
                    for (int t = 0; t < 9; ++t) *reinterpret_cast<__nv_bfloat16*>(s_a + sw128_off(row, t * P.Ck + c)) = v[t];
                }
            }
        }
        __syncwarp();
        if (warp == 0 && elect_one()) tma_store_wait_read();     // the previous tile's store has finished reading s_o ...
        fence_proxy_async();       // generic-proxy writes of the operand tile -> visible to the tensor core's async proxy
        __syncthreads();           // ... before anyone passes this barrier and rewrites it in the epilogue below
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
        // ---- epilogue: + bias, bf16, into the swizzled staging tile; ONE TMA store writes the 16 KB tile (rows past the last
        //      pixel are clipped by the tensor map), instead of 128 threads x 8 scattered 16-byte stores
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += s_bias[c + i];
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                const int chunk = (c + i) >> 3;
                store_vec<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(s_o + row * 128 + ((chunk ^ (row & 7)) << 4)), v + i);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();           // staging tile complete, accumulator drained, operand tile free
        tc_fence_after();
        if (warp == 0 && elect_one()) {
            tma_store_2d(&maps.y, s_o, j0, tile * 128);
            tma_store_commit();
        }
    }
    if (warp == 0 && elect_one()) tma_store_wait_all();
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
    if (p->y.sh != (int64_t)p->Wo * p->y.sw || p->y.sn != (int64_t)p->Ho * p->y.sh) return 0;     // pixel-contiguous: one 2-D store map
    if ((int64_t)p->N * p->Hi * p->Wi >= (1ll << 31) - 256 || encode_tiled_fn() == nullptr) return 0;
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
    A.pixels = p->N * p->Hi * p->Wi;
    A.tiles = (A.pixels + 127) / 128;
    StemMaps maps;
    {
        const uint64_t dims[2] = {(uint64_t)p->Cj, (uint64_t)A.pixels};
        const uint64_t str[2] = {1, (uint64_t)p->y.sw};
        const uint32_t box[2] = {64, 128};
        if (int rc = make_map_bf16(&maps.y, p->y.ptr, 2, dims, str, box, "dmu_conv2d/stem output")) return rc;
    }
    const int smem = (128 + 64 + 128) * 128 + 1024;
    const int ny = p->Cj / 64;
    int64_t gx = (int64_t)4 * sm_count() / ny;          // ~4 co-resident CTAs per SM hide each other's gather latency
    if (gx > A.tiles) gx = A.tiles;
    if (gx < 1) gx = 1;
    cudaError_t e = launch_pdl(stem_tc_kernel, dim3((unsigned)gx, ny), dim3(128), (size_t)smem, stream, dim3(1, 1, 1), maps, A);
    if (e != cudaSuccess) return fail("dmu_conv2d/stem: launch failed: %s", cudaGetErrorString(e));
    return check_launch("dmu_conv2d/stem");
}

}  // namespace tc
}  // namespace dmu
