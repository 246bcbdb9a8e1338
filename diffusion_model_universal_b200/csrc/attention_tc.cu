// Multi-head self-attention core on the tensor cores (models/layers/attention.py:49-61), forward, bf16, head dim 32 or 64,
// sequences S | 128 (the UNet's attention stages: S = 64, 16, 4, 1).
//
// One CTA = 128 consecutive rows of the [N * S, 3C] qkv matrix (128 / S whole images) x one head:
//   scores  D1[128 x 128] = Q[128 x d] K^T[d x 128]      tcgen05.mma, M = 128, N = 128, K = 16 x d/16, both operands K-major:
//                                                        the Q and K tiles are ONE TMA box each (64 channels x 128 rows, SWIZZLE_128B;
//                                                        a 32-wide head is one half of the 128-byte rows: a 64-byte start offset);
//                                                        the 128 keys are those of ALL images of the tile, the softmax below only
//                                                        looks at the row's own image (block-diagonal mask);
//   softmax thread = query row = TMEM lane; two reads of the row's S score columns (max, then exp / sum); the unnormalised
//           probabilities go to shared memory as bf16, K-major [128 rows x 128 keys] (zeros outside the image's block);
//   output  D2[128 x 64] = P[128 x 128] V[128 x 64]       M = 128, N = 64, K = 16 x 8; V is MN-major straight from its TMA box
//                                                        (rows = keys, 64 channels = both 32-wide heads of the pair: the other
//                                                        head's 32 columns are computed and dropped);
//   epilogue o = D2 / l (fp32) -> bf16, lse = max + log l (what dmu_attn_bwd reads).
// The scores never leave TMEM / shared memory ("flash-style"); for S <= 128 there is no key loop and no online rescaling.
// P is rounded to bf16 (2^-9 relative) before the second contraction, like every activation of the bf16 path.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace dmu {
namespace tc {

struct AttnMaps { CUtensorMap qkv; };

struct AttnTcArgs {
    __nv_bfloat16* o; int64_t o_pitch;
    float* lse;
    int rows, S, C, heads, D;
    float scale;
};

constexpr int kAtTile = 128 * 128;        // one 64-channel x 128-row box

template <int D>
__global__ void __launch_bounds__(128) attn_fwd_tc_kernel(const __grid_constant__ AttnMaps maps, const AttnTcArgs P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar_load, bar_s, bar_o;
    __shared__ uint32_t s_tmem;
    uint8_t* s_q = smem;
    uint8_t* s_k = smem + kAtTile;
    uint8_t* s_v = smem + 2 * kAtTile;
    uint8_t* s_p = smem + 3 * kAtTile;            // two 64-key chunks of 128 rows x 128 B
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * 128;
    const int h = blockIdx.y;
    const int chunk = (h * D) >> 6;               // 64-channel chunk of q / k / v that holds this head
    const int sub = (h * D) & 63;                 // channel offset of the head inside the chunk (0 or 32)

    pdl_trigger();
    if (threadIdx.x == 0) {
        mbar_init(&bar_load, 1); mbar_init(&bar_s, 1); mbar_init(&bar_o, 1);
        fence_mbar_init();
        tma_prefetch_desc(&maps.qkv);
    }
    if (warp == 1) tmem_alloc(&s_tmem, 256);
    for (int i = threadIdx.x; i < 2 * kAtTile / 16; i += blockDim.x) reinterpret_cast<uint4*>(s_p)[i] = make_uint4(0u, 0u, 0u, 0u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&bar_load, 3 * kAtTile);
            tma_load_2d(s_q, &maps.qkv, &bar_load, chunk * 64, row0);
            tma_load_2d(s_k, &maps.qkv, &bar_load, P.C + chunk * 64, row0);
            tma_load_2d(s_v, &maps.qkv, &bar_load, 2 * P.C + chunk * 64, row0);
            mbar_wait(&bar_load, 0);
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
            const uint64_t dq = smem_desc_sw128(smem_u32(s_q), 16, 1024) + (uint64_t)(sub >> 3);      // + sub * 2 bytes, 16-byte units
            const uint64_t dk = smem_desc_sw128(smem_u32(s_k), 16, 1024) + (uint64_t)(sub >> 3);
#pragma unroll
            for (int k = 0; k < D / 16; ++k) umma_bf16(tmem, dq + 2 * k, dk + 2 * k, idesc, k != 0);
            umma_commit(&bar_s);
        }
        __syncwarp();
    }

    // ---- softmax over the row's own image: keys [g * S, g * S + S) of the tile
    mbar_wait(&bar_s, 0);
    tc_fence_after();
    const int r = threadIdx.x;
    const int S = P.S;
    const int k_lo = (r / S) * S, k_hi = k_lo + S;
    const int c_lo = k_lo & ~31;                           // first 32-column chunk that touches the block
    const int nchunks = S > 32 ? S / 32 : 1;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float mx = -INFINITY;
    for (int ci = 0; ci < nchunks; ++ci) {
        float v[32];
        tmem_ld32(tmem + lane_base + (uint32_t)(c_lo + ci * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int col = c_lo + ci * 32 + i;
            if (col >= k_lo && col < k_hi) mx = fmaxf(mx, v[i] * P.scale);
        }
    }
    float l = 0.f;
    for (int ci = 0; ci < nchunks; ++ci) {
        float v[32];
        tmem_ld32(tmem + lane_base + (uint32_t)(c_lo + ci * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int col = c_lo + ci * 32 + i;
            const float p = (col >= k_lo && col < k_hi) ? __expf(v[i] * P.scale - mx) : 0.f;
            v[i] = __bfloat162float(__float2bfloat16_rn(p));      // the sum uses what the second contraction sees
            l += v[i];
        }
        // 32 keys = four 16-byte granules of this row, K-major SWIZZLE_128B inside the 64-key chunk
        const int kc = (c_lo + ci * 32) >> 6, k0 = (c_lo + ci * 32) & 63;
        uint8_t* rowp = s_p + kc * kAtTile + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 pk;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
            const int gran = (k0 >> 3) + q;
            *reinterpret_cast<uint4*>(rowp + ((gran ^ (r & 7)) << 4)) = pk;
        }
    }
    fence_proxy_async();       // generic-proxy writes of P -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();

    if (warp == 0) {
        if (elect_one()) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);       // P K-major, V MN-major
            const uint64_t dp = smem_desc_sw128(smem_u32(s_p), 16, 1024);
            const uint64_t dv = smem_desc_sw128(smem_u32(s_v), 8192, 1024);
#pragma unroll
            for (int k = 0; k < 8; ++k)      // 16 keys per step: 32 B along P's rows (next 64-key chunk after four), 16 rows of V
                umma_bf16(tmem + 128, dp + (uint64_t)((k >> 2) * (kAtTile >> 4) + (k & 3) * 2), dv + (uint64_t)(k * 128), idesc, k != 0);
            umma_commit(&bar_o);
        }
        __syncwarp();
    }
    mbar_wait(&bar_o, 0);
    tc_fence_after();
    const int grow = row0 + r;
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < D; c += 32) {
        float v[32];
        tmem_ld32(tmem + lane_base + (uint32_t)(128 + sub + c), v);
        tmem_ld_wait();
        if (grow < P.rows) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= inv;
            __nv_bfloat16* op = P.o + (int64_t)grow * P.o_pitch + h * D + c;
#pragma unroll
            for (int i = 0; i < 32; i += 8) store_vec<__nv_bfloat16>(op + i, v + i);
        }
    }
    if (P.lse && grow < P.rows) P.lse[((int64_t)(grow / S) * P.heads + h) * S + grow % S] = mx + __logf(l);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}

static bool attn_tc_enabled() {
    static const int v = [] { const char* e = getenv("DMU_ATTN_TC"); return e ? atoi(e) : 1; }();
    return v != 0;
}

int attn_tc_supported(const dmu_attn_params* p) {
    if (!attn_tc_enabled() || !p || p->dtype != DMU_BF16 || p->heads < 1) return 0;
    const int D = p->C / p->heads;
    if ((D != 32 && D != 64) || p->C % 64 != 0) return 0;
    // measured inside the training step (B = 128): at S = 16 and above the tensor-core kernel wins (S = 64, B = 256: 12.3 vs 33.8 us),
    // at S = 4 and 1 the SIMT kernel's 3.7 - 4.4 us are below this kernel's TMA -> MMA -> softmax -> MMA latency chain.
    // DMU_ATTN_TC_MIN_S=1 sends every supported shape here (tests).
    static const int min_s = [] { const char* e = getenv("DMU_ATTN_TC_MIN_S"); return e ? atoi(e) : 16; }();
    if (p->S < min_s || p->S > 128 || 128 % p->S != 0) return 0;
    if ((reinterpret_cast<uintptr_t>(p->qkv) & 15) || p->qkv_pitch % 8 || p->qkv_pitch < 3 * p->C) return 0;
    if ((reinterpret_cast<uintptr_t>(p->o) & 15) || p->o_pitch % 8) return 0;
    if ((int64_t)p->N * p->S >= (1ll << 31) - 256 || encode_tiled_fn() == nullptr) return 0;
    return 1;
}

int attn_tc_launch(const dmu_attn_params* p, cudaStream_t stream) {
    AttnMaps maps;
    AttnTcArgs A;
    memset(&A, 0, sizeof(A));
    A.o = reinterpret_cast<__nv_bfloat16*>(p->o); A.o_pitch = p->o_pitch;
    A.lse = p->lse;
    A.rows = p->N * p->S; A.S = p->S; A.C = p->C; A.heads = p->heads; A.D = p->C / p->heads;
    A.scale = 1.f / sqrtf((float)A.D);
    {
        const uint64_t dims[2] = {(uint64_t)3 * p->C, (uint64_t)A.rows};
        const uint64_t str[2] = {1, (uint64_t)p->qkv_pitch};
        const uint32_t box[2] = {64, 128};
        if (int rc = make_map_bf16(&maps.qkv, p->qkv, 2, dims, str, box, "dmu_attn_fwd/tc qkv")) return rc;
    }
    const dim3 grid((unsigned)((A.rows + 127) / 128), (unsigned)p->heads);
    const size_t smem = 5 * kAtTile + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(attn_fwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(attn_fwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    cudaError_t e = A.D == 32 ? launch_pdl(attn_fwd_tc_kernel<32>, grid, dim3(128), smem, stream, dim3(1, 1, 1), maps, A)
                              : launch_pdl(attn_fwd_tc_kernel<64>, grid, dim3(128), smem, stream, dim3(1, 1, 1), maps, A);
    if (e != cudaSuccess) return fail("dmu_attn_fwd/tc: launch failed: %s", cudaGetErrorString(e));
    return check_launch("dmu_attn_fwd/tc");
}

}  // namespace tc
}  // namespace dmu
