// Weight gradient of a 3x3 stride-1 convolution as a tcgen05 kernel that reads both activation tensors ONCE per tap group.
//
// Why: wgrad_tc_kernel (conv_tc.cu) gives a CTA two (tap, 64-channel) units and streams, per 64 pixels, two shifted boxes of q
// and one box of p: every tap pulls q through the SM's L2 port again (250 MB for the 33.6 MB of a 64->64 layer at 128x32x32),
// and the kernel is bound by exactly that (tensor pipe 13 %).  Here the contraction index is the zero-padded flat position
// space [N][H+2][W+2] of conv_halo.cu, in which every tap is a pure shift:
//   dw[a][r][s][b] = sum_Q  p_pad[Q][a] * q_pad[Q + (r-1)*PW + (s-1)][b]
// One shared-memory halo tile of q (NRq padded rows, single-row TMA boxes whose out-of-bounds pixels are the padding) and one
// tile of p (the 128 positions of the k-tile; its padding positions are TMA zero fill, so they contribute nothing) serve all
// taps of the CTA.  Both operands are MN-major (a shared-memory row = one position = 64 channels = 128 B, K = rows), and
//   * the A descriptor of tap (r,s) is the q tile's descriptor advanced by (r*PW+s) rows, and
//   * TWO taps share one M = 128 instruction: the second 64-row block of an MN-major operand sits `leading-dimension byte
//     offset` after the first, and that offset may be any multiple of 128 B - here the distance between the two taps' windows,
//     so the blocks overlap in shared memory.  (Both verified on hardware: scripts/probes/umma_mn_rowoffset.cu.)
// Taps 0..4 (two pairs + one M = 64 single) and 5..8 (two pairs) form two groups of CTAs, 3 x 64 resp. 2 x 64 TMEM columns, so a
// CTA of the other backward lane still finds TMEM on the SM; the position tiles are split over the CTAs of a group in proportion
// 4 : 3.  Accumulators go to dw by fp32 reductions, a warp = 32 consecutive b (contiguous in the staging layout).
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace dmu {
namespace tc {

struct WHaloMaps { CUtensorMap q, p; };

constexpr int kWhSlots = 3;
struct WHaloSlot { int row0, lbo, tap_a, tap_b; };      // window of tap_a starts row0 rows into the halo; lbo = byte distance to tap_b's (0: single)

struct WHaloArgs {
    int N, H, W, Ca, Cb;
    int PW, PH, NRq, NRp, tiles;
    int q_bytes, stage_bytes, stages;
    int splits[2], tiles_per_split[2], nslots[2];
    WHaloSlot slots[2][kWhSlots];
    float* dw; int64_t dw_sa, dw_sb, dw_st;
};

constexpr int kWhMaxStages = 4;

__device__ __forceinline__ int wh_floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

__global__ void __launch_bounds__(128) wgrad_halo_kernel(const __grid_constant__ WHaloMaps maps, const __grid_constant__ WHaloArgs P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[kWhMaxStages], empty_bar[kWhMaxStages], acc_bar;
    __shared__ uint32_t s_tmem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = (int)blockIdx.z < P.splits[0] ? 0 : 1;
    const int split = (int)blockIdx.z - (grp ? P.splits[0] : 0);
    const int cb0 = blockIdx.x * 64, a0 = blockIdx.y * 64;
    const int tile_lo = split * P.tiles_per_split[grp];
    const int tile_hi = min(P.tiles, tile_lo + P.tiles_per_split[grp]);
    const int nslots = P.nslots[grp];
    constexpr uint32_t kCols = 256;      // 3 x 64 accumulator columns, allocated as a power of two

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&acc_bar, 1);
        fence_mbar_init();
        tma_prefetch_desc(&maps.q);
        tma_prefetch_desc(&maps.p);
    }
    if (warp == 1) tmem_alloc(&s_tmem, kCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    pdl_wait();

    const uint32_t row_bytes = (uint32_t)P.PW * 128u;
    if (warp == 0) {
        // ------------------------------------------------ TMA producer: NRq padded rows of q, NRp padded rows of p per tile
        if (elect_one()) {
            int st = 0, par = 1;
            const uint32_t tx = (uint32_t)(P.NRq + P.NRp) * row_bytes;
            for (int tile = tile_lo; tile < tile_hi; ++tile) {
                const int Q0 = tile * 128;
                const int L0 = wh_floordiv(Q0 - P.PW - 1, P.PW), Lp = Q0 / P.PW;
                mbar_wait(&empty_bar[st], par);
                uint8_t* dq = smem + (size_t)st * P.stage_bytes;
                uint8_t* dp = dq + P.q_bytes;
                mbar_arrive_expect_tx(&full_bar[st], tx);
                int n = wh_floordiv(L0, P.PH), hp = L0 - n * P.PH;
                for (int i = 0; i < P.NRq; ++i) {
                    tma_load_4d(dq + (size_t)i * row_bytes, &maps.q, &full_bar[st], cb0, -1, hp - 1, n);
                    if (++hp == P.PH) { hp = 0; ++n; }
                }
                n = Lp / P.PH; hp = Lp - n * P.PH;
                for (int i = 0; i < P.NRp; ++i) {
                    tma_load_4d(dp + (size_t)i * row_bytes, &maps.p, &full_bar[st], a0, -1, hp - 1, n);
                    if (++hp == P.PH) { hp = 0; ++n; }
                }
                if (++st == P.stages) { st = 0; par ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer: per tile, 8 k-steps of 16 positions x the group's slots
        if (elect_one()) {
            constexpr uint32_t idesc_pair = umma_idesc_bf16(128, 64, 1, 1), idesc_single = umma_idesc_bf16(64, 64, 1, 1);
            const uint32_t smem0 = smem_u32(smem);
            int st = 0, par = 0, it = 0;
            for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
                const int Q0 = tile * 128;
                const int L0 = wh_floordiv(Q0 - P.PW - 1, P.PW), Lp = Q0 / P.PW;
                const int base_off = Q0 - P.PW - 1 - L0 * P.PW;      // halo row of padded position Q0 - PW - 1 (tap (0,0) of position Q0)
                const int p_off = Q0 - Lp * P.PW;
                mbar_wait(&full_bar[st], par);
                tc_fence_after();
                const uint32_t q_base = smem0 + (uint32_t)(st * P.stage_bytes) + (uint32_t)base_off * 128u;
                const uint64_t db = smem_desc_sw128(q_base - (uint32_t)base_off * 128u + (uint32_t)P.q_bytes + (uint32_t)p_off * 128u, 16, 1024);
#pragma unroll 1
                for (int s = 0; s < nslots; ++s) {
                    const WHaloSlot sl = P.slots[grp][s];
                    const uint64_t da = smem_desc_sw128(q_base + (uint32_t)sl.row0 * 128u, sl.lbo ? (uint32_t)sl.lbo : 16u, 1024);
                    const uint32_t idesc = sl.lbo ? idesc_pair : idesc_single;
                    const uint32_t d_tmem = tmem + (uint32_t)(s * 64);
#pragma unroll
                    for (int k = 0; k < 8; ++k)      // 16 positions = 16 rows of 128 B per step
                        umma_bf16(d_tmem, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (it | k) != 0);
                }
                umma_commit(&empty_bar[st]);
                if (++st == P.stages) { st = 0; par ^= 1; }
            }
            umma_commit(&acc_bar);
        }
        __syncwarp();
    }
    __syncwarp();

    // ---------------------------------------------------- epilogue: thread = accumulator row = (tap, channel b); columns = channels a
    if (tile_lo < tile_hi) {
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        for (int s = 0; s < nslots; ++s) {
            const WHaloSlot sl = P.slots[grp][s];
            // M = 128: row m in TMEM lane m (rows 64.. = tap_b); M = 64: row m in lane (m / 16) * 32 + m % 16
            const bool ok = sl.lbo ? true : lane < 16;
            const int tap = sl.lbo ? (threadIdx.x < 64 ? sl.tap_a : sl.tap_b) : sl.tap_a;
            const int b = cb0 + (sl.lbo ? (int)(threadIdx.x & 63) : warp * 16 + (lane & 15));
            float* dwp = P.dw + (int64_t)b * P.dw_sb + (int64_t)tap * P.dw_st + (int64_t)a0 * P.dw_sa;
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * 64 + c), v);
                tmem_ld_wait();
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) atomicAdd(dwp + (int64_t)(c + i) * P.dw_sa, v[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, kCols);
}

static inline bool wh_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool wh_nhwc_bf16_ok(const dmu_tensor4& t) {
    return t.dtype == DMU_BF16 && t.sc == 1 && wh_aligned16(t.ptr) && t.sw % 8 == 0 && t.sh % 8 == 0 && t.sn % 8 == 0;
}

static int wh_geometry(const dmu_wgrad_params* p, WHaloArgs& A) {
    memset(&A, 0, sizeof(A));
    A.N = p->N; A.H = p->Hq; A.W = p->Wq; A.Ca = p->Ca; A.Cb = p->Cb;
    A.PW = A.W + 2; A.PH = A.H + 2;
    A.NRq = 3 + (129 + A.PW - 1) / A.PW;
    A.NRp = 1 + (127 + A.PW - 1) / A.PW;
    A.tiles = (int)(((int64_t)A.N * A.PH * A.PW + 127) / 128);
    A.q_bytes = (A.NRq * A.PW * 128 + 1023) / 1024 * 1024;
    const int p_bytes = (A.NRp * A.PW * 128 + 1023) / 1024 * 1024;
    A.stage_bytes = A.q_bytes + p_bytes;
    A.stages = 3;      // measured: 20.0 us with three stages, 25 us with two (64->64 at 128x32x32); inside the step no difference
    while (A.stages > 2 && A.stages * A.stage_bytes + 1024 > 200 * 1024) --A.stages;
    if (A.stages * A.stage_bytes + 1024 > 220 * 1024) return -1;
    return A.stages * A.stage_bytes + 1024;
}

// force: any supported shape (tests, impl 5); otherwise only where it measured faster than the per-tap kernel
int wgrad_halo_supported(const dmu_wgrad_params* p, int force) {
    if (!p || !p->p.ptr || !p->q.ptr || !p->dw) return 0;
    if (p->R != 3 || p->S != 3 || p->stride != 1 || p->pad != 1) return 0;
    if (p->Hp != p->Hq || p->Wp != p->Wq) return 0;
    if (p->Hq < 8 || p->Wq < 8 || p->Wq + 2 > 256) return 0;
    if (p->Ca % 64 != 0 || p->Cb % 64 != 0) return 0;
    if (!wh_nhwc_bf16_ok(p->p) || !wh_nhwc_bf16_ok(p->q) || encode_tiled_fn() == nullptr) return 0;
    if ((int64_t)p->N * (p->Hq + 2) * (p->Wq + 2) >= (1ll << 31) - 4096) return 0;
    WHaloArgs A;
    if (wh_geometry(p, A) <= 0) return 0;
    if (force) return 1;
    static const int enabled = [] { const char* e = getenv("DMU_WGRAD_HALO"); return e ? atoi(e) : 1; }();
    if (!enabled) return 0;
    // below a few tiles per CTA the reductions of the (splits x 9 x Ca x Cb) partial sums outweigh what the operand traffic saves
    return A.tiles >= 4 * sm_count() ? 1 : 0;      // measured: 324 tiles (16x16 at B = 128) 11.7 vs 13.1 us, 100 tiles slower
}

int wgrad_halo_launch(const dmu_wgrad_params* p, cudaStream_t stream) {
    WHaloMaps maps;
    WHaloArgs A;
    const int smem = wh_geometry(p, A);
    DMU_REQUIRE(smem > 0, "dmu_conv2d_wgrad/halo: tile does not fit shared memory");
    {
        const uint64_t dims[4] = {(uint64_t)p->Cb, (uint64_t)p->Wq, (uint64_t)p->Hq, (uint64_t)p->N};
        const uint64_t str[4] = {1, (uint64_t)p->q.sw, (uint64_t)p->q.sh, (uint64_t)p->q.sn};
        const uint32_t box[4] = {64, (uint32_t)A.PW, 1, 1};
        if (int rc = make_map_bf16(&maps.q, p->q.ptr, 4, dims, str, box, "dmu_conv2d_wgrad/halo q")) return rc;
    }
    {
        const uint64_t dims[4] = {(uint64_t)p->Ca, (uint64_t)p->Wp, (uint64_t)p->Hp, (uint64_t)p->N};
        const uint64_t str[4] = {1, (uint64_t)p->p.sw, (uint64_t)p->p.sh, (uint64_t)p->p.sn};
        const uint32_t box[4] = {64, (uint32_t)A.PW, 1, 1};
        if (int rc = make_map_bf16(&maps.p, p->p.ptr, 4, dims, str, box, "dmu_conv2d_wgrad/halo p")) return rc;
    }
    A.dw = p->dw; A.dw_sa = p->dw_sa; A.dw_sb = p->dw_sb; A.dw_st = p->dw_st;
    // tap t = r * 3 + s reads the halo from row r * PW + s (relative to the window of tap 0)
    auto row_of = [&](int t) { return (t / 3) * A.PW + t % 3; };
    auto pair = [&](int ta, int tb) { return WHaloSlot{row_of(ta), (row_of(tb) - row_of(ta)) * 128, ta, tb}; };
    A.nslots[0] = 3; A.slots[0][0] = pair(0, 1); A.slots[0][1] = pair(2, 3); A.slots[0][2] = WHaloSlot{row_of(4), 0, 4, 4};
    A.nslots[1] = 2; A.slots[1][0] = pair(5, 6); A.slots[1][1] = pair(7, 8);
    // CTAs: about one wave; a group-0 CTA issues 4 units of MMA time per tile, a group-1 CTA 3
    const int per_xy = (p->Cb / 64) * (p->Ca / 64);
    int total = (sm_count() + per_xy - 1) / per_xy;
    if (total < 2) total = 2;
    int s0 = (total * 4 + 3) / 7, s1 = total - s0;
    if (s1 < 1) s1 = 1;
    for (int g = 0; g < 2; ++g) {
        int s = g ? s1 : s0;
        if (s > A.tiles) s = A.tiles;
        if (s < 1) s = 1;
        A.tiles_per_split[g] = (A.tiles + s - 1) / s;
        A.splits[g] = (A.tiles + A.tiles_per_split[g] - 1) / A.tiles_per_split[g];
    }
    dim3 grid(p->Cb / 64, p->Ca / 64, A.splits[0] + A.splits[1]);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        attr_done = true;
    }
    cudaError_t e = launch_pdl(wgrad_halo_kernel, grid, dim3(128), (size_t)smem, stream, dim3(1, 1, 1), maps, A);
    if (e != cudaSuccess) return fail("dmu_conv2d_wgrad/halo: launch failed: %s", cudaGetErrorString(e));
    if (int rc = check_launch("dmu_conv2d_wgrad/halo")) return rc;
    if (p->dbias) return dmu_colsum(&p->p, p->N, p->Hp, p->Wp, p->Ca, nullptr, 0, p->dbias, 1.0f, (dmu_stream_t)stream);
    return 0;
}

}  // namespace tc
}  // namespace dmu
