// 3x3 stride-1 convolution (fprop and dgrad) as a persistent tcgen05 kernel that reads every input pixel ONCE per CTA.
//
// Why: conv_tc.cu issues one TMA box per (tap, 64-channel chunk), i.e. it pulls the activation tile nine times through the
// SM's L2 port, plus the weights once per tile.  Measured on B200 (scripts/conv_timeline.py): an SM ingests ~60 B/clk from
// L2 whatever the grid, so that kernel is bound by bytes-per-SM (216 KB per 128-pixel tile of a 64->64 layer), not by the
// tensor pipe (23 % at best).  Here:
//   * the output tile is 128 consecutive positions of the zero-padded flat space [N][H+2][W+2]; in that space every tap is
//     a pure shift, so ONE shared-memory halo tile (NR padded rows x (W+2) pixels x 64 channels, loaded by NR single-row TMA
//     boxes whose out-of-bounds pixels are the zero padding) serves all nine taps: tap (r,s) is the same tile read through
//     a UMMA descriptor whose start address is advanced by (r*(W+2)+s) rows.  Row-granular start offsets inside the
//     128-byte swizzle pattern are legal: both TMA writes and UMMA reads swizzle on absolute shared-memory address bits
//     (scripts/probes/umma_rowoffset.cu, tma_unaligned_dst.cu, verified on hardware).
//   * the kernel is persistent (one CTA per SM, static round-robin over tiles); when the whole filter bank of the CTA's
//     output-channel tile fits (9 x Ck x NT bf16 <= ~144 KB) it is loaded once and stays resident, otherwise it streams
//     through its own mbarrier ring.
//   * warp roles: 8 epilogue warps (two per TMEM lane quarter, half of the tile's channels each), 1 TMA warp, 1 MMA warp; two
//     TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
// Positions of the padded space that are padding themselves are computed and dropped (efficiency H*W/((H+2)(W+2)): 89 %
// at 32x32, 79 % at 16x16, 64 % at 8x8); below 8x8 the per-tap kernel (with split-K) is used instead.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace dmu {
namespace tc {

extern long long* g_debug_buffer;   // conv_tc.cu (dmu_debug_set_buffer)

struct HaloMaps { CUtensorMap a, b; };

struct HaloArgs {
    int N, H, W, Ck, Cj;
    int PW, PH, NR, tiles, chunks, flip, resident;
    int a_stage_bytes, a_stages, w_stages;
    __nv_bfloat16* y; int64_t y_sn, y_sh, y_sw;
    const __nv_bfloat16* res; int64_t r_sn, r_sh, r_sw;
    const float* bias;
    const float* temb; int64_t temb_pitch;
    // narrow output (the C->3 head conv): Cj <= 4 channels written as fp32 through arbitrary strides (NCHW at the API
    // boundary); the weight box still has NT rows, the rows past Cj are TMA out-of-bounds zeros
    float* yn; int64_t n_sn, n_sh, n_sw, n_sc; int narrow;
    long long* dbg;     // development aid: CTA 0 stamps clock64 per tile (8 slots per tile)
    // fused GroupNorm(+SiLU) on the gathered operand: a = act(x * coef[n][k][0] + coef[n][k][1]) is applied to the halo tile
    // in shared memory by four dedicated warps before the MMAs read it; a (the conv's real input) is optionally written out
    // for the backward (weight gradient operand)
    const float* gn_coef; int gn_silu;
    __nv_bfloat16* a_out; int64_t a_sn, a_sh, a_sw;
    // GroupNorm statistics of the OUTPUT (dmu_conv_params.gn_fuse_mode 3): raw (sum, sum of squares) of the stored values, added to
    // st_sums[n][g][0..1]; st_sh = log2(channels per group)
    float* st_sums; int st_G, st_sh;
    unsigned long long* st_fixed;      // DMU_GN_FIXED_SUMS: the int64 fixed-point accumulators behind st_sums (order-independent)
    int run, run_sh;     // tile schedule: runs of `run` = 1 << run_sh consecutive tiles per CTA, dealt round-robin
    uint32_t pw_magic;   // floor(2^32 / PW) + 1 when every padded position * PW stays below 2^32 (floordiv_magic), else 0
};

constexpr int kMaxAStages = 4, kMaxWStages = 8;

__device__ __forceinline__ int floordiv_dev(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
// floor(a / d) for a >= -2 d by one multiply: magic = floor(2^32 / d) + 1, exact while (a + 2 d) * d < 2^32 (checked by the host;
// the per-tile integer divisions sat on the MMA warp's path between two tiles: ~150 clk each of a ~2.7k clk tile)
__device__ __forceinline__ int floordiv_magic(int a, int d, uint32_t magic) {
    return (int)(((uint64_t)(uint32_t)(a + 2 * d) * magic) >> 32) - 2;
}

// x * sigmoid(x) with one MUFU: sigmoid(x) = 0.5 * (1 + tanh(x / 2))
__device__ __forceinline__ float silu_tanh(float t) {
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * t));
    return 0.5f * t * (1.f + th);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Epilogue warps: 8 in the plain kernel - warp w and warp w + 4 share the TMEM lane quarter w % 4 and each takes half of the tile's
// channels.  Measured (ncu, 64->64 at 256x64x64): with 4 warps the epilogue of a tile (~1070 instructions per warp, one warp per
// scheduler) took ~4.3k clk against ~2.0k clk of MMA issue, i.e. the kernel was bound by its epilogue, tensor pipe 30 %.
// What the ~2650 clk of a 128-position tile are (scripts/halo_timeline.py, 64 -> 64): ~2050 clk in which the 36 MMAs issue at the pace
// of the SM's shared-memory bandwidth (36 x 6 KB of operand reads plus the 30-42 KB TMA write of a later tile at 128 B/clk), and
// ~450-600 clk between the last MMA of a tile and the first of the next (two commits, the accumulator / halo barrier waits with their
// tcgen05 fences, tile arithmetic).  Measured and dropped: a second MMA warp taking every other tile (its own accumulator and its own
// half of the halo ring) - the second issuer still starts ~270 clk after the first one's last MMA, whether it waits with try_wait or
// polls with test_wait, and DDIM-50 stayed at 1.80k img/s; the whole issue loop as one elected thread with the next tile's two barrier
// waits polled after the sixth tap (a tile 2.69k -> 2.57k clk in the timeline, the time moves into the two commits; 1.80k img/s).
template <bool GN> struct HaloCfg { static constexpr int kEpiWarps = GN ? 4 : 8; static constexpr int kThreads = (kEpiWarps + 2 + (GN ? 8 : 0)) * 32; };

template <int NT, bool GN>
__global__ void __launch_bounds__(HaloCfg<GN>::kThreads, 1) conv3x3_halo_kernel(const __grid_constant__ HaloMaps maps, const __grid_constant__ HaloArgs P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t a_full[kMaxAStages], a_empty[kMaxAStages], w_full[kMaxWStages], w_empty[kMaxWStages];
    __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
    __shared__ __align__(8) uint64_t a_ready[kMaxAStages];      // GN: transform warps -> MMA warp
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[NT];

    constexpr int kWTile = NT * 128;                       // one (tap, chunk) weight block: NT rows x 64 bf16
    constexpr int kEpiWarps = HaloCfg<GN>::kEpiWarps, kTmaWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1, kXformWarp0 = kEpiWarps + 2;
    uint8_t* smem_a = smem;
    uint8_t* smem_w = smem + (size_t)P.a_stages * P.a_stage_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j0 = blockIdx.y * NT;
    const int kblocks = 9 * P.chunks;
    // tile schedule of this CTA: runs of P.run consecutive tiles, the runs dealt round-robin over the grid (run = 1: plain
    // round-robin).  With GroupNorm statistics in the epilogue a thread's consecutive tiles should stay inside one image (the sums
    // are flushed when the image changes), while the CTAs should still walk the tensor side by side (one contiguous range per CTA
    // measured 15 % slower at 64 x 64: 148 distant DRAM streams).
    const int run = P.run, run_sh = P.run_sh;
    int t_count = 0;
    for (int base = (int)blockIdx.x * run; base < P.tiles; base += (int)gridDim.x * run) t_count += min(run, P.tiles - base);
    auto tile_at = [&](int i) { return (((i >> run_sh) * (int)gridDim.x + (int)blockIdx.x) << run_sh) + (i & (run - 1)); };

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kMaxWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps * 32); }
        for (int i = 0; i < kMaxAStages; ++i) mbar_init(&a_ready[i], 8);
        fence_mbar_init();
        tma_prefetch_desc(&maps.a);
        tma_prefetch_desc(&maps.b);
    }
    if (warp == kMmaWarp) tmem_alloc(&s_tmem, 2 * NT);
    for (int i = threadIdx.x; i < NT; i += blockDim.x)       // parameters: not produced by the previous launch
        s_bias[i] = (P.bias && j0 + i < P.Cj) ? P.bias[j0 + i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    pdl_wait();

    // TMA and MMA roles run as whole converged warps; the asynchronous instructions sit under elect_one() (tc_common.cuh)
    if (warp == kTmaWarp) {
        // ------------------------------------------------ TMA producer
        if (P.resident && elect_one()) {
            mbar_arrive_expect_tx(&w_full[0], (uint32_t)(kblocks * kWTile));
            for (int c = 0; c < P.chunks; ++c)
                for (int t = 0; t < 9; ++t)
                    tma_load_2d(smem_w + (size_t)(c * 9 + t) * kWTile, &maps.b, &w_full[0], t * P.Ck + c * 64, j0);
        }
        __syncwarp();
        int sa = 0, pa = 1, sw = 0, pw = 1;
        const uint32_t a_bytes = (uint32_t)(P.NR * P.PW) * 128u;
        for (int ti = 0; ti < t_count; ++ti) {
            const int tile = tile_at(ti);
            const int L0 = P.pw_magic ? floordiv_magic(tile * 128 - P.PW - 1, P.PW, P.pw_magic) : floordiv_dev(tile * 128 - P.PW - 1, P.PW);
            for (int c = 0; c < P.chunks; ++c) {
                long long* pdbg = (!GN && P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && ti < 64 && c == 0 && lane == 0) ? P.dbg + 8 * 64 + 8 * ti : nullptr;
                if (pdbg) pdbg[1] = clock64();
                mbar_wait(&a_empty[sa], pa);
                if (pdbg) pdbg[2] = clock64();
                if (elect_one()) {
                    mbar_arrive_expect_tx(&a_full[sa], a_bytes);
                    uint8_t* dst = smem_a + (size_t)sa * P.a_stage_bytes;
                    int n = floordiv_dev(L0, P.PH), hp = L0 - n * P.PH;
                    for (int i = 0; i < P.NR; ++i) {
                        tma_load_4d(dst + (size_t)i * P.PW * 128, &maps.a, &a_full[sa], c * 64, -1, hp - 1, n);
                        if (++hp == P.PH) { hp = 0; ++n; }
                    }
                }
                __syncwarp();
                if (pdbg) pdbg[3] = clock64();
                if (++sa == P.a_stages) { sa = 0; pa ^= 1; }
                if (!P.resident) {
                    for (int t = 0; t < 9; ++t) {
                        mbar_wait(&w_empty[sw], pw);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&w_full[sw], (uint32_t)kWTile);
                            tma_load_2d(smem_w + (size_t)sw * kWTile, &maps.b, &w_full[sw], t * P.Ck + c * 64, j0);
                        }
                        __syncwarp();
                        if (++sw == P.w_stages) { sw = 0; pw ^= 1; }
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_bf16(128, NT, 0, 0);
        if (P.resident) mbar_wait(&w_full[0], 0);
        int sa = 0, pa = 0, sw = 0, pw = 0, it = 0;
        for (; it < t_count; ++it) {
            const int tile = tile_at(it);
            const int buf = it & 1;
            long long* dbg = (P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && it < 64 && lane == 0) ? P.dbg + 8 * it : nullptr;
            if (dbg) dbg[0] = clock64();
            mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
            tc_fence_after();
            if (dbg) dbg[1] = clock64();
            const uint32_t d_tmem = tmem + (uint32_t)(buf * NT);
            const int Q0 = tile * 128;
            const int L0 = P.pw_magic ? floordiv_magic(Q0 - P.PW - 1, P.PW, P.pw_magic) : floordiv_dev(Q0 - P.PW - 1, P.PW);
            const int base_off = Q0 - P.PW - 1 - L0 * P.PW;       // smem row of padded position (Q0 - PW - 1)
            for (int c = 0; c < P.chunks; ++c) {
                if (!GN && dbg && c == 0) dbg[8 * 64] = clock64();      // (the slots of the transform warps are free in the plain kernel)
                mbar_wait(GN ? &a_ready[sa] : &a_full[sa], pa);
                if (GN) tc_fence_after();      // (a TMA fill needs no tcgen05 fence: the mbarrier's complete_tx orders it before the MMAs)
                if (dbg && c == 0) dbg[2] = clock64();
                const uint32_t a_base = smem_u32(smem_a + (size_t)sa * P.a_stage_bytes) + (uint32_t)base_off * 128u;
                if (P.resident) {
                    const uint32_t w_base = smem_u32(smem_w + (size_t)(c * 9) * kWTile);
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            const int r = t / 3, s = t % 3;
                            const int dr = P.flip ? 2 - r : r, ds = P.flip ? 2 - s : s;
                            const uint64_t da = smem_desc_sw128(a_base + (uint32_t)(dr * P.PW + ds) * 128u, 16, 1024);
                            const uint64_t db = smem_desc_sw128(w_base + (uint32_t)(t * kWTile), 16, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (c | t | k) != 0);
                        }
                        umma_commit(&a_empty[sa]);
                    }
                    __syncwarp();
                } else {
                    for (int t = 0; t < 9; ++t) {
                        const int r = t / 3, s = t % 3;
                        const int dr = P.flip ? 2 - r : r, ds = P.flip ? 2 - s : s;
                        mbar_wait(&w_full[sw], pw);
                        tc_fence_after();
                        const uint64_t da = smem_desc_sw128(a_base + (uint32_t)(dr * P.PW + ds) * 128u, 16, 1024);
                        const uint64_t db = smem_desc_sw128(smem_u32(smem_w + (size_t)sw * kWTile), 16, 1024);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (c | t | k) != 0);
                            umma_commit(&w_empty[sw]);
                            if (t == 8) umma_commit(&a_empty[sa]);
                        }
                        __syncwarp();
                        if (++sw == P.w_stages) { sw = 0; pw ^= 1; }
                    }
                }
                if (++sa == P.a_stages) { sa = 0; pa ^= 1; }
            }
            if (elect_one()) umma_commit(&acc_full[buf]);
            __syncwarp();
            if (dbg) dbg[3] = clock64();
        }
    } else if (GN && warp >= kXformWarp0) {
        // ------------------------------------------------ GroupNorm(+SiLU) transform warps (the eight after the MMA warp)
        // Tile row r is padded-flat position L0 * PW + r.  Rows that are padding (TMA zero fill) stay zero; every other row is
        // normalised in place: thread = one 16-byte chunk (8 channels) of one pixel, the chunk's logical position follows the
        // 128-byte swizzle (physical chunk ^ (row & 7); the stage base is 1024-byte aligned).
        // Warp = every 8th row of the tile (uniform validity / image bookkeeping); lane = a fixed LOGICAL 16-byte chunk (8
        // channels: its 16 coefficients stay in registers while the image does not change) of every 4th pixel, four pixels in
        // flight.  The transform of a tile is a latency chain (LDS -> affine -> SiLU -> STS) that the MMA warp waits for, so it
        // is spread over rows and unrolled rather than made instruction-lean only.  The affine part runs in fp32, SiLU on the
        // bf16x2-rounded pair: 0.5 t (1 + tanh(0.5 t)) = HMUL2 + MUFU + HFMA2.
        const int tw = warp - kXformWarp0;                // 0..7
        const int j = lane & 7, pp = lane >> 3;           // logical chunk, pixel phase 0..3
        int sa = 0, pa = 0;
        int coef_key = -1;
        float sc[8], sh[8];
        for (int ti = 0; ti < t_count; ++ti) {
            const int tile = tile_at(ti);
            const int Q0 = tile * 128;
            const int L0 = P.pw_magic ? floordiv_magic(Q0 - P.PW - 1, P.PW, P.pw_magic) : floordiv_dev(Q0 - P.PW - 1, P.PW);
            const int own0 = Q0 - L0 * P.PW;              // tile rows [own0, own0 + 128) are this tile's own output positions
            for (int c = 0; c < P.chunks; ++c) {
                long long* tdbg = (P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tw == 0 && lane == 0 && ti < 64)
                                      ? P.dbg + 8 * 64 + 8 * ti : nullptr;
                if (tdbg) tdbg[0] = clock64();
                mbar_wait(&a_full[sa], pa);
                if (tdbg) tdbg[1] = clock64();
                const uint32_t st = smem_u32(smem_a + (size_t)sa * P.a_stage_bytes);
                for (int i = tw; i < P.NR; i += 8) {
                    const int L = L0 + i;
                    const int n = floordiv_dev(L, P.PH), hp = L - n * P.PH;
                    if (n < 0 || n >= P.N || hp < 1 || hp > P.H) continue;
                    if (n * P.chunks + c != coef_key) {
                        const float4* cf = reinterpret_cast<const float4*>(P.gn_coef + ((size_t)n * P.Ck + c * 64 + j * 8) * 2);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 v = __ldg(cf + k);
                            sc[2 * k] = v.x; sh[2 * k] = v.y; sc[2 * k + 1] = v.z; sh[2 * k + 1] = v.w;
                        }
                        coef_key = n * P.chunks + c;
                    }
                    __nv_bfloat16* arow = (P.a_out && blockIdx.y == 0) ? P.a_out + (int64_t)n * P.a_sn + (int64_t)(hp - 1) * P.a_sh + c * 64 + j * 8 - P.a_sw
                                                                        : nullptr;
                    const int rbase = i * P.PW;
                    for (int wp0 = 1 + pp; wp0 <= P.W; wp0 += 16) {
                        uint4 raw[4];
                        uint32_t cell[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int r = rbase + wp0 + 4 * u;
                            cell[u] = st + (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4));
                            if (wp0 + 4 * u <= P.W) raw[u] = lds128(cell[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int wp = wp0 + 4 * u, r = rbase + wp;
                            if (wp <= P.W) {
                                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
                                uint4 outv;
                                __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&outv);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float2 f = __bfloat1622float2(h2[k]);
                                    __nv_bfloat162 t2 = __floats2bfloat162_rn(fmaf(f.x, sc[2 * k], sh[2 * k]), fmaf(f.y, sc[2 * k + 1], sh[2 * k + 1]));
                                    if (P.gn_silu) {
                                        const __nv_bfloat162 hf = __hmul2(t2, __floats2bfloat162_rn(0.5f, 0.5f));
                                        uint32_t th;
                                        asm("tanh.approx.bf16x2 %0, %1;" : "=r"(th) : "r"(*reinterpret_cast<const uint32_t*>(&hf)));
                                        t2 = __hfma2(hf, *reinterpret_cast<const __nv_bfloat162*>(&th), hf);
                                    }
                                    o2[k] = t2;
                                }
                                sts128(cell[u], outv);
                                if (arow && r >= own0 && r < own0 + 128) *reinterpret_cast<uint4*>(arow + (int64_t)wp * P.a_sw) = outv;
                            }
                        }
                    }
                }
                if (tdbg) tdbg[2] = clock64();
                fence_proxy_async();      // generic-proxy writes -> visible to the tensor core's async proxy
                __syncwarp();
                if (elect_one()) mbar_arrive(&a_ready[sa]);
                if (tdbg) tdbg[3] = clock64();
                __syncwarp();
                if (++sa == P.a_stages) { sa = 0; pa ^= 1; }
            }
        }
    } else {
        // ---------------------------------------------------- epilogue warps: thread = one output position (TMEM lane) x CH channels
        constexpr int CH = NT / (kEpiWarps / 4);             // channels per thread: the tile's channels split over warp w and warp w + 4
        constexpr int kPre = CH < 64 ? CH : 64;               // residual channels prefetched into registers
        const int row = threadIdx.x & 127;
        const int cb = (int)(threadIdx.x >> 7) * CH;          // this thread's first channel inside the tile
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int plane = P.PH * P.PW;
        // GroupNorm statistics of the output (st_sums, NT = 64): every thread keeps the sums of its rows per channel PAIR (a group is
        // >= 2 channels) for the image it is in; only when some lane of the warp moves on to another image - every plane / 128
        // tiles, the CTA's tiles being one contiguous run - are the warp's sums reduced and added to st_sums.  (Reducing every tile
        // cost the 64 x 64 layers 40 us each: the epilogue warps then no longer keep up with the MMA stream.)
        constexpr int kPairs = (NT == 64 && !GN) ? CH / 2 : 1;
        float acc_s[kPairs], acc_q[kPairs];
        int acc_img = -1;
#pragma unroll
        for (int i = 0; i < kPairs; ++i) { acc_s[i] = 0.f; acc_q[i] = 0.f; }
        auto stats_flush = [&]() {
            if constexpr (NT == 64 && !GN) {
                // the warp's lanes hold sums of at most two images.  A transposing butterfly - at offset o a lane keeps the half of
                // its values its bit o selects and adds the partner's copy of that half - leaves lane L with channel pair
                // L % kPairs's (sum, sum of squares) over the lanes that differ from L in the bits walked (all of them once the
                // halves of the warp are added, kPairs < 32); the lanes of a group then combine and one of them adds to st_sums.
                const int nA = __reduce_min_sync(0xffffffffu, acc_img >= 0 ? acc_img : 0x7fffffff);
                const int nB = __reduce_max_sync(0xffffffffu, acc_img);
#pragma unroll 1
                for (int im = nA; im <= nB && nB >= 0; ++im) {
                    const bool mine = acc_img == im;
                    float w[2 * kPairs];
#pragma unroll
                    for (int i = 0; i < kPairs; ++i) { w[2 * i] = mine ? acc_s[i] : 0.f; w[2 * i + 1] = mine ? acc_q[i] : 0.f; }
#pragma unroll
                    for (int half = kPairs; half >= 2; half >>= 1) {
                        const bool up = (lane & (half >> 1)) != 0;
#pragma unroll
                        for (int i = 0; i < half; ++i) {
                            const float keep = up ? w[half + i] : w[i];
                            const float send = up ? w[i] : w[half + i];
                            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, half >> 1);
                        }
                    }
                    float su = w[0], sq = w[1];
#pragma unroll
                    for (int o = kPairs; o < 32; o <<= 1) {      // lanes that hold the same pair
                        su += __shfl_xor_sync(0xffffffffu, su, o);
                        sq += __shfl_xor_sync(0xffffffffu, sq, o);
                    }
                    const int ppg = 1 << (P.st_sh - 1);          // channel pairs per group
                    for (int o = 1; o < ppg; o <<= 1) {
                        su += __shfl_xor_sync(0xffffffffu, su, o);
                        sq += __shfl_xor_sync(0xffffffffu, sq, o);
                    }
                    if ((lane & (ppg - 1)) == 0 && lane < kPairs) {
                        const int64_t si = ((int64_t)im * P.st_G + ((j0 + cb + 2 * lane) >> P.st_sh)) * 2;
                        if (P.st_fixed) {
                            atomicAdd(P.st_fixed + si, gn_to_fixed(su));
                            atomicAdd(P.st_fixed + si + 1, gn_to_fixed(sq));
                        } else {
                            atomicAdd(P.st_sums + si, su);
                            atomicAdd(P.st_sums + si + 1, sq);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < kPairs; ++i) { acc_s[i] = 0.f; acc_q[i] = 0.f; }
                acc_img = -1;
            }
        };
        int it = 0;
        for (; it < t_count; ++it) {
            const int tile = tile_at(it);
            const int buf = it & 1;
            const int Q = tile * 128 + row;
            const int n = Q / plane, rem = Q - n * plane;
            const int hp = rem / P.PW, wp = rem - hp * P.PW;
            const bool valid = n < P.N && hp >= 1 && hp <= P.H && wp >= 1 && wp <= P.W;
            const int ho = hp - 1, wo = wp - 1;
            if (P.st_sums) {
                if (__any_sync(0xffffffffu, valid && acc_img >= 0 && acc_img != n)) stats_flush();
                if (valid) acc_img = n;
            }
            __nv_bfloat16* yp = P.y + (int64_t)n * P.y_sn + (int64_t)ho * P.y_sh + (int64_t)wo * P.y_sw + j0 + cb;
            const __nv_bfloat16* rp = (P.res && valid) ? P.res + (int64_t)n * P.r_sn + (int64_t)ho * P.r_sh + (int64_t)wo * P.r_sw + j0 + cb : nullptr;
            const float* tp = (P.temb && valid) ? P.temb + (int64_t)n * P.temb_pitch + j0 + cb : nullptr;
            // prefetch this thread's (first 64) residual channels: the loads fly while the MMAs of this tile run
            uint4 rpre[kPre / 8];
            if (rp) {
#pragma unroll
                for (int i = 0; i < kPre / 8; ++i) rpre[i] = *reinterpret_cast<const uint4*>(rp + i * 8);
            }
            long long* dbg = (P.dbg && blockIdx.x == 0 && blockIdx.y == 0 && it < 64 && threadIdx.x == 0) ? P.dbg + 8 * it : nullptr;
            if (dbg) dbg[4] = clock64();
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after();
            if (dbg) dbg[5] = clock64();
            if (P.narrow) {
                if (cb == 0) {
                    float v[32];
                    tmem_ld32(tmem + lane_base + (uint32_t)(buf * NT), v);
                    tmem_ld_wait();
                    if (valid) {
                        float* op = P.yn + (int64_t)n * P.n_sn + (int64_t)ho * P.n_sh + (int64_t)wo * P.n_sw;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < P.narrow) op[(int64_t)j * P.n_sc] = v[j] + s_bias[j];
                    }
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[buf]);
                continue;
            }
#pragma unroll
            for (int c = 0; c < CH; c += 32) {
                float v[32];
                tmem_ld32(tmem + lane_base + (uint32_t)(buf * NT + cb + c), v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += s_bias[cb + c + i];
                    if (tp) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(tp + c + i));
                            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                        }
                    }
                    if (rp) {
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            float r8[8];
                            if (c < kPre) {
                                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rpre[(c + i) >> 3]);
#pragma unroll
                                for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h2[k]); r8[2 * k] = f.x; r8[2 * k + 1] = f.y; }
                            } else {
                                load_vec<__nv_bfloat16>(rp + c + i, r8);
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[i + k] += r8[k];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));      // what the store below writes
#pragma unroll
                    for (int i = 0; i < 32; i += 8) store_vec<__nv_bfloat16>(yp + c + i, v + i);
                }
                if constexpr (NT == 64 && !GN) {
                    if (P.st_sums && valid) {      // this row's pair sums of the stored values (flushed when the warp moves on to another image)
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            acc_s[(c + i) >> 1] += v[i] + v[i + 1];
                            acc_q[(c + i) >> 1] = fmaf(v[i], v[i], fmaf(v[i + 1], v[i + 1], acc_q[(c + i) >> 1]));
                        }
                    }
                }
            }
            if (dbg) dbg[6] = clock64();
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);      // every epilogue thread arrives: the accumulator goes back to the MMA warp
        }
        if (P.st_sums) stats_flush();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem, 2 * NT);
}

// ================================================================================================ 4x4 stride-2 transposed gather
// ConvTranspose2d(k = 4, s = 2, p = 1) and the input gradient of Conv2d(k = 4, s = 2, p = 1) (dmu_conv_params.gather = 1):
//   y[n, 2 th + a, 2 tw + b, :] = sum over the four taps (r, s) with r = 1 - a (mod 2), s = 1 - b (mod 2) of x[n, th + dh, tw + dw, :] W[r][s],
//   dh = (a + 1 - r) / 2 in {-1, 0, 1}.
// The per-tap kernel runs this as four launches' worth of CTAs (one per output parity class) whose k-loops are 4 k-blocks long: at
// 256x32x32 -> 64x64 that is 8192 CTAs bound by their fixed cost (190 us, 0.13 of the tensor peak).  Here the tile is 128 consecutive
// positions of the zero-padded flat INPUT space, in which the nine (dh, dw) are pure shifts of ONE halo tile (as in the 3x3 kernel
// above), the sixteen 64 x 64 filter blocks stay resident in shared memory, and the four parity classes are four TMEM accumulators
// of the same tile: 64 MMAs per 128 input positions = 512 output pixels, every input pixel fetched once per CTA.
// Epilogue: thread = input position x output row parity a; it writes the two horizontally adjacent output pixels (b = 0, 1).
struct HaloTArgs {
    int N, H, W, Ck, Cj;            // H, W: input extent (output 2H x 2W)
    int PW, PH, NR, tiles;
    int a_stage_bytes, a_stages;
    __nv_bfloat16* y; int64_t y_sn, y_sh, y_sw;
    const float* bias;
    uint32_t pw_magic;
};

__global__ void __launch_bounds__(320, 1) conv4x4t_halo_kernel(const __grid_constant__ HaloMaps maps, const __grid_constant__ HaloTArgs P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t a_full[kMaxAStages], a_empty[kMaxAStages], w_full, acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[64];
    constexpr int kWTile = 64 * 128, kEpiWarps = 8, kTmaWarp = 8, kMmaWarp = 9;
    uint8_t* smem_a = smem;
    uint8_t* smem_w = smem + (size_t)P.a_stages * P.a_stage_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j0 = blockIdx.y * 64;
    int t_count = 0;
    for (int t = blockIdx.x; t < P.tiles; t += gridDim.x) ++t_count;

    if (threadIdx.x == 0) {
        for (int i = 0; i < P.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        mbar_init(&w_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps * 32); }
        fence_mbar_init();
        tma_prefetch_desc(&maps.a);
        tma_prefetch_desc(&maps.b);
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_bias[i] = (P.bias && j0 + i < P.Cj) ? P.bias[j0 + i] : 0.f;
    // This kernel takes ALL 512 TMEM columns of its SM (2 buffers x 4 parity classes x 64).  Under programmatic dependent launch a
    // CTA of the previous kernel may still be on the SM and may not have allocated yet, and a CTA of the next kernel may arrive and
    // allocate before this one: either would leave two CTAs waiting for each other.  So: allocate only once the previous grid has
    // completed (its columns are free), and let the dependents in only after the allocation.
    pdl_wait();
    if (warp == kMmaWarp) tmem_alloc(&s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    pdl_trigger();

    if (warp == kTmaWarp) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&w_full, (uint32_t)(16 * kWTile));
            for (int t = 0; t < 16; ++t) tma_load_2d(smem_w + (size_t)t * kWTile, &maps.b, &w_full, t * P.Ck, j0);
            int sa = 0, pa = 1;
            const uint32_t a_bytes = (uint32_t)(P.NR * P.PW) * 128u;
            for (int ti = 0; ti < t_count; ++ti) {
                const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
                const int L0 = P.pw_magic ? floordiv_magic(tile * 128 - P.PW - 1, P.PW, P.pw_magic) : floordiv_dev(tile * 128 - P.PW - 1, P.PW);
                mbar_wait(&a_empty[sa], pa);
                mbar_arrive_expect_tx(&a_full[sa], a_bytes);
                uint8_t* dst = smem_a + (size_t)sa * P.a_stage_bytes;
                int n = floordiv_dev(L0, P.PH), hp = L0 - n * P.PH;
                for (int i = 0; i < P.NR; ++i) {
                    tma_load_4d(dst + (size_t)i * P.PW * 128, &maps.a, &a_full[sa], 0, -1, hp - 1, n);
                    if (++hp == P.PH) { hp = 0; ++n; }
                }
                if (++sa == P.a_stages) { sa = 0; pa ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == kMmaWarp) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
            mbar_wait(&w_full, 0);
            const uint32_t w_base = smem_u32(smem_w);
            int sa = 0, pa = 0;
            for (int it = 0; it < t_count; ++it) {
                const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                const int Q0 = tile * 128;
                const int L0 = P.pw_magic ? floordiv_magic(Q0 - P.PW - 1, P.PW, P.pw_magic) : floordiv_dev(Q0 - P.PW - 1, P.PW);
                const int base_off = Q0 - P.PW - 1 - L0 * P.PW;       // halo row of the (dh, dw) = (-1, -1) neighbour of position Q0
                mbar_wait(&a_full[sa], pa);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem_a + (size_t)sa * P.a_stage_bytes) + (uint32_t)base_off * 128u;
#pragma unroll
                for (int ph = 0; ph < 4; ++ph) {
                    const int a = ph >> 1, b = ph & 1;
                    const uint32_t d_tmem = tmem + (uint32_t)(buf * 256 + ph * 64);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        // parity 0: r = 1 (dh = 0), r = 3 (dh = -1); parity 1: r = 0 (dh = +1), r = 2 (dh = 0)
                        const int r = (a ? 0 : 1) + 2 * (q >> 1), sx = (b ? 0 : 1) + 2 * (q & 1);
                        const int dh = (a + 1 - r) / 2, dw = (b + 1 - sx) / 2;      // exact: the numerators are even
                        const uint64_t da = smem_desc_sw128(a_base + (uint32_t)((dh + 1) * P.PW + (dw + 1)) * 128u, 16, 1024);
                        const uint64_t db = smem_desc_sw128(w_base + (uint32_t)((r * 4 + sx) * kWTile), 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (q | k) != 0);
                    }
                }
                umma_commit(&a_empty[sa]);
                umma_commit(&acc_full[buf]);
                if (++sa == P.a_stages) { sa = 0; pa ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------- epilogue warps 0..7: thread = input position x output row parity
        const int row = threadIdx.x & 127, a = (int)(threadIdx.x >> 7);
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int plane = P.PH * P.PW;
        for (int it = 0; it < t_count; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int buf = it & 1;
            const int Q = tile * 128 + row;
            const int n = Q / plane, rem = Q - n * plane;
            const int hp = rem / P.PW, wp = rem - hp * P.PW;
            const bool valid = n < P.N && hp >= 1 && hp <= P.H && wp >= 1 && wp <= P.W;
            // the two output pixels (2 th + a, 2 tw), (2 th + a, 2 tw + 1)
            __nv_bfloat16* yp = P.y + (int64_t)n * P.y_sn + (int64_t)(2 * (hp - 1) + a) * P.y_sh + (int64_t)(2 * (wp - 1)) * P.y_sw + j0;
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 128; c += 32) {        // columns: parity class (a, b = c / 64) x 64 channels
                float v[32];
                tmem_ld32(tmem + lane_base + (uint32_t)(buf * 256 + a * 128 + c), v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += s_bias[(c & 63) + i];
                    __nv_bfloat16* op = yp + (int64_t)(c >> 6) * P.y_sw + (c & 63);
#pragma unroll
                    for (int i = 0; i < 32; i += 8) store_vec<__nv_bfloat16>(op + i, v + i);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
}

static int halo_t_geometry(const dmu_conv_params* p, HaloTArgs& A) {
    memset(&A, 0, sizeof(A));
    A.N = p->N; A.H = p->Hi; A.W = p->Wi; A.Ck = p->Ck; A.Cj = p->Cj;
    A.PW = A.W + 2; A.PH = A.H + 2;
    A.NR = 3 + (129 + A.PW - 1) / A.PW;
    A.tiles = (int)(((int64_t)A.N * A.PH * A.PW + 127) / 128);
    A.a_stage_bytes = ((A.NR * A.PW * 128) + 1023) / 1024 * 1024;
    const int budget = 214 * 1024, wbytes = 16 * 64 * 128;
    A.a_stages = (budget - wbytes) / A.a_stage_bytes;
    if (A.a_stages > kMaxAStages) A.a_stages = kMaxAStages;
    if (A.a_stages < 2) return -1;
    return A.a_stages * A.a_stage_bytes + wbytes + 1024;
}

int halo_t_supported(const dmu_conv_params* p, int force) {
    if (p->gather != 1 || p->R != 4 || p->S != 4 || p->stride != 2 || p->pad != 1) return 0;
    if (p->Ho != 2 * p->Hi || p->Wo != 2 * p->Wi) return 0;
    if (p->Hi < 4 || p->Wi < 4 || p->Wi + 2 > 256) return 0;
    if (p->Ck != 64 || p->Cj % 64 != 0) return 0;             // the sixteen filter blocks of one 64-channel chunk stay resident
    if (p->res.ptr || p->temb || p->gn_coef || p->gn_fuse_mode) return 0;
    if (p->w_sk != 1 || p->w_st != p->Ck || p->w_sn != (int64_t)16 * p->Ck) return 0;
    if ((int64_t)p->N * (p->Hi + 2) * (p->Wi + 2) >= (1ll << 31) - 4096) return 0;
    HaloTArgs A;
    if (halo_t_geometry(p, A) <= 0) return 0;
    if (force) return 1;
    static const int enabled = [] { const char* e = getenv("DMU_HALO_T"); return e ? atoi(e) : 1; }();
    // a CTA loads 128 KB of filters before its first tile: measured worth it from about two tiles per CTA (training step: 324 tiles yes, 100 no)
    return enabled && A.tiles >= 300 ? 1 : 0;
}

int halo_t_launch(const dmu_conv_params* p, cudaStream_t stream) {
    HaloMaps maps;
    HaloTArgs A;
    const int smem = halo_t_geometry(p, A);
    DMU_REQUIRE(smem > 0 && smem <= 224 * 1024, "dmu_conv2d/halo_t: tile does not fit shared memory");
    {
        const uint64_t dims[4] = {(uint64_t)p->Ck, (uint64_t)p->Wi, (uint64_t)p->Hi, (uint64_t)p->N};
        const uint64_t str[4] = {1, (uint64_t)p->x.sw, (uint64_t)p->x.sh, (uint64_t)p->x.sn};
        const uint32_t box[4] = {64, (uint32_t)A.PW, 1, 1};
        if (int rc = make_map_bf16(&maps.a, p->x.ptr, 4, dims, str, box, "dmu_conv2d/halo_t")) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)16 * p->Ck, (uint64_t)p->Cj};
        const uint64_t str[2] = {1, (uint64_t)p->w_sn};
        const uint32_t box[2] = {64, 64};
        if (int rc = make_map_bf16(&maps.b, p->w, 2, dims, str, box, "dmu_conv2d/halo_t weights")) return rc;
    }
    A.y = reinterpret_cast<__nv_bfloat16*>(p->y.ptr); A.y_sn = p->y.sn; A.y_sh = p->y.sh; A.y_sw = p->y.sw;
    A.bias = p->bias;
    if (((int64_t)A.N * A.PH * A.PW + 2 * A.PW + 256) * A.PW < (1ll << 32)) A.pw_magic = (uint32_t)((1ull << 32) / (uint64_t)A.PW) + 1u;
    const int ny = p->Cj / 64;
    int gx = sm_count() / ny;
    if (gx < 1) gx = 1;
    if (gx > A.tiles) gx = A.tiles;
    const int rounds = (A.tiles + gx - 1) / gx;
    gx = (A.tiles + rounds - 1) / rounds;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv4x4t_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        attr_done = true;
    }
    cudaError_t e = launch_pdl(conv4x4t_halo_kernel, dim3(gx, ny), dim3(320), (size_t)smem, stream, dim3(1, 1, 1), maps, A);
    if (e != cudaSuccess) return fail("dmu_conv2d/halo_t: launch failed: %s", cudaGetErrorString(e));
    return check_launch("dmu_conv2d/halo_t");
}

// ================================================================================================ 4x4 stride-2 convolution
// Conv2d(k = 4, s = 2, p = 1) (the learned downsampling) and the input gradient of ConvTranspose2d(k = 4, s = 2, p = 1)
// (dmu_conv_params.gather = 0):  y[n, ho, wo, :] = sum_{r, s} x[n, 2 ho - 1 + r, 2 wo - 1 + s, :] W[r][s].
// Row 2 ho - 1 + r lies on the parity sub-lattice ph = (r + 1) mod 2 of the input at index ho + dh, dh = floor((r - 1) / 2): on each of
// the four sub-lattices X[ph][pw][n, i, j] = x[n, 2 i + ph, 2 j + pw] (a strided tensor map) the layer is a 2 x 2 convolution
// whose taps are pure shifts in the zero-padded flat space of the OUTPUT grid.  The kernel walks the four sub-lattices like the
// 3x3 kernel walks 64-channel chunks: one halo tile per sub-lattice and output tile (a ring stage), four taps each, all sixteen
// filter blocks resident, one accumulator per tile.  The per-tap kernel pulls sixteen 16 KB boxes per tile through the SM's L2
// port instead of four ~25 KB halo tiles (measured at 256x64x64 -> 32x32: 90 us, 0.27 of the tensor peak).
struct HaloSMaps { CUtensorMap a[4]; CUtensorMap b; };
struct HaloSArgs {
    int N, H, W, Ck, Cj;            // H, W: OUTPUT extent (input 2H x 2W)
    int PW, PH, NR, tiles;
    int a_stage_bytes, a_stages;
    __nv_bfloat16* y; int64_t y_sn, y_sh, y_sw;
    const float* bias;
    uint32_t pw_magic;
};

__global__ void __launch_bounds__(320, 1) conv4x4s2_halo_kernel(const __grid_constant__ HaloSMaps maps, const __grid_constant__ HaloSArgs P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t a_full[kMaxAStages], a_empty[kMaxAStages], w_full, acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[64];
    constexpr int kWTile = 64 * 128, kEpiWarps = 8, kTmaWarp = 8, kMmaWarp = 9;
    uint8_t* smem_a = smem;
    uint8_t* smem_w = smem + (size_t)P.a_stages * P.a_stage_bytes;
    const int warp = threadIdx.x >> 5;
    const int j0 = blockIdx.y * 64;
    int t_count = 0;
    for (int t = blockIdx.x; t < P.tiles; t += gridDim.x) ++t_count;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        mbar_init(&w_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps * 32); }
        fence_mbar_init();
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.a[i]);
        tma_prefetch_desc(&maps.b);
    }
    if (warp == kMmaWarp) tmem_alloc(&s_tmem, 128);          // 2 buffers x 64 columns
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_bias[i] = (P.bias && j0 + i < P.Cj) ? P.bias[j0 + i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    pdl_wait();

    if (warp == kTmaWarp) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&w_full, (uint32_t)(16 * kWTile));
            for (int t = 0; t < 16; ++t) tma_load_2d(smem_w + (size_t)t * kWTile, &maps.b, &w_full, t * P.Ck, j0);
            int sa = 0, pa = 1;
            const uint32_t a_bytes = (uint32_t)(P.NR * P.PW) * 128u;
            for (int ti = 0; ti < t_count; ++ti) {
                const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
                const int L0 = P.pw_magic ? floordiv_magic(tile * 128 - P.PW - 1, P.PW, P.pw_magic) : floordiv_dev(tile * 128 - P.PW - 1, P.PW);
                const int n0 = floordiv_dev(L0, P.PH), hp0 = L0 - n0 * P.PH;
                for (int sub = 0; sub < 4; ++sub) {
                    mbar_wait(&a_empty[sa], pa);
                    mbar_arrive_expect_tx(&a_full[sa], a_bytes);
                    uint8_t* dst = smem_a + (size_t)sa * P.a_stage_bytes;
                    int n = n0, hp = hp0;
                    for (int i = 0; i < P.NR; ++i) {
                        tma_load_4d(dst + (size_t)i * P.PW * 128, &maps.a[sub], &a_full[sa], 0, -1, hp - 1, n);
                        if (++hp == P.PH) { hp = 0; ++n; }
                    }
                    if (++sa == P.a_stages) { sa = 0; pa ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == kMmaWarp) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
            mbar_wait(&w_full, 0);
            const uint32_t w_base = smem_u32(smem_w);
            int sa = 0, pa = 0;
            for (int it = 0; it < t_count; ++it) {
                const int tile = (int)blockIdx.x + it * (int)gridDim.x;
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                const uint32_t d_tmem = tmem + (uint32_t)(buf * 64);
                const int Q0 = tile * 128;
                const int L0 = P.pw_magic ? floordiv_magic(Q0 - P.PW - 1, P.PW, P.pw_magic) : floordiv_dev(Q0 - P.PW - 1, P.PW);
                const int base_off = Q0 - P.PW - 1 - L0 * P.PW;       // halo row of the (dh, dw) = (-1, -1) neighbour of position Q0
#pragma unroll
                for (int sub = 0; sub < 4; ++sub) {
                    const int ph = sub >> 1, pw = sub & 1;
                    mbar_wait(&a_full[sa], pa);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(smem_a + (size_t)sa * P.a_stage_bytes) + (uint32_t)base_off * 128u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        // even sub-lattice: r = 1 (dh = 0), r = 3 (dh = +1); odd sub-lattice: r = 0 (dh = -1), r = 2 (dh = 0)
                        const int r = (ph ? 0 : 1) + 2 * (q >> 1), sx = (pw ? 0 : 1) + 2 * (q & 1);
                        const int dh = (r - 1 - ph) / 2, dw = (sx - 1 - pw) / 2;      // exact: the numerators are even
                        const uint64_t da = smem_desc_sw128(a_base + (uint32_t)((dh + 1) * P.PW + (dw + 1)) * 128u, 16, 1024);
                        const uint64_t db = smem_desc_sw128(w_base + (uint32_t)((r * 4 + sx) * kWTile), 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (sub | q | k) != 0);
                    }
                    umma_commit(&a_empty[sa]);
                    if (++sa == P.a_stages) { sa = 0; pa ^= 1; }
                }
                umma_commit(&acc_full[buf]);
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------- epilogue warps 0..7: thread = output position x 32 channels
        const int row = threadIdx.x & 127, cb = (int)(threadIdx.x >> 7) * 32;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int plane = P.PH * P.PW;
        for (int it = 0; it < t_count; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int buf = it & 1;
            const int Q = tile * 128 + row;
            const int n = Q / plane, rem = Q - n * plane;
            const int hp = rem / P.PW, wp = rem - hp * P.PW;
            const bool valid = n < P.N && hp >= 1 && hp <= P.H && wp >= 1 && wp <= P.W;
            __nv_bfloat16* yp = P.y + (int64_t)n * P.y_sn + (int64_t)(hp - 1) * P.y_sh + (int64_t)(wp - 1) * P.y_sw + j0 + cb;
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after();
            float v[32];
            tmem_ld32(tmem + lane_base + (uint32_t)(buf * 64 + cb), v);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);      // the values are in registers: the accumulator goes back before the stores
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += s_bias[cb + i];
#pragma unroll
                for (int i = 0; i < 32; i += 8) store_vec<__nv_bfloat16>(yp + i, v + i);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem, 128);
}

static int halo_s_geometry(const dmu_conv_params* p, HaloSArgs& A) {
    memset(&A, 0, sizeof(A));
    A.N = p->N; A.H = p->Ho; A.W = p->Wo; A.Ck = p->Ck; A.Cj = p->Cj;
    A.PW = A.W + 2; A.PH = A.H + 2;
    A.NR = 3 + (129 + A.PW - 1) / A.PW;
    A.tiles = (int)(((int64_t)A.N * A.PH * A.PW + 127) / 128);
    A.a_stage_bytes = ((A.NR * A.PW * 128) + 1023) / 1024 * 1024;
    const int budget = 214 * 1024, wbytes = 16 * 64 * 128;
    A.a_stages = (budget - wbytes) / A.a_stage_bytes;
    if (A.a_stages > kMaxAStages) A.a_stages = kMaxAStages;
    if (A.a_stages < 2) return -1;
    return A.a_stages * A.a_stage_bytes + wbytes + 1024;
}

int halo_s_supported(const dmu_conv_params* p, int force) {
    if (p->gather != 0 || p->R != 4 || p->S != 4 || p->stride != 2 || p->pad != 1) return 0;
    if (p->Hi != 2 * p->Ho || p->Wi != 2 * p->Wo) return 0;
    if (p->Ho < 4 || p->Wo < 4 || p->Wo + 2 > 256) return 0;
    if (p->Ck != 64 || p->Cj % 64 != 0) return 0;             // the sixteen filter blocks of one 64-channel chunk stay resident
    if (p->res.ptr || p->temb || p->gn_coef || p->gn_fuse_mode) return 0;
    if (p->w_sk != 1 || p->w_st != p->Ck || p->w_sn != (int64_t)16 * p->Ck) return 0;
    if ((int64_t)p->N * (p->Ho + 2) * (p->Wo + 2) >= (1ll << 31) - 4096) return 0;
    HaloSArgs A;
    if (halo_s_geometry(p, A) <= 0) return 0;
    if (force) return 1;
    static const int enabled = [] { const char* e = getenv("DMU_HALO_S"); return e ? atoi(e) : 1; }();
    return enabled && A.tiles >= 300 ? 1 : 0;      // as for the transposed kernel
}

int halo_s_launch(const dmu_conv_params* p, cudaStream_t stream) {
    HaloSMaps maps;
    HaloSArgs A;
    const int smem = halo_s_geometry(p, A);
    DMU_REQUIRE(smem > 0 && smem <= 224 * 1024, "dmu_conv2d/halo_s: tile does not fit shared memory");
    for (int sub = 0; sub < 4; ++sub) {      // parity sub-lattice (ph, pw) of the input: [N][Ho][Wo] at twice the strides
        const int ph = sub >> 1, pw = sub & 1;
        const uint64_t dims[4] = {(uint64_t)p->Ck, (uint64_t)p->Wo, (uint64_t)p->Ho, (uint64_t)p->N};
        const uint64_t str[4] = {1, (uint64_t)p->x.sw * 2, (uint64_t)p->x.sh * 2, (uint64_t)p->x.sn};
        const uint32_t box[4] = {64, (uint32_t)A.PW, 1, 1};
        const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(p->x.ptr) + (int64_t)ph * p->x.sh + (int64_t)pw * p->x.sw;
        if (int rc = make_map_bf16(&maps.a[sub], base, 4, dims, str, box, "dmu_conv2d/halo_s")) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)16 * p->Ck, (uint64_t)p->Cj};
        const uint64_t str[2] = {1, (uint64_t)p->w_sn};
        const uint32_t box[2] = {64, 64};
        if (int rc = make_map_bf16(&maps.b, p->w, 2, dims, str, box, "dmu_conv2d/halo_s weights")) return rc;
    }
    A.y = reinterpret_cast<__nv_bfloat16*>(p->y.ptr); A.y_sn = p->y.sn; A.y_sh = p->y.sh; A.y_sw = p->y.sw;
    A.bias = p->bias;
    if (((int64_t)A.N * A.PH * A.PW + 2 * A.PW + 256) * A.PW < (1ll << 32)) A.pw_magic = (uint32_t)((1ull << 32) / (uint64_t)A.PW) + 1u;
    const int ny = p->Cj / 64;
    int gx = sm_count() / ny;
    if (gx < 1) gx = 1;
    if (gx > A.tiles) gx = A.tiles;
    const int rounds = (A.tiles + gx - 1) / gx;
    gx = (A.tiles + rounds - 1) / rounds;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv4x4s2_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        attr_done = true;
    }
    cudaError_t e = launch_pdl(conv4x4s2_halo_kernel, dim3(gx, ny), dim3(320), (size_t)smem, stream, dim3(1, 1, 1), maps, A);
    if (e != cudaSuccess) return fail("dmu_conv2d/halo_s: launch failed: %s", cudaGetErrorString(e));
    return check_launch("dmu_conv2d/halo_s");
}

static int pick_smem(const dmu_conv_params* p, int NT, HaloArgs& A) {
    const int budget = 214 * 1024;
    const int kblocks = 9 * A.chunks, wtile = NT * 128;
    A.resident = (kblocks * wtile + 2 * A.a_stage_bytes <= budget) ? 1 : 0;
    int wbytes;
    if (A.resident) {
        wbytes = kblocks * wtile;
        A.w_stages = 1;
    } else {
        A.w_stages = kMaxWStages;
        while (A.w_stages > 3 && A.w_stages * wtile + 2 * A.a_stage_bytes > budget) --A.w_stages;
        wbytes = A.w_stages * wtile;
    }
    A.a_stages = (budget - wbytes) / A.a_stage_bytes;
    if (A.a_stages > kMaxAStages) A.a_stages = kMaxAStages;
    if (A.a_stages < 2) return -1;
    return A.a_stages * A.a_stage_bytes + wbytes + 1024;
}

// geometry shared by halo_supported() and halo_launch(); returns the dynamic shared memory size (or -1)
static int halo_geometry(const dmu_conv_params* p, HaloArgs& A, int& NT) {
    memset(&A, 0, sizeof(A));
    A.N = p->N; A.H = p->Hi; A.W = p->Wi; A.Ck = p->Ck; A.Cj = p->Cj;
    A.PW = A.W + 2; A.PH = A.H + 2;
    A.NR = 3 + (129 + A.PW - 1) / A.PW;
    A.tiles = (int)(((int64_t)A.N * A.PH * A.PW + 127) / 128);
    A.chunks = A.Ck / 64;
    A.flip = p->gather;
    A.a_stage_bytes = ((A.NR * A.PW * 128) + 1023) / 1024 * 1024;
    NT = p->Cj <= 4 ? 64 : (p->Cj % 128 == 0) ? 128 : 64;
    return pick_smem(p, NT, A);
}

int halo_supported(const dmu_conv_params* p, int force) {
    if (p->R != 3 || p->S != 3 || p->stride != 1 || p->pad != 1) return 0;
    if (p->Hi != p->Ho || p->Wi != p->Wo) return 0;
    if (p->Hi < 8 || p->Wi < 8 || p->Wi + 2 > 256) return 0;     // below 8x8 most of the padded space is padding
    if ((int64_t)p->N * (p->Hi + 2) * (p->Wi + 2) >= (1ll << 31)) return 0;
    if (force) return 1;
    // Measured on B200 (scripts/halo_vs_tap.py): with the filter bank of the CTA's output-channel tile RESIDENT in shared
    // memory the halo kernel reads every activation once and wins wherever each CTA gets a few tiles (64->64 @128x32x32:
    // 16.2 vs 28.2 us, 128->64: 26.4 vs 38.4 us, 64->128 @128x8x8: 6.3 vs 8.4 us); when the filters have to stream through
    // their ring once per tile the per-tap kernel, whose CTAs share them through L2, is faster (192->64 @128x16x16: 27 vs 14 us).
    HaloArgs A;
    int NT;
    if (halo_geometry(p, A, NT) <= 0 || !A.resident) return 0;
    if (p->Cj > 128) return 0;
    if (A.tiles < 4 * (int64_t)sm_count()) return 0;      // fewer tiles per CTA: the resident-filter prologue is not amortised
    return 1;
}

// gn_fuse_mode 3: can the halo kernel (chosen by its own heuristics) add the output's GroupNorm sums in its epilogue?
int halo_stats_supported(const dmu_conv_params* p) {
    const dmu_gn_params* gn = reinterpret_cast<const dmu_gn_params*>(p->gn_fuse);
    if (!gn || !gn->sums || gn->G <= 0 || gn->C != p->Cj || gn->C % gn->G != 0) return 0;
    if (gn->N != p->N || gn->H != p->Ho || gn->W != p->Wo) return 0;
    const int cpg = gn->C / gn->G;
    if (cpg < 2 || cpg > 32 || (cpg & (cpg - 1)) != 0 || p->Cj <= 4 || p->Cj % 128 == 0 || p->gn_coef) return 0;     // 64-channel tiles, pairs
    if ((p->Hi + 2) * (p->Wi + 2) < 128) return 0;          // a warp's 32 rows must lie in at most two images
    dmu_conv_params q = *p;
    q.gn_fuse = nullptr; q.gn_fuse_mode = 0;
    return halo_supported(&q, 0);
}

int halo_launch(const dmu_conv_params* p, cudaStream_t stream) {
    HaloMaps maps;
    HaloArgs A;
    int NT;
    const int smem = halo_geometry(p, A, NT);
    const bool narrow = p->Cj <= 4;
    DMU_REQUIRE(smem > 0 && smem <= 224 * 1024, "dmu_conv2d/halo: tile does not fit shared memory");
    {
        const uint64_t dims[4] = {(uint64_t)p->Ck, (uint64_t)p->Wi, (uint64_t)p->Hi, (uint64_t)p->N};
        const uint64_t str[4] = {1, (uint64_t)p->x.sw, (uint64_t)p->x.sh, (uint64_t)p->x.sn};
        const uint32_t box[4] = {64, (uint32_t)A.PW, 1, 1};
        if (int rc = make_map_bf16(&maps.a, p->x.ptr, 4, dims, str, box, "dmu_conv2d/halo")) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)9 * p->Ck, (uint64_t)p->Cj};
        const uint64_t str[2] = {1, (uint64_t)p->w_sn};
        const uint32_t box[2] = {64, (uint32_t)NT};
        if (int rc = make_map_bf16(&maps.b, p->w, 2, dims, str, box, "dmu_conv2d/halo weights")) return rc;
    }
    if (narrow) {
        A.narrow = p->Cj;
        A.yn = reinterpret_cast<float*>(p->y.ptr); A.n_sn = p->y.sn; A.n_sh = p->y.sh; A.n_sw = p->y.sw; A.n_sc = p->y.sc;
    }
    A.y = reinterpret_cast<__nv_bfloat16*>(p->y.ptr); A.y_sn = p->y.sn; A.y_sh = p->y.sh; A.y_sw = p->y.sw;
    A.res = reinterpret_cast<const __nv_bfloat16*>(p->res.ptr); A.r_sn = p->res.sn; A.r_sh = p->res.sh; A.r_sw = p->res.sw;
    A.bias = p->bias; A.temb = p->temb; A.temb_pitch = p->temb_pitch;
    A.dbg = g_debug_buffer;
    if (p->gn_fuse_mode == 3) {
        const dmu_gn_params* gn = reinterpret_cast<const dmu_gn_params*>(p->gn_fuse);
        DMU_REQUIRE(halo_stats_supported(p), "dmu_conv2d/halo: this launch cannot accumulate the GroupNorm statistics of gn_fuse (ask dmu_conv2d_gn_fuse_supported first)");
        const int cpg = gn->C / gn->G;
        A.st_sums = gn->sums; A.st_G = gn->G; A.st_sh = 0;
        A.st_fixed = (gn->flags & DMU_GN_FIXED_SUMS) ? reinterpret_cast<unsigned long long*>(gn->sums + (int64_t)gn->N * gn->G * 2) : nullptr;
        while ((1 << A.st_sh) < cpg) ++A.st_sh;
    }
    const int ntiles_n = narrow ? 1 : p->Cj / NT;
    int gx = sm_count() / ntiles_n;
    if (gx < 1) gx = 1;
    if (gx > A.tiles) gx = A.tiles;
    // even out the tail: every CTA gets the same number of tiles (+-1)
    const int rounds = (A.tiles + gx - 1) / gx;
    gx = (A.tiles + rounds - 1) / rounds;
    if (A.st_sums) {
        // runs of up to 8 consecutive tiles per CTA (the statistics are flushed when a thread's image changes), the longest run that
        // keeps the busiest CTA within 4 % of the average
        static const int run_max = [] { const char* e = getenv("DMU_HALO_STATS_RUN_SH"); return e ? atoi(e) : 3; }();      // A/B aid
        for (int sh = run_max; sh >= 0; --sh) {
            const int run = 1 << sh, runs = (A.tiles + run - 1) / run, per = (runs + gx - 1) / gx * run;
            if ((int64_t)per * gx * 100 <= (int64_t)A.tiles * 104 || sh == 0) { A.run_sh = sh; break; }
        }
    }
    A.run = 1 << A.run_sh;
    if (((int64_t)A.N * A.PH * A.PW + 2 * A.PW + 256) * A.PW < (1ll << 32)) A.pw_magic = (uint32_t)((1ull << 32) / (uint64_t)A.PW) + 1u;
    dim3 grid(gx, ntiles_n);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(conv3x3_halo_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        cudaFuncSetAttribute(conv3x3_halo_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        cudaFuncSetAttribute(conv3x3_halo_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        cudaFuncSetAttribute(conv3x3_halo_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        attr_done = true;
    }
    cudaError_t e;
    if (p->gn_coef) {
        A.gn_coef = p->gn_coef; A.gn_silu = p->gn_silu;
        A.a_out = reinterpret_cast<__nv_bfloat16*>(p->a_out.ptr); A.a_sn = p->a_out.sn; A.a_sh = p->a_out.sh; A.a_sw = p->a_out.sw;
        e = NT == 64 ? launch_pdl(conv3x3_halo_kernel<64, true>, grid, dim3(HaloCfg<true>::kThreads), (size_t)smem, stream, dim3(1, 1, 1), maps, A)
                     : launch_pdl(conv3x3_halo_kernel<128, true>, grid, dim3(HaloCfg<true>::kThreads), (size_t)smem, stream, dim3(1, 1, 1), maps, A);
    } else {
        e = NT == 64 ? launch_pdl(conv3x3_halo_kernel<64, false>, grid, dim3(HaloCfg<false>::kThreads), (size_t)smem, stream, dim3(1, 1, 1), maps, A)
                     : launch_pdl(conv3x3_halo_kernel<128, false>, grid, dim3(HaloCfg<false>::kThreads), (size_t)smem, stream, dim3(1, 1, 1), maps, A);
    }
    if (e != cudaSuccess) return fail("dmu_conv2d/halo: launch failed: %s", cudaGetErrorString(e));
    return check_launch("dmu_conv2d/halo");
}

}  // namespace tc
}  // namespace dmu
