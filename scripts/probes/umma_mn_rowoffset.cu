// Probe for the halo weight-gradient kernel (both operands MN-major: K = padded-flat positions = shared-memory rows of 128 B,
// M / N = 64 channels inside a row, SWIZZLE_128B as a TMA box lands it):
//  1. does an MN-major descriptor work when its start address is advanced by a number of rows that is not a multiple of 8
//     (the 3x3 taps as shifted windows of ONE halo tile)?  On the A side, on the B side, with base_offset 0 or (row & 7).
//  2. M = 128 as two OVERLAPPING 64-channel blocks (leading-dimension byte offset = 128 B = one row): two taps in one MMA?
//  3. throughput of the fp32 reductions the epilogue issues: red.global.add.f32 (a warp = 128 contiguous bytes),
//     red.global.add.v4.f32 with the lanes on different rows / on contiguous 16-byte cells, every CTA adding to the same block.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I diffusion_model_universal_b200/csrc -o /tmp/probe scripts/probes/umma_mn_rowoffset.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "tc_common.cuh"

using namespace dmu::tc;
namespace dmu { char* err_buf() { static char b[256]; return b; } int fail(const char* f, ...) { printf("fail: %s\n", f); return 1; } int sm_count() { return 148; } bool pdl_enabled() { return false; }
namespace tc {
EncodeTiledFn encode_tiled_fn() { void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); return (EncodeTiledFn)p; }
} }

constexpr int QROWS = 256, PROWS = 128, NOFF = 12, NVAR = NOFF * 4 + NOFF;   // (A-side, B-side) x (base 0, base row&7) + overlapped M=128
struct Maps2 { CUtensorMap q, p; };
struct Var { int off, side, base, overlap; };

__global__ void __launch_bounds__(128) probe(const __grid_constant__ Maps2 maps, float* out, const Var* vars) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                  // 256 positions x 64 channels (the shifted operand)
    uint8_t* sP = smem + QROWS * 128;    // 128 positions x 64 channels
    __shared__ __align__(8) uint64_t bar, mbar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&mbar, 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(&s_tmem, 64);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, QROWS * 128 + PROWS * 128);
        tma_load_2d(sQ, &maps.q, &bar, 0, 0);
        tma_load_2d(sP, &maps.p, &bar, 0, 0);
    }
    mbar_wait(&bar, 0);
    for (int v = 0; v < NVAR; ++v) {
        const Var V = vars[v];
        if (threadIdx.x == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(V.overlap ? 128 : 64, 64, 1, 1);
            const uint64_t dq = smem_desc_sw128(smem_u32(sQ) + V.off * 128, V.overlap ? 128 : 8192, 1024) | ((uint64_t)(V.base & 7) << 49);
            const uint64_t dp = smem_desc_sw128(smem_u32(sP), 8192, 1024);
            for (int k = 0; k < 8; ++k) {      // K = 128 positions, 16 per MMA = 2048 B
                if (V.side == 0) umma_bf16(tmem, dq + k * 128, dp + k * 128, idesc, k != 0);
                else umma_bf16(tmem, dp + k * 128, dq + k * 128, idesc, k != 0);
            }
            umma_commit(&mbar);
        }
        mbar_wait(&mbar, v & 1);
        tc_fence_after();
        for (int c = 0; c < 64; c += 32) {
            float r[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
            tmem_ld_wait();
            for (int i = 0; i < 32; ++i) out[((size_t)v * 128 + threadIdx.x) * 64 + c + i] = r[i];
        }
        tc_fence_before(); __syncthreads(); tc_fence_after();
    }
    if (warp == 1) tmem_dealloc(tmem, 64);
}

// mode 0: scalar red, lane = consecutive float; 1: v4, lane = row (stride 2304 B), thread = 4 consecutive floats;
// 2: v4, lanes on consecutive 16-byte cells; 3: plain v4 stores to a private slab (what a two-stage reduction would write)
__global__ void __launch_bounds__(128) red_probe(float* dst, float* slab, int mode, int per_thread_floats) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float v = 1.0f;
    if (mode == 0) {
        for (int i = 0; i < per_thread_floats; ++i) atomicAdd(dst + (size_t)(i * 4 + warp) * 32 + lane, v);
    } else if (mode == 1) {
        for (int i = 0; i < per_thread_floats / 4; ++i) red_add_v4(dst + (size_t)(warp * 32 + lane) * 576 + (i % 144) * 4, v, v, v, v);
    } else if (mode == 2) {
        for (int i = 0; i < per_thread_floats / 4; ++i) red_add_v4(dst + ((size_t)(i * 4 + warp) * 32 + lane) * 4, v, v, v, v);
    } else {
        float4* s = reinterpret_cast<float4*>(slab) + (size_t)blockIdx.x * (per_thread_floats / 4) * 128;
        for (int i = 0; i < per_thread_floats / 4; ++i) s[(size_t)i * 128 + t] = make_float4(v, v, v, v);
    }
}

int main() {
    std::vector<__nv_bfloat16> hQ(QROWS * 64), hP(PROWS * 64);
    for (int r = 0; r < QROWS; ++r) for (int k = 0; k < 64; ++k) hQ[r * 64 + k] = __float2bfloat16((float)(((r * 7 + k * 3) % 17) - 8) / 8.f);
    for (int r = 0; r < PROWS; ++r) for (int k = 0; k < 64; ++k) hP[r * 64 + k] = __float2bfloat16((float)(((r * 5 + k * 11) % 13) - 6) / 4.f);
    __nv_bfloat16 *dQ, *dP; float* dout; Var* dvars;
    cudaMalloc(&dQ, hQ.size() * 2); cudaMalloc(&dP, hP.size() * 2); cudaMalloc(&dout, (size_t)NVAR * 128 * 64 * 4); cudaMalloc(&dvars, NVAR * sizeof(Var));
    cudaMemcpy(dQ, hQ.data(), hQ.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dP, hP.data(), hP.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, (size_t)NVAR * 128 * 64 * 4);
    Var vars[NVAR];
    const int offs[NOFF] = {0, 1, 2, 3, 5, 7, 8, 9, 34, 35, 68, 69};
    int nv = 0;
    for (int side = 0; side < 2; ++side)
        for (int i = 0; i < NOFF; ++i) { vars[nv++] = Var{offs[i], side, 0, 0}; vars[nv++] = Var{offs[i], side, offs[i] & 7, 0}; }
    for (int i = 0; i < NOFF; ++i) vars[nv++] = Var{offs[i], 0, 0, 1};
    cudaMemcpy(dvars, vars, sizeof(vars), cudaMemcpyHostToDevice);
    Maps2 maps;
    { cuuint64_t gd[2] = {64, QROWS}, gs[1] = {128}; cuuint32_t bd[2] = {64, QROWS}, es[2] = {1, 1};
      CUresult r = encode_tiled_fn()(&maps.q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dQ, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode q failed %d\n", (int)r); return 1; } }
    { cuuint64_t gd[2] = {64, PROWS}, gs[1] = {128}; cuuint32_t bd[2] = {64, PROWS}, es[2] = {1, 1};
      CUresult r = encode_tiled_fn()(&maps.p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dP, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode p failed %d\n", (int)r); return 1; } }
    const int smem = QROWS * 128 + PROWS * 128 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(maps, dout, dvars);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hout((size_t)NVAR * 128 * 64);
    cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
    for (int v = 0; v < NVAR; ++v) {
        const Var V = vars[v];
        double maxerr = 0; int bad = 0;
        const int M = V.overlap ? 128 : 64;
        for (int m = 0; m < M; ++m) for (int n = 0; n < 64; ++n) {
            // accumulator row m: M = 64 sits in lane (m / 16) * 32 + m % 16, M = 128 in lane m
            const int lane = V.overlap ? m : (m / 16) * 32 + m % 16;
            double ref = 0;
            for (int k = 0; k < 128; ++k) {
                double qa, pb;
                if (V.overlap) { const int sh = V.off + (m >= 64 ? 1 : 0); qa = __bfloat162float(hQ[(sh + k) * 64 + (m & 63)]); pb = __bfloat162float(hP[k * 64 + n]); }
                else if (V.side == 0) { qa = __bfloat162float(hQ[(V.off + k) * 64 + m]); pb = __bfloat162float(hP[k * 64 + n]); }
                else { qa = __bfloat162float(hP[k * 64 + m]); pb = __bfloat162float(hQ[(V.off + k) * 64 + n]); }
                ref += qa * pb;
            }
            const double err = fabs(ref - hout[((size_t)v * 128 + lane) * 64 + n]);
            if (err > 1e-2) ++bad;
            if (err > maxerr) maxerr = err;
        }
        printf("%s row offset %3d  base_offset %d  shifted on %c : max err %.4f  bad %d  -> %s\n", V.overlap ? "M=128 overlapped (LBO=128B)" : "M=64",
               V.off, V.base, V.side ? 'B' : 'A', maxerr, bad, bad ? "WRONG" : "ok");
    }

    // ---- reduction throughput
    float *dst, *slab;
    const int per_thread = 160;      // 128 threads x 160 = 20480 floats per CTA (5 accumulators of 64 x 64)
    cudaMalloc(&dst, 64 << 20); cudaMalloc(&slab, (size_t)148 * per_thread * 128 * 4 * 2);
    cudaMemset(dst, 0, 64 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"red.f32, warp = 128 contiguous B", "red.v4.f32, lanes on rows 2304 B apart", "red.v4.f32, lanes on contiguous cells", "st.v4 to a private slab"};
    for (int ctas : {37, 74, 148})
        for (int mode = 0; mode < 4; ++mode) {
            red_probe<<<ctas, 128>>>(dst, slab, mode, per_thread);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            for (int r = 0; r < 20; ++r) red_probe<<<ctas, 128>>>(dst, slab, mode, per_thread);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double us = ms * 1000.0 / 20, floats = (double)ctas * 128 * per_thread;
            printf("reduce %-42s %3d CTAs x 20480 floats: %7.2f us  (%.1f floats/ns chip, %.2f floats/clk/SM at 1.9 GHz)\n", names[mode], ctas, us, floats / us / 1000.0,
                   floats / ctas / (us * 1900.0));
        }
    return 0;
}
