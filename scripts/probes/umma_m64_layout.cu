// Probe: where do the 64 rows of an M = 64 (cta_group::1) tcgen05.mma accumulator live in tensor memory?
// (derived from umma_rowoffset.cu) -- original header: does a K-major SWIZZLE_128B UMMA A-operand descriptor work when its start address is offset by a multiple of
// 128 bytes (one tile row) that is NOT a multiple of the 1024-byte swizzle period?  (Needed to read the nine 3x3 taps of a
// convolution as shifted windows of ONE shared-memory halo tile.)  Variants: descriptor base_offset field = 0 or (row & 7).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I diffusion_model_universal_b200/csrc -o /tmp/probe scripts/probes/umma_rowoffset.cu
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

constexpr int ROWS = 256, NVAR = 24;
struct Maps2 { CUtensorMap a, b; };

__global__ void __launch_bounds__(128) probe(const __grid_constant__ Maps2 maps, float* out, const int* roffs, const int* bases) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                 // 256 rows x 128 B
    uint8_t* sB = smem + ROWS * 128;    // 64 rows x 128 B
    __shared__ __align__(8) uint64_t bar, mbar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&mbar, 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(&s_tmem, 64);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, ROWS * 128 + 64 * 128);
        tma_load_2d(sA, &maps.a, &bar, 0, 0);
        tma_load_2d(sB, &maps.b, &bar, 0, 0);
    }
    mbar_wait(&bar, 0);
    for (int v = 0; v < NVAR; ++v) {
        if (threadIdx.x == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(64, 64, 0, 0);
            uint64_t da = smem_desc_sw128(smem_u32(sA) + roffs[v] * 128, 16, 1024) | ((uint64_t)(bases[v] & 7) << 49);
            uint64_t db = smem_desc_sw128(smem_u32(sB), 16, 1024);
            for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
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

int main() {
    std::vector<__nv_bfloat16> hA(ROWS * 64), hB(64 * 64);
    for (int r = 0; r < ROWS; ++r) for (int k = 0; k < 64; ++k) hA[r * 64 + k] = __float2bfloat16((float)(((r * 7 + k * 3) % 17) - 8) / 8.f);
    for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16((float)(((n * 5 + k * 11) % 13) - 6) / 4.f);
    __nv_bfloat16 *dA, *dB; float* dout; int *droff, *dbase;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dout, NVAR * 128 * 64 * 4); cudaMalloc(&droff, NVAR * 4); cudaMalloc(&dbase, NVAR * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    int roffs[NVAR], bases[NVAR];
    const int offs[12] = {0, 1, 2, 3, 5, 7, 8, 9, 34, 35, 68, 127};
    for (int i = 0; i < 12; ++i) { roffs[2 * i] = offs[i]; bases[2 * i] = 0; roffs[2 * i + 1] = offs[i]; bases[2 * i + 1] = offs[i] & 7; }
    cudaMemcpy(droff, roffs, sizeof(roffs), cudaMemcpyHostToDevice); cudaMemcpy(dbase, bases, sizeof(bases), cudaMemcpyHostToDevice);
    Maps2 maps;
    { uint64_t dims[2] = {64, ROWS}, str[2] = {1, 64}; uint32_t box[2] = {64, ROWS};
      // box rows limited to 256
      cuuint64_t gd[2] = {dims[0], dims[1]}, gs[1] = {str[1] * 2}; cuuint32_t bd[2] = {box[0], box[1]}, es[2] = {1, 1};
      CUresult r = encode_tiled_fn()(&maps.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode a failed %d\n", (int)r); return 1; } }
    { cuuint64_t gd[2] = {64, 64}, gs[1] = {128}; cuuint32_t bd[2] = {64, 64}, es[2] = {1, 1};
      CUresult r = encode_tiled_fn()(&maps.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode b failed %d\n", (int)r); return 1; } }
    const int smem = ROWS * 128 + 64 * 128 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(maps, dout, droff, dbase);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hout(NVAR * 128 * 64);
    cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
    // variant 0 (row offset 0): match every TMEM lane against every A row
    std::vector<std::vector<double>> ref(64, std::vector<double>(64));
    for (int r = 0; r < 64; ++r) for (int n = 0; n < 64; ++n) {
        double a = 0; for (int k = 0; k < 64; ++k) a += (double)__bfloat162float(hA[r * 64 + k]) * (double)__bfloat162float(hB[n * 64 + k]);
        ref[r][n] = a;
    }
    for (int lane = 0; lane < 128; ++lane) {
        int match = -1;
        for (int r = 0; r < 64 && match < 0; ++r) {
            bool ok = true;
            for (int n = 0; n < 64 && ok; ++n) ok = fabs(ref[r][n] - hout[((size_t)0 * 128 + lane) * 64 + n]) < 1e-3;
            if (ok) match = r;
        }
        printf("lane %3d -> A row %d\n", lane, match);
    }
    return 0;
}
