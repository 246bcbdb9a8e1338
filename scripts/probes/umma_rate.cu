// Probe: issue rate of tcgen05.mma (kind::f16, bf16 operands, M = 128 per CTA) with operands already resident, i.e. the pace
// the tensor pipe itself sustains, for
//   SS mode (A and B from shared memory, K-major SWIZZLE_128B), N = 64 / 128 / 256,
//   TS mode (A from tensor memory, B from shared memory),
//   cta_group::2 (M = 256 over a CTA pair, each CTA holds half of B),
//   one or two co-resident CTAs per SM, aligned and row-shifted A start addresses.
// One thread issues `iters` k-blocks of four K=16 MMAs that cycle over `stages` operand buffers, commits once and waits.
// Operand values are whatever shared memory holds (timing only).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I diffusion_model_universal_b200/csrc -o scripts/probes/umma_rate scripts/probes/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace dmu::tc;
namespace dmu { char* err_buf() { static char b[256]; return b; } int fail(const char* f, ...) { printf("fail: %s\n", f); return 1; } int sm_count() { return 148; } bool pdl_enabled() { return false; }
namespace tc { EncodeTiledFn encode_tiled_fn() { return nullptr; } } }

struct Args { int N, ts, iters, stages, rowshift; long long* out; int M, nacc, mnmajor, commit_every, loop_flags; };

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(128) rate1(const Args P) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    const int a_bytes = (128 + 8) * 128, b_bytes = P.N * 128;
    for (int i = threadIdx.x; i < P.stages * (a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(&s_tmem, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(P.M, P.N, P.mnmajor, P.mnmajor);
        const uint32_t base = smem_u32(smem);
        long long t0 = clock64();
        for (int it = 0; it < P.iters; ++it) {
            const int s = it % P.stages;
            const uint32_t dcol = tmem + (uint32_t)((it % P.nacc) * P.N);
            const uint32_t a_addr = base + s * (a_bytes + b_bytes) + P.rowshift * 128 * (it % 3);
            const uint64_t da = smem_desc_sw128(a_addr, P.mnmajor ? 8192 : 16, 1024);
            const uint64_t db = smem_desc_sw128(base + s * (a_bytes + b_bytes) + a_bytes, P.mnmajor ? 8192 : 16, 1024);
            // K-major: 32 B per K=16 step inside the 128-byte row; MN-major: 16 k-rows = 2 KB per step
            const uint32_t kstep = P.mnmajor ? 128 : 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (P.ts) umma_bf16_ts(dcol, tmem + 256 + (uint32_t)((s * 4 + k) * 8), db + kstep * k, idesc, (it | k) != 0);
                else umma_bf16(dcol, da + kstep * k, db + kstep * k, idesc, (it | k) != 0);
            }
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        P.out[2 * blockIdx.x] = t1 - t0;
        P.out[2 * blockIdx.x + 1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}


// Same work as rate1 in SS mode, but issued the CUTLASS way: warp 0 stays converged, one elected lane issues.
__global__ void __launch_bounds__(128) rate1e(const Args P) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar, bar2[2];
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    const int a_bytes = (128 + 8) * 128, b_bytes = P.N * 128;
    for (int i = threadIdx.x; i < P.stages * (a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2[0], 1); mbar_init(&bar2[1], 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(&s_tmem, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(P.M, P.N, 0, 0);
        const uint32_t base = smem_u32(smem);
        long long t0 = clock64();
        // tight loop: two stages, descriptors precomputed, nacc in {1, 2}
        const uint64_t da0 = smem_desc_sw128(base, 16, 1024), db0 = smem_desc_sw128(base + a_bytes, 16, 1024);
        const uint64_t da1 = smem_desc_sw128(base + (a_bytes + b_bytes), 16, 1024), db1 = smem_desc_sw128(base + (a_bytes + b_bytes) + a_bytes, 16, 1024);
        const uint32_t d1 = tmem + (uint32_t)((P.nacc - 1) * P.N);
        if (elect_one()) {
#pragma unroll 1
            for (int it = 0; it < P.iters; it += 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem, da0 + 2 * k, db0 + 2 * k, idesc, (it | k) != 0);
                if (P.commit_every == 4) umma_commit(&bar2[0]);      // what a per-k-block pipeline does: one commit per 4 MMAs
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d1, da1 + 2 * k, db1 + 2 * k, idesc, (it | k) != 0);
                if (P.commit_every == 4 || P.commit_every == 8) umma_commit(&bar2[1]);
            }
        }
        __syncwarp();
        long long t1 = clock64();
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { P.out[2 * blockIdx.x] = t1 - t0; P.out[2 * blockIdx.x + 1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}


// The per-k-block issue loop of the conv kernels, piece by piece: every iteration issues 4 MMAs; loop_flags adds
//   1 = an mbarrier try_wait on an already-complete barrier,  2 = tcgen05.fence::after_thread_sync,
//   4 = leaving / re-entering the elect.sync region (+ __syncwarp) every iteration,  8 = a tcgen05.commit every iteration.
__global__ void __launch_bounds__(128) rate1f(const Args P) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar, ready, sink[8];
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    const int a_bytes = (128 + 8) * 128, b_bytes = P.N * 128;
    for (int i = threadIdx.x; i < 2 * (a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&ready, 1); for (int i = 0; i < 8; ++i) mbar_init(&sink[i], 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(&s_tmem, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(P.M, P.N, 0, 0);
        const uint32_t base = smem_u32(smem);
        long long t0 = clock64();
        int st = 0;
        for (int it = 0; it < P.iters; ++it) {
            if (P.loop_flags & 1) mbar_wait(&ready, 1);           // fresh barrier: the phase before phase 0 counts as complete
            if (P.loop_flags & 2) tc_fence_after();
            const uint64_t da = smem_desc_sw128(base + st * (a_bytes + b_bytes), 16, 1024);
            const uint64_t db = smem_desc_sw128(base + st * (a_bytes + b_bytes) + a_bytes, 16, 1024);
            if (P.loop_flags & 4) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, (it | k) != 0);
                    if (P.loop_flags & 8) umma_commit(&sink[it & 7]);
                }
                __syncwarp();
            } else if (threadIdx.x == 0 || true) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, (it | k) != 0);
                    if (P.loop_flags & 8) umma_commit(&sink[it & 7]);
                }
            }
            st ^= 1;
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { P.out[2 * blockIdx.x] = t1 - t0; P.out[2 * blockIdx.x + 1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---- cta_group::2: the pair computes D[256 x N]; each CTA holds its 128 rows of A and N/2 rows of B
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) rate2(const Args P) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_rank();
    const int a_bytes = 128 * 128, b_bytes = (P.N / 2) * 128;
    for (int i = threadIdx.x; i < P.stages * (a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); cluster_sync(); tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (rank == 0 && threadIdx.x == 0) {
        // M = 256 (bits 24-28 = M>>4), N
        const uint32_t idesc = umma_idesc_bf16(256, P.N, 0, 0);
        const uint32_t base = smem_u32(smem);
        long long t0 = clock64();
        for (int it = 0; it < P.iters; ++it) {
            const int s = it % P.stages;
            const uint64_t da = smem_desc_sw128(base + s * (a_bytes + b_bytes), 16, 1024);
            const uint64_t db = smem_desc_sw128(base + s * (a_bytes + b_bytes) + a_bytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t acc = (it | k) != 0;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc)
                    : "memory");
            }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        P.out[2 * (blockIdx.x / 2)] = t1 - t0;
        P.out[2 * (blockIdx.x / 2) + 1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads(); cluster_sync();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static void run(const char* what, int pair, int N, int ts, int stages, int rowshift, int ctas, int smem_pad, int M = 128, int nacc = 1, int mnmajor = 0, int elect = 0, int commit_every = 0, int loop_flags = -1) {
    const int iters = 2048;
    long long* dout;
    cudaMalloc(&dout, sizeof(long long) * 2 * 1024);
    cudaMemset(dout, 0, sizeof(long long) * 2 * 1024);
    Args A{N, ts, iters, stages, rowshift, dout, M, nacc, mnmajor, commit_every, loop_flags};
    const int per_stage = pair ? (128 * 128 + N / 2 * 128) : ((128 + 8) * 128 + N * 128);
    int smem = stages * per_stage + 1024 + smem_pad;
    if (pair) { cudaFuncSetAttribute(rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); rate2<<<ctas, 128, smem>>>(A); }
    else if (loop_flags >= 0) { cudaFuncSetAttribute(rate1f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); rate1f<<<ctas, 128, smem>>>(A); }
    else if (elect) { cudaFuncSetAttribute(rate1e, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); rate1e<<<ctas, 128, smem>>>(A); }
    else { cudaFuncSetAttribute(rate1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); rate1<<<ctas, 128, smem>>>(A); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s FAILED: %s\n", what, cudaGetErrorString(e)); exit(1); }
    const int nres = pair ? ctas / 2 : ctas;
    std::vector<long long> h(2 * nres);
    cudaMemcpy(h.data(), dout, sizeof(long long) * 2 * nres, cudaMemcpyDeviceToHost);
    std::vector<double> per;
    double issue = 0;
    for (int i = 0; i < nres; ++i) { per.push_back((double)h[2 * i + 1] / (iters * 4)); issue += (double)h[2 * i] / (iters * 4); }
    std::sort(per.begin(), per.end());
    const double floor_clk = (pair ? 256.0 : 128.0) * N / (256.0 * (pair ? 2 : 1));
    printf("%-44s ctas %4d  clk/MMA median %7.1f max %7.1f (issue %5.1f)  floor %5.1f  => %5.1f %% of tensor peak\n", what, ctas, per[per.size() / 2],
           per.back(), issue / nres, floor_clk, 100.0 * floor_clk / per[per.size() / 2]);
    cudaFree(dout);
}

int main() {
    // smem_pad > 114 KB forces one CTA per SM; small smem lets two co-reside
    const int ONE = 120 * 1024;
    for (int N : {64, 128, 256}) {
        char b[96];
        snprintf(b, sizeof b, "SS  N=%d 1 CTA/SM, 1 CTA only", N); run(b, 0, N, 0, 2, 0, 1, ONE - 2 * ((128 + 8) * 128 + N * 128));
        snprintf(b, sizeof b, "SS  N=%d 1 CTA/SM, 148 CTAs", N); run(b, 0, N, 0, 2, 0, 148, ONE - 2 * ((128 + 8) * 128 + N * 128));
        snprintf(b, sizeof b, "SS  N=%d 2 CTA/SM, 296 CTAs", N); run(b, 0, N, 0, 2, 0, 296, 0);
        snprintf(b, sizeof b, "SS  N=%d row-shifted A, 148 CTAs", N); run(b, 0, N, 0, 2, 1, 148, ONE - 2 * ((128 + 8) * 128 + N * 128));
        snprintf(b, sizeof b, "TS  N=%d 1 CTA/SM, 148 CTAs", N); run(b, 0, N, 1, 2, 0, 148, ONE - 2 * ((128 + 8) * 128 + N * 128));
        snprintf(b, sizeof b, "SS  cta_group::2 M=256 N=%d, 74 pairs", N); run(b, 1, N, 0, 2, 0, 148, ONE - 2 * (128 * 128 + N / 2 * 128));
    }
    const int pad64 = ONE - 2 * ((128 + 8) * 128 + 64 * 128), pad128 = ONE - 2 * ((128 + 8) * 128 + 128 * 128), pad256 = ONE - 2 * ((128 + 8) * 128 + 256 * 128);
    run("SS elect-issued N=16", 0, 16, 0, 2, 0, 148, pad64, 128, 1, 0, 1);
    run("SS elect-issued N=64", 0, 64, 0, 2, 0, 148, pad64, 128, 1, 0, 1);
    run("SS elect-issued N=128", 0, 128, 0, 2, 0, 148, pad128, 128, 1, 0, 1);
    run("SS elect-issued N=256", 0, 256, 0, 2, 0, 148, pad256, 128, 1, 0, 1);
    run("SS elect-issued N=64 2 acc", 0, 64, 0, 2, 0, 148, pad64, 128, 2, 0, 1);
    run("SS elect-issued N=128 2 acc", 0, 128, 0, 2, 0, 148, pad128, 128, 2, 0, 1);
    run("SS elect-issued N=32", 0, 32, 0, 2, 0, 148, pad64, 128, 1, 0, 1);
    run("SS elect-issued N=96", 0, 96, 0, 2, 0, 148, pad128, 128, 1, 0, 1);
    run("SS elect-issued N=192", 0, 192, 0, 2, 0, 148, pad256, 128, 1, 0, 1);
    run("SS elect-issued M=64 N=64", 0, 64, 0, 2, 0, 148, pad64, 64, 1, 0, 1);
    run("SS elect-issued M=64 N=128", 0, 128, 0, 2, 0, 148, pad128, 64, 1, 0, 1);
    run("SS elect-issued M=64 N=256", 0, 256, 0, 2, 0, 148, pad256, 64, 1, 0, 1);
    run("SS elect-issued N=64, 296 CTAs (2/SM)", 0, 64, 0, 2, 0, 296, 0, 128, 1, 0, 1);
    run("SS elect-issued N=128, 296 CTAs (2/SM)", 0, 128, 0, 2, 0, 296, 0, 128, 1, 0, 1);
    run("SS elect N=64, commit every 4 MMAs", 0, 64, 0, 2, 0, 148, pad64, 128, 1, 0, 1, 4);
    run("SS elect N=64, commit every 8 MMAs", 0, 64, 0, 2, 0, 148, pad64, 128, 1, 0, 1, 8);
    run("SS elect N=128, commit every 4 MMAs", 0, 128, 0, 2, 0, 148, pad128, 128, 1, 0, 1, 4);
    run("SS elect N=128, commit every 8 MMAs", 0, 128, 0, 2, 0, 148, pad128, 128, 1, 0, 1, 8);
    for (int f : {0, 1, 2, 4, 8, 12, 15}) {
        char b[96];
        snprintf(b, sizeof b, "k-block loop N=64 flags=%d", f);
        run(b, 0, 64, 0, 2, 0, 148, pad64, 128, 1, 0, 1, 0, f);
    }
    run("SS N=16", 0, 16, 0, 2, 0, 148, pad64);
    run("SS N=32", 0, 32, 0, 2, 0, 148, pad64);
    run("SS N=64  2 accumulators", 0, 64, 0, 2, 0, 148, pad64, 128, 2);
    run("SS N=64  4 accumulators", 0, 64, 0, 2, 0, 148, pad64, 128, 4);
    run("SS N=128 2 accumulators", 0, 128, 0, 2, 0, 148, pad128, 128, 2);
    run("TS N=64  2 accumulators", 0, 64, 1, 2, 0, 148, pad64, 128, 2);
    run("TS N=128 2 accumulators", 0, 128, 1, 2, 0, 148, pad128, 128, 2);
    run("SS M=64 N=64", 0, 64, 0, 2, 0, 148, pad64, 64);
    run("SS M=64 N=128", 0, 128, 0, 2, 0, 148, pad128, 64);
    run("SS M=64 N=256", 0, 256, 0, 2, 0, 148, pad256, 64);
    run("SS M=64 N=256 2 accumulators", 0, 256, 0, 2, 0, 148, pad256, 64, 2);
    run("SS MN-major N=64", 0, 64, 0, 2, 0, 148, pad64, 128, 1, 1);
    run("SS MN-major N=128", 0, 128, 0, 2, 0, 148, pad128, 128, 1, 1);
    run("SS 1 stage N=64 (same operands every time)", 0, 64, 0, 1, 0, 148, pad64);
    run("SS 1 stage N=128", 0, 128, 0, 1, 0, 148, pad128);
    return 0;
}
