// Probe: what does one kernel -> kernel dependency cost inside a CUDA graph on B200, for the three ways a chain of short
// dependent kernels can be linked?
//   mode 0  plain stream order (full dependency: the next grid is launched after the previous one has completed)
//   mode 1  programmatic dependent launch: next grid resident early, blocks in griddepcontrol.wait until the previous grid
//           has completed and flushed
//   mode 2  programmatic dependent launch WITHOUT griddepcontrol.wait: the previous grid's CTAs publish "my stores are done"
//           with fence + red.release on a per-launch counter, the next grid's CTAs spin (ld.acquire) on that counter
// Every kernel of the chain reads what ANOTHER CTA of the previous kernel wrote (a wrong hand-off shows up in the checksum),
// does `work` dependent FMAs per thread and writes its output.  The chain is captured into a graph and replayed.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probes/handoff scripts/probes/handoff.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(unsigned* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

__global__ void __launch_bounds__(128) link(const float* __restrict__ in, float* __restrict__ out, unsigned* flags, int idx, int mode, int work) {
    if (mode >= 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ float s[128];
    s[threadIdx.x] = (float)threadIdx.x;      // some prologue
    __syncthreads();
    if (mode == 1) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (mode == 2 && idx > 0) {
        if (threadIdx.x == 0) {
            while (ld_acquire(flags + idx - 1) < gridDim.x) {}
        }
        __syncthreads();
    }
    const int src = ((blockIdx.x + 1) % gridDim.x) * 128 + threadIdx.x;
    float v = __ldcg(in + src);
    for (int i = 0; i < work; ++i) v = fmaf(v, 1.0000001f, 1e-3f);
    out[blockIdx.x * 128 + threadIdx.x] = v + 1.f;
    if (mode == 2) {
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); red_release(flags + idx); }
    }
}

static float run(int mode, int grid, int work, int chain, double* checksum) {
    float *a, *b; unsigned* flags;
    CK(cudaMalloc(&a, 148 * 8 * 128 * 4)); CK(cudaMalloc(&b, 148 * 8 * 128 * 4)); CK(cudaMalloc(&flags, chain * 4));
    CK(cudaMemset(a, 0, 148 * 8 * 128 * 4)); CK(cudaMemset(b, 0, 148 * 8 * 128 * 4));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    CK(cudaMemsetAsync(flags, 0, chain * 4, st));
    CK(cudaMemsetAsync(a, 0, grid * 128 * 4, st));
    for (int i = 0; i < chain; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = (mode >= 1 && i > 0) ? 1 : 0;
        const float* in = (i & 1) ? b : a; float* out = (i & 1) ? a : b;
        CK(cudaLaunchKernelEx(&cfg, link, in, out, flags, i, mode, work));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 20;
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<float> h(grid * 128);
    CK(cudaMemcpy(h.data(), (chain & 1) ? b : a, grid * 128 * 4, cudaMemcpyDeviceToHost));
    double s = 0; for (float v : h) s += v;
    *checksum = s;
    cudaFree(a); cudaFree(b); cudaFree(flags); cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
    return ms * 1e3f / reps / chain;
}

int main() {
    const int chain = 200;
    printf("chain of %d dependent kernels in one graph; us per kernel\n", chain);
    printf("%6s %6s | %10s %10s %10s | checksums\n", "ctas", "work", "stream", "pdl-wait", "pdl-flags");
    for (int grid : {8, 32, 128, 148, 296}) {
        for (int work : {0, 2000, 8000}) {
            double c0, c1, c2;
            float t0 = run(0, grid, work, chain, &c0), t1 = run(1, grid, work, chain, &c1), t2 = run(2, grid, work, chain, &c2);
            printf("%6d %6d | %10.2f %10.2f %10.2f | %.6g %.6g %.6g %s\n", grid, work, t0, t1, t2, c0, c1, c2, (c0 == c1 && c1 == c2) ? "ok" : "MISMATCH");
        }
    }
    return 0;
}
