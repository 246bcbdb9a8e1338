// Shared helpers for the sm_100a kernels behind include/dmu_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <utility>

#include "../../include/dmu_b200.h"

namespace dmu {

// thread-local error text returned by dmu_last_error()
char* err_buf();
int fail(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("%s: %s", what, cudaGetErrorString(e));
    return 0;
}

#define DMU_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) return dmu::fail(__VA_ARGS__); \
    } while (0)

inline cudaStream_t as_stream(dmu_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

// ---- dtype helpers -------------------------------------------------------
__device__ __forceinline__ float ld_as_float(const void* base, int64_t idx, int dtype) {
    if (dtype == DMU_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
    return reinterpret_cast<const float*>(base)[idx];
}
__device__ __forceinline__ void st_from_float(void* base, int64_t idx, int dtype, float v) {
    if (dtype == DMU_BF16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(base)[idx] = v;
}

template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int kVec = 4;  // elements per 16 B
    static constexpr int kCode = DMU_F32;
    __device__ static __forceinline__ float to_f(float v) { return v; }
    __device__ static __forceinline__ float from_f(float v) { return v; }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kVec = 8;
    static constexpr int kCode = DMU_BF16;
    __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 16-byte vector of T <-> float[kVec]
template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float* out) {
    if constexpr (sizeof(T) == 4) {
        float4 v = *reinterpret_cast<const float4*>(p);
        out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            out[2 * i] = f.x; out[2 * i + 1] = f.y;
        }
    }
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float* in) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
    } else {
        uint4 v;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = v;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

inline bool is_nhwc(const dmu_tensor4& t) { return t.sc == 1; }

// DMU_GN_FIXED_SUMS (include/dmu_b200.h): GroupNorm raw sums accumulated as 64-bit fixed-point integers (value * 2^20) in the int64
// array behind the float array - integer atomics commute, the statistics are bit-identical from run to run.
__device__ __forceinline__ unsigned long long* gn_fixed_sums(float* sums, int N, int G) {
    return reinterpret_cast<unsigned long long*>(sums + (int64_t)N * G * 2);
}
__device__ __forceinline__ unsigned long long gn_to_fixed(float v) { return (unsigned long long)__float2ll_rn(v * 1048576.f); }
__device__ __forceinline__ float gn_from_fixed(unsigned long long v) { return (float)((double)(long long)v * (1.0 / 1048576.0)); }


// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// A recorded plan is a chain of hundreds of short dependent kernels; with the launch attribute below the next kernel's CTAs
// are scheduled (and run their prologue) while the previous kernel drains, and block in pdl_wait() until that kernel's
// memory is visible.  Contract: EVERY kernel launched through launch_pdl() executes pdl_wait() in every CTA before it
// touches global memory produced by an earlier launch (a kernel that skipped it could complete before its predecessor
// and release ITS successor too early).  Both intrinsics are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();   // DMU_PDL=0 in the environment turns the attribute off

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, dim3 cluster, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster.x * cluster.y * cluster.z > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster.x;
        attr[n].val.clusterDim.y = cluster.y;
        attr[n].val.clusterDim.z = cluster.z;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace dmu
