// GroupNorm(+SiLU) forward/backward, channel sums, activations, sinusoidal
// embedding, parameter repack and the fused Adam+EMA update.
// All HBM-bound: 128-bit accesses on NHWC rows, warp-shuffle / smem reductions,
// fp32 statistics.  Call sites: see include/dmu_b200.h.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace dmu {

constexpr int kMaxC = 1024;  // channels per pixel supported by the smem-staged kernels

// one CTA handles pixels [p0, p1) of image n; thread -> (vector v of kVec channels, pixel lane)
struct RowMap {
    int V, lanes, v, lane; bool active;
    __device__ RowMap(int C, int kVec) {
        V = C / kVec;
        lanes = blockDim.x / V;
        v = threadIdx.x % V;
        lane = threadIdx.x / V;
        active = lane < lanes;
    }
};

__device__ __forceinline__ void chunk_range(int HW, int& p0, int& p1) {
    const int per = (HW + gridDim.x - 1) / gridDim.x;
    p0 = blockIdx.x * per;
    p1 = min(HW, p0 + per);
}

// Every activation of the plans is pixel-contiguous (row pitch = W x pixel pitch, channel slices of the concat buffers
// included): skip the divide / modulo, which made these kernels instruction-bound (~100 integer instructions per access).
__device__ __forceinline__ int64_t pix_off(const dmu_tensor4& t, int n, int p, int W) {
    if (t.sh == (int64_t)W * t.sw) return (int64_t)n * t.sn + (int64_t)p * t.sw;
    return (int64_t)n * t.sn + (int64_t)(p / W) * t.sh + (int64_t)(p % W) * t.sw;
}

// 16-byte raw vectors: loads are issued kUnroll at a time before any arithmetic so that each thread keeps several
// independent 128-bit requests in flight (these passes are HBM-bound; one dependent load per iteration is latency-bound).
template <typename T>
__device__ __forceinline__ uint4 ld_raw(const T* p) { return *reinterpret_cast<const uint4*>(p); }
template <typename T>
__device__ __forceinline__ void unpack(const uint4& r, float* out) {
    if constexpr (sizeof(T) == 4) {
        out[0] = __uint_as_float(r.x); out[1] = __uint_as_float(r.y); out[2] = __uint_as_float(r.z); out[3] = __uint_as_float(r.w);
    } else {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            out[2 * i] = f.x; out[2 * i + 1] = f.y;
        }
    }
}
constexpr int kUnroll = 4;

// 16-byte asynchronous global -> shared copies (LDGSTS): the slab kernels below put their WHOLE slab in flight at once without
// holding it in registers, and consume it group by group in issue order.  A thread only ever reads back the bytes it copied
// itself, so cp.async.wait_group is all the synchronisation the data needs.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// wait until at most `pending` of this thread's committed groups are still in flight (pending < 8)
__device__ __forceinline__ void cp_async_wait_dyn(int pending) {
    switch (pending) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        case 6: cp_async_wait<6>(); break;
        default: cp_async_wait<7>(); break;
    }
}
constexpr int kSlabGroups = 8;      // commit groups per slab: consumption starts when the first eighth has landed

__device__ __forceinline__ float silu_f(float u) { return __fdividef(u, 1.f + __expf(-u)); }

// du = dy * act'(u)
__device__ __forceinline__ float act_grad(float u, float dy, int silu) {
    if (!silu) return dy;
    const float s = __fdividef(1.f, 1.f + __expf(-u));
    return dy * (s * fmaf(u, 1.f - s, 1.f));
}

// One-MUFU sigmoid for the bf16 slab kernels: sigmoid(u) = 0.5 tanh(u / 2) + 0.5 (tanh.approx.f32: ~2^-11 relative, below the
// 2^-9 of the bf16 store that follows).  These kernels are instruction-bound (profiles/r01_ncu_full_groupnorm.txt), not HBM-bound:
// the exp2 + reciprocal form costs two MUFU and three more FP instructions per element.
__device__ __forceinline__ float sigmoid_fast(float u) {
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * u));
    return fmaf(0.5f, th, 0.5f);
}
template <typename T>
__device__ __forceinline__ float act_fwd_t(float u, int silu) {
    if constexpr (sizeof(T) == 2) return silu ? u * sigmoid_fast(u) : u;
    else return silu ? silu_f(u) : u;
}
template <typename T>
__device__ __forceinline__ float act_grad_t(float u, float dy, int silu) {
    if constexpr (sizeof(T) == 2) {
        const float sg = sigmoid_fast(u);
        const float r = dy * (sg * fmaf(u, 1.f - sg, 1.f));
        return silu ? r : dy;
    } else {
        return act_grad(u, dy, silu);
    }
}



// Block-wide per-channel sum of per-thread partials WITHOUT shared-memory float atomics (those compile to a CAS spin loop
// and serialise under same-address contention).  Every thread parks its kVec partials in s_red[lane][channel]
// (lanes * C == 256 * kVec floats, whatever C is), then thread c adds the `lanes` rows of column c.
template <int kVec>
__device__ __forceinline__ void block_channel_sum(const RowMap& m, const float* a, float* s_red, float* s_out, int C) {
    if (m.active) {
#pragma unroll
        for (int i = 0; i < kVec; ++i) s_red[m.lane * C + m.v * kVec + i] = a[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int l = 0; l < m.lanes; ++l) t += s_red[l * C + c];
        s_out[c] = t;
    }
    __syncthreads();
}

// ------------------------------------------------------------------ stats
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(dmu_gn_params P) {
    pdl_trigger();
    pdl_wait();
    constexpr int kVec = Elem<T>::kVec;
    __shared__ float s_sum[kMaxC], s_sq[kMaxC];
    __shared__ float s_red[256 * kVec];
    const int n = blockIdx.y, HW = P.H * P.W, C = P.C;
    RowMap m(C, kVec);
    int p0, p1; chunk_range(HW, p0, p1);
    float a[kVec], q[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) { a[i] = 0.f; q[i] = 0.f; }
    if (m.active) {
        const T* base = reinterpret_cast<const T*>(P.x.ptr) + m.v * kVec;
        for (int p = p0 + m.lane; p < p1; p += m.lanes * kUnroll) {
            uint4 r[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int pp = p + u * m.lanes;
                r[u] = pp < p1 ? ld_raw<T>(base + pix_off(P.x, n, pp, P.W)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                float v[kVec];
                unpack<T>(r[u], v);
#pragma unroll
                for (int i = 0; i < kVec; ++i) { a[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
            }
        }
    }
    block_channel_sum<kVec>(m, a, s_red, s_sum, C);
    block_channel_sum<kVec>(m, q, s_red, s_sq, C);
    const int cpg = C / P.G;
    for (int g = threadIdx.x; g < P.G; g += blockDim.x) {
        float a = 0.f, q = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += s_sum[c]; q += s_sq[c]; }
        if (P.flags & DMU_GN_FIXED_SUMS) {
            unsigned long long* fx = gn_fixed_sums(P.sums, P.N, P.G) + ((int64_t)n * P.G + g) * 2;
            atomicAdd(fx, gn_to_fixed(a));
            atomicAdd(fx + 1, gn_to_fixed(q));
        } else {
            atomicAdd(&P.sums[((int64_t)n * P.G + g) * 2 + 0], a);
            atomicAdd(&P.sums[((int64_t)n * P.G + g) * 2 + 1], q);
        }
    }
}

// per-channel mean / rstd*gamma / beta staged in smem for image n
// fixed: read the statistics from the int64 fixed-point accumulators behind P.sums (DMU_GN_FIXED_SUMS; only the apply pass that
// directly follows the accumulating launch does - it also publishes their float value in P.sums, see gn_apply_kernel)
__device__ __forceinline__ void stage_affine(const dmu_gn_params& P, int n, float* s_mean, float* s_scale, float* s_beta, float* s_rstd,
                                             float* s_gamma = nullptr, bool fixed = false) {
    const int cpg = P.C / P.G;
    const float cnt = (float)cpg * (float)P.H * (float)P.W;
    for (int c = threadIdx.x; c < P.C; c += blockDim.x) {
        const int g = c / cpg;
        float su, sq;
        if (fixed) {
            const unsigned long long* fx = gn_fixed_sums(P.sums, P.N, P.G) + ((int64_t)n * P.G + g) * 2;
            su = gn_from_fixed(fx[0]); sq = gn_from_fixed(fx[1]);
        } else {
            su = P.sums[((int64_t)n * P.G + g) * 2 + 0]; sq = P.sums[((int64_t)n * P.G + g) * 2 + 1];
        }
        const float mean = su / cnt;
        const float var = fmaxf(sq / cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + P.eps);
        const float gam = P.gamma[c];
        s_mean[c] = mean;
        s_scale[c] = rstd * gam;
        s_beta[c] = P.beta[c];
        if (s_rstd) s_rstd[c] = rstd;
        if (s_gamma) s_gamma[c] = gam;
    }
}

// ------------------------------------------------------------------ apply
template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(dmu_gn_params P) {
    pdl_trigger();
    pdl_wait();
    constexpr int kVec = Elem<T>::kVec;
    __shared__ float s_mean[kMaxC], s_scale[kMaxC], s_beta[kMaxC];
    const int n = blockIdx.y, HW = P.H * P.W;
    const bool fixed = (P.flags & DMU_GN_FIXED_SUMS) != 0;
    stage_affine(P, n, s_mean, s_scale, s_beta, nullptr, nullptr, fixed);
    if (fixed && blockIdx.x == 0) {      // the float view of the order-independent sums, for the backward of this norm
        const unsigned long long* fx = gn_fixed_sums(P.sums, P.N, P.G) + (int64_t)n * P.G * 2;
        for (int i = threadIdx.x; i < P.G * 2; i += blockDim.x) P.sums[(int64_t)n * P.G * 2 + i] = gn_from_fixed(fx[i]);
    }
    __syncthreads();
    RowMap m(P.C, kVec);
    if (!m.active) return;
    int p0, p1; chunk_range(HW, p0, p1);
    const T* xb = reinterpret_cast<const T*>(P.x.ptr) + m.v * kVec;
    T* yb = reinterpret_cast<T*>(P.y.ptr) + m.v * kVec;
    float sc[kVec], sh[kVec];   // u = x*sc + sh
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        sc[i] = s_scale[m.v * kVec + i];
        sh[i] = s_beta[m.v * kVec + i] - s_mean[m.v * kVec + i] * sc[i];
    }
    for (int p = p0 + m.lane; p < p1; p += m.lanes * kUnroll) {
        uint4 r[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int pp = p + u * m.lanes;
            if (pp < p1) r[u] = ld_raw<T>(xb + pix_off(P.x, n, pp, P.W));
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int pp = p + u * m.lanes;
            if (pp >= p1) break;
            float v[kVec];
            unpack<T>(r[u], v);
#pragma unroll
            for (int i = 0; i < kVec; ++i) {
                v[i] = act_fwd_t<T>(fmaf(v[i], sc[i], sh[i]), P.silu);
            }
            store_vec<T>(yb + pix_off(P.y, n, pp, P.W), v);
        }
    }
}


// per-image affine form of the normalisation for the conv kernels that apply it to their operand tile (conv_halo.cu)
__global__ void __launch_bounds__(256) gn_coef_kernel(dmu_gn_params P, float* __restrict__ coef) {
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.x;
    const int cpg = P.C / P.G;
    const float cnt = (float)cpg * (float)P.H * (float)P.W;
    for (int c = threadIdx.x; c < P.C; c += blockDim.x) {
        const int g = c / cpg;
        const float su = P.sums[((int64_t)n * P.G + g) * 2 + 0], sq = P.sums[((int64_t)n * P.G + g) * 2 + 1];
        const float mean = su / cnt;
        const float rstd = rsqrtf(fmaxf(sq / cnt - mean * mean, 0.f) + P.eps);
        const float sc = rstd * P.gamma[c];
        coef[((int64_t)n * P.C + c) * 2 + 0] = sc;
        coef[((int64_t)n * P.C + c) * 2 + 1] = P.beta[c] - mean * sc;
    }
}

// ------------------------------------------------------------------ bwd reduce
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(dmu_gn_params P) {
    pdl_trigger();
    pdl_wait();
    constexpr int kVec = Elem<T>::kVec;
    constexpr int kU = 2;
    __shared__ float s_mean[kMaxC], s_scale[kMaxC], s_beta[kMaxC], s_rstd[kMaxC];
    __shared__ float s_a[kMaxC], s_b[kMaxC];
    __shared__ float s_red[256 * kVec];
    const int n = blockIdx.y, HW = P.H * P.W, C = P.C;
    stage_affine(P, n, s_mean, s_scale, s_beta, s_rstd);
    __syncthreads();
    RowMap m(C, kVec);
    int p0, p1; chunk_range(HW, p0, p1);
    float a[kVec], b[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) { a[i] = 0.f; b[i] = 0.f; }
    if (m.active) {
        float mu[kVec], sc[kVec], be[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            mu[i] = s_mean[m.v * kVec + i]; sc[i] = s_scale[m.v * kVec + i]; be[i] = s_beta[m.v * kVec + i];
        }
        const T* xb = reinterpret_cast<const T*>(P.x.ptr) + m.v * kVec;
        const T* dyb = reinterpret_cast<const T*>(P.y.ptr) + m.v * kVec;
        for (int p = p0 + m.lane; p < p1; p += m.lanes * kU) {
            uint4 rx[kU], rd[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int pp = p + u * m.lanes;
                const bool ok = pp < p1;
                rx[u] = ok ? ld_raw<T>(xb + pix_off(P.x, n, pp, P.W)) : make_uint4(0u, 0u, 0u, 0u);
                rd[u] = ok ? ld_raw<T>(dyb + pix_off(P.y, n, pp, P.W)) : make_uint4(0u, 0u, 0u, 0u);   // dy = 0 -> no contribution
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                float xv[kVec], dv[kVec];
                unpack<T>(rx[u], xv);
                unpack<T>(rd[u], dv);
#pragma unroll
                for (int i = 0; i < kVec; ++i) {
                    const float d = xv[i] - mu[i];
                    const float du = act_grad_t<T>(fmaf(d, sc[i], be[i]), dv[i], P.silu);
                    a[i] += du;
                    b[i] = fmaf(du, d, b[i]);     // sum du*(x-mean); scaled by rstd below
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kVec; ++i) b[i] *= s_rstd[m.v * kVec + i];
    }
    block_channel_sum<kVec>(m, a, s_red, s_a, C);
    block_channel_sum<kVec>(m, b, s_red, s_b, C);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(&P.red[((int64_t)n * C + c) * 2 + 0], s_a[c]);
        atomicAdd(&P.red[((int64_t)n * C + c) * 2 + 1], s_b[c]);
        if (P.dbeta) atomicAdd(&P.dbeta[c], s_a[c]);
        if (P.dgamma) atomicAdd(&P.dgamma[c], s_b[c]);
    }
}

// ------------------------------------------------------------------ bwd apply
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(dmu_gn_params P) {
    pdl_trigger();
    pdl_wait();
    constexpr int kVec = Elem<T>::kVec;
    constexpr int kU = 2;
    __shared__ float s_mean[kMaxC], s_scale[kMaxC], s_beta[kMaxC], s_rstd[kMaxC];
    __shared__ float s_A[kMaxC], s_B[kMaxC];  // per channel: group sums / cnt
    const int n = blockIdx.y, HW = P.H * P.W, C = P.C;
    stage_affine(P, n, s_mean, s_scale, s_beta, s_rstd);
    const int cpg = C / P.G;
    const float inv_cnt = 1.f / ((float)cpg * (float)HW);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g0 = (c / cpg) * cpg;
        float A = 0.f, B = 0.f;
        for (int k = g0; k < g0 + cpg; ++k) {
            const float gam = P.gamma[k];
            A += gam * P.red[((int64_t)n * C + k) * 2 + 0];
            B += gam * P.red[((int64_t)n * C + k) * 2 + 1];
        }
        s_A[c] = A * inv_cnt;
        s_B[c] = B * inv_cnt;
    }
    __syncthreads();
    RowMap m(C, kVec);
    if (!m.active) return;
    int p0, p1; chunk_range(HW, p0, p1);
    // dx = du*sc - rstd*A - (x-mean)*rstd^2*B  =  du*sc + x*k1 + k0
    float mu[kVec], sc[kVec], be[kVec], k0[kVec], k1[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const int c = m.v * kVec + i;
        const float rs = s_rstd[c];
        mu[i] = s_mean[c]; sc[i] = s_scale[c]; be[i] = s_beta[c];
        k1[i] = -rs * rs * s_B[c];
        k0[i] = -rs * s_A[c] - mu[i] * k1[i];
    }
    const T* xb = reinterpret_cast<const T*>(P.x.ptr) + m.v * kVec;
    const T* dyb = reinterpret_cast<const T*>(P.y.ptr) + m.v * kVec;
    T* dxb = reinterpret_cast<T*>(P.dx.ptr) + m.v * kVec;
    const T* a0 = P.add0.ptr ? reinterpret_cast<const T*>(P.add0.ptr) + m.v * kVec : nullptr;
    const T* a1 = P.add1.ptr ? reinterpret_cast<const T*>(P.add1.ptr) + m.v * kVec : nullptr;
    for (int p = p0 + m.lane; p < p1; p += m.lanes * kU) {
        uint4 rx[kU], rd[kU], r0[kU], r1[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int pp = p + u * m.lanes;
            if (pp < p1) {
                rx[u] = ld_raw<T>(xb + pix_off(P.x, n, pp, P.W));
                rd[u] = ld_raw<T>(dyb + pix_off(P.y, n, pp, P.W));
                if (a0) r0[u] = ld_raw<T>(a0 + pix_off(P.add0, n, pp, P.W));
                if (a1) r1[u] = ld_raw<T>(a1 + pix_off(P.add1, n, pp, P.W));
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int pp = p + u * m.lanes;
            if (pp >= p1) break;
            float xv[kVec], dv[kVec], o[kVec];
            unpack<T>(rx[u], xv);
            unpack<T>(rd[u], dv);
#pragma unroll
            for (int i = 0; i < kVec; ++i) {
                const float du = act_grad_t<T>(fmaf(xv[i] - mu[i], sc[i], be[i]), dv[i], P.silu);
                o[i] = fmaf(du, sc[i], fmaf(xv[i], k1[i], k0[i]));
            }
            if (a0) {
                float t[kVec];
                unpack<T>(r0[u], t);
#pragma unroll
                for (int i = 0; i < kVec; ++i) o[i] += t[i];
            }
            if (a1) {
                float t[kVec];
                unpack<T>(r1[u], t);
#pragma unroll
                for (int i = 0; i < kVec; ++i) o[i] += t[i];
            }
            store_vec<T>(dxb + pix_off(P.dx, n, pp, P.W), o);
        }
    }
}

// ------------------------------------------------------------------ single-pass GroupNorm (cluster per image)
// One thread-block cluster (1, 2, 4 or 8 CTAs) owns one image; every thread keeps its share of the image in registers
// (at most kHold 16-byte vectors per tensor), the per-channel partial sums of the CTAs meet through distributed shared
// memory, and the result is written from the registers: forward = 1 read + 1 write of the activation instead of
// stats (1 read) + apply (1 read + 1 write); backward = 1 read of (x, dy) instead of 2.  Also halves the launch count of
// the many latency-bound <= 16x16 tensors.
constexpr int kHold = 8;

__device__ __forceinline__ void cluster_sync_() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float* local, unsigned rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(local), r;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(r) : "memory");
    return v;
}
// totals over the cluster of two per-channel arrays (in place); every CTA ends up with the full sums
__device__ __forceinline__ void cluster_channel_total(float* s_a, float* s_b, float* s_ta, float* s_tb, int C, int cs) {
    if (cs == 1) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) { s_ta[c] = s_a[c]; s_tb[c] = s_b[c]; }
        __syncthreads();
        return;
    }
    cluster_sync_();                     // partials of every CTA are in place
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int r = 0; r < cs; ++r) { a += ld_dsmem(s_a + c, r); b += ld_dsmem(s_b + c, r); }
        s_ta[c] = a; s_tb[c] = b;
    }
    cluster_sync_();                     // nobody leaves (or overwrites its partials) while peers still read them
}

template <typename T>
__global__ void __launch_bounds__(256) gn_fwd_fused_kernel(dmu_gn_params P) {
    pdl_trigger();
    pdl_wait();
    constexpr int kVec = Elem<T>::kVec;
    __shared__ float s_a[kMaxC], s_b[kMaxC], s_ta[kMaxC], s_tb[kMaxC];
    __shared__ float s_red[256 * kVec];
    __shared__ float s_gam[kMaxC], s_bet[kMaxC];
    const int n = blockIdx.y, HW = P.H * P.W, C = P.C, cs = gridDim.x;
    RowMap m(C, kVec);
    int p0, p1; chunk_range(HW, p0, p1);
    uint4 r[kHold];
    float a[kVec], q[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) { a[i] = 0.f; q[i] = 0.f; }
    const T* xb = reinterpret_cast<const T*>(P.x.ptr) + m.v * kVec;
#pragma unroll
    for (int u = 0; u < kHold; ++u) {
        const int pp = p0 + m.lane + u * m.lanes;
        r[u] = (m.active && pp < p1) ? ld_raw<T>(xb + pix_off(P.x, n, pp, P.W)) : make_uint4(0u, 0u, 0u, 0u);
    }
    // parameters staged while the tensor loads fly (read again only after the reductions' barriers)
    for (int c = threadIdx.x; c < C; c += blockDim.x) { s_gam[c] = P.gamma[c]; s_bet[c] = P.beta[c]; }
#pragma unroll
    for (int u = 0; u < kHold; ++u) {
        float v[kVec];
        unpack<T>(r[u], v);
#pragma unroll
        for (int i = 0; i < kVec; ++i) { a[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
    }
    block_channel_sum<kVec>(m, a, s_red, s_a, C);
    block_channel_sum<kVec>(m, q, s_red, s_b, C);
    cluster_channel_total(s_a, s_b, s_ta, s_tb, C, cs);
    // group statistics -> per-channel scale/shift (reuse s_a / s_b), raw sums saved for the backward
    const int cpg = C / P.G;
    const float cnt = (float)cpg * (float)HW;
    for (int g = threadIdx.x; g < P.G; g += blockDim.x) {
        float su = 0.f, sq = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) { su += s_ta[c]; sq += s_tb[c]; }
        if (blockIdx.x == 0) {
            P.sums[((int64_t)n * P.G + g) * 2 + 0] = su;
            P.sums[((int64_t)n * P.G + g) * 2 + 1] = sq;
        }
        const float mean = su / cnt;
        const float rstd = rsqrtf(fmaxf(sq / cnt - mean * mean, 0.f) + P.eps);
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            const float sc = rstd * s_gam[c];
            s_a[c] = sc;
            s_b[c] = s_bet[c] - mean * sc;
        }
    }
    __syncthreads();
    if (!m.active) return;
    float sc[kVec], sh[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) { sc[i] = s_a[m.v * kVec + i]; sh[i] = s_b[m.v * kVec + i]; }
    T* yb = reinterpret_cast<T*>(P.y.ptr) + m.v * kVec;
#pragma unroll
    for (int u = 0; u < kHold; ++u) {
        const int pp = p0 + m.lane + u * m.lanes;
        if (pp < p1) {
            float v[kVec];
            unpack<T>(r[u], v);
#pragma unroll
            for (int i = 0; i < kVec; ++i) {
                v[i] = act_fwd_t<T>(fmaf(v[i], sc[i], sh[i]), P.silu);
            }
            store_vec<T>(yb + pix_off(P.y, n, pp, P.W), v);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_fused_kernel(dmu_gn_params P) {
    pdl_trigger();
    pdl_wait();
    constexpr int kVec = Elem<T>::kVec;
    __shared__ float s_mean[kMaxC], s_scale[kMaxC], s_beta[kMaxC], s_rstd[kMaxC];
    __shared__ float s_a[kMaxC], s_b[kMaxC], s_ta[kMaxC], s_tb[kMaxC];
    __shared__ float s_red[256 * kVec];
    __shared__ float s_gam[kMaxC];
    const int n = blockIdx.y, HW = P.H * P.W, C = P.C, cs = gridDim.x;
    RowMap m(C, kVec);
    int p0, p1; chunk_range(HW, p0, p1);
    // the tensor loads go out first: the statistics / affine staging below overlaps their latency (these launches are
    // latency chains: every dependent global round trip is ~0.7 us of a ~6 us kernel)
    uint4 rx[kHold], rd[kHold];
    const T* xb = reinterpret_cast<const T*>(P.x.ptr) + m.v * kVec;
    const T* dyb = reinterpret_cast<const T*>(P.y.ptr) + m.v * kVec;
#pragma unroll
    for (int u = 0; u < kHold; ++u) {
        const int pp = p0 + m.lane + u * m.lanes;
        const bool ok = m.active && pp < p1;
        rx[u] = ok ? ld_raw<T>(xb + pix_off(P.x, n, pp, P.W)) : make_uint4(0u, 0u, 0u, 0u);
        rd[u] = ok ? ld_raw<T>(dyb + pix_off(P.y, n, pp, P.W)) : make_uint4(0u, 0u, 0u, 0u);
    }
    stage_affine(P, n, s_mean, s_scale, s_beta, s_rstd, s_gam);
    __syncthreads();
    float mu[kVec], sc[kVec], be[kVec];
    const int cbase = m.active ? m.v * kVec : 0;
#pragma unroll
    for (int i = 0; i < kVec; ++i) { mu[i] = s_mean[cbase + i]; sc[i] = s_scale[cbase + i]; be[i] = s_beta[cbase + i]; }
    {
        float a[kVec], b[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) { a[i] = 0.f; b[i] = 0.f; }
#pragma unroll
        for (int u = 0; u < kHold; ++u) {
            float xv[kVec], dv[kVec];
            unpack<T>(rx[u], xv);
            unpack<T>(rd[u], dv);
#pragma unroll
            for (int i = 0; i < kVec; ++i) {
                const float d = xv[i] - mu[i];
                const float du = act_grad_t<T>(fmaf(d, sc[i], be[i]), dv[i], P.silu);    // dy == 0 for the padding slots
                a[i] += du;
                b[i] = fmaf(du, d, b[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < kVec; ++i) b[i] *= s_rstd[cbase + i];
        block_channel_sum<kVec>(m, a, s_red, s_a, C);
        block_channel_sum<kVec>(m, b, s_red, s_b, C);
    }
    cluster_channel_total(s_a, s_b, s_ta, s_tb, C, cs);
    const int cpg = C / P.G;
    const float inv_cnt = 1.f / ((float)cpg * (float)HW);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (blockIdx.x == 0) {
            if (P.red) {      // per-image sums, folded over the batch by dmu_gn_param_grads (or by the caller)
                P.red[((int64_t)n * C + c) * 2 + 0] = s_ta[c];
                P.red[((int64_t)n * C + c) * 2 + 1] = s_tb[c];
            }
            if (P.dbeta) atomicAdd(&P.dbeta[c], s_ta[c]);
            if (P.dgamma) atomicAdd(&P.dgamma[c], s_tb[c]);
        }
        const int g0 = (c / cpg) * cpg;
        float A = 0.f, B = 0.f;
        for (int k = g0; k < g0 + cpg; ++k) {
            const float gam = s_gam[k];
            A += gam * s_ta[k];
            B += gam * s_tb[k];
        }
        s_a[c] = A * inv_cnt;
        s_b[c] = B * inv_cnt;
    }
    __syncthreads();
    if (!m.active) return;
    float k0[kVec], k1[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const float rs = s_rstd[cbase + i];
        k1[i] = -rs * rs * s_b[cbase + i];
        k0[i] = -rs * s_a[cbase + i] - mu[i] * k1[i];
    }
    T* dxb = reinterpret_cast<T*>(P.dx.ptr) + m.v * kVec;
    const T* a0 = P.add0.ptr ? reinterpret_cast<const T*>(P.add0.ptr) + m.v * kVec : nullptr;
    const T* a1 = P.add1.ptr ? reinterpret_cast<const T*>(P.add1.ptr) + m.v * kVec : nullptr;
#pragma unroll
    for (int u = 0; u < kHold; ++u) {
        const int pp = p0 + m.lane + u * m.lanes;
        if (pp < p1) {
            float xv[kVec], dv[kVec], o[kVec];
            unpack<T>(rx[u], xv);
            unpack<T>(rd[u], dv);
#pragma unroll
            for (int i = 0; i < kVec; ++i) {
                const float du = act_grad_t<T>(fmaf(xv[i] - mu[i], sc[i], be[i]), dv[i], P.silu);
                o[i] = fmaf(du, sc[i], fmaf(xv[i], k1[i], k0[i]));
            }
            if (a0) {
                float t[kVec];
                unpack<T>(ld_raw<T>(a0 + pix_off(P.add0, n, pp, P.W)), t);
#pragma unroll
                for (int i = 0; i < kVec; ++i) o[i] += t[i];
            }
            if (a1) {
                float t[kVec];
                unpack<T>(ld_raw<T>(a1 + pix_off(P.add1, n, pp, P.W)), t);
#pragma unroll
                for (int i = 0; i < kVec; ++i) o[i] += t[i];
            }
            store_vec<T>(dxb + pix_off(P.dx, n, pp, P.W), o);
        }
    }
}

// Single-pass forward for images too large for the register-held kernel: the CTA's slab of the image is parked in SHARED
// memory (bf16 64x64x64: 64 KB per CTA of an 8-CTA cluster), statistics meet through DSMEM, and the normalised slab is written
// from shared memory: one read + one write of the activation where the two-pass pair reads it twice (the second read only
// comes from L2 while the whole tensor fits there, which a 134 MB activation does not).
template <typename T>
__global__ void __launch_bounds__(256, 3) gn_fwd_smem_kernel(dmu_gn_params P, int per) {
    pdl_trigger();
    pdl_wait();
    constexpr int kVec = Elem<T>::kVec;
    constexpr int kU2 = 4;
    extern __shared__ __align__(16) uint8_t gn_smem[];
    const int n = blockIdx.y, HW = P.H * P.W, C = P.C, cs = gridDim.x;
    uint4* s_x = reinterpret_cast<uint4*>(gn_smem);
    float* s_a = reinterpret_cast<float*>(s_x + (size_t)per * (C / kVec));
    float *s_b = s_a + C, *s_ta = s_b + C, *s_tb = s_ta + C;
    __shared__ float s_red[256 * kVec];
    RowMap m(C, kVec);
    const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
    const int cbase = m.active ? m.v * kVec : 0;
    const T* xb = reinterpret_cast<const T*>(P.x.ptr) + cbase;
    {
        float a[kVec], q[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) { a[i] = 0.f; q[i] = 0.f; }
        if (m.active) {
            // the whole slab goes in flight (kSlabGroups commit groups of this thread's pixels), then it is consumed in order
            const int mine = p1 > p0 + m.lane ? (p1 - p0 - m.lane + m.lanes - 1) / m.lanes : 0;      // pixels of this thread
            const int per_grp = (mine + kSlabGroups - 1) / kSlabGroups;
            for (int g = 0; g < kSlabGroups; ++g) {
                for (int k = g * per_grp; k < min(mine, (g + 1) * per_grp); ++k) {
                    const int pp = p0 + m.lane + k * m.lanes;
                    cp_async16(&s_x[(size_t)(pp - p0) * m.V + m.v], xb + pix_off(P.x, n, pp, P.W));
                }
                cp_async_commit();
            }
            for (int g = 0; g < kSlabGroups; ++g) {
                cp_async_wait_dyn(kSlabGroups - 1 - g);
                for (int k = g * per_grp; k < min(mine, (g + 1) * per_grp); ++k) {
                    const int pp = p0 + m.lane + k * m.lanes;
                    float v[kVec];
                    unpack<T>(s_x[(size_t)(pp - p0) * m.V + m.v], v);
#pragma unroll
                    for (int i = 0; i < kVec; ++i) { a[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
                }
            }
        }
        block_channel_sum<kVec>(m, a, s_red, s_a, C);
        block_channel_sum<kVec>(m, q, s_red, s_b, C);
    }
    cluster_channel_total(s_a, s_b, s_ta, s_tb, C, cs);
    const int cpg = C / P.G;
    const float cnt = (float)cpg * (float)HW;
    for (int g = threadIdx.x; g < P.G; g += blockDim.x) {
        float su = 0.f, sq = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) { su += s_ta[c]; sq += s_tb[c]; }
        if (blockIdx.x == 0) {
            P.sums[((int64_t)n * P.G + g) * 2 + 0] = su;
            P.sums[((int64_t)n * P.G + g) * 2 + 1] = sq;
        }
        const float mean = su / cnt;
        const float rstd = rsqrtf(fmaxf(sq / cnt - mean * mean, 0.f) + P.eps);
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            const float sc = rstd * P.gamma[c];
            s_a[c] = sc;
            s_b[c] = P.beta[c] - mean * sc;
        }
    }
    __syncthreads();
    if (!m.active) return;
    float sc[kVec], sh[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) { sc[i] = s_a[cbase + i]; sh[i] = s_b[cbase + i]; }
    T* yb = reinterpret_cast<T*>(P.y.ptr) + cbase;
    for (int p = p0 + m.lane; p < p1; p += m.lanes) {
        float v[kVec];
        unpack<T>(s_x[(size_t)(p - p0) * m.V + m.v], v);       // this thread's own entries: no barrier needed
#pragma unroll
        for (int i = 0; i < kVec; ++i) v[i] = act_fwd_t<T>(fmaf(v[i], sc[i], sh[i]), P.silu);
        store_vec<T>(yb + pix_off(P.y, n, p, P.W), v);
    }
}

// Single-pass backward for images too large for the register-held kernel above: the CTA's (x, dy) slab is parked in
// SHARED memory (bf16 32x32x64: 2 x 32 KB per CTA of a 4-CTA cluster), so the kernel needs ~64 registers and several
// CTAs share an SM.  One read of (x, dy), one write of dx; the two-pass pair it replaces read both tensors twice and ran
// at < 1 TB/s (36 + 31 us for a 16.8 MB activation).  Dynamic shared memory: [slab x | slab dy | 8 arrays of C floats].
template <typename T>
__global__ void __launch_bounds__(256, 3) gn_bwd_smem_kernel(dmu_gn_params P, int per) {
    pdl_trigger();
    pdl_wait();
    constexpr int kVec = Elem<T>::kVec;
    constexpr int kU2 = 2;
    extern __shared__ __align__(16) uint8_t gn_smem[];
    const int n = blockIdx.y, HW = P.H * P.W, C = P.C, cs = gridDim.x;
    uint4* s_x = reinterpret_cast<uint4*>(gn_smem);
    uint4* s_d = s_x + (size_t)per * (C / kVec);
    float* s_mean = reinterpret_cast<float*>(s_d + (size_t)per * (C / kVec));
    float *s_scale = s_mean + C, *s_beta = s_scale + C, *s_rstd = s_beta + C;
    float *s_a = s_rstd + C, *s_b = s_a + C, *s_ta = s_b + C, *s_tb = s_ta + C;
    __shared__ float s_red[256 * kVec];
    RowMap m(C, kVec);
    const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
    const int cbase = m.active ? m.v * kVec : 0;
    const T* xb = reinterpret_cast<const T*>(P.x.ptr) + cbase;
    const T* dyb = reinterpret_cast<const T*>(P.y.ptr) + cbase;
    // the whole (x, dy) slab goes in flight first; the statistics / affine staging below overlaps its latency
    const int mine = (m.active && p1 > p0 + m.lane) ? (p1 - p0 - m.lane + m.lanes - 1) / m.lanes : 0;      // pixels of this thread
    const int per_grp = (mine + kSlabGroups - 1) / kSlabGroups;
    for (int g = 0; g < kSlabGroups; ++g) {
        for (int k = g * per_grp; k < min(mine, (g + 1) * per_grp); ++k) {
            const int pp = p0 + m.lane + k * m.lanes;
            cp_async16(&s_x[(size_t)(pp - p0) * m.V + m.v], xb + pix_off(P.x, n, pp, P.W));
            cp_async16(&s_d[(size_t)(pp - p0) * m.V + m.v], dyb + pix_off(P.y, n, pp, P.W));
        }
        cp_async_commit();
    }
    stage_affine(P, n, s_mean, s_scale, s_beta, s_rstd);
    __syncthreads();
    float mu[kVec], sc[kVec], be[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) { mu[i] = s_mean[cbase + i]; sc[i] = s_scale[cbase + i]; be[i] = s_beta[cbase + i]; }
    {
        float a[kVec], b[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) { a[i] = 0.f; b[i] = 0.f; }
        if (m.active) {
            for (int g = 0; g < kSlabGroups; ++g) {
                cp_async_wait_dyn(kSlabGroups - 1 - g);
                for (int k = g * per_grp; k < min(mine, (g + 1) * per_grp); ++k) {
                    const int pp = p0 + m.lane + k * m.lanes;
                    float xv[kVec], dv[kVec];
                    uint4* sd = &s_d[(size_t)(pp - p0) * m.V + m.v];
                    unpack<T>(s_x[(size_t)(pp - p0) * m.V + m.v], xv);
                    unpack<T>(*sd, dv);
#pragma unroll
                    for (int i = 0; i < kVec; ++i) {
                        const float d = xv[i] - mu[i];
                        const float du = act_grad_t<T>(fmaf(d, sc[i], be[i]), dv[i], P.silu);
                        dv[i] = du;
                        a[i] += du;
                        b[i] = fmaf(du, d, b[i]);
                    }
                    // du replaces dy in the slab (in the tensor's own dtype): the second pass needs no transcendental
                    store_vec<T>(reinterpret_cast<T*>(sd), dv);
                }
            }
#pragma unroll
            for (int i = 0; i < kVec; ++i) b[i] *= s_rstd[cbase + i];
        }
        block_channel_sum<kVec>(m, a, s_red, s_a, C);
        block_channel_sum<kVec>(m, b, s_red, s_b, C);
    }
    cluster_channel_total(s_a, s_b, s_ta, s_tb, C, cs);
    const int cpg = C / P.G;
    const float inv_cnt = 1.f / ((float)cpg * (float)HW);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (blockIdx.x == 0) {
            if (P.red) {
                P.red[((int64_t)n * C + c) * 2 + 0] = s_ta[c];
                P.red[((int64_t)n * C + c) * 2 + 1] = s_tb[c];
            }
            if (P.dbeta) atomicAdd(&P.dbeta[c], s_ta[c]);
            if (P.dgamma) atomicAdd(&P.dgamma[c], s_tb[c]);
        }
        const int g0 = (c / cpg) * cpg;
        float A = 0.f, B = 0.f;
        for (int k = g0; k < g0 + cpg; ++k) {
            const float gam = P.gamma[k];
            A += gam * s_ta[k];
            B += gam * s_tb[k];
        }
        s_a[c] = A * inv_cnt;
        s_b[c] = B * inv_cnt;
    }
    __syncthreads();
    if (!m.active) return;
    float k0[kVec], k1[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
        const float rs = s_rstd[cbase + i];
        k1[i] = -rs * rs * s_b[cbase + i];
        k0[i] = -rs * s_a[cbase + i] - mu[i] * k1[i];
    }
    T* dxb = reinterpret_cast<T*>(P.dx.ptr) + cbase;
    const T* a0 = P.add0.ptr ? reinterpret_cast<const T*>(P.add0.ptr) + cbase : nullptr;
    const T* a1 = P.add1.ptr ? reinterpret_cast<const T*>(P.add1.ptr) + cbase : nullptr;
    for (int p = p0 + m.lane; p < p1; p += m.lanes * kU2) {
        uint4 r0[kU2], r1[kU2];
#pragma unroll
        for (int u = 0; u < kU2; ++u) {
            const int pp = p + u * m.lanes;
            if (pp < p1) {
                if (a0) r0[u] = ld_raw<T>(a0 + pix_off(P.add0, n, pp, P.W));
                if (a1) r1[u] = ld_raw<T>(a1 + pix_off(P.add1, n, pp, P.W));
            }
        }
#pragma unroll
        for (int u = 0; u < kU2; ++u) {
            const int pp = p + u * m.lanes;
            if (pp >= p1) break;
            float xv[kVec], dv[kVec], o[kVec];
            unpack<T>(s_x[(size_t)(pp - p0) * m.V + m.v], xv);       // this thread's own entries: no barrier needed
            unpack<T>(s_d[(size_t)(pp - p0) * m.V + m.v], dv);          // du, written by this thread in the first pass
#pragma unroll
            for (int i = 0; i < kVec; ++i) o[i] = fmaf(dv[i], sc[i], fmaf(xv[i], k1[i], k0[i]));
            if (a0) {
                float t[kVec];
                unpack<T>(r0[u], t);
#pragma unroll
                for (int i = 0; i < kVec; ++i) o[i] += t[i];
            }
            if (a1) {
                float t[kVec];
                unpack<T>(r1[u], t);
#pragma unroll
                for (int i = 0; i < kVec; ++i) o[i] += t[i];
            }
            store_vec<T>(dxb + pix_off(P.dx, n, pp, P.W), o);
        }
    }
}

// dgamma[c] += sum_n red[n,c,1], dbeta[c] += sum_n red[n,c,0] for a whole table of GroupNorm layers in one launch:
// the batch reduction of the affine-parameter gradients, kept out of the per-layer kernels (where it was a same-address
// atomic from every CTA of every image).
struct GnPgDesc { const float* red; float* dgamma; float* dbeta; int32_t C; int32_t count; };
__global__ void __launch_bounds__(256) gn_param_grads_kernel(const GnPgDesc* __restrict__ table, int N) {
    const GnPgDesc d = table[blockIdx.y];
    if (d.count > 0) N = d.count;      // rows of per-tile sums written by a fused dgrad epilogue (conv_tc.cu)
    // thread = (channel, quantity); 256 threads cover 128 channels x 2
    const int e0 = blockIdx.x * 256 + threadIdx.x;
    if (e0 >= d.C * 2) return;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    int n = 0;
    for (; n + 3 < N; n += 4) {
        t0 += d.red[(int64_t)(n + 0) * d.C * 2 + e0];
        t1 += d.red[(int64_t)(n + 1) * d.C * 2 + e0];
        t2 += d.red[(int64_t)(n + 2) * d.C * 2 + e0];
        t3 += d.red[(int64_t)(n + 3) * d.C * 2 + e0];
    }
    for (; n < N; ++n) t0 += d.red[(int64_t)n * d.C * 2 + e0];
    const float t = (t0 + t1) + (t2 + t3);
    const int c = e0 >> 1;
    if (e0 & 1) d.dgamma[c] += t;
    else d.dbeta[c] += t;
}

// cluster size for the single-pass kernels: smallest of {1,2,4,8} whose per-CTA pixel slab fits kHold vectors per thread
static int gn_fused_cluster(int HW, int C, int vec) {
    const int lanes = 256 / (C / vec);
    if (lanes < 1) return 0;
    for (int cs = 1; cs <= 8; cs <<= 1) {
        const int per = (HW + cs - 1) / cs;
        if (per <= lanes * kHold) return cs <= HW ? cs : 0;
    }
    return 0;
}

template <typename K>
static int launch_cluster(K kernel, int cs, int N, cudaStream_t stream, const dmu_gn_params& p) {
    cudaError_t e = launch_pdl(kernel, dim3(cs, N), dim3(256), 0, stream, dim3(cs, 1, 1), p);
    if (e != cudaSuccess) return fail("GroupNorm cluster launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------ column sums
// grid (pixel chunks, N): per-CTA partial sums, then atomics (out_nc / out_c must hold the running value: zero-initialised
// by the caller when a fresh sum is wanted).
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(dmu_tensor4 X, int H, int W, int C, float* out_nc, int64_t pitch,
                                                     float* out_c, float scale) {
    pdl_trigger();
    pdl_wait();
    constexpr int kVec = Elem<T>::kVec;
    __shared__ float s[kMaxC];
    __shared__ float s_red[256 * kVec];
    const int n = blockIdx.y, HW = H * W;
    RowMap m(C, kVec);
    int p0, p1; chunk_range(HW, p0, p1);
    float a[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) a[i] = 0.f;
    if (m.active) {
        const T* xb = reinterpret_cast<const T*>(X.ptr) + m.v * kVec;
        for (int p = p0 + m.lane; p < p1; p += m.lanes * kUnroll) {
            uint4 r[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int pp = p + u * m.lanes;
                r[u] = pp < p1 ? ld_raw<T>(xb + pix_off(X, n, pp, W)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                float v[kVec];
                unpack<T>(r[u], v);
#pragma unroll
                for (int i = 0; i < kVec; ++i) a[i] += v[i];
            }
        }
    }
    block_channel_sum<kVec>(m, a, s_red, s, C);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float v = s[c] * scale;
        if (out_nc) atomicAdd(&out_nc[(int64_t)n * pitch + c], v);
        if (out_c) atomicAdd(&out_c[c], v);
    }
}

// One launch for a whole table of column sums (the bias gradients and time-projection sums of one part of the backward:
// ~20 latency-bound launches become one).  1-D grid: descriptor i owns CTAs [cta0_i, cta0_i + N_i * chunks_i).
template <typename T>
__global__ void __launch_bounds__(256) colsum_multi_kernel(const dmu_colsum_desc* __restrict__ table, int n_desc) {
    pdl_trigger();
    pdl_wait();
    __shared__ int s_idx;
    for (int i = threadIdx.x; i < n_desc; i += blockDim.x) {
        const int c0 = table[i].cta0, cnt = table[i].N * table[i].chunks;
        if ((int)blockIdx.x >= c0 && (int)blockIdx.x < c0 + cnt) s_idx = i;
    }
    __syncthreads();
    const dmu_colsum_desc d = table[s_idx];
    const int local = (int)blockIdx.x - d.cta0;
    const int n = local / d.chunks, chunk = local % d.chunks;
    constexpr int kVec = Elem<T>::kVec;
    __shared__ float s[kMaxC];
    __shared__ float s_red[256 * kVec];
    const int HW = d.H * d.W, C = d.C;
    RowMap m(C, kVec);
    const int per = (HW + d.chunks - 1) / d.chunks;
    const int p0 = chunk * per, p1 = min(HW, p0 + per);
    float a[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) a[i] = 0.f;
    if (m.active) {
        const T* xb = reinterpret_cast<const T*>(d.x.ptr) + m.v * kVec;
        for (int p = p0 + m.lane; p < p1; p += m.lanes * kUnroll) {
            uint4 r[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int pp = p + u * m.lanes;
                r[u] = pp < p1 ? ld_raw<T>(xb + pix_off(d.x, n, pp, d.W)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                float v[kVec];
                unpack<T>(r[u], v);
#pragma unroll
                for (int i = 0; i < kVec; ++i) a[i] += v[i];
            }
        }
    }
    block_channel_sum<kVec>(m, a, s_red, s, C);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float v = s[c] * d.scale;
        if (d.out_nc) atomicAdd(&d.out_nc[(int64_t)n * d.pitch + c], v);
        if (d.out_c) atomicAdd(&d.out_c[c], v);
    }
}

// ------------------------------------------------------------------ SiLU + global average pool (EnergyNet head)
// models/energy_based.py:79-83:  pooled[n,c] = mean_p silu(x[n,p,c]);   backward: dx[n,p,c] = silu'(x) * g[n,c] * scale
template <typename T>
__global__ void __launch_bounds__(256) silu_pool_fwd_kernel(dmu_tensor4 X, int H, int W, int C, float* out, int64_t pitch, float scale) {
    constexpr int kVec = Elem<T>::kVec;
    __shared__ float s[kMaxC];
    __shared__ float s_red[256 * kVec];
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y, HW = H * W;
    RowMap m(C, kVec);
    int p0, p1; chunk_range(HW, p0, p1);
    float a[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) a[i] = 0.f;
    if (m.active) {
        const T* xb = reinterpret_cast<const T*>(X.ptr) + m.v * kVec;
        for (int p = p0 + m.lane; p < p1; p += m.lanes) {
            float v[kVec];
            unpack<T>(ld_raw<T>(xb + pix_off(X, n, p, W)), v);
#pragma unroll
            for (int i = 0; i < kVec; ++i) a[i] += v[i] / (1.f + expf(-v[i]));
        }
    }
    block_channel_sum<kVec>(m, a, s_red, s, C);
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&out[(int64_t)n * pitch + c], s[c] * scale);
}
template <typename T>
__global__ void __launch_bounds__(256) silu_pool_bwd_kernel(dmu_tensor4 X, dmu_tensor4 DX, int H, int W, int C, const float* g, int64_t pitch, float scale) {
    constexpr int kVec = Elem<T>::kVec;
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y, HW = H * W;
    RowMap m(C, kVec);
    if (!m.active) return;
    int p0, p1; chunk_range(HW, p0, p1);
    float gv[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) gv[i] = g[(int64_t)n * pitch + m.v * kVec + i] * scale;
    const T* xb = reinterpret_cast<const T*>(X.ptr) + m.v * kVec;
    T* db = reinterpret_cast<T*>(DX.ptr) + m.v * kVec;
    for (int p = p0 + m.lane; p < p1; p += m.lanes) {
        float v[kVec];
        unpack<T>(ld_raw<T>(xb + pix_off(X, n, p, W)), v);
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const float sg = 1.f / (1.f + expf(-v[i]));
            v[i] = gv[i] * (sg * (1.f + v[i] * (1.f - sg)));
        }
        store_vec<T>(db + pix_off(DX, n, p, W), v);
    }
}

// ------------------------------------------------------------------ second-order pieces (EnergyBasedLoss gradient penalty)
// utils/losses.py:277-285 differentiates grad_x E with create_graph=True, i.e. it back-propagates through the BACKWARD of the
// energy network.  Convolutions are linear (their double backward is the existing fprop / dgrad / wgrad kernels); the two
// nonlinear backward ops need their own derivative:
//
// (1) B(x, gamma, beta, dy) = backward of y = silu(GroupNorm(x)):   with xh = (x-mu) r, u = gamma xh + beta, du = dy phi'(u),
//     t = gamma du, M1 = mean_g t, M2 = mean_g (t xh):    dx = r (t - M1 - xh M2)            (means over one image's group)
//     Given a cotangent c of dx, S = sum c dx = r sum_i t_i w_i with w = c - mean_g c - xh mean_g(c xh).  Then
//        dS/d(dy) = r w gamma phi'(u)
//        e := dS/du = r w gamma dy phi''(u)             (phi'(u) inside t)
//        q := dS/dxh = e gamma - r (c M2 + t mean_g(c xh))
//        dS/dx  = r (q - mean_g q - xh mean_g(q xh) - r xh mean_g(t w))     (last term: the explicit factor r of dx)
//        dS/dgamma_c = sum (r w du + e xh),   dS/dbeta_c = sum e
// (2) P(h, g) = backward of pooled[n,c] = s sum_p silu(h):  dh = phi'(h) g[n,c] s;  given a cotangent c of dh:
//        dS/dh = c phi''(h) g s,   dS/dg[n,c] = s sum_p c phi'(h)
__device__ __forceinline__ void silu_d12(float u, int silu, float& d1, float& d2) {
    if (!silu) { d1 = 1.f; d2 = 0.f; return; }
    const float sg = 1.f / (1.f + expf(-u));
    d1 = sg * (1.f + u * (1.f - sg));
    d2 = sg * (1.f - sg) * (2.f + u * (1.f - 2.f * sg));
}

template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_bwd_kernel(dmu_gn_bwd2_params P) {
    constexpr int kVec = Elem<T>::kVec;
    __shared__ float s_mean[kMaxC], s_scale[kMaxC], s_beta[kMaxC], s_rstd[kMaxC];
    __shared__ float s_q[5][kMaxC];          // per-channel sums, then per-channel copies of their group totals
    __shared__ float s_red[256 * kVec];
    const int n = blockIdx.x, HW = P.H * P.W, C = P.C, cpg = C / P.G;
    {   // statistics of image n (same convention as stage_affine)
        const float cnt = (float)cpg * (float)HW;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const int g = c / cpg;
            const float su = P.sums[((int64_t)n * P.G + g) * 2 + 0], sq = P.sums[((int64_t)n * P.G + g) * 2 + 1];
            const float mean = su / cnt;
            const float rstd = rsqrtf(fmaxf(sq / cnt - mean * mean, 0.f) + P.eps);
            s_mean[c] = mean; s_rstd[c] = rstd; s_scale[c] = P.gamma[c]; s_beta[c] = P.beta[c];
        }
    }
    __syncthreads();
    RowMap m(C, kVec);
    const int cb = m.active ? m.v * kVec : 0;
    float mu[kVec], rs[kVec], ga[kVec], be[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) { mu[i] = s_mean[cb + i]; rs[i] = s_rstd[cb + i]; ga[i] = s_scale[cb + i]; be[i] = s_beta[cb + i]; }
    const T* xb = reinterpret_cast<const T*>(P.x.ptr) + cb;
    const T* dyb = reinterpret_cast<const T*>(P.dy.ptr) + cb;
    const T* ccb = reinterpret_cast<const T*>(P.c.ptr) + cb;
    const float inv_m = 1.f / ((float)cpg * (float)HW);

    auto group_totals = [&](int nq) {   // s_q[k][c] <- sum over the channels of c's group of s_q[k][.]
        __syncthreads();
        float tot[5];
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const int g0 = (c / cpg) * cpg;
            for (int k = 0; k < nq; ++k) {
                float t = 0.f;
                for (int j = g0; j < g0 + cpg; ++j) t += s_q[k][j];
                tot[k] = t;
            }
            // all reads of this group's channels by this thread are done; other threads of the group read the same values,
            // so write back only after a barrier
            for (int k = 0; k < nq; ++k) s_red[k * C + c] = tot[k];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x)
            for (int k = 0; k < nq; ++k) s_q[k][c] = s_red[k * C + c];
        __syncthreads();
    };

    // ---- pass 1: A = sum c, Bc = sum c xh, T1 = sum t, T2 = sum t xh
    {
        float a0[kVec], a1[kVec], a2[kVec], a3[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) { a0[i] = a1[i] = a2[i] = a3[i] = 0.f; }
        if (m.active) {
            for (int p = m.lane; p < HW; p += m.lanes) {
                float xv[kVec], dv[kVec], cv[kVec];
                unpack<T>(ld_raw<T>(xb + pix_off(P.x, n, p, P.W)), xv);
                unpack<T>(ld_raw<T>(dyb + pix_off(P.dy, n, p, P.W)), dv);
                unpack<T>(ld_raw<T>(ccb + pix_off(P.c, n, p, P.W)), cv);
#pragma unroll
                for (int i = 0; i < kVec; ++i) {
                    const float xh = (xv[i] - mu[i]) * rs[i];
                    float d1, d2;
                    silu_d12(fmaf(ga[i], xh, be[i]), P.silu, d1, d2);
                    const float t = ga[i] * dv[i] * d1;
                    a0[i] += cv[i]; a1[i] = fmaf(cv[i], xh, a1[i]); a2[i] += t; a3[i] = fmaf(t, xh, a3[i]);
                }
            }
        }
        block_channel_sum<kVec>(m, a0, s_red, s_q[0], C);
        block_channel_sum<kVec>(m, a1, s_red, s_q[1], C);
        block_channel_sum<kVec>(m, a2, s_red, s_q[2], C);
        block_channel_sum<kVec>(m, a3, s_red, s_q[3], C);
    }
    group_totals(4);
    float Am[kVec], Bm[kVec], M1[kVec], M2[kVec];      // group means
#pragma unroll
    for (int i = 0; i < kVec; ++i) { Am[i] = s_q[0][cb + i] * inv_m; Bm[i] = s_q[1][cb + i] * inv_m; M1[i] = s_q[2][cb + i] * inv_m; M2[i] = s_q[3][cb + i] * inv_m; }
    __syncthreads();

    // ---- pass 2: Q1 = sum q, Q2 = sum q xh, K = sum t w, dgamma, dbeta
    {
        float a0[kVec], a1[kVec], a2[kVec], a3[kVec], a4[kVec];
#pragma unroll
        for (int i = 0; i < kVec; ++i) { a0[i] = a1[i] = a2[i] = a3[i] = a4[i] = 0.f; }
        if (m.active) {
            for (int p = m.lane; p < HW; p += m.lanes) {
                float xv[kVec], dv[kVec], cv[kVec];
                unpack<T>(ld_raw<T>(xb + pix_off(P.x, n, p, P.W)), xv);
                unpack<T>(ld_raw<T>(dyb + pix_off(P.dy, n, p, P.W)), dv);
                unpack<T>(ld_raw<T>(ccb + pix_off(P.c, n, p, P.W)), cv);
#pragma unroll
                for (int i = 0; i < kVec; ++i) {
                    const float xh = (xv[i] - mu[i]) * rs[i];
                    float d1, d2;
                    silu_d12(fmaf(ga[i], xh, be[i]), P.silu, d1, d2);
                    const float du = dv[i] * d1, t = ga[i] * du;
                    const float w = cv[i] - Am[i] - xh * Bm[i];
                    const float e = rs[i] * w * ga[i] * dv[i] * d2;
                    const float q = e * ga[i] - rs[i] * (cv[i] * M2[i] + t * Bm[i]);
                    a0[i] += q; a1[i] = fmaf(q, xh, a1[i]); a2[i] = fmaf(t, w, a2[i]);
                    a3[i] += rs[i] * w * du + e * xh;      // dgamma
                    a4[i] += e;                            // dbeta
                }
            }
        }
        block_channel_sum<kVec>(m, a0, s_red, s_q[0], C);
        block_channel_sum<kVec>(m, a1, s_red, s_q[1], C);
        block_channel_sum<kVec>(m, a2, s_red, s_q[2], C);
        block_channel_sum<kVec>(m, a3, s_red, s_q[3], C);
        block_channel_sum<kVec>(m, a4, s_red, s_q[4], C);
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (P.dgamma) atomicAdd(&P.dgamma[c], s_q[3][c]);
        if (P.dbeta) atomicAdd(&P.dbeta[c], s_q[4][c]);
    }
    group_totals(3);
    float Q1[kVec], Q2[kVec], Km[kVec];
#pragma unroll
    for (int i = 0; i < kVec; ++i) { Q1[i] = s_q[0][cb + i] * inv_m; Q2[i] = s_q[1][cb + i] * inv_m; Km[i] = s_q[2][cb + i] * inv_m; }

    // ---- pass 3: outputs
    if (!m.active) return;
    T* gxb = reinterpret_cast<T*>(P.gx.ptr) + cb;
    T* gdb = P.gdy.ptr ? reinterpret_cast<T*>(P.gdy.ptr) + cb : nullptr;
    for (int p = m.lane; p < HW; p += m.lanes) {
        float xv[kVec], dv[kVec], cv[kVec], ox[kVec], od[kVec];
        unpack<T>(ld_raw<T>(xb + pix_off(P.x, n, p, P.W)), xv);
        unpack<T>(ld_raw<T>(dyb + pix_off(P.dy, n, p, P.W)), dv);
        unpack<T>(ld_raw<T>(ccb + pix_off(P.c, n, p, P.W)), cv);
#pragma unroll
        for (int i = 0; i < kVec; ++i) {
            const float xh = (xv[i] - mu[i]) * rs[i];
            float d1, d2;
            silu_d12(fmaf(ga[i], xh, be[i]), P.silu, d1, d2);
            const float t = ga[i] * dv[i] * d1;
            const float w = cv[i] - Am[i] - xh * Bm[i];
            const float e = rs[i] * w * ga[i] * dv[i] * d2;
            const float q = e * ga[i] - rs[i] * (cv[i] * M2[i] + t * Bm[i]);
            ox[i] = rs[i] * (q - Q1[i] - xh * Q2[i] - rs[i] * xh * Km[i]);
            od[i] = rs[i] * w * ga[i] * d1;
        }
        store_vec<T>(gxb + pix_off(P.gx, n, p, P.W), ox);
        if (gdb) store_vec<T>(gdb + pix_off(P.gdy, n, p, P.W), od);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) silu_pool_bwd_bwd_kernel(dmu_tensor4 X, dmu_tensor4 Cc, dmu_tensor4 GX, int H, int W, int C, const float* g, int64_t pitch,
                                                                float* gg, int64_t gg_pitch, float scale) {
    constexpr int kVec = Elem<T>::kVec;
    __shared__ float s[kMaxC];
    __shared__ float s_red[256 * kVec];
    const int n = blockIdx.y, HW = H * W;
    RowMap m(C, kVec);
    int p0, p1; chunk_range(HW, p0, p1);
    float a[kVec], gv[kVec];
    const int cb = m.active ? m.v * kVec : 0;
#pragma unroll
    for (int i = 0; i < kVec; ++i) { a[i] = 0.f; gv[i] = g[(int64_t)n * pitch + cb + i] * scale; }
    if (m.active) {
        const T* xb = reinterpret_cast<const T*>(X.ptr) + cb;
        const T* cc = reinterpret_cast<const T*>(Cc.ptr) + cb;
        T* ob = reinterpret_cast<T*>(GX.ptr) + cb;
        for (int p = p0 + m.lane; p < p1; p += m.lanes) {
            float v[kVec], cv[kVec];
            unpack<T>(ld_raw<T>(xb + pix_off(X, n, p, W)), v);
            unpack<T>(ld_raw<T>(cc + pix_off(Cc, n, p, W)), cv);
#pragma unroll
            for (int i = 0; i < kVec; ++i) {
                float d1, d2;
                silu_d12(v[i], 1, d1, d2);
                a[i] = fmaf(cv[i], d1, a[i]);
                v[i] = cv[i] * d2 * gv[i];
            }
            store_vec<T>(ob + pix_off(GX, n, p, W), v);
        }
    }
    block_channel_sum<kVec>(m, a, s_red, s, C);
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&gg[(int64_t)n * gg_pitch + c], s[c] * scale);
}

// ------------------------------------------------------------------ activations (fp32 rows)
__device__ __forceinline__ float act_f(float x, int kind) {
    if (kind == 0) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    if (kind == 1) return x / (1.f + expf(-x));
    return logf(x);
}
__device__ __forceinline__ float act_df(float x, int kind) {
    if (kind == 0) {
        const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
        const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
        return cdf + x * pdf;
    }
    if (kind == 1) {
        const float s = 1.f / (1.f + expf(-x));
        return s * (1.f + x * (1.f - s));
    }
    return 1.f / x;
}
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int kind) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = act_f(x[i], kind);
}
__global__ void act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int64_t n, int kind) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dx[i] = dy[i] * act_df(x[i], kind);
}

__global__ void sinusoidal_kernel(const void* t, int t_is_float, float* emb, int64_t batch, int dim, float neg_k) {
    const int half = dim / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < batch * half; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / half;
        const int j = (int)(i % half);
        const float tv = t_is_float ? reinterpret_cast<const float*>(t)[b] : (float)reinterpret_cast<const int64_t*>(t)[b];
        const float f = expf(__fmul_rn((float)j, neg_k));
        const float a = __fmul_rn(tv, f);
        emb[b * dim + j] = sinf(a);
        emb[b * dim + half + j] = cosf(a);
    }
}

// ------------------------------------------------------------------ repack
// Every layout change below is a permutation of a [A][B][RS] fp32 tensor whose source runs are contiguous, so each CTA
// stages a tile in shared memory with coalesced (128-byte) reads and writes it out in destination order, again in
// contiguous runs: HBM traffic = one read + one write of the filter.  (The first version gathered element-wise with a
// stride of RS floats: 8x read amplification, 143 us for the 16 M parameters of the UNet against ~25 us of traffic.)
constexpr int kRepackTile = 16 * 16 * 36;             // floats: kind 1 holds (16 o x RS <= 16) rows of 32 + 4 input channels

__device__ __forceinline__ void repack_generic(const dmu_repack_desc& d) {
    const int64_t RS = (int64_t)d.R * d.S;
    const int64_t total = (int64_t)d.O * d.I * RS;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (d.kind == 3) {
            // gradient staging [O][R][S][I] -> stored layout [O][I][R][S]
            const int64_t rs = i % RS, ci = (i / RS) % d.I, o = i / (RS * d.I);
            st_from_float(d.dst, i, d.dst_dtype, d.src[(o * RS + rs) * d.I + ci]);
            continue;
        }
        // i enumerates the destination [O][R][S][I]
        const int64_t ci = i % d.I, rs = (i / d.I) % RS, o = i / (d.I * RS);
        int64_t src;
        if (d.kind == 0) src = (o * d.I + ci) * RS + rs;        // OIHW
        else if (d.kind == 1) src = (ci * d.O + o) * RS + rs;   // IOHW (ConvTranspose2d)
        else src = i;
        st_from_float(d.dst, i, d.dst_dtype, d.src[src]);
    }
}

// 8 consecutive destination elements (16-byte aligned) from 8 floats
__device__ __forceinline__ void repack_store8(void* dst, int64_t idx, int dtype, const float* v) {
    if (dtype == DMU_BF16) {
        store_vec<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(dst) + idx, v);
    } else {
        float4* q = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + idx);
        q[0] = make_float4(v[0], v[1], v[2], v[3]);
        q[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// Shared-memory tile layout [rs][IB + 4] (kinds 0, 3) / [16 o x RS][32 + 4] (kind 1): the +4 pitch keeps the transposing
// side at <= 2-way bank conflicts and the row-wise side 16-byte aligned.  Global accesses are 16 bytes per thread with
// four independent requests in flight.
template <int RS>
__device__ __forceinline__ void repack_tiled(const dmu_repack_desc& d, float* __restrict__ tile) {
    if (d.kind == 0 || d.kind == 3) {
        // per output channel o: [I][RS] <-> [RS][I]; work item = (o, block of IB input channels)
        const int IB = d.I % 256 == 0 ? 256 : d.I < 256 ? d.I : d.I % 128 == 0 ? 128 : d.I % 64 == 0 ? 64 : 32;
        const int IBP = IB + 4;
        const int nib = d.I / IB;
        const int run4 = IB * RS / 4;
        for (int item = blockIdx.x; item < d.O * nib; item += gridDim.x) {
            const int o = item / nib, i0 = (item % nib) * IB;
            __syncthreads();
            if (d.kind == 0) {
                const float4* src = reinterpret_cast<const float4*>(d.src + ((int64_t)o * d.I + i0) * RS);   // contiguous [IB][RS]
                for (int j0 = threadIdx.x; j0 < run4; j0 += 4 * 256) {
                    float4 r[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (j0 + u * 256 < run4) r[u] = __ldg(src + j0 + u * 256);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (j0 + u * 256 >= run4) break;
                        const int j = (j0 + u * 256) * 4;
                        const float e[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) tile[((j + q) % RS) * IBP + (j + q) / RS] = e[q];
                    }
                }
                __syncthreads();
                const int c8n = IB / 8;
                for (int idx = threadIdx.x; idx < RS * c8n; idx += 256) {
                    const int rs = idx / c8n, c8 = idx % c8n;
                    const float4 lo = *reinterpret_cast<const float4*>(&tile[rs * IBP + c8 * 8]);
                    const float4 hi = *reinterpret_cast<const float4*>(&tile[rs * IBP + c8 * 8 + 4]);
                    const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                    repack_store8(d.dst, ((int64_t)o * RS + rs) * d.I + i0 + c8 * 8, d.dst_dtype, v);
                }
            } else {
                const int c4n = IB / 4;
                for (int i0_ = threadIdx.x; i0_ < RS * c4n; i0_ += 4 * 256) {
                    float4 r[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int idx = i0_ + u * 256;
                        if (idx < RS * c4n) r[u] = __ldg(reinterpret_cast<const float4*>(d.src + ((int64_t)o * RS + idx / c4n) * d.I + i0) + idx % c4n);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int idx = i0_ + u * 256;
                        if (idx >= RS * c4n) break;
                        *reinterpret_cast<float4*>(&tile[(idx / c4n) * IBP + (idx % c4n) * 4]) = r[u];
                    }
                }
                __syncthreads();
                const int64_t dst0 = ((int64_t)o * d.I + i0) * RS;
                for (int j4 = threadIdx.x; j4 < run4; j4 += 256) {
                    const int j = j4 * 4;
                    float e[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) e[q] = tile[((j + q) % RS) * IBP + (j + q) / RS];
                    if (d.dst_dtype == DMU_F32) {
                        *reinterpret_cast<float4*>(reinterpret_cast<float*>(d.dst) + dst0 + j) = make_float4(e[0], e[1], e[2], e[3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) st_from_float(d.dst, dst0 + j + q, d.dst_dtype, e[q]);
                    }
                }
            }
        }
        return;
    }
    // kind 1: src [I][O][RS] -> dst [O][RS][I]; work item = 32 input channels x 16 output channels
    const int nbo = d.O / 16, nbi = d.I / 32;
    constexpr int run = 16 * RS, run4 = run / 4, pitch = 36;
    for (int item = blockIdx.x; item < nbo * nbi; item += gridDim.x) {
        const int o0 = (item % nbo) * 16, i0 = (item / nbo) * 32;
        __syncthreads();
        // lane -> (8 input channels, 4 consecutive float4 of the run): 64-byte runs per row, <= 2-way bank conflicts
        for (int b0 = threadIdx.x; b0 < 32 * run4; b0 += 4 * 256) {
            float4 r[4];
            int ci_[4], k_[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = b0 + u * 256;                       // idx = ((k4 / 4) * 32 + ci) * 4 + k4 % 4
                const int k4 = (idx / 128) * 4 + idx % 4, ci = (idx / 4) % 32;
                ci_[u] = ci; k_[u] = k4 * 4;
                if (idx < 32 * run4) r[u] = __ldg(reinterpret_cast<const float4*>(d.src + ((int64_t)(i0 + ci) * d.O + o0) * RS) + k4);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (b0 + u * 256 >= 32 * run4) break;
                const float e[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) tile[(k_[u] + q) * pitch + ci_[u]] = e[q];
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < run * 4; idx += 256) {                        // row k, 8 consecutive ci
            const int k = idx / 4, c8 = idx % 4;
            const float4 lo = *reinterpret_cast<const float4*>(&tile[k * pitch + c8 * 8]);
            const float4 hi = *reinterpret_cast<const float4*>(&tile[k * pitch + c8 * 8 + 4]);
            const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            repack_store8(d.dst, ((int64_t)o0 * RS + k) * d.I + i0 + c8 * 8, d.dst_dtype, v);
        }
    }
}

__global__ void __launch_bounds__(256) repack_kernel(const dmu_repack_desc* __restrict__ descs) {
    __shared__ float tile[kRepackTile];
    const dmu_repack_desc d = descs[blockIdx.y];
    const int RS = d.R * d.S;
    const bool tiled = d.kind != 2 && d.I % 32 == 0 && d.O % 16 == 0 && ((reinterpret_cast<uintptr_t>(d.src) | reinterpret_cast<uintptr_t>(d.dst)) & 15) == 0;
    if (tiled && RS == 9) repack_tiled<9>(d, tile);
    else if (tiled && RS == 16) repack_tiled<16>(d, tile);
    else if (tiled && RS == 1 && d.kind == 1) repack_tiled<1>(d, tile);      // [I][O] -> [O][I] transpose of a Linear weight
    else repack_generic(d);
}

__global__ void copy4_kernel(dmu_tensor4 S, dmu_tensor4 D, int N, int H, int W, int C) {
    const int64_t total = (int64_t)N * H * W * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int w = (int)((i / C) % W);
        const int h = (int)((i / ((int64_t)C * W)) % H);
        const int n = (int)(i / ((int64_t)C * W * H));
        const float v = ld_as_float(S.ptr, n * S.sn + h * S.sh + w * S.sw + c * S.sc, S.dtype);
        st_from_float(D.ptr, n * D.sn + h * D.sh + w * D.sw + c * D.sc, D.dtype, v);
    }
}

// ------------------------------------------------------------------ Adam + EMA
__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ ema, int64_t n, float lr, float b1,
                                                       float b2, float eps, float wd, float bc1, float bc2_sqrt, float decay, float gscale,
                                                       const int64_t* __restrict__ step_dev) {
    if (step_dev) {      // bias corrections from the device-resident step count (graph replay): one double pow per CTA
        __shared__ float s_bc[2];
        if (threadIdx.x == 0) {
            const double st = (double)*step_dev;
            s_bc[0] = (float)(1.0 - pow((double)b1, st));
            s_bc[1] = (float)sqrt(1.0 - pow((double)b2, st));
        }
        __syncthreads();
        bc1 = s_bc[0]; bc2_sqrt = s_bc[1];
    }
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (int64_t)gridDim.x * blockDim.x * 4) {
        if (i + 3 < n) {
            float4 pv = *reinterpret_cast<float4*>(p + i), gv = *reinterpret_cast<const float4*>(g + i);
            float4 mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
            float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float gr = gp[k] * gscale + wd * pp[k];
                mp[k] = b1 * mp[k] + (1.f - b1) * gr;
                vp[k] = b2 * vp[k] + (1.f - b2) * gr * gr;
                const float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
                pp[k] -= (lr / bc1) * (mp[k] / denom);
            }
            *reinterpret_cast<float4*>(p + i) = pv;
            *reinterpret_cast<float4*>(m + i) = mv;
            *reinterpret_cast<float4*>(v + i) = vv;
            if (ema) {
                float4 ev = *reinterpret_cast<float4*>(ema + i);
                ev.x = decay * ev.x + (1.f - decay) * pv.x; ev.y = decay * ev.y + (1.f - decay) * pv.y;
                ev.z = decay * ev.z + (1.f - decay) * pv.z; ev.w = decay * ev.w + (1.f - decay) * pv.w;
                *reinterpret_cast<float4*>(ema + i) = ev;
            }
        } else {
            for (int64_t j = i; j < n; ++j) {
                float gr = g[j] * gscale + wd * p[j];
                m[j] = b1 * m[j] + (1.f - b1) * gr;
                v[j] = b2 * v[j] + (1.f - b2) * gr * gr;
                const float denom = sqrtf(v[j]) / bc2_sqrt + eps;
                p[j] -= (lr / bc1) * (m[j] / denom);
                if (ema) ema[j] = decay * ema[j] + (1.f - decay) * p[j];
            }
        }
    }
}

static int gn_check(const dmu_gn_params* p, const char* who, bool need_y, bool need_dx) {
    DMU_REQUIRE(p, "%s: null params", who);
    DMU_REQUIRE(p->x.ptr && p->sums && p->gamma && p->beta, "%s: null pointer", who);
    DMU_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->G > 0, "%s: non-positive dims", who);
    DMU_REQUIRE(p->C % p->G == 0, "%s: C=%d not divisible by G=%d", who, p->C, p->G);
    DMU_REQUIRE(p->C <= kMaxC, "%s: C=%d exceeds %d", who, p->C, kMaxC);
    DMU_REQUIRE(p->x.sc == 1, "%s: x must be NHWC (sc == 1)", who);
    const int vec = p->x.dtype == DMU_BF16 ? 8 : 4;
    DMU_REQUIRE(p->C % vec == 0 && p->x.sw % vec == 0 && p->x.sh % vec == 0 && p->x.sn % vec == 0, "%s: C and pitches must be multiples of %d", who, vec);
    DMU_REQUIRE(p->C / vec <= 256, "%s: C too large for one CTA row", who);
    if (need_y) {
        DMU_REQUIRE(p->y.ptr && p->y.sc == 1 && p->y.dtype == p->x.dtype, "%s: y must be NHWC of x's dtype", who);
        DMU_REQUIRE(p->y.sw % vec == 0 && p->y.sh % vec == 0 && p->y.sn % vec == 0, "%s: y pitches must be multiples of %d", who, vec);
    }
    if (need_dx) {
        DMU_REQUIRE(p->dx.ptr && p->dx.sc == 1 && p->dx.dtype == p->x.dtype && p->red, "%s: dx/red missing or wrong dtype", who);
        DMU_REQUIRE(p->dx.sw % vec == 0 && p->dx.sh % vec == 0 && p->dx.sn % vec == 0, "%s: dx pitches must be multiples of %d", who, vec);
        if (p->add0.ptr) DMU_REQUIRE(p->add0.sc == 1 && p->add0.dtype == p->x.dtype && p->add0.sw % vec == 0 && p->add0.sn % vec == 0, "%s: add0 layout", who);
        if (p->add1.ptr) DMU_REQUIRE(p->add1.sc == 1 && p->add1.dtype == p->x.dtype && p->add1.sw % vec == 0 && p->add1.sn % vec == 0, "%s: add1 layout", who);
    }
    return 0;
}

// pixel chunks per image so that the grid is about 4 waves of the SM count
static dim3 gn_grid(int N, int HW, int C, int vec) {
    const int lanes = 256 / (C / vec);
    int chunks = (HW + lanes * 4 - 1) / (lanes * 4);  // >= 4 pixels per lane
    const int want = (sm_count() * 8 + N - 1) / N;
    if (chunks > want) chunks = want;
    if (chunks < 1) chunks = 1;
    return dim3(chunks, N);
}

}  // namespace dmu

using namespace dmu;

// every kernel dispatched through this macro starts with pdl_trigger(); pdl_wait();
#define DISPATCH_T(dtype, KERNEL, grid, block, stream, ...)                                                               \
    do {                                                                                                                  \
        if ((dtype) == DMU_BF16) launch_pdl(KERNEL<__nv_bfloat16>, grid, dim3(block), 0, stream, dim3(1, 1, 1), __VA_ARGS__); \
        else launch_pdl(KERNEL<float>, grid, dim3(block), 0, stream, dim3(1, 1, 1), __VA_ARGS__);                            \
    } while (0)

extern "C" {

int dmu_gn_stats(const dmu_gn_params* p, dmu_stream_t stream) {
    if (int e = gn_check(p, "dmu_gn_stats", false, false)) return e;
    const int vec = p->x.dtype == DMU_BF16 ? 8 : 4;
    DISPATCH_T(p->x.dtype, gn_stats_kernel, gn_grid(p->N, p->H * p->W, p->C, vec), 256, as_stream(stream), *p);
    return check_launch("dmu_gn_stats");
}
int dmu_gn_coef(const dmu_gn_params* p, float* coef, dmu_stream_t stream) {
    if (int e = gn_check(p, "dmu_gn_coef", false, false)) return e;
    DMU_REQUIRE(coef, "dmu_gn_coef: null output");
    launch_pdl(gn_coef_kernel, dim3(p->N), dim3(256), 0, as_stream(stream), dim3(1, 1, 1), *p, coef);
    return check_launch("dmu_gn_coef");
}
int dmu_gn_apply(const dmu_gn_params* p, dmu_stream_t stream) {
    if (int e = gn_check(p, "dmu_gn_apply", true, false)) return e;
    const int vec = p->x.dtype == DMU_BF16 ? 8 : 4;
    DISPATCH_T(p->x.dtype, gn_apply_kernel, gn_grid(p->N, p->H * p->W, p->C, vec), 256, as_stream(stream), *p);
    return check_launch("dmu_gn_apply");
}
int dmu_gn_bwd_reduce(const dmu_gn_params* p, dmu_stream_t stream) {
    if (int e = gn_check(p, "dmu_gn_bwd_reduce", true, false)) return e;
    DMU_REQUIRE(p->red, "dmu_gn_bwd_reduce: red workspace missing");
    const int vec = p->x.dtype == DMU_BF16 ? 8 : 4;
    DISPATCH_T(p->x.dtype, gn_bwd_reduce_kernel, gn_grid(p->N, p->H * p->W, p->C, vec), 256, as_stream(stream), *p);
    return check_launch("dmu_gn_bwd_reduce");
}
int dmu_gn_bwd_apply(const dmu_gn_params* p, dmu_stream_t stream) {
    if (int e = gn_check(p, "dmu_gn_bwd_apply", true, true)) return e;
    const int vec = p->x.dtype == DMU_BF16 ? 8 : 4;
    DISPATCH_T(p->x.dtype, gn_bwd_apply_kernel, gn_grid(p->N, p->H * p->W, p->C, vec), 256, as_stream(stream), *p);
    return check_launch("dmu_gn_bwd_apply");
}

int dmu_gn_forward(const dmu_gn_params* p, dmu_stream_t stream) {
    if (int e = gn_check(p, "dmu_gn_forward", true, false)) return e;
    const int vec = p->x.dtype == DMU_BF16 ? 8 : 4;
    // measured on B200 (scripts/gn_time.py): the single-pass cluster kernel holds its slab in registers (2 CTAs per SM) and
    // wins while the tensor is small enough that launch count matters; from ~12 MB upwards the two streaming passes
    // (the second one served from L2 when the tensor fits) are faster (33.6 MB: 26.5 vs 36.9 us, 67 MB: 47 vs 98 us)
    const int64_t bytes = (int64_t)p->N * p->H * p->W * p->C * (p->x.dtype == DMU_BF16 ? 2 : 4);
    static const int fwd_fused_mb = [] { const char* e = getenv("DMU_GN_FWD_FUSED_MB"); return e ? atoi(e) : 12; }();
    const int cs = bytes <= ((int64_t)fwd_fused_mb << 20) ? gn_fused_cluster(p->H * p->W, p->C, vec) : 0;
    if (cs == 0) {
        // large tensors: shared-memory single pass when an image fits the shared memory of a cluster of <= 8 CTAs
        static const int fwd_smem = [] { const char* e = getenv("DMU_GN_FWD_SMEM"); return e ? atoi(e) : 1; }();     // A/B aid
        const int HW = p->H * p->W, V = p->C / vec;
        // (measured, scripts/gn_time.py: 16.8 MB 10.7 vs 14.3 us, 33.6 MB 22.5 vs 26.3 us; from 67 MB the two streaming passes win)
        if (fwd_smem && bytes <= (48ll << 20) && 256 / V >= 1) {
            for (int c2 = 1; c2 <= 8 && c2 <= HW; c2 <<= 1) {
                const int per = (HW + c2 - 1) / c2;
                const size_t smem = (size_t)per * V * 16 + (size_t)4 * p->C * 4;
                if (smem > 70 * 1024) continue;
                cudaError_t e;
                if (p->x.dtype == DMU_BF16) {
                    static bool a1 = (cudaFuncSetAttribute(gn_fwd_smem_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024), true);
                    (void)a1;
                    e = launch_pdl(gn_fwd_smem_kernel<__nv_bfloat16>, dim3(c2, p->N), dim3(256), smem, as_stream(stream), dim3(c2, 1, 1), *p, per);
                } else {
                    static bool a2 = (cudaFuncSetAttribute(gn_fwd_smem_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024), true);
                    (void)a2;
                    e = launch_pdl(gn_fwd_smem_kernel<float>, dim3(c2, p->N), dim3(256), smem, as_stream(stream), dim3(c2, 1, 1), *p, per);
                }
                if (e != cudaSuccess) return fail("dmu_gn_forward: cluster launch failed: %s", cudaGetErrorString(e));
                return check_launch("dmu_gn_forward");
            }
        }
        if (int e = dmu_gn_stats(p, stream)) return e;
        return dmu_gn_apply(p, stream);
    }
    if (p->x.dtype == DMU_BF16) return launch_cluster(gn_fwd_fused_kernel<__nv_bfloat16>, cs, p->N, as_stream(stream), *p);
    return launch_cluster(gn_fwd_fused_kernel<float>, cs, p->N, as_stream(stream), *p);
}
int dmu_gn_param_grads(const void* table_device, int32_t n_desc, int32_t max_c, int32_t N, dmu_stream_t stream) {
    DMU_REQUIRE(table_device && n_desc > 0 && max_c > 0 && N > 0, "dmu_gn_param_grads: bad arguments");
    gn_param_grads_kernel<<<dim3((max_c * 2 + 255) / 256, n_desc), 256, 0, as_stream(stream)>>>(reinterpret_cast<const GnPgDesc*>(table_device), N);
    return check_launch("dmu_gn_param_grads");
}

int dmu_gn_backward(const dmu_gn_params* p, dmu_stream_t stream) {
    if (int e = gn_check(p, "dmu_gn_backward", true, true)) return e;
    const int vec = p->x.dtype == DMU_BF16 ? 8 : 4;
    // measured on B200: holding (x, dy) in registers costs the single-pass backward its occupancy (225 registers), so it
    // only pays for images one CTA can hold, where it saves a launch; larger images take the two-pass kernels
    const int64_t bytes = (int64_t)p->N * p->H * p->W * p->C * (p->x.dtype == DMU_BF16 ? 2 : 4);
    const int64_t img_bytes = bytes / p->N;
    // measured (scripts/gn_time.py, round 2: slab kernel with the whole (x, dy) slab in flight by cp.async, du kept in the slab,
    // one-MUFU sigmoid): shared-memory single pass for images of >= 32 KB (32x32x64 x 128: 23 us against 32 us for the two-pass
    // pair and 46 us for the round-1 slab kernel; 16x16x128: 11 vs 17 us), the two-pass pair below that
    // measured inside the backward graph: the register-held single pass (221 registers, one CTA per SM) is the faster kernel
    // alone for small tensors but costs the step 2.4 % (41.6k vs 40.6k img/s with it off): its CTAs leave no room for the
    // weight-gradient lane, the two-pass pair does.  DMU_GN_BWD_FUSED_MB > 0 re-enables it up to that size.
    static const int bwd_fused_mb = [] { const char* e = getenv("DMU_GN_BWD_FUSED_MB"); return e ? atoi(e) : 0; }();
    const int cs = (bytes <= ((int64_t)bwd_fused_mb << 20) && gn_fused_cluster(p->H * p->W, p->C, vec) == 1) ? 1 : 0;
    static const int smem_min_kb = [] { const char* e = getenv("DMU_GN_BWD_SMEM_MIN_KB"); return e ? atoi(e) : 32; }();
    if (cs == 0 && img_bytes < ((int64_t)smem_min_kb << 10)) {
        if (int e = dmu_gn_bwd_reduce(p, stream)) return e;
        return dmu_gn_bwd_apply(p, stream);
    }
    if (cs == 0) {
        // larger images: the (x, dy) slab of each CTA of a cluster lives in shared memory (single pass)
        const int HW = p->H * p->W, V = p->C / vec;
        static const bool smem_path = !(getenv("DMU_GN_SMEM") && getenv("DMU_GN_SMEM")[0] == '0');     // A/B aid
        if (smem_path && 256 / V >= 1) {
            for (int c2 = 1; c2 <= 8 && c2 <= HW; c2 <<= 1) {
                const int per = (HW + c2 - 1) / c2;
                const size_t smem = (size_t)2 * per * V * 16 + (size_t)8 * p->C * 4;
                static const int bwd_cap_kb = [] { const char* e = getenv("DMU_GN_BWD_SMEM_KB"); return e ? atoi(e) : 72; }();
                // 72 KB (a 4-CTA cluster per 32x32x64 image, 3 CTAs per SM): slower than 8 smaller CTAs when timed alone (48.7 vs 37 us)
                // but faster inside the backward graph next to the weight-gradient lane (step 3.19 -> 3.14 ms)
                if (smem > (size_t)bwd_cap_kb * 1024) continue;
                cudaError_t e;
                if (p->x.dtype == DMU_BF16) {
                    static bool a1 = (cudaFuncSetAttribute(gn_bwd_smem_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024), true);
                    (void)a1;
                    e = launch_pdl(gn_bwd_smem_kernel<__nv_bfloat16>, dim3(c2, p->N), dim3(256), smem, as_stream(stream), dim3(c2, 1, 1), *p, per);
                } else {
                    static bool a2 = (cudaFuncSetAttribute(gn_bwd_smem_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024), true);
                    (void)a2;
                    e = launch_pdl(gn_bwd_smem_kernel<float>, dim3(c2, p->N), dim3(256), smem, as_stream(stream), dim3(c2, 1, 1), *p, per);
                }
                if (e != cudaSuccess) return fail("dmu_gn_backward: cluster launch failed: %s", cudaGetErrorString(e));
                return check_launch("dmu_gn_backward");
            }
        }
        if (int e = dmu_gn_bwd_reduce(p, stream)) return e;
        return dmu_gn_bwd_apply(p, stream);
    }
    if (p->x.dtype == DMU_BF16) return launch_cluster(gn_bwd_fused_kernel<__nv_bfloat16>, cs, p->N, as_stream(stream), *p);
    return launch_cluster(gn_bwd_fused_kernel<float>, cs, p->N, as_stream(stream), *p);
}

int dmu_colsum(const dmu_tensor4* x, int32_t N, int32_t H, int32_t W, int32_t C, float* out_nc, int64_t pitch, float* out_c,
               float scale, dmu_stream_t stream) {
    DMU_REQUIRE(x && x->ptr, "dmu_colsum: null input");
    DMU_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= kMaxC, "dmu_colsum: bad dims");
    const int vec = x->dtype == DMU_BF16 ? 8 : 4;
    DMU_REQUIRE(x->sc == 1 && C % vec == 0 && x->sw % vec == 0 && x->sn % vec == 0 && C / vec <= 256, "dmu_colsum: needs NHWC, C multiple of %d", vec);
    dmu_tensor4 t = *x;
    int Nn = N, Hh = H, Ww = W;
    if (!out_nc && t.sh == (int64_t)W * t.sw && t.sn == (int64_t)H * t.sh) {   // only the total is wanted: one flat pixel range
        Ww = N * H * W; Hh = 1; Nn = 1;
        t.sh = t.sn = (int64_t)Ww * t.sw;
    }
    dim3 grid = gn_grid(Nn, Hh * Ww, C, vec);
    if (out_c) {   // every CTA ends in C atomics on the same C addresses: keep the CTA count near one per SM
        const int cap = (sm_count() + Nn - 1) / Nn;
        if ((int)grid.x > cap) grid.x = cap;
    }
    DISPATCH_T(t.dtype, colsum_kernel, grid, 256, as_stream(stream), t, Hh, Ww, C, out_nc, pitch, out_c, scale);
    return check_launch("dmu_colsum");
}

int dmu_colsum_multi(const dmu_colsum_desc* table_device, int32_t n_desc, int32_t total_ctas, int32_t dtype, dmu_stream_t stream) {
    DMU_REQUIRE(table_device && n_desc > 0 && total_ctas > 0, "dmu_colsum_multi: bad arguments");
    const dim3 grid(total_ctas);
    if (dtype == DMU_BF16) launch_pdl(colsum_multi_kernel<__nv_bfloat16>, grid, dim3(256), 0, as_stream(stream), dim3(1, 1, 1), table_device, (int)n_desc);
    else launch_pdl(colsum_multi_kernel<float>, grid, dim3(256), 0, as_stream(stream), dim3(1, 1, 1), table_device, (int)n_desc);
    return check_launch("dmu_colsum_multi");
}

int dmu_silu_pool_fwd(const dmu_tensor4* x, int32_t N, int32_t H, int32_t W, int32_t C, float* out, int64_t pitch, float scale, dmu_stream_t stream) {
    DMU_REQUIRE(x && x->ptr && out && N > 0 && H > 0 && W > 0 && C > 0 && C <= kMaxC, "dmu_silu_pool_fwd: bad arguments");
    const int vec = x->dtype == DMU_BF16 ? 8 : 4;
    DMU_REQUIRE(x->sc == 1 && C % vec == 0 && x->sw % vec == 0 && x->sn % vec == 0 && C / vec <= 256, "dmu_silu_pool_fwd: needs NHWC, C multiple of %d", vec);
    dim3 grid = gn_grid(N, H * W, C, vec);
    const int cap = (2 * sm_count() + N - 1) / N;
    if ((int)grid.x > cap) grid.x = cap;
    DISPATCH_T(x->dtype, silu_pool_fwd_kernel, grid, 256, as_stream(stream), *x, H, W, C, out, pitch, scale);
    return check_launch("dmu_silu_pool_fwd");
}
int dmu_silu_pool_bwd(const dmu_tensor4* x, const dmu_tensor4* dx, int32_t N, int32_t H, int32_t W, int32_t C, const float* g, int64_t pitch, float scale,
                      dmu_stream_t stream) {
    DMU_REQUIRE(x && x->ptr && dx && dx->ptr && g && N > 0 && H > 0 && W > 0 && C > 0 && C <= kMaxC, "dmu_silu_pool_bwd: bad arguments");
    const int vec = x->dtype == DMU_BF16 ? 8 : 4;
    DMU_REQUIRE(x->sc == 1 && dx->sc == 1 && dx->dtype == x->dtype && C % vec == 0 && x->sw % vec == 0 && x->sn % vec == 0 && dx->sw % vec == 0 &&
                    dx->sn % vec == 0 && C / vec <= 256, "dmu_silu_pool_bwd: needs NHWC tensors of one dtype, C multiple of %d", vec);
    DISPATCH_T(x->dtype, silu_pool_bwd_kernel, gn_grid(N, H * W, C, vec), 256, as_stream(stream), *x, *dx, H, W, C, g, pitch, scale);
    return check_launch("dmu_silu_pool_bwd");
}

int dmu_gn_bwd_bwd(const dmu_gn_bwd2_params* p, dmu_stream_t stream) {
    DMU_REQUIRE(p && p->x.ptr && p->dy.ptr && p->c.ptr && p->gx.ptr && p->sums && p->gamma && p->beta, "dmu_gn_bwd_bwd: null pointer");
    DMU_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->G > 0 && p->C % p->G == 0 && p->C <= kMaxC, "dmu_gn_bwd_bwd: bad dims");
    const int vec = p->x.dtype == DMU_BF16 ? 8 : 4;
    const dmu_tensor4* ts[5] = {&p->x, &p->dy, &p->c, &p->gx, &p->gdy};
    for (int i = 0; i < 5; ++i) {
        if (!ts[i]->ptr) continue;
        DMU_REQUIRE(ts[i]->sc == 1 && ts[i]->dtype == p->x.dtype && ts[i]->sw % vec == 0 && ts[i]->sh % vec == 0 && ts[i]->sn % vec == 0,
                    "dmu_gn_bwd_bwd: tensors must be NHWC of one dtype with pitches multiple of %d", vec);
    }
    DMU_REQUIRE(p->C % vec == 0 && p->C / vec <= 256 && 5 * p->C <= 256 * vec, "dmu_gn_bwd_bwd: unsupported channel count %d", p->C);
    if (p->x.dtype == DMU_BF16) gn_bwd_bwd_kernel<__nv_bfloat16><<<p->N, 256, 0, as_stream(stream)>>>(*p);
    else gn_bwd_bwd_kernel<float><<<p->N, 256, 0, as_stream(stream)>>>(*p);
    return check_launch("dmu_gn_bwd_bwd");
}
int dmu_silu_pool_bwd_bwd(const dmu_tensor4* x, const dmu_tensor4* c, const dmu_tensor4* gx, int32_t N, int32_t H, int32_t W, int32_t C, const float* g,
                          int64_t pitch, float* gg, int64_t gg_pitch, float scale, dmu_stream_t stream) {
    DMU_REQUIRE(x && x->ptr && c && c->ptr && gx && gx->ptr && g && gg && N > 0 && H > 0 && W > 0 && C > 0 && C <= kMaxC, "dmu_silu_pool_bwd_bwd: bad arguments");
    const int vec = x->dtype == DMU_BF16 ? 8 : 4;
    DMU_REQUIRE(x->sc == 1 && c->sc == 1 && gx->sc == 1 && c->dtype == x->dtype && gx->dtype == x->dtype && C % vec == 0 && C / vec <= 256,
                "dmu_silu_pool_bwd_bwd: needs NHWC tensors of one dtype, C multiple of %d", vec);
    dim3 grid = gn_grid(N, H * W, C, vec);
    const int cap = (2 * sm_count() + N - 1) / N;
    if ((int)grid.x > cap) grid.x = cap;
    if (x->dtype == DMU_BF16) silu_pool_bwd_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(*x, *c, *gx, H, W, C, g, pitch, gg, gg_pitch, scale);
    else silu_pool_bwd_bwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(*x, *c, *gx, H, W, C, g, pitch, gg, gg_pitch, scale);
    return check_launch("dmu_silu_pool_bwd_bwd");
}

int dmu_act_fwd(const float* x, float* y, int64_t n, int32_t kind, dmu_stream_t stream) {
    DMU_REQUIRE(x && y && n >= 0 && kind >= 0 && kind <= 2, "dmu_act_fwd: bad arguments");
    if (n == 0) return 0;
    int grid = (int)((n + 255) / 256); if (grid > sm_count() * 8) grid = sm_count() * 8;
    act_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, n, kind);
    return check_launch("dmu_act_fwd");
}
int dmu_act_bwd(const float* x, const float* dy, float* dx, int64_t n, int32_t kind, dmu_stream_t stream) {
    DMU_REQUIRE(x && dy && dx && n >= 0 && kind >= 0 && kind <= 2, "dmu_act_bwd: bad arguments");
    if (n == 0) return 0;
    int grid = (int)((n + 255) / 256); if (grid > sm_count() * 8) grid = sm_count() * 8;
    act_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, dy, dx, n, kind);
    return check_launch("dmu_act_bwd");
}

int dmu_sinusoidal_embedding(const void* t, int32_t t_is_float, float* emb, int64_t batch, int32_t dim, dmu_stream_t stream) {
    DMU_REQUIRE(t && emb && batch >= 0, "dmu_sinusoidal_embedding: bad arguments");
    DMU_REQUIRE(dim >= 4 && dim % 2 == 0, "dmu_sinusoidal_embedding: dim must be even and >= 4 (embeddings.py:34 divides by half-1)");
    if (batch == 0) return 0;
    const int half = dim / 2;
    const float neg_k = (float)(-(log(10000.0) / (double)(half - 1)));
    int grid = (int)((batch * half + 255) / 256); if (grid > sm_count() * 8) grid = sm_count() * 8;
    sinusoidal_kernel<<<grid, 256, 0, as_stream(stream)>>>(t, t_is_float, emb, batch, dim, neg_k);
    return check_launch("dmu_sinusoidal_embedding");
}

int dmu_repack_weights(const dmu_repack_desc* descs_device, int32_t n_desc, int64_t max_numel, dmu_stream_t stream) {
    DMU_REQUIRE(descs_device && n_desc > 0 && max_numel > 0, "dmu_repack_weights: bad arguments");
    static const int gx_cap = [] { const char* e = getenv("DMU_REPACK_GX"); return e ? atoi(e) : 64; }();
    int gx = (int)((max_numel + 2047) / 2048); if (gx > gx_cap) gx = gx_cap;
    repack_kernel<<<dim3(gx, n_desc), 256, 0, as_stream(stream)>>>(descs_device);
    return check_launch("dmu_repack_weights");
}

int dmu_copy4(const dmu_tensor4* src, const dmu_tensor4* dst, int32_t N, int32_t H, int32_t W, int32_t C, dmu_stream_t stream) {
    DMU_REQUIRE(src && dst && src->ptr && dst->ptr, "dmu_copy4: null tensor");
    DMU_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "dmu_copy4: bad dims");
    const int64_t total = (int64_t)N * H * W * C;
    int grid = (int)((total + 255) / 256); if (grid > sm_count() * 8) grid = sm_count() * 8;
    copy4_kernel<<<grid, 256, 0, as_stream(stream)>>>(*src, *dst, N, H, W, C);
    return check_launch("dmu_copy4");
}

int dmu_adam_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, int64_t step, float ema_decay, float grad_scale, const int64_t* step_device,
                 dmu_stream_t stream) {
    DMU_REQUIRE(p && g && m && v && n >= 0 && (step >= 1 || step_device), "dmu_adam_ema: bad arguments");
    DMU_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)ema) & 15) == 0, "dmu_adam_ema: arenas (and range starts) must be 16-byte aligned");
    if (step < 1) step = 1;
    if (n == 0) return 0;
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    int grid = (int)((n / 4 + 255) / 256); if (grid > sm_count() * 8) grid = sm_count() * 8; if (grid < 1) grid = 1;
    adam_ema_kernel<<<grid, 256, 0, as_stream(stream)>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, ema_decay, grad_scale,
                                                         step_device);
    return check_launch("dmu_adam_ema");
}

}  // extern "C"
