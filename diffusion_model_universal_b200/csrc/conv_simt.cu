// Implicit-GEMM convolution on CUDA cores (fp32 FMA, fp32 accumulate).
//
// This is the exact-arithmetic path (precision = fp32) and the shape-generic
// path for the layers the tcgen05 kernels do not take (3-channel stem/head,
// Linear layers with tiny M).  One gather-GEMM covers Conv2d / ConvTranspose2d
// / Linear fprop and dgrad through the weight strides (see dmu_conv_params).
#include <stdlib.h>
#include "common.cuh"

namespace dmu {

constexpr int BM = 128, BN = 64, BK = 16;

struct PixelCoord { int n, ho, wo; bool valid; };

__device__ __forceinline__ bool gather_coord(const dmu_conv_params& P, int ho, int wo, int r, int s, int& hi, int& wi) {
    if (P.gather == 0) {
        hi = ho * P.stride - P.pad + r;
        wi = wo * P.stride - P.pad + s;
        return hi >= 0 && hi < P.Hi && wi >= 0 && wi < P.Wi;
    }
    const int hn = ho + P.pad - r, wn = wo + P.pad - s;
    if (hn < 0 || wn < 0 || (hn % P.stride) != 0 || (wn % P.stride) != 0) return false;
    hi = hn / P.stride;
    wi = wn / P.stride;
    return hi < P.Hi && wi < P.Wi;
}

// kFast: x is channel-contiguous, Ck % 8 == 0 -> an 8-run of k stays inside one tap.
template <bool kFast>
__global__ void __launch_bounds__(256) conv_gemm_kernel(dmu_conv_params P) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int M = P.N * P.Ho * P.Wo;
    const int K = P.R * P.S * P.Ck;
    const int m_base = blockIdx.x * BM;
    const int j_base = blockIdx.y * BN;

    // A loader mapping: one output pixel, 8 consecutive k
    const int a_m = tid >> 1, a_k0 = (tid & 1) * 8;
    PixelCoord pc;
    {
        const int m = m_base + a_m;
        pc.valid = m < M;
        const int mm = pc.valid ? m : 0;
        pc.wo = mm % P.Wo;
        pc.ho = (mm / P.Wo) % P.Ho;
        pc.n = mm / (P.Wo * P.Ho);
    }
    // B loader mapping
    const bool w_kcontig = (P.w_sk == 1);
    const int b_j = w_kcontig ? (tid >> 2) : ((tid & 15) * 4);
    const int b_k0 = w_kcontig ? ((tid & 3) * 4) : (tid >> 4);

    float a_reg[8], b_reg[4];

    auto load_a = [&](int kt) {
        const int k0 = kt * BK + a_k0;
#pragma unroll
        for (int i = 0; i < 8; ++i) a_reg[i] = 0.f;
        if (!pc.valid) return;
        if (kFast) {
            if (k0 >= K) return;
            const int tap = k0 / P.Ck, kc = k0 % P.Ck;
            int hi, wi;
            if (!gather_coord(P, pc.ho, pc.wo, tap / P.S, tap % P.S, hi, wi)) return;
            const int64_t off = (int64_t)pc.n * P.x.sn + (int64_t)hi * P.x.sh + (int64_t)wi * P.x.sw + kc;
            if (P.x.dtype == DMU_BF16) {
                load_vec<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(P.x.ptr) + off, a_reg);
            } else {
                load_vec<float>(reinterpret_cast<const float*>(P.x.ptr) + off, a_reg);
                load_vec<float>(reinterpret_cast<const float*>(P.x.ptr) + off + 4, a_reg + 4);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = k0 + i;
                if (k < K) {
                    const int tap = k / P.Ck, kc = k % P.Ck;
                    int hi, wi;
                    if (gather_coord(P, pc.ho, pc.wo, tap / P.S, tap % P.S, hi, wi))
                        a_reg[i] = ld_as_float(P.x.ptr, (int64_t)pc.n * P.x.sn + (int64_t)hi * P.x.sh + (int64_t)wi * P.x.sw + (int64_t)kc * P.x.sc, P.x.dtype);
                }
            }
        }
    };
    auto load_b = [&](int kt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = kt * BK + b_k0 + (w_kcontig ? i : 0);
            const int j = j_base + b_j + (w_kcontig ? 0 : i);
            float v = 0.f;
            if (k < K && j < P.Cj) {
                const int tap = k / P.Ck, kc = k % P.Ck;
                v = ld_as_float(P.w, (int64_t)j * P.w_sn + (int64_t)kc * P.w_sk + (int64_t)tap * P.w_st, P.w_dtype);
            }
            b_reg[i] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][a_k0 + i][a_m] = a_reg[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (w_kcontig) Bs[buf][b_k0 + i][b_j] = b_reg[i];
            else Bs[buf][b_k0][b_j + i] = b_reg[i];
        }
    };

    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int KT = (K + BK - 1) / BK;
    load_a(0); load_b(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) { load_a(kt + 1); load_b(kt + 1); }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < KT) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue: + bias + temb + residual, strided store
    const int j0 = j_base + tx * 4;
    if (j0 >= P.Cj) return;
    float bj[4] = {0.f, 0.f, 0.f, 0.f};
    if (P.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j0 + j < P.Cj) bj[j] = P.bias[j0 + j];
    }
    const bool vec_out = (P.y.sc == 1) && (j0 + 3 < P.Cj) && (P.Cj % 4 == 0) && (P.y.sw % 4 == 0) && (P.y.sh % 4 == 0) && (P.y.sn % 4 == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m_base + ty * 8 + i;
        if (m >= M) continue;
        const int wo = m % P.Wo, ho = (m / P.Wo) % P.Ho, n = m / (P.Wo * P.Ho);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bj[j];
        if (P.temb) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j0 + j < P.Cj) v[j] += P.temb[(int64_t)n * P.temb_pitch + j0 + j];
        }
        if (P.res.ptr) {
            const int64_t ro = (int64_t)n * P.res.sn + (int64_t)ho * P.res.sh + (int64_t)wo * P.res.sw;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j0 + j < P.Cj) v[j] += ld_as_float(P.res.ptr, ro + (int64_t)(j0 + j) * P.res.sc, P.res.dtype);
        }
        const int64_t yo = (int64_t)n * P.y.sn + (int64_t)ho * P.y.sh + (int64_t)wo * P.y.sw;
        if (vec_out) {
            if (P.y.dtype == DMU_BF16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(P.y.ptr) + yo + j0) = pk;
            } else {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(P.y.ptr) + yo + j0) = make_float4(v[0], v[1], v[2], v[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j0 + j < P.Cj) st_from_float(P.y.ptr, yo + (int64_t)(j0 + j) * P.y.sc, P.y.dtype, v[j]);
        }
    }
}

// Few output channels (Cj <= 4: the 64->3 head conv, Linear(.,1)): one thread per
// output pixel, weights broadcast from shared memory.
constexpr int kSmallMaxK = 4608;
__global__ void __launch_bounds__(128) conv_small_n_kernel(dmu_conv_params P) {
    extern __shared__ float s_w[];  // [K][4]
    const int K = P.R * P.S * P.Ck;
    for (int i = threadIdx.x; i < K * 4; i += blockDim.x) {
        const int k = i >> 2, j = i & 3;
        float v = 0.f;
        if (j < P.Cj) {
            const int tap = k / P.Ck, kc = k % P.Ck;
            v = ld_as_float(P.w, (int64_t)j * P.w_sn + (int64_t)kc * P.w_sk + (int64_t)tap * P.w_st, P.w_dtype);
        }
        s_w[i] = v;
    }
    __syncthreads();
    const int M = P.N * P.Ho * P.Wo;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int wo = m % P.Wo, ho = (m / P.Wo) % P.Ho, n = m / (P.Wo * P.Ho);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = (P.x.sc == 1) && (P.Ck % 8 == 0);
    for (int tap = 0; tap < P.R * P.S; ++tap) {
        int hi, wi;
        if (!gather_coord(P, ho, wo, tap / P.S, tap % P.S, hi, wi)) continue;
        const int64_t off = (int64_t)n * P.x.sn + (int64_t)hi * P.x.sh + (int64_t)wi * P.x.sw;
        const float* wrow = s_w + (int64_t)tap * P.Ck * 4;
        if (vec) {
            for (int kc = 0; kc < P.Ck; kc += 8) {
                float xv[8];
                if (P.x.dtype == DMU_BF16) load_vec<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(P.x.ptr) + off + kc, xv);
                else { load_vec<float>(reinterpret_cast<const float*>(P.x.ptr) + off + kc, xv); load_vec<float>(reinterpret_cast<const float*>(P.x.ptr) + off + kc + 4, xv + 4); }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wrow + (kc + i) * 4);
                    acc[0] = fmaf(xv[i], w4.x, acc[0]); acc[1] = fmaf(xv[i], w4.y, acc[1]);
                    acc[2] = fmaf(xv[i], w4.z, acc[2]); acc[3] = fmaf(xv[i], w4.w, acc[3]);
                }
            }
        } else {
            for (int kc = 0; kc < P.Ck; ++kc) {
                const float xv = ld_as_float(P.x.ptr, off + (int64_t)kc * P.x.sc, P.x.dtype);
                const float4 w4 = *reinterpret_cast<const float4*>(wrow + kc * 4);
                acc[0] = fmaf(xv, w4.x, acc[0]); acc[1] = fmaf(xv, w4.y, acc[1]);
                acc[2] = fmaf(xv, w4.z, acc[2]); acc[3] = fmaf(xv, w4.w, acc[3]);
            }
        }
    }
    const int64_t yo = (int64_t)n * P.y.sn + (int64_t)ho * P.y.sh + (int64_t)wo * P.y.sw;
    for (int j = 0; j < P.Cj; ++j) {
        float v = acc[j];
        if (P.bias) v += P.bias[j];
        if (P.temb) v += P.temb[(int64_t)n * P.temb_pitch + j];
        if (P.res.ptr) v += ld_as_float(P.res.ptr, (int64_t)n * P.res.sn + (int64_t)ho * P.res.sh + (int64_t)wo * P.res.sw + (int64_t)j * P.res.sc, P.res.dtype);
        st_from_float(P.y.ptr, yo + (int64_t)j * P.y.sc, P.y.dtype, v);
    }
}


// Skinny GEMM with split-K for the Linear layers whose M (batch rows) is too small to fill the chip with 128x64 tiles
// (time-embedding MLP, the 22-way time projection and their input gradients: M = batch, K up to 3328).
//   y[m, j] (+)= bias[j] + res[m, j] + sum_k x[m, k] * w[j*w_sn + k*w_sk]      fp32 output, pre-zeroed, fp32 atomics
constexpr int LM = 32, LN = 64, LK = 32;
// ordered = 1: the splits of one output tile form a thread-block cluster along z; every CTA parks its partial tile in shared memory
// and CTA 0 adds them in the order z = 1, 2, ... through DSMEM and stores the tile - no atomics, no zeroing, and the same bits
// on every run (the forward's Linear layers: fp32 atomics would make a bf16 forward differ from run to run by ~1e-2).
__global__ void __launch_bounds__(128) linear_splitk_kernel(dmu_conv_params P, int k_per_split, int ordered) {
    __shared__ float As[LK][LM + 1];
    __shared__ __align__(16) float Bs[LK][LN + 4];
    const int tid = threadIdx.x;
    const int M = P.N, K = P.Ck;
    const int m_base = blockIdx.x * LM, j_base = blockIdx.y * LN;
    const int k_lo = blockIdx.z * k_per_split, k_hi = min(K, k_lo + k_per_split);
    const int ty = tid >> 4, tx = tid & 15;   // 8 x 16 threads, 4 rows x 4 cols each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool w_kcontig = P.w_sk == 1;
    // fp32 operands with 16-byte aligned rows: every thread issues its 2 + 4 float4 loads of a k-step back to back (the
    // element-wise loader below is one branch + one dependent load per element: ~30 us for the 128 x 3136 x 256 products of
    // the time-embedding backward, which sit alone at the tail of the step)
    const bool vec = P.x.dtype == DMU_F32 && P.w_dtype == DMU_F32 && P.x.sc == 1 && P.x.sn % 4 == 0 && (k_lo % 4) == 0 &&
                     (reinterpret_cast<uintptr_t>(P.x.ptr) & 15) == 0 && (reinterpret_cast<uintptr_t>(P.w) & 15) == 0 &&
                     (w_kcontig ? P.w_sn % 4 == 0 : (P.w_sn == 1 && P.w_sk % 4 == 0 && j_base + LN <= P.Cj));
    for (int k0 = k_lo; k0 < k_hi; k0 += LK) {
        if (vec && k0 + LK <= k_hi) {
            const float* xp = reinterpret_cast<const float*>(P.x.ptr);
            const float* wp = reinterpret_cast<const float*>(P.w);
            float4 ra[2], rb[4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i4 = tid + 128 * u, mm = i4 >> 3, k4 = i4 & 7;
                const int m = m_base + mm;
                ra[u] = m < M ? __ldg(reinterpret_cast<const float4*>(xp + (int64_t)m * P.x.sn + k0 + 4 * k4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i4 = tid + 128 * u;
                if (w_kcontig) {
                    const int jj = i4 >> 3, k4 = i4 & 7, j = j_base + jj;
                    rb[u] = j < P.Cj ? __ldg(reinterpret_cast<const float4*>(wp + (int64_t)j * P.w_sn + k0 + 4 * k4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    const int kk = i4 >> 4, j4 = i4 & 15;
                    rb[u] = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)(k0 + kk) * P.w_sk + j_base + 4 * j4));
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i4 = tid + 128 * u, mm = i4 >> 3, k4 = i4 & 7;
                As[4 * k4 + 0][mm] = ra[u].x; As[4 * k4 + 1][mm] = ra[u].y; As[4 * k4 + 2][mm] = ra[u].z; As[4 * k4 + 3][mm] = ra[u].w;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i4 = tid + 128 * u;
                if (w_kcontig) {
                    const int jj = i4 >> 3, k4 = i4 & 7;
                    Bs[4 * k4 + 0][jj] = rb[u].x; Bs[4 * k4 + 1][jj] = rb[u].y; Bs[4 * k4 + 2][jj] = rb[u].z; Bs[4 * k4 + 3][jj] = rb[u].w;
                } else {
                    const int kk = i4 >> 4, j4 = i4 & 15;
                    *reinterpret_cast<float4*>(&Bs[kk][4 * j4]) = rb[u];
                }
            }
        } else {
        // A tile: LM rows x LK k (k fastest across threads: x rows are k-contiguous)
        for (int i = tid; i < LM * LK; i += 128) {
            const int kk = i % LK, mm = i / LK;
            const int m = m_base + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < k_hi) ? ld_as_float(P.x.ptr, (int64_t)m * P.x.sn + (int64_t)k * P.x.sc, P.x.dtype) : 0.f;
        }
        // B tile: LK k x LN j, thread order along whichever axis of w is contiguous
        for (int i = tid; i < LK * LN; i += 128) {
            const int kk = w_kcontig ? i % LK : i / LN, jj = w_kcontig ? i / LK : i % LN;
            const int j = j_base + jj, k = k0 + kk;
            Bs[kk][jj] = (j < P.Cj && k < k_hi) ? ld_as_float(P.w, (int64_t)j * P.w_sn + (int64_t)k * P.w_sk, P.w_dtype) : 0.f;
        }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < LK; ++kk) {
            float a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (ordered && gridDim.z > 1) {
        float* part = &Bs[0][0];      // LM x LN floats fit the filter tile's LK x (LN + 4) (the k-loop is over: the last __syncthreads passed)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) part[(ty * 4 + i) * LN + tx * 4 + j] = acc[i][j];
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (blockIdx.z == 0) {
            for (unsigned z = 1; z < gridDim.z; ++z) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t ra;
                        float v;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((uint32_t)__cvta_generic_to_shared(&part[(ty * 4 + i) * LN + tx * 4 + j])), "r"(z));
                        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
                        acc[i][j] += v;
                    }
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");      // nobody leaves while CTA 0 still reads
        if (blockIdx.z != 0) return;
    }
    const bool plain = gridDim.z == 1 || ordered;
    float* y = reinterpret_cast<float*>(P.y.ptr);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m_base + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jc = j_base + tx * 4 + j;
            if (jc >= P.Cj) continue;
            float v = acc[i][j];
            if (blockIdx.z == 0) {
                if (P.bias) v += P.bias[jc];
                if (P.res.ptr) v += ld_as_float(P.res.ptr, (int64_t)m * P.res.sn + (int64_t)jc * P.res.sc, P.res.dtype);
            }
            if (plain) y[(int64_t)m * P.y.sn + jc] = v;
            else atomicAdd(&y[(int64_t)m * P.y.sn + jc], v);
        }
    }
}

// ------------------------------------------------------------------ wgrad
// out[a][rs][b] += sum_pixels P[pix, a] * Q[gather(pix, rs), b]; split over pixel ranges (blockIdx.z).
constexpr int WK = 16;  // pixels per step
template <int TM>       // rows (a) per thread; BMw = 16*TM, BNw = 64
__global__ void __launch_bounds__(256) wgrad_kernel(dmu_wgrad_params P, int pix_per_split) {
    constexpr int BMw = 16 * TM, BNw = 64;
    __shared__ __align__(16) float Ps[WK][BMw + 4];
    __shared__ __align__(16) float Qs[WK][BNw + 4];
    const int tid = threadIdx.x;
    const int Mtot = P.N * P.Hp * P.Wp;
    const int ncols = P.R * P.S * P.Cb;
    const int a_base = blockIdx.y * BMw;
    const int c_base = blockIdx.x * BNw;
    const int pix0 = blockIdx.z * pix_per_split;
    const int pix1 = min(Mtot, pix0 + pix_per_split);
    const int ty = tid >> 4, tx = tid & 15;
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bias_acc[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) bias_acc[i] = 0.f;
    const bool do_bias = P.dbias != nullptr && blockIdx.x == 0 && tx == 0;

    // Q loader: thread -> pixel (tid>>4), 4 consecutive cols (tid&15)*4
    const int q_k = tid >> 4, q_c = (tid & 15) * 4;
    const bool q_fast = (P.q.sc == 1) && (P.Cb % 4 == 0);

    for (int p0 = pix0; p0 < pix1; p0 += WK) {
        // ---- load P tile [WK][BMw]
        if (TM == 4 && P.p.sc == 1 && P.p.dtype == DMU_F32 && a_base + BMw <= P.Ca && P.p.sw % 4 == 0 && P.p.sh % 4 == 0 && P.p.sn % 4 == 0 &&
            (reinterpret_cast<uintptr_t>(P.p.ptr) & 15) == 0) {
            // one float4 per thread: pixel tid >> 4, channels (tid & 15) * 4 (the element-wise loader below costs three integer
            // divisions, a dtype branch and a dependent load per element)
            const int k = tid >> 4, a4 = (tid & 15) * 4;
            const int pix = p0 + k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pix < pix1) {
                const int wo = pix % P.Wp, ho = (pix / P.Wp) % P.Hp, n = pix / (P.Wp * P.Hp);
                v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(P.p.ptr) + (int64_t)n * P.p.sn + (int64_t)ho * P.p.sh +
                                                          (int64_t)wo * P.p.sw + a_base + a4));
            }
            *reinterpret_cast<float4*>(&Ps[k][a4]) = v;
        } else
        for (int i = tid; i < WK * BMw; i += 256) {
            const int k = i / BMw, a = i % BMw;
            const int pix = p0 + k;
            float v = 0.f;
            if (pix < pix1 && a_base + a < P.Ca) {
                const int wo = pix % P.Wp, ho = (pix / P.Wp) % P.Hp, n = pix / (P.Wp * P.Hp);
                v = ld_as_float(P.p.ptr, (int64_t)n * P.p.sn + (int64_t)ho * P.p.sh + (int64_t)wo * P.p.sw + (int64_t)(a_base + a) * P.p.sc, P.p.dtype);
            }
            Ps[k][a] = v;
        }
        // ---- load Q tile [WK][64]
        {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            const int pix = p0 + q_k;
            if (pix < pix1) {
                const int wo = pix % P.Wp, ho = (pix / P.Wp) % P.Hp, n = pix / (P.Wp * P.Hp);
                const int col = c_base + q_c;
                if (q_fast) {
                    if (col < ncols) {
                        const int tap = col / P.Cb, b = col % P.Cb;
                        const int hi = ho * P.stride - P.pad + tap / P.S, wi = wo * P.stride - P.pad + tap % P.S;
                        if (hi >= 0 && hi < P.Hq && wi >= 0 && wi < P.Wq) {
                            const int64_t off = (int64_t)n * P.q.sn + (int64_t)hi * P.q.sh + (int64_t)wi * P.q.sw + b;
                            if (P.q.dtype == DMU_BF16) {
                                const __nv_bfloat16* qp = reinterpret_cast<const __nv_bfloat16*>(P.q.ptr) + off;
                                uint2 raw = *reinterpret_cast<const uint2*>(qp);
                                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
                                float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
                                v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
                            } else {
                                load_vec<float>(reinterpret_cast<const float*>(P.q.ptr) + off, v);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = col + j;
                        if (c < ncols) {
                            const int tap = c / P.Cb, b = c % P.Cb;
                            const int hi = ho * P.stride - P.pad + tap / P.S, wi = wo * P.stride - P.pad + tap % P.S;
                            if (hi >= 0 && hi < P.Hq && wi >= 0 && wi < P.Wq)
                                v[j] = ld_as_float(P.q.ptr, (int64_t)n * P.q.sn + (int64_t)hi * P.q.sh + (int64_t)wi * P.q.sw + (int64_t)b * P.q.sc, P.q.dtype);
                        }
                    }
                }
            }
            *reinterpret_cast<float4*>(&Qs[q_k][q_c]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < WK; ++k) {
            float av[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) av[i] = Ps[k][ty * TM + i];
            const float4 b = *reinterpret_cast<const float4*>(&Qs[k][tx * 4]);
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < TM; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                if (do_bias) bias_acc[i] += av[i];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int a = a_base + ty * TM + i;
        if (a >= P.Ca) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c_base + tx * 4 + j;
            if (c < ncols) {
                const int tap = c / P.Cb, b = c % P.Cb;
                atomicAdd(&P.dw[(int64_t)a * P.dw_sa + (int64_t)b * P.dw_sb + (int64_t)tap * P.dw_st], acc[i][j]);
            }
        }
        if (do_bias) atomicAdd(&P.dbias[a], bias_acc[i]);
    }
}

}  // namespace dmu

using namespace dmu;

// implemented in conv_tc.cu
extern "C" int dmu_conv2d_tc(const dmu_conv_params* p, dmu_stream_t stream);
extern "C" int dmu_conv2d_tc_supported(const dmu_conv_params* p);
extern "C" int dmu_wgrad_tc(const dmu_wgrad_params* p, dmu_stream_t stream);
extern "C" int dmu_wgrad_tc_supported(const dmu_wgrad_params* p);
// implemented in conv_edge.cu
extern "C" int dmu_conv2d_edge(const dmu_conv_params* p, dmu_stream_t stream);
extern "C" int dmu_conv2d_edge_supported(const dmu_conv_params* p);
extern "C" int dmu_wgrad_edge(const dmu_wgrad_params* p, dmu_stream_t stream);
extern "C" int dmu_wgrad_edge_supported(const dmu_wgrad_params* p);

extern "C" {

int dmu_conv2d(const dmu_conv_params* p, dmu_stream_t stream) {
    DMU_REQUIRE(p, "dmu_conv2d: null params");
    DMU_REQUIRE(p->x.ptr && p->y.ptr && p->w, "dmu_conv2d: null pointer");
    DMU_REQUIRE(p->N > 0 && p->Hi > 0 && p->Wi > 0 && p->Ck > 0 && p->Ho > 0 && p->Wo > 0 && p->Cj > 0, "dmu_conv2d: non-positive dims");
    DMU_REQUIRE(p->R > 0 && p->S > 0 && p->stride > 0 && p->pad >= 0, "dmu_conv2d: bad filter geometry");
    DMU_REQUIRE(p->gather == 0 || p->gather == 1, "dmu_conv2d: gather must be 0 or 1");
    DMU_REQUIRE((int64_t)p->N * p->Ho * p->Wo < (1ll << 31), "dmu_conv2d: too many output pixels");
    if (p->gn_fuse_mode) {
        DMU_REQUIRE(dmu_conv2d_gn_fuse_supported(p) > 0,
                    "dmu_conv2d: gn_fuse is set but this launch cannot take the GroupNorm in its epilogue (dmu_conv2d_gn_fuse_supported == 0)");
        return dmu_conv2d_tc(p, stream);
    }
    if (p->gn_coef) {
        DMU_REQUIRE(p->impl != 1 && p->impl != 3 && p->impl != 4 && dmu_conv2d_tc_supported(p),
                    "dmu_conv2d: fused GroupNorm (gn_coef) needs the halo kernel: bf16 NHWC, channels %% 64 == 0, 3x3 stride 1 pad 1, >= 8x8");
        return dmu_conv2d_tc(p, stream);
    }
    if (p->impl == 2 || p->impl == 4 || p->impl == 5) {
        DMU_REQUIRE(dmu_conv2d_tc_supported(p), "dmu_conv2d: impl=tcgen05 requested for an unsupported shape");
        return dmu_conv2d_tc(p, stream);
    }
    if (p->impl == 0 && dmu_conv2d_tc_supported(p)) return dmu_conv2d_tc(p, stream);
    if (p->impl == 3) {
        DMU_REQUIRE(dmu_conv2d_edge_supported(p), "dmu_conv2d: impl=edge requested for an unsupported shape");
        return dmu_conv2d_edge(p, stream);
    }
    if (p->impl == 0 && dmu_conv2d_edge_supported(p)) return dmu_conv2d_edge(p, stream);
    const int M = p->N * p->Ho * p->Wo;
    const int K = p->R * p->S * p->Ck;
    // Linear layer with few rows and a long contraction: split K across CTAs (fp32 rows, zeroed then accumulated)
    if (p->R == 1 && p->S == 1 && p->stride == 1 && p->pad == 0 && p->Hi == 1 && p->Wi == 1 && p->Ho == 1 && p->Wo == 1 && !p->temb &&
        p->y.dtype == DMU_F32 && p->y.sc == 1 && p->y.sn == p->Cj && p->y.ptr != p->res.ptr &&
        ((M + BM - 1) / BM) * ((p->Cj + BN - 1) / BN) < sm_count() / 2 && (int64_t)M * p->Cj * K >= (1 << 18)) {
        dim3 grid((M + LM - 1) / LM, (p->Cj + LN - 1) / LN, 1);
        int splits = (2 * sm_count() + grid.x * grid.y - 1) / (grid.x * grid.y);
        // Contractions of up to 256 (every Linear of the FORWARD: the time-embedding MLP and the 22-way time projection) run unsplit
        // with plain stores: fp32 atomics would make the forward differ from run to run in its last bits, which a bf16 network
        // amplifies to ~1e-2 (tests/test_gpu_unet.py::test_bf16_forward_is_bit_identical_from_run_to_run).  The long contractions of
        // the backward (K = 3328 of the projection's input gradient) keep the split.
        int per = (K + splits - 1) / splits;
        per = ((per + LK - 1) / LK) * LK;
        if (per < 2 * LK) per = 2 * LK;
        grid.z = (K + per - 1) / per;
        // Up to 8 splits (every Linear of the FORWARD: K <= 256) are reduced in a fixed order inside a cluster; the long contractions
        // of the backward (K = 3328 of the projection's input gradient: 16 splits and more) keep the fp32 atomics.
        static const int ordered = [] { const char* e = getenv("DMU_LINEAR_ORDERED"); return e ? atoi(e) : 1; }();      // A/B aid (0: atomics)
        if (ordered && grid.z > 1 && grid.z <= 8) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid; cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = as_stream(stream);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = grid.z;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, linear_splitk_kernel, *p, per, 1);
            if (e != cudaSuccess) return fail("dmu_conv2d/linear_splitk: cluster launch failed: %s", cudaGetErrorString(e));
            return check_launch("dmu_conv2d/linear_splitk");
        }
        if (grid.z > 1 && cudaMemsetAsync(p->y.ptr, 0, (size_t)M * p->Cj * sizeof(float), as_stream(stream)) != cudaSuccess) return check_launch("dmu_conv2d/linear zero");
        linear_splitk_kernel<<<grid, 128, 0, as_stream(stream)>>>(*p, per, 0);
        return check_launch("dmu_conv2d/linear_splitk");
    }
    if (p->Cj <= 4 && K <= kSmallMaxK) {
        const size_t smem = (size_t)K * 4 * sizeof(float);
        if (smem > 48 * 1024) cudaFuncSetAttribute(conv_small_n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv_small_n_kernel<<<(M + 127) / 128, 128, smem, as_stream(stream)>>>(*p);
        return check_launch("dmu_conv2d/small_n");
    }
    dim3 grid((M + BM - 1) / BM, (p->Cj + BN - 1) / BN);
    const int xvec = p->x.dtype == DMU_BF16 ? 8 : 4;
    const bool fast = p->x.sc == 1 && p->Ck % 8 == 0 && p->x.sw % xvec == 0 && p->x.sh % xvec == 0 && p->x.sn % xvec == 0 &&
                      (reinterpret_cast<uintptr_t>(p->x.ptr) & 15) == 0;
    if (fast) conv_gemm_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(*p);
    else conv_gemm_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(*p);
    return check_launch("dmu_conv2d/simt");
}

int dmu_conv2d_wgrad(const dmu_wgrad_params* p, dmu_stream_t stream) {
    DMU_REQUIRE(p, "dmu_conv2d_wgrad: null params");
    DMU_REQUIRE(p->p.ptr && p->q.ptr && p->dw, "dmu_conv2d_wgrad: null pointer");
    DMU_REQUIRE(p->N > 0 && p->Hp > 0 && p->Wp > 0 && p->Ca > 0 && p->Hq > 0 && p->Wq > 0 && p->Cb > 0, "dmu_conv2d_wgrad: non-positive dims");
    DMU_REQUIRE(p->R > 0 && p->S > 0 && p->stride > 0 && p->pad >= 0, "dmu_conv2d_wgrad: bad filter geometry");
    if (p->impl == 2 || p->impl == 5) {      // 5: the halo weight-gradient kernel wherever its geometry allows (tests), else as 2
        DMU_REQUIRE(dmu_wgrad_tc_supported(p), "dmu_conv2d_wgrad: impl=tcgen05 requested for an unsupported shape");
        return dmu_wgrad_tc(p, stream);
    }
    if (p->impl == 0 && dmu_wgrad_tc_supported(p)) return dmu_wgrad_tc(p, stream);
    if (p->impl == 3) {
        DMU_REQUIRE(dmu_wgrad_edge_supported(p), "dmu_conv2d_wgrad: impl=edge requested for an unsupported shape");
        return dmu_wgrad_edge(p, stream);
    }
    if (p->impl == 0 && dmu_wgrad_edge_supported(p)) return dmu_wgrad_edge(p, stream);
    const int Mtot = p->N * p->Hp * p->Wp;
    const int ncols = p->R * p->S * p->Cb;
    const bool small_a = p->Ca <= 16;
    const int BMw = small_a ? 16 : 64;
    dim3 grid((ncols + 63) / 64, (p->Ca + BMw - 1) / BMw, 1);
    const int tiles = grid.x * grid.y;
    int splits = (sm_count() * 4 + tiles - 1) / tiles;
    const int max_splits = (Mtot + 63) / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int per = (Mtot + splits - 1) / splits;
    per = ((per + WK - 1) / WK) * WK;
    splits = (Mtot + per - 1) / per;
    grid.z = splits;
    if (small_a) wgrad_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(*p, per);
    else wgrad_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(*p, per);
    return check_launch("dmu_conv2d_wgrad/simt");
}

}  // extern "C"
