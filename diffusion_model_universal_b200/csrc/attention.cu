// Low-resolution multi-head self-attention core (models/layers/attention.py:49-61).
// The whole (image, all heads) problem — S <= 64 tokens — lives in one CTA's
// shared memory; one thread owns one (head, query) row and runs a flash-style
// online softmax, so scores are never materialised in HBM.
#include "common.cuh"

namespace dmu {

template <typename T>
__device__ __forceinline__ void stage_rows(const T* src, int64_t pitch, int rows, int cols, float* dst) {
    constexpr int kVec = Elem<T>::kVec;
    const int vpr = cols / kVec;
    for (int i = threadIdx.x; i < rows * vpr; i += blockDim.x) {
        const int r = i / vpr, v = i % vpr;
        float tmp[kVec];
        load_vec<T>(src + (int64_t)r * pitch + v * kVec, tmp);
#pragma unroll
        for (int k = 0; k < kVec; ++k) dst[r * cols + v * kVec + k] = tmp[k];
    }
}

template <typename T, int D>
__global__ void __launch_bounds__(256) attn_fwd_kernel(dmu_attn_params P) {
    extern __shared__ float sm[];  // qkv [S][3C]
    const int n = blockIdx.x, S = P.S, C = P.C;
    const T* qkv = reinterpret_cast<const T*>(P.qkv) + (int64_t)n * S * P.qkv_pitch;
    stage_rows<T>(qkv, P.qkv_pitch, S, 3 * C, sm);
    __syncthreads();
    const int h = threadIdx.x / S, i = threadIdx.x % S;
    if (h >= P.heads) return;
    const float scale = rsqrtf((float)D);
    float q[D], o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { q[d] = sm[i * 3 * C + h * D + d] * scale; o[d] = 0.f; }
    float mx = -INFINITY, l = 0.f;
    for (int j = 0; j < S; ++j) {
        const float* kj = sm + j * 3 * C + C + h * D;
        const float* vj = sm + j * 3 * C + 2 * C + h * D;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) s = fmaf(q[d], kj[d], s);
        const float mn = fmaxf(mx, s);
        const float corr = expf(mx - mn);
        const float p = expf(s - mn);
        l = l * corr + p;
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = o[d] * corr + p * vj[d];
        mx = mn;
    }
    const float inv = 1.f / l;
    T* orow = reinterpret_cast<T*>(P.o) + ((int64_t)n * S + i) * P.o_pitch + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) orow[d] = Elem<T>::from_f(o[d] * inv);
    if (P.lse) P.lse[((int64_t)n * P.heads + h) * S + i] = mx + logf(l);
}

// Forward, one CTA per (image, head): KSPLIT adjacent lanes share a query and each walks every KSPLIT-th key with its own
// online softmax; the partial (max, sum, output) triples meet through warp shuffles.  The single-CTA-per-image version above
// is one serial chain of S x ~100 instructions per thread (46 us at S = 64, batch 256: 2 waves of 256 CTAs); here the chain
// is KSPLIT times shorter and 4x as many CTAs hide each other's latency.
template <typename T, int D, int KSPLIT>
__global__ void __launch_bounds__(256) attn_fwd_split_kernel(dmu_attn_params P) {
    extern __shared__ __align__(16) float sm[];  // q|k|v of this head: [S][3D + 4]
    constexpr int kRow = 3 * D + 4;
    constexpr int kVec = Elem<T>::kVec;
    const int n = blockIdx.x, h = blockIdx.y, S = P.S, C = P.C;
    const T* qkv = reinterpret_cast<const T*>(P.qkv) + (int64_t)n * S * P.qkv_pitch;
    constexpr int vpr = D / kVec;                 // 16-byte vectors per (row, q|k|v)
    for (int idx = threadIdx.x; idx < S * 3 * vpr; idx += blockDim.x) {
        const int r = idx / (3 * vpr), w = (idx / vpr) % 3, v = idx % vpr;
        float tmp[kVec];
        load_vec<T>(qkv + (int64_t)r * P.qkv_pitch + w * C + h * D + v * kVec, tmp);
#pragma unroll
        for (int k = 0; k < kVec; ++k) sm[r * kRow + w * D + v * kVec + k] = tmp[k];
    }
    __syncthreads();
    const int i = threadIdx.x / KSPLIT, part = threadIdx.x % KSPLIT;
    const bool valid = i < S;
    const int iq = valid ? i : 0;
    const float scale = rsqrtf((float)D);
    float q[D], o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { q[d] = sm[iq * kRow + d] * scale; o[d] = 0.f; }
    float mx = -INFINITY, l = 0.f;
    for (int j = part; j < S; j += KSPLIT) {
        const float* kj = sm + j * kRow + D;
        const float* vj = sm + j * kRow + 2 * D;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) s = fmaf(q[d], kj[d], s);
        const float mn = fmaxf(mx, s);
        const float corr = expf(mx - mn);
        const float p = expf(s - mn);
        l = l * corr + p;
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = o[d] * corr + p * vj[d];
        mx = mn;
    }
#pragma unroll
    for (int off = 1; off < KSPLIT; off <<= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, mx, off), l2 = __shfl_xor_sync(0xffffffffu, l, off);
        const float mn = fmaxf(mx, m2);
        // a part that saw no key (S < KSPLIT) carries (-inf, 0, 0): its weight is exp(-inf) = 0
        const float c1 = mn == -INFINITY ? 0.f : expf(mx - mn), c2 = mn == -INFINITY ? 0.f : expf(m2 - mn);
        l = l * c1 + l2 * c2;
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = o[d] * c1 + __shfl_xor_sync(0xffffffffu, o[d], off) * c2;
        mx = mn;
    }
    if (!valid) return;
    const float inv = 1.f / l;
    T* orow = reinterpret_cast<T*>(P.o) + ((int64_t)n * S + i) * P.o_pitch + h * D;
    constexpr int kPer = D / KSPLIT;              // every part writes its slice of the (identical) merged row
#pragma unroll
    for (int d = 0; d < D; ++d)
        if (d / kPer == part) orow[d] = Elem<T>::from_f(o[d] * inv);
    if (P.lse && part == 0) P.lse[((int64_t)n * P.heads + h) * S + i] = mx + logf(l);
}

// Backward, one CTA per image.  KS adjacent lanes share one (head, row): in phase 1 they split the keys of query i (dQ_i), in
// phase 2 the queries of key j (dK_j, dV_j); partial vectors meet through warp shuffles.  With KS = 1 the CTA has only
// S * heads threads (64 at 4x4) walking 2 x S x 7D dependent FMAs each; KS = 4 fills the CTA and cuts the chain fourfold.
template <typename T, int D, int KS>
__global__ void __launch_bounds__(256) attn_bwd_kernel(dmu_attn_params P) {
    extern __shared__ float sm[];
    const int n = blockIdx.x, S = P.S, C = P.C, H = P.heads;
    float* s_qkv = sm;                 // [S][3C]
    float* s_do = s_qkv + S * 3 * C;   // [S][C]
    float* s_lse = s_do + S * C;       // [H][S]
    float* s_delta = s_lse + H * S;    // [H][S]
    const T* qkv = reinterpret_cast<const T*>(P.qkv) + (int64_t)n * S * P.qkv_pitch;
    const T* dO = reinterpret_cast<const T*>(P.d_o) + (int64_t)n * S * P.do_pitch;
    stage_rows<T>(qkv, P.qkv_pitch, S, 3 * C, s_qkv);
    stage_rows<T>(dO, P.do_pitch, S, C, s_do);
    const int part = threadIdx.x % KS, hi = threadIdx.x / KS;
    const int h = hi / S, i = hi % S;
    const bool act = h < H;
    const int hh = act ? h : 0;        // idle lanes run the loops on head 0 (they take part in the shuffles) and store nothing
    const float scale = rsqrtf((float)D);
    if (act && part == 0) {
        const T* orow = reinterpret_cast<const T*>(P.o) + ((int64_t)n * S + i) * P.o_pitch + h * D;
        const T* drow = dO + (int64_t)i * P.do_pitch + h * D;
        float dl = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) dl = fmaf(Elem<T>::to_f(orow[d]), Elem<T>::to_f(drow[d]), dl);
        s_delta[h * S + i] = dl;
        s_lse[h * S + i] = P.lse[((int64_t)n * H + h) * S + i];
    }
    __syncthreads();
    T* dq_row = reinterpret_cast<T*>(P.dqkv) + ((int64_t)n * S + i) * P.dqkv_pitch;
    // phase 1: this lane group is query i -> dQ_i
    {
        float q[D], dov[D], dq[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { q[d] = s_qkv[i * 3 * C + hh * D + d]; dov[d] = s_do[i * C + hh * D + d]; dq[d] = 0.f; }
        const float lse = s_lse[hh * S + i], dl = s_delta[hh * S + i];
        for (int j = part; j < S; j += KS) {
            const float* kj = s_qkv + j * 3 * C + C + hh * D;
            const float* vj = s_qkv + j * 3 * C + 2 * C + hh * D;
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) { s = fmaf(q[d], kj[d], s); dp = fmaf(dov[d], vj[d], dp); }
            const float p = expf(s * scale - lse);
            const float ds = p * (dp - dl) * scale;
#pragma unroll
            for (int d = 0; d < D; ++d) dq[d] = fmaf(ds, kj[d], dq[d]);
        }
#pragma unroll
        for (int off = 1; off < KS; off <<= 1)
#pragma unroll
            for (int d = 0; d < D; ++d) dq[d] += __shfl_xor_sync(0xffffffffu, dq[d], off);
        if (act) {
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (d % KS == part) dq_row[h * D + d] = Elem<T>::from_f(dq[d]);
        }
    }
    // phase 2: this lane group is key/value j = i -> dK_j, dV_j
    {
        const int j = i;
        float k[D], v[D], dk[D], dv[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { k[d] = s_qkv[j * 3 * C + C + hh * D + d]; v[d] = s_qkv[j * 3 * C + 2 * C + hh * D + d]; dk[d] = 0.f; dv[d] = 0.f; }
        for (int ii = part; ii < S; ii += KS) {
            const float* qi = s_qkv + ii * 3 * C + hh * D;
            const float* doi = s_do + ii * C + hh * D;
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) { s = fmaf(qi[d], k[d], s); dp = fmaf(doi[d], v[d], dp); }
            const float p = expf(s * scale - s_lse[hh * S + ii]);
            const float ds = p * (dp - s_delta[hh * S + ii]) * scale;
#pragma unroll
            for (int d = 0; d < D; ++d) { dv[d] = fmaf(p, doi[d], dv[d]); dk[d] = fmaf(ds, qi[d], dk[d]); }
        }
#pragma unroll
        for (int off = 1; off < KS; off <<= 1)
#pragma unroll
            for (int d = 0; d < D; ++d) {
                dk[d] += __shfl_xor_sync(0xffffffffu, dk[d], off);
                dv[d] += __shfl_xor_sync(0xffffffffu, dv[d], off);
            }
        if (act) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                if (d % KS == part) {
                    dq_row[C + h * D + d] = Elem<T>::from_f(dk[d]);
                    dq_row[2 * C + h * D + d] = Elem<T>::from_f(dv[d]);
                }
            }
        }
    }
}

static int attn_check(const dmu_attn_params* p, const char* who, bool bwd) {
    DMU_REQUIRE(p, "%s: null params", who);
    DMU_REQUIRE(p->qkv && p->o, "%s: null pointer", who);
    DMU_REQUIRE(p->N > 0 && p->S > 0 && p->C > 0 && p->heads > 0, "%s: non-positive dims", who);
    DMU_REQUIRE(p->C % p->heads == 0, "%s: C=%d not divisible by heads=%d", who, p->C, p->heads);
    DMU_REQUIRE(p->S * p->heads <= 256, "%s: S*heads=%d exceeds one CTA (256 rows)", who, p->S * p->heads);
    const int vec = p->dtype == DMU_BF16 ? 8 : 4;
    DMU_REQUIRE(p->C % vec == 0 && p->qkv_pitch % vec == 0 && p->o_pitch % vec == 0, "%s: C/pitches must be multiples of %d", who, vec);
    if (bwd) {
        DMU_REQUIRE(p->d_o && p->dqkv && p->lse, "%s: null pointer", who);
        DMU_REQUIRE(p->do_pitch % vec == 0 && p->dqkv_pitch % vec == 0, "%s: pitches must be multiples of %d", who, vec);
    }
    return 0;
}

template <typename T>
static int attn_launch(const dmu_attn_params* p, cudaStream_t s, bool bwd) {
    const int D = p->C / p->heads;
    size_t smem = (size_t)p->S * 3 * p->C * sizeof(float);
    if (bwd) smem += ((size_t)p->S * p->C + 2 * (size_t)p->heads * p->S) * sizeof(float);
    DMU_REQUIRE(smem <= 227 * 1024, "dmu_attn: S=%d C=%d needs %zu B of shared memory", p->S, p->C, smem);
    const int threads = ((p->S * p->heads + 31) / 32) * 32;
#define ATTN_CASE(DD)                                                                                               \
    case DD: {                                                                                                      \
        auto kf = attn_fwd_kernel<T, DD>;                                                                           \
        const bool split4 = bwd && p->S * p->heads * 4 <= 256;                                                      \
        auto kb = split4 ? attn_bwd_kernel<T, DD, 4> : attn_bwd_kernel<T, DD, 1>;                                   \
        if (smem > 48 * 1024) {                                                                                     \
            cudaFuncSetAttribute(bwd ? (const void*)kb : (const void*)kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        }                                                                                                           \
        if (bwd) kb<<<p->N, split4 ? ((p->S * p->heads * 4 + 31) / 32) * 32 : threads, smem, s>>>(*p);              \
        else if (p->S * 4 <= 256 && DD >= 8) {                                                                      \
            const int th = ((p->S * 4 + 31) / 32) * 32;                                                             \
            attn_fwd_split_kernel<T, DD, 4><<<dim3(p->N, p->heads), th, (size_t)p->S * (3 * DD + 4) * sizeof(float), s>>>(*p); \
        } else kf<<<p->N, threads, smem, s>>>(*p);                                                                  \
        break;                                                                                                      \
    }
    switch (D) {
        ATTN_CASE(8)
        ATTN_CASE(16)
        ATTN_CASE(32)
        ATTN_CASE(64)
        default:
            return fail("dmu_attn: head dim %d unsupported (8, 16, 32, 64)", D);
    }
#undef ATTN_CASE
    return check_launch(bwd ? "dmu_attn_bwd" : "dmu_attn_fwd");
}

namespace tc {      // attention_tc.cu
int attn_tc_supported(const dmu_attn_params* p);
int attn_tc_launch(const dmu_attn_params* p, cudaStream_t stream);
}  // namespace tc

}  // namespace dmu

using namespace dmu;

extern "C" {

int dmu_attn_fwd(const dmu_attn_params* p, dmu_stream_t stream) {
    if (int e = attn_check(p, "dmu_attn_fwd", false)) return e;
    // bf16, head dim 32 / 64, S | 128: scores and probabilities on the tensor cores (attention_tc.cu)
    if (tc::attn_tc_supported(p)) return tc::attn_tc_launch(p, as_stream(stream));
    return p->dtype == DMU_BF16 ? attn_launch<__nv_bfloat16>(p, as_stream(stream), false) : attn_launch<float>(p, as_stream(stream), false);
}
int dmu_attn_bwd(const dmu_attn_params* p, dmu_stream_t stream) {
    if (int e = attn_check(p, "dmu_attn_bwd", true)) return e;
    return p->dtype == DMU_BF16 ? attn_launch<__nv_bfloat16>(p, as_stream(stream), true) : attn_launch<float>(p, as_stream(stream), true);
}

}  // extern "C"
