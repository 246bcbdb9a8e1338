/*
 * dmu_b200.h — C ABI of the B200-native denoiser hot path.
 *
 * Drop-in boundary for ChristianLin0420/diffusion-model-universal's UNet
 * denoiser path.  The reference is pure Python: the boundary
 * it exposes is the class API `BaseDiffusion.forward / loss_function /
 * generate_samples` (models/base_model.py:57-117).  That API is mirrored in
 * Python by `diffusion_model_universal_b200/` and every arithmetic step
 * behind it is one of the entry points below.  Each entry point names the
 * reference call site(s) it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless
 *    stated; the library neither allocates nor frees nor retains them.
 *  - every call enqueues work on `stream` (a cudaStream_t) and returns
 *    without synchronising.  Return value 0 = ok, non-zero = error; the
 *    message is available from dmu_last_error() (thread-local).
 *  - activations are 4-D tensors described by element strides (dmu_tensor4),
 *    so NHWC-with-pitch (internal) and NCHW fp32 (API boundary) both work.
 *  - dtype codes: DMU_F32 = 0, DMU_BF16 = 1.  Accumulation is always fp32.
 */
#ifndef DMU_B200_H
#define DMU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMU_ABI_VERSION 1
#define DMU_F32 0
#define DMU_BF16 1

typedef void* dmu_stream_t; /* cudaStream_t */

typedef struct {
    void* ptr;
    int64_t sn, sh, sw, sc; /* element strides of (n, h, w, c) */
    int32_t dtype;          /* DMU_F32 | DMU_BF16 */
    int32_t _pad;
} dmu_tensor4;

int dmu_abi_version(void);
const char* dmu_last_error(void);
/* struct sizes, so a foreign-language binding can assert its mirror matches */
int dmu_sizeof(const char* struct_name);

/* ------------------------------------------------------------------ *
 * Diffusion-process updates (fp32, any contiguous [B, inner] layout)  *
 * ------------------------------------------------------------------ */

/* models/ddpm.py:286-296 `_add_noise`:  out = sqrt(acp[t]) x0 + sqrt(1-acp[t]) noise */
int dmu_q_sample(const float* x0, const float* noise, const int64_t* t, const float* alphas_cumprod,
                 float* out, int64_t batch, int64_t inner, dmu_stream_t stream);

/* models/ddpm.py:306-329 `_reverse_diffusion_step` after the eps prediction.
 * `noise` is the randn_like draw of ddpm.py:324 (may be NULL when the caller
 * knows t == 0).  The t[0] > 0 branch of ddpm.py:311,323 is taken on device
 * from t[0]; no host sync.  `num_timesteps` is the length of the three tables:
 * a row with t == 0 inside a batch whose t[0] > 0 reads alphas_cumprod[t-1] =
 * alphas_cumprod[num_timesteps-1], like the reference's tensor index -1 (ddpm.py:311).
 * out may alias x. */
int dmu_ddpm_step(const float* x, const float* eps, const float* noise, const int64_t* t,
                  const float* betas, const float* alphas, const float* alphas_cumprod, int64_t num_timesteps,
                  float* out, int64_t batch, int64_t inner, dmu_stream_t stream);

/* models/ddim.py:97-124 `_ddim_sample` after the eps prediction; `idx`
 * indexes the S-entry tables (repaired driver).  `noise`
 * NULL <=> eta == 0 (ddim.py:117-121).  Clamps x0 to [-1,1], noise to [-3,3]. */
int dmu_ddim_step(const float* x, const float* eps, const float* noise, const int64_t* idx,
                  const float* ddim_alphas, const float* ddim_alphas_prev, const float* ddim_sigmas,
                  const float* ddim_sqrt_one_minus_alphas,
                  float* out, int64_t batch, int64_t inner, dmu_stream_t stream);

/* models/score_based.py:236-245: step = 2 (sigma*beta)^2 ; out = x + step*score + sqrt(2 step)*noise.
 * sigma is read from device memory (sigmas[k]) exactly like the reference's 0-dim tensor. */
int dmu_langevin_score_step(const float* x, const float* score, const float* noise, const float* sigmas, int64_t k,
                            float beta, float* out, int64_t n, dmu_stream_t stream);

/* models/energy_based.py:271-273 (math.sqrt repair): out = x - step*grad + sqrt_2step*noise */
int dmu_langevin_energy_step(const float* x, const float* grad, const float* noise, float step, float sqrt_2step,
                             float* out, int64_t n, dmu_stream_t stream);

/* models/energy_based.py:240-246: out = sqrt(a[t-1]/a[t]) x + sqrt((1-a[t-1])/(1-a[t])) sqrt(1-a[t]/a[t-1]) noise */
int dmu_energy_renoise(const float* x, const float* noise, const float* alphas_cumprod, int64_t t,
                       float* out, int64_t n, dmu_stream_t stream);

/* out[b,i] = a[b]*x[b,i] + c[b]*z[b,i]  (a NULL = 1).  models/score_based.py:200-201 (x + sigma*noise) and the
 * score-matching target -noise/sigma of utils/losses.py:240. */
int dmu_scale_add(const float* x, const float* z, const float* a, const float* c, float* out,
                  int64_t batch, int64_t inner, dmu_stream_t stream);

/* utils/losses.py:144-181 `_get_time_weights`, time_weight_type 'snr', without the reference's timesteps.max().item() sync:
 *   acp = cumprod(1 - linspace(1e-4, 2e-2, t_max + 1))[t];  snr = acp / (1 - acp);  v = clamp(snr / max(snr), 1e-5);
 *   w = min_weight + weight_span * (v - min(v)) / (max(v) - min(v) + 1e-5)          (weight_span = max_weight - min_weight)
 * `table` is fp32 [num_timesteps][num_timesteps]: row tm = that cumprod vector for t_max = tm (entries past tm unused), built
 * once by the host with the reference's own torch calls; t int64 [batch] with values < num_timesteps; one launch. */
int dmu_snr_time_weights(const int64_t* t, const float* table, int64_t num_timesteps, int64_t batch, float min_weight, float weight_span,
                         float* w, dmu_stream_t stream);

/* utils/losses.py:74-131 `DiffusionLoss.__call__` minus the [B]-sized time
 * weights (computed by the host with the reference's own torch ops and passed
 * as `w`, NULL = unweighted):
 *   loss = mean_b,i( w[b] * (wm d^2 + wl |d| + wh smooth_l1(d; delta)) ),  d = pred - target
 * Writes the scalar to `loss` and, when dpred != NULL, d loss / d pred.
 * `partials` is workspace of dmu_loss_workspace_floats(batch*inner) floats. */
int64_t dmu_loss_workspace_floats(int64_t numel);
int dmu_diffusion_loss(const float* pred, const float* target, const float* w,
                       float wm, float wl, float wh, float delta,
                       float* loss, float* dpred, float* partials,
                       int64_t batch, int64_t inner, dmu_stream_t stream);

/* ------------------------------------------------------------------ *
 * UNet denoiser primitives                                            *
 * ------------------------------------------------------------------ */

/* Implicit-GEMM convolution (nn.Conv2d / nn.ConvTranspose2d / nn.Linear and
 * their input gradients):  models/layers/residual.py:33,40,44,91,121,178,242,
 * models/ddpm.py:49,90, models/layers/attention.py:21-24,
 * models/layers/embeddings.py:55,57, residual.py:36 (time_mlp).
 *
 *   y[n,ho,wo,j] = bias[j] + temb[n,j] + res[n,ho,wo,j]
 *                + sum_{r,s,k} x[n, hi, wi, k] * w[j*w_sn + k*w_sk + (r*S+s)*w_st]
 *   gather 0 (conv):        hi = ho*stride - pad + r
 *   gather 1 (transposed):  hi = (ho + pad - r) / stride   when divisible
 *
 * With the weight strides this one contraction covers fprop and dgrad of
 * both layer kinds.  impl: 0 = auto, 1 = generic SIMT fp32-FMA kernel, 2 = tcgen05 (bf16, channels % 64 == 0),
 * 3 = the direct kernels for the 3-channel boundary layers (at most 4 channels on one side, stride 1),
 * 4 = tcgen05 per-tap kernel only, 5 = tcgen05 halo kernels only (3x3 stride 1 pad 1 of >= 8x8; 4x4 stride 2 pad 1 with 64 input
 * channels, conv or transposed gather: an error for any other shape); impl 2 picks between the two families by measured shape
 * heuristics.
 */
typedef struct {
    dmu_tensor4 x;       /* gathered input, channels K = Ck */
    dmu_tensor4 y;       /* output, channels J = Cj */
    dmu_tensor4 res;     /* optional residual added in the epilogue (ptr NULL = none) */
    const void* w;       /* weights, dtype w_dtype */
    int64_t w_sn, w_sk, w_st;
    const float* bias;   /* [Cj] fp32 or NULL */
    const float* temb;   /* [N, Cj] fp32 (row pitch temb_pitch) or NULL */
    int64_t temb_pitch;
    int32_t N, Hi, Wi, Ck;
    int32_t Ho, Wo, Cj;
    int32_t R, S, stride, pad;
    int32_t gather;      /* 0 conv, 1 transposed */
    int32_t w_dtype;
    int32_t impl;
    int32_t _pad;
    /* Optional split-K scratch for the tensor-core path (layers with few output pixels and a long contraction are
     * split along K over the CTAs of a thread-block cluster, which park their partial tiles here before folding
     * them).  Caller-owned, contents undefined before and after, at least dmu_conv2d_workspace_bytes(); it must not
     * be shared by launches that may run concurrently.  NULL/0 = never split. */
    void* workspace;
    int64_t workspace_bytes;
    /* Optional fused GroupNorm(+SiLU) on the INPUT (residual.py:57-58,63-64, ddpm.py:88-90: every conv of the network reads
     * act(GroupNorm(x))):  the contraction runs over a[n,h,w,k] = act(x[n,h,w,k] * gn_coef[(n*Ck+k)*2] + gn_coef[(n*Ck+k)*2+1])
     * instead of x (act = SiLU when gn_silu, identity otherwise; padding stays zero), coefficients from dmu_gn_coef.
     * a_out (ptr NULL = skip) receives a, bf16 NHWC: the weight gradient needs it.  Only the persistent halo kernel
     * implements this (ask dmu_conv2d_gn_supported); any other shape with gn_coef set is an error. */
    const float* gn_coef;
    int32_t gn_silu;
    int32_t _pad2;
    dmu_tensor4 a_out;
    /* Optional GroupNorm fused into the EPILOGUE (the <= 8x8 stages, where one output tile of a CTA holds whole images and whole
     * groups; ask dmu_conv2d_gn_fuse_supported).  gn_fuse points to a dmu_gn_params (declared below) with N, H, W = the conv's
     * output, C = Cj:
     *   gn_fuse_mode 1, forward (residual.py:57,63, attention.py:68: the next layer normalises this output):  y is written as
     *     usual and additionally gn->y = act(GroupNorm(y)); gn->sums[n, g, 0..1] receives the raw sums (plain stores).
     *     gn->x, dx, add0, add1, red are ignored.
     *   gn_fuse_mode 2, backward (this launch is the dgrad that produces dy of a GroupNorm): the convolution result is NOT
     *     stored; the epilogue applies dmu_gn_backward's arithmetic to it:  gn->dx = gn_backward(gn->x, dy; gn->sums, gamma,
     *     beta) + add0 + add1, and gn->red receives PER-TILE channel sums [tiles][C][2] = (sum du, sum du*xhat) over the
     *     images of each pixel tile (tiles = the return value of dmu_conv2d_gn_fuse_supported; fold them with
     *     dmu_gn_param_grads, desc.count = tiles).  res / bias / temb must be NULL; gn->y is ignored.
     *   gn_fuse_mode 3, statistics only (the large layers of the persistent 3x3 kernel, where an image spans many tiles): y is
     *     written as usual and the raw sums of the stored (rounded) output are ADDED to gn->sums[n, g, 0..1] with atomics
     *     (caller zeroes them, like dmu_gn_stats); the next layer's norm is then dmu_gn_apply alone - its statistics pass over
     *     the tensor is gone.  Only gn->sums, N, H, W, C, G are read.
     * Any shape dmu_conv2d_gn_fuse_supported rejects is an error with gn_fuse set. */
    const void* gn_fuse;
    int32_t gn_fuse_mode;
    int32_t _pad3;
} dmu_conv_params;
int dmu_conv2d(const dmu_conv_params* p, dmu_stream_t stream);
/* 0 when the launch cannot take p->gn_fuse (set, with its mode) in its epilogue; otherwise the number of pixel tiles of the
 * launch (>= 1; mode 2 writes that many rows of per-tile channel sums).  Pure host logic: no pointer is dereferenced except
 * p and p->gn_fuse themselves. */
int dmu_conv2d_gn_fuse_supported(const dmu_conv_params* p);
/* 1 when dmu_conv2d would run this layer on the halo kernel by its own heuristics, i.e. when the GroupNorm of its input may be
 * fused into it (gn_coef itself need not be set yet). */
int dmu_conv2d_gn_supported(const dmu_conv_params* p);
/* bytes of split-K scratch that let dmu_conv2d split every eligible layer (a constant upper bound) */
int64_t dmu_conv2d_workspace_bytes(void);

/* Weight gradient of the same contraction (autograd of the call sites above):
 *   dw[a*dw_sa + b*dw_sb + (r*S+s)*dw_st] += sum_{n,po,qo} p[n,po,qo,a] * q[n, po*stride-pad+r, qo*stride-pad+s, b]
 * p = tensor on the strided (small) grid, q = gathered tensor.  fp32 output,
 * accumulated with atomics (caller zeroes dw).  Optionally also
 * dbias[a] += sum p[.,.,.,a].
 * impl: 0 = auto, 1 = SIMT, 2 = tcgen05 (auto between the per-tap kernel and, for 3x3 stride-1 layers with many pixel
 * tiles, the halo kernel that reads p and q once per tap group), 3 = narrow-operand (3-channel) kernels, 5 = tcgen05 with the
 * halo kernel wherever its geometry allows. */
typedef struct {
    dmu_tensor4 p;
    dmu_tensor4 q;
    float* dw;
    int64_t dw_sa, dw_sb, dw_st;
    float* dbias;        /* [Ca] or NULL */
    int32_t N, Hp, Wp, Ca;
    int32_t Hq, Wq, Cb;
    int32_t R, S, stride, pad;
    int32_t impl;
} dmu_wgrad_params;
int dmu_conv2d_wgrad(const dmu_wgrad_params* p, dmu_stream_t stream);

/* nn.GroupNorm (+ nn.SiLU): residual.py:31-32,38-39, attention.py:27,68, ddpm.py:88-89,
 * energy_based.py:56-57,79-80.  NHWC tensors (sc == 1).
 *  stats:   sums[n,g,0..1] += (sum x, sum x^2) over the group  (caller zeroes sums)
 *  apply:   y = act((x - mean) * rstd * gamma + beta),  act = SiLU if silu else identity
 *  bwd_reduce: red[n,c,0..1] += (sum_p du, sum_p du*xhat), du = dy * act'(u); dgamma/dbeta += over n
 *  bwd_apply:  dx = rstd*(du*gamma - (A + xhat*Bq)/cnt) + add0 + add1
 */
typedef struct {
    dmu_tensor4 x;       /* input of the norm */
    dmu_tensor4 y;       /* output (apply) | dy (bwd) */
    dmu_tensor4 dx;      /* bwd_apply output */
    dmu_tensor4 add0, add1; /* optional addends for dx (ptr NULL = none) */
    float* sums;         /* [N, G, 2] raw sums */
    const float* gamma;  /* [C] */
    const float* beta;   /* [C] */
    float* red;          /* [N, C, 2] bwd workspace */
    float* dgamma;       /* [C] accumulated */
    float* dbeta;        /* [C] accumulated */
    int32_t N, H, W, C, G;
    int32_t silu;
    float eps;
    int32_t flags;       /* DMU_GN_FIXED_SUMS (1): order-independent statistics - see below */
} dmu_gn_params;
/* flags & DMU_GN_FIXED_SUMS: the launches that ACCUMULATE raw sums with atomics (dmu_gn_stats, a dmu_conv2d with gn_fuse_mode 3)
 * add them as 64-bit fixed-point integers (value * 2^20, two's complement) into the [N, G, 2] int64 array that FOLLOWS the float
 * array (at sums + N*G*2, 8-byte aligned, zeroed by the caller like sums), and dmu_gn_apply / the apply half of dmu_gn_forward read
 * that array and store its float value into sums[n, g, 0..1] for every later consumer (the backward).  Integer addition is
 * associative, so the statistics - and with them a bf16 forward - are bit-identical from run to run; float atomics are not
 * (two runs of a bf16 UNet differ by ~1e-2 rel-L2 once the network has amplified a few different roundings). */
#define DMU_GN_FIXED_SUMS 1
/* One-call forms: forward = stats + apply (sums must NOT be pre-zeroed: they are overwritten when the single-pass
 * cluster kernel runs and accumulated by the two-pass fallback, so zero them as for dmu_gn_stats), backward =
 * bwd_reduce + bwd_apply.  An image that fits the registers of at most 8 CTAs (a thread-block cluster, partial sums
 * exchanged through distributed shared memory) is read once; larger ones fall back to the two-pass kernels below. */
int dmu_gn_forward(const dmu_gn_params* p, dmu_stream_t stream);
int dmu_gn_backward(const dmu_gn_params* p, dmu_stream_t stream);
/* Batch reduction of the affine-parameter gradients for a table of layers in one launch:
 *   dgamma[c] += sum_n red[n,c,1];  dbeta[c] += sum_n red[n,c,0]
 * (run once after the backward of all layers, with dgamma/dbeta left NULL in their dmu_gn_params).  The table lives in
 * device memory: n_desc entries of { const float* red; float* dgamma; float* dbeta; int32_t C; int32_t count; } where
 * count > 0 overrides N for that entry (rows of per-tile sums written by a fused dgrad epilogue, dmu_conv_params.gn_fuse). */
int dmu_gn_param_grads(const void* table_device, int32_t n_desc, int32_t max_c, int32_t N, dmu_stream_t stream);
/* Per-image, per-channel affine form of the normalisation, from the sums of dmu_gn_stats:
 *   coef[(n*C+c)*2] = rstd[n,g]*gamma[c],  coef[(n*C+c)*2+1] = beta[c] - mean[n,g]*rstd[n,g]*gamma[c]   (for dmu_conv_params.gn_coef) */
int dmu_gn_coef(const dmu_gn_params* p, float* coef, dmu_stream_t stream);
int dmu_gn_stats(const dmu_gn_params* p, dmu_stream_t stream);
int dmu_gn_apply(const dmu_gn_params* p, dmu_stream_t stream);
int dmu_gn_bwd_reduce(const dmu_gn_params* p, dmu_stream_t stream);
int dmu_gn_bwd_apply(const dmu_gn_params* p, dmu_stream_t stream);

/* Per-image and total channel sums of an NHWC tensor (bias / time_mlp grads):
 *   out_nc[n*pitch + c] += sum_p x[n,p,c] (optional);  out_c[c] += sum_{n,p} x[n,p,c] (optional)
 * Both outputs are accumulated with atomics (several CTAs per image): zero them first for a fresh sum.
 * scale multiplies both (1/HW gives the spatial mean of energy_based.py:83). */
int dmu_colsum(const dmu_tensor4* x, int32_t N, int32_t H, int32_t W, int32_t C,
               float* out_nc, int64_t out_nc_pitch, float* out_c, float scale, dmu_stream_t stream);

/* A whole table of dmu_colsum calls in one launch (all tensors of one dtype; the table lives in device memory).  Per entry:
 * x is N images of H*W pixels of C channels, `chunks` CTAs per image; out_nc / out_c / scale as in dmu_colsum.
 * Used for the bias gradients (trainers/ddpm_trainer.py:543-547: every conv / Linear bias) and the per-image time-projection
 * sums (residual.py:61) of one part of the backward. */
typedef struct {
    dmu_tensor4 x;
    float* out_nc;
    int64_t pitch;
    float* out_c;
    int32_t N, H, W, C;
    int32_t chunks;                  /* CTAs per image */
    float scale;
    int32_t cta0;                    /* first CTA of this entry: prefix sum of N * chunks over the table */
    int32_t _pad;
} dmu_colsum_desc;
int dmu_colsum_multi(const dmu_colsum_desc* table_device, int32_t n_desc, int32_t total_ctas, int32_t dtype, dmu_stream_t stream);

/* EnergyNet head, models/energy_based.py:79-83: out[n*pitch + c] += scale * sum_p silu(x[n,p,c])  (scale = 1/(H*W) gives the
 * mean; out is accumulated: zero it first) and its input gradient dx[n,p,c] = silu'(x[n,p,c]) * g[n*pitch + c] * scale. */
int dmu_silu_pool_fwd(const dmu_tensor4* x, int32_t N, int32_t H, int32_t W, int32_t C, float* out, int64_t pitch, float scale,
                      dmu_stream_t stream);
int dmu_silu_pool_bwd(const dmu_tensor4* x, const dmu_tensor4* dx, int32_t N, int32_t H, int32_t W, int32_t C, const float* g,
                      int64_t pitch, float scale, dmu_stream_t stream);

/* Second-order pieces of the EnergyBasedLoss gradient penalty (utils/losses.py:277-285, create_graph=True): the derivative
 * of the BACKWARD of y = act(GroupNorm(x)) and of the SiLU+pool head.  With dx = B(x, gamma, beta, dy) the first backward
 * and c a cotangent of dx:  gx = d<c,dx>/dx,  gdy = d<c,dx>/d(dy) (optional),  dgamma/dbeta += d<c,dx>/d(gamma/beta).
 * One CTA per image (three passes): sized for EnergyNet, not tuned. */
typedef struct {
    dmu_tensor4 x, dy, c, gx, gdy;   /* NHWC, one dtype; gdy.ptr may be NULL */
    const float* sums;               /* [N, G, 2] raw sums saved by the forward */
    const float* gamma;
    const float* beta;
    float* dgamma;                   /* [C] accumulated, or NULL */
    float* dbeta;
    int32_t N, H, W, C, G;
    int32_t silu;
    float eps;
    int32_t _pad;
} dmu_gn_bwd2_params;
int dmu_gn_bwd_bwd(const dmu_gn_bwd2_params* p, dmu_stream_t stream);
/* dh = silu'(h) g[n,c] scale is the first backward of dmu_silu_pool_fwd; c = cotangent of dh:
 *   gx[n,p,ch] = c silu''(h) g[n,ch] scale;   gg[n*gg_pitch + ch] += scale * sum_p c silu'(h) */
int dmu_silu_pool_bwd_bwd(const dmu_tensor4* x, const dmu_tensor4* c, const dmu_tensor4* gx, int32_t N, int32_t H, int32_t W, int32_t C,
                          const float* g, int64_t pitch, float* gg, int64_t gg_pitch, float scale, dmu_stream_t stream);

/* Multi-head self-attention core, attention.py:49-61.  qkv: [N, S, 3C] rows
 * (pitch), heads*d = C; o: [N, S, C].  lse: [N, heads, S] fp32 (saved for bwd).
 * bwd writes dqkv given do. */
typedef struct {
    const void* qkv; int64_t qkv_pitch;
    void* o; int64_t o_pitch;
    const void* d_o; int64_t do_pitch;
    void* dqkv; int64_t dqkv_pitch;
    float* lse;
    int32_t N, S, C, heads;
    int32_t dtype;
    int32_t _pad;
} dmu_attn_params;
int dmu_attn_fwd(const dmu_attn_params* p, dmu_stream_t stream);
int dmu_attn_bwd(const dmu_attn_params* p, dmu_stream_t stream);

/* embeddings.py:24-39: emb[b, 0:half] = sin(t[b] f_i), emb[b, half:] = cos(t[b] f_i),
 * f_i = exp(-i ln(1e4)/(half-1)).  t is int64 (t_is_float = 0) or fp32. */
int dmu_sinusoidal_embedding(const void* t, int32_t t_is_float, float* emb, int64_t batch, int32_t dim, dmu_stream_t stream);

/* Elementwise activations on fp32 rows: kind 0 = exact GELU (embeddings.py:56), 1 = SiLU (score_based.py:59),
 * 2 = log (score_based.py:82).  bwd: dx = dy * act'(x). */
int dmu_act_fwd(const float* x, float* y, int64_t n, int32_t kind, dmu_stream_t stream);
int dmu_act_bwd(const float* x, const float* dy, float* dx, int64_t n, int32_t kind, dmu_stream_t stream);

/* Layout / dtype repack of parameters into kernel-friendly caches (derived,
 * never the stored format; SURVEY.md §8b "Ownership").  One launch for a
 * whole table.  kind 0: OIHW -> [O][R][S][I];  kind 1: IOHW (ConvTranspose2d)
 * -> [O][R][S][I];  kind 2: plain copy/cast ([O][I] Linear, vectors);  kind 3: the inverse of kind 0,
 * [O][R][S][I] -> [O][I][R][S] (filter gradients are accumulated in the tap-major staging layout, where the wgrad
 * kernels' atomics coalesce, and unpacked once per backward into the parameter's own layout). */
typedef struct {
    const float* src;
    void* dst;
    int32_t O, I, R, S;
    int32_t kind;
    int32_t dst_dtype;
} dmu_repack_desc;
int dmu_repack_weights(const dmu_repack_desc* descs_device, int32_t n_desc, int64_t max_numel, dmu_stream_t stream);

/* cudaMemsetAsync(ptr, 0, nbytes) on `stream` (zeroing of accumulation buffers inside a recorded plan). */
int dmu_zero(void* ptr, int64_t nbytes, dmu_stream_t stream);

/* Generic strided copy/cast between two 4-D tensors (tests, layout changes at the boundary). */
int dmu_copy4(const dmu_tensor4* src, const dmu_tensor4* dst, int32_t N, int32_t H, int32_t W, int32_t C, dmu_stream_t stream);

/* Fused Adam + EMA over a flat fp32 arena (SURVEY.md §8 f1;
 * trainers/ddpm_trainer.py:139-143,463-480):  torch.optim.Adam semantics
 * (no amsgrad, L2 weight decay), then ema = decay*ema + (1-decay)*p (ema NULL = skip).
 * `step` (>= 1) sets the bias corrections; when step_device != NULL the step count is read from that device location instead
 * (the launch can then sit in a CUDA graph that is replayed every step).  p / g / m / v / ema may point into the middle of
 * the arenas: the update of one arena range can run while the backward still fills another. */
int dmu_adam_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                 float ema_decay, float grad_scale, const int64_t* step_device, dmu_stream_t stream);

/* ------------------------------------------------------------------ *
 * Input ingest / sample formatting                   *
 * ------------------------------------------------------------------ */

/* datasets/dataset_utils.py:58-61 (T.ToTensor + T.Normalize) and trainers/ddpm_trainer.py:539 (`batch[0].to(device)`),
 * optionally fused with models/ddpm.py:286-296 `_add_noise`:
 *   x0 = (img / 255 - mean[c]) / std[c]          xt = sqrt(acp[t]) x0 + sqrt(1 - acp[t]) noise
 * img: uint8 decoded pixels, [B,H,W,C] when hwc != 0 (PIL / numpy order) else [B,C,H,W]; outputs fp32 [B,C,H,W].
 * mean / std: device pointers to `channels` floats, NULL = skip that step (ToTensor only).  x0_out or xt_out may be
 * NULL (not both); xt_out needs noise, t, alphas_cumprod.  The host->device copy carries 1 byte per value. */
int dmu_ingest_u8(const uint8_t* img, int32_t hwc, const float* mean, const float* std,
                  const float* noise, const int64_t* t, const float* alphas_cumprod,
                  float* x0_out, float* xt_out, int64_t batch, int32_t channels, int64_t hw, dmu_stream_t stream);

/* trainers/ddpm_trainer.py:821-834: torchvision `make_grid(samples, nrow, padding, pad_value)` followed by
 * `save_image`'s quantisation `mul(255).add_(0.5).clamp_(0,255).to(uint8)`, written as [grid_h, grid_w, grid_c] bytes
 * (the array PIL encodes).  Image k (row-major cell order) is read from
 *   x + (k % period) * stride_mod + (k / period) * stride_div        (fp32 [channels, height, width] each, contiguous)
 * so the trainer's "one row per sample, one column per saved denoising step" arrangement of a stacked
 * [steps, B, C, H, W] tensor is period = steps, stride_mod = B*C*H*W, stride_div = C*H*W without the cat/cat copies;
 * a plain [N,C,H,W] batch is period = N, stride_mod = C*H*W.  One image => no border; 1 channel => replicated to 3
 * (both as torchvision does).  dmu_image_grid_shape is host-only arithmetic. */
int dmu_image_grid_shape(int64_t n_images, int32_t channels, int32_t height, int32_t width, int32_t nrow, int32_t padding,
                         int64_t* grid_h, int64_t* grid_w, int32_t* grid_c);
int dmu_image_grid_u8(const float* x, int64_t n_images, int64_t period, int64_t stride_mod, int64_t stride_div,
                      int32_t channels, int32_t height, int32_t width, int32_t nrow, int32_t padding, float pad_value,
                      uint8_t* out, dmu_stream_t stream);
/* scripts/generate.py:119-133 `save_image(samples, nrow=..., normalize=True, value_range=(-1, 1))`: the same grid with
 * every image value first mapped to (clamp(v, lo, hi) - lo) / max(hi - lo, 1e-5) (torchvision's norm_ip; the range is
 * given in double like the Python scalars, the divisor is rounded to float once). */
int dmu_image_grid_range_u8(const float* x, int64_t n_images, int64_t period, int64_t stride_mod, int64_t stride_div,
                            int32_t channels, int32_t height, int32_t width, int32_t nrow, int32_t padding, float pad_value,
                            double range_lo, double range_hi, uint8_t* out, dmu_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DMU_B200_H */
