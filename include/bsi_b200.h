/*
 * bsi_b200 — C ABI of the B200-native (sm_100a) BSI hot path.
 *
 * The reference (martenlienen/bsi) is pure Python/PyTorch and has no FFI of its own; every
 * entry point below replaces a group of ATen call sites of the reference, cited as
 * file:line relative to the reference repository.  The host-side mirror of the reference's
 * Python API (bsi_b200/bsi.py, bsi_b200/models/dit.py) binds these symbols with ctypes —
 * see INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the library never
 *     allocates, frees or retains caller memory beyond the lifetime documented per call;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no device sync),
 *     so they are legal inside CUDA-graph stream capture;
 *   - return value: 0 = ok, negative = bsi_status; bsi_last_error() gives a message for the
 *     calling thread;
 *   - one host thread per device at a time (same as the reference: one process per GPU).
 */
#ifndef BSI_B200_H
#define BSI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bsi_status {
    BSI_OK = 0,
    BSI_ERR_INVALID_ARGUMENT = -1,
    BSI_ERR_CUDA = -2,
    BSI_ERR_UNSUPPORTED = -3,
    BSI_ERR_WORKSPACE = -4,
    BSI_ERR_NOT_READY = -5
} bsi_status;

/* ABI version (bumped on any signature change) and last error text of the calling thread. */
int bsi_abi_version(void);
const char* bsi_last_error(void);
/* Number of CUDA kernels this library has launched in this process (host-side counter). */
long long bsi_launch_counter(void);
/* Compute capability of the current device as major*10+minor, or negative status. */
int bsi_device_arch(void);

/* ------------------------------------------------------------------------------------------
 * Row-coefficient reference.  Per-sample scalars (c_skip, c_out, c_in, ...) are read as
 *     base[ step * step_stride + sample * sample_stride ]
 * where `step` is *step_ptr (a device int, so a captured CUDA graph can be replayed for every
 * sampler step) or 0 when step_ptr is NULL.  sample_stride = 0 broadcasts one value to the
 * whole batch (BSI.sample: same t for every sample, bsi/bsi.py:331).
 * ---------------------------------------------------------------------------------------- */
typedef struct bsi_rowref {
    const float* base;
    int32_t sample_stride;
    int32_t step_stride;
} bsi_rowref;

/* Counter-based noise: Philox4x32-10 keyed by `seed`; element e of global sample s at draw
 * index d uses counter (e/4, d, s_lo, s_hi).  Replaces torch.randn (bsi/bsi.py:325,332,415). */
typedef struct bsi_noise {
    const float* eps;      /* injected N(0,1) noise [n, D] (parity mode), or NULL for in-kernel Philox */
    uint64_t seed;         /* Philox key */
    uint64_t sample_base;  /* global index of local sample 0 (multi-GPU sharding invariance) */
    int32_t draw;          /* draw index; the kernel adds *step_ptr when step_ptr != NULL */
    const uint64_t* key_ptr; /* optional device pointer to {seed, sample_base}: overrides the two by-value fields, so a captured
                              * CUDA graph can be replayed with a new key without re-capture (NULL = use the fields above) */
} bsi_noise;

/* mu0 = rsqrt(lambda_0) * eps  (bsi/bsi.py:325-327).  sigma0 = rsqrt(lambda[0]) read from sigma0_ptr[0]. */
int bsi_sample_init(float* mu, const float* sigma0_ptr, bsi_noise noise, int64_t n, int64_t D, void* stream);

/* One fused sampler step (bsi/bsi.py:331-335 + 381-386):
 *     x_hat = c_skip*mu + c_out*f          (EDM combine; f = denoiser output)
 *     y     = x_hat + sigma*eps            (sigma = rsqrt(alpha_i))
 *     mu'   = (alpha*y + lam*mu) / lam_next
 * coef points to a [steps][8] float table {c_skip, c_out, sigma, alpha, lam, lam_next, 0, 0};
 * row = *step_ptr (or `step` if step_ptr NULL).  mu is updated in place.  Optional outputs
 * x_hat_out / y_out (sample_history, bsi/bsi.py:338-373) may be NULL.  precond = 0 means
 * x_hat = f (preconditioning=None, bsi/bsi.py:378-379). */
int bsi_step_fused(float* mu, const float* f, const float* coef, const int32_t* step_ptr, int32_t step, int32_t precond,
                   bsi_noise noise, float* x_hat_out, float* y_out, int64_t n, int64_t D, void* stream);

/* *step_ptr += 1 (single thread; last node of the captured per-step graph). */
int bsi_step_advance(int32_t* step_ptr, void* stream);

/* x_hat = c_skip*mu + c_out*f with per-row coefficients (BSI._predict_x, bsi/bsi.py:381-386). */
int bsi_edm_combine(float* x_hat, const float* mu, const float* f, bsi_rowref c_skip, bsi_rowref c_out,
                    const int32_t* step_ptr, int64_t n, int64_t D, void* stream);

/* out = scale[row] * in  (model input c_in*mu, bsi/bsi.py:385). */
int bsi_scale_rows(float* out, const float* in, bsi_rowref scale, const int32_t* step_ptr, int64_t n, int64_t D, void* stream);

/* q(mu | x, lambda): mu[r] = gamma[r]*x[r % B] + sigma[r]*eps[r]  for r in [0, R), R = n*B
 * (BSI._sample_q_mu_lambda, bsi/bsi.py:405-420; gamma=(lam-lam0)/lam, sigma=rsqrt(lam)).
 * If model_in != NULL also writes model_in[r] = c_in[r]*mu[r]. */
int bsi_q_sample(float* mu, float* model_in, const float* x, const float* gamma, const float* sigma, const float* c_in,
                 bsi_noise noise, int64_t R, int64_t B, int64_t D, void* stream);

/* Discretization.bucketize (bsi/bsi.py:32-35): idx = clamp(trunc((x - lo_edge) / dx), 0, k-1).
 * lo_edge and dx are the fp32-rounded python floats the reference feeds torch.  out_i64 and/or
 * out_u8 may be NULL. */
int bsi_bucketize(const float* x, int64_t* out_i64, uint8_t* out_u8, float lo_edge, float dx, int32_t k, int64_t numel,
                  void* stream);

/* Discretization.to_8bit_image (bsi/bsi.py:41-48): out = uint8(clamp((x - lo)/(hi - lo)*255, 0, 255)), truncating. */
int bsi_to_uint8(uint8_t* out, const float* x, float lo, float hi, int64_t numel, void* stream);

/* Discretised-Gaussian reconstruction term (BSI.reconstruction_loss, bsi/bsi.py:230-247):
 *   out[r] = - sum_d log clamp(cdf(edge[idx+1]) - cdf(edge[idx]), 1e-20),  r in [0,R), x row r % B,
 *   x_hat = c_skip[r]*mu[r] + c_out[r]*f[r]   (mu == NULL: f already is x_hat),
 *   cdf(v) = 0.5*(1+erf((v - x_hat)*inv_scale/sqrt2)), first/last bin open-ended.
 * edges: k+1 fp32 bin boundaries (torch.linspace, bsi/bsi.py:29-30), k <= 1024. */
int bsi_recon_reduce(float* out, const float* x, const float* mu, const float* f, const float* c_skip, const float* c_out,
                     const float* edges, int32_t k, float lo_edge, float dx, float inv_scale, int64_t R, int64_t B, int64_t D,
                     void* stream);

/* Squared decoding error (bsi/bsi.py:273,288,309): out[r] = sum_d (x[r % B] - x_hat[r])^2. */
int bsi_sqerr_reduce(float* out, const float* x, const float* mu, const float* f, const float* c_skip, const float* c_out,
                     int64_t R, int64_t B, int64_t D, void* stream);

/* Gradient of  sum_r w[r] * out[r]  (out from bsi_sqerr_reduce) w.r.t. the denoiser output f:
 *   grad_f[r] = -2 * w[r] * c_out[r] * (x[r % B] - x_hat[r])    (autograd of bsi/bsi.py:309-310). */
int bsi_sqerr_backward(float* grad_f, const float* w, const float* x, const float* mu, const float* f, const float* c_skip,
                       const float* c_out, int64_t R, int64_t B, int64_t D, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05 + TMEM + TMA), bf16 x bf16 -> fp32 accumulate:
 *     C[b][m][n] = epilogue( sum_k A[b][m][k] * W[b][n][k] )      (nn.Linear: y = x W^T + bias)
 * Replaces every nn.Linear of bsi/models/dit.py:33-34,71-76,79-81,154,163-165.
 * ---------------------------------------------------------------------------------------- */
typedef enum bsi_epilogue {
    BSI_EPI_BIAS_BF16 = 0,       /* out_bf16 = acc + bias                                   */
    BSI_EPI_BIAS_GELU_BF16 = 1,  /* out_bf16 = gelu_tanh(acc + bias)        (dit.py:71-76)  */
    BSI_EPI_BIAS_SILU_BF16 = 2,  /* out_bf16 = silu(acc + bias)             (dit.py:79-81)  */
    BSI_EPI_BIAS_F32 = 3,        /* out_f32  = acc + bias                                   */
    BSI_EPI_GATE_RESID_F32 = 4,  /* out_f32 += gate[row] * (acc + bias)     (dit.py:93-102) */
    BSI_EPI_POS_F32 = 5,         /* out_f32  = acc + bias + pos[m % T]      (dit.py:178)    */
    BSI_EPI_UNPATCH_F32 = 6,     /* out_f32[b,c,y,x] = acc + bias, unpatchified (dit.py:166-172) */
    BSI_EPI_MOD_SILU_BF16 = 7,   /* out_bf16 = silu(shift[b] + (1+scale[b])*(acc+bias))  (residual_block.py:19-21,45-46); conv only */
    /* training path (autograd of dit.py:71-76): the MLP's first Linear keeps its pre-activation for the backward ... */
    BSI_EPI_BIAS_GELU_DUAL_BF16 = 8, /* aux_bf16 = p = bf16(acc + bias);  out_bf16 = gelu_tanh(p)                       */
    /* ... and the data-gradient GEMM of the second Linear applies GELU' while the tile is on chip */
    BSI_EPI_MUL_GELU_GRAD_BF16 = 9   /* out_bf16 = (acc + bias) * gelu_tanh'(aux_bf16[m][n])                            */
} bsi_epilogue;

typedef struct bsi_gemm_args {
    const void* A;     /* bf16 [batch][M][lda] */
    const void* W;     /* bf16 [batch][N][ldw] */
    void* C;           /* bf16 or fp32 [batch][M][ldc] (layout per epilogue) */
    const float* bias; /* [batch][N] or NULL */
    int32_t M, N, K;
    int32_t lda, ldw, ldc;                      /* row pitches in elements; lda, ldw multiples of 8 */
    int32_t batch;                              /* >= 1 */
    int64_t stride_a, stride_w, stride_c, stride_bias; /* elements between batches (stride_a may be 0) */
    int32_t epilogue;                           /* bsi_epilogue */
    /* GATE_RESID: gate for output row m is  gate.base[step*step_stride + (m / rows_per_sample)*sample_stride + n] */
    bsi_rowref gate;
    const int32_t* step_ptr;
    int32_t rows_per_sample; /* tokens per sample T (GATE_RESID, POS, UNPATCH) */
    const float* pos;        /* POS: [T][N] fp32 */
    int32_t patch, grid_w, channels; /* UNPATCH: patch size p, patches per row, output channels; N = p*p*channels */
    void* aux;               /* GELU_DUAL: second output (pre-activation); MUL_GELU_GRAD: pre-activation input; bf16 [M][ldc], batch 1 */
} bsi_gemm_args;

int bsi_gemm_bf16(const bsi_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution (same tcgen05 kernel, A operand gathered by shifted 4-D TMA boxes):
 *     Y[b,y,x,:] = epi( sum_{tap,c} X[b, y+dy(tap), x+dx(tap), c] * W[n][tap*(C1+C2) + c] )
 * 3x3 (taps = 9, zero padding 1) or 1x1 (taps = 1) over NHWC bf16 activations; an optional second source X2
 * supplies channels [C1, C1+C2) — the torch.cat((x, x_skip)) of the U-Net up path without the concat copy.
 * Replaces nn.Conv2d at bsi/nn/residual_block.py:40,44,48, bsi/nn/attention.py:29-30, bsi/models/vdm_unet.py:71.
 * Epilogues: BIAS_BF16, BIAS_F32, MOD_SILU_BF16 (scale/shift per image), GATE_RESID_F32 (Y = resid + gate*(acc+bias),
 * gate.base NULL = 1, resid NULL = Y in place).  C1, C2 multiples of 64; W | 128; H % (128/W) == 0.
 * ---------------------------------------------------------------------------------------- */
typedef struct bsi_conv_args {
    const void* X1;      /* bf16 [B][H][W][C1] */
    const void* X2;      /* bf16 [B][H][W][C2] or NULL */
    const void* W;       /* bf16 [N][taps*(C1+C2)], K index = tap*(C1+C2) + channel, tap = (dy+1)*3 + (dx+1) */
    void* Y;             /* [B*H*W][ldc] bf16 or fp32 (per epilogue) */
    const float* bias;   /* [N] or NULL */
    const float* resid;  /* GATE_RESID: fp32 residual source [B*H*W][ldc], NULL = Y */
    int32_t B, H, Wd, C1, C2, N, taps, ldc;
    int32_t epilogue;
    bsi_rowref scale;    /* MOD_SILU: scale; GATE_RESID: gate (per image) */
    bsi_rowref shift;    /* MOD_SILU: shift */
    const int32_t* step_ptr;
    float* gn_partial;   /* optional (GATE_RESID, N = 128): [B*H*W / 128][32][2] sums and sums of squares of the fp32 output over each
                          * 128-pixel tile and each group of 4 channels -- the GroupNorm statistics of the NEXT layer for free
                          * (nn.GroupNorm(32, .), vdm_unet.py:52; consumed by bsi_groupnorm_apply_bf16) */
} bsi_conv_args;

int bsi_conv_bf16(const bsi_conv_args* args, void* stream);

/* Test hook: force the single-CTA kernel (1) or the CTA-pair / cta_group::2 kernel (2); 0 restores the automatic choice. */
int bsi_gemm_force_cta_group(int32_t cta_group);

/* Measurement aid (bench.py roofline): between begin and end every bsi_gemm_bf16 launch is bracketed by
 * CUDA events on its stream; end synchronises them and returns the summed kernel time, the algorithmic
 * flops (2*M*N*K*batch) and the number of launches.  Not legal while the stream is being captured. */
int bsi_profile_gemm_begin(void);
int bsi_profile_gemm_end(double* total_ms, double* total_flops, int32_t* launches);

/* fp32 -> bf16 cast of a [rows][cols] matrix into pitch `ld_out` (>= cols; padding zero-filled). */
int bsi_cast_bf16(void* out_bf16, const float* in, int64_t rows, int64_t cols, int64_t ld_out, void* stream);

/* LayerNorm (eps, no affine) + adaLN modulate -> bf16 (dit.py:55,66,96,101):
 *   out[m] = LN(x[m]) * (1 + scale[row(m)]) + shift[row(m)],  row(m) = m / rows_per_sample.
 * If gamma/beta != NULL (patch_decoder LayerNorm, dit.py:164) uses out = LN(x)*gamma + beta instead. */
int bsi_layernorm_mod_bf16(void* out_bf16, const float* x, bsi_rowref shift, bsi_rowref scale, const int32_t* step_ptr,
                           const float* gamma, const float* beta, int32_t rows_per_sample, int64_t M, int32_t dim, float eps,
                           void* stream);

/* Building blocks of the fp32-accurate mode (csrc/exact_kernels.cu).
 * split3: fp32 [rows][cols] (pitch ld_in) -> bf16 [rows][3*block_pitch], x = hi + lo: activation layout [hi | lo | hi] (weight_layout 0) or
 * weight layout [hi | hi | lo] (1), so that one bf16 GEMM over K = 3*block_pitch forms hi*hi + lo*hi + hi*lo; act: 0 none, 1 GELU-tanh, 2 SiLU
 * applied before the split.  attention_f32: softmax(Q K^T / sqrt(64)) V on packed fp32 QKV [B*T][3*dim] -> fp32 [B*T][dim] (dit.py:36-47).
 * layernorm_mod_f32 / patch_operand_f32: fp32-output variants of the operand builders (reference-exact angle arithmetic). */
int bsi_split3_bf16(void* out_bf16, const float* in, int64_t rows, int32_t cols, int64_t ld_in, int32_t block_pitch, int32_t weight_layout, int32_t act,
                    void* stream);
int bsi_attention_f32(float* out, const float* qkv, int32_t B, int32_t T, int32_t heads, int32_t head_dim, void* stream);
int bsi_layernorm_mod_f32(float* out_f32, const float* x, bsi_rowref shift, bsi_rowref scale, const int32_t* step_ptr, const float* gamma,
                          const float* beta, int32_t rows_per_sample, int64_t M, int32_t dim, float eps, void* stream);
int bsi_dit_patch_operand_f32(float* A, const float* mu, bsi_rowref scale, const int32_t* step_ptr, int32_t B, int32_t C, int32_t H, int32_t Wd, int32_t patch,
                              int32_t n_min, int32_t n_max, int32_t lda, void* stream);

/* Multi-head attention over packed QKV (dit.py:36-47): qkv bf16 [B*T][3*dim] with columns
 * (qkv, head, channel); out bf16 [B*T][dim] with columns (head, channel).  head_dim = 64. */
int bsi_attention_bf16(void* out_bf16, const void* qkv_bf16, int32_t B, int32_t T, int32_t heads, int32_t head_dim, void* stream);
/* Test hook: T = 256 runs the tcgen05 kernel (scores in TMEM); 1 forces the warp-level mma.sync kernel used for other T. */
int bsi_attention_force_legacy(int32_t on);
/* Development aid: cycles per phase of the tcgen05 attention kernel's timing build (environment BSI_ATT_VARIANT=9), summed over
 * CTAs since the last call: out14[0..5] = softmax warp 0 {wait S, pass 1, pass 2, wait O, epilogue, items},
 * out14[8..12] = control warp {wait Q/K + O free, S latency, wait P half 1, wait P half 2, P V tail latency}. */
int bsi_attention_debug_phases(unsigned long long* out14);
/* Same for the tcgen05 attention backward (BSI_ATT_BWD_VARIANT=9): out16[0..3] = softmax warp 0 {wait S/dP, softmax math, wait
 * accumulators, read-out}, out16[8..11] = control warp {wait inputs, wait P, wait read-out, rest}, out16[15] = items. */
int bsi_attention_backward_debug_phases(unsigned long long* out16);

/* Backward of bsi_attention_bf16 on the same packed layouts (autograd of dit.py:36-47): dqkv bf16 [B*T][3*dim] from the saved qkv,
 * the saved forward output and the upstream gradient dout bf16 [B*T][dim].  lse_ws / dsum_ws: B*heads*T floats of scratch each
 * (log2-sum-exp of the scaled scores and sum_c dout*out per query row; written by the first kernel, read by the second).
 * lse_valid != 0: lse_ws already holds the statistics (saved by bsi_attention_dropout_bf16) and the recomputation pass is skipped. */
int bsi_attention_backward_bf16(void* dqkv_bf16, float* lse_ws, float* dsum_ws, const void* qkv_bf16, const void* out_bf16, const void* dout_bf16,
                                int32_t B, int32_t T, int32_t heads, int32_t head_dim, float drop_p, uint32_t drop_seed, int32_t lse_valid,
                                void* stream);
/* bsi_attention_bf16 that also stores the per-row log2-sum-exp of the scaled scores (lse_out: B*heads*T floats) for the
 * lse_valid fast path of bsi_attention_backward_bf16 (training forward without dropout). */
int bsi_attention_lse_bf16(void* out_bf16, float* lse_out, const void* qkv_bf16, int32_t B, int32_t T, int32_t heads, int32_t head_dim, void* stream);
/* Training-mode attention with dropout on the probabilities (F.scaled_dot_product_attention(dropout_p), dit.py:43-44).  The keep mask
 * is a stateless hash of (drop_seed, head of sample, query, key) -- mix32 in csrc/common.cuh -- which bsi_attention_backward_bf16
 * regenerates from the same (drop_p, drop_seed); drop_p = 0 there means the forward ran without dropout.  lse_out (optional,
 * B*heads*T floats): log2-sum-exp per query row for the backward's lse_valid fast path. */
int bsi_attention_dropout_bf16(void* out_bf16, float* lse_out, const void* qkv_bf16, int32_t B, int32_t T, int32_t heads, int32_t head_dim,
                               float drop_p, uint32_t drop_seed, void* stream);
/* bsi_layernorm_mod_bf16 followed by nn.Dropout(drop_p) on its output (dit.py:101), mask = hash of (drop_seed, row * dim + column). */
int bsi_layernorm_mod_dropout_bf16(void* out_bf16, const float* x, bsi_rowref shift, bsi_rowref scale, int32_t rows_per_sample, int64_t M,
                                   int32_t dim, float eps, float drop_p, uint32_t drop_seed, void* stream);

/* Patch-embed operand (dit.py:149-153,228-231; fourier_features.py:24-36):
 *   A[b*T + tok][(py*p+px)*Cin + c] = bf16( feature_c( scale[b] * mu[b,:,y,x] ) )
 * Cin = C*(1 + 2*(n_max-n_min+1)) when n_max >= n_min else C; feature order: raw channels, then
 * for each channel (n, {sin,cos}).  Pitch lda >= p*p*Cin, padding zero-filled. */
int bsi_dit_patch_operand(void* A_bf16, const float* mu, bsi_rowref scale, const int32_t* step_ptr, int32_t B, int32_t C,
                          int32_t H, int32_t Wd, int32_t patch, int32_t n_min, int32_t n_max, int32_t lda, void* stream);

/* NyquistPositionalEmbedding (pos_emb.py:77-84): out_bf16[r][j] = sin(bias[j] + scale[j]*t[r]); also fp32 copy if out_f32 != NULL. */
int bsi_time_embed(void* out_bf16, float* out_f32, const float* t, const float* scale, const float* bias, int64_t rows,
                   int32_t size, void* stream);

/* ------------------------------------------------------------------------------------------
 * DiT denoiser engine (DenoisingDiT.forward, bsi/models/dit.py:174-181,225-233).
 * All device memory is caller-owned: a parameter arena (packed bf16 weights + fp32 vectors)
 * and a workspace; the engine object itself only holds offsets and shapes (host memory).
 * ---------------------------------------------------------------------------------------- */
typedef struct bsi_dit_config {
    int32_t channels, height, width; /* data_shape */
    int32_t patch, dim, depth, heads;
    int32_t fourier_n_min, fourier_n_max; /* n_max < n_min: no Fourier features */
    int32_t exact; /* 1: fp32-accurate mode -- every GEMM operand split into three bf16 terms, fp32 attention (reference eval precision,
                    * bsi/lightning/plugins.py:7-24); 3-5x slower than the bf16 engine.  0: bf16 tensor-core operands */
} bsi_dit_config;

typedef struct bsi_dit bsi_dit;

int bsi_dit_create(const bsi_dit_config* cfg, bsi_dit** out);
void bsi_dit_destroy(bsi_dit* e);
/* Bytes of the parameter arena / of the workspace for a forward at batch B (cond_rows = rows of the
 * conditioning table: B for per-sample t, k+1 for a tabulated sampler schedule). */
int64_t bsi_dit_param_bytes(const bsi_dit* e);
int64_t bsi_dit_workspace_bytes(const bsi_dit* e, int32_t B);
int64_t bsi_dit_cond_bytes(const bsi_dit* e, int32_t cond_rows);
/* Bind the arena (256-byte aligned).  Must precede bsi_dit_set_param. */
int bsi_dit_bind_params(bsi_dit* e, void* arena, int64_t bytes);
/* Pack one tensor of the reference state_dict (fp32, device) by its reference key, e.g.
 * "dit.blocks.3.attn.to_qkv.weight" (key layout: SURVEY §8 a15).  Also accepts the non-persistent
 * buffers "dit.patch_pos_embedding", "dit.t_embedding.scale", "dit.t_embedding.bias". */
int bsi_dit_set_param(bsi_dit* e, const char* key, const float* src, int64_t numel, void* stream);
/* Number of state_dict tensors still missing (0 = ready). */
int bsi_dit_missing_params(const bsi_dit* e);

/* Conditioning table: cond[layer][row][6*dim] fp32 = adaLN_modulation(t_embedding(t[row])) for all
 * layers (dit.py:79-81,90-92,177).  `scratch` >= bsi_dit_cond_scratch_bytes(rows). */
int64_t bsi_dit_cond_scratch_bytes(const bsi_dit* e, int32_t rows);
int bsi_dit_conditioning(const bsi_dit* e, float* cond, const float* t, int32_t rows, void* scratch, int64_t scratch_bytes,
                         void* stream);

/* out[B,C,H,W] = DiT(in_scale[b] * mu[b], cond rows).  Conditioning row of sample b is
 *   cond_row0 + (*step_ptr or 0) * cond_step_rows + b * cond_sample_rows
 * (per-sample t: cond_sample_rows = 1, cond_step_rows = 0;  sampler: 0 and 1). */
int bsi_dit_forward(const bsi_dit* e, float* out, const float* mu, bsi_rowref in_scale, const float* cond, int32_t cond_rows,
                    int32_t cond_row0, int32_t cond_sample_rows, int32_t cond_step_rows, const int32_t* step_ptr, int32_t B,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* Debug/introspection for parity tests: copy an intermediate of the last forward out of the workspace.
 * what: 0 = residual stream after embed+blocks [B*T][dim] fp32 (valid after forward). */
int bsi_dit_peek(const bsi_dit* e, int32_t what, float* out, int32_t B, const void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * VDM U-Net denoiser (DenoisingVDMUNet.forward, bsi/models/vdm_unet.py:92-100): building blocks and engine.
 * Activations are NHWC; convolutions go through bsi_conv_bf16.
 * ---------------------------------------------------------------------------------------- */
/* act_bf16[b,p,c] = f(GroupNorm(x)[b,p,c]*gamma[c] + beta[c]), f = SiLU if apply_silu (vdm_unet.py:52, residual_block.py:42-43);
 * x fp32 [B][HW][C]; groups of `channels_per_group` consecutive channels; raw_bf16 (optional) receives bf16(x). */
int bsi_groupnorm_act_bf16(void* act_bf16, void* raw_bf16, const float* x, const float* gamma, const float* beta, int32_t B, int32_t HW, int32_t C,
                           int32_t channels_per_group, float eps, int32_t apply_silu, void* stream);
/* The same from statistics the producing convolution left behind (bsi_conv_args.gn_partial: [B*HW/128][32][2]): one streaming pass,
 * x is read once.  C = 128; channels_per_group = 4, or 8 for one 128-channel source of GroupNorm(32, 256) over cat(x, skip). */
int bsi_groupnorm_apply_bf16(void* act_bf16, void* raw_bf16, const float* x, const float* partial, const float* gamma, const float* beta, int32_t B,
                             int32_t HW, int32_t C, int32_t channels_per_group, float eps, int32_t apply_silu, void* stream);
/* bf16 NHWC [B][HW][cpad] = cat(scale*mu, fourier(scale*mu)) zero-padded to cpad channels (vdm_unet.py:95-98); mu fp32 NCHW. */
int bsi_unet_input_bf16(void* out_bf16, const float* mu, bsi_rowref scale, const int32_t* step_ptr, int32_t B, int32_t C, int32_t HW, int32_t n_min,
                        int32_t n_max, int32_t cpad, void* stream);
/* out fp32 NCHW [B][cout][HW] = 1x1 conv of x fp32 NHWC [B][HW][C] with w fp32 [cout][C] (vdm_unet.py:72,100). */
int bsi_unet_decode(float* out, const float* x, const float* w, const float* bias, int32_t B, int32_t HW, int32_t C, int32_t cout, void* stream);
/* fp32 conv weight [N][cin][kh][kw] -> bf16 [N][taps][ctotal], channels written at [c_offset, c_offset+cpad), zero padded beyond cin. */
int bsi_pack_conv_weight(void* out_bf16, const float* w, int32_t N, int32_t cin, int32_t taps, int32_t cpad, int32_t c_offset, int32_t ctotal,
                         void* stream);
/* Single-head attention, head dim 128 (bsi/nn/attention.py:32-41): qkv bf16 [B*T][384] (q|k|v) -> out bf16 [B*T][128]. */
int bsi_attention_d128_bf16(void* out_bf16, const void* qkv_bf16, int32_t B, int32_t T, void* stream);

typedef struct bsi_unet_config {
    int32_t channels, height, width; /* data_shape */
    int32_t dim, levels, heads;      /* dim = 128, heads = 1 implemented */
    int32_t pos_size, pos_mult;      /* NyquistPositionalEmbedding size and pos_emb_mult */
    int32_t fourier_n_min, fourier_n_max;
} bsi_unet_config;
typedef struct bsi_unet bsi_unet;

int bsi_unet_create(const bsi_unet_config* cfg, bsi_unet** out);
void bsi_unet_destroy(bsi_unet* e);
int64_t bsi_unet_param_bytes(const bsi_unet* e);
int64_t bsi_unet_workspace_bytes(const bsi_unet* e, int32_t B);
int64_t bsi_unet_cond_bytes(const bsi_unet* e, int32_t cond_rows);
int64_t bsi_unet_cond_scratch_bytes(const bsi_unet* e, int32_t rows);
int bsi_unet_bind_params(bsi_unet* e, void* arena, int64_t bytes);
/* Keys of the reference state_dict, plus the buffers "pos_emb.scale" / "pos_emb.bias". */
int bsi_unet_set_param(bsi_unet* e, const char* key, const float* src, int64_t numel, void* stream);
int bsi_unet_missing_params(const bsi_unet* e);
/* cond[block][row][2*dim] fp32 = project_onto_scale_shift_block(pos_map(t[row])); blocks ordered down 0..L-1, centre 0, centre 2, up 0..L-1. */
int bsi_unet_conditioning(const bsi_unet* e, float* cond, const float* t, int32_t rows, void* scratch, int64_t scratch_bytes, void* stream);
/* out[B,C,H,W] = UNet(in_scale[b]*mu[b]); conditioning rows addressed like bsi_dit_forward. */
int bsi_unet_forward(const bsi_unet* e, float* out, const float* mu, bsi_rowref in_scale, const float* cond, int32_t cond_rows, int32_t cond_row0,
                     int32_t cond_sample_rows, int32_t cond_step_rows, const int32_t* step_ptr, int32_t B, void* workspace, int64_t workspace_bytes,
                     void* stream);

/* Weight-gradient GEMM (backward of nn.Linear, bsi/models/dit.py:33-34,71-76,79-81; groundwork for config 5):
 *     dW[n][k] += sum_m dY[m][n] * X[m][k]     dY bf16 [M][N] (pitch ldy), X bf16 [M][K] (pitch ldx), dW fp32 [N][K] (pitch ldw)
 * Accumulates into dW (autograd's "+="); the M range is split over `splits` work items per tile (0 = automatic) whose
 * partial sums are combined by TMA reduce-add, so the fp32 summation order across splits is not fixed. */
int bsi_gemm_wgrad_bf16(float* dW, const void* dY_bf16, const void* X_bf16, int32_t M, int32_t N, int32_t K, int32_t ldy, int32_t ldx,
                        int32_t ldw, int32_t splits, void* stream);

/* Elementwise / reduction pieces of the DiT backward (bsi/models/dit.py:50-55,87-103; config 5 groundwork). */
/* x_out[row] = x[row] + gate[row / rows_per_sample] * branch[row]   (torch.addcmul(x, gate, branch); x_out may be x; gate.base == NULL: gate = 1) */
int bsi_gate_residual(float* x_out, const float* x, const void* branch_bf16, bsi_rowref gate, int32_t rows_per_sample, int64_t M, int32_t D, void* stream);
/* bsi_gate_residual followed by bsi_layernorm_mod(_dropout)_bf16 on its result, in one pass over the row (dit.py:93-102):
 *   x_out = x + gate * branch (fp32, may alias x);  out = dropout(LayerNorm(x_out) * (1 + scale) + shift)   [or * gamma + beta] */
int bsi_gate_residual_layernorm_bf16(void* out_bf16, float* x_out, const float* x, const void* branch_bf16, bsi_rowref gate, bsi_rowref shift,
                                     bsi_rowref scale, const float* gamma, const float* beta, int32_t rows_per_sample, int64_t M, int32_t dim, float eps,
                                     float drop_p, uint32_t drop_seed, void* stream);
/* dbranch = gate * dx (bf16);  dgate[b][d] = sum_t dx[b,t,d] * branch[b,t,d];  dbias_part[b][d] = sum_t dbranch[b,t,d]  (either may be NULL) */
int bsi_gate_residual_backward(void* dbranch_bf16, float* dgate, float* dbias_part, const float* dx, const void* branch_bf16, bsi_rowref gate,
                               int32_t rows_per_sample, int32_t B, int32_t D, void* stream);
/* The same with per-CTA partial sums (row-pipelined kernel, D a multiple of 128 up to 1024): a CTA owns rows_per_cta consecutive rows of one sample
 * (rows_per_cta divides rows_per_sample) and writes row blockIdx of dgate_part / dbias_part, both [ceil(M / rows_per_cta)][D]; the caller adds the
 * partial rows of a sample (dgate) or all of them (bias gradient of the branch's last Linear). */
int bsi_gate_residual_backward_rows(void* dbranch_bf16, float* dgate_part, float* dbias_part, const float* dx, const void* branch_bf16, bsi_rowref gate,
                                    int32_t rows_per_sample, int32_t rows_per_cta, int64_t M, int32_t D, void* stream);

/* Finishing pass for per-CTA partial sums (bsi_layernorm_mod_backward, bsi_gate_residual_backward_rows, bsi_colsum_bf16):
 * dst[g][0..D) (+)= sum over r < rows of src[(g * rows + r)][0..D), for up to 12 independent jobs in ONE launch (fixed summation order).
 * Replaces the torch.sum / add_ calls on the partial buffers (autograd of bsi/models/dit.py:87-103: d(shift), d(scale), d(gate) per sample,
 * bias gradients over all rows). */
typedef struct bsi_reduce_job {
    const float* src;    /* [groups * rows][D], contiguous, 16-byte aligned */
    float* dst;          /* [groups] rows of pitch dst_ld */
    int32_t groups, rows, D, dst_ld;
    int32_t accumulate;  /* 1: dst += (gradient accumulation into an existing buffer), 0: dst = */
    int32_t reserved;
} bsi_reduce_job;
int bsi_reduce_rows(const bsi_reduce_job* jobs, int32_t n_jobs, void* stream);
/* fp32 [rows][cols] -> bf16 copy out[rows][ld_out] and bf16 transposed copy out_t[cols][ld_t] in one pass (either may be NULL;
 * padding beyond cols / rows is left untouched: allocate it zeroed).  The transposed weight is the operand of dX = dY W. */
int bsi_cast_transpose_bf16(void* out_bf16, void* out_t_bf16, const float* in, int32_t rows, int32_t cols, int32_t ld_out, int32_t ld_t, void* stream);
/* partial[chunk][n] = sum of a[r][n] over the chunk's rows_per_cta rows (bias gradient = sum over chunks); a bf16 [M][N], pitch ld */
int bsi_colsum_bf16(float* partial, const void* a_bf16, int64_t M, int32_t N, int64_t ld, int32_t rows_per_cta, void* stream);
/* nn.GELU(approximate="tanh") on a bf16 pre-activation, and its backward dpre = dout * gelu'(pre). */
int bsi_gelu_bf16(void* out_bf16, const void* pre_bf16, int64_t numel, void* stream);
int bsi_gelu_backward_bf16(void* dpre_bf16, const void* dout_bf16, const void* pre_bf16, int64_t numel, void* stream);
/* Backward of a = LayerNorm(x) * (1 + scale[b]) + shift[b]  (gamma == NULL)  or  a = LayerNorm(x) * gamma + beta  (gamma != NULL):
 *   dx_io[row] += dL/dx;  dscale_part / dshift_part [ceil(M / rows_per_cta)][dim]: per-CTA partial sums of da*xhat and da over
 *   rows_per_cta consecutive rows (rows_per_cta divides rows_per_sample); the caller adds the partials of a sample.
 *   drop_p > 0: the forward was bsi_layernorm_mod_dropout_bf16 with the same (drop_p, drop_seed). */
int bsi_layernorm_mod_backward(float* dx_io, float* dscale_part, float* dshift_part, const void* da_bf16, const float* x, bsi_rowref scale,
                               const float* gamma, int32_t rows_per_sample, int32_t rows_per_cta, int64_t M, int32_t dim, float eps, float drop_p,
                               uint32_t drop_seed, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer side of the training step over flat fp32 arenas of `numel` elements
 * (numel % 4 == 0).  Replaces, in two launches, Lightning's clip_grad_norm_ (config/train.yaml:40),
 * torch.optim.AdamW(fused=True) (config/task/optimizer/adamw.yaml) and the EMA copy / _foreach_lerp_
 * (bsi/tasks/ema_pytorch.py:316-341,425-434; bsi/tasks/bsi.py:196-198). */

/* out[0] = sum_i grad[i]^2, deterministic (fixed grid, fp64 final pass).
 * workspace: bsi_grad_sumsq_workspace_floats() floats of device scratch. */
int bsi_grad_sumsq(float* out, float* workspace, const float* grad, int64_t numel, void* stream);
int32_t bsi_grad_sumsq_workspace_floats(void);

typedef struct bsi_adamw_args {
    float* param;            /* fp32 weights, updated in place */
    float* grad;             /* fp32 gradients (read; zeroed afterwards if zero_grad) */
    float* exp_avg;          /* AdamW first moment */
    float* exp_avg_sq;       /* AdamW second moment */
    float* ema;              /* EMA weights (ema_mode != 0) */
    void* param_bf16;        /* optional bf16 copy of the updated weights (GEMM operand), or NULL */
    const float* grad_sumsq; /* device scalar from bsi_grad_sumsq (max_norm > 0) */
    int64_t numel;
    int64_t step;            /* optimizer step, counted from 1 (bias correction) */
    double lr, beta1, beta2, eps, weight_decay;
    float max_norm;          /* clip_grad_norm_(max_norm, 2): g *= min(max_norm / (||g|| + 1e-6), 1); <= 0 disables */
    float grad_scale;        /* gradients are multiplied by this first (1/world_size after a sum all-reduce: DDP's mean,
                                bsi/tasks/bsi.py:163-166); 0 means 1 */
    float ema_weight;        /* 1 - current_decay (EMA.get_current_decay, ema_pytorch.py:308-314) */
    int32_t ema_mode;        /* 0 none, 1 ema = param (copy_params_from_model_to_ema), 2 ema.lerp_(param, ema_weight) */
    int32_t zero_grad;       /* 1: leave the gradient arena zeroed (optimizer.zero_grad(set_to_none=False)) */
} bsi_adamw_args;

/* p *= 1 - lr*wd;  m = lerp(m, g, 1-b1);  v = v*b2 + (1-b2)*g*g;
 * p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps);  then the EMA action;  each op rounded to fp32 like
 * torch.optim.adamw._single_tensor_adamw. */
int bsi_adamw_ema_step(const bsi_adamw_args* args, void* stream);

/* EMA on its own: mode 1 ema = param, mode 2 ema.lerp_(param, weight). */
int bsi_ema_update(float* ema, const float* param, int64_t numel, float weight, int32_t mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BSI_B200_H */
