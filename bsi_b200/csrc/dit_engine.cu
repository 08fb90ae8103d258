// DiT denoiser engine: host-side orchestration of the sm_100a kernels for
// DenoisingDiT.forward (bsi/models/dit.py:174-181,225-233).  The engine owns no device memory:
// weights live in a caller-provided arena (bf16 matrices + fp32 vectors, packed from the reference
// state_dict by key), activations in a caller-provided workspace.  Every call only enqueues
// kernels on the given stream, so a whole sampler step can be captured in a CUDA graph.
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace bsi {

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct ParamSlot {
    int64_t offset = 0;  // bytes into the arena
    int64_t rows = 0, cols = 0, ld = 0;
    bool bf16 = false;   // matrix packed to bf16 (ld = pitch) or fp32 vector copied verbatim
    bool set = false;
};

}  // namespace bsi

struct bsi_dit {
    bsi_dit_config cfg;
    int T, gh, gw, cin, P, ldp, nout, nfreq;
    int64_t param_bytes = 0;
    uint8_t* arena = nullptr;
    std::map<std::string, bsi::ParamSlot> slots;

    template <typename Tp>
    Tp* ptr(const std::string& key) const {
        return reinterpret_cast<Tp*>(arena + slots.at(key).offset);
    }
    std::string blk(int l, const char* s) const { return "dit.blocks." + std::to_string(l) + "." + s; }
};

namespace bsi {

// In the fp32-accurate mode (cfg.exact) a matrix row holds three bf16 blocks [hi | hi | lo] of pitch ld each (exact_kernels.cu).
static void add_slot(bsi_dit* e, const std::string& key, int64_t rows, int64_t cols, bool bf16, int64_t ld = 0) {
    ParamSlot s;
    s.rows = rows, s.cols = cols, s.bf16 = bf16, s.ld = bf16 ? (ld ? ld : cols) : cols;
    s.offset = e->param_bytes;
    const int64_t terms = (bf16 && e->cfg.exact) ? 3 : 1;
    e->param_bytes = align_up(e->param_bytes + rows * s.ld * terms * (bf16 ? 2 : 4), 256);
    e->slots[key] = s;
}

struct Workspace {
    __nv_bfloat16 *a_patch, *xm, *qkv, *att, *h;
    float* x;
    // fp32-accurate mode: fp32 activations (f_*) and their three-term bf16 splits (s_*)
    float *f_a, *f_qkv, *f_h;
    __nv_bfloat16 *s_a, *s_h;
    int64_t bytes;
};
static Workspace carve(const bsi_dit* e, int B, uint8_t* base) {
    const int64_t M = (int64_t)B * e->T, d = e->cfg.dim;
    Workspace w{};
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        uint8_t* p = base ? base + off : nullptr;
        off = align_up(off + bytes, 1024);
        return p;
    };
    w.x = reinterpret_cast<float*>(take(M * d * 4));
    if (e->cfg.exact) {
        const int64_t wide = e->ldp > d ? e->ldp : d;
        w.f_a = reinterpret_cast<float*>(take(M * wide * 4));              // patch operand / LayerNorm output / attention output
        w.f_qkv = reinterpret_cast<float*>(take(M * 3 * d * 4));
        w.f_h = reinterpret_cast<float*>(take(M * 4 * d * 4));             // MLP pre-activation
        w.s_a = reinterpret_cast<__nv_bfloat16*>(take(M * 3 * wide * 2));  // split of f_a
        w.s_h = reinterpret_cast<__nv_bfloat16*>(take(M * 12 * d * 2));    // split of gelu(f_h)
        w.bytes = off;
        return w;
    }
    w.a_patch = reinterpret_cast<__nv_bfloat16*>(take(M * e->ldp * 2));
    w.xm = reinterpret_cast<__nv_bfloat16*>(take(M * d * 2));
    w.qkv = reinterpret_cast<__nv_bfloat16*>(take(M * 3 * d * 2));
    w.att = reinterpret_cast<__nv_bfloat16*>(take(M * d * 2));
    w.h = reinterpret_cast<__nv_bfloat16*>(take(M * 4 * d * 2));
    w.bytes = off;
    return w;
}

static int gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, const float* bias, int M, int N, int K, int epi,
                cudaStream_t st, const bsi_gemm_args* extra = nullptr) {
    bsi_gemm_args a = extra ? *extra : bsi_gemm_args{};
    a.A = A, a.W = W, a.C = C, a.bias = bias, a.M = M, a.N = N, a.K = K, a.lda = lda, a.ldw = ldw, a.ldc = ldc;
    if (a.batch < 1) a.batch = 1;
    a.epilogue = epi;
    return bsi_gemm_bf16(&a, st);
}

}  // namespace bsi

using namespace bsi;

extern "C" {

int bsi_dit_create(const bsi_dit_config* cfg, bsi_dit** out) {
    BSI_CHECK_ARG(cfg && out, "bsi_dit_create: null argument");
    BSI_CHECK_ARG(cfg->channels > 0 && cfg->patch > 0 && cfg->height % cfg->patch == 0 && cfg->width % cfg->patch == 0,
                  "bsi_dit_create: data shape (%d,%d,%d) not divisible by patch %d", cfg->channels, cfg->height, cfg->width, cfg->patch);
    BSI_CHECK_ARG(cfg->dim % 128 == 0 && cfg->dim <= 2048, "bsi_dit_create: dim=%d must be a multiple of 128 (<= 2048)", cfg->dim);
    BSI_CHECK_ARG(cfg->heads > 0 && cfg->dim == cfg->heads * 64, "bsi_dit_create: only head_dim 64 is implemented (dim=%d heads=%d)",
                  cfg->dim, cfg->heads);
    BSI_CHECK_ARG(cfg->depth > 0, "bsi_dit_create: depth must be positive");
    auto* e = new bsi_dit();
    e->cfg = *cfg;
    e->gh = cfg->height / cfg->patch, e->gw = cfg->width / cfg->patch, e->T = e->gh * e->gw;
    e->nfreq = cfg->fourier_n_max >= cfg->fourier_n_min ? cfg->fourier_n_max - cfg->fourier_n_min + 1 : 0;
    e->cin = cfg->channels * (1 + 2 * e->nfreq);
    e->P = cfg->patch * cfg->patch * e->cin;
    e->ldp = (int)align_up(e->P, 8);
    e->nout = cfg->patch * cfg->patch * cfg->channels;
    if (e->T % 128 != 0 || e->T > 512 || e->nout % 4 != 0) {
        set_error("bsi_dit_create: tokens per sample T=%d must be 128/256/384/512 and patch*patch*channels=%d a multiple of 4", e->T, e->nout);
        delete e;
        return BSI_ERR_UNSUPPORTED;
    }
    const int d = cfg->dim, L = cfg->depth;
    add_slot(e, "dit.patch_encoder.weight", d, e->P, true, e->ldp);
    add_slot(e, "dit.patch_encoder.bias", 1, d, false);
    add_slot(e, "dit.patch_pos_embedding", e->T, d, false);
    add_slot(e, "dit.t_embedding.scale", 1, d, false);
    add_slot(e, "dit.t_embedding.bias", 1, d, false);
    // adaLN first/second Linear of all layers are stored contiguously so that conditioning is two GEMM launches
    for (int l = 0; l < L; ++l) add_slot(e, e->blk(l, "adaLN_modulation.0.weight"), d, d, true);
    for (int l = 0; l < L; ++l) add_slot(e, e->blk(l, "adaLN_modulation.0.bias"), 1, d, false);
    for (int l = 0; l < L; ++l) add_slot(e, e->blk(l, "adaLN_modulation.2.weight"), 6 * d, d, true);
    for (int l = 0; l < L; ++l) add_slot(e, e->blk(l, "adaLN_modulation.2.bias"), 1, 6 * d, false);
    for (int l = 0; l < L; ++l) {
        add_slot(e, e->blk(l, "attn.to_qkv.weight"), 3 * d, d, true);
        add_slot(e, e->blk(l, "attn.to_qkv.bias"), 1, 3 * d, false);
        add_slot(e, e->blk(l, "attn.to_out.weight"), d, d, true);
        add_slot(e, e->blk(l, "attn.to_out.bias"), 1, d, false);
        add_slot(e, e->blk(l, "mlp.0.weight"), 4 * d, d, true);
        add_slot(e, e->blk(l, "mlp.0.bias"), 1, 4 * d, false);
        add_slot(e, e->blk(l, "mlp.2.weight"), d, 4 * d, true);
        add_slot(e, e->blk(l, "mlp.2.bias"), 1, d, false);
    }
    add_slot(e, "dit.patch_decoder.0.weight", 1, d, false);
    add_slot(e, "dit.patch_decoder.0.bias", 1, d, false);
    add_slot(e, "dit.patch_decoder.1.weight", e->nout, d, true);
    add_slot(e, "dit.patch_decoder.1.bias", 1, e->nout, false);
    *out = e;
    return BSI_OK;
}

void bsi_dit_destroy(bsi_dit* e) { delete e; }

int64_t bsi_dit_param_bytes(const bsi_dit* e) { return e ? e->param_bytes : 0; }
int64_t bsi_dit_workspace_bytes(const bsi_dit* e, int32_t B) { return e && B > 0 ? carve(e, B, nullptr).bytes : 0; }
int64_t bsi_dit_cond_bytes(const bsi_dit* e, int32_t rows) {
    return e && rows > 0 ? (int64_t)e->cfg.depth * rows * 6 * e->cfg.dim * 4 : 0;
}
int64_t bsi_dit_cond_scratch_bytes(const bsi_dit* e, int32_t rows) {
    if (!e || rows <= 0) return 0;
    const int64_t d = e->cfg.dim, L = e->cfg.depth;
    if (e->cfg.exact)  // fp32 embedding, its split, fp32 hidden layer of all blocks, its split
        return align_up(rows * d * 4, 1024) + align_up(rows * 3 * d * 2, 1024) + align_up(rows * L * d * 4, 1024) + align_up(rows * L * 3 * d * 2, 1024);
    return align_up(rows * d * 2, 1024) + align_up(rows * L * d * 2, 1024);
}

int bsi_dit_bind_params(bsi_dit* e, void* arena, int64_t bytes) {
    BSI_CHECK_ARG(e && arena, "bsi_dit_bind_params: null argument");
    BSI_CHECK_ARG(bytes >= e->param_bytes, "parameter arena too small: %lld < %lld", (long long)bytes, (long long)e->param_bytes);
    BSI_CHECK_ARG((reinterpret_cast<uintptr_t>(arena) & 255) == 0, "parameter arena must be 256-byte aligned");
    e->arena = reinterpret_cast<uint8_t*>(arena);
    for (auto& kv : e->slots) kv.second.set = false;
    return BSI_OK;
}

int bsi_dit_set_param(bsi_dit* e, const char* key, const float* src, int64_t numel, void* stream) {
    BSI_CHECK_ARG(e && key && src, "bsi_dit_set_param: null argument");
    BSI_CHECK_ARG(e->arena, "bsi_dit_set_param: bind the parameter arena first");
    auto it = e->slots.find(key);
    BSI_CHECK_ARG(it != e->slots.end(), "bsi_dit_set_param: unknown state_dict key '%s'", key);
    ParamSlot& s = it->second;
    BSI_CHECK_ARG(numel == s.rows * s.cols, "bsi_dit_set_param: '%s' has %lld elements, expected %lld", key, (long long)numel,
                  (long long)(s.rows * s.cols));
    if (s.bf16 && e->cfg.exact) {
        int rc = bsi_split3_bf16(e->arena + s.offset, src, s.rows, (int32_t)s.cols, s.cols, (int32_t)s.ld, /*weight layout*/ 1, 0, stream);
        if (rc != BSI_OK) return rc;
    } else if (s.bf16) {
        int rc = bsi_cast_bf16(e->arena + s.offset, src, s.rows, s.cols, s.ld, stream);
        if (rc != BSI_OK) return rc;
    } else {
        BSI_CUDA_OK(cudaMemcpyAsync(e->arena + s.offset, src, numel * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    s.set = true;
    return BSI_OK;
}

int bsi_dit_missing_params(const bsi_dit* e) {
    if (!e) return -1;
    int n = 0;
    for (auto& kv : e->slots) n += kv.second.set ? 0 : 1;
    return n;
}

int bsi_dit_conditioning(const bsi_dit* e, float* cond, const float* t, int32_t rows, void* scratch, int64_t scratch_bytes,
                         void* stream) {
    BSI_CHECK_ARG(e && cond && t && scratch && rows > 0, "bsi_dit_conditioning: bad arguments");
    if (bsi_dit_missing_params(e) != 0) {
        set_error("bsi_dit_conditioning: %d parameters not set", bsi_dit_missing_params(e));
        return BSI_ERR_NOT_READY;
    }
    if (scratch_bytes < bsi_dit_cond_scratch_bytes(e, rows)) {
        set_error("conditioning scratch too small");
        return BSI_ERR_WORKSPACE;
    }
    const int d = e->cfg.dim, L = e->cfg.depth;
    cudaStream_t st = (cudaStream_t)stream;
    if (e->cfg.exact) {
        // same chain with every operand split into three bf16 terms (exact_kernels.cu): fp32-level accuracy from the tensor cores
        uint8_t* sp = reinterpret_cast<uint8_t*>(scratch);
        float* c32 = reinterpret_cast<float*>(sp);
        sp += align_up((int64_t)rows * d * 4, 1024);
        auto* c3 = reinterpret_cast<__nv_bfloat16*>(sp);
        sp += align_up((int64_t)rows * 3 * d * 2, 1024);
        float* h32 = reinterpret_cast<float*>(sp);
        sp += align_up((int64_t)rows * L * d * 4, 1024);
        auto* h3 = reinterpret_cast<__nv_bfloat16*>(sp);
        int rc = bsi_time_embed(nullptr, c32, t, e->ptr<float>("dit.t_embedding.scale"), e->ptr<float>("dit.t_embedding.bias"), rows, d, st);
        if (rc != BSI_OK) return rc;
        if ((rc = bsi_split3_bf16(c3, c32, rows, d, d, d, 0, 0, st)) != BSI_OK) return rc;
        // the first adaLN Linear of all layers is one stacked [L*d][3d] matrix (slots are contiguous: d rows of 3d each per layer)
        rc = gemm(c3, 3 * d, e->ptr<void>(e->blk(0, "adaLN_modulation.0.weight")), 3 * d, h32, L * d, e->ptr<float>(e->blk(0, "adaLN_modulation.0.bias")), rows,
                  L * d, 3 * d, BSI_EPI_BIAS_F32, st);
        if (rc != BSI_OK) return rc;
        // SiLU + split, viewed as [rows * L][d] -> [rows * L][3d]
        if ((rc = bsi_split3_bf16(h3, h32, (int64_t)rows * L, d, d, d, 0, /*silu*/ 2, st)) != BSI_OK) return rc;
        bsi_gemm_args x{};
        x.batch = L;
        x.stride_a = 3 * d, x.stride_w = (int64_t)6 * d * 3 * d, x.stride_c = (int64_t)rows * 6 * d, x.stride_bias = 6 * d;
        return gemm(h3, L * 3 * d, e->ptr<void>(e->blk(0, "adaLN_modulation.2.weight")), 3 * d, cond, 6 * d, e->ptr<float>(e->blk(0, "adaLN_modulation.2.bias")),
                    rows, 6 * d, 3 * d, BSI_EPI_BIAS_F32, st, &x);
    }
    auto* c16 = reinterpret_cast<__nv_bfloat16*>(scratch);
    auto* h16 = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(scratch) + align_up((int64_t)rows * d * 2, 1024));
    // c = t_embedding(t)                                                   (dit.py:177)
    int rc = bsi_time_embed(c16, nullptr, t, e->ptr<float>("dit.t_embedding.scale"), e->ptr<float>("dit.t_embedding.bias"), rows, d, st);
    if (rc != BSI_OK) return rc;
    // h[:, l*d:(l+1)*d] = SiLU(c W0_l^T + b0_l) for all layers in one GEMM   (dit.py:79-80)
    rc = gemm(c16, d, e->ptr<void>(e->blk(0, "adaLN_modulation.0.weight")), d, h16, L * d, e->ptr<float>(e->blk(0, "adaLN_modulation.0.bias")),
              rows, L * d, d, BSI_EPI_BIAS_SILU_BF16, st);
    if (rc != BSI_OK) return rc;
    // cond[l] = h_l W2_l^T + b2_l, batched over layers                        (dit.py:81)
    bsi_gemm_args x{};
    x.batch = L;
    x.stride_a = d, x.stride_w = (int64_t)6 * d * d, x.stride_c = (int64_t)rows * 6 * d, x.stride_bias = 6 * d;
    return gemm(h16, L * d, e->ptr<void>(e->blk(0, "adaLN_modulation.2.weight")), d, cond, 6 * d,
                e->ptr<float>(e->blk(0, "adaLN_modulation.2.bias")), rows, 6 * d, d, BSI_EPI_BIAS_F32, st, &x);
}

int bsi_dit_forward(const bsi_dit* e, float* out, const float* mu, bsi_rowref in_scale, const float* cond, int32_t cond_rows,
                    int32_t cond_row0, int32_t cond_sample_rows, int32_t cond_step_rows, const int32_t* step_ptr, int32_t B,
                    void* workspace, int64_t workspace_bytes, void* stream) {
    BSI_CHECK_ARG(e && out && mu && in_scale.base && cond && workspace && B > 0 && cond_rows > 0, "bsi_dit_forward: bad arguments");
    // the host-known part of the conditioning row index must stay inside the table (the step part, read on the device, is the
    // caller's contract: *step_ptr * cond_step_rows + that index < cond_rows)
    BSI_CHECK_ARG(cond_row0 >= 0 && cond_sample_rows >= 0 && cond_step_rows >= 0 &&
                      (int64_t)cond_row0 + (int64_t)(B - 1) * cond_sample_rows < cond_rows,
                  "%s: conditioning rows [%d + b*%d, b < %d] exceed the table of %d rows", "bsi_dit_forward", cond_row0, cond_sample_rows, B, cond_rows);
    if (bsi_dit_missing_params(e) != 0) {
        set_error("bsi_dit_forward: %d parameters not set", bsi_dit_missing_params(e));
        return BSI_ERR_NOT_READY;
    }
    Workspace w = carve(e, B, reinterpret_cast<uint8_t*>(workspace));
    if (workspace_bytes < w.bytes) {
        set_error("workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)w.bytes);
        return BSI_ERR_WORKSPACE;
    }
    const bsi_dit_config& c = e->cfg;
    const int d = c.dim, T = e->T, M = B * T;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    // BSI_DEBUG_SKIP=ln|attn|ln,attn: timing experiments only (wrong results) -- which share of a step is a kernel family worth
    // under the power cap?  (profiles/skip_experiment_r02.txt)
    static const char* dbg = getenv("BSI_DEBUG_SKIP");
    const bool skip_ln = dbg && strstr(dbg, "ln"), skip_attn = dbg && strstr(dbg, "attn");
#define BSI_TRY(call) \
    if ((rc = (call)) != BSI_OK) return rc

    if (c.exact) {
        // ---- fp32-accurate mode: fp32 activations, three-term bf16 splits as GEMM operands (K three times as long), fp32 attention
        const int P3 = 3 * e->ldp, d3 = 3 * d;
        BSI_TRY(bsi_dit_patch_operand_f32(w.f_a, mu, in_scale, step_ptr, B, c.channels, c.height, c.width, c.patch, c.fourier_n_min, c.fourier_n_max, e->P, st));
        BSI_TRY(bsi_split3_bf16(w.s_a, w.f_a, M, e->P, e->P, e->ldp, 0, 0, st));
        {
            bsi_gemm_args x{};
            x.rows_per_sample = T, x.pos = e->ptr<float>("dit.patch_pos_embedding");
            BSI_TRY(gemm(w.s_a, P3, e->ptr<void>("dit.patch_encoder.weight"), P3, w.x, d, e->ptr<float>("dit.patch_encoder.bias"), M, d, P3, BSI_EPI_POS_F32, st, &x));
        }
        for (int l = 0; l < c.depth; ++l) {
            const float* cl = cond + ((int64_t)l * cond_rows + cond_row0) * 6 * d;
            auto part = [&](int p) {
                bsi_rowref r;
                r.base = cl + (int64_t)p * d, r.sample_stride = cond_sample_rows * 6 * d, r.step_stride = cond_step_rows * 6 * d;
                return r;
            };
            bsi_gemm_args g{};
            g.rows_per_sample = T, g.step_ptr = step_ptr;
            BSI_TRY(bsi_layernorm_mod_f32(w.f_a, w.x, part(0), part(1), step_ptr, nullptr, nullptr, T, M, d, 1e-5f, st));
            BSI_TRY(bsi_split3_bf16(w.s_a, w.f_a, M, d, d, d, 0, 0, st));
            BSI_TRY(gemm(w.s_a, d3, e->ptr<void>(e->blk(l, "attn.to_qkv.weight")), d3, w.f_qkv, 3 * d, e->ptr<float>(e->blk(l, "attn.to_qkv.bias")), M, 3 * d, d3,
                         BSI_EPI_BIAS_F32, st));
            BSI_TRY(bsi_attention_f32(w.f_a, w.f_qkv, B, T, c.heads, d / c.heads, st));
            BSI_TRY(bsi_split3_bf16(w.s_a, w.f_a, M, d, d, d, 0, 0, st));
            g.gate = part(2);
            BSI_TRY(gemm(w.s_a, d3, e->ptr<void>(e->blk(l, "attn.to_out.weight")), d3, w.x, d, e->ptr<float>(e->blk(l, "attn.to_out.bias")), M, d, d3,
                         BSI_EPI_GATE_RESID_F32, st, &g));
            BSI_TRY(bsi_layernorm_mod_f32(w.f_a, w.x, part(3), part(4), step_ptr, nullptr, nullptr, T, M, d, 1e-5f, st));
            BSI_TRY(bsi_split3_bf16(w.s_a, w.f_a, M, d, d, d, 0, 0, st));
            BSI_TRY(gemm(w.s_a, d3, e->ptr<void>(e->blk(l, "mlp.0.weight")), d3, w.f_h, 4 * d, e->ptr<float>(e->blk(l, "mlp.0.bias")), M, 4 * d, d3, BSI_EPI_BIAS_F32,
                         st));
            BSI_TRY(bsi_split3_bf16(w.s_h, w.f_h, M, 4 * d, 4 * d, 4 * d, 0, /*gelu*/ 1, st));
            g.gate = part(5);
            BSI_TRY(gemm(w.s_h, 12 * d, e->ptr<void>(e->blk(l, "mlp.2.weight")), 12 * d, w.x, d, e->ptr<float>(e->blk(l, "mlp.2.bias")), M, d, 12 * d,
                         BSI_EPI_GATE_RESID_F32, st, &g));
        }
        bsi_rowref none{};
        BSI_TRY(bsi_layernorm_mod_f32(w.f_a, w.x, none, none, nullptr, e->ptr<float>("dit.patch_decoder.0.weight"), e->ptr<float>("dit.patch_decoder.0.bias"), T, M, d,
                                      1e-5f, st));
        BSI_TRY(bsi_split3_bf16(w.s_a, w.f_a, M, d, d, d, 0, 0, st));
        bsi_gemm_args x{};
        x.rows_per_sample = T, x.patch = c.patch, x.grid_w = e->gw, x.channels = c.channels;
        BSI_TRY(gemm(w.s_a, d3, e->ptr<void>("dit.patch_decoder.1.weight"), d3, out, e->nout, e->ptr<float>("dit.patch_decoder.1.bias"), M, e->nout, d3,
                     BSI_EPI_UNPATCH_F32, st, &x));
        return BSI_OK;
    }
    // patchify + Fourier features + c_in scaling -> bf16 operand; then patch_encoder + positional table
    BSI_TRY(bsi_dit_patch_operand(w.a_patch, mu, in_scale, step_ptr, B, c.channels, c.height, c.width, c.patch, c.fourier_n_min,
                                  c.fourier_n_max, e->ldp, st));
    {
        bsi_gemm_args x{};
        x.rows_per_sample = T, x.pos = e->ptr<float>("dit.patch_pos_embedding");
        BSI_TRY(gemm(w.a_patch, e->ldp, e->ptr<void>("dit.patch_encoder.weight"), e->ldp, w.x, d, e->ptr<float>("dit.patch_encoder.bias"),
                     M, d, e->P, BSI_EPI_POS_F32, st, &x));
    }
    for (int l = 0; l < c.depth; ++l) {
        const float* cl = cond + ((int64_t)l * cond_rows + cond_row0) * 6 * d;
        auto part = [&](int p) {
            bsi_rowref r;
            r.base = cl + (int64_t)p * d, r.sample_stride = cond_sample_rows * 6 * d, r.step_stride = cond_step_rows * 6 * d;
            return r;
        };
        bsi_gemm_args g{};
        g.rows_per_sample = T, g.step_ptr = step_ptr;
        // attention branch: x += gate_msa * to_out(attn(to_qkv(modulate(norm(x), shift_msa, scale_msa))))   (dit.py:93-97)
        if (!skip_ln) BSI_TRY(bsi_layernorm_mod_bf16(w.xm, w.x, part(0), part(1), step_ptr, nullptr, nullptr, T, M, d, 1e-5f, st));
        BSI_TRY(gemm(w.xm, d, e->ptr<void>(e->blk(l, "attn.to_qkv.weight")), d, w.qkv, 3 * d, e->ptr<float>(e->blk(l, "attn.to_qkv.bias")), M,
                     3 * d, d, BSI_EPI_BIAS_BF16, st));
        if (!skip_attn) BSI_TRY(bsi_attention_bf16(w.att, w.qkv, B, T, c.heads, d / c.heads, st));
        g.gate = part(2);
        BSI_TRY(gemm(w.att, d, e->ptr<void>(e->blk(l, "attn.to_out.weight")), d, w.x, d, e->ptr<float>(e->blk(l, "attn.to_out.bias")), M, d, d,
                     BSI_EPI_GATE_RESID_F32, st, &g));
        // MLP branch: x += gate_mlp * mlp(modulate(norm(x), shift_mlp, scale_mlp))                          (dit.py:98-102)
        if (!skip_ln) BSI_TRY(bsi_layernorm_mod_bf16(w.xm, w.x, part(3), part(4), step_ptr, nullptr, nullptr, T, M, d, 1e-5f, st));
        BSI_TRY(gemm(w.xm, d, e->ptr<void>(e->blk(l, "mlp.0.weight")), d, w.h, 4 * d, e->ptr<float>(e->blk(l, "mlp.0.bias")), M, 4 * d, d,
                     BSI_EPI_BIAS_GELU_BF16, st));
        g.gate = part(5);
        BSI_TRY(gemm(w.h, 4 * d, e->ptr<void>(e->blk(l, "mlp.2.weight")), 4 * d, w.x, d, e->ptr<float>(e->blk(l, "mlp.2.bias")), M, d, 4 * d,
                     BSI_EPI_GATE_RESID_F32, st, &g));
    }
    // patch_decoder: affine LayerNorm -> Linear -> unpatchify                                               (dit.py:163-172,181)
    bsi_rowref none{};
    BSI_TRY(bsi_layernorm_mod_bf16(w.xm, w.x, none, none, nullptr, e->ptr<float>("dit.patch_decoder.0.weight"),
                                   e->ptr<float>("dit.patch_decoder.0.bias"), T, M, d, 1e-5f, st));
    {
        bsi_gemm_args x{};
        x.rows_per_sample = T, x.patch = c.patch, x.grid_w = e->gw, x.channels = c.channels;
        BSI_TRY(gemm(w.xm, d, e->ptr<void>("dit.patch_decoder.1.weight"), d, out, e->nout, e->ptr<float>("dit.patch_decoder.1.bias"), M,
                     e->nout, d, BSI_EPI_UNPATCH_F32, st, &x));
    }
#undef BSI_TRY
    return BSI_OK;
}

int bsi_dit_peek(const bsi_dit* e, int32_t what, float* out, int32_t B, const void* workspace, void* stream) {
    BSI_CHECK_ARG(e && out && workspace && B > 0, "bsi_dit_peek: bad arguments");
    Workspace w = carve(e, B, const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(workspace)));
    if (what == 0) {
        BSI_CUDA_OK(cudaMemcpyAsync(out, w.x, (int64_t)B * e->T * e->cfg.dim * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return BSI_OK;
    }
    set_error("bsi_dit_peek: unknown selector %d", what);
    return BSI_ERR_INVALID_ARGUMENT;
}

}  // extern "C"
