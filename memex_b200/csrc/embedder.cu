// embedder.cu -- mx_embedder: the sentence-embedding forward pass behind memex's SentenceEmbedder.
//
// Replaces the single line `model.encode(&segments)` of the embedder's runner thread (reference
// lib/libmemex/src/llm/embedding.rs:109) -- rust-bert's SentenceEmbeddingsModel::encode, i.e.
// BertModel forward -> masked mean-pool -> L2 normalise -- for already-tokenised input: the host
// side keeps tokenising (embedding.rs:155-198), the device gets padded ids + lengths.
//
// Per layer (post-LN BERT):                                   kernel
//   qkv  = x Wqkv^T + b                                       gemm (EPI_BIAS), fused Q|K|V
//   ctx  = softmax(q k^T / sqrt(dh) + mask) v                 attention
//   x1   = LN(ctx Wo^T + bo + x)                              gemm (EPI_BIAS_RES_LN)
//   h    = gelu(x1 W1^T + b1)                                 gemm (EPI_BIAS_GELU)
//   x    = LN(h W2^T + b2 + x1)                               gemm (EPI_BIAS_RES_LN)
// precision 0 / 2: 16-bit activations (bf16 / f16), GEMMs on tcgen05 with f32 accumulation in TMEM;
// precision 1: everything f32 on CUDA cores (validation path).  No CPU path.
#include <cstring>
#include <map>
#include <string>

#include <cstdlib>

#include <cmath>

#include "common.cuh"
#include "encoder.cuh"
#include "gemm.cuh"

using namespace mx;

namespace {

struct LayerWeights {
    void *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;  // activation dtype, [out, in]
    void *w1g = nullptr;        // T5: wi_1 (the linear half of the gated feed-forward)
    float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
    float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
};

}  // namespace

struct mx_embedder : HandleBase {
    mx_model_cfg cfg{};
    mx_model_ext ext{};
    uint32_t out_dim = 0;       // dense_out when the Dense module is present, else hidden
    int device = 0;
    int sm_count = kNumSMsDefault;
    int act = ACT_BF16;
    bool attention_tc = true;   // MX_ATTENTION_MMA=1 keeps the mma.sync kernel (A/B measurements)
    cudaStream_t stream = nullptr;
    std::vector<void *> allocs;
    float *word = nullptr, *pos = nullptr, *type0 = nullptr, *emb_g = nullptr, *emb_b = nullptr;
    void *proj_w = nullptr;     // ALBERT: [hidden, embed_dim] in the activation dtype
    float *proj_b = nullptr;
    float *dense_w = nullptr, *dense_b = nullptr, *pooled_dev = nullptr;   // Dense module (f32)
    // T5 stack
    float *t5_rel_bias = nullptr;   // [heads][2 max_pos - 1] f32: bias of (key - query), from the bucket table
    float *t5_final_g = nullptr, *t5_zero_bias = nullptr;
    float *xr = nullptr;            // [max_tokens, H] f32 residual stream
    void *hh2 = nullptr;            // [max_tokens, F] second half of the gated feed-forward
    std::vector<LayerWeights> layers;
    // workspaces sized for cfg.max_tokens
    void *x = nullptr, *x1 = nullptr, *qkv = nullptr, *ctx = nullptr, *hh = nullptr;
    int32_t *ids_dev = nullptr, *lens_dev = nullptr;   // lens_dev: [max_seqs] lengths, then [max_seqs + 1] row offsets
    bool packing = true;        // MX_ENCODER_NO_PACKING=1 keeps the padded layout for ragged batches too (A/B measurements)
    float *out_dev = nullptr;
    uint32_t max_seqs = 0;
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    KernelTimer timer;
};

namespace {

template <typename T>
int32_t dev_alloc(mx_embedder *e, T **p, size_t count)
{
    void *q = nullptr;
    cudaError_t err = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
    if (err != cudaSuccess) return fail(e, MX_ERR_SETUP, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(err));
    e->allocs.push_back(q);
    *p = static_cast<T *>(q);
    return MX_OK;
}

struct WeightTable {
    std::map<std::string, const mx_tensor *> by_name;
    const mx_tensor *get(const std::string &name, uint64_t numel, mx_embedder *e, int32_t *rc) const
    {
        auto it = by_name.find(name);
        if (it == by_name.end()) {
            // HF checkpoints of sentence-transformers models may carry a "bert." / "0.auto_model." prefix
            for (const char *pre : {"bert.", "0.auto_model.", "auto_model."}) {
                it = by_name.find(std::string(pre) + name);
                if (it != by_name.end()) break;
            }
        }
        if (it == by_name.end()) {
            *rc = fail(e, MX_ERR_SETUP, "missing weight %s", name.c_str());
            return nullptr;
        }
        if (it->second->numel != numel || !it->second->data) {
            *rc = fail(e, MX_ERR_SETUP, "weight %s has %llu elements, expected %llu", name.c_str(),
                       (unsigned long long)it->second->numel, (unsigned long long)numel);
            return nullptr;
        }
        return it->second;
    }
};

// upload an f32 host tensor; as f32 (dst_act = ACT_F32) or converted to the activation dtype
int32_t upload(mx_embedder *e, const float *host, uint64_t numel, void *dst, int dst_act, float *staging)
{
    if (dst_act == ACT_F32) {
        MX_CUDA(e, MX_ERR_SETUP, cudaMemcpyAsync(dst, host, numel * 4, cudaMemcpyHostToDevice, e->stream));
        return MX_OK;
    }
    MX_CUDA(e, MX_ERR_SETUP, cudaMemcpyAsync(staging, host, numel * 4, cudaMemcpyHostToDevice, e->stream));
    MX_CUDA(e, MX_ERR_SETUP, launch_convert_weight(staging, dst, dst_act, numel, e->stream));
    // staging is reused by the next tensor
    MX_CUDA(e, MX_ERR_SETUP, cudaStreamSynchronize(e->stream));
    return MX_OK;
}

int32_t run_gemm(mx_embedder *e, const void *A, const void *W, const float *bias, const void *residual, const float *g,
                 const float *b, void *out, uint32_t M, uint32_t N, uint32_t K, int epi, cudaStream_t st)
{
    e->timer.begin(st, 0);
    if (e->act == ACT_F32) {
        GemmRefParams p{(const float *)A, (const float *)W, bias, (const float *)residual, g, b, (float *)out, M, N, K,
                        e->cfg.ln_eps};
        MX_CUDA(e, MX_ERR_ENCODE, launch_gemm_ref(p, epi, st));
    } else {
        GemmParams p{};
        p.A = A;
        p.W = W;
        p.bias = bias;
        p.residual = residual;
        p.gamma = g;
        p.beta = b;
        p.out = out;
        p.M = M;
        p.N = N;
        p.K = K;
        p.lda = K;
        p.ldw = K;
        p.ldr = N;
        p.ldo = N;
        p.ln_eps = e->cfg.ln_eps;
        p.fmt = e->act == ACT_BF16 ? 1u : 0u;
        const char *why = nullptr;
        cudaError_t ce = launch_gemm_tc(p, epi, e->sm_count, st, &why);
        if (ce != cudaSuccess)
            return fail(e, MX_ERR_ENCODE, "tcgen05 GEMM [%u x %u x %u] failed: %s (%s)", M, N, K, cudaGetErrorString(ce),
                        why ? why : "");
    }
    e->timer.end(st);
    return MX_OK;
}

// HF T5Attention._relative_position_bucket, bidirectional, with torch's float32 arithmetic (the bucket boundaries are
// decided by its rounding)
int t5_bucket(int rel, int num_buckets, int max_distance)
{
    const int nb = num_buckets / 2;
    const int ret = rel > 0 ? nb : 0;
    const int n = rel < 0 ? -rel : rel;
    const int max_exact = nb / 2;
    if (n < max_exact) return ret + n;
    const float v = logf((float)n / (float)max_exact) / (float)log((double)max_distance / (double)max_exact) * (float)(nb - max_exact);
    const int large = max_exact + (int)v;
    return ret + std::min(large, nb - 1);
}

// weights + workspaces of the T5 stack (HF T5EncoderModel names, include/memex_b200.h)
int32_t load_t5(mx_embedder *e, const WeightTable &tab, float *staging)
{
    const mx_model_cfg &c = e->cfg;
    const mx_model_ext &x = e->ext;
    const uint64_t H = c.hidden, F = c.ffn;
    const int act = e->act;
    const size_t asz = act_size(act);
    int32_t rc = MX_OK;
    auto up_f32 = [&](const std::string &name, uint64_t numel, float **dst) -> int32_t {
        int32_t r = MX_OK;
        const mx_tensor *t = tab.get(name, numel, e, &r);
        if (!t) return r;
        if ((r = dev_alloc(e, dst, numel)) != MX_OK) return r;
        return upload(e, t->data, numel, *dst, ACT_F32, staging);
    };
    auto up_act = [&](const std::string &name, uint64_t numel, void **dst) -> int32_t {
        int32_t r = MX_OK;
        const mx_tensor *t = tab.get(name, numel, e, &r);
        if (!t) return r;
        unsigned char *p = nullptr;
        if ((r = dev_alloc(e, &p, numel * asz)) != MX_OK) return r;
        *dst = p;
        return upload(e, t->data, numel, p, act, staging);
    };
    if (tab.by_name.count("shared.weight") || !tab.by_name.count("encoder.embed_tokens.weight")) {
        if ((rc = up_f32("shared.weight", (uint64_t)c.vocab * H, &e->word)) != MX_OK) return rc;
    } else if ((rc = up_f32("encoder.embed_tokens.weight", (uint64_t)c.vocab * H, &e->word)) != MX_OK) {
        return rc;
    }
    // relative position bias: bucket table [rel_buckets, heads] -> bias of every (key - query) in (-max_pos, max_pos)
    {
        const mx_tensor *t = tab.get("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
                                     (uint64_t)x.rel_buckets * c.heads, e, &rc);
        if (!t) return rc;
        const uint32_t P = c.max_pos, span = 2 * P - 1;
        std::vector<float> table((size_t)c.heads * span);
        for (uint32_t h = 0; h < c.heads; ++h)
            for (uint32_t i = 0; i < span; ++i) {
                const int rel = (int)i - (int)(P - 1);   // key - query
                table[(size_t)h * span + i] = t->data[(size_t)t5_bucket(rel, (int)x.rel_buckets, (int)x.rel_max_distance) * c.heads + h];
            }
        if ((rc = dev_alloc(e, &e->t5_rel_bias, table.size())) != MX_OK) return rc;
        MX_CUDA(e, MX_ERR_SETUP, cudaMemcpyAsync(e->t5_rel_bias, table.data(), table.size() * 4, cudaMemcpyHostToDevice, e->stream));
        MX_CUDA(e, MX_ERR_SETUP, cudaStreamSynchronize(e->stream));
    }
    const uint64_t nz = std::max<uint64_t>(3 * H, F);
    if ((rc = dev_alloc(e, &e->t5_zero_bias, nz)) != MX_OK) return rc;
    MX_CUDA(e, MX_ERR_SETUP, cudaMemsetAsync(e->t5_zero_bias, 0, nz * 4, e->stream));
    if ((rc = up_f32("encoder.final_layer_norm.weight", H, &e->t5_final_g)) != MX_OK) return rc;
    if (x.dense_out) {
        if ((rc = up_f32("dense.linear.weight", (uint64_t)x.dense_out * H, &e->dense_w)) != MX_OK) return rc;
        if (x.dense_bias && (rc = up_f32("dense.linear.bias", x.dense_out, &e->dense_b)) != MX_OK) return rc;
    }
    e->layers.resize(c.layers);
    for (uint32_t l = 0; l < c.layers; ++l) {
        LayerWeights &w = e->layers[l];
        const std::string p = "encoder.block." + std::to_string(l) + ".layer.";
        unsigned char *wqkv = nullptr;
        if ((rc = dev_alloc(e, &wqkv, 3 * H * H * asz)) != MX_OK) return rc;
        w.wqkv = wqkv;
        const char *names[3] = {"q", "k", "v"};
        for (int j = 0; j < 3; ++j) {
            int32_t r2 = MX_OK;
            const mx_tensor *t = tab.get(p + "0.SelfAttention." + names[j] + ".weight", H * H, e, &r2);
            if (!t) return r2;
            if ((rc = upload(e, t->data, H * H, wqkv + (size_t)j * H * H * asz, act, staging)) != MX_OK) return rc;
        }
        if ((rc = up_act(p + "0.SelfAttention.o.weight", H * H, &w.wo)) != MX_OK) return rc;
        if ((rc = up_f32(p + "0.layer_norm.weight", H, &w.ln1_g)) != MX_OK) return rc;
        if ((rc = up_act(p + "1.DenseReluDense.wi_0.weight", F * H, &w.w1)) != MX_OK) return rc;
        if ((rc = up_act(p + "1.DenseReluDense.wi_1.weight", F * H, &w.w1g)) != MX_OK) return rc;
        if ((rc = up_act(p + "1.DenseReluDense.wo.weight", H * F, &w.w2)) != MX_OK) return rc;
        if ((rc = up_f32(p + "1.layer_norm.weight", H, &w.ln2_g)) != MX_OK) return rc;
    }
    const uint64_t T = c.max_tokens;
    unsigned char *ws = nullptr;
    if ((rc = dev_alloc(e, &ws, T * H * asz)) != MX_OK) return rc;
    e->x = ws;
    if ((rc = dev_alloc(e, &ws, T * H * asz)) != MX_OK) return rc;
    e->x1 = ws;
    if ((rc = dev_alloc(e, &ws, T * 3 * H * asz)) != MX_OK) return rc;
    e->qkv = ws;
    if ((rc = dev_alloc(e, &ws, T * H * asz)) != MX_OK) return rc;
    e->ctx = ws;
    if ((rc = dev_alloc(e, &ws, T * F * asz)) != MX_OK) return rc;
    e->hh = ws;
    if ((rc = dev_alloc(e, &ws, T * F * asz)) != MX_OK) return rc;
    e->hh2 = ws;
    if ((rc = dev_alloc(e, &e->xr, T * H)) != MX_OK) return rc;
    e->max_seqs = (uint32_t)std::min<uint64_t>(T, 8192);
    if ((rc = dev_alloc(e, &e->ids_dev, T)) != MX_OK) return rc;
    if ((rc = dev_alloc(e, &e->lens_dev, 2 * (uint64_t)e->max_seqs + 1)) != MX_OK) return rc;
    if ((rc = dev_alloc(e, &e->out_dev, (uint64_t)e->max_seqs * e->out_dim)) != MX_OK) return rc;
    if (x.dense_out && (rc = dev_alloc(e, &e->pooled_dev, (uint64_t)e->max_seqs * H)) != MX_OK) return rc;
    MX_CUDA(e, MX_ERR_SETUP, cudaStreamSynchronize(e->stream));
    return MX_OK;
}

// SentenceT5Base (embedding.rs:32,52): T5 encoder + mean pool + Dense + Normalize.  Padded layout; the residual stream xr
// stays in f32 (pre-norm stacks accumulate into it layer after layer), the GEMMs read the RMS-normalised activations.
int32_t forward_t5(mx_embedder *e, const int32_t *ids_dev, const int32_t *lens_dev, uint32_t B, uint32_t S, float *out_dev,
                   cudaStream_t st)
{
    const mx_model_cfg &c = e->cfg;
    const uint32_t T = B * S, H = c.hidden, F = c.ffn;
    if (S > c.max_pos) return fail(e, MX_ERR_ENCODE, "sequence length %u exceeds max_pos %u", S, c.max_pos);
    int32_t rc;
    e->timer.begin(st, 1);
    MX_CUDA(e, MX_ERR_ENCODE, launch_t5_embed(ids_dev, e->word, e->xr, T, H, c.vocab, st));
    e->timer.end(st);
    const void *delta = nullptr;
    for (uint32_t l = 0; l < c.layers; ++l) {
        const LayerWeights &w = e->layers[l];
        e->timer.begin(st, 1);
        MX_CUDA(e, MX_ERR_ENCODE, launch_t5_add_rmsnorm(e->xr, delta, w.ln1_g, c.ln_eps, e->x, e->act, T, H, st));
        e->timer.end(st);
        if ((rc = run_gemm(e, e->x, w.wqkv, e->t5_zero_bias, nullptr, nullptr, nullptr, e->qkv, T, 3 * H, H, EPI_BIAS, st)) != MX_OK) return rc;
        e->timer.begin(st, 1);
        MX_CUDA(e, MX_ERR_ENCODE, launch_attention_simt(e->qkv, lens_dev, e->ctx, e->act, B, S, H, c.heads, st, 1.0f, e->t5_rel_bias, c.max_pos));
        e->timer.end(st);
        if ((rc = run_gemm(e, e->ctx, w.wo, e->t5_zero_bias, nullptr, nullptr, nullptr, e->x1, T, H, H, EPI_BIAS, st)) != MX_OK) return rc;
        e->timer.begin(st, 1);
        MX_CUDA(e, MX_ERR_ENCODE, launch_t5_add_rmsnorm(e->xr, e->x1, w.ln2_g, c.ln_eps, e->x, e->act, T, H, st));
        e->timer.end(st);
        if ((rc = run_gemm(e, e->x, w.w1, e->t5_zero_bias, nullptr, nullptr, nullptr, e->hh, T, F, H, EPI_BIAS_GELU_TANH, st)) != MX_OK) return rc;
        if ((rc = run_gemm(e, e->x, w.w1g, e->t5_zero_bias, nullptr, nullptr, nullptr, e->hh2, T, F, H, EPI_BIAS, st)) != MX_OK) return rc;
        e->timer.begin(st, 1);
        MX_CUDA(e, MX_ERR_ENCODE, launch_gated_mul(e->hh, e->hh2, e->act, (uint64_t)T * F, st));
        e->timer.end(st);
        if ((rc = run_gemm(e, e->hh, w.w2, e->t5_zero_bias, nullptr, nullptr, nullptr, e->x1, T, H, F, EPI_BIAS, st)) != MX_OK) return rc;
        delta = e->x1;
    }
    e->timer.begin(st, 1);
    MX_CUDA(e, MX_ERR_ENCODE, launch_t5_add_rmsnorm(e->xr, delta, e->t5_final_g, c.ln_eps, e->x, e->act, T, H, st));
    if (e->ext.dense_out) {
        MX_CUDA(e, MX_ERR_ENCODE, launch_pool_normalize(e->x, e->act, lens_dev, e->pooled_dev, B, S, H, 0, nullptr, st));
        MX_CUDA(e, MX_ERR_ENCODE,
                launch_dense_tail(e->pooled_dev, e->dense_w, e->dense_b, out_dev, B, H, e->ext.dense_out, e->ext.dense_act,
                                  c.normalize, st));
    } else {
        MX_CUDA(e, MX_ERR_ENCODE, launch_pool_normalize(e->x, e->act, lens_dev, out_dev, B, S, H, c.normalize, nullptr, st));
    }
    e->timer.end(st);
    return MX_OK;
}

// one forward pass over B sequences of S tokens already in ids_dev / lens_dev
// cu_dev / n_rows: packed layout (encoder.cuh) -- the activations hold only the n_rows real tokens; nullptr = padded
int32_t forward(mx_embedder *e, const int32_t *ids_dev, const int32_t *lens_dev, const int32_t *cu_dev, uint32_t n_rows,
                uint32_t B, uint32_t S, float *out_dev, cudaStream_t st)
{
    const mx_model_cfg &c = e->cfg;
    if (e->ext.family == MX_FAMILY_T5) return forward_t5(e, ids_dev, lens_dev, B, S, out_dev, st);
    const uint32_t T = cu_dev ? n_rows : B * S, H = c.hidden, F = c.ffn;
    const uint32_t E = e->ext.embed_dim ? e->ext.embed_dim : H;
    const int epi_ffn = e->ext.ffn_act == MX_FFN_GELU_TANH ? EPI_BIAS_GELU_TANH : EPI_BIAS_GELU;
    // RoBERTa numbers its positions from padding_idx + 1: token i of a right-padded row reads row i + pos_offset
    const float *pos = e->pos + (size_t)e->ext.pos_offset * E;
    int32_t rc;
    e->timer.begin(st, 1);
    MX_CUDA(e, MX_ERR_ENCODE,
            launch_embed_ln(ids_dev, e->word, pos, e->type0, e->emb_g, e->emb_b, c.ln_eps, E == H ? e->x : e->ctx, e->act, B * S, S,
                            E, c.vocab, cu_dev, st));
    e->timer.end(st);
    // ALBERT: factorised embeddings, projected to the hidden width
    if (E != H && (rc = run_gemm(e, e->ctx, e->proj_w, e->proj_b, nullptr, nullptr, nullptr, e->x, T, H, E, EPI_BIAS, st)) != MX_OK)
        return rc;
    for (uint32_t l = 0; l < c.layers; ++l) {
        const LayerWeights &w = e->layers[e->ext.share_layers ? 0 : l];
        if ((rc = run_gemm(e, e->x, w.wqkv, w.bqkv, nullptr, nullptr, nullptr, e->qkv, T, 3 * H, H, EPI_BIAS, st)) != MX_OK) return rc;
        e->timer.begin(st, 1);
        if (e->act == ACT_F32)
            MX_CUDA(e, MX_ERR_ENCODE, launch_attention_simt(e->qkv, lens_dev, e->ctx, e->act, B, S, H, c.heads, st));

        else if (e->attention_tc && attention_tc_supported(S, H, c.heads))
            MX_CUDA(e, MX_ERR_ENCODE, launch_attention_tc(e->qkv, lens_dev, e->ctx, e->act, B, S, H, c.heads, e->sm_count, cu_dev, n_rows, st));
        else
            MX_CUDA(e, MX_ERR_ENCODE, launch_attention_mma(e->qkv, lens_dev, e->ctx, e->act, B, S, H, c.heads, st));
        e->timer.end(st);
        if ((rc = run_gemm(e, e->ctx, w.wo, w.bo, e->x, w.ln1_g, w.ln1_b, e->x1, T, H, H, EPI_BIAS_RES_LN, st)) != MX_OK) return rc;
        if ((rc = run_gemm(e, e->x1, w.w1, w.b1, nullptr, nullptr, nullptr, e->hh, T, F, H, epi_ffn, st)) != MX_OK) return rc;
        if ((rc = run_gemm(e, e->hh, w.w2, w.b2, e->x1, w.ln2_g, w.ln2_b, e->x, T, H, F, EPI_BIAS_RES_LN, st)) != MX_OK) return rc;
    }
    e->timer.begin(st, 1);
    if (e->ext.dense_out) {
        // Pooling -> Dense -> (Normalize): the order of the sentence-transformers module list
        MX_CUDA(e, MX_ERR_ENCODE, launch_pool_normalize(e->x, e->act, lens_dev, e->pooled_dev, B, S, H, 0, cu_dev, st));
        MX_CUDA(e, MX_ERR_ENCODE,
                launch_dense_tail(e->pooled_dev, e->dense_w, e->dense_b, out_dev, B, H, e->ext.dense_out, e->ext.dense_act,
                                  c.normalize, st));
    } else {
        MX_CUDA(e, MX_ERR_ENCODE, launch_pool_normalize(e->x, e->act, lens_dev, out_dev, B, S, H, c.normalize, cu_dev, st));
    }
    e->timer.end(st);
    return MX_OK;
}

// Uploads the lengths of sequences [b0, b0 + nb) and, when the batch has padding to drop and the tcgen05 attention
// kernel serves the shape, their row offsets: *cu_dev != nullptr selects the packed layout (encoder.cuh) for this chunk.
int32_t stage_lengths(mx_embedder *e, const int32_t *lens, uint32_t nb, uint32_t S, cudaStream_t st, const int32_t **cu_dev,
                      uint32_t *n_rows)
{
    *cu_dev = nullptr;
    *n_rows = nb * S;
    MX_CUDA(e, MX_ERR_ENCODE, cudaMemcpyAsync(e->lens_dev, lens, nb * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (!e->packing || e->act == ACT_F32 || !e->attention_tc || e->ext.family == MX_FAMILY_T5 ||
        !attention_tc_supported(S, e->cfg.hidden, e->cfg.heads))
        return MX_OK;
    std::vector<int32_t> cu(nb + 1);
    cu[0] = 0;
    for (uint32_t b = 0; b < nb; ++b) cu[b + 1] = cu[b] + (int32_t)std::min<uint32_t>((uint32_t)std::max(lens[b], 0), S);
    const uint32_t real = (uint32_t)cu[nb];
    if (real == 0 || real >= nb * S) return MX_OK;   // nothing to drop (or nothing at all: the padded path writes the zeros)
    int32_t *dst = e->lens_dev + e->max_seqs;
    MX_CUDA(e, MX_ERR_ENCODE, cudaMemcpyAsync(dst, cu.data(), (nb + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    *cu_dev = dst;
    *n_rows = real;
    return MX_OK;
}

}  // namespace

extern "C" {

int32_t mx_embedder_create(const mx_model_cfg *cfg, const mx_tensor *weights, uint32_t n_weights, int32_t device,
                           mx_embedder **out)
{
    return mx_embedder_create_ex(cfg, nullptr, weights, n_weights, device, out);
}

int32_t mx_embedder_out_dim(mx_embedder *e, uint32_t *dim)
{
    if (!e || !dim) return MX_ERR_INVALID;
    *dim = e->out_dim;
    return MX_OK;
}

int32_t mx_embedder_create_ex(const mx_model_cfg *cfg, const mx_model_ext *ext_in, const mx_tensor *weights, uint32_t n_weights,
                              int32_t device, mx_embedder **out)
{
    if (!cfg || !out || (!weights && n_weights)) return fail(nullptr, MX_ERR_INVALID, "null argument");
    *out = nullptr;
    mx_model_ext ext{};
    if (ext_in) ext = *ext_in;
    if (ext.embed_dim == cfg->hidden) ext.embed_dim = 0;
    if (ext.pos_offset >= cfg->max_pos) return fail(nullptr, MX_ERR_SETUP, "pos_offset %u >= max_pos %u", ext.pos_offset, cfg->max_pos);
    if (ext.dense_out > 1024) return fail(nullptr, MX_ERR_SETUP, "dense_out %u > 1024", ext.dense_out);
    if (ext.family > MX_FAMILY_T5) return fail(nullptr, MX_ERR_SETUP, "unknown model family %u", ext.family);
    const bool t5 = ext.family == MX_FAMILY_T5;
    if (t5 && (ext.d_kv == 0 || ext.d_kv * cfg->heads != cfg->hidden))
        return fail(nullptr, MX_ERR_SETUP, "T5: heads (%u) x d_kv (%u) must equal hidden (%u)", cfg->heads, ext.d_kv, cfg->hidden);
    if (t5 && (ext.rel_buckets < 4 || ext.rel_buckets % 2 != 0 || ext.rel_buckets > 256 || ext.rel_max_distance < ext.rel_buckets))
        return fail(nullptr, MX_ERR_SETUP, "T5: bad relative attention shape (buckets %u, max distance %u)", ext.rel_buckets, ext.rel_max_distance);
    if (t5 && (ext.embed_dim || ext.share_layers || ext.pos_offset))
        return fail(nullptr, MX_ERR_SETUP, "T5: embed_dim / share_layers / pos_offset do not apply");
    if (ext.dense_act > MX_ACT_TANH || ext.ffn_act > MX_FFN_GELU_TANH)
        return fail(nullptr, MX_ERR_SETUP, "unknown activation code (dense_act %u, ffn_act %u)", ext.dense_act, ext.ffn_act);
    if (ext.embed_dim && (ext.embed_dim > cfg->hidden || ext.embed_dim % 64 != 0))
        return fail(nullptr, MX_ERR_SETUP, "embed_dim %u must be a multiple of 64 and <= hidden", ext.embed_dim);
    if (cfg->layers == 0 || cfg->hidden == 0 || cfg->heads == 0 || cfg->ffn == 0 || cfg->vocab == 0 || cfg->max_pos == 0)
        return fail(nullptr, MX_ERR_SETUP, "model config has a zero dimension");
    if (cfg->hidden % cfg->heads != 0) return fail(nullptr, MX_ERR_SETUP, "hidden %% heads != 0");
    const uint32_t dh = cfg->hidden / cfg->heads;
    if (dh != 32 && dh != 64 && dh != 128) return fail(nullptr, MX_ERR_SETUP, "head_dim %u not in {32, 64, 128}", dh);
    if (cfg->hidden > 1024 || cfg->hidden % 8 != 0 || cfg->ffn % 8 != 0)
        return fail(nullptr, MX_ERR_SETUP, "hidden must be <= 1024 and hidden / ffn multiples of 8");
    if (cfg->precision > 2) return fail(nullptr, MX_ERR_SETUP, "precision must be 0 (bf16), 1 (f32) or 2 (f16)");
    const int act = cfg->precision == 1 ? ACT_F32 : (cfg->precision == 0 ? ACT_BF16 : ACT_F16);
    if (act != ACT_F32) {
        if (!gemm_tc_block_n(3 * cfg->hidden, EPI_BIAS) || !gemm_tc_block_n(cfg->ffn, EPI_BIAS_GELU) ||
            !gemm_tc_block_n(cfg->hidden, t5 ? EPI_BIAS : EPI_BIAS_RES_LN) || (ext.embed_dim && !gemm_tc_block_n(cfg->hidden, EPI_BIAS)))
            return fail(nullptr, MX_ERR_SETUP, "hidden %u / ffn %u have no tcgen05 tile configuration", cfg->hidden, cfg->ffn);
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, MX_ERR_CONNECTION, "no CUDA device: %s (this library has no CPU path)",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, MX_ERR_CONNECTION, "device %d out of range [0, %d)", device, ndev);
    MX_CUDA(nullptr, MX_ERR_CONNECTION, cudaSetDevice(device));
    cudaDeviceProp prop{};
    MX_CUDA(nullptr, MX_ERR_CONNECTION, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, MX_ERR_CONNECTION, "device %d is sm_%d%d; this library ships sm_100a code only", device,
                    prop.major, prop.minor);

    mx_embedder *e = new mx_embedder();
    e->magic = kEmbedderMagic;
    e->cfg = *cfg;
    e->ext = ext;
    e->out_dim = ext.dense_out ? ext.dense_out : cfg->hidden;
    if (e->cfg.max_tokens == 0) e->cfg.max_tokens = 256 * 256;
    if (e->cfg.type_vocab == 0) e->cfg.type_vocab = 2;
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    e->attention_tc = getenv("MX_ATTENTION_MMA") == nullptr;
    e->packing = getenv("MX_ENCODER_NO_PACKING") == nullptr;
    e->act = act;
    auto bail = [&](int32_t rc) {
        g_last_error = e->last_error;
        mx_embedder_destroy(e);
        return rc;
    };
    if ((ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(fail(e, MX_ERR_CONNECTION, "cudaStreamCreate: %s", cudaGetErrorString(ce)));

    WeightTable tab;
    for (uint32_t i = 0; i < n_weights; ++i)
        if (weights[i].name) tab.by_name[weights[i].name] = &weights[i];

    const uint64_t H = cfg->hidden, F = cfg->ffn;
    const uint64_t E = ext.embed_dim ? ext.embed_dim : H;   // width of the embedding tables
    const size_t asz = act_size(act);
    int32_t rc = MX_OK;
    float *staging = nullptr;
    const uint64_t staging_elems = std::max<uint64_t>(3 * H * H, H * F);
    if ((rc = dev_alloc(e, &staging, staging_elems)) != MX_OK) return bail(rc);

    auto up_f32 = [&](const std::string &name, uint64_t numel, float **dst) -> int32_t {
        int32_t r = MX_OK;
        const mx_tensor *t = tab.get(name, numel, e, &r);
        if (!t) return r;
        if ((r = dev_alloc(e, dst, numel)) != MX_OK) return r;
        return upload(e, t->data, numel, *dst, ACT_F32, staging);
    };
    auto up_act = [&](const std::string &name, uint64_t numel, void *dst) -> int32_t {
        int32_t r = MX_OK;
        const mx_tensor *t = tab.get(name, numel, e, &r);
        if (!t) return r;
        return upload(e, t->data, numel, dst, act, staging);
    };

    if (t5) {
        if ((rc = load_t5(e, tab, staging)) != MX_OK) return bail(rc);
        *out = e;
        return MX_OK;
    }
    if ((rc = up_f32("embeddings.word_embeddings.weight", (uint64_t)cfg->vocab * E, &e->word)) != MX_OK) return bail(rc);
    if ((rc = up_f32("embeddings.position_embeddings.weight", (uint64_t)cfg->max_pos * E, &e->pos)) != MX_OK) return bail(rc);
    if (ext.no_token_type) {
        // DistilBERT has no token-type table: the kernel adds a row of zeros
        if ((rc = dev_alloc(e, &e->type0, E)) != MX_OK) return bail(rc);
        if ((ce = cudaMemsetAsync(e->type0, 0, E * sizeof(float), e->stream)) != cudaSuccess)
            return bail(fail(e, MX_ERR_SETUP, "cudaMemset: %s", cudaGetErrorString(ce)));
    } else {
        float *type_all = nullptr;
        if ((rc = up_f32("embeddings.token_type_embeddings.weight", (uint64_t)e->cfg.type_vocab * E, &type_all)) != MX_OK)
            return bail(rc);
        e->type0 = type_all;  // rust-bert / sentence-transformers feed token_type_ids = 0
    }
    if ((rc = up_f32("embeddings.LayerNorm.weight", E, &e->emb_g)) != MX_OK) return bail(rc);
    if ((rc = up_f32("embeddings.LayerNorm.bias", E, &e->emb_b)) != MX_OK) return bail(rc);
    if (E != H) {
        unsigned char *pw = nullptr;
        if ((rc = dev_alloc(e, &pw, H * E * asz)) != MX_OK) return bail(rc);
        e->proj_w = pw;
        if ((rc = up_act("embeddings.projection.weight", H * E, e->proj_w)) != MX_OK) return bail(rc);
        if ((rc = up_f32("embeddings.projection.bias", H, &e->proj_b)) != MX_OK) return bail(rc);
    }
    if (ext.dense_out) {
        if ((rc = up_f32("dense.linear.weight", (uint64_t)ext.dense_out * H, &e->dense_w)) != MX_OK) return bail(rc);
        if (ext.dense_bias && (rc = up_f32("dense.linear.bias", ext.dense_out, &e->dense_b)) != MX_OK) return bail(rc);
    }

    const uint32_t n_layer_sets = ext.share_layers ? 1u : cfg->layers;
    e->layers.resize(n_layer_sets);
    for (uint32_t l = 0; l < n_layer_sets; ++l) {
        LayerWeights &w = e->layers[l];
        const std::string p = "encoder.layer." + std::to_string(l) + ".";
        unsigned char *wqkv = nullptr;
        if ((rc = dev_alloc(e, &wqkv, 3 * H * H * asz)) != MX_OK) return bail(rc);
        w.wqkv = wqkv;
        if ((rc = dev_alloc(e, &w.bqkv, 3 * H)) != MX_OK) return bail(rc);
        const char *qkv_names[3] = {"query", "key", "value"};
        for (int j = 0; j < 3; ++j) {
            if ((rc = up_act(p + "attention.self." + qkv_names[j] + ".weight", H * H, wqkv + (size_t)j * H * H * asz)) != MX_OK)
                return bail(rc);
            int32_t r2 = MX_OK;
            const mx_tensor *bt = tab.get(p + "attention.self." + qkv_names[j] + ".bias", H, e, &r2);
            if (!bt) return bail(r2);
            if ((rc = upload(e, bt->data, H, w.bqkv + (size_t)j * H, ACT_F32, staging)) != MX_OK) return bail(rc);
        }
        unsigned char *tmp = nullptr;
        if ((rc = dev_alloc(e, &tmp, H * H * asz)) != MX_OK) return bail(rc);
        w.wo = tmp;
        if ((rc = up_act(p + "attention.output.dense.weight", H * H, w.wo)) != MX_OK) return bail(rc);
        if ((rc = up_f32(p + "attention.output.dense.bias", H, &w.bo)) != MX_OK) return bail(rc);
        if ((rc = up_f32(p + "attention.output.LayerNorm.weight", H, &w.ln1_g)) != MX_OK) return bail(rc);
        if ((rc = up_f32(p + "attention.output.LayerNorm.bias", H, &w.ln1_b)) != MX_OK) return bail(rc);
        if ((rc = dev_alloc(e, &tmp, F * H * asz)) != MX_OK) return bail(rc);
        w.w1 = tmp;
        if ((rc = up_act(p + "intermediate.dense.weight", F * H, w.w1)) != MX_OK) return bail(rc);
        if ((rc = up_f32(p + "intermediate.dense.bias", F, &w.b1)) != MX_OK) return bail(rc);
        if ((rc = dev_alloc(e, &tmp, H * F * asz)) != MX_OK) return bail(rc);
        w.w2 = tmp;
        if ((rc = up_act(p + "output.dense.weight", H * F, w.w2)) != MX_OK) return bail(rc);
        if ((rc = up_f32(p + "output.dense.bias", H, &w.b2)) != MX_OK) return bail(rc);
        if ((rc = up_f32(p + "output.LayerNorm.weight", H, &w.ln2_g)) != MX_OK) return bail(rc);
        if ((rc = up_f32(p + "output.LayerNorm.bias", H, &w.ln2_b)) != MX_OK) return bail(rc);
    }

    const uint64_t T = e->cfg.max_tokens;
    unsigned char *ws = nullptr;
    if ((rc = dev_alloc(e, &ws, T * H * asz)) != MX_OK) return bail(rc);
    e->x = ws;
    if ((rc = dev_alloc(e, &ws, T * H * asz)) != MX_OK) return bail(rc);
    e->x1 = ws;
    if ((rc = dev_alloc(e, &ws, T * 3 * H * asz)) != MX_OK) return bail(rc);
    e->qkv = ws;
    if ((rc = dev_alloc(e, &ws, T * H * asz)) != MX_OK) return bail(rc);
    e->ctx = ws;
    if ((rc = dev_alloc(e, &ws, T * F * asz)) != MX_OK) return bail(rc);
    e->hh = ws;
    e->max_seqs = (uint32_t)std::min<uint64_t>(T, 8192);
    if ((rc = dev_alloc(e, &e->ids_dev, T)) != MX_OK) return bail(rc);
    if ((rc = dev_alloc(e, &e->lens_dev, 2 * (uint64_t)e->max_seqs + 1)) != MX_OK) return bail(rc);
    if ((rc = dev_alloc(e, &e->out_dev, (uint64_t)e->max_seqs * e->out_dim)) != MX_OK) return bail(rc);
    if (ext.dense_out && (rc = dev_alloc(e, &e->pooled_dev, (uint64_t)e->max_seqs * H)) != MX_OK) return bail(rc);
    if ((ce = cudaStreamSynchronize(e->stream)) != cudaSuccess)
        return bail(fail(e, MX_ERR_SETUP, "weight upload failed: %s", cudaGetErrorString(ce)));
    *out = e;
    return MX_OK;
}

void mx_embedder_destroy(mx_embedder *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (void *p : e->allocs) cudaFree(p);
    if (e->pinned) cudaFreeHost(e->pinned);
    if (e->stream) cudaStreamDestroy(e->stream);
    e->magic = 0;
    delete e;
}

int32_t mx_embedder_encode_device(mx_embedder *e, const int32_t *ids_dev, const int32_t *lens, uint32_t B, uint32_t S,
                                  float *out_dev, void *cuda_stream)
{
    if (!e) return MX_ERR_INVALID;
    if (!ids_dev || !lens || !out_dev) return fail(e, MX_ERR_INVALID, "null buffer");
    if (B == 0) return MX_OK;
    if (S == 0 || S + e->ext.pos_offset > e->cfg.max_pos)
        return fail(e, MX_ERR_ENCODE, "sequence length %u outside [1, max_pos - pos_offset = %u]", S, e->cfg.max_pos - e->ext.pos_offset);
    if (S > e->cfg.max_tokens) return fail(e, MX_ERR_ENCODE, "sequence length %u exceeds the workspace (%u tokens)", S, e->cfg.max_tokens);
    MX_CUDA(e, MX_ERR_CONNECTION, cudaSetDevice(e->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : e->stream;
    const uint32_t chunk = std::min(e->cfg.max_tokens / S, e->max_seqs);
    for (uint32_t b0 = 0; b0 < B; b0 += chunk) {
        const uint32_t nb = std::min(chunk, B - b0);
        const int32_t *cu_dev = nullptr;
        uint32_t n_rows = 0;
        int32_t rc = stage_lengths(e, lens + b0, nb, S, st, &cu_dev, &n_rows);
        if (rc != MX_OK) return rc;
        rc = forward(e, ids_dev + (size_t)b0 * S, e->lens_dev, cu_dev, n_rows, nb, S, out_dev + (size_t)b0 * e->out_dim, st);
        if (rc != MX_OK) return rc;
        if (b0 + chunk < B) MX_CUDA(e, MX_ERR_ENCODE, cudaStreamSynchronize(st));  // lens_dev is reused
    }
    return MX_OK;
}

int32_t mx_embedder_encode(mx_embedder *e, const int32_t *ids, const int32_t *lens, uint32_t B, uint32_t S, float *out)
{
    if (!e) return MX_ERR_INVALID;
    if (!ids || !lens || !out) return fail(e, MX_ERR_INVALID, "null buffer");
    if (B == 0) return MX_OK;
    if (S == 0 || S + e->ext.pos_offset > e->cfg.max_pos)
        return fail(e, MX_ERR_ENCODE, "sequence length %u outside [1, max_pos - pos_offset = %u]", S, e->cfg.max_pos - e->ext.pos_offset);
    if (S > e->cfg.max_tokens) return fail(e, MX_ERR_ENCODE, "sequence length %u exceeds the workspace (%u tokens)", S, e->cfg.max_tokens);
    MX_CUDA(e, MX_ERR_CONNECTION, cudaSetDevice(e->device));
    const uint32_t H = e->out_dim;   // width of one output row
    const uint32_t chunk = std::min(e->cfg.max_tokens / S, e->max_seqs);
    const size_t ids_bytes = (size_t)chunk * S * 4, out_bytes = (size_t)chunk * H * 4;
    const size_t need = ids_bytes + out_bytes;
    if (need > e->pinned_cap) {
        if (e->pinned) cudaFreeHost(e->pinned);
        e->pinned = nullptr;
        e->pinned_cap = 0;
        MX_CUDA(e, MX_ERR_ENCODE, cudaMallocHost(&e->pinned, need));
        e->pinned_cap = need;
    }
    int32_t *pin_ids = (int32_t *)e->pinned;
    float *pin_out = (float *)((char *)e->pinned + ids_bytes);
    for (uint32_t b0 = 0; b0 < B; b0 += chunk) {
        const uint32_t nb = std::min(chunk, B - b0);
        memcpy(pin_ids, ids + (size_t)b0 * S, (size_t)nb * S * 4);
        MX_CUDA(e, MX_ERR_ENCODE, cudaMemcpyAsync(e->ids_dev, pin_ids, (size_t)nb * S * 4, cudaMemcpyHostToDevice, e->stream));
        const int32_t *cu_dev = nullptr;
        uint32_t n_rows = 0;
        int32_t rc = stage_lengths(e, lens + b0, nb, S, e->stream, &cu_dev, &n_rows);
        if (rc != MX_OK) return rc;
        rc = forward(e, e->ids_dev, e->lens_dev, cu_dev, n_rows, nb, S, e->out_dev, e->stream);
        if (rc != MX_OK) return rc;
        MX_CUDA(e, MX_ERR_ENCODE, cudaMemcpyAsync(pin_out, e->out_dev, (size_t)nb * H * 4, cudaMemcpyDeviceToHost, e->stream));
        cudaError_t ce = cudaStreamSynchronize(e->stream);
        if (ce != cudaSuccess) return fail(e, MX_ERR_ENCODE, "encode failed on the device: %s", cudaGetErrorString(ce));
        memcpy(out + (size_t)b0 * H, pin_out, (size_t)nb * H * 4);
    }
    return MX_OK;
}

int32_t mx_embedder_sync(mx_embedder *e)
{
    if (!e) return MX_ERR_INVALID;
    MX_CUDA(e, MX_ERR_CONNECTION, cudaSetDevice(e->device));
    MX_CUDA(e, MX_ERR_ENCODE, cudaStreamSynchronize(e->stream));
    return MX_OK;
}

int32_t mx_embedder_set_sm_limit(mx_embedder *e, uint32_t sms)
{
    if (!e) return MX_ERR_INVALID;
    cudaDeviceProp prop{};
    MX_CUDA(e, MX_ERR_CONNECTION, cudaGetDeviceProperties(&prop, e->device));
    // the persistent kernels size their grids by this count; an even number keeps the 2-CTA clusters whole
    const uint32_t all = (uint32_t)prop.multiProcessorCount;
    e->sm_count = (int)(sms == 0 || sms >= all ? all : std::max<uint32_t>(2u, sms & ~1u));
    return MX_OK;
}

int32_t mx_embedder_set_timing(mx_embedder *e, int32_t on)
{
    if (!e) return MX_ERR_INVALID;
    cudaSetDevice(e->device);
    e->timer.reset();
    e->timer.on = on != 0;
    return MX_OK;
}

int32_t mx_embedder_get_timing(mx_embedder *e, double *gemm_ms_total, uint64_t *gemm_launches, double *other_ms_total,
                               uint64_t *other_launches)
{
    if (!e) return MX_ERR_INVALID;
    cudaSetDevice(e->device);
    e->timer.collect();
    if (gemm_ms_total) *gemm_ms_total = e->timer.total_ms[0];
    if (gemm_launches) *gemm_launches = e->timer.launches[0];
    if (other_ms_total) *other_ms_total = e->timer.total_ms[1];
    if (other_launches) *other_launches = e->timer.launches[1];
    return MX_OK;
}

}  // extern "C"
