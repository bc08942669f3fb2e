// embedder.cpp -- B200Encoder, SentenceEmbedder (the actor of reference llm/embedding.rs:77-152), model table and
// the .safetensors reader.  See memex_host.hpp.
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>

#include "../../include/memex_b200.h"
#include "json.hpp"
#include "memex_host.hpp"

namespace memex {

EmbeddingError::EmbeddingError(EmbeddingErrorKind k, const std::string &msg)
    : std::runtime_error(std::string(k == EmbeddingErrorKind::EncodingFailure ? "Failed to encode string: " : "Unable to load model: ") + msg),
      kind(k)
{
}

std::optional<Architecture> architecture_of(EmbeddingsModelType model)
{
    // shapes from the sentence-transformers model cards; the reference can only segment L12 / L6 / distilroberta
    // (embedding.rs:156-161) but encodes single texts with any of them
    switch (model) {
        case EmbeddingsModelType::AllMiniLmL6V2: return Architecture{6, 384, 12, 1536, 30522, 512, 2, 1e-12f, true, 256};
        case EmbeddingsModelType::AllMiniLmL12V2: return Architecture{12, 384, 12, 1536, 30522, 512, 2, 1e-12f, true, 128};
        case EmbeddingsModelType::BertBaseNliMeanTokens: return Architecture{12, 768, 12, 3072, 30522, 512, 2, 1e-12f, false, 128};
        case EmbeddingsModelType::AllDistilrobertaV1: {
            Architecture a{6, 768, 12, 3072, 50265, 514, 1, 1e-5f, true, 512};
            a.family = Family::Roberta;
            a.pos_offset = 2;
            a.pad_id = 1;
            return a;
        }
        case EmbeddingsModelType::DistiluseBaseMultilingualCased: {
            Architecture a{6, 768, 12, 3072, 119547, 512, 0, 1e-12f, false, 128};
            a.family = Family::DistilBert;
            a.dense_out = 512;
            a.dense_tanh = true;
            return a;
        }
        case EmbeddingsModelType::ParaphraseAlbertSmallV2: {
            Architecture a{6, 768, 12, 3072, 30000, 512, 2, 1e-12f, false, 100};
            a.family = Family::Albert;
            a.ffn_gelu_new = true;
            a.embed_dim = 128;
            a.share_layers = true;
            return a;
        }
        case EmbeddingsModelType::SentenceT5Base: {
            // sentence-transformers/sentence-t5-base: T5 v1.1 base encoder + mean pool + Dense(768 -> 768, no bias) + Normalize
            Architecture a{12, 768, 12, 2048, 32128, 512, 0, 1e-6f, true, 256};
            a.family = Family::T5;
            a.d_kv = 64;
            a.dense_out = 768;
            a.dense_bias = false;
            a.ffn_gelu_new = true;   // gated: gelu_new(wi_0 x) * wi_1 x
            return a;
        }
        default: return std::nullopt;
    }
}

namespace {
void replace_all(std::string &s, const std::string &a, const std::string &b)
{
    for (size_t pos = 0; (pos = s.find(a, pos)) != std::string::npos; pos += b.size()) s.replace(pos, a.size(), b);
}
}  // namespace

void Weights::canonicalize(Family family)
{
    static const char *const kDistil[][2] = {
        {"transformer.layer.", "encoder.layer."}, {".attention.q_lin.", ".attention.self.query."},
        {".attention.k_lin.", ".attention.self.key."}, {".attention.v_lin.", ".attention.self.value."},
        {".attention.out_lin.", ".attention.output.dense."}, {".sa_layer_norm.", ".attention.output.LayerNorm."},
        {".ffn.lin1.", ".intermediate.dense."}, {".ffn.lin2.", ".output.dense."}, {".output_layer_norm.", ".output.LayerNorm."}};
    static const char *const kAlbert[][2] = {
        {"encoder.embedding_hidden_mapping_in.", "embeddings.projection."},
        {"encoder.albert_layer_groups.0.albert_layers.0.", "encoder.layer.0."},
        {".attention.query.", ".attention.self.query."}, {".attention.key.", ".attention.self.key."},
        {".attention.value.", ".attention.self.value."}, {".attention.dense.", ".attention.output.dense."},
        {".attention.LayerNorm.", ".attention.output.LayerNorm."}, {".ffn_output.", ".output.dense."},
        {".ffn.", ".intermediate.dense."}, {".full_layer_layer_norm.", ".output.LayerNorm."}};
    for (std::string &n : names) {
        for (const char *prefix : {"roberta.", "distilbert.", "albert.", "bert."})
            if (n.rfind(prefix, 0) == 0) n = n.substr(std::strlen(prefix));
        if (n == "linear.weight" || n == "linear.bias") n = "dense." + n;
        if (family == Family::DistilBert)
            for (const auto &r : kDistil) replace_all(n, r[0], r[1]);
        else if (family == Family::Albert)
            for (const auto &r : kAlbert) replace_all(n, r[0], r[1]);
    }
}

void Weights::append(Weights &&other)
{
    for (size_t i = 0; i < other.names.size(); ++i) {
        names.push_back(std::move(other.names[i]));
        data.push_back(std::move(other.data[i]));
    }
}

// ------------------------------------------------------------------------------------------------
// safetensors: u64 header length | JSON header {name: {dtype, shape, data_offsets}} | raw little-endian data
// ------------------------------------------------------------------------------------------------
namespace {
float half_to_float(uint16_t h)
{
    const uint32_t sign = (uint32_t)(h & 0x8000) << 16;
    uint32_t exp = (h >> 10) & 0x1F, man = h & 0x3FF, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            exp = 127 - 15 + 1;
            while (!(man & 0x400)) { man <<= 1; --exp; }
            bits = sign | (exp << 23) | ((man & 0x3FF) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}
}  // namespace

Weights Weights::from_safetensors(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw EmbeddingError(EmbeddingErrorKind::SetupError, path + ": cannot open");
    uint64_t hlen = 0;
    f.read(reinterpret_cast<char *>(&hlen), 8);
    if (!f || hlen == 0 || hlen > (1ull << 28)) throw EmbeddingError(EmbeddingErrorKind::SetupError, path + ": bad safetensors header");
    std::string header(hlen, '\0');
    f.read(&header[0], (std::streamsize)hlen);
    if (!f) throw EmbeddingError(EmbeddingErrorKind::SetupError, path + ": truncated header");
    json::Value root;
    try {
        root = json::parse(header);
    } catch (const std::runtime_error &e) {
        throw EmbeddingError(EmbeddingErrorKind::SetupError, path + ": " + e.what());
    }
    Weights w;
    for (const auto &kv : root.obj) {
        if (kv.first == "__metadata__") continue;
        const json::Value *dt = kv.second.get("dtype"), *off = kv.second.get("data_offsets");
        if (!dt || !off || off->arr.size() != 2) throw EmbeddingError(EmbeddingErrorKind::SetupError, kv.first + ": malformed entry");
        const uint64_t a = (uint64_t)off->arr[0].num, b = (uint64_t)off->arr[1].num;
        std::vector<char> raw(b - a);
        f.seekg((std::streamoff)(8 + hlen + a));
        f.read(raw.data(), (std::streamsize)raw.size());
        if (!f) throw EmbeddingError(EmbeddingErrorKind::SetupError, kv.first + ": truncated data");
        std::vector<float> vals;
        if (dt->str == "F32") {
            vals.resize(raw.size() / 4);
            std::memcpy(vals.data(), raw.data(), vals.size() * 4);
        } else if (dt->str == "F16" || dt->str == "BF16") {
            vals.resize(raw.size() / 2);
            const uint16_t *h = reinterpret_cast<const uint16_t *>(raw.data());
            for (size_t i = 0; i < vals.size(); ++i) {
                if (dt->str == "F16") vals[i] = half_to_float(h[i]);
                else {
                    const uint32_t bits = (uint32_t)h[i] << 16;
                    std::memcpy(&vals[i], &bits, 4);
                }
            }
        } else {
            continue;   // integer buffers (position_ids) are not weights
        }
        // sentence-transformers checkpoints carry the names with or without a "bert." prefix
        std::string name = kv.first;
        if (name.rfind("bert.", 0) == 0) name = name.substr(5);
        w.names.push_back(name);
        w.data.push_back(std::move(vals));
    }
    return w;
}

// ------------------------------------------------------------------------------------------------
// B200Encoder
// ------------------------------------------------------------------------------------------------
B200Encoder::B200Encoder(const Architecture &arch, const Weights &weights, Precision precision, int device, uint32_t max_tokens)
    : arch_(arch)
{
    std::vector<mx_tensor> ts(weights.names.size());
    for (size_t i = 0; i < ts.size(); ++i) {
        ts[i].name = weights.names[i].c_str();
        ts[i].data = weights.data[i].data();
        ts[i].numel = weights.data[i].size();
    }
    mx_model_cfg cfg{};
    cfg.layers = arch.layers;
    cfg.hidden = arch.hidden;
    cfg.heads = arch.heads;
    cfg.ffn = arch.ffn;
    cfg.vocab = arch.vocab;
    cfg.max_pos = arch.max_pos;
    cfg.type_vocab = arch.type_vocab;
    cfg.ln_eps = arch.ln_eps;
    cfg.normalize = arch.normalize ? 1 : 0;
    // (T5 measures 1 - 3.5e-4 in bf16 at two layers, f16 meets the gate; an f16 overflow surfaces as EncodingFailure below)
    if (precision == Precision::AUTO) precision = (arch.layers <= 6 && arch.family != Family::T5) ? Precision::BF16 : Precision::F16;
    f16_ = precision == Precision::F16;
    cfg.precision = (uint32_t)precision;
    cfg.max_tokens = max_tokens;
    mx_model_ext ext{};
    ext.pos_offset = arch.pos_offset;
    ext.no_token_type = arch.family == Family::DistilBert ? 1 : 0;
    ext.dense_out = arch.dense_out;
    ext.dense_act = arch.dense_tanh ? MX_ACT_TANH : MX_ACT_IDENTITY;
    ext.dense_bias = arch.dense_bias ? 1 : 0;
    ext.ffn_act = arch.ffn_gelu_new ? MX_FFN_GELU_TANH : MX_FFN_GELU_ERF;
    ext.embed_dim = arch.embed_dim;
    ext.share_layers = arch.share_layers ? 1 : 0;
    ext.family = arch.family == Family::T5 ? MX_FAMILY_T5 : MX_FAMILY_BERT;
    ext.d_kv = arch.d_kv;
    ext.rel_buckets = arch.rel_buckets;
    ext.rel_max_distance = arch.rel_max_distance;
    int32_t rc = mx_embedder_create_ex(&cfg, &ext, ts.data(), (uint32_t)ts.size(), device, &handle_);
    if (rc != MX_OK) {
        const char *m = mx_last_error(nullptr);
        throw EmbeddingError(EmbeddingErrorKind::SetupError, m && *m ? m : ("status " + std::to_string(rc)));
    }
}

B200Encoder::~B200Encoder()
{
    if (handle_) mx_embedder_destroy(handle_);
}

std::vector<float> B200Encoder::encode_ids(const TokenBatch &b)
{
    std::vector<float> out((size_t)b.B * arch_.out_dim());
    if (b.B == 0) return out;
    int32_t rc = mx_embedder_encode(handle_, b.ids.data(), b.lens.data(), b.B, b.S, out.data());
    if (rc != MX_OK) {
        const char *m = mx_last_error(handle_);
        throw EmbeddingError(EmbeddingErrorKind::EncodingFailure, m && *m ? m : ("status " + std::to_string(rc)));
    }
    if (f16_)
        for (float v : out)
            if (!std::isfinite(v))
                throw EmbeddingError(EmbeddingErrorKind::EncodingFailure,
                                     "non-finite embedding: the model overflowed f16 activations (|x| > 65504); construct the "
                                     "encoder with Precision::BF16");
    return out;
}

// ------------------------------------------------------------------------------------------------
// SentenceEmbedder
// ------------------------------------------------------------------------------------------------
std::shared_ptr<SentenceEmbedder> SentenceEmbedder::spawn(const ModelConfig &model_config, std::shared_ptr<Encoder> encoder,
                                                          std::shared_ptr<Tokenizer> tokenizer)
{
    std::shared_ptr<SentenceEmbedder> e(new SentenceEmbedder());
    e->handle_ = std::thread([p = e.get(), model_config, encoder, tokenizer] { p->runner(model_config, encoder, tokenizer); });
    return e;
}

SentenceEmbedder::~SentenceEmbedder()
{
    {
        std::lock_guard<std::mutex> g(mu_);
        closed_ = true;   // dropping the sender ends `while let Ok(..) = receiver.recv()` (embedding.rs:102)
    }
    not_empty_.notify_all();
    not_full_.notify_all();
    if (handle_.joinable()) handle_.join();
}

// embedding.rs:94-135 embeds ONE document per message: `model.encode(&segments)` sees ~50 segments for a 42 KB text and 1 for a
// query, far from the 256-segment batches the kernels are at their best with.  This runner drains the
// channel: whatever is queued when a request arrives (plus what arrives within kBatchWaitUs, up to kBatchSegments
// segments) is segmented and tokenised on the host and goes through ONE forward pass; every caller gets exactly the rows
// of its own segments, in order.  Rows of a batch are independent (padding is masked, the packed layout drops it), so a
// request's vectors do not depend on what it was batched with.
void SentenceEmbedder::runner(ModelConfig model_config, std::shared_ptr<Encoder> encoder, std::shared_ptr<Tokenizer> tokenizer)
{
    struct Job {
        Message msg;
        std::vector<std::string> segments;
    };
    while (true) {
        std::vector<Job> jobs;
        size_t n_seg = 0;
        bool first = true;
        std::chrono::steady_clock::time_point deadline;
        while (n_seg < kBatchSegments) {
            Message msg;
            {
                std::unique_lock<std::mutex> g(mu_);
                if (first) {
                    not_empty_.wait(g, [this] { return closed_ || !channel_.empty(); });
                    if (channel_.empty()) return;
                    deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(kBatchWaitUs);
                    first = false;
                } else if (!not_empty_.wait_until(g, deadline, [this] { return closed_ || !channel_.empty(); }) || channel_.empty()) {
                    break;
                }
                msg = std::move(channel_.front());
                channel_.pop_front();
            }
            not_full_.notify_one();
            try {
                std::vector<std::string> segments =
                    msg.segment ? segment_text(model_config, msg.text, *tokenizer) : std::vector<std::string>{msg.text};
                n_seg += segments.size();
                jobs.push_back(Job{std::move(msg), std::move(segments)});
            } catch (...) {
                // the reference's runner returns Err and dies; here the caller gets the error and the actor lives on
                msg.sender.set_exception(std::current_exception());
            }
        }
        if (jobs.empty()) continue;
        try {
            std::vector<std::string> flat;
            flat.reserve(n_seg);
            for (const Job &j : jobs) flat.insert(flat.end(), j.segments.begin(), j.segments.end());
            const TokenBatch batch = tokenize_batch(*tokenizer, flat, encoder->max_seq_length());
            const std::vector<float> embeddings = encoder->encode_ids(batch);   // <- model.encode(&segments), embedding.rs:109
            ++batches_;
            const uint32_t H = encoder->hidden();
            if (embeddings.size() != flat.size() * H)   // embedding.rs:110-115
                throw EmbeddingError(EmbeddingErrorKind::EncodingFailure, "# of embeddings doesn't match # of segments");
            size_t at = 0;
            for (Job &j : jobs) {
                std::vector<EmbeddingResult> results;
                results.reserve(j.segments.size());
                for (size_t i = 0; i < j.segments.size(); ++i, ++at)
                    results.push_back({j.segments[i], std::vector<float>(embeddings.begin() + at * H, embeddings.begin() + (at + 1) * H)});
                j.msg.sender.set_value(std::move(results));
            }
        } catch (...) {
            for (Job &j : jobs) {
                try {
                    j.msg.sender.set_exception(std::current_exception());
                } catch (const std::future_error &) {   // already answered
                }
            }
        }
    }
}

std::future<std::vector<EmbeddingResult>> SentenceEmbedder::encode_async(std::string text, bool segment)
{
    Message m{std::move(text), segment, {}};
    auto fut = m.sender.get_future();
    {
        std::unique_lock<std::mutex> g(mu_);
        not_full_.wait(g, [this] { return closed_ || channel_.size() < kChannel; });   // sync_channel(100) blocks the sender
        if (closed_) throw EmbeddingError(EmbeddingErrorKind::EncodingFailure, "embedder has shut down");
        channel_.push_back(std::move(m));
    }
    not_empty_.notify_one();
    return fut;
}

std::vector<EmbeddingResult> SentenceEmbedder::encode(const std::string &text) { return encode_async(text, true).get(); }

std::optional<EmbeddingResult> SentenceEmbedder::encode_single(const std::string &text)
{
    std::vector<EmbeddingResult> v = encode_async(text, false).get();
    if (v.empty()) return std::nullopt;
    return std::move(v.back());   // value.pop(), embedding.rs:150
}

}  // namespace memex
