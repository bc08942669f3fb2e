// memex_host.hpp -- C++ host side of the B200 hot path, above the C ABI of include/memex_b200.h.
//
// The reference's host code is Rust; this image has no Rust toolchain, so the host side that a memex maintainer
// would write in Rust (INTEGRATION.md shows that source) is provided here in C++ with the SAME names, argument
// meaning and error behaviour as the reference surfaces it mirrors:
//
//   storage   lib/libmemex/src/storage/mod.rs:16-139   VectorData, VectorStoreError, trait VectorStore,
//                                                       VectorStorage, get_vector_storage
//             lib/libmemex/src/storage/local.rs:21-166 HnswStore  ->  B200Store
//   embedding lib/libmemex/src/llm/embedding.rs:10-198 EmbeddingError, EmbeddingResult, EmbeddingsModelType,
//                                                       ModelConfig, SentenceEmbedder, segment_text
//
// plus the rows SURVEY.md section 8(f) lists as "next": a process-wide registry of stores / embedders (N1: the
// reference re-loads both per task and per request), the flat on-disk matrix next to a byte-compatible
// vectors.meta.json (N2), a WordPiece tokenizer + windowing in C++ (N3) and a micro-batcher that turns concurrent
// single-query searches into the batches the scan kernels are built for (N4).
//
// All arithmetic happens in libmemex_b200.so on the GPU; nothing here computes a distance or an embedding.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

struct mx_store;
struct mx_shard_group;
struct mx_embedder;

namespace memex {

// ------------------------------------------------------------------------------------------------
// storage/mod.rs
// ------------------------------------------------------------------------------------------------
struct VectorData {   // mod.rs:16-28
    std::string _id;           // internal id of this vector / segment
    std::string document_id;   // document the segment comes from
    std::string text;          // content represented by the vector
    std::vector<float> vector;
    size_t segment_id = 0;
};

enum class StoreErrorKind {   // mod.rs:30-48
    ConnectionError, DeleteError, FileIOError, InsertionError, SearchError, SerdeError, SaveError, Unsupported
};
const char *to_string(StoreErrorKind k);

class VectorStoreError : public std::runtime_error {
public:
    VectorStoreError(StoreErrorKind kind, const std::string &msg);
    StoreErrorKind kind;
};

using VectorSearchResult = std::pair<std::string, float>;   // (doc id, score), mod.rs:51

class VectorStore {   // trait VectorStore, mod.rs:54-66
public:
    virtual ~VectorStore() = default;
    virtual void delete_(const std::string &id) = 0;
    virtual void delete_all() = 0;
    virtual void bulk_insert(const std::vector<VectorData> &data) = 0;
    virtual void insert(const VectorData &data) = 0;
    virtual std::vector<VectorSearchResult> search(const std::vector<float> &vec, size_t limit) const = 0;
    // not in the trait: nq queries in one scan (what the micro-batcher calls); default = a loop over search()
    virtual std::vector<std::vector<VectorSearchResult>> search_batch(const std::vector<std::vector<float>> &vecs,
                                                                      size_t limit) const;
    // search_batch in two halves, so that two batches may be in flight (the micro-batcher stages batch i + 1 while the
    // device answers batch i): submit returns a ticket, collect takes tickets in the order they were issued.
    // Default = the whole search at submit time, the answer parked until collect (any store works behind the batcher).
    virtual uint64_t search_batch_submit(const std::vector<std::vector<float>> &vecs, size_t limit) const;
    virtual std::vector<std::vector<VectorSearchResult>> search_batch_collect(uint64_t ticket) const;

protected:
    mutable std::unordered_map<uint64_t, std::vector<std::vector<VectorSearchResult>>> parked_;
    mutable uint64_t next_ticket_ = 0;
};

// HnswStore (local.rs:21-166) with the index replaced by the GPU row matrix.  Keeps `_id_map` exactly as the
// reference does (local.rs:24,63-64,80-83): usize row id (1-based, insertion order) -> caller's id string.
class B200Store : public VectorStore {
public:
    struct Options {
        uint32_t dim = 0;       // 0 = sized by the first insert (HnswStore takes any width; mod.rs:126 only hints 384)
        bool fp16 = false;      // rows kept as fp16 in HBM (north_star's 10Mx384 configuration)
        bool dot = false;       // dot-product metric instead of 1 - DistCosine
        int device = 0;
        uint64_t capacity = 0;
        bool save_on_insert = true;   // local.rs:66-67 "Naively save after each insert" -- per CALL here, not per row
    };
    static std::unique_ptr<B200Store> new_(const std::string &storage_path, const Options &opt);   // local.rs:95-108
    static std::unique_ptr<B200Store> new_(const std::string &storage_path) { return new_(storage_path, Options()); }
    static bool has_store(const std::string &store_path);                                           // local.rs:110-113
    static std::unique_ptr<B200Store> load(const std::string &store_path, int device = 0);          // local.rs:115-141
    void save(const std::string &store_path) const;                                                 // local.rs:143-165
    ~B200Store() override;

    void delete_(const std::string &id) override;   // local.rs:29-32: unsupported (the reference panics)
    void delete_all() override;                      // local.rs:34-53
    void bulk_insert(const std::vector<VectorData> &data) override;   // local.rs:55-60, ONE device append + ONE save
    void insert(const VectorData &data) override;                     // local.rs:62-69
    std::vector<VectorSearchResult> search(const std::vector<float> &vec, size_t limit) const override;   // local.rs:71-91
    std::vector<std::vector<VectorSearchResult>> search_batch(const std::vector<std::vector<float>> &vecs,
                                                              size_t limit) const override;
    // mx_store_search_submit / _collect: the device works on one batch while the next is staged
    uint64_t search_batch_submit(const std::vector<std::vector<float>> &vecs, size_t limit) const override;
    std::vector<std::vector<VectorSearchResult>> search_batch_collect(uint64_t ticket) const override;

    uint64_t len() const;   // hnsw.get_nb_point(), local.rs:238
    std::string storage_path;
    std::unordered_map<size_t, std::string> _id_map;
    Options options;

private:
    B200Store() = default;
    void create_handle(uint32_t dim);   // the device store is made when the width is known (Options::dim or first insert)
    std::vector<std::vector<VectorSearchResult>> map_results(const uint64_t *ids, const float *scores, const uint32_t *counts,
                                                             size_t nq, uint32_t k) const;
    std::vector<float> pack_queries(const std::vector<std::vector<float>> &vecs, size_t limit) const;
    mx_store *handle_ = nullptr;
    struct InFlight {
        uint64_t device_ticket;
        size_t nq;
        uint32_t k;
    };
    mutable std::unordered_map<uint64_t, InFlight> in_flight_;   // host ticket -> the device search behind it
};

// The same store over SEVERAL GPUs of one box, driven by ONE process (memex's server is one process: mod.rs:68-93).
// Row i of the insertion order lives on shard i % G as local row i / G, and the shard's device reports the GLOBAL
// 1-based id (mx_store_cfg.id_offset = g, id_stride = G), so `_id_map` is keyed exactly as HnswStore's (local.rs:63).
// search = mx_shard_group_search_local: every GPU scans its shard, one peer-memory exchange of the per-shard top-k
// over NVLink, merge -- no torch, no collective library (include/memex_b200.h, mx_shard_group).
// Files: <storage_path>/shard-<g>/vectors.b200.bin next to the one vectors.meta.json.
class ShardedB200Store : public VectorStore {
public:
    struct Options {
        std::vector<int> devices;   // one shard per entry; the same device may appear twice (tests on a 1-GPU box)
        uint32_t dim = 0;           // 0 = sized by the first insert
        bool fp16 = true;
        bool dot = false;
        uint64_t capacity_per_shard = 0;
        uint32_t max_batch = 64;    // queries per search_batch call the exchange buffers are sized for
        bool save_on_insert = false;
    };
    static std::unique_ptr<ShardedB200Store> new_(const std::string &storage_path, const Options &opt);
    static bool has_store(const std::string &store_path) { return B200Store::has_store(store_path); }
    static std::unique_ptr<ShardedB200Store> load(const std::string &store_path, const Options &opt);
    void save(const std::string &store_path) const;
    ~ShardedB200Store() override;

    void delete_(const std::string &id) override;
    void delete_all() override;
    void bulk_insert(const std::vector<VectorData> &data) override;
    void insert(const VectorData &data) override;
    std::vector<VectorSearchResult> search(const std::vector<float> &vec, size_t limit) const override;
    std::vector<std::vector<VectorSearchResult>> search_batch(const std::vector<std::vector<float>> &vecs,
                                                              size_t limit) const override;
    uint64_t len() const;
    size_t shards() const { return options.devices.size(); }
    std::string storage_path;
    std::unordered_map<size_t, std::string> _id_map;
    Options options;

private:
    ShardedB200Store() = default;
    void create_handles(uint32_t dim);
    std::vector<mx_store *> stores_;
    std::vector<mx_shard_group *> groups_;
};

// VectorStorage (mod.rs:68-93): every call takes the one lock, as the tokio Mutex does.
class VectorStorage {
public:
    explicit VectorStorage(std::shared_ptr<VectorStore> client);
    void add_vectors(const std::vector<VectorData> &points);
    void delete_collection();
    std::vector<VectorSearchResult> search(const std::vector<float> &query, size_t limit) const;
    std::vector<std::vector<VectorSearchResult>> search_batch(const std::vector<std::vector<float>> &queries, size_t limit) const;
    // the two halves of search_batch, each under the mutex (the device answers between them with the mutex free)
    uint64_t search_batch_submit(const std::vector<std::vector<float>> &queries, size_t limit) const;
    std::vector<std::vector<VectorSearchResult>> search_batch_collect(uint64_t ticket) const;
    std::shared_ptr<VectorStore> client;

private:
    std::shared_ptr<std::mutex> lock_;
};

// get_vector_storage (mod.rs:95-139): URI-scheme factory.  "b200://<dir>" -> <dir>/<collection>/, loaded if a
// vectors.meta.json exists there, else new.  Unlike the reference (which re-opens the index on every task and
// request, worker/src/tasks.rs:17, api handlers.rs:35,63) handles are kept in a process-wide registry keyed by
// (uri, collection) -- SURVEY.md 8(f) N1.  Options after '?' (stripped from the directory): dtype=f16|f32,
// metric=cosine|dot, device=<n>, dim=<n>; unknown keys or values -> Unsupported.  Without dim= the store takes the
// width of its first insert, as HnswStore does (so 768-d / 512-d models work behind the same URI).
VectorStorage get_vector_storage(const std::string &uri, const std::string &collection);
void drop_vector_storage_registry();   // tests

// N4: concurrent single-query searches -> one batched scan.  submit() returns a future; a worker thread collects
// up to max_batch queries (waiting at most max_wait_us after the first) and issues ONE search_batch -- in its split
// form: while the device answers batch i the worker is already collecting, staging and submitting batch i + 1.
class SearchBatcher {
public:
    SearchBatcher(VectorStorage storage, size_t max_batch = 64, uint32_t max_wait_us = 200);
    ~SearchBatcher();
    std::future<std::vector<VectorSearchResult>> submit(std::vector<float> query, size_t limit);
    uint64_t batches_issued() const { return batches_; }

private:
    struct Req {
        std::vector<float> q;
        size_t limit;
        std::promise<std::vector<VectorSearchResult>> done;
    };
    void run();
    VectorStorage storage_;
    size_t max_batch_;
    uint32_t max_wait_us_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Req> queue_;
    bool stop_ = false;
    uint64_t batches_ = 0;
    std::thread worker_;
};

// ------------------------------------------------------------------------------------------------
// llm/embedding.rs
// ------------------------------------------------------------------------------------------------
enum class EmbeddingErrorKind { EncodingFailure, SetupError };   // embedding.rs:10-16
class EmbeddingError : public std::runtime_error {
public:
    EmbeddingError(EmbeddingErrorKind kind, const std::string &msg);
    EmbeddingErrorKind kind;
};

struct EmbeddingResult {   // embedding.rs:18-22
    std::string content;
    std::vector<float> vector;
};

enum class EmbeddingsModelType {   // embedding.rs:24-33
    DistiluseBaseMultilingualCased, BertBaseNliMeanTokens, AllMiniLmL12V2, AllMiniLmL6V2, AllDistilrobertaV1,
    ParaphraseAlbertSmallV2, SentenceT5Base
};

struct ModelConfig {   // embedding.rs:57-73
    EmbeddingsModelType model = EmbeddingsModelType::AllMiniLmL12V2;
    size_t max_length = 256;
    size_t stride = 86;   // overlap roughly a third of the previous text
};

// What segment_text / the embedder need from `tokenizers::Tokenizer` (embedding.rs:163-195): ids without and with the
// model's special tokens, decoding back to text, and truncation windows.
class Tokenizer {
public:
    virtual ~Tokenizer() = default;
    virtual std::vector<int32_t> encode(const std::string &text, bool add_special_tokens) const = 0;
    virtual std::string decode(const std::vector<int32_t> &ids, bool skip_special_tokens) const = 0;
    // Tokenizer::with_truncation(max_length, stride) + encode(text, false): the first window and the overflowing ones
    std::vector<std::vector<int32_t>> encode_windows(const std::string &text, size_t max_length, size_t stride) const;

    int32_t pad_id = 0, cls_id = 101, sep_id = 102;   // [PAD] [CLS] [SEP]; RoBERTa: <pad> 1, <s> 0, </s> 2
};

// BERT WordPiece tokenizer (what `tokenizers` builds from the models' tokenizer.json: BertNormalizer(clean_text,
// handle_chinese_chars, strip_accents, lowercase) -> BertPreTokenizer -> WordPiece("##", "[UNK]", 100) with the
// WordPiece decoder, cleanup = true).  Accent stripping / lower-casing cover the BMP (tables generated from Python's
// unicodedata, gen_unicode_tables.py).
class BertTokenizer : public Tokenizer {
public:
    static std::shared_ptr<BertTokenizer> from_vocab_file(const std::string &vocab_txt, bool lowercase = true);
    static std::shared_ptr<BertTokenizer> from_vocab(const std::vector<std::string> &tokens, bool lowercase = true);

    std::vector<int32_t> encode(const std::string &text, bool add_special_tokens) const override;
    std::string decode(const std::vector<int32_t> &ids, bool skip_special_tokens) const override;

    int32_t unk_id = 100, mask_id = 103;
    size_t vocab_size() const { return id_to_token_.size(); }

private:
    std::vector<std::string> id_to_token_;
    std::unordered_map<std::string, int32_t> token_to_id_;
    bool lowercase_ = true;
    std::vector<std::u32string> pre_tokenize(const std::string &text) const;
    void wordpiece(const std::u32string &word, std::vector<int32_t> &out) const;
};

// Byte-level BPE tokenizer of the RoBERTa stacks (AllDistilrobertaV1 -- the third model segment_text accepts,
// embedding.rs:156-161): what `tokenizers` builds from that model's tokenizer.json -- no normalizer, ByteLevel
// pre-tokenizer (GPT-2 regex, add_prefix_space = false), BPE merges, ByteLevel decoder, RobertaProcessing
// (<s> ... </s>).  Specials: <s> 0, <pad> 1, </s> 2, <unk> 3, <mask> the last id the vocabulary gives it.
class ByteLevelBpeTokenizer : public Tokenizer {
public:
    // vocab.json ({"token": id}) + merges.txt ("a b" per line, first line may be a #version comment)
    static std::shared_ptr<ByteLevelBpeTokenizer> from_files(const std::string &vocab_json, const std::string &merges_txt);
    // tokens in id order (byte-level spelling, as in vocab.json), merges in rank order
    static std::shared_ptr<ByteLevelBpeTokenizer> from_vocab(const std::vector<std::string> &tokens,
                                                             const std::vector<std::pair<std::string, std::string>> &merges);

    std::vector<int32_t> encode(const std::string &text, bool add_special_tokens) const override;
    std::string decode(const std::vector<int32_t> &ids, bool skip_special_tokens) const override;

    int32_t unk_id = 3, mask_id = -1;
    size_t vocab_size() const { return id_to_token_.size(); }

private:
    std::vector<std::string> id_to_token_;
    std::unordered_map<std::string, int32_t> token_to_id_;
    std::unordered_map<std::string, uint32_t> merge_rank_;   // "left right" -> rank
    void bpe(const std::string &piece, std::vector<int32_t> &out) const;
};

// segment_text (embedding.rs:155-198): windows of max_length tokens overlapping by stride, each decoded back to
// TEXT (the first one also gets .replace(" ' ", "'"), :183).  The tokenizer is passed in: the reference re-loads it
// from the hub on every call (:163).
std::vector<std::string> segment_text(const ModelConfig &model_config, const std::string &text, const Tokenizer &tokenizer);

// rust-bert's SentenceEmbeddingsModel::tokenize step: [CLS] .. [SEP], truncate to max_seq_length, pad to the longest.
struct TokenBatch {
    std::vector<int32_t> ids;    // [B, S]
    std::vector<int32_t> lens;   // [B]
    uint32_t B = 0, S = 0;
};
TokenBatch tokenize_batch(const Tokenizer &tokenizer, const std::vector<std::string> &segments, size_t max_seq_length);

enum class Family { Bert, Roberta, DistilBert, Albert, T5 };   // the checkpoint's naming scheme / embedding layout
struct Architecture {
    uint32_t layers, hidden, heads, ffn, vocab = 30522, max_pos = 512, type_vocab = 2;
    float ln_eps = 1e-12f;
    bool normalize = true;
    uint32_t max_seq_length = 256;   // sentence_bert_config.json: what rust-bert truncates to
    // the stacks that differ from BERT only around the layers (mx_model_ext in include/memex_b200.h)
    Family family = Family::Bert;
    uint32_t pos_offset = 0;         // RoBERTa: positions start at padding_idx + 1 = 2
    int32_t pad_id = 0;              // RoBERTa: 1
    uint32_t dense_out = 0;          // sentence-transformers Dense module after pooling (0 = none)
    bool dense_tanh = false, dense_bias = true;
    bool ffn_gelu_new = false;       // ALBERT's tanh-form GELU
    uint32_t embed_dim = 0;          // ALBERT: factorised embedding width (0 = hidden)
    bool share_layers = false;       // ALBERT: one set of layer weights
    // Family::T5 (SentenceT5Base): T5 encoder stack, weights under the HF T5EncoderModel names (include/memex_b200.h)
    uint32_t d_kv = 0, rel_buckets = 32, rel_max_distance = 128;
    uint32_t out_dim() const { return dense_out ? dense_out : hidden; }
};
// every member of the enum: six stacks around BERT's post-LayerNorm block and SentenceT5Base (T5 v1.1 encoder + Dense)
std::optional<Architecture> architecture_of(EmbeddingsModelType model);

// named f32 tensors (HF BertModel names); from_safetensors reads a model.safetensors file (F32 / F16 / BF16)
struct Weights {
    std::vector<std::string> names;
    std::vector<std::vector<float>> data;
    static Weights from_safetensors(const std::string &path);
    // RoBERTa / DistilBERT / ALBERT checkpoint names -> the BERT names the C ABI takes ("roberta." / "distilbert." /
    // "albert." / "bert." prefixes dropped, a sentence-transformers Dense module's "linear.*" -> "dense.linear.*")
    void canonicalize(Family family);
    void append(Weights &&other);   // e.g. the 2_Dense/model.safetensors of a sentence-transformers checkpoint
};

// the forward pass: `model.encode(&segments)` (embedding.rs:109) on ids
class Encoder {
public:
    virtual ~Encoder() = default;
    virtual uint32_t hidden() const = 0;                     // width of one output row
    virtual uint32_t max_seq_length() const = 0;
    virtual std::vector<float> encode_ids(const TokenBatch &batch) = 0;   // [B, hidden()]; unit-norm rows if the model normalises
};
class B200Encoder : public Encoder {
public:
    // BF16 / F16: the tensor-core paths (same tcgen05 kernels and speed); F32: CUDA-core validation path; AUTO = BF16 up
    // to 6 layers, F16 beyond (bf16's 8-bit mantissa compounds to cosine 1 - 1.05e-4 against the fp32 oracle at 12 layers,
    // f16's 11 bits stay at 1 - 2e-6: tests/test_encoder_gpu.py).  An f16 overflow (|x| > 65504, exotic checkpoints)
    // surfaces as EncodingFailure naming Precision::BF16.
    enum class Precision { BF16 = 0, F32 = 1, F16 = 2, AUTO = 3 };
    B200Encoder(const Architecture &arch, const Weights &weights, Precision precision = Precision::AUTO, int device = 0,
                uint32_t max_tokens = 0);
    ~B200Encoder() override;
    uint32_t hidden() const override { return arch_.out_dim(); }
    uint32_t max_seq_length() const override { return arch_.max_seq_length; }
    std::vector<float> encode_ids(const TokenBatch &batch) override;

private:
    Architecture arch_;
    mx_embedder *handle_ = nullptr;
    bool f16_ = false;
};

// SentenceEmbedder (embedding.rs:77-152): `spawn` starts the runner thread that owns the model; requests arrive
// over a bounded channel of 100 (mpsc::sync_channel(100), :87) and are answered through a one-shot promise.
class SentenceEmbedder {
public:
    // the already-loaded encoder / tokenizer are handed in (process-wide, N1); the reference loads both inside the runner
    static std::shared_ptr<SentenceEmbedder> spawn(const ModelConfig &model_config, std::shared_ptr<Encoder> encoder,
                                                   std::shared_ptr<Tokenizer> tokenizer);
    ~SentenceEmbedder();
    std::vector<EmbeddingResult> encode(const std::string &text);                    // segment + embed each window
    std::optional<EmbeddingResult> encode_single(const std::string &text);           // one shot, truncated by the model
    std::future<std::vector<EmbeddingResult>> encode_async(std::string text, bool segment);
    uint64_t batches_run() const { return batches_; }   // forward passes issued (requests queued together share one)

private:
    static constexpr size_t kBatchSegments = 256;   // segments per forward pass (BASELINE.json config 3's batch)
    static constexpr uint32_t kBatchWaitUs = 500;   // how long the first request of a batch waits for company
    std::atomic<uint64_t> batches_{0};
    struct Message {
        std::string text;
        bool segment;
        std::promise<std::vector<EmbeddingResult>> sender;
    };
    SentenceEmbedder() = default;
    void runner(ModelConfig model_config, std::shared_ptr<Encoder> encoder, std::shared_ptr<Tokenizer> tokenizer);
    static constexpr size_t kChannel = 100;
    std::mutex mu_;
    std::condition_variable not_empty_, not_full_;
    std::deque<Message> channel_;
    bool closed_ = false;
    std::thread handle_;
};

}  // namespace memex
