// test_host.cpp -- tests of the C++ host layer; driven by tests/test_host_cpp.py.
//   test_host tokenizer <golden.json>     WordPiece / decode / windows / segment_text against the `tokenizers` package
//   test_host cpu                         host logic that needs no GPU (actor, batcher, factory errors, JSON)
//   test_host gpu <tmpdir>                the reference's own store tests (local.rs:175-242) + registry + batcher on the GPU
//   test_host encode <model.safetensors> <ids.bin> <out.bin> L H heads F vocab max_pos precision
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <sys/stat.h>

#include "../../include/memex_b200.h"
#include "json.hpp"
#include "memex_host.hpp"

using namespace memex;

static int g_checks = 0;
#define CHECK(cond)                                                                      \
    do {                                                                                 \
        ++g_checks;                                                                      \
        if (!(cond)) {                                                                   \
            std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            std::exit(1);                                                                \
        }                                                                                \
    } while (0)

static void make_dir(const std::string &p) { ::mkdir(p.c_str(), 0777); }

static std::string slurp(const std::string &p)
{
    std::ifstream f(p, std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

static std::vector<int32_t> ints(const json::Value &a)
{
    std::vector<int32_t> v;
    for (const auto &x : a.arr) v.push_back((int32_t)x.num);
    return v;
}

static int run_tokenizer(const std::string &golden)
{
    json::Value g = json::parse(slurp(golden));
    std::vector<std::string> vocab;
    for (const auto &t : g.get("vocab")->arr) vocab.push_back(t.str);
    const bool lowercase = !g.get("lowercase") || g.get("lowercase")->b;   // the cased fixture says false
    auto tok = BertTokenizer::from_vocab(vocab, lowercase);
    int n = 0;
    for (const auto &c : g.get("cases")->arr) {
        const std::string text = c.get("text")->str;
        const auto ids = tok->encode(text, false);
        if (ids != ints(*c.get("ids"))) {
            std::fprintf(stderr, "ids differ for case %d: %s\n got:", n, text.c_str());
            for (auto i : ids) std::fprintf(stderr, " %d", i);
            std::fprintf(stderr, "\nwant:");
            for (auto i : ints(*c.get("ids"))) std::fprintf(stderr, " %d", i);
            std::fprintf(stderr, "\n");
            return 1;
        }
        CHECK(tok->encode(text, true) == ints(*c.get("ids_special")));
        const std::string dec = tok->decode(ids, true);
        if (dec != c.get("decoded")->str) {
            std::fprintf(stderr, "decode differs for case %d:\n got  %s\n want %s\n", n, dec.c_str(), c.get("decoded")->str.c_str());
            return 1;
        }
        const size_t max_length = (size_t)c.get("max_length")->num, stride = (size_t)c.get("stride")->num;
        const auto windows = tok->encode_windows(text, max_length, stride);
        const auto &gw = c.get("windows")->arr;
        CHECK(windows.size() == gw.size());
        for (size_t i = 0; i < gw.size(); ++i) CHECK(windows[i] == ints(gw[i]));
        ModelConfig cfg;
        cfg.model = EmbeddingsModelType::AllMiniLmL6V2;
        cfg.max_length = max_length;
        cfg.stride = stride;
        const auto segs = segment_text(cfg, text, *tok);
        const auto &gs = c.get("segments")->arr;
        CHECK(segs.size() == gs.size());
        for (size_t i = 0; i < gs.size(); ++i) {
            if (segs[i] != gs[i].str) {
                std::fprintf(stderr, "segment %zu differs for case %d:\n got  %s\n want %s\n", i, n, segs[i].c_str(), gs[i].str.c_str());
                return 1;
            }
        }
        ++n;
    }
    // models the reference cannot segment (embedding.rs:156-161)
    ModelConfig bad;
    bad.model = EmbeddingsModelType::SentenceT5Base;
    try {
        segment_text(bad, "x", *tok);
        CHECK(false);
    } catch (const EmbeddingError &e) {
        CHECK(e.kind == EmbeddingErrorKind::SetupError);
    }
    std::printf("tokenizer ok: %d cases, %d checks\n", n, g_checks);
    return 0;
}

// byte-level BPE (all-distilroberta-v1's pipeline) against tests/golden/bpe_golden.json
static int run_bpe(const std::string &golden)
{
    json::Value g = json::parse(slurp(golden));
    std::vector<std::string> vocab;
    for (const auto &t : g.get("vocab")->arr) vocab.push_back(t.str);
    std::vector<std::pair<std::string, std::string>> merges;
    for (const auto &m : g.get("merges")->arr) merges.emplace_back(m.arr[0].str, m.arr[1].str);
    auto tok = ByteLevelBpeTokenizer::from_vocab(vocab, merges);
    CHECK(tok->cls_id == 0 && tok->pad_id == 1 && tok->sep_id == 2 && tok->unk_id == 3 && tok->mask_id == 4);
    int n = 0;
    for (const auto &c : g.get("cases")->arr) {
        const std::string text = c.get("text")->str;
        const auto ids = tok->encode(text, false);
        if (ids != ints(*c.get("ids"))) {
            std::fprintf(stderr, "ids differ for case %d: %s\n got:", n, text.c_str());
            for (auto i : ids) std::fprintf(stderr, " %d", i);
            std::fprintf(stderr, "\nwant:");
            for (auto i : ints(*c.get("ids"))) std::fprintf(stderr, " %d", i);
            std::fprintf(stderr, "\n");
            return 1;
        }
        CHECK(tok->encode(text, true) == ints(*c.get("ids_special")));
        const std::string dec = tok->decode(ids, true);
        if (dec != c.get("decoded")->str) {
            std::fprintf(stderr, "decode differs for case %d:\n got  %s\n want %s\n", n, dec.c_str(), c.get("decoded")->str.c_str());
            return 1;
        }
        const size_t max_length = (size_t)c.get("max_length")->num, stride = (size_t)c.get("stride")->num;
        const auto windows = tok->encode_windows(text, max_length, stride);
        const auto &gw = c.get("windows")->arr;
        CHECK(windows.size() == gw.size());
        for (size_t i = 0; i < gw.size(); ++i) CHECK(windows[i] == ints(gw[i]));
        ModelConfig cfg;
        cfg.model = EmbeddingsModelType::AllDistilrobertaV1;
        cfg.max_length = max_length;
        cfg.stride = stride;
        const auto segs = segment_text(cfg, text, *tok);
        const auto &gs = c.get("segments")->arr;
        CHECK(segs.size() == gs.size());
        for (size_t i = 0; i < gs.size(); ++i) {
            if (segs[i] != gs[i].str) {
                std::fprintf(stderr, "segment %zu differs for case %d:\n got  %s\n want %s\n", i, n, segs[i].c_str(), gs[i].str.c_str());
                return 1;
            }
        }
        // the embedder's batch: <s> .. </s>, padded with <pad>
        const TokenBatch tb = tokenize_batch(*tok, {text, "a"}, 12);
        CHECK(tb.B == 2 && tb.ids[0] == tok->cls_id && tb.ids[tb.lens[0] - 1] == tok->sep_id && tb.lens[0] <= 12);
        for (uint32_t i = (uint32_t)tb.lens[1]; i < tb.S; ++i) CHECK(tb.ids[tb.S + i] == tok->pad_id);
        ++n;
    }
    std::printf("bpe ok: %d cases, %d checks\n", n, g_checks);
    return 0;
}

// ---- fakes for the CPU tests --------------------------------------------------------------------
struct FakeEncoder : Encoder {
    std::atomic<int> calls{0};
    uint32_t hidden() const override { return 4; }
    uint32_t max_seq_length() const override { return 16; }
    std::vector<float> encode_ids(const TokenBatch &b) override
    {
        ++calls;
        std::vector<float> out((size_t)b.B * 4);
        for (uint32_t i = 0; i < b.B; ++i) {
            out[i * 4 + 0] = (float)b.lens[i];
            out[i * 4 + 1] = (float)b.ids[(size_t)i * b.S];          // [CLS]
            out[i * 4 + 2] = (float)b.ids[(size_t)i * b.S + 1];      // first token
            out[i * 4 + 3] = (float)b.S;
        }
        return out;
    }
};

struct FakeStore : VectorStore {
    mutable std::atomic<int> batch_calls{0};
    mutable std::atomic<int> queries{0};
    void delete_(const std::string &) override { throw VectorStoreError(StoreErrorKind::Unsupported, "x"); }
    void delete_all() override {}
    void bulk_insert(const std::vector<VectorData> &) override {}
    void insert(const VectorData &) override {}
    std::vector<VectorSearchResult> search(const std::vector<float> &v, size_t limit) const override
    {
        return search_batch({v}, limit)[0];
    }
    std::vector<std::vector<VectorSearchResult>> search_batch(const std::vector<std::vector<float>> &vs, size_t limit) const override
    {
        ++batch_calls;
        queries += (int)vs.size();
        std::this_thread::sleep_for(std::chrono::milliseconds(2));   // a "scan"
        std::vector<std::vector<VectorSearchResult>> out;
        for (const auto &v : vs) out.push_back({{"doc-" + std::to_string((int)v[0]), (float)limit}});
        return out;
    }
};

static std::shared_ptr<BertTokenizer> tiny_tokenizer()
{
    std::vector<std::string> vocab = {"[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"};
    for (const char *w : {"a", "b", "c", "d", "e", "f", "g", "h", "the", "fox", "##s", ".", ",", "'"}) vocab.push_back(w);
    return BertTokenizer::from_vocab(vocab);
}

static int run_cpu()
{
    // ---- JSON: serde_json-style id map round trip
    {
        std::string s;
        json::escape_into("a\"b\\c\n\x01", s);
        CHECK(s == "\"a\\\"b\\\\c\\n\\u0001\"");
        json::Value v = json::parse("{\"1\":\"uuid-1\",\"22\":\"x\\ty\\u00e9\",\"n\":[1,2.5,true,null,{\"k\":\"v\"}]}");
        CHECK(v.get("1")->str == "uuid-1");
        CHECK(v.get("22")->str == "x\ty\xc3\xa9");
        CHECK(v.get("n")->arr.size() == 5 && v.get("n")->arr[1].num == 2.5 && v.get("n")->arr[4].get("k")->str == "v");
        bool threw = false;
        try { json::parse("{\"a\":}"); } catch (const std::runtime_error &) { threw = true; }
        CHECK(threw);
    }
    // ---- factory errors (mod.rs:99-102,136)
    for (const char *uri : {"not a uri", "hnsw:///tmp/x", "qdrant://host", "://x"}) {
        try {
            get_vector_storage(uri, "c");
            CHECK(false);
        } catch (const VectorStoreError &e) {
            CHECK(e.kind == StoreErrorKind::Unsupported);
        }
    }
    // ---- SentenceEmbedder actor over a fake encoder
    {
        auto enc = std::make_shared<FakeEncoder>();
        auto tok = tiny_tokenizer();
        ModelConfig cfg;
        cfg.model = EmbeddingsModelType::AllMiniLmL6V2;
        cfg.max_length = 4;
        cfg.stride = 1;
        auto emb = SentenceEmbedder::spawn(cfg, enc, tok);
        // 9 tokens, windows of 4 stepping by 3: [0:4] [3:7] [6:9]
        auto res = emb->encode("a b c d e f g h the");
        CHECK(res.size() == 3);
        CHECK(res[0].content == "a b c d" && res[1].content == "d e f g" && res[2].content == "g h the");
        CHECK(res[0].vector.size() == 4 && res[0].vector[0] == 6.f /* 4 tokens + [CLS] + [SEP] */ && res[0].vector[1] == 2.f);
        CHECK(res[2].vector[0] == 5.f);
        CHECK(enc->calls == 1);   // one forward pass per document
        auto one = emb->encode_single("the fox");
        CHECK(one.has_value() && one->content == "the fox" && one->vector[0] == 4.f);
        // single shot longer than the model's window is truncated by the model (embedding.rs:144-151): 16 = max_seq_length
        auto longer = emb->encode_single("a a a a a a a a a a a a a a a a a a a a a a a a");
        CHECK(longer->vector[0] == 16.f);
        // many concurrent callers go through the bounded channel
        std::vector<std::future<std::vector<EmbeddingResult>>> futs;
        const int calls_before = enc->calls;
        for (int i = 0; i < 300; ++i) futs.push_back(emb->encode_async(i % 2 ? "a b" : "the fox a", false));
        for (int i = 0; i < 300; ++i) {
            auto r = futs[i].get();
            // every caller gets ITS row out of the shared forward pass: token count of its own text (+ [CLS] + [SEP])
            CHECK(r.size() == 1 && r[0].content == (i % 2 ? "a b" : "the fox a") && r[0].vector[0] == (i % 2 ? 4.f : 5.f));
        }
        // N4: requests queued together share forward passes (300 requests -> far fewer passes)
        CHECK(enc->calls - calls_before < 300 && emb->batches_run() >= 1);
        // an unsupported model surfaces as SetupError to the caller and the actor survives
        ModelConfig bad;
        bad.model = EmbeddingsModelType::SentenceT5Base;
        auto emb2 = SentenceEmbedder::spawn(bad, enc, tok);
        bool threw = false;
        try { emb2->encode("a"); } catch (const EmbeddingError &e) { threw = e.kind == EmbeddingErrorKind::SetupError; }
        CHECK(threw);
        CHECK(emb2->encode_single("a").has_value());
    }
    // ---- SearchBatcher: concurrent single queries become few batched scans
    {
        auto fake = std::make_shared<FakeStore>();
        VectorStorage vs(fake);
        SearchBatcher batcher(vs, 16, 20000);
        std::vector<std::future<std::vector<VectorSearchResult>>> futs;
        for (int i = 0; i < 64; ++i) futs.push_back(batcher.submit({(float)i, 0.f}, 10));
        for (int i = 0; i < 64; ++i) {
            auto r = futs[i].get();
            CHECK(r.size() == 1 && r[0].first == "doc-" + std::to_string(i) && r[0].second == 10.f);
        }
        CHECK(fake->queries == 64);
        CHECK(fake->batch_calls <= 8 && batcher.batches_issued() == (uint64_t)fake->batch_calls);   // 64 / 16 = 4 when all queue up in time
        // different limits are never mixed into one scan
        auto f1 = batcher.submit({1.f, 0.f}, 3);
        auto f2 = batcher.submit({2.f, 0.f}, 7);
        CHECK(f1.get()[0].second == 3.f && f2.get()[0].second == 7.f);
    }
    // ---- architecture table
    CHECK(architecture_of(EmbeddingsModelType::AllMiniLmL6V2)->layers == 6);
    CHECK(architecture_of(EmbeddingsModelType::AllMiniLmL12V2)->max_seq_length == 128);
    CHECK(architecture_of(EmbeddingsModelType::SentenceT5Base)->family == Family::T5);
    CHECK(architecture_of(EmbeddingsModelType::SentenceT5Base)->d_kv * 12 == 768 && !architecture_of(EmbeddingsModelType::SentenceT5Base)->dense_bias);
    CHECK(architecture_of(EmbeddingsModelType::AllDistilrobertaV1)->pos_offset == 2);
    CHECK(architecture_of(EmbeddingsModelType::DistiluseBaseMultilingualCased)->out_dim() == 512);
    CHECK(architecture_of(EmbeddingsModelType::ParaphraseAlbertSmallV2)->embed_dim == 128);
    {   // checkpoint names of the other stacks -> the BERT names the C ABI takes
        Weights w;
        w.names = {"distilbert.transformer.layer.3.attention.q_lin.weight", "transformer.layer.0.ffn.lin2.bias",
                   "distilbert.transformer.layer.1.sa_layer_norm.weight", "linear.weight"};
        w.data.resize(w.names.size());
        w.canonicalize(Family::DistilBert);
        CHECK(w.names[0] == "encoder.layer.3.attention.self.query.weight");
        CHECK(w.names[1] == "encoder.layer.0.output.dense.bias");
        CHECK(w.names[2] == "encoder.layer.1.attention.output.LayerNorm.weight");
        CHECK(w.names[3] == "dense.linear.weight");
        Weights a;
        a.names = {"albert.encoder.embedding_hidden_mapping_in.weight",
                   "encoder.albert_layer_groups.0.albert_layers.0.attention.query.bias",
                   "encoder.albert_layer_groups.0.albert_layers.0.ffn.weight",
                   "encoder.albert_layer_groups.0.albert_layers.0.ffn_output.weight",
                   "encoder.albert_layer_groups.0.albert_layers.0.full_layer_layer_norm.bias",
                   "encoder.albert_layer_groups.0.albert_layers.0.attention.LayerNorm.weight"};
        a.data.resize(a.names.size());
        a.canonicalize(Family::Albert);
        CHECK(a.names[0] == "embeddings.projection.weight");
        CHECK(a.names[1] == "encoder.layer.0.attention.self.query.bias");
        CHECK(a.names[2] == "encoder.layer.0.intermediate.dense.weight");
        CHECK(a.names[3] == "encoder.layer.0.output.dense.weight");
        CHECK(a.names[4] == "encoder.layer.0.output.LayerNorm.bias");
        CHECK(a.names[5] == "encoder.layer.0.attention.output.LayerNorm.weight");
        Weights r;
        r.names = {"roberta.embeddings.word_embeddings.weight"};
        r.data.resize(1);
        r.canonicalize(Family::Roberta);
        CHECK(r.names[0] == "embeddings.word_embeddings.weight");
    }
    // ---- no CPU path: creating a store without a CUDA device is a ConnectionError
    try {
        B200Store::Options o;
        o.dim = 384;   // a given width makes the device store at once (dim = 0 waits for the first insert)
        auto s = B200Store::new_("/tmp/mx_host_cpu_probe", o);
        std::printf("cpu ok (a CUDA device is present): %d checks\n", g_checks);
    } catch (const VectorStoreError &e) {
        CHECK(e.kind == StoreErrorKind::ConnectionError);
        std::printf("cpu ok (no CUDA device: ConnectionError as required): %d checks\n", g_checks);
    }
    return 0;
}

static std::vector<VectorData> test_data()   // local.rs:175-199
{
    return {{"test-one", "test-one", "", {0.0f, 0.1f, 0.2f}, 0},
            {"test-two", "test-two", "", {0.1f, 0.1f, 0.1f}, 0},
            {"test-three", "test-three", "", {0.3f, 0.2f, 0.1f}, 0}};
}

static int run_gpu(const std::string &tmp)
{
    B200Store::Options o3;
    o3.dim = 3;
    // test_hnsw (local.rs:201-214)
    {
        auto store = B200Store::new_(tmp + "/t1", o3);
        store->bulk_insert(test_data());
        auto results = store->search({0.1f, 0.1f, 0.1f}, 3);
        CHECK(results.size() == 3);
        CHECK(results[0].first == "test-two");
        CHECK(results[1].first == "test-three" && results[2].first == "test-one");
        CHECK(std::fabs(results[0].second - 1.0f) < 1e-6f && std::fabs(results[1].second - 0.9258201f) < 1e-6f &&
              std::fabs(results[2].second - 0.7745967f) < 1e-6f);
        CHECK(store->search({0.1f, 0.1f, 0.1f}, 2).size() == 2);   // clippy asks for 2 (examples/clippy/src/main.rs:209)
        try {
            store->delete_("test-one");
            CHECK(false);
        } catch (const VectorStoreError &e) {
            CHECK(e.kind == StoreErrorKind::Unsupported);   // the reference panics here (local.rs:29-32)
        }
        try {
            store->search({0.1f, 0.1f}, 3);
            CHECK(false);
        } catch (const VectorStoreError &e) {
            CHECK(e.kind == StoreErrorKind::SearchError);
        }
        store->delete_all();
    }
    // test_save_load (local.rs:216-227)
    {
        auto store = B200Store::new_(tmp + "/vectortest", o3);
        store->bulk_insert(test_data());
        store->save(tmp + "/vectortest");
        auto loaded = B200Store::load(tmp + "/vectortest");
        CHECK(loaded->_id_map.size() == store->_id_map.size());
        CHECK(loaded->len() == 3 && loaded->options.dim == 3);
        auto r = loaded->search({0.1f, 0.1f, 0.1f}, 3);
        CHECK(r.size() == 3 && r[0].first == "test-two");
        // the id map file is what serde_json writes for HashMap<usize, String>: a flat object, string keys
        json::Value meta = json::parse(slurp(tmp + "/vectortest/vectors.meta.json"));
        CHECK(meta.kind == json::Value::Object && meta.obj.size() == 3 && meta.get("2")->str == "test-two");
        store->delete_all();
    }
    // test_delete_all (local.rs:229-242)
    {
        auto store = B200Store::new_(tmp + "/t3", o3);
        store->bulk_insert(test_data());
        store->save(tmp + "/t3");
        store->delete_all();
        CHECK(store->_id_map.empty());
        CHECK(store->len() == 0);
        CHECK(!B200Store::has_store(tmp + "/t3"));
        try {
            B200Store::load(tmp + "/t3");
            CHECK(false);
        } catch (const VectorStoreError &) {
        }
    }
    // factory + registry (N1): the same handle comes back; a fresh process-state re-loads from disk
    {
        const std::string uri = "b200://" + tmp + "/collections";
        VectorStorage a = get_vector_storage(uri, "docs");
        std::vector<VectorData> pts;
        for (int i = 0; i < 500; ++i) {
            VectorData d;
            d._id = "seg-" + std::to_string(i);
            d.vector.assign(384, 0.f);
            d.vector[i % 384] = 1.f;
            d.vector[(i * 7 + 1) % 384] += 0.25f + 0.001f * i;
            pts.push_back(d);
        }
        a.add_vectors(pts);
        VectorStorage b = get_vector_storage(uri, "docs");
        CHECK(a.client.get() == b.client.get());
        auto r1 = b.search(pts[123].vector, 10);
        CHECK(r1.size() == 10 && r1[0].first == "seg-123" && std::fabs(r1[0].second - 1.0f) < 1e-6f);
        for (size_t i = 1; i < r1.size(); ++i) CHECK(r1[i - 1].second >= r1[i].second);
        drop_vector_storage_registry();
        VectorStorage c = get_vector_storage(uri, "docs");   // vectors.meta.json exists -> load (mod.rs:115-119)
        CHECK(c.client.get() != a.client.get());
        auto r2 = c.search(pts[123].vector, 10);
        CHECK(r2 == r1);
        // N4: the batcher's answers are the direct answers
        SearchBatcher batcher(c, 64, 2000);
        std::vector<std::future<std::vector<VectorSearchResult>>> futs;
        for (int i = 0; i < 200; ++i) futs.push_back(batcher.submit(pts[i].vector, 5));
        for (int i = 0; i < 200; ++i) {
            auto r = futs[i].get();
            CHECK(r.size() == 5 && r[0].first == "seg-" + std::to_string(i));
            if (i % 50 == 0) CHECK(r == c.search(pts[i].vector, 5));
        }
        CHECK(batcher.batches_issued() < 200);
        // the split call (two batches in flight, collected in order) gives the blocking call's answers; a third submit
        // before a collect is refused
        {
            std::vector<std::vector<float>> qa, qb;
            for (int i = 0; i < 64; ++i) qa.push_back(pts[i].vector);
            for (int i = 64; i < 73; ++i) qb.push_back(pts[i].vector);
            const auto want_a = c.search_batch(qa, 7), want_b = c.search_batch(qb, 7);
            const uint64_t ta = c.search_batch_submit(qa, 7), tb = c.search_batch_submit(qb, 7);
            bool refused = false;
            try {
                c.search_batch_submit(qa, 7);
            } catch (const VectorStoreError &e) {
                refused = e.kind == StoreErrorKind::SearchError;
            }
            CHECK(refused);
            CHECK(c.search_batch_collect(ta) == want_a);
            CHECK(c.search_batch_collect(tb) == want_b);
            const uint64_t tc = c.search_batch_submit(qb, 7);
            CHECK(c.search_batch_collect(tc) == want_b);
        }
        // more than MX_MAX_K neighbours: an error in every host, never a silent cut
        try {
            c.search(pts[0].vector, 257);
            CHECK(false);
        } catch (const VectorStoreError &e) {
            CHECK(e.kind == StoreErrorKind::SearchError);
        }
        CHECK(c.search(pts[0].vector, 256).size() == 256);
        c.delete_collection();
        CHECK(c.search(pts[0].vector, 3).empty());
        drop_vector_storage_registry();
    }
    // the factory sizes a store by its first insert (HnswStore takes any width): the 768-d and 512-d models of the
    // enum work behind the same URI; '?' options are parsed and never reach the directory name
    {
        const std::string base = tmp + "/lazy";
        VectorStorage w768 = get_vector_storage("b200+f16://" + base, "wide");
        CHECK(w768.search(std::vector<float>(768, 1.f), 3).empty());   // nothing inserted yet
        CHECK(!B200Store::has_store(base + "/wide"));
        w768.add_vectors({});                                           // an empty batch leaves no files behind
        CHECK(!B200Store::has_store(base + "/wide"));
        std::vector<VectorData> pts;
        for (int i = 0; i < 40; ++i) {
            VectorData d;
            d._id = "w-" + std::to_string(i);
            d.vector.assign(768, 0.01f);
            d.vector[i * 3] = 1.f;
            pts.push_back(d);
        }
        w768.add_vectors(pts);
        auto r = w768.search(pts[7].vector, 4);
        CHECK(r.size() == 4 && r[0].first == "w-7");
        auto *st = dynamic_cast<B200Store *>(w768.client.get());
        CHECK(st && st->options.dim == 768 && st->options.fp16);
        try {
            VectorData bad;
            bad._id = "narrow";
            bad.vector.assign(384, 1.f);
            w768.add_vectors({bad});
            CHECK(false);
        } catch (const VectorStoreError &e) {
            CHECK(e.kind == StoreErrorKind::InsertionError);
        }
        VectorStorage opt = get_vector_storage("b200://" + base + "?dtype=f16&metric=dot&dim=512&device=0", "opts");
        auto *so = dynamic_cast<B200Store *>(opt.client.get());
        CHECK(so && so->options.dim == 512 && so->options.fp16 && so->options.dot);
        CHECK(so->storage_path == base + "/opts");
        for (const char *bad_uri : {"?dtype=f64", "?colour=red", "?dim=abc"}) {
            try {
                get_vector_storage("b200://" + base + bad_uri, "x");
                CHECK(false);
            } catch (const VectorStoreError &e) {
                CHECK(e.kind == StoreErrorKind::Unsupported);
            }
        }
        // a meta file with an empty map and no matrix (what older builds left after add_vectors([])) is an empty store
        make_dir(base + "/orphan");
        { std::ofstream f(base + "/orphan/vectors.meta.json"); f << "{}"; }
        auto orphan = B200Store::load(base + "/orphan");
        CHECK(orphan->len() == 0 && orphan->search({1.f, 0.f}, 3).empty());
        drop_vector_storage_registry();
    }
    // ShardedB200Store: ONE process, several GPUs (or, on a 1-GPU box, two shards on the same device), rendezvous and
    // exchange below the C ABI (mx_shard_group).  Answers must equal the single store's: same doc ids, same score bits.
    {
        const int ndev = mx_device_count();
        ShardedB200Store::Options so;
        so.devices = ndev >= 2 ? std::vector<int>{0, 1} : std::vector<int>{0, 0};
        if (ndev >= 4) so.devices = {0, 1, 2, 3};
        so.fp16 = true;
        auto sharded = ShardedB200Store::new_(tmp + "/sharded", so);
        B200Store::Options o1;
        o1.fp16 = true;
        o1.save_on_insert = false;
        auto single = B200Store::new_(tmp + "/sharded_ref", o1);
        std::vector<VectorData> pts;
        uint32_t rng = 12345;
        auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return ((rng >> 8) & 0xffff) / 65536.0f - 0.5f; };
        const int n = 6001, d = 96;
        for (int i = 0; i < n; ++i) {
            VectorData v;
            v._id = "doc-" + std::to_string(i);
            v.vector.resize(d);
            for (auto &x : v.vector) x = rnd();
            pts.push_back(v);
        }
        pts[4000].vector = pts[17].vector;     // exact ties across shards: the lower global id must come first
        pts[4001].vector = pts[17].vector;
        // two batches, the second starting at an odd global index
        sharded->bulk_insert(std::vector<VectorData>(pts.begin(), pts.begin() + 2501));
        sharded->bulk_insert(std::vector<VectorData>(pts.begin() + 2501, pts.end()));
        single->bulk_insert(pts);
        CHECK(sharded->len() == (uint64_t)n && sharded->_id_map.size() == (size_t)n);
        for (int qi : {17, 0, 2500, 4001, 6000}) {
            auto a = sharded->search(pts[qi].vector, 10), b = single->search(pts[qi].vector, 10);
            CHECK(a.size() == 10 && a == b);
        }
        auto tie = sharded->search(pts[17].vector, 3);
        CHECK(tie[0].first == "doc-17" && tie[1].first == "doc-4000" && tie[2].first == "doc-4001");
        std::vector<std::vector<float>> batch;
        for (int i = 0; i < 20; ++i) batch.push_back(pts[i * 300].vector);     // >= 8 queries: the tcgen05 scan
        auto ra = sharded->search_batch(batch, 10), rb = single->search_batch(batch, 10);
        CHECK(ra == rb);
        // save / load with the same shard count
        sharded->save(tmp + "/sharded");
        auto again = ShardedB200Store::load(tmp + "/sharded", so);
        CHECK(again->len() == (uint64_t)n);
        CHECK(again->search_batch(batch, 10) == rb);
        again->delete_all();
        CHECK(again->len() == 0 && !ShardedB200Store::has_store(tmp + "/sharded"));
        std::printf("sharded store ok over %zu shard(s) on %d visible device(s)\n", so.devices.size(), ndev);
    }
    std::printf("gpu ok: %d checks\n", g_checks);
    return 0;
}

static int run_encode(int argc, char **argv)
{
    if (argc < 12) return 2;
    Architecture arch{(uint32_t)atoi(argv[5]), (uint32_t)atoi(argv[6]), (uint32_t)atoi(argv[7]), (uint32_t)atoi(argv[8]),
                      (uint32_t)atoi(argv[9]), (uint32_t)atoi(argv[10]), 2, 1e-12f, true, 256};
    Weights w = Weights::from_safetensors(argv[2]);
    // ids.bin: u32 B, u32 S, then B*S i32 ids, then B i32 lens
    std::ifstream f(argv[3], std::ios::binary);
    TokenBatch tb;
    f.read(reinterpret_cast<char *>(&tb.B), 4);
    f.read(reinterpret_cast<char *>(&tb.S), 4);
    tb.ids.resize((size_t)tb.B * tb.S);
    tb.lens.resize(tb.B);
    f.read(reinterpret_cast<char *>(tb.ids.data()), (std::streamsize)tb.ids.size() * 4);
    f.read(reinterpret_cast<char *>(tb.lens.data()), (std::streamsize)tb.lens.size() * 4);
    B200Encoder enc(arch, w, (B200Encoder::Precision)atoi(argv[11]), 0, tb.B * tb.S);
    std::vector<float> out = enc.encode_ids(tb);
    std::ofstream o(argv[4], std::ios::binary);
    o.write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size() * 4);
    std::printf("encode ok: %u x %u -> %zu floats\n", tb.B, tb.S, out.size());
    return 0;
}

int main(int argc, char **argv)
{
    try {
        if (argc >= 3 && !std::strcmp(argv[1], "tokenizer")) return run_tokenizer(argv[2]);
        if (argc >= 3 && !std::strcmp(argv[1], "bpe")) return run_bpe(argv[2]);
        if (argc >= 2 && !std::strcmp(argv[1], "cpu")) return run_cpu();
        if (argc >= 3 && !std::strcmp(argv[1], "gpu")) return run_gpu(argv[2]);
        if (argc >= 2 && !std::strcmp(argv[1], "encode")) return run_encode(argc, argv);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "uncaught: %s\n", e.what());
        return 1;
    }
    std::fprintf(stderr, "usage: test_host tokenizer <golden.json> | cpu | gpu <tmpdir> | encode ...\n");
    return 2;
}
