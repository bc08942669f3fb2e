// store.cpp -- B200Store / VectorStorage / get_vector_storage / SearchBatcher over the C ABI.
// Mirrors reference lib/libmemex/src/storage/{mod.rs,local.rs}; see memex_host.hpp.
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../../include/memex_b200.h"
#include "json.hpp"
#include "memex_host.hpp"

namespace memex {

namespace {

const char *kMetaFile = "vectors.meta.json";   // local.rs:19, byte-compatible: {"<usize>":"<id>",...}

StoreErrorKind kind_of(int32_t code, StoreErrorKind fallback)
{
    switch (code) {
        case MX_ERR_CONNECTION: return StoreErrorKind::ConnectionError;
        case MX_ERR_DELETE: return StoreErrorKind::DeleteError;
        case MX_ERR_FILE_IO: return StoreErrorKind::FileIOError;
        case MX_ERR_INSERTION: return StoreErrorKind::InsertionError;
        case MX_ERR_SEARCH: return StoreErrorKind::SearchError;
        case MX_ERR_SERDE: return StoreErrorKind::SerdeError;
        case MX_ERR_SAVE: return StoreErrorKind::SaveError;
        case MX_ERR_UNSUPPORTED: return StoreErrorKind::Unsupported;
        default: return fallback;
    }
}

[[noreturn]] void raise(int32_t code, const void *handle, StoreErrorKind fallback)
{
    const char *m = mx_last_error(handle);
    throw VectorStoreError(kind_of(code, fallback), m && *m ? m : ("status " + std::to_string(code)));
}

bool exists(const std::string &p)
{
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
}

void make_dirs(const std::string &path)
{
    std::string cur;
    for (size_t i = 0; i <= path.size(); ++i) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty() && !exists(cur) && ::mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST)
                throw VectorStoreError(StoreErrorKind::FileIOError, cur + ": " + std::strerror(errno));
        }
        if (i < path.size()) cur += path[i];
    }
}

std::string join(const std::string &a, const std::string &b)
{
    if (a.empty()) return b;
    return a.back() == '/' ? a + b : a + "/" + b;
}

}  // namespace

const char *to_string(StoreErrorKind k)
{
    switch (k) {
        case StoreErrorKind::ConnectionError: return "Unable to connect";
        case StoreErrorKind::DeleteError: return "DeleteError";
        case StoreErrorKind::FileIOError: return "File IO error";
        case StoreErrorKind::InsertionError: return "Unable to insert vector";
        case StoreErrorKind::SearchError: return "Unable to search";
        case StoreErrorKind::SerdeError: return "Unable to deserialize";
        case StoreErrorKind::SaveError: return "Unable to save db file";
        case StoreErrorKind::Unsupported: return "Unsupported vector db";
    }
    return "?";
}

VectorStoreError::VectorStoreError(StoreErrorKind k, const std::string &msg)
    : std::runtime_error(std::string(to_string(k)) + ": " + msg), kind(k)
{
}

std::vector<std::vector<VectorSearchResult>> VectorStore::search_batch(const std::vector<std::vector<float>> &vecs, size_t limit) const
{
    std::vector<std::vector<VectorSearchResult>> out;
    out.reserve(vecs.size());
    for (const auto &v : vecs) out.push_back(search(v, limit));
    return out;
}

uint64_t VectorStore::search_batch_submit(const std::vector<std::vector<float>> &vecs, size_t limit) const
{
    const uint64_t t = next_ticket_++;
    parked_.emplace(t, search_batch(vecs, limit));
    return t;
}

std::vector<std::vector<VectorSearchResult>> VectorStore::search_batch_collect(uint64_t ticket) const
{
    auto it = parked_.find(ticket);
    if (it == parked_.end()) throw VectorStoreError(StoreErrorKind::SearchError, "no search with ticket " + std::to_string(ticket));
    auto out = std::move(it->second);
    parked_.erase(it);
    return out;
}

// ------------------------------------------------------------------------------------------------
// B200Store
// ------------------------------------------------------------------------------------------------
std::unique_ptr<B200Store> B200Store::new_(const std::string &storage_path, const Options &opt)
{
    std::unique_ptr<B200Store> s(new B200Store());
    s->storage_path = storage_path;
    s->options = opt;
    if (opt.dim) s->create_handle(opt.dim);
    return s;
}

void B200Store::create_handle(uint32_t dim)
{
    mx_store_cfg cfg{};
    cfg.dim = dim;
    cfg.dtype = options.fp16 ? MX_DTYPE_F16 : MX_DTYPE_F32;
    cfg.metric = options.dot ? MX_METRIC_DOT : MX_METRIC_COSINE;
    cfg.device = options.device;
    cfg.capacity = options.capacity;
    cfg.id_offset = 0;
    cfg.id_stride = 1;
    int32_t rc = mx_store_create(&cfg, &handle_);
    if (rc != MX_OK) raise(rc, nullptr, StoreErrorKind::ConnectionError);
    options.dim = dim;
}

bool B200Store::has_store(const std::string &store_path) { return exists(join(store_path, kMetaFile)); }

std::unique_ptr<B200Store> B200Store::load(const std::string &store_path, int device)
{
    std::unique_ptr<B200Store> s(new B200Store());
    s->storage_path = store_path;
    s->options.device = device;
    const bool has_bin = mx_store_has_file(store_path.c_str()) != 0;
    if (has_bin) {
        int32_t rc = mx_store_load(store_path.c_str(), device, &s->handle_);
        if (rc != MX_OK) raise(rc, nullptr, StoreErrorKind::FileIOError);
    }
    std::ifstream f(join(store_path, kMetaFile), std::ios::binary);
    if (!f) throw VectorStoreError(StoreErrorKind::FileIOError, join(store_path, kMetaFile) + ": " + std::strerror(errno));
    std::stringstream ss;
    ss << f.rdbuf();
    try {
        json::Value v = json::parse(ss.str());
        if (v.kind != json::Value::Object) throw std::runtime_error("expected an object");
        for (const auto &kv : v.obj) {
            if (kv.second.kind != json::Value::String) throw std::runtime_error("expected string values");
            char *end = nullptr;
            unsigned long long id = std::strtoull(kv.first.c_str(), &end, 10);
            if (end == kv.first.c_str() || *end) throw std::runtime_error("key is not an integer: " + kv.first);
            s->_id_map[(size_t)id] = kv.second.str;
        }
    } catch (const std::runtime_error &e) {
        throw VectorStoreError(StoreErrorKind::SerdeError, e.what());
    }
    if (!has_bin) {
        // a meta file without its matrix: only an EMPTY map is a store (one that never received a row); anything
        // else is the reference's load failure (local.rs:127-129)
        if (!s->_id_map.empty())
            throw VectorStoreError(StoreErrorKind::FileIOError, join(store_path, "vectors.b200.bin") + ": No such file or directory");
        return s;
    }
    uint32_t dim = 0, dtype = 0, metric = 0;
    uint64_t cap = 0;
    mx_store_info(s->handle_, &dim, &dtype, &metric, &cap);
    s->options.dim = dim;
    s->options.fp16 = dtype == MX_DTYPE_F16;
    s->options.dot = metric == MX_METRIC_DOT;
    s->options.device = device;
    return s;
}

void B200Store::save(const std::string &store_path) const
{
    if (!handle_) return;   // nothing was ever inserted: no files, so has_store() stays false (as a fresh HnswStore dir)
    make_dirs(store_path);
    int32_t rc = mx_store_save(handle_, store_path.c_str());
    if (rc != MX_OK) raise(rc, handle_, StoreErrorKind::SaveError);
    // the id map as serde_json::to_string(&HashMap<usize, String>) renders it (local.rs:155-161)
    std::string out = "{";
    bool first = true;
    for (const auto &kv : _id_map) {
        if (!first) out += ',';
        first = false;
        out += '"' + std::to_string(kv.first) + "\":";
        json::escape_into(kv.second, out);
    }
    out += '}';
    const std::string tmp = join(store_path, std::string(kMetaFile) + ".tmp");
    {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        if (!f) throw VectorStoreError(StoreErrorKind::FileIOError, tmp + ": " + std::strerror(errno));
        f << out;
        f.flush();
        if (!f) throw VectorStoreError(StoreErrorKind::FileIOError, tmp + ": write failed");
    }
    if (std::rename(tmp.c_str(), join(store_path, kMetaFile).c_str()) != 0)
        throw VectorStoreError(StoreErrorKind::FileIOError, std::string("rename: ") + std::strerror(errno));
}

B200Store::~B200Store()
{
    if (handle_) mx_store_destroy(handle_);
}

void B200Store::delete_(const std::string &)
{
    // local.rs:29-32 is `unimplemented!()`: a panic.  The ABI reports it instead.
    if (!handle_) throw VectorStoreError(StoreErrorKind::Unsupported, "removing a single point is not supported by the file store");
    raise(mx_store_delete(handle_, 0), handle_, StoreErrorKind::Unsupported);
}

void B200Store::delete_all()
{
    mx_store_remove_file(storage_path.c_str());                 // the data file ...
    std::remove(join(storage_path, kMetaFile).c_str());         // ... and the id map (local.rs:36-46)
    if (handle_) {
        int32_t rc = mx_store_clear(handle_);
        if (rc != MX_OK) raise(rc, handle_, StoreErrorKind::DeleteError);
    }
    _id_map.clear();
}

void B200Store::bulk_insert(const std::vector<VectorData> &data)
{
    if (data.empty()) return;
    if (!handle_) {
        if (data[0].vector.empty()) throw VectorStoreError(StoreErrorKind::InsertionError, "empty vector");
        create_handle((uint32_t)data[0].vector.size());
    }
    const size_t dim = options.dim;
    std::vector<float> rows(data.size() * dim);
    for (size_t i = 0; i < data.size(); ++i) {
        if (data[i].vector.size() != dim)
            throw VectorStoreError(StoreErrorKind::InsertionError, "vector has dimension " + std::to_string(data[i].vector.size()) +
                                                                       ", store has " + std::to_string(dim));
        std::copy(data[i].vector.begin(), data[i].vector.end(), rows.begin() + i * dim);
    }
    uint64_t first = 0;
    int32_t rc = mx_store_add(handle_, rows.data(), data.size(), &first);
    if (rc != MX_OK) raise(rc, handle_, StoreErrorKind::InsertionError);
    // next_id = _id_map.len() + 1 (local.rs:63): the device's row ids and the map stay in lockstep
    for (size_t i = 0; i < data.size(); ++i) _id_map[(size_t)first + i] = data[i]._id;
    if (options.save_on_insert) {
        try {
            save(storage_path);   // `let _ = self.save(..)`: errors are ignored there too (local.rs:66-67)
        } catch (const VectorStoreError &) {
        }
    }
}

void B200Store::insert(const VectorData &data) { bulk_insert({data}); }

std::vector<float> B200Store::pack_queries(const std::vector<std::vector<float>> &vecs, size_t limit) const
{
    // one behaviour in every host (C++, Python, Rust): more than MX_MAX_K neighbours is an error, never a silent cut
    if (limit > MX_MAX_K)
        throw VectorStoreError(StoreErrorKind::SearchError, "limit " + std::to_string(limit) + " exceeds the store's maximum of " +
                                                                std::to_string(MX_MAX_K) + " neighbours per query");
    const size_t dim = options.dim, nq = vecs.size();
    std::vector<float> q(nq * dim);
    for (size_t i = 0; i < nq; ++i) {
        if (vecs[i].size() != dim)
            throw VectorStoreError(StoreErrorKind::SearchError, "query has dimension " + std::to_string(vecs[i].size()) +
                                                                    ", store has " + std::to_string(dim));
        std::copy(vecs[i].begin(), vecs[i].end(), q.begin() + i * dim);
    }
    return q;
}

std::vector<std::vector<VectorSearchResult>> B200Store::map_results(const uint64_t *ids, const float *scores,
                                                                    const uint32_t *counts, size_t nq, uint32_t k) const
{
    std::vector<std::vector<VectorSearchResult>> out(nq);
    for (size_t i = 0; i < nq; ++i) {
        out[i].reserve(counts[i]);
        for (uint32_t j = 0; j < counts[i]; ++j) {
            auto it = _id_map.find((size_t)ids[i * k + j]);
            if (it == _id_map.end())   // local.rs:80-83 panics here
                throw VectorStoreError(StoreErrorKind::SearchError, "Internal inconsistency. Id from vector store not mapped.");
            out[i].emplace_back(it->second, scores[i * k + j]);
        }
    }
    return out;
}

std::vector<std::vector<VectorSearchResult>> B200Store::search_batch(const std::vector<std::vector<float>> &vecs, size_t limit) const
{
    std::vector<std::vector<VectorSearchResult>> out(vecs.size());
    if (vecs.empty() || limit == 0 || _id_map.empty() || !handle_) return out;
    const std::vector<float> q = pack_queries(vecs, limit);
    const size_t nq = vecs.size();
    const uint32_t k = (uint32_t)limit;
    std::vector<uint64_t> ids(nq * k);
    std::vector<float> scores(nq * k);
    std::vector<uint32_t> counts(nq);
    int32_t rc = mx_store_search(handle_, q.data(), (uint32_t)nq, k, ids.data(), scores.data(), counts.data());
    if (rc != MX_OK) raise(rc, handle_, StoreErrorKind::SearchError);
    return map_results(ids.data(), scores.data(), counts.data(), nq, k);
}

uint64_t B200Store::search_batch_submit(const std::vector<std::vector<float>> &vecs, size_t limit) const
{
    // nothing for the device to do: park the empty answer (same ticket space as the device searches)
    if (vecs.empty() || limit == 0 || _id_map.empty() || !handle_) return VectorStore::search_batch_submit(vecs, limit);
    const std::vector<float> q = pack_queries(vecs, limit);
    uint64_t dev_ticket = 0;
    int32_t rc = mx_store_search_submit(handle_, q.data(), (uint32_t)vecs.size(), (uint32_t)limit, &dev_ticket);
    if (rc != MX_OK) raise(rc, handle_, StoreErrorKind::SearchError);
    const uint64_t t = next_ticket_++;
    in_flight_.emplace(t, InFlight{dev_ticket, vecs.size(), (uint32_t)limit});
    return t;
}

std::vector<std::vector<VectorSearchResult>> B200Store::search_batch_collect(uint64_t ticket) const
{
    auto it = in_flight_.find(ticket);
    if (it == in_flight_.end()) return VectorStore::search_batch_collect(ticket);
    const InFlight f = it->second;
    in_flight_.erase(it);
    std::vector<uint64_t> ids(f.nq * f.k);
    std::vector<float> scores(f.nq * f.k);
    std::vector<uint32_t> counts(f.nq);
    int32_t rc = mx_store_search_collect(handle_, f.device_ticket, ids.data(), scores.data(), counts.data());
    if (rc != MX_OK) raise(rc, handle_, StoreErrorKind::SearchError);
    return map_results(ids.data(), scores.data(), counts.data(), f.nq, f.k);
}

std::vector<VectorSearchResult> B200Store::search(const std::vector<float> &vec, size_t limit) const
{
    return search_batch({vec}, limit)[0];
}

uint64_t B200Store::len() const
{
    uint64_t n = 0;
    if (handle_) mx_store_len(handle_, &n);
    return n;
}

// ------------------------------------------------------------------------------------------------
// ShardedB200Store: one process, G GPUs
// ------------------------------------------------------------------------------------------------
namespace {
std::string shard_dir(const std::string &base, size_t g) { return join(base, "shard-" + std::to_string(g)); }

void write_id_map(const std::string &store_path, const std::unordered_map<size_t, std::string> &id_map)
{
    std::string out = "{";
    bool first = true;
    for (const auto &kv : id_map) {
        if (!first) out += ',';
        first = false;
        out += '"' + std::to_string(kv.first) + "\":";
        json::escape_into(kv.second, out);
    }
    out += '}';
    const std::string tmp = join(store_path, std::string(kMetaFile) + ".tmp");
    {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        if (!f) throw VectorStoreError(StoreErrorKind::FileIOError, tmp + ": " + std::strerror(errno));
        f << out;
        f.flush();
        if (!f) throw VectorStoreError(StoreErrorKind::FileIOError, tmp + ": write failed");
    }
    if (std::rename(tmp.c_str(), join(store_path, kMetaFile).c_str()) != 0)
        throw VectorStoreError(StoreErrorKind::FileIOError, std::string("rename: ") + std::strerror(errno));
}

std::unordered_map<size_t, std::string> read_id_map(const std::string &store_path)
{
    std::ifstream f(join(store_path, kMetaFile), std::ios::binary);
    if (!f) throw VectorStoreError(StoreErrorKind::FileIOError, join(store_path, kMetaFile) + ": " + std::strerror(errno));
    std::stringstream ss;
    ss << f.rdbuf();
    std::unordered_map<size_t, std::string> out;
    try {
        json::Value v = json::parse(ss.str());
        if (v.kind != json::Value::Object) throw std::runtime_error("expected an object");
        for (const auto &kv : v.obj) {
            if (kv.second.kind != json::Value::String) throw std::runtime_error("expected string values");
            char *end = nullptr;
            unsigned long long id = std::strtoull(kv.first.c_str(), &end, 10);
            if (end == kv.first.c_str() || *end) throw std::runtime_error("key is not an integer: " + kv.first);
            out[(size_t)id] = kv.second.str;
        }
    } catch (const std::runtime_error &e) {
        throw VectorStoreError(StoreErrorKind::SerdeError, e.what());
    }
    return out;
}
}  // namespace

std::unique_ptr<ShardedB200Store> ShardedB200Store::new_(const std::string &storage_path, const Options &opt)
{
    if (opt.devices.empty() || opt.devices.size() > 16)
        throw VectorStoreError(StoreErrorKind::ConnectionError, "a sharded store needs 1..16 devices");
    std::unique_ptr<ShardedB200Store> s(new ShardedB200Store());
    s->storage_path = storage_path;
    s->options = opt;
    if (opt.dim) s->create_handles(opt.dim);
    return s;
}

void ShardedB200Store::create_handles(uint32_t dim)
{
    const size_t G = options.devices.size();
    for (size_t g = 0; g < G; ++g) {
        mx_store_cfg cfg{};
        cfg.dim = dim;
        cfg.dtype = options.fp16 ? MX_DTYPE_F16 : MX_DTYPE_F32;
        cfg.metric = options.dot ? MX_METRIC_DOT : MX_METRIC_COSINE;
        cfg.device = options.devices[g];
        cfg.capacity = options.capacity_per_shard;
        cfg.id_offset = g;          // global id = local_row * G + g + 1: row i of the insertion order -> shard i % G
        cfg.id_stride = G;
        mx_store *h = nullptr;
        int32_t rc = mx_store_create(&cfg, &h);
        if (rc != MX_OK) raise(rc, nullptr, StoreErrorKind::ConnectionError);
        stores_.push_back(h);
        mx_shard_group *grp = nullptr;
        rc = mx_shard_group_create(options.devices[g], (uint32_t)G, (uint32_t)g, dim, std::max<uint32_t>(1, options.max_batch), MX_MAX_K, &grp);
        if (rc != MX_OK) raise(rc, nullptr, StoreErrorKind::ConnectionError);
        groups_.push_back(grp);
    }
    int32_t rc = mx_shard_group_connect_local(groups_.data(), (uint32_t)G);
    if (rc != MX_OK) raise(rc, groups_[0], StoreErrorKind::ConnectionError);
    options.dim = dim;
}

ShardedB200Store::~ShardedB200Store()
{
    for (mx_shard_group *g : groups_) mx_shard_group_destroy(g);
    for (mx_store *h : stores_) mx_store_destroy(h);
}

std::unique_ptr<ShardedB200Store> ShardedB200Store::load(const std::string &store_path, const Options &opt)
{
    if (opt.devices.empty()) throw VectorStoreError(StoreErrorKind::ConnectionError, "a sharded store needs devices");
    std::unique_ptr<ShardedB200Store> s(new ShardedB200Store());
    s->storage_path = store_path;
    s->options = opt;
    s->_id_map = read_id_map(store_path);
    const size_t G = opt.devices.size();
    if (!mx_store_has_file(shard_dir(store_path, 0).c_str())) {
        if (!s->_id_map.empty())
            throw VectorStoreError(StoreErrorKind::FileIOError, join(shard_dir(store_path, 0), "vectors.b200.bin") + ": No such file or directory");
        return s;
    }
    // the shard files hold rows in local order; ids are rebuilt from (g, G), so the files must be read with the G they
    // were written with
    uint32_t dim = 0;
    for (size_t g = 0; g < G; ++g) {
        mx_store *tmp = nullptr;
        int32_t rc = mx_store_load(shard_dir(store_path, g).c_str(), opt.devices[g], &tmp);
        if (rc != MX_OK) raise(rc, nullptr, StoreErrorKind::FileIOError);
        uint32_t d = 0, dtype = 0, metric = 0;
        uint64_t cap = 0, n = 0;
        mx_store_info(tmp, &d, &dtype, &metric, &cap);
        mx_store_len(tmp, &n);
        if (g == 0) {
            dim = d;
            s->options.fp16 = dtype == MX_DTYPE_F16;
            s->options.dot = metric == MX_METRIC_DOT;
            s->create_handles(dim);
        }
        // re-ingest through the sharded handles (global ids need id_offset / id_stride, which a plain load does not carry)
        std::vector<float> rows((size_t)std::min<uint64_t>(n, 65536) * d);
        for (uint64_t done = 0; done < n;) {
            const uint64_t c = std::min<uint64_t>(65536, n - done);
            rc = mx_store_get_rows(tmp, done, c, rows.data());
            if (rc == MX_OK) rc = mx_store_add(s->stores_[g], rows.data(), c, nullptr);
            if (rc != MX_OK) {
                mx_store_destroy(tmp);
                raise(rc, s->stores_[g], StoreErrorKind::FileIOError);
            }
            done += c;
        }
        mx_store_destroy(tmp);
    }
    if (s->len() != s->_id_map.size())
        throw VectorStoreError(StoreErrorKind::SerdeError, "shard files hold " + std::to_string(s->len()) + " rows, the id map " +
                                                               std::to_string(s->_id_map.size()) + " (written with another shard count?)");
    return s;
}

void ShardedB200Store::save(const std::string &store_path) const
{
    if (stores_.empty()) return;
    make_dirs(store_path);
    for (size_t g = 0; g < stores_.size(); ++g) {
        make_dirs(shard_dir(store_path, g));
        int32_t rc = mx_store_save(stores_[g], shard_dir(store_path, g).c_str());
        if (rc != MX_OK) raise(rc, stores_[g], StoreErrorKind::SaveError);
    }
    write_id_map(store_path, _id_map);
}

void ShardedB200Store::delete_(const std::string &)
{
    throw VectorStoreError(StoreErrorKind::Unsupported, "removing a single point is not supported by the file store");
}

void ShardedB200Store::delete_all()
{
    for (size_t g = 0; g < options.devices.size(); ++g) mx_store_remove_file(shard_dir(storage_path, g).c_str());
    std::remove(join(storage_path, kMetaFile).c_str());
    for (mx_store *h : stores_) {
        int32_t rc = mx_store_clear(h);
        if (rc != MX_OK) raise(rc, h, StoreErrorKind::DeleteError);
    }
    _id_map.clear();
}

void ShardedB200Store::bulk_insert(const std::vector<VectorData> &data)
{
    if (data.empty()) return;
    if (stores_.empty()) {
        if (data[0].vector.empty()) throw VectorStoreError(StoreErrorKind::InsertionError, "empty vector");
        create_handles((uint32_t)data[0].vector.size());
    }
    const size_t dim = options.dim, G = stores_.size();
    for (const auto &d : data)
        if (d.vector.size() != dim)
            throw VectorStoreError(StoreErrorKind::InsertionError, "vector has dimension " + std::to_string(d.vector.size()) +
                                                                       ", store has " + std::to_string(dim));
    const size_t n0 = _id_map.size();   // rows so far == next global index (next_id = len + 1, local.rs:63)
    std::vector<std::vector<float>> part(G);
    for (size_t i = 0; i < data.size(); ++i) {
        auto &p = part[(n0 + i) % G];
        p.insert(p.end(), data[i].vector.begin(), data[i].vector.end());
    }
    for (size_t g = 0; g < G; ++g) {
        if (part[g].empty()) continue;
        uint64_t first = 0;
        int32_t rc = mx_store_add(stores_[g], part[g].data(), part[g].size() / dim, &first);
        if (rc != MX_OK) raise(rc, stores_[g], StoreErrorKind::InsertionError);
    }
    for (size_t i = 0; i < data.size(); ++i) _id_map[n0 + i + 1] = data[i]._id;
    if (options.save_on_insert) {
        try {
            save(storage_path);
        } catch (const VectorStoreError &) {
        }
    }
}

void ShardedB200Store::insert(const VectorData &data) { bulk_insert({data}); }

std::vector<std::vector<VectorSearchResult>> ShardedB200Store::search_batch(const std::vector<std::vector<float>> &vecs, size_t limit) const
{
    std::vector<std::vector<VectorSearchResult>> out(vecs.size());
    if (vecs.empty() || limit == 0 || _id_map.empty() || stores_.empty()) return out;
    if (limit > MX_MAX_K)
        throw VectorStoreError(StoreErrorKind::SearchError, "limit " + std::to_string(limit) + " exceeds the store's maximum of " +
                                                                std::to_string(MX_MAX_K) + " neighbours per query");
    const size_t dim = options.dim;
    const uint32_t k = (uint32_t)limit;
    const size_t step = std::max<uint32_t>(1, options.max_batch);
    for (size_t q0 = 0; q0 < vecs.size(); q0 += step) {
        const size_t nq = std::min(step, vecs.size() - q0);
        std::vector<float> q(nq * dim);
        for (size_t i = 0; i < nq; ++i) {
            if (vecs[q0 + i].size() != dim)
                throw VectorStoreError(StoreErrorKind::SearchError, "query has dimension " + std::to_string(vecs[q0 + i].size()) +
                                                                        ", store has " + std::to_string(dim));
            std::copy(vecs[q0 + i].begin(), vecs[q0 + i].end(), q.begin() + i * dim);
        }
        std::vector<uint64_t> ids(nq * k);
        std::vector<float> scores(nq * k);
        std::vector<uint32_t> counts(nq);
        int32_t rc = mx_shard_group_search_local(groups_.data(), stores_.data(), (uint32_t)stores_.size(), q.data(), (uint32_t)nq, k,
                                                 ids.data(), scores.data(), counts.data());
        if (rc != MX_OK) raise(rc, groups_[0], StoreErrorKind::SearchError);
        for (size_t i = 0; i < nq; ++i) {
            out[q0 + i].reserve(counts[i]);
            for (uint32_t j = 0; j < counts[i]; ++j) {
                auto it = _id_map.find((size_t)ids[i * k + j]);
                if (it == _id_map.end())
                    throw VectorStoreError(StoreErrorKind::SearchError, "Internal inconsistency. Id from vector store not mapped.");
                out[q0 + i].emplace_back(it->second, scores[i * k + j]);
            }
        }
    }
    return out;
}

std::vector<VectorSearchResult> ShardedB200Store::search(const std::vector<float> &vec, size_t limit) const
{
    return search_batch({vec}, limit)[0];
}

uint64_t ShardedB200Store::len() const
{
    uint64_t total = 0;
    for (mx_store *h : stores_) {
        uint64_t n = 0;
        mx_store_len(h, &n);
        total += n;
    }
    return total;
}

// ------------------------------------------------------------------------------------------------
// VectorStorage / factory / registry
// ------------------------------------------------------------------------------------------------
VectorStorage::VectorStorage(std::shared_ptr<VectorStore> c) : client(std::move(c)), lock_(std::make_shared<std::mutex>()) {}

void VectorStorage::add_vectors(const std::vector<VectorData> &points)
{
    std::lock_guard<std::mutex> g(*lock_);
    client->bulk_insert(points);
}
void VectorStorage::delete_collection()
{
    std::lock_guard<std::mutex> g(*lock_);
    client->delete_all();
}
std::vector<VectorSearchResult> VectorStorage::search(const std::vector<float> &query, size_t limit) const
{
    std::lock_guard<std::mutex> g(*lock_);
    return client->search(query, limit);
}
std::vector<std::vector<VectorSearchResult>> VectorStorage::search_batch(const std::vector<std::vector<float>> &queries, size_t limit) const
{
    std::lock_guard<std::mutex> g(*lock_);
    return client->search_batch(queries, limit);
}
uint64_t VectorStorage::search_batch_submit(const std::vector<std::vector<float>> &queries, size_t limit) const
{
    std::lock_guard<std::mutex> g(*lock_);
    return client->search_batch_submit(queries, limit);
}
std::vector<std::vector<VectorSearchResult>> VectorStorage::search_batch_collect(uint64_t ticket) const
{
    std::lock_guard<std::mutex> g(*lock_);
    return client->search_batch_collect(ticket);
}

namespace {
std::mutex g_registry_mu;
std::unordered_map<std::string, VectorStorage> &registry()
{
    static std::unordered_map<std::string, VectorStorage> r;
    return r;
}
}  // namespace

void drop_vector_storage_registry()
{
    std::lock_guard<std::mutex> g(g_registry_mu);
    registry().clear();
}

VectorStorage get_vector_storage(const std::string &uri, const std::string &collection)
{
    // mod.rs:99-102: anything that does not parse as scheme://rest is Unsupported
    const size_t sep = uri.find("://");
    if (sep == std::string::npos || sep == 0) throw VectorStoreError(StoreErrorKind::Unsupported, uri);
    const std::string scheme = uri.substr(0, sep);
    if (scheme != "b200" && scheme != "b200+f16" && scheme != "b200+f32") throw VectorStoreError(StoreErrorKind::Unsupported, uri);
    // "<dir>?key=value&..." -- the options never reach the directory name
    std::string rest = uri.substr(sep + 3);
    B200Store::Options opt;
    opt.fp16 = scheme == "b200+f16";
    const size_t qm = rest.find('?');
    if (qm != std::string::npos) {
        std::string query = rest.substr(qm + 1);
        rest.resize(qm);
        size_t pos = 0;
        while (pos <= query.size()) {
            size_t amp = query.find('&', pos);
            if (amp == std::string::npos) amp = query.size();
            const std::string kv = query.substr(pos, amp - pos);
            pos = amp + 1;
            if (kv.empty()) continue;
            const size_t eq = kv.find('=');
            const std::string k = kv.substr(0, eq), v = eq == std::string::npos ? "" : kv.substr(eq + 1);
            auto number = [&](uint64_t max) -> uint64_t {
                char *end = nullptr;
                errno = 0;
                unsigned long long x = std::strtoull(v.c_str(), &end, 10);
                if (v.empty() || *end || errno || x > max) throw VectorStoreError(StoreErrorKind::Unsupported, uri + " (bad value for " + k + ")");
                return x;
            };
            if (k == "dtype" && (v == "f16" || v == "f32")) opt.fp16 = v == "f16";
            else if (k == "metric" && (v == "cosine" || v == "dot")) opt.dot = v == "dot";
            else if (k == "device") opt.device = (int)number(1023);
            else if (k == "dim") opt.dim = (uint32_t)number(1u << 20);
            else throw VectorStoreError(StoreErrorKind::Unsupported, uri + " (unknown option " + kv + ")");
        }
    }
    const std::string key = uri + "\n" + collection;
    std::lock_guard<std::mutex> g(g_registry_mu);
    auto it = registry().find(key);
    if (it != registry().end()) return it->second;
    // collections are stored as folders (mod.rs:109-113)
    const std::string storage = join(rest, collection);
    make_dirs(storage);
    std::shared_ptr<VectorStore> store;
    if (B200Store::has_store(storage)) {
        store = B200Store::load(storage, opt.device);
    } else {
        store = B200Store::new_(storage, opt);
    }
    VectorStorage vs(store);
    registry().emplace(key, vs);
    return vs;
}

// ------------------------------------------------------------------------------------------------
// SearchBatcher
// ------------------------------------------------------------------------------------------------
SearchBatcher::SearchBatcher(VectorStorage storage, size_t max_batch, uint32_t max_wait_us)
    : storage_(std::move(storage)), max_batch_(std::max<size_t>(1, max_batch)), max_wait_us_(max_wait_us)
{
    worker_ = std::thread([this] { run(); });
}

SearchBatcher::~SearchBatcher()
{
    {
        std::lock_guard<std::mutex> g(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (worker_.joinable()) worker_.join();
}

std::future<std::vector<VectorSearchResult>> SearchBatcher::submit(std::vector<float> query, size_t limit)
{
    Req r{std::move(query), limit, {}};
    auto fut = r.done.get_future();
    {
        std::lock_guard<std::mutex> g(mu_);
        queue_.push_back(std::move(r));
    }
    cv_.notify_one();
    return fut;
}

void SearchBatcher::run()
{
    struct Pending {
        uint64_t ticket;
        std::vector<Req> batch;
    };
    std::unique_ptr<Pending> pending;   // the batch the device is answering while the next one is gathered
    auto deliver = [this](Pending &p) {
        ++batches_;   // before the promises: a caller that has its answer sees the batch counted
        try {
            auto res = storage_.search_batch_collect(p.ticket);
            for (size_t i = 0; i < p.batch.size(); ++i) p.batch[i].done.set_value(std::move(res[i]));
        } catch (...) {
            for (auto &r : p.batch) r.done.set_exception(std::current_exception());
        }
    };
    while (true) {
        std::vector<Req> batch;
        {
            std::unique_lock<std::mutex> g(mu_);
            if (pending) {
                // a batch is on the device: take whatever has queued up meanwhile, but do not sit here waiting for more
                if (queue_.empty()) {
                    g.unlock();
                    deliver(*pending);
                    pending.reset();
                    continue;
                }
            } else {
                cv_.wait(g, [this] { return stop_ || !queue_.empty(); });
                if (queue_.empty()) return;   // stop requested and nothing left
                // the first request opens a window of max_wait_us for others to join
                const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(max_wait_us_);
                while (queue_.size() < max_batch_ && !stop_) {
                    if (cv_.wait_until(g, deadline) == std::cv_status::timeout) break;
                }
            }
            // one scan serves one `limit`: take the head's and everyone who asked for the same
            const size_t limit = queue_.front().limit;
            for (auto it = queue_.begin(); it != queue_.end() && batch.size() < max_batch_;) {
                if (it->limit == limit) {
                    batch.push_back(std::move(*it));
                    it = queue_.erase(it);
                } else {
                    ++it;
                }
            }
        }
        std::vector<std::vector<float>> qs;
        qs.reserve(batch.size());
        for (auto &r : batch) qs.push_back(std::move(r.q));
        std::unique_ptr<Pending> next;
        try {
            const uint64_t ticket = storage_.search_batch_submit(qs, batch[0].limit);   // staged + enqueued, returns at once
            next.reset(new Pending{ticket, std::move(batch)});
        } catch (...) {
            ++batches_;
            for (auto &r : batch) r.done.set_exception(std::current_exception());
        }
        if (pending) deliver(*pending);   // the older batch's answer (tickets are collected in issue order)
        pending = std::move(next);
    }
}

}  // namespace memex
