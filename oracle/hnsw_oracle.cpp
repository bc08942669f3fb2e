/*
 * oracle/hnsw_oracle.cpp -- TEST INFRASTRUCTURE ONLY (timed CPU baseline + recall report).
 *
 * CPU restatement of the search memex ships: an HNSW index over cosine distance with
 * the parameters the reference passes to hnsw_rs --
 *   Hnsw::new(max_nb_connection = 16, max_elements = 100, max_layer = 16,
 *             ef_construction = 200, DistCosine)      reference storage/local.rs:48,101
 *   hnsw.insert((&vec, id)) one point at a time       reference storage/local.rs:62-69
 *   hnsw.search(vec, limit, ef = 32)                  reference storage/local.rs:76
 *
 * hnsw_rs 0.1.20 @ 52a7f917 is a git dependency that is not vendored under
 * /root/reference and cannot be fetched offline, so this file restates the PUBLISHED
 * algorithm the crate implements (Malkov & Yashunin, "Efficient and robust approximate
 * nearest neighbor search using Hierarchical Navigable Small World graphs", 2018):
 * Alg. 1 INSERT, Alg. 2 SEARCH-LAYER, Alg. 4 SELECT-NEIGHBORS-HEURISTIC (keep pruned
 * connections, no candidate extension), Alg. 5 K-NN-SEARCH, level ~ floor(-ln U / ln M),
 * layer 0 holding up to 2M links.  The graph is randomised (in the crate too: its level
 * generator is seeded from entropy), so this is NOT a results oracle -- the results
 * oracle is the exact ranking in cosine_oracle.c, which HNSW approximates.  What this
 * file is for: timing the reference's algorithm on the GPU box's host cores and
 * reporting its recall against the exact ranking.
 *
 * Distance arithmetic is DistCosine's (see cosine_oracle.c).
 *
 * Building: memex inserts one point at a time (insert_one below).  For the bench sample the
 * build is not what is timed, so mxo_hnsw_insert_parallel inserts from several threads with a
 * lock per point's link lists and one for the entry point -- the scheme of the crate's own
 * `parallel_insert` -- which makes a 100 k-row sample affordable (2-4 ms per insert otherwise).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <vector>

extern "C" float mxo_dist_cosine(const float *a, const float *b, size_t d);

namespace {

struct Hnsw {
    uint32_t dim = 0;
    uint32_t M = 16, M0 = 32, max_layer = 16, efc = 200;
    std::vector<float> data;                                /* [n, dim] */
    std::vector<std::vector<std::vector<uint32_t>>> links;  /* [point][layer] -> neighbours */
    std::vector<uint32_t> level;
    int64_t entry = -1;
    uint32_t top_level = 0;
    std::mt19937_64 rng;
    /* visited stamps of the serial path */
    std::vector<uint32_t> stamp;
    uint32_t cur_stamp = 0;
    /* parallel build: one lock per point's link lists, one for (entry, top_level); null when serial */
    std::unique_ptr<std::mutex[]> node_mu;
    std::mutex entry_mu;

    const float *row(uint32_t i) const { return data.data() + (size_t)i * dim; }
    float dist(const float *q, uint32_t i) const { return mxo_dist_cosine(q, row(i), dim); }

    using Cand = std::pair<float, uint32_t>; /* (distance, point) */

    /* Alg. 2: returns up to ef closest to q found in `layer`, as a max-heap */
    std::priority_queue<Cand> search_layer(const float *q, const std::vector<Cand> &eps, uint32_t ef,
                                           uint32_t layer, std::vector<uint32_t> &st, uint32_t &cs)
    {
        if (++cs == 0) {
            std::fill(st.begin(), st.end(), 0u);
            cs = 1;
        }
        std::priority_queue<Cand, std::vector<Cand>, std::greater<Cand>> cand; /* min-heap */
        std::priority_queue<Cand> best;                                        /* max-heap */
        for (const Cand &e : eps) {
            st[e.second] = cs;
            cand.push(e);
            best.push(e);
        }
        while (!cand.empty()) {
            Cand c = cand.top();
            cand.pop();
            if (c.first > best.top().first && best.size() >= ef) break;
            std::vector<uint32_t> nbs_copy;
            const std::vector<uint32_t> *nbs = &links[c.second][layer];
            if (node_mu) {   /* another thread may be re-linking this point */
                std::lock_guard<std::mutex> g(node_mu[c.second]);
                nbs_copy = *nbs;
                nbs = &nbs_copy;
            }
            for (uint32_t nb : *nbs) {
                if (st[nb] == cs) continue;
                st[nb] = cs;
                float d = dist(q, nb);
                if (best.size() < ef || d < best.top().first) {
                    cand.push({d, nb});
                    best.push({d, nb});
                    if (best.size() > ef) best.pop();
                }
            }
        }
        return best;
    }

    /* Alg. 4 with keepPrunedConnections = true, extendCandidates = false */
    std::vector<uint32_t> select_heuristic(std::vector<Cand> cands, uint32_t m)
    {
        std::sort(cands.begin(), cands.end());
        std::vector<uint32_t> out;
        std::vector<Cand> pruned;
        for (const Cand &c : cands) {
            if (out.size() >= m) break;
            bool keep = true;
            for (uint32_t r : out) {
                if (dist(row(c.second), r) < c.first) {
                    keep = false;
                    break;
                }
            }
            if (keep)
                out.push_back(c.second);
            else
                pruned.push_back(c);
        }
        for (const Cand &c : pruned) {
            if (out.size() >= m) break;
            out.push_back(c.second);
        }
        return out;
    }

    uint32_t draw_level()
    {
        std::uniform_real_distribution<double> U(0.0, 1.0);
        double u = U(rng);
        if (u <= 0.0) u = 1e-300;
        uint32_t l = (uint32_t)std::floor(-std::log(u) / std::log((double)M));
        return l >= max_layer ? max_layer - 1 : l;
    }

    /* Alg. 1 for point `id`, whose row, level and (empty) link lists already exist */
    void link_point(uint32_t id, std::vector<uint32_t> &st, uint32_t &cs)
    {
        const uint32_t l = level[id];
        int64_t ent;
        uint32_t top;
        {
            std::unique_lock<std::mutex> g(entry_mu, std::defer_lock);
            if (node_mu) g.lock();
            ent = entry;
            top = top_level;
            if (ent < 0) {
                entry = id;
                top_level = l;
                return;
            }
        }
        const float *q = row(id);
        std::vector<Cand> ep{{dist(q, (uint32_t)ent), (uint32_t)ent}};
        for (int64_t lc = top; lc > (int64_t)l; --lc) {
            auto best = search_layer(q, ep, 1, (uint32_t)lc, st, cs);
            while (best.size() > 1) best.pop();
            ep = {best.top()};
        }
        for (int64_t lc = std::min(top, l); lc >= 0; --lc) {
            auto best = search_layer(q, ep, efc, (uint32_t)lc, st, cs);
            std::vector<Cand> w;
            w.reserve(best.size());
            while (!best.empty()) {
                w.push_back(best.top());
                best.pop();
            }
            uint32_t mmax = lc == 0 ? M0 : M;
            std::vector<uint32_t> nbrs = select_heuristic(w, M);
            {
                std::unique_lock<std::mutex> g;
                if (node_mu) g = std::unique_lock<std::mutex>(node_mu[id]);
                links[id][lc] = nbrs;
            }
            for (uint32_t nb : nbrs) {
                std::unique_lock<std::mutex> g;
                if (node_mu) g = std::unique_lock<std::mutex>(node_mu[nb]);
                auto &ln = links[nb][lc];
                ln.push_back(id);
                if (ln.size() > mmax) {
                    std::vector<Cand> c;
                    c.reserve(ln.size());
                    for (uint32_t x : ln) c.push_back({dist(row(nb), x), x});
                    ln = select_heuristic(c, mmax);
                }
            }
            ep = w;
        }
        if (l > top) {
            std::unique_lock<std::mutex> g(entry_mu, std::defer_lock);
            if (node_mu) g.lock();
            if (l > top_level) {
                top_level = l;
                entry = id;
            }
        }
    }

    /* rows, levels and empty link lists for n more points; returns the id of the first */
    uint32_t append_rows(const float *v, uint64_t n)
    {
        const uint32_t first = (uint32_t)level.size();
        data.insert(data.end(), v, v + (size_t)n * dim);
        for (uint64_t i = 0; i < n; ++i) {
            const uint32_t l = draw_level();
            level.push_back(l);
            links.emplace_back(l + 1);
        }
        stamp.resize(level.size(), 0u);
        return first;
    }

    void insert_one(const float *v)
    {
        const uint32_t id = append_rows(v, 1);
        link_point(id, stamp, cur_stamp);
    }

    /* Alg. 5; thread-safe for concurrent readers given caller-owned stamps */
    uint32_t search(const float *q, uint32_t k, uint32_t ef, uint64_t *ids, float *dists,
                    std::vector<uint32_t> &st, uint32_t &cs)
    {
        if (entry < 0) return 0;
        std::vector<Cand> ep{{mxo_dist_cosine(q, row((uint32_t)entry), dim), (uint32_t)entry}};
        for (int64_t lc = top_level; lc > 0; --lc) {
            auto best = search_layer(q, ep, 1, (uint32_t)lc, st, cs);
            while (best.size() > 1) best.pop();
            ep = {best.top()};
        }
        auto best = search_layer(q, ep, std::max(ef, k), 0, st, cs);
        while (best.size() > k) best.pop();
        uint32_t n = (uint32_t)best.size();
        for (uint32_t j = n; j-- > 0;) {
            ids[j] = (uint64_t)best.top().second + 1; /* 1-based d_id, local.rs:63 */
            dists[j] = best.top().first;
            best.pop();
        }
        return n;
    }
};

}  // namespace

extern "C" {

void *mxo_hnsw_new(uint32_t dim, uint64_t seed)
{
    Hnsw *h = new Hnsw();
    h->dim = dim;
    h->rng.seed(seed);
    return h;
}

void mxo_hnsw_free(void *p) { delete (Hnsw *)p; }

void mxo_hnsw_insert(void *p, const float *vecs, uint64_t n)
{
    Hnsw *h = (Hnsw *)p;
    for (uint64_t i = 0; i < n; ++i) h->insert_one(vecs + (size_t)i * h->dim);
}

/* same graph family, built by `threads` OpenMP threads (not the timed part of any measurement) */
void mxo_hnsw_insert_parallel(void *p, const float *vecs, uint64_t n, int threads)
{
    Hnsw *h = (Hnsw *)p;
    if (threads <= 1 || n < 2048) {
        for (uint64_t i = 0; i < n; ++i) h->insert_one(vecs + (size_t)i * h->dim);
        return;
    }
    /* a serial prefix gives the upper layers some shape before the threads start */
    const uint64_t serial = 1024;
    for (uint64_t i = 0; i < serial; ++i) h->insert_one(vecs + (size_t)i * h->dim);
    const uint32_t first = h->append_rows(vecs + (size_t)serial * h->dim, n - serial);
    const size_t total = h->level.size();
    h->node_mu.reset(new std::mutex[total]);
#pragma omp parallel num_threads(threads)
    {
        std::vector<uint32_t> st(total, 0u);
        uint32_t cs = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < (int64_t)(n - serial); ++i) h->link_point(first + (uint32_t)i, st, cs);
    }
    h->node_mu.reset();
}

uint64_t mxo_hnsw_len(void *p) { return ((Hnsw *)p)->level.size(); }

/* nq queries; `threads` > 1 runs independent queries on OpenMP threads (the reference
 * serialises searches behind a tokio Mutex, storage/mod.rs:85-92; threads = 1 is "as shipped"). */
void mxo_hnsw_search(void *p, const float *queries, uint32_t nq, uint32_t k, uint32_t ef,
                     int threads, uint64_t *ids_out, float *dists_out, uint32_t *counts_out)
{
    Hnsw *h = (Hnsw *)p;
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
    {
        std::vector<uint32_t> st(h->level.size(), 0u);
        uint32_t cs = 0;
#pragma omp for schedule(dynamic, 4)
        for (int64_t q = 0; q < (int64_t)nq; ++q) {
            counts_out[q] = h->search(queries + (size_t)q * h->dim, k, ef, ids_out + (size_t)q * k,
                                      dists_out + (size_t)q * k, st, cs);
            for (uint32_t j = counts_out[q]; j < k; ++j) {
                ids_out[(size_t)q * k + j] = 0;
                dists_out[(size_t)q * k + j] = 0.f;
            }
        }
    }
}

}  // extern "C"
