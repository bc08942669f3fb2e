/*
 * oracle/cosine_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the arithmetic memex's file-backed vector store performs
 * for one search, used as the parity checker for the CUDA path.  Nothing under
 * memex_b200/ may call, link or import this file; only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() do.
 *
 * What it follows
 *   - reference lib/libmemex/src/storage/local.rs:62-69   insert: ids are 1-based,
 *     next_id = len + 1.
 *   - reference lib/libmemex/src/storage/local.rs:71-91   search: neighbours come
 *     back ascending by distance, score = 1.0 - (1.0 / (1.0 / distance)) in f32.
 *   - hnsw_rs 0.1.20 (git jean-pierreBoth/hnswlib-rs rev 52a7f917, pinned in the
 *     reference's lib/libmemex/Cargo.toml:14; source NOT under /root/reference)
 *     `impl Distance<f32> for DistCosine`: per element the three products a*b,
 *     a*a, b*b are formed in f32, widened to f64 and folded left-to-right in f64;
 *     d = max(0, 1 - ab / sqrt(aa * bb)) cast to f32; if either norm is zero the
 *     distance is 0.  (Published algorithm restated from memory of that crate --
 *     it cannot be fetched offline.)
 *
 * The reference search itself is an HNSW walk (approximate).  The ranking this
 * oracle defines is the EXACT one that walk approximates: all rows ordered by
 * (distance ascending, id ascending).  oracle/hnsw_oracle.cpp restates the walk
 * for the timed CPU baseline and the recall report.
 *
 * Parity pinning: checked against the reference's only results fixture for this
 * path, local.rs:175-214 (`test_hnsw`: 3 vectors of dim 3, query [0.1,0.1,0.1],
 * first hit "test-two"), in tests/test_oracle.py.  The reference pins no score
 * value, and the crates cannot be built here, so score parity is "pinned to the
 * published formula", not to reference output.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* hnsw_rs DistCosine::eval */
float mxo_dist_cosine(const float *a, const float *b, size_t d)
{
    double ab = 0.0, aa = 0.0, bb = 0.0;
    for (size_t i = 0; i < d; ++i) {
        /* products are taken in f32 and only then widened */
        float pab = a[i] * b[i];
        float paa = a[i] * a[i];
        float pbb = b[i] * b[i];
        ab += (double)pab;
        aa += (double)paa;
        bb += (double)pbb;
    }
    if (aa > 0.0 && bb > 0.0) {
        double du = 1.0 - ab / sqrt(aa * bb);
        if (du < 0.0) du = 0.0; /* .max(0.) */
        return (float)du;
    }
    return 0.0f;
}

/* dot-product "distance" used by the store's dot metric: the score is the f64
 * left fold of the f32 products, cast to f32 (same arithmetic as `ab` above).
 * The reference has no dot metric; north_star asks for one. */
float mxo_dot(const float *a, const float *b, size_t d)
{
    double ab = 0.0;
    for (size_t i = 0; i < d; ++i) {
        float pab = a[i] * b[i];
        ab += (double)pab;
    }
    return (float)ab;
}

/* local.rs:86 -- similarity = 1.0 - (1.0 / (1.0 / distance)), all f32 */
float mxo_similarity(float distance)
{
    volatile float inv = 1.0f / distance;
    volatile float back = 1.0f / inv;
    return 1.0f - back;
}

typedef struct {
    float key; /* distance (cosine) or -dot (dot metric): smaller is better */
    uint64_t id;
} mxo_hit;

static int hit_less(const mxo_hit *x, const mxo_hit *y)
{
    if (x->key < y->key) return 1;
    if (x->key > y->key) return 0;
    return x->id < y->id;
}

/* keep the k best hits of a stream in `heap` (sorted ascending, insertion) */
static void hit_push(mxo_hit *best, uint32_t *cnt, uint32_t k, mxo_hit h)
{
    if (*cnt == k) {
        if (!hit_less(&h, &best[k - 1])) return;
        --*cnt;
    }
    uint32_t j = (*cnt)++;
    while (j > 0 && hit_less(&h, &best[j - 1])) {
        best[j] = best[j - 1];
        --j;
    }
    best[j] = h;
}

/*
 * Exact top-k of every query against the whole corpus.
 *   corpus [n, d] row-major f32, row r has id r + 1      (local.rs:63)
 *   metric 0 = cosine (score = mxo_similarity(d)), 1 = dot (score = dot)
 *   ids_out / scores_out [nq, k]; counts_out[q] = min(k, n)
 * Hits are ordered best first: (distance asc, id asc)   (local.rs:76-88)
 */
int mxo_exact_topk(const float *corpus, uint64_t n, uint32_t d, const float *queries,
                   uint32_t nq, uint32_t k, uint32_t metric, uint64_t *ids_out,
                   float *scores_out, uint32_t *counts_out)
{
    if (k == 0 || d == 0) return -1;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    for (uint32_t q = 0; q < nq; ++q) {
        const float *qv = queries + (size_t)q * d;
        mxo_hit *partial = (mxo_hit *)malloc(sizeof(mxo_hit) * (size_t)k * nthreads);
        uint32_t *pcnt = (uint32_t *)calloc(nthreads, sizeof(uint32_t));
        if (!partial || !pcnt) return -2;
#ifdef _OPENMP
#pragma omp parallel
#endif
        {
            int t = 0;
#ifdef _OPENMP
            t = omp_get_thread_num();
#endif
            mxo_hit *best = partial + (size_t)t * k;
            uint32_t cnt = 0;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
            for (int64_t r = 0; r < (int64_t)n; ++r) {
                mxo_hit h;
                /* DistCosine::eval(query, point): a = query, b = stored row */
                h.key = metric == 0 ? mxo_dist_cosine(qv, corpus + (size_t)r * d, d)
                                    : -mxo_dot(qv, corpus + (size_t)r * d, d);
                h.id = (uint64_t)r + 1;
                hit_push(best, &cnt, k, h);
            }
            pcnt[t] = cnt;
        }
        mxo_hit *best = (mxo_hit *)malloc(sizeof(mxo_hit) * k);
        uint32_t cnt = 0;
        for (int t = 0; t < nthreads; ++t)
            for (uint32_t j = 0; j < pcnt[t]; ++j) hit_push(best, &cnt, k, partial[(size_t)t * k + j]);
        for (uint32_t j = 0; j < k; ++j) {
            if (j < cnt) {
                ids_out[(size_t)q * k + j] = best[j].id;
                scores_out[(size_t)q * k + j] =
                    metric == 0 ? mxo_similarity(best[j].key) : -best[j].key;
            } else {
                ids_out[(size_t)q * k + j] = 0;
                scores_out[(size_t)q * k + j] = 0.0f;
            }
        }
        counts_out[q] = cnt;
        free(best);
        free(partial);
        free(pcnt);
    }
    return 0;
}

/* distances of chosen rows only (full-size spot checks: O(len) instead of O(n)) */
void mxo_scores_of(const float *corpus, uint32_t d, const float *query, const uint64_t *ids,
                   uint32_t len, uint32_t metric, float *scores_out)
{
    for (uint32_t j = 0; j < len; ++j) {
        const float *row = corpus + (size_t)(ids[j] - 1) * d;
        scores_out[j] = metric == 0 ? mxo_similarity(mxo_dist_cosine(query, row, d))
                                    : mxo_dot(query, row, d);
    }
}

int mxo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
