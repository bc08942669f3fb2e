// shard_group.cu -- mx_shard_group: the multi-GPU search of a row-sharded corpus BELOW the C ABI.
//
// The reference answers `VectorStorage::search` from one index (reference lib/libmemex/src/storage/mod.rs:85-92 ->
// storage/local.rs:71-91).  Here the rows are dealt over G GPUs (one mx_store each, GLOBAL ids through
// mx_store_cfg.id_offset / id_stride) and one search is: every GPU scans its shard, ONE exchange step moves the G
// per-shard top-k blobs, every GPU merges G * k candidates per query -- (key asc, id asc), the single-store ordering.
//
// A group member owns ONE exchange buffer in its GPU's memory that every peer can address:
//
//     [2][world] blob slots (stride bytes each) | [2] query slots ([max_nq, dim] f32) | flags: world x u32 (blob
//     epochs) + 1 x u32 (query epoch), zero-initialised
//
// and the rendezvous comes in two forms, both without torch and without a collective library:
//   * one process per GPU (the launch contract of bench.py / torchrun; a Rust or C++ worker per GPU):
//     mx_shard_group_export writes the buffer's 64-byte CUDA IPC handle, the host side moves the `world` handles over
//     whatever transport it has (a file, a socket, MPI, torch.distributed as plumbing), mx_shard_group_connect opens them;
//   * one process driving several GPUs (memex's server is ONE process, storage/mod.rs:68-93):
//     mx_shard_group_connect_local enables peer access between the members' devices and wires the raw pointers.
// A search is then three of this library's launches after the shard's own scan + rerank: the optional query push from
// the root (so that the host-buffer call needs no broadcast), the blob push to every peer (P2P stores over NVLink +
// release-stored epoch flags), and the merge that waits for its `world` flags inside the kernel.
// Slot reuse is safe with two parities: a peer can only push epoch e + 2 after its own merge of e + 1, which waited for
// this rank's push of e + 1, which is stream-ordered after this rank's merge (and scan) of epoch e.
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "scan.cuh"

namespace mx {
constexpr uint32_t kGroupMagic = 0x4d584752;   // "MXGR"

__global__ void wait_flag_kernel(const uint32_t *flag, uint32_t epoch)
{
    if (threadIdx.x != 0) return;
    uint32_t seen, spins = 0;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if (++spins > (1u << 27)) {
            printf("memex_b200: wait for the root's queries timed out (flag %u, epoch %u)\n", seen, epoch);
            __trap();
        }
    } while ((int32_t)(seen - epoch) < 0);
}
}  // namespace mx

using namespace mx;

struct mx_shard_group : HandleBase {
    int32_t device = 0;
    uint32_t world = 1, rank = 0, dim = 0, max_nq = 0, max_k = 0;
    uint64_t stride = 0;        // bytes of one blob slot
    uint64_t q_bytes = 0;       // bytes of one query slot
    uint64_t off_q = 0, off_flags = 0, total = 0;
    char *buf = nullptr;        // this member's exchange buffer
    void *mine = nullptr;       // this member's blob, before the push
    uint64_t peer[kMaxPeers] = {};
    bool opened[kMaxPeers] = {};   // peer[i] came from cudaIpcOpenMemHandle
    bool connected = false;
    uint32_t epoch = 0;
    cudaStream_t stream = nullptr;   // used when the caller passes no stream
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    void *dev_io = nullptr;     // host-buffer call: queries in, answer out
    size_t dev_io_cap = 0;
    IoSlot slots[2];            // mx_shard_group_search_submit / _collect: two host-buffer searches in flight
    uint64_t submit_seq = 0;
};

namespace {

int32_t group_set_device(mx_shard_group *g) { MX_CUDA(g, MX_ERR_CONNECTION, cudaSetDevice(g->device)); return MX_OK; }

}  // namespace

extern "C" {

int32_t mx_shard_group_create(int32_t device, uint32_t world, uint32_t rank, uint32_t dim, uint32_t max_nq, uint32_t max_k,
                              mx_shard_group **out)
{
    if (!out) return fail(nullptr, MX_ERR_INVALID, "null argument");
    *out = nullptr;
    if (world == 0 || world > (uint32_t)kMaxPeers || rank >= world)
        return fail(nullptr, MX_ERR_INVALID, "world must be in [1, %d] and rank below it", kMaxPeers);
    if (dim == 0 || max_nq == 0 || max_k == 0 || max_k > MX_MAX_K) return fail(nullptr, MX_ERR_INVALID, "bad group shape");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, MX_ERR_CONNECTION, "no CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, MX_ERR_CONNECTION, "device %d out of range [0, %d)", device, ndev);
    mx_shard_group *g = new mx_shard_group();
    g->magic = kGroupMagic;
    g->device = device;
    g->world = world;
    g->rank = rank;
    g->dim = dim;
    g->max_nq = max_nq;
    g->max_k = max_k;
    g->stride = (mx_topk_blob_bytes(max_nq, max_k) + 255) & ~255ull;
    g->q_bytes = (((uint64_t)max_nq * dim * sizeof(float)) + 255) & ~255ull;
    g->off_q = 2ull * world * g->stride;
    g->off_flags = g->off_q + 2 * g->q_bytes;
    g->total = g->off_flags + 256;
    auto bail = [&](const char *what, cudaError_t ce) {
        int32_t r = fail(nullptr, MX_ERR_CONNECTION, "%s: %s", what, cudaGetErrorString(ce));
        mx_shard_group_destroy(g);
        return r;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaMalloc(&g->buf, g->total)) != cudaSuccess) return bail("cudaMalloc(exchange buffer)", e);
    if ((e = cudaMalloc(&g->mine, g->stride)) != cudaSuccess) return bail("cudaMalloc(blob)", e);
    if ((e = cudaMemset(g->buf, 0, g->total)) != cudaSuccess) return bail("cudaMemset", e);
    if ((e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bail("cudaDeviceSynchronize", e);
    g->peer[rank] = reinterpret_cast<uint64_t>(g->buf);
    if (world == 1) g->connected = true;
    *out = g;
    return MX_OK;
}

void mx_shard_group_destroy(mx_shard_group *g)
{
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    for (uint32_t i = 0; i < g->world; ++i)
        if (g->opened[i]) cudaIpcCloseMemHandle(reinterpret_cast<void *>(g->peer[i]));
    cudaFree(g->buf);
    cudaFree(g->mine);
    cudaFree(g->dev_io);
    if (g->pinned) cudaFreeHost(g->pinned);
    for (IoSlot &sl : g->slots) sl.release();
    if (g->stream) cudaStreamDestroy(g->stream);
    g->magic = 0;
    delete g;
}

int32_t mx_shard_group_export(mx_shard_group *g, void *handle_out)
{
    if (!g || !handle_out) return MX_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == MX_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    int32_t rc;
    if ((rc = group_set_device(g)) != MX_OK) return rc;
    cudaIpcMemHandle_t h;
    MX_CUDA(g, MX_ERR_CONNECTION, cudaIpcGetMemHandle(&h, g->buf));
    memcpy(handle_out, &h, sizeof h);
    return MX_OK;
}

int32_t mx_shard_group_connect(mx_shard_group *g, const void *handles)
{
    if (!g || !handles) return MX_ERR_INVALID;
    if (g->connected) return fail(g, MX_ERR_INVALID, "group is already connected");
    int32_t rc;
    if ((rc = group_set_device(g)) != MX_OK) return rc;
    const char *hb = static_cast<const char *>(handles);
    for (uint32_t i = 0; i < g->world; ++i) {
        if (i == g->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hb + (size_t)i * MX_IPC_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(g, MX_ERR_CONNECTION, "cudaIpcOpenMemHandle(rank %u) failed: %s", i, cudaGetErrorString(e));
        }
        g->peer[i] = reinterpret_cast<uint64_t>(p);
        g->opened[i] = true;
    }
    g->connected = true;
    return MX_OK;
}

int32_t mx_shard_group_connect_local(mx_shard_group *const *groups, uint32_t world)
{
    if (!groups || world == 0 || world > (uint32_t)kMaxPeers) return fail(nullptr, MX_ERR_INVALID, "bad group list");
    for (uint32_t i = 0; i < world; ++i)
        if (!groups[i] || groups[i]->world != world || groups[i]->rank != i)
            return fail(nullptr, MX_ERR_INVALID, "member %u does not belong to a group of %u in rank order", i, world);
    for (uint32_t i = 0; i < world; ++i) {
        mx_shard_group *g = groups[i];
        MX_CUDA(g, MX_ERR_CONNECTION, cudaSetDevice(g->device));
        for (uint32_t j = 0; j < world; ++j) {
            if (j != i && groups[j]->device != g->device) {
                int can = 0;
                MX_CUDA(g, MX_ERR_CONNECTION, cudaDeviceCanAccessPeer(&can, g->device, groups[j]->device));
                if (!can) return fail(g, MX_ERR_CONNECTION, "device %d cannot address device %d", g->device, groups[j]->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(groups[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(g, MX_ERR_CONNECTION, "cudaDeviceEnablePeerAccess(%d -> %d): %s", g->device, groups[j]->device,
                                cudaGetErrorString(e));
                cudaGetLastError();
            }
            g->peer[j] = reinterpret_cast<uint64_t>(groups[j]->buf);
        }
        g->connected = true;
    }
    return MX_OK;
}

// phase 1 of a member's search: (queries through peer memory,) the shard's own scan + rerank into the member's blob, the
// push of that blob into every peer's slot.  Nothing in it waits for another member.
static int32_t group_scan_and_push(mx_shard_group *g, mx_store *s, const float *queries_dev, int32_t query_root, uint32_t nq,
                                   uint32_t k, cudaStream_t st, uint32_t *metric_out)
{
    if (!g->connected) return fail(g, MX_ERR_CONNECTION, "group is not connected (mx_shard_group_connect[_local] first)");
    if (nq > g->max_nq || k == 0 || k > g->max_k) return fail(g, MX_ERR_INVALID, "batch of %u x top-%u exceeds the group's %u x %u", nq, k, g->max_nq, g->max_k);
    if (query_root >= (int32_t)g->world) return fail(g, MX_ERR_INVALID, "query root %d outside the group", query_root);
    uint32_t sdim = 0, metric = 0;
    mx_store_info(s, &sdim, nullptr, &metric, nullptr);
    if (sdim != g->dim) return fail(g, MX_ERR_INVALID, "store has dimension %u, group %u", sdim, g->dim);
    *metric_out = metric;
    int32_t rc;
    if ((rc = group_set_device(g)) != MX_OK) return rc;
    const uint32_t epoch = ++g->epoch;
    const uint64_t half = (uint64_t)(epoch & 1u);
    const float *q_use = queries_dev;
    if (g->world > 1 && query_root >= 0) {
        // the root's query block through peer memory: no collective, no host round trip on the other ranks
        const uint64_t q_off = g->off_q + half * g->q_bytes;
        const uint64_t qb = (((uint64_t)nq * g->dim * sizeof(float)) + 15) & ~15ull;
        if ((uint32_t)query_root == g->rank) {
            if (!queries_dev) return fail(g, MX_ERR_INVALID, "the query root passes the queries");
            PushParams pp{};
            pp.blob = queries_dev;
            pp.blob_bytes = qb;
            for (uint32_t i = 0; i < g->world; ++i) pp.peer_base[i] = g->peer[i];
            pp.slot_offset = q_off;
            pp.flag_offset = g->off_flags;
            pp.world = g->world;
            pp.rank = g->world;            // flag index `world` = the query epoch
            pp.epoch = epoch;
            MX_CUDA(g, MX_ERR_SEARCH, launch_exchange_push(pp, st));
        } else {
            wait_flag_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const uint32_t *>(g->buf + g->off_flags) + g->world, epoch);
            count_launch();
            MX_CUDA(g, MX_ERR_SEARCH, cudaGetLastError());
            q_use = reinterpret_cast<const float *>(g->buf + q_off);
        }
    }
    if (!q_use) return fail(g, MX_ERR_INVALID, "null queries");
    rc = mx_store_search_blob_device(s, q_use, nq, k, g->mine, st);
    if (rc != MX_OK) return fail(g, rc, "shard search failed: %s", mx_last_error(s));
    if (g->world == 1) return MX_OK;
    PushParams pp{};
    pp.blob = g->mine;
    pp.blob_bytes = mx_topk_blob_bytes(nq, k);
    for (uint32_t i = 0; i < g->world; ++i) pp.peer_base[i] = g->peer[i];
    pp.slot_offset = (half * g->world + g->rank) * g->stride;
    pp.flag_offset = g->off_flags;
    pp.world = g->world;
    pp.rank = g->rank;
    pp.epoch = epoch;
    MX_CUDA(g, MX_ERR_SEARCH, launch_exchange_push(pp, st));
    return MX_OK;
}

// phase 2: the merge of the `world` blobs of the current epoch; its kernel waits for the peers' flags
static int32_t group_merge(mx_shard_group *g, uint32_t metric, uint32_t nq, uint32_t k, uint64_t *ids_dev, float *scores_dev,
                           uint32_t *counts_dev, cudaStream_t st)
{
    int32_t rc;
    if ((rc = group_set_device(g)) != MX_OK) return rc;
    if (g->world == 1)
        return mx_merge_topk_blobs_device(g->mine, mx_topk_blob_bytes(nq, k), 1, nq, k, metric, ids_dev, scores_dev, counts_dev,
                                          g->device, st);
    const uint64_t half = (uint64_t)(g->epoch & 1u);
    rc = mx_merge_topk_blobs_wait_device(g->buf + half * g->world * g->stride, g->stride, g->world, nq, k, metric, ids_dev,
                                         scores_dev, counts_dev, reinterpret_cast<const uint32_t *>(g->buf + g->off_flags),
                                         g->epoch, g->device, st);
    if (rc != MX_OK) return fail(g, rc, "merge failed: %s", mx_last_error(nullptr));
    return MX_OK;
}

int32_t mx_shard_group_search_device(mx_shard_group *g, mx_store *s, const float *queries_dev, int32_t query_root, uint32_t nq,
                                     uint32_t k, uint64_t *ids_dev, float *scores_dev, uint32_t *counts_dev, void *cuda_stream)
{
    if (!g || !s) return MX_ERR_INVALID;
    if (nq == 0) return MX_OK;
    if (!ids_dev || !scores_dev || !counts_dev) return fail(g, MX_ERR_INVALID, "null buffer");
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : g->stream;
    uint32_t metric = 0;
    int32_t rc = group_scan_and_push(g, s, queries_dev, query_root, nq, k, st, &metric);
    if (rc != MX_OK) return rc;
    return group_merge(g, metric, nq, k, ids_dev, scores_dev, counts_dev, st);
}

int32_t mx_shard_group_search(mx_shard_group *g, mx_store *s, const float *queries, int32_t query_root, uint32_t nq, uint32_t k,
                              uint64_t *ids_out, float *scores_out, uint32_t *counts_out)
{
    if (!g || !s) return MX_ERR_INVALID;
    if (!ids_out || !scores_out || !counts_out) return fail(g, MX_ERR_INVALID, "null buffer");
    if (nq == 0) return MX_OK;
    if (nq > g->max_nq || k == 0 || k > g->max_k) return fail(g, MX_ERR_INVALID, "batch of %u x top-%u exceeds the group's %u x %u", nq, k, g->max_nq, g->max_k);
    const bool have_q = query_root < 0 || (uint32_t)query_root == g->rank;
    if (have_q && !queries) return fail(g, MX_ERR_INVALID, "null queries");
    int32_t rc;
    if ((rc = group_set_device(g)) != MX_OK) return rc;
    const size_t qb = (size_t)nq * g->dim * sizeof(float);
    const size_t ib = (size_t)nq * k * sizeof(uint64_t), sb = (size_t)nq * k * sizeof(float), cb = (size_t)nq * 4;
    const size_t off_i = (qb + 255) & ~(size_t)255, off_s = off_i + ib, off_c = off_s + sb, total = off_c + cb;
    if (total > g->pinned_cap) {
        if (g->pinned) cudaFreeHost(g->pinned);
        cudaFree(g->dev_io);
        g->pinned = nullptr;
        g->dev_io = nullptr;
        g->pinned_cap = g->dev_io_cap = 0;
        MX_CUDA(g, MX_ERR_CONNECTION, cudaMallocHost(&g->pinned, total));
        MX_CUDA(g, MX_ERR_CONNECTION, cudaMalloc(&g->dev_io, total));
        g->pinned_cap = g->dev_io_cap = total;
    }
    char *hp = static_cast<char *>(g->pinned), *dp = static_cast<char *>(g->dev_io);
    if (have_q) {
        const int64_t bad_row = copy_checking_finite(reinterpret_cast<float *>(hp), queries, nq, g->dim);
        if (bad_row >= 0) return fail(g, MX_ERR_SEARCH, "non-finite value in query %lld", (long long)bad_row);
        MX_CUDA(g, MX_ERR_SEARCH, cudaMemcpyAsync(dp, hp, qb, cudaMemcpyHostToDevice, g->stream));
    }
    // the merge kernel stores the answer straight into the pinned host buffer (posted PCIe writes, never read back by the
    // device): no device-to-host copy operation at the end of the step
    rc = mx_shard_group_search_device(g, s, have_q ? reinterpret_cast<const float *>(dp) : nullptr, query_root, nq, k,
                                      reinterpret_cast<uint64_t *>(hp + off_i), reinterpret_cast<float *>(hp + off_s),
                                      reinterpret_cast<uint32_t *>(hp + off_c), g->stream);
    if (rc != MX_OK) return rc;
    MX_CUDA(g, MX_ERR_SEARCH, cudaStreamSynchronize(g->stream));
    memcpy(ids_out, hp + off_i, ib);
    memcpy(scores_out, hp + off_s, sb);
    memcpy(counts_out, hp + off_c, cb);
    return MX_OK;
}

// The same step split in two so that TWO may be in flight per member (every member submits and collects in the same order):
// while the device works on search i the host stages, copies in and enqueues search i + 1, and the GPUs never wait for a
// host.  The exchange buffers already allow it -- slots and flags alternate between two parities, and a member's push of
// epoch e + 2 is stream-ordered after its merge of e + 1 -- it is what the device-buffer call does when it is called in a loop.
int32_t mx_shard_group_search_submit(mx_shard_group *g, mx_store *s, const float *queries, int32_t query_root, uint32_t nq,
                                     uint32_t k, uint64_t *ticket_out)
{
    if (!g || !s) return MX_ERR_INVALID;
    if (!ticket_out) return fail(g, MX_ERR_INVALID, "null buffer");
    if (nq == 0 || nq > g->max_nq || k == 0 || k > g->max_k)
        return fail(g, MX_ERR_INVALID, "batch of %u x top-%u outside the group's %u x %u", nq, k, g->max_nq, g->max_k);
    const bool have_q = query_root < 0 || (uint32_t)query_root == g->rank;
    if (have_q && !queries) return fail(g, MX_ERR_INVALID, "null queries");
    IoSlot &sl = g->slots[g->submit_seq & 1];
    if (sl.busy)
        return fail(g, MX_ERR_INVALID, "two searches are already in flight: collect ticket %llu first", (unsigned long long)sl.ticket);
    int32_t rc;
    if ((rc = group_set_device(g)) != MX_OK) return rc;
    const size_t qb = (size_t)nq * g->dim * sizeof(float);
    sl.layout(qb, nq, k);
    sl.empty = false;
    MX_CUDA(g, MX_ERR_CONNECTION, sl.reserve(sl.total()));
    char *hp = static_cast<char *>(sl.pinned), *dp = static_cast<char *>(sl.dev);
    if (have_q) {
        const int64_t bad_row = copy_checking_finite(reinterpret_cast<float *>(hp), queries, nq, g->dim);
        if (bad_row >= 0) return fail(g, MX_ERR_SEARCH, "non-finite value in query %lld", (long long)bad_row);
        MX_CUDA(g, MX_ERR_SEARCH, cudaMemcpyAsync(dp, hp, qb, cudaMemcpyHostToDevice, g->stream));
    }
    rc = mx_shard_group_search_device(g, s, have_q ? reinterpret_cast<const float *>(dp) : nullptr, query_root, nq, k,
                                      reinterpret_cast<uint64_t *>(hp + sl.off_i), reinterpret_cast<float *>(hp + sl.off_s),
                                      reinterpret_cast<uint32_t *>(hp + sl.off_c), g->stream);
    if (rc != MX_OK) return rc;
    MX_CUDA(g, MX_ERR_SEARCH, cudaEventRecord(sl.done, g->stream));
    sl.busy = true;
    sl.ticket = g->submit_seq;
    *ticket_out = g->submit_seq++;
    return MX_OK;
}

int32_t mx_shard_group_search_collect(mx_shard_group *g, uint64_t ticket, uint64_t *ids_out, float *scores_out, uint32_t *counts_out)
{
    if (!g) return MX_ERR_INVALID;
    if (!ids_out || !scores_out || !counts_out) return fail(g, MX_ERR_INVALID, "null buffer");
    IoSlot &sl = g->slots[ticket & 1];
    if (!sl.busy || sl.ticket != ticket) return fail(g, MX_ERR_INVALID, "no search with ticket %llu is in flight", (unsigned long long)ticket);
    sl.busy = false;
    int32_t rc;
    if ((rc = group_set_device(g)) != MX_OK) return rc;
    MX_CUDA(g, MX_ERR_SEARCH, cudaEventSynchronize(sl.done));
    sl.copy_out(ids_out, scores_out, counts_out);
    return MX_OK;
}

// One process, `world` members on `world` devices: H2D of the query block on every member, every member's search
// enqueued (asynchronously: member 0's merge waits inside its kernel for pushes that are enqueued after it), then ONE
// D2H of member 0's answer.  This is the call behind a single-process host's VectorStore::search.
int32_t mx_shard_group_search_local(mx_shard_group *const *groups, mx_store *const *stores, uint32_t world, const float *queries,
                                    uint32_t nq, uint32_t k, uint64_t *ids_out, float *scores_out, uint32_t *counts_out)
{
    if (!groups || !stores || world == 0 || world > (uint32_t)kMaxPeers) return fail(nullptr, MX_ERR_INVALID, "bad group list");
    for (uint32_t i = 0; i < world; ++i)
        if (!groups[i] || !stores[i] || groups[i]->world != world || groups[i]->rank != i)
            return fail(nullptr, MX_ERR_INVALID, "member %u does not belong to a group of %u in rank order", i, world);
    mx_shard_group *g0 = groups[0];
    if (!queries || !ids_out || !scores_out || !counts_out) return fail(g0, MX_ERR_INVALID, "null buffer");
    if (nq == 0) return MX_OK;
    if (nq > g0->max_nq || k == 0 || k > g0->max_k) return fail(g0, MX_ERR_INVALID, "batch of %u x top-%u exceeds the group's %u x %u", nq, k, g0->max_nq, g0->max_k);
    for (size_t r = 0; r < nq; ++r) {
        uint32_t bad = 0;
        for (size_t i = 0; i < g0->dim; ++i) {
            uint32_t b;
            memcpy(&b, queries + r * g0->dim + i, 4);
            bad |= ((b & 0x7f800000u) == 0x7f800000u) ? 1u : 0u;
        }
        if (bad) return fail(g0, MX_ERR_SEARCH, "non-finite value in query %zu", r);
    }
    const size_t qb = (size_t)nq * g0->dim * sizeof(float);
    const size_t ib = (size_t)nq * k * sizeof(uint64_t), sb = (size_t)nq * k * sizeof(float), cb = (size_t)nq * 4;
    const size_t off_i = (qb + 255) & ~(size_t)255, off_s = off_i + ib, off_c = off_s + sb, total = off_c + cb;
    // buffers first: an allocation synchronises its device, which must not happen while a member's merge is waiting
    for (uint32_t i = 0; i < world; ++i) {
        mx_shard_group *g = groups[i];
        MX_CUDA(g0, MX_ERR_CONNECTION, cudaSetDevice(g->device));
        if (total > g->pinned_cap) {
            if (g->pinned) cudaFreeHost(g->pinned);
            cudaFree(g->dev_io);
            g->pinned = nullptr;
            g->dev_io = nullptr;
            g->pinned_cap = g->dev_io_cap = 0;
            MX_CUDA(g0, MX_ERR_CONNECTION, cudaMallocHost(&g->pinned, total));
            MX_CUDA(g0, MX_ERR_CONNECTION, cudaMalloc(&g->dev_io, total));
            g->pinned_cap = g->dev_io_cap = total;
        }
    }
    // phase 1 on every member (scan + rerank + push: waits for nobody), THEN phase 2 on every member (the merges wait for
    // the pushes inside their kernels): whatever the order in which the devices run, nothing waits for work that has not
    // been enqueued -- two members may even share a device (tests on a 1-GPU box)
    uint32_t metric = 0;
    for (uint32_t i = 0; i < world; ++i) {
        mx_shard_group *g = groups[i];
        MX_CUDA(g0, MX_ERR_CONNECTION, cudaSetDevice(g->device));
        char *hp = static_cast<char *>(g->pinned), *dp = static_cast<char *>(g->dev_io);
        memcpy(hp, queries, qb);
        MX_CUDA(g0, MX_ERR_SEARCH, cudaMemcpyAsync(dp, hp, qb, cudaMemcpyHostToDevice, g->stream));
        int32_t rc = group_scan_and_push(g, stores[i], reinterpret_cast<const float *>(dp), -1, nq, k, g->stream, &metric);
        if (rc != MX_OK) {
            if (g != g0) g0->last_error = g->last_error;
            return rc;
        }
    }
    for (uint32_t i = 0; i < world; ++i) {
        mx_shard_group *g = groups[i];
        char *dp = static_cast<char *>(g->dev_io);
        int32_t rc = group_merge(g, metric, nq, k, reinterpret_cast<uint64_t *>(dp + off_i), reinterpret_cast<float *>(dp + off_s),
                                 reinterpret_cast<uint32_t *>(dp + off_c), g->stream);
        if (rc != MX_OK) {
            if (g != g0) g0->last_error = g->last_error;
            return rc;
        }
    }
    MX_CUDA(g0, MX_ERR_CONNECTION, cudaSetDevice(g0->device));
    char *hp = static_cast<char *>(g0->pinned), *dp = static_cast<char *>(g0->dev_io);
    MX_CUDA(g0, MX_ERR_SEARCH, cudaMemcpyAsync(hp + off_i, dp + off_i, total - off_i, cudaMemcpyDeviceToHost, g0->stream));
    // every member's stream drains (their merges ran too: the next call may reuse the slots at once)
    for (uint32_t i = 0; i < world; ++i) {
        MX_CUDA(g0, MX_ERR_CONNECTION, cudaSetDevice(groups[i]->device));
        MX_CUDA(g0, MX_ERR_SEARCH, cudaStreamSynchronize(groups[i]->stream));
    }
    memcpy(ids_out, hp + off_i, ib);
    memcpy(scores_out, hp + off_s, sb);
    memcpy(counts_out, hp + off_c, cb);
    return MX_OK;
}

int32_t mx_shard_group_info(const mx_shard_group *g, uint32_t *world, uint32_t *rank, uint32_t *epoch, int32_t *connected)
{
    if (!g) return MX_ERR_INVALID;
    if (world) *world = g->world;
    if (rank) *rank = g->rank;
    if (epoch) *epoch = g->epoch;
    if (connected) *connected = g->connected ? 1 : 0;
    return MX_OK;
}

}  // extern "C"
