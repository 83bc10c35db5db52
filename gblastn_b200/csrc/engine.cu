// engine.cu — host orchestration of the blastn hot path and the C ABI of include/gblastn_b200.h.
//
// One search = per (resident volume, query batch):
//   chunk table -> scan kernel -> radix sorts (emission order, then diagonal group) -> diagonal /
//   ungapped kernel -> speculative gapped kernel -> D2H -> host replay (hostpost.cpp).
// Everything runs on one CUDA stream per device; the only host<->device round trips are the
// survivor / init-hit counters (needed to size the sorts) and the final D2H of init-HSPs.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>


#include "bn_device.cuh"
#include "devmem.h"
#include "dbfile.h"
#include "hostpost.h"

namespace bn {

// ---- devmem.h ------------------------------------------------------------------------------------
static thread_local DevArena *g_arena_tls = nullptr;
DevArena *&current_arena() { return g_arena_tls; }
static std::mutex g_arena_mu;
static std::vector<std::pair<uintptr_t, uintptr_t>> g_arena_ranges;
void arena_register(const DevArena &a)
{
    if (!a.base) return;
    std::lock_guard<std::mutex> lk(g_arena_mu);
    g_arena_ranges.emplace_back((uintptr_t)a.base, (uintptr_t)a.base + a.cap);
}
void arena_unregister(const DevArena &a)
{
    std::lock_guard<std::mutex> lk(g_arena_mu);
    for (size_t i = 0; i < g_arena_ranges.size(); i++)
        if (g_arena_ranges[i].first == (uintptr_t)a.base) { g_arena_ranges.erase(g_arena_ranges.begin() + (long)i); break; }
}
static bool in_arena(const void *p)
{
    std::lock_guard<std::mutex> lk(g_arena_mu);
    for (const auto &r : g_arena_ranges) if ((uintptr_t)p >= r.first && (uintptr_t)p < r.second) return true;
    return false;
}
cudaError_t dev_malloc(void **p, size_t bytes, cudaStream_t st)
{
    DevArena *a = g_arena_tls;
    if (a) {
        const size_t at = (a->used + 255) & ~size_t(255);
        a->used = at + bytes;
        if (a->base && at + bytes <= a->cap) { *p = a->base + at; return cudaSuccess; }
    }
    return cudaMallocAsync(p, bytes, st);
}
cudaError_t dev_free(void *p, cudaStream_t st)
{
    if (!p || in_arena(p)) return cudaSuccess;
    return cudaFreeAsync(p, st);
}

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CU_TRY(expr)                                                                         \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(e__ == cudaErrorMemoryAllocation ? BN_ERR_MEMORY : BN_ERR_CUDA,      \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                \
    } while (0)

// BN_TRACE >= 3: report host-side steps of a table / volume load that take longer than 1 ms
static double now_ms();
static thread_local double g_tlast = 0;
static const bool g_t3 = getenv("BN_TRACE") && atoi(getenv("BN_TRACE")) >= 3;
#define TSTEP(label) do { if (g_t3) { const double n__ = now_ms(); if (n__ - g_tlast > 1.0) fprintf(stderr, "[bn] slow step %s: %.3f ms\n", label, n__ - g_tlast); g_tlast = n__; } } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// one-shot device buffer from the stream-ordered pool (no implicit device synchronisation, unlike
// cudaMalloc): chunk tables are built per volume, and the host-buffer entry point builds one per call
template <typename T>
struct PoolBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaStream_t st = nullptr;
    cudaError_t reserve(size_t n, cudaStream_t stream)
    {
        if (n <= cap) return cudaSuccess;
        release();
        st = stream;
        cudaError_t e = dev_malloc((void **)&p, n * sizeof(T), st);
        if (e == cudaSuccess) cap = n; else p = nullptr;
        return e;
    }
    void release() { if (p) dev_free(p, st); p = nullptr; cap = 0; }
};

// grow-only pinned host buffer (target of the D2H result copies: pageable targets go through a
// driver staging copy)
template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 1024;
        cudaError_t e = cudaMallocHost((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Workspace {
    DevBuf<SeedHit> hits_a, hits_b;
    DevBuf<uint64_t> keys_a, keys_b;
    DevBuf<uint32_t> heads, leaders, buckets;
    DevBuf<uint64_t> keys_tmp;
    DevBuf<SpecResult> spec;
    DevBuf<int4> cells;
    DevBuf<DevInitHit> init;
    DevBuf<DevGapResult> gap_out;
    DevBuf<int32_t> scratch, todo;
    DevBuf<int32_t> tri_ctx, tri_sel_ctx, tri_sel_idx;   // triage: context per init-HSP / chain links / original index per selected record
    DevBuf<int32_t> lr_state, lr_ctx, lr_next, lr_list, lr_w;      // long extensions in rounds (LongRounds)
    DevBuf<unsigned long long> lr_best;
    DevBuf<uint32_t> lr_head;
    PinnedBuf<int32_t> h_sel_idx;
    DevBuf<DevInitHit> tri_init;
    DevBuf<DevGapResult> tri_gap;
    DevBuf<uint2> tri_table;
    PinnedBuf<uint2> h_table;
    DevBuf<unsigned long long> counters;
    DevBuf<uint8_t> cub_temp;
    unsigned long long *h_counters = nullptr;   // pinned
    PinnedBuf<DevInitHit> h_init;
    PinnedBuf<DevGapResult> h_gap;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // stage timers (0..5) + "search done" (6), created once
    void release()
    {
        h_init.release(); h_gap.release(); h_table.release();
        tri_ctx.release(); tri_sel_ctx.release(); tri_sel_idx.release(); tri_init.release(); tri_gap.release(); tri_table.release();
        lr_state.release(); lr_ctx.release(); lr_next.release(); lr_list.release(); lr_w.release(); lr_best.release(); lr_head.release();
        h_sel_idx.release();
        for (auto &e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
        hits_a.release(); hits_b.release(); keys_a.release(); keys_b.release();
        cells.release(); heads.release(); leaders.release(); buckets.release(); keys_tmp.release(); spec.release(); init.release(); gap_out.release(); scratch.release(); todo.release();
        counters.release(); cub_temp.release();
        if (h_counters) cudaFreeHost(h_counters);
        h_counters = nullptr;
    }
};

// One execution lane of a GPU: a stream pair and a workspace of its own.  A search (or a pipeline, or a table
// build) holds one lane from start to end, so callers on several threads (blastn -num_threads, one thread per
// volume in bn_prelim_search_volumes) run concurrently on the same device instead of queueing on a device lock:
// the engine is re-entrant per GPU as api/prelim_search_runner.hpp:94-113 expects of the word finder.
struct Lane {
    int id = -1;                          // CUDA device ordinal
    int index = 0;                        // lane number on its device
    cudaStream_t stream = nullptr;
    // high-priority stream for what follows the scan kernel in the fused pipeline: when several searches are in flight
    // on the device (job pipeline, concurrent callers) the small latency-bound kernels of the search that is ahead
    // get their blocks scheduled in front of the thousands of pending scan blocks of the searches behind it
    cudaStream_t tail_stream = nullptr;
    cudaEvent_t scan_ev = nullptr;
    cudaEvent_t alloc_ev = nullptr;
    PinnedBuf<uint8_t> stage;             // pinned staging of a pipeline job's small uploads (Stager)
    DevArena arena;                       // device memory of a pipeline job's own batch / volume (devmem.h)
    Workspace w;
    Workspace &ws() { return w; }
    std::mutex mu;                        // held by the thread that owns the lane
};

struct Gpu {
    int id = -1;
    // volume uploads of the host-buffer entry points, ONE queue per device: uploads of consecutive jobs cross PCIe one
    // after the other (each at full bandwidth, the first one done first) instead of sharing the link and all
    // finishing together; they overlap the table builds and searches on the lanes' streams
    cudaStream_t copy_stream = nullptr;
    cudaStream_t free_stream = nullptr;   // never carries work: a cudaFreeAsync on it completes at once, so the block is
                                          // reusable by the next allocation on any lane (the callers free after a host sync)
    std::vector<std::unique_ptr<Lane>> lanes;
    // lane hand-out: a caller takes all the lanes it needs in one step (a job pipeline: three + one for its traceback
    // stage) or waits holding none, so callers that need several lanes cannot deadlock on each other's partial sets
    std::mutex pool_mu;
    std::condition_variable pool_cv;
};

// RAII ownership of one lane (LaneLock) or several (acquire_lanes): free lanes are taken under the device's pool
// mutex; when there are not enough the caller waits for a release
struct LaneLock {
    Lane *lane = nullptr;
    Gpu *gpu = nullptr;
    LaneLock() = default;
    explicit LaneLock(Gpu &g) : gpu(&g)
    {
        std::unique_lock<std::mutex> lk(g.pool_mu);
        for (;;) {
            for (auto &l : g.lanes)
                if (l->mu.try_lock()) { lane = l.get(); return; }
            g.pool_cv.wait(lk);
        }
    }
    LaneLock(LaneLock &&o) noexcept : lane(o.lane), gpu(o.gpu) { o.lane = nullptr; }
    LaneLock &operator=(LaneLock &&o) noexcept { release(); lane = o.lane; gpu = o.gpu; o.lane = nullptr; return *this; }
    LaneLock(const LaneLock &) = delete;
    LaneLock &operator=(const LaneLock &) = delete;
    void release()
    {
        if (!lane) return;
        {
            std::lock_guard<std::mutex> lk(gpu->pool_mu);
            lane->mu.unlock();
        }
        gpu->pool_cv.notify_all();
        lane = nullptr;
    }
    ~LaneLock() { release(); }
    // adopt a lane whose mutex the caller already holds (acquire_lanes)
    static LaneLock adopt(Gpu &g, Lane *l) { LaneLock k; k.gpu = &g; k.lane = l; return k; }
};

// k lanes at once (k <= lanes of the device), or none while waiting
static void acquire_lanes(Gpu &g, int k, LaneLock *out)
{
    std::unique_lock<std::mutex> lk(g.pool_mu);
    for (;;) {
        std::vector<Lane *> got;
        for (auto &l : g.lanes) {
            if ((int)got.size() == k) break;
            if (l->mu.try_lock()) got.push_back(l.get());
        }
        if ((int)got.size() == k) {
            for (int i = 0; i < k; i++) out[i] = LaneLock::adopt(g, got[(size_t)i]);
            return;
        }
        for (Lane *l : got) l->mu.unlock();
        g.pool_cv.wait(lk);
    }
}

struct ChunkTable {
    std::vector<DevChunk> host;               // chunks (what seed hits, init-HSPs and the host replay refer to)
    std::vector<DevChunk> units;              // scan units of a masked volume (empty: the chunks are the units)
    std::vector<int2> ranges;                 // unmasked ranges of masked chunks
    std::vector<HostChunk> hchunks;
    std::vector<int32_t> h_block_chunk;         // sources of the asynchronous uploads: live as long as the table
    std::vector<ScanBlockDesc> h_block_desc;
    PoolBuf<DevChunk> dev, units_dev;
    PoolBuf<int2> ranges_dev;
    const DevChunk *scan_units() const { return units.empty() ? dev.p : units_dev.p; }
    int32_t n_scan_units() const { return (int32_t)(units.empty() ? host.size() : units.size()); }
    PoolBuf<int32_t> block_chunk;
    PoolBuf<ScanBlockDesc> block_desc;
    int64_t total_pos = 0;
    int64_t total_bases = 0;
    int64_t n_blocks = 0;
    int device_id = -1;
    cudaEvent_t ready = nullptr;              // recorded behind the uploads: lanes other than the builder's wait on it
    uint64_t last_use = 0;
    ChunkTable() = default;
    ChunkTable(const ChunkTable &) = delete;
    ChunkTable &operator=(const ChunkTable &) = delete;
    ~ChunkTable()
    {
        int cur = -1;
        cudaGetDevice(&cur);
        if (device_id >= 0 && cur != device_id) cudaSetDevice(device_id);
        if (ready) { cudaEventSynchronize(ready); cudaEventDestroy(ready); }      // the uploads read this object's vectors
        dev.release(); units_dev.release(); ranges_dev.release(); block_chunk.release(); block_desc.release();
        if (device_id >= 0 && cur >= 0 && cur != device_id) cudaSetDevice(cur);
    }
};

struct Volume {
    int device = 0;
    cudaEvent_t ready = nullptr;   // set while an asynchronous upload may still be in flight
    uint8_t *d_raw = nullptr;      // allocation
    uint8_t *d_packed = nullptr;   // d_raw + 64
    int64_t bytes = 0;
    std::vector<int64_t> byte_off;
    std::vector<int32_t> seq_len;
    // database masks (bn_db_set_masks): per sequence mask_first[i]..mask_first[i+1] masked [begin, end) pairs
    int32_t mask_type = 0, mask_version = 0;
    std::vector<int64_t> mask_first;
    std::vector<int32_t> mask_iv;
    // whole-volume chunk tables per table shape (a handful at most: least recently used goes first); tables of
    // partial oid ranges are temporaries of the call that needs them
    std::mutex tmu;                // guards tables, the mask fields and use_clock
    std::map<std::string, std::shared_ptr<ChunkTable>> tables;
    uint64_t use_clock = 0;
    // ambiguity data of a BLAST DB volume (bn_db_load_files): per sequence amb_first[i]..amb_first[i+1] runs
    // {start, length, blastna value} in amb_runs; empty for volumes without ambiguities
    std::vector<int64_t> amb_first;      // n_seq + 1 (empty: the volume has no ambiguity data)
    std::vector<int32_t> amb_runs;       // as given: flat {first base, bases, blastna code} in application order
    std::vector<int64_t> amb_dev_first;  // per sequence: first entry of its runs in the device table
    int4 *d_amb = nullptr;               // {first base, end, blastna code, 0}, sorted and disjoint per sequence
    int32_t *d_amb_runs = nullptr;       // (unused, kept for free_volume_dev)
    int64_t *d_amb_first = nullptr;
    bool has_ambiguity() const { return d_amb != nullptr; }
    int32_t amb_first_of(int32_t oid) const { return amb_dev_first.empty() ? 0 : (int32_t)amb_dev_first[(size_t)oid]; }
    int32_t amb_count_of(int32_t oid) const { return amb_dev_first.empty() ? 0 : (int32_t)(amb_dev_first[(size_t)oid + 1] - amb_dev_first[(size_t)oid]); }
};

struct QueryDev {
    uint8_t *query = nullptr;
    DevContext *ctx = nullptr;
    int32_t *next_pos = nullptr;
    int16_t *backbone = nullptr, *overflow = nullptr;
    int32_t *na_cells = nullptr, *na_overflow = nullptr;
    int32_t *score_table = nullptr, *matrix = nullptr;
    uint2 *qpk = nullptr;
    uint2 *prk = nullptr;
    uint4 *cinfo = nullptr;
    uint4 *qinfo = nullptr;
    uint32_t *sig = nullptr;
    uint16_t *psig = nullptr;
    uint32_t *filt = nullptr;
    DevQuery view{};
    bool ready = false;
};

struct Query {
    BnQueryBatch batch{};                 // host copy of the scalars; array pointers are NULL except contexts
    std::vector<BnContext> ctx;           // the host replay needs contexts, cutoffs and Karlin blocks only
    std::vector<CtxLite> ctx_lite;        // compact context table of the replay
    std::vector<QueryDev> dev;            // per device
    int32_t diag_array_length = 1;
    int32_t max_query_length = 0;
    bool fast_path_refused = false;       // the device-grouped word finder did not apply to this batch last time
    // direct filter of the scan kernel (scan_kernel.cu): lut == word, one-hit mode, scoring tables of the plain
    // match / mismatch form the filter evaluates in closed form
    bool direct_ok = false;
    int32_t uni_ok = 0, uni_x = 0, uni_cutoff = 0, uni_reduced = 0;
};

// Handle tables.  g_mu guards the three vectors; a call resolves its handles to shared_ptrs under it and
// works on those, so a concurrent load / free on another thread can neither move nor destroy what it uses.
// Freed slots are recycled.
static std::mutex g_mu;
static std::vector<std::unique_ptr<Gpu>> g_devices;
static std::vector<std::shared_ptr<Volume>> g_volumes;
static std::vector<std::shared_ptr<Query>> g_queries;
static bool g_inited = false;

template <typename T>
static int put_handle(std::vector<std::shared_ptr<T>> &tab, std::shared_ptr<T> obj)      // g_mu held
{
    for (size_t i = 0; i < tab.size(); i++)
        if (!tab[i]) { tab[i] = std::move(obj); return (int)i; }
    tab.push_back(std::move(obj));
    return (int)tab.size() - 1;
}

static int ensure_init()
{
    if (g_inited) return BN_OK;
    return bn_init(0, nullptr);
}

static Gpu *device_at(int d)
{
    if (d < 0 || d >= (int)g_devices.size()) return nullptr;
    return g_devices[d].get();
}

// ------------------------------------------------------------------------------------------------
// Device arrays of a query batch come from the stream-ordered pool (release threshold = never), so
// loading a new batch re-uses the previous batch's memory without touching the OS allocator.
static void free_query_dev(QueryDev &q, cudaStream_t st)
{
    void *ptrs[] = {q.query, q.ctx, q.next_pos, q.backbone, q.overflow, q.na_cells, q.na_overflow,
                    q.score_table, q.matrix, q.qpk, q.prk, q.cinfo, q.qinfo, q.sig, q.psig, q.filt};
    for (void *p : ptrs) if (p) dev_free(p, st);
    q = QueryDev{};
}

template <typename T>
static cudaError_t dev_alloc(T **dst, size_t n, cudaStream_t st)
{
    *dst = nullptr;
    if (n == 0) return cudaSuccess;
    return dev_malloc((void **)dst, n * sizeof(T), st);
}

// Pinned staging of small host->device copies (job pipeline): a cudaMemcpyAsync from pageable memory blocks the
// calling thread until the copy engine has taken the data, i.e. behind whatever volume upload is in flight; staged
// through a pinned buffer of the lane the copy is queued and the thread moves on.  The buffer belongs to one job at
// a time (the lane's previous job is complete when the next one is prepared).
struct Stager {
    uint8_t *base = nullptr;
    size_t cap = 0, used = 0;
    const void *stage(const void *src, size_t bytes)
    {
        const size_t at = (used + 255) & ~size_t(255);
        if (!base || at + bytes > cap) return src;          // does not fit: plain (blocking) copy
        memcpy(base + at, src, bytes);
        used = at + bytes;
        return base + at;
    }
};
static thread_local Stager *g_stager = nullptr;
static inline const void *staged(const void *src, size_t bytes) { return g_stager ? g_stager->stage(src, bytes) : src; }

template <typename T>
static cudaError_t upload(T **dst, const T *src, size_t n, cudaStream_t st)
{
    cudaError_t e = dev_alloc(dst, n, st);
    if (e != cudaSuccess || n == 0) return e;
    return cudaMemcpyAsync(*dst, staged(src, n * sizeof(T)), n * sizeof(T), cudaMemcpyHostToDevice, st);
}

cudaError_t launch_dust(const uint8_t *seqs, const int64_t *seq_off, const int32_t *lens, int32_t n_seq, uint32_t level,
                        uint32_t window, uint32_t linker, const int64_t *out_off, int32_t *out, int32_t *out_n,
                        int32_t *compact, int64_t *compact_off, unsigned long long *cursor, cudaStream_t st);
cudaError_t launch_build_presence(const int32_t *hashtable, int64_t hashsize, uint32_t *presence, cudaStream_t st);
cudaError_t launch_build_qpk(const uint8_t *query_start, int32_t concat_len, uint2 *qpk, int64_t nwords, cudaStream_t st);
cudaError_t launch_popc(const uint32_t *presence, int64_t nwords, uint32_t *counts, cudaStream_t st);
cudaError_t launch_build_compact(const int32_t *hashtable, const uint32_t *presence, const uint32_t *prefix,
                                 int64_t nwords, uint2 *prk, const uint4 *qinfo, uint4 *cinfo, cudaStream_t st);
cudaError_t launch_build_prk_cinfo(const uint32_t *presence, const uint32_t *prefix, int64_t nwords, uint2 *prk,
                                   const int32_t *first_qp, int64_t n_ranks, const uint4 *qinfo, uint4 *cinfo,
                                   cudaStream_t st);
cudaError_t launch_rebuild_hashtable(const DevQuery &q, int64_t hashsize, int32_t *out, cudaStream_t st);
cudaError_t build_mb_lookup_device(const uint8_t *d_query, int32_t concat_len, int32_t word_length, int32_t lut,
                                   const int32_t *d_segs, int32_t n_segs, int32_t *d_next_pos, uint32_t *d_presence,
                                   int32_t *d_first_qp, int64_t *n_launches, cudaStream_t st);

// Uploads one query batch to device d straight from the caller's arrays (no host staging copy);
// the presence bitmap and the 16-base query windows are derived on the device.
// `after_h2d` (optional) runs once every host->device copy of the batch has been queued and before the
// derivation kernels are: the host-buffer entry point starts the volume upload there, so the small
// query copies are not stuck behind it in the copy engine and the table build overlaps the upload.
static int query_to_device(Query &Q, const BnQueryBatch &src, int d, Lane *dev, const std::function<int()> *after_h2d = nullptr,
                           bool keep_async = false)
{
    QueryDev &qd = Q.dev[d];
    if (qd.ready) return BN_OK;
    CU_TRY(cudaSetDevice(dev->id));
    cudaStream_t st = dev->stream;
    const BnQueryBatch &b = Q.batch;
    std::vector<DevContext> dctx((size_t)b.num_contexts);
    for (int i = 0; i < b.num_contexts; i++) {
        const BnContext &c = Q.ctx[i];
        dctx[i] = DevContext{c.query_offset, c.query_length, c.query_index, c.frame,
                             c.x_dropoff, c.cutoff_score, c.reduced_cutoff, c.gapped_cutoff};
    }
    // ---- host -> device copies first ------------------------------------------------------------------
    TSTEP("query: begin");
    CU_TRY(upload(&qd.query, src.query_start, (size_t)b.concat_len + 2, st));
    TSTEP("query: first upload");
    CU_TRY(upload(&qd.ctx, dctx.data(), dctx.size(), st));
    CU_TRY(upload(&qd.score_table, b.nucl_score_table, (size_t)256, st));
    CU_TRY(upload(&qd.matrix, b.matrix, (size_t)256, st));
    // temporaries of the table derivation (freed stream-ordered at the end)
    int32_t *t_hashtable = nullptr, *t_first_qp = nullptr, *t_segs = nullptr;
    uint32_t *t_presence = nullptr, *t_counts = nullptr, *t_prefix = nullptr;
    const bool device_fill = b.lut_type == BN_LUT_MB && !src.hashtable;
    if (b.lut_type == BN_LUT_MB) {
        CU_TRY(dev_alloc(&qd.next_pos, (size_t)b.concat_len + 1, st));
        if (device_fill) {
            CU_TRY(upload(&t_segs, src.lookup_segments, 2 * (size_t)src.n_lookup_segments, st));
        } else {
            CU_TRY(upload(&t_hashtable, src.hashtable, (size_t)b.hashsize, st));
            CU_TRY(cudaMemcpyAsync(qd.next_pos, src.next_pos, ((size_t)b.concat_len + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        }
    } else if (b.lut_type == BN_LUT_NA) {
        static const int32_t kNoOverflow[1] = {0};
        CU_TRY(upload(&qd.na_cells, src.na_backbone, (size_t)(4 * b.hashsize), st));
        if (src.na_overflow && src.na_overflow_len > 0) CU_TRY(upload(&qd.na_overflow, src.na_overflow, (size_t)src.na_overflow_len, st));
        else CU_TRY(upload(&qd.na_overflow, kNoOverflow, (size_t)1, st));
    } else {
        static const int16_t kEmptyOverflow[2] = {-1, -1};
        CU_TRY(upload(&qd.backbone, src.backbone, (size_t)b.hashsize, st));
        if (src.overflow && b.overflow_len > 0) CU_TRY(upload(&qd.overflow, src.overflow, (size_t)b.overflow_len, st));
        else CU_TRY(upload(&qd.overflow, kEmptyOverflow, (size_t)2, st));
    }
    TSTEP("query: uploads");
    if (after_h2d) { const int rc = (*after_h2d)(); if (rc) return rc; }
    TSTEP("query: volume hook");

    // ---- derived arrays (device only) --------------------------------------------------------------------
    if (b.lut_type == BN_LUT_MB) {
        // exact presence bitmap (replaces the reference's compressed pv_array: same answers,
        // PV_TEST is only a filter in front of hashtable[index] != 0)
        CU_TRY(dev_alloc(&t_presence, (size_t)((b.hashsize + 31) / 32), st));
        if (device_fill) {
            // s_FillContigMBTable on the device: only the query bytes and the segment list cross PCIe
            CU_TRY(dev_alloc(&t_first_qp, (size_t)b.concat_len + 1, st));
            CU_TRY(cudaMemsetAsync(t_first_qp, 0, ((size_t)b.concat_len + 1) * sizeof(int32_t), st));
            CU_TRY(build_mb_lookup_device(qd.query + 1, b.concat_len, b.word_length, b.lut_word_length,
                                          t_segs, src.n_lookup_segments, qd.next_pos, t_presence,
                                          t_first_qp, nullptr, st));
        } else {
            CU_TRY(launch_build_presence(t_hashtable, b.hashsize, t_presence, st));
        }
    }
    TSTEP("query: lookup fill");
    const int64_t nw = (int64_t)((b.concat_len + 2 + 16) >> 4) + 3;
    CU_TRY(dev_alloc(&qd.qpk, (size_t)nw, st));
    CU_TRY(launch_build_qpk(qd.query, b.concat_len, qd.qpk, nw, st));
    DevQuery &v = qd.view;
    v.query = qd.query + 1;
    v.concat_len = b.concat_len;
    v.ctx = qd.ctx; v.num_contexts = b.num_contexts;
    v.lut_type = b.lut_type; v.word_length = b.word_length; v.lut_word_length = b.lut_word_length;
    v.scan_step = b.scan_step; v.hash_mask = (uint32_t)(b.hashsize - 1);
    v.next_pos = qd.next_pos;
    v.backbone = qd.backbone; v.overflow = qd.overflow;
    v.na_cells = reinterpret_cast<const int4 *>(qd.na_cells); v.na_overflow = qd.na_overflow;
    v.has_locations = Q.batch.masked_locations != nullptr;
    v.container_type = b.container_type; v.window_size = b.window_size; v.scan_range = b.scan_range;
    v.score_table = qd.score_table; v.matrix = qd.matrix; v.qpk = qd.qpk;
    v.gap_algo = b.gap_algo; v.reward = b.reward; v.penalty = b.penalty;
    v.gap_open = b.gap_open; v.gap_extend = b.gap_extend; v.gap_x_dropoff = b.gap_x_dropoff;
    if (b.lut_type == BN_LUT_MB) {
        CU_TRY(dev_alloc(&qd.qinfo, (size_t)b.concat_len + 2, st));
        // which query positions are in the table (chain heads + everything a link points to): the direct filter of the
        // scan kernel asks it of a hit's predecessor on the diagonal
        uint32_t *t_indexed = nullptr;
        CU_TRY(dev_alloc(&t_indexed, (size_t)(b.concat_len + 1 + 32) / 32 + 2, st));
        CU_TRY(launch_build_qinfo(v, qd.next_pos, b.concat_len, device_fill ? t_first_qp : t_hashtable,
                                  device_fill ? (int64_t)b.concat_len + 1 : b.hashsize, t_indexed, qd.qinfo, st));
        CU_TRY(dev_free(t_indexed, st));
        v.qinfo = qd.qinfo;
        // compact table: {presence word, rank} per 32 cells (4^lut / 4 bytes, L2-resident) + the first
        // chain element of every occupied cell in cell order.  It stands in for hashtable[] everywhere
        // on the device (mb_cell), so the 4^lut-entry table is not kept in HBM.
        const int64_t nwords = (b.hashsize + 31) / 32;
        CU_TRY(dev_alloc(&t_counts, (size_t)nwords, st));
        CU_TRY(dev_alloc(&t_prefix, (size_t)nwords, st));
        CU_TRY(dev_alloc(&qd.prk, (size_t)nwords, st));
        CU_TRY(dev_alloc(&qd.cinfo, 2 * ((size_t)b.concat_len + 2), st));
        CU_TRY(launch_popc(t_presence, nwords, t_counts, st));
        CU_TRY(dev->ws().cub_temp.reserve(prefix_sum_temp_bytes(nwords)));
        CU_TRY(prefix_sum_u32(t_counts, t_prefix, nwords, false, dev->ws().cub_temp.p, st));
        if (device_fill)
            CU_TRY(launch_build_prk_cinfo(t_presence, t_prefix, nwords, qd.prk, t_first_qp, (int64_t)b.concat_len + 1,
                                          qd.qinfo, qd.cinfo, st));
        else
            CU_TRY(launch_build_compact(t_hashtable, t_presence, t_prefix, nwords, qd.prk, qd.qinfo, qd.cinfo, st));
        v.prk = qd.prk;
        v.cinfo = qd.cinfo;
        // flank signatures of the first chain elements (scan kernel pre-filter); unfilled ranks are never addressed
        CU_TRY(dev_alloc(&qd.sig, (size_t)b.concat_len + 2, st));
        CU_TRY(launch_build_sig(qd.cinfo, (int64_t)b.concat_len + 2, qd.sig, st));
        v.sig = getenv("BN_NO_SIG") ? nullptr : qd.sig;
        // the same signatures per table cell: one 2-byte gather per scan position instead of {presence, rank} first and the
        // signature of an occupied cell after it (hashsize even: a power of 4).  32 MB at lut 12: stays in L2
        // (`profiles/r04c_gather_table_size_and_serial_trips.txt`: the gather rate does not depend on the table size up to 64 MB)
        if (v.sig && b.word_length - b.lut_word_length >= 7 && b.lut_word_length <= 12 && !getenv("BN_NO_PSIG")) {
            CU_TRY(dev_alloc(&qd.psig, (size_t)b.hashsize, st));
            CU_TRY(launch_build_psig(qd.prk, qd.sig, b.hashsize, qd.psig, st));
            v.psig = qd.psig;
        }
        // small batch: hashed presence filter for the shared-memory scan (scan_kernel_filtered).  The number of occupied
        // cells is only known on the device; concat_len bounds it.  Measured against the queue-driven kernel on a
        // 250 Mb volume: 1.75x faster at 1 % - 2 % of the filter's bits set, even at ~15 %, so the path is taken up to 1/16
        {
            const long filt_max = getenv("BN_FILT_MAX") ? atol(getenv("BN_FILT_MAX")) : (long)(FILT_BITS / 16);   // test switch, read per load
            if ((long)b.concat_len <= filt_max) {
                CU_TRY(dev_alloc(&qd.filt, (size_t)(FILT_BITS / 32), st));
                CU_TRY(launch_build_filter(t_presence, nwords, qd.filt, st));
                v.filt = qd.filt;
            }
        }
    }
    TSTEP("query: derived tables");
    {
        void *tmp[] = {t_hashtable, t_first_qp, t_segs, t_presence, t_counts, t_prefix};
        for (void *p : tmp) if (p) CU_TRY(dev_free(p, st));
    }
    TSTEP("query: frees");
    // caller's arrays may go away after bn_query_load returns (a pipeline's stay until its call does: keep_async)
    if (!keep_async) CU_TRY(cudaStreamSynchronize(st));
    qd.ready = true;
    return BN_OK;
}

// ------------------------------------------------------------------------------------------------
// Chunk table: the reference's subject split (s_GetNextSubjectChunk core/blast_engine.c:220-301,
// unmasked subjects) + scan-position prefix sums + the evolution of the diagonal container's
// `offset` (Blast_ExtendWordExit core/blast_extend.c:164-186).
// ------------------------------------------------------------------------------------------------
static const size_t kMaxCachedTables = 4;

static int build_chunk_table(Volume &V, const Query &Q, int32_t oid_begin, int32_t oid_end,
                             Lane &L, std::shared_ptr<ChunkTable> *out)
{
    cudaStream_t st = L.stream;
    const BnQueryBatch &b = Q.batch;
    const bool whole = oid_begin == 0 && oid_end == (int32_t)V.seq_len.size();
    char key[160];
    std::unique_lock<std::mutex> tlk(V.tmu);
    snprintf(key, sizeof key, "%d/%d/%d/%d/%d", b.lut_word_length, b.scan_step, b.window_size, b.word_length, V.mask_version);
    if (whole) {
        auto it = V.tables.find(key);
        if (it != V.tables.end()) {
            it->second->last_use = ++V.use_clock;
            *out = it->second;
            tlk.unlock();
            // built on another lane's stream, perhaps a moment ago
            if ((*out)->ready) CU_TRY(cudaStreamWaitEvent(st, (*out)->ready, 0));
            return BN_OK;
        }
    }

    // s_GetNextSubjectChunk (core/blast_engine.c:220-301): 200 Mb chunks with a 100-base overlap inside every
    // hard range (the whole sequence without hard masks), chunk starts rounded down to a byte; soft ranges
    // clipped to the chunk.  BlastNaWordFinder (core/na_ungapped.c:1610-1645) then scans every unmasked range
    // of a masked subject from left + (word - lut).
    auto T = std::make_shared<ChunkTable>();
    T->device_id = L.id;
    const int32_t lut = b.lut_word_length, step = b.scan_step, window = b.window_size;
    const int32_t ext_to = b.word_length - lut;
    const int32_t mt = V.mask_type;
    int32_t diag_offset = window, epoch = 0;
    int64_t prefix = 0;
    std::vector<int32_t> R;                               // unmasked ranges of the sequence, flat pairs
    for (int32_t oid = oid_begin; oid < oid_end; oid++) {
        const int32_t full = V.seq_len[oid];
        int32_t full_range[2] = {0, full};
        const int32_t *hard = full_range, *soft = full_range;
        int32_t n_hard = 1, n_soft = 1;
        if (mt) {
            const int64_t m0 = V.mask_first[(size_t)oid], m1 = V.mask_first[(size_t)oid + 1];
            const int32_t n_r = (int32_t)(m1 - m0) + 1;
            R.assign((size_t)n_r * 2, 0);
            for (int64_t k = m0; k < m1; k++) {
                R[(size_t)(2 * (k - m0) + 1)] = V.mask_iv[(size_t)(2 * k)];
                R[(size_t)(2 * (k - m0) + 2)] = V.mask_iv[(size_t)(2 * k + 1)];
            }
            R[0] = 0; R[(size_t)(2 * (n_r - 1) + 1)] = full;      // BlastSeqBlkSetSeqRanges core/blast_util.c:216-218
            if (mt == 2) { hard = R.data(); n_hard = n_r; } else { soft = R.data(); n_soft = n_r; }
        }
        int32_t hm = 0;
        int32_t next = hard[0];
        while (next < full) {
            const int32_t residual = next % 4;
            const int32_t offset = next - residual;
            DevChunk c{};
            c.byte_off = V.byte_off[oid] + offset / 4;
            c.oid = oid; c.chunk_off = offset;
            if ((int64_t)offset + BN_MAX_DBSEQ_LEN < (int64_t)hard[2 * hm + 1]) {
                c.len = BN_MAX_DBSEQ_LEN;
                next = offset + BN_MAX_DBSEQ_LEN - BN_DBSEQ_CHUNK_OVERLAP;
            } else {
                c.len = hard[2 * hm + 1] - offset;
                ++hm;
                next = hm < n_hard ? hard[2 * hm] : full;
            }
            // the chunk's unmasked ranges (chunk-relative); unmasked volume: the chunk itself
            std::vector<int2> cr;
            if (mt) {
                if (offset == 0 && residual == 0 && next == full) {
                    for (int32_t i = 0; i < n_soft; i++) cr.push_back(make_int2(soft[2 * i], soft[2 * i + 1]));
                } else if (mt != 1) cr.push_back(make_int2(residual, c.len));
                else {
                    int32_t i = 0;
                    const int32_t end = offset + c.len;
                    while (soft[2 * i + 1] < offset) ++i;
                    for (; i < n_soft && soft[2 * i] < end; ++i) cr.push_back(make_int2(soft[2 * i] - offset, soft[2 * i + 1] - offset));
                    if (cr.empty()) continue;                         // SUBJECT_SPLIT_NO_RANGE: the chunk is skipped
                    cr.front().x = std::max(cr.front().x, 0);
                    cr.back().y = std::min(cr.back().y, c.len);
                }
            }
            const int32_t ci = (int32_t)T->host.size();
            c.diag_offset = diag_offset; c.diag_epoch = epoch;
            c.parent = ci; c.p_first = 0; c.s_range = c.len;
            c.pos_prefix = prefix;
            if (!mt) {
                c.npos = c.len >= lut ? (c.len - lut) / step + 1 : 0;
                prefix += c.npos;
            } else {
                c.range_first = (int32_t)T->ranges.size(); c.n_ranges = (int32_t)cr.size();
                int64_t total = 0;
                for (const int2 &r : cr) {
                    DevChunk u = c;
                    u.p_first = r.x + ext_to; u.s_range = r.y;
                    const int32_t last = r.y - lut;
                    u.npos = last >= u.p_first ? (last - u.p_first) / step + 1 : 0;
                    u.pos_prefix = prefix + total;
                    total += u.npos;
                    T->units.push_back(u);
                    T->ranges.push_back(r);
                }
                c.npos = (int32_t)std::min<int64_t>(total, INT32_MAX);
                prefix += total;
            }
            T->total_bases += c.len;
            T->host.push_back(c);
            T->hchunks.push_back(HostChunk{oid, offset, c.len});
            if (diag_offset >= INT32_MAX / 4) { diag_offset = window; ++epoch; }
            else diag_offset += c.len + window;
        }
    }
    const std::vector<DevChunk> &SU = T->units.empty() ? T->host : T->units;         // what the scan kernel walks
    T->total_pos = prefix;
    const int ppb = scan_positions_per_block();
    T->n_blocks = (prefix + ppb - 1) / ppb;
    std::vector<int32_t> &bc = T->h_block_chunk;
    bc.assign((size_t)T->n_blocks + 1, 0);
    {
        size_t c = 0;
        const size_t n = SU.size();
        for (int64_t blk = 0; blk < T->n_blocks; blk++) {
            const int64_t g = blk * ppb;
            while (c + 1 < n && SU[c + 1].pos_prefix <= g) ++c;
            bc[(size_t)blk] = (int32_t)c;
        }
        bc[(size_t)T->n_blocks] = n ? (int32_t)n - 1 : 0;
    }
    // staged scan kernel: the slice of the volume every block copies into shared memory
    std::vector<ScanBlockDesc> &bd = T->h_block_desc;
    bd.assign((size_t)T->n_blocks, ScanBlockDesc{});
    {
        const int32_t tile_cap = scan_tile_cap(step, b.word_length), margin = scan_tile_margin();
        const int32_t maxc = scan_max_block_chunks();
        size_t c = 0;
        const size_t n = SU.size();
        for (int64_t blk = 0; blk < T->n_blocks; blk++) {
            const int64_t g0 = blk * ppb, g_last = std::min<int64_t>(g0 + ppb, prefix) - 1;
            const int32_t c_lo = bc[(size_t)blk];
            c = std::max<size_t>(c, (size_t)c_lo);
            while (c + 1 < n && SU[c + 1].pos_prefix <= g_last) ++c;
            const DevChunk &a = SU[(size_t)c_lo], &z = SU[c];
            const int64_t first_byte = a.byte_off + ((a.p_first + (g0 - a.pos_prefix) * step) >> 2);
            const int64_t last_byte = z.byte_off + (((z.p_first + (g_last - z.pos_prefix) * step) + b.word_length + 32) >> 2);
            ScanBlockDesc d{};
            d.tile_lo = (first_byte - margin) & ~int64_t(15);
            const int64_t bytes = (last_byte + margin - d.tile_lo + 15) & ~int64_t(15);
            d.c_lo = c_lo; d.c_hi = (int32_t)c;
            d.staged = (bytes <= (int64_t)tile_cap && (d.c_hi - d.c_lo) < maxc) ? 1 : 0;
            d.bytes = d.staged ? (int32_t)bytes : 0;
            bd[(size_t)blk] = d;
        }
    }
    if (!T->host.empty()) {
        CU_TRY(T->dev.reserve(T->host.size(), st));
        CU_TRY(cudaMemcpyAsync(T->dev.p, staged(T->host.data(), T->host.size() * sizeof(DevChunk)), T->host.size() * sizeof(DevChunk),
                               cudaMemcpyHostToDevice, st));
        if (!T->units.empty()) {
            CU_TRY(T->units_dev.reserve(T->units.size(), st));
            CU_TRY(cudaMemcpyAsync(T->units_dev.p, staged(T->units.data(), T->units.size() * sizeof(DevChunk)), T->units.size() * sizeof(DevChunk),
                                   cudaMemcpyHostToDevice, st));
            CU_TRY(T->ranges_dev.reserve(T->ranges.size() + 1, st));
            CU_TRY(cudaMemcpyAsync(T->ranges_dev.p, staged(T->ranges.data(), T->ranges.size() * sizeof(int2)), T->ranges.size() * sizeof(int2),
                                   cudaMemcpyHostToDevice, st));
        }
        CU_TRY(T->block_chunk.reserve(bc.size(), st));
        CU_TRY(cudaMemcpyAsync(T->block_chunk.p, staged(bc.data(), bc.size() * sizeof(int32_t)), bc.size() * sizeof(int32_t),
                               cudaMemcpyHostToDevice, st));
        CU_TRY(T->block_desc.reserve(bd.size() + 1, st));
        CU_TRY(cudaMemcpyAsync(T->block_desc.p, staged(bd.data(), bd.size() * sizeof(ScanBlockDesc)), bd.size() * sizeof(ScanBlockDesc),
                               cudaMemcpyHostToDevice, st));
        // no synchronisation: the sources are members of the table; consumers on `st` are ordered behind the
        // copies, consumers on other lanes (and the destructor) wait on `ready`
        CU_TRY(cudaEventCreateWithFlags(&T->ready, cudaEventDisableTiming));
        CU_TRY(cudaEventRecord(T->ready, st));
    }
    if (whole) {
        if (V.tables.size() >= kMaxCachedTables) {
            auto victim = V.tables.begin();
            for (auto it = V.tables.begin(); it != V.tables.end(); ++it)
                if (it->second->last_use < victim->second->last_use) victim = it;
            V.tables.erase(victim);          // a search still using it holds its own reference
        }
        T->last_use = ++V.use_clock;
        V.tables[key] = T;
    }
    *out = T;
    return BN_OK;
}

// ------------------------------------------------------------------------------------------------
struct Timer {
    cudaEvent_t a, b;
    cudaStream_t st;
    bool own;
    explicit Timer(cudaStream_t s) : st(s), own(true) { cudaEventCreate(&a); cudaEventCreate(&b); }
    // borrows a pair of long-lived events (slot 0..2) of the workspace
    Timer(cudaStream_t s, Workspace &ws, int slot) : st(s), own(false)
    {
        for (int k = 2 * slot; k < 2 * slot + 2; k++) if (!ws.ev[k]) cudaEventCreate(&ws.ev[k]);
        a = ws.ev[2 * slot]; b = ws.ev[2 * slot + 1];
    }
    ~Timer() { if (own) { cudaEventDestroy(a); cudaEventDestroy(b); } }
    void start() { cudaEventRecord(a, st); }
    void stop() { cudaEventRecord(b, st); }
    double ms() { float f = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&f, a, b); return f; }
};

static double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// two-hit mode whose single-word hits search neighbouring diagonals (core/na_ungapped.c:697, :853)
static bool serial_replay(const BnQueryBatch &b)
{
    return b.window_size > 0 && std::min(b.scan_range, b.window_size - b.word_length) > 0;
}

static int bits_for(uint64_t v) { int b = 1; while (b < 64 && (v >> b)) ++b; return b; }

struct StageCounts { int64_t n_hits = 0, lookup_hits = 0, n_init = 0, n_extended = 0; };

// What the GPU side of one search leaves for the host: counts and the pinned mirrors of the init-HSPs and
// their speculative gapped results (they belong to the workspace that was current during the search).
struct GpuOut {
    std::shared_ptr<ChunkTable> T;
    StageCounts cnt;
    DevInitHit *h_init = nullptr;
    DevGapResult *h_gap = nullptr;
    int32_t oid_begin = 0, oid_end = 0;
    double t0 = 0, t_table = 0, t_wf = 0, t_gap = 0;
    // triage (triage_kernel.cu): h_init / h_gap hold n_records selected init-HSPs (winners + undecided losers) instead
    // of all cnt.n_init, and counted_losers extensions are known to have been made without being replayed
    bool triaged = false;
    int64_t n_records = 0, counted_losers = 0;
};


static void set_direct_filter(ScanLaunch &s, const Query &Q, const ChunkTable &T, bool raw_pairs)
{
    s.direct_filter = (Q.direct_ok && !raw_pairs && T.units.empty()) ? 1 : 0;      // unmasked volumes only
    s.direct_dense = getenv("BN_NO_DIRECT_DENSE") ? 0 : 1;
    s.uni_ok = Q.uni_ok; s.uni_x = Q.uni_x; s.uni_cutoff = Q.uni_cutoff; s.uni_reduced = Q.uni_reduced;
}

// scan -> one stable radix sort on (diagonal group, global position) -> diagonal/ungapped kernel.
// Leaves the sorted seed hits in ws.hits_b and the init hits (unsorted) in ws.init.
static int run_word_finder(Lane &D, Volume &V, Query &Q, ChunkTable &T, bool raw_pairs,
                           StageCounts &cnt, BnStats *stats)
{
    Workspace &ws = D.ws();
    cudaStream_t st = D.stream;
    const DevQuery &dq = Q.dev[V.device].view;
    CU_TRY(ws.counters.reserve(16));
    if (!ws.h_counters) CU_TRY(cudaMallocHost(&ws.h_counters, 16 * sizeof(unsigned long long)));
    if (T.total_pos >= (int64_t)1 << 32) return fail(BN_ERR_OVERFLOW, "more than 2^32 scan positions in one search");
    const int gbits = bits_for((uint64_t)std::max<int64_t>(T.total_pos, 1));
    // off-diagonal two-hit search: neighbouring diagonals live in other buckets / cells -> one serial group
    const bool serial = !raw_pairs && serial_replay(Q.batch);
    const int grp_bits = (raw_pairs || serial) ? 0 : (Q.batch.container_type == BN_DIAG_HASH ? 9 : bits_for((uint64_t)Q.diag_array_length));

    Timer t_scan(st, ws, 0), t_ext(st, ws, 1);
    int64_t cap = std::max<int64_t>((int64_t)ws.hits_a.cap, std::max<int64_t>(1 << 16, T.total_pos / 16));
    for (int attempt = 0;; attempt++) {
        CU_TRY(ws.hits_a.reserve((size_t)cap));
        CU_TRY(ws.keys_a.reserve((size_t)cap));
        cap = (int64_t)std::min(ws.hits_a.cap, ws.keys_a.cap);
        CU_TRY(cudaMemsetAsync(ws.counters.p, 0, 8 * sizeof(unsigned long long), st));
        ScanLaunch s{};
        s.packed = V.d_packed; s.chunks = T.scan_units(); s.n_chunks = T.n_scan_units();
        s.total_pos = T.total_pos; s.hits = ws.hits_a.p; s.keys = ws.keys_a.p;
        s.counters = ws.counters.p; s.capacity = cap; s.block_chunk = T.block_chunk.p; s.block_desc = T.block_desc.p;
        s.raw_pairs = raw_pairs ? 1 : 0; s.gbits = gbits; s.diag_array_length = Q.diag_array_length;
        s.one_group = serial ? 1 : 0;
        s.tile_cap = scan_tile_cap(Q.batch.scan_step, Q.batch.word_length);
        set_direct_filter(s, Q, T, raw_pairs);
        t_scan.start();
        CU_TRY(launch_scan(dq, s, st));
        t_scan.stop();
        if (stats) stats->kernel_launches += 1;
        CU_TRY(cudaMemcpyAsync(ws.h_counters, ws.counters.p, 8 * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        cnt.n_hits = (int64_t)ws.h_counters[0];
        cnt.lookup_hits = (int64_t)ws.h_counters[1];
        if (cnt.n_hits <= cap) break;
        if (attempt > 2) return fail(BN_ERR_OVERFLOW, "seed-hit buffer overflow");
        cap = cnt.n_hits + cnt.n_hits / 16 + 1024;
    }
    if (stats) stats->ms_scan += t_scan.ms();
    if (cnt.n_hits >= (int64_t)INT32_MAX) return fail(BN_ERR_OVERFLOW, "more than 2^31 seed hits in one search");
    const int64_t n = cnt.n_hits;
    if (n == 0) return BN_OK;

    t_ext.start();
    CU_TRY(ws.keys_b.reserve((size_t)n));
    CU_TRY(ws.hits_b.reserve((size_t)n));
    {
        // stable radix sort on (diagonal group, global position): radix_sort.cu
        const int end_bit = std::min(64, gbits + grp_bits);
        CU_TRY(ws.cub_temp.reserve(radix_sort_temp_bytes(n)));
        bool in_b = false;
        int64_t launches = 0;
        CU_TRY(radix_sort_hits(ws.keys_a.p, ws.keys_b.p, ws.hits_a.p, ws.hits_b.p, n, end_bit, ws.cub_temp.p, &in_b, &launches, st));
        if (!in_b) { std::swap(ws.keys_a, ws.keys_b); std::swap(ws.hits_a, ws.hits_b); }     // the sorted pairs are the "b" buffers
        if (stats) stats->kernel_launches += launches;
    }
    if (raw_pairs) { t_ext.stop(); if (stats) stats->ms_extend += t_ext.ms(); return BN_OK; }

    CU_TRY(ws.cells.reserve((size_t)std::max<int64_t>(n + 2, serial ? extend_serial_cells(n, Q.batch.container_type == BN_DIAG_HASH, Q.diag_array_length) : 0)));
    CU_TRY(ws.heads.reserve((size_t)n + 1));
    CU_TRY(ws.leaders.reserve((size_t)n + 1));
    CU_TRY(ws.spec.reserve((size_t)n + 1));
    int64_t init_cap = std::max<int64_t>((int64_t)ws.init.cap, std::max<int64_t>(4096, n / 4));
    for (int attempt = 0;; attempt++) {
        CU_TRY(ws.init.reserve((size_t)init_cap));
        init_cap = (int64_t)ws.init.cap;
        CU_TRY(cudaMemsetAsync(ws.counters.p + 2, 0, 4 * sizeof(unsigned long long), st));
        ExtendLaunch e{};
        e.packed = V.d_packed; e.chunks = T.dev.p; e.ranges = T.ranges_dev.p; e.hits = ws.hits_b.p;
        e.cells = reinterpret_cast<int32_t *>(ws.cells.p); e.init = ws.init.p;
        e.counters = ws.counters.p; e.init_capacity = init_cap;
        e.spec = ws.spec.p; e.leaders = ws.leaders.p;
        e.scalar_ok = (Q.direct_ok && !getenv("BN_NO_SCALAR_LEADERS")) ? 1 : 0;
        e.uni_ok = Q.uni_ok; e.uni_x = Q.uni_x; e.uni_cutoff = Q.uni_cutoff; e.uni_reduced = Q.uni_reduced;
        if (serial) CU_TRY(launch_extend_serial(dq, e, ws.keys_b.p, n, gbits, Q.diag_array_length, st));
        else CU_TRY(launch_extend_groups(dq, e, ws.keys_b.p, ws.heads.p, n, gbits, st));
        if (stats) stats->kernel_launches += serial ? 1 : 3;
        CU_TRY(cudaMemcpyAsync(ws.h_counters, ws.counters.p, 8 * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        cnt.n_init = (int64_t)ws.h_counters[2];
        cnt.n_extended = (int64_t)ws.h_counters[3];
        if (cnt.n_init <= init_cap) break;
        if (attempt > 2) return fail(BN_ERR_OVERFLOW, "init-hit buffer overflow");
        init_cap = cnt.n_init + cnt.n_init / 16 + 1024;
    }
    t_ext.stop();
    if (stats) stats->ms_extend += t_ext.ms();
    return BN_OK;
}

static int32_t greedy_xdrop_offset(const BnQueryBatch &b)
{
    int32_t match = b.reward, mismatch = -b.penalty, xd = b.gap_x_dropoff;
    if (match % 2 == 1) { match *= 2; mismatch *= 2; xd *= 2; }
    return (xd + match / 2) / (match + mismatch) + 1;
}

// Tier-1 gapped launch over the init hits in ws.init; their number is read on the device
// (counters[2], capped at max_init), so the call needs no host knowledge of it.
static int enqueue_gapped(Lane &D, Volume &V, Query &Q, ChunkTable &T, int64_t max_init, BnStats *stats,
                          cudaStream_t on_stream = nullptr)
{
    Workspace &ws = D.ws();
    cudaStream_t st = on_stream ? on_stream : D.stream;
    const DevQuery &dq = Q.dev[V.device].view;
    const BnQueryBatch &b = Q.batch;
    CU_TRY(ws.gap_out.reserve((size_t)max_init));
    const bool greedy = b.gap_algo == BN_GAP_GREEDY;
    const int32_t xo = greedy_xdrop_offset(b);
    const int wpb = 4;                                   // greedy: warps per block
    // affine greedy (non-default gap costs with -greedy): thread-per-HSP kernel, rows for max_penalty + 1 distances
    const bool affine = greedy && (b.gap_open != 0 || b.gap_extend != 0);
    const AffineCosts ac = affine_costs(b.reward, b.penalty, b.gap_open, b.gap_extend, b.gap_x_dropoff);
    const int32_t tier = affine ? 128 : (greedy ? 254 : 1024);
    const int64_t per_thread = affine ? affine_scratch_ints(ac, tier)
                               : (greedy ? (2 * (2 * (int64_t)tier + 6) + tier + 1 + xo + 8) : 2 * (int64_t)tier);
    GappedLaunch g{};
    g.packed = V.d_packed; g.chunks = T.dev.p; g.init = ws.init.p; g.n_init = ws.counters.p + 2;
    g.max_init = max_init; g.out = ws.gap_out.p;
    g.scratch_ints_per_thread = per_thread; g.tier_d = tier; g.todo = nullptr; g.n_todo = 0;
    if (greedy && !affine) {
        // two warps per init-HSP (one per direction), rows in shared memory
        const int blocks = (int)std::min<int64_t>((max_init + 1) / 2, 148 * 8);
        g.scratch = nullptr;
        CU_TRY(launch_greedy_warp(dq, g, wpb, blocks, true, st));
    } else if (!greedy) {
        // packed DP, tier 1: one thread per init-HSP, score ring in shared memory
        const int64_t want = (max_init + gapped_threads_per_block() - 1) / gapped_threads_per_block();
        g.scratch = nullptr; g.dp_smem_ring = 1;
        g.dp_max_rows = 256;          // longer alignments go to the warp-parallel kernel (finish_gapped)
        // 16-bit ring cells when every live score of a 256-row extension fits them (gapped_kernel.cu: SmemRing16)
        const int64_t hi = (int64_t)std::max(b.reward, 1) * (g.dp_max_rows + 80), lo = (int64_t)b.gap_x_dropoff + 2 * ((int64_t)b.gap_open + b.gap_extend);
        const bool ring16 = hi < 30000 && lo < 30000 && !getenv("BN_NO_RING16");
        if (ring16) {
            g.dp_smem_ring = 2;
            g.work_counter = ws.counters.p + 7;
            CU_TRY(cudaMemsetAsync(g.work_counter, 0, sizeof(unsigned long long), st));
        }
        g.grid_blocks = (int32_t)std::max<int64_t>(1, std::min<int64_t>(ring16 ? gapped_dp_ring16_blocks() : gapped_dp_smem_blocks(), want));
        CU_TRY(launch_gapped(dq, g, st));
    } else {
        const int64_t threads = std::min<int64_t>(4096, ((max_init + 63) / 64) * 64);
        CU_TRY(ws.scratch.reserve((size_t)(per_thread * threads)));
        g.scratch = ws.scratch.p;
        g.grid_blocks = (int32_t)(threads / gapped_threads_per_block());
        CU_TRY(launch_gapped(dq, g, st));
    }
    if (stats) stats->kernel_launches += 1;
    return BN_OK;
}

// D2H of the init hits and tier-1 results, then tier 2 (worst-case scratch) for the few extensions
// that outgrew tier 1.
static int finish_gapped(Lane &D, Volume &V, Query &Q, ChunkTable &T, int64_t n_init,
                         DevInitHit *&h_init, DevGapResult *&h_gap, BnStats *stats, bool mirrored = false)
{
    Workspace &ws = D.ws();
    cudaStream_t st = D.stream;
    const DevQuery &dq = Q.dev[V.device].view;
    const BnQueryBatch &b = Q.batch;
    CU_TRY(ws.h_init.reserve((size_t)n_init + 1)); CU_TRY(ws.h_gap.reserve((size_t)n_init + 1));
    h_init = ws.h_init.p; h_gap = ws.h_gap.p;
    if (n_init == 0) return BN_OK;
    const bool greedy = b.gap_algo == BN_GAP_GREEDY;
    const int32_t xo = greedy_xdrop_offset(b);
    const int wpb = 4;
    if (!mirrored) {        // the fused pipeline has already written both arrays into the pinned mirrors
        CU_TRY(cudaMemcpyAsync(h_init, ws.init.p, (size_t)n_init * sizeof(DevInitHit), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(h_gap, ws.gap_out.p, (size_t)n_init * sizeof(DevGapResult), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
    }

    std::vector<int32_t> todo;
    if (!greedy) {
        // long alignments (the thread-per-HSP kernel gave up after its row budget): one warp each
        for (int64_t i = 0; i < n_init; i++) if (h_gap[(size_t)i].status == 2) todo.push_back((int32_t)i);
        if (!todo.empty()) {
            CU_TRY(ws.todo.reserve(todo.size()));
            CU_TRY(cudaMemcpyAsync(ws.todo.p, todo.data(), todo.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            GappedLaunch g{};
            g.packed = V.d_packed; g.chunks = T.dev.p; g.init = ws.init.p; g.n_init = ws.counters.p + 2;
            g.max_init = n_init; g.out = ws.gap_out.p;
            g.todo = ws.todo.p; g.n_todo = (int32_t)todo.size();
            if (getenv("BN_WARP_DP")) {         // the lane-per-cell formulation, kept for comparison
                const int wpb_dp = gapped_warp_per_block();
                const int blocks = (int)std::min<int64_t>(((int64_t)todo.size() + wpb_dp - 1) / wpb_dp, 148 * 4);
                CU_TRY(launch_gapped_warp(dq, g, blocks, st));
                if (stats) stats->kernel_launches += 1;
            } else {
                CU_TRY(ws.scratch.reserve(4 * todo.size() + 16));
                CU_TRY(launch_gapped_long(dq, g, reinterpret_cast<int2 *>(ws.scratch.p), st));
                if (stats) stats->kernel_launches += 2;
            }
            CU_TRY(cudaMemcpyAsync(h_gap, ws.gap_out.p, (size_t)n_init * sizeof(DevGapResult), cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            todo.clear();
        }
    }
    for (int64_t i = 0; i < n_init; i++) if (h_gap[(size_t)i].status == 1) todo.push_back((int32_t)i);
    if (!todo.empty()) {
        int32_t max_len = 0;
        for (const auto &c : T.host) max_len = std::max(max_len, c.len);
        int32_t tier;
        int64_t per_thread;
        const bool affine = greedy && (b.gap_open != 0 || b.gap_extend != 0);
        if (affine) {
            tier = std::min(10000, max_len / 2 + 1);
            per_thread = affine_scratch_ints(affine_costs(b.reward, b.penalty, b.gap_open, b.gap_extend, b.gap_x_dropoff), tier);
        } else if (greedy) {
            tier = std::min(10000, max_len / 2 + 1);
            per_thread = 2 * (2 * (int64_t)tier + 6) + tier + 1 + xo + 8;
        } else {
            tier = 256;                                                  // DP ring capacity: power of two >= longest query + 8
            while (tier < Q.max_query_length + 8) tier <<= 1;
            per_thread = 2 * (int64_t)tier;
        }
        const bool warp_greedy = greedy && !affine;
        const int tpb = warp_greedy ? wpb : gapped_threads_per_block();      // workers (warps | threads) per block
        const int hpb = warp_greedy ? wpb / 2 : tpb;                         // init-HSPs in flight per block
        int64_t blocks = std::min<int64_t>(((int64_t)todo.size() + hpb - 1) / hpb, 64);
        if (affine)     // worst-case affine rows are megabytes per thread: bound the scratch to ~2 GB
            blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, ((int64_t)1 << 29) / (per_thread * tpb)));
        CU_TRY(ws.scratch.reserve((size_t)(per_thread * blocks * tpb)));
        CU_TRY(ws.todo.reserve(todo.size()));
        CU_TRY(cudaMemcpyAsync(ws.todo.p, todo.data(), todo.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        GappedLaunch g{};
        g.packed = V.d_packed; g.chunks = T.dev.p; g.init = ws.init.p; g.n_init = ws.counters.p + 2;
        g.max_init = n_init; g.out = ws.gap_out.p;
        g.scratch = ws.scratch.p; g.scratch_ints_per_thread = per_thread; g.tier_d = tier;
        g.todo = ws.todo.p; g.n_todo = (int32_t)todo.size(); g.grid_blocks = (int32_t)blocks;
        if (warp_greedy) CU_TRY(launch_greedy_warp(dq, g, wpb, (int)blocks, false, st));
        else CU_TRY(launch_gapped(dq, g, st));
        if (stats) stats->kernel_launches += 1;
        CU_TRY(cudaMemcpyAsync(h_gap, ws.gap_out.p, (size_t)n_init * sizeof(DevGapResult), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        for (int32_t i : todo)
            if (h_gap[(size_t)i].status != 0) return fail(BN_ERR_OVERFLOW, "gapped extension scratch overflow in tier 2");
    }
    return BN_OK;
}

// The same tiers with hand-over lists built on the device, followed by the triage (triage_kernel.cu): only the
// winners and the undecided losers cross PCIe, the other losers arrive as one count.  For the large general-path
// searches of blastn mode (millions of init-HSPs, a few hundred HSPs).
static int finish_gapped_triaged(Lane &D, Volume &V, Query &Q, ChunkTable &T, int64_t n_init, GpuOut &G, BnStats *stats)
{
    Workspace &ws = D.ws();
    cudaStream_t st = D.stream;
    const DevQuery &dq = Q.dev[V.device].view;
    const BnQueryBatch &b = Q.batch;
    const bool greedy = b.gap_algo == BN_GAP_GREEDY;
    const bool affine = greedy && (b.gap_open != 0 || b.gap_extend != 0);
    const int32_t xo = greedy_xdrop_offset(b);
    const int wpb = 4;
    CU_TRY(ws.todo.reserve((size_t)n_init));
    unsigned long long *tc = ws.counters.p + 8;          // [8] status-2 count, [9] status-1 count, [10] winners, [11] undecided, [12] counted
    auto collect = [&](int32_t want, int slot, int64_t &count) -> int {
        CU_TRY(cudaMemsetAsync(tc + slot, 0, sizeof(unsigned long long), st));
        CU_TRY(launch_collect_status(ws.gap_out.p, ws.counters.p + 2, n_init, want, ws.todo.p, tc + slot, st));
        CU_TRY(cudaMemcpyAsync(ws.h_counters + 8 + slot, tc + slot, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        count = (int64_t)ws.h_counters[8 + slot];
        if (stats) stats->kernel_launches += 1;
        return BN_OK;
    };
    int64_t n2 = 0, n1 = 0;
    int rc;
    if (!greedy) {
        rc = collect(2, 0, n2);
        if (rc) return rc;
        if (n2) {
            // long alignments, in rounds (triage_kernel.cu: LongRounds): per (chunk, context) the best pending one,
            // the ones inside a box made so far are set aside
            const size_t n_cells_lr = T.host.size() * (size_t)b.num_contexts;
            CU_TRY(ws.lr_state.reserve((size_t)n2)); CU_TRY(ws.lr_ctx.reserve((size_t)n2)); CU_TRY(ws.lr_next.reserve((size_t)n2));
            CU_TRY(ws.lr_list.reserve((size_t)n2)); CU_TRY(ws.lr_w.reserve((size_t)n2));
            CU_TRY(ws.lr_best.reserve(n_cells_lr)); CU_TRY(ws.lr_head.reserve(n_cells_lr));
            CU_TRY(ws.scratch.reserve((size_t)(4 * n2 + 16)));
            // the list of long extensions lives in ws.todo; the rounds need it while ws.todo is reused: keep a copy
            CU_TRY(ws.lr_list.reserve((size_t)(2 * n2)));
            int32_t *all_long = ws.lr_list.p + n2;
            CU_TRY(cudaMemcpyAsync(all_long, ws.todo.p, (size_t)n2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
            CU_TRY(cudaMemsetAsync(ws.lr_head.p, 0, n_cells_lr * sizeof(uint32_t), st));
            LongRounds r{};
            r.init = ws.init.p; r.gap = ws.gap_out.p; r.todo = all_long; r.n_todo = (int32_t)n2;
            r.state = ws.lr_state.p; r.ctx_w = ws.lr_ctx.p; r.chain_next = ws.lr_next.p;
            r.best = ws.lr_best.p; r.chain_head = ws.lr_head.p; r.round_list = ws.lr_list.p; r.round_w = ws.lr_w.p;
            r.round_count = tc + 5; r.n_ctx = b.num_contexts; r.min_diag_separation = b.min_diag_separation;
            r.set_aside_all = getenv("BN_LONG_SET_ASIDE_ALL") ? 1 : 0;
            CU_TRY(launch_long_prepare(dq, r, st));
            const bool no_rounds = getenv("BN_NO_LONG_ROUNDS") != nullptr;      // test switch: extend all of them
            for (int round = 0;; round++) {
                int64_t n_round = n2;
                if (no_rounds) {
                    CU_TRY(cudaMemcpyAsync(ws.lr_list.p, all_long, (size_t)n2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
                } else {
                    CU_TRY(cudaMemsetAsync(ws.lr_best.p, 0xFF, n_cells_lr * sizeof(unsigned long long), st));
                    CU_TRY(cudaMemsetAsync(r.round_count, 0, sizeof(unsigned long long), st));
                    CU_TRY(launch_long_select(dq, r, st));
                    CU_TRY(cudaMemcpyAsync(ws.h_counters + 13, r.round_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                    CU_TRY(cudaStreamSynchronize(st));
                    n_round = (int64_t)ws.h_counters[13];
                    if (stats) stats->kernel_launches += 2;
                }
                if (n_round == 0) break;
                GappedLaunch g{};
                g.packed = V.d_packed; g.chunks = T.dev.p; g.init = ws.init.p; g.n_init = ws.counters.p + 2;
                g.max_init = n_init; g.out = ws.gap_out.p;
                g.todo = ws.lr_list.p; g.n_todo = (int32_t)n_round;
                if (getenv("BN_WARP_DP")) {         // the lane-per-cell formulation, kept for comparison
                    const int wpb_dp = gapped_warp_per_block();
                    const int blocks = (int)std::min<int64_t>((n_round + wpb_dp - 1) / wpb_dp, 148 * 4);
                    CU_TRY(launch_gapped_warp(dq, g, blocks, st));
                    if (stats) stats->kernel_launches += 1;
                } else {
                    CU_TRY(launch_gapped_long(dq, g, reinterpret_cast<int2 *>(ws.scratch.p), st));
                    if (stats) stats->kernel_launches += 2;
                }
                if (no_rounds) break;
                CU_TRY(launch_long_commit(dq, r, (int32_t)n_round, st));
                if (stats) stats->kernel_launches += 1;
            }
        }
    }
    // tier 2: worst-case scratch for the few that outgrew a shared-memory ring (their indices are in ws.todo)
    auto run_tier2 = [&](int64_t count) -> int {
        int32_t max_len = 0;
        for (const auto &c : T.host) max_len = std::max(max_len, c.len);
        int32_t tier;
        int64_t per_thread;
        if (affine) {
            tier = std::min(10000, max_len / 2 + 1);
            per_thread = affine_scratch_ints(affine_costs(b.reward, b.penalty, b.gap_open, b.gap_extend, b.gap_x_dropoff), tier);
        } else if (greedy) {
            tier = std::min(10000, max_len / 2 + 1);
            per_thread = 2 * (2 * (int64_t)tier + 6) + tier + 1 + xo + 8;
        } else {
            tier = 256;
            while (tier < Q.max_query_length + 8) tier <<= 1;
            per_thread = 2 * (int64_t)tier;
        }
        const bool warp_greedy = greedy && !affine;
        const int tpb = warp_greedy ? wpb : gapped_threads_per_block();
        const int hpb = warp_greedy ? wpb / 2 : tpb;
        int64_t blocks = std::min<int64_t>((count + hpb - 1) / hpb, 64);
        if (affine) blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, ((int64_t)1 << 29) / (per_thread * tpb)));
        CU_TRY(ws.scratch.reserve((size_t)(per_thread * blocks * tpb)));
        GappedLaunch g{};
        g.packed = V.d_packed; g.chunks = T.dev.p; g.init = ws.init.p; g.n_init = ws.counters.p + 2;
        g.max_init = n_init; g.out = ws.gap_out.p;
        g.scratch = ws.scratch.p; g.scratch_ints_per_thread = per_thread; g.tier_d = tier;
        g.todo = ws.todo.p; g.n_todo = (int32_t)count; g.grid_blocks = (int32_t)blocks;
        if (warp_greedy) CU_TRY(launch_greedy_warp(dq, g, wpb, (int)blocks, false, st));
        else CU_TRY(launch_gapped(dq, g, st));
        if (stats) stats->kernel_launches += 1;
        int64_t left = 0;
        const int r2 = collect(1, 1, left);
        if (r2) return r2;
        if (left) return fail(BN_ERR_OVERFLOW, "gapped extension scratch overflow in tier 2");
        return BN_OK;
    };
    rc = collect(1, 1, n1);
    if (rc) return rc;
    if (n1) { rc = run_tier2(n1); if (rc) return rc; }
    // ---- triage ---------------------------------------------------------------------------------------------
    const size_t n_ctx = (size_t)b.num_contexts, n_cells = T.host.size() * n_ctx;
    int64_t sel_cap = std::max<int64_t>((int64_t)ws.tri_init.cap, std::max<int64_t>(65536, n_init / 8));
    for (int attempt = 0;; attempt++) {
        CU_TRY(ws.tri_ctx.reserve((size_t)n_init));
        CU_TRY(ws.tri_init.reserve((size_t)sel_cap)); CU_TRY(ws.tri_gap.reserve((size_t)sel_cap));
        CU_TRY(ws.tri_sel_ctx.reserve((size_t)sel_cap)); CU_TRY(ws.tri_sel_idx.reserve((size_t)sel_cap));
        sel_cap = (int64_t)std::min(std::min(ws.tri_init.cap, ws.tri_gap.cap), std::min(ws.tri_sel_ctx.cap, ws.tri_sel_idx.cap));
        CU_TRY(ws.tri_table.reserve(n_cells));
        CU_TRY(cudaMemsetAsync(ws.tri_table.p, 0, n_cells * sizeof(uint2), st));
        CU_TRY(cudaMemsetAsync(tc + 2, 0, 3 * sizeof(unsigned long long), st));
        TriageLaunch t{};
        t.init = ws.init.p; t.gap = ws.gap_out.p; t.n_init = ws.counters.p + 2; t.max_init = n_init;
        t.ctx_of = ws.tri_ctx.p; t.sel_init = ws.tri_init.p; t.sel_gap = ws.tri_gap.p; t.sel_ctx = ws.tri_sel_ctx.p;
        t.sel_idx = ws.tri_sel_idx.p;
        t.sel_cap = sel_cap; t.tcount = tc + 2; t.table = ws.tri_table.p; t.n_ctx = (int32_t)n_ctx;
        CU_TRY(launch_triage(dq, t, st));
        if (stats) stats->kernel_launches += 2;
        CU_TRY(cudaMemcpyAsync(ws.h_counters + 10, tc + 2, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        const int64_t n_sel = (int64_t)(ws.h_counters[10] + ws.h_counters[11]);
        if (n_sel <= sel_cap) { G.n_records = n_sel; break; }
        if (attempt > 1) return fail(BN_ERR_OVERFLOW, "triage selection overflow");
        sel_cap = n_sel + n_sel / 16 + 1024;
    }
    CU_TRY(ws.h_init.reserve((size_t)G.n_records + 1)); CU_TRY(ws.h_gap.reserve((size_t)G.n_records + 1));
    G.h_init = ws.h_init.p; G.h_gap = ws.h_gap.p;
    if (G.n_records) {
        CU_TRY(cudaMemcpyAsync(G.h_init, ws.tri_init.p, (size_t)G.n_records * sizeof(DevInitHit), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(G.h_gap, ws.tri_gap.p, (size_t)G.n_records * sizeof(DevGapResult), cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(ws.h_sel_idx.reserve((size_t)G.n_records + 1));
    if (G.n_records)
        CU_TRY(cudaMemcpyAsync(ws.h_sel_idx.p, ws.tri_sel_idx.p, (size_t)G.n_records * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    const int64_t counted = (int64_t)ws.h_counters[12];
    G.counted_losers = counted;
    G.triaged = true;
    if (G.n_records + counted != n_init) return fail(BN_ERR_CUDA, "triage lost init-HSPs");

    // ---- set-aside long extensions the replay needs after all -------------------------------------------------
    // The rounds only PREDICT which long extensions the reference skips as contained.  The chunks that hold a
    // set-aside record are replayed here, exactly as the host phase will replay them; wherever the replay reaches a
    // set-aside record that is not contained, that extension is made now and the chunk is replayed again.
    {
        std::map<int32_t, std::vector<int64_t>> by_chunk;       // chunk -> records, only chunks with a set-aside record
        for (int64_t k = 0; k < G.n_records; k++)
            if (G.h_gap[k].status == 3) by_chunk[G.h_init[k].chunk];
        if (!by_chunk.empty()) {
            for (int64_t k = 0; k < G.n_records; k++) {
                auto it = by_chunk.find(G.h_init[k].chunk);
                if (it != by_chunk.end()) it->second.push_back(k);
            }
            struct Rec { HostInit h; int64_t rec; };
            for (int iteration = 0;; iteration++) {
                std::vector<int64_t> needs;                     // record positions
                for (auto &kv : by_chunk) {
                    std::vector<Rec> recs;
                    recs.reserve(kv.second.size());
                    for (int64_t k : kv.second) {
                        const DevInitHit &h = G.h_init[k];
                        const DevGapResult &g = G.h_gap[k];
                        recs.push_back(Rec{HostInit{h.chunk, h.q_off, h.s_off, h.q_start, h.s_start, h.length, h.score, h.order,
                                                    g.q_start, g.q_stop, g.s_start, g.s_stop, g.score, g.q_seed, g.s_seed, g.status}, k});
                    }
                    std::sort(recs.begin(), recs.end(), [](const Rec &x, const Rec &y) {      // Blast_InitHitListSortByScore + emission order
                        const HostInit &a = x.h, &c = y.h;
                        if (a.score != c.score) return a.score > c.score;
                        if (a.s_start != c.s_start) return a.s_start < c.s_start;
                        if (a.length != c.length) return a.length > c.length;
                        if (a.q_start != c.q_start) return a.q_start < c.q_start;
                        return a.order < c.order;
                    });
                    std::vector<HostInit> inits(recs.size());
                    for (size_t j = 0; j < recs.size(); j++) inits[j] = recs[j].h;
                    std::vector<BnHSP> scratch_out;
                    BnStats scratch_stats{};
                    int64_t needed = -1;
                    replay_gapped(b, T.hchunks[(size_t)kv.first], inits.data(), inits.size(), nullptr, scratch_out, scratch_stats,
                                  Q.ctx_lite.data(), &needed);
                    if (needed >= 0) needs.push_back(recs[(size_t)needed].rec);
                }
                if (needs.empty()) break;
                if (iteration > 10000) return fail(BN_ERR_OVERFLOW, "set-aside extensions do not converge");
                std::vector<int32_t> idx(needs.size());
                for (size_t j = 0; j < needs.size(); j++) idx[j] = ws.h_sel_idx.p[needs[j]];
                CU_TRY(ws.todo.reserve(idx.size()));
                CU_TRY(cudaMemcpyAsync(ws.todo.p, idx.data(), idx.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
                GappedLaunch g{};
                g.packed = V.d_packed; g.chunks = T.dev.p; g.init = ws.init.p; g.n_init = ws.counters.p + 2;
                g.max_init = n_init; g.out = ws.gap_out.p;
                g.todo = ws.todo.p; g.n_todo = (int32_t)idx.size();
                CU_TRY(ws.scratch.reserve(4 * idx.size() + 16));
                CU_TRY(launch_gapped_long(dq, g, reinterpret_cast<int2 *>(ws.scratch.p), st));
                if (stats) stats->kernel_launches += 2;
                int64_t overflowed = 0;
                rc = collect(1, 1, overflowed);                 // a band wider than the ring: tier 2 (its list is rebuilt in ws.todo)
                if (rc) return rc;
                if (overflowed) { rc = run_tier2(overflowed); if (rc) return rc; }
                for (size_t j = 0; j < needs.size(); j++)
                    CU_TRY(cudaMemcpyAsync(&G.h_gap[needs[j]], ws.gap_out.p + idx[j], sizeof(DevGapResult), cudaMemcpyDeviceToHost, st));
                CU_TRY(cudaStreamSynchronize(st));
            }
        }
    }
    return BN_OK;
}

static int run_gapped(Lane &D, Volume &V, Query &Q, ChunkTable &T, int64_t n_init,
                      DevInitHit *&h_init, DevGapResult *&h_gap, BnStats *stats, GpuOut *triage_into = nullptr)
{
    Workspace &ws = D.ws();
    if (n_init == 0) {
        CU_TRY(ws.h_init.reserve(1)); CU_TRY(ws.h_gap.reserve(1));
        h_init = ws.h_init.p; h_gap = ws.h_gap.p;
        return BN_OK;
    }
    Timer t(D.stream, ws, 2);
    t.start();
    int rc = enqueue_gapped(D, V, Q, T, n_init, stats);
    if (rc) return rc;
    if (triage_into) rc = finish_gapped_triaged(D, V, Q, T, n_init, *triage_into, stats);
    else rc = finish_gapped(D, V, Q, T, n_init, h_init, h_gap, stats);
    if (rc) return rc;
    t.stop();
    if (stats) stats->ms_gapped += t.ms();
    return BN_OK;
}

// ------------------------------------------------------------------------------------------------
// Fast path (diagonal HASH container): scan -> device-side bucket grouping -> speculative ungapped
// pass -> bucket replay -> tier-1 gapped extension, all queued back to back; the host learns the
// counts at ONE synchronisation after the gapped kernel.  *redo is set (and nothing else is to be
// trusted) when the device refused the fast path or a buffer was too small: the caller then runs the
// general path.
// ------------------------------------------------------------------------------------------------
// The fused pipeline of one search in two halves, so that a caller can queue the next search's kernels before it
// waits for this one (bn_prelim_search_jobs): fused_enqueue launches scan ... gapped + the result mirror without any
// host synchronisation and records `done`; fused_complete waits for that event and finishes on the host.
struct FusedState {
    int64_t cap = 0, n_limit = 0, init_cap = 0;
    int spec_enabled = 0;
    cudaEvent_t done = nullptr;        // ws.ev[6]: recorded behind the result mirror
};

static int fused_enqueue(Lane &D, Volume &V, Query &Q, ChunkTable &T, FusedState &F)
{
    Workspace &ws = D.ws();
    cudaStream_t st = D.stream;
    const DevQuery &dq = Q.dev[V.device].view;
    CU_TRY(ws.counters.reserve(16));
    if (!ws.h_counters) CU_TRY(cudaMallocHost(&ws.h_counters, 16 * sizeof(unsigned long long)));
    if (T.total_pos >= (int64_t)1 << 32) return fail(BN_ERR_OVERFLOW, "more than 2^32 scan positions in one search");
    const int gbits = bits_for((uint64_t)std::max<int64_t>(T.total_pos, 1));
    const int nb = group_sort_buckets();

    int64_t cap = std::max<int64_t>((int64_t)ws.hits_a.cap, std::max<int64_t>(1 << 16, T.total_pos / 16));
    CU_TRY(ws.hits_a.reserve((size_t)cap)); CU_TRY(ws.keys_a.reserve((size_t)cap));
    cap = (int64_t)std::min(ws.hits_a.cap, ws.keys_a.cap);
    const int64_t n_limit = std::min<int64_t>(cap, (int64_t)1 << 18);
    CU_TRY(ws.hits_b.reserve((size_t)n_limit)); CU_TRY(ws.keys_b.reserve((size_t)n_limit));
    CU_TRY(ws.keys_tmp.reserve((size_t)nb * (size_t)group_sort_bucket_cap()));
    CU_TRY(ws.cells.reserve((size_t)n_limit + 2));
    CU_TRY(ws.heads.reserve((size_t)nb + 1));
    CU_TRY(ws.leaders.reserve((size_t)n_limit + 1));
    CU_TRY(ws.spec.reserve((size_t)n_limit + 1));
    CU_TRY(ws.buckets.reserve((size_t)nb));
    CU_TRY(ws.init.reserve((size_t)std::max<int64_t>(4096, n_limit / 4)));
    const int64_t init_cap = (int64_t)ws.init.cap;
    F.cap = cap; F.n_limit = n_limit; F.init_cap = init_cap;

    Timer t_scan(st, ws, 0), t_ext(st, ws, 1), t_gap(st, ws, 2);
    CU_TRY(cudaMemsetAsync(ws.counters.p, 0, 8 * sizeof(unsigned long long), st));
    CU_TRY(cudaMemsetAsync(ws.buckets.p, 0, (size_t)nb * sizeof(uint32_t), st));
    ScanLaunch s{};
    s.packed = V.d_packed; s.chunks = T.scan_units(); s.n_chunks = T.n_scan_units();
    s.total_pos = T.total_pos; s.hits = ws.hits_a.p; s.keys = ws.keys_a.p;
    s.counters = ws.counters.p; s.capacity = cap; s.block_chunk = T.block_chunk.p; s.block_desc = T.block_desc.p;
    s.raw_pairs = 0; s.gbits = gbits; s.diag_array_length = Q.diag_array_length;
    s.tile_cap = scan_tile_cap(Q.batch.scan_step, Q.batch.word_length);
    s.bucket_count = ws.buckets.p; s.bucket_keys = ws.keys_tmp.p; s.bucket_cap = group_sort_bucket_cap();
    set_direct_filter(s, Q, T, false);
    t_scan.start();
    CU_TRY(launch_scan(dq, s, st));
    t_scan.stop();
    // everything behind the scan runs on the lane's high-priority stream
    CU_TRY(cudaEventRecord(D.scan_ev, st));
    st = D.tail_stream;
    CU_TRY(cudaStreamWaitEvent(st, D.scan_ev, 0));
    t_ext.st = st; t_gap.st = st;

    t_ext.start();
    BucketLaunch L{};
    L.hits_in = ws.hits_a.p; L.bucket_count = ws.buckets.p;
    L.keys_tmp = ws.keys_tmp.p; L.hits_out = ws.hits_b.p; L.keys_out = ws.keys_b.p;
    L.heads = ws.heads.p; L.leaders = ws.leaders.p; L.spec = ws.spec.p; L.counters = ws.counters.p;
    L.n_limit = n_limit; L.gbits = gbits; L.spec_enabled = Q.batch.window_size > 0 ? 0 : 1;
    F.spec_enabled = L.spec_enabled;
    CU_TRY(launch_bucket_group(L, st));
    ExtendLaunch e{};
    e.packed = V.d_packed; e.chunks = T.dev.p; e.ranges = T.ranges_dev.p; e.hits = ws.hits_b.p;
    e.cells = reinterpret_cast<int32_t *>(ws.cells.p); e.init = ws.init.p;
    e.counters = ws.counters.p; e.init_capacity = init_cap;
    e.spec = ws.spec.p; e.leaders = ws.leaders.p; e.n_from_device = 1;
    e.scalar_ok = (Q.direct_ok && !getenv("BN_NO_SCALAR_LEADERS")) ? 1 : 0;
    e.uni_ok = Q.uni_ok; e.uni_x = Q.uni_x; e.uni_cutoff = Q.uni_cutoff; e.uni_reduced = Q.uni_reduced;
    CU_TRY(launch_extend_grouped(dq, e, ws.keys_b.p, ws.heads.p, gbits, st));
    t_ext.stop();

    t_gap.start();
    CU_TRY(ws.h_init.reserve((size_t)init_cap + 1)); CU_TRY(ws.h_gap.reserve((size_t)init_cap + 1));
    int rc = enqueue_gapped(D, V, Q, T, init_cap, nullptr, st);
    if (rc) return rc;
    // counters + results into the pinned mirrors by one kernel, then the only synchronisation of the step
    CU_TRY(launch_mirror_results(ws.init.p, ws.gap_out.p, ws.counters.p, init_cap, ws.h_init.p, ws.h_gap.p,
                                 ws.h_counters, st));
    t_gap.stop();
    if (!ws.ev[6]) CU_TRY(cudaEventCreateWithFlags(&ws.ev[6], cudaEventDisableTiming));
    F.done = ws.ev[6];
    CU_TRY(cudaEventRecord(F.done, st));
    return BN_OK;
}

static int fused_complete(Lane &D, Volume &V, Query &Q, ChunkTable &T, const FusedState &F, StageCounts &cnt, BnStats &stats,
                          DevInitHit *&h_init, DevGapResult *&h_gap, bool *redo)
{
    *redo = false;
    Workspace &ws = D.ws();
    CU_TRY(cudaEventSynchronize(F.done));
    cnt.n_hits = (int64_t)ws.h_counters[0];
    cnt.lookup_hits = (int64_t)ws.h_counters[1];
    cnt.n_init = (int64_t)ws.h_counters[2];
    cnt.n_extended = (int64_t)ws.h_counters[3];
    if (cnt.n_hits > F.cap) {                    // scan output overflowed: grow for the general path's retry
        CU_TRY(ws.hits_a.reserve((size_t)(cnt.n_hits + cnt.n_hits / 16 + 1024)));
        CU_TRY(ws.keys_a.reserve((size_t)(cnt.n_hits + cnt.n_hits / 16 + 1024)));
    }
    if (cnt.n_hits > F.cap || ws.h_counters[6] || cnt.n_init > F.init_cap) { *redo = true; return BN_OK; }
    Timer t_scan(D.stream, ws, 0), t_ext(D.stream, ws, 1), t_gap(D.stream, ws, 2);
    const double ms_scan = t_scan.ms(), ms_ext = t_ext.ms(), ms_gap = t_gap.ms();      // before tier 2 re-uses the stream
    int rc = finish_gapped(D, V, Q, T, cnt.n_init, h_init, h_gap, &stats, true);
    if (rc) return rc;
    stats.kernel_launches += 1 + 1 + (F.spec_enabled ? 2 : 1) + 1 + 1;      // scan, grouping, extension, gapped, result mirror
    stats.ms_scan += ms_scan; stats.ms_extend += ms_ext; stats.ms_gapped += ms_gap;
    return BN_OK;
}

template <typename T>
static T *to_malloc(const std::vector<T> &v)
{
    if (v.empty()) return nullptr;
    T *p = (T *)malloc(v.size() * sizeof(T));
    if (p) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

// The GPU side of one search in two halves (see FusedState): search_gpu_begin builds / finds the chunk table and, for
// the searches the fused pipeline serves, queues all of its kernels; search_gpu_end waits for them and runs whatever
// is left (the general path from the start when the fused one did not apply or overflowed).
struct SearchPending {
    bool fused = false;
    bool allow_triage = false;
    FusedState F;
    double tw0 = 0;
};

// allow_triage: the caller's host phase takes no taps and its low_score bounds cannot move during the search
static int search_gpu_begin(Lane &D, Volume &V, Query &Q, int32_t oid_begin, int32_t oid_end, BnResults *out, GpuOut &G,
                            bool allow_triage, SearchPending &P)
{
    memset(out, 0, sizeof *out);
    G.t0 = now_ms();
    G.oid_begin = oid_begin; G.oid_end = oid_end;
    P.allow_triage = allow_triage;
    CU_TRY(cudaSetDevice(D.id));
    if (!Q.dev[V.device].ready) return fail(BN_ERR_INVALID, "query batch is not loaded on the volume's device");
    std::shared_ptr<ChunkTable> &T = G.T;
    int rc = build_chunk_table(V, Q, oid_begin, oid_end, D, &T);
    if (rc) return rc;
    out->stats.subject_bases_scanned = T->total_bases;
    if (V.ready) CU_TRY(cudaStreamWaitEvent(D.stream, V.ready, 0));
    P.tw0 = now_ms();
    const bool general = Q.batch.container_type != BN_DIAG_HASH || Q.fast_path_refused || T->total_pos <= 0 ||
                         serial_replay(Q.batch) || getenv("BN_FORCE_GENERAL") != nullptr;
    P.fused = !general;
    if (P.fused) return fused_enqueue(D, V, Q, *T, P.F);
    return BN_OK;
}

static int search_gpu_end(Lane &D, Volume &V, Query &Q, BnResults *out, GpuOut &G, SearchPending &P)
{
    CU_TRY(cudaSetDevice(D.id));
    std::shared_ptr<ChunkTable> &T = G.T;
    BnStats &stats = out->stats;
    StageCounts &cnt = G.cnt;
    const double tw0 = P.tw0;
    double tw1 = tw0;
    int rc;
    bool general = !P.fused;
    if (P.fused) {
        bool redo = false;
        rc = fused_complete(D, V, Q, *T, P.F, cnt, stats, G.h_init, G.h_gap, &redo);
        if (rc) return rc;
        if (redo) { general = true; Q.fast_path_refused = true; cnt = StageCounts{}; }
        tw1 = now_ms();
    }
    if (general) {
        rc = run_word_finder(D, V, Q, *T, false, cnt, &stats);
        if (rc) return rc;
        tw1 = now_ms();
        // triage on the device when the result set is large and a (chunk, context) table is affordable
        const bool no_triage = getenv("BN_NO_TRIAGE") != nullptr;             // test switches, read per search
        const int64_t triage_min = getenv("BN_TRIAGE_MIN") ? atoll(getenv("BN_TRIAGE_MIN")) : 100000;
        bool strands_are_contexts = true;
        for (int32_t c = 0; c < Q.batch.num_contexts && strands_are_contexts; c++) strands_are_contexts = Q.ctx_lite[(size_t)c].strand_ctx == c;
        const bool triage = P.allow_triage && !no_triage && cnt.n_init >= triage_min && strands_are_contexts &&
                            (int64_t)T->host.size() * Q.batch.num_contexts <= ((int64_t)1 << 22);
        rc = run_gapped(D, V, Q, *T, cnt.n_init, G.h_init, G.h_gap, &stats, triage ? &G : nullptr);
        if (rc) return rc;
    }
    if (!G.triaged) G.n_records = cnt.n_init;
    const double tw2 = now_ms();
    G.t_table = tw0 - G.t0; G.t_wf = tw1 - tw0; G.t_gap = tw2 - tw1;
    stats.lookup_hits = cnt.lookup_hits;
    stats.init_extends = cnt.n_init;        // the reference counts the extensions that were saved (hit_ready, core/na_ungapped.c:1000-1004)
    stats.good_init_extends = cnt.n_init;
    return BN_OK;
}

static int search_gpu_phase(Lane &D, Volume &V, Query &Q, int32_t oid_begin, int32_t oid_end, BnResults *out, GpuOut &G,
                            bool allow_triage = false)
{
    SearchPending P;
    int rc = search_gpu_begin(D, V, Q, oid_begin, oid_end, out, G, allow_triage, P);
    if (rc) return rc;
    return search_gpu_end(D, V, Q, out, G, P);
}

// Host replay of one search (containment filter, per-chunk list post-processing, chunk merge, E-values,
// low_score feedback).  Touches no device state: safe on a worker thread while the device runs the next batch.
// One search's place in a larger one (bn_prelim_search_volumes): the hit lists behind low_score are shared by all
// volumes, and the volume's sequences are numbered after those of the volumes before it.
struct HostShared {
    LowScoreTracker *tracker = nullptr;
    bool bounds_fixed = false;
    int32_t oid_base = 0;
};

static int search_host_phase(const Query &Q, const GpuOut &G, int taps, BnResults *out, const HostShared *sh = nullptr)
{
    const std::shared_ptr<ChunkTable> &T = G.T;
    const StageCounts &cnt = G.cnt;
    const DevInitHit *h_init = G.h_init;
    const DevGapResult *h_gap = G.h_gap;
    const int32_t oid_begin = G.oid_begin, oid_end = G.oid_end;
    const double t0 = G.t0, tw0 = G.t0 + G.t_table, tw1 = tw0 + G.t_wf, tw2 = tw1 + G.t_gap;
    BnStats &stats = out->stats;
    // ---- host replay -------------------------------------------------------------------------
    const double th0 = now_ms();
    static const bool trace = getenv("BN_TRACE") != nullptr;
    double t_sort = 0, t_replay = 0, t_merge = 0, t_eval = 0, t_track = 0;
    const BnQueryBatch &b = Q.batch;
    // init hits grouped by chunk, each group in the reference's order.  groups[k] = {chunk, begin, end} into
    // `inits`; a counting sort when the hits are many compared with the chunks, else a sort of the hits
    // (an nt-like volume has a million chunks and a few thousand init-HSPs).
    const size_t n_chunks = T->hchunks.size();
    const size_t n_in = (size_t)G.n_records;        // all init-HSPs, or what the device-side triage selected
    std::vector<HostInit> inits(n_in);
    struct Group { size_t chunk, lo, hi; };
    std::vector<Group> groups;
    auto host_init = [&](size_t i) {
        const DevInitHit &h = h_init[i];
        const DevGapResult &g = h_gap[i];
        return HostInit{h.chunk, h.q_off, h.s_off, h.q_start, h.s_start, h.length, h.score, h.order,
                        g.q_start, g.q_stop, g.s_start, g.s_stop, g.score, g.q_seed, g.s_seed, g.status};
    };
    // Small result sets (the usual case) take ONE sort of packed integer keys {chunk, score desc, s_start,
    // length desc, q_start, emission order} - grouping and the per-chunk Blast_InitHitListSortByScore order
    // (core/blast_extend.c:274-296) at once - and a gather; large ones are grouped by a counting sort and
    // sorted per chunk by the worker threads below.
    bool presorted = false;
    if (n_in < 16384) {
        // the pinned result mirrors were just written by the copy engine: read them once, front to back
        std::vector<SortKey> keys(n_in);
        std::vector<HostInit> all(n_in);
        for (size_t i = 0; i < n_in; i++) {
            all[i] = host_init(i);
            const HostInit &h = all[i];
            keys[i] = SortKey{((uint64_t)(uint32_t)h.chunk << 32) | (uint32_t)(INT32_MAX - h.score),
                              ((uint64_t)(uint32_t)h.s_start << 32) | (uint32_t)(INT32_MAX - h.length),
                              ((uint64_t)(uint32_t)h.q_start << 32) | h.order, (uint32_t)i, 0u};
        }
        const double tr1 = now_ms();
        sort_keys(keys);
        const double tr2 = now_ms();
        if (trace) fprintf(stderr, "[bn] host group: alloc+read %.3f sort %.3f ms\n", tr1 - th0, tr2 - tr1);
        for (size_t k = 0; k < n_in; k++) {
            inits[k] = all[(size_t)keys[k].idx];
            const size_t c = (size_t)(keys[k].k0 >> 32);
            if (groups.empty() || groups.back().chunk != c) groups.push_back(Group{c, k, k + 1});
            else groups.back().hi = k + 1;
        }
        presorted = true;
    } else if (n_in * 4 >= n_chunks) {
        std::vector<size_t> group_begin(n_chunks + 1, 0);
        for (size_t i = 0; i < n_in; i++) ++group_begin[(size_t)h_init[i].chunk + 1];
        for (size_t c = 0; c < n_chunks; c++) group_begin[c + 1] += group_begin[c];
        std::vector<size_t> cursor(group_begin.begin(), group_begin.end() - 1);
        for (size_t i = 0; i < n_in; i++) inits[cursor[(size_t)h_init[i].chunk]++] = host_init(i);
        for (size_t c = 0; c < n_chunks; c++)
            if (group_begin[c + 1] > group_begin[c]) groups.push_back(Group{c, group_begin[c], group_begin[c + 1]});
    } else {
        std::vector<uint64_t> order(n_in);
        for (size_t i = 0; i < n_in; i++) order[i] = ((uint64_t)(uint32_t)h_init[i].chunk << 32) | (uint64_t)i;
        std::sort(order.begin(), order.end());
        for (size_t k = 0; k < n_in; k++) {
            inits[k] = host_init((size_t)(order[k] & 0xFFFFFFFFull));
            const size_t c = (size_t)(order[k] >> 32);
            if (groups.empty() || groups.back().chunk != c) groups.push_back(Group{c, k, k + 1});
            else groups.back().hi = k + 1;
        }
    }

    const double t_grouped = now_ms();
    std::vector<BnHSP> final_hsps, gapped_tap;
    std::vector<BnInitHit> init_tap;
    LowScoreTracker own_tracker(b);
    LowScoreTracker &tracker = sh ? *sh->tracker : own_tracker;
    const int32_t oid_base = sh ? sh->oid_base : 0;
    struct GroupOut { std::vector<BnHSP> fresh, tap; BnStats stats{}; double t_sort = 0, t_replay = 0, t_finish = 0; };
    std::vector<GroupOut> gout(groups.size());
    // one group: sort, containment replay, per-chunk list post-processing
    auto do_group = [&](size_t gi) {
        const size_t c = groups[gi].chunk, lo = groups[gi].lo, hi = groups[gi].hi;
        GroupOut &o = gout[gi];
        const double ga = now_ms();
        if (!presorted) sort_chunk_init_hits(inits.data() + lo, inits.data() + hi);
        const double gb = now_ms();
        replay_gapped(b, T->hchunks[c], inits.data() + lo, hi - lo, tracker.low_score(), o.fresh, o.stats, Q.ctx_lite.data());
        const double gc = now_ms();
        if (taps & BN_TAP_GAPPED) o.tap = o.fresh;
        finish_chunk_list(b, o.fresh);
        o.t_sort = gb - ga; o.t_replay = gc - gb; o.t_finish = now_ms() - gc;
    };
    // The per-chunk work only meets other chunks through hit_params->low_score.  When no bound can move
    // during this search (fewer subjects than a hit list holds) the chunks are independent and, for large
    // result sets (short-read batches), are replayed by a few host threads.
    const bool bounds_fixed = sh ? sh->bounds_fixed : tracker.bounds_stay_zero((int64_t)oid_end - oid_begin);
    const bool parallel = n_in >= 16384 && groups.size() >= 2 && bounds_fixed;
    // losers the device counted instead of shipping them: each is an extension the reference makes and drops
    stats.gap_extensions += G.counted_losers;
    // runs of consecutive groups (chunks) of one subject: a subject's chunk lists are merged in order, subjects are
    // independent of each other
    struct Run { size_t lo, hi; };
    std::vector<Run> runs;
    for (size_t gi = 0; gi < groups.size(); gi++) {
        if (!runs.empty() && T->hchunks[groups[gi].chunk].oid == T->hchunks[groups[runs.back().lo].chunk].oid) runs.back().hi = gi + 1;
        else runs.push_back(Run{gi, gi + 1});
    }
    std::vector<std::vector<BnHSP>> run_out(parallel ? runs.size() : 0);
    // the chunk lists of one subject -> its final list: absolute coordinates, chunk merge, E-values, reap
    auto finish_run = [&](const Run &r, std::vector<BnHSP> &list) {
        for (size_t gi = r.lo; gi < r.hi; gi++) {
            const HostChunk &ch = T->hchunks[groups[gi].chunk];
            GroupOut &o = gout[gi];
            for (auto &h : o.fresh) { h.s_off += ch.chunk_off; h.s_end += ch.chunk_off; h.s_gapped_start += ch.chunk_off; }
            merge_chunk_lists(list, o.fresh, ch.chunk_off, ch.chunk_off == 0 ? 0 : BN_DBSEQ_CHUNK_OVERLAP);
            std::vector<BnHSP>().swap(o.fresh);
        }
        evalues_and_reap(b, list);
        if (oid_base) for (auto &h : list) h.oid += oid_base;
    };
    if (parallel) {
        // Chunks are replayed independently of each other (a 107 Mb subject is 22 of them), then every subject's chunk lists
        // are merged and evaluated: workers take chunks first, largest first, and a subject's merge goes to the worker that
        // finishes its last chunk — replay, list post-processing, merge and E-values all run in parallel, and the
        // parallelism is the number of chunks, not of subjects (C4: 150 against 7).
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const size_t n_threads = std::min<size_t>(std::min<size_t>(groups.size(), hw), 16);
        std::vector<uint32_t> order(groups.size()), run_of(groups.size());
        for (size_t ri = 0; ri < runs.size(); ri++)
            for (size_t gi = runs[ri].lo; gi < runs[ri].hi; gi++) run_of[gi] = (uint32_t)ri;
        for (size_t gi = 0; gi < groups.size(); gi++) order[gi] = (uint32_t)gi;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            const size_t na = groups[a].hi - groups[a].lo, nb = groups[b].hi - groups[b].lo;
            return na != nb ? na > nb : a < b;
        });
        std::vector<std::atomic<uint32_t>> left(runs.size());
        for (size_t ri = 0; ri < runs.size(); ri++) left[ri].store((uint32_t)(runs[ri].hi - runs[ri].lo));
        const bool finish_here = !(taps & (BN_TAP_INIT | BN_TAP_GAPPED));
        std::atomic<size_t> next{0};
        std::vector<std::thread> pool;
        auto worker = [&]() {
            for (size_t k; (k = next.fetch_add(1)) < order.size();) {
                const size_t gi = order[k], ri = run_of[gi];
                do_group(gi);
                if (left[ri].fetch_sub(1, std::memory_order_acq_rel) == 1 && finish_here) finish_run(runs[ri], run_out[ri]);
            }
        };
        for (size_t t = 1; t < n_threads; t++) pool.emplace_back(worker);
        worker();
        for (auto &th : pool) th.join();
    }
    t_sort = now_ms() - th0;

    if (parallel) {
        size_t total = 0;
        for (const auto &l : run_out) total += l.size();
        final_hsps.reserve(total);
    }
    for (size_t ri = 0; ri < runs.size(); ri++) {
        std::vector<BnHSP> serial_list;
        for (size_t gi = runs[ri].lo; gi < runs[ri].hi; gi++) {
            const size_t c = groups[gi].chunk, lo = groups[gi].lo, hi = groups[gi].hi;
            const HostChunk &ch = T->hchunks[c];
            double ta = now_ms();
            if (!parallel) do_group(gi);                  // in subject order: the bounds may move between subjects
            GroupOut &o = gout[gi];
            stats.gap_extensions += o.stats.gap_extensions;
            t_replay += now_ms() - ta;
            if (taps & BN_TAP_INIT)
                for (size_t k = lo; k < hi; k++)
                    init_tap.push_back(BnInitHit{ch.oid + oid_base, ch.chunk_off, inits[k].q_off, inits[k].s_off,
                                                 inits[k].q_start, inits[k].s_start, inits[k].length,
                                                 inits[k].score});
            if (taps & BN_TAP_GAPPED) {
                if (oid_base) for (auto &h : o.tap) h.oid += oid_base;
                gapped_tap.insert(gapped_tap.end(), o.tap.begin(), o.tap.end());
            }
        }
        double ta = now_ms();
        const bool done_by_worker = parallel && !(taps & (BN_TAP_INIT | BN_TAP_GAPPED));
        std::vector<BnHSP> &list = done_by_worker ? run_out[ri] : serial_list;
        if (!done_by_worker) finish_run(runs[ri], list);
        t_merge += now_ms() - ta; ta = now_ms();
        if (!list.empty()) {
            stats.good_extensions += (int64_t)list.size();
            final_hsps.insert(final_hsps.end(), list.begin(), list.end());
            if (!bounds_fixed) tracker.subject_done(b, list);      // the hit lists only matter when a bound can move
        }
        std::vector<BnHSP>().swap(list);
        t_track += now_ms() - ta;
    }
    stats.ms_host = now_ms() - th0;
    if (trace)
    {
        double gs = 0, gr = 0, gf = 0;
        for (const auto &o : gout) { gs += o.t_sort; gr += o.t_replay; gf += o.t_finish; }
        fprintf(stderr, "[bn] host: grouping %.3f ms, %u hardware threads\n", t_grouped - th0, std::thread::hardware_concurrency());
        fprintf(stderr, "[bn] wall: table %.3f word-finder %.3f gapped %.3f | host %.3f ms: group(+parallel) %.3f [sum over chunks: sort %.3f replay %.3f finish %.3f] serial-groups %.3f merge %.3f eval %.3f track %.3f tail %.3f (n_hits %lld n_init %lld)\n",
                tw0 - t0, tw1 - tw0, tw2 - tw1, stats.ms_host, t_sort, gs, gr, gf, t_replay, t_merge, t_eval, t_track,
                stats.ms_host - (t_sort + t_replay + t_merge + t_eval + t_track),
                (long long)cnt.n_hits, (long long)cnt.n_init);
    }

    out->n_hsps = (int64_t)final_hsps.size(); out->hsps = to_malloc(final_hsps);
    out->n_init = (int64_t)init_tap.size();   out->init = to_malloc(init_tap);
    out->n_gapped = (int64_t)gapped_tap.size(); out->gapped = to_malloc(gapped_tap);
    stats.ms_total = now_ms() - t0;
    return BN_OK;
}

static bool triage_allowed(const Query &Q, int64_t n_subjects, int taps)
{
    return taps == 0 && LowScoreTracker(Q.batch).bounds_stay_zero(n_subjects);
}

static int prelim_search_locked(Lane &D, Volume &V, Query &Q, int32_t oid_begin, int32_t oid_end,
                                int taps, BnResults *out)
{
    GpuOut G;
    int rc = search_gpu_phase(D, V, Q, oid_begin, oid_end, out, G, triage_allowed(Q, (int64_t)oid_end - oid_begin, taps));
    if (rc) return rc;
    return search_host_phase(Q, G, taps, out);
}

}  // namespace bn

using namespace bn;

// ================================================================================================
//                                             C ABI
// ================================================================================================
// Greedy traceback batch (BLAST_GreedyGappedAlignment with do_traceback): one thread per alignment with a private
// arena; alignments whose rows do not fit are retried with 16x larger arenas (fewer threads).
static int traceback_greedy_host(Lane &Dv, Volume &V, Query &Q, int32_t x_dropoff, const BnTracebackItem *items, int64_t n_items,
                                 const std::vector<DevTracebackItem> &up, BnTracebackResult **results, BnEditOp **ops, int64_t *n_ops)
{
    cudaStream_t st = Dv.stream;
    const DevQuery &dq = Q.dev[V.device].view;
    DevTracebackItem *d_items = nullptr;
    DevTracebackDir *d_out = nullptr;
    unsigned long long *d_cnt = nullptr;
    uint8_t *d_arena = nullptr, *d_todo = nullptr;
    int2 *d_ops = nullptr;
    auto release = [&]() {
        if (d_items) cudaFreeAsync(d_items, st);
        if (d_out) cudaFreeAsync(d_out, st);
        if (d_cnt) cudaFreeAsync(d_cnt, st);
        if (d_arena) cudaFreeAsync(d_arena, st);
        if (d_todo) cudaFreeAsync(d_todo, st);
        if (d_ops) cudaFreeAsync(d_ops, st);
    };
    struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};
    std::vector<DevTracebackDir> pass((size_t)(2 * n_items));
    std::vector<uint8_t> todo((size_t)n_items, 1);
    std::vector<BnEditOp> out_ops;
    std::vector<BnTracebackResult> res((size_t)n_items);
    long long rows = 0;
    for (int64_t i = 0; i < n_items; i++) rows += Q.batch.contexts[items[i].context].query_length;
    long long ops_cap = rows / 4 + 64 * n_items;
    CU_TRY(cudaMallocAsync((void **)&d_items, up.size() * sizeof(DevTracebackItem), st));
    CU_TRY(cudaMallocAsync((void **)&d_out, pass.size() * sizeof(DevTracebackDir), st));
    CU_TRY(cudaMallocAsync((void **)&d_cnt, 2 * sizeof(unsigned long long), st));
    CU_TRY(cudaMallocAsync((void **)&d_todo, (size_t)n_items, st));
    CU_TRY(cudaMemcpyAsync(d_items, up.data(), up.size() * sizeof(DevTracebackItem), cudaMemcpyHostToDevice, st));
    long long per_thread = 256ll << 10;
    int64_t left = n_items;
    for (int attempt = 0; left > 0; attempt++) {
        // 256 KB, 4 MB, 64 MB, 1 GB per alignment in flight; never more than the 4 GB budget in total
        if (attempt == 4) return fail(BN_ERR_OVERFLOW, "bn_gapped_traceback: greedy traceback scratch exhausted");
        const long long budget = 4ll << 30;
        // one WARP per alignment (traceback_greedy_warp_kernel, 4 warps per block), each with a private arena
        // affine costs (BLAST_AffineGreedyAlign's own body): one thread per alignment, 32 per block
        const bool affine = Q.batch.gap_open != 0 || Q.batch.gap_extend != 0;
        const int per_block = affine ? (per_thread >= (64ll << 20) ? 1 : 32) : 4;
        int64_t threads = std::min<int64_t>(std::min<int64_t>(left, affine ? 8192 : 148 * 16), std::max<long long>(budget / per_thread, 1));
        const int blocks = (int)((threads + per_block - 1) / per_block);
        threads = (int64_t)blocks * per_block;
        const long long arena_bytes = threads * per_thread;
        if (d_arena) { cudaFreeAsync(d_arena, st); d_arena = nullptr; }
        if (d_ops) { cudaFreeAsync(d_ops, st); d_ops = nullptr; }
        CU_TRY(cudaMallocAsync((void **)&d_arena, (size_t)arena_bytes, st));
        CU_TRY(cudaMallocAsync((void **)&d_ops, (size_t)ops_cap * sizeof(int2), st));
        CU_TRY(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), st));
        CU_TRY(cudaMemcpyAsync(d_todo, todo.data(), todo.size(), cudaMemcpyHostToDevice, st));
        TracebackLaunch L{};
        L.packed = V.d_packed; L.items = d_items; L.n = n_items; L.x_dropoff = x_dropoff; L.amb_runs = V.d_amb;
        L.arena = d_arena; L.arena_bytes = arena_bytes; L.arena_used = d_cnt;
        L.ops = d_ops; L.ops_cap = ops_cap; L.ops_used = d_cnt + 1; L.out = d_out; L.todo = d_todo;
        CU_TRY(affine ? launch_traceback_greedy_affine(dq, L, blocks, per_block, st) : launch_traceback_greedy_warp(dq, L, blocks, st));
        unsigned long long used[2] = {0, 0};
        CU_TRY(cudaMemcpyAsync(pass.data(), d_out, pass.size() * sizeof(DevTracebackDir), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(used, d_cnt, sizeof used, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        std::vector<int2> h_ops((size_t)std::min<unsigned long long>(used[1], (unsigned long long)ops_cap));
        if (!h_ops.empty()) {
            CU_TRY(cudaMemcpyAsync(h_ops.data(), d_ops, h_ops.size() * sizeof(int2), cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
        }
        bool grow_arena = false, grow_ops = false;
        for (int64_t i = 0; i < n_items; i++) {
            if (!todo[(size_t)i]) continue;
            const DevTracebackDir &l = pass[(size_t)(2 * i)], &r = pass[(size_t)(2 * i + 1)];
            if (l.status == 3) { grow_arena = true; continue; }
            if (l.status == 4) { grow_ops = true; continue; }
            const BnTracebackItem &t = items[i];
            BnTracebackResult &o = res[(size_t)i];
            o.score = l.score;
            o.query_start = t.q_start - l.a_off; o.query_stop = t.q_start + r.a_off;
            o.subject_start = t.s_start - l.b_off; o.subject_stop = t.s_start + r.b_off;
            o.esp_off = (int64_t)out_ops.size(); o.esp_n = l.n_ops; o.status = 0; o.pad = 0;
            for (int32_t k = 0; k < l.n_ops; k++) {
                const int2 e = h_ops[(size_t)(l.ops_off + k)];
                out_ops.push_back(BnEditOp{e.x, e.y});
            }
            todo[(size_t)i] = 0;
            left--;
        }
        if (grow_arena) per_thread *= 16;
        if (grow_ops) ops_cap *= 4;
    }
    *results = to_malloc(res);
    *ops = to_malloc(out_ops);
    *n_ops = (int64_t)out_ops.size();
    if (!*results || (!out_ops.empty() && !*ops)) return fail(BN_ERR_MEMORY, "bn_gapped_traceback: out of memory");
    return BN_OK;
}

extern "C" {

const char *bn_last_error(void) { return g_err.c_str(); }
const char *bn_version(void) { return "gblastn_b200 0.1.0 (sm_100a)"; }

int bn_init(int n_gpu, const int *device_ids)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_inited) return BN_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(BN_ERR_NO_DEVICE, std::string("no usable CUDA device: ") +
                                          (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    std::vector<int> ids;
    if (n_gpu <= 0 || !device_ids) { for (int i = 0; i < (n_gpu > 0 ? std::min(n_gpu, count) : count); i++) ids.push_back(i); }
    else for (int i = 0; i < n_gpu; i++) ids.push_back(device_ids[i]);
    // lanes per device: concurrent callers beyond this number wait for a lane (BN_LANES, default 6)
    int n_lanes = 6;
    if (const char *e = getenv("BN_LANES")) n_lanes = std::max(4, std::min(16, atoi(e)));      // the job pipeline needs three + one for its traceback stage
    for (int id : ids) {
        if (id < 0 || id >= count) return fail(BN_ERR_INVALID, "device id out of range");
        auto d = std::make_unique<Gpu>();
        d->id = id;
        CU_TRY(cudaSetDevice(id));
        for (int k = 0; k < n_lanes; k++) {
            auto l = std::make_unique<Lane>();
            l->id = id; l->index = k;
            CU_TRY(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking));
            {
                int lo = 0, hi = 0;
                CU_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                CU_TRY(cudaStreamCreateWithPriority(&l->tail_stream, cudaStreamNonBlocking, hi));
                CU_TRY(cudaEventCreateWithFlags(&l->scan_ev, cudaEventDisableTiming));
            }
            CU_TRY(cudaEventCreateWithFlags(&l->alloc_ev, cudaEventDisableTiming));
            d->lanes.push_back(std::move(l));
        }
        CU_TRY(cudaStreamCreateWithFlags(&d->free_stream, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
        {   // keep freed blocks in the stream-ordered pool (volumes / query tables are re-loaded often)
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, id) == cudaSuccess) {
                uint64_t keep = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
        }
        g_devices.push_back(std::move(d));
    }
    g_inited = true;
    return BN_OK;
}

static void free_volume_dev(Volume &V)
{
    Gpu *g = g_devices[(size_t)V.device].get();
    cudaSetDevice(g->id);
    cudaStream_t st = g->free_stream;
    if (V.ready) { cudaEventSynchronize(V.ready); cudaEventDestroy(V.ready); V.ready = nullptr; }
    if (V.d_raw) dev_free(V.d_raw, st);
    if (V.d_amb) cudaFreeAsync(V.d_amb, st);
    V.d_amb = nullptr;
    if (V.d_amb_runs) cudaFreeAsync(V.d_amb_runs, st);
    if (V.d_amb_first) cudaFreeAsync(V.d_amb_first, st);
    V.d_raw = nullptr; V.d_packed = nullptr; V.d_amb_runs = nullptr; V.d_amb_first = nullptr;
    std::lock_guard<std::mutex> tlk(V.tmu);
    V.tables.clear();
}

static void free_query_all(Query &Q)
{
    for (size_t d = 0; d < Q.dev.size() && d < g_devices.size(); d++)
        if (Q.dev[d].ready) { cudaSetDevice(g_devices[d]->id); free_query_dev(Q.dev[d], g_devices[d]->free_stream); }
}

void bn_release(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &q : g_queries) if (q) free_query_all(*q);
    g_queries.clear();
    for (auto &v : g_volumes) if (v) free_volume_dev(*v);
    g_volumes.clear();
    for (auto &d : g_devices) {
        cudaSetDevice(d->id);
        for (auto &l : d->lanes) {
            std::lock_guard<std::mutex> ll(l->mu);          // wait for a search still running on the lane
            cudaStreamSynchronize(l->stream);
            l->w.release();
            l->stage.release();
            if (l->arena.base) { arena_unregister(l->arena); cudaFree(l->arena.base); l->arena = DevArena{}; }
            if (l->alloc_ev) cudaEventDestroy(l->alloc_ev);
            if (l->tail_stream) cudaStreamDestroy(l->tail_stream);
            if (l->scan_ev) cudaEventDestroy(l->scan_ev);
            if (l->stream) cudaStreamDestroy(l->stream);
        }
    }
    for (auto &d : g_devices) if (d->free_stream) { cudaSetDevice(d->id); cudaStreamDestroy(d->free_stream); cudaStreamDestroy(d->copy_stream); }
    g_devices.clear();
    g_inited = false;
}

int bn_device_count(void)
{
    if (ensure_init() != BN_OK) return 0;
    return (int)g_devices.size();
}

// async: the upload runs on the device's copy stream and the call returns at once; searches on the
// volume wait for it on the device (Volume::ready)
static int db_load_impl(int device, const uint8_t *packed, int64_t packed_bytes, const int64_t *seq_byte_off,
                        const int32_t *seq_len, int32_t n_seq, bool async, std::shared_ptr<Volume> *out, Lane *D = nullptr)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!packed || !seq_byte_off || !seq_len || n_seq < 0 || !out) return fail(BN_ERR_INVALID, "bn_db_load: bad argument");
    Gpu *G = device_at(device);
    if (!G) return fail(BN_ERR_INVALID, "bn_db_load: bad device");
    LaneLock own;
    if (!D) { own = LaneLock(*G); D = own.lane; }
    for (int32_t i = 0; i < n_seq; i++) {
        const int64_t end = seq_byte_off[i] + (seq_len[i] + 3) / 4;
        if (seq_byte_off[i] < 0 || seq_len[i] < 0 || end + 16 > packed_bytes)
            return fail(BN_ERR_INVALID, "bn_db_load: sequence outside the packed buffer (16 pad bytes required)");
    }
    auto V = std::make_shared<Volume>();
    V->device = device; V->bytes = packed_bytes;
    V->byte_off.assign(seq_byte_off, seq_byte_off + n_seq);
    V->seq_len.assign(seq_len, seq_len + n_seq);
    CU_TRY(cudaSetDevice(D->id));
    // 64 readable bytes in front (reverse 16-base windows may start before the first base) and behind
    TSTEP("volume: begin");
    CU_TRY(dev_malloc((void **)&V->d_raw, (size_t)packed_bytes + 192, D->stream));
    TSTEP("volume: malloc");
    V->d_packed = V->d_raw + 64;
    cudaStream_t cs = async ? G->copy_stream : D->stream;
    if (async) {
        CU_TRY(cudaEventRecord(D->alloc_ev, D->stream));
        CU_TRY(cudaStreamWaitEvent(cs, D->alloc_ev, 0));
    }
    CU_TRY(cudaMemsetAsync(V->d_raw, 0, 64, cs));
    CU_TRY(cudaMemcpyAsync(V->d_packed, packed, (size_t)packed_bytes, cudaMemcpyHostToDevice, cs));
    CU_TRY(cudaMemsetAsync(V->d_packed + packed_bytes, 0, 128, cs));
    TSTEP("volume: copies queued");
    if (async) {
        CU_TRY(cudaEventCreateWithFlags(&V->ready, cudaEventDisableTiming));
        CU_TRY(cudaEventRecord(V->ready, cs));
    } else CU_TRY(cudaStreamSynchronize(cs));
    *out = std::move(V);
    return BN_OK;
}

int bn_db_load(int device, const uint8_t *packed, int64_t packed_bytes, const int64_t *seq_byte_off,
               const int32_t *seq_len, int32_t n_seq, int *vol_handle)
{
    if (!vol_handle) return fail(BN_ERR_INVALID, "bn_db_load: bad argument");
    std::shared_ptr<Volume> V;
    int rc = db_load_impl(device, packed, packed_bytes, seq_byte_off, seq_len, n_seq, false, &V);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    *vol_handle = put_handle(g_volumes, std::move(V));
    return BN_OK;
}

int bn_dbfile_index(const char *nin_path, const char *nsq_path, BnDbFileInfo *info, int64_t *seq_byte_off,
                    int32_t *seq_len)
{
    if (!nin_path || !nsq_path || !info) return fail(BN_ERR_INVALID, "bn_dbfile_index: bad argument");
    DbIndex idx;
    std::string err;
    if (!read_nin(nin_path, idx, err)) return fail(BN_ERR_INVALID, "bn_dbfile_index: " + err);
    MappedFile nsq;
    if (!nsq.open(nsq_path, err)) return fail(BN_ERR_INVALID, "bn_dbfile_index: " + err);
    std::vector<int64_t> off;
    std::vector<int32_t> len;
    if (!sequence_table(idx, nsq.data(), nsq.size(), off, len, err)) return fail(BN_ERR_INVALID, "bn_dbfile_index: " + err);
    memset(info, 0, sizeof *info);
    info->n_seq = idx.n_seq; info->max_len = idx.max_len;
    info->total_bases = (int64_t)idx.total_len; info->nsq_bytes = nsq.size();
    strncpy(info->title, idx.title.c_str(), sizeof info->title - 1);
    if (seq_byte_off) memcpy(seq_byte_off, off.data(), off.size() * sizeof(int64_t));
    if (seq_len) memcpy(seq_len, len.data(), len.size() * sizeof(int32_t));
    return BN_OK;
}

static int get_volume(int vol_handle, std::shared_ptr<Volume> *V);

// Ambiguity runs as the reference applies them (in order, later runs overwrite earlier ones) -> per sequence a
// sorted list of disjoint {first base, end, code} on the device.
static int install_ambiguity(Volume &V, Lane &L, const int64_t *first, const int32_t *runs)
{
    const size_t n_seq = V.seq_len.size();
    std::vector<int4> table;
    std::vector<int64_t> dev_first(n_seq + 1, 0);
    struct Iv { int32_t a, e, code; };
    std::vector<Iv> cur, next;
    for (size_t i = 0; i < n_seq; i++) {
        dev_first[i] = (int64_t)table.size();
        cur.clear();
        for (int64_t k = first[i]; k < first[i + 1]; k++) {
            const int32_t a = runs[3 * k], e = a + runs[3 * k + 1], code = runs[3 * k + 2];
            if (a < 0 || e <= a || e > V.seq_len[i] || code < 0 || code > 15)
                return fail(BN_ERR_INVALID, "ambiguity run outside its sequence");
            if (cur.empty() || cur.back().e <= a) { cur.push_back(Iv{a, e, code}); continue; }      // the usual, ascending case
            next.clear();
            bool placed = false;
            for (const Iv &v : cur) {           // cut [a, e) out of what is there, then insert it in order
                if (v.e <= a || v.a >= e) {
                    if (!placed && v.a >= e) { next.push_back(Iv{a, e, code}); placed = true; }
                    next.push_back(v);
                    continue;
                }
                if (v.a < a) next.push_back(Iv{v.a, a, v.code});
                if (!placed) { next.push_back(Iv{a, e, code}); placed = true; }
                if (v.e > e) next.push_back(Iv{e, v.e, v.code});
            }
            if (!placed) next.push_back(Iv{a, e, code});
            cur.swap(next);
        }
        for (const Iv &v : cur) table.push_back(make_int4(v.a, v.e, v.code, 0));
    }
    dev_first[n_seq] = (int64_t)table.size();
    if ((int64_t)table.size() > INT32_MAX) return fail(BN_ERR_OVERFLOW, "more than 2^31 ambiguity runs in a volume");
    CU_TRY(cudaSetDevice(L.id));
    if (V.d_amb) { CU_TRY(cudaFreeAsync(V.d_amb, L.stream)); V.d_amb = nullptr; }
    if (!table.empty()) {
        CU_TRY(cudaMallocAsync((void **)&V.d_amb, table.size() * sizeof(int4), L.stream));
        CU_TRY(cudaMemcpyAsync(V.d_amb, table.data(), table.size() * sizeof(int4), cudaMemcpyHostToDevice, L.stream));
        CU_TRY(cudaStreamSynchronize(L.stream));
    }
    V.amb_dev_first.swap(dev_first);
    if (table.empty()) V.amb_dev_first.clear();
    return BN_OK;
}

int bn_db_set_ambiguity(int vol_handle, const int64_t *first, const int32_t *runs)
{
    std::shared_ptr<Volume> V;
    int rc = get_volume(vol_handle, &V);
    if (rc) return rc;
    Gpu *g = device_at(V->device);
    if (!g) return fail(BN_ERR_INVALID, "bn_db_set_ambiguity: volume on an unknown device");
    LaneLock lock(*g);
    if (!first) {                                   // remove
        std::vector<int64_t> none(V->seq_len.size() + 1, 0);
        return install_ambiguity(*V, *lock.lane, none.data(), nullptr);
    }
    if (first[V->seq_len.size()] > 0 && !runs) return fail(BN_ERR_INVALID, "bn_db_set_ambiguity: runs is NULL");
    for (size_t i = 0; i < V->seq_len.size(); i++)
        if (first[i + 1] < first[i] || first[i] < 0) return fail(BN_ERR_INVALID, "bn_db_set_ambiguity: first[] must ascend");
    V->amb_first.assign(first, first + V->seq_len.size() + 1);
    V->amb_runs.assign(runs, runs + 3 * first[V->seq_len.size()]);
    return install_ambiguity(*V, *lock.lane, first, runs);
}

int bn_dbfile_ambiguity(const char *nin_path, const char *nsq_path, int64_t *first, int32_t **runs, int64_t *n_runs)
{
    if (!nin_path || !nsq_path || !first || !runs || !n_runs) return fail(BN_ERR_INVALID, "bn_dbfile_ambiguity: bad argument");
    DbIndex idx;
    std::string err;
    if (!read_nin(nin_path, idx, err)) return fail(BN_ERR_INVALID, "bn_dbfile_ambiguity: " + err);
    MappedFile nsq;
    if (!nsq.open(nsq_path, err)) return fail(BN_ERR_INVALID, "bn_dbfile_ambiguity: " + err);
    std::vector<int64_t> f;
    std::vector<int32_t> r;
    if (!ambiguity_table(idx, nsq.data(), nsq.size(), f, r, err)) return fail(BN_ERR_INVALID, "bn_dbfile_ambiguity: " + err);
    memcpy(first, f.data(), f.size() * sizeof(int64_t));
    *n_runs = (int64_t)(r.size() / 3);
    *runs = (int32_t *)malloc(std::max<size_t>(r.size(), 3) * sizeof(int32_t));
    if (!*runs) return fail(BN_ERR_MEMORY, "bn_dbfile_ambiguity: out of memory");
    if (!r.empty()) memcpy(*runs, r.data(), r.size() * sizeof(int32_t));
    return BN_OK;
}

int bn_dbfile_write(const char *nin_path, const char *nsq_path, const char *title, const uint8_t *packed,
                    const int64_t *seq_byte_off, const int32_t *seq_len, int32_t n_seq)
{
    if (!nin_path || !nsq_path || !packed || !seq_byte_off || !seq_len || n_seq < 0)
        return fail(BN_ERR_INVALID, "bn_dbfile_write: bad argument");
    std::string err;
    if (!write_volume(nin_path, nsq_path, title, packed, seq_byte_off, seq_len, n_seq, err))
        return fail(BN_ERR_INVALID, "bn_dbfile_write: " + err);
    return BN_OK;
}

int bn_db_load_files(int device, const char *nin_path, const char *nsq_path, int *vol_handle)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!nin_path || !nsq_path || !vol_handle) return fail(BN_ERR_INVALID, "bn_db_load_files: bad argument");
    Gpu *G = device_at(device);
    if (!G) return fail(BN_ERR_INVALID, "bn_db_load_files: bad device");
    LaneLock own(*G);
    Lane *D = own.lane;
    DbIndex idx;
    std::string err;
    if (!read_nin(nin_path, idx, err)) return fail(BN_ERR_INVALID, "bn_db_load_files: " + err);
    MappedFile nsq;
    if (!nsq.open(nsq_path, err)) return fail(BN_ERR_INVALID, "bn_db_load_files: " + err);
    auto V = std::make_shared<Volume>();
    V->device = device; V->bytes = nsq.size();
    if (!sequence_table(idx, nsq.data(), nsq.size(), V->byte_off, V->seq_len, err))
        return fail(BN_ERR_INVALID, "bn_db_load_files: " + err);
    if (!ambiguity_table(idx, nsq.data(), nsq.size(), V->amb_first, V->amb_runs, err))
        return fail(BN_ERR_INVALID, "bn_db_load_files: " + err);
    CU_TRY(cudaSetDevice(D->id));
    // the file goes to the device as it is (the bytes between sequences are ambiguity data the
    // preliminary stage never reads); zeroed pads in front and behind as in bn_db_load
    CU_TRY(cudaMallocAsync((void **)&V->d_raw, (size_t)nsq.size() + 192, D->stream));
    V->d_packed = V->d_raw + 64;
    CU_TRY(cudaMemsetAsync(V->d_raw, 0, 64, D->stream));
    CU_TRY(cudaMemcpyAsync(V->d_packed, nsq.data(), (size_t)nsq.size(), cudaMemcpyHostToDevice, D->stream));
    CU_TRY(cudaMemsetAsync(V->d_packed + nsq.size(), 0, 128, D->stream));
    CU_TRY(cudaStreamSynchronize(D->stream));
    // the traceback stage lays the ambiguity runs over the packed bases (traceback_kernel.cu: SubjAmb)
    rc = install_ambiguity(*V, *D, V->amb_first.data(), V->amb_runs.data());
    if (rc) { free_volume_dev(*V); return rc; }
    std::lock_guard<std::mutex> lk(g_mu);
    *vol_handle = put_handle(g_volumes, std::move(V));
    return BN_OK;
}

int bn_db_set_masks(int vol_handle, int mask_type, const int32_t *mask_n, const int32_t *mask_iv)
{
    std::shared_ptr<Volume> Vp;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (vol_handle < 0 || vol_handle >= (int)g_volumes.size() || !g_volumes[vol_handle])
            return fail(BN_ERR_INVALID, "bn_db_set_masks: bad handle");
        Vp = g_volumes[vol_handle];
    }
    Volume &V = *Vp;
    if (mask_type != BN_MASK_NONE && mask_type != BN_MASK_SOFT && mask_type != BN_MASK_HARD)
        return fail(BN_ERR_INVALID, "bn_db_set_masks: bad mask type");
    const size_t n = V.seq_len.size();
    std::vector<int64_t> first(n + 1, 0);
    std::vector<int32_t> iv;
    if (mask_type != BN_MASK_NONE) {
        if (!mask_n || !mask_iv) return fail(BN_ERR_INVALID, "bn_db_set_masks: NULL mask arrays");
        for (size_t i = 0; i < n; i++) {
            if (mask_n[i] < 0) return fail(BN_ERR_INVALID, "bn_db_set_masks: negative interval count");
            first[i + 1] = first[i] + mask_n[i];
        }
        iv.assign(mask_iv, mask_iv + 2 * first[n]);
        for (size_t i = 0; i < n; i++) {
            int32_t prev_end = 0;
            for (int64_t k = first[i]; k < first[i + 1]; k++) {
                const int32_t a = iv[(size_t)(2 * k)], e = iv[(size_t)(2 * k + 1)];
                if (a < prev_end || e < a || e > V.seq_len[i])
                    return fail(BN_ERR_INVALID, "bn_db_set_masks: intervals must be ascending, disjoint and inside the sequence");
                prev_end = e;
            }
        }
    }
    std::lock_guard<std::mutex> tlk(V.tmu);         // chunk-table builders read the masks under the same lock
    V.mask_type = mask_type;
    V.mask_first.swap(first);
    V.mask_iv.swap(iv);
    ++V.mask_version;                 // chunk tables are cached per mask version
    V.tables.clear();                 // tables of the old masks are of no use any more
    return BN_OK;
}

int bn_db_free(int h)
{
    std::shared_ptr<Volume> V;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (h < 0 || h >= (int)g_volumes.size() || !g_volumes[h]) return fail(BN_ERR_INVALID, "bn_db_free: bad handle");
        V.swap(g_volumes[h]);
    }
    free_volume_dev(*V);
    return BN_OK;
}

// `held`: a lane of device hook_device the caller already owns (its stream then carries the table build)
static int query_load_impl(const BnQueryBatch *b, std::shared_ptr<Query> *out, int hook_device,
                           const std::function<int()> *after_h2d, Lane *held = nullptr, bool hook_device_only = false,
                           bool keep_async = false)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!b || !out || !b->query_start || !b->contexts || b->num_contexts <= 0)
        return fail(BN_ERR_INVALID, "bn_query_load: bad argument");
    if (b->lut_type != BN_LUT_MB && b->lut_type != BN_LUT_SMALL_NA && b->lut_type != BN_LUT_NA)
        return fail(BN_ERR_UNSUPPORTED, "unknown lookup table type");
    if (b->lut_type == BN_LUT_NA && !b->na_backbone) return fail(BN_ERR_INVALID, "bn_query_load: standard blastn table arrays missing");
    if (b->lut_type == BN_LUT_MB && (!b->hashtable != !b->next_pos)) return fail(BN_ERR_INVALID, "bn_query_load: hashtable and next_pos must come together");
    if (b->lut_type == BN_LUT_MB && !b->hashtable && (!b->lookup_segments || b->n_lookup_segments <= 0))
        return fail(BN_ERR_INVALID, "bn_query_load: MB batch carries neither the table arrays nor lookup_segments");
    if (b->lut_type == BN_LUT_MB && !b->hashtable)
        for (int32_t i = 0; i + 1 < b->n_lookup_segments; i++)
            if (b->lookup_segments[2 * i + 1] >= b->lookup_segments[2 * i + 2])
                return fail(BN_ERR_INVALID, "bn_query_load: lookup_segments must be ascending and disjoint");
    if (b->lut_type == BN_LUT_SMALL_NA && !b->backbone) return fail(BN_ERR_INVALID, "bn_query_load: small table arrays missing");
    if (b->concat_len < 0 || b->concat_len >= (1 << 30)) return fail(BN_ERR_UNSUPPORTED, "bn_query_load: query batch of 2^30 bases or more");
    auto Q = std::make_shared<Query>();
    Q->batch = *b;
    Q->ctx.assign(b->contexts, b->contexts + b->num_contexts);

    for (const auto &c : Q->ctx) Q->max_query_length = std::max(Q->max_query_length, c.query_length);
    // keep only scalars + contexts on the host; remember whether masked_locations was non-NULL
    Q->batch.query_start = nullptr; Q->batch.contexts = Q->ctx.data();
    Q->batch.hashtable = nullptr; Q->batch.next_pos = nullptr; Q->batch.pv_array = nullptr;
    Q->batch.backbone = nullptr; Q->batch.overflow = nullptr;
    Q->batch.lookup_segments = nullptr; Q->batch.n_lookup_segments = 0;
    Q->batch.na_backbone = nullptr; Q->batch.na_overflow = nullptr;
    Q->batch.masked_locations = b->masked_locations ? reinterpret_cast<const int32_t *>(Q->ctx.data()) : nullptr;
    int32_t n = 1;
    while (n < b->concat_len + b->window_size) n <<= 1;   // s_BlastDiagTableNew core/blast_extend.c:46-72
    Q->diag_array_length = n;
    Q->ctx_lite = make_ctx_lite(Q->batch);
    {
        // The direct filter applies to lookup word == full word (s_BlastNaExtendDirect), one-hit mode, the extension
        // routine with the approximate pass (hash container or word >= 11, core/na_ungapped.c:720-721) and scoring
        // tables that are what blastn builds from reward / penalty: nucl_score_table[x] = sum over the four base
        // pairs of x of (pair == 0 ? reward : penalty) (core/blast_parameters.c:250-261), matrix[i][j] = reward on
        // the diagonal and penalty off it for the four bases.
        const BnQueryBatch &qb = Q->batch;
        bool ok = qb.lut_type == BN_LUT_MB && qb.word_length == qb.lut_word_length && qb.window_size == 0 &&
                  (qb.container_type == BN_DIAG_HASH || qb.word_length >= 11) && qb.reward > 0 && qb.penalty < 0;
        for (int x = 0; ok && x < 256; x++) {
            int32_t v = 0;
            for (int k = 0; k < 4; k++) v += ((x >> (2 * k)) & 3) ? qb.penalty : qb.reward;
            ok = qb.nucl_score_table[x] == v;
        }
        for (int i = 0; ok && i < 4; i++)
            for (int j = 0; ok && j < 4; j++) ok = qb.matrix[16 * i + j] == (i == j ? qb.reward : qb.penalty);
        if (getenv("BN_NO_DIRECT_FILTER")) ok = false;
        Q->direct_ok = ok;
        Q->uni_ok = 1;
        Q->uni_x = Q->ctx[0].x_dropoff; Q->uni_cutoff = Q->ctx[0].cutoff_score; Q->uni_reduced = Q->ctx[0].reduced_cutoff;
        for (const auto &c : Q->ctx)
            if (c.x_dropoff != Q->uni_x || c.cutoff_score != Q->uni_cutoff || c.reduced_cutoff != Q->uni_reduced) Q->uni_ok = 0;
    }
    Q->dev.resize(g_devices.size());
    for (size_t d = 0; d < g_devices.size(); d++) {
        if (hook_device_only && (int)d != hook_device) continue;      // a batch of the host-buffer call lives on its device only
        LaneLock own;
        Lane *L = ((int)d == hook_device) ? held : nullptr;
        if (!L) { own = LaneLock(*g_devices[d]); L = own.lane; }
        rc = query_to_device(*Q, *b, (int)d, L, (int)d == hook_device ? after_h2d : nullptr, keep_async && (int)d == hook_device);
        if (rc) { free_query_all(*Q); return rc; }
    }
    *out = std::move(Q);
    return BN_OK;
}

int bn_query_load(const BnQueryBatch *b, int *query_handle)
{
    if (!query_handle) return fail(BN_ERR_INVALID, "bn_query_load: bad argument");
    std::shared_ptr<Query> Q;
    int rc = query_load_impl(b, &Q, -1, nullptr);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    *query_handle = put_handle(g_queries, std::move(Q));
    return BN_OK;
}

int bn_query_free(int h)
{
    std::shared_ptr<Query> Q;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (h < 0 || h >= (int)g_queries.size() || !g_queries[h]) return fail(BN_ERR_INVALID, "bn_query_free: bad handle");
        Q.swap(g_queries[h]);
    }
    free_query_all(*Q);
    return BN_OK;
}

// What a call works on: shared ownership of the volume and the query batch (a concurrent bn_*_free on another
// thread only drops the table's reference) and one lane of the volume's device, held until the call returns.
struct Handles {
    std::shared_ptr<Volume> Vp;
    std::shared_ptr<Query> Qp;
    LaneLock lock;
};

static int get_volume(int vol_handle, std::shared_ptr<Volume> *V)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (vol_handle < 0 || vol_handle >= (int)g_volumes.size() || !g_volumes[vol_handle])
        return fail(BN_ERR_INVALID, "bad volume handle");
    *V = g_volumes[vol_handle];
    return BN_OK;
}

static int get_query(int query_handle, std::shared_ptr<Query> *Q)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (query_handle < 0 || query_handle >= (int)g_queries.size() || !g_queries[query_handle])
        return fail(BN_ERR_INVALID, "bad query handle");
    *Q = g_queries[query_handle];
    return BN_OK;
}

static int get_handles(int vol_handle, int query_handle, Handles &H, Volume **V, Query **Q, Lane **D)
{
    int rc = get_volume(vol_handle, &H.Vp);
    if (rc) return rc;
    rc = get_query(query_handle, &H.Qp);
    if (rc) return rc;
    Gpu *g = device_at(H.Vp->device);
    if (!g) return fail(BN_ERR_INVALID, "volume on an unknown device");
    H.lock = LaneLock(*g);
    *V = H.Vp.get(); *Q = H.Qp.get(); *D = H.lock.lane;
    return BN_OK;
}

int bn_prelim_search(int vol_handle, int query_handle, int32_t oid_begin, int32_t oid_end, int taps,
                     BnResults *out)
{
    Volume *V; Query *Q; Lane *D; Handles H;
    if (!out) return fail(BN_ERR_INVALID, "bn_prelim_search: out is NULL");
    int rc = get_handles(vol_handle, query_handle, H, &V, &Q, &D);
    if (rc) return rc;
    const int32_t n = (int32_t)V->seq_len.size();
    if (oid_begin < 0) oid_begin = 0;
    if (oid_end < 0 || oid_end > n) oid_end = n;
    return prelim_search_locked(*D, *V, *Q, oid_begin, oid_end, taps, out);
}

int bn_prelim_search_host(int device, const BnQueryBatch *batch, const uint8_t *packed,
                          int64_t packed_bytes, const int64_t *seq_byte_off, const int32_t *seq_len,
                          int32_t n_seq, int taps, BnResults *out)
{
    static const bool trace = getenv("BN_TRACE") != nullptr;
    const double t0 = now_ms();
    if (!out) return fail(BN_ERR_INVALID, "bn_prelim_search_host: out is NULL");
    int rc = ensure_init();
    if (rc) return rc;
    Gpu *G = device_at(device);
    if (!G) return fail(BN_ERR_INVALID, "bn_prelim_search_host: bad device");
    // One lane for the whole call.  The volume streams to the device on the lane's copy stream while the query
    // tables are built on its main stream; the copy is queued right behind the (small) query copies.  Volume and
    // batch live for this call only and never enter the handle tables.
    LaneLock lock(*G);
    Lane *L = lock.lane;
    std::shared_ptr<Volume> V;
    std::shared_ptr<Query> Q;
    double t1 = t0;
    const std::function<int()> start_volume = [&]() {
        const int r = db_load_impl(device, packed, packed_bytes, seq_byte_off, seq_len, n_seq, true, &V, L);
        t1 = now_ms();
        return r;
    };
    rc = query_load_impl(batch, &Q, device, &start_volume, L, true);
    const double t2 = now_ms();
    if (rc == BN_OK && !V) rc = fail(BN_ERR_INVALID, "bn_prelim_search_host: volume upload did not start");
    if (rc == BN_OK) rc = prelim_search_locked(*L, *V, *Q, 0, n_seq, taps, out);
    const double t3 = now_ms();
    if (Q) free_query_all(*Q);
    if (V) free_volume_dev(*V);
    if (trace)
        fprintf(stderr, "[bn] host-buffer call %.3f ms: query copies + volume enqueue %.3f table build %.3f search %.3f free %.3f\n",
                now_ms() - t0, t1 - t0, t2 - t1, t3 - t2, now_ms() - t3);
    return rc;
}

int bn_prelim_search_batches(int vol_handle, int32_t n_batches, const BnQueryBatch *const *batches, int taps,
                             BnResults *results)
{
    if (n_batches < 0 || (n_batches > 0 && (!batches || !results))) return fail(BN_ERR_INVALID, "bn_prelim_search_batches: bad argument");
    int rc = ensure_init();
    if (rc) return rc;
    std::shared_ptr<Volume> Vp;
    rc = get_volume(vol_handle, &Vp);
    if (rc) return rc;
    // the job pipeline with one resident volume and host-side batches
    std::vector<BnJob> jobs((size_t)n_batches);
    for (int32_t k = 0; k < n_batches; k++) {
        BnJob j{};
        j.vol_handle = vol_handle; j.query_handle = -1; j.batch = batches[k];
        jobs[(size_t)k] = j;
    }
    return bn_prelim_search_jobs(Vp->device, n_batches, jobs.data(), taps, results, nullptr);
}

static int traceback_search_impl(Lane *D, Volume *V, Query *Q, int32_t gap_x_dropoff_final, const BnHSP *hsps, int64_t n_hsps,
                                 BnTracebackHSP **out, int64_t *n_out, BnEditOp **ops_out, int64_t *n_ops_out);

// Prepare -> (device) -> complete -> host replay -> traceback, as a software pipeline over two lanes of one device
// (header: BnJob).  The caller's thread prepares job k+1 (uploads, chunk table, every kernel of the preliminary stage
// queued on the lane's stream) BEFORE it waits for job k, so the device always has the next search queued and a
// host-side volume crosses PCIe while the previous job's kernels run; one worker thread replays finished jobs on
// the host, a second one runs their traceback stage on a lane of its own.
int bn_prelim_search_jobs(int device, int32_t n_jobs, const BnJob *jobs, int taps, BnResults *results, BnTracebackOut *tb)
{
    if (n_jobs < 0 || (n_jobs > 0 && (!jobs || !results))) return fail(BN_ERR_INVALID, "bn_prelim_search_jobs: bad argument");
    int rc = ensure_init();
    if (rc) return rc;
    Gpu *g = device_at(device);
    if (!g) return fail(BN_ERR_INVALID, "bn_prelim_search_jobs: bad device");
    for (int32_t k = 0; k < n_jobs; k++) {
        memset(&results[k], 0, sizeof results[k]);
        if (tb) memset(&tb[k], 0, sizeof tb[k]);
        if (jobs[k].query_handle < 0 && !jobs[k].batch) return fail(BN_ERR_INVALID, "bn_prelim_search_jobs: job without a query batch");
        if (jobs[k].vol_handle < 0 && !jobs[k].packed) return fail(BN_ERR_INVALID, "bn_prelim_search_jobs: job without a volume");
    }
    if (n_jobs == 0) return BN_OK;
    CU_TRY(cudaSetDevice(g->id));
    // NL lanes (default five, BN_JOB_DEPTH): while the call waits for job k, jobs k+1 .. k+NL-1 are queued on the device.
    // A job's latency (scan -> grouping -> ungapped -> gapped -> mirror, ~0.4 ms for a C2 search) over the number in
    // flight bounds the rate: three lanes gave 0.135 ms per C2 search, four 0.123, five 0.121 (scan kernel alone 0.100)
    int NL = 5;
    if (const char *e = getenv("BN_JOB_DEPTH")) NL = atoi(e);
    NL = std::max(2, std::min(NL, (int)g->lanes.size() - 1));
    std::vector<LaneLock> lanes((size_t)NL + 1);               // [NL]: the traceback stage's lane (tb != NULL)
    std::vector<Stager> stagers((size_t)NL);
    acquire_lanes(*g, tb ? NL + 1 : NL, lanes.data());
    for (int i = 0; i < NL; i++) {
        if (lanes[i].lane->stage.reserve((size_t)6 << 20) == cudaSuccess) { stagers[i].base = lanes[i].lane->stage.p; stagers[i].cap = (size_t)6 << 20; }
    }

    struct JobState {
        std::shared_ptr<Volume> V;
        std::shared_ptr<Query> Q;
        bool own_v = false, own_q = false, freed = false;
        GpuOut G;
        SearchPending P;
        int rc = BN_OK;
        std::string err;
        bool host_done = false, all_done = false;
        double t_prep0 = 0, t_prep1 = 0, t_end0 = 0, t_end1 = 0, t_host0 = 0, t_host1 = 0;     // BN_TRACE timeline
        double t_prep_q = 0, t_prep_v = 0;
        // small result sets leave the lane's pinned mirrors right after the search, so the lane's next job need not
        // wait for this one's host replay
        std::vector<DevInitHit> init_copy;
        std::vector<DevGapResult> gap_copy;
        bool mirrors_free = false;
    };
    std::vector<JobState> J((size_t)n_jobs);
    const double t_call = now_ms();
    std::mutex mu;
    std::condition_variable cv;
    std::deque<int32_t> host_q, tb_q;
    bool host_closed = false, tb_closed = false;

    std::atomic<int> host_workers_live{2};
    auto host_loop = [&]() {
        for (;;) {
            int32_t k;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&]() { return !host_q.empty() || host_closed; });
                if (host_q.empty()) break;
                k = host_q.front(); host_q.pop_front();
            }
            JobState &S = J[(size_t)k];
            S.t_host0 = now_ms();
            const int r = search_host_phase(*S.Q, S.G, taps, &results[k]);
            S.t_host1 = now_ms();
            std::lock_guard<std::mutex> lk(mu);
            if (r) { S.rc = r; S.err = g_err; }
            S.host_done = true;
            if (tb && !r) tb_q.push_back(k); else S.all_done = true;
            cv.notify_all();
        }
        if (host_workers_live.fetch_sub(1) == 1) {
            std::lock_guard<std::mutex> lk(mu);
            tb_closed = true;
            cv.notify_all();
        }
    };
    std::thread host_worker(host_loop), host_worker2(host_loop);
    std::thread tb_worker;
    if (tb) tb_worker = std::thread([&]() {
        for (;;) {
            int32_t k;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&]() { return !tb_q.empty() || tb_closed; });
                if (tb_q.empty()) break;
                k = tb_q.front(); tb_q.pop_front();
            }
            JobState &S = J[(size_t)k];
            // a lane of its own: the traceback kernels run beside the next jobs' searches
            const int r = traceback_search_impl(lanes[NL].lane, S.V.get(), S.Q.get(), jobs[k].gap_x_dropoff_final, results[k].hsps,
                                                results[k].n_hsps, &tb[k].hsps, &tb[k].n_hsps, &tb[k].ops, &tb[k].n_ops);
            std::lock_guard<std::mutex> lk(mu);
            if (r) { S.rc = r; S.err = g_err; }
            S.all_done = true;
            cv.notify_all();
        }
    });

    int first_error = BN_OK;
    std::string first_msg;
    int32_t free_cursor = 0;
    // device memory of the jobs' own volumes / batches goes back to the pool as soon as every stage of the job is through
    auto release_job = [&](JobState &S) {          // every stage of the job is through
        if (S.freed) return;
        S.G.T.reset();
        if (S.own_q && S.Q) free_query_all(*S.Q);
        if (S.own_v && S.V) free_volume_dev(*S.V);
        S.freed = true;
    };
    auto release_finished = [&](bool wait_all) {
        for (; free_cursor < n_jobs; free_cursor++) {
            JobState &S = J[(size_t)free_cursor];
            if (!S.Q && !S.V) { if (!wait_all) break; continue; }     // not prepared (yet)
            {
                std::unique_lock<std::mutex> lk(mu);
                if (wait_all) cv.wait(lk, [&]() { return S.all_done; });
                else if (!S.all_done) break;
            }
            release_job(S);
        }
    };
    auto prepare = [&](int32_t k) -> int {
        JobState &S = J[(size_t)k];
        const BnJob &jb = jobs[k];
        Lane *L = lanes[k % NL].lane;
        S.t_prep0 = now_ms();
        const bool own_data = jb.query_handle < 0 || jb.vol_handle < 0;
        if (k >= NL) {
            JobState &prev = J[(size_t)k - NL];          // the lane's previous job
            std::unique_lock<std::mutex> lk(mu);
            if (prev.own_q || prev.own_v) {              // its batch / volume live in the lane's arena: every stage must be through
                cv.wait(lk, [&]() { return prev.all_done; });
                lk.unlock();
                release_job(prev);
            } else if (!prev.mirrors_free)               // the lane's pinned result mirrors are being read by its host replay
                cv.wait(lk, [&]() { return prev.host_done; });
        }
        stagers[k % NL].used = 0;
        struct UseStager { UseStager(Stager *s) { g_stager = s; } ~UseStager() { g_stager = nullptr; } } use_stager(&stagers[k % NL]);
        release_finished(false);
        // the job's own batch / volume / chunk table come from the lane's arena (devmem.h); the arena grows to what the
        // lane's previous job asked for (the first jobs of a call fall back to the pool)
        struct UseArena {
            bool on = false;
            void set(DevArena *a) { current_arena() = a; on = true; }
            void off() { if (on) current_arena() = nullptr; on = false; }
            ~UseArena() { off(); }
        } use_arena;
        if (own_data) {
            DevArena &A = L->arena;
            if (A.used > A.cap) {
                cudaStreamSynchronize(L->stream);
                cudaStreamSynchronize(L->tail_stream);
                if (A.base) { arena_unregister(A); cudaFree(A.base); A.base = nullptr; A.cap = 0; }
                const size_t want = A.used + A.used / 8 + ((size_t)1 << 20);
                if (cudaMalloc((void **)&A.base, want) == cudaSuccess) { A.cap = want; arena_register(A); }
                else { A.base = nullptr; cudaGetLastError(); }
            }
            A.used = 0;
            use_arena.set(&A);
        }
        int r = BN_OK;
        const std::function<int()> start_volume = [&]() {
            if (jb.vol_handle >= 0) return BN_OK;
            S.own_v = true;
            return db_load_impl(device, jb.packed, jb.packed_bytes, jb.seq_byte_off, jb.seq_len, jb.n_seq, true, &S.V, L);
        };
        if (jb.query_handle >= 0) {
            r = get_query(jb.query_handle, &S.Q);
            if (r == BN_OK) r = start_volume();
        } else {
            S.own_q = true;
            r = query_load_impl(jb.batch, &S.Q, device, &start_volume, L, true, true);
        }
        S.t_prep_q = now_ms();
        if (r) return r;
        if (jb.vol_handle >= 0) {
            r = get_volume(jb.vol_handle, &S.V);
            if (r) return r;
            if (S.V->device != device) return fail(BN_ERR_INVALID, "bn_prelim_search_jobs: volume lives on another device");
        } else if (!S.V) return fail(BN_ERR_INVALID, "bn_prelim_search_jobs: volume upload did not start");
        if (!S.own_v) use_arena.off();          // the chunk table of a resident volume is cached with the volume
        const int32_t n_seq = (int32_t)S.V->seq_len.size();
        r = search_gpu_begin(*L, *S.V, *S.Q, 0, n_seq, &results[k], S.G, triage_allowed(*S.Q, n_seq, taps), S.P);
        S.t_prep1 = now_ms();
        return r;
    };

    int32_t dispatched = 0, prepared = 0;
    while (prepared < std::min<int32_t>(NL - 1, n_jobs)) {
        rc = prepare(prepared);
        if (rc) { first_error = rc; first_msg = g_err; break; }
        prepared++;
    }
    for (int32_t k = 0; k < n_jobs && k < prepared; k++) {
        if (prepared < n_jobs && first_error == BN_OK) {
            rc = prepare(prepared);
            if (rc) { first_error = rc; first_msg = g_err; } else prepared++;
        }
        JobState &S = J[(size_t)k];
        S.t_end0 = now_ms();
        rc = search_gpu_end(*lanes[k % NL].lane, *S.V, *S.Q, &results[k], S.G, S.P);
        S.t_end1 = now_ms();
        if (rc) { if (first_error == BN_OK) { first_error = rc; first_msg = g_err; } break; }
        if (first_error) break;
        if (!S.G.triaged && S.G.n_records <= (int64_t)1 << 16 && S.G.h_init && S.G.h_gap) {
            S.init_copy.assign(S.G.h_init, S.G.h_init + S.G.n_records);
            S.gap_copy.assign(S.G.h_gap, S.G.h_gap + S.G.n_records);
            S.G.h_init = S.init_copy.data(); S.G.h_gap = S.gap_copy.data();
            S.mirrors_free = true;
        } else if (!S.G.triaged && S.G.n_records == 0) S.mirrors_free = true;
        {
            std::lock_guard<std::mutex> lk(mu);
            host_q.push_back(k);
            dispatched = k + 1;
            cv.notify_all();
        }
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        host_closed = true;
        cv.notify_all();
    }
    host_worker.join();
    host_worker2.join();
    if (tb_worker.joinable()) tb_worker.join();
    // jobs that were prepared but never completed still own queued device work
    for (auto &l : lanes) if (l.lane) { cudaStreamSynchronize(l.lane->stream); cudaStreamSynchronize(l.lane->tail_stream); }
    for (int32_t k = 0; k < n_jobs; k++) {
        JobState &S = J[(size_t)k];
        if (k >= dispatched) S.all_done = true;
        if (first_error == BN_OK && S.rc) { first_error = S.rc; first_msg = S.err; }
    }
    release_finished(true);
    if (getenv("BN_TRACE") && atoi(getenv("BN_TRACE")) >= 2)
        for (int32_t k = 0; k < n_jobs; k++) {
            const JobState &S = J[(size_t)k];
            fprintf(stderr, "[bn] job %3d: prepare %.3f (tables %.3f)..%.3f  complete %.3f..%.3f  host %.3f..%.3f | device spans scan %.3f ext %.3f gap %.3f\n",
                    k, S.t_prep0 - t_call, S.t_prep_q - t_call, S.t_prep1 - t_call, S.t_end0 - t_call, S.t_end1 - t_call, S.t_host0 - t_call,
                    S.t_host1 - t_call, results[k].stats.ms_scan, results[k].stats.ms_extend, results[k].stats.ms_gapped);
        }
    if (first_error) {
        for (int32_t k = 0; k < n_jobs; k++) {
            bn_results_free(&results[k]);
            if (tb) { free(tb[k].hsps); free(tb[k].ops); memset(&tb[k], 0, sizeof tb[k]); }
        }
        return fail(first_error, first_msg);
    }
    return BN_OK;
}

// database volumes shard over the GPUs, results meet on the host.  One worker thread per volume
// runs the GPU phase on a lane of the volume's device (volumes on different devices — or on different lanes of
// one device — overlap); the host replay then walks the volumes in order with ONE set of hit lists, so
// hit_params->low_score evolves exactly as in a single pass over the concatenated database
// (core/blast_engine.c:1313-1320) and prelim_hitlist_size (core/hspfilter_collector.c:328-342) is applied once,
// after the gather, never per volume.
int bn_prelim_search_volumes(int32_t n_volumes, const int *vol_handles, int query_handle, int taps,
                             int prune_hitlists, BnResults *out)
{
    if (n_volumes < 0 || (n_volumes > 0 && !vol_handles) || !out)
        return fail(BN_ERR_INVALID, "bn_prelim_search_volumes: bad argument");
    int rc = ensure_init();
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    const double t0 = now_ms();
    std::shared_ptr<Query> Qp;
    rc = get_query(query_handle, &Qp);
    if (rc) return rc;
    Query &Q = *Qp;
    const size_t nv = (size_t)n_volumes;
    std::vector<std::shared_ptr<Volume>> Vs(nv);
    std::vector<int32_t> oid_base(nv + 1, 0);
    for (size_t v = 0; v < nv; v++) {
        rc = get_volume(vol_handles[v], &Vs[v]);
        if (rc) return rc;
        if (!Q.dev[(size_t)Vs[v]->device].ready) return fail(BN_ERR_INVALID, "query batch is not loaded on a volume's device");
        const int64_t next = (int64_t)oid_base[v] + (int64_t)Vs[v]->seq_len.size();
        if (next > INT32_MAX) return fail(BN_ERR_OVERFLOW, "more than 2^31 sequences in the volumes");
        oid_base[v + 1] = (int32_t)next;
    }
    struct Part {
        GpuOut G;
        BnResults res{};
        std::vector<DevInitHit> init;         // copies of the lane's pinned mirrors (the lane moves on)
        std::vector<DevGapResult> gap;
        int rc = BN_OK;
        std::string err;
        bool done = false;
    };
    std::vector<Part> parts(nv);
    const bool volumes_triage = taps == 0 && (int64_t)oid_base[nv] <= (int64_t)LowScoreTracker(Q.batch).hitlist_size();
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<size_t> next_vol{0};
    auto worker = [&]() {
        for (size_t v; (v = next_vol.fetch_add(1)) < nv;) {
            Part &P = parts[v];
            Volume &V = *Vs[v];
            {
                LaneLock lock(*device_at(V.device));
                P.rc = search_gpu_phase(*lock.lane, V, Q, 0, (int32_t)V.seq_len.size(), &P.res, P.G, volumes_triage);
                if (P.rc) P.err = g_err;
                else {
                    P.init.assign(P.G.h_init, P.G.h_init + P.G.n_records);
                    P.gap.assign(P.G.h_gap, P.G.h_gap + P.G.n_records);
                    P.G.h_init = P.init.data(); P.G.h_gap = P.gap.data();
                }
            }
            std::lock_guard<std::mutex> lk(mu);
            P.done = true;
            cv.notify_all();
        }
    };
    size_t n_lanes_total = 0;
    for (auto &g : g_devices) n_lanes_total += g->lanes.size();
    const size_t n_workers = std::min(nv, std::max<size_t>(1, n_lanes_total));
    std::vector<std::thread> pool;
    for (size_t t = 0; t < n_workers; t++) pool.emplace_back(worker);

    // host replay in volume order = OID order of the concatenated database
    LowScoreTracker tracker(Q.batch, prune_hitlists != 0);
    const bool bounds_fixed = (int64_t)oid_base[nv] <= (int64_t)tracker.hitlist_size();
    std::vector<BnHSP> hsps, gapped;
    std::vector<BnInitHit> init;
    BnStats total{};
    int first_error = BN_OK;
    std::string first_msg;
    for (size_t v = 0; v < nv; v++) {
        Part &P = parts[v];
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&]() { return P.done; });
        }
        if (first_error) continue;
        if (P.rc) { first_error = P.rc; first_msg = P.err; continue; }
        HostShared sh;
        sh.tracker = &tracker; sh.bounds_fixed = bounds_fixed; sh.oid_base = oid_base[v];
        const int r = search_host_phase(Q, P.G, taps, &P.res, &sh);
        if (r) { first_error = r; first_msg = g_err; continue; }
        hsps.insert(hsps.end(), P.res.hsps, P.res.hsps + P.res.n_hsps);
        init.insert(init.end(), P.res.init, P.res.init + P.res.n_init);
        gapped.insert(gapped.end(), P.res.gapped, P.res.gapped + P.res.n_gapped);
        const BnStats &s = P.res.stats;
        total.lookup_hits += s.lookup_hits; total.init_extends += s.init_extends;
        total.good_init_extends += s.good_init_extends; total.gap_extensions += s.gap_extensions;
        total.good_extensions += s.good_extensions; total.subject_bases_scanned += s.subject_bases_scanned;
        total.ms_scan += s.ms_scan; total.ms_extend += s.ms_extend; total.ms_gapped += s.ms_gapped;
        total.ms_host += s.ms_host; total.kernel_launches += s.kernel_launches;
        bn_results_free(&P.res);
        P.init = std::vector<DevInitHit>(); P.gap = std::vector<DevGapResult>();
    }
    for (auto &t : pool) t.join();
    for (auto &P : parts) bn_results_free(&P.res);
    if (first_error) return fail(first_error, first_msg);
    if (prune_hitlists && !bounds_fixed) {
        // keep the lists the HSP stream still holds: per query the subjects of its hit list
        std::vector<std::vector<int32_t>> kept((size_t)Q.batch.num_queries);
        for (int32_t q = 0; q < Q.batch.num_queries; q++) kept[(size_t)q] = tracker.kept_oids(q);
        size_t o = 0;
        for (const BnHSP &h : hsps) {
            const std::vector<int32_t> &k = kept[(size_t)Q.batch.contexts[h.context].query_index];
            if (std::binary_search(k.begin(), k.end(), h.oid)) hsps[o++] = h;
        }
        hsps.resize(o);
    }
    out->n_hsps = (int64_t)hsps.size(); out->hsps = to_malloc(hsps);
    out->n_init = (int64_t)init.size(); out->init = to_malloc(init);
    out->n_gapped = (int64_t)gapped.size(); out->gapped = to_malloc(gapped);
    total.ms_total = now_ms() - t0;
    out->stats = total;
    return BN_OK;
}

void bn_results_free(BnResults *r)
{
    if (!r) return;
    free(r->hsps); free(r->init); free(r->gapped);
    memset(r, 0, sizeof *r);
}

void bn_free(void *p) { free(p); }

int bn_selftest_replay(uint64_t seed, int32_t n_cases, int64_t *n_mismatch)
{
    if (!n_mismatch || n_cases < 0) return fail(BN_ERR_INVALID, "bn_selftest_replay: bad argument");
    *n_mismatch = selftest_replay(seed, n_cases);
    return BN_OK;
}

int bn_selftest_sort(int device, int64_t n, int key_bits, uint64_t seed, int64_t *n_mismatch)
{
    if (!n_mismatch || n < 0 || n >= ((int64_t)1 << 31) || key_bits < 1 || key_bits > 64)
        return fail(BN_ERR_INVALID, "bn_selftest_sort: bad argument");
    int rc = ensure_init();
    if (rc) return rc;
    Gpu *g = device_at(device);
    if (!g) return fail(BN_ERR_INVALID, "bn_selftest_sort: bad device");
    CU_TRY(cudaSetDevice(g->id));
    *n_mismatch = 0;
    if (n == 0) return BN_OK;
    LaneLock lk(*g);
    cudaStream_t st = lk.lane->stream;
    // few distinct keys in the low digit, so that equal keys are common and stability is exercised
    std::vector<uint64_t> keys((size_t)n);
    std::vector<SeedHit> vals((size_t)n);
    std::vector<uint32_t> counts((size_t)n);
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    const uint64_t mask = key_bits == 64 ? ~0ull : (((uint64_t)1 << key_bits) - 1);
    for (int64_t i = 0; i < n; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        keys[(size_t)i] = ((x >> 11) & mask) & ~(uint64_t)((x & 3) ? 0 : 0xF0);
        vals[(size_t)i] = SeedHit{(uint32_t)i, (uint32_t)(x >> 40), (uint32_t)(x >> 20), (uint32_t)x};
        counts[(size_t)i] = (uint32_t)(x >> 58);
    }
    uint64_t *ka = nullptr, *kb = nullptr;
    SeedHit *va = nullptr, *vb = nullptr;
    uint32_t *c = nullptr;
    void *tmp = nullptr;
    auto release = [&]() { for (void *p : {(void *)ka, (void *)kb, (void *)va, (void *)vb, (void *)c, tmp}) if (p) cudaFree(p); };
    struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};
    CU_TRY(cudaMalloc((void **)&ka, (size_t)n * 8)); CU_TRY(cudaMalloc((void **)&kb, (size_t)n * 8));
    CU_TRY(cudaMalloc((void **)&va, (size_t)n * 16)); CU_TRY(cudaMalloc((void **)&vb, (size_t)n * 16));
    CU_TRY(cudaMalloc((void **)&c, (size_t)n * 4));
    CU_TRY(cudaMalloc(&tmp, std::max(radix_sort_temp_bytes(n), prefix_sum_temp_bytes(n))));
    CU_TRY(cudaMemcpyAsync(ka, keys.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(va, vals.data(), (size_t)n * 16, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(c, counts.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    bool in_b = false;
    CU_TRY(radix_sort_hits(ka, kb, va, vb, n, key_bits, tmp, &in_b, nullptr, st));
    CU_TRY(prefix_sum_u32(c, c, n, true, tmp, st));
    std::vector<uint64_t> gk((size_t)n);
    std::vector<SeedHit> gv((size_t)n);
    std::vector<uint32_t> gc((size_t)n);
    CU_TRY(cudaMemcpyAsync(gk.data(), in_b ? kb : ka, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(gv.data(), in_b ? vb : va, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(gc.data(), c, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    std::vector<uint32_t> order((size_t)n);
    for (int64_t i = 0; i < n; i++) order[(size_t)i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    uint32_t run = 0;
    int64_t bad = 0;
    for (int64_t i = 0; i < n; i++) {
        const uint32_t o = order[(size_t)i];
        if (gk[(size_t)i] != keys[o] || memcmp(&gv[(size_t)i], &vals[o], sizeof(SeedHit)) != 0) ++bad;
        run += counts[(size_t)i];
        if (gc[(size_t)i] != run) ++bad;
    }
    *n_mismatch = bad;
    return BN_OK;
}

int bn_scan_subject(int vol_handle, int query_handle, int32_t oid, int32_t chunk_off, int32_t chunk_len,
                    BnOffsetPair **pairs, int64_t *n_pairs)
{
    Volume *V; Query *Q; Lane *D; Handles H;
    if (!pairs || !n_pairs) return fail(BN_ERR_INVALID, "bn_scan_subject: NULL output");
    int rc = get_handles(vol_handle, query_handle, H, &V, &Q, &D);
    if (rc) return rc;
    if (oid < 0 || oid >= (int32_t)V->seq_len.size()) return fail(BN_ERR_INVALID, "bn_scan_subject: bad oid");
    CU_TRY(cudaSetDevice(D->id));
    if (!Q->dev[V->device].ready) return fail(BN_ERR_INVALID, "query batch is not loaded on the volume's device");
    std::shared_ptr<ChunkTable> T;
    rc = build_chunk_table(*V, *Q, oid, oid + 1, *D, &T);
    if (rc) return rc;
    if (V->ready) CU_TRY(cudaStreamWaitEvent(D->stream, V->ready, 0));
    StageCounts cnt;
    rc = run_word_finder(*D, *V, *Q, *T, true, cnt, nullptr);
    if (rc) return rc;
    // chunk_len > 0: the one subject chunk [chunk_off, chunk_off + chunk_len) (what one scansub call sequence of the
    // reference covers); chunk_len == 0: every chunk of the subject, in order.  Offsets are chunk-relative.
    int64_t want_chunk = -1;
    if (chunk_len > 0) {
        for (size_t c = 0; c < T->hchunks.size(); c++)
            if (T->hchunks[c].chunk_off == chunk_off && T->hchunks[c].len == chunk_len) want_chunk = (int64_t)c;
        if (want_chunk < 0) return fail(BN_ERR_INVALID, "bn_scan_subject: the subject has no chunk with this offset and length");
    } else if (chunk_off != 0) return fail(BN_ERR_INVALID, "bn_scan_subject: chunk_off without chunk_len");
    std::vector<SeedHit> h((size_t)cnt.n_hits);
    if (cnt.n_hits) {
        CU_TRY(cudaMemcpyAsync(h.data(), D->ws().hits_b.p, h.size() * sizeof(SeedHit), cudaMemcpyDeviceToHost, D->stream));
        CU_TRY(cudaStreamSynchronize(D->stream));
    }
    BnOffsetPair *o = (BnOffsetPair *)malloc(std::max<size_t>(h.size(), 1) * sizeof(BnOffsetPair));
    if (!o) return fail(BN_ERR_MEMORY, "bn_scan_subject: out of memory");
    size_t n_out = 0;
    for (size_t i = 0; i < h.size(); i++) {
        if (want_chunk >= 0 && (int64_t)h[i].chunk != want_chunk) continue;
        o[n_out].q_off = h[i].q_off; o[n_out].s_off = h[i].s_off;
        ++n_out;
    }
    *pairs = o; *n_pairs = (int64_t)n_out;
    return BN_OK;
}

int bn_word_finder(int vol_handle, int query_handle, int32_t oid_begin, int32_t oid_end,
                   BnInitHit **init, int64_t *n_init)
{
    BnResults r;
    if (!init || !n_init) return fail(BN_ERR_INVALID, "bn_word_finder: NULL output");
    int rc = bn_prelim_search(vol_handle, query_handle, oid_begin, oid_end, BN_TAP_INIT, &r);
    if (rc) return rc;
    *init = r.init; *n_init = r.n_init;
    r.init = nullptr;
    bn_results_free(&r);
    return BN_OK;
}

int bn_get_gapped_score(int vol_handle, int query_handle, int32_t oid, int32_t chunk_off,
                        const BnInitHit *init, int64_t n_init, const int32_t *low_score,
                        BnHSP **hsps, int64_t *n_hsps)
{
    Volume *V; Query *Q; Lane *D; Handles H;
    if (!hsps || !n_hsps || n_init < 0 || (n_init > 0 && !init)) return fail(BN_ERR_INVALID, "bn_get_gapped_score: bad argument");
    int rc = get_handles(vol_handle, query_handle, H, &V, &Q, &D);
    if (rc) return rc;
    if (oid < 0 || oid >= (int32_t)V->seq_len.size()) return fail(BN_ERR_INVALID, "bn_get_gapped_score: bad oid");
    CU_TRY(cudaSetDevice(D->id));
    if (!Q->dev[V->device].ready) return fail(BN_ERR_INVALID, "query batch is not loaded on the volume's device");
    std::shared_ptr<ChunkTable> T;
    rc = build_chunk_table(*V, *Q, oid, oid + 1, *D, &T);
    if (rc) return rc;
    int32_t chunk = -1;
    for (size_t c = 0; c < T->hchunks.size(); c++) if (T->hchunks[c].chunk_off == chunk_off) chunk = (int32_t)c;
    if (chunk < 0) return fail(BN_ERR_INVALID, "bn_get_gapped_score: no subject chunk starts at chunk_off");
    if (V->ready) CU_TRY(cudaStreamWaitEvent(D->stream, V->ready, 0));
    *hsps = nullptr; *n_hsps = 0;
    if (n_init == 0) return BN_OK;

    Workspace &ws = D->ws();
    cudaStream_t st = D->stream;
    std::vector<DevInitHit> up((size_t)n_init);
    {
        // caller-supplied records are read by the gapped kernels: the seed and the ungapped segment must lie
        // inside one query context and inside the chunk
        const int32_t clen = T->hchunks[(size_t)chunk].len;
        const BnQueryBatch &b = Q->batch;
        auto context_of = [&](int32_t q) {
            int32_t lo = 0, hi = b.num_contexts;
            while (lo < hi - 1) { const int32_t m = (lo + hi) / 2; if (b.contexts[m].query_offset > q) hi = m; else lo = m; }
            return lo;
        };
        for (int64_t i = 0; i < n_init; i++) {
            const BnInitHit &h = init[i];
            bool ok = h.q_off >= 0 && h.q_off < b.concat_len && h.s_off >= 0 && h.s_off < clen && h.length >= 0 &&
                      h.q_start >= 0 && (int64_t)h.q_start + h.length <= b.concat_len &&
                      h.s_start >= 0 && (int64_t)h.s_start + h.length <= clen;
            if (ok) {
                const BnContext &c = b.contexts[context_of(h.q_off)];
                ok = h.q_off >= c.query_offset && h.q_off < c.query_offset + c.query_length &&
                     h.q_start >= c.query_offset && h.q_start + h.length <= c.query_offset + c.query_length;
            }
            if (!ok) return fail(BN_ERR_INVALID, "bn_get_gapped_score: init hit " + std::to_string(i) + " lies outside the query context or the subject chunk");
            up[(size_t)i] = DevInitHit{chunk, h.q_off, h.s_off, h.q_start, h.s_start, h.length, h.score, (uint32_t)i};
        }
    }
    CU_TRY(ws.counters.reserve(16));
    CU_TRY(ws.init.reserve((size_t)n_init));
    const unsigned long long n_ull = (unsigned long long)n_init;
    CU_TRY(cudaMemsetAsync(ws.counters.p, 0, 8 * sizeof(unsigned long long), st));
    CU_TRY(cudaMemcpyAsync(ws.counters.p + 2, &n_ull, sizeof n_ull, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(ws.init.p, up.data(), up.size() * sizeof(DevInitHit), cudaMemcpyHostToDevice, st));
    DevInitHit *h_init = nullptr;
    DevGapResult *h_gap = nullptr;
    rc = run_gapped(*D, *V, *Q, *T, n_init, h_init, h_gap, nullptr);
    if (rc) return rc;
    std::vector<HostInit> inits((size_t)n_init);
    for (size_t i = 0; i < inits.size(); i++) {
        const DevInitHit &h = h_init[i];
        const DevGapResult &g = h_gap[i];
        inits[i] = HostInit{h.chunk, h.q_off, h.s_off, h.q_start, h.s_start, h.length, h.score, h.order,
                            g.q_start, g.q_stop, g.s_start, g.s_stop, g.score, g.q_seed, g.s_seed, g.status};
    }
    sort_init_hits(inits);
    std::vector<BnHSP> out;
    BnStats stats{};
    replay_gapped(Q->batch, T->hchunks[(size_t)chunk], inits.data(), inits.size(), low_score, out, stats, Q->ctx_lite.data());
    *hsps = to_malloc(out);
    *n_hsps = (int64_t)out.size();
    if (!out.empty() && !*hsps) return fail(BN_ERR_MEMORY, "bn_get_gapped_score: out of memory");
    return BN_OK;
}

// BLAST_GappedAlignmentWithTraceback (core/blast_gapalign.c:3994-4155) for a batch of start points; the
// alignments run on the device (traceback_kernel.cu), the two directions are joined here exactly like
// Blast_PrelimEditBlockToGapEditScript (:2455-2517) and the leading / trailing gap pruning of :4115-4150.
static int traceback_core(Lane *D, Volume *V, Query *Q, int32_t gap_x_dropoff_final,
                          const BnTracebackItem *items, int64_t n_items,
                          BnTracebackResult **results, BnEditOp **ops, int64_t *n_ops)
{
    // the caller holds the lane D and has made the device current
    *results = nullptr; *ops = nullptr; *n_ops = 0;
    const bool greedy = Q->batch.gap_algo == BN_GAP_GREEDY;
    if (!greedy && Q->batch.gap_extend <= 0)
        return fail(BN_ERR_UNSUPPORTED, "bn_gapped_traceback: dynamic programming needs gap_extend > 0");
    if (!Q->dev[V->device].ready) return fail(BN_ERR_INVALID, "query batch is not loaded on the volume's device");
    if (n_items == 0) return BN_OK;
    if (V->ready) CU_TRY(cudaStreamWaitEvent(D->stream, V->ready, 0));
    cudaStream_t st = D->stream;
    const DevQuery &dq = Q->dev[V->device].view;

    std::vector<DevTracebackItem> up((size_t)n_items);
    long long rows = 0;
    for (int64_t i = 0; i < n_items; i++) {
        const BnTracebackItem &t = items[i];
        if (t.oid < 0 || t.oid >= (int32_t)V->seq_len.size() || t.context < 0 || t.context >= Q->batch.num_contexts)
            return fail(BN_ERR_INVALID, "bn_gapped_traceback: bad oid or context");
        const int32_t qlen = Q->batch.contexts[t.context].query_length;
        if (t.s_shift < 0 || t.s_length < 0 || (int64_t)t.s_shift + t.s_length > V->seq_len[(size_t)t.oid] ||
            t.q_start < 0 || t.q_start >= qlen || t.s_start < 0 || t.s_start >= t.s_length)
            return fail(BN_ERR_INVALID, "bn_gapped_traceback: start point outside the sequences");
        up[(size_t)i] = DevTracebackItem{V->byte_off[(size_t)t.oid], t.context, t.s_shift, t.s_length, t.q_start, t.s_start, 0,
                                         V->amb_first_of(t.oid), V->amb_count_of(t.oid), 0};
        rows += qlen + 2;
    }
    if (greedy) return traceback_greedy_host(*D, *V, *Q, gap_x_dropoff_final, items, n_items, up, results, ops, n_ops);
    // arena: row tables (12 B per query row) + script rows (band width ~ 2 (X / gap_extend) + slack) + run lists
    const int32_t xd = std::max(gap_x_dropoff_final, Q->batch.gap_open + Q->batch.gap_extend);
    const long long band = 2ll * (xd / Q->batch.gap_extend + 3) + 32;
    long long arena_bytes = 2 * rows * (12 + band + 8) + 2 * n_items * 8192ll + (16ll << 20);
    long long ops_cap = rows / 2 + 64 * n_items;

    DevTracebackItem *d_items = nullptr;
    DevTracebackDir *d_out = nullptr;
    unsigned long long *d_cnt = nullptr;
    uint8_t *d_arena = nullptr;
    int2 *d_ops = nullptr;
    std::vector<DevTracebackDir> dirs((size_t)(2 * n_items));
    std::vector<int2> h_ops;
    auto release = [&]() {
        if (d_items) cudaFreeAsync(d_items, st);
        if (d_out) cudaFreeAsync(d_out, st);
        if (d_cnt) cudaFreeAsync(d_cnt, st);
        if (d_arena) cudaFreeAsync(d_arena, st);
        if (d_ops) cudaFreeAsync(d_ops, st);
        d_items = nullptr; d_out = nullptr; d_cnt = nullptr; d_arena = nullptr; d_ops = nullptr;
    };
    struct Guard { decltype(release) &f; ~Guard() { f(); } } guard{release};
    CU_TRY(cudaMallocAsync((void **)&d_items, up.size() * sizeof(DevTracebackItem), st));
    CU_TRY(cudaMallocAsync((void **)&d_out, dirs.size() * sizeof(DevTracebackDir), st));
    CU_TRY(cudaMallocAsync((void **)&d_cnt, 2 * sizeof(unsigned long long), st));
    CU_TRY(cudaMemcpyAsync(d_items, up.data(), up.size() * sizeof(DevTracebackItem), cudaMemcpyHostToDevice, st));
    unsigned long long used[2] = {0, 0};
    int2 *d_wide_ring = nullptr;
    uint8_t *d_wide_pf = nullptr;
    struct WideGuard { int2 *&r; uint8_t *&p; cudaStream_t st; ~WideGuard() { if (r) cudaFreeAsync(r, st); if (p) cudaFreeAsync(p, st); } } wide_guard{d_wide_ring, d_wide_pf, st};
    for (int attempt = 0; attempt < 4; attempt++) {
        if (d_arena) { cudaFreeAsync(d_arena, st); d_arena = nullptr; }
        if (d_ops) { cudaFreeAsync(d_ops, st); d_ops = nullptr; }
        CU_TRY(cudaMallocAsync((void **)&d_arena, (size_t)arena_bytes, st));
        CU_TRY(cudaMallocAsync((void **)&d_ops, (size_t)ops_cap * sizeof(int2), st));
        CU_TRY(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), st));
        TracebackLaunch L{};
        L.packed = V->d_packed; L.items = d_items; L.n = n_items; L.x_dropoff = gap_x_dropoff_final; L.amb_runs = V->d_amb;
        L.arena = d_arena; L.arena_bytes = arena_bytes; L.arena_used = d_cnt;
        L.ops = d_ops; L.ops_cap = ops_cap; L.ops_used = d_cnt + 1; L.out = d_out;
        const int wpb = traceback_warps_per_block();
        const int blocks = (int)std::min<int64_t>((2 * n_items + wpb - 1) / wpb, d_wide_ring ? 148 : 148 * 8);
        L.wide_ring = d_wide_ring; L.wide_pf = d_wide_pf;
        CU_TRY(launch_traceback_dp(dq, L, blocks, st));
        CU_TRY(cudaMemcpyAsync(dirs.data(), d_out, dirs.size() * sizeof(DevTracebackDir), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(used, d_cnt, sizeof used, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        bool grow_arena = false, grow_ops = false, too_wide = false;
        for (const DevTracebackDir &d : dirs) {
            too_wide |= d.status == 1;
            grow_arena |= d.status == 3;
            grow_ops |= d.status == 4;
        }
        if (too_wide) {
            // a band beyond the shared-memory ring (2 X_final / gap_extend above ~1000): the batch runs again with the
            // rings in global memory, one grid's worth of them
            if (d_wide_ring) return fail(BN_ERR_OVERFLOW, "bn_gapped_traceback: alignment band wider than the device ring");
            const size_t cells = (size_t)148 * (size_t)wpb * (size_t)traceback_wide_cells();
            CU_TRY(cudaMallocAsync((void **)&d_wide_ring, cells * sizeof(int2), st));
            CU_TRY(cudaMallocAsync((void **)&d_wide_pf, cells, st));
            --attempt;
            continue;
        }
        if (!grow_arena && !grow_ops) break;
        if (attempt == 3) return fail(BN_ERR_OVERFLOW, "bn_gapped_traceback: traceback scratch exhausted");
        if (grow_arena) arena_bytes = std::max<long long>(2 * arena_bytes, (long long)used[0] + (64ll << 20));
        if (grow_ops) ops_cap = std::max<long long>(2 * ops_cap, (long long)used[1] + 1024);
    }
    h_ops.resize((size_t)std::min<unsigned long long>(used[1], (unsigned long long)ops_cap));
    if (!h_ops.empty()) {
        CU_TRY(cudaMemcpyAsync(h_ops.data(), d_ops, h_ops.size() * sizeof(int2), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
    }

    std::vector<BnTracebackResult> res((size_t)n_items);
    std::vector<BnEditOp> out_ops;
    out_ops.reserve(h_ops.size());
    const int32_t gap_open = Q->batch.gap_open, gap_extend = Q->batch.gap_extend;
    for (int64_t i = 0; i < n_items; i++) {
        const BnTracebackItem &t = items[i];
        const DevTracebackDir &l = dirs[(size_t)(2 * i)], &r = dirs[(size_t)(2 * i + 1)];
        BnTracebackResult &o = res[(size_t)i];
        int32_t score_left = l.score, score_right = 0;
        o.query_start = t.q_start - l.a_off + 1;
        o.subject_start = t.s_start - l.b_off + 1;
        if (r.ran) {
            score_right = r.score;
            o.query_stop = t.q_start + r.a_off + 1;
            o.subject_stop = t.s_start + r.b_off + 1;
        } else {
            o.query_stop = t.q_start - 1;
            o.subject_stop = t.s_start - 1;
        }
        // Blast_PrelimEditBlockToGapEditScript: rev as it is, then fwd back to front, equal ops at the seam merged
        const int2 *rev = h_ops.data() + l.ops_off, *fwd = h_ops.data() + r.ops_off;
        const int32_t nr = l.n_ops, nf = r.ran ? r.n_ops : 0;
        const size_t first = out_ops.size();
        for (int32_t k = 0; k < nr; k++) out_ops.push_back(BnEditOp{rev[k].x, rev[k].y});
        if (nf > 0) {
            int32_t k = nf - 1;
            if (nr > 0 && fwd[nf - 1].x == rev[nr - 1].x) { out_ops.back().num += fwd[nf - 1].y; k = nf - 2; }
            for (; k >= 0; k--) out_ops.push_back(BnEditOp{fwd[k].x, fwd[k].y});
        }
        // leading / trailing gaps are cut off (core/blast_gapalign.c:4115-4150)
        size_t size = out_ops.size() - first;
        if (size && out_ops[first].op_type != 3) {
            score_left += gap_open + out_ops[first].num * gap_extend;
            if (out_ops[first].op_type == 0) o.subject_start += out_ops[first].num;
            else o.query_start += out_ops[first].num;
            out_ops.erase(out_ops.begin() + (long)first);
            size--;
        }
        if (size && out_ops[first + size - 1].op_type != 3) {
            score_right += gap_open + out_ops[first + size - 1].num * gap_extend;
            if (out_ops[first + size - 1].op_type == 0) o.subject_stop -= out_ops[first + size - 1].num;
            else o.query_stop -= out_ops[first + size - 1].num;
            out_ops.pop_back();
            size--;
        }
        o.score = score_left + score_right;
        o.esp_off = (int64_t)first;
        o.esp_n = (int32_t)size;
        o.status = 0;
    }
    *results = to_malloc(res);
    *ops = to_malloc(out_ops);
    *n_ops = (int64_t)out_ops.size();
    if (!*results || (!out_ops.empty() && !*ops)) return fail(BN_ERR_MEMORY, "bn_gapped_traceback: out of memory");
    return BN_OK;
}

int bn_gapped_traceback(int vol_handle, int query_handle, int32_t gap_x_dropoff_final,
                        const BnTracebackItem *items, int64_t n_items,
                        BnTracebackResult **results, BnEditOp **ops, int64_t *n_ops)
{
    Volume *V; Query *Q; Lane *D; Handles H;
    if (!results || !ops || !n_ops || n_items < 0 || (n_items > 0 && !items))
        return fail(BN_ERR_INVALID, "bn_gapped_traceback: bad argument");
    int rc = get_handles(vol_handle, query_handle, H, &V, &Q, &D);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(D->id));
    return traceback_core(D, V, Q, gap_x_dropoff_final, items, n_items, results, ops, n_ops);
}

// The per-HSP body of Blast_TracebackFromHSPList (core/blast_traceback.c:490-571) for a list of preliminary HSPs:
// start point (BLAST_CheckStartForGappedAlignment :97-153, BlastGetOffsetsForGappedAlignment
// core/blast_gapalign.c:3059-3131, BlastGetStartForGappedAlignmentNucl :3134-3182) and AdjustSubjectRange (:3608-3636)
// on the device, then the alignment with traceback.
// start points + alignments for a list of preliminary HSPs; the caller holds the lane D and has made the device current
static int traceback_hsps_core(Lane *D, Volume *V, Query *Q, int32_t gap_x_dropoff_final, const BnHSP *hsps, int64_t n_hsps,
                               std::vector<BnTracebackItem> &all, std::vector<BnTracebackResult> &res,
                               BnEditOp **ops, int64_t *n_ops)
{
    *ops = nullptr; *n_ops = 0;
    all.clear(); res.clear();
    if (!Q->dev[V->device].ready) return fail(BN_ERR_INVALID, "query batch is not loaded on the volume's device");
    if (n_hsps == 0) return BN_OK;
    if (V->ready) CU_TRY(cudaStreamWaitEvent(D->stream, V->ready, 0));
    cudaStream_t st = D->stream;
    std::vector<DevTracebackHsp> up((size_t)n_hsps);
    for (int64_t i = 0; i < n_hsps; i++) {
        const BnHSP &h = hsps[i];
        if (h.oid < 0 || h.oid >= (int32_t)V->seq_len.size() || h.context < 0 || h.context >= Q->batch.num_contexts)
            return fail(BN_ERR_INVALID, "bn_traceback_hsps: bad oid or context");
        const int32_t slen = V->seq_len[(size_t)h.oid], qlen = Q->batch.contexts[h.context].query_length;
        if (h.q_off < 0 || h.q_end > qlen || h.q_off >= h.q_end || h.s_off < 0 || h.s_end > slen || h.s_off >= h.s_end)
            return fail(BN_ERR_INVALID, "bn_traceback_hsps: HSP outside the sequences");
        if (!(h.q_gapped_start == 0 && h.s_gapped_start == 0) &&
            (h.q_gapped_start < h.q_off || h.q_gapped_start >= h.q_end || h.s_gapped_start < h.s_off || h.s_gapped_start >= h.s_end))
            return fail(BN_ERR_INVALID, "bn_traceback_hsps: gapped start outside the HSP");
        up[(size_t)i] = DevTracebackHsp{V->byte_off[(size_t)h.oid], slen, h.context, h.q_off, h.q_end, h.s_off, h.s_end,
                                        h.q_gapped_start, h.s_gapped_start, V->amb_first_of(h.oid), V->amb_count_of(h.oid)};
    }
    DevTracebackHsp *d_h = nullptr;
    DevTracebackItem *d_it = nullptr;
    std::vector<DevTracebackItem> its((size_t)n_hsps);
    CU_TRY(cudaMallocAsync((void **)&d_h, up.size() * sizeof(DevTracebackHsp), st));
    CU_TRY(cudaMallocAsync((void **)&d_it, its.size() * sizeof(DevTracebackItem), st));
    CU_TRY(cudaMemcpyAsync(d_h, up.data(), up.size() * sizeof(DevTracebackHsp), cudaMemcpyHostToDevice, st));
    cudaError_t e = launch_traceback_start(Q->dev[V->device].view, V->d_packed, V->d_amb, d_h, n_hsps, d_it, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(its.data(), d_it, its.size() * sizeof(DevTracebackItem), cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(d_h, st); cudaFreeAsync(d_it, st);
    CU_TRY(e);
    CU_TRY(cudaStreamSynchronize(st));
    std::vector<BnTracebackItem> found;
    std::vector<int64_t> where;
    all.resize((size_t)n_hsps);
    for (int64_t i = 0; i < n_hsps; i++) {
        const DevTracebackItem &t = its[(size_t)i];
        BnTracebackItem b{hsps[i].oid, hsps[i].context, t.s_shift, t.s_length, t.q_start, t.s_start};
        if (!t.pad) b.oid = -1;                        // no start point: the reference drops the HSP (:514-518)
        else { found.push_back(b); where.push_back(i); }
        all[(size_t)i] = b;
    }
    BnTracebackResult *r = nullptr;
    int rc = traceback_core(D, V, Q, gap_x_dropoff_final, found.data(), (int64_t)found.size(), &r, ops, n_ops);
    if (rc) return rc;
    res.resize((size_t)n_hsps);
    for (auto &x : res) { memset(&x, 0, sizeof x); x.status = -1; }
    for (size_t k = 0; k < where.size(); k++) res[(size_t)where[k]] = r[k];
    free(r);
    return BN_OK;
}

int bn_traceback_hsps(int vol_handle, int query_handle, int32_t gap_x_dropoff_final,
                      const BnHSP *hsps, int64_t n_hsps, BnTracebackItem **items_out,
                      BnTracebackResult **results, BnEditOp **ops, int64_t *n_ops)
{
    Volume *V; Query *Q; Lane *D; Handles H;
    if (!items_out || !results || !ops || !n_ops || n_hsps < 0 || (n_hsps > 0 && !hsps))
        return fail(BN_ERR_INVALID, "bn_traceback_hsps: bad argument");
    int rc = get_handles(vol_handle, query_handle, H, &V, &Q, &D);
    if (rc) return rc;
    *items_out = nullptr; *results = nullptr; *ops = nullptr; *n_ops = 0;
    CU_TRY(cudaSetDevice(D->id));
    std::vector<BnTracebackItem> all;
    std::vector<BnTracebackResult> res;
    rc = traceback_hsps_core(D, V, Q, gap_x_dropoff_final, hsps, n_hsps, all, res, ops, n_ops);
    if (rc) return rc;
    if (n_hsps == 0) return BN_OK;
    *items_out = to_malloc(all);
    *results = to_malloc(res);
    if (!*items_out || !*results) return fail(BN_ERR_MEMORY, "bn_traceback_hsps: out of memory");
    return BN_OK;
}

// BLAST_ComputeTraceback (core/blast_traceback.c:1375-1640) for blastn / megablast database searches: every
// preliminary HSP is aligned with traceback on the device, the list logic of Blast_TracebackFromHSPList (:336-790)
// is replayed on the host (hostpost.cpp), the re-evaluation and identity counts run on the device again.
static int traceback_search_impl(Lane *D, Volume *V, Query *Q, int32_t gap_x_dropoff_final, const BnHSP *hsps, int64_t n_hsps,
                                 BnTracebackHSP **out, int64_t *n_out, BnEditOp **ops_out, int64_t *n_ops_out);
int bn_traceback_search(int vol_handle, int query_handle, int32_t gap_x_dropoff_final,
                        const BnHSP *hsps, int64_t n_hsps,
                        BnTracebackHSP **out, int64_t *n_out, BnEditOp **ops_out, int64_t *n_ops_out)
{
    Volume *V; Query *Q; Lane *D; Handles H;
    if (!out || !n_out || !ops_out || !n_ops_out || n_hsps < 0 || (n_hsps > 0 && !hsps))
        return fail(BN_ERR_INVALID, "bn_traceback_search: bad argument");
    int rc = get_handles(vol_handle, query_handle, H, &V, &Q, &D);
    if (rc) return rc;
    return traceback_search_impl(D, V, Q, gap_x_dropoff_final, hsps, n_hsps, out, n_out, ops_out, n_ops_out);
}

static int traceback_search_impl(Lane *D, Volume *V, Query *Q, int32_t gap_x_dropoff_final, const BnHSP *hsps, int64_t n_hsps,
                                 BnTracebackHSP **out, int64_t *n_out, BnEditOp **ops_out, int64_t *n_ops_out)
{
    int rc;
    *out = nullptr; *n_out = 0; *ops_out = nullptr; *n_ops_out = 0;
    CU_TRY(cudaSetDevice(D->id));
    const BnQueryBatch &b = Q->batch;
    const bool greedy = b.gap_algo == BN_GAP_GREEDY;
    static const bool trace = getenv("BN_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    std::vector<BnTracebackItem> items;
    std::vector<BnTracebackResult> res;
    BnEditOp *ops = nullptr; int64_t n_ops = 0;
    rc = traceback_hsps_core(D, V, Q, gap_x_dropoff_final, hsps, n_hsps, items, res, &ops, &n_ops);
    if (rc) return rc;
    struct FreeOps { BnEditOp *&p; ~FreeOps() { free(p); } } free_ops{ops};
    if (n_hsps == 0) return BN_OK;
    const double t1 = now();

    // device pass over a set of HSPs with their edit scripts: re-evaluation where asked for, then the identity count
    auto run_post = [&](const std::vector<DevTracebackPost> &post, std::vector<int2> &pops, std::vector<DevTracebackPostOut> &pout) -> int {
        pout.resize(post.size());
        if (post.empty()) return BN_OK;
        cudaStream_t st = D->stream;
        DevTracebackPost *d_p = nullptr; int2 *d_o = nullptr; DevTracebackPostOut *d_r = nullptr;
        CU_TRY(cudaMallocAsync((void **)&d_p, post.size() * sizeof(DevTracebackPost), st));
        CU_TRY(cudaMallocAsync((void **)&d_o, std::max<size_t>(pops.size(), 1) * sizeof(int2), st));
        CU_TRY(cudaMallocAsync((void **)&d_r, post.size() * sizeof(DevTracebackPostOut), st));
        cudaError_t e = cudaMemcpyAsync(d_p, post.data(), post.size() * sizeof(DevTracebackPost), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && !pops.empty()) e = cudaMemcpyAsync(d_o, pops.data(), pops.size() * sizeof(int2), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = launch_traceback_reevaluate(Q->dev[V->device].view, V->d_packed, V->d_amb, d_p, (int64_t)post.size(), d_o, d_r, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pout.data(), d_r, pout.size() * sizeof(DevTracebackPostOut), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && !pops.empty()) e = cudaMemcpyAsync(pops.data(), d_o, pops.size() * sizeof(int2), cudaMemcpyDeviceToHost, st);
        cudaFreeAsync(d_p, st); cudaFreeAsync(d_o, st); cudaFreeAsync(d_r, st);
        CU_TRY(e);
        CU_TRY(cudaStreamSynchronize(st));
        return BN_OK;
    };
    // hit_options->percent_identity / min_hit_length with DP tracebacks: the reference counts the identities of every
    // alignment it makes and tests it before the HSP enters the containment tree (core/blast_traceback.c:658-669), so
    // the counts of ALL speculative alignments are needed before the list replay
    const bool filter_on = identity_filter_on(b);
    std::vector<DevTracebackPostOut> raw_ident;
    if (filter_on && !greedy) {
        std::vector<DevTracebackPost> post;
        std::vector<int2> pops;
        std::vector<int64_t> who((size_t)n_hsps, -1);
        for (int64_t i = 0; i < n_hsps; i++) {
            if (items[(size_t)i].oid < 0) continue;
            const BnTracebackResult &r = res[(size_t)i];
            const int32_t oid = hsps[i].oid, sh = items[(size_t)i].s_shift;
            DevTracebackPost p{};
            p.byte_off = V->byte_off[(size_t)oid]; p.esp_off = (int64_t)pops.size();
            p.seq_len = V->seq_len[(size_t)oid]; p.context = hsps[i].context;
            p.q_off = r.query_start; p.q_end = r.query_stop; p.s_off = r.subject_start + sh; p.s_end = r.subject_stop + sh;
            p.score = r.score; p.esp_n = r.esp_n; p.reevaluate = 0;
            p.amb_first = V->amb_first_of(oid); p.amb_n = V->amb_count_of(oid);
            for (int32_t x = 0; x < r.esp_n; x++) pops.push_back(make_int2(ops[r.esp_off + x].op_type, ops[r.esp_off + x].num));
            who[(size_t)i] = (int64_t)post.size();
            post.push_back(p);
        }
        std::vector<DevTracebackPostOut> pout;
        rc = run_post(post, pops, pout);
        if (rc) return rc;
        raw_ident.assign((size_t)n_hsps, DevTracebackPostOut{});
        for (int64_t i = 0; i < n_hsps; i++) if (who[(size_t)i] >= 0) raw_ident[(size_t)i] = pout[(size_t)who[(size_t)i]];
    }

    // lists = HSPs of one (subject, query) pair in the order given (the preliminary lists are sorted by score)
    std::vector<int64_t> order((size_t)n_hsps);
    for (int64_t i = 0; i < n_hsps; i++) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
        if (hsps[x].oid != hsps[y].oid) return hsps[x].oid < hsps[y].oid;
        return b.contexts[hsps[x].context].query_index < b.contexts[hsps[y].context].query_index;
    });
    struct List { int32_t oid, query_index; std::vector<TbHsp> arr; size_t extra_start; };
    std::vector<List> lists;
    std::vector<TbCand> cand;
    for (size_t k = 0; k < order.size();) {
        const int32_t oid = hsps[order[k]].oid, qi = b.contexts[hsps[order[k]].context].query_index;
        cand.clear();
        size_t e = k;
        for (; e < order.size() && hsps[order[e]].oid == oid && b.contexts[hsps[order[e]].context].query_index == qi; e++) {
            const int64_t i = order[e];
            TbCand c;
            c.pre = hsps[i];
            c.has_start = items[(size_t)i].oid >= 0;
            c.s_shift = items[(size_t)i].s_shift; c.q_start = items[(size_t)i].q_start; c.s_start = items[(size_t)i].s_start;
            c.res = res[(size_t)i];
            c.ops = c.has_start ? ops + res[(size_t)i].esp_off : nullptr;
            c.num_ident = raw_ident.empty() ? 0 : raw_ident[(size_t)i].num_ident;
            c.align_length = raw_ident.empty() ? 0 : raw_ident[(size_t)i].align_length;
            cand.push_back(c);
        }
        List L; L.oid = oid; L.query_index = qi; L.extra_start = 0;
        traceback_list_stage1(b, V->seq_len[(size_t)oid], cand.data(), cand.size(), L.arr, L.extra_start);
        if (!L.arr.empty()) lists.push_back(std::move(L));
        k = e;
    }
    const double t2 = now();
    // device pass: re-evaluation (greedy: every HSP; otherwise the trimmed ones) + identities
    std::vector<DevTracebackPost> post;
    std::vector<int2> pops;
    std::vector<std::pair<size_t, size_t>> who;
    for (size_t li = 0; li < lists.size(); li++) {
        List &L = lists[li];
        for (size_t j = 0; j < L.arr.size(); j++) {
            TbHsp &h = L.arr[j];
            if (!h.alive) continue;
            DevTracebackPost p{};
            p.byte_off = V->byte_off[(size_t)h.oid]; p.esp_off = (int64_t)pops.size();
            p.seq_len = V->seq_len[(size_t)h.oid]; p.context = h.context;
            p.q_off = h.q_off; p.q_end = h.q_end; p.s_off = h.s_off; p.s_end = h.s_end; p.score = h.score;
            p.esp_n = (int32_t)h.esp.size();
            p.reevaluate = (greedy || j >= L.extra_start) ? 1 : 0;
            p.amb_first = V->amb_first_of(h.oid); p.amb_n = V->amb_count_of(h.oid);
            for (const BnEditOp &o : h.esp) pops.push_back(make_int2(o.op_type, o.num));
            post.push_back(p);
            who.emplace_back(li, j);
        }
    }
    std::vector<DevTracebackPostOut> pout;
    rc = run_post(post, pops, pout);
    if (rc) return rc;
    for (size_t k = 0; k < post.size(); k++) {
        TbHsp &h = lists[who[k].first].arr[who[k].second];
        const DevTracebackPostOut &o = pout[k];
        if (o.deleted) { h.alive = false; continue; }
        // Blast_HSPTestIdentityAndLength after the re-evaluation (core/blast_traceback.c:733-735)
        if (filter_on && post[k].reevaluate && hsp_fails_identity_or_length(b, o.num_ident, o.align_length)) { h.alive = false; continue; }
        h.q_off = o.q_off; h.q_end = o.q_end; h.s_off = o.s_off; h.s_end = o.s_end; h.score = o.score;
        h.num_ident = o.num_ident;
        h.esp.clear();
        for (int32_t x = o.first; x <= o.last; x++) {
            const int2 e = pops[(size_t)(post[k].esp_off + x)];
            h.esp.push_back(BnEditOp{e.x, e.y});
        }
    }
    const double t3 = now();
    // second half of the list logic, then the per-query hit lists (Blast_HSPResultsSortByEvalue, s_BlastPruneExtraHits)
    std::vector<size_t> keep;
    for (size_t li = 0; li < lists.size(); li++) {
        traceback_list_stage2(b, V->seq_len[(size_t)lists[li].oid], lists[li].arr);
        if (!lists[li].arr.empty()) keep.push_back(li);
    }
    std::stable_sort(keep.begin(), keep.end(), [&](size_t x, size_t y) {
        if (lists[x].query_index != lists[y].query_index) return lists[x].query_index < lists[y].query_index;
        return traceback_list_before(lists[x].arr, lists[y].arr);
    });
    std::vector<BnTracebackHSP> result;
    std::vector<BnEditOp> result_ops;
    int32_t cur_q = -1, in_q = 0;
    for (size_t li : keep) {
        const List &L = lists[li];
        if (L.query_index != cur_q) { cur_q = L.query_index; in_q = 0; }
        if (b.hitlist_size > 0 && in_q >= b.hitlist_size) continue;
        in_q++;
        for (const TbHsp &h : L.arr) {
            BnTracebackHSP o{};
            o.query_index = L.query_index; o.oid = h.oid; o.context = h.context;
            o.q_off = h.q_off; o.q_end = h.q_end; o.s_off = h.s_off; o.s_end = h.s_end;
            o.score = h.score; o.num_ident = h.num_ident; o.esp_n = (int32_t)h.esp.size();
            o.esp_off = (int64_t)result_ops.size(); o.evalue = h.evalue; o.bit_score = h.bit_score;
            result_ops.insert(result_ops.end(), h.esp.begin(), h.esp.end());
            result.push_back(o);
        }
    }
    if (trace)
        fprintf(stderr, "[bn] traceback stage %.3f ms: start points + alignments %.3f, list stage 1 %.3f, re-evaluation %.3f, "
                        "list stage 2 + results %.3f (%lld HSPs in, %zu out)\n", now() - t0, t1 - t0, t2 - t1, t3 - t2, now() - t3,
                (long long)n_hsps, result.size());
    *out = to_malloc(result); *n_out = (int64_t)result.size();
    *ops_out = to_malloc(result_ops); *n_ops_out = (int64_t)result_ops.size();
    if ((!result.empty() && !*out) || (!result_ops.empty() && !*ops_out)) return fail(BN_ERR_MEMORY, "bn_traceback_search: out of memory");
    return BN_OK;
}

int bn_dust_mask_batch(int device, const uint8_t *seqs, const int32_t *lens, int32_t n_queries, int32_t level, int32_t window,
                       int32_t linker, int32_t **mask_n, int32_t **mask_iv, int64_t *n_intervals)
{
    if (!mask_n || !mask_iv || !n_intervals || n_queries < 0 || (n_queries > 0 && (!seqs || !lens)))
        return fail(BN_ERR_INVALID, "bn_dust_mask_batch: bad argument");
    *mask_n = nullptr; *mask_iv = nullptr; *n_intervals = 0;
    int rc = ensure_init();
    if (rc) return rc;
    Gpu *G = device_at(device);
    if (!G) return fail(BN_ERR_INVALID, "bn_dust_mask_batch: bad device");
    // the parameter ranges of CSymDustMasker's constructor (symdust.cpp:213-226)
    const uint32_t lv = (level >= 2 && level <= 64) ? (uint32_t)level : 20u;
    const uint32_t w = (window >= 8 && window <= 64) ? (uint32_t)window : 64u;
    const uint32_t lk = (linker >= 1 && linker <= 32) ? (uint32_t)linker : 1u;
    const size_t n = (size_t)n_queries;
    std::vector<int64_t> seq_off(n + 1, 0), out_off(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        if (lens[i] < 0) return fail(BN_ERR_INVALID, "bn_dust_mask_batch: negative length");
        seq_off[i + 1] = seq_off[i] + lens[i];
        out_off[i + 1] = out_off[i] + 2 * ((int64_t)lens[i] / 2 + 1);
    }
    const double t0 = now_ms();
    double t_dev0 = t0, t_dev1 = t0;
    std::vector<int32_t> counts(n, 0), flat;
    std::vector<int64_t> coff(n, 0);
    if (n > 0) {
        LaneLock lock(*G);
        Lane *L = lock.lane;
        CU_TRY(cudaSetDevice(L->id));
        cudaStream_t st = L->stream;
        uint8_t *d_seq = nullptr; int64_t *d_soff = nullptr, *d_ooff = nullptr, *d_coff = nullptr;
        int32_t *d_len = nullptr, *d_out = nullptr, *d_n = nullptr, *d_compact = nullptr;
        unsigned long long *d_cursor = nullptr, total = 0;
        cudaError_t e = cudaMallocAsync((void **)&d_seq, (size_t)std::max<int64_t>(seq_off[n], 1), st);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_soff, (n + 1) * sizeof(int64_t), st);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_ooff, (n + 1) * sizeof(int64_t), st);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_coff, n * sizeof(int64_t), st);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_len, n * sizeof(int32_t), st);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_out, (size_t)out_off[n] * sizeof(int32_t), st);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_compact, (size_t)out_off[n] * sizeof(int32_t), st);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_n, n * sizeof(int32_t), st);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_cursor, sizeof(unsigned long long), st);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), st);
        if (e == cudaSuccess && seq_off[n] > 0) e = cudaMemcpyAsync(d_seq, seqs, (size_t)seq_off[n], cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_soff, seq_off.data(), (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ooff, out_off.data(), (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_len, lens, n * sizeof(int32_t), cudaMemcpyHostToDevice, st);
        if (getenv("BN_TRACE")) { cudaStreamSynchronize(st); t_dev0 = now_ms(); }
        if (e == cudaSuccess) e = launch_dust(d_seq, d_soff, d_len, n_queries, lv, w, lk, d_ooff, d_out, d_n, d_compact, d_coff, d_cursor, st);
        if (getenv("BN_TRACE")) { cudaStreamSynchronize(st); t_dev1 = now_ms(); }
        if (e == cudaSuccess) e = cudaMemcpyAsync(counts.data(), d_n, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(coff.data(), d_coff, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_cursor, sizeof total, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess && total > 0) {
            flat.resize((size_t)total);
            e = cudaMemcpyAsync(flat.data(), d_compact, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        }
        void *ptrs[] = {d_seq, d_soff, d_ooff, d_coff, d_len, d_out, d_compact, d_n, d_cursor};
        for (void *p : ptrs) if (p) cudaFreeAsync(p, st);
        CU_TRY(e);
        CU_TRY(cudaStreamSynchronize(st));
    }
    std::vector<int32_t> iv;
    int64_t n_redone = 0;
    const double t_host0 = now_ms();
    int32_t *mn = (int32_t *)malloc(sizeof(int32_t) * std::max<size_t>(n, 1));
    if (!mn) return fail(BN_ERR_MEMORY, "bn_dust_mask_batch: out of memory");
    for (size_t i = 0; i < n; i++) {
        if (counts[i] >= 0) {
            mn[i] = counts[i];
            iv.insert(iv.end(), flat.begin() + coff[i], flat.begin() + coff[i] + 2 * (int64_t)counts[i]);
        } else {            // the kernel's list of perfect intervals overflowed: the host routine redoes this query
            ++n_redone;
            int32_t *one = nullptr, cnt = 0;
            rc = bn_dust_mask(seqs + seq_off[i], lens[i], level, window, linker, &one, &cnt);
            if (rc) { free(mn); return fail(rc, "bn_dust_mask_batch: host pass failed"); }
            mn[i] = cnt;
            iv.insert(iv.end(), one, one + 2 * (size_t)cnt);
            free(one);
        }
    }
    int32_t *out = (int32_t *)malloc(sizeof(int32_t) * std::max<size_t>(iv.size(), 2));
    if (!out) { free(mn); return fail(BN_ERR_MEMORY, "bn_dust_mask_batch: out of memory"); }
    if (!iv.empty()) memcpy(out, iv.data(), iv.size() * sizeof(int32_t));
    *mask_n = mn; *mask_iv = out; *n_intervals = (int64_t)(iv.size() / 2);
    if (getenv("BN_TRACE"))
        fprintf(stderr, "[bn] dust batch %.3f ms: uploads %.3f kernel %.3f results %.3f + host %.3f (%d queries, %lld bases, %zu intervals, %lld redone on the host)\n",
                now_ms() - t0, t_dev0 - t0, t_dev1 - t_dev0, t_host0 - t_dev1, now_ms() - t_host0, n_queries, (long long)seq_off[n], iv.size() / 2,
                (long long)n_redone);
    return BN_OK;
}

int bn_query_download_lookup(int query_handle, int device, int32_t *hashtable, int32_t *next_pos)
{
    std::shared_ptr<Query> Qp;
    int rc = get_query(query_handle, &Qp);
    if (rc) return rc;
    Query &Q = *Qp;
    Gpu *G = device_at(device);
    if (!G || !Q.dev[device].ready || Q.batch.lut_type != BN_LUT_MB || !hashtable || !next_pos)
        return fail(BN_ERR_INVALID, "bn_query_download_lookup: bad argument");
    LaneLock lock(*G);
    Lane *D = lock.lane;
    CU_TRY(cudaSetDevice(D->id));
    cudaStream_t st = D->stream;
    int32_t *tmp = nullptr;
    CU_TRY(cudaMallocAsync((void **)&tmp, (size_t)Q.batch.hashsize * sizeof(int32_t), st));
    CU_TRY(launch_rebuild_hashtable(Q.dev[device].view, Q.batch.hashsize, tmp, st));
    CU_TRY(cudaMemcpyAsync(hashtable, tmp, (size_t)Q.batch.hashsize * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(next_pos, Q.dev[device].next_pos, ((size_t)Q.batch.concat_len + 1) * sizeof(int32_t),
                           cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaFreeAsync(tmp, st));
    CU_TRY(cudaStreamSynchronize(st));
    return BN_OK;
}

int bn_bench_scan(int vol_handle, int query_handle, int iters, double *ms_per_launch,
                  int64_t *bases_per_launch, int64_t *hits)
{
    Volume *V; Query *Q; Lane *D; Handles H;
    int rc = get_handles(vol_handle, query_handle, H, &V, &Q, &D);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(D->id));
    if (!Q->dev[V->device].ready) return fail(BN_ERR_INVALID, "query batch is not loaded on the volume's device");
    std::shared_ptr<ChunkTable> T;
    rc = build_chunk_table(*V, *Q, 0, (int32_t)V->seq_len.size(), *D, &T);
    if (rc) return rc;
    if (V->ready) CU_TRY(cudaStreamWaitEvent(D->stream, V->ready, 0));
    Workspace &ws = D->ws();
    cudaStream_t st = D->stream;
    CU_TRY(ws.counters.reserve(16));
    int64_t cap = std::max<int64_t>((int64_t)ws.hits_a.cap, std::max<int64_t>(1 << 16, T->total_pos / 16));
    CU_TRY(ws.hits_a.reserve((size_t)cap)); CU_TRY(ws.keys_a.reserve((size_t)cap));
    cap = (int64_t)std::min(ws.hits_a.cap, ws.keys_a.cap);
    ScanLaunch s{};
    s.packed = V->d_packed; s.chunks = T->scan_units(); s.n_chunks = T->n_scan_units();
    s.total_pos = T->total_pos; s.hits = ws.hits_a.p; s.keys = ws.keys_a.p;
    s.counters = ws.counters.p; s.capacity = cap; s.block_chunk = T->block_chunk.p; s.block_desc = T->block_desc.p; s.raw_pairs = 0;
    s.gbits = bits_for((uint64_t)std::max<int64_t>(T->total_pos, 1)); s.diag_array_length = Q->diag_array_length;
    s.tile_cap = scan_tile_cap(Q->batch.scan_step, Q->batch.word_length);
    set_direct_filter(s, *Q, *T, false);
    const DevQuery &dq = Q->dev[V->device].view;
    Timer t(st);
    double total = 0;
    unsigned long long h_c[2] = {0, 0};
    for (int i = 0; i < iters; i++) {
        CU_TRY(cudaMemsetAsync(ws.counters.p, 0, 8 * sizeof(unsigned long long), st));
        t.start();
        CU_TRY(launch_scan(dq, s, st));
        t.stop();
        total += t.ms();
    }
    CU_TRY(cudaMemcpyAsync(h_c, ws.counters.p, sizeof h_c, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (ms_per_launch) *ms_per_launch = iters > 0 ? total / iters : 0;
    if (bases_per_launch) *bases_per_launch = T->total_bases;
    if (hits) *hits = (int64_t)h_c[0];
    return BN_OK;
}

}  // extern "C"
