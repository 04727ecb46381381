// capi.cu — host engine and the C-ABI of include/b200_hnsw.h.
//
// Owns the index state in HBM (vectors, level-0 adjacency matrix, upper-level rows), the
// level draw and insertion order of faiss's hnsw_add_vertices, the
// batch schedule of the GPU construction, and the launch of the traversal kernels.
// There is no CPU fallback anywhere in this file: every path ends in a kernel launch.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <random>
#include <string>
#include <vector>

#include "../../include/b200_hnsw.h"
#include "engine.h"

namespace bh {
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches += n; }
}  // namespace bh

namespace {

thread_local std::string t_last_error;

int fail(const std::string& msg) {
    t_last_error = msg;
    return 1;
}

#define BH_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n, cudaStream_t s, bool keep = false, size_t keep_n = 0) {
        if (n <= cap) return cudaSuccess;
        T* np = nullptr;
        cudaError_t e = cudaMalloc(&np, n * sizeof(T));
        if (e != cudaSuccess) return e;
        if (keep && p && keep_n) {
            e = cudaMemcpyAsync(np, p, keep_n * sizeof(T), cudaMemcpyDeviceToDevice, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) {
                cudaFree(np);
                return e;
            }
        }
        if (p) cudaFree(p);
        p = np;
        cap = n;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Page-locked, device-mapped host buffer (staging for pageable callers; the kernels read / write it
// directly over PCIe, see bh_index_search).
template <class T>
struct PinBuf {
    T* p = nullptr;
    T* dev = nullptr;  // the same memory as the device sees it
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        cudaError_t e = cudaHostAlloc((void**)&p, n * sizeof(T), cudaHostAllocPortable | cudaHostAllocMapped);
        if (e != cudaSuccess) {
            p = nullptr;
            return e;
        }
        e = cudaHostGetDevicePointer((void**)&dev, p, 0);
        if (e != cudaSuccess) {
            release();
            return e;
        }
        cap = n;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        dev = nullptr;
        cap = 0;
    }
};

// Resources of one in-flight search call. An index keeps a small pool of these, so concurrent
// bh_index_search calls on one handle run side by side (faiss: `search` is const and thread-safe).
// A context has kLanes streams: a pageable host batch is cut into chunks that go round-robin over
// the lanes, so staging chunk i+1 on the CPU overlaps the traversal of chunk i, and the chunks'
// kernels overlap each other's tails.
constexpr int kLanes = 3;
constexpr int kMaxSearchCtx = 8;
struct SearchLane {
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t done = nullptr;
    PinBuf<float> hq, hD;
    PinBuf<int64_t> hI;
    PinBuf<int32_t> hS;
    int64_t pending_i0 = -1, pending_m = 0;  // chunk whose results sit in hD/hI (not yet copied out)
};
struct SearchCtx {
    SearchLane lane[kLanes];
    DevBuf<int> counters;    // one work counter per lane
    DevBuf<float> q_d;       // re-aligned / zero-padded queries of a *_device call
    DevBuf<uint8_t> sel_d;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // search_device launches on lane 0 overlap their tails (programmatic dependent launch), which forbids
    // a memset between two launches: each launch takes the next counter of a ring; one half of the ring
    // is re-zeroed (an ordinary stream operation, i.e. after everything before it has finished) whenever
    // the other half comes into use
    static constexpr int kRing = 64;
    DevBuf<int> ring;
    unsigned ring_seq = 0;
    bool busy = false;
    cudaError_t init(cudaStream_t primary) {
        cudaError_t e;
        for (int i = 0; i < kLanes; i++) {
            if (i == 0 && primary) {
                lane[i].stream = primary;
            } else {
                if ((e = cudaStreamCreateWithFlags(&lane[i].stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
                lane[i].own_stream = true;
            }
            if ((e = cudaEventCreateWithFlags(&lane[i].done, cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        if ((e = cudaEventCreate(&ev0)) != cudaSuccess) return e;
        if ((e = cudaEventCreate(&ev1)) != cudaSuccess) return e;
        if ((e = ring.reserve(kRing, lane[0].stream)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(ring.p, 0, kRing * sizeof(int), lane[0].stream)) != cudaSuccess) return e;
        return counters.reserve(kLanes, lane[0].stream);
    }
    void destroy() {
        for (int i = 0; i < kLanes; i++) {
            SearchLane& l = lane[i];
            if (l.stream) cudaStreamSynchronize(l.stream);
            l.hq.release(); l.hD.release(); l.hI.release(); l.hS.release();
            if (l.done) cudaEventDestroy(l.done);
            if (l.own_stream && l.stream) cudaStreamDestroy(l.stream);
            l.stream = nullptr;
            l.done = nullptr;
        }
        counters.release(); q_d.release(); sel_d.release();
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        ring.release();
        ev0 = ev1 = nullptr;
    }
};

int ceil_log2(long long v) {
    int b = 0;
    while ((1ll << b) < v) b++;
    return b;
}

}  // namespace

struct bh_index {
    int d = 0;   // the caller's dimension (faiss index.d)
    int dp = 0;  // stored row width in elements: d rounded up to a whole number of 16-byte chunks (4 fp32 /
                 // 8 fp16), the tail zero-filled — adds exact zeros to L2 and IP, so distances are unchanged
    int M = 0, metric = BH_METRIC_L2, device = 0;
    int storage = BH_STORAGE_F32;  // BH_STORAGE_F16 / BH_STORAGE_BF16: rows held in 16 bits (opt-in)
    int efSearch = 16, efConstruction = 40;  // faiss HNSW defaults (App. A.1)
    bool check_relative_distance = true;
    bh_build_params bp{0, 0, 0, 0, 0};
    // add / reset / import hold `rw` exclusively; searches, reconstruct and export share it (faiss: search
    // is const and may run concurrently, add must not overlap anything). Each search call works in a
    // SearchCtx taken from `pool`.
    mutable std::shared_mutex rw;
    mutable std::mutex pool_mu;
    mutable std::condition_variable pool_cv;
    mutable std::vector<std::unique_ptr<SearchCtx>> pool;
    std::vector<double> assign_probas;
    std::vector<int> cum_nn;
    std::mt19937 rng{12345};

    std::vector<int32_t> levels;      // level+1 per vertex (faiss hnsw.levels)
    std::vector<int32_t> upper_base;  // first upper row per vertex, -1 for level-0-only vertices
    int64_t n_upper_rows = 0;
    int64_t ntotal = 0;
    int entry_point = -1, max_level = -1;

    DevBuf<float> vecs;
    DevBuf<int32_t> nbr0, upper_base_d, upper_nbr, slot_head;
    DevBuf<uint8_t> nver0, nverU;  // verified prefix per adjacency row (build_kernels.cu)
    int64_t slot_level0 = 0;  // slot numbering base used when slot_head was laid out
    int64_t row_cap = 0, upper_row_cap = 0;  // rows every per-row / per-upper-row buffer can hold
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int num_sms = 148;
    size_t smem_optin = 227 * 1024;
    DevBuf<int> counter;
    DevBuf<unsigned long long> build_counters;  // [6], see bh_index_last_build_counters
    unsigned long long last_build_counters[6] = {0, 0, 0, 0, 0, 0};
    mutable std::atomic<float> last_search_ms{0.f};
    float last_build_ms = 0.f;
    // build scratch
    DevBuf<int4> items_d;
    DevBuf<unsigned long long> cand_lists;
    DevBuf<int32_t> cand_counts, e_slot, e_src, e_dst, e_level, e_next;
    DevBuf<float> e_dist;

    int deg0() const { return 2 * M; }
    bool half() const { return storage != BH_STORAGE_F32; }  // 16-bit rows
    int fmt16() const { return storage == BH_STORAGE_BF16 ? 2 : 1; }
    int row_floats() const { return half() ? dp / 2 : dp; }  // stored row size in 4-byte units
    void set_padded_dim() { dp = half() ? (d + 7) / 8 * 8 : (d + 3) / 4 * 4; }
    mutable DevBuf<float> conv_d;                           // fp32 staging for fp16 conversion
    mutable std::mutex conv_mu;                             // reconstruct_n calls may overlap

    bh::GraphView view() const {
        bh::GraphView g;
        g.vecs = vecs.p;
        g.nbr0 = nbr0.p;
        g.upper_base = upper_base_d.p;
        g.upper_nbr = upper_nbr.p;
        g.d = dp;
        g.nchunk = row_floats() / 4;
        g.half = half() ? fmt16() : 0;
        g.deg0 = deg0();
        g.degU = M;
        g.entry_point = entry_point;
        g.max_level = max_level;
        g.is_l2 = metric == BH_METRIC_L2;
        return g;
    }

    // App. A.2 — HNSW::set_default_probas(M, 1/ln M)
    void set_default_probas() {
        const float levelMult = (float)(1.0 / std::log((double)M));
        int nn = 0;
        cum_nn.push_back(0);
        for (int level = 0;; level++) {
            float proba = (float)(std::exp(-level / (double)levelMult) * (1 - std::exp(-1 / (double)levelMult)));
            if (proba < 1e-9) break;
            assign_probas.push_back(proba);
            nn += level == 0 ? M * 2 : M;
            cum_nn.push_back(nn);
        }
    }
    // App. A.2 — HNSW::random_level
    int random_level() {
        double f = rng() / float(rng.max());
        for (size_t level = 0; level < assign_probas.size(); level++) {
            if (f < assign_probas[level]) return (int)level;
            f -= assign_probas[level];
        }
        return (int)assign_probas.size() - 1;
    }

    // Visited table per query (beam.cuh). Every policy may forget without changing results, so the
    // size is a pure performance knob: bytes of shared memory per resident query against re-scored
    // vertices. Default = set-associative with 16-bit quotient slots: measured by replaying the
    // traversal on a 1M x 128 graph (scripts/visited_policy_sim.py), 32 remembered vertices per list
    // entry (4 buckets x 8 ways x ef) keep the re-scored fraction at 2-3 % for ef 16..512, against
    // 12-32 % for the round-1 clear-and-re-seed table of the same bytes.
    //   policy 0 (auto): req_bits == 0 -> set-associative, auto size; req_bits > 0 -> exact table of
    //                    2^req_bits slots (the checker mode: exact while it does not fill)
    //   policy 1: exact / clear-and-re-seed;  policy 2: set-associative (4 << req_bits bytes)
    int min_hash_bits(int ef) const { return std::max(8, ceil_log2(((long long)(ef + deg0()) * 4 + 2) / 3 + 1)); }
    void pick_visited(int ef, int policy, int req_bits, int& mode, int& bits) const {
        const bool exact = policy == 1 || (policy == 0 && req_bits > 0);
        if (exact) {
            mode = bh::kVisitedExact;
            const int lo = min_hash_bits(ef);  // 3/4 of the slots must hold the ef-list plus one full row
            if (req_bits > 0) {
                bits = std::min(std::max(req_bits, lo), 16);
            } else {
                int b = std::min(ceil_log2((long long)ef * 4), 10);
                bits = std::min(std::max(std::max(b, 9), lo), 15);
            }
            return;
        }
        if (req_bits > 0) {
            bits = std::min(std::max(req_bits, 4), 15);
        } else if (const char* e = getenv("BH_VISITED_BITS")) {  // experiments only
            bits = atoi(e);
        } else {
            // Measured (profiles/r2_visited_sweep.jsonl, 1M x 128): residency beats the last few per cent
            // of re-scoring — at ef=128 an 8 KB table re-scores 2 % but runs 13 % slower than a 4 KB one
            // that re-scores 9 %, because only 20 instead of 24 queries stay resident per SM. So: the
            // largest table up to 64 remembered vertices per list entry that still lets the full set of
            // query groups (24 per SM; 16 for rows wider than 512 B, which run the 128-register variant)
            // fit in shared memory; never below 2 KB.
            const int ideal = std::min(std::max(ceil_log2((long long)ef * 8) + 2, 9), 13);
            // (and not more than ~200 KB of the SM for the 24 groups: ef=384 with a 4 KB table fits 24
            // groups in 217 KB but runs 12 % slower than with 2 KB — the L1 that is left matters)
            const size_t groups = row_floats() / 4 > 32 ? 16 : 24;
            const size_t budget = std::min<size_t>(smem_optin - (groups / 4) * 1024, 200 * 1024);
            bits = 9;
            for (int b = ideal; b > 9; b--)
                if (groups * bh::beam_group_smem(dp, ef, b, deg0()) <= budget) {
                    bits = b;
                    break;
                }
        }
        // 16-bit slots name an id exactly only while ntotal <= buckets * 2^16
        mode = (ntotal <= (1ll << (bits + 14))) ? bh::kVisitedAssoc16 : bh::kVisitedAssoc32;
    }
    // Warps cooperating on one query. With enough work items to fill the GPU, one warp per query
    // keeps the most queries in flight (throughput regime). With few items (small query batches,
    // the early construction rounds) the machine is idle and a hop's latency is what matters:
    // W warps score a hop's ~50 vectors in one gather round instead of 3-6 serial ones.
    int auto_warps(int ef, int hash_bits, int req, long long n_items) const {
        if (req == 1 || req == 2 || req == 4 || req == 8) return req;
        const size_t s = bh::beam_group_smem(dp, ef, hash_bits, deg0());
        // measured on an idle B200 (scripts/small_batch.py): 11 us/hop at W=1, 6.5-7 us at W=4;
        // W=8 never wins, and beyond ~1k items W=1's higher residency wins.
        int w = 1;
        if (n_items <= 3LL * num_sms) w = 4;
        else if (n_items <= 5LL * num_sms) w = 2;
        // one warp per query needs four queries' state per CTA
        if (w == 1 && 4 * s > smem_optin) w = 2;
        if (w == 2 && 2 * s > smem_optin) w = 4;
        return w;
    }
    // auto_warps, then widened (W -> 2W: half as many query groups per CTA) until the CTA's groups fit in
    // shared memory. False when even one group per CTA does not fit.
    bool pick_warps(int ef, int hb, int rk, int req, long long n_items, int& W) const {
        W = auto_warps(ef + rk, hb, req, n_items);
        const size_t gs = bh::beam_group_smem(dp, ef, hb, deg0(), rk);
        auto groups = [](int w) { return w >= 4 ? 1 : 4 / w; };
        while ((size_t)groups(W) * gs > smem_optin && W < 4) W *= 2;
        return (size_t)groups(W) * gs <= smem_optin;
    }
    // Register/occupancy variant of the one-warp-per-query kernel (beam_kernel.cu): 1 = 80 regs,
    // 6 CTAs/SM, used while 24 queries' state fits in one SM's shared memory; then 3 = 96 regs, 5 CTAs;
    // else 0 = 128 regs, 4 CTAs.
    int beam_variant(int ef, int hash_bits) const {
        const char* e = getenv("BH_BEAM_VARIANT");
        if (e) return atoi(e);
        // rows wider than 32 chunks (512 B) have fewer teams per warp, so the halved-R variants keep too
        // few bytes in flight per warp (measured, ef=256: d=256 0.86 vs 0.92 of the roof, d=512 0.94 vs
        // 1.01, d=768 0.88 vs 1.00) -> full-R, 128-register variant
        if (row_floats() / 4 > 32) return 0;
        const size_t gs = bh::beam_group_smem(dp, ef, hash_bits, deg0());
        if (24 * gs <= smem_optin - 6 * 1024) return 1;
        if (20 * gs <= smem_optin - 5 * 1024) return 3;  // 5 CTAs/SM, <=96 regs
        return 0;
    }

    // Row capacity is committed (row_cap) only after EVERY per-row buffer has reached it, and each
    // buffer is checked on its own (DevBuf::reserve is a no-op when large enough): an allocation
    // failure half-way leaves row_cap at the old value, so the next add() retries the buffers that
    // are still short instead of trusting the ones that already grew.
    int ensure_capacity(int64_t n_new_total, int64_t upper_rows_total) {
        const int64_t old_n = ntotal;
        const int rf = row_floats();
        int64_t want = row_cap;
        if (n_new_total > row_cap) want = std::max<int64_t>(n_new_total, row_cap * 3 / 2);
        BH_CUDA(vecs.reserve((size_t)want * rf, stream, true, (size_t)old_n * rf));
        BH_CUDA(nbr0.reserve((size_t)want * deg0(), stream, true, (size_t)old_n * deg0()));
        BH_CUDA(upper_base_d.reserve((size_t)want, stream, true, (size_t)old_n));
        BH_CUDA(nver0.reserve((size_t)want, stream, true, (size_t)old_n));
        row_cap = want;
        int64_t want_up = upper_row_cap;
        if (upper_rows_total > upper_row_cap) want_up = std::max<int64_t>(upper_rows_total, upper_row_cap * 3 / 2);
        BH_CUDA(upper_nbr.reserve((size_t)want_up * M, stream, true, (size_t)n_upper_rows * M));
        BH_CUDA(nverU.reserve((size_t)want_up + 1, stream, true, (size_t)n_upper_rows));
        upper_row_cap = want_up;
        // pending-list heads: one per adjacency row, all -1 between batches
        const size_t need = (size_t)row_cap + (size_t)upper_row_cap + 1;
        if (need > slot_head.cap || row_cap != slot_level0) {
            BH_CUDA(slot_head.reserve(need, stream));
            BH_CUDA(cudaMemsetAsync(slot_head.p, 0xFF, slot_head.cap * sizeof(int32_t), stream));
            slot_level0 = row_cap;
        }
        return 0;
    }

    void free_all() {
        nver0.release(); nverU.release();
        vecs.release(); nbr0.release(); upper_base_d.release(); upper_nbr.release(); slot_head.release();
        conv_d.release(); build_counters.release();
        counter.release();
        for (auto& c : pool) c->destroy();
        pool.clear();
        items_d.release(); cand_lists.release(); cand_counts.release();
        e_slot.release(); e_src.release(); e_dst.release(); e_level.release(); e_next.release(); e_dist.release();
    }
};

namespace {

// Host rows [m][d] -> [m][dp] with a zero tail (d == dp: one straight copy).
void copy_rows_padded(float* dst, const float* src, int64_t m, int d, int dp) {
    if (d == dp) {
        std::memcpy(dst, src, (size_t)m * d * sizeof(float));
        return;
    }
    for (int64_t i = 0; i < m; i++) {
        std::memcpy(dst + (size_t)i * dp, src + (size_t)i * d, (size_t)d * sizeof(float));
        std::memset(dst + (size_t)i * dp + d, 0, (size_t)(dp - d) * sizeof(float));
    }
}

// Same on the stream: `src` is host or device memory, `dst` device [m][dp] fp32.
cudaError_t copy_rows_padded_async(float* dst, const float* src, int64_t m, int d, int dp,
                                   cudaMemcpyKind kind, cudaStream_t stream) {
    if (d == dp) return cudaMemcpyAsync(dst, src, (size_t)m * d * sizeof(float), kind, stream);
    cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)m * dp * sizeof(float), stream);
    if (e != cudaSuccess) return e;
    return cudaMemcpy2DAsync(dst, (size_t)dp * sizeof(float), src, (size_t)d * sizeof(float),
                             (size_t)d * sizeof(float), (size_t)m, kind, stream);
}

// storage->add: rows [n0, n0+n) into HBM (zero-padded to dp), converting to fp16 on the device when
// that storage is on
int upload_vectors(bh_index* h, int64_t n0, int64_t n, const float* x) {
    const int d = h->d, dp = h->dp;
    if (!h->half()) {
        BH_CUDA(copy_rows_padded_async(h->vecs.p + (size_t)n0 * dp, x, n, d, dp, cudaMemcpyHostToDevice, h->stream));
        return 0;
    }
    const int64_t chunk = std::min<int64_t>(n, 1 << 20);
    BH_CUDA(h->conv_d.reserve((size_t)chunk * dp, h->stream));
    char* dst = reinterpret_cast<char*>(h->vecs.p);
    for (int64_t i0 = 0; i0 < n; i0 += chunk) {
        const int64_t m = std::min(chunk, n - i0);
        BH_CUDA(copy_rows_padded_async(h->conv_d.p, x + (size_t)i0 * d, m, d, dp, cudaMemcpyHostToDevice, h->stream));
        BH_CUDA(bh::launch_f32_to_f16(h->conv_d.p, dst + (size_t)(n0 + i0) * dp * 2, (size_t)m * dp, h->stream, h->fmt16()));
    }
    return 0;
}

// ---- search contexts -----------------------------------------------------------------------
SearchCtx* acquire_ctx(const bh_index* h, bool primary_only) {
    std::unique_lock<std::mutex> lk(h->pool_mu);
    for (;;) {
        const size_t lim = primary_only ? 1 : h->pool.size();
        for (size_t i = 0; i < lim; i++)
            if (!h->pool[i]->busy) {
                h->pool[i]->busy = true;
                return h->pool[i].get();
            }
        if (!primary_only && h->pool.size() < (size_t)kMaxSearchCtx) {
            std::unique_ptr<SearchCtx> c(new SearchCtx());
            const cudaError_t e = c->init(nullptr);
            if (e != cudaSuccess) {
                c->destroy();
                fail(std::string("search: creating a search context failed: ") + cudaGetErrorString(e));
                return nullptr;
            }
            c->busy = true;
            h->pool.push_back(std::move(c));
            return h->pool.back().get();
        }
        h->pool_cv.wait(lk);
    }
}
struct CtxLease {  // gives the context back on every return path
    const bh_index* h;
    SearchCtx* c;
    ~CtxLease() {
        if (!c) return;
        {
            std::lock_guard<std::mutex> lk(h->pool_mu);
            c->busy = false;
        }
        h->pool_cv.notify_all();
    }
};

// Work counter for an overlapping launch on the context's lane 0: the next slot of the ring; entering a
// half of the ring re-zeroes the OTHER half (its launches are >= 32 calls old) with an ordinary memset.
int next_ring_counter(SearchCtx& c, cudaStream_t st, int** counter) {
    const unsigned seq = c.ring_seq++;
    const unsigned slot = seq % SearchCtx::kRing, half = SearchCtx::kRing / 2;
    if (slot % half == 0) BH_CUDA(cudaMemsetAsync(c.ring.p + (slot == 0 ? half : 0), 0, half * sizeof(int), st));
    *counter = c.ring.p + slot;
    return 0;
}

// One traversal launch: n queries at xq_d (device-addressable, 16-byte aligned rows of dp floats).
// `overlap`: launch with programmatic stream serialisation, so that this launch's CTAs may start filling
// the SM slots the PREVIOUS search launch on the stream frees while it drains (see search_device).
int search_device_impl(const bh_index* h, cudaStream_t stream, int* counter, int64_t n, const float* xq_d,
                       int64_t k, float* D_d, int64_t* I_d, int32_t* stats_d, const bh_search_params* params,
                       const uint8_t* sel_dev = nullptr, int n_shard_out = 0,
                       unsigned long long* const* shard_out = nullptr, bool overlap = false, bool solo = true) {
    const int efS = (params && params->efSearch > 0) ? params->efSearch : h->efSearch;
    bool crd = h->check_relative_distance;
    if (params && params->check_relative_distance == 1) crd = true;
    if (params && params->check_relative_distance == 2) crd = false;
    const int ef = (int)std::max<int64_t>(efS, k);
    if (ef > 4096) return fail("max(efSearch, k) > 4096 is not supported");
    const int rk = sel_dev ? (int)k : 0;  // selector-filtered result list lives beside the candidate list
    int vmode = 0, hb = 0;
    h->pick_visited(ef + rk, params ? params->visited_policy : 0, params ? params->hash_bits : 0, vmode, hb);
    int W = 1;
    if (!h->pick_warps(ef, hb, rk, params ? params->warps_per_query : 0, n, W))
        return fail("efSearch/hash_bits need more shared memory than one SM has");
    bh::BeamTask t{};
    t.queries = xq_d;
    t.k = (int)k;
    t.D = D_d;
    t.I = I_d;
    t.items = nullptr;
    t.n_items = (int)n;
    t.ef = ef;
    t.ef_stop = crd ? efS : INT_MAX;
    t.max_steps = crd ? INT_MAX : efS;
    t.hash_bits = hb;
    t.visited_mode = vmode;
    t.stats = stats_d;
    t.sel = sel_dev;
    t.counter = counter;
    t.n_shard_out = n_shard_out;
    for (int p = 0; p < n_shard_out; p++) t.shard_out[p] = shard_out[p];
    t.pdl = overlap ? 1 : 0;
    // drain-mode prefetch (beam.cuh run()): -1.5..4 % for a launch that drains on an otherwise idle GPU (one
    // synchronous call; 10k..1k queries), but +3 % when the next launch is already filling the freed slots
    // (overlapping launches, the chunks of a pageable batch) and +2 % in construction rounds — off there
    t.drain_prefetch = (overlap || !solo) ? 0 : 1;
    if (!overlap) BH_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));  // (overlap: the caller hands out a zeroed counter)
    BH_CUDA(bh::launch_beam(h->view(), t, W, h->beam_variant(ef + rk, hb), h->num_sms, stream, nullptr));
    bh::count_launch();
    return 0;
}

// faiss hnsw_add_vertices' insertion order: buckets by level, highest first, each bucket
// shuffled by RandomGenerator(789) (App. A.7).
void faiss_insertion_order(const std::vector<int32_t>& levels, int64_t n0, int64_t n,
                           std::vector<int32_t>& out) {
    std::vector<int> hist;
    std::vector<int> order(n);
    for (int64_t i = 0; i < n; i++) {
        int pt_level = levels[i + n0] - 1;
        while (pt_level >= (int)hist.size()) hist.push_back(0);
        hist[pt_level]++;
    }
    std::vector<int> offs(hist.size() + 1, 0);
    for (size_t i = 0; i + 1 < hist.size(); i++) offs[i + 1] = offs[i] + hist[i];
    for (int64_t i = 0; i < n; i++) {
        int pt_level = levels[i + n0] - 1;
        order[offs[pt_level]++] = (int)(i + n0);
    }
    std::mt19937 rng2(789);
    out.clear();
    out.reserve(n);
    int i1 = (int)n;
    for (int pt_level = (int)hist.size() - 1; pt_level >= 0; pt_level--) {
        int i0 = i1 - hist[pt_level];
        for (int j = i0; j < i1; j++) std::swap(order[j], order[j + rng2() % (i1 - j)]);
        for (int i = i0; i < i1; i++) out.push_back(order[i]);
        i1 = i0;
    }
}

int add_impl(bh_index* h, int64_t n, const float* x, const int32_t* preset_levels,
             const int32_t* order_in) {
    if (n < 0) return fail("add: n < 0");
    if (n == 0) return 0;
    if (!x) return fail("add: x is null");
    if (h->ntotal + n > (int64_t)INT32_MAX - 1) return fail("add: more than 2^31-2 vectors per index");
    BH_CUDA(cudaSetDevice(h->device));
    const int64_t n0 = h->ntotal;
    const int d = h->dp, M = h->M, deg0 = h->deg0();
    // Argument errors are rejected BEFORE the level RNG advances or any state changes; later failures
    // (allocation, launch configuration) are undone by `rollback` below.
    const int efc = h->efConstruction;
    if (efc < 1 || efc > 4096) return fail("efConstruction must be in [1, 4096]");
    int vmode = 0, hb = 0;
    h->pick_visited(efc, h->bp.visited_policy, h->bp.hash_bits, vmode, hb);
    if (bh::beam_group_smem(d, efc, hb, deg0) > h->smem_optin)
        return fail("efConstruction/hash_bits need more shared memory than one SM has");
    if (preset_levels)
        for (int64_t i = 0; i < n; i++)
            if (preset_levels[i] < 1 || preset_levels[i] > (int)h->assign_probas.size())
                return fail("add: preset level out of range");
    if (order_in) {
        std::vector<char> seen(n, 0);
        for (int64_t i = 0; i < n; i++) {
            const int64_t r = (int64_t)order_in[i] - n0;
            if (r < 0 || r >= n || seen[r]) return fail("add: order is not a permutation of the new ids");
            seen[r] = 1;
        }
    }
    const bool dbg = getenv("BH_DEBUG_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!dbg) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[bh add] %-28s %8.2f ms\n", what,
                std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };

    // From here on host state changes (level RNG, levels, ntotal, entry point). Any failure before the
    // rounds have been enqueued and completed puts all of it back, so a failed add() — allocation
    // failure included — leaves the index answering searches exactly as before. (Device rows beyond
    // the old ntotal may have been written; nothing reads them.) A kernel fault inside the rounds is
    // a sticky CUDA error and takes the context down whatever the host state says.
    struct Rollback {
        bh_index* h;
        std::mt19937 rng;
        size_t nlev;
        int64_t n_upper_rows, ntotal;
        int entry_point, max_level;
        bool armed = true;
        ~Rollback() {
            if (!armed) return;
            h->rng = rng;
            h->levels.resize(nlev);
            h->upper_base.resize(nlev);
            h->n_upper_rows = n_upper_rows;
            h->ntotal = ntotal;
            h->entry_point = entry_point;
            h->max_level = max_level;
        }
    } rollback{h, h->rng, h->levels.size(), h->n_upper_rows, h->ntotal, h->entry_point, h->max_level};

    // -- prepare_level_tab (App. A.2): levels, row allocation
    std::vector<int32_t> new_levels(n);
    for (int64_t i = 0; i < n; i++) {
        if (preset_levels) {
            new_levels[i] = preset_levels[i];
        } else {
            new_levels[i] = h->random_level() + 1;
        }
    }
    int64_t upper_rows = h->n_upper_rows;
    std::vector<int32_t> new_ub(n);
    for (int64_t i = 0; i < n; i++) {
        if (new_levels[i] > 1) {
            new_ub[i] = (int32_t)upper_rows;
            upper_rows += new_levels[i] - 1;
        } else {
            new_ub[i] = -1;
        }
    }
    if (upper_rows > INT32_MAX) return fail("add: too many upper rows");
    lap("level draw");
    if (int rc = h->ensure_capacity(n0 + n, upper_rows)) return rc;
    lap("ensure_capacity (cudaMalloc)");

    // -- storage->add: vectors into HBM; new rows = -1
    if (int rc = upload_vectors(h, n0, n, x)) return rc;
    BH_CUDA(cudaMemsetAsync(h->nbr0.p + (size_t)n0 * deg0, 0xFF, (size_t)n * deg0 * sizeof(int32_t), h->stream));
    if (upper_rows > h->n_upper_rows)
        BH_CUDA(cudaMemsetAsync(h->upper_nbr.p + (size_t)h->n_upper_rows * M, 0xFF,
                                (size_t)(upper_rows - h->n_upper_rows) * M * sizeof(int32_t), h->stream));
    BH_CUDA(cudaMemcpyAsync(h->upper_base_d.p + n0, new_ub.data(), (size_t)n * sizeof(int32_t),
                            cudaMemcpyHostToDevice, h->stream));
    BH_CUDA(cudaMemsetAsync(h->nver0.p + n0, 0, (size_t)n, h->stream));
    if (upper_rows > h->n_upper_rows)
        BH_CUDA(cudaMemsetAsync(h->nverU.p + h->n_upper_rows, 0, (size_t)(upper_rows - h->n_upper_rows), h->stream));
    BH_CUDA(cudaStreamSynchronize(h->stream));  // new_ub / x may go out of scope
    lap("H2D vectors + row init");
    h->levels.insert(h->levels.end(), new_levels.begin(), new_levels.end());
    h->upper_base.insert(h->upper_base.end(), new_ub.begin(), new_ub.end());
    h->n_upper_rows = upper_rows;
    h->ntotal = n0 + n;

    // -- insertion order
    std::vector<int32_t> order;
    if (order_in) {
        order.assign(order_in, order_in + n);  // validated above
    } else {
        faiss_insertion_order(h->levels, n0, n, order);
    }

    // -- batch schedule. A round inserts points concurrently against the graph as it stood at the
    //    start of the round (faiss's OpenMP build has the same blindness between the points its
    //    threads are inserting at one moment). Rounds are kept small relative to the graph.
    // auto: 10240 points per round (~3 waves of resident insertion searches), growing with the graph
    // beyond 2.6M vertices (a round never exceeds 1/256 of the graph there), capped at 64k
    const bool auto_batch = h->bp.max_batch <= 0;
    const int max_batch = auto_batch ? 10240 : h->bp.max_batch;
    const bool auto_div = h->bp.batch_divisor <= 0;
    // auto: a round is at most 1/32 of the graph (measured at 1M x 128, efC=200: 1/64 -> 1/32 takes the build
    // from 1.50 s to 1.39 s at -0.05 pt recall@10; 1/24 costs -0.2 pt; profiles/README.md)
    const int divisor = auto_div ? 32 : h->bp.batch_divisor;
    struct Round { int64_t item_begin, item_end; int new_entry, new_max_level; };
    std::vector<int4> items;
    std::vector<Round> rounds;
    items.reserve((size_t)n + n / 8);
    {
        int cur_max_level = h->max_level;
        int64_t in_graph = n0;
        int64_t pos = 0;
        if (h->entry_point < 0) {  // very first vertex: becomes the entry point, no links
            const int pt = order[0];
            cur_max_level = h->levels[pt] - 1;
            h->entry_point = pt;
            h->max_level = cur_max_level;
            in_graph = 1;
            pos = 1;
        }
        while (pos < n) {
            int64_t cap = max_batch;
            if (auto_batch) cap = std::min<int64_t>(65536, std::max<int64_t>(cap, in_graph / 256));
            // While the graph is below 1/32 of the size this call will reach, rounds may be coarser
            // (1/16 of the graph): those vertices are <= 3 % of the final graph and their rows are
            // reworked by the back-links of everything inserted later.
            const int div_now = (auto_div && in_graph * 32 < n0 + n) ? std::min(divisor, 16) : divisor;
            int64_t target = std::max<int64_t>(1, std::min<int64_t>(cap, in_graph / div_now));
            Round r{(int64_t)items.size(), 0, -1, -1};
            int64_t cnt = 0;
            while (pos < n && cnt < target) {
                const int pt = order[pos];
                const int pt_level = h->levels[pt] - 1;
                if (pt_level > cur_max_level) {
                    if (cnt > 0) break;  // close the round; this point gets a round of its own
                    for (int l = cur_max_level; l >= 0; l--) items.push_back(make_int4(pt, l, pt_level, 0));
                    r.new_entry = pt;
                    r.new_max_level = pt_level;
                    cur_max_level = pt_level;
                    pos++;
                    cnt++;
                    break;
                }
                for (int l = pt_level; l >= 0; l--) items.push_back(make_int4(pt, l, pt_level, 0));
                pos++;
                cnt++;
            }
            r.item_end = (int64_t)items.size();
            in_graph += cnt;
            rounds.push_back(r);
        }
    }
    if (items.empty()) {
        rollback.armed = false;
        return 0;
    }
    lap("order + round schedule");
    size_t max_items = 0;
    for (const Round& r : rounds) max_items = std::max(max_items, (size_t)(r.item_end - r.item_begin));


    BH_CUDA(h->items_d.reserve(items.size(), h->stream));
    lap("  items alloc");
    BH_CUDA(cudaMemcpyAsync(h->items_d.p, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice,
                            h->stream));
    lap("  items H2D");
    BH_CUDA(h->cand_lists.reserve(max_items * efc, h->stream));
    BH_CUDA(h->cand_counts.reserve(max_items, h->stream));
    lap("  cand alloc");
    const size_t max_edges = max_items * deg0;
    BH_CUDA(h->e_slot.reserve(max_edges, h->stream));
    BH_CUDA(h->e_src.reserve(max_edges, h->stream));
    BH_CUDA(h->e_dst.reserve(max_edges, h->stream));
    BH_CUDA(h->e_level.reserve(max_edges, h->stream));
    BH_CUDA(h->e_next.reserve(max_edges, h->stream));
    BH_CUDA(h->e_dist.reserve(max_edges, h->stream));

    BH_CUDA(h->build_counters.reserve(6, h->stream));
    BH_CUDA(cudaMemsetAsync(h->build_counters.p, 0, 6 * sizeof(unsigned long long), h->stream));
    lap("scratch alloc + items H2D");
    BH_CUDA(cudaEventRecord(h->ev0, h->stream));
    for (const Round& r : rounds) {
        const int n_items = (int)(r.item_end - r.item_begin);
        if (n_items > 0) {
            bh::GraphView g = h->view();
            bh::BeamTask t{};
            t.items = h->items_d.p + r.item_begin;
            t.out_lists = h->cand_lists.p;
            t.out_counts = h->cand_counts.p;
            t.n_items = n_items;
            t.ef = efc;
            t.ef_stop = INT_MAX;
            t.max_steps = INT_MAX;
            t.hash_bits = hb;
            t.visited_mode = vmode;
            t.stats = nullptr;
            t.build_counters = h->build_counters.p;
            t.counter = h->counter.p;
            t.drain_prefetch = 0;
            BH_CUDA(cudaMemsetAsync(h->counter.p, 0, sizeof(int), h->stream));
            int W = 1;
            if (!h->pick_warps(efc, hb, 0, h->bp.warps_per_query, n_items, W))
                return fail("efConstruction/hash_bits need more shared memory than one SM has");
            bh::BuildBatch b{};
            b.items = t.items;
            b.cand_lists = h->cand_lists.p;
            b.cand_counts = h->cand_counts.p;
            b.n_items = n_items;
            b.efc = efc;
            b.edge_dst_slot = h->e_slot.p;
            b.edge_src = h->e_src.p;
            b.edge_dst = h->e_dst.p;
            b.edge_level = h->e_level.p;
            b.edge_dist = h->e_dist.p;
            b.edge_next = h->e_next.p;
            b.slot_head = h->slot_head.p;
            b.n_level0 = h->slot_level0;
            b.nver0 = h->nver0.p;
            b.nverU = h->nverU.p;
            // rows with up to 16 unverified members take the incremental shrink (measured at 1M x 128: 8 -> 16 takes
            // the build from 1.47 s to 1.38 s; 24 costs shared memory / occupancy again: 1.42 s)
            b.max_special = std::min(16, std::max(2, (16 * 1024) / (4 * h->row_floats())));  // (wide rows: a larger
            // staging budget than 16 KB per warp was measured slower at 768-d and 960-d)
            b.build_counters = h->build_counters.p;
            // rows above 512 B: selection + forward links run in the traversal kernel's epilogue (see launch_one)
            const bool fuse = h->row_floats() / 4 > 32;
            BH_CUDA(bh::launch_beam(g, t, W, h->beam_variant(efc, hb), h->num_sms, h->stream, nullptr, fuse ? &b : nullptr));
            if (!fuse) BH_CUDA(bh::launch_select_and_link(g, b, h->num_sms, h->stream));
            BH_CUDA(bh::launch_backlinks(g, b, h->num_sms, h->stream));
            bh::count_launch(fuse ? 2 : 3);
        }
        if (r.new_entry >= 0) {  // add_with_locks tail: a taller point becomes the entry point
            h->entry_point = r.new_entry;
            h->max_level = r.new_max_level;
        }
    }
    BH_CUDA(cudaEventRecord(h->ev1, h->stream));
    lap("enqueue rounds (host)");
    BH_CUDA(cudaStreamSynchronize(h->stream));
    lap("wait for device");
    BH_CUDA(cudaEventElapsedTime(&h->last_build_ms, h->ev0, h->ev1));
    BH_CUDA(cudaMemcpy(h->last_build_counters, h->build_counters.p, 6 * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost));
    rollback.armed = false;
    return 0;
}

}  // namespace

// =================================================================== C-ABI
extern "C" {

const char* bh_last_error(void) { return t_last_error.c_str(); }
const char* bh_version(void) { return "b200-hnsw 0.1 (sm_100a)"; }
int64_t bh_launch_count(void) { return (int64_t)bh::g_launches.load(); }

int bh_index_create(bh_index** out, int d, int M, int metric, int device) {
    if (!out) return fail("create: out is null");
    *out = nullptr;
    if (d <= 0 || d > 2048) return fail("create: d must be in [1, 2048]");
    if (M < 2 || 2 * M > bh::kMaxDeg) return fail("create: M must be in [2, 64]");
    if (metric != BH_METRIC_L2 && metric != BH_METRIC_INNER_PRODUCT)
        return fail("create: metric must be METRIC_L2 (1) or METRIC_INNER_PRODUCT (0)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(std::string("create: no CUDA device (this engine has no CPU fallback): ") +
                    cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail("create: bad device ordinal");
    BH_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    BH_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("create: this library is built for sm_100a (B200) only");
    bh_index* h = new bh_index();
    h->d = d;
    h->set_padded_dim();
    h->M = M;
    h->metric = metric;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    h->set_default_probas();
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        h->counter.reserve(4, h->stream) != cudaSuccess) {
        delete h;
        return fail("create: stream/event/alloc failed");
    }
    {  // primary search context: its lane 0 is the index's stream
        std::unique_ptr<SearchCtx> c(new SearchCtx());
        if (c->init(h->stream) != cudaSuccess) {
            c->destroy();
            bh_index_free(h);
            return fail("create: search context allocation failed");
        }
        h->pool.push_back(std::move(c));
    }
    *out = h;
    return 0;
}

int bh_index_free(bh_index* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    h->free_all();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

static int reset_locked(bh_index* h) {
    BH_CUDA(cudaSetDevice(h->device));
    BH_CUDA(cudaStreamSynchronize(h->stream));
    h->levels.clear();
    h->upper_base.clear();
    h->n_upper_rows = 0;
    h->ntotal = 0;
    h->entry_point = -1;
    h->max_level = -1;
    // faiss IndexHNSW::reset → hnsw.reset() keeps the RNG state; so do we.
    return 0;
}

int bh_index_reset(bh_index* h) {
    if (!h) return fail("null index");
    std::unique_lock<std::shared_mutex> lk(h->rw);
    return reset_locked(h);
}

int bh_index_set_vector_storage(bh_index* h, int storage) {
    if (!h) return fail("null index");
    if (storage != BH_STORAGE_F32 && storage != BH_STORAGE_F16 && storage != BH_STORAGE_BF16)
        return fail("unknown storage kind");
    std::unique_lock<std::shared_mutex> lk(h->rw);
    if (h->ntotal != 0) return fail("set_vector_storage: the index is not empty");
    if ((storage != BH_STORAGE_F32) != h->half()) {  // capacities are counted in rows of the old width: start over
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        h->vecs.release();
        h->nbr0.release();
        h->upper_base_d.release();
        h->nver0.release();
        h->slot_level0 = 0;
        h->row_cap = 0;
    }
    h->storage = storage;
    h->set_padded_dim();
    return 0;
}
int bh_index_get_vector_storage(const bh_index* h) { return h ? h->storage : -1; }

int bh_index_train(bh_index* h, int64_t, const float*) {
    if (!h) return fail("null index");
    return 0;  // IndexFlat storage needs no training; is_trained is always true
}

int bh_index_add(bh_index* h, int64_t n, const float* x) {
    if (!h) return fail("null index");
    std::unique_lock<std::shared_mutex> lk(h->rw);
    return add_impl(h, n, x, nullptr, nullptr);
}

int bh_index_add_ex(bh_index* h, int64_t n, const float* x, const int32_t* levels, const int32_t* order) {
    if (!h) return fail("null index");
    std::unique_lock<std::shared_mutex> lk(h->rw);
    return add_impl(h, n, x, levels, order);
}

int bh_index_search_device(const bh_index* h, int64_t n, const float* x, int64_t k, float* distances,
                           int64_t* labels, const bh_search_params* params) {
    if (!h) return fail("null index");
    if (n < 0 || k <= 0) return fail("search: need n >= 0 and k > 0");
    if (n == 0) return 0;
    if (!x || !distances || !labels) return fail("search_device: null buffer");
    if (n > INT32_MAX) return fail("search_device: n too large for one call");
    std::shared_lock<std::shared_mutex> lk(h->rw);
    if (h->ntotal == 0) return fail("search_device: empty index");
    BH_CUDA(cudaSetDevice(h->device));
    if (params && params->sel_bitmap && params->sel_bitmap_bytes < (h->ntotal + 7) / 8)
        return fail("search: sel_bitmap is smaller than (ntotal + 7) / 8 bytes");
    // always the primary context: its lane 0 is the index's own stream (bh_index_stream)
    CtxLease lease{h, acquire_ctx(h, true)};
    if (!lease.c) return 1;
    SearchCtx& c = *lease.c;
    cudaStream_t st = c.lane[0].stream;
    // Each query row is fetched with one bulk (TMA) copy, which needs a 16-byte aligned source. Rows of a
    // d % 4 != 0 index, or a buffer at an odd offset (a tensor view), are first re-laid into an aligned,
    // zero-padded staging buffer on the stream — never handed to the kernel as they are.
    if (h->d != h->dp || (reinterpret_cast<uintptr_t>(x) & 15) != 0) {
        BH_CUDA(c.q_d.reserve((size_t)n * h->dp, st));
        BH_CUDA(copy_rows_padded_async(c.q_d.p, x, n, h->d, h->dp, cudaMemcpyDeviceToDevice, st));
        x = c.q_d.p;
    }
    // Back-to-back calls (a serving loop enqueueing batch after batch) overlap: a batch ends with a drain
    // phase — once its work counter runs out the resident query groups finish one by one and for about one
    // query's latency the SMs are half empty (~13 % of a 10k-query step at efSearch=64). Launched with
    // programmatic stream serialisation, the next batch's CTAs take the freed slots at once. Every other
    // stream operation (copies, events, the caller's kernels, add()) still waits for all earlier launches.
    int* counter = nullptr;
    if (int rc = next_ring_counter(c, st, &counter)) return rc;
    return search_device_impl(h, st, counter, n, x, k, distances, labels, params ? params->stats : nullptr,
                              params, params ? params->sel_bitmap : nullptr, 0, nullptr, true);
}

int bh_index_search(const bh_index* h, int64_t n, const float* x, int64_t k, float* distances,
                    int64_t* labels, const bh_search_params* params) {
    if (!h) return fail("null index");
    if (n < 0 || k <= 0) return fail("search: need n >= 0 and k > 0");
    if (n == 0) return 0;
    if (!x || !distances || !labels) return fail("search: null buffer");
    std::shared_lock<std::shared_mutex> lk(h->rw);
    if (h->ntotal == 0) {  // HNSW::search returns at once on an empty graph: heaps stay (FLT_MAX,-1)
        const float pad = h->metric == BH_METRIC_L2 ? FLT_MAX : -FLT_MAX;
        for (int64_t i = 0; i < n * k; i++) {
            distances[i] = pad;
            labels[i] = -1;
        }
        return 0;
    }
    BH_CUDA(cudaSetDevice(h->device));
    CtxLease lease{h, acquire_ctx(h, false)};
    if (!lease.c) return 1;
    SearchCtx& c = *lease.c;
    SearchLane& l0 = c.lane[0];
    const uint8_t* sel_dev = nullptr;
    if (params && params->sel_bitmap) {
        const int64_t need = (h->ntotal + 7) / 8;
        if (params->sel_bitmap_bytes < need) return fail("search: sel_bitmap is smaller than (ntotal + 7) / 8 bytes");
        BH_CUDA(c.sel_d.reserve((size_t)need, l0.stream));
        BH_CUDA(cudaMemcpyAsync(c.sel_d.p, params->sel_bitmap, (size_t)need, cudaMemcpyHostToDevice, l0.stream));
        BH_CUDA(cudaStreamSynchronize(l0.stream));  // the other lanes read it too
        sel_dev = c.sel_d.p;
    }
    int32_t* stats_host = params ? params->stats : nullptr;

    // (1) Zero-copy path: the caller's buffers are page-locked (cudaHostAlloc / cudaHostRegister, e.g.
    // torch pinned tensors) and the rows are TMA-aligned, so they are device-addressable under UVA: the
    // kernel reads each query straight from host memory with its bulk copy and writes the k results back
    // over PCIe — the transfers overlap the traversal instead of bracketing it.
    if (!stats_host && h->d == h->dp && n <= INT32_MAX) {
        auto dev_ptr = [](const void* p) -> void* {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
        };
        void* xd = dev_ptr(x);
        void* dd = dev_ptr(distances);
        void* ld = dev_ptr(labels);
        if (xd && dd && ld && (reinterpret_cast<uintptr_t>(xd) & 15) == 0) {
            BH_CUDA(cudaEventRecord(c.ev0, l0.stream));
            if (int rc = search_device_impl(h, l0.stream, c.counters.p, n, (const float*)xd, k, (float*)dd,
                                            (int64_t*)ld, nullptr, params, sel_dev))
                return rc;
            BH_CUDA(cudaEventRecord(c.ev1, l0.stream));
            BH_CUDA(cudaStreamSynchronize(l0.stream));
            float ms = 0.f;
            BH_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
            h->last_search_ms = ms;
            return 0;
        }
    }

    // (2) Pageable callers (a faiss drop-in passes plain numpy / malloc memory): the batch is cut into
    // chunks that go round-robin over the context's lanes. Per chunk the CPU copies (and zero-pads) the
    // queries into the lane's page-locked staging, the lane's kernel reads them from there and writes
    // D / I (/ stats) into page-locked result staging, and the CPU copies those out when the lane is
    // next needed or at the end. Staging chunk i+1 overlaps the traversal of chunk i, and the chunks'
    // kernels fill each other's tails; there is no separate H2D / D2H phase.
    const int d = h->d, dp = h->dp;
    // a mid-size batch: one chunk per lane while the staging copy is small beside the traversal (rows up to
    // 512 B), four chunks for wider rows, where a smaller first chunk starts the GPU sooner (measured, 10k
    // queries: 1M x 128 efSearch 64: 2.80 ms with 3 chunks, 2.88 with 4, 3.00 with 6-8, 2.91 with 1;
    // 300k x 960 efSearch 48: 12.9 / 12.4 / 12.5 / 14.1 ms; profiles/README.md)
    const int nchunks = (size_t)dp * sizeof(float) <= 512 ? kLanes : 4;
    int64_t chunk = (n + nchunks - 1) / nchunks;
    chunk = std::max<int64_t>(chunk, 2048);            // small batches: one launch
    chunk = std::min<int64_t>(chunk, 32768);           // bounds the staging memory
    chunk = std::min<int64_t>(chunk, n);
    auto drain = [&](SearchLane& l) -> int {
        if (l.pending_i0 < 0) return 0;
        BH_CUDA(cudaEventSynchronize(l.done));
        std::memcpy(distances + (size_t)l.pending_i0 * k, l.hD.p, (size_t)l.pending_m * k * sizeof(float));
        std::memcpy(labels + (size_t)l.pending_i0 * k, l.hI.p, (size_t)l.pending_m * k * sizeof(int64_t));
        if (stats_host)
            std::memcpy(stats_host + (size_t)l.pending_i0 * 4, l.hS.p, (size_t)l.pending_m * 4 * sizeof(int32_t));
        l.pending_i0 = -1;
        return 0;
    };
    struct PendingGuard {  // an error return must not leave a chunk marked pending for the next call
        SearchCtx& c;
        ~PendingGuard() {
            for (int i = 0; i < kLanes; i++) {
                if (c.lane[i].pending_i0 >= 0) cudaStreamSynchronize(c.lane[i].stream);
                c.lane[i].pending_i0 = -1;
            }
        }
    } guard{c};
    BH_CUDA(cudaEventRecord(c.ev0, l0.stream));
    int li = 0;
    for (int64_t i0 = 0; i0 < n; i0 += chunk, li = (li + 1) % kLanes) {
        const int64_t m = std::min(chunk, n - i0);
        SearchLane& l = c.lane[li];
        if (int rc = drain(l)) return rc;
        BH_CUDA(l.hq.reserve((size_t)chunk * dp));
        BH_CUDA(l.hD.reserve((size_t)chunk * k));
        BH_CUDA(l.hI.reserve((size_t)chunk * k));
        if (stats_host) BH_CUDA(l.hS.reserve((size_t)chunk * 4));
        copy_rows_padded(l.hq.p, x + (size_t)i0 * d, m, d, dp);
        if (int rc = search_device_impl(h, l.stream, c.counters.p + li, m, l.hq.dev, k, l.hD.dev, l.hI.dev,
                                        stats_host ? l.hS.dev : nullptr, params, sel_dev, 0, nullptr, false,
                                        /*solo=*/chunk >= n))
            return rc;
        BH_CUDA(cudaEventRecord(l.done, l.stream));
        l.pending_i0 = i0;
        l.pending_m = m;
    }
    for (int i = 1; i < kLanes; i++)  // device-side span of the call: lane 0 waits for the others
        if (c.lane[i].pending_i0 >= 0) BH_CUDA(cudaStreamWaitEvent(l0.stream, c.lane[i].done, 0));
    BH_CUDA(cudaEventRecord(c.ev1, l0.stream));
    for (int i = 0; i < kLanes; i++)
        if (int rc = drain(c.lane[(li + i) % kLanes])) return rc;  // oldest chunk first
    BH_CUDA(cudaEventSynchronize(c.ev1));
    float ms = 0.f;
    BH_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    h->last_search_ms = ms;
    return 0;
}

int bh_index_reconstruct(const bh_index* h, int64_t key, float* out) {
    if (!h) return fail("null index");
    if (key < 0 || key >= h->ntotal) return fail("reconstruct: key out of range");
    return bh_index_reconstruct_n(h, key, 1, out);
}

int bh_index_reconstruct_n(const bh_index* h, int64_t i0, int64_t ni, float* out) {
    if (!h) return fail("null index");
    if (!out) return fail("reconstruct_n: null buffer");
    std::shared_lock<std::shared_mutex> lk(h->rw);
    if (i0 < 0 || ni < 0 || i0 + ni > h->ntotal) return fail("reconstruct_n: range out of bounds");
    if (ni == 0) return 0;
    BH_CUDA(cudaSetDevice(h->device));
    // own stream-less copies (cudaMemcpy*): reconstruct may run beside searches on the index's stream
    const int d = h->d, dp = h->dp;
    if (!h->half()) {
        BH_CUDA(cudaMemcpy2D(out, (size_t)d * sizeof(float), h->vecs.p + (size_t)i0 * dp, (size_t)dp * sizeof(float),
                             (size_t)d * sizeof(float), (size_t)ni, cudaMemcpyDeviceToHost));
    } else {  // widen on the device (exact), then copy: the caller always sees fp32
        std::lock_guard<std::mutex> cl(h->conv_mu);
        const int64_t chunk = std::min<int64_t>(ni, 1 << 20);
        BH_CUDA(h->conv_d.reserve((size_t)chunk * dp, nullptr));
        const char* src = reinterpret_cast<const char*>(h->vecs.p);
        for (int64_t j0 = 0; j0 < ni; j0 += chunk) {
            const int64_t m = std::min(chunk, ni - j0);
            BH_CUDA(bh::launch_f16_to_f32(src + (size_t)(i0 + j0) * dp * 2, h->conv_d.p, (size_t)m * dp, nullptr, h->fmt16()));
            BH_CUDA(cudaMemcpy2D(out + (size_t)j0 * d, (size_t)d * sizeof(float), h->conv_d.p, (size_t)dp * sizeof(float),
                                 (size_t)d * sizeof(float), (size_t)m, cudaMemcpyDeviceToHost));
        }
    }
    return 0;
}

int bh_selector_range_to_bitmap(int64_t ntotal, int64_t imin, int64_t imax, uint8_t* bitmap) {
    if (ntotal < 0 || !bitmap) return fail("selector: bad arguments");
    std::memset(bitmap, 0, (size_t)((ntotal + 7) / 8));
    for (int64_t i = std::max<int64_t>(imin, 0); i < std::min(imax, ntotal); i++) bitmap[i >> 3] |= (uint8_t)(1u << (i & 7));
    return 0;
}
int bh_selector_batch_to_bitmap(int64_t ntotal, int64_t n, const int64_t* ids, uint8_t* bitmap) {
    if (ntotal < 0 || n < 0 || !bitmap || (n > 0 && !ids)) return fail("selector: bad arguments");
    std::memset(bitmap, 0, (size_t)((ntotal + 7) / 8));
    for (int64_t j = 0; j < n; j++)
        if (ids[j] >= 0 && ids[j] < ntotal) bitmap[ids[j] >> 3] |= (uint8_t)(1u << (ids[j] & 7));
    return 0;
}
int bh_selector_not(int64_t ntotal, uint8_t* bitmap) {
    if (ntotal < 0 || !bitmap) return fail("selector: bad arguments");
    const int64_t nb = (ntotal + 7) / 8;
    for (int64_t b = 0; b < nb; b++) bitmap[b] = (uint8_t)~bitmap[b];
    if (ntotal & 7) bitmap[nb - 1] &= (uint8_t)((1u << (ntotal & 7)) - 1u);  // ids >= ntotal stay non-members
    return 0;
}

int64_t bh_index_ntotal(const bh_index* h) { return h ? h->ntotal : -1; }
int bh_index_d(const bh_index* h) { return h ? h->d : -1; }
int bh_index_M(const bh_index* h) { return h ? h->M : -1; }
int bh_index_metric(const bh_index* h) { return h ? h->metric : -1; }
int bh_index_entry_point(const bh_index* h) { return h ? h->entry_point : -1; }
int bh_index_max_level(const bh_index* h) { return h ? h->max_level : -1; }
int bh_index_get_ef_search(const bh_index* h) { return h ? h->efSearch : -1; }
int bh_index_set_ef_search(bh_index* h, int ef) {
    if (!h) return fail("null index");
    if (ef < 1) return fail("efSearch must be >= 1");
    h->efSearch = ef;
    return 0;
}
int bh_index_get_ef_construction(const bh_index* h) { return h ? h->efConstruction : -1; }
int bh_index_set_ef_construction(bh_index* h, int ef) {
    if (!h) return fail("null index");
    if (ef < 1) return fail("efConstruction must be >= 1");
    h->efConstruction = ef;
    return 0;
}
int bh_index_set_check_relative_distance(bh_index* h, int on) {
    if (!h) return fail("null index");
    h->check_relative_distance = on != 0;
    return 0;
}
int bh_index_set_build_params(bh_index* h, const bh_build_params* p) {
    if (!h || !p) return fail("null argument");
    h->bp = *p;
    return 0;
}

int64_t bh_index_neighbors_size(const bh_index* h) {
    if (!h) return -1;
    int64_t s = 0;
    for (int64_t i = 0; i < h->ntotal; i++) s += h->cum_nn[h->levels[i]];
    return s;
}

int bh_index_export_graph(const bh_index* h, int32_t* levels, uint64_t* offsets, int32_t* neighbors) {
    if (!h) return fail("null index");
    std::shared_lock<std::shared_mutex> lk(h->rw);
    BH_CUDA(cudaSetDevice(h->device));
    const int64_t n = h->ntotal;
    const int deg0 = h->deg0(), M = h->M;
    std::vector<int32_t> l0((size_t)n * deg0), up((size_t)h->n_upper_rows * M);
    if (n) BH_CUDA(cudaMemcpyAsync(l0.data(), h->nbr0.p, l0.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (!up.empty())
        BH_CUDA(cudaMemcpyAsync(up.data(), h->upper_nbr.p, up.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    BH_CUDA(cudaStreamSynchronize(h->stream));
    uint64_t off = 0;
    for (int64_t i = 0; i < n; i++) {
        if (levels) levels[i] = h->levels[i];
        if (offsets) offsets[i] = off;
        if (neighbors) {
            std::memcpy(neighbors + off, l0.data() + (size_t)i * deg0, deg0 * sizeof(int32_t));
            for (int l = 1; l < h->levels[i]; l++)
                std::memcpy(neighbors + off + h->cum_nn[l], up.data() + ((size_t)h->upper_base[i] + l - 1) * M,
                            M * sizeof(int32_t));
        }
        off += h->cum_nn[h->levels[i]];
    }
    if (offsets) offsets[n] = off;
    return 0;
}

int bh_index_import_graph(bh_index* h, int64_t n, const float* x, const int32_t* levels,
                          const int32_t* neighbors, int64_t nneighbors, int entry_point, int max_level) {
    if (!h) return fail("null index");
    if (n <= 0 || !x || !levels || !neighbors) return fail("import: bad arguments");
    if (n > (int64_t)INT32_MAX - 1) return fail("import: too many vectors");
    if (entry_point < 0 || entry_point >= n) return fail("import: entry point out of range");
    std::unique_lock<std::shared_mutex> lk(h->rw);
    BH_CUDA(cudaSetDevice(h->device));
    const int M = h->M, deg0 = h->deg0();
    std::vector<int32_t> ub(n);
    int64_t upper_rows = 0, need = 0;
    for (int64_t i = 0; i < n; i++) {
        if (levels[i] < 1 || levels[i] > (int)h->assign_probas.size()) return fail("import: level out of range");
        ub[i] = levels[i] > 1 ? (int32_t)upper_rows : -1;
        upper_rows += levels[i] - 1;
        need += h->cum_nn[levels[i]];
    }
    if (need != nneighbors) return fail("import: neighbors array size does not match levels");
    if (max_level < 0 || levels[entry_point] - 1 != max_level)
        return fail("import: entry point's level must equal max_level");
    std::vector<int32_t> l0((size_t)n * deg0), up((size_t)upper_rows * M);
    int64_t off = 0;
    for (int64_t i = 0; i < n; i++) {
        std::memcpy(l0.data() + (size_t)i * deg0, neighbors + off, deg0 * sizeof(int32_t));
        for (int l = 1; l < levels[i]; l++)
            std::memcpy(up.data() + ((size_t)ub[i] + l - 1) * M, neighbors + off + h->cum_nn[l], M * sizeof(int32_t));
        off += h->cum_nn[levels[i]];
    }
    for (size_t i = 0; i < l0.size(); i++)
        if (l0[i] < -1 || l0[i] >= n) return fail("import: neighbor id out of range");
    for (size_t i = 0; i < up.size(); i++)
        if (up[i] < -1 || up[i] >= n) return fail("import: neighbor id out of range");
    // all checks passed: only now drop the old contents (a rejected import leaves the index as it was)
    if (int rc = reset_locked(h)) return rc;
    if (int rc = h->ensure_capacity(n, upper_rows)) return rc;
    if (int rc = upload_vectors(h, 0, n, x)) return rc;
    BH_CUDA(cudaMemcpyAsync(h->nbr0.p, l0.data(), l0.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    if (!up.empty())
        BH_CUDA(cudaMemcpyAsync(h->upper_nbr.p, up.data(), up.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    BH_CUDA(cudaMemcpyAsync(h->upper_base_d.p, ub.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    BH_CUDA(cudaMemsetAsync(h->nver0.p, 0, (size_t)n, h->stream));  // imported rows: nothing verified yet
    if (upper_rows) BH_CUDA(cudaMemsetAsync(h->nverU.p, 0, (size_t)upper_rows, h->stream));
    BH_CUDA(cudaStreamSynchronize(h->stream));
    h->levels.assign(levels, levels + n);
    h->upper_base = ub;
    h->n_upper_rows = upper_rows;
    h->ntotal = n;
    h->entry_point = entry_point;
    h->max_level = max_level;
    return 0;
}

void* bh_index_stream(const bh_index* h) { return h ? (void*)h->stream : nullptr; }
int bh_index_synchronize(const bh_index* h) {
    if (!h) return fail("null index");
    BH_CUDA(cudaSetDevice(h->device));
    BH_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}
float bh_index_last_build_ms(const bh_index* h) { return h ? h->last_build_ms : -1.f; }
int bh_index_last_build_counters(const bh_index* h, uint64_t out[6]) {
    if (!h || !out) return fail("null argument");
    for (int i = 0; i < 6; i++) out[i] = h->last_build_counters[i];
    return 0;
}
float bh_index_last_search_ms(const bh_index* h) { return h ? h->last_search_ms.load() : -1.f; }

int bh_merge_topk_device(int nshard, int64_t nq, int64_t k, int metric, const float* D_all,
                         const int64_t* I_all, const int64_t* id_offsets, float* D_out, int64_t* I_out,
                         void* stream) {
    if (nshard < 1 || nshard > bh::kMaxShards || nq < 0 || k <= 0 || k > 4096) return fail("merge: bad shape");
    if (!D_all || !I_all || !id_offsets || !D_out || !I_out) return fail("merge: null buffer");
    bh::ShardOffsets off{};
    for (int i = 0; i < nshard; i++) off.v[i] = id_offsets[i];
    // enqueued on `stream`, no allocation and no synchronisation: the caller's stream orders it
    BH_CUDA(bh::launch_merge_topk(nshard, nq, (int)k, metric == BH_METRIC_L2, D_all, I_all, off, D_out, I_out,
                                  (cudaStream_t)stream));
    return 0;
}

}  // extern "C"

// =================================================================== sharded search (one box)
//
// Replaces faiss::IndexShards(successive_ids = true) + merge_knn_results for the GPUs of
// one box. Rank r owns the contiguous id range [offset_r, offset_r + ntotal_r) and an independent graph.
// A search is collective: every rank calls it with the same queries; each rank's traversal kernel writes
// its k results per query as packed 8-byte keys (distance bits, LOCAL id) straight into every rank's
// gather buffer — peer memory over NVLink — a one-warp kernel then raises this rank's flag in every peer,
// and the merge kernel waits for all flags and merges. No library collective, no host round trip, no
// cross-stream wait: three launches on the index's stream.
struct bh_shards {
    bh_index* local = nullptr;
    int rank = 0, nranks = 1;
    int64_t max_q = 0, max_k = 0;
    unsigned long long* arena = nullptr;  // [kShardRing][nranks][max_q * max_k] packed keys, then the flags
    size_t parity_elems = 0;              // nranks * max_q * max_k (one ring slot)
    unsigned long long* flags = nullptr;  // [nranks * kFlagStride], slot r = last epoch rank r published
    unsigned long long* peer_arena[bh::kMaxPeers] = {};
    bool peer_ipc[bh::kMaxPeers] = {};
    int64_t ntotals[bh::kMaxPeers] = {};
    bool connected = false;
    unsigned long long epoch = 0;
    int* status = nullptr;      // mapped host word: 1 = a peer did not publish within the timeout
    int* status_dev = nullptr;
    int timeout_ms = 20000;
    // pipelined mode: flag + merge kernels run on `xstream`, so consecutive traversal launches stay adjacent on
    // the index's stream and overlap their drain phases; evE[e % ring] = traversal e done, evF = merge e done
    int pipelined = 0;
    cudaStream_t xstream = nullptr;
    cudaEvent_t evE[16] = {}, evF[16] = {};
    std::mutex mu;
};
constexpr int kShardRing = 16;  // gather-buffer ring: call e uses slot e % 16 on every rank

namespace {
struct ShardBlob {  // BH_SHARDS_BLOB_BYTES
    uint32_t magic;
    int32_t rank, nranks, device;
    int64_t pid, ntotal, max_q, max_k;
    uint64_t raw_ptr;
    cudaIpcMemHandle_t ipc;
    char pad[BH_SHARDS_BLOB_BYTES - 4 - 12 - 32 - 8 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(ShardBlob) == BH_SHARDS_BLOB_BYTES, "blob layout");
constexpr uint32_t kBlobMagic = 0x62685348u;

unsigned long long* gather_slot(unsigned long long* arena, size_t parity_elems, int parity, int src_rank,
                                int64_t n, int64_t k) {
    // lists of one call are packed [src_rank][n][k] at the front of the parity's half
    return arena + (size_t)parity * parity_elems + (size_t)src_rank * n * k;  // parity = ring slot
}
}  // namespace

extern "C" {

int bh_shards_create(bh_shards** out, bh_index* local, int rank, int nranks, int64_t max_queries, int64_t max_k) {
    if (!out) return fail("shards_create: out is null");
    *out = nullptr;
    if (!local) return fail("shards_create: null index");
    if (nranks < 1 || nranks > bh::kMaxPeers || rank < 0 || rank >= nranks)
        return fail("shards_create: need 0 <= rank < nranks <= 16");
    if (max_queries < 1 || max_k < 1 || max_k > 4096) return fail("shards_create: bad max_queries / max_k");
    BH_CUDA(cudaSetDevice(local->device));
    std::unique_ptr<bh_shards> s(new bh_shards());
    s->local = local;
    s->rank = rank;
    s->nranks = nranks;
    s->max_q = max_queries;
    s->max_k = max_k;
    s->parity_elems = (size_t)nranks * max_queries * max_k;
    const size_t flag_elems = (size_t)nranks * bh::kFlagStride;
    const size_t bytes = (kShardRing * s->parity_elems + flag_elems) * sizeof(unsigned long long);
    BH_CUDA(cudaMalloc((void**)&s->arena, bytes));  // plain cudaMalloc: exportable with cudaIpcGetMemHandle
    s->flags = s->arena + kShardRing * s->parity_elems;
    cudaError_t e = cudaMemset(s->arena, 0, bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->xstream, cudaStreamNonBlocking);
    for (int i = 0; i < kShardRing && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&s->evE[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->evF[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s->status, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *s->status = 0;
        e = cudaHostGetDevicePointer((void**)&s->status_dev, s->status, 0);
    }
    if (e != cudaSuccess) {
        cudaFree(s->arena);
        if (s->status) cudaFreeHost(s->status);
        if (s->xstream) cudaStreamDestroy(s->xstream);
        for (int i = 0; i < kShardRing; i++) {
            if (s->evE[i]) cudaEventDestroy(s->evE[i]);
            if (s->evF[i]) cudaEventDestroy(s->evF[i]);
        }
        return fail(std::string("shards_create: ") + cudaGetErrorString(e));
    }
    s->peer_arena[rank] = s->arena;
    s->ntotals[rank] = local->ntotal;
    s->connected = nranks == 1;
    *out = s.release();
    return 0;
}

int bh_shards_free(bh_shards* s) {
    if (!s) return 0;
    cudaSetDevice(s->local->device);
    cudaStreamSynchronize(s->local->stream);
    if (s->xstream) {
        cudaStreamSynchronize(s->xstream);
        cudaStreamDestroy(s->xstream);
    }
    for (int i = 0; i < kShardRing; i++) {
        if (s->evE[i]) cudaEventDestroy(s->evE[i]);
        if (s->evF[i]) cudaEventDestroy(s->evF[i]);
    }
    for (int p = 0; p < s->nranks; p++)
        if (s->peer_ipc[p] && s->peer_arena[p]) cudaIpcCloseMemHandle(s->peer_arena[p]);
    if (s->arena) cudaFree(s->arena);
    if (s->status) cudaFreeHost(s->status);
    delete s;
    return 0;
}

int bh_shards_export(bh_shards* s, void* blob) {
    if (!s || !blob) return fail("shards_export: null argument");
    BH_CUDA(cudaSetDevice(s->local->device));
    ShardBlob b{};
    b.magic = kBlobMagic;
    b.rank = s->rank;
    b.nranks = s->nranks;
    b.device = s->local->device;
    b.pid = (int64_t)getpid();
    b.ntotal = s->local->ntotal;
    b.max_q = s->max_q;
    b.max_k = s->max_k;
    b.raw_ptr = (uint64_t)(uintptr_t)s->arena;
    BH_CUDA(cudaIpcGetMemHandle(&b.ipc, s->arena));
    std::memcpy(blob, &b, sizeof(b));
    return 0;
}

int bh_shards_connect(bh_shards* s, const void* blobs) {
    if (!s || !blobs) return fail("shards_connect: null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    BH_CUDA(cudaSetDevice(s->local->device));
    const ShardBlob* b = static_cast<const ShardBlob*>(blobs);
    for (int p = 0; p < s->nranks; p++) {
        if (b[p].magic != kBlobMagic || b[p].rank != p || b[p].nranks != s->nranks)
            return fail("shards_connect: blob " + std::to_string(p) + " is not rank " + std::to_string(p) + "'s export");
        if (b[p].max_q != s->max_q || b[p].max_k != s->max_k)
            return fail("shards_connect: ranks were created with different max_queries / max_k");
    }
    for (int p = 0; p < s->nranks; p++) {
        s->ntotals[p] = b[p].ntotal;
        if (p == s->rank || s->peer_arena[p]) continue;  // re-connect only refreshes the shard sizes
        if (b[p].pid == (int64_t)getpid()) {  // same process: the pointer itself, with peer access if needed
            if (b[p].device != s->local->device) {
                int can = 0;
                BH_CUDA(cudaDeviceCanAccessPeer(&can, s->local->device, b[p].device));
                if (!can) return fail("shards_connect: no peer access between the two devices");
                const cudaError_t e = cudaDeviceEnablePeerAccess(b[p].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
            s->peer_arena[p] = reinterpret_cast<unsigned long long*>((uintptr_t)b[p].raw_ptr);
        } else {
            void* ptr = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&ptr, b[p].ipc, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(std::string("shards_connect: cudaIpcOpenMemHandle(rank ") + std::to_string(p) +
                            "): " + cudaGetErrorString(e));
            }
            s->peer_arena[p] = static_cast<unsigned long long*>(ptr);
            s->peer_ipc[p] = true;
        }
    }
    s->connected = true;
    return 0;
}

// first half: traverse the local shard, publish the packed lists (to every connected rank, or only to
// this rank's own buffer when `publish_to_peers` is 0 — the caller then moves them, see bh_shards_gather)
int bh_shards_post(bh_shards* s, int64_t n, const float* x, int64_t k, const bh_search_params* params,
                   int publish_to_peers) {
    if (!s) return fail("null shards handle");
    const bh_index* h = s->local;
    if (n < 1 || n > s->max_q || k < 1 || k > s->max_k) return fail("shards_post: n / k exceed max_queries / max_k");
    if (!x) return fail("shards_post: null queries");
    if (publish_to_peers && !s->connected) return fail("shards_post: not connected (bh_shards_connect)");
    if (params && (params->stats || params->sel_bitmap)) return fail("shards_post: stats / selectors are per shard; not supported here");
    std::lock_guard<std::mutex> slk(s->mu);
    std::shared_lock<std::shared_mutex> lk(h->rw);
    BH_CUDA(cudaSetDevice(h->device));
    if (*s->status) return fail("sharded search: a peer did not publish its results in time (earlier call)");
    CtxLease lease{h, acquire_ctx(h, true)};
    if (!lease.c) return 1;
    SearchCtx& c = *lease.c;
    cudaStream_t st = c.lane[0].stream;
    s->epoch++;
    const int parity = (int)(s->epoch % kShardRing);
    const bool piped = s->pipelined && publish_to_peers;
    // Pipelined flow control: rank A's call e stores into slot e % R of every peer, which that peer's merge
    // e-R must have finished reading. Every R/2-th call this stream waits for its OWN merge of R/2 calls
    // earlier (call e'): that merge needed every rank's flag e'-R/2, each raised (in order, on that rank's
    // exchange stream) after the rank's merge e'-R/2-1 — so all merges up to e'-R/2-1 >= e-R are done for
    // the calls e', ..., e'+R/2-1. (The wait breaks the launch adjacency once per R/2 calls.)
    constexpr int kHalfRing = kShardRing / 2;
    if (piped && s->epoch % kHalfRing == 0 && s->epoch >= (unsigned)kShardRing)
        BH_CUDA(cudaStreamWaitEvent(st, s->evF[(s->epoch - kHalfRing) % kShardRing], 0));
    unsigned long long* outs[bh::kMaxPeers];
    int n_out = 0;
    if (publish_to_peers) {
        for (int p = 0; p < s->nranks; p++)
            outs[n_out++] = gather_slot(s->peer_arena[p], s->parity_elems, parity, s->rank, n, k);
    } else {
        outs[n_out++] = gather_slot(s->arena, s->parity_elems, parity, s->rank, n, k);
    }
    if (h->ntotal == 0) {  // an empty shard publishes empty lists
        for (int p = 0; p < n_out; p++) BH_CUDA(cudaMemsetAsync(outs[p], 0xFF, (size_t)n * k * 8, st));
    } else {
        if (h->d != h->dp || (reinterpret_cast<uintptr_t>(x) & 15) != 0) {
            BH_CUDA(c.q_d.reserve((size_t)n * h->dp, st));
            BH_CUDA(copy_rows_padded_async(c.q_d.p, x, n, h->d, h->dp, cudaMemcpyDeviceToDevice, st));
            x = c.q_d.p;
        }
        int* counter = c.counters.p;
        if (piped)
            if (int rc = next_ring_counter(c, st, &counter)) return rc;
        if (int rc = search_device_impl(h, st, counter, n, x, k, nullptr, nullptr, nullptr, params, nullptr, n_out, outs,
                                        piped))
            return rc;
    }
    cudaStream_t xs = st;
    if (piped) {  // flag (and later the merge) on the exchange stream, behind this traversal only
        BH_CUDA(cudaEventRecord(s->evE[parity], st));
        BH_CUDA(cudaStreamWaitEvent(s->xstream, s->evE[parity], 0));
        xs = s->xstream;
    }
    if (publish_to_peers && s->nranks > 1) {
        bh::PeerFlags pf{};
        for (int p = 0; p < s->nranks; p++) pf.v[p] = s->peer_arena[p] + kShardRing * s->parity_elems;
        BH_CUDA(bh::launch_shard_signal(s->nranks, s->rank, pf, s->epoch, xs));
    }
    return 0;
}

// device pointer of the current call's gather buffer, [nranks][n][k] packed keys (this rank's slice at index
// `rank`): a caller that exchanges with its own collective (e.g. one in-place NCCL all-gather of n*k*8 bytes
// per rank) fills the other slices, then calls bh_shards_collect with wait_for_peers = 0
int bh_shards_gather(bh_shards* s, int64_t n, int64_t k, void** ptr) {
    if (!s || !ptr) return fail("null argument");
    if (n < 1 || n > s->max_q || k < 1 || k > s->max_k) return fail("shards_gather: n / k exceed max_queries / max_k");
    std::lock_guard<std::mutex> lk(s->mu);
    *ptr = gather_slot(s->arena, s->parity_elems, (int)(s->epoch % kShardRing), 0, n, k);
    return 0;
}

// second half: wait for every rank's lists of the current call (wait_for_peers), merge, write D / I with
// global ids (local id + the owning shard's offset, successive_ids)
int bh_shards_collect(bh_shards* s, int64_t n, int64_t k, float* distances, int64_t* labels, int wait_for_peers) {
    if (!s) return fail("null shards handle");
    const bh_index* h = s->local;
    if (n < 1 || n > s->max_q || k < 1 || k > s->max_k) return fail("shards_collect: n / k exceed max_queries / max_k");
    if (!distances || !labels) return fail("shards_collect: null buffer");
    std::lock_guard<std::mutex> slk(s->mu);
    BH_CUDA(cudaSetDevice(h->device));
    bh::ShardOffsets off{};
    int64_t acc = 0;
    for (int p = 0; p < s->nranks; p++) {
        off.v[p] = acc;
        acc += s->ntotals[p];
    }
    const bool wait = wait_for_peers && s->nranks > 1;
    const int slot = (int)(s->epoch % kShardRing);
    const bool piped = s->pipelined && wait_for_peers;
    BH_CUDA(bh::launch_merge_packed(s->nranks, n, (int)k, h->metric == BH_METRIC_L2,
                                    gather_slot(s->arena, s->parity_elems, slot, 0, n, k), off,
                                    distances, labels, wait ? s->flags : nullptr, s->epoch, s->status_dev,
                                    s->timeout_ms, piped ? s->xstream : h->stream));
    if (piped) BH_CUDA(cudaEventRecord(s->evF[slot], s->xstream));
    return 0;
}

// Pipelined mode (off by default): the flag and merge kernels of a search go to a separate exchange stream, so
// the index's stream carries nothing but traversal launches, which then overlap their drain phases call after
// call and no longer wait for slower peers' flags. D / I of a call are complete only after
// bh_shards_join. Pass different D / I buffers to calls whose results you have not joined yet — or the same
// ones if only the last call's results matter (merges run in call order). Join before switching the mode.
int bh_shards_set_pipelined(bh_shards* s, int on) {
    if (!s) return fail("null shards handle");
    std::lock_guard<std::mutex> lk(s->mu);
    s->pipelined = on ? 1 : 0;
    return 0;
}
// make `stream` (cudaStream_t as void*; NULL = the local index's stream) wait for the most recent call's merge
int bh_shards_join(bh_shards* s, void* stream) {
    if (!s) return fail("null shards handle");
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->pipelined || s->epoch == 0) return 0;
    BH_CUDA(cudaSetDevice(s->local->device));
    BH_CUDA(cudaStreamWaitEvent(stream ? (cudaStream_t)stream : s->local->stream, s->evF[s->epoch % kShardRing], 0));
    return 0;
}

// this rank's own lists of the current call as (D, I) with LOCAL ids (device buffers [n][k])
int bh_shards_local_lists(bh_shards* s, int64_t n, int64_t k, float* distances, int64_t* labels) {
    if (!s || !distances || !labels) return fail("null argument");
    if (n < 1 || n > s->max_q || k < 1 || k > s->max_k) return fail("shards_local_lists: n / k exceed max_queries / max_k");
    const bh_index* h = s->local;
    std::lock_guard<std::mutex> lk(s->mu);
    BH_CUDA(cudaSetDevice(h->device));
    BH_CUDA(bh::launch_unpack(gather_slot(s->arena, s->parity_elems, (int)(s->epoch % kShardRing), s->rank, n, k), n * k,
                              h->metric == BH_METRIC_L2, distances, labels, h->stream));
    return 0;
}

// the collective call: post + collect. x / distances / labels are device pointers on this rank's GPU;
// enqueued on the local index's stream, returns without synchronising.
int bh_shards_search_device(bh_shards* s, int64_t n, const float* x, int64_t k, float* distances, int64_t* labels,
                            const bh_search_params* params) {
    if (!s) return fail("null shards handle");
    if (n < 0 || k <= 0) return fail("search: need n >= 0 and k > 0");
    for (int64_t i0 = 0; i0 < n; i0 += s->max_q) {  // batches beyond max_queries: one exchange per slice
        const int64_t m = std::min<int64_t>(s->max_q, n - i0);
        if (int rc = bh_shards_post(s, m, x + (size_t)i0 * s->local->d, k, params, 1)) return rc;
        if (int rc = bh_shards_collect(s, m, k, distances + (size_t)i0 * k, labels + (size_t)i0 * k, 1)) return rc;
    }
    return 0;
}

int bh_shards_set_ntotals(bh_shards* s, const int64_t* ntotals) {
    if (!s || !ntotals) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    for (int p = 0; p < s->nranks; p++) s->ntotals[p] = ntotals[p];
    return 0;
}

int bh_shards_status(bh_shards* s) { return s ? *s->status : -1; }

}  // extern "C"
