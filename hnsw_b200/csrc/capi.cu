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
#include <mutex>
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
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) return e;
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

int ceil_log2(long long v) {
    int b = 0;
    while ((1ll << b) < v) b++;
    return b;
}

}  // namespace

struct bh_index {
    int d = 0, M = 0, metric = BH_METRIC_L2, device = 0;
    int storage = BH_STORAGE_F32;  // BH_STORAGE_F16: rows held as fp16 (opt-in)
    int efSearch = 16, efConstruction = 40;  // faiss HNSW defaults (App. A.1)
    bool check_relative_distance = true;
    bh_build_params bp{0, 0, 0, 0};
    mutable std::mutex mu;  // one operation at a time per handle (stream, counter and staging are shared)
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
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int num_sms = 148;
    size_t smem_optin = 227 * 1024;
    DevBuf<int> counter;
    DevBuf<unsigned long long> build_counters;  // [6], see bh_index_last_build_counters
    unsigned long long last_build_counters[6] = {0, 0, 0, 0, 0, 0};
    // search staging
    mutable DevBuf<float> q_d, D_d;
    mutable DevBuf<int64_t> I_d;
    mutable DevBuf<int32_t> stats_d;
    mutable DevBuf<uint8_t> sel_d;
    mutable float last_search_ms = 0.f;
    float last_build_ms = 0.f;
    // build scratch
    DevBuf<int4> items_d;
    DevBuf<unsigned long long> cand_lists;
    DevBuf<int32_t> cand_counts, e_slot, e_src, e_dst, e_level, e_next;
    DevBuf<float> e_dist;

    int deg0() const { return 2 * M; }
    bool half() const { return storage == BH_STORAGE_F16; }
    int row_floats() const { return half() ? d / 2 : d; }  // stored row size in 4-byte units
    mutable DevBuf<float> conv_d;                           // fp32 staging for fp16 conversion

    bh::GraphView view() const {
        bh::GraphView g;
        g.vecs = vecs.p;
        g.nbr0 = nbr0.p;
        g.upper_base = upper_base_d.p;
        g.upper_nbr = upper_nbr.p;
        g.d = d;
        g.nchunk = row_floats() / 4;
        g.half = half() ? 1 : 0;
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

    // Visited-hash slots per query. The table is "forgetful" (beam.cuh), so its size is a pure
    // performance knob: small tables keep many queries resident per SM at the price of a few
    // re-scored vertices. Minimum: 3/4 of the slots must hold the ef-list plus one full row.
    int min_hash_bits(int ef) const { return std::max(8, ceil_log2(((long long)(ef + deg0()) * 4 + 2) / 3 + 1)); }
    int auto_hash_bits(int ef, int req) const {
        const int lo = min_hash_bits(ef);
        if (req > 0) return std::min(std::max(req, lo), 16);
        // measured on 1M x 128 (profiles/README.md): ~4 slots per list entry is the sweet spot
        int b = std::min(ceil_log2((long long)ef * 4), 10);
        b = std::max(b, 9);
        return std::min(std::max(b, lo), 15);
    }
    // Warps cooperating on one query. With enough work items to fill the GPU, one warp per query
    // keeps the most queries in flight (throughput regime). With few items (small query batches,
    // the early construction rounds) the machine is idle and a hop's latency is what matters:
    // W warps score a hop's ~50 vectors in one gather round instead of 3-6 serial ones.
    int auto_warps(int ef, int hash_bits, int req, long long n_items) const {
        if (req == 1 || req == 2 || req == 4 || req == 8) return req;
        const size_t s = bh::beam_group_smem(d, ef, hash_bits, deg0());
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
        const size_t gs = bh::beam_group_smem(d, ef, hash_bits, deg0());
        if (24 * gs <= smem_optin - 6 * 1024) return 1;
        if (20 * gs <= smem_optin - 5 * 1024) return 3;  // 5 CTAs/SM, <=96 regs
        return 0;
    }

    int ensure_capacity(int64_t n_new_total, int64_t upper_rows_total) {
        const int64_t old_n = ntotal;
        const int rf = row_floats();
        if ((size_t)n_new_total > (size_t)(vecs.cap / rf)) {
            int64_t cap = std::max<int64_t>(n_new_total, (int64_t)(vecs.cap / rf) * 3 / 2);
            BH_CUDA(vecs.reserve((size_t)cap * rf, stream, true, (size_t)old_n * rf));
            BH_CUDA(nbr0.reserve((size_t)cap * deg0(), stream, true, (size_t)old_n * deg0()));
            BH_CUDA(upper_base_d.reserve((size_t)cap, stream, true, (size_t)old_n));
            BH_CUDA(nver0.reserve((size_t)cap, stream, true, (size_t)old_n));
        }
        if ((size_t)upper_rows_total * M > upper_nbr.cap) {
            size_t cap = std::max<size_t>((size_t)upper_rows_total * M, upper_nbr.cap * 3 / 2);
            BH_CUDA(upper_nbr.reserve(cap, stream, true, (size_t)n_upper_rows * M));
            BH_CUDA(nverU.reserve(cap / M + 1, stream, true, (size_t)n_upper_rows));
        }
        // pending-list heads: one per adjacency row, all -1 between batches
        const int64_t level0_rows = (int64_t)(vecs.cap / row_floats());
        const size_t need = (size_t)level0_rows + upper_nbr.cap / M + 1;
        if (need > slot_head.cap || level0_rows != slot_level0) {
            BH_CUDA(slot_head.reserve(need, stream));
            BH_CUDA(cudaMemsetAsync(slot_head.p, 0xFF, slot_head.cap * sizeof(int32_t), stream));
            slot_level0 = level0_rows;
        }
        return 0;
    }

    void free_all() {
        nver0.release(); nverU.release();
        vecs.release(); nbr0.release(); upper_base_d.release(); upper_nbr.release(); slot_head.release();
        sel_d.release(); conv_d.release(); build_counters.release();
        counter.release(); q_d.release(); D_d.release(); I_d.release(); stats_d.release();
        items_d.release(); cand_lists.release(); cand_counts.release();
        e_slot.release(); e_src.release(); e_dst.release(); e_level.release(); e_next.release(); e_dist.release();
    }
};

namespace {

// storage->add: rows [n0, n0+n) into HBM, converting to fp16 on the device when that storage is on
int upload_vectors(bh_index* h, int64_t n0, int64_t n, const float* x) {
    const int d = h->d;
    if (!h->half()) {
        BH_CUDA(cudaMemcpyAsync(h->vecs.p + (size_t)n0 * d, x, (size_t)n * d * sizeof(float),
                                cudaMemcpyHostToDevice, h->stream));
        return 0;
    }
    const int64_t chunk = std::min<int64_t>(n, 1 << 20);
    BH_CUDA(h->conv_d.reserve((size_t)chunk * d, h->stream));
    char* dst = reinterpret_cast<char*>(h->vecs.p);
    for (int64_t i0 = 0; i0 < n; i0 += chunk) {
        const int64_t m = std::min(chunk, n - i0);
        BH_CUDA(cudaMemcpyAsync(h->conv_d.p, x + (size_t)i0 * d, (size_t)m * d * sizeof(float),
                                cudaMemcpyHostToDevice, h->stream));
        BH_CUDA(bh::launch_f32_to_f16(h->conv_d.p, dst + (size_t)(n0 + i0) * d * 2, (size_t)m * d, h->stream));
    }
    return 0;
}

int search_device_impl(const bh_index* h, int64_t n, const float* xq_d, int64_t k, float* D_d,
                       int64_t* I_d, int32_t* stats_d, const bh_search_params* params,
                       const uint8_t* sel_dev = nullptr) {
    const int efS = (params && params->efSearch > 0) ? params->efSearch : h->efSearch;
    bool crd = h->check_relative_distance;
    if (params && params->check_relative_distance == 1) crd = true;
    if (params && params->check_relative_distance == 2) crd = false;
    const int ef = (int)std::max<int64_t>(efS, k);
    if (ef > 4096) return fail("max(efSearch, k) > 4096 is not supported");
    const int rk = sel_dev ? (int)k : 0;  // selector-filtered result list lives beside the candidate list
    const int hb = h->auto_hash_bits(ef + rk, params ? params->hash_bits : 0);
    int W = h->auto_warps(ef + rk, hb, params ? params->warps_per_query : 0, n);
    int G = W >= 4 ? 1 : 4 / W;
    while (G * bh::beam_group_smem(h->d, ef, hb, h->deg0(), rk) > h->smem_optin && W < 4) {
        W *= 2;
        G = W >= 4 ? 1 : 4 / W;
    }
    if (G * bh::beam_group_smem(h->d, ef, hb, h->deg0(), rk) > h->smem_optin)
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
    t.stats = stats_d;
    t.sel = sel_dev;
    t.counter = h->counter.p;
    BH_CUDA(cudaMemsetAsync(h->counter.p, 0, sizeof(int), h->stream));
    BH_CUDA(bh::launch_beam(h->view(), t, W, h->beam_variant(ef + rk, hb), h->num_sms, h->stream, nullptr));
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
    const int d = h->d, M = h->M, deg0 = h->deg0();
    // Everything that can reject the call is checked BEFORE the level RNG advances or any state
    // changes, so a failed add() leaves the index exactly as it was.
    const int efc = h->efConstruction;
    if (efc < 1 || efc > 4096) return fail("efConstruction must be in [1, 4096]");
    const int hb = h->auto_hash_bits(efc, h->bp.hash_bits);
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
    const int divisor = auto_div ? 64 : h->bp.batch_divisor;
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
    if (items.empty()) return 0;
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
            t.stats = nullptr;
            t.build_counters = h->build_counters.p;
            t.counter = h->counter.p;
            BH_CUDA(cudaMemsetAsync(h->counter.p, 0, sizeof(int), h->stream));
            const int W = h->auto_warps(efc, hb, h->bp.warps_per_query, n_items);
            BH_CUDA(bh::launch_beam(g, t, W, h->beam_variant(efc, hb), h->num_sms, h->stream, nullptr));
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
            b.max_special = std::min(8, std::max(2, (16 * 1024) / (4 * h->row_floats())));
            b.build_counters = h->build_counters.p;
            BH_CUDA(bh::launch_select_and_link(g, b, h->num_sms, h->stream));
            BH_CUDA(bh::launch_backlinks(g, b, h->num_sms, h->stream));
            bh::count_launch(3);
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
    if (d <= 0 || d % 4 != 0 || d > 2048) return fail("create: d must be a multiple of 4 in [4, 2048]");
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
    std::lock_guard<std::mutex> lk(h->mu);
    return reset_locked(h);
}

int bh_index_set_vector_storage(bh_index* h, int storage) {
    if (!h) return fail("null index");
    if (storage != BH_STORAGE_F32 && storage != BH_STORAGE_F16) return fail("unknown storage kind");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->ntotal != 0) return fail("set_vector_storage: the index is not empty");
    if (storage == BH_STORAGE_F16 && h->d % 8 != 0) return fail("fp16 storage needs d % 8 == 0");
    if (storage != h->storage) {  // capacities are counted in rows of the old width: start over
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        h->vecs.release();
        h->nbr0.release();
        h->upper_base_d.release();
        h->nver0.release();
        h->slot_level0 = 0;
    }
    h->storage = storage;
    return 0;
}
int bh_index_get_vector_storage(const bh_index* h) { return h ? h->storage : -1; }

int bh_index_train(bh_index* h, int64_t, const float*) {
    if (!h) return fail("null index");
    return 0;  // IndexFlat storage needs no training; is_trained is always true
}

int bh_index_add(bh_index* h, int64_t n, const float* x) {
    if (!h) return fail("null index");
    std::lock_guard<std::mutex> lk(h->mu);
    return add_impl(h, n, x, nullptr, nullptr);
}

int bh_index_add_ex(bh_index* h, int64_t n, const float* x, const int32_t* levels, const int32_t* order) {
    if (!h) return fail("null index");
    std::lock_guard<std::mutex> lk(h->mu);
    return add_impl(h, n, x, levels, order);
}

int bh_index_search_device(const bh_index* h, int64_t n, const float* x, int64_t k, float* distances,
                           int64_t* labels, const bh_search_params* params) {
    if (!h) return fail("null index");
    if (n < 0 || k <= 0) return fail("search: need n >= 0 and k > 0");
    if (n == 0) return 0;
    if (h->ntotal == 0) return fail("search_device: empty index");
    if (n > INT32_MAX) return fail("search_device: n too large for one call");
    std::lock_guard<std::mutex> lk(h->mu);
    BH_CUDA(cudaSetDevice(h->device));
    if (params && params->sel_bitmap && params->sel_bitmap_bytes < (h->ntotal + 7) / 8)
        return fail("search: sel_bitmap is smaller than (ntotal + 7) / 8 bytes");
    return search_device_impl(h, n, x, k, distances, labels, params ? params->stats : nullptr, params,
                              params ? params->sel_bitmap : nullptr);
}

int bh_index_search(const bh_index* h, int64_t n, const float* x, int64_t k, float* distances,
                    int64_t* labels, const bh_search_params* params) {
    if (!h) return fail("null index");
    if (n < 0 || k <= 0) return fail("search: need n >= 0 and k > 0");
    if (n == 0) return 0;
    if (!x || !distances || !labels) return fail("search: null buffer");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->ntotal == 0) {  // HNSW::search returns at once on an empty graph: heaps stay (FLT_MAX,-1)
        const float pad = h->metric == BH_METRIC_L2 ? FLT_MAX : -FLT_MAX;
        for (int64_t i = 0; i < n * k; i++) {
            distances[i] = pad;
            labels[i] = -1;
        }
        return 0;
    }
    BH_CUDA(cudaSetDevice(h->device));
    const uint8_t* sel_dev = nullptr;
    if (params && params->sel_bitmap) {
        const int64_t need = (h->ntotal + 7) / 8;
        if (params->sel_bitmap_bytes < need) return fail("search: sel_bitmap is smaller than (ntotal + 7) / 8 bytes");
        BH_CUDA(h->sel_d.reserve((size_t)need, h->stream));
        BH_CUDA(cudaMemcpyAsync(h->sel_d.p, params->sel_bitmap, (size_t)need, cudaMemcpyHostToDevice, h->stream));
        sel_dev = h->sel_d.p;
    }
    // Zero-copy path: when the caller's buffers are page-locked (cudaHostAlloc / cudaHostRegister,
    // e.g. torch pinned tensors) they are device-addressable under UVA, so the kernel reads each
    // query straight from host memory with its TMA bulk copy and writes the k results back over
    // PCIe — the transfers overlap the traversal instead of bracketing it.
    if (n <= INT32_MAX && !(params && params->stats)) {
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
        if (xd && dd && ld) {
            BH_CUDA(cudaEventRecord(h->ev0, h->stream));
            if (int rc = search_device_impl(h, n, (const float*)xd, k, (float*)dd, (int64_t*)ld, nullptr, params, sel_dev))
                return rc;
            BH_CUDA(cudaEventRecord(h->ev1, h->stream));
            BH_CUDA(cudaStreamSynchronize(h->stream));
            BH_CUDA(cudaEventElapsedTime(&h->last_search_ms, h->ev0, h->ev1));
            return 0;
        }
    }
    const int64_t chunk = 1 << 18;
    const int64_t nb = std::min(n, chunk);
    BH_CUDA(h->q_d.reserve((size_t)nb * h->d, h->stream));
    BH_CUDA(h->D_d.reserve((size_t)nb * k, h->stream));
    BH_CUDA(h->I_d.reserve((size_t)nb * k, h->stream));
    int32_t* stats_host = params ? params->stats : nullptr;
    if (stats_host) BH_CUDA(h->stats_d.reserve((size_t)nb * 4, h->stream));
    float total_ms = 0.f;
    for (int64_t i0 = 0; i0 < n; i0 += chunk) {
        const int64_t m = std::min(chunk, n - i0);
        BH_CUDA(cudaMemcpyAsync(h->q_d.p, x + (size_t)i0 * h->d, (size_t)m * h->d * sizeof(float),
                                cudaMemcpyHostToDevice, h->stream));
        BH_CUDA(cudaEventRecord(h->ev0, h->stream));
        if (int rc = search_device_impl(h, m, h->q_d.p, k, h->D_d.p, h->I_d.p,
                                        stats_host ? h->stats_d.p : nullptr, params, sel_dev))
            return rc;
        BH_CUDA(cudaEventRecord(h->ev1, h->stream));
        BH_CUDA(cudaMemcpyAsync(distances + (size_t)i0 * k, h->D_d.p, (size_t)m * k * sizeof(float),
                                cudaMemcpyDeviceToHost, h->stream));
        BH_CUDA(cudaMemcpyAsync(labels + (size_t)i0 * k, h->I_d.p, (size_t)m * k * sizeof(int64_t),
                                cudaMemcpyDeviceToHost, h->stream));
        if (stats_host)
            BH_CUDA(cudaMemcpyAsync(stats_host + (size_t)i0 * 4, h->stats_d.p, (size_t)m * 4 * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, h->stream));
        BH_CUDA(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        BH_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        total_ms += ms;
    }
    h->last_search_ms = total_ms;
    return 0;
}

int bh_index_reconstruct(const bh_index* h, int64_t key, float* out) {
    if (!h) return fail("null index");
    if (key < 0 || key >= h->ntotal) return fail("reconstruct: key out of range");
    return bh_index_reconstruct_n(h, key, 1, out);
}

int bh_index_reconstruct_n(const bh_index* h, int64_t i0, int64_t ni, float* out) {
    if (!h) return fail("null index");
    if (i0 < 0 || ni < 0 || i0 + ni > h->ntotal) return fail("reconstruct_n: range out of bounds");
    if (ni == 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    BH_CUDA(cudaSetDevice(h->device));
    if (!h->half()) {
        BH_CUDA(cudaMemcpyAsync(out, h->vecs.p + (size_t)i0 * h->d, (size_t)ni * h->d * sizeof(float),
                                cudaMemcpyDeviceToHost, h->stream));
    } else {  // widen on the device (exact), then copy: the caller always sees fp32
        const int64_t chunk = std::min<int64_t>(ni, 1 << 20);
        BH_CUDA(h->conv_d.reserve((size_t)chunk * h->d, h->stream));
        const char* src = reinterpret_cast<const char*>(h->vecs.p);
        for (int64_t j0 = 0; j0 < ni; j0 += chunk) {
            const int64_t m = std::min(chunk, ni - j0);
            BH_CUDA(bh::launch_f16_to_f32(src + (size_t)(i0 + j0) * h->d * 2, h->conv_d.p, (size_t)m * h->d, h->stream));
            BH_CUDA(cudaMemcpyAsync(out + (size_t)j0 * h->d, h->conv_d.p, (size_t)m * h->d * sizeof(float),
                                    cudaMemcpyDeviceToHost, h->stream));
            BH_CUDA(cudaStreamSynchronize(h->stream));
        }
    }
    BH_CUDA(cudaStreamSynchronize(h->stream));
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
    std::lock_guard<std::mutex> lk(h->mu);
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
float bh_index_last_search_ms(const bh_index* h) { return h ? h->last_search_ms : -1.f; }

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
