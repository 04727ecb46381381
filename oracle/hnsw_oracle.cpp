// oracle/hnsw_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the semantics of faiss::IndexHNSWFlat (train / add / search),
// which is the implementation the reference repo points at
// (/root/reference/README.md:2 "The project is based on faiss and optimized for HNSW").
//
// PARITY UNPINNED: the reference mount holds no source, no tests and no golden
// vectors, faiss itself is not installed and cannot be fetched, and the reference
// pins no faiss version. This file therefore follows upstream faiss *as specified in
// SURVEY.md Appendix A* (files named there: faiss/impl/HNSW.{h,cpp},
// faiss/IndexHNSW.cpp, faiss/utils/Heap.h, faiss/utils/random.cpp) and each function
// below cites the appendix paragraph it restates. It is pinned only by
//   (1) std::mt19937 known answers (standard-defined),
//   (2) brute-force exact kNN (recall), and
//   (3) structural invariants of the built graph, and
//   (4) a 7-vertex graph worked by hand from the published algorithm (neighbour rows,
//       visit order, results) that the oracle and the GPU must both reproduce
// — see tests/test_oracle.py and tests/test_handworked_golden.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library. The product (hnsw_b200/) never does.
//
// Distance arithmetic has two modes (orc_set_team):
//   team == 0 : "native" — 16 independent partial sums, auto-vectorised (AVX-512 here);
//               this is the mode the CPU baseline is timed in.
//   team == T : emulates, bit for bit, the summation order of the CUDA kernels
//               (hnsw_b200/csrc/beam.cuh, build_kernels.cu): lane j of a T-lane team owns the float4
//               chunks j, j+T, j+2T…, accumulates them with one fmaf chain, then the
//               team is reduced by an xor-butterfly. Used by the parity tests so that
//               oracle and GPU traverse the same graph along the same path.

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <queue>
#include <random>
#include <string>
#include <vector>

#include <omp.h>

namespace {

using idx_t = int64_t;
using storage_idx_t = int32_t;

constexpr int METRIC_INNER_PRODUCT = 0;  // faiss/MetricType.h values
constexpr int METRIC_L2 = 1;

// ---------------------------------------------------------------- distances (A.3)

inline float l2_native(const float* a, const float* b, int d) {
    float acc[16] = {0};
    int i = 0;
    for (; i + 16 <= d; i += 16)
        for (int j = 0; j < 16; j++) {
            float t = a[i + j] - b[i + j];
            acc[j] += t * t;
        }
    float s = 0;
    for (int j = 0; j < 16; j++) s += acc[j];
    for (; i < d; i++) {
        float t = a[i] - b[i];
        s += t * t;
    }
    return s;
}

inline float ip_native(const float* a, const float* b, int d) {
    float acc[16] = {0};
    int i = 0;
    for (; i + 16 <= d; i += 16)
        for (int j = 0; j < 16; j++) acc[j] += a[i + j] * b[i + j];
    float s = 0;
    for (int j = 0; j < 16; j++) s += acc[j];
    for (; i < d; i++) s += a[i] * b[i];
    return s;
}

// Bit-exact emulation of the CUDA team reduction (see header comment).
// CE = elements per 16-byte chunk of the stored vector: 4 for fp32 storage, 8 for fp16 storage.
template <bool L2>
inline float team_order(const float* a, const float* b, int d, int T, int CE = 4) {
    float lane[32];
    const int nchunk = d / CE;
    for (int j = 0; j < T; j++) {
        float acc = 0.f;
        for (int c = j; c < nchunk; c += T) {
            for (int e = 0; e < CE; e++) {
                float x = a[CE * c + e], y = b[CE * c + e];
                if (L2) {
                    float t = x - y;
                    acc = std::fmaf(t, t, acc);
                } else {
                    acc = std::fmaf(x, y, acc);
                }
            }
        }
        lane[j] = acc;
    }
    for (int off = T / 2; off >= 1; off >>= 1) {
        float nxt[32];
        for (int j = 0; j < T; j++) nxt[j] = lane[j] + lane[j ^ off];
        std::memcpy(lane, nxt, sizeof(float) * T);
    }
    return lane[0];
}

// Round to IEEE binary16 (round-to-nearest-even) and back, like __float2half_rn on the device.
inline float round_to_half(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = x & 0x80000000u;
    x &= 0x7FFFFFFFu;
    float out;
    if (x >= 0x7F800000u) {                    // inf / nan
        out = f;
        return out;
    }
    if (x >= 0x477FF000u) {                    // >= 65520: rounds to +-inf in fp16
        uint32_t inf = sign | 0x7F800000u;
        std::memcpy(&out, &inf, 4);
        return out;
    }
    float a;
    std::memcpy(&a, &x, 4);
    if (x < 0x38800000u) {                     // subnormal in fp16: quantum 2^-24
        const float q = 5.9604644775390625e-08f;
        a = std::nearbyintf(a / q) * q;
    } else {                                   // normal: keep 10 mantissa bits, RNE
        const uint32_t lsb = (x >> 13) & 1u;
        x += 0xFFFu + lsb;
        x &= 0xFFFFE000u;
        std::memcpy(&a, &x, 4);
    }
    uint32_t r;
    std::memcpy(&r, &a, 4);
    r |= sign;
    std::memcpy(&out, &r, 4);
    return out;
}

// Round to bfloat16 (round-to-nearest-even) and back, like the engine's f32_to_bf16_kernel.
inline float round_to_bf16(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u)
        u = (u | 0x00400000u) & 0xFFFF0000u;
    else
        u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
    std::memcpy(&f, &u, 4);
    return f;
}

struct Oracle;

// DistanceComputer (A.3): query↔stored and stored↔stored; IP is negated so that
// "smaller is better" holds everywhere inside the graph code.
struct DistanceComputer {
    const Oracle* o;
    const float* q = nullptr;
    explicit DistanceComputer(const Oracle* o_) : o(o_) {}
    void set_query(const float* x) { q = x; }
    inline float pair(const float* a, const float* b) const;
    inline float operator()(storage_idx_t i) const;
    inline float symmetric_dis(storage_idx_t i, storage_idx_t j) const;
};

// VisitedTable (A.12)
struct VisitedTable {
    std::vector<uint8_t> visited;
    uint8_t visno = 1;
    explicit VisitedTable(size_t n) : visited(n, 0) {}
    void set(int no) { visited[no] = visno; }
    bool get(int no) const { return visited[no] == visno; }
    void advance() {
        visno++;
        if (visno == 250) {
            std::fill(visited.begin(), visited.end(), 0);
            visno = 1;
        }
    }
};

// Binary max-heap on parallel (val,id) arrays — the restatement of faiss/utils/Heap.h
// maxheap_push / pop / replace_top used by MinimaxHeap and the result heap.
inline bool heap_gt(float v1, idx_t i1, float v2, idx_t i2) {
    return v1 > v2 || (v1 == v2 && i1 > i2);
}
template <class ID>
inline void maxheap_sift_down(size_t k, float* val, ID* ids, size_t i, float v, ID id) {
    for (;;) {
        size_t c1 = 2 * i + 1, c2 = c1 + 1;
        if (c1 >= k) break;
        size_t c = (c2 < k && heap_gt(val[c2], ids[c2], val[c1], ids[c1])) ? c2 : c1;
        if (!heap_gt(val[c], ids[c], v, id)) break;
        val[i] = val[c];
        ids[i] = ids[c];
        i = c;
    }
    val[i] = v;
    ids[i] = id;
}
template <class ID>
inline void maxheap_push(size_t k /*size after push*/, float* val, ID* ids, float v, ID id) {
    size_t i = k - 1;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (!heap_gt(v, id, val[p], ids[p])) break;
        val[i] = val[p];
        ids[i] = ids[p];
        i = p;
    }
    val[i] = v;
    ids[i] = id;
}
template <class ID>
inline void maxheap_pop(size_t k /*size before pop*/, float* val, ID* ids) {
    float v = val[k - 1];
    ID id = ids[k - 1];
    maxheap_sift_down(k - 1, val, ids, 0, v, id);
}
template <class ID>
inline void maxheap_replace_top(size_t k, float* val, ID* ids, float v, ID id) {
    maxheap_sift_down(k, val, ids, 0, v, id);
}

// MinimaxHeap (A.6 / SURVEY §8a5): capacity-n buffer organised as a max-heap for
// eviction; pop_min is a linear scan and leaves the popped slot in place with id=-1.
struct MinimaxHeap {
    int n, k = 0, nvalid = 0;
    std::vector<storage_idx_t> ids;
    std::vector<float> dis;
    explicit MinimaxHeap(int n_) : n(n_), ids(n_), dis(n_) {}
    void push(storage_idx_t i, float v) {
        if (k == n) {
            if (v >= dis[0]) return;
            if (ids[0] != -1) --nvalid;
            maxheap_pop(k--, dis.data(), ids.data());
        }
        maxheap_push(++k, dis.data(), ids.data(), v, i);
        ++nvalid;
    }
    int size() const { return nvalid; }
    int pop_min(float* vmin_out) {
        int i = k - 1;
        while (i >= 0 && ids[i] == -1) i--;
        if (i < 0) return -1;
        int imin = i;
        float vmin = dis[i];
        for (i--; i >= 0; i--)
            if (ids[i] != -1 && dis[i] < vmin) {
                vmin = dis[i];
                imin = i;
            }
        if (vmin_out) *vmin_out = vmin;
        int ret = ids[imin];
        ids[imin] = -1;
        --nvalid;
        return ret;
    }
    int count_below(float thresh) const {
        int c = 0;
        for (int i = 0; i < k; i++)
            if (dis[i] < thresh) c++;
        return c;
    }
};

struct NodeDistCloser {  // top of a priority_queue = farthest
    float d;
    int id;
    bool operator<(const NodeDistCloser& o) const { return d < o.d; }
};
struct NodeDistFarther {  // top of a priority_queue = nearest
    float d;
    int id;
    bool operator<(const NodeDistFarther& o) const { return d > o.d; }
};

struct QueryStats {
    int ndis0 = 0, nhops0 = 0, ndis_up = 0, nhops_up = 0;
};

struct Oracle {
    int d, M, metric;
    int efConstruction = 40, efSearch = 16;  // A.1 defaults
    bool check_relative_distance = true;
    int team = 0;
    int half_storage = 0;  // 1: vectors rounded to IEEE fp16 on add, 2: to bfloat16 (the engine's opt-in storage modes)
    std::vector<double> assign_probas;
    std::vector<int> cum_nneighbor_per_level;
    std::vector<int> levels;      // level+1 per vertex
    std::vector<size_t> offsets;  // offsets[i] .. offsets[i+1] = all rows of vertex i
    std::vector<storage_idx_t> neighbors;
    storage_idx_t entry_point = -1;
    int max_level = -1;
    std::mt19937 rng{12345};
    std::vector<float> xb;
    idx_t ntotal = 0;
    std::string last_error;

    Oracle(int d_, int M_, int metric_) : d(d_), M(M_), metric(metric_) {
        set_default_probas(M, 1.0 / std::log((double)M));
        offsets.push_back(0);
    }

    // A.2 — set_default_probas
    void set_default_probas(int M_, float levelMult) {
        int nn = 0;
        cum_nneighbor_per_level.push_back(0);
        for (int level = 0;; level++) {
            float proba = std::exp(-level / (double)levelMult) * (1 - std::exp(-1 / (double)levelMult));
            if (proba < 1e-9) break;
            assign_probas.push_back(proba);
            nn += level == 0 ? M_ * 2 : M_;
            cum_nneighbor_per_level.push_back(nn);
        }
    }
    int nb_neighbors(int level) const {
        return cum_nneighbor_per_level[level + 1] - cum_nneighbor_per_level[level];
    }
    int cum_nb_neighbors(int level) const { return cum_nneighbor_per_level[level]; }
    void neighbor_range(idx_t no, int level, size_t* begin, size_t* end) const {
        size_t o = offsets[no];
        *begin = o + cum_nb_neighbors(level);
        *end = o + cum_nb_neighbors(level + 1);
    }
    // A.2 — random_level; RandomGenerator::rand_float() = mt() / float(mt.max())
    int random_level() {
        double f = rng() / float(rng.max());
        for (size_t level = 0; level < assign_probas.size(); level++) {
            if (f < assign_probas[level]) return (int)level;
            f -= assign_probas[level];
        }
        return (int)assign_probas.size() - 1;
    }
    // A.2 / A.7 — prepare_level_tab
    int prepare_level_tab(size_t n, bool preset_levels) {
        size_t n0 = offsets.size() - 1;
        if (!preset_levels) {
            for (size_t i = 0; i < n; i++) levels.push_back(random_level() + 1);
        }
        int ml = 0;
        for (size_t i = 0; i < n; i++) {
            int pt_level = levels[i + n0] - 1;
            ml = std::max(ml, pt_level);
            offsets.push_back(offsets.back() + cum_nb_neighbors(pt_level + 1));
            neighbors.resize(offsets.back(), -1);
        }
        return ml;
    }
};

inline float DistanceComputer::pair(const float* a, const float* b) const {
    const int ce = o->half_storage ? 8 : 4;
    if (o->metric == METRIC_L2)
        return o->team ? team_order<true>(a, b, o->d, o->team, ce) : l2_native(a, b, o->d);
    float s = o->team ? team_order<false>(a, b, o->d, o->team, ce) : ip_native(a, b, o->d);
    return -s;  // NegativeDistanceComputer
}
inline float DistanceComputer::operator()(storage_idx_t i) const {
    return pair(q, o->xb.data() + (size_t)i * o->d);
}
inline float DistanceComputer::symmetric_dis(storage_idx_t i, storage_idx_t j) const {
    return pair(o->xb.data() + (size_t)i * o->d, o->xb.data() + (size_t)j * o->d);
}

// A.4 — greedy_update_nearest
void greedy_update_nearest(const Oracle& h, DistanceComputer& qdis, int level,
                           storage_idx_t& nearest, float& d_nearest, QueryStats* st) {
    for (;;) {
        storage_idx_t prev = nearest;
        size_t begin, end;
        h.neighbor_range(nearest, level, &begin, &end);
        if (st) st->nhops_up++;
        for (size_t i = begin; i < end; i++) {
            storage_idx_t v = h.neighbors[i];
            if (v < 0) break;
            if (st) st->ndis_up++;
            float dis = qdis(v);
            if (dis < d_nearest) {
                nearest = v;
                d_nearest = dis;
            }
        }
        if (nearest == prev) return;
    }
}

// A.6 — search_from_candidates (bounded queue; the default branch)
// `sel` (may be null) is faiss's IDSelectorBitmap: id i is a member iff bit (i & 7) of byte i >> 3 is
// set. As in faiss the selector filters what enters the RESULT heap only; traversal is unchanged.
inline bool is_member(const uint8_t* sel, idx_t id) { return !sel || ((sel[id >> 3] >> (id & 7)) & 1); }

int search_from_candidates(const Oracle& h, DistanceComputer& qdis, int k, idx_t* I, float* D,
                           MinimaxHeap& candidates, VisitedTable& vt, int level, int efSearch,
                           bool do_dis_check, QueryStats* st, const uint8_t* sel = nullptr) {
    int nres = 0;
    for (int i = 0; i < candidates.k; i++) {
        idx_t v1 = candidates.ids[i];
        float dd = candidates.dis[i];
        if (is_member(sel, v1)) {
            if (nres < k) {
                maxheap_push<idx_t>(++nres, D, I, dd, v1);
            } else if (dd < D[0]) {
                maxheap_replace_top<idx_t>(nres, D, I, dd, v1);
            }
        }
        vt.set((int)v1);
    }
    int nstep = 0;
    while (candidates.size() > 0) {
        float d0 = 0;
        int v0 = candidates.pop_min(&d0);
        if (do_dis_check) {
            if (candidates.count_below(d0) >= efSearch) break;
        }
        size_t begin, end;
        h.neighbor_range(v0, level, &begin, &end);
        if (st) st->nhops0++;
        for (size_t j = begin; j < end; j++) {
            int v1 = h.neighbors[j];
            if (v1 < 0) break;
            if (vt.get(v1)) continue;
            vt.set(v1);
            if (st) st->ndis0++;
            float dd = qdis(v1);
            if (is_member(sel, v1)) {
                if (nres < k) {
                    maxheap_push<idx_t>(++nres, D, I, dd, (idx_t)v1);
                } else if (dd < D[0]) {
                    maxheap_replace_top<idx_t>(nres, D, I, dd, (idx_t)v1);
                }
            }
            candidates.push(v1, dd);
        }
        nstep++;
        if (!do_dis_check && nstep > efSearch) break;
    }
    return nres;
}

// A.5 — HNSW::search (upper_beam == 1, search_bounded_queue == true only)
int hnsw_search(const Oracle& h, DistanceComputer& qdis, int k, idx_t* I, float* D,
                VisitedTable& vt, int efSearch, QueryStats* st, const uint8_t* sel = nullptr) {
    if (h.entry_point == -1) return 0;
    storage_idx_t nearest = h.entry_point;
    float d_nearest = qdis(nearest);
    for (int level = h.max_level; level >= 1; level--)
        greedy_update_nearest(h, qdis, level, nearest, d_nearest, st);
    int ef = std::max(efSearch, k);
    MinimaxHeap candidates(ef);
    candidates.push(nearest, d_nearest);
    int nres = search_from_candidates(h, qdis, k, I, D, candidates, vt, 0, efSearch,
                                      h.check_relative_distance, st, sel);
    vt.advance();
    return nres;
}

// A.10 — shrink_neighbor_list (vector-output overload)
void shrink_neighbor_list(DistanceComputer& qdis, std::priority_queue<NodeDistFarther>& input,
                          std::vector<NodeDistFarther>& output, int max_size) {
    while (!input.empty()) {
        NodeDistFarther v1 = input.top();
        input.pop();
        float dist_v1_q = v1.d;
        bool good = true;
        for (const NodeDistFarther& v2 : output) {
            float dist_v1_v2 = qdis.symmetric_dis(v2.id, v1.id);
            if (dist_v1_v2 < dist_v1_q) {
                good = false;
                break;
            }
        }
        if (good) {
            output.push_back(v1);
            if ((int)output.size() >= max_size) return;
        }
    }
}
// A.10 — shrink_neighbor_list (in-place overload on the "closer" queue)
void shrink_neighbor_list(DistanceComputer& qdis, std::priority_queue<NodeDistCloser>& rs1,
                          int max_size) {
    if ((int)rs1.size() < max_size) return;
    std::priority_queue<NodeDistFarther> rs;
    std::vector<NodeDistFarther> ret;
    while (!rs1.empty()) {
        rs.push({rs1.top().d, rs1.top().id});
        rs1.pop();
    }
    shrink_neighbor_list(qdis, rs, ret, max_size);
    for (const NodeDistFarther& c : ret) rs1.push({c.d, c.id});
}

// A.11 — add_link
void add_link(Oracle& h, DistanceComputer& qdis, storage_idx_t src, storage_idx_t dest, int level) {
    size_t begin, end;
    h.neighbor_range(src, level, &begin, &end);
    if (h.neighbors[end - 1] == -1) {
        size_t i = end;
        while (i > begin) {
            if (h.neighbors[i - 1] != -1) break;
            i--;
        }
        h.neighbors[i] = dest;
        return;
    }
    std::priority_queue<NodeDistCloser> rs;
    rs.push({qdis.symmetric_dis(src, dest), dest});
    for (size_t i = begin; i < end; i++) {
        storage_idx_t neigh = h.neighbors[i];
        rs.push({qdis.symmetric_dis(src, neigh), neigh});
    }
    shrink_neighbor_list(qdis, rs, (int)(end - begin));
    size_t i = begin;
    while (!rs.empty()) {
        h.neighbors[i++] = rs.top().id;
        rs.pop();
    }
    while (i < end) h.neighbors[i++] = -1;
}

// A.9 — search_neighbors_to_add
void search_neighbors_to_add(Oracle& h, DistanceComputer& qdis,
                             std::priority_queue<NodeDistCloser>& results, int entry_point,
                             float d_entry_point, int level, VisitedTable& vt) {
    std::priority_queue<NodeDistFarther> candidates;
    candidates.push({d_entry_point, entry_point});
    results.push({d_entry_point, entry_point});
    vt.set(entry_point);
    while (!candidates.empty()) {
        const NodeDistFarther& cur = candidates.top();
        if (cur.d > results.top().d) break;
        int cur_node = cur.id;
        candidates.pop();
        size_t begin, end;
        h.neighbor_range(cur_node, level, &begin, &end);
        for (size_t i = begin; i < end; i++) {
            storage_idx_t node = h.neighbors[i];
            if (node < 0) break;
            if (vt.get(node)) continue;
            vt.set(node);
            float dis = qdis(node);
            if ((int)results.size() < h.efConstruction || results.top().d > dis) {
                results.push({dis, node});
                candidates.push({dis, node});
                if ((int)results.size() > h.efConstruction) results.pop();
            }
        }
    }
    vt.advance();
}

// A.9 — add_links_starting_from
void add_links_starting_from(Oracle& h, DistanceComputer& ptdis, storage_idx_t pt_id,
                             storage_idx_t nearest, float d_nearest, int level, omp_lock_t* locks,
                             VisitedTable& vt) {
    std::priority_queue<NodeDistCloser> link_targets;
    search_neighbors_to_add(h, ptdis, link_targets, nearest, d_nearest, level, vt);
    int Mlev = h.nb_neighbors(level);
    shrink_neighbor_list(ptdis, link_targets, Mlev);
    std::vector<storage_idx_t> nbrs;
    nbrs.reserve(link_targets.size());
    while (!link_targets.empty()) {
        storage_idx_t other = link_targets.top().id;
        add_link(h, ptdis, pt_id, other, level);
        nbrs.push_back(other);
        link_targets.pop();
    }
    omp_unset_lock(&locks[pt_id]);
    for (storage_idx_t other : nbrs) {
        omp_set_lock(&locks[other]);
        add_link(h, ptdis, other, pt_id, level);
        omp_unset_lock(&locks[other]);
    }
    omp_set_lock(&locks[pt_id]);
}

// A.8 — add_with_locks
void add_with_locks(Oracle& h, DistanceComputer& ptdis, int pt_level, int pt_id,
                    std::vector<omp_lock_t>& locks, VisitedTable& vt) {
    storage_idx_t nearest;
#pragma omp critical
    {
        nearest = h.entry_point;
        if (nearest == -1) {
            h.max_level = pt_level;
            h.entry_point = pt_id;
        }
    }
    if (nearest < 0) return;
    omp_set_lock(&locks[pt_id]);
    int level = h.max_level;
    float d_nearest = ptdis(nearest);
    for (; level > pt_level; level--) greedy_update_nearest(h, ptdis, level, nearest, d_nearest, nullptr);
    for (; level >= 0; level--)
        add_links_starting_from(h, ptdis, pt_id, nearest, d_nearest, level, locks.data(), vt);
    omp_unset_lock(&locks[pt_id]);
    if (pt_level > h.max_level) {
        h.max_level = pt_level;
        h.entry_point = pt_id;
    }
}

// A.7 — hnsw_add_vertices. `order_out`, when non-null, receives the insertion order
// (length n) so the GPU build can be driven through the identical sequence in tests.
void hnsw_add_vertices(Oracle& h, size_t n0, size_t n, const float* x, bool preset_levels,
                       int nthreads, int32_t* order_out) {
    if (n == 0) return;
    size_t ntotal = n0 + n;
    h.prepare_level_tab(n, preset_levels);
    std::vector<omp_lock_t> locks(ntotal);
    for (size_t i = 0; i < ntotal; i++) omp_init_lock(&locks[i]);

    std::vector<int> hist;
    std::vector<int> order(n);
    {
        for (size_t i = 0; i < n; i++) {
            int pt_level = h.levels[i + n0] - 1;
            while (pt_level >= (int)hist.size()) hist.push_back(0);
            hist[pt_level]++;
        }
        std::vector<int> offs(hist.size() + 1, 0);
        for (size_t i = 0; i + 1 < hist.size(); i++) offs[i + 1] = offs[i] + hist[i];
        for (size_t i = 0; i < n; i++) {
            storage_idx_t pt_id = (storage_idx_t)(i + n0);
            int pt_level = h.levels[pt_id] - 1;
            order[offs[pt_level]++] = pt_id;
        }
    }
    {
        std::mt19937 rng2(789);  // RandomGenerator rng2(789); rand_int(m) = mt() % m
        int i1 = (int)n;
        for (int pt_level = (int)hist.size() - 1; pt_level >= 0; pt_level--) {
            int i0 = i1 - hist[pt_level];
            for (int j = i0; j < i1; j++) std::swap(order[j], order[j + rng2() % (i1 - j)]);
            const bool par = nthreads > 1 && i1 > i0 + 100;
#pragma omp parallel num_threads(par ? nthreads : 1)
            {
                VisitedTable vt(ntotal);
                DistanceComputer dis(&h);
#pragma omp for schedule(static)
                for (int i = i0; i < i1; i++) {
                    storage_idx_t pt_id = order[i];
                    dis.set_query(x + (size_t)(pt_id - n0) * h.d);
                    add_with_locks(h, dis, pt_level, pt_id, locks, vt);
                }
            }
            i1 = i0;
        }
    }
    if (order_out) {
        // insertion sequence: highest bucket first, each bucket in its shuffled order
        size_t w = 0;
        int i1 = (int)n;
        for (int pt_level = (int)hist.size() - 1; pt_level >= 0; pt_level--) {
            int i0 = i1 - hist[pt_level];
            for (int i = i0; i < i1; i++) order_out[w++] = order[i];
            i1 = i0;
        }
    }
    for (size_t i = 0; i < ntotal; i++) omp_destroy_lock(&locks[i]);
}

}  // namespace

// ------------------------------------------------------------------ C interface
extern "C" {

void* orc_create(int d, int M, int metric) {
    if (d <= 0 || M < 2 || (metric != METRIC_L2 && metric != METRIC_INNER_PRODUCT)) return nullptr;
    return new Oracle(d, M, metric);
}
void orc_free(void* p) { delete static_cast<Oracle*>(p); }
void orc_set_ef_construction(void* p, int v) { static_cast<Oracle*>(p)->efConstruction = v; }
void orc_set_ef_search(void* p, int v) { static_cast<Oracle*>(p)->efSearch = v; }
void orc_set_check_relative_distance(void* p, int v) {
    static_cast<Oracle*>(p)->check_relative_distance = v != 0;
}
// 0 = native SIMD order; T in {1,2,4,8,16,32} = CUDA team order (d % 4 == 0 required)
int orc_set_team(void* p, int T) {
    Oracle* o = static_cast<Oracle*>(p);
    if (T != 0 && (o->d % 4 != 0 || T < 1 || T > 32 || (T & (T - 1)))) return 1;
    o->team = T;
    return 0;
}
// fp16 storage emulation: must be set before the first add/import; needs d % 8 == 0 in team mode
int orc_set_half_storage(void* p, int on) {
    Oracle* o = static_cast<Oracle*>(p);
    if (o->ntotal != 0 || (on && o->d % 8 != 0)) return 1;
    o->half_storage = on;  // 0 off, 1 fp16, 2 bf16
    return 0;
}
float orc_round_to_half(float f) { return round_to_half(f); }
int64_t orc_ntotal(void* p) { return static_cast<Oracle*>(p)->ntotal; }
int orc_entry_point(void* p) { return static_cast<Oracle*>(p)->entry_point; }
int orc_max_level(void* p) { return static_cast<Oracle*>(p)->max_level; }
int orc_max_threads() { return omp_get_max_threads(); }
int64_t orc_neighbors_size(void* p) { return (int64_t) static_cast<Oracle*>(p)->neighbors.size(); }
int orc_n_levels_table(void* p) { return (int) static_cast<Oracle*>(p)->assign_probas.size(); }
void orc_get_assign_probas(void* p, double* out) {
    Oracle* o = static_cast<Oracle*>(p);
    std::copy(o->assign_probas.begin(), o->assign_probas.end(), out);
}
void orc_get_cum_nneighbor(void* p, int* out) {
    Oracle* o = static_cast<Oracle*>(p);
    std::copy(o->cum_nneighbor_per_level.begin(), o->cum_nneighbor_per_level.end(), out);
}

// IndexHNSW::add (A.7): storage->add then hnsw_add_vertices. nthreads<=1 = sequential,
// deterministic. order_out may be null.
int orc_add(void* p, int64_t n, const float* x, int nthreads, int32_t* order_out) {
    Oracle* o = static_cast<Oracle*>(p);
    if (n < 0) return 1;
    size_t n0 = (size_t)o->ntotal;
    o->xb.insert(o->xb.end(), x, x + (size_t)n * o->d);
    if (o->half_storage)
        for (size_t i = n0 * o->d; i < o->xb.size(); i++)
            o->xb[i] = o->half_storage == 2 ? round_to_bf16(o->xb[i]) : round_to_half(o->xb[i]);
    o->ntotal += n;
    // The build reads vectors from o->xb (which may have been reallocated), so pass
    // the stored copy, not the caller's pointer.
    hnsw_add_vertices(*o, n0, (size_t)n, o->xb.data() + n0 * o->d,
                      o->levels.size() == (size_t)o->ntotal, nthreads, order_out);
    return 0;
}

// add() with caller-supplied levels (level + 1 per point), as faiss allows by filling hnsw.levels before
// add: prepare_level_tab then skips the random draw (A.7). Used by the hand-worked golden test.
int orc_add_preset(void* p, int64_t n, const float* x, const int* levels, int nthreads) {
    Oracle* o = static_cast<Oracle*>(p);
    if (n < 0 || (size_t)o->ntotal != o->levels.size()) return 1;
    for (int64_t i = 0; i < n; i++) {
        if (levels[i] < 1 || levels[i] > (int)o->assign_probas.size()) return 2;
    }
    o->levels.insert(o->levels.end(), levels, levels + n);
    return orc_add(p, n, x, nthreads, nullptr);
}

// Draw levels for the next n points exactly as add() would (advances the index RNG) —
// lets a test feed the same levels to the GPU engine.
void orc_peek_levels(void* p, int64_t n, int* levels_out) {
    Oracle* o = static_cast<Oracle*>(p);
    std::mt19937 saved = o->rng;
    for (int64_t i = 0; i < n; i++) levels_out[i] = o->random_level() + 1;
    o->rng = saved;
}

// IndexHNSW::search: per query heapify → HNSW::search → reorder ascending;
// IP distances are negated back at the end. stats (may be null) is int32[nq][4] =
// {ndis level0, nhops level0, ndis upper, nhops upper}.
int orc_search_sel(void* p, int64_t nq, const float* xq, int64_t k, float* D, int64_t* I,
                   int ef_search, int nthreads, int32_t* stats, const uint8_t* sel);

int orc_search(void* p, int64_t nq, const float* xq, int64_t k, float* D, int64_t* I,
               int ef_search /*<=0: index default*/, int nthreads, int32_t* stats) {
    return orc_search_sel(p, nq, xq, k, D, I, ef_search, nthreads, stats, nullptr);
}

int orc_search_sel(void* p, int64_t nq, const float* xq, int64_t k, float* D, int64_t* I,
                   int ef_search, int nthreads, int32_t* stats, const uint8_t* sel) {
    Oracle* o = static_cast<Oracle*>(p);
    if (k <= 0) return 1;
    const int ef = ef_search > 0 ? ef_search : o->efSearch;
#pragma omp parallel num_threads(nthreads > 1 ? nthreads : 1)
    {
        VisitedTable vt((size_t)o->ntotal);
        DistanceComputer dis(o);
#pragma omp for schedule(guided)
        for (int64_t i = 0; i < nq; i++) {
            idx_t* idxi = I + i * k;
            float* simi = D + i * k;
            dis.set_query(xq + i * o->d);
            for (int64_t j = 0; j < k; j++) {  // maxheap_heapify with no input
                simi[j] = FLT_MAX;
                idxi[j] = -1;
            }
            QueryStats st;
            int nres = hnsw_search(*o, dis, (int)k, idxi, simi, vt, ef, &st, sel);
            // maxheap_reorder: valid entries ascending, then (FLT_MAX,-1) padding
            for (int m = nres; m > 1; m--) {
                float v = simi[0];
                idx_t id = idxi[0];
                maxheap_pop<idx_t>(m, simi, idxi);
                simi[m - 1] = v;
                idxi[m - 1] = id;
            }
            if (stats) {
                stats[4 * i + 0] = st.ndis0;
                stats[4 * i + 1] = st.nhops0;
                stats[4 * i + 2] = st.ndis_up;
                stats[4 * i + 3] = st.nhops_up;
            }
        }
    }
    if (o->metric == METRIC_INNER_PRODUCT)
        for (int64_t i = 0; i < nq * k; i++) D[i] = -D[i];
    return 0;
}

// Graph export / import in faiss's own layout.
void orc_export_graph(void* p, int* levels, uint64_t* offsets, int32_t* neighbors) {
    Oracle* o = static_cast<Oracle*>(p);
    std::copy(o->levels.begin(), o->levels.end(), levels);
    for (size_t i = 0; i < o->offsets.size(); i++) offsets[i] = o->offsets[i];
    std::copy(o->neighbors.begin(), o->neighbors.end(), neighbors);
}
int orc_import(void* p, int64_t n, const float* x, const int* levels, const int32_t* neighbors,
               int64_t nneigh, int entry_point, int max_level) {
    Oracle* o = static_cast<Oracle*>(p);
    o->xb.assign(x, x + (size_t)n * o->d);
    if (o->half_storage)
        for (float& v : o->xb) v = o->half_storage == 2 ? round_to_bf16(v) : round_to_half(v);
    o->ntotal = n;
    o->levels.assign(levels, levels + n);
    o->offsets.assign(1, 0);
    for (int64_t i = 0; i < n; i++) {
        if (levels[i] < 1 || levels[i] >= (int)o->cum_nneighbor_per_level.size()) return 1;
        o->offsets.push_back(o->offsets.back() + o->cum_nb_neighbors(levels[i]));
    }
    if ((int64_t)o->offsets.back() != nneigh) return 2;
    o->neighbors.assign(neighbors, neighbors + nneigh);
    o->entry_point = entry_point;
    o->max_level = max_level;
    return 0;
}

// Stand-alone pieces exposed for unit / property tests.
float orc_distance(void* p, const float* a, const float* b) {
    Oracle* o = static_cast<Oracle*>(p);
    DistanceComputer dc(o);
    return dc.pair(a, b);
}
// mt19937 anchors: first outputs of the two generators the build uses.
uint32_t orc_mt19937_first(uint32_t seed) {
    std::mt19937 g(seed);
    return (uint32_t)g();
}
// shrink_neighbor_list on an explicit candidate set: ids[n] with distances dq[n] to a
// base vector (not necessarily stored); returns kept ids, nearest first.
int orc_shrink(void* p, int n, const int32_t* ids, const float* dq, int max_size, int32_t* out) {
    Oracle* o = static_cast<Oracle*>(p);
    DistanceComputer dc(o);
    std::priority_queue<NodeDistFarther> in;
    for (int i = 0; i < n; i++) in.push({dq[i], ids[i]});
    std::vector<NodeDistFarther> res;
    if (n < max_size) {
        while (!in.empty()) {
            res.push_back(in.top());
            in.pop();
        }
    } else {
        shrink_neighbor_list(dc, in, res, max_size);
    }
    for (size_t i = 0; i < res.size(); i++) out[i] = res[i].id;
    return (int)res.size();
}

}  // extern "C"
