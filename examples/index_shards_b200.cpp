// index_shards_b200.cpp — a faiss::IndexShards(successive_ids = true) replacement over the GPUs of one
// box, written against the C-ABI only (include/b200_hnsw.h): no Python, no torch, no NCCL.
//
// One process, one host thread per GPU ("rank"). Every rank builds an IndexHNSWFlat over its contiguous
// slice of the database, the ranks exchange their bootstrap blobs through plain host memory, and a search
// is the collective bh_shards_search_device: each rank's traversal kernel stores its per-query top-k
// straight into every rank's gather buffer over NVLink, raises a flag, and merges when all flags are up.
// The program checks the merged result against an exact host-side merge of the per-shard results.
//
//   g++ -O2 -std=c++17 -I include examples/index_shards_b200.cpp -L hnsw_b200 -lb200hnsw \
//       -L/usr/local/cuda/lib64 -lcudart -lpthread -o index_shards_b200
//   LD_LIBRARY_PATH=hnsw_b200 ./index_shards_b200 [n_per_shard] [d]
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "b200_hnsw.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        if ((call) != 0) {                                                           \
            std::fprintf(stderr, "%s failed: %s\n", #call, bh_last_error());         \
            std::exit(1);                                                            \
        }                                                                            \
    } while (0)

struct Barrier {  // all rank threads meet here (C++17: no std::barrier)
    std::atomic<int> count{0}, gen{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        const int g = gen.load();
        if (count.fetch_add(1) + 1 == n) {
            count = 0;
            gen++;
        } else {
            while (gen.load() == g) std::this_thread::yield();
        }
    }
};

int main(int argc, char** argv) {
    const int64_t n_shard = argc > 1 ? std::atoll(argv[1]) : 20000;
    const int d = argc > 2 ? std::atoi(argv[2]) : 64;
    const int M = 16, k = 10, nq = 1000, efSearch = 64;
    int ngpu = 0;
    if (cudaGetDeviceCount(&ngpu) != cudaSuccess || ngpu < 1) {
        std::fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    const int R = std::min(ngpu, 8);
    std::mt19937 rng(7);
    std::normal_distribution<float> nd;
    std::vector<float> xb((size_t)R * n_shard * d), xq((size_t)nq * d);
    for (auto& v : xb) v = nd(rng);
    for (auto& v : xq) v = nd(rng);

    std::vector<unsigned char> blobs((size_t)R * BH_SHARDS_BLOB_BYTES);
    std::vector<std::vector<float>> D(R, std::vector<float>((size_t)nq * k)), Dl = D;
    std::vector<std::vector<int64_t>> I(R, std::vector<int64_t>((size_t)nq * k)), Il = I;
    Barrier bar(R);
    std::vector<std::thread> th;
    for (int r = 0; r < R; r++)
        th.emplace_back([&, r] {
            cudaSetDevice(r);
            bh_index* idx = nullptr;
            CHECK(bh_index_create(&idx, d, M, BH_METRIC_L2, r));
            CHECK(bh_index_set_ef_construction(idx, 100));
            CHECK(bh_index_add(idx, n_shard, xb.data() + (size_t)r * n_shard * d));   // this rank's slice
            bh_shards* sh = nullptr;
            CHECK(bh_shards_create(&sh, idx, r, R, nq, k));
            CHECK(bh_shards_export(sh, blobs.data() + (size_t)r * BH_SHARDS_BLOB_BYTES));
            bar.wait();                                   // every blob is in place
            CHECK(bh_shards_connect(sh, blobs.data()));
            bar.wait();
            float *xq_d, *D_d;
            int64_t* I_d;
            cudaMalloc(&xq_d, xq.size() * 4);
            cudaMalloc(&D_d, (size_t)nq * k * 4);
            cudaMalloc(&I_d, (size_t)nq * k * 8);
            cudaMemcpy(xq_d, xq.data(), xq.size() * 4, cudaMemcpyHostToDevice);   // the "broadcast"
            bh_search_params p{};
            p.efSearch = efSearch;
            for (int rep = 0; rep < 3; rep++)             // collective: every rank calls it
                CHECK(bh_shards_search_device(sh, nq, xq_d, k, D_d, I_d, &p));
            CHECK(bh_index_synchronize(idx));
            if (bh_shards_status(sh) != 0) std::fprintf(stderr, "rank %d: a peer timed out\n", r);
            cudaMemcpy(D[r].data(), D_d, (size_t)nq * k * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(I[r].data(), I_d, (size_t)nq * k * 8, cudaMemcpyDeviceToHost);
            // this rank's own lists, for the exactness check below
            CHECK(bh_index_search(idx, nq, xq.data(), k, Dl[r].data(), Il[r].data(), &p));
            bar.wait();                                   // nobody frees a buffer a peer may still write
            cudaFree(xq_d);
            cudaFree(D_d);
            cudaFree(I_d);
            bh_shards_free(sh);
            bh_index_free(idx);
        });
    for (auto& t : th) t.join();

    // exact merge of the per-shard lists on the host: by distance, ties by shard then position
    int64_t bad = 0;
    for (int q = 0; q < nq; q++) {
        std::vector<std::tuple<float, int, int, int64_t>> all;
        for (int r = 0; r < R; r++)
            for (int i = 0; i < k; i++) {
                const int64_t id = Il[r][(size_t)q * k + i];
                all.emplace_back(Dl[r][(size_t)q * k + i], r, i, id < 0 ? -1 : id + r * n_shard);
            }
        std::sort(all.begin(), all.end());
        for (int r = 0; r < R; r++)
            for (int i = 0; i < k; i++)
                if (I[r][(size_t)q * k + i] != std::get<3>(all[i]) || D[r][(size_t)q * k + i] != std::get<0>(all[i])) bad++;
    }
    std::printf("%d shards x %lld x %d: merged top-%d of %d queries on every rank %s the exact host merge\n", R,
                (long long)n_shard, d, k, nq, bad ? "DIFFERS FROM" : "==");
    return bad ? 1 : 0;
}
