// cuclark_b200 — internal handle and cross-TU prototypes (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/cuclark_b200.h"
#include "common.cuh"

namespace cuclark {

constexpr int COUNTER_DENSE = 0;       // reads routed to the dense fallback
constexpr int COUNTER_TRUNC = 1;       // rows with more than row_pairs targets
constexpr int COUNTER_LOOKUPS = 2;     // uint64 at [2],[3]
constexpr int N_COUNTERS = 8;

// Per-call scratch: counters and the list of reads for the dense fallback.
struct Scratch {
    uint32_t* d_counters = nullptr;    // N_COUNTERS
    uint32_t* d_dense_list = nullptr;
    uint32_t dense_cap = 0;
    uint32_t* d_dense_hist = nullptr;  // dense_blocks * n_targets counters of the dense fallback, all zero between calls.
                                       // Every batch / pipeline slot has its own, so calls on different streams are independent
};

struct Batch {
    Scratch scratch;
    uint32_t* h_counters = nullptr;    // pinned copy of the counters
    uint32_t* h_ptr = nullptr;         // pinned
    uint16_t* h_cont = nullptr;
    uint16_t* h_final = nullptr;
    uint16_t* h_rows = nullptr;
    uint32_t* d_ptr = nullptr;
    uint16_t* d_cont = nullptr;
    uint16_t* d_final = nullptr;
    uint16_t* d_rows = nullptr;
    size_t n_reads = 0, n_cont = 0;
    bool ready = false, queried = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
};

struct TextPipe;   // stream.cu
struct RouteCtx;   // route.cu
constexpr int ROUTE_MAX_RANKS = CUCLARK_ROUTE_MAX_RANKS;

}  // namespace cuclark

struct cuclark_db {
    cuclark_config cfg;
    int row_pairs;
    int key_bytes;
    // device table
    uint4* d_table = nullptr;
    uint4* d_ovf = nullptr;            // overflow table
    cuclark::TableView view{};
    uint64_t n_entries = 0, n_spilled = 0, n_spill_buckets = 0;
    int src_sfactor = 1;               // -s the table was built with, and the sizes of its source files
    uint64_t src_bytes[3] = {0, 0, 0}; //   (.sz/.ky/.lb; 0 when built from arrays or synthetic): table cache header
    uint64_t src_mtime_ns[3] = {0, 0, 0};
    // classify scratch (one-shot calls; batches carry their own)
    cuclark::Scratch scratch;
    uint32_t* d_dense_hist = nullptr;  // dense_blocks * n_targets, for callers whose Scratch has none: those dense kernels
    cudaEvent_t dense_chain = nullptr; //   are serialised across streams through this event
    std::mutex dense_mu;               //   (wait + launch + record must not interleave between host threads)
    int dense_blocks = 0;
    int classify_blocks_per_sm[6] = {0, 0, 0, 0, 0, 0};
    int sm_count = 0;
    cudaStream_t stream = nullptr;     // library-owned default stream
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // last-call stats
    uint64_t last_lookups = 0, last_dense = 0, last_trunc = 0;
    double last_ms = 0;
    // batches
    std::vector<cuclark::Batch> batches;
    size_t batch_max_reads = 0, batch_max_cont = 0;
    bool batch_rows = false;
    // text pipeline (slots are allocated on first use and kept)
    cuclark::TextPipe* text_pipe = nullptr;
    // table-partitioned mode by k-mer routing (route.cu)
    cuclark::RouteCtx* route = nullptr;
};

namespace cuclark {

// kernels launched by the hot path since the library was loaded (cuclark_kernel_launches): classify, merge,
// text pipeline and k-mer routing launch sites count themselves; table builders and generators do not
extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launches(int n) { g_kernel_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// table.cu
int table_build_from_arrays(cuclark_db* db, const uint8_t* sz, const void* ky, const uint16_t* lb,
                            uint64_t n_entries_file, int sfactor, const char* base_path);
int table_build_synthetic(cuclark_db* db, uint32_t seed, uint32_t n_targets, uint64_t genome_len, int light_gap);
void table_free(cuclark_db* db);
int table_clone(cuclark_db* src, cuclark_db* dst);
void table_plan(const cuclark_config& cfg, uint64_t n_entries, cuclark_table_plan* out);
int table_save(cuclark_db* db, const char* path);
int table_load(cuclark_db* db, const char* path, const char* src_base, int sfactor);

// classify.cu
int classify_launch(cuclark_db* db, const Scratch& sc, const uint32_t* d_ptr, const uint16_t* d_cont, size_t n_reads,
                    uint16_t* d_final, uint16_t* d_rows, cudaStream_t st);
int merge_rows_launch(cuclark_db* db, const uint16_t* d_parts, int n_parts, size_t n_reads, uint16_t* d_rows_out,
                      uint16_t* d_final, cudaStream_t st);
int synth_reads_launch(uint32_t seed, uint32_t genome_seed, uint32_t n_targets, uint64_t genome_len,
                       uint64_t first_read, size_t n_reads, int read_len, int pct_random, int sub_per_10k,
                       uint32_t* d_ptr, uint16_t* d_cont, cudaStream_t st);
int synth_fastq_launch(uint32_t seed, uint32_t genome_seed, uint32_t n_targets, uint64_t genome_len, uint64_t first_read,
                       size_t n_reads, int read_len, int pct_random, int sub_per_10k, uint8_t* d_text, cudaStream_t st, int mate = 0);
// stream.cu
void text_pipe_free(cuclark_db* db);

// route.cu
int route_alloc(cuclark_db* db, int n_ranks, size_t max_containers);
void route_free(cuclark_db* db);
int route_export(cuclark_db* db, void* handle64, uint64_t* bytes);
int route_import(cuclark_db* db, int peer_rank, const void* handle64);
int route_connect(cuclark_db* const* dbs, int n);
int route_scatter(cuclark_db* db, const uint32_t* d_ptr, const uint16_t* d_cont, size_t n_reads, size_t n_cont, cudaStream_t st);
int route_probe(cuclark_db* db, cudaStream_t st);
int route_gather(cuclark_db* db, const Scratch& sc, const uint32_t* d_ptr, const uint16_t* d_cont, size_t n_reads, size_t n_cont,
                 uint16_t* d_final, uint16_t* d_rows, cudaStream_t st);
int route_stats(cuclark_db* db, cuclark_route_stats* out);

int gather_bench_launch(cuclark_db* db, uint64_t n_probes, int bytes_per_probe, int ilp, int iters, double* ms_out);

}  // namespace cuclark
