// cuclark_b200 — C ABI (include/cuclark_b200.h): handle, database loading,
// batch plumbing (pinned staging, one stream + event per batch), host calls.
//
// Mirrors the life cycle of the reference's CuClarkDB<HKMERr>
// (src/CuClarkDB.cu:84-1033): ctor -> read -> malloc -> readyBatch ->
// queryBatch -> waitForBatch -> freeBatchMemory, with these differences:
//  * one process drives ONE device (multi-GPU = one process per GPU);
//  * every batch has its own stream, so H2D of batch i+1 overlaps the kernels
//    of batch i (the reference issues everything on the default stream);
//  * final results AND rows come back in one D2H per batch.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "internal.h"

namespace cuclark {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_kernel_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

static int key_bytes_for(int k, uint64_t htsize) {
    // src/main.cc:278-316
    const size_t t_b = (size_t)(log((double)htsize) / log(4.0));
    if ((size_t)k <= t_b + 8) return 2;
    if ((size_t)k <= t_b + 16) return 4;
    return 8;
}

static int use_device(cuclark_db* db) {
    CK(cudaSetDevice(db->cfg.device));
    return CUCLARK_OK;
}

static void free_scratch(Scratch& sc) {
    cudaFree(sc.d_counters); cudaFree(sc.d_dense_list); cudaFree(sc.d_dense_hist);
    sc = Scratch{};
}

// hist_words > 0: the scratch gets its own dense-fallback histogram (dense_blocks * n_targets counters)
static int alloc_scratch(Scratch& sc, uint32_t dense_cap, size_t hist_words = 0) {
    sc.dense_cap = dense_cap;
    CK(cudaMalloc(&sc.d_counters, N_COUNTERS * sizeof(uint32_t)));
    CK(cudaMalloc(&sc.d_dense_list, (size_t)dense_cap * 4));
    CK(cudaMemset(sc.d_counters, 0, N_COUNTERS * sizeof(uint32_t)));
    if (hist_words) {
        CK(cudaMalloc(&sc.d_dense_hist, hist_words * 4));
        CK(cudaMemset(sc.d_dense_hist, 0, hist_words * 4));
    }
    return CUCLARK_OK;
}

static void free_batches(cuclark_db* db) {
    for (auto& b : db->batches) {
        if (b.stream) cudaStreamSynchronize(b.stream);
        free_scratch(b.scratch);
        cudaFreeHost(b.h_counters);
        cudaFreeHost(b.h_ptr); cudaFreeHost(b.h_cont); cudaFreeHost(b.h_final); cudaFreeHost(b.h_rows);
        cudaFree(b.d_ptr); cudaFree(b.d_cont); cudaFree(b.d_final); cudaFree(b.d_rows);
        if (b.done) cudaEventDestroy(b.done);
        if (b.stream) cudaStreamDestroy(b.stream);
    }
    db->batches.clear();
}

static int take_counters(cuclark_db* db, const uint32_t* c, uint32_t dense_cap) {
    db->last_dense = c[COUNTER_DENSE];
    db->last_trunc = c[COUNTER_TRUNC];
    db->last_lookups = (uint64_t)c[COUNTER_LOOKUPS] | ((uint64_t)c[COUNTER_LOOKUPS + 1] << 32);
    if (c[COUNTER_DENSE] > dense_cap) {
        set_error("%u reads needed the dense fallback, capacity is %u", c[COUNTER_DENSE], dense_cap);
        return CUCLARK_ERR_NOMEM;
    }
    return CUCLARK_OK;
}

static int fetch_counters(cuclark_db* db, cudaStream_t st) {
    uint32_t c[N_COUNTERS];
    CK(cudaMemcpyAsync(c, db->scratch.d_counters, sizeof c, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return take_counters(db, c, db->scratch.dense_cap);
}

}  // namespace cuclark

using namespace cuclark;

extern "C" {

const char* cuclark_last_error(void) { return g_err; }
int cuclark_version(void) { return 200; }
uint64_t cuclark_kernel_launches(void) { return g_kernel_launches.load(std::memory_order_relaxed); }

int cuclark_create(const cuclark_config* cfg, cuclark_db** out) {
    if (!cfg || !out) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    *out = nullptr;
    if (cfg->k < 2 || cfg->k > 32) { set_error("The k-mer length should be in [2,32]."); return CUCLARK_ERR_ARG; }
    if (cfg->htsize < 2) { set_error("htsize must be the HTSIZE of the database files"); return CUCLARK_ERR_ARG; }
    if (cfg->n_targets < 1 || cfg->n_targets > 65535) { set_error("n_targets must be in [1,65535]"); return CUCLARK_ERR_ARG; }
    if (cfg->shard_count > 1 && (cfg->shard_index < 0 || cfg->shard_index >= cfg->shard_count)) { set_error("bad shard index"); return CUCLARK_ERR_ARG; }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (there is no CPU fallback)");
        return CUCLARK_ERR_NO_DEVICE;
    }
    if (cfg->device < 0 || cfg->device >= n_dev) { set_error("device %d not present (%d devices)", cfg->device, n_dev); return CUCLARK_ERR_NO_DEVICE; }
    cuclark_db* db = new cuclark_db();
    db->cfg = *cfg;
    if (db->cfg.shard_count < 1) { db->cfg.shard_count = 1; db->cfg.shard_index = 0; }
    if (db->cfg.layout == 0) {                        // CUCLARK_LAYOUT=1|2|3 forces a table layout for handles that left it open
        const char* e = getenv("CUCLARK_LAYOUT");     // (the command line has no flag for it)
        if (e && e[0] >= '1' && e[0] <= '3' && !e[1]) db->cfg.layout = e[0] - '0';
    }
    db->key_bytes = cfg->key_bytes ? cfg->key_bytes : key_bytes_for(cfg->k, cfg->htsize);
    db->row_pairs = cfg->row_pairs > 0 ? cfg->row_pairs
                                       : (cfg->htsize == CUCLARK_HTSIZE_LIGHT ? CUCLARK_MAXHITS_LIGHT : CUCLARK_MAXHITS_FULL);
    if (db->key_bytes != 2 && db->key_bytes != 4 && db->key_bytes != 8) { delete db; set_error("key_bytes must be 2, 4 or 8"); return CUCLARK_ERR_ARG; }
    if (db->row_pairs > 63) { delete db; set_error("row_pairs must be <= 63"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) { delete db; return rc; }
    // A probe needs ONE 32-byte sector; by default the L2 fetches the whole 128-byte line from
    // HBM on a sector miss, which quadruples DRAM traffic for random probes (ncu:
    // dram__bytes_read = 132 B per lookup). Ask for sector-granular fetches.
    {
        size_t gran = 32;
        if (const char* e = getenv("CUCLARK_L2_FETCH")) gran = (size_t)atoi(e);
        if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaGetLastError();
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { delete db; set_error("cudaGetDeviceProperties failed"); return CUCLARK_ERR_CUDA; }
    db->sm_count = prop.multiProcessorCount;
    db->dense_blocks = db->sm_count;
    auto fail = [&](const char* what) { set_error("%s: %s", what, cudaGetErrorString(cudaGetLastError())); cuclark_destroy(db); return CUCLARK_ERR_CUDA; };
    if (cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    if (cudaEventCreate(&db->ev0) != cudaSuccess || cudaEventCreate(&db->ev1) != cudaSuccess) return fail("event");
    if (cudaEventCreateWithFlags(&db->dense_chain, cudaEventDisableTiming) != cudaSuccess) return fail("event");
    if (alloc_scratch(db->scratch, 1u << 20) != CUCLARK_OK) return fail("scratch");
    const size_t hist_bytes = (size_t)db->dense_blocks * cfg->n_targets * 4;
    if (cudaMalloc(&db->d_dense_hist, hist_bytes) != cudaSuccess) return fail("dense histograms");
    cudaMemset(db->d_dense_hist, 0, hist_bytes);
    *out = db;
    return CUCLARK_OK;
}

int cuclark_destroy(cuclark_db* db) {
    if (!db) return CUCLARK_OK;
    cudaSetDevice(db->cfg.device);
    cudaDeviceSynchronize();
    free_batches(db);
    text_pipe_free(db);
    route_free(db);
    table_free(db);
    free_scratch(db->scratch);
    cudaFree(db->d_dense_hist);
    if (db->dense_chain) cudaEventDestroy(db->dense_chain);
    if (db->ev0) cudaEventDestroy(db->ev0);
    if (db->ev1) cudaEventDestroy(db->ev1);
    if (db->stream) cudaStreamDestroy(db->stream);
    delete db;
    return CUCLARK_OK;
}

int cuclark_load_db_files(cuclark_db* db, const char* base, int sfactor) {
    if (!db || !base) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    table_free(db);
    return table_build_from_arrays(db, nullptr, nullptr, nullptr, 0, sfactor, base);
}

int cuclark_load_db_arrays(cuclark_db* db, const uint8_t* sz, const void* ky, const uint16_t* lb,
                           uint64_t n_entries, int sfactor) {
    if (!db || !sz || (n_entries && (!ky || !lb))) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    table_free(db);
    return table_build_from_arrays(db, sz, ky, lb, n_entries, sfactor, nullptr);
}

int cuclark_save_table(cuclark_db* db, const char* path) {
    if (!db || !path) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    return table_save(db, path);
}

int cuclark_load_table(cuclark_db* db, const char* path, const char* src_base, int sfactor) {
    if (!db || !path) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    table_free(db);
    return table_load(db, path, src_base, sfactor);
}

int cuclark_clone_table(cuclark_db* src, cuclark_db* dst) {
    if (!src || !dst || src == dst) { set_error("bad argument"); return CUCLARK_ERR_ARG; }
    return table_clone(src, dst);
}

int cuclark_device_info(int device, int* n_devices, uint64_t* free_bytes, uint64_t* total_bytes) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (there is no CPU fallback)");
        return CUCLARK_ERR_NO_DEVICE;
    }
    if (n_devices) *n_devices = n;
    if (free_bytes || total_bytes) {
        if (device < 0 || device >= n) { set_error("device %d not present (%d devices)", device, n); return CUCLARK_ERR_NO_DEVICE; }
        size_t f = 0, t = 0;
        CK(cudaSetDevice(device));
        CK(cudaMemGetInfo(&f, &t));
        if (free_bytes) *free_bytes = f;
        if (total_bytes) *total_bytes = t;
    }
    return CUCLARK_OK;
}

int cuclark_build_db_synthetic(cuclark_db* db, uint32_t seed, uint32_t n_targets, uint64_t genome_len, int light_gap) {
    if (!db) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    if ((int)n_targets > db->cfg.n_targets) { set_error("n_targets exceeds the handle's n_targets"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    table_free(db);
    return table_build_synthetic(db, seed, n_targets, genome_len, light_gap);
}

int cuclark_plan_table(const cuclark_config* cfg_in, uint64_t n_entries, cuclark_table_plan* out) {
    if (!cfg_in || !out) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    if (cfg_in->k < 2 || cfg_in->k > 32) { set_error("The k-mer length should be in [2,32]."); return CUCLARK_ERR_ARG; }
    cuclark_config cfg = *cfg_in;
    if (cfg.shard_count < 1) { cfg.shard_count = 1; cfg.shard_index = 0; }
    if (cfg.layout == 0) {
        const char* e = getenv("CUCLARK_LAYOUT");
        if (e && e[0] >= '1' && e[0] <= '3' && !e[1]) cfg.layout = e[0] - '0';
    }
    table_plan(cfg, n_entries, out);
    return CUCLARK_OK;
}

int cuclark_get_stats(cuclark_db* db, cuclark_stats* s) {
    if (!db || !s) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    memset(s, 0, sizeof *s);
    s->n_entries = db->n_entries;
    s->n_buckets = db->view.M;
    s->n_local_buckets = db->view.n_local;
    s->table_bytes = (db->view.n_local + db->view.n_ovf) * 32;
    s->n_spilled = db->n_spilled;
    s->n_spill_buckets = db->n_spill_buckets;
    s->layout = db->view.layout;
    s->k = db->cfg.k;
    s->lookups = db->last_lookups;
    s->dense_reads = db->last_dense;
    s->truncated_rows = db->last_trunc;
    s->last_kernel_ms = db->last_ms;
    return CUCLARK_OK;
}

/* ---- batches ------------------------------------------------------------------ */
int cuclark_batches_alloc(cuclark_db* db, int n_batches, size_t max_reads, size_t max_containers, int want_rows) {
    if (!db || n_batches < 1) { set_error("bad argument"); return CUCLARK_ERR_ARG; }
    if (max_containers > 0xFFFFFFFFull) { set_error("ERROR: Batch overflow. Please increase the number of batches (-b <numberofbatches>)."); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    free_batches(db);
    db->batch_max_reads = max_reads;
    db->batch_max_cont = max_containers;
    db->batch_rows = want_rows != 0;
    const size_t pitch = 2 * (size_t)db->row_pairs + 2;
    db->batches.resize(n_batches);
    for (auto& b : db->batches) {
        CK(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
        rc = alloc_scratch(b.scratch, 1u << 18, (size_t)db->dense_blocks * db->cfg.n_targets);
        if (rc) return rc;
        CK(cudaMallocHost(&b.h_counters, N_COUNTERS * sizeof(uint32_t)));
        CK(cudaMallocHost(&b.h_ptr, (max_reads + 1) * 4));
        CK(cudaMallocHost(&b.h_cont, (max_containers + 4) * 2));
        CK(cudaMallocHost(&b.h_final, (max_reads + 1) * 10));
        CK(cudaMalloc(&b.d_ptr, (max_reads + 1) * 4));
        CK(cudaMalloc(&b.d_cont, (max_containers + 4) * 2));
        CK(cudaMalloc(&b.d_final, (max_reads + 1) * 10));
        if (want_rows) {
            CK(cudaMallocHost(&b.h_rows, (max_reads + 1) * pitch * 2));
            CK(cudaMalloc(&b.d_rows, (max_reads + 1) * pitch * 2));
        }
    }
    return CUCLARK_OK;
}

int cuclark_batch_buffers(cuclark_db* db, int batch, uint32_t** reads_ptr, uint16_t** containers,
                          uint16_t** final5, uint16_t** rows) {
    if (!db || batch < 0 || batch >= (int)db->batches.size()) { set_error("no such batch"); return CUCLARK_ERR_ARG; }
    Batch& b = db->batches[batch];
    if (reads_ptr) *reads_ptr = b.h_ptr;
    if (containers) *containers = b.h_cont;
    if (final5) *final5 = b.h_final;
    if (rows) *rows = b.h_rows;
    return CUCLARK_OK;
}

int cuclark_batch_ready(cuclark_db* db, int batch, size_t n_reads, size_t n_containers) {
    if (!db || batch < 0 || batch >= (int)db->batches.size()) { set_error("no such batch"); return CUCLARK_ERR_ARG; }
    if (n_reads > db->batch_max_reads || n_containers > db->batch_max_cont) { set_error("batch larger than allocated"); return CUCLARK_ERR_ARG; }
    Batch& b = db->batches[batch];
    b.n_reads = n_reads;
    b.n_cont = n_containers;
    b.ready = true;
    b.queried = false;
    return CUCLARK_OK;
}

int cuclark_batch_query(cuclark_db* db, int batch) {
    if (!db || batch < 0 || batch >= (int)db->batches.size()) { set_error("no such batch"); return CUCLARK_ERR_ARG; }
    Batch& b = db->batches[batch];
    if (!b.ready) { set_error("batch %d is not ready", batch); return CUCLARK_ERR_STATE; }
    int rc = use_device(db);
    if (rc) return rc;
    const size_t pitch = 2 * (size_t)db->row_pairs + 2;
    CK(cudaMemcpyAsync(b.d_ptr, b.h_ptr, (b.n_reads + 1) * 4, cudaMemcpyHostToDevice, b.stream));
    CK(cudaMemcpyAsync(b.d_cont, b.h_cont, b.n_cont * 2, cudaMemcpyHostToDevice, b.stream));
    rc = classify_launch(db, b.scratch, b.d_ptr, b.d_cont, b.n_reads, b.d_final, db->batch_rows ? b.d_rows : nullptr, b.stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(b.h_counters, b.scratch.d_counters, N_COUNTERS * sizeof(uint32_t), cudaMemcpyDeviceToHost, b.stream));
    CK(cudaMemcpyAsync(b.h_final, b.d_final, b.n_reads * 10, cudaMemcpyDeviceToHost, b.stream));
    if (db->batch_rows) CK(cudaMemcpyAsync(b.h_rows, b.d_rows, b.n_reads * pitch * 2, cudaMemcpyDeviceToHost, b.stream));
    CK(cudaEventRecord(b.done, b.stream));
    b.queried = true;
    return CUCLARK_OK;
}

int cuclark_batch_wait(cuclark_db* db, int batch) {
    if (!db || batch < 0 || batch >= (int)db->batches.size()) { set_error("no such batch"); return CUCLARK_ERR_ARG; }
    Batch& b = db->batches[batch];
    if (!b.queried) { set_error("batch %d was not queried", batch); return CUCLARK_ERR_STATE; }
    CK(cudaEventSynchronize(b.done));
    return take_counters(db, b.h_counters, b.scratch.dense_cap);
}

int cuclark_batches_free(cuclark_db* db) {
    if (!db) return CUCLARK_OK;
    cudaSetDevice(db->cfg.device);
    free_batches(db);
    return CUCLARK_OK;
}

/* ---- one-shot calls --------------------------------------------------------------- */
int cuclark_classify_device(cuclark_db* db, const uint32_t* d_reads_ptr, const uint16_t* d_containers,
                            size_t n_reads, uint16_t* d_final5, uint16_t* d_rows, void* stream) {
    if (!db || (!d_reads_ptr && n_reads)) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    return classify_launch(db, db->scratch, d_reads_ptr, d_containers, n_reads, d_final5, d_rows,
                           stream ? (cudaStream_t)stream : db->stream);
}

int cuclark_classify_host(cuclark_db* db, const uint32_t* reads_ptr, const uint16_t* containers, size_t n_reads,
                          uint16_t* final5, uint16_t* rows) {
    if (!db || !reads_ptr || (!final5 && !rows)) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    const size_t n_cont = reads_ptr[n_reads];
    if (n_cont && !containers) { set_error("null containers"); return CUCLARK_ERR_ARG; }
    const size_t pitch = 2 * (size_t)db->row_pairs + 2;
    uint32_t* d_ptr = nullptr; uint16_t *d_cont = nullptr, *d_final = nullptr, *d_rows = nullptr;
    cudaStream_t st = db->stream;
    auto cleanup = [&]() { cudaFree(d_ptr); cudaFree(d_cont); cudaFree(d_final); cudaFree(d_rows); };
#define CKF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(e_)); cleanup(); return CUCLARK_ERR_CUDA; } } while (0)
    CKF(cudaMalloc(&d_ptr, (n_reads + 1) * 4));
    CKF(cudaMalloc(&d_cont, (n_cont + 4) * 2));
    if (final5) CKF(cudaMalloc(&d_final, (n_reads + 1) * 10));
    if (rows) CKF(cudaMalloc(&d_rows, (n_reads + 1) * pitch * 2));
    CKF(cudaMemcpyAsync(d_ptr, reads_ptr, (n_reads + 1) * 4, cudaMemcpyHostToDevice, st));
    if (n_cont) CKF(cudaMemcpyAsync(d_cont, containers, n_cont * 2, cudaMemcpyHostToDevice, st));
    CKF(cudaEventRecord(db->ev0, st));
    rc = classify_launch(db, db->scratch, d_ptr, d_cont, n_reads, d_final, d_rows, st);
    if (rc) { cleanup(); return rc; }
    CKF(cudaEventRecord(db->ev1, st));
    if (final5) CKF(cudaMemcpyAsync(final5, d_final, n_reads * 10, cudaMemcpyDeviceToHost, st));
    if (rows) CKF(cudaMemcpyAsync(rows, d_rows, n_reads * pitch * 2, cudaMemcpyDeviceToHost, st));
    rc = fetch_counters(db, st);
    float ms = 0;
    cudaEventElapsedTime(&ms, db->ev0, db->ev1);
    db->last_ms = ms;
    cleanup();
#undef CKF
    return rc;
}

int cuclark_merge_rows_device(cuclark_db* db, const uint16_t* d_rows_parts, int n_parts, size_t n_reads,
                              uint16_t* d_rows_out, uint16_t* d_final5, void* stream) {
    if (!db || !d_rows_parts) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    return merge_rows_launch(db, d_rows_parts, n_parts, n_reads, d_rows_out, d_final5,
                             stream ? (cudaStream_t)stream : db->stream);
}

int cuclark_synth_reads_device(cuclark_db* db, uint32_t seed, uint32_t genome_seed, uint32_t n_targets,
                               uint64_t genome_len, uint64_t first_read, size_t n_reads, int read_len,
                               int pct_random, int sub_per_10k, uint32_t* d_reads_ptr, uint16_t* d_containers,
                               void* stream) {
    if (!db || !d_reads_ptr || !d_containers) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    return synth_reads_launch(seed, genome_seed, n_targets, genome_len, first_read, n_reads, read_len, pct_random,
                              sub_per_10k, d_reads_ptr, d_containers, stream ? (cudaStream_t)stream : db->stream);
}

int cuclark_synth_fastq_device(cuclark_db* db, uint32_t seed, uint32_t genome_seed, uint32_t n_targets,
                               uint64_t genome_len, uint64_t first_read, size_t n_reads, int read_len,
                               int pct_random, int sub_per_10k, uint8_t* d_text, void* stream) {
    if (!db || !d_text) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    return synth_fastq_launch(seed, genome_seed, n_targets, genome_len, first_read, n_reads, read_len, pct_random,
                              sub_per_10k, d_text, stream ? (cudaStream_t)stream : db->stream);
}

int cuclark_synth_fastq_pair_device(cuclark_db* db, uint32_t seed, uint32_t genome_seed, uint32_t n_targets,
                                    uint64_t genome_len, uint64_t first_read, size_t n_reads, int read_len,
                                    int pct_random, int sub_per_10k, int mate, uint8_t* d_text, void* stream) {
    if (!db || !d_text || (mate != 1 && mate != 2)) { set_error("bad argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    return synth_fastq_launch(seed, genome_seed, n_targets, genome_len, first_read, n_reads, read_len, pct_random,
                              sub_per_10k, d_text, stream ? (cudaStream_t)stream : db->stream, mate);
}

int cuclark_gather_bench(cuclark_db* db, uint64_t n_probes, int bytes_per_probe, int ilp, int iters, double* ms_out) {
    if (!db || !ms_out || iters < 1) { set_error("bad argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    return gather_bench_launch(db, n_probes, bytes_per_probe, ilp, iters, ms_out);
}

/* ---- table-partitioned mode by k-mer routing (route.cu) ---------------------------------------------- */
int cuclark_route_alloc(cuclark_db* db, int n_ranks, size_t max_containers) {
    if (!db) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    return rc ? rc : route_alloc(db, n_ranks, max_containers);
}
int cuclark_route_free(cuclark_db* db) {
    if (db) route_free(db);
    return CUCLARK_OK;
}
int cuclark_route_export(cuclark_db* db, void* handle64, uint64_t* region_bytes) {
    if (!db || !handle64) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    return rc ? rc : route_export(db, handle64, region_bytes);
}
int cuclark_route_import(cuclark_db* db, int peer_rank, const void* handle64) {
    if (!db || !handle64) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    return rc ? rc : route_import(db, peer_rank, handle64);
}
int cuclark_route_connect(cuclark_db* const* dbs, int n) {
    if (!dbs || n < 1 || n > CUCLARK_ROUTE_MAX_RANKS) { set_error("bad argument"); return CUCLARK_ERR_ARG; }
    return route_connect(dbs, n);
}
int cuclark_route_scatter(cuclark_db* db, const uint32_t* d_reads_ptr, const uint16_t* d_containers, size_t n_reads,
                          size_t n_containers, void* stream) {
    if (!db || (!d_reads_ptr && n_reads)) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    return rc ? rc : route_scatter(db, d_reads_ptr, d_containers, n_reads, n_containers, stream ? (cudaStream_t)stream : db->stream);
}
int cuclark_route_probe(cuclark_db* db, void* stream) {
    if (!db) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    return rc ? rc : route_probe(db, stream ? (cudaStream_t)stream : db->stream);
}
int cuclark_route_gather(cuclark_db* db, const uint32_t* d_reads_ptr, const uint16_t* d_containers, size_t n_reads,
                         size_t n_containers, uint16_t* d_final5, uint16_t* d_rows, void* stream) {
    if (!db || (!d_reads_ptr && n_reads)) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    return rc ? rc : route_gather(db, db->scratch, d_reads_ptr, d_containers, n_reads, n_containers, d_final5, d_rows,
                                  stream ? (cudaStream_t)stream : db->stream);
}
int cuclark_route_get_stats(cuclark_db* db, cuclark_route_stats* out) {
    if (!db || !out) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    return route_stats(db, out);
}

// scatter on every rank | all ranks wait for all scatters | probe | all wait for all probes | gather
int cuclark_classify_routed_device(cuclark_db* const* dbs, int n, const uint32_t* const* d_reads_ptr,
                                   const uint16_t* const* d_containers, const size_t* n_reads, const size_t* n_containers,
                                   uint16_t* const* d_final5, uint16_t* const* d_rows) {
    if (!dbs || n < 1 || n > CUCLARK_ROUTE_MAX_RANKS || !d_reads_ptr || !d_containers || !n_reads || !n_containers) {
        set_error("bad argument");
        return CUCLARK_ERR_ARG;
    }
    cudaEvent_t ev[2][CUCLARK_ROUTE_MAX_RANKS] = {};
    int rc = CUCLARK_OK;
    auto done = [&](int r) {
        for (int ph = 0; ph < 2; ph++)
            for (int i = 0; i < n; i++)
                if (ev[ph][i]) { cudaSetDevice(dbs[i]->cfg.device); cudaEventDestroy(ev[ph][i]); }
        return r;
    };
    for (int i = 0; i < n && !rc; i++) {
        if (!dbs[i]) { set_error("null handle"); return done(CUCLARK_ERR_ARG); }
        if ((rc = use_device(dbs[i]))) break;
        for (int ph = 0; ph < 2; ph++)
            if (cudaEventCreateWithFlags(&ev[ph][i], cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate failed"); return done(CUCLARK_ERR_CUDA); }
    }
    for (int i = 0; i < n && !rc; i++) {
        if ((rc = use_device(dbs[i]))) break;
        rc = route_scatter(dbs[i], d_reads_ptr[i], d_containers[i], n_reads[i], n_containers[i], dbs[i]->stream);
        if (!rc && cudaEventRecord(ev[0][i], dbs[i]->stream) != cudaSuccess) { set_error("cudaEventRecord failed"); rc = CUCLARK_ERR_CUDA; }
    }
    for (int i = 0; i < n && !rc; i++) {
        if ((rc = use_device(dbs[i]))) break;
        for (int j = 0; j < n; j++)
            if (j != i && cudaStreamWaitEvent(dbs[i]->stream, ev[0][j], 0) != cudaSuccess) { set_error("cudaStreamWaitEvent failed"); rc = CUCLARK_ERR_CUDA; }
        if (!rc) rc = route_probe(dbs[i], dbs[i]->stream);
        if (!rc && cudaEventRecord(ev[1][i], dbs[i]->stream) != cudaSuccess) { set_error("cudaEventRecord failed"); rc = CUCLARK_ERR_CUDA; }
    }
    for (int i = 0; i < n && !rc; i++) {
        if ((rc = use_device(dbs[i]))) break;
        for (int j = 0; j < n; j++)
            if (j != i && cudaStreamWaitEvent(dbs[i]->stream, ev[1][j], 0) != cudaSuccess) { set_error("cudaStreamWaitEvent failed"); rc = CUCLARK_ERR_CUDA; }
        if (!rc) rc = route_gather(dbs[i], dbs[i]->scratch, d_reads_ptr[i], d_containers[i], n_reads[i], n_containers[i],
                                   d_final5 ? d_final5[i] : nullptr, d_rows ? d_rows[i] : nullptr, dbs[i]->stream);
    }
    // the events may be destroyed once the work that waits on them is enqueued; the counters come back with the sync
    for (int i = 0; i < n && !rc; i++) {
        if ((rc = use_device(dbs[i]))) break;
        rc = fetch_counters(dbs[i], dbs[i]->stream);
    }
    return done(rc);
}

/* counters of the last classify on `stream` (synchronises that stream) */
int cuclark_sync_stats(cuclark_db* db, void* stream) {
    if (!db) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    int rc = use_device(db);
    if (rc) return rc;
    return fetch_counters(db, stream ? (cudaStream_t)stream : db->stream);
}

}  // extern "C"
