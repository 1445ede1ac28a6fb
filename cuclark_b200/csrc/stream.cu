// cuclark_b200 — the text pipeline: raw FASTA/FASTQ bytes in, result CSV out.
//
// Replaces the body of CuCLARK::getObjectsDataComputeFullGPU + printExtendedResultsSynced
// (src/CuCLARK_hh.hh:1335-1790, 1951-2139). The reference indexes and packs on the
// host (OpenMP over batches), copies packed batches to the device, and prints one line
// per read with fprintf. Here the host only cuts the input into chunks at record
// boundaries and moves bytes; every chunk goes through
//     H2D text -> line table -> records -> 2-bit pack -> classify -> CSV text -> D2H
// on its own stream, `n_slots` chunks in flight (one host thread per slot), and the
// CSV pieces are handed to the sink in file order.
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "internal.h"
#include "textpipe.cuh"

namespace cuclark {

struct TextSlot {
    TextSlotDev d;
    uint8_t* h_text = nullptr;       // pinned staging for pageable sources
    char* h_csv = nullptr;           // pinned
    ChunkInfo* h_info = nullptr;     // pinned
    uint32_t* h_counters = nullptr;  // pinned
    Scratch scratch;
    cudaStream_t stream = nullptr;
};

struct TextPipe {
    std::vector<TextSlot> slots;
    size_t chunk_bytes = 0;
    bool extended = false;
    int row_pairs = 0;
    // device name table
    char* d_name_chars = nullptr;
    uint32_t* d_name_off = nullptr;
    NameTable names;
    std::vector<std::string> host_names;   // [0] = "NA"
};

namespace {

template <typename T>
int dmalloc(T*& p, size_t count) {
    CK(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
    return CUCLARK_OK;
}

void free_slot(TextSlot& s) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    TextSlotDev& d = s.d;
    cudaFree(d.text); cudaFree(d.line_start); cudaFree(d.hdr_line); cudaFree(d.name_s); cudaFree(d.name_e);
    cudaFree(d.seq_s); cudaFree(d.seq_e); cudaFree(d.len); cudaFree(d.reads_ptr); cudaFree(d.cont);
    cudaFree(d.final5); cudaFree(d.rows); cudaFree(d.csv_off); cudaFree(d.csv); cudaFree(d.tile_a); cudaFree(d.tile_b);
    cudaFree(d.info);
    cudaFreeHost(s.h_text); cudaFreeHost(s.h_csv); cudaFreeHost(s.h_info); cudaFreeHost(s.h_counters);
    cudaFree(s.scratch.d_counters); cudaFree(s.scratch.d_dense_list);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = TextSlot{};
}

int alloc_slot(TextSlot& s, size_t C, bool extended, int row_pairs) {
    TextSlotDev& d = s.d;
    d.cap_bytes = C;
    d.cap_lines = C / 6 + 64;
    d.cap_reads = C / 16 + 64;
    d.cap_cont = C / 2 + 1024;
    d.cap_csv = std::max<size_t>(2 * C, 1 << 20);
    d.cap_tiles = ((std::max(C / 4096, d.cap_reads / 2048) + 8) | 1) + 1;      // even, so the totals behind it are aligned
    int rc;
#define A(call) do { rc = (call); if (rc) return rc; } while (0)
    A(dmalloc(d.text, C + 256));
    A(dmalloc(d.line_start, d.cap_lines + 2));
    A(dmalloc(d.hdr_line, d.cap_reads + 1));
    A(dmalloc(d.name_s, d.cap_reads)); A(dmalloc(d.name_e, d.cap_reads));
    A(dmalloc(d.seq_s, d.cap_reads)); A(dmalloc(d.seq_e, d.cap_reads)); A(dmalloc(d.len, d.cap_reads));
    A(dmalloc(d.reads_ptr, d.cap_reads + 1));
    A(dmalloc(d.cont, d.cap_cont + 8));
    A(dmalloc(d.final5, (d.cap_reads + 1) * 5));
    if (extended) A(dmalloc(d.rows, (d.cap_reads + 1) * (size_t)(2 * row_pairs + 2)));
    A(dmalloc(d.csv_off, d.cap_reads + 1));
    A(dmalloc(d.csv, d.cap_csv));
    A(dmalloc(d.tile_a, d.cap_tiles + 4)); A(dmalloc(d.tile_b, d.cap_tiles + 4));
    A(dmalloc(d.info, 1));
#undef A
    CK(cudaMemset(d.text, '\n', C + 256));
    CK(cudaMallocHost(&s.h_text, C));
    CK(cudaMallocHost(&s.h_csv, d.cap_csv));
    CK(cudaMallocHost(&s.h_info, sizeof(ChunkInfo)));
    CK(cudaMallocHost(&s.h_counters, N_COUNTERS * sizeof(uint32_t)));
    s.scratch.dense_cap = 1u << 16;
    CK(cudaMalloc(&s.scratch.d_counters, N_COUNTERS * sizeof(uint32_t)));
    CK(cudaMalloc(&s.scratch.d_dense_list, (size_t)s.scratch.dense_cap * 4));
    CK(cudaMemset(s.scratch.d_counters, 0, N_COUNTERS * sizeof(uint32_t)));
    CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    return CUCLARK_OK;
}

}  // namespace

void text_pipe_free(cuclark_db* db) {
    TextPipe* tp = db->text_pipe;
    if (!tp) return;
    for (auto& s : tp->slots) free_slot(s);
    cudaFree(tp->d_name_chars); cudaFree(tp->d_name_off);
    delete tp;
    db->text_pipe = nullptr;
}

namespace {

int ensure_pipe(cuclark_db* db, size_t chunk_bytes, int n_slots, bool extended, const char* const* target_names) {
    TextPipe* tp = db->text_pipe;
    if (tp && (tp->chunk_bytes != chunk_bytes || (int)tp->slots.size() != n_slots || tp->extended != extended ||
               tp->row_pairs != db->row_pairs)) {
        text_pipe_free(db);
        tp = nullptr;
    }
    if (!tp) {
        tp = new TextPipe();
        db->text_pipe = tp;
        tp->chunk_bytes = chunk_bytes; tp->extended = extended; tp->row_pairs = db->row_pairs;
        tp->slots.resize(n_slots);
        for (auto& s : tp->slots) {
            int rc = alloc_slot(s, chunk_bytes, extended, db->row_pairs);
            if (rc) { text_pipe_free(db); return rc; }
        }
    }
    // names: [0] = "NA" (src/CuCLARK_hh.hh:1879-1883)
    std::vector<std::string> names;
    names.push_back("NA");
    for (int t = 0; t < db->cfg.n_targets; t++) {
        if (target_names && target_names[t]) names.push_back(target_names[t]);
        else names.push_back("T" + std::to_string(t));
    }
    if (names != tp->host_names) {
        cudaFree(tp->d_name_chars); cudaFree(tp->d_name_off);
        tp->d_name_chars = nullptr; tp->d_name_off = nullptr;
        std::string chars;
        std::vector<uint32_t> off;
        uint32_t max_len = 0;
        for (auto& s : names) { off.push_back((uint32_t)chars.size()); chars += s; max_len = std::max<uint32_t>(max_len, (uint32_t)s.size()); }
        off.push_back((uint32_t)chars.size());
        CK(cudaMalloc(&tp->d_name_chars, chars.size() + 1));
        CK(cudaMalloc(&tp->d_name_off, off.size() * 4));
        CK(cudaMemcpy(tp->d_name_chars, chars.data(), chars.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(tp->d_name_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
        tp->names.chars = tp->d_name_chars; tp->names.off = tp->d_name_off;
        tp->names.n_names = (uint32_t)names.size(); tp->names.max_len = max_len;
        tp->host_names = names;
    }
    return CUCLARK_OK;
}

const char* tp_err_text(uint32_t e) {
    if (e & TP_ERR_LINES) return "chunk holds more lines than the slot can index (lines shorter than 6 bytes on average); raise chunk_bytes";
    if (e & TP_ERR_READS) return "chunk holds more reads than the slot can hold (records shorter than 16 bytes on average); raise chunk_bytes";
    if (e & TP_ERR_CONT) return "packed reads of the chunk exceed the container buffer";
    if (e & TP_ERR_CSV) return "CSV text of the chunk exceeds the output buffer";
    return "unknown";
}

// One run over a text buffer, shared by the worker threads.
struct Job {
    cuclark_db* db;
    TextPipe* tp;
    const uint8_t* text;
    size_t n;
    bool fastq, paired, extended, src_pinned;
    cuclark_sink_fn sink;
    void* user;
    cuclark_text_arrays* arrays;     // debug hook: collect index/containers/results instead of CSV
    // chunking
    std::mutex mu;
    std::condition_variable cv;
    size_t cursor = 0;
    uint64_t next_seq = 0, turn = 0;
    bool failed = false;
    int rc = CUCLARK_OK;
    std::string err;
    // totals (guarded by the turn)
    uint64_t n_reads = 0, n_cont = 0, lookups = 0, csv_bytes = 0, dense = 0, trunc = 0, n_chunks = 0;

    void fail(int code, const std::string& msg) {
        std::lock_guard<std::mutex> g(mu);
        if (!failed) { failed = true; rc = code; err = msg; }
        cv.notify_all();
    }
};

// Largest end <= want that is the first byte of a record (or n).
// FASTA: '>' at a line start (src/CuCLARK_hh.hh:1376). FASTQ: a line starting with '@' whose
// line after next starts with '+': in 4-line FASTQ only a header satisfies this (a quality
// line starting with '@' is followed by a header and then a sequence line, which never
// starts with '+'); the reference snaps its batch starts with a similar test (:1430-1471).
size_t record_boundary(const uint8_t* t, size_t n, size_t start, size_t want, bool fastq) {
    if (want >= n) return n;
    for (size_t p = want; p > start + 1; p--) {
        if (t[p - 1] != '\n') continue;
        if (!fastq) { if (t[p] == '>') return p; continue; }
        if (t[p] != '@') continue;
        const uint8_t* l1 = (const uint8_t*)memchr(t + p, '\n', n - p);
        if (!l1) continue;
        const uint8_t* l2 = (const uint8_t*)memchr(l1 + 1, '\n', n - (size_t)(l1 + 1 - t));
        if (!l2 || (size_t)(l2 + 1 - t) >= n) continue;
        if (l2[1] == '+') return p;
    }
    return start;      // no boundary inside the window: a record larger than a chunk
}

#define JCK(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            J.fail(CUCLARK_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e_));          \
            return;                                                                                       \
        }                                                                                                 \
    } while (0)
#define JRC(call)                                                                                         \
    do {                                                                                                  \
        int r_ = (call);                                                                                  \
        if (r_) { J.fail(r_, cuclark_last_error()); return; }                                             \
    } while (0)

// waits until it is chunk `seq`'s turn to emit; false if the job failed meanwhile
bool wait_turn(Job& J, uint64_t seq) {
    std::unique_lock<std::mutex> lk(J.mu);
    J.cv.wait(lk, [&] { return J.failed || J.turn == seq; });
    return !J.failed;
}
void end_turn(Job& J) {
    std::lock_guard<std::mutex> g(J.mu);
    J.turn++;
    J.cv.notify_all();
}

template <typename T>
bool copy_out(Job& J, T* dst, size_t cap, uint64_t at, const void* dsrc, size_t count, cudaStream_t st) {
    if (!dst) return true;
    if (at + count > cap) { J.fail(CUCLARK_ERR_NOMEM, "cuclark_text_debug: output array too small"); return false; }
    if (count && cudaMemcpyAsync(dst + at, dsrc, count * sizeof(T), cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        J.fail(CUCLARK_ERR_CUDA, "D2H failed");
        return false;
    }
    return true;
}

void worker(Job& J, TextSlot& S) {
    cuclark_db* db = J.db;
    if (cudaSetDevice(db->cfg.device) != cudaSuccess) { J.fail(CUCLARK_ERR_CUDA, "cudaSetDevice failed"); return; }
    const TextSlotDev& d = S.d;
    const int k = db->cfg.k;
    const size_t pitch = 2 * (size_t)db->row_pairs + 2;
    const bool want_rows = J.extended || (J.arrays && J.arrays->rows);
    for (;;) {
        size_t start, end;
        uint64_t seq;
        {
            std::lock_guard<std::mutex> g(J.mu);
            if (J.failed || J.cursor >= J.n) return;
            start = J.cursor;
            end = record_boundary(J.text, J.n, start, std::min(J.n, start + d.cap_bytes), J.fastq);
            if (end <= start) {
                J.failed = true; J.rc = CUCLARK_ERR_ARG;
                J.err = "a single record is larger than chunk_bytes (" + std::to_string(d.cap_bytes) + "); raise chunk_bytes";
                J.cv.notify_all();
                return;
            }
            J.cursor = end;
            seq = J.next_seq++;
        }
        const uint32_t nb = (uint32_t)(end - start);
        const uint8_t* src = J.text + start;
        if (!J.src_pinned) { memcpy(S.h_text, src, nb); src = S.h_text; }
        JCK(cudaMemcpyAsync(d.text, src, nb, cudaMemcpyHostToDevice, S.stream));
        JRC(tp_index_launch(d, nb, J.fastq, S.stream));
        JCK(cudaMemcpyAsync(S.h_info, d.info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, S.stream));
        JCK(cudaStreamSynchronize(S.stream));
        if (S.h_info->err) { J.fail(CUCLARK_ERR_NOMEM, tp_err_text(S.h_info->err)); return; }
        const uint32_t n_reads = S.h_info->n_reads;
        JRC(tp_pack_launch(d, nb, n_reads, k, S.stream));
        JCK(cudaMemcpyAsync(S.h_info, d.info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, S.stream));
        JCK(cudaStreamSynchronize(S.stream));
        if (S.h_info->err) { J.fail(CUCLARK_ERR_NOMEM, tp_err_text(S.h_info->err)); return; }
        const uint64_t n_cont = S.h_info->n_cont;
        const bool classify = !J.arrays || J.arrays->final5 || J.arrays->rows;
        if (classify) {
            JRC(classify_launch(db, S.scratch, d.reads_ptr, d.cont, n_reads, d.final5, want_rows ? d.rows : nullptr, S.stream));
            JCK(cudaMemcpyAsync(S.h_counters, S.scratch.d_counters, N_COUNTERS * sizeof(uint32_t), cudaMemcpyDeviceToHost, S.stream));
        }
        if (J.arrays) {
            // ---- debug hook: hand the intermediate arrays back, in file order ----
            JCK(cudaStreamSynchronize(S.stream));
            if (!wait_turn(J, seq)) return;
            cuclark_text_arrays& A = *J.arrays;
            const uint64_t r0 = J.n_reads, c0 = J.n_cont;
            std::vector<uint32_t> tmp(n_reads + 1);
            auto widen = [&](uint64_t* dst, const uint32_t* dsrc, uint64_t add) -> bool {
                if (!dst) return true;
                if (r0 + n_reads > A.cap_reads) { J.fail(CUCLARK_ERR_NOMEM, "cuclark_text_debug: output array too small"); return false; }
                if (cudaMemcpy(tmp.data(), dsrc, (size_t)n_reads * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { J.fail(CUCLARK_ERR_CUDA, "D2H failed"); return false; }
                for (uint32_t i = 0; i < n_reads; i++) dst[r0 + i] = (uint64_t)tmp[i] + add;
                return true;
            };
            bool ok = widen(A.name_s, d.name_s, start) && widen(A.name_e, d.name_e, start) && widen(A.seq_s, d.seq_s, start) &&
                      widen(A.seq_e, d.seq_e, start) && widen(A.len, d.len, 0);
            if (ok && A.reads_ptr) {
                if (r0 + n_reads + 1 > A.cap_reads + 1 || c0 + n_cont > 0xFFFFFFFFull) { J.fail(CUCLARK_ERR_NOMEM, "cuclark_text_debug: output array too small"); ok = false; }
                else if (cudaMemcpy(tmp.data(), d.reads_ptr, (size_t)(n_reads + 1) * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { J.fail(CUCLARK_ERR_CUDA, "D2H failed"); ok = false; }
                else for (uint32_t i = 0; i <= n_reads; i++) A.reads_ptr[r0 + i] = (uint32_t)(c0 + tmp[i]);
            }
            ok = ok && copy_out(J, A.containers, A.cap_containers, c0, d.cont, n_cont, S.stream);
            ok = ok && copy_out(J, A.final5, A.cap_reads * 5, r0 * 5, d.final5, (size_t)n_reads * 5, S.stream);
            ok = ok && copy_out(J, A.rows, A.cap_reads * pitch, r0 * pitch, d.rows, (size_t)n_reads * pitch, S.stream);
            if (ok && cudaStreamSynchronize(S.stream) != cudaSuccess) { J.fail(CUCLARK_ERR_CUDA, "sync failed"); ok = false; }
            if (!ok) return;
            J.n_reads += n_reads; J.n_cont += n_cont; J.n_chunks++;
            if (classify) {
                J.lookups += (uint64_t)S.h_counters[COUNTER_LOOKUPS] | ((uint64_t)S.h_counters[COUNTER_LOOKUPS + 1] << 32);
                J.dense += S.h_counters[COUNTER_DENSE]; J.trunc += S.h_counters[COUNTER_TRUNC];
            }
            end_turn(J);
            continue;
        }
        // ---- CSV text, in groups that fit the output buffer ----
        const size_t max_line = 39 + (J.extended ? 2 * (size_t)db->cfg.n_targets + 4 * (size_t)db->row_pairs : 0) + 64 +
                                2 * (size_t)J.tp->names.max_len;
        const uint32_t group = (uint32_t)std::min<size_t>(std::max<size_t>(d.cap_csv / max_line, 1), 0x7FFFFFFF);
        bool have_turn = false;
        for (uint32_t first = 0; first < n_reads; first += group) {
            const uint32_t cnt = std::min(group, n_reads - first);
            JRC(tp_csv_launch(d, J.tp->names, first, cnt, k, J.paired, J.extended, db->row_pairs, (uint32_t)db->cfg.n_targets, S.stream));
            JCK(cudaMemcpyAsync(S.h_info, d.info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, S.stream));
            JCK(cudaStreamSynchronize(S.stream));
            if (S.h_info->err) { J.fail(CUCLARK_ERR_NOMEM, tp_err_text(S.h_info->err)); return; }
            const size_t bytes = S.h_info->csv_bytes;
            if (bytes) {
                JCK(cudaMemcpyAsync(S.h_csv, d.csv, bytes, cudaMemcpyDeviceToHost, S.stream));
                JCK(cudaStreamSynchronize(S.stream));
            }
            if (!have_turn) { if (!wait_turn(J, seq)) return; have_turn = true; }
            if (bytes && J.sink && J.sink(J.user, S.h_csv, bytes) != 0) {
                J.fail(CUCLARK_ERR_IO, "the CSV sink reported an error");
                return;
            }
            J.csv_bytes += bytes;
        }
        if (!have_turn && !wait_turn(J, seq)) return;
        if (S.h_counters[COUNTER_DENSE] > S.scratch.dense_cap) { J.fail(CUCLARK_ERR_NOMEM, "too many reads needed the dense fallback"); return; }
        J.n_reads += n_reads; J.n_cont += n_cont; J.n_chunks++;
        J.lookups += (uint64_t)S.h_counters[COUNTER_LOOKUPS] | ((uint64_t)S.h_counters[COUNTER_LOOKUPS + 1] << 32);
        J.dense += S.h_counters[COUNTER_DENSE]; J.trunc += S.h_counters[COUNTER_TRUNC];
        end_turn(J);
    }
}

int run_text(cuclark_db* db, const uint8_t* text, size_t n, const cuclark_text_opts* o, cuclark_sink_fn sink, void* user,
             cuclark_text_arrays* arrays, cuclark_text_stats* out) {
    if (!db || (!text && n)) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    if (!db->d_table) { set_error("no database loaded"); return CUCLARK_ERR_STATE; }
    if (out) memset(out, 0, sizeof *out);
    const auto t0 = std::chrono::steady_clock::now();
    if (n == 0 || (text[0] != '>' && text[0] != '@')) {
        set_error("Failed to recognize the format of the file.");            // src/CuCLARK_hh.hh:1535-1538
        return CUCLARK_ERR_FORMAT;
    }
    CK(cudaSetDevice(db->cfg.device));
    size_t chunk = o && o->chunk_bytes ? o->chunk_bytes : (size_t)64 << 20;
    chunk = std::min<size_t>(std::max<size_t>(chunk, 4096), (size_t)1 << 30);
    chunk = (chunk + 255) & ~(size_t)255;
    int n_slots = o && o->n_slots > 0 ? o->n_slots : 4;
    n_slots = std::min(n_slots, 16);
    const bool extended = o && o->extended;
    int rc = ensure_pipe(db, chunk, n_slots, extended || (arrays && arrays->rows), o ? o->target_names : nullptr);
    if (rc) return rc;
    Job J;
    J.db = db; J.tp = db->text_pipe; J.text = text; J.n = n;
    J.fastq = text[0] == '@';
    J.paired = o && o->paired; J.extended = extended;
    J.sink = sink; J.user = user; J.arrays = arrays;
    cudaPointerAttributes attr;
    J.src_pinned = cudaPointerGetAttributes(&attr, text) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (sink && !arrays) {
        // header line (src/CuCLARK_hh.hh:1957-1972)
        std::string h = "Object_ID";
        if (extended) for (size_t t = 1; t < J.tp->host_names.size(); t++) { h += ","; h += J.tp->host_names[t]; }
        h += ",Length,Gamma,1st_assignment,score1,2nd_assignment,score2,confidence\n";
        if (sink(user, h.data(), h.size()) != 0) { set_error("the CSV sink reported an error"); return CUCLARK_ERR_IO; }
        J.csv_bytes += h.size();
    }
    const size_t n_chunks_est = (n + chunk - 1) / chunk;
    const int n_threads = (int)std::min<size_t>(n_slots, std::max<size_t>(n_chunks_est, 1));
    std::vector<std::thread> threads;
    for (int i = 1; i < n_threads; i++) threads.emplace_back([&J, i] { worker(J, J.tp->slots[i]); });
    worker(J, J.tp->slots[0]);
    for (auto& t : threads) t.join();
    if (J.failed) { set_error("%s", J.err.c_str()); return J.rc; }
    db->last_lookups = J.lookups; db->last_dense = J.dense; db->last_trunc = J.trunc;
    if (out) {
        out->n_reads = J.n_reads; out->lookups = J.lookups; out->csv_bytes = J.csv_bytes; out->n_chunks = J.n_chunks;
        out->n_containers = J.n_cont; out->dense_reads = J.dense; out->truncated_rows = J.trunc;
        out->text_bytes = n;
        out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return CUCLARK_OK;
}

int file_sink(void* user, const char* data, size_t n) {
    return fwrite(data, 1, n, (FILE*)user) == n ? 0 : -1;
}

}  // namespace
}  // namespace cuclark

using namespace cuclark;

extern "C" {

int cuclark_classify_text(cuclark_db* db, const uint8_t* text, size_t n, const cuclark_text_opts* opts,
                          cuclark_sink_fn sink, void* user, cuclark_text_stats* out) {
    return run_text(db, text, n, opts, sink, user, nullptr, out);
}

int cuclark_text_debug(cuclark_db* db, const uint8_t* text, size_t n, const cuclark_text_opts* opts,
                       cuclark_text_arrays* arrays, cuclark_text_stats* out) {
    if (!arrays) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    return run_text(db, text, n, opts, nullptr, nullptr, arrays, out);
}

int cuclark_classify_file(cuclark_db* db, const char* objects_path, const char* csv_path, const cuclark_text_opts* opts,
                          cuclark_text_stats* out) {
    if (!db || !objects_path || !csv_path) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    const int fd = open(objects_path, O_RDONLY);
    struct stat sb;
    if (fd < 0 || fstat(fd, &sb) != 0 || sb.st_size == 0) {
        if (fd >= 0) close(fd);
        set_error("Failed to open %s", objects_path);                        // src/CuCLARK_hh.hh:524-528
        return CUCLARK_ERR_IO;
    }
    const size_t n = (size_t)sb.st_size;
    void* map = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) { close(fd); set_error("Failed to mmapping the file."); return CUCLARK_ERR_IO; }
    madvise(map, n, MADV_SEQUENTIAL);
    FILE* f = fopen(csv_path, "w");
    if (!f) { munmap(map, n); close(fd); set_error("Failed to create/open file result: %s", csv_path); return CUCLARK_ERR_IO; }
    std::vector<char> iobuf(8 << 20);
    setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
    int rc = run_text(db, (const uint8_t*)map, n, opts, file_sink, f, nullptr, out);
    if (fclose(f) != 0 && rc == CUCLARK_OK) { set_error("failed to write %s", csv_path); rc = CUCLARK_ERR_IO; }
    munmap(map, n);
    close(fd);
    return rc;
}

}  // extern "C"
