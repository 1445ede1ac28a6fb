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
    void* d_arena = nullptr;         // every device array of the slot is carved from ONE allocation,
    void* h_arena = nullptr;         //   every pinned one from another: a slot costs two driver calls
    bool with_stage = false;         // h_text is present
};

struct TextPipe {
    std::vector<TextSlot> slots;     // allocated lazily, each by the thread that first uses it
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

void free_slot(TextSlot& s) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.d_arena);
    cudaFreeHost(s.h_arena);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = TextSlot{};
}

// bump allocator over one arena: first pass (base == nullptr) only sizes it
struct Carver {
    uint8_t* base;
    size_t off = 0;
    template <typename T>
    void take(T*& p, size_t count) {
        off = (off + 255) & ~(size_t)255;
        p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
    }
};

// Pinned memory costs ~0.6 ms per MiB to allocate on the B200 host (16 slots of 64 MiB chunks took
// 3.6 s, the classification of 2 M reads 40 ms), so a slot is sized tightly and allocated by its own
// thread while the other slots already work. `with_stage`: the source is pageable and needs h_text.
int alloc_slot(TextSlot& s, size_t C, bool extended, int row_pairs, bool with_stage, size_t hist_words) {
    TextSlotDev& d = s.d;
    d.cap_bytes = C;
    d.cap_lines = C / 6 + 64;
    d.cap_reads = C / 16 + 64;
    d.cap_cont = C / 2 + 1024;
    // CSV of a chunk that does not fit is produced in groups (worker()): half a chunk holds a whole
    // chunk of FASTQ or of FASTA reads of >= 100 bp
    d.cap_csv = std::max<size_t>(C / 2, 1 << 20);
    d.cap_tiles = ((std::max(C / 4096, d.cap_reads / 2048) + 8) | 1) + 1;      // even, so the totals behind it are aligned
    s.scratch.dense_cap = 1u << 16;
    auto carve_dev = [&](Carver& c) {
        c.take(d.text, C + 256);
        c.take(d.line_start, d.cap_lines + 2);
        c.take(d.hdr_line, d.cap_reads + 1);
        c.take(d.name_s, d.cap_reads); c.take(d.name_e, d.cap_reads);
        c.take(d.seq_s, d.cap_reads); c.take(d.seq_e, d.cap_reads); c.take(d.len, d.cap_reads);
        c.take(d.reads_ptr, d.cap_reads + 1);
        c.take(d.cont, d.cap_cont + 8);
        c.take(d.final5, (d.cap_reads + 1) * 5);
        if (extended) c.take(d.rows, (d.cap_reads + 1) * (size_t)(2 * row_pairs + 2)); else d.rows = nullptr;
        c.take(d.csv_off, d.cap_reads + 1);
        c.take(d.csv, d.cap_csv);
        c.take(d.tile_a, d.cap_tiles + 4); c.take(d.tile_b, d.cap_tiles + 4);
        c.take(d.info, 1);
        c.take(s.scratch.d_counters, (size_t)N_COUNTERS);
        c.take(s.scratch.d_dense_list, (size_t)s.scratch.dense_cap);
        c.take(s.scratch.d_dense_hist, hist_words);          // the slot's own dense-fallback histogram: no cross-stream chain
    };
    auto carve_host = [&](Carver& c) {
        c.take(s.h_info, 1);
        c.take(s.h_counters, (size_t)N_COUNTERS);
        c.take(s.h_csv, d.cap_csv);
        if (with_stage) c.take(s.h_text, C); else s.h_text = nullptr;
    };
    Carver size_d{nullptr}, size_h{nullptr};
    carve_dev(size_d); carve_host(size_h);
    CK(cudaMalloc(&s.d_arena, size_d.off + 256));
    CK(cudaMallocHost(&s.h_arena, size_h.off + 256));
    Carver cd{static_cast<uint8_t*>(s.d_arena)}, ch{static_cast<uint8_t*>(s.h_arena)};
    carve_dev(cd); carve_host(ch);
    s.with_stage = with_stage;
    CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CK(cudaMemsetAsync(d.text, '\n', C + 256, s.stream));
    CK(cudaMemsetAsync(s.scratch.d_counters, 0, N_COUNTERS * sizeof(uint32_t), s.stream));
    CK(cudaMemsetAsync(s.scratch.d_dense_hist, 0, hist_words * 4, s.stream));
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
    uint64_t out_off = 0;            // next CSV byte offset (owned by the turn holder)
    char* out_buf = nullptr;         // optional destination buffer
    size_t out_cap = 0;
    bool out_pinned = false;
    bool failed = false;
    int rc = CUCLARK_OK;
    std::string err;
    // totals (guarded by the turn)
    uint64_t n_reads = 0, n_cont = 0, lookups = 0, dense = 0, trunc = 0, n_chunks = 0;
    // CUCLARK_TIMING=1: seconds per phase, summed over the slot threads (guarded by mu)
    bool timing = false;
    double t_phase[7] = {0, 0, 0, 0, 0, 0, 0};   // stage-in, h2d+index, pack, classify+csv, d2h, sink, slot allocation

    // table-partitioned (routed) runs: the ranks work in lockstep rounds
    int n_ranks = 0;
    int bar_count = 0;
    uint64_t bar_gen = 0;
    int round_chunks = 0;            // ranks that got a chunk in the current round

    void fail(int code, const std::string& msg) {
        std::lock_guard<std::mutex> g(mu);
        if (!failed) { failed = true; rc = code; err = msg; }
        cv.notify_all();
    }
    // all n_ranks threads meet here; false if the job failed meanwhile (nobody is left waiting)
    bool barrier() {
        std::unique_lock<std::mutex> lk(mu);
        if (failed) return false;
        const uint64_t gen = bar_gen;
        if (++bar_count == n_ranks) { bar_count = 0; bar_gen++; cv.notify_all(); return true; }
        cv.wait(lk, [&] { return failed || bar_gen != gen; });
        return !failed;
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

// First record start at or after `from` (or n): the end of a record that is larger than a chunk.
size_t record_boundary_after(const uint8_t* t, size_t n, size_t from, bool fastq) {
    for (size_t p = std::max<size_t>(from, 1); p < n; p++) {
        if (t[p - 1] != '\n') {
            const uint8_t* nl = (const uint8_t*)memchr(t + p, '\n', n - p);
            if (!nl) return n;
            p = (size_t)(nl - t);            // the loop increment moves to the line start
            continue;
        }
        if (!fastq) { if (t[p] == '>') return p; continue; }
        if (t[p] != '@') continue;
        const uint8_t* l1 = (const uint8_t*)memchr(t + p, '\n', n - p);
        if (!l1) return n;
        const uint8_t* l2 = (const uint8_t*)memchr(l1 + 1, '\n', n - (size_t)(l1 + 1 - t));
        if (!l2 || (size_t)(l2 + 1 - t) >= n) return n;
        if (l2[1] == '+') return p;
    }
    return n;
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

struct PhaseClock {
    Job& J;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    std::chrono::steady_clock::time_point t;
    explicit PhaseClock(Job& j) : J(j) { if (J.timing) t = std::chrono::steady_clock::now(); }
    void lap(int phase) {
        if (!J.timing) return;
        const auto n = std::chrono::steady_clock::now();
        acc[phase] += std::chrono::duration<double>(n - t).count();
        t = n;
    }
    void skip() { if (J.timing) t = std::chrono::steady_clock::now(); }
    ~PhaseClock() {
        if (!J.timing) return;
        std::lock_guard<std::mutex> g(J.mu);
        for (int i = 0; i < 7; i++) J.t_phase[i] += acc[i];
    }
};

void worker(Job& J, cuclark_db* db, TextSlot& S) {
    PhaseClock pc(J);
    const TextPipe* tp = db->text_pipe;
    if (cudaSetDevice(db->cfg.device) != cudaSuccess) { J.fail(CUCLARK_ERR_CUDA, "cudaSetDevice failed"); return; }
    if (S.stream && !J.src_pinned && !S.with_stage) free_slot(S);        // allocated for a pinned source earlier
    if (!S.stream) {
        {   // nothing left to do? then do not pay for a slot
            std::lock_guard<std::mutex> g(J.mu);
            if (J.failed || J.cursor >= J.n) return;
        }
        // one slot at a time, in thread order: the driver serialises the allocations anyway, and this way
        // the first threads are already classifying while the later ones still wait for their memory
        static std::mutex alloc_mu;
        int rc;
        {
            std::lock_guard<std::mutex> g(alloc_mu);
            rc = alloc_slot(S, tp->chunk_bytes, tp->extended, tp->row_pairs, !J.src_pinned, (size_t)db->dense_blocks * db->cfg.n_targets);
        }
        if (rc) { free_slot(S); J.fail(rc, cuclark_last_error()); return; }
        pc.lap(6);
    }
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
            end = record_boundary(J.text, J.n, start, std::min(J.n, start + S.d.cap_bytes), J.fastq);
            if (end <= start) {
                // one record larger than the slot (the reference takes records of any size,
                // src/CuCLARK_hh.hh:1377-1389): the chunk is that record alone and the slot grows to hold it
                end = record_boundary_after(J.text, J.n, start + S.d.cap_bytes, J.fastq);
                if (end - start > ((size_t)1 << 31)) {
                    J.failed = true; J.rc = CUCLARK_ERR_ARG;
                    J.err = "a single record of more than 2 GiB";
                    J.cv.notify_all();
                    return;
                }
            }
            J.cursor = end;
            seq = J.next_seq++;
        }
        const uint32_t nb = (uint32_t)(end - start);
        if (nb > S.d.cap_bytes) {
            static std::mutex grow_mu;
            std::lock_guard<std::mutex> g(grow_mu);
            const bool stage = S.with_stage;
            free_slot(S);
            const int rc = alloc_slot(S, ((size_t)nb + (1u << 20)) & ~(size_t)255, tp->extended, tp->row_pairs, stage, (size_t)db->dense_blocks * db->cfg.n_targets);
            if (rc) { free_slot(S); J.fail(rc, cuclark_last_error()); return; }
        }
        const uint8_t* src = J.text + start;
        pc.skip();
        if (!J.src_pinned) { memcpy(S.h_text, src, nb); src = S.h_text; }
        pc.lap(0);
        JCK(cudaMemcpyAsync(d.text, src, nb, cudaMemcpyHostToDevice, S.stream));
        JRC(tp_index_launch(d, nb, J.fastq, S.stream));
        JCK(cudaMemcpyAsync(S.h_info, d.info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, S.stream));
        JCK(cudaStreamSynchronize(S.stream));
        if (S.h_info->err) { J.fail(CUCLARK_ERR_NOMEM, tp_err_text(S.h_info->err)); return; }
        const uint32_t n_reads = S.h_info->n_reads;
        pc.lap(1);
        JRC(tp_pack_launch(d, nb, n_reads, k, S.stream));
        JCK(cudaMemcpyAsync(S.h_info, d.info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, S.stream));
        JCK(cudaStreamSynchronize(S.stream));
        if (S.h_info->err) { J.fail(CUCLARK_ERR_NOMEM, tp_err_text(S.h_info->err)); return; }
        const uint64_t n_cont = S.h_info->n_cont;
        pc.lap(2);
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
                                2 * (size_t)tp->names.max_len;
        const uint32_t group = (uint32_t)std::min<size_t>(std::max<size_t>(d.cap_csv / max_line, 1), 0x7FFFFFFF);
        auto account = [&] {                             // called while holding the turn
            J.n_reads += n_reads; J.n_cont += n_cont; J.n_chunks++;
            J.lookups += (uint64_t)S.h_counters[COUNTER_LOOKUPS] | ((uint64_t)S.h_counters[COUNTER_LOOKUPS + 1] << 32);
            J.dense += S.h_counters[COUNTER_DENSE]; J.trunc += S.h_counters[COUNTER_TRUNC];
        };
        bool have_turn = false;
        for (uint32_t first = 0; first < n_reads; first += group) {
            const uint32_t cnt = std::min(group, n_reads - first);
            JRC(tp_csv_launch(d, tp->names, first, cnt, k, J.paired, J.extended, db->row_pairs, (uint32_t)db->cfg.n_targets, S.stream));
            JCK(cudaMemcpyAsync(S.h_info, d.info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, S.stream));
            JCK(cudaStreamSynchronize(S.stream));
            if (S.h_info->err) { J.fail(CUCLARK_ERR_NOMEM, tp_err_text(S.h_info->err)); return; }
            if (S.h_counters[COUNTER_DENSE] > S.scratch.dense_cap) { J.fail(CUCLARK_ERR_NOMEM, "too many reads needed the dense fallback"); return; }
            const size_t bytes = S.h_info->csv_bytes;
            pc.lap(3);
            // the turn only hands out the output offset (file order); the copies of different chunks overlap
            if (!have_turn) { if (!wait_turn(J, seq)) return; have_turn = true; }
            const uint64_t offset = J.out_off;
            J.out_off += bytes;
            if (first + cnt >= n_reads) { account(); end_turn(J); }
            pc.skip();
            if (!bytes) continue;
            char* dst = S.h_csv;
            if (J.out_buf) {
                if (offset + bytes > J.out_cap) { J.fail(CUCLARK_ERR_NOMEM, "the output buffer is too small for the CSV"); return; }
                if (J.out_pinned) dst = J.out_buf + offset;
            }
            JCK(cudaMemcpyAsync(dst, d.csv, bytes, cudaMemcpyDeviceToHost, S.stream));
            JCK(cudaStreamSynchronize(S.stream));
            pc.lap(4);
            if (J.out_buf && !J.out_pinned) memcpy(J.out_buf + offset, S.h_csv, bytes);
            if (J.sink && J.sink(J.user, dst, bytes, offset) != 0) { J.fail(CUCLARK_ERR_IO, "the CSV sink reported an error"); return; }
            pc.lap(5);
        }
        if (n_reads == 0) {
            if (!wait_turn(J, seq)) return;
            account();
            end_turn(J);
        }
    }
}

// Table-partitioned run (the reference's `-d N`, src/CuClarkDB.cu:546-574, 886-974): rank g = handle g holds shard g
// and one slot. The ranks work in lockstep ROUNDS: each takes the next chunk of the input (in rank order, so the
// CSV stays in file order), indexes and packs it, scatters its k-mers (route.cu); when all have, every shard probes
// what is addressed to it; when all have, each rank gathers its labels, formats its CSV and hands it to the sink.
void worker_routed(Job& J, cuclark_db* const* dbs, int g) {
    cuclark_db* db = dbs[g];
    const TextPipe* tp = db->text_pipe;
    TextSlot& S = db->text_pipe->slots[0];
    if (cudaSetDevice(db->cfg.device) != cudaSuccess) { J.fail(CUCLARK_ERR_CUDA, "cudaSetDevice failed"); return; }
    if (S.stream && !J.src_pinned && !S.with_stage) free_slot(S);
    if (!S.stream) {
        static std::mutex alloc_mu;
        std::lock_guard<std::mutex> lk(alloc_mu);
        const int rc = alloc_slot(S, tp->chunk_bytes, tp->extended, tp->row_pairs, !J.src_pinned, (size_t)db->dense_blocks * db->cfg.n_targets);
        if (rc) { free_slot(S); J.fail(rc, cuclark_last_error()); return; }
    }
    const TextSlotDev& d = S.d;
    const int k = db->cfg.k;
    for (uint64_t round = 0;; round++) {
        // ---- this rank's chunk of the round: cut in rank order
        size_t start = 0, end = 0;
        const uint64_t seq = round * (uint64_t)J.n_ranks + (uint64_t)g;
        {
            std::unique_lock<std::mutex> lk(J.mu);
            J.cv.wait(lk, [&] { return J.failed || J.next_seq == seq; });
            if (J.failed) return;
            if (g == 0) J.round_chunks = 0;
            if (J.cursor < J.n) {
                start = J.cursor;
                end = record_boundary(J.text, J.n, start, std::min(J.n, start + d.cap_bytes), J.fastq);
                if (end <= start) {
                    J.failed = true; J.rc = CUCLARK_ERR_ARG;
                    J.err = "a single record is larger than chunk_bytes (" + std::to_string(d.cap_bytes) + "); raise chunk_bytes";
                    J.cv.notify_all();
                    return;
                }
                J.cursor = end;
                J.round_chunks++;
            }
            J.next_seq++;
            J.cv.notify_all();
        }
        const uint32_t nb = (uint32_t)(end - start);
        uint32_t n_reads = 0;
        uint64_t n_cont = 0;
        if (nb) {
            const uint8_t* src = J.text + start;
            if (!J.src_pinned) { memcpy(S.h_text, src, nb); src = S.h_text; }
            JCK(cudaMemcpyAsync(d.text, src, nb, cudaMemcpyHostToDevice, S.stream));
            JRC(tp_index_launch(d, nb, J.fastq, S.stream));
            JCK(cudaMemcpyAsync(S.h_info, d.info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, S.stream));
            JCK(cudaStreamSynchronize(S.stream));
            if (S.h_info->err) { J.fail(CUCLARK_ERR_NOMEM, tp_err_text(S.h_info->err)); return; }
            n_reads = S.h_info->n_reads;
            JRC(tp_pack_launch(d, nb, n_reads, k, S.stream));
            JCK(cudaMemcpyAsync(S.h_info, d.info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, S.stream));
            JCK(cudaStreamSynchronize(S.stream));
            if (S.h_info->err) { J.fail(CUCLARK_ERR_NOMEM, tp_err_text(S.h_info->err)); return; }
            n_cont = S.h_info->n_cont;
        }
        JRC(route_scatter(db, d.reads_ptr, d.cont, n_reads, n_cont, S.stream));
        JCK(cudaStreamSynchronize(S.stream));
        if (!J.barrier()) return;                            // every rank's k-mers are in its arena
        bool any;
        { std::lock_guard<std::mutex> lk(J.mu); any = J.round_chunks > 0; }
        if (!any) return;                                    // the input is used up (all ranks see the same count)
        JRC(route_probe(db, S.stream));
        JCK(cudaStreamSynchronize(S.stream));
        if (!J.barrier()) return;                            // every label is back with the rank that asked
        JRC(route_gather(db, S.scratch, d.reads_ptr, d.cont, n_reads, n_cont, d.final5, J.extended ? d.rows : nullptr, S.stream));
        JCK(cudaMemcpyAsync(S.h_counters, S.scratch.d_counters, N_COUNTERS * sizeof(uint32_t), cudaMemcpyDeviceToHost, S.stream));
        JCK(cudaStreamSynchronize(S.stream));
        if (S.h_counters[COUNTER_DENSE] > S.scratch.dense_cap) { J.fail(CUCLARK_ERR_NOMEM, "too many reads needed the dense fallback"); return; }
        cuclark_route_stats rs;
        JRC(route_stats(db, &rs));
        if (rs.err) { J.fail(CUCLARK_ERR_NOMEM, "routing arena exhausted"); return; }
        // ---- CSV of this rank's chunk, handed over in file order (chunk `seq`)
        const size_t max_line = 39 + (J.extended ? 2 * (size_t)db->cfg.n_targets + 4 * (size_t)db->row_pairs : 0) + 64 +
                                2 * (size_t)tp->names.max_len;
        const uint32_t group = (uint32_t)std::min<size_t>(std::max<size_t>(d.cap_csv / max_line, 1), 0x7FFFFFFF);
        if (!wait_turn(J, seq)) return;
        for (uint32_t first = 0; first < n_reads; first += group) {
            const uint32_t cnt = std::min(group, n_reads - first);
            JRC(tp_csv_launch(d, tp->names, first, cnt, k, J.paired, J.extended, db->row_pairs, (uint32_t)db->cfg.n_targets, S.stream));
            JCK(cudaMemcpyAsync(S.h_info, d.info, sizeof(ChunkInfo), cudaMemcpyDeviceToHost, S.stream));
            JCK(cudaStreamSynchronize(S.stream));
            if (S.h_info->err) { J.fail(CUCLARK_ERR_NOMEM, tp_err_text(S.h_info->err)); return; }
            const size_t bytes = S.h_info->csv_bytes;
            const uint64_t offset = J.out_off;
            J.out_off += bytes;
            if (!bytes) continue;
            char* dst = S.h_csv;
            if (J.out_buf) {
                if (offset + bytes > J.out_cap) { J.fail(CUCLARK_ERR_NOMEM, "the output buffer is too small for the CSV"); return; }
                if (J.out_pinned) dst = J.out_buf + offset;
            }
            JCK(cudaMemcpyAsync(dst, d.csv, bytes, cudaMemcpyDeviceToHost, S.stream));
            JCK(cudaStreamSynchronize(S.stream));
            if (J.out_buf && !J.out_pinned) memcpy(J.out_buf + offset, S.h_csv, bytes);
            if (J.sink && J.sink(J.user, dst, bytes, offset) != 0) { J.fail(CUCLARK_ERR_IO, "the CSV sink reported an error"); return; }
        }
        J.n_reads += n_reads; J.n_cont += n_cont; J.n_chunks += nb ? 1 : 0;
        J.lookups += rs.lookups;
        J.dense += S.h_counters[COUNTER_DENSE]; J.trunc += S.h_counters[COUNTER_TRUNC];
        end_turn(J);
    }
}

int run_text(cuclark_db* const* dbs, int n_dbs, const uint8_t* text, size_t n, const cuclark_text_opts* o,
             cuclark_sink_fn sink, void* user, char* out_buf, size_t out_cap, cuclark_text_arrays* arrays,
             cuclark_text_stats* out) {
    if (!dbs || n_dbs < 1 || (!text && n)) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    for (int i = 0; i < n_dbs; i++) {
        if (!dbs[i]) { set_error("null argument"); return CUCLARK_ERR_ARG; }
        if (!dbs[i]->d_table) { set_error("no database loaded"); return CUCLARK_ERR_STATE; }
        if (dbs[i]->cfg.shard_count > 1 && (dbs[i]->cfg.shard_count != n_dbs || dbs[i]->cfg.shard_index != i)) {
            set_error("table-partitioned run: pass the %d shard handles in shard order (handle %d is shard %d of %d)", dbs[i]->cfg.shard_count, i,
                      dbs[i]->cfg.shard_index, dbs[i]->cfg.shard_count);
            return CUCLARK_ERR_STATE;
        }
        if (dbs[i]->cfg.k != dbs[0]->cfg.k || dbs[i]->cfg.n_targets != dbs[0]->cfg.n_targets || dbs[i]->row_pairs != dbs[0]->row_pairs) {
            set_error("handles of a multi-device run must share k, n_targets and row_pairs");
            return CUCLARK_ERR_ARG;
        }
    }
    if (out) memset(out, 0, sizeof *out);
    const auto t0 = std::chrono::steady_clock::now();
    if (n == 0 || (text[0] != '>' && text[0] != '@')) {
        set_error("Failed to recognize the format of the file.");            // src/CuCLARK_hh.hh:1535-1538
        return CUCLARK_ERR_FORMAT;
    }
    // default chunk: small enough that a one-shot run (the CLI) spends little on pinned slot memory and
    // has many chunks to overlap, large enough to keep the per-chunk synchronisations negligible
    size_t chunk = o && o->chunk_bytes ? o->chunk_bytes : (n <= ((size_t)4 << 30) ? (size_t)4 << 20 : (size_t)16 << 20);
    chunk = std::min<size_t>(std::max<size_t>(chunk, 4096), (size_t)1 << 30);
    chunk = (chunk + 255) & ~(size_t)255;
    int n_slots = o && o->n_slots > 0 ? o->n_slots : 4;
    n_slots = std::min(n_slots, 16);
    const bool routed = dbs[0]->cfg.shard_count > 1;         // table-partitioned: one slot per rank, lockstep rounds
    if (routed) {
        if (arrays) { set_error("cuclark_text_debug needs the whole table on one device"); return CUCLARK_ERR_STATE; }
        n_slots = 1;
        if (!(o && o->chunk_bytes)) chunk = std::min<size_t>((size_t)32 << 20, std::max<size_t>((n / n_dbs + 4095) & ~(size_t)4095, (size_t)1 << 20));
    }
    const bool extended = o && o->extended;
    for (int i = 0; i < n_dbs; i++) {
        CK(cudaSetDevice(dbs[i]->cfg.device));
        int rc = ensure_pipe(dbs[i], chunk, n_slots, extended || (arrays && arrays->rows), o ? o->target_names : nullptr);
        if (rc) return rc;
    }
    if (routed) {
        // routing buffers for one chunk per rank (every k-mer starts in a data container: <= chunk/2 + slack containers)
        const size_t cap_cont = chunk / 2 + 1024 + 8;
        bool fresh = false;
        for (int i = 0; i < n_dbs; i++) {
            CK(cudaSetDevice(dbs[i]->cfg.device));
            cuclark_route_stats rs;
            if (!dbs[i]->route || route_stats(dbs[i], &rs) != CUCLARK_OK || rs.map_bytes != 8 * cap_cont * 4) {
                int rc = route_alloc(dbs[i], n_dbs, cap_cont);
                if (rc) return rc;
                fresh = true;
            }
        }
        if (fresh) { int rc = route_connect(dbs, n_dbs); if (rc) return rc; }
    }
    const auto t_setup = std::chrono::steady_clock::now();
    Job J;
    J.timing = getenv("CUCLARK_TIMING") != nullptr;
    J.text = text; J.n = n;
    J.fastq = text[0] == '@';
    J.paired = o && o->paired; J.extended = extended;
    J.sink = sink; J.user = user; J.arrays = arrays;
    J.out_buf = out_buf; J.out_cap = out_cap;
    cudaPointerAttributes attr;
    J.src_pinned = cudaPointerGetAttributes(&attr, text) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    J.out_pinned = out_buf && cudaPointerGetAttributes(&attr, out_buf) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if ((sink || out_buf) && !arrays) {
        // header line (src/CuCLARK_hh.hh:1957-1972)
        const TextPipe* tp = dbs[0]->text_pipe;
        std::string h = "Object_ID";
        if (extended) for (size_t t = 1; t < tp->host_names.size(); t++) { h += ","; h += tp->host_names[t]; }
        h += ",Length,Gamma,1st_assignment,score1,2nd_assignment,score2,confidence\n";
        if (out_buf) {
            if (h.size() > out_cap) { set_error("the output buffer is too small for the CSV"); return CUCLARK_ERR_NOMEM; }
            memcpy(out_buf, h.data(), h.size());
        }
        if (sink && sink(user, h.data(), h.size(), 0) != 0) { set_error("the CSV sink reported an error"); return CUCLARK_ERR_IO; }
        J.out_off = h.size();
    }
    // one host thread per slot; slots of all devices pull chunks from the same cursor. A slot costs
    // ~10 ms of (serialised) driver time to allocate: short inputs get fewer of them
    const size_t n_chunks_est = (n + chunk - 1) / chunk;
    const size_t max_threads = std::max<size_t>(std::min<size_t>(n_chunks_est, 2), n_chunks_est / 8);
    std::vector<std::thread> threads;
    size_t started = 0;
    if (routed) {
        J.n_ranks = n_dbs;
        for (int i = 0; i < n_dbs; i++) threads.emplace_back([&J, dbs, i] { worker_routed(J, dbs, i); });
    } else
    for (int s = 0; s < n_slots && started < max_threads; s++)
        for (int i = 0; i < n_dbs && started < max_threads; i++, started++) {
            cuclark_db* db = dbs[i];
            TextSlot* slot = &db->text_pipe->slots[s];
            threads.emplace_back([&J, db, slot] { worker(J, db, *slot); });
        }
    for (auto& t : threads) t.join();
    if (J.timing) {
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[cuclark timing] setup %.1f ms, run %.1f ms with %zu threads (%zu chunks of <= %zu MiB); thread-seconds: "
                        "slot alloc %.3f, stage-in %.3f, h2d+index %.3f, pack %.3f, classify+csv %.3f, d2h %.3f, sink %.3f\n",
                std::chrono::duration<double, std::milli>(t_setup - t0).count(),
                std::chrono::duration<double, std::milli>(t1 - t_setup).count(), threads.size(), (size_t)J.n_chunks, chunk >> 20,
                J.t_phase[6], J.t_phase[0], J.t_phase[1], J.t_phase[2], J.t_phase[3], J.t_phase[4], J.t_phase[5]);
    }
    if (J.failed) { set_error("%s", J.err.c_str()); return J.rc; }
    dbs[0]->last_lookups = J.lookups; dbs[0]->last_dense = J.dense; dbs[0]->last_trunc = J.trunc;
    if (out) {
        out->n_reads = J.n_reads; out->lookups = J.lookups; out->csv_bytes = J.out_off; out->n_chunks = J.n_chunks;
        out->n_containers = J.n_cont; out->dense_reads = J.dense; out->truncated_rows = J.trunc;
        out->text_bytes = n;
        out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return CUCLARK_OK;
}

// called concurrently from the slot threads, each with its own byte range of the file
int file_sink(void* user, const char* data, size_t n, uint64_t offset) {
    const int fd = (int)(intptr_t)user;
    while (n) {
        const ssize_t w = pwrite(fd, data, n, (off_t)offset);
        if (w <= 0) return -1;
        data += w; n -= (size_t)w; offset += (uint64_t)w;
    }
    return 0;
}

}  // namespace
}  // namespace cuclark

using namespace cuclark;

extern "C" {

int cuclark_classify_text(cuclark_db* db, const uint8_t* text, size_t n, const cuclark_text_opts* opts,
                          cuclark_sink_fn sink, void* user, cuclark_text_stats* out) {
    return run_text(&db, 1, text, n, opts, sink, user, nullptr, 0, nullptr, out);
}

int cuclark_classify_text_buffer(cuclark_db* const* dbs, int n_dbs, const uint8_t* text, size_t n,
                                 const cuclark_text_opts* opts, char* out, size_t out_cap, size_t* out_len,
                                 cuclark_text_stats* stats) {
    if (!out || !out_len) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    cuclark_text_stats st;
    const int rc = run_text(dbs, n_dbs, text, n, opts, nullptr, nullptr, out, out_cap, nullptr, &st);
    *out_len = rc == CUCLARK_OK ? (size_t)st.csv_bytes : 0;
    if (stats) *stats = st;
    return rc;
}

int cuclark_classify_text_multi(cuclark_db* const* dbs, int n_dbs, const uint8_t* text, size_t n,
                                const cuclark_text_opts* opts, cuclark_sink_fn sink, void* user,
                                cuclark_text_stats* out) {
    return run_text(dbs, n_dbs, text, n, opts, sink, user, nullptr, 0, nullptr, out);
}

int cuclark_text_debug(cuclark_db* db, const uint8_t* text, size_t n, const cuclark_text_opts* opts,
                       cuclark_text_arrays* arrays, cuclark_text_stats* out) {
    if (!arrays) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    return run_text(&db, 1, text, n, opts, nullptr, nullptr, nullptr, 0, arrays, out);
}

int cuclark_classify_file_multi(cuclark_db* const* dbs, int n_dbs, const char* objects_path, const char* csv_path,
                                const cuclark_text_opts* opts, cuclark_text_stats* out) {
    if (!dbs || n_dbs < 1 || !objects_path || !csv_path) { set_error("null argument"); return CUCLARK_ERR_ARG; }
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
    const int ofd = open(csv_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (ofd < 0) { munmap(map, n); close(fd); set_error("Failed to create/open file result: %s", csv_path); return CUCLARK_ERR_IO; }
    int rc = run_text(dbs, n_dbs, (const uint8_t*)map, n, opts, file_sink, (void*)(intptr_t)ofd, nullptr, 0, nullptr, out);
    if (close(ofd) != 0 && rc == CUCLARK_OK) { set_error("failed to write %s", csv_path); rc = CUCLARK_ERR_IO; }
    munmap(map, n);
    close(fd);
    return rc;
}

int cuclark_classify_file(cuclark_db* db, const char* objects_path, const char* csv_path, const cuclark_text_opts* opts,
                          cuclark_text_stats* out) {
    return cuclark_classify_file_multi(&db, 1, objects_path, csv_path, opts, out);
}

}  // extern "C"
