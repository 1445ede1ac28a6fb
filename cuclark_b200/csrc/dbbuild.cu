// cuclark_b200 — database construction on the device.
//
// Replaces the host-only, single-threaded builder of the reference
//   makeSpecificTargetSets   src/CuCLARK_hh.hh:691-1112   (scan targets, add every k-mer)
//   EHashtable::addElement   src/HashTableStorage_hh.hh:484-523  (canonical, multiplicity)
//   RemoveCommon             src/HashTableStorage_hh.hh:242-292  (keep k-mers of one label)
//   hTable::write            src/hashTable_hh.hh:591-663         (.sz/.ky/.lb)
// and writes byte-identical files. The reference inserts k-mer by k-mer into HTSIZE
// std::vector-like buckets on the host (146 GB RAM and hours at bacterial scale,
// README.md:93). Here the reference's own bucket function r = c mod HTSIZE is the key of a
// two-pass counting sort in HBM:
//   pass A  every k-mer occurrence of every target: count[r]++
//           (per file: blank header lines, compact the nucleotides, roll the k-mers)
//   scan    count -> segment offsets (tile bases in 64 bit, 32-bit offsets inside a tile)
//   pass B  the same k-mers again: (q = c div HTSIZE, label) scattered into bucket r's segment
//   reduce  one thread per bucket: sort its segment by q, keep a k-mer iff all its copies
//           carry the same label and it was seen more than minCount times; the number of
//           kept entries IS the .sz byte
//   write   scan of the kept counts -> .ky / .lb in bucket order, ascending q in a bucket
// All stages are HBM-bound integer work (random 4-byte atomics in passes A/B, streaming
// elsewhere); nothing is kept per k-mer on the host.
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "internal.h"

namespace cuclark {

namespace {

constexpr int DBT = 256;
constexpr int DB_BYTES = 16;
constexpr int DB_TILE = DBT * DB_BYTES;
constexpr int SEG_TILE = 2048;            // buckets per offset tile
constexpr uint8_t BLANK = 1;              // byte that replaces header text: class "other"

constexpr uint32_t DBERR_BUCKET_255 = 1;  // a bucket keeps >= 256 entries: cannot be stored (src/hashTable_hh.hh:620-625)

__device__ __forceinline__ uint32_t blk_excl_scan(uint32_t v, uint32_t& total) {
    __shared__ uint32_t ws[DBT / 32];
    __shared__ uint32_t wtotal;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < DBT / 32 ? ws[lane] : 0, xi = x;
#pragma unroll
        for (int o = 1; o < DBT / 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, xi, o);
            if (lane >= o) xi += t;
        }
        if (lane < DBT / 32) ws[lane] = xi - x;
        if (lane == DBT / 32 - 1) wtotal = xi;
    }
    __syncthreads();
    const uint32_t res = inc - v + ws[w];
    total = wtotal;
    __syncthreads();
    return res;
}

// single block: exclusive scan of 64-bit tile sums in place; *total = sum
__global__ void __launch_bounds__(DBT) k_scan_u64(uint64_t* tiles, uint64_t n, uint64_t* total) {
    __shared__ uint64_t sh[DBT];
    uint64_t carry = 0;
    for (uint64_t b = 0; b < n; b += DBT) {
        const uint64_t i = b + threadIdx.x;
        const uint64_t v = i < n ? tiles[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < DBT; o <<= 1) {
            const uint64_t t = threadIdx.x >= (unsigned)o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) tiles[i] = carry + sh[threadIdx.x] - v;
        carry += sh[DBT - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

// ---- per-file text stages -----------------------------------------------------------------
// (1) line starts: a '>' ANYWHERE in a line makes the scanner skip the rest of that line and
// restart the k-mer window (m_table['>'] = -2, src/CuCLARK_hh.hh:746-760). One warp per line
// finds the first '>' and overwrites [it, end of line) with BLANK.
__global__ void k_db_mark_lines(const uint8_t* __restrict__ text, uint32_t n, uint32_t* line_flag) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    line_flag[p] = (p == 0 || text[p - 1] == '\n') ? 1u : 0u;
}

// From every '>' to the end of its line the text is blanked: the reference skips from a '>' to the next newline
// (m_table['>'] == -2, src/CuCLARK_hh.hh:956-974). One thread per byte finds the '>' (they are rare: one per record),
// and walks its header line. (Round 1 gave every LINE a warp that searched it for a '>': a single-line 4 Mbp genome
// was scanned by one warp, 35 ms per file and pass — 83 of the 106 s of the bacterial-scale build.)
__global__ void __launch_bounds__(DBT) k_db_blank_headers(uint8_t* text, uint32_t n) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || text[p] != '>') return;
    for (uint32_t q = p; q < n && text[q] != '\n'; q++) text[q] = BLANK;      // (a second '>' of the line: its thread repeats the tail)
}

// class of a byte after header blanking: 0..3 forward code (m_table: A0 C1 G2 T/U3), 4 newline, 5 break
__device__ __forceinline__ int db_class(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        case '\n': return 4;
        default: return 5;
    }
}

// (2) compaction: codes[j] = forward code of the j-th nucleotide, brk[j] = 1 if a break
// (non-nucleotide other than '\n') lies between nucleotide j-1 and j.
__global__ void __launch_bounds__(DBT) k_db_count_nt(const uint8_t* __restrict__ text, uint32_t n, uint32_t* tiles,
                                                     int flag_mode, const uint32_t* __restrict__ flags) {
    const uint32_t base = blockIdx.x * DB_TILE + threadIdx.x * DB_BYTES;
    uint32_t c = 0;
    for (int i = 0; i < DB_BYTES; i++)
        if (base + i < n) c += flag_mode ? flags[base + i] : (uint32_t)(db_class(text[base + i]) < 4);
    uint32_t total;
    blk_excl_scan(c, total);
    if (threadIdx.x == 0) tiles[blockIdx.x] = total;
}

__global__ void __launch_bounds__(DBT) k_scan_u32_tiles(uint32_t* tiles, uint32_t n_tiles, uint32_t* total_out) {
    uint32_t carry = 0;
    for (uint32_t b = 0; b < n_tiles; b += DBT) {
        const uint32_t i = b + threadIdx.x;
        const uint32_t v = i < n_tiles ? tiles[i] : 0u;
        uint32_t total;
        const uint32_t ex = blk_excl_scan(v, total);
        if (i < n_tiles) tiles[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(DBT) k_db_compact(const uint8_t* __restrict__ text, uint32_t n,
                                                    const uint32_t* __restrict__ tiles, uint8_t* codes, uint8_t* brk) {
    const uint32_t base = blockIdx.x * DB_TILE + threadIdx.x * DB_BYTES;
    int cls[DB_BYTES];
    uint32_t c = 0;
    for (int i = 0; i < DB_BYTES; i++) {
        cls[i] = base + i < n ? db_class(text[base + i]) : 4;
        c += cls[i] < 4;
    }
    uint32_t total;
    uint32_t j = blk_excl_scan(c, total) + tiles[blockIdx.x];
    for (int i = 0; i < DB_BYTES; i++) {
        if (cls[i] < 4) codes[j++] = (uint8_t)cls[i];
        else if (cls[i] == 5) brk[j] = 1;          // j = index of the next nucleotide (array has n_nt + 1 entries)
    }
}

// line table from flags: line_start[rank] = p for every p with flag
__global__ void __launch_bounds__(DBT) k_db_emit_flagged(const uint32_t* __restrict__ flags, uint32_t n,
                                                         const uint32_t* __restrict__ tiles, uint32_t* out) {
    const uint32_t base = blockIdx.x * DB_TILE + threadIdx.x * DB_BYTES;
    uint32_t c = 0;
    for (int i = 0; i < DB_BYTES; i++) if (base + i < n) c += flags[base + i];
    uint32_t total;
    uint32_t j = blk_excl_scan(c, total) + tiles[blockIdx.x];
    for (int i = 0; i < DB_BYTES; i++) if (base + i < n && flags[base + i]) out[j++] = base + i;
}

// ---- k-mer emission -------------------------------------------------------------------------
struct Emit {
    uint32_t* count;           // HTSIZE counters (pass A: ++; pass B: cursor, counts down)
    const uint32_t* local;     // pass B: exclusive offset of bucket r inside its tile
    const uint64_t* tile_base; // pass B: 64-bit base of each tile of SEG_TILE buckets
    void* keys;                // pass B: quotients, key_bytes wide (4 or 8 here; 2-byte files use 4 internally)
    uint16_t* labels;
    uint64_t htsize, magic;
    int wide;                  // keys are uint64
    int pass_b;
    uint16_t label;
};

__device__ __forceinline__ void emit_kmer(const Emit& e, uint64_t R, int k) {
    const uint64_t c = canonical(R, k);
    uint64_t q, r;
    divmod_M(c, e.htsize, e.magic, q, r);
    if (!e.pass_b) { atomicAdd(&e.count[r], 1u); return; }
    const uint32_t left = atomicSub(&e.count[r], 1u);                  // left-1 = my slot inside the segment
    const uint64_t pos = e.tile_base[r / SEG_TILE] + e.local[r] + (left - 1);
    if (e.wide) static_cast<uint64_t*>(e.keys)[pos] = q;
    else static_cast<uint32_t*>(e.keys)[pos] = (uint32_t)q;
    e.labels[pos] = e.label;
}

// full variant: every window of k nucleotides inside a run (src/CuCLARK_hh.hh:914-975). The integer is
// the R form (first base in the high bits, complement code), as the scanner's _km_r.
__global__ void __launch_bounds__(DBT) k_db_kmers_full(const uint8_t* __restrict__ codes, const uint8_t* __restrict__ brk,
                                                       uint32_t n_nt, int k, Emit e) {
    constexpr uint32_t RUN = 64;
    const uint64_t j0 = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * RUN;
    if (j0 >= n_nt) return;
    const uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t R = 0;
    uint32_t len = 0;
    const uint64_t end = min((uint64_t)n_nt, j0 + RUN + k - 1);
    for (uint64_t p = j0; p < end; p++) {
        len = (p > j0 && brk[p]) ? 1 : len + 1;
        R = ((R << 2) | (3u - codes[p])) & mask;
        if (len >= (uint32_t)k && p - (k - 1) < j0 + RUN) emit_kmer(e, R, k);
    }
}

// light variant: consecutive NON-overlapping k-mers of each run; the i-th completed k-mer of the FILE
// is added iff i % gap == 0 (src/CuCLARK_hh.hh:705-767).
// a run starts at nucleotide 0 and after every break
__global__ void k_db_run_flags(const uint8_t* __restrict__ brk, uint32_t n_nt, uint32_t* flags) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_nt) flags[j] = (j == 0 || brk[j]) ? 1u : 0u;
}

__global__ void k_db_run_windows(const uint32_t* __restrict__ run_start, uint32_t n_runs, uint32_t n_nt, int k,
                                 uint32_t* windows) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint32_t s = run_start[r], e = r + 1 < n_runs ? run_start[r + 1] : n_nt;
    windows[r] = (e - s) / (uint32_t)k;
}

__global__ void k_db_kmers_light(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ run_start,
                                 const uint32_t* __restrict__ win_base, uint32_t n_runs, uint32_t n_windows, int k,
                                 int gap, Emit e) {
    const uint64_t i = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * (uint64_t)gap;     // window index in the file
    if (i >= n_windows) return;
    // run containing window i: last r with win_base[r] <= i
    uint32_t lo = 0, hi = n_runs - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (win_base[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const uint64_t s = (uint64_t)run_start[lo] + (i - win_base[lo]) * (uint64_t)k;
    uint64_t R = 0;
    for (int j = 0; j < k; j++) R = (R << 2) | (3u - codes[s + j]);
    emit_kmer(e, R, k);
}

// ---- bucket offsets ---------------------------------------------------------------------------
__global__ void __launch_bounds__(DBT) k_db_tile_offsets(const uint32_t* __restrict__ count, uint64_t htsize, uint32_t* local,
                                                         uint64_t* tile_sum) {
    // one block per tile of SEG_TILE buckets (8 per thread)
    const uint64_t base = (uint64_t)blockIdx.x * SEG_TILE + threadIdx.x * (SEG_TILE / DBT);
    uint32_t v[SEG_TILE / DBT], s = 0;
    for (int i = 0; i < SEG_TILE / DBT; i++) { v[i] = base + i < htsize ? count[base + i] : 0u; s += v[i]; }
    uint32_t total;
    uint32_t run = blk_excl_scan(s, total);
    for (int i = 0; i < SEG_TILE / DBT; i++) {
        if (base + i < htsize) local[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// ---- per-bucket reduce: sort by q, RemoveCommon, compact ------------------------------------------
template <typename K>
__device__ __forceinline__ void sift_down(K* q, uint16_t* l, uint64_t start, uint64_t end) {
    uint64_t root = start;
    for (;;) {
        uint64_t child = 2 * root + 1;
        if (child > end) return;
        if (child + 1 <= end && (q[child] < q[child + 1] || (q[child] == q[child + 1] && l[child] < l[child + 1]))) child++;
        if (q[root] < q[child] || (q[root] == q[child] && l[root] < l[child])) {
            const K tq = q[root]; q[root] = q[child]; q[child] = tq;
            const uint16_t tl = l[root]; l[root] = l[child]; l[child] = tl;
            root = child;
        } else return;
    }
}

template <typename K>
__global__ void __launch_bounds__(DBT) k_db_reduce(K* keys, uint16_t* labels, const uint32_t* __restrict__ local,
                                                   const uint64_t* __restrict__ tile_base, uint32_t* kept, uint64_t htsize,
                                                   uint32_t min_count, uint64_t total_entries, uint32_t* err) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= htsize) return;
    const uint64_t s = tile_base[r / SEG_TILE] + local[r];
    uint64_t e;
    if (r + 1 >= htsize) e = total_entries;
    else e = tile_base[(r + 1) / SEG_TILE] + local[r + 1];
    const uint64_t n = e - s;
    if (n == 0) { kept[r] = 0; return; }
    K* q = keys + s;
    uint16_t* l = labels + s;
    if (n > 1) {
        if (n <= 24) {                                           // insertion sort by (q, label)
            for (uint64_t i = 1; i < n; i++) {
                const K kq = q[i];
                const uint16_t kl = l[i];
                uint64_t j = i;
                while (j > 0 && (q[j - 1] > kq || (q[j - 1] == kq && l[j - 1] > kl))) { q[j] = q[j - 1]; l[j] = l[j - 1]; j--; }
                q[j] = kq; l[j] = kl;
            }
        } else {                                                 // heap sort: in place, O(n log n) for pathological buckets
            for (uint64_t st = (n - 2) / 2 + 1; st-- > 0;) sift_down(q, l, st, n - 1);
            for (uint64_t end = n - 1; end > 0; end--) {
                const K tq = q[0]; q[0] = q[end]; q[end] = tq;
                const uint16_t tl = l[0]; l[0] = l[end]; l[end] = tl;
                sift_down(q, l, (uint64_t)0, end - 1);
            }
        }
    }
    // runs of equal q: keep iff one label and min(copies, 254) > min_count
    // (lElement::AddToCount saturates below 255, src/dataType.hh:335-336; RemoveCommon :257)
    uint64_t w = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i + 1;
        while (j < n && q[j] == q[i]) j++;
        const bool one_label = l[j - 1] == l[i];                 // sorted by label inside the run
        const uint64_t copies = min(j - i, (uint64_t)254);
        if (one_label && copies > min_count) { q[w] = q[i]; l[w] = l[i]; w++; }
        i = j;
    }
    kept[r] = (uint32_t)w;
    if (w >= 256) atomicOr(err, DBERR_BUCKET_255);
}

template <typename K>
__global__ void __launch_bounds__(DBT) k_db_write(const K* __restrict__ keys, const uint16_t* __restrict__ labels,
                                                  const uint32_t* __restrict__ local, const uint64_t* __restrict__ tile_base,
                                                  const uint32_t* __restrict__ kept, const uint32_t* __restrict__ out_local,
                                                  const uint64_t* __restrict__ out_tile_base, uint64_t htsize, uint64_t r0,
                                                  uint64_t nr, uint64_t out0, uint8_t* sz, void* ky, uint16_t* lb, int key_bytes) {
    // buckets [r0, r0 + nr) of one output window; out0 = output index of the window's first entry
    const uint64_t r = r0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= r0 + nr || r >= htsize) return;
    const uint32_t n = kept[r];
    sz[r - r0] = (uint8_t)n;
    if (!n) return;
    const uint64_t s = tile_base[r / SEG_TILE] + local[r];
    const uint64_t o = out_tile_base[r / SEG_TILE] + out_local[r] - out0;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t q = keys[s + i];
        if (key_bytes == 2) static_cast<uint16_t*>(ky)[o + i] = (uint16_t)q;
        else if (key_bytes == 4) static_cast<uint32_t*>(ky)[o + i] = (uint32_t)q;
        else static_cast<uint64_t*>(ky)[o + i] = q;
        lb[o + i] = labels[s + i];
    }
}

struct FileMap {
    const uint8_t* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open_file(const char* path) {
        fd = open(path, O_RDONLY);
        struct stat sb;
        if (fd < 0 || fstat(fd, &sb) != 0) { if (fd >= 0) close(fd); fd = -1; return false; }
        n = (size_t)sb.st_size;
        if (n == 0) { p = nullptr; return true; }
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); fd = -1; return false; }
        p = (const uint8_t*)m;
        return true;
    }
    ~FileMap() {
        if (p) munmap((void*)p, n);
        if (fd >= 0) close(fd);
    }
};

// A FASTQ target (src/CuCLARK_hh.hh:986-1080 full variant, :769-860 light): the first line is skipped, the sequence
// line is scanned, its newline resets the window and the next three lines ('+', qualities, the next header) are
// skipped. The same k-mers come out of the FASTA scanner for the text ">\n<sequence line>\n" per record (a '>' line
// is skipped and resets the window, :956-974), so a FASTQ target is rewritten into that form on the host.
void fastq_target_as_fasta(const uint8_t* p, size_t n, std::vector<uint8_t>& out) {
    out.clear();
    out.reserve(n / 2 + 16);
    size_t i = 0;
    auto skip_line = [&] {
        const uint8_t* nl = (const uint8_t*)memchr(p + i, '\n', n - i);
        i = nl ? (size_t)(nl - p) + 1 : n;
    };
    skip_line();                                     // first header
    while (i < n) {
        const uint8_t* nl = (const uint8_t*)memchr(p + i, '\n', n - i);
        const size_t e = nl ? (size_t)(nl - p) : n;
        out.push_back('>'); out.push_back('\n');
        out.insert(out.end(), p + i, p + e);
        out.push_back('\n');
        i = nl ? e + 1 : n;
        skip_line(); skip_line(); skip_line();       // '+', qualities, next header
    }
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return CUCLARK_OK;
        cudaFree(p);
        p = nullptr; cap = 0;
        bytes = bytes + bytes / 4 + 4096;
        CK(cudaMalloc(&p, bytes));
        cap = bytes;
        return CUCLARK_OK;
    }
    ~DevBuf() { cudaFree(p); }
    template <typename T> T* as() { return static_cast<T*>(p); }
};

// per-file scratch, grown on demand
struct FileScratch {
    DevBuf text, flags, line_start, tiles, codes, brk, run_start, windows;
    uint32_t* d_small = nullptr;      // [0] n_lines, [1] n_nt, [2] n_runs, [3] n_windows
    ~FileScratch() { cudaFree(d_small); }
};

int write_all(int fd, const void* buf, size_t n) {
    const char* p = (const char*)buf;
    while (n) {
        const ssize_t w = write(fd, p, n);
        if (w <= 0) return -1;
        p += w; n -= (size_t)w;
    }
    return 0;
}

// one pass over one target file: emits every k-mer the reference scanner adds
int scan_file(FileScratch& fs, const uint8_t* host, size_t n_bytes, int k, int light_gap, Emit e, uint64_t* n_kmers,
              uint64_t* n_nt_out, cudaStream_t st) {
    if (n_bytes >= 0xFFFF0000ull) { set_error("target files of 4 GB or more are not supported by the device builder"); return CUCLARK_ERR_ARG; }
    const uint32_t n = (uint32_t)n_bytes;
    const uint32_t n_tiles = (n + DB_TILE - 1) / DB_TILE;
    int rc;
#define R(call) do { rc = (call); if (rc) return rc; } while (0)
    R(fs.text.reserve((size_t)n + 64));
    R(fs.flags.reserve((size_t)n * 4 + 64));
    R(fs.tiles.reserve((size_t)(n_tiles + 8) * 4));
    if (!fs.d_small) CK(cudaMalloc(&fs.d_small, 64));
    uint8_t* text = fs.text.as<uint8_t>();
    uint32_t* flags = fs.flags.as<uint32_t>();
    uint32_t* tiles = fs.tiles.as<uint32_t>();
    CK(cudaMemcpyAsync(text, host, n, cudaMemcpyHostToDevice, st));
    // lines
    k_db_mark_lines<<<(n + 255) / 256, 256, 0, st>>>(text, n, flags);
    k_db_count_nt<<<n_tiles, DBT, 0, st>>>(text, n, tiles, 1, flags);
    k_scan_u32_tiles<<<1, DBT, 0, st>>>(tiles, n_tiles, fs.d_small + 0);
    uint32_t h_small[4];
    CK(cudaMemcpyAsync(h_small, fs.d_small, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t n_lines = h_small[0];
    R(fs.line_start.reserve((size_t)(n_lines + 2) * 4));
    k_db_emit_flagged<<<n_tiles, DBT, 0, st>>>(flags, n, tiles, fs.line_start.as<uint32_t>());
    k_db_blank_headers<<<(n + DBT - 1) / DBT, DBT, 0, st>>>(text, n);
    // compaction
    k_db_count_nt<<<n_tiles, DBT, 0, st>>>(text, n, tiles, 0, nullptr);
    k_scan_u32_tiles<<<1, DBT, 0, st>>>(tiles, n_tiles, fs.d_small + 1);
    CK(cudaMemcpyAsync(h_small + 1, fs.d_small + 1, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t n_nt = h_small[1];
    if (n_nt_out) *n_nt_out += n_nt;
    if (n_nt < (uint32_t)k) return CUCLARK_OK;
    R(fs.codes.reserve((size_t)n_nt + 64));
    R(fs.brk.reserve((size_t)n_nt + 64));
    uint8_t* codes = fs.codes.as<uint8_t>();
    uint8_t* brk = fs.brk.as<uint8_t>();
    CK(cudaMemsetAsync(brk, 0, (size_t)n_nt + 1, st));
    k_db_compact<<<n_tiles, DBT, 0, st>>>(text, n, tiles, codes, brk);
    if (!light_gap) {
        const uint64_t threads = ((uint64_t)n_nt + 63) / 64;
        k_db_kmers_full<<<(uint32_t)((threads + DBT - 1) / DBT), DBT, 0, st>>>(codes, brk, n_nt, k, e);
        CK(cudaGetLastError());
        if (n_kmers) *n_kmers += 0;      // occurrences are counted on the device (sum of the bucket counters)
        return CUCLARK_OK;
    }
    // light: run table (a run starts at nucleotide 0 and after every break), windows per run, their prefix sums
    const uint32_t nt_tiles = (n_nt + DB_TILE - 1) / DB_TILE;
    R(fs.flags.reserve((size_t)n_nt * 4 + 64));
    R(fs.tiles.reserve((size_t)(std::max(nt_tiles, n_tiles) + 8) * 4));
    flags = fs.flags.as<uint32_t>();
    tiles = fs.tiles.as<uint32_t>();
    k_db_run_flags<<<(n_nt + 255) / 256, 256, 0, st>>>(brk, n_nt, flags);
    k_db_count_nt<<<nt_tiles, DBT, 0, st>>>(nullptr, n_nt, tiles, 1, flags);
    k_scan_u32_tiles<<<1, DBT, 0, st>>>(tiles, nt_tiles, fs.d_small + 2);
    CK(cudaMemcpyAsync(h_small + 2, fs.d_small + 2, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t n_runs = h_small[2];
    R(fs.run_start.reserve((size_t)(n_runs + 2) * 4));
    R(fs.windows.reserve((size_t)(n_runs + 2) * 4 + 64));
    uint32_t* run_start = fs.run_start.as<uint32_t>();
    uint32_t* windows = fs.windows.as<uint32_t>();
    k_db_emit_flagged<<<nt_tiles, DBT, 0, st>>>(flags, n_nt, tiles, run_start);
    k_db_run_windows<<<(n_runs + 255) / 256, 256, 0, st>>>(run_start, n_runs, n_nt, k, windows);
    // exclusive scan of windows (n_runs is small next to n_nt: one block)
    k_scan_u32_tiles<<<1, DBT, 0, st>>>(windows, n_runs, fs.d_small + 3);
    CK(cudaMemcpyAsync(h_small + 3, fs.d_small + 3, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t n_windows = h_small[3];
    if (n_kmers) *n_kmers += (n_windows + light_gap - 1) / light_gap;
    if (n_windows) {
        const uint64_t threads = ((uint64_t)n_windows + light_gap - 1) / light_gap;
        k_db_kmers_light<<<(uint32_t)((threads + 255) / 256), 256, 0, st>>>(codes, run_start, windows, n_runs, n_windows, k, light_gap, e);
    }
    CK(cudaGetLastError());
#undef R
    return CUCLARK_OK;
}

int key_bytes_auto(int k, uint64_t htsize) {
    // src/main.cc:278-316
    const size_t t_b = (size_t)(log((double)htsize) / log(4.0));
    if ((size_t)k <= t_b + 8) return 2;
    if ((size_t)k <= t_b + 16) return 4;
    return 8;
}

template <typename K>
int reduce_and_write(K* keys, uint16_t* labels, uint32_t* local, uint64_t* tile_base, uint32_t* kept, uint32_t* out_local,
                     uint64_t* out_tile, uint64_t n_seg_tiles, uint64_t htsize, uint32_t min_count, uint64_t total,
                     int key_bytes, const char* out_base, uint32_t* d_err, uint64_t* total_kept, cudaStream_t st) {
    k_db_reduce<K><<<(uint32_t)((htsize + DBT - 1) / DBT), DBT, 0, st>>>(keys, labels, local, tile_base, kept, htsize, min_count,
                                                                        total, d_err);
    CK(cudaGetLastError());
    uint32_t h_err = 0;
    CK(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_err & DBERR_BUCKET_255) {
        set_error("This table can not be stored on disk: Some bucket list size exceeds 255.");     // src/hashTable_hh.hh:620-625
        return CUCLARK_ERR_BUILD;
    }
    // output offsets of the kept entries
    uint64_t* d_total = out_tile + n_seg_tiles;
    k_db_tile_offsets<<<(uint32_t)n_seg_tiles, DBT, 0, st>>>(kept, htsize, out_local, out_tile);
    k_scan_u64<<<1, DBT, 0, st>>>(out_tile, n_seg_tiles, d_total);
    std::vector<uint64_t> h_out_tile(n_seg_tiles + 1);
    CK(cudaMemcpyAsync(h_out_tile.data(), out_tile, (n_seg_tiles + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *total_kept = h_out_tile[n_seg_tiles];
    // write the three files window by window (windows end at tile boundaries)
    const std::string base(out_base);
    const int fsz = open((base + ".sz").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    const int fky = open((base + ".ky").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    const int flb = open((base + ".lb").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    auto close_all = [&] { if (fsz >= 0) close(fsz); if (fky >= 0) close(fky); if (flb >= 0) close(flb); };
    if (fsz < 0 || fky < 0 || flb < 0) { close_all(); set_error("cannot create %s.{sz,ky,lb}", out_base); return CUCLARK_ERR_IO; }
    const uint64_t WIN_TILES = (64ull << 20) / SEG_TILE, WIN_ENTRIES = 128ull << 20;
    uint8_t *d_sz = nullptr, *h_sz = nullptr;
    uint8_t *d_ky = nullptr, *h_ky = nullptr;
    uint16_t *d_lb = nullptr, *h_lb = nullptr;
    size_t cap_entries = 0;
    const size_t win_buckets = WIN_TILES * SEG_TILE;
    int rc = CUCLARK_OK;
    auto cleanup = [&] { cudaFree(d_sz); cudaFree(d_ky); cudaFree(d_lb); cudaFreeHost(h_sz); cudaFreeHost(h_ky); cudaFreeHost(h_lb); close_all(); };
#define W(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(e_)); cleanup(); return CUCLARK_ERR_CUDA; } } while (0)
    W(cudaMalloc(&d_sz, win_buckets));
    W(cudaMallocHost(&h_sz, win_buckets));
    for (uint64_t t0 = 0; t0 < n_seg_tiles;) {
        uint64_t t1 = t0 + 1;
        while (t1 < n_seg_tiles && t1 - t0 < WIN_TILES && h_out_tile[t1 + 1] - h_out_tile[t0] <= WIN_ENTRIES) t1++;
        const uint64_t r0 = t0 * SEG_TILE, nr = std::min<uint64_t>((t1 - t0) * SEG_TILE, htsize - r0);
        const uint64_t out0 = h_out_tile[t0], ne = h_out_tile[t1] - out0;
        if (ne > cap_entries) {
            cudaFree(d_ky); cudaFree(d_lb); cudaFreeHost(h_ky); cudaFreeHost(h_lb);
            d_ky = h_ky = nullptr; d_lb = h_lb = nullptr;
            cap_entries = std::max<uint64_t>(ne, 1 << 20);
            W(cudaMalloc(&d_ky, cap_entries * key_bytes)); W(cudaMalloc(&d_lb, cap_entries * 2));
            W(cudaMallocHost(&h_ky, cap_entries * key_bytes)); W(cudaMallocHost(&h_lb, cap_entries * 2));
        }
        k_db_write<K><<<(uint32_t)((nr + DBT - 1) / DBT), DBT, 0, st>>>(keys, labels, local, tile_base, kept, out_local, out_tile,
                                                                       htsize, r0, nr, out0, d_sz, d_ky, d_lb, key_bytes);
        W(cudaGetLastError());
        W(cudaMemcpyAsync(h_sz, d_sz, nr, cudaMemcpyDeviceToHost, st));
        if (ne) {
            W(cudaMemcpyAsync(h_ky, d_ky, ne * key_bytes, cudaMemcpyDeviceToHost, st));
            W(cudaMemcpyAsync(h_lb, d_lb, ne * 2, cudaMemcpyDeviceToHost, st));
        }
        W(cudaStreamSynchronize(st));
        if (write_all(fsz, h_sz, nr) || (ne && (write_all(fky, h_ky, ne * key_bytes) || write_all(flb, h_lb, ne * 2)))) {
            set_error("write to %s.{sz,ky,lb} failed", out_base);
            rc = CUCLARK_ERR_IO;
            break;
        }
        t0 = t1;
    }
#undef W
    cleanup();
    return rc;
}

}  // namespace
}  // namespace cuclark

using namespace cuclark;

extern "C" int cuclark_build_database(const cuclark_build_opts* o, const char* const* target_files,
                                      const uint16_t* target_labels, size_t n_files, const char* out_base,
                                      cuclark_build_stats* stats) {
    if (!o || !target_files || !target_labels || !out_base) { set_error("null argument"); return CUCLARK_ERR_ARG; }
    if (o->k < 2 || o->k > 32) { set_error("The k-mer length should be in [2,32]."); return CUCLARK_ERR_ARG; }
    if (o->htsize < 2) { set_error("htsize must be HTSIZE of the variant"); return CUCLARK_ERR_ARG; }
    if (o->light_gap < 0) { set_error("bad gap"); return CUCLARK_ERR_ARG; }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (there is no CPU fallback)");
        return CUCLARK_ERR_NO_DEVICE;
    }
    if (o->device < 0 || o->device >= n_dev) { set_error("device %d not present", o->device); return CUCLARK_ERR_NO_DEVICE; }
    const bool timing = getenv("CUCLARK_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[cuclark timing] database build: %s %.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t_last).count());
        t_last = n;
    };
    CK(cudaSetDevice(o->device));
    CK(cudaFree(nullptr));
    lap("CUDA context");
    if (stats) memset(stats, 0, sizeof *stats);
    const int k = o->k;
    const uint64_t htsize = o->htsize;
    const int key_bytes = o->key_bytes ? o->key_bytes : key_bytes_auto(k, htsize);
    if (key_bytes != 2 && key_bytes != 4 && key_bytes != 8) { set_error("key_bytes must be 2, 4 or 8"); return CUCLARK_ERR_ARG; }
    // the largest quotient must fit the key type (the reference picks the type from k, src/main.cc:278-316)
    const long double max_q = (k == 32 ? 18446744073709551615.0L : (long double)((1ull << (2 * k)) - 1)) / (long double)htsize;
    if ((key_bytes == 2 && max_q >= 65536.0L) || (key_bytes == 4 && max_q >= 4294967296.0L)) {
        set_error("k = %d does not fit %d-byte keys with HTSIZE %llu", k, key_bytes, (unsigned long long)htsize);
        return CUCLARK_ERR_ARG;
    }
    const bool wide = key_bytes == 8;
    cudaStream_t st = nullptr;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const uint64_t n_seg_tiles = (htsize + SEG_TILE - 1) / SEG_TILE;
    uint32_t *count = nullptr, *local = nullptr, *out_local = nullptr, *d_err = nullptr;
    uint64_t *tile_base = nullptr, *out_tile = nullptr;
    void* keys = nullptr;
    uint16_t* labels = nullptr;
    FileScratch fs;
    std::vector<uint8_t> fastq_text;
    auto cleanup = [&] {
        cudaFree(count); cudaFree(local); cudaFree(out_local); cudaFree(d_err); cudaFree(tile_base); cudaFree(out_tile);
        cudaFree(keys); cudaFree(labels);
        cudaStreamDestroy(st);
    };
#define B(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(e_)); cleanup(); return CUCLARK_ERR_CUDA; } } while (0)
    B(cudaMalloc(&count, htsize * 4));
    B(cudaMalloc(&local, htsize * 4));
    B(cudaMalloc(&tile_base, (n_seg_tiles + 2) * 8));
    B(cudaMalloc(&d_err, 64));
    B(cudaMemsetAsync(count, 0, htsize * 4, st));
    B(cudaMemsetAsync(d_err, 0, 64, st));
    Emit e;
    memset(&e, 0, sizeof e);
    e.count = count; e.htsize = htsize; e.magic = ~0ull / htsize; e.wide = wide;
    // the two passes over the target files
    uint64_t total = 0, n_nt = 0;
    for (int pass = 0; pass < 2; pass++) {
        e.pass_b = pass;
        uint64_t nt_pass = 0;
        for (size_t f = 0; f < n_files; f++) {
            FileMap fm;
            if (!fm.open_file(target_files[f])) {
                if (pass == 0) fprintf(stderr, "Failed to open %s\n", target_files[f]);           // src/CuCLARK_hh.hh:701-704
                continue;
            }
            if (fm.n == 0) continue;
            const uint8_t* text = fm.p;
            size_t text_n = fm.n;
            if (fm.p[0] == '@') {                    // FASTQ target
                fastq_target_as_fasta(fm.p, fm.n, fastq_text);
                text = fastq_text.data(); text_n = fastq_text.size();
                if (text_n == 0) continue;
            } else if (fm.p[0] != '>') {
                set_error("%s: targets must be FASTA or FASTQ files (first byte '%c'; spectrum files are not supported)", target_files[f], fm.p[0]);
                cleanup();
                return CUCLARK_ERR_FORMAT;
            }
            e.label = target_labels[f];
            const int rc = scan_file(fs, text, text_n, k, o->light_gap, e, nullptr, &nt_pass, st);
            if (rc) { cleanup(); return rc; }
            B(cudaStreamSynchronize(st));
        }
        n_nt = nt_pass;
        lap(pass == 0 ? "pass 1 over the targets (count)" : "pass 2 over the targets (scatter)");
        if (pass == 0) {
            uint64_t* d_total = tile_base + n_seg_tiles;
            k_db_tile_offsets<<<(uint32_t)n_seg_tiles, DBT, 0, st>>>(count, htsize, local, tile_base);
            k_scan_u64<<<1, DBT, 0, st>>>(tile_base, n_seg_tiles, d_total);
            B(cudaGetLastError());
            B(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st));
            B(cudaStreamSynchronize(st));
            B(cudaMalloc(&keys, std::max<uint64_t>(total, 1) * (wide ? 8 : 4)));
            B(cudaMalloc(&labels, std::max<uint64_t>(total, 1) * 2));
            e.local = local; e.tile_base = tile_base; e.keys = keys; e.labels = labels;
        }
    }
    // count[] is all zero again: reuse it for the kept sizes
    B(cudaMalloc(&out_local, htsize * 4));
    B(cudaMalloc(&out_tile, (n_seg_tiles + 2) * 8));
    uint64_t total_kept = 0;
    int rc;
    if (wide)
        rc = reduce_and_write<uint64_t>((uint64_t*)keys, labels, local, tile_base, count, out_local, out_tile, n_seg_tiles, htsize,
                                        o->min_count, total, key_bytes, out_base, d_err, &total_kept, st);
    else
        rc = reduce_and_write<uint32_t>((uint32_t*)keys, labels, local, tile_base, count, out_local, out_tile, n_seg_tiles, htsize,
                                        o->min_count, total, key_bytes, out_base, d_err, &total_kept, st);
    lap("per-bucket sort, RemoveCommon and file write");
    if (stats) {
        stats->n_nucleotides = n_nt;
        stats->n_kmers_added = total;
        stats->n_kmers_kept = total_kept;
        stats->key_bytes = key_bytes;
    }
#undef B
    cleanup();
    return rc;
}
