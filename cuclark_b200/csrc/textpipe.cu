// cuclark_b200 — text stages on the device.
//
// The reference indexes and packs reads on the host, byte by byte, with OpenMP
// over batches only (src/CuCLARK_hh.hh:1340-1534 index, :1616-1708 pack) and
// formats the CSV with one fprintf per read (:1951-2139). Here the raw
// FASTA/FASTQ bytes of a chunk are copied to HBM once and everything else
// happens there:
//
//   k_tp_count / k_tp_lines   newline + header-line table of the chunk (two passes over the
//                             text, 16 bytes per thread, block scans; the chunk stays in L2)
//   k_tp_records_*            one thread per record: name span, sequence span, Length
//   k_tp_pack<false/true>     one warp per read: count containers, then (after a scan that
//                             yields readsPointer) write them — the reference's exact format
//   k_tp_csv<false/true>      one thread per read: line length, then (after a scan) the bytes,
//                             with an exact integer implementation of printf("%g") (fmt_g.h)
//
// All are streaming byte kernels bound by HBM/L2 bandwidth; algorithmic bytes per
// read: text bytes in (x2 passes for the line table, x2 for pack), 2*(1+ceil(L/8)) + 4
// out, CSV line out.
#include <algorithm>

#include "fmt_g.h"
#include "internal.h"
#include "textpipe.cuh"

namespace cuclark {

namespace {

constexpr int TPT = 256;                 // threads per block
constexpr int TPB = 16;                  // bytes per thread
constexpr int TP_TILE = TPT * TPB;       // bytes per block
constexpr int SCAN_ITEMS = 8;            // scan: items per thread
constexpr int SCAN_TILE = TPT * SCAN_ITEMS;

// exclusive scan of one value per thread over a 256-thread block; *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t& total) {
    __shared__ uint32_t ws[TPT / 32];
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
        uint32_t x = lane < TPT / 32 ? ws[lane] : 0, xi = x;
#pragma unroll
        for (int o = 1; o < TPT / 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, xi, o);
            if (lane >= o) xi += t;
        }
        if (lane < TPT / 32) ws[lane] = xi - x;
        if (lane == TPT / 32 - 1) wtotal = xi;
    }
    __syncthreads();
    const uint32_t res = inc - v + ws[w];
    total = wtotal;
    __syncthreads();
    return res;
}

// ---- generic in-place exclusive scan of uint32 arrays ---------------------------
__global__ void __launch_bounds__(TPT) k_scan_reduce(const uint32_t* __restrict__ data, uint32_t n, uint32_t* tiles) {
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) s += base + i < n ? data[base + i] : 0u;
    uint32_t total;
    block_excl_scan(s, total);
    if (threadIdx.x == 0) tiles[blockIdx.x] = total;
}

// single block: exclusive scan of tiles[0..n_tiles) in place; 64-bit total
__global__ void __launch_bounds__(TPT) k_scan_tiles(uint32_t* tiles, uint32_t n_tiles, uint64_t* total_out) {
    uint64_t carry = 0;
    for (uint32_t b = 0; b < n_tiles; b += TPT) {
        const uint32_t i = b + threadIdx.x;
        const uint32_t v = i < n_tiles ? tiles[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_excl_scan(v, total);
        if (i < n_tiles) tiles[i] = (uint32_t)carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(TPT) k_scan_apply(uint32_t* data, uint32_t n, const uint32_t* __restrict__ tiles,
                                                    const uint64_t* __restrict__ total) {
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = base + i < n ? data[base + i] : 0u; s += v[i]; }
    uint32_t bt;
    uint32_t run = block_excl_scan(s, bt) + tiles[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) data[n] = (uint32_t)*total;
}

// data[0..n) -> exclusive prefix sums in place, data[n] = total (low 32 bits), *total64 = total
int scan_inplace(uint32_t* data, uint32_t n, uint32_t* tiles, size_t cap_tiles, uint64_t* total64, cudaStream_t st) {
    const uint32_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (n_tiles > cap_tiles) { set_error("scan: %u tiles exceed the scratch (%zu)", n_tiles, cap_tiles); return CUCLARK_ERR_NOMEM; }
    if (n_tiles) k_scan_reduce<<<n_tiles, TPT, 0, st>>>(data, n, tiles);
    k_scan_tiles<<<1, TPT, 0, st>>>(tiles, n_tiles, total64);
    if (n_tiles) k_scan_apply<<<n_tiles, TPT, 0, st>>>(data, n, tiles, total64);
    else CK(cudaMemsetAsync(data, 0, 4, st));
    CK(cudaGetLastError());
    count_launches(n_tiles ? 3 : 1);
    return CUCLARK_OK;
}

// ---- stage 1a: line table -----------------------------------------------------------
struct ByteVec { uint8_t b[TPB]; };

__device__ __forceinline__ ByteVec load16(const uint8_t* p) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    ByteVec r;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 16; i++) r.b[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
    return r;
}

// newlines and header lines ('>' at a line start) of this thread's 16 bytes
__device__ __forceinline__ void count16(const uint8_t* text, uint32_t n, uint32_t base, uint32_t& nl, uint32_t& hd,
                                        ByteVec& v, uint8_t& prev) {
    nl = 0; hd = 0;
    prev = '\n';                                       // the chunk starts at a line start
    if (base >= n) return;
    v = load16(text + base);
    if (base) prev = text[base - 1];
    uint8_t p = prev;
#pragma unroll
    for (int i = 0; i < TPB; i++) {
        if (base + i < n) {
            nl += v.b[i] == '\n';
            hd += (v.b[i] == '>') & (p == '\n');
        }
        p = v.b[i];
    }
}

__global__ void __launch_bounds__(TPT) k_tp_count(const uint8_t* __restrict__ text, uint32_t n, uint32_t* tile_nl,
                                                  uint32_t* tile_hd) {
    const uint32_t base = blockIdx.x * TP_TILE + threadIdx.x * TPB;
    uint32_t nl, hd;
    ByteVec v;
    uint8_t prev;
    count16(text, n, base, nl, hd, v, prev);
    uint32_t tn, th;
    block_excl_scan(nl, tn);
    block_excl_scan(hd, th);
    if (threadIdx.x == 0) { tile_nl[blockIdx.x] = tn; tile_hd[blockIdx.x] = th; }
}

// totals -> ChunkInfo; sentinel entries of the line table
__global__ void k_tp_finish_counts(const uint8_t* __restrict__ text, uint32_t n, const uint64_t* tot_nl,
                                   const uint64_t* tot_hd, uint32_t* line_start, uint32_t cap_lines, ChunkInfo* info) {
    const uint32_t nl = (uint32_t)*tot_nl, hd = (uint32_t)*tot_hd;
    const bool open_tail = n > 0 && text[n - 1] != '\n';
    info->n_newlines = nl;
    info->n_headers = hd;
    info->n_lines = nl + (open_tail ? 1u : 0u);
    info->err = 0;
    info->n_reads = 0;
    info->n_cont = 0;
    info->csv_bytes = 0;
    if (nl + 2 > cap_lines) { info->err |= TP_ERR_LINES; return; }
    line_start[0] = 0;
    // line L spans [line_start[L], line_start[L+1] - 1): give an unterminated last line an end
    if (open_tail) line_start[nl + 1] = n + 1;
}

__global__ void __launch_bounds__(TPT) k_tp_lines(const uint8_t* __restrict__ text, uint32_t n,
                                                  const uint32_t* __restrict__ tile_nl, const uint32_t* __restrict__ tile_hd,
                                                  uint32_t* line_start, uint32_t* hdr_line, uint32_t cap_lines,
                                                  uint32_t cap_reads, ChunkInfo* info, bool fasta) {
    const uint32_t base = blockIdx.x * TP_TILE + threadIdx.x * TPB;
    uint32_t nl, hd;
    ByteVec v;
    uint8_t prev;
    count16(text, n, base, nl, hd, v, prev);
    uint32_t tn, th;
    uint32_t l = block_excl_scan(nl, tn) + tile_nl[blockIdx.x];     // newlines before this thread's bytes
    uint32_t h = block_excl_scan(hd, th) + tile_hd[blockIdx.x];
    if (base >= n || (info->err & TP_ERR_LINES)) return;
    uint8_t p = prev;
    bool too_many_reads = false;
#pragma unroll
    for (int i = 0; i < TPB; i++) {
        if (base + i < n) {
            if (fasta && v.b[i] == '>' && p == '\n') {
                if (h < cap_reads) hdr_line[h] = l; else too_many_reads = true;
                h++;
            }
            if (v.b[i] == '\n') { l++; line_start[l] = base + i + 1; }   // l <= n_newlines < cap_lines - 1
        }
        p = v.b[i];
    }
    if (too_many_reads) atomicOr(&info->err, TP_ERR_READS);
}

__device__ __forceinline__ bool is_sep(uint8_t c) { return c == ' ' || c == '\t' || c == '\n'; }

// Name = bytes after '>'/'@' up to the first separator; the first name byte is never tested
// (src/CuCLARK_hh.hh:1369-1372, 1498-1501). A header with an empty name makes the reference
// lose the record structure; here the name is empty and the record structure is kept.
__device__ __forceinline__ void name_span(const uint8_t* text, uint32_t n, uint32_t hs, uint32_t he, uint32_t& ns,
                                          uint32_t& ne) {
    ns = min(hs + 1, n);
    if (ns >= he) { ne = ns; return; }
    uint32_t i = ns + 1;
    while (i < he && !is_sep(text[i])) i++;
    ne = i;                                             // he is the newline (a separator) or n
}

// FASTQ: strictly four lines per record (src/CuCLARK_hh.hh:1474-1523)
__global__ void __launch_bounds__(TPT) k_tp_records_fastq(const uint8_t* __restrict__ text, uint32_t n,
                                                          const uint32_t* __restrict__ line_start, ChunkInfo* info,
                                                          uint32_t cap_reads, uint32_t* name_s, uint32_t* name_e,
                                                          uint32_t* seq_s, uint32_t* seq_e, uint32_t* len) {
    if (info->err) return;
    const uint32_t n_lines = info->n_lines;
    uint32_t R = (n_lines + 3) / 4;
    // a record starts only if at least one byte follows its '@' (:1520 `if ((++i) >= iNext) break`)
    if (R > 1 && line_start[4 * (R - 1)] + 1 >= n) R--;
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) {
        info->n_reads = min(R, cap_reads);
        if (R > cap_reads) atomicOr(&info->err, TP_ERR_READS);
    }
    if (r >= R || r >= cap_reads) return;
    const uint32_t L0 = 4 * r;
    const uint32_t hs = line_start[L0], he = line_start[L0 + 1] - 1;
    uint32_t ns, ne;
    name_span(text, n, hs, min(he, n), ns, ne);
    uint32_t ss = n, se = n;
    if (L0 + 1 < n_lines) { ss = line_start[L0 + 1]; se = line_start[L0 + 2] - 1; }
    name_s[r] = ns; name_e[r] = ne; seq_s[r] = ss; seq_e[r] = se; len[r] = se - ss;
}

// FASTA: a record runs from a header line to the line before the next header; multi-line
// sequences allowed; Length = bytes - one per line (src/CuCLARK_hh.hh:1362-1390)
__global__ void __launch_bounds__(TPT) k_tp_records_fasta(const uint8_t* __restrict__ text, uint32_t n,
                                                          const uint32_t* __restrict__ line_start,
                                                          const uint32_t* __restrict__ hdr_line, ChunkInfo* info,
                                                          uint32_t cap_reads, uint32_t* name_s, uint32_t* name_e,
                                                          uint32_t* seq_s, uint32_t* seq_e, uint32_t* len) {
    if (info->err) return;
    const uint32_t n_lines = info->n_lines, R = info->n_headers;
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) info->n_reads = min(R, cap_reads);
    if (r >= R || r >= cap_reads) return;
    const uint32_t h = hdr_line[r];
    const uint32_t h2 = r + 1 < R ? hdr_line[r + 1] : n_lines;
    const uint32_t hs = line_start[h], he = min(line_start[h + 1] - 1, n);
    uint32_t ns, ne;
    name_span(text, n, hs, he, ns, ne);
    const uint32_t lines = h2 - h - 1;
    const uint32_t ss = min(line_start[h + 1], n);
    const uint32_t se = lines ? line_start[h2] - 1 : ss;
    name_s[r] = ns; name_e[r] = ne; seq_s[r] = ss; seq_e[r] = se; len[r] = se - ss + 1 - lines;
}

// ---- stage 1b: 2-bit packing ----------------------------------------------------------
// 0..3: packed (complement) code A=3 C=2 G=1 T/U=0 (m_rTable, src/CuCLARK_hh.hh:291-295);
// 4: newline, transparent; 5: anything else ends the part (:1646-1697)
__device__ __forceinline__ int nt_class(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 3;
        case 'C': case 'c': return 2;
        case 'G': case 'g': return 1;
        case 'T': case 't': case 'U': case 'u': return 0;
        case '\n': return 4;
        default: return 5;
    }
}

constexpr int PACK_WARPS = 8;

// A part header is a uint16 (src/CuCLARK_hh.hh:1616-1708 sums the run length into one): a run of 65,536
// nucleotides or more wraps it in the reference, whose kernel then reads data containers as headers
// (undefined results). Here such a run becomes several parts of at most MAX_PART nt that OVERLAP by k-1 nt, so
// that every k-mer of the run is still looked up exactly once (deviation Q8; oracle: orc_pack).
constexpr uint32_t MAX_PART = 65535;

// appends the `cnt` staged codes sbuf[fill .. fill+cnt) to the open part (run nucleotides so far, fill = run & 7
// of them pending in sbuf[0..fill)): full containers go out, the remainder is moved to the front of sbuf
__device__ __forceinline__ void pack_flush(uint16_t* out, uint32_t limit, uint32_t cc, uint32_t run, uint32_t cnt,
                                           uint8_t* sbuf, int lane) {
    const uint32_t fill = run & 7;
    const uint32_t total = fill + cnt, full = total >> 3;
    if ((uint32_t)lane < full) {
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) w = (w << 2) | sbuf[8 * lane + i];
        const uint32_t idx = cc + 1 + (run >> 3) + lane;
        if (idx < limit) out[idx] = (uint16_t)w;
    }
    const uint8_t keep = (uint32_t)lane < (total & 7) ? sbuf[8 * full + lane] : 0;
    __syncwarp();
    if ((uint32_t)lane < (total & 7)) sbuf[lane] = keep;
    __syncwarp();
}

// One warp packs (WRITE) or sizes (!WRITE) one read. Returns the number of containers.
// `limit` (WRITE only) = containers this read owns: writes beyond it belong to parts that the
// reference drops by rewinding its write index (:1699-1703) and must not touch the next read.
template <bool WRITE>
__device__ __forceinline__ uint32_t pack_read(const uint8_t* __restrict__ text, uint32_t s, uint32_t e, int k,
                                              uint16_t* out, uint32_t limit, uint8_t* sbuf, int lane) {
    uint32_t cc = 0;          // containers of the finished, kept parts
    uint32_t run = 0;         // nucleotides of the open part
    const uint32_t lt = (1u << lane) - 1;
    for (uint32_t p0 = s; p0 <= e; p0 += 32) {
        const uint32_t p = p0 + lane;
        int cls = 6;                                         // beyond the read
        if (p < e) cls = nt_class(text[p]);
        else if (p == e) cls = 5;                            // the end of the read closes the open part
        const uint32_t ntm = __ballot_sync(0xFFFFFFFFu, cls < 4);
        const uint32_t brk = __ballot_sync(0xFFFFFFFFu, cls == 5);
        int lo = 0;
        for (;;) {
            const uint32_t rest = lo < 32 ? (brk >> lo) << lo : 0u;
            int hi = rest ? __ffs(rest) - 1 : 32;
            uint32_t segmask = (hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1)) & ~((1u << lo) - 1);
            uint32_t nts = ntm & segmask;
            uint32_t cnt = __popc(nts);
            bool split = false;                              // the part is full: close it BEFORE lane hi
            if (run + cnt > MAX_PART) {
                cnt = MAX_PART - run;
                hi = (int)__fns(nts, 0, (int)cnt + 1);       // the first nucleotide that does not fit
                segmask &= (1u << hi) - 1;
                nts &= segmask;
                split = true;
            }
            if (cnt) {
                if (WRITE) {
                    if ((nts >> lane) & 1) sbuf[(run & 7) + __popc(nts & lt)] = (uint8_t)cls;
                    __syncwarp();
                    pack_flush(out, limit, cc, run, cnt, sbuf, lane);
                }
                run += cnt;
            }
            if (hi >= 32) break;
            // a break at lane hi (or a full part) closes the open part
            if (run) {
                const uint32_t hdr = run;                        // <= MAX_PART: fits the uint16 header
                if (WRITE && lane == 0) {
                    const uint32_t fill = run & 7;
                    if (fill) {
                        uint32_t w = 0;
                        for (uint32_t i = 0; i < fill; i++) w = (w << 2) | sbuf[i];
                        const uint32_t idx = cc + 1 + (run >> 3);
                        if (idx < limit) out[idx] = (uint16_t)(w << (2 * (8 - fill)));
                    }
                    if (hdr >= (uint32_t)k && cc < limit) out[cc] = (uint16_t)hdr;
                }
                if (WRITE) __syncwarp();
                const uint32_t old_cc = cc;
                if (hdr >= (uint32_t)k) cc += 1 + ((run + 7) >> 3);
                run = 0;
                if (split) {
                    // the next part starts with the last k-1 nucleotides of the one just closed
                    const uint32_t ov = (uint32_t)k - 1;
                    if (WRITE) {
                        if ((uint32_t)lane < ov) {
                            const uint32_t j = MAX_PART - ov + lane;
                            const uint32_t idx = old_cc + 1 + (j >> 3);
                            sbuf[lane] = idx < limit ? (uint8_t)((out[idx] >> (2 * (7 - (j & 7)))) & 3u) : 0;
                        }
                        __syncwarp();
                        pack_flush(out, limit, cc, 0, ov, sbuf, lane);
                    }
                    run = ov;
                }
            }
            lo = split ? hi : hi + 1;
            if (lo >= 32) break;
        }
    }
    return cc;
}

template <bool WRITE>
__global__ void __launch_bounds__(PACK_WARPS * 32) k_tp_pack(const uint8_t* __restrict__ text, uint32_t n_reads, int k,
                                                             const uint32_t* __restrict__ seq_s,
                                                             const uint32_t* __restrict__ seq_e,
                                                             const uint32_t* __restrict__ len, uint32_t* reads_ptr,
                                                             uint16_t* cont, const ChunkInfo* info) {
    __shared__ uint8_t sbuf_all[PACK_WARPS][48];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * PACK_WARPS + w;
    if (r >= n_reads) return;
    if (WRITE && info->err) return;                           // the containers would not fit: the host reports it
    const bool live = len[r] >= (uint32_t)k;                  // :1633
    if (WRITE) {
        const uint32_t o = reads_ptr[r], lim = reads_ptr[r + 1] - o;
        if (live && lim) pack_read<true>(text, seq_s[r], seq_e[r], k, cont + o, lim, sbuf_all[w], lane);
    } else {
        const uint32_t c = live ? pack_read<false>(text, seq_s[r], seq_e[r], k, nullptr, 0, sbuf_all[w], lane) : 0u;
        if (lane == 0) reads_ptr[r] = c;
    }
}

__global__ void k_tp_check_cont(ChunkInfo* info, uint64_t cap_cont) {
    if (info->n_cont > cap_cont || info->n_cont > 0xFFFFFFFFull) info->err |= TP_ERR_CONT;
}

// ---- stage 4': CSV ---------------------------------------------------------------------
struct CsvParams {
    const uint8_t* text;
    const uint32_t *name_s, *name_e, *len;
    const uint16_t* final5;
    const uint16_t* rows;
    NameTable names;
    uint32_t first, n;
    int k, row_pairs;
    uint32_t n_targets;
    bool paired, extended;
    uint32_t* off;
    char* out;
    const ChunkInfo* info;
};

template <bool WRITE>
struct Sink {
    char* p;
    uint32_t n = 0;
    __device__ __forceinline__ void put(char c) { if (WRITE) p[n] = c; n++; }
    __device__ __forceinline__ void put_u32(uint32_t v) {
        char t[10];
        const int d = fmt_u32(v, t);
        for (int i = 0; i < d; i++) put(t[i]);
    }
    __device__ __forceinline__ void put_g(double d) {
        char t[16];
        const int m = fmt_g(d, t);
        for (int i = 0; i < m; i++) put(t[i]);
    }
    __device__ __forceinline__ void put_name(const NameTable& nt, uint32_t idx) {
        if (idx >= nt.n_names) idx = 0;
        for (uint32_t i = nt.off[idx]; i < nt.off[idx + 1]; i++) put(nt.chars[i]);
    }
};

// src/CuCLARK_hh.hh:2014-2031 (extended columns), :2097-2135 (line)
template <bool WRITE>
__device__ __forceinline__ uint32_t csv_line(const CsvParams& P, uint32_t r, char* dst) {
    Sink<WRITE> s{dst};
    uint32_t nl = P.name_e[r] - P.name_s[r];
    if (nl >= 40) nl = 39;                                   // OBJECTNAMEMAX, src/parameters.hh:46
    const uint8_t* nm = P.text + P.name_s[r];
    for (uint32_t i = 0; i < nl; i++) {
        const char c = (char)nm[i];
        if (!c) break;                                       // "%s" stops at a NUL
        s.put(c);
    }
    if (P.extended) {
        const uint16_t* R = P.rows + (size_t)r * (2 * P.row_pairs + 2);
        uint32_t w = 0;
        const uint32_t cnt = min((uint32_t)R[0], (uint32_t)P.row_pairs);
        for (uint32_t i = 0; i < cnt; i++) {
            const uint32_t t = R[1 + 2 * i];
            for (; w < t; w++) { s.put(','); s.put('0'); }
            s.put(',');
            s.put_u32(R[2 + 2 * i]);
            w++;
        }
        for (; w < P.n_targets; w++) { s.put(','); s.put('0'); }
    }
    const uint16_t* v = P.final5 + (size_t)r * 5;
    const uint32_t total = v[0], i1 = v[1], best = v[2], i2 = v[3], sbest = v[4];
    const uint32_t norm = P.paired ? P.len[r] - 1u : P.len[r];    // NBN = 1, :2112
    s.put(','); s.put_u32(norm);
    s.put(','); s.put_g(csv_gamma(total, norm, P.k));
    s.put(','); s.put_name(P.names, i1);
    s.put(','); s.put_u32(best);
    s.put(','); s.put_name(P.names, i2);
    s.put(','); s.put_u32(sbest);
    s.put(','); s.put_g(csv_confidence(best, sbest));
    s.put('\n');
    return s.n;
}

template <bool WRITE>
__global__ void __launch_bounds__(TPT) k_tp_csv(const CsvParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    if (WRITE && P.info->err) return;                        // the text would not fit: the host reports it
    const uint32_t r = P.first + i;
    if (WRITE) csv_line<true>(P, r, P.out + P.off[i]);
    else P.off[i] = csv_line<false>(P, r, nullptr);
}

__global__ void k_tp_check_csv(ChunkInfo* info, uint64_t cap_csv) {
    if (info->csv_bytes > cap_csv) info->err |= TP_ERR_CSV;
}

}  // namespace

int tp_index_launch(const TextSlotDev& s, uint32_t n, bool fastq, cudaStream_t st) {
    const uint32_t n_tiles = (n + TP_TILE - 1) / TP_TILE;
    if (n_tiles > s.cap_tiles) { set_error("chunk of %u bytes exceeds the slot", n); return CUCLARK_ERR_ARG; }
    uint64_t* tot_nl = reinterpret_cast<uint64_t*>(s.tile_a + s.cap_tiles);       // two uint64 after each tile array
    uint64_t* tot_hd = reinterpret_cast<uint64_t*>(s.tile_b + s.cap_tiles);
    if (n_tiles) k_tp_count<<<n_tiles, TPT, 0, st>>>(s.text, n, s.tile_a, s.tile_b);
    k_scan_tiles<<<1, TPT, 0, st>>>(s.tile_a, n_tiles, tot_nl);
    k_scan_tiles<<<1, TPT, 0, st>>>(s.tile_b, n_tiles, tot_hd);
    k_tp_finish_counts<<<1, 1, 0, st>>>(s.text, n, tot_nl, tot_hd, s.line_start, (uint32_t)s.cap_lines, s.info);
    if (n_tiles)
        k_tp_lines<<<n_tiles, TPT, 0, st>>>(s.text, n, s.tile_a, s.tile_b, s.line_start, s.hdr_line, (uint32_t)s.cap_lines,
                                            (uint32_t)s.cap_reads, s.info, !fastq);
    // records: the grid covers the most reads the chunk can hold (fastq: >= 2 bytes per line)
    const uint64_t max_reads = fastq ? (uint64_t)n / 8 + 1 : (uint64_t)n / 2 + 1;
    const uint32_t grid = (uint32_t)((std::min<uint64_t>(max_reads, s.cap_reads) + TPT - 1) / TPT);
    if (fastq)
        k_tp_records_fastq<<<grid, TPT, 0, st>>>(s.text, n, s.line_start, s.info, (uint32_t)s.cap_reads, s.name_s, s.name_e,
                                                 s.seq_s, s.seq_e, s.len);
    else
        k_tp_records_fasta<<<grid, TPT, 0, st>>>(s.text, n, s.line_start, s.hdr_line, s.info, (uint32_t)s.cap_reads,
                                                 s.name_s, s.name_e, s.seq_s, s.seq_e, s.len);
    CK(cudaGetLastError());
    count_launches(n_tiles ? 6 : 4);
    return CUCLARK_OK;
}

int tp_pack_launch(const TextSlotDev& s, uint32_t n_bytes, uint32_t n_reads, int k, cudaStream_t st) {
    (void)n_bytes;
    if (n_reads > s.cap_reads) { set_error("too many reads for the slot"); return CUCLARK_ERR_ARG; }
    const uint32_t grid = (n_reads + PACK_WARPS - 1) / PACK_WARPS;
    if (grid) k_tp_pack<false><<<grid, PACK_WARPS * 32, 0, st>>>(s.text, n_reads, k, s.seq_s, s.seq_e, s.len, s.reads_ptr, nullptr, s.info);
    int rc = scan_inplace(s.reads_ptr, n_reads, s.tile_a, s.cap_tiles, &s.info->n_cont, st);
    if (rc) return rc;
    k_tp_check_cont<<<1, 1, 0, st>>>(s.info, s.cap_cont);
    if (grid) k_tp_pack<true><<<grid, PACK_WARPS * 32, 0, st>>>(s.text, n_reads, k, s.seq_s, s.seq_e, s.len, s.reads_ptr, s.cont, s.info);
    CK(cudaGetLastError());
    count_launches(grid ? 3 : 1);
    return CUCLARK_OK;
}

int tp_csv_launch(const TextSlotDev& s, const NameTable& names, uint32_t first, uint32_t n, int k, bool paired,
                  bool extended, int row_pairs, uint32_t n_targets, cudaStream_t st) {
    CsvParams P;
    P.text = s.text; P.name_s = s.name_s; P.name_e = s.name_e; P.len = s.len;
    P.final5 = s.final5; P.rows = s.rows; P.names = names; P.first = first; P.n = n;
    P.k = k; P.row_pairs = row_pairs; P.n_targets = n_targets; P.paired = paired; P.extended = extended;
    P.off = s.csv_off; P.out = s.csv; P.info = s.info;
    const uint32_t grid = (n + TPT - 1) / TPT;
    if (grid) k_tp_csv<false><<<grid, TPT, 0, st>>>(P);
    int rc = scan_inplace(s.csv_off, n, s.tile_a, s.cap_tiles, &s.info->csv_bytes, st);
    if (rc) return rc;
    k_tp_check_csv<<<1, 1, 0, st>>>(s.info, s.cap_csv);
    if (grid) k_tp_csv<true><<<grid, TPT, 0, st>>>(P);
    CK(cudaGetLastError());
    count_launches(grid ? 3 : 1);
    return CUCLARK_OK;
}

}  // namespace cuclark
