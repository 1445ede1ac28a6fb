// cuclark_b200 — the classification kernels.
//
// Replaces queryKernel/queryElement/mergeKernel/resultKernel of the reference
// (src/CuClarkDB.cu:1045-1471). One WARP per read (the reference: one 64-thread
// block per read, one thread per k-mer, each re-reading up to 8 containers):
//   * lane j keeps 32 nucleotides of the current part as one 64-bit word; a
//     k-mer is a funnel shift of two neighbouring words fetched by shuffle
//     (no shared-memory staging, no per-thread re-assembly);
//   * hashed tables: canonical k-mer -> (home bucket, exact quotient) -> ONE 256-bit
//     load of the 32-byte sector bucket; ILP_ROUNDS independent probes per lane are in
//     flight before any is consumed;
//   * LOCAL tables (common.cuh): the home is chosen by the k-mer's canonical minimizer
//     (rolling window minimum of m-mer hashes across lanes), so the ~4.5 consecutive
//     k-mers that share one probe the same two 128-byte lines and adjacent lanes
//     coalesce; 1.5x the lookups of the hashed kernel, bound by the integer ALU pipe;
//   * per-read target counts live in a 64-slot per-warp hash table in shared
//     memory (the reference zeroes and scans numTargets counters per read), with
//     a register fast path while a read has hit a single target;
//   * top-1/top-2 by one warp max-reduction each over (hits << 16 | ~target):
//     identical to the reference's ascending scan with strict '>' (lowest
//     target index wins ties); sparse rows are emitted in
//     ascending target order only when asked for.
// Reads that hit more than 64 distinct targets are handed to an exact dense
// fallback kernel (the reference is undefined beyond MAXHITS).
#include <algorithm>

#include "hits.cuh"
#include "internal.h"
#include "kmerwin.cuh"
#include "synth.cuh"

namespace cuclark {

namespace {

#ifndef CUCLARK_WPB
#define CUCLARK_WPB 8
#endif
constexpr int WARPS_PER_BLOCK = CUCLARK_WPB;
constexpr int ILP_ROUNDS = 4;         // independent probes per lane
#ifndef CUCLARK_ILP_LOCAL
#define CUCLARK_ILP_LOCAL 1
#endif
#ifndef CUCLARK_LOCAL_BRANCHFREE
#define CUCLARK_LOCAL_BRANCHFREE 1
#endif
#ifndef CUCLARK_LOCAL_SKIP_EMPTY_ROW
#define CUCLARK_LOCAL_SKIP_EMPTY_ROW 1      // the look-ahead row of a part's last group usually holds no m-mer: do not hash it
#endif
#ifndef CUCLARK_LOCAL_MIN_BLOCKS
#define CUCLARK_LOCAL_MIN_BLOCKS 4
#endif
constexpr int ILP_LOCAL = CUCLARK_ILP_LOCAL;                // LOCAL layout: rows per group (more resident warps instead)
constexpr int LOCAL_MIN_BLOCKS = CUCLARK_LOCAL_MIN_BLOCKS;  // resident blocks per SM asked of ptxas for the LOCAL kernel
constexpr int READS_PER_CHUNK = 31;   // one coalesced load of 32 container offsets delimits 31 reads
constexpr int COUNTER_CHUNK = 4;      // dynamic work counter (reset with the other counters)

struct ClassifyParams {
    TableView t;
    const uint32_t* reads_ptr;
    const uint16_t* cont;
    uint32_t n_reads;
    uint32_t n_targets;
    HitSink out;
};

template <int LAYOUT, bool ROWS>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, LAYOUT == LAYOUT_LOCAL ? LOCAL_MIN_BLOCKS : 2) k_classify(const ClassifyParams p) {
    __shared__ uint32_t s_key[WARPS_PER_BLOCK][TSLOTS];
    __shared__ uint32_t s_cnt[WARPS_PER_BLOCK][TSLOTS];
    __shared__ uint16_t s_row[ROWS ? WARPS_PER_BLOCK : 1][ROWS ? 2 * MAX_ROW_PAIRS + 2 : 2];
    __shared__ uint32_t s_ptr[WARPS_PER_BLOCK][36];     // 32 container offsets, first read, number of reads, current read

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* tkey = s_key[wib];
    uint32_t* tcnt = s_cnt[wib];
    uint32_t* sptr = s_ptr[wib];
    tab_clear(tkey, tcnt, lane);

    constexpr int ILP = LAYOUT == LAYOUT_LOCAL ? ILP_LOCAL : ILP_ROUNDS;
    const TableView& T = p.t;
    const int k = T.k;
    const int kshift = 64 - 2 * k;
    uint32_t my_lookups = 0;          // per lane, flushed after every chunk of 31 reads (and before it can wrap)

    auto flush_lookups = [&] {             // one atomic per warp and chunk of 31 reads
        const unsigned long long tot = (unsigned long long)__reduce_add_sync(0xFFFFFFFFu, my_lookups & 0xFFFFu) +
                                       ((unsigned long long)__reduce_add_sync(0xFFFFFFFFu, my_lookups >> 16) << 16);
        if (lane == 0 && tot) atomicAdd(reinterpret_cast<unsigned long long*>(p.out.counters + COUNTER_LOOKUPS), tot);
        my_lookups = 0;
    };
    // Each warp pulls chunks of 31 consecutive reads from a global counter (dynamic
    // balance); one coalesced load fetches the chunk's 32 container offsets.
    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&p.out.counters[COUNTER_CHUNK], 1u);
        chunk = __shfl_sync(0xFFFFFFFFu, chunk, 0);
        const uint64_t base64 = (uint64_t)chunk * READS_PER_CHUNK;
        if (base64 >= p.n_reads) break;
        {
            const uint32_t base = (uint32_t)base64;
            // the chunk's state is parked in shared memory: it is touched once per read, the registers are the row loop's
            __syncwarp();
            sptr[lane] = p.reads_ptr[min(base + lane, p.n_reads)];   // entry i: start of read i = end of read i-1
            if (lane == 0) { sptr[32] = base; sptr[33] = min((uint32_t)READS_PER_CHUNK, p.n_reads - base); sptr[34] = 0; }
            __syncwarp();
        }

        // prefetched first 32 containers of the next read: lane 0 its first part header, lane j container j-1
        uint32_t pf = 0;
        {
            const uint32_t p0 = sptr[0], e0 = sptr[1];
            if (p0 + lane < e0) pf = p.cont[p0 + lane];
        }

        for (;;) {
            __syncwarp();
            const uint32_t ri = sptr[34];                   // (the loop counter lives in shared memory too)
            if (ri >= sptr[33]) break;
            if (__any_sync(0xFFFFFFFFu, my_lookups >= (1u << 30))) flush_lookups();     // a read adds < 2^30 per lane
            uint32_t pos = sptr[ri];
            const uint32_t end = sptr[ri + 1];
            const uint32_t cur = pf;
            if (ri + 1 < sptr[33]) {                        // prefetch the next read while this one probes
                const uint32_t pn = end;
                const uint32_t en = sptr[ri + 2];
                pf = pn + lane < en ? p.cont[pn + lane] : 0;
            }
            __syncwarp();
            if (lane == 0) sptr[34] = ri + 1;
            const uint32_t cur_hdr = __shfl_sync(0xFFFFFFFFu, cur, 0);
            uint32_t cur_win = __shfl_down_sync(0xFFFFFFFFu, cur, 1);                    // lane 31: loaded on demand
            WarpHits hits;
            bool first_part = true;
            while (pos < end) {
                const uint32_t Lraw = first_part ? cur_hdr : (uint32_t)p.cont[pos];
                const uint32_t first = pos + 1;
                uint32_t L;
                const uint32_t ncont = part_extent(Lraw, first, end, L);
                pos = first + ncont;                         // start of the next part
                const int nk = (int)L - k + 1;
                for (int cb = 0; cb < nk; cb += 32 * CHUNK_ROUNDS) {
                    // lane j: nucleotides [cb + 32j, cb + 32j + 32) of the part, MSB first
                    const uint32_t co = (uint32_t)cb >> 3;   // first container of this chunk
                    const int nwin = (int)min(4u, (ncont - co + 31) >> 5);
                    uint32_t wv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t ci = first + co + 32 * u + lane;
                        if (u == 0 && first_part && cb == 0) wv[u] = lane < 31 ? cur_win : ncont > 31 ? (uint32_t)p.cont[first + 31] : 0u;
                        else wv[u] = (u < nwin && ci < pos) ? (uint32_t)p.cont[ci] : 0u;
                    }
                    const uint64_t W = assemble_words(wv, nwin, lane);
                    const int rounds = min(CHUNK_ROUNDS, (nk - cb + 31) >> 5);
                    bool have_carry = false;                 // LOCAL: hash row handed over from the previous group
                    uint64_t carry_z = 0, carry_c = 0;       // carry_z bit 61: the k-mer stands in its canonical form
                    uint32_t carry_oh = 0;
                    const int m_limit = (int)L - (k - LOCAL_W + 1) - cb - lane;   // an m-mer starts at row i iff 32 i <= m_limit
                    for (int r0 = 0; r0 < rounds; r0 += ILP) {
                        uint64_t q[ILP];
                        uint32_t lb[ILP];
                        Sector sec[ILP];
                        bool live[ILP];
                        uint64_t cc[LAYOUT == LAYOUT_LOCAL ? ILP : 1];   // LOCAL: the k-mer itself (overflow key)
                        Sector secb[LAYOUT == LAYOUT_LOCAL ? ILP : 1];   // LOCAL: sector of the second candidate line
                        if (LAYOUT == LAYOUT_LOCAL) {
                            // ---- minimizer of every k-mer of the ILP rows. Each lane hashes the FIRST m-mer of
                            // its position (rows r0..r0+ILP: the windows of the last row reach 7 positions
                            // into the next one); a windowed minimum over 8 consecutive positions by doubling
                            // across lanes gives every k-mer its minimizer: the LEFTMOST smallest hash in read
                            // orientation. Where the smallest hash occurs twice the canonical k-mer may have
                            // its other home (rightmost): the table keeps such k-mers in the overflow table
                            // and flags both homes (common.cuh, "TIES").
                            constexpr int NR = ILP + 1;
                            const int m = k - LOCAL_W + 1, mbits = 2 * m;
                            const uint64_t mmask = (~0ull) >> (64 - mbits);
                            uint64_t cs[ILP];                 // canonical k-mers
                            uint64_t zs[NR];                  // mixed canonical m-mer | (a < b) << 63 | (a > b) << 62 | (x <= rc) << 61
                            uint32_t AL[NR];
#pragma unroll
                            for (int R = 0; R < NR; R++) {
                                const int i = r0 + R;
                                uint64_t c;
                                uint32_t oh;
                                if (R == 0 && have_carry) {           // warp-uniform: the previous group's tail row
                                    c = carry_c; zs[R] = carry_z; oh = carry_oh;
                                } else if (CUCLARK_LOCAL_SKIP_EMPTY_ROW && (i > 31 || 32 * i > m_limit + lane)) {
                                    // warp-uniform: no m-mer of the part starts in this row (the look-ahead row behind
                                    // the last k-mers, unless the part ends within 7 nt of a row boundary)
                                    c = 0; zs[R] = 0; oh = LOCAL_ORDER_MAX + 1;
                                } else {
                                    uint64_t x, rc;
                                    const uint64_t hi = shfl64(W, i & 31), lo = shfl64(W, (i + 1) & 31);
                                    x = window64(hi, lo, 2 * lane) >> kshift;
                                    rc = revcomp2(x, k);
                                    // first m-mer of x and its reverse complement (= last m-mer of rc); the x of a
                                    // row past the 32 words has undefined low bits, which neither of the two touches
                                    const uint64_t a = x >> (2 * (LOCAL_W - 1)), b = rc & mmask;
                                    const bool lt = a < b;
                                    const uint64_t z = local_mix(lt ? a : b, mbits);
                                    oh = local_order(z, mbits);
                                    if (i > 31 || 32 * i > m_limit) oh = LOCAL_ORDER_MAX + 1;   // no m-mer here
                                    const bool kf = x <= rc;
                                    c = kf ? x : rc;
                                    zs[R] = z | ((uint64_t)lt << 63) | ((uint64_t)(!lt && a != b) << 62) | ((uint64_t)kf << 61);
                                }
                                if (R < ILP) cs[R < ILP ? R : 0] = c;
                                if (R == NR - 1) { carry_c = c; carry_z = zs[R]; carry_oh = oh; }
                                AL[R] = (oh << 8) | (uint32_t)(32 * R + lane);
                            }
                            have_carry = true;
#pragma unroll
                            for (int lvl = 0; lvl < 3; lvl++) {            // windows 2, 4, 8
                                const int sft = 1 << lvl;
                                const int src = (lane + sft) & 31;
                                const bool wrap = lane + sft >= 32;
                                uint32_t TL[NR];
#pragma unroll
                                for (int R = 0; R < NR; R++) TL[R] = __shfl_sync(0xFFFFFFFFu, AL[R], src);
#pragma unroll
                                for (int R = 0; R < NR; R++)     // (lanes near 31 of the last row take wrapped values: nobody reads them)
                                    AL[R] = min(AL[R], (wrap && R + 1 < NR) ? TL[R + 1 < NR ? R + 1 : R] : TL[R]);
                            }
#pragma unroll
                            for (int j = 0; j < ILP; j++) {
                                const int i = r0 + j;
                                const bool valid = i < rounds && cb + 32 * i + lane < nk;
                                const uint64_t c = cs[j];
                                const bool is_fwd = (zs[j] >> 61) & 1ull;
                                const uint32_t pos = AL[j] & 255u;                            // 32 j + lane + offset
                                const int src = (int)(pos & 31u);
                                const uint64_t z0 = shfl64(zs[j], src), z1 = shfl64(zs[j + 1], src);
                                const uint64_t zf = (pos >> 5) == (uint32_t)j ? z0 : z1;
                                const int o_read = (int)((pos - (uint32_t)(32 * j + lane)) & 7u);
                                uint32_t zq, line;
                                local_divmod(zf & ((1ull << 61) - 1), (uint32_t)T.NL, T.nl_m32, T.nl_sh, zq, line);
                                const int o_c = is_fwd ? o_read : LOCAL_W - 1 - o_read;
                                const bool f = (zf >> (is_fwd ? 63 : 62)) & 1ull;
                                const uint32_t rest = local_rest(c, o_c, m);
                                q[j] = (uint64_t)local_key_lo(zq, rest) | ((uint64_t)local_key_hi(rest, o_c, f) << 32);
                                cc[j] = c;
                                // the sector of the A line and of the B line (second choice of the table builder):
                                // two independent loads, no dependent second round trip
                                const uint64_t lb64 = (uint64_t)line * 4 + (uint64_t)(o_c & 3) - T.lo;
                                lb[j] = (uint32_t)lb64;
                                live[j] = valid && lb64 < T.n_local;
                                my_lookups += valid;
                                const uint32_t rel_b = local_alt_rel(line - T.line_lo, zq, T.line_n, false);
#if CUCLARK_LOCAL_BRANCHFREE
                                // no branch around the loads: a lane without a k-mer (tail of a part, other shard) reads sector 0
                                const uint64_t sa_ = live[j] ? (uint64_t)lb[j] : 0ull;
                                const uint64_t sb_ = live[j] ? (uint64_t)rel_b * 4 + (uint64_t)(o_c & 3) : 0ull;
                                sec[j] = load_sector_line(T.buckets + 2 * sa_);
                                secb[j] = load_sector_line(T.buckets + 2 * sb_);
#else
                                if (live[j]) {
                                    sec[j] = load_sector_line(T.buckets + 2 * (uint64_t)lb[j]);
                                    secb[j] = load_sector_line(T.buckets + 2 * ((uint64_t)rel_b * 4 + (uint64_t)(o_c & 3)));
                                }
#endif
                            }
#pragma unroll
                            for (int j = 0; j < ILP; j++) {
                                uint32_t label = NO_LABEL;
#if CUCLARK_LOCAL_BRANCHFREE
                                {
                                    const uint32_t la_ = match_sector<LAYOUT_LOCAL>(sec[j], q[j]);
                                    const uint32_t lb_ = match_sector<LAYOUT_LOCAL>(secb[j], q[j] | ((uint64_t)LOCAL_ALT_BIT << 32));
                                    label = la_ == NO_LABEL ? lb_ : la_;
                                    // both candidate sectors full at build time: the k-mer may be in the overflow table
                                    // (0.3 % of the entries after the builder's rescue pass): one warp-uniform branch
                                    const bool ask_ovf = live[j] && label == NO_LABEL && sector_overflowed(sec[j]) && sector_overflowed(secb[j]);
                                    if (__any_sync(0xFFFFFFFFu, ask_ovf)) { if (ask_ovf) label = ovf_lookup(T, cc[j]); }
                                    if (!live[j] || label >= p.n_targets) label = NO_LABEL;
                                }
#else
                                if (live[j]) {
                                    label = match_sector<LAYOUT_LOCAL>(sec[j], q[j]);
                                    if (label == NO_LABEL) label = match_sector<LAYOUT_LOCAL>(secb[j], q[j] | ((uint64_t)LOCAL_ALT_BIT << 32));
                                    // both candidate sectors full at build time: the k-mer may be in the overflow table (rare)
                                    if (label == NO_LABEL && sector_overflowed(sec[j]) && sector_overflowed(secb[j])) label = ovf_lookup(T, cc[j]);
                                    if (label >= p.n_targets) label = NO_LABEL;
                                }
#endif
                                hits.add(label, tkey, tcnt, lane);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < ILP; j++) {
                                const int i = r0 + j;
                                const uint64_t hi = shfl64(W, i & 31), lo = shfl64(W, (i + 1) & 31);
                                const uint64_t x = window64(hi, lo, 2 * lane) >> kshift;
                                const bool valid = i < rounds && cb + 32 * i + lane < nk;
                                uint64_t b;
                                divmod_M(canonical(x, k), T.M, T.magic, q[j], b);
                                const uint64_t lb64 = b - T.lo;
                                lb[j] = (uint32_t)lb64;              // n_local < 2^32 (checked at build)
                                live[j] = valid && lb64 < T.n_local;
                                my_lookups += valid;
                                if (live[j]) sec[j] = load_sector(T.buckets + 2 * (uint64_t)lb[j]);
                            }
#pragma unroll
                            for (int j = 0; j < ILP; j++) {
                                uint32_t label = NO_LABEL;
                                if (live[j]) {
                                    label = match_sector<LAYOUT>(sec[j], q[j]);
                                    if (label == NO_LABEL && sector_overflowed(sec[j]))
                                        label = ovf_lookup(T, q[j] * T.M + ((uint64_t)lb[j] + T.lo));
                                    if (label >= p.n_targets) label = NO_LABEL;
                                }
                                hits.add(label, tkey, tcnt, lane);
                            }
                        }
                    }
                }
                first_part = false;
            }

        // ---- per-read result --------------------------------------------------
        __syncwarp();
        hits.finish<ROWS>(p.out, sptr[32] + sptr[34] - 1, tkey, tcnt, s_row[wib], lane);
        }   // reads of the chunk
        flush_lookups();
    }       // chunks
}

// Exact fallback for reads that hit more than TSLOTS distinct targets: one block
// per read, dense per-target counters in global scratch (the reference's own
// scheme, src/CuClarkDB.cu:1064-1074, 1156-1243), ascending scan by thread 0.
template <int LAYOUT>
__global__ void __launch_bounds__(256) k_classify_dense(const ClassifyParams p, uint32_t* hist_all) {
    uint32_t* hist = hist_all + (size_t)blockIdx.x * p.n_targets;
    const uint32_t n_list = min(p.out.counters[COUNTER_DENSE], p.out.dense_cap);
    const int k = p.t.k;
    for (uint32_t li = blockIdx.x; li < n_list; li += gridDim.x) {
        const uint32_t read = p.out.dense_list[li];
        uint32_t pos = p.reads_ptr[read];
        const uint32_t end = p.reads_ptr[read + 1];
        while (pos < end) {
            uint32_t L;
            const uint32_t ncont = part_extent(p.cont[pos], pos + 1, end, L);      // clamped as in k_classify
            const uint16_t* c = p.cont + pos + 1;
            pos += 1 + ncont;
            const int nk = (int)L - k + 1;
            for (int w = threadIdx.x; w < nk; w += blockDim.x) {
                uint64_t x = 0;
                for (int j = 0; j < k; j++) {
                    const int qn = w + j;
                    x = (x << 2) | ((c[qn >> 3] >> (2 * (7 - (qn & 7)))) & 3u);
                }
                const uint64_t rc = revcomp2(x, k);
                const uint32_t label = table_lookup<LAYOUT>(p.t, x <= rc ? x : rc, x <= rc);
                if (label < p.n_targets) atomicAdd(&hist[label], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) dense_emit(hist, p.n_targets, read, p.out);
        __syncthreads();
    }
}

// mergeKernel + resultKernel (src/CuClarkDB.cu:1321-1471) for n_parts shards:
// n_parts-way merge of ascending sparse rows, summing equal targets, with the
// top-2 scan folded in. One thread per read.
__global__ void k_merge_rows(const uint16_t* __restrict__ parts, int n_parts, uint32_t n_reads, int row_pairs,
                             uint16_t* rows_out, uint16_t* final5, uint32_t* counters) {
    const uint32_t read = blockIdx.x * blockDim.x + threadIdx.x;
    if (read >= n_reads) return;
    const int pitch = 2 * row_pairs + 2;
    const size_t part_stride = (size_t)n_reads * pitch;
    int idx[16];
    for (int g = 0; g < n_parts; g++) idx[g] = 0;
    uint16_t best = 0, sbest = 0, ib = 0, isb = 0, sum = 0;
    uint32_t n = 0;
    uint16_t* out = rows_out ? rows_out + (size_t)read * pitch : nullptr;
    // an input row that was cut at row_pairs (its shard saw more targets than a row holds) makes the merged
    // result of this read inexact: it is counted in truncated_rows, which callers must find zero
    bool cut = false;
    for (int g = 0; g < n_parts; g++) cut = cut || parts[g * part_stride + (size_t)read * pitch] > (uint16_t)row_pairs;
    for (;;) {
        uint32_t tmin = 0x10000u;
        for (int g = 0; g < n_parts; g++) {
            const uint16_t* r = parts + g * part_stride + (size_t)read * pitch;
            const int cnt = min((int)r[0], row_pairs);
            if (idx[g] < cnt) tmin = min(tmin, (uint32_t)r[1 + 2 * idx[g]]);
        }
        if (tmin == 0x10000u) break;
        uint16_t h = 0;
        for (int g = 0; g < n_parts; g++) {
            const uint16_t* r = parts + g * part_stride + (size_t)read * pitch;
            const int cnt = min((int)r[0], row_pairs);
            if (idx[g] < cnt && r[1 + 2 * idx[g]] == tmin) { h = (uint16_t)(h + r[2 + 2 * idx[g]]); idx[g]++; }
        }
        if (h > best) { sbest = best; isb = ib; best = h; ib = (uint16_t)(tmin + 1); }
        else if (h > sbest) { sbest = h; isb = (uint16_t)(tmin + 1); }
        sum = (uint16_t)(sum + h);
        if (out && n < (uint32_t)row_pairs) { out[1 + 2 * n] = (uint16_t)tmin; out[2 + 2 * n] = h; }
        n++;
    }
    if (out) {
        out[0] = (uint16_t)n;
        for (uint32_t i = 1 + 2 * min(n, (uint32_t)row_pairs); i < (uint32_t)pitch; i++) out[i] = 0;
    }
    if (cut || (out && n > (uint32_t)row_pairs)) atomicAdd(&counters[COUNTER_TRUNC], 1u);
    if (final5) {
        uint16_t* f = final5 + (size_t)read * 5;
        f[0] = sum; f[1] = ib; f[2] = best; f[3] = isb; f[4] = sbest;
    }
}

// ---- synthetic reads, packed (device twin of synth.read_codes + oracle pack) ----
__global__ void k_synth_reads(uint32_t seed, uint32_t genome_seed, uint32_t n_targets, uint64_t genome_len,
                              uint64_t first_read, uint32_t n_reads, int read_len, int pct_random, int sub_per_10k,
                              uint32_t* reads_ptr, uint16_t* cont) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t per_read = 1 + (read_len + 7) / 8;
    if (r == 0) reads_ptr[n_reads] = n_reads * per_read;
    if (r >= n_reads) return;
    const uint64_t i = first_read + r;
    const uint64_t h1 = synth::key(synth::TAG_READ, seed, i, 0), h2 = synth::key(synth::TAG_READ, seed, i, 1);
    const bool is_random = (h1 % 100) < (uint64_t)pct_random;
    const uint32_t target = (uint32_t)((h1 >> 8) % n_targets);
    const bool is_rc = (h1 >> 40) & 1;
    const uint64_t pos = h2 % (genome_len - read_len + 1);
    uint16_t* c = cont + (size_t)r * per_read;
    reads_ptr[r] = r * per_read;
    c[0] = (uint16_t)read_len;
    uint32_t word = 0, fill = 0, ci = 1;
    for (int j = 0; j < read_len; j++) {
        uint32_t code;
        if (is_random) {
            code = (uint32_t)(synth::key(synth::TAG_RBASE, seed, i, (uint64_t)(j >> 5)) >> (2 * (j & 31))) & 3u;
        } else if (is_rc) {
            code = 3u - synth::genome_base(genome_seed, target, pos + (read_len - 1 - j));
        } else {
            code = synth::genome_base(genome_seed, target, pos + j);
        }
        if (sub_per_10k) {
            const uint64_t sh = synth::key(synth::TAG_SUB, seed, i, (uint64_t)j);
            if ((sh % 10000) < (uint64_t)sub_per_10k) code = (code + 1 + (uint32_t)((sh >> 20) % 3)) & 3u;
        }
        word = (word << 2) | (3u - code);        // packed code is the complement code
        if (++fill == 8) { c[ci++] = (uint16_t)word; word = 0; fill = 0; }
    }
    if (fill) c[ci] = (uint16_t)(word << (2 * (8 - fill)));
}

// Same reads as k_synth_reads, as 4-line FASTQ text with fixed-width records:
// "@r%09llu\n" + bases + "\n+\n" + 'I' * read_len + "\n"  (16 + 2*read_len bytes each).
// mate = 0: the single-end read; 1 / 2: the mates of a pair (BASELINE configs[2]): mate 1 is that same read, mate 2 lies
// 2 x read_len further along the same target (clamped to its end) on the OPPOSITE strand, with its own substitutions.
__global__ void k_synth_fastq(uint32_t seed, uint32_t genome_seed, uint32_t n_targets, uint64_t genome_len,
                              uint64_t first_read, uint32_t n_reads, int read_len, int pct_random, int sub_per_10k,
                              uint8_t* text, int mate) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const size_t rec = 16 + 2 * (size_t)read_len;
    uint8_t* o = text + (size_t)r * rec;
    const uint64_t i = first_read + r;
    const uint64_t h1 = synth::key(synth::TAG_READ, seed, i, 0), h2 = synth::key(synth::TAG_READ, seed, i, 1);
    const bool is_random = (h1 % 100) < (uint64_t)pct_random;
    const uint32_t target = (uint32_t)((h1 >> 8) % n_targets);
    bool is_rc = (h1 >> 40) & 1;
    uint64_t pos = h2 % (genome_len - read_len + 1);
    const uint64_t jo = mate == 2 ? 1024 : 0;            // mate 2 draws its random bases / substitutions from other counters
    if (mate == 2) { pos = min(pos + 2 * (uint64_t)read_len, genome_len - read_len); is_rc = !is_rc; }
    o[0] = '@'; o[1] = 'r';
    uint64_t v = i % 1000000000ull;
    for (int d = 8; d >= 0; d--) { o[2 + d] = (uint8_t)('0' + v % 10); v /= 10; }
    o[11] = '\n';
    uint8_t* sq = o + 12;
    for (int j = 0; j < read_len; j++) {
        uint32_t code;
        if (is_random) code = (uint32_t)(synth::key(synth::TAG_RBASE, seed, i, jo + (uint64_t)(j >> 5)) >> (2 * (j & 31))) & 3u;
        else if (is_rc) code = 3u - synth::genome_base(genome_seed, target, pos + (read_len - 1 - j));
        else code = synth::genome_base(genome_seed, target, pos + j);
        if (sub_per_10k) {
            const uint64_t sh = synth::key(synth::TAG_SUB, seed, i, jo + (uint64_t)j);
            if ((sh % 10000) < (uint64_t)sub_per_10k) code = (code + 1 + (uint32_t)((sh >> 20) % 3)) & 3u;
        }
        sq[j] = (uint8_t)"ACGT"[code];
    }
    uint8_t* t = sq + read_len;
    t[0] = '\n'; t[1] = '+'; t[2] = '\n';
    for (int j = 0; j < read_len; j++) t[3 + j] = 'I';
    t[3 + read_len] = '\n';
}

// ---- random-sector gather: the roofline denominator ----------------------------
template <int BYTES, int ILP>
__global__ void __launch_bounds__(256) k_gather(const uint4* __restrict__ base, uint64_t n_units, uint64_t n_probes,
                                                uint64_t salt, uint32_t* sink) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t stride = gridDim.x * (uint64_t)blockDim.x;
    uint32_t acc = 0;
    // n_probes is rounded up to a multiple of stride*ILP by the host, so every
    // load below is unconditional and all ILP of them are in flight together
    for (uint64_t i = tid; i < n_probes; i += stride * ILP) {
        uint4 v[ILP][BYTES / 16];
        const uint4* ptr[ILP];
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            const uint64_t h = synth::mix64((i + j * stride) ^ salt);
            ptr[j] = base + __umul64hi(h, n_units) * (BYTES / 16);      // uniform in [0, n_units)
        }
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            if (BYTES == 32) {
                asm("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(v[j][0].x), "=r"(v[j][0].y), "=r"(v[j][0].z), "=r"(v[j][0].w),
                      "=r"(v[j][1].x), "=r"(v[j][1].y), "=r"(v[j][1].z), "=r"(v[j][1].w)
                    : "l"(ptr[j]));
            } else {
#pragma unroll
                for (int q = 0; q < BYTES / 16; q++)
                    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                        : "=r"(v[j][q].x), "=r"(v[j][q].y), "=r"(v[j][q].z), "=r"(v[j][q].w)
                        : "l"(ptr[j] + q));
            }
        }
#pragma unroll
        for (int j = 0; j < ILP; j++)
#pragma unroll
            for (int q = 0; q < BYTES / 16; q++) acc ^= v[j][q].x ^ v[j][q].y ^ v[j][q].z ^ v[j][q].w;
    }
    if (acc == 0x12345678u) *sink = acc;      // keep the loads alive
}

template <int BYTES>
void gather_dispatch(int ilp, int blocks, const uint4* base, uint64_t n_units, uint64_t n_probes, uint64_t salt,
                     uint32_t* sink, cudaStream_t st) {
    switch (ilp) {
        case 1: k_gather<BYTES, 1><<<blocks, 256, 0, st>>>(base, n_units, n_probes, salt, sink); break;
        case 2: k_gather<BYTES, 2><<<blocks, 256, 0, st>>>(base, n_units, n_probes, salt, sink); break;
        case 8: k_gather<BYTES, 8><<<blocks, 256, 0, st>>>(base, n_units, n_probes, salt, sink); break;
        default: k_gather<BYTES, 4><<<blocks, 256, 0, st>>>(base, n_units, n_probes, salt, sink); break;
    }
}

}  // namespace

template <typename K>
static int blocks_per_sm(K kernel) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, WARPS_PER_BLOCK * 32, 0) != cudaSuccess || n < 1) n = 1;
    return n;
}

int classify_launch(cuclark_db* db, const Scratch& sc, const uint32_t* d_ptr, const uint16_t* d_cont, size_t n_reads,
                    uint16_t* d_final, uint16_t* d_rows, cudaStream_t st) {
    if (!db->d_table) { set_error("no database loaded"); return CUCLARK_ERR_STATE; }
    if (n_reads > 0xFFFFFFF0ull) { set_error("too many reads in one call"); return CUCLARK_ERR_ARG; }
    if (db->row_pairs > MAX_ROW_PAIRS) { set_error("row_pairs > %d", MAX_ROW_PAIRS); return CUCLARK_ERR_ARG; }
    ClassifyParams p;
    p.t = db->view;
    p.reads_ptr = d_ptr; p.cont = d_cont; p.n_reads = (uint32_t)n_reads;
    p.n_targets = (uint32_t)db->cfg.n_targets;
    p.out.final5 = d_final; p.out.rows = d_rows; p.out.row_pairs = db->row_pairs;
    p.out.counters = sc.d_counters; p.out.dense_list = sc.d_dense_list; p.out.dense_cap = sc.dense_cap;
    CK(cudaMemsetAsync(sc.d_counters, 0, N_COUNTERS * sizeof(uint32_t), st));
    if (n_reads == 0) return CUCLARK_OK;
    const int layout = db->view.layout;
    const bool narrow = layout == LAYOUT_NARROW, local = layout == LAYOUT_LOCAL;
    // persistent grid: SM count x resident blocks per SM (one wave), capped by the work
    const int variant = (narrow ? 0 : local ? 4 : 2) + (d_rows ? 1 : 0);
    void (*kern)(const ClassifyParams) =
        d_rows ? (narrow ? k_classify<LAYOUT_NARROW, true> : local ? k_classify<LAYOUT_LOCAL, true> : k_classify<LAYOUT_WIDE, true>)
               : (narrow ? k_classify<LAYOUT_NARROW, false> : local ? k_classify<LAYOUT_LOCAL, false> : k_classify<LAYOUT_WIDE, false>);
    if (!db->classify_blocks_per_sm[variant]) db->classify_blocks_per_sm[variant] = blocks_per_sm(kern);
    const int blocks_needed = (int)((n_reads + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    const int blocks = std::min(blocks_needed, db->sm_count * db->classify_blocks_per_sm[variant]);
    kern<<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(p);
    CK(cudaGetLastError());
    // exact fallback; exits at once when the list is empty. A scratch with its own histogram (batches, pipeline slots)
    // makes the call independent of every other stream; the handle's shared one is chained across streams.
    auto dense = [&](uint32_t* hist) {
        if (narrow) k_classify_dense<LAYOUT_NARROW><<<db->dense_blocks, 256, 0, st>>>(p, hist);
        else if (local) k_classify_dense<LAYOUT_LOCAL><<<db->dense_blocks, 256, 0, st>>>(p, hist);
        else k_classify_dense<LAYOUT_WIDE><<<db->dense_blocks, 256, 0, st>>>(p, hist);
    };
    if (sc.d_dense_hist) {
        dense(sc.d_dense_hist);
        CK(cudaGetLastError());
    } else {
        std::lock_guard<std::mutex> dense_guard(db->dense_mu);
        CK(cudaStreamWaitEvent(st, db->dense_chain, 0));
        dense(db->d_dense_hist);
        CK(cudaGetLastError());
        CK(cudaEventRecord(db->dense_chain, st));
    }
    count_launches(2);
    return CUCLARK_OK;
}

int merge_rows_launch(cuclark_db* db, const uint16_t* d_parts, int n_parts, size_t n_reads, uint16_t* d_rows_out,
                      uint16_t* d_final, cudaStream_t st) {
    if (n_parts < 1 || n_parts > 16) { set_error("n_parts must be 1..16"); return CUCLARK_ERR_ARG; }
    if (n_reads == 0) return CUCLARK_OK;
    k_merge_rows<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(d_parts, n_parts, (uint32_t)n_reads, db->row_pairs,
                                                                   d_rows_out, d_final, db->scratch.d_counters);
    CK(cudaGetLastError());
    count_launches(1);
    return CUCLARK_OK;
}

int synth_reads_launch(uint32_t seed, uint32_t genome_seed, uint32_t n_targets, uint64_t genome_len,
                       uint64_t first_read, size_t n_reads, int read_len, int pct_random, int sub_per_10k,
                       uint32_t* d_ptr, uint16_t* d_cont, cudaStream_t st) {
    if (read_len < 1 || read_len > 65535 || genome_len < (uint64_t)read_len) { set_error("bad read_len"); return CUCLARK_ERR_ARG; }
    const uint64_t per_read = 1 + (read_len + 7) / 8;
    if (n_reads * per_read > 0xFFFFFFFFull) { set_error("batch exceeds 2^32 containers"); return CUCLARK_ERR_ARG; }
    k_synth_reads<<<(unsigned)((n_reads + 256) / 256), 256, 0, st>>>(seed, genome_seed, n_targets, genome_len, first_read,
                                                                     (uint32_t)n_reads, read_len, pct_random, sub_per_10k,
                                                                     d_ptr, d_cont);
    CK(cudaGetLastError());
    return CUCLARK_OK;
}

int synth_fastq_launch(uint32_t seed, uint32_t genome_seed, uint32_t n_targets, uint64_t genome_len, uint64_t first_read,
                       size_t n_reads, int read_len, int pct_random, int sub_per_10k, uint8_t* d_text, cudaStream_t st, int mate) {
    if (mate < 0 || mate > 2 || (mate && genome_len < 3ull * (uint64_t)read_len)) { set_error("bad mate / genome too short for pairs"); return CUCLARK_ERR_ARG; }
    if (read_len < 1 || read_len > 65535 || genome_len < (uint64_t)read_len) { set_error("bad read_len"); return CUCLARK_ERR_ARG; }
    if (n_reads > 0xFFFFFFF0ull) { set_error("too many reads"); return CUCLARK_ERR_ARG; }
    if (n_reads)
        k_synth_fastq<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(seed, genome_seed, n_targets, genome_len, first_read,
                                                                         (uint32_t)n_reads, read_len, pct_random, sub_per_10k, d_text, mate);
    CK(cudaGetLastError());
    return CUCLARK_OK;
}

int gather_bench_launch(cuclark_db* db, uint64_t n_probes, int bytes_per_probe, int ilp, int iters, double* ms_out) {
    if (!db->d_table) { set_error("no database loaded"); return CUCLARK_ERR_STATE; }
    if (bytes_per_probe != 32 && bytes_per_probe != 64 && bytes_per_probe != 128) { set_error("bytes_per_probe must be 32, 64 or 128"); return CUCLARK_ERR_ARG; }
    const uint64_t n_units = db->view.n_local * 32 / bytes_per_probe;
    if (n_units == 0) { set_error("table too small"); return CUCLARK_ERR_ARG; }
    const int blocks = db->sm_count * 8;
    const uint64_t quantum = (uint64_t)blocks * 256 * (ilp == 1 || ilp == 2 || ilp == 8 ? ilp : 4);
    n_probes = (n_probes + quantum - 1) / quantum * quantum;
    cudaStream_t st = db->stream;
    for (int it = -1; it < iters; it++) {
        if (it == 0) CK(cudaEventRecord(db->ev0, st));
        const uint64_t salt = 0x5bd1e995ull * (uint64_t)(it + 2);
        if (bytes_per_probe == 32) gather_dispatch<32>(ilp, blocks, db->d_table, n_units, n_probes, salt, db->scratch.d_counters + 6, st);
        else if (bytes_per_probe == 64) gather_dispatch<64>(ilp, blocks, db->d_table, n_units, n_probes, salt, db->scratch.d_counters + 6, st);
        else gather_dispatch<128>(ilp, blocks, db->d_table, n_units, n_probes, salt, db->scratch.d_counters + 6, st);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(db->ev1, st));
    CK(cudaEventSynchronize(db->ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, db->ev0, db->ev1));
    *ms_out = ms / iters;
    return CUCLARK_OK;
}

}  // namespace cuclark
