// cuclark_b200 — device table construction ("re-bucketing").
//
// Replaces CuClarkDB::read / swapDbParts (reference src/CuClarkDB.cu:462-858):
// instead of per-part prefix-sum pointers + key and label arrays copied as they
// lie in the files, the (bucket r, quotient q, label) triples of <base>.sz/.ky/.lb
// (format: src/hashTable_hh.hh:591-663) are turned back into canonical k-mers
// c = q*HTSIZE + r and scattered ON THE DEVICE into 32-byte sector buckets
// (layout in common.cuh). The whole table stays resident in HBM: no swap cycles,
// no 32-bit bucket pointers.
#include <fcntl.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "internal.h"
#include "synth.cuh"

namespace cuclark {

namespace {

constexpr uint32_t ERR_OVF_LIST = 1, ERR_COUNTER = 2, ERR_NO_SLOT = 4;
constexpr double OVF_LOAD = 0.75;                  // mean entries per 3-slot overflow bucket
constexpr double OVF_LOAD_LOCAL = 1.2;             // LOCAL spills ~13% of the entries: a denser overflow table
constexpr int CHUNK_THREADS = 1024;
constexpr uint32_t CHUNK_BLOCKS = 8192;            // 8.4M reference buckets per chunk
constexpr int LOAD_IO_THREADS = 12;                // host threads filling a staging set of the file loader

struct BuildCtx {
    uint4* table;
    uint32_t* cnt8;        // one byte per local bucket, packed 4 per word
    uint64_t M, magic, lo, n_local, htsize;
    uint64_t* ovf_c;       // k-mers that did not fit their home bucket
    uint16_t* ovf_l;
    uint32_t ovf_cap;
    uint32_t* flags;       // [0] overflow count, [1] error bits
    unsigned long long* inserted;   // entries homed in this shard
    uint4* ovf;            // overflow table (allocated after phase 1)
    uint32_t* ovf_cnt8;
    uint64_t n_ovf;
    uint64_t NL;           // LOCAL: lines (M = 4 NL)
    int k, layout;
};

// global home bucket b and in-bucket key q of the canonical k-mer c
template <int LAYOUT>
__device__ __forceinline__ void home_of(uint64_t c, uint64_t M, uint64_t magic, uint64_t NL, int k, uint64_t& q, uint64_t& b) {
    if (LAYOUT == LAYOUT_LOCAL) local_locate(c, k, NL, b, q);
    else divmod_M(c, M, magic, q, b);
}
__device__ __forceinline__ void home_of_rt(int layout, uint64_t c, uint64_t M, uint64_t magic, uint64_t NL, int k, uint64_t& q, uint64_t& b) {
    if (layout == LAYOUT_LOCAL) local_locate(c, k, NL, b, q);
    else divmod_M(c, M, magic, q, b);
}

template <int LAYOUT>
__device__ __forceinline__ void write_slot(uint4* table, uint64_t lb, uint32_t slot, uint64_t q, uint32_t label) {
    uint32_t* w = reinterpret_cast<uint32_t*>(table + 2 * lb);
    if (LAYOUT == LAYOUT_NARROW) {
        w[slot] = (uint32_t)q;
        atomicOr(&w[5 + (slot >> 1)], label << (16 * (slot & 1)));
    } else if (LAYOUT == LAYOUT_LOCAL) {
        w[slot] = (uint32_t)q;
        atomicOr(&w[4 + (slot >> 1)], label << (16 * (slot & 1)));
        atomicAnd(&w[6], ~(0xFFu << (8 * slot)) | ((uint32_t)(q >> 32) << (8 * slot)));   // byte was 0xFF
    } else {
        reinterpret_cast<uint64_t*>(w)[slot] = q;
        atomicOr(&w[6 + (slot >> 1)], label << (16 * (slot & 1)));
    }
}

__device__ __forceinline__ uint32_t claim_slot(uint32_t* cnt8, uint32_t* flags, uint64_t lb) {
    const uint32_t sh = 8 * (uint32_t)(lb & 3);
    const uint32_t old = (atomicAdd(&cnt8[lb >> 2], 1u << sh) >> sh) & 0xFFu;
    if (old == 255u) atomicOr(&flags[1], ERR_COUNTER);
    return old;
}

template <int LAYOUT>
__device__ __forceinline__ bool insert_home(const BuildCtx& x, uint64_t c, uint32_t label) {
    constexpr uint32_t SLOTS = LAYOUT == LAYOUT_NARROW ? NARROW_SLOTS : LAYOUT == LAYOUT_LOCAL ? LOCAL_SLOTS : WIDE_SLOTS;
    uint64_t q, b;
    if (LAYOUT == LAYOUT_LOCAL) {
        // two candidate sectors (A line, B line): take the emptier one, the other if that one fills up
        // under our feet; the overflow table only when both are full
        uint64_t sa, sb, sr, qr;
        const bool tie = local_locate_both(c, x.k, x.NL, sa, q, sr, qr);
        if (tie) {
            // two possible homes (common.cuh, "TIES"): lives in the overflow table of every shard that holds one of
            // them (k_place_spills flags their sectors); counted by the shard of the first home
            const bool mine = sa - x.lo < x.n_local;
            if (mine || sr - x.lo < x.n_local) {
                const uint32_t i = atomicAdd(&x.flags[0], 1u);
                if (i < x.ovf_cap) { x.ovf_c[i] = c; x.ovf_l[i] = (uint16_t)label; }
                else atomicOr(&x.flags[1], ERR_OVF_LIST);
            }
            return mine;
        }
        sb = local_alt_sector(sa, q, x.lo >> 2, (uint32_t)(x.n_local >> 2));
        const uint64_t la = sa - x.lo, lbb = sb - x.lo;
        if (la >= x.n_local) return false;           // homed in another shard
        const uint32_t ca = (__ldcg(&x.cnt8[la >> 2]) >> (8 * (uint32_t)(la & 3))) & 0xFFu;
        const uint32_t cb = (__ldcg(&x.cnt8[lbb >> 2]) >> (8 * (uint32_t)(lbb & 3))) & 0xFFu;
        const bool b_first = cb < ca && lbb != la;
        for (int t = 0; t < 2; t++) {
            const bool use_b = (t == 0) == b_first;
            if (use_b && lbb == la) continue;
            const uint64_t l = use_b ? lbb : la;
            if (((__ldcg(&x.cnt8[l >> 2]) >> (8 * (uint32_t)(l & 3))) & 0xFFu) >= SLOTS) continue;   // full: do not wrap the byte
            const uint32_t slot = claim_slot(x.cnt8, x.flags, l);
            if (slot < SLOTS) {
                write_slot<LAYOUT>(x.table, l, slot, use_b ? q | ((uint64_t)LOCAL_ALT_BIT << 32) : q, label);
                return true;
            }
        }
        const uint32_t i = atomicAdd(&x.flags[0], 1u);
        if (i < x.ovf_cap) { x.ovf_c[i] = c; x.ovf_l[i] = (uint16_t)label; }
        else atomicOr(&x.flags[1], ERR_OVF_LIST);
        return true;
    }
    home_of<LAYOUT>(c, x.M, x.magic, x.NL, x.k, q, b);
    const uint64_t lb = b - x.lo;
    if (lb >= x.n_local) return false;               // homed in another shard
    const uint32_t slot = claim_slot(x.cnt8, x.flags, lb);
    if (slot < SLOTS) {
        write_slot<LAYOUT>(x.table, lb, slot, q, label);
    } else {
        const uint32_t i = atomicAdd(&x.flags[0], 1u);
        if (i < x.ovf_cap) { x.ovf_c[i] = c; x.ovf_l[i] = (uint16_t)label; }
        else atomicOr(&x.flags[1], ERR_OVF_LIST);
    }
    return true;
}

__global__ void k_init_table(uint4* table, uint64_t n_local, int layout) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;   // one uint4 each
    if (i >= 2 * n_local) return;
    uint4 v;
    if (layout == LAYOUT_NARROW) {
        v = (i & 1) ? make_uint4(0xFFFFFFFFu, 0u, 0u, 0u) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    } else if (layout == LAYOUT_LOCAL) {
        v = (i & 1) ? make_uint4(0u, 0u, 0xFFFFFFFFu, 0u) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    } else {
        v = (i & 1) ? make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    }
    table[i] = v;
}

// One thread per reference bucket of the chunk; block-wide exclusive scan of
// the bucket sizes gives each bucket its offset into the chunk's keys/labels.
template <int LAYOUT>
__global__ void __launch_bounds__(CHUNK_THREADS) k_insert_chunk(BuildCtx x, const uint8_t* __restrict__ sz,
                                                                const uint8_t* __restrict__ keep,
                                                                const uint64_t* __restrict__ coarse, uint64_t r0,
                                                                uint32_t nb, const void* __restrict__ keys,
                                                                const uint16_t* __restrict__ labels, int key_bytes) {
    __shared__ uint32_t warp_sum[CHUNK_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t r = blockIdx.x * CHUNK_THREADS + tid;
    const uint32_t s = r < nb ? sz[r] : 0;
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t v = warp_sum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if (lane >= o) v += u;
        }
        warp_sum[lane] = v;
    }
    __syncthreads();
    const uint64_t off = coarse[blockIdx.x] + (wid ? warp_sum[wid - 1] : 0) + (incl - s);
    if (s == 0 || (keep && !keep[r])) return;
    unsigned long long mine = 0;
    for (uint32_t i = 0; i < s; i++) {
        uint64_t q;
        if (key_bytes == 4) q = static_cast<const uint32_t*>(keys)[off + i];
        else if (key_bytes == 2) q = static_cast<const uint16_t*>(keys)[off + i];
        else q = static_cast<const uint64_t*>(keys)[off + i];
        const uint64_t c = q * x.htsize + (r0 + r);
        mine += insert_home<LAYOUT>(x, c, labels[off + i]);
    }
    if (mine) atomicAdd(x.inserted, mine);
}

// Entries whose home bucket was full: flag the home bucket and put the full
// k-mer into the overflow table (linear probing over WIDE buckets).
__global__ void k_place_spills(BuildCtx x, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t c = x.ovf_c[i];
    if (c == OVF_EMPTY) return;                      // housed in the lines after all (k_local_rescue)
    const uint32_t label = x.ovf_l[i];
    uint64_t q, b;
    if (x.layout == LAYOUT_LOCAL) {
        // both candidate sectors were full, or a tie k-mer: flag the A and B sector of its home(s) held by this shard
        uint64_t sec[2], key[2];
        const bool tie = local_locate_both(c, x.k, x.NL, sec[0], key[0], sec[1], key[1]);
        for (int h = 0; h < (tie ? 2 : 1); h++) {
            if (sec[h] - x.lo >= x.n_local) continue;
            const uint64_t sb = local_alt_sector(sec[h], key[h], x.lo >> 2, (uint32_t)(x.n_local >> 2));
            atomicOr(reinterpret_cast<uint32_t*>(x.table + 2 * (sec[h] - x.lo)) + 7, 1u << 16);
            atomicOr(reinterpret_cast<uint32_t*>(x.table + 2 * (sb - x.lo)) + 7, 1u << 16);
        }
    } else {
        home_of_rt(x.layout, c, x.M, x.magic, x.NL, x.k, q, b);
        atomicOr(reinterpret_cast<uint32_t*>(x.table + 2 * (b - x.lo)) + 7, 1u << 16);
    }
    uint64_t ob = ovf_home(c, x.n_ovf);
    for (uint64_t n_try = 0; n_try < x.n_ovf; n_try++) {
        // the counter saturates logically at WIDE_SLOTS; do not wrap its byte
        const uint32_t sh = 8 * (uint32_t)(ob & 3);
        if (((__ldcg(&x.ovf_cnt8[ob >> 2]) >> sh) & 0xFFu) < (uint32_t)WIDE_SLOTS) {
            const uint32_t slot = claim_slot(x.ovf_cnt8, x.flags, ob);
            if (slot < (uint32_t)WIDE_SLOTS) { write_slot<LAYOUT_WIDE>(x.ovf, ob, slot, c, label); return; }
        }
        if (++ob == x.n_ovf) ob = 0;
    }
    atomicOr(&x.flags[1], ERR_NO_SLOT);
}

// ---- LOCAL: rescue of entries that found both candidate sectors full ------------------------------------
// The builder is insert-only ("the emptier of the two sectors"), which left 2.9 % of the entries of the bacterial-scale
// table in the overflow table — and an overflow probe is a DEPENDENT second memory round trip that 30-40 % of the
// kernel's rows paid for one or two of their lanes (ncu, profiles/r02_classify_local_sass_profile.md). An entry whose
// sectors A and B are full can usually still be housed: one of the eight entries sitting there moves to ITS other
// candidate sector (depth-1 cuckoo displacement; at 2.7 entries per 4-slot sector nearly always one of the eight
// alternatives has room). A line and B line of a minimizer lie in the same block of LOCAL_ALT_BLOCK lines, and so do
// the alternatives of everything stored in them: one lock per block serialises the moves, blocks proceed in parallel.
// Tie k-mers (two possible homes) stay in the overflow table.
__device__ __forceinline__ int local_free_slot(volatile uint32_t* w) {
    const uint32_t hi = w[6];
#pragma unroll
    for (int s = 0; s < LOCAL_SLOTS; s++) if (((hi >> (8 * s)) & 0xFFu) == 0xFFu) return s;
    return -1;
}
__device__ __forceinline__ void local_store_slot(volatile uint32_t* w, int s, uint64_t key, uint32_t label) {
    w[s] = (uint32_t)key;
    w[6] = (w[6] & ~(0xFFu << (8 * s))) | ((uint32_t)(key >> 32) << (8 * s));
    const int lw = 4 + (s >> 1), sh = 16 * (s & 1);
    w[lw] = (w[lw] & ~(0xFFFFu << sh)) | (label << sh);
}
__device__ __forceinline__ bool local_try_house(const BuildCtx& x, uint64_t l, uint64_t key, uint32_t label) {
    volatile uint32_t* w = reinterpret_cast<volatile uint32_t*>(x.table + 2 * l);
    const int s = local_free_slot(w);
    if (s < 0) return false;
    local_store_slot(w, s, key, label);
    return true;
}
// make room in local sector l by moving one of its entries to that entry's other candidate sector; if that one is
// full too, DEPTH more levels of the same (the entries of one sector often share their alternatives: they come in
// clumps of a few minimizers, so one level left 0.33 % of the entries in the overflow table)
template <int DEPTH>
__device__ bool local_displace(const BuildCtx& x, uint64_t l, uint64_t key, uint32_t label, uint64_t from) {
    volatile uint32_t* w = reinterpret_cast<volatile uint32_t*>(x.table + 2 * l);
    const uint32_t line_n = (uint32_t)(x.n_local >> 2);
    const uint32_t rel = (uint32_t)(l >> 2), sub = (uint32_t)(l & 3);
    for (int pass = 0; pass < (DEPTH > 0 ? 2 : 1); pass++) {                // first the cheap moves, then the deeper ones
        for (int s = 0; s < LOCAL_SLOTS; s++) {
            const uint32_t v_hi = (w[6] >> (8 * s)) & 0xFFu, v_lo = w[s];
            if (v_hi == 0xFFu) continue;                                    // (free: local_try_house would have taken it)
            const bool in_b = v_hi & LOCAL_ALT_BIT;
            const uint32_t other_rel = local_alt_rel(rel, v_lo & ((1u << LOCAL_ZQ_BITS) - 1u), line_n, in_b);
            const uint64_t other = (uint64_t)other_rel * 4 + sub;
            if (other == l || other == from) continue;
            const uint32_t lw = w[4 + (s >> 1)];
            const uint32_t v_label = (s & 1) ? (lw >> 16) : (lw & 0xFFFFu);
            const uint64_t v_key = (uint64_t)v_lo | ((uint64_t)(v_hi ^ LOCAL_ALT_BIT) << 32);
            bool moved;
            if (pass == 0) moved = local_try_house(x, other, v_key, v_label);
            else if constexpr (DEPTH > 0) moved = local_displace<DEPTH - 1>(x, other, v_key, v_label, l);
            else moved = false;
            if (!moved) continue;
            local_store_slot(w, s, key, label);
            return true;
        }
    }
    return false;
}
__global__ void k_local_rescue(BuildCtx x, uint32_t n, uint32_t* locks, unsigned long long* rescued) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t c = x.ovf_c[i];
    const uint32_t label = x.ovf_l[i];
    uint64_t sa, key, sr, kr;
    if (local_locate_both(c, x.k, x.NL, sa, key, sr, kr)) return;           // tie: both homes flagged, lives in overflow
    const uint64_t la = sa - x.lo;
    if (la >= x.n_local) return;
    const uint64_t lb = local_alt_sector(sa, key, x.lo >> 2, (uint32_t)(x.n_local >> 2)) - x.lo;
    const uint64_t key_b = key | ((uint64_t)LOCAL_ALT_BIT << 32);
    uint32_t* lock = locks + ((la >> 2) / LOCAL_ALT_BLOCK);
    bool done = false, housed = false;
    while (!done) {
        if (atomicCAS(lock, 0u, 1u) == 0u) {
            __threadfence();
            housed = local_try_house(x, la, key, label) || (lb != la && local_try_house(x, lb, key_b, label)) ||
                     local_displace<2>(x, la, key, label, ~0ull) || (lb != la && local_displace<2>(x, lb, key_b, label, ~0ull));
            __threadfence();
            atomicExch(lock, 0u);
            done = true;
        }
    }
    if (housed) { x.ovf_c[i] = OVF_EMPTY; atomicAdd(rescued, 1ull); }
}

__global__ void k_table_stats(const uint4* table, uint64_t n_local, unsigned long long* out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    bool spill = false;
    if (i < n_local) spill = ((table[2 * i + 1].w >> 16) & 1u) != 0;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, spill);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

// ---- synthetic database (bench.py) -------------------------------------------
// Every overlapping k-mer of every target (full variant, src/CuCLARK_hh.hh:896-975).
template <int LAYOUT>
__global__ void k_synth_insert_full(BuildCtx x, uint32_t seed, uint32_t n_targets, uint64_t genome_len, int k,
                                    uint64_t runs_per_target) {
    constexpr int RUN = 64;
    const uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= runs_per_target * n_targets) return;
    const uint32_t t = (uint32_t)(g / runs_per_target);
    const uint64_t p0 = (g % runs_per_target) * RUN;
    const uint64_t nk = genome_len - k + 1;
    const uint64_t mask = (~0ull) >> (64 - 2 * k);
    uint64_t R = 0, word = 0;
    unsigned long long mine = 0;
    for (uint64_t p = p0; p < p0 + RUN + k - 1 && p < genome_len; p++) {
        if (p == p0 || (p & 31) == 0) word = synth::genome_word(seed, t, p >> 5);
        const uint32_t code = (uint32_t)(word >> (2 * (p & 31))) & 3u;
        R = ((R << 2) | (3u - code)) & mask;
        if (p >= p0 + k - 1 && p - (k - 1) < nk) mine += insert_home<LAYOUT>(x, canonical(R, k), t);
    }
    if (mine) atomicAdd(x.inserted, mine);
}

// Every gap-th non-overlapping k-mer (light variant, src/CuCLARK_hh.hh:705-767).
template <int LAYOUT>
__global__ void k_synth_insert_light(BuildCtx x, uint32_t seed, uint32_t n_targets, uint64_t genome_len, int k,
                                     int gap, uint64_t per_target) {
    const uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= per_target * n_targets) return;
    const uint32_t t = (uint32_t)(g / per_target);
    const uint64_t start = (g % per_target) * (uint64_t)gap * k;
    uint64_t R = 0;
    for (int j = 0; j < k; j++) R = (R << 2) | (3u - synth::genome_base(seed, t, start + j));
    if (insert_home<LAYOUT>(x, canonical(R, k), t)) atomicAdd(x.inserted, 1ull);
}

// RemoveCommon (src/HashTableStorage_hh.hh:242-292) for the synthetic builder:
// a k-mer inserted more than once is kept once if all copies carry the same
// label and removed entirely otherwise. Copies of one k-mer sit in its home
// bucket and/or along its overflow probe sequence. Two read-only passes compute
// a verdict per slot (no races), a third applies them.
template <int LAYOUT>
struct DedupeView {
    uint4* table; uint4* ovf; uint64_t M, lo, n_local, n_ovf;
    uint8_t* del_main; uint8_t* del_ovf;
    uint64_t NL; int k;
};

template <int LAYOUT> struct SlotsOf { static constexpr int n = LAYOUT == LAYOUT_NARROW ? NARROW_SLOTS : LAYOUT == LAYOUT_LOCAL ? LOCAL_SLOTS : WIDE_SLOTS; };

// key, label and occupancy of slot s of a home bucket
template <int LAYOUT>
__device__ __forceinline__ bool read_slot(const uint32_t* w, int s, uint64_t& key, uint32_t& label) {
    if (LAYOUT == LAYOUT_NARROW) {
        key = w[s];
        const uint32_t lw = w[5 + (s >> 1)];
        label = (s & 1) ? (lw >> 16) : (lw & 0xFFFFu);
        return w[s] != 0xFFFFFFFFu;
    } else if (LAYOUT == LAYOUT_LOCAL) {
        const uint32_t hi = (w[6] >> (8 * s)) & 0xFFu;
        key = (uint64_t)w[s] | ((uint64_t)hi << 32);
        const uint32_t lw = w[4 + (s >> 1)];
        label = (s & 1) ? (lw >> 16) : (lw & 0xFFFFu);
        return hi != 0xFFu;
    } else {
        key = reinterpret_cast<const uint64_t*>(w)[s];
        const uint32_t lw = w[6 + (s >> 1)];
        label = (s & 1) ? (lw >> 16) : (lw & 0xFFFFu);
        return key != OVF_EMPTY;
    }
}

__device__ __forceinline__ uint64_t wide_key(const uint32_t* w, int s) { return reinterpret_cast<const uint64_t*>(w)[s]; }
__device__ __forceinline__ uint32_t wide_label(const uint32_t* w, int s) { const uint32_t lw = w[6 + (s >> 1)]; return (s & 1) ? (lw >> 16) : (lw & 0xFFFFu); }

// scan the overflow probe sequence of c: sets differ if a copy with another label
// exists; returns the number of copies seen strictly before (stop_b, stop_s)
// (pass stop_b = ~0 to count all copies).
__device__ __forceinline__ uint32_t ovf_scan(const uint4* ovf, uint64_t n_ovf, uint64_t c, uint32_t label,
                                             uint64_t stop_b, int stop_s, bool& differ, uint32_t& total) {
    uint32_t before = 0;
    bool passed = false;
    total = 0;
    uint64_t b = ovf_home(c, n_ovf);
    for (uint64_t n = 0; n < n_ovf; n++) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(ovf + 2 * b);
        for (int s = 0; s < WIDE_SLOTS; s++) {
            if (b == stop_b && s == stop_s) { passed = true; continue; }
            if (wide_key(w, s) != c) continue;
            total++;
            if (!passed) before++;
            if (wide_label(w, s) != label) differ = true;
        }
        if (wide_key(w, WIDE_SLOTS - 1) == OVF_EMPTY) break;
        if (++b == n_ovf) b = 0;
    }
    return before;
}

// LOCAL: the copies of one k-mer (candidate sectors la / lbb, local indices; key q without the alt bit) in the
// main table other than slot (my_lb, my_slot). Order of the copies: A slots, then B slots. Sets differ if a copy
// carries another label, any if there is a copy at all; returns whether a copy precedes (my_lb, my_slot).
__device__ __forceinline__ bool local_main_copies(const uint4* table, uint64_t la, uint64_t lbb, uint64_t q, uint32_t label,
                                                  uint64_t my_lb, int my_slot, bool& differ, bool& any) {
    bool earlier = false, passed = false;
    any = false;
    for (int t = 0; t < 2; t++) {
        if (t && lbb == la) break;
        const uint64_t l = t ? lbb : la;
        const uint64_t key = t ? q | ((uint64_t)LOCAL_ALT_BIT << 32) : q;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(table + 2 * l);
        for (int s = 0; s < LOCAL_SLOTS; s++) {
            if (l == my_lb && s == my_slot) { passed = true; continue; }
            uint64_t kk; uint32_t ll;
            if (!read_slot<LAYOUT_LOCAL>(w, s, kk, ll) || kk != key) continue;
            any = true;
            if (ll != label) differ = true;
            if (!passed) earlier = true;
        }
    }
    return earlier;
}
__device__ __forceinline__ bool sector_flagged(const uint4* table, uint64_t l) {
    return (reinterpret_cast<const uint32_t*>(table + 2 * l)[7] >> 16) & 1u;
}

template <int LAYOUT>
__global__ void k_dedupe_main(DedupeView<LAYOUT> v) {
    constexpr int SLOTS = SlotsOf<LAYOUT>::n;
    const uint64_t lb = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (lb >= v.n_local) return;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(v.table + 2 * lb);
    const bool flagged = (w[7] >> 16) & 1u;
    uint64_t keys[SLOTS];
    uint32_t labels[SLOTS];
    bool used[SLOTS];
    for (int s = 0; s < SLOTS; s++) used[s] = read_slot<LAYOUT>(w, s, keys[s], labels[s]);
    uint32_t del = 0;
    if (LAYOUT == LAYOUT_LOCAL) {
        const uint64_t line_lo = v.lo >> 2;
        const uint32_t line_n = (uint32_t)(v.n_local >> 2);
        for (int i = 0; i < SLOTS; i++) {
            if (!used[i]) continue;
            const uint64_t c = local_rebuild2((lb + v.lo) >> 2, keys[i], v.k, v.NL, line_lo, line_n);
            uint64_t sa, sb, q;
            local_locate2(c, v.k, v.NL, line_lo, line_n, sa, sb, q);
            bool differ = false, any;
            const bool earlier = local_main_copies(v.table, sa - v.lo, sb - v.lo, q, labels[i], lb, i, differ, any);
            if (v.n_ovf && sector_flagged(v.table, sa - v.lo) && sector_flagged(v.table, sb - v.lo)) {
                uint32_t total;
                ovf_scan(v.ovf, v.n_ovf, c, labels[i], ~0ull, 0, differ, total);
            }
            if (differ || earlier) del |= 1u << i;  // main copies precede overflow copies
        }
    } else {
    for (int i = 0; i < SLOTS; i++) {
        if (!used[i]) continue;
        bool differ = false, earlier = false;
        for (int j = 0; j < SLOTS; j++) {
            if (j == i || !used[j] || keys[j] != keys[i]) continue;
            if (labels[j] != labels[i]) differ = true;
            if (j < i) earlier = true;
        }
        if (flagged && v.n_ovf) {
            uint32_t total;
            const uint64_t c = LAYOUT == LAYOUT_LOCAL ? local_rebuild((lb + v.lo) >> 2, keys[i], v.k, v.NL) : keys[i] * v.M + (lb + v.lo);
            ovf_scan(v.ovf, v.n_ovf, c, labels[i], ~0ull, 0, differ, total);
        }
        if (differ || earlier) del |= 1u << i;      // main copies precede overflow copies
    }
    }
    v.del_main[lb] = (uint8_t)del;
}

template <int LAYOUT>
__global__ void k_dedupe_ovf(DedupeView<LAYOUT> v, uint64_t magic) {
    constexpr int SLOTS = SlotsOf<LAYOUT>::n;
    const uint64_t ob = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (ob >= v.n_ovf) return;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(v.ovf + 2 * ob);
    uint32_t del = 0;
    for (int s = 0; s < WIDE_SLOTS; s++) {
        const uint64_t c = wide_key(w, s);
        if (c == OVF_EMPTY) break;
        const uint32_t label = wide_label(w, s);
        bool differ = false;
        uint32_t total;
        const uint32_t before = ovf_scan(v.ovf, v.n_ovf, c, label, ob, s, differ, total);
        if (LAYOUT == LAYOUT_LOCAL) {               // copies in the two candidate sectors
            uint64_t sa, sb, q;
            local_locate2(c, v.k, v.NL, v.lo >> 2, (uint32_t)(v.n_local >> 2), sa, sb, q);
            bool in_main = false;                   // (tie k-mers never have copies in the lines)
            if (sa - v.lo < v.n_local) local_main_copies(v.table, sa - v.lo, sb - v.lo, q, label, ~0ull, 0, differ, in_main);
            if (differ || in_main || before) del |= 1u << s;
            continue;
        }
        // copies in the home bucket
        uint64_t q, b;
        home_of<LAYOUT>(c, v.M, magic, v.NL, v.k, q, b);
        const uint32_t* hw = reinterpret_cast<const uint32_t*>(v.table + 2 * (b - v.lo));
        bool in_main = false;
        for (int j = 0; j < SLOTS; j++) {
            uint64_t kk; uint32_t ll;
            if (!read_slot<LAYOUT>(hw, j, kk, ll)) continue;
            if (kk != q) continue;
            in_main = true;
            if (ll != label) differ = true;
        }
        if (differ || in_main || before) del |= 1u << s;
    }
    v.del_ovf[ob] = (uint8_t)del;
}

template <int LAYOUT>
__global__ void k_dedupe_apply(DedupeView<LAYOUT> v, unsigned long long* removed) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    unsigned long long gone = 0;
    if (i < v.n_local && v.del_main[i]) {
        uint32_t* w = reinterpret_cast<uint32_t*>(v.table + 2 * i);
        const uint32_t del = v.del_main[i];
        for (int s = 0; s < SlotsOf<LAYOUT>::n; s++) {
            if (!((del >> s) & 1u)) continue;
            if (LAYOUT == LAYOUT_NARROW) w[s] = 0xFFFFFFFFu;
            else if (LAYOUT == LAYOUT_LOCAL) { w[s] = 0xFFFFFFFFu; w[6] |= 0xFFu << (8 * s); }
            else reinterpret_cast<uint64_t*>(w)[s] = OVF_EMPTY;
            gone++;
        }
    }
    if (i < v.n_ovf && v.del_ovf[i]) {
        uint32_t* w = reinterpret_cast<uint32_t*>(v.ovf + 2 * i);
        const uint32_t del = v.del_ovf[i];
        for (int s = 0; s < WIDE_SLOTS; s++) {
            if (!((del >> s) & 1u)) continue;
            reinterpret_cast<uint64_t*>(w)[s] = OVF_TOMBSTONE;     // keeps the probe sequence intact
            gone++;
        }
    }
    if (gone) atomicAdd(removed, gone);
}

// ---- host side ------------------------------------------------------------------
struct Geometry {
    int layout;
    uint64_t M, lo, n_local;
    uint64_t NL;            // LOCAL: lines; M = 4 NL
};

// LOCAL needs mix(minimizer) div NL to fit LOCAL_ZQ_BITS: NL >= 4^m / 2^19 (m = k - 7)
uint64_t local_min_lines(int k) {
    const int mbits = 2 * (k - LOCAL_W + 1);
    return mbits > LOCAL_ZQ_BITS ? (1ull << (mbits - LOCAL_ZQ_BITS)) : 1;
}

// constants of local_divmod() (common.cuh) for this table
void set_local_divmod(TableView& v) {
    v.nl_m32 = 0; v.nl_sh = 0;
    if (v.layout != LAYOUT_LOCAL || v.NL < 2) return;
    const int mbits = 2 * (v.k - LOCAL_W + 1);
    v.nl_sh = mbits > 32 ? mbits - 32 : 0;
    v.nl_m32 = (uint32_t)((((__uint128_t)1) << (32 + v.nl_sh)) / v.NL);
}

uint64_t pow4(int k) { return k >= 32 ? 0 : (1ull << (2 * k)); }   // 0 means 2^64

Geometry choose_geometry(const cuclark_config& cfg, uint64_t n_entries, double grow, bool allow_auto_local = true) {
    Geometry g;
    const double narrow_load = cfg.bucket_load > 0 ? cfg.bucket_load : 2.6;
    const double wide_load = cfg.bucket_load > 0 ? std::min(cfg.bucket_load, 2.0) : 1.5;
    // quotient c / M must stay below 2^32-1 in the narrow layout
    uint64_t m_min_narrow = 0;
    bool narrow_possible = cfg.k < 32;
    if (narrow_possible) m_min_narrow = pow4(cfg.k) / 0xFFFFFFFFull + 2;
    uint64_t m_narrow = (uint64_t)((double)n_entries / narrow_load * grow) + 64;
    uint64_t m_wide = (uint64_t)((double)n_entries / wide_load * grow) + 64;
    int layout = cfg.layout;
    // LOCAL needs at least local_min_lines(k) lines (key width); a table that would be mostly
    // empty at that size (small database at large k) uses the hashed layouts instead
    // (and enough distinct minimizers: about 1 in 9 canonical m-mers ever is one)
    const double local_load = cfg.bucket_load > 0 ? std::min(cfg.bucket_load, 3.8) : 3.0;
    const double local_lines = (double)n_entries / (4.0 * local_load) * grow + 16;
    // CUCLARK_ALLOW_SPARSE_TABLE=1 (parity tests): an explicitly requested LOCAL table is built at its minimum
    // size even when the database leaves it mostly empty — k=31 needs 2^29 lines = 68.7 GB whatever it holds, and
    // this is how the kernel instantiation of the bacterial-scale run meets the oracle on a small database
    const bool allow_sparse = cfg.layout == LAYOUT_LOCAL && getenv("CUCLARK_ALLOW_SPARSE_TABLE") != nullptr;
    const bool local_fits = cfg.k >= LOCAL_MIN_K && cfg.k <= 32 &&
        (allow_sparse || local_min_lines(cfg.k) * 128ull <= (8ull << 30) || (double)local_min_lines(cfg.k) <= 4.0 * local_lines) &&
        (double)pow4(cfg.k - LOCAL_W + 1) / 9.0 >= 4.0 * local_lines &&
        local_lines < 1.0e9 && (double)local_min_lines(cfg.k) < 1.0e9;     // 4 NL < 2^32 sectors, 3 NL < 2^32 (local_divmod)
    // automatic choice: minimizer lines for a single-device table that fills them (at bacterial scale they
    // serve ~1.25x the lookups of the hashed sectors; a small database would pay the minimum size for nothing)
    if (layout == 0 && allow_auto_local && cfg.shard_count <= 1 && local_fits && cfg.bucket_load <= 0 &&
        (double)local_min_lines(cfg.k) <= 1.25 * local_lines && 4.0 * local_lines < 4.2e9 && !getenv("CUCLARK_NO_LOCAL"))
        layout = LAYOUT_LOCAL;
    if (layout == LAYOUT_LOCAL && local_fits) {
        // entries per 4-slot sector (two candidate lines per minimizer even out the clumps of ~4.5 k-mers
        // that share one)
        uint64_t nl = (uint64_t)local_lines;
        nl = std::max(nl, local_min_lines(cfg.k)) | 1ull;
        g.layout = LAYOUT_LOCAL;
        g.NL = nl;
        g.M = 4 * nl;
        const uint64_t G = cfg.shard_count > 1 ? cfg.shard_count : 1, i = cfg.shard_count > 1 ? cfg.shard_index : 0;
        g.lo = 4 * (uint64_t)((__uint128_t)nl * i / G);          // shards hold whole lines
        g.n_local = 4 * (uint64_t)((__uint128_t)nl * (i + 1) / G) - g.lo;
        return g;
    }
    if (layout == LAYOUT_LOCAL) layout = 0;                      // k too small / table too small for minimizer lines
    g.NL = 0;
    if (layout == 0) {
        if (!narrow_possible) layout = LAYOUT_WIDE;
        else {
            const uint64_t need = std::max(m_narrow, m_min_narrow);
            const uint64_t small = (256ull << 20) / 32;
            layout = (need <= small || (double)need <= 1.25 * (double)m_wide) ? LAYOUT_NARROW : LAYOUT_WIDE;
        }
    }
    if (layout == LAYOUT_NARROW && !narrow_possible) layout = LAYOUT_WIDE;
    g.layout = layout;
    g.M = layout == LAYOUT_NARROW ? std::max(m_narrow, m_min_narrow) : m_wide;
    g.M |= 1ull;
    const uint64_t G = cfg.shard_count > 1 ? cfg.shard_count : 1, i = cfg.shard_count > 1 ? cfg.shard_index : 0;
    g.lo = (uint64_t)((__uint128_t)g.M * i / G);
    const uint64_t hi = (uint64_t)((__uint128_t)g.M * (i + 1) / G);
    g.n_local = hi - g.lo;
    return g;   // callers check n_local < 2^32 (the classify kernel keeps local bucket ids in 32 bit)
}

struct BuildBuffers {
    uint4* table = nullptr;
    uint4* ovf = nullptr;
    uint32_t* ovf_cnt8 = nullptr;
    uint8_t* del_main = nullptr;
    uint8_t* del_ovf = nullptr;
    uint32_t* cnt8 = nullptr;
    uint64_t* ovf_c = nullptr;
    uint16_t* ovf_l = nullptr;
    uint32_t* flags = nullptr;
    unsigned long long* counters = nullptr;   // [0] inserted, [1] spill buckets, [2] removed
    void free_temp() {
        cudaFree(cnt8); cudaFree(ovf_c); cudaFree(ovf_l); cudaFree(flags); cudaFree(counters);
        cudaFree(ovf_cnt8); cudaFree(del_main); cudaFree(del_ovf);
        cnt8 = nullptr; ovf_c = nullptr; ovf_l = nullptr; flags = nullptr; counters = nullptr;
        ovf_cnt8 = nullptr; del_main = nullptr; del_ovf = nullptr;
    }
    void free_all() { free_temp(); cudaFree(table); cudaFree(ovf); table = nullptr; ovf = nullptr; }
};

int alloc_build(const Geometry& g, uint64_t n_expected, BuildBuffers& b, BuildCtx& x, uint64_t htsize, int k) {
    x.k = k;
    if (g.n_local >= 0xFFFFFFFFull) { set_error("shard of %llu buckets exceeds 2^32: use more shards", (unsigned long long)g.n_local); return CUCLARK_ERR_ARG; }
    const uint64_t G = 1;
    (void)G;
    uint64_t ovf_cap64 = (g.layout == LAYOUT_LOCAL ? n_expected / 3 : n_expected / 6) + (1u << 16);
    if (ovf_cap64 > 0xFFFFFF00ull) ovf_cap64 = 0xFFFFFF00ull;
    if (cudaMalloc(&b.table, g.n_local * 32) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc of %.2f GB table failed", g.n_local * 32 / 1e9); return CUCLARK_ERR_NOMEM; }
    CK(cudaMalloc(&b.cnt8, (g.n_local / 4 + 1) * 4));
    CK(cudaMalloc(&b.ovf_c, ovf_cap64 * 8));
    CK(cudaMalloc(&b.ovf_l, ovf_cap64 * 2));
    CK(cudaMalloc(&b.flags, 16));
    CK(cudaMalloc(&b.counters, 32));
    CK(cudaMemset(b.cnt8, 0, (g.n_local / 4 + 1) * 4));
    CK(cudaMemset(b.flags, 0, 16));
    CK(cudaMemset(b.counters, 0, 32));
    const uint64_t n4 = 2 * g.n_local;
    k_init_table<<<(unsigned)((n4 + 255) / 256), 256>>>(b.table, g.n_local, g.layout);
    CK(cudaGetLastError());
    x.table = b.table; x.cnt8 = b.cnt8; x.M = g.M; x.magic = (uint64_t)((((__uint128_t)1) << 64) / g.M);
    x.lo = g.lo; x.n_local = g.n_local; x.htsize = htsize;
    x.NL = g.NL; x.layout = g.layout;
    x.ovf_c = b.ovf_c; x.ovf_l = b.ovf_l; x.ovf_cap = (uint32_t)ovf_cap64; x.flags = b.flags;
    x.inserted = b.counters;
    return CUCLARK_OK;
}

// spills, stats, error check; returns CUCLARK_ERR_BUILD if the geometry was too tight
int finish_build(cuclark_db* db, const Geometry& g, BuildBuffers& b, BuildCtx& x, bool dedupe) {
    uint32_t flags[2];
    CK(cudaMemcpy(flags, b.flags, 8, cudaMemcpyDeviceToHost));
    if (flags[1]) { set_error("table build: bucket counter/overflow-list exhausted (flags %u)", flags[1]); return CUCLARK_ERR_BUILD; }
    const uint32_t n_ovf_entries = flags[0];
    uint64_t n_ovf = 0;
    uint64_t n_left = n_ovf_entries;                 // entries that really go to the overflow table
    if (n_ovf_entries && g.layout == LAYOUT_LOCAL && !getenv("CUCLARK_NO_RESCUE")) {
        const uint64_t n_locks = (g.n_local >> 2) / LOCAL_ALT_BLOCK + 1;
        uint32_t* locks = nullptr;
        CK(cudaMalloc(&locks, n_locks * 4));
        CK(cudaMemset(locks, 0, n_locks * 4));
        CK(cudaMemset(b.counters + 3, 0, 8));
        k_local_rescue<<<(n_ovf_entries + 127) / 128, 128>>>(x, n_ovf_entries, locks, b.counters + 3);
        unsigned long long rescued = 0;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(&rescued, b.counters + 3, 8, cudaMemcpyDeviceToHost);
        cudaFree(locks);
        if (e != cudaSuccess) { set_error("k_local_rescue failed: %s", cudaGetErrorString(e)); return CUCLARK_ERR_CUDA; }
        n_left = n_ovf_entries - rescued;
    }
    if (n_ovf_entries) {
        double ovf_load = g.layout == LAYOUT_LOCAL ? OVF_LOAD_LOCAL : OVF_LOAD;
        if (const char* e = getenv("CUCLARK_OVF_LOAD")) { const double v = atof(e); if (v >= 0.2 && v <= 2.4) ovf_load = v; }   // tuning knob
        n_ovf = (uint64_t)((double)n_left / ovf_load) + 64;
        if (cudaMalloc(&b.ovf, n_ovf * 32) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc of overflow table failed"); return CUCLARK_ERR_NOMEM; }
        CK(cudaMalloc(&b.ovf_cnt8, (n_ovf / 4 + 1) * 4));
        CK(cudaMemset(b.ovf_cnt8, 0, (n_ovf / 4 + 1) * 4));
        k_init_table<<<(unsigned)((2 * n_ovf + 255) / 256), 256>>>(b.ovf, n_ovf, LAYOUT_WIDE);
        CK(cudaGetLastError());
        x.ovf = b.ovf; x.ovf_cnt8 = b.ovf_cnt8; x.n_ovf = n_ovf;
        k_place_spills<<<(n_ovf_entries + 255) / 256, 256>>>(x, n_ovf_entries);
        CK(cudaGetLastError());
    }
    if (dedupe) {
        CK(cudaMalloc(&b.del_main, g.n_local));
        CK(cudaMalloc(&b.del_ovf, n_ovf + 1));
        const unsigned blocks = (unsigned)((std::max(g.n_local, n_ovf) + 255) / 256);
        if (g.layout == LAYOUT_NARROW) {
            DedupeView<LAYOUT_NARROW> v{b.table, b.ovf, g.M, g.lo, g.n_local, n_ovf, b.del_main, b.del_ovf, g.NL, x.k};
            k_dedupe_main<LAYOUT_NARROW><<<(unsigned)((g.n_local + 255) / 256), 256>>>(v);
            if (n_ovf) k_dedupe_ovf<LAYOUT_NARROW><<<(unsigned)((n_ovf + 255) / 256), 256>>>(v, x.magic);
            else CK(cudaMemset(b.del_ovf, 0, 1));
            k_dedupe_apply<LAYOUT_NARROW><<<blocks, 256>>>(v, b.counters + 2);
        } else if (g.layout == LAYOUT_LOCAL) {
            DedupeView<LAYOUT_LOCAL> v{b.table, b.ovf, g.M, g.lo, g.n_local, n_ovf, b.del_main, b.del_ovf, g.NL, x.k};
            k_dedupe_main<LAYOUT_LOCAL><<<(unsigned)((g.n_local + 255) / 256), 256>>>(v);
            if (n_ovf) k_dedupe_ovf<LAYOUT_LOCAL><<<(unsigned)((n_ovf + 255) / 256), 256>>>(v, x.magic);
            else CK(cudaMemset(b.del_ovf, 0, 1));
            k_dedupe_apply<LAYOUT_LOCAL><<<blocks, 256>>>(v, b.counters + 2);
        } else {
            DedupeView<LAYOUT_WIDE> v{b.table, b.ovf, g.M, g.lo, g.n_local, n_ovf, b.del_main, b.del_ovf, g.NL, x.k};
            k_dedupe_main<LAYOUT_WIDE><<<(unsigned)((g.n_local + 255) / 256), 256>>>(v);
            if (n_ovf) k_dedupe_ovf<LAYOUT_WIDE><<<(unsigned)((n_ovf + 255) / 256), 256>>>(v, x.magic);
            else CK(cudaMemset(b.del_ovf, 0, 1));
            k_dedupe_apply<LAYOUT_WIDE><<<blocks, 256>>>(v, b.counters + 2);
        }
        CK(cudaGetLastError());
    }
    k_table_stats<<<(unsigned)((g.n_local + 255) / 256), 256>>>(b.table, g.n_local, b.counters + 1);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(flags, b.flags, 8, cudaMemcpyDeviceToHost));
    if (flags[1]) { set_error("table build: overflow table full (flags %u)", flags[1]); return CUCLARK_ERR_BUILD; }
    unsigned long long counters[3];
    CK(cudaMemcpy(counters, b.counters, 24, cudaMemcpyDeviceToHost));
    db->n_entries = counters[0] - counters[2];
    db->n_spilled = n_left;
    db->n_spill_buckets = counters[1];
    db->d_table = b.table;
    db->d_ovf = b.ovf;
    db->view.buckets = b.table;
    db->view.ovf = b.ovf;
    db->view.n_ovf = n_ovf;
    db->view.M = g.M;
    db->view.magic = x.magic;
    db->view.lo = g.lo;
    db->view.n_local = g.n_local;
    db->view.layout = g.layout;
    db->view.k = db->cfg.k;
    db->view.NL = g.NL;
    db->view.magicNL = g.NL ? (uint64_t)((((__uint128_t)1) << 64) / g.NL) : 0;
    set_local_divmod(db->view);
    db->view.line_lo = (uint32_t)(g.lo >> 2);
    db->view.line_n = (uint32_t)(g.n_local >> 2);
    b.free_temp();
    return CUCLARK_OK;
}

}  // namespace

void table_plan(const cuclark_config& cfg, uint64_t n_entries, cuclark_table_plan* out) {
    const Geometry g = choose_geometry(cfg, n_entries, 1.0);
    out->layout = g.layout;
    out->n_buckets = g.M;
    out->n_local_buckets = g.n_local;
    out->home_bytes = g.n_local * 32;
}

// The loaded table of `src` copied device to device into `dst` (another device, same configuration): the files
// are read and re-bucketed ONCE, the replicas travel over NVLink ("read files once, broadcast").
int table_clone(cuclark_db* src, cuclark_db* dst) {
    if (!src->d_table) { set_error("no database loaded in the source handle"); return CUCLARK_ERR_STATE; }
    if (src->cfg.k != dst->cfg.k || src->cfg.htsize != dst->cfg.htsize || src->cfg.n_targets != dst->cfg.n_targets ||
        src->cfg.shard_index != dst->cfg.shard_index || src->cfg.shard_count != dst->cfg.shard_count || src->key_bytes != dst->key_bytes) {
        set_error("cuclark_clone_table: the handles differ in k, HTSIZE, n_targets, key width or shard");
        return CUCLARK_ERR_ARG;
    }
    // direct NVLink copies need peer access in both directions (without it the copy is staged through the host: 36 GB/s)
    if (src->cfg.device != dst->cfg.device) {
        const int pair[2][2] = {{dst->cfg.device, src->cfg.device}, {src->cfg.device, dst->cfg.device}};
        for (auto& pr : pair) {
            int can = 0;
            if (cudaSetDevice(pr[0]) == cudaSuccess && cudaDeviceCanAccessPeer(&can, pr[0], pr[1]) == cudaSuccess && can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(pr[1], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return CUCLARK_ERR_CUDA; }
            }
            cudaGetLastError();
        }
    }
    CK(cudaSetDevice(dst->cfg.device));
    table_free(dst);
    const size_t tb = src->view.n_local * 32, ob = src->view.n_ovf * 32;
    uint4 *t = nullptr, *o = nullptr;
    if (cudaMalloc(&t, tb) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc of %.2f GB table failed", tb / 1e9); return CUCLARK_ERR_NOMEM; }
    if (ob && cudaMalloc(&o, ob) != cudaSuccess) { cudaGetLastError(); cudaFree(t); set_error("cudaMalloc of overflow table failed"); return CUCLARK_ERR_NOMEM; }
    cudaError_t e = cudaMemcpyPeerAsync(t, dst->cfg.device, src->d_table, src->cfg.device, tb, dst->stream);
    if (e == cudaSuccess && ob) e = cudaMemcpyPeerAsync(o, dst->cfg.device, src->d_ovf, src->cfg.device, ob, dst->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dst->stream);
    if (e != cudaSuccess) { cudaFree(t); cudaFree(o); set_error("device-to-device copy of the table failed: %s", cudaGetErrorString(e)); return CUCLARK_ERR_CUDA; }
    dst->d_table = t; dst->d_ovf = o;
    dst->view = src->view;
    dst->view.buckets = t; dst->view.ovf = o;
    dst->n_entries = src->n_entries; dst->n_spilled = src->n_spilled; dst->n_spill_buckets = src->n_spill_buckets;
    dst->src_sfactor = src->src_sfactor;
    for (int i = 0; i < 3; i++) { dst->src_bytes[i] = src->src_bytes[i]; dst->src_mtime_ns[i] = src->src_mtime_ns[i]; }
    return CUCLARK_OK;
}

void table_free(cuclark_db* db) {
    if (db->d_table) cudaFree(db->d_table);
    if (db->d_ovf) cudaFree(db->d_ovf);
    db->d_table = nullptr;
    db->d_ovf = nullptr;
    db->view = TableView{};
    db->n_entries = db->n_spilled = db->n_spill_buckets = 0;
    db->src_sfactor = 1;
    db->src_bytes[0] = db->src_bytes[1] = db->src_bytes[2] = 0;
    db->src_mtime_ns[0] = db->src_mtime_ns[1] = db->src_mtime_ns[2] = 0;
}

// ---- loading <base>.sz/.ky/.lb (or the same three arrays from host memory) ----------------------
namespace {

void file_stamp(const std::string& p, uint64_t& size, uint64_t& mtime_ns);

// One staging set: a chunk of reference buckets with its keys and labels, host (pinned) and device.
struct LoadStage {
    uint8_t *h_sz = nullptr, *h_keep = nullptr, *d_sz = nullptr, *d_keep = nullptr;
    uint64_t *h_rel = nullptr, *d_rel = nullptr;
    uint8_t *h_keys = nullptr, *d_keys = nullptr;
    uint16_t *h_labels = nullptr, *d_labels = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    bool busy = false;
    int init(uint64_t max_buckets, uint64_t max_entries, int kb, bool keep) {
        CK(cudaMallocHost(&h_sz, max_buckets)); CK(cudaMalloc(&d_sz, max_buckets));
        if (keep) { CK(cudaMallocHost(&h_keep, max_buckets)); CK(cudaMalloc(&d_keep, max_buckets)); }
        CK(cudaMallocHost(&h_rel, (CHUNK_BLOCKS + 1) * 8)); CK(cudaMalloc(&d_rel, (CHUNK_BLOCKS + 1) * 8));
        CK(cudaMallocHost(&h_keys, max_entries * kb)); CK(cudaMalloc(&d_keys, max_entries * kb));
        CK(cudaMallocHost(&h_labels, max_entries * 2)); CK(cudaMalloc(&d_labels, max_entries * 2));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        return CUCLARK_OK;
    }
    ~LoadStage() {
        if (st) cudaStreamSynchronize(st);
        cudaFreeHost(h_sz); cudaFreeHost(h_keep); cudaFreeHost(h_rel); cudaFreeHost(h_keys); cudaFreeHost(h_labels);
        cudaFree(d_sz); cudaFree(d_keep); cudaFree(d_rel); cudaFree(d_keys); cudaFree(d_labels);
        if (done) cudaEventDestroy(done);
        if (st) cudaStreamDestroy(st);
    }
};

bool pread_all(int fd, void* dst, size_t n, uint64_t off) {
    uint8_t* p = static_cast<uint8_t*>(dst);
    while (n) {
        const ssize_t r = pread(fd, p, n, (off_t)off);
        if (r <= 0) return false;
        p += r; n -= (size_t)r; off += (uint64_t)r;
    }
    return true;
}

// runs f(t) on n_threads host threads
template <typename F>
void parallel_for_threads(int n_threads, F f) {
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back([&f, t] { f(t); });
    f(0);
    for (auto& x : th) x.join();
}

}  // namespace

int table_build_from_arrays(cuclark_db* db, const uint8_t* sz_in, const void* ky, const uint16_t* lb,
                            uint64_t n_entries_file, int sfactor, const char* base_path) {
    const uint64_t H = db->cfg.htsize;
    const int kb = db->key_bytes;
    const bool timing = getenv("CUCLARK_TIMING") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    // files: .sz is mapped (scanned in place by several threads), .ky/.lb are read chunk by chunk
    struct Files {
        int sz = -1, ky = -1, lb = -1;
        void* map = nullptr; size_t map_n = 0;
        ~Files() { if (map) munmap(map, map_n); if (sz >= 0) close(sz); if (ky >= 0) close(ky); if (lb >= 0) close(lb); }
    } fs;
    const uint8_t* sz = sz_in;
    if (base_path) {
        const std::string b(base_path);
        fs.sz = open((b + ".sz").c_str(), O_RDONLY);
        if (fs.sz < 0) { set_error("Failed to open %s.sz", base_path); return CUCLARK_ERR_IO; }
        fs.ky = open((b + ".ky").c_str(), O_RDONLY);
        if (fs.ky < 0) { set_error("Failed to open %s.ky", base_path); return CUCLARK_ERR_IO; }
        fs.lb = open((b + ".lb").c_str(), O_RDONLY);
        if (fs.lb < 0) { set_error("Failed to open %s.lb", base_path); return CUCLARK_ERR_IO; }
        struct stat sb;
        if (fstat(fs.sz, &sb) != 0 || (uint64_t)sb.st_size < H) { set_error("%s.sz is short (%llu of %llu bytes)", base_path, (unsigned long long)sb.st_size, (unsigned long long)H); return CUCLARK_ERR_IO; }
        fs.map = mmap(nullptr, H, PROT_READ, MAP_PRIVATE, fs.sz, 0);
        if (fs.map == MAP_FAILED) { fs.map = nullptr; set_error("cannot map %s.sz", base_path); return CUCLARK_ERR_IO; }
        fs.map_n = H;
        madvise(fs.map, H, MADV_WILLNEED);
        sz = static_cast<const uint8_t*>(fs.map);
    }

    // host pass over the bucket sizes, on several threads: entries before every 1024-bucket block,
    // -s sampling (src/CuClarkDB.cu:511-524: of the non-empty buckets numbered 1,2,3... keep every
    // sfactor-th) and the kept-entry count
    const uint64_t n_blocks = (H + CHUNK_THREADS - 1) / CHUNK_THREADS;
    std::vector<uint64_t> coarse(n_blocks + 1);
    std::vector<uint8_t> keep;
    if (sfactor > 1) keep.assign(H, 0);
    const int n_thr = (int)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)std::thread::hardware_concurrency(), 16, n_blocks / 64 + 1}));
    std::vector<uint64_t> part_total(n_thr + 1, 0), part_nonzero(n_thr + 1, 0), part_kept(n_thr, 0);
    auto block_range = [&](int t, uint64_t& b0, uint64_t& b1) { b0 = n_blocks * t / n_thr; b1 = n_blocks * (t + 1) / n_thr; };
    parallel_for_threads(n_thr, [&](int t) {
        uint64_t b0, b1; block_range(t, b0, b1);
        uint64_t total = 0, nonzero = 0;
        for (uint64_t bk = b0; bk < b1; bk++) {
            coarse[bk] = total;                                   // relative to the thread's range; rebased below
            const uint64_t r1 = std::min<uint64_t>((bk + 1) * CHUNK_THREADS, H);
            for (uint64_t r = bk * CHUNK_THREADS; r < r1; r++) { total += sz[r]; nonzero += sz[r] != 0; }
        }
        part_total[t + 1] = total; part_nonzero[t + 1] = nonzero;
    });
    for (int t = 0; t < n_thr; t++) { part_total[t + 1] += part_total[t]; part_nonzero[t + 1] += part_nonzero[t]; }
    const uint64_t total = part_total[n_thr];
    parallel_for_threads(n_thr, [&](int t) {
        uint64_t b0, b1; block_range(t, b0, b1);
        for (uint64_t bk = b0; bk < b1; bk++) coarse[bk] += part_total[t];
        if (sfactor > 1) {
            uint64_t nonzero = part_nonzero[t], kept = 0;
            const uint64_t r1 = std::min<uint64_t>(b1 * CHUNK_THREADS, H);
            for (uint64_t r = b0 * CHUNK_THREADS; r < r1; r++) {
                if (!sz[r]) continue;
                if ((++nonzero % (uint64_t)sfactor) == 0) { keep[r] = 1; kept += sz[r]; }
            }
            part_kept[t] = kept;
        }
    });
    coarse[n_blocks] = total;
    uint64_t kept = total;
    if (sfactor > 1) { kept = 0; for (int t = 0; t < n_thr; t++) kept += part_kept[t]; }
    db->src_sfactor = sfactor > 1 ? sfactor : 1;
    db->src_bytes[0] = base_path ? H : 0;
    db->src_bytes[1] = base_path ? total * (uint64_t)kb : 0;
    db->src_bytes[2] = base_path ? total * 2 : 0;
    for (int i = 0; i < 3; i++) db->src_mtime_ns[i] = 0;
    if (base_path) {
        const char* ext[3] = {".sz", ".ky", ".lb"};
        for (int i = 0; i < 3; i++) { uint64_t sz_i; file_stamp(std::string(base_path) + ext[i], sz_i, db->src_mtime_ns[i]); }
    }
    if (!base_path && total != n_entries_file) { set_error("bucket sizes sum to %llu entries but %llu were passed", (unsigned long long)total, (unsigned long long)n_entries_file); return CUCLARK_ERR_ARG; }
    const double ms_scan = since(t_begin);
    const auto t_build = std::chrono::steady_clock::now();

    // largest chunk, for the staging sets
    uint64_t max_chunk_entries = 1;
    for (uint64_t b0 = 0; b0 < n_blocks; b0 += CHUNK_BLOCKS) {
        const uint64_t b1 = std::min<uint64_t>(b0 + CHUNK_BLOCKS, n_blocks);
        max_chunk_entries = std::max(max_chunk_entries, coarse[b1] - coarse[b0]);
    }
    const uint64_t max_chunk_buckets = std::min<uint64_t>((uint64_t)CHUNK_BLOCKS * CHUNK_THREADS, H);
    LoadStage stage[2];
    for (auto& s : stage) {
        const int rc = s.init(max_chunk_buckets, max_chunk_entries, kb, sfactor > 1);
        if (rc) return rc;
    }

    int rc = CUCLARK_ERR_BUILD;
    double grow = 1.0;
    bool auto_local = true, grow_next = false;   // an automatically chosen LOCAL table falls back to the hashed layouts if it cannot be built
    for (int attempt = 0; attempt < 6 && rc == CUCLARK_ERR_BUILD; attempt++) {
        if (grow_next) grow *= 1.3;
        const Geometry g = choose_geometry(db->cfg, kept, grow, auto_local);
        const bool is_auto_local = db->cfg.layout == 0 && g.layout == LAYOUT_LOCAL;
        grow_next = !is_auto_local;
        if (is_auto_local) auto_local = false;
        BuildBuffers bb;
        BuildCtx x{};
        rc = alloc_build(g, kept / (db->cfg.shard_count > 1 ? db->cfg.shard_count : 1), bb, x, H, db->cfg.k);
        if (rc != CUCLARK_OK) { bb.free_all(); if (is_auto_local && rc != CUCLARK_ERR_ARG) { cudaGetLastError(); rc = CUCLARK_ERR_BUILD; continue; } break; }
        if (cudaDeviceSynchronize() != cudaSuccess) {              // the table is initialised on the default stream
            set_error("table initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
            bb.free_all();
            return CUCLARK_ERR_CUDA;
        }
        uint64_t n_chunk = 0;
        for (uint64_t b0 = 0; b0 < n_blocks && rc == CUCLARK_OK; b0 += CHUNK_BLOCKS) {
            const uint64_t b1 = std::min<uint64_t>(b0 + CHUNK_BLOCKS, n_blocks);
            const uint64_t r0 = b0 * CHUNK_THREADS, r1 = std::min<uint64_t>(b1 * CHUNK_THREADS, H);
            const uint64_t e0 = coarse[b0], ne = coarse[b1] - e0;
            if (ne == 0) continue;
            // the host fills one staging set while the device works on the other
            LoadStage& S = stage[n_chunk++ & 1];
            cudaError_t e = S.busy ? cudaEventSynchronize(S.done) : cudaSuccess;
            if (e == cudaSuccess) {
                // the staging set is filled by LOAD_IO_THREADS host threads: one thread moved ~5 GB/s out of the page cache,
                // which made the host side (not PCIe, not the insert kernel) the limit of the load (36 GB of files: 6 s)
                std::atomic<int> short_file{0};
                parallel_for_threads(LOAD_IO_THREADS, [&](int t) {
                    auto slice = [&](uint64_t n, uint64_t& lo, uint64_t& hi) { lo = n * (uint64_t)t / LOAD_IO_THREADS; hi = n * (uint64_t)(t + 1) / LOAD_IO_THREADS; };
                    uint64_t lo, hi;
                    slice(r1 - r0, lo, hi);
                    memcpy(S.h_sz + lo, sz + r0 + lo, hi - lo);
                    if (sfactor > 1) memcpy(S.h_keep + lo, keep.data() + r0 + lo, hi - lo);
                    if (t == 0) for (uint64_t b = b0; b <= b1; b++) S.h_rel[b - b0] = coarse[b] - e0;
                    slice(ne, lo, hi);                       // whole entries per thread
                    if (base_path) {
                        if (!pread_all(fs.ky, S.h_keys + lo * kb, (hi - lo) * kb, (e0 + lo) * kb)) short_file |= 1;
                        if (!pread_all(fs.lb, S.h_labels + lo, (hi - lo) * 2, (e0 + lo) * 2)) short_file |= 2;
                    } else {
                        memcpy(S.h_keys + lo * kb, static_cast<const uint8_t*>(ky) + (e0 + lo) * kb, (hi - lo) * kb);
                        memcpy(S.h_labels + lo, lb + e0 + lo, (hi - lo) * 2);
                    }
                });
                if (short_file) { set_error("%s%s is short", base_path, (short_file & 1) ? ".ky" : ".lb"); rc = CUCLARK_ERR_IO; break; }
                auto cp = [&](void* dst, const void* src, size_t n) { return cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, S.st); };
                e = cp(S.d_sz, S.h_sz, r1 - r0);
                if (e == cudaSuccess && sfactor > 1) e = cp(S.d_keep, S.h_keep, r1 - r0);
                if (e == cudaSuccess) e = cp(S.d_rel, S.h_rel, (b1 - b0 + 1) * 8);
                if (e == cudaSuccess) e = cp(S.d_keys, S.h_keys, ne * kb);
                if (e == cudaSuccess) e = cp(S.d_labels, S.h_labels, ne * 2);
            }
            if (e != cudaSuccess) { set_error("DB upload failed: %s", cudaGetErrorString(e)); rc = CUCLARK_ERR_CUDA; break; }
            const unsigned blocks = (unsigned)(b1 - b0);
            if (g.layout == LAYOUT_NARROW)
                k_insert_chunk<LAYOUT_NARROW><<<blocks, CHUNK_THREADS, 0, S.st>>>(x, S.d_sz, S.d_keep, S.d_rel, r0, (uint32_t)(r1 - r0), S.d_keys, S.d_labels, kb);
            else if (g.layout == LAYOUT_LOCAL)
                k_insert_chunk<LAYOUT_LOCAL><<<blocks, CHUNK_THREADS, 0, S.st>>>(x, S.d_sz, S.d_keep, S.d_rel, r0, (uint32_t)(r1 - r0), S.d_keys, S.d_labels, kb);
            else
                k_insert_chunk<LAYOUT_WIDE><<<blocks, CHUNK_THREADS, 0, S.st>>>(x, S.d_sz, S.d_keep, S.d_rel, r0, (uint32_t)(r1 - r0), S.d_keys, S.d_labels, kb);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaEventRecord(S.done, S.st);
            if (e != cudaSuccess) { set_error("insert kernel failed: %s", cudaGetErrorString(e)); rc = CUCLARK_ERR_CUDA; break; }
            S.busy = true;
        }
        for (auto& S : stage) {
            if (!S.busy) continue;
            const cudaError_t e = cudaStreamSynchronize(S.st);
            S.busy = false;
            if (e != cudaSuccess && rc == CUCLARK_OK) { set_error("insert kernel failed: %s", cudaGetErrorString(e)); rc = CUCLARK_ERR_CUDA; }
        }
        if (rc == CUCLARK_OK) rc = finish_build(db, g, bb, x, false);
        if (rc != CUCLARK_OK) bb.free_all();
        if (is_auto_local && (rc == CUCLARK_ERR_NOMEM || rc == CUCLARK_ERR_CUDA)) { cudaGetLastError(); rc = CUCLARK_ERR_BUILD; }   // hashed retry
    }
    if (timing)
        fprintf(stderr, "[cuclark timing] database load: host pass over %llu bucket sizes %.1f ms on %d threads, "
                        "upload + device re-bucketing of %llu entries %.1f ms (table layout %d)\n",
                (unsigned long long)H, ms_scan, n_thr, (unsigned long long)kept, since(t_build), db->view.layout);
    return rc;
}

int table_build_synthetic(cuclark_db* db, uint32_t seed, uint32_t n_targets, uint64_t genome_len, int light_gap) {
    const int k = db->cfg.k;
    if (genome_len < (uint64_t)k) { set_error("genome shorter than k"); return CUCLARK_ERR_ARG; }
    uint64_t per_target, expected;
    if (light_gap > 0) {
        const uint64_t nk = genome_len / k;
        per_target = (nk + light_gap - 1) / light_gap;
    } else {
        per_target = genome_len - k + 1;
    }
    expected = per_target * n_targets;
    int rc = CUCLARK_ERR_BUILD;
    double grow = 1.0;
    bool auto_local = true, grow_next = false;
    for (int attempt = 0; attempt < 6 && rc == CUCLARK_ERR_BUILD; attempt++) {
        if (grow_next) grow *= 1.3;
        const Geometry g = choose_geometry(db->cfg, expected, grow, auto_local);
        const bool is_auto_local = db->cfg.layout == 0 && g.layout == LAYOUT_LOCAL;
        grow_next = !is_auto_local;
        if (is_auto_local) auto_local = false;
        BuildBuffers bb;
        BuildCtx x{};
        rc = alloc_build(g, expected / (db->cfg.shard_count > 1 ? db->cfg.shard_count : 1), bb, x, db->cfg.htsize, db->cfg.k);
        if (rc != CUCLARK_OK) { bb.free_all(); if (is_auto_local && rc != CUCLARK_ERR_ARG) { cudaGetLastError(); rc = CUCLARK_ERR_BUILD; continue; } break; }
        if (light_gap > 0) {
            const uint64_t n = per_target * n_targets;
            const unsigned blocks = (unsigned)((n + 255) / 256);
            if (g.layout == LAYOUT_NARROW) k_synth_insert_light<LAYOUT_NARROW><<<blocks, 256>>>(x, seed, n_targets, genome_len, k, light_gap, per_target);
            else if (g.layout == LAYOUT_LOCAL) k_synth_insert_light<LAYOUT_LOCAL><<<blocks, 256>>>(x, seed, n_targets, genome_len, k, light_gap, per_target);
            else k_synth_insert_light<LAYOUT_WIDE><<<blocks, 256>>>(x, seed, n_targets, genome_len, k, light_gap, per_target);
        } else {
            const uint64_t runs = (per_target + 63) / 64;
            const uint64_t n = runs * n_targets;
            const unsigned blocks = (unsigned)((n + 127) / 128);
            if (g.layout == LAYOUT_NARROW) k_synth_insert_full<LAYOUT_NARROW><<<blocks, 128>>>(x, seed, n_targets, genome_len, k, runs);
            else if (g.layout == LAYOUT_LOCAL) k_synth_insert_full<LAYOUT_LOCAL><<<blocks, 128>>>(x, seed, n_targets, genome_len, k, runs);
            else k_synth_insert_full<LAYOUT_WIDE><<<blocks, 128>>>(x, seed, n_targets, genome_len, k, runs);
        }
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { set_error("synthetic insert failed: %s", cudaGetErrorString(e)); bb.free_all(); return CUCLARK_ERR_CUDA; }
        rc = finish_build(db, g, bb, x, true);
        if (rc != CUCLARK_OK) bb.free_all();
        if (is_auto_local && (rc == CUCLARK_ERR_NOMEM || rc == CUCLARK_ERR_CUDA)) { cudaGetLastError(); rc = CUCLARK_ERR_BUILD; }   // hashed retry
    }
    return rc;
}

// ---- table cache: the device layout as one file --------------------------------------------------
// SURVEY.md section 8(f)-3. The reference loader reads .sz/.ky/.lb serially into pinned memory and
// rebuilds its bucket pointers on every start (src/CuClarkDB.cu:462-808); here the re-bucketed table
// can be written once and streamed back: file -> pinned double buffer -> H2D, no rebuild.
namespace {

constexpr char CACHE_MAGIC[8] = {'C', 'U', 'C', 'B', '2', 'T', 'B', 'L'};
constexpr uint32_t CACHE_VERSION = 4;   // 2: LOCAL tables have two candidate lines per minimizer; 3: the second one within 32 KB of the first;
                                        // 4: source files identified by size AND modification time, LOCAL block size in the header
constexpr size_t CACHE_IO_BYTES = 64ull << 20;

struct CacheHeader {                 // 192 bytes, little endian
    char magic[8];
    uint32_t version, header_bytes;
    uint32_t k, layout, n_targets, key_bytes, shard_index, shard_count, sfactor, reserved;
    uint64_t htsize, M, lo, n_local, n_ovf, n_entries, n_spilled, n_spill_buckets;
    uint64_t src_bytes[3];           // sizes of the .sz/.ky/.lb the table was built from (0: not from files)
    uint64_t checksum;               // wrapping sum of the payload's 64-bit words times their position parity
    uint64_t src_mtime_ns[3];        // modification times of those files: a database rebuilt in place keeps its sizes
    uint64_t local_alt_block;        // LOCAL: lines per block of the second candidate line (compile-time geometry)
    uint64_t pad[2];
};
static_assert(sizeof(CacheHeader) == 192, "cache header layout");

// order-sensitive checksum of a device buffer of 64-bit words: sum of w[i] * (2*(i mod 2^20) + 1)
__global__ void k_checksum(const uint64_t* __restrict__ w, uint64_t n, unsigned long long* out) {
    uint64_t acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += gridDim.x * (uint64_t)blockDim.x)
        acc += w[i] * (2 * (i & 0xFFFFFu) + 1);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, (unsigned long long)acc);
}

int device_checksum(cuclark_db* db, const uint4* table, uint64_t n_table, const uint4* ovf, uint64_t n_ovf, uint64_t* out) {
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, 16));
    CK(cudaMemset(d, 0, 16));
    const int blocks = db->sm_count * 8;
    if (n_table) k_checksum<<<blocks, 256>>>(reinterpret_cast<const uint64_t*>(table), n_table * 4, d);
    if (n_ovf) k_checksum<<<blocks, 256>>>(reinterpret_cast<const uint64_t*>(ovf), n_ovf * 4, d + 1);
    unsigned long long h[2];
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("checksum kernel failed: %s", cudaGetErrorString(e)); return CUCLARK_ERR_CUDA; }
    *out = h[0] ^ (h[1] * 0x9E3779B97F4A7C15ull);
    return CUCLARK_OK;
}

// Table <-> file through CACHE_IO_THREADS host threads, each with its own pinned buffer and stream: piece i of the
// payload belongs to thread i mod T, which reads it (pread) and copies it up (or copies it down and writes it) while
// the other threads do the same for their pieces. One thread moved 5 GB/s out of /dev/shm (r01_cache_bench.json);
// the PCIe link takes ~55 GB/s, so the file side needs the parallelism.
constexpr int CACHE_IO_THREADS = 12;

bool pwrite_all(int fd, const void* src, size_t n, uint64_t off) {
    const uint8_t* p = static_cast<const uint8_t*>(src);
    while (n) {
        const ssize_t w = pwrite(fd, p, n, (off_t)off);
        if (w <= 0) return false;
        p += w; n -= (size_t)w; off += (uint64_t)w;
    }
    return true;
}

// to_device: file[file_off ..) -> d_buf; else d_buf -> file[file_off ..)
int stream_file(int device, int fd, uint64_t file_off, void* d_buf, uint64_t bytes, bool to_device, const char* path) {
    const uint64_t n_pieces = (bytes + CACHE_IO_BYTES - 1) / CACHE_IO_BYTES;
    if (!n_pieces) return CUCLARK_OK;
    const int T = (int)std::min<uint64_t>(CACHE_IO_THREADS, n_pieces);
    std::vector<int> rcs(T, CUCLARK_OK);
    std::vector<std::string> errs(T);
    parallel_for_threads(T, [&](int t) {
        auto fail = [&](int rc, const std::string& m) { rcs[t] = rc; errs[t] = m; };
        if (cudaSetDevice(device) != cudaSuccess) return fail(CUCLARK_ERR_CUDA, "cudaSetDevice failed");
        void* h = nullptr;
        cudaStream_t st = nullptr;
        if (cudaMallocHost(&h, CACHE_IO_BYTES) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            if (h) cudaFreeHost(h);
            return fail(CUCLARK_ERR_NOMEM, "pinned staging buffer for the table cache");
        }
        uint8_t* dev = static_cast<uint8_t*>(d_buf);
        for (uint64_t i = (uint64_t)t; i < n_pieces && rcs[t] == CUCLARK_OK; i += (uint64_t)T) {
            const uint64_t off = i * CACHE_IO_BYTES, n = std::min<uint64_t>(CACHE_IO_BYTES, bytes - off);
            if (to_device) {
                if (!pread_all(fd, h, n, file_off + off)) { fail(CUCLARK_ERR_IO, std::string(path) + " is truncated"); break; }
                if (cudaMemcpyAsync(dev + off, h, n, cudaMemcpyHostToDevice, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
                    fail(CUCLARK_ERR_CUDA, "H2D copy of the table cache failed");
            } else {
                if (cudaMemcpyAsync(h, dev + off, n, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
                    fail(CUCLARK_ERR_CUDA, "D2H copy of the table failed");
                    break;
                }
                if (!pwrite_all(fd, h, n, file_off + off)) fail(CUCLARK_ERR_IO, std::string("write to ") + path + " failed");
            }
        }
        cudaGetLastError();
        cudaStreamDestroy(st);
        cudaFreeHost(h);
    });
    for (int t = 0; t < T; t++)
        if (rcs[t] != CUCLARK_OK) { set_error("%s", errs[t].c_str()); return rcs[t]; }
    return CUCLARK_OK;
}

// size and modification time (ns) of a file; zeros if it cannot be stat'ed
void file_stamp(const std::string& p, uint64_t& size, uint64_t& mtime_ns) {
    struct stat sb;
    size = mtime_ns = 0;
    if (stat(p.c_str(), &sb) != 0) return;
    size = (uint64_t)sb.st_size;
    mtime_ns = (uint64_t)sb.st_mtim.tv_sec * 1000000000ull + (uint64_t)sb.st_mtim.tv_nsec;
}

}  // namespace

int table_save(cuclark_db* db, const char* path) {
    if (!db->d_table) { set_error("no database loaded"); return CUCLARK_ERR_STATE; }
    CacheHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, CACHE_MAGIC, 8);
    h.version = CACHE_VERSION; h.header_bytes = sizeof h;
    h.k = db->cfg.k; h.layout = db->view.layout; h.n_targets = db->cfg.n_targets; h.key_bytes = db->key_bytes;
    h.shard_index = db->cfg.shard_index; h.shard_count = db->cfg.shard_count; h.sfactor = db->src_sfactor;
    h.htsize = db->cfg.htsize; h.M = db->view.M; h.lo = db->view.lo; h.n_local = db->view.n_local; h.n_ovf = db->view.n_ovf;
    h.n_entries = db->n_entries; h.n_spilled = db->n_spilled; h.n_spill_buckets = db->n_spill_buckets;
    for (int i = 0; i < 3; i++) { h.src_bytes[i] = db->src_bytes[i]; h.src_mtime_ns[i] = db->src_mtime_ns[i]; }
    h.local_alt_block = LOCAL_ALT_BLOCK;
    CK(cudaDeviceSynchronize());
    int rc = device_checksum(db, db->d_table, h.n_local, db->d_ovf, h.n_ovf, &h.checksum);
    if (rc) return rc;
    const std::string tmp = std::string(path) + ".tmp";
    const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) { set_error("Failed to open %s for writing", tmp.c_str()); return CUCLARK_ERR_IO; }
    if (!pwrite_all(fd, &h, sizeof h, 0)) { set_error("write to %s failed", tmp.c_str()); rc = CUCLARK_ERR_IO; }
    if (rc == CUCLARK_OK) rc = stream_file(db->cfg.device, fd, sizeof h, db->d_table, h.n_local * 32, false, tmp.c_str());
    if (rc == CUCLARK_OK && h.n_ovf) rc = stream_file(db->cfg.device, fd, sizeof h + h.n_local * 32, db->d_ovf, h.n_ovf * 32, false, tmp.c_str());
    if (close(fd) != 0 && rc == CUCLARK_OK) { set_error("write to %s failed", tmp.c_str()); rc = CUCLARK_ERR_IO; }
    if (rc == CUCLARK_OK && rename(tmp.c_str(), path) != 0) { set_error("cannot rename %s to %s", tmp.c_str(), path); rc = CUCLARK_ERR_IO; }
    if (rc != CUCLARK_OK) remove(tmp.c_str());
    return rc;
}

int table_load(cuclark_db* db, const char* path, const char* src_base, int sfactor) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { set_error("Failed to open %s", path); return CUCLARK_ERR_IO; }
    struct Closer { int fd; ~Closer() { close(fd); } } closer{fd};
    CacheHeader h;
    if (!pread_all(fd, &h, sizeof h, 0) || memcmp(h.magic, CACHE_MAGIC, 8) != 0 || h.header_bytes != sizeof h) {
        set_error("%s is not a cuclark_b200 table cache", path); return CUCLARK_ERR_FORMAT;
    }
    if (h.version != CACHE_VERSION) { set_error("%s: cache version %u, this library reads %u", path, h.version, CACHE_VERSION); return CUCLARK_ERR_FORMAT; }
    const int want_s = sfactor > 1 ? sfactor : 1;
    if ((int)h.k != db->cfg.k || h.htsize != db->cfg.htsize || (int)h.n_targets != db->cfg.n_targets ||
        (int)h.shard_index != db->cfg.shard_index || (int)h.shard_count != db->cfg.shard_count || (int)h.sfactor != want_s ||
        (db->cfg.layout != 0 && (int)h.layout != db->cfg.layout)) {
        set_error("%s was written for another configuration (k=%u htsize=%llu targets=%u shard %u/%u s=%u layout %u)", path, h.k,
                  (unsigned long long)h.htsize, h.n_targets, h.shard_index, h.shard_count, h.sfactor, h.layout);
        return CUCLARK_ERR_FORMAT;
    }
    // geometry: everything the kernels index with comes from here, so it is checked before anything is allocated
    // (n_local < 2^32 and n_ovf < 2^40 also keep the byte counts below far from wrapping)
    if ((h.layout != LAYOUT_NARROW && h.layout != LAYOUT_WIDE && h.layout != LAYOUT_LOCAL) || h.M == 0 || h.n_local == 0 || h.n_local >= 0xFFFFFFFFull ||
        h.n_ovf >= (1ull << 40) || h.lo > h.M || h.n_local > h.M - h.lo ||
        (h.layout == LAYOUT_NARROW && (h.k >= 32 || pow4((int)h.k) / h.M >= 0xFFFFFFFFull)) ||
        (h.layout == LAYOUT_LOCAL && ((h.M & 3) || (h.lo & 3) || (h.n_local & 3) || (int)h.k < LOCAL_MIN_K || h.M / 4 < local_min_lines((int)h.k) ||
                                      h.local_alt_block != LOCAL_ALT_BLOCK))) {
        set_error("%s: inconsistent geometry", path); return CUCLARK_ERR_FORMAT;
    }
    if (src_base) {
        // the cache belongs to the files it was built from: same sizes AND same modification times (a database
        // rebuilt in place with as many k-mers keeps its sizes)
        const std::string b(src_base);
        const char* ext[3] = {".sz", ".ky", ".lb"};
        for (int i = 0; i < 3; i++) {
            uint64_t sz_i, mt_i;
            file_stamp(b + ext[i], sz_i, mt_i);
            if (sz_i != h.src_bytes[i] || mt_i != h.src_mtime_ns[i]) {
                set_error("%s does not belong to %s%s (size or modification time differs)", path, src_base, ext[i]); return CUCLARK_ERR_FORMAT;
            }
        }
    }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (uint64_t)sb.st_size != sizeof h + (h.n_local + h.n_ovf) * 32) { set_error("%s is truncated", path); return CUCLARK_ERR_FORMAT; }
    uint4 *table = nullptr, *ovf = nullptr;
    if (cudaMalloc(&table, h.n_local * 32) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc of %.2f GB table failed", h.n_local * 32 / 1e9); return CUCLARK_ERR_NOMEM; }
    if (h.n_ovf && cudaMalloc(&ovf, h.n_ovf * 32) != cudaSuccess) { cudaGetLastError(); cudaFree(table); set_error("cudaMalloc of overflow table failed"); return CUCLARK_ERR_NOMEM; }
    int rc = stream_file(db->cfg.device, fd, sizeof h, table, h.n_local * 32, true, path);
    if (rc == CUCLARK_OK && h.n_ovf) rc = stream_file(db->cfg.device, fd, sizeof h + h.n_local * 32, ovf, h.n_ovf * 32, true, path);
    uint64_t sum = 0;
    if (rc == CUCLARK_OK) rc = device_checksum(db, table, h.n_local, ovf, h.n_ovf, &sum);
    if (rc == CUCLARK_OK && sum != h.checksum) { set_error("%s: checksum mismatch (file is corrupt)", path); rc = CUCLARK_ERR_FORMAT; }
    if (rc != CUCLARK_OK) { cudaFree(table); cudaFree(ovf); return rc; }
    db->d_table = table; db->d_ovf = ovf;
    db->view = TableView{};
    db->view.buckets = table; db->view.ovf = ovf;
    db->view.M = h.M; db->view.magic = (uint64_t)((((__uint128_t)1) << 64) / h.M);
    db->view.lo = h.lo; db->view.n_local = h.n_local; db->view.n_ovf = h.n_ovf;
    db->view.layout = (int)h.layout; db->view.k = db->cfg.k;
    if (h.layout == LAYOUT_LOCAL) { db->view.NL = h.M / 4; db->view.magicNL = (uint64_t)((((__uint128_t)1) << 64) / db->view.NL); set_local_divmod(db->view); db->view.line_lo = (uint32_t)(h.lo >> 2); db->view.line_n = (uint32_t)(h.n_local >> 2); }
    db->n_entries = h.n_entries; db->n_spilled = h.n_spilled; db->n_spill_buckets = h.n_spill_buckets;
    db->src_sfactor = (int)h.sfactor;
    for (int i = 0; i < 3; i++) { db->src_bytes[i] = h.src_bytes[i]; db->src_mtime_ns[i] = h.src_mtime_ns[i]; }
    return CUCLARK_OK;
}

}  // namespace cuclark
