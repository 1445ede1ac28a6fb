// cuclark_b200 — shared device/host definitions.
//
// Device table layout ("sector buckets"), designed for one 32-byte HBM sector
// per probe. The reference keeps three arrays (bucket pointers, keys, labels;
// src/CuClarkDB.cu:1268-1313) and touches >= 3 sectors per lookup; here every
// canonical k-mer c has exactly one home bucket of 32 bytes:
//
//     b   = c mod M            (M = number of device buckets, our own modulus)
//     key = c div M            (exact quotient: no lossy fingerprint)
//
//   NARROW (key fits 32 bit, i.e. M > 4^k / (2^32-1)):
//     word 0..4   key[0..4]            (0xFFFFFFFF = empty)
//     word 5      label[0] | label[1] << 16
//     word 6      label[2] | label[3] << 16
//     word 7      label[4] | meta << 16
//   WIDE (any k <= 32, any M; used for small tables and the overflow table):
//     word 0..5   key[0..2] as uint64  (all-ones = empty)
//     word 6      label[0] | label[1] << 16
//     word 7      label[2] | meta << 16
//
//   LOCAL (k >= 19, large tables): the home of a k-mer is chosen by its canonical MINIMIZER
//     (the m-mer, m = k-7, with the smallest hash among the 8 m-mers of the k-mer; both
//     strands give the same one), so that the ~4.5 consecutive k-mers of a read that share a
//     minimizer probe the SAME 128-byte line: line = mix(minimizer) mod NL, sector within the
//     line = minimizer offset & 3 (consecutive k-mers -> consecutive offsets -> the group is
//     spread evenly over the 4 sectors of its line). Adjacent lanes of a warp then coalesce
//     into one line request instead of one random sector each. The slot still identifies the
//     k-mer exactly: key = mix(minimizer) div NL (<= 19 bit) | the 7 other nucleotides (14 bit)
//     | offset (3 bit) | strand of the minimizer (1 bit) = 37 bit.
//     TWO candidate lines per minimizer (A = mix mod NL, B = A + step(key) inside the shard): an entry
//     goes to the emptier one; a lookup loads the sector of both (independent loads, no second round
//     trip) and a bit of the key tells an entry living in its B line from one living in its A line.
//     Lines fill in clumps of ~4.5 entries, one choice would overflow ~16% of the entries.
//     word 0..3   key[0..3] low 32 bits      (all-ones + high byte 0xFF = empty)
//     word 4      label[0] | label[1] << 16
//     word 5      label[2] | label[3] << 16
//     word 6      key high bits, one byte per slot
//     word 7      meta << 16
//
//   meta bit 0 = "overflowed": more k-mers are homed here than the bucket has
//   slots; the excess lives in the OVERFLOW table, a second, sparsely filled
//   array of WIDE buckets keyed by the full canonical k-mer, addressed by an
//   independent hash and resolved by linear probing over buckets (a bucket with
//   a free slot ends the probe sequence; removed entries leave a tombstone).
//   A lookup reads its home sector and, only when the key is not there AND the
//   bucket is flagged, continues in the overflow table. At ~2.5 entries per
//   5-slot bucket ~4% of the buckets are flagged: ~1.04 sectors per lookup for
//   hits and misses alike.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace cuclark {

constexpr int LAYOUT_NARROW = 1;
constexpr int LAYOUT_WIDE = 2;
constexpr int LAYOUT_LOCAL = 3;
constexpr int NARROW_SLOTS = 5;
constexpr int WIDE_SLOTS = 3;
constexpr int LOCAL_SLOTS = 4;
constexpr int LOCAL_W = 8;             // m-mers per k-mer (m = k - LOCAL_W + 1)
constexpr int LOCAL_ZQ_BITS = 19;      // bits of mix(minimizer) div NL kept in the key
constexpr int LOCAL_REST_BITS = 2 * (LOCAL_W - 1);
constexpr int LOCAL_MIN_K = 19;        // m >= 12: the order hash takes the top 24 bits of a >= 24-bit value
constexpr uint32_t LOCAL_ORDER_MAX = 0xFFFFFEu;    // 24-bit order hash, all-ones reserved for "no m-mer"
constexpr uint32_t NO_LABEL = 0xFFFFFFFFu;
constexpr uint64_t OVF_EMPTY = ~0ull;
constexpr uint64_t OVF_TOMBSTONE = ~0ull - 1;
constexpr uint64_t OVF_HASH_MULT = 0x9E3779B97F4A7C15ull;

struct TableView {
    const uint4* buckets;   // 2 x uint4 per bucket, 32-byte aligned
    const uint4* ovf;       // overflow table (WIDE buckets, full k-mers), may be null
    uint64_t M;             // global number of buckets (the modulus)
    uint64_t magic;         // floor(2^64 / M)
    uint64_t lo;            // this shard holds home buckets [lo, lo+n_local)
    uint64_t n_local;
    uint64_t n_ovf;         // buckets in the overflow table
    uint64_t NL;            // LOCAL: number of 128-byte lines (M = 4 NL), and floor(2^64 / NL)
    uint64_t magicNL;
    uint32_t nl_m32;        // LOCAL: floor(2^(32 + nl_sh) / NL) for local_divmod()
    int nl_sh;              //        max(0, 2m - 32)
    uint32_t line_lo, line_n;   // LOCAL: this shard's lines [line_lo, line_lo + line_n)
    int layout;
    int k;
};

// ---- k-mer arithmetic -------------------------------------------------------
// Reverse complement of a 2k-bit code, as src/CuClarkDB.cu:1255-1263 defines it
// (reverse the 2-bit groups, complement, shift down): brev reverses all 64
// bits, the swap puts each pair back in order.
__host__ __device__ __forceinline__ uint64_t revcomp2(uint64_t x, int k) {
#ifdef __CUDA_ARCH__
    // on the two 32-bit halves: reverse the bits (and the halves), swap the bits of every pair, complement
    uint32_t rl = __brev((uint32_t)(x >> 32)), rh = __brev((uint32_t)x);
    rl = ~(((rl >> 1) & 0x55555555u) | ((rl & 0x55555555u) << 1));
    rh = ~(((rh >> 1) & 0x55555555u) | ((rh & 0x55555555u) << 1));
    return (((uint64_t)rh << 32) | rl) >> (64 - 2 * k);
#else
    uint64_t r = x;
    r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
    r = ((r >> 2) & 0x3333333333333333ull) | ((r & 0x3333333333333333ull) << 2);
    r = ((r >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((r & 0x0F0F0F0F0F0F0F0Full) << 4);
    r = ((r >> 8) & 0x00FF00FF00FF00FFull) | ((r & 0x00FF00FF00FF00FFull) << 8);
    r = ((r >> 16) & 0x0000FFFF0000FFFFull) | ((r & 0x0000FFFF0000FFFFull) << 16);
    r = (r >> 32) | (r << 32);
    r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
    return (~r) >> (64 - 2 * k);
#endif
}

__host__ __device__ __forceinline__ uint64_t canonical(uint64_t x, int k) {
    uint64_t r = revcomp2(x, k);
    return x < r ? x : r;
}

// c = q*M + r with a precomputed floor(2^64/M); c < 2^64, exact after one fix-up
// (the estimate is q or q-1 because c * (2^64 mod M) / M < 2^64).
__device__ __forceinline__ void divmod_M(uint64_t c, uint64_t M, uint64_t magic, uint64_t& q, uint64_t& r) {
    q = __umul64hi(c, magic);
    r = c - q * M;
    if (r >= M) { r -= M; q++; }
}

// ---- LOCAL layout: minimizer-addressed lines ---------------------------------------
// Invertible mixing of an nbits-wide value (24 <= nbits <= 50): the xor-shift by ceil(nbits/2) is an
// involution, the multiplier is odd, so local_unmix() undoes it exactly.
constexpr uint64_t LOCAL_MUL = 0xD6E8FEB86659FD93ull;

__host__ __device__ __forceinline__ uint64_t local_mix(uint64_t u, int nbits) {
    // fold the upper half into the lower one, then multiply: the top bits of the product (the order of the
    // m-mers) and its low bits (the line) both depend on every bit of u
    const uint64_t mask = (~0ull) >> (64 - nbits);
    u ^= (uint32_t)(u >> ((nbits + 1) >> 1));       // the shifted value has at most 32 bits (nbits <= 64)
    return (u * LOCAL_MUL) & mask;
}
__host__ __device__ __forceinline__ uint64_t inv_odd64(uint64_t a) {   // a * x == 1 mod 2^64
    uint64_t x = a;
    for (int i = 0; i < 6; i++) x *= 2 - a * x;
    return x;
}
__host__ __device__ __forceinline__ uint64_t local_unmix(uint64_t z, int nbits) {
    const uint64_t mask = (~0ull) >> (64 - nbits);
    z = (z * inv_odd64(LOCAL_MUL)) & mask;
    z ^= z >> ((nbits + 1) >> 1);
    return z;
}
// order of the m-mers inside a k-mer: the top 24 bits of the mixed canonical m-mer
__host__ __device__ __forceinline__ uint32_t local_order(uint64_t z, int nbits) {
    const uint32_t h = (uint32_t)(z >> (nbits - 24));
    return h > LOCAL_ORDER_MAX ? LOCAL_ORDER_MAX : h;
}
// (zq, line) = divmod(z, NL) for z < 4^m in 32-bit arithmetic: the quotient estimate from the top 32 bits of z and
// m32 = floor(2^(32+sh) / NL) is at most 2 too small (NL >= 2^(sh+13)), and 3 NL < 2^32, so the remainder lives in 32 bits
__host__ __device__ __forceinline__ void local_divmod(uint64_t z, uint32_t NL, uint32_t m32, int sh, uint32_t& zq, uint32_t& line) {
#ifdef __CUDA_ARCH__
    uint32_t q = __umulhi((uint32_t)(z >> sh), m32);
#else
    uint32_t q = (uint32_t)(((uint64_t)(uint32_t)(z >> sh) * m32) >> 32);
#endif
    uint32_t r = (uint32_t)z - q * NL;
    if (r >= NL) { r -= NL; q++; }
    if (r >= NL) { r -= NL; q++; }
    zq = q;
    line = r;
}
// canonical m-mer at offset o of the 2k-bit code y (first nucleotide in the high bits);
// fwd tells whether the form standing in y is the strictly smaller one
__host__ __device__ __forceinline__ uint64_t local_mmer(uint64_t y, int o, int m, bool& fwd) {
    const uint64_t a = (y >> (2 * (LOCAL_W - 1 - o))) & ((~0ull) >> (64 - 2 * m));
    const uint64_t b = revcomp2(a, m);
    fwd = a < b;
    return fwd ? a : b;
}
// the 2(W-1) bits of c around its m-mer at offset o
__host__ __device__ __forceinline__ uint32_t local_rest(uint64_t c, int o, int m) {
    const uint32_t lomask = (1u << (2 * (LOCAL_W - 1 - o))) - 1u;
    return ((uint32_t)(c >> (2 * m)) & ~lomask) | ((uint32_t)c & lomask);
}
// key = zq | rest << 19 | o << 33 | fwd << 36, as its low word and its high byte
__host__ __device__ __forceinline__ uint32_t local_key_lo(uint64_t zq, uint32_t rest) { return (uint32_t)zq | (rest << LOCAL_ZQ_BITS); }
__host__ __device__ __forceinline__ uint32_t local_key_hi(uint32_t rest, int o, bool fwd) {
    return (rest >> (32 - LOCAL_ZQ_BITS)) | ((uint32_t)o << 1) | ((uint32_t)fwd << 4);
}
// Home of the canonical k-mer c: its minimizer is the m-mer with the smallest order hash,
// the LEFTMOST such offset in c's own orientation on ties. Returns the global sector index
// (line * 4 + (offset & 3)) and the 37-bit key. Reference form: the classify kernel computes
// the same thing with a rolling window minimum across lanes.
//
// TIES. The kernel takes the leftmost smallest hash in READ orientation, which is the rightmost one of the
// canonical k-mer when the read shows the other strand. A k-mer whose smallest order hash occurs at more than
// one offset (the same m-mer twice: low-complexity sequence; or two m-mers colliding in 24 bits) therefore has
// two possible homes. Such k-mers are never stored in the lines: the builder puts them into the overflow table and
// flags the sectors of BOTH homes, so whichever one a lookup computes, it finds the flags and goes to the overflow
// table. local_locate_both() reports the second home (rightmost offset) and whether the k-mer is such a tie.
__host__ __device__ __forceinline__ void local_home_at(uint64_t c, int m, uint64_t NL, int o, uint64_t& sector, uint64_t& key) {
    bool fwd;
    const uint64_t z = local_mix(local_mmer(c, o, m, fwd), 2 * m);
    const uint64_t line = z % NL, zq = z / NL;
    const uint32_t rest = local_rest(c, o, m);
    sector = line * 4 + (uint64_t)(o & 3);
    key = (uint64_t)local_key_lo(zq, rest) | ((uint64_t)local_key_hi(rest, o, fwd) << 32);
}
__host__ __device__ __forceinline__ bool local_min_offsets(uint64_t c, int k, int& o_left, int& o_right) {
    const int m = k - LOCAL_W + 1;
    uint32_t best = 0xFFFFFFFFu;
    o_left = o_right = 0;
    for (int o = 0; o < LOCAL_W; o++) {
        bool fwd;
        const uint32_t h = local_order(local_mix(local_mmer(c, o, m, fwd), 2 * m), 2 * m);
        if (h < best) { best = h; o_left = o_right = o; }
        else if (h == best) o_right = o;
    }
    return o_left != o_right;
}
__host__ __device__ __forceinline__ void local_locate(uint64_t c, int k, uint64_t NL, uint64_t& sector, uint64_t& key) {
    int ol, orr;
    local_min_offsets(c, k, ol, orr);
    local_home_at(c, k - LOCAL_W + 1, NL, ol, sector, key);
}
__host__ __device__ __forceinline__ bool local_locate_both(uint64_t c, int k, uint64_t NL, uint64_t& sec_l, uint64_t& key_l,
                                                           uint64_t& sec_r, uint64_t& key_r) {
    int ol, orr;
    const bool tie = local_min_offsets(c, k, ol, orr);
    local_home_at(c, k - LOCAL_W + 1, NL, ol, sec_l, key_l);
    local_home_at(c, k - LOCAL_W + 1, NL, orr, sec_r, key_r);
    return tie;
}
// B sector (global index) of a home whose A sector is sec_a; sec_a itself when the A line is outside the shard
__host__ __device__ __forceinline__ uint64_t local_alt_sector(uint64_t sec_a, uint64_t key, uint64_t line_lo, uint32_t line_n);
// ---- second candidate line -----------------------------------------------------------
constexpr uint32_t LOCAL_ALT_BIT = 1u << 5;            // in the key's high byte: the entry lives in its B line
// The B line is the A line rotated inside its block of LOCAL_ALT_BLOCK lines by a step that is a function of the part of the mixed
// minimizer the key keeps (zq), so that the A line is recoverable from (B line, key).
// The block is 256 lines = 32 KB: measured on an 8 G-entry table (96 GB, past the reach of the TLBs) 38 G lookups/s with
// the B line anywhere in the table, 46 with it in the same 2 MB page, 53 within 128 KB, 58 within 32 KB or 16 KB; the
// smaller the block the less even the fill (overflowed entries 4.87 % -> 4.91 % at 32 KB, 5.2 % at 4 KB).
#ifndef CUCLARK_ALT_BLOCK_LINES
#define CUCLARK_ALT_BLOCK_LINES 256
#endif
constexpr uint32_t LOCAL_ALT_BLOCK = CUCLARK_ALT_BLOCK_LINES;   // power of two
__host__ __device__ __forceinline__ uint32_t local_mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// A line -> B line (inverse = false) or B line -> A line (inverse = true), as indices relative to the shard's
// first line: a rotation by step(zq) inside the block of the line (the last block may be shorter)
__host__ __device__ __forceinline__ uint32_t local_alt_rel(uint32_t rel, uint32_t zq, uint32_t line_n, bool inverse) {
    const uint32_t start = rel & ~(LOCAL_ALT_BLOCK - 1u);
    const uint32_t size = line_n - start < LOCAL_ALT_BLOCK ? line_n - start : LOCAL_ALT_BLOCK;
    if (size < 2) return rel;
    const uint32_t step = 1u + local_mulhi32(zq * 0x9E3779B1u + 0x7F4A7C15u, size - 1u);     // 1 .. size-1
    uint32_t off = rel - start;
    off = inverse ? (off >= step ? off - step : off + size - step) : (off + step >= size ? off + step - size : off + step);
    return start + off;
}
// B line (global index) of the k-mer with A line `line_a` and key part zq; shard lines [line_lo, line_lo + line_n)
__host__ __device__ __forceinline__ uint64_t local_alt_line(uint64_t line_a, uint32_t zq, uint64_t line_lo, uint32_t line_n) {
    return line_lo + local_alt_rel((uint32_t)(line_a - line_lo), zq, line_n, false);
}
__host__ __device__ __forceinline__ uint64_t local_alt_line_inv(uint64_t line_b, uint32_t zq, uint64_t line_lo, uint32_t line_n) {
    return line_lo + local_alt_rel((uint32_t)(line_b - line_lo), zq, line_n, true);
}
__host__ __device__ __forceinline__ uint64_t local_alt_sector(uint64_t sec_a, uint64_t key, uint64_t line_lo, uint32_t line_n) {
    const uint64_t line_a = sec_a >> 2;
    if (line_a - line_lo >= line_n) return sec_a;
    return local_alt_line(line_a, (uint32_t)(key & ((1ull << LOCAL_ZQ_BITS) - 1)), line_lo, line_n) * 4 + (sec_a & 3);
}
// both candidate sectors (global indices) of c; sec_b == sec_a when the A line is not in this shard
__host__ __device__ __forceinline__ void local_locate2(uint64_t c, int k, uint64_t NL, uint64_t line_lo, uint32_t line_n,
                                                       uint64_t& sec_a, uint64_t& sec_b, uint64_t& key) {
    local_locate(c, k, NL, sec_a, key);
    sec_b = local_alt_sector(sec_a, key, line_lo, line_n);
}
// the k-mer stored as `key` (with or without LOCAL_ALT_BIT) in a sector of global line `line`
__host__ __device__ __forceinline__ uint64_t local_rebuild(uint64_t line, uint64_t key, int k, uint64_t NL);
__host__ __device__ __forceinline__ uint64_t local_rebuild2(uint64_t line, uint64_t key, int k, uint64_t NL, uint64_t line_lo, uint32_t line_n) {
    const uint64_t alt = (uint64_t)LOCAL_ALT_BIT << 32;
    if (key & alt) line = local_alt_line_inv(line, (uint32_t)(key & ((1ull << LOCAL_ZQ_BITS) - 1)), line_lo, line_n);
    return local_rebuild(line, key & ~alt, k, NL);
}

// inverse of local_locate: the canonical k-mer stored as `key` in (a sector of) `line`
__host__ __device__ __forceinline__ uint64_t local_rebuild(uint64_t line, uint64_t key, int k, uint64_t NL) {
    const int m = k - LOCAL_W + 1;
    const uint64_t zq = key & ((1ull << LOCAL_ZQ_BITS) - 1);
    const uint64_t rest = (key >> LOCAL_ZQ_BITS) & ((1ull << LOCAL_REST_BITS) - 1);
    const int o = (int)((key >> (LOCAL_ZQ_BITS + LOCAL_REST_BITS)) & 7u);
    const bool fwd = (key >> (LOCAL_ZQ_BITS + LOCAL_REST_BITS + 3)) & 1u;
    const uint64_t u = local_unmix(zq * NL + line, 2 * m);
    const uint64_t mm = fwd ? u : revcomp2(u, m);            // !fwd: the larger (or the equal) form stands in c
    const int lo_bits = 2 * (LOCAL_W - 1 - o);
    const uint64_t lo = rest & ((1ull << lo_bits) - 1), hi = rest >> lo_bits;
    return (o ? hi << (2 * m + lo_bits) : 0) | (mm << lo_bits) | lo;
}

// ---- 32-byte probe ------------------------------------------------------------
struct Sector { uint32_t w[8]; };

__device__ __forceinline__ Sector load_sector(const uint4* p) {
    Sector s;
    // .L2::64B: on B200 a plain sector miss fills the whole 128-byte line from HBM (ncu: 128 B of
    // DRAM reads per probe); with this qualifier the fill is 64 B, the smallest the ISA offers.
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(s.w[0]), "=r"(s.w[1]), "=r"(s.w[2]), "=r"(s.w[3]),
                   "=r"(s.w[4]), "=r"(s.w[5]), "=r"(s.w[6]), "=r"(s.w[7])
                 : "l"(p));
    return s;
}

// Match `key` among the slots of a sector. Returns the label or NO_LABEL.
// Empty slots hold an all-ones key, which no valid quotient equals (M is chosen
// so that quotients stay below it).
template <int LAYOUT>
__device__ __forceinline__ uint32_t match_sector(const Sector& s, uint64_t key) {
    uint32_t label = NO_LABEL;
    if (LAYOUT == LAYOUT_NARROW) {
        const uint32_t k32 = (uint32_t)key;
#pragma unroll
        for (int i = 0; i < NARROW_SLOTS; i++) {
            const uint32_t lw = s.w[5 + (i >> 1)];
            if (s.w[i] == k32) label = (i & 1) ? (lw >> 16) : (lw & 0xFFFFu);
        }
    } else if (LAYOUT == LAYOUT_LOCAL) {
        const uint32_t k32 = (uint32_t)key;
        const uint32_t H = s.w[6] ^ ((uint32_t)(key >> 32) * 0x01010101u);     // byte i is 0 iff the high byte of slot i matches
        if (((s.w[0] ^ k32) | __byte_perm(H, 0, 0x4440)) == 0u) label = s.w[4] & 0xFFFFu;
        if (((s.w[1] ^ k32) | __byte_perm(H, 0, 0x4441)) == 0u) label = s.w[4] >> 16;
        if (((s.w[2] ^ k32) | __byte_perm(H, 0, 0x4442)) == 0u) label = s.w[5] & 0xFFFFu;
        if (((s.w[3] ^ k32) | __byte_perm(H, 0, 0x4443)) == 0u) label = s.w[5] >> 16;
    } else {
#pragma unroll
        for (int i = 0; i < WIDE_SLOTS; i++) {
            const uint64_t kk = (uint64_t)s.w[2 * i] | ((uint64_t)s.w[2 * i + 1] << 32);
            const uint32_t lw = s.w[6 + (i >> 1)];
            if (kk == key) label = (i & 1) ? (lw >> 16) : (lw & 0xFFFFu);
        }
    }
    return label;
}

// LOCAL probes: the whole 128-byte line is wanted in L2 (the neighbouring lanes / the next
// round read its other sectors)
__device__ __forceinline__ Sector load_sector_line(const uint4* p) {
    Sector s;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(s.w[0]), "=r"(s.w[1]), "=r"(s.w[2]), "=r"(s.w[3]),
                   "=r"(s.w[4]), "=r"(s.w[5]), "=r"(s.w[6]), "=r"(s.w[7])
                 : "l"(p));
    return s;
}

__device__ __forceinline__ bool sector_overflowed(const Sector& s) { return (s.w[7] >> 16) & 1u; }

__host__ __device__ __forceinline__ uint64_t ovf_home(uint64_t c, uint64_t n_ovf) {
#ifdef __CUDA_ARCH__
    return __umul64hi(c * OVF_HASH_MULT, n_ovf);
#else
    return (uint64_t)(((__uint128_t)(c * OVF_HASH_MULT) * n_ovf) >> 64);
#endif
}

// Overflow table: linear probing over WIDE buckets holding the full k-mer, from bucket b on
// (n0 buckets of the sequence have been looked at already).
__device__ __forceinline__ uint32_t ovf_lookup_from(const TableView& t, uint64_t c, uint64_t b, uint64_t n0) {
    for (uint64_t n = n0; n < t.n_ovf; n++) {
        const Sector s = load_sector(t.ovf + 2 * b);
        const uint32_t label = match_sector<LAYOUT_WIDE>(s, c);
        if (label != NO_LABEL) return label;
        if (((uint64_t)s.w[4] | ((uint64_t)s.w[5] << 32)) == OVF_EMPTY) return NO_LABEL;
        if (++b == t.n_ovf) b = 0;
    }
    return NO_LABEL;
}
__device__ __forceinline__ uint32_t ovf_lookup(const TableView& t, uint64_t c) {
    uint64_t b = ovf_home(c, t.n_ovf);
    for (uint64_t n = 0; n < t.n_ovf; n++) {
        const Sector s = load_sector(t.ovf + 2 * b);
        const uint32_t label = match_sector<LAYOUT_WIDE>(s, c);
        if (label != NO_LABEL) return label;
        // a free slot ends the sequence (slots fill in order, so test the last one)
        if (((uint64_t)s.w[4] | ((uint64_t)s.w[5] << 32)) == OVF_EMPTY) return NO_LABEL;
        if (++b == t.n_ovf) b = 0;
    }
    return NO_LABEL;
}

// Full lookup of one canonical k-mer (used by the slow paths; the hot kernel
// inlines the same steps so that it can batch the home-sector loads).
// `fwd`: the read showed the k-mer in its canonical form. It only matters for LOCAL tie k-mers (two possible homes,
// possibly in two shards): the hot kernel takes the leftmost smallest hash in READ orientation, i.e. the rightmost
// one of the canonical form when the read shows the other strand — the slow paths must pick the same home, so
// that in a sharded table exactly one shard answers for such a k-mer whichever path a read takes on each shard.
template <int LAYOUT>
__device__ __forceinline__ uint32_t table_lookup(const TableView& t, uint64_t c, bool fwd = true) {
    uint64_t q, b;
    if (LAYOUT == LAYOUT_LOCAL) {
        uint64_t sa, sb;
        int ol, orr;
        local_min_offsets(c, t.k, ol, orr);
        local_home_at(c, t.k - LOCAL_W + 1, t.NL, fwd ? ol : orr, sa, q);
        sb = local_alt_sector(sa, q, t.line_lo, t.line_n);
        const uint64_t la = sa - t.lo;
        if (la >= t.n_local) return NO_LABEL;       // other shard
        const Sector A = load_sector(t.buckets + 2 * la), B = load_sector(t.buckets + 2 * (sb - t.lo));
        uint32_t label = match_sector<LAYOUT_LOCAL>(A, q);
        if (label == NO_LABEL) label = match_sector<LAYOUT_LOCAL>(B, q | ((uint64_t)LOCAL_ALT_BIT << 32));
        if (label == NO_LABEL && sector_overflowed(A) && sector_overflowed(B)) label = ovf_lookup(t, c);
        return label;
    }
    divmod_M(c, t.M, t.magic, q, b);
    const uint64_t lb = b - t.lo;
    if (lb >= t.n_local) return NO_LABEL;           // other shard (b < lo wraps around)
    const Sector s = load_sector(t.buckets + 2 * lb);
    uint32_t label = match_sector<LAYOUT>(s, q);
    if (label == NO_LABEL && sector_overflowed(s)) label = ovf_lookup(t, c);
    return label;
}

// ---- error handling -------------------------------------------------------------
void set_error(const char* fmt, ...);
#define CK(call)                                                                       \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            cuclark::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                               __FILE__, __LINE__);                                    \
            return CUCLARK_ERR_CUDA;                                                   \
        }                                                                              \
    } while (0)

}  // namespace cuclark
