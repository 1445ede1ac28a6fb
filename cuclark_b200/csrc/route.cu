// cuclark_b200 — table-partitioned classification by ROUTING K-MERS to the shard that holds them.
//
// The reference's multi-GPU mode partitions the table by bucket range, sends every read batch to every device
// (src/CuClarkDB.cu:546-574, 886-895) and merges the per-device sparse rows pairwise (:953-974, mergeKernel
// :1321-1415): every device extracts and looks at every k-mer of every read, so N devices buy capacity, not speed.
// Here the READS are partitioned too. Rank g of N (one per GPU; one process per GPU under torchrun, or N handles
// in one process as the reference's `-d N`) holds shard g of the table and classifies ITS reads in three steps:
//
//   scatter  (k_route_scatter, on g)  stages 2+3a: every canonical k-mer of g's reads is appended to an ARENA in
//            g's own HBM, in blocks of 256 entries that each belong to ONE owner shard (the shard whose bucket
//            range holds the k-mer's home bucket); a per-owner BLOCK LIST names the blocks. pos_of[slot]
//            remembers where the k-mer of every read position went.
//   probe    (k_route_probe, on d)    stage 3b: shard d walks the block lists addressed to it in EVERY rank's
//            region — k-mers are loaded straight out of the peer's HBM over NVLink (coalesced 256-byte
//            requests), probed in d's table, and the 2-byte labels are stored straight back into the peer's
//            label array. Transfer and probe are one kernel: no staging copy, no collective call.
//   gather   (k_route_gather, on g)   stage 4: per read, the labels of its k-mers (through pos_of) feed the same
//            per-read hit counting, top-2 and sparse rows as the single-table kernel (hits.cuh).
//
// Between the steps all ranks meet at a barrier (the caller's: stream-ordered NCCL all-reduce under torchrun,
// CUDA events between the devices of one process). NVLink carries 8 bytes per k-mer one way and 2 bytes back.
// Because a canonical k-mer has exactly one home shard the labels, and so every count, equal the single-table
// result bit for bit.
// Table layouts: hashed shards (NARROW / WIDE: any table size; the arena entry is the canonical k-mer, the owner is the
// shard of its home bucket) and LOCAL shards (tables of up to 2^30 lines in all, i.e. config-2 scale: the scatter
// kernel runs the minimizer front end, the owner is the shard of the A line, the arena entry is (sector, 37-bit key);
// consecutive k-mers of a read share their minimizer, so they land next to each other in the owner's block and its
// probes coalesce into line requests exactly as in the single-table kernel — a probe costs ~0.44 line requests
// instead of one random sector).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "hits.cuh"
#include "internal.h"
#include "kmerwin.cuh"
#include "local_rows.cuh"

namespace cuclark {

namespace {

#ifndef CUCLARK_ROUTE_BLK
#define CUCLARK_ROUTE_BLK 1024
#endif
// arena entries per block = one TMA bulk copy. 8 GPUs, hashed shards, probe phase per 1.2 G k-mers: 256 entries (2 KB
// copies, 4 stages, 2 CTAs per SM) 40.2 ms; 1,024 entries (8 KB copies, 2 stages, 1 CTA per SM) 35.7 ms — the per-copy
// cost of a remote bulk copy, not NVLink bandwidth (260 GB/s per GPU), was what the smaller blocks paid.
constexpr uint32_t BLK = CUCLARK_ROUTE_BLK;
constexpr uint32_t SUB = 256;                     // entries a warp holds in registers at a time while probing a block
static_assert(BLK % SUB == 0, "block = whole sub-blocks");
constexpr uint64_t SENTINEL = ~0ull;              // unused arena entry (never a canonical k-mer, also at k = 32)
constexpr uint16_t LABEL_NONE = 0xFFFF;
constexpr int R_WARPS = 8;
constexpr int R_READS_PER_CHUNK = 31;
constexpr int SCATTER_BLOCKS_PER_SM = 4;
constexpr uint32_t ROUTE_ERR_ARENA = 1u;

// Head of a rank's region; the peers read nblk[] (how many blocks are addressed to them).
struct RouteHeader {
    uint32_t n_blocks;                            // blocks handed out of the arena by the last scatter
    uint32_t err;
    uint32_t chunk;                               // dynamic work counters of the three kernels
    uint32_t chunk_gather;
    uint32_t nblk[ROUTE_MAX_RANKS];
    uint32_t pad[ROUTE_MAX_RANKS - 4];
    unsigned long long lookups;
    unsigned long long probed;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// addresses inside a region whose base is `base` (the same layout on every rank)
struct RegionView {
    RouteHeader* hdr;
    uint32_t* blocklist;                          // [n_ranks][cap_blocks]
    uint64_t* arena;                              // [cap_blocks * BLK]: canonical k-mers (hashed shards) or slot keys (LOCAL shards)
    uint16_t* labels;                             // [cap_blocks * BLK]
    uint32_t* sec32;                              // [cap_blocks * BLK], LOCAL shards: A sector of the entry, relative to its owner's shard
};
__host__ __device__ inline RegionView region_view(uint8_t* base, int n_ranks, uint32_t cap_blocks) {
    RegionView v;
    size_t off = 0;
    v.hdr = reinterpret_cast<RouteHeader*>(base); off = align256(sizeof(RouteHeader));
    v.blocklist = reinterpret_cast<uint32_t*>(base + off); off = align256(off + (size_t)n_ranks * cap_blocks * 4);
    v.arena = reinterpret_cast<uint64_t*>(base + off); off = align256(off + (size_t)cap_blocks * BLK * 8);
    v.labels = reinterpret_cast<uint16_t*>(base + off); off = align256(off + (size_t)cap_blocks * BLK * 2);
    v.sec32 = reinterpret_cast<uint32_t*>(base + off);
    return v;
}
size_t region_bytes(int n_ranks, uint32_t cap_blocks, bool with_sec) {
    size_t off = align256(sizeof(RouteHeader));
    off = align256(off + (size_t)n_ranks * cap_blocks * 4);
    off = align256(off + (size_t)cap_blocks * BLK * 8);
    off = align256(off + (size_t)cap_blocks * BLK * 2);
    if (with_sec) off = align256(off + (size_t)cap_blocks * BLK * 4);
    return off;
}

struct ScatterParams {
    const uint32_t* reads_ptr;
    const uint16_t* cont;
    uint32_t n_reads;
    int k;
    uint64_t M, magic;                            // the GLOBAL bucket count of the sharded table
    int n_ranks;
    uint64_t lo[ROUTE_MAX_RANKS + 1];             // shard i holds home buckets [lo[i], lo[i+1]) (LOCAL: lines)
    float inv_width;                              // n_ranks / M (LOCAL: / NL), for the owner estimate
    uint32_t NL, nl_m32;                          // LOCAL: the GLOBAL line count and the constants of local_divmod()
    int nl_sh;
    RegionView mine;
    uint32_t cap_blocks;
    uint32_t* pos_of;                             // [8 * n_cont]: arena index of the k-mer at (container, nucleotide)
};

// shard that holds home bucket b: estimate by a float product, exact by the boundaries
__device__ __forceinline__ int owner_of(const ScatterParams& p, uint64_t b) {
    int d = min(p.n_ranks - 1, (int)((float)b * p.inv_width));
    while (b < p.lo[d]) d--;
    while (b >= p.lo[d + 1]) d++;
    return d;
}

// ---- scatter: k-mers of this rank's reads into per-owner blocks of the local arena -------------------------
template <bool LOCAL>
__global__ void __launch_bounds__(R_WARPS * 32, SCATTER_BLOCKS_PER_SM) k_route_scatter(const ScatterParams p) {
    __shared__ uint32_t s_base[R_WARPS][ROUTE_MAX_RANKS];     // arena index of the owner's open block
    __shared__ uint32_t s_used[R_WARPS][ROUTE_MAX_RANKS];     // entries used in it (BLK = none open)
    __shared__ uint32_t s_ptr[R_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* base_of = s_base[wib];
    uint32_t* used_of = s_used[wib];
    uint32_t* sptr = s_ptr[wib];
    if (lane < ROUTE_MAX_RANKS) { base_of[lane] = 0; used_of[lane] = BLK; }
    __syncwarp();
    const int k = p.k, kshift = 64 - 2 * k;
    const uint32_t lt = (1u << lane) - 1u;
    unsigned long long my_lookups = 0;
    RouteHeader* H = p.mine.hdr;

    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&H->chunk, 1u);
        chunk = __shfl_sync(0xFFFFFFFFu, chunk, 0);
        const uint64_t base64 = (uint64_t)chunk * R_READS_PER_CHUNK;
        if (base64 >= p.n_reads) break;
        const uint32_t base = (uint32_t)base64;
        const uint32_t nr = min((uint32_t)R_READS_PER_CHUNK, p.n_reads - base);
        __syncwarp();
        sptr[lane] = p.reads_ptr[min(base + lane, p.n_reads)];
        __syncwarp();
        for (uint32_t ri = 0; ri < nr; ri++) {
            uint32_t pos = sptr[ri];
            const uint32_t end = sptr[ri + 1];
            while (pos < end) {
                const uint32_t first = pos + 1;
                uint32_t L;
                const uint32_t ncont = part_extent(p.cont[pos], first, end, L);
                pos = first + ncont;
                const int nk = (int)L - k + 1;
                for (int cb = 0; cb < nk; cb += 32 * CHUNK_ROUNDS) {
                    const uint32_t co = (uint32_t)cb >> 3;
                    const int nwin = (int)min(4u, (ncont - co + 31) >> 5);
                    uint32_t wv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t ci = first + co + 32 * u + lane;
                        wv[u] = (u < nwin && ci < pos) ? (uint32_t)p.cont[ci] : 0u;
                    }
                    const uint64_t W = assemble_words(wv, nwin, lane);
                    const int rounds = min(CHUNK_ROUNDS, (nk - cb + 31) >> 5);
                    LocalCarry carry;                        // LOCAL: hash row handed over from the previous round
                    const int m_limit = (int)L - (k - LOCAL_W + 1) - cb - lane;
                    for (int i = 0; i < rounds; i++) {
                        const int w = cb + 32 * i + lane;
                        const bool valid = w < nk;
                        uint64_t entry;                      // what the owner needs: the k-mer, or (LOCAL) its slot key ...
                        uint32_t entry_sec = 0;              // ... and its A sector inside the owner's shard
                        int d = -1;
                        if (LOCAL) {
                            const LocalProbe P = local_row(carry, W, i, m_limit, k, kshift, p.NL, p.nl_m32, p.nl_sh, lane);
                            if (valid) {
                                d = owner_of(p, (uint64_t)P.line);
                                entry_sec = (uint32_t)(((uint64_t)P.line - p.lo[d]) * 4 + (uint64_t)(P.o_c & 3));
                            }
                            entry = P.key;
                        } else {
                            const uint64_t hi = shfl64(W, i & 31), lo = shfl64(W, (i + 1) & 31);
                            entry = canonical(window64(hi, lo, 2 * lane) >> kshift, k);
                            uint64_t q, b;
                            divmod_M(entry, p.M, p.magic, q, b);
                            if (valid) d = owner_of(p, b);
                        }
                        my_lookups += valid;
                        // lanes with the same owner take consecutive entries of that owner's open block
                        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, d);
                        const int leader = __ffs(grp) - 1;
                        uint32_t at = 0;
                        if (valid && lane == leader) {
                            const uint32_t cnt = __popc(grp);
                            uint32_t used = used_of[d], bs = base_of[d];
                            if (used + cnt > BLK) {
                                // the tail of the open block stays unused, then a fresh block of the arena
                                for (uint32_t t = used; t < BLK; t++) p.mine.arena[bs + t] = SENTINEL;
                                const uint32_t blk = atomicAdd(&H->n_blocks, 1u);
                                if (blk < p.cap_blocks) {
                                    p.mine.blocklist[(size_t)d * p.cap_blocks + atomicAdd(&H->nblk[d], 1u)] = blk;
                                    bs = blk * BLK;
                                } else {
                                    atomicOr(&H->err, ROUTE_ERR_ARENA);      // cannot happen (capacity is an upper bound)
                                    bs = 0;
                                }
                                used = 0;
                                base_of[d] = bs;
                            }
                            used_of[d] = used + cnt;
                            at = bs + used;
                        }
                        at = __shfl_sync(0xFFFFFFFFu, at, leader < 0 ? 0 : leader) + __popc(grp & lt);
                        if (valid) {
                            p.mine.arena[at] = entry;
                            if (LOCAL) p.mine.sec32[at] = entry_sec;
                            p.pos_of[8u * first + (uint32_t)w] = at;
                        }
                        __syncwarp();
                    }
                }
            }
        }
    }
    // close the open blocks
    __syncwarp();
    for (int d = 0; d < p.n_ranks; d++) {
        const uint32_t used = used_of[d], bs = base_of[d];
        for (uint32_t t = used + lane; t < BLK; t += 32) p.mine.arena[bs + t] = SENTINEL;
    }
    for (int o = 16; o; o >>= 1) my_lookups += __shfl_xor_sync(0xFFFFFFFFu, my_lookups, o);
    if (lane == 0 && my_lookups) atomicAdd(&H->lookups, my_lookups);
}

// ---- probe: this shard answers the blocks addressed to it in every rank's region ---------------------------
struct ProbeParams {
    TableView t;
    int n_ranks, rank;
    uint32_t cap_blocks;
    uint32_t n_targets;
    uint8_t* region[ROUTE_MAX_RANKS];             // every rank's region as this device addresses it
    RouteHeader* my_hdr;
};

// One block of 256 k-mers: probe them in this shard's table, store the labels into the asker's label array.
template <int LAYOUT>
__device__ __forceinline__ void probe_block(const ProbeParams& p, const uint64_t (&c)[SUB / 32], uint16_t* dst, int lane,
                                            unsigned long long& probed) {
    const TableView& T = p.t;
#pragma unroll
    for (int h = 0; h < (int)(SUB / 32); h += 4) {
        Sector sec[4];
        uint64_t q[4], b[4];
        bool live[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            divmod_M(c[h + j], T.M, T.magic, q[j], b[j]);
            b[j] -= T.lo;
            live[j] = c[h + j] != SENTINEL && b[j] < T.n_local;
            if (live[j]) sec[j] = load_sector(T.buckets + 2 * b[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t label = NO_LABEL;
            if (live[j]) {
                label = match_sector<LAYOUT>(sec[j], q[j]);
                if (label == NO_LABEL && sector_overflowed(sec[j])) label = ovf_lookup(T, c[h + j]);
                if (label >= p.n_targets) label = NO_LABEL;
                probed++;
            }
            dst[32 * (h + j) + lane] = label == NO_LABEL ? LABEL_NONE : (uint16_t)label;      // 64-byte stores, posted
        }
    }
}

// The same for a LOCAL shard: the entry is (A sector inside this shard, 37-bit key). Consecutive entries of a block
// come from consecutive k-mers of a read, share their minimizer and so their two candidate lines: the 32 lanes'
// loads of one round coalesce into ~14 line requests. Two entries per lane in flight (2 x 2 sectors of registers).
__device__ __forceinline__ void probe_block_local(const ProbeParams& p, const uint64_t (&key)[SUB / 32], const uint32_t (&sa)[SUB / 32],
                                                  uint16_t* dst, int lane, unsigned long long& probed) {
    const TableView& T = p.t;
#pragma unroll
    for (int h = 0; h < (int)(SUB / 32); h += 2) {
        Sector A[2], B[2];
        bool live[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            live[j] = key[h + j] != SENTINEL && (uint64_t)sa[h + j] < T.n_local;
            const uint32_t a = live[j] ? sa[h + j] : 0u;                 // lanes without an entry read sector 0
            const uint32_t zq = (uint32_t)key[h + j] & ((1u << LOCAL_ZQ_BITS) - 1u);
            const uint32_t b = live[j] ? local_alt_rel(a >> 2, zq, T.line_n, false) * 4u + (a & 3u) : 0u;
            A[j] = load_sector_line(T.buckets + 2 * (uint64_t)a);
            B[j] = load_sector_line(T.buckets + 2 * (uint64_t)b);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t la = match_sector<LAYOUT_LOCAL>(A[j], key[h + j]);
            const uint32_t lb = match_sector<LAYOUT_LOCAL>(B[j], key[h + j] | ((uint64_t)LOCAL_ALT_BIT << 32));
            uint32_t label = la == NO_LABEL ? lb : la;
            const bool ask_ovf = live[j] && label == NO_LABEL && sector_overflowed(A[j]) && sector_overflowed(B[j]);
            if (__any_sync(0xFFFFFFFFu, ask_ovf)) {
                if (ask_ovf) label = ovf_lookup(T, local_rebuild((uint64_t)T.line_lo + (sa[h + j] >> 2), key[h + j], T.k, T.NL));
            }
            if (!live[j] || label >= p.n_targets) label = NO_LABEL;
            probed += live[j];
            dst[32 * (h + j) + lane] = label == NO_LABEL ? LABEL_NONE : (uint16_t)label;
        }
    }
}

#if CUCLARK_ROUTE_BLK == 256
// Every warp walks its share of the block list addressed to this shard in rank g's region. The k-mers of the NEXT
// block (and the list entry after it) are already on their way over NVLink while the current block is probed:
// a remote load takes microseconds, and without the prefetch the probe rate fell from 33 G/s (2 GPUs, half of the
// blocks remote) to 22 G/s (8 GPUs, 7/8 remote).
template <int LAYOUT>
__global__ void __launch_bounds__(256, 2) k_route_probe(const ProbeParams p) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long probed = 0;
    for (int gi = 0; gi < p.n_ranks; gi++) {
        const int g = (p.rank + gi) % p.n_ranks;              // start with the own region, then round the ring
        const RegionView R = region_view(p.region[g], p.n_ranks, p.cap_blocks);
        const uint32_t nb = min(__ldcv(&R.hdr->nblk[p.rank]), p.cap_blocks);
        const uint32_t* list = R.blocklist + (size_t)p.rank * p.cap_blocks;
        uint32_t i = warp;
        if (i >= nb) continue;
        uint32_t blk = __ldcv(&list[i]);
        uint32_t blk_next = i + n_warps < nb ? __ldcv(&list[i + n_warps]) : 0;
        uint64_t c[BLK / 32];
#pragma unroll
        for (int j = 0; j < (int)(BLK / 32); j++) c[j] = __ldcv(&R.arena[(size_t)blk * BLK + 32 * j + lane]);     // 256-byte requests
        for (;;) {
            const uint32_t i_next = i + n_warps;
            const bool more = i_next < nb;
            uint64_t cn[BLK / 32];
            uint32_t blk_next2 = 0;
            if (more) {
#pragma unroll
                for (int j = 0; j < (int)(BLK / 32); j++) cn[j] = __ldcv(&R.arena[(size_t)blk_next * BLK + 32 * j + lane]);
                if (i_next + n_warps < nb) blk_next2 = __ldcv(&list[i_next + n_warps]);
            }
            probe_block<LAYOUT>(p, c, R.labels + (size_t)blk * BLK, lane, probed);
            if (!more) break;
#pragma unroll
            for (int j = 0; j < (int)(BLK / 32); j++) c[j] = cn[j];
            blk = blk_next; blk_next = blk_next2; i = i_next;
        }
    }
    for (int o = 16; o; o >>= 1) probed += __shfl_xor_sync(0xFFFFFFFFu, probed, o);
    if (lane == 0 && probed) atomicAdd(&p.my_hdr->probed, probed);
}

#endif

// ---- the same walk with the k-mer blocks brought in by the TMA (bulk async copies into shared memory) ------
// Each warp owns a ring of PROBE_STAGES 2 KB stages and one mbarrier per stage; lane 0 arms the barrier with the
// byte count and issues `cp.async.bulk` from the (peer) arena, the warp waits on the barrier's phase before it
// reads the stage. Up to PROBE_STAGES blocks per warp are in flight over NVLink while one is being probed, without
// holding them in registers. A warp takes a CONTIGUOUS range of the block list so that 32 list entries arrive
// with one coalesced load.
#ifndef CUCLARK_PROBE_STAGES
#define CUCLARK_PROBE_STAGES (CUCLARK_ROUTE_BLK >= 1024 ? 2 : 4)
#endif
#ifndef CUCLARK_PROBE_CTAS
#define CUCLARK_PROBE_CTAS (CUCLARK_ROUTE_BLK >= 1024 ? 1 : 2)
#endif
#ifndef CUCLARK_PROBE_WARPS
#define CUCLARK_PROBE_WARPS 8
#endif
constexpr int PROBE_STAGES = CUCLARK_PROBE_STAGES;
constexpr int PROBE_CTAS_PER_SM = CUCLARK_PROBE_CTAS;
constexpr int PROBE_WARPS = CUCLARK_PROBE_WARPS;
template <bool LOCAL>
struct alignas(128) ProbeWarpSmem {
    uint64_t kmers[PROBE_STAGES][BLK];
    uint32_t secs[LOCAL ? PROBE_STAGES : 1][LOCAL ? BLK : 4];
    uint64_t bar[PROBE_STAGES];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

template <int LAYOUT>
__global__ void __launch_bounds__(PROBE_WARPS * 32, PROBE_CTAS_PER_SM) k_route_probe_tma(const ProbeParams p) {
    extern __shared__ __align__(128) uint8_t probe_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr bool LOCAL = LAYOUT == LAYOUT_LOCAL;
    ProbeWarpSmem<LOCAL>& S = reinterpret_cast<ProbeWarpSmem<LOCAL>*>(probe_smem)[wib];
    if (lane == 0)
        for (int s = 0; s < PROBE_STAGES; s++) mbar_init(&S.bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t warp = blockIdx.x * PROBE_WARPS + wib, n_warps = gridDim.x * PROBE_WARPS;
    unsigned long long probed = 0;
    uint32_t n_issued = 0, n_used = 0;           // blocks this warp has requested / consumed so far (stage = n % STAGES)
    for (int gi = 0; gi < p.n_ranks; gi++) {
        const int g = (p.rank + gi) % p.n_ranks;
        const RegionView R = region_view(p.region[g], p.n_ranks, p.cap_blocks);
        const uint32_t nb = min(__ldcv(&R.hdr->nblk[p.rank]), p.cap_blocks);
        const uint32_t* list = R.blocklist + (size_t)p.rank * p.cap_blocks;
        const uint32_t lo = (uint32_t)((uint64_t)nb * warp / n_warps), hi = (uint32_t)((uint64_t)nb * (warp + 1) / n_warps);
        uint32_t entry_next = lo + lane < hi ? __ldcv(&list[lo + lane]) : 0;
        for (uint32_t base = lo; base < hi; base += 32) {
            const uint32_t entry = entry_next;                                   // block ids of this group of <= 32 blocks
            const uint32_t m = min(32u, hi - base);
            entry_next = base + 32 + lane < hi ? __ldcv(&list[base + 32 + lane]) : 0;
            auto issue = [&](uint32_t t) {                                       // warp-uniform t
                const uint32_t blk = __shfl_sync(0xFFFFFFFFu, entry, t);
                const uint32_t st = n_issued % PROBE_STAGES;
                if (lane == 0) {
                    mbar_expect_tx(&S.bar[st], LOCAL ? BLK * 12 : BLK * 8);
                    bulk_g2s(S.kmers[st], R.arena + (size_t)blk * BLK, BLK * 8, &S.bar[st]);
                    if (LOCAL) bulk_g2s(S.secs[st], R.sec32 + (size_t)blk * BLK, BLK * 4, &S.bar[st]);
                }
                n_issued++;
            };
            for (uint32_t t = 0; t < min((uint32_t)PROBE_STAGES, m); t++) issue(t);
            for (uint32_t t = 0; t < m; t++) {
                const uint32_t st = n_used % PROBE_STAGES, parity = (n_used / PROBE_STAGES) & 1u;
                mbar_wait(&S.bar[st], parity);
                const uint32_t blk = __shfl_sync(0xFFFFFFFFu, entry, t);
#pragma unroll 1
                for (uint32_t sub = 0; sub < BLK; sub += SUB) {
                    uint64_t c[SUB / 32];
                    uint32_t sa[LOCAL ? SUB / 32 : 1];
#pragma unroll
                    for (int j = 0; j < (int)(SUB / 32); j++) {
                        c[j] = S.kmers[st][sub + 32 * j + lane];
                        if (LOCAL) sa[LOCAL ? j : 0] = S.secs[st][sub + 32 * j + lane];
                    }
                    if (sub + SUB == BLK) {                                       // everything read: the stage is free again
                        n_used++;
                        __syncwarp();
                        if (t + PROBE_STAGES < m) issue(t + PROBE_STAGES);
                    }
                    if constexpr (LOCAL) probe_block_local(p, c, sa, R.labels + (size_t)blk * BLK + sub, lane, probed);
                    else probe_block<LAYOUT>(p, c, R.labels + (size_t)blk * BLK + sub, lane, probed);
                }
            }
        }
    }
    for (int o = 16; o; o >>= 1) probed += __shfl_xor_sync(0xFFFFFFFFu, probed, o);
    if (lane == 0 && probed) atomicAdd(&p.my_hdr->probed, probed);
}

// ---- gather: labels of this rank's reads -> per-read result -----------------------------------------------
struct GatherParams {
    const uint32_t* reads_ptr;
    const uint16_t* cont;
    uint32_t n_reads;
    int k;
    uint32_t n_targets;
    const uint32_t* pos_of;
    const uint16_t* labels;
    RouteHeader* hdr;
    HitSink out;
};

template <bool ROWS>
__global__ void __launch_bounds__(R_WARPS * 32, 4) k_route_gather(const GatherParams p) {
    __shared__ uint32_t s_key[R_WARPS][TSLOTS];
    __shared__ uint32_t s_cnt[R_WARPS][TSLOTS];
    __shared__ uint16_t s_row[ROWS ? R_WARPS : 1][ROWS ? 2 * MAX_ROW_PAIRS + 2 : 2];
    __shared__ uint32_t s_ptr[R_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* tkey = s_key[wib];
    uint32_t* tcnt = s_cnt[wib];
    uint32_t* sptr = s_ptr[wib];
    tab_clear(tkey, tcnt, lane);
    const int k = p.k;
    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&p.hdr->chunk_gather, 1u);
        chunk = __shfl_sync(0xFFFFFFFFu, chunk, 0);
        const uint64_t base64 = (uint64_t)chunk * R_READS_PER_CHUNK;
        if (base64 >= p.n_reads) break;
        const uint32_t base = (uint32_t)base64;
        const uint32_t nr = min((uint32_t)R_READS_PER_CHUNK, p.n_reads - base);
        __syncwarp();
        sptr[lane] = p.reads_ptr[min(base + lane, p.n_reads)];
        __syncwarp();
        for (uint32_t ri = 0; ri < nr; ri++) {
            uint32_t pos = sptr[ri];
            const uint32_t end = sptr[ri + 1];
            WarpHits hits;
            while (pos < end) {
                const uint32_t first = pos + 1;
                uint32_t L;
                const uint32_t ncont = part_extent(p.cont[pos], first, end, L);
                pos = first + ncont;
                const int nk = (int)L - k + 1;
                for (int w0 = 0; w0 < nk; w0 += 128) {
                    uint32_t lab[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {                  // four independent pos_of -> label chains
                        const int w = w0 + 32 * j + lane;
                        lab[j] = NO_LABEL;
                        if (w < nk) {
                            const uint32_t v = __ldcg(&p.labels[p.pos_of[8u * first + (uint32_t)w]]);      // written by the peers: L2, not L1
                            if (v < p.n_targets) lab[j] = v;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (w0 + 32 * j < nk) hits.add(lab[j], tkey, tcnt, lane);      // warp-uniform condition
                }
            }
            __syncwarp();
            hits.finish<ROWS>(p.out, base + ri, tkey, tcnt, s_row[wib], lane);
        }
    }
}

// exact fallback for reads with more than 64 distinct targets: dense per-target counters (as k_classify_dense)
__global__ void __launch_bounds__(256) k_route_dense(const GatherParams p, uint32_t* hist_all) {
    uint32_t* hist = hist_all + (size_t)blockIdx.x * p.n_targets;
    const uint32_t n_list = min(p.out.counters[COUNTER_DENSE], p.out.dense_cap);
    for (uint32_t li = blockIdx.x; li < n_list; li += gridDim.x) {
        const uint32_t read = p.out.dense_list[li];
        uint32_t pos = p.reads_ptr[read];
        const uint32_t end = p.reads_ptr[read + 1];
        while (pos < end) {
            const uint32_t first = pos + 1;
            uint32_t L;
            const uint32_t ncont = part_extent(p.cont[pos], first, end, L);
            pos = first + ncont;
            const int nk = (int)L - p.k + 1;
            for (int w = threadIdx.x; w < nk; w += blockDim.x) {
                const uint32_t v = __ldcg(&p.labels[p.pos_of[8u * first + (uint32_t)w]]);
                if (v < p.n_targets) atomicAdd(&hist[v], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) dense_emit(hist, p.n_targets, read, p.out);
        __syncthreads();
    }
}

}  // namespace

// ---- host side ----------------------------------------------------------------------------------------
struct RouteCtx {
    int n_ranks = 0, rank = 0;
    size_t cap_cont = 0;
    uint32_t cap_blocks = 0;
    uint8_t* region = nullptr;
    size_t bytes = 0;
    uint8_t* peer[ROUTE_MAX_RANKS] = {};
    bool peer_ipc[ROUTE_MAX_RANKS] = {};
    uint32_t* pos_of = nullptr;
    bool local = false;                       // LOCAL shards: entries are (sector, key), the scatter runs the minimizer front end
    RouteHeader* h_hdr = nullptr;             // pinned copy for stats
    int sm_count = 0;
    uint64_t last_lookups = 0, last_probed = 0, last_blocks = 0;
};

void route_free(cuclark_db* db) {
    RouteCtx* r = db->route;
    if (!r) return;
    cudaSetDevice(db->cfg.device);
    cudaDeviceSynchronize();
    for (int i = 0; i < r->n_ranks; i++)
        if (r->peer_ipc[i] && r->peer[i]) cudaIpcCloseMemHandle(r->peer[i]);
    cudaFree(r->region);
    cudaFree(r->pos_of);
    cudaFreeHost(r->h_hdr);
    delete r;
    db->route = nullptr;
}

int route_alloc(cuclark_db* db, int n_ranks, size_t max_containers) {
    if (!db->d_table) { set_error("no database loaded"); return CUCLARK_ERR_STATE; }
    if (n_ranks < 1 || n_ranks > ROUTE_MAX_RANKS) { set_error("n_ranks must be 1..%d", ROUTE_MAX_RANKS); return CUCLARK_ERR_ARG; }
    if (db->cfg.shard_count != n_ranks) { set_error("the handle holds shard %d of %d, not of %d", db->cfg.shard_index, db->cfg.shard_count, n_ranks); return CUCLARK_ERR_ARG; }
    if (max_containers == 0 || max_containers > ((size_t)1 << 28)) { set_error("max_containers must be in [1, 2^28] per call"); return CUCLARK_ERR_ARG; }
    route_free(db);
    RouteCtx* r = new RouteCtx();
    r->n_ranks = n_ranks; r->rank = db->cfg.shard_index; r->cap_cont = max_containers; r->sm_count = db->sm_count;
    // every k-mer starts at a nucleotide of a data container: at most 8 per container. On top, each warp of the
    // scatter kernel leaves at most one open block per owner and wastes < 32 entries per block it closes.
    const uint64_t warps = (uint64_t)db->sm_count * SCATTER_BLOCKS_PER_SM * R_WARPS;
    const uint64_t entries = 8ull * max_containers;
    const uint64_t blocks = (entries + (BLK - 32) - 1) / (BLK - 32) + warps * n_ranks + 64;
    if (blocks * BLK >= 0xFFFFFFFFull) { delete r; set_error("routing arena exceeds 2^32 entries: lower max_containers"); return CUCLARK_ERR_ARG; }
    r->cap_blocks = (uint32_t)blocks;
    r->local = db->view.layout == LAYOUT_LOCAL;
    r->bytes = region_bytes(n_ranks, r->cap_blocks, r->local);
    if (cudaMalloc(&r->region, r->bytes) != cudaSuccess) { cudaGetLastError(); delete r; set_error("cudaMalloc of the %.2f GB routing region failed", r->bytes / 1e9); return CUCLARK_ERR_NOMEM; }
    if (cudaMalloc(&r->pos_of, 8 * max_containers * 4) != cudaSuccess) { cudaGetLastError(); cudaFree(r->region); delete r; set_error("cudaMalloc of the position map failed"); return CUCLARK_ERR_NOMEM; }
    if (cudaMallocHost(&r->h_hdr, sizeof(RouteHeader)) != cudaSuccess) { cudaGetLastError(); cudaFree(r->region); cudaFree(r->pos_of); delete r; set_error("cudaMallocHost failed"); return CUCLARK_ERR_NOMEM; }
    CK(cudaMemset(r->region, 0, align256(sizeof(RouteHeader))));
    r->peer[r->rank] = r->region;
    db->route = r;
    return CUCLARK_OK;
}

int route_export(cuclark_db* db, void* handle64, uint64_t* bytes) {
    RouteCtx* r = db->route;
    if (!r) { set_error("cuclark_route_alloc first"); return CUCLARK_ERR_STATE; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, r->region));
    memcpy(handle64, &h, 64);
    if (bytes) *bytes = r->bytes;
    return CUCLARK_OK;
}

int route_import(cuclark_db* db, int peer_rank, const void* handle64) {
    RouteCtx* r = db->route;
    if (!r) { set_error("cuclark_route_alloc first"); return CUCLARK_ERR_STATE; }
    if (peer_rank < 0 || peer_rank >= r->n_ranks || peer_rank == r->rank) { set_error("bad peer rank"); return CUCLARK_ERR_ARG; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    r->peer[peer_rank] = static_cast<uint8_t*>(ptr);
    r->peer_ipc[peer_rank] = true;
    return CUCLARK_OK;
}

// handles of ONE process, in rank order: plain peer pointers (peer access enabled between the devices)
int route_connect(cuclark_db* const* dbs, int n) {
    for (int i = 0; i < n; i++) {
        if (!dbs[i] || !dbs[i]->route) { set_error("cuclark_route_alloc every handle first"); return CUCLARK_ERR_STATE; }
        RouteCtx* r = dbs[i]->route;
        if (r->n_ranks != n || r->rank != i) { set_error("handles must be passed in rank order (handle %d is rank %d of %d)", i, r->rank, r->n_ranks); return CUCLARK_ERR_ARG; }
        if (r->cap_blocks != dbs[0]->route->cap_blocks || r->local != dbs[0]->route->local) { set_error("all ranks must allocate the same capacity and hold the same table layout"); return CUCLARK_ERR_ARG; }
    }
    for (int i = 0; i < n; i++) {
        CK(cudaSetDevice(dbs[i]->cfg.device));
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            if (dbs[j]->cfg.device != dbs[i]->cfg.device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, dbs[i]->cfg.device, dbs[j]->cfg.device));
                if (!can) { set_error("device %d cannot access device %d", dbs[i]->cfg.device, dbs[j]->cfg.device); return CUCLARK_ERR_CUDA; }
                const cudaError_t e = cudaDeviceEnablePeerAccess(dbs[j]->cfg.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); return CUCLARK_ERR_CUDA; }
                cudaGetLastError();
            }
            dbs[i]->route->peer[j] = dbs[j]->route->region;
            dbs[i]->route->peer_ipc[j] = false;
        }
    }
    return CUCLARK_OK;
}

static int route_check(cuclark_db* db, size_t n_reads, size_t n_cont, bool need_peers) {
    RouteCtx* r = db->route;
    if (!r) { set_error("cuclark_route_alloc first"); return CUCLARK_ERR_STATE; }
    if (n_cont > r->cap_cont) { set_error("%zu containers exceed the routing capacity %zu", n_cont, r->cap_cont); return CUCLARK_ERR_ARG; }
    if (n_reads > 0xFFFFFFF0ull) { set_error("too many reads in one call"); return CUCLARK_ERR_ARG; }
    if (need_peers)
        for (int i = 0; i < r->n_ranks; i++)
            if (!r->peer[i]) { set_error("rank %d's region is not connected (cuclark_route_import / cuclark_route_connect)", i); return CUCLARK_ERR_STATE; }
    return CUCLARK_OK;
}

int route_scatter(cuclark_db* db, const uint32_t* d_ptr, const uint16_t* d_cont, size_t n_reads, size_t n_cont, cudaStream_t st) {
    int rc = route_check(db, n_reads, n_cont, false);
    if (rc) return rc;
    RouteCtx* r = db->route;
    CK(cudaMemsetAsync(r->region, 0, align256(sizeof(RouteHeader)), st));
    if (n_reads == 0) return CUCLARK_OK;
    ScatterParams p;
    p.reads_ptr = d_ptr; p.cont = d_cont; p.n_reads = (uint32_t)n_reads; p.k = db->cfg.k;
    p.M = db->view.M; p.magic = db->view.magic; p.n_ranks = r->n_ranks;
    p.NL = (uint32_t)db->view.NL; p.nl_m32 = db->view.nl_m32; p.nl_sh = db->view.nl_sh;
    const uint64_t units = r->local ? db->view.NL : p.M;          // shards are ranges of lines (LOCAL) or of buckets: choose_geometry
    for (int i = 0; i <= r->n_ranks; i++) p.lo[i] = (uint64_t)((__uint128_t)units * i / r->n_ranks);
    p.inv_width = (float)((double)r->n_ranks / (double)units);
    p.mine = region_view(r->region, r->n_ranks, r->cap_blocks);
    p.cap_blocks = r->cap_blocks;
    p.pos_of = r->pos_of;
    const int blocks = (int)std::min<size_t>((n_reads + R_WARPS - 1) / R_WARPS, (size_t)r->sm_count * SCATTER_BLOCKS_PER_SM);
    if (r->local) k_route_scatter<true><<<blocks, R_WARPS * 32, 0, st>>>(p);
    else k_route_scatter<false><<<blocks, R_WARPS * 32, 0, st>>>(p);
    CK(cudaGetLastError());
    count_launches(1);
    return CUCLARK_OK;
}

int route_probe(cuclark_db* db, cudaStream_t st) {
    int rc = route_check(db, 0, 0, true);
    if (rc) return rc;
    RouteCtx* r = db->route;
    ProbeParams p;
    p.t = db->view; p.n_ranks = r->n_ranks; p.rank = r->rank; p.cap_blocks = r->cap_blocks;
    p.n_targets = (uint32_t)db->cfg.n_targets;
    for (int i = 0; i < ROUTE_MAX_RANKS; i++) p.region[i] = i < r->n_ranks ? r->peer[i] : nullptr;
    p.my_hdr = reinterpret_cast<RouteHeader*>(r->region);
    const int blocks = r->sm_count * PROBE_CTAS_PER_SM;
    static const bool use_ldg = getenv("CUCLARK_ROUTE_LDG") != nullptr;       // the register-prefetch variant, for A/B runs (hashed shards)
    if (r->local) {
        const size_t smem = sizeof(ProbeWarpSmem<true>) * PROBE_WARPS;
        CK(cudaFuncSetAttribute(k_route_probe_tma<LAYOUT_LOCAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_route_probe_tma<LAYOUT_LOCAL><<<blocks, PROBE_WARPS * 32, smem, st>>>(p);
#if CUCLARK_ROUTE_BLK == 256
    } else if (use_ldg) {
        if (db->view.layout == LAYOUT_NARROW) k_route_probe<LAYOUT_NARROW><<<blocks, 256, 0, st>>>(p);
        else k_route_probe<LAYOUT_WIDE><<<blocks, 256, 0, st>>>(p);
#endif
    } else {
        const size_t smem = sizeof(ProbeWarpSmem<false>) * PROBE_WARPS;
        if (db->view.layout == LAYOUT_NARROW) {
            CK(cudaFuncSetAttribute(k_route_probe_tma<LAYOUT_NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_route_probe_tma<LAYOUT_NARROW><<<blocks, PROBE_WARPS * 32, smem, st>>>(p);
        } else {
            CK(cudaFuncSetAttribute(k_route_probe_tma<LAYOUT_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_route_probe_tma<LAYOUT_WIDE><<<blocks, PROBE_WARPS * 32, smem, st>>>(p);
        }
    }
    CK(cudaGetLastError());
    count_launches(1);
    return CUCLARK_OK;
}

int route_gather(cuclark_db* db, const Scratch& sc, const uint32_t* d_ptr, const uint16_t* d_cont, size_t n_reads, size_t n_cont,
                 uint16_t* d_final, uint16_t* d_rows, cudaStream_t st) {
    int rc = route_check(db, n_reads, n_cont, false);
    if (rc) return rc;
    if (db->row_pairs > MAX_ROW_PAIRS) { set_error("row_pairs > %d", MAX_ROW_PAIRS); return CUCLARK_ERR_ARG; }
    RouteCtx* r = db->route;
    CK(cudaMemsetAsync(sc.d_counters, 0, N_COUNTERS * sizeof(uint32_t), st));
    const RegionView v = region_view(r->region, r->n_ranks, r->cap_blocks);
    // stats of this call: header -> pinned copy (read by cuclark_sync_stats / route_stats after the stream is idle)
    CK(cudaMemcpyAsync(r->h_hdr, v.hdr, sizeof(RouteHeader), cudaMemcpyDeviceToHost, st));
    if (n_reads == 0) return CUCLARK_OK;
    GatherParams p;
    p.reads_ptr = d_ptr; p.cont = d_cont; p.n_reads = (uint32_t)n_reads; p.k = db->cfg.k;
    p.n_targets = (uint32_t)db->cfg.n_targets; p.pos_of = r->pos_of; p.labels = v.labels; p.hdr = v.hdr;
    p.out.final5 = d_final; p.out.rows = d_rows; p.out.row_pairs = db->row_pairs;
    p.out.counters = sc.d_counters; p.out.dense_list = sc.d_dense_list; p.out.dense_cap = sc.dense_cap;
    const int blocks = (int)std::min<size_t>((n_reads + R_WARPS - 1) / R_WARPS, (size_t)r->sm_count * 8);
    if (d_rows) k_route_gather<true><<<blocks, R_WARPS * 32, 0, st>>>(p);
    else k_route_gather<false><<<blocks, R_WARPS * 32, 0, st>>>(p);
    CK(cudaGetLastError());
    if (sc.d_dense_hist) {
        k_route_dense<<<db->dense_blocks, 256, 0, st>>>(p, sc.d_dense_hist);
        CK(cudaGetLastError());
    } else {
        std::lock_guard<std::mutex> dense_guard(db->dense_mu);
        CK(cudaStreamWaitEvent(st, db->dense_chain, 0));
        k_route_dense<<<db->dense_blocks, 256, 0, st>>>(p, db->d_dense_hist);
        CK(cudaGetLastError());
        CK(cudaEventRecord(db->dense_chain, st));
    }
    count_launches(2);
    return CUCLARK_OK;
}

int route_stats(cuclark_db* db, cuclark_route_stats* out) {
    RouteCtx* r = db->route;
    if (!r) { set_error("cuclark_route_alloc first"); return CUCLARK_ERR_STATE; }
    memset(out, 0, sizeof *out);
    out->n_ranks = r->n_ranks; out->rank = r->rank;
    out->region_bytes = r->bytes; out->map_bytes = 8 * r->cap_cont * 4;
    out->cap_blocks = r->cap_blocks;
    out->lookups = r->h_hdr->lookups; out->probed = r->h_hdr->probed;
    out->blocks = r->h_hdr->n_blocks; out->err = r->h_hdr->err;
    uint64_t remote = 0;
    for (int i = 0; i < r->n_ranks; i++) if (i != r->rank) remote += r->h_hdr->nblk[i];
    out->blocks_remote = remote;
    return CUCLARK_OK;
}

}  // namespace cuclark
