// cuclark_b200 — per-read target counts of one warp and the read's result (stage 4).
//
// Replaces the reference's dense shared-memory histogram + compaction + resultKernel
// (src/CuClarkDB.cu:1064-1074, 1156-1243, 1421-1471) for the warp-per-read kernels: k_classify (classify.cu) and
// k_route_gather (route.cu) feed the labels of a row of 32 k-mers with add(), finish() emits the read's
// uint16[5] result (and its sparse row) or hands the read to the dense fallback.
//   * while every hit of the read went to ONE target (the usual case) each lane counts its own hits;
//   * otherwise a 64-slot per-warp hash table in shared memory fed by match_any group leaders;
//   * top-1/top-2 by one warp max-reduction each over (hits << 16 | ~target): identical to the reference's
//     ascending scan with strict '>' (lowest target index wins ties);
//   * more than 64 distinct targets: `overflow`, the read goes to the exact dense fallback.
#pragma once
#include "internal.h"

namespace cuclark {

constexpr int TSLOTS = 64;            // per-warp hash slots
constexpr int MAX_ROW_PAIRS = 63;
constexpr uint32_t HIT_EMPTY = 0xFFFFFFFFu;

// where results go (part of the kernel parameters)
struct HitSink {
    uint16_t* final5;
    uint16_t* rows;
    int row_pairs;
    uint32_t* counters;
    uint32_t* dense_list;
    uint32_t dense_cap;
};

// leader-only insert of (label, n) into the warp's table; false if it is full
__device__ __forceinline__ bool tab_add(uint32_t* tkey, uint32_t* tcnt, uint32_t label, uint32_t n) {
    uint32_t slot = (label * 0x9E3779B1u) >> 26;          // 6 bits
#pragma unroll 1
    for (int i = 0; i < TSLOTS; i++) {
        const uint32_t old = atomicCAS(&tkey[slot], HIT_EMPTY, label);
        if (old == HIT_EMPTY || old == label) { atomicAdd(&tcnt[slot], n); return true; }
        slot = (slot + 1) & (TSLOTS - 1);
    }
    return false;
}

__device__ __forceinline__ void tab_clear(uint32_t* tkey, uint32_t* tcnt, int lane) {
    tkey[lane] = HIT_EMPTY; tkey[lane + 32] = HIT_EMPTY; tcnt[lane] = 0; tcnt[lane + 32] = 0;
    __syncwarp();
}

struct WarpHits {
    uint32_t first_label = NO_LABEL, total = 0, my_same = 0;
    bool table_mode = false, overflow = false;

    // hits of one row of 32 k-mers into the read's counters (warp-collective); label = NO_LABEL for a miss
    __device__ __forceinline__ void add(const uint32_t label, uint32_t* tkey, uint32_t* tcnt, int lane) {
        uint32_t hitmask = 0;
        if (!table_mode) {
            if (first_label == NO_LABEL) {
                hitmask = __ballot_sync(0xFFFFFFFFu, label != NO_LABEL);
                if (!hitmask) return;
                first_label = __shfl_sync(0xFFFFFFFFu, label, __ffs(hitmask) - 1);
            }
            const bool hit = label != NO_LABEL;
            if (!__any_sync(0xFFFFFFFFu, hit && label != first_label)) { my_same += hit; return; }
            // a second target: fold the per-lane counts and continue in the shared-memory table
            const uint32_t first_cnt = __reduce_add_sync(0xFFFFFFFFu, my_same);
            total = first_cnt;
            table_mode = true;
            if (lane == 0 && first_cnt) tab_add(tkey, tcnt, first_label, first_cnt);
            __syncwarp();
        }
        hitmask = __ballot_sync(0xFFFFFFFFu, label != NO_LABEL);
        if (!hitmask) return;
        total += __popc(hitmask);
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, label);
        bool ok = true;
        if (label != NO_LABEL && lane == __ffs(grp) - 1) ok = tab_add(tkey, tcnt, label, __popc(grp));
        if (__any_sync(0xFFFFFFFFu, !ok)) overflow = true;
        __syncwarp();
    }

    // the read's result; srow = 2*MAX_ROW_PAIRS+2 uint16 of shared memory of this warp (ROWS only).
    // Leaves the warp's table empty for the next read.
    template <bool ROWS>
    __device__ __forceinline__ void finish(const HitSink& p, uint32_t read, uint32_t* tkey, uint32_t* tcnt, uint16_t* srow,
                                           int lane) {
        const int pitch = 2 * p.row_pairs + 2;
        uint16_t* row = ROWS ? p.rows + (size_t)read * pitch : nullptr;
        if (overflow) {
            // the exact dense fallback will write the result; clear the table
            tab_clear(tkey, tcnt, lane);
            if (lane == 0) {
                const uint32_t i = atomicAdd(&p.counters[COUNTER_DENSE], 1u);
                if (i < p.dense_cap) p.dense_list[i] = read;
            }
            return;
        }
        uint32_t v_sum, v_i1, v_h1, v_i2, v_h2;
        if (!table_mode) {
            const uint32_t h = __reduce_add_sync(0xFFFFFFFFu, my_same) & 0xFFFFu;
            v_sum = h; v_i1 = h ? first_label + 1 : 0; v_h1 = h; v_i2 = 0; v_h2 = 0;
            if (ROWS) {
                for (int i = lane; i < pitch; i += 32) {
                    uint16_t v = 0;
                    if (h) v = i == 0 ? 1 : i == 1 ? (uint16_t)first_label : i == 2 ? (uint16_t)h : 0;
                    row[i] = v;
                }
            }
        } else {
            // each lane owns slots lane and lane+32
            const uint32_t l0 = tkey[lane], l1 = tkey[lane + 32];
            const uint32_t h0 = l0 == HIT_EMPTY ? 0 : (tcnt[lane] & 0xFFFFu);
            const uint32_t h1 = l1 == HIT_EMPTY ? 0 : (tcnt[lane + 32] & 0xFFFFu);
            const uint32_t k0 = h0 ? (h0 << 16) | (0xFFFFu - l0) : 0;
            const uint32_t k1 = h1 ? (h1 << 16) | (0xFFFFu - l1) : 0;
            const uint32_t best = __reduce_max_sync(0xFFFFFFFFu, max(k0, k1));
            const uint32_t e0 = k0 == best ? 0 : k0, e1 = k1 == best ? 0 : k1;
            const uint32_t second = __reduce_max_sync(0xFFFFFFFFu, max(e0, e1));
            v_sum = total & 0xFFFFu;
            v_h1 = best >> 16; v_i1 = best ? (0xFFFFu - (best & 0xFFFFu)) + 1 : 0;
            v_h2 = second >> 16; v_i2 = second ? (0xFFFFu - (second & 0xFFFFu)) + 1 : 0;
            if (ROWS) {
                for (int i = lane; i < pitch; i += 32) srow[i] = 0;
                __syncwarp();
                // rank of each occupied slot = number of occupied slots with a smaller target
                uint32_t r0 = 0, r1 = 0;
                for (int s = 0; s < TSLOTS; s++) {
                    const uint32_t ls = tkey[s];
                    const bool occ = ls != HIT_EMPTY && (tcnt[s] & 0xFFFFu);
                    r0 += occ && ls < l0;
                    r1 += occ && ls < l1;
                }
                if (h0 && r0 < (uint32_t)p.row_pairs) { srow[1 + 2 * r0] = (uint16_t)l0; srow[2 + 2 * r0] = (uint16_t)h0; }
                if (h1 && r1 < (uint32_t)p.row_pairs) { srow[1 + 2 * r1] = (uint16_t)l1; srow[2 + 2 * r1] = (uint16_t)h1; }
                const uint32_t n = __popc(__ballot_sync(0xFFFFFFFFu, h0 != 0)) + __popc(__ballot_sync(0xFFFFFFFFu, h1 != 0));
                if (lane == 0) {
                    srow[0] = (uint16_t)n;
                    if (n > (uint32_t)p.row_pairs) atomicAdd(&p.counters[COUNTER_TRUNC], 1u);
                }
                __syncwarp();
                for (int i = lane; i < pitch; i += 32) row[i] = srow[i];
            }
            tab_clear(tkey, tcnt, lane);
        }
        if (p.final5 && lane < 5) {
            const uint32_t v = lane == 0 ? v_sum : lane == 1 ? v_i1 : lane == 2 ? v_h1 : lane == 3 ? v_i2 : v_h2;
            p.final5[(size_t)read * 5 + lane] = (uint16_t)v;
        }
    }
};

// The dense fallback's epilogue: ascending scan of one read's dense per-target counters by ONE thread
// (resultKernel's scan, src/CuClarkDB.cu:1430-1465), sparse row, counters zeroed again.
__device__ __forceinline__ void dense_emit(uint32_t* hist, uint32_t n_targets, uint32_t read, const HitSink& p) {
    const int pitch = 2 * p.row_pairs + 2;
    uint16_t best = 0, sbest = 0, ib = 0, isb = 0, sum = 0;
    uint32_t n = 0;
    uint16_t* row = p.rows ? p.rows + (size_t)read * pitch : nullptr;
    if (row) for (int i = 0; i < pitch; i++) row[i] = 0;
    for (uint32_t t = 0; t < n_targets; t++) {
        const uint16_t h = (uint16_t)hist[t];
        hist[t] = 0;
        if (!h) continue;
        if (h > best) { sbest = best; isb = ib; best = h; ib = (uint16_t)(t + 1); }
        else if (h > sbest) { sbest = h; isb = (uint16_t)(t + 1); }
        sum = (uint16_t)(sum + h);
        if (row && n < (uint32_t)p.row_pairs) { row[1 + 2 * n] = (uint16_t)t; row[2 + 2 * n] = h; }
        n++;
    }
    if (row) { row[0] = (uint16_t)n; if (n > (uint32_t)p.row_pairs) atomicAdd(&p.counters[COUNTER_TRUNC], 1u); }
    if (p.final5) {
        uint16_t* f = p.final5 + (size_t)read * 5;
        f[0] = sum; f[1] = ib; f[2] = best; f[3] = isb; f[4] = sbest;
    }
}

}  // namespace cuclark
