// cuclark_b200 — k-mer windows of a packed part, shared by the warp-per-read kernels (stage 2).
//
// Replaces the per-thread re-assembly of queryKernel (src/CuClarkDB.cu:1090-1135: every thread concatenates up
// to 8 containers from shared memory): lane j of the warp keeps 32 nucleotides of the current part as one 64-bit
// word, MSB first; the k-mer that starts at nucleotide 32 i + lane is a funnel shift of the words of lanes i and
// i+1 fetched by shuffle. No shared-memory staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cuclark {

constexpr int CHUNK_ROUNDS = 31;      // rounds of 32 k-mers served by one set of 32 words

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

// top 64 bits of the 128-bit value (hi:lo) << s, 0 <= s <= 62: the 32 nucleotides that start s/2 nucleotides into
// word hi. Two 32-bit funnel shifts (which take the amount mod 32) over operands picked by s >= 32.
__device__ __forceinline__ uint64_t window64(uint64_t hi, uint64_t lo, int s) {
    const uint32_t h1 = (uint32_t)(hi >> 32), h0 = (uint32_t)hi, l1 = (uint32_t)(lo >> 32), l0 = (uint32_t)lo;
    const bool big = s >= 32;
    const uint32_t a = big ? h0 : h1, b = big ? l1 : h0, c = big ? l0 : l1;
    return ((uint64_t)__funnelshift_l(b, a, s) << 32) | __funnelshift_l(c, b, s);
}

// 64-bit words (32 nt each, MSB first) of up to 128 consecutive containers held as four
// 32-lane windows: lane j gets containers 4j..4j+3. Only the first nwin windows are live.
__device__ __forceinline__ uint64_t assemble_words(const uint32_t (&wv)[4], int nwin, int lane) {
    uint64_t W = 0;
    const int src = 4 * (lane & 7);
#pragma unroll
    for (int u = 0; u < 4; u++) {
        if (u < nwin) {                                   // warp-uniform
            uint64_t w = 0;
#pragma unroll
            for (int t = 0; t < 4; t++)
                w |= (uint64_t)__shfl_sync(0xFFFFFFFFu, wv[u], src + t) << (48 - 16 * t);
            if ((lane >> 3) == u) W = w;
        }
    }
    return W;
}

// A part header that lies about its part (a foreign packer: the reference's own uint16 header wraps at 65,536 nt)
// must not carry the loads past the read's containers: the part is clamped to the read.
// first = index of the part's first data container; returns its container count, L = its nucleotides.
__device__ __forceinline__ uint32_t part_extent(uint32_t header, uint32_t first, uint32_t end, uint32_t& L) {
    const uint32_t ncont = min((header + 7) >> 3, end - first);
    L = min(header, 8u * ncont);
    return ncont;
}

}  // namespace cuclark
