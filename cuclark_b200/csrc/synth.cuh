// Counter-based synthetic data, device twin of cuclark_b200/synth.py.
// Every base is a pure function of (seed, target, position); every read of
// (seed, read index). Keep the constants in sync with synth.py.
#pragma once
#include <stdint.h>

namespace cuclark {
namespace synth {

constexpr uint64_t TAG_GENOME = 0x47, TAG_READ = 0x52, TAG_RBASE = 0x62, TAG_SUB = 0x73;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__host__ __device__ __forceinline__ uint64_t key(uint64_t tag, uint32_t seed, uint64_t a, uint64_t b) {
    uint64_t h = mix64((tag << 56) ^ ((uint64_t)(seed & 0xFFFFu) << 40) ^ b);
    return mix64(h ^ (a * 0x9E3779B97F4A7C15ull));
}

// 32 bases (2 bit each, base p%32 in bits 2*(p%32)) of word w of a genome
__host__ __device__ __forceinline__ uint64_t genome_word(uint32_t seed, uint32_t target, uint64_t w) {
    return key(TAG_GENOME, seed, target, w);
}

__host__ __device__ __forceinline__ uint32_t genome_base(uint32_t seed, uint32_t target, uint64_t p) {
    return (uint32_t)(genome_word(seed, target, p >> 5) >> (2 * (p & 31))) & 3u;
}

}  // namespace synth
}  // namespace cuclark
