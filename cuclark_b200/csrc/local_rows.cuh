// cuclark_b200 — LOCAL layout: the minimizer front end for one row of 32 k-mers, as a function.
//
// k_classify<LOCAL> (classify.cu) carries this computation inline for ILP rows per group; this is its one-row form,
// used by the table-partitioned scatter kernel (route.cu), which needs every k-mer's line (-> owner shard), sector and
// key but probes nothing itself. Both must give the same (line, key) for every k-mer: the parity tests hold each of
// them to the oracle (tests/test_gpu_local_layout.py, tests/test_gpu_route.py).
//
// Each lane hashes the FIRST m-mer of its position in row i and in row i+1 (the windows of a row reach 7 positions into
// the next one; the next row's hashes are handed over to the next call instead of being recomputed); a windowed
// minimum over 8 consecutive positions by doubling across lanes gives every k-mer its minimizer — the LEFTMOST smallest
// hash in read orientation (common.cuh, "TIES").
#pragma once
#include "common.cuh"
#include "kmerwin.cuh"

namespace cuclark {

struct LocalCarry {                  // the row handed over from the previous call (reset per chunk of 32 words)
    bool have = false;
    uint64_t z = 0, c = 0;           // z bit 63/62: m-mer strand flags, bit 61: the k-mer stands in its canonical form
    uint32_t oh = 0;
};

struct LocalProbe {
    uint64_t c;                      // canonical k-mer
    uint64_t key;                    // 37-bit slot key (without the alt bit)
    uint32_t zq, line;               // mix(minimizer) div / mod NL: line is the GLOBAL A line
    int o_c;                         // offset of the minimizer in the canonical k-mer (sector = o_c & 3)
};

// W: this lane's 32 nucleotides of the chunk; i: row (k-mers 32 i + lane); m_limit = L - m - cb - lane
__device__ __forceinline__ LocalProbe local_row(LocalCarry& S, uint64_t W, int i, int m_limit, int k, int kshift, uint32_t NL,
                                                uint32_t nl_m32, int nl_sh, int lane) {
    const int m = k - LOCAL_W + 1, mbits = 2 * m;
    const uint64_t mmask = (~0ull) >> (64 - mbits);
    uint64_t cs = 0, zs[2];
    uint32_t AL[2];
#pragma unroll
    for (int R = 0; R < 2; R++) {
        const int ir = i + R;
        uint64_t c;
        uint32_t oh;
        if (R == 0 && S.have) {                      // warp-uniform
            c = S.c; zs[0] = S.z; oh = S.oh;
        } else {
            const uint64_t hi = shfl64(W, ir & 31), lo = shfl64(W, (ir + 1) & 31);
            const uint64_t x = window64(hi, lo, 2 * lane) >> kshift;
            const uint64_t rc = revcomp2(x, k);
            const uint64_t a = x >> (2 * (LOCAL_W - 1)), b = rc & mmask;
            const bool lt = a < b;
            const uint64_t z = local_mix(lt ? a : b, mbits);
            oh = local_order(z, mbits);
            if (ir > 31 || 32 * ir > m_limit) oh = LOCAL_ORDER_MAX + 1;      // no m-mer here
            const bool kf = x <= rc;
            c = kf ? x : rc;
            zs[R] = z | ((uint64_t)lt << 63) | ((uint64_t)(!lt && a != b) << 62) | ((uint64_t)kf << 61);
        }
        if (R == 0) cs = c;
        else { S.c = c; S.z = zs[1]; S.oh = oh; }
        AL[R] = (oh << 8) | (uint32_t)(32 * R + lane);
    }
    S.have = true;
#pragma unroll
    for (int lvl = 0; lvl < 3; lvl++) {              // windows 2, 4, 8
        const int sft = 1 << lvl;
        const int src = (lane + sft) & 31;
        const bool wrap = lane + sft >= 32;
        const uint32_t t0 = __shfl_sync(0xFFFFFFFFu, AL[0], src), t1 = __shfl_sync(0xFFFFFFFFu, AL[1], src);
        AL[0] = min(AL[0], wrap ? t1 : t0);
        AL[1] = min(AL[1], t1);                      // (lanes near 31 of the second row take wrapped values: nobody reads them)
    }
    LocalProbe P;
    P.c = cs;
    const bool is_fwd = (zs[0] >> 61) & 1ull;
    const uint32_t pos = AL[0] & 255u;               // lane + offset (may reach into the second row)
    const int src = (int)(pos & 31u);
    const uint64_t z0 = shfl64(zs[0], src), z1 = shfl64(zs[1], src);
    const uint64_t zf = (pos >> 5) == 0u ? z0 : z1;
    const int o_read = (int)((pos - (uint32_t)lane) & 7u);
    local_divmod(zf & ((1ull << 61) - 1), NL, nl_m32, nl_sh, P.zq, P.line);
    P.o_c = is_fwd ? o_read : LOCAL_W - 1 - o_read;
    const bool f = (zf >> (is_fwd ? 63 : 62)) & 1ull;
    const uint32_t rest = local_rest(cs, P.o_c, m);
    P.key = (uint64_t)local_key_lo(P.zq, rest) | ((uint64_t)local_key_hi(rest, P.o_c, f) << 32);
    return P;
}

}  // namespace cuclark
