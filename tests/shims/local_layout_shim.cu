// CPU test shim: the LOCAL layout's address arithmetic (csrc/common.cuh, __host__ __device__) checked on the host.
//   * local_rebuild(local_locate(c)) == c and the key fits 37 bits;
//   * the classify kernel's route to the same home — minimizer chosen in READ orientation (leftmost smallest
//     hash for a forward-canonical k-mer, rightmost for a reverse-canonical one), mixed m-mer, strand flags —
//     gives the same (sector, key) as local_locate on the canonical k-mer, also for periodic sequences whose
//     k-mers hold the same m-mer several times and for palindromic m-mers.
#include <stdint.h>
#include <stdio.h>

#include "../../cuclark_b200/csrc/common.cuh"

using namespace cuclark;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

extern "C" long shim_local_check(int k, uint64_t NL, long n, int* first_bad) {
    const int m = k - LOCAL_W + 1;
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    const uint64_t mmask = (~0ull) >> (64 - 2 * m);
    long ties = 0;
    for (long t = 0; t < n; t++) {
        uint64_t x = rnd() & kmask;
        if (t % 4 == 0) {                                   // periodic (period 1..8 nt), sometimes with one change
            const int period = 1 + (int)(rnd() % 8);
            const uint64_t unit = rnd() & ((1ull << (2 * period)) - 1);
            x = 0;
            for (int i = 0; i < 64; i += 2 * period) x = (x << (2 * period)) | unit;
            x &= kmask;
            if (t % 8 == 0) x ^= 1ull << (rnd() % (2 * k));
        }
        const uint64_t rc = revcomp2(x, k);
        const bool is_fwd = x <= rc;
        const uint64_t c = is_fwd ? x : rc;
        uint64_t sec, key;
        local_locate(c, k, NL, sec, key);
        if (key >> 37) { *first_bad = 1; return t; }
        if (local_rebuild(sec >> 2, key, k, NL) != c) { *first_bad = 2; return t; }
        // the kernel: leftmost smallest hash in READ orientation
        uint32_t best = 0xFFFFFFFFu;
        int offL = 0, offR = 0;
        for (int o = 0; o < LOCAL_W; o++) {
            const uint64_t a = (x >> (2 * (LOCAL_W - 1 - o))) & mmask, b = revcomp2(a, m);
            const uint32_t h = local_order(local_mix(a < b ? a : b, 2 * m), 2 * m);
            if (h < best) { best = h; offL = offR = o; } else if (h == best) offR = o;
        }
        const int o_read = offL;
        const uint64_t a = (x >> (2 * (LOCAL_W - 1 - o_read))) & mmask, b = revcomp2(a, m);
        const uint64_t z = local_mix(a < b ? a : b, 2 * m);
        if (local_unmix(z, 2 * m) != (a < b ? a : b)) { *first_bad = 3; return t; }
        const int o_c = is_fwd ? o_read : LOCAL_W - 1 - o_read;
        const bool f = is_fwd ? a < b : a > b;
        const uint32_t rest = local_rest(c, o_c, m);
        const uint64_t q = (uint64_t)local_key_lo(z / NL, rest) | ((uint64_t)local_key_hi(rest, o_c, f) << 32);
        const uint64_t qsec = (z % NL) * 4 + (uint64_t)(o_c & 3);
        // the builder: one home, or two for a tie (then the k-mer lives in the overflow table and both homes are flagged)
        uint64_t sl, kl, sr, kr;
        const bool tie = local_locate_both(c, k, NL, sl, kl, sr, kr);
        if (tie != (offL != offR)) { *first_bad = 11; return t; }
        if (sl != sec || kl != key) { *first_bad = 12; return t; }
        ties += tie;
        if (!tie) {
            if (q != key || qsec != sec) { *first_bad = 4; return t; }
        } else if (!((q == kl && qsec == sl) || (q == kr && qsec == sr))) { *first_bad = 13; return t; }
        if (local_rebuild(qsec >> 2, q, k, NL) != c) { *first_bad = 14; return t; }
        // second candidate line: inside the shard, never the A line itself, and invertible from (B line, key)
        {
            const uint64_t line_lo = (NL / 3), line_n64 = NL - line_lo - (NL / 5);        // a shard in the middle
            uint64_t sa, sb, k2;
            local_locate2(c, k, NL, line_lo, (uint32_t)line_n64, sa, sb, k2);
            if (sa != sec || k2 != key) { *first_bad = 6; return t; }
            const uint64_t la = sa >> 2;
            if (la - line_lo < line_n64) {
                const uint64_t lbn = sb >> 2;
                if (lbn == la || lbn - line_lo >= line_n64 || (sb & 3) != (sa & 3)) { *first_bad = 7; return t; }
                const uint64_t altkey = key | ((uint64_t)LOCAL_ALT_BIT << 32);
                if (local_rebuild2(lbn, altkey, k, NL, line_lo, (uint32_t)line_n64) != c) { *first_bad = 8; return t; }
                if (local_rebuild2(la, key, k, NL, line_lo, (uint32_t)line_n64) != c) { *first_bad = 9; return t; }
            } else if (sb != sa) { *first_bad = 10; return t; }
        }
        // the kernel's 32-bit divmod
        const int sh = 2 * m > 32 ? 2 * m - 32 : 0;
        const uint32_t m32 = (uint32_t)((((__uint128_t)1) << (32 + sh)) / NL);
        uint32_t dq, dr;
        local_divmod(z, (uint32_t)NL, m32, sh, dq, dr);
        if (dq != z / NL || dr != z % NL) { *first_bad = 5; return t; }
    }
    *first_bad = 0;
    return ties;
}
