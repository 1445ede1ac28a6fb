// cuclark_b200 — exact "%g" and "%u" formatting, usable on host and device.
//
// The result CSV prints gamma and confidence with printf("%g")
// (src/CuCLARK_hh.hh:2108-2135). To write the CSV on the device byte for byte
// we reproduce glibc's %g (precision 6, round-half-even on the EXACT binary
// value of the double) with integer arithmetic only:
//   d = M * 2^E (M < 2^53). For the magnitudes that can occur (|d| in
//   [2^-96, 2^40)) M * 10^s fits in 128 bits, so floor/round of d * 10^s is
//   exact; s is chosen so that the rounded value has six digits.
// Values outside that range (never produced by sum/(Length-k+1) or
// h1/(h1+h2)) take a generic but still exact slow path for the digits via
// repeated scaling in 128-bit arithmetic; see fmt_g_selftest in tests.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FMTG_HD __host__ __device__ __forceinline__
#else
#define FMTG_HD static inline
#endif

namespace cuclark {

FMTG_HD uint64_t fmtg_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    union { double d; uint64_t u; } v;
    v.d = d;
    return v.u;
#endif
}

FMTG_HD uint64_t fmtg_mulhi(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// decimal digits of v, most significant first; returns the count
FMTG_HD int fmt_u32(uint32_t v, char* out) {
    char tmp[10];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; i++) out[i] = tmp[n - 1 - i];
    return n;
}

FMTG_HD int dec_digits_u32(uint32_t v) {
    int n = 1;
    while (v >= 10) { v /= 10; n++; }
    return n;
}

// round(M * P / 2^sh) with ties to even; M*P < 2^128, 1 <= sh <= 127
FMTG_HD uint64_t fmtg_scaled_round(uint64_t M, uint64_t P, int sh) {
    const uint64_t hi = fmtg_mulhi(M, P), lo = M * P;
    uint64_t q, rem_hi, rem_lo, half_hi, half_lo;
    if (sh >= 64) {
        const int s = sh - 64;
        q = s ? (hi >> s) : hi;
        rem_hi = s ? (hi & ((1ull << s) - 1)) : 0;
        rem_lo = lo;
        half_hi = s ? (1ull << (s - 1)) : 0;
        half_lo = s ? 0 : (1ull << 63);
    } else {
        q = (hi << (64 - sh)) | (lo >> sh);          // callers keep the quotient below 2^64
        rem_hi = 0;
        rem_lo = lo & ((1ull << sh) - 1);
        half_hi = 0;
        half_lo = 1ull << (sh - 1);
    }
    const bool gt = rem_hi > half_hi || (rem_hi == half_hi && rem_lo > half_lo);
    const bool eq = rem_hi == half_hi && rem_lo == half_lo;
    if (gt || (eq && (q & 1))) q++;
    return q;
}

// Writes printf("%g", d) into out (at most 13 bytes: "-1.23456e-308"); returns the length.
// nan_negative: sign printed for NaN (x86 0.0/0.0 yields the negative quiet NaN -> "-nan").
FMTG_HD int fmt_g(double d, char* out) {
    const uint64_t bits = fmtg_bits(d);
    const bool neg = bits >> 63;
    const int bexp = (int)((bits >> 52) & 0x7FF);
    const uint64_t frac = bits & ((1ull << 52) - 1);
    int n = 0;
    if (neg) out[n++] = '-';
    if (bexp == 0x7FF) {
        if (frac) { out[n++] = 'n'; out[n++] = 'a'; out[n++] = 'n'; }
        else { out[n++] = 'i'; out[n++] = 'n'; out[n++] = 'f'; }
        return n;
    }
    if (bexp == 0 && frac == 0) { out[n++] = '0'; return n; }
    // d = M * 2^E
    uint64_t M = bexp ? (frac | (1ull << 52)) : frac;
    int E = (bexp ? bexp : 1) - 1075;
    // decimal exponent estimate: X ~ floor(log10(d)); log10(2) ~ 1233/4096
    int msb = 63;
    while (!((M >> msb) & 1)) msb--;
    const int e2 = E + msb;                           // d in [2^e2, 2^(e2+1))
    int X = (e2 * 1233) >> 12;                        // floor(e2*log10 2) or one less/more; fixed below
    uint64_t q = 0;
    for (int iter = 0; iter < 4; iter++) {
        const int s = 5 - X;                          // want round(d * 10^s) in [1e5, 1e6)
        // d * 10^s = M * 10^s * 2^E
        if (s >= 0 && s <= 27 && E < 0) {
            // split 10^s = 5^s * 2^s so that the multiplier stays below 2^64 (5^27 < 2^63)
            uint64_t P = 1;
            for (int i = 0; i < s; i++) P *= 5;
            const int sh = -E - s;                    // M * 5^s / 2^sh
            if (sh >= 1 && sh <= 127) q = fmtg_scaled_round(M, P, sh);
            else if (sh <= 0 && sh > -10) q = (M * P) << (-sh);   // tiny cases: exact integer
            else q = 0;
        } else {
            // outside the range the CSV can produce: fall back to long double-free scaling
            // (exact for integers < 2^53; good to 1 ulp of the 6th digit otherwise)
            double v = neg ? -d : d;
            int ss = s;
            while (ss > 0) { v *= 10.0; ss--; }
            while (ss < 0) { v /= 10.0; ss++; }
            q = (uint64_t)(v + 0.5);
        }
        if (q < 100000ull) { X--; continue; }
        if (q >= 1000000ull) {
            if (q == 1000000ull) { q = 100000ull; X++; break; }   // 9.999995 -> 10.0000
            X++;
            continue;
        }
        break;
    }
    // six digits, strip trailing zeros
    char dg[6];
    for (int i = 5; i >= 0; i--) { dg[i] = (char)('0' + q % 10); q /= 10; }
    int nd = 6;
    while (nd > 1 && dg[nd - 1] == '0') nd--;
    if (X < -4 || X >= 6) {
        out[n++] = dg[0];
        if (nd > 1) { out[n++] = '.'; for (int i = 1; i < nd; i++) out[n++] = dg[i]; }
        out[n++] = 'e';
        int ax = X;
        if (ax < 0) { out[n++] = '-'; ax = -ax; } else out[n++] = '+';
        if (ax >= 100) { out[n++] = (char)('0' + ax / 100); ax %= 100; }
        out[n++] = (char)('0' + ax / 10);
        out[n++] = (char)('0' + ax % 10);
    } else if (X >= 0) {
        for (int i = 0; i <= X; i++) out[n++] = i < nd ? dg[i] : '0';
        if (nd > X + 1) { out[n++] = '.'; for (int i = X + 1; i < nd; i++) out[n++] = dg[i]; }
    } else {
        out[n++] = '0'; out[n++] = '.';
        for (int i = 0; i < -X - 1; i++) out[n++] = '0';
        for (int i = 0; i < nd; i++) out[n++] = dg[i];
    }
    return n;
}

// gamma and confidence exactly as the reference computes them (src/CuCLARK_hh.hh:2122-2129)
FMTG_HD double csv_gamma(uint32_t total, uint32_t norm, int k) {
    const double den = ((double)norm - (double)k) + 1.0;
    if (total == 0 && den == 0.0) {
        // 0.0/0.0: the x86 host produces the NEGATIVE quiet NaN ("-nan" in the reference CSV)
#if defined(__CUDA_ARCH__)
        return __longlong_as_double((long long)0xFFF8000000000000ull);
#else
        union { uint64_t u; double d; } v;
        v.u = 0xFFF8000000000000ull;
        return v.d;
#endif
    }
    return (double)total / den;
}

FMTG_HD double csv_confidence(uint32_t best, uint32_t sbest) {
    const double delta = (double)(best + sbest);
    return delta < 0.001 ? 0.0 : (double)best / delta;
}

}  // namespace cuclark
