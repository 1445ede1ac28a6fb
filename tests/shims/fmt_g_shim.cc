// Test shim: exposes the product's host/device %g formatter (cuclark_b200/csrc/fmt_g.h) to ctypes.
#include "../../cuclark_b200/csrc/fmt_g.h"

extern "C" {
int shim_fmt_g(double d, char* out) { int n = cuclark::fmt_g(d, out); out[n] = 0; return n; }
int shim_gamma(unsigned total, unsigned norm, int k, char* out) { int n = cuclark::fmt_g(cuclark::csv_gamma(total, norm, k), out); out[n] = 0; return n; }
int shim_conf(unsigned best, unsigned sbest, char* out) { int n = cuclark::fmt_g(cuclark::csv_confidence(best, sbest), out); out[n] = 0; return n; }
// bulk: every a/b with a in [a0,a1), b in [b0,b1): returns number of mismatches against snprintf("%g")
long shim_sweep(unsigned a0, unsigned a1, unsigned b0, unsigned b1, unsigned* bad_a, unsigned* bad_b) {
    long bad = 0;
    char x[32], y[32];
    for (unsigned a = a0; a < a1; a++)
        for (unsigned b = b0; b < b1; b++) {
            const double d = (double)a / (double)b;
            int n = cuclark::fmt_g(d, x); x[n] = 0;
            __builtin_snprintf(y, sizeof y, "%g", d);
            if (__builtin_strcmp(x, y)) { if (!bad) { *bad_a = a; *bad_b = b; } bad++; }
        }
    return bad;
}
}
