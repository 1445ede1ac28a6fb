/*
 * TEST INFRASTRUCTURE ONLY (oracle/). Never linked into or called from the product.
 *
 * Thin C wrapper around the UNMODIFIED reference host hash table, compiled
 * against the reference headers where they lie under /root/reference/src
 * (see oracle/Makefile; no reference source is copied into this repo).
 *
 * It gives the ground truth for stage 3 of the hot path:
 *   EHashtable<HKMERr, rElement>::Read          src/HashTableStorage_hh.hh:146-152
 *     -> hTable::read                           src/hashTable_hh.hh:666-946
 *   EHashtable::queryElement(uint64, label)     src/HashTableStorage_hh.hh:128-131
 *     -> hTable::find(const uint64_t&, ILBL&)   src/hashTable_hh.hh:476-513
 * and doubles as host baseline B2: OpenMP over k-mers.
 *
 * The variant (HTSIZE, MAXHITS) is fixed at compile time by which
 * parameters.hh the include path resolves to, exactly as src/Makefile:29-34
 * does; two shared objects are built: libref_lookup_full.so / _light.so.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <iostream>
#include <sstream>
#include <fstream>
#include <unistd.h>
#include <sys/types.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <sys/mman.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "HashTableStorage_hh.hh"

namespace {
struct AnyTable {
    int key_bytes;
    EHashtable<uint16_t, rElement>* t16;
    EHashtable<uint32_t, rElement>* t32;
    EHashtable<uint64_t, rElement>* t64;
};
}  // namespace

extern "C" {

/* HTSIZE / MAXHITS the object was compiled with (to assert the variant). */
uint64_t ref_htsize(void) { return (uint64_t)HTSIZE; }
int ref_maxhits(void) { return MAXHITS; }

/* Key width the reference CLI would pick for this k (src/main.cc:278-316). */
int ref_key_bytes(int k) {
    size_t t_b = log(HTSIZE) / log(4.0);
    if ((size_t)k <= t_b + 8) return 2;
    if ((size_t)k <= t_b + 16) return 4;
    return 8;
}

void* ref_db_open(const char* base, int k, int sfactor, int threads) {
    std::vector<std::string> none;
    AnyTable* a = new AnyTable();
    a->key_bytes = ref_key_bytes(k);
    a->t16 = NULL; a->t32 = NULL; a->t64 = NULL;
    size_t fsz = 0;
    bool ok = false;
    if (a->key_bytes == 2) {
        a->t16 = new EHashtable<uint16_t, rElement>(k, none, none);
        ok = a->t16->Read(base, fsz, threads, sfactor, false);
    } else if (a->key_bytes == 4) {
        a->t32 = new EHashtable<uint32_t, rElement>(k, none, none);
        ok = a->t32->Read(base, fsz, threads, sfactor, false);
    } else {
        a->t64 = new EHashtable<uint64_t, rElement>(k, none, none);
        ok = a->t64->Read(base, fsz, threads, sfactor, false);
    }
    if (!ok) { delete a->t16; delete a->t32; delete a->t64; delete a; return NULL; }
    return a;
}

/* labels[i] = label index of k-mer i, or -1 on a miss. kmers are in either
 * orientation (find() canonicalises). Returns the number of hits. */
long ref_db_query(void* h, const uint64_t* kmers, long n, int32_t* labels, int threads) {
    AnyTable* a = (AnyTable*)h;
    long hits = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) reduction(+ : hits) schedule(static)
#endif
    for (long i = 0; i < n; i++) {
        ILBL l = 0;
        bool f;
        if (a->t32) f = a->t32->queryElement(kmers[i], l);
        else if (a->t16) f = a->t16->queryElement(kmers[i], l);
        else f = a->t64->queryElement(kmers[i], l);
        labels[i] = f ? (int32_t)l : -1;
        hits += f;
    }
    return hits;
}

void ref_db_close(void* h) {
    AnyTable* a = (AnyTable*)h;
    if (!a) return;
    delete a->t16; delete a->t32; delete a->t64;
    delete a;
}

}  // extern "C"
