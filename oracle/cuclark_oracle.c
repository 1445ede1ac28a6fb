/*
 * oracle/cuclark_oracle.c — TEST INFRASTRUCTURE ONLY (see cuclark_oracle.h).
 *
 * CPU restatement of the CuCLARK classification hot path, written from the
 * reference's behaviour. Citations are relative to /root/reference.
 * Plain C11 + OpenMP. Scalar and deliberately simple: this is the checker,
 * not the product, and the product never links or loads it.
 */
#define _GNU_SOURCE
#include "cuclark_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ===== nucleotide classes ==================================================
 * src/CuCLARK_hh.hh:263-300 (m_table / m_rTable / m_separators / m_Letter):
 * ACGT/acgt and U/u are nucleotides; the packed code is the COMPLEMENT code
 * A=3 C=2 G=1 T/U=0 (m_rTable).                                              */
static int nt_rcode(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 3;
        case 'C': case 'c': return 2;
        case 'G': case 'g': return 1;
        case 'T': case 't': case 'U': case 'u': return 0;
        default: return -1;
    }
}
static int is_sep(uint8_t c) { return c == ' ' || c == '\t' || c == '\n'; }
static int is_letter(uint8_t c) { return (c >= 65 && c < 91) || (c >= 97 && c < 123); }

/* ===== canonical k-mer =======================================================
 * src/hashTable_hh.hh:478-489 == src/CuClarkDB.cu:1255-1266: reverse the 2-bit
 * groups over 64 bit, complement, shift down to 2k bits; canonical = min.      */
static uint64_t revcomp(uint64_t x, int k) {
    uint64_t r = 0;
    for (int j = 0; j < 32; j++) {           /* reverse order of 2-bit groups */
        r = (r << 2) | (x & 3u);
        x >>= 2;
    }
    return (~r) >> (64 - 2 * k);
}
uint64_t orc_canonical(uint64_t kmer, int k) {
    uint64_t r = revcomp(kmer, k);
    return kmer < r ? kmer : r;
}

/* ===== table ================================================================
 * File format: src/hashTable_hh.hh:591-663 (.sz one uint8 per bucket, .ky
 * quotients bucket-major ascending, .lb uint16 labels).
 * Load-time sampling: src/CuClarkDB.cu:497-524 — among non-empty buckets
 * numbered 1,2,3,... keep those with n % s == 0 (s <= 1 keeps all).           */
orc_db* orc_db_from_arrays(uint64_t htsize, int k, int key_bytes, const uint8_t* sz,
                           const void* ky, const uint16_t* lb, int sfactor) {
    orc_db* db = (orc_db*)calloc(1, sizeof(orc_db));
    db->htsize = htsize; db->k = k; db->key_bytes = key_bytes;
    db->start = (uint64_t*)malloc((htsize + 1) * sizeof(uint64_t));
    uint64_t kept = 0, nonzero = 0;
    for (uint64_t r = 0; r < htsize; r++) {
        db->start[r] = kept;
        if (sz[r]) {
            nonzero++;
            if (sfactor <= 1 || (nonzero % (uint64_t)sfactor) == 0) kept += sz[r];
        }
    }
    db->start[htsize] = kept;
    db->n = kept;
    db->keys = (uint64_t*)malloc((kept ? kept : 1) * sizeof(uint64_t));
    db->labels = (uint16_t*)malloc((kept ? kept : 1) * sizeof(uint16_t));
    uint64_t src = 0;
    nonzero = 0;
    for (uint64_t r = 0; r < htsize; r++) {
        if (!sz[r]) continue;
        nonzero++;
        int keep = (sfactor <= 1 || (nonzero % (uint64_t)sfactor) == 0);
        if (keep) {
            uint64_t dst = db->start[r];
            for (unsigned j = 0; j < sz[r]; j++) {
                uint64_t q;
                if (key_bytes == 2) q = ((const uint16_t*)ky)[src + j];
                else if (key_bytes == 4) q = ((const uint32_t*)ky)[src + j];
                else q = ((const uint64_t*)ky)[src + j];
                db->keys[dst + j] = q;
                db->labels[dst + j] = lb[src + j];
            }
        }
        src += sz[r];
    }
    return db;
}

static void* slurp(const char* path, size_t* size) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void* p = malloc(n > 0 ? (size_t)n : 1);
    if (n > 0 && fread(p, 1, (size_t)n, f) != (size_t)n) { free(p); fclose(f); return NULL; }
    fclose(f);
    *size = (size_t)n;
    return p;
}

orc_db* orc_db_load(const char* base, uint64_t htsize, int k, int key_bytes, int sfactor) {
    size_t L = strlen(base) + 8, s1, s2, s3;
    char* p = (char*)malloc(L);
    snprintf(p, L, "%s.sz", base); uint8_t* sz = (uint8_t*)slurp(p, &s1);
    snprintf(p, L, "%s.ky", base); void* ky = slurp(p, &s2);
    snprintf(p, L, "%s.lb", base); uint16_t* lb = (uint16_t*)slurp(p, &s3);
    free(p);
    orc_db* db = NULL;
    if (sz && ky && lb && s1 == htsize && s2 / key_bytes == s3 / 2)
        db = orc_db_from_arrays(htsize, k, key_bytes, sz, ky, lb, sfactor);
    free(sz); free(ky); free(lb);
    return db;
}

/* ---- synthetic genomes: twin of cuclark_b200/synth.py (mix64/_key/genome word) ---- */
static uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static uint64_t synth_key(uint64_t tag, uint32_t seed, uint64_t a, uint64_t b) {
    uint64_t h = mix64((tag << 56) ^ ((uint64_t)(seed & 0xFFFFu) << 40) ^ b);
    return mix64(h ^ (a * 0x9E3779B97F4A7C15ull));
}
static unsigned synth_base(uint32_t seed, uint32_t t, uint64_t p) {
    return (unsigned)(synth_key(0x47, seed, t, p >> 5) >> (2 * (p & 31))) & 3u;
}

typedef struct { uint64_t key; uint16_t label; } kl_t;

static void radix_sort_kl(kl_t* a, kl_t* tmp, size_t n) {
    for (int pass = 0; pass < 8; pass++) {
        size_t cnt[257] = {0};
        int sh = 8 * pass;
        for (size_t i = 0; i < n; i++) cnt[((a[i].key >> sh) & 0xFF) + 1]++;
        if (cnt[1] == n) continue;                 /* byte is all zero */
        for (int b = 0; b < 256; b++) cnt[b + 1] += cnt[b];
        for (size_t i = 0; i < n; i++) tmp[cnt[(a[i].key >> sh) & 0xFF]++] = a[i];
        memcpy(a, tmp, n * sizeof(kl_t));
    }
}

/* Builder semantics: src/CuCLARK_hh.hh:694-767 (light scanner), :896-975 (full
 * scanner), src/HashTableStorage_hh.hh:484-523 (addElement, canonical = min) and
 * :242-292 (RemoveCommon: keep k-mers of exactly one target); file layout
 * src/hashTable_hh.hh:591-663. */
orc_db* orc_db_build_synth(uint32_t seed, uint32_t n_targets, uint64_t genome_len, int k, uint64_t htsize,
                           int key_bytes, int light_gap, int threads, const char* write_base) {
    uint64_t per = light_gap > 0 ? ((genome_len / k) + light_gap - 1) / light_gap : genome_len - k + 1;
    size_t n = (size_t)per * n_targets;
    kl_t* a = (kl_t*)malloc(n * sizeof(kl_t));
    kl_t* tmp = (kl_t*)malloc(n * sizeof(kl_t));
    uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (uint32_t t = 0; t < n_targets; t++) {
        kl_t* out = a + (size_t)t * per;
        if (light_gap > 0) {
            for (uint64_t j = 0; j < per; j++) {
                uint64_t start = j * (uint64_t)light_gap * k, R = 0;
                for (int i = 0; i < k; i++) R = (R << 2) | (3u - synth_base(seed, t, start + i));
                uint64_t c = orc_canonical(R, k);
                out[j].key = ((c % htsize) << 32) | (c / htsize);     /* needs quotient < 2^32 */
                out[j].label = (uint16_t)t;
            }
        } else {
            uint64_t R = 0, word = 0;
            for (uint64_t p = 0; p < genome_len; p++) {
                if ((p & 31) == 0) word = synth_key(0x47, seed, t, p >> 5);
                R = ((R << 2) | (3u - ((unsigned)(word >> (2 * (p & 31))) & 3u))) & mask;
                if (p + 1 >= (uint64_t)k) {
                    uint64_t c = orc_canonical(R, k);
                    out[p + 1 - k].key = ((c % htsize) << 32) | (c / htsize);
                    out[p + 1 - k].label = (uint16_t)t;
                }
            }
        }
    }
    radix_sort_kl(a, tmp, n);
    free(tmp);
    /* unique + RemoveCommon */
    size_t m = 0;
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        int common = 0;
        while (j < n && a[j].key == a[i].key) { common |= a[j].label != a[i].label; j++; }
        if (!common) a[m++] = a[i];
        i = j;
    }
    orc_db* db = (orc_db*)calloc(1, sizeof(orc_db));
    db->htsize = htsize; db->k = k; db->key_bytes = key_bytes; db->n = m;
    db->start = (uint64_t*)calloc(htsize + 1, sizeof(uint64_t));
    db->keys = (uint64_t*)malloc((m ? m : 1) * sizeof(uint64_t));
    db->labels = (uint16_t*)malloc((m ? m : 1) * sizeof(uint16_t));
    for (size_t i = 0; i < m; i++) {
        db->start[(a[i].key >> 32) + 1]++;
        db->keys[i] = a[i].key & 0xFFFFFFFFull;
        db->labels[i] = a[i].label;
    }
    free(a);
    if (write_base) {
        size_t L = strlen(write_base) + 8;
        char* p = (char*)malloc(L);
        uint8_t* sz = (uint8_t*)malloc(htsize);
        for (uint64_t r = 0; r < htsize; r++) sz[r] = (uint8_t)db->start[r + 1];
        snprintf(p, L, "%s.sz", write_base);
        FILE* f = fopen(p, "wb"); if (f) { fwrite(sz, 1, htsize, f); fclose(f); }
        free(sz);
        snprintf(p, L, "%s.ky", write_base);
        f = fopen(p, "wb");
        if (f) {
            for (size_t i = 0; i < m; i++) {
                if (key_bytes == 4) { uint32_t v = (uint32_t)db->keys[i]; fwrite(&v, 4, 1, f); }
                else if (key_bytes == 2) { uint16_t v = (uint16_t)db->keys[i]; fwrite(&v, 2, 1, f); }
                else fwrite(&db->keys[i], 8, 1, f);
            }
            fclose(f);
        }
        snprintf(p, L, "%s.lb", write_base);
        f = fopen(p, "wb"); if (f) { fwrite(db->labels, 2, m, f); fclose(f); }
        free(p);
    }
    for (uint64_t r = 0; r < htsize; r++) db->start[r + 1] += db->start[r];
    return db;
}

void orc_db_free(orc_db* db) {
    if (!db) return;
    free(db->start); free(db->keys); free(db->labels); free(db);
}
uint64_t orc_db_size(const orc_db* db) { return db->n; }

/* src/hashTable_hh.hh:476-513 (host) == src/CuClarkDB.cu:1249-1314 (device):
 * c = canonical; q = c / HTSIZE, r = c % HTSIZE; shard range test on r
 * (CuClarkDB.cu:1272); empty bucket or q outside [first,last] -> miss; linear
 * ascending scan while key <= q.                                               */
int orc_db_find(const orc_db* db, uint64_t kmer, uint64_t lo, uint64_t hi, uint16_t* label) {
    uint64_t c = orc_canonical(kmer, db->k);
    uint64_t q = c / db->htsize, r = c % db->htsize;
    if (r < lo || r >= hi) return 0;
    uint64_t b = db->start[r], e = db->start[r + 1];
    if (e == b) return 0;
    if (db->keys[b] > q || db->keys[e - 1] < q) return 0;
    for (uint64_t i = b; i < e && db->keys[i] <= q; i++)
        if (db->keys[i] == q) { *label = db->labels[i]; return 1; }
    return 0;
}

long orc_db_query(const orc_db* db, const uint64_t* kmers, long n, int32_t* labels, int threads) {
    long hits = 0;
#pragma omp parallel for num_threads(threads) reduction(+ : hits) schedule(static)
    for (long i = 0; i < n; i++) {
        uint16_t l = 0;
        int f = orc_db_find(db, kmers[i], 0, db->htsize, &l);
        labels[i] = f ? (int32_t)l : -1;
        hits += f;
    }
    return hits;
}

void orc_db_entries(const orc_db* db, uint64_t* kmers, uint16_t* labels) {
    for (uint64_t r = 0; r < db->htsize; r++)
        for (uint64_t i = db->start[r]; i < db->start[r + 1]; i++) {
            kmers[i] = db->keys[i] * db->htsize + r;
            labels[i] = db->labels[i];
        }
}

/* ===== read indexing ======================================================== */
static void ix_push(orc_index* ix, size_t ns, size_t ne, size_t ss, size_t se, size_t len) {
    if (ix->n == ix->cap) {
        ix->cap = ix->cap ? ix->cap * 2 : 1024;
        ix->name_s = (size_t*)realloc(ix->name_s, ix->cap * sizeof(size_t));
        ix->name_e = (size_t*)realloc(ix->name_e, ix->cap * sizeof(size_t));
        ix->seq_s = (size_t*)realloc(ix->seq_s, ix->cap * sizeof(size_t));
        ix->seq_e = (size_t*)realloc(ix->seq_e, ix->cap * sizeof(size_t));
        ix->len = (size_t*)realloc(ix->len, ix->cap * sizeof(size_t));
    }
    ix->name_s[ix->n] = ns; ix->name_e[ix->n] = ne;
    ix->seq_s[ix->n] = ss; ix->seq_e[ix->n] = se; ix->len[ix->n] = len;
    ix->n++;
}

/* Byte read with the bound the reference lacks (it may read one byte past the
 * mmap when a file ends inside a header; we treat that byte as NUL).           */
static uint8_t at(const uint8_t* m, size_t nb, size_t i) { return i < nb ? m[i] : 0; }

/* FASTA: src/CuCLARK_hh.hh:1340-1404. Batch r starts scanning at byte
 * r*(1+nb/n_batches), skips to the next '>', and owns records until it stands
 * on a '>' at or beyond its upper bound. Name = bytes after '>' up to the
 * first separator (the first name byte is never tested, :1369). Sequence =
 * everything up to the last newline before the next '>'; Length = bytes minus
 * one per line (:1377-1389) — an empty sequence therefore has Length 1.        */
static void index_fasta(const uint8_t* m, size_t nb, size_t n_batches, orc_index* ix) {
    size_t step = 1 + nb / n_batches;
    for (size_t r = 0; r < n_batches; r++) {
        ix->batch_first[r] = ix->n;
        size_t i = step * r, hi = step * (r + 1);
        while (i < nb && m[i] != '>') i++;
        if (i >= nb) continue;                 /* reference would run off the map */
        i++;
        for (;;) {
            size_t ns = i;
            while (i < nb && !is_sep(at(m, nb, ++i))) {}
            size_t ne = i;
            while (i < nb && m[i++] != '\n') {}
            size_t ss = i, se = i, lines = 0;
            while (i < nb && m[i] != '>') {
                while (i < nb && m[i] != '\n') i++;
                lines++;
                se = i++;
            }
            ix_push(ix, ns, ne, ss, se, se - ss + 1 - lines);
            if ((i >= hi && at(m, nb, i) == '>') || i >= nb) break;
            i++;
        }
    }
    ix->batch_first[n_batches] = ix->n;
}

/* FASTQ: src/CuCLARK_hh.hh:1405-1534. Batch starts are snapped to a record by
 * looking at the six line starts after the byte stride: a line starting with
 * '@' whose next line is letters only and whose line after that starts with
 * '+' (:1430-1471). Records are strictly four lines; Length = bytes of line 2. */
static size_t next_line(const uint8_t* m, size_t nb, size_t i) {
    while (i < nb && m[i++] != '\n') {}
    return i;
}
static void index_fastq(const uint8_t* m, size_t nb, size_t n_batches, orc_index* ix) {
    size_t step = 1 + nb / n_batches;
    size_t* pos = (size_t*)calloc(n_batches + 1, sizeof(size_t));
    pos[0] = 1;
    for (size_t r = 1; r < n_batches; r++) {
        size_t p[7];
        p[0] = step * r;
        for (int j = 1; j <= 6; j++) p[j] = next_line(m, nb, p[j - 1]);
        for (int j = 1; j <= 4; j++) {
            if (at(m, nb, p[j]) != '@') continue;
            size_t i = p[j + 1];
            while (is_letter(at(m, nb, i++))) {}
            if (i == p[j + 2] && at(m, nb, i) == '+') { pos[r] = p[j] + 1; break; }
        }
    }
    for (size_t r = 0; r < n_batches; r++) {
        ix->batch_first[r] = ix->n;
        size_t inext = r + 1 < n_batches ? pos[r + 1] : nb;
        size_t i = pos[r];
        if (i == 0 || i >= nb) continue;      /* unsnapped / empty batch */
        for (;;) {
            size_t ns = i;
            while (i < nb && !is_sep(at(m, nb, ++i))) {}
            size_t ne = i;
            while (i < nb && m[i++] != '\n') {}
            size_t ss = i;
            while (i < nb && m[i] != '\n') i++;
            size_t se = i++;
            ix_push(ix, ns, ne, ss, se, se - ss);
            i = next_line(m, nb, i);
            i = next_line(m, nb, i);
            if (++i >= inext) break;
        }
    }
    ix->batch_first[n_batches] = ix->n;
    free(pos);
}

int orc_index_reads(const uint8_t* map, size_t nb, size_t n_batches, orc_index* out) {
    memset(out, 0, sizeof(*out));
    if (n_batches < 1) n_batches = 1;
    out->n_batches = n_batches;
    out->batch_first = (size_t*)calloc(n_batches + 1, sizeof(size_t));
    if (nb == 0) return -1;
    if (map[0] == '>') index_fasta(map, nb, n_batches, out);
    else if (map[0] == '@') index_fastq(map, nb, n_batches, out);
    else return -1;                            /* :1535-1538 */
    return 0;
}

void orc_index_free(orc_index* ix) {
    free(ix->name_s); free(ix->name_e); free(ix->seq_s); free(ix->seq_e); free(ix->len);
    free(ix->batch_first);
    memset(ix, 0, sizeof(*ix));
}

/* ===== packing ==============================================================
 * src/CuCLARK_hh.hh:1616-1708. Per read with Length >= k the bytes
 * [seq_s, seq_e) are split into maximal runs of nucleotides; '\n' is
 * transparent, any other byte ends the run. Each run of >= k nucleotides
 * becomes a "part": one header container holding the run length followed by
 * ceil(len/8) containers, 8 nt each, MSB first, complement code, last one
 * left-aligned. Shorter runs leave nothing behind (the reference overwrites
 * them, :1646-1660 and :1699-1703).
 * DEVIATION Q8: the header is a uint16 sum, so in the reference it
 * wraps for runs of 65,536 nt or more and its kernel then reads data containers
 * as headers (undefined results). Here — oracle and device packer alike — such a
 * run becomes several parts of at most 65,535 nt that overlap by k-1 nt, so every
 * k-mer of the run is still looked up exactly once.                             */
#define ORC_MAX_PART 65535u
size_t orc_pack_bound(const orc_index* ix, size_t first, size_t n) {
    size_t tot = 0;
    for (size_t i = first; i < first + n; i++) {
        size_t bytes = ix->seq_e[i] - ix->seq_s[i];
        tot += bytes / 8 + 2 + bytes / 16 + 2; /* generous: header per possible part */
        tot += (bytes / ORC_MAX_PART + 1) * 8; /* overlap of split parts */
    }
    return tot + 8;
}

size_t orc_pack(const uint8_t* map, const orc_index* ix, size_t first, size_t n, int k,
                uint32_t* reads_ptr, uint16_t* cont) {
    size_t cc = 0;
    for (size_t r = 0; r < n; r++) {
        size_t g = first + r;
        reads_ptr[r] = (uint32_t)cc;
        if (ix->len[g] < (size_t)k) continue;                    /* :1633 */
        size_t i = ix->seq_s[g], e = ix->seq_e[g];
        while (i < e) {
            /* skip to the start of a run */
            while (i < e && nt_rcode(map[i]) < 0) i++;
            if (i >= e) break;
            size_t hdr = cc++;
            size_t runlen = 0;
            size_t recent[32];                   /* byte positions of the last nucleotides */
            uint16_t word = 0;
            unsigned fill = 0;
            while (i < e) {
                int c = nt_rcode(map[i]);
                if (c >= 0) {
                    if (runlen == ORC_MAX_PART) {
                        /* full part and the run goes on: the next part starts k-1 nucleotides back */
                        i = recent[(runlen - (size_t)(k - 1)) & 31];
                        break;
                    }
                    recent[runlen & 31] = i;
                    word = (uint16_t)((word << 2) | (unsigned)c);
                    if (++fill == 8) { cont[cc++] = word; fill = 0; word = 0; }
                    runlen++;
                    i++;
                } else if (map[i] == '\n') {
                    i++;
                } else {
                    break;
                }
            }
            if (fill) cont[cc++] = (uint16_t)(word << (2 * (8 - fill)));
            if (runlen < (size_t)k) cc = hdr;            /* drop the short run */
            else cont[hdr] = (uint16_t)runlen;
        }
    }
    reads_ptr[n] = (uint32_t)cc;
    return cc;
}

/* ===== extraction ===========================================================
 * src/CuClarkDB.cu:1090-1135. Every window of k consecutive nucleotides of a
 * part is one k-mer; its integer is the 2k-bit window of the MSB-first 2-bit
 * stream (first base in the HIGH bits, complement code) = the "R form".        */
static uint64_t part_kmer(const uint16_t* c, size_t p, int k) {
    uint64_t v = 0;
    for (int j = 0; j < k; j++) {
        size_t q = p + (size_t)j;
        unsigned nt = (c[q >> 3] >> (2 * (7 - (q & 7)))) & 3u;
        v = (v << 2) | nt;
    }
    return v;
}

uint64_t orc_extract(const uint32_t* reads_ptr, const uint16_t* cont, size_t n, int k, uint64_t* out) {
    uint64_t cnt = 0;
    for (size_t r = 0; r < n; r++) {
        uint32_t p = reads_ptr[r], e = reads_ptr[r + 1];
        while (p < e) {
            unsigned L = cont[p++];
            const uint16_t* c = cont + p;
            p += (L - 1) / 8 + 1;
            for (size_t w = 0; w + (size_t)k <= L; w++) {
                if (out) out[cnt] = part_kmer(c, w, k);
                cnt++;
            }
        }
    }
    return cnt;
}

/* ===== classify =============================================================
 * Per read: dense per-target counters (src/CuClarkDB.cu:1064-1074, 1156-1170),
 * sparse row of non-zero counters in ascending target order (:1181-1243; the
 * numTargets <= 32 uninitialised-count quirk, SURVEY.md A.7-Q3, is defined as
 * the true count), then the top-2 scan (resultKernel, :1421-1471).
 * Rows keep the first row_pairs pairs; row[0] is the true distinct count.      */
static void top2_scan(const uint32_t* hits, int n_targets, uint16_t* f5) {
    uint16_t best = 0, sbest = 0, ib = 0, isb = 0, sum = 0;
    for (int t = 0; t < n_targets; t++) {
        uint16_t h = (uint16_t)hits[t];
        if (!h) continue;
        if (h > best) { sbest = best; isb = ib; best = h; ib = (uint16_t)(t + 1); }
        else if (h > sbest) { sbest = h; isb = (uint16_t)(t + 1); }
        sum = (uint16_t)(sum + h);
    }
    f5[0] = sum; f5[1] = ib; f5[2] = best; f5[3] = isb; f5[4] = sbest;
}

uint64_t orc_classify(const orc_db* db, const uint32_t* reads_ptr, const uint16_t* cont, size_t n,
                      int n_targets, int row_pairs, uint64_t part_lo, uint64_t part_hi,
                      uint16_t* rows, uint16_t* final5, int threads) {
    uint64_t lookups = 0;
    int k = db->k;
    size_t pitch = (size_t)(2 * row_pairs + 2);
#pragma omp parallel num_threads(threads) reduction(+ : lookups)
    {
        uint32_t* hits = (uint32_t*)calloc((size_t)n_targets, sizeof(uint32_t));
#pragma omp for schedule(dynamic, 256)
        for (size_t r = 0; r < n; r++) {
            uint32_t p = reads_ptr[r], e = reads_ptr[r + 1];
            while (p < e) {
                unsigned L = cont[p++];
                const uint16_t* c = cont + p;
                p += (L - 1) / 8 + 1;
                /* rolling form of part_kmer(): same windows, same integers */
                uint64_t R = 0, mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
                for (size_t q = 0; q < L; q++) {
                    R = ((R << 2) | ((c[q >> 3] >> (2 * (7 - (q & 7)))) & 3u)) & mask;
                    if (q + 1 < (size_t)k) continue;
                    uint16_t lab;
                    lookups++;
                    if (orc_db_find(db, R, part_lo, part_hi, &lab) && lab < n_targets) hits[lab]++;
                }
            }
            top2_scan(hits, n_targets, final5 + 5 * r);
            uint16_t* row = rows ? rows + pitch * r : NULL;
            if (row) memset(row, 0, pitch * sizeof(uint16_t));
            unsigned cnt = 0;
            for (int t = 0; t < n_targets; t++) {
                if (!hits[t]) continue;
                if (row && cnt < (unsigned)row_pairs) { row[1 + 2 * cnt] = (uint16_t)t; row[2 + 2 * cnt] = (uint16_t)hits[t]; }
                cnt++;
                hits[t] = 0;
            }
            if (row) row[0] = (uint16_t)cnt;
        }
        free(hits);
    }
    return lookups;
}

/* src/CuClarkDB.cu:1321-1415: two-way merge of ascending sparse rows, summing
 * equal targets. Defined here for rows whose counts fit row_pairs.             */
void orc_merge_rows(const uint16_t* a, const uint16_t* b, size_t n, int row_pairs, uint16_t* out) {
    size_t pitch = (size_t)(2 * row_pairs + 2);
    for (size_t r = 0; r < n; r++) {
        const uint16_t* A = a + pitch * r; const uint16_t* B = b + pitch * r;
        uint16_t* O = out + pitch * r;
        unsigned na = A[0] < row_pairs ? A[0] : (unsigned)row_pairs;
        unsigned nb = B[0] < row_pairs ? B[0] : (unsigned)row_pairs;
        unsigned ia = 0, ib = 0, no = 0;
        memset(O, 0, pitch * sizeof(uint16_t));
        while (ia < na || ib < nb) {
            uint16_t t, h;
            if (ib >= nb || (ia < na && A[1 + 2 * ia] < B[1 + 2 * ib])) { t = A[1 + 2 * ia]; h = A[2 + 2 * ia]; ia++; }
            else if (ia >= na || B[1 + 2 * ib] < A[1 + 2 * ia]) { t = B[1 + 2 * ib]; h = B[2 + 2 * ib]; ib++; }
            else { t = A[1 + 2 * ia]; h = (uint16_t)(A[2 + 2 * ia] + B[2 + 2 * ib]); ia++; ib++; }
            if (no < (unsigned)row_pairs) { O[1 + 2 * no] = t; O[2 + 2 * no] = h; }
            no++;
        }
        O[0] = (uint16_t)no;
    }
}

/* src/CuClarkDB.cu:1421-1471 */
void orc_result_from_rows(const uint16_t* rows, size_t n, int row_pairs, uint16_t* final5) {
    size_t pitch = (size_t)(2 * row_pairs + 2);
    for (size_t r = 0; r < n; r++) {
        const uint16_t* R = rows + pitch * r;
        uint16_t best = 0, sbest = 0, ib = 0, isb = 0, sum = 0;
        unsigned cnt = R[0] < row_pairs ? R[0] : (unsigned)row_pairs;
        for (unsigned i = 0; i < cnt; i++) {
            uint16_t t = R[1 + 2 * i], h = R[2 + 2 * i];
            if (h > best) { sbest = best; isb = ib; best = h; ib = (uint16_t)(t + 1); }
            else if (h > sbest) { sbest = h; isb = (uint16_t)(t + 1); }
            sum = (uint16_t)(sum + h);
        }
        uint16_t* f = final5 + 5 * r;
        f[0] = sum; f[1] = ib; f[2] = best; f[3] = isb; f[4] = sbest;
    }
}

/* ===== CSV ==================================================================
 * src/CuCLARK_hh.hh:1951-2139. Header, then per read: name (<= 39 bytes),
 * [extended: one count per target in label order,] Length (paired: minus the
 * joining N), gamma = sum / (Length - k + 1), names and scores of the two best
 * targets ("NA" for none), confidence = h1 / (h1 + h2) (0 when both are 0),
 * doubles printed with %g.                                                     */
int orc_write_csv(const char* path, const uint8_t* map, const orc_index* ix, int k, int paired,
                  const char* const* names, int n_targets, const uint16_t* final5,
                  const uint16_t* rows, int row_pairs) {
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    fputs("Object_ID", f);
    if (rows) for (int t = 0; t < n_targets; t++) fprintf(f, ",%s", names[t]);
    fputs(",Length,Gamma,1st_assignment,score1,2nd_assignment,score2,confidence\n", f);
    size_t pitch = (size_t)(2 * row_pairs + 2);
    for (size_t r = 0; r < ix->n; r++) {
        const uint16_t* v = final5 + 5 * r;
        char name[40];
        size_t nl = ix->name_e[r] - ix->name_s[r];
        if (nl >= 40) nl = 39;
        memcpy(name, map + ix->name_s[r], nl);
        name[nl] = 0;
        fputs(name, f);   /* a NUL inside the name truncates it, as %s does */
        if (rows) {
            const uint16_t* R = rows + pitch * r;
            unsigned w = 0;
            for (unsigned i = 0; i < R[0] && i < (unsigned)row_pairs; i++) {
                for (; w < R[1 + 2 * i]; w++) fputs(",0", f);
                fprintf(f, ",%d", (int)R[2 + 2 * i]);
                w++;
            }
            for (; w < (unsigned)n_targets; w++) fputs(",0", f);
        }
        uint32_t norm = (uint32_t)(paired ? ix->len[r] - 1 : ix->len[r]);
        uint32_t total = v[0], i1 = v[1], best = v[2], i2 = v[3], sbest = v[4];
        double gamma = (double)total / (((double)norm - (double)k) + 1.0);
        double delta = (double)(best + sbest);
        delta = (delta < 0.001) ? 0 : ((double)best) / delta;
        fprintf(f, ",%u,%g,%s,%u,%s,%u,%g\n", norm, gamma, i1 ? names[i1 - 1] : "NA", best,
                i2 ? names[i2 - 1] : "NA", sbest, delta);
    }
    fclose(f);
    return 0;
}

/* ===== paired-end merge =====================================================
 * src/file.cc:205-268: FASTQ only; ids are the first token of the header split
 * on ' ', '/', '\t', '@' and must match; output ">id\n<seq1>N<seq2>\n".        */
static char* get_line(FILE* f) {
    char* line = NULL; size_t cap = 0;
    ssize_t n = getline(&line, &cap, f);
    if (n < 0) { free(line); return NULL; }
    if (n > 0 && line[n - 1] == '\n') line[n - 1] = 0;
    return line;
}
static void first_token(const char* s, char* out, size_t cap) {
    const char* seps = " /\t@";
    while (*s && strchr(seps, *s)) s++;
    size_t j = 0;
    while (*s && !strchr(seps, *s) && j + 1 < cap) out[j++] = *s++;
    out[j] = 0;
}
int orc_merge_paired(const char* f1, const char* f2, const char* outp) {
    FILE* a = fopen(f1, "r"); FILE* b = fopen(f2, "r");
    if (!a || !b) return -1;
    FILE* o = fopen(outp, "wb");
    if (!o) return -1;
    int rc = 0;
    char *l1, *l2;
    int c1 = fgetc(a), c2 = fgetc(b);
    if (c1 != c2) { fclose(a); fclose(b); fclose(o); return -4; }   /* different formats */
    if (c1 != '@') { fclose(a); fclose(b); fclose(o); return -5; }  /* must be FASTQ */
    rewind(a); rewind(b);
    while ((l1 = get_line(a)) && (l2 = get_line(b))) {
        if (l1[0] == '@' && l2[0] == '@') {
            char id1[4096], id2[4096];
            first_token(l1, id1, sizeof id1); first_token(l2, id2, sizeof id2);
            if (strcmp(id1, id2)) { rc = -2; free(l1); free(l2); break; }
            char* s1 = get_line(a); char* s2 = get_line(b);
            if (!s1 || !s2) { rc = -3; free(l1); free(l2); free(s1); free(s2); break; }
            fprintf(o, ">%s\n%sN%s\n", id1, s1, s2);
            free(s1); free(s2);
            for (int j = 0; j < 2; j++) { free(get_line(a)); free(get_line(b)); }
        }
        free(l1); free(l2);
    }
    fclose(a); fclose(b); fclose(o);
    return rc;
}
