/*
 * cuclark_b200.h — C ABI of libcuclark_b200.so
 *
 * B200-native replacement for the GPU driver class of CuCLARK,
 * `CuClarkDB<HKMERr>` (reference: src/CuClarkDB.cuh:98-150, implementation
 * src/CuClarkDB.cu:84-1033, kernels src/CuClarkDB.cu:1045-1471), i.e. the seam
 * between the host orchestrator (src/CuCLARK_hh.hh) and the device:
 * load the target-specific k-mer database, take batches of 2-bit packed reads,
 * return per-read (sum, best, hits, second, hits) and, on request, the sparse
 * per-target hit rows.
 *
 * Plain C types only. Every function returns CUCLARK_OK (0) or a negative
 * error code; cuclark_last_error() gives the message of the calling thread's
 * last failure. There is NO CPU fallback: without a CUDA device
 * cuclark_create() fails with CUCLARK_ERR_NO_DEVICE.
 *
 * Data formats are the reference's own:
 *   - database files  <base>.sz / .ky / .lb        src/hashTable_hh.hh:591-663
 *   - packed reads    readsPointer[n+1] (uint32 container offsets) +
 *                     containers (uint16: per part one length header, then
 *                     8 nt per container, MSB first, code A=3 C=2 G=1 T=0)
 *                                                  src/CuCLARK_hh.hh:1616-1708
 *   - sparse row      uint16[2*row_pairs+2] = [n, (target, hits)*n] ascending
 *                                                  src/CuClarkDB.cu:1181-1243
 *   - final result    uint16[5] = [sum, best+1, hits, second+1, hits]
 *                                                  src/CuClarkDB.cu:1421-1471
 */
#ifndef CUCLARK_B200_H
#define CUCLARK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUCLARK_OK 0
#define CUCLARK_ERR_ARG (-1)
#define CUCLARK_ERR_NO_DEVICE (-2)
#define CUCLARK_ERR_CUDA (-3)
#define CUCLARK_ERR_IO (-4)        /* a database file is missing/short: CuClarkDB::read returns false */
#define CUCLARK_ERR_NOMEM (-5)
#define CUCLARK_ERR_STATE (-6)     /* call order violated (no DB loaded, batch not ready, ...) */
#define CUCLARK_ERR_BUILD (-7)     /* device table could not be built (bucket overflow) */
#define CUCLARK_ERR_FORMAT (-8)    /* input is neither FASTA nor FASTQ (src/CuCLARK_hh.hh:1535-1538) */

/* reference constants (src/parameters.hh:38-48, src/parameters_light_hh:39-49) */
#define CUCLARK_HTSIZE_FULL 1610612741ull
#define CUCLARK_HTSIZE_LIGHT 57777779ull
#define CUCLARK_MAXHITS_FULL 15
#define CUCLARK_MAXHITS_LIGHT 23

typedef struct cuclark_db cuclark_db;

/* Replaces the CuClarkDB constructor arguments (src/CuClarkDB.cu:85-108) and
 * the compile-time variant (HTSIZE/MAXHITS by header swap, src/Makefile:26-34),
 * which is a run-time choice here. */
typedef struct cuclark_config {
    int k;                /* k-mer length, 2..32 (src/main.cc:107-112)                       */
    uint64_t htsize;      /* HTSIZE of the database FILES (full or light)                    */
    int key_bytes;        /* width of a .ky element: 2, 4, 8; 0 = as src/main.cc:278-316     */
    int n_targets;        /* number of labels (CuClarkDB ctor _numTargets)                   */
    int row_pairs;        /* MAXHITS: (target,hits) pairs per sparse row; 0 = by htsize      */
    int device;           /* CUDA device ordinal                                             */
    int shard_index;      /* table-partitioned mode: this handle holds shard i of n          */
    int shard_count;      /*   (contiguous ranges of device buckets); 1 = whole table        */
    double bucket_load;   /* mean entries per 32-byte bucket; 0 = default                    */
    int layout;           /* 0 auto, 1 narrow (5 x 32-bit key), 2 wide (3 x 64-bit key),     */
                          /* 3 local (minimizer-addressed 128-byte lines of 4 x 4 37-bit     */
                          /* keys, two candidate lines per minimizer; k >= 19 and a table of */
                          /* at least 4^(k-7)/2^19 lines; auto takes it for single-device    */
                          /* tables that fill that minimum; env CUCLARK_LAYOUT=1|2|3 decides */
                          /* for handles that pass 0, CUCLARK_NO_LOCAL=1 keeps auto hashed)  */
} cuclark_config;

typedef struct cuclark_stats {
    uint64_t n_entries;        /* entries held by this shard                                 */
    uint64_t n_buckets;        /* global bucket count M                                      */
    uint64_t n_local_buckets;  /* buckets held by this shard                                 */
    uint64_t table_bytes;      /* device bytes of this shard's table                         */
    uint64_t n_spilled;        /* entries stored outside their home bucket                   */
    uint64_t n_spill_buckets;  /* buckets with maxdisp > 0                                   */
    int layout;                /* 1 narrow, 2 wide, 3 local                                  */
    int k;
    uint64_t lookups;          /* k-mers looked up by the last classify call                 */
    uint64_t dense_reads;      /* reads of the last call that took the dense fallback        */
    uint64_t truncated_rows;   /* rows of the last call with more than row_pairs targets     */
    double last_kernel_ms;     /* device time of the last classify (CUDA events)             */
} cuclark_stats;

const char* cuclark_last_error(void);
int cuclark_version(void);
/* Kernels the hot path (classification, row merge, text pipeline, k-mer routing) has launched in this process since
 * the library was loaded; bench.py reports the difference over its timed region as `gpu_launches`. */
uint64_t cuclark_kernel_launches(void);

/* CuClarkDB::CuClarkDB / ~CuClarkDB (src/CuClarkDB.cu:85-208, 215-260) */
int cuclark_create(const cuclark_config* cfg, cuclark_db** out);
int cuclark_destroy(cuclark_db* db);

/* CuClarkDB::read + swapDbParts + sync (src/CuClarkDB.cu:462-858): loads
 * <base>.sz/.ky/.lb, applies -s sampling (:511-524), re-buckets on the device.
 * CUCLARK_ERR_IO if a file cannot be opened (the reference returns false). */
int cuclark_load_db_files(cuclark_db* db, const char* base, int sfactor);
/* same from host arrays (sz: htsize bytes; ky: n_entries keys of key_bytes; lb) */
int cuclark_load_db_arrays(cuclark_db* db, const uint8_t* sz, const void* ky, const uint16_t* lb,
                           uint64_t n_entries, int sfactor);
/* The table of `src` copied device to device into `dst` (same configuration, another device): a replicated
 * multi-GPU run reads and re-buckets the files once and sends the replicas over NVLink instead of rebuilding the
 * table per device (the reference never replicates: src/CuClarkDB.cu:546-574). */
int cuclark_clone_table(cuclark_db* src, cuclark_db* dst);
/* CuClarkDB::CuClarkDB device enumeration (src/CuClarkDB.cu:108-138): number of CUDA devices and, for `device`,
 * its free and total memory (each pointer may be NULL). CUCLARK_ERR_NO_DEVICE without a GPU. */
int cuclark_device_info(int device, int* n_devices, uint64_t* free_bytes, uint64_t* total_bytes);
/* Synthetic database built ON the device (bench.py at BASELINE sizes): all
 * canonical k-mers of n_targets seeded random genomes of genome_len bases
 * (stride 1 = every overlapping k-mer, as the full variant; the light variant
 * samples every `stride`-th non-overlapping k-mer), label = target, k-mers seen
 * in more than one target removed (RemoveCommon, src/HashTableStorage_hh.hh:242-292). */
int cuclark_build_db_synthetic(cuclark_db* db, uint32_t seed, uint32_t n_targets, uint64_t genome_len,
                               int light_gap);

/* ---- table cache -----------------------------------------------------------
 * The reference re-reads .sz/.ky/.lb and rebuilds its bucket pointers on every start
 * (CuClarkDB::read, src/CuClarkDB.cu:462-808). cuclark_save_table writes the loaded table in its
 * DEVICE layout (192-byte header, sector buckets, overflow table; checksummed) so that a later
 * cuclark_load_table streams it back file -> pinned buffer -> HBM without rebuilding.
 * cuclark_load_table returns CUCLARK_ERR_IO if `path` cannot be opened and CUCLARK_ERR_FORMAT if
 * the file is not a cache, is corrupt, was written for another k / HTSIZE / n_targets / shard /
 * sampling factor, or (src_base != NULL) if <src_base>.sz/.ky/.lb are not the files it was
 * built from (compared by size); the caller then falls back to cuclark_load_db_files. */
int cuclark_save_table(cuclark_db* db, const char* path);
int cuclark_load_table(cuclark_db* db, const char* path, const char* src_base, int sfactor);

/* What cuclark_load_db_* would build for a database of n_entries k-mers under cfg, without touching a device:
 * the layout the automatic choice takes, the bucket count and the bytes of the home table (the overflow table
 * comes on top: ~3 % of the entries at 27 B each). For sizing a run and for the reference's dbParts logic
 * (src/CuClarkDB.cu:543-574 decides the number of swap cycles from the same figures; here it is always 1). */
typedef struct cuclark_table_plan {
    int layout;                /* 1 narrow, 2 wide, 3 local                                  */
    uint64_t n_buckets;        /* global 32-byte buckets                                     */
    uint64_t n_local_buckets;  /* buckets of this shard (cfg.shard_index of cfg.shard_count) */
    uint64_t home_bytes;       /* n_local_buckets * 32                                       */
} cuclark_table_plan;
int cuclark_plan_table(const cuclark_config* cfg, uint64_t n_entries, cuclark_table_plan* out);

int cuclark_get_stats(cuclark_db* db, cuclark_stats* out);
/* Fetch the counters (lookups, dense_reads, truncated_rows) of the last
 * cuclark_classify_device call; synchronises `stream` (NULL = library stream). */
int cuclark_sync_stats(cuclark_db* db, void* stream);

/* ---- batches: CuClarkDB::malloc / readyBatch / queryBatch / waitForBatch /
 * freeBatchMemory (src/CuClarkDB.cu:318-460, 861-1033). The library owns
 * pinned host buffers, the caller fills them (pack loop) and reads results. */
int cuclark_batches_alloc(cuclark_db* db, int n_batches, size_t max_reads, size_t max_containers,
                          int want_rows);
int cuclark_batch_buffers(cuclark_db* db, int batch, uint32_t** reads_ptr, uint16_t** containers,
                          uint16_t** final5, uint16_t** rows);
int cuclark_batch_ready(cuclark_db* db, int batch, size_t n_reads, size_t n_containers);
int cuclark_batch_query(cuclark_db* db, int batch);   /* async: H2D, kernels, D2H, event */
int cuclark_batch_wait(cuclark_db* db, int batch);
int cuclark_batches_free(cuclark_db* db);

/* One-shot convenience over host buffers (pageable or pinned). rows may be NULL. */
int cuclark_classify_host(cuclark_db* db, const uint32_t* reads_ptr, const uint16_t* containers,
                          size_t n_reads, uint16_t* final5, uint16_t* rows);

/* Device-resident inputs and outputs (all pointers are device pointers on
 * cfg.device; stream is a cudaStream_t or NULL). d_final5 and d_rows may each
 * be NULL. With shard_count > 1 only rows are meaningful (merge them first). */
int cuclark_classify_device(cuclark_db* db, const uint32_t* d_reads_ptr, const uint16_t* d_containers,
                            size_t n_reads, uint16_t* d_final5, uint16_t* d_rows, void* stream);

/* mergeKernel + resultKernel (src/CuClarkDB.cu:1321-1471) generalised to n_parts
 * shards: d_rows_parts holds n_parts consecutive row arrays [part][read][pitch].
 * Writes merged rows (may be NULL) and final results (may be NULL). */
int cuclark_merge_rows_device(cuclark_db* db, const uint16_t* d_rows_parts, int n_parts, size_t n_reads,
                              uint16_t* d_rows_out, uint16_t* d_final5, void* stream);

/* ---- table-partitioned mode by k-mer routing ---------------------------------------------------------
 * The reference's multi-GPU mode (`-d N`) partitions the table by bucket range, copies every read batch to
 * every device (src/CuClarkDB.cu:546-574, 886-895) and merges the per-device rows (:953-974): every device
 * looks at every k-mer. Here rank g of N holds shard g (cfg.shard_index / shard_count, hashed layout) AND only
 * its own reads: the canonical k-mers of the reads are bucketed by owner shard in g's HBM (scatter), every
 * shard probes the k-mers addressed to it straight out of the peers' HBM over NVLink and stores the 2-byte
 * labels straight back (probe), and g counts the labels per read (gather). The result equals the
 * single-table result bit for bit. All ranks must pass a barrier between scatter and probe and between
 * probe and gather (stream-ordered NCCL all-reduce between processes, events between the devices of one
 * process — cuclark_classify_routed_device does it for the handles of one process). */
#define CUCLARK_ROUTE_MAX_RANKS 16
typedef struct cuclark_route_stats {
    int n_ranks, rank;
    uint64_t region_bytes;     /* arena + labels + block lists of this rank (peer-visible)               */
    uint64_t map_bytes;        /* position map (local)                                                   */
    uint64_t cap_blocks;       /* arena blocks of 256 entries                                            */
    uint64_t lookups;          /* k-mers of this rank's reads in the last scatter                        */
    uint64_t probed;           /* k-mers this shard probed for all ranks in the last probe               */
    uint64_t blocks;           /* arena blocks used by the last scatter                                  */
    uint64_t blocks_remote;    /* of which addressed to other ranks: x 256 x (8 + 2) bytes cross NVLink  */
    uint32_t err;
} cuclark_route_stats;
/* Buffers for calls of up to max_containers containers (<= 2^28); the same value on every rank. */
int cuclark_route_alloc(cuclark_db* db, int n_ranks, size_t max_containers);
int cuclark_route_free(cuclark_db* db);
/* One process per GPU: a 64-byte CUDA IPC handle of this rank's region, to be opened by the other ranks. */
int cuclark_route_export(cuclark_db* db, void* handle64, uint64_t* region_bytes);
int cuclark_route_import(cuclark_db* db, int peer_rank, const void* handle64);
/* One process, N handles (in rank order; devices may differ or coincide): peer access + plain pointers. */
int cuclark_route_connect(cuclark_db* const* dbs, int n);
int cuclark_route_scatter(cuclark_db* db, const uint32_t* d_reads_ptr, const uint16_t* d_containers, size_t n_reads,
                          size_t n_containers, void* stream);
int cuclark_route_probe(cuclark_db* db, void* stream);
int cuclark_route_gather(cuclark_db* db, const uint32_t* d_reads_ptr, const uint16_t* d_containers, size_t n_reads,
                         size_t n_containers, uint16_t* d_final5, uint16_t* d_rows, void* stream);
/* figures of the last scatter/probe; valid once the gather's stream is idle (cuclark_sync_stats) */
int cuclark_route_get_stats(cuclark_db* db, cuclark_route_stats* out);
/* The three steps for the N connected handles of ONE process, with the barriers between them (CUDA events across
 * the devices): per rank its own device-resident reads and outputs. This is what the CLI's -d N runs when the
 * table does not fit one device. */
int cuclark_classify_routed_device(cuclark_db* const* dbs, int n, const uint32_t* const* d_reads_ptr,
                                   const uint16_t* const* d_containers, const size_t* n_reads, const size_t* n_containers,
                                   uint16_t* const* d_final5, uint16_t* const* d_rows);

/* ---- database construction on the device -------------------------------------------------------
 * makeSpecificTargetSets + EHashtable::addElement + RemoveCommon + hTable::write
 * (src/CuCLARK_hh.hh:691-1112, src/HashTableStorage_hh.hh:242-292 and 484-523,
 * src/hashTable_hh.hh:591-663): scans the FASTA target files, takes every overlapping k-mer
 * (light_gap = 0, cuCLARK) or every light_gap-th non-overlapping one (cuCLARK-l), keeps the
 * canonical k-mers that occur under exactly one label more than min_count times and writes
 * <out_base>.sz/.ky/.lb byte for byte as the reference does. target_labels[i] is the label index
 * of file i (labels numbered in order of first appearance in the targets file). */
typedef struct cuclark_build_opts {
    int k;
    uint64_t htsize;      /* CUCLARK_HTSIZE_FULL or CUCLARK_HTSIZE_LIGHT                               */
    int key_bytes;        /* 0 = as src/main.cc:278-316                                               */
    int light_gap;        /* 0 = full variant; >= 1: -g of cuCLARK-l (default there: 4)               */
    uint32_t min_count;   /* -t                                                                       */
    int device;
} cuclark_build_opts;
typedef struct cuclark_build_stats {
    uint64_t n_nucleotides;   /* ACGTU bytes outside header lines                                     */
    uint64_t n_kmers_added;   /* addElement calls                                                     */
    uint64_t n_kmers_kept;    /* entries written                                                      */
    int key_bytes;
} cuclark_build_stats;
int cuclark_build_database(const cuclark_build_opts* opts, const char* const* target_files,
                           const uint16_t* target_labels, size_t n_files, const char* out_base,
                           cuclark_build_stats* stats);

/* ---- text in, CSV out: CuCLARK::getObjectsDataComputeFullGPU + printExtendedResultsSynced --------
 * (src/CuCLARK_hh.hh:1335-1790 and 1951-2139). The reference indexes reads (:1340-1534) and packs
 * them (:1616-1708) on the host and prints one CSV line per read with fprintf (:2097-2136). Here
 * the RAW FASTA/FASTQ bytes are copied to the device in chunks cut at record boundaries and
 * indexing, 2-bit packing, classification and CSV formatting (including printf's "%g") all run
 * there; the sink receives the CSV text in file order, starting with the header line. */
typedef struct cuclark_text_opts {
    int paired;                        /* Length column = Length - 1: mates joined by one N (:2112)  */
    int extended;                      /* --extended: one hit-count column per target (:2014-2031)   */
    const char* const* target_names;   /* n_targets labels in label order (NULL: "T<i>")             */
    size_t chunk_bytes;                /* text bytes per chunk; 0 = 4 MiB (16 MiB above 4 GiB input) */
    int n_slots;                       /* chunks in flight (host threads, streams); 0 = 4            */
} cuclark_text_opts;

typedef struct cuclark_text_stats {
    uint64_t n_reads;
    uint64_t lookups;
    uint64_t n_containers;
    uint64_t text_bytes;               /* H2D payload                                                */
    uint64_t csv_bytes;                /* D2H payload (+ header)                                     */
    uint64_t n_chunks;
    uint64_t dense_reads;
    uint64_t truncated_rows;
    double seconds;                    /* wall time of the call                                      */
} cuclark_text_stats;

/* Receives the CSV text piece by piece: `data[0..n)` belongs at byte `offset` of the result. Pieces
 * may arrive from several threads at once and out of order (their ranges are disjoint and together
 * cover the file); the header line comes first. Return 0 to continue, non-zero to abort
 * (-> CUCLARK_ERR_IO). */
typedef int (*cuclark_sink_fn)(void* user, const char* data, size_t n, uint64_t offset);

/* `text` is host memory (pinned memory is copied from directly, pageable memory is staged). */
int cuclark_classify_text(cuclark_db* db, const uint8_t* text, size_t n, const cuclark_text_opts* opts,
                          cuclark_sink_fn sink, void* user, cuclark_text_stats* out);
/* Same, into a caller-provided host buffer; a pinned buffer receives the device's text directly. */
int cuclark_classify_text_buffer(cuclark_db* const* dbs, int n_dbs, const uint8_t* text, size_t n,
                                 const cuclark_text_opts* opts, char* out, size_t out_cap, size_t* out_len,
                                 cuclark_text_stats* stats);
/* CuCLARK::runSimple (src/CuCLARK_hh.hh:512-573): mmap `objects_path`, write `csv_path`.
 * CUCLARK_ERR_IO if the input is missing or empty ("Failed to open"). */
int cuclark_classify_file(cuclark_db* db, const char* objects_path, const char* csv_path,
                          const cuclark_text_opts* opts, cuclark_text_stats* out);

/* Read-partitioned multi-GPU in one process (the CLI's -d N): one handle per device, each holding
 * the WHOLE table; chunks of the input are handed to whichever device has a free slot, the CSV
 * stays in file order. (The reference's -d N partitions the table instead, src/CuClarkDB.cu:546-574;
 * that mode is cfg.shard_count + cuclark_merge_rows_device.) */
int cuclark_classify_text_multi(cuclark_db* const* dbs, int n_dbs, const uint8_t* text, size_t n,
                                const cuclark_text_opts* opts, cuclark_sink_fn sink, void* user,
                                cuclark_text_stats* out);
int cuclark_classify_file_multi(cuclark_db* const* dbs, int n_dbs, const char* objects_path, const char* csv_path,
                                const cuclark_text_opts* opts, cuclark_text_stats* out);

/* Same pipeline, but instead of CSV text the intermediate arrays come back (host pointers, each
 * may be NULL): the read index (absolute byte offsets into `text`), the packed reads in the
 * reference's format, and the per-read results. For parity tests of the device-side stages. */
typedef struct cuclark_text_arrays {
    size_t cap_reads;                  /* capacity of the per-read arrays (reads_ptr: +1)            */
    size_t cap_containers;
    uint64_t *name_s, *name_e, *seq_s, *seq_e, *len;
    uint32_t* reads_ptr;
    uint16_t* containers;
    uint16_t* final5;
    uint16_t* rows;
} cuclark_text_arrays;
int cuclark_text_debug(cuclark_db* db, const uint8_t* text, size_t n, const cuclark_text_opts* opts,
                       cuclark_text_arrays* arrays, cuclark_text_stats* out);

/* ---- synthetic reads generated on the device (bench.py): packed format ---- */
/* Fills d_reads_ptr[n_reads+1] and d_containers (n_reads * (1+ceil(read_len/8)))
 * with reads sampled from the synthetic genomes of cuclark_build_db_synthetic. */
int cuclark_synth_reads_device(cuclark_db* db, uint32_t seed, uint32_t genome_seed, uint32_t n_targets,
                               uint64_t genome_len, uint64_t first_read, size_t n_reads, int read_len,
                               int pct_random, int sub_per_10k, uint32_t* d_reads_ptr,
                               uint16_t* d_containers, void* stream);

/* The same reads as 4-line FASTQ text, fixed-width records of 16 + 2*read_len bytes:
 * "@r<9 digits>\n<bases>\n+\n<'I' x read_len>\n". d_text holds n_reads records. */
int cuclark_synth_fastq_device(cuclark_db* db, uint32_t seed, uint32_t genome_seed, uint32_t n_targets,
                               uint64_t genome_len, uint64_t first_read, size_t n_reads, int read_len,
                               int pct_random, int sub_per_10k, uint8_t* d_text, void* stream);

/* The two files of a paired-end run (BASELINE configs[2]): mate = 1 writes the reads above, mate = 2 their mates —
 * same id, same target, 2 x read_len further along, opposite strand, own substitutions. */
int cuclark_synth_fastq_pair_device(cuclark_db* db, uint32_t seed, uint32_t genome_seed, uint32_t n_targets,
                                    uint64_t genome_len, uint64_t first_read, size_t n_reads, int read_len,
                                    int pct_random, int sub_per_10k, int mate, uint8_t* d_text, void* stream);

/* ---- roofline probe: random 32-byte-sector gather over this table ---------- */
/* Reads n_probes uniformly random sectors of the loaded table per launch
 * (ilp independent loads per thread); returns average ms over `iters` launches. */
int cuclark_gather_bench(cuclark_db* db, uint64_t n_probes, int bytes_per_probe, int ilp, int iters,
                         double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* CUCLARK_B200_H */
