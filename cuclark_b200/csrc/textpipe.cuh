// cuclark_b200 — device-side text stages: raw FASTA/FASTQ bytes -> read index ->
// 2-bit containers, and per-read results -> CSV text. Prototypes shared by
// textpipe.cu (kernels) and stream.cu (host pipeline).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cuclark {

constexpr uint32_t TP_ERR_LINES = 1u;      // more lines than the slot can index
constexpr uint32_t TP_ERR_READS = 2u;      // more reads than the slot can hold
constexpr uint32_t TP_ERR_CONT = 4u;       // packed reads exceed the container buffer
constexpr uint32_t TP_ERR_CSV = 8u;        // CSV text exceeds the output buffer

// Written by the device, mirrored into pinned host memory after each stage.
struct ChunkInfo {
    uint32_t n_newlines;
    uint32_t n_headers;      // lines starting with '>'
    uint32_t n_lines;
    uint32_t n_reads;
    uint32_t err;
    uint32_t pad;
    uint64_t n_cont;         // containers of the packed chunk
    uint64_t csv_bytes;      // bytes of the CSV group last formatted
};

// Device arrays of one pipeline slot (capacities fixed at allocation).
struct TextSlotDev {
    uint8_t* text = nullptr;         // cap_bytes (+ padding)
    uint32_t* line_start = nullptr;  // cap_lines + 2
    uint32_t* hdr_line = nullptr;    // cap_reads + 1 (FASTA)
    uint32_t* name_s = nullptr;      // cap_reads each, offsets into text
    uint32_t* name_e = nullptr;
    uint32_t* seq_s = nullptr;
    uint32_t* seq_e = nullptr;
    uint32_t* len = nullptr;         // the Length column
    uint32_t* reads_ptr = nullptr;   // cap_reads + 1: counts, scanned in place into container offsets
    uint16_t* cont = nullptr;        // cap_cont
    uint16_t* final5 = nullptr;      // cap_reads * 5
    uint16_t* rows = nullptr;        // cap_reads * pitch (extended only)
    uint32_t* csv_off = nullptr;     // cap_reads + 1
    char* csv = nullptr;             // cap_csv
    uint32_t* tile_a = nullptr;      // scan scratch
    uint32_t* tile_b = nullptr;
    ChunkInfo* info = nullptr;
    size_t cap_bytes = 0, cap_lines = 0, cap_reads = 0, cap_cont = 0, cap_csv = 0, cap_tiles = 0;
};

// Target names on the device: names[0] = "NA", names[t + 1] = label t.
struct NameTable {
    const char* chars = nullptr;
    const uint32_t* off = nullptr;   // n_names + 1
    uint32_t n_names = 0;
    uint32_t max_len = 0;
};

// stage 1a: line table + per-read name/sequence spans (src/CuCLARK_hh.hh:1340-1534)
int tp_index_launch(const TextSlotDev& s, uint32_t n_bytes, bool fastq, cudaStream_t st);
// stage 1b: containers (src/CuCLARK_hh.hh:1616-1708); n_reads is the value read back from info
int tp_pack_launch(const TextSlotDev& s, uint32_t n_bytes, uint32_t n_reads, int k, cudaStream_t st);
// stage 4': CSV lines of reads [first, first + n) (src/CuCLARK_hh.hh:1951-2139)
int tp_csv_launch(const TextSlotDev& s, const NameTable& names, uint32_t first, uint32_t n, int k, bool paired,
                  bool extended, int row_pairs, uint32_t n_targets, cudaStream_t st);

}  // namespace cuclark
