// CuClarkDB_b200.cc — the reference-side binding of libcuclark_b200.so.
//
// Same public interface as the reference's GPU driver class `CuClarkDB<HKMERr>` (src/CuClarkDB.cuh:98-150);
// every method forwards to the C ABI of include/cuclark_b200.h. It REPLACES src/CuClarkDB.cu in the reference's
// build (src/Makefile:10: CUCLARKCC = CuClarkDB_b200.cc main.cc analyser.cc file.cc kmersConversion.cc, plain
// g++ -fopenmp, link with -lcuclark_b200); nothing else of the reference tree changes: its orchestrator
// (src/CuCLARK_hh.hh) indexes and packs the reads, fills the pinned buffers this class hands out, and prints the
// CSV from the flat result arrays, exactly as before.
//
// The class declaration of the reference header is kept as it is, so the adapter's own state (the library handle,
// the flat result arrays) lives in a side table keyed by `this` instead of in new members.
// oracle/Makefile (`make adapter`) compiles this file against the UNMODIFIED reference sources where they lie;
// tests/test_adapter.py runs the reference's own orchestrator over the library and compares its CSV with the
// unmodified reference binary's.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <iostream>
#include <map>
#include <mutex>
#include <vector>

#include "CuClarkDB.cuh"
#include "cuclark_b200.h"

namespace {

struct B200State {
    cuclark_db* db = nullptr;
    size_t n_batches = 0;
    std::vector<uint16_t*> b_final, b_rows;         // the library's pinned per-batch result buffers
    std::vector<size_t> b_reads, b_first;            // reads of batch b, its first read in the flat arrays
    std::vector<char> b_copied;
    RESULTS *flat_final = nullptr, *flat_full = nullptr;
    size_t row_size = 0, final_row_size = 0;
    std::mutex mu;
};

std::mutex g_mu;
std::map<const void*, B200State*> g_state;

B200State* state_of(const void* self, bool create = false) {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_state.find(self);
    if (it != g_state.end()) return it->second;
    if (!create) { std::cerr << "CuClarkDB_b200: object used before construction" << std::endl; exit(1); }
    return g_state[self] = new B200State();
}

// CUERR / CUMEMERR of the reference (src/CuClarkDB.cu:45-63): message on stderr, exit(1)
void b200_ok(int rc, const char* what) {
    if (rc == CUCLARK_OK) return;
    std::cerr << "CUERR '" << cuclark_last_error() << "' in " << what << std::endl;
    exit(1);
}

}  // namespace

template <typename HKMERr>
CuClarkDB<HKMERr>::CuClarkDB() {}

// src/CuClarkDB.cu:85-208: numDevices == 0 means "all"; this adapter drives device 0 (the library's own command line
// spreads a run over devices, csrc/cli_main.cc), a request for more devices than present fails as the reference's does
template <typename HKMERr>
CuClarkDB<HKMERr>::CuClarkDB(const size_t _numDevices, const uint8_t _k, const size_t _numBatches, const size_t _numTargets) {
    m_k = _k; m_numTargets = _numTargets; m_numBatches = _numBatches;
    int n_dev = 0;
    std::cerr << "Checking for CUDA devices: ";
    if (cuclark_device_info(0, &n_dev, NULL, NULL) != CUCLARK_OK || (size_t)n_dev < _numDevices) {
        std::cerr << "Not enough CUDA devices found: " << n_dev << " of " << _numDevices << std::endl;
        exit(1);
    }
    m_numDevices = 1;
    std::cerr << n_dev << " found, using 1 through libcuclark_b200 " << cuclark_version() << std::endl;
    cuclark_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.k = _k; cfg.htsize = HTSIZE; cfg.key_bytes = (int)sizeof(HKMERr);
    cfg.n_targets = (int)_numTargets; cfg.row_pairs = MAXHITS; cfg.device = 0; cfg.shard_count = 1;
    B200State* s = state_of(this, true);
    b200_ok(cuclark_create(&cfg, &s->db), "cuclark_create");
    s->n_batches = _numBatches;
}

template <typename HKMERr>
CuClarkDB<HKMERr>::~CuClarkDB() {
    B200State* s = state_of(this);
    freeBatchMemory();
    cuclark_destroy(s->db);
    std::lock_guard<std::mutex> g(g_mu);
    g_state.erase(this);
    delete s;
}

// src/CuClarkDB.cu:462-808: false if a database file is missing (the caller then builds the database)
template <typename HKMERr>
bool CuClarkDB<HKMERr>::read(const char* _filename, size_t& _fileSize, size_t& _dbParts, const ITYPE& _modCollision,
                             const bool& /*_isfastLoadingRequested*/) {
    B200State* s = state_of(this);
    const int rc = cuclark_load_db_files(s->db, _filename, (int)_modCollision);
    if (rc == CUCLARK_ERR_IO) return false;
    b200_ok(rc, "cuclark_load_db_files");
    cuclark_stats st;
    cuclark_get_stats(s->db, &st);
    _fileSize = (size_t)st.table_bytes;
    _dbParts = 1;                  // the whole table is resident: no swap cycles (src/CuClarkDB.cu:543-574)
    m_dbParts = 1;
    return true;
}

template <typename HKMERr> bool CuClarkDB<HKMERr>::swapDbParts() { return false; }   // src/CuClarkDB.cu:814-858
template <typename HKMERr> bool CuClarkDB<HKMERr>::sync() { return true; }

// src/CuClarkDB.cu:318-415: pinned read buffers per batch for the caller's pack loop, flat result arrays
template <typename HKMERr>
size_t CuClarkDB<HKMERr>::malloc(size_t _numReads, size_t _maxReads, size_t _maxReadsInContainers,
                                 std::vector<ITYPE>& _indexBatches, RESULTS*& _fullResults, size_t _resultRowSize,
                                 RESULTS*& _finalResults, size_t _finalResultsRowSize, bool _isExtended,
                                 std::vector<uint32_t*>& _readsPointer, std::vector<CONTAINER*>& _readsInCon) {
    B200State* s = state_of(this);
    if (_resultRowSize != (size_t)(2 * MAXHITS + 2) || _finalResultsRowSize != 5) {
        std::cerr << "CuClarkDB_b200: unexpected result row sizes" << std::endl;
        exit(1);
    }
    const size_t nb = s->n_batches;
    b200_ok(cuclark_batches_alloc(s->db, (int)nb, _maxReads, _maxReadsInContainers, _isExtended ? 1 : 0), "cuclark_batches_alloc");
    _readsPointer.assign(nb, NULL);
    _readsInCon.assign(nb, NULL);
    s->b_final.assign(nb, NULL); s->b_rows.assign(nb, NULL);
    s->b_reads.assign(nb, 0); s->b_first.assign(nb, 0); s->b_copied.assign(nb, 0);
    for (size_t b = 0; b < nb; b++) {
        b200_ok(cuclark_batch_buffers(s->db, (int)b, &_readsPointer[b], &_readsInCon[b], &s->b_final[b], &s->b_rows[b]), "cuclark_batch_buffers");
        s->b_first[b] = _indexBatches[b];
    }
    s->row_size = _resultRowSize; s->final_row_size = _finalResultsRowSize;
    s->flat_final = (RESULTS*)calloc(_numReads * _finalResultsRowSize + 1, sizeof(RESULTS));
    s->flat_full = _isExtended ? (RESULTS*)calloc(_numReads * _resultRowSize + 1, sizeof(RESULTS)) : NULL;
    _finalResults = s->flat_final;
    if (_isExtended) _fullResults = s->flat_full;
    return (_maxReads + 1) * sizeof(uint32_t) + _maxReadsInContainers * sizeof(CONTAINER);
}

template <typename HKMERr>
void CuClarkDB<HKMERr>::freeBatchMemory() {
    B200State* s = state_of(this);
    cuclark_batches_free(s->db);
    free(s->flat_final); free(s->flat_full);
    s->flat_final = s->flat_full = NULL;
}

// called concurrently from the OpenMP threads, distinct batches (src/CuCLARK_hh.hh:1735)
template <typename HKMERr>
bool CuClarkDB<HKMERr>::readyBatch(const size_t _batchId, const size_t _numReads, const size_t _containerCount) {
    B200State* s = state_of(this);
    s->b_reads[_batchId] = _numReads;
    s->b_copied[_batchId] = 0;
    b200_ok(cuclark_batch_ready(s->db, (int)_batchId, _numReads, _containerCount), "cuclark_batch_ready");
    return true;
}

// asynchronous: H2D, kernels, D2H and the batch event are enqueued (src/CuClarkDB.cu:861-1033)
template <typename HKMERr>
bool CuClarkDB<HKMERr>::queryBatch(const size_t _batchId, const bool /*_isExtended*/, const bool /*_isFollowup*/) {
    b200_ok(cuclark_batch_query(state_of(this)->db, (int)_batchId), "cuclark_batch_query");
    return true;
}

// blocks on the batch, then lays its results into the flat arrays the orchestrator's writer reads
// (h_results[i] = _fullResults + rowSize * indexBatches[i], src/CuClarkDB.cu:410-414)
template <typename HKMERr>
bool CuClarkDB<HKMERr>::waitForBatch(size_t _batchId) {
    B200State* s = state_of(this);
    b200_ok(cuclark_batch_wait(s->db, (int)_batchId), "cuclark_batch_wait");
    std::lock_guard<std::mutex> g(s->mu);
    if (!s->b_copied[_batchId]) {
        const size_t n = s->b_reads[_batchId], first = s->b_first[_batchId];
        memcpy(s->flat_final + first * s->final_row_size, s->b_final[_batchId], n * s->final_row_size * sizeof(RESULTS));
        if (s->flat_full && s->b_rows[_batchId])
            memcpy(s->flat_full + first * s->row_size, s->b_rows[_batchId], n * s->row_size * sizeof(RESULTS));
        s->b_copied[_batchId] = 1;
    }
    return true;
}

template <typename HKMERr> bool CuClarkDB<HKMERr>::checkBatch(size_t) { return true; }

template <typename HKMERr>
bool CuClarkDB<HKMERr>::getFinalResult(const size_t _batchId, RESULTS* _finalResult) {
    B200State* s = state_of(this);
    waitForBatch(_batchId);
    memcpy(_finalResult, s->b_final[_batchId], s->b_reads[_batchId] * 5 * sizeof(RESULTS));
    return true;
}

template class CuClarkDB<uint16_t>;      // src/CuClarkDB.cu:1474-1476
template class CuClarkDB<uint32_t>;
template class CuClarkDB<uint64_t>;
