"""Host-side mirror of the reference's GPU driver class over the C ABI.

``CuClarkDB`` follows ``CuClarkDB<HKMERr>`` of the reference
(src/CuClarkDB.cuh:98-150): same method names, argument meaning and error
behaviour (``read`` returns False when a database file is missing; CUDA
failures raise), so the parity tests read like a test of the reference class.
Everything goes through ``libcuclark_b200.so`` (include/cuclark_b200.h) with
ctypes; there is no Python or CPU implementation behind it — if the library
is missing or there is no GPU the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcuclark_b200.so")

HTSIZE_FULL = 1610612741     # src/parameters.hh:39
HTSIZE_LIGHT = 57777779      # src/parameters_light_hh:40
MAXHITS_FULL = 15            # src/parameters.hh:44
MAXHITS_LIGHT = 23           # src/parameters_light_hh:45
FINAL_ROW = 5                # m_finalResultsRowSize, src/CuCLARK_hh.hh:1593

# every symbol include/cuclark_b200.h declares
ABI_SYMBOLS = [
    "cuclark_last_error", "cuclark_version", "cuclark_kernel_launches", "cuclark_create", "cuclark_destroy",
    "cuclark_load_db_files", "cuclark_load_db_arrays", "cuclark_build_db_synthetic",
    "cuclark_get_stats", "cuclark_sync_stats",
    "cuclark_batches_alloc", "cuclark_batch_buffers", "cuclark_batch_ready", "cuclark_batch_query",
    "cuclark_batch_wait", "cuclark_batches_free",
    "cuclark_classify_host", "cuclark_classify_device", "cuclark_merge_rows_device",
    "cuclark_synth_reads_device", "cuclark_synth_fastq_device", "cuclark_gather_bench",
    "cuclark_classify_text", "cuclark_classify_file", "cuclark_text_debug",
    "cuclark_classify_text_multi", "cuclark_classify_file_multi", "cuclark_classify_text_buffer",
    "cuclark_build_database", "cuclark_save_table", "cuclark_load_table", "cuclark_plan_table",
    "cuclark_route_alloc", "cuclark_route_free", "cuclark_route_export", "cuclark_route_import", "cuclark_route_connect",
    "cuclark_route_scatter", "cuclark_route_probe", "cuclark_route_gather", "cuclark_route_get_stats",
    "cuclark_classify_routed_device", "cuclark_clone_table", "cuclark_device_info", "cuclark_synth_fastq_pair_device",
]


class CuclarkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cuclark_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("k", C.c_int), ("htsize", C.c_uint64), ("key_bytes", C.c_int), ("n_targets", C.c_int),
                ("row_pairs", C.c_int), ("device", C.c_int), ("shard_index", C.c_int), ("shard_count", C.c_int),
                ("bucket_load", C.c_double), ("layout", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("n_entries", C.c_uint64), ("n_buckets", C.c_uint64), ("n_local_buckets", C.c_uint64),
                ("table_bytes", C.c_uint64), ("n_spilled", C.c_uint64), ("n_spill_buckets", C.c_uint64),
                ("layout", C.c_int), ("k", C.c_int), ("lookups", C.c_uint64), ("dense_reads", C.c_uint64),
                ("truncated_rows", C.c_uint64), ("last_kernel_ms", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class TextOpts(C.Structure):
    _fields_ = [("paired", C.c_int), ("extended", C.c_int), ("target_names", C.POINTER(C.c_char_p)),
                ("chunk_bytes", C.c_size_t), ("n_slots", C.c_int)]


class TextStats(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("lookups", C.c_uint64), ("n_containers", C.c_uint64),
                ("text_bytes", C.c_uint64), ("csv_bytes", C.c_uint64), ("n_chunks", C.c_uint64),
                ("dense_reads", C.c_uint64), ("truncated_rows", C.c_uint64), ("seconds", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class TextArrays(C.Structure):
    _fields_ = [("cap_reads", C.c_size_t), ("cap_containers", C.c_size_t),
                ("name_s", C.c_void_p), ("name_e", C.c_void_p), ("seq_s", C.c_void_p), ("seq_e", C.c_void_p),
                ("len", C.c_void_p), ("reads_ptr", C.c_void_p), ("containers", C.c_void_p),
                ("final5", C.c_void_p), ("rows", C.c_void_p)]


class BuildOpts(C.Structure):
    _fields_ = [("k", C.c_int), ("htsize", C.c_uint64), ("key_bytes", C.c_int), ("light_gap", C.c_int),
                ("min_count", C.c_uint32), ("device", C.c_int)]


class BuildStats(C.Structure):
    _fields_ = [("n_nucleotides", C.c_uint64), ("n_kmers_added", C.c_uint64), ("n_kmers_kept", C.c_uint64),
                ("key_bytes", C.c_int)]


class RouteStats(C.Structure):
    _fields_ = [("n_ranks", C.c_int), ("rank", C.c_int), ("region_bytes", C.c_uint64), ("map_bytes", C.c_uint64),
                ("cap_blocks", C.c_uint64), ("lookups", C.c_uint64), ("probed", C.c_uint64), ("blocks", C.c_uint64),
                ("blocks_remote", C.c_uint64), ("err", C.c_uint32)]


SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64)

_lib = None


def load_library():
    """dlopen libcuclark_b200.so and declare the prototypes. Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    # CUCLARK_LIB: another build of the same library (kernel variants under lib/variants/, tools/kernel_variants.py)
    path = os.environ.get("CUCLARK_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: build it with `python -m cuclark_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(path)
    u64, u32, u16, u8, sz, vp, ci = C.c_uint64, C.c_uint32, C.c_uint16, C.c_uint8, C.c_size_t, C.c_void_p, C.c_int
    P = C.POINTER
    lib.cuclark_last_error.restype = C.c_char_p
    lib.cuclark_version.restype = ci
    lib.cuclark_create.argtypes = [P(Config), P(vp)]
    lib.cuclark_destroy.argtypes = [vp]
    lib.cuclark_load_db_files.argtypes = [vp, C.c_char_p, ci]
    lib.cuclark_load_db_arrays.argtypes = [vp, vp, vp, vp, u64, ci]
    lib.cuclark_build_db_synthetic.argtypes = [vp, u32, u32, u64, ci]
    lib.cuclark_get_stats.argtypes = [vp, P(Stats)]
    lib.cuclark_sync_stats.argtypes = [vp, vp]
    lib.cuclark_batches_alloc.argtypes = [vp, ci, sz, sz, ci]
    lib.cuclark_batch_buffers.argtypes = [vp, ci, P(vp), P(vp), P(vp), P(vp)]
    lib.cuclark_batch_ready.argtypes = [vp, ci, sz, sz]
    lib.cuclark_batch_query.argtypes = [vp, ci]
    lib.cuclark_batch_wait.argtypes = [vp, ci]
    lib.cuclark_batches_free.argtypes = [vp]
    lib.cuclark_classify_host.argtypes = [vp, vp, vp, sz, vp, vp]
    lib.cuclark_classify_device.argtypes = [vp, vp, vp, sz, vp, vp, vp]
    lib.cuclark_merge_rows_device.argtypes = [vp, vp, ci, sz, vp, vp, vp]
    lib.cuclark_synth_reads_device.argtypes = [vp, u32, u32, u32, u64, u64, sz, ci, ci, ci, vp, vp, vp]
    lib.cuclark_synth_fastq_device.argtypes = [vp, u32, u32, u32, u64, u64, sz, ci, ci, ci, vp, vp]
    lib.cuclark_gather_bench.argtypes = [vp, u64, ci, ci, ci, P(C.c_double)]
    lib.cuclark_classify_text.argtypes = [vp, vp, sz, P(TextOpts), SINK_FN, vp, P(TextStats)]
    lib.cuclark_classify_file.argtypes = [vp, C.c_char_p, C.c_char_p, P(TextOpts), P(TextStats)]
    lib.cuclark_text_debug.argtypes = [vp, vp, sz, P(TextOpts), P(TextArrays), P(TextStats)]
    lib.cuclark_classify_text_multi.argtypes = [P(vp), ci, vp, sz, P(TextOpts), SINK_FN, vp, P(TextStats)]
    lib.cuclark_classify_text_buffer.argtypes = [P(vp), ci, vp, sz, P(TextOpts), vp, sz, P(sz), P(TextStats)]
    lib.cuclark_build_database.argtypes = [P(BuildOpts), P(C.c_char_p), P(u16), sz, C.c_char_p, P(BuildStats)]
    lib.cuclark_save_table.argtypes = [vp, C.c_char_p]
    lib.cuclark_load_table.argtypes = [vp, C.c_char_p, C.c_char_p, ci]
    lib.cuclark_classify_file_multi.argtypes = [P(vp), ci, C.c_char_p, C.c_char_p, P(TextOpts), P(TextStats)]
    if not hasattr(lib, "cuclark_route_alloc"):         # (variant build from before the routing entry points)
        _lib = lib
        for name in ABI_SYMBOLS:
            fn = getattr(lib, name, None)
            if fn is not None and name != "cuclark_last_error":
                fn.restype = ci
        return lib
    lib.cuclark_synth_fastq_pair_device.argtypes = [vp, u32, u32, u32, u64, u64, sz, ci, ci, ci, ci, vp, vp]
    lib.cuclark_clone_table.argtypes = [vp, vp]
    lib.cuclark_device_info.argtypes = [ci, P(ci), P(u64), P(u64)]
    lib.cuclark_route_alloc.argtypes = [vp, ci, sz]
    lib.cuclark_route_free.argtypes = [vp]
    lib.cuclark_route_export.argtypes = [vp, vp, P(u64)]
    lib.cuclark_route_import.argtypes = [vp, ci, vp]
    lib.cuclark_route_connect.argtypes = [P(vp), ci]
    lib.cuclark_route_scatter.argtypes = [vp, vp, vp, sz, sz, vp]
    lib.cuclark_route_probe.argtypes = [vp, vp]
    lib.cuclark_route_gather.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp]
    lib.cuclark_route_get_stats.argtypes = [vp, P(RouteStats)]
    lib.cuclark_classify_routed_device.argtypes = [P(vp), ci, P(vp), P(vp), P(sz), P(sz), P(vp), P(vp)]
    variant = bool(os.environ.get("CUCLARK_LIB"))       # an older build for A/B timing may lack the newest entry points
    for name in ABI_SYMBOLS:
        fn = getattr(lib, name, None)
        if fn is None:
            if variant:
                continue
            raise AttributeError(f"{name} is declared in include/cuclark_b200.h but not exported by {path}")
        if name not in ("cuclark_last_error", "cuclark_kernel_launches"):
            fn.restype = ci
    if hasattr(lib, "cuclark_kernel_launches"):
        lib.cuclark_kernel_launches.restype = u64
    _lib = lib
    return lib


class TablePlan(C.Structure):
    _fields_ = [("layout", C.c_int), ("n_buckets", C.c_uint64), ("n_local_buckets", C.c_uint64), ("home_bytes", C.c_uint64)]


def kernel_launches() -> int:
    """Hot-path kernels launched by the library in this process so far (cuclark_kernel_launches)."""
    lib = load_library()
    return int(lib.cuclark_kernel_launches()) if hasattr(lib, "cuclark_kernel_launches") else 0


def plan_table(k: int, n_entries: int, htsize: int = HTSIZE_FULL, n_targets: int = 1, shard=(0, 1), bucket_load: float = 0.0,
               layout: int = 0) -> dict:
    """cuclark_plan_table: the layout and size the loader would choose, no device needed."""
    lib = load_library()
    cfg = Config(k, htsize, 0, n_targets, 0, 0, shard[0], shard[1], bucket_load, layout)
    out = TablePlan()
    rc = lib.cuclark_plan_table(C.byref(cfg), C.c_uint64(n_entries), C.byref(out))
    if rc != 0:
        raise CuclarkError(rc, lib.cuclark_last_error().decode())
    return {f: getattr(out, f) for f, _ in TablePlan._fields_}


def key_bytes_for(k: int, htsize: int) -> int:
    """Width of a .ky element as the reference CLI picks it (src/main.cc:278-316)."""
    import math
    t_b = int(math.log(htsize) / math.log(4.0))
    return 2 if k <= t_b + 8 else 4 if k <= t_b + 16 else 8


def _np_from(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype)
    ct = {np.uint32: C.c_uint32, np.uint16: C.c_uint16}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))


def build_database(target_files, target_labels, out_base: str, k: int, light: bool = False, light_gap: int = 0,
                   min_count: int = 0, device: int = 0, htsize: int | None = None) -> dict:
    """makeSpecificTargetSets + RemoveCommon + write on the device: FASTA targets -> <out_base>.sz/.ky/.lb.
    `target_labels[i]` = label index of file i (order of first appearance in the targets file)."""
    lib = load_library()
    hts = htsize if htsize is not None else (HTSIZE_LIGHT if light else HTSIZE_FULL)
    o = BuildOpts(k, hts, 0, (light_gap or 4) if light else 0, min_count, device)
    files = (C.c_char_p * len(target_files))(*[f.encode() for f in target_files])
    labels = (C.c_uint16 * len(target_labels))(*target_labels)
    st = BuildStats()
    rc = lib.cuclark_build_database(C.byref(o), files, labels, len(target_files), out_base.encode(), C.byref(st))
    if rc != 0:
        raise CuclarkError(rc, lib.cuclark_last_error().decode())
    return {f: getattr(st, f) for f, _ in st._fields_}


class CuClarkDB:
    """Mirror of ``CuClarkDB<HKMERr>`` (src/CuClarkDB.cuh:98-150) for ONE device.

    ctor(numDevices, k, numBatches, numTargets) becomes
    ``CuClarkDB(k, n_targets, light=..., device=..., shard=(i, n))``: the
    variant is a run-time choice and multi-GPU is one process per GPU.
    """

    def __init__(self, k: int, n_targets: int, light: bool = False, device: int = 0, shard=(0, 1),
                 row_pairs: int = 0, bucket_load: float = 0.0, layout: int = 0, htsize: int | None = None):
        self._lib = load_library()
        self.htsize = htsize if htsize is not None else (HTSIZE_LIGHT if light else HTSIZE_FULL)
        self.k = k
        self.n_targets = n_targets
        self.row_pairs = row_pairs or (MAXHITS_LIGHT if self.htsize == HTSIZE_LIGHT else MAXHITS_FULL)
        self.row_size = 2 * self.row_pairs + 2          # m_resultRowSize, src/CuCLARK_hh.hh:1586-1590
        self.key_bytes = key_bytes_for(k, self.htsize)
        cfg = Config(k, self.htsize, self.key_bytes, n_targets, self.row_pairs, device, shard[0], shard[1],
                     bucket_load, layout)
        h = C.c_void_p()
        self._h = None
        self._check(self._lib.cuclark_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._batches = 0
        self._want_rows = False

    # -- plumbing -------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise CuclarkError(rc, self._lib.cuclark_last_error().decode())

    def close(self):
        if self._h:
            self._lib.cuclark_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- database (CuClarkDB::read, src/CuClarkDB.cu:462) ------------------------
    def read(self, filename: str, mod_collision: int = 1) -> bool:
        """Load ``<filename>.sz/.ky/.lb``; False if a file is missing (the
        reference's contract, the caller then rebuilds or exits)."""
        rc = self._lib.cuclark_load_db_files(self._h, filename.encode(), mod_collision)
        if rc == -4:      # CUCLARK_ERR_IO
            return False
        self._check(rc)
        return True

    def load_arrays(self, sz, ky, lb, mod_collision: int = 1):
        sz = np.ascontiguousarray(sz, np.uint8)
        lb = np.ascontiguousarray(lb, np.uint16)
        ky = np.ascontiguousarray(ky)
        assert ky.dtype.itemsize == self.key_bytes, "key dtype must match the reference's key width for this k"
        assert sz.size == self.htsize
        self._check(self._lib.cuclark_load_db_arrays(self._h, sz.ctypes.data, ky.ctypes.data, lb.ctypes.data,
                                                      ky.size, mod_collision))

    def save_table(self, path: str):
        """Write the loaded table in its device layout (cuclark_save_table)."""
        self._check(self._lib.cuclark_save_table(self._h, path.encode()))

    def load_table(self, path: str, src_base: str | None = None, mod_collision: int = 1) -> bool:
        """Stream a table cache back into HBM. False if the file is missing, is no cache, is corrupt
        or belongs to another configuration / other source files (the caller then calls read())."""
        rc = self._lib.cuclark_load_table(self._h, path.encode(), src_base.encode() if src_base else None,
                                          mod_collision)
        if rc in (-4, -8):      # CUCLARK_ERR_IO, CUCLARK_ERR_FORMAT
            return False
        self._check(rc)
        return True

    def clone_table_from(self, src: "CuClarkDB"):
        """Take a device-to-device copy of `src`'s loaded table (cuclark_clone_table)."""
        self._check(self._lib.cuclark_clone_table(src._h, self._h))

    def build_synthetic(self, seed: int, n_targets: int, genome_len: int, light_gap: int = 0):
        self._check(self._lib.cuclark_build_db_synthetic(self._h, seed, n_targets, genome_len, light_gap))

    def swapDbParts(self) -> bool:
        """The table is fully resident: there is never another part (src/CuClarkDB.cu:814-858)."""
        return False

    def sync(self) -> bool:
        return True

    def stats(self, sync_stream=None, sync: bool = False) -> dict:
        if sync:
            self._check(self._lib.cuclark_sync_stats(self._h, sync_stream))
        s = Stats()
        self._check(self._lib.cuclark_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    # -- batches (malloc/readyBatch/queryBatch/waitForBatch/freeBatchMemory) ------
    def malloc(self, n_batches: int, max_reads: int, max_containers: int, is_extended: bool = False):
        """Allocate pinned + device buffers; returns per-batch numpy views
        (reads_ptr, containers, final, rows) over the library's pinned memory."""
        self._check(self._lib.cuclark_batches_alloc(self._h, n_batches, max_reads, max_containers, int(is_extended)))
        self._batches, self._want_rows = n_batches, is_extended
        views = []
        for b in range(n_batches):
            p, c, f, r = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
            self._check(self._lib.cuclark_batch_buffers(self._h, b, C.byref(p), C.byref(c), C.byref(f), C.byref(r)))
            views.append((_np_from(p, max_reads + 1, np.uint32), _np_from(c, max_containers, np.uint16),
                          _np_from(f, max_reads * FINAL_ROW, np.uint16).reshape(max_reads, FINAL_ROW),
                          _np_from(r, max_reads * self.row_size, np.uint16).reshape(max_reads, self.row_size)
                          if is_extended else None))
        return views

    def readyBatch(self, batch_id: int, n_reads: int, container_count: int) -> bool:
        self._check(self._lib.cuclark_batch_ready(self._h, batch_id, n_reads, container_count))
        return True

    def queryBatch(self, batch_id: int, is_extended: bool = False, is_followup: bool = False) -> bool:
        self._check(self._lib.cuclark_batch_query(self._h, batch_id))
        return True

    def waitForBatch(self, batch_id: int) -> bool:
        self._check(self._lib.cuclark_batch_wait(self._h, batch_id))
        return True

    def freeBatchMemory(self):
        self._check(self._lib.cuclark_batches_free(self._h))
        self._batches = 0

    # -- one-shot ------------------------------------------------------------------
    def classify(self, reads_ptr, containers, want_rows: bool = False):
        """Host arrays in (reference packed format), host arrays out."""
        ptr = np.ascontiguousarray(reads_ptr, np.uint32)
        cont = np.ascontiguousarray(containers, np.uint16)
        n = ptr.size - 1
        final = np.zeros((n, FINAL_ROW), np.uint16)
        rows = np.zeros((n, self.row_size), np.uint16) if want_rows else None
        self._check(self._lib.cuclark_classify_host(self._h, ptr.ctypes.data, cont.ctypes.data if cont.size else None, n,
                                                     final.ctypes.data, rows.ctypes.data if want_rows else None))
        return final, rows

    def classify_device(self, d_ptr: int, d_cont: int, n_reads: int, d_final: int = 0, d_rows: int = 0, stream: int = 0):
        self._check(self._lib.cuclark_classify_device(self._h, d_ptr, d_cont, n_reads, d_final or None, d_rows or None,
                                                       stream or None))

    def merge_rows_device(self, d_parts: int, n_parts: int, n_reads: int, d_rows_out: int = 0, d_final: int = 0,
                          stream: int = 0):
        self._check(self._lib.cuclark_merge_rows_device(self._h, d_parts, n_parts, n_reads, d_rows_out or None,
                                                         d_final or None, stream or None))

    # -- table-partitioned mode by k-mer routing (include/cuclark_b200.h, csrc/route.cu) ------------------------
    def route_alloc(self, n_ranks: int, max_containers: int):
        self._check(self._lib.cuclark_route_alloc(self._h, n_ranks, max_containers))

    def route_free(self):
        self._check(self._lib.cuclark_route_free(self._h))

    def route_export(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's region (one process per GPU: send it to the other ranks)."""
        buf = C.create_string_buffer(64)
        self._check(self._lib.cuclark_route_export(self._h, buf, None))
        return buf.raw

    def route_import(self, peer_rank: int, handle: bytes):
        buf = C.create_string_buffer(handle, 64)
        self._check(self._lib.cuclark_route_import(self._h, peer_rank, buf))

    def route_scatter(self, d_ptr: int, d_cont: int, n_reads: int, n_cont: int, stream: int = 0):
        self._check(self._lib.cuclark_route_scatter(self._h, d_ptr, d_cont, n_reads, n_cont, stream or None))

    def route_probe(self, stream: int = 0):
        self._check(self._lib.cuclark_route_probe(self._h, stream or None))

    def route_gather(self, d_ptr: int, d_cont: int, n_reads: int, n_cont: int, d_final: int = 0, d_rows: int = 0,
                     stream: int = 0):
        self._check(self._lib.cuclark_route_gather(self._h, d_ptr, d_cont, n_reads, n_cont, d_final or None,
                                                    d_rows or None, stream or None))

    def route_stats(self) -> dict:
        s = RouteStats()
        self._check(self._lib.cuclark_route_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in RouteStats._fields_}

    def synth_reads_device(self, seed, genome_seed, n_targets, genome_len, first_read, n_reads, read_len,
                           pct_random, sub_per_10k, d_ptr: int, d_cont: int, stream: int = 0):
        self._check(self._lib.cuclark_synth_reads_device(self._h, seed, genome_seed, n_targets, genome_len, first_read,
                                                          n_reads, read_len, pct_random, sub_per_10k, d_ptr, d_cont,
                                                          stream or None))

    def synth_fastq_device(self, seed, genome_seed, n_targets, genome_len, first_read, n_reads, read_len,
                           pct_random, sub_per_10k, d_text: int, stream: int = 0):
        self._check(self._lib.cuclark_synth_fastq_device(self._h, seed, genome_seed, n_targets, genome_len, first_read,
                                                          n_reads, read_len, pct_random, sub_per_10k, d_text,
                                                          stream or None))

    def synth_fastq_pair_device(self, seed, genome_seed, n_targets, genome_len, first_read, n_reads, read_len,
                                pct_random, sub_per_10k, mate: int, d_text: int, stream: int = 0):
        self._check(self._lib.cuclark_synth_fastq_pair_device(self._h, seed, genome_seed, n_targets, genome_len, first_read,
                                                               n_reads, read_len, pct_random, sub_per_10k, mate, d_text,
                                                               stream or None))

    def gather_bench(self, n_probes: int, bytes_per_probe: int = 32, ilp: int = 4, iters: int = 5) -> float:
        ms = C.c_double()
        self._check(self._lib.cuclark_gather_bench(self._h, n_probes, bytes_per_probe, ilp, iters, C.byref(ms)))
        return ms.value

    # -- text in, CSV out (CuCLARK::runSimple -> getObjectsDataComputeFullGPU -> printExtendedResultsSynced) ----
    def _text_opts(self, names, paired, extended, chunk_bytes, n_slots):
        o = TextOpts()
        o.paired, o.extended, o.chunk_bytes, o.n_slots = int(paired), int(extended), chunk_bytes, n_slots
        if names is not None:
            assert len(names) == self.n_targets
            arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
            o.target_names = arr
            o._keep = arr
        return o

    def classify_text(self, data, names=None, paired=False, extended=False, chunk_bytes=0, n_slots=0):
        """Raw FASTA/FASTQ bytes (anything with the buffer protocol) -> (CSV bytes, stats), through the
        sink callback (pieces arrive with their byte offset, possibly out of order)."""
        buf = np.frombuffer(data, np.uint8)
        o = self._text_opts(names, paired, extended, chunk_bytes, n_slots)
        out = bytearray()

        def sink(_user, ptr, m, off):
            if len(out) < off + m:
                out.extend(bytes(off + m - len(out)))
            out[off:off + m] = C.string_at(ptr, m)
            return 0

        cb = SINK_FN(sink)
        st = TextStats()
        self._check(self._lib.cuclark_classify_text(self._h, buf.ctypes.data, buf.size, C.byref(o), cb, None, C.byref(st)))
        assert len(out) == st.csv_bytes
        return bytes(out), st.as_dict()

    def classify_text_buffer(self, text_addr: int, n: int, out_addr: int, out_cap: int, names=None, paired=False,
                             extended=False, chunk_bytes=0, n_slots=0, peers=()):
        """Host addresses in and out (pinned memory is copied from/to directly). `peers`: more handles
        (other devices, same database) for a read-partitioned multi-GPU run. Returns (csv_len, stats)."""
        o = self._text_opts(names, paired, extended, chunk_bytes, n_slots)
        hs = [self._h] + [p._h for p in peers]
        arr = (C.c_void_p * len(hs))(*hs)
        st = TextStats()
        ln = C.c_size_t()
        self._check(self._lib.cuclark_classify_text_buffer(arr, len(hs), text_addr, n, C.byref(o), out_addr, out_cap,
                                                            C.byref(ln), C.byref(st)))
        return ln.value, st.as_dict()

    def classify_file(self, objects_path: str, csv_path: str, names=None, paired=False, extended=False,
                      chunk_bytes=0, n_slots=0) -> dict:
        o = self._text_opts(names, paired, extended, chunk_bytes, n_slots)
        st = TextStats()
        self._check(self._lib.cuclark_classify_file(self._h, objects_path.encode(), csv_path.encode(), C.byref(o),
                                                     C.byref(st)))
        return st.as_dict()

    def text_debug(self, data: bytes, max_reads: int, max_containers: int, chunk_bytes=0, n_slots=0,
                   classify=True, want_rows=False):
        """Device-side index + pack (+ classify) of raw text; returns the intermediate arrays."""
        buf = np.frombuffer(data, np.uint8)
        a = TextArrays()
        a.cap_reads, a.cap_containers = max_reads, max_containers
        arrs = {k: np.zeros(max_reads, np.uint64) for k in ("name_s", "name_e", "seq_s", "seq_e", "len")}
        arrs["reads_ptr"] = np.zeros(max_reads + 1, np.uint32)
        arrs["containers"] = np.zeros(max_containers, np.uint16)
        if classify:
            arrs["final5"] = np.zeros((max_reads, FINAL_ROW), np.uint16)
            if want_rows:
                arrs["rows"] = np.zeros((max_reads, self.row_size), np.uint16)
        for k, v in arrs.items():
            setattr(a, k, v.ctypes.data)
        o = self._text_opts(None, False, False, chunk_bytes, n_slots)
        st = TextStats()
        self._check(self._lib.cuclark_text_debug(self._h, buf.ctypes.data, buf.size, C.byref(o), C.byref(a), C.byref(st)))
        n, nc = st.n_reads, st.n_containers
        out = {k: (v[:n + 1] if k == "reads_ptr" else v[:nc] if k == "containers" else v[:n]) for k, v in arrs.items()}
        return out, st.as_dict()


def device_info(device: int = 0) -> dict:
    lib = load_library()
    n, f, t = C.c_int(), C.c_uint64(), C.c_uint64()
    rc = lib.cuclark_device_info(device, C.byref(n), C.byref(f), C.byref(t))
    if rc != 0:
        raise CuclarkError(rc, lib.cuclark_last_error().decode())
    return {"n_devices": n.value, "free_bytes": f.value, "total_bytes": t.value}


def route_connect(handles):
    """N handles of THIS process, in rank order (cuclark_route_connect): peer access + plain pointers."""
    lib = load_library()
    arr = (C.c_void_p * len(handles))(*[h._h for h in handles])
    rc = lib.cuclark_route_connect(arr, len(handles))
    if rc != 0:
        raise CuclarkError(rc, lib.cuclark_last_error().decode())


def classify_routed_device(handles, d_ptrs, d_conts, n_reads, n_conts, d_finals=None, d_rows=None):
    """cuclark_classify_routed_device: scatter | barrier | probe | barrier | gather for the connected handles of
    this process; per rank its device pointers (ints) and counts. Synchronises every rank's library stream."""
    lib = load_library()
    n = len(handles)
    vp = C.c_void_p * n
    szs = C.c_size_t * n
    rc = lib.cuclark_classify_routed_device(
        vp(*[h._h for h in handles]), n, vp(*d_ptrs), vp(*d_conts), szs(*n_reads), szs(*n_conts),
        vp(*[x or None for x in d_finals]) if d_finals else None, vp(*[x or None for x in d_rows]) if d_rows else None)
    if rc != 0:
        raise CuclarkError(rc, lib.cuclark_last_error().decode())
