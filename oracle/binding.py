"""ctypes binding of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Imported by tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline``
/ ``--impl reference`` legs. Nothing under ``cuclark_b200/`` imports this.

``Oracle``      -> oracle/liboracle.so   (our C restatement, "port")
``RefLookup``   -> oracle/_ref/libref_lookup_{light,full}.so (the reference's
                   own hTable::find compiled from /root/reference/src)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HTSIZE_FULL = 1610612741   # src/parameters.hh:39
HTSIZE_LIGHT = 57777779    # src/parameters_light_hh:40
MAXHITS_FULL = 15          # src/parameters.hh:44
MAXHITS_LIGHT = 23         # src/parameters_light_hh:45


def key_bytes_for(k: int, htsize: int) -> int:
    """Key width the reference CLI picks (src/main.cc:278-316)."""
    import math
    t_b = int(math.log(htsize) / math.log(4.0))
    if k <= t_b + 8:
        return 2
    if k <= t_b + 16:
        return 4
    return 8


def build_port() -> str:
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "cuclark_oracle.c")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return so


class _Index(C.Structure):
    _fields_ = [("n", C.c_size_t), ("cap", C.c_size_t),
                ("name_s", C.POINTER(C.c_size_t)), ("name_e", C.POINTER(C.c_size_t)),
                ("seq_s", C.POINTER(C.c_size_t)), ("seq_e", C.POINTER(C.c_size_t)),
                ("len", C.POINTER(C.c_size_t)),
                ("n_batches", C.c_size_t), ("batch_first", C.POINTER(C.c_size_t))]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleDB:
    def __init__(self, lib, handle, htsize, k):
        self._lib, self._h, self.htsize, self.k = lib, handle, htsize, k

    @property
    def size(self) -> int:
        return self._lib.orc_db_size(self._h)

    def entries(self):
        n = self.size
        km = np.empty(n, np.uint64)
        lb = np.empty(n, np.uint16)
        self._lib.orc_db_entries(self._h, _p(km, C.c_uint64), _p(lb, C.c_uint16))
        return km, lb

    def query(self, kmers: np.ndarray, threads: int = 1):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        out = np.empty(kmers.size, np.int32)
        hits = self._lib.orc_db_query(self._h, _p(kmers, C.c_uint64), kmers.size, _p(out, C.c_int32), threads)
        return out, hits

    def close(self):
        if self._h:
            self._lib.orc_db_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ReadIndex:
    """Numpy view of an orc_index (copied out)."""

    def __init__(self, ix: _Index):
        n = ix.n
        get = lambda p: np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        self.n = n
        self.name_s, self.name_e = get(ix.name_s), get(ix.name_e)
        self.seq_s, self.seq_e, self.len = get(ix.seq_s), get(ix.seq_e), get(ix.len)
        self.batch_first = np.ctypeslib.as_array(ix.batch_first, shape=(ix.n_batches + 1,)).copy()


class Oracle:
    def __init__(self):
        self.lib = lib = C.CDLL(build_port())
        u64, u32, u16, u8, sz = C.c_uint64, C.c_uint32, C.c_uint16, C.c_uint8, C.c_size_t
        P = C.POINTER
        lib.orc_db_from_arrays.restype = C.c_void_p
        lib.orc_db_from_arrays.argtypes = [u64, C.c_int, C.c_int, P(u8), C.c_void_p, P(u16), C.c_int]
        lib.orc_db_load.restype = C.c_void_p
        lib.orc_db_load.argtypes = [C.c_char_p, u64, C.c_int, C.c_int, C.c_int]
        lib.orc_db_build_synth.restype = C.c_void_p
        lib.orc_db_build_synth.argtypes = [C.c_uint32, C.c_uint32, u64, C.c_int, u64, C.c_int, C.c_int, C.c_int,
                                           C.c_char_p]
        lib.orc_db_free.argtypes = [C.c_void_p]
        lib.orc_db_size.restype = u64
        lib.orc_db_size.argtypes = [C.c_void_p]
        lib.orc_canonical.restype = u64
        lib.orc_canonical.argtypes = [u64, C.c_int]
        lib.orc_db_query.restype = C.c_long
        lib.orc_db_query.argtypes = [C.c_void_p, P(u64), C.c_long, P(C.c_int32), C.c_int]
        lib.orc_db_entries.argtypes = [C.c_void_p, P(u64), P(u16)]
        lib.orc_index_reads.restype = C.c_int
        lib.orc_index_reads.argtypes = [P(u8), sz, sz, P(_Index)]
        lib.orc_index_free.argtypes = [P(_Index)]
        lib.orc_pack_bound.restype = sz
        lib.orc_pack_bound.argtypes = [P(_Index), sz, sz]
        lib.orc_pack.restype = sz
        lib.orc_pack.argtypes = [P(u8), P(_Index), sz, sz, C.c_int, P(u32), P(u16)]
        lib.orc_extract.restype = u64
        lib.orc_extract.argtypes = [P(u32), P(u16), sz, C.c_int, P(u64)]
        lib.orc_classify.restype = u64
        lib.orc_classify.argtypes = [C.c_void_p, P(u32), P(u16), sz, C.c_int, C.c_int, u64, u64,
                                     P(u16), P(u16), C.c_int]
        lib.orc_merge_rows.argtypes = [P(u16), P(u16), sz, C.c_int, P(u16)]
        lib.orc_result_from_rows.argtypes = [P(u16), sz, C.c_int, P(u16)]
        lib.orc_write_csv.restype = C.c_int
        lib.orc_write_csv.argtypes = [C.c_char_p, P(u8), P(_Index), C.c_int, C.c_int, P(C.c_char_p), C.c_int,
                                      P(u16), P(u16), C.c_int]
        lib.orc_merge_paired.restype = C.c_int
        lib.orc_merge_paired.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]

    # -- table ---------------------------------------------------------------
    def db_load(self, base: str, htsize: int, k: int, sfactor: int = 1) -> OracleDB:
        h = self.lib.orc_db_load(base.encode(), htsize, k, key_bytes_for(k, htsize), sfactor)
        if not h:
            raise FileNotFoundError(base)
        return OracleDB(self.lib, h, htsize, k)

    def db_from_arrays(self, htsize, k, sz, ky, lb, sfactor: int = 1) -> OracleDB:
        sz = np.ascontiguousarray(sz, np.uint8)
        lb = np.ascontiguousarray(lb, np.uint16)
        kb = ky.dtype.itemsize
        ky = np.ascontiguousarray(ky)
        h = self.lib.orc_db_from_arrays(htsize, k, kb, _p(sz, C.c_uint8), ky.ctypes.data, _p(lb, C.c_uint16), sfactor)
        return OracleDB(self.lib, h, htsize, k)

    def db_build_synth(self, seed: int, n_targets: int, genome_len: int, k: int, htsize: int, light_gap: int = 0,
                       threads: int = 1, write_base: str | None = None) -> OracleDB:
        h = self.lib.orc_db_build_synth(seed, n_targets, genome_len, k, htsize, key_bytes_for(k, htsize), light_gap,
                                        threads, write_base.encode() if write_base else None)
        return OracleDB(self.lib, h, htsize, k)

    def canonical(self, kmer: int, k: int) -> int:
        return self.lib.orc_canonical(int(kmer), k)

    # -- reads ---------------------------------------------------------------
    def index(self, data: bytes | np.ndarray, n_batches: int = 1):
        buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
        ix = _Index()
        rc = self.lib.orc_index_reads(_p(buf, C.c_uint8), buf.size, n_batches, C.byref(ix))
        if rc != 0:
            self.lib.orc_index_free(C.byref(ix))
            raise ValueError("Failed to recognize the format of the file.")
        return ix, buf

    def index_np(self, data, n_batches: int = 1) -> ReadIndex:
        ix, _ = self.index(data, n_batches)
        out = ReadIndex(ix)
        self.lib.orc_index_free(C.byref(ix))
        return out

    def pack(self, ix: _Index, buf: np.ndarray, k: int, first: int = 0, n: int | None = None):
        n = ix.n - first if n is None else n
        bound = self.lib.orc_pack_bound(C.byref(ix), first, n)
        ptr = np.zeros(n + 1, np.uint32)
        cont = np.zeros(bound, np.uint16)
        cc = self.lib.orc_pack(_p(buf, C.c_uint8), C.byref(ix), first, n, k, _p(ptr, C.c_uint32), _p(cont, C.c_uint16))
        return ptr, cont[:cc].copy()

    def extract(self, ptr, cont, k: int):
        ptr = np.ascontiguousarray(ptr, np.uint32)
        cont = np.ascontiguousarray(cont, np.uint16)
        n = ptr.size - 1
        cnt = self.lib.orc_extract(_p(ptr, C.c_uint32), _p(cont, C.c_uint16), n, k, None)
        out = np.empty(cnt, np.uint64)
        self.lib.orc_extract(_p(ptr, C.c_uint32), _p(cont, C.c_uint16), n, k, _p(out, C.c_uint64))
        return out

    def count_kmers(self, ptr, cont, k: int) -> int:
        ptr = np.ascontiguousarray(ptr, np.uint32)
        cont = np.ascontiguousarray(cont, np.uint16)
        return self.lib.orc_extract(_p(ptr, C.c_uint32), _p(cont, C.c_uint16), ptr.size - 1, k, None)

    def classify(self, db: OracleDB, ptr, cont, n_targets: int, row_pairs: int, want_rows: bool = True,
                 part=None, threads: int = 1):
        ptr = np.ascontiguousarray(ptr, np.uint32)
        cont = np.ascontiguousarray(cont, np.uint16)
        if cont.size == 0:
            cont = np.zeros(1, np.uint16)
        n = ptr.size - 1
        lo, hi = part if part is not None else (0, db.htsize)
        rows = np.zeros((n, 2 * row_pairs + 2), np.uint16) if want_rows else None
        final = np.zeros((n, 5), np.uint16)
        lookups = self.lib.orc_classify(db._h, _p(ptr, C.c_uint32), _p(cont, C.c_uint16), n, n_targets, row_pairs,
                                        lo, hi, _p(rows, C.c_uint16) if want_rows else None,
                                        _p(final, C.c_uint16), threads)
        return final, rows, lookups

    def merge_rows(self, a, b, row_pairs: int):
        out = np.zeros_like(a)
        self.lib.orc_merge_rows(_p(a, C.c_uint16), _p(b, C.c_uint16), a.shape[0], row_pairs, _p(out, C.c_uint16))
        return out

    def result_from_rows(self, rows, row_pairs: int):
        out = np.zeros((rows.shape[0], 5), np.uint16)
        self.lib.orc_result_from_rows(_p(rows, C.c_uint16), rows.shape[0], row_pairs, _p(out, C.c_uint16))
        return out

    def write_csv(self, path: str, ix: _Index, buf: np.ndarray, k: int, paired: bool, names, final, rows=None,
                  row_pairs: int = 15):
        arr = (C.c_char_p * len(names))(*[s.encode() for s in names])
        final = np.ascontiguousarray(final, np.uint16)
        rc = self.lib.orc_write_csv(path.encode(), _p(buf, C.c_uint8), C.byref(ix), k, int(paired), arr, len(names),
                                    _p(final, C.c_uint16),
                                    _p(np.ascontiguousarray(rows, np.uint16), C.c_uint16) if rows is not None else None,
                                    row_pairs)
        if rc:
            raise OSError(path)

    def free_index(self, ix: _Index):
        self.lib.orc_index_free(C.byref(ix))

    def merge_paired(self, f1: str, f2: str, out: str) -> int:
        return self.lib.orc_merge_paired(f1.encode(), f2.encode(), out.encode())


class RefLookup:
    """The reference's own host table (hTable::read + find), variant fixed at build."""

    def __init__(self, light: bool):
        so = os.path.join(HERE, "_ref", "libref_lookup_light.so" if light else "libref_lookup_full.so")
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        self.lib = lib = C.CDLL(so)
        lib.ref_htsize.restype = C.c_uint64
        lib.ref_db_open.restype = C.c_void_p
        lib.ref_db_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
        lib.ref_db_query.restype = C.c_long
        lib.ref_db_query.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_long, C.POINTER(C.c_int32), C.c_int]
        lib.ref_db_close.argtypes = [C.c_void_p]
        assert lib.ref_htsize() == (HTSIZE_LIGHT if light else HTSIZE_FULL)
        self._h = None

    def open(self, base: str, k: int, sfactor: int = 1, threads: int = 1):
        self._h = self.lib.ref_db_open(base.encode(), k, sfactor, threads)
        if not self._h:
            raise FileNotFoundError(base)
        return self

    def query(self, kmers: np.ndarray, threads: int = 1):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        out = np.empty(kmers.size, np.int32)
        hits = self.lib.ref_db_query(self._h, _p(kmers, C.c_uint64), kmers.size, _p(out, C.c_int32), threads)
        return out, hits

    def close(self):
        if self._h:
            self.lib.ref_db_close(self._h)
            self._h = None
