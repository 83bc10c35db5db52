"""Python binding of libgblastn_b200.so (the C ABI of include/gblastn_b200.h).

The library is the product: this module only marshals numpy arrays into the POD structs.
There is no fallback — if the shared library is missing, or no CUDA device is usable, calls
raise (the product path must fail loudly, never route through the oracle).
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GBLASTN_B200_LIB") or os.path.join(_HERE, "libgblastn_b200.so")      # override: kernel experiments

EXPORTS = [
    "bn_init", "bn_release", "bn_device_count", "bn_last_error", "bn_version",
    "bn_db_load", "bn_db_free", "bn_query_load", "bn_query_free",
    "bn_prelim_search", "bn_prelim_search_host", "bn_prelim_search_volumes", "bn_results_free",
    "bn_scan_subject", "bn_word_finder", "bn_free", "bn_bench_scan", "bn_query_download_lookup", "bn_get_gapped_score", "bn_gapped_traceback", "bn_traceback_hsps", "bn_traceback_search",
    "bn_dbfile_index", "bn_db_load_files", "bn_dbfile_write", "bn_dbfile_ambiguity", "bn_db_set_ambiguity", "bn_prelim_search_batches", "bn_prelim_search_jobs", "bn_db_set_masks", "bn_selftest_replay", "bn_selftest_sort",
    "bn_setup_create", "bn_setup_batch", "bn_setup_kbp_std", "bn_setup_kbp_gap",
    "bn_setup_gap_x_dropoff_final", "bn_setup_longest_chain", "bn_setup_free", "bn_dust_mask", "bn_dust_mask_batch",
]


class BnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gblastn_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BnError(-1, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH)
        _lib.bn_last_error.restype = C.c_char_p
        _lib.bn_version.restype = C.c_char_p
        _lib.bn_setup_batch.restype = C.POINTER(abi.BnQueryBatch)
        _lib.bn_setup_kbp_std.restype = C.POINTER(C.c_double)
        _lib.bn_setup_kbp_gap.restype = C.POINTER(C.c_double)
    return _lib


def _check(rc):
    if rc != 0:
        raise BnError(rc, lib().bn_last_error().decode())


def init(n_gpu=0, device_ids=None):
    ids = None
    if device_ids is not None:
        ids = (C.c_int * len(device_ids))(*device_ids)
        n_gpu = len(device_ids)
    _check(lib().bn_init(C.c_int(n_gpu), ids))


def release():
    lib().bn_release()


def device_count() -> int:
    return int(lib().bn_device_count())


class Volume:
    """A database volume resident in one GPU's HBM."""

    def __init__(self, vol, device=0):
        self.packed = np.ascontiguousarray(vol.packed, dtype=np.uint8)
        self.byte_off = np.ascontiguousarray(vol.byte_off, dtype=np.int64)
        self.seq_len = np.ascontiguousarray(vol.seq_len, dtype=np.int32)
        h = C.c_int(-1)
        _check(lib().bn_db_load(C.c_int(device), self.packed.ctypes.data_as(C.c_void_p),
                                C.c_int64(self.packed.shape[0]),
                                self.byte_off.ctypes.data_as(C.c_void_p),
                                self.seq_len.ctypes.data_as(C.c_void_p),
                                C.c_int32(self.seq_len.shape[0]), C.byref(h)))
        self.handle = h.value
        self.device = device

    def free(self):
        if self.handle >= 0:
            lib().bn_db_free(C.c_int(self.handle))
            self.handle = -1

    def set_ambiguity(self, first, runs):
        """Ambiguity runs (first: int64[n_seq + 1]; runs: int32[n, 3] = first base, bases, blastna code); None removes."""
        if first is None:
            _check(lib().bn_db_set_ambiguity(C.c_int(self.handle), None, None))
            return
        f = np.ascontiguousarray(first, dtype=np.int64)
        r = np.ascontiguousarray(runs, dtype=np.int32).reshape(-1)
        if r.size == 0:
            r = np.zeros(3, np.int32)
        _check(lib().bn_db_set_ambiguity(C.c_int(self.handle), f.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p)))

    def set_masks(self, masks, mask_type=abi.BN_MASK_SOFT):
        """Database masks: `masks` = per sequence a list of half-open (begin, end) masked intervals
        (ascending, disjoint); mask_type BN_MASK_SOFT / BN_MASK_HARD; None removes them."""
        if masks is None:
            _check(lib().bn_db_set_masks(C.c_int(self.handle), C.c_int(abi.BN_MASK_NONE), None, None))
            return
        mn = np.ascontiguousarray([len(m) for m in masks], dtype=np.int32)
        flat = [x for m in masks for iv in m for x in iv]
        miv = np.ascontiguousarray(flat if flat else [0, 0], dtype=np.int32)
        _check(lib().bn_db_set_masks(C.c_int(self.handle), C.c_int(mask_type), mn.ctypes.data_as(C.c_void_p),
                                     miv.ctypes.data_as(C.c_void_p)))


class FileVolume(Volume):
    """A volume loaded from BLAST database files (.nin + .nsq) straight into HBM."""

    def __init__(self, nin_path, nsq_path, device=0):
        h = C.c_int(-1)
        _check(lib().bn_db_load_files(C.c_int(device), str(nin_path).encode(), str(nsq_path).encode(), C.byref(h)))
        self.handle = h.value
        self.device = device


def dbfile_index(nin_path, nsq_path):
    """Host-only parse of a volume's index: (info dict, byte offsets into the .nsq, lengths)."""
    info = abi.BnDbFileInfo()
    _check(lib().bn_dbfile_index(str(nin_path).encode(), str(nsq_path).encode(), C.byref(info), None, None))
    off = np.zeros(info.n_seq, dtype=np.int64)
    ln = np.zeros(info.n_seq, dtype=np.int32)
    _check(lib().bn_dbfile_index(str(nin_path).encode(), str(nsq_path).encode(), C.byref(info),
                                 off.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p)))
    d = {"n_seq": info.n_seq, "max_len": info.max_len, "total_bases": info.total_bases,
         "nsq_bytes": info.nsq_bytes, "title": info.title.decode(errors="replace")}
    return d, off, ln


def dbfile_ambiguity(nin_path, nsq_path):
    """Ambiguity runs of a volume: (first: int64[n_seq + 1], runs: int32[n, 3] = first base, bases, blastna code)."""
    info, _, _ = dbfile_index(nin_path, nsq_path)
    first = np.zeros(info["n_seq"] + 1, dtype=np.int64)
    p, n = C.POINTER(C.c_int32)(), C.c_int64(0)
    _check(lib().bn_dbfile_ambiguity(str(nin_path).encode(), str(nsq_path).encode(), first.ctypes.data_as(C.c_void_p),
                                     C.byref(p), C.byref(n)))
    try:
        runs = np.ctypeslib.as_array(p, shape=(max(n.value, 1) * 3,)).copy()[: n.value * 3].reshape(-1, 3)
    finally:
        lib().bn_free(p)
    return first, runs


def dbfile_write(nin_path, nsq_path, vol, title="synthetic"):
    packed = np.ascontiguousarray(vol.packed, dtype=np.uint8)
    boff = np.ascontiguousarray(vol.byte_off, dtype=np.int64)
    slen = np.ascontiguousarray(vol.seq_len, dtype=np.int32)
    _check(lib().bn_dbfile_write(str(nin_path).encode(), str(nsq_path).encode(), title.encode(),
                                 packed.ctypes.data_as(C.c_void_p), boff.ctypes.data_as(C.c_void_p),
                                 slen.ctypes.data_as(C.c_void_p), C.c_int32(slen.shape[0])))


class Query:
    """A query batch (lookup table + parameters) loaded for every device in use."""

    def __init__(self, holder):
        self.holder = holder       # keeps the numpy arrays alive
        batch = holder.batch if hasattr(holder, "batch") else holder
        h = C.c_int(-1)
        _check(lib().bn_query_load(C.byref(batch), C.byref(h)))
        self.handle = h.value

    def free(self):
        if self.handle >= 0:
            lib().bn_query_free(C.c_int(self.handle))
            self.handle = -1


_STAT_FIELDS = tuple(k for k, _ in abi.BnStats._fields_)


def _results(res: abi.BnResults) -> dict:
    try:
        return {
            "hsps": abi.struct_array(res.hsps, res.n_hsps, abi.HSP_DTYPE),
            "init": abi.struct_array(res.init, res.n_init, abi.INIT_DTYPE),
            "gapped": abi.struct_array(res.gapped, res.n_gapped, abi.HSP_DTYPE),
            "stats": {k: getattr(res.stats, k) for k in _STAT_FIELDS},
        }
    finally:
        lib().bn_results_free(C.byref(res))


def prelim_search(volume: Volume, query: Query, oid_begin=0, oid_end=-1, taps=0) -> dict:
    res = abi.BnResults()
    _check(lib().bn_prelim_search(C.c_int(volume.handle), C.c_int(query.handle),
                                  C.c_int32(oid_begin), C.c_int32(oid_end), C.c_int(taps),
                                  C.byref(res)))
    return _results(res)


def prelim_search_volumes(volumes, query: Query, taps=0, prune_hitlists=False) -> dict:
    """One query batch against several resident volumes (any devices); OIDs are those of the concatenated database
    and the hit-list rules (low_score, prelim_hitlist_size) are applied once, across the volumes."""
    n = len(volumes)
    hs = (C.c_int * n)(*[v.handle for v in volumes])
    res = abi.BnResults()
    _check(lib().bn_prelim_search_volumes(C.c_int32(n), hs, C.c_int(query.handle), C.c_int(taps),
                                          C.c_int(1 if prune_hitlists else 0), C.byref(res)))
    return _results(res)


def prelim_search_batches(volume: Volume, holders, taps=0) -> list:
    """Pipelined search of several query batches against one resident volume (one result dict per batch)."""
    n = len(holders)
    batches = [h.batch if hasattr(h, "batch") else h for h in holders]
    ptrs = (C.POINTER(abi.BnQueryBatch) * n)(*[C.pointer(b) for b in batches])
    res = (abi.BnResults * n)()
    _check(lib().bn_prelim_search_batches(C.c_int(volume.handle), C.c_int32(n), ptrs, C.c_int(taps), res))
    return [_results(res[k]) for k in range(n)]


class JobBatch:
    """A prepared call of bn_prelim_search_jobs: the ctypes job array is built once, `run()` is the C call alone (what
    bench.py times), `results()` converts and frees what it returned.  jobs: list of dicts with
         volume = engine.Volume (resident)  |  host_volume = synth.Volume (uploaded as part of the job)
         query  = engine.Query (resident)   |  batch = BnQueryBatch / holder (loaded as part of the job)
         gap_x_dropoff_final (with traceback=True)"""

    def __init__(self, jobs, device=0, taps=0, traceback=False):
        n = len(jobs)
        self.n, self.device, self.taps, self.traceback = n, device, taps, traceback
        self.arr = (abi.BnJob * n)()
        self.keep = []
        for k, j in enumerate(jobs):
            a = self.arr[k]
            if j.get("volume") is not None:
                a.vol_handle = j["volume"].handle
            else:
                v = j["host_volume"]
                packed = v.packed if (isinstance(v.packed, np.ndarray) and v.packed.dtype == np.uint8 and v.packed.flags.c_contiguous) \
                    else np.ascontiguousarray(v.packed, dtype=np.uint8)
                boff = np.ascontiguousarray(v.byte_off, dtype=np.int64)
                slen = np.ascontiguousarray(v.seq_len, dtype=np.int32)
                self.keep += [packed, boff, slen]
                a.vol_handle = -1
                a.packed = packed.ctypes.data
                a.packed_bytes = packed.shape[0]
                a.seq_byte_off = boff.ctypes.data
                a.seq_len = slen.ctypes.data
                a.n_seq = slen.shape[0]
            if j.get("query") is not None:
                a.query_handle = j["query"].handle
            else:
                b = j["batch"]
                b = b.batch if hasattr(b, "batch") else b
                self.keep.append(b)
                a.query_handle = -1
                a.batch = C.pointer(b)
            a.gap_x_dropoff_final = int(j.get("gap_x_dropoff_final", 0))
        self.res = (abi.BnResults * n)()
        self.tb = (abi.BnTracebackOut * n)() if traceback else None
        self._fn = lib().bn_prelim_search_jobs
        self._pending = False

    def run(self):
        if self._pending:
            self.results()
        _check(self._fn(C.c_int(self.device), C.c_int32(self.n), self.arr, C.c_int(self.taps), self.res, self.tb))
        self._pending = True

    def results(self) -> list:
        assert self._pending, "run() first"
        self._pending = False
        out = []
        for k in range(self.n):
            if self.traceback:
                t = self.tb[k]
                try:
                    pair = (abi.struct_array(C.c_void_p(t.hsps), t.n_hsps, abi.TB_HSP_DTYPE),
                            abi.struct_array(C.c_void_p(t.ops), t.n_ops, abi.EDIT_OP_DTYPE))
                finally:
                    lib().bn_free(C.c_void_p(t.hsps))
                    lib().bn_free(C.c_void_p(t.ops))
            d = _results(self.res[k])
            if self.traceback:
                d["tb"] = pair
            out.append(d)
        return out


def prelim_search_jobs(jobs, device=0, taps=0, traceback=False) -> list:
    """The job pipeline (bn_prelim_search_jobs), one call: one result dict per job; with traceback=True each also
    carries "tb" = (final HSPs, ops)."""
    jb = JobBatch(jobs, device=device, taps=taps, traceback=traceback)
    jb.run()
    return jb.results()


def prelim_search_host(holder, vol, device=0, taps=0) -> dict:
    """Reference-facing path: host buffers in, H2D + search + D2H inside the call."""
    batch = holder.batch if hasattr(holder, "batch") else holder
    packed = np.ascontiguousarray(vol.packed, dtype=np.uint8)
    boff = np.ascontiguousarray(vol.byte_off, dtype=np.int64)
    slen = np.ascontiguousarray(vol.seq_len, dtype=np.int32)
    res = abi.BnResults()
    _check(lib().bn_prelim_search_host(C.c_int(device), C.byref(batch),
                                       packed.ctypes.data_as(C.c_void_p), C.c_int64(packed.shape[0]),
                                       boff.ctypes.data_as(C.c_void_p), slen.ctypes.data_as(C.c_void_p),
                                       C.c_int32(slen.shape[0]), C.c_int(taps), C.byref(res)))
    return _results(res)


def scan_subject(volume: Volume, query: Query, oid: int, chunk_off=0, chunk_len=0) -> np.ndarray:
    """Scan tap of one subject: every chunk (chunk_len 0) or the chunk [chunk_off, chunk_off + chunk_len)."""
    p = C.POINTER(abi.BnOffsetPair)()
    n = C.c_int64(0)
    _check(lib().bn_scan_subject(C.c_int(volume.handle), C.c_int(query.handle), C.c_int32(oid),
                                 C.c_int32(chunk_off), C.c_int32(chunk_len), C.byref(p), C.byref(n)))
    try:
        return abi.struct_array(p, n.value, abi.PAIR_DTYPE)
    finally:
        lib().bn_free(p)


def get_gapped_score(volume: Volume, query: Query, oid: int, chunk_off: int, init: np.ndarray,
                     low_score=None) -> np.ndarray:
    """BlastGetGappedScoreType drop-in for one subject chunk; `init` is an INIT_DTYPE array."""
    init = np.ascontiguousarray(init, dtype=abi.INIT_DTYPE)
    p = C.POINTER(abi.BnHSP)()
    n = C.c_int64(0)
    ls = None
    if low_score is not None:
        ls_arr = np.ascontiguousarray(low_score, dtype=np.int32)
        ls = ls_arr.ctypes.data_as(C.c_void_p)
    _check(lib().bn_get_gapped_score(C.c_int(volume.handle), C.c_int(query.handle), C.c_int32(oid),
                                     C.c_int32(chunk_off), init.ctypes.data_as(C.c_void_p),
                                     C.c_int64(init.shape[0]), ls, C.byref(p), C.byref(n)))
    try:
        return abi.struct_array(p, n.value, abi.HSP_DTYPE)
    finally:
        lib().bn_free(p)


def gapped_traceback(volume: Volume, query: Query, gap_x_dropoff_final: int, items: np.ndarray):
    """BLAST_GappedAlignmentWithTraceback drop-in for a batch of start points (`items`: TB_ITEM_DTYPE).
    Returns (results: TB_RESULT_DTYPE array, ops: EDIT_OP_DTYPE array)."""
    items = np.ascontiguousarray(items, dtype=abi.TB_ITEM_DTYPE)
    pr, po, n = C.c_void_p(), C.c_void_p(), C.c_int64(0)
    _check(lib().bn_gapped_traceback(C.c_int(volume.handle), C.c_int(query.handle), C.c_int32(gap_x_dropoff_final),
                                     items.ctypes.data_as(C.c_void_p), C.c_int64(items.shape[0]),
                                     C.byref(pr), C.byref(po), C.byref(n)))
    try:
        return (abi.struct_array(pr, items.shape[0], abi.TB_RESULT_DTYPE),
                abi.struct_array(po, n.value, abi.EDIT_OP_DTYPE))
    finally:
        lib().bn_free(pr)
        lib().bn_free(po)


def traceback_hsps(volume: Volume, query: Query, gap_x_dropoff_final: int, hsps: np.ndarray):
    """Start point + alignment with traceback for preliminary HSPs (`hsps`: HSP_DTYPE array).
    Returns (items: TB_ITEM_DTYPE, results: TB_RESULT_DTYPE, ops: EDIT_OP_DTYPE)."""
    hsps = np.ascontiguousarray(hsps, dtype=abi.HSP_DTYPE)
    pi, pr, po, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64(0)
    _check(lib().bn_traceback_hsps(C.c_int(volume.handle), C.c_int(query.handle), C.c_int32(gap_x_dropoff_final),
                                   hsps.ctypes.data_as(C.c_void_p), C.c_int64(hsps.shape[0]),
                                   C.byref(pi), C.byref(pr), C.byref(po), C.byref(n)))
    try:
        return (abi.struct_array(pi, hsps.shape[0], abi.TB_ITEM_DTYPE),
                abi.struct_array(pr, hsps.shape[0], abi.TB_RESULT_DTYPE),
                abi.struct_array(po, n.value, abi.EDIT_OP_DTYPE))
    finally:
        lib().bn_free(pi)
        lib().bn_free(pr)
        lib().bn_free(po)


def traceback_search(volume: Volume, query: Query, gap_x_dropoff_final: int, hsps: np.ndarray):
    """The traceback stage for the preliminary lists `hsps` (HSP_DTYPE).  Returns (final HSPs: TB_HSP_DTYPE, ops)."""
    hsps = np.ascontiguousarray(hsps, dtype=abi.HSP_DTYPE)
    po, pe, n, ne = C.c_void_p(), C.c_void_p(), C.c_int64(0), C.c_int64(0)
    _check(lib().bn_traceback_search(C.c_int(volume.handle), C.c_int(query.handle), C.c_int32(gap_x_dropoff_final),
                                     hsps.ctypes.data_as(C.c_void_p), C.c_int64(hsps.shape[0]),
                                     C.byref(po), C.byref(n), C.byref(pe), C.byref(ne)))
    try:
        return abi.struct_array(po, n.value, abi.TB_HSP_DTYPE), abi.struct_array(pe, ne.value, abi.EDIT_OP_DTYPE)
    finally:
        lib().bn_free(po)
        lib().bn_free(pe)


def download_lookup(query: Query, device=0):
    """Parity tap: (hashtable, next_pos) of an MB batch as resident on `device`."""
    batch = query.holder.batch if hasattr(query.holder, "batch") else query.holder
    ht = np.zeros(int(batch.hashsize), dtype=np.int32)
    nx = np.zeros(int(batch.concat_len) + 1, dtype=np.int32)
    _check(lib().bn_query_download_lookup(C.c_int(query.handle), C.c_int(device),
                                          ht.ctypes.data_as(C.c_void_p), nx.ctypes.data_as(C.c_void_p)))
    return ht, nx


def bench_scan(volume: Volume, query: Query, iters: int):
    ms = C.c_double(0)
    bases = C.c_int64(0)
    hits = C.c_int64(0)
    _check(lib().bn_bench_scan(C.c_int(volume.handle), C.c_int(query.handle), C.c_int(iters),
                               C.byref(ms), C.byref(bases), C.byref(hits)))
    return ms.value, bases.value, hits.value


def dust_mask(query: np.ndarray, level=20, window=64, linker=1):
    """Symmetric DUST of one query (blastna bytes): list of inclusive (from, to) intervals, as `masks` of setup.Setup."""
    q = np.ascontiguousarray(query, dtype=np.uint8)
    p = C.POINTER(C.c_int32)()
    n = C.c_int32(0)
    _check(lib().bn_dust_mask(q.ctypes.data_as(C.c_void_p), C.c_int32(q.shape[0]), C.c_int32(level), C.c_int32(window),
                              C.c_int32(linker), C.byref(p), C.byref(n)))
    try:
        return [(int(p[2 * i]), int(p[2 * i + 1])) for i in range(n.value)]
    finally:
        lib().bn_free(p)


def dust_mask_batch(queries, level=20, window=64, linker=1, device=0):
    """Symmetric DUST of a whole query batch on the device (bn_dust_mask_batch): one list of (from, to) per query."""
    lens = np.ascontiguousarray([len(q) for q in queries], dtype=np.int32)
    cat = np.ascontiguousarray(np.concatenate(queries) if len(queries) else np.zeros(0, np.uint8), dtype=np.uint8)
    pn, pv, total = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.c_int64(0)
    _check(lib().bn_dust_mask_batch(C.c_int(device), cat.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p),
                                    C.c_int32(len(queries)), C.c_int32(level), C.c_int32(window), C.c_int32(linker),
                                    C.byref(pn), C.byref(pv), C.byref(total)))
    try:
        counts = np.ctypeslib.as_array(pn, shape=(max(len(queries), 1),))[:len(queries)].copy()
        flat = np.ctypeslib.as_array(pv, shape=(max(2 * total.value, 2),))[:2 * total.value].copy()
    finally:
        lib().bn_free(pn)
        lib().bn_free(pv)
    out, at = [], 0
    for c in counts:
        out.append([(int(flat[2 * (at + i)]), int(flat[2 * (at + i) + 1])) for i in range(int(c))])
        at += int(c)
    return out
