"""ctypes binding of oracle/_ref/libblastref.so (the reference engine + our tap driver).

TEST INFRASTRUCTURE: importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libblastref.so")
# the same reference engine + driver with oracle/shim linked in: cfg.seam = 1 routes the word finder and the gapped
# stage through libgblastn_b200.so (needs a GPU); loaded only by tests/test_shim_hybrid.py
SHIM_LIB_PATH = os.path.join(_HERE, "_ref", "libblastshim.so")


class RefConfig(C.Structure):
    _fields_ = [
        ("task", C.c_int32), ("word_size", C.c_int32), ("reward", C.c_int32),
        ("penalty", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32),
        ("greedy", C.c_int32), ("window_size", C.c_int32), ("scan_range", C.c_int32),
        ("min_diag_separation", C.c_int32), ("hitlist_size", C.c_int32),
        ("mask_at_hash", C.c_int32),
        ("xdrop_ungap", C.c_double), ("xdrop_gap", C.c_double), ("xdrop_gap_final", C.c_double),
        ("evalue", C.c_double), ("low_score_perc", C.c_double),
        ("db_length", C.c_int64), ("db_num_seqs", C.c_int32), ("num_threads", C.c_int32),
        ("taps", C.c_int32), ("prelim_only", C.c_int32),
        ("smask_type", C.c_int32), ("smask_n", C.c_void_p), ("smask_iv", C.c_void_p),
        ("hsp_num_max", C.c_int32), ("amb_first", C.c_void_p), ("amb_runs", C.c_void_p), ("seam", C.c_int32),
        ("percent_identity", C.c_double), ("min_hit_length", C.c_int32),
    ]


class RefTable(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_int32)), ("rows", C.c_int64), ("cap", C.c_int64),
                ("ncol", C.c_int32)]


class RefResult(C.Structure):
    _fields_ = [
        ("scan", RefTable), ("init", RefTable), ("gapped", RefTable), ("final_", RefTable),
        ("num_contexts", C.c_int32),
        ("ctx_query_offset", C.POINTER(C.c_int32)), ("ctx_query_length", C.POINTER(C.c_int32)),
        ("ctx_length_adjustment", C.POINTER(C.c_int32)),
        ("ctx_eff_searchsp", C.POINTER(C.c_int64)),
        ("ctx_x_dropoff", C.POINTER(C.c_int32)), ("ctx_cutoff_score", C.POINTER(C.c_int32)),
        ("ctx_reduced_cutoff", C.POINTER(C.c_int32)), ("ctx_gapped_cutoff", C.POINTER(C.c_int32)),
        ("ctx_kbp_std", C.POINTER(C.c_double)), ("ctx_kbp_gap", C.POINTER(C.c_double)),
        ("gap_x_dropoff", C.c_int32), ("gap_x_dropoff_final", C.c_int32),
        ("container_type", C.c_int32), ("round_down", C.c_int32),
        ("nucl_score_table", C.c_int32 * 256), ("matrix", C.c_int32 * 256),
        ("lut_type", C.c_int32), ("lut_word_length", C.c_int32), ("word_length", C.c_int32),
        ("scan_step", C.c_int32), ("longest_chain", C.c_int32), ("pv_array_bts", C.c_int32),
        ("hashsize", C.c_int64), ("next_pos_len", C.c_int64), ("pv_len", C.c_int64),
        ("overflow_len", C.c_int64),
        ("hashtable", C.POINTER(C.c_int32)), ("next_pos", C.POINTER(C.c_int32)),
        ("pv_array", C.POINTER(C.c_uint32)),
        ("backbone", C.POINTER(C.c_int16)), ("overflow", C.POINTER(C.c_int16)),
        ("n_masked_locations", C.c_int32), ("masked_locations", C.POINTER(C.c_int32)),
        ("concat_len", C.c_int32), ("concat_query", C.POINTER(C.c_uint8)),
        ("lookup_hits", C.c_int64), ("init_extends", C.c_int64),
        ("good_init_extends", C.c_int64), ("gap_extensions", C.c_int64),
        ("good_extensions", C.c_int64),
        ("seconds_prelim", C.c_double), ("status", C.c_int32),
        ("na_backbone", C.POINTER(C.c_int32)), ("na_overflow", C.POINTER(C.c_int32)),
        ("na_overflow_len", C.c_int64),
        ("tb_calls", RefTable), ("tb_ops", RefTable), ("tb_final", RefTable),
        ("seconds_traceback", C.c_double),
        ("kept", RefTable),
    ]


_lib = None
_shim_lib = None
_use_shim = False


def available() -> bool:
    return os.path.exists(LIB_PATH)


def shim_available() -> bool:
    return os.path.exists(SHIM_LIB_PATH)


def lib():
    global _lib, _shim_lib
    if _use_shim:
        if _shim_lib is None:
            _shim_lib = C.CDLL(SHIM_LIB_PATH)
            _shim_lib.ref_search.restype = C.c_int
            _shim_lib.ref_free_result.restype = None
        return _shim_lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_search.restype = C.c_int
        _lib.ref_free_result.restype = None
    return _lib


class use_shim_library:
    """Context manager: calls inside go to libblastshim.so (reference engine + B200 seams)."""

    def __enter__(self):
        global _use_shim
        _use_shim = True

    def __exit__(self, *a):
        global _use_shim
        _use_shim = False


TAP_SCAN, TAP_INIT, TAP_GAPPED, TAP_LUT, TAP_TRACEBACK = 1, 2, 4, 8, 16

TB_CALL_COLS = ("kind", "oid", "context", "s_shift", "q_start", "s_start", "q_len", "s_len",
                "score", "query_start", "query_stop", "subject_start", "subject_stop", "esp_off", "esp_n")
TB_FINAL_COLS = ("query_index", "oid", "context", "q_off", "q_end", "s_off", "s_end", "score", "num_ident",
                 "evalue_lo", "evalue_hi", "bits_lo", "bits_hi", "esp_off", "esp_n")

SCAN_COLS = ("oid", "chunk_off", "q_off", "s_off")
INIT_COLS = ("oid", "chunk_off", "q_off", "s_off", "q_start", "s_start", "length", "score")
GAPPED_COLS = ("oid", "chunk_off", "context", "q_off", "q_end", "s_off", "s_end", "score",
               "q_gapped_start", "s_gapped_start")
FINAL_COLS = ("oid", "context", "q_off", "q_end", "s_off", "s_end", "score",
              "q_gapped_start", "s_gapped_start", "evalue_lo", "evalue_hi")


def _tab(t: RefTable) -> np.ndarray:
    if t.rows == 0:
        return np.zeros((0, t.ncol), dtype=np.int32)
    a = np.ctypeslib.as_array(t.data, shape=(t.rows * t.ncol,))
    return a.reshape(t.rows, t.ncol).copy()


def _arr(ptr, n, dtype):
    if not ptr or n <= 0:
        return None
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def default_config(task="megablast", **kw) -> RefConfig:
    cfg = RefConfig()
    cfg.task = 0 if task == "megablast" else 1
    cfg.gap_open = -1
    cfg.gap_extend = -1
    cfg.greedy = -1
    cfg.min_diag_separation = -1
    cfg.low_score_perc = -1.0
    cfg.mask_at_hash = 1
    cfg.num_threads = 1
    cfg.prelim_only = 1
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    return cfg


def traceback_calls(queries, volume, items, cfg: RefConfig, ambiguity=None):
    """The reference's alignment-with-traceback routine on arbitrary start points.  `items`: int32 array (n, 6) of
    {oid, context, s_shift, s_length, q_start, s_start}.  Returns the same dict as search(); the calls are in
    tb_calls / tb_ops."""
    return search(queries, volume, cfg, tb_items=np.ascontiguousarray(items, dtype=np.int32).reshape(-1, 6),
                  ambiguity=ambiguity)


def search(queries, volume, cfg: RefConfig | None = None, *, task="megablast", masks=None,
           subject_masks=None, subject_mask_type=1, tb_items=None, ambiguity=None, **kw):
    """Run the reference preliminary search. `queries`: list of uint8 blastna arrays;
    `volume`: gblastn_b200.synth.Volume (or any object with packed/byte_off/seq_len);
    `masks`: optional list (per query) of [(left, right)] inclusive plus-strand intervals;
    `subject_masks`: optional list (per subject) of [(begin, end)] half-open masked intervals (database masks,
    soft = 1 / hard = 2).  Returns a dict of numpy arrays."""
    if cfg is None:
        cfg = default_config(task, **kw)
    keep = []
    if ambiguity is not None:       # (first: int64[n + 1], runs: int32[k, 3]) as gblastn_b200.engine.dbfile_ambiguity returns
        af = np.ascontiguousarray(ambiguity[0], dtype=np.int64)
        ar = np.ascontiguousarray(ambiguity[1], dtype=np.int32).reshape(-1)
        if ar.size == 0:
            ar = np.zeros(3, np.int32)
        keep += [af, ar]
        cfg.amb_first, cfg.amb_runs = af.ctypes.data, ar.ctypes.data
    else:
        cfg.amb_first, cfg.amb_runs = None, None
    if subject_masks is not None:
        sn = np.ascontiguousarray([len(m) for m in subject_masks], dtype=np.int32)
        sflat = [x for m in subject_masks for iv in m for x in iv]
        siv = np.ascontiguousarray(sflat if sflat else [0, 0], dtype=np.int32)
        keep += [sn, siv]
        cfg.smask_type = int(subject_mask_type)
        cfg.smask_n = sn.ctypes.data
        cfg.smask_iv = siv.ctypes.data
    else:
        cfg.smask_type = 0
        cfg.smask_n = None
        cfg.smask_iv = None
    qcat = np.ascontiguousarray(np.concatenate(queries) if len(queries) else np.zeros(0, np.uint8),
                                dtype=np.uint8)
    qlens = np.ascontiguousarray([len(q) for q in queries], dtype=np.int32)
    packed = np.ascontiguousarray(volume.packed, dtype=np.uint8)
    boff = np.ascontiguousarray(volume.byte_off, dtype=np.int64)
    slen = np.ascontiguousarray(volume.seq_len, dtype=np.int32)
    if masks is not None:
        mn = np.ascontiguousarray([len(m) for m in masks], dtype=np.int32)
        flat = [x for m in masks for iv in m for x in iv]
        miv = np.ascontiguousarray(flat if flat else [0], dtype=np.int32)
        mn_p, miv_p = mn.ctypes.data_as(C.c_void_p), miv.ctypes.data_as(C.c_void_p)
    else:
        mn_p, miv_p = None, None
    res = RefResult()
    if tb_items is not None:
        lib().ref_traceback_calls.restype = C.c_int
        st = lib().ref_traceback_calls(C.byref(cfg), C.c_int32(len(queries)),
                                       qcat.ctypes.data_as(C.c_void_p), qlens.ctypes.data_as(C.c_void_p),
                                       mn_p, miv_p, C.c_int32(slen.shape[0]),
                                       packed.ctypes.data_as(C.c_void_p), boff.ctypes.data_as(C.c_void_p),
                                       slen.ctypes.data_as(C.c_void_p), C.c_int32(tb_items.shape[0]),
                                       tb_items.ctypes.data_as(C.c_void_p), C.byref(res))
    else:
        st = lib().ref_search(C.byref(cfg), C.c_int32(len(queries)),
                              qcat.ctypes.data_as(C.c_void_p), qlens.ctypes.data_as(C.c_void_p),
                              mn_p, miv_p, C.c_int32(slen.shape[0]),
                              packed.ctypes.data_as(C.c_void_p), boff.ctypes.data_as(C.c_void_p),
                              slen.ctypes.data_as(C.c_void_p), C.byref(res))
    try:
        n = res.num_contexts
        out = {
            "status": st,
            "scan": _tab(res.scan), "init": _tab(res.init), "gapped": _tab(res.gapped),
            "final": _tab(res.final_),
            "kept": _tab(res.kept) if res.kept.ncol else np.zeros((0, 4), np.int32),
            "tb_calls": _tab(res.tb_calls), "tb_ops": _tab(res.tb_ops), "tb_final": _tab(res.tb_final),
            "num_contexts": n,
            "ctx_query_offset": _arr(res.ctx_query_offset, n, np.int32),
            "ctx_query_length": _arr(res.ctx_query_length, n, np.int32),
            "ctx_length_adjustment": _arr(res.ctx_length_adjustment, n, np.int32),
            "ctx_eff_searchsp": _arr(res.ctx_eff_searchsp, n, np.int64),
            "ctx_x_dropoff": _arr(res.ctx_x_dropoff, n, np.int32),
            "ctx_cutoff_score": _arr(res.ctx_cutoff_score, n, np.int32),
            "ctx_reduced_cutoff": _arr(res.ctx_reduced_cutoff, n, np.int32),
            "ctx_gapped_cutoff": _arr(res.ctx_gapped_cutoff, n, np.int32),
            "ctx_kbp_std": None if n == 0 else _arr(res.ctx_kbp_std, 4 * n, np.float64).reshape(n, 4),
            "ctx_kbp_gap": None if n == 0 else _arr(res.ctx_kbp_gap, 4 * n, np.float64).reshape(n, 4),
            "gap_x_dropoff": res.gap_x_dropoff, "gap_x_dropoff_final": res.gap_x_dropoff_final,
            "container_type": res.container_type, "round_down": res.round_down,
            "nucl_score_table": np.array(res.nucl_score_table, dtype=np.int32),
            "matrix": np.array(res.matrix, dtype=np.int32).reshape(16, 16),
            "lut_type": res.lut_type, "lut_word_length": res.lut_word_length,
            "word_length": res.word_length, "scan_step": res.scan_step,
            "longest_chain": res.longest_chain, "pv_array_bts": res.pv_array_bts,
            "hashsize": res.hashsize,
            "hashtable": _arr(res.hashtable, res.hashsize, np.int32),
            "next_pos": _arr(res.next_pos, res.next_pos_len, np.int32),
            "pv_array": _arr(res.pv_array, res.pv_len, np.uint32),
            "backbone": _arr(res.backbone, res.hashsize, np.int16),
            "overflow": _arr(res.overflow, res.overflow_len, np.int16),
            "na_backbone": _arr(res.na_backbone, 4 * res.hashsize, np.int32) if res.lut_type == 2 else None,
            "na_overflow": _arr(res.na_overflow, res.na_overflow_len, np.int32) if res.lut_type == 2 else None,
            "n_masked_locations": res.n_masked_locations,
            "masked_locations": (_arr(res.masked_locations, 2 * max(res.n_masked_locations, 0), np.int32)
                                 if res.n_masked_locations > 0 else None),
            "concat_query": _arr(res.concat_query, res.concat_len + 2, np.uint8),
            "lookup_hits": res.lookup_hits, "init_extends": res.init_extends,
            "good_init_extends": res.good_init_extends, "gap_extensions": res.gap_extensions,
            "good_extensions": res.good_extensions, "seconds_prelim": res.seconds_prelim,
            "seconds_traceback": res.seconds_traceback,
        }
    finally:
        lib().ref_free_result(C.byref(res))
    return out


def evalue_bits(final: np.ndarray) -> np.ndarray:
    """uint64 bit patterns of the E-values in a `final` table."""
    lo = final[:, 9].astype(np.uint32).astype(np.uint64)
    hi = final[:, 10].astype(np.uint32).astype(np.uint64)
    return lo | (hi << np.uint64(32))
