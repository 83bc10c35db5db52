"""ctypes mirror of include/gblastn_b200.h (POD structs of the C ABI)."""
from __future__ import annotations

import ctypes as C
import numpy as np

BN_OK = 0
BN_ERR_INVALID, BN_ERR_MEMORY, BN_ERR_NO_DEVICE, BN_ERR_CUDA, BN_ERR_UNSUPPORTED, BN_ERR_OVERFLOW = 1, 2, 3, 4, 5, 6
BN_LUT_MB, BN_LUT_SMALL_NA, BN_LUT_NA = 0, 1, 2
BN_DIAG_ARRAY, BN_DIAG_HASH = 0, 1
BN_GAP_DP, BN_GAP_GREEDY = 0, 1
BN_TAP_INIT, BN_TAP_GAPPED = 2, 4
BN_MASK_NONE, BN_MASK_SOFT, BN_MASK_HARD = 0, 1, 2


class BnContext(C.Structure):
    _fields_ = [
        ("query_offset", C.c_int32), ("query_length", C.c_int32), ("query_index", C.c_int32),
        ("frame", C.c_int32), ("is_valid", C.c_int32), ("length_adjustment", C.c_int32),
        ("eff_searchsp", C.c_int64),
        ("x_dropoff", C.c_int32), ("cutoff_score", C.c_int32), ("reduced_cutoff", C.c_int32),
        ("gapped_cutoff", C.c_int32),
        ("gap_lambda", C.c_double), ("gap_logK", C.c_double),
    ]


class BnQueryBatch(C.Structure):
    _fields_ = [
        ("query_start", C.c_void_p), ("concat_len", C.c_int32), ("num_contexts", C.c_int32),
        ("contexts", C.POINTER(BnContext)), ("num_queries", C.c_int32),
        ("lut_type", C.c_int32), ("word_length", C.c_int32), ("lut_word_length", C.c_int32),
        ("scan_step", C.c_int32), ("hashsize", C.c_int64),
        ("hashtable", C.c_void_p), ("next_pos", C.c_void_p), ("pv_array", C.c_void_p),
        ("pv_array_bts", C.c_int32),
        ("backbone", C.c_void_p), ("overflow", C.c_void_p), ("overflow_len", C.c_int64),
        ("masked_locations", C.c_void_p), ("n_masked_locations", C.c_int32),
        ("container_type", C.c_int32), ("window_size", C.c_int32), ("scan_range", C.c_int32),
        ("nucl_score_table", C.c_int32 * 256), ("matrix", C.c_int32 * 256),
        ("gap_algo", C.c_int32), ("reward", C.c_int32), ("penalty", C.c_int32),
        ("gap_open", C.c_int32), ("gap_extend", C.c_int32), ("gap_x_dropoff", C.c_int32),
        ("min_diag_separation", C.c_int32), ("round_down", C.c_int32),
        ("hsp_num_max", C.c_int32), ("hitlist_size", C.c_int32),
        ("evalue_cutoff", C.c_double), ("low_score_perc", C.c_double),
        ("lookup_segments", C.c_void_p), ("n_lookup_segments", C.c_int32),
        ("na_backbone", C.c_void_p), ("na_overflow", C.c_void_p), ("na_overflow_len", C.c_int64),
        ("percent_identity", C.c_double), ("min_hit_length", C.c_int32), ("reserved0", C.c_int32),
    ]


class BnJob(C.Structure):
    _fields_ = [("vol_handle", C.c_int), ("query_handle", C.c_int), ("batch", C.POINTER(BnQueryBatch)),
                ("packed", C.c_void_p), ("packed_bytes", C.c_int64), ("seq_byte_off", C.c_void_p),
                ("seq_len", C.c_void_p), ("n_seq", C.c_int32), ("gap_x_dropoff_final", C.c_int32)]


class BnTracebackOut(C.Structure):
    _fields_ = [("hsps", C.c_void_p), ("n_hsps", C.c_int64), ("ops", C.c_void_p), ("n_ops", C.c_int64)]


class BnOffsetPair(C.Structure):
    _fields_ = [("q_off", C.c_uint32), ("s_off", C.c_uint32)]


class BnInitHit(C.Structure):
    _fields_ = [("oid", C.c_int32), ("chunk_off", C.c_int32), ("q_off", C.c_int32),
                ("s_off", C.c_int32), ("q_start", C.c_int32), ("s_start", C.c_int32),
                ("length", C.c_int32), ("score", C.c_int32)]


class BnHSP(C.Structure):
    _fields_ = [("oid", C.c_int32), ("context", C.c_int32), ("q_off", C.c_int32),
                ("q_end", C.c_int32), ("s_off", C.c_int32), ("s_end", C.c_int32),
                ("score", C.c_int32), ("q_gapped_start", C.c_int32),
                ("s_gapped_start", C.c_int32), ("chunk_off", C.c_int32), ("evalue", C.c_double)]


class BnStats(C.Structure):
    _fields_ = [("lookup_hits", C.c_int64), ("init_extends", C.c_int64),
                ("good_init_extends", C.c_int64), ("gap_extensions", C.c_int64),
                ("good_extensions", C.c_int64), ("subject_bases_scanned", C.c_int64),
                ("ms_scan", C.c_double), ("ms_extend", C.c_double), ("ms_gapped", C.c_double),
                ("ms_host", C.c_double), ("ms_total", C.c_double), ("kernel_launches", C.c_int64)]


class BnResults(C.Structure):
    _fields_ = [("hsps", C.POINTER(BnHSP)), ("n_hsps", C.c_int64),
                ("init", C.POINTER(BnInitHit)), ("n_init", C.c_int64),
                ("gapped", C.POINTER(BnHSP)), ("n_gapped", C.c_int64),
                ("stats", BnStats)]


class BnDbFileInfo(C.Structure):
    _fields_ = [("n_seq", C.c_int32), ("max_len", C.c_int32), ("total_bases", C.c_int64),
                ("nsq_bytes", C.c_int64), ("title", C.c_char * 256)]


class BnSetupOptions(C.Structure):
    _fields_ = [("task", C.c_int32), ("word_size", C.c_int32), ("reward", C.c_int32),
                ("penalty", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32),
                ("greedy", C.c_int32), ("window_size", C.c_int32), ("scan_range", C.c_int32),
                ("min_diag_separation", C.c_int32), ("hitlist_size", C.c_int32),
                ("mask_at_hash", C.c_int32),
                ("xdrop_ungap", C.c_double), ("xdrop_gap", C.c_double),
                ("xdrop_gap_final", C.c_double), ("evalue", C.c_double),
                ("low_score_perc", C.c_double),
                ("db_length", C.c_int64), ("db_num_seqs", C.c_int32),
                ("avg_subject_length", C.c_int32), ("device_lookup", C.c_int32),
                ("hsp_num_max", C.c_int32),
                ("percent_identity", C.c_double), ("min_hit_length", C.c_int32)]


HSP_DTYPE = np.dtype([("oid", "<i4"), ("context", "<i4"), ("q_off", "<i4"), ("q_end", "<i4"),
                      ("s_off", "<i4"), ("s_end", "<i4"), ("score", "<i4"),
                      ("q_gapped_start", "<i4"), ("s_gapped_start", "<i4"), ("chunk_off", "<i4"),
                      ("evalue", "<f8")])
INIT_DTYPE = np.dtype([("oid", "<i4"), ("chunk_off", "<i4"), ("q_off", "<i4"), ("s_off", "<i4"),
                       ("q_start", "<i4"), ("s_start", "<i4"), ("length", "<i4"), ("score", "<i4")])
PAIR_DTYPE = np.dtype([("q_off", "<u4"), ("s_off", "<u4")])

TB_ITEM_DTYPE = np.dtype([("oid", "<i4"), ("context", "<i4"), ("s_shift", "<i4"), ("s_length", "<i4"),
                          ("q_start", "<i4"), ("s_start", "<i4")])
TB_RESULT_DTYPE = np.dtype([("score", "<i4"), ("query_start", "<i4"), ("query_stop", "<i4"),
                            ("subject_start", "<i4"), ("subject_stop", "<i4"), ("esp_n", "<i4"),
                            ("esp_off", "<i8"), ("status", "<i4"), ("pad", "<i4")])
EDIT_OP_DTYPE = np.dtype([("op_type", "<i4"), ("num", "<i4")])
TB_HSP_DTYPE = np.dtype([("query_index", "<i4"), ("oid", "<i4"), ("context", "<i4"), ("q_off", "<i4"), ("q_end", "<i4"),
                         ("s_off", "<i4"), ("s_end", "<i4"), ("score", "<i4"), ("num_ident", "<i4"), ("esp_n", "<i4"),
                         ("esp_off", "<i8"), ("evalue", "<f8"), ("bit_score", "<f8")])
assert HSP_DTYPE.itemsize == C.sizeof(BnHSP)
assert INIT_DTYPE.itemsize == C.sizeof(BnInitHit)


def struct_array(ptr, n, dtype) -> np.ndarray:
    """Copy n records behind a ctypes pointer into a numpy structured array."""
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = int(n) * dtype.itemsize
    out = np.empty(int(n), dtype=dtype)
    C.memmove(out.ctypes.data, C.cast(ptr, C.c_void_p).value, nbytes)      # one copy, straight into the array
    return out


class BatchHolder:
    """A BnQueryBatch plus the numpy arrays that keep its pointers alive."""

    def __init__(self):
        self.batch = BnQueryBatch()
        self.keep = []

    def ptr(self, arr, dtype):
        if arr is None:
            return None
        a = np.ascontiguousarray(arr, dtype=dtype)
        self.keep.append(a)
        return a.ctypes.data
