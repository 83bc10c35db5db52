"""ctypes binding of oracle/liboracle.so (our plain-C restatement) — TEST INFRASTRUCTURE.

Also holds `batch_from_reference`, which turns the reference engine's own dumped set-up
(lookup-table arrays, cutoffs, Karlin blocks) into the BnQueryBatch the C ABI consumes: the
"reference host feeds the drop-in" configuration.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from gblastn_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

TAP_SCAN, TAP_INIT, TAP_GAPPED = 1, 2, 4


class PortResults(C.Structure):
    _fields_ = [("hsps", C.POINTER(abi.BnHSP)), ("n_hsps", C.c_int64),
                ("init", C.POINTER(abi.BnInitHit)), ("n_init", C.c_int64),
                ("gapped", C.POINTER(abi.BnHSP)), ("n_gapped", C.c_int64),
                ("scan", C.POINTER(abi.BnOffsetPair)), ("scan_oid", C.POINTER(C.c_int32)),
                ("scan_chunk", C.POINTER(C.c_int32)), ("n_scan", C.c_int64),
                ("stats", abi.BnStats)]


_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.port_prelim_search.restype = C.c_int
    return _lib


def batch_from_reference(r: dict, *, task: str, cfg=None, use_pv=True) -> abi.BatchHolder:
    """BnQueryBatch from a refdriver.search(..., taps|=TAP_LUT) result."""
    mb = task == "megablast"
    h = abi.BatchHolder()
    b = h.batch
    n = r["num_contexts"]
    ctxs = (abi.BnContext * n)()
    for c in range(n):
        x = ctxs[c]
        x.query_offset = int(r["ctx_query_offset"][c])
        x.query_length = int(r["ctx_query_length"][c])
        x.query_index = c // 2
        x.frame = 1 if c % 2 == 0 else -1
        x.is_valid = 1
        x.length_adjustment = int(r["ctx_length_adjustment"][c])
        x.eff_searchsp = int(r["ctx_eff_searchsp"][c])
        x.x_dropoff = int(r["ctx_x_dropoff"][c])
        x.cutoff_score = int(r["ctx_cutoff_score"][c])
        x.reduced_cutoff = int(r["ctx_reduced_cutoff"][c])
        x.gapped_cutoff = int(r["ctx_gapped_cutoff"][c])
        x.gap_lambda = float(r["ctx_kbp_gap"][c, 0])
        x.gap_logK = float(r["ctx_kbp_gap"][c, 2])
    h.keep.append(ctxs)
    b.contexts = ctxs
    b.num_contexts = n
    b.num_queries = n // 2
    b.query_start = h.ptr(r["concat_query"], np.uint8)
    b.concat_len = int(r["concat_query"].shape[0] - 2)
    b.lut_type = int(r["lut_type"])
    b.word_length = int(r["word_length"])
    b.lut_word_length = int(r["lut_word_length"])
    b.scan_step = int(r["scan_step"])
    b.hashsize = int(r["hashsize"])
    if b.lut_type == abi.BN_LUT_MB:
        b.hashtable = h.ptr(r["hashtable"], np.int32)
        b.next_pos = h.ptr(r["next_pos"], np.int32)
        b.pv_array = h.ptr(r["pv_array"], np.uint32) if use_pv else None
        b.pv_array_bts = int(r["pv_array_bts"])
    elif b.lut_type == abi.BN_LUT_SMALL_NA:
        b.backbone = h.ptr(r["backbone"], np.int16)
        ov = r["overflow"] if r["overflow"] is not None else np.zeros(1, np.int16)
        b.overflow = h.ptr(ov, np.int16)
        b.overflow_len = int(ov.shape[0])
    else:
        b.na_backbone = h.ptr(r["na_backbone"], np.int32)
        ov = r["na_overflow"] if r["na_overflow"] is not None else np.zeros(1, np.int32)
        b.na_overflow = h.ptr(ov, np.int32)
        b.na_overflow_len = int(ov.shape[0])
    if r["n_masked_locations"] is not None and r["n_masked_locations"] >= 0:
        ml = r["masked_locations"] if r["masked_locations"] is not None else np.zeros(2, np.int32)
        b.masked_locations = h.ptr(ml, np.int32)
        b.n_masked_locations = int(max(r["n_masked_locations"], 0))
    b.container_type = int(r["container_type"])
    b.window_size = int(cfg.window_size) if cfg is not None else 0
    b.scan_range = int(cfg.scan_range) if cfg is not None else 0
    for i in range(256):
        b.nucl_score_table[i] = int(r["nucl_score_table"][i])
        b.matrix[i] = int(r["matrix"].reshape(-1)[i])

    def opt(name, dflt_mb, dflt_bn, none=0):
        v = getattr(cfg, name) if cfg is not None else none
        return (dflt_mb if mb else dflt_bn) if v == none else v

    b.reward = opt("reward", 1, 2)
    b.penalty = opt("penalty", -2, -3)
    b.gap_open = opt("gap_open", 0, 5, none=-1)
    b.gap_extend = opt("gap_extend", 0, 2, none=-1)
    greedy = opt("greedy", 1, 0, none=-1)
    b.gap_algo = abi.BN_GAP_GREEDY if greedy else abi.BN_GAP_DP
    b.gap_x_dropoff = int(r["gap_x_dropoff"])
    b.min_diag_separation = opt("min_diag_separation", 6, 50, none=-1)
    # sbp->round_down as the reference computed it (s_GetNuclValuesArray, core/blast_stat.c:3207-3345)
    b.round_down = int(r["round_down"])
    b.hsp_num_max = 0
    b.percent_identity = float(cfg.percent_identity) if cfg is not None else 0.0
    b.min_hit_length = int(cfg.min_hit_length) if cfg is not None else 0
    b.hitlist_size = (cfg.hitlist_size if cfg is not None and cfg.hitlist_size else 500)
    b.evalue_cutoff = (cfg.evalue if cfg is not None and cfg.evalue > 0 else 10.0)
    lsp = cfg.low_score_perc if cfg is not None else -1.0
    b.low_score_perc = 0.15 if lsp < 0 else lsp
    return h


def search(holder: abi.BatchHolder, volume, taps=TAP_INIT | TAP_GAPPED, subject_masks=None,
           subject_mask_type=1) -> dict:
    packed = np.ascontiguousarray(volume.packed, dtype=np.uint8)
    boff = np.ascontiguousarray(volume.byte_off, dtype=np.int64)
    slen = np.ascontiguousarray(volume.seq_len, dtype=np.int32)
    res = PortResults()
    if subject_masks is not None:
        sn = np.ascontiguousarray([len(m) for m in subject_masks], dtype=np.int32)
        flat = [x for m in subject_masks for iv in m for x in iv]
        siv = np.ascontiguousarray(flat if flat else [0, 0], dtype=np.int32)
        mt, sn_p, siv_p = int(subject_mask_type), sn.ctypes.data_as(C.c_void_p), siv.ctypes.data_as(C.c_void_p)
    else:
        mt, sn_p, siv_p = 0, None, None
    lib().port_prelim_search_masked.restype = C.c_int
    st = lib().port_prelim_search_masked(C.byref(holder.batch), packed.ctypes.data_as(C.c_void_p),
                                         boff.ctypes.data_as(C.c_void_p), slen.ctypes.data_as(C.c_void_p),
                                         C.c_int32(slen.shape[0]), C.c_int(taps), C.c_int32(mt), sn_p, siv_p,
                                         C.byref(res))
    try:
        out = {
            "status": st,
            "hsps": abi.struct_array(res.hsps, res.n_hsps, abi.HSP_DTYPE),
            "init": abi.struct_array(res.init, res.n_init, abi.INIT_DTYPE),
            "gapped": abi.struct_array(res.gapped, res.n_gapped, abi.HSP_DTYPE),
            "scan": abi.struct_array(res.scan, res.n_scan, abi.PAIR_DTYPE),
            "scan_oid": (np.ctypeslib.as_array(res.scan_oid, shape=(res.n_scan,)).copy()
                         if res.n_scan else np.zeros(0, np.int32)),
            "scan_chunk": (np.ctypeslib.as_array(res.scan_chunk, shape=(res.n_scan,)).copy()
                           if res.n_scan else np.zeros(0, np.int32)),
            "stats": {k: getattr(res.stats, k) for k, _ in abi.BnStats._fields_},
        }
    finally:
        lib().port_results_free(C.byref(res))
    return out


# ---- helpers to compare against refdriver tables ------------------------------------------------
def init_table(a: np.ndarray) -> np.ndarray:
    return np.stack([a[k] for k in ("oid", "chunk_off", "q_off", "s_off", "q_start", "s_start",
                                     "length", "score")], axis=1).astype(np.int32) if a.size else np.zeros((0, 8), np.int32)


def gapped_table(a: np.ndarray) -> np.ndarray:
    return np.stack([a[k] for k in ("oid", "chunk_off", "context", "q_off", "q_end", "s_off", "s_end",
                                     "score", "q_gapped_start", "s_gapped_start")], axis=1).astype(np.int32) if a.size else np.zeros((0, 10), np.int32)


def final_table(a: np.ndarray) -> np.ndarray:
    if not a.size:
        return np.zeros((0, 11), np.int32)
    bits = a["evalue"].view(np.uint64)
    lo = (bits & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32)
    hi = (bits >> np.uint64(32)).astype(np.uint32).view(np.int32)
    cols = [a[k] for k in ("oid", "context", "q_off", "q_end", "s_off", "s_end", "score",
                           "q_gapped_start", "s_gapped_start")] + [lo, hi]
    return np.stack(cols, axis=1).astype(np.int32)
