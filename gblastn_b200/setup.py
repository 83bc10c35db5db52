"""Python face of the host-side set-up mirror (bn_setup_* in the C ABI)."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import abi, engine


class Setup:
    """Owns a BnSetup; `.batch` is the BnQueryBatch to hand to engine.Query."""

    def __init__(self, queries, *, task="megablast", db_length, db_num_seqs, masks=None, **kw):
        lib = engine.lib()
        opt = abi.BnSetupOptions()
        opt.task = 0 if task == "megablast" else 1
        opt.gap_open = -1
        opt.gap_extend = -1
        opt.greedy = -1
        opt.min_diag_separation = -1
        opt.low_score_perc = -1.0
        opt.mask_at_hash = 1
        opt.db_length = int(db_length)
        opt.db_num_seqs = int(db_num_seqs)
        opt.avg_subject_length = int(db_length // max(db_num_seqs, 1))
        for k, v in kw.items():
            if not hasattr(opt, k):
                raise KeyError(k)
            setattr(opt, k, v)
        self.opt = opt
        qcat = np.ascontiguousarray(np.concatenate(queries), dtype=np.uint8)
        qlens = np.ascontiguousarray([len(q) for q in queries], dtype=np.int32)
        if masks is not None:
            mn = np.ascontiguousarray([len(m) for m in masks], dtype=np.int32)
            flat = [x for m in masks for iv in m for x in iv]
            miv = np.ascontiguousarray(flat if flat else [0], dtype=np.int32)
            mn_p, miv_p = mn.ctypes.data_as(C.c_void_p), miv.ctypes.data_as(C.c_void_p)
        else:
            mn_p, miv_p = None, None
        self._h = C.c_void_p()
        rc = lib.bn_setup_create(C.byref(opt), C.c_int32(len(queries)), qcat.ctypes.data_as(C.c_void_p),
                                 qlens.ctypes.data_as(C.c_void_p), mn_p, miv_p, C.byref(self._h))
        if rc != 0:
            raise engine.BnError(rc, "bn_setup_create failed (unsupported option combination or invalid query)")
        self.batch = lib.bn_setup_batch(self._h).contents
        self.n_contexts = self.batch.num_contexts

    # --- views used by tests -------------------------------------------------------------------
    def _arr(self, ptr, n, ctype, dtype):
        if not ptr or n <= 0:
            return None
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).astype(dtype, copy=True)

    @property
    def hashtable(self):
        return self._arr(self.batch.hashtable, self.batch.hashsize, C.c_int32, np.int32)

    @property
    def na_backbone(self):
        return self._arr(self.batch.na_backbone, 4 * self.batch.hashsize, C.c_int32, np.int32)

    @property
    def na_overflow(self):
        return self._arr(self.batch.na_overflow, self.batch.na_overflow_len, C.c_int32, np.int32)

    @property
    def next_pos(self):
        return self._arr(self.batch.next_pos, self.batch.concat_len + 1, C.c_int32, np.int32)

    @property
    def pv_array(self):
        return self._arr(self.batch.pv_array, self.batch.hashsize >> self.batch.pv_array_bts, C.c_uint32, np.uint32)

    @property
    def backbone(self):
        return self._arr(self.batch.backbone, self.batch.hashsize, C.c_int16, np.int16)

    @property
    def overflow(self):
        return self._arr(self.batch.overflow, self.batch.overflow_len, C.c_int16, np.int16)

    @property
    def concat_query(self):
        return self._arr(self.batch.query_start, self.batch.concat_len + 2, C.c_uint8, np.uint8)

    @property
    def masked_locations(self):
        return self._arr(self.batch.masked_locations, 2 * self.batch.n_masked_locations, C.c_int32, np.int32)

    def contexts(self):
        return [self.batch.contexts[i] for i in range(self.n_contexts)]

    def kbp_std(self):
        p = engine.lib().bn_setup_kbp_std(self._h)
        return np.ctypeslib.as_array(p, shape=(4 * self.n_contexts,)).reshape(-1, 4).copy()

    def kbp_gap(self):
        p = engine.lib().bn_setup_kbp_gap(self._h)
        return np.ctypeslib.as_array(p, shape=(4 * self.n_contexts,)).reshape(-1, 4).copy()

    def gap_x_dropoff_final(self):
        return int(engine.lib().bn_setup_gap_x_dropoff_final(self._h))

    def longest_chain(self):
        return int(engine.lib().bn_setup_longest_chain(self._h))

    def free(self):
        if self._h:
            engine.lib().bn_setup_free(self._h)
            self._h = C.c_void_p()
