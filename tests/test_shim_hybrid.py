"""The compiled boundary: oracle/shim/gblastn_b200_shim.c routes the reference engine's two seams —
aux_struct->WordFinder and aux_struct->GetGappedScore (core/blast_engine.c:926,940) — into libgblastn_b200.so, and
the reference's own Blast_RunPreliminarySearch drives everything else (subject loop, chunking, purge / sort / merge,
E-values, HSP stream, hit lists, low_score).  The hybrid must produce what the pure reference produces, tap for tap.
"""
import ctypes as C

import numpy as np
import pytest

from tests import cases


def _need_shim():
    from oracle import refdriver as R
    if not R.available() or not R.shim_available():
        pytest.skip("oracle/_ref/libblastref.so / libblastshim.so not built (needs /root/reference)")
    return R


def test_shim_library_exports_and_refuses_without_device(built):
    """CPU: the hybrid library loads, exports the binding, and a seam-1 search without a GPU fails with a status
    (no crash, no fallback to the reference's own seams)."""
    R = _need_shim()
    lib = C.CDLL(R.SHIM_LIB_PATH)
    for sym in ("bnshim_word_finder", "bnshim_get_gapped_score", "bnshim_prelim_begin", "bnshim_prelim_end",
                "bnshim_attach_volume", "bnshim_attach_resident_volume", "bnshim_detach_volume", "ref_search"):
        assert hasattr(lib, sym), sym
    # the pure reference library must not contain (or depend on) the product
    pure = C.CDLL(R.LIB_PATH)
    assert not hasattr(pure, "bnshim_attach_resident_volume")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present: the refusal path is for boxes without one")
    task, cfgkw, vol, qs = cases.make_case("mb_lut11_hash_indels")
    with R.use_shim_library():
        r = R.search(qs, vol, R.default_config(task, seam=1, **cfgkw))
    assert r["status"] != 0 and r["final"].shape[0] == 0
    # ... and the pure library refuses seam 1 outright
    r = R.search(qs, vol, R.default_config(task, seam=1, **cfgkw))
    assert r["status"] == 91


HYBRID_CASES = ["c1_megablast_10kb_vs_1mb", "mb_lut11_hash_indels", "mb_smallna_diagarray", "blastn_mb11_dp",
                "blastn_smallna_dp", "mb_lut12_stride17", "mb_with_N", "mb_ntlike_many_subjects",
                "blastn_ws7_na_table", "mb_two_hit_w40_hash", "mb_repeat_family_hitlist5",
                "blastn_repeat_family_hitlist5"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", HYBRID_CASES)
def test_hybrid_equals_pure_reference(name):
    R = _need_shim()
    task, cfgkw, vol, qs = cases.make_case(name)
    taps = R.TAP_INIT | R.TAP_GAPPED
    pure = R.search(qs, vol, R.default_config(task, taps=taps, **cfgkw))
    assert pure["status"] == 0
    with R.use_shim_library():
        hyb = R.search(qs, vol, R.default_config(task, taps=taps, seam=1, **cfgkw))
    assert hyb["status"] == 0
    assert np.array_equal(hyb["init"], pure["init"]), "init-hit lists handed to the reference differ"
    assert np.array_equal(hyb["gapped"], pure["gapped"]), "HSP lists handed to the reference differ"
    assert np.array_equal(hyb["final"], pure["final"]), "HSP stream differs (coordinates, scores or E-value bits)"
    assert np.array_equal(hyb["kept"], pure["kept"]), "hit lists held by the stream differ"
    assert pure["final"].shape[0] > 0


@pytest.mark.gpu
def test_hybrid_masked_query():
    """lut->masked_locations reaches the engine through the binding (s_TypeOfWord re-probing)."""
    R = _need_shim()
    task, cfgkw, vol, qs = cases.make_case("mb_lut11_hash_indels")
    rng = np.random.default_rng(5)
    masks = [[(int(a), int(a) + 40)] for a in rng.integers(0, 400, size=len(qs))]
    pure = R.search(qs, vol, R.default_config(task, taps=R.TAP_INIT, **cfgkw), masks=masks)
    with R.use_shim_library():
        hyb = R.search(qs, vol, R.default_config(task, taps=R.TAP_INIT, seam=1, **cfgkw), masks=masks)
    assert pure["status"] == 0 and hyb["status"] == 0 and pure["n_masked_locations"] > 0
    assert np.array_equal(hyb["init"], pure["init"]) and np.array_equal(hyb["final"], pure["final"])


@pytest.mark.gpu
def test_hybrid_four_caller_threads():
    """num_threads = 4: four reference threads, each with its own seqsrc copy and OID range, call the seams at the
    same time (api/prelim_search_runner.hpp:94-113); the engine serves them on separate lanes."""
    R = _need_shim()
    task, cfgkw, vol, qs = cases.make_case("mb_ntlike_many_subjects")
    pure = R.search(qs, vol, R.default_config(task, num_threads=4, **cfgkw))
    with R.use_shim_library():
        hyb = R.search(qs, vol, R.default_config(task, num_threads=4, seam=1, **cfgkw))
    assert pure["status"] == 0 and hyb["status"] == 0
    assert pure["final"].shape[0] > 0 and np.array_equal(hyb["final"], pure["final"])


@pytest.mark.gpu
def test_hybrid_multi_chunk_subject():
    """A subject longer than MAX_DBSEQ_LEN: the reference splits it and calls the seams once per chunk."""
    R = _need_shim()
    from gblastn_b200 import synth
    vol = synth.random_volume([200_000_300, 50_000], seed=91)
    qs = synth.planted_queries(vol, 30, 800, seed=92, planted_frac=0.9, sub_rate=0.02)
    # plant a query across the chunk seam at 199 999 900 .. 200 000 000
    b = vol.bases(0)[199_999_500:200_000_300].copy()
    qs.append(np.ascontiguousarray(b, dtype=np.uint8))
    pure = R.search(qs, vol, R.default_config("megablast", taps=R.TAP_INIT | R.TAP_GAPPED))
    with R.use_shim_library():
        hyb = R.search(qs, vol, R.default_config("megablast", taps=R.TAP_INIT | R.TAP_GAPPED, seam=1))
    assert pure["status"] == 0 and hyb["status"] == 0
    assert len(set(pure["init"][:, 1].tolist())) == 2, "both chunks must produce init hits"
    assert np.array_equal(hyb["init"], pure["init"]) and np.array_equal(hyb["gapped"], pure["gapped"])
    assert np.array_equal(hyb["final"], pure["final"])
