"""bn_dust_mask (gblastn_b200/csrc/dust.cpp) against the reference's own symmetric DUST — c++/src/algo/dustmask/symdust.cpp
compiled in place against stand-in object-manager headers (oracle/Makefile target `dust`, oracle/dust_driver.cpp) — and
against the committed golden intervals (tests/golden/dust_golden.json, written by tests/golden/make_dust_golden.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libdustref.so")


def ref_dust(q, level=20, window=64, linker=1):
    lib = C.CDLL(REF)
    s = bytes(np.array([65, 67, 71, 84] + [78] * 12, dtype=np.uint8)[np.asarray(q, dtype=np.uint8)])
    out, n = C.POINTER(C.c_int)(), C.c_int(0)
    lib.ref_dust(s, len(s), level, window, linker, C.byref(out), C.byref(n))
    r = [(out[2 * i], out[2 * i + 1]) for i in range(n.value)]
    lib.ref_dust_free(out)
    return r


def dust_cases():
    """Seeded sequences rich in what DUST reacts to: homopolymers, short tandem repeats of period 1-6 with and without
    interruptions, repeats at both ends, near-threshold stretches, ambiguity codes, very short sequences."""
    rng = np.random.default_rng(2024)
    cases = []
    for k in range(160):
        n = int(rng.integers(1, 1500)) if k % 7 else int(rng.integers(0, 12))
        q = rng.integers(0, 4, size=n, dtype=np.uint8)
        for _ in range(int(rng.integers(0, 5))):
            if n < 20:
                break
            period = int(rng.integers(1, 7))
            unit = rng.integers(0, 4, size=period, dtype=np.uint8)
            m = int(rng.integers(6, min(n, 200)))
            a = int(rng.integers(0, n - m + 1)) if rng.random() < 0.8 else (0 if rng.random() < 0.5 else n - m)
            rep = np.tile(unit, m // period + 1)[:m].copy()
            noise = rng.random(m) < rng.choice([0.0, 0.02, 0.1])
            rep[noise] = rng.integers(0, 4, size=int(noise.sum()), dtype=np.uint8)
            q[a:a + m] = rep
        if k % 9 == 0 and n > 30:
            q[rng.integers(0, n, size=3)] = 14          # N
        cases.append(q)
    cases.append(np.zeros(5000, np.uint8))               # one long homopolymer
    cases.append(np.tile(np.array([1, 0], np.uint8), 400))
    return cases


@pytest.mark.parametrize("params", [(20, 64, 1), (10, 32, 5), (40, 64, 1), (20, 16, 32)])
def test_dust_matches_reference(built, params):
    from gblastn_b200 import engine as E
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libdustref.so not built (needs /root/reference)")
    masked = 0
    for q in dust_cases():
        got = E.dust_mask(q, *params)
        want = ref_dust(q, *params)
        assert got == want, f"len {len(q)} params {params}: {got[:4]} vs {want[:4]}"
        masked += len(want)
    assert masked > 100


def test_dust_matches_golden(built):
    from gblastn_b200 import engine as E
    gold = json.load(open(os.path.join(HERE, "golden", "dust_golden.json")))
    cases = dust_cases()
    assert len(gold["intervals"]) == len(cases)
    for q, want in zip(cases, gold["intervals"]):
        assert E.dust_mask(q) == [tuple(x) for x in want]


@pytest.mark.gpu
@pytest.mark.parametrize("params", [(20, 64, 1), (10, 32, 5), (40, 64, 1), (20, 16, 32)])
def test_dust_kernel_equals_host_routine(params):
    """bn_dust_mask_batch (one thread per query on the device) == bn_dust_mask per query — which the CPU tests pin to
    the reference's symdust.cpp and to the golden intervals — on the seeded cases, a C5-style batch of 5 kb queries with
    low-complexity inserts, and empty / tiny sequences."""
    from gblastn_b200 import engine as E, synth
    qs = dust_cases()
    vol = synth.random_volume([300_000], seed=5)
    big = synth.add_low_complexity(synth.planted_queries(vol, 60, 5000, seed=6, planted_frac=0.5, sub_rate=0.02), seed=7, frac=0.5)
    qs = qs + big + [np.zeros(0, np.uint8), np.zeros(2, np.uint8), np.array([0, 1, 2], np.uint8)]
    got = E.dust_mask_batch(qs, *params)
    assert len(got) == len(qs)
    masked = 0
    for q, g in zip(qs, got):
        want = E.dust_mask(q, *params)
        assert g == want, f"len {len(q)} params {params}: {g[:4]} vs {want[:4]}"
        masked += len(want)
    assert masked > 100
    if os.path.exists(REF):
        for q, g in list(zip(qs, got))[::7]:
            assert g == ref_dust(q, *params)


def test_dust_masks_reach_the_lookup_table(built):
    """DUST intervals as query masks of the product set-up == the reference's set-up given the same intervals:
    words inside the masked stretches stay out of the lookup table (mask-at-hash)."""
    from gblastn_b200 import engine as E, setup as S, synth
    from oracle import refdriver as R
    if not R.available():
        pytest.skip("reference library not present")
    vol = synth.random_volume([200_000], seed=5)
    qs = synth.add_low_complexity(synth.planted_queries(vol, 12, 1500, seed=6), seed=7, frac=1.0)
    masks = [E.dust_mask(q) for q in qs]
    assert sum(len(m) for m in masks) >= 12
    r = R.search(qs, vol, R.default_config("megablast", taps=R.TAP_LUT), masks=masks)
    s = S.Setup(qs, task="megablast", db_length=vol.total_bases, db_num_seqs=vol.n_seqs, masks=masks)
    try:
        assert np.array_equal(s.hashtable, r["hashtable"]) and np.array_equal(s.next_pos, r["next_pos"])
        assert r["n_masked_locations"] > 0
    finally:
        s.free()
