"""BLAST database volume files (.nin + .nsq): reader against fixtures from the reference's own test
data and an independent numpy parse; writer round trip.  Host-only (no GPU)."""
import json
import os

import numpy as np
import pytest

from gblastn_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXPECTED = json.load(open(os.path.join(GOLD, "dbfile_expected.json")))
REF_DATA = "/root/reference/c++/src/algo/blast/unit_tests/api/data"


def _prefix(name):
    p = os.path.join(GOLD, name)
    # "seqn" in dbfile_expected.json is the 2004-sequence volume of unit_tests/api/data, which stays in the reference
    # tree; tests/golden/seqn.* is the 100-sequence volume of unit_tests/seqdb_reader/data (tests/test_ambiguity.py)
    if name != "seqn" and os.path.exists(p + ".nin"):
        return p
    p = os.path.join(REF_DATA, name)
    return p if os.path.exists(p + ".nin") else None


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_reader_matches_reference_fixtures(name):
    from gblastn_b200 import engine as E
    prefix = _prefix(name)
    if prefix is None:
        pytest.skip("fixture lives in the reference tree only")
    want = EXPECTED[name]
    info, off, ln = E.dbfile_index(prefix + ".nin", prefix + ".nsq")
    assert info["title"] == want["title"]
    assert info["n_seq"] == want["n_seq"] and info["total_bases"] == want["total_bases"]
    assert info["max_len"] == want["max_len"]
    assert off[:8].tolist() == want["first_offsets"] and ln[:8].tolist() == want["first_lengths"]
    assert int((ln.astype(np.int64) * (np.arange(ln.size) + 1)).sum()) == want["length_checksum"]
    assert info["nsq_bytes"] == os.path.getsize(prefix + ".nsq")


def test_writer_round_trip(tmp_path):
    """ragged lengths incl. 1-3 base sequences and multiples of 4 (empty last byte)"""
    from gblastn_b200 import engine as E
    vol = synth.random_volume([1000, 13, 27, 4001, 4, 5, 6, 7, 1, 2, 3, 250_000, 16], seed=5)
    nin, nsq = str(tmp_path / "v.nin"), str(tmp_path / "v.nsq")
    E.dbfile_write(nin, nsq, vol, title="round trip")
    info, off, ln = E.dbfile_index(nin, nsq)
    assert info["title"] == "round trip" and info["n_seq"] == vol.n_seqs
    assert np.array_equal(ln, vol.seq_len) and info["total_bases"] == vol.total_bases
    raw = np.fromfile(nsq, dtype=np.uint8)
    assert raw[0] == 0
    for i in range(vol.n_seqs):
        whole, rem = int(ln[i]) // 4, int(ln[i]) & 3
        assert np.array_equal(raw[off[i]: off[i] + whole], vol.packed[vol.byte_off[i]: vol.byte_off[i] + whole])
        last = int(raw[off[i] + whole])
        assert last & 3 == rem
        if rem:
            m = (0xFF << (8 - 2 * rem)) & 0xFF
            assert last & m == int(vol.packed[vol.byte_off[i] + whole]) & m


def test_reader_rejects_malformed(tmp_path):
    from gblastn_b200 import engine as E
    vol = synth.random_volume([500, 90], seed=6)
    nin, nsq = str(tmp_path / "v.nin"), str(tmp_path / "v.nsq")
    E.dbfile_write(nin, nsq, vol)
    raw = bytearray(open(nin, "rb").read())
    bad = str(tmp_path / "bad.nin")
    open(bad, "wb").write(bytes([0, 0, 0, 5]) + bytes(raw[4:]))          # wrong format version
    with pytest.raises(E.BnError):
        E.dbfile_index(bad, nsq)
    open(bad, "wb").write(bytes(raw[:-6]))                                # truncated offset arrays
    with pytest.raises(E.BnError):
        E.dbfile_index(bad, nsq)
    open(str(tmp_path / "short.nsq"), "wb").write(open(nsq, "rb").read()[:40])   # index points past the file
    with pytest.raises(E.BnError):
        E.dbfile_index(nin, str(tmp_path / "short.nsq"))
