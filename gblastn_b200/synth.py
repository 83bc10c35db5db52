"""Seeded synthetic query / database generators.

DB volumes are written directly in the reference's subject wire format: ncbi2na, 4 bases
per byte, first base in the two most significant bits (inc-core/blast_util.h:52-55), every
sequence starting on a byte boundary and occupying len//4 + 1 bytes whose last byte carries
len % 4 in its low two bits, as in a .nsq volume (seqdb_reader/seqdbvol.cpp:1734-1815).
Queries are blastna bytes (0..3 = A,C,G,T; 14 = N; core/blast_encoding.c:126).
"""
from __future__ import annotations

import dataclasses
import numpy as np

PAD_BYTES = 16  # readable slack after the last sequence (scanners over-read <= 3 bytes)


@dataclasses.dataclass
class Volume:
    packed: np.ndarray      # uint8, all sequences back to back + PAD_BYTES
    byte_off: np.ndarray    # int64[n]
    seq_len: np.ndarray     # int32[n]

    @property
    def n_seqs(self) -> int:
        return int(self.seq_len.shape[0])

    @property
    def total_bases(self) -> int:
        return int(self.seq_len.astype(np.int64).sum())

    def bases(self, oid: int) -> np.ndarray:
        """Unpacked bases (uint8 0..3) of one sequence."""
        L = int(self.seq_len[oid])
        b0 = int(self.byte_off[oid])
        raw = self.packed[b0:b0 + (L + 3) // 4]
        out = np.empty((raw.shape[0], 4), dtype=np.uint8)
        out[:, 0] = raw >> 6
        out[:, 1] = (raw >> 4) & 3
        out[:, 2] = (raw >> 2) & 3
        out[:, 3] = raw & 3
        return out.reshape(-1)[:L]


def pack_bases(bases: np.ndarray) -> np.ndarray:
    """bases uint8 0..3 -> ncbi2na bytes (len//4 + 1 bytes, remainder count in last byte)."""
    L = int(bases.shape[0])
    nfull = L // 4
    out = np.zeros(nfull + 1, dtype=np.uint8)
    if nfull:
        b = bases[:nfull * 4].reshape(nfull, 4).astype(np.uint8)
        out[:nfull] = (b[:, 0] << 6) | (b[:, 1] << 4) | (b[:, 2] << 2) | b[:, 3]
    last = 0
    for k in range(L - nfull * 4):
        last |= int(bases[nfull * 4 + k]) << (6 - 2 * k)
    out[nfull] = last | (L & 3)
    return out


def make_volume_from_bases(seqs: list[np.ndarray]) -> Volume:
    chunks, offs, lens, pos = [], [], [], 0
    for s in seqs:
        p = pack_bases(np.asarray(s, dtype=np.uint8))
        chunks.append(p)
        offs.append(pos)
        lens.append(len(s))
        pos += p.shape[0]
    chunks.append(np.zeros(PAD_BYTES, dtype=np.uint8))
    return Volume(np.concatenate(chunks), np.asarray(offs, dtype=np.int64),
                  np.asarray(lens, dtype=np.int32))


def random_volume(seq_lens, seed: int) -> Volume:
    """iid uniform ACGT volume, generated directly in packed form (fast for Gb sizes)."""
    rng = np.random.default_rng(seed)
    seq_lens = np.asarray(seq_lens, dtype=np.int64)
    nbytes = seq_lens // 4 + 1
    offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
    total = int(nbytes.sum())
    packed = rng.integers(0, 256, size=total + PAD_BYTES, dtype=np.uint8)
    packed[total:] = 0
    last = offs + nbytes - 1
    rem = (seq_lens & 3).astype(np.uint8)
    keep = np.array([0x00, 0xC0, 0xF0, 0xFC], dtype=np.uint8)[rem]
    packed[last] = (packed[last] & keep) | rem
    return Volume(packed, offs, seq_lens.astype(np.int32))


_COMP = np.array([3, 2, 1, 0, 5, 4, 7, 6, 8, 9, 13, 12, 11, 10, 14, 15], dtype=np.uint8)


def revcomp(q: np.ndarray) -> np.ndarray:
    return _COMP[q[::-1]]


def mutate(seq: np.ndarray, rng, sub_rate: float, indel_rate: float = 0.0) -> np.ndarray:
    seq = seq.copy()
    n = seq.shape[0]
    if sub_rate > 0:
        m = rng.random(n) < sub_rate
        seq[m] = (seq[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
    if indel_rate > 0:
        out = []
        ev = rng.random(n)
        for i in range(n):
            if ev[i] < indel_rate / 2:
                continue                       # deletion
            out.append(seq[i])
            if ev[i] > 1 - indel_rate / 2:
                out.append(np.uint8(rng.integers(0, 4)))  # insertion
        seq = np.asarray(out, dtype=np.uint8)
    return seq


def planted_queries(vol: Volume, n_queries: int, qlen: int, seed: int, *, planted_frac=0.8,
                    sub_rate=0.02, indel_rate=0.0, rc_frac=0.5, n_frac=0.0):
    """80 % planted (mutated substrings of DB sequences, half reverse-complemented),
    20 % random queries; exact length qlen."""
    rng = np.random.default_rng(seed)
    lens64 = vol.seq_len.astype(np.int64)
    ok = np.nonzero(lens64 >= qlen + 8)[0]
    queries = []
    for _ in range(n_queries):
        if ok.size and rng.random() < planted_frac:
            oid = int(ok[rng.integers(0, ok.size)])
            L = int(lens64[oid])
            take = qlen + 8
            start = int(rng.integers(0, L - take + 1))
            b0 = int(vol.byte_off[oid]) + start // 4
            raw = vol.packed[b0:b0 + take // 4 + 2]
            un = np.empty((raw.shape[0], 4), dtype=np.uint8)
            un[:, 0] = raw >> 6
            un[:, 1] = (raw >> 4) & 3
            un[:, 2] = (raw >> 2) & 3
            un[:, 3] = raw & 3
            seg = un.reshape(-1)[start % 4: start % 4 + take]
            q = mutate(seg, rng, sub_rate, indel_rate)
            if q.shape[0] < qlen:
                q = np.concatenate([q, rng.integers(0, 4, size=qlen - q.shape[0], dtype=np.uint8)])
            q = q[:qlen]
            if rng.random() < rc_frac:
                q = revcomp(q)
        else:
            q = rng.integers(0, 4, size=qlen, dtype=np.uint8)
        if n_frac > 0:
            q = q.copy()
            q[rng.random(qlen) < n_frac] = 14
        queries.append(np.ascontiguousarray(q, dtype=np.uint8))
    return queries


def add_low_complexity(queries, seed: int, frac=0.3):
    """low-complexity stretches — poly-A (30-80 bases), (CA)n, (GGA)n — written over random places
    of `frac` of the queries (lengths unchanged), so that DUST has something to mask."""
    rng = np.random.default_rng(seed)
    out = []
    for q in queries:
        q = q.copy()
        if rng.random() < frac and q.shape[0] > 200:
            for _ in range(int(rng.integers(1, 3))):
                kind = int(rng.integers(0, 3))
                n = int(rng.integers(30, 81))
                unit = (np.array([0], np.uint8), np.array([1, 0], np.uint8), np.array([2, 2, 0], np.uint8))[kind]
                rep = np.tile(unit, n // unit.shape[0] + 1)[:n]
                a = int(rng.integers(0, q.shape[0] - n))
                q[a:a + n] = rep
        out.append(np.ascontiguousarray(q, dtype=np.uint8))
    return out
