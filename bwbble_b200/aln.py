"""The .aln wire format between `bwbble align` and `bwbble aln2sam` (align.c:345-382, :430-483)."""
from __future__ import annotations

import struct
from typing import List, NamedTuple, Tuple


class AlnHit(NamedTuple):
    score: int
    L: int
    U: int
    num_mm: int
    num_gapo: int
    num_gape: int
    aln_length: int
    pairs: Tuple[int, ...]     # state | run<<2, path scanned from its last element to its first


def parse_aln(buf: bytes) -> List[List[AlnHit]]:
    """Parse a binary .aln stream into per-read hit lists (alnsf2alns_bin, align.c:430-483)."""
    out = []
    p = 0
    n_buf = len(buf)
    while p < n_buf:
        (n,) = struct.unpack_from("<i", buf, p)
        p += 4
        hits = []
        for _ in range(n):
            score, L, U, mm, go, ge, alen, npairs = struct.unpack_from("<iQQiiiii", buf, p)
            p += 40
            pairs = struct.unpack_from("<%di" % npairs, buf, p)
            p += 4 * npairs
            hits.append(AlnHit(score, L, U, mm, go, ge, alen, tuple(pairs)))
        out.append(hits)
    return out


def record_end(buf: bytes, p: int = 0) -> int:
    """offset just past the record (one read: count + hits) that starts at p"""
    (n,) = struct.unpack_from("<i", buf, p)
    p += 4
    for _ in range(n):
        (npairs,) = struct.unpack_from("<i", buf, p + 36)
        p += 40 + 4 * npairs
    return p


def first_difference(a: bytes, b: bytes):
    """(read index, hits_a, hits_b) of the first read whose records differ, or None."""
    pa, pb = parse_aln(a), parse_aln(b)
    for i in range(max(len(pa), len(pb))):
        ha = pa[i] if i < len(pa) else None
        hb = pb[i] if i < len(pb) else None
        if ha != hb:
            return i, ha, hb
    return None
