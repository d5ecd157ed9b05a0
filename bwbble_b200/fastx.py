"""FASTQ ingest matching fastq2reads (io.c:410-515): 4-line records, '@' resync, nt4 codes."""
from __future__ import annotations

import numpy as np

from .synth import Reads

_NT4 = np.full(256, 4, dtype=np.uint8)
for _ch, _v in ((b"Aa", 0), (b"Gg", 1), (b"Cc", 2), (b"Tt", 3)):
    for _b in _ch:
        _NT4[_b] = _v


def read_fastq(path: str, with_quals: bool = False) -> Reads:
    names, seqs, quals = [], [], []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    i = 0
    n = len(lines)
    while i < n:
        if not lines[i].startswith(b"@"):
            i += 1
            continue
        if i + 3 >= n:
            break
        names.append(lines[i][1:257].decode(errors="replace"))
        seqs.append(_NT4[np.frombuffer(lines[i + 1], dtype=np.uint8)])
        if with_quals:
            quals.append(lines[i + 3].decode(errors="replace"))
        i += 4
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    seq = np.concatenate(seqs) if seqs else np.zeros(0, dtype=np.uint8)
    return Reads(np.ascontiguousarray(seq, dtype=np.uint8), offsets, names, {"quals": quals} if with_quals else {})
