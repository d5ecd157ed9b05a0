"""FM-index files of the reference: <fasta>.bwt (store_bwt, bwt.c:66-82) and <fasta>.ann."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib


@dataclass
class BwtIndex:
    """In-memory image of a .bwt file; field meaning as bwt_t (bwt.h:19-40)."""
    length: int
    sa0_index: int
    C: np.ndarray      # uint64[17]
    bwt: np.ndarray    # uint32[num_words], 8 symbols per word, first symbol in the top nibble
    O: np.ndarray      # uint64[num_occ*16], inclusive checkpoints every 128 rows
    SA: Optional[np.ndarray] = None   # uint64[num_sa], every 32nd suffix-array value

    @property
    def num_words(self) -> int:
        return len(self.bwt)

    @property
    def num_occ(self) -> int:
        return len(self.O) // 16

    def symbols(self) -> np.ndarray:
        """Unpacked BWT codes (uint8[length]) -- for tests on small indexes."""
        sh = (28 - 4 * np.arange(8, dtype=np.uint32))[None, :]
        return ((self.bwt[:, None] >> sh) & 15).astype(np.uint8).reshape(-1)[: self.length]


def load_bwt(path: str, load_sa: bool = False) -> BwtIndex:
    """load_bwt (bwt.c:90-125)."""
    with open(path, "rb") as f:
        hdr = np.fromfile(f, dtype=np.uint64, count=5)
        if len(hdr) != 5:
            raise IOError("short .bwt header: %s" % path)
        length, num_words, num_sa, num_occ, sa0 = (int(x) for x in hdr)
        C = np.fromfile(f, dtype=np.uint64, count=17)
        bwt = np.fromfile(f, dtype=np.uint32, count=num_words)
        O = np.fromfile(f, dtype=np.uint64, count=num_occ * 16)
        SA = np.fromfile(f, dtype=np.uint64, count=num_sa) if load_sa else None
        if len(bwt) != num_words or len(O) != num_occ * 16 or (load_sa and len(SA) != num_sa):
            raise IOError("truncated .bwt: %s" % path)
    return BwtIndex(length, sa0, C, bwt, O, SA)


def build_index(fasta_path: str, write_ref: bool = False, aligner=None) -> str:
    """`bwbble index <fasta>` (bwt.c:29-63): writes <fasta>.bwt and <fasta>.ann with the native
    host builder (bwbble_b200/csrc/index_build.cpp) or, given an Aligner, with the suffix sort and the
    BWT / checkpoint passes on its first device (K7, index_build_gpu.cu).  Returns the .bwt path."""
    if aligner is not None:
        import ctypes as C
        rounds = C.c_int(0)
        rc = _lib.lib().bwb_index_build_device(aligner._ctx, os.fsencode(fasta_path), int(write_ref), C.byref(rounds))
        _lib.check(rc, aligner._ctx)
        aligner.last_index_sort_rounds = rounds.value
        return fasta_path + ".bwt"
    rc = _lib.lib().bwb_index_build(os.fsencode(fasta_path), int(write_ref))
    _lib.check(rc)
    return fasta_path + ".bwt"
