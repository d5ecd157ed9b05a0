"""FASTQ ingest with fastq2reads' record grammar (io.c:410-515): the native reader of the streaming entry point
(bwb_fastq_parse, fastq_stream.cpp), so every entry point sees a file the same way -- records start at the next '@',
the name is the rest of that line, the base line goes through nt4_table, anything up to and including the '+' line is
skipped, the quality line must be as long as the base line (it may end with the file)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .synth import Reads


def read_fastq(path: str, with_quals: bool = False) -> Reads:
    L = _lib.lib()
    seq, off, names, quals = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    n, nb, qb = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    rc = L.bwb_fastq_parse(os.fsencode(path), C.byref(seq), C.byref(off), C.byref(n), C.byref(names), C.byref(nb),
                           C.byref(quals) if with_quals else None, C.byref(qb) if with_quals else None)
    if rc != 0:
        raise _lib.BwbError(rc, "cannot parse %s (missing file, truncated record, or a quality line whose length "
                                "differs from its base line)" % path)
    try:
        offsets = np.frombuffer(C.string_at(off, (n.value + 1) * 8), dtype=np.uint64).copy()
        s = np.frombuffer(C.string_at(seq, int(offsets[-1])), dtype=np.uint8).copy()
        nm = C.string_at(names, nb.value).split(b"\0")[:n.value]
        meta = {}
        if with_quals:
            meta["quals"] = [q.decode(errors="replace") for q in C.string_at(quals, qb.value).split(b"\0")[:n.value]]
    finally:
        for p in (seq, off, names, quals):
            if p:
                L.bwb_free(p)
    return Reads(s, offsets, [x.decode(errors="replace") for x in nm], meta)
