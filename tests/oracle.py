"""ctypes access to the CPU restatement under oracle/ (TEST INFRASTRUCTURE: the checker only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libbwbble_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "bwbble")
REF_SRC = "/root/reference/mg-aligner"


class OrcBwt(C.Structure):
    _fields_ = [("length", C.c_uint64), ("num_words", C.c_uint64), ("num_sa", C.c_uint64), ("num_occ", C.c_uint64),
                ("sa0_index", C.c_uint64), ("C", C.c_uint64 * 17), ("bwt", C.c_void_p), ("O", C.c_void_p),
                ("SA", C.c_void_p)]


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "max_diff", "max_gapo", "max_gape", "max_entries", "mm_score", "gapo_score", "gape_score",
        "seed_length", "max_diff_seed", "max_best", "no_indel_length", "matched_Ncontig",
        "use_precalc", "is_multiref", "n_threads")]


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_O", "n_O_shortcut", "n_Oalpha", "n_Oalpha_shortcut", "pops", "pushes",
                                          "exact_tail_calls", "max_heap", "hits", "max_list")]


class OrcList(C.Structure):
    _fields_ = [("v", C.c_void_p), ("n", C.c_int), ("cap", C.c_int)]


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_bwt_load.restype = C.POINTER(OrcBwt)
        L.orc_bwt_load.argtypes = [C.c_char_p, C.c_int]
        L.orc_bwt_free.argtypes = [C.POINTER(OrcBwt)]
        L.orc_O.restype = C.c_uint64
        L.orc_O.argtypes = [C.POINTER(OrcBwt), C.c_uint, C.c_uint64]
        L.orc_O_alphabet.argtypes = [C.POINTER(OrcBwt), C.c_uint64, C.POINTER(C.c_uint64 * 16), C.c_int]
        L.orc_exact_match_bounded.restype = C.c_int
        L.orc_exact_match_bounded.argtypes = [C.POINTER(OrcBwt), C.c_void_p, C.c_int, C.c_uint64, C.c_uint64,
                                              C.POINTER(OrcParams), C.POINTER(OrcList)]
        L.orc_precalc_entry.restype = C.c_int
        L.orc_precalc_entry.argtypes = [C.POINTER(OrcBwt), C.POINTER(OrcParams), C.c_uint32, C.POINTER(OrcList)]
        L.orc_read2index.restype = C.c_long
        L.orc_read2index.argtypes = [C.c_void_p, C.c_int]
        L.orc_calculate_d.argtypes = [C.POINTER(OrcBwt), C.c_void_p, C.c_int, C.c_void_p, C.POINTER(OrcParams)]
        L.orc_align.restype = C.c_int
        L.orc_align.argtypes = [C.POINTER(OrcBwt), C.POINTER(OrcParams), C.c_void_p, C.c_void_p, C.c_uint64,
                                C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(OrcStats)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_score_mask.argtypes = [C.POINTER(C.c_uint64 * 16), C.c_int]
        L.orc_default_params.argtypes = [C.POINTER(OrcParams)]
        _lib = L
    return _lib


def to_orc_params(p) -> OrcParams:
    """bwbble_b200 Params (or dict of field names) -> OrcParams."""
    o = OrcParams()
    lib().orc_default_params(C.byref(o))
    names = [n for n, _ in OrcParams._fields_]
    if isinstance(p, dict):
        for k, v in p.items():
            setattr(o, k, int(v))
    elif p is not None:
        for n in names:
            setattr(o, n, int(getattr(p, n)))
    return o


class Oracle:
    def __init__(self, bwt_path: str):
        self.h = lib().orc_bwt_load(bwt_path.encode(), 0)
        self.length = int(self.h.contents.length)

    def close(self):
        if self.h:
            lib().orc_bwt_free(self.h)
            self.h = None

    def O(self, c: int, i: int) -> int:
        return int(lib().orc_O(self.h, c, C.c_uint64(i & 0xFFFFFFFFFFFFFFFF)))

    def O_alphabet(self, i: int, inc: int) -> np.ndarray:
        occ = (C.c_uint64 * 16)()
        lib().orc_O_alphabet(self.h, C.c_uint64(i & 0xFFFFFFFFFFFFFFFF), C.byref(occ), inc)
        return np.array(list(occ), dtype=np.uint64)

    def exact_match(self, read: np.ndarray, params=None) -> np.ndarray:
        p = to_orc_params(params)
        read = np.ascontiguousarray(read, dtype=np.uint8)
        lst = OrcList()
        lib().orc_exact_match_bounded(self.h, read.ctypes.data, len(read) - 1, 0, self.length - 1, C.byref(p), C.byref(lst))
        if lst.n == 0:
            out = np.zeros((0, 2), dtype=np.uint64)
        else:
            out = np.frombuffer(C.string_at(lst.v, lst.n * 16), dtype=np.uint64).reshape(-1, 2).copy()
        if lst.v:
            lib().orc_free(lst.v)
        return out

    def precalc_entry(self, index: int, params=None) -> np.ndarray:
        """row `index` of the reference's .pre table (align.c:200-224) as an (n, 2) array of (L, U)"""
        p = to_orc_params(params)
        lst = OrcList()
        lib().orc_precalc_entry(self.h, C.byref(p), index, C.byref(lst))
        out = (np.zeros((0, 2), dtype=np.uint64) if lst.n == 0 else
               np.frombuffer(C.string_at(lst.v, lst.n * 16), dtype=np.uint64).reshape(-1, 2).copy())
        if lst.v:
            lib().orc_free(lst.v)
        return out

    def calculate_d(self, read: np.ndarray, dlen: int = 0, params=None) -> np.ndarray:
        p = to_orc_params(params)
        read = np.ascontiguousarray(read, dtype=np.uint8)
        dlen = dlen if dlen > 0 else len(read)
        D = np.zeros((dlen + 1, 2), dtype=np.int32)
        lib().orc_calculate_d(self.h, read.ctypes.data, dlen, D.ctypes.data, C.byref(p))
        return D

    def align(self, seq: np.ndarray, offsets: np.ndarray, params=None, threads: int = 1):
        """-> (.aln bytes, stats dict)"""
        p = to_orc_params(params)
        p.n_threads = threads
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        buf = C.c_void_p()
        ln = C.c_uint64()
        st = OrcStats()
        rc = lib().orc_align(self.h, C.byref(p), seq.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                             C.byref(buf), C.byref(ln), C.byref(st))
        if rc:
            raise RuntimeError("orc_align failed: %d" % rc)
        try:
            data = C.string_at(buf, ln.value)
        finally:
            lib().orc_free(buf)
        return data, {n: int(getattr(st, n)) for n, _ in OrcStats._fields_}


def pushed_scores(reset: bool = True) -> set:
    """scores of every heap entry the oracle pushed since the last reset (process-wide test hook)"""
    m = (C.c_uint64 * 16)()
    lib().orc_score_mask(C.byref(m), 1 if reset else 0)
    return {64 * k + b for k in range(16) for b in range(64) if (int(m[k]) >> b) & 1}


def have_reference_sources() -> bool:
    return os.path.isdir(REF_SRC)


def ensure_ref_binary() -> str:
    """Compile the unmodified reference into oracle/_ref/ (only where /root/reference exists)."""
    if not os.path.exists(REF_BIN):
        if not have_reference_sources():
            return ""
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)
    return REF_BIN
