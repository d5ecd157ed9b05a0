"""ctypes binding of libbwbble_b200.so (the C ABI in include/bwbble_b200.h).

The library is built in-tree by `make -C bwbble_b200/csrc` (or __graft_entry__.build()).  There is
no Python or CPU fallback: if the shared object is missing, or CUDA is unusable when a context is
created, the import / call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BWBBLE_B200_LIB") or os.path.join(_HERE, "libbwbble_b200.so")   # env override: A/B builds


class BwbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("bwbble_b200 error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    """Mirror of aln_params_t (mg-aligner/align.h:48-79); defaults = set_default_aln_params (align.c:22-38)."""
    _fields_ = [(n, C.c_int32) for n in (
        "max_diff", "max_gapo", "max_gape", "max_entries", "mm_score", "gapo_score", "gape_score",
        "seed_length", "max_diff_seed", "max_best", "no_indel_length", "matched_Ncontig",
        "use_precalc", "is_multiref", "n_threads")]


class GapRun(C.Structure):
    _fields_ = [("start", C.c_uint8), ("len", C.c_uint8), ("state", C.c_uint8), ("pad", C.c_uint8)]


class Hit(C.Structure):
    _fields_ = [("L", C.c_uint64), ("U", C.c_uint64), ("score", C.c_int32),
                ("num_mm", C.c_uint8), ("num_gapo", C.c_uint8), ("num_gape", C.c_uint8), ("aln_length", C.c_uint8),
                ("n_runs", C.c_uint8), ("pad", C.c_uint8 * 3), ("read_id", C.c_uint32), ("runs", GapRun * 4)]


assert C.sizeof(Hit) == 48 and C.sizeof(Params) == 60

# name -> (restype, argtypes); the test-suite checks that every symbol of the header is exported
SIGNATURES = {
    "bwb_default_params": (None, [C.POINTER(Params)]),
    "bwb_score_buckets": (C.c_int, [C.POINTER(Params), C.POINTER(C.c_uint8), C.c_int]),
    "bwb_set_seed_carry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "bwb_seed_donor_plan": (C.c_int, [C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "bwb_create": (C.c_void_p, [C.POINTER(C.c_int), C.c_int]),
    "bwb_destroy": (None, [C.c_void_p]),
    "bwb_last_error": (C.c_char_p, [C.c_void_p]),
    "bwb_device_count": (C.c_int, [C.c_void_p]),
    "bwb_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_longlong]),
    "bwb_set_stream": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "bwb_index_upload": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                   C.c_void_p, C.c_uint64]),
    "bwb_index_load_file": (C.c_int, [C.c_void_p, C.c_char_p]),
    "bwb_index_num_blocks": (C.c_uint64, [C.c_void_p]),
    "bwb_index_download_blocks": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bwb_occ": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "bwb_occ_alphabet": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "bwb_occ_bench": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_float),
                                C.POINTER(C.c_uint64)]),
    "bwb_exact_match": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "bwb_calculate_d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "bwb_lower_bounds": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]),
    "bwb_align": (C.c_int, [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "bwb_reads_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "bwb_reads_free": (None, [C.c_void_p]),
    "bwb_align_resident": (C.c_int, [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "bwb_results_fetch": (C.c_int, [C.c_void_p]),
    "bwb_results_num_reads": (C.c_uint64, [C.c_void_p]),
    "bwb_results_num_hits": (C.c_uint64, [C.c_void_p]),
    "bwb_results_counts": (C.c_void_p, [C.c_void_p]),
    "bwb_results_hits": (C.c_void_p, [C.c_void_p]),
    "bwb_results_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64 * 8)]),
    "bwb_results_kernel_ms": (C.c_double, [C.c_void_p]),
    "bwb_results_k3_ms": (C.c_double, [C.c_void_p]),
    "bwb_results_aln_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "bwb_results_write_aln": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "bwb_results_free": (None, [C.c_void_p]),
    "bwb_free": (None, [C.c_void_p]),
    "bwb_index_build": (C.c_int, [C.c_char_p, C.c_int]),
    "bwb_index_build_device": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]),
    "bwb_align_fastq": (C.c_longlong, [C.c_void_p, C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p,
                                       C.c_uint64, C.c_int, C.c_uint64]),
    "bwb_fastq_parse": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "bwb_sa_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "bwb_index_load_file_sa": (C.c_int, [C.c_void_p, C.c_char_p]),
    "bwb_results_locations": (C.c_void_p, [C.c_void_p]),
    "bwb_precalc_build": (C.c_int, [C.c_void_p, C.c_int]),
    "bwb_precalc_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]),
    "bwb_precalc_load_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "bwb_precalc_write": (C.c_int, [C.c_void_p, C.c_char_p]),
    "bwb_precalc_num_intervals": (C.c_uint64, [C.c_void_p]),
    "bwb_precalc_row": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "bwb_results_write_sam": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_char_p), C.c_uint64, C.c_int, C.c_char_p, C.c_int, C.c_int]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared object (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `make -C bwbble_b200/csrc` or __graft_entry__.build(); "
                "bwbble_b200 has no Python/CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, ctx=None) -> None:
    if rc != 0:
        msg = lib().bwb_last_error(ctx)
        raise BwbError(rc, msg.decode() if msg else "?")
