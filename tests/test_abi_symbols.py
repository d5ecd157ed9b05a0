"""The C-ABI shared library loads and exports every symbol include/bwbble_b200.h declares; without a
GPU the product path fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

from bwbble_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "bwbble_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bwb_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libbwbble_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "%s has no ctypes signature in bwbble_b200/_lib.py" % n
    assert sorted(_lib.SIGNATURES) == names


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_lib.Params) == 60          # aln_params_t: 15 ints (align.h:48-79)
    assert ctypes.sizeof(_lib.Hit) == 48
    assert _lib.Hit.read_id.offset == 28 and _lib.Hit.runs.offset == 32


def test_default_params_are_the_reference_defaults():
    p = _lib.Params()
    _lib.lib().bwb_default_params(ctypes.byref(p))
    got = {n: getattr(p, n) for n, _ in _lib.Params._fields_}
    assert got == dict(max_diff=0, max_gapo=1, max_gape=6, max_entries=3000000, mm_score=3, gapo_score=11, gape_score=4,
                       seed_length=32, max_diff_seed=2, max_best=30, no_indel_length=5, matched_Ncontig=0,
                       use_precalc=0, is_multiref=1, n_threads=1)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from bwbble_b200 import Aligner, BwbError
    with pytest.raises(BwbError) as ei:
        Aligner()
    assert "no CPU fallback" in str(ei.value)


def test_product_package_never_uses_the_oracle():
    """Nothing under bwbble_b200/ or include/ may import, link or execute anything under oracle/."""
    bad = []
    for top in ("bwbble_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")) or f == "Makefile":
                    src = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"import\s+oracle|from\s+oracle|oracle/|libbwbble_oracle|orc_[a-z_]+\(", src):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
