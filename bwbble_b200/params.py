"""Alignment parameters: the reference's aln_params_t (align.h:48-79) and CLI flags (main.c:100-117)."""
from __future__ import annotations

from ._lib import Params

_FLAG2FIELD = {"M": "mm_score", "O": "gapo_score", "E": "gape_score", "n": "max_diff", "k": "max_diff_seed",
               "o": "max_gapo", "e": "max_gape", "l": "seed_length", "m": "max_entries", "t": "n_threads"}


def default_params(**kw) -> Params:
    """set_default_aln_params (align.c:22-38) + overrides by field name or by CLI flag letter."""
    p = Params(max_diff=0, max_gapo=1, max_gape=6, max_entries=3000000, mm_score=3, gapo_score=11, gape_score=4,
               seed_length=32, max_diff_seed=2, max_best=30, no_indel_length=5, matched_Ncontig=0,
               use_precalc=0, is_multiref=1, n_threads=1)
    for k, v in kw.items():
        setattr(p, _FLAG2FIELD.get(k, k), int(v))
    return p


def params_to_cli(p: Params) -> list:
    """Command-line flags that make the reference CLI use exactly these parameters."""
    out = []
    for flag, field in _FLAG2FIELD.items():
        out += ["-" + flag, str(getattr(p, field))]
    if not p.is_multiref:
        out.append("-S")
    if p.use_precalc:
        out.append("-P")
    return out
