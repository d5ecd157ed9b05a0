"""Access to tests/golden/ (outputs of the UNMODIFIED reference, see tests/golden/make_golden.py)."""
import gzip
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))
_FLAG2FIELD = {"-M": "mm_score", "-O": "gapo_score", "-E": "gape_score", "-n": "max_diff", "-k": "max_diff_seed",
               "-o": "max_gapo", "-e": "max_gape", "-l": "seed_length", "-m": "max_entries", "-t": "n_threads"}


def grid():
    return MANIFEST["grid"]


def pgrid():
    """-P cases (and -S -P): tag -> flags"""
    return MANIFEST.get("pgrid", {})


def mixed():
    """mixed-length cases (reads of 14..60 bases on g.fa): tag -> flags"""
    return MANIFEST.get("mixed", {})


def flags_to_kwargs(flags):
    kw, i = {}, 0
    while i < len(flags):
        if flags[i] == "-P":
            kw["use_precalc"] = 1
            i += 1
        elif flags[i] == "-S":
            kw["is_multiref"] = 0
            i += 1
        else:
            kw[_FLAG2FIELD[flags[i]]] = int(flags[i + 1])
            i += 2
    return kw


def golden_bytes(name):
    return open(os.path.join(GOLDEN, name), "rb").read()


def materialise_index(tmpdir, name="g.fa"):
    """<name> + .bwt (+ .ann) as the reference wrote them, in tmpdir; returns the fasta path.
    name = "g21.fa": the multi-genome the reference's shipped sim_chr21_N100.fastq reads were planted into."""
    fa = os.path.join(str(tmpdir), name)
    with open(fa, "wb") as f:
        f.write(golden_bytes(name))
    with open(fa + ".bwt", "wb") as f:
        f.write(gzip.decompress(golden_bytes(name + ".bwt.gz")))
    with open(fa + ".ann", "wb") as f:
        f.write(golden_bytes(name + ".ann"))
    return fa
