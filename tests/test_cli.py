"""python -m bwbble_b200: flag parsing mirrors the reference CLI (main.c:100-117); `index` reproduces the reference's files."""
import hashlib
import subprocess
import sys
import os

import pytest

import golden_util as G
from bwbble_b200.__main__ import build_parser, params_from_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_align_flags_map_to_aln_params():
    a = build_parser().parse_args(["align", "-n", "4", "-o", "2", "-e", "3", "-k", "3", "-l", "20", "-M", "2", "-O", "5",
                                   "-E", "1", "-m", "200", "-S", "-P", "g.fa", "r.fq", "o.aln"])
    p = params_from_args(a)
    assert (p.max_diff, p.max_gapo, p.max_gape, p.max_diff_seed, p.seed_length) == (4, 2, 3, 3, 20)
    assert (p.mm_score, p.gapo_score, p.gape_score, p.max_entries) == (2, 5, 1, 200)
    assert p.is_multiref == 0 and p.use_precalc == 1
    d = params_from_args(build_parser().parse_args(["align", "g.fa", "r.fq", "o.aln"]))
    assert (d.max_diff, d.max_gapo, d.max_gape, d.seed_length, d.max_diff_seed, d.is_multiref, d.use_precalc) == (0, 1, 6, 32, 2, 1, 0)


def test_index_subcommand_writes_the_reference_files(tmp_path):
    fa = str(tmp_path / "g.fa")
    open(fa, "wb").write(G.golden_bytes("g.fa"))
    r = subprocess.run([sys.executable, "-m", "bwbble_b200", "index", fa], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert hashlib.md5(open(fa + ".bwt", "rb").read()).hexdigest() == G.MANIFEST["md5"]["g.fa.bwt"]
    assert open(fa + ".ann", "rb").read() == G.golden_bytes("g.fa.ann")


@pytest.mark.gpu
def test_align_subcommand_writes_the_reference_aln_and_sam(tmp_path):
    fa = G.materialise_index(tmp_path)
    fq = os.path.join(G.GOLDEN, "r.fq")
    aln, sam = str(tmp_path / "o.aln"), str(tmp_path / "o.sam")
    r = subprocess.run([sys.executable, "-m", "bwbble_b200", "align", "-n", "3", "--sam", sam, "--sam-max-diff", "3", fa, fq, aln],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert open(aln, "rb").read() == G.golden_bytes("aln_n3.aln")
    assert open(sam, "rb").read() == G.golden_bytes("sam_n3.sam")
    # -P: the table is built on the device and stored next to the index on first use, like the reference does
    alnp = str(tmp_path / "p.aln")
    r = subprocess.run([sys.executable, "-m", "bwbble_b200", "align", "-P", "-n", "3", fa, fq, alnp],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert open(alnp, "rb").read() == G.golden_bytes("aln_P_n3.aln")
    assert hashlib.md5(open(fa + ".pre", "rb").read()).hexdigest() == G.MANIFEST["pre"]["multi"]["md5"]
