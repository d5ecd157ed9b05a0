"""End to end through the reference's OWN host: oracle/_ref/bwbble_gpu = the reference's main.o, align.o,
bwt.o, io.o, is.o, exact_match.o linked with bwbble_b200/csrc/shim/bwbble_shim.c + libbwbble_b200.so
(`make -C oracle dropin`, built where /root/reference exists; the binary travels to the GPU box).
CLI, index files, .aln and SAM must be indistinguishable from the reference's."""
import os
import subprocess

import pytest

import golden_util as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_BIN = os.path.join(ROOT, "oracle", "_ref", "bwbble_gpu")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bwbble")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(GPU_BIN) and os.path.exists(REF_BIN)),
                                 reason="drop-in binaries not built (make -C oracle ref dropin needs /root/reference)")]


@pytest.mark.parametrize("tag", ["n0", "n3", "n4_o2_e3_k3_l20", "n3_t4"])
def test_dropin_cli_writes_the_reference_aln_and_sam(tmp_path, tag):
    fa = G.materialise_index(tmp_path)
    fq = os.path.join(G.GOLDEN, "r.fq")
    flags = G.grid()[tag]
    aln = str(tmp_path / "out.aln")
    r = subprocess.run([GPU_BIN, "align", *flags, fa, fq, aln], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Processed" in r.stdout
    assert open(aln, "rb").read() == G.golden_bytes("aln_%s.aln" % tag)
    if "sam_%s.sam" % tag in G.MANIFEST["md5"]:
        sam = str(tmp_path / "out.sam")
        n = flags[flags.index("-n") + 1]
        subprocess.run([REF_BIN, "aln2sam", "-n", n, fa, fq, aln, sam], check=True, stdout=subprocess.DEVNULL, timeout=600)
        assert open(sam, "rb").read() == G.golden_bytes("sam_%s.sam" % tag)


@pytest.mark.parametrize("tag", ["mixed_n3", "mixed_n3_t3"])
def test_dropin_cli_on_reads_shorter_than_the_seed(tmp_path, tag):
    """SURVEY Q6 through the reference's own main(): serial entry point and -t 3 (OpenMP entry point) on mixed.fq"""
    fa = G.materialise_index(tmp_path)
    aln = str(tmp_path / "out.aln")
    r = subprocess.run([GPU_BIN, "align", *G.mixed()[tag], fa, os.path.join(G.GOLDEN, "mixed.fq"), aln],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert open(aln, "rb").read() == G.golden_bytes("aln_%s.aln" % tag)


def test_dropin_single_genome_mode_equals_reference_binary(tmp_path):
    """-S through the drop-in CLI against the reference binary run on the same files."""
    fa = G.materialise_index(tmp_path)
    fq = os.path.join(G.GOLDEN, "r.fq")
    a, b = str(tmp_path / "gpu.aln"), str(tmp_path / "ref.aln")
    for binary, out in ((GPU_BIN, a), (REF_BIN, b)):
        r = subprocess.run([binary, "align", "-S", "-n", "3", fa, fq, out], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:]
    assert open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.parametrize("tag", ["P_n3", "SP_n3"])
def test_dropin_precalc_P(tmp_path, tag):
    """`bwbble_gpu align -P`: the reference host loads <fasta>.pre (written here by K0c, so that the run does not
    spend minutes in the reference's own precalc_sa_intervals) and hands the table to the shim."""
    from bwbble_b200 import Aligner
    fa = G.materialise_index(tmp_path)
    fq = os.path.join(G.GOLDEN, "r.fq")
    flags = G.pgrid()[tag]
    with Aligner(heap_pool_mb=256) as al:
        al.load_index(fa + ".bwt")
        al.build_precalc("-S" not in flags)
        al.write_precalc(fa + ".pre")
    aln = str(tmp_path / "out.aln")
    r = subprocess.run([GPU_BIN, "align", *flags, fa, fq, aln], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Pre-calculating" not in r.stdout          # the table came from the file
    assert open(aln, "rb").read() == G.golden_bytes("aln_%s.aln" % tag)


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_dropin_cli_on_two_gpus(tmp_path):
    """BWBBLE_GPUS=2: the shim shards every batch over two devices inside the one process; same .aln bytes"""
    fa = G.materialise_index(tmp_path)
    fq = os.path.join(G.GOLDEN, "r.fq")
    aln = str(tmp_path / "out.aln")
    env = dict(os.environ, BWBBLE_GPUS="2")
    r = subprocess.run([GPU_BIN, "align", *G.grid()["n3"], fa, fq, aln], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert open(aln, "rb").read() == G.golden_bytes("aln_n3.aln")
