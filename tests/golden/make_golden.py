"""Generate the golden fixtures with the UNMODIFIED reference (run here, where /root/reference exists):

    python tests/golden/make_golden.py

Builds oracle/_ref/bwbble from /root/reference/mg-aligner (make -C oracle ref), generates a small
seeded multi-genome + reads, and records what the real reference produces:
    g.fa, r.fq                      inputs (committed so the fixtures do not depend on the generator)
    g.fa.bwt.gz, g.fa.ann           `bwbble index g.fa`
    aln_<tag>.aln                   `bwbble align <flags> g.fa r.fq out.aln`   for every entry of GRID
    sam_<tag>.sam                   `bwbble aln2sam -n <n> g.fa r.fq out.aln out.sam` (SAM = "next" row)
    aln_<tag>.aln for PGRID         the same with -P (12-mer seed table); the 68 MB g.fa.pre the
                                    reference computes on the way is pinned by its md5 and its interval count
    manifest.json                   tag -> flags, md5 of every file
`python tests/golden/make_golden.py --precalc-only` adds the PGRID files to an existing golden set;
`--shipped-fastq` adds the reference's own test_data/sim_chr21_N100.fastq (BASELINE configs[0]) with the outputs
the reference produces for it on a small multi-genome the reads' loci were planted into (g21.fa);
`--mixed-lengths` adds mixed.fq (reads of 14..60 bases, many no longer than the seed: SURVEY Q6) and its .aln files.
The reference repository ships no golden vectors of its own, so these ARE the pin.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GRID = {
    "n0": ["-n", "0"],
    "n1": ["-n", "1"],
    "n2": ["-n", "2"],
    "n3": ["-n", "3"],
    "n5": ["-n", "5"],
    "n3_t4": ["-n", "3", "-t", "4"],
    "n4_o2_e3_k3_l20": ["-n", "4", "-o", "2", "-e", "3", "-k", "3", "-l", "20"],
    "n3_M2_O5_E2": ["-n", "3", "-M", "2", "-O", "5", "-E", "2"],
    "n4_l0": ["-n", "4", "-l", "0"],
    "n6_o2_M4_O4_E4": ["-n", "6", "-o", "2", "-M", "4", "-O", "4", "-E", "4"],
    "n3_o0": ["-n", "3", "-o", "0"],
    "n4_m200": ["-n", "4", "-m", "200"],
}

# -P cases; multi-genome and -S tables differ, so each mode gets its own directory / .pre file
PGRID = {
    "P_n0": ["-P", "-n", "0"],
    "P_n3": ["-P", "-n", "3"],
    "P_n5": ["-P", "-n", "5"],
    "P_n4_o2_e3_k3_l20": ["-P", "-n", "4", "-o", "2", "-e", "3", "-k", "3", "-l", "20"],
    "SP_n3": ["-S", "-P", "-n", "3"],
}


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def add_precalc(ref, manifest):
    import numpy as np
    import golden_util
    run = lambda *a: subprocess.run([ref, *a], check=True, stdout=subprocess.DEVNULL)
    manifest["pgrid"] = PGRID
    manifest["pre"] = {}
    for mode in ("multi", "single"):
        with tempfile.TemporaryDirectory() as d:
            fa = golden_util.materialise_index(d)
            fq = os.path.join(HERE, "r.fq")
            for tag, flags in PGRID.items():
                if ("-S" in flags) != (mode == "single"):
                    continue
                aln = os.path.join(d, "out.aln")
                run("align", *flags, fa, fq, aln)          # the first run of a mode writes g.fa.pre (minutes)
                shutil.copy(aln, os.path.join(HERE, "aln_%s.aln" % tag))
                manifest["md5"]["aln_%s.aln" % tag] = md5(aln)
            raw = np.fromfile(fa + ".pre", dtype=np.uint8)
            manifest["pre"][mode] = {"md5": md5(fa + ".pre"), "bytes": int(raw.size),
                                     "intervals": int((raw.size - 4 * (1 << 24)) // 16)}


def add_shipped_fastq(ref, manifest):
    """BASELINE configs[0]: the one real input the reference ships, test_data/sim_chr21_N100.fastq (100 reads of
    100 bp simulated from chr21; mg-aligner/README.md:33-38).  Its FASTA is not in the tree, so the reads are planted
    into a small multi-genome (each read's locus + random spacers, some sites turned into IUPAC SNP codes) and the
    reference's own index / align / aln2sam outputs on it are recorded."""
    import numpy as np
    src = "/root/reference/test_data/sim_chr21_N100.fastq"
    shutil.copy(src, os.path.join(HERE, "sim_chr21_N100.fastq"))
    os.chmod(os.path.join(HERE, "sim_chr21_N100.fastq"), 0o644)
    lines = open(src).read().split("\n")
    seqs = [lines[i + 1] for i in range(0, len(lines) - 3, 4)]
    rng = np.random.default_rng(2121)
    iupac = {"A": "RWM", "C": "YSM", "G": "RSK", "T": "YWK"}
    parts = []
    for k, sq in enumerate(seqs):
        sq = list(sq.upper())
        for pos in rng.choice(len(sq), size=2, replace=False):          # two SNP sites per locus
            if sq[pos] in iupac:
                sq[pos] = iupac[sq[pos]][int(rng.integers(0, 3))]
        if k % 7 == 3:                                                    # a substitution the read has to pay for
            pos = int(rng.integers(20, 80))
            sq[pos] = "ACGT"[("ACGT".index(sq[pos]) + 1) % 4] if sq[pos] in "ACGT" else sq[pos]
        parts.append("".join(sq))
        parts.append("".join("ACGT"[i] for i in rng.integers(0, 4, size=int(rng.integers(30, 90)))))
    genome = "".join(parts)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "g21.fa")
        with open(fa, "w") as f:
            f.write(">chr21_loci planted sim_chr21_N100 reads\n")
            for i in range(0, len(genome), 60):
                f.write(genome[i:i + 60] + "\n")
        run = lambda *a: subprocess.run([ref, *a], check=True, stdout=subprocess.DEVNULL)
        run("index", fa)
        shutil.copy(fa, os.path.join(HERE, "g21.fa"))
        shutil.copy(fa + ".ann", os.path.join(HERE, "g21.fa.ann"))
        with open(fa + ".bwt", "rb") as s2, gzip.GzipFile(os.path.join(HERE, "g21.fa.bwt.gz"), "wb", mtime=0) as dst:
            dst.write(s2.read())
        manifest["md5"]["g21.fa.bwt"] = md5(fa + ".bwt")
        manifest["shipped"] = {"sim_n0": ["-n", "0"], "sim_n5": ["-n", "5"]}
        for tag, flags in manifest["shipped"].items():
            aln = os.path.join(d, "out.aln")
            run("align", *flags, fa, src, aln)
            shutil.copy(aln, os.path.join(HERE, "aln_%s.aln" % tag))
            manifest["md5"]["aln_%s.aln" % tag] = md5(aln)
            sam = os.path.join(d, "out.sam")
            run("aln2sam", "-n", flags[1], fa, src, aln, sam)
            shutil.copy(sam, os.path.join(HERE, "sam_%s.sam" % tag))
            manifest["md5"]["sam_%s.sam" % tag] = md5(sam)
        for f in ("g21.fa", "g21.fa.ann", "sim_chr21_N100.fastq"):
            manifest["md5"][f] = md5(os.path.join(HERE, f))


MIXED = {
    "mixed_n3": ["-n", "3"],
    "mixed_n3_t3": ["-n", "3", "-t", "3"],
    "mixed_n2_k1_l40": ["-n", "2", "-k", "1", "-l", "40"],
    "mixed_P_n3": ["-P", "-n", "3"],
}


def add_mixed_lengths(ref, manifest):
    """reads no longer than the seed (-l, default 32) do not get a D_seed of their own -- they consult the
    array the previous longer read of their thread left behind (inexact_match.c:36,62-64 serial; :121,141-143 OpenMP;
    with -P a read skipped for an N in its 12-mer leaves it untouched, :50-57).  mixed.fq = 400 reads of 14..60 bases."""
    import golden_util
    from bwbble_b200 import synth
    run = lambda *a: subprocess.run([ref, *a], check=True, stdout=subprocess.DEVNULL)
    g = synth.make_genome(101, 24000, n_records=2, snp_rate=0.015, tri_frac=0.08, n_bubbles=16, n_frac=0.03,
                          n_repeat_copies=8, repeat_len=200, n_microsats=3, lowercase_frac=0.01)      # = g.fa (checked below)
    reads = synth.make_reads(g, 108, 400, 60, 2, n_base_frac=0.004, bubble_frac=0.1, ragged=(14, 60))    # (seed chosen so that the serial and the OpenMP driver disagree)
    reads.names = ["m%d" % i for i in range(reads.n)]
    manifest["mixed"] = MIXED
    with tempfile.TemporaryDirectory() as d:
        fa = golden_util.materialise_index(d)
        chk = os.path.join(d, "chk.fa")
        g.write_fasta(chk)
        assert open(chk, "rb").read() == open(fa, "rb").read(), "synth.make_genome no longer reproduces g.fa"
        fq = os.path.join(HERE, "mixed.fq")
        reads.write_fastq(fq)
        manifest["md5"]["mixed.fq"] = md5(fq)
        for tag, flags in MIXED.items():
            aln = os.path.join(d, "out.aln")
            run("align", *flags, fa, fq, aln)              # -P: the reference first writes g.fa.pre (minutes)
            shutil.copy(aln, os.path.join(HERE, "aln_%s.aln" % tag))
            manifest["md5"]["aln_%s.aln" % tag] = md5(aln)


def main():
    if "--mixed-lengths" in sys.argv:
        import oracle
        ref = oracle.ensure_ref_binary()
        assert ref, "/root/reference is required to (re)generate the golden files"
        manifest = json.load(open(os.path.join(HERE, "manifest.json")))
        add_mixed_lengths(ref, manifest)
        json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
        return
    if "--shipped-fastq" in sys.argv:
        import oracle
        ref = oracle.ensure_ref_binary()
        assert ref, "/root/reference is required to (re)generate the golden files"
        manifest = json.load(open(os.path.join(HERE, "manifest.json")))
        add_shipped_fastq(ref, manifest)
        json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
        return
    if "--precalc-only" in sys.argv:
        import oracle
        ref = oracle.ensure_ref_binary()
        assert ref, "/root/reference is required to (re)generate the golden files"
        manifest = json.load(open(os.path.join(HERE, "manifest.json")))
        add_precalc(ref, manifest)
        json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
        return
    import oracle
    from bwbble_b200 import synth
    ref = oracle.ensure_ref_binary()
    assert ref, "/root/reference is required to (re)generate the golden files"
    g = synth.make_genome(101, 24000, n_records=2, snp_rate=0.015, tri_frac=0.08, n_bubbles=16, n_frac=0.03,
                          n_repeat_copies=8, repeat_len=200, n_microsats=3, lowercase_frac=0.01)
    reads = synth.make_reads(g, 102, 160, 100, 3, indel_frac=0.2, n_base_frac=0.004, bubble_frac=0.1)
    short = synth.make_reads(g, 103, 40, 150, 4, indel_frac=0.3, n_base_frac=0.01, ragged=(36, 150))
    with tempfile.TemporaryDirectory() as d:
        fa, fq = os.path.join(d, "g.fa"), os.path.join(d, "r.fq")
        g.write_fasta(fa)
        reads.write_fastq(fq)
        with open(fq, "ab") as f:                      # ragged 36..150 bp reads appended
            tmp = os.path.join(d, "s.fq")
            short.names = ["s%d" % i for i in range(short.n)]
            short.write_fastq(tmp)
            f.write(open(tmp, "rb").read())
        run = lambda *a: subprocess.run([ref, *a], check=True, stdout=subprocess.DEVNULL)
        run("index", fa)
        manifest = {"grid": GRID, "md5": {}}
        shutil.copy(fa, os.path.join(HERE, "g.fa"))
        shutil.copy(fq, os.path.join(HERE, "r.fq"))
        shutil.copy(fa + ".ann", os.path.join(HERE, "g.fa.ann"))
        with open(fa + ".bwt", "rb") as src, gzip.GzipFile(os.path.join(HERE, "g.fa.bwt.gz"), "wb", mtime=0) as dst:
            dst.write(src.read())
        manifest["md5"]["g.fa.bwt"] = md5(fa + ".bwt")
        for tag, flags in GRID.items():
            aln = os.path.join(d, "out.aln")
            run("align", *flags, fa, fq, aln)
            shutil.copy(aln, os.path.join(HERE, "aln_%s.aln" % tag))
            manifest["md5"]["aln_%s.aln" % tag] = md5(aln)
            if tag in ("n0", "n3", "n4_o2_e3_k3_l20"):
                sam = os.path.join(d, "out.sam")
                n = flags[flags.index("-n") + 1]
                run("aln2sam", "-n", n, fa, fq, aln, sam)
                shutil.copy(sam, os.path.join(HERE, "sam_%s.sam" % tag))
                manifest["md5"]["sam_%s.sam" % tag] = md5(sam)
        for f in ("g.fa", "r.fq", "g.fa.ann"):
            manifest["md5"][f] = md5(os.path.join(HERE, f))
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
    add_precalc(ref, manifest)
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
    print("golden files written to", HERE)


if __name__ == "__main__":
    main()
