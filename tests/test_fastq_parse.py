"""The native FASTQ reader (bwb_fastq_parse / bwb_align_fastq, fastq_stream.cpp) against fastq2reads' record grammar
(io.c:410-515), restated here line by line: records start at the next '@' wherever it is; the name is the rest of
that line (256 characters kept); the base line is mapped through nt4_table (io.h:113-130); everything up to and
including the '+' line is skipped; the quality line must be as long as the base line and may end with the file.
Host-only: no device needed."""
import numpy as np
import pytest

from bwbble_b200 import BwbError
from bwbble_b200.fastx import read_fastq

NT4 = {c: v for cs, v in (("Aa", 0), ("Gg", 1), ("Cc", 2), ("Tt", 3)) for c in cs}


def grammar(data: bytes):
    """byte-at-a-time restatement (what the round-1 parser and the reference's fgetc loop do)"""
    i, n = 0, len(data)
    reads = []

    def until(ch):
        nonlocal i
        out = bytearray()
        while i < n and data[i] != ch:
            out.append(data[i])
            i += 1
        ok = i < n
        i += 1
        return bytes(out), ok

    while True:
        _, ok = until(ord("@"))
        if not ok:
            return reads
        name, ok = until(10)
        assert ok
        bases, ok = until(10)
        assert ok
        _, ok = until(ord("+"))
        assert ok
        _, ok = until(10)
        assert ok
        qual, _ = until(10)          # may end with the file
        if len(qual) != len(bases):
            raise ValueError("quality length")
        reads.append((name[:256], [NT4.get(chr(b), 4) for b in bases], qual))


CASES = {
    "plain": b"@r1\nACGT\n+\n2222\n@r2 desc\nNNAC\n+r2\n!!!!\n",
    "no_trailing_newline": b"@r1\nACGTAC\n+\n222222",
    "garbage_before_and_between": b"junk line\n\n@r1\nAC\n+\n22\nstray text\n\n@r2\nGT\n+\n22\n",
    "lowercase_and_iupac": b"@x\nacgtRYKMn\n+\n222222222\n",
    "crlf": b"@r1\r\nACGT\r\n+\r\n2222\r\n",                    # '\r' is a base like any other (code 4) and counts
    "at_sign_in_quality": b"@r1\nACGT\n+\n@@@@\n@r2\nTTTT\n+\n2222\n",
    "long_name": b"@" + b"n" * 400 + b"\nAC\n+\n22\n",
    "empty_read": b"@e\n\n+\n\n@r\nA\n+\n2\n",
    "many": b"".join(b"@r%d\n%s\n+\n%s\n" % (k, b"ACGTN"[: 1 + k % 5] * (1 + k % 7), b"2" * ((1 + k % 5) * (1 + k % 7))) for k in range(5000)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_native_reader_follows_the_reference_grammar(tmp_path, name):
    data = CASES[name]
    p = tmp_path / "x.fq"
    p.write_bytes(data)
    exp = grammar(data)
    got = read_fastq(str(p), with_quals=True)
    assert got.n == len(exp)
    for k, (nm, codes, qual) in enumerate(exp):
        assert list(got.read(k)) == codes, (name, k)
        assert got.names[k].encode() == nm
        assert got.meta["quals"][k].encode() == qual


def test_quality_length_mismatch_is_an_error(tmp_path):
    p = tmp_path / "bad.fq"
    p.write_bytes(b"@r1\nACGT\n+\n222\n")
    with pytest.raises(BwbError):
        read_fastq(str(p))
    p.write_bytes(b"@r1\nACGT\n")                 # truncated record
    with pytest.raises(BwbError):
        read_fastq(str(p))
    with pytest.raises(BwbError):
        read_fastq(str(tmp_path / "missing.fq"))


def test_records_longer_than_the_read_buffer(tmp_path):
    """a 9 MB base line spans several 4 MB reader blocks"""
    n = 9 * (1 << 20) + 123
    rng = np.random.default_rng(1)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].tobytes()
    p = tmp_path / "big.fq"
    p.write_bytes(b"@big\n" + bases + b"\n+\n" + b"2" * n + b"\n@s\nAC\n+\n22\n")
    got = read_fastq(str(p))
    assert got.n == 2 and len(got.read(0)) == n and list(got.read(1)) == [0, 2]
    exp = np.frombuffer(bases, dtype=np.uint8)
    lut = np.full(256, 4, dtype=np.uint8)
    for c, v in NT4.items():
        lut[ord(c)] = v
    assert (got.read(0) == lut[exp]).all()
