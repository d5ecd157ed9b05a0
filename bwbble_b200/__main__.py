"""Command line of the Python mirror, with the reference's sub-commands and flags (mg-aligner/main.c:60-160):

    python -m bwbble_b200 index [--device] <seq_fasta>
    python -m bwbble_b200 align [-M -O -E -n -k -o -e -l -m -t -S -P] [--gpus N] [--sam out.sam] <seq_fasta> <reads_fastq> <output_aln>

`index` writes <fasta>.bwt / .ann (host builder, or the device builder K7 with --device); `align` streams the FASTQ through
the device path and writes the binary .aln the reference's `aln2sam` reads (and, with --sam, the SAM itself).  The drop-in
for the reference's own C host is the shim; this entry point needs no reference code at all."""
from __future__ import annotations

import argparse
import sys

from .params import default_params

_INT_FLAGS = {"M": "mismatch penalty", "O": "gap open penalty", "E": "gap extend penalty", "n": "max differences",
              "k": "max differences in the seed", "o": "max gap opens", "e": "max gap extensions", "l": "seed length",
              "m": "max heap entries", "t": "threads (ignored by the device path)"}


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="python -m bwbble_b200", description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    ix = sub.add_parser("index", help="bwbble index")
    ix.add_argument("--device", action="store_true", help="suffix sort and BWT passes on the GPU (K7)")
    ix.add_argument("--ref", action="store_true", help="also write <fasta>.ref")
    ix.add_argument("fasta")
    al = sub.add_parser("align", help="bwbble align")
    for f, h in _INT_FLAGS.items():
        al.add_argument("-" + f, type=int, default=None, help=h)
    al.add_argument("-S", action="store_true", help="single-genome mode")
    al.add_argument("-P", action="store_true", help="seed the search from the 12-mer table <fasta>.pre")
    al.add_argument("--gpus", type=int, default=1, help="number of devices to shard the reads over")
    al.add_argument("--sam", default=None, help="also write this SAM file (aln2sam -n <max differences + 1 ... 6>)")
    al.add_argument("--sam-max-diff", type=int, default=6, help="aln2sam's -n (default 6, main.c:142)")
    al.add_argument("--batch", type=int, default=0, help="reads per launch (0: 8 M)")
    al.add_argument("fasta")
    al.add_argument("fastq")
    al.add_argument("aln")
    return ap


def params_from_args(args) -> "Params":  # noqa: F821
    kw = {f: getattr(args, f) for f in _INT_FLAGS if getattr(args, f) is not None}
    p = default_params(**kw)
    if args.S:
        p.is_multiref = 0
    if args.P:
        p.use_precalc = 1
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    if args.cmd == "index":
        from .index import build_index
        if args.device:
            from .align import Aligner
            with Aligner([0]) as al:
                build_index(args.fasta, write_ref=args.ref, aligner=al)
        else:
            build_index(args.fasta, write_ref=args.ref)
        print("index written: %s.bwt" % args.fasta)
        return 0
    from .align import align_reads
    n = align_reads(args.fasta, args.fastq, args.aln, params_from_args(args), devices=list(range(args.gpus)),
                    batch=args.batch, sam_path=args.sam, max_mm=args.sam_max_diff)
    print("Processed %d reads." % n)
    return 0


if __name__ == "__main__":
    sys.exit(main())
