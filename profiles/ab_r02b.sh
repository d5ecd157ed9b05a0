#!/bin/bash
# Round-2 (second session) A/B of K4 builds on one B200 box.  Libraries under bwbble_b200/ab/ (built by hand, see
# profiles/r02_ab_log.md): BASE = commit a6636a1; A3N = compact bucket heads; A3 = + single-interval tail levels in
# registers; A3H = + upper index block requested before classification; A4 / A4H = the same at 4 blocks/SM (128 regs).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/r02b_ab.log
: > $OUT
echo "== parity (default build = A3)" | tee -a $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4 | tee -a $OUT
for v in BASE A3N A3 A3H A4 A4H; do
  echo "== chr21 $v" | tee -a $OUT
  BWBBLE_B200_LIB=$PWD/bwbble_b200/ab/lib_$v.so timeout 300 python bench.py --quick --batch 2097152 --steps 2 --warmup 1 2>>gpurun_out/r02b_ab.err | cut -c1-400 | tee -a $OUT
done
for v in A3 A4 A4H; do
  echo "== g300 $v" | tee -a $OUT
  BWBBLE_B200_LIB=$PWD/bwbble_b200/ab/lib_$v.so timeout 400 python bench.py --quick --workload g300 --steps 2 --warmup 1 2>>gpurun_out/r02b_ab.err | cut -c1-400 | tee -a $OUT
done
