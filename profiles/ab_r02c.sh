#!/bin/bash
# Round-2 A/B, third call: cooperative index-block loads (COOP) in K4.  A4H = the default build (4 blocks/SM, early
# request of the upper block); C4 = COOP at 4 blocks/SM; C3 = COOP at 3 blocks/SM (168 registers, no spills).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
OUT=gpurun_out/r02c_ab.log
: > $OUT
for v in A4H C4 C3; do
  echo "== chr21 $v" | tee -a $OUT
  BWBBLE_B200_LIB=$PWD/bwbble_b200/ab/lib_$v.so timeout 300 python bench.py --quick --batch 2097152 --steps 2 --warmup 1 2>>gpurun_out/r02c_ab.err | cut -c1-330 | tee -a $OUT
done
echo "== parity C4" | tee -a $OUT
BWBBLE_B200_LIB=$PWD/bwbble_b200/ab/lib_C4.so timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_short_reads.py -x -q 2>&1 | tail -4 | tee -a $OUT
