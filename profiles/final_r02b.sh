#!/bin/bash
# Round 2, final measurements of the second session (one gpurun call): default bench line, ncu launch list, ncu --set full of K4.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02b_bench_chr21.json 2> gpurun_out/r02b_bench_chr21.log
tail -c 600 gpurun_out/r02b_bench_chr21.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv \
    python bench.py --quick --steps 2 --warmup 1 --batch 1048576 > gpurun_out/r02b_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_search_l -c 1 -f -o gpurun_out/r02b_k4 \
    python bench.py --quick --steps 1 --warmup 0 --batch 393216 > gpurun_out/r02b_k4_ncu.log 2>&1
ls -la gpurun_out | tail -8
