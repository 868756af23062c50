#!/bin/bash
# Evidence run on the GPU box (one gpurun call): ncu launch list of the bench command, one full capture of a warm
# step's three kernels, warm-cache DRAM / L2 traffic of the same kernels, compute-sanitizer memcheck + racecheck on
# small shapes. Everything lands in gpurun_out/ (summaries are copied into profiles/ by hand).
set -x
R=${1:-r02}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches.csv \
  python bench.py --steps 16 --warmup 3 > gpurun_out/${R}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"amil_tile2|amil_hidden|gemm2_tc" -s 6 -c 6 \
  -f -o gpurun_out/${R}_prof_step python tools/prof_step.py > gpurun_out/${R}_ncu_full.log 2>&1
timeout 600 ncu --cache-control none --clock-control none \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum \
  -k regex:"amil_tile2|amil_hidden|gemm2_tc" -s 9 -c 6 --csv --log-file gpurun_out/${R}_warm_traffic.csv \
  env STEPS=6 python tools/prof_step.py > gpurun_out/${R}_ncu_warm.log 2>&1
N=700 L=256 D=256 STEPS=2 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/prof_step.py \
  > gpurun_out/${R}_memcheck_small.log 2>&1; echo "memcheck small rc=$?" >> gpurun_out/${R}_memcheck_small.log
N=1500 L=512 D=384 STEPS=2 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/prof_step.py \
  > gpurun_out/${R}_memcheck_big.log 2>&1; echo "memcheck big rc=$?" >> gpurun_out/${R}_memcheck_big.log
N=700 L=256 D=256 STEPS=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/prof_step.py \
  > gpurun_out/${R}_racecheck_small.log 2>&1; echo "racecheck small rc=$?" >> gpurun_out/${R}_racecheck_small.log
N=600 L=512 D=384 STEPS=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/prof_step.py \
  > gpurun_out/${R}_racecheck_big.log 2>&1; echo "racecheck big rc=$?" >> gpurun_out/${R}_racecheck_big.log
tail -n 4 gpurun_out/${R}_memcheck_*.log gpurun_out/${R}_racecheck_*.log
