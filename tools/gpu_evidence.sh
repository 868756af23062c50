#!/bin/bash
# Evidence run on the GPU box (one gpurun call): ncu launch list of the bench command, one full capture of a warm
# step's three kernels, warm-cache DRAM / L2 traffic of the same kernels. Everything lands in gpurun_out/ (summaries are copied into profiles/ by hand).
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
# (compute-sanitizer is closed on this pool: rc 86, gpurun_out/r02a_memcheck_*.log — not attempted again)
