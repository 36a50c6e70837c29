#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, one GPU): launch lists of c4 and c5 and one full capture of each dominant kernel.
# Every ncu command runs only after the identical plain command exited 0 (B200_PROFILING.md).
mkdir -p gpurun_out
K4='regex:zq_sweep|pre_sweep|post_sweep|epilogue|p_dirichlet|moments'
K5='regex:tetra|p_dirichlet|moments'
B4="python bench.py --workload c4 --steps 12 --warmup 3 --no-cpu --no-e2e"
B5="python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu --no-e2e"
$B4 > gpurun_out/r2_plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$K4" -s 30 -c 50 --csv --log-file gpurun_out/r2_launches_c4.csv $B4 > /dev/null 2>&1
$B4 > gpurun_out/r2_plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:zq_sweep -s 4 -c 1 -f -o gpurun_out/r2_zq_sweep_final $B4 > gpurun_out/r2_ncu_c4.log 2>&1
$B5 > gpurun_out/r2_plain_c5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$K5" -s 40 -c 60 --csv --log-file gpurun_out/r2_launches_c5.csv $B5 > /dev/null 2>&1
$B5 > gpurun_out/r2_plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tetra_zs|tetra_geno" -s 6 -c 2 -f -o gpurun_out/r2_tetra_final $B5 > gpurun_out/r2_ncu_c5.log 2>&1
ls -la gpurun_out/r2_zq_sweep_final.ncu-rep gpurun_out/r2_tetra_final.ncu-rep gpurun_out/r2_launches_c4.csv gpurun_out/r2_launches_c5.csv
