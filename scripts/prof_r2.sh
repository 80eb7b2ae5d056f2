# round-2 profile captures (run under gpurun, one GPU)
set -x
# 1. every launch of the bench command with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-parity --no-partition --no-train > gpurun_out/r2_launches_bench.log 2>&1
# 2. full sections of every kernel of ONE eager step at the bench batch (the second of two steps: 17 launches per step;
#    the static embedders of the first step run on the exact-fp32 kernels, which the name filter leaves out)
ncu --set full --clock-control none --import-source on -k regex:"tc_|step_epilogue|halo" -s 17 -c 17 -o gpurun_out/prof_r2_step \
    python scripts/prof_step.py 32 2 > gpurun_out/prof_r2_step.log 2>&1
ls -la gpurun_out/*.ncu-rep
