#!/bin/bash
# usage (on the GPU box, from the repo root): bash tools/ncu_capture.sh <tag>
# 1. the bench command without ncu (must exit 0), 2. launch list with per-launch device times,
# 3. ONE `ncu --set full` capture of the streaming-fold stepper (vx_sim_kernel<SIM_FOLD, 0, 0>, picked
# by its mangled name; the second job's launch).  Everything lands in gpurun_out/<tag>_*.
tag=${1:-cap}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --target-samples 0 --grid-boxes 0"
sha256sum covid19-vaccination-model_b200/csrc/vaxmc_kernels.cuh > gpurun_out/${tag}_kernels.sha256
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
    -k regex:vx_sim_kernelILi1ELi0ELi0E -s 1 -c 1 -f \
    -o gpurun_out/${tag}_fold $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out/${tag}_*
