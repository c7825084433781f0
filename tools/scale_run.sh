#!/bin/bash
# usage (GPU box, repo root): bash tools/scale_run.sh <tag> <N> [extra]
# one torchrun bench line at N ranks (+ at N = 8 the wide 5-year sweep and the NCCL bit-exactness script)
tag=$1; N=$2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node $N --master-port 29700 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${tag}_bench_n$N.log 2>&1; echo rc=$?
grep -c '^{' gpurun_out/${tag}_bench_n$N.log
if [ "$3" = "full" ]; then
  $TR --nproc-per-node $N --master-port 29701 bench.py --gpus $N --workload wide5y --samples $((100000000/N)) --steps 3 --warmup 1 --target-samples 0 --grid-boxes 0 > gpurun_out/${tag}_bench_n${N}_wide.log 2>&1; echo rc=$?
  $TR --nproc-per-node $N --master-port 29702 tests/multigpu_bitexact.py > gpurun_out/${tag}_bitexact_n$N.log 2>&1; echo rc=$?
  tail -3 gpurun_out/${tag}_bitexact_n$N.log | cut -c1-200
fi
