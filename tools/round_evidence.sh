#!/bin/bash
# usage (GPU box, repo root): bash tools/round_evidence.sh <tag>
# the single-GPU evidence of a round: GPU test suite, the default bench line, the reference arm,
# small-job latencies, launch list + one full ncu capture of the fold stepper
tag=${1:-ev}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -2 gpurun_out/${tag}_tests.log
python bench.py > gpurun_out/${tag}_bench_n1.log 2>&1; tail -c 600 gpurun_out/${tag}_bench_n1.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.log 2>&1; tail -c 300 gpurun_out/${tag}_bench_ref.log
python tools/latency_small_jobs.py > gpurun_out/${tag}_latency.jsonl 2>&1; tail -2 gpurun_out/${tag}_latency.jsonl
bash tools/ncu_capture.sh ${tag}
