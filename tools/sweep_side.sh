cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streamed or sharding or refinement or grid or usa_1e6 or wide or max_running or explicit or quantiles_fan" > gpurun_out/r2k_tests.log 2>&1; tail -3 gpurun_out/r2k_tests.log
for s in 0 32768 98304 163840 229376; do
VX_OPTIONS=side_samples=$s python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --target-samples 0 --grid-boxes 0 > gpurun_out/r2k_side_$s.log 2>&1
python - $s <<'PY'
import json,sys
n=sys.argv[1]
for l in open(f"gpurun_out/r2k_side_{n}.log"):
    if l.startswith("{"):
        d=json.loads(l); print(n, round(d["ms_per_step"],4), d["breakdown_ms_per_step"], d["bitexact"]["sha256"][:12])
PY
done
