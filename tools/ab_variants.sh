cd $GRAFT_REPO_ROOT
P=$PWD/covid19-vaccination-model_b200
run() { n=$1; shift; VAXMC_LIB=$P/libvaxmc_$n.so VAXMC_NVCC_EXTRA="$*" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --target-samples 0 --grid-boxes 0 > gpurun_out/r2h_$n.log 2>&1; python - $n <<'PY'
import json,sys
n=sys.argv[1]
for l in open(f"gpurun_out/r2h_{n}.log"):
    if l.startswith("{"):
        d=json.loads(l); print(n, round(d["ms_per_step"],4), round(d["roofline"]["ms_per_launch"],4), d["breakdown_ms_per_step"]["pilot"], d["bitexact"]["sha256"][:12])
PY
}
run base -DVX_ZCLAMP
run e1
run e8 -DVX_SMEM_STATE
run e7 -DVX_SMEM_STATE -DVX_HOIST_CF
run e7b -DVX_HOIST_CF
run base -DVX_ZCLAMP
run e7 -DVX_SMEM_STATE -DVX_HOIST_CF
