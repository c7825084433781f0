#!/bin/bash
# A/B of compile-time variants of libvaxmc on the GPU box.
#   here:    for v in base:"" x:"-DVX_HITS_PER_QUAD"; do n=${v%%:*}; f=${v#*:}; \
#              VAXMC_LIB=$PWD/covid19-vaccination-model_b200/libvaxmc_$n.so VAXMC_NVCC_EXTRA="$f" \
#              python covid19-vaccination-model_b200/build.py; done       (the .so files travel with gpurun)
#   there:   bash tools/ab_variants.sh <tag> base:"" x:"-DVX_HITS_PER_QUAD"
# Prints ms per job, ms per fold launch, pilot ms and the digest of the fixed 3e6-sample job per variant.
cd ${GRAFT_REPO_ROOT:-.}
P=$PWD/covid19-vaccination-model_b200
tag=$1; shift
for v in "$@"; do
  n=${v%%:*}; f=${v#*:}
  VAXMC_LIB=$P/libvaxmc_$n.so VAXMC_NVCC_EXTRA="$f" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e \
      --target-samples 0 --grid-boxes 0 > gpurun_out/${tag}_$n.log 2>&1
  python - $n gpurun_out/${tag}_$n.log <<'PY'
import json, sys
n, f = sys.argv[1:3]
for l in open(f):
    if l.startswith("{"):
        d = json.loads(l)
        print(n, round(d["ms_per_step"], 4), round(d["roofline"]["ms_per_launch"], 4), round(d["breakdown_ms_per_step"]["pilot"], 4),
              d["bitexact"]["sha256"][:12])
PY
done
