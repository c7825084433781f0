#!/bin/bash
# usage (GPU box): bash tools/sweep_opts.sh <tag> "<VX_OPTIONS 1>" "<VX_OPTIONS 2>" ...   (use "-" for none)
cd $GRAFT_REPO_ROOT
tag=$1; shift
i=0
for o in "$@"; do
i=$((i+1)); [ "$o" = "-" ] && o=""
VX_OPTIONS="$o" timeout 300 python bench.py --steps ${STEPS:-15} --warmup 3 --no-cpu-baseline --no-e2e --target-samples 0 --grid-boxes 0 > gpurun_out/${tag}_$i.log 2>&1
python - "$o" gpurun_out/${tag}_$i.log <<'PY'
import json,sys
o,f=sys.argv[1:3]
ok=False
for l in open(f):
    if l.startswith("{"):
        d=json.loads(l); b=d["breakdown_ms_per_step"]; ok=True
        print("%-50s total %.4f pilot %.4f fold %.4f nonfold %.4f  side %s  %s" % (o, d["ms_per_step"], b["pilot"], b["fold"], b["non_fold"], d["job_split"]["side_samples_stepped_beside_the_pilot"], d["bitexact"]["sha256"][:10]))
if not ok: print(o, "FAILED", open(f).read()[-400:])
PY
done
