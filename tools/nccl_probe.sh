#!/bin/bash
# usage (GPU box, N GPUs): bash tools/nccl_probe.sh <tag> <N>: the all-reduce of the slab under NCCL's choices
tag=$1; N=$2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
B="bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --target-samples 0 --grid-boxes 0"
run() { name=$1; shift; env "$@" $TR --master-port 29710 $B > gpurun_out/${tag}_$name.log 2>&1; python - $name gpurun_out/${tag}_$name.log <<'PY'
import json,sys
for l in open(sys.argv[2]):
    if l.startswith("{"):
        d=json.loads(l); b=d["breakdown_ms_per_step"]; print(sys.argv[1], "ms %.3f merge %.3f fold %.3f pilot %.3f" % (d["ms_per_step"], b["merge_allreduce"], b["fold"], b["pilot"]))
PY
}
run default X=1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING $TR --master-port 29711 bench.py --gpus $N --steps 2 --warmup 1 --no-e2e --target-samples 0 --grid-boxes 0 > gpurun_out/${tag}_debug.log 2>&1
grep -iE "nvls|algo|Connected|channels|comm 0x.*rank 0" gpurun_out/${tag}_debug.log | head -30
run nvls NCCL_ALGO=NVLS
run tree NCCL_ALGO=Tree
run ring_ll128 NCCL_ALGO=Ring NCCL_PROTO=LL128
run ring_simple NCCL_ALGO=Ring NCCL_PROTO=Simple
