#!/bin/bash
# the bench line on all GPUs of the box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TAG=${1:-r2b}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "benchN rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_$TAG.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'])
    for k in ('batch_split','slab_parity','slab3d','slab3d_strong','solve3d'):
        if k in d: print(k, {a:b for a,b in d[k].items() if a not in ('workload','scaling','solver')})
except Exception as e:
    print('bench parse', e)
PY
grep -i "error\|illegal\|Traceback" gpurun_out/bench_n${N}_$TAG.err | head -5
