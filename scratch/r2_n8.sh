#!/bin/bash
# round-2: the z-slab transports and the bench line on all GPUs of the box (8)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tests/helpers/slab_check.py 128 > gpurun_out/slab_p2p_128_n$N.log 2>&1; echo "slab p2p 128 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 tests/helpers/slab_check.py 256 > gpurun_out/slab_p2p_256_n$N.log 2>&1; echo "slab p2p 256 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n${N}_r2a.json 2> gpurun_out/bench_n${N}_r2a.err; echo "benchN rc=$?"
grep -h "transport\|slab \|single" gpurun_out/slab_p2p_*_n$N.log
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_r2a.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'])
    for k in ('batch_split','slab_parity','slab3d','slab3d_strong','solve3d'):
        if k in d: print(k, {a:b for a,b in d[k].items() if a not in ('workload','scaling','solver')})
except Exception as e:
    print('bench parse', e)
PY
grep -i "error\|illegal\|Traceback" gpurun_out/bench_n${N}_r2a.err | head -5
