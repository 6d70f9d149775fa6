#!/bin/bash
# round-2: slab transports only (needs >= 2 GPUs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/slab_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/slab_tests.log
for n in 128 256; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tests/helpers/slab_check.py $n > gpurun_out/slab_p2p_$n.log 2>&1; echo "slab p2p $n rc=$?"
KEFFB200_SLAB_NOFUSE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 tests/helpers/slab_check.py $n > gpurun_out/slab_nofuse_$n.log 2>&1; echo "slab nofuse $n rc=$?"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/benchN.json 2> gpurun_out/benchN.err; echo "benchN rc=$?"
grep -h "transport\|slab \|single" gpurun_out/slab_p2p_*.log gpurun_out/slab_nofuse_*.log
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/benchN.json').read().strip().splitlines()[-1])
    for k in ('batch_split','slab_parity','slab3d','slab3d_strong','solve3d'):
        if k in d: print(k, {a:b for a,b in d[k].items() if a not in ('workload','scaling','solver')})
except Exception as e:
    print('bench parse', e)
PY
grep -i "error\|illegal" gpurun_out/benchN.err | head -5
