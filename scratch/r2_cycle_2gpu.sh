#!/bin/bash
# round-2 3D cycle on 2 GPUs: all GPU tests (incl. slab), short bench at N=1 (3D legs) with p FP32 vs KEFFB200_P64
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
for k in ('stencil3d','stencil3d_f32in','solve3d','solve3d_1024'):
    if k in d: print(k, {a:b for a,b in d[k].items() if a not in ('workload','scaling','solver','kernel','bound','unit','peak')})
PY
KEFFB200_P64=1 timeout 300 python scratch/prof_mg.py 512 500000 2>&1 | tail -3
timeout 300 python scratch/prof_mg.py 512 500000 2>&1 | tail -3
