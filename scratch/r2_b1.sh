#!/bin/bash
# round-2 single-GPU cycle: GPU tests, batch A/B over the reliable-update window, short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests.log
for d in 1e-4 1e-5 1e-6 1e-8; do
  echo "== DELTA $d"; KEFFB200_BATCH_DELTA=$d timeout 300 python scratch/ab_batch.py 10000 2>&1 | grep "^mg:\|worst"
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
for k in ('stencil3d','solve3d','solve3d_1024'):
    if k in d: print(k, {a:b for a,b in d[k].items() if a not in ('workload','scaling','solver','kernel')})
PY
