#!/bin/bash
# streamed host batches: batch tests + bench (e2e pinned / pageable), against the three-piece scheme
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "batch or image or onchip or direct or fp32 or keff2d" 2>&1 | tail -3
for v in "" "KEFFB200_BATCH_PIECES=1"; do
  echo "== $v"
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-3d 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'pageable',round(d['e2e']['pageable_value']),'launches',d['gpu_launches'], 'e2e ms', d['e2e']['ms_per_step'])"
done
