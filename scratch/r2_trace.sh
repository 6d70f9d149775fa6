#!/bin/bash
# phase trace of iteration 3: 1024x1024x128 per GPU in N slabs (N = all GPUs), beside 512^3 on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
KEFFB200_TRACE=3 timeout 150 python scratch/prof_mg.py 512 500000 > gpurun_out/trace_1gpu.log 2>&1; echo "1gpu rc=$?"; grep "trace\|solve" gpurun_out/trace_1gpu.log | tail -3
if [ "$N" -gt 1 ]; then
KEFFB200_TRACE=3 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 tests/helpers/slab_check.py 1024 $((128*N)) > gpurun_out/trace_n$N.log 2>&1; echo "slab rc=$?"
grep "trace\|slab " gpurun_out/trace_n$N.log | tail -20
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29722 tests/helpers/slab_check.py 256 > gpurun_out/slab256_n$N.log 2>&1; echo "slab256 rc=$?"
grep "slab \|single" gpurun_out/slab256_n$N.log
fi
