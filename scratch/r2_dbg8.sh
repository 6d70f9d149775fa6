#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 scratch/dbg_slabT.py 1 2 > gpurun_out/dbg_slabT_pre.log 2>&1; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|NCCL version" gpurun_out/dbg_slabT_pre.log | tail -20
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 scratch/dbg_slabT.py 0 2 > gpurun_out/dbg_slabT_nopre.log 2>&1; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|NCCL version" gpurun_out/dbg_slabT_nopre.log | tail -20
