#!/bin/bash
# round-2 GPU check: tests, slab transports, short bench (1 GPU and all GPUs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
N=$(nvidia-smi -L | wc -l)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench1 rc=$?"
if [ "$N" -ge 2 ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tests/helpers/slab_check.py 128 > gpurun_out/slab_p2p.log 2>&1; echo "slab p2p rc=$?"
  KEFFB200_SLAB_NCCL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 tests/helpers/slab_check.py 128 > gpurun_out/slab_nccl.log 2>&1; echo "slab nccl rc=$?"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/benchN.json 2> gpurun_out/benchN.err; echo "benchN rc=$?"
  KEFFB200_SLAB_NCCL=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/benchN_nccl.json 2> gpurun_out/benchN_nccl.err; echo "benchN nccl rc=$?"
fi
grep -h "transport\|slab \|single" gpurun_out/slab_*.log
