#!/bin/bash
# batch-kernel development cycle: batch parity tests, then the A/B script (multigrid kernel against the additive one)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "batch or image or onchip or direct or keff2d" > gpurun_out/gpu_tests_batch.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests_batch.log
timeout 300 python scratch/ab_batch.py 10000 2>&1 | grep -v "^$" | tail -12
