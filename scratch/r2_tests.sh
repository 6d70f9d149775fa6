#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/gpu_tests.log
