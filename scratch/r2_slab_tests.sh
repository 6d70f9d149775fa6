#!/bin/bash
# the z-slab tests (need 2 GPUs): both transports against the single-GPU solve, one process driving all GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 400 python -m pytest tests/test_gpu_slab.py tests/test_host_cpp.py -m gpu -q -x 2>&1 | tail -3
