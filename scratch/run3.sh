#!/bin/bash
timeout 200 python scratch/ab_batch.py 1480 2>&1 | grep -E "^mg|^add|gap|\(128, 64\)|\(16, 16\)"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r1i.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/gpu_tests_r1i.log
timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1
