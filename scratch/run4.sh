#!/bin/bash
timeout 200 python scratch/ab_batch.py 1480 2>&1 | grep -E "^mg|^add|gap"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r1j.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/gpu_tests_r1j.log
timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1
KEFFB200_BATCH_DELTA=1e-5 timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1
