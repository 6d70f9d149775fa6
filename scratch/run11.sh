#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r1l.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/gpu_tests_r1l.log
timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_batch2d_mg -c 1 -s 1 -o gpurun_out/prof_bmg_r1l -f python scratch/prof_batch.py 10000 > gpurun_out/ncu_bmg_r1l.log 2>&1; echo "ncu full rc $?"
