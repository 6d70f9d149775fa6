#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch" 2>&1 | tail -2
timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1
KEFFB200_BATCH_DELTA=1e-5 timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1
