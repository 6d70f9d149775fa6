#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch" 2>&1 | tail -2
timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1
for w in "0.595,1.614" "0.85,0.85" "0.65,1.5"; do echo "== wc $w"; KEFFB200_BATCH_WC=$w timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1; done
for o in 0.8 0.9; do echo "== omega $o"; KEFFB200_BATCH_OMEGA=$o timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1; done
echo "== delta 1e-5"; KEFFB200_BATCH_DELTA=1e-5 timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1
timeout 200 python scratch/ab_batch.py 1480 2>&1 | grep -E "iters mg|^mg"
