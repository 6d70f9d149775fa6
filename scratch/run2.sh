#!/bin/bash
timeout 200 python scratch/ab_batch.py 1480 2>&1 | grep -E "^mg|^add|gap|\(128, 64\)|\(16, 16\)"
for d in 1e-5 1e-6 1e-8; do echo "== delta $d"; KEFFB200_BATCH_DELTA=$d timeout 100 python scratch/prof_batch.py 1480 2>&1 | tail -1; done
for o in 0.8 0.9; do echo "== omega $o"; KEFFB200_BATCH_OMEGA=$o timeout 100 python scratch/prof_batch.py 1480 2>&1 | tail -1; done
