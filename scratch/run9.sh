#!/bin/bash
for d in 1 0; do echo "== deep $d"; KEFFB200_BATCH_DEEP=$d timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1; done
