#!/bin/bash
for o in 0.75 0.8; do for w in "0.7,1.4" "0.75,1.3" "0.8,1.2" "0.7,1.5" "0.65,1.4"; do
echo "== omega0 $o wc $w: $(KEFFB200_BATCH_OMEGA=$o KEFFB200_BATCH_WC=$w timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1)"; done; done
