#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_mg.py -x -q 2>&1 | tail -2
for om in "0.9" "0.595,1.614" "0.9,0.9" "0.526,0.8,1.665"; do
  echo "== omega $om"; KEFFB200_MG_OMEGA=$om timeout 200 python scratch/mg_check.py 128 512 2>&1 | grep "mg:" 
done
