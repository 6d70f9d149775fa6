#!/bin/bash
timeout 120 python scratch/dbg_batch.py 10000 10000 0 2>&1 | tail -2
timeout 120 python scratch/dbg_batch.py 0 10000 0 2>&1 | tail -1
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scratch/dbg_batch.py 10000 10000 0 > gpurun_out/sanitizer_r1k.log 2>&1; echo "sanitizer rc $?"; grep -v "^$" gpurun_out/sanitizer_r1k.log | head -40
