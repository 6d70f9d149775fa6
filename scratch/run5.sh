#!/bin/bash
timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1
KEFFB200_L2_PERSIST=1 timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -2
timeout 200 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum -k regex:k_batch2d_mg -c 1 -s 1 python scratch/prof_batch.py 1480 2>&1 | grep -E "dram__|gpu__time"
KEFFB200_L2_PERSIST=1 timeout 200 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum -k regex:k_batch2d_mg -c 1 -s 1 python scratch/prof_batch.py 1480 2>&1 | grep -E "dram__|gpu__time|persist"
