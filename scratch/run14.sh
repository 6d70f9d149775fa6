#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch" 2>&1 | tail -2
timeout 100 python scratch/prof_batch.py 10000 2>&1 | tail -1
echo "== 3D default"; timeout 200 python scratch/prof_stencil.py 512 500000 2>&1 | tail -1
echo "== 3D nu0=1 w 0.8"; KEFFB200_MG_OMEGA0=0.8 timeout 200 python scratch/prof_stencil.py 512 500000 2>&1 | tail -1
echo "== 3D nu0=1 w 0.85"; KEFFB200_MG_OMEGA0=0.85 timeout 200 python scratch/prof_stencil.py 512 500000 2>&1 | tail -1
echo "== 3D nu0=1 w 0.8, coarse .7,1.4"; KEFFB200_MG_OMEGA0=0.8 KEFFB200_MG_OMEGA=0.7,1.4 timeout 200 python scratch/prof_stencil.py 512 500000 2>&1 | tail -1
