#!/bin/bash
timeout 200 python scratch/ab_stencil.py 2>&1 | grep -E "apply|solve|rror|rel err"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mg.py -x -q -m gpu 2>&1 | tail -2
