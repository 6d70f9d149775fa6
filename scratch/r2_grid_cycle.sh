#!/bin/bash
# grid-solver development cycle on one GPU: the 3D parity tests, then the 512^3 solve with its phase trace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python -m pytest tests -m gpu -q -x -k "mg or schwarz or bernoulli or ragged or volume or fp32 or known or operator" 2>&1 | tail -2
KEFFB200_TRACE=3 timeout 200 python scratch/prof_mg.py 512 500000 2>&1 | grep "trace\|solve" | sed 's/\[keffb200 trace\] rank 0 iteration 3: //' | tail -2
