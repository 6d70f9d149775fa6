#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_rNN.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/gpu_tests_rNN.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_rNN.json 2> gpurun_out/bench_rNN.err; echo "bench rc $?"; cut -c1-300 gpurun_out/bench_rNN.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
