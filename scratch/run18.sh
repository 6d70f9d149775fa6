#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r1o.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/gpu_tests_r1o.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1o.json 2> gpurun_out/bench_r1o.err; echo "bench rc $?"; cut -c1-300 gpurun_out/bench_r1o.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
