#!/bin/bash
# GPU call 1: A/B of the stencil kernels, full GPU tests, bench (both arms), ncu of the bench's batch kernel
mkdir -p gpurun_out
timeout 300 python scratch/ab_stencil.py > gpurun_out/ab_ws.log 2>&1; echo "ab ws rc $?"
KEFFB200_STENCIL_SYNC=1 timeout 300 python scratch/ab_stencil.py > gpurun_out/ab_sync.log 2>&1; echo "ab sync rc $?"
grep -E "apply|solve|Error|error" gpurun_out/ab_ws.log gpurun_out/ab_sync.log | tail -12
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r1h.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/gpu_tests_r1h.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; echo "bench rc $?"; cat gpurun_out/bench_r1h.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1h.json 2> gpurun_out/bench_ref_r1h.err; echo "ref rc $?"; cat gpurun_out/bench_ref_r1h.json
timeout 300 ncu --set full --clock-control none -k regex:k_batch2d_onchip -c 1 -s 1 -o gpurun_out/prof_batch_r1h -f python scratch/prof_batch.py 10000 > gpurun_out/ncu_batch_r1h.log 2>&1; echo "ncu rc $?"
